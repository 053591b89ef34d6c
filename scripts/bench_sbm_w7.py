"""scalar_base_mult device-resident at several batch sizes, 6-bit against 7-bit windows (S256_BM_W7_MIN selects)."""
import importlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import importlib, os, sys, json, numpy as np, torch
sys.path.insert(0, %r)
pkg = importlib.import_module("secp256k1-voi_b200")
eng = pkg.Engine(device=0, max_batch=1 << 20)
out = {}
for n in (20000, 40000, 65536, 131072, 262144, 1 << 20):
    ks = torch.from_numpy(pkg.synth.base_mult_scalars(n)).cuda()
    for _ in range(3): r = eng.scalar_base_mult(ks)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): r = eng.scalar_base_mult(ks)
    b.record(); torch.cuda.synchronize()
    out[n] = round(a.elapsed_time(b) / 10, 4)
chk = eng.scalar_base_mult(torch.from_numpy(pkg.synth.base_mult_scalars(200000)).cuda())[0].cpu().numpy()
import hashlib
out["sha"] = hashlib.sha256(chk.tobytes()).hexdigest()[:16]
print(json.dumps(out))
''' % ROOT
for name, v in (("w6", str(1 << 40)), ("w7", "16385")):
    p = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, S256_BM_W7_MIN=v), capture_output=True, text=True, timeout=600)
    print(name, p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-600:], flush=True)
