"""ScalarBaseMult at small batch sizes for the lane-split widths (S256_BM_T = lanes per scalar)."""
import importlib, os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import importlib, sys, json, numpy as np, torch
sys.path.insert(0, %r)
pkg = importlib.import_module("secp256k1-voi_b200")
from oracle import oracle as orc
eng = pkg.Engine(device=0, max_batch=1 << 16)
out = {}
for n in (256, 1024, 4096, 8192, 16384):
    ks_h = pkg.synth.base_mult_scalars(n)
    ks = torch.from_numpy(ks_h).cuda()
    got, st = eng.scalar_base_mult(ks)
    exp, est = orc.batch_scalar_base_mult(ks_h)
    ok = bool(np.array_equal(got.cpu().numpy(), exp) and np.array_equal(st.cpu().numpy(), est))
    for _ in range(5): eng.scalar_base_mult(ks)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): eng.scalar_base_mult(ks)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 50
    out[n] = {"ok": ok, "us": ms * 1e3, "Mops": n / ms / 1e3}
print(json.dumps(out))
''' % ROOT
for t in ("0", "16", "32"):
    p = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, S256_BM_T=t), capture_output=True, text=True)
    print("T =", t or "default", p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-500:], flush=True)
