"""Times the device-resident MSM (n = 2^20 and 2^17) for each prebuilt variant .so; checks the closed form."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import importlib, os, sys, json, numpy as np, torch
sys.path.insert(0, %r)
pkg = importlib.import_module("secp256k1-voi_b200")
out = {}
eng = pkg.Engine(device=0, max_batch=1 << 20)
w = pkg.synth.msm_batch(1 << 20, eng.scalar_base_mult)
exp, _ = eng.scalar_base_mult(np.frombuffer(w["closed_form_scalar"], np.uint8).reshape(1, 32).copy())
for lg in (20, 17):
    n = 1 << lg
    dk, dp = torch.from_numpy(w["k32"][:n]).cuda(), torch.from_numpy(w["pt65"][:n]).cuda()
    for _ in range(3): r = eng.msm(dk, dp)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): r = eng.msm(dk, dp)
    b.record(); torch.cuda.synchronize()
    out[str(lg)] = round(a.elapsed_time(b) / 10, 4)
    if lg == 20:
        got = r[0] if isinstance(r, tuple) else r
        out["ok"] = bool(np.array_equal(np.asarray(got.cpu() if hasattr(got, "cpu") else got).reshape(-1)[:65], np.asarray(exp).reshape(-1)[:65]))
print(json.dumps(out))
''' % ROOT
for lib in sorted(glob.glob(os.path.join(ROOT, "secp256k1-voi_b200", "lib", "variants", "*.so"))):
    p = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, S256_LIB=lib), capture_output=True, text=True, timeout=600)
    print(os.path.basename(lib)[:-3], p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-600:], flush=True)
