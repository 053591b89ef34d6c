"""Per prebuilt variant (.so in lib/variants): ScalarBaseMult at n = 4096 / 2^20 and the one-GPU MSM at 2^17 / 2^20."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import importlib, sys, json, numpy as np, torch
sys.path.insert(0, %r)
pkg = importlib.import_module("secp256k1-voi_b200")
eng = pkg.Engine(device=0, max_batch=1 << 20)
def timed(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
out = {}
for n in (4096, 1 << 20):
    ks = torch.from_numpy(pkg.synth.base_mult_scalars(n)).cuda()
    out["sbm_us_%%d" %% n] = timed(lambda: eng.scalar_base_mult(ks), 30) * 1e3
w = pkg.synth.msm_batch(1 << 20, eng.scalar_base_mult)
for lg in (17, 20):
    dk, dp = torch.from_numpy(w["k32"][:1 << lg].copy()).cuda(), torch.from_numpy(w["pt65"][:1 << lg].copy()).cuda()
    out["msm_ms_2p%%d" %% lg] = timed(lambda: eng.msm(dk, dp), 10)
print(json.dumps(out))
''' % ROOT
for lib in sorted(glob.glob(os.path.join(ROOT, "secp256k1-voi_b200", "lib", "variants", "*.so"))):
    p = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, S256_LIB=lib), capture_output=True, text=True, timeout=600)
    print(os.path.basename(lib)[:-3], p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-400:], flush=True)
