"""Times the ct ScalarMult kernel for each prebuilt variant .so."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import importlib, os, sys, json, numpy as np, torch
sys.path.insert(0, %r)
pkg = importlib.import_module("secp256k1-voi_b200")
n = 1 << 19
eng = pkg.Engine(device=0, max_batch=n)
we = pkg.synth.ecdh_batch(n, eng.scalar_base_mult)
de = [torch.from_numpy(we[k]).cuda() for k in ("k32", "pt65")]
for _ in range(2): got, st = eng.scalar_mult(*de)
exp, _ = eng.scalar_base_mult(torch.from_numpy(we["closed_form_scalar"]).cuda())
ok = bool(torch.equal(got, exp))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): eng.scalar_mult(*de)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(json.dumps({"ok": ok, "ms": ms, "per_s": n / (ms * 1e-3)}))
''' % ROOT
for lib in sorted(glob.glob(os.path.join(ROOT, "secp256k1-voi_b200", "lib", "variants", "*.so"))):
    env = dict(os.environ, S256_LIB=lib)
    p = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
    print(os.path.basename(lib)[:-3], p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-400:], flush=True)
