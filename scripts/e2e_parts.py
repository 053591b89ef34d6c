"""e2e (host pinned buffers through the C ABI) for different sub-chunk counts."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import importlib, os, sys, json, time, numpy as np, torch
sys.path.insert(0, %r)
pkg = importlib.import_module("secp256k1-voi_b200")
n = 1 << 20
eng = pkg.Engine(device=0, max_batch=n)
w = pkg.synth.ecdsa_batch(n, eng.scalar_base_mult)
h = [torch.from_numpy(w[k]).pin_memory().numpy() for k in ("pk65", "digest32", "sig64")]
for _ in range(2): ok = eng.ecdsa_verify(*h)
assert np.array_equal(ok, w["expected"])
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): ok = eng.ecdsa_verify(*h)
dt = (time.perf_counter() - t0) / 10
print(json.dumps({"ms": dt * 1e3, "verifies_per_s": n / dt}))
''' % ROOT
for parts in (1, 2, 16):
    env = dict(os.environ, S256_PIPE_PARTS=str(parts))
    p = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=900)
    print(parts, p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-400:], flush=True)
