"""One traced host-pointer verify call (S256_TRACE=1 prints the device timeline of the pipeline stages)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
pkg = importlib.import_module("secp256k1-voi_b200")
n = 1 << 20
eng = pkg.Engine(device=0, max_batch=n)
w = pkg.synth.ecdsa_batch(n, eng.scalar_base_mult)
h = [torch.from_numpy(w[k]).pin_memory().numpy() for k in ("pk65", "digest32", "sig64")]
os.environ.pop("S256_TRACE", None)
for _ in range(3): ok = eng.ecdsa_verify(*h)
os.environ["S256_TRACE"] = "1"
ok = eng.ecdsa_verify(*h)
os.environ.pop("S256_TRACE", None)
ts = []
for _ in range(10):
    t0 = time.perf_counter(); eng.ecdsa_verify(*h); ts.append(time.perf_counter() - t0)
print("ms per call: min %.3f median %.3f" % (min(ts) * 1e3, sorted(ts)[5] * 1e3))
