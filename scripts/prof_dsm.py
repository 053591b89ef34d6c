"""Small driver for ncu: a few verify passes at 2^18 items (short kernels)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
pkg = importlib.import_module("secp256k1-voi_b200")
n = 1 << int(os.environ.get("LOG2N", "18"))
eng = pkg.Engine(device=0, max_batch=n)
w = pkg.synth.ecdsa_batch(n, eng.scalar_base_mult)
d = [torch.from_numpy(w[k]).cuda() for k in ("pk65", "digest32", "sig64")]
for _ in range(int(os.environ.get("PASSES", "3"))):
    ok = eng.ecdsa_verify(*d)
torch.cuda.synchronize()
assert np.array_equal(ok.cpu().numpy(), w["expected"])
print("ok", n)
