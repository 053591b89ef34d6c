"""Drives each hot kernel a few times (for ncu -k filters)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
pkg = importlib.import_module("secp256k1-voi_b200")
n = 1 << int(os.environ.get("LOG2N", "18"))
which = os.environ.get("WHICH", "verify")
eng = pkg.Engine(device=0, max_batch=n)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
if which == "verify":
    w = pkg.synth.ecdsa_batch(n, eng.scalar_base_mult)
    d = [cu(w[k]) for k in ("pk65", "digest32", "sig64")]
    for _ in range(3): eng.ecdsa_verify(*d)
elif which == "ecdh":
    we = pkg.synth.ecdh_batch(n, eng.scalar_base_mult)
    de = [cu(we[k]) for k in ("k32", "pt65")]
    for _ in range(3): eng.ecdh(*de)
elif which == "sbm":
    ks = cu(pkg.synth.base_mult_scalars(n))
    for _ in range(3): eng.scalar_base_mult(ks)
torch.cuda.synchronize()
print("ok")
