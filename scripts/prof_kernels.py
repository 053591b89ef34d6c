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
elif which == "h2c":
    import time
    engp = pkg.Engine(device=0, max_batch=n, pinned_outputs=True)
    msgs = np.frombuffer(b"".join(pkg.synth.D(b"h2c", i) for i in range(n)), np.uint8).reshape(n, 32)
    hm = torch.from_numpy(msgs.copy()).pin_memory().numpy()
    dst = b"QUUX-V01-CS02-with-secp256k1_XMD:SHA-256_SSWU_RO_"
    for ro in (True, False):
        for _ in range(2): engp.hash_to_curve(dst, hm, random_oracle=ro)
        t0 = time.perf_counter()
        for _ in range(3): engp.hash_to_curve(dst, hm, random_oracle=ro)
        print("h2c ro=%s: %.3f ms per 2^%d call, %.1f M/s" % (ro, (time.perf_counter() - t0) / 3 * 1e3, int(np.log2(n)), n * 3 / (time.perf_counter() - t0) / 1e6))
torch.cuda.synchronize()
print("ok")
