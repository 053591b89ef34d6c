import importlib, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
pkg = importlib.import_module("secp256k1-voi_b200")
eng = pkg.Engine(device=0, max_batch=1 << 20)
out = {}
for n in (256, 4096, 16384, 32768, 65536, 1 << 18, 1 << 20):
    ks = torch.from_numpy(pkg.synth.base_mult_scalars(n)).cuda()
    for _ in range(3): eng.scalar_base_mult(ks)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): eng.scalar_base_mult(ks)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    out[n] = {"ms": ms, "ops_per_s": n / (ms * 1e-3)}
    print(n, out[n], flush=True)
w = pkg.synth.ecdsa_batch(4096, eng.scalar_base_mult)
d = [torch.from_numpy(w[k]).cuda() for k in ("pk65", "digest32", "sig64")]
for _ in range(3): eng.ecdsa_verify(*d)
torch.cuda.synchronize(); a.record()
for _ in range(20): eng.ecdsa_verify(*d)
b.record(); torch.cuda.synchronize()
print("verify 4096", a.elapsed_time(b) / 20, "ms")
