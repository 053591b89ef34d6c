import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("secp256k1-voi_b200")
eng = pkg.Engine(device=0, max_batch=1 << 17)
for n in (12000, 16384, 20000, 32768, 40000):
    ks = torch.from_numpy(pkg.synth.base_mult_scalars(n)).cuda()
    for _ in range(3): eng.scalar_base_mult(ks)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): eng.scalar_base_mult(ks)
    b.record(); torch.cuda.synchronize()
    print(n, round(a.elapsed_time(b) / 20, 4), "ms")
