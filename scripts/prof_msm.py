"""Drives the MSM a few times (for ncu launch lists); TIME=1 prints wall-clock per call (pinned host buffers in)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
pkg = importlib.import_module("secp256k1-voi_b200")
n = 1 << int(os.environ.get("LOG2N", "20"))
eng = pkg.Engine(device=0, max_batch=n)
wm = pkg.synth.msm_batch(n, eng.scalar_base_mult)
hk, hp = torch.from_numpy(wm["k32"]).pin_memory().numpy(), torch.from_numpy(wm["pt65"]).pin_memory().numpy()
for _ in range(3): r, st = eng.msm(hk, hp)
print("ok", st)
if os.environ.get("TIME"):
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); eng.msm(hk, hp); ts.append(time.perf_counter() - t0)
    print({"n": n, "ms_per_msm_min": min(ts) * 1e3, "ms_per_msm_median": sorted(ts)[5] * 1e3})
