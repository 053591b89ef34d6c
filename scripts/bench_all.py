"""Device-resident throughput of every hot-path entry point (BASELINE.json configs 1-5 + recovery),
with the MAC32-based fraction of the live-measured IMAD.WIDE peak.  One JSON object to stdout."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
pkg = importlib.import_module("secp256k1-voi_b200")
LOG2N = int(os.environ.get("LOG2N", "20"))
n = 1 << LOG2N
eng = pkg.Engine(device=0, max_batch=n)
peak = max(eng.microbench_imad(8192)[0] for _ in range(2))
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

out = {"imad_wide_peak_mac32_per_s": peak, "n": n}
def report(name, key, items, ms):
    mac = pkg.mac32_per_item(key)
    out[name] = {"items": items, "ms": ms, "items_per_s": items / (ms * 1e-3), "mac32_per_item": mac,
                 "frac_of_imad_wide_peak": items / (ms * 1e-3) * mac / peak if mac else None}

ks4096 = cu(pkg.synth.base_mult_scalars(4096)); ksn = cu(pkg.synth.base_mult_scalars(n))
report("scalar_base_mult_4096", "scalar_base_mult", 4096, timeit(lambda: eng.scalar_base_mult(ks4096), reps=20))
report("scalar_base_mult", "scalar_base_mult", n, timeit(lambda: eng.scalar_base_mult(ksn)))
w = pkg.synth.ecdsa_batch(n, eng.scalar_base_mult)
d = [cu(w[k]) for k in ("pk65", "digest32", "sig64")]
report("ecdsa_verify", "ecdsa_verify", n, timeit(lambda: eng.ecdsa_verify(*d)))
assert np.array_equal(eng.ecdsa_verify(*d).cpu().numpy(), w["expected"])
sig65 = torch.cat([d[2], torch.zeros((n, 1), dtype=torch.uint8, device="cuda")], dim=1).contiguous()
report("ecdsa_recover", "ecdsa_recover", n, timeit(lambda: eng.ecdsa_recover(d[1], sig65)))
ws = pkg.synth.schnorr_batch(n, eng.scalar_base_mult)
ds = [cu(ws[k]) for k in ("pkx32", "msg", "sig64")]
report("schnorr_verify", "schnorr_verify", n, timeit(lambda: eng.schnorr_verify(*ds)))
assert np.array_equal(eng.schnorr_verify(*ds).cpu().numpy(), ws["expected"])
we = pkg.synth.ecdh_batch(n, eng.scalar_base_mult)
de = [cu(we[k]) for k in ("k32", "pt65")]
report("scalar_mult_ct", "scalar_mult", n, timeit(lambda: eng.scalar_mult(*de)))
report("ecdh", "ecdh", n, timeit(lambda: eng.ecdh(*de)))
got, st = eng.scalar_mult(*de); exp, _ = eng.scalar_base_mult(cu(we["closed_form_scalar"]))
assert torch.equal(got, exp)
u1 = cu(pkg.synth.base_mult_scalars(n, start=7)); 
report("double_scalar_mult", "double_scalar_mult_basepoint_vartime", n, timeit(lambda: eng.double_scalar_mult_basepoint_vartime(u1, de[0], de[1])))
wm = pkg.synth.msm_batch(n, eng.scalar_base_mult)
hk, hp = torch.from_numpy(wm["k32"]).pin_memory().numpy(), torch.from_numpy(wm["pt65"]).pin_memory().numpy()
t0 = time.perf_counter()
for _ in range(3): r, st = eng.msm(hk, hp)
ms = (time.perf_counter() - t0) / 3 * 1e3
exp, est = eng.scalar_base_mult(np.frombuffer(wm["closed_form_scalar"], np.uint8))
assert st == est[0] and np.array_equal(r, exp[0])
out["msm_vartime_host_buffers"] = {"items": n, "ms": ms, "items_per_s": n / (ms * 1e-3)}
print(json.dumps(out, indent=1))
