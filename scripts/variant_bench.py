"""Times ECDSA verify (device-resident) for each prebuilt variant .so; one subprocess per variant."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import importlib, os, sys, json, numpy as np, torch
sys.path.insert(0, %r)
pkg = importlib.import_module("secp256k1-voi_b200")
n = 1 << int(os.environ.get("LOG2N", "19"))
eng = pkg.Engine(device=0, max_batch=n)
w = pkg.synth.ecdsa_batch(n, eng.scalar_base_mult)
d = [torch.from_numpy(w[k]).cuda() for k in ("pk65", "digest32", "sig64")]
for _ in range(3): ok = eng.ecdsa_verify(*d)
torch.cuda.synchronize()
good = bool(np.array_equal(ok.cpu().numpy(), w["expected"]))
eng.profile_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ok = eng.ecdsa_verify(*d)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
dsm_ms, cnt = eng.profile_read()
ks = torch.from_numpy(pkg.synth.base_mult_scalars(n)).cuda()
for _ in range(2): eng.scalar_base_mult(ks)
torch.cuda.synchronize(); e0.record()
for _ in range(3): eng.scalar_base_mult(ks)
e1.record(); torch.cuda.synchronize()
sbm = n * 3 / (e0.elapsed_time(e1) * 1e-3)
print(json.dumps({"ok": good, "verifies_per_s": n / (ms * 1e-3), "ms": ms, "dsm_ms": dsm_ms / max(cnt, 1), "sbm_per_s": sbm}))
''' % ROOT
res = {}
libs = sorted(glob.glob(os.path.join(ROOT, "secp256k1-voi_b200", "lib", "variants", "*.so")))
for lib in libs:
    env = dict(os.environ, S256_LIB=lib)
    p = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
    name = os.path.basename(lib)[:-3]
    try:
        res[name] = json.loads(p.stdout.strip().splitlines()[-1])
    except Exception:
        res[name] = {"error": (p.stderr or p.stdout)[-400:]}
    print(name, res[name], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "variants.json"), "w"), indent=1)
