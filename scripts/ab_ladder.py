"""A/B: frame-form (shared-memory operand) ladder vs register-form ladder, same library."""
import importlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import importlib, os, sys, json, numpy as np, torch
sys.path.insert(0, %r)
pkg = importlib.import_module("secp256k1-voi_b200")
n = 1 << int(os.environ.get("LOG2N", "20"))
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
print(json.dumps({"ok": good, "verifies_per_s": n / (ms * 1e-3), "ms": ms, "dsm_ms": dsm_ms / max(cnt, 1)}))
''' % ROOT
for mode in ("vm", "reg"):
    env = dict(os.environ, S256_LADDER=mode)
    p = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=900)
    print(mode, p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-500:], flush=True)
