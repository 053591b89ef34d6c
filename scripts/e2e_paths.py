"""e2e (pinned host buffers through the C ABI) of the non-headline paths, with and without the sub-chunk pipeline."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import importlib, os, sys, json, time, numpy as np, torch
sys.path.insert(0, %r)
pkg = importlib.import_module("secp256k1-voi_b200")
n = 1 << 20
eng = pkg.Engine(device=0, max_batch=n, pinned_outputs=True)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
gen = lambda k: tuple(np.array(a) for a in eng.scalar_base_mult(k))  # copies: pinned results are reused per call
def t(f, reps=6):
    for _ in range(2): r = f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): r = f()
    return (time.perf_counter() - t0) / reps * 1e3, r
out = {}
ks = pin(pkg.synth.base_mult_scalars(n))
ms, (pts, st) = t(lambda: eng.scalar_base_mult(ks)); out["scalar_base_mult"] = ms
we = pkg.synth.ecdh_batch(n, gen)
k, p = pin(we["k32"]), pin(we["pt65"])
ms, (x, st) = t(lambda: eng.ecdh(k, p)); out["ecdh"] = ms
x = x.copy(); st = st.copy()
exp, _ = gen(we["closed_form_scalar"])
assert np.array_equal(x, exp[:, 1:33]) and (st == 1).all()
ws = pkg.synth.schnorr_batch(n, gen)
a, b, c = pin(ws["pkx32"]), pin(ws["msg"]), pin(ws["sig64"])
ms, ok = t(lambda: eng.schnorr_verify(a, b, c)); out["schnorr_verify"] = ms
assert np.array_equal(ok, ws["expected"])
w = pkg.synth.ecdsa_batch(n, gen, corrupt_every=0)
priv = pin(np.frombuffer(b"".join(pkg.synth._nonzero_mod_n(v).to_bytes(32, "big") for v in pkg.synth._stream_ints(b"key", 0, n, pkg.synth.SEED)), np.uint8).reshape(n, 32))
dg = pin(w["digest32"])
ms, (sig, rec, st) = t(lambda: eng.ecdsa_sign_rfc6979(priv, dg)); out["ecdsa_sign_rfc6979"] = ms
sig65 = pin(np.concatenate([sig.copy(), rec.copy().reshape(-1, 1)], axis=1))
ms, (pk, st) = t(lambda: eng.ecdsa_recover(dg, sig65)); out["ecdsa_recover"] = ms
assert np.array_equal(pk, w["pk65"]) and (st == 1).all()
print(json.dumps({k_: round(v, 3) for k_, v in out.items()}))
''' % ROOT
for cuts in os.environ.get("CUTS_LIST", "4,16,40").split(";"):
    env = dict(os.environ, S256_PIPE_CUTS=cuts)
    p = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=1200)
    print("S256_PIPE_CUTS=" + cuts, p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-600:], flush=True)
