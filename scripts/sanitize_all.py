"""Every entry point once at small sizes (driven under compute-sanitizer)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
pkg = importlib.import_module("secp256k1-voi_b200")
eng = pkg.Engine(device=0, max_batch=4096)
for n in (1, 33, 1500):
    ks = pkg.synth.base_mult_scalars(n)
    pk, st = eng.scalar_base_mult(ks)
    w = pkg.synth.ecdsa_batch(n, eng.scalar_base_mult)
    assert np.array_equal(eng.ecdsa_verify(w["pk65"], w["digest32"], w["sig64"]), w["expected"])
    sig65 = np.concatenate([w["sig64"], np.zeros((n, 1), np.uint8)], axis=1)
    eng.ecdsa_recover(w["digest32"], sig65)
    ws = pkg.synth.schnorr_batch(n, eng.scalar_base_mult)
    assert np.array_equal(eng.schnorr_verify(ws["pkx32"], ws["msg"], ws["sig64"]), ws["expected"])
    we = pkg.synth.ecdh_batch(n, eng.scalar_base_mult)
    eng.scalar_mult(we["k32"], we["pt65"]); eng.ecdh(we["k32"], we["pt65"])
    eng.double_scalar_mult_basepoint_vartime(we["k32"], ks, we["pt65"])
    for vt in (True, False):
        eng.msm(we["k32"], we["pt65"], vartime=vt)
    comp = np.concatenate([2 + (we["pt65"][:, 64:] & 1), we["pt65"][:, 1:33]], axis=1)
    eng.point_decompress(comp)
print("sanitize_all ok")
