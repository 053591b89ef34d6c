"""ctypes view of tests/hostsim/hostsim.cpp (CPU simulation of the kernel
logic; test infrastructure only, see the header of hostsim.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_SO = os.path.join(_HERE, "_build", "hostsim.so")
_lib = None


def _deps():
    csrc = os.path.join(_ROOT, "secp256k1-voi_b200", "csrc")
    return [os.path.join(_HERE, "hostsim.cpp")] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]


def lib():
    global _lib
    if _lib is None:
        stale = not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in _deps())
        if stale:
            os.makedirs(os.path.dirname(_SO), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
                                   "-o", _SO, os.path.join(_HERE, "hostsim.cpp")])
        _lib = C.CDLL(_SO)
    return _lib


def _a(x, w):
    return np.ascontiguousarray(x, dtype=np.uint8).reshape(-1, w)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def ecdsa_verify(pk, dg, sig, flags=0):
    pk, dg, sig = _a(pk, 65), _a(dg, 32), _a(sig, 64)
    n = len(pk); ok = np.zeros(n, np.uint8)
    lib().sim_ecdsa_verify(_p(pk), _p(dg), _p(sig), C.c_uint32(flags), C.c_size_t(n), _p(ok))
    return ok


def ecdsa_recover(dg, sig65):
    dg, sig65 = _a(dg, 32), _a(sig65, 65)
    n = len(dg); out = np.zeros((n, 65), np.uint8); st = np.zeros(n, np.uint8)
    lib().sim_ecdsa_recover(_p(dg), _p(sig65), C.c_size_t(n), _p(out), _p(st))
    return out, st


def schnorr_verify(pkx, msg, sig):
    pkx, sig = _a(pkx, 32), _a(sig, 64)
    n = len(pkx)
    msg = np.ascontiguousarray(msg, dtype=np.uint8).reshape(n, -1)
    ok = np.zeros(n, np.uint8)
    lib().sim_schnorr_verify(_p(pkx), _p(msg), C.c_size_t(msg.shape[1]), _p(sig), C.c_size_t(n), _p(ok))
    return ok


def double_scalar_mult(u1, u2, pts):
    u1, u2, pts = _a(u1, 32), _a(u2, 32), _a(pts, 65)
    n = len(u1); out = np.zeros((n, 65), np.uint8); st = np.zeros(n, np.uint8)
    lib().sim_double_scalar_mult(_p(u1), _p(u2), _p(pts), C.c_size_t(n), _p(out), _p(st))
    return out, st


def scalar_base_mult(k):
    k = _a(k, 32); n = len(k); out = np.zeros((n, 65), np.uint8); st = np.zeros(n, np.uint8)
    lib().sim_scalar_base_mult(_p(k), C.c_size_t(n), _p(out), _p(st))
    return out, st


def gen_table(wbits, nwin):
    out = np.zeros(((2 ** wbits - 1) * nwin, 64), np.uint8)
    lib().sim_gen_table(wbits, nwin, _p(out))
    return out


def field_op(op, a, b):
    a, b = _a(a, 32), _a(b, 32); n = len(a); out = np.zeros((n, 32), np.uint8)
    lib().sim_field_op(op, _p(a), _p(b), C.c_size_t(n), _p(out))
    return out


def scalar_mult(k, pts):
    k, pts = _a(k, 32), _a(pts, 65); n = len(k); out = np.zeros((n, 65), np.uint8); st = np.zeros(n, np.uint8)
    lib().sim_scalar_mult(_p(k), _p(pts), C.c_size_t(n), 0, _p(out), _p(st))
    return out, st


def ecdh(k, pts):
    k, pts = _a(k, 32), _a(pts, 65); n = len(k); out = np.zeros((n, 32), np.uint8); st = np.zeros(n, np.uint8)
    lib().sim_scalar_mult(_p(k), _p(pts), C.c_size_t(n), 1, _p(out), _p(st))
    return out, st


def point_decompress(p33):
    p33 = _a(p33, 33); n = len(p33); out = np.zeros((n, 65), np.uint8); st = np.zeros(n, np.uint8)
    lib().sim_point_decompress(_p(p33), C.c_size_t(n), _p(out), _p(st))
    return out, st


def msm(k, pts, vartime=True, force_c=0):
    k, pts = _a(k, 32), _a(pts, 65); n = len(k)
    if len(pts) != n:
        raise ValueError("secp256k1: len(scalars) != len(points)")
    out = np.zeros(65, np.uint8); st = C.c_uint8(0)
    lib().sim_msm(_p(k), _p(pts), C.c_size_t(n), int(vartime), int(force_c), _p(out), C.byref(st), None)
    return out, st.value


def msm_partial(k, pts, vartime=True, force_c=0):
    k, pts = _a(k, 32), _a(pts, 65); n = len(k)
    out = np.zeros(65, np.uint8); part = np.zeros(96, np.uint8); st = C.c_uint8(0)
    lib().sim_msm(_p(k), _p(pts), C.c_size_t(n), int(vartime), int(force_c), _p(out), C.byref(st), _p(part))
    return part, (1 if st.value in (1, 2) else 0)


def msm_combine(parts):
    parts = _a(parts, 96); out = np.zeros(65, np.uint8); st = C.c_uint8(0)
    lib().sim_msm_combine(_p(parts), C.c_size_t(len(parts)), _p(out), C.byref(st))
    return out, st.value


def ecdsa_sign_rfc6979(priv, digest):
    priv, digest = _a(priv, 32), _a(digest, 32); n = len(priv)
    sig = np.zeros((n, 64), np.uint8); rec = np.zeros(n, np.uint8); st = np.zeros(n, np.uint8)
    lib().sim_ecdsa_sign_rfc6979(_p(priv), _p(digest), C.c_size_t(n), _p(sig), _p(rec), _p(st))
    return sig, rec, st


def schnorr_sign(priv, msg, aux):
    priv, aux = _a(priv, 32), _a(aux, 32); n = len(priv)
    msg = np.ascontiguousarray(msg, dtype=np.uint8).reshape(n, -1)
    sig = np.zeros((n, 64), np.uint8); st = np.zeros(n, np.uint8)
    lib().sim_schnorr_sign(_p(priv), _p(msg), C.c_size_t(msg.shape[1]), _p(aux), C.c_size_t(n), _p(sig), _p(st))
    return sig, st


def hash_to_curve(dst, msgs, random_oracle=True):
    import hashlib
    dst = bytes(dst)
    if len(dst) > 255:  # the library pre-hashes oversize DSTs on the host
        dst = hashlib.sha256(b"H2C-OVERSIZE-DST-" + dst).digest()
    m = np.ascontiguousarray(msgs, dtype=np.uint8); n, ml = m.shape
    out = np.zeros((n, 65), np.uint8); st = np.zeros(n, np.uint8)
    lib().sim_hash_to_curve(dst, C.c_size_t(len(dst)), _p(m), C.c_size_t(ml), C.c_size_t(n), int(random_oracle), _p(out), _p(st))
    return out, st


def expand_message_xmd(dst, msgs, length):
    import hashlib
    dst = bytes(dst)
    if len(dst) > 255:
        dst = hashlib.sha256(b"H2C-OVERSIZE-DST-" + dst).digest()
    m = np.ascontiguousarray(msgs, dtype=np.uint8); n, ml = m.shape
    out = np.zeros((n, length), np.uint8)
    for i in range(n):
        row = np.zeros(length, np.uint8)
        lib().sim_expand_xmd(dst, C.c_size_t(len(dst)), _p(m[i:i + 1].copy()), C.c_size_t(ml), int(length), _p(row))
        out[i] = row
    return out


def jac_vs_complete(a65, b65, ops):
    """The same chain of doublings / additions on the Jacobian and on the complete formulas: (out, status) of each."""
    a, b = _a(a65, 65), _a(b65, 65)
    ops = np.ascontiguousarray(ops, dtype=np.uint8)
    oj, oc = np.zeros(65, np.uint8), np.zeros(65, np.uint8)
    sj, sc_ = C.c_uint8(0), C.c_uint8(0)
    lib().sim_jac_vs_complete(_p(a), _p(b), _p(ops), C.c_size_t(len(ops)), _p(oj), C.byref(sj), _p(oc), C.byref(sc_))
    return (oj, sj.value), (oc, sc_.value)
