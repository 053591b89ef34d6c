"""ctypes view of the CPU oracle (oracle/secp256k1_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Nothing under
secp256k1-voi_b200/ may import this module.

Byte conventions follow the reference (all big-endian): scalars / digests /
x-coordinates 32 B, compact signatures r||s 64 B (r||s||v 65 B), public keys
SEC 1 uncompressed 65 B.  Status bytes: 0 invalid, 1 ok, 2 identity.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_SRC = os.path.join(_HERE, "secp256k1_oracle.c")

ST_INVALID, ST_OK, ST_IDENTITY = 0, 1, 2
FLAG_REJECT_MALLEABLE = 1


def build(force=False):
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(_SRC)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "_build/liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_init()
    return _lib


def _buf(b, n=None):
    b = bytes(b)
    if n is not None and len(b) != n:
        raise ValueError(f"expected {n} bytes, got {len(b)}")
    return b


def _out(n):
    return C.create_string_buffer(n)


# ---- single-item helpers --------------------------------------------------
def fe_set_bytes(b):
    o = _out(32); d = lib().orc_fe_set_bytes(_buf(b, 32), o); return o.raw, d


def fe_bytes_are_canonical(b):
    return bool(lib().orc_fe_bytes_are_canonical(_buf(b, 32)))


def fe_mul(a, b):
    o = _out(32); lib().orc_fe_mul(_buf(a, 32), _buf(b, 32), o); return o.raw


def fe_invert(a):
    o = _out(32); lib().orc_fe_invert(_buf(a, 32), o); return o.raw


def fe_sqrt(a):
    o = _out(32); ok = lib().orc_fe_sqrt(_buf(a, 32), o); return o.raw, ok


def sc_set_bytes(b):
    o = _out(32); d = lib().orc_sc_set_bytes(_buf(b, 32), o); return o.raw, d


def sc_bytes_are_canonical(b):
    return bool(lib().orc_sc_bytes_are_canonical(_buf(b, 32)))


def sc_is_gt_half_n(b):
    return lib().orc_sc_is_gt_half_n(_buf(b, 32))


def sc_mul(a, b):
    o = _out(32); lib().orc_sc_mul(_buf(a, 32), _buf(b, 32), o); return o.raw


def sc_add(a, b):
    o = _out(32); lib().orc_sc_add(_buf(a, 32), _buf(b, 32), o); return o.raw


def sc_invert(a):
    o = _out(32); lib().orc_sc_invert(_buf(a, 32), o); return o.raw


def sc_split_glv(k):
    a, b = _out(32), _out(32); lib().orc_sc_split_glv(_buf(k, 32), a, b); return a.raw, b.raw


def sha256(d):
    o = _out(32); d = bytes(d); lib().orc_sha256(d, C.c_size_t(len(d)), o); return o.raw


def gen_table_bytes():
    o = _out(522240); lib().orc_gen_table_bytes(o); return o.raw


def point_decode(b):
    b = bytes(b); o = _out(65); st = lib().orc_point_decode(b, C.c_size_t(len(b)), o); return o.raw, st


def scalar_base_mult(k, vartime=False):
    o = _out(65)
    f = lib().orc_scalar_base_mult_vartime if vartime else lib().orc_scalar_base_mult
    st = f(_buf(k, 32), o)
    return o.raw, st


def scalar_mult(k, pt65, mode=0):
    """mode 0: ct GLV ScalarMult, 1: vartime GLV, 2: bit-serial trivial."""
    o = _out(65); st = lib().orc_scalar_mult(_buf(k, 32), _buf(pt65, 65), mode, o); return o.raw, st


def ecdh(k, pt65):
    o = _out(32); st = lib().orc_ecdh(_buf(k, 32), _buf(pt65, 65), o); return o.raw, st


def double_scalar_mult_basepoint_vartime(u1, u2, pt65):
    o = _out(65)
    st = lib().orc_double_scalar_mult_basepoint_vartime(_buf(u1, 32), _buf(u2, 32), _buf(pt65, 65), o)
    return o.raw, st


def point_add(a65, a_st, b65, b_st):
    o = _out(65); st = lib().orc_point_add(_buf(a65, 65), a_st, _buf(b65, 65), b_st, o); return o.raw, st


def msm(ks, pts, vartime=True):
    ks, pts = bytes(ks), bytes(pts)
    n = len(ks) // 32
    assert len(pts) == 65 * n
    o = _out(65); st = lib().orc_msm(ks, pts, C.c_size_t(n), int(vartime), o); return o.raw, st


def ecdsa_verify(pk65, digest32, sig64, flags=0):
    return lib().orc_ecdsa_verify(_buf(pk65, 65), _buf(digest32, 32), _buf(sig64, 64), C.c_uint32(flags))


def ecdsa_recover(digest32, sig65):
    o = _out(65); st = lib().orc_ecdsa_recover(_buf(digest32, 32), _buf(sig65, 65), o); return o.raw, st


def ecdsa_sign_rfc6979(priv32, digest32):
    sig = _out(64); rec = C.c_uint8(0)
    st = lib().orc_ecdsa_sign_rfc6979(_buf(priv32, 32), _buf(digest32, 32), sig, C.byref(rec))
    return sig.raw, rec.value, st


def hmac_sha256(key32, msg):
    o = _out(32); msg = bytes(msg); lib().orc_hmac_sha256(_buf(key32, 32), msg, C.c_size_t(len(msg)), o); return o.raw


def hash_to_curve(dst, msg, random_oracle=True):
    dst, msg = bytes(dst), bytes(msg); o = _out(65)
    st = lib().orc_hash_to_curve(dst, C.c_size_t(len(dst)), msg, C.c_size_t(len(msg)), int(random_oracle), o)
    return o.raw, st


def expand_message_xmd(dst, msg, length):
    dst, msg = bytes(dst), bytes(msg); o = _out(length)
    ok = lib().orc_expand_message_xmd(dst, C.c_size_t(len(dst)), msg, C.c_size_t(len(msg)), o, C.c_size_t(length))
    return o.raw if ok else None


def map_to_curve(u48):
    o = _out(65); st = lib().orc_map_to_curve(_buf(u48, 48), o); return o.raw, st


def schnorr_sign(priv32, msg, aux32):
    sig = _out(64); msg = bytes(msg)
    st = lib().orc_schnorr_sign(_buf(priv32, 32), msg, C.c_size_t(len(msg)), _buf(aux32, 32), sig)
    return sig.raw, st


def schnorr_verify(pkx32, msg, sig64):
    msg = bytes(msg)
    return lib().orc_schnorr_verify(_buf(pkx32, 32), msg, C.c_size_t(len(msg)), _buf(sig64, 64))


# ---- threaded batch drivers (numpy uint8 arrays, row-major) ----------------
def _np(a, width):
    a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1, width)
    return a


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def default_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def batch_scalar_base_mult(k, vartime=False, threads=None):
    k = _np(k, 32); n = len(k)
    out = np.zeros((n, 65), np.uint8); st = np.zeros(n, np.uint8)
    lib().orc_batch_scalar_base_mult(_p(k), C.c_size_t(n), int(vartime), _p(out), _p(st), threads or default_threads())
    return out, st


def batch_scalar_mult(k, pts, threads=None):
    k = _np(k, 32); pts = _np(pts, 65); n = len(k)
    out = np.zeros((n, 65), np.uint8); st = np.zeros(n, np.uint8)
    lib().orc_batch_scalar_mult(_p(k), _p(pts), C.c_size_t(n), _p(out), _p(st), threads or default_threads())
    return out, st


def batch_ecdh(k, pts, threads=None):
    k = _np(k, 32); pts = _np(pts, 65); n = len(k)
    out = np.zeros((n, 32), np.uint8); st = np.zeros(n, np.uint8)
    lib().orc_batch_ecdh(_p(k), _p(pts), C.c_size_t(n), _p(out), _p(st), threads or default_threads())
    return out, st


def batch_double_scalar_mult(u1, u2, pts, threads=None):
    u1 = _np(u1, 32); u2 = _np(u2, 32); pts = _np(pts, 65); n = len(u1)
    out = np.zeros((n, 65), np.uint8); st = np.zeros(n, np.uint8)
    lib().orc_batch_double_scalar_mult(_p(u1), _p(u2), _p(pts), C.c_size_t(n), _p(out), _p(st), threads or default_threads())
    return out, st


def batch_ecdsa_verify(pk, digest, sig, flags=0, threads=None):
    pk = _np(pk, 65); digest = _np(digest, 32); sig = _np(sig, 64); n = len(pk)
    ok = np.zeros(n, np.uint8)
    lib().orc_batch_ecdsa_verify(_p(pk), _p(digest), _p(sig), C.c_uint32(flags), C.c_size_t(n), _p(ok), threads or default_threads())
    return ok


def batch_ecdsa_recover(digest, sig65, threads=None):
    digest = _np(digest, 32); sig65 = _np(sig65, 65); n = len(digest)
    out = np.zeros((n, 65), np.uint8); st = np.zeros(n, np.uint8)
    lib().orc_batch_ecdsa_recover(_p(digest), _p(sig65), C.c_size_t(n), _p(out), _p(st), threads or default_threads())
    return out, st


def batch_schnorr_verify(pkx, msg, sig, threads=None):
    pkx = _np(pkx, 32); sig = _np(sig, 64); n = len(pkx)
    msg = np.ascontiguousarray(msg, dtype=np.uint8).reshape(n, -1) if n else np.zeros((0, 32), np.uint8)
    ok = np.zeros(n, np.uint8)
    lib().orc_batch_schnorr_verify(_p(pkx), _p(msg), C.c_size_t(msg.shape[1]), _p(sig), C.c_size_t(n), _p(ok), threads or default_threads())
    return ok


def batch_ecdsa_sign_rfc6979(priv, digest, threads=None):
    priv = _np(priv, 32); digest = _np(digest, 32); n = len(priv)
    sig = np.zeros((n, 64), np.uint8); rec = np.zeros(n, np.uint8); st = np.zeros(n, np.uint8)
    lib().orc_batch_ecdsa_sign_rfc6979(_p(priv), _p(digest), C.c_size_t(n), _p(sig), _p(rec), _p(st), threads or default_threads())
    return sig, rec, st


# ---- OpenSSL sanity anchor (oracle/openssl_baseline.c): measurement only -------------------------------------------
def openssl_ecdsa_verify(pk, digest, sig, reps=1, threads=None):
    """(ok bytes, seconds for `reps` passes of ECDSA_do_verify on `threads` threads), or None when libcrypto / its headers
    are not there to build against."""
    so = os.path.join(_HERE, "_build", "libosslbase.so")
    src = os.path.join(_HERE, "openssl_baseline.c")
    try:
        if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-B", "_build/libosslbase.so"], stdout=subprocess.DEVNULL,
                                  stderr=subprocess.DEVNULL)
        L = C.CDLL(so)
    except (OSError, subprocess.CalledProcessError):
        return None
    pk, digest, sig = _np(pk, 65), _np(digest, 32), _np(sig, 64)
    n = len(pk)
    ok = np.zeros(n, np.uint8)
    secs = C.c_double(0)
    rc = L.ossl_batch_ecdsa_verify(_p(pk), _p(digest), _p(sig), C.c_size_t(n), int(reps), int(threads or default_threads()),
                                   _p(ok), C.byref(secs))
    if rc != 0:
        return None
    return ok, secs.value
