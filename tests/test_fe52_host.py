"""The FP64-pipe field multiplication experiment (csrc/fe52.cuh) compiled
for the host -- fma() under FE_TOWARDZERO in place of fma.rz.f64 -- against Python integers:
the result is congruent to a*b mod p, limbs below 2^52, top limb inside the loose bound."""
import ctypes as C
import os
import random
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
P = 2**256 - 2**32 - 977
M52 = 2**52 - 1
TOP_MAX = 2**48 + 2**6


def _lib():
    out = os.path.join(HERE, "cpp", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libfe52_host.so")
    src = os.path.join(HERE, "cpp", "fe52_host.cpp")
    hdr = os.path.join(HERE, "..", "secp256k1-voi_b200", "csrc", "fe52.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-frounding-math", "-ffp-contract=off", "-x", "c++", "-shared", "-fPIC", "-o", so, src])
    return C.CDLL(so)


def _limbs(x):
    return [(x >> (52 * i)) & M52 for i in range(4)] + [x >> 208]


def _value(l):
    return sum(int(v) << (52 * i) for i, v in enumerate(l))


def test_fe52_mul_matches_integers():
    lib = _lib()
    rng = random.Random(20261017)
    top = (TOP_MAX << 208) | (2**208 - 1)  # every limb at its loose maximum
    edge = [0, 1, 2, P - 1, P, P + 1, 2**256 - 1, 2**256, top, 2**255, 2**52 - 1, 2**52, 2**208 - 1, 977, 2**32 + 977]
    pairs = [(x, y) for x in edge for y in edge]
    pairs += [(rng.getrandbits(256), rng.getrandbits(256)) for _ in range(20000)]
    pairs += [(rng.getrandbits(256) | (2**256 - 2**200), rng.getrandbits(256) | (2**256 - 2**200)) for _ in range(2000)]
    # limbs of all ones / sparse limbs stress the column carries
    pairs += [(_value([rng.choice([0, M52, 1, M52 - 1]) for _ in range(4)] + [rng.choice([0, 2**48 - 1, TOP_MAX])]),
               _value([rng.choice([0, M52, 1, M52 - 1]) for _ in range(4)] + [rng.choice([0, 2**48 - 1, TOP_MAX])])) for _ in range(4000)]
    n = len(pairs)
    a = np.array([_limbs(x) for x, _ in pairs], dtype=np.uint64)
    b = np.array([_limbs(y) for _, y in pairs], dtype=np.uint64)
    assert int(a[:, 4].max()) <= TOP_MAX and int(b[:, 4].max()) <= TOP_MAX
    r = np.zeros((n, 5), dtype=np.uint64)
    u64p = C.POINTER(C.c_uint64)
    lib.fe52_mul_host(a.ctypes.data_as(u64p), b.ctypes.data_as(u64p), r.ctypes.data_as(u64p), C.c_size_t(n))
    for (x, y), l in zip(pairs, r):
        assert all(int(v) <= M52 for v in l[:4]) and int(l[4]) <= TOP_MAX
        assert _value(l) % P == (x * y) % P
    # a chain of dependent products stays inside the bound (outputs fed back as inputs)
    x = np.array([_limbs(top)], dtype=np.uint64)
    acc, want = x.copy(), top
    for _ in range(200):
        out = np.zeros((1, 5), dtype=np.uint64)
        lib.fe52_mul_host(acc.ctypes.data_as(u64p), x.ctypes.data_as(u64p), out.ctypes.data_as(u64p), C.c_size_t(1))
        want = want * top % P
        assert int(out[0, 4]) <= TOP_MAX and _value(out[0]) % P == want
        acc = out
