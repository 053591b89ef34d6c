"""Parity checks shared by the CPU host-simulation tests (kernel logic,
portable arithmetic) and the GPU tests (the product, through the C ABI).
`be` is a backend exposing the Engine method names; `o` is the oracle."""
import hashlib
import importlib

import numpy as np

from conftest import load_golden

synth = importlib.import_module("secp256k1-voi_b200.synth")

N = synth.N
P = synth.P
H = bytes.fromhex


def b32(x):
    return int(x).to_bytes(32, "big")


def rows(lst, w):
    return np.frombuffer(b"".join(lst), np.uint8).reshape(-1, w).copy()


def oracle_base_mult(o):
    return lambda k: o.batch_scalar_base_mult(k)


# ---------------------------------------------------------------------------
def check_field_ops(be, rng_seed=7, n=256):
    rng = np.random.default_rng(rng_seed)
    edge = [0, 1, 2, P - 1, P, P + 1, 2**256 - 1, P - 2, 977, 2**32 + 977, 2**255, N, N - 1, 2**32 - 1, 2**224]
    A = edge + [int.from_bytes(rng.bytes(32), "big") for _ in range(n)]
    B = list(reversed(edge)) + [int.from_bytes(rng.bytes(32), "big") for _ in range(n)]
    # worst cases for the folds: products / sums that land just below 2^256
    A += [2**256 - 1] * 4 + [P - 1, P - 1]
    B += [2**256 - 1, P, P - 1, 2, P - 1, 2]
    a, b = rows([b32(x) for x in A], 32), rows([b32(x) for x in B], 32)
    ops = [(0, lambda x, y: x * y % P), (1, lambda x, y: (x + y) % P), (2, lambda x, y: (x - y) % P),
           (3, lambda x, y: pow(x % P, P - 2, P)), (5, lambda x, y: x * 21 % P), (6, lambda x, y: x * x % P),
           (16, lambda x, y: (x % N) * (y % N) % N), (17, lambda x, y: (x % N + y % N) % N),
           (18, lambda x, y: pow(x % N, N - 2, N))]
    for op, f in ops:
        got = be.debug_field_op(op, a, b)
        for i, (x, y) in enumerate(zip(A, B)):
            assert int.from_bytes(got[i].tobytes(), "big") == f(x, y), (op, i, hex(x), hex(y))
    # sqrt: squares give a root, non-residues give zero (field_sqrt_ratio.go:14-63)
    got = be.debug_field_op(4, a, b)
    for i, x in enumerate(A):
        r = int.from_bytes(got[i].tobytes(), "big")
        xm = x % P
        if pow(xm, (P - 1) // 2, P) in (0, 1):
            assert r * r % P == xm, (i, hex(x))
        else:
            assert r == 0, (i, hex(x))


def check_gen_table(be):
    k = load_golden("kats.json")
    tb = be.debug_gen_table(8, 32)  # byte-for-byte internal/gentable/point_mul_table.bin
    assert hashlib.sha256(tb.tobytes()).hexdigest() == k["gentable_sha256"]


def check_base_mult(be, o, n=64):
    ks = synth.base_mult_scalars(n)
    got, st = be.scalar_base_mult(ks)
    exp, est = o.batch_scalar_base_mult(ks)
    assert np.array_equal(st, est)
    assert np.array_equal(got, exp)
    assert st[0] == 2 and st[4] == 2 and st[1] == 1  # 0*G, n*G = identity
    kats = load_golden("kats.json")
    assert got[1].tobytes().hex() == kats["g_uncompressed"]


def check_rfc6979_and_kats(be, o):
    doc = load_golden("rfc6979.json")
    privs = rows([H(r["priv"]) for r in doc["rows"]], 32)
    pk, st = o.batch_scalar_base_mult(privs)
    got_pk, got_st = be.scalar_base_mult(privs)
    assert np.array_equal(got_pk, pk) and np.array_equal(got_st, st)
    dg = rows([H(r["digest"]) for r in doc["rows"]], 32)
    sig = rows([H(r["r"]) + H(r["s"]) for r in doc["rows"]], 64)
    assert be.ecdsa_verify(pk, dg, sig).tolist() == [1] * len(pk)
    assert be.ecdsa_verify(pk, dg, sig, 1).tolist() == [1] * len(pk)
    bad = sig.copy(); bad[:, 5] ^= 1
    assert be.ecdsa_verify(pk, dg, bad).tolist() == [0] * len(pk)


def check_wycheproof_ecdsa(be, o, limit=None):
    cases = load_golden("wycheproof_ecdsa.json")["cases"]
    if limit:
        # keep every edge-flagged case, thin out the plain ones
        cases = [c for i, c in enumerate(cases) if i % limit == 0 or not c["valid"]
                 or any(f in ("EdgeCaseShamirMultiplication", "PointDuplication", "ArithmeticError", "SmallRandS",
                              "SpecialCaseHash", "EdgeCasePublicKey") for f in c["flags"])]
    pk = rows([H(c["pk"]) for c in cases], 65)
    dg = rows([H(c["digest"])[:32] for c in cases], 32)
    sig = rows([H(c["r"]) + H(c["s"]) for c in cases], 64)
    got = be.ecdsa_verify(pk, dg, sig)
    exp = np.array([c["valid"] for c in cases], np.uint8)
    bad = np.nonzero(got != exp)[0]
    assert len(bad) == 0, [cases[i] for i in bad[:3]]
    # exhaustive-recovery cross-check (wycheproof_test.go:421-438)
    sig65 = np.concatenate([np.concatenate([sig, np.full((len(sig), 1), v, np.uint8)], axis=1) for v in range(4)])
    q, st = be.ecdsa_recover(np.tile(dg, (4, 1)), sig65)
    eq, est = o.batch_ecdsa_recover(np.tile(dg, (4, 1)), sig65)
    assert np.array_equal(st, est)
    assert np.array_equal(q, eq)
    rec = np.zeros(len(sig), bool)
    for v in range(4):
        sl = slice(v * len(sig), (v + 1) * len(sig))
        rec |= (st[sl] == 1) & (q[sl] == pk).all(axis=1)
    assert np.array_equal(rec.astype(np.uint8), exp)


def check_bip340(be):
    doc = load_golden("bip340.json")["rows"]
    by_len = {}
    for r in doc:
        by_len.setdefault(len(r["msg"]) // 2, []).append(r)
    for mlen, rs in by_len.items():
        pk = rows([H(r["pk"]) for r in rs], 32)
        sig = rows([H(r["sig"]) for r in rs], 64)
        msg = rows([H(r["msg"]) for r in rs], mlen) if mlen else np.zeros((len(rs), 0), np.uint8)
        got = be.schnorr_verify(pk, msg, sig)
        assert got.tolist() == [int(r["valid"]) for r in rs], (mlen, got.tolist())


def check_ecdsa_synth(be, o, n=256):
    w = synth.ecdsa_batch(n, oracle_base_mult(o))
    got = be.ecdsa_verify(w["pk65"], w["digest32"], w["sig64"])
    assert np.array_equal(got, w["expected"])
    exp = o.batch_ecdsa_verify(w["pk65"], w["digest32"], w["sig64"])
    assert np.array_equal(got, exp)
    # RejectMalleable: odd items were not low-s normalised
    got1 = be.ecdsa_verify(w["pk65"], w["digest32"], w["sig64"], 1)
    exp1 = o.batch_ecdsa_verify(w["pk65"], w["digest32"], w["sig64"], 1)
    assert np.array_equal(got1, exp1)
    assert got1.sum() < got.sum()


def check_schnorr_synth(be, o, n=128):
    w = synth.schnorr_batch(n, oracle_base_mult(o))
    got = be.schnorr_verify(w["pkx32"], w["msg"], w["sig64"])
    assert np.array_equal(got, w["expected"])
    exp = o.batch_schnorr_verify(w["pkx32"], w["msg"], w["sig64"])
    assert np.array_equal(got, exp)


def edge_ecdsa_inputs(o):
    """Adversarial rows: bad encodings and the exceptional points of the ladder."""
    g, _ = o.scalar_base_mult(b32(1))
    d = 0x1234567890ABCDEF1234567890ABCDEF
    q, _ = o.scalar_base_mult(b32(d))
    z = hashlib.sha256(b"edge").digest()
    zi = int.from_bytes(z, "big") % N

    def sign(k, dd=d, zz=zi):
        R, _ = o.scalar_base_mult(b32(k))
        r = int.from_bytes(R[1:33], "big") % N
        s = pow(k, -1, N) * (zz + r * dd) % N
        return b32(r) + b32(s)

    good = sign(0xC0FFEE)
    out = []

    def add(pk, dg, sg):
        out.append((bytes(pk), bytes(dg), bytes(sg)))

    add(q, z, good)
    add(q, z, b32(0) + good[32:])                      # r = 0
    add(q, z, good[:32] + b32(0))                      # s = 0
    add(q, z, b32(N) + good[32:])                      # r = n (non-canonical)
    add(q, z, good[:32] + b32(N))                      # s = n
    add(q, z, good[:32] + b32(N + 5))                  # s > n
    add(q, z, b32(2**256 - 1) + b32(2**256 - 1))
    add(b"\x05" + q[1:], z, good)                      # bad prefix
    add(b"\x04" + b32(P) + q[33:], z, good)            # x = p (non-canonical)
    add(q[:33] + b32(int.from_bytes(q[33:], "big") ^ 1), z, good)  # off curve
    add(b"\x04" + bytes(64), z, good)                  # (0,0) off curve
    hs = int.from_bytes(good[32:], "big")
    add(q, z, good[:32] + b32(N - hs))                 # the malleable twin verifies too
    # e = 0 (digest = n -> reduces to 0), and digest with all bits set
    add(q, bytes(32), sign(0xABCDEF, zz=0))
    add(q, b32(N), sign(0xABCDEF, zz=0))
    add(q, b"\xff" * 32, sign(0x77, zz=(2**256 - 1) % N))
    # P = G and P = -G (table entries collide with the fixed-base half)
    add(g, z, sign(0x1337, dd=1))
    ng = g[:33] + b32(P - int.from_bytes(g[33:], "big"))
    add(ng, z, sign(0x1337, dd=N - 1))
    # u1*G = -u2*Q  => R = infinity: choose s so that z + r*d = 0 is impossible
    # for a real signature, so craft r, s directly: u1 = z/s, u2 = r/s, want z + r*d = 0.
    r_inf = (-zi) * pow(d, -1, N) % N
    add(q, z, b32(r_inf) + b32(1))
    add(q, z, b32(r_inf) + b32(0xDEADBEEF))
    # u1*G = u2*Q (doubling inside the final add): z = r*d  ->  R = 2*u1*G
    r_dbl = zi * pow(d, -1, N) % N
    add(q, z, b32(r_dbl) + b32(3))
    # r + n < p branch: r small such that x(R) = r + n (cannot forge; must simply be rejected)
    add(q, z, b32(5) + b32(7))
    add(q, z, b32(P - N - 1) + b32(7))
    add(q, z, b32(P - N) + b32(7))
    pk = rows([x[0] for x in out], 65)
    dg = rows([x[1] for x in out], 32)
    sg = rows([x[2] for x in out], 64)
    return pk, dg, sg


def check_ecdsa_edges(be, o):
    pk, dg, sg = edge_ecdsa_inputs(o)
    for flags in (0, 1):
        got = be.ecdsa_verify(pk, dg, sg, flags)
        exp = o.batch_ecdsa_verify(pk, dg, sg, flags)
        assert np.array_equal(got, exp), (flags, got.tolist(), exp.tolist())
    assert got[0] == 1 or exp[0] == 0


def check_double_scalar_mult(be, o, n=96):
    rng = np.random.default_rng(11)
    u1 = [int.from_bytes(rng.bytes(32), "big") for _ in range(n)]
    u2 = [int.from_bytes(rng.bytes(32), "big") for _ in range(n)]
    d = [int.from_bytes(rng.bytes(32), "big") % (N - 1) + 1 for _ in range(n)]
    # exceptional structure: zeros, u1*G = -u2*P, u1*G = u2*P, tiny / huge scalars, P = +-G
    d[0], u1[0], u2[0] = 1, 0, 0
    d[1], u1[1], u2[1] = 5, 0, 1
    d[2], u1[2], u2[2] = 5, 1, 0
    d[3], u1[3], u2[3] = 7, (N - 7 * 9) % N, 9          # sum = identity
    d[4], u1[4], u2[4] = 7, 63, 9                        # final add is a doubling
    d[5], u1[5], u2[5] = 1, 2**255, N - 1
    d[6], u1[6], u2[6] = N - 1, 1, 1                     # G + (-G) = identity
    d[7], u1[7], u2[7] = 3, N, N + 1                     # reduced like NewScalarFromBytes
    d[8], u1[8], u2[8] = 11, 2**256 - 1, 2**256 - 1
    for j, s in enumerate(load_golden("kats.json")["glv_split_scalars"]):
        if 9 + j < n:
            u2[9 + j] = int(s, 16)
    pts, _ = o.batch_scalar_base_mult(rows([b32(x) for x in d], 32))
    U1, U2 = rows([b32(x) for x in u1], 32), rows([b32(x) for x in u2], 32)
    got, st = be.double_scalar_mult_basepoint_vartime(U1, U2, pts)
    exp, est = o.batch_double_scalar_mult(U1, U2, pts)
    assert np.array_equal(st, est), (st.tolist(), est.tolist())
    assert np.array_equal(got, exp)
    assert st[0] == 2 and st[3] == 2 and st[6] == 2
    # invalid points are reported, not computed
    bad = pts.copy(); bad[0, 0] = 2; bad[1, 64] ^= 1
    got, st = be.double_scalar_mult_basepoint_vartime(U1, U2, bad)
    assert st[0] == 0 and st[1] == 0 and np.array_equal(st[2:], est[2:])
    assert not got[0].any() and not got[1].any()


def check_recover_synth(be, o, n=64):
    w = synth.ecdsa_batch(n, oracle_base_mult(o), corrupt_every=0)
    sig65 = np.concatenate([np.concatenate([w["sig64"], np.full((n, 1), v, np.uint8)], axis=1) for v in (0, 1, 2, 3, 4, 27)])
    dg = np.tile(w["digest32"], (6, 1))
    got, st = be.ecdsa_recover(dg, sig65)
    exp, est = o.batch_ecdsa_recover(dg, sig65)
    assert np.array_equal(st, est)
    assert np.array_equal(got, exp)
    hit = np.zeros(n, bool)
    for v in range(4):
        sl = slice(v * n, (v + 1) * n)
        hit |= (st[sl] == 1) & (got[sl] == w["pk65"]).all(axis=1)
    assert hit.all()
    assert not st[4 * n:].any()  # v >= 4 is an error (point_s11n.go:246-248)
