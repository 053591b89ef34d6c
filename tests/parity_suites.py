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
    # rare carry paths of the folds (fe.cuh / fe_vt.cuh): after the first pass the low limbs are
    # ff..f0 | ffffffff | ffffffff, so adding delta = 2^32 + 977 ripples past limb 2 (shallow: stops
    # at limb 3; deep: runs off the top and wraps a second time); likewise for borrows
    D = 2**32 + 977
    X3 = 2**96 - 16
    A += [P - 1, 2**256 - 1, 2**256 - 2, 1, 1, 7 * 2**96 + 5]
    B += [X3 + D + 1, 2**96, 2**256 - 1, 2**256 + 1 - (7 * 2**96 + 5), 2**256 - 4, 2**256 - 3]
    for r in range(21):  # a * 21 = 2^256 + t, t = X3 + r * 2^96: the top fold of mul_small ripples past limb 2
        if (2**256 + X3 + r * 2**96) % 21 == 0:
            A.append((2**256 + X3 + r * 2**96) // 21)
            B.append(3)
            break
    # a * b whose once-folded value lo + hi * delta equals 2^256 + t with t = X3 + (random high limbs):
    # the top fold of the multiplication ripples past limb 2
    made = 0
    for trial in range(200):
        t = X3 + (int.from_bytes(rng.bytes(1), "big") << 96)
        a_ = int.from_bytes(rng.bytes(10), "big") | (1 << 79) | 1
        hi = (-(2**256 + t) * pow(P, -1, a_)) % a_      # makes hi * p + 2^256 + t divisible by a_
        if hi * D <= t:
            continue
        prod = hi * P + 2**256 + t                        # = hi * 2^256 + lo with lo + hi * delta = 2^256 + t
        if prod % a_:
            continue
        b_ = prod // a_
        if b_ < 2**256 and (prod % 2**256) + (prod >> 256) * D == 2**256 + t:
            A.append(a_); B.append(b_); made += 1
            if made == 4:
                break
    assert made == 4
    a, b = rows([b32(x) for x in A], 32), rows([b32(x) for x in B], 32)
    fe_ops = [(0, lambda x, y: x * y % P), (1, lambda x, y: (x + y) % P), (2, lambda x, y: (x - y) % P),
              (5, lambda x, y: x * 21 % P), (6, lambda x, y: x * x % P)]
    # 8a: the bits shifted out fold back as (a >> 253) * delta; inputs chosen so that the fold ripples
    A += [2**256 - 1, (7 << 253) | (2**93 - 1 << 0), (1 << 253) | (2**93 - 2)]
    B += [0, 0, 0]
    # the fused small multiples of the Jacobian formulas (2a, 3a, a - 2b, a - 8b; fe_vt.cuh): top bits shifted out plus
    # the chain's carry / borrow fold together (k up to 8), with results that make the k * delta fold ripple and wrap
    yy = int.from_bytes(rng.bytes(31), "big")
    for k in (1, 2, 3, 7):
        bb = (k << 253) + (yy >> 3)
        A += [(8 * bb + 5) % 2**256, (8 * bb + 5 + 2**96) % 2**256, (2 * bb + 5) % 2**256, (2 * bb + 2**96 + 3) % 2**256, 0, 5]
        B += [bb] * 6
    A += [2**256 - 1 - r for r in range(4)] + [(2**256 + X3) // 3, (2**256 + X3 + 2**96) // 3, (2**257 + X3) // 3 + 1]
    B += [2**256 - 1] * 4 + [0, 0, 0]
    a, b = rows([b32(x) for x in A], 32), rows([b32(x) for x in B], 32)
    ops = fe_ops + [(8 + op, f) for op, f in fe_ops] + [(15, lambda x, y: x * 8 % P)] + [
        (20, lambda x, y: 2 * x % P), (21, lambda x, y: 3 * x % P), (22, lambda x, y: (x - 2 * y) % P),
        (23, lambda x, y: (x - 8 * y) % P), (24, lambda x, y: 2 * x * y % P), (25, lambda x, y: (x * x - y * y) % P)] + [
        (3, lambda x, y: pow(x % P, P - 2, P)), (7, lambda x, y: pow(x % P, P - 2, P)),
        (19, lambda x, y: pow(x % N, N - 2, N)),
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


def check_base_mult_edges(be, o, n=20000):
    """The large-batch fixed-base kernel (7-bit windows, Jacobian accumulator whose exceptional cases are argued away in
    kernels.cuh): scalars at the boundaries of the signed recoding -- every single window digit at +-1, at the largest
    magnitude 2^(WB-1), runs of carries, the top window with its small range, values around n and n/2 -- placed at
    EVEN and odd indices (the host simulation alternates the two formulas by index)."""
    vals = [0, 1, 2, N - 1, N - 2, N, N + 1, 2**256 - 1, N // 2, N // 2 + 1, 2**252, 2**252 - 1, 2**252 + 1, 15 * 2**252,
            2**255, 2**256 - N, 2 * (2**256 - N)]
    for wb in (6, 7):
        nw = (257 + wb - 1) // wb
        for w in range(nw):
            vals += [(1 << (wb * w)) % N, ((1 << (wb - 1)) << (wb * w)) % N, (((1 << (wb - 1)) + 1) << (wb * w)) % N]
        vals += [sum(((1 << (wb - 1)) + 1) << (wb * w) for w in range(nw - 1)) % N,
                 sum(((1 << wb) - 1) << (wb * w) for w in range(nw - 1)) % N, sum(1 << (wb * w) for w in range(nw)) % N]
    vals = [v for v in vals for _ in (0, 1)]     # each value at an even and at an odd index
    ks = synth.base_mult_scalars(n, start=1000)
    edge = rows([b32(v % 2**256) for v in vals], 32)
    ks[16:16 + len(edge)] = edge
    got, st = be.scalar_base_mult(ks)
    m = 16 + len(edge) + 64                       # the oracle on the crafted prefix and a few random rows
    exp, est = o.batch_scalar_base_mult(ks[:m])
    assert np.array_equal(st[:m], est)
    assert np.array_equal(got[:m], exp)
    return got, st


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
    # boundaries of the signed 22-bit comb recoding of u1 (kernels.cuh comb_digit): a digit of exactly
    # 2^21 (last table entry), 2^21 + 1 (-> -(2^21 - 1) with a carry), a run of carries, the top window
    edge_u1 = [2**21, 2**21 + 1, 2**22 - 1, 2**22, (2**21 + 1) * sum(2**(22 * w) for w in range(11)),
               sum((2**22 - 1) << (22 * w) for w in range(11)), N - 1, N - 2**21, 2**242, 2**255 + 2**21 + 1]
    for j, v in enumerate(edge_u1):
        if 40 + j < n:
            u1[40 + j] = v
    # the exceptional branches of the Jacobian ladder (csrc/jac.cuh) away from the final addition: the accumulator
    # equals the comb entry of a middle window (doubling), is its negative (identity, then an assignment from the
    # identity), and the same with the collision in the top window / with the lambda half of u2 empty or alone
    mid = [(1, 2**22 + 3 * 2**44, 2**22), (N - 1, 2**22 + 5 * 2**44, 2**22), (1, 7 * 2**242, 7 * 2**242 % N),
           (N - 1, 7 * 2**242, 7 * 2**242 % N), (1, 2**66 + 2**110, 2**66), (3, 3 * 2**44 + 1, 2**44),
           (N - 3, 3 * 2**44 + 2**200, 2**44), (1, 1, 1), (1, 2, 1), (2, 2, 1), (N - 2, 2, 1), (1, 16, 16), (1, 17, 16)]
    for j, (dd, a, b) in enumerate(mid):
        if 50 + j < n:
            d[50 + j], u1[50 + j], u2[50 + j] = dd, a, b
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


def check_double_scalar_mult_small_multiples(be, o):
    """Small multiples of G everywhere: P = j G, u2 small, u1 = +-(j u2) + delta.  The partial sums of the two halves
    collide all the time (equal points: the doubling branch of the Jacobian ladder; opposite points: its identity
    flag; near misses on either side), in the first comb window, where every one of these u1 lives, and -- for the rows
    with 2^22 and 2^44 factors -- in the second and third."""
    rows_ = []
    for j in (1, 2, 3, 5, 16, 17, 31):
        for u2 in (0, 1, 2, 3, 15, 16, 17, 32, 33, 511):
            for sign in (1, -1):
                for delta in (-1, 0, 1):
                    for scale in (1, 2**22, 2**44):
                        rows_.append((j, (sign * j * u2 * scale + delta) % N, u2 * scale % N))
    d = [r[0] for r in rows_]
    pts, _ = o.batch_scalar_base_mult(rows([b32(x) for x in d], 32))
    U1, U2 = rows([b32(r[1]) for r in rows_], 32), rows([b32(r[2]) for r in rows_], 32)
    got, st = be.double_scalar_mult_basepoint_vartime(U1, U2, pts)
    exp, est = o.batch_double_scalar_mult(U1, U2, pts)
    assert np.array_equal(st, est), np.nonzero(st != est)[0][:10].tolist()
    assert np.array_equal(got, exp), np.nonzero((got != exp).any(axis=1))[0][:10].tolist()
    assert (st == 2).sum() >= len(rows_) // 12  # the identity really occurs (sign = -1, delta = 0, and u2 = 0 with delta = 0)


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


def check_scalar_mult_ecdh(be, o, n=64):
    w = synth.ecdh_batch(n, oracle_base_mult(o))
    k, pts = w["k32"].copy(), w["pt65"].copy()
    # edge scalars: 0, 1, 2, n-1, n (-> 0), n+1, 2^128, lambda, GLV boundary values
    edge = [0, 1, 2, N - 1, N, N + 1, 2**128, 0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72]
    edge += [int(s, 16) for s in load_golden("kats.json")["glv_split_scalars"]]
    for j, v in enumerate(edge[:n]):
        k[j] = np.frombuffer(b32(v % 2**256), np.uint8)
    got, st = be.scalar_mult(k, pts)
    exp, est = o.batch_scalar_mult(k, pts)
    assert np.array_equal(st, est), (st.tolist(), est.tolist())
    assert np.array_equal(got, exp)
    assert st[0] == 2 and st[4] == 2
    # closed form on the untouched tail: k * (d*G) == (k*d mod n) * G
    cf, _ = o.batch_scalar_base_mult(w["closed_form_scalar"])
    assert np.array_equal(got[len(edge):], cf[len(edge):])
    x, xst = be.ecdh(k, pts)
    ex, exst = o.batch_ecdh(k, pts)
    assert np.array_equal(xst, exst) and np.array_equal(x, ex)
    # libsecp256k1 KAT (point_test.go:242-261)
    kat = load_golden("kats.json")
    got, st = be.scalar_mult(rows([H(kat["libsecp_xn"])], 32), rows([H(kat["libsecp_a"])], 65))
    assert st[0] == 1 and got[0].tobytes().hex() == kat["libsecp_b"]
    # invalid points are rejected, not multiplied
    bad = pts[:4].copy(); bad[0, 0] = 3; bad[1, 40] ^= 1; bad[2, 1:33] = 0xFF
    got, st = be.scalar_mult(k[8:12], bad)
    assert st[:3].tolist() == [0, 0, 0] and st[3] == 1 and not got[:3].any()


def check_wycheproof_ecdh(be, o, limit=None):
    cases = load_golden("wycheproof_ecdh.json")["cases"]
    if limit:
        cases = [c for i, c in enumerate(cases) if i % limit == 0 or c["shared"] == "" or len(c["flags"]) > 0 and c["flags"] != ["Normal"]][:400]
    unc = [c for c in cases if len(c["point"]) == 130]
    cmp_ = [c for c in cases if len(c["point"]) == 66]
    # compressed encodings go through the engine's decompressor first
    if cmp_:
        out, st = be.point_decompress(rows([H(c["point"]) for c in cmp_], 33))
        for c, o65, s in zip(cmp_, out, st):
            assert (s == 1) == (c["shared"] != ""), c
            c["point65"] = o65.tobytes()
    for c in unc:
        c["point65"] = H(c["point"])
    live = [c for c in cases if "point65" in c and (len(c["point"]) == 130 or c["shared"] != "")]
    k = rows([H(c["priv"]) for c in live], 32)
    pts = rows([c["point65"] for c in live], 65)
    x, st = be.ecdh(k, pts)
    for c, xi, s in zip(live, x, st):
        if c["shared"] == "":
            assert s == 0, c
        else:
            assert s == 1 and xi.tobytes().hex() == c["shared"], c


def check_point_decompress(be, o, n=64):
    w = synth.ecdh_batch(n, oracle_base_mult(o))
    comp = np.zeros((n, 33), np.uint8)
    comp[:, 0] = 2 + (w["pt65"][:, 64] & 1)
    comp[:, 1:] = w["pt65"][:, 1:33]
    out, st = be.point_decompress(comp)
    assert st.tolist() == [1] * n and np.array_equal(out, w["pt65"])
    flip = comp.copy(); flip[:, 0] ^= 1
    out, st = be.point_decompress(flip)
    assert st.all() and np.array_equal(out[:, :33], w["pt65"][:, :33]) and not np.array_equal(out, w["pt65"])
    bad = comp[:4].copy(); bad[0, 0] = 4; bad[1, 1:] = 0xFF; bad[2, 1:] = 0; bad[2, 32] = 5  # x = 5: x^3+7 = 132 is a non-residue?
    out, st = be.point_decompress(bad)
    exp = [o.point_decode(bytes(b))[1] for b in bad]
    assert st.tolist() == exp
    kats = load_golden("kats.json")
    out, st = be.point_decompress(rows([H(kats["g_compressed"])], 33))
    assert st[0] == 1 and out[0].tobytes().hex() == kats["g_uncompressed"]


def check_point_compress(be, o, n=64):
    """(*Point).CompressedBytes (point_s11n.go:90-117): 02 | 03 by the parity of y, then X; the round trip through
    the decompressor; rows that are not points of the curve are refused; the generator's known encoding."""
    w = synth.ecdh_batch(n, oracle_base_mult(o))
    exp = np.zeros((n, 33), np.uint8)
    exp[:, 0] = 2 + (w["pt65"][:, 64] & 1)
    exp[:, 1:] = w["pt65"][:, 1:33]
    out, st = be.point_compress(w["pt65"])
    assert st.tolist() == [1] * n and np.array_equal(out, exp)
    back, st2 = be.point_decompress(out)
    assert st2.all() and np.array_equal(back, w["pt65"])
    bad = w["pt65"][:4].copy(); bad[0, 0] = 2; bad[1, 64] ^= 1; bad[2, 1:33] = 0xFF; bad[3, 1:] = 0
    out, st = be.point_compress(bad)
    assert st.tolist() == [0, 0, 0, 0] and not out.any()
    kats = load_golden("kats.json")
    out, st = be.point_compress(rows([H(kats["g_uncompressed"])], 65))
    assert st[0] == 1 and out[0].tobytes().hex() == kats["g_compressed"]


def check_msm(be, o, sizes=(0, 1, 2, 31, 32, 33, 64, 300), big=None, heavy=None):
    """point_mul_multi_test.go:14-70 (sizes 0, 1, 32, 64 vs sum of ScalarMult) and
    the closed form of config 5: sum s_i * (d_i G) == (sum s_i d_i mod n) G."""
    w = synth.msm_batch(max(sizes), oracle_base_mult(o))
    for n in sizes:
        k, pts = w["k32"][:n], w["pt65"][:n]
        for vt in (True, False):
            if not vt and n > 64:
                continue
            got, st = be.msm(k, pts, vartime=vt)
            exp, est = o.msm(k.tobytes(), pts.tobytes(), vartime=vt)
            assert st == est, (n, vt, st, est)
            assert got.tobytes() == exp, (n, vt)
    # structure: zero scalars, repeated points, cancelling pairs, scalars = n-1 / 2^255
    n = 40
    k, pts = w["k32"][:n].copy(), w["pt65"][:n].copy()
    k[0] = 0
    k[1] = np.frombuffer(b32(N - 1), np.uint8)
    k[2] = np.frombuffer(b32(2**255), np.uint8)
    k[3] = np.frombuffer(b32(2**256 - 1), np.uint8)   # reduced like NewScalarFromBytes
    pts[5] = pts[4]
    pts[7] = pts[6]; k[7] = np.frombuffer(b32((N - int.from_bytes(k[6].tobytes(), "big")) % N), np.uint8)  # cancels
    pts[11] = pts[10]; k[11] = k[10]            # the same term twice: equal points meet in every bucket (doubling branch)
    pts[13] = pts[12]; pts[14] = pts[12]; k[13] = k[12]; k[14] = k[12]   # and three times
    got, st = be.msm(k, pts)
    exp, est = o.msm(k.tobytes(), pts.tobytes())
    assert (st, got.tobytes()) == (est, exp)
    # everything cancels -> identity
    k2 = np.zeros((2, 32), np.uint8); k2[0, 31] = 5; k2[1] = np.frombuffer(b32(N - 5), np.uint8)
    p2 = np.concatenate([pts[8:9], pts[8:9]])
    got, st = be.msm(k2, p2)
    assert st == 2 and not got.any()
    # an undecodable point poisons the whole product (the reference cannot even build the Point)
    bad = pts.copy(); bad[9, 64] ^= 1
    got, st = be.msm(k, bad)
    assert st == 0 and not got.any()
    # length mismatch panics in the reference (point_mul_multi.go:27-29)
    import pytest
    with pytest.raises(ValueError):
        be.msm(k[:3], pts[:4])
    # skewed scalars: every point lands in the same bucket of every window (exercises bucket slicing)
    ns = min(max(sizes), 300)
    ks = np.tile(np.frombuffer(b32(0x123456789ABCDEF0FEDCBA9876543210 << 64 | 0x55AA), np.uint8), (ns, 1))
    got, st = be.msm(ks, w["pt65"][:ns])
    exp, est = o.msm(ks.tobytes(), w["pt65"][:ns].tobytes())
    assert (st, got.tobytes()) == (est, exp)
    if heavy:
        # one bucket per window holds every point (> 64 slices of 64: the super-slice level);
        # k * sum P_i = (k * sum d_i mod n) * G
        wh = synth.msm_batch(heavy, oracle_base_mult(o))
        kv = 0x0123456789ABCDEF0FEDCBA987654321_1122334455667788_99AABBCCDDEEFF00
        kh = np.tile(np.frombuffer(b32(kv), np.uint8), (heavy, 1))
        got, st = be.msm(kh, wh["pt65"])
        exp, est = o.scalar_base_mult(b32(kv * wh["key_sum"] % N))
        assert st == est and got.tobytes() == exp
    if big:
        wb = synth.msm_batch(big, oracle_base_mult(o))
        got, st = be.msm(wb["k32"], wb["pt65"])
        exp, est = o.scalar_base_mult(wb["closed_form_scalar"])
        assert st == est and got.tobytes() == exp


def check_msm_sharded(be, o, n=256, shards=4):
    """Config 5's multi-GPU shape on one device: per-shard partial sums, one gather, one combine."""
    w = synth.msm_batch(n, oracle_base_mult(o))
    per = n // shards
    parts = []
    for g in range(shards):
        p, st = be.msm_partial(w["k32"][g * per:(g + 1) * per], w["pt65"][g * per:(g + 1) * per])
        assert st == 1
        parts.append(p)
    got, st = be.msm_combine(np.stack(parts))
    exp, est = o.scalar_base_mult(w["closed_form_scalar"])
    assert st == est and got.tobytes() == exp
    # garbage partials are rejected
    bad = np.stack(parts); bad[1, 5] ^= 1
    got, st = be.msm_combine(bad)
    assert st == 0
    got, st = be.msm_combine(np.zeros((0, 96), np.uint8))
    assert st == 2


def check_sign_rfc6979(be, o, n=64):
    """PrivateKey.Sign(RFC6979SHA256(), digest): the 20 sign KATs of the reference
    (secec/ecdsa_k_test.go:244-278), then seeded keys/digests against the oracle, then the
    sign -> verify and sign -> recover round trips."""
    doc = load_golden("rfc6979.json")["rows"]
    priv = rows([H(r["priv"]) for r in doc], 32)
    dg = rows([H(r["digest"]) for r in doc], 32)
    sig, rec, st = be.ecdsa_sign_rfc6979(priv, dg)
    assert st.tolist() == [1] * len(doc)
    for r, s_ in zip(doc, sig):
        assert s_.tobytes().hex() == r["r"] + r["s"], r
    ks = synth.base_mult_scalars(n, start=1000)           # arbitrary 32-byte strings
    ks[0] = 0; ks[1] = np.frombuffer(b32(N), np.uint8); ks[2] = 0xFF        # invalid private keys
    ks[3] = np.frombuffer(b32(1), np.uint8); ks[4] = np.frombuffer(b32(N - 1), np.uint8)
    dgs = synth.base_mult_scalars(n, start=5000)
    dgs[5] = 0; dgs[6] = 0xFF; dgs[7] = np.frombuffer(b32(N), np.uint8)
    sig, rec, st = be.ecdsa_sign_rfc6979(ks, dgs)
    esig, erec, est = o.batch_ecdsa_sign_rfc6979(ks, dgs)
    assert np.array_equal(st, est) and np.array_equal(sig, esig) and np.array_equal(rec, erec)
    assert st[:3].tolist() == [0, 0, 0] and not sig[:3].any()
    good = st == 1
    pk, pst = o.batch_scalar_base_mult(ks[good])
    assert o.batch_ecdsa_verify(pk, dgs[good], sig[good], 1).all()   # low-s signatures verify
    q, qst = o.batch_ecdsa_recover(dgs[good], np.concatenate([sig[good], rec[good, None]], axis=1))
    assert (qst == 1).all() and np.array_equal(q, pk)


def check_schnorr_sign(be, o, n=48):
    """SchnorrPrivateKey.Sign: the BIP-340 vector rows that carry a secret key (bit-exact signature),
    then seeded keys against the oracle and sign -> verify."""
    doc = [r for r in load_golden("bip340.json")["rows"] if r["sk"]]
    by_len = {}
    for r in doc:
        by_len.setdefault(len(r["msg"]) // 2, []).append(r)
    for mlen, rs in by_len.items():
        priv = rows([H(r["sk"]) for r in rs], 32)
        aux = rows([H(r["aux"]) for r in rs], 32)
        msg = rows([H(r["msg"]) for r in rs], mlen) if mlen else np.zeros((len(rs), 0), np.uint8)
        sig, st = be.schnorr_sign(priv, msg, aux)
        assert st.tolist() == [1] * len(rs)
        for r, s_ in zip(rs, sig):
            assert s_.tobytes().hex() == r["sig"], r["index"]
    priv = synth.base_mult_scalars(n, start=3000)
    priv[0] = 0; priv[1] = np.frombuffer(b32(N), np.uint8); priv[2] = np.frombuffer(b32(N - 1), np.uint8)
    priv[3] = np.frombuffer(b32(1), np.uint8)
    msg = synth.base_mult_scalars(n, start=4000)
    aux = synth.base_mult_scalars(n, start=6000); aux[4] = 0; aux[5] = 0xFF
    sig, st = be.schnorr_sign(priv, msg, aux)
    exp = [o.schnorr_sign(priv[i].tobytes(), msg[i].tobytes(), aux[i].tobytes()) for i in range(n)]
    assert st.tolist() == [e[1] for e in exp]
    assert sig.tobytes() == b"".join(e[0] for e in exp)
    assert st[0] == 0 and st[1] == 0 and st[2] == 1 and not sig[:2].any()
    good = st == 1
    pk, _ = o.batch_scalar_base_mult(priv[good])
    assert o.batch_schnorr_verify(pk[:, 1:33].copy(), msg[good], sig[good]).all()


def check_hash_to_curve(be, o, n=32):
    """RFC 9380 suite vectors (secec/h2c/h2c_test.go:35-120), expander vectors that fit 96 bytes,
    then seeded messages against the oracle, both suites, odd message lengths, an oversize DST."""
    doc = load_golden("h2c.json")
    for s_ in doc["suites"]:
        by_len = {}
        for v in s_["vectors"]:
            by_len.setdefault(len(v["msg"]), []).append(v)
        for mlen, vs in by_len.items():
            msgs = rows([v["msg"].encode() for v in vs], mlen) if mlen else np.zeros((len(vs), 0), np.uint8)
            out, st = be.hash_to_curve(s_["dst"].encode(), msgs, s_["random_oracle"])
            assert st.tolist() == [1] * len(vs)
            for v, o65 in zip(vs, out):
                assert o65[1:].tobytes().hex() == v["Px"] + v["Py"], v["msg"][:20]
    for e in doc["expand"]:
        for tc in e["tests"]:
            if tc["len"] > 96:
                continue
            m = np.frombuffer(tc["msg"].encode(), np.uint8).reshape(1, -1) if tc["msg"] else np.zeros((1, 0), np.uint8)
            got = be.expand_message_xmd(e["dst"].encode(), m, tc["len"])
            assert got[0].tobytes().hex() == tc["uniform_bytes"], tc["msg"][:20]
    for mlen in (0, 1, 31, 64, 119):
        msgs = np.frombuffer(b"".join(synth.D(b"h2c", i)[:1] * mlen if mlen else b"" for i in range(n)), np.uint8).reshape(n, mlen) \
            if mlen else np.zeros((n, 0), np.uint8)
        if mlen:
            msgs = (msgs + np.arange(n, dtype=np.uint8)[:, None] * 7 + np.arange(mlen, dtype=np.uint8)[None, :]).astype(np.uint8)
        for ro in (True, False):
            for dst in (b"QUUX-V01-CS02-with-secp256k1_XMD:SHA-256_SSWU_RO_", b"d", b"z" * 300):
                out, st = be.hash_to_curve(dst, msgs, ro)
                for i in range(0, n, max(1, n // 8)):
                    exp, est = o.hash_to_curve(dst, msgs[i].tobytes(), ro)
                    assert (st[i], out[i].tobytes()) == (est, exp), (mlen, ro, len(dst), i)
