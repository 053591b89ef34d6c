"""Property tests (hypothesis) over the host-simulated kernel logic, in the spirit of the reference's
randomised consistency tests (point_test.go:262-347: fast path == bit-serial double-and-add after
rescale) but seeded and biased towards the values where windowed / GLV / carry code breaks."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

import hostsim as hs
import parity_suites as ps

N, P = ps.N, ps.P
EDGE = [0, 1, 2, 3, 7, 8, 15, 16, 17, 31, 32, N - 1, N - 2, N, N + 1, N // 2, N // 2 + 1, 2**128 - 1, 2**128, 2**128 + 1,
        2**255, 2**256 - 1, 2**64 - 1, 2**192, 0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72,
        0xAC9C52B33FA3CF1F5AD9E3FD77ED9BA4A880B9FC8EC739C2E0CFC810B51283CF]
# scalars: edges, sparse / dense bit patterns, runs of identical nibbles (window recoding carries), random
scalar = st.one_of(
    st.sampled_from(EDGE),
    st.integers(0, 2**256 - 1),
    st.builds(lambda nib, cnt, sh: (int(("%x" % nib) * cnt, 16) << sh) % 2**256, st.integers(1, 15), st.integers(1, 64), st.integers(0, 200)),
    st.builds(lambda a, b: (1 << a) | (1 << b), st.integers(0, 255), st.integers(0, 255)),
    st.builds(lambda a: (N - a) % 2**256, st.integers(0, 2**40)),
)
SET = dict(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck))


def b32(x):
    return int(x).to_bytes(32, "big")


def rows(ints):
    return np.frombuffer(b"".join(b32(x) for x in ints), np.uint8).reshape(-1, 32).copy()


@settings(**SET)
@given(st.lists(scalar, min_size=1, max_size=6))
def test_base_mult_matches_trivial(oracle, ks):
    got, stt = hs.scalar_base_mult(rows(ks))
    g, _ = oracle.scalar_base_mult(b32(1))
    for k, o65, s in zip(ks, got, stt):
        exp, est = oracle.scalar_mult(b32(k), g, 2)  # bit-serial double-and-add
        assert (s, o65.tobytes()) == (est, exp)


@settings(**SET)
@given(st.lists(st.tuples(scalar, scalar, st.integers(1, 2**64)), min_size=1, max_size=4))
def test_dsm_and_ct_mult_match_trivial(oracle, items):
    u1 = [a for a, _, _ in items]; u2 = [b for _, b, _ in items]
    pts, _ = oracle.batch_scalar_base_mult(rows([d for _, _, d in items]))
    got, stt = hs.double_scalar_mult(rows(u1), rows(u2), pts)
    ctg, cts = hs.scalar_mult(rows(u2), pts)
    g, _ = oracle.scalar_base_mult(b32(1))
    for i in range(len(items)):
        a, ast = oracle.scalar_mult(b32(u1[i]), g, 2)
        b, bst = oracle.scalar_mult(b32(u2[i]), pts[i].tobytes(), 2)
        exp, est = oracle.point_add(a, ast, b, bst)
        assert (stt[i], got[i].tobytes()) == (est, exp)
        assert (cts[i], ctg[i].tobytes()) == (bst, b)   # constant-time ladder == bit-serial too


@settings(**SET)
@given(st.lists(st.tuples(scalar, st.integers(1, 2**64)), min_size=0, max_size=40), st.sampled_from([0, 4, 5, 8, 13, 16]))
def test_msm_matches_sum_of_products(oracle, items, c):
    ks = rows([k for k, _ in items]) if items else np.zeros((0, 32), np.uint8)
    pts = oracle.batch_scalar_base_mult(rows([d for _, d in items]))[0] if items else np.zeros((0, 65), np.uint8)
    got, s = hs.msm(ks, pts, vartime=True, force_c=c if len(items) else 0)
    total = sum((k % N) * d for k, d in items) % N
    exp, est = oracle.scalar_base_mult(b32(total))
    assert (s, got.tobytes()) == (est, exp)


@settings(**SET)
@given(st.integers(0, 2**256 - 1), st.integers(0, 2**256 - 1))
def test_field_ops(a, b):
    A, B = rows([a]), rows([b])
    assert int.from_bytes(hs.field_op(0, A, B)[0].tobytes(), "big") == a * b % P
    assert int.from_bytes(hs.field_op(1, A, B)[0].tobytes(), "big") == (a + b) % P
    assert int.from_bytes(hs.field_op(2, A, B)[0].tobytes(), "big") == (a - b) % P
    assert int.from_bytes(hs.field_op(16, A, B)[0].tobytes(), "big") == (a % N) * (b % N) % N


def test_safegcd_inversion_matches_fermat_and_pow():
    """modinv.cuh (Bernstein-Yang divsteps) against Python's pow and against the reference's Fermat chains
    (field_invert.go:11, scalar_invert.go:11) kept as fe_invert_fermat / sc_invert_fermat: random values plus
    the structured ones a limb-boundary or sign bug would trip over (powers of two, all-ones runs, values
    next to the moduli, non-canonical field inputs), and Invert(0) = 0."""
    rng = np.random.default_rng(2024)
    vals = [int.from_bytes(rng.bytes(32), "big") for _ in range(1500)]
    vals += [2**k for k in range(256)] + [2**k - 1 for k in range(1, 257)]
    vals += [P - 1 - k for k in range(32)] + [N - 1 - k for k in range(32)] + list(range(32)) + [P + k for k in range(32)]
    vals += [(2**30) ** k for k in range(9)] + [(2**30) ** k - 1 for k in range(1, 9)]
    a = rows(vals)
    inv_p, fer_p = hs.field_op(3, a, a), hs.field_op(7, a, a)
    inv_n, fer_n = hs.field_op(18, a, a), hs.field_op(19, a, a)
    for v, ip, fp, i_n, fn in zip(vals, inv_p, fer_p, inv_n, fer_n):
        assert int.from_bytes(ip.tobytes(), "big") == pow(v % P, P - 2, P), hex(v)
        assert ip.tobytes() == fp.tobytes(), hex(v)
        assert int.from_bytes(i_n.tobytes(), "big") == pow(v % N, N - 2, N), hex(v)
        assert i_n.tobytes() == fn.tobytes(), hex(v)


# ---------------------------------------------------------------------------
# The argument that lets the large-batch fixed-base kernel use the INCOMPLETE mixed Jacobian addition
# (csrc/kernels.cuh, item_base_mult_ct_jac): with ascending windows the accumulator A = sum_{v<w} d_v B^v is never
# congruent to +- the entry E = d_w B^w, and "all digits so far are zero" is the same as "A is the identity".
# The recoding below restates the kernel's (ct_window_bits + carry, digits in [-(B/2 - 1), B/2]).
# ---------------------------------------------------------------------------
def _signed_digits(k, wb):
    nw = (257 + wb - 1) // wb
    out, carry = [], 0
    for w in range(nw):
        v = ((k >> (wb * w)) & ((1 << wb) - 1)) + carry
        carry = 1 if v > (1 << (wb - 1)) else 0
        out.append(v - (carry << wb))
    assert carry == 0 and sum(d << (wb * w) for w, d in enumerate(out)) == k
    return out


def test_fixed_base_jacobian_argument_constants():
    c = 2**256 - N
    for wb in (6, 7):
        B, nw = 1 << wb, (257 + wb - 1) // wb
        for w in range(1, nw):
            assert (B // 2) * (B**w - 1) * 100 < 51 * (B - 1) * B**w            # |A| < 0.51 B^w
        assert wb * (nw - 1) == 252                                             # the top window starts at bit 252
        assert (B // 2 + 1) * B**(nw - 2) < 2**252 < N                          # below the top window |A -+ E| < n
        # top window: E = n + A would have to be a multiple of 2^252 with -0.51 * 2^252 < A < 0
        for m in range(0, 64):
            A = c - m * 2**252
            assert not (-51 * 2**252 < 100 * A < 0), m


@settings(**SET)
@given(st.lists(scalar, min_size=1, max_size=8))
def test_fixed_base_jacobian_argument_on_scalars(ks):
    for k in ks:
        k %= N                                                                  # sc_from_be32 reduces
        for wb in (6, 7):
            B, digits = 1 << wb, _signed_digits(k, wb)
            A = 0
            for w, d in enumerate(digits):
                if d != 0:
                    E = d * B**w
                    assert (A - E) % N != 0 and (A + E) % N != 0, (hex(k), wb, w)   # no doubling, no cancellation
                assert (A % N == 0) == all(x == 0 for x in digits[:w]), (hex(k), wb, w)
                A += d * B**w
            assert A == k


def test_fixed_base_jacobian_argument_on_structured_scalars():
    c = 2**256 - N
    ks = set()
    for base in (0, c, 2 * c, N - c, N // 2, 2**252, 15 * 2**252, N - 2**252, (N + c) // 2):
        for d in range(-70, 71):
            ks.add((base + d) % N)
            ks.add((base + d * 2**252) % N)
            ks.add((base + d * 2**245) % N)
    for wb in (6, 7):
        B = 1 << wb
        for w in range(0, (257 + wb - 1) // wb):
            for d in (1, B // 2 - 1, B // 2, B // 2 + 1, B - 1):
                ks.add((d * B**w) % N)
                ks.add((N - d * B**w) % N)
    test_fixed_base_jacobian_argument_on_scalars.hypothesis.inner_test(sorted(ks))


# ---------------------------------------------------------------------------
# The Jacobian formulas of the variable-time ladders (csrc/jac.cuh) against the complete ones (csrc/point.cuh): random
# chains of doublings and additions of +-A and +-B from the identity, with B chosen so that the chain runs into the
# explicit branches -- B = A, B = -A, B = 2A, B = -2A (equal / opposite accumulator and addend after one or two steps).
# ---------------------------------------------------------------------------
def _pt(o, k):
    out, st = o.batch_scalar_base_mult(rows([k % N]))
    assert st[0] == 1
    return out[0]


@settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
@given(st.integers(1, 2**64), st.sampled_from([1, -1, 2, -2, 3, 5, 7, 2**31]), st.lists(st.integers(0, 4), min_size=1, max_size=24))
def test_jacobian_chain_equals_complete_chain(a, rel, ops):
    from oracle import oracle as o
    o.lib()
    A = _pt(o, a)
    B = _pt(o, (a * rel) % N if abs(rel) <= 2 else rel)
    (oj, sj), (oc, sc_) = hs.jac_vs_complete(A, B, ops)
    assert sj == sc_ and np.array_equal(oj, oc), (a, rel, ops)


def test_jacobian_chain_exceptional_branches_are_reached():
    from oracle import oracle as o
    o.lib()
    A, A2, nA2 = _pt(o, 9), _pt(o, 18), _pt(o, N - 18)
    cases = [
        (A, A, [3, 1]),          # A + A: doubling inside the addition
        (A, A, [3, 2]),          # A - A: the identity flag
        (A, A, [3, 2, 1, 0, 3]), # ... then an assignment from the identity, a doubling, an addition
        (A, A2, [3, 0, 1]),      # 2A + 2A
        (A, nA2, [3, 0, 1]),     # 2A - 2A
        (A, A2, [0, 0, 3]),      # doublings of the identity, then an assignment
        (A, A2, [3, 0, 2, 2]),   # 2A - 2A - 2A = -2A
    ]
    for a_, b_, ops in cases:
        (oj, sj), (oc, sc_) = hs.jac_vs_complete(a_, b_, ops)
        assert sj == sc_ and np.array_equal(oj, oc), ops
    (oj, sj), _ = hs.jac_vs_complete(A, A, [3, 2])
    assert sj == 2 and not oj.any()                    # the identity is reported, not encoded
