#!/usr/bin/env python3
"""Generates secp256k1-voi_b200/csrc/fe_mul_gen.cuh: the F_p multiplier and squarer cores in SPLIT form.

Why a second form.  The first multiplier (fe.cuh: fe_mul_wide + fe_reduce_wide_pre) merges its even- and
odd-aligned accumulators into one 16-limb product and then reduces it.  ptxas turns the odd-aligned half of that
reduction into IMAD.WIDE with register-pair addends it has to assemble with IMAD.MOV, and every fresh top limb of a
row costs a SEL plus an IMAD.MOV of zero: 19 non-multiplying instructions per product on the multiplier's own pipe
(profiles/r01_sass_mix_k_dsm.txt).  Here nothing is ever re-paired:

  product   rows as before (two carry chains per row, on different accumulators), but a chain that can carry out
            continues into the NEXT row's top product (a fresh aligned pair with a zero addend), and the other
            chain of that row, three products long, ripples its carry through that pair with two adds;
  reduce    H = hi(E) + hi(O) is formed with adds, H_even * 977 accumulates onto the E pairs and H_odd * 977 onto
            the O pairs (both already aligned), the "<< 32" half of 2^256 = 2^32 + 977 is a limb shift in the last
            two add chains.

64 + 8 wide multiplications, ~40 adds, no register moves by construction.  The squarer doubles E and O separately
by funnel shifts and adds the squares a_i^2 through the multiplier's accumulate input.

The instruction lists are EMULATED below against Python integers (random, extreme and crafted-carry inputs) before
the header is written; there is no GPU on the build box.  Value model: E[k] sits at limb k, O[k] at limb k + 1.
"""
import os
import random

M32 = 0xFFFFFFFF
P = 2**256 - 2**32 - 977


class Stmt:
    """One asm statement: a list of ops sharing one carry flag."""

    def __init__(self):
        self.ops = []

    def op(self, name, dst, *src):
        self.ops.append((name, dst, src))
        return self


class Prog:
    def __init__(self):
        self.stmts = []

    def stmt(self):
        s = Stmt()
        self.stmts.append(s)
        return s

    # ----- emulation --------------------------------------------------------------------------------------------
    def run(self, regs):
        regs = dict(regs)

        def val(x):
            if isinstance(x, int):
                return x
            return regs[x]

        for s in self.stmts:
            cc = 0
            for name, dst, src in s.ops:
                if name == "cc0":  # a carry that is provably zero is threaded into the next chain: check it
                    assert cc == 0, "threaded carry is not zero"
                    continue
                if name in ("mul.lo", "mul.hi"):
                    p = val(src[0]) * val(src[1])
                    v = (p & M32) if name == "mul.lo" else (p >> 32)
                    regs[dst] = v
                    continue
                if name == "shf.l":  # funnel shift left by src[2]: high word of (hi:lo) << k
                    lo, hi, k = val(src[0]), val(src[1]), src[2]
                    regs[dst] = (((hi << 32) | lo) << k >> 32) & M32
                    continue
                if name == "shl":
                    regs[dst] = (val(src[0]) << src[1]) & M32
                    continue
                if name == "shr":
                    regs[dst] = val(src[0]) >> src[1]
                    continue
                base, cin, cout = name, False, False
                if base.endswith(".cc"):
                    cout = True
                    base = base[:-3]
                if base in ("madc.lo", "madc.hi", "addc"):
                    cin = True
                if base in ("mad.lo", "madc.lo"):
                    v = ((val(src[0]) * val(src[1])) & M32) + val(src[2])
                elif base in ("mad.hi", "madc.hi"):
                    v = ((val(src[0]) * val(src[1])) >> 32) + val(src[2])
                elif base in ("add", "addc"):
                    v = val(src[0]) + val(src[1])
                else:
                    raise ValueError(name)
                if cin:
                    v += cc
                regs[dst] = v & M32
                if cout:
                    cc = v >> 32
                else:
                    assert v >> 32 == 0, ("carry lost", name, dst)
                    cc = 0
            # a statement may end on a .cc op only if its carry is provably zero
            assert cc == 0, "carry lost at the end of a statement"
        return regs

    # ----- C++ emission -----------------------------------------------------------------------------------------
    def emit(self, cname):
        """cname(reg) -> C++ lvalue for a register name."""
        out = []
        for s in self.stmts:
            written, read_first = [], []
            seen_w = set()
            for name, dst, src in s.ops:
                if name == "cc0":
                    continue
                for x in src:
                    if isinstance(x, str) and x not in seen_w and x not in read_first:
                        read_first.append(x)
                if dst not in seen_w:
                    seen_w.add(dst)
                    written.append(dst)
            outs = [(("+r" if d in read_first else "=r"), d) for d in written]
            ins = [x for x in read_first if x not in seen_w]
            idx = {d: n for n, (_, d) in enumerate(outs)}
            for n, x in enumerate(ins):
                idx[x] = len(outs) + n

            def ref(x):
                return str(x) if isinstance(x, int) else "%" + str(idx[x])

            lines = []
            for name, dst, src in s.ops:
                if name == "cc0":
                    continue
                if name == "shf.l":
                    lines.append(f"shf.l.wrap.b32 {ref(dst)},{ref(src[0])},{ref(src[1])},{src[2]};")
                elif name == "shl":
                    lines.append(f"shl.b32 {ref(dst)},{ref(src[0])},{src[1]};")
                elif name == "shr":
                    lines.append(f"shr.u32 {ref(dst)},{ref(src[0])},{src[1]};")
                else:
                    lines.append(f"{name}.u32 {ref(dst)}," + ",".join(ref(x) for x in src) + ";")
            # early-clobber: an output written before all inputs are read must not share a register with an input
            o = ", ".join(f'"{("=&r" if k == "=r" else k)}"({cname(d)})' for k, d in outs)
            i = ", ".join(f'"r"({cname(x)})' for x in ins)
            body = ' "\n        "'.join(" ".join(lines[k:k + 4]) for k in range(0, len(lines), 4))
            out.append(f'    asm("{body}"\n        : {o}\n        : {i});')
        return "\n".join(out)


def mulw(s, lo, hi, x, y):
    s.op("mul.lo", lo, x, y).op("mul.hi", hi, x, y)


def acc_name(pos):
    """(array, index) of the aligned pair that holds a product at limb position pos."""
    return ("e", pos) if pos % 2 == 0 else ("o", pos - 1)


def build_product(p, A, B, mode="defer"):
    """E, O <- A * B (A, B: lists of 8 register names).  E[k] at limb k, O[k] at limb k + 1.

    Row i adds a_j * b_i at limb i + j: the even j form chain A(i) on one array, the odd j chain B(i) on the other.
    A chain whose last pair is live can carry out; the modes differ in where that carry goes:

    "defer"   B(i) ends on its own top product (a fresh pair: zero addend, absorbs the carry); A(i) runs over four
              live pairs and its carry c_i (limb i + 8) is captured with a select and added when H is formed.
              Pure wavefront: the k-th product of a chain only waits for the k-th product of the chain before it.
    "ripple"  A(i) continues into row i + 1's top product (fresh pair), B(i + 1) is three products long and ripples
              its carry through that pair with two adds (the second is a final add: IMAD.X on the multiplier's pipe).
    "thread"  as "ripple", but the ripple's provably-zero carry is threaded into the next chain on the same array, so
              every add has a consumer for its carry and stays on the ALU pipe; this serialises the chains of an
              array (measured: the ladders are latency-bound, so this form lost to the wavefront forms).
    Returns (statement to thread the H chain into, or None; list of deferred carries [(name, H index)])."""
    live = set()
    thread = mode == "thread"

    def pair(pos):
        arr, k = acc_name(pos)
        return f"{arr}{k}", f"{arr}{k + 1}"

    # row 0: all pairs fresh
    s = p.stmt()
    for j in range(8):
        lo, hi = pair(j)
        mulw(s, lo, hi, A[j], B[0])
        live.update((lo, hi))
    open_stmt = {"e": None, "o": None}  # statement whose last op left a provably-zero carry to thread
    deferred = []

    def chain_a(i):
        arr = acc_name(i)[0]
        threaded = thread and open_stmt[arr] is not None
        s = open_stmt[arr] if threaded else p.stmt()
        open_stmt[arr] = None
        first = True
        for j in (0, 2, 4, 6):
            lo, hi = pair(i + j)
            assert lo in live and hi in live, (i, j)
            s.op(("madc.lo.cc" if threaded else "mad.lo.cc") if first else "madc.lo.cc", lo, A[j], B[i], lo)
            s.op("madc.hi.cc", hi, A[j], B[i], hi)
            first = False
        if mode == "defer":
            s.op("addc", f"c{i}", 0, 0)
            deferred.append((f"c{i}", i))
        elif i < 7:
            lo, hi = pair(i + 1 + 7)
            assert lo not in live and hi not in live
            s.op("madc.lo.cc", lo, A[7], B[i + 1], 0)
            s.op("madc.hi", hi, A[7], B[i + 1], 0)
            live.update((lo, hi))
        else:
            k = acc_name(i + 6)[1]
            top = f"{arr}{k + 2}"
            assert top not in live
            s.op("addc", top, 0, 0)
            live.add(top)

    def chain_b(i):
        arr = acc_name(i + 1)[0]
        s = p.stmt()
        first = True
        js = (1, 3, 5, 7) if (i == 1 or mode == "defer") else (1, 3, 5)
        for j in js:
            lo, hi = pair(i + j)
            if lo not in live:
                assert j == 7
                s.op("madc.lo.cc", lo, A[j], B[i], 0)
                s.op("madc.hi", hi, A[j], B[i], 0)
                live.update((lo, hi))
            else:
                s.op("mad.lo.cc" if first else "madc.lo.cc", lo, A[j], B[i], lo)
                s.op("madc.hi.cc", hi, A[j], B[i], hi)
            first = False
        if len(js) == 3:
            lo, hi = pair(i + 7)
            assert lo in live and hi in live
            s.op("addc.cc", lo, lo, 0)
            if thread:
                s.op("addc.cc", hi, hi, 0)
                s.op("cc0", None)
                open_stmt[arr] = s
            else:
                s.op("addc", hi, hi, 0)
        return s

    last_e = None
    for i in range(1, 8):
        chain_a(i)
        sb = chain_b(i)
        if acc_name(i + 1)[0] == "e":
            last_e = sb
    return (last_e if thread else None), deferred


def build_reduce(p, after=None, deferred=(), thread_tail=False, o14=True):
    """(E[0..15], O[0..14]) -> r0..r7, t8, t9 with value = sum r_k 2^(32k) + 2^256 (t8 + 2^32 t9) (mod p).
    `after`: a statement that ended on a provably-zero carry (threaded into the H chain, mode "thread").
    `deferred`: carries (name, k) still owed to limb 8 + k.  `o14`: False if O has no index 14 (mode "defer")."""
    # deferred carries first: G = hi(E) + C (limbs 9..15)
    hi_e = {k: f"e{8 + k}" for k in range(8)}
    if deferred:
        cs = dict((k, n) for n, k in deferred)
        ks = sorted(cs)
        assert ks == list(range(ks[0], 8))
        s = p.stmt()
        for k in ks:
            name = "add.cc" if k == ks[0] else ("addc.cc" if k < 7 else "addc")
            s.op(name, f"g{k}", f"e{8 + k}", cs[k])
            hi_e[k] = f"g{k}"
    # H = hi(E) + hi(O): h_k at limb 8 + k; the product is < 2^512, so no carry leaves h7
    s = after if after is not None else p.stmt()
    for k in range(8):
        name = ("addc.cc" if after is not None else "add.cc") if k == 0 else ("addc.cc" if (k < 7 or thread_tail) else "addc")
        s.op(name, f"h{k}", hi_e[k], f"o{7 + k}" if (k < 7 or o14) else 0)
    if thread_tail:
        s.op("cc0", None)
    else:
        s = p.stmt()
    # E pairs += H_even * 977; the carry out of limb 7 is parked in o7 (limb 8), which H has just absorbed
    for n, k in enumerate((0, 2, 4, 6)):
        s.op("madc.lo.cc" if (thread_tail or n) else "mad.lo.cc", f"e{k}", f"h{k}", 977, f"e{k}")
        s.op("madc.hi.cc", f"e{k + 1}", f"h{k}", 977, f"e{k + 1}")
    s.op("addc", "o7", 0, 0)
    # O pairs += H_odd * 977 (pairs at limbs 1..8); o7 <= 1 + 976 + 1: no carry out
    s = p.stmt()
    for n, k in enumerate((1, 3, 5, 7)):
        s.op("mad.lo.cc" if n == 0 else "madc.lo.cc", f"o{k - 1}", f"h{k}", 977, f"o{k - 1}")
        s.op("madc.hi.cc" if k < 7 else "madc.hi", f"o{k}", f"h{k}", 977, f"o{k}")
    # r = E + (O << 32) + (H << 32): two add chains over limbs 1..8, the second one's carry is t9
    s = p.stmt()
    for k in range(1, 9):
        a = f"e{k}" if k < 8 else 0
        s.op("add.cc" if k == 1 else ("addc.cc" if (k < 8 or thread_tail) else "addc"), f"s{k}", a, f"o{k - 1}")
    if thread_tail:
        s.op("cc0", None)  # o7 <= 1 + 976 + 1: limb 8 cannot carry
    else:
        s = p.stmt()
    for k in range(1, 9):
        s.op("addc.cc" if (thread_tail or k > 1) else "add.cc", f"s{k}", f"s{k}", f"h{k - 1}")
    s.op("addc", "t9", 0, 0)


def build_product_acc(p, C, D):
    """E, O += C * D on top of a finished first product (every pair is live now, so every chain can carry out).
    The carries are captured with selects (cA_i at limb i + 8, cB_i at limb i + 9) and added when H is formed -- a ripple
    through all the live limbs above would put up to eight dependent adds behind every chain.
    Returns {limb: [carry names]} for limbs 8..16."""
    def pair(pos):
        arr, k = acc_name(pos)
        return f"{arr}{k}", f"{arr}{k + 1}"

    owed = {}
    for i in range(8):
        for kind, js, nm in (("A", (0, 2, 4, 6), f"ca{i}"), ("B", (1, 3, 5, 7), f"cb{i}")):
            s = p.stmt()
            for n, j in enumerate(js):
                lo, hi = pair(i + j)
                s.op("mad.lo.cc" if n == 0 else "madc.lo.cc", lo, C[j], D[i], lo)
                s.op("madc.hi.cc", hi, C[j], D[i], hi)
            s.op("addc", nm, 0, 0)
            owed.setdefault(i + js[-1] + 2, []).append(nm)
    return owed


def build_reduce2(p, owed):
    """Reduction of a SUM of two products: H = hi(E) + hi(O) + owed carries has nine limbs (h8 <= 1 since the sum is below
    2^513); h0..h7 go through the same chains as for one product, h8 * 2^256 * delta is added to the top word pair
    (t8, t9), which the fold handles for any small t9."""
    # The owed carries are never summed with each other (ptxas turns a sum of two captured flags into predicated constant
    # moves on the multiplier's pipe): the A-chain carries (one per limb 8..15) join hi(E), the B-chain carries (one per
    # limb 9..16) join hi(O), and the two sums are added.
    ca = {limb: [n for n in names if n.startswith("ca")] for limb, names in owed.items()}
    cb = {limb: [n for n in names if n.startswith("cb")] for limb, names in owed.items()}
    assert all(len(ca.get(8 + k, [])) == 1 for k in range(8)) and all(len(cb.get(9 + k, [])) == 1 for k in range(8))
    s = p.stmt()
    for k in range(8):
        s.op("add.cc" if k == 0 else "addc.cc", f"g{k}", f"e{8 + k}", ca[8 + k][0])
    s.op("addc", "g8", 0, 0)
    s = p.stmt()
    for k in range(1, 8):
        s.op("add.cc" if k == 1 else "addc.cc", f"q{k}", f"o{7 + k}", cb[8 + k][0])
    s.op("addc", "q8", cb[16][0], 0)
    # H = G + Q (q0 = o7)
    s = p.stmt()
    for k in range(8):
        s.op("add.cc" if k == 0 else "addc.cc", f"h{k}", f"g{k}", "o7" if k == 0 else f"q{k}")
    s.op("addc", "h8", "g8", "q8")
    # from here on as for one product
    s = p.stmt()
    for n, k in enumerate((0, 2, 4, 6)):
        s.op("madc.lo.cc" if n else "mad.lo.cc", f"e{k}", f"h{k}", 977, f"e{k}")
        s.op("madc.hi.cc", f"e{k + 1}", f"h{k}", 977, f"e{k + 1}")
    s.op("addc", "o7", 0, 0)
    s = p.stmt()
    for n, k in enumerate((1, 3, 5, 7)):
        s.op("mad.lo.cc" if n == 0 else "madc.lo.cc", f"o{k - 1}", f"h{k}", 977, f"o{k - 1}")
        s.op("madc.hi.cc" if k < 7 else "madc.hi", f"o{k}", f"h{k}", 977, f"o{k}")
    s = p.stmt()
    for k in range(1, 9):
        a = f"e{k}" if k < 8 else 0
        s.op("add.cc" if k == 1 else ("addc.cc" if k < 8 else "addc"), f"s{k}", a, f"o{k - 1}")
    s = p.stmt()
    for k in range(1, 9):
        s.op("addc.cc" if k > 1 else "add.cc", f"s{k}", f"s{k}", f"h{k - 1}")
    s.op("addc", "t9", 0, 0)
    # + h8 * delta on (t8, t9): h8 * 977 <= 977, h8 <= 1
    s = p.stmt()
    s.op("mad.lo.cc", "s8", "h8", 977, "s8")
    s.op("addc", "t9", "t9", "h8")


def build_square(p, A):
    """E, O <- cross = sum_{i<j} a_i a_j 2^(32(i+j)) in split form, with no carry ever leaving a chain.

    Position q holds min(q, 14 - q) // 2 + (1 if ...) products; they are dealt to nested chains, innermost first, so
    that every chain starts and ends on a FRESH pair (zero addend, absorbs the carry) and runs over live pairs in
    between: O: [7], [5..9], [3..11], [1..13]; E: [6..8], [4..10], [2..12].  28 wide multiplications, nothing else."""
    live = set()

    def pair(pos):
        arr, k = acc_name(pos)
        return f"{arr}{k}", f"{arr}{k + 1}"

    by_pos = {}
    for i in range(8):
        for j in range(i + 1, 8):
            by_pos.setdefault(i + j, []).append((i, j))
    for parity, centre in ((1, 7), (0, 7)):
        positions = sorted(q for q in by_pos if q % 2 == parity)
        nchains = max(len(by_pos[q]) for q in positions)
        # chain c (0 = innermost) covers the positions that still have a product left when it is formed
        for c in range(nchains):
            span = [q for q in positions if by_pos[q]]
            # innermost first: the positions with the most products left
            most = max(len(by_pos[q]) for q in span)
            span = [q for q in span if len(by_pos[q]) == most] if c == 0 else \
                   [q for q in span if len(by_pos[q]) >= nchains - c]
            assert span == list(range(span[0], span[-1] + 1, 2)), span
            s = p.stmt()
            for n, q in enumerate(span):
                i, j = by_pos[q].pop()
                lo, hi = pair(q)
                fresh = lo not in live
                assert fresh == (n == 0 or n == len(span) - 1), (span, q)
                last = n == len(span) - 1
                s.op("mad.lo.cc" if n == 0 else "madc.lo.cc", lo, A[i], A[j], 0 if fresh else lo)
                s.op("madc.hi" if last else "madc.hi.cc", hi, A[i], A[j], 0 if fresh else hi)
                live.update((lo, hi))
    assert all(not v for v in by_pos.values())
    return live


def double_and_squares(p, live, A, thread=False):
    """E <- 2E + sum a_i^2 2^(64 i), O <- 2O.  Cross sums: E limbs 2..13, O idx 0..13 (cross < 2^511).
    Returns the statement that ends on a provably-zero carry (threaded into the H chain)."""
    assert sorted(int(x[1:]) for x in live if x[0] == "e") == list(range(2, 14))
    assert sorted(int(x[1:]) for x in live if x[0] == "o") == list(range(0, 14))
    # 2 * O: o14 receives the bit shifted out of o13
    s = p.stmt()
    s.op("shr", "o14", "o13", 31)
    for k in range(13, 0, -1):
        s.op("shf.l", f"o{k}", f"o{k - 1}", f"o{k}", 1)
    s.op("shf.l", "o0", 0, "o0", 1)
    # 2 * E over limbs 2..13; the bit shifted out of e13 (limb 14) joins a_7^2 below
    s = p.stmt()
    s.op("shr", "x14", "e13", 31)
    for k in range(13, 2, -1):
        s.op("shf.l", f"e{k}", f"e{k - 1}", f"e{k}", 1)
    s.op("shf.l", "e2", 0, "e2", 1)
    # a_7^2 apart (fresh pair), so that no pair with one live and one fresh half is ever needed
    s = p.stmt()
    mulw(s, "y14", "y15", A[7], A[7])
    # squares: pair (2i, 2i+1) += a_i^2, one carry chain from limb 0 to limb 13, then limbs 14, 15 by adds
    s = p.stmt()
    for i in range(7):
        lo, hi = 2 * i, 2 * i + 1
        s.op("mad.lo.cc" if i == 0 else "madc.lo.cc", f"e{lo}", A[i], A[i], f"e{lo}" if lo >= 2 else 0)
        s.op("madc.hi.cc", f"e{hi}", A[i], A[i], f"e{hi}" if lo >= 2 else 0)
    s.op("addc.cc", "e14", "x14", "y14")
    if thread:
        s.op("addc.cc", "e15", "y15", 0)
        s.op("cc0", None)
    else:
        s.op("addc", "e15", "y15", 0)
    return s


def value_split(regs):
    v = 0
    for k in range(16):
        v += regs.get(f"e{k}", 0) << (32 * k)
    for k in range(15):
        v += regs.get(f"o{k}", 0) << (32 * (k + 1))
    return v


def value_reduced(regs):
    v = regs["e0"]
    for k in range(1, 8):
        v += regs[f"s{k}"] << (32 * k)
    return v + ((regs["s8"] + (regs["t9"] << 32)) << 256)


def limbs(x):
    return [(x >> (32 * k)) & M32 for k in range(8)]


def test_inputs(trials):
    rnd = random.Random(7)
    special = [0, 1, P - 1, P, P + 1, 2**256 - 1, 2**256 - 2**32, 2**255, 2**32 - 1, 2**224 - 1, (2**256 - 1) ^ (M32 << 96)]
    cases = [(x, y) for x in special for y in special]
    for _ in range(trials):
        pick = lambda: [rnd.choice([0, 1, M32, M32 - 1, 0x80000000, rnd.getrandbits(32)]) for _ in range(8)]
        cases.append((sum(v << (32 * k) for k, v in enumerate(pick())), sum(v << (32 * k) for k, v in enumerate(pick()))))
    for _ in range(trials):
        cases.append((rnd.getrandbits(256), rnd.getrandbits(256)))
    return cases


def main():
    A = [f"a{k}" for k in range(8)]
    B = [f"b{k}" for k in range(8)]
    # Two forms of each core, same arithmetic, different carry plumbing (see build_product):
    #   suffix _w ("wavefront"): mode "ripple", nothing threaded -- the shared out-of-line body of the variable-time
    #       ladders, which are bound by dependent latency (measured 21.55 ms against 21.70 threaded / 21.72 deferred
    #       / 22.09 for the merged round-1 form, k_dsm at 2^20);
    #   no suffix: mode "thread", tail threaded -- the constant-time kernels, where the multiplier is inlined and the
    #       instruction count on the multiplier's pipe decides (fixed-base mult 153.0 M/s against 150.1 / 145.5).
    forms = {"": ("thread", True), "_w": (os.environ.get("S256_GEN_MODE_W", "ripple"), False)}
    cases = test_inputs(3000)
    top_max = 0
    progs = {}
    for suffix, (mode, tail) in forms.items():
        mul = Prog()
        last_e, deferred = build_product(mul, A, B, mode)
        build_reduce(mul, last_e, deferred, thread_tail=tail, o14=(mode != "defer"))
        sqr = Prog()
        live = build_square(sqr, A)
        last = double_and_squares(sqr, live, A, thread=tail)
        build_reduce(sqr, last if tail else None, thread_tail=tail)
        sqr_only = Prog()
        double_and_squares(sqr_only, build_square(sqr_only, A), A)
        prod_only = Prog()
        build_product(prod_only, A, B, mode)
        for x, y in cases:
            regs = {f"a{k}": v for k, v in enumerate(limbs(x))}
            regs.update({f"b{k}": v for k, v in enumerate(limbs(y))})
            pr = prod_only.run(regs)  # product phase alone (deferred carries are owed to limb i + 8)
            assert value_split(pr) + sum(pr.get(f"c{i}", 0) << (32 * (i + 8)) for i in range(1, 8)) == x * y, (hex(x), hex(y))
            out = mul.run(regs)
            assert value_reduced(out) % P == (x * y) % P
            assert out["t9"] <= 1
            top_max = max(top_max, out["s8"] + (out["t9"] << 32))
            assert value_split(sqr_only.run(regs)) == x * x, hex(x)
            out = sqr.run(regs)
            assert value_reduced(out) % P == (x * x) % P
            assert out["t9"] <= 1
        progs[suffix] = (mul, sqr, mode)

    # a * b + c * d with one reduction (wavefront form only: the shared body the variable-time formulas call)
    Cn = [f"c{k}" for k in range(8)]
    Dn = [f"d{k}" for k in range(8)]
    fused = Prog()
    build_product(fused, A, B, "ripple")
    owed = build_product_acc(fused, Cn, Dn)
    build_reduce2(fused, owed)
    rnd = random.Random(11)
    t9_max = 0
    for n_, (x, y) in enumerate(cases):
        z, w = cases[(n_ * 7 + 3) % len(cases)]
        if n_ % 3 == 0:
            z, w = rnd.getrandbits(256), rnd.getrandbits(256)
        if n_ % 50 == 1:
            x = y = z = w = 2**256 - 1
        regs = {f"a{k}": v for k, v in enumerate(limbs(x))}
        regs.update({f"b{k}": v for k, v in enumerate(limbs(y))})
        regs.update({f"c{k}": v for k, v in enumerate(limbs(z))})
        regs.update({f"d{k}": v for k, v in enumerate(limbs(w))})
        out = fused.run(regs)
        assert value_reduced(out) % P == (x * y + z * w) % P, (hex(x), hex(y), hex(z), hex(w))
        assert out["h8"] <= 1
        t9_max = max(t9_max, out["t9"])
    assert t9_max <= 3

    def cn(prefix_map):
        def f(r):
            for pre, fmt in prefix_map.items():
                if r.startswith(pre) and r[len(pre):].isdigit():
                    return fmt.format(int(r[len(pre):]))
            return r
        return f

    names = cn({"a": "a[{}]", "b": "b[{}]", "s": "r[{}]"})

    def wide(nm):
        return sum(1 for s in nm.stmts for o in s.ops if o[0] in ("mul.hi", "madc.hi.cc", "madc.hi", "mad.hi"))

    def alu(nm):
        return sum(1 for s in nm.stmts for o in s.ops if o[0].startswith(("add", "sh")))

    decl_e = ", ".join(f"e{k}" for k in range(16))
    decl_o = ", ".join(f"o{k}" for k in range(15))
    decl_h = ", ".join(f"h{k}" for k in range(8))
    mul0, sqr0, _ = progs[""]
    hdr = f'''// fe_mul_gen.cuh -- GENERATED by tools/gen_fe_mul.py; do not edit.
// F_p multiplier / squarer cores in split form (even- and odd-aligned accumulators are reduced without ever being
// merged or re-paired): {wide(mul0)} wide multiplications + {alu(mul0)} adds per product, {wide(sqr0)} + {alu(sqr0)} (adds and shifts) per square.
// Two forms with the same arithmetic: fe_mul_core / fe_sqr_core (carries threaded: fewest instructions on the multiplier's
// pipe, for inlined constant-time code) and fe_mul_core_w / fe_sqr_core_w (wavefront: shortest dependent chains, for the
// shared body the latency-bound ladders call).  Every instruction list below was emulated against Python integers on
// {len(cases)} inputs before emission.
// Output: r[0..7] + 2^256 * (r[8] + 2^32 * t9) is congruent to the product mod p, t9 in {{0, 1}}.
#pragma once
#include <stdint.h>

namespace s256 {{
'''
    for suffix, (mul, sqr, mode) in progs.items():
        extra = "\n    uint32_t c1, c2, c3, c4, c5, c6, c7, g1, g2, g3, g4, g5, g6, g7;" if mode == "defer" else ""
        hdr += f'''
__device__ __forceinline__ void fe_mul_core{suffix}(uint32_t r[9], uint32_t &t9, const uint32_t a[8], const uint32_t b[8]) {{
    uint32_t {decl_e};
    uint32_t {decl_o};
    uint32_t {decl_h};{extra}
{mul.emit(names)}
    r[0] = e0;
}}

__device__ __forceinline__ void fe_sqr_core{suffix}(uint32_t r[9], uint32_t &t9, const uint32_t a[8]) {{
    uint32_t {decl_e};
    uint32_t {decl_o};
    uint32_t {decl_h}, x14, y14, y15;
{sqr.emit(names)}
    r[0] = e0;
}}
'''
    names4 = cn({"a": "a[{}]", "b": "b[{}]", "c": "c[{}]", "d": "d[{}]", "s": "r[{}]"})
    # register names of the fused program that collide with the operand prefixes are renamed for emission
    def ren(r):
        if r is None or isinstance(r, int):
            return r
        if r[0] == "c" and r[1] in "ab":        # captured carries ca0.. / cb0..
            return "k" + r[1:]
        if r[0] == "d" and r[1:].isdigit() and int(r[1:]) >= 8:   # owed sums d9..d15 (d0..d7 are operand limbs)
            return "w" + r[1:]
        return r
    for st in fused.stmts:
        st.ops = [(nm, ren(dst), tuple(ren(x) for x in src)) for nm, dst, src in st.ops]
    decl_k = ", ".join([f"ka{i}" for i in range(8)] + [f"kb{i}" for i in range(8)])
    decl_w = ", ".join(f"q{i}" for i in range(1, 9))
    decl_g = ", ".join(f"g{i}" for i in range(9))
    hdr += f'''
// r + 2^256 (r[8] + 2^32 t9) == a * b + c * d (mod p) with ONE reduction ({wide(fused)} wide multiplications instead of 2 x {wide(mul0)}), t9 <= 3.
// The complete formulas end in three such sums (X3, Y3, Z3); fewer calls and fewer reduction tails is what the
// latency-bound ladders gain from it.
__device__ __forceinline__ void fe_mul2add_core_w(uint32_t r[9], uint32_t &t9, const uint32_t a[8], const uint32_t b[8],
                                                  const uint32_t c[8], const uint32_t d[8]) {{
    uint32_t {decl_e};
    uint32_t {decl_o};
    uint32_t {decl_h}, h8;
    uint32_t {decl_k};
    uint32_t {decl_w}, {decl_g};
{fused.emit(names4)}
    r[0] = e0;
}}
'''
    hdr += "\n}  // namespace s256\n"
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "secp256k1-voi_b200", "csrc", "fe_mul_gen.cuh")
    open(path, "w").write(hdr)
    print(f"wrote {path}: mul {wide(mul0)} wide + {alu(mul0)} alu, sqr {wide(sqr0)} wide + {alu(sqr0)} alu; "
          f"{len(progs)} forms emulated on {len(cases)} inputs each; max top word seen {top_max:#x}")


if __name__ == "__main__":
    main()
