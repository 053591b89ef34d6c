#!/usr/bin/env python3
"""Generates secp256k1-voi_b200/csrc/fe_sqr_gen.cuh: the 8-limb squaring.

a^2 = 2 * sum_{i<j} a_i a_j 2^(32(i+j)) + sum_i a_i^2 2^(64 i):
28 cross products + 8 squares = 36 IMAD.WIDE.U32 instead of 64.  The cross
products are accumulated, like fe_mul_wide, into an even-aligned and an
odd-aligned accumulator with mad.lo.cc / madc.hi.cc chains; rows have different
lengths, so the generator tracks which limbs are live, threads the carry of
every chain through the live limbs above it, and emits one self-contained asm
statement per chain.  The same instruction list is EMULATED here on random and
extreme inputs against Python integers before the header is written (there is
no GPU on the build box)."""
import os
import random

M32 = 0xFFFFFFFF


class Prog:
    def __init__(self):
        self.stmts = []      # list of (asm lines, outputs(+r/=r), inputs)
        self.init = {"e": set(), "o": set()}

    # ---- emission helpers: every statement is a list of ops for the emulator too
    def chain(self, acc, base, pairs):
        """acc[base + 2k, base + 2k + 1] += a[i] * a[j] for k, (i, j) in enumerate(pairs)."""
        ops = []
        init = self.init[acc]
        first = True
        last_hi_was_live = False
        for k, (i, j) in enumerate(pairs):
            lo, hi = base + 2 * k, base + 2 * k + 1
            ops.append(("mad.lo.cc" if first else "madc.lo.cc", (acc, lo), i, j, (acc, lo) if lo in init else None))
            ops.append(("madc.hi.cc", (acc, hi), i, j, (acc, hi) if hi in init else None))
            last_hi_was_live = hi in init
            init.add(lo)
            init.add(hi)
            first = False
        nxt = base + 2 * len(pairs)
        if last_hi_was_live:
            # a carry may leave the chain: thread it through the live limbs above, park it in the first free one
            while nxt < 16 and nxt in init:
                ops.append(("addc.cc", (acc, nxt), None, None, (acc, nxt)))
                nxt += 1
            if nxt < 16:
                ops.append(("addc", (acc, nxt), None, None, None))
                init.add(nxt)
        self.stmts.append(ops)

    # ---- emulation
    def run(self, a):
        regs = {}
        for ops in self.stmts:
            cc = 0
            for op, dst, i, j, addend in ops:
                add = regs[addend] if addend is not None else 0
                if op in ("mad.lo.cc", "madc.lo.cc"):
                    v = ((a[i] * a[j]) & M32) + add + (cc if op.startswith("madc") else 0)
                elif op == "madc.hi.cc":
                    v = ((a[i] * a[j]) >> 32) + add + cc
                elif op in ("addc.cc", "addc"):
                    v = add + cc
                else:
                    raise ValueError(op)
                regs[dst] = v & M32
                cc = v >> 32
                if op == "addc":
                    assert cc == 0
            assert cc == 0, "carry lost at the end of a chain"
        return regs

    # ---- C++ emission
    def emit(self):
        out = []
        for ops in self.stmts:
            outs, ins, lines = [], [], []

            def ref(kind, name):
                lst = outs if kind in ("+r", "=r") else ins
                for n, (k, nm) in enumerate(lst):
                    if nm == name:
                        return n if lst is outs else None
                lst.append((kind, name))
                return None
            # first pass: classify destination limbs (read-modify-write => "+r", fresh => "=r")
            dst_kind = {}
            for op, dst, i, j, addend in ops:
                nm = f"{dst[0]}{dst[1]}"
                if addend is not None and addend == dst:
                    dst_kind[nm] = "+r"
                else:
                    dst_kind.setdefault(nm, "=r")
            names_out = list(dst_kind.keys())
            names_in = []
            for op, dst, i, j, addend in ops:
                for x in (i, j):
                    if x is not None and f"a[{x}]" not in names_in:
                        names_in.append(f"a[{x}]")
            idx = {nm: n for n, nm in enumerate(names_out)}
            for n, nm in enumerate(names_in):
                idx[nm] = len(names_out) + n
            for op, dst, i, j, addend in ops:
                d = f"%{idx[f'{dst[0]}{dst[1]}']}"
                ad = f"%{idx[f'{addend[0]}{addend[1]}']}" if addend is not None else "0"
                if op.startswith("mad"):
                    lines.append(f"{op}.u32 {d},%{idx[f'a[{i}]']},%{idx[f'a[{j}]']},{ad};")
                elif op == "addc.cc":
                    lines.append(f"addc.cc.u32 {d},{ad},0;")
                else:
                    lines.append(f"addc.u32 {d},0,0;")
            o = ", ".join(f'"{dst_kind[nm]}"({nm})' for nm in names_out)
            i_ = ", ".join(f'"r"({nm})' for nm in names_in)
            out.append('    asm("' + " ".join(lines) + '"\n        : ' + o + "\n        : " + i_ + ");")
        return "\n".join(out)


def build():
    p = Prog()
    for i in range(7):
        ev = [(i, j) for j in range(i + 2, 8, 2)]   # i + j even  -> E at limb i + j
        od = [(i, j) for j in range(i + 1, 8, 2)]   # i + j odd   -> O at limb i + j, index i + j - 1
        if od:
            p.chain("o", i + od[0][1] - 1, od)
        if ev:
            p.chain("e", i + ev[0][1], ev)
    return p


def reference_check(p, trials=2000):
    rnd = random.Random(1)
    cases = [[M32] * 8, [0] * 8, [1] + [0] * 7, [M32] + [0] * 7, [0] * 7 + [M32]]
    cases += [[rnd.choice([0, 1, M32, M32 - 1, rnd.getrandbits(32)]) for _ in range(8)] for _ in range(trials)]
    for a in cases:
        regs = p.run(a)
        cross = 0
        for (acc, k), v in regs.items():
            cross += v << (32 * (k if acc == "e" else k + 1))
        want = sum(a[i] * a[j] << (32 * (i + j)) for i in range(8) for j in range(i + 1, 8))
        assert cross == want, a
    return len(cases)


HEADER = '''// fe_sqr_gen.cuh -- GENERATED by tools/gen_fe_sqr.py; do not edit.
// 8-limb squaring: 28 cross products (even / odd aligned carry chains), doubled,
// plus 8 squares = 36 IMAD.WIDE.U32 (fe_mul_wide needs 64).  The instruction list
// below was emulated against Python integers on {n} inputs before emission.
#pragma once
#include <stdint.h>

namespace s256 {{

__device__ __forceinline__ void fe_sqr_wide(uint32_t r[16], const uint32_t a[8]) {{
    uint32_t {decl};
{chains}
    // live limbs: e{elive}, o{olive} (o[k] sits at limb k + 1)
    // cross = e + (o << 32)
    uint32_t c[16];
{merge}
    // 2 * cross (cross < 2^511)
    uint32_t d[16];
    d[0] = c[0] << 1;
#pragma unroll
    for (int k = 1; k < 16; k++) asm("shf.l.wrap.b32 %0,%1,%2,1;" : "=r"(d[k]) : "r"(c[k - 1]), "r"(c[k]));
    // + squares a_i^2 at limb 2i
    uint32_t s[16];
#pragma unroll
    for (int k = 0; k < 8; k++) asm("mul.lo.u32 %0,%2,%2; mul.hi.u32 %1,%2,%2;" : "=r"(s[2 * k]), "=r"(s[2 * k + 1]) : "r"(a[k]));
    asm("add.cc.u32 %0,%16,%32; addc.cc.u32 %1,%17,%33; addc.cc.u32 %2,%18,%34; addc.cc.u32 %3,%19,%35;"
        "addc.cc.u32 %4,%20,%36; addc.cc.u32 %5,%21,%37; addc.cc.u32 %6,%22,%38; addc.cc.u32 %7,%23,%39;"
        "addc.cc.u32 %8,%24,%40; addc.cc.u32 %9,%25,%41; addc.cc.u32 %10,%26,%42; addc.cc.u32 %11,%27,%43;"
        "addc.cc.u32 %12,%28,%44; addc.cc.u32 %13,%29,%45; addc.cc.u32 %14,%30,%46; addc.u32 %15,%31,%47;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3]), "r"(d[4]), "r"(d[5]), "r"(d[6]), "r"(d[7]),
          "r"(d[8]), "r"(d[9]), "r"(d[10]), "r"(d[11]), "r"(d[12]), "r"(d[13]), "r"(d[14]), "r"(d[15]),
          "r"(s[0]), "r"(s[1]), "r"(s[2]), "r"(s[3]), "r"(s[4]), "r"(s[5]), "r"(s[6]), "r"(s[7]),
          "r"(s[8]), "r"(s[9]), "r"(s[10]), "r"(s[11]), "r"(s[12]), "r"(s[13]), "r"(s[14]), "r"(s[15]));
}}

}}  // namespace s256
'''


def merge_code(p):
    """c[0..15] = e + (o << 32) over the live limbs only, as one carry chain."""
    e, o = p.init["e"], p.init["o"]
    lines, first = [], True
    ins = []
    for k in range(16):
        terms = []
        if k in e:
            terms.append(f"e{k}")
        if k - 1 in o:
            terms.append(f"o{k - 1}")
        if first:
            if len(terms) == 2:
                lines.append((k, "add.cc.u32", terms))
                first = False
            else:
                # no carry can exist yet: plain copy
                lines.append((k, "mov", terms))
        else:
            lines.append((k, "addc.cc.u32" if k < 15 else "addc.u32", terms))
    # emulate-time sanity is covered by reference_check on e/o; emit C++
    out = []
    asm_ops, outs, inputs = [], [], []
    for k, op, terms in lines:
        if op == "mov":
            out.append(f"    c[{k}] = {terms[0] if terms else '0u'};")
            continue
        a = terms[0] if len(terms) > 0 else None
        b = terms[1] if len(terms) > 1 else None
        outs.append(f"c[{k}]")
        def reg(x):
            if x is None:
                return "0"
            if x not in inputs:
                inputs.append(x)
            return "IN" + str(inputs.index(x))
        asm_ops.append((op, len(outs) - 1, reg(a), reg(b)))
    nout = len(outs)
    txt = []
    for op, d, a, b in asm_ops:
        fa = a if a == "0" else "%" + str(nout + int(a[2:]))
        fb = b if b == "0" else "%" + str(nout + int(b[2:]))
        txt.append(f"{op} %{d},{fa},{fb};")
    out.append('    asm("' + " ".join(txt) + '"\n        : ' + ", ".join(f'"=r"({o_})' for o_ in outs) +
               "\n        : " + ", ".join(f'"r"({i_})' for i_ in inputs) + ");")
    return "\n".join(out)


def merge_check(p, trials=500):
    """The merge is an add chain over live limbs; check its structure numerically."""
    rnd = random.Random(2)
    for _ in range(trials):
        a = [rnd.choice([0, M32, rnd.getrandbits(32)]) for _ in range(8)]
        regs = p.run(a)
        carry, c = 0, []
        started = False
        for k in range(16):
            ev = regs.get(("e", k), 0) if k in p.init["e"] else 0
            ov = regs.get(("o", k - 1), 0) if (k - 1) in p.init["o"] else 0
            both = (k in p.init["e"]) and ((k - 1) in p.init["o"])
            if not started and not both:
                v = ev + ov
                assert v <= M32
            else:
                started = True
                v = ev + ov + carry
            c.append(v & M32)
            carry = v >> 32
        assert carry == 0
        cross = sum(v << (32 * k) for k, v in enumerate(c))
        want = sum(a[i] * a[j] << (32 * (i + j)) for i in range(8) for j in range(i + 1, 8))
        assert cross == want
        assert ((2 * cross + sum(a[i] * a[i] << (64 * i) for i in range(8))) == sum(x << (32 * i) for i, x in enumerate(a)) ** 2)


def main():
    p = build()
    n = reference_check(p)
    merge_check(p)
    e, o = sorted(p.init["e"]), sorted(p.init["o"])
    decl = ", ".join([f"e{k}" for k in e] + [f"o{k}" for k in o])
    src = HEADER.format(n=n, decl=decl, chains=p.emit(), merge=merge_code(p), elive=e, olive=o)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "secp256k1-voi_b200", "csrc", "fe_sqr_gen.cuh")
    open(path, "w").write(src)
    nmad = sum(1 for ops in p.stmts for op in ops if op[0].endswith("hi.cc"))
    print(f"wrote {path}: {nmad} cross IMAD.WIDE + 8 squares; emulated on {n} inputs; live e={e} o={o}")


if __name__ == "__main__":
    main()
