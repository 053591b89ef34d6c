// fe_vt.cuh -- variable-time flavour of the F_p operations, for PUBLIC data only.
//
// The constant-time code (fe.cuh) finishes every addition, subtraction and reduction with a full
// 8-limb ripple of the 2^256 = delta fold, because a shortcut would branch on the data.  The
// verification ladder (DoubleScalarMultBasepointVartime, point_mul_glv.go:307) and the vartime MSM
// (point_mul_multi.go:73) handle public values and are variable time in the reference too, so here
// the ripple stops after limb 2 and continues behind a branch only if a carry really leaves that limb
// (probability 2^-32 per operation for random data; any input stays correct).  That removes 5 of the
// ~21 instructions of an addition and 5 of a reduction; the ladder kernel is sensitive to its
// instruction count, not only to its multiplier work.
//
// fe_ops<false> is the constant-time set, fe_ops<true> this one; the group law (point.cuh) is a
// template over it.  Off the device both are the same portable code.
#pragma once
#include "fe.cuh"

namespace s256 {

#if S256_PTX

S256_D void fe_fold_carry_vt(fe &r, uint32_t c) {
    uint32_t c3;
    uint32_t t = c ? S256_DELTA_LO : 0u;  // c in {0, 1}; a product or a negation here would issue on the multiplier's pipe
    asm("add.cc.u32 %0,%0,%4; addc.cc.u32 %1,%1,%5; addc.cc.u32 %2,%2,0; addc.u32 %3,0,0;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "=r"(c3)
        : "r"(t), "r"(c));
    if (c3) {
        uint32_t c2;
        asm("add.cc.u32 %0,%0,1; addc.cc.u32 %1,%1,0; addc.cc.u32 %2,%2,0; addc.cc.u32 %3,%3,0; addc.cc.u32 %4,%4,0;"
            "addc.u32 %5,0,0;"
            : "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7]), "=r"(c2));
        if (c2) {  // wrapped: r < delta now, add delta once more, no further carry possible
            asm("add.cc.u32 %0,%0,%3; addc.cc.u32 %1,%1,1; addc.u32 %2,%2,0;"
                : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2])
                : "r"(S256_DELTA_LO));
        }
    }
}
S256_D void fe_add_vt(fe &r, const fe &a, const fe &b) {
    uint32_t c = fe_add_raw(r, a, b);
    fe_fold_carry_vt(r, c);
}
S256_D void fe_sub_vt(fe &r, const fe &a, const fe &b) {
    uint32_t bw = fe_sub_raw(r, a, b);
    uint32_t one = bw & 1u, t = bw & S256_DELTA_LO, b3;
    asm("sub.cc.u32 %0,%0,%4; subc.cc.u32 %1,%1,%5; subc.cc.u32 %2,%2,0; subc.u32 %3,0,0;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "=r"(b3)
        : "r"(t), "r"(one));
    if (b3) {
        uint32_t bw2;
        asm("sub.cc.u32 %0,%0,1; subc.cc.u32 %1,%1,0; subc.cc.u32 %2,%2,0; subc.cc.u32 %3,%3,0; subc.cc.u32 %4,%4,0;"
            "subc.u32 %5,0,0;"
            : "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7]), "=r"(bw2));
        if (bw2) {  // wrapped again: r >= 2^256 - delta now, subtracting delta cannot borrow
            asm("sub.cc.u32 %0,%0,%8; subc.cc.u32 %1,%1,1; subc.cc.u32 %2,%2,0; subc.cc.u32 %3,%3,0;"
                "subc.cc.u32 %4,%4,0; subc.cc.u32 %5,%5,0; subc.cc.u32 %6,%6,0; subc.u32 %7,%7,0;"
                : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]),
                  "+r"(r.v[7])
                : "r"(S256_DELTA_LO));
        }
    }
}
// out = t[0..7] + (t8 + 2^32 * t9) * delta.  The common case is three instructions on limbs 0..2 after the one wide
// multiplication: (out0, out1) = t8 * 977 + (t1 : t0) with the carry into limb 2, then "+ t8 << 32".  Anything
// else -- a carry leaving limb 2 (probability ~2^-30) or t9 = 1 (the high half within 2^-22 of 2^256) -- takes the
// slow path behind one branch.  No multiplication by t9, no zero register for a 64-bit addend: nothing here issues
// on the multiplier's pipe except the one product.
S256_D void fe_fold_top_vt(fe &out, uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3, uint32_t t4, uint32_t t5,
                           uint32_t t6, uint32_t t7, uint32_t t8, uint32_t t9) {
    uint32_t c3, c3b;  // both carries are captured (never "c3 += carry": a final add issues as IMAD.X)
    asm("mad.lo.cc.u32 %0,%4,%5,%6; madc.hi.cc.u32 %1,%4,%5,%7; addc.cc.u32 %2,%8,0; addc.u32 %3,0,0;"
        : "=&r"(out.v[0]), "=&r"(out.v[1]), "=&r"(out.v[2]), "=&r"(c3)
        : "r"(t8), "r"(S256_DELTA_LO), "r"(t0), "r"(t1), "r"(t2));
    asm("add.cc.u32 %0,%0,%3; addc.cc.u32 %1,%1,0; addc.u32 %2,0,0;" : "+r"(out.v[1]), "+r"(out.v[2]), "=r"(c3b) : "r"(t8));
    out.v[3] = t3; out.v[4] = t4; out.v[5] = t5; out.v[6] = t6; out.v[7] = t7;
    if (c3 | c3b | t9) {  // rare: + t9 * (977 << 32) + (t9 << 64) + ((c3 + c3b) << 96), then the wrap of a carry out of limb 7
        uint32_t c;
        c3 += c3b;
        asm("add.cc.u32 %0,%0,%8; addc.cc.u32 %1,%1,%9; addc.cc.u32 %2,%2,%10; addc.cc.u32 %3,%3,0; addc.cc.u32 %4,%4,0;"
            "addc.cc.u32 %5,%5,0; addc.cc.u32 %6,%6,0; addc.u32 %7,0,0;"
            : "+r"(out.v[1]), "+r"(out.v[2]), "+r"(out.v[3]), "+r"(out.v[4]), "+r"(out.v[5]), "+r"(out.v[6]), "+r"(out.v[7]),
              "=r"(c)
            : "r"((0u - t9) & S256_DELTA_LO), "r"(t9), "r"(c3));
        if (c) {  // out < 2^67 now; one more delta, no carry possible
            asm("add.cc.u32 %0,%0,%3; addc.cc.u32 %1,%1,1; addc.u32 %2,%2,0;"
                : "+r"(out.v[0]), "+r"(out.v[1]), "+r"(out.v[2])
                : "r"(S256_DELTA_LO));
        }
    }
}
// The same fold for a top word pair with a small t9 (<= 7): what the sum of two products leaves (fe_mul2add_core_w).  The
// fast path adds t8 * delta + t9 * (977 << 32) + (t9 << 64) on limbs 0..2; a carry leaving limb 2 takes the slow path.
S256_D void fe_fold_top2_vt(fe &out, uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3, uint32_t t4, uint32_t t5,
                            uint32_t t6, uint32_t t7, uint32_t t8, uint32_t t9) {
    uint32_t c3, c3b, c3c;
    uint32_t t9m = t9 * S256_DELTA_LO;
    asm("mad.lo.cc.u32 %0,%4,%5,%6; madc.hi.cc.u32 %1,%4,%5,%7; addc.cc.u32 %2,%8,0; addc.u32 %3,0,0;"
        : "=&r"(out.v[0]), "=&r"(out.v[1]), "=&r"(out.v[2]), "=&r"(c3)
        : "r"(t8), "r"(S256_DELTA_LO), "r"(t0), "r"(t1), "r"(t2));
    asm("add.cc.u32 %0,%0,%3; addc.cc.u32 %1,%1,%4; addc.u32 %2,0,0;" : "+r"(out.v[1]), "+r"(out.v[2]), "=r"(c3b) : "r"(t8), "r"(t9));
    asm("add.cc.u32 %0,%0,%3; addc.cc.u32 %1,%1,0; addc.u32 %2,0,0;" : "+r"(out.v[1]), "+r"(out.v[2]), "=r"(c3c) : "r"(t9m));
    out.v[3] = t3; out.v[4] = t4; out.v[5] = t5; out.v[6] = t6; out.v[7] = t7;
    if (c3 | c3b | c3c) {  // rare: ripple the carries out of limb 2, then the wrap of a carry out of limb 7
        uint32_t c;
        c3 += c3b + c3c;
        asm("add.cc.u32 %0,%0,%6; addc.cc.u32 %1,%1,0; addc.cc.u32 %2,%2,0; addc.cc.u32 %3,%3,0; addc.cc.u32 %4,%4,0;"
            "addc.u32 %5,0,0;"
            : "+r"(out.v[3]), "+r"(out.v[4]), "+r"(out.v[5]), "+r"(out.v[6]), "+r"(out.v[7]), "=r"(c)
            : "r"(c3));
        if (c) {  // out < 2^70 now; one more delta, no carry possible
            asm("add.cc.u32 %0,%0,%3; addc.cc.u32 %1,%1,1; addc.u32 %2,%2,0;"
                : "+r"(out.v[0]), "+r"(out.v[1]), "+r"(out.v[2])
                : "r"(S256_DELTA_LO));
        }
    }
}
// r = a * b + c * d with one reduction
S256_D void fe_mul2add_inline_vt(fe &r, const fe &a, const fe &b, const fe &c, const fe &d) {
    uint32_t w[9], t9;
    fe_mul2add_core_w(w, t9, a.v, b.v, c.v, d.v);
    fe_fold_top2_vt(r, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], w[8], t9);
}
S256_D void fe_mul_inline_vt(fe &r, const fe &a, const fe &b) {
#if !defined(S256_MUL_MERGED)
    uint32_t w[9], t9;
    fe_mul_core_w(w, t9, a.v, b.v);
    fe_fold_top_vt(r, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], w[8], t9);
#else
    uint32_t w[16], t8, t9;
    fe_mul_wide(w, a.v, b.v);
    fe_reduce_wide_pre(w, t8, t9);
    fe_fold_top_vt(r, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], t8, t9);
#endif
}
S256_D void fe_sqr_inline_vt(fe &r, const fe &a) {
#if !defined(S256_NO_SQR) && !defined(S256_MUL_MERGED)
    uint32_t w[9], t9;
    fe_sqr_core_w(w, t9, a.v);
    fe_fold_top_vt(r, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], w[8], t9);
#elif !defined(S256_NO_SQR)
    uint32_t w[16], t8, t9;
    fe_sqr_wide(w, a.v);
    fe_reduce_wide_pre(w, t8, t9);
    fe_fold_top_vt(r, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], t8, t9);
#else
    fe_mul_inline_vt(r, a, a);
#endif
}
#ifndef S256_MUL_INLINE
static __device__ __noinline__ fe fe_mul_call_vt(fe a, fe b) {
    fe r;
    fe_mul_inline_vt(r, a, b);
    return r;
}
static __device__ __noinline__ fe fe_sqr_call_vt(fe a) {
    fe r;
    fe_sqr_inline_vt(r, a);
    return r;
}
S256_D void fe_mul_vt(fe &r, const fe &a, const fe &b) { r = fe_mul_call_vt(a, b); }
S256_D void fe_sqr_vt(fe &r, const fe &a) { r = fe_sqr_call_vt(a); }
#ifndef S256_NO_FUSED
static __device__ __noinline__ fe fe_mul2add_call_vt(fe a, fe b, fe c, fe d) {
    fe r;
    fe_mul2add_inline_vt(r, a, b, c, d);
    return r;
}
S256_D void fe_mul2add_vt(fe &r, const fe &a, const fe &b, const fe &c, const fe &d) { r = fe_mul2add_call_vt(a, b, c, d); }
#else
S256_D void fe_mul2add_vt(fe &r, const fe &a, const fe &b, const fe &c, const fe &d) {
    fe t, u;
    fe_mul_vt(t, a, b);
    fe_mul_vt(u, c, d);
    fe_add_vt(r, t, u);
}
#endif
#else
S256_D void fe_mul2add_vt(fe &r, const fe &a, const fe &b, const fe &c, const fe &d) { fe_mul2add_inline_vt(r, a, b, c, d); }
S256_D void fe_mul_vt(fe &r, const fe &a, const fe &b) { fe_mul_inline_vt(r, a, b); }
S256_D void fe_sqr_vt(fe &r, const fe &a) { fe_sqr_inline_vt(r, a); }
#endif
// The fused a b + c d for the CONSTANT-TIME flavour: the same core (its carry-outs are captured with selects, no
// branch), a full-ripple fold of the top word pair (t9 <= 3) and a masked second fold -- fe_fold_top of fe.cuh with a
// small multiple in place of its 0 / 1 mask.
S256_D void fe_fold_top2_ct(fe &out, uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3, uint32_t t4, uint32_t t5,
                            uint32_t t6, uint32_t t7, uint32_t t8, uint32_t t9) {
    uint32_t u0, u1, u2, c;
    uint32_t t9d = t9 * S256_DELTA_LO;
    asm("mul.lo.u32 %0,%2,%3; mul.hi.u32 %1,%2,%3;" : "=r"(u0), "=r"(u1) : "r"(t8), "r"(S256_DELTA_LO));
    asm("add.cc.u32 %0,%0,%2; addc.u32 %1,%3,0;" : "+r"(u1), "=r"(u2) : "r"(t8), "r"(t9));
    asm("add.cc.u32 %0,%0,%2; addc.u32 %1,%1,0;" : "+r"(u1), "+r"(u2) : "r"(t9d));
    asm("add.cc.u32 %0,%9,%17; addc.cc.u32 %1,%10,%18; addc.cc.u32 %2,%11,%19; addc.cc.u32 %3,%12,0;"
        "addc.cc.u32 %4,%13,0; addc.cc.u32 %5,%14,0; addc.cc.u32 %6,%15,0; addc.cc.u32 %7,%16,0; addc.u32 %8,0,0;"
        : "=r"(out.v[0]), "=r"(out.v[1]), "=r"(out.v[2]), "=r"(out.v[3]), "=r"(out.v[4]), "=r"(out.v[5]),
          "=r"(out.v[6]), "=r"(out.v[7]), "=r"(c)
        : "r"(t0), "r"(t1), "r"(t2), "r"(t3), "r"(t4), "r"(t5), "r"(t6), "r"(t7), "r"(u0), "r"(u1), "r"(u2));
    // on carry the value is < 2^68 now: one more delta, no carry possible; masked, not branched
    asm("add.cc.u32 %0,%0,%3; addc.cc.u32 %1,%1,%4; addc.u32 %2,%2,0;"
        : "+r"(out.v[0]), "+r"(out.v[1]), "+r"(out.v[2])
        : "r"((0u - c) & S256_DELTA_LO), "r"(c));
}
S256_D void fe_mul2add_inline_ct(fe &r, const fe &a, const fe &b, const fe &c, const fe &d) {
    uint32_t w[9], t9;
    fe_mul2add_core_w(w, t9, a.v, b.v, c.v, d.v);
    fe_fold_top2_ct(r, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], w[8], t9);
}
#ifndef S256_MUL_INLINE
static __device__ __noinline__ fe fe_mul2add_call_ct(fe a, fe b, fe c, fe d) {
    fe r;
    fe_mul2add_inline_ct(r, a, b, c, d);
    return r;
}
S256_D void fe_mul2add_ct(fe &r, const fe &a, const fe &b, const fe &c, const fe &d) { r = fe_mul2add_call_ct(a, b, c, d); }
#else
S256_D void fe_mul2add_ct(fe &r, const fe &a, const fe &b, const fe &c, const fe &d) { fe_mul2add_inline_ct(r, a, b, c, d); }
#endif
S256_D void fe_mul_small_vt(fe &r, const fe &a, uint32_t k) {
    uint32_t e[8], t8;
    fe_mul_small_pre(e, t8, a, k);
    fe_fold_top_vt(r, e[0], e[1], e[2], e[3], e[4], e[5], e[6], e[7], t8, 0u);
}
// r = 21a = a + 4a + 16a (b3 of the complete formulas) WITHOUT the multiplier: two funnel-shift passes and two add
// chains on the ALU pipe, which has headroom in the ladders, instead of 8 wide multiplications plus the register
// pairing ptxas wraps around them on the pipe that bounds the kernel.  Returns limbs + top (< 32) for the fold.
S256_D void fe_mul21_pre(uint32_t e[8], uint32_t &t8, const fe &a) {
    uint32_t x[8], c1, c2;
#pragma unroll
    for (int i = 7; i >= 1; i--) x[i] = __funnelshift_l(a.v[i - 1], a.v[i], 2);
    x[0] = __funnelshift_l(0u, a.v[0], 2);
    asm("add.cc.u32 %0,%9,%17; addc.cc.u32 %1,%10,%18; addc.cc.u32 %2,%11,%19; addc.cc.u32 %3,%12,%20;"
        "addc.cc.u32 %4,%13,%21; addc.cc.u32 %5,%14,%22; addc.cc.u32 %6,%15,%23; addc.cc.u32 %7,%16,%24;"
        "addc.u32 %8,0,0;"
        : "=r"(e[0]), "=r"(e[1]), "=r"(e[2]), "=r"(e[3]), "=r"(e[4]), "=r"(e[5]), "=r"(e[6]), "=r"(e[7]), "=r"(c1)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]));
#pragma unroll
    for (int i = 7; i >= 1; i--) x[i] = __funnelshift_l(a.v[i - 1], a.v[i], 4);
    x[0] = __funnelshift_l(0u, a.v[0], 4);
    asm("add.cc.u32 %0,%0,%9; addc.cc.u32 %1,%1,%10; addc.cc.u32 %2,%2,%11; addc.cc.u32 %3,%3,%12;"
        "addc.cc.u32 %4,%4,%13; addc.cc.u32 %5,%5,%14; addc.cc.u32 %6,%6,%15; addc.cc.u32 %7,%7,%16;"
        "addc.u32 %8,0,0;"
        : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]), "=r"(c2)
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]));
    t8 = (a.v[7] >> 30) + (a.v[7] >> 28) + c1 + c2;
}
// out = e[0..7] + t * delta for a small t (< 2^16): t * 977 fits one word, so the fold is three adds on limbs 0..2
S256_D void fe_fold_small_vt(fe &out, const uint32_t e[8], uint32_t t) {
    uint32_t c3;
    uint32_t u = t * S256_DELTA_LO;
    asm("add.cc.u32 %0,%4,%7; addc.cc.u32 %1,%5,%8; addc.cc.u32 %2,%6,0; addc.u32 %3,0,0;"
        : "=r"(out.v[0]), "=r"(out.v[1]), "=r"(out.v[2]), "=r"(c3)
        : "r"(e[0]), "r"(e[1]), "r"(e[2]), "r"(u), "r"(t));
    out.v[3] = e[3]; out.v[4] = e[4]; out.v[5] = e[5]; out.v[6] = e[6]; out.v[7] = e[7];
    if (c3) {
        uint32_t c2;
        asm("add.cc.u32 %0,%0,1; addc.cc.u32 %1,%1,0; addc.cc.u32 %2,%2,0; addc.cc.u32 %3,%3,0; addc.cc.u32 %4,%4,0;"
            "addc.u32 %5,0,0;"
            : "+r"(out.v[3]), "+r"(out.v[4]), "+r"(out.v[5]), "+r"(out.v[6]), "+r"(out.v[7]), "=r"(c2));
        if (c2) {  // wrapped: out is tiny now, one more delta cannot carry
            asm("add.cc.u32 %0,%0,%3; addc.cc.u32 %1,%1,1; addc.u32 %2,%2,0;"
                : "+r"(out.v[0]), "+r"(out.v[1]), "+r"(out.v[2])
                : "r"(S256_DELTA_LO));
        }
    }
}
S256_D void fe_mul21_vt(fe &r, const fe &a) {
    uint32_t e[8], t8;
    fe_mul21_pre(e, t8, a);
    fe_fold_small_vt(r, e, t8);
}
// r = 8a: one funnel-shift pass and a fold of the three bits shifted out, instead of three additions
S256_D void fe_mul8_vt(fe &r, const fe &a) {
    uint32_t top = a.v[7] >> 29;
#pragma unroll
    for (int i = 7; i >= 1; i--) r.v[i] = __funnelshift_l(a.v[i - 1], a.v[i], 3);
    r.v[0] = a.v[0] << 3;
    uint32_t t = top * S256_DELTA_LO, c3;
    asm("add.cc.u32 %0,%0,%4; addc.cc.u32 %1,%1,%5; addc.cc.u32 %2,%2,0; addc.u32 %3,0,0;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "=r"(c3)
        : "r"(t), "r"(top));
    if (c3) {
        uint32_t c2;
        asm("add.cc.u32 %0,%0,1; addc.cc.u32 %1,%1,0; addc.cc.u32 %2,%2,0; addc.cc.u32 %3,%3,0; addc.cc.u32 %4,%4,0;"
            "addc.u32 %5,0,0;"
            : "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7]), "=r"(c2));
        if (c2) {  // wrapped: r < 8 * delta now, one more delta cannot carry
            asm("add.cc.u32 %0,%0,%3; addc.cc.u32 %1,%1,1; addc.u32 %2,%2,0;"
                : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2])
                : "r"(S256_DELTA_LO));
        }
    }
}

// ---- small multiples fused with the neighbouring addition / subtraction (Jacobian formulas, jac.cuh) ----
// Each replaces two or three carry chains of dependent adds by one: the multiple is formed with funnel shifts
// (independent of each other), the bits shifted out of limb 7 and the chain's carry / borrow are folded together.
// r -= k * delta for a small k (k * 977 fits a word): three subtractions on limbs 0..2, the rest behind a branch
S256_D void fe_fold_borrow_k_vt(fe &r, uint32_t k) {
    uint32_t t = k * S256_DELTA_LO, b3;
    asm("sub.cc.u32 %0,%0,%4; subc.cc.u32 %1,%1,%5; subc.cc.u32 %2,%2,0; subc.u32 %3,0,0;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "=r"(b3)
        : "r"(t), "r"(k));
    if (b3) {
        uint32_t bw2;
        asm("sub.cc.u32 %0,%0,1; subc.cc.u32 %1,%1,0; subc.cc.u32 %2,%2,0; subc.cc.u32 %3,%3,0; subc.cc.u32 %4,%4,0;"
            "subc.u32 %5,0,0;"
            : "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7]), "=r"(bw2));
        if (bw2) {  // wrapped again: r >= 2^256 - k * delta now, subtracting delta cannot borrow
            asm("sub.cc.u32 %0,%0,%8; subc.cc.u32 %1,%1,1; subc.cc.u32 %2,%2,0; subc.cc.u32 %3,%3,0;"
                "subc.cc.u32 %4,%4,0; subc.cc.u32 %5,%5,0; subc.cc.u32 %6,%6,0; subc.u32 %7,%7,0;"
                : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]),
                  "+r"(r.v[7])
                : "r"(S256_DELTA_LO));
        }
    }
}
// x = a << s (limbs), returns the s bits shifted out
template <int SH>
S256_D uint32_t fe_shl_limbs(uint32_t x[8], const fe &a) {
#pragma unroll
    for (int i = 7; i >= 1; i--) x[i] = __funnelshift_l(a.v[i - 1], a.v[i], SH);
    x[0] = a.v[0] << SH;
    return a.v[7] >> (32 - SH);
}
// r = 2a
S256_D void fe_mul2_vt(fe &r, const fe &a) {
    uint32_t x[8];
    uint32_t top = fe_shl_limbs<1>(x, a);
    fe_fold_small_vt(r, x, top);
}
// r = 3a = a + 2a
S256_D void fe_mul3_vt(fe &r, const fe &a) {
    fe x, e;
    uint32_t top = fe_shl_limbs<1>(x.v, a);
    uint32_t c = fe_add_raw(e, a, x);
    fe_fold_small_vt(r, e.v, top + c);
}
// r = a - 2b
S256_D void fe_sub2_vt(fe &r, const fe &a, const fe &b) {
    fe x;
    uint32_t top = fe_shl_limbs<1>(x.v, b);
    uint32_t bw = fe_sub_raw(r, a, x);
    fe_fold_borrow_k_vt(r, top + (bw & 1u));
}
// r = a - 8b
S256_D void fe_submul8_vt(fe &r, const fe &a, const fe &b) {
    fe x;
    uint32_t top = fe_shl_limbs<3>(x.v, b);
    uint32_t bw = fe_sub_raw(r, a, x);
    fe_fold_borrow_k_vt(r, top + (bw & 1u));
}

#else  // portable: one implementation serves both flavours

S256_HD void fe_add_vt(fe &r, const fe &a, const fe &b) { fe_add(r, a, b); }
S256_HD void fe_sub_vt(fe &r, const fe &a, const fe &b) { fe_sub(r, a, b); }
S256_HD void fe_mul_vt(fe &r, const fe &a, const fe &b) { fe_mul(r, a, b); }
S256_HD void fe_sqr_vt(fe &r, const fe &a) { fe_sqr(r, a); }
S256_HD void fe_mul2add_vt(fe &r, const fe &a, const fe &b, const fe &c, const fe &d) {
    fe t, u;
    fe_mul(t, a, b);
    fe_mul(u, c, d);
    fe_add(r, t, u);
}
S256_HD void fe_mul_small_vt(fe &r, const fe &a, uint32_t k) { fe_mul_small(r, a, k); }
S256_HD void fe_mul8_vt(fe &r, const fe &a) {
    fe t;
    fe_add(t, a, a);
    fe_add(t, t, t);
    fe_add(r, t, t);
}

S256_HD void fe_mul2_vt(fe &r, const fe &a) { fe_add(r, a, a); }
S256_HD void fe_mul3_vt(fe &r, const fe &a) {
    fe t;
    fe_add(t, a, a);
    fe_add(r, t, a);
}
S256_HD void fe_sub2_vt(fe &r, const fe &a, const fe &b) {
    fe t;
    fe_sub(t, a, b);
    fe_sub(r, t, b);
}
S256_HD void fe_submul8_vt(fe &r, const fe &a, const fe &b) {
    fe t;
    fe_mul8_vt(t, b);
    fe_sub(r, a, t);
}

#endif

template <bool VT>
struct fe_ops;
template <>
struct fe_ops<false> {
    S256_HD static void add(fe &r, const fe &a, const fe &b) { fe_add(r, a, b); }
    S256_HD static void sub(fe &r, const fe &a, const fe &b) { fe_sub(r, a, b); }
    S256_HD static void mul(fe &r, const fe &a, const fe &b) { fe_mul(r, a, b); }
    S256_HD static void sqr(fe &r, const fe &a) { fe_sqr(r, a); }
    // a b + c d and a b - c d: one fused, branch-free call on the device (fe_mul2add_ct), two products and an addition elsewhere
#if S256_PTX && !defined(S256_NO_FUSED) && !defined(S256_NO_FUSED_CT)
    S256_HD static void mul2add(fe &r, const fe &a, const fe &b, const fe &c, const fe &d) { fe_mul2add_ct(r, a, b, c, d); }
    S256_HD static void mul2sub(fe &r, const fe &a, const fe &b, const fe &c, const fe &d) {
        fe nd;
        fe_neg(nd, d);
        fe_mul2add_ct(r, a, b, c, nd);
    }
#else
    S256_HD static void mul2add(fe &r, const fe &a, const fe &b, const fe &c, const fe &d) {
        fe t, u;
        fe_mul(t, a, b);
        fe_mul(u, c, d);
        fe_add(r, t, u);
    }
    S256_HD static void mul2sub(fe &r, const fe &a, const fe &b, const fe &c, const fe &d) {
        fe t, u;
        fe_mul(t, a, b);
        fe_mul(u, c, d);
        fe_sub(r, t, u);
    }
#endif
#if S256_PTX && !defined(S256_B3_MULT)
    // 21a and 8a by shifts and adds with the branch-free fold (the constant-time flavour of fe_mul21_vt / fe_mul8_vt)
    S256_HD static void mul_small(fe &r, const fe &a, uint32_t k) {
        if (k == 21u) {
            uint32_t e[8], t8;
            fe_mul21_pre(e, t8, a);
            fe_fold_top(r, e[0], e[1], e[2], e[3], e[4], e[5], e[6], e[7], t8, 0u);
        } else {
            fe_mul_small(r, a, k);
        }
    }
    S256_HD static void mul8(fe &r, const fe &a) {
        uint32_t e[8];
#pragma unroll
        for (int i = 7; i >= 1; i--) e[i] = __funnelshift_l(a.v[i - 1], a.v[i], 3);
        e[0] = a.v[0] << 3;
        fe_fold_top(r, e[0], e[1], e[2], e[3], e[4], e[5], e[6], e[7], a.v[7] >> 29, 0u);
    }
#else
    S256_HD static void mul_small(fe &r, const fe &a, uint32_t k) { fe_mul_small(r, a, k); }
    S256_HD static void mul8(fe &r, const fe &a) {
        fe t;
        fe_add(t, a, a);
        fe_add(t, t, t);
        fe_add(r, t, t);
    }
#endif
    S256_HD static void mul2(fe &r, const fe &a) { fe_add(r, a, a); }
    S256_HD static void mul3(fe &r, const fe &a) {
        fe t;
        fe_add(t, a, a);
        fe_add(r, t, a);
    }
    S256_HD static void sub2(fe &r, const fe &a, const fe &b) {
        fe t;
        fe_sub(t, a, b);
        fe_sub(r, t, b);
    }
    S256_HD static void submul8(fe &r, const fe &a, const fe &b) {
        fe t;
        mul8(t, b);
        fe_sub(r, a, t);
    }
};
template <>
struct fe_ops<true> {
    S256_HD static void add(fe &r, const fe &a, const fe &b) { fe_add_vt(r, a, b); }
    S256_HD static void sub(fe &r, const fe &a, const fe &b) { fe_sub_vt(r, a, b); }
    S256_HD static void mul(fe &r, const fe &a, const fe &b) { fe_mul_vt(r, a, b); }
    S256_HD static void sqr(fe &r, const fe &a) { fe_sqr_vt(r, a); }
    S256_HD static void mul2add(fe &r, const fe &a, const fe &b, const fe &c, const fe &d) { fe_mul2add_vt(r, a, b, c, d); }
    S256_HD static void mul2sub(fe &r, const fe &a, const fe &b, const fe &c, const fe &d) {
        fe z = fe_zero(), nd;
        fe_sub_vt(nd, z, d);
        fe_mul2add_vt(r, a, b, c, nd);
    }
#if S256_PTX && !defined(S256_B3_MULT)
    S256_HD static void mul_small(fe &r, const fe &a, uint32_t k) {
        if (k == 21u) fe_mul21_vt(r, a); else fe_mul_small_vt(r, a, k);
    }
#else
    S256_HD static void mul_small(fe &r, const fe &a, uint32_t k) { fe_mul_small_vt(r, a, k); }
#endif
    S256_HD static void mul8(fe &r, const fe &a) { fe_mul8_vt(r, a); }
#ifndef S256_NO_FUSED_SMALL
    S256_HD static void mul2(fe &r, const fe &a) { fe_mul2_vt(r, a); }
    S256_HD static void mul3(fe &r, const fe &a) { fe_mul3_vt(r, a); }
    S256_HD static void sub2(fe &r, const fe &a, const fe &b) { fe_sub2_vt(r, a, b); }
    S256_HD static void submul8(fe &r, const fe &a, const fe &b) { fe_submul8_vt(r, a, b); }
#else
    S256_HD static void mul2(fe &r, const fe &a) { fe_add_vt(r, a, a); }
    S256_HD static void mul3(fe &r, const fe &a) {
        fe t;
        fe_add_vt(t, a, a);
        fe_add_vt(r, t, a);
    }
    S256_HD static void sub2(fe &r, const fe &a, const fe &b) {
        fe t;
        fe_sub_vt(t, a, b);
        fe_sub_vt(r, t, b);
    }
    S256_HD static void submul8(fe &r, const fe &a, const fe &b) {
        fe t;
        fe_mul8_vt(t, b);
        fe_sub_vt(r, a, t);
    }
#endif
};

}  // namespace s256
