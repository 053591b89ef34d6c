// fe.cuh -- arithmetic in F_p, p = 2^256 - 2^32 - 977, for sm_100a.
//
// Replaces the reference's fiat-crypto 4x64 Montgomery code
// (internal/fiat/secp256k1montgomery/secp256k1montgomery.go:87,418,750,802,844
// behind internal/field/field.go:61-104).  Montgomery form is unobservable at
// the reference's byte boundary, so elements are kept in PLAIN form as
// 8 x 32-bit little-endian limbs, "weakly reduced": any value in [0, 2^256)
// congruent to the element.  Canonical (< p) form is produced only by
// fe_normalize(), which every compare / parity / encode goes through.
//
// Device path: 64 IMAD.WIDE.U32 for the 8x8 limb products, written as
// mad.lo.cc / madc.hi.cc carry chains split into even- and odd-aligned
// accumulators so that no chain ever waits on the other; then the special
// reduction 2^256 = 2^32 + 977 (mod p): 8 more IMAD.WIDE for hi*977, the
// "<< 32" part is a limb shift.  73 MAC32 per modmul.
//
// Host path (S256_HOSTSIM or !__CUDA_ARCH__): the same functions in portable
// C so that kernel logic can be exercised by tests/hostsim on a CPU-only box.
// It is never linked into the product library's compute path.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define S256_HD __host__ __device__ __forceinline__
#define S256_D __device__ __forceinline__
#else
#define S256_HD inline
#define S256_D inline
#endif

#if defined(__CUDA_ARCH__) && !defined(S256_FE_PORTABLE)
#define S256_PTX 1
#else
#define S256_PTX 0
#endif

namespace s256 {

struct fe {
    uint32_t v[8];
};

// p = 2^256 - 2^32 - 977; delta = 2^256 - p = 2^32 + 977
#define S256_P0 0xFFFFFC2Fu
#define S256_P1 0xFFFFFFFEu
#define S256_DELTA_LO 977u

S256_HD fe fe_zero() {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}
S256_HD fe fe_from_u32(uint32_t x) {
    fe r = fe_zero();
    r.v[0] = x;
    return r;
}
S256_HD fe fe_one() { return fe_from_u32(1); }

// ---------------------------------------------------------------------------
// add / sub with the 2^256 = delta fold.  Results stay in [0, 2^256).
// ---------------------------------------------------------------------------
#if S256_PTX

// r += c * delta where c in {0,1}.  A carry out is only possible when r was within delta of 2^256, after which r is
// < delta and delta is added once more (no further carry possible).  Branch-free: the second addition is masked,
// because the constant-time kernels run on these (point_mul_table.go:168, point_mul_glv.go:257 are branch-free too).
S256_D void fe_fold_carry(fe &r, uint32_t c) {
    uint32_t c2;
    uint32_t t = (0u - c) & S256_DELTA_LO;
    asm("add.cc.u32 %0,%0,%9; addc.cc.u32 %1,%1,%10; addc.cc.u32 %2,%2,0; addc.cc.u32 %3,%3,0;"
        "addc.cc.u32 %4,%4,0; addc.cc.u32 %5,%5,0; addc.cc.u32 %6,%6,0; addc.cc.u32 %7,%7,0; addc.u32 %8,0,0;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]),
          "+r"(r.v[7]), "=r"(c2)
        : "r"(t), "r"(c));
    asm("add.cc.u32 %0,%0,%3; addc.cc.u32 %1,%1,%4; addc.u32 %2,%2,0;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2])
        : "r"((0u - c2) & S256_DELTA_LO), "r"(c2));
}

// r = a + b mod 2^256, returns the carry
S256_D uint32_t fe_add_raw(fe &r, const fe &a, const fe &b) {
    uint32_t c;
    asm("add.cc.u32 %0,%9,%17; addc.cc.u32 %1,%10,%18; addc.cc.u32 %2,%11,%19; addc.cc.u32 %3,%12,%20;"
        "addc.cc.u32 %4,%13,%21; addc.cc.u32 %5,%14,%22; addc.cc.u32 %6,%15,%23; addc.cc.u32 %7,%16,%24;"
        "addc.u32 %8,0,0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
          "=r"(r.v[7]), "=r"(c)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    return c;
}
S256_D void fe_add(fe &r, const fe &a, const fe &b) {
    uint32_t c = fe_add_raw(r, a, b);
    fe_fold_carry(r, c);
}

// r = a - b mod 2^256, returns the borrow as 0 / 0xFFFFFFFF
S256_D uint32_t fe_sub_raw(fe &r, const fe &a, const fe &b) {
    uint32_t bw;
    asm("sub.cc.u32 %0,%9,%17; subc.cc.u32 %1,%10,%18; subc.cc.u32 %2,%11,%19; subc.cc.u32 %3,%12,%20;"
        "subc.cc.u32 %4,%13,%21; subc.cc.u32 %5,%14,%22; subc.cc.u32 %6,%15,%23; subc.cc.u32 %7,%16,%24;"
        "subc.u32 %8,0,0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
          "=r"(r.v[7]), "=r"(bw)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    return bw;
}
// on borrow the true value is r - 2^256 = r - delta (mod p)
S256_D void fe_fold_borrow(fe &r, uint32_t bw) {
    uint32_t bw2;
    uint32_t one = bw & 1u;
    uint32_t t = bw & S256_DELTA_LO;
    asm("sub.cc.u32 %0,%0,%9; subc.cc.u32 %1,%1,%10; subc.cc.u32 %2,%2,0; subc.cc.u32 %3,%3,0;"
        "subc.cc.u32 %4,%4,0; subc.cc.u32 %5,%5,0; subc.cc.u32 %6,%6,0; subc.cc.u32 %7,%7,0; subc.u32 %8,0,0;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]),
          "+r"(r.v[7]), "=r"(bw2)
        : "r"(t), "r"(one));
    // wrapped again (bw2 = ~0): r >= 2^256 - delta now, subtracting delta once more cannot borrow; masked, not
    // branched, for the constant-time kernels
    asm("sub.cc.u32 %0,%0,%8; subc.cc.u32 %1,%1,%9; subc.cc.u32 %2,%2,0; subc.cc.u32 %3,%3,0;"
        "subc.cc.u32 %4,%4,0; subc.cc.u32 %5,%5,0; subc.cc.u32 %6,%6,0; subc.u32 %7,%7,0;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]),
          "+r"(r.v[7])
        : "r"(bw2 & S256_DELTA_LO), "r"(bw2 & 1u));
}
S256_D void fe_sub(fe &r, const fe &a, const fe &b) {
    uint32_t bw = fe_sub_raw(r, a, b);
    fe_fold_borrow(r, bw);
}

#else  // portable

S256_HD void fe_fold_carry(fe &r, uint32_t c) {
    for (int round = 0; round < 2 && c; round++) {
        uint64_t acc = (uint64_t)r.v[0] + (uint64_t)c * S256_DELTA_LO;
        r.v[0] = (uint32_t)acc;
        acc = (acc >> 32) + r.v[1] + c;
        r.v[1] = (uint32_t)acc;
        for (int i = 2; i < 8; i++) {
            acc = (acc >> 32) + r.v[i];
            r.v[i] = (uint32_t)acc;
        }
        c = (uint32_t)(acc >> 32);
    }
}
S256_HD void fe_add(fe &r, const fe &a, const fe &b) {
    uint64_t acc = 0;
    for (int i = 0; i < 8; i++) {
        acc = (acc >> 32) + a.v[i] + b.v[i];
        r.v[i] = (uint32_t)acc;
    }
    fe_fold_carry(r, (uint32_t)(acc >> 32));
}
S256_HD void fe_sub(fe &r, const fe &a, const fe &b) {
    int64_t acc = 0;
    for (int i = 0; i < 8; i++) {
        acc = (acc >> 32) + (int64_t)a.v[i] - (int64_t)b.v[i];
        r.v[i] = (uint32_t)acc;
    }
    uint32_t bw = (uint32_t)((acc >> 32) & 1);
    for (int round = 0; round < 2 && bw; round++) {
        int64_t s = (int64_t)r.v[0] - S256_DELTA_LO;
        r.v[0] = (uint32_t)s;
        s = (s >> 32) + (int64_t)r.v[1] - 1;
        r.v[1] = (uint32_t)s;
        for (int i = 2; i < 8; i++) {
            s = (s >> 32) + (int64_t)r.v[i];
            r.v[i] = (uint32_t)s;
        }
        bw = (uint32_t)((s >> 32) & 1);
    }
}
#endif

S256_HD void fe_neg(fe &r, const fe &a) {
    fe z = fe_zero();
    fe_sub(r, z, a);
}
S256_HD void fe_dbl(fe &r, const fe &a) { fe_add(r, a, a); }

// ---------------------------------------------------------------------------
// 8x8 -> 16 limb product and the special reduction
// ---------------------------------------------------------------------------
#if S256_PTX

// x[0..7] += (a0,a1,a2,a3) * b placed at limb offsets 0,2,4,6; carry -> top
#define S256_CHAIN_C(x0, x1, x2, x3, x4, x5, x6, x7, top, a0, a1, a2, a3, b)                                     \
    asm("mad.lo.cc.u32 %0,%9,%13,%0; madc.hi.cc.u32 %1,%9,%13,%1;"                                               \
        "madc.lo.cc.u32 %2,%10,%13,%2; madc.hi.cc.u32 %3,%10,%13,%3;"                                            \
        "madc.lo.cc.u32 %4,%11,%13,%4; madc.hi.cc.u32 %5,%11,%13,%5;"                                            \
        "madc.lo.cc.u32 %6,%12,%13,%6; madc.hi.cc.u32 %7,%12,%13,%7; addc.u32 %8,0,0;"                           \
        : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7), "=r"(top)              \
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b))
// same, but x7 is fresh (written, not read); cannot carry out
#define S256_CHAIN_X1(x0, x1, x2, x3, x4, x5, x6, x7, a0, a1, a2, a3, b)                                         \
    asm("mad.lo.cc.u32 %0,%8,%12,%0; madc.hi.cc.u32 %1,%8,%12,%1;"                                               \
        "madc.lo.cc.u32 %2,%9,%12,%2; madc.hi.cc.u32 %3,%9,%12,%3;"                                              \
        "madc.lo.cc.u32 %4,%10,%12,%4; madc.hi.cc.u32 %5,%10,%12,%5;"                                            \
        "madc.lo.cc.u32 %6,%11,%12,%6; madc.hi.u32 %7,%11,%12,0;"                                                \
        : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "=r"(x7)                         \
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b))
// same, but x6 and x7 are both fresh
#define S256_CHAIN_X2(x0, x1, x2, x3, x4, x5, x6, x7, a0, a1, a2, a3, b)                                         \
    asm("mad.lo.cc.u32 %0,%8,%12,%0; madc.hi.cc.u32 %1,%8,%12,%1;"                                               \
        "madc.lo.cc.u32 %2,%9,%12,%2; madc.hi.cc.u32 %3,%9,%12,%3;"                                              \
        "madc.lo.cc.u32 %4,%10,%12,%4; madc.hi.cc.u32 %5,%10,%12,%5;"                                            \
        "madc.lo.cc.u32 %6,%11,%12,0; madc.hi.u32 %7,%11,%12,0;"                                                 \
        : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "=r"(x6), "=r"(x7)                         \
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b))
#define S256_MULW(lo, hi, a, b) asm("mul.lo.u32 %0,%2,%3; mul.hi.u32 %1,%2,%3;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b))

// r[0..15] = a * b
S256_D void fe_mul_wide(uint32_t r[16], const uint32_t a[8], const uint32_t b[8]) {
    uint32_t e[16], o[15];  // o[k] sits at limb k + 1
    // row 0
    S256_MULW(e[0], e[1], a[0], b[0]);
    S256_MULW(e[2], e[3], a[2], b[0]);
    S256_MULW(e[4], e[5], a[4], b[0]);
    S256_MULW(e[6], e[7], a[6], b[0]);
    S256_MULW(o[0], o[1], a[1], b[0]);
    S256_MULW(o[2], o[3], a[3], b[0]);
    S256_MULW(o[4], o[5], a[5], b[0]);
    S256_MULW(o[6], o[7], a[7], b[0]);
    // row 1
    S256_CHAIN_C(o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], o[8], a[0], a[2], a[4], a[6], b[1]);
    S256_CHAIN_X2(e[2], e[3], e[4], e[5], e[6], e[7], e[8], e[9], a[1], a[3], a[5], a[7], b[1]);
    // row 2
    S256_CHAIN_C(e[2], e[3], e[4], e[5], e[6], e[7], e[8], e[9], e[10], a[0], a[2], a[4], a[6], b[2]);
    S256_CHAIN_X1(o[2], o[3], o[4], o[5], o[6], o[7], o[8], o[9], a[1], a[3], a[5], a[7], b[2]);
    // row 3
    S256_CHAIN_C(o[2], o[3], o[4], o[5], o[6], o[7], o[8], o[9], o[10], a[0], a[2], a[4], a[6], b[3]);
    S256_CHAIN_X1(e[4], e[5], e[6], e[7], e[8], e[9], e[10], e[11], a[1], a[3], a[5], a[7], b[3]);
    // row 4
    S256_CHAIN_C(e[4], e[5], e[6], e[7], e[8], e[9], e[10], e[11], e[12], a[0], a[2], a[4], a[6], b[4]);
    S256_CHAIN_X1(o[4], o[5], o[6], o[7], o[8], o[9], o[10], o[11], a[1], a[3], a[5], a[7], b[4]);
    // row 5
    S256_CHAIN_C(o[4], o[5], o[6], o[7], o[8], o[9], o[10], o[11], o[12], a[0], a[2], a[4], a[6], b[5]);
    S256_CHAIN_X1(e[6], e[7], e[8], e[9], e[10], e[11], e[12], e[13], a[1], a[3], a[5], a[7], b[5]);
    // row 6
    S256_CHAIN_C(e[6], e[7], e[8], e[9], e[10], e[11], e[12], e[13], e[14], a[0], a[2], a[4], a[6], b[6]);
    S256_CHAIN_X1(o[6], o[7], o[8], o[9], o[10], o[11], o[12], o[13], a[1], a[3], a[5], a[7], b[6]);
    // row 7
    S256_CHAIN_C(o[6], o[7], o[8], o[9], o[10], o[11], o[12], o[13], o[14], a[0], a[2], a[4], a[6], b[7]);
    S256_CHAIN_X1(e[8], e[9], e[10], e[11], e[12], e[13], e[14], e[15], a[1], a[3], a[5], a[7], b[7]);
    // r = e + (o << 32); the product is < 2^512 so the last add cannot carry
    r[0] = e[0];
    asm("add.cc.u32 %0,%15,%30; addc.cc.u32 %1,%16,%31; addc.cc.u32 %2,%17,%32; addc.cc.u32 %3,%18,%33;"
        "addc.cc.u32 %4,%19,%34; addc.cc.u32 %5,%20,%35; addc.cc.u32 %6,%21,%36; addc.cc.u32 %7,%22,%37;"
        "addc.cc.u32 %8,%23,%38; addc.cc.u32 %9,%24,%39; addc.cc.u32 %10,%25,%40; addc.cc.u32 %11,%26,%41;"
        "addc.cc.u32 %12,%27,%42; addc.cc.u32 %13,%28,%43; addc.u32 %14,%29,%44;"
        : "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(e[8]), "r"(e[9]),
          "r"(e[10]), "r"(e[11]), "r"(e[12]), "r"(e[13]), "r"(e[14]), "r"(e[15]), "r"(o[0]), "r"(o[1]), "r"(o[2]),
          "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]),
          "r"(o[12]), "r"(o[13]), "r"(o[14]));
}

// out = (t[0..7] + 2^256 * (t8 + 2^32 * t9)) mod-p-folded into [0, 2^256); t9 in {0,1}
// u = T * (2^32 + 977), T = t8 + t9 * 2^32 < 2^33 + 1  ->  u < 2^66
S256_D void fe_top_times_delta(uint32_t &u0, uint32_t &u1, uint32_t &u2, uint32_t t8, uint32_t t9) {
    S256_MULW(u0, u1, t8, S256_DELTA_LO);
    uint32_t t9d = (0u - t9) & S256_DELTA_LO;  // t9 in {0, 1}
    asm("add.cc.u32 %0,%0,%2; addc.u32 %1,%3,0;" : "+r"(u1), "=r"(u2) : "r"(t8), "r"(t9));
    asm("add.cc.u32 %0,%0,%2; addc.u32 %1,%1,0;" : "+r"(u1), "+r"(u2) : "r"(t9d));
}
S256_D void fe_fold_top(fe &out, uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3, uint32_t t4, uint32_t t5,
                        uint32_t t6, uint32_t t7, uint32_t t8, uint32_t t9) {
    uint32_t u0, u1, u2, c;
    fe_top_times_delta(u0, u1, u2, t8, t9);
    asm("add.cc.u32 %0,%9,%17; addc.cc.u32 %1,%10,%18; addc.cc.u32 %2,%11,%19; addc.cc.u32 %3,%12,0;"
        "addc.cc.u32 %4,%13,0; addc.cc.u32 %5,%14,0; addc.cc.u32 %6,%15,0; addc.cc.u32 %7,%16,0; addc.u32 %8,0,0;"
        : "=r"(out.v[0]), "=r"(out.v[1]), "=r"(out.v[2]), "=r"(out.v[3]), "=r"(out.v[4]), "=r"(out.v[5]),
          "=r"(out.v[6]), "=r"(out.v[7]), "=r"(c)
        : "r"(t0), "r"(t1), "r"(t2), "r"(t3), "r"(t4), "r"(t5), "r"(t6), "r"(t7), "r"(u0), "r"(u1), "r"(u2));
    // on carry out < 2^66 now: one more delta, no carry possible; masked, not branched (constant-time kernels)
    asm("add.cc.u32 %0,%0,%3; addc.cc.u32 %1,%1,%4; addc.u32 %2,%2,0;"
        : "+r"(out.v[0]), "+r"(out.v[1]), "+r"(out.v[2])
        : "r"((0u - c) & S256_DELTA_LO), "r"(c));
}

// out = r[0..15] mod p (weak)
// r[0..7] + 2^256 * (t8 + 2^32 * t9) = r[0..15] folded once
S256_D void fe_reduce_wide_pre(uint32_t r[16], uint32_t &t8, uint32_t &t9) {
    uint32_t o[8];
    const uint32_t d = S256_DELTA_LO;
    // lo += hi_even * 977 (carry -> t8)
    S256_CHAIN_C(r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], t8, r[8], r[10], r[12], r[14], d);
    // o = hi_odd * 977, sitting at limbs 1..8
    S256_MULW(o[0], o[1], r[9], d);
    S256_MULW(o[2], o[3], r[11], d);
    S256_MULW(o[4], o[5], r[13], d);
    S256_MULW(o[6], o[7], r[15], d);
    // limbs 1..8 += o   (t8 <= 1, o[7] < 977: no carry out of limb 8)
    asm("add.cc.u32 %0,%0,%8; addc.cc.u32 %1,%1,%9; addc.cc.u32 %2,%2,%10; addc.cc.u32 %3,%3,%11;"
        "addc.cc.u32 %4,%4,%12; addc.cc.u32 %5,%5,%13; addc.cc.u32 %6,%6,%14; addc.u32 %7,%7,%15;"
        : "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(t8)
        : "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]));
    // limbs 1..8 += hi (the "<< 32" half of delta); carry -> t9
    asm("add.cc.u32 %0,%0,%9; addc.cc.u32 %1,%1,%10; addc.cc.u32 %2,%2,%11; addc.cc.u32 %3,%3,%12;"
        "addc.cc.u32 %4,%4,%13; addc.cc.u32 %5,%5,%14; addc.cc.u32 %6,%6,%15; addc.cc.u32 %7,%7,%16;"
        "addc.u32 %8,0,0;"
        : "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(t8), "=r"(t9)
        : "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
}
S256_D void fe_reduce_wide(fe &out, uint32_t r[16]) {
    uint32_t t8, t9;
    fe_reduce_wide_pre(r, t8, t9);
    fe_fold_top(out, r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], t8, t9);
}

}  // namespace s256
#include "fe_mul_gen.cuh"
namespace s256 {
// The multiplier the curve code uses: the split-form core of fe_mul_gen.cuh (no register re-pairing, see
// tools/gen_fe_mul.py).  -DS256_MUL_MERGED selects the first form (merge the accumulators, then reduce) for A/B runs;
// fe_mul_wide itself stays in use as the 8x8 product of the Z_n arithmetic (sc.cuh).
S256_D void fe_mul_inline(fe &r, const fe &a, const fe &b) {
#ifndef S256_MUL_MERGED
    uint32_t w[9], t9;
    fe_mul_core(w, t9, a.v, b.v);
    fe_fold_top(r, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], w[8], t9);
#else
    uint32_t w[16];
    fe_mul_wide(w, a.v, b.v);
    fe_reduce_wide(r, w);
#endif
}
#ifndef S256_NO_SQR
}  // namespace s256
#include "fe_sqr_gen.cuh"
namespace s256 {
// dedicated squaring: 36 + 9 MAC32 (tools/gen_fe_sqr.py, tools/gen_fe_mul.py)
S256_D void fe_sqr_inline(fe &r, const fe &a) {
#ifndef S256_MUL_MERGED
    uint32_t w[9], t9;
    fe_sqr_core(w, t9, a.v);
    fe_fold_top(r, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], w[8], t9);
#else
    uint32_t w[16];
    fe_sqr_wide(w, a.v);
    fe_reduce_wide(r, w);
#endif
}
#else
S256_D void fe_sqr_inline(fe &r, const fe &a) { fe_mul_inline(r, a, a); }
#endif
#ifndef S256_MUL_INLINE
// Out of line on purpose: one ~2.4 KB body shared by every call site stays
// resident in the instruction caches (the fully inlined ladder is ~140 KB of
// straight-line code and stalls on instruction fetch, see profiles/), and the
// arguments travel in registers (no stack traffic).
static __device__ __noinline__ fe fe_mul_call(fe a, fe b) {
    fe r;
    fe_mul_inline(r, a, b);
    return r;
}
static __device__ __noinline__ fe fe_sqr_call(fe a) {
    fe r;
    fe_sqr_inline(r, a);
    return r;
}
S256_D void fe_mul(fe &r, const fe &a, const fe &b) { r = fe_mul_call(a, b); }
S256_D void fe_sqr(fe &r, const fe &a) { r = fe_sqr_call(a); }
#else
S256_D void fe_mul(fe &r, const fe &a, const fe &b) { fe_mul_inline(r, a, b); }
S256_D void fe_sqr(fe &r, const fe &a) { fe_sqr_inline(r, a); }
#endif

// r = a * k for a small constant k (< 2^16): 8 IMAD.WIDE + one fold
S256_D void fe_mul_small_pre(uint32_t e[8], uint32_t &t8, const fe &a, uint32_t k) {
    uint32_t o[8];
    S256_MULW(e[0], e[1], a.v[0], k);
    S256_MULW(e[2], e[3], a.v[2], k);
    S256_MULW(e[4], e[5], a.v[4], k);
    S256_MULW(e[6], e[7], a.v[6], k);
    S256_MULW(o[0], o[1], a.v[1], k);
    S256_MULW(o[2], o[3], a.v[3], k);
    S256_MULW(o[4], o[5], a.v[5], k);
    S256_MULW(o[6], o[7], a.v[7], k);
    asm("add.cc.u32 %0,%0,%8; addc.cc.u32 %1,%1,%9; addc.cc.u32 %2,%2,%10; addc.cc.u32 %3,%3,%11;"
        "addc.cc.u32 %4,%4,%12; addc.cc.u32 %5,%5,%13; addc.cc.u32 %6,%6,%14; addc.u32 %7,%15,0;"
        : "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]), "=r"(t8)
        : "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]));
}
S256_D void fe_mul_small(fe &r, const fe &a, uint32_t k) {
    uint32_t e[8], t8;
    fe_mul_small_pre(e, t8, a, k);
    fe_fold_top(r, e[0], e[1], e[2], e[3], e[4], e[5], e[6], e[7], t8, 0u);
}

#else  // portable

S256_HD void fe_reduce_wide_portable(fe &out, const uint32_t w[16]) {
    // t = lo + hi * (2^32 + 977): 10 limbs
    uint32_t t[10];
    uint64_t acc = 0;
    for (int i = 0; i < 9; i++) {
        acc += (i < 8) ? w[i] : 0;
        if (i < 8) acc += (uint64_t)w[8 + i] * S256_DELTA_LO;
        uint64_t carry = acc >> 32;
        uint64_t lowpart = acc & 0xFFFFFFFFu;
        if (i >= 1) lowpart += w[8 + i - 1];
        t[i] = (uint32_t)lowpart;
        acc = carry + (lowpart >> 32);
    }
    t[9] = (uint32_t)acc;
    // fold T = t8 + t9 * 2^32 (< 2^34) once more
    uint64_t T = (uint64_t)t[8] | ((uint64_t)t[9] << 32);
    fe r;
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
    // u = T * 977 + (T << 32), three limbs and a bit
    unsigned __int128 u = (unsigned __int128)T * S256_DELTA_LO + ((unsigned __int128)T << 32);
    acc = 0;
    for (int i = 0; i < 8; i++) {
        acc += r.v[i];
        if (i < 4) acc += (uint32_t)(u >> (32 * i));
        r.v[i] = (uint32_t)acc;
        acc >>= 32;
    }
    fe_fold_carry(r, (uint32_t)acc);
    out = r;
}
S256_HD void fe_mul(fe &r, const fe &a, const fe &b) {
    uint32_t w[16];
    for (int i = 0; i < 16; i++) w[i] = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t carry = 0;
        for (int j = 0; j < 8; j++) {
            uint64_t t = (uint64_t)a.v[i] * b.v[j] + w[i + j] + carry;
            w[i + j] = (uint32_t)t;
            carry = t >> 32;
        }
        w[i + 8] = (uint32_t)carry;
    }
    fe_reduce_wide_portable(r, w);
}
S256_HD void fe_sqr(fe &r, const fe &a) { fe_mul(r, a, a); }
S256_HD void fe_mul_inline(fe &r, const fe &a, const fe &b) { fe_mul(r, a, b); }
S256_HD void fe_sqr_inline(fe &r, const fe &a) { fe_mul(r, a, a); }
S256_HD void fe_mul_small(fe &r, const fe &a, uint32_t k) {
    fe kk = fe_from_u32(k);
    fe_mul(r, a, kk);
}
#endif

// ---------------------------------------------------------------------------
// canonical form, predicates, selects
// ---------------------------------------------------------------------------

// r = a mod p, canonical (< p): a >= p  <=>  a + delta carries out of 2^256
S256_HD void fe_normalize(fe &r, const fe &a) {
    uint32_t s[8];
    uint64_t acc = (uint64_t)a.v[0] + S256_DELTA_LO;
    s[0] = (uint32_t)acc;
    acc = (acc >> 32) + a.v[1] + 1u;
    s[1] = (uint32_t)acc;
#pragma unroll
    for (int i = 2; i < 8; i++) {
        acc = (acc >> 32) + a.v[i];
        s[i] = (uint32_t)acc;
    }
    uint32_t ge = (uint32_t)(acc >> 32);  // 1 iff a >= p
    uint32_t m = 0u - ge;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (s[i] & m) | (a.v[i] & ~m);
}
// 1 iff a == 0 (mod p)
S256_HD uint32_t fe_is_zero(const fe &a) {
    uint32_t z = a.v[0] | a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5] | a.v[6] | a.v[7];
    uint32_t q = (a.v[0] ^ S256_P0) | (a.v[1] ^ S256_P1) | ~a.v[2] | ~a.v[3] | ~a.v[4] | ~a.v[5] | ~a.v[6] | ~a.v[7];
    return (uint32_t)((z == 0) | (q == 0));
}
S256_HD uint32_t fe_equal(const fe &a, const fe &b) {
    fe d;
    fe_sub(d, a, b);
    return fe_is_zero(d);
}
// parity of the canonical value (field.go:191-197)
S256_HD uint32_t fe_is_odd(const fe &a) {
    fe n;
    fe_normalize(n, a);
    return n.v[0] & 1u;
}
// r = ctrl ? b : a, branch-free (field.go:172-175 ConditionalSelect)
S256_HD void fe_cmov(fe &r, const fe &a, const fe &b, uint32_t ctrl) {
    uint32_t m = 0u - (ctrl & 1u);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (a.v[i] & ~m) | (b.v[i] & m);
}

// 1 iff the 8-limb value is < p (canonical); limbs straight from bytes
S256_HD uint32_t fe_limbs_are_canonical(const fe &a) {
    uint64_t acc = (uint64_t)a.v[0] + S256_DELTA_LO;
    acc = (acc >> 32) + a.v[1] + 1u;
#pragma unroll
    for (int i = 2; i < 8; i++) acc = (acc >> 32) + a.v[i];
    return 1u - (uint32_t)(acc >> 32);
}

// big-endian 32 bytes <-> limbs (internal/helpers/helpers.go:47-65)
S256_HD void fe_from_be32(fe &r, const uint8_t *b) {
#if defined(__CUDA_ARCH__)
    if ((((size_t)b) & 15u) == 0) {  // aligned rows (x-only keys, signatures): two 128-bit loads
        const uint4 hi = *reinterpret_cast<const uint4 *>(b), lo = *reinterpret_cast<const uint4 *>(b + 16);
        r.v[7] = __byte_perm(hi.x, 0, 0x0123); r.v[6] = __byte_perm(hi.y, 0, 0x0123);
        r.v[5] = __byte_perm(hi.z, 0, 0x0123); r.v[4] = __byte_perm(hi.w, 0, 0x0123);
        r.v[3] = __byte_perm(lo.x, 0, 0x0123); r.v[2] = __byte_perm(lo.y, 0, 0x0123);
        r.v[1] = __byte_perm(lo.z, 0, 0x0123); r.v[0] = __byte_perm(lo.w, 0, 0x0123);
        return;
    }
#endif
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint8_t *q = b + 4 * (7 - i);
        r.v[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
    }
}
// caller passes a canonical element
S256_HD void fe_to_be32(uint8_t *b, const fe &a) {
#if defined(__CUDA_ARCH__)
    // a 16-byte aligned destination (32- and 64-byte rows of the batch buffers): two 128-bit stores
    if ((((size_t)b) & 15u) == 0) {
        uint4 hi, lo;
        hi.x = __byte_perm(a.v[7], 0, 0x0123); hi.y = __byte_perm(a.v[6], 0, 0x0123);
        hi.z = __byte_perm(a.v[5], 0, 0x0123); hi.w = __byte_perm(a.v[4], 0, 0x0123);
        lo.x = __byte_perm(a.v[3], 0, 0x0123); lo.y = __byte_perm(a.v[2], 0, 0x0123);
        lo.z = __byte_perm(a.v[1], 0, 0x0123); lo.w = __byte_perm(a.v[0], 0, 0x0123);
        reinterpret_cast<uint4 *>(b)[0] = hi;
        reinterpret_cast<uint4 *>(b)[1] = lo;
        return;
    }
#endif
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint8_t *q = b + 4 * (7 - i);
        q[0] = (uint8_t)(a.v[i] >> 24);
        q[1] = (uint8_t)(a.v[i] >> 16);
        q[2] = (uint8_t)(a.v[i] >> 8);
        q[3] = (uint8_t)a.v[i];
    }
}

// ---------------------------------------------------------------------------
// exponentiation chains (per thread).  Same shape as the reference's
// addchain-generated routines: internal/field/field_invert.go:11-140
// (x^(p-2), 255 S + 15 M, Invert(0) = 0) and field_sqrt_ratio.go:65-185
// (x^((p+1)/4), 253 S + 13 M).
// ---------------------------------------------------------------------------
S256_HD void fe_sqr_n(fe &r, const fe &a, int n) {
    r = a;
#pragma unroll 1
    for (int i = 0; i < n; i++) fe_sqr(r, r);
}
S256_HD void fe_pow_x223(fe &x223, fe &x22, fe &x2, fe &x3, const fe &a) {
    fe t, x6, x9, x11, x44, x88, x176;
    fe_sqr(t, a); fe_mul(x2, t, a);
    fe_sqr(t, x2); fe_mul(x3, t, a);
    fe_sqr_n(t, x3, 3); fe_mul(x6, t, x3);
    fe_sqr_n(t, x6, 3); fe_mul(x9, t, x3);
    fe_sqr_n(t, x9, 2); fe_mul(x11, t, x2);
    fe_sqr_n(t, x11, 11); fe_mul(x22, t, x11);
    fe_sqr_n(t, x22, 22); fe_mul(x44, t, x22);
    fe_sqr_n(t, x44, 44); fe_mul(x88, t, x44);
    fe_sqr_n(t, x88, 88); fe_mul(x176, t, x88);
    fe_sqr_n(t, x176, 44); fe_mul(t, t, x44);
    fe_sqr_n(t, t, 3); fe_mul(x223, t, x3);
}
// x^(p-2) by the reference's addition chain (field_invert.go:11): kept as the cross-check of fe_invert
S256_HD void fe_invert_fermat(fe &r, const fe &a) {
    fe x223, x22, x2, x3, t;
    fe_pow_x223(x223, x22, x2, x3, a);
    fe_sqr_n(t, x223, 23); fe_mul(t, t, x22);
    fe_sqr_n(t, t, 5); fe_mul(t, t, a);
    fe_sqr_n(t, t, 3); fe_mul(t, t, x2);
    fe_sqr_n(t, t, 2); fe_mul(r, t, a);
}
// returns 1 and a root iff a is a square; else 0 and r = 0
S256_HD uint32_t fe_sqrt(fe &r, const fe &a) {
    fe x223, x22, x2, x3, t, chk;
    fe_pow_x223(x223, x22, x2, x3, a);
    fe_sqr_n(t, x223, 23); fe_mul(t, t, x22);
    fe_sqr_n(t, t, 6); fe_mul(t, t, x2);
    fe_sqr_n(t, t, 2);
    fe_sqr(chk, t);
    uint32_t ok = fe_equal(chk, a);
    fe z = fe_zero();
    fe_cmov(r, z, t, ok);
    return ok;
}

}  // namespace s256

#include "modinv.cuh"

namespace s256 {
// a^-1 mod p, Invert(0) = 0 (field_invert.go:11): safegcd (modinv.cuh), constant time
S256_HD void fe_invert(fe &r, const fe &a) {
    fe t;
    fe_normalize(t, a);
    mi_invert(r.v, t.v, mi_modulus_p());
}
}  // namespace s256
