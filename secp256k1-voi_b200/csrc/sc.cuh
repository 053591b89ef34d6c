// sc.cuh -- arithmetic in Z_n (the secp256k1 group order) and scalar recoding.
//
// Replaces the reference's fiat-crypto scalar Montgomery code behind scalar.go
// (scalar.go:66-206, scalar_invert.go:11) and the GLV split of
// point_mul_glv.go:59-189.  Plain (non-Montgomery) 8 x 32-bit little-endian
// limbs, always fully reduced to [0, n).  n = 2^256 - NC with NC a 129-bit
// constant, so a 512-bit product is reduced by folding hi * NC three times.
// All routines are branch-free on the scalar value (masks, no data-dependent
// control flow): the constant-time entry points feed secrets through them.
//
// Z_n work is ~2 % of a verification, so this is portable C (the compiler
// emits IMAD.WIDE.U32 + IADD3 for the 32x32->64 accumulations); the same code
// compiles for the CPU-side host simulation used by tests/hostsim.
#pragma once
#include "fe.cuh"

namespace s256 {

struct sc {
    uint32_t v[8];
};

// Constants live twice under nvcc (a __constant__ copy for device code, a plain
// copy for host code); S256_K(name) picks the right one for the current pass.
#if defined(__CUDACC__)
#define S256_CONST(name, n, ...)                                        \
    static __device__ __constant__ const uint32_t name##_d[n] = {__VA_ARGS__}; \
    static const uint32_t name##_h[n] = {__VA_ARGS__};
#else
#define S256_CONST(name, n, ...) static const uint32_t name##_h[n] = {__VA_ARGS__};
#endif
#if defined(__CUDA_ARCH__)
#define S256_K(name) name##_d
#define S256_NOINLINE __noinline__
#else
#define S256_K(name) name##_h
#define S256_NOINLINE
#endif

// n, 2^256 - n, floor(n / 2)  (scalar.go:17-38)
S256_CONST(SC_N, 8, 0xD0364141u, 0xBFD25E8Cu, 0xAF48A03Bu, 0xBAAEDCE6u, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu)
S256_CONST(SC_NC, 5, 0x2FC9BEBFu, 0x402DA173u, 0x50B75FC4u, 0x45512319u, 0x00000001u)
S256_CONST(SC_HALF_N, 8, 0x681B20A0u, 0xDFE92F46u, 0x57A4501Du, 0x5D576E73u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x7FFFFFFFu)
// GLV constants (point_mul_glv.go:41-56)
S256_CONST(SC_NEG_LAMBDA, 8, 0xB51283CFu, 0xE0CFC810u, 0x8EC739C2u, 0xA880B9FCu, 0x77ED9BA4u, 0x5AD9E3FDu, 0x3FA3CF1Fu, 0xAC9C52B3u)
S256_CONST(SC_NEG_B1, 8, 0x0ABFE4C3u, 0x6F547FA9u, 0x010E8828u, 0xE4437ED6u, 0, 0, 0, 0)
S256_CONST(SC_NEG_B2, 8, 0x3DB1562Cu, 0xD765CDA8u, 0x0774346Du, 0x8A280AC5u, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu)
S256_CONST(SC_G1, 8, 0x45DBB031u, 0xE893209Au, 0x71E8CA7Fu, 0x3DAA8A14u, 0x9284EB15u, 0xE86C90E4u, 0xA7D46BCDu, 0x3086D221u)
S256_CONST(SC_G2, 8, 0x8AC47F71u, 0x1571B4AEu, 0x9DF506C6u, 0x221208ACu, 0x0ABFE4C4u, 0x6F547FA9u, 0x010E8828u, 0xE4437ED6u)

S256_HD sc sc_zero() {
    sc r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}
S256_HD sc sc_one() {
    sc r = sc_zero();
    r.v[0] = 1;
    return r;
}
S256_HD sc sc_const(const uint32_t *c) {
    sc r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = c[i];
    return r;
}
S256_HD uint32_t sc_is_zero(const sc &a) {
    return (uint32_t)((a.v[0] | a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5] | a.v[6] | a.v[7]) == 0);
}
S256_HD uint32_t sc_equal(const sc &a, const sc &b) {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) d |= a.v[i] ^ b.v[i];
    return (uint32_t)(d == 0);
}
// r = a - m (8 limbs), returns borrow (1 iff a < m)
S256_HD uint32_t sc_sub_limbs(uint32_t r[8], const uint32_t a[8], const uint32_t m[8]) {
    int64_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        acc = (acc >> 32) + (int64_t)a[i] - (int64_t)m[i];
        r[i] = (uint32_t)acc;
    }
    return (uint32_t)((acc >> 32) & 1);
}
// a (any 256-bit value < 2n) -> a mod n; returns 1 iff a subtraction happened
S256_HD uint32_t sc_reduce_once(sc &r, const uint32_t a[8], uint32_t carry_in) {
    uint32_t d[8];
    uint32_t borrow = sc_sub_limbs(d, a, S256_K(SC_N));
    uint32_t ge = carry_in | (1u - borrow);
    uint32_t m = 0u - ge;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (d[i] & m) | (a[i] & ~m);
    return ge;
}
// scalar.go:123-131 SetBytes: big-endian bytes, reduced once; returns didReduce
S256_HD uint32_t sc_from_be32(sc &r, const uint8_t *b) {
    uint32_t l[8];
#if S256_PTX
    // rows of the batch buffers are 16-byte aligned: two 128-bit loads instead of 32 byte loads
    if ((((size_t)b) & 15u) == 0) {
        const uint4 hi = *reinterpret_cast<const uint4 *>(b), lo = *reinterpret_cast<const uint4 *>(b + 16);
        l[7] = __byte_perm(hi.x, 0, 0x0123); l[6] = __byte_perm(hi.y, 0, 0x0123);
        l[5] = __byte_perm(hi.z, 0, 0x0123); l[4] = __byte_perm(hi.w, 0, 0x0123);
        l[3] = __byte_perm(lo.x, 0, 0x0123); l[2] = __byte_perm(lo.y, 0, 0x0123);
        l[1] = __byte_perm(lo.z, 0, 0x0123); l[0] = __byte_perm(lo.w, 0, 0x0123);
        return sc_reduce_once(r, l, 0);
    }
#endif
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint8_t *q = b + 4 * (7 - i);
        l[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
    }
    return sc_reduce_once(r, l, 0);
}
S256_HD void sc_to_be32(uint8_t *b, const sc &a) {
#if defined(__CUDA_ARCH__)
    if ((((size_t)b) & 15u) == 0) {  // aligned rows: two 128-bit stores
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
// scalar.go:190-206
S256_HD uint32_t sc_is_gt_half_n(const sc &a) {
    uint32_t d[8];
    uint32_t borrow = sc_sub_limbs(d, a.v, S256_K(SC_HALF_N));
    uint32_t nz = d[0] | d[1] | d[2] | d[3] | d[4] | d[5] | d[6] | d[7];
    return (1u - borrow) & (uint32_t)(nz != 0);
}
S256_HD void sc_add(sc &r, const sc &a, const sc &b) {
    uint32_t s[8];
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        acc = (acc >> 32) + a.v[i] + b.v[i];
        s[i] = (uint32_t)acc;
    }
    sc_reduce_once(r, s, (uint32_t)(acc >> 32));
}
S256_HD void sc_neg(sc &r, const sc &a) {
    uint32_t d[8];
    sc_sub_limbs(d, S256_K(SC_N), a.v);
    uint32_t m = 0u - (1u - sc_is_zero(a));
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = d[i] & m;
}
S256_HD void sc_cmov(sc &r, const sc &a, const sc &b, uint32_t ctrl) {
    uint32_t m = 0u - (ctrl & 1u);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (a.v[i] & ~m) | (b.v[i] & m);
}

// out[0..LO+?] = lo[0..7] + hi[0..HN-1] * NC ; OUTN limbs written
template <int HN, int OUTN>
S256_HD void sc_fold(uint32_t *out, const uint32_t *lo, const uint32_t *hi) {
    uint32_t t[OUTN];
#pragma unroll
    for (int i = 0; i < OUTN; i++) t[i] = (i < 8) ? lo[i] : 0u;
#pragma unroll
    for (int i = 0; i < HN; i++) {
        uint64_t carry = 0;
#pragma unroll
        for (int j = 0; j < 5; j++) {
            if (i + j < OUTN) {
                uint64_t x = (uint64_t)hi[i] * S256_K(SC_NC)[j] + t[i + j] + carry;
                t[i + j] = (uint32_t)x;
                carry = x >> 32;
            }
        }
#pragma unroll
        for (int k = i + 5; k < OUTN; k++) {
            uint64_t x = (uint64_t)t[k] + carry;
            t[k] = (uint32_t)x;
            carry = x >> 32;
        }
    }
#pragma unroll
    for (int i = 0; i < OUTN; i++) out[i] = t[i];
}

// w[0..15] -> w mod n
S256_HD void sc_reduce512(sc &r, const uint32_t w[16]) {
    uint32_t x[14], y[10], z[9];
    sc_fold<8, 14>(x, w, w + 8);  // < 2^386
    sc_fold<6, 10>(y, x, x + 8);  // x >> 256 < 2^130 (limbs 8..13; 13 is 0) ; y < 2^260
    sc_fold<1, 9>(z, y, y + 8);   // y >> 256 < 2^4 ; z < 2^256 + 2^134
    // z[8] in {0,1}: fold it (z_lo < 2^134 when set, so no further carry)
    uint32_t c = z[8];
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        acc = (acc >> 32) + z[i] + ((i < 5) ? (uint64_t)(S256_K(SC_NC)[i] & (0u - c)) : 0u);
        z[i] = (uint32_t)acc;
    }
    sc_reduce_once(r, z, 0);
}
S256_HD void sc_mul_wide(uint32_t w[16], const uint32_t a[8], const uint32_t b[8]) {
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t carry = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            uint64_t t = (uint64_t)a[i] * b[j] + w[i + j] + carry;
            w[i + j] = (uint32_t)t;
            carry = t >> 32;
        }
        w[i + 8] = (uint32_t)carry;
    }
}
#if S256_PTX
// Device form of the product and of the first two folds: the same even/odd mad.cc carry chains as
// fe_mul_wide (fe.cuh) instead of 64-bit C arithmetic, whose dependent IMAD / IADD3 / IADD3.X triples
// made one sc_mul ~3100 cycles of latency -- and the batched-inversion kernel, 600 dependent sc_mul
// per thread on few threads, is latency bound.
// out[0..12] = lo[0..7] + hi[0..7] * (2^256 - n),  2^256 - n = NC[0..3] + 2^128
S256_D void sc_fold_ptx(uint32_t out[13], const uint32_t lo[8], const uint32_t hi[8]) {
    const uint32_t c0 = 0x2FC9BEBFu, c1 = 0x402DA173u, c2 = 0x50B75FC4u, c3 = 0x45512319u;
    uint32_t e[12], o[11];
    // hi * NC[0..3]: rows 0..3 of the schoolbook product
    S256_MULW(e[0], e[1], hi[0], c0);
    S256_MULW(e[2], e[3], hi[2], c0);
    S256_MULW(e[4], e[5], hi[4], c0);
    S256_MULW(e[6], e[7], hi[6], c0);
    S256_MULW(o[0], o[1], hi[1], c0);
    S256_MULW(o[2], o[3], hi[3], c0);
    S256_MULW(o[4], o[5], hi[5], c0);
    S256_MULW(o[6], o[7], hi[7], c0);
    S256_CHAIN_C(o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], o[8], hi[0], hi[2], hi[4], hi[6], c1);
    S256_CHAIN_X2(e[2], e[3], e[4], e[5], e[6], e[7], e[8], e[9], hi[1], hi[3], hi[5], hi[7], c1);
    S256_CHAIN_C(e[2], e[3], e[4], e[5], e[6], e[7], e[8], e[9], e[10], hi[0], hi[2], hi[4], hi[6], c2);
    S256_CHAIN_X1(o[2], o[3], o[4], o[5], o[6], o[7], o[8], o[9], hi[1], hi[3], hi[5], hi[7], c2);
    S256_CHAIN_C(o[2], o[3], o[4], o[5], o[6], o[7], o[8], o[9], o[10], hi[0], hi[2], hi[4], hi[6], c3);
    S256_CHAIN_X1(e[4], e[5], e[6], e[7], e[8], e[9], e[10], e[11], hi[1], hi[3], hi[5], hi[7], c3);
    // P = e + (o << 32): 12 limbs, no carry out (P < 2^384)
    asm("add.cc.u32 %0,%0,%11; addc.cc.u32 %1,%1,%12; addc.cc.u32 %2,%2,%13; addc.cc.u32 %3,%3,%14;"
        "addc.cc.u32 %4,%4,%15; addc.cc.u32 %5,%5,%16; addc.cc.u32 %6,%6,%17; addc.cc.u32 %7,%7,%18;"
        "addc.cc.u32 %8,%8,%19; addc.cc.u32 %9,%9,%20; addc.u32 %10,%10,%21;"
        : "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]), "+r"(e[8]), "+r"(e[9]),
          "+r"(e[10]), "+r"(e[11])
        : "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(o[8]), "r"(o[9]),
          "r"(o[10]));
    // + lo (limbs 0..7), carry rippling through limbs 8..11 into limb 12
    asm("add.cc.u32 %0,%0,%13; addc.cc.u32 %1,%1,%14; addc.cc.u32 %2,%2,%15; addc.cc.u32 %3,%3,%16;"
        "addc.cc.u32 %4,%4,%17; addc.cc.u32 %5,%5,%18; addc.cc.u32 %6,%6,%19; addc.cc.u32 %7,%7,%20;"
        "addc.cc.u32 %8,%8,0; addc.cc.u32 %9,%9,0; addc.cc.u32 %10,%10,0; addc.cc.u32 %11,%11,0; addc.u32 %12,0,0;"
        : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]), "+r"(e[8]),
          "+r"(e[9]), "+r"(e[10]), "+r"(e[11]), "=r"(out[12])
        : "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]));
    // + hi << 128 (limbs 4..11)
    asm("add.cc.u32 %0,%0,%9; addc.cc.u32 %1,%1,%10; addc.cc.u32 %2,%2,%11; addc.cc.u32 %3,%3,%12;"
        "addc.cc.u32 %4,%4,%13; addc.cc.u32 %5,%5,%14; addc.cc.u32 %6,%6,%15; addc.cc.u32 %7,%7,%16; addc.u32 %8,%8,0;"
        : "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]), "+r"(e[8]), "+r"(e[9]), "+r"(e[10]), "+r"(e[11]), "+r"(out[12])
        : "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]));
#pragma unroll
    for (int i = 0; i < 12; i++) out[i] = e[i];
}
S256_D void sc_reduce512_ptx(sc &r, const uint32_t w[16]) {
    uint32_t x[13], y[13], z[9], h2[8];
    sc_fold_ptx(x, w, w + 8);  // < 2^386: limbs 0..12, x[12] <= 2
#pragma unroll
    for (int i = 0; i < 8; i++) h2[i] = i < 5 ? x[8 + i] : 0u;
    sc_fold_ptx(y, x, h2);     // x >> 256 < 2^130  ->  y < 2^260: y[8] < 16, y[9..12] = 0
    sc_fold<1, 9>(z, y, y + 8);
    uint32_t c = z[8];
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        acc = (acc >> 32) + z[i] + ((i < 5) ? (uint64_t)(S256_K(SC_NC)[i] & (0u - c)) : 0u);
        z[i] = (uint32_t)acc;
    }
    sc_reduce_once(r, z, 0);
}
#endif

#if S256_PTX
// out of line (one body for ~60 call sites), operands and result in registers: passing references
// would pin every scalar to the local-memory stack
static __device__ __noinline__ sc sc_mul_call(sc a, sc b) {
    uint32_t w[16];
    sc r;
    fe_mul_wide(w, a.v, b.v);
    sc_reduce512_ptx(r, w);
    return r;
}
S256_D void sc_mul(sc &r, const sc &a, const sc &b) { r = sc_mul_call(a, b); }
// 36 instead of 64 products (fe_sqr_gen.cuh): the inversion chain is 253 squarings
static __device__ __noinline__ sc sc_sqr_call(sc a) {
    uint32_t w[16];
    sc r;
    fe_sqr_wide(w, a.v);
    sc_reduce512_ptx(r, w);
    return r;
}
S256_D void sc_sqr(sc &r, const sc &a) { r = sc_sqr_call(a); }
#else
#if defined(__CUDACC__)
static __host__ __device__ S256_NOINLINE
#else
inline
#endif
void sc_mul(sc &r, const sc &a, const sc &b) {
    uint32_t w[16];
    sc_mul_wide(w, a.v, b.v);
    sc_reduce512(r, w);
}
S256_HD void sc_sqr(sc &r, const sc &a) { sc_mul(r, a, a); }
#endif

// scalar_invert.go:11-303 -- x^(n-2), Invert(0) = 0.  The top 127 exponent
// bits are ones (run-of-ones chain), the low 129 go through a 4-bit window.
#if defined(__CUDACC__)
static __host__ __device__ S256_NOINLINE
#else
inline
#endif
void sc_invert_fermat(sc &r, const sc &a) {
    sc tbl[16];
    tbl[0] = sc_one();
    tbl[1] = a;
#pragma unroll 1
    for (int i = 2; i < 16; i++) sc_mul(tbl[i], tbl[i - 1], a);
    sc t, x7, x14, x28, x56, x63, x;
    // 2^127 - 1
    sc_sqr(t, tbl[3]);            // 6
    sc_mul(x7, t, a);             // 7 -> a^7 = ones(3)
    // ones(k): a^(2^k - 1)
    sc o3 = x7, o6, o7;
    t = o3;
#pragma unroll 1
    for (int i = 0; i < 3; i++) sc_sqr(t, t);
    sc_mul(o6, t, o3);
    sc_sqr(t, o6);
    sc_mul(o7, t, a);
    t = o7;
#pragma unroll 1
    for (int i = 0; i < 7; i++) sc_sqr(t, t);
    sc_mul(x14, t, o7);
    t = x14;
#pragma unroll 1
    for (int i = 0; i < 14; i++) sc_sqr(t, t);
    sc_mul(x28, t, x14);
    t = x28;
#pragma unroll 1
    for (int i = 0; i < 28; i++) sc_sqr(t, t);
    sc_mul(x56, t, x28);
    t = x56;
#pragma unroll 1
    for (int i = 0; i < 7; i++) sc_sqr(t, t);
    sc_mul(x63, t, o7);
    t = x63;
#pragma unroll 1
    for (int i = 0; i < 63; i++) sc_sqr(t, t);
    sc_mul(t, t, x63);  // ones(126)
    sc_sqr(t, t);
    sc_mul(x, t, a);    // ones(127)
    sc_sqr(x, x);       // exponent bit 128 of n-2 is 0
    // low 128 bits of n - 2: limbs 3..0 of n, with the last nibble 1 -> f (0x...4141 - 2 = 0x...413F)
#pragma unroll 1
    for (int i = 31; i >= 0; i--) {
#pragma unroll 1
        for (int k = 0; k < 4; k++) sc_sqr(x, x);
        uint32_t limb = S256_K(SC_N)[i >> 3];
        if ((i >> 3) == 0) limb -= 2u;
        uint32_t nib = (limb >> ((i & 7) * 4)) & 0xFu;
        if (nib) sc_mul(x, x, tbl[nib]);  // exponent is public
    }
    r = x;
}
// a^-1 mod n, Invert(0) = 0 (scalar_invert.go:11): safegcd (modinv.cuh), constant time; a is canonical
#if defined(__CUDACC__)
static __host__ __device__ S256_NOINLINE
#else
inline
#endif
void sc_invert(sc &r, const sc &a) {
    mi_invert(r.v, a.v, mi_modulus_n());
}

// point_mul_glv.go:119-189 -- round(k * g / 2^384): limbs 12..15 of the
// product plus the rounding bit (bit 383).
S256_HD void sc_mul_shift384(sc &r, const sc &k, const uint32_t *g) {
    uint32_t w[16], gg[8];
#pragma unroll
    for (int i = 0; i < 8; i++) gg[i] = g[i];
    sc_mul_wide(w, k.v, gg);
    uint64_t acc = (uint64_t)w[12] + (w[11] >> 31);
    r.v[0] = (uint32_t)acc;
    acc = (acc >> 32) + w[13];
    r.v[1] = (uint32_t)acc;
    acc = (acc >> 32) + w[14];
    r.v[2] = (uint32_t)acc;
    acc = (acc >> 32) + w[15];
    r.v[3] = (uint32_t)acc;
    r.v[4] = r.v[5] = r.v[6] = r.v[7] = 0;
}
// point_mul_glv.go:59-117 -- k = k1 + k2 * lambda (mod n)
S256_HD void sc_split_glv(sc &k1, sc &k2, const sc &k) {
    sc c1, c2, t;
    sc_mul_shift384(c1, k, S256_K(SC_G1));
    sc_mul_shift384(c2, k, S256_K(SC_G2));
    sc_mul(k2, c1, sc_const(S256_K(SC_NEG_B1)));
    sc_mul(t, c2, sc_const(S256_K(SC_NEG_B2)));
    sc_add(k2, k2, t);
    sc_mul(k1, k2, sc_const(S256_K(SC_NEG_LAMBDA)));
    sc_add(k1, k, k1);
}
// Split plus the sign normalisation of point_mul_glv.go:213-220 / :263-269:
// returns magnitudes < 2^128 (4 limbs each) and the two negate flags.
S256_HD void sc_split_glv_abs(uint32_t m1[4], uint32_t &neg1, uint32_t m2[4], uint32_t &neg2, const sc &k) {
    sc k1, k2, n1, n2;
    sc_split_glv(k1, k2, k);
    neg1 = sc_is_gt_half_n(k1);
    neg2 = sc_is_gt_half_n(k2);
    sc_neg(n1, k1);
    sc_neg(n2, k2);
    sc_cmov(k1, k1, n1, neg1);
    sc_cmov(k2, k2, n2, neg2);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        m1[i] = k1.v[i];
        m2[i] = k2.v[i];
    }
}

// Signed fixed-window recoding of a magnitude m < 2^128, window W bits:
// m = sum d_i * 2^(W*i), d_i in [-(2^(W-1) - 1), 2^(W-1)], ND = ceil(129 / W)
// digits.  Branch-free.  Digits are stored as int8.
template <int W>
struct glv_recode {
    static constexpr int ND = (129 + W - 1) / W;
    S256_HD static void run(int8_t *d, const uint32_t m[4]) {
        uint32_t carry = 0;
#pragma unroll
        for (int i = 0; i < ND; i++) {
            int bit = i * W;
            uint32_t v = 0;
            if (bit < 128) {
                int limb = bit >> 5, sh = bit & 31;
                v = m[limb] >> sh;
                if (sh + W > 32 && limb + 1 < 4) v |= m[limb + 1] << (32 - sh);
                v &= (1u << W) - 1u;
            }
            v += carry;                                   // 0 .. 2^W
            carry = (v + (1u << (W - 1)) - 1u) >> W;      // 1 iff v > 2^(W-1)
            d[i] = (int8_t)((int32_t)v - (int32_t)(carry << W));
        }
    }
};

}  // namespace s256
