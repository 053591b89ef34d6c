// microbench.cuh -- integer-pipe throughput probes (roofline denominators).
// Operands are loop-variant (fed from other accumulators) so ptxas cannot
// strength-reduce the multiplies; the SASS of every variant is checked in
// profiles/r01_microbench_sass.txt.  The host reports "ops" per second over
// all SMs, an op being one wide multiply-add (or one 32-bit op for MB_ADDC).
#pragma once
#include <stdint.h>

namespace s256 {

enum {
    MB_MAD_WIDE = 0,    // mad.wide.u32 d64 = a*b + d64 (no carry flag)   -> IMAD.WIDE.U32
    MB_MADC_CHAIN = 1,  // mad.lo.cc/madc.hi.cc chains of 4 + addc         -> IMAD.WIDE.U32(.X) with carry predicates
    MB_MAD_LO = 2,      // mad.lo.u32                                      -> IMAD
    MB_MAD_HI = 3,      // mad.hi.u32                                      -> IMAD.HI.U32
    MB_ADDC_CHAIN = 4,  // add.cc/addc.cc chains of 8                      -> IADD3 / IADD3.X
    MB_MADC_PAIR = 5,   // mad.lo.cc + madc.hi.cc + addc (carry OUT only)
    MB_MIX = 6,         // mad.wide interleaved with add.cc chains (dual-pipe issue)
    MB_DFMA = 7,        // fma.rz.f64, eight independent chains                           -> DFMA (FP64 pipe)
    MB_DFMA_IMAD = 8,   // 8 DFMA + 4 carry-chained wide MADs per repeat (2 : 1, the ratio of the pipe rates): do the two pipes overlap?
    MB_DFMA_PROD = 9,   // one 52x52-bit product the FP64 way: 2 DFMA + 1 DADD + two 64-bit integer adds
    MB_DFMA_PROD_IMAD = 10,  // MB_DFMA_PROD and a carry-chained wide MAD side by side (1 : 1)
    MB_NVARIANTS = 11
};

template <int V>
__global__ void __launch_bounds__(256) k_int_probe(uint32_t seed, int iters, unsigned long long *sink) {
    uint32_t a = seed ^ (threadIdx.x * 2654435761u), b = seed + blockIdx.x * 40503u + 1u;
    uint32_t r0 = a, r1 = b, r2 = a + 1, r3 = b + 2, r4 = a + 3, r5 = b + 4, r6 = a + 5, r7 = b + 6;
    uint32_t s0 = b, s1 = a, s2 = b + 1, s3 = a + 2, s4 = b + 3, s5 = a + 4, s6 = b + 5, s7 = a + 6, t = 0, u = 0;
    unsigned long long c0 = a, c1 = b, c2 = a + 1, c3 = b + 2, c4 = a + 3, c5 = b + 4, c6 = a + 5, c7 = b + 6;
    // FP64 probes: operands below 2^52 held as doubles, the splitting constants of the
    // 52-bit-limb product (hi = fma_rz(a, b, 2^104) - 2^104, lo = fma_rz(a, b, 2^104 + 2^52 - hi') - 2^52)
    double d0 = (double)(a & 0xfffff) + 1.0, d1 = d0 + 3.0, d2 = d0 + 5.0, d3 = d0 + 7.0, d4 = d0 + 11.0, d5 = d0 + 13.0, d6 = d0 + 17.0,
           d7 = d0 + 19.0;
    const double dm = 0.99999904632568359375 + (double)(b & 7) * 1.1102230246251565e-16, dy = 1.0 + (double)(a & 3);
    const double C1 = 20282409603651670423947251286016.0 /* 2^104 */, C2 = C1 + 4503599627370496.0 /* + 2^52 */;
    const double db = 4503599627370495.0 - (double)(b & 0xffff);
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (V == MB_MAD_WIDE) {
                // eight 64-bit accumulators; the multiplicand is the low word of a neighbour
                asm volatile(
                    "mad.wide.u32 %0,%8,%16,%0; mad.wide.u32 %1,%9,%16,%1; mad.wide.u32 %2,%10,%16,%2; mad.wide.u32 %3,%11,%16,%3;"
                    "mad.wide.u32 %4,%12,%16,%4; mad.wide.u32 %5,%13,%16,%5; mad.wide.u32 %6,%14,%16,%6; mad.wide.u32 %7,%15,%16,%7;"
                    : "+l"(c0), "+l"(c1), "+l"(c2), "+l"(c3), "+l"(c4), "+l"(c5), "+l"(c6), "+l"(c7)
                    : "r"((uint32_t)c1), "r"((uint32_t)c2), "r"((uint32_t)c3), "r"((uint32_t)c4), "r"((uint32_t)c5),
                      "r"((uint32_t)c6), "r"((uint32_t)c7), "r"((uint32_t)c0), "r"(b));
            } else if (V == MB_MADC_CHAIN) {  // 8 wide MADs per trip: two chains of four, multiplier from the other chain
                asm volatile(
                    "mad.lo.cc.u32 %0,%9,%13,%0; madc.hi.cc.u32 %1,%9,%13,%1; madc.lo.cc.u32 %2,%10,%13,%2; madc.hi.cc.u32 %3,%10,%13,%3;"
                    "madc.lo.cc.u32 %4,%11,%13,%4; madc.hi.cc.u32 %5,%11,%13,%5; madc.lo.cc.u32 %6,%12,%13,%6; madc.hi.cc.u32 %7,%12,%13,%7;"
                    "addc.u32 %8,%8,0;"
                    : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "+r"(t)
                    : "r"(s0), "r"(s2), "r"(s4), "r"(s6), "r"(s1));
                asm volatile(
                    "mad.lo.cc.u32 %0,%9,%13,%0; madc.hi.cc.u32 %1,%9,%13,%1; madc.lo.cc.u32 %2,%10,%13,%2; madc.hi.cc.u32 %3,%10,%13,%3;"
                    "madc.lo.cc.u32 %4,%11,%13,%4; madc.hi.cc.u32 %5,%11,%13,%5; madc.lo.cc.u32 %6,%12,%13,%6; madc.hi.cc.u32 %7,%12,%13,%7;"
                    "addc.u32 %8,%8,0;"
                    : "+r"(s0), "+r"(s1), "+r"(s2), "+r"(s3), "+r"(s4), "+r"(s5), "+r"(s6), "+r"(s7), "+r"(u)
                    : "r"(r0), "r"(r2), "r"(r4), "r"(r6), "r"(r1));
            } else if (V == MB_MAD_LO) {
                asm volatile(
                    "mad.lo.u32 %0,%1,%8,%0; mad.lo.u32 %1,%2,%8,%1; mad.lo.u32 %2,%3,%8,%2; mad.lo.u32 %3,%4,%8,%3;"
                    "mad.lo.u32 %4,%5,%8,%4; mad.lo.u32 %5,%6,%8,%5; mad.lo.u32 %6,%7,%8,%6; mad.lo.u32 %7,%0,%8,%7;"
                    : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) : "r"(b));
            } else if (V == MB_MAD_HI) {
                asm volatile(
                    "mad.hi.u32 %0,%1,%8,%0; mad.hi.u32 %1,%2,%8,%1; mad.hi.u32 %2,%3,%8,%2; mad.hi.u32 %3,%4,%8,%3;"
                    "mad.hi.u32 %4,%5,%8,%4; mad.hi.u32 %5,%6,%8,%5; mad.hi.u32 %6,%7,%8,%6; mad.hi.u32 %7,%0,%8,%7;"
                    : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) : "r"(b));
            } else if (V == MB_ADDC_CHAIN) {
                asm volatile(
                    "add.cc.u32 %0,%0,%8; addc.cc.u32 %1,%1,%9; addc.cc.u32 %2,%2,%10; addc.cc.u32 %3,%3,%11;"
                    "addc.cc.u32 %4,%4,%12; addc.cc.u32 %5,%5,%13; addc.cc.u32 %6,%6,%14; addc.u32 %7,%7,%15;"
                    : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7)
                    : "r"(s0), "r"(s1), "r"(s2), "r"(s3), "r"(s4), "r"(s5), "r"(s6), "r"(s7));
                asm volatile(
                    "add.cc.u32 %0,%0,%8; addc.cc.u32 %1,%1,%9; addc.cc.u32 %2,%2,%10; addc.cc.u32 %3,%3,%11;"
                    "addc.cc.u32 %4,%4,%12; addc.cc.u32 %5,%5,%13; addc.cc.u32 %6,%6,%14; addc.u32 %7,%7,%15;"
                    : "+r"(s0), "+r"(s1), "+r"(s2), "+r"(s3), "+r"(s4), "+r"(s5), "+r"(s6), "+r"(s7)
                    : "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(r4), "r"(r5), "r"(r6), "r"(r7));
            } else if (V == MB_MADC_PAIR) {  // 4 wide MADs with carry-out only + 4 addc
                asm volatile(
                    "mad.lo.cc.u32 %0,%10,%14,%0; madc.hi.cc.u32 %1,%10,%14,%1; addc.u32 %8,%8,0;"
                    "mad.lo.cc.u32 %2,%11,%14,%2; madc.hi.cc.u32 %3,%11,%14,%3; addc.u32 %9,%9,0;"
                    "mad.lo.cc.u32 %4,%12,%14,%4; madc.hi.cc.u32 %5,%12,%14,%5; addc.u32 %8,%8,0;"
                    "mad.lo.cc.u32 %6,%13,%14,%6; madc.hi.cc.u32 %7,%13,%14,%7; addc.u32 %9,%9,0;"
                    : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "+r"(t), "+r"(u)
                    : "r"(s0), "r"(s2), "r"(s4), "r"(s6), "r"(s1));
                s0 += r1; s2 += r3; s4 += r5; s6 += r7; s1 ^= r0;
            } else if (V == MB_MIX) {  // 4 mad.wide + 8 carry-chained adds
                asm volatile(
                    "{ .reg .u64 d;\n"
                    "mov.b64 d,{%0,%1}; mad.wide.u32 d,%2,%16,d; mov.b64 {%0,%1},d; add.cc.u32 %8,%8,%0; addc.cc.u32 %9,%9,%1;\n"
                    "mov.b64 d,{%2,%3}; mad.wide.u32 d,%4,%16,d; mov.b64 {%2,%3},d; addc.cc.u32 %10,%10,%2; addc.cc.u32 %11,%11,%3;\n"
                    "mov.b64 d,{%4,%5}; mad.wide.u32 d,%6,%16,d; mov.b64 {%4,%5},d; addc.cc.u32 %12,%12,%4; addc.cc.u32 %13,%13,%5;\n"
                    "mov.b64 d,{%6,%7}; mad.wide.u32 d,%0,%16,d; mov.b64 {%6,%7},d; addc.cc.u32 %14,%14,%6; addc.u32 %15,%15,%7; }\n"
                    : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "+r"(s0), "+r"(s1),
                      "+r"(s2), "+r"(s3), "+r"(s4), "+r"(s5), "+r"(s6), "+r"(s7)
                    : "r"(b));
            } else if (V == MB_DFMA || V == MB_DFMA_IMAD) {
                asm volatile(
                    "fma.rz.f64 %0,%0,%8,%9; fma.rz.f64 %1,%1,%8,%9; fma.rz.f64 %2,%2,%8,%9; fma.rz.f64 %3,%3,%8,%9;"
                    "fma.rz.f64 %4,%4,%8,%9; fma.rz.f64 %5,%5,%8,%9; fma.rz.f64 %6,%6,%8,%9; fma.rz.f64 %7,%7,%8,%9;"
                    : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3), "+d"(d4), "+d"(d5), "+d"(d6), "+d"(d7) : "d"(dm), "d"(dy));
                if (V == MB_DFMA_IMAD) {
                    asm volatile(
                        "mad.lo.cc.u32 %0,%9,%13,%0; madc.hi.cc.u32 %1,%9,%13,%1; madc.lo.cc.u32 %2,%10,%13,%2; madc.hi.cc.u32 %3,%10,%13,%3;"
                        "madc.lo.cc.u32 %4,%11,%13,%4; madc.hi.cc.u32 %5,%11,%13,%5; madc.lo.cc.u32 %6,%12,%13,%6; madc.hi.cc.u32 %7,%12,%13,%7;"
                        "addc.u32 %8,%8,0;"
                        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "+r"(t)
                        : "r"(s0), "r"(s2), "r"(s4), "r"(s6), "r"(s1));
                    s0 += r1; s2 += r3; s4 += r5; s6 += r7; s1 ^= r0;
                }
            } else if (V == MB_DFMA_PROD || V == MB_DFMA_PROD_IMAD) {
                // four products per repeat; the next multiplicand is the low half just computed
#define S256_MB_PROD(D, ACC_HI, ACC_LO)                                                                              \
    asm volatile("{ .reg .f64 h, l, sb; .reg .u64 x;\n"                                                              \
                 "fma.rz.f64 h,%0,%3,%4; sub.rz.f64 sb,%5,h; fma.rz.f64 l,%0,%3,sb;\n"                                \
                 "mov.b64 x,h; add.u64 %1,%1,x; mov.b64 x,l; add.u64 %2,%2,x; sub.rz.f64 %0,l,%6; }\n"                \
                 : "+d"(D), "+l"(ACC_HI), "+l"(ACC_LO) : "d"(db), "d"(C1), "d"(C2), "d"(4503599627370496.0));
                S256_MB_PROD(d0, c0, c1) S256_MB_PROD(d1, c2, c3) S256_MB_PROD(d2, c4, c5) S256_MB_PROD(d3, c6, c7)
                if (V == MB_DFMA_PROD_IMAD) {
                    asm volatile(
                        "mad.lo.cc.u32 %0,%5,%7,%0; madc.hi.cc.u32 %1,%5,%7,%1; madc.lo.cc.u32 %2,%6,%7,%2; madc.hi.cc.u32 %3,%6,%7,%3;"
                        "addc.u32 %4,%4,0;"
                        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(t) : "r"(s0), "r"(s2), "r"(s1));
                    s0 += r1; s2 += r3; s1 ^= r0;
                }
            }
        }
    }
    c0 ^= (unsigned long long)__double_as_longlong(d0) ^ __double_as_longlong(d1) ^ __double_as_longlong(d2) ^ __double_as_longlong(d3);
    c1 ^= (unsigned long long)__double_as_longlong(d4) ^ __double_as_longlong(d5) ^ __double_as_longlong(d6) ^ __double_as_longlong(d7);
    unsigned long long cc = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
    uint32_t s = (uint32_t)cc ^ (uint32_t)(cc >> 32) ^ r0 ^ r1 ^ r2 ^ r3 ^ r4 ^ r5 ^ r6 ^ r7 ^ s0 ^ s1 ^ s2 ^ s3 ^ s4 ^ s5 ^ s6 ^ s7 ^ t ^ u;
    if (s == 0x12345678u) sink[0] = s;
}

// operations per loop trip per thread (8 unrolled repeats)
static inline double mb_ops_per_trip(int v) {
    switch (v) {
        case MB_MAD_WIDE: return 8 * 8;
        case MB_MADC_CHAIN: return 8 * 8;
        case MB_MAD_LO: return 8 * 8;
        case MB_MAD_HI: return 8 * 8;
        case MB_ADDC_CHAIN: return 16 * 8;
        case MB_MADC_PAIR: return 4 * 8;
        case MB_MIX: return 4 * 8;
        case MB_DFMA: return 8 * 8;           // DFMAs
        case MB_DFMA_IMAD: return 12 * 8;     // 8 DFMAs + 4 wide MADs
        case MB_DFMA_PROD: return 4 * 8;      // 52x52-bit products
        case MB_DFMA_PROD_IMAD: return 6 * 8; // 4 FP64 products + 2 wide MADs
        default: return 0;
    }
}

}  // namespace s256
