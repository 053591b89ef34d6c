// fe52.cuh -- EXPERIMENT, not on any product path: F_p multiplication on the FP64 pipe.
//
// Why: the B200 issues DFMA at 58.7 /clk/SM but IMAD.WIDE.U32 at 28.8 (profiles/
// r01_microbench_fp64_pipe.json), and one DFMA pair yields a 52x52 -> 104-bit product
// (2704 bit^2) where one IMAD.WIDE yields 32x32 (1024 bit^2).  The two do NOT overlap
// (same probe: 8 DFMA + 4 IMAD.WIDE take the sum of their times), so the question is
// only which pipe form multiplies 256-bit numbers in fewer issue cycles.
//
// Representation: five integer-valued doubles v[i] in [0, 2^52), value = sum v[i] 2^(52 i),
// loosely reduced: v[4] <= 2^48 + 2^6, i.e. value < 2^256 + 2^214 (not canonical).
//
// Product of two limbs (both < 2^52), exact, three FP64 operations:
//     h = fma_rz(a, b, 2^104)            in [2^104, 2^105): mantissa = floor(a b / 2^52)
//     l = fma_rz(a, b, 2^104 + 2^52 - h) in [2^52, 2^53):   mantissa = a b mod 2^52
// The bit patterns of h and l are added into 64-bit integer column sums (ALU pipe); the
// exponent fields add up to per-column constants that the accumulators start from, negated.
// Reduction: 2^260 = 16 (2^32 + 977) =: R (mod p), a 37-bit constant: the five high columns
// are carried to 52 bits, turned into doubles and multiplied by R the same way.
#pragma once
#include <stdint.h>
#include <string.h>
#if !defined(__CUDA_ARCH__)
#include <math.h>
#endif

namespace s256 {

#if defined(__CUDACC__)
#define FE52_HD __host__ __device__ __forceinline__
#else
#define FE52_HD static inline
#endif

struct fe52 {
    double v[5];
};

// host callers set fesetround(FE_TOWARDZERO) around these (tests/test_fe52_host.py does)
FE52_HD double fe52_fma_rz(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rz(a, b, c);
#else
    return fma(a, b, c);
#endif
}
FE52_HD double fe52_sub_rz(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rz(a, b);
#else
    return a - b;
#endif
}
FE52_HD uint64_t fe52_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    memcpy(&u, &d, 8);
    return u;
#endif
}
// integer below 2^52 -> double: exponent of 2^52 OR-ed in, then one subtraction
FE52_HD double fe52_from_u52(uint64_t x) {
    uint64_t u = x | 0x4330000000000000ull;
#if defined(__CUDA_ARCH__)
    return __dsub_rz(__longlong_as_double((long long)u), 4503599627370496.0);
#else
    double d;
    memcpy(&d, &u, 8);
    return d - 4503599627370496.0;
#endif
}

#define FE52_EXP_L (0x433ull << 52) /* exponent field of [2^52, 2^53)   */
#define FE52_EXP_H (0x467ull << 52) /* exponent field of [2^104, 2^105) */
#define FE52_M52 0xFFFFFFFFFFFFFull

FE52_HD void fe52_mul(fe52 &r, const fe52 &a, const fe52 &b) {
    const double C1 = 20282409603651670423947251286016.0;  // 2^104
    const double C2 = C1 + 4503599627370496.0;             // 2^104 + 2^52
    // columns 0..9 start from minus the exponent fields they are going to collect:
    // column k takes nl(k) low halves and nl(k-1) high halves, nl = 1,2,3,4,5,4,3,2,1
    uint64_t c[10];
#pragma unroll
    for (int k = 0; k < 10; k++) {
        int nl = k <= 4 ? k + 1 : (k <= 8 ? 9 - k : 0);
        int nh = k == 0 ? 0 : (k - 1 <= 4 ? k : 10 - k);
        c[k] = 0 - ((uint64_t)nl * FE52_EXP_L + (uint64_t)nh * FE52_EXP_H);
    }
#pragma unroll
    for (int i = 0; i < 5; i++) {
#pragma unroll
        for (int j = 0; j < 5; j++) {
            double h = fe52_fma_rz(a.v[i], b.v[j], C1);
            double l = fe52_fma_rz(a.v[i], b.v[j], fe52_sub_rz(C2, h));
            c[i + j] += fe52_bits(l);
            c[i + j + 1] += fe52_bits(h);
        }
    }
    // every column is now below 10 * 2^52.  Carry columns 4..9 so that 5..9 fit 52 bits
    // (column 9 is below 2^46 for loosely reduced inputs and takes the last carry as is).
#pragma unroll
    for (int k = 4; k < 9; k++) {
        c[k + 1] += c[k] >> 52;
        c[k] &= FE52_M52;
    }
    // fold: column k+5 times R into columns k (low half) and k+1 (high half, 37 bits)
    const double R = 68719492368.0;  // 2^36 + 15632 = 2^260 mod p
    uint64_t top = 0 - FE52_EXP_H;   // the high half of the last fold product: column 5 again
#pragma unroll
    for (int k = 0; k < 5; k++) {
        double x = fe52_from_u52(c[k + 5]);
        double h = fe52_fma_rz(x, R, C1);
        double l = fe52_fma_rz(x, R, fe52_sub_rz(C2, h));
        c[k] += fe52_bits(l) - FE52_EXP_L;
        if (k < 4) c[k + 1] += fe52_bits(h) - FE52_EXP_H;
        else top += fe52_bits(h);
    }
    // top < 2^37 stands at 2^260 again: top * R = (top >> 16) 2^52 + ((top & 0xffff) << 36) + top * 15632
    c[0] += ((top & 0xFFFFull) << 36) + top * 15632ull;
    c[1] += top >> 16;
    // bits 48.. of column 4 stand at 2^256 = 2^32 + 977 (mod p); the carry that reaches column 4
    // afterwards is below 2^6, which the loose bound allows
    uint64_t t = c[4] >> 48;
    c[4] &= 0xFFFFFFFFFFFFull;
    c[0] += t * 4294968273ull;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        c[k + 1] += c[k] >> 52;
        c[k] &= FE52_M52;
    }
#pragma unroll
    for (int k = 0; k < 5; k++) r.v[k] = fe52_from_u52(c[k]);
}

}  // namespace s256
