// modinv.cuh -- modular inversion by the Bernstein-Yang "safegcd" divstep iteration, constant time.
//
// Replaces the reference's Fermat chains (internal/field/field_invert.go:11 x^(p-2), 255 S + 15 M;
// scalar_invert.go:11 x^(n-2), 253 S + 40 M) wherever an inverse is needed: the value is the same
// unique inverse (and Invert(0) = 0, as in the reference), the cost is ~4x lower -- 600 divsteps in
// 20 batches of 30, each batch a 2x2 integer matrix applied to (f, g) and to (d, e) mod m -- and the
// dependent chain is much shorter, which is what the latency-bound batched-inversion kernels feel.
// The algorithm and its bound (590 divsteps suffice for 256-bit inputs; 600 are run) are those of
// D. J. Bernstein and B.-Y. Yang, "Fast constant-time gcd computation and modular inversion"
// (TCHES 2019/3), in the signed-30-bit-limb form popularised by libsecp256k1's modinv32; written
// here from the published description.  No branch and no memory address depends on the data.
#pragma once
#include <stdint.h>

#include "fe.cuh"

namespace s256 {

struct mi_s30 {
    int32_t v[9];  // value = sum v[i] * 2^(30 i); limbs 0..7 in (-2^30, 2^30), limb 8 small, signed
};
struct mi_modulus {
    mi_s30 m;
    uint32_t inv30;  // m^-1 mod 2^30
};
#define S256_M30 0x3FFFFFFF

S256_HD void mi_from_limbs(mi_s30 &r, const uint32_t a[8]) {
    r.v[0] = (int32_t)(a[0] & S256_M30);
    r.v[1] = (int32_t)(((a[0] >> 30) | (a[1] << 2)) & S256_M30);
    r.v[2] = (int32_t)(((a[1] >> 28) | (a[2] << 4)) & S256_M30);
    r.v[3] = (int32_t)(((a[2] >> 26) | (a[3] << 6)) & S256_M30);
    r.v[4] = (int32_t)(((a[3] >> 24) | (a[4] << 8)) & S256_M30);
    r.v[5] = (int32_t)(((a[4] >> 22) | (a[5] << 10)) & S256_M30);
    r.v[6] = (int32_t)(((a[5] >> 20) | (a[6] << 12)) & S256_M30);
    r.v[7] = (int32_t)(((a[6] >> 18) | (a[7] << 14)) & S256_M30);
    r.v[8] = (int32_t)(a[7] >> 16);
}
// caller passes a normalised value: every limb in [0, 2^30), limb 8 in [0, 2^16)
S256_HD void mi_to_limbs(uint32_t a[8], const mi_s30 &s) {
    const uint32_t *v = reinterpret_cast<const uint32_t *>(s.v);
    a[0] = v[0] | (v[1] << 30);
    a[1] = (v[1] >> 2) | (v[2] << 28);
    a[2] = (v[2] >> 4) | (v[3] << 26);
    a[3] = (v[3] >> 6) | (v[4] << 24);
    a[4] = (v[4] >> 8) | (v[5] << 22);
    a[5] = (v[5] >> 10) | (v[6] << 20);
    a[6] = (v[6] >> 12) | (v[7] << 18);
    a[7] = (v[7] >> 14) | (v[8] << 16);
}

struct mi_trans {
    int32_t u, v, q, r;
};

// 30 divsteps on the low bits of (f, g); returns the new zeta = -(delta + 1/2) and the transition matrix
// t with  2^30 * (f', g') = t * (f, g).
S256_HD int32_t mi_divsteps_30(int32_t zeta, uint32_t f0, uint32_t g0, mi_trans &t) {
    uint32_t u = 1, v = 0, q = 0, r = 1, f = f0, g = g0;
#pragma unroll 6
    for (int i = 0; i < 30; i++) {
        uint32_t mask1 = (uint32_t)(zeta >> 31);  // zeta < 0
        uint32_t mask2 = 0u - (g & 1u);           // g odd
        uint32_t x = (f ^ mask1) - mask1, y = (u ^ mask1) - mask1, z = (v ^ mask1) - mask1;
        g += x & mask2;
        q += y & mask2;
        r += z & mask2;
        mask1 &= mask2;
        zeta = (int32_t)((uint32_t)zeta ^ mask1) - 1;  // -zeta - 2 if both, else zeta - 1
        f += g & mask1;
        u += q & mask1;
        v += r & mask1;
        g >>= 1;
        u <<= 1;
        v <<= 1;
    }
    t.u = (int32_t)u;
    t.v = (int32_t)v;
    t.q = (int32_t)q;
    t.r = (int32_t)r;
    return zeta;
}

// (d, e) <- t * (d, e) / 2^30 mod m, limbs kept in (-2^30, 2^30) with d, e in (-2m, m)
S256_HD void mi_update_de(mi_s30 &d, mi_s30 &e, const mi_trans &t, const mi_modulus &mod) {
    const int32_t u = t.u, v = t.v, q = t.q, r = t.r;
    int32_t sd = d.v[8] >> 31, se = e.v[8] >> 31;
    int32_t md = (u & sd) + (v & se), me = (q & sd) + (r & se);
    int32_t di = d.v[0], ei = e.v[0];
    int64_t cd = (int64_t)u * di + (int64_t)v * ei;
    int64_t ce = (int64_t)q * di + (int64_t)r * ei;
    // multiples of m that clear the bottom 30 bits
    md -= (int32_t)((mod.inv30 * (uint32_t)cd + (uint32_t)md) & S256_M30);
    me -= (int32_t)((mod.inv30 * (uint32_t)ce + (uint32_t)me) & S256_M30);
    cd += (int64_t)mod.m.v[0] * md;
    ce += (int64_t)mod.m.v[0] * me;
    cd >>= 30;
    ce >>= 30;
#pragma unroll
    for (int i = 1; i < 9; i++) {
        di = d.v[i];
        ei = e.v[i];
        cd += (int64_t)u * di + (int64_t)v * ei;
        ce += (int64_t)q * di + (int64_t)r * ei;
        cd += (int64_t)mod.m.v[i] * md;
        ce += (int64_t)mod.m.v[i] * me;
        d.v[i - 1] = (int32_t)cd & S256_M30;
        cd >>= 30;
        e.v[i - 1] = (int32_t)ce & S256_M30;
        ce >>= 30;
    }
    d.v[8] = (int32_t)cd;
    e.v[8] = (int32_t)ce;
}
// (f, g) <- t * (f, g) / 2^30 (exact)
S256_HD void mi_update_fg(mi_s30 &f, mi_s30 &g, const mi_trans &t) {
    const int32_t u = t.u, v = t.v, q = t.q, r = t.r;
    int32_t fi = f.v[0], gi = g.v[0];
    int64_t cf = (int64_t)u * fi + (int64_t)v * gi;
    int64_t cg = (int64_t)q * fi + (int64_t)r * gi;
    cf >>= 30;
    cg >>= 30;
#pragma unroll
    for (int i = 1; i < 9; i++) {
        fi = f.v[i];
        gi = g.v[i];
        cf += (int64_t)u * fi + (int64_t)v * gi;
        cg += (int64_t)q * fi + (int64_t)r * gi;
        f.v[i - 1] = (int32_t)cf & S256_M30;
        cf >>= 30;
        g.v[i - 1] = (int32_t)cg & S256_M30;
        cg >>= 30;
    }
    f.v[8] = (int32_t)cf;
    g.v[8] = (int32_t)cg;
}
// r in (-2m, m), negated if sign < 0, brought to [0, m)
S256_HD void mi_normalize(mi_s30 &r, int32_t sign, const mi_modulus &mod) {
    int32_t cond_add = r.v[8] >> 31;
#pragma unroll
    for (int i = 0; i < 9; i++) r.v[i] += mod.m.v[i] & cond_add;
    int32_t cond_negate = sign >> 31;
#pragma unroll
    for (int i = 0; i < 9; i++) r.v[i] = (r.v[i] ^ cond_negate) - cond_negate;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.v[i + 1] += r.v[i] >> 30;
        r.v[i] &= S256_M30;
    }
    cond_add = r.v[8] >> 31;
#pragma unroll
    for (int i = 0; i < 9; i++) r.v[i] += mod.m.v[i] & cond_add;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.v[i + 1] += r.v[i] >> 30;
        r.v[i] &= S256_M30;
    }
}

// out = in^-1 mod m (0 for 0); in must be < m
S256_HD void mi_invert(uint32_t out[8], const uint32_t in[8], const mi_modulus &mod) {
    mi_s30 d, e, f = mod.m, g;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        d.v[i] = 0;
        e.v[i] = 0;
    }
    e.v[0] = 1;
    mi_from_limbs(g, in);
    int32_t zeta = -1;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int i = 0; i < 20; i++) {  // 600 divsteps
        mi_trans t;
        zeta = mi_divsteps_30(zeta, (uint32_t)f.v[0], (uint32_t)g.v[0], t);
        mi_update_de(d, e, t, mod);
        mi_update_fg(f, g, t);
    }
    // g = 0 and f = +-gcd = +-1 (or +-m for in = 0, where d = 0)
    mi_normalize(d, f.v[8], mod);
    mi_to_limbs(out, d);
}

S256_HD mi_modulus mi_modulus_p() {
    mi_modulus m;
    m.m.v[0] = 0x3FFFFC2F; m.m.v[1] = 0x3FFFFFFB; m.m.v[2] = 0x3FFFFFFF; m.m.v[3] = 0x3FFFFFFF; m.m.v[4] = 0x3FFFFFFF;
    m.m.v[5] = 0x3FFFFFFF; m.m.v[6] = 0x3FFFFFFF; m.m.v[7] = 0x3FFFFFFF; m.m.v[8] = 0xFFFF;
    m.inv30 = 0x2DDACACFu;
    return m;
}
S256_HD mi_modulus mi_modulus_n() {
    mi_modulus m;
    m.m.v[0] = 0x10364141; m.m.v[1] = 0x3F497A33; m.m.v[2] = 0x348A03BB; m.m.v[3] = 0x2BB739AB; m.m.v[4] = 0x3FFFFEBA;
    m.m.v[5] = 0x3FFFFFFF; m.m.v[6] = 0x3FFFFFFF; m.m.v[7] = 0x3FFFFFFF; m.m.v[8] = 0xFFFF;
    m.inv30 = 0x2A774EC1u;
    return m;
}

}  // namespace s256
