// jac.cuh -- Jacobian-coordinate group law (x = X / Z^2, y = Y / Z^3) for the VARIABLE-TIME ladders only.
//
// The reference runs every path on the Renes-Costello-Batina complete formulas (point_projective.go:24-273), and so do
// the constant-time kernels and the Point API here (point.cuh).  The verification ladder
// (DoubleScalarMultBasepointVartime, point_mul_glv.go:307-317) handles public data, is variable time in the reference
// too, and only its affine result is observable -- so its 125 doublings may use the a = 0 Jacobian doubling
// (2 M + 5 S = 371 MAC32 against 519 for Algorithm 9) and its additions the mixed Jacobian addition against an AFFINE
// per-item table (6 M + 3 S + one fused pair = 710 against 849 for Algorithm 7).  These formulas are incomplete; the
// exceptional cases the complete formulas absorb are handled explicitly, by branches (public data):
//   * accumulator at infinity          -> a flag kept by the caller; the addition becomes an assignment
//   * addend == accumulator (H = R = 0) -> the doubling
//   * addend == -accumulator (H = 0)    -> the flag is set
//   * doubling of a point with Y = 0    -> cannot occur (no 2-torsion on a prime-order curve)
// SURVEY.md section 6 names the Wycheproof groups that exercise them (PointDuplication, EdgeCaseShamirMultiplication,
// ArithmeticError); tests/parity_suites.py runs them plus crafted u1 G = +-u2 P rows against the oracle.
#pragma once
#include "point.cuh"

namespace s256 {

// v = 2p (a = 0).  p must not be at infinity for the result to mean anything; (0 : 1 : 0) maps to Z = 0 harmlessly.
template <bool VT = true>
S256_HD void jac_double(pt &v, const pt &p) {
    typedef fe_ops<VT> F;
    fe a, b, c, d, e, f, t, z3;
    F::sqr(a, p.x);
    F::sqr(b, p.y);
    F::sqr(c, b);
#ifndef S256_JDBL_3M4S
    F::add(t, p.x, b);  // D = 2 ((X + B)^2 - A - C)
    F::sqr(t, t);
    F::sub(t, t, a);
    F::sub(t, t, c);
    F::mul2(d, t);
#else
    F::mul(t, p.x, b);  // D = 4 X B
    F::mul2(t, t);
    F::mul2(d, t);
#endif
    F::mul3(e, a);  // E = 3 A
    F::sqr(f, e);
    F::mul(z3, p.y, p.z);
    F::mul2(v.z, z3);
    F::sub2(t, f, d);  // X3 = F - 2 D
    v.x = t;
    F::sub(t, d, t);
    F::mul(f, e, t);  // Y3 = E (D - X3) - 8 C
    F::submul8(v.y, f, c);
}
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ pt jac_double_call(pt a) {
    pt r;
    jac_double<true>(r, a);
    return r;
}
#else
static inline pt jac_double_call(pt a) {
    pt r;
    jac_double<true>(r, a);
    return r;
}
#endif

// v = p + (x2, y2) for a finite p whose sum with the addend is known not to be exceptional (table construction:
// k P + P with 2 <= k < 16 on a curve of prime order).
template <bool VT = true>
S256_HD void jac_add_mixed_nocheck(pt &v, const pt &p, const fe &x2, const fe &y2) {
    typedef fe_ops<VT> F;
    fe zz, u2, s2, h, r, hh, hhh, w, t, x3;
    F::sqr(zz, p.z);
    F::mul(u2, x2, zz);
    F::mul(s2, y2, p.z);
    F::mul(s2, s2, zz);
    F::sub(h, u2, p.x);
    F::sub(r, s2, p.y);
    F::sqr(hh, h);
    F::mul(hhh, h, hh);
    F::mul(w, p.x, hh);
    F::mul(v.z, p.z, h);
    F::sqr(x3, r);
    F::sub(x3, x3, hhh);
    F::sub2(x3, x3, w);
    F::sub(t, w, x3);
    F::mul2sub(v.y, r, t, p.y, hhh);
    v.x = x3;
}

// acc += (x2, y2), every case handled; inf is the caller's "accumulator is the identity" flag.
template <bool VT = true>
S256_HD void jac_add_mixed_var(pt &acc, uint32_t &inf, const fe &x2, const fe &y2) {
    typedef fe_ops<VT> F;
    if (inf) {
        acc.x = x2;
        acc.y = y2;
        acc.z = fe_one();
        inf = 0u;
        return;
    }
    fe zz, u2, s2, h, r, hh, hhh, w, t, x3;
    F::sqr(zz, acc.z);
    F::mul(u2, x2, zz);
    F::mul(s2, y2, acc.z);
    F::mul(s2, s2, zz);
    F::sub(h, u2, acc.x);
    F::sub(r, s2, acc.y);
    if (fe_is_zero(h)) {
        if (fe_is_zero(r))
            acc = jac_double_call(acc);
        else
            inf = 1u;
        return;
    }
    F::sqr(hh, h);
    F::mul(hhh, h, hh);
    F::mul(w, acc.x, hh);
    F::mul(acc.z, acc.z, h);
    F::sqr(x3, r);
    F::sub(x3, x3, hhh);
    F::sub2(x3, x3, w);
    F::sub(t, w, x3);
    F::mul2sub(acc.y, r, t, acc.y, hhh);
    acc.x = x3;
}

// Jacobian -> the homogeneous projective form every consumer of the ladder's result expects: (X Z : Y : Z^3)
template <bool VT = true>
S256_HD void jac_to_projective(pt &v, const pt &p, uint32_t inf) {
    typedef fe_ops<VT> F;
    if (inf) {
        pt_set_identity(v);
        return;
    }
    fe zz;
    F::sqr(zz, p.z);
    F::mul(v.x, p.x, p.z);
    v.y = p.y;
    F::mul(v.z, zz, p.z);
}

}  // namespace s256
