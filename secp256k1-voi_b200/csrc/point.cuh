// point.cuh -- secp256k1 group law in homogeneous projective coordinates.
//
// The Renes-Costello-Batina complete formulas for a = 0, b3 = 21, kept
// branch-free exactly as the reference uses them (point_projective.go:24-120
// Algorithm 7, :123-205 Algorithm 8, :208-273 Algorithm 9); identity is
// (0 : 1 : 0) (point.go:42-49).  Multiplications by b3 go through
// fe_mul_small (8 MAC32 instead of 73).  Only affine values are observable
// (point_test.go:359-390), so temporaries / scheduling differ freely.
// Each formula is a template over the field-operation set: <false> (default) is the constant-time
// one, <true> the variable-time one of fe_vt.cuh for the paths that are variable time in the
// reference as well (verification ladder, vartime MSM).
#pragma once
#include "fe.cuh"
#include "fe_vt.cuh"

namespace s256 {

struct pt {
    fe x, y, z;
};
struct apt {
    fe x, y;
};

#define S256_B3 21u

S256_HD void pt_set_identity(pt &r) {
    r.x = fe_zero();
    r.y = fe_one();
    r.z = fe_zero();
}
S256_HD uint32_t pt_is_identity(const pt &p) { return fe_is_zero(p.z); }  // point.go:148-152
S256_HD void pt_from_affine(pt &r, const apt &a) {
    r.x = a.x;
    r.y = a.y;
    r.z = fe_one();
}
S256_HD void pt_cmov(pt &r, const pt &a, const pt &b, uint32_t ctrl) {
    fe_cmov(r.x, a.x, b.x, ctrl);
    fe_cmov(r.y, a.y, b.y, ctrl);
    fe_cmov(r.z, a.z, b.z, ctrl);
}
// y -> -y iff ctrl (point.go:100-106 ConditionalNegate)
S256_HD void fe_cneg(fe &r, const fe &a, uint32_t ctrl) {
    fe n;
    fe_neg(n, a);
    fe_cmov(r, a, n, ctrl);
}

// v = p + q, complete (12 M + 2 m3b + 19 a).
template <bool VT = false>
S256_HD void pt_add(pt &v, const pt &p, const pt &q) {
    typedef fe_ops<VT> F;
    fe t0, t1, t2, t3, t4, x3, y3, z3;
    F::mul(t0, p.x, q.x);
    F::mul(t1, p.y, q.y);
    F::mul(t2, p.z, q.z);
    F::add(t3, p.x, p.y);
    F::add(t4, q.x, q.y);
    F::mul(t3, t3, t4);
    F::add(t4, t0, t1);
    F::sub(t3, t3, t4);
    F::add(t4, p.y, p.z);
    F::add(x3, q.y, q.z);
    F::mul(t4, t4, x3);
    F::add(x3, t1, t2);
    F::sub(t4, t4, x3);
    F::add(x3, p.x, p.z);
    F::add(y3, q.x, q.z);
    F::mul(x3, x3, y3);
    F::add(y3, t0, t2);
    F::sub(y3, x3, y3);
    F::add(x3, t0, t0);
    F::add(t0, x3, t0);
    F::mul_small(t2, t2, S256_B3);
    F::add(z3, t1, t2);
    F::sub(t1, t1, t2);
    F::mul_small(y3, y3, S256_B3);
    // X3 = t3 t1 - t4 y3, Y3 = t1 z3 + y3 t0, Z3 = z3 t4 + t0 t3: three sums of two products
    F::mul2sub(v.x, t3, t1, t4, y3);
    F::mul2add(v.y, t1, z3, y3, t0);
    F::mul2add(v.z, z3, t4, t0, t3);
}

// v = p + (x2, y2, 1); complete for every p, addend must not be the identity
// (11 M + 2 m3b + 13 a).
template <bool VT = false>
S256_HD void pt_add_mixed(pt &v, const pt &p, const fe &x2, const fe &y2) {
    typedef fe_ops<VT> F;
    fe t0, t1, t2, t3, t4, x3, y3, z3;
    F::mul(t0, p.x, x2);
    F::mul(t1, p.y, y2);
    F::add(t3, x2, y2);
    F::add(t4, p.x, p.y);
    F::mul(t3, t3, t4);
    F::add(t4, t0, t1);
    F::sub(t3, t3, t4);
    F::mul(t4, y2, p.z);
    F::add(t4, t4, p.y);
    F::mul(y3, x2, p.z);
    F::add(y3, y3, p.x);
    F::add(x3, t0, t0);
    F::add(t0, x3, t0);
    F::mul_small(t2, p.z, S256_B3);
    F::add(z3, t1, t2);
    F::sub(t1, t1, t2);
    F::mul_small(y3, y3, S256_B3);
    F::mul2sub(v.x, t3, t1, t4, y3);
    F::mul2add(v.y, t1, z3, y3, t0);
    F::mul2add(v.z, z3, t4, t0, t3);
}

// v = 2p, complete (6 M + 2 S + 1 m3b + 9 a).
template <bool VT = false>
S256_HD void pt_double(pt &v, const pt &p) {
    typedef fe_ops<VT> F;
    fe t0, t1, t2, x3, y3, z3;
    F::sqr(t0, p.y);
    F::mul8(z3, t0);
    F::mul(t1, p.y, p.z);
    F::sqr(t2, p.z);
    F::mul_small(t2, t2, S256_B3);
    fe z8 = z3, t23;
    F::add(y3, t0, t2);
    F::mul(z3, t1, z8);
    F::add(t1, t2, t2);
    F::add(t23, t1, t2);
    F::sub(t0, t0, t23);
    F::mul(t1, p.x, p.y);          // (before v.y is written: v may alias p)
    F::mul2add(y3, t2, z8, t0, y3);  // Y3 = t2 (8 Y^2) + (Y^2 - 3 t2)(Y^2 + t2)
    F::mul(x3, t0, t1);
    F::add(x3, x3, x3);
    v.x = x3;
    v.y = y3;
    v.z = z3;
}

// y^2 == x^3 + 7 (point_s11n.go:298-307)
S256_HD void fe_curve_rhs(fe &yy, const fe &x) {
    fe t;
    fe_sqr(t, x);
    fe_mul(t, t, x);
    fe seven = fe_from_u32(7);
    fe_add(yy, t, seven);
}
S256_HD uint32_t apt_on_curve(const apt &a) {
    fe yy, y2;
    fe_curve_rhs(yy, a.x);
    fe_sqr(y2, a.y);
    return fe_equal(yy, y2);
}

// beta (point_mul_glv.go:44): lambda * (x, y) = (beta * x, y)
S256_HD fe fe_beta() {
    fe b;
    b.v[0] = 0x719501EEu; b.v[1] = 0xC1396C28u; b.v[2] = 0x12F58995u; b.v[3] = 0x9CF04975u;
    b.v[4] = 0xAC3434E9u; b.v[5] = 0x6E64479Eu; b.v[6] = 0x657C0710u; b.v[7] = 0x7AE96A2Bu;
    return b;
}
// the generator (point.go:18-21)
S256_HD apt apt_generator() {
    apt g;
    g.x.v[0] = 0x16F81798u; g.x.v[1] = 0x59F2815Bu; g.x.v[2] = 0x2DCE28D9u; g.x.v[3] = 0x029BFCDBu;
    g.x.v[4] = 0xCE870B07u; g.x.v[5] = 0x55A06295u; g.x.v[6] = 0xF9DCBBACu; g.x.v[7] = 0x79BE667Eu;
    g.y.v[0] = 0xFB10D4B8u; g.y.v[1] = 0x9C47D08Fu; g.y.v[2] = 0xA6855419u; g.y.v[3] = 0xFD17B448u;
    g.y.v[4] = 0x0E1108A8u; g.y.v[5] = 0x5DA4FBFCu; g.y.v[6] = 0x26A3C465u; g.y.v[7] = 0x483ADA77u;
    return g;
}
// n as a field element (point_s11n.go feN, used by RecoverPoint :264)
S256_HD fe fe_group_order() {
    fe n;
    n.v[0] = 0xD0364141u; n.v[1] = 0xBFD25E8Cu; n.v[2] = 0xAF48A03Bu; n.v[3] = 0xBAAEDCE6u;
    n.v[4] = 0xFFFFFFFEu; n.v[5] = 0xFFFFFFFFu; n.v[6] = 0xFFFFFFFFu; n.v[7] = 0xFFFFFFFFu;
    return n;
}

}  // namespace s256
