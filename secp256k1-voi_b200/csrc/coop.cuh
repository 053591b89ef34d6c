// coop.cuh -- group law for the serial tails (Horner over the MSM windows): one point operation
// spread over the 8 lanes of a lane group.
//
// A lone thread running the complete formulas (point_projective.go:24-273) is bound by the latency of
// its multiplications, 12 / 8 of them back to back.  In both formulas the multiplications fall into
// two rounds of mutually independent products (6 + 6 for the addition, 4 + 4 for the doubling), so
// each lane computes one product per round and the results are exchanged with shuffles: two
// multiplication latencies per operation instead of 12 / 8.  Every lane of a group holds the same
// point before and after the call.  Device only; results are the same group elements as pt_add /
// pt_double (point.cuh), which the host-simulated tests exercise instead.
#pragma once
#include "point.cuh"

#if defined(__CUDACC__)
namespace s256 {

// the value lane `src` of each 8-lane group holds
__device__ __forceinline__ fe fe_group_bcast(const fe &m, int src) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, m.v[i], src, 8);
    return r;
}
__device__ __forceinline__ fe fe_pick(bool c, const fe &a, const fe &b) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = c ? a.v[i] : b.v[i];
    return r;
}

// p = 2p; j = lane index inside the group (0..7; lanes 4..7 repeat 0..3).  Whole warp must call.
__device__ __noinline__ void pt_double_coop(pt &p, int j) {
    int jj = j & 3;
    // round 1: Y^2, Y Z, Z^2, X Y
    fe a = fe_pick(jj == 3, p.x, fe_pick(jj == 2, p.z, p.y));
    fe b = fe_pick(jj == 0 || jj == 3, p.y, p.z);
    fe m;
    fe_mul(m, a, b);
    fe yy = fe_group_bcast(m, 0), yz = fe_group_bcast(m, 1), zz = fe_group_bcast(m, 2), xy = fe_group_bcast(m, 3);
    fe z8, t2, y3, t0, t;
    fe_add(z8, yy, yy);
    fe_add(z8, z8, z8);
    fe_add(z8, z8, z8);
    fe_mul_small(t2, zz, S256_B3);
    fe_add(y3, yy, t2);
    fe_add(t, t2, t2);
    fe_add(t, t, t2);
    fe_sub(t0, yy, t);
    // round 2: t2 z8, (Y Z) z8, t0 y3, t0 (X Y)
    a = fe_pick(jj == 0, t2, fe_pick(jj == 1, yz, t0));
    b = fe_pick(jj < 2, z8, fe_pick(jj == 2, y3, xy));
    fe_mul(m, a, b);
    fe x3 = fe_group_bcast(m, 0), c = fe_group_bcast(m, 2), d = fe_group_bcast(m, 3);
    p.z = fe_group_bcast(m, 1);
    fe_add(p.y, x3, c);
    fe_add(p.x, d, d);
}

// p = p + q (complete).  Lanes 6 and 7 repeat lane 0.  Whole warp must call.
__device__ __noinline__ void pt_add_coop(pt &p, const pt &q, int j) {
    int jj = j < 6 ? j : 0;
    fe s1, s2, a, b, m;
    // round 1: X1X2, Y1Y2, Z1Z2, (X1+Y1)(X2+Y2), (Y1+Z1)(Y2+Z2), (X1+Z1)(X2+Z2)
    fe_add(s1, p.x, p.y);
    fe_add(s2, p.y, p.z);
    fe_add(a, p.x, p.z);
    a = fe_pick(jj == 0, p.x, fe_pick(jj == 1, p.y, fe_pick(jj == 2, p.z, fe_pick(jj == 3, s1, fe_pick(jj == 4, s2, a)))));
    fe_add(s1, q.x, q.y);
    fe_add(s2, q.y, q.z);
    fe_add(b, q.x, q.z);
    b = fe_pick(jj == 0, q.x, fe_pick(jj == 1, q.y, fe_pick(jj == 2, q.z, fe_pick(jj == 3, s1, fe_pick(jj == 4, s2, b)))));
    fe_mul(m, a, b);
    fe t0 = fe_group_bcast(m, 0), t1 = fe_group_bcast(m, 1), t2 = fe_group_bcast(m, 2);
    fe t3 = fe_group_bcast(m, 3), t4 = fe_group_bcast(m, 4), y3 = fe_group_bcast(m, 5);
    fe t, z3;
    fe_add(t, t0, t1);
    fe_sub(t3, t3, t);
    fe_add(t, t1, t2);
    fe_sub(t4, t4, t);
    fe_add(t, t0, t2);
    fe_sub(y3, y3, t);
    fe_add(t, t0, t0);
    fe_add(t0, t, t0);
    fe_mul_small(t2, t2, S256_B3);
    fe_add(z3, t1, t2);
    fe_sub(t1, t1, t2);
    fe_mul_small(y3, y3, S256_B3);
    // round 2: t4 y3, t3 t1, y3 t0, t1 z3, t0 t3, z3 t4
    a = fe_pick(jj == 0, t4, fe_pick(jj == 1, t3, fe_pick(jj == 2, y3, fe_pick(jj == 3, t1, fe_pick(jj == 4, t0, z3)))));
    b = fe_pick(jj == 0, y3, fe_pick(jj == 1, t1, fe_pick(jj == 2, t0, fe_pick(jj == 3, z3, fe_pick(jj == 4, t3, t4)))));
    fe_mul(m, a, b);
    fe m0 = fe_group_bcast(m, 0), m1 = fe_group_bcast(m, 1), m2 = fe_group_bcast(m, 2);
    fe m3 = fe_group_bcast(m, 3), m4 = fe_group_bcast(m, 4), m5 = fe_group_bcast(m, 5);
    fe_sub(p.x, m1, m0);
    fe_add(p.y, m3, m2);
    fe_add(p.z, m5, m4);
}

}  // namespace s256
#endif
