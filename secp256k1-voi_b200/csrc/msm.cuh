// msm.cuh -- Pippenger multi-scalar multiplication  sum_i k_i * P_i.
//
// Replaces the reference's Straus routine (point_mul_multi.go:25-117), whose
// l x 15 x 104 B tables do not scale to l = 2^20 (1.6 GB) and which costs
// ~1000 modmuls per point; only the resulting point is observable, so the
// algorithm is free.  Variable time in the scalars (MultiScalarMultVartime,
// :73); the constant-time flavour (:25) is served by the ct ladder per item
// plus a sum (api.cu).
//
//   digits : k -> signed c-bit digits d_w in [-(2^(c-1)-1), 2^(c-1)], nwin windows
//   sort   : counting sort of (window, |d|) keys: count (atomicAdd) -> exclusive
//            scan -> scatter; the order inside a bucket is arbitrary, which is
//            harmless because only the group element matters
//   slices : buckets are cut into slices of <= MSM_SLICE entries; one thread per slice does
//            mixed complete additions (inputs are affine).  Slicing bounds the serial work of
//            dense buckets (e.g. the partial top window) and of skewed scalar distributions.
//            Slices are handed to threads in order of decreasing length (a 65-bin counting
//            sort), so the lanes of a warp run the same number of additions
//   windows: R_w = sum_j j * B_{w,j} in two levels, a bucket being the sum of its slices.  Level 1:
//            a thread owns seg consecutive buckets and forms (run, sum) = (sum B_j,
//            sum (j-lo+1) B_j); the CTA turns the position weights lo_t = t * seg into a suffix
//            scan of the runs (sum_t t * run_t = sum_{t>=1} S_t, S_t = sum_{u>=t} run_u): no
//            scalar multiple of a point is ever computed, only log2(seg) doublings.  Level 2 is
//            the same step over the CTAs' (run, sum) pairs, one CTA per window
//   final  : Horner over the windows (c doublings each) with the 8-lane cooperative group law of
//            coop.cuh, because this chain of 256 - c doublings is the serial tail of the MSM
#pragma once
#include "fe.cuh"
#include "point.cuh"
#include "jac.cuh"
#include "sc.cuh"

namespace s256 {

constexpr int MSM_MAX_C = 16;
constexpr int MSM_SLICE = 64;   // entries per slice
constexpr int MSM_SUPER = 64;   // slices per super-slice (buckets of more than MSM_SUPER slices only)
constexpr int MSM_WT = 128;     // threads of a window-stage CTA
// buckets one thread reduces in the window stage (a power of two, like every window size):
// small, because the stage is latency bound -- each thread runs 2 * seg dependent additions
S256_HD int msm_seg_for(int nbw) {
    int seg = nbw >> 9;
    if (seg < 1) seg = 1;
    if (seg > 8) seg = 8;
    return seg;
}
S256_HD int msm_log2(int v) {
    int l = 0;
    while ((1 << (l + 1)) <= v) l++;
    return l;
}
// CTAs of MSM_WT threads a window of nbw buckets needs (<= MSM_WT, so one CTA folds them)
S256_HD int msm_parts_for(int nbw) {
    int per = MSM_WT * msm_seg_for(nbw);
    return (nbw + per - 1) / per;
}
constexpr int MSM_MAX_WIN = 64;  // c = 4 -> 64 windows

// Windows 0 .. nwin-2 use signed digits (2^(c-1) buckets each).  The top window keeps
// its digit UNSIGNED (value + incoming carry, at most 2^top_bits) so that no carry-only
// window exists: such a window would put half of all points into one bucket.
struct msm_plan {
    int c;         // window bits
    int nwin;      // windows = ceil(bits / c)
    int nb;        // buckets per signed window = 2^(c-1)
    int nb_top;    // buckets of the top window = 2^top_bits, top_bits = bits - c*(nwin-1)
    int total;     // all buckets
    int slice;     // entries per slice (<= MSM_SLICE): one thread folds one slice
};

// bits = 128: the variable-time MSM splits every scalar with the lambda endomorphism (k = k1 + k2 lambda, |k1|, |k2| <
// 2^128, point_mul_glv.go:59-117) and runs Pippenger over 2n points (P_i, lambda P_i = (beta x_i, y_i)) with 128-bit
// scalars: the same number of bucket accumulations, but HALF the buckets and half the doublings of the serial Horner
// tail (k_msm_final: (nwin - 1) c = 112 instead of 240), which is what a small per-GPU share of a sharded MSM waits for.
S256_HD msm_plan msm_plan_for_c(int c, int bits = 256) {
    msm_plan p;
    p.c = c;
    p.nwin = (bits + c - 1) / c;
    p.nb = 1 << (c - 1);
    p.nb_top = 1 << (bits - c * (p.nwin - 1));
    p.total = (p.nwin - 1) * p.nb + p.nb_top;
    p.slice = MSM_SLICE;
    return p;
}
// Window width by a cost model instead of the round-1 rule c = floor(log2 n) - 4: one bucket accumulation per point and
// window, about eight of those per bucket for the window stage (measured: 0.13-0.17 ns per accumulated entry, 0.8-2 ns
// per bucket), and NO width whose unsigned top window is only a few bits wide -- with c = 15 and 256-bit scalars the top
// window has 2^1 buckets, half of all points land in one of them, and that bucket's 4096 slices are folded by a lone
// thread (k_msm_superslices 0.41 ms, the digit kernels' atomics on one counter: n = 2^19 took 3.17 ms against 3.62 ms
// for 2^20).  n = the number of (virtual) points that carry a `bits`-bit scalar.
S256_HD msm_plan msm_make_plan(size_t n, int bits = 256) {
    int best = 4;
    double best_cost = -1.0;
    for (int c = 4; c <= MSM_MAX_C; c++) {
        msm_plan p = msm_plan_for_c(c, bits);
        int top_bits = bits - c * (p.nwin - 1);
        if ((n >> top_bits) > 2048 && c != MSM_MAX_C) continue;  // a top bucket of more than 32 slices
        double cost = (double)n * p.nwin + 8.0 * (double)p.total;
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = c;
        }
    }
    msm_plan p = msm_plan_for_c(best, bits);
    // A slice is folded by ONE thread, one dependent mixed addition after the other (~4 us each for a lone warp): with
    // few entries there are not enough slices to fill the GPU and the kernel lasts as long as its longest chain
    // (n = 2^17: 41 k slices of 64, 0.53 ms).  Shorter slices for smaller inputs: more threads, shorter chains; the
    // extra slice sums of a bucket are folded in the window stage.
    double entries = (double)n * p.nwin;
    p.slice = entries >= 8.0e6 ? MSM_SLICE : (entries >= 4.0e6 ? MSM_SLICE / 2 : MSM_SLICE / 4);
    return p;
}
S256_HD int msm_window_buckets(const msm_plan &p, int w) { return w == p.nwin - 1 ? p.nb_top : p.nb; }

// digits of a reduced scalar: d[w] in [-(2^(c-1)-1), 2^(c-1)] for w < nwin-1, d[nwin-1] in [0, 2^top_bits]
S256_HD void msm_digits(int32_t *d, const sc &k, const msm_plan &p) {
    uint32_t carry = 0;
    for (int w = 0; w < p.nwin; w++) {
        int bit = w * p.c;
        int limb = bit >> 5, sh = bit & 31;
        uint32_t v = k.v[limb] >> sh;
        if (sh + p.c > 32 && limb + 1 < 8) v |= k.v[limb + 1] << (32 - sh);
        v &= (1u << p.c) - 1u;  // bits above 255 are zero by construction
        v += carry;
        if (w == p.nwin - 1) {
            d[w] = (int32_t)v;
        } else {
            carry = (v + (1u << (p.c - 1)) - 1u) >> p.c;
            d[w] = (int32_t)v - (int32_t)(carry << p.c);
        }
    }
}

// The two halves of a scalar for the endomorphism form: magnitudes below 2^128 (zero-extended to 8 limbs, so that
// msm_digits reads them like any scalar) and their signs.  Virtual point 2i is P_i with k1, 2i + 1 is lambda P_i with k2.
S256_HD void msm_glv_halves(sc &m1, uint32_t &neg1, sc &m2, uint32_t &neg2, const sc &k) {
    uint32_t a[4], b[4];
    sc_split_glv_abs(a, neg1, b, neg2, k);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        m1.v[i] = i < 4 ? a[i] : 0u;
        m2.v[i] = i < 4 ? b[i] : 0u;
    }
}
// the virtual points of P: (x, y) and (beta x, y)
S256_HD void msm_glv_points(apt &p0, apt &p1, const apt &p) {
    p0 = p;
    fe_mul(p1.x, p.x, fe_beta());
    p1.y = p.y;
}

// slice / bucket accumulation: entries hold (point index << 1) | negate
S256_HD void msm_bucket_sum(pt &out, const uint32_t *entries, uint32_t start, uint32_t end, const apt *aff) {
#ifndef S256_MSM_RCB
    // Jacobian accumulator, mixed Jacobian additions with their exceptional cases handled by branches (jac.cuh): the
    // vartime MSM handles public data like the verification ladder, and equal points in one bucket (repeated inputs,
    // or P and -P with opposite digits) do reach the doubling and the identity here
    fe_ops<true> f;
    pt acc;
    acc.x = acc.y = acc.z = fe_zero();
    uint32_t inf = 1u;
    if (start < end) {
        uint32_t v = entries[start];
        apt a = aff[v >> 1];
        for (uint32_t e = start; e < end; e++) {
            uint32_t vn = v;
            apt an = a;
            if (e + 1 < end) {
                vn = entries[e + 1];
                an = aff[vn >> 1];
            }
            if (v & 1u) {
                fe z = fe_zero();
                fe_sub_vt(a.y, z, a.y);
            }
            jac_add_mixed_var(f, acc, inf, a.x, a.y);
            v = vn;
            a = an;
        }
    }
    jac_to_projective(f, out, acc, inf);
#else
    pt acc;
    pt_set_identity(acc);
    if (start < end) {
        // the next point is fetched before the current addition: the gather latency hides behind it
        uint32_t v = entries[start];
        apt a = aff[v >> 1];
        for (uint32_t e = start; e < end; e++) {
            uint32_t vn = v;
            apt an = a;
            if (e + 1 < end) {
                vn = entries[e + 1];
                an = aff[vn >> 1];
            }
            if (v & 1u) {
                fe z = fe_zero();
                fe_sub_vt(a.y, z, a.y);
            }
            pt_add_mixed<true>(acc, acc, a.x, a.y);
            v = vn;
            a = an;
        }
    }
    out = acc;
#endif
}

// slices of bucket b: max(1, ceil(count / MSM_SLICE))
S256_HD uint32_t msm_slices_of(uint32_t count, int slice) {
    return count == 0 ? 1u : (count + (uint32_t)slice - 1u) / (uint32_t)slice;
}

// A bucket of more than MSM_SUPER slices (> 4096 entries: equal or adversarial scalars) gets a second
// level: the thread of every MSM_SUPER-th slice folds the next MSM_SUPER slice sums into its own slot.
// [s0, s1) = the slices of the bucket that owns slice s.
S256_HD void msm_superslice_fold(pt *slice_sum, uint32_t s, uint32_t s0, uint32_t s1) {
    if (s1 - s0 <= (uint32_t)MSM_SUPER || ((s - s0) % (uint32_t)MSM_SUPER) != 0) return;
    uint32_t e = s + (uint32_t)MSM_SUPER;
    if (e > s1) e = s1;
    pt acc = slice_sum[s];
    for (uint32_t q = s + 1; q < e; q++) {
        pt t = slice_sum[q];
        pt_add<true>(acc, acc, t);
    }
    slice_sum[s] = acc;
}

// bucket b = sum of its slices (almost always exactly one), or of its super-slices
S256_HD void msm_bucket_from_slices(pt &out, const pt *slice_sum, const uint32_t *sl_off, uint32_t b) {
    uint32_t s0 = sl_off[b], s1 = sl_off[b + 1];
    uint32_t step = s1 - s0 > (uint32_t)MSM_SUPER ? (uint32_t)MSM_SUPER : 1u;
    out = slice_sum[s0];
    for (uint32_t s = s0 + step; s < s1; s += step) {
        pt q = slice_sum[s];
        pt_add<true>(out, out, q);
    }
}

// run = sum_{j in [lo, hi)} B_j and sum = sum_j (j - lo + 1) * B_j for the window whose first
// bucket is `base` (lo < hi).  The weights are relative to lo: the caller scales by position.
S256_HD void msm_segment_pair(pt &run, pt &sum, const pt *slice_sum, const uint32_t *sl_off, uint32_t base, int lo,
                              int hi) {
    msm_bucket_from_slices(run, slice_sum, sl_off, base + (uint32_t)(hi - 1));
    sum = run;
    for (int j = hi - 1; j > lo; j--) {
        pt b;
        msm_bucket_from_slices(b, slice_sum, sl_off, base + (uint32_t)(j - 1));
        pt_add<true>(run, run, b);
        pt_add<true>(sum, sum, run);
    }
}
// v = sum + 2^log2w * s   (the position weight of a thread's / CTA's run total)
S256_HD void msm_weigh(pt &v, const pt &sum, const pt &s, int log2w) {
    pt m = s;
    for (int k = 0; k < log2w; k++) pt_double<true>(m, m);
    pt_add<true>(v, sum, m);
}

// Horner over window results, highest first; window w is the sum of `parts` partials at win[w * stride ..]
S256_HD void msm_horner(pt &out, const pt *win, const msm_plan &p, int parts, int stride) {
    pt acc;
    pt_set_identity(acc);
    for (int w = p.nwin - 1; w >= 0; w--) {
        if (w != p.nwin - 1)
            for (int k = 0; k < p.c; k++) pt_double<true>(acc, acc);
        for (int q = 0; q < parts; q++) {
            pt t = win[w * stride + q];
            pt_add<true>(acc, acc, t);
        }
    }
    out = acc;
}
// which entry range does slice s cover?  binary search for the bucket (returned), then the range
S256_HD uint32_t msm_slice_range(uint32_t &start, uint32_t &end, uint32_t s, const uint32_t *sl_off,
                                 const uint32_t *offsets, uint32_t total_buckets, int slice) {
    uint32_t lo = 0, hi = total_buckets;  // invariant: sl_off[lo] <= s < sl_off[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (sl_off[mid] <= s) lo = mid; else hi = mid;
    }
    uint32_t k = s - sl_off[lo];
    start = offsets[lo] + k * (uint32_t)slice;
    end = start + (uint32_t)slice;
    if (end > offsets[lo + 1]) end = offsets[lo + 1];
    if (start > end) start = end;
    return lo;  // the bucket
}

// projective point <-> 96-byte big-endian X || Y || Z (the cross-GPU partial)
S256_HD void pt_to_be96(uint8_t *b, const pt &p) {
    fe t;
    fe_normalize(t, p.x); fe_to_be32(b, t);
    fe_normalize(t, p.y); fe_to_be32(b + 32, t);
    fe_normalize(t, p.z); fe_to_be32(b + 64, t);
}
S256_HD void pt_from_be96(pt &p, const uint8_t *b) {
    fe_from_be32(p.x, b);
    fe_from_be32(p.y, b + 32);
    fe_from_be32(p.z, b + 64);
}
// Y^2 Z == X^3 + 7 Z^3, or the identity (0 : y : 0)
S256_HD uint32_t pt_on_curve(const pt &p) {
    fe l, r, t, z2;
    fe_sqr(l, p.y); fe_mul(l, l, p.z);
    fe_sqr(t, p.x); fe_mul(r, t, p.x);
    fe_sqr(z2, p.z); fe_mul(t, z2, p.z);
    fe_mul_small(t, t, 7u);
    fe_add(r, r, t);
    uint32_t ident = fe_is_zero(p.z) & fe_is_zero(p.x) & (1u - fe_is_zero(p.y));
    return (fe_equal(l, r) & (1u - fe_is_zero(p.z))) | ident;
}

}  // namespace s256
