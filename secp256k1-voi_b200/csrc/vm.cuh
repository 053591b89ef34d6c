// vm.cuh -- the point formulas as sequences of slot operations ("frame" form).
//
// Why: in register form every out-of-line fe_mul call makes ptxas marshal 16
// argument / 8 result registers with IMAD.MOV.U32, i.e. on the same (fmaheavy)
// pipe the multiplies need (profiles/: 1048 IMAD.MOV for 51 call sites, ~14 %
// of that pipe), and inlining instead blows the instruction cache.  Here each
// thread owns VM_SLOTS field-element slots in SHARED memory; mul / sqr / add /
// sub are small out-of-line routines that take slot addresses, load operands
// with LDS.128 and store the result with STS.128 -- operand traffic moves to
// the otherwise idle LSU pipe, the ladder body shrinks to a few KB of calls,
// and the kernel needs ~64 registers instead of 128-168.
//
// The formulas (Renes-Costello-Batina Alg. 7/8/9, point_projective.go:24-273)
// are written once against a Frame type: DevFrame (shared memory, device) and
// HostFrame (plain arrays, portable arithmetic) so that tests/hostsim executes
// the identical operation sequences.
#pragma once
#include "fe.cuh"
#include "point.cuh"

namespace s256 {

constexpr int VM_SLOTS = 13;
// slot map: accumulator, addend, temporaries
enum { SX = 0, SY = 1, SZ = 2, AX = 3, AY = 4, AZ = 5, T0 = 6, T1 = 7, T2 = 8, T3 = 9, T4 = 10, T5 = 11, TU = 12 };

// ---------------------------------------------------------------------------
// host frame (also the semantic definition of every op)
// ---------------------------------------------------------------------------
struct HostFrame {
    fe s[VM_SLOTS];
    S256_HD void mul(int d, int a, int b) { fe_mul(s[d], s[a], s[b]); }
    S256_HD void sqr(int d, int a) { fe_sqr(s[d], s[a]); }
    S256_HD void add(int d, int a, int b) { fe_add(s[d], s[a], s[b]); }
    S256_HD void sub(int d, int a, int b) { fe_sub(s[d], s[a], s[b]); }
    S256_HD void mul21(int d, int a) { fe_mul_small(s[d], s[a], S256_B3); }
    S256_HD void mul_beta(int d, int a) { fe b = fe_beta(); fe_mul(s[d], s[a], b); }
    S256_HD void neg(int d, int a) { fe_neg(s[d], s[a]); }
    S256_HD void set(int d, const fe &v) { s[d] = v; }
    S256_HD fe get(int a) const { return s[a]; }
    S256_HD void load_pt(int d, const pt *p) { s[d] = p->x; s[d + 1] = p->y; s[d + 2] = p->z; }
    S256_HD void load_apt(int d, const apt *p) { s[d] = p->x; s[d + 1] = p->y; }
    S256_HD void store_pt(pt *p, int a) const { p->x = s[a]; p->y = s[a + 1]; p->z = s[a + 2]; }
};

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------
// device frame: slot `k`, limb group g (0: limbs 0-3, 1: limbs 4-7) of thread t
// lives at uint4 index (2k + g) * TPB + t  -> LDS.128 / STS.128, conflict free.
// ---------------------------------------------------------------------------
template <int TPB>
struct DevFrameOps {
    static __device__ __forceinline__ void ld(fe &r, const uint4 *sm, uint32_t idx) {
        uint4 lo = sm[idx], hi = sm[idx + TPB];
        r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
        r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    }
    static __device__ __forceinline__ void st(uint4 *sm, uint32_t idx, const fe &r) {
        sm[idx] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
        sm[idx + TPB] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
    }
    // the out-of-line workers: three small integers in, no field element ever crosses the call ABI
    static __device__ __noinline__ void w_mul(uint32_t d, uint32_t a, uint32_t b) {
        extern __shared__ uint4 vm_smem[];
        fe x, y, r;
        ld(x, vm_smem, a);
        ld(y, vm_smem, b);
        fe_mul_inline(r, x, y);
        st(vm_smem, d, r);
    }
    static __device__ __noinline__ void w_sqr(uint32_t d, uint32_t a) {
        extern __shared__ uint4 vm_smem[];
        fe x, r;
        ld(x, vm_smem, a);
        fe_sqr_inline(r, x);
        st(vm_smem, d, r);
    }
    static __device__ __noinline__ void w_mul_beta(uint32_t d, uint32_t a) {
        extern __shared__ uint4 vm_smem[];
        fe x, r;
        ld(x, vm_smem, a);
        const fe b = fe_beta();
        fe_mul_inline(r, x, b);
        st(vm_smem, d, r);
    }
    static __device__ __noinline__ void w_mul21(uint32_t d, uint32_t a) {
        extern __shared__ uint4 vm_smem[];
        fe x, r;
        ld(x, vm_smem, a);
        fe_mul_small(r, x, S256_B3);
        st(vm_smem, d, r);
    }
    static __device__ __noinline__ void w_add(uint32_t d, uint32_t a, uint32_t b) {
        extern __shared__ uint4 vm_smem[];
        fe x, y, r;
        ld(x, vm_smem, a);
        ld(y, vm_smem, b);
        fe_add(r, x, y);
        st(vm_smem, d, r);
    }
    static __device__ __noinline__ void w_sub(uint32_t d, uint32_t a, uint32_t b) {
        extern __shared__ uint4 vm_smem[];
        fe x, y, r;
        ld(x, vm_smem, a);
        ld(y, vm_smem, b);
        fe_sub(r, x, y);
        st(vm_smem, d, r);
    }
};

template <int TPB>
struct DevFrame {
    using W = DevFrameOps<TPB>;
    uint32_t t;  // threadIdx.x
    __device__ __forceinline__ uint32_t at(int k) const { return (uint32_t)(2 * k) * TPB + t; }
    __device__ __forceinline__ void mul(int d, int a, int b) { W::w_mul(at(d), at(a), at(b)); }
    __device__ __forceinline__ void sqr(int d, int a) { W::w_sqr(at(d), at(a)); }
    __device__ __forceinline__ void add(int d, int a, int b) { W::w_add(at(d), at(a), at(b)); }
    __device__ __forceinline__ void sub(int d, int a, int b) { W::w_sub(at(d), at(a), at(b)); }
    __device__ __forceinline__ void mul21(int d, int a) { W::w_mul21(at(d), at(a)); }
    __device__ __forceinline__ void mul_beta(int d, int a) { W::w_mul_beta(at(d), at(a)); }
    __device__ __forceinline__ void neg(int d, int a) {
        extern __shared__ uint4 vm_smem[];
        fe x, r;
        W::ld(x, vm_smem, at(a));
        fe_neg(r, x);
        W::st(vm_smem, at(d), r);
    }
    __device__ __forceinline__ void set(int d, const fe &v) {
        extern __shared__ uint4 vm_smem[];
        W::st(vm_smem, at(d), v);
    }
    __device__ __forceinline__ fe get(int a) const {
        extern __shared__ uint4 vm_smem[];
        fe r;
        W::ld(r, vm_smem, at(a));
        return r;
    }
    // global <-> slots, 128-bit accesses (pt / apt are 16-byte aligned arrays of limbs)
    __device__ __forceinline__ void load_words(int d, const uint4 *g, int nfe) {
        extern __shared__ uint4 vm_smem[];
#pragma unroll
        for (int k = 0; k < 3; k++)
            if (k < nfe) {
                vm_smem[at(d + k)] = g[2 * k];
                vm_smem[at(d + k) + TPB] = g[2 * k + 1];
            }
    }
    __device__ __forceinline__ void load_pt(int d, const pt *p) { load_words(d, reinterpret_cast<const uint4 *>(p), 3); }
    __device__ __forceinline__ void load_apt(int d, const apt *p) { load_words(d, reinterpret_cast<const uint4 *>(p), 2); }
    __device__ __forceinline__ void store_pt(pt *p, int a) const {
        extern __shared__ uint4 vm_smem[];
        uint4 *g = reinterpret_cast<uint4 *>(p);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            g[2 * k] = vm_smem[at(a + k)];
            g[2 * k + 1] = vm_smem[at(a + k) + TPB];
        }
    }
};
#endif  // __CUDACC__

// ---------------------------------------------------------------------------
// formulas.  Accumulator in (SX, SY, SZ), addend in (AX, AY[, AZ]); the
// accumulator is overwritten with the result.  Addend slots are preserved.
// ---------------------------------------------------------------------------

// point_projective.go:24-120 -- complete addition, 12 M + 2 m3b + 19 a
template <class F>
S256_HD void vm_pt_add(F &f) {
    f.mul(T0, SX, AX);
    f.mul(T1, SY, AY);
    f.mul(T2, SZ, AZ);
    f.add(T3, SX, SY); f.add(TU, AX, AY); f.mul(T3, T3, TU);
    f.add(T4, SY, SZ); f.add(TU, AY, AZ); f.mul(T4, T4, TU);
    f.add(T5, SX, SZ); f.add(TU, AX, AZ); f.mul(T5, T5, TU);
    f.add(TU, T0, T1); f.sub(T3, T3, TU);
    f.add(TU, T1, T2); f.sub(T4, T4, TU);
    f.add(TU, T0, T2); f.sub(T5, T5, TU);   // Y3 of the paper's step 6
    f.add(TU, T0, T0); f.add(T0, TU, T0);   // 3 * t0
    f.mul21(T2, T2);
    f.add(SZ, T1, T2);
    f.sub(T1, T1, T2);
    f.mul21(T5, T5);
    f.mul(SX, T4, T5);
    f.mul(T2, T3, T1);
    f.sub(SX, T2, SX);
    f.mul(T5, T5, T0);
    f.mul(T1, T1, SZ);
    f.add(SY, T1, T5);
    f.mul(T0, T0, T3);
    f.mul(SZ, SZ, T4);
    f.add(SZ, SZ, T0);
}

// point_projective.go:123-205 -- mixed addition (addend affine, not the identity), 11 M + 2 m3b + 13 a
template <class F>
S256_HD void vm_pt_add_mixed(F &f) {
    f.mul(T0, SX, AX);
    f.mul(T1, SY, AY);
    f.add(T3, AX, AY); f.add(T4, SX, SY); f.mul(T3, T3, T4);
    f.add(T4, T0, T1); f.sub(T3, T3, T4);
    f.mul(T4, AY, SZ); f.add(T4, T4, SY);
    f.mul(T5, AX, SZ); f.add(T5, T5, SX);
    f.add(TU, T0, T0); f.add(T0, TU, T0);
    f.mul21(T2, SZ);
    f.add(SZ, T1, T2);
    f.sub(T1, T1, T2);
    f.mul21(T5, T5);
    f.mul(SX, T4, T5);
    f.mul(T2, T3, T1);
    f.sub(SX, T2, SX);
    f.mul(T5, T5, T0);
    f.mul(T1, T1, SZ);
    f.add(SY, T1, T5);
    f.mul(T0, T0, T3);
    f.mul(SZ, SZ, T4);
    f.add(SZ, SZ, T0);
}

// point_projective.go:208-273 -- complete doubling, 6 M + 2 S + 1 m3b + 9 a
template <class F>
S256_HD void vm_pt_double(F &f) {
    f.sqr(T0, SY);
    f.add(T3, T0, T0); f.add(T3, T3, T3); f.add(T3, T3, T3);   // 8 * Y^2
    f.mul(T1, SY, SZ);
    f.sqr(T2, SZ);
    f.mul21(T2, T2);
    f.mul(T4, T2, T3);
    f.add(T5, T0, T2);
    f.mul(SZ, T1, T3);
    f.add(T1, T2, T2); f.add(T2, T1, T2);
    f.sub(T0, T0, T2);
    f.mul(T5, T0, T5);
    f.mul(T1, SX, SY);
    f.add(SY, T4, T5);
    f.mul(SX, T0, T1);
    f.add(SX, SX, SX);
}

template <class F>
S256_HD void vm_set_identity(F &f) {
    f.set(SX, fe_zero());
    f.set(SY, fe_one());
    f.set(SZ, fe_zero());
}

}  // namespace s256
