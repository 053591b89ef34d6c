// api.cu -- C ABI (include/secp256k1_b200.h) over the sm_100a kernels.
//
// Host side only orchestrates: device buffers, copies, launches.  There is no
// CPU implementation of any curve operation in this library: without a CUDA
// device every entry point fails with S256_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "../../include/secp256k1_b200.h"
#include "kernels.cuh"
#include "microbench.cuh"
#include "msm.cuh"
#include <cub/device/device_scan.cuh>
#include "launchers.h"

using namespace s256;

// ---------------------------------------------------------------------------
// __global__ wrappers
// ---------------------------------------------------------------------------
#define S256_TPB 128

__global__ void __launch_bounds__(S256_TPB) k_gen_table(apt *out, int wb, size_t total) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    uint32_t w = (uint32_t)(idx >> wb), d = (uint32_t)(idx & ((1u << wb) - 1u));
    apt a;
    item_gen_multiple(a, w, d, wb);
    out[idx] = a;
}
// the signed constant-time table: out[w][j] = (j + 1) * 2^(CT_WB*w) * G
__global__ void __launch_bounds__(S256_TPB) k_gen_ct_table(apt *out) {
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint32_t)(CT_NW * CT_SZ)) return;
    apt a;
    item_gen_multiple(a, idx / CT_SZ, idx % CT_SZ + 1u, CT_WB);
    out[idx] = a;
}

__global__ void __launch_bounds__(S256_TPB) k_decode_uncompressed(const uint8_t *pt65, size_t n, apt *aff,
                                                                  uint8_t *pvalid) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    apt a;
    pvalid[i] = item_decode_uncompressed(a, pt65 + 65 * i);
    aff[i] = a;
}

__global__ void __launch_bounds__(S256_TPB) k_decode_compressed(const uint8_t *pt33, size_t n, apt *aff,
                                                                uint8_t *pvalid) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    apt a;
    pvalid[i] = item_decode_compressed(a, pt33 + 33 * i);
    aff[i] = a;
}
// affine (validated) -> 65-byte encoding; invalid -> zeros
__global__ void __launch_bounds__(S256_TPB) k_encode_affine(const apt *aff, const uint8_t *pvalid, size_t n,
                                                            uint8_t *out65, uint8_t *status) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    apt a = aff[i];
    uint32_t ok = pvalid[i] != 0;
    uint8_t *o = out65 + 65 * i;
    if (ok) {
        o[0] = 0x04;
        fe_to_be32(o + 1, a.x);
        fe_to_be32(o + 33, a.y);
    } else {
        for (int b = 0; b < 65; b++) o[b] = 0;
    }
    status[i] = ok ? ST_OK : ST_INVALID;
}

// BIP-340 lift_x: x-only key, even y (secec/bitcoin/schnorr.go:257-275)
__global__ void __launch_bounds__(S256_TPB) k_decode_xonly(const uint8_t *pkx32, size_t n, apt *aff, uint8_t *pvalid) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    apt a;
    pvalid[i] = item_decode_xonly(a, pkx32 + 32 * i);
    aff[i] = a;
}

// RecoverPoint (point_s11n.go:245-282): x = r (+ n if v & 2), parity v & 1
__global__ void __launch_bounds__(S256_TPB) k_decode_recover(const uint8_t *sig65, size_t n, apt *aff,
                                                             uint8_t *pvalid) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    apt a;
    pvalid[i] = item_decode_recover(a, sig65 + 65 * i);
    aff[i] = a;
}

template <int K, bool RECOVER>
__global__ void __launch_bounds__(S256_TPB) k_ecdsa_scalars(const uint8_t *digest32, const uint8_t *sig, size_t n,
                                                            uint32_t flags, sc *u1, int8_t *dig1, int8_t *dig2,
                                                            uint8_t *sfl) {
    size_t stride = (n + K - 1) / K;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= stride) return;
    if (!RECOVER)
        group_ecdsa_scalars<K>(t, stride, n, digest32, sig, flags, u1, dig1, dig2, sfl);
    else
        group_recover_scalars<K>(t, stride, n, digest32, sig, u1, dig1, dig2, sfl);
}

__global__ void __launch_bounds__(S256_TPB) k_plain_scalars(const uint8_t *u1b, const uint8_t *u2b, size_t n, sc *u1,
                                                            int8_t *dig1, int8_t *dig2, uint8_t *sfl) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    item_plain_scalars(i, n, u1b, u2b, u1, dig1, dig2, sfl);
}

// secec/bitcoin/schnorr.go:420-449: r < p, s < n (zero allowed), e = H(r||P||m) mod n;
// R = s*G + (-e)*P (:244-245)
__global__ void __launch_bounds__(S256_TPB) k_schnorr_scalars(const uint8_t *pkx32, const uint8_t *msg, size_t msg_len,
                                                              const uint8_t *sig64, size_t n, sc *u1, int8_t *dig1,
                                                              int8_t *dig2, uint8_t *sfl) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    item_schnorr_scalars(i, n, pkx32, msg, msg_len, sig64, u1, dig1, dig2, sfl);
}

#ifndef S256_DSM_MINB
#define S256_DSM_MINB 4
#endif
__global__ void __launch_bounds__(S256_TPB, S256_DSM_MINB)
    k_dsm(size_t n, const apt *aff, const sc *u1, const int8_t *dig1, const int8_t *dig2, const uint8_t *sfl, pt *tbl,
          pt *res, const apt *comb) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    item_dsm(i, n, aff, u1, dig1, dig2, sfl, tbl, res, comb);
}

#ifndef S256_VM_MINB
#define S256_VM_MINB 4
#endif
__global__ void __launch_bounds__(S256_TPB, S256_VM_MINB)
    k_dsm_vm(size_t n, const apt *aff, const sc *u1, const int8_t *dig1, const int8_t *dig2, const uint8_t *sfl, pt *tbl,
             pt *res, const apt *comb) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    DevFrame<S256_TPB> f{threadIdx.x};
    item_dsm_vm(f, i, n, aff, u1, dig1, dig2, sfl, tbl, res, comb);
}
constexpr size_t VM_SMEM_BYTES = (size_t)VM_SLOTS * 2 * S256_TPB * sizeof(uint4);

#ifndef S256_SM_MINB
#define S256_SM_MINB 4
#endif
__global__ void __launch_bounds__(S256_TPB, S256_SM_MINB)
    k_scalar_mult_ct(size_t n, const apt *aff, const uint8_t *k32, pt *tbl, pt *res) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#ifdef S256_CT_TABLE_GLOBAL
    CtTableGlobal T{tbl + i * (size_t)DSM_TS};
#else
    CtTableShared<S256_TPB> T{threadIdx.x};
#endif
    item_scalar_mult_ct(i, aff, k32, T, res);
}
#ifdef S256_CT_TABLE_GLOBAL
constexpr size_t CT_SMEM_BYTES = 0;
#else
constexpr size_t CT_SMEM_BYTES = (size_t)CTM_TS * 6 * S256_TPB * sizeof(uint4);
#endif
static void s256_launch_scalar_mult_ct(size_t n, const apt *aff, const uint8_t *k32, pt *tbl, pt *res, cudaStream_t s) {
    if (n == 0) return;
    k_scalar_mult_ct<<<(unsigned)((n + S256_TPB - 1) / S256_TPB), S256_TPB, CT_SMEM_BYTES, s>>>(n, aff, k32, tbl, res);
}


__global__ void __launch_bounds__(S256_TPB) k_ecdsa_finish(size_t n, const pt *res, const uint8_t *sig64,
                                                           const uint8_t *pvalid, const uint8_t *sfl, uint8_t *ok) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pt R = res[i];
    uint32_t valid = (uint32_t)(pvalid[i] != 0) & (uint32_t)(sfl[i] & SFL_VALID);
    ok[i] = item_ecdsa_finish(R, sig64 + 64 * i, valid);
}

// in-status = pvalid (may be null) AND sfl valid bit (may be null)
template <int K>
__global__ void __launch_bounds__(S256_TPB) k_finish_affine(size_t n, const pt *res, const uint8_t *pvalid,
                                                            const uint8_t *sfl, uint8_t *comb_status, int mode,
                                                            uint8_t *out, uint8_t *status, const uint8_t *sig64) {
    size_t stride = (n + K - 1) / K;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= stride) return;
    group_finish<K>(t, stride, n, res, pvalid, sfl, comb_status, mode, out, status, sig64);
}

// ---- deterministic signing (kernels.cuh) ----
__global__ void __launch_bounds__(S256_TPB) k_rfc6979_nonce(const uint8_t *priv32, const uint8_t *digest32, size_t n,
                                                            uint8_t *kbuf, uint8_t *valid) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    valid[i] = item_rfc6979_nonce(kbuf + 32 * i, priv32 + 32 * i, digest32 + 32 * i);
}
template <int K>
__global__ void __launch_bounds__(S256_TPB) k_sign_finish(size_t n, const uint8_t *priv32, const uint8_t *digest32,
                                                          const uint8_t *kbuf, const uint8_t *valid, const uint8_t *r65,
                                                          uint8_t *sig64, uint8_t *recid, uint8_t *status) {
    size_t stride = (n + K - 1) / K;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= stride) return;
    group_sign_finish<K>(t, stride, n, priv32, digest32, kbuf, valid, r65, sig64, recid, status);
}

__global__ void __launch_bounds__(S256_TPB) k_schnorr_nonce(const uint8_t *priv32, const uint8_t *p65, const uint8_t *msg,
                                                            size_t msg_len, const uint8_t *aux32, size_t n,
                                                            uint8_t *kbuf, uint8_t *valid) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    valid[i] = item_schnorr_nonce(kbuf + 32 * i, priv32 + 32 * i, p65 + 65 * i, msg + msg_len * i, msg_len, aux32 + 32 * i);
}
__global__ void __launch_bounds__(S256_TPB) k_schnorr_sign_finish(const uint8_t *priv32, const uint8_t *p65,
                                                                  const uint8_t *r65, const uint8_t *kbuf,
                                                                  const uint8_t *msg, size_t msg_len,
                                                                  const uint8_t *valid, size_t n, uint8_t *sig64,
                                                                  uint8_t *status) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    item_schnorr_sign_finish(sig64 + 64 * i, status + i, priv32 + 32 * i, p65 + 65 * i, r65 + 65 * i, kbuf + 32 * i,
                             msg + msg_len * i, msg_len, valid[i]);
}

// ---- Pippenger MSM (msm.cuh) -------------------------------------------------
template <bool SCATTER>
__global__ void __launch_bounds__(S256_TPB) k_msm_digits(const uint8_t *k32, size_t n, msm_plan plan, uint32_t *counts,
                                                         uint32_t *cursor, uint32_t *entries) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sc k;
    sc_from_be32(k, k32 + 32 * i);
    int32_t d[MSM_MAX_WIN];
    msm_digits(d, k, plan);
    for (int w = 0; w < plan.nwin; w++) {
        int32_t dw = d[w];
        if (dw == 0) continue;
        uint32_t mag = (uint32_t)(dw < 0 ? -dw : dw);
        uint32_t b = (uint32_t)w * (uint32_t)plan.nb + (mag - 1u);
        if (!SCATTER) {
            atomicAdd(&counts[b], 1u);
        } else {
            uint32_t pos = atomicAdd(&cursor[b], 1u);
            entries[pos] = ((uint32_t)i << 1) | (uint32_t)(dw < 0);
        }
    }
}
// nsl[b] = slices of bucket b (counts -> slice counts), then scanned into sl_off
__global__ void __launch_bounds__(S256_TPB) k_msm_slice_counts(uint32_t total, const uint32_t *counts, uint32_t *nsl) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < total) nsl[b] = msm_slices_of(counts[b]);
    if (b == total) nsl[b] = 0;
}
__global__ void __launch_bounds__(S256_TPB) k_msm_slices(uint32_t max_slices, uint32_t total, const uint32_t *sl_off,
                                                         const uint32_t *offsets, const uint32_t *entries,
                                                         const apt *aff, pt *slice_sum) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= max_slices || s >= sl_off[total]) return;
    uint32_t st, en;
    msm_slice_range(st, en, s, sl_off, offsets, total);
    pt r;
    msm_bucket_sum(r, entries, st, en, aff);
    slice_sum[s] = r;
}
#define S256_MSM_WT 128
// grid (blocks per window, nwin): every thread reduces MSM_SEG buckets, the CTA folds its threads
__global__ void __launch_bounds__(S256_MSM_WT) k_msm_windows(msm_plan plan, const pt *slice_sum, const uint32_t *sl_off,
                                                             pt *winpart, int parts) {
    __shared__ pt sh[S256_MSM_WT];
    int w = blockIdx.y, t = threadIdx.x;
    int nbw = msm_window_buckets(plan, w);
    int seg = msm_seg_for(nbw);
    int lo = (blockIdx.x * S256_MSM_WT + t) * seg, hi = lo + seg;
    if (hi > nbw) hi = nbw;
    pt s;
    if (lo < hi)
        msm_segment(s, slice_sum, sl_off, (uint32_t)w * (uint32_t)plan.nb, lo, hi);
    else
        pt_set_identity(s);
    sh[t] = s;
    __syncthreads();
    for (int stride = S256_MSM_WT / 2; stride >= 1; stride >>= 1) {
        if (t < stride) {
            pt a = sh[t], b = sh[t + stride];
            pt_add(a, a, b);
            sh[t] = a;
        }
        __syncthreads();
    }
    if (t == 0) winpart[w * parts + blockIdx.x] = sh[0];
}
// win[w * parts] = sum of the `parts` CTA partials of window w (one thread per window)
__global__ void k_msm_fold(int nwin, pt *winpart, int parts) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwin) return;
    pt acc = winpart[w * parts];
    for (int q = 1; q < parts; q++) {
        pt t = winpart[w * parts + q];
        pt_add(acc, acc, t);
    }
    winpart[w * parts] = acc;
}
// acc (device, projective) += Horner(window partials); first = overwrite
__global__ void k_msm_final(msm_plan plan, const pt *winpart, int parts, pt *acc, int first) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    pt r;
    msm_horner(r, winpart, plan, 1, parts);
    if (!first) {
        pt a = *acc;
        pt_add(r, r, a);
    }
    *acc = r;
}
// out[t] = sum of in[t], in[t + nout], ...   (tree levels of the constant-time MSM)
__global__ void __launch_bounds__(S256_TPB) k_reduce_points(const pt *in, size_t n, pt *out, size_t nout) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nout) return;
    pt acc;
    pt_set_identity(acc);
    for (size_t i = t; i < n; i += nout) {
        pt q = in[i];
        pt_add(acc, acc, q);
    }
    out[t] = acc;
}
__global__ void k_acc_point(const pt *in, pt *acc, int first) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    pt r = *in;
    if (!first) {
        pt a = *acc;
        pt_add(r, r, a);
    }
    *acc = r;
}
__global__ void __launch_bounds__(S256_TPB) k_any_invalid(const uint8_t *pvalid, size_t n, uint32_t *flag) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && pvalid[i] == 0) atomicOr(flag, 1u);
}
// partial96 rows -> one projective sum; rows must be points on the curve (or the identity)
__global__ void k_combine_partials(const uint8_t *partials96, size_t m, pt *out, uint32_t *flag) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    pt acc;
    pt_set_identity(acc);
    for (size_t j = 0; j < m; j++) {
        pt q;
        pt_from_be96(q, partials96 + 96 * j);
        uint32_t ok = fe_limbs_are_canonical(q.x) & fe_limbs_are_canonical(q.y) & fe_limbs_are_canonical(q.z) &
                      pt_on_curve(q);
        if (!ok) {
            atomicOr(flag, 1u);
            continue;
        }
        pt_add(acc, acc, q);
    }
    *out = acc;
}
__global__ void k_export_partial(const pt *acc, uint8_t *out96) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    pt a = *acc;
    pt_to_be96(out96, a);
}

__global__ void __launch_bounds__(S256_TPB) k_field_op(int op, const uint8_t *a32, const uint8_t *b32, size_t n,
                                                       uint8_t *out32) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (op < 16) {
        fe a, b, r;
        fe_from_be32(a, a32 + 32 * i);
        fe_from_be32(b, b32 + 32 * i);
        switch (op) {
            case 0: fe_mul(r, a, b); break;
            case 1: fe_add(r, a, b); break;
            case 2: fe_sub(r, a, b); break;
            case 3: fe_invert(r, a); break;
            case 4: fe_sqrt(r, a); break;
            case 5: fe_mul_small(r, a, 21u); break;
            case 6: fe_sqr(r, a); break;
            default: r = fe_zero();
        }
        fe_normalize(r, r);
        fe_to_be32(out32 + 32 * i, r);
    } else {
        sc a, b, r;
        sc_from_be32(a, a32 + 32 * i);
        sc_from_be32(b, b32 + 32 * i);
        switch (op) {
            case 16: sc_mul(r, a, b); break;
            case 17: sc_add(r, a, b); break;
            case 18: sc_invert(r, a); break;
            default: r = sc_zero();
        }
        sc_to_be32(out32 + 32 * i, r);
    }
}

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
struct s256_ctx {
    int device = -1;
    size_t cap = 0;
    std::mutex mu;
    cudaStream_t stream = nullptr, stream2 = nullptr;
    std::string last_err;
    std::atomic<uint64_t> launches{0};
    // constant tables
    apt *comb = nullptr;    // [COMB_NW][COMB_SZ]
    apt *ct_tab = nullptr;  // [CT_NW][CT_SZ]
    // per-chunk scratch
    apt *aff = nullptr;
    sc *u1 = nullptr;
    int8_t *dig1 = nullptr, *dig2 = nullptr;
    uint8_t *sfl = nullptr, *pvalid = nullptr, *cstat = nullptr;
    pt *tbl = nullptr, *res = nullptr;
    // staging for the host-pointer entry points
    uint8_t *in_a = nullptr, *in_b = nullptr, *in_c = nullptr, *out = nullptr, *st = nullptr;
    size_t in_b_bytes = 0;
    unsigned long long *sink = nullptr;
    // MSM scratch (allocated on first use)
    size_t msm_cap = 0;
    uint32_t *msm_counts = nullptr, *msm_offsets = nullptr, *msm_cursor = nullptr, *msm_entries = nullptr;
    uint32_t *msm_flag = nullptr;
    uint32_t *msm_nsl = nullptr, *msm_sloff = nullptr;
    size_t msm_max_slices = 0;
    pt *msm_buckets = nullptr, *msm_win = nullptr, *msm_acc = nullptr, *msm_tmp = nullptr;
    void *msm_cub = nullptr;
    size_t msm_cub_bytes = 0;
    // optional per-kernel timing of the dominant kernel (bench.py roofline)
    bool profiling = false;
    // sub-chunks per host-pointer call (S256_PIPE_PARTS).  Measured (scripts/e2e_parts.py): splitting does not
    // pay -- the batched-inversion kernel is latency bound, so its cost multiplies with the part count.
    int pipe_parts = 1;
    cudaEvent_t ev_decode = nullptr;
    bool use_reg_ladder = true;  // S256_LADDER=vm selects the frame-form ladder (A/B measurements)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> dsm_events;
};

#define CK(call)                                                                     \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) {                                                     \
            ctx->last_err = std::string(#call) + ": " + cudaGetErrorString(e_);      \
            return S256_ERR_CUDA;                                                    \
        }                                                                            \
    } while (0)

static inline unsigned grid_for(size_t n) { return (unsigned)((n + S256_TPB - 1) / S256_TPB); }
#define LAUNCH(ctx, kern, grid, smem, strm, ...)                  \
    do {                                                          \
        kern<<<(grid), S256_TPB, (smem), (strm)>>>(__VA_ARGS__);  \
        (ctx)->launches.fetch_add(1, std::memory_order_relaxed);  \
    } while (0)

struct dev_guard {
    int prev = -1;
    explicit dev_guard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~dev_guard() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

extern "C" const char *s256_strerror(int code) {
    switch (code) {
        case S256_SUCCESS: return "success";
        case S256_ERR_NO_DEVICE: return "no usable CUDA device (sm_100 required; there is no CPU fallback)";
        case S256_ERR_CUDA: return "CUDA runtime error";
        case S256_ERR_ARG: return "invalid argument";
        case S256_ERR_NOMEM: return "out of memory";
        case S256_ERR_UNIMPLEMENTED: return "not implemented";
        case S256_ERR_NCCL: return "NCCL error";
        default: return "unknown error";
    }
}
extern "C" const char *s256_last_cuda_error(const s256_ctx *ctx) { return ctx ? ctx->last_err.c_str() : ""; }
extern "C" int s256_device(const s256_ctx *ctx) { return ctx ? ctx->device : -1; }
extern "C" uint64_t s256_launch_count(const s256_ctx *ctx) { return ctx ? ctx->launches.load() : 0; }

extern "C" void s256_free(s256_ctx *ctx) {
    if (!ctx) return;
    {
        dev_guard g(ctx->device);
        void *ptrs[] = {ctx->comb, ctx->ct_tab, ctx->aff, ctx->u1,   ctx->dig1, ctx->dig2, ctx->sfl, ctx->pvalid,
                        ctx->cstat, ctx->tbl,   ctx->res, ctx->in_a, ctx->in_b, ctx->in_c, ctx->out, ctx->st,
                        ctx->sink, ctx->msm_counts, ctx->msm_offsets, ctx->msm_cursor, ctx->msm_entries, ctx->msm_flag,
                        ctx->msm_buckets, ctx->msm_win, ctx->msm_acc, ctx->msm_tmp, ctx->msm_cub, ctx->msm_nsl, ctx->msm_sloff};
        for (void *p : ptrs)
            if (p) cudaFree(p);
        if (ctx->stream) cudaStreamDestroy(ctx->stream);
        if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
        if (ctx->ev_decode) cudaEventDestroy(ctx->ev_decode);
    }
    delete ctx;
}

static int ctx_alloc(s256_ctx *ctx) {
    size_t cap = ctx->cap;
    CK(cudaMalloc(&ctx->comb, sizeof(apt) * COMB_NW * COMB_SZ));
    CK(cudaMalloc(&ctx->ct_tab, sizeof(apt) * CT_NW * CT_SZ));
    CK(cudaMalloc(&ctx->aff, sizeof(apt) * cap));
    CK(cudaMalloc(&ctx->u1, sizeof(sc) * cap));
    CK(cudaMalloc(&ctx->dig1, (size_t)DSM_ND * cap));
    CK(cudaMalloc(&ctx->dig2, (size_t)DSM_ND * cap));
    CK(cudaMalloc(&ctx->sfl, cap));
    CK(cudaMalloc(&ctx->pvalid, cap));
    CK(cudaMalloc(&ctx->cstat, cap));
    CK(cudaMalloc(&ctx->tbl, sizeof(pt) * DSM_TS * cap));
    CK(cudaMalloc(&ctx->res, sizeof(pt) * cap));
    CK(cudaMalloc(&ctx->in_a, 65 * cap));
    CK(cudaMalloc(&ctx->in_c, 65 * cap));
    CK(cudaMalloc(&ctx->in_b, 32 * cap));
    ctx->in_b_bytes = 32 * cap;
    CK(cudaMalloc(&ctx->out, 65 * cap));
    CK(cudaMalloc(&ctx->st, cap));
    CK(cudaMalloc(&ctx->sink, 8));
    return S256_SUCCESS;
}

extern "C" int s256_init(s256_ctx **out, int device, size_t max_batch) {
    if (!out) return S256_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return S256_ERR_NO_DEVICE;
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return S256_ERR_NO_DEVICE;
    if (device >= count) return S256_ERR_ARG;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return S256_ERR_NO_DEVICE;
    if (prop.major != 10) return S256_ERR_NO_DEVICE;  // sm_100a code only
    s256_ctx *ctx = new (std::nothrow) s256_ctx;
    if (!ctx) return S256_ERR_NOMEM;
    ctx->device = device;
    ctx->cap = max_batch ? max_batch : ((size_t)1 << 20);
    dev_guard g(device);
    int rc = ctx_alloc(ctx);
    if (rc == S256_SUCCESS && (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
                               cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess))
        rc = S256_ERR_CUDA;
    if (rc == S256_SUCCESS && cudaEventCreateWithFlags(&ctx->ev_decode, cudaEventDisableTiming) != cudaSuccess)
        rc = S256_ERR_CUDA;
    if (rc == S256_SUCCESS) {
        // generator tables (reference: package init, point_mul_table.go:75-100,147-160)
        size_t total = (size_t)COMB_NW * COMB_SZ;
        LAUNCH(ctx, k_gen_table, grid_for(total), 0, ctx->stream, ctx->comb, COMB_WB, total);
        LAUNCH(ctx, k_gen_ct_table, grid_for((size_t)CT_NW * CT_SZ), 0, ctx->stream, ctx->ct_tab);
        s256_ct_kernels_init();
        if (CT_SMEM_BYTES)
            cudaFuncSetAttribute(k_scalar_mult_ct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CT_SMEM_BYTES);
        cudaFuncSetAttribute(k_dsm_vm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VM_SMEM_BYTES);
        const char *pp = getenv("S256_PIPE_PARTS");
        if (pp && atoi(pp) >= 1 && atoi(pp) <= 16) ctx->pipe_parts = atoi(pp);
        const char *lad = getenv("S256_LADDER");
        ctx->use_reg_ladder = !(lad && std::string(lad) == "vm");  // register form is the faster one (profiles/)
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            fprintf(stderr, "s256_init: table generation failed: %s\n", cudaGetErrorString(e));
            rc = S256_ERR_CUDA;
        }
    }
    if (rc != S256_SUCCESS) {
        s256_free(ctx);
        return rc;
    }
    *out = ctx;
    return S256_SUCCESS;
}

// ---------------------------------------------------------------------------
// device-pointer pipelines (one chunk <= cap)
// ---------------------------------------------------------------------------
// A window into the per-item scratch arrays starting at item `off`: sub-chunks of one call work on
// disjoint windows, so they can be in flight on different streams at the same time.
struct view {
    apt *aff;
    sc *u1;
    int8_t *dig1, *dig2;
    uint8_t *sfl, *pvalid, *cstat;
    pt *tbl, *res;
    uint8_t *in_a, *in_b, *in_c, *out, *st;
};
static view view_at(const s256_ctx *ctx, size_t off) {
    view v;
    v.aff = ctx->aff + off;
    v.u1 = ctx->u1 + off;
    v.dig1 = ctx->dig1 + (size_t)DSM_ND * off;  // [digit][item] inside the window
    v.dig2 = ctx->dig2 + (size_t)DSM_ND * off;
    v.sfl = ctx->sfl + off;
    v.pvalid = ctx->pvalid + off;
    v.cstat = ctx->cstat + off;
    v.tbl = ctx->tbl + (size_t)DSM_TS * off;
    v.res = ctx->res + off;
    v.in_a = ctx->in_a + 65 * off;
    v.in_b = ctx->in_b + 32 * off;
    v.in_c = ctx->in_c + 65 * off;
    v.out = ctx->out + 65 * off;
    v.st = ctx->st + off;
    return v;
}

static void enqueue_dsm(s256_ctx *ctx, const view &v, size_t n, cudaStream_t s) {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ctx->profiling) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, s);
    }
    if (ctx->use_reg_ladder)
        LAUNCH(ctx, k_dsm, grid_for(n), 0, s, n, v.aff, v.u1, v.dig1, v.dig2, v.sfl, v.tbl, v.res,
               ctx->comb);
    else
        LAUNCH(ctx, k_dsm_vm, grid_for(n), VM_SMEM_BYTES, s, n, v.aff, v.u1, v.dig1, v.dig2, v.sfl,
               v.tbl, v.res, ctx->comb);
    if (ctx->profiling) {
        cudaEventRecord(e1, s);
        ctx->dsm_events.emplace_back(e0, e1);
    }
}
#ifndef S256_INV_K
#define S256_INV_K 32
#endif
constexpr int INV_K = S256_INV_K;
// Inversion group size by batch size: K items share one Fermat chain but are processed serially by
// one thread, so small batches use small groups (n = 4096 with K = 32 would run on 128 threads).
static inline int inv_k_for(size_t n) { return n >= ((size_t)1 << 19) ? INV_K : (n >= ((size_t)1 << 16) ? 4 : 1); }
#define DISPATCH_K(n, CALL)               \
    do {                                  \
        switch (inv_k_for(n)) {           \
            case 1: { constexpr int KK = 1; CALL; } break;  \
            case 4: { constexpr int KK = 4; CALL; } break;  \
            default: { constexpr int KK = INV_K; CALL; } break; \
        }                                 \
    } while (0)
constexpr int MSM_MAX_PARTS = 16;  // (2^16 buckets) / (128 threads * 32 buckets)
static inline unsigned grid_for_groups(size_t n, int k) { return grid_for((n + k - 1) / k); }

// The scalar kernel needs digest + signature only and the decode kernel the public keys only: when the
// caller passes a second stream, decode runs there (behind the key copy) and joins through an event.
static int chunk_ecdsa_verify(s256_ctx *ctx, const view &v, const uint8_t *pk, const uint8_t *dg, const uint8_t *sig, uint32_t flags,
                              size_t n, uint8_t *ok, cudaStream_t s, cudaStream_t s_decode = nullptr) {
    DISPATCH_K(n, LAUNCH(ctx, (k_ecdsa_scalars<KK, false>), grid_for_groups(n, KK), 0, s, dg, sig, n, flags, v.u1,
                         v.dig1, v.dig2, v.sfl));
    if (s_decode && s_decode != s) {
        LAUNCH(ctx, k_decode_uncompressed, grid_for(n), 0, s_decode, pk, n, v.aff, v.pvalid);
        cudaEventRecord(ctx->ev_decode, s_decode);
        cudaStreamWaitEvent(s, ctx->ev_decode, 0);
    } else {
        LAUNCH(ctx, k_decode_uncompressed, grid_for(n), 0, s, pk, n, v.aff, v.pvalid);
    }
    enqueue_dsm(ctx, v, n, s);
    LAUNCH(ctx, k_ecdsa_finish, grid_for(n), 0, s, n, v.res, sig, v.pvalid, v.sfl, ok);
    return S256_SUCCESS;
}
static int chunk_ecdsa_recover(s256_ctx *ctx, const view &v, const uint8_t *dg, const uint8_t *sig65, size_t n, uint8_t *pk65,
                               uint8_t *status, cudaStream_t s) {
    LAUNCH(ctx, k_decode_recover, grid_for(n), 0, s, sig65, n, v.aff, v.pvalid);
    DISPATCH_K(n, LAUNCH(ctx, (k_ecdsa_scalars<KK, true>), grid_for_groups(n, KK), 0, s, dg, sig65, n, 0u, v.u1,
                         v.dig1, v.dig2, v.sfl));
    enqueue_dsm(ctx, v, n, s);
    DISPATCH_K(n, LAUNCH(ctx, k_finish_affine<KK>, grid_for_groups(n, KK), 0, s, n, v.res, v.pvalid, v.sfl, v.cstat,
           3, pk65, status, (const uint8_t *)nullptr));
    return S256_SUCCESS;
}
static int chunk_schnorr_verify(s256_ctx *ctx, const view &v, const uint8_t *pkx, const uint8_t *msg, size_t msg_len,
                                const uint8_t *sig, size_t n, uint8_t *ok, cudaStream_t s) {
    LAUNCH(ctx, k_decode_xonly, grid_for(n), 0, s, pkx, n, v.aff, v.pvalid);
    LAUNCH(ctx, k_schnorr_scalars, grid_for(n), 0, s, pkx, msg, msg_len, sig, n, v.u1, v.dig1, v.dig2,
           v.sfl);
    enqueue_dsm(ctx, v, n, s);
    DISPATCH_K(n, LAUNCH(ctx, k_finish_affine<KK>, grid_for_groups(n, KK), 0, s, n, v.res, v.pvalid, v.sfl, v.cstat,
           2, (uint8_t *)nullptr, ok, sig));
    return S256_SUCCESS;
}
static int chunk_dsm(s256_ctx *ctx, const view &v, const uint8_t *u1, const uint8_t *u2, const uint8_t *pt65, size_t n,
                     uint8_t *out65, uint8_t *status, cudaStream_t s) {
    LAUNCH(ctx, k_decode_uncompressed, grid_for(n), 0, s, pt65, n, v.aff, v.pvalid);
    LAUNCH(ctx, k_plain_scalars, grid_for(n), 0, s, u1, u2, n, v.u1, v.dig1, v.dig2, v.sfl);
    enqueue_dsm(ctx, v, n, s);
    DISPATCH_K(n, LAUNCH(ctx, k_finish_affine<KK>, grid_for_groups(n, KK), 0, s, n, v.res, v.pvalid,
           (const uint8_t *)nullptr, v.cstat, 0, out65, status, (const uint8_t *)nullptr));
    return S256_SUCCESS;
}
static int chunk_base_mult(s256_ctx *ctx, const view &v, const uint8_t *k32, size_t n, uint8_t *out65, uint8_t *status,
                           cudaStream_t s) {
    s256_launch_base_mult_ct(k32, n, ctx->ct_tab, v.res, s);
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    DISPATCH_K(n, LAUNCH(ctx, k_finish_affine<KK>, grid_for_groups(n, KK), 0, s, n, v.res, (const uint8_t *)nullptr,
           (const uint8_t *)nullptr, v.cstat, 0, out65, status, (const uint8_t *)nullptr));
    return S256_SUCCESS;
}

// Point.ScalarMult / PrivateKey.ECDH: decode (public) -> ct ladder -> batched affine
static int chunk_scalar_mult(s256_ctx *ctx, const view &v, const uint8_t *k32, const uint8_t *pt65, size_t n, int mode, uint8_t *out,
                             uint8_t *status, cudaStream_t s) {
    LAUNCH(ctx, k_decode_uncompressed, grid_for(n), 0, s, pt65, n, v.aff, v.pvalid);
    s256_launch_scalar_mult_ct(n, v.aff, k32, v.tbl, v.res, s);
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    DISPATCH_K(n, LAUNCH(ctx, k_finish_affine<KK>, grid_for_groups(n, KK), 0, s, n, v.res, v.pvalid,
           (const uint8_t *)nullptr, v.cstat, mode, out, status, (const uint8_t *)nullptr));
    return S256_SUCCESS;
}

// PrivateKey.Sign(RFC6979SHA256(), digest): nonce -> k*G (ct) -> affine -> (r, s, v).  Scratch use: k in
// v.u1 (32 B/item), R in v.out (65 B/item), validity in v.pvalid; the nonce buffer is wiped afterwards.
static int chunk_sign(s256_ctx *ctx, const view &v, const uint8_t *priv32, const uint8_t *digest32, size_t n,
                      uint8_t *sig64, uint8_t *recid, uint8_t *status, cudaStream_t s) {
    uint8_t *kbuf = reinterpret_cast<uint8_t *>(v.u1);
    LAUNCH(ctx, k_rfc6979_nonce, grid_for(n), 0, s, priv32, digest32, n, kbuf, v.pvalid);
    s256_launch_base_mult_ct(kbuf, n, ctx->ct_tab, v.res, s);
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    DISPATCH_K(n, LAUNCH(ctx, k_finish_affine<KK>, grid_for_groups(n, KK), 0, s, n, v.res, (const uint8_t *)nullptr,
                         (const uint8_t *)nullptr, v.cstat, 0, v.out, v.sfl, (const uint8_t *)nullptr));
    DISPATCH_K(n, LAUNCH(ctx, k_sign_finish<KK>, grid_for_groups(n, KK), 0, s, n, priv32, digest32, kbuf, v.pvalid, v.out,
                         sig64, recid, status));
    CK(cudaMemsetAsync(kbuf, 0, 32 * n, s));
    CK(cudaMemsetAsync(v.res, 0, sizeof(pt) * n, s));
    return S256_SUCCESS;
}

// SchnorrPrivateKey.Sign: P = d'G -> nonce -> R = k'G -> finish.  Byte scratch: P in v.out, R and k' in the
// per-item table area (1536 B/item, unused by this path); both secret buffers are wiped afterwards.
static int chunk_schnorr_sign(s256_ctx *ctx, const view &v, const uint8_t *priv32, const uint8_t *msg, size_t msg_len,
                              const uint8_t *aux32, size_t n, uint8_t *sig64, uint8_t *status, cudaStream_t s) {
    uint8_t *arena = reinterpret_cast<uint8_t *>(v.tbl);
    uint8_t *r65 = arena, *kbuf = arena + 65 * n;
    s256_launch_base_mult_ct(priv32, n, ctx->ct_tab, v.res, s);
    DISPATCH_K(n, LAUNCH(ctx, k_finish_affine<KK>, grid_for_groups(n, KK), 0, s, n, v.res, (const uint8_t *)nullptr,
                         (const uint8_t *)nullptr, v.cstat, 0, v.out, v.sfl, (const uint8_t *)nullptr));
    LAUNCH(ctx, k_schnorr_nonce, grid_for(n), 0, s, priv32, v.out, msg, msg_len, aux32, n, kbuf, v.pvalid);
    s256_launch_base_mult_ct(kbuf, n, ctx->ct_tab, v.res, s);
    ctx->launches.fetch_add(2, std::memory_order_relaxed);
    DISPATCH_K(n, LAUNCH(ctx, k_finish_affine<KK>, grid_for_groups(n, KK), 0, s, n, v.res, (const uint8_t *)nullptr,
                         (const uint8_t *)nullptr, v.cstat, 0, r65, v.sfl, (const uint8_t *)nullptr));
    LAUNCH(ctx, k_schnorr_sign_finish, grid_for(n), 0, s, priv32, v.out, r65, kbuf, msg, msg_len, v.pvalid, n, sig64,
           status);
    CK(cudaMemsetAsync(kbuf, 0, 32 * n, s));
    CK(cudaMemsetAsync(v.res, 0, sizeof(pt) * n, s));
    return S256_SUCCESS;
}

// Runs `body(offset, count)` over chunks of at most cap items.
template <typename F>
static int for_chunks(s256_ctx *ctx, size_t n, F body) {
    for (size_t off = 0; off < n; off += ctx->cap) {
        size_t c = n - off < ctx->cap ? n - off : ctx->cap;
        int rc = body(off, c);
        if (rc != S256_SUCCESS) return rc;
    }
    return S256_SUCCESS;
}
// Host-pointer calls: the chunk is cut into sub-chunks that alternate between two streams, each
// doing its own H2D -> kernels -> D2H on a disjoint scratch window, so the copies of one sub-chunk
// overlap the kernels of the other.  body(view, global offset, count, stream).
template <typename F>
static int pipelined(s256_ctx *ctx, size_t n, F body) {
    const size_t min_sub = 65536;
    for (size_t off = 0; off < n; off += ctx->cap) {
        size_t c = n - off < ctx->cap ? n - off : ctx->cap;
        size_t parts = c / min_sub;
        if (parts > (size_t)ctx->pipe_parts) parts = (size_t)ctx->pipe_parts;
        if (parts < 1) parts = 1;
        size_t sub = (c + parts - 1) / parts;
        sub = (sub + 127) & ~(size_t)127;
        int k = 0;
        for (size_t so = 0; so < c; so += sub, k++) {
            size_t sc_ = c - so < sub ? c - so : sub;
            int rc = body(view_at(ctx, so), off + so, sc_, (k & 1) ? ctx->stream2 : ctx->stream);
            if (rc != S256_SUCCESS) return rc;
        }
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream2));
    }
    return S256_SUCCESS;
}
static int check_launch(s256_ctx *ctx) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        ctx->last_err = std::string("kernel launch: ") + cudaGetErrorString(e);
        return S256_ERR_CUDA;
    }
    return S256_SUCCESS;
}

// ---------------------------------------------------------------------------
// exported entry points
// ---------------------------------------------------------------------------
#define ENTER(ctx)                       \
    if (!(ctx)) return S256_ERR_ARG;     \
    std::lock_guard<std::mutex> lk_((ctx)->mu); \
    dev_guard dg_((ctx)->device)

extern "C" int s256_ecdsa_verify_dev(s256_ctx *ctx, const uint8_t *pk, const uint8_t *dg, const uint8_t *sig,
                                     uint32_t flags, size_t n, uint8_t *ok, void *stream) {
    ENTER(ctx);
    if (n && (!pk || !dg || !sig || !ok)) return S256_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = for_chunks(ctx, n, [&](size_t off, size_t c) {
        return chunk_ecdsa_verify(ctx, view_at(ctx, 0), pk + 65 * off, dg + 32 * off, sig + 64 * off, flags, c, ok + off, s);
    });
    return rc != S256_SUCCESS ? rc : check_launch(ctx);
}
extern "C" int s256_ecdsa_verify(s256_ctx *ctx, const uint8_t *pk, const uint8_t *dg, const uint8_t *sig,
                                 uint32_t flags, size_t n, uint8_t *ok) {
    ENTER(ctx);
    if (n && (!pk || !dg || !sig || !ok)) return S256_ERR_ARG;
    int rc = pipelined(ctx, n, [&](const view &v, size_t off, size_t c, cudaStream_t s) {
        // digest + signature first: the batched inversion starts while the keys are still in flight
        cudaStream_t s2 = (s == ctx->stream) ? ctx->stream2 : ctx->stream;
        CK(cudaMemcpyAsync(v.in_b, dg + 32 * off, 32 * c, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(v.in_c, sig + 64 * off, 64 * c, cudaMemcpyHostToDevice, s));
        bool split = ctx->pipe_parts == 1;
        CK(cudaMemcpyAsync(v.in_a, pk + 65 * off, 65 * c, cudaMemcpyHostToDevice, split ? s2 : s));
        int r = chunk_ecdsa_verify(ctx, v, v.in_a, v.in_b, v.in_c, flags, c, v.st, s, split ? s2 : s);
        if (r != S256_SUCCESS) return r;
        CK(cudaMemcpyAsync(ok + off, v.st, c, cudaMemcpyDeviceToHost, s));
        return S256_SUCCESS;
    });
    return rc != S256_SUCCESS ? rc : check_launch(ctx);
}

extern "C" int s256_ecdsa_recover_dev(s256_ctx *ctx, const uint8_t *dg, const uint8_t *sig65, size_t n, uint8_t *pk65,
                                      uint8_t *status, void *stream) {
    ENTER(ctx);
    if (n && (!dg || !sig65 || !pk65 || !status)) return S256_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = for_chunks(ctx, n, [&](size_t off, size_t c) {
        return chunk_ecdsa_recover(ctx, view_at(ctx, 0), dg + 32 * off, sig65 + 65 * off, c, pk65 + 65 * off, status + off, s);
    });
    return rc != S256_SUCCESS ? rc : check_launch(ctx);
}
extern "C" int s256_ecdsa_recover(s256_ctx *ctx, const uint8_t *dg, const uint8_t *sig65, size_t n, uint8_t *pk65,
                                  uint8_t *status) {
    ENTER(ctx);
    if (n && (!dg || !sig65 || !pk65 || !status)) return S256_ERR_ARG;
    int rc = pipelined(ctx, n, [&](const view &v, size_t off, size_t c, cudaStream_t s) {
        CK(cudaMemcpyAsync(v.in_b, dg + 32 * off, 32 * c, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(v.in_c, sig65 + 65 * off, 65 * c, cudaMemcpyHostToDevice, s));
        int r = chunk_ecdsa_recover(ctx, v, v.in_b, v.in_c, c, v.out, v.st, s);
        if (r != S256_SUCCESS) return r;
        CK(cudaMemcpyAsync(pk65 + 65 * off, v.out, 65 * c, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(status + off, v.st, c, cudaMemcpyDeviceToHost, s));
        return S256_SUCCESS;
    });
    return rc != S256_SUCCESS ? rc : check_launch(ctx);
}

extern "C" int s256_schnorr_verify_dev(s256_ctx *ctx, const uint8_t *pkx, const uint8_t *msg, size_t msg_len,
                                       const uint8_t *sig, size_t n, uint8_t *ok, void *stream) {
    ENTER(ctx);
    if (n && (!pkx || (!msg && msg_len) || !sig || !ok)) return S256_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = for_chunks(ctx, n, [&](size_t off, size_t c) {
        return chunk_schnorr_verify(ctx, view_at(ctx, 0), pkx + 32 * off, msg + msg_len * off, msg_len, sig + 64 * off, c, ok + off, s);
    });
    return rc != S256_SUCCESS ? rc : check_launch(ctx);
}
extern "C" int s256_schnorr_verify(s256_ctx *ctx, const uint8_t *pkx, const uint8_t *msg, size_t msg_len,
                                   const uint8_t *sig, size_t n, uint8_t *ok) {
    ENTER(ctx);
    if (n && (!pkx || (!msg && msg_len) || !sig || !ok)) return S256_ERR_ARG;
    cudaStream_t s = ctx->stream;
    size_t need = (msg_len ? msg_len : 1) * (n < ctx->cap ? n : ctx->cap);
    if (need > ctx->in_b_bytes) {
        if (ctx->in_b) cudaFree(ctx->in_b);
        ctx->in_b = nullptr;
        ctx->in_b_bytes = 0;
        CK(cudaMalloc(&ctx->in_b, need));
        ctx->in_b_bytes = need;
    }
    int rc = for_chunks(ctx, n, [&](size_t off, size_t c) {
        CK(cudaMemcpyAsync(ctx->in_a, pkx + 32 * off, 32 * c, cudaMemcpyHostToDevice, s));
        if (msg_len) CK(cudaMemcpyAsync(ctx->in_b, msg + msg_len * off, msg_len * c, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->in_c, sig + 64 * off, 64 * c, cudaMemcpyHostToDevice, s));
        int r = chunk_schnorr_verify(ctx, view_at(ctx, 0), ctx->in_a, ctx->in_b, msg_len, ctx->in_c, c, ctx->st, s);
        if (r != S256_SUCCESS) return r;
        CK(cudaMemcpyAsync(ok + off, ctx->st, c, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        return S256_SUCCESS;
    });
    return rc != S256_SUCCESS ? rc : check_launch(ctx);
}

extern "C" int s256_double_scalar_mult_basepoint_vartime_dev(s256_ctx *ctx, const uint8_t *u1, const uint8_t *u2,
                                                             const uint8_t *pt65, size_t n, uint8_t *out65,
                                                             uint8_t *status, void *stream) {
    ENTER(ctx);
    if (n && (!u1 || !u2 || !pt65 || !out65 || !status)) return S256_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = for_chunks(ctx, n, [&](size_t off, size_t c) {
        return chunk_dsm(ctx, view_at(ctx, 0), u1 + 32 * off, u2 + 32 * off, pt65 + 65 * off, c, out65 + 65 * off, status + off, s);
    });
    return rc != S256_SUCCESS ? rc : check_launch(ctx);
}
extern "C" int s256_double_scalar_mult_basepoint_vartime(s256_ctx *ctx, const uint8_t *u1, const uint8_t *u2,
                                                         const uint8_t *pt65, size_t n, uint8_t *out65,
                                                         uint8_t *status) {
    ENTER(ctx);
    if (n && (!u1 || !u2 || !pt65 || !out65 || !status)) return S256_ERR_ARG;
    int rc = pipelined(ctx, n, [&](const view &v, size_t off, size_t c, cudaStream_t s) {
        CK(cudaMemcpyAsync(v.in_a, pt65 + 65 * off, 65 * c, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(v.in_b, u1 + 32 * off, 32 * c, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(v.in_c, u2 + 32 * off, 32 * c, cudaMemcpyHostToDevice, s));
        int r = chunk_dsm(ctx, v, v.in_b, v.in_c, v.in_a, c, v.out, v.st, s);
        if (r != S256_SUCCESS) return r;
        CK(cudaMemcpyAsync(out65 + 65 * off, v.out, 65 * c, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(status + off, v.st, c, cudaMemcpyDeviceToHost, s));
        return S256_SUCCESS;
    });
    return rc != S256_SUCCESS ? rc : check_launch(ctx);
}

extern "C" int s256_scalar_base_mult_dev(s256_ctx *ctx, const uint8_t *k32, size_t n, uint8_t *out65, uint8_t *status,
                                         void *stream) {
    ENTER(ctx);
    if (n && (!k32 || !out65 || !status)) return S256_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = for_chunks(ctx, n, [&](size_t off, size_t c) {
        return chunk_base_mult(ctx, view_at(ctx, 0), k32 + 32 * off, c, out65 + 65 * off, status + off, s);
    });
    return rc != S256_SUCCESS ? rc : check_launch(ctx);
}
extern "C" int s256_scalar_base_mult(s256_ctx *ctx, const uint8_t *k32, size_t n, uint8_t *out65, uint8_t *status) {
    ENTER(ctx);
    if (n && (!k32 || !out65 || !status)) return S256_ERR_ARG;
    int rc = pipelined(ctx, n, [&](const view &v, size_t off, size_t c, cudaStream_t s) {
        CK(cudaMemcpyAsync(v.in_b, k32 + 32 * off, 32 * c, cudaMemcpyHostToDevice, s));
        int r = chunk_base_mult(ctx, v, v.in_b, c, v.out, v.st, s);
        if (r != S256_SUCCESS) return r;
        CK(cudaMemcpyAsync(out65 + 65 * off, v.out, 65 * c, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(status + off, v.st, c, cudaMemcpyDeviceToHost, s));
        return S256_SUCCESS;
    });
    return rc != S256_SUCCESS ? rc : check_launch(ctx);
}

static int scalar_mult_common_dev(s256_ctx *ctx, const uint8_t *k32, const uint8_t *pt65, size_t n, int mode,
                                  uint8_t *out, uint8_t *status, void *stream) {
    ENTER(ctx);
    if (n && (!k32 || !pt65 || !out || !status)) return S256_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    size_t w = mode == 1 ? 32 : 65;
    int rc = for_chunks(ctx, n, [&](size_t off, size_t c) {
        return chunk_scalar_mult(ctx, view_at(ctx, 0), k32 + 32 * off, pt65 + 65 * off, c, mode, out + w * off, status + off, s);
    });
    return rc != S256_SUCCESS ? rc : check_launch(ctx);
}
static int scalar_mult_common_host(s256_ctx *ctx, const uint8_t *k32, const uint8_t *pt65, size_t n, int mode,
                                   uint8_t *out, uint8_t *status) {
    ENTER(ctx);
    if (n && (!k32 || !pt65 || !out || !status)) return S256_ERR_ARG;
    size_t w = mode == 1 ? 32 : 65;
    int rc = pipelined(ctx, n, [&](const view &v, size_t off, size_t c, cudaStream_t s) {
        CK(cudaMemcpyAsync(v.in_a, pt65 + 65 * off, 65 * c, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(v.in_b, k32 + 32 * off, 32 * c, cudaMemcpyHostToDevice, s));
        int r = chunk_scalar_mult(ctx, v, v.in_b, v.in_a, c, mode, v.out, v.st, s);
        if (r != S256_SUCCESS) return r;
        CK(cudaMemcpyAsync(out + w * off, v.out, w * c, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(status + off, v.st, c, cudaMemcpyDeviceToHost, s));
        return S256_SUCCESS;
    });
    return rc != S256_SUCCESS ? rc : check_launch(ctx);
}
extern "C" int s256_scalar_mult(s256_ctx *ctx, const uint8_t *k32, const uint8_t *pt65, size_t n, uint8_t *out65,
                                uint8_t *status) {
    return scalar_mult_common_host(ctx, k32, pt65, n, 0, out65, status);
}
extern "C" int s256_scalar_mult_dev(s256_ctx *ctx, const uint8_t *k32, const uint8_t *pt65, size_t n, uint8_t *out65,
                                    uint8_t *status, void *stream) {
    return scalar_mult_common_dev(ctx, k32, pt65, n, 0, out65, status, stream);
}
extern "C" int s256_ecdh(s256_ctx *ctx, const uint8_t *k32, const uint8_t *pt65, size_t n, uint8_t *x32,
                         uint8_t *status) {
    return scalar_mult_common_host(ctx, k32, pt65, n, 1, x32, status);
}
extern "C" int s256_ecdh_dev(s256_ctx *ctx, const uint8_t *k32, const uint8_t *pt65, size_t n, uint8_t *x32,
                             uint8_t *status, void *stream) {
    return scalar_mult_common_dev(ctx, k32, pt65, n, 1, x32, status, stream);
}
// NewPointFromBytes on compressed encodings (point_s11n.go:140): 33 B -> 65 B + status
extern "C" int s256_point_decompress(s256_ctx *ctx, const uint8_t *pt33, size_t n, uint8_t *out65, uint8_t *status) {
    ENTER(ctx);
    if (n && (!pt33 || !out65 || !status)) return S256_ERR_ARG;
    cudaStream_t s = ctx->stream;
    int rc = for_chunks(ctx, n, [&](size_t off, size_t c) {
        CK(cudaMemcpyAsync(ctx->in_a, pt33 + 33 * off, 33 * c, cudaMemcpyHostToDevice, s));
        LAUNCH(ctx, k_decode_compressed, grid_for(c), 0, s, ctx->in_a, c, ctx->aff, ctx->pvalid);
        LAUNCH(ctx, k_encode_affine, grid_for(c), 0, s, ctx->aff, ctx->pvalid, c, ctx->out, ctx->st);
        CK(cudaMemcpyAsync(out65 + 65 * off, ctx->out, 65 * c, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(status + off, ctx->st, c, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        return S256_SUCCESS;
    });
    return rc != S256_SUCCESS ? rc : check_launch(ctx);
}

extern "C" int s256_ecdsa_sign_rfc6979_dev(s256_ctx *ctx, const uint8_t *priv32, const uint8_t *digest32, size_t n,
                                           uint8_t *sig64, uint8_t *recid, uint8_t *status, void *stream) {
    ENTER(ctx);
    if (n && (!priv32 || !digest32 || !sig64 || !recid || !status)) return S256_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = for_chunks(ctx, n, [&](size_t off, size_t c) {
        return chunk_sign(ctx, view_at(ctx, 0), priv32 + 32 * off, digest32 + 32 * off, c, sig64 + 64 * off, recid + off,
                          status + off, s);
    });
    return rc != S256_SUCCESS ? rc : check_launch(ctx);
}
extern "C" int s256_ecdsa_sign_rfc6979(s256_ctx *ctx, const uint8_t *priv32, const uint8_t *digest32, size_t n,
                                       uint8_t *sig64, uint8_t *recid, uint8_t *status) {
    ENTER(ctx);
    if (n && (!priv32 || !digest32 || !sig64 || !recid || !status)) return S256_ERR_ARG;
    int rc = pipelined(ctx, n, [&](const view &v, size_t off, size_t c, cudaStream_t s) {
        CK(cudaMemcpyAsync(v.in_a, priv32 + 32 * off, 32 * c, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(v.in_b, digest32 + 32 * off, 32 * c, cudaMemcpyHostToDevice, s));
        // outputs staged in v.in_c (sig64), v.cstat is busy inside finish_affine -> recid in v.in_a + 32*cap? use tail of in_c
        uint8_t *d_sig = v.in_c, *d_rec = v.in_c + 64 * c, *d_st = v.st;
        int r = chunk_sign(ctx, v, v.in_a, v.in_b, c, d_sig, d_rec, d_st, s);
        if (r != S256_SUCCESS) return r;
        CK(cudaMemcpyAsync(sig64 + 64 * off, d_sig, 64 * c, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(recid + off, d_rec, c, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(status + off, d_st, c, cudaMemcpyDeviceToHost, s));
        CK(cudaMemsetAsync(v.in_a, 0, 32 * c, s));  // wipe the staged private keys
        return S256_SUCCESS;
    });
    return rc != S256_SUCCESS ? rc : check_launch(ctx);
}

extern "C" int s256_schnorr_sign_dev(s256_ctx *ctx, const uint8_t *priv32, const uint8_t *msg, size_t msg_len,
                                     const uint8_t *aux32, size_t n, uint8_t *sig64, uint8_t *status, void *stream) {
    ENTER(ctx);
    if (n && (!priv32 || (!msg && msg_len) || !aux32 || !sig64 || !status)) return S256_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = for_chunks(ctx, n, [&](size_t off, size_t c) {
        return chunk_schnorr_sign(ctx, view_at(ctx, 0), priv32 + 32 * off, msg + msg_len * off, msg_len, aux32 + 32 * off,
                                  c, sig64 + 64 * off, status + off, s);
    });
    return rc != S256_SUCCESS ? rc : check_launch(ctx);
}
extern "C" int s256_schnorr_sign(s256_ctx *ctx, const uint8_t *priv32, const uint8_t *msg, size_t msg_len,
                                 const uint8_t *aux32, size_t n, uint8_t *sig64, uint8_t *status) {
    ENTER(ctx);
    if (n && (!priv32 || (!msg && msg_len) || !aux32 || !sig64 || !status)) return S256_ERR_ARG;
    cudaStream_t s = ctx->stream;
    size_t need = (msg_len ? msg_len : 1) * (n < ctx->cap ? n : ctx->cap);
    if (need > ctx->in_b_bytes) {
        if (ctx->in_b) cudaFree(ctx->in_b);
        ctx->in_b = nullptr;
        ctx->in_b_bytes = 0;
        CK(cudaMalloc(&ctx->in_b, need));
        ctx->in_b_bytes = need;
    }
    int rc = for_chunks(ctx, n, [&](size_t off, size_t c) {
        view v = view_at(ctx, 0);
        CK(cudaMemcpyAsync(ctx->in_a, priv32 + 32 * off, 32 * c, cudaMemcpyHostToDevice, s));
        if (msg_len) CK(cudaMemcpyAsync(ctx->in_b, msg + msg_len * off, msg_len * c, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->in_c, aux32 + 32 * off, 32 * c, cudaMemcpyHostToDevice, s));
        uint8_t *d_sig = reinterpret_cast<uint8_t *>(v.aff);  // 64 B per item, unused by this path
        int r = chunk_schnorr_sign(ctx, v, ctx->in_a, ctx->in_b, msg_len, ctx->in_c, c, d_sig, ctx->st, s);
        if (r != S256_SUCCESS) return r;
        CK(cudaMemcpyAsync(sig64 + 64 * off, d_sig, 64 * c, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(status + off, ctx->st, c, cudaMemcpyDeviceToHost, s));
        CK(cudaMemsetAsync(ctx->in_a, 0, 32 * c, s));
        CK(cudaStreamSynchronize(s));
        return S256_SUCCESS;
    });
    return rc != S256_SUCCESS ? rc : check_launch(ctx);
}

// PublicKey.Verify with EncodingASN1 (secec/ecdsa.go:171-228): parse on the host (codecs.cpp), then the
// same batch as the compact path; rows the parser rejects come back false.
extern "C" int s256_ecdsa_verify_asn1(s256_ctx *ctx, const uint8_t *pk65, const uint8_t *digest32, const uint8_t *der,
                                      const size_t *offsets, uint32_t flags, size_t n, uint8_t *ok) {
    if (!ctx || (n && (!pk65 || !digest32 || !der || !offsets || !ok))) return S256_ERR_ARG;
    std::vector<uint8_t> sig(64 * n), parsed(n);
    int rc = s256_parse_asn1_signatures(der, offsets, n, sig.data(), parsed.data());
    if (rc != S256_SUCCESS) return rc;
    rc = s256_ecdsa_verify(ctx, pk65, digest32, sig.data(), flags, n, ok);
    if (rc != S256_SUCCESS) return rc;
    for (size_t i = 0; i < n; i++) ok[i] &= parsed[i];
    return S256_SUCCESS;
}
// bitcoin.VerifyASN1 (secec/bitcoin/ecdsa_shitcoin.go:29-35)
extern "C" int s256_bitcoin_verify_asn1(s256_ctx *ctx, const uint8_t *pk65, const uint8_t *digest32, const uint8_t *der,
                                        const size_t *offsets, size_t n, uint8_t *ok) {
    if (!ctx || (n && (!pk65 || !digest32 || !der || !offsets || !ok))) return S256_ERR_ARG;
    std::vector<uint8_t> bip(n);
    int rc = s256_is_valid_signature_encoding_bip0066(der, offsets, n, bip.data());
    if (rc != S256_SUCCESS) return rc;
    // strip the sighash byte of the rows that passed; rejected rows become empty (and fail to parse)
    std::vector<size_t> off2(n + 1);
    std::vector<uint8_t> der2;
    der2.reserve(offsets[n] - offsets[0]);
    for (size_t i = 0; i < n; i++) {
        off2[i] = der2.size();
        if (bip[i]) der2.insert(der2.end(), der + offsets[i], der + offsets[i + 1] - 1);
    }
    off2[n] = der2.size();
    if (der2.empty()) der2.push_back(0);
    rc = s256_ecdsa_verify_asn1(ctx, pk65, digest32, der2.data(), off2.data(), S256_FLAG_REJECT_MALLEABLE, n, ok);
    if (rc != S256_SUCCESS) return rc;
    for (size_t i = 0; i < n; i++) ok[i] &= bip[i];
    return S256_SUCCESS;
}

// ---------------------------------------------------------------------------
// MSM
// ---------------------------------------------------------------------------
static size_t msm_entries_capacity(const s256_ctx *ctx) {
    // nwin * n entries; the planner uses c >= 12 once n >= 2^16 (nwin <= 22), and c >= 4 (nwin <= 64) below
    size_t cap = ctx->cap;
    size_t small = (size_t)MSM_MAX_WIN * (cap < 65536 ? cap : 65536), large = (size_t)22 * cap;
    return small > large ? small : large;
}
static int msm_ensure(s256_ctx *ctx) {
    if (ctx->msm_cap) return S256_SUCCESS;
    size_t total = 0;
    for (int c = 4; c <= MSM_MAX_C; c++) {
        size_t t = (size_t)msm_plan_for_c(c).total;
        if (t > total) total = t;
    }
    CK(cudaMalloc(&ctx->msm_counts, (total + 1) * 4));
    CK(cudaMalloc(&ctx->msm_offsets, (total + 1) * 4));
    CK(cudaMalloc(&ctx->msm_cursor, (total + 1) * 4));
    CK(cudaMalloc(&ctx->msm_entries, msm_entries_capacity(ctx) * 4));
    // slice sums: one per bucket at least, plus one per MSM_SLICE entries
    ctx->msm_max_slices = total + msm_entries_capacity(ctx) / MSM_SLICE + 1;
    CK(cudaMalloc(&ctx->msm_buckets, ctx->msm_max_slices * sizeof(pt)));
    CK(cudaMalloc(&ctx->msm_nsl, (total + 1) * 4));
    CK(cudaMalloc(&ctx->msm_sloff, (total + 1) * 4));
    CK(cudaMalloc(&ctx->msm_win, (size_t)MSM_MAX_WIN * MSM_MAX_PARTS * sizeof(pt)));
    CK(cudaMalloc(&ctx->msm_acc, sizeof(pt)));
    CK(cudaMalloc(&ctx->msm_tmp, 4096 * sizeof(pt)));
    CK(cudaMalloc(&ctx->msm_flag, 4));
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, ctx->msm_counts, ctx->msm_offsets, (int)(total + 1));
    ctx->msm_cub_bytes = bytes;
    CK(cudaMalloc(&ctx->msm_cub, bytes));
    ctx->msm_cap = ctx->cap;
    return S256_SUCCESS;
}

// one chunk (device pointers): msm_acc (+)= sum k_i P_i
static int chunk_msm(s256_ctx *ctx, const uint8_t *k32, const uint8_t *pt65, size_t n, int vartime, int first,
                     cudaStream_t s) {
    LAUNCH(ctx, k_decode_uncompressed, grid_for(n), 0, s, pt65, n, ctx->aff, ctx->pvalid);
    LAUNCH(ctx, k_any_invalid, grid_for(n), 0, s, ctx->pvalid, n, ctx->msm_flag);
    if (!vartime || n < 32) {
        // constant-time flavour (and tiny inputs): ct ladder per item, then a sum tree
        s256_launch_scalar_mult_ct(n, ctx->aff, k32, ctx->tbl, ctx->res, s);
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
        size_t m = n < 2048 ? (n < 32 ? 1 : 32) : 2048;
        LAUNCH(ctx, k_reduce_points, grid_for(m), 0, s, ctx->res, n, ctx->msm_tmp, m);
        if (m > 32) {
            LAUNCH(ctx, k_reduce_points, grid_for(32), 0, s, ctx->msm_tmp, m, ctx->msm_tmp + 2048, (size_t)32);
            LAUNCH(ctx, k_reduce_points, 1, 0, s, ctx->msm_tmp + 2048, (size_t)32, ctx->msm_tmp + 2048 + 32, (size_t)1);
            k_acc_point<<<1, 1, 0, s>>>(ctx->msm_tmp + 2048 + 32, ctx->msm_acc, first);
        } else if (m > 1) {
            LAUNCH(ctx, k_reduce_points, 1, 0, s, ctx->msm_tmp, m, ctx->msm_tmp + 2048, (size_t)1);
            k_acc_point<<<1, 1, 0, s>>>(ctx->msm_tmp + 2048, ctx->msm_acc, first);
        } else {
            k_acc_point<<<1, 1, 0, s>>>(ctx->msm_tmp, ctx->msm_acc, first);
        }
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
        return S256_SUCCESS;
    }
    msm_plan pl = msm_make_plan(n);
    while (pl.c < MSM_MAX_C && (size_t)pl.nwin * n > msm_entries_capacity(ctx)) pl = msm_plan_for_c(pl.c + 1);
    uint32_t total = (uint32_t)pl.total;
    CK(cudaMemsetAsync(ctx->msm_counts, 0, ((size_t)total + 1) * 4, s));
    LAUNCH(ctx, k_msm_digits<false>, grid_for(n), 0, s, k32, n, pl, ctx->msm_counts, ctx->msm_cursor, ctx->msm_entries);
    size_t bytes = ctx->msm_cub_bytes;
    CK(cub::DeviceScan::ExclusiveSum(ctx->msm_cub, bytes, ctx->msm_counts, ctx->msm_offsets, (int)(total + 1), s));
    CK(cudaMemcpyAsync(ctx->msm_cursor, ctx->msm_offsets, (size_t)total * 4, cudaMemcpyDeviceToDevice, s));
    LAUNCH(ctx, k_msm_digits<true>, grid_for(n), 0, s, k32, n, pl, ctx->msm_counts, ctx->msm_cursor, ctx->msm_entries);
    // buckets -> slices of <= MSM_SLICE entries
    LAUNCH(ctx, k_msm_slice_counts, grid_for((size_t)total + 1), 0, s, total, ctx->msm_counts, ctx->msm_nsl);
    CK(cub::DeviceScan::ExclusiveSum(ctx->msm_cub, bytes, ctx->msm_nsl, ctx->msm_sloff, (int)(total + 1), s));
    size_t max_slices = (size_t)total + ((size_t)pl.nwin * n) / MSM_SLICE + 1;
    if (max_slices > ctx->msm_max_slices) max_slices = ctx->msm_max_slices;
    LAUNCH(ctx, k_msm_slices, grid_for(max_slices), 0, s, (uint32_t)max_slices, total, ctx->msm_sloff, ctx->msm_offsets,
           ctx->msm_entries, ctx->aff, ctx->msm_buckets);
    int parts = 1;
    for (int w = 0; w < pl.nwin; w += pl.nwin - 1 > 0 ? pl.nwin - 1 : 1) {  // first and top window cover both sizes
        int nbw = msm_window_buckets(pl, w);
        int p = (nbw + S256_MSM_WT * msm_seg_for(nbw) - 1) / (S256_MSM_WT * msm_seg_for(nbw));
        if (p > parts) parts = p;
    }
    k_msm_windows<<<dim3(parts, pl.nwin), S256_MSM_WT, 0, s>>>(pl, ctx->msm_buckets, ctx->msm_sloff, ctx->msm_win, parts);
    k_msm_fold<<<1, 64, 0, s>>>(pl.nwin, ctx->msm_win, parts);
    k_msm_final<<<1, 1, 0, s>>>(pl, ctx->msm_win, parts, ctx->msm_acc, first);
    ctx->launches.fetch_add(3, std::memory_order_relaxed);
    return S256_SUCCESS;
}

// host pointers -> msm_acc holds the projective sum; *invalid = 1 if a point failed to decode
static int msm_run(s256_ctx *ctx, const uint8_t *k32, const uint8_t *pt65, size_t n, int vartime, uint32_t *invalid) {
    int rc = msm_ensure(ctx);
    if (rc != S256_SUCCESS) return rc;
    cudaStream_t s = ctx->stream;
    CK(cudaMemsetAsync(ctx->msm_flag, 0, 4, s));
    if (n == 0) {
        pt id;
        pt_set_identity(id);
        CK(cudaMemcpyAsync(ctx->msm_acc, &id, sizeof(pt), cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
    }
    int first = 1;
    rc = for_chunks(ctx, n, [&](size_t off, size_t c) {
        CK(cudaMemcpyAsync(ctx->in_a, pt65 + 65 * off, 65 * c, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->in_b, k32 + 32 * off, 32 * c, cudaMemcpyHostToDevice, s));
        int r = chunk_msm(ctx, ctx->in_b, ctx->in_a, c, vartime, first, s);
        first = 0;
        return r;
    });
    if (rc != S256_SUCCESS) return rc;
    CK(cudaMemcpyAsync(invalid, ctx->msm_flag, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return check_launch(ctx);
}
// msm_acc -> 65-byte encoding + status
static int msm_finish(s256_ctx *ctx, uint8_t *out65, uint8_t *status) {
    cudaStream_t s = ctx->stream;
    LAUNCH(ctx, k_finish_affine<INV_K>, 1, 0, s, (size_t)1, ctx->msm_acc, (const uint8_t *)nullptr,
           (const uint8_t *)nullptr, ctx->cstat, 0, ctx->out, ctx->st, (const uint8_t *)nullptr);
    CK(cudaMemcpyAsync(out65, ctx->out, 65, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(status, ctx->st, 1, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return check_launch(ctx);
}
extern "C" int s256_msm(s256_ctx *ctx, const uint8_t *k32, const uint8_t *pt65, size_t n, int vartime, uint8_t *out65,
                        uint8_t *status) {
    ENTER(ctx);
    if (!out65 || !status || (n && (!k32 || !pt65))) return S256_ERR_ARG;
    uint32_t invalid = 0;
    int rc = msm_run(ctx, k32, pt65, n, vartime, &invalid);
    if (rc != S256_SUCCESS) return rc;
    if (invalid) {
        memset(out65, 0, 65);
        *status = S256_ST_INVALID;
        return S256_SUCCESS;
    }
    return msm_finish(ctx, out65, status);
}
extern "C" int s256_msm_partial(s256_ctx *ctx, const uint8_t *k32, const uint8_t *pt65, size_t n, int vartime,
                                uint8_t *partial96, uint8_t *status) {
    ENTER(ctx);
    if (!partial96 || !status || (n && (!k32 || !pt65))) return S256_ERR_ARG;
    uint32_t invalid = 0;
    int rc = msm_run(ctx, k32, pt65, n, vartime, &invalid);
    if (rc != S256_SUCCESS) return rc;
    if (invalid) {
        memset(partial96, 0, 96);
        *status = S256_ST_INVALID;
        return S256_SUCCESS;
    }
    cudaStream_t s = ctx->stream;
    k_export_partial<<<1, 1, 0, s>>>(ctx->msm_acc, ctx->out);
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    CK(cudaMemcpyAsync(partial96, ctx->out, 96, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *status = S256_ST_OK;
    return check_launch(ctx);
}
extern "C" int s256_msm_combine(s256_ctx *ctx, const uint8_t *partials96, size_t m, uint8_t *out65, uint8_t *status) {
    ENTER(ctx);
    if (!out65 || !status || (m && !partials96)) return S256_ERR_ARG;
    if (96 * m > 65 * ctx->cap) return S256_ERR_ARG;
    int rc = msm_ensure(ctx);
    if (rc != S256_SUCCESS) return rc;
    cudaStream_t s = ctx->stream;
    CK(cudaMemsetAsync(ctx->msm_flag, 0, 4, s));
    if (m) CK(cudaMemcpyAsync(ctx->in_a, partials96, 96 * m, cudaMemcpyHostToDevice, s));
    k_combine_partials<<<1, 1, 0, s>>>(ctx->in_a, m, ctx->msm_acc, ctx->msm_flag);
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    uint32_t invalid = 0;
    CK(cudaMemcpyAsync(&invalid, ctx->msm_flag, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (invalid) {
        memset(out65, 0, 65);
        *status = S256_ST_INVALID;
        return S256_SUCCESS;
    }
    return msm_finish(ctx, out65, status);
}

// ---------------------------------------------------------------------------
// debug / measurement
// ---------------------------------------------------------------------------
extern "C" int s256_debug_gen_table(s256_ctx *ctx, int wbits, int nwin, uint8_t *out) {
    ENTER(ctx);
    if (!out || wbits < 1 || wbits > 16 || nwin < 1 || wbits * nwin > 256) return S256_ERR_ARG;
    size_t total = (size_t)nwin << wbits;
    apt *d = nullptr;
    CK(cudaMalloc(&d, total * sizeof(apt)));
    LAUNCH(ctx, k_gen_table, grid_for(total), 0, ctx->stream, d, wbits, total);
    apt *h = (apt *)malloc(total * sizeof(apt));
    cudaError_t e = cudaMemcpyAsync(h, d, total * sizeof(apt), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) {
        free(h);
        ctx->last_err = cudaGetErrorString(e);
        return S256_ERR_CUDA;
    }
    // byte serialisation only (BE X || Y per entry, d = 0 skipped)
    for (int w = 0; w < nwin; w++)
        for (size_t dgt = 1; dgt < ((size_t)1 << wbits); dgt++) {
            const apt &a = h[((size_t)w << wbits) + dgt];
            fe_to_be32(out, a.x);
            fe_to_be32(out + 32, a.y);
            out += 64;
        }
    free(h);
    return S256_SUCCESS;
}

extern "C" int s256_debug_field_op(s256_ctx *ctx, int op, const uint8_t *a32, const uint8_t *b32, size_t n,
                                   uint8_t *out32) {
    ENTER(ctx);
    if (n && (!a32 || !b32 || !out32)) return S256_ERR_ARG;
    if (n > ctx->cap) return S256_ERR_ARG;
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(ctx->in_a, a32, 32 * n, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->in_c, b32, 32 * n, cudaMemcpyHostToDevice, s));
    LAUNCH(ctx, k_field_op, grid_for(n), 0, s, op, ctx->in_a, ctx->in_c, n, ctx->out);
    CK(cudaMemcpyAsync(out32, ctx->out, 32 * n, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return check_launch(ctx);
}

// Per-kernel timing of k_dsm with CUDA events on the launching stream: enable, run
// steps, then read (synchronises on the recorded events and clears them).
extern "C" int s256_profile_enable(s256_ctx *ctx, int enable) {
    ENTER(ctx);
    for (auto &pr : ctx->dsm_events) {
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    ctx->dsm_events.clear();
    ctx->profiling = enable != 0;
    return S256_SUCCESS;
}
extern "C" int s256_profile_read(s256_ctx *ctx, double *dsm_ms_total, uint64_t *dsm_launches) {
    ENTER(ctx);
    double total = 0;
    uint64_t cnt = 0;
    for (auto &pr : ctx->dsm_events) {
        CK(cudaEventSynchronize(pr.second));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, pr.first, pr.second));
        total += ms;
        cnt++;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    ctx->dsm_events.clear();
    if (dsm_ms_total) *dsm_ms_total = total;
    if (dsm_launches) *dsm_launches = cnt;
    return S256_SUCCESS;
}

// Integer-pipe probes (microbench.cuh): operations per second for one variant.
template <int V>
static void launch_probe(int blocks, int iters, cudaStream_t s, unsigned long long *sink) {
    k_int_probe<V><<<blocks, 256, 0, s>>>(12345u, iters, sink);
}
extern "C" int s256_microbench_variant(s256_ctx *ctx, int variant, int iters, double *ops_per_s, double *ms_out) {
    ENTER(ctx);
    if (iters < 1 || variant < 0 || variant >= MB_NVARIANTS) return S256_ERR_ARG;
    typedef void (*fn_t)(int, int, cudaStream_t, unsigned long long *);
    static const fn_t fns[MB_NVARIANTS] = {launch_probe<0>, launch_probe<1>, launch_probe<2>, launch_probe<3>,
                                           launch_probe<4>, launch_probe<5>, launch_probe<6>};
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int blocks = 148 * 8;
    fns[variant](blocks, 16, ctx->stream, ctx->sink);
    CK(cudaEventRecord(e0, ctx->stream));
    fns[variant](blocks, iters, ctx->stream, ctx->sink);
    CK(cudaEventRecord(e1, ctx->stream));
    ctx->launches.fetch_add(2);
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ops_per_s) *ops_per_s = (double)blocks * 256 * (double)iters * mb_ops_per_trip(variant) / (ms * 1e-3);
    if (ms_out) *ms_out = ms;
    return check_launch(ctx);
}

extern "C" int s256_microbench_imad(s256_ctx *ctx, int iters, double *mac32_per_s, double *ms) {
    // the carry-chained form is the one the field multiplier issues, and the fastest MAC32 form measured
    return s256_microbench_variant(ctx, MB_MADC_CHAIN, iters, mac32_per_s, ms);
}

// MAC32 per item actually executed (DESIGN.md "work per item"): F_p mul = 73 (64 + 9),
// F_p square = 45 (36 + 9), small-constant mul = 9, Z_n modmul = 139 (64 + 40 + 30 + 5).
extern "C" double s256_mac32_per_item(const char *name) {
    const double M = 73, S = 45, SM = 9, ZN = 139;
    const double dbl = 6 * M + 2 * S + SM, add = 12 * M + 2 * SM, mix = 11 * M + 2 * SM;
    const double inv_fe = 255 * S + 15 * M, sqrt_fe = 254 * S + 13 * M + 2 * S + M, inv_sc = 330 * ZN;
    const double oncurve = 2 * S + M;
    const double table = (DSM_TS / 2) * dbl + (DSM_TS / 2 - 1) * mix;
    const double ladder = (DSM_ND - 1) * DSM_W * dbl + 2 * DSM_ND * add + DSM_ND * M;
    const double comb = COMB_NW * mix;
    const double dsm = table + ladder + comb;
    const double split = 3 * ZN + 2 * 64;
    const double affine = (3 + 2) * M + inv_fe / INV_K;
    std::string s(name ? name : "");
    if (s == "k_dsm") return dsm;  // the ladder kernel alone
    if (s == "ecdsa_verify") return oncurve + (5 * ZN + inv_sc / INV_K + split) + dsm + 2 * M;
    if (s == "ecdsa_recover") return sqrt_fe + (6 * ZN + inv_sc / INV_K + split) + dsm + affine;
    if (s == "schnorr_verify") return sqrt_fe + (ZN + split) + dsm + affine;
    if (s == "double_scalar_mult_basepoint_vartime") return oncurve + split + dsm + affine;
    if (s == "scalar_base_mult") return CT_NW * mix + affine;
    if (s == "schnorr_sign") return 2 * (CT_NW * mix + affine) + 2 * ZN;  // + ~9 SHA-256 blocks
    if (s == "ecdsa_sign_rfc6979") return CT_NW * mix + affine + (5 * ZN + inv_sc / INV_K);  // + 22 SHA-256 blocks
    if (s == "scalar_mult" || s == "ecdh") {
        const double tab = (CTM_TS / 2) * dbl + (CTM_TS / 2 - 1) * mix;
        const double lad = (CTM_ND - 1) * CTM_W * dbl + 2 * CTM_ND * add + CTM_ND * M;
        return oncurve + split + tab + lad + affine;
    }
    return 0.0;
}
