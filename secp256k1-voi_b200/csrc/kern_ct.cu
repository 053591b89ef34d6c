// kern_ct.cu -- the constant-time kernels (secret scalars).
#ifndef S256_BM_CALL
#define S256_MUL_INLINE 1
#endif
#include "kernels.cuh"
#include "launchers.h"
#include <cstdio>
#include <cstdlib>

using namespace s256;
// 16 warps per SM in both configurations; the table (NW * SZ * 64 bytes of shared memory per CTA) decides
// how many CTAs fit: 88 KB (6-bit windows) twice with 256 threads, 53 KB (5-bit) four times with 128.
#define S256_TPB 128       // the lane-split kernels
#define S256_TPB_BIG 256   // the throughput kernel
#ifndef S256_BM_MINB
#define S256_BM_MINB 4
#endif
#define S256_BM_MINB_BIG 2
#ifndef S256_BM_W7_MIN_DEFAULT
#define S256_BM_W7_MIN_DEFAULT 16385  // batches from this size on take the 7-bit kernel: faster than the 6-bit one at every
                                       // size above the lane-split range (20 000: 0.304 vs 0.308 ms ... 2^20: 6.56 vs 6.84 ms)
#endif
#ifndef S256_BM_SPLIT4_MAX
#define S256_BM_SPLIT4_MAX 16384
#endif
template <int WB>
__host__ __device__ constexpr size_t ct_bytes_big() { return (size_t)ct_cfg<WB>::NW * ct_cfg<WB>::SZ * sizeof(apt); }
constexpr size_t CT_BYTES_BIG = ct_bytes_big<CT_WB>();
// the large-batch flavour: 7-bit windows, 37 additions instead of 43 over a 152 KB table -- one CTA of 512 threads per SM
// (the same 16 warps); staging 152 KB per CTA only pays when every CTA has many scalars to work through
constexpr int CT_WB_HUGE = 7;
#define S256_TPB_HUGE 512
constexpr size_t CT_BYTES_HUGE = ct_bytes_big<CT_WB_HUGE>();
// lane-split kernels: every window's row is followed by 16 bytes of padding (bank spreading, kernels.cuh)
constexpr size_t CT_ROW_SMALL = (size_t)ct_cfg<CT_WB_SMALL>::SZ * sizeof(apt), CT_ROW_SMALL_PAD = CT_ROW_SMALL + 16;
constexpr size_t CT_BYTES_SMALL = (size_t)ct_cfg<CT_WB_SMALL>::NW * CT_ROW_SMALL_PAD;

template <int WB>
__device__ __forceinline__ void base_mult_ct_body(const uint8_t *k32, size_t n, const apt *tab_g, pt *res) {
    extern __shared__ uint4 smem_raw[];
    apt *tab = reinterpret_cast<apt *>(smem_raw);
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(tab_g);
        const int nvec = (int)(ct_bytes_big<WB>() / 16);
        for (int v = threadIdx.x; v < nvec; v += blockDim.x) smem_raw[v] = src[v];
    }
    __syncthreads();
    // warps of 32 consecutive items are dealt round-robin over the CTAs (warp g -> CTA g % grid), so that a
    // batch that does not fill every resident CTA still loads all SMs evenly
    const size_t G = gridDim.x, wpb = blockDim.x / 32, lane = threadIdx.x & 31;
    for (size_t g = blockIdx.x + G * (threadIdx.x / 32); g * 32 < n; g += G * wpb) {
        size_t i = g * 32 + lane;
        if (i >= n) break;
        sc k;
        sc_from_be32(k, k32 + 32 * i);
        pt acc;
#ifndef S256_BM_RCB
        item_base_mult_ct_jac<WB>(acc, k, tab);  // Jacobian accumulator, result in homogeneous form (kernels.cuh)
#else
        item_base_mult_ct<WB>(acc, k, tab);
#endif
        res[i] = acc;
    }
}
__global__ void __launch_bounds__(S256_TPB_BIG, S256_BM_MINB_BIG)
    k_base_mult_ct(const uint8_t *k32, size_t n, const apt *tab_g, pt *res) {
    base_mult_ct_body<CT_WB>(k32, n, tab_g, res);
}
__global__ void __launch_bounds__(S256_TPB_HUGE, 1)
    k_base_mult_ct_w7(const uint8_t *k32, size_t n, const apt *tab_g, pt *res) {
    base_mult_ct_body<CT_WB_HUGE>(k32, n, tab_g, res);
}


// T lanes per scalar (T divides 32): partial sums, then a shuffle tree of complete additions
template <int T>
__global__ void __launch_bounds__(S256_TPB, S256_BM_MINB)
    k_base_mult_ct_split(const uint8_t *k32, size_t n, const apt *tab_g, pt *res) {
    extern __shared__ uint4 smem_raw[];
    apt *tab = reinterpret_cast<apt *>(smem_raw);
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(tab_g);
        constexpr int per_row = (int)(CT_ROW_SMALL / 16), nvec = ct_cfg<CT_WB_SMALL>::NW * per_row;
        for (int v = threadIdx.x; v < nvec; v += blockDim.x) smem_raw[v + v / per_row] = src[v];   // + one uint4 per row
    }
    __syncthreads();
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t item = gid / T;
    int part = (int)(gid % T);
    bool live = item < n;  // uniform inside each group of T lanes
    sc k;
    sc_from_be32(k, k32 + 32 * (live ? item : 0));
    pt acc;
    item_base_mult_ct_part<CT_WB_SMALL>(acc, k, tab, part, T, CT_ROW_SMALL_PAD);
#pragma unroll
    for (int off = T / 2; off >= 1; off >>= 1) {
        pt o;
        uint32_t *dst = o.x.v;
        const uint32_t *src = acc.x.v;
#pragma unroll
        for (int q = 0; q < 24; q++) dst[q] = __shfl_down_sync(0xffffffffu, src[q], off, T);
        pt_add(acc, acc, o);
    }
    if (live && part == 0) res[item] = acc;
}

static int g_sm_count = 148;  // B200; replaced by the device's own count at init
void s256_ct_kernels_init() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
        g_sm_count = sms;
    cudaFuncSetAttribute(k_base_mult_ct_split<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CT_BYTES_SMALL);
    cudaFuncSetAttribute(k_base_mult_ct_split<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CT_BYTES_SMALL);
    cudaFuncSetAttribute(k_base_mult_ct_split<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CT_BYTES_SMALL);
    cudaFuncSetAttribute(k_base_mult_ct_split<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CT_BYTES_SMALL);
    cudaFuncSetAttribute(k_base_mult_ct_split<16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_base_mult_ct_split<32>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_base_mult_ct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CT_BYTES_BIG);
    cudaFuncSetAttribute(k_base_mult_ct_w7, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CT_BYTES_HUGE);
    cudaFuncSetAttribute(k_base_mult_ct_w7, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    // ask for the largest shared-memory carve-out: without the hint the driver sizes it for ONE CTA and the
    // second (fourth) CTA of an SM waits for the first to retire
    cudaFuncSetAttribute(k_base_mult_ct, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_base_mult_ct_split<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_base_mult_ct_split<8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (getenv("S256_TRACE")) {
        int a = 0, b = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_base_mult_ct, S256_TPB_BIG, CT_BYTES_BIG);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_base_mult_ct_split<8>, S256_TPB, CT_BYTES_SMALL);
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, k_base_mult_ct);
        fprintf(stderr, "[s256 trace] k_base_mult_ct: %d CTAs/SM (regs %d, static smem %zu, dyn %zu); split<8>: %d CTAs/SM\n", a,
                fa.numRegs, fa.sharedSizeBytes, CT_BYTES_BIG, b);
    }
}
// tab_big: [NW(6)][SZ(6)], tab_small: [NW(5)][SZ(5)], tab_huge: [NW(7)][SZ(7)] (ctx->ct_tab, ctx->ct_tab_small, ctx->ct_tab_huge)
void s256_launch_base_mult_ct(const uint8_t *k32, size_t n, const apt *tab_big, const apt *tab_small, const apt *tab_huge,
                              pt *res, cudaStream_t s) {
    if (n == 0) return;
    static const size_t w7_min = getenv("S256_BM_W7_MIN") ? (size_t)atoll(getenv("S256_BM_W7_MIN")) : (size_t)S256_BM_W7_MIN_DEFAULT;
    if (tab_huge && n >= w7_min) {
        size_t warps7 = (n + 31) / 32;
        unsigned g7 = warps7 < (size_t)g_sm_count ? (unsigned)warps7 : (unsigned)g_sm_count;
        k_base_mult_ct_w7<<<g7, S256_TPB_HUGE, CT_BYTES_HUGE, s>>>(k32, n, tab_huge, res);
        return;
    }
    // small batches are latency bound: deal the windows of each scalar to 8 / 4 lanes (16 lanes measured
    // slower at n = 4096: 0.227 against 0.198 ms)
    static const int force_t = getenv("S256_BM_T") ? atoi(getenv("S256_BM_T")) : 0;  // tuning knob: lanes per scalar
    if (force_t == 16 || force_t == 32) {
        if (force_t == 16)
            k_base_mult_ct_split<16><<<(unsigned)((n * 16 + S256_TPB - 1) / S256_TPB), S256_TPB, CT_BYTES_SMALL, s>>>(k32, n, tab_small, res);
        else
            k_base_mult_ct_split<32><<<(unsigned)((n * 32 + S256_TPB - 1) / S256_TPB), S256_TPB, CT_BYTES_SMALL, s>>>(k32, n, tab_small, res);
        return;
    }
    if (n <= 8192) {
        k_base_mult_ct_split<8><<<(unsigned)((n * 8 + S256_TPB - 1) / S256_TPB), S256_TPB, CT_BYTES_SMALL, s>>>(k32, n, tab_small, res);
        return;
    }
    if (n <= (size_t)S256_BM_SPLIT4_MAX) {
        k_base_mult_ct_split<4><<<(unsigned)((n * 4 + S256_TPB - 1) / S256_TPB), S256_TPB, CT_BYTES_SMALL, s>>>(k32, n, tab_small, res);
        return;
    }
    size_t warps = (n + 31) / 32;
    unsigned maxg = (unsigned)g_sm_count * S256_BM_MINB_BIG;
    unsigned grid = warps < maxg ? (unsigned)warps : maxg;
    k_base_mult_ct<<<grid, S256_TPB_BIG, CT_BYTES_BIG, s>>>(k32, n, tab_big, res);
}
