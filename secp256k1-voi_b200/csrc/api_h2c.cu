// api_h2c.cu -- hash to curve (RFC 9380): kernels (h2c.cuh) and entry points.
#include "ctx.h"
#include "h2c.cuh"

__global__ void __launch_bounds__(S256_TPB) k_hash_to_curve(const uint8_t *dst, int dst_len, const uint8_t *msg,
                                                            size_t msg_len, size_t n, int ro, pt *res) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pt r;
    item_hash_to_curve(r, dst, dst_len, msg + msg_len * i, msg_len, ro);
    res[i] = r;
}
__global__ void __launch_bounds__(S256_TPB) k_expand_xmd(const uint8_t *dst, int dst_len, const uint8_t *msg,
                                                         size_t msg_len, size_t n, int len, uint8_t *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    h2c_expand_xmd(out + (size_t)len * i, len, dst, dst_len, msg + msg_len * i, msg_len);
}

// RFC 9380 section 5.3.3 / h2c_expand_message.go:52-62: an oversize DST is replaced by its hash (host side:
// public bytes, one hash per call)
static int prepare_dst(const uint8_t *dst, size_t dst_len, uint8_t buf[H2C_MAX_DST], int *out_len) {
    if (!dst || dst_len == 0) return S256_ERR_ARG;  // errInvalidDomainSep
    if (dst_len > (size_t)H2C_MAX_DST) {
        sha_stream c;
        sha_init(c);
        sha_update(c, (const uint8_t *)"H2C-OVERSIZE-DST-", 17);
        sha_update(c, dst, dst_len);
        sha_final(c, buf);
        *out_len = 32;
    } else {
        memcpy(buf, dst, dst_len);
        *out_len = (int)dst_len;
    }
    return S256_SUCCESS;
}
// grows the message staging buffer (ctx->in_b) like the Schnorr entry points do
static int ensure_msg_staging(s256_ctx *ctx, size_t msg_len, size_t n) {
    size_t need = (msg_len ? msg_len : 1) * (n < ctx->cap ? n : ctx->cap);
    if (int grc = grow_in_b(ctx, need)) return grc;
    return S256_SUCCESS;
}

extern "C" int s256_hash_to_curve(s256_ctx *ctx, const uint8_t *dst, size_t dst_len, const uint8_t *msg, size_t msg_len,
                                  size_t n, int random_oracle, uint8_t *out65, uint8_t *status) {
    ENTER(ctx);
    scratch_guard sg_(ctx, ctx->stream, true);
    if (n && ((!msg && msg_len) || !out65 || !status)) return S256_ERR_ARG;
    uint8_t dbuf[H2C_MAX_DST];
    int dl = 0;
    int rc = prepare_dst(dst, dst_len, dbuf, &dl);
    if (rc != S256_SUCCESS) return rc;
    rc = ensure_msg_staging(ctx, msg_len, n);
    if (rc != S256_SUCCESS) return rc;
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(ctx->in_c, dbuf, (size_t)dl, cudaMemcpyHostToDevice, s));
    rc = for_chunks(ctx, n, [&](size_t off, size_t c) {
        view v = view_at(ctx, 0);
        if (msg_len) CK(cudaMemcpyAsync(ctx->in_b, msg + msg_len * off, msg_len * c, cudaMemcpyHostToDevice, s));
        LAUNCH(ctx, k_hash_to_curve, grid_for(c), 0, s, ctx->in_c, dl, ctx->in_b, msg_len, c, random_oracle ? 1 : 0, v.res);
        s256_launch_finish_affine(ctx, c, v.res, nullptr, nullptr, v.cstat, 0, v.out, v.st, nullptr, s);
        CK(cudaMemcpyAsync(out65 + 65 * off, v.out, 65 * c, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(status + off, v.st, c, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        return S256_SUCCESS;
    });
    return rc != S256_SUCCESS ? rc : check_launch(ctx);
}

extern "C" int s256_expand_message_xmd(s256_ctx *ctx, const uint8_t *dst, size_t dst_len, const uint8_t *msg,
                                       size_t msg_len, size_t n, size_t len_in_bytes, uint8_t *out) {
    ENTER(ctx);
    scratch_guard sg_(ctx, ctx->stream, true);
    if (n && ((!msg && msg_len) || !out)) return S256_ERR_ARG;
    if (len_in_bytes == 0 || len_in_bytes > 96) return S256_ERR_ARG;  // the suites here need 48 or 96
    uint8_t dbuf[H2C_MAX_DST];
    int dl = 0;
    int rc = prepare_dst(dst, dst_len, dbuf, &dl);
    if (rc != S256_SUCCESS) return rc;
    rc = ensure_msg_staging(ctx, msg_len, n);
    if (rc != S256_SUCCESS) return rc;
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(ctx->in_c, dbuf, (size_t)dl, cudaMemcpyHostToDevice, s));
    uint8_t *d_out = reinterpret_cast<uint8_t *>(ctx->res);  // 96 B per item
    rc = for_chunks(ctx, n, [&](size_t off, size_t c) {
        if (msg_len) CK(cudaMemcpyAsync(ctx->in_b, msg + msg_len * off, msg_len * c, cudaMemcpyHostToDevice, s));
        LAUNCH(ctx, k_expand_xmd, grid_for(c), 0, s, ctx->in_c, dl, ctx->in_b, msg_len, c, (int)len_in_bytes, d_out);
        CK(cudaMemcpyAsync(out + len_in_bytes * off, d_out, len_in_bytes * c, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        return S256_SUCCESS;
    });
    return rc != S256_SUCCESS ? rc : check_launch(ctx);
}
