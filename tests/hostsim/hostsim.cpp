// hostsim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles secp256k1-voi_b200/csrc/kernels.cuh with g++ (portable limb
// arithmetic instead of PTX) and drives the *same* item/group functions in the
// same order as api.cu's pipelines, one "thread" at a time.  It exists because
// the build container has no GPU: kernel logic (recoding, ladders, batched
// inversion, encodings) is debugged here against the oracle before GPU time
// is spent.  It is never linked into, or called by, the product library.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../secp256k1-voi_b200/csrc/kernels.cuh"
#include "../../secp256k1-voi_b200/csrc/msm.cuh"
#include "../../secp256k1-voi_b200/csrc/h2c.cuh"

using namespace s256;
#define EXPORT extern "C" __attribute__((visibility("default")))

static std::vector<apt> g_ct, g_ct_small;
static constexpr int K = 32;

static void ensure_tables() {
    if (!g_ct.empty()) return;
    // both resident tables, as s256_init builds them: 6-bit windows (throughput kernel), 5-bit (lane-split kernels)
    constexpr int SZB = ct_cfg<CT_WB>::SZ, SZS = ct_cfg<CT_WB_SMALL>::SZ;
    g_ct.resize((size_t)ct_cfg<CT_WB>::NW * SZB);
    for (size_t idx = 0; idx < g_ct.size(); idx++)
        item_gen_multiple(g_ct[idx], (uint32_t)(idx / SZB), (uint32_t)(idx % SZB) + 1u, CT_WB);
    g_ct_small.resize((size_t)ct_cfg<CT_WB_SMALL>::NW * SZS);
    for (size_t idx = 0; idx < g_ct_small.size(); idx++)
        item_gen_multiple(g_ct_small[idx], (uint32_t)(idx / SZS), (uint32_t)(idx % SZS) + 1u, CT_WB_SMALL);
}
// The comb has 12 * 2^21 entries (1.5 GiB); the simulation fills only the entries a batch uses, in
// lazily zeroed memory (calloc: untouched pages cost nothing).
static apt *g_comb_tab = nullptr;
static uint8_t *g_comb_have = nullptr;
static void ensure_comb_entry(uint32_t w, uint32_t mag) {  // entry mag * 2^(WB*w) * G, mag in [1, 2^(WB-1)]
    if (!g_comb_tab) {
        g_comb_tab = (apt *)calloc((size_t)COMB_NW * COMB_SZ, sizeof(apt));
        g_comb_have = (uint8_t *)calloc((size_t)COMB_NW * COMB_SZ, 1);
    }
    size_t idx = (size_t)w * COMB_SZ + (mag - 1u);
    if (!g_comb_have[idx]) {
        item_gen_multiple(g_comb_tab[idx], w, mag, COMB_WB);
        g_comb_have[idx] = 1;
    }
}

struct scratch {
    std::vector<apt> aff;
    std::vector<sc> u1;
    std::vector<int8_t> dig1, dig2;
    std::vector<uint8_t> sfl, pvalid, cstat;
    std::vector<pt> tbl, res;
    explicit scratch(size_t n)
        : aff(n), u1(n), dig1(n * DSM_ND), dig2(n * DSM_ND), sfl(n), pvalid(n), cstat(n), tbl(n * DSM_TSTRIDE), res(n) {}
};

static void run_dsm(scratch &s, size_t n) {
    for (size_t i = 0; i < n; i++) {
        uint32_t carry = 0;
        for (int w = 0; w < COMB_NW; w++) {
            int32_t d = comb_digit(s.u1[i], w, carry);
            if (d) ensure_comb_entry((uint32_t)w, (uint32_t)(d < 0 ? -d : d));
        }
    }
    if (!g_comb_tab) ensure_comb_entry(0, 1);
    for (size_t i = 0; i < n; i++)
        item_dsm(i, n, s.aff.data(), s.u1.data(), s.dig1.data(), s.dig2.data(), s.sfl.data(), s.tbl.data(), s.res.data(),
                 g_comb_tab);
}
static void run_finish(scratch &s, size_t n, bool use_pvalid, bool use_sfl, int mode, uint8_t *out, uint8_t *status,
                       const uint8_t *sig) {
    size_t stride = (n + K - 1) / K;
    for (size_t t = 0; t < stride; t++)
        group_finish<K>(t, stride, n, s.res.data(), use_pvalid ? s.pvalid.data() : nullptr,
                        use_sfl ? s.sfl.data() : nullptr, s.cstat.data(), mode, out, status, sig);
}

EXPORT void sim_ecdsa_verify(const uint8_t *pk, const uint8_t *dg, const uint8_t *sig, uint32_t flags, size_t n, uint8_t *ok) {
    scratch s(n);
    for (size_t i = 0; i < n; i++) s.pvalid[i] = item_decode_uncompressed(s.aff[i], pk + 65 * i);
    size_t stride = (n + K - 1) / K;
    for (size_t t = 0; t < stride; t++)
        group_ecdsa_scalars<K>(t, stride, n, dg, sig, flags, s.u1.data(), s.dig1.data(), s.dig2.data(), s.sfl.data());
    run_dsm(s, n);
    for (size_t i = 0; i < n; i++) {
        uint32_t valid = (uint32_t)(s.pvalid[i] != 0) & (uint32_t)(s.sfl[i] & SFL_VALID);
        ok[i] = item_ecdsa_finish(s.res[i], sig + 64 * i, valid);
    }
}
EXPORT void sim_ecdsa_recover(const uint8_t *dg, const uint8_t *sig65, size_t n, uint8_t *pk65, uint8_t *status) {
    scratch s(n);
    for (size_t i = 0; i < n; i++) s.pvalid[i] = item_decode_recover(s.aff[i], sig65 + 65 * i);
    size_t stride = (n + K - 1) / K;
    for (size_t t = 0; t < stride; t++)
        group_recover_scalars<K>(t, stride, n, dg, sig65, s.u1.data(), s.dig1.data(), s.dig2.data(), s.sfl.data());
    run_dsm(s, n);
    run_finish(s, n, true, true, 3, pk65, status, nullptr);
}
EXPORT void sim_schnorr_verify(const uint8_t *pkx, const uint8_t *msg, size_t msg_len, const uint8_t *sig, size_t n, uint8_t *ok) {
    scratch s(n);
    for (size_t i = 0; i < n; i++) s.pvalid[i] = item_decode_xonly(s.aff[i], pkx + 32 * i);
    for (size_t i = 0; i < n; i++)
        item_schnorr_scalars(i, n, pkx, msg, msg_len, sig, s.u1.data(), s.dig1.data(), s.dig2.data(), s.sfl.data());
    run_dsm(s, n);
    run_finish(s, n, true, true, 2, nullptr, ok, sig);
}
EXPORT void sim_double_scalar_mult(const uint8_t *u1, const uint8_t *u2, const uint8_t *pt65, size_t n, uint8_t *out65, uint8_t *status) {
    scratch s(n);
    for (size_t i = 0; i < n; i++) s.pvalid[i] = item_decode_uncompressed(s.aff[i], pt65 + 65 * i);
    for (size_t i = 0; i < n; i++) item_plain_scalars(i, n, u1, u2, s.u1.data(), s.dig1.data(), s.dig2.data(), s.sfl.data());
    run_dsm(s, n);
    run_finish(s, n, true, false, 0, out65, status, nullptr);
}
EXPORT void sim_scalar_base_mult(const uint8_t *k32, size_t n, uint8_t *out65, uint8_t *status) {
    ensure_tables();
    scratch s(n);
    for (size_t i = 0; i < n; i++) {
        sc k;
        sc_from_be32(k, k32 + 32 * i);
        if (n <= 8) {  // exercise the lane-split form (T = 8 then T = 4) the way the kernel folds it
            int T = (i & 1) ? 4 : 8;
            pt part[8];
            for (int p = 0; p < T; p++) item_base_mult_ct_part<CT_WB_SMALL>(part[p], k, g_ct_small.data(), p, T);
            for (int off = T / 2; off >= 1; off >>= 1)
                for (int p = 0; p < off; p++) pt_add(part[p], part[p], part[p + off]);
            s.res[i] = part[0];
        } else if (i & 1) {
            item_base_mult_ct(s.res[i], k, g_ct.data());
        } else {  // the throughput kernels' form: Jacobian accumulator (kernels.cuh)
            item_base_mult_ct_jac(s.res[i], k, g_ct.data());
        }
    }
    run_finish(s, n, false, false, 0, out65, status, nullptr);
}
EXPORT void sim_scalar_mult(const uint8_t *k32, const uint8_t *pt65, size_t n, int mode, uint8_t *out, uint8_t *status) {
    scratch s(n);
    for (size_t i = 0; i < n; i++) s.pvalid[i] = item_decode_uncompressed(s.aff[i], pt65 + 65 * i);
    for (size_t i = 0; i < n; i++) {
        CtTableGlobal T{s.tbl.data() + i * (size_t)DSM_TSTRIDE};
        item_scalar_mult_ct_affine(i, s.aff.data(), k32, T, s.tbl.data() + i * (size_t)DSM_TSTRIDE, s.res.data());
    }
    run_finish(s, n, true, false, mode, out, status, nullptr);
}
EXPORT void sim_point_decompress(const uint8_t *pt33, size_t n, uint8_t *out65, uint8_t *status) {
    for (size_t i = 0; i < n; i++) {
        apt a;
        uint8_t ok = item_decode_compressed(a, pt33 + 33 * i);
        memset(out65 + 65 * i, 0, 65);
        if (ok) {
            out65[65 * i] = 4;
            fe_to_be32(out65 + 65 * i + 1, a.x);
            fe_to_be32(out65 + 65 * i + 33, a.y);
        }
        status[i] = ok ? ST_OK : ST_INVALID;
    }
}
// Pippenger flow of api.cu's chunk_msm, sequentially: digits -> counting sort -> bucket sums ->
// window running sums (256 "threads" + tree) -> Horner.  force_c > 0 overrides the window size.
EXPORT void sim_msm(const uint8_t *k32, const uint8_t *pt65, size_t n, int vartime, int force_c, uint8_t *out65, uint8_t *status,
                    uint8_t *partial96) {
    scratch s(n ? n : 1);
    bool invalid = false;
    for (size_t i = 0; i < n; i++) {
        s.pvalid[i] = item_decode_uncompressed(s.aff[i], pt65 + 65 * i);
        invalid |= !s.pvalid[i];
    }
    pt acc;
    pt_set_identity(acc);
    if (n && (!vartime || n < 32) && force_c == 0) {
        for (size_t i = 0; i < n; i++) {
            CtTableGlobal T{s.tbl.data() + i * (size_t)DSM_TSTRIDE};
            item_scalar_mult_ct_affine(i, s.aff.data(), k32, T, s.tbl.data() + i * (size_t)DSM_TSTRIDE, s.res.data());
        }
        for (size_t i = 0; i < n; i++) pt_add(acc, acc, s.res[i]);
    } else if (n) {
        // endomorphism form, as chunk_msm runs it: 2n virtual points (P, lambda P) with the 128-bit halves of the scalars
        const size_t nv = 2 * n;
        msm_plan pl = msm_make_plan(nv, 128);
        if (force_c) pl = msm_plan_for_c(force_c, 128);
        size_t total = (size_t)pl.total;
        std::vector<uint32_t> counts(total + 1, 0), offsets(total + 1, 0), cursor(total + 1, 0), entries((size_t)pl.nwin * nv);
        std::vector<int32_t> dig((size_t)pl.nwin * nv);
        std::vector<apt> aff2(nv);
        std::vector<uint32_t> hneg(nv);
        for (size_t i = 0; i < n; i++) {
            sc k, m[2];
            sc_from_be32(k, k32 + 32 * i);
            uint32_t ng[2];
            msm_glv_halves(m[0], ng[0], m[1], ng[1], k);
            msm_glv_points(aff2[2 * i], aff2[2 * i + 1], s.aff[i]);
            for (int h = 0; h < 2; h++) {
                size_t v = 2 * i + h;
                hneg[v] = ng[h];
                int32_t d[MSM_MAX_WIN];
                msm_digits(d, m[h], pl);
                for (int w = 0; w < pl.nwin; w++) {
                    dig[(size_t)w * nv + v] = d[w];
                    if (d[w]) counts[(size_t)w * pl.nb + (size_t)(std::abs(d[w]) - 1)]++;
                }
            }
        }
        for (size_t b = 0; b < total; b++) offsets[b + 1] = offsets[b] + counts[b];
        cursor = offsets;
        for (size_t v = nv; v-- > 0;)  // reverse order: bucket order must not matter
            for (int w = 0; w < pl.nwin; w++) {
                int32_t d = dig[(size_t)w * nv + v];
                if (d) entries[cursor[(size_t)w * pl.nb + (size_t)(std::abs(d) - 1)]++] = ((uint32_t)v << 1) | ((uint32_t)(d < 0) ^ hneg[v]);
            }
        // slices of <= MSM_SLICE entries, exactly as the kernels cut them
        std::vector<uint32_t> sl_off(total + 1, 0);
        for (size_t b = 0; b < total; b++) sl_off[b + 1] = sl_off[b] + msm_slices_of(counts[b], pl.slice);
        std::vector<pt> slice_sum(sl_off[total]);
        std::vector<uint32_t> bucket_of(sl_off[total]);
        for (uint32_t sidx = 0; sidx < sl_off[total]; sidx++) {
            uint32_t st, en;
            bucket_of[sidx] = msm_slice_range(st, en, sidx, sl_off.data(), offsets.data(), (uint32_t)total, pl.slice);
            msm_bucket_sum(slice_sum[sidx], entries.data(), st, en, aff2.data());
        }
        for (uint32_t sidx = 0; sidx < sl_off[total]; sidx++)
            msm_superslice_fold(slice_sum.data(), sidx, sl_off[bucket_of[sidx]], sl_off[bucket_of[sidx] + 1]);
        // window stage exactly as k_msm_windows / k_msm_windows2 run it: (run, sum) per thread, suffix
        // scan of the runs across the CTA, position weights as doublings, tree sum
        const int T = MSM_WT;
        auto cta_weighted = [&](std::vector<pt> &run, std::vector<pt> &sum, int lg, pt &R, pt &A) {
            std::vector<pt> sh = run;
            for (int d = 1; d < T; d <<= 1) {
                std::vector<pt> nx = sh;
                for (int t = 0; t + d < T; t++) pt_add(nx[t], sh[t], sh[t + d]);
                sh = nx;
            }
            R = sh[0];
            std::vector<pt> v(T);
            v[0] = sum[0];
            for (int t = 1; t < T; t++) msm_weigh(v[t], sum[t], sh[t], lg);
            for (int stride = T / 2; stride >= 1; stride >>= 1)
                for (int t = 0; t < stride; t++) pt_add(v[t], v[t], v[t + stride]);
            A = v[0];
        };
        int parts = std::max(msm_parts_for(pl.nb), msm_parts_for(pl.nb_top));
        std::vector<pt> part((size_t)pl.nwin * parts * 2), win(pl.nwin);
        for (int w = 0; w < pl.nwin; w++) {
            int nbw = msm_window_buckets(pl, w), seg = msm_seg_for(nbw), pw = msm_parts_for(nbw);
            for (int blk = 0; blk < pw; blk++) {
                std::vector<pt> run(T), sum(T);
                for (int t = 0; t < T; t++) {
                    int lo = (blk * T + t) * seg, hi = lo + seg;
                    if (hi > nbw) hi = nbw;
                    if (lo < hi) msm_segment_pair(run[t], sum[t], slice_sum.data(), sl_off.data(), (uint32_t)w * (uint32_t)pl.nb, lo, hi);
                    else { pt_set_identity(run[t]); pt_set_identity(sum[t]); }
                }
                cta_weighted(run, sum, msm_log2(seg), part[((size_t)w * parts + blk) * 2], part[((size_t)w * parts + blk) * 2 + 1]);
            }
            std::vector<pt> run(T), sum(T);
            for (int t = 0; t < T; t++) {
                if (t < pw) { run[t] = part[((size_t)w * parts + t) * 2]; sum[t] = part[((size_t)w * parts + t) * 2 + 1]; }
                else { pt_set_identity(run[t]); pt_set_identity(sum[t]); }
            }
            pt R;
            cta_weighted(run, sum, msm_log2(T * seg), R, win[w]);
        }
        parts = 1;
        msm_horner(acc, win.data(), pl, 1, parts);
    }
    memset(out65, 0, 65);
    if (partial96) pt_to_be96(partial96, acc);
    if (invalid) { *status = ST_INVALID; if (partial96) memset(partial96, 0, 96); return; }
    scratch f(1);
    f.res[0] = acc;
    run_finish(f, 1, false, false, 0, out65, status, nullptr);
}
EXPORT void sim_msm_combine(const uint8_t *partials96, size_t m, uint8_t *out65, uint8_t *status) {
    pt acc;
    pt_set_identity(acc);
    bool bad = false;
    for (size_t j = 0; j < m; j++) {
        pt q;
        pt_from_be96(q, partials96 + 96 * j);
        uint32_t ok = fe_limbs_are_canonical(q.x) & fe_limbs_are_canonical(q.y) & fe_limbs_are_canonical(q.z) & pt_on_curve(q);
        if (!ok) { bad = true; continue; }
        pt_add(acc, acc, q);
    }
    memset(out65, 0, 65);
    if (bad) { *status = ST_INVALID; return; }
    scratch f(1);
    f.res[0] = acc;
    run_finish(f, 1, false, false, 0, out65, status, nullptr);
}
EXPORT void sim_ecdsa_sign_rfc6979(const uint8_t *priv32, const uint8_t *digest32, size_t n, uint8_t *sig64, uint8_t *recid,
                                   uint8_t *status) {
    ensure_tables();
    scratch s(n);
    std::vector<uint8_t> kbuf(32 * n), valid(n), r65(65 * n), rst(n);
    for (size_t i = 0; i < n; i++) valid[i] = item_rfc6979_nonce(kbuf.data() + 32 * i, priv32 + 32 * i, digest32 + 32 * i);
    for (size_t i = 0; i < n; i++) {
        sc k;
        sc_from_be32(k, kbuf.data() + 32 * i);
        item_base_mult_ct(s.res[i], k, g_ct.data());
    }
    run_finish(s, n, false, false, 0, r65.data(), rst.data(), nullptr);
    size_t stride = (n + K - 1) / K;
    for (size_t t = 0; t < stride; t++)
        group_sign_finish<K>(t, stride, n, priv32, digest32, kbuf.data(), valid.data(), r65.data(), sig64, recid, status);
}
EXPORT void sim_schnorr_sign(const uint8_t *priv32, const uint8_t *msg, size_t msg_len, const uint8_t *aux32, size_t n,
                             uint8_t *sig64, uint8_t *status) {
    ensure_tables();
    scratch s(n);
    std::vector<uint8_t> p65(65 * n), r65(65 * n), kbuf(32 * n), valid(n), st(n);
    for (size_t i = 0; i < n; i++) {
        sc k;
        sc_from_be32(k, priv32 + 32 * i);
        item_base_mult_ct(s.res[i], k, g_ct.data());
    }
    run_finish(s, n, false, false, 0, p65.data(), st.data(), nullptr);
    for (size_t i = 0; i < n; i++)
        valid[i] = item_schnorr_nonce(kbuf.data() + 32 * i, priv32 + 32 * i, p65.data() + 65 * i, msg + msg_len * i, msg_len,
                                      aux32 + 32 * i);
    for (size_t i = 0; i < n; i++) {
        sc k;
        sc_from_be32(k, kbuf.data() + 32 * i);
        item_base_mult_ct(s.res[i], k, g_ct.data());
    }
    run_finish(s, n, false, false, 0, r65.data(), st.data(), nullptr);
    for (size_t i = 0; i < n; i++)
        item_schnorr_sign_finish(sig64 + 64 * i, status + i, priv32 + 32 * i, p65.data() + 65 * i, r65.data() + 65 * i,
                                 kbuf.data() + 32 * i, msg + msg_len * i, msg_len, valid[i]);
}
EXPORT void sim_hash_to_curve(const uint8_t *dst, size_t dst_len, const uint8_t *msg, size_t msg_len, size_t n, int ro,
                              uint8_t *out65, uint8_t *status) {
    scratch s(n ? n : 1);
    for (size_t i = 0; i < n; i++) item_hash_to_curve(s.res[i], dst, (int)dst_len, msg + msg_len * i, msg_len, ro);
    run_finish(s, n, false, false, 0, out65, status, nullptr);
}
EXPORT void sim_expand_xmd(const uint8_t *dst, size_t dst_len, const uint8_t *msg, size_t msg_len, int len, uint8_t *out) {
    h2c_expand_xmd(out, len, dst, (int)dst_len, msg, msg_len);
}
// Differential check of the two group laws: the same chain of operations on the Jacobian formulas (jac.cuh, with
// their explicit exceptional cases) and on the complete ones (point.cuh).  ops[j]: 0 = double, 1 = add b, 2 = add -b,
// 3 = add a, 4 = add -a; the accumulator starts at the identity.  Both results leave through the shared conversion.
EXPORT void sim_jac_vs_complete(const uint8_t *a65, const uint8_t *b65, const uint8_t *ops, size_t nops, uint8_t *out_jac65,
                                uint8_t *st_jac, uint8_t *out_rcb65, uint8_t *st_rcb) {
    scratch s(2);
    apt A, B;
    if (!item_decode_uncompressed(A, a65) || !item_decode_uncompressed(B, b65)) {
        *st_jac = *st_rcb = 0;
        return;
    }
    fe_ops<true> f;
    pt j, c;
    j.x = j.y = j.z = fe_zero();
    uint32_t inf = 1u;
    pt_set_identity(c);
    for (size_t k = 0; k < nops; k++) {
        if (ops[k] == 0) {
            jac_double(f, j, j);
            pt_double<true>(c, c);
        } else {
            apt q = (ops[k] == 1 || ops[k] == 2) ? B : A;
            if (ops[k] == 2 || ops[k] == 4) fe_neg(q.y, q.y);
            jac_add_mixed_var(f, j, inf, q.x, q.y);
            pt_add_mixed<true>(c, c, q.x, q.y);
        }
    }
    jac_to_projective(f, s.res[0], j, inf);
    s.res[1] = c;
    uint8_t out[130], st[2];
    run_finish(s, 2, false, false, 0, out, st, nullptr);
    memcpy(out_jac65, out, 65);
    memcpy(out_rcb65, out + 65, 65);
    *st_jac = st[0];
    *st_rcb = st[1];
}
EXPORT void sim_gen_table(int wbits, int nwin, uint8_t *out) {
    for (int w = 0; w < nwin; w++)
        for (uint32_t d = 1; d < (1u << wbits); d++) {
            apt a;
            item_gen_multiple(a, (uint32_t)w, d, wbits);
            fe_to_be32(out, a.x);
            fe_to_be32(out + 32, a.y);
            out += 64;
        }
}
EXPORT void sim_field_op(int op, const uint8_t *a32, const uint8_t *b32, size_t n, uint8_t *out32) {
    for (size_t i = 0; i < n; i++) {
        if (op < 16 || op >= 20) {
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
                case 7: fe_invert_fermat(r, a); break;
                case 8: fe_mul_vt(r, a, b); break;
                case 9: fe_add_vt(r, a, b); break;
                case 10: fe_sub_vt(r, a, b); break;
                case 13: fe_mul_small_vt(r, a, 21u); break;
                case 14: fe_sqr_vt(r, a); break;
                case 15: fe_mul8_vt(r, a); break;
                case 20: fe_mul2_vt(r, a); break;
                case 21: fe_mul3_vt(r, a); break;
                case 22: fe_sub2_vt(r, a, b); break;
                case 23: fe_submul8_vt(r, a, b); break;
                case 24: fe_ops<false>::mul2add(r, a, b, b, a); break;  // the constant-time fused a b + c d (device)
                case 25: fe_ops<false>::mul2sub(r, a, a, b, b); break;
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
                case 19: sc_invert_fermat(r, a); break;
                default: r = sc_zero();
            }
            sc_to_be32(out32 + 32 * i, r);
        }
    }
}
