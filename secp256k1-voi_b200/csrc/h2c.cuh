// h2c.cuh -- hash to curve for secp256k1 (RFC 9380, suites secp256k1_XMD:SHA-256_SSWU_RO_ / _NU_).
//
// Replaces secec/h2c/h2c.go:25-63 (hash_to_curve / encode_to_curve),
// secec/h2c/h2c_expand_message.go:33-139 (expand_message_xmd, SHA-256),
// point_h2c.go:23-55 (SetUniformBytes), internal/swu/swu.go:70-199 (simplified SWU on the
// 3-isogenous curve E' and the isogeny map) and internal/field/field_sqrt_ratio.go:25-63,
// field_reduce.go:24-64.  One item per thread, branch-free in the data (the reference is
// constant time here too).  The three inversions of one map (tv4, x_den, y_den) share a single
// Fermat chain through Montgomery's trick; a vanishing isogeny denominator yields the identity.
// Constants: RFC 9380 section 8.7 and appendix E.1 (the values in internal/swu/swu.go:13-66).
#pragma once
#include "fe.cuh"
#include "point.cuh"
#include "sc.cuh"
#include "sha256.cuh"

namespace s256 {

S256_CONST(H2C_A, 8, 0x1A444533u, 0x405447C0u, 0xCB6F0E5Du, 0xE953D363u, 0xF0F5D272u, 0xA08A5558u, 0xDD661ADCu, 0x3F8731ABu)
S256_CONST(H2C_K10, 8, 0xAAAAA8C7u, 0x8E38E38Du, 0xE38E38E3u, 0x38E38E38u, 0x8E38E38Eu, 0xE38E38E3u, 0x38E38E38u, 0x8E38E38Eu)
S256_CONST(H2C_K11, 8, 0xF17C6581u, 0xDFFF1044u, 0x0BF63B92u, 0xD595D2FCu, 0xA7FD44C5u, 0xB9F315CEu, 0x0BC321D5u, 0x07D3D4C8u)
S256_CONST(H2C_K12, 8, 0x3D9DD262u, 0x4ECBD0B5u, 0x037C4031u, 0xE4506144u, 0xCA25CAECu, 0xE2A413DEu, 0x23F234E6u, 0x534C328Du)
S256_CONST(H2C_K13, 8, 0xAAAAA88Cu, 0x8E38E38Du, 0xE38E38E3u, 0x38E38E38u, 0x8E38E38Eu, 0xE38E38E3u, 0x38E38E38u, 0x8E38E38Eu)
S256_CONST(H2C_K20, 8, 0x781EB49Bu, 0x9FE6B745u, 0x42F8487Du, 0x86CD4095u, 0xB7B640DDu, 0x9CA34CCBu, 0x3D94918Au, 0xD3577119u)
S256_CONST(H2C_K21, 8, 0x2A8C6D14u, 0xC52A5661u, 0x1F5E41BBu, 0x06D36B64u, 0x1B542254u, 0xF7C4B2D5u, 0x4383DC1Du, 0xEDADC6F6u)
S256_CONST(H2C_K30, 8, 0x8E38E23Cu, 0xA12F684Bu, 0x12F684BDu, 0x2F684BDAu, 0xF684BDA1u, 0x684BDA12u, 0x84BDA12Fu, 0x4BDA12F6u)
S256_CONST(H2C_K31, 8, 0x201D71A3u, 0xDFFC90FCu, 0xD686DA6Fu, 0x647AB046u, 0x12A0A6D5u, 0xA9D0A54Bu, 0xD5CB7C0Fu, 0xC75E0C32u)
S256_CONST(H2C_K32, 8, 0x9ECEE931u, 0xA765E85Au, 0x01BE2018u, 0x722830A2u, 0x6512E576u, 0x715209EFu, 0x91F91A73u, 0x29A61946u)
S256_CONST(H2C_K33, 8, 0x38E38D84u, 0x84BDA12Fu, 0x4BDA12F6u, 0xBDA12F68u, 0xDA12F684u, 0xA12F684Bu, 0x12F684BDu, 0x2F684BDAu)
S256_CONST(H2C_K40, 8, 0xFFFFF93Bu, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu)
S256_CONST(H2C_K41, 8, 0x685C2573u, 0xDFB425D2u, 0xC8E8D978u, 0x9467C1BFu, 0x2722C298u, 0xD5E9E663u, 0xB8BDB49Fu, 0x7A06534Bu)
S256_CONST(H2C_K42, 8, 0xBFD2A76Fu, 0xA7BF8192u, 0x2F0D6299u, 0x0A3D2116u, 0xA8FE337Eu, 0xF3A70C3Fu, 0x6545CA2Cu, 0x6484AA71u)
S256_CONST(H2C_C2, 8, 0x303C4A59u, 0x286729C8u, 0xA74789DDu, 0xEC184F00u, 0x8F842AFEu, 0x7AD13FB3u, 0x724013E5u, 0x31FDF302u)
S256_CONST(H2C_Z, 8, 0xFFFFFC24u, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu)

S256_HD fe fe_const(const uint32_t *c) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = c[i];
    return r;
}

constexpr int H2C_MAX_DST = 255;

// expand_message_xmd with SHA-256; len <= 96 here (hash_to_field of one or two elements, L = 48).
// dst_len must be in [1, 255] (the host pre-hashes oversize DSTs, RFC 9380 section 5.3.3).
S256_HD void h2c_expand_xmd(uint8_t *out, int len, const uint8_t *dst, int dst_len, const uint8_t *msg, size_t msg_len) {
    uint8_t b0[32], bi[32], x[3], dl = (uint8_t)dst_len;
    sha_stream c;
    sha_init(c);
    for (int i = 0; i < 64; i++) {  // Z_pad
        uint8_t z = 0;
        sha_update(c, &z, 1);
    }
    sha_update(c, msg, msg_len);
    x[0] = (uint8_t)(len >> 8); x[1] = (uint8_t)len; x[2] = 0;
    sha_update(c, x, 3);
    sha_update(c, dst, (size_t)dst_len);
    sha_update(c, &dl, 1);
    sha_final(c, b0);
    int ell = (len + 31) / 32, off = 0;
    for (int i = 1; i <= ell; i++) {
        uint8_t t[32];
        for (int k = 0; k < 32; k++) t[k] = (i == 1) ? b0[k] : (uint8_t)(b0[k] ^ bi[k]);
        sha_init(c);
        sha_update(c, t, 32);
        x[0] = (uint8_t)i;
        sha_update(c, x, 1);
        sha_update(c, dst, (size_t)dst_len);
        sha_update(c, &dl, 1);
        sha_final(c, bi);
        for (int k = 0; k < 32 && off < len; k++) out[off++] = bi[k];
    }
}

// 48 big-endian bytes -> element: lo + hi * 2^256, 2^256 = 2^32 + 977 (mod p)
S256_HD void fe_from_wide48(fe &r, const uint8_t *b) {
    fe lo, hi = fe_zero(), d = fe_zero();
    fe_from_be32(lo, b + 16);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint8_t *q = b + 4 * (3 - i);
        hi.v[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
    }
    d.v[0] = S256_DELTA_LO;
    d.v[1] = 1;
    fe_mul(hi, hi, d);
    fe_add(r, lo, hi);
}

// x^((p-3)/4), (p-3)/4 = 2^254 - 2^30 - 245 (field_sqrt_ratio.go:65-185 pow3mod4)
S256_HD void fe_pow_p34(fe &r, const fe &a) {
    fe x223, x22, x2, x3, t;
    fe_pow_x223(x223, x22, x2, x3, a);
    fe_sqr_n(t, x223, 23); fe_mul(t, t, x22);
    fe_sqr_n(t, t, 5); fe_mul(t, t, a);
    fe_sqr_n(t, t, 3); fe_mul(r, t, x2);
}
// RFC 9380 F.2.1.2 sqrt_ratio for p = 3 mod 4: (isQR, y) with y = sqrt(u/v) or sqrt(Z*u/v)
S256_HD uint32_t fe_sqrt_ratio(fe &z, const fe &u, const fe &v) {
    fe tv1, tv2, tv3, y1, y2;
    fe_sqr(tv1, v);
    fe_mul(tv2, u, v);
    fe_mul(tv1, tv1, tv2);
    fe_pow_p34(y1, tv1);
    fe_mul(y1, y1, tv2);
    fe_mul(y2, y1, fe_const(S256_K(H2C_C2)));
    fe_sqr(tv3, y1);
    fe_mul(tv3, tv3, v);
    uint32_t qr = fe_equal(tv3, u);
    fe_cmov(z, y2, y1, qr);
    return qr;
}

// SetUniformBytes (point_h2c.go:23-55): u -> point on E (projective; identity on the exceptional case)
S256_HD void h2c_map_to_curve(pt &out, const uint8_t *u48) {
    const fe A = fe_const(S256_K(H2C_A)), Z = fe_const(S256_K(H2C_Z)), one = fe_one();
    fe u, tv1, tv2, tv3, tv4, tv5, tv6, x, y, y1, nt;
    fe_from_wide48(u, u48);
    // --- simplified SWU on E' (swu.go:70-147) ---
    fe_sqr(tv1, u);
    fe_mul(tv1, Z, tv1);
    fe_sqr(tv2, tv1);
    fe_add(tv2, tv2, tv1);
    fe_add(tv3, tv2, one);
    fe_mul_small(tv3, tv3, 1771u);  // B'
    uint32_t sel = fe_is_zero(tv2);
    fe_neg(nt, tv2);
    fe_cmov(tv4, nt, Z, sel);
    fe_mul(tv4, A, tv4);
    fe_sqr(tv2, tv3);
    fe_sqr(tv6, tv4);
    fe_mul(tv5, A, tv6);
    fe_add(tv2, tv2, tv5);
    fe_mul(tv2, tv2, tv3);
    fe_mul(tv6, tv6, tv4);
    fe_mul_small(tv5, tv6, 1771u);
    fe_add(tv2, tv2, tv5);
    fe_mul(x, tv1, tv3);
    uint32_t is_sq = fe_sqrt_ratio(y1, tv2, tv6);
    fe_mul(y, tv1, u);
    fe_mul(y, y, y1);
    fe_cmov(x, x, tv3, is_sq);
    fe_cmov(y, y, y1, is_sq);
    uint32_t flip = fe_is_odd(u) ^ fe_is_odd(y);
    fe_cneg(y, y, flip);
    // x' = x / tv4 is deferred: X' = x, denominator tv4 (never zero)
    // --- isogeny map (swu.go:149-199) in terms of x' = x / tv4 ---
    // work with the affine x' after one shared inversion of tv4 * x_den * y_den; first get x' itself:
    // x_den and y_den are polynomials in x', so clear tv4 by homogenising: x' = x / d, d = tv4.
    fe d = tv4, d2, d3, X = x, X2, X3, xn, xd, yn, yd, t;
    fe_sqr(d2, d);
    fe_mul(d3, d2, d);
    fe_sqr(X2, X);
    fe_mul(X3, X2, X);
    // x_num * d^3 = k13 X^3 + k12 X^2 d + k11 X d^2 + k10 d^3
    fe_mul(xn, fe_const(S256_K(H2C_K13)), X3);
    fe_mul(t, fe_const(S256_K(H2C_K12)), X2); fe_mul(t, t, d); fe_add(xn, xn, t);
    fe_mul(t, fe_const(S256_K(H2C_K11)), X); fe_mul(t, t, d2); fe_add(xn, xn, t);
    fe_mul(t, fe_const(S256_K(H2C_K10)), d3); fe_add(xn, xn, t);
    // x_den * d^2 = X^2 + k21 X d + k20 d^2
    fe_mul(xd, fe_const(S256_K(H2C_K21)), X); fe_mul(xd, xd, d); fe_add(xd, xd, X2);
    fe_mul(t, fe_const(S256_K(H2C_K20)), d2); fe_add(xd, xd, t);
    // y_num * d^3 = k33 X^3 + k32 X^2 d + k31 X d^2 + k30 d^3
    fe_mul(yn, fe_const(S256_K(H2C_K33)), X3);
    fe_mul(t, fe_const(S256_K(H2C_K32)), X2); fe_mul(t, t, d); fe_add(yn, yn, t);
    fe_mul(t, fe_const(S256_K(H2C_K31)), X); fe_mul(t, t, d2); fe_add(yn, yn, t);
    fe_mul(t, fe_const(S256_K(H2C_K30)), d3); fe_add(yn, yn, t);
    // y_den * d^3 = X^3 + k42 X^2 d + k41 X d^2 + k40 d^3
    fe_mul(yd, fe_const(S256_K(H2C_K42)), X2); fe_mul(yd, yd, d); fe_add(yd, yd, X3);
    fe_mul(t, fe_const(S256_K(H2C_K41)), X); fe_mul(t, t, d2); fe_add(yd, yd, t);
    fe_mul(t, fe_const(S256_K(H2C_K40)), d3); fe_add(yd, yd, t);
    // x = (xn / d^3) / (xd / d^2) = xn / (xd * d);  y = Y * (yn / d^3) / (yd / d^3) = Y * yn / yd
    uint32_t bad = fe_is_zero(xd) | fe_is_zero(yd);
    fe xdd;
    fe_mul(xdd, xd, d);
    // projective point with Z3 = xdd * yd: X3 = xn * yd, Y3 = Y * yn * xdd
    pt r, id;
    fe_mul(r.x, xn, yd);
    fe_mul(r.y, y, yn);
    fe_mul(r.y, r.y, xdd);
    fe_mul(r.z, xdd, yd);
    pt_set_identity(id);
    pt_cmov(out, r, id, bad);
}

// hash_to_curve (ro = 1) / encode_to_curve (ro = 0): secec/h2c/h2c.go:25-63
S256_HD void item_hash_to_curve(pt &out, const uint8_t *dst, int dst_len, const uint8_t *msg, size_t msg_len, int ro) {
    uint8_t ub[96];
    h2c_expand_xmd(ub, ro ? 96 : 48, dst, dst_len, msg, msg_len);
    pt q0;
    h2c_map_to_curve(q0, ub);
    if (ro) {
        pt q1;
        h2c_map_to_curve(q1, ub + 48);
        pt_add(q0, q0, q1);
    }
    out = q0;
}

}  // namespace s256
