/*
 * secp256k1_oracle.c -- CPU restatement of the secp256k1-voi hot path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  It is the checker (tests/, smoke(),
 * bench.py's cpu_baseline / --impl reference legs), never the product: nothing
 * under secp256k1-voi_b200/ may include, link or call it.
 *
 * The reference (Go, gitlab.com/yawning/secp256k1-voi) cannot be compiled in
 * this image (no Go toolchain), so this is a from-scratch plain-C restatement
 * of the reference's *algorithms* -- 4x64-bit-limb Montgomery arithmetic like
 * the fiat code, the Renes-Costello-Batina complete formulas, the 8-bit /
 * 4-bit precomputed generator tables, the GLV split with 4-bit fixed windows,
 * Straus multi-scalar multiplication, one Fermat inversion per affine
 * conversion -- so that it serves both as the bit-exact oracle and as the
 * "reference-algorithm" CPU baseline.  Parity is pinned: tests/test_oracle_*.py
 * check it against every vector the reference's own tests hold for this path
 * (Wycheproof ECDSA/ECDH, BIP-340, RFC 6979, libsecp256k1 KATs, the embedded
 * generator table's sha256) and against OpenSSL.
 *
 * Each function cites the reference file:line whose behaviour it follows
 * (paths relative to the reference root).
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <pthread.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef uint8_t u8;

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* 256-bit helpers                                                            */
/* ------------------------------------------------------------------------- */

typedef struct { u64 v[4]; } u256; /* little-endian limbs */

/* internal/helpers/helpers.go:47-65 -- big-endian bytes <-> 4x64 LE limbs */
static void be32_to_limbs(u64 l[4], const u8 b[32]) {
    for (int i = 0; i < 4; i++) {
        u64 w = 0;
        for (int j = 0; j < 8; j++) w = (w << 8) | b[(3 - i) * 8 + j];
        l[i] = w;
    }
}
static void limbs_to_be32(u8 b[32], const u64 l[4]) {
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) b[(3 - i) * 8 + j] = (u8)(l[i] >> (56 - 8 * j));
}
static u64 sub256(u64 r[4], const u64 a[4], const u64 b[4]) {
    u64 borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - b[i] - borrow;
        r[i] = (u64)d;
        borrow = (u64)(d >> 64) & 1;
    }
    return borrow;
}
static u64 add256(u64 r[4], const u64 a[4], const u64 b[4]) {
    u64 carry = 0;
    for (int i = 0; i < 4; i++) {
        u128 s = (u128)a[i] + b[i] + carry;
        r[i] = (u64)s;
        carry = (u64)(s >> 64);
    }
    return carry;
}
static int is_zero256(const u64 a[4]) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
static int eq256(const u64 a[4], const u64 b[4]) {
    return ((a[0] ^ b[0]) | (a[1] ^ b[1]) | (a[2] ^ b[2]) | (a[3] ^ b[3])) == 0;
}

/* ------------------------------------------------------------------------- */
/* Generic 4x64 word-by-word Montgomery arithmetic (the shape of the fiat code:*/
/* internal/fiat/secp256k1montgomery/secp256k1montgomery.go:87 and its scalar  */
/* twin; 36 64x64 multiplies per product, fully reduced outputs).              */
/* ------------------------------------------------------------------------- */

typedef struct {
    u64 m[4];    /* modulus */
    u64 minv;    /* -m^-1 mod 2^64 */
    u64 r2[4];   /* R^2 mod m */
    u64 one[4];  /* R mod m */
} mont_ctx;

static mont_ctx FP, FN;

static const u64 P_LIMBS[4] = {0xFFFFFFFEFFFFFC2FULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL};
static const u64 N_LIMBS[4] = {0xBFD25E8CD0364141ULL, 0xBAAEDCE6AF48A03BULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL};
/* scalar.go:33-38 -- n >> 1 */
static const u64 HALF_N[4] = {0xDFE92F46681B20A0ULL, 0x5D576E7357A4501DULL, 0xFFFFFFFFFFFFFFFFULL, 0x7FFFFFFFFFFFFFFFULL};

static void mont_mul(u64 r[4], const u64 a[4], const u64 b[4], const mont_ctx *c) {
    u64 t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5;
    const u64 *m = c->m;
    for (int i = 0; i < 4; i++) {
        u128 acc;
        u64 ai = a[i];
        acc = (u128)ai * b[0] + t0; t0 = (u64)acc; acc >>= 64;
        acc += (u128)ai * b[1] + t1; t1 = (u64)acc; acc >>= 64;
        acc += (u128)ai * b[2] + t2; t2 = (u64)acc; acc >>= 64;
        acc += (u128)ai * b[3] + t3; t3 = (u64)acc; acc >>= 64;
        acc += t4; t4 = (u64)acc; t5 = (u64)(acc >> 64);
        u64 q = t0 * c->minv;
        acc = (u128)q * m[0] + t0; acc >>= 64;
        acc += (u128)q * m[1] + t1; t0 = (u64)acc; acc >>= 64;
        acc += (u128)q * m[2] + t2; t1 = (u64)acc; acc >>= 64;
        acc += (u128)q * m[3] + t3; t2 = (u64)acc; acc >>= 64;
        acc += t4; t3 = (u64)acc; t4 = t5 + (u64)(acc >> 64);
    }
    u64 t[4] = {t0, t1, t2, t3}, d[4];
    u64 borrow = sub256(d, t, m);
    /* result = t - m if (t4:t) >= m */
    int ge = (t4 != 0) || (borrow == 0);
    for (int i = 0; i < 4; i++) r[i] = ge ? d[i] : t[i];
}
static void mont_add(u64 r[4], const u64 a[4], const u64 b[4], const mont_ctx *c) {
    u64 s[4], d[4];
    u64 carry = add256(s, a, b);
    u64 borrow = sub256(d, s, c->m);
    int ge = carry || !borrow;
    for (int i = 0; i < 4; i++) r[i] = ge ? d[i] : s[i];
}
static void mont_sub(u64 r[4], const u64 a[4], const u64 b[4], const mont_ctx *c) {
    u64 d[4], s[4];
    u64 borrow = sub256(d, a, b);
    add256(s, d, c->m);
    for (int i = 0; i < 4; i++) r[i] = borrow ? s[i] : d[i];
}
static void mont_neg(u64 r[4], const u64 a[4], const mont_ctx *c) {
    u64 z[4] = {0, 0, 0, 0};
    mont_sub(r, z, a, c);
}
static void mont_to(u64 r[4], const u64 a[4], const mont_ctx *c) { mont_mul(r, a, c->r2, c); }
static void mont_from(u64 r[4], const u64 a[4], const mont_ctx *c) {
    u64 one[4] = {1, 0, 0, 0};
    mont_mul(r, a, one, c);
}
static void mont_ctx_init(mont_ctx *c, const u64 m[4]) {
    memcpy(c->m, m, 32);
    /* Newton iteration for m^-1 mod 2^64, then negate. */
    u64 inv = 1;
    for (int i = 0; i < 6; i++) inv *= 2 - m[0] * inv;
    c->minv = (u64)0 - inv;
    /* R mod m = 2^256 - m (valid because m > 2^255). */
    u64 z[4] = {0, 0, 0, 0};
    sub256(c->one, z, m);
    /* R^2 mod m by 256 modular doublings of R mod m. */
    u64 x[4];
    memcpy(x, c->one, 32);
    for (int i = 0; i < 256; i++) mont_add(x, x, x, c);
    memcpy(c->r2, x, 32);
}

/* ------------------------------------------------------------------------- */
/* Field elements mod p (Montgomery domain), internal/field/field.go          */
/* ------------------------------------------------------------------------- */

typedef struct { u64 v[4]; } fe;

static fe FE_ZERO, FE_ONE, FE_B, FE_B3, FE_BETA, FE_N;

static inline void fe_mul(fe *r, const fe *a, const fe *b) { mont_mul(r->v, a->v, b->v, &FP); }   /* field.go:73 */
static inline void fe_sqr(fe *r, const fe *a) { mont_mul(r->v, a->v, a->v, &FP); }                /* field.go:79 */
static inline void fe_add(fe *r, const fe *a, const fe *b) { mont_add(r->v, a->v, b->v, &FP); }   /* field.go:61 */
static inline void fe_sub(fe *r, const fe *a, const fe *b) { mont_sub(r->v, a->v, b->v, &FP); }   /* field.go:67 */
static inline void fe_neg(fe *r, const fe *a) { mont_neg(r->v, a->v, &FP); }                      /* field.go:85 */
static inline int fe_is_zero(const fe *a) { return is_zero256(a->v); }
static inline int fe_eq(const fe *a, const fe *b) { return eq256(a->v, b->v); }
static void fe_pow2k(fe *r, const fe *a, int k) { /* field.go:91 */
    fe t = *a;
    for (int i = 0; i < k; i++) fe_sqr(&t, &t);
    *r = t;
}
/* field.go:115-137 -- SetBytes reduces (value - p if >= p) and reports it;
 * field.go:123 SetCanonicalBytes rejects >= p. */
static int fe_set_bytes(fe *r, const u8 b[32]) {
    u64 l[4], d[4];
    be32_to_limbs(l, b);
    u64 borrow = sub256(d, l, P_LIMBS);
    int did_reduce = !borrow;
    if (did_reduce) memcpy(l, d, 32);
    mont_to(r->v, l, &FP);
    return did_reduce;
}
static int fe_set_canonical_bytes(fe *r, const u8 b[32]) { /* 1 = ok */
    u64 l[4], d[4];
    be32_to_limbs(l, b);
    if (!sub256(d, l, P_LIMBS)) return 0;
    mont_to(r->v, l, &FP);
    return 1;
}
static void fe_bytes(u8 b[32], const fe *a) { /* field.go:140 */
    u64 l[4];
    mont_from(l, a->v, &FP);
    limbs_to_be32(b, l);
}
static int fe_is_odd(const fe *a) { /* field.go:191-197 */
    u64 l[4];
    mont_from(l, a->v, &FP);
    return (int)(l[0] & 1);
}
static void fe_set_u64(fe *r, u64 x) {
    u64 l[4] = {x, 0, 0, 0};
    mont_to(r->v, l, &FP);
}

/* internal/field/field_invert.go:11-140 -- x^(p-2); Invert(0) = 0.
 * Same cost class (255 squarings + 15 multiplications) through the classic
 * run-of-ones chain for p - 2 = 2^256 - 2^32 - 979. */
static void fe_pow_common(fe *x223, fe *x22, fe *x2, const fe *a) {
    fe x3, x6, x9, x11, x44, x88, x176, x220, t;
    fe_sqr(&t, a); fe_mul(x2, &t, a);
    fe_sqr(&t, x2); fe_mul(&x3, &t, a);
    fe_pow2k(&t, &x3, 3); fe_mul(&x6, &t, &x3);
    fe_pow2k(&t, &x6, 3); fe_mul(&x9, &t, &x3);
    fe_pow2k(&t, &x9, 2); fe_mul(&x11, &t, x2);
    fe_pow2k(&t, &x11, 11); fe_mul(x22, &t, &x11);
    fe_pow2k(&t, x22, 22); fe_mul(&x44, &t, x22);
    fe_pow2k(&t, &x44, 44); fe_mul(&x88, &t, &x44);
    fe_pow2k(&t, &x88, 88); fe_mul(&x176, &t, &x88);
    fe_pow2k(&t, &x176, 44); fe_mul(&x220, &t, &x44);
    fe_pow2k(&t, &x220, 3); fe_mul(x223, &t, &x3);
}
static void fe_invert(fe *r, const fe *a) {
    fe x223, x22, x2, t;
    fe_pow_common(&x223, &x22, &x2, a);
    /* p-2 = 1^223 0 1^22 0000 1 0 1 1 0 1 */
    fe_pow2k(&t, &x223, 23); fe_mul(&t, &t, &x22);
    fe_pow2k(&t, &t, 5); fe_mul(&t, &t, a);
    fe_pow2k(&t, &t, 3); fe_mul(&t, &t, &x2);
    fe_pow2k(&t, &t, 2); fe_mul(r, &t, a);
}
/* internal/field/field_sqrt_ratio.go:14-63 with v = 1: candidate a^((p+1)/4),
 * returns (root, 1) if it squares back to a, else (0, 0). */
static int fe_sqrt(fe *r, const fe *a) {
    fe x223, x22, x2, t, chk;
    fe_pow_common(&x223, &x22, &x2, a);
    /* (p+1)/4 = 1^223 0 1^22 0000 11 00 */
    fe_pow2k(&t, &x223, 23); fe_mul(&t, &t, &x22);
    fe_pow2k(&t, &t, 6); fe_mul(&t, &t, &x2);
    fe_pow2k(&t, &t, 2);
    fe_sqr(&chk, &t);
    if (!fe_eq(&chk, a)) { *r = FE_ZERO; return 0; }
    *r = t;
    return 1;
}

/* ------------------------------------------------------------------------- */
/* Scalars mod n (Montgomery domain), scalar.go                               */
/* ------------------------------------------------------------------------- */

typedef struct { u64 v[4]; } sc;

static sc SC_ZERO, SC_ONE, SC_NEG_LAMBDA, SC_NEG_B1, SC_NEG_B2, SC_G1, SC_G2;

static inline void sc_mul(sc *r, const sc *a, const sc *b) { mont_mul(r->v, a->v, b->v, &FN); } /* scalar.go:78 */
static inline void sc_sqr(sc *r, const sc *a) { mont_mul(r->v, a->v, a->v, &FN); }
static inline void sc_add(sc *r, const sc *a, const sc *b) { mont_add(r->v, a->v, b->v, &FN); } /* scalar.go:66 */
static inline void sc_neg(sc *r, const sc *a) { mont_neg(r->v, a->v, &FN); }                    /* scalar.go:96 */
static inline int sc_is_zero(const sc *a) { return is_zero256(a->v); }
static inline int sc_eq(const sc *a, const sc *b) { return eq256(a->v, b->v); }
/* scalar.go:123-131 + reduceSaturated :272-292 -- reduce once, report. */
static int sc_set_bytes(sc *r, const u8 b[32]) {
    u64 l[4], d[4];
    be32_to_limbs(l, b);
    u64 borrow = sub256(d, l, N_LIMBS);
    int did_reduce = !borrow;
    if (did_reduce) memcpy(l, d, 32);
    mont_to(r->v, l, &FN);
    return did_reduce;
}
static int sc_set_canonical_bytes(sc *r, const u8 b[32]) { /* scalar.go:137-145; 1 = ok */
    u64 l[4], d[4];
    be32_to_limbs(l, b);
    if (!sub256(d, l, N_LIMBS)) return 0;
    mont_to(r->v, l, &FN);
    return 1;
}
static void sc_limbs(u64 l[4], const sc *a) { mont_from(l, a->v, &FN); }
static void sc_bytes(u8 b[32], const sc *a) { /* scalar.go:148-158 */
    u64 l[4];
    sc_limbs(l, a);
    limbs_to_be32(b, l);
}
static int sc_is_gt_half_n(const sc *a) { /* scalar.go:190-206 */
    u64 l[4], d[4];
    sc_limbs(l, a);
    u64 borrow = sub256(d, l, HALF_N);
    return !borrow && !is_zero256(d);
}
static void sc_set_limbs(sc *r, const u64 l[4]) { mont_to(r->v, l, &FN); }
/* scalar_invert.go:11-303 -- x^(n-2), Invert(0) = 0.  n - 2's top 127 bits are
 * ones; the low 129 bits are consumed with a 4-bit fixed window. */
static void sc_invert(sc *r, const sc *a) {
    sc tbl[16], t, x;
    tbl[0] = SC_ONE;
    tbl[1] = *a;
    for (int i = 2; i < 16; i++) sc_mul(&tbl[i], &tbl[i - 1], a);
    u64 e[4];
    u64 two[4] = {2, 0, 0, 0};
    sub256(e, N_LIMBS, two);
    /* x = a^(2^127 - 1): run-of-ones chain. */
    sc x2, x3, x6, x7, x14, x28, x56, x63, x126;
    sc_sqr(&t, a); sc_mul(&x2, &t, a);
    sc_sqr(&t, &x2); sc_mul(&x3, &t, a);
    t = x3; for (int i = 0; i < 3; i++) sc_sqr(&t, &t); sc_mul(&x6, &t, &x3);
    sc_sqr(&t, &x6); sc_mul(&x7, &t, a);
    t = x7; for (int i = 0; i < 7; i++) sc_sqr(&t, &t); sc_mul(&x14, &t, &x7);
    t = x14; for (int i = 0; i < 14; i++) sc_sqr(&t, &t); sc_mul(&x28, &t, &x14);
    t = x28; for (int i = 0; i < 28; i++) sc_sqr(&t, &t); sc_mul(&x56, &t, &x28);
    t = x56; for (int i = 0; i < 7; i++) sc_sqr(&t, &t); sc_mul(&x63, &t, &x7);
    t = x63; for (int i = 0; i < 63; i++) sc_sqr(&t, &t); sc_mul(&x126, &t, &x63);
    sc_sqr(&t, &x126); sc_mul(&x, &t, a); /* 2^127 - 1 */
    /* remaining 129 bits: bit 128 first (it is 0), then 32 nibbles. */
    sc_sqr(&x, &x);
    if ((e[2] >> 0) & 1) sc_mul(&x, &x, a);
    for (int i = 31; i >= 0; i--) {
        for (int k = 0; k < 4; k++) sc_sqr(&x, &x);
        unsigned nib = (unsigned)((e[i / 16] >> ((i % 16) * 4)) & 0xF);
        if (nib) sc_mul(&x, &x, &tbl[nib]);
    }
    *r = x;
}

/* ------------------------------------------------------------------------- */
/* Points: homogeneous projective, identity (0,1,0).  point.go:31-59          */
/* ------------------------------------------------------------------------- */

typedef struct { fe x, y, z; } pt;
typedef struct { fe x, y; } apt;

static pt PT_G;

static void pt_identity(pt *r) { r->x = FE_ZERO; r->y = FE_ONE; r->z = FE_ZERO; } /* point.go:42-49 */
static int pt_is_identity(const pt *p) { return fe_is_zero(&p->z); }             /* point.go:148-152 */
static void pt_neg(pt *r, const pt *p) { r->x = p->x; fe_neg(&r->y, &p->y); r->z = p->z; } /* point.go:89 */

/* point_projective.go:24-120 -- RCB'16 Algorithm 7 (a = 0, b3 = 21). */
static void pt_add(pt *v, const pt *p, const pt *q) {
    fe t0, t1, t2, t3, t4, x3, y3, z3;
    fe_mul(&t0, &p->x, &q->x); fe_mul(&t1, &p->y, &q->y); fe_mul(&t2, &p->z, &q->z);
    fe_add(&t3, &p->x, &p->y); fe_add(&t4, &q->x, &q->y); fe_mul(&t3, &t3, &t4);
    fe_add(&t4, &t0, &t1); fe_sub(&t3, &t3, &t4); fe_add(&t4, &p->y, &p->z);
    fe_add(&x3, &q->y, &q->z); fe_mul(&t4, &t4, &x3); fe_add(&x3, &t1, &t2);
    fe_sub(&t4, &t4, &x3); fe_add(&x3, &p->x, &p->z); fe_add(&y3, &q->x, &q->z);
    fe_mul(&x3, &x3, &y3); fe_add(&y3, &t0, &t2); fe_sub(&y3, &x3, &y3);
    fe_add(&x3, &t0, &t0); fe_add(&t0, &x3, &t0); fe_mul(&t2, &FE_B3, &t2);
    fe_add(&z3, &t1, &t2); fe_sub(&t1, &t1, &t2); fe_mul(&y3, &FE_B3, &y3);
    fe_mul(&x3, &t4, &y3); fe_mul(&t2, &t3, &t1); fe_sub(&x3, &t2, &x3);
    fe_mul(&y3, &y3, &t0); fe_mul(&t1, &t1, &z3); fe_add(&y3, &t1, &y3);
    fe_mul(&t0, &t0, &t3); fe_mul(&z3, &z3, &t4); fe_add(&z3, &z3, &t0);
    v->x = x3; v->y = y3; v->z = z3;
}
/* point_projective.go:123-205 -- RCB'16 Algorithm 8 (Z2 = 1; addend != inf). */
static void pt_add_mixed(pt *v, const pt *p, const fe *x2, const fe *y2) {
    fe t0, t1, t2, t3, t4, x3, y3, z3;
    fe_mul(&t0, &p->x, x2); fe_mul(&t1, &p->y, y2); fe_add(&t3, x2, y2);
    fe_add(&t4, &p->x, &p->y); fe_mul(&t3, &t3, &t4); fe_add(&t4, &t0, &t1);
    fe_sub(&t3, &t3, &t4); fe_mul(&t4, y2, &p->z); fe_add(&t4, &t4, &p->y);
    fe_mul(&y3, x2, &p->z); fe_add(&y3, &y3, &p->x); fe_add(&x3, &t0, &t0);
    fe_add(&t0, &x3, &t0); fe_mul(&t2, &FE_B3, &p->z); fe_add(&z3, &t1, &t2);
    fe_sub(&t1, &t1, &t2); fe_mul(&y3, &FE_B3, &y3); fe_mul(&x3, &t4, &y3);
    fe_mul(&t2, &t3, &t1); fe_sub(&x3, &t2, &x3); fe_mul(&y3, &y3, &t0);
    fe_mul(&t1, &t1, &z3); fe_add(&y3, &t1, &y3); fe_mul(&t0, &t0, &t3);
    fe_mul(&z3, &z3, &t4); fe_add(&z3, &z3, &t0);
    v->x = x3; v->y = y3; v->z = z3;
}
/* point_projective.go:208-273 -- RCB'16 Algorithm 9. */
static void pt_double(pt *v, const pt *p) {
    fe t0, t1, t2, x3, y3, z3;
    fe_sqr(&t0, &p->y); fe_add(&z3, &t0, &t0); fe_add(&z3, &z3, &z3);
    fe_add(&z3, &z3, &z3); fe_mul(&t1, &p->y, &p->z); fe_sqr(&t2, &p->z);
    fe_mul(&t2, &FE_B3, &t2); fe_mul(&x3, &t2, &z3); fe_add(&y3, &t0, &t2);
    fe_mul(&z3, &t1, &z3); fe_add(&t1, &t2, &t2); fe_add(&t2, &t1, &t2);
    fe_sub(&t0, &t0, &t2); fe_mul(&y3, &t0, &y3); fe_add(&y3, &x3, &y3);
    fe_mul(&t1, &p->x, &p->y); fe_mul(&x3, &t0, &t1); fe_add(&x3, &x3, &x3);
    v->x = x3; v->y = y3; v->z = z3;
}
/* point_projective.go:278-302 -- one Fermat inversion; identity -> (0,1,0). */
static void pt_rescale(pt *v, const pt *p) {
    if (pt_is_identity(p)) { pt_identity(v); return; }
    fe a;
    fe_invert(&a, &p->z);
    fe_mul(&v->x, &a, &p->x);
    fe_mul(&v->y, &a, &p->y);
    v->z = FE_ONE;
}
/* point.go:134-145 -- cross-multiplied equality. */
static int pt_eq(const pt *a, const pt *b) {
    fe l, r;
    fe_mul(&l, &a->x, &b->z); fe_mul(&r, &b->x, &a->z);
    if (!fe_eq(&l, &r)) return 0;
    fe_mul(&l, &a->y, &b->z); fe_mul(&r, &b->y, &a->z);
    return fe_eq(&l, &r);
}

/* point_s11n.go:302-307 -- x^3 + 7 */
static void maybe_yy(fe *yy, const fe *x) {
    fe_sqr(yy, x); fe_mul(yy, yy, x); fe_add(yy, yy, &FE_B);
}
/* Status codes shared with include/secp256k1_b200.h */
#define ST_INVALID 0
#define ST_OK 1
#define ST_IDENTITY 2

/* point_s11n.go:178-213 -- 04 || X || Y, canonical coords, on curve. */
static int pt_set_uncompressed(pt *v, const u8 b[65]) {
    if (b[0] != 0x04) return 0;
    fe x, y, yy, y2;
    if (!fe_set_canonical_bytes(&x, b + 1)) return 0;
    if (!fe_set_canonical_bytes(&y, b + 33)) return 0;
    maybe_yy(&yy, &x);
    fe_sqr(&y2, &y);
    if (!fe_eq(&yy, &y2)) return 0;
    v->x = x; v->y = y; v->z = FE_ONE;
    return 1;
}
/* point_s11n.go:140-172 -- 02/03 || X, sqrt, parity select. */
static int pt_set_compressed(pt *v, const u8 b[33]) {
    if (b[0] != 0x02 && b[0] != 0x03) return 0;
    fe x, y, yy, yn;
    if (!fe_set_canonical_bytes(&x, b + 1)) return 0;
    maybe_yy(&yy, &x);
    if (!fe_sqrt(&y, &yy)) return 0;
    fe_neg(&yn, &y);
    v->x = x;
    v->y = (fe_is_odd(&y) == (b[0] & 1)) ? y : yn;
    v->z = FE_ONE;
    return 1;
}
/* point_s11n.go:66-88 -- returns ST_IDENTITY (out zeroed) or ST_OK. */
static int pt_uncompressed_bytes(u8 out[65], const pt *p) {
    if (pt_is_identity(p)) { memset(out, 0, 65); return ST_IDENTITY; }
    pt s;
    pt_rescale(&s, p);
    out[0] = 0x04;
    fe_bytes(out + 1, &s.x);
    fe_bytes(out + 33, &s.y);
    return ST_OK;
}

/* ------------------------------------------------------------------------- */
/* Generator tables.  point_mul_table.go:68-100,147-160;                      */
/* internal/gentable/point_mul_table.go:16-49                                 */
/* ------------------------------------------------------------------------- */

static apt (*G_HUGE)[255]; /* [32][255]: (j+1) * 256^i * G */
static apt G_ODD[32][15];  /* [32][15]:  (j+1) * 16 * 256^i * G */

static void gen_tables(void) {
    G_HUGE = malloc(sizeof(apt) * 32 * 255);
    pt base = PT_G;
    for (int i = 0; i < 32; i++) {
        pt acc = base, s;
        for (int j = 0; j < 255; j++) {
            pt_rescale(&s, &acc);
            G_HUGE[i][j].x = s.x;
            G_HUGE[i][j].y = s.y;
            pt_add(&acc, &acc, &base);
        }
        for (int k = 0; k < 8; k++) pt_double(&base, &base);
    }
    for (int i = 0; i < 32; i++)
        for (int j = 0; j < 15; j++) G_ODD[i][j] = G_HUGE[i][(16 + (j << 4)) - 1];
}
/* Serialised like internal/gentable/point_mul_table.go:28-46 (BE X || Y). */
EXPORT void orc_gen_table_bytes(u8 *out /* 522240 */) {
    for (int i = 0; i < 32; i++)
        for (int j = 0; j < 255; j++) {
            fe_bytes(out, &G_HUGE[i][j].x); out += 32;
            fe_bytes(out, &G_HUGE[i][j].y); out += 32;
        }
}

/* point_mul_table.go:51-60 -- [1..15]P by 7 doublings + 7 additions. */
static void pt_table15(pt tbl[15], const pt *p) {
    tbl[0] = *p;
    for (int i = 1; i < 15; i += 2) {
        pt_double(&tbl[i], &tbl[i / 2]);
        pt_add(&tbl[i + 1], &tbl[i], p);
    }
}

/* ------------------------------------------------------------------------- */
/* Scalar multiplication                                                      */
/* ------------------------------------------------------------------------- */

/* point_mul_table.go:168-194 -- constant-time: 64 lookups + 64 mixed adds, the
 * idx == 0 case resolved by select (the oracle scans nothing: only values are
 * observable). */
static void pt_scalar_base_mult(pt *v, const sc *s) {
    u8 b[32];
    sc_bytes(b, s);
    pt_identity(v);
    for (int i = 0; i < 32; i++) {
        int ti = 31 - i;
        unsigned hi = b[i] >> 4, lo = b[i] & 0xF;
        pt tmp;
        if (hi) { pt_add_mixed(&tmp, v, &G_ODD[ti][hi - 1].x, &G_ODD[ti][hi - 1].y); *v = tmp; }
        if (lo) { pt_add_mixed(&tmp, v, &G_HUGE[ti][lo - 1].x, &G_HUGE[ti][lo - 1].y); *v = tmp; }
    }
}
/* point_mul_table.go:197-211 -- vartime, 8-bit windows. */
static void pt_scalar_base_mult_vartime(pt *v, const sc *s) {
    u8 b[32];
    sc_bytes(b, s);
    pt_identity(v);
    for (int i = 0; i < 32; i++) {
        if (!b[i]) continue;
        const apt *e = &G_HUGE[31 - i][b[i] - 1];
        pt_add_mixed(v, v, &e->x, &e->y);
    }
}

/* point_mul_glv.go:119-189 -- round(k * g / 2^384) via 4x4 schoolbook. */
static void sc_mul_g_floored_div(sc *r, const sc *k, const sc *g) {
    u64 a[4], b[4], c[8] = {0};
    sc_limbs(a, k);
    sc_limbs(b, g);
    for (int i = 0; i < 4; i++) {
        u64 u = 0;
        for (int j = 0; j < 4; j++) {
            u128 t = (u128)a[i] * b[j] + c[i + j] + u;
            c[i + j] = (u64)t;
            u = (u64)(t >> 64);
        }
        c[i + 4] = u;
    }
    u64 should_add = (c[5] >> 63) & 1;
    u64 l[4] = {c[6] + should_add, 0, 0, 0};
    l[1] = c[7] + (l[0] < should_add);
    sc_set_limbs(r, l);
}
/* point_mul_glv.go:59-117 */
static void sc_split_glv(sc *k1, sc *k2, const sc *k) {
    sc c1, c2, t;
    sc_mul_g_floored_div(&c1, k, &SC_G1);
    sc_mul_g_floored_div(&c2, k, &SC_G2);
    sc_mul(k2, &c1, &SC_NEG_B1);
    sc_mul(&t, &c2, &SC_NEG_B2);
    sc_add(k2, k2, &t);
    sc_mul(k1, k2, &SC_NEG_LAMBDA);
    sc_add(k1, k, k1);
}
/* point_mul_glv.go:191-200 */
static void pt_mul_beta(pt *v, const pt *p) { fe_mul(&v->x, &p->x, &FE_BETA); v->y = p->y; v->z = p->z; }

/* point_mul_glv.go:203-254 (vartime) and :257-303 (constant time): identical
 * values; the vartime flavour skips zero nibbles, the ct one adds the identity. */
static void pt_scalar_mult_glv(pt *v, const sc *s, const pt *p, int vartime) {
    pt pee = *p, pp;
    pt_mul_beta(&pp, p);
    sc k1, k2;
    sc_split_glv(&k1, &k2, s);
    if (sc_is_gt_half_n(&k1)) { sc_neg(&k1, &k1); pt_neg(&pee, &pee); }
    if (sc_is_gt_half_n(&k2)) { sc_neg(&k2, &k2); pt_neg(&pp, &pp); }
    pt t1[15], t2[15], id;
    pt_table15(t1, &pee);
    pt_table15(t2, &pp);
    pt_identity(&id);
    pt_identity(v);
    u8 b1[32], b2[32];
    sc_bytes(b1, &k1);
    sc_bytes(b2, &k2);
    for (int i = 16; i < 32; i++) {
        if (i != 16) for (int k = 0; k < 4; k++) pt_double(v, v);
        unsigned n1 = b1[i] >> 4, n2 = b2[i] >> 4;
        if (n1) pt_add(v, v, &t1[n1 - 1]); else if (!vartime) pt_add(v, v, &id);
        if (n2) pt_add(v, v, &t2[n2 - 1]); else if (!vartime) pt_add(v, v, &id);
        for (int k = 0; k < 4; k++) pt_double(v, v);
        n1 = b1[i] & 0xF; n2 = b2[i] & 0xF;
        if (n1) pt_add(v, v, &t1[n1 - 1]); else if (!vartime) pt_add(v, v, &id);
        if (n2) pt_add(v, v, &t2[n2 - 1]); else if (!vartime) pt_add(v, v, &id);
    }
}
/* point_mul_glv.go:307-317 */
static void pt_double_scalar_mult_basepoint_vartime(pt *v, const sc *u1, const sc *u2, const pt *p) {
    pt a, b;
    pt_scalar_base_mult_vartime(&a, u1);
    pt_scalar_mult_glv(&b, u2, p, 1);
    pt_add(v, &a, &b);
}
/* point_test.go:392-416 scalarMultTrivial -- bit-serial double-and-add, the
 * reference's own cross-check for every fast path. */
static void pt_scalar_mult_trivial(pt *v, const sc *s, const pt *p) {
    u8 b[32];
    sc_bytes(b, s);
    pt q;
    pt_identity(&q);
    for (int i = 0; i < 256; i++) {
        pt_double(&q, &q);
        if ((b[i / 8] >> (7 - (i % 8))) & 1) pt_add(&q, &q, p);
    }
    *v = q;
}
/* point_mul_multi.go:25-117 -- Straus with shared doublings; l == 1 -> GLV. */
static void pt_multi_scalar_mult(pt *v, const sc *scalars, const pt *points, size_t l, int vartime) {
    if (l == 1) { pt_scalar_mult_glv(v, &scalars[0], &points[0], vartime); return; }
    pt_identity(v);
    if (l == 0) return;
    pt (*tbls)[15] = malloc(sizeof(pt) * 15 * l);
    u8 (*sb)[32] = malloc(32 * l);
    pt id;
    pt_identity(&id);
    for (size_t j = 0; j < l; j++) { pt_table15(tbls[j], &points[j]); sc_bytes(sb[j], &scalars[j]); }
    for (int i = 0; i < 32; i++) {
        if (i) for (int k = 0; k < 4; k++) pt_double(v, v);
        for (size_t j = 0; j < l; j++) {
            unsigned nb = sb[j][i] >> 4;
            if (nb) pt_add(v, v, &tbls[j][nb - 1]); else if (!vartime) pt_add(v, v, &id);
        }
        for (int k = 0; k < 4; k++) pt_double(v, v);
        for (size_t j = 0; j < l; j++) {
            unsigned nb = sb[j][i] & 0xF;
            if (nb) pt_add(v, v, &tbls[j][nb - 1]); else if (!vartime) pt_add(v, v, &id);
        }
    }
    free(tbls);
    free(sb);
}

/* ------------------------------------------------------------------------- */
/* SHA-256 (FIPS 180-4) -- stands in for Go's crypto/sha256 used at           */
/* secec/bitcoin/schnorr.go:309-320.                                          */
/* ------------------------------------------------------------------------- */

static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
typedef struct { uint32_t h[8]; u8 buf[64]; u64 len; } sha256_ctx;
static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static void sha256_block(uint32_t h[8], const u8 *p) {
    uint32_t w[64], a, b, c, d, e, f, g, hh;
    for (int i = 0; i < 16; i++) w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    a = h[0]; b = h[1]; c = h[2]; d = h[3]; e = h[4]; f = h[5]; g = h[6]; hh = h[7];
    for (int i = 0; i < 64; i++) {
        uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + K256[i] + w[i];
        uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
static void sha256_init(sha256_ctx *c) {
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(c->h, iv, 32);
    c->len = 0;
}
static void sha256_update(sha256_ctx *c, const u8 *d, size_t n) {
    size_t fill = (size_t)(c->len & 63);
    c->len += n;
    if (fill) {
        size_t take = 64 - fill;
        if (take > n) take = n;
        memcpy(c->buf + fill, d, take);
        d += take; n -= take; fill += take;
        if (fill < 64) return;
        sha256_block(c->h, c->buf);
    }
    while (n >= 64) { sha256_block(c->h, d); d += 64; n -= 64; }
    if (n) memcpy(c->buf, d, n);
}
static void sha256_final(sha256_ctx *c, u8 out[32]) {
    u64 bits = c->len * 8;
    u8 pad = 0x80;
    sha256_update(c, &pad, 1);
    u8 z = 0;
    while ((c->len & 63) != 56) sha256_update(c, &z, 1);
    u8 lb[8];
    for (int i = 0; i < 8; i++) lb[i] = (u8)(bits >> (56 - 8 * i));
    sha256_update(c, lb, 8);
    for (int i = 0; i < 8; i++) { out[4 * i] = (u8)(c->h[i] >> 24); out[4 * i + 1] = (u8)(c->h[i] >> 16); out[4 * i + 2] = (u8)(c->h[i] >> 8); out[4 * i + 3] = (u8)c->h[i]; }
}
EXPORT void orc_sha256(const u8 *d, size_t n, u8 out[32]) {
    sha256_ctx c;
    sha256_init(&c);
    sha256_update(&c, d, n);
    sha256_final(&c, out);
}

/* ------------------------------------------------------------------------- */
/* Protocols                                                                  */
/* ------------------------------------------------------------------------- */

#define FLAG_REJECT_MALLEABLE 1u

/* secec/ecdsa.go:171-228 (EncodingCompact path) -> secec/s11n.go:129-145 ->
 * secec/ecdsa.go:392-470.  pk65 is decoded like secec.NewPublicKey
 * (secec/secec.go:188-216 -> point_s11n.go:178). */
static int ecdsa_verify_one(const u8 pk65[65], const u8 digest32[32], const u8 sig64[64], uint32_t flags) {
    pt q;
    if (!pt_set_uncompressed(&q, pk65)) return 0;
    sc r, s, e, sinv, u1, u2, v;
    if (!sc_set_canonical_bytes(&r, sig64) || sc_is_zero(&r)) return 0;
    if (!sc_set_canonical_bytes(&s, sig64 + 32) || sc_is_zero(&s)) return 0;
    if ((flags & FLAG_REJECT_MALLEABLE) && sc_is_gt_half_n(&s)) return 0;
    sc_set_bytes(&e, digest32); /* hashToScalar, ecdsa.go:477-486 */
    sc_invert(&sinv, &s);
    sc_mul(&u1, &e, &sinv);
    sc_mul(&u2, &r, &sinv);
    pt R;
    pt_double_scalar_mult_basepoint_vartime(&R, &u1, &u2, &q);
    if (pt_is_identity(&R)) return 0;
    pt sR;
    pt_rescale(&sR, &R);
    u8 xb[32];
    fe_bytes(xb, &sR.x);
    sc_set_bytes(&v, xb);
    return sc_eq(&v, &r);
}

/* point_s11n.go:245-282 */
static int pt_recover(pt *R, const sc *r, u8 rec_id) {
    if (rec_id >= 4) return 0;
    u8 rb[32], xb[33];
    sc_bytes(rb, r);
    fe x, xn;
    if (!fe_set_canonical_bytes(&x, rb)) return 0; /* cannot happen: n < p */
    unsigned x_gt_n = (rec_id >> 1) & 1;
    fe_add(&xn, &x, &FE_N);
    if (x_gt_n) x = xn;
    fe_bytes(xb + 1, &x);
    sc chk;
    int did_reduce = sc_set_bytes(&chk, xb + 1);
    if (!((unsigned)did_reduce == x_gt_n && sc_eq(&chk, r))) return 0;
    xb[0] = 0x02 + (rec_id & 1);
    return pt_set_compressed(R, xb);
}
/* secec/ecdsa.go:244-282; sig65 = r || s || v parsed like
 * secec/s11n.go:156-168. */
static int ecdsa_recover_one(const u8 digest32[32], const u8 sig65[65], u8 pk65[65]) {
    sc r, s, e, nege, rinv, u1, u2;
    memset(pk65, 0, 65);
    if (!sc_set_canonical_bytes(&r, sig65) || sc_is_zero(&r)) return ST_INVALID;
    if (!sc_set_canonical_bytes(&s, sig65 + 32) || sc_is_zero(&s)) return ST_INVALID;
    pt R, Q;
    if (!pt_recover(&R, &r, sig65[64])) return ST_INVALID;
    sc_set_bytes(&e, digest32);
    sc_neg(&nege, &e);
    sc_invert(&rinv, &r);
    sc_mul(&u1, &nege, &rinv);
    sc_mul(&u2, &s, &rinv);
    pt_double_scalar_mult_basepoint_vartime(&Q, &u1, &u2, &R);
    if (pt_is_identity(&Q)) return ST_INVALID; /* secec.go:206-209 */
    pt_uncompressed_bytes(pk65, &Q);
    return ST_OK;
}

/* secec/bitcoin/schnorr.go:221-253, :257-275 (lift_x), :420-478. */
static const u8 *bip340_challenge_midstate(void) {
    static u8 tag2[64];
    static int init;
    if (!init) {
        orc_sha256((const u8 *)"BIP0340/challenge", 17, tag2);
        memcpy(tag2 + 32, tag2, 32);
        init = 1;
    }
    return tag2;
}
static int schnorr_verify_one(const u8 pkx32[32], const u8 *msg, size_t msg_len, const u8 sig64[64]) {
    u8 cp[33];
    cp[0] = 0x02;
    memcpy(cp + 1, pkx32, 32);
    pt P;
    if (!pt_set_compressed(&P, cp)) return 0;
    fe rfe;
    if (!fe_set_canonical_bytes(&rfe, sig64)) return 0; /* r >= p */
    sc s, e;
    if (!sc_set_canonical_bytes(&s, sig64 + 32)) return 0; /* s >= n; zero allowed */
    sha256_ctx c;
    u8 eb[32];
    sha256_init(&c);
    sha256_update(&c, bip340_challenge_midstate(), 64);
    sha256_update(&c, sig64, 32);
    sha256_update(&c, pkx32, 32);
    sha256_update(&c, msg, msg_len);
    sha256_final(&c, eb);
    sc_set_bytes(&e, eb);
    sc_neg(&e, &e);
    pt R;
    pt_double_scalar_mult_basepoint_vartime(&R, &s, &e, &P);
    if (pt_is_identity(&R)) return 0;
    u8 ub[65];
    pt_uncompressed_bytes(ub, &R);
    if (ub[64] & 1) return 0;
    return memcmp(ub + 1, sig64, 32) == 0;
}

/* secec/ecdsa_k_rfc6979.go:36-145 -- HMAC_DRBG(SHA-256) keyed with int2octets(x) || bits2octets(h1);
 * every Read returns one 32-byte T = V after V = HMAC_K(V); a rejected T is followed by
 * K = HMAC_K(V || 0x00), V = HMAC_K(V) (delayed to the next Read). */
static void hmac_sha256(u8 out[32], const u8 key[32], const u8 *msg, size_t len) {
    u8 pad[64], inner[32];
    sha256_ctx c;
    memset(pad, 0x36, 64);
    for (int i = 0; i < 32; i++) pad[i] ^= key[i];
    sha256_init(&c); sha256_update(&c, pad, 64); sha256_update(&c, msg, len); sha256_final(&c, inner);
    memset(pad, 0x5c, 64);
    for (int i = 0; i < 32; i++) pad[i] ^= key[i];
    sha256_init(&c); sha256_update(&c, pad, 64); sha256_update(&c, inner, 32); sha256_final(&c, out);
}
typedef struct { u8 v[32], k[32]; int need_update; } drbg6979;
static void drbg_init(drbg6979 *g, const u8 x32[32], const u8 h32[32]) {
    u8 m[97];
    memset(g->v, 0x01, 32);
    memset(g->k, 0x00, 32);
    g->need_update = 0;
    for (int oct = 0; oct < 2; oct++) {
        memcpy(m, g->v, 32); m[32] = (u8)oct; memcpy(m + 33, x32, 32); memcpy(m + 65, h32, 32);
        hmac_sha256(g->k, g->k, m, 97);
        hmac_sha256(g->v, g->k, g->v, 32);
    }
}
static void drbg_read(drbg6979 *g, u8 out[32]) {
    if (g->need_update) {
        u8 m[33];
        memcpy(m, g->v, 32); m[32] = 0;
        hmac_sha256(g->k, g->k, m, 33);
        hmac_sha256(g->v, g->k, g->v, 32);
    }
    hmac_sha256(g->v, g->k, g->v, 32);
    memcpy(out, g->v, 32);
    g->need_update = 1;
}
/* secec/ecdsa.go:284-390 with rand = RFC6979SHA256() (:505-506): e = leftmost 32 digest bytes mod n,
 * k from the DRBG by rejection (sampleRandomScalar :523-545), R = k*G, r = x(R) mod n,
 * s = (r*d + e)/k, low-s normalisation, recovery id = (didReduce << 1 | yOdd) ^ negated.
 * priv32 must be canonical and non-zero (NewPrivateKey, secec/secec.go:141-160). */
static int ecdsa_sign_rfc6979_one(const u8 priv32[32], const u8 digest32[32], u8 sig64[64], u8 *recid) {
    sc d, e, k, r, s, kinv;
    memset(sig64, 0, 64);
    *recid = 0;
    if (!sc_set_canonical_bytes(&d, priv32) || sc_is_zero(&d)) return ST_INVALID;
    sc_set_bytes(&e, digest32);
    u8 db[32], eb[32], t[32];
    sc_bytes(db, &d);
    sc_bytes(eb, &e);
    drbg6979 g;
    drbg_init(&g, db, eb);
    for (;;) {
        int ok = 0;
        for (int i = 0; i < 8 && !ok; i++) {
            drbg_read(&g, t);
            ok = (sc_set_bytes(&k, t) == 0) && !sc_is_zero(&k);
        }
        if (!ok) return ST_INVALID;
        pt R, sR;
        pt_scalar_base_mult(&R, &k);
        pt_rescale(&sR, &R);
        u8 xb[32];
        fe_bytes(xb, &sR.x);
        int did_reduce = sc_set_bytes(&r, xb);
        if (sc_is_zero(&r)) continue;
        sc_invert(&kinv, &k);
        sc_mul(&s, &r, &d);
        sc_add(&s, &s, &e);
        sc_mul(&s, &s, &kinv);
        if (sc_is_zero(&s)) continue;
        int neg = sc_is_gt_half_n(&s);
        if (neg) sc_neg(&s, &s);
        *recid = (u8)(((did_reduce << 1) | fe_is_odd(&sR.y)) ^ neg);
        sc_bytes(sig64, &r);
        sc_bytes(sig64 + 32, &s);
        return ST_OK;
    }
}

/* secec/bitcoin/schnorr.go:309-320 */
static void tagged_hash3(u8 out[32], const char *tag, const u8 *a, size_t al, const u8 *b, size_t bl, const u8 *c, size_t cl) {
    u8 th[32];
    sha256_ctx x;
    orc_sha256((const u8 *)tag, strlen(tag), th);
    sha256_init(&x);
    sha256_update(&x, th, 32); sha256_update(&x, th, 32);
    if (al) sha256_update(&x, a, al);
    if (bl) sha256_update(&x, b, bl);
    if (cl) sha256_update(&x, c, cl);
    sha256_final(&x, out);
}
/* secec/bitcoin/schnorr.go:322-400 signSchnorr (+ NewSchnorrPrivateKeyFromECDSA :161-180 for d and P):
 * d' canonical non-zero; P = d'G; d = d' or n - d' (even y); t = d xor H_aux(aux);
 * k' = H_nonce(t || Px || m) mod n (zero is an error); R = k'G; k = k' or n - k';
 * e = H_challenge(Rx || Px || m) mod n; sig = Rx || (k + e d). */
static int schnorr_sign_one(const u8 priv32[32], const u8 *msg, size_t msg_len, const u8 aux32[32], u8 sig64[64]) {
    sc dp, d, kp, k, e, sum;
    memset(sig64, 0, 64);
    if (!sc_set_canonical_bytes(&dp, priv32) || sc_is_zero(&dp)) return ST_INVALID;
    pt P, R, sP, sR;
    pt_scalar_base_mult(&P, &dp);
    pt_rescale(&sP, &P);
    u8 px[32], rx[32], db[32], t[32], rnd[32], eb[32];
    fe_bytes(px, &sP.x);
    d = dp;
    if (fe_is_odd(&sP.y)) sc_neg(&d, &dp);
    sc_bytes(db, &d);
    tagged_hash3(t, "BIP0340/aux", aux32, 32, NULL, 0, NULL, 0);
    for (int i = 0; i < 32; i++) t[i] ^= db[i];
    tagged_hash3(rnd, "BIP0340/nonce", t, 32, px, 32, msg, msg_len);
    sc_set_bytes(&kp, rnd);
    if (sc_is_zero(&kp)) return ST_INVALID;
    pt_scalar_base_mult(&R, &kp);
    pt_rescale(&sR, &R);
    fe_bytes(rx, &sR.x);
    k = kp;
    if (fe_is_odd(&sR.y)) sc_neg(&k, &kp);
    tagged_hash3(eb, "BIP0340/challenge", rx, 32, px, 32, msg, msg_len);
    sc_set_bytes(&e, eb);
    sc_mul(&sum, &e, &d);
    sc_add(&sum, &k, &sum);
    memcpy(sig64, rx, 32);
    sc_bytes(sig64 + 32, &sum);
    return ST_OK;
}

/* ------------------------------------------------------------------------- */
/* Hash to curve (RFC 9380): secec/h2c/ (all .go files), point_h2c.go, internal/swu/swu.go  */
/* ------------------------------------------------------------------------- */
static void hex_to_limbs(u64 l[4], const char *h);
static fe H_A, H_B, H_Z, H_C2, H_K[4][4];
static void fe_from_hex(fe *r, const char *h) { u64 l[4]; hex_to_limbs(l, h); mont_to(r->v, l, &FP); }
static void h2c_init(void) {
    fe_from_hex(&H_A, "3f8731abdd661adca08a5558f0f5d272e953d363cb6f0e5d405447c01a444533");  /* swu.go feA */
    fe_set_u64(&H_B, 1771);
    fe t; fe_set_u64(&t, 11); fe_neg(&H_Z, &t);
    fe_from_hex(&H_C2, "31fdf302724013e57ad13fb38f842afeec184f00a74789dd286729c8303c4a59");  /* field_sqrt_ratio.go:10 */
    static const char *k[4][4] = {
        {"8e38e38e38e38e38e38e38e38e38e38e38e38e38e38e38e38e38e38daaaaa8c7", "07d3d4c80bc321d5b9f315cea7fd44c5d595d2fc0bf63b92dfff1044f17c6581",
         "534c328d23f234e6e2a413deca25caece4506144037c40314ecbd0b53d9dd262", "8e38e38e38e38e38e38e38e38e38e38e38e38e38e38e38e38e38e38daaaaa88c"},
        {"d35771193d94918a9ca34ccbb7b640dd86cd409542f8487d9fe6b745781eb49b", "edadc6f64383dc1df7c4b2d51b54225406d36b641f5e41bbc52a56612a8c6d14", 0, 0},
        {"4bda12f684bda12f684bda12f684bda12f684bda12f684bda12f684b8e38e23c", "c75e0c32d5cb7c0fa9d0a54b12a0a6d5647ab046d686da6fdffc90fc201d71a3",
         "29a6194691f91a73715209ef6512e576722830a201be2018a765e85a9ecee931", "2f684bda12f684bda12f684bda12f684bda12f684bda12f684bda12f38e38d84"},
        {"fffffffffffffffffffffffffffffffffffffffffffffffffffffffefffff93b", "7a06534bb8bdb49fd5e9e6632722c2989467c1bfc8e8d978dfb425d2685c2573",
         "6484aa716545ca2cf3a70c3fa8fe337e0a3d21162f0d6299a7bf8192bfd2a76f", 0}};
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) if (k[i][j]) fe_from_hex(&H_K[i][j], k[i][j]);
}
/* secec/h2c/h2c_expand_message.go:33-139 with SHA-256 */
static int expand_message_xmd(u8 *out, size_t len, const u8 *dst, size_t dst_len, const u8 *msg, size_t msg_len) {
    u8 dbuf[32], b0[32], bi[32], zpad[64] = {0}, x[3];
    if (len == 0 || len > 65535 || dst_len == 0) return 0;
    if (dst_len > 255) {
        sha256_ctx c; sha256_init(&c);
        sha256_update(&c, (const u8 *)"H2C-OVERSIZE-DST-", 17); sha256_update(&c, dst, dst_len); sha256_final(&c, dbuf);
        dst = dbuf; dst_len = 32;
    }
    size_t ell = (len + 31) / 32;
    if (ell > 255) return 0;
    u8 dl = (u8)dst_len;
    sha256_ctx c; sha256_init(&c);
    sha256_update(&c, zpad, 64); sha256_update(&c, msg, msg_len);
    x[0] = (u8)(len >> 8); x[1] = (u8)len; x[2] = 0; sha256_update(&c, x, 3);
    sha256_update(&c, dst, dst_len); sha256_update(&c, &dl, 1); sha256_final(&c, b0);
    sha256_init(&c); sha256_update(&c, b0, 32); x[0] = 1; sha256_update(&c, x, 1);
    sha256_update(&c, dst, dst_len); sha256_update(&c, &dl, 1); sha256_final(&c, bi);
    size_t off = 0;
    for (size_t i = 1; i <= ell; i++) {
        if (i > 1) {
            u8 t[32];
            for (int k = 0; k < 32; k++) t[k] = b0[k] ^ bi[k];
            sha256_init(&c); sha256_update(&c, t, 32); x[0] = (u8)i; sha256_update(&c, x, 1);
            sha256_update(&c, dst, dst_len); sha256_update(&c, &dl, 1); sha256_final(&c, bi);
        }
        size_t take = len - off < 32 ? len - off : 32;
        memcpy(out + off, bi, take);
        off += take;
    }
    return 1;
}
/* internal/field/field_reduce.go:24-64 for 48-byte inputs: big-endian integer mod p */
static void fe_set_wide48(fe *r, const u8 b[48]) {
    u8 lo[32], hi[32] = {0};
    memcpy(lo, b + 16, 32);
    memcpy(hi + 16, b, 16);
    fe l, h, two256;
    fe_set_bytes(&l, lo);
    fe_set_bytes(&h, hi);
    u64 d[4] = {0x1000003D1ULL, 0, 0, 0};  /* 2^256 mod p */
    mont_to(two256.v, d, &FP);
    fe_mul(&h, &h, &two256);
    fe_add(r, &l, &h);
}
/* internal/field/field_sqrt_ratio.go:25-63 (RFC 9380 F.2.1.2, p = 3 mod 4) */
static int fe_sqrt_ratio(fe *z, const fe *u, const fe *v) {
    fe tv1, tv2, tv3, y1, y2, x223, x22, x2;
    fe_sqr(&tv1, v); fe_mul(&tv2, u, v); fe_mul(&tv1, &tv1, &tv2);
    /* y1 = tv1^((p-3)/4): (p-3)/4 = 2^254 - 2^30 - 245 = 1^223 0 1^22 0000 10 11 */
    fe_pow_common(&x223, &x22, &x2, &tv1);
    fe t; fe_pow2k(&t, &x223, 23); fe_mul(&t, &t, &x22);
    fe_pow2k(&t, &t, 5); fe_mul(&t, &t, &tv1);     /* ...0000 1 */
    fe_pow2k(&t, &t, 3); fe_mul(&t, &t, &x2);      /* 0 11 */
    y1 = t;
    fe_mul(&y1, &y1, &tv2);
    fe_mul(&y2, &y1, &H_C2);
    fe_sqr(&tv3, &y1); fe_mul(&tv3, &tv3, v);
    int qr = fe_eq(&tv3, u);
    *z = qr ? y1 : y2;
    return qr;
}
/* internal/swu/swu.go:70-147 -- map_to_curve_simple_swu on E' (RFC 9380 F.2) */
static void swu_map(fe *x, fe *y, const fe *u) {
    fe tv1, tv2, tv3, tv4, tv5, tv6, y1, nt;
    fe_sqr(&tv1, u); fe_mul(&tv1, &H_Z, &tv1);
    fe_sqr(&tv2, &tv1); fe_add(&tv2, &tv2, &tv1);
    fe_add(&tv3, &tv2, &FE_ONE); fe_mul(&tv3, &H_B, &tv3);
    int sel = fe_is_zero(&tv2);
    fe_neg(&nt, &tv2);
    tv4 = sel ? H_Z : nt;
    fe_mul(&tv4, &H_A, &tv4);
    fe_sqr(&tv2, &tv3); fe_sqr(&tv6, &tv4); fe_mul(&tv5, &H_A, &tv6);
    fe_add(&tv2, &tv2, &tv5); fe_mul(&tv2, &tv2, &tv3);
    fe_mul(&tv6, &tv6, &tv4); fe_mul(&tv5, &H_B, &tv6); fe_add(&tv2, &tv2, &tv5);
    fe_mul(x, &tv1, &tv3);
    int is_sq = fe_sqrt_ratio(&y1, &tv2, &tv6);
    fe_mul(y, &tv1, u); fe_mul(y, y, &y1);
    if (is_sq) { *x = tv3; *y = y1; }
    if (fe_is_odd(u) != fe_is_odd(y)) fe_neg(y, y);
    fe_invert(&tv4, &tv4);
    fe_mul(x, x, &tv4);
}
/* internal/swu/swu.go:149-199 -- 3-isogeny E' -> E; 0 if a denominator vanishes */
static int iso_map(fe *xo, fe *yo, const fe *X, const fe *Y) {
    fe XX, XXX, xn, xd, yn, yd, t;
    fe_sqr(&XX, X); fe_mul(&XXX, &XX, X);
    fe_mul(&xn, &H_K[0][3], &XXX); fe_mul(&t, &H_K[0][2], &XX); fe_add(&xn, &xn, &t);
    fe_mul(&t, &H_K[0][1], X); fe_add(&xn, &xn, &t); fe_add(&xn, &xn, &H_K[0][0]);
    fe_mul(&xd, &H_K[1][1], X); fe_add(&xd, &xd, &XX); fe_add(&xd, &xd, &H_K[1][0]);
    int xz = fe_is_zero(&xd);
    fe_invert(&xd, &xd); fe_mul(xo, &xn, &xd);
    fe_mul(&yn, &H_K[2][3], &XXX); fe_mul(&t, &H_K[2][2], &XX); fe_add(&yn, &yn, &t);
    fe_mul(&t, &H_K[2][1], X); fe_add(&yn, &yn, &t); fe_add(&yn, &yn, &H_K[2][0]);
    fe_mul(&yd, &H_K[3][2], &XX); fe_mul(&t, &H_K[3][1], X); fe_add(&yd, &yd, &t);
    fe_add(&yd, &yd, &XXX); fe_add(&yd, &yd, &H_K[3][0]);
    int yz = fe_is_zero(&yd);
    fe_invert(&yd, &yd); fe_mul(yo, &yn, &yd); fe_mul(yo, Y, yo);
    return !(xz | yz);
}
/* point_h2c.go:23-55 SetUniformBytes (48 bytes) */
static void pt_set_uniform48(pt *v, const u8 b[48]) {
    fe u, xp, yp, x, y;
    fe_set_wide48(&u, b);
    swu_map(&xp, &yp, &u);
    if (iso_map(&x, &y, &xp, &yp)) { v->x = x; v->y = y; v->z = FE_ONE; } else pt_identity(v);
}
/* secec/h2c/h2c.go:25-63 */
static int hash_to_curve_one(const u8 *dst, size_t dst_len, const u8 *msg, size_t msg_len, int ro, u8 out65[65]) {
    u8 ub[96];
    memset(out65, 0, 65);
    if (!expand_message_xmd(ub, ro ? 96 : 48, dst, dst_len, msg, msg_len)) return ST_INVALID;
    pt q0, q1, r;
    pt_set_uniform48(&q0, ub);
    if (ro) { pt_set_uniform48(&q1, ub + 48); pt_add(&r, &q0, &q1); } else r = q0;
    return pt_uncompressed_bytes(out65, &r);
}

/* ------------------------------------------------------------------------- */
/* Exported single-item entry points (ctypes)                                 */
/* ------------------------------------------------------------------------- */

static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static void hex_to_limbs(u64 l[4], const char *h) {
    u8 b[32];
    for (int i = 0; i < 32; i++) {
        unsigned v = 0;
        for (int k = 0; k < 2; k++) {
            char ch = h[2 * i + k];
            v = v * 16 + (unsigned)(ch <= '9' ? ch - '0' : (ch | 32) - 'a' + 10);
        }
        b[i] = (u8)v;
    }
    be32_to_limbs(l, b);
}
static void do_init(void) {
    mont_ctx_init(&FP, P_LIMBS);
    mont_ctx_init(&FN, N_LIMBS);
    memset(&FE_ZERO, 0, sizeof FE_ZERO);
    memcpy(FE_ONE.v, FP.one, 32);
    fe_set_u64(&FE_B, 7);   /* point.go:18-21 */
    fe_set_u64(&FE_B3, 21); /* point_projective.go:21 */
    u64 l[4];
    hex_to_limbs(l, "7ae96a2b657c07106e64479eac3434e99cf0497512f58995c1396c28719501ee"); /* point_mul_glv.go:44 */
    mont_to(FE_BETA.v, l, &FP);
    mont_to(FE_N.v, N_LIMBS, &FP); /* point_s11n.go feN */
    memset(&SC_ZERO, 0, sizeof SC_ZERO);
    memcpy(SC_ONE.v, FN.one, 32);
    hex_to_limbs(l, "ac9c52b33fa3cf1f5ad9e3fd77ed9ba4a880b9fc8ec739c2e0cfc810b51283cf"); sc_set_limbs(&SC_NEG_LAMBDA, l); /* :41 */
    hex_to_limbs(l, "00000000000000000000000000000000e4437ed6010e88286f547fa90abfe4c3"); sc_set_limbs(&SC_NEG_B1, l);     /* :47 */
    hex_to_limbs(l, "fffffffffffffffffffffffffffffffe8a280ac50774346dd765cda83db1562c"); sc_set_limbs(&SC_NEG_B2, l);     /* :50 */
    hex_to_limbs(l, "3086d221a7d46bcde86c90e49284eb153daa8a1471e8ca7fe893209a45dbb031"); sc_set_limbs(&SC_G1, l);         /* :53 */
    hex_to_limbs(l, "e4437ed6010e88286f547fa90abfe4c4221208ac9df506c61571b4ae8ac47f71"); sc_set_limbs(&SC_G2, l);         /* :56 */
    hex_to_limbs(l, "79be667ef9dcbbac55a06295ce870b07029bfcdb2dce28d959f2815b16f81798"); mont_to(PT_G.x.v, l, &FP);
    hex_to_limbs(l, "483ada7726a3c4655da4fbfc0e1108a8fd17b448a68554199c47d08ffb10d4b8"); mont_to(PT_G.y.v, l, &FP);
    PT_G.z = FE_ONE;
    h2c_init();
    gen_tables();
}
EXPORT void orc_init(void) { pthread_once(&g_once, do_init); }

EXPORT int orc_fe_set_bytes(const u8 in[32], u8 out[32]) { orc_init(); fe a; int d = fe_set_bytes(&a, in); fe_bytes(out, &a); return d; }
EXPORT int orc_fe_bytes_are_canonical(const u8 in[32]) { orc_init(); fe a; return fe_set_canonical_bytes(&a, in); }
EXPORT void orc_fe_mul(const u8 a[32], const u8 b[32], u8 out[32]) { orc_init(); fe x, y; fe_set_bytes(&x, a); fe_set_bytes(&y, b); fe_mul(&x, &x, &y); fe_bytes(out, &x); }
EXPORT void orc_fe_invert(const u8 a[32], u8 out[32]) { orc_init(); fe x; fe_set_bytes(&x, a); fe_invert(&x, &x); fe_bytes(out, &x); }
EXPORT int orc_fe_sqrt(const u8 a[32], u8 out[32]) { orc_init(); fe x; fe_set_bytes(&x, a); int ok = fe_sqrt(&x, &x); fe_bytes(out, &x); return ok; }
EXPORT int orc_sc_set_bytes(const u8 in[32], u8 out[32]) { orc_init(); sc a; int d = sc_set_bytes(&a, in); sc_bytes(out, &a); return d; }
EXPORT int orc_sc_bytes_are_canonical(const u8 in[32]) { orc_init(); sc a; return sc_set_canonical_bytes(&a, in); }
EXPORT int orc_sc_is_gt_half_n(const u8 in[32]) { orc_init(); sc a; sc_set_bytes(&a, in); return sc_is_gt_half_n(&a); }
EXPORT void orc_sc_mul(const u8 a[32], const u8 b[32], u8 out[32]) { orc_init(); sc x, y; sc_set_bytes(&x, a); sc_set_bytes(&y, b); sc_mul(&x, &x, &y); sc_bytes(out, &x); }
EXPORT void orc_sc_add(const u8 a[32], const u8 b[32], u8 out[32]) { orc_init(); sc x, y; sc_set_bytes(&x, a); sc_set_bytes(&y, b); sc_add(&x, &x, &y); sc_bytes(out, &x); }
EXPORT void orc_sc_invert(const u8 a[32], u8 out[32]) { orc_init(); sc x; sc_set_bytes(&x, a); sc_invert(&x, &x); sc_bytes(out, &x); }
EXPORT void orc_sc_split_glv(const u8 k[32], u8 k1[32], u8 k2[32]) { orc_init(); sc a, b, c; sc_set_bytes(&a, k); sc_split_glv(&b, &c, &a); sc_bytes(k1, &b); sc_bytes(k2, &c); }

/* Point decode: len 33 or 65 (SetBytes, point_s11n.go:218-231; the 1-byte
 * identity encoding is handled by callers).  Returns ST_OK / ST_INVALID. */
static int decode_point(pt *p, const u8 *b, size_t len) {
    if (len == 65) return pt_set_uncompressed(p, b);
    if (len == 33) return pt_set_compressed(p, b);
    return 0;
}
EXPORT int orc_point_decode(const u8 *in, size_t len, u8 out65[65]) {
    orc_init();
    pt p;
    memset(out65, 0, 65);
    if (len == 1) return in[0] == 0x00 ? ST_IDENTITY : ST_INVALID; /* point_s11n.go:211-217 */
    if (!decode_point(&p, in, len)) return ST_INVALID;
    return pt_uncompressed_bytes(out65, &p);
}
EXPORT int orc_scalar_base_mult(const u8 k32[32], u8 out65[65]) {
    orc_init();
    sc k; pt v;
    sc_set_bytes(&k, k32);
    pt_scalar_base_mult(&v, &k);
    return pt_uncompressed_bytes(out65, &v);
}
EXPORT int orc_scalar_base_mult_vartime(const u8 k32[32], u8 out65[65]) {
    orc_init();
    sc k; pt v;
    sc_set_bytes(&k, k32);
    pt_scalar_base_mult_vartime(&v, &k);
    return pt_uncompressed_bytes(out65, &v);
}
/* mode: 0 = ct GLV (ScalarMult), 1 = vartime GLV, 2 = bit-serial trivial */
EXPORT int orc_scalar_mult(const u8 k32[32], const u8 pt65[65], int mode, u8 out65[65]) {
    orc_init();
    sc k; pt p, v;
    memset(out65, 0, 65);
    if (!pt_set_uncompressed(&p, pt65)) return ST_INVALID;
    sc_set_bytes(&k, k32);
    if (mode == 2) pt_scalar_mult_trivial(&v, &k, &p);
    else pt_scalar_mult_glv(&v, &k, &p, mode);
    return pt_uncompressed_bytes(out65, &v);
}
/* secec/secec.go:53-56 -- x(k * P); identity -> error. */
EXPORT int orc_ecdh(const u8 k32[32], const u8 pt65[65], u8 x32[32]) {
    u8 out[65];
    int st = orc_scalar_mult(k32, pt65, 0, out);
    memset(x32, 0, 32);
    if (st != ST_OK) return st == ST_IDENTITY ? ST_IDENTITY : ST_INVALID;
    memcpy(x32, out + 1, 32);
    return ST_OK;
}
EXPORT int orc_double_scalar_mult_basepoint_vartime(const u8 u1[32], const u8 u2[32], const u8 pt65[65], u8 out65[65]) {
    orc_init();
    sc a, b; pt p, v;
    memset(out65, 0, 65);
    if (!pt_set_uncompressed(&p, pt65)) return ST_INVALID;
    sc_set_bytes(&a, u1);
    sc_set_bytes(&b, u2);
    pt_double_scalar_mult_basepoint_vartime(&v, &a, &b, &p);
    return pt_uncompressed_bytes(out65, &v);
}
EXPORT int orc_point_add(const u8 a65[65], int a_st, const u8 b65[65], int b_st, u8 out65[65]) {
    orc_init();
    pt a, b, v;
    pt_identity(&a); pt_identity(&b);
    if (a_st == ST_OK && !pt_set_uncompressed(&a, a65)) return ST_INVALID;
    if (b_st == ST_OK && !pt_set_uncompressed(&b, b65)) return ST_INVALID;
    pt_add(&v, &a, &b);
    return pt_uncompressed_bytes(out65, &v);
}
EXPORT int orc_msm(const u8 *k32, const u8 *pt65, size_t n, int vartime, u8 out65[65]) {
    orc_init();
    sc *ks = malloc(sizeof(sc) * (n ? n : 1));
    pt *ps = malloc(sizeof(pt) * (n ? n : 1));
    int st = ST_OK;
    memset(out65, 0, 65);
    for (size_t i = 0; i < n; i++) {
        sc_set_bytes(&ks[i], k32 + 32 * i);
        if (!pt_set_uncompressed(&ps[i], pt65 + 65 * i)) { st = ST_INVALID; break; }
    }
    if (st == ST_OK) {
        pt v;
        pt_multi_scalar_mult(&v, ks, ps, n, vartime);
        st = pt_uncompressed_bytes(out65, &v);
    }
    free(ks); free(ps);
    return st;
}
EXPORT int orc_ecdsa_verify(const u8 pk65[65], const u8 digest32[32], const u8 sig64[64], uint32_t flags) {
    orc_init();
    return ecdsa_verify_one(pk65, digest32, sig64, flags);
}
EXPORT int orc_ecdsa_recover(const u8 digest32[32], const u8 sig65[65], u8 pk65[65]) {
    orc_init();
    return ecdsa_recover_one(digest32, sig65, pk65);
}
EXPORT int orc_ecdsa_sign_rfc6979(const u8 priv32[32], const u8 digest32[32], u8 sig64[64], u8 *recid) {
    orc_init();
    return ecdsa_sign_rfc6979_one(priv32, digest32, sig64, recid);
}
EXPORT void orc_hmac_sha256(const u8 key[32], const u8 *msg, size_t len, u8 out[32]) { hmac_sha256(out, key, msg, len); }
EXPORT int orc_hash_to_curve(const u8 *dst, size_t dst_len, const u8 *msg, size_t msg_len, int ro, u8 out65[65]) {
    orc_init();
    return hash_to_curve_one(dst, dst_len, msg, msg_len, ro, out65);
}
EXPORT int orc_expand_message_xmd(const u8 *dst, size_t dst_len, const u8 *msg, size_t msg_len, u8 *out, size_t len) {
    return expand_message_xmd(out, len, dst, dst_len, msg, msg_len);
}
EXPORT int orc_map_to_curve(const u8 u48[48], u8 out65[65]) {
    orc_init();
    pt q;
    pt_set_uniform48(&q, u48);
    return pt_uncompressed_bytes(out65, &q);
}
EXPORT int orc_schnorr_sign(const u8 priv32[32], const u8 *msg, size_t msg_len, const u8 aux32[32], u8 sig64[64]) {
    orc_init();
    return schnorr_sign_one(priv32, msg, msg_len, aux32, sig64);
}
EXPORT int orc_schnorr_verify(const u8 pkx32[32], const u8 *msg, size_t msg_len, const u8 sig64[64]) {
    orc_init();
    return schnorr_verify_one(pkx32, msg, msg_len, sig64);
}

/* ------------------------------------------------------------------------- */
/* Threaded batch drivers (parity at size, and the reported CPU baseline).    */
/* Items are split into contiguous slices, one per thread.                    */
/* ------------------------------------------------------------------------- */

enum { OP_SBM, OP_SBM_VT, OP_SMUL, OP_ECDH, OP_DSM, OP_VERIFY, OP_RECOVER, OP_SCHNORR, OP_SIGN };
typedef struct {
    int op; size_t lo, hi;
    const u8 *a, *b, *c; size_t msg_len; uint32_t flags;
    u8 *out, *status;
} job_t;
static void *job_run(void *arg) {
    job_t *j = arg;
    for (size_t i = j->lo; i < j->hi; i++) {
        switch (j->op) {
        case OP_SBM: j->status[i] = (u8)orc_scalar_base_mult(j->a + 32 * i, j->out + 65 * i); break;
        case OP_SBM_VT: j->status[i] = (u8)orc_scalar_base_mult_vartime(j->a + 32 * i, j->out + 65 * i); break;
        case OP_SMUL: j->status[i] = (u8)orc_scalar_mult(j->a + 32 * i, j->b + 65 * i, 0, j->out + 65 * i); break;
        case OP_ECDH: j->status[i] = (u8)orc_ecdh(j->a + 32 * i, j->b + 65 * i, j->out + 32 * i); break;
        case OP_DSM: j->status[i] = (u8)orc_double_scalar_mult_basepoint_vartime(j->a + 32 * i, j->b + 32 * i, j->c + 65 * i, j->out + 65 * i); break;
        case OP_VERIFY: j->status[i] = (u8)ecdsa_verify_one(j->a + 65 * i, j->b + 32 * i, j->c + 64 * i, j->flags); break;
        case OP_RECOVER: j->status[i] = (u8)ecdsa_recover_one(j->a + 32 * i, j->b + 65 * i, j->out + 65 * i); break;
        case OP_SIGN: j->status[i] = (u8)ecdsa_sign_rfc6979_one(j->a + 32 * i, j->b + 32 * i, j->out + 64 * i, (u8 *)j->c + i); break;
        case OP_SCHNORR: j->status[i] = (u8)schnorr_verify_one(j->a + 32 * i, j->b + j->msg_len * i, j->msg_len, j->c + 64 * i); break;
        }
    }
    return NULL;
}
static void run_batch(job_t proto, size_t n, int nthreads) {
    orc_init();
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
    job_t *jobs = malloc(sizeof(job_t) * nthreads);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = proto;
        jobs[t].lo = n * t / nthreads;
        jobs[t].hi = n * (t + 1) / nthreads;
        pthread_create(&th[t], NULL, job_run, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}
EXPORT void orc_batch_scalar_base_mult(const u8 *k32, size_t n, int vartime, u8 *out65, u8 *status, int nthreads) {
    job_t j = {0}; j.op = vartime ? OP_SBM_VT : OP_SBM; j.a = k32; j.out = out65; j.status = status;
    run_batch(j, n, nthreads);
}
EXPORT void orc_batch_scalar_mult(const u8 *k32, const u8 *pt65, size_t n, u8 *out65, u8 *status, int nthreads) {
    job_t j = {0}; j.op = OP_SMUL; j.a = k32; j.b = pt65; j.out = out65; j.status = status;
    run_batch(j, n, nthreads);
}
EXPORT void orc_batch_ecdh(const u8 *k32, const u8 *pt65, size_t n, u8 *x32, u8 *status, int nthreads) {
    job_t j = {0}; j.op = OP_ECDH; j.a = k32; j.b = pt65; j.out = x32; j.status = status;
    run_batch(j, n, nthreads);
}
EXPORT void orc_batch_double_scalar_mult(const u8 *u1, const u8 *u2, const u8 *pt65, size_t n, u8 *out65, u8 *status, int nthreads) {
    job_t j = {0}; j.op = OP_DSM; j.a = u1; j.b = u2; j.c = pt65; j.out = out65; j.status = status;
    run_batch(j, n, nthreads);
}
EXPORT void orc_batch_ecdsa_verify(const u8 *pk65, const u8 *digest32, const u8 *sig64, uint32_t flags, size_t n, u8 *ok, int nthreads) {
    job_t j = {0}; j.op = OP_VERIFY; j.a = pk65; j.b = digest32; j.c = sig64; j.flags = flags; j.status = ok;
    run_batch(j, n, nthreads);
}
EXPORT void orc_batch_ecdsa_recover(const u8 *digest32, const u8 *sig65, size_t n, u8 *pk65, u8 *status, int nthreads) {
    job_t j = {0}; j.op = OP_RECOVER; j.a = digest32; j.b = sig65; j.out = pk65; j.status = status;
    run_batch(j, n, nthreads);
}
EXPORT void orc_batch_schnorr_verify(const u8 *pkx32, const u8 *msg, size_t msg_len, const u8 *sig64, size_t n, u8 *ok, int nthreads) {
    job_t j = {0}; j.op = OP_SCHNORR; j.a = pkx32; j.b = msg; j.msg_len = msg_len; j.c = sig64; j.status = ok;
    run_batch(j, n, nthreads);
}
EXPORT void orc_batch_ecdsa_sign_rfc6979(const u8 *priv32, const u8 *digest32, size_t n, u8 *sig64, u8 *recid, u8 *status, int nthreads) {
    job_t j = {0}; j.op = OP_SIGN; j.a = priv32; j.b = digest32; j.c = recid; j.out = sig64; j.status = status;
    run_batch(j, n, nthreads);
}
