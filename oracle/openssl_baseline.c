/* openssl_baseline.c -- MEASUREMENT INFRASTRUCTURE, not product code and not the oracle.
 *
 * BASELINE.md section 3.3 asks for an OpenSSL ECDSA_do_verify throughput on the GPU box's host cores beside the port of
 * the reference's algorithm, as a sanity anchor for that port.  This file drives the system libcrypto (OpenSSL 3.0.x in
 * this image) over a batch of (public key, digest, compact signature) rows with a pthread pool.  Keys and signatures are
 * parsed into EC_KEY / ECDSA_SIG objects BEFORE the timed region (the most generous reading for the CPU); the timed
 * region is ECDSA_do_verify only.  Only bench.py's cpu_baseline leg loads it; nothing under secp256k1-voi_b200/ does.
 */
#define OPENSSL_SUPPRESS_DEPRECATED 1
#include <openssl/bn.h>
#include <openssl/ec.h>
#include <openssl/ecdsa.h>
#include <openssl/obj_mac.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <time.h>

typedef struct {
    EC_KEY **keys;
    ECDSA_SIG **sigs;
    const uint8_t *digest32;
    uint8_t *ok;
    size_t lo, hi;
    int reps;
} job_t;

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    for (int r = 0; r < j->reps; r++)
        for (size_t i = j->lo; i < j->hi; i++)
            j->ok[i] = (j->keys[i] && j->sigs[i] && ECDSA_do_verify(j->digest32 + 32 * i, 32, j->sigs[i], j->keys[i]) == 1);
    return NULL;
}

static double now(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

/* returns 0 on success; *seconds = wall time of `reps` passes of ECDSA_do_verify over the n rows on `threads` threads */
int ossl_batch_ecdsa_verify(const uint8_t *pk65, const uint8_t *digest32, const uint8_t *sig64, size_t n, int reps,
                            int threads, uint8_t *ok, double *seconds) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    EC_GROUP *grp = EC_GROUP_new_by_curve_name(NID_secp256k1);
    if (!grp) return -1;
    EC_KEY **keys = (EC_KEY **)calloc(n ? n : 1, sizeof(*keys));
    ECDSA_SIG **sigs = (ECDSA_SIG **)calloc(n ? n : 1, sizeof(*sigs));
    if (!keys || !sigs) return -2;
    for (size_t i = 0; i < n; i++) {
        EC_KEY *k = EC_KEY_new();
        EC_POINT *p = EC_POINT_new(grp);
        if (k && p && EC_KEY_set_group(k, grp) == 1 && EC_POINT_oct2point(grp, p, pk65 + 65 * i, 65, NULL) == 1 &&
            EC_KEY_set_public_key(k, p) == 1) {
            keys[i] = k;
        } else if (k) {
            EC_KEY_free(k);
        }
        if (p) EC_POINT_free(p);
        BIGNUM *r = BN_bin2bn(sig64 + 64 * i, 32, NULL), *s = BN_bin2bn(sig64 + 64 * i + 32, 32, NULL);
        ECDSA_SIG *sg = ECDSA_SIG_new();
        if (sg && r && s && ECDSA_SIG_set0(sg, r, s) == 1) {
            sigs[i] = sg;
        } else {
            if (sg) ECDSA_SIG_free(sg);
            if (r) BN_free(r);
            if (s) BN_free(s);
        }
    }
    pthread_t th[256];
    job_t jobs[256];
    double t0 = now();
    for (int t = 0; t < threads; t++) {
        jobs[t] = (job_t){keys, sigs, digest32, ok, n * (size_t)t / (size_t)threads, n * (size_t)(t + 1) / (size_t)threads, reps};
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    *seconds = now() - t0;
    for (size_t i = 0; i < n; i++) {
        if (keys[i]) EC_KEY_free(keys[i]);
        if (sigs[i]) ECDSA_SIG_free(sigs[i]);
    }
    free(keys);
    free(sigs);
    EC_GROUP_free(grp);
    return 0;
}
