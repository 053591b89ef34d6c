// Host build of the FP64-pipe field multiplication experiment (csrc/fe52.cuh) for
// tests/test_fe52_host.py: the same source, fma() under FE_TOWARDZERO standing in for fma.rz.f64.
#include <fenv.h>
#include "../../secp256k1-voi_b200/csrc/fe52.cuh"

extern "C" void fe52_mul_host(const uint64_t *a, const uint64_t *b, uint64_t *r, size_t n) {
    const int old = fegetround();
    fesetround(FE_TOWARDZERO);
    for (size_t i = 0; i < n; i++) {
        s256::fe52 x, y, z;
        for (int k = 0; k < 5; k++) {
            x.v[k] = (double)a[5 * i + k];
            y.v[k] = (double)b[5 * i + k];
        }
        s256::fe52_mul(z, x, y);
        for (int k = 0; k < 5; k++) r[5 * i + k] = (uint64_t)z.v[k];
    }
    fesetround(old);
}
