// launchers.h -- host-callable launch functions of kernels that live in their own
// translation unit (kern_ct.cu is compiled with fe_mul inlined: its single
// add-site loop fits the instruction cache and measured 11 % faster that way).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "point.cuh"
#include "sc.cuh"

void s256_ct_kernels_init();
void s256_launch_base_mult_ct(const uint8_t *k32, size_t n, const s256::apt *tab_big, const s256::apt *tab_small,
                              const s256::apt *tab_huge, s256::pt *res, cudaStream_t s);
