// ecg_microbench.cuh — integer-pipe microbenchmark kernels behind ecg_microbench() (roofline denominators).
#pragma once
#include "ecg_kernels.cuh"

// ------------------------------------------------------------------------------------------------
// integer-pipe microbenchmarks (roofline denominators)
__global__ void __launch_bounds__(256) mb_imad_wide_kernel(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3 + 1, b0 = seed ^ 0x9E3779B9u, b1 = b0 + blockIdx.x;
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = a0 + i;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    // 4 independent chains of 4 IMAD.WIDE.U32.X each = 16 per iteration, x4 unrolled = 64
#pragma unroll
    for (int u = 0; u < 4; u++) {
      mad_wide_cc(r[0], r[1], a0, b0);
      madc_wide_cc(r[2], r[3], a1, b0);
      madc_wide_cc(r[4], r[5], a0, b1);
      madc_wide_cc(r[6], r[7], a1, b1);
      mad_wide_cc(r[8], r[9], a1, b0);
      madc_wide_cc(r[10], r[11], a0, b1);
      madc_wide_cc(r[12], r[13], a1, b1);
      madc_wide_cc(r[14], r[15], a0, b0);
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s ^= r[i];
  if (s == 0x12345678u) out[0] = s;
}
__global__ void __launch_bounds__(256) mb_imad_kernel(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a = seed + threadIdx.x, b = seed ^ 0x9E3779B9u;
  uint32_t r[8];
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = a + i;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < 8; i++) r[i] = r[i] * a + b;
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= r[i];
  if (s == 0x12345678u) out[0] = s;
}
__global__ void __launch_bounds__(256) mb_iadd_kernel(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a = seed + threadIdx.x;
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = a + i;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      r[0] = add_cc(r[0], r[8]);
#pragma unroll
      for (int i = 1; i < 8; i++) r[i] = addc_cc(r[i], r[8 + i]);
      r[8] = add_cc(r[8], r[1]);
#pragma unroll
      for (int i = 1; i < 8; i++) r[8 + i] = addc_cc(r[8 + i], r[(i + 1) & 7]);
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s ^= r[i];
  if (s == 0x12345678u) out[0] = s;
}
template <class F>
__global__ void __launch_bounds__(256) mb_fmul_kernel(uint32_t* out, int iters, uint32_t seed) {
  Fe a, b;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a.v[i] = seed * (i + 1) + threadIdx.x;
    b.v[i] = (seed ^ 0x9E3779B9u) * (i + 3) + blockIdx.x;
  }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    F::mul(a, a, b);
    F::mul(b, b, a);
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= a.v[i] ^ b.v[i];
  if (s == 0x12345678u) out[0] = s;
}

