// Shared helpers for libfrtm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/frtm_b200.h"

namespace frtm {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

// Every launch goes through this so failures surface as FRTM_ELAUNCH with context and launches are counted.
#define FRTM_CHECK_LAUNCH(name)                                                     \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      frtm::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));      \
      return FRTM_ELAUNCH;                                                          \
    }                                                                               \
    frtm::count_launch();                                                           \
  } while (0)

#define FRTM_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      frtm::set_error(__VA_ARGS__);        \
      return FRTM_EINVAL;                  \
    }                                      \
  } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum, result valid in every thread. `red` = shared float[32].
__device__ __forceinline__ float block_sum(float v, float *red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += red[i];  // fixed order -> deterministic
  return t;
}

// ATen's bilinear source index (align_corners=False): upsample_bilinear2d / area_pixel_compute_source_index.
__device__ __forceinline__ void bilinear_src(int dst, float scale, int in_size, int &i0, int &i1, float &lam) {
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  lam = src - (float)i0;
}

}  // namespace frtm
