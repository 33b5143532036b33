// Types shared by the GN/CG operator kernels (target_model.cu: CUDA-core kernels, gn_apply_tc.cu: tcgen05 kernel).
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>

namespace frtm {

// Pointer table for object-batched launches: rows {samples, stencil, uty, weights, filt, cg_state, gate_count,
// samples_split} x n_obj.
constexpr int GA_TABLE_ROWS = 8;
struct GaArgs {
  const float *X, *S, *T, *sw, *pvec;   // single-object form (table == nullptr)
  const __half *XS;                     // split-fp16 tile image of X (frtm_split_samples) or null
  const long long *table;               // batched form: blockIdx.y = object
  float *partial;                       // [n_obj][cap][c*9]
  float *dbg;                           // optional debug dump of one CTA's score / v maps (tests only), else null
  int n_obj, cap, c, h, w, use_y;
};

// Split tile image of one sample (c, hw) fp32 -> [ntiles][hi|lo][c rows][64 px] fp16 of 16*x, rows in the 128-byte
// swizzled shared-memory layout of a K-major (pixels contiguous) UMMA operand.  ntiles is even.
constexpr int GC_TILE = 64;
__host__ __device__ inline int gc_ntiles(int hw) { return ((hw + 2 * GC_TILE - 1) / (2 * GC_TILE)) * 2; }
__host__ __device__ inline int64_t gc_sample_halves(int c, int hw) { return (int64_t)gc_ntiles(hw) * 2 * c * GC_TILE; }

// true if the tensor-core operator kernel supports this problem shape
bool gn_apply_tc_supported(int c, int h, int w);
// launches the tensor-core operator kernel; grid = (cap, n_obj)
int gn_apply_tc_launch(const GaArgs &a, cudaStream_t st);

}  // namespace frtm

namespace frtm {
constexpr float GC_ACT_SCALE = 16.f;

// One work item of the split: 8 consecutive pixels of channel row r in tile j -> one 16-byte chunk in each plane.
// item index = (j * c + r) * 8 + ch8;  src = the sample (c, hw) fp32;  dst = the sample's tile image.
__device__ __forceinline__ void gc_split_item(const float *__restrict__ src, __half *__restrict__ dst, int c, int hw, int64_t item) {
  const int ch8 = (int)(item & 7);
  const int64_t jr = item >> 3;
  const int r = (int)(jr % c);
  const int j = (int)(jr / c);
  const int q0 = j * GC_TILE + ch8 * 8;
  __align__(16) __half hh[8];
  __align__(16) __half ll[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float v = (q0 + e < hw) ? src[(int64_t)r * hw + q0 + e] * GC_ACT_SCALE : 0.f;
    hh[e] = __float2half_rn(v);
    ll[e] = __float2half_rn(v - __half2float(hh[e]));
  }
  const int64_t off = ((int64_t)(j * 2) * c + r) * GC_TILE + ((ch8 ^ (r & 7)) << 3);
  *reinterpret_cast<uint4 *>(dst + off) = *reinterpret_cast<const uint4 *>(hh);
  *reinterpret_cast<uint4 *>(dst + off + (int64_t)c * GC_TILE) = *reinterpret_cast<const uint4 *>(ll);
}
}  // namespace frtm
