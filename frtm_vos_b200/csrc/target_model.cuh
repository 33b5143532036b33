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

// Operator image of one memory sample, streamed by the tensor-core operator kernel (gn_apply_tc.cu):
//   [ntiles][hi|lo][c rows][64 px] fp16   split tile image of the features, 16*x = hi + lo, rows in the 128-byte
//                                         swizzled shared-memory layout of a UMMA operand, pixels beyond hw zero
//   [nchunks][10][256 px] fp32            stencil taps 0..8 and U^T w^2 y of 256 consecutive pixels per chunk
// ntiles is even.  Same bytes per element as the fp32 arrays it mirrors (plus padding of the last tile / chunk).
constexpr int GC_TILE = 64;
constexpr int GC_CHUNK_PX = 256;
constexpr int GC_CHUNK_BYTES = 10 * GC_CHUNK_PX * 4;
constexpr float GC_ACT_SCALE = 16.f;
__host__ __device__ inline int gc_ntiles(int hw) { return ((hw + 2 * GC_TILE - 1) / (2 * GC_TILE)) * 2; }
__host__ __device__ inline int gc_nchunks(int hw) { return (hw + GC_CHUNK_PX - 1) / GC_CHUNK_PX; }
__host__ __device__ inline int64_t gc_sample_bytes(int c, int hw) {
  return (int64_t)gc_ntiles(hw) * 2 * c * GC_TILE * 2 + (int64_t)gc_nchunks(hw) * GC_CHUNK_BYTES;
}
__host__ __device__ inline int64_t gc_sample_items(int c, int hw) {
  return (int64_t)gc_ntiles(hw) * c * 8 + (int64_t)gc_nchunks(hw) * 10 * (GC_CHUNK_PX / 4);
}

// true if the tensor-core operator kernel supports this problem shape
bool gn_apply_tc_supported(int c, int h, int w);
// launches the tensor-core operator kernel over all (object, sample) pairs
int gn_apply_tc_launch(const GaArgs &a, cudaStream_t st);

// One work item of the image build.  Items [0, ntiles*c*8): 8 consecutive pixels of channel row r in tile j -> one
// 16-byte chunk in each plane.  Remaining items: 4 consecutive pixels of one stencil / uty row of one chunk.
__device__ __forceinline__ void gc_image_item(const float *__restrict__ x, const float *__restrict__ stencil,
                                              const float *__restrict__ uty, uint8_t *__restrict__ img, int c, int hw,
                                              int64_t item) {
  const int ntiles = gc_ntiles(hw);
  const int64_t nx = (int64_t)ntiles * c * 8;
  if (item < nx) {
    const int ch8 = (int)(item & 7);
    const int64_t jr = item >> 3;
    const int r = (int)(jr % c);
    const int j = (int)(jr / c);
    const int q0 = j * GC_TILE + ch8 * 8;
    __align__(16) __half hh[8];
    __align__(16) __half ll[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float v = (q0 + e < hw) ? x[(int64_t)r * hw + q0 + e] * GC_ACT_SCALE : 0.f;
      hh[e] = __float2half_rn(v);
      ll[e] = __float2half_rn(v - __half2float(hh[e]));
    }
    __half *dst = reinterpret_cast<__half *>(img);
    const int64_t off = ((int64_t)(j * 2) * c + r) * GC_TILE + ((ch8 ^ (r & 7)) << 3);
    *reinterpret_cast<uint4 *>(dst + off) = *reinterpret_cast<const uint4 *>(hh);
    *reinterpret_cast<uint4 *>(dst + off + (int64_t)c * GC_TILE) = *reinterpret_cast<const uint4 *>(ll);
  } else {
    const int64_t e = item - nx;                       // (chunk m, row t, group of 4 pixels)
    const int g4 = (int)(e % (GC_CHUNK_PX / 4));
    const int t = (int)((e / (GC_CHUNK_PX / 4)) % 10);
    const int m = (int)(e / (10 * (GC_CHUNK_PX / 4)));
    const int q0 = m * GC_CHUNK_PX + g4 * 4;
    const float *src = t < 9 ? stencil + (int64_t)t * hw : uty;
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = (q0 + u < hw) ? src[q0 + u] : 0.f;
    float *dst = reinterpret_cast<float *>(img + (int64_t)ntiles * 2 * c * GC_TILE * 2) + ((int64_t)m * 10 + t) * GC_CHUNK_PX + g4 * 4;
    *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

}  // namespace frtm
