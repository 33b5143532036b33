// Warp-level tensor-core helpers (ldmatrix / mma.sync m16n8k16 f16 -> f32) and the split-fp16 utilities shared by the
// single-pass GN/CG operator kernels (gn_apply_mma.cu: one CTA per sample; gn_apply_cl.cu: one cluster per sample).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace frtm {

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// d += a (16x16, row) x b (16x8, col), fp16 operands, fp32 accumulate
__device__ __forceinline__ void hmma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - f.x, x1 - f.y);
  hi = *reinterpret_cast<const uint32_t *>(&h);
  lo = *reinterpret_cast<const uint32_t *>(&l);
}

// power of two that brings a non-negative float (given by its bits) into [2^9, 2^10); 1 for 0 / inf / nan
__device__ __forceinline__ float scale_from_bits(unsigned bits) {
  const int e = (int)(bits >> 23);                    // biased exponent: value = 1.m x 2^(e-127)
  if (bits == 0u || e == 0 || e >= 255) return 1.f;
  const int se = min(max(127 + 9 - (e - 127), 27), 227);   // 2^(9 - (e-127)), clamped like pow2_scale
  return __uint_as_float((unsigned)se << 23);
}

}  // namespace frtm
