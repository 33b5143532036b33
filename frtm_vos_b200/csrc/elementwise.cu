// Bandwidth-bound glue kernels of the backbone / refinement network / mask merge (NHWC fp32).
#include "common.cuh"
#include <cuda_fp16.h>

namespace frtm {

// ---------------------------------------------------------------- normalise ------------------------------------
__global__ void normalize_kernel(const uint8_t *__restrict__ img, int64_t HW, int64_t total, float4 *__restrict__ out,
                                 float s0, float s1, float s2, float b0, float b1, float b2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t b = i / HW, p = i - b * HW;
  const uint8_t *src = img + b * 3 * HW + p;
  float4 v;
  // mul then add, separately rounded, exactly like  norm_weight * input.float() + norm_bias
  v.x = __fadd_rn(__fmul_rn(s0, (float)src[0]), b0);
  v.y = __fadd_rn(__fmul_rn(s1, (float)src[HW]), b1);
  v.z = __fadd_rn(__fmul_rn(s2, (float)src[2 * HW]), b2);
  v.w = 0.f;
  out[i] = v;
}

// ---------------------------------------------------------------- stem patches --------------------------------
// im2col of the 7x7 / stride-2 / pad-3 stem on the normalised image, written directly as the split fp16 planes of a
// (B,Ho,Wo,192) tensor with k = ky*24 + c*8 + kx (kx = 7 and k >= 168 are zero): every (ky, c) row of a patch is one
// aligned 16-byte chunk of 7 consecutive input bytes.  The stem then runs as a 1x1 tensor-core conv with K = 192
// (weights permuted the same way) instead of 7.7 GMAC of fp32 CUDA-core FMAs per 8 frames.
// One thread = one chunk (one 16-byte store per plane).
__global__ void __launch_bounds__(256) stem_patches_kernel(const uint8_t *__restrict__ img, int B, int H, int W, int Ho, int Wo,
                                                           __half *__restrict__ hi, __half *__restrict__ lo, float s0, float s1,
                                                           float s2, float b0, float b1, float b2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * Ho * Wo * 24;
  if (i >= total) return;
  const int j = (int)(i % 24);
  const int64_t pix = i / 24;
  const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho);
  const int b = (int)(pix / ((int64_t)Wo * Ho));
  __align__(16) __half hh[8];
  __align__(16) __half ll[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { hh[e] = __float2half_rn(0.f); ll[e] = hh[e]; }
  const int ky = j / 3, c = j - ky * 3;
  const int iy = 2 * oy + ky - 3;
  if (j < 21 && iy >= 0 && iy < H) {
    const float sc = c == 0 ? s0 : (c == 1 ? s1 : s2), bc = c == 0 ? b0 : (c == 1 ? b1 : b2);
    const uint8_t *row = img + (((int64_t)b * 3 + c) * H + iy) * (int64_t)W;
    const int ix0 = 2 * ox - 3;
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) {
      const int ix = ix0 + kx;
      if (ix >= 0 && ix < W) {
        // mul then add, separately rounded, exactly like  norm_weight * input.float() + norm_bias
        const float v = __fadd_rn(__fmul_rn(sc, (float)row[ix]), bc) * 16.f;
        hh[kx] = __float2half_rn(v);
        ll[kx] = __float2half_rn(v - __half2float(hh[kx]));
      }
    }
  }
  *reinterpret_cast<uint4 *>(hi + pix * 192 + j * 8) = *reinterpret_cast<const uint4 *>(hh);
  *reinterpret_cast<uint4 *>(lo + pix * 192 + j * 8) = *reinterpret_cast<const uint4 *>(ll);
}

// ---------------------------------------------------------------- max pool -------------------------------------
__global__ void maxpool_kernel(const float *__restrict__ x, int B, int H, int W, int C, int Ho, int Wo,
                               float *__restrict__ y, float *__restrict__ y_nchw, __half *__restrict__ y_hi = nullptr,
                               __half *__restrict__ y_lo = nullptr) {
  const int C4 = C / 4;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * Ho * Wo * C4;
  if (i >= total) return;
  const int c4 = (int)(i % C4);
  int64_t r = i / C4;
  const int ox = (int)(r % Wo);
  r /= Wo;
  const int oy = (int)(r % Ho);
  const int b = (int)(r / Ho);
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  for (int dy = 0; dy < 3; ++dy) {
    const int iy = oy * 2 - 1 + dy;
    if (iy < 0 || iy >= H) continue;
    for (int dx = 0; dx < 3; ++dx) {
      const int ix = ox * 2 - 1 + dx;
      if (ix < 0 || ix >= W) continue;
      const float4 v = *reinterpret_cast<const float4 *>(x + (((int64_t)b * H + iy) * W + ix) * C + c4 * 4);
      m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
  }
  if (y) *reinterpret_cast<float4 *>(y + (((int64_t)b * Ho + oy) * Wo + ox) * C + c4 * 4) = m;
  if (y_hi) {   // split planes of 16*x for the tensor-core conv that follows (split_kernel's conversion)
    const float v[4] = {m.x * 16.f, m.y * 16.f, m.z * 16.f, m.w * 16.f};
    __half h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      h[k] = __float2half_rn(v[k]);
      l[k] = __float2half_rn(v[k] - __half2float(h[k]));
    }
    *reinterpret_cast<uint2 *>(y_hi + i * 4) = *reinterpret_cast<const uint2 *>(h);
    *reinterpret_cast<uint2 *>(y_lo + i * 4) = *reinterpret_cast<const uint2 *>(l);
  }
  if (y_nchw) {
    const int64_t hw = (int64_t)Ho * Wo, pix = (int64_t)oy * Wo + ox;
    float *d = y_nchw + ((int64_t)b * C + c4 * 4) * hw + pix;
    d[0] = m.x; d[hw] = m.y; d[2 * hw] = m.z; d[3 * hw] = m.w;
  }
}

// ---------------------------------------------------------------- bilinear -------------------------------------
template <int VEC>
__global__ void resize_bilinear_kernel(const float *__restrict__ x, int B, int H, int W, int C, int ldx,
                                       float *__restrict__ y, int Ho, int Wo, int ldy, int coff, int accumulate,
                                       float sh, float sw) {
  const int CV = C / VEC;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * Ho * Wo * CV;
  if (i >= total) return;
  const int cv = (int)(i % CV);
  int64_t r = i / CV;
  const int ox = (int)(r % Wo);
  r /= Wo;
  const int oy = (int)(r % Ho);
  const int b = (int)(r / Ho);
  int y0, y1, x0, x1;
  float ly, lx;
  bilinear_src(oy, sh, H, y0, y1, ly);
  bilinear_src(ox, sw, W, x0, x1, lx);
  const float hy = 1.f - ly, hx = 1.f - lx;
  const float *p00 = x + (((int64_t)b * H + y0) * W + x0) * ldx + cv * VEC;
  const float *p01 = x + (((int64_t)b * H + y0) * W + x1) * ldx + cv * VEC;
  const float *p10 = x + (((int64_t)b * H + y1) * W + x0) * ldx + cv * VEC;
  const float *p11 = x + (((int64_t)b * H + y1) * W + x1) * ldx + cv * VEC;
  float *d = y + (((int64_t)b * Ho + oy) * Wo + ox) * ldy + coff + cv * VEC;
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    float v = hy * (hx * p00[k] + lx * p01[k]) + ly * (hx * p10[k] + lx * p11[k]);
    d[k] = accumulate ? d[k] + v : v;
  }
}

// ---------------------------------------------------------------- bicubic x2 -----------------------------------
__constant__ float kCubicE[4] = {-0.10546875f, 0.87890625f, 0.26171875f, -0.03515625f};

// One thread = two horizontally adjacent 4x4 input windows (a 4x5 patch, 4 channels) = a 2x4 block of outputs.  The
// polyphase filters are outer products of the same two 4-tap rows, so the patch is filtered along x once (four phase
// values per input row) and then along y: 20 loads and 384 FMAs per 8 outputs instead of 32 loads and 640 — the kernel
// is instruction-bound, not bandwidth-bound.
__global__ void __launch_bounds__(256) pyrup_bicubic_kernel(const float *__restrict__ x, int B, int H, int W, int C,
                                                            float *__restrict__ y, __half *__restrict__ y_hi,
                                                            __half *__restrict__ y_lo) {
  const int C4 = C / 4, Ho = 2 * H, Wo = 2 * W;
  const int NP = (W + 2) / 2;                                   // window pairs per row (windows n = 0..W)
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * (H + 1) * NP * C4;
  if (i >= total) return;
  const int c4 = (int)(i % C4);
  int64_t r = i / C4;
  const int n = 2 * (int)(r % NP);
  r /= NP;
  const int m = (int)(r % (H + 1));
  const int b = (int)(r / (H + 1));
  const float E0 = kCubicE[0], E1 = kCubicE[1], E2 = kCubicE[2], E3 = kCubicE[3];
  // window rows m-2..m+1, cols n-2..n+2 (replicate padding); x pass -> phases: window n even/odd, window n+1 even/odd
  float4 hx[4][4];
#pragma unroll
  for (int ky = 0; ky < 4; ++ky) {
    const int sy = min(max(m + ky - 2, 0), H - 1);
    float4 v[5];
#pragma unroll
    for (int kx = 0; kx < 5; ++kx) {
      const int sx = min(max(n + kx - 2, 0), W - 1);
      v[kx] = *reinterpret_cast<const float4 *>(x + (((int64_t)b * H + sy) * W + sx) * C + c4 * 4);
    }
#define FRTM_DOT4(a0, a1, a2, a3, w0, w1, w2, w3)                                                         \
  make_float4((w0 * a0.x + w1 * a1.x) + (w2 * a2.x + w3 * a3.x), (w0 * a0.y + w1 * a1.y) + (w2 * a2.y + w3 * a3.y), \
              (w0 * a0.z + w1 * a1.z) + (w2 * a2.z + w3 * a3.z), (w0 * a0.w + w1 * a1.w) + (w2 * a2.w + w3 * a3.w))
    hx[ky][0] = FRTM_DOT4(v[0], v[1], v[2], v[3], E0, E1, E2, E3);
    hx[ky][1] = FRTM_DOT4(v[0], v[1], v[2], v[3], E3, E2, E1, E0);
    hx[ky][2] = FRTM_DOT4(v[1], v[2], v[3], v[4], E0, E1, E2, E3);
    hx[ky][3] = FRTM_DOT4(v[1], v[2], v[3], v[4], E3, E2, E1, E0);
  }
#pragma unroll
  for (int ry = 0; ry < 2; ++ry) {
    const int oy = 2 * m + ry - 1;                              // pre-crop row 2m + ry -> cropped 2m + ry - 1
    if (oy < 0 || oy >= Ho) continue;
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) {
      const int ox = 2 * n + ph - 1;
      if (ox < 0 || ox >= Wo) continue;
      const float4 acc = ry ? FRTM_DOT4(hx[0][ph], hx[1][ph], hx[2][ph], hx[3][ph], E3, E2, E1, E0)
                            : FRTM_DOT4(hx[0][ph], hx[1][ph], hx[2][ph], hx[3][ph], E0, E1, E2, E3);
      const int64_t o = (((int64_t)b * Ho + oy) * Wo + ox) * C + c4 * 4;
      if (y) *reinterpret_cast<float4 *>(y + o) = acc;
      if (y_hi) {   // split planes of 16*x for the tensor-core conv that follows (conv_tc.cu)
        const __half2 h01 = __floats2half2_rn(acc.x * 16.f, acc.y * 16.f), h23 = __floats2half2_rn(acc.z * 16.f, acc.w * 16.f);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(acc.x * 16.f - f01.x, acc.y * 16.f - f01.y);
        const __half2 l23 = __floats2half2_rn(acc.z * 16.f - f23.x, acc.w * 16.f - f23.y);
        uint2 ph_, pl_;
        ph_.x = *reinterpret_cast<const uint32_t *>(&h01); ph_.y = *reinterpret_cast<const uint32_t *>(&h23);
        pl_.x = *reinterpret_cast<const uint32_t *>(&l01); pl_.y = *reinterpret_cast<const uint32_t *>(&l23);
        *reinterpret_cast<uint2 *>(y_hi + o) = ph_;
        *reinterpret_cast<uint2 *>(y_lo + o) = pl_;
      }
    }
  }
#undef FRTM_DOT4
}

// ---------------------------------------------------------------- global average pool ---------------------------
constexpr int GAP_CHUNK = 64;   // pixels per stage-1 block

// stage 1: block = 4 pixel groups x 64 channel lanes; a thread sums the 16 pixels of its group with ALL loads in flight
// (a strided loop with 4 loads in flight per thread made the launch a chain of 16 memory round trips: 17 us even for a
// 15 x 27 map), then a smem reduce.  Fixed order -> deterministic.
__global__ void __launch_bounds__(256) gap_stage1_kernel(const float *__restrict__ x, int HW, int C, int ldx, int nchunks,
                                                         float *__restrict__ part) {
  __shared__ float red[4][64];
  const int b = blockIdx.y, ch = blockIdx.x;
  const int p0 = ch * GAP_CHUNK, p1 = min(p0 + GAP_CHUNK, HW);
  const int lane = threadIdx.x & 63, grp = threadIdx.x >> 6;
  for (int c0 = 0; c0 < C; c0 += 64) {
    const int c = c0 + lane;
    float v[GAP_CHUNK / 4];
    const float *base = x + (int64_t)b * HW * ldx + c;
#pragma unroll
    for (int k = 0; k < GAP_CHUNK / 4; ++k) {
      const int p = p0 + grp + 4 * k;
      v[k] = (c < C && p < p1) ? base[(int64_t)p * ldx] : 0.f;
    }
#pragma unroll
    for (int w = GAP_CHUNK / 8; w > 0; w >>= 1)
#pragma unroll
      for (int k = 0; k < w; ++k) v[k] += v[k + w];
    red[grp][lane] = v[0];
    __syncthreads();
    if (grp == 0 && c < C) part[((int64_t)b * nchunks + ch) * C + c] = (red[0][lane] + red[1][lane]) + (red[2][lane] + red[3][lane]);
    __syncthreads();
  }
}
// stage 2: block = one image; 4 chunk groups x 64 channel lanes, 4 partial sums in flight per thread, smem reduce
__global__ void __launch_bounds__(256) gap_stage2_kernel(const float *__restrict__ part, int HW, int C, int nchunks,
                                                         float *__restrict__ out) {
  __shared__ float red[4][64];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 63, grp = threadIdx.x >> 6;
  for (int c0 = 0; c0 < C; c0 += 64) {
    const int c = c0 + lane;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (c < C) {
      const float *base = part + (int64_t)b * nchunks * C + c;
      int k = grp;
      for (; k + 12 < nchunks; k += 16) {
        s0 += base[(int64_t)k * C];
        s1 += base[(int64_t)(k + 4) * C];
        s2 += base[(int64_t)(k + 8) * C];
        s3 += base[(int64_t)(k + 12) * C];
      }
      for (; k < nchunks; k += 4) s0 += base[(int64_t)k * C];
    }
    red[grp][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (grp == 0 && c < C) out[(int64_t)b * C + c] = ((red[0][lane] + red[1][lane]) + (red[2][lane] + red[3][lane])) / (float)HW;
    __syncthreads();
  }
}

// ---------------------------------------------------------------- CAB -------------------------------------------
// gate = sigmoid(W2 relu(W1 [sp | dp] + b1) + b2) per image (seg_network.py:34-37).  One block per image, blockDim = 4 * C:
// four threads share an output row (a quarter of the dot product each, float4 loads, two accumulators), then a shared
// memory reduce — the single 128-long FMA chain per thread it replaces took 14 us per launch.
__global__ void cab_gate_kernel(const float *__restrict__ sp, const float *__restrict__ dp, const float *__restrict__ w1,
                                const float *__restrict__ b1, const float *__restrict__ w2, const float *__restrict__ b2,
                                int C, float *__restrict__ gate) {
  extern __shared__ float sm[];  // pooled[2C] + hidden[C] + partial[4C]
  float *pooled = sm, *hid = sm + 2 * C, *part = sm + 3 * C;
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    pooled[i] = sp[(int64_t)b * C + i];
    pooled[C + i] = dp[(int64_t)b * C + i];
  }
  __syncthreads();
  const int o = threadIdx.x >> 2, q = threadIdx.x & 3;
  {
    const int K = 2 * C, kq = K / 4;
    const float *wr = w1 + (int64_t)o * K + q * kq, *pv = pooled + q * kq;
    float s0 = 0.f, s1 = 0.f;
    for (int k = 0; k < kq; k += 8) {
      const float4 a0 = *reinterpret_cast<const float4 *>(wr + k), a1 = *reinterpret_cast<const float4 *>(wr + k + 4);
      s0 = fmaf(a0.x, pv[k], s0); s0 = fmaf(a0.y, pv[k + 1], s0); s0 = fmaf(a0.z, pv[k + 2], s0); s0 = fmaf(a0.w, pv[k + 3], s0);
      s1 = fmaf(a1.x, pv[k + 4], s1); s1 = fmaf(a1.y, pv[k + 5], s1); s1 = fmaf(a1.z, pv[k + 6], s1); s1 = fmaf(a1.w, pv[k + 7], s1);
    }
    part[threadIdx.x] = s0 + s1;
  }
  __syncthreads();
  if (q == 0) hid[o] = fmaxf(((part[4 * o] + part[4 * o + 1]) + (part[4 * o + 2] + part[4 * o + 3])) + b1[o], 0.f);
  __syncthreads();
  {
    const int kq = C / 4;
    const float *wr = w2 + (int64_t)o * C + q * kq, *hv = hid + q * kq;
    float s0 = 0.f, s1 = 0.f;
    for (int k = 0; k < kq; k += 8) {
      const float4 a0 = *reinterpret_cast<const float4 *>(wr + k), a1 = *reinterpret_cast<const float4 *>(wr + k + 4);
      s0 = fmaf(a0.x, hv[k], s0); s0 = fmaf(a0.y, hv[k + 1], s0); s0 = fmaf(a0.z, hv[k + 2], s0); s0 = fmaf(a0.w, hv[k + 3], s0);
      s1 = fmaf(a1.x, hv[k + 4], s1); s1 = fmaf(a1.y, hv[k + 5], s1); s1 = fmaf(a1.z, hv[k + 6], s1); s1 = fmaf(a1.w, hv[k + 7], s1);
    }
    part[threadIdx.x] = s0 + s1;
  }
  __syncthreads();
  if (q == 0) {
    const float s = ((part[4 * o] + part[4 * o + 1]) + (part[4 * o + 2] + part[4 * o + 3])) + b2[o];
    gate[(int64_t)b * C + o] = 1.f / (1.f + expf(-s));
  }
}

// CAB gate straight from the two maps (C = 64): stage 1 of both global average pools in ONE launch, then one block per
// image that finishes both pools (the arithmetic of gap_stage2_kernel, term for term) and runs the gate on them — two
// launches instead of five small, latency-bound ones per refinement level.  Bit-identical to the separate kernels.
__global__ void __launch_bounds__(256) gap_stage1_pair_kernel(const float *__restrict__ xa, int HWa, int lda, int nca,
                                                              float *__restrict__ parta, const float *__restrict__ xb, int HWb,
                                                              int ldb, int ncb, float *__restrict__ partb, int C) {
  __shared__ float red[4][64];
  const bool second = (int)blockIdx.x >= nca;
  const float *x = second ? xb : xa;
  const int HW = second ? HWb : HWa, ldx = second ? ldb : lda, nchunks = second ? ncb : nca;
  float *part = second ? partb : parta;
  const int b = blockIdx.y, ch = second ? blockIdx.x - nca : blockIdx.x;
  const int p0 = ch * GAP_CHUNK, p1 = min(p0 + GAP_CHUNK, HW);
  const int lane = threadIdx.x & 63, grp = threadIdx.x >> 6;
  for (int c0 = 0; c0 < C; c0 += 64) {
    const int c = c0 + lane;
    float v[GAP_CHUNK / 4];
    const float *base = x + (int64_t)b * HW * ldx + c;
#pragma unroll
    for (int k = 0; k < GAP_CHUNK / 4; ++k) {
      const int p = p0 + grp + 4 * k;
      v[k] = (c < C && p < p1) ? base[(int64_t)p * ldx] : 0.f;
    }
#pragma unroll
    for (int w = GAP_CHUNK / 8; w > 0; w >>= 1)
#pragma unroll
      for (int k = 0; k < w; ++k) v[k] += v[k + w];
    red[grp][lane] = v[0];
    __syncthreads();
    if (grp == 0 && c < C) part[((int64_t)b * nchunks + ch) * C + c] = (red[0][lane] + red[1][lane]) + (red[2][lane] + red[3][lane]);
    __syncthreads();
  }
}

// pooled[c] of image b from the stage-1 partial sums: gap_stage2_kernel for C = 64 (lane = channel, 4 chunk groups)
__device__ __forceinline__ void gap_finish64(const float *__restrict__ part, int b, int HW, int nchunks, float (*red)[64],
                                             float *__restrict__ pooled) {
  const int lane = threadIdx.x & 63, grp = threadIdx.x >> 6;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  const float *base = part + (int64_t)b * nchunks * 64 + lane;
  int k = grp;
  for (; k + 12 < nchunks; k += 16) {
    s0 += base[(int64_t)k * 64];
    s1 += base[(int64_t)(k + 4) * 64];
    s2 += base[(int64_t)(k + 8) * 64];
    s3 += base[(int64_t)(k + 12) * 64];
  }
  for (; k < nchunks; k += 4) s0 += base[(int64_t)k * 64];
  red[grp][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (grp == 0) pooled[lane] = ((red[0][lane] + red[1][lane]) + (red[2][lane] + red[3][lane])) / (float)HW;
  __syncthreads();
}

__global__ void __launch_bounds__(256) cab_pool_gate64_kernel(const float *__restrict__ parta, int HWa, int nca,
                                                              const float *__restrict__ partb, int HWb, int ncb,
                                                              const float *__restrict__ deep_pool, const float *__restrict__ w1,
                                                              const float *__restrict__ b1, const float *__restrict__ w2,
                                                              const float *__restrict__ b2, float *__restrict__ gate,
                                                              float *__restrict__ pool_out) {
  constexpr int C = 64;
  __shared__ float red[4][64];
  __shared__ __align__(16) float pooled[2 * C];
  __shared__ __align__(16) float hid[C];
  __shared__ float part[4 * C];
  const int b = blockIdx.x;
  gap_finish64(parta, b, HWa, nca, red, pooled);
  if (partb) gap_finish64(partb, b, HWb, ncb, red, pooled + C);
  else if (threadIdx.x < C) pooled[C + threadIdx.x] = deep_pool[(int64_t)b * C + threadIdx.x];
  __syncthreads();
  if (pool_out && threadIdx.x < 2 * C) pool_out[(int64_t)b * 2 * C + threadIdx.x] = pooled[threadIdx.x];
  // the gate: cab_gate_kernel at C = 64 (blockDim = 4 C)
  const int o = threadIdx.x >> 2, q = threadIdx.x & 3;
  {
    const int K = 2 * C, kq = K / 4;
    const float *wr = w1 + (int64_t)o * K + q * kq, *pv = pooled + q * kq;
    float s0 = 0.f, s1 = 0.f;
    for (int k = 0; k < kq; k += 8) {
      const float4 a0 = *reinterpret_cast<const float4 *>(wr + k), a1 = *reinterpret_cast<const float4 *>(wr + k + 4);
      s0 = fmaf(a0.x, pv[k], s0); s0 = fmaf(a0.y, pv[k + 1], s0); s0 = fmaf(a0.z, pv[k + 2], s0); s0 = fmaf(a0.w, pv[k + 3], s0);
      s1 = fmaf(a1.x, pv[k + 4], s1); s1 = fmaf(a1.y, pv[k + 5], s1); s1 = fmaf(a1.z, pv[k + 6], s1); s1 = fmaf(a1.w, pv[k + 7], s1);
    }
    part[threadIdx.x] = s0 + s1;
  }
  __syncthreads();
  if (q == 0) hid[o] = fmaxf(((part[4 * o] + part[4 * o + 1]) + (part[4 * o + 2] + part[4 * o + 3])) + b1[o], 0.f);
  __syncthreads();
  {
    const int kq = C / 4;
    const float *wr = w2 + (int64_t)o * C + q * kq, *hv = hid + q * kq;
    float s0 = 0.f, s1 = 0.f;
    for (int k = 0; k < kq; k += 8) {
      const float4 a0 = *reinterpret_cast<const float4 *>(wr + k), a1 = *reinterpret_cast<const float4 *>(wr + k + 4);
      s0 = fmaf(a0.x, hv[k], s0); s0 = fmaf(a0.y, hv[k + 1], s0); s0 = fmaf(a0.z, hv[k + 2], s0); s0 = fmaf(a0.w, hv[k + 3], s0);
      s1 = fmaf(a1.x, hv[k + 4], s1); s1 = fmaf(a1.y, hv[k + 5], s1); s1 = fmaf(a1.z, hv[k + 6], s1); s1 = fmaf(a1.w, hv[k + 7], s1);
    }
    part[threadIdx.x] = s0 + s1;
  }
  __syncthreads();
  if (q == 0) {
    const float s = ((part[4 * o] + part[4 * o + 1]) + (part[4 * o + 2] + part[4 * o + 3])) + b2[o];
    gate[(int64_t)b * C + o] = 1.f / (1.f + expf(-s));
  }
}

__global__ void cab_apply_kernel(const float4 *__restrict__ sh, const float *__restrict__ gate,
                                 const float *__restrict__ deeper, int vec, int HW, int C, int64_t total4,
                                 float4 *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int C4 = C / 4;
  const int c = (int)(i % C4) * 4;
  const int64_t b = i / ((int64_t)HW * C4);
  const float4 g = *reinterpret_cast<const float4 *>(gate + b * C + c);
  const float4 d = vec ? *reinterpret_cast<const float4 *>(deeper + b * C + c)
                       : reinterpret_cast<const float4 *>(deeper)[i];
  const float4 s = sh[i];
  // shallower * sigmoid(...) then + deeper: two roundings like the reference (seg_network.py:38-39)
  float4 o;
  o.x = __fadd_rn(__fmul_rn(s.x, g.x), d.x);
  o.y = __fadd_rn(__fmul_rn(s.y, g.y), d.y);
  o.z = __fadd_rn(__fmul_rn(s.z, g.z), d.z);
  o.w = __fadd_rn(__fmul_rn(s.w, g.w), d.w);
  out[i] = o;
}

// ---------------------------------------------------------------- layout helpers --------------------------------
__global__ void scatter_channel_kernel(const float *__restrict__ src, int64_t total, float *__restrict__ dst, int ld,
                                       int coff, int nzero) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float *d = dst + i * ld + coff;
  d[0] = src[i];
  for (int k = 1; k <= nzero; ++k) d[k] = 0.f;
}

__global__ void broadcast_objects_kernel(const float *__restrict__ src, int N, int HW, int C, int lds,
                                         float *__restrict__ dst, int ldd, int64_t total4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int C4 = C / 4;
  const int c = (int)(i % C4) * 4;
  int64_t r = i / C4;
  const int p = (int)(r % HW);
  const int64_t bo = r / HW;     // f*N + n
  const int64_t f = bo / N;
  const float4 v = *reinterpret_cast<const float4 *>(src + (f * HW + p) * lds + c);
  *reinterpret_cast<float4 *>(dst + (bo * HW + p) * ldd + c) = v;
}

// NHWC -> NCHW through a 32x32 shared tile (coalesced on both sides)
__global__ void nhwc_to_nchw_kernel(const float *__restrict__ x, int HW, int C, int ldx, float *__restrict__ y) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int p = p0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (p < HW && c < C) ? x[((int64_t)b * HW + p) * ldx + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, p = p0 + threadIdx.x;
    if (p < HW && c < C) y[((int64_t)b * C + c) * HW + p] = tile[threadIdx.x][j];
  }
}

// ---------------------------------------------------------------- merge -----------------------------------------
// One thread per pixel.  Mirrors model/tracker.py:203-221 (sigmoid of live objects, suppression under new objects'
// start masks, clamp, background = min(1-p), softmax over p/(1-p), first-occurrence argmax gate) followed by the
// label rule of :143-150 applied to the merged masks.  `src` holds logits for objects whose bit is set in
// logit_mask and probabilities (start masks) for the others.  masks[] doubles as per-pixel scratch.
__device__ __forceinline__ int softmax_argmax(const float *__restrict__ m, int N, int HW, int p, float z0, float &den_out,
                                              float &mx_out) {
  float mx = z0;
  for (int n = 0; n < N; ++n) {
    const float c = m[(int64_t)(n + 1) * HW + p];
    mx = fmaxf(mx, c / (1.f - c));
  }
  float den = expf(z0 - mx);
  for (int n = 0; n < N; ++n) {
    const float c = m[(int64_t)(n + 1) * HW + p];
    den += expf(c / (1.f - c) - mx);
  }
  float best = expf(z0 - mx) / den;
  int arg = 0;
  for (int n = 0; n < N; ++n) {
    const float c = m[(int64_t)(n + 1) * HW + p];
    const float s = expf(c / (1.f - c) - mx) / den;
    if (s > best) { best = s; arg = n + 1; }
  }
  den_out = den;
  mx_out = mx;
  return arg;
}

// Register-resident variant for N <= NMAX objects: the same arithmetic in the same order as merge_masks_kernel below, but every
// per-object value stays in registers between the passes (one global read of src, one global write of masks).
template <int NMAX>
__global__ void __launch_bounds__(256) merge_masks_reg_kernel(const float *__restrict__ src, unsigned long long logit_mask,
                                                              const uint8_t *__restrict__ suppress, int N, int HW,
                                                              const uint8_t *__restrict__ lut, int single,
                                                              float *__restrict__ masks, uint8_t *__restrict__ labels,
                                                              int *__restrict__ counts, int counts_stride) {
  src += (int64_t)blockIdx.y * N * HW;
  masks += (int64_t)blockIdx.y * (N + 1) * HW;
  labels += (int64_t)blockIdx.y * HW;
  counts += (int64_t)blockIdx.y * counts_stride;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const float lo = 1e-7f, hi = 1.f - 1e-7f;
  float m[NMAX];
#pragma unroll
  for (int n = 0; n < NMAX; ++n) m[n] = 0.f;
  if (p < HW) {
    float bg = INFINITY;
    const float keep = suppress ? (float)(1 - (int)suppress[p]) : 1.f;
#pragma unroll
    for (int n = 0; n < NMAX; ++n) {
      if (n < N) {
        float pr = src[(int64_t)n * HW + p];
        if ((logit_mask >> n) & 1ull) pr = (1.f / (1.f + expf(-pr))) * keep;
        pr = fminf(fmaxf(pr, lo), hi);
        m[n] = pr;
        bg = fminf(bg, 1.f - pr);
      }
    }
    // first softmax / argmax (softmax_argmax on the register copy).  Every odds ratio z = m / (1 - m) and every
    // exponential exp(z - max) is evaluated ONCE and kept in registers (the same expressions on the same inputs as the
    // three separate passes of merge_masks_kernel, hence the same bits; a third of the divisions and exponentials)
    const float z0 = bg / (1.f - bg);
    float mx = z0;
    float z[NMAX], e[NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) {
      z[n] = 0.f; e[n] = 0.f;
      if (n < N) { z[n] = m[n] / (1.f - m[n]); mx = fmaxf(mx, z[n]); }
    }
    const float e0 = expf(z0 - mx);
    float den = e0;
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
      if (n < N) { e[n] = expf(z[n] - mx); den += e[n]; }
    float best = e0 / den;
    int arg = 0;
#pragma unroll
    for (int n = 0; n < NMAX; ++n) {
      if (n < N) {
        const float sn = e[n] / den;
        e[n] = sn;                                   // the softmax value itself from here on
        if (sn > best) { best = sn; arg = n + 1; }
      }
    }
    masks[p] = (arg == 0) ? e0 / den : 0.f;
    float bg2 = INFINITY;
#pragma unroll
    for (int n = 0; n < NMAX; ++n) {
      if (n < N) {
        const float sn = (arg == n + 1) ? e[n] : 0.f;
        m[n] = sn;
        masks[(int64_t)(n + 1) * HW + p] = sn;
        bg2 = fminf(bg2, 1.f - fminf(fmaxf(sn, lo), hi));
      }
    }
    int lab;
    if (single) {
      lab = m[0] > 0.5f ? 1 : 0;
    } else {
      const float zb = bg2 / (1.f - bg2);
      float mx2 = zb;
#pragma unroll
      for (int n = 0; n < NMAX; ++n) {
        if (n < N) {
          const float c = fminf(fmaxf(m[n], lo), hi);
          z[n] = c / (1.f - c);
          mx2 = fmaxf(mx2, z[n]);
        }
      }
      const float eb = expf(zb - mx2);
      float den2 = eb;
#pragma unroll
      for (int n = 0; n < NMAX; ++n)
        if (n < N) { e[n] = expf(z[n] - mx2); den2 += e[n]; }
      float best2 = eb / den2;
      lab = 0;
#pragma unroll
      for (int n = 0; n < NMAX; ++n) {
        if (n < N) {
          const float s2 = e[n] / den2;
          if (s2 > best2) { best2 = s2; lab = n + 1; }
        }
      }
    }
    labels[p] = lut[lab];
  }
  // per-object count of pixels > 0.5 (update gate), one atomic per warp per object (integer -> deterministic)
#pragma unroll
  for (int n = 0; n < NMAX; ++n) {
    if (n < N) {
      const bool on = (p < HW) && m[n] > 0.5f;
      const unsigned b = __ballot_sync(0xffffffffu, on);
      if ((threadIdx.x & 31) == 0 && b) atomicAdd(&counts[n], __popc(b));
    }
  }
}

__global__ void merge_masks_kernel(const float *__restrict__ src, unsigned long long logit_mask,
                                   const uint8_t *__restrict__ suppress, int N, int HW, const uint8_t *__restrict__ lut,
                                   int single, float *__restrict__ masks, uint8_t *__restrict__ labels,
                                   int *__restrict__ counts, int counts_stride) {
  // blockIdx.y = frame of a block of consecutive frames (independent merges, one launch)
  src += (int64_t)blockIdx.y * N * HW;
  masks += (int64_t)blockIdx.y * (N + 1) * HW;
  labels += (int64_t)blockIdx.y * HW;
  counts += (int64_t)blockIdx.y * counts_stride;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const float lo = 1e-7f, hi = 1.f - 1e-7f;
  if (p < HW) {
    float bg = INFINITY;
    const float keep = suppress ? (float)(1 - (int)suppress[p]) : 1.f;
    for (int n = 0; n < N; ++n) {
      float pr = src[(int64_t)n * HW + p];
      if ((logit_mask >> n) & 1ull) pr = (1.f / (1.f + expf(-pr))) * keep;
      pr = fminf(fmaxf(pr, lo), hi);
      masks[(int64_t)(n + 1) * HW + p] = pr;
      bg = fminf(bg, 1.f - pr);
    }
    float den, mx;
    const float z0 = bg / (1.f - bg);
    const int arg = softmax_argmax(masks, N, HW, p, z0, den, mx);
    masks[p] = (arg == 0) ? expf(z0 - mx) / den : 0.f;
    float bg2 = INFINITY;
    for (int n = 0; n < N; ++n) {
      const float c = masks[(int64_t)(n + 1) * HW + p];
      const float s = (arg == n + 1) ? expf(c / (1.f - c) - mx) / den : 0.f;
      masks[(int64_t)(n + 1) * HW + p] = s;
      bg2 = fminf(bg2, 1.f - fminf(fmaxf(s, lo), hi));
    }
    int lab;
    if (single) {
      lab = masks[(int64_t)HW + p] > 0.5f ? 1 : 0;
    } else {
      // second clamp / softmax / argmax on the merged masks.  Only the winner is non-zero, so evaluate the rule
      // on clamped values without disturbing the stored masks.
      float mx2 = bg2 / (1.f - bg2);
      for (int n = 0; n < N; ++n) {
        const float c = fminf(fmaxf(masks[(int64_t)(n + 1) * HW + p], lo), hi);
        mx2 = fmaxf(mx2, c / (1.f - c));
      }
      float den2 = expf(bg2 / (1.f - bg2) - mx2);
      for (int n = 0; n < N; ++n) {
        const float c = fminf(fmaxf(masks[(int64_t)(n + 1) * HW + p], lo), hi);
        den2 += expf(c / (1.f - c) - mx2);
      }
      float best = expf(bg2 / (1.f - bg2) - mx2) / den2;
      lab = 0;
      for (int n = 0; n < N; ++n) {
        const float c = fminf(fmaxf(masks[(int64_t)(n + 1) * HW + p], lo), hi);
        const float s = expf(c / (1.f - c) - mx2) / den2;
        if (s > best) { best = s; lab = n + 1; }
      }
    }
    labels[p] = lut[lab];
  }
  // per-object count of pixels > 0.5 (update gate), one atomic per warp per object (integer -> deterministic)
  for (int n = 0; n < N; ++n) {
    const bool on = (p < HW) && masks[(int64_t)(n + 1) * HW + p] > 0.5f;
    const unsigned b = __ballot_sync(0xffffffffu, on);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&counts[n], __popc(b));
  }
}

// t[b][p][tap] = sum_c w[tap][c] x[b][p][c]  (9 taps, output channel stride 12, channels 9..11 zero)
__global__ void __launch_bounds__(256) tapmaps_kernel(const float *__restrict__ x, int64_t npix, int C, const float *__restrict__ w,
                                                      float *__restrict__ y) {
  extern __shared__ float ws[];  // [9][C]
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.f;
  const float4 *px = reinterpret_cast<const float4 *>(x + p * C);
  for (int c4 = 0; c4 < C / 4; ++c4) {
    const float4 v = px[c4];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float *wt = ws + t * C + c4 * 4;
      acc[t] = fmaf(v.x, wt[0], acc[t]);
      acc[t] = fmaf(v.y, wt[1], acc[t]);
      acc[t] = fmaf(v.z, wt[2], acc[t]);
      acc[t] = fmaf(v.w, wt[3], acc[t]);
    }
  }
  float4 *o = reinterpret_cast<float4 *>(y + p * 12);
  o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  o[2] = make_float4(acc[8], 0.f, 0.f, 0.f);
}

// out[b][y][x] = bias + sum_tap v[b][y+dy][x+dx][tap]   (zero outside the image = the conv's zero padding)
__global__ void __launch_bounds__(256) shift_sum9_kernel(const float *__restrict__ v, int B, int H, int W,
                                                         const float *__restrict__ bias, float *__restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * H * W) return;
  const int xw = (int)(idx % W), yh = (int)((idx / W) % H);
  const int64_t b = idx / ((int64_t)W * H);
  float acc = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int iy = yh + t / 3 - 1, ix = xw + t % 3 - 1;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) acc += v[((b * H + iy) * W + ix) * 12 + t];
  }
  out[idx] = acc + (bias ? bias[0] : 0.f);
}


// ---------------------------------------------------------------- fused tail of the upsampler --------------------
// out[b][Y][X] = bias + sum_tap R_tap[Y+dy][X+dx]  (zero outside the image = the final conv's zero padding), where
// R_tap = bilinear_(H,W)( pyrup_bicubic( t_tap ) ) and t (B,h,w,12) are the 9 tap maps of the final 3x3 conv contracted at
// low resolution (seg_network.py:139-145 by linearity).  Along each axis bilinear o bicubic-x2 is ONE 5-tap window on the
// low-resolution samples (the two upsampled neighbours of a bilinear source position share or straddle one 4-tap bicubic
// window), with weights that depend on the destination index and the tap shift only.  One block = a 16x64 output tile:
// the tap maps of the tile (with halo) are staged once, a row pass applies the x windows of the three column shifts
// (every thread keeps the windows of its own column in registers), a column pass the y windows of the three row shifts
// (from a small table) — three block barriers per tile, nothing but t is read and nothing but the logits are written.
constexpr int FT_TY = 16, FT_TX = 64, FT_UY = 24, FT_UX = 76, FT_NTY = 16, FT_NTX = 42;
constexpr int FT_SMEM = (9 * FT_NTY * (FT_NTX + 1) + 9 * FT_NTY * FT_TX + 3 * FT_TY * 8) * 4;

// 5-tap window of bilinear(pyrup_bicubic(.)) at destination index d (already clamped to the image; ok = false -> zero
// weights): samples lo + base .. lo + base + 4 of the low-resolution axis.
__device__ __forceinline__ void fused_axis_window(int d, bool ok, float scale, int up_size, int lo, int &base, float (&a)[5]) {
  int u0, u1;
  float lam;
  bilinear_src(d, scale, up_size, u0, u1, lam);
  // upsampled (cropped) index u = pre-crop u + 1 = 2n (+1): window n-2 .. n+1, even -> kCubicE, odd -> reversed
  const int n0 = (u0 + 1) >> 1, n1 = (u1 + 1) >> 1;
  const bool odd0 = (u0 + 1) & 1, odd1 = (u1 + 1) & 1;
  const float w0 = ok ? 1.f - lam : 0.f, w1 = ok ? lam : 0.f;
  float c0[4], c1[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    c0[k] = w0 * (odd0 ? kCubicE[3 - k] : kCubicE[k]);
    c1[k] = w1 * (odd1 ? kCubicE[3 - k] : kCubicE[k]);
  }
  const bool same = n1 == n0;                               // else n1 = n0 + 1
  a[0] = c0[0] + (same ? c1[0] : 0.f);
  a[1] = c0[1] + (same ? c1[1] : c1[0]);
  a[2] = c0[2] + (same ? c1[2] : c1[1]);
  a[3] = c0[3] + (same ? c1[3] : c1[2]);
  a[4] = same ? 0.f : c1[3];
  base = n0 - 2 - lo;
}

__global__ void __launch_bounds__(256) upsample_tapsum_kernel(const float *__restrict__ t12, int B, int h, int w, int H,
                                                              int W, const float *__restrict__ bias,
                                                              float *__restrict__ out) {
  extern __shared__ __align__(16) float ft_smem[];
  float (*T)[FT_NTY][FT_NTX + 1] = reinterpret_cast<float (*)[FT_NTY][FT_NTX + 1]>(ft_smem);                 // [9]
  float (*R)[FT_NTY][FT_TX] = reinterpret_cast<float (*)[FT_NTY][FT_TX]>(ft_smem + 9 * FT_NTY * (FT_NTX + 1));   // [9]
  float *WY = ft_smem + 9 * FT_NTY * (FT_NTX + 1) + 9 * FT_NTY * FT_TX;   // [3][FT_TY][8]: a[0..4], base (int bits)
  const int Hu = 2 * h, Wu = 2 * w;
  const float sy = (float)Hu / (float)H, sx = (float)Wu / (float)W;
  const int b = blockIdx.z, Y0 = blockIdx.y * FT_TY, X0 = blockIdx.x * FT_TX;
  const int tid = threadIdx.x;
  // upsampled (pre-resize) rows / columns the tile touches, and the low-resolution window behind them
  int i0, i1;
  float lam;
  bilinear_src(max(Y0 - 1, 0), sy, Hu, i0, i1, lam);
  const int uy_lo = i0;
  bilinear_src(min(Y0 + FT_TY, H - 1), sy, Hu, i0, i1, lam);
  const int nuy = i1 - uy_lo + 1;
  bilinear_src(max(X0 - 1, 0), sx, Wu, i0, i1, lam);
  const int ux_lo = i0;
  bilinear_src(min(X0 + FT_TX, W - 1), sx, Wu, i0, i1, lam);
  const int nux = i1 - ux_lo + 1;
  const int ty_lo = ((uy_lo + 1) >> 1) - 2, nty = ((uy_lo + nuy) >> 1) + 1 - ty_lo + 1;
  const int tx_lo = ((ux_lo + 1) >> 1) - 2, ntx = ((ux_lo + nux) >> 1) + 1 - tx_lo + 1;
  // (the host checks nty <= FT_NTY, ntx <= FT_NTX for the given sizes)
  for (int i = tid; i < nty * ntx; i += 256) {
    const int ry = i / ntx, rx = i - ry * ntx;
    const int sy_ = min(max(ty_lo + ry, 0), h - 1), sx_ = min(max(tx_lo + rx, 0), w - 1);   // replicate padding
    const float4 *src = reinterpret_cast<const float4 *>(t12 + (((int64_t)b * h + sy_) * w + sx_) * 12);
    const float4 a0 = src[0], a1 = src[1], a2 = src[2];
    T[0][ry][rx] = a0.x; T[1][ry][rx] = a0.y; T[2][ry][rx] = a0.z; T[3][ry][rx] = a0.w;
    T[4][ry][rx] = a1.x; T[5][ry][rx] = a1.y; T[6][ry][rx] = a1.z; T[7][ry][rx] = a1.w;
    T[8][ry][rx] = a2.x;
  }
  // own output pixels: rows tid/64 + 4k, column tid%64
  const int lx = tid & 63, ly = tid >> 6;
  if (tid < 3 * FT_TY) {                                    // y windows of the tile's rows under the three row shifts
    const int d = tid / FT_TY, r = tid - d * FT_TY;
    const int Ys = Y0 + r + d - 1;
    int base;
    float a[5];
    fused_axis_window(min(max(Ys, 0), H - 1), Ys >= 0 && Ys < H && Y0 + r < H, sy, Hu, ty_lo, base, a);
    float *dst = WY + tid * 8;
    dst[0] = a[0]; dst[1] = a[1]; dst[2] = a[2]; dst[3] = a[3]; dst[4] = a[4];
    dst[5] = __int_as_float(base);
  }
  int bx[3], bx4[3];
  float ax[3][5];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const int Xs = X0 + lx + d - 1;
    fused_axis_window(min(max(Xs, 0), W - 1), Xs >= 0 && Xs < W && X0 + lx < W, sx, Wu, tx_lo, bx[d], ax[d]);
    bx4[d] = min(bx[d] + 4, ntx - 1);                        // (its weight is zero whenever the clamp acts)
  }
  __syncthreads();
  // row pass: R[tap][ry][X] = x window of the tap's column shift applied to the tap map's row ry
  for (int ry = ly; ry < nty; ry += 4) {
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int dx = tap % 3;
      const float *row = &T[tap][ry][0];
      const float *r0 = row + bx[dx];
      R[tap][ry][lx] = ((ax[dx][0] * r0[0] + ax[dx][1] * r0[1]) + (ax[dx][2] * r0[2] + ax[dx][3] * r0[3])) + ax[dx][4] * row[bx4[dx]];
    }
  }
  __syncthreads();
  // column pass: the y window of the tap's row shift over R, summed over the taps
  const float bv = bias ? bias[0] : 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int Yl = ly + 4 * k;
    float acc = 0.f;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const float4 a03 = *reinterpret_cast<const float4 *>(WY + (dy * FT_TY + Yl) * 8);
      const float2 a4b = *reinterpret_cast<const float2 *>(WY + (dy * FT_TY + Yl) * 8 + 4);
      const int base = __float_as_int(a4b.y), r4 = min(base + 4, nty - 1);
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const float *col = &R[dy * 3 + dx][0][lx];
        acc += ((a03.x * col[base * FT_TX] + a03.y * col[(base + 1) * FT_TX]) +
                (a03.z * col[(base + 2) * FT_TX] + a03.w * col[(base + 3) * FT_TX])) + a4b.x * col[r4 * FT_TX];
      }
    }
    const int Y = Y0 + Yl;
    if (X0 + lx < W && Y < H) out[((int64_t)b * H + Y) * W + X0 + lx] = acc + bv;
  }
}

// host-side check that every tile of the given problem fits the shared-memory windows of upsample_tapsum_kernel
static bool upsample_tapsum_fits(int h, int w, int H, int W) {
  const int Hu = 2 * h, Wu = 2 * w;
  if (Hu < H || Wu < W) return false;
  const double sy = (double)Hu / H, sx = (double)Wu / W;
  const int nuy = (int)((FT_TY + 2) * sy) + 3, nux = (int)((FT_TX + 2) * sx) + 3;
  return nuy <= FT_UY && nux <= FT_UX && nuy / 2 + 5 <= FT_NTY && nux / 2 + 5 <= FT_NTX;
}

// Bicubic resize (ATen upsample_bicubic2d, align_corners=False, A = -0.75): source = scale (dst + 0.5) - 0.5 WITHOUT the
// clamp at 0 the bilinear mode has, 4 x 4 taps at floor(source) - 1 .. + 2 with indices clamped to the image.
__device__ __forceinline__ void cubic_coeffs(float t, float (&w)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x3 = 2.f - t, u = 1.f - t;
  w[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  w[1] = ((A + 2.f) * t - (A + 3.f)) * t * t + 1.f;
  w[2] = ((A + 2.f) * u - (A + 3.f)) * u * u + 1.f;
  w[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}
__global__ void __launch_bounds__(256) resize_bicubic_kernel(const float *__restrict__ x, int B, int H, int W, int C, int ldx,
                                                             float *__restrict__ y, int Ho, int Wo, int ldy, float sh, float sw) {
  const int C4 = C >> 2;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * Ho * Wo * C4) return;
  const int c4 = (int)(idx % C4);
  int64_t r = idx / C4;
  const int ox = (int)(r % Wo); r /= Wo;
  const int oy = (int)(r % Ho);
  const int b = (int)(r / Ho);
  const float fy = sh * ((float)oy + 0.5f) - 0.5f, fx = sw * ((float)ox + 0.5f) - 0.5f;
  const float fy0 = floorf(fy), fx0 = floorf(fx);
  float wy[4], wx[4];
  cubic_coeffs(fy - fy0, wy);
  cubic_coeffs(fx - fx0, wx);
  const int iy = (int)fy0, ix = (int)fx0;
  const float *xb = x + (int64_t)b * H * W * ldx + c4 * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int yy = min(max(iy - 1 + j, 0), H - 1);
    float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int xx = min(max(ix - 1 + i, 0), W - 1);
      const float4 v = *reinterpret_cast<const float4 *>(xb + ((int64_t)yy * W + xx) * ldx);
      row.x = fmaf(wx[i], v.x, row.x); row.y = fmaf(wx[i], v.y, row.y); row.z = fmaf(wx[i], v.z, row.z); row.w = fmaf(wx[i], v.w, row.w);
    }
    acc.x = fmaf(wy[j], row.x, acc.x); acc.y = fmaf(wy[j], row.y, acc.y); acc.z = fmaf(wy[j], row.z, acc.z); acc.w = fmaf(wy[j], row.w, acc.w);
  }
  *reinterpret_cast<float4 *>(y + (((int64_t)b * Ho + oy) * Wo + ox) * ldy + c4 * 4) = acc;
}


// labels[p] = lut[ argmax softmax( p / (1 - p) ) ] over {background = min_i (1 - p_i), p_1 .. p_N} after the clamp to
// [1e-7, 1 - 1e-7] — the single-stage label rule of model/tracker.py:143-150 and ytvos_validation/tracker.py:53-62,106-107
// applied to RAW object probabilities (first maximum wins, as torch.argmax)
__global__ void __launch_bounds__(256) labels_from_probs_kernel(const float *__restrict__ src, int N, int HW, const uint8_t *__restrict__ lut,
                                                                uint8_t *__restrict__ labels) {
  src += (int64_t)blockIdx.y * N * HW;
  labels += (int64_t)blockIdx.y * HW;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const float lo = 1e-7f, hi = 1.f - 1e-7f;
  float bg = INFINITY;
  for (int n = 0; n < N; ++n) bg = fminf(bg, 1.f - fminf(fmaxf(src[(int64_t)n * HW + p], lo), hi));
  const float z0 = bg / (1.f - bg);
  float mx = z0;
  for (int n = 0; n < N; ++n) {
    const float c = fminf(fmaxf(src[(int64_t)n * HW + p], lo), hi);
    mx = fmaxf(mx, c / (1.f - c));
  }
  float den = expf(z0 - mx);
  for (int n = 0; n < N; ++n) {
    const float c = fminf(fmaxf(src[(int64_t)n * HW + p], lo), hi);
    den += expf(c / (1.f - c) - mx);
  }
  float best = expf(z0 - mx) / den;
  int lab = 0;
  for (int n = 0; n < N; ++n) {
    const float c = fminf(fmaxf(src[(int64_t)n * HW + p], lo), hi);
    const float sv = expf(c / (1.f - c) - mx) / den;
    if (sv > best) { best = sv; lab = n + 1; }
  }
  labels[p] = lut[lab];
}


// out[i] = x[i] > thr ? 1 : 0   (binary training labels of the "thresh" update method, ytvos_validation/discriminator.py:364-367)
__global__ void __launch_bounds__(256) threshold_kernel(const float *__restrict__ x, int64_t n, float thr, float *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = x[i] > thr ? 1.f : 0.f;
}


// out[n][p] = sigmoid(logit[n][p]) * (1 - suppress[p])   (per-object probabilities before the merge,
// ytvos_validation/tracker.py:133-139,176-178)
__global__ void __launch_bounds__(256) sigmoid_suppress_kernel(const float *__restrict__ logits, const uint8_t *__restrict__ suppress,
                                                               int N, int HW, float *__restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const float keep = suppress ? (float)(1 - (int)suppress[p]) : 1.f;
  for (int n = 0; n < N; ++n) out[(int64_t)n * HW + p] = (1.f / (1.f + expf(-logits[(int64_t)n * HW + p]))) * keep;
}

}  // namespace frtm

using namespace frtm;

extern "C" int frtm_normalize_u8(const uint8_t *img, int B, int H, int W, float *out, void *stream) {
  FRTM_REQUIRE(img && out && B > 0 && H > 0 && W > 0, "normalize: bad arguments");
  // constants rounded exactly like the reference: `1 / 255 / stds` is evaluated by torch as reciprocal(stds) * (1/255),
  // `-means / stds` as a true division, both in fp32
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  float s[3], b[3];
  for (int i = 0; i < 3; ++i) {
    volatile float rcp = 1.0f / stdv[i];
    s[i] = rcp * (float)(1.0 / 255.0);
    b[i] = -mean[i] / stdv[i];
  }
  const int64_t HW = (int64_t)H * W, total = HW * B;
  normalize_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(img, HW, total, reinterpret_cast<float4 *>(out),
                                                                       s[0], s[1], s[2], b[0], b[1], b[2]);
  FRTM_CHECK_LAUNCH("normalize_u8");
  return FRTM_OK;
}

extern "C" int frtm_stem_patches_u8(const uint8_t *img, int B, int H, int W, void *hi, void *lo, void *stream) {
  FRTM_REQUIRE(img && hi && lo && B > 0 && H > 0 && W > 0, "stem_patches: bad arguments");
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  float s[3], b[3];
  for (int i = 0; i < 3; ++i) {        // same constants, rounded the same way, as frtm_normalize_u8
    volatile float rcp = 1.0f / stdv[i];
    s[i] = rcp * (float)(1.0 / 255.0);
    b[i] = -mean[i] / stdv[i];
  }
  const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
  const int64_t total = (int64_t)B * Ho * Wo * 24;
  stem_patches_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(img, B, H, W, Ho, Wo, (__half *)hi, (__half *)lo, s[0], s[1],
                                                                          s[2], b[0], b[1], b[2]);
  FRTM_CHECK_LAUNCH("stem_patches");
  return FRTM_OK;
}

extern "C" int frtm_maxpool3x3s2_nhwc(const float *x, int B, int H, int W, int C, float *y, float *y_nchw, void *stream) {
  FRTM_REQUIRE(x && y && C % 4 == 0, "maxpool: bad arguments");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const int64_t total = (int64_t)B * Ho * Wo * (C / 4);
  maxpool_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x, B, H, W, C, Ho, Wo, y, y_nchw);
  FRTM_CHECK_LAUNCH("maxpool3x3s2");
  return FRTM_OK;
}

extern "C" int frtm_maxpool3x3s2_split_nhwc(const float *x, int B, int H, int W, int C, float *y, void *y_hi, void *y_lo,
                                            void *stream) {
  FRTM_REQUIRE(x && y_hi && y_lo && C % 4 == 0, "maxpool_split: bad arguments");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const int64_t total = (int64_t)B * Ho * Wo * (C / 4);
  maxpool_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x, B, H, W, C, Ho, Wo, y, nullptr, (__half *)y_hi, (__half *)y_lo);
  FRTM_CHECK_LAUNCH("maxpool3x3s2_split");
  return FRTM_OK;
}

extern "C" int frtm_resize_bilinear_nhwc(const float *x, int B, int H, int W, int C, int ldx, float *y, int Ho, int Wo,
                                         int ldy, int y_coff, int accumulate, void *stream) {
  FRTM_REQUIRE(x && y && B > 0 && C > 0, "resize_bilinear: bad arguments");
  const float sh = (float)H / (float)Ho, sw = (float)W / (float)Wo;
  cudaStream_t st = (cudaStream_t)stream;
  if (C % 4 == 0) {
    const int64_t total = (int64_t)B * Ho * Wo * (C / 4);
    resize_bilinear_kernel<4><<<cdiv(total, 256), 256, 0, st>>>(x, B, H, W, C, ldx, y, Ho, Wo, ldy, y_coff, accumulate, sh, sw);
  } else {
    const int64_t total = (int64_t)B * Ho * Wo * C;
    resize_bilinear_kernel<1><<<cdiv(total, 256), 256, 0, st>>>(x, B, H, W, C, ldx, y, Ho, Wo, ldy, y_coff, accumulate, sh, sw);
  }
  FRTM_CHECK_LAUNCH("resize_bilinear");
  return FRTM_OK;
}

extern "C" int frtm_resize_bicubic_nhwc(const float *x, int B, int H, int W, int C, int ldx, float *y, int Ho, int Wo, int ldy,
                                        void *stream) {
  FRTM_REQUIRE(x && y && B > 0 && C > 0 && C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "resize_bicubic: C, ldx, ldy must be multiples of 4");
  const int64_t total = (int64_t)B * Ho * Wo * (C / 4);
  resize_bicubic_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x, B, H, W, C, ldx, y, Ho, Wo, ldy, (float)H / (float)Ho,
                                                                            (float)W / (float)Wo);
  FRTM_CHECK_LAUNCH("resize_bicubic");
  return FRTM_OK;
}

extern "C" int frtm_pyrup_bicubic_nhwc(const float *x, int B, int H, int W, int C, float *y, void *y_hi, void *y_lo,
                                       void *stream) {
  FRTM_REQUIRE(x && (y || y_hi) && C % 4 == 0 && (!y_hi || y_lo), "pyrup_bicubic: bad arguments");
  const int64_t total = (int64_t)B * (H + 1) * ((W + 2) / 2) * (C / 4);
  pyrup_bicubic_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x, B, H, W, C, y, (__half *)y_hi, (__half *)y_lo);
  FRTM_CHECK_LAUNCH("pyrup_bicubic");
  return FRTM_OK;
}

extern "C" int64_t frtm_global_avgpool_workspace(int B, int HW, int C) {
  return (int64_t)B * cdiv(HW, GAP_CHUNK) * C * sizeof(float);
}

extern "C" int frtm_global_avgpool_nhwc(const float *x, int B, int HW, int C, int ldx, float *out, float *workspace,
                                        int64_t workspace_bytes, void *stream) {
  FRTM_REQUIRE(x && out && workspace, "global_avgpool: null pointer");
  FRTM_REQUIRE(workspace_bytes >= frtm_global_avgpool_workspace(B, HW, C), "global_avgpool: workspace too small");
  const int nchunks = cdiv(HW, GAP_CHUNK);
  cudaStream_t st = (cudaStream_t)stream;
  gap_stage1_kernel<<<dim3(nchunks, B), 256, 0, st>>>(x, HW, C, ldx, nchunks, workspace);
  FRTM_CHECK_LAUNCH("gap_stage1");
  gap_stage2_kernel<<<B, 256, 0, st>>>(workspace, HW, C, nchunks, out);
  FRTM_CHECK_LAUNCH("gap_stage2");
  return FRTM_OK;
}

extern "C" int frtm_cab_gate(const float *sp, const float *dp, const float *w1, const float *b1, const float *w2,
                             const float *b2, int B, int C, float *gate, void *stream) {
  FRTM_REQUIRE(sp && dp && w1 && b1 && w2 && b2 && gate, "cab_gate: null pointer");
  FRTM_REQUIRE(C % 32 == 0 && C <= 256 && (reinterpret_cast<uintptr_t>(w1) & 15) == 0 && (reinterpret_cast<uintptr_t>(w2) & 15) == 0,
               "cab_gate: C must be a multiple of 32 (<= 256) and the weights 16-byte aligned");
  cab_gate_kernel<<<B, 4 * C, 7 * C * sizeof(float), (cudaStream_t)stream>>>(sp, dp, w1, b1, w2, b2, C, gate);
  FRTM_CHECK_LAUNCH("cab_gate");
  return FRTM_OK;
}

extern "C" int64_t frtm_cab_gate_from_maps_workspace(int B, int HWs, int HWd, int C) {
  return (int64_t)B * (cdiv(HWs, GAP_CHUNK) + cdiv(HWd > 0 ? HWd : 0, GAP_CHUNK)) * C * sizeof(float);
}

extern "C" int frtm_cab_gate_from_maps(const float *shallow, int HWs, int lds, const float *deeper, int HWd, int ldd,
                                       const float *deep_pool, int B, int C, const float *w1, const float *b1, const float *w2,
                                       const float *b2, float *gate, float *pool_out, float *workspace, int64_t workspace_bytes,
                                       void *stream) {
  FRTM_REQUIRE(shallow && (deeper || deep_pool) && w1 && b1 && w2 && b2 && gate && workspace, "cab_gate_from_maps: null pointer");
  FRTM_REQUIRE(C == 64 && (reinterpret_cast<uintptr_t>(w1) & 15) == 0 && (reinterpret_cast<uintptr_t>(w2) & 15) == 0,
               "cab_gate_from_maps: C must be 64 and the weights 16-byte aligned (else: global_avgpool + cab_gate)");
  FRTM_REQUIRE(B > 0 && B <= 65535 && HWs > 0 && (!deeper || HWd > 0), "cab_gate_from_maps: bad sizes");
  FRTM_REQUIRE(workspace_bytes >= frtm_cab_gate_from_maps_workspace(B, HWs, deeper ? HWd : 0, C),
               "cab_gate_from_maps: workspace too small");
  const int nca = cdiv(HWs, GAP_CHUNK), ncb = deeper ? cdiv(HWd, GAP_CHUNK) : 0;
  float *parta = workspace, *partb = deeper ? workspace + (int64_t)B * nca * C : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  gap_stage1_pair_kernel<<<dim3(nca + ncb, B), 256, 0, st>>>(shallow, HWs, lds, nca, parta, deeper, HWd, ldd, ncb, partb, C);
  FRTM_CHECK_LAUNCH("gap_stage1_pair");
  cab_pool_gate64_kernel<<<B, 256, 0, st>>>(parta, HWs, nca, partb, HWd, ncb, deep_pool, w1, b1, w2, b2, gate, pool_out);
  FRTM_CHECK_LAUNCH("cab_pool_gate64");
  return FRTM_OK;
}

extern "C" int frtm_cab_apply_nhwc(const float *shallow, const float *gate, const float *deeper, int deeper_is_vector,
                                   int B, int HW, int C, float *out, void *stream) {
  FRTM_REQUIRE(shallow && gate && deeper && out && C % 4 == 0, "cab_apply: bad arguments");
  const int64_t total4 = (int64_t)B * HW * (C / 4);
  cab_apply_kernel<<<cdiv(total4, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(shallow), gate,
                                                                        deeper, deeper_is_vector, HW, C, total4,
                                                                        reinterpret_cast<float4 *>(out));
  FRTM_CHECK_LAUNCH("cab_apply");
  return FRTM_OK;
}

// CAB with the deeper level resized on the fly (seg_network.py:38-39: shallower * gate + F.interpolate(deeper, bilinear)):
// the (B,Hd,Wd,C) map is 4x smaller than the resized one and stays in L2, so the resize kernel's write of the full-size
// map and this kernel's read of it disappear.  Same bilinear arithmetic as resize_bilinear_kernel.
namespace frtm {
__global__ void __launch_bounds__(256) cab_apply_resized_kernel(const float4 *__restrict__ sh, const float *__restrict__ gate,
                                                                const float *__restrict__ deeper, int H, int W, int C, int Hd,
                                                                int Wd, float shy, float shx, int64_t total4,
                                                                float4 *__restrict__ out, __half *__restrict__ y_hi,
                                                                __half *__restrict__ y_lo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int C4 = C / 4;
  const int c = (int)(i % C4) * 4;
  int64_t r = i / C4;
  const int ox = (int)(r % W);
  r /= W;
  const int oy = (int)(r % H);
  const int64_t b = r / H;
  int y0, y1, x0, x1;
  float ly, lx;
  bilinear_src(oy, shy, Hd, y0, y1, ly);
  bilinear_src(ox, shx, Wd, x0, x1, lx);
  const float hy = 1.f - ly, hx = 1.f - lx;
  const float4 p00 = *reinterpret_cast<const float4 *>(deeper + ((b * Hd + y0) * Wd + x0) * C + c);
  const float4 p01 = *reinterpret_cast<const float4 *>(deeper + ((b * Hd + y0) * Wd + x1) * C + c);
  const float4 p10 = *reinterpret_cast<const float4 *>(deeper + ((b * Hd + y1) * Wd + x0) * C + c);
  const float4 p11 = *reinterpret_cast<const float4 *>(deeper + ((b * Hd + y1) * Wd + x1) * C + c);
  float4 d;
  d.x = hy * (hx * p00.x + lx * p01.x) + ly * (hx * p10.x + lx * p11.x);
  d.y = hy * (hx * p00.y + lx * p01.y) + ly * (hx * p10.y + lx * p11.y);
  d.z = hy * (hx * p00.z + lx * p01.z) + ly * (hx * p10.z + lx * p11.z);
  d.w = hy * (hx * p00.w + lx * p01.w) + ly * (hx * p10.w + lx * p11.w);
  const float4 g = *reinterpret_cast<const float4 *>(gate + b * C + c);
  const float4 s = sh[i];
  float4 o;
  o.x = __fadd_rn(__fmul_rn(s.x, g.x), d.x);
  o.y = __fadd_rn(__fmul_rn(s.y, g.y), d.y);
  o.z = __fadd_rn(__fmul_rn(s.z, g.z), d.z);
  o.w = __fadd_rn(__fmul_rn(s.w, g.w), d.w);
  if (out) out[i] = o;
  if (y_hi) {   // split planes of 16*x for the tensor-core conv that follows (the same conversion as split_kernel)
    const float v[4] = {o.x * 16.f, o.y * 16.f, o.z * 16.f, o.w * 16.f};
    __half h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      h[k] = __float2half_rn(v[k]);
      l[k] = __float2half_rn(v[k] - __half2float(h[k]));
    }
    *reinterpret_cast<uint2 *>(y_hi + i * 4) = *reinterpret_cast<const uint2 *>(h);
    *reinterpret_cast<uint2 *>(y_lo + i * 4) = *reinterpret_cast<const uint2 *>(l);
  }
}
}  // namespace frtm

extern "C" int frtm_cab_apply_resized_nhwc(const float *shallow, const float *gate, const float *deeper, int B, int H, int W, int C,
                                           int Hd, int Wd, float *out, void *y_hi, void *y_lo, void *stream) {
  FRTM_REQUIRE(shallow && gate && deeper && (out || y_hi) && (!y_hi || y_lo) && C % 4 == 0 && Hd > 0 && Wd > 0,
               "cab_apply_resized: bad arguments");
  const int64_t total4 = (int64_t)B * H * W * (C / 4);
  frtm::cab_apply_resized_kernel<<<cdiv(total4, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4 *>(shallow), gate, deeper, H, W, C, Hd, Wd, (float)Hd / (float)H, (float)Wd / (float)W, total4,
      reinterpret_cast<float4 *>(out), (__half *)y_hi, (__half *)y_lo);
  FRTM_CHECK_LAUNCH("cab_apply_resized");
  return FRTM_OK;
}

extern "C" int frtm_scatter_channel_nhwc(const float *src, int B, int HW, float *dst, int ld, int coff, int nzero,
                                         void *stream) {
  FRTM_REQUIRE(src && dst && coff + nzero < ld, "scatter_channel: bad arguments");
  const int64_t total = (int64_t)B * HW;
  scatter_channel_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(src, total, dst, ld, coff, nzero);
  FRTM_CHECK_LAUNCH("scatter_channel");
  return FRTM_OK;
}

extern "C" int frtm_broadcast_objects_nhwc(const float *src, int F, int N, int HW, int C, int lds, float *dst, int ldd,
                                           void *stream) {
  FRTM_REQUIRE(src && dst && C % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0, "broadcast_objects: bad arguments");
  const int64_t total4 = (int64_t)F * N * HW * (C / 4);
  broadcast_objects_kernel<<<cdiv(total4, 256), 256, 0, (cudaStream_t)stream>>>(src, N, HW, C, lds, dst, ldd, total4);
  FRTM_CHECK_LAUNCH("broadcast_objects");
  return FRTM_OK;
}

extern "C" int frtm_nhwc_to_nchw(const float *x, int B, int HW, int C, int ldx, float *y, void *stream) {
  FRTM_REQUIRE(x && y, "nhwc_to_nchw: null pointer");
  dim3 grid(cdiv(HW, 32), cdiv(C, 32), B);
  nhwc_to_nchw_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(x, HW, C, ldx, y);
  FRTM_CHECK_LAUNCH("nhwc_to_nchw");
  return FRTM_OK;
}

extern "C" int frtm_merge_masks(const float *src, uint64_t logit_mask, const uint8_t *suppress, int N, int HW,
                                const uint8_t *lut, int single_object, float *masks, uint8_t *labels, int *counts,
                                void *stream) {
  FRTM_REQUIRE(src && lut && masks && labels && counts && N > 0 && N <= 64, "merge_masks: bad arguments");
  if (N <= 8)
    merge_masks_reg_kernel<8><<<cdiv(HW, 256), 256, 0, (cudaStream_t)stream>>>(src, (unsigned long long)logit_mask, suppress, N, HW,
                                                                               lut, single_object, masks, labels, counts, 0);
  else
    merge_masks_kernel<<<cdiv(HW, 256), 256, 0, (cudaStream_t)stream>>>(src, (unsigned long long)logit_mask, suppress, N,
                                                                        HW, lut, single_object, masks, labels, counts, 0);
  FRTM_CHECK_LAUNCH("merge_masks");
  return FRTM_OK;
}

extern "C" int frtm_sigmoid_suppress(const float *logits, const uint8_t *suppress, int N, int HW, float *out, void *stream) {
  FRTM_REQUIRE(logits && out && N > 0 && HW > 0, "sigmoid_suppress: bad arguments");
  sigmoid_suppress_kernel<<<cdiv(HW, 256), 256, 0, (cudaStream_t)stream>>>(logits, suppress, N, HW, out);
  FRTM_CHECK_LAUNCH("sigmoid_suppress");
  return FRTM_OK;
}

extern "C" int frtm_threshold_f32(const float *x, int64_t n, float thr, float *out, void *stream) {
  FRTM_REQUIRE(x && out && n >= 0, "threshold_f32: bad arguments");
  if (n == 0) return FRTM_OK;
  threshold_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, thr, out);
  FRTM_CHECK_LAUNCH("threshold_f32");
  return FRTM_OK;
}

extern "C" int frtm_labels_from_probs(const float *src, int F, int N, int HW, const uint8_t *lut, uint8_t *labels, void *stream) {
  FRTM_REQUIRE(src && lut && labels && F > 0 && N > 0 && HW > 0, "labels_from_probs: bad arguments");
  labels_from_probs_kernel<<<dim3(cdiv(HW, 256), F), 256, 0, (cudaStream_t)stream>>>(src, N, HW, lut, labels);
  FRTM_CHECK_LAUNCH("labels_from_probs");
  return FRTM_OK;
}

extern "C" int frtm_merge_masks_frames(const float *src, int F, uint64_t logit_mask, int N, int HW, const uint8_t *lut,
                                       int single_object, float *masks, uint8_t *labels, int *counts, int counts_stride,
                                       void *stream) {
  FRTM_REQUIRE(src && lut && masks && labels && counts && N > 0 && N <= 64 && F > 0 && F <= 65535 && counts_stride >= N,
               "merge_masks_frames: bad arguments");
  if (N <= 8)
    merge_masks_reg_kernel<8><<<dim3(cdiv(HW, 256), F), 256, 0, (cudaStream_t)stream>>>(src, (unsigned long long)logit_mask, nullptr, N,
                                                                                        HW, lut, single_object, masks, labels,
                                                                                        counts, counts_stride);
  else
    merge_masks_kernel<<<dim3(cdiv(HW, 256), F), 256, 0, (cudaStream_t)stream>>>(src, (unsigned long long)logit_mask, nullptr, N, HW,
                                                                                 lut, single_object, masks, labels, counts,
                                                                                 counts_stride);
  FRTM_CHECK_LAUNCH("merge_masks_frames");
  return FRTM_OK;
}

extern "C" int frtm_tapmaps_nhwc(const float *x, int64_t npix, int C, const float *w9c, float *y12, void *stream) {
  FRTM_REQUIRE(x && w9c && y12 && C % 4 == 0 && C <= 1024, "tapmaps: bad arguments");
  tapmaps_kernel<<<cdiv(npix, 256), 256, 9 * C * sizeof(float), (cudaStream_t)stream>>>(x, npix, C, w9c, y12);
  FRTM_CHECK_LAUNCH("tapmaps");
  return FRTM_OK;
}

extern "C" int frtm_shift_sum9(const float *v12, int B, int H, int W, const float *bias, float *out, void *stream) {
  FRTM_REQUIRE(v12 && out, "shift_sum9: null pointer");
  shift_sum9_kernel<<<cdiv((int64_t)B * H * W, 256), 256, 0, (cudaStream_t)stream>>>(v12, B, H, W, bias, out);
  FRTM_CHECK_LAUNCH("shift_sum9");
  return FRTM_OK;
}

extern "C" int frtm_upsample_tapsum(const float *t12, int B, int h, int w, int H, int W, const float *bias, float *out,
                                    void *stream) {
  FRTM_REQUIRE(t12 && out && B > 0, "upsample_tapsum: bad arguments");
  FRTM_REQUIRE(upsample_tapsum_fits(h, w, H, W), "upsample_tapsum: (%d,%d) -> x2 -> (%d,%d) does not fit the fused tile windows", h,
               w, H, W);
  FRTM_REQUIRE(B <= 65535, "upsample_tapsum: batch too large");
  const dim3 grid(cdiv(W, FT_TX), cdiv(H, FT_TY), B);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(upsample_tapsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM);
    if (e != cudaSuccess) { set_error("upsample_tapsum: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return FRTM_ELAUNCH; }
    attr_set = true;
  }
  upsample_tapsum_kernel<<<grid, 256, FT_SMEM, (cudaStream_t)stream>>>(t12, B, h, w, H, W, bias, out);
  FRTM_CHECK_LAUNCH("upsample_tapsum");
  return FRTM_OK;
}
extern "C" int64_t frtm_upsample_tapsum_supported(int h, int w, int H, int W) { return upsample_tapsum_fits(h, w, H, W) ? 1 : 0; }
