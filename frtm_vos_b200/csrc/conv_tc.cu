// Tensor-core convolution for sm_100a: tcgen05.mma (kind::f16) with fp32 accumulation in TMEM, operands staged in
// shared memory by TMA (activations: 4-D tiled tensor maps with zero OOB fill = the conv's zero padding; weights:
// pre-swizzled tiles fetched with one bulk async copy), mbarrier producer/consumer pipeline, warp-specialised roles.
//
// Precision: the reference computes these convs in fp32 and the parity budget is 1e-3 on the logits, which no single
// fp16/bf16/tf32 pass meets.  Operands are therefore split x = hi + lo (two fp16 planes, x pre-scaled
// by 2^4, weights pre-scaled per output channel by a power of two so `lo` stays out of the fp16 subnormals) and each
// k-step issues three MMAs into the same accumulator:  hi*hi + hi*lo + lo*hi  (the lo*lo term is < 2^-22 relative).
// The tensor-core accumulator TRUNCATES on every accumulate (measured on B200: all-positive operands give a relative
// bias of -K * 2^-27, i.e. -1.7e-5 at K=2304), which compounds over ResNet-101's depth past the 1e-3 logit budget.  The
// accumulation is therefore two-level: the MMA warp accumulates TC_FOLD k-blocks into one of two TMEM slots, the
// epilogue warps fold each finished slot into fp32 registers with round-to-nearest adds while the MMA warp fills the
// other slot.  The bias becomes ~3*TC_FOLD ulp, independent of K.
// The epilogue undoes the power-of-two scales exactly, adds bias / residual, applies ReLU and writes fp32 NHWC and/or
// the split planes for the next conv and/or an NCHW copy.
//
// GEMM view per CTA: D[128 pixels (8x16 spatial tile), BN couts] += A[128, 64] * B[64, BN] per (filter tap, 64-channel
// chunk).  Stride-1 "same" convolutions only (1x1 and 3x3 here); everything else stays on conv_simt.cu.
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include "tc_ptx.cuh"

namespace frtm {

constexpr int TC_TH = 8, TC_TW = 16;          // spatial tile -> 128 GEMM rows
constexpr int TC_BK = 64;                     // fp16 elements per k-block = one 128-byte swizzle row
constexpr int TC_A_BYTES = 128 * 128;         // one A plane tile (128 rows x 128 B)
constexpr float TC_ACT_SCALE = 16.f;
constexpr int TC_FOLD = 2;                   // k-blocks accumulated inside the tensor core between fp32 register folds

struct TcArgs {
  const __half *wt;        // [ntile][tap][kchunk][hi|lo][BN rows x 128 B, SW128 image]
  const float *oscale;     // [CoutPad] = 1 / (act_scale * weight_scale[n])   (exact powers of two)
  const float *bias;       // [Cout] or null
  const float *res;        // fp32 NHWC residual or null
  const __half *res_hi, *res_lo;  // or split residual (scaled by TC_ACT_SCALE), channel stride ldrh
  float *y;                // fp32 NHWC out (ldy, y_coff) or null
  float *y_nchw;           // or null
  __half *y_hi, *y_lo;     // split out planes (channel stride ldyh, offset yh_coff) or null
  const float *r1_score;   // (B,Ho,Wo) fp32: a 65th input channel handled as an fp32 rank-1 term (TSE.transform), or null
  const float *r1_w;       // [9][Cout] its 3x3 weights
  const float *r1_bias;    // [Cout] bias added after the rank-1 term, or null
  float *y_extra;          // (B,Ho,Wo): receives output channel `extra_ch` (the next conv's score channel), or null
  int extra_ch, yh_cout;   // yh_cout: number of leading output channels written to the split planes
  const float *tapw;       // [9][Cout] weights of a following 3x3 -> 1 conv, contracted per pixel in the epilogue, or null
  float *y_tap;            // (B,Ho,Wo,12): the 9 tap maps  sum_c tapw[tap][c] * out[c]  (needs Cout <= BN), or null
  int ldr, ldrh, ldy, y_coff, ldyh, yh_coff;
  int B, H, W, Ho, Wo, stride, Cout, kh, kw, pad, relu, tiles_x, tiles_y, kchunks;
};


// ----------------------------------------------------------------------------------------------------------------
// kernel: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue (warp 2 owns the TMEM allocation)
// ----------------------------------------------------------------------------------------------------------------
// ----------------------------------------------------------------------------------------------------------------
// epilogue pieces shared by the tensor-core conv kernels (threads 64..191 = the four epilogue warps)
// ----------------------------------------------------------------------------------------------------------------
// shared-memory staging of the per-channel scale / bias (and the optional rank-1 / tap weights) of N tile n0
template <int BN, bool R1, bool TAP>
__device__ __forceinline__ void tc_epilogue_stage(const TcArgs &a, int n0, float *s_osc) {
  float *s_bias = s_osc + BN;
  float *s_tap = s_bias + BN;             // [9][BN]
  float *s_r1 = s_tap + 9 * BN;           // [10][BN]: rank-1 weights of the 9 taps, then the bias that follows them
  for (int i = threadIdx.x - 64; i < BN; i += 128) {
    s_osc[i] = a.oscale[n0 + i];
    s_bias[i] = (a.bias != nullptr && n0 + i < a.Cout) ? a.bias[n0 + i] : 0.f;
  }
  if constexpr (TAP) {
    for (int i = threadIdx.x - 64; i < 9 * BN; i += 128) {
      const int t = i / BN, ch = i - t * BN;
      s_tap[i] = (n0 + ch < a.Cout) ? a.tapw[t * a.Cout + n0 + ch] : 0.f;
    }
  }
  if constexpr (R1) {
    for (int i = threadIdx.x - 64; i < 10 * BN; i += 128) {
      const int t = i / BN, ch = i - t * BN;
      float v = 0.f;
      if (n0 + ch < a.Cout) v = t < 9 ? a.r1_w[t * a.Cout + n0 + ch] : (a.r1_bias ? a.r1_bias[n0 + ch] : 0.f);
      s_r1[i] = v;
    }
  }
}

// output pixel (b, py, px) with its BN accumulated channels: unscale, bias, rank-1 term, residual, ReLU, all outputs
template <int BN, bool R1, bool TAP>
__device__ __forceinline__ void tc_epilogue_store(const TcArgs &a, const float (&acc)[BN], int b, int py, int px, int n0,
                                                  const float *s_osc) {
  const float *s_bias = s_osc + BN;
  const float *s_tap = s_bias + BN;
  const float *s_r1 = s_tap + 9 * BN;
  const bool valid = py < a.Ho && px < a.Wo;
  const int64_t pix = ((int64_t)b * a.Ho + py) * a.Wo + px;
  float sv[9];
  if constexpr (R1) {
    const float *sb = a.r1_score + (int64_t)b * a.Ho * a.Wo;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
      sv[t] = (valid && yy >= 0 && yy < a.Ho && xx >= 0 && xx < a.Wo) ? sb[yy * a.Wo + xx] : 0.f;
    }
  }
    const bool vec_f32 = (a.ldy % 4 == 0) && (a.y_coff % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.y) & 15) == 0);
    const bool vec_res = (a.ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.res) & 15) == 0);
    const bool vec_h = (a.ldyh % 8 == 0) && (a.yh_coff % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.y_hi) & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(a.y_lo) & 15) == 0);
    const bool vec_rh = (a.ldrh % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.res_hi) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(a.res_lo) & 15) == 0);
    float tp[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) tp[t] = 0.f;
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = acc[c0 + j];
      if (!valid || n0 + c0 >= a.Cout) continue;
      const bool full = n0 + c0 + 16 <= a.Cout;
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaf(v[j], s_osc[c0 + j], s_bias[c0 + j]);
      if constexpr (R1) {     // same order as frtm_rank1_finish: taps 0..8, then the bias
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaf(s_r1[t * BN + c0 + j], sv[t], v[j]);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += s_r1[9 * BN + c0 + j];
      }
      if (a.res) {
        const float *r = a.res + pix * a.ldr + n0 + c0;
        if (full && vec_res) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 t = *reinterpret_cast<const float4 *>(r + j);
            v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + c0 + j < a.Cout) v[j] += r[j];
        }
      }
      if (a.res_hi) {
        const __half *rh = a.res_hi + pix * a.ldrh + n0 + c0, *rl = a.res_lo + pix * a.ldrh + n0 + c0;
        if (full && vec_rh) {
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            const uint4 uh = *reinterpret_cast<const uint4 *>(rh + j), ul = *reinterpret_cast<const uint4 *>(rl + j);
            const __half2 *h2 = reinterpret_cast<const __half2 *>(&uh), *l2 = reinterpret_cast<const __half2 *>(&ul);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 fh = __half22float2(h2[k]), fl = __half22float2(l2[k]);
              v[j + 2 * k] += (fh.x + fl.x) * (1.f / TC_ACT_SCALE);
              v[j + 2 * k + 1] += (fh.y + fl.y) * (1.f / TC_ACT_SCALE);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + c0 + j < a.Cout) v[j] += (__half2float(rh[j]) + __half2float(rl[j])) * (1.f / TC_ACT_SCALE);
        }
      }
      if (a.relu) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if constexpr (TAP) {
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
          for (int j = 0; j < 16; ++j) tp[t] = fmaf(v[j], s_tap[t * BN + c0 + j], tp[t]);
      }
      if (a.y) {
        float *dst = a.y + pix * a.ldy + a.y_coff + n0 + c0;
        if (full && vec_f32) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4 *>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + c0 + j < a.Cout) dst[j] = v[j];
        }
      }
      if (a.y_hi) {
        __half *dh = a.y_hi + pix * a.ldyh + a.yh_coff + n0 + c0, *dl = a.y_lo + pix * a.ldyh + a.yh_coff + n0 + c0;
        __align__(16) __half hh[16];
        __align__(16) __half ll[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float sc = v[j] * TC_ACT_SCALE;
          hh[j] = __float2half_rn(sc);
          ll[j] = __float2half_rn(sc - __half2float(hh[j]));
        }
        if (n0 + c0 + 16 <= a.yh_cout && vec_h) {
          *reinterpret_cast<uint4 *>(dh) = *reinterpret_cast<const uint4 *>(hh);
          *reinterpret_cast<uint4 *>(dh + 8) = *reinterpret_cast<const uint4 *>(hh + 8);
          *reinterpret_cast<uint4 *>(dl) = *reinterpret_cast<const uint4 *>(ll);
          *reinterpret_cast<uint4 *>(dl + 8) = *reinterpret_cast<const uint4 *>(ll + 8);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + c0 + j < a.yh_cout) { dh[j] = hh[j]; dl[j] = ll[j]; }
        }
      }
      if (R1 && a.y_extra && a.extra_ch >= n0 + c0 && a.extra_ch < n0 + c0 + 16) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n0 + c0 + j == a.extra_ch) a.y_extra[pix] = v[j];
      }
      if (a.y_nchw) {
        const int64_t hw = (int64_t)a.Ho * a.Wo, p = (int64_t)py * a.Wo + px;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n0 + c0 + j < a.Cout) a.y_nchw[((int64_t)b * a.Cout + n0 + c0 + j) * hw + p] = v[j];
      }
    }
    if (TAP && valid) {
      float4 *o = reinterpret_cast<float4 *>(a.y_tap + pix * 12);
      o[0] = make_float4(tp[0], tp[1], tp[2], tp[3]);
      o[1] = make_float4(tp[4], tp[5], tp[6], tp[7]);
      o[2] = make_float4(tp[8], 0.f, 0.f, 0.f);
    }
}

// PACK: issue the passes as A_hi x [B_hi | B_lo] (N = 2*BN) + A_lo x B_hi — needs 4*BN TMEM columns; tiles whose
// 4*BN would push the allocation to all 512 columns (one CTA per SM) keep the three separate N = BN products.
// R1 / TAP: epilogue features compiled in only where used (rank-1 score channel; tap-map contraction) — the unrolled
// epilogue is most of the kernel's code, and the plain instances must not pay for them in instruction-cache misses.
template <int BN, int STAGES, bool R1, bool TAP>
__global__ void __launch_bounds__(192) conv_tc_kernel(const __grid_constant__ CUtensorMap tm_hi,
                                                      const __grid_constant__ CUtensorMap tm_lo, const TcArgs a) {
  constexpr int B_BYTES = BN * 128;
  constexpr int STAGE_BYTES = 2 * TC_A_BYTES + 2 * B_BYTES;
  constexpr bool PACK = (4 * BN <= 256) || BN == 128;
  constexpr int SLOT = PACK ? 2 * BN : BN;     // accumulator slot: [hi*hi + lo*hi | hi*lo] or one sum
  constexpr int COLS = tmem_cols(2 * SLOT);    // two accumulator slots
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_full = base + STAGES * STAGE_BYTES;       // STAGES x 8 B
  const uint32_t bar_empty = bar_full + 8 * STAGES;
  const uint32_t bar_accf = bar_empty + 8 * STAGES;            // 2 x 8 B: accumulator slot full
  const uint32_t bar_acce = bar_accf + 16;                     // 2 x 8 B: accumulator slot drained
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gen + STAGES * STAGE_BYTES + 16 * STAGES + 32);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int tiles_per_img = a.tiles_x * a.tiles_y;
  const int b = blockIdx.x / tiles_per_img;
  const int tr = blockIdx.x - b * tiles_per_img;
  const int y0 = (tr / a.tiles_x) * TC_TH, x0 = (tr % a.tiles_x) * TC_TW;
  const int ntile = blockIdx.y;
  const int ntaps = a.kh * a.kw;
  const int nkb = ntaps * a.kchunks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_lo)) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int j = 0; j < 2; ++j) {
      mbar_init(bar_accf + 8 * j, 1);
      mbar_init(bar_acce + 8 * j, 4);      // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Producer and MMA warps walk their loops converged and issue through one elected lane (elect.sync): control flow and
  // descriptor arithmetic stay warp-uniform, so the TMA / tcgen05 instructions are issued back to back from uniform
  // registers instead of through the per-lane serialisation loops an `if (lane == 0)` region compiles to.
  if (warp == 0) {
    {
      const __half *wbase = a.wt + (size_t)ntile * nkb * (2 * B_BYTES / 2);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        const int tap = kb / a.kchunks, kc = kb - tap * a.kchunks;
        const int ky = tap / a.kw, kx = tap - ky * a.kw;
        const uint32_t sa = base + s * STAGE_BYTES;
        if (elect_one()) {
          mbar_expect_tx(bar_full + 8 * s, STAGE_BYTES);
          tma_load_4d(sa, &tm_hi, kc * TC_BK, x0 * a.stride + kx - a.pad, y0 * a.stride + ky - a.pad, b, bar_full + 8 * s);
          tma_load_4d(sa + TC_A_BYTES, &tm_lo, kc * TC_BK, x0 * a.stride + kx - a.pad, y0 * a.stride + ky - a.pad, b, bar_full + 8 * s);
          bulk_load(sa + 2 * TC_A_BYTES, wbase + (size_t)kb * (2 * B_BYTES / 2), 2 * B_BYTES, bar_full + 8 * s);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    {
      // per k-step:  D[0:2BN) (+)= A_hi x [B_hi | B_lo]  (one N = 2*BN product: the hi and lo weight rows are contiguous in
      // the stage) and  D[0:BN) += A_lo x B_hi  — two reads of the A tile per k-step instead of three
      constexpr uint32_t idesc = umma_idesc(128, BN), idesc2 = umma_idesc(128, 2 * BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(bar_full + 8 * s, ph);
        tc_fence_after();
        const int grp = kb / TC_FOLD, first = (kb % TC_FOLD) == 0;
        const uint32_t slot = grp & 1;
        if (first) {                        // the epilogue must have drained this slot (two groups ago)
          mbar_wait(bar_acce + 8 * slot, ((grp >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        const uint32_t tacc = tmem_base + slot * SLOT;
        const uint32_t sa = base + s * STAGE_BYTES;
        const uint64_t a_hi = umma_desc(sa), a_lo = umma_desc(sa + TC_A_BYTES);
        const uint64_t b_hl = umma_desc(sa + 2 * TC_A_BYTES), b_lo = umma_desc(sa + 2 * TC_A_BYTES + B_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);     // +32 bytes along K inside the swizzle atom
            if (PACK) {
              umma_f16(tacc, a_hi + adv, b_hl + adv, idesc2, (first && k == 0) ? 0u : 1u);
              umma_f16(tacc, a_lo + adv, b_hl + adv, idesc, 1u);
            } else {
              umma_f16(tacc, a_hi + adv, b_hl + adv, idesc, (first && k == 0) ? 0u : 1u);
              umma_f16(tacc, a_hi + adv, b_lo + adv, idesc, 1u);
              umma_f16(tacc, a_lo + adv, b_hl + adv, idesc, 1u);
            }
          }
          umma_commit(bar_empty + 8 * s);     // frees the stage once the MMAs above have consumed it
          if ((kb % TC_FOLD) == TC_FOLD - 1 || kb == nkb - 1) umma_commit(bar_accf + 8 * slot);   // slot ready to fold
        }
        __syncwarp();
      }
    }
  } else {
    // ---- epilogue: TMEM lane = tile row = pixel; a warp may only touch lanes 32*(warp%4) .. +31 ----
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int py = y0 + row / TC_TW, px = x0 + row % TC_TW;
    const int n0 = ntile * BN;
    // per-channel output scale and bias (and the optional rank-1 / tap weights) staged once per CTA (overlaps the main loop)
    float *s_osc = reinterpret_cast<float *>(gen + STAGES * STAGE_BYTES + 16 * STAGES + 48);
    tc_epilogue_stage<BN, R1, TAP>(a, n0, s_osc);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    // fold every finished accumulator slot into fp32 registers (round-to-nearest adds)
    float acc[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) acc[j] = 0.f;
    const int ngrp = (nkb + TC_FOLD - 1) / TC_FOLD;
#pragma unroll 1
    for (int grp = 0; grp < ngrp; ++grp) {
      const uint32_t slot = grp & 1;
      mbar_wait(bar_accf + 8 * slot, (grp >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 16) {
        float t[16];
        const uint32_t col = tmem_base + ((uint32_t)(q * 32) << 16) + slot * SLOT + (uint32_t)c0;
        tmem_ld16(col, t);
        if (PACK) {
          float t2[16];
          tmem_ld16(col + BN, t2);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[c0 + j] += t[j] + t2[j];
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[c0 + j] += t[j];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_acce + 8 * slot) : "memory");
    }
    tc_epilogue_store<BN, R1, TAP>(a, acc, b, py, px, n0, s_osc);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(COLS) : "memory");
  }
}

// ----------------------------------------------------------------------------------------------------------------
// CTA-pair kernel for the wide convolutions (Cout a multiple of 128: the bottleneck stages of ResNet-101 and the deep
// stages of ResNet-18 — 1x1 and 3x3, stride 1 and 2).  The general kernel above moves 48 KB from L2 per 128x64x64
// product and is bound by the L2 -> shared-memory path (~42 B/clk/SM) at a quarter of the tensor rate.  Here two CTAs of
// a cluster (the two SMs of a TPC) run ONE tcgen05.mma.cta_group::2 product of M = 256 pixels x N = 2*HALF output
// channels: each CTA stages its own 128-pixel A tile and HALF of the weight rows, so a k-block costs a CTA 32 + HALF/4 KB
// for 128 x 2*HALF x 64 MACs — a third (HALF = 128) or half (HALF = 64) of the general kernel's traffic per MAC — and
// every product runs at the full N >= 128 issue rate.  The three split products hi*hi + hi*lo + lo*hi go into the same
// accumulator; accumulation is two-level as above (TC_FOLD k-blocks per TMEM slot, eight epilogue warps fold the slots into
// fp32 registers: warp = (lane quarter, column half)).  Protocol: both producers load into their own shared memory and
// count the bytes on the LEADER's full barrier; the leader's MMA warp issues for the pair and its commits arrive on the
// empty / accumulator-full barriers of both CTAs (multicast); the epilogue warps of both CTAs arrive on the leader's
// accumulator-drained barrier.
// Weights: the N tile 64 packing of TcArgs::wt, read through a 2-D tensor map (rows of 128 B, pre-swizzled image).
// ----------------------------------------------------------------------------------------------------------------
constexpr int P2_THREADS = 64 + 8 * 32;
#ifndef P2_FOLD_N
#define P2_FOLD_N 1
#endif
// k-blocks per TMEM slot between register folds.  The three split products share one accumulator here (12 truncating
// accumulations per k-block against 8 at the main magnitude in the packed scheme of the general kernel), so the pair
// kernel folds every k-block: 12 per fold instead of 16.
constexpr int P2_FOLD = P2_FOLD_N;

// Epilogue of a 128-pixel x BN-channel tile staged through shared memory (the pipeline stages are free once the last
// accumulator slot has been committed): the TMEM-row-per-thread layout of the accumulators would make every global access a
// 32 x 16-byte scatter and needs a fully unrolled store routine (tens of KB of code, instruction-cache bound); through the
// tile a small rolled loop reads / writes 8 consecutive channels per thread, i.e. whole 128-byte lines per quarter warp.
// `tile` = [128][BN + 4] fp32 holding acc * oscale + bias; called by the 256 epilogue threads (te = 0..255) after a barrier.
template <int BN>
__device__ __forceinline__ void pair_epilogue_store(const TcArgs &a, float *tile, int te, int b, int y0, int x0, int n0) {
  constexpr int LDT = BN + 4, G = BN / 8, ITERS = 128 * G / 256, U = 4;
  static_assert(ITERS % U == 0, "pair_epilogue_store: iteration count");
  const bool vec_f32 = (a.ldy % 4 == 0) && (a.y_coff % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.y) & 15) == 0);
  const bool vec_res = (a.ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.res) & 15) == 0);
  const bool vec_h = (a.ldyh % 8 == 0) && (a.yh_coff % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.y_hi) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(a.y_lo) & 15) == 0);
  const bool vec_rh = (a.ldrh % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.res_hi) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(a.res_lo) & 15) == 0);
  const bool post = a.res || a.res_hi || a.relu;
  // U iterations at a time with all their residual loads issued first: one load round trip per thread and iteration would
  // leave the loop latency-bound (8 KB in flight per SM)
#pragma unroll 1
  for (int it = 0; it < ITERS; it += U) {
    int64_t pix[U];
    int pp[U], cc[U];
    bool ok[U];
    uint4 rh4[U], rl4[U];
    float4 rf[U][2];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = te + (it + u) * 256;
      pp[u] = idx / G;
      cc[u] = (idx - pp[u] * G) * 8;
      const int py = y0 + pp[u] / TC_TW, px = x0 + pp[u] % TC_TW;
      ok[u] = py < a.Ho && px < a.Wo && b < a.B;
      pix[u] = ((int64_t)b * a.Ho + py) * a.Wo + px;
      rh4[u] = make_uint4(0u, 0u, 0u, 0u); rl4[u] = rh4[u];
      rf[u][0] = make_float4(0.f, 0.f, 0.f, 0.f); rf[u][1] = rf[u][0];
      if (ok[u] && a.res_hi) {
        const __half *rh = a.res_hi + pix[u] * a.ldrh + n0 + cc[u], *rl = a.res_lo + pix[u] * a.ldrh + n0 + cc[u];
        if (vec_rh) {
          rh4[u] = *reinterpret_cast<const uint4 *>(rh);
          rl4[u] = *reinterpret_cast<const uint4 *>(rl);
        } else {
          __half *hh = reinterpret_cast<__half *>(&rh4[u]), *ll = reinterpret_cast<__half *>(&rl4[u]);
#pragma unroll
          for (int j = 0; j < 8; ++j) { hh[j] = rh[j]; ll[j] = rl[j]; }
        }
      }
      if (ok[u] && a.res) {
        const float *r = a.res + pix[u] * a.ldr + n0 + cc[u];
        if (vec_res) {
          rf[u][0] = *reinterpret_cast<const float4 *>(r);
          rf[u][1] = *reinterpret_cast<const float4 *>(r + 4);
        } else {
          float *f = reinterpret_cast<float *>(&rf[u][0]);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = r[j];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
      const int c = cc[u];
      float *tp = tile + pp[u] * LDT + c;
      float v[8];
      *reinterpret_cast<float4 *>(v) = *reinterpret_cast<const float4 *>(tp);
      *reinterpret_cast<float4 *>(v + 4) = *reinterpret_cast<const float4 *>(tp + 4);
      {
        const float *f = reinterpret_cast<const float *>(&rf[u][0]);
        const __half2 *h2 = reinterpret_cast<const __half2 *>(&rh4[u]), *l2 = reinterpret_cast<const __half2 *>(&rl4[u]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 fh = __half22float2(h2[k]), fl = __half22float2(l2[k]);
          v[2 * k] += f[2 * k] + (fh.x + fl.x) * (1.f / TC_ACT_SCALE);
          v[2 * k + 1] += f[2 * k + 1] + (fh.y + fl.y) * (1.f / TC_ACT_SCALE);
        }
      }
      if (a.relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (a.y) {
        float *dst = a.y + pix[u] * a.ldy + a.y_coff + n0 + c;
        if (vec_f32) {
          *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4 *>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = v[j];
        }
      }
      if (a.y_hi && n0 + c < a.yh_cout) {
        __half *dh = a.y_hi + pix[u] * a.ldyh + a.yh_coff + n0 + c, *dl = a.y_lo + pix[u] * a.ldyh + a.yh_coff + n0 + c;
        __align__(16) __half hh[8];
        __align__(16) __half ll[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float sc = v[j] * TC_ACT_SCALE;
          hh[j] = __float2half_rn(sc);
          ll[j] = __float2half_rn(sc - __half2float(hh[j]));
        }
        if (n0 + c + 8 <= a.yh_cout && vec_h) {
          *reinterpret_cast<uint4 *>(dh) = *reinterpret_cast<const uint4 *>(hh);
          *reinterpret_cast<uint4 *>(dl) = *reinterpret_cast<const uint4 *>(ll);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (n0 + c + j < a.yh_cout) { dh[j] = hh[j]; dl[j] = ll[j]; }
        }
      }
      if (a.y_nchw && post) {               // the NCHW pass below reads the finished values
        *reinterpret_cast<float4 *>(tp) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(tp + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
  }
  if (a.y_nchw) {                           // exported feature map: lanes along the 16 contiguous pixels of a tile row
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int64_t hw = (int64_t)a.Ho * a.Wo;
#pragma unroll 1
    for (int idx = te; idx < 128 * BN; idx += 256) {
      const int p = idx & 127, c = idx >> 7;
      const int py = y0 + p / TC_TW, px = x0 + p % TC_TW;
      if (py < a.Ho && px < a.Wo && b < a.B) a.y_nchw[((int64_t)b * a.Cout + n0 + c) * hw + (int64_t)py * a.Wo + px] = tile[p * LDT + c];
    }
  }
}

template <int HALF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P2_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                const __grid_constant__ CUtensorMap tm_w, const TcArgs a) {
  constexpr int BN = 2 * HALF;                       // output channels of the pair's tile
  constexpr int B_PLANE = HALF * 128;                // one weight plane (hi or lo) of this CTA's rows
  constexpr int STAGE_BYTES = 2 * TC_A_BYTES + 2 * B_PLANE;
  constexpr int STAGES = HALF == 128 ? 3 : 4;
  constexpr int SLOT = BN;
  constexpr int COLS = tmem_cols(2 * SLOT);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_full = base + STAGES * STAGE_BYTES;       // leader's are the ones in use
  const uint32_t bar_empty = bar_full + 8 * STAGES;
  const uint32_t bar_accf = bar_empty + 8 * STAGES;            // 2 x 8 B: accumulator slot full (both CTAs)
  const uint32_t bar_acce = bar_accf + 16;                     // 2 x 8 B: slot drained by the 16 epilogue warps of the pair (leader's)
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gen + STAGES * STAGE_BYTES + 16 * STAGES + 32);
  float *s_osc = reinterpret_cast<float *>(gen + STAGES * STAGE_BYTES + 16 * STAGES + 48);   // [2 halves][osc HALF | bias HALF]

#ifdef TC2_TIMING
  __shared__ long long tc2_stamps[10][8];
  const long long t_start = clock64();
#define TC2_STAMP(i) do { if (lane == 0) tc2_stamps[warp][i] = clock64() - t_start; } while (0)
#else
#define TC2_STAMP(i) do {} while (0)
#endif
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t rank = pair_ctarank();
  const int tiles_per_img = a.tiles_x * a.tiles_y;
  const int tile = blockIdx.x;                       // the pair owns tiles 2p, 2p+1; the last one may be past the end
  const int b = tile / tiles_per_img;
  const int tr = tile - b * tiles_per_img;
  const int y0 = (tr / a.tiles_x) * TC_TH, x0 = (tr % a.tiles_x) * TC_TW;
  const int ntile = blockIdx.y;
  const int ntaps = a.kh * a.kw;
  const int nkb = ntaps * a.kchunks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_lo)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_w)) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int j = 0; j < 2; ++j) {
      mbar_init(bar_accf + 8 * j, 1);
      mbar_init(bar_acce + 8 * j, 16);     // one arrive per epilogue warp of either CTA
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();                          // (the cluster barrier below implies it; racecheck only knows this one)
  pair_sync_all();                          // barriers of both CTAs initialised, TMEM of both allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  TC2_STAMP(0);

  if (warp == 0) {
    // ---- producer (both CTAs): own A tile (hi, lo) + this CTA's HALF weight rows (hi, lo), bytes counted by the leader ----
    const uint32_t lead_full = mapa_shared(bar_full, 0);
    const int sub0 = (ntile * BN + (int)rank * HALF) / 64;       // first 64-row weight tile of this CTA
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(bar_empty + 8 * s, ph ^ 1);
      const int tap = kb / a.kchunks, kc = kb - tap * a.kchunks;
      const int ky = tap / a.kw, kx = tap - ky * a.kw;
      const uint32_t sa = base + s * STAGE_BYTES;
      if (elect_one()) {
        if (rank == 0) mbar_expect_tx(bar_full + 8 * s, 2 * STAGE_BYTES);
        const uint32_t fb = lead_full + 8 * s;
        tma2_load_4d(sa, &tm_hi, kc * TC_BK, x0 * a.stride + kx - a.pad, y0 * a.stride + ky - a.pad, b, fb);
        tma2_load_4d(sa + TC_A_BYTES, &tm_lo, kc * TC_BK, x0 * a.stride + kx - a.pad, y0 * a.stride + ky - a.pad, b, fb);
#pragma unroll
        for (int j = 0; j < HALF / 64; ++j) {
          const int row = ((sub0 + j) * nkb + kb) * 128;          // [hi 64 rows | lo 64 rows] of weight tile sub0 + j
          tma2_load_2d(sa + 2 * TC_A_BYTES + j * 8192, &tm_w, 0, row, fb);
          tma2_load_2d(sa + 2 * TC_A_BYTES + B_PLANE + j * 8192, &tm_w, 0, row + 64, fb);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ---- MMA issuer (leader): per k-step  D += A_hi x B_hi,  D += A_hi x B_lo,  D += A_lo x B_hi  over the pair ----
      constexpr uint32_t idesc = umma_idesc(256, BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(bar_full + 8 * s, ph);
        tc_fence_after();
        if (kb == 0) TC2_STAMP(1);
        const int grp = kb / P2_FOLD, first = (kb % P2_FOLD) == 0;
        const uint32_t slot = grp & 1;
        if (first) {                        // both CTAs' epilogues must have drained this slot (two groups ago)
          mbar_wait(bar_acce + 8 * slot, ((grp >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        const uint32_t tacc = tmem_base + slot * SLOT;
        const uint32_t sa = base + s * STAGE_BYTES;
        const uint64_t a_hi = umma_desc(sa), a_lo = umma_desc(sa + TC_A_BYTES);
        const uint64_t b_hi = umma_desc(sa + 2 * TC_A_BYTES), b_lo = umma_desc(sa + 2 * TC_A_BYTES + B_PLANE);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            umma2_f16(tacc, a_hi + adv, b_hi + adv, idesc, (first && k == 0) ? 0u : 1u);
            umma2_f16(tacc, a_hi + adv, b_lo + adv, idesc, 1u);
            umma2_f16(tacc, a_lo + adv, b_hi + adv, idesc, 1u);
          }
          umma2_commit(bar_empty + 8 * s);
          if ((kb % P2_FOLD) == P2_FOLD - 1 || kb == nkb - 1) umma2_commit(bar_accf + 8 * slot);
        }
        __syncwarp();
      }
      TC2_STAMP(2);
    }
  } else {
    // ---- epilogue (both CTAs): warp = (TMEM lane quarter q, column half h); a thread owns HALF channels of one pixel ----
    const int q = warp & 3, h = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int n0 = ntile * BN;
    for (int i = threadIdx.x - 64; i < BN; i += P2_THREADS - 64) {
      const int hh = i / HALF, ii = i - hh * HALF;
      s_osc[hh * 2 * HALF + ii] = a.oscale[n0 + i];
      s_osc[hh * 2 * HALF + HALF + ii] = (a.bias != nullptr && n0 + i < a.Cout) ? a.bias[n0 + i] : 0.f;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const uint32_t lead_acce = mapa_shared(bar_acce, 0);
    float acc[HALF];
#pragma unroll
    for (int j = 0; j < HALF; ++j) acc[j] = 0.f;
    const int ngrp = (nkb + P2_FOLD - 1) / P2_FOLD;
#pragma unroll 1
    for (int grp = 0; grp < ngrp; ++grp) {
      const uint32_t slot = grp & 1;
      mbar_wait(bar_accf + 8 * slot, (grp >> 1) & 1);
      tc_fence_after();
      const uint32_t col0 = tmem_base + ((uint32_t)(q * 32) << 16) + slot * SLOT + (uint32_t)(h * HALF);
#pragma unroll
      for (int c0 = 0; c0 < HALF; c0 += 32) {
        float t[32];
        tmem_ld32(col0 + c0, t);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c0 + j] += t[j];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive_local(bar_acce + 8 * slot);
        else mbar_arrive_cluster(lead_acce + 8 * slot);
      }
    }
    TC2_STAMP(3);
    // every MMA of the pair has completed: the pipeline stages are free and become the staging tile [128][BN + 4] fp32
    float *tile = reinterpret_cast<float *>(gen);
    {
      const float *so = s_osc + h * 2 * HALF, *sb = so + HALF;
      float *trow = tile + row * (BN + 4) + h * HALF;
#pragma unroll
      for (int c0 = 0; c0 < HALF; c0 += 4)
        *reinterpret_cast<float4 *>(trow + c0) =
            make_float4(fmaf(acc[c0], so[c0], sb[c0]), fmaf(acc[c0 + 1], so[c0 + 1], sb[c0 + 1]),
                        fmaf(acc[c0 + 2], so[c0 + 2], sb[c0 + 2]), fmaf(acc[c0 + 3], so[c0 + 3], sb[c0 + 3]));
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    TC2_STAMP(4);
    pair_epilogue_store<BN>(a, tile, (int)threadIdx.x - 64, b, y0, x0, n0);
    TC2_STAMP(5);
  }
  tc_fence_before();
  TC2_STAMP(6);
  pair_sync_all();                          // no CTA of the pair leaves while the other may still signal it or use its TMEM
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(COLS) : "memory");
  }
#ifdef TC2_TIMING
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.y == 0 && (blockIdx.x < 2 || blockIdx.x == gridDim.x - 2))
    printf("tc2 cta %d: setup %lld | mma: first_full %lld issued %lld | epi w2: folds %lld staged %lld stored %lld | w9: folds %lld staged %lld stored %lld | final %lld end %lld\n",
           blockIdx.x, tc2_stamps[0][0], tc2_stamps[1][1], tc2_stamps[1][2], tc2_stamps[2][3], tc2_stamps[2][4], tc2_stamps[2][5],
           tc2_stamps[9][3], tc2_stamps[9][4], tc2_stamps[9][5], tc2_stamps[2][6], clock64() - t_start);
#endif
}

// ----------------------------------------------------------------------------------------------------------------
// Persistent form of the CTA-pair kernel for the shapes with several tiles per SM (the K = 64..256 expand convs of the
// bottleneck stages, the stage-2 convs): there the launch is its epilogue — a 128 x 256 tile means 128 KB of residual reads
// and 128 KB of split-plane writes per CTA, DRAM-bound when every CTA of a wave stores at the same time while the tensor pipe
// idles, followed by the next wave's cold prologue.  Here one pair per TPC walks the (pixel-pair, N tile) items; pair tile
// N = 128 (HALF = 64) so that three 48 KB stages leave room for a DEDICATED [128][132] fp32 staging tile, and the 512 TMEM
// columns are four accumulator slots (one k-block each): while the eight epilogue warps stage and store item i, the
// producer and the MMA warp run up to four k-blocks — whole items at K <= 256 — ahead.
// ----------------------------------------------------------------------------------------------------------------
constexpr int P2P_STAGES = 3, P2P_SLOTS = 4;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P2_THREADS, 1)
conv_tc2p_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                 const __grid_constant__ CUtensorMap tm_w, const TcArgs a, const int n_items, const int ntiles_n) {
  constexpr int HALF = 64, BN = 128;
  constexpr int B_PLANE = HALF * 128;
  constexpr int STAGE_BYTES = 2 * TC_A_BYTES + 2 * B_PLANE;            // 48 KB
  constexpr int TILE_BYTES = 128 * (BN + 4) * 4;                       // staging tile
  constexpr int SLOT = BN;
  constexpr int COLS = 512;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
  float *tile = reinterpret_cast<float *>(gen + P2P_STAGES * STAGE_BYTES);
  constexpr int TAIL = P2P_STAGES * STAGE_BYTES + TILE_BYTES;
  const uint32_t bar_full = base + TAIL;
  const uint32_t bar_empty = bar_full + 8 * P2P_STAGES;
  const uint32_t bar_accf = bar_empty + 8 * P2P_STAGES;               // P2P_SLOTS x 8 B
  const uint32_t bar_acce = bar_accf + 8 * P2P_SLOTS;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gen + TAIL + 16 * P2P_STAGES + 16 * P2P_SLOTS);
  float *s_osc = reinterpret_cast<float *>(gen + TAIL + 16 * P2P_STAGES + 16 * P2P_SLOTS + 16);   // [2 halves][osc 64 | bias 64]

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t rank = pair_ctarank();
  const int tiles_per_img = a.tiles_x * a.tiles_y;
  const int nkb = a.kh * a.kw * a.kchunks;
  const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  // phase totals of the first pair (timing builds only: `make tc2timing`, FRTM_B200_LIB=.../libfrtm_b200_timing.so)
#ifdef TC2_TIMING
  __shared__ long long p2p_t[10][8];
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long t_begin = clock64();
  long long tlast = t_begin;
#define P2P_T(k) do { const long long t__ = clock64(); tacc[k] += t__ - tlast; tlast = t__; } while (0)
#else
#define P2P_T(k) do {} while (0)
#endif

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_lo)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_w)) : "memory");
    for (int s = 0; s < P2P_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int j = 0; j < P2P_SLOTS; ++j) {
      mbar_init(bar_accf + 8 * j, 1);
      mbar_init(bar_acce + 8 * j, 16);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  pair_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---- producer (both CTAs) ----
    const uint32_t lead_full = mapa_shared(bar_full, 0);
    uint32_t it = 0;
    for (int item = cluster; item < n_items; item += n_clusters) {
      const int pairi = item / ntiles_n, ntile = item - pairi * ntiles_n;
      const int tl = 2 * pairi + (int)rank;
      const int b = tl / tiles_per_img, tr = tl - b * tiles_per_img;
      const int y0 = (tr / a.tiles_x) * TC_TH, x0 = (tr % a.tiles_x) * TC_TW;
      const int sub0 = (ntile * BN + (int)rank * HALF) / 64;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % P2P_STAGES;
        mbar_wait(bar_empty + 8 * s, ((it / P2P_STAGES) & 1) ^ 1);
        const int tap = kb / a.kchunks, kc = kb - tap * a.kchunks;
        const int ky = tap / a.kw, kx = tap - ky * a.kw;
        const uint32_t sa = base + s * STAGE_BYTES;
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(bar_full + 8 * s, 2 * STAGE_BYTES);
          const uint32_t fb = lead_full + 8 * s;
          tma2_load_4d(sa, &tm_hi, kc * TC_BK, x0 * a.stride + kx - a.pad, y0 * a.stride + ky - a.pad, b, fb);
          tma2_load_4d(sa + TC_A_BYTES, &tm_lo, kc * TC_BK, x0 * a.stride + kx - a.pad, y0 * a.stride + ky - a.pad, b, fb);
          const int row = (sub0 * nkb + kb) * 128;
          tma2_load_2d(sa + 2 * TC_A_BYTES, &tm_w, 0, row, fb);
          tma2_load_2d(sa + 2 * TC_A_BYTES + B_PLANE, &tm_w, 0, row + 64, fb);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ---- MMA issuer (leader): one k-block per accumulator slot ----
      constexpr uint32_t idesc = umma_idesc(256, BN);
      uint32_t it = 0;
      for (int item = cluster; item < n_items; item += n_clusters) {
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % P2P_STAGES;
          const uint32_t slot = it % P2P_SLOTS;
          P2P_T(0);
          mbar_wait(bar_acce + 8 * slot, ((it / P2P_SLOTS) & 1) ^ 1);      // drained by both CTAs' epilogues
          P2P_T(1);
          mbar_wait(bar_full + 8 * s, (it / P2P_STAGES) & 1);
          P2P_T(2);
          tc_fence_after();
          const uint32_t tacc = tmem_base + slot * SLOT;
          const uint32_t sa = base + s * STAGE_BYTES;
          const uint64_t a_hi = umma_desc(sa), a_lo = umma_desc(sa + TC_A_BYTES);
          const uint64_t b_hi = umma_desc(sa + 2 * TC_A_BYTES), b_lo = umma_desc(sa + 2 * TC_A_BYTES + B_PLANE);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {
              const uint64_t adv = (uint64_t)(k * 32 >> 4);
              umma2_f16(tacc, a_hi + adv, b_hi + adv, idesc, k == 0 ? 0u : 1u);
              umma2_f16(tacc, a_hi + adv, b_lo + adv, idesc, 1u);
              umma2_f16(tacc, a_lo + adv, b_hi + adv, idesc, 1u);
            }
            umma2_commit(bar_empty + 8 * s);
            umma2_commit(bar_accf + 8 * slot);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ---- epilogue (both CTAs): warp = (TMEM lane quarter q, column half h) ----
    const int q = warp & 3, h = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int te = (int)threadIdx.x - 64;
    const uint32_t lead_acce = mapa_shared(bar_acce, 0);
    uint32_t it = 0;
    for (int item = cluster; item < n_items; item += n_clusters) {
      const int pairi = item / ntiles_n, ntile = item - pairi * ntiles_n;
      const int tl = 2 * pairi + (int)rank;
      const int b = tl / tiles_per_img, tr = tl - b * tiles_per_img;
      const int y0 = (tr / a.tiles_x) * TC_TH, x0 = (tr % a.tiles_x) * TC_TW;
      const int n0 = ntile * BN;
      if (te < BN) {                          // this item's scales / biases (read by the tile write below, after barrier A)
        const int hh = te / HALF, ii = te - hh * HALF;
        s_osc[hh * 2 * HALF + ii] = a.oscale[n0 + te];
        s_osc[hh * 2 * HALF + HALF + ii] = (a.bias != nullptr && n0 + te < a.Cout) ? a.bias[n0 + te] : 0.f;
      }
      float acc[HALF];
#pragma unroll
      for (int j = 0; j < HALF; ++j) acc[j] = 0.f;
      P2P_T(0);
#pragma unroll 1
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t slot = it % P2P_SLOTS;
        mbar_wait(bar_accf + 8 * slot, (it / P2P_SLOTS) & 1);
        tc_fence_after();
        const uint32_t col0 = tmem_base + ((uint32_t)(q * 32) << 16) + slot * SLOT + (uint32_t)(h * HALF);
#pragma unroll
        for (int c0 = 0; c0 < HALF; c0 += 32) {
          float t[32];
          tmem_ld32(col0 + c0, t);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[c0 + j] += t[j];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) mbar_arrive_local(bar_acce + 8 * slot);
          else mbar_arrive_cluster(lead_acce + 8 * slot);
        }
      }
      P2P_T(2);
      asm volatile("bar.sync 1, 256;" ::: "memory");      // A: scales staged, the previous item's store loop is over
      P2P_T(3);
      {
        const float *so = s_osc + h * 2 * HALF, *sb = so + HALF;
        float *trow = tile + row * (BN + 4) + h * HALF;
#pragma unroll
        for (int c0 = 0; c0 < HALF; c0 += 4)
          *reinterpret_cast<float4 *>(trow + c0) =
              make_float4(fmaf(acc[c0], so[c0], sb[c0]), fmaf(acc[c0 + 1], so[c0 + 1], sb[c0 + 1]),
                          fmaf(acc[c0 + 2], so[c0 + 2], sb[c0 + 2]), fmaf(acc[c0 + 3], so[c0 + 3], sb[c0 + 3]));
      }
      P2P_T(4);
      asm volatile("bar.sync 1, 256;" ::: "memory");      // B: tile complete
      P2P_T(5);
      pair_epilogue_store<BN>(a, tile, te, b, y0, x0, n0);
      P2P_T(6);
    }
  }
#ifdef TC2_TIMING
  if (lane == 0) for (int i = 0; i < 8; ++i) p2p_t[warp][i] = tacc[i];
#endif
  tc_fence_before();
  pair_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(COLS) : "memory");
  }
#ifdef TC2_TIMING
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x < 2)
    printf("tc2p cta %d total %lld | mma: other %lld wait_drained %lld wait_full %lld | epi w2: other %lld folds+waits %lld barA %lld "
           "tile %lld barB %lld store %lld | epi w9: folds+waits %lld barA %lld tile %lld barB %lld store %lld\n",
           blockIdx.x, clock64() - t_begin, p2p_t[1][0], p2p_t[1][1], p2p_t[1][2], p2p_t[2][0], p2p_t[2][2], p2p_t[2][3],
           p2p_t[2][4], p2p_t[2][5], p2p_t[2][6], p2p_t[9][2], p2p_t[9][3], p2p_t[9][4], p2p_t[9][5], p2p_t[9][6]);
#endif
}

// ----------------------------------------------------------------------------------------------------------------
// 3x3 stride-1 convolution with 64 input channels — the shape of every 3x3 conv of the refinement network and of the
// first backbone stage, i.e. most of the conv time at 480p.  The general kernel above is bound by L2 -> shared-memory
// operand traffic (every CTA re-fetches each tap's shifted A tile and all the weights: 48 KB per 128x64x64 product,
// 10.5 TB/s measured).  Here
//   * the weights of all 9 taps (hi + lo planes, 9 x 2 x BN x 128 B) are fetched ONCE per CTA and stay in shared memory;
//     the CTA is persistent and walks output tiles with stride gridDim.x;
//   * an output tile is 16 rows x 8 pixels, and its input arrives as three column-shifted SLABS (dx = 0,1,2) of
//     18 rows x 8 pixels x 64 channels: with 8-pixel rows a vertical shift dy is a shift by exactly one 1024-byte
//     swizzle atom, so the three vertical taps of a slab are three UMMA descriptors into the same bytes.
// 110 KB of A per tile instead of 295 KB, no weight traffic in the steady state.  Accumulation, fold and epilogue are
// those of the general kernel (packed hi|lo products, TC_FOLD taps per TMEM slot, fp32 register fold).
// ----------------------------------------------------------------------------------------------------------------
constexpr int S3_TH = 16, S3_TW = 8;
constexpr int S3_SLAB_BYTES = (S3_TH + 2) * S3_TW * 128;      // one plane of one slab: 18 rows x 8 px x 64 ch fp16
constexpr int S3_STAGES = 4;                                   // ring of slab PLANES (hi and lo are separate stages)
constexpr int S3_FOLD = 3;                                     // taps per TMEM slot = one slab (K = 192 per product)
constexpr int S3_THREADS = 64 + 2 * 128;                       // producer warp, MMA warp, two epilogue groups of 4 warps
constexpr int S3_NBAR = 2 * S3_STAGES + 8 + 1;

// The epilogue of a tile (three TMEM folds + scale / bias / residual / ReLU + the stores) takes longer than its MMAs, and
// a persistent CTA has no second resident CTA to hide it behind: two epilogue groups take alternate tiles, each with
// its own pair of accumulator slots, so the tensor core runs one tile ahead of the stores.
template <int BN, bool R1, bool TAP>
__global__ void __launch_bounds__(S3_THREADS, 1) conv_tc3_kernel(const __grid_constant__ CUtensorMap tm_hi,
                                                                 const __grid_constant__ CUtensorMap tm_lo, const TcArgs a) {
  constexpr int B_BYTES = BN * 128;
  constexpr int W_BYTES = 9 * 2 * B_BYTES;
  constexpr int STAGE_BYTES = S3_SLAB_BYTES;                    // one plane of one slab
  constexpr int SLOT = 2 * BN;
  constexpr int COLS = tmem_cols(4 * SLOT);                      // 2 groups x 2 slots
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t wsm = base;                                     // weights [tap][hi|lo][BN x 128 B]
  const uint32_t slabs = base + W_BYTES;                         // S3_STAGES x [hi|lo] slab
  constexpr int TAIL = W_BYTES + S3_STAGES * STAGE_BYTES;
  const uint32_t bar_full = base + TAIL;                         // S3_STAGES x 8 B
  const uint32_t bar_empty = bar_full + 8 * S3_STAGES;
  const uint32_t bar_accf = bar_empty + 8 * S3_STAGES;           // [group][slot]: accumulator slot full
  const uint32_t bar_acce = bar_accf + 32;                       // [group][slot]: accumulator slot drained
  const uint32_t bar_w = bar_acce + 32;                          // weights landed
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gen + TAIL + 8 * S3_NBAR);
  float *s_osc = reinterpret_cast<float *>(gen + TAIL + 8 * S3_NBAR + 16);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int tiles_per_img = a.tiles_x * a.tiles_y;
  const int ntiles = a.B * tiles_per_img;
  constexpr int NGRP = 9 / S3_FOLD;                              // folds per tile

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_lo)) : "memory");
    for (int s = 0; s < S3_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int j = 0; j < 4; ++j) {
      mbar_init(bar_accf + 8 * j, 1);
      mbar_init(bar_acce + 8 * j, 4);      // one arrive per epilogue warp of the group
    }
    mbar_init(bar_w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---- producer: the weights once, then three slabs per tile ----
    if (elect_one()) {
      mbar_expect_tx(bar_w, W_BYTES);
#pragma unroll 1
      for (int t = 0; t < 9; ++t)
        bulk_load(wsm + t * 2 * B_BYTES, reinterpret_cast<const uint8_t *>(a.wt) + (size_t)t * 2 * B_BYTES, 2 * B_BYTES, bar_w);
    }
    __syncwarp();
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int b = tile / tiles_per_img, tr = tile - b * tiles_per_img;
      const int y0 = (tr / a.tiles_x) * S3_TH, x0 = (tr % a.tiles_x) * S3_TW;
      // the hi and the lo plane of a slab are separate ring stages: a stage is refilled as soon as the products that
      // read ITS plane have completed, which keeps twice as many refills in flight for the same shared memory
      for (int st = 0; st < 6; ++st, ++it) {
        const int dx = st >> 1;
        const int s = it % S3_STAGES;
        mbar_wait(bar_empty + 8 * s, ((it / S3_STAGES) & 1) ^ 1);
        const uint32_t sa = slabs + s * STAGE_BYTES;
        if (elect_one()) {
          mbar_expect_tx(bar_full + 8 * s, STAGE_BYTES);
          tma_load_4d(sa, (st & 1) ? &tm_lo : &tm_hi, 0, x0 + dx - 1, y0 - 1, b, bar_full + 8 * s);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: per slab (dx) the three vertical taps, each 4 k-steps of the packed products, one TMEM slot ----
    constexpr uint32_t idesc = umma_idesc(128, BN), idesc2 = umma_idesc(128, 2 * BN);
    mbar_wait(bar_w, 0);
    uint32_t it = 0, nt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++nt) {
      const uint32_t eg = nt & 1;                                // epilogue group of this tile
      const uint32_t gc = (nt >> 1) * NGRP;                      // folds this group has seen before this tile
      for (int dx = 0; dx < 3; ++dx) {
        const uint32_t grp = gc + dx, slot = grp & 1;
        mbar_wait(bar_acce + 8 * (2 * eg + slot), ((grp >> 1) & 1) ^ 1);   // the group has drained this slot
        tc_fence_after();
        const uint32_t tacc = tmem_base + (2 * eg + slot) * SLOT;
        // hi plane: A_hi x [B_hi | B_lo] (N = 2*BN) for the three vertical taps; then lo plane: A_lo x B_hi (N = BN)
#pragma unroll
        for (int pl = 0; pl < 2; ++pl, ++it) {
          const int s = it % S3_STAGES;
          mbar_wait(bar_full + 8 * s, (it / S3_STAGES) & 1);
          const uint64_t dA = umma_desc(slabs + s * STAGE_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const uint64_t a_p = dA + (uint64_t)(dy * 1024 >> 4);
              const uint64_t b_hl = umma_desc(wsm + (dy * 3 + dx) * 2 * B_BYTES);
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k) {
                const uint64_t adv = (uint64_t)(k * 32 >> 4);
                if (pl == 0) umma_f16(tacc, a_p + adv, b_hl + adv, idesc2, (dy == 0 && k == 0) ? 0u : 1u);
                else umma_f16(tacc, a_p + adv, b_hl + adv, idesc, 1u);
              }
            }
            umma_commit(bar_empty + 8 * s);
            if (pl == 1) umma_commit(bar_accf + 8 * (2 * eg + slot));
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ---- epilogue groups: warps 2-5 take the even tiles of this CTA, warps 6-9 the odd ones ----
    const int eg = (warp - 2) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    tc_epilogue_stage<BN, R1, TAP>(a, 0, s_osc);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    uint32_t nt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++nt) {
      if ((int)(nt & 1) != eg) continue;
      const uint32_t gc = (nt >> 1) * NGRP;
      const int b = tile / tiles_per_img, tr = tile - b * tiles_per_img;
      const int y0 = (tr / a.tiles_x) * S3_TH, x0 = (tr % a.tiles_x) * S3_TW;
      float acc[BN];
#pragma unroll
      for (int j = 0; j < BN; ++j) acc[j] = 0.f;
#pragma unroll 1
      for (int g = 0; g < NGRP; ++g) {
        const uint32_t grp = gc + g, slot = grp & 1;
        mbar_wait(bar_accf + 8 * (2 * eg + slot), (grp >> 1) & 1);
        tc_fence_after();
        const uint32_t col0 = tmem_base + ((uint32_t)(q * 32) << 16) + (2 * eg + slot) * SLOT;
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
          if (c0 + 32 <= BN) {
            float t[32], t2[32];
            tmem_ld32(col0 + c0, t);
            tmem_ld32(col0 + BN + c0, t2);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[c0 + j] += t[j] + t2[j];
          } else {
            float t[16], t2[16];
            tmem_ld16(col0 + c0, t);
            tmem_ld16(col0 + BN + c0, t2);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[c0 + j] += t[j] + t2[j];
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_acce + 8 * (2 * eg + slot)) : "memory");
      }
      tc_epilogue_store<BN, R1, TAP>(a, acc, b, y0 + row / S3_TW, x0 + row % S3_TW, 0, s_osc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(COLS) : "memory");
  }
}

// ----------------------------------------------------------------------------------------------------------------
// 1x1 stride-1 convolution with up to 256 input channels and one N tile (the RRB / TSE 1x1 convs of the refinement
// network, the stem over its patches).  These are pure streaming kernels — 64 MACs per loaded element — and the
// general kernel spends them one CTA per 128 pixels, each paying barrier / TMEM / tensor-map setup, one exposed load and
// an unoverlapped epilogue (2.4 TB/s measured).  Same remedy as the slab kernel: persistent CTAs, weights resident,
// a 4-stage ring of 16x8-pixel tiles x 64 channels, one TMEM slot per tile (K <= 256), two epilogue groups.
// ----------------------------------------------------------------------------------------------------------------
constexpr int S1_STAGES = 4;
constexpr int S1_TILE_BYTES = 128 * 128;                       // one plane of one k-chunk of a tile
constexpr int S1_NBAR = 2 * S1_STAGES + 8 + 1;

template <int BN, bool R1, bool TAP>
__global__ void __launch_bounds__(S3_THREADS, 1) conv_tc1_kernel(const __grid_constant__ CUtensorMap tm_hi,
                                                                 const __grid_constant__ CUtensorMap tm_lo, const TcArgs a) {
  constexpr int B_BYTES = BN * 128;
  constexpr int STAGE_BYTES = 2 * S1_TILE_BYTES;
  constexpr int SLOT = 2 * BN;
  constexpr int COLS = tmem_cols(4 * SLOT);                      // 2 groups x 2 slots
  const int KC = a.kchunks;
  const int W_BYTES = KC * 2 * B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t ring = base;                                    // S1_STAGES x [hi|lo] tile chunk
  const uint32_t wsm = base + S1_STAGES * STAGE_BYTES;           // weights [kc][hi|lo][BN x 128 B] (up to 4 chunks)
  constexpr int TAIL = S1_STAGES * STAGE_BYTES + 4 * 2 * B_BYTES;
  const uint32_t bar_full = base + TAIL;
  const uint32_t bar_empty = bar_full + 8 * S1_STAGES;
  const uint32_t bar_accf = bar_empty + 8 * S1_STAGES;           // [group][slot]
  const uint32_t bar_acce = bar_accf + 32;
  const uint32_t bar_w = bar_acce + 32;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gen + TAIL + 8 * S1_NBAR);
  float *s_osc = reinterpret_cast<float *>(gen + TAIL + 8 * S1_NBAR + 16);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int tiles_per_img = a.tiles_x * a.tiles_y;
  const int ntiles = a.B * tiles_per_img;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_lo)) : "memory");
    for (int s = 0; s < S1_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int j = 0; j < 4; ++j) {
      mbar_init(bar_accf + 8 * j, 1);
      mbar_init(bar_acce + 8 * j, 4);
    }
    mbar_init(bar_w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bar_w, W_BYTES);
      for (int kc = 0; kc < KC; ++kc)
        bulk_load(wsm + kc * 2 * B_BYTES, reinterpret_cast<const uint8_t *>(a.wt) + (size_t)kc * 2 * B_BYTES, 2 * B_BYTES, bar_w);
    }
    __syncwarp();
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int b = tile / tiles_per_img, tr = tile - b * tiles_per_img;
      const int y0 = (tr / a.tiles_x) * S3_TH, x0 = (tr % a.tiles_x) * S3_TW;
      for (int kc = 0; kc < KC; ++kc, ++it) {
        const int s = it % S1_STAGES;
        mbar_wait(bar_empty + 8 * s, ((it / S1_STAGES) & 1) ^ 1);
        const uint32_t sa = ring + s * STAGE_BYTES;
        if (elect_one()) {
          mbar_expect_tx(bar_full + 8 * s, STAGE_BYTES);
          tma_load_4d(sa, &tm_hi, kc * TC_BK, x0, y0, b, bar_full + 8 * s);
          tma_load_4d(sa + S1_TILE_BYTES, &tm_lo, kc * TC_BK, x0, y0, b, bar_full + 8 * s);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc(128, BN), idesc2 = umma_idesc(128, 2 * BN);
    mbar_wait(bar_w, 0);
    uint32_t it = 0, nt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++nt) {
      const uint32_t eg = nt & 1, grp = nt >> 1, slot = grp & 1;
      mbar_wait(bar_acce + 8 * (2 * eg + slot), ((grp >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + (2 * eg + slot) * SLOT;
      for (int kc = 0; kc < KC; ++kc, ++it) {
        const int s = it % S1_STAGES;
        mbar_wait(bar_full + 8 * s, (it / S1_STAGES) & 1);
        const uint64_t a_hi = umma_desc(ring + s * STAGE_BYTES), a_lo = a_hi + (uint64_t)(S1_TILE_BYTES >> 4);
        const uint64_t b_hl = umma_desc(wsm + kc * 2 * B_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            umma_f16(tacc, a_hi + adv, b_hl + adv, idesc2, (kc == 0 && k == 0) ? 0u : 1u);
            umma_f16(tacc, a_lo + adv, b_hl + adv, idesc, 1u);
          }
          umma_commit(bar_empty + 8 * s);
          if (kc == KC - 1) umma_commit(bar_accf + 8 * (2 * eg + slot));
        }
        __syncwarp();
      }
    }
  } else {
    const int eg = (warp - 2) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    tc_epilogue_stage<BN, R1, TAP>(a, 0, s_osc);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    uint32_t nt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++nt) {
      if ((int)(nt & 1) != eg) continue;
      const uint32_t grp = nt >> 1, slot = grp & 1;
      const int b = tile / tiles_per_img, tr = tile - b * tiles_per_img;
      const int y0 = (tr / a.tiles_x) * S3_TH, x0 = (tr % a.tiles_x) * S3_TW;
      float acc[BN];
      mbar_wait(bar_accf + 8 * (2 * eg + slot), (grp >> 1) & 1);
      tc_fence_after();
      const uint32_t col0 = tmem_base + ((uint32_t)(q * 32) << 16) + (2 * eg + slot) * SLOT;
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float t[32], t2[32];
        tmem_ld32(col0 + c0, t);
        tmem_ld32(col0 + BN + c0, t2);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c0 + j] = t[j] + t2[j];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_acce + 8 * (2 * eg + slot)) : "memory");
      tc_epilogue_store<BN, R1, TAP>(a, acc, b, y0 + row / S3_TW, x0 + row % S3_TW, 0, s_osc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(COLS) : "memory");
  }
}

// ----------------------------------------------------------------------------------------------------------------
// The 7x7 / stride-2 stem (torchvision resnet.py conv1 + bn1 + relu, model/feature_extractor.py:42-46) straight from the
// uint8 image.  The stem is a 1x1 tensor-core conv over im2col patches (K = ky*24 + c*8 + kx = 192, frtm_stem_patches_u8);
// materialising the patches costs 570 MB of writes and the same again of reads per 8 frames of 480x854 for a 3 MB input.
// Here eight builder warps write the patches of a 16x8-pixel tile directly into the shared-memory image the MMA reads:
// one thread = one 16-byte chunk (the 7 kx taps of one (pixel, ky, c) + a zero) of the hi and of the lo plane, placed at
// its 128-byte-swizzled position (row = pixel, chunk j of k-block kc at j ^ (pixel % 8)); a fence.proxy.async publishes
// the generic-proxy writes to the tensor core.  Everything else is the streaming 1x1 kernel: weights resident, a 4-stage
// ring of k-blocks, one TMEM slot per tile, fp32 fold + bias (folded BatchNorm) + ReLU in the epilogue.  The patch values
// are computed exactly as stem_patches_kernel computes them, so the result is bit-identical to the two-kernel path.
// warp 0: weights; warp 1: MMA; warps 2-5: epilogue; warps 6-13: builders.
// ----------------------------------------------------------------------------------------------------------------
constexpr int ST_THREADS = 448, ST_BUILD0 = 192, ST_NBUILD = 256;
constexpr int ST_KC = 3;                                         // k-blocks of 64: K = 192
constexpr int ST_RROWS = 2 * S3_TH + 5, ST_RCOLS = 24;           // image region behind a tile: 37 rows x (2*8+5 = 21 -> 24) columns

__global__ void __launch_bounds__(ST_THREADS, 1) conv_stem_kernel(const uint8_t *__restrict__ img, const TcArgs a, float s0, float s1,
                                                                  float s2, float b0, float b1, float b2) {
  constexpr int BN = 64;
  constexpr int B_BYTES = BN * 128;
  constexpr int STAGE_BYTES = 2 * S1_TILE_BYTES;
  constexpr int SLOT = 2 * BN;
  constexpr int COLS = tmem_cols(2 * SLOT);
  constexpr int W_BYTES = ST_KC * 2 * B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t ring = base;
  const uint32_t wsm = base + S1_STAGES * STAGE_BYTES;
  constexpr int TAIL = S1_STAGES * STAGE_BYTES + W_BYTES;
  const uint32_t bar_full = base + TAIL;                         // S1_STAGES x 8 B: one arrive per builder warp
  const uint32_t bar_empty = bar_full + 8 * S1_STAGES;
  const uint32_t bar_accf = bar_empty + 8 * S1_STAGES;           // 2 x 8 B
  const uint32_t bar_acce = bar_accf + 16;
  const uint32_t bar_w = bar_acce + 16;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gen + TAIL + 16 * S1_STAGES + 40);
  float *s_osc = reinterpret_cast<float *>(gen + TAIL + 16 * S1_STAGES + 64);          // osc[64] | bias[64] (+ unused tap / r1 areas)
  uint8_t *region = gen + TAIL + 16 * S1_STAGES + 64 + 84 * BN;                          // 2 x [hi|lo][3][ST_RROWS][ST_RCOLS] halves

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int tiles_per_img = a.tiles_x * a.tiles_y;
  const int ntiles = a.B * tiles_per_img;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < S1_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, ST_NBUILD / 32);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int j = 0; j < 2; ++j) {
      mbar_init(bar_accf + 8 * j, 1);
      mbar_init(bar_acce + 8 * j, 4);
    }
    mbar_init(bar_w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bar_w, W_BYTES);
      for (int kc = 0; kc < ST_KC; ++kc)
        bulk_load(wsm + kc * 2 * B_BYTES, reinterpret_cast<const uint8_t *>(a.wt) + (size_t)kc * 2 * B_BYTES, 2 * B_BYTES, bar_w);
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc(128, BN), idesc2 = umma_idesc(128, 2 * BN);
    mbar_wait(bar_w, 0);
    uint32_t it = 0, nt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++nt) {
      const uint32_t slot = nt & 1;
      mbar_wait(bar_acce + 8 * slot, ((nt >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + slot * SLOT;
      for (int kc = 0; kc < ST_KC; ++kc, ++it) {
        const int s = it % S1_STAGES;
        mbar_wait(bar_full + 8 * s, (it / S1_STAGES) & 1);
        tc_fence_after();
        const uint64_t a_hi = umma_desc(ring + s * STAGE_BYTES), a_lo = a_hi + (uint64_t)(S1_TILE_BYTES >> 4);
        const uint64_t b_hl = umma_desc(wsm + kc * 2 * B_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            umma_f16(tacc, a_hi + adv, b_hl + adv, idesc2, (kc == 0 && k == 0) ? 0u : 1u);
            umma_f16(tacc, a_lo + adv, b_hl + adv, idesc, 1u);
          }
          umma_commit(bar_empty + 8 * s);
          if (kc == ST_KC - 1) umma_commit(bar_accf + 8 * slot);
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ---- epilogue: tiles alternate between the two TMEM slots ----
    const int q = warp & 3;
    const int row = q * 32 + lane;
    tc_epilogue_stage<BN, false, false>(a, 0, s_osc);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    uint32_t nt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++nt) {
      const uint32_t slot = nt & 1;
      const int b = tile / tiles_per_img, tr = tile - b * tiles_per_img;
      const int y0 = (tr / a.tiles_x) * S3_TH, x0 = (tr % a.tiles_x) * S3_TW;
      float acc[BN];
      mbar_wait(bar_accf + 8 * slot, (nt >> 1) & 1);
      tc_fence_after();
      const uint32_t col0 = tmem_base + ((uint32_t)(q * 32) << 16) + slot * SLOT;
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float t[32], t2[32];
        tmem_ld32(col0 + c0, t);
        tmem_ld32(col0 + BN + c0, t2);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c0 + j] = t[j] + t2[j];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_acce + 8 * slot) : "memory");
      tc_epilogue_store<BN, false, false>(a, acc, b, y0 + row / S3_TW, x0 + row % S3_TW, 0, s_osc);
    }
  } else {
    // ---- builders: the tile's image region -> shared memory, then its patches k-block by k-block into the ring.  The
    //      region of the NEXT tile is fetched into registers (all loads in flight) before the current tile's patches are
    //      built and stored to the other region buffer afterwards, so its global-memory latency hides behind the build ----
    const int tb = (int)threadIdx.x - ST_BUILD0;
    constexpr int RBYTES = 3 * ST_RROWS * ST_RCOLS, RPT = (RBYTES + ST_NBUILD - 1) / ST_NBUILD;
    uint32_t nxt[RPT];      // one register per byte (a packed uint8_t array would make every load wait for its byte insert); 256 = outside
    auto fetch = [&](int tile) {
      const int b = tile / tiles_per_img, tr = tile - b * tiles_per_img;
      const int ry0 = 2 * (tr / a.tiles_x) * S3_TH - 3, rx0 = 2 * (tr % a.tiles_x) * S3_TW - 3;
#pragma unroll
      for (int k = 0; k < RPT; ++k) {
        const int i = tb + k * ST_NBUILD;
        const int c = i / (ST_RROWS * ST_RCOLS), rr = (i / ST_RCOLS) % ST_RROWS, cc = i % ST_RCOLS;
        const int iy = ry0 + rr, ix = rx0 + cc;
        nxt[k] = 256u;
        if (i < RBYTES && tile < ntiles && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W)
          nxt[k] = img[(((int64_t)b * 3 + c) * a.H + iy) * (int64_t)a.W + ix];
      }
    };
    // the region as split fp16 planes of 16 * normalised value (exactly stem_patches_kernel's arithmetic: mul then add,
    // separately rounded), zero outside the image: every input pixel is converted once per tile instead of once per
    // patch it appears in (~12 x), and a patch chunk becomes a plain 14-byte gather per plane
    auto stash = [&](__half *dst) {
#pragma unroll
      for (int k = 0; k < RPT; ++k) {
        const int i = tb + k * ST_NBUILD;
        if (i < RBYTES) {
          const int c = i / (ST_RROWS * ST_RCOLS);
          const float sc = c == 0 ? s0 : (c == 1 ? s1 : s2), bc = c == 0 ? b0 : (c == 1 ? b1 : b2);
          const float v = nxt[k] < 256u ? __fadd_rn(__fmul_rn(sc, (float)nxt[k]), bc) * 16.f : 0.f;
          const __half h = __float2half_rn(v);
          dst[i] = h;
          dst[RBYTES + i] = __float2half_rn(v - __half2float(h));
        }
      }
    };
    __half *region0 = reinterpret_cast<__half *>(region);        // two buffers of [hi | lo][3][rows][cols] halves, tiles alternate
    fetch(blockIdx.x);
    stash(region0);
    asm volatile("bar.sync 2, 256;" ::: "memory");
#ifdef TC2_TIMING
    long long tacc[6] = {0, 0, 0, 0, 0, 0}, tlast = clock64();
#define ST_T(k) do { const long long t__ = clock64(); tacc[k] += t__ - tlast; tlast = t__; } while (0)
#else
#define ST_T(k) do {} while (0)
#endif
    uint32_t it = 0, nt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++nt) {
      const int b = tile / tiles_per_img, tr = tile - b * tiles_per_img;
      const int y0 = (tr / a.tiles_x) * S3_TH, x0 = (tr % a.tiles_x) * S3_TW;
      (void)b;
      const __half *reg_h = region0 + (nt & 1) * 2 * RBYTES, *reg_l = reg_h + RBYTES;
      ST_T(0);
      fetch(tile + (int)gridDim.x);
      ST_T(1);
      for (int kc = 0; kc < ST_KC; ++kc, ++it) {
        const int s = it % S1_STAGES;
        mbar_wait(bar_empty + 8 * s, ((it / S1_STAGES) & 1) ^ 1);
        ST_T(2);
        uint8_t *st_hi = gen + s * STAGE_BYTES, *st_lo = st_hi + S1_TILE_BYTES;
#pragma unroll 2
        for (int i = tb; i < 128 * 8; i += ST_NBUILD) {
          const int r = i >> 3, jc = i & 7, j = kc * 8 + jc;
          const int ty = r >> 3, tx = r & 7;
          const int ky = min(j / 3, 6), c = j - (j / 3) * 3;
          // seven consecutive halves of a region row (even start: 4-byte aligned), the eighth is the zero pad
          const int ro = (c * ST_RROWS + 2 * ty + ky) * ST_RCOLS + 2 * tx;
          const uint32_t *gh = reinterpret_cast<const uint32_t *>(reg_h + ro), *gl = reinterpret_cast<const uint32_t *>(reg_l + ro);
          uint32_t ph[4], pl[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) { ph[e] = gh[e]; pl[e] = gl[e]; }
          ph[3] &= 0xffffu; pl[3] &= 0xffffu;
          if (j >= 21) {
#pragma unroll
            for (int e = 0; e < 4; ++e) { ph[e] = 0u; pl[e] = 0u; }
          }
          const int off = r * 128 + ((jc ^ (r & 7)) << 4);
          *reinterpret_cast<uint4 *>(st_hi + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
          *reinterpret_cast<uint4 *>(st_lo + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        }
        ST_T(3);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_full + 8 * s) : "memory");
        ST_T(4);
      }
      stash(region0 + ((nt & 1) ^ 1) * 2 * RBYTES);
      asm volatile("bar.sync 2, 256;" ::: "memory");              // next region complete, this one no longer read
      ST_T(5);
    }
#ifdef TC2_TIMING
    if (tb == 0 && blockIdx.x == 0)
      printf("stem cta 0 builder: tiles %u | other %lld fetch-issue %lld wait-empty %lld build %lld fence+arrive %lld stash+bar %lld\n", nt,
             tacc[0], tacc[1], tacc[2], tacc[3], tacc[4], tacc[5]);
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(COLS) : "memory");
  }
}

// fp32 NHWC (ldx) -> two fp16 planes hi/lo of x * 2^4 (channel stride ldh)
__global__ void split_kernel(const float *__restrict__ x, int64_t npix, int C, int ldx, __half *__restrict__ hi,
                             __half *__restrict__ lo, int ldh) {
  const int C4 = C / 4;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * C4) return;
  const int64_t p = i / C4;
  const int c = (int)(i - p * C4) * 4;
  const float4 v = *reinterpret_cast<const float4 *>(x + p * ldx + c);
  const float s[4] = {v.x * TC_ACT_SCALE, v.y * TC_ACT_SCALE, v.z * TC_ACT_SCALE, v.w * TC_ACT_SCALE};
  __half h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = __float2half_rn(s[j]);
    l[j] = __float2half_rn(s[j] - __half2float(h[j]));
  }
  *reinterpret_cast<uint2 *>(hi + p * ldh + c) = *reinterpret_cast<uint2 *>(h);
  *reinterpret_cast<uint2 *>(lo + p * ldh + c) = *reinterpret_cast<uint2 *>(l);
}

// Finishes a conv whose input is cat(64 channels, 1 score channel) (TSE.transform, model/seg_network.py:15,19-20): the
// tensor-core kernel has produced the 64-channel part in y; this adds the score channel's 3x3 contribution, the bias and
// the ReLU, and emits what the next layer needs: split planes for outputs [0,64) and/or fp32, and output channel 64 (when
// Cout == 65) as a separate fp32 map — the next conv's score channel.
__global__ void __launch_bounds__(256)
rank1_finish_kernel(const float *__restrict__ yin, int ldin, int n_obj, float *__restrict__ yout, int ldout,
                    const float *__restrict__ s, const float *__restrict__ wx, const float *__restrict__ bias, int B, int H,
                    int W, int Cout, int relu, __half *__restrict__ yh, __half *__restrict__ yl, int ldh,
                    float *__restrict__ extra) {
  // thread = (pixel, group of 4 output channels); the 65th output channel (if any) is a 17th, scalar group
  extern __shared__ float wsm[];   // wx [9][Cout] then bias [Cout]
  for (int i = threadIdx.x; i < 9 * Cout; i += blockDim.x) wsm[i] = wx[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) wsm[9 * Cout + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int ngrp = (Cout + 3) >> 2;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * H * W * ngrp;
  if (idx >= total) return;
  const int grp = (int)(idx % ngrp);
  const int64_t pix = idx / ngrp;
  const int64_t hw = (int64_t)W * H;
  const int px = (int)(pix % W);
  const int py = (int)((pix / W) % H);
  const int64_t img = pix / hw;
  const float *sb = s + img * hw;
  float sv[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
    sv[t] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? sb[yy * W + xx] : 0.f;
  }
  // the 64-channel part may be shared by the n_obj objects of a frame (it depends on backbone features only)
  const float *yi = yin + ((img / n_obj) * hw + (pix - img * hw)) * ldin;
  const int n0 = grp * 4;
  const int nv = min(4, Cout - n0);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (nv == 4 && (ldin & 3) == 0) {
    const float4 v = *reinterpret_cast<const float4 *>(yi + n0);
    acc[0] = v.x; acc[1] = v.y; acc[2] = v.z; acc[3] = v.w;
  } else {
    for (int j = 0; j < nv; ++j) acc[j] = yi[n0 + j];
  }
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nv) acc[j] = fmaf(wsm[t * Cout + n0 + j], sv[t], acc[j]);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j < nv) {
      acc[j] += wsm[9 * Cout + n0 + j];
      if (relu) acc[j] = fmaxf(acc[j], 0.f);
    }
  }
  if (yout) {
    float *yo = yout + pix * ldout + n0;
    if (nv == 4 && (ldout & 3) == 0) *reinterpret_cast<float4 *>(yo) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else for (int j = 0; j < nv; ++j) yo[j] = acc[j];
  }
  if (n0 < 64) {
    if (yh) {
      __align__(8) __half hh[4];
      __align__(8) __half ll[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float sc = acc[j] * TC_ACT_SCALE;
        hh[j] = __float2half_rn(sc);
        ll[j] = __float2half_rn(sc - __half2float(hh[j]));
      }
      *reinterpret_cast<uint2 *>(yh + pix * ldh + n0) = *reinterpret_cast<const uint2 *>(hh);
      *reinterpret_cast<uint2 *>(yl + pix * ldh + n0) = *reinterpret_cast<const uint2 *>(ll);
    }
  } else if (extra) {
    extra[pix] = acc[0];
  }
}

// Device-side packer for 1x1 weights that change at run time (the target models' projection matrices): W (Cout,Cin) fp32
// -> the tile layout of TcArgs::wt with per-row power-of-two scaling, and oscale.  One block per output row.
__global__ void __launch_bounds__(256) pack_tc_1x1_kernel(const float *__restrict__ W, int Cout, int Cin, int BN,
                                                          __half *__restrict__ wt, float *__restrict__ oscale) {
  __shared__ float red[32];
  const int n = blockIdx.x;                       // 0 .. cout_pad-1
  const int nkc = Cin / TC_BK;
  const int ntile = n / BN, r = n - ntile * BN;
  float amax = 0.f;
  if (n < Cout)
    for (int k = threadIdx.x; k < Cin; k += blockDim.x) amax = fmaxf(amax, fabsf(W[(int64_t)n * Cin + k]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
  __syncthreads();
  amax = 0.f;
  for (int i = 0; i < (blockDim.x + 31) / 32; ++i) amax = fmaxf(amax, red[i]);
  float scale = 1.f;
  if (amax > 0.f) {
    int e;
    frexpf(amax, &e);                              // amax = m * 2^e, m in [0.5, 1)  ->  floor(log2(amax)) = e - 1
    scale = ldexpf(1.f, 10 - e);                   // max |w| * scale in [2^9, 2^10)
  }
  if (threadIdx.x == 0) oscale[n] = (n < Cout) ? 1.f / (TC_ACT_SCALE * scale) : 1.f;
  for (int k = threadIdx.x; k < Cin; k += blockDim.x) {
    const float v = (n < Cout) ? W[(int64_t)n * Cin + k] * scale : 0.f;
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    const int kc = k / TC_BK, kk = k - kc * TC_BK;
    const int chunk = kk >> 3, e8 = kk & 7;
    const int64_t tile = ((int64_t)ntile * nkc + kc) * 2;
    const int64_t off = (int64_t)r * 64 + ((chunk ^ (r & 7)) << 3) + e8;
    wt[(tile + 0) * BN * 64 + off] = h;
    wt[(tile + 1) * BN * 64 + off] = l;
  }
}

// ----------------------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int make_act_map(CUtensorMap *tm, const __half *ptr, int B, int H, int W, int C, int ld, int stride, int box_w = 0,
                        int box_h = 0) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("conv_tc: cuTensorMapEncodeTiled entry point not available"); return FRTM_ELAUNCH; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  // with a traversal stride s the box spans TW*s x TH*s input pixels and TMA keeps every s-th one (ceil(box/s) elements)
  cuuint32_t box[4] = {TC_BK, (cuuint32_t)(box_w ? box_w : TC_TW * stride), (cuuint32_t)(box_h ? box_h : TC_TH * stride), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half *>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return FRTM_ELAUNCH; }
  return FRTM_OK;
}

// the pre-swizzled weight image as rows of 128 B; a box is one plane (hi or lo) of one 64-row weight tile of one k-block
static int make_weight_map(CUtensorMap *tm, const __half *wt, int64_t rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("conv_tc: cuTensorMapEncodeTiled entry point not available"); return FRTM_ELAUNCH; }
  cuuint64_t dims[2] = {64, (cuuint64_t)rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, 64};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half *>(wt), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled (weights) failed (%d)", (int)r); return FRTM_ELAUNCH; }
  return FRTM_OK;
}

template <int HALF>
static int launch_tc2(const CUtensorMap &mh, const CUtensorMap &ml, const CUtensorMap &mw, const TcArgs &a, cudaStream_t st) {
  constexpr int stages = HALF == 128 ? 3 : 4;
  constexpr int smem = stages * (2 * TC_A_BYTES + 2 * HALF * 128) + 16 * stages + 48 + 16 * HALF + 1024;
  static_assert(smem <= 227 * 1024, "conv_tc2: shared memory budget");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("conv_tc2: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return FRTM_ELAUNCH; }
    configured = true;
  }
  const int ntiles = a.B * a.tiles_x * a.tiles_y;
  dim3 grid((unsigned)(2 * cdiv(ntiles, 2)), (unsigned)(a.Cout / (2 * HALF)));
  conv_tc2_kernel<HALF><<<grid, P2_THREADS, smem, st>>>(mh, ml, mw, a);
  FRTM_CHECK_LAUNCH("conv_tc2");
  return FRTM_OK;
}

static int launch_tc2p(const CUtensorMap &mh, const CUtensorMap &ml, const CUtensorMap &mw, const TcArgs &a, int num_sms, cudaStream_t st) {
  constexpr int smem = P2P_STAGES * (2 * TC_A_BYTES + 2 * 64 * 128) + 128 * 132 * 4 + 16 * P2P_STAGES + 16 * P2P_SLOTS + 16 + 1024 + 1024;
  static_assert(smem <= 227 * 1024, "conv_tc2p: shared memory budget");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc2p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("conv_tc2p: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return FRTM_ELAUNCH; }
    configured = true;
  }
  const int ntiles = a.B * a.tiles_x * a.tiles_y, ntiles_n = a.Cout / 128;
  const int n_items = cdiv(ntiles, 2) * ntiles_n;
  const int clusters = n_items < num_sms / 2 ? n_items : num_sms / 2;
  conv_tc2p_kernel<<<2 * clusters, P2_THREADS, smem, st>>>(mh, ml, mw, a, n_items, ntiles_n);
  FRTM_CHECK_LAUNCH("conv_tc2p");
  return FRTM_OK;
}

template <int BN, int STAGES, bool R1, bool TAP>
static int launch_tc(const CUtensorMap &mh, const CUtensorMap &ml, const TcArgs &a, dim3 grid, cudaStream_t st) {
  constexpr int smem = STAGES * (2 * TC_A_BYTES + 2 * BN * 128) + 16 * STAGES + 48 + 84 * BN + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES, R1, TAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return FRTM_ELAUNCH; }
    configured = true;
  }
  conv_tc_kernel<BN, STAGES, R1, TAP><<<grid, 192, smem, st>>>(mh, ml, a);
  FRTM_CHECK_LAUNCH("conv_tc");
  return FRTM_OK;
}

template <int BN, bool R1, bool TAP>
static int launch_tc3(const CUtensorMap &mh, const CUtensorMap &ml, const TcArgs &a, cudaStream_t st) {
  constexpr int smem = 9 * 2 * BN * 128 + S3_STAGES * S3_SLAB_BYTES + 8 * S3_NBAR + 16 + 84 * BN + 1024;
  static_assert(smem <= 227 * 1024, "conv_tc3: shared memory budget");
  static bool configured = false;
  static int num_sms = 0;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc3_kernel<BN, R1, TAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("conv_tc3: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return FRTM_ELAUNCH; }
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
    configured = true;
  }
  const int ntiles = a.B * a.tiles_x * a.tiles_y;
  conv_tc3_kernel<BN, R1, TAP><<<ntiles < num_sms ? ntiles : num_sms, S3_THREADS, smem, st>>>(mh, ml, a);
  FRTM_CHECK_LAUNCH("conv_tc3");
  return FRTM_OK;
}

template <int BN>
static int launch_tc1(const CUtensorMap &mh, const CUtensorMap &ml, const TcArgs &a, cudaStream_t st) {
  constexpr int smem = S1_STAGES * 2 * S1_TILE_BYTES + 4 * 2 * BN * 128 + 8 * S1_NBAR + 16 + 84 * BN + 1024;
  static_assert(smem <= 227 * 1024, "conv_tc1: shared memory budget");
  static bool configured = false;
  static int num_sms = 0;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc1_kernel<BN, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("conv_tc1: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return FRTM_ELAUNCH; }
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
    configured = true;
  }
  const int ntiles = a.B * a.tiles_x * a.tiles_y;
  conv_tc1_kernel<BN, false, false><<<ntiles < num_sms ? ntiles : num_sms, S3_THREADS, smem, st>>>(mh, ml, a);
  FRTM_CHECK_LAUNCH("conv_tc1");
  return FRTM_OK;
}

}  // namespace frtm

using namespace frtm;

extern "C" int frtm_stem_conv_u8(const uint8_t *img, int B, int H, int W, const void *wt, const float *oscale, const float *bias,
                                 float *y, int ldy, int relu, void *stream) {
  FRTM_REQUIRE(img && wt && oscale && y && B > 0 && H > 0 && W > 0, "stem_conv: bad arguments");
  FRTM_REQUIRE(ldy >= 64 && (reinterpret_cast<uintptr_t>(wt) & 15) == 0, "stem_conv: 64 output channels, 16-byte aligned weights");
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  float s[3], b[3];
  for (int i = 0; i < 3; ++i) {        // same constants, rounded the same way, as frtm_normalize_u8 / frtm_stem_patches_u8
    volatile float rcp = 1.0f / stdv[i];
    s[i] = rcp * (float)(1.0 / 255.0);
    b[i] = -mean[i] / stdv[i];
  }
  TcArgs a;
  a.wt = (const __half *)wt; a.oscale = oscale; a.bias = bias; a.res = nullptr; a.res_hi = nullptr; a.res_lo = nullptr;
  a.y = y; a.y_nchw = nullptr; a.y_hi = nullptr; a.y_lo = nullptr; a.tapw = nullptr; a.y_tap = nullptr;
  a.r1_score = nullptr; a.r1_w = nullptr; a.r1_bias = nullptr; a.y_extra = nullptr; a.extra_ch = -1; a.yh_cout = 64;
  a.ldr = 0; a.ldrh = 0; a.ldy = ldy; a.y_coff = 0; a.ldyh = 0; a.yh_coff = 0;
  a.B = B; a.H = H; a.W = W; a.Cout = 64; a.kh = 1; a.kw = 1; a.pad = 0; a.relu = relu; a.stride = 1;
  a.Ho = (H + 6 - 7) / 2 + 1; a.Wo = (W + 6 - 7) / 2 + 1;
  a.tiles_x = cdiv(a.Wo, S3_TW); a.tiles_y = cdiv(a.Ho, S3_TH); a.kchunks = ST_KC;
  constexpr int smem = S1_STAGES * 2 * S1_TILE_BYTES + ST_KC * 2 * 64 * 128 + 16 * S1_STAGES + 64 + 84 * 64 +
                       2 * 2 * 3 * ST_RROWS * ST_RCOLS * 2 + 64 + 1024;
  static_assert(smem <= 227 * 1024, "conv_stem: shared memory budget");
  static bool configured = false;
  static int num_sms = 0;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("stem_conv: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return FRTM_ELAUNCH; }
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
    configured = true;
  }
  const int ntiles = a.B * a.tiles_x * a.tiles_y;
  conv_stem_kernel<<<ntiles < num_sms ? ntiles : num_sms, ST_THREADS, smem, (cudaStream_t)stream>>>(img, a, s[0], s[1], s[2], b[0],
                                                                                                    b[1], b[2]);
  FRTM_CHECK_LAUNCH("stem_conv");
  return FRTM_OK;
}

extern "C" int frtm_split_f16(const float *x, int64_t npix, int C, int ldx, void *hi, void *lo, int ldh, void *stream) {
  FRTM_REQUIRE(x && hi && lo && C % 4 == 0 && ldx % 4 == 0 && ldh % 8 == 0, "split_f16: bad arguments");
  const int64_t total = npix * (C / 4);
  split_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x, npix, C, ldx, (__half *)hi, (__half *)lo, ldh);
  FRTM_CHECK_LAUNCH("split_f16");
  return FRTM_OK;
}

extern "C" int frtm_conv2d_tc(const void *x_hi, const void *x_lo, int B, int H, int W, int Cin, int ldx, const void *wt,
                              const float *oscale, int bn_tile, const float *bias, const float *res, int ldr,
                              const void *res_hi, const void *res_lo, int ldrh, float *y, int ldy, int y_coff, float *y_nchw,
                              void *y_hi, void *y_lo, int ldyh, int yh_coff, int yh_cout, const float *tapw, float *y_tap,
                              const float *r1_score, const float *r1_w, const float *r1_bias, float *y_extra, int extra_ch,
                              int Cout, int kh, int kw, int stride, int relu, int kernel_select, void *stream) {
  FRTM_REQUIRE(x_hi && x_lo && wt && oscale && (y || y_nchw || y_hi || y_tap), "conv2d_tc: null pointer");
  FRTM_REQUIRE(kernel_select >= 0 && kernel_select <= 4,
               "conv2d_tc: kernel_select must be 0 (automatic), 1 (general tile kernel), 2 / 3 (CTA-pair kernel, N = 128 / 256) or 4 "
               "(persistent CTA-pair kernel)");
  const bool special = kernel_select == 0;
  FRTM_REQUIRE(!r1_score || r1_w, "conv2d_tc: the rank-1 score channel needs its weights");
  FRTM_REQUIRE(!y_extra || (extra_ch >= 0 && extra_ch < Cout), "conv2d_tc: extra_ch out of range");
  FRTM_REQUIRE(!y_tap || (tapw && Cout <= bn_tile && (reinterpret_cast<uintptr_t>(y_tap) & 15) == 0),
               "conv2d_tc: the tap-map output needs tapw, a single N tile (Cout <= bn_tile) and a 16-byte aligned buffer");
  FRTM_REQUIRE(Cin % TC_BK == 0 && ldx % 8 == 0, "conv2d_tc: Cin must be a multiple of 64 and ldx of 8 (got %d, %d)", Cin, ldx);
  FRTM_REQUIRE(kh == kw && (kh == 1 || kh == 3), "conv2d_tc: only 1x1 (pad 0) and 3x3 (pad 1) convolutions");
  FRTM_REQUIRE(stride == 1 || stride == 2, "conv2d_tc: stride must be 1 or 2");
  FRTM_REQUIRE((reinterpret_cast<uintptr_t>(x_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(x_lo) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(wt) & 15) == 0, "conv2d_tc: operands must be 16-byte aligned");
  FRTM_REQUIRE(!y_hi || (y_lo && ldyh % 2 == 0), "conv2d_tc: bad split output");
  // 3x3 / stride 1 / 64 input channels / one N tile of 32 or 64: the slab kernel with resident weights
  const bool slab3 = kh == 3 && stride == 1 && Cin == TC_BK && Cout <= bn_tile && (bn_tile == 64 || bn_tile == 32) &&
                     !(y_tap && bn_tile != 32) && !(r1_score && bn_tile != 64) && special;
  // 1x1 / stride 1 / up to 256 input channels / one N tile of 32 or 64, plain epilogue: the persistent streaming kernel
  const bool stream1 = kh == 1 && stride == 1 && Cin <= 4 * TC_BK && Cout <= bn_tile && (bn_tile == 64 || bn_tile == 32) &&
                       !y_tap && !r1_score && special;
  const int box_w = (slab3 || stream1) ? S3_TW : 0, box_h = slab3 ? S3_TH + 2 : (stream1 ? S3_TH : 0);
  CUtensorMap mh, ml;
  int rc = make_act_map(&mh, (const __half *)x_hi, B, H, W, Cin, ldx, stride, box_w, box_h);
  if (rc) return rc;
  rc = make_act_map(&ml, (const __half *)x_lo, B, H, W, Cin, ldx, stride, box_w, box_h);
  if (rc) return rc;
  TcArgs a;
  a.wt = (const __half *)wt; a.oscale = oscale; a.bias = bias; a.res = res; a.res_hi = (const __half *)res_hi;
  a.res_lo = (const __half *)res_lo; a.y = y; a.y_nchw = y_nchw; a.y_hi = (__half *)y_hi; a.y_lo = (__half *)y_lo;
  a.tapw = tapw; a.y_tap = y_tap;
  a.r1_score = r1_score; a.r1_w = r1_w; a.r1_bias = r1_bias; a.y_extra = y_extra; a.extra_ch = extra_ch;
  a.yh_cout = (yh_cout > 0 && yh_cout < Cout) ? yh_cout : Cout;
  a.ldr = ldr; a.ldrh = ldrh; a.ldy = ldy; a.y_coff = y_coff; a.ldyh = ldyh; a.yh_coff = yh_coff;
  a.B = B; a.H = H; a.W = W; a.Cout = Cout; a.kh = kh; a.kw = kw; a.pad = kh / 2; a.relu = relu; a.stride = stride;
  a.Ho = (H + 2 * a.pad - kh) / stride + 1; a.Wo = (W + 2 * a.pad - kw) / stride + 1;
  a.tiles_x = cdiv(a.Wo, TC_TW); a.tiles_y = cdiv(a.Ho, TC_TH); a.kchunks = Cin / TC_BK;
  const int ntiles_n = cdiv(Cout, bn_tile);
  dim3 grid((unsigned)(B * a.tiles_x * a.tiles_y), (unsigned)ntiles_n);
  cudaStream_t st = (cudaStream_t)stream;
  const bool r1 = r1_score != nullptr, tap = y_tap != nullptr;
  FRTM_REQUIRE(!(r1 && tap), "conv2d_tc: rank-1 input and tap-map output cannot be combined");
  // wide convs (Cout a multiple of 128, weights in the N tile 64 packing, plain epilogue): the CTA-pair kernel.  Pair tile
  // N = 256 when that still gives every SM a CTA, else N = 128 (twice the CTAs at 1.5x the L2 traffic per MAC).
  const bool pair_ok = bn_tile == 64 && Cout % 128 == 0 && !r1 && !tap;
  FRTM_REQUIRE(kernel_select < 2 || (pair_ok && (kernel_select != 3 || Cout % 256 == 0)),
               "conv2d_tc: the CTA-pair kernels need the N tile 64 packing, Cout %% 128 == 0 (N = 256: %% 256) and a plain epilogue");
  if (kernel_select >= 2 || (special && pair_ok)) {
    static int num_sms = 0;
    if (num_sms == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
    }
    const int pairs = cdiv((int64_t)B * a.tiles_x * a.tiles_y, 2);
    // several N = 128 items per TPC: the persistent kernel (store of item i behind the products of item i+1); else one
    // tile per CTA, N = 256 when that still gives 0.7 CTAs per SM
    bool persistent = kernel_select == 4, wide = kernel_select == 3;
    if (kernel_select == 0) {
      persistent = 2 * pairs * (Cout / 128) >= 5 * (num_sms / 2);
      wide = !persistent && Cout % 256 == 0 && 10 * (2 * pairs * (Cout / 256)) >= 7 * num_sms;
    }
    CUtensorMap mw;
    rc = make_weight_map(&mw, (const __half *)wt, (int64_t)(Cout / 64) * a.kh * a.kw * a.kchunks * 128);
    if (rc) return rc;
    if (persistent) return launch_tc2p(mh, ml, mw, a, num_sms, st);
    return wide ? launch_tc2<128>(mh, ml, mw, a, st) : launch_tc2<64>(mh, ml, mw, a, st);
  }
  if (stream1) {
    a.tiles_x = cdiv(a.Wo, S3_TW); a.tiles_y = cdiv(a.Ho, S3_TH);
    return bn_tile == 64 ? launch_tc1<64>(mh, ml, a, st) : launch_tc1<32>(mh, ml, a, st);
  }
  if (slab3) {
    a.tiles_x = cdiv(a.Wo, S3_TW); a.tiles_y = cdiv(a.Ho, S3_TH);
    if (bn_tile == 64) return r1 ? launch_tc3<64, true, false>(mh, ml, a, st) : launch_tc3<64, false, false>(mh, ml, a, st);
    return tap ? launch_tc3<32, false, true>(mh, ml, a, st) : launch_tc3<32, false, false>(mh, ml, a, st);
  }
  FRTM_REQUIRE(!r1 || bn_tile == 64 || bn_tile == 80, "conv2d_tc: the rank-1 score channel is built for N tiles 64 and 80");
  FRTM_REQUIRE(!tap || bn_tile == 32, "conv2d_tc: the tap-map output is built for the N tile 32");
  FRTM_REQUIRE(r1 || !y_extra, "conv2d_tc: y_extra needs the rank-1 epilogue");
  switch (bn_tile) {
    case 32: return tap ? launch_tc<32, 2, false, true>(mh, ml, a, grid, st) : launch_tc<32, 2, false, false>(mh, ml, a, grid, st);
    case 64: return r1 ? launch_tc<64, 2, true, false>(mh, ml, a, grid, st) : launch_tc<64, 2, false, false>(mh, ml, a, grid, st);
    case 80: return r1 ? launch_tc<80, 2, true, false>(mh, ml, a, grid, st) : launch_tc<80, 2, false, false>(mh, ml, a, grid, st);
    case 128: return launch_tc<128, 3, false, false>(mh, ml, a, grid, st);   // one CTA per SM anyway (TMEM): a third stage hides the L2 latency
    default: set_error("conv2d_tc: unsupported N tile %d (32, 64, 80, 128)", bn_tile); return FRTM_EINVAL;
  }
}

extern "C" int frtm_rank1_finish(const float *y_in, int ldin, int n_obj, float *y_out, int ldout, const float *score,
                                 const float *wx, const float *bias, int B, int H, int W, int Cout, int relu, void *y_hi,
                                 void *y_lo, int ldh, float *extra, void *stream) {
  FRTM_REQUIRE(y_in && score && wx && Cout <= 65 && Cout >= 1 && n_obj >= 1, "rank1_finish: bad arguments");
  FRTM_REQUIRE(!y_hi || (ldh % 4 == 0), "rank1_finish: split planes need a channel stride that is a multiple of 4");
  const int64_t total = (int64_t)B * H * W * ((Cout + 3) / 4);
  rank1_finish_kernel<<<cdiv(total, 256), 256, 10 * Cout * sizeof(float), (cudaStream_t)stream>>>(
      y_in, ldin, n_obj, y_out, ldout, score, wx, bias, B, H, W, Cout, relu, (__half *)y_hi, (__half *)y_lo, ldh, extra);
  FRTM_CHECK_LAUNCH("rank1_finish");
  return FRTM_OK;
}

extern "C" int frtm_pack_tc_1x1(const float *W, int Cout, int Cin, int bn_tile, void *wt, float *oscale, void *stream) {
  FRTM_REQUIRE(W && wt && oscale && Cin % TC_BK == 0 && bn_tile % 8 == 0, "pack_tc_1x1: bad arguments");
  const int cout_pad = cdiv(Cout, bn_tile) * bn_tile;
  pack_tc_1x1_kernel<<<cout_pad, 256, 0, (cudaStream_t)stream>>>(W, Cout, Cin, bn_tile, (__half *)wt, oscale);
  FRTM_CHECK_LAUNCH("pack_tc_1x1");
  return FRTM_OK;
}
