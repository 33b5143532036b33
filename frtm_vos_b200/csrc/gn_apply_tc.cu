// GN/CG operator  g_i = X_i^T [ sw_i (S_i (X_i * p) - use_y t_i) ]  for one memory sample per CTA, with both
// contractions over the sample on the tensor cores (tcgen05, fp32 accumulation in TMEM) so the kernel is bound by
// the bytes of the sample and not by fp32 FMA issue (the CUDA-core kernels in target_model.cu need 18 FMAs per
// loaded element, 8.2 FLOP/B — above the fp32 ridge at HBM speed).
//
// The sample is read as its split tile image (frtm_split_samples, written once at insert time): tiles of
// [96 channel rows][64 pixels] fp16, hi and lo planes with 16 x = hi + lo, rows in the 128-byte swizzled layout, so
// one bulk copy per tile lands an operand both contractions can use without any data movement by threads:
//   phase 1  tap maps   Y[q][tap]  = sum_c X[c][q] p[c][tap]     A = tile pair as an MN-major operand (M = 128 pixels,
//                                                                K = channels), B = p (16 x c, K-major), N = 16
//   phase 2  scores     s[q] = sum_tap Y[tap][q + tap];  v = sw (S s - use_y t)      CUDA cores, shared memory only
//   phase 3  gradient   g[c][tap] = sum_q X[c][q] v[q - tap]     A = the same tile as a K-major operand (M = channels,
//                                                                K = 64 pixels), B = 9 shifted copies of v (16 x 64)
// p, v are split the same way (per-launch / per-sample power-of-two scale, hi + lo fp16) and every product is issued
// as hi*hi + hi*lo + lo*hi.  Phase 3 accumulates GC_FOLD tiles (K = 256) inside the tensor core and folds the slot into
// fp32 registers with round-to-nearest adds (the tensor-core accumulator truncates, see conv_tc.cu).
// Streaming: warp 0 issues the bulk copies through a 3-slot mbarrier ring — 2 * ntiles tiles per sample, the first pass
// from HBM, the second from L2; warp 1 issues the MMAs; warps 2-5 drain TMEM, run phase 2 and build the v operand.
// 97 KB of shared memory and 64 TMEM columns per CTA: two CTAs per SM, so one sample's phase 3 overlaps another's phase 1.
#include "common.cuh"
#include "target_model.cuh"
#include "tc_ptx.cuh"

namespace frtm {

constexpr int GC_SLOTS = 3;
constexpr int GC_FOLD = 4;
constexpr int GC_THREADS = 192;
constexpr int GC_VSLOT_BYTES = 4096;       // [hi|lo][16 rows][64] fp16
constexpr int GC_NBARS = 18;

// SW128 shared-memory matrix descriptor with explicit leading/stride byte offsets (MN-major operands use both)
__device__ __forceinline__ uint64_t umma_desc_ls(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// power of two that brings amax into [2^9, 2^10)
__device__ __forceinline__ float pow2_scale(float amax) {
  if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
  int e;
  frexpf(amax, &e);
  e = max(min(10 - e, 100), -100);
  return ldexpf(1.f, e);
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void drain_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__global__ void __launch_bounds__(GC_THREADS, 2) gn_apply_tc_kernel(const GaArgs a, int ntiles, int tile_bytes) {
  const float *__restrict__ S = a.S, *__restrict__ T = a.T, *__restrict__ sw = a.sw, *__restrict__ pvec = a.pvec;
  const __half *__restrict__ XS = a.XS;
  float *__restrict__ partial = a.partial;
  const int c = a.c, h = a.h, w = a.w, use_y = a.use_y;
  if (a.table) {
    const int o = blockIdx.y;
    S = reinterpret_cast<const float *>(a.table[1 * a.n_obj + o]);
    T = reinterpret_cast<const float *>(a.table[2 * a.n_obj + o]);
    sw = reinterpret_cast<const float *>(a.table[3 * a.n_obj + o]);
    pvec = reinterpret_cast<const float *>(a.table[(use_y ? 4 : 5) * a.n_obj + o]);
    XS = reinterpret_cast<const __half *>(a.table[7 * a.n_obj + o]);
    partial += (int64_t)o * a.cap * c * 9;
  }
  const int i = blockIdx.x;
  const int n = c * 9;
  const int hw = h * w, wp = w + 2, npad = (h + 2) * wp;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float wgt = sw[i];
  if (wgt == 0.f) {
    for (int k = tid; k < n; k += GC_THREADS) partial[(int64_t)i * n + k] = 0.f;
    return;
  }

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t ring = base;
  const uint32_t vring = ring + GC_SLOTS * tile_bytes;
  uint8_t *vring_g = gen + GC_SLOTS * tile_bytes;
  float *sp = reinterpret_cast<float *>(vring_g + 2 * GC_VSLOT_BYTES);
  float *vp = sp + npad;
  uint8_t *tail = reinterpret_cast<uint8_t *>(vp + npad);
  tail = gen + (((tail - gen) + 15) & ~(size_t)15);
  const uint32_t bars = base + (uint32_t)(tail - gen);
  float *red = reinterpret_cast<float *>(tail + 8 * GC_NBARS);            // 8 floats
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(red + 8);
  const uint32_t bar_full = bars, bar_empty = bars + 8 * 3, bar_accfull = bars + 8 * 6, bar_accfree = bars + 8 * 8,
                 bar_vready = bars + 8 * 10, bar_vfree = bars + 8 * 12, bar_foldfull = bars + 8 * 14,
                 bar_foldfree = bars + 8 * 16;
  const int plane_bytes = tile_bytes >> 1;

  if (tid == 0) {
    for (int s = 0; s < 3; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_accfull + 8 * s, 1); mbar_init(bar_accfree + 8 * s, 4);
      mbar_init(bar_vready + 8 * s, 4); mbar_init(bar_vfree + 8 * s, 1);
      mbar_init(bar_foldfull + 8 * s, 1); mbar_init(bar_foldfree + 8 * s, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- p operand: B[16 taps][c] (K-major, two 64-channel chunks), scaled to [2^9, 2^10), split hi + lo ----
  float amax = 0.f;
  for (int k = tid; k < n; k += GC_THREADS) amax = fmaxf(amax, fabsf(pvec[k]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if (lane == 0) red[warp] = amax;
  for (int k = tid; k < 2 * npad; k += GC_THREADS) sp[k] = 0.f;      // sp and vp are contiguous
  __syncthreads();
  amax = 0.f;
#pragma unroll
  for (int k = 0; k < GC_THREADS / 32; ++k) amax = fmaxf(amax, red[k]);
  const float pscale = pow2_scale(amax);
  for (int idx = tid; idx < 16 * 128; idx += GC_THREADS) {
    const int t = idx >> 7, ch = idx & 127;
    const float v = (t < 9 && ch < c) ? pvec[ch * 9 + t] * pscale : 0.f;
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    const int chunk = ch >> 6, kk = ch & 63;
    const int off = chunk * GC_VSLOT_BYTES + t * 128 + (((kk >> 3) ^ (t & 7)) << 4) + (kk & 7) * 2;
    *reinterpret_cast<__half *>(vring_g + off) = hi;
    *reinterpret_cast<__half *>(vring_g + off + 2048) = lo;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int npairs = ntiles >> 1;

  if (warp == 0) {
    // ===== producer: 2 * ntiles bulk copies of one tile (hi + lo planes are contiguous) =====
    if (lane == 0) {
      const uint8_t *src = reinterpret_cast<const uint8_t *>(XS) + (int64_t)i * ntiles * tile_bytes;
      for (int it = 0; it < 2 * ntiles; ++it) {
        const int s = it % GC_SLOTS;
        mbar_wait(bar_empty + 8 * s, ((it / GC_SLOTS) & 1) ^ 1);
        mbar_expect_tx(bar_full + 8 * s, tile_bytes);
        const int j = it < ntiles ? it : it - ntiles;
        bulk_load(ring + s * tile_bytes, src + (int64_t)j * tile_bytes, tile_bytes, bar_full + 8 * s);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc3 = umma_idesc(128, 16);
      constexpr uint32_t idesc1 = idesc3 | (1u << 15);        // A is MN-major (pixels contiguous, K = channel rows)
      // phase 1: one M = 128 (two tiles) x N = 16 x K = c product per tile pair
      for (int tp = 0; tp < npairs; ++tp) {
        const int it0 = 2 * tp, it1 = it0 + 1;
        const int s0 = it0 % GC_SLOTS, s1 = it1 % GC_SLOTS;
        mbar_wait(bar_full + 8 * s0, (it0 / GC_SLOTS) & 1);
        mbar_wait(bar_full + 8 * s1, (it1 / GC_SLOTS) & 1);
        const int as = tp & 1;
        mbar_wait(bar_accfree + 8 * as, ((tp >> 1) & 1) ^ 1);
        tc_fence_after();
        const int slo = min(s0, s1), shi = max(s0, s1);       // TMEM lanes 0-63 <- the tile in the lower slot
        const uint32_t lbo = (uint32_t)(shi - slo) * tile_bytes;
        const uint32_t a0 = ring + slo * tile_bytes;
        const uint32_t tacc = tmem_base + as * 16;
        for (int ks = 0; ks < c / 16; ++ks) {
          const uint64_t a_hi = umma_desc_ls(a0 + ks * 2048, lbo, 1024);
          const uint64_t a_lo = umma_desc_ls(a0 + plane_bytes + ks * 2048, lbo, 1024);
          const uint32_t bb = vring + (ks >> 2) * GC_VSLOT_BYTES + (ks & 3) * 32;
          const uint64_t b_hi = umma_desc_ls(bb, 0, 1024), b_lo = umma_desc_ls(bb + 2048, 0, 1024);
          umma_f16(tacc, a_hi, b_hi, idesc1, ks > 0 ? 1u : 0u);
          umma_f16(tacc, a_hi, b_lo, idesc1, 1u);
          umma_f16(tacc, a_lo, b_hi, idesc1, 1u);
        }
        umma_commit(bar_empty + 8 * s0);
        umma_commit(bar_empty + 8 * s1);
        umma_commit(bar_accfull + 8 * as);
      }
      // phase 3: M = 128 (c channel rows valid) x N = 16 x K = 64 pixels per tile, GC_FOLD tiles per accumulator slot
      for (int j = 0; j < ntiles; ++j) {
        const int it = ntiles + j, s = it % GC_SLOTS;
        mbar_wait(bar_full + 8 * s, (it / GC_SLOTS) & 1);
        const int vs = j & 1;
        mbar_wait(bar_vready + 8 * vs, (j >> 1) & 1);
        const int grp = j / GC_FOLD, fs = grp & 1;
        const bool first = (j % GC_FOLD) == 0;
        if (first) mbar_wait(bar_foldfree + 8 * fs, ((grp >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t a0 = ring + s * tile_bytes, b0 = vring + vs * GC_VSLOT_BYTES;
        const uint32_t tacc = tmem_base + 32 + fs * 16;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t a_hi = umma_desc_ls(a0 + kk * 32, 0, 1024), a_lo = umma_desc_ls(a0 + plane_bytes + kk * 32, 0, 1024);
          const uint64_t b_hi = umma_desc_ls(b0 + kk * 32, 0, 1024), b_lo = umma_desc_ls(b0 + 2048 + kk * 32, 0, 1024);
          umma_f16(tacc, a_hi, b_hi, idesc3, (first && kk == 0) ? 0u : 1u);
          umma_f16(tacc, a_hi, b_lo, idesc3, 1u);
          umma_f16(tacc, a_lo, b_hi, idesc3, 1u);
        }
        umma_commit(bar_empty + 8 * s);
        umma_commit(bar_vfree + 8 * vs);
        if ((j % GC_FOLD) == GC_FOLD - 1 || j == ntiles - 1) umma_commit(bar_foldfull + 8 * fs);
      }
    }
  } else {
    // ===== drain warps: thread dt owns TMEM lane dt =====
    const int q4 = warp & 3;
    const int dt = q4 * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const float yscale = 1.f / (GC_ACT_SCALE * pscale);
    // ---- phase 1 drain: scatter the 9 tap values of the own pixel into the padded score map, one tap per round so
    //      no two threads ever add to the same entry in a round (fixed order -> deterministic) ----
    for (int tp = 0; tp < npairs; ++tp) {
      const int as = tp & 1;
      mbar_wait(bar_accfull + 8 * as, (tp >> 1) & 1);
      tc_fence_after();
      float y[16];
      tmem_ld16(tlane + as * 16, y);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_accfree + 8 * as);
      const bool swapped = ((2 * tp) % GC_SLOTS) > ((2 * tp + 1) % GC_SLOTS);
      const int tile = 2 * tp + (((dt >> 6) & 1) ^ (swapped ? 1 : 0));
      const int q = tile * GC_TILE + (dt & 63);
      const bool valid = q < hw;
      const int py = q / w, px = q - py * w;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int yy = py - (t / 3 - 1), xx = px - (t % 3 - 1);
        if (valid && yy >= 0 && yy < h && xx >= 0 && xx < w) sp[(yy + 1) * wp + xx + 1] += y[t] * yscale;
        drain_sync();
      }
    }
    // ---- phase 2: v = sw (S s - use_y t), its maximum, and clearing of the operand ring (it held p until now) ----
    const float *Si = S + (int64_t)i * 9 * hw, *Ti = T + (int64_t)i * hw;
    float vmax = 0.f;
    for (int q = dt; q < hw; q += 128) {
      const int py = q / w, px = q - py * w;
      float av = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) av = fmaf(__ldg(Si + t * hw + q), sp[(py + t / 3) * wp + px + t % 3], av);
      if (use_y) av -= __ldg(Ti + q);
      av *= wgt;
      vp[(py + 1) * wp + px + 1] = av;
      vmax = fmaxf(vmax, fabsf(av));
    }
    {
      uint4 *z = reinterpret_cast<uint4 *>(vring_g) + dt * 4;
      const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
      z[0] = zero; z[1] = zero; z[2] = zero; z[3] = zero;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (lane == 0) red[q4] = vmax;
    drain_sync();
    const float vscale = pow2_scale(fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])));
    if (a.dbg != nullptr && use_y && blockIdx.x == 0 && blockIdx.y == 0) {
      for (int k = dt; k < 2 * npad; k += 128) a.dbg[k] = sp[k];
    }
    // ---- phase 3: build the shifted-v operand of each tile (thread = tap row t, 8-pixel chunk), fold finished slots ----
    float g[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) g[t] = 0.f;
    const int t = dt >> 3, ch8 = dt & 7;
    const int dy = t / 3 - 1, dx = t % 3 - 1;
    auto fold = [&](int grp) {
      const int fs = grp & 1;
      mbar_wait(bar_foldfull + 8 * fs, (grp >> 1) & 1);
      tc_fence_after();
      float f[16];
      tmem_ld16(tlane + 32 + fs * 16, f);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_foldfree + 8 * fs);
#pragma unroll
      for (int u = 0; u < 16; ++u) g[u] += f[u];
    };
    for (int j = 0; j < ntiles; ++j) {
      const int vs = j & 1;
      mbar_wait(bar_vfree + 8 * vs, ((j >> 1) & 1) ^ 1);
      if (t < 9) {
        const int q0 = j * GC_TILE + ch8 * 8;
        int py = q0 / w, px = q0 - py * w;
        __align__(16) __half hh[8];
        __align__(16) __half ll[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float v = (q0 + e < hw) ? vp[(py + 1 - dy) * wp + px + 1 - dx] * vscale : 0.f;
          hh[e] = __float2half_rn(v);
          ll[e] = __float2half_rn(v - __half2float(hh[e]));
          if (++px == w) { px = 0; ++py; }
        }
        uint8_t *row = vring_g + vs * GC_VSLOT_BYTES + t * 128 + ((ch8 ^ (t & 7)) << 4);
        *reinterpret_cast<uint4 *>(row) = *reinterpret_cast<const uint4 *>(hh);
        *reinterpret_cast<uint4 *>(row + 2048) = *reinterpret_cast<const uint4 *>(ll);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_vready + 8 * vs);
      if (j > 0 && (j % GC_FOLD) == 0) fold(j / GC_FOLD - 1);
    }
    fold((ntiles - 1) / GC_FOLD);
    if (dt < c) {
      const float gscale = 1.f / (GC_ACT_SCALE * vscale);
      float *dst = partial + (int64_t)i * n + dt * 9;
#pragma unroll
      for (int u = 0; u < 9; ++u) dst[u] = g[u] * gscale;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64) : "memory");
  }
}

// n samples (c, hw) fp32 -> their split tile images
__global__ void __launch_bounds__(256) split_samples_kernel(const float *__restrict__ src, __half *__restrict__ dst, int c,
                                                            int hw, int64_t items_per_sample) {
  const int k = blockIdx.y;
  const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= items_per_sample) return;
  gc_split_item(src + (int64_t)k * c * hw, dst + (int64_t)k * gc_sample_halves(c, hw), c, hw, item);
}

static size_t gc_smem_bytes(int c, int h, int w) {
  const size_t npad = (size_t)(h + 2) * (w + 2);
  return 1024 + (size_t)GC_SLOTS * 2 * c * 128 + 2 * GC_VSLOT_BYTES + 2 * npad * 4 + 16 + 8 * GC_NBARS + 32 + 16;
}

bool gn_apply_tc_supported(int c, int h, int w) {
  return c % 16 == 0 && c >= 16 && c <= 128 && gc_smem_bytes(c, h, w) <= 227 * 1024;
}

int gn_apply_tc_launch(const GaArgs &a, cudaStream_t st) {
  const size_t smem = gc_smem_bytes(a.c, a.h, a.w);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(gn_apply_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("gn_apply_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return FRTM_ELAUNCH; }
    cudaFuncSetAttribute(gn_apply_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    configured = smem;
  }
  const dim3 grid(a.cap, a.table ? a.n_obj : 1);
  gn_apply_tc_kernel<<<grid, GC_THREADS, smem, st>>>(a, gc_ntiles(a.h * a.w), 2 * a.c * 128);
  FRTM_CHECK_LAUNCH("gn_apply_tc");
  return FRTM_OK;
}

}  // namespace frtm

using namespace frtm;

extern "C" int64_t frtm_split_sample_bytes(int c, int hw) { return gc_sample_halves(c, hw) * 2; }

extern "C" int frtm_split_samples(const float *samples, int n, int c, int hw, void *split, void *stream) {
  FRTM_REQUIRE(samples && split && n >= 1 && c % 8 == 0, "split_samples: bad arguments (c must be a multiple of 8)");
  FRTM_REQUIRE((reinterpret_cast<uintptr_t>(split) & 15) == 0, "split_samples: output must be 16-byte aligned");
  const int64_t items = (int64_t)gc_ntiles(hw) * c * 8;
  split_samples_kernel<<<dim3(cdiv(items, 256), n), 256, 0, (cudaStream_t)stream>>>(samples, (__half *)split, c, hw, items);
  FRTM_CHECK_LAUNCH("split_samples");
  return FRTM_OK;
}
