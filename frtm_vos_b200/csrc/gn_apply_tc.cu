// GN/CG operator  g_i = X_i^T [ sw_i (S_i (X_i * p) - use_y t_i) ]  over the frame memory, with both contractions over
// a sample on the tensor cores (tcgen05, fp32 accumulation in TMEM) so the kernel is bound by the bytes of the sample
// and not by fp32 FMA issue (the CUDA-core kernels in target_model.cu need 18 FMAs per loaded element, 8.2 FLOP/B —
// above the fp32 ridge at HBM speed).
//
// Each sample is read as its operator image (target_model.cuh, written once at insert time): tiles of
// [c channel rows][64 pixels] fp16, hi and lo planes with 16 x = hi + lo, rows in the 128-byte swizzled layout, so one
// bulk copy per tile lands an operand both contractions can use without any data movement by threads, followed by
// the stencil / U^T w^2 y rows in 256-pixel chunks:
//   phase 1  tap maps   Y[q][tap]  = sum_c X[c][q] p[c][tap]     A = tile pair as an MN-major operand (M = 128 pixels,
//                                                                K = channels), B = p (16 x c, K-major), N = 16
//   phase 2  scores     s[q] = sum_tap Y[tap][q + tap];  v = sw (S s - use_y t)      CUDA cores, shared memory only
//   phase 3  gradient   g[c][tap] = sum_q X[c][q] v[q - tap]     A = the same tile as a K-major operand (M = channels,
//                                                                K = 64 pixels), B = 9 shifted copies of v (16 x 64)
// p, v are split the same way (per-object / per-sample power-of-two scale, hi + lo fp16) and every product is issued
// as hi*hi + hi*lo + lo*hi.  Phase 3 accumulates GC_FOLD tiles (K = 256) inside the tensor core and folds the slot into
// fp32 registers with round-to-nearest adds (the tensor-core accumulator truncates, see conv_tc.cu).
//
// One CTA per sample, two CTAs per SM (115 KB of shared memory, 256 TMEM columns each), so every active sample of a
// 3-object update is in flight at once and one sample's phase 3 overlaps another's phase 1.  Warp 0 streams the image
// through a 3-slot mbarrier ring (ntiles tiles from HBM, the stencil chunks, the ntiles tiles again in REVERSE order, so
// the most recently streamed tiles — the ones still in L2 — are re-read first); warp 1 issues the MMAs; warps 2-5 drain
// TMEM, run phase 2 and build the v operand.  Measured (profiles/r01_gn_tc_timeline.md; -DFRTM_DEBUG_NO_MMA removes only
// 14 % of the time): the kernel is bound by the serial hand-offs between the roles, not by the tensor pipe or by
// bandwidth, so work per hand-off is batched — two tile pairs (256 pixels) per phase-1 accumulator slot, drained with
// one gather pass per group; two tiles per v-operand slot; A_hi x [B_hi | B_lo] as one N = 32 product plus A_lo x B_hi —
// and the per-sample partial reduction and the CG vector step run in the kernel's tail (last CTA per object).
#include "common.cuh"
#include "target_model.cuh"
#include "tc_ptx.cuh"

namespace frtm {

constexpr int GC_MAX_SLOTS = 8;            // ring slots (one tile each): as many as fit, see gc_slots()
constexpr int GC_FOLD = 4;                 // tiles accumulated inside the tensor core between fp32 register folds
constexpr int GC_VSTEP = 2;                // tiles per v-operand slot
constexpr int GC_PGROUP = 2;               // tile pairs per phase-1 accumulator slot
constexpr int GC_THREADS = 192;
constexpr int GC_VTILE_BYTES = 4096;       // [hi|lo][16 rows][64] fp16
constexpr int GC_VSLOT_BYTES = GC_VSTEP * GC_VTILE_BYTES;
constexpr int GC_NBARS = 2 * GC_MAX_SLOTS + 12;
constexpr int GC_P1_COLS = 48;             // phase 1: one 16-column accumulator per pass (hi*hi, hi*lo, lo*hi) of a pair
constexpr int GC_P3_COLS = 48;             // phase 3: A_hi x [B_hi|B_lo] (32 columns) and A_lo x B_hi (16) of one tile parity
constexpr int GC_TMEM_COLS = 256;          // 2 slots x max(GC_PGROUP * GC_P1_COLS, GC_VSTEP * GC_P3_COLS) = 192

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

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void drain_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// Timeline tracing (debug builds only, -DGC_TRACE): CTA (0,0) of an RHS launch stamps globaltimer values into a.dbg
// (as 64-bit words after the 2*npad floats of the map dump): slot k of role r at dbg64[r * 64 + k].
#ifdef GC_TRACE
#define GC_STAMP(role, k)                                                                                        \
  do {                                                                                                           \
    if (a.dbg != nullptr && use_y && blockIdx.x == 0 && blockIdx.y == 0 && (k) < 64) {                           \
      unsigned long long t__;                                                                                    \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                                                    \
      reinterpret_cast<unsigned long long *>(a.dbg + 2 * npad + 2)[(role) * 64 + (k)] = t__;                     \
    }                                                                                                            \
  } while (0)
#else
#define GC_STAMP(role, k) do { } while (0)
#endif

// the operator for the sample of this CTA (returns early, after zeroing its partial row, for inactive memory slots)
__device__ __forceinline__ void gc_sample(const GaArgs &a, const GcParams &P) {
  const int c = a.c, h = a.h, w = a.w, use_y = a.use_y;
  const int n = c * 9;
  const int hw = h * w, wp = w + 2, npad = (h + 2) * wp;
  const int tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_idx();
  const int ntiles = P.ntiles, nchunks = P.nchunks, tile_bytes = P.tile_bytes, NS = P.slots;
  const int cps = tile_bytes >= 2 * GC_CHUNK_BYTES ? 2 : 1;      // stencil chunks per ring slot
  const int ncl = (nchunks + cps - 1) / cps;                      // stencil loads
  const int plane_bytes = tile_bytes >> 1;
  const int npairs = ntiles >> 1;
  const int i = blockIdx.x;
  const float *sw = a.sw, *pvec = a.pvec;
  const __half *xs = a.XS;
  float *part = a.partial + (int64_t)i * n;
  if (a.table) {
    const int o = blockIdx.y;
    sw = reinterpret_cast<const float *>(a.table[3 * a.n_obj + o]);
    // RHS pass linearises at the filter itself, CG passes apply the operator to the direction p (= cg_state[0:n])
    pvec = reinterpret_cast<const float *>(a.table[(use_y ? 4 : 5) * a.n_obj + o]);
    xs = reinterpret_cast<const __half *>(a.table[7 * a.n_obj + o]);
    part += (int64_t)o * a.cap * n;
  }
  GC_STAMP(0, 0);
  const float wgt = sw[i];
  if (wgt == 0.f) {
    for (int k = tid; k < n; k += GC_THREADS) part[k] = 0.f;
    return;
  }
  const uint8_t *img = reinterpret_cast<const uint8_t *>(xs) + (int64_t)i * P.image_bytes;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t ring = base;
  // operand buffer: phase 1 keeps p here (2 chunks x [hi|lo][16][64]), phase 3 the two v-operand slots
  const uint32_t obuf = ring + NS * tile_bytes;
  uint8_t *obuf_g = gen + NS * tile_bytes;
  float *sp = reinterpret_cast<float *>(obuf_g + 2 * GC_VSLOT_BYTES);
  float *vp = sp + npad;
  float *ybuf = vp + npad;                       // [9][GC_PGROUP * 128] tap values of one phase-1 drain group
  uint8_t *tail = reinterpret_cast<uint8_t *>(ybuf + 9 * GC_PGROUP * 128);
  tail = gen + (((tail - gen) + 15) & ~(size_t)15);
  const uint32_t bars = base + (uint32_t)(tail - gen);
  float *red = reinterpret_cast<float *>(tail + 8 * GC_NBARS);            // 8 floats
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(red + 8);
  const uint32_t bar_full = bars, bar_empty = bars + 8 * GC_MAX_SLOTS;
  const uint32_t bar_accfull = bars + 8 * (2 * GC_MAX_SLOTS), bar_accfree = bar_accfull + 16, bar_vready = bar_accfull + 32,
                 bar_vfree = bar_accfull + 48, bar_foldfull = bar_accfull + 64, bar_foldfree = bar_accfull + 80;

  if (tid == 0) {
    for (int s = 0; s < GC_MAX_SLOTS; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_accfull + 8 * s, 1); mbar_init(bar_accfree + 8 * s, 4);
      mbar_init(bar_vready + 8 * s, 4); mbar_init(bar_vfree + 8 * s, 1);
      mbar_init(bar_foldfull + 8 * s, 1); mbar_init(bar_foldfree + 8 * s, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(GC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- p operand: B[16 taps][c] (K-major, two 64-channel chunks), scaled to [2^9, 2^10), split hi + lo ----
  constexpr int PV = 6;                                   // c*9 <= 1024 values over 192 threads
  float pv[PV];
  float amax = 0.f;
#pragma unroll
  for (int k = 0; k < PV; ++k) {
    const int idx = tid + k * GC_THREADS;
    pv[k] = idx < n ? pvec[idx] : 0.f;
    amax = fmaxf(amax, fabsf(pv[k]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if (lane == 0) red[warp] = amax;
  for (int k = tid; k < (2 * GC_VSLOT_BYTES) / 16; k += GC_THREADS) reinterpret_cast<uint4 *>(obuf_g)[k] = make_uint4(0u, 0u, 0u, 0u);
  for (int k = tid; k < 2 * npad; k += GC_THREADS) sp[k] = 0.f;      // sp and vp are contiguous
  __syncthreads();
  amax = 0.f;
#pragma unroll
  for (int k = 0; k < GC_THREADS / 32; ++k) amax = fmaxf(amax, red[k]);
  const float pscale = pow2_scale(amax);
#pragma unroll
  for (int k = 0; k < PV; ++k) {
    const int idx = tid + k * GC_THREADS;
    if (idx < n) {
      const int ch = idx / 9, t = idx - ch * 9;
      const float v = pv[k] * pscale;
      const __half hi = __float2half_rn(v);
      const __half lo = __float2half_rn(v - __half2float(hi));
      const int chunk = ch >> 6, k6 = ch & 63;
      const int off = chunk * GC_VTILE_BYTES + t * 128 + (((k6 >> 3) ^ (t & 7)) << 4) + (k6 & 7) * 2;
      *reinterpret_cast<__half *>(obuf_g + off) = hi;
      *reinterpret_cast<__half *>(obuf_g + off + 2048) = lo;
    }
  }
  fence_async_smem();
  tc_fence_before();
  // the producer only needs the barriers (first __syncthreads above): it signals its share of the setup and starts
  // streaming while the other warps finish the p operand
  if (warp == 0) asm volatile("bar.arrive 2, %0;" ::"n"(GC_THREADS) : "memory");
  else asm volatile("bar.sync 2, %0;" ::"n"(GC_THREADS) : "memory");
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) GC_STAMP(0, 1);

  if (warp == 0) {
    // ===== producer: ntiles tiles, nchunks stencil chunks, ntiles tiles again =====
    {
      const uint8_t *st = img + (int64_t)ntiles * tile_bytes;
      for (int it = 0; it < 2 * ntiles + ncl; ++it) {
        const int s = it % NS;
        mbar_wait(bar_empty + 8 * s, ((it / NS) & 1) ^ 1);
        const uint8_t *src;
        uint32_t bytes = tile_bytes;
        if (it < ntiles) src = img + (int64_t)it * tile_bytes;
        else if (it < ntiles + ncl) {
          const int m0 = (it - ntiles) * cps;
          src = st + (int64_t)m0 * GC_CHUNK_BYTES;
          bytes = (uint32_t)min(cps, nchunks - m0) * GC_CHUNK_BYTES;
        } else src = img + (int64_t)(ntiles - 1 - (it - ntiles - ncl)) * tile_bytes;     // second pass: last tile first
        if (elect_one()) {
          mbar_expect_tx(bar_full + 8 * s, bytes);
          bulk_load(ring + s * tile_bytes, src, bytes, bar_full + 8 * s);
          GC_STAMP(1, it);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the schedule, one elected lane issues =====
    {
      constexpr uint32_t idesc16 = umma_idesc(128, 16), idesc32 = umma_idesc(128, 32);
      constexpr uint32_t idesc1 = idesc16 | (1u << 15);        // A is MN-major (pixels contiguous, K = channel rows)
      constexpr uint32_t idesc1_32 = idesc32 | (1u << 15);
      // This warp is a single serial instruction stream, so its instruction count per product IS the kernel's critical
      // path (measured: ~150 instructions per MMA pair cost more than the MMAs).  Descriptors are therefore built once
      // per operand and advanced by adding constants, the k loops are unrolled with compile-time offsets, and one elected
      // region issues all the products of a hand-off.
      const uint64_t dP = umma_desc_ls(obuf, 0, 1024);           // p operand, k-step ks at + (ks>>2)*4 KB + (ks&3)*32 B
      const int nks = c / 16;
      // phase 1: per tile pair (M = 128 pixels), per k-step (16 channels):  D[0:32) (+)= A_hi x [p_hi | p_lo]  as one
      // N = 32 product and  D[32:48) (+)= A_lo x p_hi  (N = 16); GC_PGROUP pairs per accumulator slot
      for (int tp = 0; tp < npairs; ++tp) {
        const int it0 = 2 * tp, it1 = it0 + 1;
        const int s0 = it0 % NS, s1 = it1 % NS;
        const int grp = tp / GC_PGROUP, as = grp & 1, sub = tp - grp * GC_PGROUP;
        if (sub == 0) {     // the drain warps have read this accumulator slot out of TMEM (tcgen05.ld) before arriving
          mbar_wait(bar_accfree + 8 * as, ((grp >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        mbar_wait(bar_full + 8 * s0, (it0 / NS) & 1);
        mbar_wait(bar_full + 8 * s1, (it1 / NS) & 1);
        if (lane == 0) GC_STAMP(2, tp);
        const int slo = min(s0, s1), shi = max(s0, s1);       // TMEM lanes 0-63 <- the tile in the lower slot
        const uint64_t dA = umma_desc_ls(ring + slo * tile_bytes, (uint32_t)(shi - slo) * tile_bytes, 1024);
        const uint64_t dAlo = dA + (uint64_t)(plane_bytes >> 4);
        const uint32_t tacc = tmem_base + as * (GC_P1_COLS * GC_PGROUP) + sub * GC_P1_COLS;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            if (ks < nks) {
              const uint64_t ka = (uint64_t)(ks * 2048 >> 4);                              // 16 channel rows = 2 KB
              const uint64_t kb = (uint64_t)(((ks >> 2) * GC_VTILE_BYTES + (ks & 3) * 32) >> 4);
              if (ks == 0) {
                umma_f16(tacc, dA, dP, idesc1_32, 0u);
                umma_f16(tacc + 32, dAlo, dP, idesc1, 0u);
              } else {
                umma_f16(tacc, dA + ka, dP + kb, idesc1_32, 1u);
                umma_f16(tacc + 32, dAlo + ka, dP + kb, idesc1, 1u);
              }
            }
          }
          umma_commit(bar_empty + 8 * s0);
          umma_commit(bar_empty + 8 * s1);
          if (sub == GC_PGROUP - 1 || tp == npairs - 1) umma_commit(bar_accfull + 8 * as);
        }
        __syncwarp();
        if (lane == 0) GC_STAMP(0, 2 + tp);
      }
      // phase 3: per tile  D[0:32) (+)= A_hi x [v_hi | v_lo],  D[32:48) (+)= A_lo x v_hi ;  M = 128 (c channel rows valid),
      // K = 64 pixels per tile, GC_FOLD tiles per accumulator slot, GC_VSTEP tiles per v-operand slot, each tile of a
      // slot with its own accumulator columns
      const int nvs = (ntiles + GC_VSTEP - 1) / GC_VSTEP;
      for (int vstep = 0; vstep < nvs; ++vstep) {
        const int vs = vstep & 1;
        const int j0 = vstep * GC_VSTEP;
        const int nt = min(GC_VSTEP, ntiles - j0);
        const int grp = j0 / GC_FOLD, fs = grp & 1;
        const bool first = (j0 % GC_FOLD) == 0;
        if (first) {
          mbar_wait(bar_foldfree + 8 * fs, ((grp >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        mbar_wait(bar_vready + 8 * vs, (vstep >> 1) & 1);
        uint64_t dX[GC_VSTEP];
        int sl[GC_VSTEP];
#pragma unroll
        for (int u = 0; u < GC_VSTEP; ++u) {
          const int it = ntiles + ncl + j0 + u;
          sl[u] = it % NS;
          if (u < nt) mbar_wait(bar_full + 8 * sl[u], (it / NS) & 1);
          dX[u] = umma_desc_ls(ring + sl[u] * tile_bytes, 0, 1024);
        }
        if (lane == 0) GC_STAMP(2, 16 + vstep);
        const uint64_t dV = umma_desc_ls(obuf + vs * GC_VSLOT_BYTES, 0, 1024);
        const uint64_t pl = (uint64_t)(plane_bytes >> 4);
        const uint32_t tacc = tmem_base + fs * (GC_P3_COLS * GC_VSTEP);
        const uint32_t acc0 = first ? 0u : 1u;
        if (elect_one()) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
#pragma unroll
            for (int u = 0; u < GC_VSTEP; ++u) {
              if (u < nt) {
                const uint64_t kk = (uint64_t)(k4 * 32 >> 4);
                const uint64_t dB = dV + (uint64_t)(u * GC_VTILE_BYTES >> 4) + kk;      // rows 0-15 hi taps, rows 16-31 lo taps
                umma_f16(tacc + u * GC_P3_COLS, dX[u] + kk, dB, idesc32, k4 == 0 ? acc0 : 1u);
                umma_f16(tacc + u * GC_P3_COLS + 32, dX[u] + pl + kk, dB, idesc16, k4 == 0 ? acc0 : 1u);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < GC_VSTEP; ++u)
            if (u < nt) umma_commit(bar_empty + 8 * sl[u]);
          umma_commit(bar_vfree + 8 * vs);
          const int jl = j0 + nt - 1;
          if ((jl % GC_FOLD) == GC_FOLD - 1 || jl == ntiles - 1) umma_commit(bar_foldfull + 8 * fs);
        }
        __syncwarp();
      }
    }
  } else {
    // ===== drain warps: thread dt owns TMEM lane dt =====
    const int q4 = warp & 3;
    const int dt = q4 * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const float yscale = 1.f / (GC_ACT_SCALE * pscale);
    // ---- phase 1 drain: scatter the 9 tap values of the own pixels into the padded score map, one tap per round so
    //      no two threads ever add to the same entry in a round (fixed order -> deterministic) ----
    const int ngrp1 = (npairs + GC_PGROUP - 1) / GC_PGROUP;
    for (int grp = 0; grp < ngrp1; ++grp) {
      const int as = grp & 1;
      const int np = min(GC_PGROUP, npairs - grp * GC_PGROUP);
      mbar_wait(bar_accfull + 8 * as, (grp >> 1) & 1);
      tc_fence_after();
      if (dt == 0) GC_STAMP(3, grp);
      float y[GC_PGROUP][16];
#pragma unroll
      for (int u = 0; u < GC_PGROUP; ++u) {
        if (u < np) {
          float y1[16], y2[16];
          const uint32_t col = tlane + as * (GC_P1_COLS * GC_PGROUP) + u * GC_P1_COLS;
          tmem_ld16(col, y[u]);
          tmem_ld16(col + 16, y1);
          tmem_ld16(col + 32, y2);
#pragma unroll
          for (int t = 0; t < 9; ++t) y[u][t] += y1[t] + y2[t];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_accfree + 8 * as);
      // The group's tap values go to shared memory as plain stores; then every score-map entry the group can reach
      // (its pixels +- one row and one column) is owned by exactly one thread, which gathers the up to nine contributions
      // and does ONE read-modify-write: two barriers per group instead of nine conflict-free scatter rounds.
      constexpr int GPX = GC_PGROUP * 128;
      const int g0 = grp * GPX;
#pragma unroll
      for (int u = 0; u < GC_PGROUP; ++u) {
        const int tp = grp * GC_PGROUP + u;
        const bool swapped = ((2 * tp) % NS) > ((2 * tp + 1) % NS);
        const int tile = 2 * tp + (((dt >> 6) & 1) ^ (swapped ? 1 : 0));
        const int li = (tile - 2 * grp * GC_PGROUP) * GC_TILE + (dt & 63);
        const bool on = u < np && g0 + li < hw;
#pragma unroll
        for (int t = 0; t < 9; ++t) ybuf[t * GPX + li] = on ? y[u][t] * yscale : 0.f;
      }
      drain_sync();
      const int tlo = max(g0 - (w + 1), 0), thi = min(g0 + GPX + w + 1, hw);      // targets [tlo, thi)
      for (int qt = tlo + dt; qt < thi; qt += 128) {
        const int ty = qt / w, tx = qt - ty * w;
        float sum = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int ddy = t / 3 - 1, ddx = t % 3 - 1;
          const int sy = ty + ddy, sx = tx + ddx;
          const int li = qt + ddy * w + ddx - g0;                   // source pixel, local to the group
          if (sy >= 0 && sy < h && sx >= 0 && sx < w && li >= 0 && li < GPX) sum += ybuf[t * GPX + li];
        }
        sp[(ty + 1) * wp + tx + 1] += sum;
      }
      drain_sync();
    }
    // ---- phase 2: v = sw (S s - use_y t) from the stencil chunks in the ring, and its maximum ----
    if (dt == 0) GC_STAMP(3, 15);
    float vmax = 0.f;
    for (int m = 0; m < nchunks; ++m) {
      const int it = ntiles + m / cps, s = it % NS;
      if (m % cps == 0) mbar_wait(bar_full + 8 * s, (it / NS) & 1);
      const float *ck = reinterpret_cast<const float *>(gen + (size_t)s * tile_bytes + (size_t)(m % cps) * GC_CHUNK_BYTES);
#pragma unroll
      for (int u = 0; u < GC_CHUNK_PX / 128; ++u) {
        const int pxl = dt + u * 128;
        const int q = m * GC_CHUNK_PX + pxl;
        if (q < hw) {
          const int py = q / w, px = q - py * w;
          float av = 0.f;
#pragma unroll
          for (int t = 0; t < 9; ++t) av = fmaf(ck[t * GC_CHUNK_PX + pxl], sp[(py + t / 3) * wp + px + t % 3], av);
          if (use_y) av -= ck[9 * GC_CHUNK_PX + pxl];
          av *= wgt;
          vp[(py + 1) * wp + px + 1] = av;
          vmax = fmaxf(vmax, fabsf(av));
        }
      }
      if (m % cps == cps - 1 || m == nchunks - 1) {
        drain_sync();
        if (dt == 0) mbar_arrive(bar_empty + 8 * s);
      }
    }
    // the operand buffer held p until now (every phase-1 product has completed): clear it for the v operand
    {
      uint4 *z = reinterpret_cast<uint4 *>(obuf_g) + dt * ((2 * GC_VSLOT_BYTES) / 16 / 128);
#pragma unroll
      for (int k = 0; k < (2 * GC_VSLOT_BYTES) / 16 / 128; ++k) z[k] = make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (lane == 0) red[q4] = vmax;
    drain_sync();
    const float vscale = pow2_scale(fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])));
    if (a.dbg != nullptr && use_y && blockIdx.x == 0 && blockIdx.y == 0) {
      for (int k = dt; k < 2 * npad; k += 128) a.dbg[k] = sp[k];
    }
    // ---- phase 3: build the shifted-v operand (thread = tap row, 8-pixel chunk of each tile), fold finished slots ----
    float gacc[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) gacc[u] = 0.f;
    auto fold = [&](int grp) {
      const int fs = grp & 1;
      mbar_wait(bar_foldfull + 8 * fs, (grp >> 1) & 1);
      tc_fence_after();
      float f[GC_VSTEP][32], f2[GC_VSTEP][16];
#pragma unroll
      for (int u = 0; u < GC_VSTEP; ++u) {
        const uint32_t col = tlane + fs * (GC_P3_COLS * GC_VSTEP) + u * GC_P3_COLS;
        tmem_ld32(col, f[u]);
        tmem_ld16(col + 32, f2[u]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_foldfree + 8 * fs);
#pragma unroll
      for (int u = 0; u < GC_VSTEP; ++u)
#pragma unroll
        for (int t = 0; t < 9; ++t) gacc[t] += f[u][t] + f[u][16 + t] + f2[u][t];
    };
    const int nvsteps = (ntiles + GC_VSTEP - 1) / GC_VSTEP;
    constexpr int STEPS_PER_FOLD = GC_FOLD / GC_VSTEP;
    if (dt == 0) GC_STAMP(3, 16);
    for (int vstep = 0; vstep < nvsteps; ++vstep) {
      const int vs = vstep & 1;
      mbar_wait(bar_vfree + 8 * vs, ((vstep >> 1) & 1) ^ 1);
      if (dt == 0) GC_STAMP(3, 17 + vstep);
      // 9 tap rows x GC_VSTEP tiles x 16 half-chunks of 4 pixels, spread over all 128 threads
      for (int item = dt; item < 9 * GC_VSTEP * 16; item += 128) {
        const int t = item / (GC_VSTEP * 16), rem = item - t * (GC_VSTEP * 16);
        const int u = rem >> 4, c8 = (rem >> 1) & 7, half = rem & 1;
        const int ddy = t / 3 - 1, ddx = t % 3 - 1;
        const int q0 = (ntiles - 1 - (vstep * GC_VSTEP + u)) * GC_TILE + c8 * 8 + half * 4;   // second pass runs backwards
        int py = q0 / w, px = q0 - py * w;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[e] = (q0 + e < hw) ? vp[(py + 1 - ddy) * wp + px + 1 - ddx] * vscale : 0.f;
          if (++px == w) { px = 0; ++py; }
        }
        const __half2 h01 = __floats2half2_rn(v[0], v[1]), h23 = __floats2half2_rn(v[2], v[3]);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(v[0] - f01.x, v[1] - f01.y), l23 = __floats2half2_rn(v[2] - f23.x, v[3] - f23.y);
        uint8_t *row = obuf_g + vs * GC_VSLOT_BYTES + u * GC_VTILE_BYTES + t * 128 + ((c8 ^ (t & 7)) << 4) + half * 8;
        uint2 ph, pl;
        ph.x = *reinterpret_cast<const uint32_t *>(&h01); ph.y = *reinterpret_cast<const uint32_t *>(&h23);
        pl.x = *reinterpret_cast<const uint32_t *>(&l01); pl.y = *reinterpret_cast<const uint32_t *>(&l23);
        *reinterpret_cast<uint2 *>(row) = ph;
        *reinterpret_cast<uint2 *>(row + 2048) = pl;
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_vready + 8 * vs);
      if (vstep > 0 && (vstep % STEPS_PER_FOLD) == 0) fold(vstep / STEPS_PER_FOLD - 1);
    }
    fold((ntiles - 1) / GC_FOLD);
    if (dt == 0) GC_STAMP(3, 40);
    if (dt < c) {
      const float gscale = 1.f / (GC_ACT_SCALE * vscale);
#pragma unroll
      for (int u = 0; u < 9; ++u) part[dt * 9 + u] = gacc[u] * gscale;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(GC_TMEM_COLS) : "memory");
  }
}

__global__ void __launch_bounds__(GC_THREADS, 2) gn_apply_tc_kernel(const GaArgs a, const GcParams P, const GcFuse F) {
  gc_sample(a, P);
  gc_fused_tail<GC_THREADS, 6>(a, F);
}

// n samples (c, hw) fp32 + their stencils -> operator images
__global__ void __launch_bounds__(256) build_images_kernel(const float *__restrict__ x, const float *__restrict__ stencil,
                                                           const float *__restrict__ uty, uint8_t *__restrict__ img, int c,
                                                           int hw, int64_t items_per_sample, int64_t image_bytes) {
  const int k = blockIdx.y;
  const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= items_per_sample) return;
  gc_image_item(x + (int64_t)k * c * hw, stencil + (int64_t)k * 9 * hw, uty + (int64_t)k * hw, img + (int64_t)k * image_bytes, c,
                hw, item);
}

static size_t gc_fixed_smem(int h, int w) {
  const size_t npad = (size_t)(h + 2) * (w + 2);
  return 1024 + 2 * GC_VSLOT_BYTES + 2 * npad * 4 + 9 * GC_PGROUP * 128 * 4 + 16 + 8 * GC_NBARS + 32 + 16;
}
// Ring depth: measured on B200 (3 objects x 69 samples at 30x54) two CTAs per SM with 3 slots each (0.37 ms per update)
// beat one CTA per SM with 8 slots (0.48 ms): a sample's phases are serialised by the role hand-offs, not by load
// latency, so a second resident sample hides more than a deeper ring does.
constexpr int GC_RING_SLOTS = 3;
static int gc_slots(int c, int h, int w) {
  const size_t tile = (size_t)2 * c * 128;
  const size_t fixed = gc_fixed_smem(h, w);
  if (fixed + 3 * tile > 227 * 1024) return 0;
  const int s = (int)((227 * 1024 - fixed) / tile);
  return s > GC_RING_SLOTS ? GC_RING_SLOTS : s;
}

bool gn_apply_tc_supported(int c, int h, int w) {
  // the stencil chunks travel through the tile ring, so a tile slot must hold one (c >= 40); p is staged in registers
  return c % 16 == 0 && c >= 48 && c * 9 <= 6 * GC_THREADS && c * 9 <= 1024 && c <= 128 && w < 65536 && gc_slots(c, h, w) >= 3;
}

int gn_apply_tc_launch(const GaArgs &a, const GcFuse &fuse, cudaStream_t st) {
  GcParams P;
  const int hw = a.h * a.w;
  P.ntiles = gc_ntiles(hw); P.nchunks = gc_nchunks(hw); P.tile_bytes = 2 * a.c * 128; P.slots = gc_slots(a.c, a.h, a.w);
  P.image_bytes = gc_sample_bytes(a.c, hw);
  const size_t smem = gc_fixed_smem(a.h, a.w) + (size_t)P.slots * P.tile_bytes;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(gn_apply_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("gn_apply_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return FRTM_ELAUNCH; }
    cudaFuncSetAttribute(gn_apply_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    configured = smem;
  }
  const dim3 grid(a.cap, a.table ? a.n_obj : 1);
  gn_apply_tc_kernel<<<grid, GC_THREADS, smem, st>>>(a, P, fuse);
  FRTM_CHECK_LAUNCH("gn_apply_tc");
  return FRTM_OK;
}

}  // namespace frtm

using namespace frtm;

extern "C" int64_t frtm_split_sample_bytes(int c, int hw) { return gc_sample_bytes(c, hw); }

extern "C" int frtm_split_samples(const float *samples, const float *stencil, const float *uty, int n, int c, int hw,
                                  void *split, void *stream) {
  FRTM_REQUIRE(samples && stencil && uty && split && n >= 1 && c % 8 == 0, "split_samples: bad arguments (c must be a multiple of 8)");
  FRTM_REQUIRE((reinterpret_cast<uintptr_t>(split) & 15) == 0, "split_samples: output must be 16-byte aligned");
  const int64_t items = gc_sample_items(c, hw);
  build_images_kernel<<<dim3(cdiv(items, 256), n), 256, 0, (cudaStream_t)stream>>>(samples, stencil, uty, (uint8_t *)split, c, hw,
                                                                                 items, gc_sample_bytes(c, hw));
  FRTM_CHECK_LAUNCH("split_samples");
  return FRTM_OK;
}
