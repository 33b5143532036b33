// GN/CG operator  g_i = X_i^T [ sw_i (S_i (X_i * p) - use_y t_i) ]  over the frame memory in ONE pass over every sample.
//
// The operator is bound by the bytes of the samples (8.2 FLOP/B — above the fp32 ridge of the CUDA cores at HBM speed, so
// both contractions run as warp-level tensor-core tiles), and the sample does not fit in shared memory (622 KB at 480p),
// so the two-pass kernel (gn_apply_tc.cu) read it twice.  Here the three phases of a sample run as a sliding window over
// its operator image (target_model.cuh), 128 pixels (two image tiles, 48 KB at c = 96) per step:
//
//   step b:  P1(b)    tap maps   Y[tap][q] = sum_c X[c][q] p[c][tap]        q in block b          (mma.sync, A = X^T)
//            s        scores     s[q] = sum_tap Y[tap][q + off(tap)]        q <  128(b+1) -  (w+1)
//            v        residual   v[q] = sw (sum_tap S[tap][q] s[q + off] - use_y t[q])   q < 128(b+1) - 2(w+1)
//            P3(b-2)  gradient   g[c][tap] += sum_q X[c][q] v[q - off(tap)] q in block b-2       (mma.sync, A = X)
//
// P3 of a block needs v one row beyond the block, v needs s one row further and s needs Y one row further: 3(w+1) <= 256
// pixels, i.e. a lag of two blocks.  Blocks b-2 .. b stay resident in a 4-slot ring filled by bulk copies (one in flight
// while a step computes), so every byte of the image is read from HBM exactly once and the DRAM traffic of a launch is
// its algorithmic bytes.  All eight warps walk the steps together (four block barriers per step, no producer/consumer
// role hand-offs); in P1 warp k owns the k-th 16-pixel m-tile of the block, in P3 the k-th 16-pixel k-step.
//
// Arithmetic: operands are split fp16 pairs (16 x = hi + lo in the image; p scaled per object, v per warp-step by a power
// of two) and every product is hi*hi + hi*lo + lo*hi.  Taps 0-7 are one n-tile; tap 8 shares a second n-tile between its
// hi and lo columns, so a product costs five m16n8k16 instructions.  mma.sync accumulates at most 2 k-steps x 3 products
// before the result is folded into fp32 registers with ordinary adds; the 8 warps' partial gradients are summed in a fixed
// order, the per-sample rows by the fused tail (target_model.cuh) — results are deterministic.
#include "common.cuh"
#include "target_model.cuh"
#include "tc_ptx.cuh"

namespace frtm {

constexpr int GM_THREADS = 256;
constexpr int GM_WARPS = 8;
constexpr int GM_BLK = 128;               // pixels per step
constexpr int GM_SLOTS = 4;               // ring slots of one block each
constexpr int GM_YPX = 512;               // tap-map ring, pixels (4 blocks)
constexpr int GM_YSTRIDE = GM_YPX + 4;    // row stride in floats: rows two taps apart land 8 banks apart
constexpr int GM_RING = 1024;             // score / residual rings, pixels
constexpr int GM_MAXC = 96;

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

// Phase timeline (timing builds only, `make timing` -> libfrtm_b200_timing.so): thread 0 of CTA (0,0) accumulates the
// clock64 time it spends in every phase of a sample and prints the totals.
#ifdef GM_TIMING
#define GM_T(k) do { if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) { const long long t__ = clock64(); tacc[k] += t__ - tlast; tlast = t__; } } while (0)
#else
#define GM_T(k) do { } while (0)
#endif

template <int C>
__device__ __forceinline__ void gm_sample(const GaArgs &a, const GcParams &P) {
  constexpr int KS = C / 16;                          // channel k-steps of P1 = channel m-tiles of P3
  const int h = a.h, w = a.w, use_y = a.use_y;
  constexpr int n = C * 9;
  const int hw = h * w, lag = w + 1;
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  const int nblocks = P.ntiles >> 1;
  const uint32_t tile_bytes = (uint32_t)P.tile_bytes, blk_bytes = 2u * tile_bytes;
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
  const float wgt = sw[i];
  if (wgt == 0.f) {
    for (int k = tid; k < n; k += GM_THREADS) part[k] = 0.f;
    return;
  }
  const uint8_t *img = reinterpret_cast<const uint8_t *>(xs) + (int64_t)i * P.image_bytes;
  const float *sten = reinterpret_cast<const float *>(img + (int64_t)P.ntiles * tile_bytes);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t ring = base;
  float *ybuf = reinterpret_cast<float *>(gen + GM_SLOTS * blk_bytes);       // [9][GM_YSTRIDE]
  float *sring = ybuf + 9 * GM_YSTRIDE;
  float *vring = sring + GM_RING;
  float *red = vring + GM_RING;                                               // 8 floats
  const uint32_t bars = base + GM_SLOTS * blk_bytes + (9 * GM_YSTRIDE + 2 * GM_RING + 8) * 4;   // 8-byte aligned

  if (tid == 0) {
    for (int s = 0; s < GM_SLOTS; ++s) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int b = 0; b < GM_SLOTS && b < nblocks; ++b) {
      mbar_expect_tx(bars + 8 * b, blk_bytes);
      bulk_load(ring + b * blk_bytes, img + (int64_t)b * blk_bytes, tile_bytes, bars + 8 * b);
      bulk_load(ring + b * blk_bytes + tile_bytes, img + (int64_t)b * blk_bytes + tile_bytes, tile_bytes, bars + 8 * b);
    }
  }

  // ---- p: scaled to [2^9, 2^10), staged in shared memory (the tap-map ring is free until the first step), then every
  //      thread keeps its B fragments of all channel k-steps in registers ----
  float pscale;
  {
    float amax = 0.f;
    for (int k = tid; k < n; k += GM_THREADS) {
      const float v = pvec[k];
      ybuf[k] = v;
      amax = fmaxf(amax, fabsf(v));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0) red[wp] = amax;
    __syncthreads();                                   // also publishes the barrier initialisation
    amax = 0.f;
#pragma unroll
    for (int k = 0; k < GM_WARPS; ++k) amax = fmaxf(amax, red[k]);
    pscale = pow2_scale(amax);
  }
  const int g = lane >> 2, k0 = (lane & 3) * 2;        // fragment coordinates: row / column group, k pair
  uint32_t pbh[KS][2], pbl[KS][2], pb8[KS][2];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int ch = ks * 16 + u * 8 + k0;
      split2(ybuf[ch * 9 + g] * pscale, ybuf[(ch + 1) * 9 + g] * pscale, pbh[ks][u], pbl[ks][u]);
      uint32_t h8, l8;
      split2(ybuf[ch * 9 + 8] * pscale, ybuf[(ch + 1) * 9 + 8] * pscale, h8, l8);
      pb8[ks][u] = g == 0 ? h8 : (g == 1 ? l8 : 0u);   // tap 8: column 0 = hi, column 1 = lo
    }
  }
  __syncthreads();                                     // p staging is read: the tap-map ring may be written
  const float yscale = 1.f / (GC_ACT_SCALE * pscale);

  // per-lane ldmatrix offsets inside an image tile (row r of an 8x8 matrix = lane & 7, matrix id = lane >> 3)
  const int lr = lane & 7, lid = lane >> 3;
  const int jj = wp >> 2;                              // image tile of the block this warp's 16 pixels live in
  const int ch0 = (wp & 3) * 2;                        // their first 16-byte chunk in the tile row
  // P1: A = X^T (m = pixel, k = channel): matrices (k 0-7 | m 0-7), (k 0-7 | m 8-15), (k 8-15 | m 0-7), (k 8-15 | m 8-15)
  const uint32_t off1 = jj * tile_bytes + (uint32_t)((lid >> 1) * 8 + lr) * 128u + (uint32_t)(((ch0 + (lid & 1)) ^ lr) << 4);
  // P3: A = X (m = channel, k = pixel): matrices (m 0-7 | k 0-7), (m 8-15 | k 0-7), (m 0-7 | k 8-15), (m 8-15 | k 8-15)
  const uint32_t off3 = jj * tile_bytes + (uint32_t)((lid & 1) * 8 + lr) * 128u + (uint32_t)(((ch0 + (lid >> 1)) ^ lr) << 4);
  const uint32_t plane = (uint32_t)C * 128u;           // hi -> lo plane of a tile

  float gacc[KS][4], gacc8[KS][2];
#pragma unroll
  for (int m = 0; m < KS; ++m) {
#pragma unroll
    for (int u = 0; u < 4; ++u) gacc[m][u] = 0.f;
    gacc8[m][0] = 0.f; gacc8[m][1] = 0.f;
  }

#ifdef GM_TIMING
  long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
#endif
  int s_hi = 0, v_hi = 0;
  for (int b = 0; b < nblocks + 2; ++b) {
    const int s_lo = s_hi, v_lo = v_hi;
    s_hi = b >= nblocks - 1 ? hw : GM_BLK * (b + 1) - lag;
    v_hi = b >= nblocks - 1 ? hw : max(s_hi - lag, 0);
    // stencil rows of this step's residual pixels: issued now, consumed after P1 and the score gather
    float st[10];
    const int qv = v_lo + tid;
    if (tid < GM_BLK && qv < v_hi) {
      const float *src = sten + (int64_t)(qv >> 8) * (10 * GC_CHUNK_PX) + (qv & (GC_CHUNK_PX - 1));
#pragma unroll
      for (int t = 0; t < 10; ++t) st[t] = (t < 9 || use_y) ? __ldg(src + t * GC_CHUNK_PX) : 0.f;
    }

    // ---------------- P1(b) ----------------
    if (b < nblocks) {
      const int slot = b % GM_SLOTS;
      GM_T(0);
      mbar_wait(bars + 8 * slot, (b / GM_SLOTS) & 1);
      GM_T(1);
      const uint32_t t1 = ring + slot * blk_bytes + off1;
      float ys[4] = {0.f, 0.f, 0.f, 0.f}, y8[2] = {0.f, 0.f};
#pragma unroll
      for (int kp = 0; kp < KS; kp += 2) {
        float d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f}, d3[4] = {0.f, 0.f, 0.f, 0.f},
              d4[4] = {0.f, 0.f, 0.f, 0.f}, d5[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = kp; ks < kp + 2 && ks < KS; ++ks) {
          uint32_t ah[4], al[4];
          ldsm_x4_t(t1 + ks * 2048, ah);
          ldsm_x4_t(t1 + ks * 2048 + plane, al);
          hmma(d1, ah, pbh[ks][0], pbh[ks][1]);
          hmma(d2, ah, pbl[ks][0], pbl[ks][1]);
          hmma(d3, al, pbh[ks][0], pbh[ks][1]);
          hmma(d4, ah, pb8[ks][0], pb8[ks][1]);
          hmma(d5, al, pb8[ks][0], pb8[ks][1]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) ys[u] += (d1[u] + d2[u]) + d3[u];
        y8[0] += (d4[0] + d4[1]) + (d5[0] + d5[1]);
        y8[1] += (d4[2] + d4[3]) + (d5[2] + d5[3]);
      }
      // D rows = pixels g, g + 8 of the m-tile; columns = taps k0, k0 + 1 (tap 8: column 0/1 sum, lanes with k0 == 0)
      const int pr = ((b * GM_BLK) & (GM_YPX - 1)) + wp * 16 + g;
      ybuf[k0 * GM_YSTRIDE + pr] = ys[0] * yscale;
      ybuf[(k0 + 1) * GM_YSTRIDE + pr] = ys[1] * yscale;
      ybuf[k0 * GM_YSTRIDE + pr + 8] = ys[2] * yscale;
      ybuf[(k0 + 1) * GM_YSTRIDE + pr + 8] = ys[3] * yscale;
      if (k0 == 0) {
        ybuf[8 * GM_YSTRIDE + pr] = y8[0] * yscale;
        ybuf[8 * GM_YSTRIDE + pr + 8] = y8[1] * yscale;
      }
    }
    GM_T(2);
    __syncthreads();
    GM_T(3);

    // ---------------- scores ----------------
    for (int q = s_lo + tid; q < s_hi; q += GM_THREADS) {
      const int y = q / w, x = q - y * w;
      float sum = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int dy = t / 3 - 1, dx = t % 3 - 1;
        if ((unsigned)(y + dy) < (unsigned)h && (unsigned)(x + dx) < (unsigned)w)
          sum += ybuf[t * GM_YSTRIDE + ((q + dy * w + dx) & (GM_YPX - 1))];
      }
      sring[q & (GM_RING - 1)] = sum;
    }
    GM_T(4);
    __syncthreads();
    GM_T(5);

    // ---------------- residual ----------------
    for (int q = qv; q < v_hi; q += GM_BLK) {
      if (tid >= GM_BLK) break;
      if (q != qv) {                                   // tail steps cover more than one round: load directly
        const float *src = sten + (int64_t)(q >> 8) * (10 * GC_CHUNK_PX) + (q & (GC_CHUNK_PX - 1));
#pragma unroll
        for (int t = 0; t < 10; ++t) st[t] = (t < 9 || use_y) ? __ldg(src + t * GC_CHUNK_PX) : 0.f;
      }
      const int y = q / w, x = q - y * w;
      float av = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int dy = t / 3 - 1, dx = t % 3 - 1;
        const bool ok = (unsigned)(y + dy) < (unsigned)h && (unsigned)(x + dx) < (unsigned)w;
        av = fmaf(st[t], ok ? sring[(q + dy * w + dx) & (GM_RING - 1)] : 0.f, av);
      }
      if (use_y) av -= st[9];
      vring[q & (GM_RING - 1)] = av * wgt;
    }
    GM_T(6);
    __syncthreads();
    GM_T(7);

    // ---------------- P3(b - 2) ----------------
    if (b >= 2) {
      const int bb = b - 2;
      const int slot = bb % GM_SLOTS;
      const int qb = bb * GM_BLK + wp * 16;
      // B = shifted residual: B[k = pixel][n = tap] = v(y - dy, x - dx), zero outside the map
      float vv[4], v8[4];
      {
        const int q0 = qb + k0;
        int y = q0 / w, x = q0 - y * w;
        const int dy = g / 3 - 1, dx = g % 3 - 1;       // tap g (0-7)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          // k = k0, k0 + 1, k0 + 8, k0 + 9
          const int q = q0 + (e & 1) + (e >> 1) * 8;
          int yy = y, xx = x + (e & 1) + (e >> 1) * 8;
          if (xx >= w) { xx -= w; ++yy; }
          if (xx >= w) { xx -= w; ++yy; }
          const bool in = q < hw;
          const bool ok = in && (unsigned)(yy - dy) < (unsigned)h && (unsigned)(xx - dx) < (unsigned)w;
          vv[e] = ok ? vring[(q - dy * w - dx) & (GM_RING - 1)] : 0.f;
          const bool ok8 = in && g < 2 && yy >= 1 && xx >= 1;                  // tap 8: dy = dx = +1
          v8[e] = ok8 ? vring[(q - w - 1) & (GM_RING - 1)] : 0.f;
        }
      }
      float vmax = fmaxf(fmaxf(fabsf(vv[0]), fabsf(vv[1])), fmaxf(fabsf(vv[2]), fabsf(vv[3])));
      vmax = fmaxf(vmax, fmaxf(fmaxf(fabsf(v8[0]), fabsf(v8[1])), fmaxf(fabsf(v8[2]), fabsf(v8[3]))));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
      const float vscale = pow2_scale(vmax);
      const float gscale = 1.f / (GC_ACT_SCALE * vscale);
      uint32_t bh[2], bl[2], b8[2];
      split2(vv[0] * vscale, vv[1] * vscale, bh[0], bl[0]);
      split2(vv[2] * vscale, vv[3] * vscale, bh[1], bl[1]);
      {
        uint32_t h0, l0, h1, l1;
        split2(v8[0] * vscale, v8[1] * vscale, h0, l0);
        split2(v8[2] * vscale, v8[3] * vscale, h1, l1);
        b8[0] = g == 0 ? h0 : (g == 1 ? l0 : 0u);
        b8[1] = g == 0 ? h1 : (g == 1 ? l1 : 0u);
      }
      const uint32_t t3 = ring + slot * blk_bytes + off3;
#pragma unroll
      for (int m = 0; m < KS; ++m) {
        uint32_t ah[4], al[4];
        ldsm_x4(t3 + m * 2048, ah);
        ldsm_x4(t3 + m * 2048 + plane, al);
        float d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f}, d3[4] = {0.f, 0.f, 0.f, 0.f},
              d4[4] = {0.f, 0.f, 0.f, 0.f}, d5[4] = {0.f, 0.f, 0.f, 0.f};
        hmma(d1, ah, bh[0], bh[1]);
        hmma(d2, ah, bl[0], bl[1]);
        hmma(d3, al, bh[0], bh[1]);
        hmma(d4, ah, b8[0], b8[1]);
        hmma(d5, al, b8[0], b8[1]);
#pragma unroll
        for (int u = 0; u < 4; ++u) gacc[m][u] = fmaf((d1[u] + d2[u]) + d3[u], gscale, gacc[m][u]);
        gacc8[m][0] = fmaf((d4[0] + d4[1]) + (d5[0] + d5[1]), gscale, gacc8[m][0]);
        gacc8[m][1] = fmaf((d4[2] + d4[3]) + (d5[2] + d5[3]), gscale, gacc8[m][1]);
      }
      GM_T(8);
      __syncthreads();                                 // every warp is done with block b - 2: its slot takes block b + 2
      GM_T(9);
      if (tid == 0 && b + 2 < nblocks) {
        const int nb = b + 2;
        const uint32_t dst = ring + slot * blk_bytes;
        const uint8_t *src = img + (int64_t)nb * blk_bytes;
        mbar_expect_tx(bars + 8 * slot, blk_bytes);
        bulk_load(dst, src, tile_bytes, bars + 8 * slot);
        bulk_load(dst + tile_bytes, src + tile_bytes, tile_bytes, bars + 8 * slot);
      }
    }
  }

#ifdef GM_TIMING
  if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0)
    printf("gm timeline (clocks, %d steps): stencil-issue %lld | wait-load %lld | P1 %lld | bar %lld | scores %lld | bar %lld | residual %lld | bar %lld "
           "| P3 %lld | bar %lld\n", nblocks + 2, tacc[0], tacc[1], tacc[2], tacc[3], tacc[4], tacc[5], tacc[6], tacc[7], tacc[8], tacc[9]);
#endif
  // ---- the 8 warps' partial gradients, summed in a fixed order (the ring is idle: every block was consumed) ----
  float *gred = reinterpret_cast<float *>(gen);        // [GM_WARPS][n]
  {
    float *mine = gred + wp * n;
#pragma unroll
    for (int m = 0; m < KS; ++m) {
      const int ch = m * 16 + g;
      mine[ch * 9 + k0] = gacc[m][0];
      mine[ch * 9 + k0 + 1] = gacc[m][1];
      mine[(ch + 8) * 9 + k0] = gacc[m][2];
      mine[(ch + 8) * 9 + k0 + 1] = gacc[m][3];
      if (k0 == 0) {
        mine[ch * 9 + 8] = gacc8[m][0];
        mine[(ch + 8) * 9 + 8] = gacc8[m][1];
      }
    }
  }
  __syncthreads();
  for (int k = tid; k < n; k += GM_THREADS) {
    float v[GM_WARPS];
#pragma unroll
    for (int u = 0; u < GM_WARPS; ++u) v[u] = gred[u * n + k];
    part[k] = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
  }
}

__global__ void __launch_bounds__(GM_THREADS, 1) gn_apply_mma_kernel(const GaArgs a, const GcParams P, const GcFuse F) {
  gm_sample<GM_MAXC>(a, P);
  gc_fused_tail<GM_THREADS, 4>(a, F);
}

static size_t gm_smem(int c) {
  return 1024 + (size_t)GM_SLOTS * 4 * c * 128 + (size_t)(9 * GM_YSTRIDE + 2 * GM_RING + 8) * 4 + 8 * GM_SLOTS + 64;
}

bool gn_apply_mma_supported(int c, int h, int w) {
  // the window lags are measured in rows of the map: three rows (+3) must fit in two 128-pixel blocks; the fragment
  // walk advances a pixel by up to 8 inside a row
  return c == GM_MAXC && w >= 8 && 3 * (w + 1) <= 2 * GM_BLK && h >= 1 && gm_smem(c) <= 227 * 1024;
}

int gn_apply_mma_launch(const GaArgs &a, const GcFuse &fuse, cudaStream_t st) {
  GcParams P;
  const int hw = a.h * a.w;
  P.ntiles = gc_ntiles(hw); P.nchunks = gc_nchunks(hw); P.tile_bytes = 2 * a.c * 128; P.slots = GM_SLOTS;
  P.image_bytes = gc_sample_bytes(a.c, hw);
  const size_t smem = gm_smem(a.c);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gn_apply_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("gn_apply_mma: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return FRTM_ELAUNCH; }
    configured = true;
  }
  const dim3 grid(a.cap, a.table ? a.n_obj : 1);
  gn_apply_mma_kernel<<<grid, GM_THREADS, smem, st>>>(a, P, fuse);
  FRTM_CHECK_LAUNCH("gn_apply_mma");
  return FRTM_OK;
}

}  // namespace frtm
