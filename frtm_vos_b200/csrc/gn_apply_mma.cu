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
// hi and lo columns, so a product costs five m16n8k16 instructions.  The warps are few (one CTA per SM, 8 warps), so the
// instruction streams are written for instruction-level parallelism: all ldmatrix loads of a phase are issued first, the
// products run as independent accumulator chains (P1: 5 chains over the 6 channel k-steps; P3: 12 chains, one per channel
// m-tile and tap group, continued across up to 4 steps inside the tensor core before they are folded into fp32 sums), the
// score / residual phases are branch-free.  The P3 operand is scaled by the running maximum of |v| of the sample (tracked
// by the residual phase with one warp-reduce + shared atomicMax per warp), which changes rarely; the accumulators are
// folded first when it does.  The 8 warps' partial gradients are summed in a fixed order, the per-sample rows by the fused
// tail (target_model.cuh) — results are deterministic.
#include "common.cuh"
#include "target_model.cuh"
#include "tc_ptx.cuh"
#include "mma_sync.cuh"

namespace frtm {

constexpr int GM_THREADS = 256;
constexpr int GM_WARPS = 8;
constexpr int GM_BLK = 128;               // pixels per step
constexpr int GM_SLOTS = 4;               // ring slots of one block each
constexpr int GM_YPX = 512;               // tap-map ring, pixels (4 blocks)
constexpr int GM_YSTRIDE = GM_YPX + 4;    // row stride in floats: rows two taps apart land 8 banks apart
constexpr int GM_RING = 1024;             // score / residual rings, pixels
constexpr int GM_MAXC = 96;

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
  constexpr int FOLD = 4;                             // P3 steps accumulated inside the tensor core between fp32 folds
  const int h = a.h, w = a.w, use_y = a.use_y;
  constexpr int n = C * 9;
  const int hw = h * w, lag = w + 1;
  const int tid = threadIdx.x, lane = tid & 31, wp = uniform_warp_idx();
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
  float *red = vring + GM_RING;                                               // 8 floats: [0..7] warp maxima, then the running |v| max
  unsigned *vmaxbits = reinterpret_cast<unsigned *>(red) + 7;
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
    constexpr int PV = (n + GM_THREADS - 1) / GM_THREADS;
    float pv[PV];
#pragma unroll
    for (int k = 0; k < PV; ++k) pv[k] = tid + k * GM_THREADS < n ? pvec[tid + k * GM_THREADS] : 0.f;   // all loads in flight
    float amax = 0.f;
#pragma unroll
    for (int k = 0; k < PV; ++k) {
      if (tid + k * GM_THREADS < n) ybuf[tid + k * GM_THREADS] = pv[k];
      amax = fmaxf(amax, fabsf(pv[k]));
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
  if (tid == 0) *vmaxbits = 0u;                        // (red[0..7] were consumed before the barrier)
  const float yscale = 1.f / (GC_ACT_SCALE * pscale);

  // exact q / w for q < 65536 without an integer division
  const unsigned magic = 0xFFFFFFFFu / (unsigned)w + 1u;
  // per-lane ldmatrix offsets inside an image tile (row r of an 8x8 matrix = lane & 7, matrix id = lane >> 3)
  const int lr = lane & 7, lid = lane >> 3;
  const int jj = wp >> 2;                              // image tile of the block this warp's 16 pixels live in
  const int ch0 = (wp & 3) * 2;                        // their first 16-byte chunk in the tile row
  // P1: A = X^T (m = pixel, k = channel): matrices (k 0-7 | m 0-7), (k 0-7 | m 8-15), (k 8-15 | m 0-7), (k 8-15 | m 8-15)
  const uint32_t off1 = jj * tile_bytes + (uint32_t)((lid >> 1) * 8 + lr) * 128u + (uint32_t)(((ch0 + (lid & 1)) ^ lr) << 4);
  // P3: A = X (m = channel, k = pixel): matrices (m 0-7 | k 0-7), (m 8-15 | k 0-7), (m 0-7 | k 8-15), (m 8-15 | k 8-15)
  const uint32_t off3 = jj * tile_bytes + (uint32_t)((lid & 1) * 8 + lr) * 128u + (uint32_t)(((ch0 + (lid >> 1)) ^ lr) << 4);
  const uint32_t plane = (uint32_t)C * 128u;           // hi -> lo plane of a tile

  // P3 state: accumulators inside the tensor core (acc, at scale cur_scale), folded sums in fp32 (gsum)
  float acc[KS][4], acc8[KS][4], gsum[KS][4], gsum8[KS][2];
#pragma unroll
  for (int m = 0; m < KS; ++m) {
#pragma unroll
    for (int u = 0; u < 4; ++u) { acc[m][u] = 0.f; acc8[m][u] = 0.f; gsum[m][u] = 0.f; }
    gsum8[m][0] = 0.f; gsum8[m][1] = 0.f;
  }
  float cur_scale = 0.f;
  int nacc = 0;
  auto fold = [&]() {
    const float gs = 1.f / (GC_ACT_SCALE * cur_scale);
#pragma unroll
    for (int m = 0; m < KS; ++m) {
#pragma unroll
      for (int u = 0; u < 4; ++u) { gsum[m][u] = fmaf(acc[m][u], gs, gsum[m][u]); acc[m][u] = 0.f; }
      gsum8[m][0] = fmaf(acc8[m][0] + acc8[m][1], gs, gsum8[m][0]);
      gsum8[m][1] = fmaf(acc8[m][2] + acc8[m][3], gs, gsum8[m][1]);
#pragma unroll
      for (int u = 0; u < 4; ++u) acc8[m][u] = 0.f;
    }
    nacc = 0;
  };
  // tap of this lane's B column (taps 0-7) and the position of its first k pixel, advanced by one block per P3 step
  const int tdy = g / 3 - 1, tdx = g % 3 - 1, toff = tdy * w + tdx;
  const int step_y = (int)__umulhi((unsigned)GM_BLK, magic), step_x = GM_BLK - step_y * w;
  int py, px;
  {
    const unsigned q0 = (unsigned)(wp * 16 + k0);
    py = (int)__umulhi(q0, magic);
    px = (int)q0 - py * w;
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
    if (qv < v_hi) {
      const float *src = sten + (int64_t)(qv >> 8) * (10 * GC_CHUNK_PX) + (qv & (GC_CHUNK_PX - 1));
#pragma unroll
      for (int t = 0; t < 10; ++t) st[t] = (t < 9 || use_y) ? __ldg(src + t * GC_CHUNK_PX) : 0.f;
    }

    // ---------------- P1(b): five accumulator chains over the six channel k-steps ----------------
    if (b < nblocks) {
      const int slot = b % GM_SLOTS;
      GM_T(0);
      mbar_wait(bars + 8 * slot, (b / GM_SLOTS) & 1);
      GM_T(1);
      const uint32_t t1 = ring + slot * blk_bytes + off1;
      uint32_t ah[KS][4], al[KS][4];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        ldsm_x4_t(t1 + ks * 2048, ah[ks]);
        ldsm_x4_t(t1 + ks * 2048 + plane, al[ks]);
      }
      float d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f}, d3[4] = {0.f, 0.f, 0.f, 0.f},
            d4[4] = {0.f, 0.f, 0.f, 0.f}, d5[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        hmma(d1, ah[ks], pbh[ks][0], pbh[ks][1]);
        hmma(d4, ah[ks], pb8[ks][0], pb8[ks][1]);
        hmma(d2, ah[ks], pbl[ks][0], pbl[ks][1]);
        hmma(d3, al[ks], pbh[ks][0], pbh[ks][1]);
        hmma(d5, al[ks], pb8[ks][0], pb8[ks][1]);
      }
      // D rows = pixels g, g + 8 of the m-tile; columns = taps k0, k0 + 1 (tap 8: column 0/1 sum, lanes with k0 == 0)
      const int pr = ((b * GM_BLK) & (GM_YPX - 1)) + wp * 16 + g;
      ybuf[k0 * GM_YSTRIDE + pr] = ((d1[0] + d2[0]) + d3[0]) * yscale;
      ybuf[(k0 + 1) * GM_YSTRIDE + pr] = ((d1[1] + d2[1]) + d3[1]) * yscale;
      ybuf[k0 * GM_YSTRIDE + pr + 8] = ((d1[2] + d2[2]) + d3[2]) * yscale;
      ybuf[(k0 + 1) * GM_YSTRIDE + pr + 8] = ((d1[3] + d2[3]) + d3[3]) * yscale;
      if (k0 == 0) {
        ybuf[8 * GM_YSTRIDE + pr] = ((d4[0] + d4[1]) + (d5[0] + d5[1])) * yscale;
        ybuf[8 * GM_YSTRIDE + pr + 8] = ((d4[2] + d4[3]) + (d5[2] + d5[3])) * yscale;
      }
    }
    GM_T(2);
    __syncthreads();
    GM_T(3);

    // ---------------- scores (branch-free gather; every step's range fits one round of the CTA) ----------------
    for (int q = s_lo + tid; q < s_hi; q += GM_THREADS) {
      const int y = (int)__umulhi((unsigned)q, magic), x = q - y * w;
      float sum = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int dy = t / 3 - 1, dx = t % 3 - 1;
        const bool ok = (unsigned)(y + dy) < (unsigned)h && (unsigned)(x + dx) < (unsigned)w;
        const float yv = ybuf[t * GM_YSTRIDE + ((q + dy * w + dx) & (GM_YPX - 1))];
        sum += ok ? yv : 0.f;
      }
      sring[q & (GM_RING - 1)] = sum;
    }
    GM_T(4);
    __syncthreads();
    GM_T(5);

    // ---------------- residual, and the running maximum of |v| (the scale of the P3 operand) ----------------
    {
      unsigned vb = 0u;
      for (int q0 = v_lo, round = 0; q0 < v_hi; q0 += GM_THREADS, ++round) {      // CTA-uniform trip count
        const int q = q0 + tid;
        if (q < v_hi) {
          if (round) {                                   // tail steps cover more than one round: load directly
            const float *src = sten + (int64_t)(q >> 8) * (10 * GC_CHUNK_PX) + (q & (GC_CHUNK_PX - 1));
#pragma unroll
            for (int t = 0; t < 10; ++t) st[t] = (t < 9 || use_y) ? __ldg(src + t * GC_CHUNK_PX) : 0.f;
          }
          const int y = (int)__umulhi((unsigned)q, magic), x = q - y * w;
          float av = 0.f;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const int dy = t / 3 - 1, dx = t % 3 - 1;
            const bool ok = (unsigned)(y + dy) < (unsigned)h && (unsigned)(x + dx) < (unsigned)w;
            const float sv = sring[(q + dy * w + dx) & (GM_RING - 1)];
            av = fmaf(st[t], ok ? sv : 0.f, av);
          }
          if (use_y) av -= st[9];
          av *= wgt;
          vring[q & (GM_RING - 1)] = av;
          vb = max(vb, __float_as_uint(fabsf(av)));
        }
      }
      vb = __reduce_max_sync(0xffffffffu, vb);
      if (lane == 0 && vb) atomicMax(vmaxbits, vb);
    }
    GM_T(6);
    __syncthreads();
    GM_T(7);

    // ---------------- P3(b - 2): six independent accumulator chains (channel m-tiles) x 5 products ----------------
    if (b >= 2) {
      const int bb = b - 2;
      const int slot = bb % GM_SLOTS;
      const uint32_t t3 = ring + slot * blk_bytes + off3;
      uint32_t ah[KS][4], al[KS][4];
#pragma unroll
      for (int m = 0; m < KS; ++m) {                       // the A operand does not depend on v: issue its loads first
        ldsm_x4(t3 + m * 2048, ah[m]);
        ldsm_x4(t3 + m * 2048 + plane, al[m]);
      }
      // scale of the operand: the running max of |v| over everything computed so far (changes rarely; the tensor-core
      // accumulators are folded into fp32 first when it does)
      const float vscale = scale_from_bits(*vmaxbits);
      if (vscale != cur_scale || nacc == FOLD) {
        if (nacc) fold();
        cur_scale = vscale;
      }
      // B = shifted residual: B[k = pixel][n = tap] = v(y - dy, x - dx), zero outside the map
      float vv[4], v8[4];
      {
        const int q0 = bb * GM_BLK + wp * 16 + k0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {                      // k = k0, k0 + 1, k0 + 8, k0 + 9
          const int dq = (e & 1) + (e >> 1) * 8;
          const int q = q0 + dq;
          int yy = py, xx = px + dq;
          if (xx >= w) { xx -= w; ++yy; }
          if (xx >= w) { xx -= w; ++yy; }
          const bool in = q < hw;
          const bool ok = in && (unsigned)(yy - tdy) < (unsigned)h && (unsigned)(xx - tdx) < (unsigned)w;
          const bool ok8 = in && g < 2 && yy >= 1 && xx >= 1;                  // tap 8: dy = dx = +1
          const float a0 = vring[(q - toff) & (GM_RING - 1)], a8 = vring[(q - w - 1) & (GM_RING - 1)];
          vv[e] = ok ? a0 * vscale : 0.f;
          v8[e] = ok8 ? a8 * vscale : 0.f;
        }
        px += step_x; py += step_y;                        // next P3 step: one block further
        if (px >= w) { px -= w; ++py; }
      }
      uint32_t bh[2], bl[2], b8[2];
      split2(vv[0], vv[1], bh[0], bl[0]);
      split2(vv[2], vv[3], bh[1], bl[1]);
      {
        uint32_t h0, l0, h1, l1;
        split2(v8[0], v8[1], h0, l0);
        split2(v8[2], v8[3], h1, l1);
        b8[0] = g == 0 ? h0 : (g == 1 ? l0 : 0u);
        b8[1] = g == 0 ? h1 : (g == 1 ? l1 : 0u);
      }
#pragma unroll
      for (int m = 0; m < KS; ++m) hmma(acc[m], ah[m], bh[0], bh[1]);
#pragma unroll
      for (int m = 0; m < KS; ++m) hmma(acc8[m], ah[m], b8[0], b8[1]);
#pragma unroll
      for (int m = 0; m < KS; ++m) hmma(acc[m], ah[m], bl[0], bl[1]);
#pragma unroll
      for (int m = 0; m < KS; ++m) hmma(acc8[m], al[m], b8[0], b8[1]);
#pragma unroll
      for (int m = 0; m < KS; ++m) hmma(acc[m], al[m], bh[0], bh[1]);
      ++nacc;
      GM_T(8);
      __syncthreads();                                 // every warp is done with block b - 2: its slot takes block b + 2
      GM_T(9);
      if (wp == 0 && b + 2 < nblocks) {                // warp-uniform branch + elected lane: addresses stay in uniform registers
        const int nb = b + 2;
        const uint32_t dst = ring + slot * blk_bytes;
        const uint8_t *src = img + (int64_t)nb * blk_bytes;
        if (elect_one()) {
          mbar_expect_tx(bars + 8 * slot, blk_bytes);
          bulk_load(dst, src, tile_bytes, bars + 8 * slot);
          bulk_load(dst + tile_bytes, src + tile_bytes, tile_bytes, bars + 8 * slot);
        }
        __syncwarp();
      }
    }
  }
  if (nacc) fold();

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
      mine[ch * 9 + k0] = gsum[m][0];
      mine[ch * 9 + k0 + 1] = gsum[m][1];
      mine[(ch + 8) * 9 + k0] = gsum[m][2];
      mine[(ch + 8) * 9 + k0 + 1] = gsum[m][3];
      if (k0 == 0) {
        mine[ch * 9 + 8] = gsum8[m][0];
        mine[(ch + 8) * 9 + 8] = gsum8[m][1];
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
  return c == GM_MAXC && w >= 8 && 3 * (w + 1) <= 2 * GM_BLK && h >= 1 && h * w < 65536 && gm_smem(c) <= 227 * 1024;
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
