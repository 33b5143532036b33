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
#define GM_T(k) do { if (tid == 0 && blockIdx.x == 0) { const long long t__ = clock64(); tacc[k] += t__ - tlast; tlast = t__; } } while (0)
#else
#define GM_T(k) do { } while (0)
#endif

// One unit of work: P3 blocks [pb0, pb1) of memory slot `i` of object `o` (a whole sample: [0, nblocks)); P1 runs two blocks
// further on either side (the window's lag).  The unit's gradient goes to `part` (n floats).
template <int C>
__device__ __forceinline__ void gm_sample(const GaArgs &a, const GcParams &P, int o, int i, int pb0, int pb1, float *part) {
  constexpr int KS = C / 16;                          // channel k-steps of P1 = channel m-tiles of P3
  constexpr int FOLD = 4;                             // P3 steps accumulated inside the tensor core between fp32 folds
  const int h = a.h, w = a.w, use_y = a.use_y;
  constexpr int n = C * 9;
  const int hw = h * w, lag = w + 1;
  const int tid = threadIdx.x, lane = tid & 31, wp = uniform_warp_idx();
#ifdef GM_TIMING
  const long long t_entry = clock64();
#endif
  const int nblocks = P.ntiles >> 1;
  const int fb = max(pb0 - 2, 0), lb = min(pb1 + 2, nblocks);          // P1 blocks [fb, lb)
  const uint32_t tile_bytes = (uint32_t)P.tile_bytes, blk_bytes = 2u * tile_bytes;
  const float *sw = a.sw, *pvec = a.pvec;
  const __half *xs = a.XS;
  if (a.table) {
    sw = reinterpret_cast<const float *>(a.table[3 * a.n_obj + o]);
    // RHS pass linearises at the filter itself, CG passes apply the operator to the direction p (= cg_state[0:n])
    pvec = reinterpret_cast<const float *>(a.table[(use_y ? 4 : 5) * a.n_obj + o]);
    xs = reinterpret_cast<const __half *>(a.table[7 * a.n_obj + o]);
  }
  const float wgt = sw[i];
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
    for (int b = 0; b < GM_SLOTS && fb + b < lb; ++b) {
      mbar_expect_tx(bars + 8 * b, blk_bytes);
      bulk_load(ring + b * blk_bytes, img + (int64_t)(fb + b) * blk_bytes, tile_bytes, bars + 8 * b);
      bulk_load(ring + b * blk_bytes + tile_bytes, img + (int64_t)(fb + b) * blk_bytes + tile_bytes, tile_bytes, bars + 8 * b);
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
    const unsigned q0 = (unsigned)(pb0 * GM_BLK + wp * 16 + k0);
    py = (int)__umulhi(q0, magic);
    px = (int)q0 - py * w;
  }

#ifdef GM_TIMING
  long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
  const long long t_loop = tlast;
#endif
  // scores are needed two rows beyond the unit's P3 pixels, residuals one row beyond
  const int s_end = min(GM_BLK * pb1 + 2 * lag, hw), v_end = min(GM_BLK * pb1 + lag, hw);
  int s_hi = max(GM_BLK * pb0 - 2 * lag, 0), v_hi = max(GM_BLK * pb0 - lag, 0);
  for (int b = fb; b < lb + 2; ++b) {
    const int s_lo = s_hi, v_lo = v_hi;
    s_hi = b >= lb - 1 ? s_end : min(s_end, max(GM_BLK * (b + 1) - lag, s_lo));
    v_hi = b >= lb - 1 ? v_end : min(v_end, max(s_hi - lag, v_lo));
    // stencil rows of this step's residual pixels: issued now, consumed after P1 and the score gather
    float st[10];
    const int qv = v_lo + tid;
    if (qv < v_hi) {
      const float *src = sten + (int64_t)(qv >> 8) * (10 * GC_CHUNK_PX) + (qv & (GC_CHUNK_PX - 1));
#pragma unroll
      for (int t = 0; t < 10; ++t) st[t] = (t < 9 || use_y) ? __ldg(src + t * GC_CHUNK_PX) : 0.f;
    }

    // ---------------- P1(b): five accumulator chains over the six channel k-steps ----------------
    if (b < lb) {
      const int slot = (b - fb) % GM_SLOTS;
      GM_T(0);
      mbar_wait(bars + 8 * slot, ((b - fb) / GM_SLOTS) & 1);
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
    if (b - 2 >= fb) {                                   // block b - 2 leaves the window (P3 on it if it belongs to the unit)
      const int bb = b - 2;
      const int slot = (bb - fb) % GM_SLOTS;
      if (bb >= pb0 && bb < pb1) {
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
      }
      GM_T(8);
      __syncthreads();                                 // every warp is done with block b - 2: its slot takes block b + 2
      GM_T(9);
      if (wp == 0 && b + 2 < lb) {                     // warp-uniform branch + elected lane: addresses stay in uniform registers
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
  const long long t_end = clock64();
  if (tid == 0 && blockIdx.x == 0)
    printf("gm prologue %lld | loop %lld clocks\n", t_loop - t_entry, t_end - t_loop);
  if (tid == 0 && blockIdx.x == 0)
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

// Units of a launch.  The U active samples of the work list run as whole samples, one CTA each (one CTA per SM, CTAs are
// dispatched in index order), except the R = U mod nsm samples of the last, partial wave: those are cut into k = min(4,
// nsm / R) parts each (P3 block ranges; the parts recompute two blocks of P1 on either side), so the last wave fills the
// machine with units a k-th as long instead of leaving nsm - R SMs idle for a whole sample time.
struct GmUnits {
  GnListWs ws;
  int nsm;
};
__device__ __forceinline__ int gm_unit_of_item(int t, int n_whole, int k) { return t < n_whole ? t : n_whole + (t - n_whole) * k; }

__global__ void __launch_bounds__(GM_THREADS, 1) gn_apply_mma_kernel(const GaArgs a, const GcParams P, const GcFuse F, const GmUnits Q) {
  constexpr int n = GM_MAXC * 9;
  __shared__ float red[32];
  __shared__ int s_last;
  const int nblocks = P.ntiles >> 1;
  const int U = Q.ws.list.hdr[0];
  const int R = U % Q.nsm;
  int k = R ? min(4, Q.nsm / R) : 1;
  k = max(1, min(k, nblocks / 2));                   // a part is at least two blocks
  const int n_whole = k > 1 ? U - R : U;
  const int units = n_whole + (U - n_whole) * k;
  const int b = blockIdx.x;
#ifdef GM_TIMING
  unsigned long long g0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
#endif
  if (b >= units) return;
  const int item_idx = b < n_whole ? b : n_whole + (b - n_whole) / k;
  const int part_idx = b < n_whole ? 0 : (b - n_whole) % k, parts = b < n_whole ? 1 : k;
  const uint32_t item = Q.ws.list.items[item_idx];
  const int o = (int)(item >> 16), slot = (int)(item & 0xffffu);
  float *row = Q.ws.rows + (int64_t)b * n;
  gm_sample<GM_MAXC>(a, P, o, slot, part_idx * nblocks / parts, (part_idx + 1) * nblocks / parts, row);
#ifdef GM_TIMING
  if (threadIdx.x == 0 && (b == 0 || b == 100 || b == 147 || b == 148 || b == 200 || b == units - 1)) {
    unsigned long long g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    printf("gm cta %d of %d units: start %llu end %llu (ns mod 1e6) = %llu ns\n", b, units, g0 % 1000000ull, g1 % 1000000ull, g1 - g0);
  }
#endif
  if (!F.enabled) return;

  // ---- reduction + CG vector step in the tail of the launch (as gc_fused_tail, over the object's unit rows) ----
  const int r0 = gm_unit_of_item(Q.ws.list.hdr[1 + o], n_whole, k), r1 = gm_unit_of_item(Q.ws.list.hdr[2 + o], n_whole, k);
  const int n_o = r1 - r0, ngrp = (n_o + GC_RGROUP - 1) / GC_RGROUP;
  const int grp = (b - r0) / GC_RGROUP, gsize = min(GC_RGROUP, n_o - grp * GC_RGROUP);
  int *cnt = Q.ws.counters + (int64_t)o * (1 + Q.ws.ngrp_max);
  float *gs = Q.ws.gsum + (int64_t)o * Q.ws.ngrp_max * n;
  __threadfence();                                   // this unit's row is visible before its ticket
  __syncthreads();
  if (threadIdx.x == 0) {
    const int ticket = atomicAdd(cnt + 1 + grp, 1);
    s_last = (ticket == gsize - 1) ? 1 : 0;
    if (s_last) cnt[1 + grp] = 0;                    // every ticket of this group has been drawn: reset for the next launch
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  {
    const float *part = Q.ws.rows + ((int64_t)r0 + (int64_t)grp * GC_RGROUP) * n;
    float *dst = gs + (int64_t)grp * n;
    for (int t = threadIdx.x; t < n; t += GM_THREADS) {
      float v[GC_RGROUP];
#pragma unroll
      for (int u = 0; u < GC_RGROUP; ++u) v[u] = u < gsize ? __ldcg(part + (int64_t)u * n + t) : 0.f;
      dst[t] = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int ticket = atomicAdd(cnt, 1);
    s_last = (ticket == ngrp - 1) ? 1 : 0;
    if (s_last) cnt[0] = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  CgVec cg = F.cg;
  const int *gate = F.gate;
  if (a.table) {
    float *cgst = reinterpret_cast<float *>(a.table[5 * a.n_obj + o]);
    cg.f = reinterpret_cast<float *>(a.table[4 * a.n_obj + o]);
    cg.p = cgst; cg.rprev = cgst + cg.n; cg.rho = cgst + 2 * cg.n; cg.hasp = cgst + 2 * cg.n + 1;
    cg.r += (int64_t)o * 3 * cg.n; cg.x += (int64_t)o * 3 * cg.n; cg.q += (int64_t)o * 3 * cg.n;
    gate = reinterpret_cast<const int *>(a.table[6 * a.n_obj + o]);
  }
  cg.partial = gs;                                   // the vector step sums the group rows
  cg.cap = ngrp;
  if (gate && gate[0] < F.min_px) return;
  cg_vector_step_cta<GM_THREADS, 4>(cg, F.mode, red);
}

static size_t gm_smem(int c) {
  return 1024 + (size_t)GM_SLOTS * 4 * c * 128 + (size_t)(9 * GM_YSTRIDE + 2 * GM_RING + 8) * 4 + 8 * GM_SLOTS + 64;
}

bool gn_apply_mma_supported(int c, int h, int w) {
  // the window lags are measured in rows of the map: three rows (+3) must fit in two 128-pixel blocks; the fragment
  // walk advances a pixel by up to 8 inside a row
  return c == GM_MAXC && w >= 8 && 3 * (w + 1) <= 2 * GM_BLK && h >= 1 && h * w < 65536 && gm_smem(c) <= 227 * 1024;
}

int gn_apply_mma_launch(const GaArgs &a, const GcFuse &fuse, const GnListWs &ws, cudaStream_t st) {
  GcParams P;
  const int hw = a.h * a.w;
  P.ntiles = gc_ntiles(hw); P.nchunks = gc_nchunks(hw); P.tile_bytes = 2 * a.c * 128; P.slots = GM_SLOTS;
  P.image_bytes = gc_sample_bytes(a.c, hw);
  const size_t smem = gm_smem(a.c);
  static int nsm = 0;
  if (nsm == 0) {
    cudaError_t e = cudaFuncSetAttribute(gn_apply_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("gn_apply_mma: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return FRTM_ELAUNCH; }
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    nsm = v < 160 ? v : 160;                           // the unit rows are sized for up to 160 extra part units
  }
  GmUnits Q;
  Q.ws = ws; Q.nsm = nsm;
  // at most (n_obj * cap) whole samples + nsm part units; CTAs beyond the launch's unit count exit at once
  gn_apply_mma_kernel<<<a.n_obj * a.cap + nsm, GM_THREADS, smem, st>>>(a, P, fuse, Q);
  FRTM_CHECK_LAUNCH("gn_apply_mma");
  return FRTM_OK;
}

}  // namespace frtm
