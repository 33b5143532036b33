// GN/CG operator  g = sum_i X_i^T [ sw_i (S_i (X_i * p) - use_y t_i) ]  with every sample held ON CHIP by a thread-block
// cluster while its three phases run.
//
// A sample (c x h*w features as split fp16 planes, 622 KB at 480p) does not fit in one SM's shared memory, which forced
// the single-CTA kernels to stream it in 128-pixel steps with four block barriers per step (gn_apply_mma.cu) or to read it
// twice (gn_apply_tc.cu).  Here a cluster of CS CTAs (4 at 480p, 8 at 720p) owns one sample at a time: CTA r keeps the
// image tiles of its pixel range [a_r, b_r) (<= 8 tiles of 64 pixels, one bulk copy each) in its shared memory and runs
//
//   P1   tap maps   Y[tap][q] = sum_c X[c][q] p[c][tap]            q in [a_r, b_r)      mma.sync, A = X^T (ldmatrix.trans)
//   --   cluster barrier; the 9 x (w+1) tap values on either side of the range are copied from the neighbours (DSMEM)
//   s    scores     s[q] = sum_tap Y[tap][q + off(tap)]
//   --   cluster barrier; halo of s from the neighbours
//   v    residual   v[q] = sw (sum_tap S[tap][q] s[q + off] - use_y t[q])
//   --   cluster barrier; halo of v from the neighbours
//   P3   gradient   g[c][tap] += sum_q X[c][q] v[q - off(tap)]     q in [a_r, b_r)      mma.sync, A = X (ldmatrix)
//
// so the two tensor-core phases are long uninterrupted instruction streams (26 m-tiles / k-steps spread over 8 warps),
// three cluster barriers replace ~60 block barriers per sample, and every byte of the image is read from HBM exactly
// once.  The clusters are persistent: the active (object, slot) pairs are compacted once per update
// (gn_build_items_kernel), every cluster takes a contiguous range of them, and as soon as P3 has consumed a tile its
// slot is refilled with the same tile of the cluster's NEXT sample, so the copies of sample k+1 run behind the phases of
// sample k.  Gradients stay in registers across the samples of an object (fp32 sums, the tensor-core accumulators are
// folded at the end of every sample with that sample's operand scale); when the object changes or the range ends the 8
// warps, then the CS CTAs (through rank 0's shared memory), are summed in a fixed order and rank 0 writes ONE row per
// (object, cluster).  The cluster that delivers an object's last row sums the rows (fixed order) and runs the CG vector
// step, as the other operator kernels do: one launch per operator application, deterministic results.
//
// Arithmetic as in gn_apply_mma.cu: split fp16 operands (16 x = hi + lo in the image; p scaled per object, v per CTA and
// sample by a power of two), hi*hi + hi*lo + lo*hi, taps 0-7 in one n-tile and tap 8's hi | lo columns in a second one.
#include "common.cuh"
#include "target_model.cuh"
#include "tc_ptx.cuh"
#include "mma_sync.cuh"

namespace frtm {

// Phase timeline (timing builds only, `make timing`): thread 0 of ranks 0 and 1 of cluster 0 accumulates clock64 per phase.
#ifdef GM_TIMING
#define CL_T(k) do { if (tid == 0 && cid == 0 && rank < 2) { const long long t__ = clock64(); tacc[k] += t__ - tlast; tlast = t__; } } while (0)
#else
#define CL_T(k) do { } while (0)
#endif

constexpr int CL_THREADS = 256;
constexpr int CL_WARPS = 8;
constexpr int CL_C = 96;
constexpr int CL_MAXSLOT = 8;              // image tiles per CTA
constexpr int CL_HALO = 96;                // halo pixels kept on either side of the own range (>= w + 1, multiple of 32)
constexpr int CL_MAXITEMS = 96;            // samples per cluster per launch (own range of the compacted list)

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() { cluster_arrive(); cluster_wait(); }
// shared::cluster address of `addr` (a shared::cta address of this CTA's layout) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float dsmem_ld(uint32_t caddr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(caddr) : "memory");
  return v;
}
__device__ __forceinline__ void dsmem_st(uint32_t caddr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(caddr), "f"(v) : "memory");
}
// asynchronous store into a peer CTA's shared memory that signals the peer's mbarrier (complete_tx) — no fences, no
// cluster barrier: the halo exchange between neighbouring CTAs of a cluster is a push
__device__ __forceinline__ void dsmem_push(uint32_t caddr, float v, uint32_t cbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(caddr), "r"(__float_as_uint(v)),
               "r"(cbar)
               : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_arrive_cta(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

struct ClParams {
  int ntiles, tile_bytes, cs, nslot, nclusters, rowcap;
  int64_t image_bytes;
  ClList list;
  float *rows;        // [n_obj][rowcap][n]  one row per (object, cluster)
  int *tickets;       // [n_obj] zero-initialised; whoever draws the last ticket of an object resets it
};

__device__ __forceinline__ int cl_cluster_of(int u, int U, int NC) { return (int)(((long long)(u + 1) * NC - 1) / U); }

__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t &r0, uint32_t &r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}

__global__ void __launch_bounds__(CL_THREADS, 1) gn_apply_cl_kernel(const GaArgs a, const ClParams P, const GcFuse F) {
  constexpr int C = CL_C, KS = C / 16, n = C * 9;
  __shared__ int s_last;
  __shared__ float red32[32];
  const int h = a.h, w = a.w, use_y = a.use_y, hw = h * w, lag = w + 1;
  const int tid = threadIdx.x, lane = tid & 31, wp = uniform_warp_idx();
  const int rank = (int)cluster_ctarank(), cid = (int)cluster_id_x(), CS = P.cs;
  const uint32_t tile_bytes = (uint32_t)P.tile_bytes;
  // own tiles / pixels of every sample
  const int T0 = rank * P.ntiles / CS, T1 = (rank + 1) * P.ntiles / CS, ntl = T1 - T0;
  const int a0 = T0 * GC_TILE, own_px = max(min(T1 * GC_TILE, hw) - a0, 0);
  const int left_px = rank > 0 ? a0 - (rank - 1) * P.ntiles / CS * GC_TILE : 0;      // own pixels of the left neighbour
  const int npx = GC_TILE * ntl;                                                      // pixels of the own tiles (incl. padding)

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
  const int YS = 2 * CL_HALO + GC_TILE * P.nslot + 4;                   // row stride of the tap maps (== 4 mod 32)
  const int EXT = 2 * CL_HALO + GC_TILE * P.nslot;                      // s / v arrays with halo
  const int YCAP = max(9 * YS, 5 * n);                                  // the tap-map area doubles as p / flush staging (5 n floats)
  const uint32_t ring = base;
  // fixed block behind the ring (16-byte aligned): zero chunk | barriers | image pointers
  uint8_t *fix = gen + (size_t)P.nslot * tile_bytes;
  uint4 *zero16 = reinterpret_cast<uint4 *>(fix);                           // 16 zero bytes (ldmatrix rows of the unused taps)
  const uint32_t bars = smem_u32(fix + 16);
  const uint32_t bar_full = bars, bar_cons = bars + 8 * CL_MAXSLOT, bar_halo = bars + 16 * CL_MAXSLOT;   // + Y, s, v halo barriers
  unsigned long long *item_img = reinterpret_cast<unsigned long long *>(fix + 16 + 16 * CL_MAXSLOT + 32);   // [CL_MAXITEMS]
  // shifted-v operand of P3: fp16 [18 rows: taps 0-7 hi | taps 0-7 lo | tap 8 hi | tap 8 lo][VP] (also the p staging area)
  const int VP = GC_TILE * ntl + 8;                                      // row pitch in halves (pitch bytes == 16 mod 128)
  const int VCAP = max(18 * (GC_TILE * P.nslot + 8) * 2, n * 4);        // bytes
  __half *vhi = reinterpret_cast<__half *>(fix + 16 + 16 * CL_MAXSLOT + 32 + 8 * CL_MAXITEMS);
  __half *vlo = vhi + 8 * VP, *v8h = vlo + 8 * VP, *v8l = v8h + VP;
  float *ybuf = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(vhi) + VCAP);   // [9][YS] tap maps
  float *sext = ybuf + YCAP;
  float *vext = sext + EXT;
  float *red = vext + EXT;                                                   // 8 floats
  unsigned *vmaxbits = reinterpret_cast<unsigned *>(red + 8);
  int *issued = reinterpret_cast<int *>(red + 10);                          // [CL_MAXSLOT]
  float *item_wgt = red + 10 + CL_MAXSLOT;                                   // [CL_MAXITEMS] sample weights
  int *item_obj = reinterpret_cast<int *>(item_wgt + CL_MAXITEMS);          // [CL_MAXITEMS]
  uint16_t *maskv = reinterpret_cast<uint16_t *>(item_obj + CL_MAXITEMS);  // [64 * CL_MAXSLOT] tap validity of the own pixels
  const uint32_t ybuf_s = smem_u32(ybuf), sext_s = smem_u32(sext), vext_s = smem_u32(vext);
  const int nnb = (rank > 0 ? 1 : 0) + (rank + 1 < CS ? 1 : 0);         // neighbours that push halos into this CTA

  // ---- the cluster's range of the work list; everything an item needs from global memory is fetched up front ----
  // (fewer samples than clusters: the first U clusters take one each, so the clusters that hold rows of an object are
  //  always a contiguous interval; U == 0: no cluster has items and cl_cluster_of is never called)
  const int U = P.list.hdr[0];
  const int NC = min(P.nclusters, max(U, 1));
  const int u_lo = cid < NC ? (int)((long long)cid * U / NC) : 0, u_hi = cid < NC ? (int)((long long)(cid + 1) * U / NC) : 0;
  const int nitems = u_hi - u_lo;
  for (int k = tid; k < nitems; k += CL_THREADS) {
    const uint32_t item = P.list.items[u_lo + k];
    const int o = (int)(item >> 16), slot = (int)(item & 0xffffu);
    const __half *xs = a.table ? reinterpret_cast<const __half *>(a.table[7 * a.n_obj + o]) : a.XS;
    const float *swp = a.table ? reinterpret_cast<const float *>(a.table[3 * a.n_obj + o]) : a.sw;
    item_obj[k] = o;
    item_img[k] = (unsigned long long)(reinterpret_cast<const uint8_t *>(xs) + (int64_t)slot * P.image_bytes);
    item_wgt[k] = swp[slot];
  }
  const unsigned magic = 0xFFFFFFFFu / (unsigned)w + 1u;   // exact q / w for q < 65536
  for (int lp = tid; lp < GC_TILE * CL_MAXSLOT; lp += CL_THREADS) {
    // bit t: pixel q + off(t) is inside the map (t = 3 (dy + 1) + dx + 1); bit 15: q itself is a pixel of the map
    const int q = a0 + lp;
    unsigned m = 0u;
    if (lp < npx && q < hw) {
      const int y = (int)__umulhi((unsigned)q, magic), x = q - y * w;
      m = 0x8000u;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int dy = t / 3 - 1, dx = t % 3 - 1;
        if ((unsigned)(y + dy) < (unsigned)h && (unsigned)(x + dx) < (unsigned)w) m |= 1u << t;
      }
    }
    maskv[lp] = (uint16_t)m;
  }
  if (tid == 0) {
    for (int s = 0; s < CL_MAXSLOT; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_cons + 8 * s, 4); issued[s] = 0; }
    for (int s = 0; s < 3; ++s) mbar_init(bar_halo + 8 * s, 1);
    *zero16 = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // first sample: all own tiles
  if (wp == 0 && nitems > 0) {
    const uint8_t *img = reinterpret_cast<const uint8_t *>(item_img[0]) + (int64_t)T0 * tile_bytes;
    if (elect_one()) {
      for (int j = 0; j < ntl; ++j) {
        mbar_expect_tx(bar_full + 8 * j, tile_bytes);
        bulk_load(ring + j * tile_bytes, img + (int64_t)j * tile_bytes, tile_bytes, bar_full + 8 * j);
      }
    }
    __syncwarp();
  }
  cluster_sync_all();                                  // every CTA of the cluster is resident before any DSMEM access

  const int g = lane >> 2, k0 = (lane & 3) * 2;
  const int lr = lane & 7, lid = lane >> 3;
  const uint32_t plane = (uint32_t)C * 128u;

  float gsum[KS][4], gsum8[KS][2];
#pragma unroll
  for (int m = 0; m < KS; ++m) {
#pragma unroll
    for (int u = 0; u < 4; ++u) gsum[m][u] = 0.f;
    gsum8[m][0] = 0.f; gsum8[m][1] = 0.f;
  }
  uint32_t pbh[KS][2], pbl[KS][2], pb8[KS][2];
  float yscale = 0.f;
  int cur_obj = -1;

  // the row of (object o, this cluster) is complete in global memory: draw the object's ticket; the last cluster sums the
  // rows (fixed order) and runs the CG vector step.  Rank 0 only; CTA-local synchronisation only.
  auto finish_object = [&](int o, int c_lo, int nrows, bool have_row) {
    if (have_row) {
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        const int ticket = atomicAdd(P.tickets + o, 1);
        s_last = ticket == nrows - 1 ? 1 : 0;
        if (s_last) P.tickets[o] = 0;
      }
      __syncthreads();
      if (!s_last) return;
      __threadfence();
    }
    CgVec cg = F.cg;
    const int *gate = F.gate;
    if (a.table) {
      float *cgst = reinterpret_cast<float *>(a.table[5 * a.n_obj + o]);
      cg.f = reinterpret_cast<float *>(a.table[4 * a.n_obj + o]);
      cg.p = cgst; cg.rprev = cgst + cg.n; cg.rho = cgst + 2 * cg.n; cg.hasp = cgst + 2 * cg.n + 1;
      cg.r += (int64_t)o * 3 * cg.n; cg.x += (int64_t)o * 3 * cg.n; cg.q += (int64_t)o * 3 * cg.n;
      gate = reinterpret_cast<const int *>(a.table[6 * a.n_obj + o]);
    }
    cg.partial = P.rows + ((int64_t)o * P.rowcap + c_lo) * n;
    cg.cap = nrows;
    if (gate && gate[0] < F.min_px) return;
    cg_vector_step_cta<CL_THREADS, 4>(cg, F.mode, red32);
  };

  // Sum of the 8 warps, then of the CS CTAs (fixed orders) through the tap-map area; rank 0 writes the row of
  // (object, this cluster).  Called by every CTA of the cluster at the same point.
  auto flush = [&](int o) {
    float *stage = ybuf;                                 // [4][n]
    float *outv = ybuf + 4 * n;                          // [n]
    for (int round = 0; round < 2; ++round) {
      if ((wp >> 2) == round) {
        float *mine = stage + (wp & 3) * n;
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
      for (int k = tid; k < n; k += CL_THREADS) {
        const float sum4 = (stage[k] + stage[n + k]) + (stage[2 * n + k] + stage[3 * n + k]);
        outv[k] = round ? outv[k] + sum4 : sum4;
      }
      __syncthreads();
    }
#pragma unroll
    for (int m = 0; m < KS; ++m) {
#pragma unroll
      for (int u = 0; u < 4; ++u) gsum[m][u] = 0.f;
      gsum8[m][0] = 0.f; gsum8[m][1] = 0.f;
    }
    // cross-CTA: ranks 1.. deliver their vectors into rank 0's stage rows, four ranks per exchange
    for (int r0 = 1; r0 < CS; r0 += 4) {
      cluster_sync_all();                                // rank 0's stage rows are free
      if (rank >= r0 && rank < r0 + 4) {
        const uint32_t dst0 = dsmem_addr(ybuf_s + (uint32_t)((rank - r0) * n) * 4u, 0u);
        for (int k = tid; k < n; k += CL_THREADS) dsmem_st(dst0 + 4u * k, outv[k]);
      }
      cluster_sync_all();
      if (rank == 0) {
        const int cnt = min(4, CS - r0);
        for (int k = tid; k < n; k += CL_THREADS) {
          float acc = outv[k];
          for (int r = 0; r < cnt; ++r) acc += stage[r * n + k];
          outv[k] = acc;
        }
        __syncthreads();
      }
    }
    if (rank == 0) {
      float *row = P.rows + ((int64_t)o * P.rowcap + cid) * n;
      for (int k = tid; k < n; k += CL_THREADS) row[k] = outv[k];
    }
    // the staging area is the tap-map area, whose halo columns the neighbours push into as soon as they run the next
    // sample's P1: nobody proceeds before rank 0 has read its staging rows
    cluster_sync_all();
    if (rank != 0) return;
    const int f0 = P.list.hdr[1 + o], f1 = P.list.hdr[2 + o];
    const int c_lo = cl_cluster_of(f0, U, NC), c_hi = cl_cluster_of(f1 - 1, U, NC);
    finish_object(o, c_lo, c_hi - c_lo + 1, true);
  };

  // objects without a single active sample have no rows: their vector step runs on the regularisation terms alone
  if (cid == 0 && rank == 0) {
    for (int o = 0; o < a.n_obj; ++o)
      if (P.list.hdr[1 + o] == P.list.hdr[2 + o]) finish_object(o, 0, 0, false);
  }

#ifdef GM_TIMING
  long long tacc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
#endif
  for (int it = 0; it < nitems; ++it) {
    CL_T(0);
    const int o = item_obj[it];
    const uint32_t par = (uint32_t)it & 1u;
    const float wgt = item_wgt[it];
    const float *sten = reinterpret_cast<const float *>(reinterpret_cast<const uint8_t *>(item_img[it]) + (int64_t)P.ntiles * tile_bytes);
    const uint8_t *img_next = it + 1 < nitems ? reinterpret_cast<const uint8_t *>(item_img[it + 1]) + (int64_t)T0 * tile_bytes : nullptr;

    if (o != cur_obj) {
      // ---- new object: deliver the previous one, then p (scaled to [2^9, 2^10)) as B fragments in registers ----
      if (cur_obj >= 0) flush(cur_obj);
      cur_obj = o;
      const float *pvec = a.table ? reinterpret_cast<const float *>(a.table[(use_y ? 4 : 5) * a.n_obj + o]) : a.pvec;
      constexpr int PV = (n + CL_THREADS - 1) / CL_THREADS;
      float pv[PV];
#pragma unroll
      for (int k = 0; k < PV; ++k) pv[k] = tid + k * CL_THREADS < n ? pvec[tid + k * CL_THREADS] : 0.f;
      float amax = 0.f;
      float *pst = reinterpret_cast<float *>(vhi);       // staging in the operand area (free between samples)
#pragma unroll
      for (int k = 0; k < PV; ++k) {
        if (tid + k * CL_THREADS < n) pst[tid + k * CL_THREADS] = pv[k];
        amax = fmaxf(amax, fabsf(pv[k]));
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, d));
      if (lane == 0) red[wp] = amax;
      __syncthreads();
      amax = 0.f;
#pragma unroll
      for (int k = 0; k < CL_WARPS; ++k) amax = fmaxf(amax, red[k]);
      const float pscale = pow2_scale(amax);
      yscale = 1.f / (GC_ACT_SCALE * pscale);
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int ch = ks * 16 + u * 8 + k0;
          split2(pst[ch * 9 + g] * pscale, pst[(ch + 1) * 9 + g] * pscale, pbh[ks][u], pbl[ks][u]);
          uint32_t h8, l8;
          split2(pst[ch * 9 + 8] * pscale, pst[(ch + 1) * 9 + 8] * pscale, h8, l8);
          pb8[ks][u] = g == 0 ? h8 : (g == 1 ? l8 : 0u);
        }
      }
      __syncthreads();                                   // staging consumed before the operand rows are written again
    }
    if (tid == 0) *vmaxbits = 0u;
    CL_T(1);

    // ---------------- P1: m-tiles of 16 pixels, round-robin over the warps, in tile (= arrival) order ----------------
    for (int mt = wp; mt < 4 * ntl; mt += CL_WARPS) {
      const int j = mt >> 2, ch0 = (mt & 3) * 2;
      mbar_wait(bar_full + 8 * j, par);
      const uint32_t t1 = ring + j * tile_bytes + (uint32_t)((lid >> 1) * 8 + lr) * 128u + (uint32_t)(((ch0 + (lid & 1)) ^ lr) << 4);
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
      const int pr = CL_HALO + mt * 16 + g;
      ybuf[k0 * YS + pr] = ((d1[0] + d2[0]) + d3[0]) * yscale;
      ybuf[(k0 + 1) * YS + pr] = ((d1[1] + d2[1]) + d3[1]) * yscale;
      ybuf[k0 * YS + pr + 8] = ((d1[2] + d2[2]) + d3[2]) * yscale;
      ybuf[(k0 + 1) * YS + pr + 8] = ((d1[3] + d2[3]) + d3[3]) * yscale;
      if (k0 == 0) {
        ybuf[8 * YS + pr] = ((d4[0] + d4[1]) + (d5[0] + d5[1])) * yscale;
        ybuf[8 * YS + pr + 8] = ((d4[2] + d4[3]) + (d5[2] + d5[3])) * yscale;
      }
    }
    CL_T(2);
    __syncthreads();                                     // this CTA's tap maps are written
    // halo exchange #1: push the first / last lag pixels of the own tap maps into the neighbours' halo columns
    // (thread = (pixel k of the halo, tap parity); asynchronous stores that complete on the receiver's barrier)
    if (nnb) {
      if (tid == 0) mbar_expect_tx(bar_halo, (uint32_t)(nnb * 9 * lag * 4));
      const int k = tid & 127, t0 = tid >> 7;
      if (k < lag) {
        if (rank > 0) {
          const uint32_t dst = dsmem_addr(ybuf_s + (uint32_t)(CL_HALO + left_px + k) * 4u, rank - 1), bar = dsmem_addr(bar_halo, rank - 1);
#pragma unroll
          for (int u = 0; u < 5; ++u) { const int t = t0 + 2 * u; if (t < 9) dsmem_push(dst + (uint32_t)(t * YS) * 4u, ybuf[t * YS + CL_HALO + k], bar); }
        }
        if (rank + 1 < CS) {
          const uint32_t dst = dsmem_addr(ybuf_s + (uint32_t)(CL_HALO - lag + k) * 4u, rank + 1), bar = dsmem_addr(bar_halo, rank + 1);
#pragma unroll
          for (int u = 0; u < 5; ++u) { const int t = t0 + 2 * u; if (t < 9) dsmem_push(dst + (uint32_t)(t * YS) * 4u, ybuf[t * YS + CL_HALO + npx - lag + k], bar); }
        }
      }
    }
    // stencil rows of the own pixels, one thread per pixel (two rounds): issued now, consumed after the score phase
    constexpr int VR = (CL_MAXSLOT * GC_TILE) / CL_THREADS;
    float st[VR][10];
#pragma unroll
    for (int r = 0; r < VR; ++r) {
      const int lp = r * CL_THREADS + tid;
      const int q = a0 + lp;
      if (lp < own_px) {
        const float *src = sten + (int64_t)(q >> 8) * (10 * GC_CHUNK_PX) + (q & (GC_CHUNK_PX - 1));
#pragma unroll
        for (int t = 0; t < 10; ++t) st[r][t] = (t < 9 || use_y) ? __ldg(src + t * GC_CHUNK_PX) : 0.f;
      }
    }
    CL_T(3);
    if (nnb) mbar_wait(bar_halo, par);                   // the neighbours' tap values have landed in the halo columns
    CL_T(4);
    __syncthreads();
    CL_T(5);
    // ---------------- scores of the own pixels ----------------
    for (int lp = tid; lp < own_px; lp += CL_THREADS) {
      const unsigned m = maskv[lp];
      const float *yp = ybuf + CL_HALO + lp;
      float sum = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int dy = t / 3 - 1, dx = t % 3 - 1;
        const float yv = yp[t * YS + dy * w + dx];
        sum += (m >> t) & 1u ? yv : 0.f;
      }
      sext[CL_HALO + lp] = sum;
    }
    CL_T(6);
    __syncthreads();
    CL_T(7);
    if (nnb) {                                           // halo exchange #2: scores
      if (tid == 0) mbar_expect_tx(bar_halo + 8, (uint32_t)(nnb * lag * 4));
      if (tid < lag) {
        if (rank > 0) dsmem_push(dsmem_addr(sext_s + (uint32_t)(CL_HALO + left_px + tid) * 4u, rank - 1), sext[CL_HALO + tid], dsmem_addr(bar_halo + 8, rank - 1));
      } else if (tid >= 128 && tid - 128 < lag) {
        if (rank + 1 < CS) dsmem_push(dsmem_addr(sext_s + (uint32_t)(CL_HALO - lag + tid - 128) * 4u, rank + 1), sext[CL_HALO + npx - lag + tid - 128], dsmem_addr(bar_halo + 8, rank + 1));
      }
      mbar_wait(bar_halo + 8, par);
    }
    CL_T(8);
    // ---------------- residual of the own pixels, maximum of |v| ----------------
    {
      unsigned vb = 0u;
#pragma unroll
      for (int r = 0; r < VR; ++r) {
        const int lp = r * CL_THREADS + tid;
        if (lp < own_px) {
          const unsigned m = maskv[lp];
          const float *sp = sext + CL_HALO + lp;
          float av = 0.f;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const int dy = t / 3 - 1, dx = t % 3 - 1;
            const float sv = sp[dy * w + dx];
            av = fmaf(st[r][t], (m >> t) & 1u ? sv : 0.f, av);
          }
          if (use_y) av -= st[r][9];
          av *= wgt;
          vext[CL_HALO + lp] = av;
          vb = max(vb, __float_as_uint(fabsf(av)));
        }
      }
      vb = __reduce_max_sync(0xffffffffu, vb);
      if (lane == 0 && vb) atomicMax(vmaxbits, vb);
    }
    CL_T(9);
    __syncthreads();
    CL_T(10);
    if (nnb) {                                           // halo exchange #3: residual
      if (tid == 0) mbar_expect_tx(bar_halo + 16, (uint32_t)(nnb * lag * 4));
      if (tid < lag) {
        if (rank > 0) dsmem_push(dsmem_addr(vext_s + (uint32_t)(CL_HALO + left_px + tid) * 4u, rank - 1), vext[CL_HALO + tid], dsmem_addr(bar_halo + 16, rank - 1));
      } else if (tid >= 128 && tid - 128 < lag) {
        if (rank + 1 < CS) dsmem_push(dsmem_addr(vext_s + (uint32_t)(CL_HALO - lag + tid - 128) * 4u, rank + 1), vext[CL_HALO + npx - lag + tid - 128], dsmem_addr(bar_halo + 16, rank + 1));
      }
      mbar_wait(bar_halo + 16, par);
      // the operand scale covers the halo values too
      float hv = 0.f;
      if (tid < lag) { if (rank > 0) hv = vext[CL_HALO - lag + tid]; }
      else if (tid >= 128 && tid - 128 < lag) { if (rank + 1 < CS) hv = vext[CL_HALO + npx + tid - 128]; }
      const unsigned hb = __reduce_max_sync(0xffffffffu, __float_as_uint(fabsf(hv)));
      if (lane == 0 && hb) atomicMax(vmaxbits, hb);
    }
    __syncthreads();
    CL_T(11);

    // ---------------- the shifted-v operand of P3, built once for all warps: fp16 hi / lo rows [tap][pixel] ----------------
    const float vscale = scale_from_bits(*vmaxbits);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      // B[k = pixel][n = tap t] = v(y - dy, x - dx) = v[q - off(t)], inside the map iff tap 8 - t of q is; two pixels per store
      const int dy = t / 3 - 1, dx = t % 3 - 1;
      for (int pr = tid; pr < (npx >> 1); pr += CL_THREADS) {
        const int lp = pr * 2;
        const unsigned m0 = maskv[lp], m1 = maskv[lp + 1];
        const float *vp = vext + CL_HALO + lp - (dy * w + dx);
        const float x0 = ((m0 >> (8 - t)) & 1u) ? vp[0] * vscale : 0.f, x1 = ((m1 >> (8 - t)) & 1u) ? vp[1] * vscale : 0.f;
        uint32_t hh, ll;
        split2(x0, x1, hh, ll);
        if (t < 8) {
          *reinterpret_cast<uint32_t *>(vhi + t * VP + lp) = hh;
          *reinterpret_cast<uint32_t *>(vlo + t * VP + lp) = ll;
        } else {
          *reinterpret_cast<uint32_t *>(v8h + lp) = hh;
          *reinterpret_cast<uint32_t *>(v8l + lp) = ll;
        }
      }
    }
    __syncthreads();

    // ---------------- P3: k-steps of 16 pixels, round-robin over the warps; consumed tiles are refilled ----------------
    {
      float acc[KS][4], acc8[KS][4];
#pragma unroll
      for (int m = 0; m < KS; ++m) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { acc[m][u] = 0.f; acc8[m][u] = 0.f; }
      }
      // ldmatrix row addresses of the B operand: x4 = (hi | k 0-7), (hi | k 8-15), (lo | k 0-7), (lo | k 8-15), row = tap
      const uint32_t bsrc = smem_u32((lid < 2 ? vhi : vlo) + lr * VP) + (uint32_t)(lid & 1) * 16u;
      // x2 for tap 8: row 0 = hi, row 1 = lo, rows 2-7 = zeros
      const uint32_t b8src = lr == 0 ? smem_u32(v8h) + (uint32_t)(lid & 1) * 16u
                           : lr == 1 ? smem_u32(v8l) + (uint32_t)(lid & 1) * 16u : smem_u32(zero16);
      const uint32_t b8step = lr < 2 ? 32u : 0u;
      for (int ks = wp; ks < 4 * ntl; ks += CL_WARPS) {
        const int j = ks >> 2, ch0 = (ks & 3) * 2;
        const uint32_t t3 = ring + j * tile_bytes + (uint32_t)((lid & 1) * 8 + lr) * 128u + (uint32_t)(((ch0 + (lid >> 1)) ^ lr) << 4);
        uint32_t ah[KS][4], al[KS][4], bb[4], b8[2];
        ldsm_x4(bsrc + (uint32_t)ks * 32u, bb);
        ldsm_x2(b8src + (uint32_t)ks * b8step, b8[0], b8[1]);
#pragma unroll
        for (int m = 0; m < KS; ++m) {
          ldsm_x4(t3 + m * 2048, ah[m]);
          ldsm_x4(t3 + m * 2048 + plane, al[m]);
        }
#pragma unroll
        for (int m = 0; m < KS; ++m) hmma(acc[m], ah[m], bb[0], bb[1]);
#pragma unroll
        for (int m = 0; m < KS; ++m) hmma(acc8[m], ah[m], b8[0], b8[1]);
#pragma unroll
        for (int m = 0; m < KS; ++m) hmma(acc[m], ah[m], bb[2], bb[3]);
#pragma unroll
        for (int m = 0; m < KS; ++m) hmma(acc8[m], al[m], b8[0], b8[1]);
#pragma unroll
        for (int m = 0; m < KS; ++m) hmma(acc[m], al[m], bb[0], bb[1]);
        // this warp is done with tile j (its ldmatrix loads have returned: the products above consumed them); the warp that
        // completes the tile's count refills the slot with the same tile of the cluster's next sample
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cta(bar_cons + 8 * j);
          if (img_next != nullptr && mbar_test(bar_cons + 8 * j, par) && atomicExch(issued + j, it + 1) != it + 1) {
            mbar_expect_tx(bar_full + 8 * j, tile_bytes);
            bulk_load(ring + j * tile_bytes, img_next + (int64_t)j * tile_bytes, tile_bytes, bar_full + 8 * j);
          }
        }
        __syncwarp();
      }
      // fold this sample's accumulators (operand scale of this sample) into the fp32 sums
      const float gs = 1.f / (GC_ACT_SCALE * vscale);
#pragma unroll
      for (int m = 0; m < KS; ++m) {
#pragma unroll
        for (int u = 0; u < 4; ++u) gsum[m][u] = fmaf(acc[m][u], gs, gsum[m][u]);
        gsum8[m][0] = fmaf(acc8[m][0] + acc8[m][1], gs, gsum8[m][0]);
        gsum8[m][1] = fmaf(acc8[m][2] + acc8[m][3], gs, gsum8[m][1]);
      }
    }
    CL_T(12);
    // a tile whose last arrival did not see the completed phase (another warp's arrive raced its test) is refilled here;
    // the barrier also keeps the next sample's tap maps from overwriting the operand rows other warps still read
    __syncthreads();
    CL_T(13);
    if (wp == 0 && img_next != nullptr) {
      if (elect_one()) {
        for (int j = 0; j < ntl; ++j) {
          if (atomicExch(issued + j, it + 1) != it + 1) {
            mbar_expect_tx(bar_full + 8 * j, tile_bytes);
            bulk_load(ring + j * tile_bytes, img_next + (int64_t)j * tile_bytes, tile_bytes, bar_full + 8 * j);
          }
        }
      }
      __syncwarp();
    }
  }
  CL_T(14);
  if (cur_obj >= 0) flush(cur_obj);
  CL_T(15);
#ifdef GM_TIMING
  if (tid == 0 && cid == 0 && rank < 2)
    printf("cl timeline rank %d (clocks, %d items): top %lld | setup %lld | P1 %lld | sten-issue %lld | csync1 %lld | haloY %lld | scores %lld | "
           "csync2 %lld | haloS %lld | resid %lld | csync3 %lld | haloV %lld | Vbuild+P3 %lld | bar %lld | tail %lld | flush %lld\n", rank, nitems,
           tacc[0], tacc[1], tacc[2], tacc[3], tacc[4], tacc[5], tacc[6], tacc[7], tacc[8], tacc[9], tacc[10], tacc[11], tacc[12], tacc[13],
           tacc[14], tacc[15]);
#endif
  cluster_sync_all();                                    // no CTA leaves while a neighbour may still read its shared memory
}

// -------------------------------------------------------------------------------------------------------------------
static size_t cl_smem(int c, int nslot) {
  const size_t ys = 2 * CL_HALO + GC_TILE * nslot + 4, ext = 2 * CL_HALO + GC_TILE * nslot;
  const size_t ycap = 9 * ys > (size_t)5 * c * 9 ? 9 * ys : (size_t)5 * c * 9;
  const size_t vcap = (size_t)18 * (GC_TILE * nslot + 8) * 2 > (size_t)c * 9 * 4 ? (size_t)18 * (GC_TILE * nslot + 8) * 2 : (size_t)c * 9 * 4;
  // ring | zero chunk, barriers, image pointers | P3 operand | tap maps | s, v | red, vmax, issued | item weights, objects | masks
  return 1024 + (size_t)nslot * 2 * c * 128 + 16 + 16 * CL_MAXSLOT + 32 + 8 * CL_MAXITEMS + vcap +
         (ycap + 2 * ext + 10 + CL_MAXSLOT + 2 * CL_MAXITEMS) * 4 + 2 * GC_TILE * CL_MAXSLOT + 16;
}
// smallest cluster whose CTAs can hold their share of a sample (<= 8 tiles, shared memory) and whose ranges cover the
// neighbours' halos; 16 is the hardware maximum (non-portable size, opted into at launch)
static int cl_cluster_size(int c, int ntiles, int w) {
  for (int cs = 1; cs <= 16; cs *= 2) {
    const int per = (ntiles + cs - 1) / cs;
    if (per > CL_MAXSLOT || cl_smem(c, per) > 227 * 1024) continue;
    if (cs > 1 && ((ntiles / cs) < 1 || (ntiles / cs) * GC_TILE < w + 1)) continue;
    return cs;
  }
  return 0;
}

bool gn_apply_cl_supported(int c, int h, int w) {
  const int hw = h * w, ntiles = gc_ntiles(hw);
  if (c != CL_C || w < 8 || w + 1 > CL_HALO || hw >= 65536) return false;
  return cl_cluster_size(c, ntiles, w) != 0;
}

struct ClPlan {
  ClParams P;
  size_t smem;
};

static int cl_plan(const GaArgs &a, const GnListWs &ws, ClPlan &plan) {
  const int hw = a.h * a.w, n = a.c * 9;
  ClParams &P = plan.P;
  P.ntiles = gc_ntiles(hw); P.tile_bytes = 2 * a.c * 128; P.cs = cl_cluster_size(a.c, P.ntiles, a.w);
  P.nslot = (P.ntiles + P.cs - 1) / P.cs; P.image_bytes = gc_sample_bytes(a.c, hw); P.rowcap = 160;
  plan.smem = cl_smem(a.c, P.nslot);
  P.rows = ws.rows; P.tickets = ws.tickets; P.list = ws.list;
  (void)n;
  // persistent clusters: as many as the device can hold at once for this cluster size / shared-memory footprint
  static int cached_nc[17] = {0};
  static size_t cached_smem[17] = {0};
  if (cached_nc[P.cs] == 0 || cached_smem[P.cs] != plan.smem) {
    cudaError_t e = cudaFuncSetAttribute(gn_apply_cl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
    if (e == cudaSuccess && P.cs > 8) e = cudaFuncSetAttribute(gn_apply_cl_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) { set_error("gn_apply_cl: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return FRTM_ELAUNCH; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(P.cs * 64); cfg.blockDim = dim3(CL_THREADS); cfg.dynamicSmemBytes = plan.smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = P.cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nc = 0;
    e = cudaOccupancyMaxActiveClusters(&nc, gn_apply_cl_kernel, &cfg);
    if (e != cudaSuccess || nc <= 0) {
      cudaGetLastError();
      set_error("gn_apply_cl: cudaOccupancyMaxActiveClusters failed (%s)", cudaGetErrorString(e));
      return FRTM_ELAUNCH;
    }
    cached_nc[P.cs] = nc < 160 ? nc : 160;
    cached_smem[P.cs] = plan.smem;
  }
  P.nclusters = cached_nc[P.cs];
  return FRTM_OK;
}

int gn_apply_cl_launch(const GaArgs &a, const GcFuse &fuse, const GnListWs &ws, cudaStream_t st) {
  ClPlan plan;
  if (int rc = cl_plan(a, ws, plan)) return rc;
  FRTM_REQUIRE(((int64_t)a.n_obj * a.cap + plan.P.nclusters - 1) / plan.P.nclusters + 1 <= CL_MAXITEMS,
               "gn_apply_cl: %d x %d memory slots over %d clusters exceed the per-cluster work list", a.n_obj, a.cap, plan.P.nclusters);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(plan.P.cs * plan.P.nclusters); cfg.blockDim = dim3(CL_THREADS); cfg.dynamicSmemBytes = plan.smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = plan.P.cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gn_apply_cl_kernel, a, plan.P, fuse);
  if (e != cudaSuccess) { set_error("gn_apply_cl: launch failed: %s", cudaGetErrorString(e)); return FRTM_ELAUNCH; }
  FRTM_CHECK_LAUNCH("gn_apply_cl");
  return FRTM_OK;
}

}  // namespace frtm
