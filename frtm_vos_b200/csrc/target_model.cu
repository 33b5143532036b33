// Target-model kernels: 3x3 correlation and its adjoints, hinge pixel weights, the stencil form of
// U^T diag(w^2) U, the frame memory, and the closed-form Gauss-Newton / Polak-Ribiere CG update.
//
// Math (model/discriminator.py:38-64, model/optimizer.py:77-157):
//   residual  r0 = W (U s(theta) - y),  W_i = pw_i sqrt(sw_i),  U = bilinear (h,w)->(H,W), align_corners=False
//   J^T J v   = Js^T [ sw_i * (U^T pw_i^2 U) ] Js v            -> 9-tap spatially varying stencil S_i on (h,w)
//   J^T r0    = Js^T [ sw_i * (S_i s_i - U^T pw_i^2 y_i) ]
// with Js the Jacobian of the low-resolution score map.  S_i and t_i = U^T pw_i^2 y_i are built once per memory
// sample at insert time, so the full-resolution maps are never touched inside CG.
#include "common.cuh"
#include "target_model.cuh"
#include <cstring>
#include <algorithm>

namespace frtm {

// ---------------------------------------------------------------- 3x3 correlation --------------------------------
// out[n][pix] = sum_c sum_tap x[n][c][pix+tap] f[fi(n)][c][tap]      (+= if accumulate)
__global__ void __launch_bounds__(512) corr3x3_kernel(const float *__restrict__ x, const float *__restrict__ filt,
                                                      const int *__restrict__ fidx, int c, int h, int w,
                                                      float *__restrict__ out, int accumulate,
                                                      const float *__restrict__ skip_if_zero) {
  // block = 128 pixels x 4 channel groups; partial sums are combined through shared memory in a fixed order
  extern __shared__ float fs[];  // [c][9] then [4][128] partials
  float *part = fs + c * 9;
  const int n = blockIdx.y;
  if (skip_if_zero && skip_if_zero[n] == 0.f) return;
  const float *f = filt + (int64_t)(fidx ? fidx[n] : 0) * c * 9;
  for (int i = threadIdx.x; i < c * 9; i += blockDim.x) fs[i] = f[i];
  __syncthreads();
  const int hw = h * w;
  const int lp = threadIdx.x & 127, cg = threadIdx.x >> 7;
  const int pix = blockIdx.x * 128 + lp;
  float acc = 0.f;
  if (pix < hw) {
    const int py = pix / w, px = pix - py * w;
    bool ok[9];
    int off[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
      ok[t] = yy >= 0 && yy < h && xx >= 0 && xx < w;
      off[t] = yy * w + xx;
    }
    const float *xn = x + (int64_t)n * c * hw;
    const int c0 = (c * cg) / 4, c1 = (c * (cg + 1)) / 4;
    for (int ch = c0; ch < c1; ++ch) {
      const float *xc = xn + (int64_t)ch * hw;
#pragma unroll
      for (int t = 0; t < 9; ++t)
        if (ok[t]) acc = fmaf(xc[off[t]], fs[ch * 9 + t], acc);
    }
  }
  part[cg * 128 + lp] = acc;
  __syncthreads();
  if (cg == 0 && pix < hw) {
    const float r = (part[lp] + part[128 + lp]) + (part[256 + lp] + part[384 + lp]);
    float *o = out + (int64_t)n * hw + pix;
    *o = accumulate ? *o + r : r;
  }
}

// v[n][pix] = sw[n] * ( sum_tap S[n][tap][pix] s[n][pix+tap] - use_y * t[n][pix] )
__global__ void stencil_apply_kernel(const float *__restrict__ S, const float *__restrict__ s, const float *__restrict__ t,
                                     const float *__restrict__ sw, int h, int w, int use_y, float *__restrict__ v) {
  const int n = blockIdx.y, hw = h * w;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= hw) return;
  const float wgt = sw[n];
  if (wgt == 0.f) { v[(int64_t)n * hw + pix] = 0.f; return; }
  const int py = pix / w, px = pix - py * w;
  const float *Sn = S + (int64_t)n * 9 * hw, *sn = s + (int64_t)n * hw;
  float acc = 0.f;
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) {
    const int yy = py + tp / 3 - 1, xx = px + tp % 3 - 1;
    if (yy >= 0 && yy < h && xx >= 0 && xx < w) acc = fmaf(Sn[(int64_t)tp * hw + pix], sn[yy * w + xx], acc);
  }
  if (use_y) acc -= t[(int64_t)n * hw + pix];
  v[(int64_t)n * hw + pix] = wgt * acc;
}

// partial[n][ch][tap] = sum_pix x[n][ch][pix+tap] v[n][pix]    one warp per channel, 8 channels per block
__global__ void __launch_bounds__(256) corr3x3_grad_filter_kernel(const float *__restrict__ x, const float *__restrict__ v,
                                                                  int c, int h, int w, float *__restrict__ partial,
                                                                  const float *__restrict__ skip_if_zero) {
  const int n = blockIdx.y;
  const int ch = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (ch >= c) return;
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.f;
  const int hw = h * w;
  if (!(skip_if_zero && skip_if_zero[n] == 0.f)) {
    const float *xc = x + ((int64_t)n * c + ch) * hw, *vn = v + (int64_t)n * hw;
    for (int pix = lane; pix < hw; pix += 32) {
      const float vv = vn[pix];
      const int py = pix / w, px = pix - py * w;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) acc[t] = fmaf(xc[yy * w + xx], vv, acc[t]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float s = warp_sum(acc[t]);
    if (lane == 0) partial[((int64_t)n * c + ch) * 9 + t] = s;
  }
}

// out[i] = sum_n partial[n][i]   (fixed order)
__global__ void reduce_rows_kernel(const float *__restrict__ partial, int rows, int64_t n, float *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += partial[(int64_t)r * n + i];
  out[i] = s;
}


// ---------------------------------------------------------------- fused J^T S J apply ----------------------------
// One CTA per memory sample computes that sample's contribution to  g = X^T [ sw (S (X * p) - use_y t) ]  in a single
// launch, streaming the sample (c x hw floats, NCHW) from HBM once and from L2 once:
//   phase 1  tap maps   Y[tap][q] = sum_c X[c][q] p[c][tap]      (thread = 2 consecutive pixels, 18 FMAs per 8-byte load)
//   phase 2  scores     s[q] = sum_tap Y[tap][q + tap];  v = sw (S s - use_y t)      (shared memory only)
//   phase 3  gradient   g[c][tap] = sum_q X[c][q] v[q - tap]     (warp = 4 channels, lane = pixel pair, 72 FMAs per
//                                                                 12 shared loads; one shuffle reduction at the end)
// Both passes are written around the SOURCE pixel q so every loaded element of X feeds 9 FMAs without touching its
// neighbours; the spatial shifts are applied to the small maps (Y, v) held in shared memory instead.
// Blackwell packed fp32 FMA (FFMA2): two independent IEEE fp32 FMAs per issue slot.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void ffma2(unsigned long long &d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

constexpr int GA_THREADS = 832;   // 26 warps: 810 pixel pairs of a 30x54 map in one round; 24 channel quads in phase 3

template <bool FAST>   // FAST: even width -> 8-byte loads and a shared 3x4 window per pixel pair
__global__ void __launch_bounds__(GA_THREADS, 1) gn_apply_kernel(const GaArgs a) {
  const float *__restrict__ X = a.X, *__restrict__ S = a.S, *__restrict__ T = a.T, *__restrict__ sw = a.sw,
                           *__restrict__ pvec = a.pvec;
  float *__restrict__ partial = a.partial;
  const int c = a.c, h = a.h, w = a.w, use_y = a.use_y;
  if (a.table) {
    const int o = blockIdx.y;
    X = reinterpret_cast<const float *>(a.table[0 * a.n_obj + o]);
    S = reinterpret_cast<const float *>(a.table[1 * a.n_obj + o]);
    T = reinterpret_cast<const float *>(a.table[2 * a.n_obj + o]);
    sw = reinterpret_cast<const float *>(a.table[3 * a.n_obj + o]);
    // RHS pass linearises at the filter itself, CG passes apply the operator to the direction p (= cg_state[0:n])
    pvec = reinterpret_cast<const float *>(a.table[(use_y ? 4 : 5) * a.n_obj + o]);
    partial += (int64_t)o * a.cap * c * 9;
  }
  extern __shared__ __align__(16) float sm[];
  const int hw = h * w, wp = w + 2, npad = (h + 2) * wp;
  float *Y = sm;                       // [9][hw]
  float *sp = Y + 9 * hw;              // padded scores
  float *vp = sp + npad;               // padded v
  float *ps = vp + npad;               // [c][12], read with 16-byte loads
  ps = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(ps) + 15) & ~(uintptr_t)15);
  const int i = blockIdx.x;
  const int n = c * 9;
  const float wgt = sw[i];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (wgt == 0.f) {
    for (int k = tid; k < n; k += GA_THREADS) partial[(int64_t)i * n + k] = 0.f;
    return;
  }
  for (int k = tid; k < c * 12; k += GA_THREADS) {
    const int ch = k / 12, t = k - ch * 12;
    ps[k] = t < 9 ? pvec[ch * 9 + t] : 0.f;
  }
  for (int k = tid; k < 2 * npad; k += GA_THREADS) sp[k] = 0.f;   // sp and vp are contiguous
  __syncthreads();
  const float *Xi = X + (int64_t)i * c * hw;
  const int npairs = (hw + 1) >> 1;  // FAST: w even, so a pixel pair never straddles two rows and hw is even
  const int stride2 = hw >> 1;
  auto load2 = [&](const float *chan_base, int g) -> float2 {
    if (FAST) return __ldg(reinterpret_cast<const float2 *>(chan_base) + g);
    float2 r;
    r.x = __ldg(chan_base + 2 * g);
    r.y = (2 * g + 1 < hw) ? __ldg(chan_base + 2 * g + 1) : 0.f;
    return r;
  };

  // ---- phase 1 ---- thread = pixel pair; taps are processed two at a time with packed FMAs (p pairs come straight out
  // of the 16-byte shared loads, the pixel value is duplicated into both halves); the next 4 channels are in flight
  for (int g = tid; g < npairs; g += GA_THREADS) {
    unsigned long long acc[2][5];
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int t = 0; t < 5; ++t) acc[k][t] = 0ull;
    float2 nx[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) nx[u] = load2(Xi + (int64_t)u * hw, g);
#pragma unroll 1
    for (int ch = 0; ch < c; ch += 4) {
      float2 x2[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) x2[u] = nx[u];
      if (ch + 4 < c) {
#pragma unroll
        for (int u = 0; u < 4; ++u) nx[u] = load2(Xi + (int64_t)(ch + 4 + u) * hw, g);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const ulonglong2 pa = *reinterpret_cast<const ulonglong2 *>(ps + (ch + u) * 12);       // (p0,p1) (p2,p3)
        const ulonglong2 pb = *reinterpret_cast<const ulonglong2 *>(ps + (ch + u) * 12 + 4);   // (p4,p5) (p6,p7)
        const unsigned long long pc = *reinterpret_cast<const unsigned long long *>(ps + (ch + u) * 12 + 8);  // (p8,0)
        const unsigned long long xa = pack2(x2[u].x, x2[u].x), xb = pack2(x2[u].y, x2[u].y);
        ffma2(acc[0][0], xa, pa.x); ffma2(acc[0][1], xa, pa.y); ffma2(acc[0][2], xa, pb.x); ffma2(acc[0][3], xa, pb.y);
        ffma2(acc[0][4], xa, pc);
        ffma2(acc[1][0], xb, pa.x); ffma2(acc[1][1], xb, pa.y); ffma2(acc[1][2], xb, pb.x); ffma2(acc[1][3], xb, pb.y);
        ffma2(acc[1][4], xb, pc);
      }
    }
    float y[2][10];
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int t = 0; t < 5; ++t) unpack2(acc[k][t], y[k][2 * t], y[k][2 * t + 1]);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      Y[t * hw + 2 * g] = y[0][t];
      if (2 * g + 1 < hw) Y[t * hw + 2 * g + 1] = y[1][t];
    }
  }
  __syncthreads();

  // ---- phase 2: s[q] = sum_tap Y[tap][q + tap], then v = sw (S s - use_y t) ----
  for (int q = tid; q < hw; q += GA_THREADS) {
    const int py = q / w, px = q - py * w;
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
      if (yy >= 0 && yy < h && xx >= 0 && xx < w) s += Y[t * hw + yy * w + xx];
    }
    sp[(py + 1) * wp + px + 1] = s;
  }
  __syncthreads();
  const float *Si = S + (int64_t)i * 9 * hw, *Ti = T + (int64_t)i * hw;
  for (int q = tid; q < hw; q += GA_THREADS) {
    const int py = q / w, px = q - py * w;
    float a = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) a = fmaf(Si[t * hw + q], sp[(py + t / 3) * wp + px + t % 3], a);
    if (use_y) a -= Ti[q];
    vp[(py + 1) * wp + px + 1] = wgt * a;
  }
  __syncthreads();

  // ---- phase 3: warp = channel quad, lane = pixel pair; the next pair's loads are in flight during the FMAs ----
  const int nquad = c >> 2;
  for (int quad = warp; quad < nquad; quad += GA_THREADS / 32) {
    float acc[4][9];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[u][t] = 0.f;
    const float *xo = Xi + (int64_t)quad * 4 * hw;
    float2 nx[4];
    if (lane < npairs) {
#pragma unroll
      for (int u = 0; u < 4; ++u) nx[u] = load2(xo + (int64_t)u * hw, lane);
    }
    int py = (2 * lane) / w, px = 2 * lane - py * w;   // advanced incrementally: +64 pixels per iteration
#pragma unroll 1
    for (int g = lane; g < npairs; g += 32) {
      float2 x2[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) x2[u] = nx[u];
      if (g + 32 < npairs) {
#pragma unroll
        for (int u = 0; u < 4; ++u) nx[u] = load2(xo + (int64_t)u * hw, g + 32);
      }
      const int q = 2 * g;
      if (g != lane) {
        px += 64;
        while (px >= w) { px -= w; ++py; }
      }
      if (FAST) {
        // v window of the pair (same row): rows py-1..py+1 (padded +1), columns px-1..px+2 (padded +1) -> 12 loads
        float vw[3][4];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) vw[r][cc] = vp[(py + r) * wp + px + cc];
        // tap t=(dy,dx): pixel k needs v[(py - dy, px + k - dx)] = vw[1 - dy][k + 1 - dx]
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int dy = t / 3 - 1, dx = t % 3 - 1;
          const float v0 = vw[1 - dy][1 - dx], v1 = vw[1 - dy][2 - dx];
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u][t] = fmaf(x2[u].x, v0, fmaf(x2[u].y, v1, acc[u][t]));
        }
      } else {
        const int q1 = q + 1;
        const int py1 = q1 / w, px1 = q1 - py1 * w;
        const bool has1 = q1 < hw;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int dy = t / 3 - 1, dx = t % 3 - 1;
          const float v0 = vp[(py + 1 - dy) * wp + px + 1 - dx];
          const float v1 = has1 ? vp[(py1 + 1 - dy) * wp + px1 + 1 - dx] : 0.f;
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u][t] = fmaf(x2[u].x, v0, fmaf(x2[u].y, v1, acc[u][t]));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float r = warp_sum(acc[u][t]);
        if (lane == 0) partial[(int64_t)i * n + (quad * 4 + u) * 9 + t] = r;
      }
  }
}

// ---------------------------------------------------------------- fused J^T S J apply, bulk-copy staged ----------
// Same three phases as gn_apply_kernel, but the sample is streamed into shared memory by a producer warp with
// cp.async.bulk (TMA 1-D bulk copies) through an mbarrier full/empty ring, so no load latency is ever exposed to the
// 26 compute warps and no registers are tied up by loads in flight:
//   phase 1 stages = CQ whole channels (contiguous CQ*hw floats, one bulk copy),  consumer thread = pixel pair
//   phase 3 stages = all c channels x 64 pixels (c row segments of 256 B),         consumer warp = channel quad
constexpr int GT_CONSUMER_WARPS = 26;
constexpr int GT_THREADS = (GT_CONSUMER_WARPS + 1) * 32;
constexpr int GT_SLAB = 64;           // pixels per phase-3 stage (one pixel pair per lane)

__device__ __forceinline__ uint32_t gt_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gt_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
  }
  printf("frtm gn_apply_tma: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
  __trap();
}

template <int CQ>
__global__ void __launch_bounds__(GT_THREADS, 1) gn_apply_tma_kernel(const GaArgs a, int NST, int stage_floats) {
  const float *__restrict__ X = a.X, *__restrict__ S = a.S, *__restrict__ T = a.T, *__restrict__ sw = a.sw,
                           *__restrict__ pvec = a.pvec;
  float *__restrict__ partial = a.partial;
  const int c = a.c, h = a.h, w = a.w, use_y = a.use_y;
  if (a.table) {
    const int o = blockIdx.y;
    X = reinterpret_cast<const float *>(a.table[0 * a.n_obj + o]);
    S = reinterpret_cast<const float *>(a.table[1 * a.n_obj + o]);
    T = reinterpret_cast<const float *>(a.table[2 * a.n_obj + o]);
    sw = reinterpret_cast<const float *>(a.table[3 * a.n_obj + o]);
    pvec = reinterpret_cast<const float *>(a.table[(use_y ? 4 : 5) * a.n_obj + o]);
    partial += (int64_t)o * a.cap * c * 9;
  }
  extern __shared__ __align__(128) float sm[];
  const int hw = h * w, wp = w + 2, npad = (h + 2) * wp;
  float *Y = sm;                                   // [9][hw]
  float *sp = Y + 9 * hw;                          // padded scores
  float *vp = sp + npad;                           // padded v
  float *ps = vp + npad;                           // [c][12], read with 16-byte loads
  ps = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(ps) + 15) & ~(uintptr_t)15);
  float *ring = ps + c * 12;
  ring = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(ring) + 127) & ~(uintptr_t)127);
  uint64_t *bars = reinterpret_cast<uint64_t *>(ring + (size_t)NST * stage_floats);   // full[NST] | empty[NST]
  const uint32_t bar_full = gt_smem_u32(bars), bar_empty = bar_full + 8 * NST;

  const int i = blockIdx.x;
  const int n = c * 9;
  const float wgt = sw[i];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (wgt == 0.f) {
    for (int k = tid; k < n; k += GT_THREADS) partial[(int64_t)i * n + k] = 0.f;
    return;
  }
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_full + 8 * s), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_empty + 8 * s), "r"(GT_CONSUMER_WARPS));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int k = tid; k < c * 12; k += GT_THREADS) {
    const int ch = k / 12, t = k - ch * 12;
    ps[k] = t < 9 ? pvec[ch * 9 + t] : 0.f;
  }
  for (int k = tid; k < 2 * npad; k += GT_THREADS) sp[k] = 0.f;
  __syncthreads();

  const float *Xi = X + (int64_t)i * c * hw;
  const int npairs = hw >> 1;
  const int nst1 = c / CQ;                              // phase-1 stages

  if (warp == GT_CONSUMER_WARPS) {
    // ===== producer warp =====
    // phase 1 only: large contiguous copies (CQ whole channels each).  Phase 3 wants all channels x few pixels, i.e.
    // c small row segments per stage; issued as separate bulk copies those are request-rate bound (measured 2.4x slower
    // than direct loads), so phase 3 reads its second pass straight from L2 with register prefetch instead.
    if (lane == 0) {
      for (int it = 0; it < nst1; ++it) {
        const int s = it % NST;
        const uint32_t ph = (it / NST) & 1;
        gt_mbar_wait(bar_empty + 8 * s, ph ^ 1);
        const uint32_t dst = gt_smem_u32(ring + (size_t)s * stage_floats);
        const uint32_t bytes = (uint32_t)CQ * hw * 4;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_full + 8 * s), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(reinterpret_cast<uint64_t>(Xi + (int64_t)it * CQ * hw)), "r"(bytes), "r"(bar_full + 8 * s) : "memory");
      }
    }
  } else {
    // ===== compute warps =====
    // ---- phase 1: thread = pixel pair, stage = CQ channels ----
    const bool own = tid < npairs;                 // host guarantees npairs <= 832
    unsigned long long acc[2][5];
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int t = 0; t < 5; ++t) acc[k][t] = 0ull;
    for (int it = 0; it < nst1; ++it) {
      const int s = it % NST;
      if (lane == 0) gt_mbar_wait(bar_full + 8 * s, (it / NST) & 1);   // one poller per warp
      __syncwarp();
      const float *st = ring + (size_t)s * stage_floats;
      if (own) {
        const float *xs = st + 2 * tid;
        const float *pq = ps + it * CQ * 12;
#pragma unroll
        for (int u = 0; u < CQ; ++u) {
          const float2 x2 = *reinterpret_cast<const float2 *>(xs + (size_t)u * hw);
          const float *pp = pq + u * 12;
          const ulonglong2 pa = *reinterpret_cast<const ulonglong2 *>(pp);
          const ulonglong2 pb = *reinterpret_cast<const ulonglong2 *>(pp + 4);
          const unsigned long long pc = *reinterpret_cast<const unsigned long long *>(pp + 8);
          const unsigned long long xa = pack2(x2.x, x2.x), xb = pack2(x2.y, x2.y);
          ffma2(acc[0][0], xa, pa.x); ffma2(acc[0][1], xa, pa.y); ffma2(acc[0][2], xa, pb.x); ffma2(acc[0][3], xa, pb.y);
          ffma2(acc[0][4], xa, pc);
          ffma2(acc[1][0], xb, pa.x); ffma2(acc[1][1], xb, pa.y); ffma2(acc[1][2], xb, pb.x); ffma2(acc[1][3], xb, pb.y);
          ffma2(acc[1][4], xb, pc);
        }
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_empty + 8 * s) : "memory");
    }
    if (own) {
      float y[2][10];
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int t = 0; t < 5; ++t) unpack2(acc[k][t], y[k][2 * t], y[k][2 * t + 1]);
#pragma unroll
      for (int t = 0; t < 9; ++t) *reinterpret_cast<float2 *>(Y + t * hw + 2 * tid) = make_float2(y[0][t], y[1][t]);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(GT_CONSUMER_WARPS * 32) : "memory");

    // ---- phase 2 ----
    for (int q = tid; q < hw; q += GT_CONSUMER_WARPS * 32) {
      const int py = q / w, px = q - py * w;
      float sv = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) sv += Y[t * hw + yy * w + xx];
      }
      sp[(py + 1) * wp + px + 1] = sv;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(GT_CONSUMER_WARPS * 32) : "memory");
    const float *Si = S + (int64_t)i * 9 * hw, *Ti = T + (int64_t)i * hw;
    for (int q = tid; q < hw; q += GT_CONSUMER_WARPS * 32) {
      const int py = q / w, px = q - py * w;
      float av = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) av = fmaf(Si[t * hw + q], sp[(py + t / 3) * wp + px + t % 3], av);
      if (use_y) av -= Ti[q];
      vp[(py + 1) * wp + px + 1] = wgt * av;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(GT_CONSUMER_WARPS * 32) : "memory");

    // ---- phase 3: warp = channel quad, lane = pixel pair, second pass over the sample straight from L2 ----
    const int nquad = c >> 2;
    for (int quad = warp; quad < nquad; quad += GT_CONSUMER_WARPS) {
      float g[4][9];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int t = 0; t < 9; ++t) g[u][t] = 0.f;
      const float2 *xo = reinterpret_cast<const float2 *>(Xi + (int64_t)quad * 4 * hw) + lane;
      const int stride2 = hw >> 1;
      float2 nx[4];
      if (lane < npairs) {
#pragma unroll
        for (int u = 0; u < 4; ++u) nx[u] = __ldg(xo + (int64_t)u * stride2);
      }
      int py = (2 * lane) / w, px = 2 * lane - py * w;
#pragma unroll 1
      for (int gp = lane; gp < npairs; gp += 32) {
        float2 x2[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x2[u] = nx[u];
        xo += 32;
        if (gp + 32 < npairs) {
#pragma unroll
          for (int u = 0; u < 4; ++u) nx[u] = __ldg(xo + (int64_t)u * stride2);
        }
        float vw[3][4];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float2 v01 = *reinterpret_cast<const float2 *>(vp + (py + r) * wp + px);
          const float2 v23 = *reinterpret_cast<const float2 *>(vp + (py + r) * wp + px + 2);
          vw[r][0] = v01.x; vw[r][1] = v01.y; vw[r][2] = v23.x; vw[r][3] = v23.y;
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int dy = t / 3 - 1, dx = t % 3 - 1;
          const float v0 = vw[1 - dy][1 - dx], v1 = vw[1 - dy][2 - dx];
#pragma unroll
          for (int u = 0; u < 4; ++u) g[u][t] = fmaf(x2[u].x, v0, fmaf(x2[u].y, v1, g[u][t]));
        }
        px += 64;
        while (px >= w) { px -= w; ++py; }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float r = warp_sum(g[u][t]);
          if (lane == 0) partial[(int64_t)i * n + (quad * 4 + u) * 9 + t] = r;
        }
    }
  }
}

// ---------------------------------------------------------------- pixel weights ----------------------------------
__global__ void __launch_bounds__(1024) pixel_count_kernel(const float *__restrict__ y, int HW, int threshold,
                                                           float *__restrict__ px) {
  __shared__ float red[32];
  const int k = blockIdx.x;
  float s = 0.f;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const float v = y[(int64_t)k * HW + i];
    s += threshold ? (v > 0.5f ? 1.f : 0.f) : v;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) px[k] = s;
}

__global__ void pixel_weights_kernel(const float *__restrict__ y, const float *__restrict__ px, const int *__restrict__ ipx,
                                     int HW, float tf, int threshold, float *__restrict__ w) {
  const int k = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HW) return;
  const float cnt = ipx ? (float)ipx[k] : px[k];
  float af = cnt / (float)HW;                       // discriminator.py:127
  const float small = cnt < 10.f ? 1.f : 0.f;       // :131-132
  af = small * tf + (1.f - small) * af;
  const float big = af > tf ? 1.f : 0.f;            // :134-135
  const float tfe = big * af + (1.f - big) * tf;
  const float wf = tfe / af, wb = (1.f - tfe) / (1.f - af);
  float v = y[(int64_t)k * HW + i];
  if (threshold) v = v > 0.5f ? 1.f : 0.f;
  w[(int64_t)k * HW + i] = sqrtf(wf * v + wb * (1.f - v));
}

// ---------------------------------------------------------------- stencil build ----------------------------------
__device__ __forceinline__ float tent(int idx, int i0, int i1, float lam) {
  return (idx == i0 ? 1.f - lam : 0.f) + (idx == i1 ? lam : 0.f);
}

// One warp per low-res pixel a=(i,j): gathers the ~ (2H/h)x(2W/w) high-res window that a contributes to.  The bilinear
// weights are separable, so each lane keeps the column factors of its (at most two) window columns in registers and
// the row factors are computed once per window row: 12 FMAs and two loads per high-res pixel, no index arithmetic.
__global__ void __launch_bounds__(256) build_stencil_kernel(const float *__restrict__ pw, const float *__restrict__ y, int H,
                                                            int W, int h, int w, float *__restrict__ stencil,
                                                            float *__restrict__ uty) {
  const int k = blockIdx.y, hw = h * w;
  const int a = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (a >= hw) return;
  const int i = a / w, j = a - i * w;
  const float sh = (float)h / (float)H, sw = (float)w / (float)W;
  // conservative window: all Y whose source row can be in (i-1, i+1)
  int Y0 = (int)floorf(((float)i - 1.f + 0.5f) / sh - 0.5f) - 1, Y1 = (int)ceilf(((float)i + 1.f + 0.5f) / sh - 0.5f) + 1;
  int X0 = (int)floorf(((float)j - 1.f + 0.5f) / sw - 0.5f) - 1, X1 = (int)ceilf(((float)j + 1.f + 0.5f) / sw - 0.5f) + 1;
  Y0 = max(Y0, 0); X0 = max(X0, 0); Y1 = min(Y1, H - 1); X1 = min(X1, W - 1);
  float acc[9], accy = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.f;
  const float *pwk = pw + (int64_t)k * H * W, *yk = y + (int64_t)k * H * W;
  // The columns that really touch j (source x in (j-1, j+1)) are one interval of at most 2W/w <= 32 columns inside the
  // conservative window (~2W/w + 4): start the pass at the first of them, and the second pass — four live lanes repeating
  // all the row work — is needed only when the interval is wider than a warp.
  {
    const int X = X0 + lane;
    int x0, x1;
    float lx;
    bilinear_src(min(X, X1), sw, w, x0, x1, lx);
    const unsigned live = __ballot_sync(0xffffffffu, X <= X1 && tent(j, x0, x1, lx) != 0.f);
    if (live != 0u) X0 += __ffs(live) - 1;
  }
  for (int xb = 0; xb < X1 - X0 + 1; xb += 32) {
    const int X = X0 + xb + lane;
    float cx[3] = {0.f, 0.f, 0.f}, cxc = 0.f;
    if (X <= X1) {
      int x0, x1;
      float lx;
      bilinear_src(X, sw, w, x0, x1, lx);
      cxc = tent(j, x0, x1, lx);                           // U[(Y,X), (.,j)] column factor
#pragma unroll
      for (int d = 0; d < 3; ++d) cx[d] = cxc * tent(j + d - 1, x0, x1, lx);
    }
    if (__ballot_sync(0xffffffffu, cxc != 0.f) == 0u) continue;
    // rows in batches of BS_ROWS with all their loads in flight (one load pair per row and trip made the warp a chain of
    // ~40 memory round trips); the accumulation order over Y is unchanged
    constexpr int BS_ROWS = 8;
    for (int Yb = Y0; Yb <= Y1; Yb += BS_ROWS) {
      float pv[BS_ROWS], yv[BS_ROWS];
#pragma unroll
      for (int u = 0; u < BS_ROWS; ++u) {
        const bool in = Yb + u <= Y1 && cxc != 0.f;
        pv[u] = in ? __ldg(pwk + (int64_t)(Yb + u) * W + X) : 0.f;
        yv[u] = in ? __ldg(yk + (int64_t)(Yb + u) * W + X) : 0.f;
      }
      // the row factors are the same in every lane: lane u computes those of row Yb + u once, the loop broadcasts them
      float f_cyc, f_cy[3];
      {
        int y0, y1;
        float ly;
        bilinear_src(min(Yb + (lane & (BS_ROWS - 1)), Y1), sh, h, y0, y1, ly);
        f_cyc = tent(i, y0, y1, ly);
#pragma unroll
        for (int d = 0; d < 3; ++d) f_cy[d] = f_cyc * tent(i + d - 1, y0, y1, ly);
      }
#pragma unroll
      for (int u = 0; u < BS_ROWS; ++u) {
        const int Y = Yb + u;
        if (Y > Y1) break;                                   // warp-uniform
        const float cyc = __shfl_sync(0xffffffffu, f_cyc, u);
        if (cyc == 0.f) continue;                            // warp-uniform
        float cy[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) cy[d] = __shfl_sync(0xffffffffu, f_cy[d], u);
        if (cxc != 0.f) {
          const float p = pv[u];
          const float p2 = p * p;
          accy = fmaf(p2 * (cyc * cxc), yv[u], accy);
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const float r = p2 * cy[dy];
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) acc[dy * 3 + dx] = fmaf(r, cx[dx], acc[dy * 3 + dx]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float s = warp_sum(acc[t]);
    if (lane == 0) stencil[((int64_t)k * 9 + t) * hw + a] = s;
  }
  accy = warp_sum(accy);
  if (lane == 0) uty[(int64_t)k * hw + a] = accy;
}

// ---------------------------------------------------------------- memory -----------------------------------------
// state: {current_size, prev_replace_ind (-1 none), slot chosen now (-1 = skipped), inserts so far}
// one step of the replace-minimum policy with its sample-weight update (memory.py:65-92); one warp
__device__ __forceinline__ void memory_policy_step(float *__restrict__ sw, int cap, float lr, int *__restrict__ state, int lane) {
  // lanes stride over the capacity, reductions by shuffle (first-minimum tie break = lowest index)
  const int size = state[0], prev = state[1];
  __syncwarp();
  int r = 0;
  if (size == 0 || lr == 1.f) {
    for (int i = lane; i < cap; i += 32) sw[i] = (i == 0) ? 1.f : 0.f;
  } else {
    float best = INFINITY;
    int bi = 0x7fffffff;
    for (int i = lane; i < cap; i += 32) {
      const float v = sw[i];
      if (v < best) { best = v; bi = i; }          // ascending i within a lane keeps the first minimum
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    r = bi;                                        // like torch.min(sw, 0): first index of the minimum
    if (prev < 0) {
      for (int i = lane; i < cap; i += 32) sw[i] = (i == r) ? lr : sw[i] / (1.f - lr);
    } else {
      const float pv = sw[prev];
      __syncwarp();
      if (lane == 0) sw[r] = pv / (1.f - lr);
    }
  }
  __syncwarp();
  double tot = 0.0;
  for (int i = lane; i < cap; i += 32) tot += (double)sw[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  const float ft = (float)tot;
  for (int i = lane; i < cap; i += 32) sw[i] = sw[i] / ft;
  __syncwarp();
  if (lane == 0) {
    state[1] = r;
    state[2] = r;
    state[0] = min(size + 1, cap);
    state[3] += 1;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(32) memory_next_slot_kernel(float *__restrict__ sw, int cap, float lr, int *__restrict__ state,
                                                              const int *__restrict__ gate_count, int min_px) {
  const int lane = threadIdx.x;
  if (gate_count && gate_count[0] < min_px) { if (lane == 0) state[2] = -1; return; }
  memory_policy_step(sw, cap, lr, state, lane);
}

// The policy steps of all frames of a track block, one warp per object: the frames of an object are sequential (every
// insert changes the weights the next choice looks at), the objects are independent.  slots[f * n_obj + o] = slot of
// frame f (-1 = gated out).  table rows: 6 = sample weights, 7 = policy state (see MemBlockArgs).
__global__ void __launch_bounds__(32) memory_slots_block_kernel(const long long *__restrict__ table, int n_obj, int nF, int cap,
                                                                float lr, const int *__restrict__ gate_counts, int min_px,
                                                                int *__restrict__ slots) {
  const int o = blockIdx.x, lane = threadIdx.x;
  float *sw = reinterpret_cast<float *>(table[6 * n_obj + o]);
  int *state = reinterpret_cast<int *>(table[7 * n_obj + o]);
  for (int f = 0; f < nF; ++f) {
    if (gate_counts[f * n_obj + o] < min_px) {
      if (lane == 0) { state[2] = -1; slots[f * n_obj + o] = -1; }
      __syncwarp();
      continue;
    }
    memory_policy_step(sw, cap, lr, state, lane);
    if (lane == 0) slots[f * n_obj + o] = state[2];
    __syncwarp();
  }
}

// All five pieces of one sample (projected features, soft label, pixel weights, stencil, U^T w^2 y) in one launch, plus
// the split tile image of the features that the tensor-core operator kernel streams (gn_apply_tc.cu).
struct InsertArgs {
  const float *src[5];
  float *dst[5];
  int64_t n[5];
  int64_t total;
  uint8_t *split;         // [cap][gc_sample_bytes] operator images or null
  int c, hw;
  int64_t split_items;
};
__global__ void memory_insert_kernel(const InsertArgs a, const int *__restrict__ state) {
  const int slot = state[2];
  if (slot < 0) return;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.total) {
    i -= a.total;
    if (i < a.split_items)
      gc_image_item(a.src[0], a.src[3], a.src[4], a.split + (int64_t)slot * gc_sample_bytes(a.c, a.hw), a.c, a.hw, i);
    return;
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    if (i < a.n[k]) {
      a.dst[k][(int64_t)slot * a.n[k] + i] = a.src[k][i];
      return;
    }
    i -= a.n[k];
  }
}

// All inserts of a track block in one launch: blockIdx.y = f * n_obj + o copies sample (f, o) into the slot the policy
// kernel chose for it.  table rows (each n_obj pointers): samples, labels, pixel weights, stencil, uty, operator images.
// A slot chosen twice inside the block keeps the LAST frame's sample, as the sequential inserts would leave it.
struct MemBlockArgs {
  const long long *table;
  const int *slots;
  const float *src[5];      // features, labels, pixel weights, stencil, uty of all (frame, object) rows, row-major
  int64_t n[5];
  int64_t total, split_items;
  int n_obj, nF, c, hw;
};
template <int V>   // elements per thread of the copy part: 4 (every piece a multiple of 4 floats, 16-byte aligned) or 1
__global__ void __launch_bounds__(256) memory_insert_block_kernel(const MemBlockArgs a, int copy_blocks) {
  const int row = blockIdx.y, f = row / a.n_obj, o = row - f * a.n_obj;
  __shared__ int s_slot;
  if (threadIdx.x == 0) {
    int slot = a.slots[row];
    for (int g = f + 1; g < a.nF && slot >= 0; ++g)
      if (a.slots[g * a.n_obj + o] == slot) slot = -1;
    s_slot = slot;
  }
  __syncthreads();
  const int slot = s_slot;
  if (slot < 0) return;
  if ((int)blockIdx.x >= copy_blocks) {            // operator image of the sample
    const int64_t i = (int64_t)(blockIdx.x - copy_blocks) * blockDim.x + threadIdx.x;
    uint8_t *split = reinterpret_cast<uint8_t *>(a.table[5 * a.n_obj + o]);
    if (i < a.split_items && split)
      gc_image_item(a.src[0] + (int64_t)row * a.n[0], a.src[3] + (int64_t)row * a.n[3], a.src[4] + (int64_t)row * a.n[4],
                    split + (int64_t)slot * gc_sample_bytes(a.c, a.hw), a.c, a.hw, i);
    return;
  }
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (i >= a.total) return;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    if (i < a.n[k]) {
      float *dst = reinterpret_cast<float *>(a.table[k * a.n_obj + o]);
      if (!dst) return;
      const float *src = a.src[k] + (int64_t)row * a.n[k] + i;
      dst += (int64_t)slot * a.n[k] + i;
      if (V == 4) *reinterpret_cast<float4 *>(dst) = *reinterpret_cast<const float4 *>(src);
      else *dst = *src;
      return;
    }
    i -= a.n[k];
  }
}

// ---------------------------------------------------------------- CG vector kernels (filter-only problem) --------
// cg_state layout: p[n] | r_prev[n] | rho | has_p | pad | pad           (persists across updates)
// work layout    : b/r[n] | x[n] | q[n] | scal[8]

// mode 0: finish RHS ( r = b = -(sum partial + reg^2 f) ), x = 0, apply the forgetting factor, then first direction.
// mode 1: finish A p ( q = sum partial + reg^2 p ), alpha step, optional residual update, next direction.
// mode 2: like mode 1 but last CG iteration of the GN step: no residual update, no new direction, f += x.
__global__ void __launch_bounds__(1024) cg_vector_kernel(CgVec s, int mode, const int *__restrict__ gate, int min_px,
                                                         const long long *__restrict__ table, int n_obj) {
  if (table) {   // object-batched: one CTA per object, pointers from the table
    const int o = blockIdx.x;
    float *cgst = reinterpret_cast<float *>(table[5 * n_obj + o]);
    s.f = reinterpret_cast<float *>(table[4 * n_obj + o]);
    s.p = cgst; s.rprev = cgst + s.n; s.rho = cgst + 2 * s.n; s.hasp = cgst + 2 * s.n + 1;
    s.r += (int64_t)o * 3 * s.n; s.x += (int64_t)o * 3 * s.n; s.q += (int64_t)o * 3 * s.n;
    s.partial += (int64_t)o * s.cap * s.n;
    gate = reinterpret_cast<const int *>(table[6 * n_obj + o]);
  }
  if (gate && gate[0] < min_px) return;
  __shared__ float red[32];
  const int t = threadIdx.x;
  const bool act = t < s.n;
  float g = 0.f;
  if (act) {   // fixed summation order, 8 independent loads in flight
    float g8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int k = 0;
    for (; k + 8 <= s.cap; k += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) g8[u] += s.partial[(int64_t)(k + u) * s.n + t];
    }
    for (; k < s.cap; ++k) g8[0] += s.partial[(int64_t)k * s.n + t];
    g = ((g8[0] + g8[1]) + (g8[2] + g8[3])) + ((g8[4] + g8[5]) + (g8[6] + g8[7]));
  }
  float r = 0.f, p = 0.f, rp = 0.f;
  float rho = *s.rho;
  // direction_forget_factor == 0: the CG state is reset at the start of every run (optimizer.py:93-103)
  const bool hasp = *s.hasp != 0.f && !(mode == 0 && s.forget == 0.f);
  if (mode == 0) {
    if (act) {
      r = -(g + s.reg2 * s.f[t]);
      s.x[t] = 0.f;
      p = hasp ? s.p[t] : 0.f;
      rp = hasp ? s.rprev[t] : 0.f;
    }
    if (hasp) rho = rho / s.forget;                    // optimizer.py:104-105
  } else {
    float q = 0.f;
    if (act) {
      p = s.p[t];
      q = g + s.reg2 * p;
      r = s.r[t];
    }
    const float pq = block_sum(act ? p * q : 0.f, red);
    const float alpha = rho / pq;                      // standard_alpha, optimizer.py:134
    if (act) {
      rp = r;                                          // r_prev = r.clone()
      s.rprev[t] = rp;
      const float xn = s.x[t] + alpha * p;
      s.x[t] = xn;
      if (mode == 1) r = r - alpha * q;
      else s.f[t] += xn;                               // theta += step_alpha * delta_x
      s.r[t] = r;
    }
    if (mode == 2) return;
  }
  // next direction (optimizer.py:115-128): z = r / diag_M ; rho = <r,z> ; beta = max((rho - <r_prev,z>)/rho1, 0)
  const float z = r * s.minv;
  const float rho_new = block_sum(act ? r * z : 0.f, red);
  float pn = z;
  if (mode != 0 || hasp) {
    const float rho2 = block_sum(act ? rp * z : 0.f, red);
    const float beta = fmaxf((rho_new - rho2) / rho, 0.f);
    pn = z + p * beta;
  }
  if (act) {
    s.p[t] = pn;
    if (mode == 0) s.r[t] = r;
  }
  __syncthreads();
  if (t == 0) {
    *s.rho = rho_new;
    *s.hasp = 1.f;
  }
}

// ---------------------------------------------------------------- small vector helpers (joint problem) -----------
constexpr int VB = 64;  // blocks used by the multi-block vector kernels

__device__ __forceinline__ float seg_sum(const float *__restrict__ part, int stride) {
  float s = 0.f;
  for (int i = 0; i < VB; ++i) s += part[i * stride];
  return s;
}

// part[b][0] = sum_chunk (r0 r0 m0 + r1 r1 m1), part[b][1] = sum_chunk (rp0 r0 m0 + rp1 r1 m1)
__global__ void __launch_bounds__(256) joint_dots_rz_kernel(const float *r0, const float *rp0, int64_t n0, float m0,
                                                            const float *r1, const float *rp1, int64_t n1, float m1,
                                                            float *part) {
  __shared__ float red[32];
  float a = 0.f, b = 0.f;
  const int64_t tot = n0 + n1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
    const bool first = i < n0;
    const float r = first ? r0[i] : r1[i - n0];
    const float rp = first ? rp0[i] : rp1[i - n0];
    const float z = r * (first ? m0 : m1);
    a = fmaf(r, z, a);
    b = fmaf(rp, z, b);
  }
  a = block_sum(a, red);
  b = block_sum(b, red);
  if (threadIdx.x == 0) { part[blockIdx.x * 2] = a; part[blockIdx.x * 2 + 1] = b; }
}

// scal: [0]=rho [1]=has_p ; p = z + beta p
__global__ void __launch_bounds__(256) joint_direction_kernel(float *p0, const float *r0, int64_t n0, float m0, float *p1,
                                                              const float *r1, int64_t n1, float m1, const float *part,
                                                              const float *scal_in, float *scal_out, float forget_div) {
  const float rho_new = seg_sum(part, 2), rho2 = seg_sum(part + 1, 2);
  const bool hasp = scal_in[1] != 0.f;
  const float rho1 = scal_in[0] / forget_div;   // forget_div = forget on the first iteration of a run, else 1
  const float beta = hasp ? fmaxf((rho_new - rho2) / rho1, 0.f) : 0.f;
  const int64_t tot = n0 + n1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < n0) p0[i] = r0[i] * m0 + (hasp ? p0[i] * beta : 0.f);
    else { const int64_t k = i - n0; p1[k] = r1[k] * m1 + (hasp ? p1[k] * beta : 0.f); }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { scal_out[0] = rho_new; scal_out[1] = 1.f; }
}

__global__ void __launch_bounds__(256) joint_dot_pq_kernel(const float *p0, const float *q0, int64_t n0, const float *p1,
                                                           const float *q1, int64_t n1, float *part) {
  __shared__ float red[32];
  float a = 0.f;
  const int64_t tot = n0 + n1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x)
    a = (i < n0) ? fmaf(p0[i], q0[i], a) : fmaf(p1[i - n0], q1[i - n0], a);
  a = block_sum(a, red);
  if (threadIdx.x == 0) part[blockIdx.x] = a;
}

__global__ void __launch_bounds__(256) joint_step_kernel(float *x0, float *r0, float *rp0, const float *p0, const float *q0,
                                                         int64_t n0, float *x1, float *r1, float *rp1, const float *p1,
                                                         const float *q1, int64_t n1, const float *part,
                                                         const float *scal, int update_r) {
  const float alpha = scal[0] / seg_sum(part, 1);
  const int64_t tot = n0 + n1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
    float *x = i < n0 ? x0 + i : x1 + (i - n0);
    float *r = i < n0 ? r0 + i : r1 + (i - n0);
    float *rp = i < n0 ? rp0 + i : rp1 + (i - n0);
    const float p = i < n0 ? p0[i] : p1[i - n0];
    const float q = i < n0 ? q0[i] : q1[i - n0];
    *rp = *r;
    *x = *x + alpha * p;
    if (update_r) *r = *r - alpha * q;
  }
}

// y = a*x + b*y ; used for q = g + reg^2 p, b = -(g + reg^2 theta), theta += x
__global__ void axpby_kernel(float *y, const float *x, float a, float b, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a * x[i] + b * y[i];
}
// out = a*x + b*y
__global__ void axpby_out_kernel(float *out, const float *x, float a, const float *y, float b, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a * x[i] + b * y[i];
}
// (c, C) <-> (C, cpad) transposes of the projection matrix
__global__ void transpose_pad_kernel(const float *src, int rows, int cols, float *dst, int ld) {  // dst[col][row]
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  dst[(int64_t)c * ld + r] = src[i];
}
// dst[c][r] = src[r][c]   (rows x cols -> cols x rows), 32 x 32 tiles through shared memory
__global__ void transpose_tiled_kernel(const float *__restrict__ src, int rows, int cols, float *__restrict__ dst) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? src[(int64_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[(int64_t)c * rows + r] = tile[threadIdx.x][j];
  }
}
__global__ void untranspose_kernel(const float *src, int rows, int cols, int ld, float *dst) {  // dst[r][c] = src[c][r]
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  dst[i] = src[(int64_t)c * ld + r];
}


}  // namespace frtm

using namespace frtm;

// =====================================================================================================================
// C ABI
// =====================================================================================================================
extern "C" int frtm_corr3x3_nchw(const float *x, const float *filt, const int *filter_index, int NB, int c, int h, int w,
                                 float *out, void *stream) {
  FRTM_REQUIRE(x && filt && out && NB > 0 && c > 0, "corr3x3: bad arguments");
  FRTM_REQUIRE(c * 9 * sizeof(float) <= 48 * 1024, "corr3x3: too many channels");
  dim3 grid(cdiv(h * w, 128), NB);
  corr3x3_kernel<<<grid, 512, (c * 9 + 512) * sizeof(float), (cudaStream_t)stream>>>(x, filt, filter_index, c, h, w, out, 0, nullptr);
  FRTM_CHECK_LAUNCH("corr3x3");
  return FRTM_OK;
}

extern "C" int frtm_pixel_weights(const float *y, int K, int HW, float tf, int threshold, float *w, float *workspace,
                                  const int *counts, void *stream) {
  FRTM_REQUIRE(y && w && (workspace || counts) && K > 0, "pixel_weights: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (!counts) {
    pixel_count_kernel<<<K, 1024, 0, st>>>(y, HW, threshold, workspace);
    FRTM_CHECK_LAUNCH("pixel_count");
  }
  pixel_weights_kernel<<<dim3(cdiv(HW, 256), K), 256, 0, st>>>(y, workspace, counts, HW, tf, threshold, w);
  FRTM_CHECK_LAUNCH("pixel_weights");
  return FRTM_OK;
}

extern "C" int frtm_build_stencil(const float *pw, const float *y, int K, int H, int W, int h, int w, float *stencil,
                                  float *uty, void *stream) {
  FRTM_REQUIRE(pw && y && stencil && uty && K > 0, "build_stencil: bad arguments");
  FRTM_REQUIRE(H >= h && W >= w, "build_stencil: expects an upsampling geometry (H>=h, W>=w)");
  build_stencil_kernel<<<dim3(cdiv(h * w, 8), K), 256, 0, (cudaStream_t)stream>>>(pw, y, H, W, h, w, stencil, uty);
  FRTM_CHECK_LAUNCH("build_stencil");
  return FRTM_OK;
}

extern "C" int frtm_memory_next_slot(float *weights, int capacity, float lr, int *state, const int *gate_count, int min_px,
                                     void *stream) {
  FRTM_REQUIRE(weights && state && capacity > 0, "memory_next_slot: bad arguments");
  memory_next_slot_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(weights, capacity, lr, state, gate_count, min_px);
  FRTM_CHECK_LAUNCH("memory_next_slot");
  return FRTM_OK;
}

extern "C" int frtm_memory_insert(const float *feat, int feat_elems, const float *label, const float *pw, int HW,
                                  const float *stencil, const float *uty, int hw, float *mem_samples, float *mem_labels,
                                  float *mem_pw, float *mem_stencil, float *mem_uty, void *mem_split, const int *state,
                                  void *stream) {
  FRTM_REQUIRE(feat && mem_samples && state, "memory_insert: bad arguments");
  InsertArgs a;
  const float *src[5] = {feat, label, pw, stencil, uty};
  float *dst[5] = {mem_samples, mem_labels, mem_pw, mem_stencil, mem_uty};
  const int64_t n[5] = {feat_elems, HW, HW, 9 * (int64_t)hw, hw};
  a.total = 0;
  for (int k = 0; k < 5; ++k) {
    const bool on = src[k] != nullptr && dst[k] != nullptr;
    a.src[k] = src[k]; a.dst[k] = dst[k]; a.n[k] = on ? n[k] : 0;
    a.total += a.n[k];
  }
  a.split = (uint8_t *)mem_split; a.c = 0; a.hw = hw; a.split_items = 0;
  if (mem_split) {
    FRTM_REQUIRE(hw > 0 && feat_elems % hw == 0 && (feat_elems / hw) % 8 == 0, "memory_insert: operator image needs c %% 8 == 0");
    FRTM_REQUIRE(stencil && uty, "memory_insert: operator image needs the sample's stencil and uty");
    a.c = feat_elems / hw;
    a.split_items = gc_sample_items(a.c, hw);
  }
  memory_insert_kernel<<<cdiv(a.total + a.split_items, 256), 256, 0, (cudaStream_t)stream>>>(a, state);
  FRTM_CHECK_LAUNCH("memory_insert");
  return FRTM_OK;
}

extern "C" int frtm_memory_insert_block(const void *table, int n_obj, int n_frames, int capacity, float lr, const int *gate_counts,
                                        int min_px, const float *feat, int feat_elems, const float *labels, const float *pw,
                                        int HW, const float *stencil, const float *uty, int hw, int with_split, int with_fullres,
                                        int *slots, void *stream) {
  FRTM_REQUIRE(table && gate_counts && feat && labels && pw && stencil && uty && slots, "memory_insert_block: null pointer");
  FRTM_REQUIRE(n_obj >= 1 && n_frames >= 1 && n_obj * n_frames <= 65535 && capacity >= 1, "memory_insert_block: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  memory_slots_block_kernel<<<n_obj, 32, 0, st>>>((const long long *)table, n_obj, n_frames, capacity, lr, gate_counts, min_px, slots);
  FRTM_CHECK_LAUNCH("memory_slots_block");
  MemBlockArgs a;
  a.table = (const long long *)table; a.slots = slots;
  const float *src[5] = {feat, labels, pw, stencil, uty};
  // without the full-resolution mirrors (labels, pixel weights) the copy grid covers a sixth of the elements
  const int64_t n[5] = {feat_elems, with_fullres ? HW : 0, with_fullres ? HW : 0, 9 * (int64_t)hw, hw};
  a.total = 0;
  for (int k = 0; k < 5; ++k) { a.src[k] = src[k]; a.n[k] = n[k]; a.total += n[k]; }
  a.n_obj = n_obj; a.nF = n_frames; a.hw = hw; a.c = 0; a.split_items = 0;
  if (with_split) {
    FRTM_REQUIRE(hw > 0 && feat_elems % hw == 0 && (feat_elems / hw) % 8 == 0, "memory_insert_block: operator image needs c %% 8 == 0");
    a.c = feat_elems / hw;
    a.split_items = gc_sample_items(a.c, hw);
  }
  bool vec = true;
  for (int k = 0; k < 5; ++k) vec = vec && n[k] % 4 == 0 && (reinterpret_cast<uintptr_t>(src[k]) & 15) == 0;
  const int copy_blocks = cdiv(vec ? a.total / 4 : a.total, 256);
  dim3 grid((unsigned)(copy_blocks + cdiv(a.split_items, 256)), (unsigned)(n_obj * n_frames));
  if (vec) memory_insert_block_kernel<4><<<grid, 256, 0, st>>>(a, copy_blocks);   // (the memory's own arrays come from the
  else memory_insert_block_kernel<1><<<grid, 256, 0, st>>>(a, copy_blocks);      // allocator: 256-byte aligned)
  FRTM_CHECK_LAUNCH("memory_insert_block");
  return FRTM_OK;
}

// ---- filter-only GN/CG ------------------------------------------------------------------------------------------
namespace frtm {
// ---- work list + workspace of the list-driven operator kernels -------------------------------------------------------
__global__ void __launch_bounds__(256) gn_build_items_kernel(const float *sw_single, const long long *table, int n_obj, int cap,
                                                             ClList L) {
  __shared__ int wsum[8];
  __shared__ int base_s;
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  if (tid == 0) base_s = 0;
  __syncthreads();
  const int total = n_obj * cap;
  for (int e0 = 0; e0 < total; e0 += 256) {
    const int e = e0 + tid;
    const int o = e < total ? e / cap : 0, slot = e < total ? e - o * cap : 0;
    const float *sw = table ? reinterpret_cast<const float *>(table[3 * n_obj + o]) : sw_single;
    const bool act = e < total && sw[slot] != 0.f;
    const unsigned m = __ballot_sync(0xffffffffu, act);
    const int before = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) wsum[wp] = __popc(m);
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { if (k < wp) woff += wsum[k]; tot += wsum[k]; }
    const int idx = base_s + woff + before;
    if (act) L.items[idx] = ((uint32_t)o << 16) | (uint32_t)slot;
    if (e < total && slot == 0) L.hdr[1 + o] = idx;
    __syncthreads();
    if (tid == 0) base_s += tot;
    __syncthreads();
  }
  if (tid == 0) { L.hdr[0] = base_s; L.hdr[1 + n_obj] = base_s; }
}


int gn_items_build(const GaArgs &a, ClList list, cudaStream_t st) {
  gn_build_items_kernel<<<1, 256, 0, st>>>(a.sw, a.table, a.n_obj, a.cap, list);
  FRTM_CHECK_LAUNCH("gn_build_items");
  return FRTM_OK;
}
static int64_t gn_list_ngrp(int cap) { return (cap + 160 + GC_RGROUP - 1) / GC_RGROUP; }
int64_t gn_list_workspace_bytes(int n_obj, int cap, int c) {
  const int64_t n = (int64_t)c * 9, ngrp = gn_list_ngrp(cap);
  // rows [n_obj*cap + 160*n_obj][n] | gsum [n_obj][ngrp][n] | counters [n_obj][1 + ngrp] | tickets [n_obj] | hdr [2 + n_obj] | items
  return (((int64_t)n_obj * (cap + 160)) * n + n_obj * ngrp * n + n_obj * (1 + ngrp) + n_obj + (2 + n_obj) + (int64_t)n_obj * cap + 16) * 4;
}
GnListWs gn_list_workspace(float *base, int n_obj, int cap, int c) {
  const int64_t n = (int64_t)c * 9, ngrp = gn_list_ngrp(cap);
  GnListWs w;
  w.rows = base;
  w.gsum = base + ((int64_t)n_obj * (cap + 160)) * n;
  w.counters = reinterpret_cast<int *>(w.gsum + n_obj * ngrp * n);
  w.tickets = w.counters + n_obj * (1 + ngrp);
  w.list.hdr = w.tickets + n_obj;
  w.list.items = reinterpret_cast<uint32_t *>(w.list.hdr + 2 + n_obj);
  w.ngrp_max = (int)ngrp;
  return w;
}

}  // namespace frtm

extern "C" int64_t frtm_gn_update_workspace(int cap, int c, int h, int w) {
  const int64_t n = (int64_t)c * 9, hw = (int64_t)h * w;
  // s[cap][hw] | v[cap][hw] | partial[cap][n] | r[n] | x[n] | q[n] | tickets[1 + ngroups] | group sums[ngroups][n]
  const int64_t ngrp = (cap + GC_RGROUP - 1) / GC_RGROUP;
  return (2 * cap * hw + cap * n + 3 * n + 64 + (1 + ngrp) + 8 + ngrp * n) * (int64_t)sizeof(float) + gn_list_workspace_bytes(1, cap, c);
}

extern "C" int64_t frtm_gn_operator_kind(int c, int h, int w) {
  return gn_apply_mma_supported(c, h, w) ? 3 : (gn_apply_tc_supported(c, h, w) ? 2 : 1);
}

static int gn_update_impl(const float *samples, const __half *samples_split, const float *stencil, const float *uty,
                          const float *weights, const long long *table, bool table_has_split, int n_obj, int cap, int c, int h, int w, float *filt, float *cg_state,
                          const int *cg_iters, int n_gn, float reg, float precond, float forget, const int *gate_count,
                          int min_px, int operator_select, float *workspace, int64_t workspace_bytes, cudaStream_t st) {
  FRTM_REQUIRE(cg_iters && workspace && n_obj >= 1, "gn_update: null pointer");
  FRTM_REQUIRE(c * 9 <= 1024, "gn_update: filter too large for the single-block CG kernel (c*9 <= 1024)");
  FRTM_REQUIRE(workspace_bytes >= n_obj * frtm_gn_update_workspace(cap, c, h, w), "gn_update: workspace too small");
  FRTM_REQUIRE(forget >= 0.f, "gn_update: direction_forget_factor must be >= 0 (0 resets the CG state at every run, optimizer.py:102-103)");
  const int n = c * 9, hw = h * w;
  FRTM_REQUIRE(c % 4 == 0, "gn_update: needs c %% 4 == 0 (got %d)", c);
  const bool fast = (w % 2 == 0);
  const size_t ga_smem = ((size_t)9 * hw + 2 * (size_t)(h + 2) * (w + 2) + (size_t)c * 12) * sizeof(float) + 16;
  FRTM_REQUIRE(ga_smem <= 227 * 1024, "gn_update: feature map %dx%d too large for the shared-memory resident tap maps", h, w);
  static size_t ga_configured = 0;
  if (ga_smem > ga_configured) {
    cudaError_t e = cudaFuncSetAttribute(gn_apply_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ga_smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gn_apply_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ga_smem);
    if (e != cudaSuccess) { set_error("gn_update: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return FRTM_ELAUNCH; }
    ga_configured = ga_smem;
  }
  // workspace: partial[n_obj][cap][n] | per object r[n] x[n] q[n]
  float *partial = workspace;
  float *vecs = partial + (int64_t)n_obj * cap * n;
  GaArgs ga;
  ga.X = samples; ga.S = stencil; ga.T = uty; ga.sw = weights; ga.pvec = filt; ga.table = table; ga.partial = partial;
  ga.XS = samples_split; ga.dbg = nullptr;
  // tensor-core operator kernel: needs the split tile image of the samples (written at insert time)
  // Operator kernels over the operator images (written at insert time): the single-pass mma.sync kernel where the shape
  // allows it, else the two-pass tcgen05 kernel, else CUDA cores.  operator_select forces one (tests, A-B measurements).
  const bool have_image = table ? table_has_split : samples_split != nullptr;
  FRTM_REQUIRE(operator_select >= 0 && operator_select <= 4, "gn_update: operator_select must be 0..4");
  FRTM_REQUIRE(operator_select < 2 || have_image, "gn_update: the tensor-core operators need the operator images");
  FRTM_REQUIRE(operator_select != 2 || gn_apply_tc_supported(c, h, w), "gn_update: shape not supported by the two-pass operator");
  FRTM_REQUIRE(operator_select != 3 || gn_apply_mma_supported(c, h, w), "gn_update: shape not supported by the single-pass operator");
  // the cluster operator gives every cluster a contiguous range of the active samples (at most 96 per launch and cluster)
  const bool cl_ok = gn_apply_cl_supported(c, h, w) && (int64_t)n_obj * cap <= 16 * 96 && n_obj < 65536 && cap < 65536;
  FRTM_REQUIRE(operator_select != 4 || cl_ok, "gn_update: shape not supported by the cluster operator");
  // (measured on B200, profiles/r02_gn_operator.md: the sliding-window kernel is the fastest of the three on the BASELINE
  //  shapes, so it is what 0 resolves to; the cluster kernel runs only when asked for)
  const bool use_cl = operator_select == 4;
  const bool use_mma = !use_cl && (operator_select == 3 || (operator_select == 0 && have_image && gn_apply_mma_supported(c, h, w)));
  const bool use_tc = use_cl || use_mma || operator_select == 2 || (operator_select == 0 && have_image && gn_apply_tc_supported(c, h, w));
  ga.n_obj = n_obj; ga.cap = cap; ga.c = c; ga.h = h; ga.w = w; ga.use_y = 1;
  CgVec cg;
  cg.f = filt; cg.p = cg_state; cg.rprev = cg_state ? cg_state + n : nullptr; cg.rho = cg_state ? cg_state + 2 * n : nullptr;
  cg.hasp = cg_state ? cg_state + 2 * n + 1 : nullptr;
  cg.r = vecs; cg.x = vecs + n; cg.q = vecs + 2 * n;
  cg.partial = partial; cg.n = n; cg.cap = cap; cg.reg2 = reg * reg; cg.minv = 1.f / precond; cg.forget = forget;
  const dim3 grid(cap, table ? n_obj : 1);
  // bulk-copy staged kernel: needs an even width (pixel pairs stay inside a row), hw % 4 == 0 (16-byte bulk copies),
  // one pixel pair per compute thread, one channel quad per compute warp, and the ring must fit in shared memory
  int CQ = 0, NST = 0, stage_floats = 0;
  size_t gt_smem = 0;
  if (fast && hw % 4 == 0 && hw / 2 <= GT_CONSUMER_WARPS * 32 && c % 4 == 0) {
    const size_t fixed = ((size_t)9 * hw + 2 * (size_t)(h + 2) * (w + 2) + (size_t)c * 12) * sizeof(float) + 256;
    for (int cq = 4; cq >= 1 && CQ == 0; cq >>= 1) {
      if (c % cq) continue;
      const int sf = (cq * hw + 31) / 32 * 32;
      for (int nst = 4; nst >= 2; --nst) {
        const size_t total = fixed + (size_t)nst * sf * sizeof(float) + 16 * nst;
        if (total <= 227 * 1024) { CQ = cq; NST = nst; stage_floats = sf; gt_smem = total; break; }
      }
    }
  }
  static size_t gt_configured = 0;
  if (CQ && gt_smem > gt_configured) {
    cudaError_t e = cudaFuncSetAttribute(gn_apply_tma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gt_smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gn_apply_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gt_smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gn_apply_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gt_smem);
    if (e != cudaSuccess) { cudaGetLastError(); CQ = 0; } else gt_configured = gt_smem;
  }
  // tensor-core path: the operator kernel's last CTA per object runs the CG vector step itself (tickets in the workspace)
  GcFuse fuse;
  fuse.cg = cg; fuse.gate = gate_count; fuse.mode = 0; fuse.min_px = min_px; fuse.enabled = use_tc ? 1 : 0;
  const int ngrp = (cap + GC_RGROUP - 1) / GC_RGROUP;
  fuse.counters = reinterpret_cast<int *>(vecs + (int64_t)n_obj * 3 * n);
  fuse.gsum = vecs + (int64_t)n_obj * 3 * n + (((int64_t)n_obj * (1 + ngrp) + 3) & ~(int64_t)3);
  if (use_tc) {
    if (cudaMemsetAsync(fuse.counters, 0, sizeof(int) * n_obj * (1 + ngrp), st) != cudaSuccess) {
      set_error("gn_update: cudaMemsetAsync failed");
      return FRTM_ELAUNCH;
    }
  }
  // the list-driven operators' region sits behind the n_obj older layouts: rows per unit / (object, cluster), group sums,
  // tickets, the work list of this update (built once: the sample weights do not change between operator applications)
  float *ws_list = workspace + (n_obj * (frtm_gn_update_workspace(cap, c, h, w) - gn_list_workspace_bytes(1, cap, c))) / (int64_t)sizeof(float);
  const GnListWs lws = gn_list_workspace(ws_list, n_obj, cap, c);
  if (use_cl || use_mma) {
    if (cudaMemsetAsync(lws.counters, 0, sizeof(int) * ((int64_t)n_obj * (2 + lws.ngrp_max)), st) != cudaSuccess) {
      set_error("gn_update: cudaMemsetAsync failed");
      return FRTM_ELAUNCH;
    }
    if (int rc = gn_items_build(ga, lws.list, st)) return rc;
  }
  auto launch_apply = [&]() -> int {
    if (use_tc) {
      const int rc = use_cl ? gn_apply_cl_launch(ga, fuse, lws, st)
                            : use_mma ? gn_apply_mma_launch(ga, fuse, lws, st) : gn_apply_tc_launch(ga, fuse, st);
      if (rc == FRTM_OK) count_launch(-1);                  // counted again by FRTM_CHECK_LAUNCH at the call site
      return rc;
    }
    if (CQ == 4) gn_apply_tma_kernel<4><<<grid, GT_THREADS, gt_smem, st>>>(ga, NST, stage_floats);
    else if (CQ == 2) gn_apply_tma_kernel<2><<<grid, GT_THREADS, gt_smem, st>>>(ga, NST, stage_floats);
    else if (CQ == 1) gn_apply_tma_kernel<1><<<grid, GT_THREADS, gt_smem, st>>>(ga, NST, stage_floats);
    else if (fast) gn_apply_kernel<true><<<grid, GA_THREADS, ga_smem, st>>>(ga);
    else gn_apply_kernel<false><<<grid, GA_THREADS, ga_smem, st>>>(ga);
    return FRTM_OK;
  };
  // gating: the tiny vector kernel checks the gate and skips all arithmetic; the streaming kernel is harmless (it only
  // writes workspace), so it is launched unconditionally to keep the stream free of host syncs.
  for (int gi = 0; gi < n_gn; ++gi) {
    const int iters = cg_iters[gi];
    if (iters <= 0) continue;
    ga.use_y = 1; ga.pvec = filt;        // RHS: partial_i = X_i^T sw_i (S_i (X_i * f) - t_i)
    fuse.mode = 0;
    if (int rc = launch_apply()) return rc;
    FRTM_CHECK_LAUNCH("gn_update/apply(rhs)");
    if (!use_tc) {
      cg_vector_kernel<<<table ? n_obj : 1, 1024, 0, st>>>(cg, 0, gate_count, min_px, table, n_obj);
      FRTM_CHECK_LAUNCH("gn_update/cg(rhs)");
    }
    ga.use_y = 0; ga.pvec = cg.p;
    for (int it = 0; it < iters; ++it) {
      fuse.mode = it == iters - 1 ? 2 : 1;
      if (int rc = launch_apply()) return rc;
      FRTM_CHECK_LAUNCH("gn_update/apply");
      if (!use_tc) {
        cg_vector_kernel<<<table ? n_obj : 1, 1024, 0, st>>>(cg, fuse.mode, gate_count, min_px, table, n_obj);
        FRTM_CHECK_LAUNCH("gn_update/cg");
      }
    }
  }
  return FRTM_OK;
}

extern "C" int frtm_gn_update(const float *samples, const void *samples_split, const float *stencil, const float *uty,
                              const float *weights, int cap, int c, int h, int w, float *filt, float *cg_state,
                              const int *cg_iters, int n_gn, float reg, float precond, float forget, const int *gate_count,
                              int min_px, int operator_select, float *workspace, int64_t workspace_bytes, void *stream) {
  FRTM_REQUIRE(samples && stencil && uty && weights && filt && cg_state, "gn_update: null pointer");
  return gn_update_impl(samples, (const __half *)samples_split, stencil, uty, weights, nullptr, false, 1, cap, c, h, w, filt, cg_state, cg_iters, n_gn, reg, precond,
                        forget, gate_count, min_px, operator_select, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int frtm_gn_update_batched(const void *table, int n_obj, int has_split, int cap, int c, int h, int w,
                                      const int *cg_iters, int n_gn, float reg, float precond, float forget, int min_px,
                                      int operator_select, float *workspace, int64_t workspace_bytes, void *stream) {
  FRTM_REQUIRE(table && n_obj >= 1, "gn_update_batched: null table");
  return gn_update_impl(nullptr, nullptr, nullptr, nullptr, nullptr, reinterpret_cast<const long long *>(table), has_split != 0, n_obj, cap, c, h, w,
                        nullptr, nullptr, cg_iters, n_gn, reg, precond, forget, nullptr, min_px, operator_select, workspace, workspace_bytes,
                        (cudaStream_t)stream);
}

// ---- rank-9 form of the joint problem's operator ---------------------------------------------------------------
// The target model's output is ONE channel, so the two products of J and J^T with the projection matrix collapse:
//   F * (dP x)            = conv3x3(x; G),  G[ch][tap] = sum_c F[c][tap] dP[c][ch]          (C x 9 instead of c x C)
//   J_P^T v = x^T (V F^T) = (x^T V) F^T,    H[ch][tap] = sum_p x[p][ch] v[p - tap]           (C x 9), gP[c][ch] = sum_tap F[c][tap] H[ch][tap]
// i.e. 2 x 9 FMAs per element of x instead of 2 x 96: the two dense (c x C x pixels) contractions per operator application
// (the reference's conv / conv-backward of the 1x1 projection, model/discriminator.py:168-175) become two streaming passes
// over x.  Mathematically identical, the summation order differs.
constexpr int JG_LD = 12;                  // G / tap-map rows padded to 12 floats (three 16-byte loads)

// G[ch][t] = sum_c Pt[ch][c] F[c][t]
__global__ void __launch_bounds__(256) joint_G_kernel(const float *__restrict__ Pt, const float *__restrict__ F, int C, int c,
                                                      float *__restrict__ G) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * JG_LD) return;
  const int ch = i / JG_LD, t = i - ch * JG_LD;
  float acc = 0.f;
  if (t < 9)
    for (int cc = 0; cc < c; ++cc) acc = fmaf(Pt[(int64_t)ch * c + cc], F[cc * 9 + t], acc);
  G[i] = acc;
}

// Y[p][t] = sum_ch xT[ch][p] G[ch][t]     xT: (C, NP) channel-major copy of x; block = 32 pixels x 8 channel slices
__global__ void __launch_bounds__(256) joint_tapmaps_kernel(const float *__restrict__ xT, const float *__restrict__ G, int C, int NP,
                                                            float *__restrict__ Y) {
  __shared__ float red[8][9][33];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int p = blockIdx.x * 32 + lane;
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.f;
  const int c0 = (int)((int64_t)C * wp / 8), c1 = (int)((int64_t)C * (wp + 1) / 8);
  if (p < NP) {
#pragma unroll 8                                   // eight independent 128-byte loads in flight per warp
    for (int ch = c0; ch < c1; ++ch) {
      const float xv = xT[(int64_t)ch * NP + p];
      const float4 g0 = __ldg(reinterpret_cast<const float4 *>(G + (int64_t)ch * JG_LD));
      const float4 g1 = __ldg(reinterpret_cast<const float4 *>(G + (int64_t)ch * JG_LD + 4));
      const float g8 = __ldg(G + (int64_t)ch * JG_LD + 8);
      acc[0] = fmaf(xv, g0.x, acc[0]); acc[1] = fmaf(xv, g0.y, acc[1]); acc[2] = fmaf(xv, g0.z, acc[2]);
      acc[3] = fmaf(xv, g0.w, acc[3]); acc[4] = fmaf(xv, g1.x, acc[4]); acc[5] = fmaf(xv, g1.y, acc[5]);
      acc[6] = fmaf(xv, g1.z, acc[6]); acc[7] = fmaf(xv, g1.w, acc[7]); acc[8] = fmaf(xv, g8, acc[8]);
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) red[wp][t][lane] = acc[t];
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * 32; i += 256) {
    const int t = i >> 5, l = i & 31;
    const float sum = ((red[0][t][l] + red[1][t][l]) + (red[2][t][l] + red[3][t][l])) +
                      ((red[4][t][l] + red[5][t][l]) + (red[6][t][l] + red[7][t][l]));
    if (blockIdx.x * 32 + l < NP) Y[(int64_t)(blockIdx.x * 32 + l) * JG_LD + t] = sum;
  }
}

// s[n][pix] = sum_t Y[n*hw + pix + off(t)][t]   (zero outside the map)
__global__ void __launch_bounds__(256) joint_gather_scores_kernel(const float *__restrict__ Y, int h, int w, float *__restrict__ s) {
  const int n = blockIdx.y, hw = h * w;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= hw) return;
  const int py = pix / w, px = pix - py * w;
  float acc = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
    if (yy >= 0 && yy < h && xx >= 0 && xx < w) acc += Y[((int64_t)n * hw + yy * w + xx) * JG_LD + t];
  }
  s[(int64_t)n * hw + pix] = acc;
}

// Hp[split][ch][t] = sum_{p in split} x[p][ch] v[p - off(t)]    x NHWC (NP, C); thread = channel; split = JH_PIX pixels of a sample
constexpr int JH_PIX = 128;
__global__ void __launch_bounds__(256) joint_gradx_kernel(const float *__restrict__ x, const float *__restrict__ v, int C, int h, int w,
                                                          int splits_per_sample, float *__restrict__ Hp) {
  __shared__ __align__(16) float Vs[JH_PIX][JG_LD];
  const int hw = h * w;
  const int n = blockIdx.y / splits_per_sample, sp = blockIdx.y - n * splits_per_sample;
  const int p0 = sp * JH_PIX, np = min(JH_PIX, hw - p0);
  for (int i = threadIdx.x; i < JH_PIX * JG_LD; i += 256) {
    const int lp = i / JG_LD, t = i - lp * JG_LD;
    float val = 0.f;
    if (lp < np && t < 9) {
      const int pix = p0 + lp, py = pix / w, px = pix - py * w;
      const int yy = py - (t / 3 - 1), xx = px - (t % 3 - 1);
      if (yy >= 0 && yy < h && xx >= 0 && xx < w) val = v[(int64_t)n * hw + yy * w + xx];
    }
    Vs[lp][t] = val;
  }
  __syncthreads();
  const int ch = blockIdx.x * 256 + threadIdx.x;
  if (ch >= C) return;
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.f;
  const float *xp = x + ((int64_t)n * hw + p0) * C + ch;
#pragma unroll 8
  for (int lp = 0; lp < np; ++lp) {
    const float xv = xp[(int64_t)lp * C];
    const float4 a = *reinterpret_cast<const float4 *>(&Vs[lp][0]), b = *reinterpret_cast<const float4 *>(&Vs[lp][4]);
    const float v8 = Vs[lp][8];
    acc[0] = fmaf(xv, a.x, acc[0]); acc[1] = fmaf(xv, a.y, acc[1]); acc[2] = fmaf(xv, a.z, acc[2]);
    acc[3] = fmaf(xv, a.w, acc[3]); acc[4] = fmaf(xv, b.x, acc[4]); acc[5] = fmaf(xv, b.y, acc[5]);
    acc[6] = fmaf(xv, b.z, acc[6]); acc[7] = fmaf(xv, b.w, acc[7]); acc[8] = fmaf(xv, v8, acc[8]);
  }
  float *dst = Hp + ((int64_t)blockIdx.y * C + ch) * 9;
#pragma unroll
  for (int t = 0; t < 9; ++t) dst[t] = acc[t];
}

// gP[ch][cc] = sum_t F[cc][t] H[ch][t],  H = sum_split Hp (fixed order); block = 8 channels
__global__ void __launch_bounds__(256) joint_gP_kernel(const float *__restrict__ Hp, int nsplit, const float *__restrict__ F, int C, int c,
                                                       float *__restrict__ gP) {
  __shared__ float Hs[8][9];
  const int ch0 = blockIdx.x * 8;
  if (threadIdx.x < 72) {
    const int lc = threadIdx.x / 9, t = threadIdx.x - lc * 9;
    float acc = 0.f;
    if (ch0 + lc < C)
      for (int sidx = 0; sidx < nsplit; ++sidx) acc += Hp[((int64_t)sidx * C + ch0 + lc) * 9 + t];
    Hs[lc][t] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 8 * c; i += 256) {
    const int lc = i / c, cc = i - lc * c;
    if (ch0 + lc >= C) continue;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc = fmaf(F[cc * 9 + t], Hs[lc][t], acc);
    gP[(int64_t)(ch0 + lc) * c + cc] = acc;
  }
}

// ---- joint (project, filter) GN/CG of Discriminator.init --------------------------------------------------------
namespace {
struct InitWs {
  // sizes
  int K, C, c, h, w, hw, nP, nF, sps, nsplit;
  // buffers (all float)
  float *Pt, *cx, *xT, *G, *Y, *s, *v, *Hp, *fpart, *rP, *rF, *rpP, *rpF, *pP, *pF, *qP, *qF, *xP, *xF, *scal, *dpart;
  int64_t total;
  InitWs(int K_, int C_, int c_, int h_, int w_, float *base) : K(K_), C(C_), c(c_), h(h_), w(w_) {
    hw = h * w; nP = C * c; nF = c * 9; sps = cdiv(hw, JH_PIX); nsplit = K * sps;
    int64_t o = 0;
    auto take = [&](int64_t n) { float *p = base ? base + o : nullptr; o += (n + 3) / 4 * 4; return p; };
    Pt = take(nP); cx = take((int64_t)K * c * hw); xT = take((int64_t)K * hw * C); G = take((int64_t)C * JG_LD);
    Y = take((int64_t)K * hw * JG_LD); s = take((int64_t)K * hw); v = take((int64_t)K * hw);
    Hp = take((int64_t)nsplit * C * 9); fpart = take((int64_t)K * nF);
    rP = take(nP); rF = take(nF); rpP = take(nP); rpF = take(nF); pP = take(nP); pF = take(nF); qP = take(nP); qF = take(nF);
    xP = take(nP); xF = take(nF); scal = take(16); dpart = take(2 * VB);
    total = o;
  }
};

// g = J^T [ sw (S Js(d) - use_y t) ]  for the joint problem at the current (Pt, F); d = (dP, dF) or null for the RHS
int joint_products(const InitWs &W, const float *x, const float *stencil, const float *uty, const float *sw, const float *F,
                   const float *dP, const float *dF, float *gP, float *gF, cudaStream_t st) {
  const size_t fsm = (size_t)W.nF * sizeof(float), csm = fsm + 512 * sizeof(float);
  const int NP = W.K * W.hw;
  dim3 gpix(cdiv(W.hw, 128), W.K), gpix256(cdiv(W.hw, 256), W.K), ggrad(cdiv(W.c, 8), W.K);
  if (dP == nullptr) {  // RHS: s = F * (P x)
    corr3x3_kernel<<<gpix, 512, csm, st>>>(W.cx, F, nullptr, W.c, W.h, W.w, W.s, 0, nullptr);
    FRTM_CHECK_LAUNCH("gn_init/score");
  } else {              // J d: s = F * (dP x) + dF * (P x), the first term as conv3x3(x; G(dP, F))
    joint_G_kernel<<<cdiv((int64_t)W.C * JG_LD, 256), 256, 0, st>>>(dP, F, W.C, W.c, W.G);
    FRTM_CHECK_LAUNCH("gn_init/G");
    joint_tapmaps_kernel<<<cdiv(NP, 32), 256, 0, st>>>(W.xT, W.G, W.C, NP, W.Y);
    FRTM_CHECK_LAUNCH("gn_init/tapmaps");
    joint_gather_scores_kernel<<<gpix256, 256, 0, st>>>(W.Y, W.h, W.w, W.s);
    FRTM_CHECK_LAUNCH("gn_init/score(dP)");
    corr3x3_kernel<<<gpix, 512, csm, st>>>(W.cx, dF, nullptr, W.c, W.h, W.w, W.s, 1, nullptr);
    FRTM_CHECK_LAUNCH("gn_init/score(dF)");
  }
  stencil_apply_kernel<<<gpix256, 256, 0, st>>>(stencil, W.s, uty, sw, W.h, W.w, dP == nullptr ? 1 : 0, W.v);
  FRTM_CHECK_LAUNCH("gn_init/stencil");
  corr3x3_grad_filter_kernel<<<ggrad, 256, 0, st>>>(W.cx, W.v, W.c, W.h, W.w, W.fpart, nullptr);
  FRTM_CHECK_LAUNCH("gn_init/gradF");
  reduce_rows_kernel<<<cdiv(W.nF, 256), 256, 0, st>>>(W.fpart, W.K, W.nF, gF);
  FRTM_CHECK_LAUNCH("gn_init/gradF.reduce");
  // J_P^T v = (x^T V) F^T
  joint_gradx_kernel<<<dim3(cdiv(W.C, 256), W.nsplit), 256, 0, st>>>(x, W.v, W.C, W.h, W.w, W.sps, W.Hp);
  FRTM_CHECK_LAUNCH("gn_init/gradx");
  joint_gP_kernel<<<cdiv(W.C, 8), 256, 0, st>>>(W.Hp, W.nsplit, F, W.C, W.c, gP);
  FRTM_CHECK_LAUNCH("gn_init/gradP");
  return FRTM_OK;
}
}  // namespace

extern "C" int64_t frtm_gn_init_workspace(int K, int C, int c, int h, int w) {
  InitWs W(K, C, c, h, w, nullptr);
  return W.total * (int64_t)sizeof(float);
}

// mode 0: full optimisation.  mode 1 (probe, for teacher-forced parity tests): out_b = RHS at (P,F), out_Ad = A (dP,dF).
static int gn_init_impl(const float *x, const float *stencil, const float *uty, const float *sw, int K, int C, int c, int h,
                        int w, float *P, float *F, const int *cg_iters, int n_gn, float regP, float regF, float mP, float mF,
                        float forget, float *ws, int64_t ws_bytes, cudaStream_t st, const float *dP, const float *dF,
                        float *out_bP, float *out_bF, float *out_AP, float *out_AF) {
  FRTM_REQUIRE(x && stencil && uty && sw && P && F && ws, "gn_init: null pointer");
  FRTM_REQUIRE(C % 4 == 0 && c % 4 == 0 && c <= 128 && c * 9 <= 1024, "gn_init: unsupported channel counts C=%d c=%d", C, c);
  FRTM_REQUIRE(ws_bytes >= frtm_gn_init_workspace(K, C, c, h, w), "gn_init: workspace too small");
  FRTM_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "gn_init: workspace must be 16-byte aligned");
  InitWs W(K, C, c, h, w, ws);
  const int64_t nP = W.nP, nF = W.nF;
  int rc;
  auto blocks = [](int64_t n) { return cdiv(n, 256); };
  transpose_pad_kernel<<<blocks(nP), 256, 0, st>>>(P, c, C, W.Pt, c);   // Pt[C][c]
  FRTM_CHECK_LAUNCH("gn_init/transpose");
  {   // channel-major copy of the features for the tap-map pass (x itself is pixel-major), once per call
    const int NP = K * h * w;
    transpose_tiled_kernel<<<dim3(cdiv(C, 32), cdiv(NP, 32)), dim3(32, 8), 0, st>>>(x, NP, C, W.xT);
    FRTM_CHECK_LAUNCH("gn_init/xT");
  }
  cudaMemsetAsync(W.scal, 0, 16 * sizeof(float), st);
  const bool probe = dP != nullptr;
  int scal_cur = 0;
  for (int gi = 0; gi < (probe ? 1 : n_gn); ++gi) {
    // cx = P x  (K,c,h,w)
    rc = frtm_conv2d_nhwc(x, K, h, w, C, C, W.Pt, nullptr, nullptr, 0, nullptr, 0, 0, W.cx, c, 1, 1, 1, 0, 0, st);
    if (rc) return rc;
    rc = joint_products(W, x, stencil, uty, sw, F, nullptr, nullptr, W.qP, W.qF, st);
    if (rc) return rc;
    // r = b = -(g + reg^2 theta)
    axpby_out_kernel<<<blocks(nP), 256, 0, st>>>(W.rP, W.qP, -1.f, W.Pt, -regP * regP, nP);
    FRTM_CHECK_LAUNCH("gn_init/rhsP");
    axpby_out_kernel<<<blocks(nF), 256, 0, st>>>(W.rF, W.qF, -1.f, F, -regF * regF, nF);
    FRTM_CHECK_LAUNCH("gn_init/rhsF");
    if (probe) {
      untranspose_kernel<<<blocks(nP), 256, 0, st>>>(W.rP, c, C, c, out_bP);
      FRTM_CHECK_LAUNCH("gn_init/probe.b");
      cudaMemcpyAsync(out_bF, W.rF, nF * sizeof(float), cudaMemcpyDeviceToDevice, st);
      transpose_pad_kernel<<<blocks(nP), 256, 0, st>>>(dP, c, C, W.pP, c);
      FRTM_CHECK_LAUNCH("gn_init/probe.dP");
      rc = joint_products(W, x, stencil, uty, sw, F, W.pP, dF, W.qP, W.qF, st);
      if (rc) return rc;
      axpby_kernel<<<blocks(nP), 256, 0, st>>>(W.qP, W.pP, regP * regP, 1.f, nP);
      FRTM_CHECK_LAUNCH("gn_init/probe.qP");
      axpby_kernel<<<blocks(nF), 256, 0, st>>>(W.qF, dF, regF * regF, 1.f, nF);
      FRTM_CHECK_LAUNCH("gn_init/probe.qF");
      untranspose_kernel<<<blocks(nP), 256, 0, st>>>(W.qP, c, C, c, out_AP);
      FRTM_CHECK_LAUNCH("gn_init/probe.A");
      cudaMemcpyAsync(out_AF, W.qF, nF * sizeof(float), cudaMemcpyDeviceToDevice, st);
      return FRTM_OK;
    }
    cudaMemsetAsync(W.xP, 0, nP * sizeof(float), st);
    cudaMemsetAsync(W.xF, 0, nF * sizeof(float), st);
    const int iters = cg_iters[gi];
    if (forget == 0.f) cudaMemsetAsync(W.scal, 0, 16 * sizeof(float), st);   // factor 0: the CG state is reset at every run (optimizer.py:102-103)
    for (int it = 0; it < iters; ++it) {
      float *sin = W.scal + 4 * scal_cur, *sout = W.scal + 4 * (scal_cur ^ 1);
      joint_dots_rz_kernel<<<VB, 256, 0, st>>>(W.rP, W.rpP, nP, 1.f / mP, W.rF, W.rpF, nF, 1.f / mF, W.dpart);
      FRTM_CHECK_LAUNCH("gn_init/dots_rz");
      joint_direction_kernel<<<VB, 256, 0, st>>>(W.pP, W.rP, nP, 1.f / mP, W.pF, W.rF, nF, 1.f / mF, W.dpart, sin, sout,
                                                 it == 0 && forget > 0.f ? forget : 1.f);
      FRTM_CHECK_LAUNCH("gn_init/direction");
      scal_cur ^= 1;
      rc = joint_products(W, x, stencil, uty, sw, F, W.pP, W.pF, W.qP, W.qF, st);
      if (rc) return rc;
      axpby_kernel<<<blocks(nP), 256, 0, st>>>(W.qP, W.pP, regP * regP, 1.f, nP);
      FRTM_CHECK_LAUNCH("gn_init/qP");
      axpby_kernel<<<blocks(nF), 256, 0, st>>>(W.qF, W.pF, regF * regF, 1.f, nF);
      FRTM_CHECK_LAUNCH("gn_init/qF");
      joint_dot_pq_kernel<<<VB, 256, 0, st>>>(W.pP, W.qP, nP, W.pF, W.qF, nF, W.dpart);
      FRTM_CHECK_LAUNCH("gn_init/dot_pq");
      joint_step_kernel<<<VB, 256, 0, st>>>(W.xP, W.rP, W.rpP, W.pP, W.qP, nP, W.xF, W.rF, W.rpF, W.pF, W.qF, nF, W.dpart,
                                            W.scal + 4 * scal_cur, it < iters - 1 ? 1 : 0);
      FRTM_CHECK_LAUNCH("gn_init/step");
    }
    if (iters > 0) {
      axpby_kernel<<<blocks(nP), 256, 0, st>>>(W.Pt, W.xP, 1.f, 1.f, nP);   // theta += delta
      FRTM_CHECK_LAUNCH("gn_init/updP");
      axpby_kernel<<<blocks(nF), 256, 0, st>>>(F, W.xF, 1.f, 1.f, nF);
      FRTM_CHECK_LAUNCH("gn_init/updF");
    }
  }
  untranspose_kernel<<<blocks(nP), 256, 0, st>>>(W.Pt, c, C, c, P);
  FRTM_CHECK_LAUNCH("gn_init/untranspose");
  return FRTM_OK;
}

// The joint optimisation is ~650 launches with a fixed schedule.  When the caller passes the same buffers again (the
// Python side stages every object through persistent buffers), the whole sequence is captured once into a CUDA graph and
// replayed: one launch per object instead of 650.  First call with a given signature runs eagerly (it also performs the
// one-time cudaFuncSetAttribute calls), the second captures + instantiates, later calls replay.
#include <map>
#include <vector>
#include <mutex>
namespace {
struct InitKey {
  std::vector<long long> v;
  bool operator<(const InitKey &o) const { return v < o.v; }
};
struct InitGraph { int calls = 0; int kernels = 1; cudaGraphExec_t exec = nullptr; long long stamp = 0; };
std::map<InitKey, InitGraph> g_init_graphs;
std::mutex g_init_mutex;
long long g_init_clock = 0;
constexpr size_t INIT_GRAPH_CAP = 16;      // signatures kept (one per staged buffer set / stream); least recently used goes first
// key.v[6] is the workspace pointer of the signature
void drop_init_graphs_locked(const void *workspace) {
  for (auto it = g_init_graphs.begin(); it != g_init_graphs.end();) {
    if (workspace == nullptr || it->first.v[6] == (long long)reinterpret_cast<uintptr_t>(workspace)) {
      if (it->second.exec) cudaGraphExecDestroy(it->second.exec);
      it = g_init_graphs.erase(it);
    } else ++it;
  }
}
}  // namespace

// Forget the captured optimisation graphs that use `workspace` (NULL: all of them).  The caller owns the staged buffers a
// graph was captured over; it calls this before freeing them (frtm_vos_b200.model.optimizer drops a staging set when its
// small LRU overflows and when the tracker is cleared), so no graph outlives its memory.
extern "C" int frtm_gn_init_release(const void *workspace) {
  std::lock_guard<std::mutex> lock(g_init_mutex);
  drop_init_graphs_locked(workspace);
  return FRTM_OK;
}

extern "C" int frtm_gn_init(const float *x_nhwc, const float *stencil, const float *uty, const float *sw, int K, int C, int c,
                            int h, int w, float *P, float *F, const int *cg_iters, int n_gn, float regP, float regF,
                            float precondP, float precondF, float forget, float *workspace, int64_t workspace_bytes,
                            void *stream) {
  FRTM_REQUIRE(cg_iters && n_gn >= 0, "gn_init: bad schedule");
  cudaStream_t st = (cudaStream_t)stream;
  auto eager = [&]() {
    return gn_init_impl(x_nhwc, stencil, uty, sw, K, C, c, h, w, P, F, cg_iters, n_gn, regP, regF, precondP, precondF, forget,
                        workspace, workspace_bytes, st, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  };
  InitKey key;
  for (const void *p : {(const void *)x_nhwc, (const void *)stencil, (const void *)uty, (const void *)sw, (const void *)P,
                        (const void *)F, (const void *)workspace, (const void *)stream})
    key.v.push_back((long long)reinterpret_cast<uintptr_t>(p));
  for (int q : {K, C, c, h, w, n_gn}) key.v.push_back(q);
  for (int i = 0; i < n_gn; ++i) key.v.push_back(cg_iters[i]);
  for (float f : {regP, regF, precondP, precondF, forget}) { long long b = 0; memcpy(&b, &f, sizeof(float)); key.v.push_back(b); }
  std::lock_guard<std::mutex> lock(g_init_mutex);
  if (g_init_graphs.size() >= INIT_GRAPH_CAP && g_init_graphs.find(key) == g_init_graphs.end()) {
    auto oldest = g_init_graphs.begin();
    for (auto it = g_init_graphs.begin(); it != g_init_graphs.end(); ++it)
      if (it->second.stamp < oldest->second.stamp) oldest = it;
    if (oldest->second.exec) cudaGraphExecDestroy(oldest->second.exec);
    g_init_graphs.erase(oldest);
  }
  InitGraph &g = g_init_graphs[key];
  g.stamp = ++g_init_clock;
  g.calls += 1;
  if (g.calls < 2 || g.calls < 0) return eager();
  // The legacy default stream cannot be captured: fork to an internal stream (event in / event out) for graph work.
  static cudaStream_t side = nullptr;
  static cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  cudaStream_t gs = st;
  const bool fork = (st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread);
  if (fork) {
    if (!side) {
      if (cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&ev_out, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError(); side = nullptr; g.calls = -1000000; return eager();
      }
    }
    gs = side;
    cudaEventRecord(ev_in, st);
    cudaStreamWaitEvent(gs, ev_in, 0);
  }
  auto join = [&]() {
    if (fork) { cudaEventRecord(ev_out, gs); cudaStreamWaitEvent(st, ev_out, 0); }
  };
  if (g.exec == nullptr) {
    // second call with this signature: capture on gs
    cudaError_t e = cudaStreamBeginCapture(gs, cudaStreamCaptureModeRelaxed);
    if (e != cudaSuccess) { cudaGetLastError(); g.calls = -1000000; join(); return eager(); }
    const int64_t before = frtm_launch_count();
    const int rc = gn_init_impl(x_nhwc, stencil, uty, sw, K, C, c, h, w, P, F, cg_iters, n_gn, regP, regF, precondP, precondF,
                                forget, workspace, workspace_bytes, gs, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    g.kernels = (int)(frtm_launch_count() - before);     // kernel nodes in the graph (counted once here, per replay below)
    count_launch(-g.kernels);
    cudaGraph_t graph = nullptr;
    e = cudaStreamEndCapture(gs, &graph);
    if (rc != FRTM_OK || e != cudaSuccess || graph == nullptr) {
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      g.calls = -1000000;                     // never try again for this signature
      join();
      return rc != FRTM_OK ? rc : eager();
    }
    e = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { cudaGetLastError(); g.exec = nullptr; g.calls = -1000000; join(); return eager(); }
  }
  cudaError_t e = cudaGraphLaunch(g.exec, gs);
  join();
  if (e != cudaSuccess) { set_error("gn_init: cudaGraphLaunch: %s", cudaGetLastError() == cudaSuccess ? cudaGetErrorString(e) : "launch failed"); return FRTM_ELAUNCH; }
  count_launch(g.kernels);                               // kernels executed by the replayed graph
  return FRTM_OK;
}

extern "C" int frtm_gn_init_probe(const float *x_nhwc, const float *stencil, const float *uty, const float *sw, int K, int C,
                                  int c, int h, int w, float *P, float *F, const float *dP, const float *dF, float regP,
                                  float regF, float *out_bP, float *out_bF, float *out_AP, float *out_AF, float *workspace,
                                  int64_t workspace_bytes, void *stream) {
  FRTM_REQUIRE(dP && dF && out_bP && out_bF && out_AP && out_AF, "gn_init_probe: null pointer");
  return gn_init_impl(x_nhwc, stencil, uty, sw, K, C, c, h, w, P, F, nullptr, 1, regP, regF, 1.f, 1.f, 1.f, workspace,
                      workspace_bytes, (cudaStream_t)stream, dP, dF, out_bP, out_bF, out_AP, out_AF);
}

// Stand-alone building blocks exported for the parity tests.
extern "C" int frtm_stencil_apply(const float *stencil, const float *s, const float *uty, const float *sw, int NB, int h, int w,
                                  int use_y, float *v, void *stream) {
  FRTM_REQUIRE(stencil && s && uty && sw && v, "stencil_apply: null pointer");
  stencil_apply_kernel<<<dim3(cdiv(h * w, 256), NB), 256, 0, (cudaStream_t)stream>>>(stencil, s, uty, sw, h, w, use_y, v);
  FRTM_CHECK_LAUNCH("stencil_apply");
  return FRTM_OK;
}
