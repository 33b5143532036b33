// Types shared by the GN/CG operator kernels (target_model.cu: CUDA-core kernels, gn_apply_tc.cu: tcgen05 kernel).
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>

namespace frtm {

// Pointer table for object-batched launches: rows {samples, stencil, uty, weights, filt, cg_state, gate_count,
// samples_split} x n_obj.
constexpr int GA_TABLE_ROWS = 8;
struct GaArgs {
  const float *X, *S, *T, *sw, *pvec;   // single-object form (table == nullptr)
  const __half *XS;                     // split-fp16 tile image of X (frtm_split_samples) or null
  const long long *table;               // batched form: blockIdx.y = object
  float *partial;                       // [n_obj][cap][c*9]
  float *dbg;                           // optional debug dump of one CTA's score / v maps (tests only), else null
  int n_obj, cap, c, h, w, use_y;
};

// Operator image of one memory sample, streamed by the tensor-core operator kernel (gn_apply_tc.cu):
//   [ntiles][hi|lo][c rows][64 px] fp16   split tile image of the features, 16*x = hi + lo, rows in the 128-byte
//                                         swizzled shared-memory layout of a UMMA operand, pixels beyond hw zero
//   [nchunks][10][256 px] fp32            stencil taps 0..8 and U^T w^2 y of 256 consecutive pixels per chunk
// ntiles is even.  Same bytes per element as the fp32 arrays it mirrors (plus padding of the last tile / chunk).
constexpr int GC_TILE = 64;
constexpr int GC_CHUNK_PX = 256;
constexpr int GC_CHUNK_BYTES = 10 * GC_CHUNK_PX * 4;
constexpr float GC_ACT_SCALE = 16.f;
__host__ __device__ inline int gc_ntiles(int hw) { return ((hw + 2 * GC_TILE - 1) / (2 * GC_TILE)) * 2; }
__host__ __device__ inline int gc_nchunks(int hw) { return (hw + GC_CHUNK_PX - 1) / GC_CHUNK_PX; }
__host__ __device__ inline int64_t gc_sample_bytes(int c, int hw) {
  return (int64_t)gc_ntiles(hw) * 2 * c * GC_TILE * 2 + (int64_t)gc_nchunks(hw) * GC_CHUNK_BYTES;
}
__host__ __device__ inline int64_t gc_sample_items(int c, int hw) {
  return (int64_t)gc_ntiles(hw) * c * 8 + (int64_t)gc_nchunks(hw) * 10 * (GC_CHUNK_PX / 4);
}

struct GcParams {
  int ntiles, nchunks, tile_bytes, slots;
  int64_t image_bytes;
};

// power of two that brings amax into [2^9, 2^10)
__device__ __forceinline__ float pow2_scale(float amax) {
  if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
  int e;
  frexpf(amax, &e);
  e = max(min(10 - e, 100), -100);
  return ldexpf(1.f, e);
}

// Operator kernels over the operator images.  Both take one CTA per (memory slot, object), write the sample's gradient
// partial and run the reduction + CG vector step in the tail of the launch (gc_fused_tail below):
//   gn_apply_mma_*  single pass over the image (sliding window, warp-level mma.sync tiles) — the default
//   gn_apply_tc_*   two passes (tcgen05 / TMEM, producer / issuer / drain roles) — kept for shapes the first does not take
bool gn_apply_tc_supported(int c, int h, int w);
bool gn_apply_mma_supported(int c, int h, int w);
struct GcFuse;
int gn_apply_tc_launch(const GaArgs &a, const GcFuse &fuse, cudaStream_t st);
// Work list of an update, built once per update by gn_items_build (the sample weights do not change between its operator
// applications): hdr[0] = U (active samples of all objects), hdr[1 + o] = first item of object o (hdr[1 + n_obj] = U),
// items[u] = (object << 16) | slot, object-major, slots ascending.
struct ClList {
  int *hdr;
  uint32_t *items;
};
int gn_items_build(const GaArgs &a, ClList list, cudaStream_t st);

// Workspace of the list-driven operator kernels (floats): see gn_list_workspace() in target_model.cu.
struct GnListWs {
  ClList list;
  float *rows;        // sliding-window kernel: one row per unit [n_obj*cap + 160][n]; cluster kernel: [n_obj][160][n]
  float *gsum;        // [n_obj][ngrp_max][n] group sums (sliding-window kernel)
  int *counters;      // [n_obj][1 + ngrp_max] tickets, zero between launches
  int *tickets;       // [n_obj] (cluster kernel)
  int ngrp_max;
};
int64_t gn_list_workspace_bytes(int n_obj, int cap, int c);
GnListWs gn_list_workspace(float *base, int n_obj, int cap, int c);

// gn_apply_cl_*: one sample per thread-block cluster, resident on chip through its three phases (gn_apply_cl.cu); persistent
// clusters over a work list built once per update (prepare), own workspace (rows per (object, cluster), tickets, the list)
bool gn_apply_cl_supported(int c, int h, int w);
int gn_apply_cl_launch(const GaArgs &a, const GcFuse &fuse, const GnListWs &ws, cudaStream_t st);
int gn_apply_mma_launch(const GaArgs &a, const GcFuse &fuse, const GnListWs &ws, cudaStream_t st);

// One work item of the image build.  Items [0, ntiles*c*8): 8 consecutive pixels of channel row r in tile j -> one
// 16-byte chunk in each plane.  Remaining items: 4 consecutive pixels of one stencil / uty row of one chunk.
__device__ __forceinline__ void gc_image_item(const float *__restrict__ x, const float *__restrict__ stencil,
                                              const float *__restrict__ uty, uint8_t *__restrict__ img, int c, int hw,
                                              int64_t item) {
  const int ntiles = gc_ntiles(hw);
  const int64_t nx = (int64_t)ntiles * c * 8;
  if (item < nx) {
    const int ch8 = (int)(item & 7);
    const int64_t jr = item >> 3;
    const int r = (int)(jr % c);
    const int j = (int)(jr / c);
    const int q0 = j * GC_TILE + ch8 * 8;
    __align__(16) __half hh[8];
    __align__(16) __half ll[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float v = (q0 + e < hw) ? x[(int64_t)r * hw + q0 + e] * GC_ACT_SCALE : 0.f;
      hh[e] = __float2half_rn(v);
      ll[e] = __float2half_rn(v - __half2float(hh[e]));
    }
    __half *dst = reinterpret_cast<__half *>(img);
    const int64_t off = ((int64_t)(j * 2) * c + r) * GC_TILE + ((ch8 ^ (r & 7)) << 3);
    *reinterpret_cast<uint4 *>(dst + off) = *reinterpret_cast<const uint4 *>(hh);
    *reinterpret_cast<uint4 *>(dst + off + (int64_t)c * GC_TILE) = *reinterpret_cast<const uint4 *>(ll);
  } else {
    const int64_t e = item - nx;                       // (chunk m, row t, group of 4 pixels)
    const int g4 = (int)(e % (GC_CHUNK_PX / 4));
    const int t = (int)((e / (GC_CHUNK_PX / 4)) % 10);
    const int m = (int)(e / (10 * (GC_CHUNK_PX / 4)));
    const int q0 = m * GC_CHUNK_PX + g4 * 4;
    const float *src = t < 9 ? stencil + (int64_t)t * hw : uty;
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = (q0 + u < hw) ? src[q0 + u] : 0.f;
    float *dst = reinterpret_cast<float *>(img + (int64_t)ntiles * 2 * c * GC_TILE * 2) + ((int64_t)m * 10 + t) * GC_CHUNK_PX + g4 * 4;
    *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[3]);
  }
}


// ---------------------------------------------------------------- CG vector step (filter-only problem) -------------
// cg_state layout: p[n] | r_prev[n] | rho | has_p | pad | pad           (persists across updates)
// work layout    : b/r[n] | x[n] | q[n]
struct CgVec {
  float *f, *p, *rprev, *rho, *hasp;  // persistent
  float *r, *x, *q;                   // per-run scratch
  const float *partial;               // [cap][n] per-sample gradient partials
  int n, cap;
  float reg2, minv, forget;
};

// What the operator kernel needs to run the CG vector step itself when its last CTA of an object retires.
constexpr int GC_RGROUP = 8;      // samples per first-level reduction group
struct GcFuse {
  CgVec cg;            // template: pointers of the single-object form, or per-object offsets applied from the table
  const int *gate;     // single-object gate (table form: row 6)
  int *counters;       // [n_obj][1 + ngroups] zero-initialised tickets (object, then its groups of GC_RGROUP samples);
                       // whoever draws the last ticket resets the counter
  float *gsum;         // [n_obj][ngroups][n] group sums of the per-sample partials
  int mode, min_px, enabled;
};

// One CG vector step by the whole CTA (NT threads, EPT = ceil(n / NT) elements per thread):
// mode 0: finish RHS ( r = b = -(sum partial + reg^2 f) ), x = 0, apply the forgetting factor, then first direction.
// mode 1: finish A p ( q = sum partial + reg^2 p ), alpha step, residual update, next direction.
// mode 2: like mode 1 but last CG iteration of the GN step: no residual update, no new direction, f += x.
// Same recurrences as cg_vector_kernel (optimizer.py:98-153); partials are read with ld.global.cg (written by other SMs).
template <int NT, int EPT>
__device__ __forceinline__ void cg_vector_step_cta(const CgVec &s, int mode, float *red) {
  const int tid = threadIdx.x;
  float g[EPT], r[EPT], p[EPT], rp[EPT], q[EPT];
  {   // fixed summation order; all EPT x 8 loads of a round are independent and in flight together
    float g8[EPT][8];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      r[e] = 0.f; p[e] = 0.f; rp[e] = 0.f; q[e] = 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) g8[e][u] = 0.f;
    }
    int k = 0;
    for (; k + 8 <= s.cap; k += 8) {
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const int t = tid + e * NT;
        if (t < s.n) {
#pragma unroll
          for (int u = 0; u < 8; ++u) g8[e][u] += __ldcg(s.partial + (int64_t)(k + u) * s.n + t);
        }
      }
    }
    for (; k < s.cap; ++k) {
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const int t = tid + e * NT;
        if (t < s.n) g8[e][0] += __ldcg(s.partial + (int64_t)k * s.n + t);
      }
    }
#pragma unroll
    for (int e = 0; e < EPT; ++e)
      g[e] = ((g8[e][0] + g8[e][1]) + (g8[e][2] + g8[e][3])) + ((g8[e][4] + g8[e][5]) + (g8[e][6] + g8[e][7]));
  }
  float rho = *s.rho;
  // direction_forget_factor == 0: the CG state is reset at the start of every run (optimizer.py:93-103)
  const bool hasp = *s.hasp != 0.f && !(mode == 0 && s.forget == 0.f);
  if (mode == 0) {
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int t = tid + e * NT;
      if (t < s.n) {
        r[e] = -(g[e] + s.reg2 * s.f[t]);
        s.x[t] = 0.f;
        p[e] = hasp ? s.p[t] : 0.f;
        rp[e] = hasp ? s.rprev[t] : 0.f;
      }
    }
    if (hasp) rho = rho / s.forget;                    // optimizer.py:104-105
  } else {
    float pq_l = 0.f;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int t = tid + e * NT;
      if (t < s.n) {
        p[e] = s.p[t];
        q[e] = g[e] + s.reg2 * p[e];
        r[e] = s.r[t];
        pq_l += p[e] * q[e];
      }
    }
    const float pq = block_sum(pq_l, red);
    const float alpha = rho / pq;                      // standard_alpha, optimizer.py:134
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int t = tid + e * NT;
      if (t < s.n) {
        rp[e] = r[e];                                  // r_prev = r.clone()
        s.rprev[t] = rp[e];
        const float xn = s.x[t] + alpha * p[e];
        s.x[t] = xn;
        if (mode == 1) r[e] = r[e] - alpha * q[e];
        else s.f[t] += xn;                             // theta += step_alpha * delta_x
        s.r[t] = r[e];
      }
    }
    if (mode == 2) return;
  }
  // next direction (optimizer.py:115-128): z = r / diag_M ; rho = <r,z> ; beta = max((rho - <r_prev,z>)/rho1, 0)
  float rz = 0.f, rpz = 0.f;
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    const float z = r[e] * s.minv;
    rz += r[e] * z;
    rpz += rp[e] * z;
  }
  const float rho_new = block_sum(rz, red);
  float beta = 0.f;
  const bool use_beta = (mode != 0 || hasp);
  if (use_beta) {
    const float rho2 = block_sum(rpz, red);
    beta = fmaxf((rho_new - rho2) / rho, 0.f);
  }
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    const int t = tid + e * NT;
    if (t < s.n) {
      const float z = r[e] * s.minv;
      s.p[t] = use_beta ? z + p[e] * beta : z;
      if (mode == 0) s.r[t] = r[e];
    }
  }
  __syncthreads();
  if (tid == 0) {
    *s.rho = rho_new;
    *s.hasp = 1.f;
  }
}

// The per-sample partials are reduced and the CG vector step is run by the operator kernel itself, so an operator
// application is ONE launch instead of two: the CTA that retires last within a group of GC_RGROUP samples sums the
// group's rows (fixed order), the one that completes the last group of an object sums the group rows (fixed order) and
// advances the Polak-Ribiere recurrences.  Tickets are global atomics; every sum has a fixed order -> deterministic.
// Called by every thread of every CTA of an operator launch after its partial row is written.
template <int NT, int EPT>
__device__ __forceinline__ void gc_fused_tail(const GaArgs &a, const GcFuse &F) {
  if (!F.enabled) return;
  __shared__ float red[32];
  __shared__ int s_last;
  const int o = a.table ? blockIdx.y : 0;
  const int n = a.c * 9;
  const int ngrp = (a.cap + GC_RGROUP - 1) / GC_RGROUP;
  const int grp = blockIdx.x / GC_RGROUP;
  const int gsize = min(GC_RGROUP, a.cap - grp * GC_RGROUP);
  int *cnt = F.counters + (int64_t)o * (1 + ngrp);
  __threadfence();                                   // this CTA's partial row is visible before its ticket
  __syncthreads();
  if (threadIdx.x == 0) {
    const int ticket = atomicAdd(cnt + 1 + grp, 1);
    s_last = (ticket == gsize - 1) ? 1 : 0;
    if (s_last) cnt[1 + grp] = 0;                    // every ticket of this group has been drawn: reset for the next launch
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  {   // group sum: rows grp*8 .. +gsize of this object's partials
    const float *part = a.partial + ((int64_t)o * a.cap + (int64_t)grp * GC_RGROUP) * n;
    float *dst = F.gsum + ((int64_t)o * ngrp + grp) * n;
    for (int t = threadIdx.x; t < n; t += NT) {
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
  cg.partial = F.gsum + (int64_t)o * ngrp * n;       // the vector step sums the group rows
  cg.cap = ngrp;
  if (gate && gate[0] < F.min_px) return;
  cg_vector_step_cta<NT, EPT>(cg, F.mode, red);
}

}  // namespace frtm
