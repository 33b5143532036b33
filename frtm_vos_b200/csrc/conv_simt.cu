// fp32 implicit-GEMM convolution on CUDA cores (NHWC).  This is the exact-arithmetic conv of the library: it serves
// the awkward shapes (7x7 stem with 3 input channels, 65-channel TSE convs, Cout < 16) and is the on-device
// comparator for the tcgen05 path in conv_tc.cu.  GEMM view:  M = B*Ho*Wo pixels, N = Cout, K = kh * (kw*Cin)
// where, thanks to NHWC, the kw*Cin values of one filter row are contiguous in memory for a given output pixel.
#include "common.cuh"

namespace frtm {

struct ConvArgs {
  const float *x, *w, *bias, *res;
  float *y, *y_nchw;
  int B, H, W, Cin, ldx, ldr, ldy, y_coff, Cout, CoutPad, kh, kw, stride, pad, relu, Ho, Wo, M, rowlen;
};

constexpr int BM = 128, BK = 16, APAD = 4;

template <int BN>
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvArgs a) {
  constexpr int TN = BN / 16;  // outputs per thread along N (4 or 2)
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int t = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  // ---- A-load bookkeeping: this thread fetches k-quad `kq` of rows lm and lm+64 ----
  const int kq = t >> 6;
  int iy0[2], ix0[2];
  int64_t xbase[2];
  bool mval[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int m = m0 + (t & 63) + 64 * i;
    mval[i] = m < a.M;
    int mm = mval[i] ? m : 0;
    int ox = mm % a.Wo;
    int r = mm / a.Wo;
    int oy = r % a.Ho;
    int b = r / a.Ho;
    iy0[i] = oy * a.stride - a.pad;
    ix0[i] = ox * a.stride - a.pad;
    xbase[i] = (int64_t)b * a.H * a.W;
  }
  // ---- B-load bookkeeping ----
  constexpr int BQ = BN / 4;  // float4 per weight row
  const bool bload = t < BK * BQ;
  const int bk = t / BQ, bn = (t % BQ) * 4;

  const int ksteps_row = (a.rowlen + BK - 1) / BK;
  const int nsteps = a.kh * ksteps_row;

  float4 ra[2], rb;
  auto fetch = [&](int step) {
    const int ky = step / ksteps_row;
    const int kk = (step - ky * ksteps_row) * BK + kq * 4;
    const int kx = kk / a.Cin, c = kk - kx * a.Cin;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int iy = iy0[i] + ky, ix = ix0[i] + kx;
      ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mval[i] && kk < a.rowlen && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W)
        ra[i] = *reinterpret_cast<const float4 *>(a.x + (xbase[i] + (int64_t)iy * a.W + ix) * a.ldx + c);
    }
    rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bload) {
      const int kb = (step - ky * ksteps_row) * BK + bk;
      if (kb < a.rowlen && n0 + bn < a.CoutPad)
        rb = *reinterpret_cast<const float4 *>(a.w + ((int64_t)ky * a.rowlen + kb) * a.CoutPad + n0 + bn);
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int lm = (t & 63) + 64 * i;
      As[buf][kq * 4 + 0][lm] = ra[i].x;
      As[buf][kq * 4 + 1][lm] = ra[i].y;
      As[buf][kq * 4 + 2][lm] = ra[i].z;
      As[buf][kq * 4 + 3][lm] = ra[i].w;
    }
    if (bload) *reinterpret_cast<float4 *>(&Bs[buf][bk][bn]) = rb;
  };

  const int tx = t & 15, ty = t >> 4;
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  fetch(0);
  stash(0);
  __syncthreads();
  for (int step = 0; step < nsteps; ++step) {
    const int buf = step & 1;
    if (step + 1 < nsteps) fetch(step + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8 + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[TN];
      if (TN == 4) {
        const float4 b4 = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 4]);
        bv[0] = b4.x; bv[1] = b4.y; bv[2] = b4.z; bv[3] = b4.w;
      } else {
        const float2 b2 = *reinterpret_cast<const float2 *>(&Bs[buf][k][tx * 2]);
        bv[0] = b2.x; bv[1] = b2.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (step + 1 < nsteps) {
      stash(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue: bias + residual + relu, NHWC (and optional NCHW) stores ----
  const int nb = n0 + tx * TN;
  float bsv[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) bsv[j] = (a.bias != nullptr && nb + j < a.Cout) ? a.bias[nb + j] : 0.f;
  const int HWo = a.Ho * a.Wo;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= a.M) continue;
    float v[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      v[j] = acc[i][j] + bsv[j];
      if (a.res != nullptr && nb + j < a.Cout) v[j] += a.res[(int64_t)m * a.ldr + nb + j];
      if (a.relu) v[j] = fmaxf(v[j], 0.f);
    }
    if (a.y != nullptr) {
      float *dst = a.y + (int64_t)m * a.ldy + a.y_coff + nb;
#pragma unroll
      for (int j = 0; j < TN; ++j)
        if (nb + j < a.Cout) dst[j] = v[j];
    }
    if (a.y_nchw != nullptr) {
      const int b = m / HWo, pix = m - b * HWo;
#pragma unroll
      for (int j = 0; j < TN; ++j)
        if (nb + j < a.Cout) a.y_nchw[((int64_t)b * a.Cout + nb + j) * HWo + pix] = v[j];
    }
  }
}

// Final conv of the refinement network: 3x3, C -> 1.  One warp per output pixel quad would be overkill; each thread
// owns one pixel and streams the C channels of its 9 neighbours as float4 (NHWC keeps them contiguous).
__global__ void __launch_bounds__(256) conv3x3_to1_kernel(const float *__restrict__ x, int B, int H, int W, int C,
                                                          const float *__restrict__ w, const float *__restrict__ bias,
                                                          float *__restrict__ y) {
  extern __shared__ float ws[];  // [9][C]
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * H * W;
  if (idx >= total) return;
  const int xw = (int)(idx % W);
  const int yh = (int)((idx / W) % H);
  const int64_t b = idx / ((int64_t)W * H);
  float acc = 0.f;
  for (int dy = -1; dy <= 1; ++dy) {
    const int iy = yh + dy;
    if (iy < 0 || iy >= H) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      const int ix = xw + dx;
      if (ix < 0 || ix >= W) continue;
      const float4 *px = reinterpret_cast<const float4 *>(x + ((b * H + iy) * W + ix) * C);
      const float *wt = ws + ((dy + 1) * 3 + (dx + 1)) * C;
      for (int c4 = 0; c4 < C / 4; ++c4) {
        const float4 v = px[c4];
        acc = fmaf(v.x, wt[c4 * 4 + 0], acc);
        acc = fmaf(v.y, wt[c4 * 4 + 1], acc);
        acc = fmaf(v.z, wt[c4 * 4 + 2], acc);
        acc = fmaf(v.w, wt[c4 * 4 + 3], acc);
      }
    }
  }
  y[idx] = acc + (bias ? bias[0] : 0.f);
}

}  // namespace frtm

using namespace frtm;

extern "C" int frtm_conv2d_nhwc(const float *x, int B, int H, int W, int Cin, int ldx, const float *w, const float *bias,
                                const float *res, int ldr, float *y, int ldy, int y_coff, float *y_nchw, int Cout,
                                int kh, int kw, int stride, int pad, int relu, void *stream) {
  FRTM_REQUIRE(x && w && (y || y_nchw), "conv2d: null pointer");
  FRTM_REQUIRE(B > 0 && H > 0 && W > 0 && Cout > 0 && kh > 0 && kw > 0 && stride > 0, "conv2d: bad shape");
  FRTM_REQUIRE(Cin > 0 && Cin % 4 == 0 && ldx % 4 == 0 && ldx >= Cin, "conv2d: Cin (%d) / ldx (%d) must be multiples of 4", Cin, ldx);
  FRTM_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
               "conv2d: x and w must be 16-byte aligned");
  ConvArgs a;
  a.x = x; a.w = w; a.bias = bias; a.res = res; a.y = y; a.y_nchw = y_nchw;
  a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.ldx = ldx; a.ldr = ldr; a.ldy = ldy; a.y_coff = y_coff;
  a.Cout = Cout; a.CoutPad = (Cout + 3) / 4 * 4; a.kh = kh; a.kw = kw; a.stride = stride; a.pad = pad; a.relu = relu;
  a.Ho = (H + 2 * pad - kh) / stride + 1;
  a.Wo = (W + 2 * pad - kw) / stride + 1;
  FRTM_REQUIRE(a.Ho > 0 && a.Wo > 0, "conv2d: empty output");
  const int64_t M = (int64_t)B * a.Ho * a.Wo;
  FRTM_REQUIRE(M < (1ll << 31), "conv2d: too many output pixels");
  a.M = (int)M;
  a.rowlen = kw * Cin;
  cudaStream_t st = (cudaStream_t)stream;
  if (Cout > 32) {
    dim3 grid(cdiv(a.M, BM), cdiv(Cout, 64));
    conv_simt_kernel<64><<<grid, 256, 0, st>>>(a);
  } else {
    dim3 grid(cdiv(a.M, BM), 1);
    conv_simt_kernel<32><<<grid, 256, 0, st>>>(a);
  }
  FRTM_CHECK_LAUNCH("conv2d_nhwc");
  return FRTM_OK;
}

extern "C" int frtm_conv3x3_to1_nhwc(const float *x, int B, int H, int W, int C, const float *w, const float *bias,
                                     float *y, void *stream) {
  FRTM_REQUIRE(x && w && y, "conv3x3_to1: null pointer");
  FRTM_REQUIRE(C % 4 == 0 && C <= 1024, "conv3x3_to1: C must be a multiple of 4");
  const int64_t total = (int64_t)B * H * W;
  conv3x3_to1_kernel<<<cdiv(total, 256), 256, 9 * C * sizeof(float), (cudaStream_t)stream>>>(x, B, H, W, C, w, bias, y);
  FRTM_CHECK_LAUNCH("conv3x3_to1");
  return FRTM_OK;
}
