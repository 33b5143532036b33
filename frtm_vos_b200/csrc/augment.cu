// First-frame augmentation rendering on the device: affine warp (bicubic / nearest, constant-0
// border), per-channel 2-D filter, alpha paste.  B200-native counterpart of the reference's only native component,
// lib/_npp/nppig.cpp (nppiWarpAffine per channel) plus the torch conv2d / elementwise ops around it
// (model/augmenter.py:342-396).  Geometry follows cv2.warpAffine: `M` maps source -> destination, the kernel samples the
// source at M^-1 (x, y); bicubic uses the Keys kernel with a = -0.75 like OpenCV, evaluated at the exact source
// coordinate (OpenCV quantises it to 1/32 px), taps outside the image read 0.
#include "common.cuh"

namespace frtm {

__device__ __forceinline__ void cubic_weights(float t, float w[4]) {
  const float a = -0.75f;
  w[0] = ((a * (t + 1.f) - 5.f * a) * (t + 1.f) + 8.f * a) * (t + 1.f) - 4.f * a;
  w[1] = ((a + 2.f) * t - (a + 3.f)) * t * t + 1.f;
  w[2] = ((a + 2.f) * (1.f - t) - (a + 3.f)) * (1.f - t) * (1.f - t) + 1.f;
  w[3] = 1.f - w[0] - w[1] - w[2];
}

template <typename TI>
__global__ void warp_affine_kernel(const TI *__restrict__ src, int C, int H, int W, float *__restrict__ dst_f,
                                   uint8_t *__restrict__ dst_u8, int Ho, int Wo, float m00, float m01, float m02, float m10,
                                   float m11, float m12, int nearest, float lo, float hi) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= Wo) return;
  const float sx = m00 * x + m01 * y + m02, sy = m10 * x + m11 * y + m12;   // inverse map: destination -> source
  for (int c = 0; c < C; ++c) {
    const TI *s = src + (int64_t)c * H * W;
    float v = 0.f;
    if (nearest) {
      const int ix = (int)floorf(sx + 0.5f), iy = (int)floorf(sy + 0.5f);
      if (ix >= 0 && ix < W && iy >= 0 && iy < H) v = (float)s[(int64_t)iy * W + ix];
    } else {
      const float fx = floorf(sx), fy = floorf(sy);
      float wx[4], wy[4];
      cubic_weights(sx - fx, wx);
      cubic_weights(sy - fy, wy);
      const int ix = (int)fx - 1, iy = (int)fy - 1;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int yy = iy + j;
        if (yy < 0 || yy >= H) continue;
        float row = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int xx = ix + i;
          if (xx >= 0 && xx < W) row = fmaf(wx[i], (float)s[(int64_t)yy * W + xx], row);
        }
        v = fmaf(wy[j], row, v);
      }
      v = fminf(fmaxf(v, lo), hi);
    }
    const int64_t o = ((int64_t)c * Ho + y) * Wo + x;
    if (dst_f) dst_f[o] = v;
    if (dst_u8) dst_u8[o] = (uint8_t)v;
  }
}

// dst[c][y][x] = sum_{j,i} k[j][i] src[c][y + j - kh/2][x + i - kw/2]   (zero padding, cross-correlation like F.conv2d)
__global__ void filter2d_kernel(const float *__restrict__ src, int C, int H, int W, const float *__restrict__ k, int kh, int kw,
                                float *__restrict__ dst) {
  extern __shared__ float ks[];
  for (int i = threadIdx.x; i < kh * kw; i += blockDim.x) ks[i] = k[i];
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, c = blockIdx.z;
  if (x >= W) return;
  const float *s = src + (int64_t)c * H * W;
  float acc = 0.f;
  for (int j = 0; j < kh; ++j) {
    const int yy = y + j - kh / 2;
    if (yy < 0 || yy >= H) continue;
    for (int i = 0; i < kw; ++i) {
      const int xx = x + i - kw / 2;
      if (xx >= 0 && xx < W) acc = fmaf(ks[j * kw + i], s[(int64_t)yy * W + xx], acc);
    }
  }
  dst[((int64_t)c * H + y) * W + x] = acc;
}

// out = uint8( rgb * a + canvas * (1 - a) ),  a = rgba[3] / 255      (augmenter.py:391-394)
__global__ void alpha_paste_kernel(const float *__restrict__ rgba, const float *__restrict__ canvas, int64_t HW,
                                   uint8_t *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HW) return;
  const float a = rgba[3 * HW + i] / 255.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = rgba[c * HW + i] * a + canvas[c * HW + i] * (1.f - a);
    out[c * HW + i] = (uint8_t)v;
  }
}

// Nearest-neighbour warp of a uint8 mask reproducing cv2.warpAffine(INTER_NEAREST) bit for bit: OpenCV inverts the
// matrix in double precision, quantises the source coordinate to 1/1024 px per term (cvRound = round-half-even) and
// rounds with +512 >> 10.  Also counts the pixels equal to `count_value` (the validity test of augmenter.py:453-471).
__global__ void warp_mask_cv_kernel(const uint8_t *__restrict__ src, int H, int W, uint8_t *__restrict__ dst, int Ho, int Wo,
                                    double i00, double i01, double i02, double i10, double i11, double i12, int count_value,
                                    int *__restrict__ count) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  bool hit = false;
  if (x < Wo) {
    const int adx = __double2int_rn(i00 * x * 1024.0), ady = __double2int_rn(i10 * x * 1024.0);
    const int X0 = __double2int_rn((i01 * y + i02) * 1024.0) + 512, Y0 = __double2int_rn((i11 * y + i12) * 1024.0) + 512;
    const int sx = (X0 + adx) >> 10, sy = (Y0 + ady) >> 10;
    uint8_t v = 0;
    if (sx >= 0 && sx < W && sy >= 0 && sy < H) v = src[(int64_t)sy * W + sx];
    dst[(int64_t)y * Wo + x] = v;
    hit = v == count_value;
  }
  const unsigned b = __ballot_sync(0xffffffffu, hit);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, __popc(b));
}

// All candidate masks of an augmentation round in ONE launch: blockIdx.z selects the transform (inverse matrices as kernel
// arguments, no upload), same arithmetic as warp_mask_cv_kernel.
constexpr int WARP_BATCH = 32;
struct WarpBatch {
  double inv[WARP_BATCH][6];
};
__global__ void warp_mask_cv_batch_kernel(const uint8_t *__restrict__ src, int H, int W, uint8_t *__restrict__ dst, int Ho, int Wo,
                                          const __grid_constant__ WarpBatch B, int count_value, int *__restrict__ counts) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, k = blockIdx.z;
  const double i00 = B.inv[k][0], i01 = B.inv[k][1], i02 = B.inv[k][2], i10 = B.inv[k][3], i11 = B.inv[k][4], i12 = B.inv[k][5];
  bool hit = false;
  if (x < Wo) {
    const int adx = __double2int_rn(i00 * x * 1024.0), ady = __double2int_rn(i10 * x * 1024.0);
    const int X0 = __double2int_rn((i01 * y + i02) * 1024.0) + 512, Y0 = __double2int_rn((i11 * y + i12) * 1024.0) + 512;
    const int sx = (X0 + adx) >> 10, sy = (Y0 + ady) >> 10;
    uint8_t v = 0;
    if (sx >= 0 && sx < W && sy >= 0 && sy < H) v = src[(int64_t)sy * W + sx];
    dst[((int64_t)k * Ho + y) * Wo + x] = v;
    hit = v == count_value;
  }
  const unsigned b = __ballot_sync(0xffffffffu, hit);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(counts + k, __popc(b));
}

}  // namespace frtm

using namespace frtm;

// cv::warpAffine's own inversion of a 2x3 matrix (imgwarp.cpp), in double
static void cv_invert_affine(const double *M_in, double *M) {
  for (int i = 0; i < 6; ++i) M[i] = M_in[i];
  double D = M[0] * M[4] - M[1] * M[3];
  D = D != 0 ? 1. / D : 0;
  const double A11 = M[4] * D, A22 = M[0] * D;
  M[0] = A11; M[1] *= -D; M[3] *= -D; M[4] = A22;
  const double b1 = -M[0] * M[2] - M[1] * M[5], b2 = -M[3] * M[2] - M[4] * M[5];
  M[2] = b1; M[5] = b2;
}

extern "C" int frtm_warp_mask_nearest_batch(const uint8_t *src, int H, int W, uint8_t *dst, int Ho, int Wo, int n,
                                            const double *M_host, int count_value, int *counts, void *stream) {
  FRTM_REQUIRE(src && dst && M_host && counts && n >= 1 && Ho >= 1 && Ho <= 65535, "warp_mask_nearest_batch: bad arguments");
  for (int k0 = 0; k0 < n; k0 += WARP_BATCH) {
    const int nb = n - k0 < WARP_BATCH ? n - k0 : WARP_BATCH;
    WarpBatch B;
    for (int k = 0; k < WARP_BATCH; ++k) cv_invert_affine(M_host + 6 * (k0 + (k < nb ? k : 0)), B.inv[k]);
    warp_mask_cv_batch_kernel<<<dim3(cdiv(Wo, 128), Ho, nb), 128, 0, (cudaStream_t)stream>>>(
        src, H, W, dst + (int64_t)k0 * Ho * Wo, Ho, Wo, B, count_value, counts + k0);
    FRTM_CHECK_LAUNCH("warp_mask_nearest_batch");
  }
  return FRTM_OK;
}

static bool invert_affine(const double *M, float *inv) {
  const double det = M[0] * M[4] - M[1] * M[3];
  if (det == 0.0) return false;
  const double a = M[4] / det, b = -M[1] / det, c = -M[3] / det, d = M[0] / det;
  inv[0] = (float)a; inv[1] = (float)b; inv[2] = (float)(-(a * M[2] + b * M[5]));
  inv[3] = (float)c; inv[4] = (float)d; inv[5] = (float)(-(c * M[2] + d * M[5]));
  return true;
}

extern "C" int frtm_warp_affine(const void *src, int src_is_u8, int C, int H, int W, float *dst_f32, uint8_t *dst_u8, int Ho,
                                int Wo, const double *M_host, int nearest, float clamp_lo, float clamp_hi, void *stream) {
  FRTM_REQUIRE(src && (dst_f32 || dst_u8) && M_host && C > 0, "warp_affine: bad arguments");
  float inv[6];
  FRTM_REQUIRE(invert_affine(M_host, inv), "warp_affine: singular transform");
  dim3 grid(cdiv(Wo, 128), Ho);
  cudaStream_t st = (cudaStream_t)stream;
  if (src_is_u8)
    warp_affine_kernel<uint8_t><<<grid, 128, 0, st>>>((const uint8_t *)src, C, H, W, dst_f32, dst_u8, Ho, Wo, inv[0], inv[1], inv[2],
                                                      inv[3], inv[4], inv[5], nearest, clamp_lo, clamp_hi);
  else
    warp_affine_kernel<float><<<grid, 128, 0, st>>>((const float *)src, C, H, W, dst_f32, dst_u8, Ho, Wo, inv[0], inv[1], inv[2],
                                                    inv[3], inv[4], inv[5], nearest, clamp_lo, clamp_hi);
  FRTM_CHECK_LAUNCH("warp_affine");
  return FRTM_OK;
}

extern "C" int frtm_filter2d(const float *src, int C, int H, int W, const float *kernel, int kh, int kw, float *dst, void *stream) {
  FRTM_REQUIRE(src && kernel && dst && kh > 0 && kw > 0 && kh * kw <= 4096, "filter2d: bad arguments");
  filter2d_kernel<<<dim3(cdiv(W, 128), H, C), 128, kh * kw * sizeof(float), (cudaStream_t)stream>>>(src, C, H, W, kernel, kh, kw, dst);
  FRTM_CHECK_LAUNCH("filter2d");
  return FRTM_OK;
}

extern "C" int frtm_alpha_paste(const float *rgba, const float *canvas, int H, int W, uint8_t *out, void *stream) {
  FRTM_REQUIRE(rgba && canvas && out, "alpha_paste: null pointer");
  const int64_t HW = (int64_t)H * W;
  alpha_paste_kernel<<<cdiv(HW, 256), 256, 0, (cudaStream_t)stream>>>(rgba, canvas, HW, out);
  FRTM_CHECK_LAUNCH("alpha_paste");
  return FRTM_OK;
}

extern "C" int frtm_warp_mask_nearest(const uint8_t *src, int H, int W, uint8_t *dst, int Ho, int Wo, const double *M_host,
                                      int count_value, int *count, void *stream) {
  FRTM_REQUIRE(src && dst && M_host && count, "warp_mask_nearest: null pointer");
  // cv::warpAffine's own inversion (imgwarp.cpp), in double
  double M[6];
  for (int i = 0; i < 6; ++i) M[i] = M_host[i];
  double D = M[0] * M[4] - M[1] * M[3];
  D = D != 0 ? 1. / D : 0;
  const double A11 = M[4] * D, A22 = M[0] * D;
  M[0] = A11; M[1] *= -D; M[3] *= -D; M[4] = A22;
  const double b1 = -M[0] * M[2] - M[1] * M[5], b2 = -M[3] * M[2] - M[4] * M[5];
  M[2] = b1; M[5] = b2;
  warp_mask_cv_kernel<<<dim3(cdiv(Wo, 128), Ho), 128, 0, (cudaStream_t)stream>>>(src, H, W, dst, Ho, Wo, M[0], M[1], M[2], M[3],
                                                                                 M[4], M[5], count_value, count);
  FRTM_CHECK_LAUNCH("warp_mask_nearest");
  return FRTM_OK;
}
