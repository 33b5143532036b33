// Error string, launch counter, version.
#include "common.cuh"
#include <atomic>

namespace frtm {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace frtm

extern "C" const char *frtm_last_error(void) { return frtm::g_err; }
extern "C" int frtm_version(void) { return 100; }
extern "C" int64_t frtm_launch_count(void) { return frtm::g_launches.load(std::memory_order_relaxed); }
/* A caller that replays launches of this library through a CUDA graph it captured itself reports the kernels of a replay. */
extern "C" int64_t frtm_count_launches(int64_t n) { frtm::count_launch((int)n); return frtm::g_launches.load(std::memory_order_relaxed); }

// Small host->device constant uploads WITHOUT a memcpy (a pageable cudaMemcpy would synchronise the caller with all
// work queued on the stream): the values travel as kernel arguments.
namespace frtm {
struct Vals16 { float f[16]; int i[16]; };
__global__ void fill_values_kernel(float *fdst, int nf, int *idst, int ni, const Vals16 v) {
  const int t = threadIdx.x;
  if (t < nf) fdst[t] = v.f[t];
  if (t < ni) idst[t] = v.i[t];
}
}  // namespace frtm

extern "C" int frtm_fill_small(float *fdst, const float *fvals_host, int nf, int *idst, const int *ivals_host, int ni,
                               void *stream) {
  FRTM_REQUIRE(nf >= 0 && nf <= 16 && ni >= 0 && ni <= 16 && (nf == 0 || (fdst && fvals_host)) && (ni == 0 || (idst && ivals_host)),
               "fill_small: at most 16 floats and 16 ints");
  frtm::Vals16 v;
  for (int k = 0; k < 16; ++k) { v.f[k] = k < nf ? fvals_host[k] : 0.f; v.i[k] = k < ni ? ivals_host[k] : 0; }
  frtm::fill_values_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(fdst, nf, idst, ni, v);
  FRTM_CHECK_LAUNCH("fill_small");
  return FRTM_OK;
}

namespace frtm {
__global__ void fill_u8_kernel(uint8_t *dst, int n, const Vals16 v) {
  if ((int)threadIdx.x < n) dst[threadIdx.x] = (uint8_t)v.i[threadIdx.x];
}
}  // namespace frtm

extern "C" int frtm_fill_u8(uint8_t *dst, const int *vals_host, int n, void *stream) {
  FRTM_REQUIRE(dst && vals_host && n >= 0 && n <= 16, "fill_u8: at most 16 values");
  frtm::Vals16 v;
  for (int k = 0; k < 16; ++k) { v.f[k] = 0.f; v.i[k] = k < n ? vals_host[k] : 0; }
  frtm::fill_u8_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(dst, n, v);
  FRTM_CHECK_LAUNCH("fill_u8");
  return FRTM_OK;
}

namespace frtm {
struct Vals64 { long long v[64]; };
__global__ void fill_i64_kernel(long long *dst, int n, const Vals64 v) {
  const int t = threadIdx.x;
  if (t < n) dst[t] = v.v[t];
}
}  // namespace frtm

extern "C" int frtm_fill_i64(void *dst, const int64_t *vals_host, int n, void *stream) {
  FRTM_REQUIRE(dst && vals_host && n >= 0, "fill_i64: bad arguments");
  for (int o = 0; o < n; o += 64) {
    frtm::Vals64 v;
    const int m = n - o < 64 ? n - o : 64;
    for (int k = 0; k < 64; ++k) v.v[k] = k < m ? (long long)vals_host[o + k] : 0;
    frtm::fill_i64_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(reinterpret_cast<long long *>(dst) + o, m, v);
    FRTM_CHECK_LAUNCH("fill_i64");
  }
  return FRTM_OK;
}
