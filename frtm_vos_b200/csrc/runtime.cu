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
