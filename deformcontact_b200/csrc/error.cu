#include <stdarg.h>
#include <atomic>
#include "common.cuh"

namespace dcb {
unsigned long long launches();
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::atomic<unsigned long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
unsigned long long launches() { return g_launches.load(std::memory_order_relaxed); }
int sm_count() {
  static std::atomic<int> cache[64];
  int d = 0;
  cudaGetDevice(&d);
  int n = cache[d & 63].load(std::memory_order_relaxed);
  if (n <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || n <= 0) n = kSMs;
    cache[d & 63].store(n, std::memory_order_relaxed);
  }
  return n;
}
}  // namespace dcb

extern "C" int dc_version(void) { return 100; }
extern "C" const char* dc_last_error(void) { return dcb::g_err; }

extern "C" uint64_t dc_launch_count(void) { return dcb::launches(); }
