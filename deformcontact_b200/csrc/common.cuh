// Shared helpers for libdcb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/dcb200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libdcb200 is written for sm_100a only"
#endif

namespace dcb {

void set_error(const char* fmt, ...);

#define DC_REQUIRE(cond, code, ...)        \
  do {                                     \
    if (!(cond)) {                         \
      dcb::set_error(__VA_ARGS__);         \
      return (code);                       \
    }                                      \
  } while (0)

#define DC_CUDA(call)                                                                     \
  do {                                                                                    \
    cudaError_t _e = (call);                                                              \
    if (_e != cudaSuccess) {                                                              \
      dcb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return DC_ECUDA;                                                                    \
    }                                                                                     \
  } while (0)

void count_launches(int n);
// every kernel launch site is followed by DC_LAUNCHED(n) (n = launches since the last check)
#define DC_LAUNCHED(n)                  \
  do {                                  \
    dcb::count_launches(n);             \
    DC_CUDA(cudaPeekAtLastError());     \
  } while (0)
#define DC_LAUNCH_CHECK() DC_LAUNCHED(1)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// carve aligned sub-buffers out of a caller-provided workspace
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* p = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
  size_t used() const { return align_up(off, 256); }
};

constexpr int kSMs = 148;   // B200; device code only (prefetch-distance heuristics).  Host launch sites use sm_count().

// SM count of the current device (queried once per device, cached).
int sm_count();

// cudaFuncSetAttribute is per device: a once-flag per (call site, device) instead of a process-wide bool, so a process
// that drives a second GPU sets the >48 KB dynamic shared memory / carve-out attributes there too.
struct DeviceOnce {
  unsigned long long mask = 0;   // bit d set = done on device d (benign if raced: the attribute calls are idempotent)
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    const unsigned long long bit = 1ull << (d & 63);
    if (mask & bit) return false;
    mask |= bit;
    return true;
  }
};

}  // namespace dcb
