// Device helpers shared by the brute-force (knn.cu) and the uniform-grid (knn_grid.cu) neighbour searches:
// the (distance bits << 32 | index) key, the bit-exact squared distance, the lane-distributed top-k set.
#pragma once
#include "common.cuh"

namespace dcb {
// brute-force launches of knn.cu (also the fallback of knn_grid.cu); `skip`: device flag, kernels return when *skip != 0
int launch_knn_brute(const float* pos, const int64_t* ptr, int64_t B, int64_t N, int kk, int loop, int32_t* nbr_out,
                     const int* skip, cudaStream_t st);
int launch_radius_brute(const float* pos, const int64_t* ptr, int64_t B, int64_t N, float r2, int cap, int loop,
                        int32_t* nbr_out, int32_t* count_out, const int* skip, cudaStream_t st);
}  // namespace dcb

namespace {
using namespace dcb;

constexpr unsigned long long KEY_INF = ~0ull;

__device__ __forceinline__ float sqdist(float qx, float qy, float qz, float px, float py, float pz) {
  float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ unsigned long long shfl64(unsigned long long v, int src) {
  unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, src);
  unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src);
  return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ unsigned long long shfl_xor64(unsigned long long v, int m) {
  unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, m);
  unsigned hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), m);
  return ((unsigned long long)hi << 32) | lo;
}

template <int SLOTS>
__device__ __forceinline__ void warp_bitonic_sort(unsigned long long (&k)[SLOTS], int lane) {
  constexpr int n = 32 * SLOTS;
#pragma unroll
  for (int size = 2; size <= n; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= 32) {
        const int ss = stride / 32;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
          if ((s & ss) == 0) {
            const bool asc = (((s * 32 + lane) & size) == 0);
            unsigned long long a = k[s], b = k[s | ss];
            if ((a > b) == asc) { k[s] = b; k[s | ss] = a; }
          }
        }
      } else {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
          const bool asc = (((s * 32 + lane) & size) == 0);
          unsigned long long o = shfl_xor64(k[s], stride);
          const bool lower = (lane & stride) == 0;
          const bool keepmin = (lower == asc);
          k[s] = keepmin ? (k[s] < o ? k[s] : o) : (k[s] > o ? k[s] : o);
        }
      }
    }
  }
}

// One candidate (already tested against the 32-bit prefilter by the caller's ballot) set: insert every lane's
// surviving key into query qi's lane-distributed top-k set, in ballot order.
// Warp-wide maximum of 64-bit keys with two 32-bit REDUX operations (high words, then low words among the lanes
// that hold the maximal high word).
__device__ __forceinline__ unsigned long long warp_max64(unsigned long long v) {
  const unsigned hi = __reduce_max_sync(0xffffffffu, (unsigned)(v >> 32));
  const unsigned lo = __reduce_max_sync(0xffffffffu, (unsigned)(v >> 32) == hi ? (unsigned)v : 0u);
  return ((unsigned long long)hi << 32) | lo;
}

// Insert every lane's surviving key into a query's lane-distributed top-k set, in lane order.  Invariant:
// `thresh` is the maximum of the set (KEY_INF placeholders included), so the slot to replace is the one equal to it.
template <int SLOTS>
__device__ __forceinline__ void knn_insert(unsigned long long (&keys)[SLOTS], unsigned long long& thresh, unsigned& thr_hi,
                                           unsigned long long key, int lane) {
  unsigned m = __ballot_sync(0xffffffffu, key < thresh);
  while (m) {
    const int b = __ffs(m) - 1;
    m &= m - 1;
    const unsigned long long ck = shfl64(key, b);
    if (ck < thresh) {
      int lslot = -1;
#pragma unroll
      for (int s = SLOTS - 1; s >= 0; --s)
        if (keys[s] == thresh) lslot = s;
      const int owner = __ffs(__ballot_sync(0xffffffffu, lslot >= 0)) - 1;   // several lanes may hold a KEY_INF placeholder
      unsigned long long lmax = 0ull;
#pragma unroll
      for (int s = 0; s < SLOTS; ++s) {
        if (lane == owner && s == lslot) keys[s] = ck;
        lmax = keys[s] > lmax ? keys[s] : lmax;
      }
      thresh = warp_max64(lmax);
      thr_hi = (unsigned)(thresh >> 32);
    }
  }
}

// Sort a query's candidate set and write its table row: ascending (distance, index), the self match removed when
// loop = 0 (torch_cluster searches k+1 and drops the query itself), -1 padding up to W.
template <int SLOTS>
__device__ __forceinline__ void knn_emit(unsigned long long (&keys)[SLOTS], int64_t q, int kk, int loop, int W,
                                         int32_t* __restrict__ out, int lane) {
  constexpr int CAP = 32 * SLOTS;
  const int ndummy = CAP - kk;
  warp_bitonic_sort<SLOTS>(keys, lane);
  // rank of the self match among the real entries (or CAP if absent / loop)
  int self_rank = CAP;
  bool valid[SLOTS];
  int idx[SLOTS], rank[SLOTS];
  int nvalid = 0;
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    rank[s] = s * 32 + lane - ndummy;
    valid[s] = rank[s] >= 0 && keys[s] != KEY_INF;
    idx[s] = (int)(unsigned)(keys[s] & 0xffffffffull);
    const bool is_self = valid[s] && !loop && (int64_t)idx[s] == q;
    const unsigned sm = __ballot_sync(0xffffffffu, is_self);
    if (sm) self_rank = s * 32 + (__ffs(sm) - 1) - ndummy;
    nvalid += __popc(__ballot_sync(0xffffffffu, valid[s]));
  }
  const int nout = nvalid - (self_rank < CAP ? 1 : 0);
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    if (valid[s] && rank[s] != self_rank) {
      const int p = rank[s] - (rank[s] > self_rank ? 1 : 0);
      if (p < W) out[q * W + p] = idx[s];
    }
  }
  for (int p = nout + lane; p < W; p += 32) out[q * W + p] = -1;
}
}  // namespace
