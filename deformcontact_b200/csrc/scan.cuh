// Device-wide exclusive scan of uint32 arrays (three-phase, deterministic) shared by the CSR
// builder and the neighbour-table compaction.
#pragma once
#include "common.cuh"

namespace dcb {
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o) v += t;
  }
  return v;
}

// exclusive scan of `total` per-thread values over a 256-thread block; returns prefix, sets block total
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total, uint32_t* sm /*>=9*/) {
  uint32_t inc = warp_incl_scan(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 31) sm[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t s = (l < SCAN_THREADS / 32) ? sm[l] : 0;
    uint32_t si = warp_incl_scan(s);
    if (l < SCAN_THREADS / 32) sm[l] = si - s;
    if (l == SCAN_THREADS / 32 - 1) sm[SCAN_THREADS / 32] = si;
  }
  __syncthreads();
  uint32_t pre = sm[w] + inc - v;
  *total = sm[SCAN_THREADS / 32];
  __syncthreads();
  return pre;
}

static __global__ void scan_reduce_kernel(const uint32_t* __restrict__ in, int64_t L, uint32_t* __restrict__ blocksum) {
  __shared__ uint32_t sm[16];
  int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j)
    if (base + j < L) s += in[base + j];
  uint32_t tot;
  block_excl_scan(s, &tot, sm);
  if (threadIdx.x == 0) blocksum[blockIdx.x] = tot;
}

static __global__ void scan_blocksums_kernel(uint32_t* __restrict__ blocksum, int64_t nb) {
  __shared__ uint32_t sm[16];
  __shared__ uint32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < nb; base += SCAN_THREADS) {
    int64_t i = base + threadIdx.x;
    uint32_t v = i < nb ? blocksum[i] : 0;
    uint32_t tot;
    uint32_t pre = block_excl_scan(v, &tot, sm);
    uint32_t carry = carry_s;
    if (i < nb) blocksum[i] = carry + pre;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
}

static __global__ void scan_apply_kernel(uint32_t* __restrict__ data, int64_t L, const uint32_t* __restrict__ blocksum) {
  __shared__ uint32_t sm[16];
  int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    v[j] = (base + j < L) ? data[base + j] : 0;
    s += v[j];
  }
  uint32_t tot;
  uint32_t pre = block_excl_scan(s, &tot, sm) + blocksum[blockIdx.x];
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    if (base + j < L) data[base + j] = pre;
    pre += v[j];
  }
}


static inline int64_t scan_num_blocks(int64_t L) { return cdiv(L > 0 ? L : 1, SCAN_CHUNK); }
// in-place exclusive scan of data[0..L); blocksum must hold scan_num_blocks(L) uint32
static inline int exclusive_scan_u32(uint32_t* data, int64_t L, uint32_t* blocksum, cudaStream_t st) {
  if (L <= 0) return DC_OK;
  int64_t nb = scan_num_blocks(L);
  scan_reduce_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(data, L, blocksum);
  scan_blocksums_kernel<<<1, SCAN_THREADS, 0, st>>>(blocksum, nb);
  scan_apply_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(data, L, blocksum);
  DC_LAUNCHED(3);
  return DC_OK;
}
}  // namespace dcb
