// K5 — CSR construction: stable LSD radix sort of the edge list by one endpoint.
// Replaces gcn_norm's scatter/deg bookkeeping (PyG gcn_conv.py) and makes every later
// per-receiver reduction a deterministic, atomics-free segmented sum in original edge order.
#include "common.cuh"
#include "scan.cuh"

namespace {
using namespace dcb;

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_MAXBINS = 256;

__global__ void make_keys_kernel(const int64_t* __restrict__ ei, int64_t E, int64_t N, int group_by, int drop,
                                 uint32_t* __restrict__ keys, int32_t* __restrict__ vals) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  int64_t s = ei[e], t = ei[E + e];
  bool bad = (s < 0) | (s >= N) | (t < 0) | (t >= N) | (drop && s == t);
  int64_t k = group_by ? s : t;
  keys[e] = bad ? (uint32_t)N : (uint32_t)k;
  vals[e] = (int32_t)e;
}

__global__ void rs_hist_kernel(const uint32_t* __restrict__ keys, int64_t E, int shift, uint32_t mask, int nbins,
                               uint32_t* __restrict__ counts, int nblocks) {
  __shared__ uint32_t hist[RS_MAXBINS];
  for (int i = threadIdx.x; i < nbins; i += RS_THREADS) hist[i] = 0;
  __syncthreads();
  int64_t start = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    int64_t i = start + r * RS_THREADS + threadIdx.x;
    if (i < E) atomicAdd(&hist[(keys[i] >> shift) & mask], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nbins; i += RS_THREADS) counts[(size_t)i * nblocks + blockIdx.x] = hist[i];
}

__global__ void rs_scatter_kernel(const uint32_t* __restrict__ keys_in, const int32_t* __restrict__ vals_in,
                                  uint32_t* __restrict__ keys_out, int32_t* __restrict__ vals_out, int64_t E,
                                  int shift, uint32_t mask, int nbins, const uint32_t* __restrict__ offsets,
                                  int nblocks) {
  __shared__ uint32_t base[RS_MAXBINS];
  __shared__ uint32_t wcnt[RS_WARPS][RS_MAXBINS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < nbins) base[tid] = offsets[(size_t)tid * nblocks + blockIdx.x];
  int64_t start = (int64_t)blockIdx.x * RS_TILE;
  for (int r = 0; r < RS_ITEMS; ++r) {
    for (int w = 0; w < RS_WARPS; ++w) wcnt[w][tid] = 0;
    __syncthreads();
    int64_t i = start + r * RS_THREADS + tid;
    bool valid = i < E;
    uint32_t key = valid ? keys_in[i] : 0u;
    int32_t val = valid ? vals_in[i] : 0;
    uint32_t d = valid ? ((key >> shift) & mask) : (uint32_t)RS_MAXBINS;
    uint32_t peers = __match_any_sync(0xffffffffu, d);
    uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank == 0) wcnt[warp][d] = __popc(peers);
    __syncthreads();
    if (tid < nbins) {
      uint32_t run = base[tid];
#pragma unroll
      for (int w = 0; w < RS_WARPS; ++w) {
        uint32_t c = wcnt[w][tid];
        wcnt[w][tid] = run;
        run += c;
      }
      base[tid] = run;
    }
    __syncthreads();
    if (valid) {
      uint32_t dst = wcnt[warp][d] + rank;
      keys_out[dst] = key;
      vals_out[dst] = val;
    }
    __syncthreads();
  }
}

__global__ void csr_rowptr_kernel(const uint32_t* __restrict__ keys, int64_t E, int64_t N, int32_t* __restrict__ rowptr) {
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v > N) return;
  int64_t lo = 0, hi = E;  // first position with key >= v
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < (uint32_t)v) lo = mid + 1; else hi = mid;
  }
  rowptr[v] = (int32_t)lo;
}

__global__ void csr_edges_kernel(const int32_t* __restrict__ vals, const int64_t* __restrict__ ei, int64_t E, int group_by,
                                 int32_t* __restrict__ nbr, int32_t* __restrict__ eid) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= E) return;
  int32_t e = vals[i];
  eid[i] = e;
  nbr[i] = (int32_t)ei[group_by ? (E + e) : (int64_t)e];
}

__global__ void deg_inv_sqrt_kernel(const int32_t* __restrict__ rowptr, int64_t N, int add, float* __restrict__ dis) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  int d = rowptr[i + 1] - rowptr[i] + add;
  dis[i] = d > 0 ? __fdiv_rn(1.0f, __fsqrt_rn((float)d)) : 0.0f;
}

struct SortPlan {
  int passes, dbits, nbins;
  int64_t nblocks, L, nscan;
};
SortPlan plan_sort(int64_t N, int64_t E) {
  SortPlan p;
  int bits = 1;
  while ((1ll << bits) <= N) ++bits;  // keys in [0, N]
  p.passes = (bits + 7) / 8;
  p.dbits = (bits + p.passes - 1) / p.passes;
  p.nbins = 1 << p.dbits;
  p.nblocks = cdiv(E > 0 ? E : 1, RS_TILE);
  p.L = (int64_t)p.nbins * p.nblocks;
  p.nscan = cdiv(p.L, SCAN_CHUNK);
  return p;
}
}  // namespace

extern "C" size_t dc_csr_build_workspace_bytes(int64_t N, int64_t E) {
  if (E < 0 || N < 0) return 0;
  SortPlan p = plan_sort(N, E);
  dcb::Carver c(nullptr);
  int64_t e = E > 0 ? E : 1;
  c.take<uint32_t>(e); c.take<uint32_t>(e); c.take<int32_t>(e); c.take<int32_t>(e);
  c.take<uint32_t>(p.L); c.take<uint32_t>(p.nscan);
  return c.used();
}

extern "C" int dc_csr_build(const int64_t* edge_index, int64_t E, int64_t N, int group_by, int drop_self_loops,
                            int32_t* rowptr, int32_t* nbr, int32_t* eid, void* workspace, size_t workspace_bytes,
                            dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(E >= 0 && N >= 0 && N < (1ll << 31) - 1 && E < (1ll << 31) - 1, DC_EINVAL, "csr_build: N=%lld E=%lld out of int32 range",
             (long long)N, (long long)E);
  DC_REQUIRE(rowptr && (E == 0 || (edge_index && nbr && eid)), DC_EINVAL, "csr_build: null pointer");
  DC_REQUIRE(workspace_bytes >= dc_csr_build_workspace_bytes(N, E), DC_EWORKSPACE, "csr_build: workspace %zu < %zu",
             workspace_bytes, dc_csr_build_workspace_bytes(N, E));
  if (E == 0) {
    DC_CUDA(cudaMemsetAsync(rowptr, 0, (N + 1) * sizeof(int32_t), st));
    return DC_OK;
  }
  DC_REQUIRE(workspace, DC_EINVAL, "csr_build: null workspace");
  SortPlan p = plan_sort(N, E);
  dcb::Carver c(workspace);
  uint32_t* k0 = c.take<uint32_t>(E);
  uint32_t* k1 = c.take<uint32_t>(E);
  int32_t* v0 = c.take<int32_t>(E);
  int32_t* v1 = c.take<int32_t>(E);
  uint32_t* counts = c.take<uint32_t>(p.L);
  uint32_t* bsum = c.take<uint32_t>(p.nscan);
  make_keys_kernel<<<(unsigned)cdiv(E, 256), 256, 0, st>>>(edge_index, E, N, group_by, drop_self_loops, k0, v0);
  DC_LAUNCH_CHECK();
  uint32_t mask = (uint32_t)p.nbins - 1u;
  for (int pass = 0; pass < p.passes; ++pass) {
    int shift = pass * p.dbits;
    rs_hist_kernel<<<(unsigned)p.nblocks, RS_THREADS, 0, st>>>(k0, E, shift, mask, p.nbins, counts, (int)p.nblocks);
    if (int rc = exclusive_scan_u32(counts, p.L, bsum, st)) return rc;
    rs_scatter_kernel<<<(unsigned)p.nblocks, RS_THREADS, 0, st>>>(k0, v0, k1, v1, E, shift, mask, p.nbins, counts,
                                                                 (int)p.nblocks);
    DC_LAUNCHED(2);
    uint32_t* tk = k0; k0 = k1; k1 = tk;
    int32_t* tv = v0; v0 = v1; v1 = tv;
  }
  csr_rowptr_kernel<<<(unsigned)cdiv(N + 1, 256), 256, 0, st>>>(k0, E, N, rowptr);
  csr_edges_kernel<<<(unsigned)cdiv(E, 256), 256, 0, st>>>(v0, edge_index, E, group_by, nbr, eid);
  DC_LAUNCHED(2);
  return DC_OK;
}

extern "C" int dc_deg_inv_sqrt(const int32_t* rowptr, int64_t N, int add_self_loop, float* dis, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(N >= 0 && (N == 0 || (rowptr && dis)), DC_EINVAL, "deg_inv_sqrt: bad args");
  if (N == 0) return DC_OK;
  deg_inv_sqrt_kernel<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(rowptr, N, add_self_loop ? 1 : 0, dis);
  DC_LAUNCH_CHECK();
  return DC_OK;
}
