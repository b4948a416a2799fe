// K4 — kNN / radius neighbour search (torch_cluster.knn_graph / radius_graph semantics as called
// at utils/pointcloud_utils.py:10,12).  Tiled brute force: a CTA stages TILE candidate points in
// shared memory; each warp owns QW queries (8 for k <= 31) and evaluates 64 candidates per step (two per
// lane) against all of them with no synchronisation — 2*QW independent distance chains per lane — then
// one warp vote tells whether any pair beat its query's current k-th distance (a 32-bit compare on the
// distance bits); only then the survivors are inserted with warp shuffles into a per-query candidate
// set spread across the lanes.  A warp-shuffle bitonic sort orders the final set.  Keys are
// (fp32 distance bits << 32 | index): one unsigned compare gives "ascending distance, lower index
// wins ties".  Distances use individually rounded mul/add in a fixed association so they are
// bit-identical to the oracle's.
#include "common.cuh"
#include "scan.cuh"
#include "knn_common.cuh"

namespace {
using namespace dcb;

constexpr int KNN_THREADS = 256;
constexpr int KNN_WARPS = KNN_THREADS / 32;
constexpr int RQW = 8;                  // radius search: queries per warp
constexpr int RQB = KNN_WARPS * RQW;   // ... per block
constexpr int TILE = 1024;

// largest g in [0, B) with ptr[g] <= q
__device__ __forceinline__ int find_graph(const int64_t* __restrict__ ptr, int64_t B, int64_t q) {
  int64_t lo = 0, hi = B - 1;
  while (lo < hi) {
    int64_t mid = (lo + hi + 1) >> 1;
    if (ptr[mid] <= q) lo = mid; else hi = mid - 1;
  }
  return (int)lo;
}

// QW queries per warp; every step evaluates 64 candidates (two per lane) against all QW queries with no
// synchronisation at all (2*QW independent distance chains per lane), then ONE warp vote decides whether any pair
// beat its query's current k-th distance; only then the (rare, after the first few hundred candidates) insertion
// path runs.  Invalid queries carry NaN coordinates and a zero threshold, so they never vote.
template <int SLOTS, int QW>
__global__ void __launch_bounds__(KNN_THREADS)
knn_kernel(const float* __restrict__ pos, const int64_t* __restrict__ ptr, int64_t B, int64_t N, int kk, int loop, int W,
           int32_t* __restrict__ out, const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;   // device-side dispatch (knn_grid.cu): the grid search took this cloud
  __shared__ float sx[TILE], sy[TILE], sz[TILE];
  constexpr int QB = KNN_WARPS * QW;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t qb0 = (int64_t)blockIdx.x * QB;
  const int64_t qb1 = min(N, qb0 + QB) - 1;
  const int64_t cand_lo = ptr[find_graph(ptr, B, qb0)];
  const int64_t cand_hi = ptr[find_graph(ptr, B, qb1) + 1];

  float qx[QW], qy[QW], qz[QW];
  int64_t lo[QW], hi[QW];
  bool qv[QW];
  unsigned long long keys[QW][SLOTS], thresh[QW];
  unsigned thr_hi[QW];
#pragma unroll
  for (int qi = 0; qi < QW; ++qi) {
    const int64_t q = qb0 + warp * QW + qi;
    qv[qi] = q < N;
    const int64_t qq = qv[qi] ? q : (N - 1);
    const int g = find_graph(ptr, B, qq);
    lo[qi] = ptr[g];
    hi[qi] = ptr[g + 1];
    const float nan = __int_as_float(0x7fc00000);
    qx[qi] = qv[qi] ? pos[3 * qq] : nan; qy[qi] = qv[qi] ? pos[3 * qq + 1] : nan; qz[qi] = qv[qi] ? pos[3 * qq + 2] : nan;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) keys[qi][s] = (s * 32 + lane) < kk ? KEY_INF : 0ull;  // 0 = permanently unused slot
    thresh[qi] = qv[qi] ? KEY_INF : 0ull;
    thr_hi[qi] = qv[qi] ? 0xffffffffu : 0u;
  }

  for (int64_t t0 = cand_lo; t0 < cand_hi; t0 += TILE) {
    __syncthreads();
    const int tile_n = (int)min((int64_t)TILE, cand_hi - t0);
    for (int i = tid; i < tile_n; i += KNN_THREADS) {
      const int64_t c = t0 + i;
      sx[i] = pos[3 * c]; sy[i] = pos[3 * c + 1]; sz[i] = pos[3 * c + 2];
    }
    __syncthreads();
    // whole tile inside every valid query's own point cloud (the common case: one graph per CTA) -> no range tests
    bool all_inside = true;
#pragma unroll
    for (int qi = 0; qi < QW; ++qi) all_inside = all_inside && (!qv[qi] || (t0 >= lo[qi] && t0 + tile_n <= hi[qi]));
    for (int j0 = 0; j0 < tile_n; j0 += 64) {
      // TILE is a multiple of 64, so both reads stay inside the arrays (stale values past tile_n are masked below)
      const int ja = j0 + lane, jb = ja + 32;
      const float pxa = sx[ja], pya = sy[ja], pza = sz[ja];
      const float pxb = sx[jb], pyb = sy[jb], pzb = sz[jb];
      const int64_t ca = t0 + ja, cb = t0 + jb;
      unsigned da[QW], db[QW];
      bool pass = false;
      if (all_inside && j0 + 64 <= tile_n) {
#pragma unroll
        for (int qi = 0; qi < QW; ++qi) {
          da[qi] = __float_as_uint(sqdist(qx[qi], qy[qi], qz[qi], pxa, pya, pza));
          db[qi] = __float_as_uint(sqdist(qx[qi], qy[qi], qz[qi], pxb, pyb, pzb));
          // d >= 0 (or NaN for an invalid query): unsigned order == float order; ties on the distance go on to the exact compare
          pass = pass || da[qi] <= thr_hi[qi] || db[qi] <= thr_hi[qi];
        }
        if (!__any_sync(0xffffffffu, pass)) continue;
        // which (query, half) pairs have a survivor: one OR-reduction of a per-lane bit mask instead of 2*QW votes
        unsigned bits = 0;
#pragma unroll
        for (int qi = 0; qi < QW; ++qi)
          bits |= (da[qi] <= thr_hi[qi] ? (1u << qi) : 0u) | (db[qi] <= thr_hi[qi] ? (1u << (QW + qi)) : 0u);
        const unsigned any = __reduce_or_sync(0xffffffffu, bits);
#pragma unroll
        for (int qi = 0; qi < QW; ++qi) {
          if (any & (1u << qi))
            knn_insert<SLOTS>(keys[qi], thresh[qi], thr_hi[qi], ((unsigned long long)da[qi] << 32) | (unsigned)ca, lane);
          if (any & (1u << (QW + qi)))
            knn_insert<SLOTS>(keys[qi], thresh[qi], thr_hi[qi], ((unsigned long long)db[qi] << 32) | (unsigned)cb, lane);
        }
      } else {
        const bool inba = ja < tile_n, inbb = jb < tile_n;
#pragma unroll
        for (int qi = 0; qi < QW; ++qi) {
          const bool oka = inba && qv[qi] && ca >= lo[qi] && ca < hi[qi];
          const bool okb = inbb && qv[qi] && cb >= lo[qi] && cb < hi[qi];
          const unsigned a = __float_as_uint(sqdist(qx[qi], qy[qi], qz[qi], pxa, pya, pza));
          const unsigned b = __float_as_uint(sqdist(qx[qi], qy[qi], qz[qi], pxb, pyb, pzb));
          if (__ballot_sync(0xffffffffu, oka && a <= thr_hi[qi]))
            knn_insert<SLOTS>(keys[qi], thresh[qi], thr_hi[qi], oka ? (((unsigned long long)a << 32) | (unsigned)ca) : KEY_INF, lane);
          if (__ballot_sync(0xffffffffu, okb && b <= thr_hi[qi]))
            knn_insert<SLOTS>(keys[qi], thresh[qi], thr_hi[qi], okb ? (((unsigned long long)b << 32) | (unsigned)cb) : KEY_INF, lane);
        }
      }
    }
  }

#pragma unroll
  for (int qi = 0; qi < QW; ++qi) {
    const int64_t q = qb0 + warp * QW + qi;
    if (!qv[qi]) continue;  // warp-uniform
    knn_emit<SLOTS>(keys[qi], q, kk, loop, W, out, lane);
  }
}

__global__ void __launch_bounds__(KNN_THREADS)
radius_kernel(const float* __restrict__ pos, const int64_t* __restrict__ ptr, int64_t B, int64_t N, float r2, int cap,
              int loop, int W, int32_t* __restrict__ out, int32_t* __restrict__ count_out, const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;   // device-side dispatch (knn_grid.cu): the grid search took this cloud
  __shared__ float sx[TILE], sy[TILE], sz[TILE];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t qb0 = (int64_t)blockIdx.x * RQB;
  const int64_t qb1 = min(N, qb0 + RQB) - 1;
  const int64_t cand_lo = ptr[find_graph(ptr, B, qb0)];
  const int64_t cand_hi = ptr[find_graph(ptr, B, qb1) + 1];
  float qx[RQW], qy[RQW], qz[RQW];
  int64_t lo[RQW], hi[RQW];
  bool qv[RQW];
  int cnt[RQW];
  float r2q[RQW];   // the radius test of a query that is full is switched off (-1)
#pragma unroll
  for (int qi = 0; qi < RQW; ++qi) {
    const int64_t q = qb0 + warp * RQW + qi;
    qv[qi] = q < N;
    const int64_t qq = qv[qi] ? q : (N - 1);
    const int g = find_graph(ptr, B, qq);
    lo[qi] = ptr[g]; hi[qi] = ptr[g + 1];
    const float nan = __int_as_float(0x7fc00000);
    qx[qi] = qv[qi] ? pos[3 * qq] : nan; qy[qi] = qv[qi] ? pos[3 * qq + 1] : nan; qz[qi] = qv[qi] ? pos[3 * qq + 2] : nan;
    cnt[qi] = 0;
    r2q[qi] = r2;
  }
  for (int64_t t0 = cand_lo; t0 < cand_hi; t0 += TILE) {
    __syncthreads();
    const int tile_n = (int)min((int64_t)TILE, cand_hi - t0);
    for (int i = tid; i < tile_n; i += KNN_THREADS) {
      const int64_t c = t0 + i;
      sx[i] = pos[3 * c]; sy[i] = pos[3 * c + 1]; sz[i] = pos[3 * c + 2];
    }
    __syncthreads();
    bool all_full = true, all_inside = true;
#pragma unroll
    for (int qi = 0; qi < RQW; ++qi) {
      all_full = all_full && (!qv[qi] || cnt[qi] >= cap);
      all_inside = all_inside && (!qv[qi] || (t0 >= lo[qi] && t0 + tile_n <= hi[qi]));
    }
    if (all_full) continue;  // still takes part in the tile loads above
    for (int j0 = 0; j0 < tile_n; j0 += 64) {
      // two candidates per lane, 2*RQW vote-free distance chains, then one warp vote (hits are rare for a small radius);
      // TILE is a multiple of 64, so both reads stay inside the arrays (stale values past tile_n are masked)
      const int ja = j0 + lane, jb = ja + 32;
      const float pxa = sx[ja], pya = sy[ja], pza = sz[ja];
      const float pxb = sx[jb], pyb = sy[jb], pzb = sz[jb];
      const int64_t ca = t0 + ja, cb = t0 + jb;
      const bool fast = all_inside && j0 + 64 <= tile_n;
      const bool inba = ja < tile_n, inbb = jb < tile_n;
      bool ha[RQW], hb[RQW];
      bool pass = false;
      if (fast) {   // whole step inside every query's own point cloud: nothing but the distance tests
#pragma unroll
        for (int qi = 0; qi < RQW; ++qi) {
          ha[qi] = sqdist(qx[qi], qy[qi], qz[qi], pxa, pya, pza) < r2q[qi];   // NaN for an invalid query: never a hit
          hb[qi] = sqdist(qx[qi], qy[qi], qz[qi], pxb, pyb, pzb) < r2q[qi];
          pass = pass || ha[qi] || hb[qi];
        }
      } else {
#pragma unroll
        for (int qi = 0; qi < RQW; ++qi) {
          ha[qi] = inba && ca >= lo[qi] && ca < hi[qi] && sqdist(qx[qi], qy[qi], qz[qi], pxa, pya, pza) < r2q[qi];
          hb[qi] = inbb && cb >= lo[qi] && cb < hi[qi] && sqdist(qx[qi], qy[qi], qz[qi], pxb, pyb, pzb) < r2q[qi];
          pass = pass || ha[qi] || hb[qi];
        }
      }
      if (!__any_sync(0xffffffffu, pass)) continue;
#pragma unroll
      for (int qi = 0; qi < RQW; ++qi) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {   // ascending candidate index: first the lower 32, then the upper 32
          const bool hit = half ? hb[qi] : ha[qi];
          const unsigned m = __ballot_sync(0xffffffffu, hit);
          if (m) {
            const int p = cnt[qi] + __popc(m & ((1u << lane) - 1u));
            if (hit && p < cap) out[(qb0 + warp * RQW + qi) * W + p] = (int32_t)(half ? cb : ca);
            cnt[qi] = min(cap, cnt[qi] + __popc(m));
            if (cnt[qi] >= cap) r2q[qi] = -1.f;   // full: no further candidate can hit (hits past `cap` are dropped above)
          }
        }
      }
    }
  }
  __syncwarp();
  // drop the self match (if it made it into the first `cap` hits), pad with -1
#pragma unroll
  for (int qi = 0; qi < RQW; ++qi) {
    const int64_t q = qb0 + warp * RQW + qi;
    if (!qv[qi]) continue;
    int32_t* row = out + q * W;
    const int n = cnt[qi];
    int self_pos = n;
    if (!loop) {
      for (int p0 = 0; p0 < n; p0 += 32) {
        const int p = p0 + lane;
        const unsigned sm = __ballot_sync(0xffffffffu, p < n && (int64_t)row[p] == q);
        if (sm) { self_pos = p0 + __ffs(sm) - 1; break; }
      }
    }
    const int nout = n - (self_pos < n ? 1 : 0);
    if (self_pos < n) {
      for (int p0 = self_pos; p0 < n - 1; p0 += 32) {  // shift left by one, chunk by chunk in order
        const int p = p0 + lane;
        int32_t v = (p < n - 1) ? row[p + 1] : 0;
        __syncwarp();
        if (p < n - 1) row[p] = v;
        __syncwarp();
      }
    }
    for (int p = nout + lane; p < W; p += 32) row[p] = -1;
    if (lane == 0 && count_out) count_out[q] = nout;
  }
}

__global__ void nbr_count_kernel(const int32_t* __restrict__ nbr, int64_t N, int W, uint32_t* __restrict__ cnt) {
  int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (q >= N) return;
  int c = 0;
  for (int p = 0; p < W; ++p) c += nbr[q * W + p] >= 0;
  cnt[q] = c;
}

__global__ void nbr_emit_kernel(const int32_t* __restrict__ nbr, int64_t N, int W, const uint32_t* __restrict__ off,
                                int64_t* __restrict__ ei, int64_t cap, int64_t* __restrict__ num_out) {
  int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (q >= N) return;
  int64_t o = off[q];
  for (int p = 0; p < W; ++p) {
    int v = nbr[q * W + p];
    if (v >= 0) {
      if (o < cap) { ei[o] = v; ei[cap + o] = q; }
      ++o;
    }
  }
  if (q == N - 1) *num_out = o;
}
}  // namespace

// The two brute-force launches, shared with knn_grid.cu (its fallback for clouds a uniform grid cannot split).  `skip`
// (device int, may be NULL): the kernels return at once when *skip != 0.
namespace dcb {
int launch_knn_brute(const float* pos, const int64_t* ptr, int64_t B, int64_t N, int kk, int loop, int32_t* nbr_out,
                     const int* skip, cudaStream_t st) {
  // 8 queries per warp while the candidate set is one key per lane; fewer for wide sets (register budget)
  if (kk <= 32) knn_kernel<1, 8><<<(unsigned)cdiv(N, KNN_WARPS * 8), KNN_THREADS, 0, st>>>(pos, ptr, B, N, kk, loop, kk, nbr_out, skip);
  else if (kk <= 64) knn_kernel<2, 4><<<(unsigned)cdiv(N, KNN_WARPS * 4), KNN_THREADS, 0, st>>>(pos, ptr, B, N, kk, loop, kk, nbr_out, skip);
  else knn_kernel<4, 4><<<(unsigned)cdiv(N, KNN_WARPS * 4), KNN_THREADS, 0, st>>>(pos, ptr, B, N, kk, loop, kk, nbr_out, skip);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

int launch_radius_brute(const float* pos, const int64_t* ptr, int64_t B, int64_t N, float r2, int cap, int loop,
                        int32_t* nbr_out, int32_t* count_out, const int* skip, cudaStream_t st) {
  radius_kernel<<<(unsigned)cdiv(N, RQB), KNN_THREADS, 0, st>>>(pos, ptr, B, N, r2, cap, loop, cap, nbr_out, count_out, skip);
  DC_LAUNCH_CHECK();
  return DC_OK;
}
}  // namespace dcb

extern "C" int dc_knn(const float* pos, const int64_t* ptr, int64_t B, int64_t N, int32_t k, int loop, int32_t* nbr_out,
                      dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(N >= 0 && B >= 1 && k >= 1, DC_EINVAL, "knn: bad sizes N=%lld B=%lld k=%d", (long long)N, (long long)B, k);
  if (N == 0) return DC_OK;
  DC_REQUIRE(pos && ptr && nbr_out, DC_EINVAL, "knn: null pointer");
  const int kk = k + (loop ? 0 : 1);
  DC_REQUIRE(kk <= 128, DC_ENOSUP, "knn: k=%d exceeds the supported maximum (127, or 128 with loop)", k);
  return dcb::launch_knn_brute(pos, ptr, B, N, kk, loop, nbr_out, nullptr, st);
}

extern "C" int dc_radius(const float* pos, const int64_t* ptr, int64_t B, int64_t N, float r, int32_t max_nbr, int loop,
                         int32_t* nbr_out, int32_t* count_out, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(N >= 0 && B >= 1 && max_nbr >= 1, DC_EINVAL, "radius: bad sizes");
  if (N == 0) return DC_OK;
  DC_REQUIRE(pos && ptr && nbr_out, DC_EINVAL, "radius: null pointer");
  const int cap = max_nbr + (loop ? 0 : 1);
  const float r2 = r * r;  // fp32 product, as torch_cluster
  return dcb::launch_radius_brute(pos, ptr, B, N, r2, cap, loop, nbr_out, count_out, nullptr, st);
}

extern "C" size_t dc_nbr_to_edge_index_workspace_bytes(int64_t N) {
  dcb::Carver c(nullptr);
  c.take<uint32_t>(N > 0 ? N : 1);
  c.take<uint32_t>(scan_num_blocks(N));
  return c.used();
}

extern "C" int dc_nbr_to_edge_index(const int32_t* nbr, int64_t N, int32_t W, int64_t* edge_index, int64_t edge_cap,
                                    int64_t* num_edges_out, void* workspace, size_t workspace_bytes, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(N >= 0 && W >= 0 && num_edges_out, DC_EINVAL, "nbr_to_edge_index: bad args");
  if (N == 0 || W == 0) {
    DC_CUDA(cudaMemsetAsync(num_edges_out, 0, sizeof(int64_t), st));
    return DC_OK;
  }
  DC_REQUIRE(nbr && edge_index && workspace, DC_EINVAL, "nbr_to_edge_index: null pointer");
  DC_REQUIRE(workspace_bytes >= dc_nbr_to_edge_index_workspace_bytes(N), DC_EWORKSPACE, "nbr_to_edge_index: workspace");
  DC_REQUIRE((int64_t)N * W < (1ll << 32), DC_ENOSUP, "nbr_to_edge_index: N*W too large");
  dcb::Carver c(workspace);
  uint32_t* cnt = c.take<uint32_t>(N);
  uint32_t* bsum = c.take<uint32_t>(scan_num_blocks(N));
  nbr_count_kernel<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(nbr, N, W, cnt);
  DC_LAUNCH_CHECK();
  if (int rc = exclusive_scan_u32(cnt, N, bsum, st)) return rc;
  nbr_emit_kernel<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(nbr, N, W, cnt, edge_index, edge_cap, num_edges_out);
  DC_LAUNCH_CHECK();
  return DC_OK;
}
