// K4g — exact kNN / radius search on a uniform grid for ONE large point cloud (torch_cluster.knn_graph /
// radius_graph semantics as called at utils/pointcloud_utils.py:10,12 with batch=None).  Same results, bit for bit,
// as the brute-force kernels of knn.cu (same distance arithmetic, same (distance, index) keys, same top-k set), but
// every query only visits the cells around it:
//   1. bounding box -> cell size h for ~8 points per cell (all parameters computed ON THE DEVICE: no host sync);
//   2. counting sort of the points by cell (integer atomics; the order inside a cell does not influence the result);
//   3. one warp per query walks Chebyshev rings of cells around its own; the x-run of cells of a ring row is ONE
//      contiguous range of the sorted array; candidates are tested 32 at a time and inserted into the
//      lane-distributed top-k set; the walk stops once the k-th distance is provably smaller than the distance to
//      anything outside the visited box (kNN) / the box provably contains the ball of radius r (radius search).
// The stop test is conservative (margins far above the fp32 rounding of the cell assignment and of the distances),
// so unvisited points can never belong to the answer; ties are resolved by the exact key compare as in knn.cu.
// Radius search keeps the `cap` SMALLEST INDICES among the hits (the brute-force kernel's "first by index" rule) with
// the same top-k machinery on keys = index.
// Batched form (dc_knn_grid_batched, kNN only): ONE GRID PER GRAPH of a batch of large clouds (Batch.ptr) in the same launches —
// per-graph bounding box and cell size, the graphs' cells laid end to end in one cell array (graph b owns cells
// [cell_base_b, cell_base_b + cells_b), cell_base_b = 2 (ptr[b] / 8) + 66 b in closed form), one counting sort for the whole batch (the
// sorted range of graph b is [ptr[b], ptr[b+1]) again), every query walks the rings of its own graph's grid only.
#include <cstddef>
#include "common.cuh"
#include "scan.cuh"
#include "knn_common.cuh"

namespace {
using namespace dcb;

struct GridParams {
  float lo[3];
  float inv_h, h, margin;   // margin: absolute slack of the stop test (>> rounding of the cell assignment)
  int dim[3];
  int cells;
  int use_grid;            // device-side dispatch: 1 = the grid search runs, 0 = the brute-force kernels take the cloud (batched: entry 0 decides)
  int cell_base;           // first cell of this graph in the batch's cell array (0 for a single cloud)
  int64_t ptr[2];          // {0, N}: the one-cloud `ptr` the brute-force kernels expect
  unsigned long long cost; // sum over cells of count^2 (x27 = candidate visits of the first ring)
};
static_assert(offsetof(GridParams, use_grid) == 40 && offsetof(GridParams, ptr) == 48 && sizeof(GridParams) == 72,
              "GridParams layout is read by ops.grid_took_it");

constexpr int BB_THREADS = 256;
constexpr int BB_BLOCKS = 296;   // 2 x 148 partial bounding boxes

__global__ void __launch_bounds__(BB_THREADS)
bbox_partial_kernel(const float* __restrict__ pos, int64_t N, float* __restrict__ partial /*[BB_BLOCKS][6]*/) {
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float v = pos[3 * i + d];
      mn[d] = fminf(mn[d], v);
      mx[d] = fmaxf(mx[d], v);
    }
  }
  __shared__ float sm[BB_THREADS / 32][6];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { sm[threadIdx.x >> 5][d] = mn[d]; sm[threadIdx.x >> 5][3 + d] = mx[d]; }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    float v = sm[0][threadIdx.x];
    for (int w = 1; w < BB_THREADS / 32; ++w) v = threadIdx.x < 3 ? fminf(v, sm[w][threadIdx.x]) : fmaxf(v, sm[w][threadIdx.x]);
    partial[blockIdx.x * 6 + threadIdx.x] = v;
  }
}

// cell size for `target_cells` cells over the longest extent of the box [mn, mx], grid dimensions (<= max_cells cells)
__device__ GridParams make_grid_params(const float (&mn)[3], const float (&mx)[3], int64_t target_cells, int64_t max_cells, int64_t N) {
  float ext[3], emax = 0.f;
  for (int d = 0; d < 3; ++d) { ext[d] = mx[d] - mn[d]; emax = fmaxf(emax, ext[d]); }
  GridParams g;
  if (!(emax > 0.f) || !isfinite(emax)) {   // all points coincide (or non-finite input, or no points): one cell, brute force inside it
    for (int d = 0; d < 3; ++d) { g.lo[d] = isfinite(mn[d]) ? mn[d] : 0.f; g.dim[d] = 1; }
    g.h = 1.f; g.inv_h = 0.f; g.margin = 0.f; g.cells = 1;
  } else {
    float h = emax / cbrtf((float)target_cells);
    for (;;) {   // dims = floor(ext/h)+1 per axis; grow h until the grid fits the workspace
      int64_t c = 1;
      for (int d = 0; d < 3; ++d) { g.dim[d] = (int)fminf(floorf(ext[d] / h), 2.0e6f) + 1; c *= g.dim[d]; }
      if (c <= max_cells) { g.cells = (int)c; break; }
      h *= 1.1f;
    }
    for (int d = 0; d < 3; ++d) g.lo[d] = mn[d];
    g.h = h; g.inv_h = 1.f / h;
    float amax = 0.f;
    for (int d = 0; d < 3; ++d) amax = fmaxf(amax, fmaxf(fabsf(mn[d]), fabsf(mx[d])));
    g.margin = 1.0e-5f * (emax + amax);   // ~100 ulp of the largest coordinate / extent
  }
  g.use_grid = 1; g.cell_base = 0; g.ptr[0] = 0; g.ptr[1] = N; g.cost = 0ull;
  return g;
}

// one thread: final bounding box of the cloud and its grid
__global__ void grid_params_kernel(const float* __restrict__ partial, int nblocks, int64_t target_cells, int64_t max_cells,
                                   int64_t N, GridParams* __restrict__ gp) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int b = 0; b < nblocks; ++b)
    for (int d = 0; d < 3; ++d) { mn[d] = fminf(mn[d], partial[b * 6 + d]); mx[d] = fmaxf(mx[d], partial[b * 6 + 3 + d]); }
  *gp = make_grid_params(mn, mx, target_cells, max_cells, N);
}

// ---- batched: one CTA per graph -> its bounding box and grid; cells of graph b start at 2 (ptr[b] / 8) + 66 b
__host__ __device__ inline int64_t grid_target_cells(int64_t N) { return N / 8 > 1 ? N / 8 : 1; }
__host__ __device__ inline int64_t grid_max_cells(int64_t N) { return 2 * grid_target_cells(N) + 64; }
__host__ __device__ inline int64_t batched_cell_base(int64_t first_point, int64_t b) { return 2 * (first_point / 8) + 66 * b; }

__global__ void __launch_bounds__(BB_THREADS)
grid_params_batched_kernel(const float* __restrict__ pos, const int64_t* __restrict__ gptr, GridParams* __restrict__ gp) {
  const int b = blockIdx.x;
  const int64_t p0 = gptr[b], p1 = gptr[b + 1];
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int64_t i = p0 + threadIdx.x; i < p1; i += blockDim.x) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float v = pos[3 * i + d];
      mn[d] = fminf(mn[d], v);
      mx[d] = fmaxf(mx[d], v);
    }
  }
  __shared__ float sm[BB_THREADS / 32][6];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { sm[threadIdx.x >> 5][d] = mn[d]; sm[threadIdx.x >> 5][3 + d] = mx[d]; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < BB_THREADS / 32; ++w)
      for (int d = 0; d < 3; ++d) { sm[0][d] = fminf(sm[0][d], sm[w][d]); sm[0][3 + d] = fmaxf(sm[0][3 + d], sm[w][3 + d]); }
    const float lo[3] = {sm[0][0], sm[0][1], sm[0][2]}, hi[3] = {sm[0][3], sm[0][4], sm[0][5]};
    const int64_t n = p1 - p0;
    GridParams g = make_grid_params(lo, hi, grid_target_cells(n), grid_max_cells(n), n);
    g.cell_base = (int)batched_cell_base(p0, b);
    g.ptr[0] = p0; g.ptr[1] = p1;
    gp[b] = g;
  }
}

// graph of point / sorted position i: the last b with gptr[b] <= i (empty graphs are skipped by construction)
__device__ __forceinline__ int find_graph(const int64_t* __restrict__ gptr, int B, int64_t i) {
  int lo = 0, hi = B;   // invariant: gptr[lo] <= i < gptr[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(gptr + mid) <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// A uniform grid only pays while no cell holds a large share of the cloud (one far outlier, two distant clusters, ... put
// almost everything into a few cells and every query would scan them 32 candidates at a time).  cost = sum count^2 estimates
// the candidate visits (x27 for the first ring); the tiled brute-force kernel needs N^2 pair tests at a ~5x lower price
// each.  The decision is taken on the device (no host sync): both searches are launched, one returns immediately.
__global__ void __launch_bounds__(256)
grid_cost_kernel(const uint32_t* __restrict__ start, GridParams* __restrict__ gp, int64_t cells_batched) {
  const int cells = cells_batched > 0 ? (int)cells_batched : gp->cells;
  unsigned long long c = 0ull;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += gridDim.x * blockDim.x) {
    const unsigned long long n = start[i + 1] - start[i];
    c += n * n;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&gp->cost, c);
}

__global__ void grid_decide_kernel(GridParams* __restrict__ gp, int64_t N) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double visits = 27.0 * 5.0 * (double)gp->cost;
  gp->use_grid = visits <= (double)N * (double)N ? 1 : 0;
}
// batched: the brute-force price is the sum of n_b^2; one decision for the whole batch, in entry 0.  A pair test of the brute-force
// kernel is NOT five times cheaper here: scans of a few thousand candidates spend most of their 64-candidate steps in the top-k
// insertion path (0.16-0.49 T pair distances/s at 1000-5000 points per graph against 1.78 T for one 200k cloud,
// profiles/r02c_knn_batch_lab.txt), so candidate visits are weighed like pair tests.
__global__ void grid_decide_batched_kernel(GridParams* __restrict__ gp, const int64_t* __restrict__ gptr, int B) {
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  double pairs = 0.0;
  for (int b = threadIdx.x; b < B; b += 32) { const double n = (double)(gptr[b + 1] - gptr[b]); pairs += n * n; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pairs += __shfl_xor_sync(0xffffffffu, pairs, o);
  if (threadIdx.x == 0) gp->use_grid = 27.0 * (double)gp->cost <= pairs ? 1 : 0;
}

__device__ __forceinline__ int cell_coord(float p, float lo, float inv_h, int dim) {
  const int c = (int)floorf((p - lo) * inv_h);
  return min(max(c, 0), dim - 1);
}

__global__ void __launch_bounds__(256)
cell_count_kernel(const float* __restrict__ pos, int64_t N, const GridParams* __restrict__ gp, uint32_t* __restrict__ cell_of,
                  uint32_t* __restrict__ count, const int64_t* __restrict__ gptr, int B) {
  GridParams g = *gp;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    if (gptr) g = gp[find_graph(gptr, B, i)];
    const int cx = cell_coord(pos[3 * i], g.lo[0], g.inv_h, g.dim[0]);
    const int cy = cell_coord(pos[3 * i + 1], g.lo[1], g.inv_h, g.dim[1]);
    const int cz = cell_coord(pos[3 * i + 2], g.lo[2], g.inv_h, g.dim[2]);
    const uint32_t c = (uint32_t)g.cell_base + ((uint32_t)cz * g.dim[1] + cy) * g.dim[0] + cx;
    cell_of[i] = c;
    atomicAdd(count + c, 1u);
  }
}

__global__ void __launch_bounds__(256)
cell_scatter_kernel(const float* __restrict__ pos, int64_t N, const uint32_t* __restrict__ cell_of,
                    const uint32_t* __restrict__ start, uint32_t* __restrict__ cursor, float4* __restrict__ sorted) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t c = cell_of[i];
    const uint32_t p = start[c] + atomicAdd(cursor + c, 1u);
    sorted[p] = make_float4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], __int_as_float((int)i));
  }
}

constexpr int GRID_THREADS = 256;

// MODE 0: kNN (keys = distance bits << 32 | index);  MODE 1: radius (keys = index of the hits, d2 < r2)
template <int SLOTS, int MODE>
__global__ void __launch_bounds__(GRID_THREADS)
grid_search_kernel(const float4* __restrict__ sorted, const uint32_t* __restrict__ start, const GridParams* __restrict__ gp,
                   int64_t N, int kk, int loop, int W, float r2, int32_t* __restrict__ out, int32_t* __restrict__ count_out,
                   const int64_t* __restrict__ gptr, int B) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (GRID_THREADS / 32) + (threadIdx.x >> 5);
  if (w >= N) return;   // warp-uniform
  if (!gp->use_grid) return;   // the brute-force kernels take this cloud / batch (see grid_decide_kernel)
  const GridParams g = gptr ? gp[find_graph(gptr, B, w)] : *gp;   // batched: sorted position w lies in its graph's point range
  const float4 qp = sorted[w];
  const int64_t q = (int64_t)__float_as_int(qp.w);
  const int cx = cell_coord(qp.x, g.lo[0], g.inv_h, g.dim[0]);
  const int cy = cell_coord(qp.y, g.lo[1], g.inv_h, g.dim[1]);
  const int cz = cell_coord(qp.z, g.lo[2], g.inv_h, g.dim[2]);
  unsigned long long keys[SLOTS], thresh = KEY_INF;
  unsigned thr_hi = 0xffffffffu;
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) keys[s] = (s * 32 + lane) < kk ? KEY_INF : 0ull;
  const int rmax = max(g.dim[0], max(g.dim[1], g.dim[2]));

  auto scan_range = [&](uint32_t beg, uint32_t end) {
    for (uint32_t j0 = beg; j0 < end; j0 += 32) {   // warp-uniform bounds
      const uint32_t j = j0 + lane;
      unsigned long long key = KEY_INF;
      if (j < end) {
        const float4 c = __ldg(sorted + j);
        const float d = sqdist(qp.x, qp.y, qp.z, c.x, c.y, c.z);
        if (MODE == 0) key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_int(c.w);
        else if (d < r2) key = (unsigned long long)(unsigned)__float_as_int(c.w);
      }
      if (__ballot_sync(0xffffffffu, (unsigned)(key >> 32) <= thr_hi && key != KEY_INF))
        knn_insert<SLOTS>(keys, thresh, thr_hi, key, lane);
    }
  };

  for (int r = 0;; ++r) {
    const int z0 = max(cz - r, 0), z1 = min(cz + r, g.dim[2] - 1);
    const int y0 = max(cy - r, 0), y1 = min(cy + r, g.dim[1] - 1);
    const int x0 = max(cx - r, 0), x1 = min(cx + r, g.dim[0] - 1);
    for (int z = z0; z <= z1; ++z) {
      for (int y = y0; y <= y1; ++y) {
        const uint32_t row = (uint32_t)g.cell_base + ((uint32_t)z * g.dim[1] + y) * g.dim[0];
        if (abs(z - cz) == r || abs(y - cy) == r) {
          scan_range(start[row + x0], start[row + x1 + 1]);       // the whole x-run of the ring row: one contiguous range
        } else {                                                  // interior row: only its two end cells are new
          if (cx - r >= 0) scan_range(start[row + cx - r], start[row + cx - r + 1]);
          if (cx + r < g.dim[0]) scan_range(start[row + cx + r], start[row + cx + r + 1]);
        }
      }
    }
    if (r >= rmax) break;   // the whole grid has been visited
    // distance from the query to the nearest face of the visited box that still has cells behind it
    float m = INFINITY;
    if (cx - r > 0) m = fminf(m, qp.x - (g.lo[0] + (float)(cx - r) * g.h));
    if (cx + r < g.dim[0] - 1) m = fminf(m, (g.lo[0] + (float)(cx + r + 1) * g.h) - qp.x);
    if (cy - r > 0) m = fminf(m, qp.y - (g.lo[1] + (float)(cy - r) * g.h));
    if (cy + r < g.dim[1] - 1) m = fminf(m, (g.lo[1] + (float)(cy + r + 1) * g.h) - qp.y);
    if (cz - r > 0) m = fminf(m, qp.z - (g.lo[2] + (float)(cz - r) * g.h));
    if (cz + r < g.dim[2] - 1) m = fminf(m, (g.lo[2] + (float)(cz + r + 1) * g.h) - qp.z);
    if (m == INFINITY) break;   // nothing outside the box
    const float ms = m - g.margin;
    if (ms > 0.f) {
      const float bound = ms * ms * 0.9999f;
      if (MODE == 0) {
        if (thr_hi != 0xffffffffu && __uint_as_float(thr_hi) < bound) break;   // set full and k-th distance inside the box
      } else {
        if (r2 <= bound) break;   // the box contains the whole ball
      }
    }
  }

  if (MODE == 0) {
    knn_emit<SLOTS>(keys, q, kk, loop, W, out, lane);
  } else {
    // keys are plain indices: ascending index order = the brute-force kernel's emission order; drop self, pad, count
    warp_bitonic_sort<SLOTS>(keys, lane);
    constexpr int CAP = 32 * SLOTS;
    const int ndummy = CAP - kk;
    int self_rank = CAP, nvalid = 0;
    bool valid[SLOTS];
    int rank[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      rank[s] = s * 32 + lane - ndummy;
      valid[s] = rank[s] >= 0 && keys[s] != KEY_INF;
      const unsigned sm = __ballot_sync(0xffffffffu, valid[s] && !loop && (int64_t)(unsigned)keys[s] == q);
      if (sm) self_rank = s * 32 + (__ffs(sm) - 1) - ndummy;
      nvalid += __popc(__ballot_sync(0xffffffffu, valid[s]));
    }
    const int nout = nvalid - (self_rank < CAP ? 1 : 0);
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      if (valid[s] && rank[s] != self_rank) {
        const int p = rank[s] - (rank[s] > self_rank ? 1 : 0);
        if (p < W) out[q * W + p] = (int32_t)(unsigned)keys[s];
      }
    }
    for (int p = nout + lane; p < W; p += 32) out[q * W + p] = -1;
    if (lane == 0 && count_out) count_out[q] = nout;
  }
}

// cells of the batched cell array: the closed-form bases leave room for every graph's own maximum (see batched_cell_base)
inline int64_t batched_total_cells(int64_t N, int64_t B) { return 2 * (N / 8) + 66 * B + 2; }

struct GridWs {
  GridParams* gp;
  float* partial;
  uint32_t *cell_of, *count, *cursor, *bsum;
  float4* sorted;
  size_t bytes;
};

GridWs carve_grid(void* ws, int64_t N, int64_t B = 1) {
  Carver c(ws);
  GridWs g;
  const int64_t mc = B > 1 ? batched_total_cells(N, B) : grid_max_cells(N);
  g.gp = c.take<GridParams>(B > 1 ? B : 1);
  g.partial = c.take<float>(BB_BLOCKS * 6);
  g.cell_of = c.take<uint32_t>(N);
  g.count = c.take<uint32_t>(mc + 1);    // counts, then (in place) the exclusive scan = cell starts; [cells] = N
  g.cursor = c.take<uint32_t>(mc + 1);
  g.bsum = c.take<uint32_t>(scan_num_blocks(mc + 1));
  g.sorted = c.take<float4>(N);
  g.bytes = c.used();
  return g;
}

// steps 1 + 2: bounding box, grid parameters, counting sort by cell
int build_grid(const float* pos, int64_t N, const GridWs& g, cudaStream_t st) {
  const int64_t mc = grid_max_cells(N);
  DC_CUDA(cudaMemsetAsync(g.count, 0, (mc + 1) * sizeof(uint32_t), st));
  DC_CUDA(cudaMemsetAsync(g.cursor, 0, (mc + 1) * sizeof(uint32_t), st));
  bbox_partial_kernel<<<BB_BLOCKS, BB_THREADS, 0, st>>>(pos, N, g.partial);
  grid_params_kernel<<<1, 32, 0, st>>>(g.partial, BB_BLOCKS, grid_target_cells(N), mc, N, g.gp);
  const unsigned nb = (unsigned)std::min<int64_t>(cdiv(N, 256), (int64_t)sm_count() * 16);
  cell_count_kernel<<<nb, 256, 0, st>>>(pos, N, g.gp, g.cell_of, g.count, nullptr, 1);
  DC_LAUNCHED(3);
  if (int rc = exclusive_scan_u32(g.count, mc + 1, g.bsum, st)) return rc;   // unused cells keep start = N
  cell_scatter_kernel<<<nb, 256, 0, st>>>(pos, N, g.cell_of, g.count, g.cursor, g.sorted);
  grid_cost_kernel<<<(unsigned)std::min<int64_t>(cdiv(mc, 256), 64), 256, 0, st>>>(g.count, g.gp, 0);
  grid_decide_kernel<<<1, 32, 0, st>>>(g.gp, N);
  DC_LAUNCHED(3);
  return DC_OK;
}

// the same for a batch: per-graph boxes and grids (one CTA per graph), ONE counting sort over the batch's cell array
int build_grid_batched(const float* pos, const int64_t* gptr, int64_t B, int64_t N, const GridWs& g, cudaStream_t st) {
  const int64_t mc = batched_total_cells(N, B);
  DC_CUDA(cudaMemsetAsync(g.count, 0, (mc + 1) * sizeof(uint32_t), st));
  DC_CUDA(cudaMemsetAsync(g.cursor, 0, (mc + 1) * sizeof(uint32_t), st));
  grid_params_batched_kernel<<<(unsigned)B, BB_THREADS, 0, st>>>(pos, gptr, g.gp);
  const unsigned nb = (unsigned)std::min<int64_t>(cdiv(N, 256), (int64_t)sm_count() * 16);
  cell_count_kernel<<<nb, 256, 0, st>>>(pos, N, g.gp, g.cell_of, g.count, gptr, (int)B);
  DC_LAUNCHED(2);
  if (int rc = exclusive_scan_u32(g.count, mc + 1, g.bsum, st)) return rc;   // empty cells (and the gaps between graphs) keep the next start
  cell_scatter_kernel<<<nb, 256, 0, st>>>(pos, N, g.cell_of, g.count, g.cursor, g.sorted);
  grid_cost_kernel<<<(unsigned)std::min<int64_t>(cdiv(mc, 256), 64), 256, 0, st>>>(g.count, g.gp, mc);
  grid_decide_batched_kernel<<<1, 32, 0, st>>>(g.gp, gptr, (int)B);
  DC_LAUNCHED(3);
  return DC_OK;
}
}  // namespace

namespace {
__global__ void __launch_bounds__(256)
extract_order_kernel(const float4* __restrict__ sorted, int64_t N, int32_t* __restrict__ order) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < N) order[i] = __float_as_int(sorted[i].w);
}
}  // namespace

extern "C" size_t dc_knn_grid_workspace_bytes(int64_t N) { return carve_grid(nullptr, N > 0 ? N : 1).bytes; }

extern "C" int dc_knn_grid(const float* pos, int64_t N, int32_t k, int loop, int32_t* nbr_out, int32_t* order_out, void* workspace,
                           size_t workspace_bytes, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(N >= 0 && k >= 1, DC_EINVAL, "knn_grid: bad sizes N=%lld k=%d", (long long)N, k);
  if (N == 0) return DC_OK;
  DC_REQUIRE(pos && nbr_out && workspace, DC_EINVAL, "knn_grid: null pointer");
  DC_REQUIRE(N < (1ll << 31), DC_ENOSUP, "knn_grid: N exceeds 32-bit indices");
  const int kk = k + (loop ? 0 : 1);
  DC_REQUIRE(kk <= 128, DC_ENOSUP, "knn_grid: k=%d exceeds the supported maximum (127, or 128 with loop)", k);
  const GridWs g = carve_grid(workspace, N);
  DC_REQUIRE(workspace_bytes >= g.bytes, DC_EWORKSPACE, "knn_grid: workspace too small");
  if (int rc = build_grid(pos, N, g, st)) return rc;
  const unsigned grid = (unsigned)cdiv(N, GRID_THREADS / 32);
  if (kk <= 32) grid_search_kernel<1, 0><<<grid, GRID_THREADS, 0, st>>>(g.sorted, g.count, g.gp, N, kk, loop, kk, 0.f, nbr_out, nullptr, nullptr, 1);
  else if (kk <= 64) grid_search_kernel<2, 0><<<grid, GRID_THREADS, 0, st>>>(g.sorted, g.count, g.gp, N, kk, loop, kk, 0.f, nbr_out, nullptr, nullptr, 1);
  else grid_search_kernel<4, 0><<<grid, GRID_THREADS, 0, st>>>(g.sorted, g.count, g.gp, N, kk, loop, kk, 0.f, nbr_out, nullptr, nullptr, 1);
  DC_LAUNCH_CHECK();
  if (int rc = launch_knn_brute(pos, g.gp->ptr, 1, N, kk, loop, nbr_out, &g.gp->use_grid, st)) return rc;   // runs iff use_grid == 0
  if (order_out) {
    extract_order_kernel<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(g.sorted, N, order_out);
    DC_LAUNCH_CHECK();
  }
  return DC_OK;
}

extern "C" size_t dc_knn_grid_batched_workspace_bytes(int64_t N, int64_t B) {
  return carve_grid(nullptr, N > 0 ? N : 1, B > 1 ? B : 2).bytes;
}

extern "C" int dc_knn_grid_batched(const float* pos, const int64_t* ptr, int64_t B, int64_t N, int32_t k, int loop, int32_t* nbr_out,
                                   void* workspace, size_t workspace_bytes, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(N >= 0 && B >= 1 && k >= 1, DC_EINVAL, "knn_grid_batched: bad sizes N=%lld B=%lld k=%d", (long long)N, (long long)B, k);
  if (N == 0) return DC_OK;
  DC_REQUIRE(pos && ptr && nbr_out && workspace, DC_EINVAL, "knn_grid_batched: null pointer");
  DC_REQUIRE(N < (1ll << 31) && B < (1ll << 24) && batched_total_cells(N, B) < (1ll << 31), DC_ENOSUP, "knn_grid_batched: batch too large");
  const int kk = k + (loop ? 0 : 1);
  DC_REQUIRE(kk <= 128, DC_ENOSUP, "knn_grid_batched: k=%d exceeds the supported maximum (127, or 128 with loop)", k);
  const GridWs g = carve_grid(workspace, N, B > 1 ? B : 2);
  DC_REQUIRE(workspace_bytes >= g.bytes, DC_EWORKSPACE, "knn_grid_batched: workspace too small");
  if (int rc = build_grid_batched(pos, ptr, B, N, g, st)) return rc;
  const unsigned grid = (unsigned)cdiv(N, GRID_THREADS / 32);
  if (kk <= 32) grid_search_kernel<1, 0><<<grid, GRID_THREADS, 0, st>>>(g.sorted, g.count, g.gp, N, kk, loop, kk, 0.f, nbr_out, nullptr, ptr, (int)B);
  else if (kk <= 64) grid_search_kernel<2, 0><<<grid, GRID_THREADS, 0, st>>>(g.sorted, g.count, g.gp, N, kk, loop, kk, 0.f, nbr_out, nullptr, ptr, (int)B);
  else grid_search_kernel<4, 0><<<grid, GRID_THREADS, 0, st>>>(g.sorted, g.count, g.gp, N, kk, loop, kk, 0.f, nbr_out, nullptr, ptr, (int)B);
  DC_LAUNCH_CHECK();
  return launch_knn_brute(pos, ptr, B, N, kk, loop, nbr_out, &g.gp->use_grid, st);   // runs iff the batch was handed back (use_grid == 0)
}

extern "C" int dc_radius_grid_batched(const float* pos, const int64_t* ptr, int64_t B, int64_t N, float r, int32_t max_nbr, int loop,
                                      int32_t* nbr_out, int32_t* count_out, void* workspace, size_t workspace_bytes,
                                      dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(N >= 0 && B >= 1 && max_nbr >= 1, DC_EINVAL, "radius_grid_batched: bad sizes");
  if (N == 0) return DC_OK;
  DC_REQUIRE(pos && ptr && nbr_out && workspace, DC_EINVAL, "radius_grid_batched: null pointer");
  DC_REQUIRE(N < (1ll << 31) && B < (1ll << 24) && batched_total_cells(N, B) < (1ll << 31), DC_ENOSUP, "radius_grid_batched: batch too large");
  const int cap = max_nbr + (loop ? 0 : 1);
  DC_REQUIRE(cap <= 128, DC_ENOSUP, "radius_grid_batched: max_num_neighbors=%d exceeds the supported maximum (127)", max_nbr);
  const GridWs g = carve_grid(workspace, N, B > 1 ? B : 2);
  DC_REQUIRE(workspace_bytes >= g.bytes, DC_EWORKSPACE, "radius_grid_batched: workspace too small");
  if (int rc = build_grid_batched(pos, ptr, B, N, g, st)) return rc;
  const float r2 = r * r;  // fp32 product, as torch_cluster / dc_radius
  const unsigned grid = (unsigned)cdiv(N, GRID_THREADS / 32);
  if (cap <= 32) grid_search_kernel<1, 1><<<grid, GRID_THREADS, 0, st>>>(g.sorted, g.count, g.gp, N, cap, loop, cap, r2, nbr_out, count_out, ptr, (int)B);
  else if (cap <= 64) grid_search_kernel<2, 1><<<grid, GRID_THREADS, 0, st>>>(g.sorted, g.count, g.gp, N, cap, loop, cap, r2, nbr_out, count_out, ptr, (int)B);
  else grid_search_kernel<4, 1><<<grid, GRID_THREADS, 0, st>>>(g.sorted, g.count, g.gp, N, cap, loop, cap, r2, nbr_out, count_out, ptr, (int)B);
  DC_LAUNCH_CHECK();
  return launch_radius_brute(pos, ptr, B, N, r2, cap, loop, nbr_out, count_out, &g.gp->use_grid, st);
}

extern "C" int dc_radius_grid(const float* pos, int64_t N, float r, int32_t max_nbr, int loop, int32_t* nbr_out,
                              int32_t* count_out, int32_t* order_out, void* workspace, size_t workspace_bytes, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(N >= 0 && max_nbr >= 1, DC_EINVAL, "radius_grid: bad sizes");
  if (N == 0) return DC_OK;
  DC_REQUIRE(pos && nbr_out && workspace, DC_EINVAL, "radius_grid: null pointer");
  DC_REQUIRE(N < (1ll << 31), DC_ENOSUP, "radius_grid: N exceeds 32-bit indices");
  const int cap = max_nbr + (loop ? 0 : 1);
  DC_REQUIRE(cap <= 128, DC_ENOSUP, "radius_grid: max_num_neighbors=%d exceeds the supported maximum (127)", max_nbr);
  const GridWs g = carve_grid(workspace, N);
  DC_REQUIRE(workspace_bytes >= g.bytes, DC_EWORKSPACE, "radius_grid: workspace too small");
  if (int rc = build_grid(pos, N, g, st)) return rc;
  const float r2 = r * r;  // fp32 product, as torch_cluster / dc_radius
  const unsigned grid = (unsigned)cdiv(N, GRID_THREADS / 32);
  if (cap <= 32) grid_search_kernel<1, 1><<<grid, GRID_THREADS, 0, st>>>(g.sorted, g.count, g.gp, N, cap, loop, cap, r2, nbr_out, count_out, nullptr, 1);
  else if (cap <= 64) grid_search_kernel<2, 1><<<grid, GRID_THREADS, 0, st>>>(g.sorted, g.count, g.gp, N, cap, loop, cap, r2, nbr_out, count_out, nullptr, 1);
  else grid_search_kernel<4, 1><<<grid, GRID_THREADS, 0, st>>>(g.sorted, g.count, g.gp, N, cap, loop, cap, r2, nbr_out, count_out, nullptr, 1);
  DC_LAUNCH_CHECK();
  if (int rc = launch_radius_brute(pos, g.gp->ptr, 1, N, r2, cap, loop, nbr_out, count_out, &g.gp->use_grid, st)) return rc;
  if (order_out) {
    extract_order_kernel<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(g.sorted, N, order_out);
    DC_LAUNCH_CHECK();
  }
  return DC_OK;
}

extern "C" int dc_cell_order(const float* pos, int64_t N, int32_t* order, void* workspace, size_t workspace_bytes,
                             dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(N >= 0, DC_EINVAL, "cell_order: negative size");
  if (N == 0) return DC_OK;
  DC_REQUIRE(pos && order && workspace, DC_EINVAL, "cell_order: null pointer");
  DC_REQUIRE(N < (1ll << 31), DC_ENOSUP, "cell_order: N exceeds 32-bit indices");
  const GridWs g = carve_grid(workspace, N);
  DC_REQUIRE(workspace_bytes >= g.bytes, DC_EWORKSPACE, "cell_order: workspace too small");
  if (int rc = build_grid(pos, N, g, st)) return rc;
  extract_order_kernel<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(g.sorted, N, order);
  DC_LAUNCH_CHECK();
  return DC_OK;
}
