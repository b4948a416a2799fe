// N3 — on-GPU batch assembly: the step BEFORE the message-passing path.
// Reference: Batch.from_data_list + .to(device) (train.py:36-44; PyG data/batch.py, data/collate.py),
// mesh_to_graph (utils/graph_utils.py:12-16), _feature_rigid (loaders/common.py:6-19),
// _create_rigid_pointcloud (loaders/common.py:25-30: create_sphere + translate).
//
// The host packs the per-sample arrays back to back into ONE pinned staging buffer (indices stay
// graph-local, int32 or int64), copies it once, and these kernels produce the batched layout:
// edge_index with the cumulative node offsets added, the `batch` vector, mesh half-edges for all
// graphs of the batch in one launch, the 25-d collider features, and instanced collider spheres.
// All index arithmetic is exact; the feature kernel uses the same rounding sequence as dc_posenc.
#include "common.cuh"

namespace {
using namespace dcb;

constexpr int kPtrSmem = 2048;   // segment tables up to this many entries are staged in shared memory

// largest g in [0, B) with ptr[g] <= i   (ptr ascending, ptr[0] = 0; empty segments are skipped)
__device__ __forceinline__ int seg_of(const int64_t* ptr, int B, int64_t i) {
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (ptr[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// stage a [B+1] table in shared memory when it fits; returns the pointer to search
__device__ __forceinline__ const int64_t* stage_ptr(const int64_t* __restrict__ g, int B, int64_t* s) {
  if (B + 1 > kPtrSmem) return g;
  for (int i = threadIdx.x; i <= B; i += blockDim.x) s[i] = g[i];
  return s;
}

__global__ void __launch_bounds__(256)
batch_vector_kernel(const int64_t* __restrict__ node_ptr, int B, int64_t N, int64_t* __restrict__ batch) {
  __shared__ int64_t sp[kPtrSmem];
  const int64_t* p = stage_ptr(node_ptr, B, sp);
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
    batch[i] = seg_of(p, B, i);
}

template <typename T>
__global__ void __launch_bounds__(256)
edges_offset_kernel(const T* __restrict__ local, int64_t lstride, const int64_t* __restrict__ edge_ptr,
                    const int64_t* __restrict__ node_ptr, int B, int64_t E, int64_t* __restrict__ out, int64_t ostride) {
  __shared__ int64_t se[kPtrSmem];
  const int64_t* ep = stage_ptr(edge_ptr, B, se);
  __syncthreads();
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t off = node_ptr[seg_of(ep, B, e)];
    out[e] = (int64_t)local[e] + off;
    out[ostride + e] = (int64_t)local[lstride + e] + off;
  }
}

// one thread per half-edge; (a,b),(b,c),(c,a) per triangle in triangle order (utils/graph_utils.py:12)
template <typename T>
__global__ void __launch_bounds__(256)
mesh_edges_batched_kernel(const T* __restrict__ tri, const int64_t* __restrict__ tri_ptr,
                          const int64_t* __restrict__ node_ptr, int B, int64_t ntri, int64_t tri_per_graph,
                          int64_t nodes_per_graph, int64_t* __restrict__ out, int64_t ostride) {
  __shared__ int64_t st[kPtrSmem];
  const int64_t* tp = tri_ptr ? stage_ptr(tri_ptr, B, st) : nullptr;
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < 3 * ntri; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i / 3;
    const int c = (int)(i - 3 * t);
    int64_t off, row;
    if (tp) {                       // ragged: every graph brings its own triangles
      const int g = seg_of(tp, B, t);
      off = node_ptr[g];
      row = t;
    } else {                        // instanced: every graph reuses the same template mesh
      const int64_t g = t / tri_per_graph;
      off = g * nodes_per_graph;
      row = t - g * tri_per_graph;
    }
    out[i] = (int64_t)tri[3 * row + c] + off;
    out[ostride + i] = (int64_t)tri[3 * row + (c + 1) % 3] + off;
  }
}

// x[n] = [head[g][0..H) | pos | sin(pos), cos(pos), sin(2 pos), cos(2 pos), sin(4 pos), cos(4 pos)]
// (loaders/common.py:18 cat order; to_log_freq(pos, 3, 1) column order of utils/pos_encoding.py:32-44)
__global__ void __launch_bounds__(256)
node_features_kernel(const float* __restrict__ pos, const float* __restrict__ head, int H,
                     const int64_t* __restrict__ node_ptr, int B, int64_t N, float* __restrict__ out, int64_t ldo) {
  __shared__ int64_t sp[kPtrSmem];
  const int64_t* p = (H > 0) ? stage_ptr(node_ptr, B, sp) : nullptr;
  __syncthreads();
  const int W = H + 21;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N * W; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / W;
    const int j = (int)(i - n * W);
    float v;
    if (j < H) {
      v = head[(int64_t)seg_of(p, B, n) * H + j];
    } else {
      const int c = j - H, grp = c / 3, d = c - 3 * grp;
      const float x = pos[3 * n + d];
      if (grp == 0) v = x;
      else {
        const int b = (grp - 1) >> 1;
        const float a = __fmul_rn(x, (float)(1 << b));
        v = (grp & 1) ? sinf(a) : cosf(a);
      }
    }
    out[n * ldo + j] = v;
  }
}

// pos[g*V + v] = float(tmpl[v] + center[g])   (Open3D translate is an fp64 add; utils/graph_utils.py:10 casts to fp32)
__global__ void __launch_bounds__(256)
instance_points_kernel(const double* __restrict__ tmpl, const double* __restrict__ centers, int64_t B, int64_t V,
                       float* __restrict__ pos) {
  const int64_t total = B * V * 3;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = i / (3 * V);
    const int64_t r = i - g * 3 * V;
    pos[i] = __double2float_rn(__dadd_rn(tmpl[r], centers[3 * g + (r % 3)]));
  }
}

inline unsigned grid_for(int64_t n) { return (unsigned)std::min<int64_t>(cdiv(n, 256), (int64_t)sm_count() * 16); }
}  // namespace

extern "C" int dc_batch_vector(const int64_t* node_ptr, int64_t B, int64_t N, int64_t* batch, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(B >= 0 && N >= 0 && B < (1ll << 31), DC_EINVAL, "batch_vector: bad sizes");
  if (N == 0) return DC_OK;
  DC_REQUIRE(node_ptr && batch && B >= 1, DC_EINVAL, "batch_vector: null pointer / no graphs");
  batch_vector_kernel<<<grid_for(N), 256, 0, st>>>(node_ptr, (int)B, N, batch);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

extern "C" int dc_edges_offset(const void* local, int32_t index_bytes, int64_t local_stride, const int64_t* edge_ptr,
                               const int64_t* node_ptr, int64_t B, int64_t E, int64_t* edge_index, int64_t edge_stride,
                               dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(B >= 0 && E >= 0 && B < (1ll << 31), DC_EINVAL, "edges_offset: bad sizes");
  DC_REQUIRE(index_bytes == 4 || index_bytes == 8, DC_EINVAL, "edges_offset: index_bytes must be 4 or 8");
  if (E == 0) return DC_OK;
  DC_REQUIRE(local && edge_ptr && node_ptr && edge_index && B >= 1 && local_stride >= E && edge_stride >= E, DC_EINVAL,
             "edges_offset: bad args");
  if (index_bytes == 4)
    edges_offset_kernel<int32_t><<<grid_for(E), 256, 0, st>>>((const int32_t*)local, local_stride, edge_ptr, node_ptr, (int)B, E,
                                                              edge_index, edge_stride);
  else
    edges_offset_kernel<int64_t><<<grid_for(E), 256, 0, st>>>((const int64_t*)local, local_stride, edge_ptr, node_ptr, (int)B, E,
                                                              edge_index, edge_stride);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

extern "C" int dc_mesh_edges_batched(const void* triangles, int32_t index_bytes, const int64_t* tri_ptr,
                                     const int64_t* node_ptr, int64_t B, int64_t num_tri, int64_t tri_per_graph,
                                     int64_t nodes_per_graph, int64_t* edge_index, int64_t edge_stride, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(B >= 0 && num_tri >= 0 && B < (1ll << 31), DC_EINVAL, "mesh_edges_batched: bad sizes");
  DC_REQUIRE(index_bytes == 4 || index_bytes == 8, DC_EINVAL, "mesh_edges_batched: index_bytes must be 4 or 8");
  if (num_tri == 0) return DC_OK;
  DC_REQUIRE(triangles && edge_index && B >= 1 && edge_stride >= 3 * num_tri, DC_EINVAL, "mesh_edges_batched: bad args");
  if (tri_ptr) DC_REQUIRE(node_ptr != nullptr, DC_EINVAL, "mesh_edges_batched: ragged form needs node_ptr");
  else DC_REQUIRE(tri_per_graph >= 1 && nodes_per_graph >= 0 && num_tri == B * tri_per_graph, DC_EINVAL,
                  "mesh_edges_batched: instanced form needs num_tri == B * tri_per_graph");
  if (index_bytes == 4)
    mesh_edges_batched_kernel<int32_t><<<grid_for(3 * num_tri), 256, 0, st>>>((const int32_t*)triangles, tri_ptr, node_ptr, (int)B, num_tri,
                                                                              tri_per_graph, nodes_per_graph, edge_index, edge_stride);
  else
    mesh_edges_batched_kernel<int64_t><<<grid_for(3 * num_tri), 256, 0, st>>>((const int64_t*)triangles, tri_ptr, node_ptr, (int)B, num_tri,
                                                                              tri_per_graph, nodes_per_graph, edge_index, edge_stride);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

extern "C" int dc_node_features(const float* pos, const float* head, int32_t head_width, const int64_t* node_ptr, int64_t B,
                                int64_t N, float* out, int64_t ldo, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(N >= 0 && B >= 0 && B < (1ll << 31) && head_width >= 0 && head_width <= 64, DC_EINVAL, "node_features: bad sizes");
  if (N == 0) return DC_OK;
  DC_REQUIRE(pos && out && ldo >= head_width + 21, DC_EINVAL, "node_features: bad args");
  if (head_width > 0) DC_REQUIRE(head && node_ptr && B >= 1, DC_EINVAL, "node_features: per-graph head needs head, node_ptr");
  node_features_kernel<<<grid_for(N * (head_width + 21)), 256, 0, st>>>(pos, head, head_width, node_ptr, (int)B, N, out, ldo);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

extern "C" int dc_instance_points(const double* tmpl, const double* centers, int64_t B, int64_t V, float* pos,
                                  dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(B >= 0 && V >= 0, DC_EINVAL, "instance_points: bad sizes");
  if (B == 0 || V == 0) return DC_OK;
  DC_REQUIRE(tmpl && centers && pos, DC_EINVAL, "instance_points: null pointer");
  instance_points_kernel<<<grid_for(B * V * 3), 256, 0, st>>>(tmpl, centers, B, V, pos);
  DC_LAUNCH_CHECK();
  return DC_OK;
}
