// Small kernels around the message-passing layers: bias-gradient column sums, ReLU backward,
// mesh -> edge list (utils/graph_utils.py:12-13), positional encoding (utils/pos_encoding.py:6-44)
// and the GAT attention scalars (PyG gat_conv.py / utils/softmax.py).
#include <algorithm>
#include "common.cuh"

namespace {
using namespace dcb;

constexpr int CS_ROWS = 2048;  // rows per stage-1 block

// stage 1: partial[chunk, n] = sum over the chunk's rows, fixed order (8 interleaved row lanes, then tree in smem)
__global__ void __launch_bounds__(256)
colsum_stage1(const float* __restrict__ X, int64_t ldx, int64_t M, int64_t N, float* __restrict__ partial) {
  __shared__ float sm[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t n = (int64_t)blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * CS_ROWS;
  const int64_t r1 = min(M, r0 + CS_ROWS);
  float s = 0.f;
  if (n < N)
    for (int64_t r = r0 + ty; r < r1; r += 8) s += X[r * ldx + n];
  sm[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = sm[0][tx];
#pragma unroll
    for (int i = 1; i < 8; ++i) t += sm[i][tx];
    partial[(int64_t)blockIdx.y * N + n] = t;
  }
}
// stage 1 of the fused ReLU backward + bias gradient: dX = dY * (Y > 0) written once, its column sums accumulated on the
// way in exactly the order of colsum_stage1 (so dc_relu_bwd_colsum == dc_relu_bwd followed by dc_colsum, bit for bit)
__global__ void __launch_bounds__(256)
relu_bwd_colsum_stage1(const float* __restrict__ Y, int64_t ldy, const float* __restrict__ dY, int64_t lddy, float* __restrict__ dX,
                       int64_t lddx, int64_t M, int64_t N, float* __restrict__ partial) {
  __shared__ float sm[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t n = (int64_t)blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * CS_ROWS;
  const int64_t r1 = min(M, r0 + CS_ROWS);
  float s = 0.f;
  if (n < N) {
#pragma unroll 4
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const float g = Y[r * ldy + n] > 0.f ? dY[r * lddy + n] : 0.f;
      dX[r * lddx + n] = g;
      s += g;
    }
  }
  sm[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = sm[0][tx];
#pragma unroll
    for (int i = 1; i < 8; ++i) t += sm[i][tx];
    partial[(int64_t)blockIdx.y * N + n] = t;
  }
}
// float4 forms of the two stage-1 kernels (N % 4 == 0, 16-byte aligned rows): a thread owns 4 adjacent columns, so every global
// access is 16 bytes per lane (512 contiguous bytes per warp and row) instead of 4.  Per column the rows are still summed by the
// same 8 interleaved row lanes and the same tree: bit-identical to the scalar kernels.  TXL = column lanes per block (32: 128 columns
// per block; 8: 32 columns per block, four times the blocks — for short matrices, where 128-column blocks leave most SMs idle:
// 64 blocks for the 64000 x 256 gradients of a 32-graph shard).
template <int TXL>
__global__ void __launch_bounds__(TXL * 8)
colsum_stage1_v4(const float* __restrict__ X, int64_t ldx, int64_t M, int64_t N, float* __restrict__ partial) {
  __shared__ float4 sm[8][TXL + 1];
  const int tx = threadIdx.x % TXL, ty = threadIdx.x / TXL;
  const int64_t n = ((int64_t)blockIdx.x * TXL + tx) * 4;
  const int64_t r0 = (int64_t)blockIdx.y * CS_ROWS;
  const int64_t r1 = min(M, r0 + CS_ROWS);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n < N) {
#pragma unroll 8
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const float4 x = __ldcs(reinterpret_cast<const float4*>(X + r * ldx + n));
      s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
    }
  }
  sm[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float4 t = sm[0][tx];
#pragma unroll
    for (int i = 1; i < 8; ++i) { const float4 u = sm[i][tx]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
    *reinterpret_cast<float4*>(partial + (int64_t)blockIdx.y * N + n) = t;
  }
}
template <int TXL>
__global__ void __launch_bounds__(TXL * 8)
relu_bwd_colsum_stage1_v4(const float* __restrict__ Y, int64_t ldy, const float* __restrict__ dY, int64_t lddy, float* __restrict__ dX,
                          int64_t lddx, int64_t M, int64_t N, float* __restrict__ partial) {
  __shared__ float4 sm[8][TXL + 1];
  const int tx = threadIdx.x % TXL, ty = threadIdx.x / TXL;
  const int64_t n = ((int64_t)blockIdx.x * TXL + tx) * 4;
  const int64_t r0 = (int64_t)blockIdx.y * CS_ROWS;
  const int64_t r1 = min(M, r0 + CS_ROWS);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n < N) {
#pragma unroll 4
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const float4 y = *reinterpret_cast<const float4*>(Y + r * ldy + n);
      const float4 d = *reinterpret_cast<const float4*>(dY + r * lddy + n);
      const float4 g = make_float4(y.x > 0.f ? d.x : 0.f, y.y > 0.f ? d.y : 0.f, y.z > 0.f ? d.z : 0.f, y.w > 0.f ? d.w : 0.f);
      *reinterpret_cast<float4*>(dX + r * lddx + n) = g;
      s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
    }
  }
  sm[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float4 t = sm[0][tx];
#pragma unroll
    for (int i = 1; i < 8; ++i) { const float4 u = sm[i][tx]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
    *reinterpret_cast<float4*>(partial + (int64_t)blockIdx.y * N + n) = t;
  }
}
__global__ void colsum_stage2(const float* __restrict__ partial, int64_t chunks, int64_t N, float* __restrict__ out) {
  int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int64_t c = 0; c < chunks; ++c) s += partial[c * N + n];
  out[n] = s;
}

__global__ void relu_bwd_kernel(const float* __restrict__ Y, const float* __restrict__ dY, float* __restrict__ dX, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dX[i] = Y[i] > 0.f ? dY[i] : 0.f;
}
__global__ void relu_bwd_kernel4(const float4* __restrict__ Y, const float4* __restrict__ dY, float4* __restrict__ dX, int64_t n4) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 y = Y[i], g = dY[i];
  dX[i] = make_float4(y.x > 0.f ? g.x : 0.f, y.y > 0.f ? g.y : 0.f, y.z > 0.f ? g.z : 0.f, y.w > 0.f ? g.w : 0.f);
}

__global__ void mesh_edges_kernel(const int64_t* __restrict__ tri, int64_t T, int64_t offset, int64_t* __restrict__ ei,
                                  int64_t stride, int64_t start) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // over 3T half-edges
  if (i >= 3 * T) return;
  int64_t t = i / 3, c = i % 3;
  ei[start + i] = tri[3 * t + c] + offset;
  ei[stride + start + i] = tri[3 * t + (c + 1) % 3] + offset;
}

__global__ void posenc_kernel(const float* __restrict__ pos, int64_t N, float* __restrict__ out, int64_t ldo, int col0) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // over N*3
  if (i >= 3 * N) return;
  int64_t n = i / 3;
  int d = (int)(i % 3);
  float x = pos[i];
  float* o = out + n * ldo + col0 + d;
  o[0] = x;
  float f = 1.f;
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    float a = __fmul_rn(x, f);
    o[3 * (1 + 2 * b)] = sinf(a);
    o[3 * (2 + 2 * b)] = cosf(a);
    f *= 2.f;
  }
}

// one warp per (node, head)
__global__ void __launch_bounds__(256)
gat_scores_kernel(const float* __restrict__ xs, int64_t ld, int64_t N, int H, int C, const float* __restrict__ att_src,
                  const float* __restrict__ att_dst, float* __restrict__ a_src, float* __restrict__ a_dst) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= N * H) return;
  const int64_t n = w / H;
  const int h = (int)(w % H);
  const float* row = xs + n * ld + (int64_t)h * C;
  float s = 0.f, d = 0.f;
  for (int c = lane; c < C; c += 32) {
    float v = row[c];
    s = fmaf(v, att_src[h * C + c], s);
    d = fmaf(v, att_dst[h * C + c], d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    d += __shfl_xor_sync(0xffffffffu, d, o);
  }
  if (lane == 0) { a_src[w] = s; a_dst[w] = d; }
}

__device__ __forceinline__ float leaky(float z, float slope) { return z > 0.f ? z : z * slope; }

// one thread per receiver (degrees are ~10); sums in CSR order then the self loop, like the reference scatter
__global__ void gat_softmax_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr,
                                   const int32_t* __restrict__ eid, const float* __restrict__ a_src,
                                   const float* __restrict__ a_dst, float slope, int64_t N, float* __restrict__ alpha_edge,
                                   float* __restrict__ alpha_self) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int beg = rowptr[i], end = rowptr[i + 1];
  const float ad = a_dst[i];
  const float es = leaky(a_src[i] + ad, slope);
  float mx = es;
  for (int p = beg; p < end; ++p) mx = fmaxf(mx, leaky(a_src[nbr[p]] + ad, slope));
  float sum = 0.f;
  for (int p = beg; p < end; ++p) sum += expf(leaky(a_src[nbr[p]] + ad, slope) - mx);
  const float xself = expf(es - mx);
  sum += xself;
  sum += 1e-16f;
  for (int p = beg; p < end; ++p) alpha_edge[eid[p]] = expf(leaky(a_src[nbr[p]] + ad, slope) - mx) / sum;
  alpha_self[i] = xself / sum;
}

// backward of the attention scalars, one thread per receiver:
//  dalpha_e = <dout[i], xs[nbr]>,  de = alpha * (dalpha - sum alpha*dalpha),  dz = de * leaky'(z)
// writes dz_edge[eid], dz_self[i], da_dst[i] = sum dz (edges + self)
__global__ void gat_bwd_edge_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr,
                                    const int32_t* __restrict__ eid, const float* __restrict__ a_src,
                                    const float* __restrict__ a_dst, float slope, const float* __restrict__ alpha_edge,
                                    const float* __restrict__ alpha_self, const float* __restrict__ xs, int64_t ldx,
                                    const float* __restrict__ dout, int64_t ldd, int C, int64_t N,
                                    float* __restrict__ dz_edge, float* __restrict__ dz_self, float* __restrict__ da_dst) {
  // one warp per receiver: lanes split the feature dot products
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= N) return;
  const int beg = rowptr[i], end = rowptr[i + 1];
  const float* g = dout + i * ldd;
  // pass 1: dot products -> stash dalpha in dz_edge / dz_self, accumulate sum alpha*dalpha
  float acc = 0.f;
  for (int p = beg; p <= end; ++p) {
    const bool self = p == end;
    const int64_t j = self ? i : nbr[p];
    const float* row = xs + j * ldx;
    float d = 0.f;
    for (int c = lane; c < C; c += 32) d = fmaf(g[c], row[c], d);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    const float al = self ? alpha_self[i] : alpha_edge[eid[p]];
    acc = fmaf(al, d, acc);
    if (lane == 0) {
      if (self) dz_self[i] = d; else dz_edge[eid[p]] = d;
    }
  }
  __syncwarp();
  const float ad = a_dst[i];
  float tot = 0.f;
  for (int p = beg + lane; p <= end; p += 32) {
    const bool self = p == end;
    const int e = self ? 0 : eid[p];
    const float al = self ? alpha_self[i] : alpha_edge[e];
    const float dal = self ? dz_self[i] : dz_edge[e];
    const float z = (self ? a_src[i] : a_src[nbr[p]]) + ad;
    const float dz = al * (dal - acc) * (z > 0.f ? 1.f : slope);
    if (self) dz_self[i] = dz; else dz_edge[e] = dz;
    tot += dz;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  if (lane == 0) da_dst[i] = tot;
}

// out[i] = init[i] + sum_{p in row i} val[eid[p]]   (one thread per row, CSR order)
__global__ void segment_sum_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ eid,
                                   const float* __restrict__ val, const float* __restrict__ init, int64_t N,
                                   float* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  float s = init ? init[i] : 0.f;
  for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) s += val[eid[p]];
  out[i] = s;
}
}  // namespace

extern "C" size_t dc_colsum_workspace_bytes(int64_t M, int64_t N) {
  return dcb::align_up((size_t)dcb::cdiv(M > 0 ? M : 1, CS_ROWS) * (N > 0 ? N : 1) * sizeof(float), 256);
}
extern "C" int dc_colsum(const float* X, int64_t ldx, int64_t M, int64_t N, float* out, void* workspace,
                         size_t workspace_bytes, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(M >= 0 && N >= 0, DC_EINVAL, "colsum: negative size");
  if (N == 0) return DC_OK;
  DC_REQUIRE(out, DC_EINVAL, "colsum: null out");
  if (M == 0) { DC_CUDA(cudaMemsetAsync(out, 0, N * sizeof(float), st)); return DC_OK; }
  DC_REQUIRE(X && workspace && workspace_bytes >= dc_colsum_workspace_bytes(M, N), DC_EWORKSPACE, "colsum: workspace");
  int64_t chunks = cdiv(M, CS_ROWS);
  DC_REQUIRE(chunks <= 65535, DC_ENOSUP, "colsum: M too large");
  if (N % 4 == 0 && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0) {
    if (cdiv(N, 128) * chunks >= 2 * sm_count()) {
      dim3 grid4((unsigned)cdiv(N, 128), (unsigned)chunks);
      colsum_stage1_v4<32><<<grid4, 256, 0, st>>>(X, ldx, M, N, static_cast<float*>(workspace));
    } else {
      dim3 grid4((unsigned)cdiv(N, 32), (unsigned)chunks);
      colsum_stage1_v4<8><<<grid4, 64, 0, st>>>(X, ldx, M, N, static_cast<float*>(workspace));
    }
  } else {
    dim3 grid((unsigned)cdiv(N, 32), (unsigned)chunks);
    colsum_stage1<<<grid, 256, 0, st>>>(X, ldx, M, N, static_cast<float*>(workspace));
  }
  colsum_stage2<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(static_cast<float*>(workspace), chunks, N, out);
  DC_LAUNCHED(2);
  return DC_OK;
}

extern "C" int dc_relu_bwd_colsum(const float* Y, int64_t ldy, const float* dY, int64_t lddy, float* dX, int64_t lddx, int64_t M,
                                  int64_t N, float* colsum, void* workspace, size_t workspace_bytes, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(M >= 0 && N >= 0, DC_EINVAL, "relu_bwd_colsum: negative size");
  if (N == 0) return DC_OK;
  DC_REQUIRE(colsum, DC_EINVAL, "relu_bwd_colsum: null colsum");
  if (M == 0) { DC_CUDA(cudaMemsetAsync(colsum, 0, N * sizeof(float), st)); return DC_OK; }
  DC_REQUIRE(Y && dY && dX && ldy >= N && lddy >= N && lddx >= N, DC_EINVAL, "relu_bwd_colsum: bad operands");
  DC_REQUIRE(workspace && workspace_bytes >= dc_colsum_workspace_bytes(M, N), DC_EWORKSPACE, "relu_bwd_colsum: workspace");
  int64_t chunks = cdiv(M, CS_ROWS);
  DC_REQUIRE(chunks <= 65535, DC_ENOSUP, "relu_bwd_colsum: M too large");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (N % 4 == 0 && ldy % 4 == 0 && lddy % 4 == 0 && lddx % 4 == 0 && al16(Y) && al16(dY) && al16(dX) && al16(workspace)) {
    if (cdiv(N, 128) * chunks >= 2 * sm_count()) {
      dim3 grid4((unsigned)cdiv(N, 128), (unsigned)chunks);
      relu_bwd_colsum_stage1_v4<32><<<grid4, 256, 0, st>>>(Y, ldy, dY, lddy, dX, lddx, M, N, static_cast<float*>(workspace));
    } else {
      dim3 grid4((unsigned)cdiv(N, 32), (unsigned)chunks);
      relu_bwd_colsum_stage1_v4<8><<<grid4, 64, 0, st>>>(Y, ldy, dY, lddy, dX, lddx, M, N, static_cast<float*>(workspace));
    }
  } else {
    dim3 grid((unsigned)cdiv(N, 32), (unsigned)chunks);
    relu_bwd_colsum_stage1<<<grid, 256, 0, st>>>(Y, ldy, dY, lddy, dX, lddx, M, N, static_cast<float*>(workspace));
  }
  colsum_stage2<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(static_cast<float*>(workspace), chunks, N, colsum);
  DC_LAUNCHED(2);
  return DC_OK;
}

extern "C" int dc_relu_bwd(const float* Y, const float* dY, float* dX, int64_t n, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n <= 0) return DC_OK;
  DC_REQUIRE(Y && dY && dX, DC_EINVAL, "relu_bwd: null pointer");
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (n % 4 == 0 && al(Y) && al(dY) && al(dX))
    relu_bwd_kernel4<<<(unsigned)cdiv(n / 4, 256), 256, 0, st>>>((const float4*)Y, (const float4*)dY, (float4*)dX, n / 4);
  else
    relu_bwd_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(Y, dY, dX, n);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

namespace {
// out[i, :] = in[perm[i], :]
__global__ void __launch_bounds__(256)
permute_rows_kernel4(const float4* __restrict__ in, int64_t ldin4, const int32_t* __restrict__ perm, float4* __restrict__ out,
                     int64_t ldout4, int64_t N, int nvec) {
  const int64_t total = N * nvec;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / nvec;
    const int c = (int)(i - r * nvec);
    out[r * ldout4 + c] = __ldg(in + (int64_t)perm[r] * ldin4 + c);
  }
}
__global__ void __launch_bounds__(256)
permute_rows_kernel1(const float* __restrict__ in, int64_t ldin, const int32_t* __restrict__ perm, float* __restrict__ out,
                     int64_t ldout, int64_t N, int F) {
  const int64_t total = N * F;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / F;
    const int c = (int)(i - r * F);
    out[r * ldout + c] = __ldg(in + (int64_t)perm[r] * ldin + c);
  }
}
}  // namespace

extern "C" int dc_permute_rows(const float* in, int64_t ldin, const int32_t* perm, float* out, int64_t ldout, int64_t N, int32_t F,
                               dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(N >= 0 && F >= 0, DC_EINVAL, "permute_rows: negative size");
  if (N == 0 || F == 0) return DC_OK;
  DC_REQUIRE(in && perm && out && in != out && ldin >= F && ldout >= F, DC_EINVAL, "permute_rows: bad args");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const unsigned grid = (unsigned)std::min<int64_t>(cdiv(N * (int64_t)F, 256 * 4), (int64_t)sm_count() * 32);
  if (F % 4 == 0 && ldin % 4 == 0 && ldout % 4 == 0 && al16(in) && al16(out))
    permute_rows_kernel4<<<grid ? grid : 1, 256, 0, st>>>(reinterpret_cast<const float4*>(in), ldin / 4, perm,
                                                          reinterpret_cast<float4*>(out), ldout / 4, N, F / 4);
  else
    permute_rows_kernel1<<<grid ? grid : 1, 256, 0, st>>>(in, ldin, perm, out, ldout, N, F);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

extern "C" int dc_mesh_edges(const int64_t* triangles, int64_t T, int64_t offset, int64_t* edge_index, int64_t edge_stride,
                             int64_t edge_start, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (T <= 0) return DC_OK;
  DC_REQUIRE(triangles && edge_index && edge_start >= 0 && edge_start + 3 * T <= edge_stride, DC_EINVAL, "mesh_edges: bad args");
  mesh_edges_kernel<<<(unsigned)cdiv(3 * T, 256), 256, 0, st>>>(triangles, T, offset, edge_index, edge_stride, edge_start);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

extern "C" int dc_posenc(const float* pos, int64_t N, float* out, int64_t ldo, int32_t col0, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (N <= 0) return DC_OK;
  DC_REQUIRE(pos && out && col0 >= 0 && ldo >= col0 + 21, DC_EINVAL, "posenc: bad args");
  posenc_kernel<<<(unsigned)cdiv(3 * N, 256), 256, 0, st>>>(pos, N, out, ldo, col0);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

extern "C" int dc_gat_scores(const float* xs, int64_t ld, int64_t N, int32_t H, int32_t C, const float* att_src,
                             const float* att_dst, float* a_src, float* a_dst, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (N <= 0) return DC_OK;
  DC_REQUIRE(xs && att_src && att_dst && a_src && a_dst && H >= 1 && C >= 1 && ld >= (int64_t)H * C, DC_EINVAL, "gat_scores: bad args");
  gat_scores_kernel<<<(unsigned)cdiv(N * H, 8), 256, 0, st>>>(xs, ld, N, H, C, att_src, att_dst, a_src, a_dst);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

extern "C" int dc_gat_softmax(const int32_t* rowptr, const int32_t* nbr, const int32_t* eid, const float* a_src,
                              const float* a_dst, float slope, int64_t N, float* alpha_edge, float* alpha_self,
                              dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (N <= 0) return DC_OK;
  DC_REQUIRE(rowptr && a_src && a_dst && alpha_edge && alpha_self, DC_EINVAL, "gat_softmax: null pointer");
  gat_softmax_kernel<<<(unsigned)cdiv(N, 128), 128, 0, st>>>(rowptr, nbr, eid, a_src, a_dst, slope, N, alpha_edge, alpha_self);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

extern "C" int dc_gat_bwd_edge(const int32_t* rowptr, const int32_t* nbr, const int32_t* eid, const float* a_src,
                               const float* a_dst, float slope, const float* alpha_edge, const float* alpha_self,
                               const float* xs, int64_t ldx, const float* dout, int64_t ldd, int32_t C, int64_t N,
                               float* dz_edge, float* dz_self, float* da_dst, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (N <= 0) return DC_OK;
  DC_REQUIRE(rowptr && a_src && a_dst && alpha_edge && alpha_self && xs && dout && dz_edge && dz_self && da_dst,
             DC_EINVAL, "gat_bwd_edge: null pointer");
  gat_bwd_edge_kernel<<<(unsigned)cdiv(N, 8), 256, 0, st>>>(rowptr, nbr, eid, a_src, a_dst, slope, alpha_edge, alpha_self,
                                                            xs, ldx, dout, ldd, C, N, dz_edge, dz_self, da_dst);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

extern "C" int dc_segment_sum(const int32_t* rowptr, const int32_t* eid, const float* val, const float* init, int64_t N,
                              float* out, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (N <= 0) return DC_OK;
  DC_REQUIRE(rowptr && val && out, DC_EINVAL, "segment_sum: null pointer");
  segment_sum_kernel<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(rowptr, eid, val, init, N, out);
  DC_LAUNCH_CHECK();
  return DC_OK;
}
