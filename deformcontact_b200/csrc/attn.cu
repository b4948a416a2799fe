// N1 — dense row softmax and its backward for the cross attention of models/model.py:7-21
//   scores = L(x_resting) L(x_rigid)^T;  attn = softmax(scores, dim=-1);  out = attn x_rigid     (no 1/sqrt(d) scale)
// The three matrix products run on the tcgen05 3xTF32 GEMMs (gemm_tc.cu, N-chunked); these two kernels are the
// element-wise pieces in between: one CTA per row, the row cached in registers (N <= 4096) so the score matrix is
// read once and written once.  Deterministic (fixed reduction tree).
#include "common.cuh"

namespace {
using namespace dcb;

constexpr int SM_THREADS = 256;
constexpr int SM_ITEMS = 16;   // x 256 threads = rows of up to 4096 columns held in registers

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();   // sm may still be read from the previous reduction
  if (l == 0) sm[w] = v;
  __syncthreads();
  float r = sm[0];
#pragma unroll
  for (int i = 1; i < SM_THREADS / 32; ++i) r = is_max ? fmaxf(r, sm[i]) : r + sm[i];
  return r;
}

// in place: S[m, :] <- exp(S[m, :] - max) / sum
template <bool CACHED>
__global__ void __launch_bounds__(SM_THREADS)
softmax_rows_kernel(float* __restrict__ S, long long ld, int N) {
  __shared__ float sm[SM_THREADS / 32];
  float* row = S + (long long)blockIdx.x * ld;
  float v[SM_ITEMS];
  float mx = -INFINITY;
  if (CACHED) {
#pragma unroll
    for (int i = 0; i < SM_ITEMS; ++i) {
      const int c = threadIdx.x + i * SM_THREADS;
      v[i] = c < N ? row[c] : -INFINITY;
      mx = fmaxf(mx, v[i]);
    }
  } else {
    for (int c = threadIdx.x; c < N; c += SM_THREADS) mx = fmaxf(mx, row[c]);
  }
  mx = block_reduce(mx, true, sm);
  float sum = 0.f;
  if (CACHED) {
#pragma unroll
    for (int i = 0; i < SM_ITEMS; ++i) {
      v[i] = expf(v[i] - mx);   // exp(-inf) = 0 for the padding
      sum += v[i];
    }
  } else {
    for (int c = threadIdx.x; c < N; c += SM_THREADS) {
      const float e = expf(row[c] - mx);
      row[c] = e;
      sum += e;
    }
  }
  sum = block_reduce(sum, false, sm);
  const float inv = 1.0f / sum;
  if (CACHED) {
#pragma unroll
    for (int i = 0; i < SM_ITEMS; ++i) {
      const int c = threadIdx.x + i * SM_THREADS;
      if (c < N) row[c] = v[i] * inv;
    }
  } else {
    for (int c = threadIdx.x; c < N; c += SM_THREADS) row[c] *= inv;
  }
}

// in place on dP: dS[m, n] = P[m, n] * (dP[m, n] - sum_j dP[m, j] P[m, j])
template <bool CACHED>
__global__ void __launch_bounds__(SM_THREADS)
softmax_bwd_rows_kernel(const float* __restrict__ P, long long ldp, float* __restrict__ dP, long long ldd, int N) {
  __shared__ float sm[SM_THREADS / 32];
  const float* prow = P + (long long)blockIdx.x * ldp;
  float* drow = dP + (long long)blockIdx.x * ldd;
  float pv[SM_ITEMS], dv[SM_ITEMS];
  float dot = 0.f;
  if (CACHED) {
#pragma unroll
    for (int i = 0; i < SM_ITEMS; ++i) {
      const int c = threadIdx.x + i * SM_THREADS;
      pv[i] = c < N ? prow[c] : 0.f;
      dv[i] = c < N ? drow[c] : 0.f;
      dot += pv[i] * dv[i];
    }
  } else {
    for (int c = threadIdx.x; c < N; c += SM_THREADS) dot += prow[c] * drow[c];
  }
  dot = block_reduce(dot, false, sm);
  if (CACHED) {
#pragma unroll
    for (int i = 0; i < SM_ITEMS; ++i) {
      const int c = threadIdx.x + i * SM_THREADS;
      if (c < N) drow[c] = pv[i] * (dv[i] - dot);
    }
  } else {
    for (int c = threadIdx.x; c < N; c += SM_THREADS) drow[c] = prow[c] * (drow[c] - dot);
  }
}
// out[m] = sum_n A[m, n] * B[m, n]; one warp per row, lane-strided partial sums + shuffle tree (fixed order)
__global__ void __launch_bounds__(256)
rowdot_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B, long long ldb, long long M, int N,
              float* __restrict__ out) {
  const long long row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float* a = A + row * lda;
  const float* b = B + row * ldb;
  float s = 0.f;
  for (int c = lane; c < N; c += 32) s += a[c] * b[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = s;
}
}  // namespace

extern "C" int dc_rowdot(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t N, float* out, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(M >= 0 && N >= 0, DC_EINVAL, "rowdot: negative size");
  if (M == 0) return DC_OK;
  DC_REQUIRE(A && B && out && lda >= N && ldb >= N && N < (1ll << 31), DC_EINVAL, "rowdot: bad arguments");
  rowdot_kernel<<<(unsigned)cdiv(M, 8), 256, 0, st>>>(A, lda, B, ldb, M, (int)N, out);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

extern "C" int dc_softmax_rows(float* S, int64_t ld, int64_t M, int64_t N, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(M >= 0 && N >= 0, DC_EINVAL, "softmax_rows: negative size");
  if (M == 0 || N == 0) return DC_OK;
  DC_REQUIRE(S && ld >= N && M < (1ll << 31) && N < (1ll << 31), DC_EINVAL, "softmax_rows: bad arguments");
  if (N <= SM_THREADS * SM_ITEMS)
    softmax_rows_kernel<true><<<(unsigned)M, SM_THREADS, 0, st>>>(S, ld, (int)N);
  else
    softmax_rows_kernel<false><<<(unsigned)M, SM_THREADS, 0, st>>>(S, ld, (int)N);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

extern "C" int dc_softmax_bwd_rows(const float* P, int64_t ldp, float* dP, int64_t ldd, int64_t M, int64_t N, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(M >= 0 && N >= 0, DC_EINVAL, "softmax_bwd_rows: negative size");
  if (M == 0 || N == 0) return DC_OK;
  DC_REQUIRE(P && dP && ldp >= N && ldd >= N && M < (1ll << 31) && N < (1ll << 31), DC_EINVAL, "softmax_bwd_rows: bad arguments");
  if (N <= SM_THREADS * SM_ITEMS)
    softmax_bwd_rows_kernel<true><<<(unsigned)M, SM_THREADS, 0, st>>>(P, ldp, dP, ldd, (int)N);
  else
    softmax_bwd_rows_kernel<false><<<(unsigned)M, SM_THREADS, 0, st>>>(P, ldp, dP, ldd, (int)N);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// ------------------------------------------------------------------------------------------
// N2 — the training losses of train.py:47-58 in one pass over the CSR pair the encoder already built:
//   L1 displacement loss   sum_{i,c} |pred[i,c] - tgt[i,c]|                               (nn.L1Loss, train.py:21,52)
//   gradient consistency   sum_e || (tgt[i]-tgt[j]) - (pred[i]-pred[j]) ||_2, e = (j -> i)  (models/losses.py:12-17)
// and their gradients with respect to pred, which do not depend on anything upstream but two scalars:
//   gl[i,c] = sign(pred - tgt),   gc[i] = - sum_{e into i} d_e/|d_e| + sum_{e out of i} d_e/|d_e|   (0 where |d_e| = 0)
// One thread per node walks its in-edges (by-target CSR) and out-edges (by-source CSR) in CSR order: deterministic, no
// atomics.  partial[i] = {consistency terms of the edges into i, L1 terms of node i}; the caller sums the columns.
namespace {
__global__ void __launch_bounds__(256)
edge_loss_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr, const int32_t* __restrict__ rowptr_t,
                 const int32_t* __restrict__ nbr_t, const float* __restrict__ pred, const float* __restrict__ tgt, long long N,
                 float* __restrict__ partial, float* __restrict__ gc, float* __restrict__ gl) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float px = pred[3 * i], py = pred[3 * i + 1], pz = pred[3 * i + 2];
  const float tx = tgt[3 * i], ty = tgt[3 * i + 1], tz = tgt[3 * i + 2];
  const float ex = tx - px, ey = ty - py, ez = tz - pz;   // (tgt - pred)[i]; d_e = e[i] - e[j] for e = (j -> i)
  float c = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
  for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) {       // edges into i: d = e_i - e_j, dL/dpred_i = -d/|d|
    const long long j = nbr[p];
    const float dx = ex - (tgt[3 * j] - pred[3 * j]), dy = ey - (tgt[3 * j + 1] - pred[3 * j + 1]), dz = ez - (tgt[3 * j + 2] - pred[3 * j + 2]);
    const float n = sqrtf(dx * dx + dy * dy + dz * dz);
    c += n;
    if (n > 0.f) { const float r = 1.0f / n; gx -= dx * r; gy -= dy * r; gz -= dz * r; }
  }
  for (int p = rowptr_t[i]; p < rowptr_t[i + 1]; ++p) {   // edges out of i (i is the source j of e = (i -> k)): d = e_k - e_i, dL/dpred_i = +d/|d|
    const long long k = nbr_t[p];
    const float dx = (tgt[3 * k] - pred[3 * k]) - ex, dy = (tgt[3 * k + 1] - pred[3 * k + 1]) - ey, dz = (tgt[3 * k + 2] - pred[3 * k + 2]) - ez;
    const float n = sqrtf(dx * dx + dy * dy + dz * dz);
    if (n > 0.f) { const float r = 1.0f / n; gx += dx * r; gy += dy * r; gz += dz * r; }
  }
  partial[2 * i] = c;
  partial[2 * i + 1] = fabsf(ex) + fabsf(ey) + fabsf(ez);
  gc[3 * i] = gx; gc[3 * i + 1] = gy; gc[3 * i + 2] = gz;
  gl[3 * i] = (float)((px > tx) - (px < tx)); gl[3 * i + 1] = (float)((py > ty) - (py < ty)); gl[3 * i + 2] = (float)((pz > tz) - (pz < tz));
}
}  // namespace

extern "C" int dc_edge_loss(const int32_t* rowptr, const int32_t* nbr, const int32_t* rowptr_t, const int32_t* nbr_t, const float* pred,
                            const float* tgt, int64_t N, float* partial, float* gc, float* gl, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(N >= 0, DC_EINVAL, "edge_loss: negative size");
  if (N == 0) return DC_OK;
  DC_REQUIRE(rowptr && rowptr_t && pred && tgt && partial && gc && gl, DC_EINVAL, "edge_loss: null pointer");
  edge_loss_kernel<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(rowptr, nbr, rowptr_t, nbr_t, pred, tgt, N, partial, gc, gl);
  DC_LAUNCH_CHECK();
  return DC_OK;
}
