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
