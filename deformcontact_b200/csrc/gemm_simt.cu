// K2/K3 (fp32 SIMT variant) — the Linear GEMMs inside the convs, any shape / stride / transpose.
//   C[M,N] = act( sum_s opA(A_s)[M,K_s] opB(B_s)[K_s,N] + bias ) (+C)
// 128x128x16 tiles, 256 threads, 8x8 register micro-tiles, register-prefetch double buffering,
// optional split-K with a fixed-order second stage (deterministic; used for weight gradients
// whose reduction runs over all nodes).  The tcgen05 3xTF32 kernel (gemm_tc.cu) takes over the
// large K-major shapes; this kernel is the exact-fp32 path for everything else.
#include "common.cuh"

namespace dcb {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

struct GemmParams {
  const float* A[4];
  const float* B[4];
  int64_t lda[4], ldb[4];
  int K[4];
  int nseg;
  int64_t M, N;
  float* C;
  int64_t ldc;
  const float* bias;
  int relu, accumulate;
  int splits, kb_per_split, kb_total;
  float* partial;  // [splits, M, N] when splits > 1
  int vecA[4], vecB[4];
};

// element (r, kk) of a [128 x 16] tile.  KCONTIG: src[(row0+r)*ld + k0+kk]; else src[(k0+kk)*ld + row0+r]
template <bool KCONTIG>
__device__ __forceinline__ void load_tile(const float* __restrict__ src, int64_t ld, int64_t row0, int64_t R, int k0,
                                          int K, bool vec, float (&reg)[8]) {
  const int t = threadIdx.x;
  if (KCONTIG) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t r = row0 + (t >> 2) + 64 * i;
      const int kk = k0 + (t & 3) * 4;
      if (vec && r < R && kk + 3 < K) {
        float4 v = __ldg(reinterpret_cast<const float4*>(src + r * ld + kk));
        reg[i * 4 + 0] = v.x; reg[i * 4 + 1] = v.y; reg[i * 4 + 2] = v.z; reg[i * 4 + 3] = v.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) reg[i * 4 + j] = (r < R && kk + j < K) ? __ldg(src + r * ld + kk + j) : 0.f;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int kk = k0 + (t >> 5) + 8 * i;
      const int64_t r = row0 + (t & 31) * 4;
      if (vec && kk < K && r + 3 < R) {
        float4 v = __ldg(reinterpret_cast<const float4*>(src + (int64_t)kk * ld + r));
        reg[i * 4 + 0] = v.x; reg[i * 4 + 1] = v.y; reg[i * 4 + 2] = v.z; reg[i * 4 + 3] = v.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) reg[i * 4 + j] = (kk < K && r + j < R) ? __ldg(src + (int64_t)kk * ld + r + j) : 0.f;
      }
    }
  }
}

template <bool KCONTIG>
__device__ __forceinline__ void store_tile(float (*sm)[BM + PAD], const float (&reg)[8]) {
  const int t = threadIdx.x;
  if (KCONTIG) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sm[(t & 3) * 4 + j][(t >> 2) + 64 * i] = reg[i * 4 + j];
  } else {
#pragma unroll
    for (int i = 0; i < 2; ++i)
      *reinterpret_cast<float4*>(&sm[(t >> 5) + 8 * i][(t & 31) * 4]) =
          make_float4(reg[i * 4 + 0], reg[i * 4 + 1], reg[i * 4 + 2], reg[i * 4 + 3]);
  }
}

__device__ __forceinline__ void locate_kb(const GemmParams& p, int kb, int& seg, int& k0) {
  seg = 0;
#pragma unroll 1
  while (seg < p.nseg - 1) {
    int nk = (p.K[seg] + BK - 1) / BK;
    if (kb < nk) break;
    kb -= nk;
    ++seg;
  }
  k0 = kb * BK;
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) sgemm_kernel(const GemmParams p) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  const int kb_beg = blockIdx.z * p.kb_per_split;
  const int kb_end = min(p.kb_total, kb_beg + p.kb_per_split);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[8], rb[8];
  int seg, k0;
  if (kb_beg < kb_end) {
    locate_kb(p, kb_beg, seg, k0);
    load_tile<!TA>(p.A[seg], p.lda[seg], m0, p.M, k0, p.K[seg], p.vecA[seg], ra);
    load_tile<TB>(p.B[seg], p.ldb[seg], n0, p.N, k0, p.K[seg], p.vecB[seg], rb);
    store_tile<!TA>(As[0], ra);
    store_tile<TB>(Bs[0], rb);
  }
  __syncthreads();
  int cur = 0;
  for (int kb = kb_beg; kb < kb_end; ++kb) {
    const bool more = kb + 1 < kb_end;
    if (more) {
      locate_kb(p, kb + 1, seg, k0);
      load_tile<!TA>(p.A[seg], p.lda[seg], m0, p.M, k0, p.K[seg], p.vecA[seg], ra);
      load_tile<TB>(p.B[seg], p.ldb[seg], n0, p.N, k0, p.K[seg], p.vecB[seg], rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      store_tile<!TA>(As[cur ^ 1], ra);
      store_tile<TB>(Bs[cur ^ 1], rb);
    }
    __syncthreads();
    cur ^= 1;
  }

  const bool direct = p.splits == 1;
  float* dst = direct ? p.C : p.partial + (size_t)blockIdx.z * p.M * p.N;
  const int64_t ldd = direct ? p.ldc : p.N;
  const bool vst = ((ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (m >= p.M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int64_t n = n0 + jh * 64 + tx * 4;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] = acc[i][jh * 4 + j];
        if (direct && n + j < p.N) {
          if (p.bias) v[j] += p.bias[n + j];
          if (p.accumulate) v[j] += dst[m * ldd + n + j];
          if (p.relu) v[j] = fmaxf(v[j], 0.f);
        }
      }
      if (vst && n + 3 < p.N) {
        *reinterpret_cast<float4*>(dst + m * ldd + n) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < p.N) dst[m * ldd + n + j] = v[j];
      }
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, int64_t M, int64_t N,
                                     float* __restrict__ C, int64_t ldc, const float* __restrict__ bias, int relu,
                                     int accumulate) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  int64_t m = idx / N, n = idx % N;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += partial[(size_t)z * M * N + idx];
  if (bias) s += bias[n];
  if (accumulate) s += C[m * ldc + n];
  if (relu) s = fmaxf(s, 0.f);
  C[m * ldc + n] = s;
}

static int plan_splits(int64_t M, int64_t N, int kb_total) {
  int64_t tiles = cdiv(M, BM) * cdiv(N, BN);
  if (tiles >= sm_count() || kb_total < 64) return 1;
  int64_t want = cdiv(2 * sm_count(), tiles);
  int64_t maxs = kb_total / 16;  // at least 16 k-blocks (256 k) per split
  int64_t s = want < maxs ? want : maxs;
  return (int)(s < 1 ? 1 : s);
}

size_t gemm_simt_workspace_bytes(int64_t M, int64_t N, int64_t Ktot) {
  int kb_total = (int)cdiv(Ktot, BK);
  int s = plan_splits(M, N, kb_total);
  return s > 1 ? align_up((size_t)s * M * N * sizeof(float), 256) : 0;
}

int gemm_simt(const dc_gemm_seg* segs, int nseg, int transA, int transB, int64_t M, int64_t N, float* C, int64_t ldc,
              const float* bias, int relu, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  GemmParams p{};
  p.nseg = nseg;
  int kb_total = 0;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  for (int s = 0; s < nseg; ++s) {
    DC_REQUIRE(segs[s].A && segs[s].B && segs[s].K >= 0 && segs[s].K < (1ll << 31), DC_EINVAL, "gemm: bad segment %d", s);
    p.A[s] = segs[s].A; p.B[s] = segs[s].B; p.lda[s] = segs[s].lda; p.ldb[s] = segs[s].ldb; p.K[s] = (int)segs[s].K;
    p.vecA[s] = (segs[s].lda % 4 == 0) && al16(segs[s].A);
    p.vecB[s] = (segs[s].ldb % 4 == 0) && al16(segs[s].B);
    kb_total += (int)cdiv(segs[s].K, BK);
  }
  p.M = M; p.N = N; p.C = C; p.ldc = ldc; p.bias = bias; p.relu = relu; p.accumulate = accumulate;
  p.kb_total = kb_total;
  int splits = nseg == 1 ? plan_splits(M, N, kb_total) : 1;
  if (splits > 1) {
    size_t need = align_up((size_t)splits * M * N * sizeof(float), 256);
    DC_REQUIRE(workspace && workspace_bytes >= need, DC_EWORKSPACE, "gemm: split-K workspace %zu < %zu", workspace_bytes, need);
    p.partial = static_cast<float*>(workspace);
  }
  p.kb_per_split = (int)cdiv(kb_total > 0 ? kb_total : 1, splits);
  splits = (int)cdiv(kb_total > 0 ? kb_total : 1, p.kb_per_split);
  p.splits = splits;
  dim3 grid((unsigned)cdiv(M, BM), (unsigned)cdiv(N, BN), (unsigned)splits);
  DC_REQUIRE(grid.y <= 65535u && grid.z <= 65535u, DC_ENOSUP, "gemm: N=%lld too large for grid", (long long)N);
  if (!transA && transB) sgemm_kernel<false, true><<<grid, 256, 0, st>>>(p);
  else if (!transA && !transB) sgemm_kernel<false, false><<<grid, 256, 0, st>>>(p);
  else if (transA && !transB) sgemm_kernel<true, false><<<grid, 256, 0, st>>>(p);
  else sgemm_kernel<true, true><<<grid, 256, 0, st>>>(p);
  DC_LAUNCH_CHECK();
  if (splits > 1) {
    splitk_reduce_kernel<<<(unsigned)cdiv(M * N, 256), 256, 0, st>>>(p.partial, splits, M, N, C, ldc, bias, relu, accumulate);
    DC_LAUNCH_CHECK();
  }
  return DC_OK;
}

}  // namespace dcb
