// K1 — CSR-by-receiver gather / segmented sum (forward) and, on the by-source CSR, its
// transpose (backward wrt features).  Replaces MessagePassing.propagate(aggr='add') with
// message = w_e * x_j  (PyG message_passing.py; call sites models/model.py:71,77): no [E,F]
// temporaries, no scatter atomics; each receiver's sum runs sequentially in CSR order with
// separately rounded fp32 mul and add, i.e. the order and rounding of the CPU reference.
//
// Layout: h [N, ldh] fp32 row-major.  A group of LPN lanes owns one receiver; each lane keeps
// VPL float4 accumulators (LPN * VPL * 4 >= F).  The group first loads up to LPN neighbour ids
// and weights with one coalesced load, then walks them with shuffles, keeping UNROLL row
// gathers (128-bit each) in flight per lane.
#include "common.cuh"

namespace {
using namespace dcb;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <int LPN, int VPL, int UNROLL>
__global__ void __launch_bounds__(256)
spmm_vec_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr, const float* __restrict__ dis,
                const float* __restrict__ edge_w, const int32_t* __restrict__ ewi, const float* __restrict__ self_w, const float* __restrict__ h,
                int64_t ldh, float* __restrict__ out, int64_t ldo, const float* __restrict__ add, int64_t ldadd,
                int64_t N, int nvec /* F/4 */, int self_loop, const float* __restrict__ bias, int relu) {
  constexpr int GROUPS = 256 / LPN;
  const int tid = threadIdx.x;
  const int gl = tid % LPN;            // lane within group
  const int64_t node = (int64_t)blockIdx.x * GROUPS + tid / LPN;
  const unsigned gmask = (LPN == 32) ? 0xffffffffu : (((1u << LPN) - 1u) << ((tid & 31) / LPN * LPN));
  if (node >= N) return;  // whole group exits together
  const int beg = rowptr[node], end = rowptr[node + 1];
  const float di = dis ? dis[node] : 1.0f;

  float4 acc[VPL];
  bool act[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    act[v] = (gl + v * LPN) < nvec;
    acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (add != nullptr && act[v]) acc[v] = ldg4(add + node * ldadd + 4 * (gl + v * LPN));
  }

  for (int base = beg; base < end; base += LPN) {
    const int me = base + gl;
    int nb = 0;
    float w = 0.f;
    if (me < end) {
      nb = nbr[me];
      w = dis ? __fmul_rn(dis[nb], di) : 1.0f;
      if (edge_w) { const float ew = edge_w[ewi ? ewi[me] : me]; w = dis ? __fmul_rn(w, ew) : ew; }
    }
    const int cnt = min(LPN, end - base);
    for (int j = 0; j < cnt; j += UNROLL) {
      float4 val[UNROLL][VPL];
      float wj[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int src = __shfl_sync(gmask, nb, j + u, LPN);
        wj[u] = __shfl_sync(gmask, w, j + u, LPN);
        const bool ok = (j + u) < cnt;
        const float* row = h + (int64_t)src * ldh;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          val[u][v] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok && act[v]) val[u][v] = ldg4(row + 4 * (gl + v * LPN));
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if ((j + u) < cnt) {
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            acc[v].x = __fadd_rn(acc[v].x, __fmul_rn(wj[u], val[u][v].x));
            acc[v].y = __fadd_rn(acc[v].y, __fmul_rn(wj[u], val[u][v].y));
            acc[v].z = __fadd_rn(acc[v].z, __fmul_rn(wj[u], val[u][v].z));
            acc[v].w = __fadd_rn(acc[v].w, __fmul_rn(wj[u], val[u][v].w));
          }
        }
      }
    }
  }
  if (self_loop) {
    const float ws = self_w ? self_w[node] : __fmul_rn(di, di);
    const float* row = h + node * ldh;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      if (act[v]) {
        float4 x = ldg4(row + 4 * (gl + v * LPN));
        acc[v].x = __fadd_rn(acc[v].x, __fmul_rn(ws, x.x));
        acc[v].y = __fadd_rn(acc[v].y, __fmul_rn(ws, x.y));
        acc[v].z = __fadd_rn(acc[v].z, __fmul_rn(ws, x.z));
        acc[v].w = __fadd_rn(acc[v].w, __fmul_rn(ws, x.w));
      }
    }
  }
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    if (!act[v]) continue;
    if (bias) {
      const float4 b = ldg4(bias + 4 * (gl + v * LPN));
      acc[v].x = __fadd_rn(acc[v].x, b.x); acc[v].y = __fadd_rn(acc[v].y, b.y);
      acc[v].z = __fadd_rn(acc[v].z, b.z); acc[v].w = __fadd_rn(acc[v].w, b.w);
    }
    if (relu) {
      acc[v].x = fmaxf(acc[v].x, 0.f); acc[v].y = fmaxf(acc[v].y, 0.f);
      acc[v].z = fmaxf(acc[v].z, 0.f); acc[v].w = fmaxf(acc[v].w, 0.f);
    }
    *reinterpret_cast<float4*>(out + node * ldo + 4 * (gl + v * LPN)) = acc[v];
  }
}

// Generic path: any F, any alignment; one lane per feature column (strided by LPN).
template <int LPN>
__global__ void __launch_bounds__(256)
spmm_scalar_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr, const float* __restrict__ dis,
                   const float* __restrict__ edge_w, const int32_t* __restrict__ ewi, const float* __restrict__ self_w, const float* __restrict__ h,
                   int64_t ldh, float* __restrict__ out, int64_t ldo, const float* __restrict__ add, int64_t ldadd,
                   int64_t N, int F, int self_loop, const float* __restrict__ bias, int relu) {
  constexpr int GROUPS = 256 / LPN;
  const int tid = threadIdx.x;
  const int gl = tid % LPN;
  const int64_t node = (int64_t)blockIdx.x * GROUPS + tid / LPN;
  if (node >= N) return;
  const int beg = rowptr[node], end = rowptr[node + 1];
  const float di = dis ? dis[node] : 1.0f;
  for (int f = gl; f < F; f += LPN) {
    float acc = add ? add[node * ldadd + f] : 0.f;
    for (int e = beg; e < end; ++e) {
      const int nb = nbr[e];
      float w = dis ? __fmul_rn(dis[nb], di) : 1.0f;
      if (edge_w) { const float ew = edge_w[ewi ? ewi[e] : e]; w = dis ? __fmul_rn(w, ew) : ew; }
      acc = __fadd_rn(acc, __fmul_rn(w, __ldg(h + (int64_t)nb * ldh + f)));
    }
    if (self_loop) {
      const float ws = self_w ? self_w[node] : __fmul_rn(di, di);
      acc = __fadd_rn(acc, __fmul_rn(ws, __ldg(h + node * ldh + f)));
    }
    if (bias) acc = __fadd_rn(acc, bias[f]);
    if (relu) acc = fmaxf(acc, 0.f);
    out[node * ldo + f] = acc;
  }
}

template <int LPN, int VPL, int UNROLL>
void launch_vec(const int32_t* rowptr, const int32_t* nbr, const float* dis, const float* edge_w, const int32_t* ewi, const float* self_w,
                const float* h, int64_t ldh, float* out, int64_t ldo, const float* add, int64_t ldadd, int64_t N, int F,
                int self_loop, const float* bias, int relu, cudaStream_t st) {
  constexpr int GROUPS = 256 / LPN;
  spmm_vec_kernel<LPN, VPL, UNROLL><<<(unsigned)cdiv(N, GROUPS), 256, 0, st>>>(rowptr, nbr, dis, edge_w, ewi, self_w, h, ldh,
                                                                                out, ldo, add, ldadd, N, F / 4, self_loop, bias, relu);
}
}  // namespace

extern "C" int dc_spmm(const int32_t* rowptr, const int32_t* nbr, const float* dis, const float* edge_w,
                       const int32_t* ewi, const float* self_w, const float* h, int64_t ldh, float* out, int64_t ldo, const float* add,
                       int64_t ldadd, int64_t N, int32_t F, int self_loop, const float* bias, int relu, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(N >= 0 && F >= 0, DC_EINVAL, "spmm: negative size");
  if (N == 0 || F == 0) return DC_OK;
  DC_REQUIRE(rowptr && h && out, DC_EINVAL, "spmm: null pointer");  // nbr may be NULL for an edgeless graph
  DC_REQUIRE(ldh >= F && ldo >= F && (!add || ldadd >= F), DC_EINVAL, "spmm: leading dimension < F");
  DC_REQUIRE(h != out, DC_EINVAL, "spmm: out must not alias h");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  bool vec = (F % 4 == 0) && (ldh % 4 == 0) && (ldo % 4 == 0) && al16(h) && al16(out) &&
             (!add || ((ldadd % 4 == 0) && al16(add))) && (!bias || al16(bias)) && F <= 1024;
  if (vec) {
    const int nv = F / 4;
#define DC_SPMM_ARGS rowptr, nbr, dis, edge_w, ewi, self_w, h, ldh, out, ldo, add, ldadd, N, F, self_loop, bias, relu, st
    if (nv <= 8) launch_vec<8, 1, 4>(DC_SPMM_ARGS);
    else if (nv <= 16) launch_vec<16, 1, 4>(DC_SPMM_ARGS);
    else if (nv <= 32) launch_vec<32, 1, 4>(DC_SPMM_ARGS);
    else if (nv <= 64) launch_vec<32, 2, 4>(DC_SPMM_ARGS);
    else if (nv <= 128) launch_vec<32, 4, 2>(DC_SPMM_ARGS);
    else launch_vec<32, 8, 1>(DC_SPMM_ARGS);
#undef DC_SPMM_ARGS
  } else {
    if (F <= 8)
      spmm_scalar_kernel<8><<<(unsigned)cdiv(N, 32), 256, 0, st>>>(rowptr, nbr, dis, edge_w, ewi, self_w, h, ldh, out, ldo,
                                                                   add, ldadd, N, F, self_loop, bias, relu);
    else if (F <= 16)
      spmm_scalar_kernel<16><<<(unsigned)cdiv(N, 16), 256, 0, st>>>(rowptr, nbr, dis, edge_w, ewi, self_w, h, ldh, out, ldo,
                                                                    add, ldadd, N, F, self_loop, bias, relu);
    else
      spmm_scalar_kernel<32><<<(unsigned)cdiv(N, 8), 256, 0, st>>>(rowptr, nbr, dis, edge_w, ewi, self_w, h, ldh, out, ldo,
                                                                   add, ldadd, N, F, self_loop, bias, relu);
  }
  DC_LAUNCH_CHECK();
  return DC_OK;
}
