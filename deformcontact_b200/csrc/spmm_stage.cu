// K1 v10 — TMA-staged hop chain.
//
// Same contract as dc_spmm_chain (K1 v9): up to DC_MAX_CHAIN consecutive hops  out_k = add_k + A in_k  of one layer in ONE
// launch, a CTA owning one (tile of receivers = whole graphs of the block-diagonal batch) x (feature slice) across the hops;
// same per-receiver summation order and rounding (sequential in CSR order, separately rounded mul and add): bit-identical.
//
// What changes is where the gathers are served from.  v6-v9 gather source-row slices through L1: a 2000-row tile has
// 2000 distinct 128-byte lines per slice (row pitch 1 KB), 250 KB of lines for an L1 that also streams the edge records
// — measured hit rate 59 %, L2->L1 traffic 3.5x the compulsory bytes, 10 warps per scheduler waiting on the long scoreboard
// (profiles/r01_ncu_spmm_v9_chain.csv).  Here the tile's slice of the hop input is brought into SHARED MEMORY once, by TMA:
//   * one elected thread issues cp.async.bulk.tensor.2d boxes [R rows x LANES*4 floats] of the strided [rows, slice] view
//     (no swizzle; out-of-range rows / columns are zero-filled by the TMA unit), completion on an mbarrier: no registers, no
//     LSU instructions, full-line requests, every row slice crosses L2 -> SM exactly once per hop;
//   * the slice width adapts to the tile: LANES = 4..8 float4 lanes per receiver so that rows x LANES x 16 B fits the
//     227 KB of shared memory (2000 rows -> 7 lanes = 112-byte slices, <= 1760 rows -> 8 lanes = 128-byte slices); the
//     F/4 float4 columns are dealt to ceil(F4 / LANES) slices of nearly equal width;
//   * receivers then gather from shared memory (ld.shared.v4, ~30 cycles instead of an L2 round trip), 8 in flight per lane,
//     and store their row slice straight to global memory;
//   * between hops: proxy fence + block barrier, then the next hop's input (the rows this CTA just wrote, L2-resident) is
//     staged the same way.
// Any feature width with F % 4 == 0 qualifies (no F % 32 rule), so the 24 / 28-wide layer-1 hops chain too.
#include <cuda.h>

#include "common.cuh"

namespace {
using namespace dcb;

constexpr int SG_THREADS = 1024;
constexpr int SG_SMEM_MAX = 232448;        // 227 KB opt-in limit per CTA
constexpr int SG_BAR_BYTES = 128;          // mbarrier slot in front of the rows (keeps them 128-byte aligned)
constexpr int SG_ALIGN_SLACK = 128;        // the dynamic shared memory base is aligned up to 128 bytes in the kernel
constexpr int SG_MAX_CHUNK_ROWS = 256;     // TMA box dimension limit

struct SgHops {
  const float* add[DC_MAX_CHAIN];
  float* out[DC_MAX_CHAIN];
  unsigned ldadd[DC_MAX_CHAIN], ldout[DC_MAX_CHAIN];
};

__device__ __forceinline__ uint32_t sg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sg_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void sg_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sg_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void sg_tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ float4 sg_lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 sg_ld_coherent4(const float* p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void sg_mul_add(float4& acc, float w, const float4& v) {
  acc.x = __fadd_rn(acc.x, __fmul_rn(w, v.x));
  acc.y = __fadd_rn(acc.y, __fmul_rn(w, v.y));
  acc.z = __fadd_rn(acc.z, __fmul_rn(w, v.z));
  acc.w = __fadd_rn(acc.w, __fmul_rn(w, v.w));
}

// LANES float4 lanes per receiver; G = 32 / LANES receivers per warp (lanes >= G * LANES idle)
template <int LANES>
__global__ void __launch_bounds__(SG_THREADS, 1)
spmm_stage_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                  const __grid_constant__ CUtensorMap map2, const __grid_constant__ CUtensorMap map3,
                  const int32_t* __restrict__ rowptr, const int2* __restrict__ edges, const float* __restrict__ self_w,
                  const SgHops hops, int num_hops, int N, int F4, int self_loop, const int32_t* __restrict__ tile_ptr,
                  int n_slices, int tile_nodes, int chunk_rows) {
  extern __shared__ uint8_t sg_smem[];
  constexpr int G = 32 / LANES;
  constexpr int PITCH = LANES * 16;
  const uint32_t bar = (sg_smem_u32(sg_smem) + 127u) & ~127u;   // TMA destinations must be 128-byte aligned
  const uint32_t rows_base = bar + SG_BAR_BYTES;

  const int slice = blockIdx.x % n_slices;
  const int tile = blockIdx.x / n_slices;
  const int t0 = tile_ptr ? tile_ptr[tile] : tile * tile_nodes;
  const int t1 = tile_ptr ? tile_ptr[tile + 1] : min(N, t0 + tile_nodes);
  const int c0 = (int)(((long long)slice * F4) / n_slices);           // first float4 column of this slice
  const int width = (int)(((long long)(slice + 1) * F4) / n_slices) - c0;   // <= LANES
  const int n_chunks = (t1 - t0 + chunk_rows - 1) / chunk_rows;

  const int lane = threadIdx.x & 31;
  const int grp = lane / LANES, gl = lane % LANES;
  const bool active = grp < G && gl < width;
  const int groups_per_cta = (SG_THREADS / 32) * G;
  const int first = (threadIdx.x >> 5) * G + grp;

  if (threadIdx.x == 0) {
    sg_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  for (int hop = 0; hop < num_hops; ++hop) {
    if (threadIdx.x == 0 && n_chunks > 0) {
      const CUtensorMap* m = hop == 0 ? &map0 : hop == 1 ? &map1 : hop == 2 ? &map2 : &map3;
      sg_mbar_expect_tx(bar, (uint32_t)(n_chunks * chunk_rows * PITCH));
      for (int c = 0; c < n_chunks; ++c)
        sg_tma_load_2d(rows_base + (uint32_t)(c * chunk_rows * PITCH), m, bar, c0 * 4, t0 + c * chunk_rows);
    }
    const float* add = hops.add[hop];
    float* out = hops.out[hop];
    const unsigned ldadd = hops.ldadd[hop], ldo = hops.ldout[hop];
    int node = t0 + first;
    int beg = 0, end = 0;
    if (active && node < t1) { beg = __ldg(rowptr + node); end = __ldg(rowptr + node + 1); }
    if (n_chunks > 0) sg_mbar_wait(bar, (uint32_t)(hop & 1));
    if (active) {
      const uint32_t my = rows_base + (uint32_t)(gl * 16) - (uint32_t)t0 * PITCH;   // row r of the graph sits at my + r * PITCH
      const int colf = (c0 + gl) * 4;
      for (; node < t1; node += groups_per_cta) {
        int begn = 0, endn = 0;
        if (node + groups_per_cta < t1) { begn = __ldg(rowptr + node + groups_per_cta); endn = __ldg(rowptr + node + groups_per_cta + 1); }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (add != nullptr) acc = sg_ld_coherent4(add + (size_t)((unsigned)node * ldadd) + colf);
        int p = beg;
        for (; p + 8 <= end; p += 8) {
          int2 e[8];
          float4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) e[u] = __ldg(edges + p + u);
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = sg_lds4(my + (uint32_t)e[u].x * PITCH);
#pragma unroll
          for (int u = 0; u < 8; ++u) sg_mul_add(acc, __int_as_float(e[u].y), v[u]);
        }
        if (p + 4 <= end) {
          int2 e[4];
          float4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) e[u] = __ldg(edges + p + u);
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = sg_lds4(my + (uint32_t)e[u].x * PITCH);
#pragma unroll
          for (int u = 0; u < 4; ++u) sg_mul_add(acc, __int_as_float(e[u].y), v[u]);
          p += 4;
        }
        for (; p < end; ++p) {
          const int2 e = __ldg(edges + p);
          sg_mul_add(acc, __int_as_float(e.y), sg_lds4(my + (uint32_t)e.x * PITCH));
        }
        if (self_loop) sg_mul_add(acc, self_w[node], sg_lds4(my + (uint32_t)node * PITCH));
        *reinterpret_cast<float4*>(out + (size_t)((unsigned)node * ldo) + colf) = acc;
        beg = begn; end = endn;
      }
    }
    // rows stored by the generic proxy must be visible to the TMA (async proxy) reads of the next hop, and every gather of
    // this hop must be done before the staging buffer is overwritten
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
  }
}

typedef CUresult (*SgEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
SgEncodeFn sg_encode() {
  static SgEncodeFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<SgEncodeFn>(sym);
  }
  return fn;
}

// float4 lanes per receiver for tiles of up to `rows` rows staged in chunks of `chunk` rows; 0 = does not fit
int sg_lanes(int64_t rows, int chunk) {
  const int64_t padded = cdiv(rows, chunk) * chunk;
  int64_t lanes = (SG_SMEM_MAX - SG_BAR_BYTES - SG_ALIGN_SLACK) / (padded * 16);
  if (lanes > 8) lanes = 8;
  return lanes < 4 ? 0 : (int)lanes;
}
// rows per TMA box: <= 256 (box dimension limit) and a multiple of 8 so that every box lands 128-byte aligned
// (chunk * LANES * 16 bytes) whatever LANES is
int sg_chunk_rows(int64_t rows) {
  const int64_t n = cdiv(rows, SG_MAX_CHUNK_ROWS);
  return (int)(cdiv(cdiv(rows, n < 1 ? 1 : n), 8) * 8);
}
}  // namespace

// 1 if dc_spmm_stage can run tiles of up to max_tile_rows receivers (the whole tile slice must fit in shared memory)
extern "C" int dc_spmm_stage_supported(int64_t max_tile_rows, int32_t F) {
  if (max_tile_rows <= 0 || F <= 0 || F % 4) return 0;
  return sg_lanes(max_tile_rows, sg_chunk_rows(max_tile_rows)) > 0 ? 1 : 0;
}

extern "C" int dc_spmm_stage(const int32_t* rowptr, const void* edges, const float* self_w, const dc_hop_t* hops, int32_t num_hops,
                             int64_t N, int32_t F, int self_loop, const int32_t* tile_ptr, int64_t n_tiles, int32_t tile_nodes,
                             int64_t max_tile_rows, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(N >= 0 && F >= 0 && num_hops >= 0, DC_EINVAL, "spmm_stage: negative size");
  if (N == 0 || F == 0 || num_hops == 0) return DC_OK;
  DC_REQUIRE(num_hops <= DC_MAX_CHAIN, DC_EINVAL, "spmm_stage: at most %d hops", DC_MAX_CHAIN);
  DC_REQUIRE(rowptr && hops && edges, DC_EINVAL, "spmm_stage: null pointer");
  DC_REQUIRE(!self_loop || self_w, DC_EINVAL, "spmm_stage: self_loop needs self_w");
  DC_REQUIRE(F % 4 == 0 && (reinterpret_cast<uintptr_t>(edges) & 7) == 0, DC_ENOSUP, "spmm_stage: needs F %% 4 == 0");
  DC_REQUIRE(N < (1ll << 31), DC_ENOSUP, "spmm_stage: N exceeds 32-bit offsets");
  if (!tile_ptr) {
    DC_REQUIRE(tile_nodes > 0, DC_EINVAL, "spmm_stage: tile_nodes must be > 0 without tile_ptr");
    n_tiles = cdiv(N, tile_nodes);
    max_tile_rows = tile_nodes < N ? tile_nodes : N;
  }
  DC_REQUIRE(n_tiles > 0 && max_tile_rows > 0, DC_EINVAL, "spmm_stage: no tiles");
  const int chunk = sg_chunk_rows(max_tile_rows);
  const int lanes = sg_lanes(max_tile_rows, chunk);
  DC_REQUIRE(lanes > 0, DC_ENOSUP, "spmm_stage: a tile of %lld rows does not fit in shared memory (use dc_spmm_chain)",
             (long long)max_tile_rows);
  const int F4 = F / 4;
  const int n_slices = (int)cdiv(F4, lanes);
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  SgEncodeFn enc = sg_encode();
  DC_REQUIRE(enc, DC_ECUDA, "spmm_stage: cuTensorMapEncodeTiled not available");
  CUtensorMap maps[DC_MAX_CHAIN];
  SgHops a;
  for (int k = 0; k < DC_MAX_CHAIN; ++k) {
    const dc_hop_t& hp = hops[k < num_hops ? k : 0];
    if (k < num_hops) {
      DC_REQUIRE(hp.in && hp.out && hp.in != hp.out, DC_EINVAL, "spmm_stage: hop %d: null / aliased in and out", k);
      DC_REQUIRE(hp.ldin % 4 == 0 && hp.ldout % 4 == 0 && hp.ldin >= F && hp.ldout >= F && al16(hp.in) && al16(hp.out) &&
                     (!hp.add || (hp.ldadd % 4 == 0 && hp.ldadd >= F && al16(hp.add))),
                 DC_ENOSUP, "spmm_stage: hop %d: needs 16-byte aligned rows", k);
      DC_REQUIRE((uint64_t)N * (uint64_t)hp.ldout < (1ull << 32) && (!hp.add || (uint64_t)N * (uint64_t)hp.ldadd < (1ull << 32)),
                 DC_ENOSUP, "spmm_stage: hop %d: N*ld exceeds 32-bit element offsets", k);
    }
    // strided [N rows, F columns] view of the hop input; box = [lanes * 4 floats, chunk rows], dense in shared memory
    cuuint64_t gdim[2] = {(cuuint64_t)F, (cuuint64_t)N};
    cuuint64_t gstr[1] = {(cuuint64_t)hp.ldin * 4};
    cuuint32_t box[2] = {(cuuint32_t)(lanes * 4), (cuuint32_t)chunk};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&maps[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(hp.in), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DC_REQUIRE(r == CUDA_SUCCESS, DC_ECUDA, "spmm_stage: cuTensorMapEncodeTiled failed (%d)", (int)r);
    a.add[k] = hp.add; a.out[k] = hp.out;
    a.ldadd[k] = (unsigned)hp.ldadd; a.ldout[k] = (unsigned)hp.ldout;
  }
  const size_t smem = (size_t)SG_BAR_BYTES + SG_ALIGN_SLACK + (size_t)cdiv(max_tile_rows, chunk) * chunk * lanes * 16;
  const unsigned grid = (unsigned)(n_tiles * n_slices);
#define SG_LAUNCH(L)                                                                                                          \
  do {                                                                                                                        \
    static DeviceOnce once;                                                                                                   \
    if (once.first()) DC_CUDA(cudaFuncSetAttribute(spmm_stage_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, SG_SMEM_MAX)); \
    spmm_stage_kernel<L><<<grid, SG_THREADS, smem, st>>>(maps[0], maps[1], maps[2], maps[3], rowptr, static_cast<const int2*>(edges), \
                                                         self_w, a, num_hops, (int)N, F4, self_loop, tile_ptr, n_slices, tile_nodes, chunk); \
  } while (0)
  switch (lanes) {
    case 8: SG_LAUNCH(8); break;
    case 7: SG_LAUNCH(7); break;
    case 6: SG_LAUNCH(6); break;
    case 5: SG_LAUNCH(5); break;
    default: SG_LAUNCH(4); break;
  }
#undef SG_LAUNCH
  DC_LAUNCH_CHECK();
  return DC_OK;
}
