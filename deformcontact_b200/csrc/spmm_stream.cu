// K1 v11 — "stream" hop chain.  Same contract, same (tile x 128-byte feature slice) CTA mapping, same summation order and
// rounding as the K1 v9 chain kernel (spmm.cu: spmm_chain_kernel) — bit-identical — with the per-warp latency chain of
// v9 removed.  What the ncu source view of v9 showed (profiles/r02_ncu_spmm_chain_c5_source.csv): a warp iteration is
//   rowptr -> 8 edge records -> 8 row gathers -> 64 mul/add -> store,
// strictly in that order, so every iteration drains the memory pipeline twice: 20 % of all stall samples sit on the first
// use of the edge records, 39 % on the first use of each gathered row, and nothing is in flight while the warp computes and
// stores.  Here
//   * an 8-lane group owns a CONTIGUOUS range of receivers of the tile (balanced by edges + rows with a binary search in
//     rowptr), so its edges are ONE contiguous stream of packed records;
//   * the records of the stream are copied into a small per-group shared-memory ring by cp.async three chunks (24 edges)
//     ahead: reading a record is a 29-cycle broadcast LDS, never a global round trip, and costs no registers;
//   * the row gathers form a ROLLING window: slot j of 8 is refilled with the gather of edge s + 8 the moment edge s has
//     been accumulated, across receiver boundaries, so 8 x 16 B per lane are in flight at all times;
//   * receiver boundaries (store, next addend) are handled inside the stream; the next row pointer and the next addend are
//     loaded one receiver ahead.
#include <stdlib.h>

#include "common.cuh"

namespace dcb {
namespace {

constexpr int SS_CHUNK = 8;                       // records per cp.async chunk = one per lane of the group
constexpr int SS_SLOTS = 4;                       // ring depth in chunks: consuming c, gathering from c + 1, c + 2 / c + 3 in flight
constexpr int SS_RING = SS_CHUNK * SS_SLOTS;      // 32 records
constexpr int SS_RING_BYTES = SS_RING * 8 + 16;   // + 16: consecutive groups start 4 banks apart (conflict-free 16-byte broadcasts)
constexpr int SS_ROW_WEIGHT = 2;                  // balancing key: edges + 2 * rows (a receiver costs about two edges: store + addend)

struct StreamHops {
  const float* in[DC_MAX_CHAIN];
  const float* add[DC_MAX_CHAIN];
  float* out[DC_MAX_CHAIN];
  unsigned ldin[DC_MAX_CHAIN], ldadd[DC_MAX_CHAIN], ldout[DC_MAX_CHAIN];
};

__device__ __forceinline__ float4 ss_ld4(const char* p) {   // coherent: rows written earlier in this launch are read back
  float4 v;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void ss_mul_add(float4& acc, float w, const float4& v) {   // two roundings per term, like the oracle
  acc.x = __fadd_rn(acc.x, __fmul_rn(w, v.x));
  acc.y = __fadd_rn(acc.y, __fmul_rn(w, v.y));
  acc.z = __fadd_rn(acc.z, __fmul_rn(w, v.z));
  acc.w = __fadd_rn(acc.w, __fmul_rn(w, v.w));
}
__device__ __forceinline__ void ss_cp8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ss_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void ss_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ int4 ss_lds_rec2(uint32_t a) {   // two consecutive records (16-byte aligned)
  int4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
  return r;
}
__device__ __forceinline__ int2 ss_lds_rec(uint32_t a) {
  int2 r;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a));
  return r;
}

// BATCH = gathers per wait point.  The hardware tracks outstanding loads with six counting scoreboards per warp: a wait
// on one of them returns when EVERY load assigned to it has landed, so a window refilled one load at a time (first version of
// this kernel: 0.66 ms per hop against 0.35 for v9) makes each use wait for the load issued just before it.  The window is
// therefore two batches: batch A is consumed (one wait) and refilled while batch B is in flight, and vice versa.
template <int THREADS, int BATCH>
__global__ void __launch_bounds__(THREADS, 1)
spmm_stream_kernel(const int32_t* __restrict__ rowptr, const int2* __restrict__ edges, const StreamHops hops, int num_hops, int N,
                   const int32_t* __restrict__ tile_ptr, int n_slices, int tile_nodes) {
  constexpr int GROUPS = THREADS / 8;
  constexpr int WIN = 2 * BATCH;                 // gather window in edges
  constexpr int AHEAD = WIN / SS_CHUNK;          // the records gathered from in round c belong to chunk c + AHEAD
  static_assert(BATCH == 4 || BATCH == 8, "batch of 4 or 8 gathers");
  static_assert(AHEAD + 3 <= SS_SLOTS + 1 && SS_RING % WIN == 0, "ring too small");
  __shared__ __align__(16) unsigned char ring_mem[GROUPS * SS_RING_BYTES];
  __shared__ int bounds[GROUPS + 1];

  const int slice = blockIdx.x % n_slices;
  const int tile = blockIdx.x / n_slices;
  const int t0 = tile_ptr ? tile_ptr[tile] : tile * tile_nodes;
  const int t1 = tile_ptr ? tile_ptr[tile + 1] : min(N, t0 + tile_nodes);
  const int gl = threadIdx.x & 7;
  const int grp = threadIdx.x >> 3;
  const int col = slice * 32 + gl * 4;

  // ---- receivers of the tile -> GROUPS contiguous ranges of equal cost (edges + SS_ROW_WEIGHT per receiver)
  if (threadIdx.x <= GROUPS) {
    const int e0 = __ldg(rowptr + t0);
    const long long total = (long long)(__ldg(rowptr + t1) - e0) + (long long)SS_ROW_WEIGHT * (t1 - t0);
    const long long target = total * threadIdx.x / GROUPS;
    int lo = t0, hi = t1;   // first r with key(r) >= target; key(t0) = 0, key(t1) = total
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const long long key = (long long)(__ldg(rowptr + mid) - e0) + (long long)SS_ROW_WEIGHT * (mid - t0);
      if (key < target) lo = mid + 1; else hi = mid;
    }
    bounds[threadIdx.x] = lo;
  }
  __syncthreads();
  const int r0 = bounds[grp], r1 = bounds[grp + 1];
  const int S0 = r0 < r1 ? __ldg(rowptr + r0) : 0;
  const int len = r0 < r1 ? __ldg(rowptr + r1) - S0 : 0;               // edges of this group's stream
  const int len_max = __reduce_max_sync(0xffffffffu, len);             // warp-uniform round count
  const uint32_t rb = (uint32_t)__cvta_generic_to_shared(ring_mem) + (uint32_t)(grp * SS_RING_BYTES);
  const int2* __restrict__ recs = edges + S0;

  for (int hop = 0; hop < num_hops; ++hop) {
    // byte pointers and 32-bit byte pitches: a row address is ONE IMAD.WIDE.U32
    const char* hcol = reinterpret_cast<const char*>(hops.in[hop] + col);
    const unsigned ldh = hops.ldin[hop] * 4u;
    const char* add = reinterpret_cast<const char*>(hops.add[hop]);
    const unsigned ldadd = hops.ldadd[hop] * 4u, ldo = hops.ldout[hop] * 4u;
    char* out = reinterpret_cast<char*>(hops.out[hop] + col);
    if (add) add += col * 4;

    // ---- records: chunks 0 .. AHEAD + 1 on their way before anything else
#pragma unroll
    for (int c = 0; c < AHEAD + 2; ++c) {
      const int q = c * SS_CHUNK + gl;
      if (q < len) ss_cp8(rb + (uint32_t)((q & (SS_RING - 1)) * 8), recs + q);
      ss_commit();
    }
    // Receiver state: RAW absolute row ends (no arithmetic on a freshly loaded row pointer: the value is first needed one
    // receiver later, when it has long landed) and running pointers (one add per receiver instead of an index multiply).
    int rows_left = r1 - r0;                                           // receivers not yet stored, the current one included
    int rend = 0x7fffffff, rend1 = 0x7fffffff;                         // end of the current / next receiver's edges (CSR positions)
    const int32_t* rp_next = rowptr + r0 + 2;                          // row end to load at the next receiver change
    char* out_row = out + (size_t)(unsigned)r0 * ldo;
    const char* add_row = add ? add + (size_t)(unsigned)r0 * ldadd : nullptr;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), addn = acc;
    if (rows_left > 0) {
      rend = __ldg(rowptr + r0 + 1);
      if (rows_left > 1) rend1 = __ldg(rp_next);
      ++rp_next;
      if (add) {
        acc = ss_ld4(add_row);
        if (rows_left > 1) addn = ss_ld4(add_row + ldadd);
        add_row += 2 * (size_t)ldadd;
      }
    }
    ss_wait<2>();        // chunks 0 .. AHEAD - 1 have landed
    __syncwarp();
    float4 v[WIN];
    float w[WIN];
#pragma unroll
    for (int j = 0; j < WIN; ++j) {
      v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      w[j] = 0.f;
      if (j < len) {
        const int2 rec = ss_lds_rec(rb + (uint32_t)(j * 8));
        w[j] = __int_as_float(rec.y);
        v[j] = ss_ld4(hcol + (size_t)(unsigned)rec.x * ldh);
      }
    }

    // receiver r is complete when the stream position reaches its end (the loop also walks over receivers without edges)
    auto next_receiver = [&](int qa) {   // qa: CSR position of the edge about to be accumulated
      while (qa >= rend) {
        *reinterpret_cast<float4*>(out_row) = acc;
        out_row += ldo;
        acc = addn;
        rend = rend1;
        --rows_left;
        if (rows_left > 1) {
          rend1 = __ldg(rp_next);
          if (add) addn = ss_ld4(add_row);
        }
        ++rp_next;
        add_row += ldadd;
      }
    };
    for (int base = 0; base < len_max; base += WIN) {
      const uint32_t ring_next = rb + (uint32_t)(((base + WIN) & (SS_RING - 1)) * 8);   // record of edge base + WIN (WIN | SS_RING)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if ((h * BATCH) % SS_CHUNK == 0) {   // a new chunk of 8 edges starts here (compile-time condition)
          const int cc = (base + h * BATCH) / SS_CHUNK;
          ss_wait<1>();      // chunk cc + AHEAD has landed (only the chunk after it may still be in flight)
          __syncwarp();      // ... for every lane of the group; and every lane is done reading the slot refilled below
          const int q = (cc + AHEAD + 2) * SS_CHUNK + gl;
          if (q < len) ss_cp8(rb + (uint32_t)((q & (SS_RING - 1)) * 8), recs + q);
          ss_commit();
        }
        if (base + WIN + (h + 1) * BATCH <= len) {
          // ---- steady state: every edge of this batch and of its refill exists -> no per-edge predicates.
          // Consume batch h (ONE wait for its BATCH gathers, issued a batch ago) ...
#pragma unroll
          for (int jj = 0; jj < BATCH; ++jj) {
            const int j = h * BATCH + jj;
            if (S0 + base + j >= rend) next_receiver(S0 + base + j);
            ss_mul_add(acc, w[j], v[j]);
          }
          // ... and refill it with the gathers of the edges WIN further down the stream (records: chunk cc + AHEAD, landed)
          int4 rec[BATCH / 2];   // two records per 16-byte shared-memory load: fewer loads in flight, fewer scoreboards taken
#pragma unroll
          for (int jj = 0; jj < BATCH / 2; ++jj) rec[jj] = ss_lds_rec2(ring_next + (uint32_t)((h * BATCH + 2 * jj) * 8));
#pragma unroll
          for (int jj = 0; jj < BATCH / 2; ++jj) {
            const int j = h * BATCH + 2 * jj;
            w[j] = __int_as_float(rec[jj].y);
            v[j] = ss_ld4(hcol + (size_t)(unsigned)rec[jj].x * ldh);
            w[j + 1] = __int_as_float(rec[jj].w);
            v[j + 1] = ss_ld4(hcol + (size_t)(unsigned)rec[jj].z * ldh);
          }
        } else {
          // ---- head / tail of the stream: the same with every edge predicated
#pragma unroll
          for (int jj = 0; jj < BATCH; ++jj) {
            const int j = h * BATCH + jj;
            const int q = base + j;
            if (q < len) {
              next_receiver(S0 + q);
              ss_mul_add(acc, w[j], v[j]);
            }
          }
#pragma unroll
          for (int jj = 0; jj < BATCH; ++jj) {
            const int j = h * BATCH + jj;
            if (base + j + WIN < len) {
              const int2 rec = ss_lds_rec(ring_next + (uint32_t)(j * 8));
              w[j] = __int_as_float(rec.y);
              v[j] = ss_ld4(hcol + (size_t)(unsigned)rec.x * ldh);
            }
          }
        }
      }
    }
    // ---- the last receiver with edges, and any trailing receivers without
    while (rows_left > 0) {
      *reinterpret_cast<float4*>(out_row) = acc;
      out_row += ldo;
      acc = addn;
      --rows_left;
      if (rows_left > 1 && add) addn = ss_ld4(add_row);
      add_row += ldadd;
    }
    ss_wait<0>();
    __syncthreads();   // every row slice of this hop is stored (and visible to the block) before the next hop gathers it
  }
}

}  // namespace
}  // namespace dcb

using namespace dcb;

extern "C" int dc_spmm_stream(const int32_t* rowptr, const void* edges, const dc_hop_t* hops, int32_t num_hops, int64_t N, int32_t F,
                              const int32_t* tile_ptr, int64_t n_tiles, int32_t tile_nodes, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(N >= 0 && F >= 0 && num_hops >= 0, DC_EINVAL, "spmm_stream: negative size");
  if (N == 0 || F == 0 || num_hops == 0) return DC_OK;
  DC_REQUIRE(num_hops <= DC_MAX_CHAIN, DC_EINVAL, "spmm_stream: at most %d hops", DC_MAX_CHAIN);
  DC_REQUIRE(rowptr && hops && edges, DC_EINVAL, "spmm_stream: null pointer");
  DC_REQUIRE(F % 32 == 0 && (reinterpret_cast<uintptr_t>(edges) & 7) == 0, DC_ENOSUP,
             "spmm_stream: needs F %% 32 == 0 and packed 8-byte edge records (use dc_spmm_chain / dc_spmm_lean)");
  DC_REQUIRE(N < (1ll << 31), DC_ENOSUP, "spmm_stream: N exceeds 32-bit offsets");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  StreamHops a;
  for (int k = 0; k < DC_MAX_CHAIN; ++k) {
    const dc_hop_t& hp = hops[k < num_hops ? k : 0];
    if (k < num_hops) {
      DC_REQUIRE(hp.in && hp.out && hp.in != hp.out, DC_EINVAL, "spmm_stream: hop %d: null / aliased in and out", k);
      DC_REQUIRE(hp.ldin % 4 == 0 && hp.ldout % 4 == 0 && hp.ldin >= F && hp.ldout >= F && al16(hp.in) && al16(hp.out) &&
                     (!hp.add || (hp.ldadd % 4 == 0 && hp.ldadd >= F && al16(hp.add))),
                 DC_ENOSUP, "spmm_stream: hop %d: needs 16-byte aligned rows", k);
      DC_REQUIRE((uint64_t)N * (uint64_t)hp.ldout < (1ull << 32) && (uint64_t)N * (uint64_t)hp.ldin < (1ull << 32) &&
                     (!hp.add || (uint64_t)N * (uint64_t)hp.ldadd < (1ull << 32)),
                 DC_ENOSUP, "spmm_stream: hop %d: N*ld exceeds 32-bit element offsets", k);
    }
    a.in[k] = hp.in; a.add[k] = hp.add; a.out[k] = hp.out;
    a.ldin[k] = (unsigned)hp.ldin; a.ldadd[k] = (unsigned)hp.ldadd; a.ldout[k] = (unsigned)hp.ldout;
  }
  if (!tile_ptr) {
    DC_REQUIRE(tile_nodes > 0, DC_EINVAL, "spmm_stream: tile_nodes must be > 0 without tile_ptr");
    n_tiles = cdiv(N, tile_nodes);
  }
  DC_REQUIRE(n_tiles > 0, DC_EINVAL, "spmm_stream: no tiles");
  const int n_slices = F / 32;
  // Configurations (A/B by DCB200_K1_STREAM_CFG, scripts/k1_chain_lab.py): threads x gathers per batch (window = 2 batches)
  //   0: 768 x 4 (96 groups, 80 registers)   1: 1024 x 4 (128 groups, 64 registers)   2: 512 x 8 (64 groups, 16 in the window)
  //   3: 512 x 4 (128 registers: nothing spilled or rematerialised)   4: 640 x 4
  static int cfg = -1;
  if (cfg < 0) {
    const char* e = getenv("DCB200_K1_STREAM_CFG");
    cfg = e ? atoi(e) : 4;   // 640 x 4 measured best (profiles/r02_k1_experiments.txt)
    if (cfg < 0 || cfg > 4) cfg = 4;
  }
  static DeviceOnce carve;
  if (carve.first()) {   // static shared memory = record rings (264 B per group): ask for the smallest carve-out that holds them
    cudaFuncSetAttribute(spmm_stream_kernel<768, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, 14);
    cudaFuncSetAttribute(spmm_stream_kernel<1024, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, 28);
    cudaFuncSetAttribute(spmm_stream_kernel<512, 8>, cudaFuncAttributePreferredSharedMemoryCarveout, 14);
    cudaFuncSetAttribute(spmm_stream_kernel<512, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, 14);
    cudaFuncSetAttribute(spmm_stream_kernel<640, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, 14);
  }
  const unsigned grid = (unsigned)(n_tiles * n_slices);
  const int2* e2 = static_cast<const int2*>(edges);
  if (cfg == 1) spmm_stream_kernel<1024, 4><<<grid, 1024, 0, st>>>(rowptr, e2, a, num_hops, (int)N, tile_ptr, n_slices, tile_nodes);
  else if (cfg == 2) spmm_stream_kernel<512, 8><<<grid, 512, 0, st>>>(rowptr, e2, a, num_hops, (int)N, tile_ptr, n_slices, tile_nodes);
  else if (cfg == 3) spmm_stream_kernel<512, 4><<<grid, 512, 0, st>>>(rowptr, e2, a, num_hops, (int)N, tile_ptr, n_slices, tile_nodes);
  else if (cfg == 4) spmm_stream_kernel<640, 4><<<grid, 640, 0, st>>>(rowptr, e2, a, num_hops, (int)N, tile_ptr, n_slices, tile_nodes);
  else spmm_stream_kernel<768, 4><<<grid, 768, 0, st>>>(rowptr, e2, a, num_hops, (int)N, tile_ptr, n_slices, tile_nodes);
  DC_LAUNCH_CHECK();
  return DC_OK;
}
