// K1 v7 — "edge blocks": the CSR-by-receiver gather / segmented sum (spmm.cu) driven by a sliced-ELL
// (SELL-4-sigma) copy of the CSR that is built once per batch and reused by every hop of every layer.
//
// Why: the tile x 128-byte-slice kernels (v2..v6) are bound by instruction issue and by L1TEX wavefronts, not
// by HBM (profiles/r01_ncu_spmm_*.csv): 8 lanes share a receiver, so the 4 receivers of a warp read their
// edge records from 4 different places (2-4 wavefronts per 8-byte record load) and rows of unequal degree
// diverge.  Here
//   * a warp's work item is a UNIT of 4 receivers ("slots") whose edge records are interleaved
//     [pair of edges][slot][2 records]: one 128-bit load per lane fetches two records and the four groups of a
//     warp read one contiguous 64-byte segment (1 wavefront per 2 edges instead of 4-8);
//   * slots are sorted by degree (descending, stable) inside windows of 256 receivers of a tile, so the four
//     rows of a unit have (nearly) equal length: no divergence, padding bounded by E + 2N records;
//   * fp32 mul and add are issued as packed FMUL2 / FFMA2(m, 1.0, acc) (two IEEE-rounded lanes per
//     instruction; the multiplier 1.0 is a kernel argument so ptxas cannot contract mul+add into one fma):
//     half the FP instructions, same bits;
//   * one IMAD.WIDE per gather address, no shuffles, rows deeper than 64 edges continue from the packed CSR.
// The per-receiver summation order is still the CSR (= original edge) order with separately rounded mul and
// add, so results are bit-identical to dc_spmm and to the CPU reference order.
#include "common.cuh"
#include "scan.cuh"

namespace {
using namespace dcb;
typedef unsigned long long u64;

constexpr int BLK_WINDOW = 256;   // slots sorted together (sigma)
constexpr int BLK_DEPTH = 64;     // edges per row held in blocks; deeper rows continue from the CSR records

// ---------------------------------------------------------------------------------- build
// one CTA per tile; each window of 256 receivers is ranked by degree (stable, descending)
__global__ void __launch_bounds__(BLK_WINDOW)
blk_order_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ tile_ptr,
                 const int32_t* __restrict__ tile_unit_ptr, int4* __restrict__ slots, uint32_t* __restrict__ unit_cnt, int C) {
  __shared__ int s_key[BLK_WINDOW];
  __shared__ int s_node[BLK_WINDOW];
  __shared__ int s_deg[BLK_WINDOW];
  const int tile = blockIdx.x, tid = threadIdx.x;
  const int t0 = tile_ptr[tile], n_t = tile_ptr[tile + 1] - t0;
  const int ub = tile_unit_ptr[tile], units_t = tile_unit_ptr[tile + 1] - ub;
  for (int w0 = 0; w0 < n_t; w0 += BLK_WINDOW) {
    const int i = w0 + tid;
    const bool valid = i < n_t;
    int deg = -1;
    if (valid) deg = rowptr[t0 + i + 1] - rowptr[t0 + i];
    const int key = min(deg, BLK_DEPTH);
    s_key[tid] = key;
    __syncthreads();
    int rank = 0;
#pragma unroll 8
    for (int j = 0; j < BLK_WINDOW; ++j) {
      const int kj = s_key[j];
      rank += (kj > key) || (kj == key && j < tid);
    }
    s_node[rank] = valid ? t0 + i : -1;
    s_deg[rank] = max(deg, 0);
    __syncthreads();
    const int slot_in_tile = w0 + tid;
    if (slot_in_tile < units_t * C) {
      const int q = tid & ~(C - 1);
      const int md = (min(s_deg[q], BLK_DEPTH) + 1) & ~1;   // unit depth: max degree (first of the sorted C), even
      const int mn = min(s_deg[q + C - 1], BLK_DEPTH);      // unit minimum (pads have degree 0)
      slots[(size_t)ub * C + slot_in_tile] = make_int4(s_node[tid], s_deg[tid], 0, md | (mn << 8));
      if ((tid & (C - 1)) == 0) unit_cnt[ub + slot_in_tile / C] = (uint32_t)(md * C);
    }
    __syncthreads();
  }
}

// one warp per unit: copy the first min(deg, 64) records of its C rows into the interleaved layout
// [pair of edges][slot][2]; bit 16 of the slot's depth word flags units with a source outside their own tile
__global__ void __launch_bounds__(256)
blk_fill_kernel(const int32_t* __restrict__ rowptr, const int2* __restrict__ edges, int4* __restrict__ slots,
                const uint32_t* __restrict__ unit_off, int64_t n_units, int2* __restrict__ recs, int64_t rec_cap,
                int* __restrict__ status, int C, const int32_t* __restrict__ tile_ptr,
                const int32_t* __restrict__ tile_unit_ptr, int n_tiles, int32_t* __restrict__ tile_rec_ptr) {
  const int64_t unit = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (unit >= n_units) return;
  const int lane = threadIdx.x & 31, g = lane & (C - 1);
  const int4 s = slots[unit * C + g];
  const int md = s.w & 0xff;
  const uint32_t roff = unit_off[unit];
  if ((int64_t)roff + (int64_t)md * C > rec_cap) {   // cannot happen with the documented capacity bound
    if (lane == 0) atomicExch(status, 1);
    return;
  }
  int lo = 0, hi = n_tiles;   // tile of this unit: last t with tile_unit_ptr[t] <= unit
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tile_unit_ptr[mid] <= unit) lo = mid; else hi = mid;
  }
  const int t0 = tile_ptr[lo], t1 = tile_ptr[lo + 1];
  if (lane == 0) {   // record offsets at the tile boundaries (empty tiles share the offset of the next unit)
    if (unit == tile_unit_ptr[lo])
      for (int t = lo; t >= 0 && tile_unit_ptr[t] == unit; --t) tile_rec_ptr[t] = (int32_t)roff;
    if (unit == n_units - 1)
      for (int t = n_tiles; t > lo && tile_unit_ptr[t] == n_units; --t) tile_rec_ptr[t] = (int32_t)(roff + (uint32_t)(md * C));
  }
  const int beg = s.x >= 0 ? rowptr[s.x] : 0;
  const int dcap = min(s.y, BLK_DEPTH);
  bool outside = false;
  for (int u = lane / C; u < md; u += 32 / C) {
    int2 r = make_int2(0, 0);
    if (u < dcap) {
      r = edges[beg + u];
      outside |= (r.x < t0) || (r.x >= t1);
    }
    recs[(size_t)roff + (size_t)(u >> 1) * (2 * C) + g * 2 + (u & 1)] = r;
  }
  if (s.y > BLK_DEPTH) outside = true;   // deep rows are finished from the CSR by the checked path
  const unsigned any_out = __ballot_sync(0xffffffffu, outside);
  if (lane < C) {
    slots[unit * C + lane].z = (int)roff;
    if (any_out) slots[unit * C + lane].w = s.w | (1 << 16);
  }
}

// ---------------------------------------------------------------------------------- gather
__device__ __forceinline__ u64 pack2(float a, float b) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }

// acc = fl(acc + fl(w * v)) on two packed fp32 lanes; `one2` = {1.0f, 1.0f} from a kernel argument
__device__ __forceinline__ void mul_add2(u64& acc, u64 v, u64 w2, u64 one2) {
  u64 m;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(m) : "l"(v), "l"(w2));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(m), "l"(one2));
}
__device__ __forceinline__ void add2(u64& acc, u64 v, u64 one2) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(v), "l"(one2)); }

__device__ __forceinline__ ulonglong2 ldg_row(const char* p) { return __ldg(reinterpret_cast<const ulonglong2*>(p)); }
__device__ __forceinline__ int4 ldg_rec(const int4* p) {
  int4 v;
  asm volatile("ld.global.nc.L1::evict_first.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ int4 ld_slot(const int4* p) {
  int4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// One CTA per SM owns (tile of receivers) x (128-byte feature slice): the tile's source-row slices (~2000 x 128 B)
// are reused out of L1.  A warp takes one unit (4 receivers x 8 lanes x float4) per iteration: 4 record loads
// (128-bit, one 64-byte segment per warp), 8 row gathers in flight per lane, 32 packed FP instructions.
// Rows shorter than the unit's depth read pad records {0, 0} (row 0, never accumulated).
template <int THREADS, bool FULL>
__global__ void __launch_bounds__(THREADS, 1)
spmm_blk_kernel(const int4* __restrict__ slots, const int4* __restrict__ recs, const int32_t* __restrict__ rowptr,
                const int2* __restrict__ edges, const int32_t* __restrict__ tile_unit_ptr, int n_tile_slices, int n_slices,
                const float* __restrict__ self_w, const float* __restrict__ h, unsigned ldb /* bytes */, float* __restrict__ out,
                unsigned ldob, const float* __restrict__ add, unsigned ldaddb, int F, int self_loop,
                const float* __restrict__ bias, int relu, float one, int prefetch, const int32_t* __restrict__ tile_ptr,
                const int32_t* __restrict__ tile_rec_ptr, int l2_ahead, int l2_sectors) {
  constexpr int UPC = THREADS / 32;   // units per CTA iteration
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 3, gl = lane & 7;
  const u64 one2 = pack2(one, one);
  const int4 pad = make_int4(-1, 0, 0, 0);

  for (int ts = blockIdx.x; ts < n_tile_slices; ts += gridDim.x) {
    const int slice = ts % n_slices, tile = ts / n_slices;
    if (l2_ahead > 0) {
      // The gather loop below finds the compulsory HBM traffic one dependent batch at a time (slot -> records ->
      // rows), far too little of it in flight to keep HBM busy.  So while working on this tile slice, pull the
      // tile slice this CTA (or, for a one-shot grid, the SM) takes `l2_ahead` steps later into L2 as one deep
      // stream: its 128-byte row slices, its share of the tile's edge records and slot records.
      // l2_ahead == 3: this tile slice itself (the stream runs ahead of the gather loop inside the CTA)
      const int tsn = l2_ahead == 3 ? ts : ts + l2_ahead * ((int)gridDim.x >= n_tile_slices ? kSMs : (int)gridDim.x);
      if (tsn < n_tile_slices) {
        const int sl = tsn % n_slices, tl = tsn / n_slices;
        const int r0 = tile_ptr[tl], r1 = tile_ptr[tl + 1];
        const char* base = reinterpret_cast<const char*>(h + sl * 32);
        const int nsec = l2_sectors ? min(4, (F - sl * 32) / 8) : 1;   // prefetches per row slice (one per 32-byte sector, or one per line)
        for (int r = r0 + (int)threadIdx.x; r < r1; r += THREADS) {
          const char* p = base + (size_t)(unsigned)r * ldb;
          for (int q = 0; q < nsec; ++q) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + q * 32));
        }
        // records / slots of the tile: split over the slices (every slice of a tile runs at about the same time)
        const size_t q0 = (size_t)(unsigned)tile_rec_ptr[tl] * 8, q1 = (size_t)(unsigned)tile_rec_ptr[tl + 1] * 8;
        for (size_t o = q0 + ((size_t)threadIdx.x * n_slices + sl) * 128; o < q1; o += (size_t)THREADS * n_slices * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(recs) + o));
        const size_t z0 = (size_t)tile_unit_ptr[tl] * 64, z1 = (size_t)tile_unit_ptr[tl + 1] * 64;
        for (size_t o = z0 + ((size_t)threadIdx.x * n_slices + sl) * 128; o < z1; o += (size_t)THREADS * n_slices * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(slots) + o));
      }
    }
    const int col = slice * 32 + gl * 4;
    if (!FULL && col >= F) continue;   // lane idle for the tail slice (no collectives below)
    const int u1 = tile_unit_ptr[tile + 1];
    int unit = tile_unit_ptr[tile] + warp;
    if (unit >= u1) continue;          // warp-uniform
    const char* __restrict__ hb = reinterpret_cast<const char*>(h + col);
    int4 s = ld_slot(slots + (size_t)unit * 4 + g);
    for (; unit < u1; unit += UPC) {
      const int4 sn = unit + UPC < u1 ? ld_slot(slots + (size_t)(unit + UPC) * 4 + g) : pad;   // slot record one unit ahead
      const int node = s.x, deg = s.y;
      const int md = s.w & 0xff, mn = (s.w >> 8) & 0xff;
      const int4* rp = recs + ((size_t)(unsigned)s.z >> 1) + g;
      // the warp's next unit lies UPC units further; with equal depths its records start UPC * md * 4 records ahead
      if (prefetch && lane < md)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(rp - g) + (size_t)(UPC * md) * 32 + lane * 32));
      u64 a0 = 0ull, a1 = 0ull;
      if (add != nullptr && node >= 0) {
        const ulonglong2 t = __ldcs(reinterpret_cast<const ulonglong2*>(reinterpret_cast<const char*>(add + col) + (size_t)(unsigned)node * ldaddb));
        a0 = t.x; a1 = t.y;
      }
      int b = 0;
#pragma unroll 1
      for (; b + 8 <= md; b += 8, rp += 16) {   // full chunks: 8 gathers in flight (warp-uniform trip count)
        int4 r[4];
        ulonglong2 v[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) r[j] = ldg_rec(rp + j * 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[2 * j] = ldg_row(hb + (size_t)(unsigned)r[j].x * ldb);
          v[2 * j + 1] = ldg_row(hb + (size_t)(unsigned)r[j].z * ldb);
        }
        if (mn >= b + 8) {   // warp-uniform: all four rows have 8 more edges — no predicates
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float w0 = __int_as_float(r[j].y), w1 = __int_as_float(r[j].w);
            const u64 p0 = pack2(w0, w0), p1 = pack2(w1, w1);
            mul_add2(a0, v[2 * j].x, p0, one2);
            mul_add2(a1, v[2 * j].y, p0, one2);
            mul_add2(a0, v[2 * j + 1].x, p1, one2);
            mul_add2(a1, v[2 * j + 1].y, p1, one2);
          }
        } else {
          const int cnt = deg - b;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (2 * j < cnt) {
              const float w0 = __int_as_float(r[j].y);
              const u64 p0 = pack2(w0, w0);
              mul_add2(a0, v[2 * j].x, p0, one2);
              mul_add2(a1, v[2 * j].y, p0, one2);
            }
            if (2 * j + 1 < cnt) {
              const float w1 = __int_as_float(r[j].w);
              const u64 p1 = pack2(w1, w1);
              mul_add2(a0, v[2 * j + 1].x, p1, one2);
              mul_add2(a1, v[2 * j + 1].y, p1, one2);
            }
          }
        }
      }
#pragma unroll 1
      for (; b < md; b += 2, rp += 4) {   // tail of the unit, one pair of edges at a time
        const int4 r = ldg_rec(rp);
        const ulonglong2 v0 = ldg_row(hb + (size_t)(unsigned)r.x * ldb);
        const ulonglong2 v1 = ldg_row(hb + (size_t)(unsigned)r.z * ldb);
        if (b < deg) {
          const float w0 = __int_as_float(r.y);
          const u64 p0 = pack2(w0, w0);
          mul_add2(a0, v0.x, p0, one2);
          mul_add2(a1, v0.y, p0, one2);
        }
        if (b + 1 < deg) {
          const float w1 = __int_as_float(r.w);
          const u64 p1 = pack2(w1, w1);
          mul_add2(a0, v1.x, p1, one2);
          mul_add2(a1, v1.y, p1, one2);
        }
      }
      if (deg > BLK_DEPTH) {   // deep rows: the rest of the row, in order, from the packed CSR
        const int beg = rowptr[node];
        for (int p = beg + BLK_DEPTH; p < beg + deg; ++p) {
          const int2 e = __ldg(edges + p);
          const ulonglong2 v = ldg_row(hb + (size_t)(unsigned)e.x * ldb);
          const float w = __int_as_float(e.y);
          const u64 w2 = pack2(w, w);
          mul_add2(a0, v.x, w2, one2);
          mul_add2(a1, v.y, w2, one2);
        }
      }
      if (node >= 0) {
        if (self_loop) {
          const ulonglong2 v = ldg_row(hb + (size_t)(unsigned)node * ldb);
          const float w = self_w[node];
          const u64 w2 = pack2(w, w);
          mul_add2(a0, v.x, w2, one2);
          mul_add2(a1, v.y, w2, one2);
        }
        if (bias) {
          const ulonglong2 b4 = __ldg(reinterpret_cast<const ulonglong2*>(bias + col));
          add2(a0, b4.x, one2);
          add2(a1, b4.y, one2);
        }
        float4 o;
        unpack2(a0, o.x, o.y);
        unpack2(a1, o.z, o.w);
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        __stcs(reinterpret_cast<float4*>(reinterpret_cast<char*>(out + col) + (size_t)(unsigned)node * ldob), o);
      }
      s = sn;
    }
  }
}
}  // namespace

extern "C" size_t dc_blocks_workspace_bytes(int64_t n_units) {
  if (n_units < 0) n_units = 0;
  return align_up((size_t)(n_units + 1) * 4, 256) + align_up((size_t)scan_num_blocks(n_units) * 4 + 4, 256) + 256;
}

extern "C" int64_t dc_blocks_record_capacity(int64_t num_nodes, int64_t num_edges, int64_t n_tiles) {
  // per window of 256 slots (sorted by degree, rows capped at 64): sum_u 4*even(max_u) <= sum deg + 4*64 + 256
  const int64_t windows = num_nodes / BLK_WINDOW + n_tiles + 1;
  return num_edges + windows * (8 * BLK_DEPTH + BLK_WINDOW) + 64;   // valid for unit sizes 4 and 8
}

extern "C" int dc_blocks_build(const int32_t* rowptr, const void* edges, const int32_t* tile_ptr, const int32_t* tile_unit_ptr,
                               int64_t n_tiles, int64_t n_units, int32_t unit_size, void* slots, void* recs, int64_t rec_capacity,
                               int32_t* tile_rec_ptr, int32_t* status, void* workspace, size_t workspace_bytes,
                               dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(n_tiles >= 0 && n_units >= 0, DC_EINVAL, "blocks_build: negative size");
  DC_REQUIRE(unit_size == 4 || unit_size == 8, DC_EINVAL, "blocks_build: unit_size must be 4 or 8");
  if (n_tiles == 0 || n_units == 0) return DC_OK;
  DC_REQUIRE(rowptr && tile_ptr && tile_unit_ptr && slots && recs && tile_rec_ptr && status && workspace, DC_EINVAL,
             "blocks_build: null pointer");
  DC_REQUIRE(workspace_bytes >= dc_blocks_workspace_bytes(n_units), DC_EWORKSPACE, "blocks_build: workspace too small");
  DC_REQUIRE(n_units < (1ll << 29) && rec_capacity < (1ll << 32), DC_ENOSUP, "blocks_build: too many units / records");
  Carver cv(workspace);
  uint32_t* unit_cnt = cv.take<uint32_t>(n_units + 1);
  uint32_t* bsum = cv.take<uint32_t>(scan_num_blocks(n_units) + 1);
  blk_order_kernel<<<(unsigned)n_tiles, BLK_WINDOW, 0, st>>>(rowptr, tile_ptr, tile_unit_ptr, static_cast<int4*>(slots), unit_cnt,
                                                             unit_size);
  DC_LAUNCH_CHECK();
  if (int rc = exclusive_scan_u32(unit_cnt, n_units, bsum, st)) return rc;
  blk_fill_kernel<<<(unsigned)cdiv(n_units, 8), 256, 0, st>>>(rowptr, static_cast<const int2*>(edges), static_cast<int4*>(slots),
                                                              unit_cnt, n_units, static_cast<int2*>(recs), rec_capacity, status, unit_size,
                                                              tile_ptr, tile_unit_ptr, (int)n_tiles, tile_rec_ptr);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

extern "C" int dc_spmm_blocks(const void* slots, const void* recs, const int32_t* rowptr, const void* edges, const int32_t* tile_ptr,
                              const int32_t* tile_unit_ptr, const int32_t* tile_rec_ptr, int64_t n_tiles, const float* self_w, const float* h, int64_t ldh,
                              float* out, int64_t ldo, const float* add, int64_t ldadd, int32_t F, int self_loop,
                              const float* bias, int relu, int flags, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(n_tiles >= 0 && F >= 0, DC_EINVAL, "spmm_blocks: negative size");
  if (n_tiles == 0 || F == 0) return DC_OK;
  DC_REQUIRE(slots && recs && rowptr && tile_ptr && tile_unit_ptr && tile_rec_ptr && h && out, DC_EINVAL, "spmm_blocks: null pointer");
  DC_REQUIRE(h != out, DC_EINVAL, "spmm_blocks: out must not alias h");
  DC_REQUIRE(!self_loop || self_w, DC_EINVAL, "spmm_blocks: self_loop needs self_w");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  DC_REQUIRE((F % 4 == 0) && (ldh % 4 == 0) && (ldo % 4 == 0) && ldh >= F && ldo >= F && al16(h) && al16(out) && al16(recs) &&
                 al16(slots) && (!add || ((ldadd % 4 == 0) && ldadd >= F && al16(add))) && (!bias || al16(bias)),
             DC_ENOSUP, "spmm_blocks: needs F %% 4 == 0 and 16-byte aligned rows (use dc_spmm)");
  DC_REQUIRE(ldh < (1ll << 30) && ldo < (1ll << 30) && ldadd < (1ll << 30), DC_ENOSUP, "spmm_blocks: row stride too large");
  const int n_slices = (F + 31) / 32;
  const int64_t n_ts = n_tiles * n_slices;
  DC_REQUIRE(n_ts < (1ll << 31), DC_ENOSUP, "spmm_blocks: too many tile slices");
  const unsigned grid = (flags & 1) ? (unsigned)(n_ts < sm_count() ? n_ts : sm_count()) : (unsigned)n_ts;
  const bool full = F % 32 == 0, t768 = (flags >> 2) & 1;
  const int prefetch = (flags >> 1) & 1;
  const int l2_ahead = (flags >> 3) & 3;   // 0 = off, else distance (in grid strides) of the L2 prefetch stream
#define DC_BLK_LAUNCH(T, FU)                                                                                                   \
  do {                                                                                                                         \
    cudaFuncSetAttribute(spmm_blk_kernel<T, FU>, cudaFuncAttributePreferredSharedMemoryCarveout, 0); /* idempotent */          \
    spmm_blk_kernel<T, FU><<<grid, T, 0, st>>>(static_cast<const int4*>(slots), static_cast<const int4*>(recs), rowptr,        \
                                               static_cast<const int2*>(edges), tile_unit_ptr, (int)n_ts, n_slices, self_w, h, \
                                               (unsigned)(ldh * 4), out, (unsigned)(ldo * 4), add, (unsigned)(ldadd * 4), F,   \
                                               self_loop, bias, relu, 1.0f, prefetch, tile_ptr, tile_rec_ptr, l2_ahead,        \
                                               (flags >> 5) & 1);                                                              \
  } while (0)
  if (t768) { if (full) DC_BLK_LAUNCH(768, true); else DC_BLK_LAUNCH(768, false); }
  else { if (full) DC_BLK_LAUNCH(1024, true); else DC_BLK_LAUNCH(1024, false); }
#undef DC_BLK_LAUNCH
  DC_LAUNCH_CHECK();
  return DC_OK;
}

