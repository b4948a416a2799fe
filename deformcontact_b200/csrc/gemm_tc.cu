// K2 (tensor-core variant) — tcgen05 kind::tf32 GEMM with a 3-term error-compensated split
// ("3xTF32"): C = A_hi*B_hi + A_hi*B_lo + A_lo*B_hi, fp32 accumulation in TMEM.  A single
// TF32 pass keeps 10 mantissa bits (~5e-4 per product) and fails the 1e-5 parity bar; the split
// restores fp32-class accuracy at 1/3 of the TF32 rate, still several times the FP32 SIMT peak.
//
//   C[M,N] = act( sum_s A_s[M,K_s] * B_s + bias ) (+C),  A_s row-major (K-major), any B layout.
//
// Structure (one persistent CTA per SM, 384 threads, warp-specialised):
//   warp 0      TMA producer: A tile (raw fp32) + pre-split B_hi / B_lo tiles, 128B-swizzled,
//               into a 2-stage smem ring (mbarrier complete_tx).
//   warps 8-11  split warps: in place A -> A_hi (mantissa masked to tf32), A_lo = A - A_hi into a
//               second buffer (same swizzled positions), fence.proxy.async, arrive.
//   warp 1      MMA issuer: one elected thread issues 12 tcgen05.mma (M=128, N<=256, K=8) per
//               32-wide k-block; tcgen05.commit frees the stage.  The tensor core truncates its
//               fp32 accumulator on every MMA (measured: error grows with the number of
//               accumulation steps), so the two small cross terms go to a SECOND TMEM accumulator
//               and the main one sees 3x fewer steps; they are summed in fp32 RN in the epilogue.
//   warp 2      TMEM allocator (512 columns = main + cross-term accumulator, 256 columns each).
//   warps 4-7   epilogue: tcgen05.ld (32 lanes x 32 columns) of both accumulators, add, + bias,
//               ReLU, transpose through a padded smem tile, 128-bit coalesced row stores.
// B (the layer weights, <= 1 MB) is split into hi/lo and packed K-major by a tiny pre-kernel.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace dcb {
namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;           // 32 fp32 = 128 bytes = one swizzle atom row
constexpr int TC_STAGES = 2;
// The tensor core truncates (rounds toward zero) its fp32 TMEM accumulator on every MMA: a systematic shrink
// proportional to the chain length.  The main accumulator is therefore drained into C (fp32 RN adds) every
// TC_DRAIN_KB k-blocks (= 4 * TC_DRAIN_KB accumulation steps); measured r01: 128-step chains give 3-4e-6
// relative error with a bias that the model's unscaled softmax attention amplifies ~30x downstream.
constexpr int TC_DRAIN_KB = 8;
constexpr int TC_THREADS = 384;
constexpr int TC_MAX_N = 256;
constexpr int A_TILE_BYTES = TC_BM * TC_BK * 4;        // 16 KB
constexpr int B_TILE_BYTES = TC_MAX_N * TC_BK * 4;     // 32 KB (sized for N = 256)
constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;  // A_hi(raw), A_lo, B_hi, B_lo = 96 KB
constexpr int EPI_ROW = 36;                            // padded staging row (floats): conflict-free 128-bit access
constexpr int EPI_STAGE_FLOATS = 4 * 32 * EPI_ROW;
constexpr int SMEM_BYTES = TC_STAGES * STAGE_BYTES + EPI_STAGE_FLOATS * 4 + 256 + 1024;  // + barriers + align slack

struct TcParams {
  int nseg;
  int kb_seg[4];  // k-blocks per segment
  int kb_total;
  int M, N;
  float* C;
  long long ldc;
  const float* bias;
  int relu, accumulate;
  int m_tiles;
  int drain_kb;   // main-accumulator chain length in k-blocks
  int n_chunks;   // output columns are covered in chunks of TC_MAX_N (N > 256: score-matrix shaped GEMMs)
  int n_mma;      // MMA N per chunk (N itself when N <= chunk width, else the chunk width with TMA zero fill past N)
  int n_chunk;    // chunk width: 256, or 128 when that is what it takes to give every SM a tile
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);  // start address, 16-byte units
  d |= (uint64_t)1 << 16;                   // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapA3,
               const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  float* epi_stage = reinterpret_cast<float*>(base_ptr + TC_STAGES * STAGE_BYTES);
  const uint32_t bar_base = base + TC_STAGES * STAGE_BYTES + EPI_STAGE_FLOATS * 4;
  // barriers (8 B each): full[2], xform[2], empty[2], tmem_full[2], tmem_empty[2], then tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto xform_bar = [&](int s) { return bar_base + 16u + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 32u + 8u * s; };
  auto tfull_bar = [&](int a) { return bar_base + 48u + 8u * a; };
  auto tempty_bar = [&](int a) { return bar_base + 64u + 8u * a; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(base_ptr + TC_STAGES * STAGE_BYTES + EPI_STAGE_FLOATS * 4 + 80);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(xform_bar(s), 128);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar(0), 1);
    mbar_init(tempty_bar(0), 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_smem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int N = p.N;
  const int total_tiles = p.m_tiles * p.n_chunks;   // tile t -> (m tile t / n_chunks, column chunk t % n_chunks)
  const uint32_t b_bytes = (uint32_t)p.n_mma * TC_BK * 4;
  const uint32_t stage_tx = A_TILE_BYTES + 2 * b_bytes;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const CUtensorMap* maps[4] = {&mapA0, &mapA1, &mapA2, &mapA3};
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.n_chunks) * TC_BM;
        const int n0 = (tile % p.n_chunks) * p.n_chunk;
        int kb = 0;
        for (int s = 0; s < p.nseg; ++s) {
          for (int j = 0; j < p.kb_seg[s]; ++j, ++kb, ++it) {
            const int st = it % TC_STAGES;
            const uint32_t ph = (it / TC_STAGES) & 1;
            mbar_wait(empty_bar(st), ph ^ 1);
            const uint32_t sbase = base + st * STAGE_BYTES;
            mbar_expect_tx(full_bar(st), stage_tx);
            tma_load_2d(sbase, maps[s], full_bar(st), j * TC_BK, m0);
            tma_load_2d(sbase + 2 * A_TILE_BYTES, &mapBhi, full_bar(st), kb * TC_BK, n0);
            tma_load_2d(sbase + 2 * A_TILE_BYTES + B_TILE_BYTES, &mapBlo, full_bar(st), kb * TC_BK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.n_mma >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int it = 0, dcount = 0;
      const uint32_t d_main = tmem_base, d_cross = tmem_base + TC_MAX_N;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        for (int kb0 = 0; kb0 < p.kb_total; kb0 += p.drain_kb, ++dcount) {
          const int kb1 = min(p.kb_total, kb0 + p.drain_kb);
          mbar_wait(tempty_bar(0), (dcount & 1) ^ 1);
          tc_fence_after();
          for (int kb = kb0; kb < kb1; ++kb, ++it) {
            const int st = it % TC_STAGES;
            const uint32_t ph = (it / TC_STAGES) & 1;
            mbar_wait(full_bar(st), ph);
            mbar_wait(xform_bar(st), ph);
            tc_fence_after();
            const uint32_t sbase = base + st * STAGE_BYTES;
            const uint64_t a_hi = make_desc(sbase), a_lo = make_desc(sbase + A_TILE_BYTES);
            const uint64_t b_hi = make_desc(sbase + 2 * A_TILE_BYTES), b_lo = make_desc(sbase + 2 * A_TILE_BYTES + B_TILE_BYTES);
#pragma unroll
            for (int kk = 0; kk < TC_BK / 8; ++kk) {
              const uint64_t adv = (uint64_t)(kk * 32 >> 4);  // +32 bytes per K=8 step inside the swizzle atom
              umma_tf32(d_cross, a_lo + adv, b_hi + adv, idesc, (kb > 0 || kk > 0) ? 1u : 0u);   // whole-tile chain (tiny values)
              umma_tf32(d_cross, a_hi + adv, b_lo + adv, idesc, 1u);
              umma_tf32(d_main, a_hi + adv, b_hi + adv, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);  // restarted every drain
            }
            umma_commit(empty_bar(st));
          }
          umma_commit(tfull_bar(0));
        }
      }
    }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ split warps: A -> A_hi (in place), A_lo
    const int t = threadIdx.x - 256;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < p.kb_total; ++kb, ++it) {
        const int st = it % TC_STAGES;
        const uint32_t ph = (it / TC_STAGES) & 1;
        mbar_wait(full_bar(st), ph);
        uint4* a = reinterpret_cast<uint4*>(base_ptr + st * STAGE_BYTES);
        uint4* lo = reinterpret_cast<uint4*>(base_ptr + st * STAGE_BYTES + A_TILE_BYTES);
#pragma unroll
        for (int i = 0; i < A_TILE_BYTES / 16 / 128; ++i) {
          const int idx = t + i * 128;
          uint4 v = a[idx];
          uint4 h = make_uint4(v.x & 0xFFFFE000u, v.y & 0xFFFFE000u, v.z & 0xFFFFE000u, v.w & 0xFFFFE000u);
          uint4 l;
          l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x));
          l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y));
          l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z));
          l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w));
          a[idx] = h;
          lo[idx] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(xform_bar(st));
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp - 4;  // TMEM lane quarter == warp % 4
    float* stg = epi_stage + q * 32 * EPI_ROW;
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                        (!p.bias || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
    int dcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
     const int n0 = (tile % p.n_chunks) * p.n_chunk;
     const int ncols = min(p.n_chunk, N - n0);
     for (int kb0 = 0; kb0 < p.kb_total; kb0 += p.drain_kb, ++dcount) {
      const bool first_drain = kb0 == 0, last_drain = kb0 + p.drain_kb >= p.kb_total;
      const bool acc_c = first_drain ? (p.accumulate != 0) : true;   // later drains add onto this tile's partial C
      const bool do_relu = last_drain && p.relu;
      mbar_wait(tfull_bar(0), dcount & 1);
      tc_fence_after();
      const int row0 = (tile / p.n_chunks) * TC_BM + q * 32;
      for (int c = 0; c < ncols; c += 32) {
        uint32_t r[32], x[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c;
#define DC_TMEM_LD32(R, ADDR)                                                                                            \
  asm volatile(                                                                                                         \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                         \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, " \
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                                                                 \
      : "=r"(R[0]), "=r"(R[1]), "=r"(R[2]), "=r"(R[3]), "=r"(R[4]), "=r"(R[5]), "=r"(R[6]), "=r"(R[7]), "=r"(R[8]),    \
        "=r"(R[9]), "=r"(R[10]), "=r"(R[11]), "=r"(R[12]), "=r"(R[13]), "=r"(R[14]), "=r"(R[15]), "=r"(R[16]),         \
        "=r"(R[17]), "=r"(R[18]), "=r"(R[19]), "=r"(R[20]), "=r"(R[21]), "=r"(R[22]), "=r"(R[23]), "=r"(R[24]),        \
        "=r"(R[25]), "=r"(R[26]), "=r"(R[27]), "=r"(R[28]), "=r"(R[29]), "=r"(R[30]), "=r"(R[31])                       \
      : "r"(ADDR))
        DC_TMEM_LD32(r, taddr);
        if (last_drain) {   // the cross-term accumulator is read once, at the end of the tile
          DC_TMEM_LD32(x, taddr + TC_MAX_N);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = 0u;
        }
#undef DC_TMEM_LD32
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 v;
          v.x = __uint_as_float(r[j]) + __uint_as_float(x[j]);
          v.y = __uint_as_float(r[j + 1]) + __uint_as_float(x[j + 1]);
          v.z = __uint_as_float(r[j + 2]) + __uint_as_float(x[j + 2]);
          v.w = __uint_as_float(r[j + 3]) + __uint_as_float(x[j + 3]);
          *reinterpret_cast<float4*>(stg + lane * EPI_ROW + j) = v;
        }
        __syncwarp();
        const int c4 = (lane & 7) * 4;
        const int col = n0 + c + c4;
        if (vec_ok && col + 3 < N) {
          const float4 bv = (p.bias && first_drain) ? *reinterpret_cast<const float4*>(p.bias + col) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int it4 = 0; it4 < 8; ++it4) {
            const int rr = it4 * 4 + (lane >> 3);
            const int row = row0 + rr;
            if (row < p.M) {
              float4 v = *reinterpret_cast<const float4*>(stg + rr * EPI_ROW + c4);
              v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
              float4* dst = reinterpret_cast<float4*>(p.C + (long long)row * p.ldc + col);
              if (acc_c) { const float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
              if (do_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
              *dst = v;
            }
          }
        } else {
          for (int it4 = 0; it4 < 8; ++it4) {
            const int rr = it4 * 4 + (lane >> 3);
            const int row = row0 + rr;
            for (int e = 0; e < 4; ++e) {
              if (row < p.M && col + e < N) {
                float v = stg[rr * EPI_ROW + c4 + e] + ((p.bias && first_drain) ? p.bias[col + e] : 0.f);
                float* dst = p.C + (long long)row * p.ldc + col + e;
                if (acc_c) v += *dst;
                if (do_relu) v = fmaxf(v, 0.f);
                *dst = v;
              }
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(0));
     }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// K3 — weight-gradient GEMM  C[M,N] = A^T B  with A stored [K, M], B stored [K, N] row-major and a very
// long contraction K (= number of nodes).  Both operands are "MN-major" for the tensor core: the TMA
// boxes are [32 k-rows x 128 B] and the shared-memory descriptors use the MN-major SWIZZLE_128B canonical
// layout (LBO = 4096 B between 32-element chunks along M/N, SBO = 1024 B between groups of 8 k).
// Neither operand can be pre-split (both are activations), so the split warps produce hi / lo for both
// tiles in shared memory.  The contraction is cut into chunks of TN_CHUNK_KB k-blocks; every CTA adds
// its chunks' partial tiles (fp32 RN, fixed order) into a private slab of the workspace and a second
// kernel sums the slabs in CTA order  =>  deterministic, and each TMEM accumulation chain stays short
// (the tensor core truncates its accumulator on every MMA).
constexpr int TN_CHUNK_KB = 8;   // 8 k-blocks = 256 contraction elements (32 steps) per TMEM accumulation chain

struct TnParams {
  int M, N, K;      // C is [M, N]; contraction length K
  int m_tiles, chunks, items;
  float* partial;   // [gridDim.x, m_tiles * 128, N]
  unsigned lbo, sbo, kadv, major_bits, layout;  // shared-memory descriptor parameters (bytes)
};

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;  // leading byte offset
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;  // stride byte offset
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;           // 1 = SWIZZLE_128B_BASE32B (the only MN-major layout for tf32)
  return d;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_tn_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  float* epi_stage = reinterpret_cast<float*>(base_ptr + TC_STAGES * STAGE_BYTES);
  const uint32_t bar_base = base + TC_STAGES * STAGE_BYTES + EPI_STAGE_FLOATS * 4;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto xform_bar = [&](int s) { return bar_base + 16u + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 32u + 8u * s; };
  const uint32_t tfull_bar = bar_base + 48u, tempty_bar = bar_base + 64u;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(base_ptr + TC_STAGES * STAGE_BYTES + EPI_STAGE_FLOATS * 4 + 80);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(xform_bar(s), 128);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_smem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int N = p.N;
  const int n_chunks32 = N / 32;                      // 32-element chunks along N (N % 32 == 0)
  const uint32_t stage_tx = (uint32_t)(4 + n_chunks32) * 4096u;
  const int kb_total = (p.K + TC_BK - 1) / TC_BK;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int w = blockIdx.x; w < p.items; w += gridDim.x) {
        const int mt = w % p.m_tiles, ch = w / p.m_tiles;
        const int kb0 = ch * TN_CHUNK_KB, kb1 = min(kb_total, kb0 + TN_CHUNK_KB);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int st = it % TC_STAGES;
          const uint32_t ph = (it / TC_STAGES) & 1;
          mbar_wait(empty_bar(st), ph ^ 1);
          const uint32_t sbase = base + st * STAGE_BYTES;
          mbar_expect_tx(full_bar(st), stage_tx);
#pragma unroll
          for (int c = 0; c < 4; ++c) tma_load_2d(sbase + c * 4096, &mapA, full_bar(st), mt * TC_BM + c * 32, kb * TC_BK);
          for (int c = 0; c < n_chunks32; ++c)
            tma_load_2d(sbase + 2 * A_TILE_BYTES + c * 4096, &mapB, full_bar(st), c * 32, kb * TC_BK);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // a_major = b_major = 1 (MN-major), tf32 x tf32 -> f32
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (p.major_bits << 15) | ((uint32_t)(N >> 3) << 17) |
                             ((uint32_t)(TC_BM >> 4) << 24);
      const uint32_t d_main = tmem_base, d_cross = tmem_base + TC_MAX_N;
      int it = 0, tcount = 0;
      for (int w = blockIdx.x; w < p.items; w += gridDim.x, ++tcount) {
        const int ch = w / p.m_tiles;
        const int kb0 = ch * TN_CHUNK_KB, kb1 = min(kb_total, kb0 + TN_CHUNK_KB);
        mbar_wait(tempty_bar, (tcount & 1) ^ 1);
        tc_fence_after();
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int st = it % TC_STAGES;
          const uint32_t ph = (it / TC_STAGES) & 1;
          mbar_wait(full_bar(st), ph);
          mbar_wait(xform_bar(st), ph);
          tc_fence_after();
          const uint32_t sbase = base + st * STAGE_BYTES;
          const uint64_t a_hi = make_desc_mn(sbase, p.lbo, p.sbo, p.layout), a_lo = make_desc_mn(sbase + A_TILE_BYTES, p.lbo, p.sbo, p.layout);
          const uint64_t b_hi = make_desc_mn(sbase + 2 * A_TILE_BYTES, p.lbo, p.sbo, p.layout),
                         b_lo = make_desc_mn(sbase + 2 * A_TILE_BYTES + B_TILE_BYTES, p.lbo, p.sbo, p.layout);
#pragma unroll
          for (int kk = 0; kk < TC_BK / 8; ++kk) {
            const uint64_t adv = (uint64_t)((kk * p.kadv) >> 4);  // next group of 8 k-rows
            const uint32_t first = (kb > kb0 || kk > 0) ? 1u : 0u;
            umma_tf32(d_cross, a_lo + adv, b_hi + adv, idesc, first);
            umma_tf32(d_cross, a_hi + adv, b_lo + adv, idesc, 1u);
            umma_tf32(d_main, a_hi + adv, b_hi + adv, idesc, first);
          }
          umma_commit(empty_bar(st));
        }
        umma_commit(tfull_bar);
      }
    }
  } else if (warp >= 8) {
    const int t = threadIdx.x - 256;
    const int b_vec = n_chunks32 * 256;  // 16-byte vectors in the B tile
    int it = 0;
    for (int w = blockIdx.x; w < p.items; w += gridDim.x) {
      const int ch = w / p.m_tiles;
      const int kb0 = ch * TN_CHUNK_KB, kb1 = min(kb_total, kb0 + TN_CHUNK_KB);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int st = it % TC_STAGES;
        const uint32_t ph = (it / TC_STAGES) & 1;
        mbar_wait(full_bar(st), ph);
        uint4* a = reinterpret_cast<uint4*>(base_ptr + st * STAGE_BYTES);
        uint4* alo = reinterpret_cast<uint4*>(base_ptr + st * STAGE_BYTES + A_TILE_BYTES);
        uint4* b = reinterpret_cast<uint4*>(base_ptr + st * STAGE_BYTES + 2 * A_TILE_BYTES);
        uint4* blo = reinterpret_cast<uint4*>(base_ptr + st * STAGE_BYTES + 2 * A_TILE_BYTES + B_TILE_BYTES);
        auto split = [](uint4* src, uint4* lo, int idx) {
          uint4 v = src[idx];
          uint4 h = make_uint4(v.x & 0xFFFFE000u, v.y & 0xFFFFE000u, v.z & 0xFFFFE000u, v.w & 0xFFFFE000u);
          uint4 l;
          l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x));
          l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y));
          l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z));
          l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w));
          src[idx] = h;
          lo[idx] = l;
        };
#pragma unroll
        for (int i = 0; i < A_TILE_BYTES / 16 / 128; ++i) split(a, alo, t + i * 128);
        for (int idx = t; idx < b_vec; idx += 128) split(b, blo, idx);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(xform_bar(st));
      }
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    float* stg = epi_stage + q * 32 * EPI_ROW;
    float* slab = p.partial + (size_t)blockIdx.x * ((size_t)p.m_tiles * TC_BM) * N;
    unsigned seen = 0;  // m-tiles this CTA has already written (first visit stores, later visits add)
    int tcount = 0;
    for (int w = blockIdx.x; w < p.items; w += gridDim.x, ++tcount) {
      const int mt = w % p.m_tiles;
      const bool add = (seen >> mt) & 1u;
      seen |= 1u << mt;
      mbar_wait(tfull_bar, tcount & 1);
      tc_fence_after();
      const int row0 = mt * TC_BM + q * 32;
      for (int c = 0; c < N; c += 32) {
        uint32_t r[32], x[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c;
#define DC_TMEM_LD32(R, ADDR)                                                                                            \
  asm volatile(                                                                                                         \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                         \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, " \
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                                                                 \
      : "=r"(R[0]), "=r"(R[1]), "=r"(R[2]), "=r"(R[3]), "=r"(R[4]), "=r"(R[5]), "=r"(R[6]), "=r"(R[7]), "=r"(R[8]),    \
        "=r"(R[9]), "=r"(R[10]), "=r"(R[11]), "=r"(R[12]), "=r"(R[13]), "=r"(R[14]), "=r"(R[15]), "=r"(R[16]),         \
        "=r"(R[17]), "=r"(R[18]), "=r"(R[19]), "=r"(R[20]), "=r"(R[21]), "=r"(R[22]), "=r"(R[23]), "=r"(R[24]),        \
        "=r"(R[25]), "=r"(R[26]), "=r"(R[27]), "=r"(R[28]), "=r"(R[29]), "=r"(R[30]), "=r"(R[31])                       \
      : "r"(ADDR))
        DC_TMEM_LD32(r, taddr);
        DC_TMEM_LD32(x, taddr + TC_MAX_N);
#undef DC_TMEM_LD32
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 v;
          v.x = __uint_as_float(r[j]) + __uint_as_float(x[j]);
          v.y = __uint_as_float(r[j + 1]) + __uint_as_float(x[j + 1]);
          v.z = __uint_as_float(r[j + 2]) + __uint_as_float(x[j + 2]);
          v.w = __uint_as_float(r[j + 3]) + __uint_as_float(x[j + 3]);
          *reinterpret_cast<float4*>(stg + lane * EPI_ROW + j) = v;
        }
        __syncwarp();
        const int col = c + (lane & 7) * 4;
#pragma unroll
        for (int it4 = 0; it4 < 8; ++it4) {
          const int rr = it4 * 4 + (lane >> 3);
          float4 v = *reinterpret_cast<const float4*>(stg + rr * EPI_ROW + (lane & 7) * 4);
          float4* dst = reinterpret_cast<float4*>(slab + (size_t)(row0 + rr) * N + col);
          if (add) { const float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
          *dst = v;
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(tempty_bar);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// C[m, n] = (accumulate ? C : 0) + sum over CTA slabs in CTA order
__global__ void tn_reduce_kernel(const float* __restrict__ partial, int slabs, int m_pad, int M, int N, float* __restrict__ C,
                                 long long ldc, int accumulate) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)M * N) return;
  const int m = (int)(idx / N), n = (int)(idx % N);
  float s = 0.f;
  for (int z = 0; z < slabs; ++z) s += partial[((size_t)z * m_pad + m) * N + n];
  if (accumulate) s += C[m * ldc + n];
  C[m * ldc + n] = s;
}

// B_s (either layout) -> packed K-major [N, Ktot] hi / lo
struct PackParams {
  const float* B[4];
  long long ldb[4];
  int K[4], koff[4];
  int nseg, N, Ktot, transB;
};
__global__ void pack_b_kernel(const PackParams p, float* __restrict__ hi, float* __restrict__ lo) {
  const int seg = blockIdx.z;
  const int k = blockIdx.x * 32 + threadIdx.x, n = blockIdx.y * 8 + threadIdx.y;
  if (seg >= p.nseg || k >= p.K[seg] || n >= p.N) return;
  const float b = p.transB ? p.B[seg][(long long)n * p.ldb[seg] + k] : p.B[seg][(long long)k * p.ldb[seg] + n];
  const float h = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u);
  const long long o = (long long)n * p.Ktot + p.koff[seg] + k;
  hi[o] = h;
  lo[o] = b - h;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(sym);
  }
  return fn;
}

int make_map(CUtensorMap* map, const float* ptr, uint64_t inner, uint64_t rows, uint64_t ld_elems, uint32_t box_rows,
             CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeFn enc = get_encode();
  DC_REQUIRE(enc, DC_ECUDA, "gemm_tc: cuTensorMapEncodeTiled not available");
  cuuint64_t gdim[2] = {inner, rows};
  cuuint64_t gstr[1] = {ld_elems * 4};
  cuuint32_t box[2] = {TC_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DC_REQUIRE(r == CUDA_SUCCESS, DC_ECUDA, "gemm_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return DC_OK;
}
}  // namespace

static bool tn_supported(const dc_gemm_seg* segs, int nseg, int transB, int64_t M, int64_t N) {
  if (nseg != 1 || transB) return false;
  if (N % 32 != 0 || N < 32 || N > TC_MAX_N) return false;
  if (M < 1 || M > 32 * TC_BM || segs[0].K < 1 || segs[0].K >= (1ll << 31)) return false;
  if (segs[0].lda % 4 != 0 || segs[0].ldb % 4 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(segs[0].A) & 15) || (reinterpret_cast<uintptr_t>(segs[0].B) & 15)) return false;
  return true;
}

bool gemm_tc_supported(const dc_gemm_seg* segs, int nseg, int transA, int transB, int64_t M, int64_t N, const float* C,
                       int64_t ldc, int accumulate, bool for_auto) {
  (void)C; (void)ldc; (void)accumulate;
  // AUTO never picks the tensor path for weight gradients: the contraction runs over all nodes with heavy
  // cancellation, where the tensor core's truncating TMEM accumulation (not the 3xTF32 split) costs ~30x in
  // accuracy vs fp32 FFMA with round-to-nearest (measured r01: 1.4e-4 vs 2e-6 on the end-to-end gradient).
  // DC_GEMM_TF32X3 / DC_GEMM_PREFER_TC still select it explicitly.
  if (transA) return !for_auto && tn_supported(segs, nseg, transB, M, N);
  if (nseg < 1 || nseg > 4) return false;
  // N <= 256: one chunk, MMA N = N (multiple of 16).  N > 256: chunks of 256 columns, the last one zero-filled by TMA.
  if (N < 16 || (N <= TC_MAX_N && N % 16 != 0) || N >= (1ll << 24)) return false;
  if (M < 1 || M >= (1ll << 31)) return false;
  int64_t ktot = 0;
  for (int s = 0; s < nseg; ++s) {
    // a single segment may have any K % 4 == 0 (TMA zero-fills the last k-block of A and of the packed B)
    if (segs[s].K <= 0 || (nseg > 1 ? segs[s].K % TC_BK != 0 : segs[s].K % 4 != 0)) return false;
    if (segs[s].lda % 4 != 0 || (reinterpret_cast<uintptr_t>(segs[s].A) & 15)) return false;
    ktot += segs[s].K;
  }
  // AUTO: small problems are not worth the packing pass / persistent launch
  return !for_auto || (double)M * (double)N * (double)ktot >= 1.0e8;
}

static int tn_grid(int64_t M, int64_t K) {
  const int64_t items = cdiv(M, TC_BM) * cdiv(cdiv(K, TC_BK), TN_CHUNK_KB);
  return (int)(items < kSMs ? items : kSMs);
}

size_t gemm_tc_workspace_bytes(int64_t M, int64_t N, int64_t Ktot, int transA, int transB) {
  (void)transB;
  if (transA) {
    if (N > TC_MAX_N) return 0;
    if (M > 32 * TC_BM) return 0;
    return align_up((size_t)tn_grid(M, Ktot) * cdiv(M, TC_BM) * TC_BM * N * sizeof(float), 256) + 256;
  }
  return align_up((size_t)2 * N * Ktot * sizeof(float), 256) + 256;
}

static int gemm_tc_tn(const dc_gemm_seg* segs, int64_t M, int64_t N, float* C, int64_t ldc, const float* bias, int relu,
                      int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  DC_REQUIRE(!bias && !relu, DC_ENOSUP, "gemm_tc: bias / relu are not supported on the transposed (weight-gradient) path");
  const int64_t K = segs[0].K;
  const size_t need = gemm_tc_workspace_bytes(M, N, K, 1, 0);
  DC_REQUIRE(workspace && workspace_bytes >= need, DC_EWORKSPACE, "gemm_tc: workspace %zu < %zu", workspace_bytes, need);
  float* partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  TnParams p{};
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.m_tiles = (int)cdiv(M, TC_BM);
  p.chunks = (int)cdiv(cdiv(K, TC_BK), TN_CHUNK_KB);
  p.items = p.m_tiles * p.chunks;
  p.partial = partial;
  // MN-major SWIZZLE_128B canonical layout: LBO = next 32-element chunk along M/N (one TMA box, 4096 B),
  // SBO = next group of 8 along K (1024 B); one K=8 MMA step advances by SBO.
  // MN-major tf32 operands: SWIZZLE_128B_BASE32B is the only layout the tensor core accepts (32-byte swizzle
  // atoms; TMA side: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).  LBO = next 32-element chunk along M/N (one TMA box,
  // 4096 B); SBO = next group of 4 k-rows (512 B); one K=8 MMA step advances the start address by 1024 B.
  // (Established on hardware in round 1: plain SWIZZLE_128B with the transpose bits set silently yields zeros.)
  p.lbo = 4096; p.sbo = 512; p.kadv = 1024; p.major_bits = 3; p.layout = 1;
  const CUtensorMapSwizzle tma_swz = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  const int grid = tn_grid(M, K);
  const size_t slab = (size_t)p.m_tiles * TC_BM * N * sizeof(float);
  DC_CUDA(cudaMemsetAsync(partial, 0, slab * grid, st));
  CUtensorMap mA, mB;
  if (int rc = make_map(&mA, segs[0].A, (uint64_t)M, (uint64_t)K, (uint64_t)segs[0].lda, 32, tma_swz)) return rc;
  if (int rc = make_map(&mB, segs[0].B, (uint64_t)N, (uint64_t)K, (uint64_t)segs[0].ldb, 32, tma_swz)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    DC_CUDA(cudaFuncSetAttribute(gemm_tc_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  gemm_tc_tn_kernel<<<grid, TC_THREADS, SMEM_BYTES, st>>>(mA, mB, p);
  DC_LAUNCH_CHECK();
  tn_reduce_kernel<<<(unsigned)cdiv(M * N, 256), 256, 0, st>>>(partial, grid, p.m_tiles * TC_BM, (int)M, (int)N, C, ldc, accumulate);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

int gemm_tc(const dc_gemm_seg* segs, int nseg, int transA, int transB, int64_t M, int64_t N, float* C, int64_t ldc,
            const float* bias, int relu, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (transA) {
    DC_REQUIRE(tn_supported(segs, nseg, transB, M, N), DC_ENOSUP, "gemm_tc: transposed shape unsupported");
    return gemm_tc_tn(segs, M, N, C, ldc, bias, relu, accumulate, workspace, workspace_bytes, st);
  }
  int64_t ktot = 0;
  for (int s = 0; s < nseg; ++s) ktot += segs[s].K;
  const size_t need = gemm_tc_workspace_bytes(M, N, ktot, transA, transB);
  DC_REQUIRE(workspace && workspace_bytes >= need, DC_EWORKSPACE, "gemm_tc: workspace %zu < %zu", workspace_bytes, need);
  float* bhi = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  float* blo = bhi + N * ktot;

  PackParams pp{};
  pp.nseg = nseg; pp.N = (int)N; pp.Ktot = (int)ktot; pp.transB = transB;
  int maxk = 0, koff = 0;
  for (int s = 0; s < nseg; ++s) {
    pp.B[s] = segs[s].B; pp.ldb[s] = segs[s].ldb; pp.K[s] = (int)segs[s].K; pp.koff[s] = koff;
    koff += (int)segs[s].K;
    maxk = maxk > (int)segs[s].K ? maxk : (int)segs[s].K;
  }
  dim3 pgrid((unsigned)cdiv(maxk, 32), (unsigned)cdiv(N, 8), (unsigned)nseg);
  pack_b_kernel<<<pgrid, dim3(32, 8), 0, st>>>(pp, bhi, blo);
  DC_LAUNCH_CHECK();

  CUtensorMap maps[4], mbhi, mblo;
  TcParams p{};
  p.nseg = nseg;
  for (int s = 0; s < 4; ++s) {
    const int ss = s < nseg ? s : 0;
    if (int rc = make_map(&maps[s], segs[ss].A, (uint64_t)segs[ss].K, (uint64_t)M, (uint64_t)segs[ss].lda, TC_BM)) return rc;
    p.kb_seg[s] = s < nseg ? (int)cdiv(segs[s].K, TC_BK) : 0;
  }
  p.m_tiles = (int)cdiv(M, TC_BM);
  // 256-column chunks keep the B tile traffic per flop lowest; fall back to 128 when 256 would leave SMs without a tile
  p.n_chunk = (N > 128 && (int64_t)p.m_tiles * cdiv(N, TC_MAX_N) < kSMs) ? 128 : TC_MAX_N;
  p.n_chunks = (int)cdiv(N, p.n_chunk);
  p.n_mma = N <= p.n_chunk ? (int)N : p.n_chunk;
  if (int rc = make_map(&mbhi, bhi, (uint64_t)ktot, (uint64_t)N, (uint64_t)ktot, (uint32_t)p.n_mma)) return rc;
  if (int rc = make_map(&mblo, blo, (uint64_t)ktot, (uint64_t)N, (uint64_t)ktot, (uint32_t)p.n_mma)) return rc;
  p.kb_total = (int)cdiv(ktot, TC_BK);
  p.M = (int)M; p.N = (int)N; p.C = C; p.ldc = ldc; p.bias = bias; p.relu = relu; p.accumulate = accumulate;
  p.m_tiles = (int)cdiv(M, TC_BM);
  p.drain_kb = TC_DRAIN_KB;
  if (const char* e = getenv("DCB200_DRAIN_KB")) p.drain_kb = atoi(e) > 0 ? atoi(e) : TC_DRAIN_KB;

  static bool attr_set = false;
  if (!attr_set) {
    DC_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  const int64_t tiles = (int64_t)p.m_tiles * p.n_chunks;
  const int grid = (int)(tiles < kSMs ? tiles : kSMs);
  gemm_tc_kernel<<<grid, TC_THREADS, SMEM_BYTES, st>>>(maps[0], maps[1], maps[2], maps[3], mbhi, mblo, p);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

}  // namespace dcb
