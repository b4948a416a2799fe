// L1 gather-throughput lab (B200): how many bytes per clock does one SM deliver for L1-RESIDENT row-slice gathers,
// as a function of the load width and of the number of distinct 128-byte lines one warp instruction touches?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/l1_lab scripts/l1_lab.cu && ./scripts/l1_lab
// Each CTA (one per SM) owns a private 64 KB table of 512 rows x 128 B (L1 resident after the first pass) and a warp issues
// UNROLL independent loads per iteration at pseudo-random rows:
//   v4 x 4 rows : LDG.128, 8 lanes per row, 4 rows per instruction   (the K1 mapping: 8 lanes x float4 per receiver)
//   v4 x 1 row  : LDG.128, all 4 quarter-warps in the same 512-byte span (coalesced)
//   v2 x 2 rows : LDG.64, 16 lanes per row
//   v2 x 1 row  : LDG.64, 256 contiguous bytes
//   v1 x 1 row  : LDG.32, 32 lanes on one 128-byte line
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int ROWS = 512;          // x 128 B = 64 KB per CTA
constexpr int UNROLL = 8;

template <int MODE>
__global__ void __launch_bounds__(1024, 1) gather_kernel(const float* __restrict__ table, float* __restrict__ sink, int iters) {
  const float* tab = table + (size_t)blockIdx.x * ROWS * 32;
  const int lane = threadIdx.x & 31;
  unsigned s = (threadIdx.x >> 5) * 2654435761u + 12345u + blockIdx.x;   // warp-uniform random stream
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    float r[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      s = s * 1664525u + 1013904223u;
      const unsigned w = s;
      if (MODE == 0) {          // LDG.128, 4 rows per instruction
        const unsigned row = (w >> 8) + (lane >> 3) * 131u;
        const float4 v = *reinterpret_cast<const float4*>(tab + (row % ROWS) * 32 + (lane & 7) * 4);
        r[u] = v.x + v.y + v.z + v.w;
      } else if (MODE == 1) {   // LDG.128, 512 contiguous bytes
        const unsigned row = (w >> 8) % (ROWS / 4);
        const float4 v = *reinterpret_cast<const float4*>(tab + row * 128 + lane * 4);
        r[u] = v.x + v.y + v.z + v.w;
      } else if (MODE == 2) {   // LDG.64, 2 rows per instruction
        const unsigned row = (w >> 8) + (lane >> 4) * 131u;
        const float2 v = *reinterpret_cast<const float2*>(tab + (row % ROWS) * 32 + (lane & 15) * 2);
        r[u] = v.x + v.y;
      } else if (MODE == 3) {   // LDG.64, 256 contiguous bytes
        const unsigned row = (w >> 8) % (ROWS / 2);
        const float2 v = *reinterpret_cast<const float2*>(tab + row * 64 + lane * 2);
        r[u] = v.x + v.y;
      } else {                  // LDG.32, one line
        const unsigned row = (w >> 8) % ROWS;
        r[u] = tab[row * 32 + lane];
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += r[u];
  }
  if (acc == 123.456f) sink[threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, int bytes_per_instr, const float* table, float* sink, int sms) {
  const int iters = 2000;
  gather_kernel<MODE><<<sms, 1024>>>(table, sink, 50);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  gather_kernel<MODE><<<sms, 1024>>>(table, sink, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double instr_per_sm = 32.0 * iters * UNROLL;                 // warp instructions per SM
  const double bytes_per_sm = instr_per_sm * bytes_per_instr;
  printf("%-14s %8.3f ms   %7.1f GB/s per SM   %6.1f B/clk per SM at %d MHz (nominal)   %5.2f clk per warp instruction\n", name, ms,
         bytes_per_sm / (ms * 1e-3) / 1e9, bytes_per_sm / (ms * 1e-3) / (clk_khz * 1e3), clk_khz / 1000,
         (ms * 1e-3) * (clk_khz * 1e3) / instr_per_sm);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float *table, *sink;
  cudaMalloc(&table, (size_t)sms * ROWS * 128);
  cudaMalloc(&sink, 4096);
  cudaMemset(table, 0, (size_t)sms * ROWS * 128);
  cudaFuncSetAttribute(gather_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
  run<0>("v4 x 4 rows", 512, table, sink, sms);
  run<1>("v4 x 1 span", 512, table, sink, sms);
  run<2>("v2 x 2 rows", 256, table, sink, sms);
  run<3>("v2 x 1 span", 256, table, sink, sms);
  run<4>("v1 x 1 row", 128, table, sink, sms);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
