// K2 (tensor-core variant) — tcgen05 kind::tf32 GEMM with 3-term error-compensated split.
// Placeholder until the kernel lands: reports "unsupported" so dc_gemm uses the fp32 SIMT path.
#include "common.cuh"

namespace dcb {
bool gemm_tc_supported(const dc_gemm_seg*, int, int, int, int64_t, int64_t, const float*, int64_t, int) { return false; }
size_t gemm_tc_workspace_bytes(int64_t, int64_t, int64_t, int, int) { return 0; }
int gemm_tc(const dc_gemm_seg*, int, int, int, int64_t, int64_t, float*, int64_t, const float*, int, int, void*, size_t,
            cudaStream_t) {
  set_error("gemm_tc: not built");
  return DC_ENOSUP;
}
}  // namespace dcb
