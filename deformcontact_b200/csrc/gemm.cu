// dc_gemm — C-ABI entry for the layer GEMMs; dispatches between the exact-fp32 SIMT kernel
// (gemm_simt.cu) and the tcgen05 3xTF32 tensor-core kernel (gemm_tc.cu).
#include "common.cuh"

namespace dcb {
size_t gemm_simt_workspace_bytes(int64_t M, int64_t N, int64_t Ktot);
int gemm_simt(const dc_gemm_seg* segs, int nseg, int transA, int transB, int64_t M, int64_t N, float* C, int64_t ldc,
              const float* bias, int relu, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t st);
bool gemm_tc_supported(const dc_gemm_seg* segs, int nseg, int transA, int transB, int64_t M, int64_t N, const float* C,
                       int64_t ldc, int accumulate, bool for_auto);
size_t gemm_tc_workspace_bytes(int64_t M, int64_t N, int64_t Ktot, int transA, int transB);
int gemm_tc(const dc_gemm_seg* segs, int nseg, int transA, int transB, int64_t M, int64_t N, float* C, int64_t ldc,
            const float* bias, int relu, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t st);
}  // namespace dcb

extern "C" size_t dc_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K_total, int transA, int transB) {
  if (M < 0 || N < 0 || K_total < 0) return 0;
  size_t a = dcb::gemm_simt_workspace_bytes(M, N, K_total);
  size_t b = dcb::gemm_tc_workspace_bytes(M, N, K_total, transA, transB);
  return a > b ? a : b;
}

extern "C" int dc_gemm(const dc_gemm_seg* segs, int nseg, int transA, int transB, int64_t M, int64_t N, float* C,
                       int64_t ldc, const float* bias, int relu, int accumulate, int precision, void* workspace,
                       size_t workspace_bytes, dc_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DC_REQUIRE(segs && nseg >= 1 && nseg <= 4, DC_EINVAL, "gemm: nseg must be 1..4");
  DC_REQUIRE(M >= 0 && N >= 0, DC_EINVAL, "gemm: negative size");
  if (M == 0 || N == 0) return DC_OK;
  DC_REQUIRE(C && ldc >= N, DC_EINVAL, "gemm: bad C / ldc");
  DC_REQUIRE(precision >= DC_GEMM_AUTO && precision <= DC_GEMM_PREFER_TC, DC_EINVAL, "gemm: unknown precision %d", precision);
  if (precision == DC_GEMM_TF32X3) {
    bool tc_ok = dcb::gemm_tc_supported(segs, nseg, transA, transB, M, N, C, ldc, accumulate, false);
    DC_REQUIRE(tc_ok, DC_ENOSUP, "gemm: shape/layout not supported by the tcgen05 path");
    return dcb::gemm_tc(segs, nseg, transA, transB, M, N, C, ldc, bias, relu, accumulate, workspace, workspace_bytes, st);
  }
  if ((precision == DC_GEMM_AUTO || precision == DC_GEMM_PREFER_TC) &&
      dcb::gemm_tc_supported(segs, nseg, transA, transB, M, N, C, ldc, accumulate, precision == DC_GEMM_AUTO))
    return dcb::gemm_tc(segs, nseg, transA, transB, M, N, C, ldc, bias, relu, accumulate, workspace, workspace_bytes, st);
  return dcb::gemm_simt(segs, nseg, transA, transB, M, N, C, ldc, bias, relu, accumulate, workspace, workspace_bytes, st);
}
