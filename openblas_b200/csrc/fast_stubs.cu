/* fast_stubs.cu -- placeholders for roofline kernel families not built yet: they decline every
 * problem so the dispatcher falls through to the generic kernel.  Each entry disappears from
 * this file when its real kernel lands. */
#include "gemm_common.cuh"
namespace b200 {
#ifndef HAVE_DGEMM_DMMA
cudaError_t launch_dgemm_dmma(const DeviceGemm &, cudaStream_t) { return cudaErrorNotSupported; }
#endif
#ifndef HAVE_ZGEMM_DMMA
cudaError_t launch_zgemm_dmma(const DeviceGemm &, cudaStream_t) { return cudaErrorNotSupported; }
#endif
#ifndef HAVE_SGEMM_FFMA
cudaError_t launch_sgemm_ffma(const DeviceGemm &, cudaStream_t) { return cudaErrorNotSupported; }
#endif
#ifndef HAVE_CGEMM_FFMA
cudaError_t launch_cgemm_ffma(const DeviceGemm &, cudaStream_t) { return cudaErrorNotSupported; }
#endif
#ifndef HAVE_SBGEMM_TCGEN05
cudaError_t launch_sbgemm_tcgen05(const DeviceGemm &, cudaStream_t) { return cudaErrorNotSupported; }
#endif
}  // namespace b200
