// Host-side launcher of one class kernel (explicitly instantiated in gen/eri_inst_*.cu).
#pragma once
#include "eri_config.cuh"

namespace lb200 {

template <int LA, int LB, int LC, int LD, int MODE>
cudaError_t launch_class(const EriParams& p, const RowInfo* rows, int num_sms,
                         cudaStream_t stream) {
  using C = Cfg<LA, LB, LC, LD, MODE>;
  if constexpr (!C::FITS) {
    return cudaErrorInvalidConfiguration;
  } else {
    auto kern = eri_class_kernel<LA, LB, LC, LD, C::T, MODE>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           C::SMEM_BYTES);
    if (err != cudaSuccess) return err;
    long long grid = (long long)num_sms * C::CTAS_PER_SM;
    if (!p.ntasks_dev) {
      const long long need = ((long long)p.ntasks + C::TEAMS_PER_CTA - 1) / C::TEAMS_PER_CTA;
      if (need < grid) grid = need;
    }
    if (grid < 1) return cudaSuccess;
    kern<<<(unsigned)grid, C::THREADS, C::SMEM_BYTES, stream>>>(p, rows);
    return cudaGetLastError();
  }
}

// one entry point per class; mode selects the instantiation
template <int LA, int LB, int LC, int LD>
cudaError_t launch_class_any(const EriParams& p, const RowInfo* rows, int mode, int num_sms,
                             cudaStream_t stream) {
  if (mode == kModeFock) return launch_class<LA, LB, LC, LD, kModeFock>(p, rows, num_sms, stream);
  return launch_class<LA, LB, LC, LD, kModeStoreCart>(p, rows, num_sms, stream);
}

}  // namespace lb200
