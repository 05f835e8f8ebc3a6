// xo_common.cuh -- launch helpers shared by the kernels' C-ABI wrappers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/xo_b200.h"

namespace xo {

// SM count of the current device (148 on B200); cached per device ordinal.
inline int sm_count() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < 64 && cache[dev] > 0) return cache[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  if (dev >= 0 && dev < 64) cache[dev] = n;
  return n;
}

// grid for a grid-stride kernel: a whole multiple of the SM count (<= ctas_per_sm
// resident CTAs each), never more CTAs than there is work.
inline int64_t clamp_grid(int64_t work_ctas, int sms, int ctas_per_sm) {
  int64_t full = (int64_t)sms * ctas_per_sm;
  if (work_ctas >= full) return full;
  return work_ctas < 1 ? 1 : work_ctas;
}

inline int launch_status() {
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? 0 : -(int)err;
}

inline int cuda_status(cudaError_t err) { return err == cudaSuccess ? 0 : -(int)err; }

}  // namespace xo
