// xo_common.cuh -- launch helpers shared by the kernels' C-ABI wrappers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#ifndef XO_NO_NVTX
#include <nvtx3/nvToolsExt.h>
#endif

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

// kernels launched by this library in this process (xo_diag_launch_count; bench.py's `gpu_launches`)
extern std::atomic<long long> g_launch_count;
inline void note_launch() { g_launch_count.fetch_add(1, std::memory_order_relaxed); }

// NVTX range around an entry point's launches (tracing); header-only NVTX v3, a no-op
// unless a profiler is attached.  -DXO_NO_NVTX compiles it out.
struct NvtxScope {
#ifndef XO_NO_NVTX
  explicit NvtxScope(const char *name) { nvtxRangePushA(name); }
  ~NvtxScope() { nvtxRangePop(); }
#else
  explicit NvtxScope(const char *) {}
#endif
};
#define XO_NVTX_SCOPE(name) xo::NvtxScope _xo_nvtx_scope(name)

inline int launch_status() {
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? 0 : -(int)err;
}

inline int cuda_status(cudaError_t err) { return err == cudaSuccess ? 0 : -(int)err; }

// Programmatic dependent launch (sm_90+): a kernel launched with `pdl = true` may begin while the
// preceding kernel of the stream is still running; it must execute pdl_wait() before touching
// anything that kernel writes.  The preceding kernel lets it start early with
// pdl_launch_dependents().  Both are no-ops for a kernel launched the ordinary way.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#endif

#ifdef __CUDACC__
// every ordinary launch of the library goes through here (counted)
template <class... KArgs, class... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  note_launch();
  kernel<<<grid, block, smem, st>>>(KArgs(args)...);
}
#endif

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  note_launch();
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace xo
