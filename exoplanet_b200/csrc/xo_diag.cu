// xo_diag.cu -- measurement helpers exported next to the product entry points.
//
//  xo_diag_dfma_peak: the fp64 ceiling the fused kernels are quoted against (SURVEY.md 8d ceiling (B),
//    BASELINE.md section 2 "must be measured by the builder"): a register-resident DFMA loop on every SM,
//    timed with CUDA events on the stream it is launched on.  8 independent accumulator chains per
//    thread, 4 x 256-thread CTAs per SM, so the fp64 pipe's issue slot is never left waiting on a dependency.
//  xo_diag_launch_count: kernels this library has launched in this process (bench.py's `gpu_launches`).
#include <cuda_runtime.h>

#include <atomic>

#include "xo_common.cuh"

namespace xo {
std::atomic<long long> g_launch_count{0};
}

namespace {

constexpr int kDfmaIlp = 8;
constexpr int kDfmaInner = 64;  // DFMAs per chain per outer iteration (unrolled)

__global__ void __launch_bounds__(256) dfma_peak_kernel(double *__restrict__ sink, int outer, double a, double b) {
  double x[kDfmaIlp];
#pragma unroll
  for (int j = 0; j < kDfmaIlp; ++j) x[j] = (double)(threadIdx.x + j) * 1e-3;
  for (int it = 0; it < outer; ++it) {
#pragma unroll
    for (int u = 0; u < kDfmaInner; ++u) {
#pragma unroll
      for (int j = 0; j < kDfmaIlp; ++j) x[j] = fma(x[j], a, b);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < kDfmaIlp; ++j) s += x[j];
  if (s == 123.456) sink[0] = s;  // never true for the constants used; keeps the chains alive
}

}  // namespace

extern "C" {

long long xo_diag_launch_count(void) { return xo::g_launch_count.load(std::memory_order_relaxed); }

// status word at the head of a workspace (celerite entry points): lanes with a non-finite result, expired
// inter-CTA waits.  Copies 16 bytes device -> host on `stream` and waits for them.
int xo_read_status(const void *workspace, unsigned long long *bad_lanes, unsigned long long *wait_failures, void *stream) {
  if (!workspace) return XO_EINVAL;
  unsigned long long h[2] = {0, 0};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemcpyAsync(h, workspace, sizeof(h), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return -(int)e;
  if (bad_lanes) *bad_lanes = h[0];
  if (wait_failures) *wait_failures = h[1];
  return 0;
}

int xo_diag_dfma_peak(int reps, double *dfma_per_s, double *best_ms, void *stream) {
  if (reps < 1 || !dfma_per_s) return XO_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int sms = xo::sm_count();
  const int grid = sms * 4, block = 256, outer = 4096;
  double *sink = nullptr;
  cudaError_t err = cudaMalloc(&sink, sizeof(double));
  if (err != cudaSuccess) return -(int)err;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int w = 0; w < 2; ++w) dfma_peak_kernel<<<grid, block, 0, st>>>(sink, outer, 0.999999, 1e-7);
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0, st);
    dfma_peak_kernel<<<grid, block, 0, st>>>(sink, outer, 0.999999, 1e-7);
    cudaEventRecord(e1, st);
    err = cudaEventSynchronize(e1);
    if (err != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  if (err != cudaSuccess) return -(int)err;
  const double n = (double)grid * block * (double)outer * kDfmaInner * kDfmaIlp;  // thread-level DFMAs per launch
  *dfma_per_s = n / (best * 1e-3);
  if (best_ms) *best_ms = best;
  return xo::launch_status();
}

}  // extern "C"
