// xo_ops.cu -- the three reference ops as standalone sm_100a kernels + their C ABI.
//   kepler                (replaces /root/reference/src/exoplanet/orbits/keplerian.py:333)
//   quad_solution_vector  (replaces /root/reference/src/exoplanet/light_curves/limb_dark.py:24)
//   contact_points        (replaces /root/reference/src/exoplanet/orbits/keplerian.py:744-753)
// All are elementwise fp64: no tensor cores, no shared memory needed; the levers are
// coalesced 16-byte loads/stores, two independent elements per thread for ILP, streaming
// cache hints (no reuse), and a grid sized as a multiple of the SM count.
#include <cuda_runtime.h>

#include <cstdlib>

#include "xo_common.cuh"
#include "xo_contact.cuh"
#include "xo_math.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ double2 ld2(const double *p) { return __ldcs(reinterpret_cast<const double2 *>(p)); }
__device__ __forceinline__ void st2(double *p, double a, double b) { __stcs(reinterpret_cast<double2 *>(p), make_double2(a, b)); }

// ---- kepler ---------------------------------------------------------------
template <int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks) kepler_fwd_kernel(const double *__restrict__ M,
                                                                          const double *__restrict__ e,
                                                                          double *__restrict__ sinf,
                                                                          double *__restrict__ cosf, int64_t n,
                                                                          int64_t e_inner) {
  const int64_t npair = n >> 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // software pipeline: the loads of iteration k+1 are in flight while iteration k computes
  double2 m = make_double2(0.0, 0.0), ee = make_double2(0.0, 0.0);
  if (i < npair) {
    m = ld2(M + 2 * i);
    ee = (e_inner == 1) ? ld2(e + 2 * i) : make_double2(__ldg(e + (2 * i) / e_inner), __ldg(e + (2 * i + 1) / e_inner));
  }
  for (; i < npair; i += stride) {
    const double2 mc = m, ec = ee;
    const int64_t nx = i + stride;
    if (nx < npair) {
      m = ld2(M + 2 * nx);
      ee = (e_inner == 1) ? ld2(e + 2 * nx) : make_double2(__ldg(e + (2 * nx) / e_inner), __ldg(e + (2 * nx + 1) / e_inner));
    }
    double s0, c0, s1, c1;
    xo::kepler_solve(mc.x, ec.x, s0, c0);
    xo::kepler_solve(mc.y, ec.y, s1, c1);
    st2(sinf + 2 * i, s0, s1);
    st2(cosf + 2 * i, c0, c1);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    int64_t i = n - 1;
    double s, c;
    xo::kepler_solve(M[i], e[i / e_inner], s, c);
    sinf[i] = s;
    cosf[i] = c;
  }
}

// unaligned fallback (pointer not 16-byte aligned): one element per thread
__global__ void __launch_bounds__(kThreads) kepler_fwd_kernel_scalar(const double *__restrict__ M,
                                                                     const double *__restrict__ e,
                                                                     double *__restrict__ sinf,
                                                                     double *__restrict__ cosf, int64_t n,
                                                                     int64_t e_inner) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double s, c;
    xo::kepler_solve(M[i], e[i / e_inner], s, c);
    sinf[i] = s;
    cosf[i] = c;
  }
}

// two elements per thread, 16-byte streaming loads/stores when every pointer allows it
template <bool kVec>
__global__ void __launch_bounds__(kThreads) kepler_vjp_kernel(const double *__restrict__ sinf,
                                                              const double *__restrict__ cosf,
                                                              const double *__restrict__ e,
                                                              const double *__restrict__ bs,
                                                              const double *__restrict__ bc, double *__restrict__ bM,
                                                              double *__restrict__ be, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  auto one = [](double sf, double cf, double ee, double gs, double gc, double &oM, double &oe) {
    double dfdM, dfde;
    xo::kepler_jac(sf, cf, ee, dfdM, dfde);
    const double bf = gs * cf - gc * sf;
    oM = bf * dfdM;
    oe = bf * dfde;
  };
  if (kVec) {
    const int64_t npair = n >> 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npair; i += stride) {
      const double2 sf = ld2(sinf + 2 * i), cf = ld2(cosf + 2 * i), ee = ld2(e + 2 * i), gs = ld2(bs + 2 * i), gc = ld2(bc + 2 * i);
      double m0, e0, m1, e1;
      one(sf.x, cf.x, ee.x, gs.x, gc.x, m0, e0);
      one(sf.y, cf.y, ee.y, gs.y, gc.y, m1, e1);
      st2(bM + 2 * i, m0, m1);
      st2(be + 2 * i, e0, e1);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
      const int64_t i = n - 1;
      one(sinf[i], cosf[i], e[i], bs[i], bc[i], bM[i], be[i]);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      double m0, e0;
      one(__ldcs(sinf + i), __ldcs(cosf + i), __ldcs(e + i), __ldcs(bs + i), __ldcs(bc + i), m0, e0);
      __stcs(bM + i, m0);
      __stcs(be + i, e0);
    }
  }
}

// ---- quad_solution_vector --------------------------------------------------
// One point per thread; the [n][3] outputs are staged through shared memory so each
// warp writes 768 contiguous bytes with fully coalesced 8-byte stores.
template <bool kJac>
__global__ void __launch_bounds__(kThreads) qsv_fwd_kernel(const double *__restrict__ b, const double *__restrict__ r,
                                                           double *__restrict__ s, double *__restrict__ dsdb,
                                                           double *__restrict__ dsdr, int64_t n) {
  __shared__ double stage[(kJac ? 3 : 1) * 3 * kThreads];
  const int64_t tiles = (n + kThreads - 1) / kThreads;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t base = tile * kThreads;
    const int64_t i = base + threadIdx.x;
    double S[3] = {0, 0, 0}, B[3] = {0, 0, 0}, R[3] = {0, 0, 0};
    if (i < n) xo::qsv<kJac>(__ldcs(b + i), __ldcs(r + i), S, B, R);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      stage[3 * threadIdx.x + k] = S[k];
      if (kJac) {
        stage[3 * kThreads + 3 * threadIdx.x + k] = B[k];
        stage[6 * kThreads + 3 * threadIdx.x + k] = R[k];
      }
    }
    __syncthreads();
    const int64_t cnt = min((int64_t)kThreads, n - base) * 3;
    for (int64_t j = threadIdx.x; j < cnt; j += kThreads) {
      __stcs(s + 3 * base + j, stage[j]);
      if (kJac) {
        __stcs(dsdb + 3 * base + j, stage[3 * kThreads + j]);
        __stcs(dsdr + 3 * base + j, stage[6 * kThreads + j]);
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads) qsv_vjp_kernel(const double *__restrict__ bs, const double *__restrict__ dsdb,
                                                           const double *__restrict__ dsdr, double *__restrict__ bb,
                                                           double *__restrict__ br, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double g0 = bs[3 * i], g1 = bs[3 * i + 1], g2 = bs[3 * i + 2];
    bb[i] = g0 * dsdb[3 * i] + g1 * dsdb[3 * i + 1] + g2 * dsdb[3 * i + 2];
    br[i] = g0 * dsdr[3 * i] + g1 * dsdr[3 * i + 1] + g2 * dsdr[3 * i + 2];
  }
}

// ---- contact_points (solver in xo_contact.cuh) -----------------------------
__global__ void contact_points_kernel(const double *a, const double *e, const double *cw, const double *sw,
                                      const double *ci, const double *si, const double *L, double *Ml, double *Mr,
                                      int32_t *flag, int64_t n) {
  (void)si;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double l, r;
  flag[i] = xo::cp_solve(a[i], e[i], cw[i], sw[i], ci[i], L[i], l, r);
  Ml[i] = l;
  Mr[i] = r;
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" {

int xo_kepler_fwd_bcast(const double *M, const double *e, int64_t e_inner, double *sinf, double *cosf, int64_t n,
                        void *stream) {
  XO_NVTX_SCOPE("xo_kepler_fwd_bcast");
  if (n < 0 || e_inner < 1) return XO_EINVAL;
  if (n == 0) return 0;
  if (!M || !e || !sinf || !cosf) return XO_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int sms = xo::sm_count();
  bool vec = aligned16(M) && aligned16(sinf) && aligned16(cosf) && (e_inner != 1 || aligned16(e));
  if (vec) {
    static const int mb = [] {
      const char *v = getenv("XO_KEPLER_MINBLOCKS");
      return v && *v ? atoi(v) : 2;  // 126 regs: loop-invariant constants stay in registers (measured fastest)
    }();
    int64_t work = (n / 2 + kThreads - 1) / kThreads;
    int grid = (int)xo::clamp_grid(work, sms, mb);
    if (mb == 2)
      xo::launch(kepler_fwd_kernel<2>, grid, kThreads, 0, st, M, e, sinf, cosf, n, e_inner);
    else if (mb == 3)
      xo::launch(kepler_fwd_kernel<3>, grid, kThreads, 0, st, M, e, sinf, cosf, n, e_inner);
    else
      xo::launch(kepler_fwd_kernel<4>, grid, kThreads, 0, st, M, e, sinf, cosf, n, e_inner);
  } else {
    int64_t work = (n + kThreads - 1) / kThreads;
    int grid = (int)xo::clamp_grid(work, sms, 8);
    xo::launch(kepler_fwd_kernel_scalar, grid, kThreads, 0, st, M, e, sinf, cosf, n, e_inner);
  }
  return xo::launch_status();
}

int xo_kepler_fwd(const double *M, const double *e, double *sinf, double *cosf, int64_t n, void *stream) {
  XO_NVTX_SCOPE("xo_kepler_fwd");
  return xo_kepler_fwd_bcast(M, e, 1, sinf, cosf, n, stream);
}

int xo_kepler_vjp(const double *sinf, const double *cosf, const double *e, const double *bsinf, const double *bcosf,
                  double *bM, double *be, int64_t n, void *stream) {
  XO_NVTX_SCOPE("xo_kepler_vjp");
  if (n < 0) return XO_EINVAL;
  if (n == 0) return 0;
  if (!sinf || !cosf || !e || !bsinf || !bcosf || !bM || !be) return XO_EINVAL;
  const bool vec = aligned16(sinf) && aligned16(cosf) && aligned16(e) && aligned16(bsinf) && aligned16(bcosf) && aligned16(bM) &&
                   aligned16(be);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vec) {
    int grid = (int)xo::clamp_grid((n / 2 + kThreads - 1) / kThreads, xo::sm_count(), 8);
    xo::launch(kepler_vjp_kernel<true>, grid, kThreads, 0, st, sinf, cosf, e, bsinf, bcosf, bM, be, n);
  } else {
    int grid = (int)xo::clamp_grid((n + kThreads - 1) / kThreads, xo::sm_count(), 8);
    xo::launch(kepler_vjp_kernel<false>, grid, kThreads, 0, st, sinf, cosf, e, bsinf, bcosf, bM, be, n);
  }
  return xo::launch_status();
}

int xo_qsv_fwd(const double *b, const double *r, double *s, double *dsdb, double *dsdr, int64_t n, void *stream) {
  XO_NVTX_SCOPE("xo_qsv_fwd");
  if (n < 0) return XO_EINVAL;
  if (n == 0) return 0;
  if (!b || !r || !s || ((dsdb == nullptr) != (dsdr == nullptr))) return XO_EINVAL;
  int grid = (int)xo::clamp_grid((n + kThreads - 1) / kThreads, xo::sm_count(), 4);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dsdb)
    xo::launch(qsv_fwd_kernel<true>, grid, kThreads, 0, st, b, r, s, dsdb, dsdr, n);
  else
    xo::launch(qsv_fwd_kernel<false>, grid, kThreads, 0, st, b, r, s, nullptr, nullptr, n);
  return xo::launch_status();
}

int xo_qsv_vjp(const double *bs, const double *dsdb, const double *dsdr, double *bb, double *br, int64_t n, void *stream) {
  XO_NVTX_SCOPE("xo_qsv_vjp");
  if (n < 0) return XO_EINVAL;
  if (n == 0) return 0;
  if (!bs || !dsdb || !dsdr || !bb || !br) return XO_EINVAL;
  int grid = (int)xo::clamp_grid((n + kThreads - 1) / kThreads, xo::sm_count(), 8);
  xo::launch(qsv_vjp_kernel, grid, kThreads, 0, static_cast<cudaStream_t>(stream), bs, dsdb, dsdr, bb, br, n);
  return xo::launch_status();
}

int xo_contact_points(const double *a, const double *e, const double *cosw, const double *sinw, const double *cosi,
                      const double *sini, const double *L, double *M_left, double *M_right, int32_t *flag, int64_t n,
                      void *stream) {
  XO_NVTX_SCOPE("xo_contact_points");
  if (n < 0) return XO_EINVAL;
  if (n == 0) return 0;
  if (!a || !e || !cosw || !sinw || !cosi || !sini || !L || !M_left || !M_right || !flag) return XO_EINVAL;
  int threads = 64;
  int grid = (int)((n + threads - 1) / threads);
  xo::launch(contact_points_kernel, grid, threads, 0, static_cast<cudaStream_t>(stream), a, e, cosw, sinw, cosi, sini, L, M_left,
                                                                                M_right, flag, n);
  return xo::launch_status();
}

}  // extern "C"
