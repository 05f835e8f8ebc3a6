// xo_celerite.cu -- celerite2 SHOTerm GP log-likelihood as a parallel scan (BASELINE
// config 4).  celerite2 (>=0.3.1, /root/reference/setup.py:36) has NO call site in the
// reference; the contract here is celerite2's documented SHOTerm/GaussianProcess
// log_likelihood, pinned against a dense Cholesky oracle.
//
// The celerite recurrence is a Riccati (linear-fractional) recursion, so it is run in
// its equivalent 2-state Kalman-filter form, which has an associative combine
// (Sarkka & Garcia-Fernandez 2021).  Three phases, every thread owns a run of RUN consecutive cadences:
//   1. gp_aggregate: walk the run once from a *known* start (P = 0), tracking the affine dependence of the filtered mean
//      on the unknown start state and the information (eta, J) the run's data carry about it -> one scan element per
//      run, no 2x2 inverses; then an in-CTA inclusive scan of the run elements (shuffle scan + 4 warp totals);
//   2. gp_cta_scan: one CTA scans the per-CTA totals and applies each exclusive prefix to the stationary prior -> the
//      state (m, Z = Sigma_inf - P) entering every CTA;
//   3. gp_loglike: each thread applies its in-CTA prefix to that state and re-walks its run with plain Kalman steps,
//      accumulating -1/2 (v^2/D + log D); the last CTA of a chain folds the per-CTA sums in CTA order.
// A CTA's 128 runs are one contiguous tile of the series: it is read with coalesced loads into shared memory (row
// pitch RUN + 1: conflict-free walks), so the dependent chains of the walks never wait for global memory (L1 prefetch
// of each run's lines + operands loaded one step ahead was measured first: the resident CTAs' lines do not fit in L1,
// long_scoreboard stayed at 7-13 stalls per issue).  Two one-pass formulations (tile staged in shared memory, one read of the data, a fixed
// inter-CTA tree with flags instead of the second kernel) were built and measured in round 2: 0.98 ms (128-thread
// tiles) and 1.27 ms (one-warp tiles) against 0.58 ms for this form at 16 chains x 1e6 cadences -- the scans are
// instruction/latency-bound, not bandwidth-bound, and the serial tree look-back costs more than the second read of the
// data saves (profiles/r02_gp_onepass.txt); they live in the history (commits "celerite: one-pass ...").
// SHO in (a, c, h) form so celerite2's eps-regularised coefficients are matched exactly:
//   underdamped k(tau) = a e^{-c tau}(cos h tau + (c/h) sin h tau), overdamped with cosh/sinh.
#include <cuda_runtime.h>

#include <cstdlib>

#include "xo_common.cuh"
#include "xo_math.cuh"

namespace {

#ifndef XO_GP_RUN
#define XO_GP_RUN 16  // cadences per thread (build-time knob)
#endif
constexpr int kRun = XO_GP_RUN;
// the reverse (smoother / adjoint) passes carry more state per step: shorter runs, measured 27% faster at 16 chains
constexpr int kRunB = 8;
constexpr int kThreadsGp = 128;
constexpr int kWarpsGp = kThreadsGp / 32;
constexpr unsigned kFullGp = 0xffffffffu;
constexpr int kAggDoubles = 14;

struct Sho {
  double a, c, h, w2;  // variance, damping, (pseudo)frequency, c^2 +- h^2
  bool under;
};

__device__ __forceinline__ Sho make_sho(double S0, double w0, double Q, double eps) {
  Sho k;
  k.a = S0 * w0 * Q;
  k.c = 0.5 * w0 / Q;
  k.under = Q >= 0.5;
  double f = sqrt(fmax(k.under ? 4.0 * Q * Q - 1.0 : 1.0 - 4.0 * Q * Q, eps));
  k.h = k.c * f;
  k.w2 = k.under ? k.c * k.c + k.h * k.h : k.c * k.c - k.h * k.h;
  return k;
}

// Cf = e^{-c dt} cos(h dt) (cosh when overdamped), Sf = e^{-c dt} sin(h dt) / h (sinh)
__device__ __forceinline__ void sho_cs(const Sho &k, double dt, double &Cf, double &Sf) {
  if (k.under) {
    double sn, cs;
    sincos(k.h * dt, &sn, &cs);
    double ex = exp(-k.c * dt);
    Cf = ex * cs;
    Sf = ex * sn / k.h;
  } else {
    double ep = exp(-(k.c - k.h) * dt), en = exp(-(k.c + k.h) * dt);
    Cf = 0.5 * (ep + en);
    Sf = 0.5 * (ep - en) / k.h;
  }
}

// (Cf, Sf) of an earlier step of the same thread.  On a (nearly) uniform grid consecutive steps
// differ only by the rounding of the time stamps (|delta| ~ 1e-10 dt for Kepler/TESS-like series), and
//   d/dt (Cf, Sf) = (-c Cf -+ h^2 Sf,  Cf - c Sf)
// carries the cached pair to the new dt with a second-order error <= (1e-7 dt (c + h))^2 / 2 --
// 4 flops instead of sincos + exp.  Irregular sampling falls through to the full evaluation.
struct ShoCache {
  double dt = -1.0, Cf = 1.0, Sf = 0.0;
};
__device__ __forceinline__ void sho_cs(const Sho &k, double dt, ShoCache &cc, double &Cf, double &Sf) {
  const double d = dt - cc.dt;
  if (cc.dt > 0.0 && fabs(d) <= 1e-7 * cc.dt) {
    const double h2 = k.under ? k.h * k.h : -k.h * k.h;
    Cf = fma(d, -fma(k.c, cc.Cf, h2 * cc.Sf), cc.Cf);
    Sf = fma(d, fma(-k.c, cc.Sf, cc.Cf), cc.Sf);
    return;
  }
  sho_cs(k, dt, Cf, Sf);
  cc.dt = dt;
  cc.Cf = Cf;
  cc.Sf = Sf;
}

// transition matrix over dt
__device__ __forceinline__ void sho_phi(const Sho &k, double dt, ShoCache &cc, double &f00, double &f01, double &f10,
                                        double &f11) {
  double Cf, Sf;
  sho_cs(k, dt, cc, Cf, Sf);
  f00 = fma(k.c, Sf, Cf);
  f01 = Sf;
  f10 = -k.w2 * Sf;
  f11 = fma(-k.c, Sf, Cf);
}

struct Agg {
  double A00, A01, A10, A11, b0, b1, C00, C01, C11, e0, e1, J00, J01, J11;
};

__device__ __forceinline__ Agg agg_identity() { return Agg{1, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; }

// combine(earlier i, later j)
__device__ __forceinline__ Agg agg_combine(const Agg &i, const Agg &j) {
  // M = I + C_i J_j ; X = M^-1
  double m00 = 1.0 + i.C00 * j.J00 + i.C01 * j.J01, m01 = i.C00 * j.J01 + i.C01 * j.J11;
  double m10 = i.C01 * j.J00 + i.C11 * j.J01, m11 = 1.0 + i.C01 * j.J01 + i.C11 * j.J11;
  double rdet = 1.0 / (m00 * m11 - m01 * m10);
  double x00 = m11 * rdet, x01 = -m01 * rdet, x10 = -m10 * rdet, x11 = m00 * rdet;
  // G = A_j X
  double g00 = j.A00 * x00 + j.A01 * x10, g01 = j.A00 * x01 + j.A01 * x11;
  double g10 = j.A10 * x00 + j.A11 * x10, g11 = j.A10 * x01 + j.A11 * x11;
  Agg o;
  o.A00 = g00 * i.A00 + g01 * i.A10;
  o.A01 = g00 * i.A01 + g01 * i.A11;
  o.A10 = g10 * i.A00 + g11 * i.A10;
  o.A11 = g10 * i.A01 + g11 * i.A11;
  double u0 = i.b0 + i.C00 * j.e0 + i.C01 * j.e1, u1 = i.b1 + i.C01 * j.e0 + i.C11 * j.e1;
  o.b0 = g00 * u0 + g01 * u1 + j.b0;
  o.b1 = g10 * u0 + g11 * u1 + j.b1;
  // T = G C_i ; C = T A_j^T + C_j
  double t00 = g00 * i.C00 + g01 * i.C01, t01 = g00 * i.C01 + g01 * i.C11;
  double t10 = g10 * i.C00 + g11 * i.C01, t11 = g10 * i.C01 + g11 * i.C11;
  o.C00 = t00 * j.A00 + t01 * j.A01 + j.C00;
  o.C01 = 0.5 * ((t00 * j.A10 + t01 * j.A11) + (t10 * j.A00 + t11 * j.A01)) + j.C01;
  o.C11 = t10 * j.A10 + t11 * j.A11 + j.C11;
  // H = A_i^T X^T
  double h00 = i.A00 * x00 + i.A10 * x01, h01 = i.A00 * x10 + i.A10 * x11;
  double h10 = i.A01 * x00 + i.A11 * x01, h11 = i.A01 * x10 + i.A11 * x11;
  double v0 = j.e0 - (j.J00 * i.b0 + j.J01 * i.b1), v1 = j.e1 - (j.J01 * i.b0 + j.J11 * i.b1);
  o.e0 = h00 * v0 + h01 * v1 + i.e0;
  o.e1 = h10 * v0 + h11 * v1 + i.e1;
  // S = H J_j ; J = S A_i + J_i
  double s00 = h00 * j.J00 + h01 * j.J01, s01 = h00 * j.J01 + h01 * j.J11;
  double s10 = h10 * j.J00 + h11 * j.J01, s11 = h10 * j.J01 + h11 * j.J11;
  o.J00 = s00 * i.A00 + s01 * i.A10 + i.J00;
  o.J01 = 0.5 * ((s00 * i.A01 + s01 * i.A11) + (s10 * i.A00 + s11 * i.A10)) + i.J01;
  o.J11 = s10 * i.A01 + s11 * i.A11 + i.J11;
  return o;
}

// apply an aggregate to a Gaussian state (m, P)
__device__ __forceinline__ void agg_apply(const Agg &g, double &m0, double &m1, double &P00, double &P01, double &P11) {
  double q00 = 1.0 + P00 * g.J00 + P01 * g.J01, q01 = P00 * g.J01 + P01 * g.J11;
  double q10 = P01 * g.J00 + P11 * g.J01, q11 = 1.0 + P01 * g.J01 + P11 * g.J11;
  double rdet = 1.0 / (q00 * q11 - q01 * q10);
  double x00 = q11 * rdet, x01 = -q01 * rdet, x10 = -q10 * rdet, x11 = q00 * rdet;
  double g00 = g.A00 * x00 + g.A01 * x10, g01 = g.A00 * x01 + g.A01 * x11;
  double g10 = g.A10 * x00 + g.A11 * x10, g11 = g.A10 * x01 + g.A11 * x11;
  double u0 = m0 + P00 * g.e0 + P01 * g.e1, u1 = m1 + P01 * g.e0 + P11 * g.e1;
  double n0 = g00 * u0 + g01 * u1 + g.b0, n1 = g10 * u0 + g11 * u1 + g.b1;
  double t00 = g00 * P00 + g01 * P01, t01 = g00 * P01 + g01 * P11;
  double t10 = g10 * P00 + g11 * P01, t11 = g10 * P01 + g11 * P11;
  double c00 = t00 * g.A00 + t01 * g.A01 + g.C00;
  double c01 = 0.5 * ((t00 * g.A10 + t01 * g.A11) + (t10 * g.A00 + t11 * g.A01)) + g.C01;
  double c11 = t10 * g.A10 + t11 * g.A11 + g.C11;
  m0 = n0;
  m1 = n1;
  P00 = c00;
  P01 = c01;
  P11 = c11;
}

__device__ __forceinline__ Agg agg_shfl_up(const Agg &a, int delta) {
  Agg o;
  const double *src = &a.A00;
  double *dst = &o.A00;
#pragma unroll
  for (int k = 0; k < kAggDoubles; ++k) dst[k] = __shfl_up_sync(kFullGp, src[k], delta);
  return o;
}

__device__ __forceinline__ void agg_store(double *p, const Agg &a) {
  const double *src = &a.A00;
#pragma unroll
  for (int k = 0; k < kAggDoubles; ++k) p[k] = src[k];
}
__device__ __forceinline__ Agg agg_load(const double *p) {
  Agg a;
  double *dst = &a.A00;
#pragma unroll
  for (int k = 0; k < kAggDoubles; ++k) dst[k] = p[k];
  return a;
}

// inclusive scan of one Agg per thread across the CTA (thread order); returns inclusive value
__device__ Agg cta_inclusive_scan(Agg v, Agg *warp_tot /* smem [kWarps] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    Agg o = agg_shfl_up(v, d);
    if (lane >= d) v = agg_combine(o, v);
  }
  if (lane == 31) warp_tot[warp] = v;
  __syncthreads();
  if (warp > 0) {
    Agg pre = warp_tot[0];
    for (int w = 1; w < warp; ++w) pre = agg_combine(pre, warp_tot[w]);
    v = agg_combine(pre, v);
  }
  return v;
}


// status word at the head of every celerite workspace (see xo_read_status): [0] lanes whose result was not finite,
// [1] reserved (expired waits of the one-pass formulations; always 0 in this form)
struct GpStatus {
  unsigned long long bad_lanes, wait_failures;
};

// sum of log D over a run as the log of a product of mantissas plus an integer exponent sum: one DMUL and a
// few integer instructions per cadence instead of a log; mantissas are in [0.5, 1), so a run of <= 64 cadences
// cannot underflow, and the exponent sum is exact for ANY finite positive D (no D^4 overflow).
struct LogProd {
  double mant = 1.0, slow = 0.0;
  int expo = 0;
  __device__ __forceinline__ void mul(double D) {
    const int hi = __double2hiint(D);
    const int ef = (hi >> 20) & 0x7ff;
    if (hi > 0 && ef != 0 && ef != 0x7ff) {  // positive, normal
      expo += ef - 1022;
      mant *= __hiloint2double((hi & 0x800fffff) | 0x3fe00000, __double2loint(D));
    } else {
      slow += log(D);  // denormal, zero, negative, inf, nan: exact library semantics (NaN / -inf propagate)
    }
  }
  __device__ __forceinline__ double value() const { return log(mant) + slow + (double)expo * 0.69314718055994530941723212145818; }
};


#ifndef XO_GP_RUN_LONG
#define XO_GP_RUN_LONG 32  // run length used when there is parallelism to spare: halves the in-CTA scan work per cadence
#endif
// A CTA's 128 runs are one contiguous tile of the series: it is read from HBM with coalesced loads into shared memory
// (row pitch RUN + 1 doubles: thread r then walks row r conflict-free), the walks never wait for global memory.
template <int RUN>
struct TileSmem {
  double t[kThreadsGp * (RUN + 1)];
  double d[kThreadsGp * (RUN + 1)];  // diag + jitter
  double y[kThreadsGp * (RUN + 1)];
  double t_halo;                      // the time stamp before the tile's first cadence (t[0] itself for the first tile)
};
template <int RUN>
__device__ __forceinline__ void stage_tile(TileSmem<RUN> &sm, const double *__restrict__ t, const double *__restrict__ diag,
                                           const double *__restrict__ yb, double jit, int64_t base, int64_t N) {
  const int nvalid = (int)max((int64_t)0, min((int64_t)kThreadsGp * RUN, N - base));
#pragma unroll 4
  for (int idx = threadIdx.x; idx < kThreadsGp * RUN; idx += kThreadsGp) {
    const int at = (idx / RUN) * (RUN + 1) + idx % RUN;
    if (idx < nvalid) {
      sm.t[at] = __ldg(t + base + idx);
      sm.d[at] = __ldg(diag + base + idx) + jit;
      sm.y[at] = __ldg(yb + base + idx);
    }
  }
  if (threadIdx.x == 0) sm.t_halo = __ldg(t + (base > 0 ? base - 1 : 0));
  __syncthreads();
}

// ---- phase 1 ---------------------------------------------------------------
template <int RUN>
__global__ void __launch_bounds__(kThreadsGp) gp_aggregate_kernel(const double *__restrict__ t, const double *__restrict__ y,
                                                                  const double *__restrict__ diag, int64_t N,
                                                                  const double *__restrict__ S0, const double *__restrict__ w0,
                                                                  const double *__restrict__ Qf,
                                                                  const double *__restrict__ jitter, double eps, int64_t T,
                                                                  double *__restrict__ pref, double *__restrict__ cta_tot,
                                                                  unsigned *__restrict__ done, GpStatus *__restrict__ status) {
  extern __shared__ __align__(16) unsigned char gp_smem_raw[];
  TileSmem<RUN> &sm = *reinterpret_cast<TileSmem<RUN> *>(gp_smem_raw);
  __shared__ Agg warp_tot[kWarpsGp];
  const int64_t bch = blockIdx.y;
  const int64_t g = (int64_t)blockIdx.x * kThreadsGp + threadIdx.x;
  const int64_t s = g * RUN, e = min(s + RUN, N);
  const double *yb = y + bch * N;
  if (threadIdx.x == 0 && blockIdx.x == 0) {  // per-call scratch the later kernels of this stream-ordered sequence count in
    done[bch] = 0u;
    if (bch == 0) status->bad_lanes = status->wait_failures = 0ull;
  }
  const Sho k = make_sho(S0[bch], w0[bch], Qf[bch], eps);
  const double jit = jitter ? jitter[bch] : 0.0;
  stage_tile<RUN>(sm, t, diag, yb, jit, (int64_t)blockIdx.x * kThreadsGp * RUN, N);
  const double *ts = sm.t + threadIdx.x * (RUN + 1), *ds = sm.d + threadIdx.x * (RUN + 1), *ys = sm.y + threadIdx.x * (RUN + 1);
  const double t_before = threadIdx.x > 0 ? sm.t[(threadIdx.x - 1) * (RUN + 1) + RUN - 1] : sm.t_halo;
  Agg v = agg_identity();
  if (s < e) {
    double A00 = 1, A01 = 0, A10 = 0, A11 = 1, b0 = 0, b1 = 0, e0 = 0, e1 = 0, J00 = 0, J01 = 0, J11 = 0;
    double Z00 = k.a, Z01 = 0.0, Z11 = k.a * k.w2;  // P = 0  <=>  Z = Sigma_inf
    double tprev = t_before;
    ShoCache cc;
    const int len = (int)(e - s);
#pragma unroll 1
    for (int j = 0; j < len; ++j) {
      const double ti = ts[j], di = ds[j], yi = ys[j];
      double f00, f01, f10, f11;
      sho_phi(k, ti - tprev, cc, f00, f01, f10, f11);
      tprev = ti;
      // A <- Phi A, b <- Phi b, Z <- Phi Z Phi^T
      double a00 = f00 * A00 + f01 * A10, a01 = f00 * A01 + f01 * A11;
      double a10 = f10 * A00 + f11 * A10, a11 = f10 * A01 + f11 * A11;
      double c0 = f00 * b0 + f01 * b1, c1 = f10 * b0 + f11 * b1;
      double z00 = f00 * Z00 + f01 * Z01, z01 = f00 * Z01 + f01 * Z11;
      double z10 = f10 * Z00 + f11 * Z01, z11 = f10 * Z01 + f11 * Z11;
      double p00 = z00 * f00 + z01 * f01, p01 = z00 * f10 + z01 * f11, p11 = z10 * f10 + z11 * f11;
      const double pm = k.a - p00;
      const double D = pm + di;
      const double rD = xo::rcp_fast(D);
      const double K0 = pm * rD, K1 = -p01 * rD;
      const double vb = yi - c0;
      J00 = fma(a00 * a00, rD, J00);
      J01 = fma(a00 * a01, rD, J01);
      J11 = fma(a01 * a01, rD, J11);
      e0 = fma(a00 * vb, rD, e0);
      e1 = fma(a01 * vb, rD, e1);
      A00 = a00 - K0 * a00;
      A01 = a01 - K0 * a01;
      A10 = a10 - K1 * a00;
      A11 = a11 - K1 * a01;
      b0 = fma(K0, vb, c0);
      b1 = fma(K1, vb, c1);
      Z00 = fma(D * K0, K0, p00);
      Z01 = fma(D * K0, K1, p01);
      Z11 = fma(D * K1, K1, p11);
    }
    v = Agg{A00, A01, A10, A11, b0, b1, k.a - Z00, -Z01, k.a * k.w2 - Z11, e0, e1, J00, J01, J11};
  }
  v = cta_inclusive_scan(v, warp_tot);
  if (g < T) agg_store(pref + (bch * T + g) * kAggDoubles, v);
  if (threadIdx.x == kThreadsGp - 1) agg_store(cta_tot + (bch * gridDim.x + blockIdx.x) * kAggDoubles, v);
}

// ---- phase 2: scan the CTA totals, emit the state entering each CTA ---------
__global__ void __launch_bounds__(kThreadsGp) gp_cta_scan_kernel(const double *__restrict__ cta_tot, int nctas,
                                                                 const double *__restrict__ S0,
                                                                 const double *__restrict__ w0,
                                                                 const double *__restrict__ Qf, double eps,
                                                                 double *__restrict__ cta_state) {
  __shared__ Agg warp_tot[kWarpsGp];
  __shared__ Agg carry_s;
  const int64_t bch = blockIdx.x;
  const Sho k = make_sho(S0[bch], w0[bch], Qf[bch], eps);
  Agg carry = agg_identity();
  for (int base = 0; base < nctas; base += kThreadsGp) {
    const int idx = base + threadIdx.x;
    Agg v = idx < nctas ? agg_load(cta_tot + (bch * nctas + idx) * kAggDoubles) : agg_identity();
    v = cta_inclusive_scan(v, warp_tot);
    Agg incl = agg_combine(carry, v);
    // exclusive prefix of idx = inclusive of idx-1
    Agg prev = agg_shfl_up(incl, 1);
    __shared__ Agg lastlane[kWarpsGp];
    if ((threadIdx.x & 31) == 31) lastlane[threadIdx.x >> 5] = incl;
    __syncthreads();
    Agg excl;
    if ((threadIdx.x & 31) == 0)
      excl = (threadIdx.x == 0) ? carry : lastlane[(threadIdx.x >> 5) - 1];
    else
      excl = prev;
    if (idx < nctas) {
      double m0 = 0.0, m1 = 0.0, P00 = k.a, P01 = 0.0, P11 = k.a * k.w2;
      agg_apply(excl, m0, m1, P00, P01, P11);
      double *o = cta_state + (bch * nctas + idx) * 5;
      o[0] = m0;
      o[1] = m1;
      o[2] = k.a - P00;  // Z = Sigma_inf - P
      o[3] = -P01;
      o[4] = k.a * k.w2 - P11;
    }
    if (threadIdx.x == kThreadsGp - 1) carry_s = incl;
    __syncthreads();
    carry = carry_s;
    __syncthreads();
  }
}


// ---- phase 3 ---------------------------------------------------------------
template <int RUN, int REC /* 0: none, 1: (v/D, K0, K1) per cadence for the smoother, 2: state entering every run of kRunB cadences */>
__global__ void __launch_bounds__(kThreadsGp) gp_loglike_kernel(const double *__restrict__ t, const double *__restrict__ y,
                                                                const double *__restrict__ diag, int64_t N,
                                                                const double *__restrict__ S0, const double *__restrict__ w0,
                                                                const double *__restrict__ Qf, const double *__restrict__ jitter,
                                                                double eps, int64_t T, const double *__restrict__ pref,
                                                                const double *__restrict__ cta_state, double *__restrict__ cta_ll,
                                                                unsigned *__restrict__ done, GpStatus *__restrict__ status,
                                                                double *__restrict__ loglike, double *__restrict__ rec, int64_t Tb) {
  extern __shared__ __align__(16) unsigned char gp_smem_raw[];
  TileSmem<RUN> &sm = *reinterpret_cast<TileSmem<RUN> *>(gp_smem_raw);
  __shared__ double red[kWarpsGp];
  __shared__ int is_last;
  const int64_t bch = blockIdx.y;
  const int64_t g = (int64_t)blockIdx.x * kThreadsGp + threadIdx.x;
  const int64_t s = g * RUN, e = min(s + RUN, N);
  const double *yb = y + bch * N;
  const Sho k = make_sho(S0[bch], w0[bch], Qf[bch], eps);
  const double jit = jitter ? jitter[bch] : 0.0;
  stage_tile<RUN>(sm, t, diag, yb, jit, (int64_t)blockIdx.x * kThreadsGp * RUN, N);
  const double *ts = sm.t + threadIdx.x * (RUN + 1), *ds = sm.d + threadIdx.x * (RUN + 1), *ys = sm.y + threadIdx.x * (RUN + 1);
  const double t_before = threadIdx.x > 0 ? sm.t[(threadIdx.x - 1) * (RUN + 1) + RUN - 1] : sm.t_halo;
  double acc = 0.0;
  if (s < e) {
    const double *st = cta_state + (bch * gridDim.x + blockIdx.x) * 5;
    double m0 = st[0], m1 = st[1], Z00 = st[2], Z01 = st[3], Z11 = st[4];
    if (threadIdx.x > 0) {
      double P00 = k.a - Z00, P01 = -Z01, P11 = k.a * k.w2 - Z11;
      Agg pre = agg_load(pref + (bch * T + g - 1) * kAggDoubles);
      agg_apply(pre, m0, m1, P00, P01, P11);
      Z00 = k.a - P00;
      Z01 = -P01;
      Z11 = k.a * k.w2 - P11;
    }
    double tprev = t_before;
    ShoCache cc;
    LogProd lp;
    const int len = (int)(e - s);
#pragma unroll 1
    for (int j = 0; j < len; ++j) {
      const int64_t i = s + j;
      const double ti = ts[j], di = ds[j], yi = ys[j];
      double f00, f01, f10, f11;
      sho_phi(k, ti - tprev, cc, f00, f01, f10, f11);
      tprev = ti;
      if (REC == 2 && ((i - s) % kRunB) == 0) {  // state entering every run of kRunB cadences: where the reverse pass restarts
        double *rb = rec + bch * 5 * Tb + i / kRunB;
        rb[0] = m0;
        rb[Tb] = m1;
        rb[2 * Tb] = Z00;
        rb[3 * Tb] = Z01;
        rb[4 * Tb] = Z11;
      }
      double c0 = f00 * m0 + f01 * m1, c1 = f10 * m0 + f11 * m1;
      double z00 = f00 * Z00 + f01 * Z01, z01 = f00 * Z01 + f01 * Z11;
      double z10 = f10 * Z00 + f11 * Z01, z11 = f10 * Z01 + f11 * Z11;
      double p00 = z00 * f00 + z01 * f01, p01 = z00 * f10 + z01 * f11, p11 = z10 * f10 + z11 * f11;
      const double pm = k.a - p00;
      const double D = pm + di;
      const double rD = xo::rcp_fast(D);
      const double K0 = pm * rD, K1 = -p01 * rD;
      const double v = yi - c0;
      acc = fma(v * v, rD, acc);
      lp.mul(D);
      if (REC == 1) {  // what the backward (disturbance-smoother) pass needs
        double *rb = rec + bch * 3 * N;
        rb[i] = v * rD;
        rb[N + i] = K0;
        rb[2 * N + i] = K1;
      }
      m0 = fma(K0, v, c0);
      m1 = fma(K1, v, c1);
      Z00 = fma(D * K0, K0, p00);
      Z01 = fma(D * K0, K1, p01);
      Z11 = fma(D * K1, K1, p11);
    }
    acc += lp.value();
    if (!isfinite(acc)) atomicAdd(&status->bad_lanes, 1ull);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFullGp, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s4 = 0.0;
    for (int w = 0; w < kWarpsGp; ++w) s4 += red[w];
    cta_ll[bch * gridDim.x + blockIdx.x] = s4;
    __threadfence();
    is_last = atomicAdd(&done[bch], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last && threadIdx.x < 32) {  // the last CTA of the chain folds the per-CTA sums in CTA order (bit-reproducible)
    __threadfence();
    double tot = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) tot += __ldcg(cta_ll + bch * gridDim.x + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(kFullGp, tot, o);
    if (threadIdx.x == 0) {
      loglike[bch] = -0.5 * tot - 0.5 * (double)N * 1.8378770664093454835606594728112;  // log(2 pi)
      done[bch] = 0u;  // the reverse pass counts its CTAs in the same word
    }
  }
}

// ---------------------------------------------------------------------------
// alpha = K^-1 y by the disturbance smoother (Durbin & Koopman 2012, sec. 4.5), run as an
// affine scan in reverse time: with w_k = v_k / D_k, K_k the Kalman gain and Phi_{k+1} the next
// transition,   R_{k-1} = (I - K_k h)^T Phi_{k+1}^T R_k + h^T w_k,  R_{N-1} = 0,
//               alpha_k = w_k - (Phi_{k+1} K_k) . R_k.
// Same three phases as the forward pass, thread g owning run T-1-g so the scan runs right-to-left;
// the scan element is an affine map (A 2x2, b 2), combine = composition.
// ---------------------------------------------------------------------------
struct Aff {
  double a00, a01, a10, a11, b0, b1;
};
__device__ __forceinline__ Aff aff_identity() { return Aff{1, 0, 0, 1, 0, 0}; }
// cur after prev:  x -> A_cur (A_prev x + b_prev) + b_cur
__device__ __forceinline__ Aff aff_after(const Aff &prev, const Aff &cur) {
  Aff o;
  o.a00 = cur.a00 * prev.a00 + cur.a01 * prev.a10;
  o.a01 = cur.a00 * prev.a01 + cur.a01 * prev.a11;
  o.a10 = cur.a10 * prev.a00 + cur.a11 * prev.a10;
  o.a11 = cur.a10 * prev.a01 + cur.a11 * prev.a11;
  o.b0 = cur.a00 * prev.b0 + cur.a01 * prev.b1 + cur.b0;
  o.b1 = cur.a10 * prev.b0 + cur.a11 * prev.b1 + cur.b1;
  return o;
}
__device__ __forceinline__ Aff aff_shfl_up(const Aff &a, int d) {
  Aff o;
  const double *src = &a.a00;
  double *dst = &o.a00;
#pragma unroll
  for (int k = 0; k < 6; ++k) dst[k] = __shfl_up_sync(kFullGp, src[k], d);
  return o;
}
__device__ Aff aff_cta_inclusive_scan(Aff v, Aff *warp_tot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    Aff o = aff_shfl_up(v, d);
    if (lane >= d) v = aff_after(o, v);
  }
  if (lane == 31) warp_tot[warp] = v;
  __syncthreads();
  if (warp > 0) {
    Aff pre = warp_tot[0];
    for (int w = 1; w < warp; ++w) pre = aff_after(pre, warp_tot[w]);
    v = aff_after(pre, v);
  }
  return v;
}

// one backward step k of run walking: given R (= R_k) returns alpha_k and updates R to R_{k-1}
__device__ __forceinline__ double gp_bwd_step(const Sho &k, const double *__restrict__ t, const double *__restrict__ rb,
                                              int64_t N, int64_t i, ShoCache &cc, double &R0, double &R1) {
  const double w = __ldg(rb + i), K0 = __ldg(rb + N + i), K1 = __ldg(rb + 2 * N + i);
  double f00 = 1.0, f01 = 0.0, f10 = 0.0, f11 = 1.0;
  if (i + 1 < N) sho_phi(k, __ldg(t + i + 1) - __ldg(t + i), cc, f00, f01, f10, f11);
  // q = Phi_{k+1}^T R
  const double q0 = f00 * R0 + f10 * R1, q1 = f01 * R0 + f11 * R1;
  const double kq = K0 * q0 + K1 * q1;  // (Phi K) . R = K . (Phi^T R)
  const double alpha = w - kq;
  R0 = q0 - kq + w;  // (I - K h)^T q + h^T w,  h = (1, 0)
  R1 = q1;
  return alpha;
}

__global__ void __launch_bounds__(kThreadsGp) gp_bwd_aggregate_kernel(const double *__restrict__ t, int64_t N,
                                                                      const double *__restrict__ S0,
                                                                      const double *__restrict__ w0,
                                                                      const double *__restrict__ Qf, double eps,
                                                                      int64_t T, const double *__restrict__ rec,
                                                                      double *__restrict__ bpref,
                                                                      double *__restrict__ bcta) {
  __shared__ Aff warp_tot[kWarpsGp];
  const int64_t bch = blockIdx.y;
  const Sho k = make_sho(S0[bch], w0[bch], Qf[bch], eps);
  const int64_t g = (int64_t)blockIdx.x * kThreadsGp + threadIdx.x;
  const int64_t run = T - 1 - g;  // reversed ownership
  Aff v = aff_identity();
  if (run >= 0) {
    const int64_t s = run * kRunB, e = min(s + kRunB, N);
    const double *rb = rec + bch * 3 * N;
    // the run's map R_{e-1} -> R_{s-1}: push the two basis vectors and the origin through it
    double o0 = 0.0, o1 = 0.0, x0 = 1.0, x1 = 0.0, y0 = 0.0, y1 = 1.0;
    ShoCache cc;
    for (int64_t i = e - 1; i >= s; --i) {
      const double K0 = __ldg(rb + N + i), K1 = __ldg(rb + 2 * N + i), w = __ldg(rb + i);
      double f00 = 1.0, f01 = 0.0, f10 = 0.0, f11 = 1.0;
      if (i + 1 < N) sho_phi(k, __ldg(t + i + 1) - __ldg(t + i), cc, f00, f01, f10, f11);
      auto lin = [&](double &r0, double &r1, double add) {
        const double q0 = f00 * r0 + f10 * r1, q1 = f01 * r0 + f11 * r1;
        r0 = q0 - (K0 * q0 + K1 * q1) + add;
        r1 = q1;
      };
      lin(o0, o1, w);
      lin(x0, x1, w);
      lin(y0, y1, w);
    }
    v = Aff{x0 - o0, y0 - o0, x1 - o1, y1 - o1, o0, o1};
  }
  v = aff_cta_inclusive_scan(v, warp_tot);
  if (g < T) {
    double *o = bpref + (bch * T + g) * 6;
    o[0] = v.a00; o[1] = v.a01; o[2] = v.a10; o[3] = v.a11; o[4] = v.b0; o[5] = v.b1;
  }
  if (threadIdx.x == kThreadsGp - 1) {
    double *o = bcta + (bch * gridDim.x + blockIdx.x) * 6;
    o[0] = v.a00; o[1] = v.a01; o[2] = v.a10; o[3] = v.a11; o[4] = v.b0; o[5] = v.b1;
  }
}

// R entering each CTA (from the right): exclusive scan of the CTA totals applied to R = 0
__global__ void __launch_bounds__(kThreadsGp) gp_bwd_cta_scan_kernel(const double *__restrict__ bcta, int nctas,
                                                                     double *__restrict__ bstate) {
  __shared__ Aff warp_tot[kWarpsGp];
  __shared__ Aff lastlane[kWarpsGp];
  __shared__ Aff carry_s;
  const int64_t bch = blockIdx.x;
  Aff carry = aff_identity();
  for (int base = 0; base < nctas; base += kThreadsGp) {
    const int idx = base + threadIdx.x;
    Aff v = aff_identity();
    if (idx < nctas) {
      const double *p = bcta + (bch * nctas + idx) * 6;
      v = Aff{p[0], p[1], p[2], p[3], p[4], p[5]};
    }
    v = aff_cta_inclusive_scan(v, warp_tot);
    Aff incl = aff_after(carry, v);
    Aff prev = aff_shfl_up(incl, 1);
    if ((threadIdx.x & 31) == 31) lastlane[threadIdx.x >> 5] = incl;
    __syncthreads();
    Aff excl = ((threadIdx.x & 31) == 0) ? (threadIdx.x == 0 ? carry : lastlane[(threadIdx.x >> 5) - 1]) : prev;
    if (idx < nctas) {
      double *o = bstate + (bch * nctas + idx) * 2;
      o[0] = excl.b0;  // the map applied to R = 0
      o[1] = excl.b1;
    }
    if (threadIdx.x == kThreadsGp - 1) carry_s = incl;
    __syncthreads();
    carry = carry_s;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreadsGp) gp_bwd_solve_kernel(const double *__restrict__ t, int64_t N,
                                                                  const double *__restrict__ S0,
                                                                  const double *__restrict__ w0,
                                                                  const double *__restrict__ Qf, double eps, int64_t T,
                                                                  const double *__restrict__ rec,
                                                                  const double *__restrict__ bpref,
                                                                  const double *__restrict__ bstate,
                                                                  double *__restrict__ alpha) {
  const int64_t bch = blockIdx.y;
  const Sho k = make_sho(S0[bch], w0[bch], Qf[bch], eps);
  const int64_t g = (int64_t)blockIdx.x * kThreadsGp + threadIdx.x;
  const int64_t run = T - 1 - g;
  if (run < 0) return;
  const int64_t s = run * kRunB, e = min(s + kRunB, N);
  const double *st = bstate + (bch * gridDim.x + blockIdx.x) * 2;
  double R0 = st[0], R1 = st[1];
  if (threadIdx.x > 0) {  // apply the in-CTA exclusive prefix (inclusive of the previous thread)
    const double *p = bpref + (bch * T + g - 1) * 6;
    const double n0 = p[0] * R0 + p[1] * R1 + p[4], n1 = p[2] * R0 + p[3] * R1 + p[5];
    R0 = n0;
    R1 = n1;
  }
  const double *rb = rec + bch * 3 * N;
  double *ab = alpha + bch * N;
  ShoCache cc;
  for (int64_t i = e - 1; i >= s; --i) ab[i] = gp_bwd_step(k, t, rb, N, i, cc, R0, R1);
}

// ---------------------------------------------------------------------------
// Reverse-mode gradient of the log-likelihood w.r.t. (S0, w0, Q), diag[N] and y[N] -- what
// celerite2's own reverse pass provides, here as a second parallel scan.
//
// Forward step i maps the state s = (m, Z) to
//   cvec = Phi m,  Pz = Phi Z Phi^T,  k = (a - Pz00, -Pz01),  D = k0 + diag_i,  v = y_i - cvec0,
//   l_i = -1/2 (v^2/D + log D),       m' = cvec + k v/D,       Z' = Pz + k k^T / D.
// Its adjoint  lambda_{i-1} = J_i^T lambda_i + d l_i / d s_{i-1}  is AFFINE in the 5-vector
// lambda = (mbar0, mbar1, Zbar00, Zbar01, Zbar11), with block-triangular J^T (mbar does not depend
// on Zbar), so it runs as a reverse affine scan exactly like the disturbance smoother above:
// thread g owns run T-1-g, composes its run's map (24 doubles), in-CTA shuffle scan, one CTA over
// the CTA totals, then every thread re-walks its run backwards with the now known incoming lambda
// and accumulates the parameter cotangents.  The forward quantities of a step are recomputed from
// the state entering every run of kRunB cadences, which the forward scan leaves behind.
// ---------------------------------------------------------------------------
constexpr int kAdjDoubles = 24;
struct Adj {
  double A[4];  // mbar_in = A mbar_out + om
  double B[6];  // Zbar_in = B mbar_out + C Zbar_out + oz   (B 3x2, C 3x3, row-major)
  double C[9];
  double om[2], oz[3];
};
__device__ __forceinline__ Adj adj_identity() {
  Adj a;
  double *d = a.A;
#pragma unroll
  for (int i = 0; i < kAdjDoubles; ++i) d[i] = 0.0;
  a.A[0] = a.A[3] = 1.0;
  a.C[0] = a.C[4] = a.C[8] = 1.0;
  return a;
}
// cur after prev:  lambda -> cur(prev(lambda))
__device__ __forceinline__ Adj adj_after(const Adj &p, const Adj &c) {
  Adj o;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
#pragma unroll
    for (int q = 0; q < 2; ++q) o.A[2 * r + q] = c.A[2 * r] * p.A[q] + c.A[2 * r + 1] * p.A[2 + q];
    o.om[r] = c.A[2 * r] * p.om[0] + c.A[2 * r + 1] * p.om[1] + c.om[r];
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int q = 0; q < 2; ++q)
      o.B[2 * r + q] = c.B[2 * r] * p.A[q] + c.B[2 * r + 1] * p.A[2 + q] + c.C[3 * r] * p.B[q] + c.C[3 * r + 1] * p.B[2 + q] +
                       c.C[3 * r + 2] * p.B[4 + q];
#pragma unroll
    for (int q = 0; q < 3; ++q) o.C[3 * r + q] = c.C[3 * r] * p.C[q] + c.C[3 * r + 1] * p.C[3 + q] + c.C[3 * r + 2] * p.C[6 + q];
    o.oz[r] = c.B[2 * r] * p.om[0] + c.B[2 * r + 1] * p.om[1] + c.C[3 * r] * p.oz[0] + c.C[3 * r + 1] * p.oz[1] +
              c.C[3 * r + 2] * p.oz[2] + c.oz[r];
  }
  return o;
}
__device__ __forceinline__ void adj_apply(const Adj &a, double (&L)[5]) {
  const double m0 = a.A[0] * L[0] + a.A[1] * L[1] + a.om[0], m1 = a.A[2] * L[0] + a.A[3] * L[1] + a.om[1];
  double z[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    z[r] = a.B[2 * r] * L[0] + a.B[2 * r + 1] * L[1] + a.C[3 * r] * L[2] + a.C[3 * r + 1] * L[3] + a.C[3 * r + 2] * L[4] + a.oz[r];
  L[0] = m0;
  L[1] = m1;
  L[2] = z[0];
  L[3] = z[1];
  L[4] = z[2];
}

__device__ __forceinline__ Adj adj_shfl_up(const Adj &a, int d) {
  Adj o;
  const double *src = a.A;
  double *dst = o.A;
#pragma unroll
  for (int k = 0; k < kAdjDoubles; ++k) dst[k] = __shfl_up_sync(kFullGp, src[k], d);
  return o;
}
__device__ __forceinline__ void adj_store(double *p, const Adj &a) {
  const double *src = a.A;
#pragma unroll
  for (int k = 0; k < kAdjDoubles; ++k) p[k] = src[k];
}
__device__ __forceinline__ Adj adj_load(const double *p) {
  Adj a;
  double *dst = a.A;
#pragma unroll
  for (int k = 0; k < kAdjDoubles; ++k) dst[k] = p[k];
  return a;
}
__device__ Adj adj_cta_inclusive_scan(Adj v, Adj *warp_tot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {
    Adj o = adj_shfl_up(v, d);
    if (lane >= d) v = adj_after(o, v);
  }
  if (lane == 31) warp_tot[warp] = v;
  __syncthreads();
  if (warp > 0) {
    Adj pre = warp_tot[0];
    for (int w = 1; w < warp; ++w) pre = adj_after(pre, warp_tot[w]);
    v = adj_after(pre, v);
  }
  return v;
}

// forward quantities of step i, recomputed from the recorded state entering it
struct StepFwd {
  double f00, f01, f10, f11, Cf, Sf, dt;
  double m0, m1;
  double w00, w01, w10, w11;  // W = Phi Z
  double k0, k1, rD, w;       // k, 1/D, v/D
};
__device__ __forceinline__ StepFwd gp_step_fwd(const Sho &k, double dt, double yi, double di, double m0, double m1, double Z00,
                                               double Z01, double Z11, ShoCache &cc) {
  StepFwd s;
  s.dt = dt;
  sho_cs(k, s.dt, cc, s.Cf, s.Sf);
  s.f00 = fma(k.c, s.Sf, s.Cf);
  s.f01 = s.Sf;
  s.f10 = -k.w2 * s.Sf;
  s.f11 = fma(-k.c, s.Sf, s.Cf);
  s.m0 = m0;
  s.m1 = m1;
  s.w00 = s.f00 * Z00 + s.f01 * Z01;
  s.w01 = s.f00 * Z01 + s.f01 * Z11;
  s.w10 = s.f10 * Z00 + s.f11 * Z01;
  s.w11 = s.f10 * Z01 + s.f11 * Z11;
  const double p00 = s.w00 * s.f00 + s.w01 * s.f01, p01 = s.w00 * s.f10 + s.w01 * s.f11;
  s.k0 = k.a - p00;
  s.k1 = -p01;
  s.rD = xo::rcp_fast(s.k0 + di);
  s.w = (yi - (s.f00 * s.m0 + s.f01 * s.m1)) * s.rD;
  return s;
}

// lambda_i -> lambda_{i-1}.  kForce: include d l_i / d(v, D) (the affine part); kAcc: also return
// acc = (vbar, Dbar, abar, Phibar00, Phibar01, Phibar10, Phibar11) of this step.
template <bool kForce, bool kAcc>
__device__ __forceinline__ void gp_step_adj(const StepFwd &s, double (&L)[5], double *acc) {
  const double mb0 = L[0], mb1 = L[1], zb00 = L[2], zb01 = L[3], zb11 = L[4];
  const double km = s.k0 * mb0 + s.k1 * mb1;
  const double q = zb00 * s.k0 * s.k0 + zb01 * s.k0 * s.k1 + zb11 * s.k1 * s.k1;
  double vbar = km * s.rD;
  double Dbar = -(km * s.w + q * s.rD) * s.rD;
  if (kForce) {
    vbar -= s.w;
    Dbar += 0.5 * (s.w * s.w - s.rD);
  }
  const double kb0 = mb0 * s.w + (2.0 * zb00 * s.k0 + zb01 * s.k1) * s.rD;
  const double kb1 = mb1 * s.w + (zb01 * s.k0 + 2.0 * zb11 * s.k1) * s.rD;
  const double cb0 = mb0 - vbar, cb1 = mb1;
  const double pb00 = zb00 - kb0 - Dbar, hb = 0.5 * (zb01 - kb1), pb11 = zb11;
  // Zbar = Phi^T Sbar Phi, Sbar = [[pb00, hb], [hb, pb11]]
  const double u00 = pb00 * s.f00 + hb * s.f10, u01 = pb00 * s.f01 + hb * s.f11;
  const double u10 = hb * s.f00 + pb11 * s.f10, u11 = hb * s.f01 + pb11 * s.f11;
  L[2] = s.f00 * u00 + s.f10 * u10;
  L[3] = 2.0 * (s.f00 * u01 + s.f10 * u11);
  L[4] = s.f01 * u01 + s.f11 * u11;
  L[0] = s.f00 * cb0 + s.f10 * cb1;
  L[1] = s.f01 * cb0 + s.f11 * cb1;
  if (kAcc) {
    acc[0] = vbar;
    acc[1] = Dbar;
    acc[2] = kb0 + Dbar;
    // Phibar = cbar m^T + 2 Sbar W
    acc[3] = cb0 * s.m0 + 2.0 * (pb00 * s.w00 + hb * s.w10);
    acc[4] = cb0 * s.m1 + 2.0 * (pb00 * s.w01 + hb * s.w11);
    acc[5] = cb1 * s.m0 + 2.0 * (hb * s.w00 + pb11 * s.w10);
    acc[6] = cb1 * s.m1 + 2.0 * (hb * s.w01 + pb11 * s.w11);
  }
}

// ---- reverse (adjoint) pass: three kernels, thread g owns run T-1-g of kRunB cadences ---------------------------
// The forward pass leaves only the state entering every run of kRunB cadences (5 doubles per run: 5 B per cadence instead
// of the 40 B per cadence of a per-step record); each thread re-derives the states inside its run into shared memory
// (thread-contiguous, conflict-free) before walking the run backwards.
struct RunStates {
  double st[5][kRunB][kThreadsGp];
};
struct GradSmem {
  TileSmem<kRunB> tile;  // the CTA's 128 runs: t, diag + jitter, y (the final kernel overwrites y / diag rows with d/dy, d/d diag)
  RunStates R;
};
// The CTA with index bx owns runs T-1-(128 bx + tid), tid = 0..127: the contiguous block of time-runs
// [r_lo, r_lo + 128) with r_lo = T - 128 (bx + 1) (negative for the last CTA: clipped, its first rows are then unused).
struct GradTile {
  int64_t r_lo, base;  // first time-run of the tile (may be negative), first cadence staged
  int row0;            // smem row of time-run max(r_lo, 0)
};
__device__ __forceinline__ GradTile grad_tile(int64_t T) {
  GradTile g;
  g.r_lo = T - (int64_t)kThreadsGp * (blockIdx.x + 1);
  const int64_t first = g.r_lo > 0 ? g.r_lo : 0;
  g.base = first * kRunB;
  g.row0 = 0;
  return g;
}
// states entering every step of the run in smem row `row` (cadences [s, e)) -> R.st[.][j][tid]
__device__ __forceinline__ void gp_rederive(const Sho &k, const TileSmem<kRunB> &tile, int row, const double *__restrict__ rp, int64_t Tb,
                                            int len, RunStates &R) {
  const int tid = threadIdx.x;
  const double *ts = tile.t + row * (kRunB + 1), *ds = tile.d + row * (kRunB + 1), *ys = tile.y + row * (kRunB + 1);
  double m0 = __ldg(rp), m1 = __ldg(rp + Tb), Z00 = __ldg(rp + 2 * Tb), Z01 = __ldg(rp + 3 * Tb), Z11 = __ldg(rp + 4 * Tb);
  double tprev = row > 0 ? tile.t[(row - 1) * (kRunB + 1) + kRunB - 1] : tile.t_halo;
  ShoCache cc;
#pragma unroll 1
  for (int j = 0; j < len; ++j) {
    R.st[0][j][tid] = m0;
    R.st[1][j][tid] = m1;
    R.st[2][j][tid] = Z00;
    R.st[3][j][tid] = Z01;
    R.st[4][j][tid] = Z11;
    if (j + 1 == len) break;
    const double ti = ts[j];
    double f00, f01, f10, f11;
    sho_phi(k, ti - tprev, cc, f00, f01, f10, f11);
    tprev = ti;
    double c0 = f00 * m0 + f01 * m1, c1 = f10 * m0 + f11 * m1;
    double z00 = f00 * Z00 + f01 * Z01, z01 = f00 * Z01 + f01 * Z11;
    double z10 = f10 * Z00 + f11 * Z01, z11 = f10 * Z01 + f11 * Z11;
    double p00 = z00 * f00 + z01 * f01, p01 = z00 * f10 + z01 * f11, p11 = z10 * f10 + z11 * f11;
    const double pm = k.a - p00;
    const double D = pm + ds[j];
    const double rD = xo::rcp_fast(D);
    const double K0 = pm * rD, K1 = -p01 * rD;
    const double vv = ys[j] - c0;
    m0 = fma(K0, vv, c0);
    m1 = fma(K1, vv, c1);
    Z00 = fma(D * K0, K0, p00);
    Z01 = fma(D * K0, K1, p01);
    Z11 = fma(D * K1, K1, p11);
  }
}
// forward quantities of step j of the run in smem row `row`
__device__ __forceinline__ StepFwd gp_step_at(const Sho &k, const TileSmem<kRunB> &tile, int row, int j, const RunStates &R, ShoCache &cc) {
  const int tid = threadIdx.x;
  const double *ts = tile.t + row * (kRunB + 1);
  const double tp = j > 0 ? ts[j - 1] : (row > 0 ? tile.t[(row - 1) * (kRunB + 1) + kRunB - 1] : tile.t_halo);
  return gp_step_fwd(k, ts[j] - tp, tile.y[row * (kRunB + 1) + j], tile.d[row * (kRunB + 1) + j], R.st[0][j][tid], R.st[1][j][tid],
                     R.st[2][j][tid], R.st[3][j][tid], R.st[4][j][tid], cc);
}

// the same step for a lambda whose mean part is zero and stays zero (the three Zbar basis vectors of a run's map: mbar
// never depends on Zbar): only L[2..4] are read and written
__device__ __forceinline__ void gp_step_adj_z(const StepFwd &s, double (&L)[5]) {
  const double zb00 = L[2], zb01 = L[3], zb11 = L[4];
  const double q = zb00 * s.k0 * s.k0 + zb01 * s.k0 * s.k1 + zb11 * s.k1 * s.k1;
  const double Dbar = -q * s.rD * s.rD;
  const double kb0 = (2.0 * zb00 * s.k0 + zb01 * s.k1) * s.rD;
  const double kb1 = (zb01 * s.k0 + 2.0 * zb11 * s.k1) * s.rD;
  const double pb00 = zb00 - kb0 - Dbar, hb = 0.5 * (zb01 - kb1), pb11 = zb11;
  const double u00 = pb00 * s.f00 + hb * s.f10, u01 = pb00 * s.f01 + hb * s.f11;
  const double u10 = hb * s.f00 + pb11 * s.f10, u11 = hb * s.f01 + pb11 * s.f11;
  L[2] = s.f00 * u00 + s.f10 * u10;
  L[3] = 2.0 * (s.f00 * u01 + s.f10 * u11);
  L[4] = s.f01 * u01 + s.f11 * u11;
}

__global__ void __launch_bounds__(kThreadsGp, 3)
    gp_grad_aggregate_kernel(const double *__restrict__ t, const double *__restrict__ y, const double *__restrict__ diag, int64_t N,
                             const double *__restrict__ S0, const double *__restrict__ w0, const double *__restrict__ Qf,
                             const double *__restrict__ jitter, double eps, int64_t T, const double *__restrict__ rs,
                             double *__restrict__ gpref, double *__restrict__ gcta) {
  extern __shared__ __align__(16) unsigned char gp_smem_raw[];
  GradSmem &sm = *reinterpret_cast<GradSmem *>(gp_smem_raw);
  RunStates &R = sm.R;
  __shared__ Adj warp_tot[kWarpsGp];
  const int64_t bch = blockIdx.y;
  const Sho k = make_sho(S0[bch], w0[bch], Qf[bch], eps);
  const double jit = jitter ? jitter[bch] : 0.0;
  const int64_t g = (int64_t)blockIdx.x * kThreadsGp + threadIdx.x;
  const int64_t run = T - 1 - g;  // reversed ownership
  const GradTile gt = grad_tile(T);
  stage_tile<kRunB>(sm.tile, t, diag, y + bch * N, jit, gt.base, N);
  Adj v = adj_identity();
  if (run >= 0) {
    const int64_t s = run * kRunB, e = min(s + kRunB, N);
    const int row = (int)(run - (gt.r_lo > 0 ? gt.r_lo : 0)), len = (int)(e - s);
    gp_rederive(k, sm.tile, row, rs + (size_t)bch * 5 * T + run, T, len, R);
    // the run's map lambda_{e-1} -> lambda_{s-1}: push the origin (with the forcing) and the five basis vectors through it
    double O[5] = {0, 0, 0, 0, 0}, X0[5] = {1, 0, 0, 0, 0}, X1[5] = {0, 1, 0, 0, 0}, X2[5] = {0, 0, 1, 0, 0},
           X3[5] = {0, 0, 0, 1, 0}, X4[5] = {0, 0, 0, 0, 1};
    ShoCache cc;
#pragma unroll 1
    for (int j = len - 1; j >= 0; --j) {
      const StepFwd f = gp_step_at(k, sm.tile, row, j, R, cc);
      gp_step_adj<true, false>(f, O, nullptr);
      gp_step_adj<false, false>(f, X0, nullptr);
      gp_step_adj<false, false>(f, X1, nullptr);
      gp_step_adj_z(f, X2);
      gp_step_adj_z(f, X3);
      gp_step_adj_z(f, X4);
    }
    v.A[0] = X0[0]; v.A[1] = X1[0]; v.A[2] = X0[1]; v.A[3] = X1[1];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      v.B[2 * q] = X0[2 + q];
      v.B[2 * q + 1] = X1[2 + q];
      v.C[3 * q] = X2[2 + q];
      v.C[3 * q + 1] = X3[2 + q];
      v.C[3 * q + 2] = X4[2 + q];
      v.oz[q] = O[2 + q];
    }
    v.om[0] = O[0];
    v.om[1] = O[1];
  }
  v = adj_cta_inclusive_scan(v, warp_tot);
  if (g < T) adj_store(gpref + (bch * T + g) * kAdjDoubles, v);
  if (threadIdx.x == kThreadsGp - 1) adj_store(gcta + (bch * gridDim.x + blockIdx.x) * kAdjDoubles, v);
}

// lambda entering each CTA (from the right): exclusive scan of the CTA totals applied to lambda = 0
__global__ void __launch_bounds__(kThreadsGp) gp_grad_cta_scan_kernel(const double *__restrict__ gcta, int nctas,
                                                                      double *__restrict__ gstate) {
  __shared__ Adj warp_tot[kWarpsGp];
  __shared__ Adj lastlane[kWarpsGp];
  __shared__ Adj carry_s;
  const int64_t bch = blockIdx.x;
  Adj carry = adj_identity();
  for (int base = 0; base < nctas; base += kThreadsGp) {
    const int idx = base + threadIdx.x;
    Adj v = idx < nctas ? adj_load(gcta + (bch * nctas + idx) * kAdjDoubles) : adj_identity();
    v = adj_cta_inclusive_scan(v, warp_tot);
    Adj incl = adj_after(carry, v);
    Adj prev = adj_shfl_up(incl, 1);
    if ((threadIdx.x & 31) == 31) lastlane[threadIdx.x >> 5] = incl;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) prev = threadIdx.x == 0 ? carry : lastlane[(threadIdx.x >> 5) - 1];
    if (idx < nctas) {
      double *o = gstate + (bch * nctas + idx) * 5;
      o[0] = prev.om[0];
      o[1] = prev.om[1];
      o[2] = prev.oz[0];
      o[3] = prev.oz[1];
      o[4] = prev.oz[2];
    }
    if (threadIdx.x == kThreadsGp - 1) carry_s = incl;
    __syncthreads();
    carry = carry_s;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreadsGp, 3)
    gp_grad_final_kernel(const double *__restrict__ t, const double *__restrict__ y, const double *__restrict__ diag, int64_t N,
                         const double *__restrict__ S0, const double *__restrict__ w0, const double *__restrict__ Qf,
                         const double *__restrict__ jitter, double eps, int64_t T, const double *__restrict__ rs,
                         const double *__restrict__ gpref, const double *__restrict__ gstate, double *__restrict__ gy,
                         double *__restrict__ gdiag, double *__restrict__ gpart /* [B][nctas][6] */, unsigned *__restrict__ done,
                         GpStatus *__restrict__ status, double *__restrict__ ghyper /* [B][3] */,
                         double *__restrict__ gsum /* [B][2] or null */) {
  extern __shared__ __align__(16) unsigned char gp_smem_raw[];
  GradSmem &sm = *reinterpret_cast<GradSmem *>(gp_smem_raw);
  RunStates &R = sm.R;
  __shared__ double red[6][kWarpsGp];
  __shared__ int is_last;
  const int64_t bch = blockIdx.y;
  const Sho k = make_sho(S0[bch], w0[bch], Qf[bch], eps);
  const double jit = jitter ? jitter[bch] : 0.0;
  const int64_t g = (int64_t)blockIdx.x * kThreadsGp + threadIdx.x;
  const int64_t run = T - 1 - g;
  const GradTile gt = grad_tile(T);
  stage_tile<kRunB>(sm.tile, t, diag, y + bch * N, jit, gt.base, N);
  double ga = 0.0, gc = 0.0, gh = 0.0, gw2 = 0.0;  // cotangents of a, c, h and w2 = c^2 +- h^2
  double sy = 0.0, sd = 0.0;                        // sums of d/dy and d/d diag over the run
  if (run >= 0) {
    const int64_t s = run * kRunB, e = min(s + kRunB, N);
    const int row = (int)(run - (gt.r_lo > 0 ? gt.r_lo : 0)), len = (int)(e - s);
    gp_rederive(k, sm.tile, row, rs + (size_t)bch * 5 * T + run, T, len, R);
    const double *st = gstate + (bch * gridDim.x + blockIdx.x) * 5;
    double Lm[5] = {st[0], st[1], st[2], st[3], st[4]};
    if (threadIdx.x > 0) adj_apply(adj_load(gpref + (bch * T + g - 1) * kAdjDoubles), Lm);
    ShoCache cc;
#pragma unroll 1
    for (int j = len - 1; j >= 0; --j) {
      const StepFwd f = gp_step_at(k, sm.tile, row, j, R, cc);
      double acc[7];
      gp_step_adj<true, true>(f, Lm, acc);
      sm.tile.y[row * (kRunB + 1) + j] = acc[0];  // d / d y_i     (y_i, diag_i are dead from here on: every step reads its own only)
      sm.tile.d[row * (kRunB + 1) + j] = acc[1];  // d / d diag_i
      sy += acc[0];
      sd += acc[1];
      ga += acc[2];
      // Phi = [[Cf + c Sf, Sf], [-w2 Sf, Cf - c Sf]]
      const double d0011 = acc[3] - acc[6];
      const double Cb = acc[3] + acc[6];
      const double Sb = k.c * d0011 + acc[4] - k.w2 * acc[5];
      gc += f.Sf * d0011 - f.dt * (f.Cf * Cb + f.Sf * Sb);
      gw2 -= f.Sf * acc[5];
      const double dCdh = (k.under ? -f.dt : f.dt) * k.h * f.Sf;
      gh += dCdh * Cb + (f.dt * f.Cf - f.Sf) / k.h * Sb;
    }
    if (!isfinite(ga + gc + gh + gw2)) atomicAdd(&status->bad_lanes, 1ull);
  }
  double vals[6] = {ga, gc, gh, gw2, sy, sd};
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double x = vals[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFullGp, x, o);
    if ((threadIdx.x & 31) == 0) red[j][threadIdx.x >> 5] = x;
  }
  __syncthreads();
  // per-cadence gradients out of the tile, coalesced
  if (gy || gdiag) {
    // only the rows this CTA owns: the CTA holding the start of the series (r_lo < 0) staged rows of its neighbour too
    const int nrows = gt.r_lo < 0 ? kThreadsGp + (int)gt.r_lo : kThreadsGp;
    const int nvalid = (int)max((int64_t)0, min((int64_t)nrows * kRunB, N - gt.base));
#pragma unroll 4
    for (int idx = threadIdx.x; idx < kThreadsGp * kRunB; idx += kThreadsGp) {
      if (idx < nvalid) {
        const int at = (idx / kRunB) * (kRunB + 1) + idx % kRunB;
        if (gy) gy[bch * N + gt.base + idx] = sm.tile.y[at];
        if (gdiag) gdiag[bch * N + gt.base + idx] = sm.tile.d[at];
      }
    }
  }
  if (threadIdx.x < 6) {
    double s4 = 0.0;
    for (int w = 0; w < kWarpsGp; ++w) s4 += red[threadIdx.x][w];
    gpart[(bch * gridDim.x + blockIdx.x) * 6 + threadIdx.x] = s4;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    is_last = atomicAdd(&done[bch], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last && threadIdx.x < 32) {
    // fold the per-CTA cotangents of (a, c, h, w2) in CTA order and chain them to (S0, w0, Q):
    //   a = S0 w0 Q,  c = w0 / (2Q),  h = c f,  f = sqrt(max(+-(4Q^2 - 1), eps)),  w2 = c^2 +- h^2
    __threadfence();
    const int lane = threadIdx.x;
    double fa = 0.0, fc = 0.0, fh = 0.0, fw2 = 0.0, fy = 0.0, fd = 0.0;
    for (int i = lane; i < (int)gridDim.x; i += 32) {
      const double *p = gpart + (bch * gridDim.x + i) * 6;
      fa += __ldcg(p);
      fc += __ldcg(p + 1);
      fh += __ldcg(p + 2);
      fw2 += __ldcg(p + 3);
      fy += __ldcg(p + 4);
      fd += __ldcg(p + 5);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      fa += __shfl_xor_sync(kFullGp, fa, o);
      fc += __shfl_xor_sync(kFullGp, fc, o);
      fh += __shfl_xor_sync(kFullGp, fh, o);
      fw2 += __shfl_xor_sync(kFullGp, fw2, o);
      fy += __shfl_xor_sync(kFullGp, fy, o);
      fd += __shfl_xor_sync(kFullGp, fd, o);
    }
    if (lane == 0) {
      if (gsum) {
        gsum[2 * bch] = fy;
        gsum[2 * bch + 1] = fd;
      }
      const double sgn = k.under ? 1.0 : -1.0;
      fc += 2.0 * k.c * fw2;
      fh += sgn * 2.0 * k.h * fw2;
      const double s0 = S0[bch], w = w0[bch], Q = Qf[bch];
      const double arg = sgn * (4.0 * Q * Q - 1.0);
      const double f = sqrt(fmax(arg, eps));
      const double dfdQ = arg > eps ? sgn * 4.0 * Q / f : 0.0;
      const double dcdQ = -k.c / Q;
      ghyper[3 * bch + 0] = fa * w * Q;
      ghyper[3 * bch + 1] = fa * s0 * Q + fc * (0.5 / Q) + fh * (0.5 / Q) * f;
      ghyper[3 * bch + 2] = fa * s0 * w + fc * dcdQ + fh * (dcdQ * f + k.c * dfdQ);
      done[bch] = 0u;
    }
  }
}

struct GpLayout {
  int run;         // cadences per thread of the forward pass (16; XO_GP_RUN=32 selects XO_GP_RUN_LONG)
  int64_t T;       // runs (threads) per chain
  int nctas;       // CTAs per chain
  int64_t Tb;      // the same for the reverse passes (runs of kRunB)
  int nctas_b;
  size_t off_done, off_pref, off_tot, off_state, off_ll, off_rec, off_bpref, off_bcta, off_bstate, off_rs, off_gpref, off_gcta, off_gstate,
      off_gpart, bytes;
};
inline int gp_run_length(int64_t N, int64_t B) {
  // the long run halves the in-CTA scan work per cadence but also the number of threads
  static const int forced = [] {
    const char *v = getenv("XO_GP_RUN");
    return v && *v ? atoi(v) : 0;
  }();
  (void)N;
  (void)B;
  // with the tiles staged in shared memory the long run costs occupancy (104 KB per CTA): kept as a build / run-time knob
  return forced == XO_GP_RUN_LONG ? XO_GP_RUN_LONG : kRun;
}
inline GpLayout gp_layout(int64_t N, int64_t B, bool solve = false, bool grad = false) {
  GpLayout L;
  L.run = gp_run_length(N, B);
  // sized for the short run (the larger scratch), so that a workspace query does not depend on the run picked later
  const int64_t Tmax = (N + kRun - 1) / kRun;
  const int nctas_max = (int)((Tmax + kThreadsGp - 1) / kThreadsGp);
  L.T = (N + L.run - 1) / L.run;
  L.nctas = (int)((L.T + kThreadsGp - 1) / kThreadsGp);
  L.Tb = (N + kRunB - 1) / kRunB;
  L.nctas_b = (int)((L.Tb + kThreadsGp - 1) / kThreadsGp);
  size_t o = 256;  // GpStatus
  auto take_bytes = [&](size_t n) {
    size_t at = o;
    o += (n + 255) & ~(size_t)255;
    return at;
  };
  auto take = [&](size_t n) { return take_bytes(n * sizeof(double)); };
  L.off_done = take_bytes((size_t)B * sizeof(unsigned));
  L.off_pref = take((size_t)B * Tmax * kAggDoubles);
  L.off_tot = take((size_t)B * nctas_max * kAggDoubles);
  L.off_state = take((size_t)B * nctas_max * 5);
  L.off_ll = take((size_t)B * nctas_max);
  L.off_rec = L.off_bpref = L.off_bcta = L.off_bstate = 0;
  if (solve) {
    L.off_rec = take((size_t)B * 3 * N);
    L.off_bpref = take((size_t)B * L.Tb * 6);
    L.off_bcta = take((size_t)B * L.nctas_b * 6);
    L.off_bstate = take((size_t)B * L.nctas_b * 2);
  }
  L.off_rs = L.off_gpref = L.off_gcta = L.off_gstate = L.off_gpart = 0;
  if (grad) {
    L.off_rs = take((size_t)B * 5 * L.Tb);
    L.off_gpref = take((size_t)B * L.Tb * kAdjDoubles);
    L.off_gcta = take((size_t)B * L.nctas_b * kAdjDoubles);
    L.off_gstate = take((size_t)B * L.nctas_b * 5);
    L.off_gpart = take((size_t)B * L.nctas_b * 6);
  }
  L.bytes = o;
  return L;
}

// the forward scan: three stream-ordered kernels (the last one folds the per-CTA sums)
template <int REC>
int gp_forward(const double *t, const double *y, const double *diag, int64_t N, const double *S0, const double *w0, const double *Q,
               const double *jitter, int64_t B, double eps, double *loglike, double *rec, char *ws, const GpLayout &L, cudaStream_t st) {
  auto at = [&](size_t off) { return reinterpret_cast<double *>(ws + off); };
  unsigned *done = reinterpret_cast<unsigned *>(ws + L.off_done);
  GpStatus *status = reinterpret_cast<GpStatus *>(ws);
  dim3 grid(L.nctas, (unsigned)B);
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {  // tiles above 48 KB of shared memory need the opt-in, once per device
    cudaError_t e = cudaFuncSetAttribute(gp_aggregate_kernel<kRun>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem<kRun>));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gp_loglike_kernel<kRun, REC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem<kRun>));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gp_aggregate_kernel<XO_GP_RUN_LONG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)sizeof(TileSmem<XO_GP_RUN_LONG>));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gp_loglike_kernel<XO_GP_RUN_LONG, REC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)sizeof(TileSmem<XO_GP_RUN_LONG>));
    if (e != cudaSuccess) return -(int)e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  if (L.run == kRun) {
    xo::launch(gp_aggregate_kernel<kRun>, grid, dim3(kThreadsGp), sizeof(TileSmem<kRun>), st, t, y, diag, N, S0, w0, Q, jitter, eps, L.T, at(L.off_pref),
               at(L.off_tot), done, status);
    xo::launch(gp_cta_scan_kernel, dim3((unsigned)B), dim3(kThreadsGp), 0, st, (const double *)at(L.off_tot), L.nctas, S0, w0, Q, eps,
               at(L.off_state));
    xo::launch(gp_loglike_kernel<kRun, REC>, grid, dim3(kThreadsGp), sizeof(TileSmem<kRun>), st, t, y, diag, N, S0, w0, Q, jitter, eps, L.T,
               (const double *)at(L.off_pref), (const double *)at(L.off_state), at(L.off_ll), done, status, loglike, rec, L.Tb);
  } else {
    xo::launch(gp_aggregate_kernel<XO_GP_RUN_LONG>, grid, dim3(kThreadsGp), sizeof(TileSmem<XO_GP_RUN_LONG>), st, t, y, diag, N, S0, w0, Q, jitter, eps, L.T, at(L.off_pref),
               at(L.off_tot), done, status);
    xo::launch(gp_cta_scan_kernel, dim3((unsigned)B), dim3(kThreadsGp), 0, st, (const double *)at(L.off_tot), L.nctas, S0, w0, Q, eps,
               at(L.off_state));
    xo::launch(gp_loglike_kernel<XO_GP_RUN_LONG, REC>, grid, dim3(kThreadsGp), sizeof(TileSmem<XO_GP_RUN_LONG>), st, t, y, diag, N, S0, w0, Q, jitter, eps, L.T,
               (const double *)at(L.off_pref), (const double *)at(L.off_state), at(L.off_ll), done, status, loglike, rec, L.Tb);
  }
  return 0;
}

}  // namespace

namespace xo {

// xo_celerite_sho_grad plus a per-chain jitter term added to diag and the per-chain sums gsum[B][2] =
// (sum_n d/dy_n, sum_n d/d diag_n): what the joint transit + GP model needs (d/d mean = -gsum[0],
// d/d jitter = gsum[1]).  Six kernels, all on `stream`.
int celerite_sho_grad_impl(const double *t, const double *y, const double *diag, int64_t N, const double *S0, const double *w0,
                           const double *Q, const double *jitter, int64_t B, double eps, double *loglike, double *ghyper,
                           double *gy, double *gdiag, double *gsum, void *workspace, size_t workspace_bytes, void *stream) {
  if (N < 0 || B < 0) return XO_EINVAL;
  if (B == 0) return 0;
  if (N == 0) return XO_EINVAL;
  if (B > 65535) return XO_ERANGE;
  if (!t || !y || !diag || !S0 || !w0 || !Q || !loglike || !ghyper || !workspace) return XO_EINVAL;
  GpLayout L = gp_layout(N, B, false, true);
  if (workspace_bytes < L.bytes) return XO_EWORKSPACE;
  char *ws = static_cast<char *>(workspace);
  auto at = [&](size_t off) { return reinterpret_cast<double *>(ws + off); };
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  XO_NVTX_SCOPE("xo_celerite_sho_grad");
  int rc = gp_forward<2>(t, y, diag, N, S0, w0, Q, jitter, B, eps, loglike, at(L.off_rs), ws, L, st);
  if (rc) return rc;
  dim3 gridb(L.nctas_b, (unsigned)B);
  static bool gattr[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !gattr[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gp_grad_aggregate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GradSmem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gp_grad_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GradSmem));
    if (e != cudaSuccess) return -(int)e;
    if (dev >= 0 && dev < 64) gattr[dev] = true;
  }
  xo::launch(gp_grad_aggregate_kernel, gridb, dim3(kThreadsGp), sizeof(GradSmem), st, t, y, diag, N, S0, w0, Q, jitter, eps, L.Tb,
             (const double *)at(L.off_rs), at(L.off_gpref), at(L.off_gcta));
  xo::launch(gp_grad_cta_scan_kernel, dim3((unsigned)B), dim3(kThreadsGp), 0, st, (const double *)at(L.off_gcta), L.nctas_b, at(L.off_gstate));
  xo::launch(gp_grad_final_kernel, gridb, dim3(kThreadsGp), sizeof(GradSmem), st, t, y, diag, N, S0, w0, Q, jitter, eps, L.Tb, (const double *)at(L.off_rs),
             (const double *)at(L.off_gpref), (const double *)at(L.off_gstate), gy, gdiag, at(L.off_gpart),
             reinterpret_cast<unsigned *>(ws + L.off_done), reinterpret_cast<GpStatus *>(ws), ghyper, gsum);
  return xo::launch_status();
}

}  // namespace xo

extern "C" {

size_t xo_celerite_sho_workspace(int64_t N, int64_t B) {
  if (N <= 0 || B <= 0) return 0;
  return gp_layout(N, B).bytes;
}

int xo_celerite_sho_loglike(const double *t, const double *y, const double *diag, int64_t N, const double *S0,
                            const double *w0, const double *Q, int64_t B, double eps, double *loglike, void *workspace,
                            size_t workspace_bytes, void *stream) {
  if (N < 0 || B < 0) return XO_EINVAL;
  if (B == 0) return 0;
  if (N == 0) return XO_EINVAL;
  if (false) return XO_ERANGE;
  if (!t || !y || !diag || !S0 || !w0 || !Q || !loglike || !workspace) return XO_EINVAL;
  XO_NVTX_SCOPE("xo_celerite_sho_loglike");
  GpLayout L = gp_layout(N, B);
  if (workspace_bytes < L.bytes) return XO_EWORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = gp_forward<0>(t, y, diag, N, S0, w0, Q, nullptr, B, eps, loglike, nullptr, static_cast<char *>(workspace), L, st);
  return rc ? rc : xo::launch_status();
}

size_t xo_celerite_sho_solve_workspace(int64_t N, int64_t B) {
  if (N <= 0 || B <= 0) return 0;
  return gp_layout(N, B, true).bytes;
}

// log-likelihood AND alpha = K^-1 y (so that d loglike / d y = -alpha): forward scan, then the
// disturbance smoother as a reverse affine scan.  alpha[B][N].
int xo_celerite_sho_solve(const double *t, const double *y, const double *diag, int64_t N, const double *S0,
                          const double *w0, const double *Q, int64_t B, double eps, double *loglike, double *alpha,
                          void *workspace, size_t workspace_bytes, void *stream) {
  if (N < 0 || B < 0) return XO_EINVAL;
  if (B == 0) return 0;
  if (N == 0) return XO_EINVAL;
  if (B > 65535 || false) return XO_ERANGE;
  if (!t || !y || !diag || !S0 || !w0 || !Q || !loglike || !alpha || !workspace) return XO_EINVAL;
  GpLayout L = gp_layout(N, B, true);
  if (workspace_bytes < L.bytes) return XO_EWORKSPACE;
  char *ws = static_cast<char *>(workspace);
  auto at = [&](size_t off) { return reinterpret_cast<double *>(ws + off); };
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  XO_NVTX_SCOPE("xo_celerite_sho_solve");
  int rc = gp_forward<1>(t, y, diag, N, S0, w0, Q, nullptr, B, eps, loglike, at(L.off_rec), ws, L, st);
  if (rc) return rc;
  dim3 gridb(L.nctas_b, (unsigned)B);
  xo::launch(gp_bwd_aggregate_kernel, gridb, kThreadsGp, 0, st, t, N, S0, w0, Q, eps, L.Tb, at(L.off_rec), at(L.off_bpref),
                                                       at(L.off_bcta));
  xo::launch(gp_bwd_cta_scan_kernel, (unsigned)B, kThreadsGp, 0, st, at(L.off_bcta), L.nctas_b, at(L.off_bstate));
  xo::launch(gp_bwd_solve_kernel, gridb, kThreadsGp, 0, st, t, N, S0, w0, Q, eps, L.Tb, at(L.off_rec), at(L.off_bpref),
                                                   at(L.off_bstate), alpha);
  return xo::launch_status();
}

size_t xo_celerite_sho_grad_workspace(int64_t N, int64_t B) {
  if (N <= 0 || B <= 0) return 0;
  return gp_layout(N, B, false, true).bytes;
}

// log-likelihood and its exact reverse-mode gradient: ghyper[B][3] = d/d(S0, w0, Q), and (each
// optional) gy[B][N] = d/dy (= -K^-1 y), gdiag[B][N] = d/d diag.
int xo_celerite_sho_grad(const double *t, const double *y, const double *diag, int64_t N, const double *S0,
                         const double *w0, const double *Q, int64_t B, double eps, double *loglike, double *ghyper,
                         double *gy, double *gdiag, void *workspace, size_t workspace_bytes, void *stream) {
  return xo::celerite_sho_grad_impl(t, y, diag, N, S0, w0, Q, nullptr, B, eps, loglike, ghyper, gy, gdiag, nullptr, workspace,
                                    workspace_bytes, stream);
}

}  // extern "C"
