// xo_sampler.cu -- the transit posterior in unconstrained coordinates and a whole lock-step HMC transition as C-ABI
// entry points: what turns model+gradient evaluations per second into samples per second.
//
// What it replaces: the PyMC model of /root/reference/docs/tutorials/data-and-models.md:458-509 (priors on log period,
// t0, impact parameter, log radius ratio, Kipping's (q1, q2) limb-darkening parameterisation
// /root/reference/src/exoplanet/distributions.py:98-102, a mean flux) sampled by NUTS one chain per OS process
// (docs/user/multiprocessing.rst).  Here every chain of the batch advances together on the device; the change of
// variables, its chain rule and the priors are two tiny kernels around the fused model evaluation (no torch op,
// no autograd graph), and one call runs all L leapfrog steps of a transition.
//
// Unconstrained vector q[C][D], D = 6 P + 3:
//   per planet  log period, t0, sqrt(e) cos w, sqrt(e) sin w, b, log(r/R*)
//   per chain   logit q1, logit q2 (u1 = 2 sqrt(q1) q2, u2 = sqrt(q1)(1 - 2 q2)), mean flux
// with independent Normal priors on q: prior_mu / prior_sigma are [D] (prior_chain_stride = 0, shared by the chains) or
// [C][prior_chain_stride] (one row per chain).  R* = M* = 1, m_planet = 0.
#include <cuda_runtime.h>
#include <curand_kernel.h>

#include "xo_common.cuh"
#include "xo_posterior.cuh"
#include "xo_transit.cuh"

namespace {

using xo::chain_rule;
using xo::kNpl;
using xo::unpack_planet;
using xo::unpack_star;
constexpr unsigned kFullS = 0xffffffffu;

__device__ __forceinline__ double warp_sum_s(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullS, v, o);
  return v;
}

__global__ void posterior_unpack_kernel(const double *__restrict__ q, int64_t C, int P, int D, double *__restrict__ par,
                                        double *__restrict__ star, double *__restrict__ mean) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C * P) {
    const int64_t c = i / P;
    const int p = (int)(i % P);
    unpack_planet(q + c * D + p * kNpl, par + i * XO_NPAR);
  } else if (i < C * P + C) {
    const int64_t c = i - C * P;
    unpack_star(q + c * D + kNpl * P, star + c * XO_NSTAR, mean + c);
  }
}

// one warp per chain: logp = loglike + log prior, grad = chain rule + prior
__global__ void posterior_finish_kernel(const double *__restrict__ q, const double *__restrict__ gpar, const double *__restrict__ gstar,
                                        const double *__restrict__ gmean, const double *__restrict__ loglike, int64_t C, int P,
                                        int D, const double *__restrict__ mu, const double *__restrict__ sigma, int64_t pstride,
                                        double *__restrict__ logp, double *__restrict__ grad) {
  const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  const int lane = threadIdx.x & 31;
  const double lpr = warp_sum_s(chain_rule(q + c * D, gpar + c * P * XO_NPAR, gstar + c * XO_NSTAR, gmean[c], P, D, mu + c * pstride,
                                           sigma + c * pstride, lane, grad + c * D));
  if (lane == 0) logp[c] = loglike[c] + lpr;
}

// ---- HMC ---------------------------------------------------------------------------------------------------------
// begin: p ~ N(0, M), M = 1 / inv_mass; h0; half kick with the current gradient; first drift; unpack
__global__ void hmc_begin_kernel(const double *__restrict__ q, const double *__restrict__ logp, const double *__restrict__ grad,
                                 const double *__restrict__ eps, const double *__restrict__ inv_mass, int64_t C, int P, int D,
                                 unsigned long long seed, unsigned long long counter, double *__restrict__ qn,
                                 double *__restrict__ mom, double *__restrict__ h0, double *__restrict__ par,
                                 double *__restrict__ star, double *__restrict__ mean) {
  const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  const int lane = threadIdx.x & 31;
  curandStatePhilox4_32_10_t rng;
  curand_init(seed, (unsigned long long)(c * 32 + lane), counter * 64ull, &rng);
  const double e = eps[c];
  double kin = 0.0;
  for (int j = lane; j < D; j += 32) {
    const double im = inv_mass[c * D + j];
    double pj = curand_normal_double(&rng) * rsqrt(im);
    kin += 0.5 * pj * pj * im;
    pj = fma(0.5 * e, grad[c * D + j], pj);
    mom[c * D + j] = pj;
    qn[c * D + j] = fma(e * im, pj, q[c * D + j]);
  }
  kin = warp_sum_s(kin);
  if (lane == 0) h0[c] = -logp[c] + kin;
  __syncwarp();
  if (lane < P) unpack_planet(qn + c * D + lane * kNpl, par + (c * P + lane) * XO_NPAR);
  if (lane == 31) unpack_star(qn + c * D + kNpl * P, star + c * XO_NSTAR, mean + c);
}

// after the model evaluation at qn: logp_n, grad_n; kick (full, or half on the last step); drift + unpack unless last
__global__ void hmc_leap_kernel(double *__restrict__ qn, const double *__restrict__ gpar, const double *__restrict__ gstar,
                                const double *__restrict__ gmean, const double *__restrict__ loglike,
                                const double *__restrict__ eps, const double *__restrict__ inv_mass, int64_t C, int P, int D,
                                const double *__restrict__ mu, const double *__restrict__ sigma, int64_t pstride, int is_last,
                                double *__restrict__ mom, double *__restrict__ logp_n, double *__restrict__ grad_n,
                                double *__restrict__ par, double *__restrict__ star, double *__restrict__ mean) {
  const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  const int lane = threadIdx.x & 31;
  const double lpr = warp_sum_s(chain_rule(qn + c * D, gpar + c * P * XO_NPAR, gstar + c * XO_NSTAR, gmean[c], P, D, mu + c * pstride,
                                           sigma + c * pstride, lane, grad_n + c * D));
  if (lane == 0) logp_n[c] = loglike[c] + lpr;
  __syncwarp();
  const double e = eps[c];
  for (int j = lane; j < D; j += 32) {
    const double pj = fma(is_last ? 0.5 * e : e, grad_n[c * D + j], mom[c * D + j]);
    mom[c * D + j] = pj;
    if (!is_last) qn[c * D + j] = fma(e * inv_mass[c * D + j], pj, qn[c * D + j]);
  }
  if (is_last) return;
  __syncwarp();
  if (lane < P) unpack_planet(qn + c * D + lane * kNpl, par + (c * P + lane) * XO_NPAR);
  if (lane == 31) unpack_star(qn + c * D + kNpl * P, star + c * XO_NSTAR, mean + c);
}

// Metropolis accept / reject per chain; non-finite energy -> reject
__global__ void hmc_end_kernel(const double *__restrict__ qn, const double *__restrict__ mom, const double *__restrict__ logp_n,
                               const double *__restrict__ grad_n, const double *__restrict__ h0,
                               const double *__restrict__ inv_mass, int64_t C, int D, unsigned long long seed,
                               unsigned long long counter, double *__restrict__ q, double *__restrict__ logp,
                               double *__restrict__ grad, double *__restrict__ accept_prob, double *__restrict__ energy_error) {
  const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  const int lane = threadIdx.x & 31;
  double kin = 0.0;
  for (int j = lane; j < D; j += 32) {
    const double pj = mom[c * D + j];
    kin += 0.5 * pj * pj * inv_mass[c * D + j];
  }
  kin = warp_sum_s(kin);
  const double h1 = -logp_n[c] + kin;
  double dh = h0[c] - h1;
  if (!isfinite(dh)) dh = -INFINITY;
  const double a = exp(fmin(dh, 0.0));
  curandStatePhilox4_32_10_t rng;
  curand_init(seed, (unsigned long long)(c * 32), counter * 64ull + 32ull, &rng);
  const bool acc = curand_uniform_double(&rng) < a;  // the same draw in every lane
  if (acc) {
    for (int j = lane; j < D; j += 32) {
      q[c * D + j] = qn[c * D + j];
      grad[c * D + j] = grad_n[c * D + j];
    }
    if (lane == 0) logp[c] = logp_n[c];
  }
  if (lane == 0) {
    accept_prob[c] = a;
    energy_error[c] = -dh;
  }
}

struct SamplerLayout {
  size_t off_par, off_star, off_mean, off_gpar, off_gstar, off_gmean, off_ll, off_qn, off_mom, off_h0, off_lpn, off_gn, off_model,
      model_bytes, bytes;
};
inline SamplerLayout sampler_layout(int64_t N, int64_t C, int64_t P) {
  SamplerLayout L;
  const int64_t D = kNpl * P + 3;
  size_t o = 0;
  auto take = [&](size_t n) {
    size_t at = o;
    o += (n * sizeof(double) + 255) & ~(size_t)255;
    return at;
  };
  L.off_par = take((size_t)C * P * XO_NPAR);
  L.off_star = take((size_t)C * XO_NSTAR);
  L.off_mean = take((size_t)C);
  L.off_gpar = take((size_t)C * P * XO_NPAR);
  L.off_gstar = take((size_t)C * XO_NSTAR);
  L.off_gmean = take((size_t)C);
  L.off_ll = take((size_t)C);
  L.off_qn = take((size_t)C * D);
  L.off_mom = take((size_t)C * D);
  L.off_h0 = take((size_t)C);
  L.off_lpn = take((size_t)C);
  L.off_gn = take((size_t)C * D);
  L.model_bytes = xo_transit_model_workspace(N, C, P);
  L.off_model = take((L.model_bytes + 7) / 8);
  L.bytes = o;
  return L;
}

}  // namespace

extern "C" {

size_t xo_transit_posterior_workspace(int64_t N, int64_t C, int64_t P) {
  if (N <= 0 || C <= 0 || P <= 0) return 0;
  return sampler_layout(N, C, P).bytes;
}

int xo_transit_posterior_logp_grad(const double *t, const double *y, const double *ivar, int64_t N, const double *moments,
                                   const double *q, int64_t C, int64_t P, const double *prior_mu, const double *prior_sigma,
                                   int64_t prior_chain_stride, const double *dtk, const double *wk, int K, double texp, int use_in_transit, double lognorm,
                                   double *logp, double *grad, void *workspace, size_t workspace_bytes, void *stream) {
  if (int rc = xo::check_batch(N, C, P)) return rc;
  if (C == 0) return 0;
  if (N == 0 || P == 0) return XO_EINVAL;
  if (!t || !y || !ivar || !q || !prior_mu || !prior_sigma || !dtk || !wk || !logp || !grad || !workspace) return XO_EINVAL;
  if (prior_chain_stride != 0 && prior_chain_stride < 6 * P + 3) return XO_EINVAL;
  const SamplerLayout L = sampler_layout(N, C, P);
  if (workspace_bytes < L.bytes) return XO_EWORKSPACE;
  XO_NVTX_SCOPE("xo_transit_posterior_logp_grad");
  char *ws = static_cast<char *>(workspace);
  auto at = [&](size_t off) { return reinterpret_cast<double *>(ws + off); };
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int D = kNpl * (int)P + 3;
  if (xo::model_split_supported()) {
    // the model's own prologue and epilogue kernels apply the change of variables, its chain rule and the priors in place:
    // 3 launches per posterior evaluation instead of 5
    xo::PosteriorIO io;
    io.q = q;
    io.D = D;
    io.mu = prior_mu;
    io.sigma = prior_sigma;
    io.pstride = prior_chain_stride;
    io.par = at(L.off_par);
    io.star = at(L.off_star);
    io.mean = at(L.off_mean);
    io.logp = logp;
    io.grad = grad;
    return xo::model_logp_grad_phases(t, y, ivar, N, at(L.off_par), at(L.off_star), at(L.off_mean), C, P, dtk, wk, K, texp, use_in_transit,
                                      lognorm, moments, at(L.off_ll), at(L.off_gpar), at(L.off_gstar), at(L.off_gmean), ws + L.off_model,
                                      L.model_bytes, stream, 3, &io);
  }
  xo::launch(posterior_unpack_kernel, (unsigned)((C * P + C + 127) / 128), 128, 0, st, q, C, (int)P, D, at(L.off_par), at(L.off_star),
             at(L.off_mean));
  int rc = xo_transit_model_logp_grad(t, y, ivar, N, at(L.off_par), at(L.off_star), at(L.off_mean), C, P, dtk, wk, K, texp,
                                      use_in_transit, lognorm, moments, at(L.off_ll), at(L.off_gpar), at(L.off_gstar), at(L.off_gmean),
                                      ws + L.off_model, L.model_bytes, stream);
  if (rc) return rc;
  xo::launch(posterior_finish_kernel, (unsigned)((C + 3) / 4), 128, 0, st, q, at(L.off_gpar), at(L.off_gstar), at(L.off_gmean),
             at(L.off_ll), C, (int)P, D, prior_mu, prior_sigma, prior_chain_stride, logp, grad);
  return xo::launch_status();
}

int xo_hmc_transition(const double *t, const double *y, const double *ivar, int64_t N, const double *moments, int64_t C, int64_t P,
                      const double *prior_mu, const double *prior_sigma, int64_t prior_chain_stride, const double *dtk, const double *wk,
                      int K, double texp, int use_in_transit, double lognorm, double *q, double *logp, double *grad, const double *eps,
                      const double *inv_mass, int n_leapfrog, unsigned long long seed, unsigned long long counter,
                      double *accept_prob, double *energy_error, void *workspace, size_t workspace_bytes, void *stream) {
  if (int rc = xo::check_batch(N, C, P)) return rc;
  if (C == 0) return 0;
  if (N == 0 || P == 0 || n_leapfrog < 1) return XO_EINVAL;
  if (prior_chain_stride != 0 && prior_chain_stride < 6 * P + 3) return XO_EINVAL;
  if (!t || !y || !ivar || !q || !logp || !grad || !eps || !inv_mass || !prior_mu || !prior_sigma || !dtk || !wk || !accept_prob ||
      !energy_error || !workspace)
    return XO_EINVAL;
  const SamplerLayout L = sampler_layout(N, C, P);
  if (workspace_bytes < L.bytes) return XO_EWORKSPACE;
  XO_NVTX_SCOPE("xo_hmc_transition");
  char *ws = static_cast<char *>(workspace);
  auto at = [&](size_t off) { return reinterpret_cast<double *>(ws + off); };
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int D = kNpl * (int)P + 3;
  const unsigned grid = (unsigned)((C + 3) / 4);
  xo::launch(hmc_begin_kernel, grid, 128, 0, st, (const double *)q, (const double *)logp, (const double *)grad, eps, inv_mass, C, (int)P, D,
             seed, counter, at(L.off_qn), at(L.off_mom), at(L.off_h0), at(L.off_par), at(L.off_star), at(L.off_mean));
  for (int i = 0; i < n_leapfrog; ++i) {
    int rc = xo_transit_model_logp_grad(t, y, ivar, N, at(L.off_par), at(L.off_star), at(L.off_mean), C, P, dtk, wk, K, texp,
                                        use_in_transit, lognorm, moments, at(L.off_ll), at(L.off_gpar), at(L.off_gstar),
                                        at(L.off_gmean), ws + L.off_model, L.model_bytes, stream);
    if (rc) return rc;
    xo::launch(hmc_leap_kernel, grid, 128, 0, st, at(L.off_qn), (const double *)at(L.off_gpar), (const double *)at(L.off_gstar),
               (const double *)at(L.off_gmean), (const double *)at(L.off_ll), eps, inv_mass, C, (int)P, D, prior_mu, prior_sigma,
               prior_chain_stride, (int)(i + 1 == n_leapfrog), at(L.off_mom), at(L.off_lpn), at(L.off_gn), at(L.off_par), at(L.off_star), at(L.off_mean));
  }
  xo::launch(hmc_end_kernel, grid, 128, 0, st, (const double *)at(L.off_qn), (const double *)at(L.off_mom), (const double *)at(L.off_lpn),
             (const double *)at(L.off_gn), (const double *)at(L.off_h0), inv_mass, C, D, seed, counter, q, logp, grad, accept_prob,
             energy_error);
  return xo::launch_status();
}

}  // extern "C"
