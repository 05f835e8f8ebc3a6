// xo_posterior.cuh -- the change of variables of the transit posterior (csrc/xo_sampler.cu): unconstrained q -> sampler-side
// parameters, and its chain rule + Normal priors.  Shared by the stand-alone kernels of xo_sampler.cu and by the model
// prologue / epilogue kernels of xo_model.cu, which apply it in place when a posterior evaluation asks them to (two
// launches fewer per evaluation).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/xo_b200.h"

namespace xo {

// what a posterior evaluation hands to the model kernels (q == nullptr: plain model evaluation)
struct PosteriorIO {
  const double *q = nullptr;  // [C][D], D = 6 P + 3
  int D = 0;
  const double *mu = nullptr, *sigma = nullptr;  // priors: [D] (pstride = 0) or [C][pstride]
  int64_t pstride = 0;
  double *par = nullptr, *star = nullptr, *mean = nullptr;  // written by the prologue (the logp kernel and the fold read them)
  double *logp = nullptr, *grad = nullptr;                  // [C], [C][D]: written by the epilogue
};

#ifdef __CUDACC__
constexpr int kNpl = 6;

// q[c] -> par[c][P][8], star[c][4], mean[c]; one thread per (chain, planet) plus one per chain
__device__ __forceinline__ void unpack_planet(const double *q, double *par) {
  const double h = q[2], k = q[3];
  par[0] = exp(q[0]);
  par[1] = q[1];
  par[2] = h * h + k * k;
  par[3] = atan2(k, h);
  par[4] = q[4];
  par[5] = exp(q[5]);
  par[6] = 0.0;
  par[7] = 0.0;
}
__device__ __forceinline__ void unpack_star(const double *qs, double *star, double *mean) {
  const double q1 = 1.0 / (1.0 + exp(-qs[0])), q2 = 1.0 / (1.0 + exp(-qs[1]));
  const double sq = sqrt(q1);
  star[0] = 1.0;
  star[1] = 1.0;
  star[2] = 2.0 * sq * q2;
  star[3] = sq * (1.0 - 2.0 * q2);
  *mean = qs[2];
}

// chain rule of the change of variables + the Normal priors; lane-parallel over the D coordinates of one chain.
// Returns this lane's share of the log prior (to be summed over the warp).
__device__ __forceinline__ double chain_rule(const double *q, const double *gpar, const double *gstar, double gmean, int P, int D,
                                             const double *mu, const double *sigma, int lane, double *grad) {
  double lprior = 0.0;
  for (int j = lane; j < D; j += 32) {
    double g;
    if (j < kNpl * P) {
      const int p = j / kNpl, i = j % kNpl;
      const double *qp = q + p * kNpl, *gp = gpar + p * XO_NPAR;
      const double h = qp[2], k = qp[3], e = h * h + k * k;
      switch (i) {
        case 0: g = gp[0] * exp(qp[0]); break;
        case 1: g = gp[1]; break;
        case 2: g = gp[2] * 2.0 * h + (e > 0.0 ? gp[3] * (-k / e) : 0.0); break;
        case 3: g = gp[2] * 2.0 * k + (e > 0.0 ? gp[3] * (h / e) : 0.0); break;
        case 4: g = gp[4]; break;
        default: g = gp[5] * exp(qp[5]); break;
      }
    } else {
      const double *qs = q + kNpl * P;
      const int i = j - kNpl * P;
      if (i == 2) {
        g = gmean;
      } else {
        const double q1 = 1.0 / (1.0 + exp(-qs[0])), q2 = 1.0 / (1.0 + exp(-qs[1]));
        const double sq = sqrt(q1);
        if (i == 0) {
          const double gq1 = gstar[2] * q2 / sq + gstar[3] * (1.0 - 2.0 * q2) / (2.0 * sq);
          g = gq1 * q1 * (1.0 - q1);
        } else {
          const double gq2 = 2.0 * sq * (gstar[2] - gstar[3]);
          g = gq2 * q2 * (1.0 - q2);
        }
      }
    }
    const double z = (q[j] - mu[j]) / sigma[j];
    lprior -= 0.5 * z * z;
    grad[j] = g - z / sigma[j];
  }
  return lprior;
}

#endif  // __CUDACC__

}  // namespace xo
