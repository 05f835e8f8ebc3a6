// xo_transit.cuh -- host-side pieces of the fused logp+grad launch shared with xo_model.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "xo_common.cuh"

namespace xo {

constexpr int kNBase = 5;  // per-warp partial layout: lp, gmean, gcl[3], then gel[P][XO_NGRAD]

// work split of one chain: slots (warps) and cadences per slot
struct Split {
  int slots;
  int64_t L;
};
Split split_transit(int64_t N, int64_t C);
int check_batch(int64_t N, int64_t C, int64_t P);
int transit_min_blocks();

// single-launch model mode of the logp kernel (parameters in, parameter gradients out)
struct ModelArgs {
  const double *par = nullptr;   // [C][P][XO_NPAR]
  const double *star = nullptr;  // [C][XO_NSTAR]
  double texp = 0.0;
  int use_in_transit = 1;
  int elements_only = 0;      // 1: plain element-level launch that only uses `order`
  int stage_from_params = 1;  // 0: element blocks were written by model_prologue_kernel (elem, cl valid)
  double lognorm = 0.0;
  unsigned *counter = nullptr;  // [C], zero on entry; left zero on exit
  const int *order = nullptr;   // [C] chain processed by blockIdx.y (heaviest first), or null = identity
  double *logp = nullptr, *gpar = nullptr, *gstar = nullptr, *gmean = nullptr;
  // light-curve VJP mode only: the cotangent is -ybar[C][N], one value per (chain, cadence) shared by every planet
  // (d/d lc of a likelihood of the residual y - mean - sum_p lc), and gmean is taken from -vjp_gsum[2 c]
  int vjp_shared = 0;
  const double *vjp_gsum = nullptr;
  // TTVOrbit (/root/reference/src/exoplanet/orbits/ttv.py:175-187): per (chain, planet) the sorted table of transit times
  // ttv_times[C][P][ttv_nt] (all_transit_times; ttv_count[p] valid entries).  A cadence is measured from the NEAREST transit
  // (bin edges half-way between consecutive transits) instead of from elem[XO_EL_T0]; ttv_grad[C][P][ttv_nt] receives
  // d / d transit time (zero-initialised by the caller; accumulated with atomics).
  const double *ttv_times = nullptr;
  const int *ttv_count = nullptr;
  int ttv_nt = 0;
  double *ttv_grad = nullptr;
  // light-curve VJP mode only: light-travel delay (keplerian.py:411-470); the element blocks carry R_star in XO_EL_RSTAR
  int light_delay = 0;
  // per-cadence exposure times texp_vec[N] (limb_dark.py:199-205; element-level launches): dtk then holds the UNIT stencil
  const double *texp_vec = nullptr;
  // work partition by expected in-transit cadences (element-level launches): weights[C][P] (>= 1 each, see
  // transit_weight()), written by the kernel that derived the element blocks; rows[C] receives (first row, rows) of each
  // chain in `partial` for the fold.  null: every chain gets gridDim.x CTAs.
  const uint32_t *weights = nullptr;
  int2 *rows = nullptr;
  int nchains = 0;    // with weights: number of chains (the grid is then one-dimensional, grid_ctas CTAs)
  int grid_ctas = 0;  // with weights: transit_weighted_grid()
};

// Expected drains (32 in-transit cadences each) of one planet over N cadences, + 1: the unit of the work partition
__host__ __device__ inline uint32_t transit_weight(double period, double tstart, double tend, int64_t N) {
  const double d = (tend - tstart) / period;
  const double duty = !(d < 1.0) ? 1.0 : (d > 0.0 ? d : 0.0);  // NaN, infinite window, window >= period: every cadence
  return 1u + (uint32_t)(duty * (double)N * (1.0 / 32.0));
}
constexpr int64_t kMaxWeighted = 4096;  // chains * planets up to which the logp kernel builds the partition table in every CTA
bool transit_weighted(int64_t C, int64_t P);
// CTAs of a weighted launch: whole waves of the device (every CTA slot busy until the end) once the launch exceeds one wave
int64_t transit_weighted_grid(int64_t N, int64_t C);
int64_t transit_weighted_rows(int64_t N, int64_t C);  // rows of `partial` such a launch writes (one per warp)

// light curve forward with an optional transit-time table (lc[C][N][P])
void launch_transit_lc_fwd(const double *t, int64_t N, const double *elem, const double *cl, int64_t C, int P, const double *dtk,
                           const double *wk, int K, const double *ttv_times, const int *ttv_count, int ttv_nt, double *lc,
                           cudaStream_t st, bool light_delay = false, const double *texp_vec = nullptr);

// enqueue transit_logp_grad_kernel; partial[C][slots][kNBase + P*XO_NGRAD]; model != nullptr
// selects the single-launch model mode (elem / cl are then ignored)
void launch_transit_logp_grad(const double *t, const double *y, const double *ivar, int64_t N, const double *elem,
                              const double *cl, const double *mean, int64_t C, int P, const double *dtk,
                              const double *wk, int K, Split sp, const double *moments, double *partial, cudaStream_t st,
                              const ModelArgs *model, bool pdl = false);

// the model evaluation in phases (xo_model.cu): 1 = prologue only, 2 = logp kernel + epilogue, 3 = all.
// A host driver can run phase 1 on a side stream while the light curve is still being uploaded.
bool model_split_supported();
int model_logp_grad_phases(const double *t, const double *y, const double *ivar, int64_t N, const double *par,
                           const double *star, const double *mean, int64_t C, int64_t P, const double *dtk, const double *wk,
                           int K, double texp, int use_in_transit, double lognorm, const double *moments, double *logp,
                           double *gpar, double *gstar, double *gmean, void *workspace, size_t workspace_bytes, void *stream,
                           int phases, const struct PosteriorIO *posterior = nullptr);

// resid[C][N] = y[n] - mean[c] - sum_p lc[c][n][p]   (what a noise model of the residuals consumes)
void launch_transit_resid(const double *t, const double *y, int64_t N, const double *elem, const double *cl, const double *mean,
                          int64_t C, int P, const double *dtk, const double *wk, int K, double *resid, cudaStream_t st,
                          const double *ttv_times = nullptr, const int *ttv_count = nullptr, int ttv_nt = 0);
// fold of the per-warp partials + reverse pass of the prologue (xo_model.cu)
int launch_model_epilogue(const double *par, const double *star, const double *partial, int slots, int64_t C, int P, double lognorm,
                          double *logp, double *gpar, double *gstar, double *gmean, cudaStream_t st, bool pdl, const int2 *rows = nullptr);

// celerite gradient with per-chain jitter and per-chain sums of d/dy, d/d diag (xo_celerite.cu)
int celerite_sho_grad_impl(const double *t, const double *y, const double *diag, int64_t N, const double *S0, const double *w0,
                           const double *Q, const double *jitter, int64_t B, double eps, double *loglike, double *ghyper,
                           double *gy, double *gdiag, double *gsum, void *workspace, size_t workspace_bytes, void *stream);

// order[rank] = chain, heaviest expected in-transit work first
void launch_chain_order(const double *elem, int64_t C, int P, int *order, cudaStream_t st);

}  // namespace xo
