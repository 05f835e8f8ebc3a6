// xo_model.cu -- the whole "model + gradient" evaluation from sampler-side parameters:
//   prologue   (period, t0, e, omega, b, r/R*, m_pl, K; R*, M*, u1, u2)  ->  element block + c
//              = KeplerianOrbit.__init__ (/root/reference/src/exoplanet/orbits/keplerian.py:133-281),
//                the in_transit window (:729-767, contact_points op :744-753) and
//                get_cl (/root/reference/src/exoplanet/light_curves/limb_dark.py:11-18)
//   main       transit_logp_grad_kernel (xo_transit.cu)
//   epilogue   fold of the per-warp partials + the prologue's reverse pass, so one C call returns
//              d logp / d parameters -- what PyMC's NUTS asks for at every leapfrog step.
// Three launches, no host round trip, O(#params) bytes in and out per chain.
#include <cuda_runtime.h>

#include <cstdlib>

#include "xo_common.cuh"
#include "xo_posterior.cuh"
#include "xo_prologue.cuh"

namespace {

// pio.q != nullptr (a posterior evaluation, csrc/xo_sampler.cu): the sampler-side parameters are derived here from the
// unconstrained q instead of being read -- and written out for the fold; the pair of a planet works from its own local copy
template <bool kPost>
__global__ void model_prologue_kernel(const double *__restrict__ par, const double *__restrict__ star, int64_t C, int P,
                                      double texp, int use_in_transit, double *__restrict__ elem,
                                      double *__restrict__ cl, int64_t N, uint32_t *__restrict__ weights, uint32_t base_weight,
                                      const xo::PosteriorIO pio) {
  xo::pdl_launch_dependents();  // the logp kernel may start its data-only preamble now
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t idx = tid >> 1;
  const bool active = idx < C * P;
  const int64_t ii = active ? idx : 0;
  const int64_t c = ii / P;
  double lpar[XO_NPAR], lstar[XO_NSTAR];
  const double *ppar = par + ii * XO_NPAR, *pstar = star + c * XO_NSTAR;
  if (kPost) {  // (a template flag: the plain evaluation keeps its global-memory operands and its instruction stream)
    const int p = (int)(ii % P);
    const double *qc = pio.q + c * pio.D;
    double lmean;
    xo::unpack_planet(qc + p * xo::kNpl, lpar);
    xo::unpack_star(qc + xo::kNpl * P, lstar, &lmean);
    if (active && (tid & 1) == 0) {
#pragma unroll
      for (int j = 0; j < XO_NPAR; ++j) pio.par[ii * XO_NPAR + j] = lpar[j];
      if (p == 0) {
#pragma unroll
        for (int j = 0; j < XO_NSTAR; ++j) pio.star[c * XO_NSTAR + j] = lstar[j];
        pio.mean[c] = lmean;
      }
    }
    ppar = lpar;
    pstar = lstar;
  }
  xo::prologue_pair(ppar, pstar, texp, use_in_transit, (int)(tid & 1), active, elem + ii * XO_NEL);
  if (active && (tid & 1) == 0 && idx % P == 0) xo::limb_coeffs(pstar[2], pstar[3], cl + 3 * c);
  if (weights) {  // expected in-transit work of this (chain, planet): the logp kernel deals its CTAs out by these
    __syncwarp();  // the pair's element block is complete
    if (active && (tid & 1) == 0) {
      const double *el = elem + ii * XO_NEL;
      // base_weight: what sweeping the chain's N cadences costs when the sweep has to read them all (no data moments)
      weights[ii] = xo::transit_weight(el[XO_EL_PERIOD], el[XO_EL_TSTART], el[XO_EL_TEND], N) + (idx % P == 0 ? base_weight : 0u);
    }
  }
}

// one warp per chain.  pio.q != nullptr: the parameters are re-derived from q for the parameter-only half (it runs before
// pdl_wait(), i.e. possibly before the prologue kernel two launches back has written them), and after the fold the chain
// rule of the change of variables and the priors turn (logp, gpar, gstar, gmean) into the posterior's logp and d/dq.
template <bool kPost>
__global__ void model_epilogue_kernel(const double *__restrict__ par, const double *__restrict__ star,
                                      const double *__restrict__ partial, int slots, int64_t C, int P, double lognorm,
                                      double *__restrict__ logp, double *__restrict__ gpar, double *__restrict__ gstar,
                                      double *__restrict__ gmean, const int2 *__restrict__ rows, const xo::PosteriorIO pio) {
  const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  const int nacc = xo::kNBase + P * XO_NGRAD;
  const int lane = threadIdx.x & 31;
  double lpar[XO_NPAR], lstar[XO_NSTAR];
  const double *ppar = par + (c * P + (lane < P ? lane : 0)) * XO_NPAR, *pstar = star + c * XO_NSTAR;
  if (kPost) {
    const double *qc = pio.q + c * pio.D;
    double lmean;
    xo::unpack_planet(qc + (lane < P ? lane : 0) * xo::kNpl, lpar);
    xo::unpack_star(qc + xo::kNpl * P, lstar, &lmean);
    ppar = lpar;
    pstar = lstar;
  }
  // parameter-only half first: as a programmatic dependent of the logp kernel this overlaps its tail
  const xo::EpiPre pre = xo::epilogue_pre(ppar, pstar);
  xo::pdl_wait();
  // the chain's rows of `partial`: slots of them, or what the weighted partition of the logp kernel dealt it
  const int64_t first = rows ? rows[c].x : c * slots;
  const int count = rows ? rows[c].y : slots;
  xo::epilogue_chain<true>(pre, pstar, partial + first * nacc, count, P, lognorm, logp + c, gpar + c * P * XO_NPAR,
                           gstar + c * XO_NSTAR, gmean + c);
  if (kPost) {
    __syncwarp();  // the lanes' gpar / gstar / gmean / logp are visible to the whole warp
    const double *gpar_c = gpar + c * P * XO_NPAR, *gstar_c = gstar + c * XO_NSTAR;
    const double gm = *(const volatile double *)(gmean + c), ll = *(const volatile double *)(logp + c);
    double lpr = xo::chain_rule(pio.q + c * pio.D, gpar_c, gstar_c, gm, P, pio.D, pio.mu + c * pio.pstride, pio.sigma + c * pio.pstride,
                                lane, pio.grad + c * pio.D);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lpr += __shfl_xor_sync(0xffffffffu, lpr, o);
    if (lane == 0) pio.logp[c] = ll + lpr;
  }
}

struct ModelLayout {
  size_t off_elem, off_cl, off_partial, off_counter, off_weights, off_rows, bytes;
  xo::Split sp;
};
inline ModelLayout model_layout(int64_t N, int64_t C, int64_t P) {
  ModelLayout L;
  L.sp = xo::split_transit(N, C);
  size_t o = 0;
  auto take = [&](size_t n) {
    size_t at = o;
    o += (n * sizeof(double) + 255) & ~(size_t)255;
    return at;
  };
  L.off_elem = take((size_t)C * P * XO_NEL);
  L.off_cl = take((size_t)C * 3);
  const size_t rows_uniform = (size_t)C * L.sp.slots, rows_weighted = (size_t)xo::transit_weighted_rows(N, C);
  L.off_partial = take((rows_uniform > rows_weighted ? rows_uniform : rows_weighted) * (xo::kNBase + P * XO_NGRAD));
  L.off_counter = take((size_t)C);  // per-chain arrival counters of the single-launch mode
  L.off_weights = take(((size_t)C * P + 1) / 2);  // uint32 [C][P]
  L.off_rows = take((size_t)C);                   // int2 [C]
  L.bytes = o;
  return L;
}

}  // namespace

namespace xo {

static int model_launches() {
  // XO_MODEL_LAUNCHES: 3 = prologue kernel, logp kernel, epilogue kernel (default: measured fastest);
  //                    2 = prologue kernel + logp kernel whose last CTA per chain runs the epilogue;
  //                    1 = one kernel (every CTA re-derives its chain's elements at start)
  static const int launches = [] {
    const char *v = getenv("XO_MODEL_LAUNCHES");
    return v && *v ? atoi(v) : 3;
  }();
  return launches;
}

bool model_split_supported() { return model_launches() == 3; }

int launch_model_epilogue(const double *par, const double *star, const double *partial, int slots, int64_t C, int P, double lognorm,
                          double *logp, double *gpar, double *gstar, double *gmean, cudaStream_t st, bool pdl, const int2 *rows) {
  cudaError_t le = xo::launch_pdl(model_epilogue_kernel<false>, dim3((unsigned)((C + 3) / 4)), dim3(128), st, pdl, par, star, partial, slots, C, P,
                                  lognorm, logp, gpar, gstar, gmean, rows, xo::PosteriorIO{});
  return le == cudaSuccess ? 0 : -(int)le;
}

// phases: 1 = prologue only (may run on another stream than the rest), 2 = logp kernel + epilogue
// (the caller has ordered them after the prologue), 3 = everything on `stream`
int model_logp_grad_phases(const double *t, const double *y, const double *ivar, int64_t N, const double *par,
                           const double *star, const double *mean, int64_t C, int64_t P, const double *dtk, const double *wk,
                           int K, double texp, int use_in_transit, double lognorm, const double *moments, double *logp,
                           double *gpar, double *gstar, double *gmean, void *workspace, size_t workspace_bytes, void *stream,
                           int phases, const PosteriorIO *posterior) {
  if (int rc = xo::check_batch(N, C, P)) return rc;
  if (K < 1 || K > XO_MAX_STENCIL) return XO_ERANGE;
  if (C == 0) return 0;
  if (N == 0 || P == 0) return XO_EINVAL;
  if (!t || !y || !ivar || !par || !star || !mean || !dtk || !wk || !logp || !gpar || !gstar || !gmean || !workspace)
    return XO_EINVAL;
  ModelLayout L = model_layout(N, C, P);
  if (workspace_bytes < L.bytes) return XO_EWORKSPACE;
  char *ws = static_cast<char *>(workspace);
  double *elem = reinterpret_cast<double *>(ws + L.off_elem);
  double *cl = reinterpret_cast<double *>(ws + L.off_cl);
  double *partial = reinterpret_cast<double *>(ws + L.off_partial);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int launches = model_launches();
  const int64_t cp = C * P;
  if (posterior && (launches != 3 || phases != 3)) return XO_EINVAL;  // the in-place change of variables needs the 3-kernel chain
  const PosteriorIO io = posterior ? *posterior : PosteriorIO{};
  if (launches == 1 || launches == 2) {
    if (phases != 3) return XO_EINVAL;
    unsigned *counter = reinterpret_cast<unsigned *>(ws + L.off_counter);
    cudaError_t e = cudaMemsetAsync(counter, 0, (size_t)C * sizeof(unsigned), st);
    if (e != cudaSuccess) return -(int)e;
    xo::ModelArgs ma;
    ma.par = par;
    ma.star = star;
    ma.texp = texp;
    ma.use_in_transit = use_in_transit;
    ma.stage_from_params = launches == 1;
    ma.lognorm = lognorm;
    ma.counter = counter;
    ma.logp = logp;
    ma.gpar = gpar;
    ma.gstar = gstar;
    ma.gmean = gmean;
    if (launches == 2)
      xo::launch(model_prologue_kernel<false>, (unsigned)((2 * cp + 63) / 64), 64, 0, st, par, star, C, (int)P, texp, use_in_transit, elem, cl, N,
                 (uint32_t *)nullptr, 0u, xo::PosteriorIO{});
    xo::launch_transit_logp_grad(t, y, ivar, N, elem, cl, mean, C, (int)P, dtk, wk, K, L.sp, moments, partial, st, &ma);
    return xo::launch_status();
  }
  static const bool pdl = [] {  // XO_MODEL_PDL=0: plain stream order between the three kernels
    const char *v = getenv("XO_MODEL_PDL");
    return !(v && *v == '0');
  }();
  // CTAs of the logp kernel dealt to the chains by their expected in-transit work (weights from the prologue)
  const bool weighted = xo::transit_weighted(C, P);
  uint32_t *weights = weighted ? reinterpret_cast<uint32_t *>(ws + L.off_weights) : nullptr;
  int2 *rows = weighted ? reinterpret_cast<int2 *>(ws + L.off_rows) : nullptr;
  if (phases & 1) {
    // without the data moments every cadence is read and summed in the sweep: ~40 instructions per 32 cadences against
    // ~1500 per drain of 32 in-transit samples
    const uint32_t base_w = moments ? 0u : (uint32_t)(N / 1200);
    if (io.q)
      xo::launch(model_prologue_kernel<true>, (unsigned)((2 * cp + 63) / 64), 64, 0, st, par, star, C, (int)P, texp, use_in_transit, elem, cl,
                 N, weights, base_w, io);
    else
      xo::launch(model_prologue_kernel<false>, (unsigned)((2 * cp + 63) / 64), 64, 0, st, par, star, C, (int)P, texp, use_in_transit, elem, cl,
                 N, weights, base_w, io);
  }
  if (!(phases & 2)) return xo::launch_status();
  static const int lpt = [] {
    const char *v = getenv("XO_MODEL_HEAVY_FIRST");
    return v && *v ? atoi(v) : 0;  // measured slower on B200: heavy CTAs co-resident on an SM contend
  }();
  xo::ModelArgs ord;
  const xo::ModelArgs *pord = nullptr;
  if (lpt && C > 1 && !weighted) {
    int *order = reinterpret_cast<int *>(ws + L.off_counter);  // the counter slot is free in this mode
    xo::launch_chain_order(elem, C, (int)P, order, st);
    ord.order = order;
    ord.elements_only = 1;
    pord = &ord;
  }
  if (weighted) {
    ord.weights = weights;
    ord.rows = rows;
    ord.nchains = (int)C;
    ord.grid_ctas = (int)xo::transit_weighted_grid(N, C);
    ord.elements_only = 1;
    pord = &ord;
  }
  // heavy-first inserts the ordering kernel between prologue and logp kernel: no early start then
  // (nor when the prologue ran on another stream: the kernel preceding this one is then unrelated)
  xo::launch_transit_logp_grad(t, y, ivar, N, elem, cl, mean, C, (int)P, dtk, wk, K, L.sp, moments, partial, st, pord,
                               pdl && !ord.order && phases == 3);
  cudaError_t le = io.q ? xo::launch_pdl(model_epilogue_kernel<true>, dim3((unsigned)((C + 3) / 4)), dim3(128), st, pdl, par, star,
                                         (const double *)partial, L.sp.slots, C, (int)P, lognorm, logp, gpar, gstar, gmean,
                                         (const int2 *)rows, io)
                        : xo::launch_pdl(model_epilogue_kernel<false>, dim3((unsigned)((C + 3) / 4)), dim3(128), st, pdl, par, star,
                                         (const double *)partial, L.sp.slots, C, (int)P, lognorm, logp, gpar, gstar, gmean,
                                         (const int2 *)rows, io);
  if (le != cudaSuccess) return -(int)le;
  return xo::launch_status();
}

}  // namespace xo

extern "C" {

size_t xo_transit_model_workspace(int64_t N, int64_t C, int64_t P) {
  if (N <= 0 || C <= 0 || P <= 0) return 0;
  return model_layout(N, C, P).bytes;
}

int xo_transit_model_logp_grad(const double *t, const double *y, const double *ivar, int64_t N, const double *par,
                               const double *star, const double *mean, int64_t C, int64_t P, const double *dtk,
                               const double *wk, int K, double texp, int use_in_transit, double lognorm,
                               const double *moments, double *logp, double *gpar, double *gstar, double *gmean,
                               void *workspace, size_t workspace_bytes, void *stream) {
  XO_NVTX_SCOPE("xo_transit_model_logp_grad");
  return xo::model_logp_grad_phases(t, y, ivar, N, par, star, mean, C, P, dtk, wk, K, texp, use_in_transit, lognorm, moments, logp,
                                    gpar, gstar, gmean, workspace, workspace_bytes, stream, 3);
}

// element blocks only (prologue), for callers that want to inspect / reuse them
int xo_elements_from_params(const double *par, const double *star, int64_t C, int64_t P, double texp, int use_in_transit,
                            double *elem, double *cl, void *stream) {
  XO_NVTX_SCOPE("xo_elements_from_params");
  if (C < 0 || P < 0) return XO_EINVAL;
  if (P > XO_MAX_PLANETS) return XO_ERANGE;
  if (C == 0 || P == 0) return 0;
  if (!par || !star || !elem || !cl) return XO_EINVAL;
  const int64_t cp = C * P;
  xo::launch(model_prologue_kernel<false>, (unsigned)((2 * cp + 63) / 64), 64, 0, static_cast<cudaStream_t>(stream), par, star, C, (int)P, texp,
             use_in_transit, elem, cl, (int64_t)0, (uint32_t *)nullptr, 0u, xo::PosteriorIO{});
  return xo::launch_status();
}

}  // extern "C"
