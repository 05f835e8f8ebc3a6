/* xo_b200.h -- C ABI of libxo_b200.so, the B200 (sm_100a) engine behind the
 * operator boundary of exoplanet-dev/exoplanet.
 *
 * The boundary being replaced is the module object `ops` that the reference binds at
 *   /root/reference/src/exoplanet/compat.py:56   (from exoplanet_core.pymc import ops)
 * and touches at exactly four call sites:
 *   ops.kepler                /root/reference/src/exoplanet/orbits/keplerian.py:333, :818
 *   ops.quad_solution_vector  /root/reference/src/exoplanet/light_curves/limb_dark.py:24
 *   ops.contact_points        /root/reference/src/exoplanet/orbits/keplerian.py:744-753
 * exoplanet-core's own FFI for those ops is a pybind11 `driver` that fills flat,
 * contiguous fp64 host arrays in place; the `*_host` entry points below have that
 * shape (plain pointers + element count) and are what a maintainer would bind.
 * The device-pointer entry points are the same kernels for callers that keep
 * their arrays in HBM (torch CUDA tensors): no torch types cross this ABI.
 *
 * Conventions
 *  - every array is contiguous fp64 unless stated; sizes are int64_t element counts;
 *  - device entry points take device pointers and a `stream` (cudaStream_t passed as
 *    void*; NULL = the legacy default stream), enqueue asynchronously, never allocate,
 *    never free caller memory; scratch is passed in explicitly (workspace);
 *  - `*_host` entry points take HOST pointers, run on CUDA device `device`,
 *    copy in, launch, copy out and synchronise before returning; their scratch is
 *    stream-ordered (cudaMallocAsync/cudaFreeAsync), nothing persists;
 *  - return value: 0 = ok; negative = -(cudaError_t); positive = XO_E* argument error.
 *    No C++ exception crosses the ABI.  NaN in -> NaN out; every iteration is bounded.
 *  - the library is stateless and re-entrant.
 */
#ifndef XO_B200_H
#define XO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XO_ABI_VERSION 1

/* argument errors */
#define XO_EINVAL 1   /* bad size / null pointer */
#define XO_ERANGE 2   /* K (stencil) or P (planets) above the compiled maximum */
#define XO_EWORKSPACE 3 /* workspace too small */

/* Element block: one per (chain, planet), XO_NEL doubles, array-of-structs
 * [C][P][XO_NEL].  It is what KeplerianOrbit.__init__ derives per planet
 * (keplerian.py:133-281) plus the in_transit() window (keplerian.py:729-767). */
#define XO_NEL 16
#define XO_EL_T0 0      /* reference transit time t0                    */
#define XO_EL_TREF 1    /* tref = t_periastron - t0   (keplerian.py:279) */
#define XO_EL_N 2       /* mean motion n = 2 pi / period (:146)          */
#define XO_EL_ECC 3     /* e (0 => circular branch, :331-332)            */
#define XO_EL_COSW 4
#define XO_EL_SINW 5
#define XO_EL_COSI 6
#define XO_EL_SINI 7
#define XO_EL_AOR 8     /* a / R_star                                    */
#define XO_EL_ROR 9     /* r_planet / R_star                             */
#define XO_EL_PERIOD 10
#define XO_EL_TSTART 11 /* in-transit window on dt = mod(t-t0+P/2,P)-P/2; */
#define XO_EL_TEND 12   /* (-inf,+inf) = evaluate every cadence          */
#define XO_EL_KRV 13    /* RV semi-amplitude (keplerian.py:658-670)      */
#define XO_NGRAD 10     /* d/d(t0,tref,n,e,cosw,sinw,cosi,sini,aor,ror)  */
#define XO_NGRAD_RV 7   /* d/d(t0,tref,n,e,cosw,sinw,Krv)                */
#define XO_NPAR 8       /* per (chain, planet): period, t0, ecc, omega, b, r/R*, m_planet, K_rv */
#define XO_NSTAR 4      /* per chain: R_star, M_star, u1, u2                              */
#define XO_MAX_STENCIL 33 /* exposure-time sub-samples per cadence        */
#define XO_MAX_PLANETS 8

const char *xo_version(void);
int xo_abi_version(void);
int xo_device_count(void);
/* human-readable text for a return value of any entry point */
const char *xo_strerror(int status);

/* ---- kepler(M, e) -> (sin f, cos f)          replaces keplerian.py:333 ---- */
int xo_kepler_fwd(const double *M, const double *e, double *sinf, double *cosf, int64_t n, void *stream);
/* e given once per element block of `inner` consecutive M values (the reference
 * broadcasts a per-planet scalar, keplerian.py:333); n = total elements        */
int xo_kepler_fwd_bcast(const double *M, const double *e, int64_t e_stride_inner, double *sinf, double *cosf, int64_t n,
                        void *stream);
/* VJP of the op: (bsinf, bcosf) -> (bM, be) */
int xo_kepler_vjp(const double *sinf, const double *cosf, const double *e, const double *bsinf, const double *bcosf,
                  double *bM, double *be, int64_t n, void *stream);

/* ---- quad_solution_vector(b, r) -> s[n][3]  replaces limb_dark.py:24 ----
 * dsdb / dsdr ([n][3], may both be NULL) are the Jacobians the VJP consumes. */
int xo_qsv_fwd(const double *b, const double *r, double *s, double *dsdb, double *dsdr, int64_t n, void *stream);
int xo_qsv_vjp(const double *bs, const double *dsdb, const double *dsdr, double *bb, double *br, int64_t n, void *stream);

/* ---- contact_points(a,e,cosw,sinw,cosi,sini,L) -> (M_left,M_right,flag)
 *      replaces keplerian.py:744-753; flag != 0 => caller evaluates all cadences */
int xo_contact_points(const double *a, const double *e, const double *cosw, const double *sinw, const double *cosi,
                      const double *sini, const double *L, double *M_left, double *M_right, int32_t *flag, int64_t n,
                      void *stream);

/* ---- fused L3 path (keplerian.py:324-409 + limb_dark.py:163-252) ----
 * t[N]; elem[C][P][XO_NEL]; cl[C][3] (get_cl, limb_dark.py:11-18);
 * stencil dtk[K] (already multiplied by texp) and wk[K] (limb_dark.py:181-198), K=1,dtk=0,wk=1 for texp=None.
 * lc[C][N][P]: exact 0.0 outside the window / for z<=0 / for b>=1+r. */
int xo_transit_lc_fwd(const double *t, int64_t N, const double *elem, const double *cl, int64_t C, int64_t P,
                      const double *dtk, const double *wk, int K, double *lc, void *stream);

/* reverse pass of xo_transit_lc_fwd: bar_lc[C][N][P] -> gel[C][P][XO_NGRAD], gcl[C][3];
 * workspace >= xo_transit_logp_grad_workspace(N, C, P) + 16 * C bytes */
int xo_transit_lc_vjp(const double *t, int64_t N, const double *elem, const double *cl, int64_t C, int64_t P,
                      const double *dtk, const double *wk, int K, const double *bar_lc, double *gel, double *gcl,
                      void *workspace, size_t workspace_bytes, void *stream);

/* Gaussian log-likelihood and its gradient in one launch:
 *   logp[c] = lognorm - 1/2 sum_n ivar[n] (y[n] - mean[c] - sum_p lc[c][n][p])^2
 *   gel[C][P][XO_NGRAD], gcl[C][3], gmean[C].
 * workspace: xo_transit_logp_grad_workspace(N, C, P) bytes of device memory.
 * moments: NULL, or the block written by xo_transit_data_moments(t, y, ivar) (a device buffer of
 *   XO_MOMENTS_LEN doubles; the first 8 are the result, the rest scratch of the reduction) --
 *   (y_ref, sum ivar, sum ivar (y-y_ref), sum ivar (y-y_ref)^2, t_is_sorted, 0, 0, 0), y_ref a value
 *   inside the data's range so the shifted sums carry no cancellation; with it the
 *   out-of-transit part of the sums is closed-form, the sweep reads t only, and for sorted t whole
 *   128-cadence steps are skipped by one interval test (same result to rounding). */
size_t xo_transit_logp_grad_workspace(int64_t N, int64_t C, int64_t P);
#define XO_MOMENTS_LEN 1024
int xo_transit_data_moments(const double *t, const double *y, const double *ivar, int64_t N,
                            double *moments /*[XO_MOMENTS_LEN]*/,
                            void *stream);
/* Optional time index behind the moments (sorted t only): allocate xo_transit_moments_len(N) doubles instead of
 * XO_MOMENTS_LEN, call xo_transit_data_moments and then xo_transit_data_index on the same stream.  The block then
 * also holds table[j] = #{n : t[n] < t[0] + j w} on a uniform grid of N/2 bins (moments[5] = number of bins,
 * moments[6] = 1/w), and the logp kernels find the cadences of every transit window by table look-up instead of
 * testing time stamps (same result, same evaluated set).  A block without the index (moments[5] = 0) keeps working. */
size_t xo_transit_moments_len(int64_t N); /* in doubles */
int xo_transit_data_index(const double *t, int64_t N, double *moments, size_t moments_len, void *stream);
/* both in ONE pass over the data (the index costs a few integer operations per cadence on top of the moments' loads): what
 * the per-call host entry points and the resident data handle use */
int xo_transit_data_moments_index(const double *t, const double *y, const double *ivar, int64_t N, double *moments,
                                  size_t moments_len, void *stream);
int xo_transit_logp_grad(const double *t, const double *y, const double *ivar, int64_t N, const double *elem,
                         const double *cl, const double *mean, int64_t C, int64_t P, const double *dtk, const double *wk,
                         int K, double lognorm, const double *moments, double *logp, double *gel, double *gcl,
                         double *gmean, void *workspace, size_t workspace_bytes, void *stream);

/* The same evaluation from sampler-side parameters: prologue (KeplerianOrbit.__init__,
 * keplerian.py:133-281; in_transit window :729-767; get_cl limb_dark.py:11-18) + logp +
 * the prologue's reverse pass.  par[C][P][XO_NPAR], star[C][XO_NSTAR], mean[C] ->
 * logp[C], gpar[C][P][XO_NPAR], gstar[C][XO_NSTAR], gmean[C].  texp = exposure time used to
 * widen the window (0 if none; the stencil dtk/wk carries the integration itself). */
size_t xo_transit_model_workspace(int64_t N, int64_t C, int64_t P);
int xo_transit_model_logp_grad(const double *t, const double *y, const double *ivar, int64_t N, const double *par,
                               const double *star, const double *mean, int64_t C, int64_t P, const double *dtk,
                               const double *wk, int K, double texp, int use_in_transit, double lognorm,
                               const double *moments, double *logp, double *gpar, double *gstar, double *gmean,
                               void *workspace, size_t workspace_bytes, void *stream);
/* prologue only: elem[C][P][XO_NEL], cl[C][3] */
int xo_elements_from_params(const double *par, const double *star, int64_t C, int64_t P, double texp, int use_in_transit,
                            double *elem, double *cl, void *stream);

/* ---- the re-entrant callers of the same kernel: TTVOrbit and light-travel delay ----
 * TTVOrbit (/root/reference/src/exoplanet/orbits/ttv.py:35-187): ttv_times[C][P][ttv_nt] holds, per (chain, planet), the sorted
 * transit times (TTVOrbit.all_transit_times), ttv_count[P] (DEVICE int array) valid entries each; NULL = ordinary orbit.
 * A cadence is measured from its NEAREST transit (bin edges half-way between consecutive transits, ttv.py:159-187) instead
 * of from elem[XO_EL_T0]; elem[XO_EL_T0] / XO_EL_PERIOD must be the linear ephemeris the table departs from (transit j near
 * t0 + j P).  gttv[C][P][ttv_nt] = d / d transit time (zeroed by the call, accumulated with one atomic per transit and warp:
 * sums of more than two partials may differ in the last bit between runs).
 * Light-travel delay (orbits/keplerian.py:411-470, get_relative_position(light_delay=True) -> limb_dark.py:215): light_delay
 * != 0 and elem[XO_EL_RSTAR] = R_star in R_sun; every sample is evaluated at the retarded time t - D(position, line-of-sight
 * velocity and acceleration at t) -- a second Kepler solve -- and the reverse pass carries d D / d elements (sin i included;
 * d / d R_star is not provided).  The reference evaluates every cadence in this mode (in_transit raises for light_delay,
 * keplerian.py:720-723): pass the window of an ordinary orbit widened by 2 a (1 + e) / c, or (-inf, +inf).
 * Per-cadence exposure times (/root/reference/src/exoplanet/light_curves/limb_dark.py:199-205, tests/light_curves_test.py:105):
 * texp_vec[N] != NULL scales the stencil of every cadence, t_i + texp_vec[i] * dtk[k]; dtk then holds the UNIT stencil
 * (offsets in exposure times) and the element windows must be wide enough for the longest exposure (or infinite, the
 * reference's use_in_transit=False).  Not combined with light_delay in the reverse pass. */
#define XO_EL_RSTAR 14
int xo_transit_lc_fwd_ex(const double *t, int64_t N, const double *elem, const double *cl, int64_t C, int64_t P,
                         const double *dtk, const double *wk, int K, const double *ttv_times, const int *ttv_count, int ttv_nt,
                         int light_delay, const double *texp_vec, double *lc, void *stream);
int xo_transit_lc_vjp_ex(const double *t, int64_t N, const double *elem, const double *cl, int64_t C, int64_t P,
                         const double *dtk, const double *wk, int K, const double *ttv_times, const int *ttv_count, int ttv_nt,
                         int light_delay, const double *texp_vec, const double *bar_lc, double *gel, double *gcl, double *gttv,
                         void *workspace, size_t workspace_bytes, void *stream);
int xo_transit_logp_grad_ttv(const double *t, const double *y, const double *ivar, int64_t N, const double *elem,
                             const double *cl, const double *mean, int64_t C, int64_t P, const double *dtk, const double *wk,
                             int K, double lognorm, const double *moments, const double *ttv_times, const int *ttv_count,
                             int ttv_nt, double *logp, double *gel, double *gcl, double *gmean, double *gttv, void *workspace,
                             size_t workspace_bytes, void *stream);

/* relative position / R_star and RV shape for every (chain, cadence, planet):
 * X,Y,Z,rv each [C][N][P], any may be NULL (keplerian.py:380-409, 658-670) */
int xo_orbit_eval(const double *t, int64_t N, const double *elem, int64_t C, int64_t P, double *X, double *Y, double *Z,
                  double *rv, void *stream);

/* RV Gaussian log-likelihood + gradient:
 *   model[c][n] = gamma[c] + sum_p Krv_p (cos(w+f) + e cos w);  grv[C][P][XO_NGRAD_RV], ggamma[C] */
size_t xo_rv_logp_grad_workspace(int64_t N, int64_t C, int64_t P);
int xo_rv_logp_grad(const double *t, const double *y, const double *ivar, int64_t N, const double *elem,
                    const double *gamma, int64_t C, int64_t P, double lognorm, double *logp, double *grv, double *ggamma,
                    void *workspace, size_t workspace_bytes, void *stream);

/* ---- celerite SHO GP log-likelihood (celerite2 SHOTerm; no reference call site) ----
 * B chains share (t, diag); y[B][N] (residuals), params S0[B], w0[B], Q[B]; out loglike[B].
 * eps as celerite2's SHOTerm(eps=1e-5). */
size_t xo_celerite_sho_workspace(int64_t N, int64_t B);
int xo_celerite_sho_loglike(const double *t, const double *y, const double *diag, int64_t N, const double *S0,
                            const double *w0, const double *Q, int64_t B, double eps, double *loglike, void *workspace,
                            size_t workspace_bytes, void *stream);

/* log-likelihood and alpha = K^-1 y in one call (d loglike / d y = -alpha, the cotangent a mean
 * model such as the transit light curve needs); alpha[B][N]. */
size_t xo_celerite_sho_solve_workspace(int64_t N, int64_t B);
int xo_celerite_sho_solve(const double *t, const double *y, const double *diag, int64_t N, const double *S0,
                          const double *w0, const double *Q, int64_t B, double eps, double *loglike, double *alpha,
                          void *workspace, size_t workspace_bytes, void *stream);

/* log-likelihood and its exact reverse-mode gradient (what celerite2's own backward pass gives):
 * ghyper[B][3] = d loglike / d(S0, w0, Q); optional gy[B][N] = d/dy (= -K^-1 y) and
 * gdiag[B][N] = d/d diag (sum over chains for a shared jitter term).  The adjoint of the Kalman
 * recursion is affine in a 5-vector and runs as a second (reverse) parallel scan. */
size_t xo_celerite_sho_grad_workspace(int64_t N, int64_t B);
int xo_celerite_sho_grad(const double *t, const double *y, const double *diag, int64_t N, const double *S0,
                         const double *w0, const double *Q, int64_t B, double eps, double *loglike, double *ghyper,
                         double *gy, double *gdiag, void *workspace, size_t workspace_bytes, void *stream);

/* ---- joint transit + celerite SHO GP model (BASELINE configs[3]) ----
 * loglike[c] = log N(y - mean[c] - sum_p lc(par[c], star[c]); 0, K_SHO(S0[c], w0[c], Q[c]) + diag + jitter[c])
 * and its gradient w.r.t. every parameter, one stream-ordered call (6 kernels + 1 memset + 1 strided copy, no
 * host round trip): par[C][P][XO_NPAR], star[C][XO_NSTAR], mean[C], S0/w0/Q[C], jitter[C] or NULL (added to
 * diag[N]); outputs loglike[C], gpar, gstar, gmean[C], ghyper[C][3] = d/d(S0, w0, Q), gjitter[C] or NULL.
 * The first 16 bytes of the workspace are the status word of xo_read_status. */
size_t xo_transit_gp_workspace(int64_t N, int64_t C, int64_t P);
int xo_transit_gp_logp_grad(const double *t, const double *y, const double *diag, int64_t N, const double *par,
                            const double *star, const double *mean, const double *S0, const double *w0, const double *Q,
                            const double *jitter, int64_t C, int64_t P, const double *dtk, const double *wk, int K, double texp,
                            int use_in_transit, double eps, double *loglike, double *gpar, double *gstar, double *gmean,
                            double *ghyper, double *gjitter, void *workspace, size_t workspace_bytes, void *stream);

/* ---- posterior in unconstrained coordinates + lock-step HMC ----
 * The PyMC model of /root/reference/docs/tutorials/data-and-models.md:458-509 as one device-resident evaluation:
 * q[C][D], D = 6 P + 3 = per planet (log period, t0, sqrt(e) cos w, sqrt(e) sin w, b, log r/R*), per chain
 * (logit q1, logit q2, mean flux) with Kipping's (q1, q2) -> (u1, u2) (/root/reference/src/exoplanet/distributions.py:98-102),
 * R* = M* = 1, independent Normal priors on q: prior_mu / prior_sigma are [D] with prior_chain_stride = 0 (shared by the
 * chains) or [C][prior_chain_stride] (one row per chain).
 *   logp[c] = Gaussian log-likelihood of the light curve + log prior (without the priors' normalisation constants)
 *   grad[C][D] = d logp / d q      (change of variables, its chain rule and the priors are two small kernels
 *                                   around xo_transit_model_logp_grad; moments may be NULL)
 * xo_hmc_transition advances every chain by one HMC transition of n_leapfrog steps (momenta ~ N(0, 1/inv_mass),
 * per-chain step sizes eps[C], Metropolis accept/reject) entirely on `stream`: 2 + 4 n_leapfrog kernel launches, no
 * host synchronisation; q / logp / grad are the chains' current state (in and out).  Random numbers: Philox with
 * (seed, chain, counter): pass a different `counter` for every transition. */
size_t xo_transit_posterior_workspace(int64_t N, int64_t C, int64_t P);
int xo_transit_posterior_logp_grad(const double *t, const double *y, const double *ivar, int64_t N, const double *moments,
                                   const double *q, int64_t C, int64_t P, const double *prior_mu, const double *prior_sigma,
                                   int64_t prior_chain_stride, const double *dtk, const double *wk, int K, double texp, int use_in_transit, double lognorm,
                                   double *logp, double *grad, void *workspace, size_t workspace_bytes, void *stream);
int xo_hmc_transition(const double *t, const double *y, const double *ivar, int64_t N, const double *moments, int64_t C, int64_t P,
                      const double *prior_mu, const double *prior_sigma, int64_t prior_chain_stride, const double *dtk, const double *wk,
                      int K, double texp, int use_in_transit, double lognorm, double *q, double *logp, double *grad, const double *eps,
                      const double *inv_mass, int n_leapfrog, unsigned long long seed, unsigned long long counter,
                      double *accept_prob, double *energy_error, void *workspace, size_t workspace_bytes, void *stream);

/* ---- host-buffer entry points (the shape of exoplanet-core's driver) ---- */
int xo_kepler_host(const double *M, const double *e, double *sinf, double *cosf, int64_t n, int device);
int xo_qsv_host(const double *b, const double *r, double *s, double *dsdb, double *dsdr, int64_t n, int device);
int xo_contact_points_host(const double *a, const double *e, const double *cosw, const double *sinw, const double *cosi,
                           const double *sini, const double *L, double *M_left, double *M_right, int32_t *flag,
                           int64_t n, int device);
int xo_transit_lc_host(const double *t, int64_t N, const double *elem, const double *cl, int64_t C, int64_t P,
                       const double *dtk, const double *wk, int K, double *lc, int device);
int xo_transit_logp_grad_host(const double *t, const double *y, const double *ivar, int64_t N, const double *elem,
                              const double *cl, const double *mean, int64_t C, int64_t P, const double *dtk,
                              const double *wk, int K, double lognorm, double *logp, double *gel, double *gcl,
                              double *gmean, int device);
int xo_transit_model_logp_grad_host(const double *t, const double *y, const double *ivar, int64_t N, const double *par,
                                    const double *star, const double *mean, int64_t C, int64_t P, const double *dtk,
                                    const double *wk, int K, double texp, int use_in_transit, double lognorm,
                                    double *logp, double *gpar, double *gstar, double *gmean, int device);


/* ---- resident data set (what a compiled PyTensor function holds as constants) ----
 * The light curve (t, y, ivar) is uploaded once and owned by the handle; each evaluation moves
 * only O(#params) bytes: par/star/mean in, logp/gradients out (all HOST buffers). */
typedef struct xo_transit_data xo_transit_data;
int xo_transit_data_create(const double *t, const double *y, const double *ivar, int64_t N, int device,
                           xo_transit_data **out);
int xo_transit_data_destroy(xo_transit_data *data);
int xo_transit_model_logp_grad_resident(xo_transit_data *data, const double *par, const double *star, const double *mean,
                                        int64_t C, int64_t P, const double *dtk, const double *wk, int K, double texp,
                                        int use_in_transit, double lognorm, int use_moments, double *logp, double *gpar,
                                        double *gstar, double *gmean);

/* ---- measurement helpers (no reference counterpart) ----
 * xo_diag_dfma_peak: register-resident DFMA loop on every SM of the current device, best of `reps` launches timed
 *   with CUDA events on `stream` -> thread-level DFMA per second (x2 = flop/s): the measured fp64 ceiling the
 *   fused kernels are quoted against (SURVEY.md 8d ceiling (B)).
 * xo_diag_launch_count: number of kernels this library has launched in this process so far. */
/* Device status word.  The celerite entry points keep, in the first 16 bytes of their workspace, (a) the number of
 * lanes whose result was not finite (NaN / inf input, non-positive variance) and (b) the number of inter-CTA waits of
 * the one-pass scans that expired (never expected; the result must then be discarded).  They are written by the
 * asynchronous launches; this call copies them to the host on `stream` and synchronises it.  This is SURVEY.md 8(b)'s
 * "count of non-converged/NaN lanes written to a device status word" for entry points that cannot return it
 * themselves because they do not synchronise. */
int xo_read_status(const void *workspace, unsigned long long *bad_lanes, unsigned long long *wait_failures, void *stream);

int xo_diag_dfma_peak(int reps, double *dfma_per_s, double *best_ms, void *stream);
long long xo_diag_launch_count(void);

/* ---- the tree bookkeeping of a lock-step No-U-Turn sampler (csrc/xo_nuts.cu; SURVEY.md 8f #4: what PyMC's NUTS,
 * /root/reference/docs/tutorials/data-and-models.md:505-509, does per chain and process) -------------------------------
 * Multinomial NUTS with the generalised U-turn criterion over C chains in step; D coordinates per chain; any logp+gradient
 * callback in between (the fused transit posterior above is one).  `state` is one block of xo_nuts_state_layout() doubles
 * (flags 0.0 / 1.0): offsets[0..13] the per-chain scalars h0, log_w, lp_prop, active, depth, diverged, acc_sum, n_leaves,
 * dh_max, sgn, sub_ok, s_logw, s_lp, go_right ([C] each); offsets[14..28] the per-coordinate arrays ql, pl, gl, qr, pr,
 * gr, q_prop, g_prop, p_sum, q, p, g, s_q, s_g, s_psum ([C][D] each); offsets[29], [30] the U-turn checkpoints ck_p, ck_s
 * ([max_depth + 1][C][D] each).  One doubling j: xo_nuts_start (u_dir[C] uniform: direction; the new half starts from that
 * edge), then for leaf = 0 .. 2^j - 1: xo_nuts_leaf_pre (half kick, drift: state.q is where logp is wanted),
 * <caller evaluates logp[C], grad[C][D] at state.q>, xo_nuts_leaf_post (u_take[C] uniform); then xo_nuts_merge
 * (u_merge[C] uniform) -- state.active then says which chains keep growing.  The uniforms come from the caller, so the
 * kernels are deterministic and can be checked state for state against the tensor implementation. */
size_t xo_nuts_state_layout(int64_t C, int D, int max_depth, int64_t *offsets /*[31] or NULL*/);
int xo_nuts_start(double *state, int64_t C, int D, int max_depth, const double *u_dir, const double *logp_cur, void *stream);
int xo_nuts_leaf_pre(double *state, int64_t C, int D, int max_depth, const double *eps, const double *inv_mass, void *stream);
int xo_nuts_leaf_post(double *state, int64_t C, int D, int max_depth, const double *eps, const double *inv_mass, const double *logp,
                      const double *grad, const double *u_take, int leaf, void *stream);
int xo_nuts_merge(double *state, int64_t C, int D, int max_depth, const double *inv_mass, const double *u_merge, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* XO_B200_H */
