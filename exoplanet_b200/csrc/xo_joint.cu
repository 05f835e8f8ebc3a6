// xo_joint.cu -- BASELINE configs[3]: transit mean model + celerite2 SHOTerm GP noise model, log-likelihood and
// its gradient w.r.t. every sampler-side parameter in ONE stream-ordered call with no host round trip:
//
//   prologue            (period, t0, e, omega, b, r/R*, ...; R*, M*, u1, u2) -> element blocks + limb-darkening vector
//   residual            r[c][n] = y[n] - mean[c] - sum_p lc[c][n][p]                        (transit_resid_kernel)
//   GP scan + adjoint   logL[c], d logL/d(S0, w0, Q), g[c][n] = d logL/d r, sum_n g, sum_n d logL/d diag_n
//                                                                          (gp_forward_kernel, gp_backward_kernel)
//   light-curve VJP     cotangent -g[c][n] shared by the planets -> d logL/d elements, d/d c  (transit_logp_grad_kernel, VJP mode)
//   epilogue            fold + reverse pass of the prologue -> d logL/d parameters           (model_epilogue_kernel)
//
// What it replaces in a PyMC model written against the reference: LimbDarkLightCurve.get_light_curve
// (/root/reference/src/exoplanet/light_curves/limb_dark.py:99-232) as the mean of
// celerite2.pymc.GaussianProcess(SHOTerm).marginal (celerite2 is an optional extra, /root/reference/setup.py:36),
// and pytensor.grad of the resulting log-likelihood.
#include <cuda_runtime.h>

#include "xo_common.cuh"
#include "xo_transit.cuh"

namespace {

struct JointLayout {
  size_t off_elem, off_cl, off_resid, off_gres, off_gsum, off_scratch, off_transit, off_gp, transit_bytes, gp_bytes, bytes;
};
inline JointLayout joint_layout(int64_t N, int64_t C, int64_t P) {
  JointLayout L;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t at = o;
    o += (bytes + 255) & ~(size_t)255;
    return at;
  };
  L.off_gp = take(L.gp_bytes = xo_celerite_sho_grad_workspace(N, C));  // first: its head is the status word
  L.off_elem = take(sizeof(double) * (size_t)C * P * XO_NEL);
  L.off_cl = take(sizeof(double) * (size_t)C * 3);
  L.off_resid = take(sizeof(double) * (size_t)C * N);
  L.off_gres = take(sizeof(double) * (size_t)C * N);
  L.off_gsum = take(sizeof(double) * (size_t)C * 2);
  L.off_scratch = take(sizeof(double) * (size_t)C);
  L.off_transit = take(L.transit_bytes = xo_transit_logp_grad_workspace(N, C, P));
  L.bytes = o;
  return L;
}

}  // namespace

extern "C" {

size_t xo_transit_gp_workspace(int64_t N, int64_t C, int64_t P) {
  if (N <= 0 || C <= 0 || P <= 0) return 0;
  return joint_layout(N, C, P).bytes;
}

int xo_transit_gp_logp_grad(const double *t, const double *y, const double *diag, int64_t N, const double *par,
                            const double *star, const double *mean, const double *S0, const double *w0, const double *Q,
                            const double *jitter, int64_t C, int64_t P, const double *dtk, const double *wk, int K, double texp,
                            int use_in_transit, double eps, double *loglike, double *gpar, double *gstar, double *gmean,
                            double *ghyper, double *gjitter, void *workspace, size_t workspace_bytes, void *stream) {
  if (int rc = xo::check_batch(N, C, P)) return rc;
  if (K < 1 || K > XO_MAX_STENCIL) return XO_ERANGE;
  if (C == 0) return 0;
  if (N == 0 || P == 0) return XO_EINVAL;
  if (!t || !y || !diag || !par || !star || !mean || !S0 || !w0 || !Q || !dtk || !wk || !loglike || !gpar || !gstar || !gmean ||
      !ghyper || !workspace)
    return XO_EINVAL;
  const JointLayout L = joint_layout(N, C, P);
  if (workspace_bytes < L.bytes) return XO_EWORKSPACE;
  XO_NVTX_SCOPE("xo_transit_gp_logp_grad");
  char *ws = static_cast<char *>(workspace);
  auto at = [&](size_t off) { return reinterpret_cast<double *>(ws + off); };
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double *elem = at(L.off_elem), *cl = at(L.off_cl), *resid = at(L.off_resid), *gres = at(L.off_gres), *gsum = at(L.off_gsum);
  int rc = xo_elements_from_params(par, star, C, P, texp, use_in_transit, elem, cl, stream);
  if (rc) return rc;
  xo::launch_transit_resid(t, y, N, elem, cl, mean, C, (int)P, dtk, wk, K, resid, st);
  rc = xo::celerite_sho_grad_impl(t, resid, diag, N, S0, w0, Q, jitter, C, eps, loglike, ghyper, gres, nullptr, gsum, ws + L.off_gp,
                                  L.gp_bytes, stream);
  if (rc) return rc;
  // reverse pass of the light curve with the cotangent -gres[c][n] (shared by the planets); gmean = -sum_n gres
  const xo::Split sp = xo::split_transit(N, C);
  xo::ModelArgs ma;
  ma.elements_only = 1;
  ma.vjp_shared = 1;
  ma.vjp_gsum = gsum;
  double *partial = at(L.off_transit);
  xo::launch_transit_logp_grad(t, gres, nullptr, N, elem, cl, nullptr, C, (int)P, dtk, wk, K, sp, nullptr, partial, st, &ma);
  rc = xo::launch_model_epilogue(par, star, partial, sp.slots, C, (int)P, 0.0, at(L.off_scratch), gpar, gstar, gmean, st, false);
  if (rc) return rc;
  if (gjitter) {
    cudaError_t e = cudaMemcpy2DAsync(gjitter, sizeof(double), gsum + 1, 2 * sizeof(double), sizeof(double), (size_t)C,
                                      cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return -(int)e;
  }
  return xo::launch_status();
}

}  // extern "C"
