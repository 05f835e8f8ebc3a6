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

#include "xo_common.cuh"
#include "xo_contact.cuh"
#include "xo_math.cuh"
#include "xo_transit.cuh"

namespace {

constexpr double kGGrav = 2942.2062175044193;  // R_sun^3 / M_sun / day^2 (orbits/constants.py:32)
constexpr unsigned kFullM = 0xffffffffu;

// two adjacent lanes per (chain, planet): lane parity = contact side, so the two bracketed
// Newton solves of the in-transit window run concurrently; even lane writes the element block
__global__ void model_prologue_kernel(const double *__restrict__ par, const double *__restrict__ star, int64_t C, int P,
                                      double texp, int use_in_transit, double *__restrict__ elem,
                                      double *__restrict__ cl) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t idx = tid >> 1;
  const int side = (int)(tid & 1);
  const bool active = idx < C * P;
  const int64_t ii = active ? idx : 0;
  const int64_t c = ii / P;
  const double *q = par + ii * XO_NPAR;
  const double period = q[0], t0 = q[1], e = q[2], omega = q[3], b = q[4], ror = q[5], mpl = q[6], K = q[7];
  const double R = star[c * XO_NSTAR + 0], Ms = star[c * XO_NSTAR + 1];
  const double a = cbrt(kGGrav * (Ms + mpl) * period * period / (4.0 * xo::kPi * xo::kPi));
  const double n = 2.0 * xo::kPi / period;
  double sw, cw;
  sincos(omega, &sw, &cw);
  const double y = sqrt(1.0 - e) * cw, x = sqrt(1.0 + e) * (1.0 + sw);
  const double E0 = 2.0 * atan2(y, x);
  const double M0 = E0 - e * sin(E0);
  const double ome2 = (1.0 - e) * (1.0 + e);
  const double F = (1.0 + e * sw) / ome2;
  const double cosi = F * R / a * b;
  const double sini = sqrt((1.0 - cosi) * (1.0 + cosi));
  double tc = side ? INFINITY : -INFINITY;  // this lane's contact time relative to t0
  int flag = 0;
  if (use_in_transit) {
    double Mc;
    flag = xo::cp_solve_side(a, e, cw, sw, cosi, R * (1.0 + ror), side ? 1.0 : -1.0, Mc);
    if (flag == 0) {
      const double hp = 0.5 * period;
      tc = (Mc - M0) / n + hp;
      tc = tc - period * floor(tc / period) - hp;
      if (side == 0) {
        if (tc > 0.0) tc -= period;
        tc -= 0.5 * texp;
      } else {
        if (tc < 0.0) tc += period;
        tc += 0.5 * texp;
      }
    }
  }
  const double tc_other = __shfl_xor_sync(0xffffffffu, tc, 1);
  const int flag_other = __shfl_xor_sync(0xffffffffu, flag, 1);
  if (!active || side != 0) return;
  double ts = tc, te = tc_other;
  if (flag | flag_other) {
    ts = -INFINITY;
    te = INFINITY;
  }
  double *el = elem + idx * XO_NEL;
  el[XO_EL_T0] = t0;
  el[XO_EL_TREF] = -M0 / n;
  el[XO_EL_N] = n;
  el[XO_EL_ECC] = e;
  el[XO_EL_COSW] = cw;
  el[XO_EL_SINW] = sw;
  el[XO_EL_COSI] = cosi;
  el[XO_EL_SINI] = sini;
  el[XO_EL_AOR] = a / R;
  el[XO_EL_ROR] = ror;
  el[XO_EL_PERIOD] = period;
  el[XO_EL_TSTART] = ts;
  el[XO_EL_TEND] = te;
  el[XO_EL_KRV] = K;
  el[14] = 0.0;
  el[15] = 0.0;
  if (idx % P == 0) {
    const double u1 = star[c * XO_NSTAR + 2], u2 = star[c * XO_NSTAR + 3];
    const double c0 = 1.0 - u1 - 1.5 * u2, c1 = u1 + 2.0 * u2, c2 = -0.25 * u2;
    const double rn = 1.0 / (xo::kPi * (c0 + c1 / 1.5));
    cl[3 * c + 0] = c0 * rn;
    cl[3 * c + 1] = c1 * rn;
    cl[3 * c + 2] = c2 * rn;
  }
}

// one warp per chain: lane p < P folds planet p's 10 partial sums and runs the reverse pass
// of the prologue; lanes 8..12 fold (logp, gmean, gcl[3]).
__global__ void model_epilogue_kernel(const double *__restrict__ par, const double *__restrict__ star,
                                      const double *__restrict__ partial, int slots, int64_t C, int P, double lognorm,
                                      double *__restrict__ logp, double *__restrict__ gpar, double *__restrict__ gstar,
                                      double *__restrict__ gmean) {
  const int lane = threadIdx.x & 31;
  const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  const int nacc = xo::kNBase + P * XO_NGRAD;
  const double *src = partial + c * slots * nacc;
  // fold: lane l sums slots l, l+32, ... of every accumulator (ascending), then a fixed-order
  // butterfly combines the 32 lane sums -> deterministic, and 32x less serial latency
  double g[XO_NGRAD];
#pragma unroll
  for (int j = 0; j < XO_NGRAD; ++j) g[j] = 0.0;
  double base = 0.0;
  for (int j = 0; j < nacc; ++j) {
    double acc = 0.0;
    for (int s = lane; s < slots; s += 32) acc += src[(int64_t)s * nacc + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFullM, acc, o);
    if (j < xo::kNBase) {
      if (lane == 8 + j) base = acc;
    } else {
      const int pj = (j - xo::kNBase) / XO_NGRAD, k = (j - xo::kNBase) % XO_NGRAD;
      if (lane == pj) {
#pragma unroll
        for (int kk = 0; kk < XO_NGRAD; ++kk)
          if (kk == k) g[kk] = acc;
      }
    }
  }
  const double R = star[c * XO_NSTAR + 0], Ms = star[c * XO_NSTAR + 1];
  double gR = 0.0, gMs = 0.0;
  if (lane < P) {
    const double *q = par + (c * P + lane) * XO_NPAR;
    const double period = q[0], e = q[2], omega = q[3], b = q[4], mpl = q[6];
    const double mt = Ms + mpl;
    const double a = cbrt(kGGrav * mt * period * period / (4.0 * xo::kPi * xo::kPi));
    const double n = 2.0 * xo::kPi / period;
    double sw, cw;
    sincos(omega, &sw, &cw);
    const double sme = sqrt(1.0 - e), spe = sqrt(1.0 + e);
    const double y = sme * cw, x = spe * (1.0 + sw);
    const double E0 = 2.0 * atan2(y, x);
    double sE0, cE0;
    sincos(E0, &sE0, &cE0);
    const double M0 = E0 - e * sE0;
    const double ome2 = (1.0 - e) * (1.0 + e);
    const double F = (1.0 + e * sw) / ome2;
    const double cosi = F * R / a * b;
    const double sini = sqrt((1.0 - cosi) * (1.0 + cosi));
    // cotangents of the element block: g = (t0, tref, n, e, cw, sw, cosi, sini, aor, ror)
    double gt0 = g[0], ge = g[3], gcw = g[4], gsw = g[5];
    double ga = g[8] / R;
    gR += -g[8] * a / (R * R);
    double gcosi = g[6] + (sini > 0.0 ? -g[7] * cosi / sini : 0.0);
    double gF = gcosi * R * b / a;
    gR += gcosi * F * b / a;
    ga += -gcosi * F * R * b / (a * a);
    double gb = gcosi * F * R / a;
    ge += gF * (sw / ome2 + (1.0 + e * sw) * 2.0 * e / (ome2 * ome2));
    gsw += gF * e / ome2;
    double gM0 = -g[1] / n;
    double gn = g[2] + g[1] * M0 / (n * n);
    double gP = -gn * n / period;
    double gE0 = gM0 * (1.0 - e * cE0);
    ge += -gM0 * sE0;
    const double rr = 1.0 / (x * x + y * y);
    double gy = 2.0 * gE0 * x * rr, gx = -2.0 * gE0 * y * rr;
    gcw += gy * sme;
    ge += -gy * cw / (2.0 * sme);
    gsw += gx * spe;
    ge += gx * (1.0 + sw) / (2.0 * spe);
    double gomega = -gcw * sw + gsw * cw;
    gP += ga * (2.0 / 3.0) * a / period;
    double gmt = ga * a / (3.0 * mt);
    gMs += gmt;
    double *o = gpar + (c * P + lane) * XO_NPAR;
    o[0] = gP;
    o[1] = gt0;
    o[2] = ge;
    o[3] = gomega;
    o[4] = gb;
    o[5] = g[9];
    o[6] = gmt;
    o[7] = 0.0;  // K does not enter the transit likelihood
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {  // planets live in lanes 0..7
    gR += __shfl_xor_sync(kFullM, gR, o);
    gMs += __shfl_xor_sync(kFullM, gMs, o);
  }
  const double lp = __shfl_sync(kFullM, base, 8), gm = __shfl_sync(kFullM, base, 9);
  const double q0 = __shfl_sync(kFullM, base, 10), q1 = __shfl_sync(kFullM, base, 11), q2 = __shfl_sync(kFullM, base, 12);
  if (lane == 0) {
    logp[c] = lp + lognorm;
    gmean[c] = gm;
    // reverse of get_cl
    const double u1 = star[c * XO_NSTAR + 2], u2 = star[c * XO_NSTAR + 3];
    const double c0 = 1.0 - u1 - 1.5 * u2, c1 = u1 + 2.0 * u2, c2 = -0.25 * u2;
    const double norm = xo::kPi * (c0 + c1 / 1.5), rn = 1.0 / norm;
    double gnorm = -(q0 * c0 + q1 * c1 + q2 * c2) * rn * rn;
    double gc0 = q0 * rn + gnorm * xo::kPi, gc1 = q1 * rn + gnorm * xo::kPi / 1.5, gc2 = q2 * rn;
    gstar[c * XO_NSTAR + 0] = gR;
    gstar[c * XO_NSTAR + 1] = gMs;
    gstar[c * XO_NSTAR + 2] = -gc0 + gc1;
    gstar[c * XO_NSTAR + 3] = -1.5 * gc0 + 2.0 * gc1 - 0.25 * gc2;
  }
}

struct ModelLayout {
  size_t off_elem, off_cl, off_partial, bytes;
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
  L.off_partial = take((size_t)C * L.sp.slots * (xo::kNBase + P * XO_NGRAD));
  L.bytes = o;
  return L;
}

}  // namespace

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
  const int64_t cp = C * P;
  model_prologue_kernel<<<(unsigned)((2 * cp + 63) / 64), 64, 0, st>>>(par, star, C, (int)P, texp, use_in_transit, elem, cl);
  xo::launch_transit_logp_grad(t, y, ivar, N, elem, cl, mean, C, (int)P, dtk, wk, K, L.sp, moments, partial, st);
  model_epilogue_kernel<<<(unsigned)((C + 3) / 4), 128, 0, st>>>(par, star, partial, L.sp.slots, C, (int)P, lognorm, logp,
                                                                 gpar, gstar, gmean);
  return xo::launch_status();
}

// element blocks only (prologue), for callers that want to inspect / reuse them
int xo_elements_from_params(const double *par, const double *star, int64_t C, int64_t P, double texp, int use_in_transit,
                            double *elem, double *cl, void *stream) {
  if (C < 0 || P < 0) return XO_EINVAL;
  if (P > XO_MAX_PLANETS) return XO_ERANGE;
  if (C == 0 || P == 0) return 0;
  if (!par || !star || !elem || !cl) return XO_EINVAL;
  const int64_t cp = C * P;
  model_prologue_kernel<<<(unsigned)((2 * cp + 63) / 64), 64, 0, static_cast<cudaStream_t>(stream)>>>(
      par, star, C, (int)P, texp, use_in_transit, elem, cl);
  return xo::launch_status();
}

}  // extern "C"
