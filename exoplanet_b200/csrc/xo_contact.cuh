// xo_contact.cuh -- first/last contact of the sky-projected orbit with a circle of radius L
// (the op at /root/reference/src/exoplanet/orbits/keplerian.py:744-753).
#pragma once
#include <math.h>

#include "xo_math.cuh"

namespace xo {

// g(u) = p^2 (sin^2 u + cos^2 i cos^2 u) - L^2 (1 + e cos f)^2,  f = pi/2 - w + u, |u| <= pi/2.
// Bracketed Newton (never leaves [lo, hi]), bounded iterations.  One call solves ONE side
// (dir = -1: first contact, +1: last contact) so two lanes can work on a planet concurrently.
__device__ __forceinline__ void cp_eval(double u, double p2, double L2, double e, double cw, double sw, double ci2,
                                        double &g, double &dg) {
  double su, cu, d0, d1;
  sincos_0pi(fabs(u), su, cu, d0, d1);  // |u| <= pi/2
  su = u < 0.0 ? -su : su;
  double cf = sw * cu - cw * su;
  double dcf = -sw * su - cw * cu;
  double opc = 1.0 + e * cf;
  g = p2 * (su * su + ci2 * cu * cu) - L2 * opc * opc;
  dg = 2.0 * p2 * su * cu * (1.0 - ci2) - 2.0 * L2 * opc * e * dcf;
}

// returns 0 and the mean anomaly of the contact on side `dir`, or a non-zero flag.
// kFast (the fused model's prologue, where the window only selects which cadences are evaluated
// and the light curve is exactly 0 at its edges): stop Newton as soon as the last step was below
// 3e-8 -- the iterate it produced is then within ~1e-15 |g''/g'| of the root -- and seed it with
// the two-term series of asin instead of the library call.
template <bool kFast = false>
__device__ inline int cp_solve_side(double a, double e, double cw, double sw, double ci, double L, double dir, double &Mout) {
  const double p = a * (1.0 - e) * (1.0 + e), p2 = p * p, L2 = L * L, ci2 = ci * ci;
  Mout = 0.0;
  double g, dg;
  cp_eval(0.0, p2, L2, e, cw, sw, ci2, g, dg);
  if (!(g < 0.0)) return 1;
  // circular-like estimate of the half-angle sizes the first bracket and seeds Newton
  const double opc0 = 1.0 + e * sw;
  const double arg = kFast ? (L2 * opc0 * opc0 * rcp_fast(p2) - ci2) * rcp_fast(fmax(1.0 - ci2, 1e-30))
                           : (L2 * opc0 * opc0 / p2 - ci2) / fmax(1.0 - ci2, 1e-300);
  double u0;
  if (kFast) {
    const double sq = sqrt_fast(fmax(arg, 0.0));
    u0 = (arg > 0.0 && arg < 0.25) ? sq * fma(arg, fma(arg, 0.075, 1.0 / 6.0), 1.0) : (arg > 0.0 && arg < 1.0 ? asin(sq) : 0.25 * kPi);
  } else {
    u0 = (arg > 0.0 && arg < 1.0) ? asin(sqrt(arg)) : 0.25 * kPi;
  }
  double lo = 0.0, hi = fmin(1.5 * u0 + 1e-6, 0.5 * kPi);
  bool found = false;
#pragma unroll 1
  for (int it = 0; it < 64; ++it) {
    cp_eval(dir * hi, p2, L2, e, cw, sw, ci2, g, dg);
    if (g > 0.0) {
      found = true;
      break;
    }
    if (hi >= 0.5 * kPi) break;
    lo = hi;
    hi = fmin(1.5 * hi + 1e-6, 0.5 * kPi);
  }
  if (!found) return 2;
  double x = fmin(fmax(u0, lo), hi);
#pragma unroll 1
  for (int it = 0; it < 60; ++it) {
    cp_eval(dir * x, p2, L2, e, cw, sw, ci2, g, dg);
    if (g > 0.0) hi = x; else lo = x;
    // (a zero or denormal slope gives inf / NaN here, which the bracket test below turns into a bisection)
    double xn = kFast ? fma(-g, rcp_fast(dir * dg), x) : x - g / (dir * dg);
    if (!(xn >= lo && xn <= hi)) xn = 0.5 * (lo + hi);  // (>=, <=: a converged iterate sits ON the bracket edge it just set)
    const double step = fabs(xn - x);
    x = xn;
    if (step <= (kFast ? 3e-8 : 1e-15) * fabs(x) + 1e-300 || hi - lo <= 4e-16 * hi) break;
  }
  double su, cu, d0, d1;
  sincos_0pi(x, su, cu, d0, d1);
  su *= dir;
  const double cf = sw * cu - cw * su;  // cos(f0 + u), cos f0 = sin w, sin f0 = cos w
  const double sf = cw * cu + sw * su;
  const double den = 1.0 + e * cf;
  if (kFast) {
    const double rden = rcp_fast(den);
    const double sE = sqrt_fast((1.0 - e) * (1.0 + e)) * sf * rden, cE = (e + cf) * rden;
    const double E = atan2_pos(fabs(sE), cE);  // |E| in [0, pi]
    Mout = (sE < 0.0 ? -E : E) - e * sE;
  } else {
    const double sE = sqrt((1.0 - e) * (1.0 + e)) * sf / den, cE = (e + cf) / den;
    Mout = atan2(sE, cE) - e * sE;
  }
  return 0;
}

__device__ inline int cp_solve(double a, double e, double cw, double sw, double ci, double L, double &Ml, double &Mr) {
  int f0 = cp_solve_side<false>(a, e, cw, sw, ci, L, -1.0, Ml);
  int f1 = cp_solve_side<false>(a, e, cw, sw, ci, L, 1.0, Mr);
  if (f0 | f1) Ml = Mr = 0.0;
  return f0 ? f0 : f1;
}

}  // namespace xo
