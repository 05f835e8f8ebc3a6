// xo_contact.cuh -- first/last contact of the sky-projected orbit with a circle of radius L
// (the op at /root/reference/src/exoplanet/orbits/keplerian.py:744-753).
#pragma once
#include <math.h>

#include "xo_math.cuh"

namespace xo {

// g(u) = p^2 (sin^2 u + cos^2 i cos^2 u) - L^2 (1 + e cos f)^2,  f = pi/2 - w + u.
// One thread per planet; bracketed Newton (never leaves [lo, hi]), bounded iterations.
__device__ __forceinline__ void cp_eval(double u, double p2, double L2, double e, double cw, double sw, double ci2,
                                        double &g, double &dg) {
  double su, cu;
  sincos(u, &su, &cu);
  double cf = sw * cu - cw * su;
  double dcf = -sw * su - cw * cu;
  double opc = 1.0 + e * cf;
  g = p2 * (su * su + ci2 * cu * cu) - L2 * opc * opc;
  dg = 2.0 * p2 * su * cu * (1.0 - ci2) - 2.0 * L2 * opc * e * dcf;
}

__device__ inline int cp_solve(double a, double e, double cw, double sw, double ci, double L, double &Ml, double &Mr) {
  const double p = a * (1.0 - e) * (1.0 + e), p2 = p * p, L2 = L * L, ci2 = ci * ci;
  Ml = Mr = 0.0;
  double g0, dg0;
  cp_eval(0.0, p2, L2, e, cw, sw, ci2, g0, dg0);
  if (!(g0 < 0.0)) return 1;
  // circular-like estimate of the half-angle, used to size the first bracket
  double opc0 = 1.0 + e * sw;
  double arg = (L2 * opc0 * opc0 / p2 - ci2) / fmax(1.0 - ci2, 1e-300);
  double u0 = (arg > 0.0 && arg < 1.0) ? asin(sqrt(arg)) : 0.25 * xo::kPi;
  double root[2];
#pragma unroll 1
  for (int side = 0; side < 2; ++side) {
    const double dir = side ? 1.0 : -1.0;
    double lo = 0.0, hi = fmin(1.5 * u0 + 1e-6, 0.5 * xo::kPi);
    double g, dg;
    bool found = false;
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
      cp_eval(dir * hi, p2, L2, e, cw, sw, ci2, g, dg);
      if (g > 0.0) {
        found = true;
        break;
      }
      if (hi >= 0.5 * xo::kPi) break;
      lo = hi;
      hi = fmin(1.5 * hi + 1e-6, 0.5 * xo::kPi);
    }
    if (!found) return 2;
    double x = 0.5 * (lo + hi);
#pragma unroll 1
    for (int it = 0; it < 100; ++it) {
      cp_eval(dir * x, p2, L2, e, cw, sw, ci2, g, dg);
      if (g > 0.0) hi = x; else lo = x;
      double dx = -g / (dir * dg);
      double xn = x + dx;
      if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
      if (fabs(xn - x) <= 4e-16 * fabs(x) || hi - lo <= 4e-16 * hi) {
        x = xn;
        break;
      }
      x = xn;
    }
    root[side] = dir * x;
  }
#pragma unroll 1
  for (int side = 0; side < 2; ++side) {
    double su, cu;
    sincos(root[side], &su, &cu);
    double cf = sw * cu - cw * su;  // cos(f0 + u), cos f0 = sin w, sin f0 = cos w
    double sf = cw * cu + sw * su;
    double den = 1.0 + e * cf;
    double sE = sqrt((1.0 - e) * (1.0 + e)) * sf / den, cE = (e + cf) / den;
    double M = atan2(sE, cE) - e * sE;
    if (side) Mr = M; else Ml = M;
  }
  return 0;
}


}  // namespace xo
