// xo_math.cuh -- per-cadence fp64 math of the hot path, written for sm_100a.
//
// Everything here is a __host__ __device__ inline so the *same* arithmetic can be
// compiled into a CPU harness by the unit tests (tests/hostcheck) -- that harness is
// test infrastructure, not a fallback: the product library only launches kernels.
//
// What each routine restates (reference file:line are under /root/reference):
//   kepler_solve   -- the op at src/exoplanet/orbits/keplerian.py:333 (exoplanet-core
//                     `kepler`): Markley (1995) cubic starter + one 5th-order correction,
//                     reorganised for the GPU: no data-dependent loop, one sincos, the
//                     corrected sin/cos by angle addition, 4 reciprocals, cancellation-
//                     free E - sin E / 1 - cos E so the e->1, M->0 corner stays accurate.
//   qsv            -- the op at src/exoplanet/light_curves/limb_dark.py:24
//                     (`quad_solution_vector`): Agol, Luger & Foreman-Mackey (2020)
//                     uniform / linear / quadratic terms and their b, r derivatives, with
//                     three Bulirsch cel integrals advanced in lock-step on one AGM ladder.
//   lc_point       -- the PyTensor glue between them: keplerian.py:324-334 (mean anomaly),
//                     :283-322,:400-409 (radius, rotation), limb_dark.py:216-252.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define XO_HD __host__ __device__ __forceinline__
#else
#define XO_HD inline
#endif

namespace xo {

constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double kTwoPiHi = 6.283185307179586232e+00;   // fp64(2 pi)
constexpr double kTwoPiLo = 2.449293598294706414e-16;   // 2 pi - fp64(2 pi)
constexpr double kInvTwoPi = 0.159154943091895335768883763372514362;
constexpr double kPio2Hi = 1.570796326794896558e+00;
constexpr double kPio2Lo = 6.123233995736766036e-17;
constexpr double kTwoOverPi = 0.636619772367581343075535053490057448;

// fp64 constants that are not encodable as instruction immediates live in ONE table in
// __constant__ memory so that DFMA/DMUL read them as c[bank][offset] operands; otherwise ptxas
// rebuilds every 64-bit literal with two MOVs per use (~80 extra instructions per Kepler solve).
#define XO_CONST_TABLE                                                                                         \
  {                                                                                                            \
    /* 0..6  sin kernel S(u), highest degree first */                                                          \
    -7.58426609767779964718e-13, 1.60584895500232502802e-10, -2.50521059616691626032e-8,                       \
        0.0000027557319218543807191, -0.000198412698412640218792, 0.00833333333333333099958,                  \
        -0.166666666666666666651, /* 7..13 cos kernel C(u) */                                                  \
        4.74452016112717801505e-14, -1.1470437495511350934e-11, 2.08767556405153041563e-9,                     \
        -2.75573192209605647402e-7, 0.0000248015873015840664244, -0.00138888888888888875915,                  \
        0.0416666666666666666658, /* 14..19 range reduction */                                                 \
        0.159154943091895335768883763372514362, 6.283185307179586232e+00, 2.449293598294706414e-16,            \
        0.636619772367581343075535053490057448, 1.570796326794896558e+00, 6.123233995736766036e-17,            \
        /* 20..25 small-angle sin/cos series: -1/5040, 1/120, -1/6, -1/720, 1/24, pi */                        \
        -1.0 / 5040.0, 1.0 / 120.0, -1.0 / 6.0, -1.0 / 720.0, 1.0 / 24.0, 3.14159265358979323846264338327950288, \
        /* 26..31 solution-vector constants: 2pi/3, 1/(2pi), 4pi, 2pi, 1/(9pi), 1/3 */                         \
        2.0 * kPi / 3.0, 0.5 / kPi, 4.0 * kPi, 2.0 * kPi, 1.0 / (9.0 * kPi), 1.0 / 3.0, /* fp64-folded like the oracle */ \
        /* 32..42 atan kernel A(u), LOWEST degree first (tools/gen_poly.py); 43 pi/4; 44 tan(pi/8) */            \
        -0.333333333333333295843, 0.199999999999947778464, -0.14285714284510125851, 0.111111110025998437107, -0.0909090406207583984612, 0.0769217118408174259129, -0.066643423247421068204, 0.0585668160968349331121, -0.0507778892612066834491, 0.0390110323660379116755, -0.0189077255845882088276, \
        0.785398163397448309615660845819875721, 0.414213562373095048801688724209698079                         \
  }
namespace detail {
#if defined(__CUDACC__)
static __constant__ double c_tab[45] = XO_CONST_TABLE;
#endif
static const double h_tab[45] = XO_CONST_TABLE;
}  // namespace detail
#if defined(__CUDA_ARCH__)
#define XO_K(i) (::xo::detail::c_tab[i])
#else
#define XO_K(i) (::xo::detail::h_tab[i])
#endif

XO_HD double rcp(double x) {
#if defined(__CUDA_ARCH__)
  return __drcp_rn(x);
#else
  return 1.0 / x;
#endif
}

// ~1 ulp reciprocal without the IEEE fix-up paths of __drcp_rn: MUFU.RCP64H seed
// (rcp.approx.ftz.f64, ~20 bits) + two Newton steps.  x must be a normal, finite, non-zero
// double (all call sites divide by quantities bounded away from 0 and inf).
XO_HD double rcp_fast(double x) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
#else
  return 1.0 / x;
#endif
}

// reciprocal to ~1e-12 relative (seed + one Newton step): enough where the quotient is a
// small correction term
XO_HD double rcp_mid(double x) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  return fma(r, fma(e, e, e), r);
#else
  return 1.0 / x;
#endif
}

// ~1 ulp square root of a normal positive double: MUFU.RSQ64H seed + two coupled
// Newton (Goldschmidt) steps.  sqrt_fast(0) = 0.
XO_HD double sqrt_fast(double x) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double g = x * r, h = 0.5 * r;
  double e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  return x > 0.0 ? g : 0.0;
#else
  return sqrt(x);
#endif
}

// ---------------------------------------------------------------------------
// sin/cos on [0, pi] that also return E - sin E and 1 - cos E free of cancellation.
// Polynomials from tools/gen_poly.py (|x| <= 1.02 pi/4; abs err 5e-18 / 1e-18).
// ---------------------------------------------------------------------------
XO_HD void sincos_0pi(double E, double &s, double &c, double &EmS, double &omC) {
  double qf = rint(E * XO_K(17));  // 0, 1 or 2
  double x = fma(-qf, XO_K(18), E);
  x = fma(-qf, XO_K(19), x);
  double u = x * x;
  double ps = XO_K(0);
  ps = fma(ps, u, XO_K(1));
  ps = fma(ps, u, XO_K(2));
  ps = fma(ps, u, XO_K(3));
  ps = fma(ps, u, XO_K(4));
  ps = fma(ps, u, XO_K(5));
  ps = fma(ps, u, XO_K(6));
  double pc = XO_K(7);
  pc = fma(pc, u, XO_K(8));
  pc = fma(pc, u, XO_K(9));
  pc = fma(pc, u, XO_K(10));
  pc = fma(pc, u, XO_K(11));
  pc = fma(pc, u, XO_K(12));
  pc = fma(pc, u, XO_K(13));
  double xs = x * u * ps;               // sin x - x
  double cm1 = fma(u * u, pc, -0.5 * u);  // cos x - 1
  double sx = x + xs, cx = 1.0 + cm1;
  if (qf == 0.0) {
    s = sx;
    c = cx;
    EmS = -xs;
    omC = -cm1;
  } else if (qf == 1.0) {  // E = pi/2 + x
    s = cx;
    c = -sx;
    EmS = E - cx;
    omC = 1.0 + sx;
  } else {  // E = pi + x
    s = -sx;
    c = -cx;
    EmS = E + sx;
    omC = 2.0 + cm1;
  }
}

// full-range sincos for the circular branch (keplerian.py:331-332): M already reduced
// to [-pi, pi].
XO_HD void sincos_reduced(double Mr, double &s, double &c) {
  double a = fabs(Mr), ss, cc, d0, d1;
  sincos_0pi(fmin(a, kPi), ss, cc, d0, d1);
  s = Mr < 0.0 ? -ss : ss;
  c = cc;
}

XO_HD double reduce_2pi(double M) {
  double k = rint(M * XO_K(14));
  double r = fma(-k, XO_K(15), M);
  return fma(-k, XO_K(16), r);
}

// ---------------------------------------------------------------------------
// Kepler: (M, e) -> (sin f, cos f).  0 <= e < 1.
//
// The Markley (1995) starter only has to land within ~3e-4 (relative) of the root -- the
// 5th-order correction that follows squares-and-more that away -- so it is evaluated in
// fp32 on the FMA/MUFU pipes, which run beside the fp64 pipe that bounds this kernel.
// 1 - e is formed in fp64 first so the e -> 1 corner keeps its relative accuracy.
// ---------------------------------------------------------------------------
// single-instruction fp32 approximations (flush-to-zero, no denormal fix-up code)
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float rcp_f32(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rsqrt_f32(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lg2_f32(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_f32(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
#endif

XO_HD double markley_start_f32(double Ma, double e, double ome) {
  const float M = (float)Ma, ef = (float)e, om = (float)ome;  // F2F: 4 issue slots; bit surgery costs ~24 (measured slower)
  constexpr float fac = (float)(1.0 / (kPi - 6.0 / kPi));
  constexpr float pif = (float)kPi;
#if defined(__CUDA_ARCH__)
  const float rope = rcp_f32(1.0f + ef);
#else
  const float rope = 1.0f / (1.0f + ef);
#endif
  const float alpha = fmaf(1.6f * fac * rope, pif - M, 3.0f * pif * fac);
  const float d = fmaf(alpha, ef, 3.0f * om);
  const float ad = alpha * d;
  const float M2 = M * M;
  const float r = fmaf(3.0f * ad, d - om, M2) * M;
  const float q = fmaf(2.0f * ad, om, -M2);
  const float q2 = q * q;
  const float disc = fmaxf(fmaf(q2, q, r * r), 1e-37f);
#if defined(__CUDA_ARCH__)
  const float y = fabsf(r) + disc * rsqrt_f32(disc);
  float w = ex2_f32((2.0f / 3.0f) * lg2_f32(y));  // cbrt(y)^2
#else
  const float y = fabsf(r) + sqrtf(disc);
  float w = cbrtf(y);
  w = w * w;
#endif
  const float W = fmaf(w, w + q, q2);
  const float num = fmaf(2.0f * r, w, M * W), den = d * W;
#if defined(__CUDA_ARCH__)
  float E = num * rcp_f32(den);
#else
  float E = num / den;
#endif
  if (!(den > 0.0f) || !(E >= 0.0f)) E = M;  // underflow guard (M absurdly small): Newton recovers
  return (double)E;
}

XO_HD void kepler_solve(double M, double e, double &sinf, double &cosf) {
  double Mr = reduce_2pi(M);
  double Ma = fmin(fabs(Mr), XO_K(25));
  double ome = 1.0 - e;
  double ope = 1.0 + e;
  double E = markley_start_f32(Ma, e, ome);
  // one 5th-order correction: root of the quartic Taylor polynomial in dE,
  // f0 + f1 dE + f2 dE^2/2 + f3 dE^3/6 - f2 dE^4/24 = 0, by fixed-point on 1/f1
  double s, c, EmS, omC;
  sincos_0pi(E, s, c, EmS, omC);
  double f0 = fma(e, EmS, fma(E, ome, -Ma));
  double f1 = fma(e, omC, ome);
  double f2 = e * s;
  double f3 = e * c;
  double rf1 = rcp_mid(f1);
  double h2 = 0.5 * f2, h3 = -f3 * XO_K(22), h4 = -f2 * XO_K(24);
  double dE = -f0 * rf1;
#pragma unroll
  for (int it = 0; it < 3; ++it) dE = -fma(dE * dE, fma(dE, fma(dE, h4, h3), h2), f0) * rf1;
  // sin/cos(E + dE) by angle addition (|dE| < ~1e-2)
  double d2 = dE * dE;
  double sd = dE * fma(d2, fma(d2, fma(d2, XO_K(20), XO_K(21)), XO_K(22)), 1.0);
  double cdm1 = d2 * fma(d2, fma(d2, XO_K(23), XO_K(24)), -0.5);
  double s1 = s + fma(c, sd, s * cdm1);
  double omC1 = omC - fma(c, cdm1, -s * sd);
  double den = fma(e, omC1, ome);  // 1 - e cos E
  double rden = rcp_fast(den);
  double sq = sqrt_fast(ome * ope);
  const double nanprop = (M - M) + (e - e);  // NaN / inf in -> NaN out (fmin/fmax above drop NaNs)
  sinf = fma((Mr < 0.0 ? -sq : sq) * s1, rden, nanprop);
  cosf = fma(ome - omC1, rden, nanprop);  // (cos E - e) without cancellation
}

// VJP factors of the op
XO_HD void kepler_jac(double sinf, double cosf, double e, double &dfdM, double &dfde) {
  double ome2 = (1.0 - e) * (1.0 + e);
  double opc = fma(e, cosf, 1.0);
  double r = rcp_fast(ome2);
  dfdM = opc * opc * r * sqrt_fast(r);
  dfde = (2.0 + e * cosf) * sinf * r;
}

// atan2(y, x) for y >= 0 (result in [0, pi]): two argument reductions to |s| <= tan(pi/8), then
// s + s^3 A(s^2).  Replaces the general-purpose library atan2 (about 4x the instructions, half of
// them constant moves) on the ingress/egress branch of the solution vector.
XO_HD double atan2_pos(double y, double x) {
  const double ax = fabs(x);
  const double mn = fmin(y, ax), mx = fmax(y, ax);
  if (!(mx > 0.0)) return 0.0;
  const bool big = mn > XO_K(44) * mx;
  const double s = (big ? mn - mx : mn) * rcp_fast(big ? mn + mx : mx);
  const double u = s * s;
  double p = XO_K(42);
  p = fma(p, u, XO_K(41));
  p = fma(p, u, XO_K(40));
  p = fma(p, u, XO_K(39));
  p = fma(p, u, XO_K(38));
  p = fma(p, u, XO_K(37));
  p = fma(p, u, XO_K(36));
  p = fma(p, u, XO_K(35));
  p = fma(p, u, XO_K(34));
  p = fma(p, u, XO_K(33));
  p = fma(p, u, XO_K(32));
  double r = fma(s * u, p, s);
  if (big) r += XO_K(43);
  if (y > ax) r = 2.0 * XO_K(43) - r;
  if (x < 0.0) r = XO_K(25) - r;
  return r;
}

// ---------------------------------------------------------------------------
// Three Bulirsch (1969) cel integrals on one AGM ladder:
//   G  = cel(kc, p,  a, b)   (p > 0)
//   Ek = cel(kc, 1,  1, kc^2) = E(m)
//   Em = cel(kc, 1,  1, 0)    = (E(m) - (1-m) K(m)) / m
// ---------------------------------------------------------------------------
XO_HD void cel3(double kc, double kc2, double p, double a, double b, double &G, double &Ek, double &Em) {
  constexpr double CA = 1.0e-8;  // quadratic convergence: error ~ CA^2
  double e = kc, em = 1.0;
  p = sqrt_fast(p);
  b *= rcp_fast(p);
  double p1 = 1.0, aE = 1.0, bE = kc2, aM = 1.0, bM = 0.0;
#pragma unroll 1
  for (int it = 0; it < 12; ++it) {
    double rp = rcp_fast(p), rp1 = rcp_fast(p1);
    double f = a;
    a = fma(b, rp, a);
    double g = e * rp;
    b = fma(f, g, b);
    b += b;
    p += g;
    double g1 = e * rp1;
    f = aE;
    aE = fma(bE, rp1, aE);
    bE = fma(f, g1, bE);
    bE += bE;
    f = aM;
    aM = fma(bM, rp1, aM);
    bM = fma(f, g1, bM);
    bM += bM;
    p1 += g1;
    double gp = em;
    em += kc;
    if (fabs(gp - kc) <= gp * CA) break;
    kc = sqrt_fast(e);
    kc += kc;
    e = kc * em;
  }
  double t = 0.5 * XO_K(25) * rcp_fast(em);
  G = t * fma(a, em, b) * rcp_fast(em + p);
  double r1 = rcp_fast(em + p1);
  Ek = t * fma(aE, em, bE) * r1;
  Em = t * fma(aM, em, bM) * r1;
}

// ---------------------------------------------------------------------------
// The measure-zero configurations of quad_solution_vector (b = 0, b = r, b = 1 - r) in their own NON-inlined functions.
// Inlined, their r-only subexpressions (divisions, acos, sqrt) are loop-invariant in the sweep of the logp kernel -- r is the
// chain's radius ratio -- and the compiler hoists them in front of every drain: ~360 instructions each time, 11% of that
// kernel's instruction count (profiles/r02_transit.txt, "line 440").  Results are returned by value (registers).
// ---------------------------------------------------------------------------
#if defined(__CUDACC__)
#define XO_RARE static __host__ __device__ __noinline__
#else
#define XO_RARE static
#endif
enum QsvKind { kB0, kHalf, kEqSmall, kEqLarge, kTangent, kLimb, kInside };
struct QsvEq {
  int kind;
  double kc2, aux;
};
XO_RARE QsvEq qsv_eq_kind(double r, double r2) {  // b == r: which closed form, and the modulus of its elliptic integrals
  QsvEq o{kHalf, 1.0, 0.0};
  if (r == 0.5) return o;
  if (r < 0.5) {
    o.kind = kEqSmall;
    o.kc2 = 1.0 - 4.0 * r2;  // m = 4 r^2; G = K(m)
  } else {
    o.kind = kEqLarge;
    o.aux = 0.25 / r2;  // m = 1/(4 r^2)
    o.kc2 = 1.0 - o.aux;
  }
  return o;
}
struct QsvLam {
  double Lam, d1b, d1r;
};
XO_RARE QsvLam qsv_rare(int kind, double r, double r2, double aux, double onembmr2, double G, double E, double Em) {
  QsvLam o;
  if (kind == kB0) {
    const double x = sqrt(fmax(0.0, 1.0 - r2));
    o.Lam = -2.0 / 3.0 * x * x * x;
    o.d1b = 0.0;
    o.d1r = -2.0 * kPi * r * x;
  } else if (kind == kTangent) {
    const double sqA = sqrt(onembmr2);
    o.Lam = 2.0 / (3.0 * kPi) * acos(1.0 - 2.0 * r) - 4.0 / (9.0 * kPi) * (3.0 + 2.0 * r - 8.0 * r2) * sqrt(r * (1.0 - r)) -
            (r > 0.5 ? 2.0 / 3.0 : 0.0);
    o.d1r = -4.0 * r * sqA;
    o.d1b = 4.0 * r * sqA / 3.0;
  } else if (kind == kHalf) {
    o.Lam = 1.0 / 3.0 - 4.0 / (9.0 * kPi);
    o.d1r = -4.0 * r;
    o.d1b = 4.0 * r / 3.0;
  } else if (kind == kEqSmall) {
    o.Lam = 1.0 / 3.0 + 2.0 / (9.0 * kPi) * (4.0 * (2.0 * r2 - 1.0) * E + (1.0 - 4.0 * r2) * G);
    o.d1r = -4.0 * r * E;
    o.d1b = 4.0 * r / 3.0 * (2.0 * Em - E);
  } else {  // kEqLarge
    o.Lam = 1.0 / 3.0 + 16.0 * r / (9.0 * kPi) * (2.0 * r2 - 1.0) * E - (1.0 - 4.0 * r2) * (3.0 - 8.0 * r2) / (9.0 * kPi * r) * G;
    o.d1r = -8.0 * r2 * aux * Em;
    o.d1b = 8.0 * r2 / 3.0 * aux * (2.0 * E - Em);
  }
  return o;
}

// ---------------------------------------------------------------------------
// quad_solution_vector(b, r): s[3] and (optionally) ds/db, ds/dr.
// b may be negative: |b| is used and d/db carries sign(b).
// ---------------------------------------------------------------------------
template <bool kJac>
XO_HD void qsv(double b_in, double r, double *s, double *dsdb, double *dsdr) {
  double b = fabs(b_in);
  if (kJac) {
    dsdb[0] = dsdb[1] = dsdb[2] = 0.0;
    dsdr[0] = dsdr[1] = dsdr[2] = 0.0;
  }
  // the four kite factors; the one subtraction in each that can cancel is exact
  double t1 = 1.0 + b + r;
  double t2 = (r <= 2.0) ? (1.0 - r) + b : 1.0 + (b - r);  // 1 + b - r
  double t3 = (b <= 2.0) ? (1.0 - b) + r : 1.0 + (r - b);  // 1 - b + r
  double t4 = (b >= 0.5) ? (b - 1.0) + r : (r - 1.0) + b;  // b + r - 1
  if (!(r > 0.0) || !(t3 > 0.0)) {                         // unocculted (r <= 0, b >= 1 + r, NaN)
    bool bad = (b != b) || (r != r);
    s[0] = bad ? NAN : XO_K(25);
    s[1] = bad ? NAN : XO_K(26);
    s[2] = bad ? NAN : 0.0;
    return;
  }
  if (t2 <= 0.0) {  // fully occulted
    s[0] = s[1] = s[2] = 0.0;
    return;
  }
  double r2 = r * r, b2 = b * b;
  double bmr = b - r, bpr = b + r;
  double onembmr2 = t2 * t3;   // 1 - (b-r)^2
  double onembpr2 = -t4 * t1;  // 1 - (b+r)^2
  double fourbr = 4.0 * b * r;
  double kap0, sk0, ck0, s0, eta;
  if (t4 <= 0.0) {  // occultor entirely inside the disk
    kap0 = XO_K(25);
    sk0 = 0.0;
    ck0 = -1.0;
    s0 = XO_K(25) * (1.0 - r2);
    eta = 0.5 * r2 * fma(2.0, b2, r2);
  } else {
    double D = sqrt_fast(t1 * t2 * t3 * t4);  // twice the kite area
    double c0 = fma(r - 1.0, r + 1.0, b2);
    double c1 = fma(1.0 - r, 1.0 + r, b2);
    kap0 = atan2_pos(D, c0);
    double kap1 = atan2_pos(D, c1);
    double rtwobr = rcp_fast(2.0 * b * r);
    sk0 = D * rtwobr;
    ck0 = c0 * rtwobr;
    s0 = XO_K(25) - (kap1 + r2 * kap0 - 0.5 * D);
    eta = (kap1 + r2 * fma(2.0, b2, r2) * kap0 - 0.25 * (1.0 + 5.0 * r2 + b2) * D) * XO_K(27);
  }
  s[0] = s0;
  s[2] = fma(XO_K(28), eta, 2.0 * s0) - XO_K(29);
  if (kJac) {
    double gg = 2.0 - 4.0 * b2 - 4.0 * r2;
    dsdb[0] = 2.0 * r * sk0;
    dsdr[0] = -2.0 * r * kap0;
    dsdr[2] = -r * (2.0 * gg * kap0 + 16.0 * b * r * sk0);
    dsdb[2] = r * (2.0 * gg * sk0 + 8.0 * b * r * (kap0 + sk0 * ck0));
  }
  // linear term  s1 = 2pi/3 (1 - Theta(r - b)) - pi Lambda.  Every regime that needs elliptic
  // integrals first reduces to ONE cel3 call (so a warp whose lanes sit in different regimes
  // walks the AGM ladder once, together), then finishes with its own closed form.
  int kind;
  double kc2 = 1.0, cp = 1.0, ca = 1.0, cb = 1.0;  // cel(kc, p, a, b) arguments
  double aux = 0.0, qq = 0.0;
  if (b == 0.0) {
    kind = kB0;
  } else if (b == r) {
    const QsvEq eq = qsv_eq_kind(r, r2);
    kind = eq.kind;
    kc2 = eq.kc2;
    aux = eq.aux;
  } else if (t4 == 0.0) {
    kind = kTangent;  // b = 1 - r
  } else if (t4 > 0.0) {
    kind = kLimb;  // k^2 < 1: the limb crosses the occultor
    kc2 = -onembpr2 * rcp_fast(fourbr);
    cp = bmr * bmr * kc2;
    ca = 0.0;
    cb = 3.0 * kc2 * bmr * bpr;
  } else {
    kind = kInside;  // k^2 > 1
    const double rA = rcp_fast(onembmr2);
    kc2 = onembpr2 * rA;
    qq = bmr * rcp_fast(bpr);
    const double mu = 3.0 * qq * rA;
    cp = qq * qq * kc2;
    ca = 1.0 + mu;
    cb = cp + mu;
  }
  double G = 0.0, E = 0.0, Em = 0.0;
  if (kind == kEqSmall || kind == kEqLarge || kind >= kLimb) cel3(sqrt_fast(kc2), kc2, cp, ca, cb, G, E, Em);
  double Lam, d1b, d1r;
  if (kind == kLimb) {
    const double rsqbr = rcp_fast(sqrt_fast(b * r));
    Lam = onembmr2 * (G + fma(-3.0 + 6.0 * r2 + 2.0 * b * r, Em, -fourbr * E)) * rsqbr * XO_K(30);
    d1r = -2.0 * r * onembmr2 * Em * rsqbr;
    d1b = 2.0 * r * onembmr2 * (2.0 * E - Em) * rsqbr * XO_K(31);
  } else if (kind == kInside) {
    const double sqA = sqrt_fast(onembmr2);
    Lam = 2.0 * sqA * fma(onembpr2, G, -(4.0 - 7.0 * r2 - b2) * E) * XO_K(30);
    d1r = -4.0 * r * sqA * E;
    d1b = 4.0 * r * sqA * XO_K(31) * (2.0 * Em - E);
  } else {  // b = 0, b = r, b = 1 - r
    const QsvLam o = qsv_rare(kind, r, r2, aux, onembmr2, G, E, Em);
    Lam = o.Lam;
    d1b = o.d1b;
    d1r = o.d1r;
  }
  s[1] = fma(-XO_K(25), Lam, r > b ? 0.0 : XO_K(26));
  if (kJac) {
    dsdb[1] = d1b;
    dsdr[1] = d1r;
    if (b_in < 0.0) {
      dsdb[0] = -dsdb[0];
      dsdb[1] = -dsdb[1];
      dsdb[2] = -dsdb[2];
    }
  }
}

// ---------------------------------------------------------------------------
// Element block (include/xo_b200.h; XO_EL_* indices) and the per-sample evaluation.
// `el` usually points into shared memory (one block per chain and planet).
// ---------------------------------------------------------------------------
// dt = mod(t - t0 + P/2, P) - P/2 in [tstart, tend]   (keplerian.py:731,769)
XO_HD bool in_window(const double *el, double rperiod, double t) {
  double period = el[10];
  double hp = 0.5 * period;
  double x = (t - el[0]) + hp;
  double dt = fma(-period, floor(x * rperiod), x) - hp;
  return (dt >= el[11]) & (dt <= el[12]);
}

// true anomaly at time tt (keplerian.py:324-334); M = ((t - t0) - tref) n, in that order
template <typename EP>
XO_HD void true_anomaly(EP el, double tt, double &sf, double &cf, double &tm) {
  tm = (tt - el[0]) - el[1];
  double M = tm * el[2];
  double e = el[3];
  if (e == 0.0) {
    sincos_reduced(reduce_2pi(M), sf, cf);
  } else {
    kepler_solve(M, e, sf, cf);
  }
}

// One light-curve sample: returns lc = s(b, ror) . cl - 1 for z > 0, else exactly 0
// (limb_dark.py:216-252).  With kGrad, ACCUMULATES w * d lc/d(t0,tref,n,e,cosw,sinw,cosi,
// sini,aor,ror) into g[10] and w * s into gs[3] (= w * d lc / d cl).
// kStride: element stride of g / gs (1 for plain arrays; blockDim for per-thread columns of a
// shared-memory matrix, which keeps the 13 accumulators out of registers during the heavy math).
// Light-travel delay (keplerian.py:411-470, _get_retarded_position with a -> -a, z0 = 0): the body is seen where it was
// at the retarded time t - D, with D from its position, line-of-sight velocity and acceleration at t:
//   z = R rr si swf,  vz = -n R aor si (e cw + cwf) / sqrt(1 - e^2),  az = -n^2 ((1 + e cf) / (1 - e^2))^3 z   [R_sun, days]
//   D = (c/az) (u - sqrt(u^2 + 2 az z / c^2)), u = 1 + vz/c       (|az| < 1e-10:  D = -z / (c + vz))
// R = r_star is element XO_EL_RSTAR (slot 14; 0 = no light delay).  kGrad: also d D / d (t0, tref, n, e, cw, sw, si, aor)
// into dD[0..7] (si at index 6, aor at 7) -- the chain rule of the second solve needs them.
constexpr double kCLight = 37231.66360672704;  // R_sun / day (orbits/constants.py:36)
template <bool kGrad>
XO_HD double light_delay(const double *el, double tt, double R, double *dD) {
  double sf, cf, tm;
  true_anomaly(el, tt, sf, cf, tm);
  const double n = el[2], e = el[3], cw = el[4], sw = el[5], si = el[7], aor = el[8];
  const double ome2 = (1.0 - e) * (1.0 + e), opc = fma(e, cf, 1.0);
  const double rso = 1.0 / sqrt(ome2);
  const double rr = aor * ome2 / opc;
  const double cwf = cw * cf - sw * sf, swf = sw * cf + cw * sf;
  const double z = R * rr * si * swf;
  const double kv = -n * R * aor * si * rso;  // d vz / d (e cw + cwf)
  const double vz = kv * fma(e, cw, cwf);
  const double q = opc / ome2;
  const double az = -n * n * q * q * q * z;
  const double rc = 1.0 / kCLight, u = fma(vz, rc, 1.0);
  double D, Dz, Dvz, Daz;
  if (fabs(az) < 1.0e-10) {
    const double rden = 1.0 / (kCLight + vz);
    D = -z * rden;
    Dz = -rden;
    Dvz = z * rden * rden;
    Daz = 0.0;
  } else {
    const double s = sqrt(fma(u, u, 2.0 * az * z * rc * rc));
    D = (kCLight / az) * (u - s);
    Dz = -rc / s;
    Dvz = (1.0 - u / s) / az;
    Daz = -D / az - z * rc / (az * s);
  }
  if (kGrad) {
    const double bz = Dz - Daz * n * n * q * q * q, bvz = Dvz, bq = Daz * 3.0 * az / q;
    const double brr = bz * R * si * swf, bswf = bz * R * si * rr, bcwf = bvz * kv;
    const double bcf = -brr * rr * e / opc + bswf * sw + bcwf * cw + bq * e / ome2;
    const double bsf = bswf * cw - bcwf * sw;
    double be = brr * aor * (-2.0 * e * opc - ome2 * cf) / (opc * opc) + bvz * (kv * cw + vz * e / ome2) +
                bq * (cf * ome2 + 2.0 * e * opc) / (ome2 * ome2);
    const double dfdM = opc * opc * rso * rso * rso, dfde = (2.0 + e * cf) * sf / ome2;
    const double bf = bsf * cf - bcf * sf;
    const double bM = bf * dfdM;
    be += bf * dfde;
    dD[0] = dD[1] = -bM * n;
    dD[2] = bM * tm + bvz * vz / n + Daz * 2.0 * az / n;
    dD[3] = be;
    dD[4] = bswf * sf + bcwf * cf + bvz * kv * e;
    dD[5] = bswf * cf - bcwf * sf;
    dD[6] = bz * R * rr * swf + bvz * vz / si;
    dD[7] = brr * rr / aor + bvz * vz / aor;
  }
  return D;
}

// EP: the element pointer's type.  `const volatile double *` makes every use of an element a fresh (broadcast) shared-memory
// load instead of a value the compiler keeps in a register across the whole evaluation: ~20 fewer live registers in the
// logp kernel, which is what lets it fit 128 (xo_transit.cu, XO_ELEM_VOLATILE).
template <bool kGrad, int kStride = 1, bool kDelay = false, typename EP = const double *>
XO_HD double lc_point(EP el, const double *cl, double tt, double w, double *g, double *gs) {
  double dD[8];
  bool delayed = false;
  if (kDelay && el[14] > 0.0) {  // XO_EL_RSTAR: light-travel delay on
    tt -= light_delay<kGrad>(const_cast<const double *>(el), tt, el[14], dD);
    delayed = true;
  }
  double sf, cf, tm;
  true_anomaly(el, tt, sf, cf, tm);
  const double e = el[3], cw = el[4], sw = el[5], ci = el[6], si = el[7], aor = el[8], ror = el[9];
  double ome2 = (1.0 - e) * (1.0 + e);
  double opc = fma(e, cf, 1.0);
  double ropc = rcp_fast(opc);
  double rr = aor * ome2 * ropc;  // |r| / R_star   (keplerian.py:403 with a -> -a)
  double cwf = cw * cf - sw * sf;
  double swf = sw * cf + cw * sf;
  double z = rr * si * swf;  // Z / R_star > 0 <=> in front of the star (limb_dark.py:252)
  if (!(z > 0.0)) return 0.0;
  double ci2 = ci * ci;
  double h = fma(cwf, cwf, ci2 * swf * swf);
  double sh = sqrt_fast(h);
  double b = rr * sh;
  if (b >= 1.0 + ror) return 0.0;
  double S[3], B[3], R[3];
  qsv<kGrad>(b, ror, S, B, R);
  double lc = fma(S[0], cl[0], fma(S[1], cl[1], S[2] * cl[2])) - 1.0;
  if (!kGrad) return lc;
  gs[0] = fma(w, S[0], gs[0]);
  gs[kStride] = fma(w, S[1], gs[kStride]);
  gs[2 * kStride] = fma(w, S[2], gs[2 * kStride]);
  double dldb = w * fma(B[0], cl[0], fma(B[1], cl[1], B[2] * cl[2]));
  double dldr = w * fma(R[0], cl[0], fma(R[1], cl[1], R[2] * cl[2]));
  double db_drr = sh;
  double db_dcwf = 0.0, db_dswf = 0.0, db_dci = 0.0;
  if (sh > 0.0) {
    double rsh = rr * rcp_fast(sh);
    db_dcwf = rsh * cwf;
    db_dswf = rsh * ci2 * swf;
    db_dci = rsh * ci * swf * swf;
  }
  double rome2 = rcp_fast(ome2);
  double dfdM = opc * opc * rome2 * sqrt_fast(rome2);
  double dfde = (2.0 + e * cf) * sf * rome2;
  double drr_dcf = -rr * e * ropc;
  double drr_de = aor * (-2.0 * e * opc - ome2 * cf) * ropc * ropc;
  double db_dcf = db_drr * drr_dcf + db_dcwf * cw + db_dswf * sw;
  double db_dsf = db_dswf * cw - db_dcwf * sw;
  double db_df = db_dsf * cf - db_dcf * sf;
  double dl_dM = dldb * db_df * dfdM;
  double gt = -dl_dM * el[2];
  g[0] += gt;
  g[kStride] += gt;
  g[2 * kStride] = fma(dl_dM, tm, g[2 * kStride]);
  g[3 * kStride] = fma(dldb, db_df * dfde + db_drr * drr_de, g[3 * kStride]);
  g[4 * kStride] = fma(dldb, db_dcwf * cf + db_dswf * sf, g[4 * kStride]);
  g[5 * kStride] = fma(dldb, db_dswf * cf - db_dcwf * sf, g[5 * kStride]);
  g[6 * kStride] = fma(dldb, db_dci, g[6 * kStride]);
  g[8 * kStride] = fma(dldb, db_drr * rr * rcp_fast(aor), g[8 * kStride]);
  g[9 * kStride] += dldr;
  if (kDelay && delayed) {  // t' = t - D(theta): d lc / d theta += (d lc / d t') (-d D / d theta), d lc / d t' = -gt
    g[0] = fma(gt, dD[0], g[0]);
    g[kStride] = fma(gt, dD[1], g[kStride]);
    g[2 * kStride] = fma(gt, dD[2], g[2 * kStride]);
    g[3 * kStride] = fma(gt, dD[3], g[3 * kStride]);
    g[4 * kStride] = fma(gt, dD[4], g[4 * kStride]);
    g[5 * kStride] = fma(gt, dD[5], g[5 * kStride]);
    g[7 * kStride] = fma(gt, dD[6], g[7 * kStride]);
    g[8 * kStride] = fma(gt, dD[7], g[8 * kStride]);
  }
  return lc;
}

}  // namespace xo
