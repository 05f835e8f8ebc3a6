// xo_prologue.cuh -- the O(P) prologue of the transit model and its reverse pass, as warp-level
// device routines shared by the stand-alone kernels (xo_model.cu) and the single-launch model
// kernel (xo_transit.cu).
//   prologue: (period, t0, e, omega, b, r/R*, m_pl, K; R*, M*) -> element block, = KeplerianOrbit.__init__
//             (/root/reference/src/exoplanet/orbits/keplerian.py:133-281) + in_transit window (:729-767)
//   epilogue: fold of the per-warp partial sums + reverse pass of the prologue and of get_cl
//             (/root/reference/src/exoplanet/light_curves/limb_dark.py:11-18)
#pragma once
#include <cuda_runtime.h>

#include "xo_common.cuh"
#include "xo_contact.cuh"
#include "xo_math.cuh"
#include "xo_transit.cuh"

namespace xo {

constexpr double kGGrav = 2942.2062175044193;  // R_sun^3 / M_sun / day^2 (orbits/constants.py:32)

// Two adjacent lanes per planet (side = lane parity = contact side): the two bracketed Newton
// solves of the in-transit window run concurrently.  EVERY lane of the warp must call this
// (it shuffles); inactive lanes pass active = false.  The even lane of an active pair writes el[16].
__device__ inline void prologue_pair(const double *q, const double *st, double texp, int use_in_transit, int side,
                                     bool active, double *el) {
  const double period = q[0], t0 = q[1], e = q[2], omega = q[3], b = q[4], ror = q[5], mpl = q[6], K = q[7];
  const double R = st[0], Ms = st[1];
  // This kernel is one short dependent chain per lane (every instruction costs its full latency),
  // so it is written for instruction count: reciprocals by MUFU seed + Newton, the in-range
  // polynomial sincos / atan2 of xo_math.cuh, sin E0 and cos E0 from tan(E0 / 2) without a second
  // trigonometric call.
  const double rper = rcp_fast(period);
  const double a = cbrt(kGGrav * (Ms + mpl) * period * period * (1.0 / (4.0 * kPi * kPi)));
  const double n = 2.0 * kPi * rper;
  double sw, cw;
  if (fabs(omega) <= kPi) {
    double d0, d1;
    sincos_0pi(fabs(omega), sw, cw, d0, d1);
    sw = omega < 0.0 ? -sw : sw;
  } else {
    sincos(omega, &sw, &cw);
  }
  // E0 = 2 atan2(y, x), x >= 0:  sin E0 = 2xy / (x^2 + y^2),  cos E0 = (x^2 - y^2) / (x^2 + y^2)
  const double y = sqrt_fast(1.0 - e) * cw, x = sqrt_fast(1.0 + e) * (1.0 + sw);
  const double hE = atan2_pos(fabs(y), x);
  const double E0 = 2.0 * (y < 0.0 ? -hE : hE);
  const double x2y2 = fma(x, x, y * y);
  const double sE0 = x2y2 > 0.0 ? 2.0 * x * y * rcp_fast(x2y2) : 0.0;
  const double M0 = E0 - e * sE0;
  const double ome2 = (1.0 - e) * (1.0 + e);
  const double F = (1.0 + e * sw) * rcp_fast(ome2);
  const double ra = rcp_fast(a);
  const double cosi = F * R * ra * b;
  const double sini = sqrt_fast((1.0 - cosi) * (1.0 + cosi));
  double tc = side ? INFINITY : -INFINITY;  // this lane's contact time relative to t0
  int flag = 0;
  if (use_in_transit) {
    double Mc;
    flag = cp_solve_side<true>(a, e, cw, sw, cosi, R * (1.0 + ror), side ? 1.0 : -1.0, Mc);
    if (flag == 0) {
      const double hp = 0.5 * period;
      tc = (Mc - M0) * period * (1.0 / (2.0 * kPi)) + hp;
      tc = tc - period * floor(tc * rper) - hp;
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
  el[XO_EL_T0] = t0;
  el[XO_EL_TREF] = -M0 * period * (1.0 / (2.0 * kPi));
  el[XO_EL_N] = n;
  el[XO_EL_ECC] = e;
  el[XO_EL_COSW] = cw;
  el[XO_EL_SINW] = sw;
  el[XO_EL_COSI] = cosi;
  el[XO_EL_SINI] = sini;
  el[XO_EL_AOR] = a * rcp_fast(R);
  el[XO_EL_ROR] = ror;
  el[XO_EL_PERIOD] = period;
  el[XO_EL_TSTART] = ts;
  el[XO_EL_TEND] = te;
  el[XO_EL_KRV] = K;
  el[14] = 0.0;
  el[15] = 0.0;
}

__device__ inline void limb_coeffs(double u1, double u2, double *cl3) {
  const double c0 = 1.0 - u1 - 1.5 * u2, c1 = u1 + 2.0 * u2, c2 = -0.25 * u2;
  const double rn = 1.0 / (kPi * (c0 + c1 / 1.5));
  cl3[0] = c0 * rn;
  cl3[1] = c1 * rn;
  cl3[2] = c2 * rn;
}

// Parameter-only part of the reverse prologue of planet `lane` (everything that does not need the
// accumulated cotangents): computed before the partial sums are waited for.
struct EpiPre {
  double period, e, b, mt, a, n, sw, cw, sme, spe, x, y, sE0, cE0, M0, ome2, F, cosi, sini;
  // reciprocals the reverse pass multiplies by (computed here, before the partial sums are waited for)
  double rR, ra, rome2, rn, rper, rsme, rspe, rmt, rsini, rxy;
};
__device__ inline EpiPre epilogue_pre(const double *q, const double *st) {
  EpiPre r;
  const double R = st[0], Ms = st[1];
  r.period = q[0];
  r.e = q[2];
  r.b = q[4];
  r.mt = Ms + q[6];
  // same instruction-lean forms as prologue_pair (this half overlaps the logp kernel's tail, but the
  // sooner it is done the sooner the fold can start)
  r.a = cbrt(kGGrav * r.mt * r.period * r.period * (1.0 / (4.0 * kPi * kPi)));
  r.n = 2.0 * kPi * rcp_fast(r.period);
  if (fabs(q[3]) <= kPi) {
    double d0, d1;
    sincos_0pi(fabs(q[3]), r.sw, r.cw, d0, d1);
    r.sw = q[3] < 0.0 ? -r.sw : r.sw;
  } else {
    sincos(q[3], &r.sw, &r.cw);
  }
  r.sme = sqrt_fast(1.0 - r.e);
  r.spe = sqrt_fast(1.0 + r.e);
  r.y = r.sme * r.cw;
  r.x = r.spe * (1.0 + r.sw);
  const double hE = atan2_pos(fabs(r.y), r.x);
  const double E0 = 2.0 * (r.y < 0.0 ? -hE : hE);
  const double x2y2 = fma(r.x, r.x, r.y * r.y);
  const double rxy = x2y2 > 0.0 ? rcp_fast(x2y2) : 0.0;
  r.sE0 = 2.0 * r.x * r.y * rxy;
  r.cE0 = x2y2 > 0.0 ? (r.x - r.y) * (r.x + r.y) * rxy : 1.0;
  r.M0 = E0 - r.e * r.sE0;
  r.ome2 = (1.0 - r.e) * (1.0 + r.e);
  r.F = (1.0 + r.e * r.sw) * rcp_fast(r.ome2);
  r.cosi = r.F * R * rcp_fast(r.a) * r.b;
  r.sini = sqrt_fast((1.0 - r.cosi) * (1.0 + r.cosi));
  r.rR = rcp_fast(R);
  r.ra = rcp_fast(r.a);
  r.rome2 = rcp_fast(r.ome2);
  r.rper = rcp_fast(r.period);
  r.rn = r.period * (1.0 / (2.0 * kPi));
  r.rsme = rcp_fast(r.sme);
  r.rspe = rcp_fast(r.spe);
  r.rmt = rcp_fast(r.mt);
  r.rsini = r.sini > 0.0 ? rcp_fast(r.sini) : 0.0;
  r.rxy = rxy;
  return r;
}

#if defined(__CUDACC__)  // (tests/hostcheck compiles the prologue above for the CPU; what follows needs a warp)
// Warp-collective epilogue of ONE chain: lane l sums slots l, l+32, ... of every accumulator
// (ascending), a fixed-order butterfly combines the 32 lane sums (deterministic), lane p < P runs
// the reverse pass of planet p's prologue (pre = epilogue_pre of that planet), lane 0 the reverse
// of get_cl.  Accumulators are folded five at a time (nacc = 5 + 10 P) so that the loads and the
// shuffle trees of a group overlap.  kCoherent: read the partials through L2 (they were written by
// other CTAs of the same launch).
template <bool kCoherent>
__device__ inline void epilogue_chain(const EpiPre &pre, const double *st, const double *src, int slots, int P,
                                      double lognorm, double *logp, double *gpar_c, double *gstar_c, double *gmean) {
  constexpr unsigned kFullM = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int nacc = kNBase + P * XO_NGRAD;
  double g[XO_NGRAD];
#pragma unroll
  for (int j = 0; j < XO_NGRAD; ++j) g[j] = 0.0;
  double base = 0.0;
  for (int j0 = 0; j0 < nacc; j0 += 5) {
    double a5[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int s = lane; s < slots; s += 32) {
      const double *ptr = src + (int64_t)s * nacc + j0;
#pragma unroll
      for (int u = 0; u < 5; ++u) a5[u] += kCoherent ? __ldcg(ptr + u) : ptr[u];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < 5; ++u) a5[u] += __shfl_xor_sync(kFullM, a5[u], o);
    }
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const int j = j0 + u;
      if (j < kNBase) {
        if (lane == 8 + j) base = a5[u];
      } else {
        const int pj = (j - kNBase) / XO_NGRAD, k = (j - kNBase) % XO_NGRAD;
        if (lane == pj) {
#pragma unroll
          for (int kk = 0; kk < XO_NGRAD; ++kk)
            if (kk == k) g[kk] = a5[u];
        }
      }
    }
  }
  const double R = st[0];
  double gR = 0.0, gMs = 0.0;
  if (lane < P) {
    const double e = pre.e, b = pre.b, a = pre.a, n = pre.n, sw = pre.sw, cw = pre.cw;
    const double sme = pre.sme, spe = pre.spe, x = pre.x, y = pre.y, sE0 = pre.sE0, cE0 = pre.cE0, M0 = pre.M0;
    const double F = pre.F, cosi = pre.cosi;
    // cotangents of the element block: g = (t0, tref, n, e, cw, sw, cosi, sini, aor, ror)
    const double rR = pre.rR, ra = pre.ra, rome2 = pre.rome2, rn = pre.rn;
    double gt0 = g[0], ge = g[3], gcw = g[4], gsw = g[5];
    double ga = g[8] * rR;
    gR += -g[8] * a * rR * rR;
    double gcosi = g[6] - g[7] * cosi * pre.rsini;
    double gF = gcosi * R * b * ra;
    gR += gcosi * F * b * ra;
    ga += -gcosi * F * R * b * ra * ra;
    double gb = gcosi * F * R * ra;
    ge += gF * (sw * rome2 + (1.0 + e * sw) * 2.0 * e * rome2 * rome2);
    gsw += gF * e * rome2;
    double gM0 = -g[1] * rn;
    double gn = g[2] + g[1] * M0 * rn * rn;
    double gP = -gn * n * pre.rper;
    double gE0 = gM0 * (1.0 - e * cE0);
    ge += -gM0 * sE0;
    const double rr = pre.rxy;
    double gy = 2.0 * gE0 * x * rr, gx = -2.0 * gE0 * y * rr;
    gcw += gy * sme;
    ge += -gy * cw * 0.5 * pre.rsme;
    gsw += gx * spe;
    ge += gx * (1.0 + sw) * 0.5 * pre.rspe;
    double gomega = -gcw * sw + gsw * cw;
    gP += ga * (2.0 / 3.0) * a * pre.rper;
    double gmt = ga * a * (1.0 / 3.0) * pre.rmt;
    gMs += gmt;
    double *o = gpar_c + lane * XO_NPAR;
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
    *logp = lp + lognorm;
    *gmean = gm;
    // reverse of get_cl
    const double u1 = st[2], u2 = st[3];
    const double c0 = 1.0 - u1 - 1.5 * u2, c1 = u1 + 2.0 * u2, c2 = -0.25 * u2;
    const double norm = kPi * (c0 + c1 / 1.5), rn = 1.0 / norm;
    double gnorm = -(q0 * c0 + q1 * c1 + q2 * c2) * rn * rn;
    double gc0 = q0 * rn + gnorm * kPi, gc1 = q1 * rn + gnorm * kPi / 1.5, gc2 = q2 * rn;
    gstar_c[0] = gR;
    gstar_c[1] = gMs;
    gstar_c[2] = -gc0 + gc1;
    gstar_c[3] = -1.5 * gc0 + 2.0 * gc1 - 0.25 * gc2;
  }
}

// convenience: both halves back to back (single-launch modes of the logp kernel)
template <bool kCoherent>
__device__ inline void epilogue_chain(const double *par_c, const double *st, const double *src, int slots, int P,
                                      double lognorm, double *logp, double *gpar_c, double *gstar_c, double *gmean) {
  const int lane = threadIdx.x & 31;
  const EpiPre pre = epilogue_pre(par_c + (lane < P ? lane : 0) * XO_NPAR, st);
  epilogue_chain<kCoherent>(pre, st, src, slots, P, lognorm, logp, gpar_c, gstar_c, gmean);
}
#endif  // __CUDACC__

}  // namespace xo
