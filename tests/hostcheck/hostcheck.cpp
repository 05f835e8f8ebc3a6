// TEST INFRASTRUCTURE ONLY: compiles exoplanet_b200/csrc/xo_math.cuh (the exact
// __host__ __device__ arithmetic the kernels run) for the CPU so the logic can be
// checked against the oracle without a GPU.  Never loaded by the product package.
#include <cmath>
#include <cstdint>

#include "../../exoplanet_b200/csrc/xo_math.cuh"

// The contact-point solver and the model prologue are __device__ code; for this host build the
// qualifiers are defined away and the one warp shuffle of prologue_pair (the two lanes of a planet
// exchanging their contact times) is replaced by a mailbox, so that calling it for side 1 and then
// for side 0 reproduces what a lane pair does.
#define __device__
#define __forceinline__ inline
namespace hc_stub {
static double box_d = 0.0;
static int box_i = 0;
}  // namespace hc_stub
static inline double __shfl_xor_sync(unsigned, double v, int) {
  double r = hc_stub::box_d;
  hc_stub::box_d = v;
  return r;
}
static inline int __shfl_xor_sync(unsigned, int v, int) {
  int r = hc_stub::box_i;
  hc_stub::box_i = v;
  return r;
}
#include "../../exoplanet_b200/csrc/xo_contact.cuh"
#include "../../exoplanet_b200/csrc/xo_prologue.cuh"  // its warp-collective epilogue is compiled under nvcc only

extern "C" {
void hc_kepler(const double *M, const double *e, double *s, double *c, int64_t n) {
  for (int64_t i = 0; i < n; ++i) xo::kepler_solve(M[i], e[i], s[i], c[i]);
}
void hc_qsv(const double *b, const double *r, double *s, double *dsdb, double *dsdr, int64_t n) {
  for (int64_t i = 0; i < n; ++i) xo::qsv<true>(b[i], r[i], s + 3 * i, dsdb + 3 * i, dsdr + 3 * i);
}
void hc_qsv_nojac(const double *b, const double *r, double *s, int64_t n) {
  for (int64_t i = 0; i < n; ++i) xo::qsv<false>(b[i], r[i], s + 3 * i, nullptr, nullptr);
}
// lc and gradient at times t for ONE element block
void hc_lc_point(const double *el, const double *cl, const double *t, double *lc, double *g, double *gs, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    for (int j = 0; j < 10; ++j) g[10 * i + j] = 0.0;
    for (int j = 0; j < 3; ++j) gs[3 * i + j] = 0.0;
    lc[i] = xo::lc_point<true>(el, cl, t[i], 1.0, g + 10 * i, gs + 3 * i);
  }
}
// contact points, one call per side like the kernels; fast = the fused prologue's variant
void hc_contact(const double *a, const double *e, const double *cw, const double *sw, const double *ci, const double *L, int fast,
                double *Ml, double *Mr, int32_t *flag, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    int f0, f1;
    if (fast) {
      f0 = xo::cp_solve_side<true>(a[i], e[i], cw[i], sw[i], ci[i], L[i], -1.0, Ml[i]);
      f1 = xo::cp_solve_side<true>(a[i], e[i], cw[i], sw[i], ci[i], L[i], 1.0, Mr[i]);
    } else {
      f0 = xo::cp_solve_side<false>(a[i], e[i], cw[i], sw[i], ci[i], L[i], -1.0, Ml[i]);
      f1 = xo::cp_solve_side<false>(a[i], e[i], cw[i], sw[i], ci[i], L[i], 1.0, Mr[i]);
    }
    flag[i] = f0 ? f0 : f1;
  }
}
// the model prologue for n (chain, planet) pairs: par[n][8], star[n][4] -> el[n][16]
void hc_prologue(const double *par, const double *star, double texp, int use_in_transit, double *el, int64_t n) {
  double scratch[16];
  for (int64_t i = 0; i < n; ++i) {
    xo::prologue_pair(par + 8 * i, star + 4 * i, texp, use_in_transit, 1, true, scratch);  // odd lane: sends its contact
    xo::prologue_pair(par + 8 * i, star + 4 * i, texp, use_in_transit, 0, true, el + 16 * i);  // even lane: receives, writes
  }
}
void hc_in_window(const double *el, const double *t, int32_t *out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = xo::in_window(el, 1.0 / el[10], t[i]) ? 1 : 0;
}
}
