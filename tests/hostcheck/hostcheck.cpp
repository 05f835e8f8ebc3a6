// TEST INFRASTRUCTURE ONLY: compiles exoplanet_b200/csrc/xo_math.cuh (the exact
// __host__ __device__ arithmetic the kernels run) for the CPU so the logic can be
// checked against the oracle without a GPU.  Never loaded by the product package.
#include <cstdint>

#include "../../exoplanet_b200/csrc/xo_math.cuh"

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
void hc_in_window(const double *el, const double *t, int32_t *out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = xo::in_window(el, 1.0 / el[10], t[i]) ? 1 : 0;
}
}
