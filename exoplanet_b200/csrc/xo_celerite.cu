// xo_celerite.cu -- celerite2 SHOTerm GP log-likelihood as a parallel scan (BASELINE
// config 4).  celerite2 (>=0.3.1, /root/reference/setup.py:36) has NO call site in the
// reference; the contract here is celerite2's documented SHOTerm/GaussianProcess
// log_likelihood, pinned against a dense Cholesky oracle.
//
// The celerite recurrence is a Riccati (linear-fractional) recursion, so it is run in
// its equivalent 2-state Kalman-filter form, which has an associative combine
// (Sarkka & Garcia-Fernandez 2021).  One kernel, one read of the data (gp_forward_kernel): a warp stages a tile of
// 512 cadences of (t, diag, y) in shared memory, every lane walks a run of kRun consecutive cadences once from a
// *known* start (P = 0), tracking the affine dependence of the filtered mean on the unknown start state and the
// information (eta, J) the run's data carry about it -> one scan element per run, no 2x2 inverses; warp scan; a
// fixed 32-ary tree over the tile totals of the chain (flags + bounded waits, tiles handed out by ticket) gives
// the state entering the tile; every lane then re-walks its run from shared memory with plain Kalman steps,
// accumulating -1/2 (v^2/D + log D); the last tile of a chain folds the per-tile sums in order.  A tile is ONE
// warp on purpose: the tree look-back is a serial stretch of a few thousand dependent instructions, and with
// one-warp CTAs nobody waits at a barrier for it -- the other 15 resident warps of the SM keep computing.
// SHO in (a, c, h) form so celerite2's eps-regularised coefficients are matched exactly:
//   underdamped k(tau) = a e^{-c tau}(cos h tau + (c/h) sin h tau), overdamped with cosh/sinh.
#include <cuda_runtime.h>

#include "xo_common.cuh"
#include "xo_math.cuh"

namespace {

#ifndef XO_GP_RUN
#define XO_GP_RUN 16  // cadences per thread (build-time knob)
#endif
constexpr int kRun = XO_GP_RUN;
// the reverse (smoother / adjoint) passes carry more state per step: shorter runs, measured 27% faster at 16 chains
constexpr int kRunB = 8;
constexpr int kThreadsGp = 128;
constexpr int kWarpsGp = kThreadsGp / 32;
constexpr unsigned kFullGp = 0xffffffffu;
constexpr int kAggDoubles = 14;

struct Sho {
  double a, c, h, w2;  // variance, damping, (pseudo)frequency, c^2 +- h^2
  bool under;
};

__device__ __forceinline__ Sho make_sho(double S0, double w0, double Q, double eps) {
  Sho k;
  k.a = S0 * w0 * Q;
  k.c = 0.5 * w0 / Q;
  k.under = Q >= 0.5;
  double f = sqrt(fmax(k.under ? 4.0 * Q * Q - 1.0 : 1.0 - 4.0 * Q * Q, eps));
  k.h = k.c * f;
  k.w2 = k.under ? k.c * k.c + k.h * k.h : k.c * k.c - k.h * k.h;
  return k;
}

// Cf = e^{-c dt} cos(h dt) (cosh when overdamped), Sf = e^{-c dt} sin(h dt) / h (sinh)
__device__ __forceinline__ void sho_cs(const Sho &k, double dt, double &Cf, double &Sf) {
  if (k.under) {
    double sn, cs;
    sincos(k.h * dt, &sn, &cs);
    double ex = exp(-k.c * dt);
    Cf = ex * cs;
    Sf = ex * sn / k.h;
  } else {
    double ep = exp(-(k.c - k.h) * dt), en = exp(-(k.c + k.h) * dt);
    Cf = 0.5 * (ep + en);
    Sf = 0.5 * (ep - en) / k.h;
  }
}

// (Cf, Sf) of an earlier step of the same thread.  On a (nearly) uniform grid consecutive steps
// differ only by the rounding of the time stamps (|delta| ~ 1e-10 dt for Kepler/TESS-like series), and
//   d/dt (Cf, Sf) = (-c Cf -+ h^2 Sf,  Cf - c Sf)
// carries the cached pair to the new dt with a second-order error <= (1e-7 dt (c + h))^2 / 2 --
// 4 flops instead of sincos + exp.  Irregular sampling falls through to the full evaluation.
struct ShoCache {
  double dt = -1.0, Cf = 1.0, Sf = 0.0;
};
__device__ __forceinline__ void sho_cs(const Sho &k, double dt, ShoCache &cc, double &Cf, double &Sf) {
  const double d = dt - cc.dt;
  if (cc.dt > 0.0 && fabs(d) <= 1e-7 * cc.dt) {
    const double h2 = k.under ? k.h * k.h : -k.h * k.h;
    Cf = fma(d, -fma(k.c, cc.Cf, h2 * cc.Sf), cc.Cf);
    Sf = fma(d, fma(-k.c, cc.Sf, cc.Cf), cc.Sf);
    return;
  }
  sho_cs(k, dt, Cf, Sf);
  cc.dt = dt;
  cc.Cf = Cf;
  cc.Sf = Sf;
}

// transition matrix over dt
__device__ __forceinline__ void sho_phi(const Sho &k, double dt, ShoCache &cc, double &f00, double &f01, double &f10,
                                        double &f11) {
  double Cf, Sf;
  sho_cs(k, dt, cc, Cf, Sf);
  f00 = fma(k.c, Sf, Cf);
  f01 = Sf;
  f10 = -k.w2 * Sf;
  f11 = fma(-k.c, Sf, Cf);
}

struct Agg {
  double A00, A01, A10, A11, b0, b1, C00, C01, C11, e0, e1, J00, J01, J11;
};

__device__ __forceinline__ Agg agg_identity() { return Agg{1, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; }

// combine(earlier i, later j)
__device__ __forceinline__ Agg agg_combine(const Agg &i, const Agg &j) {
  // M = I + C_i J_j ; X = M^-1
  double m00 = 1.0 + i.C00 * j.J00 + i.C01 * j.J01, m01 = i.C00 * j.J01 + i.C01 * j.J11;
  double m10 = i.C01 * j.J00 + i.C11 * j.J01, m11 = 1.0 + i.C01 * j.J01 + i.C11 * j.J11;
  double rdet = 1.0 / (m00 * m11 - m01 * m10);
  double x00 = m11 * rdet, x01 = -m01 * rdet, x10 = -m10 * rdet, x11 = m00 * rdet;
  // G = A_j X
  double g00 = j.A00 * x00 + j.A01 * x10, g01 = j.A00 * x01 + j.A01 * x11;
  double g10 = j.A10 * x00 + j.A11 * x10, g11 = j.A10 * x01 + j.A11 * x11;
  Agg o;
  o.A00 = g00 * i.A00 + g01 * i.A10;
  o.A01 = g00 * i.A01 + g01 * i.A11;
  o.A10 = g10 * i.A00 + g11 * i.A10;
  o.A11 = g10 * i.A01 + g11 * i.A11;
  double u0 = i.b0 + i.C00 * j.e0 + i.C01 * j.e1, u1 = i.b1 + i.C01 * j.e0 + i.C11 * j.e1;
  o.b0 = g00 * u0 + g01 * u1 + j.b0;
  o.b1 = g10 * u0 + g11 * u1 + j.b1;
  // T = G C_i ; C = T A_j^T + C_j
  double t00 = g00 * i.C00 + g01 * i.C01, t01 = g00 * i.C01 + g01 * i.C11;
  double t10 = g10 * i.C00 + g11 * i.C01, t11 = g10 * i.C01 + g11 * i.C11;
  o.C00 = t00 * j.A00 + t01 * j.A01 + j.C00;
  o.C01 = 0.5 * ((t00 * j.A10 + t01 * j.A11) + (t10 * j.A00 + t11 * j.A01)) + j.C01;
  o.C11 = t10 * j.A10 + t11 * j.A11 + j.C11;
  // H = A_i^T X^T
  double h00 = i.A00 * x00 + i.A10 * x01, h01 = i.A00 * x10 + i.A10 * x11;
  double h10 = i.A01 * x00 + i.A11 * x01, h11 = i.A01 * x10 + i.A11 * x11;
  double v0 = j.e0 - (j.J00 * i.b0 + j.J01 * i.b1), v1 = j.e1 - (j.J01 * i.b0 + j.J11 * i.b1);
  o.e0 = h00 * v0 + h01 * v1 + i.e0;
  o.e1 = h10 * v0 + h11 * v1 + i.e1;
  // S = H J_j ; J = S A_i + J_i
  double s00 = h00 * j.J00 + h01 * j.J01, s01 = h00 * j.J01 + h01 * j.J11;
  double s10 = h10 * j.J00 + h11 * j.J01, s11 = h10 * j.J01 + h11 * j.J11;
  o.J00 = s00 * i.A00 + s01 * i.A10 + i.J00;
  o.J01 = 0.5 * ((s00 * i.A01 + s01 * i.A11) + (s10 * i.A00 + s11 * i.A10)) + i.J01;
  o.J11 = s10 * i.A01 + s11 * i.A11 + i.J11;
  return o;
}

// apply an aggregate to a Gaussian state (m, P)
__device__ __forceinline__ void agg_apply(const Agg &g, double &m0, double &m1, double &P00, double &P01, double &P11) {
  double q00 = 1.0 + P00 * g.J00 + P01 * g.J01, q01 = P00 * g.J01 + P01 * g.J11;
  double q10 = P01 * g.J00 + P11 * g.J01, q11 = 1.0 + P01 * g.J01 + P11 * g.J11;
  double rdet = 1.0 / (q00 * q11 - q01 * q10);
  double x00 = q11 * rdet, x01 = -q01 * rdet, x10 = -q10 * rdet, x11 = q00 * rdet;
  double g00 = g.A00 * x00 + g.A01 * x10, g01 = g.A00 * x01 + g.A01 * x11;
  double g10 = g.A10 * x00 + g.A11 * x10, g11 = g.A10 * x01 + g.A11 * x11;
  double u0 = m0 + P00 * g.e0 + P01 * g.e1, u1 = m1 + P01 * g.e0 + P11 * g.e1;
  double n0 = g00 * u0 + g01 * u1 + g.b0, n1 = g10 * u0 + g11 * u1 + g.b1;
  double t00 = g00 * P00 + g01 * P01, t01 = g00 * P01 + g01 * P11;
  double t10 = g10 * P00 + g11 * P01, t11 = g10 * P01 + g11 * P11;
  double c00 = t00 * g.A00 + t01 * g.A01 + g.C00;
  double c01 = 0.5 * ((t00 * g.A10 + t01 * g.A11) + (t10 * g.A00 + t11 * g.A01)) + g.C01;
  double c11 = t10 * g.A10 + t11 * g.A11 + g.C11;
  m0 = n0;
  m1 = n1;
  P00 = c00;
  P01 = c01;
  P11 = c11;
}

// ---------------------------------------------------------------------------
// Inter-tile plumbing of the one-pass scans: tickets, flags, bounded spin-waits, a fixed 32-ary tree.
//
// A tile is the work of ONE WARP (a CTA of 32 threads): while a warp waits for its predecessors no other warp
// idles at a barrier, and the other resident warps of the SM (16 forward, 12 reverse) keep the pipes busy.
// Tiles are handed out by an atomic ticket (chain-major: ticket = chain * tiles + tile), so a warp only ever waits
// on tiles with a LOWER ticket, i.e. on warps that are already running: forward progress does not depend on the
// hardware's CTA scheduling order.  A flag is published with a release store after the payload; readers poll it
// with acquire loads and read the payload through L2 (ld.cg).  Every wait is bounded (kSpinLimit polls with
// back-off): on expiry the warp records a failure in the status word and carries on, so the kernel always
// terminates (include/xo_b200.h: "every iteration is bounded").
//
// The exclusive prefix of tile k over the earlier tiles of its chain comes from a FIXED tree, so the result is
// bit-reproducible: level 0 = tile totals, level 1 = totals of groups of 32 tiles, level 2 = totals of groups of
// 1024 tiles.  prefix(k) = [level-2 totals before k's level-2 group] o [level-1 totals before k's group, inside that
// level-2 group] o [tile totals before k inside its group] -- at most 31 + 31 entries below level 2, reduced in one
// pass by a warp holding two entries per lane.  The last tile of a group publishes the group total BEFORE it waits
// for anything outside the group (else the group totals would be published one after the other, a serial chain).
// ---------------------------------------------------------------------------
constexpr unsigned kSpinLimit = 1u << 22;  // x >= 40 ns back-off: > 0.15 s, far beyond any real wait

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// true when *flag became non-zero within the bound
__device__ __forceinline__ bool wait_flag(const unsigned *flag) {
  unsigned ns = 20;
  for (unsigned it = 0; it < kSpinLimit; ++it) {
    if (ld_acquire_u32(flag) != 0u) return true;
    __nanosleep(ns);
    if (ns < 160) ns += 20;
  }
  return false;
}

// status word at the head of every celerite workspace (see xo_read_status): [0] lanes whose result was
// not finite, [1] expired waits (a broken scan: the result must be discarded)
struct GpStatus {
  unsigned long long bad_lanes, wait_failures;
};

struct AggOps {
  using E = Agg;
  static constexpr int kDoubles = kAggDoubles;
  static __device__ __forceinline__ E identity() { return agg_identity(); }
  static __device__ __forceinline__ E combine(const E &earlier, const E &later) { return agg_combine(earlier, later); }
  static __device__ __forceinline__ double *data(E &e) { return &e.A00; }
  static __device__ __forceinline__ const double *data(const E &e) { return &e.A00; }
};

template <class O>
__device__ __forceinline__ typename O::E elem_shfl_up(const typename O::E &a, int delta) {
  typename O::E o;
#pragma unroll
  for (int k = 0; k < O::kDoubles; ++k) O::data(o)[k] = __shfl_up_sync(kFullGp, O::data(a)[k], delta);
  return o;
}
template <class O>
__device__ __forceinline__ typename O::E elem_shfl(const typename O::E &a, int src) {
  typename O::E o;
#pragma unroll
  for (int k = 0; k < O::kDoubles; ++k) O::data(o)[k] = __shfl_sync(kFullGp, O::data(a)[k], src);
  return o;
}
template <class O>
__device__ __forceinline__ typename O::E elem_load_cg(const double *p) {
  typename O::E a;
#pragma unroll
  for (int k = 0; k < O::kDoubles; ++k) O::data(a)[k] = __ldcg(p + k);
  return a;
}
template <class O>
__device__ __forceinline__ void elem_store(double *p, const typename O::E &a) {
#pragma unroll
  for (int k = 0; k < O::kDoubles; ++k) p[k] = O::data(a)[k];
}

// inclusive scan over the 32 lanes of a warp (lane order = scan order)
template <class O>
__device__ __forceinline__ typename O::E warp_inclusive_scan(typename O::E v) {
  const int lane = threadIdx.x & 31;
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {
    typename O::E o = elem_shfl_up<O>(v, d);
    if (lane >= d) v = O::combine(o, v);
  }
  return v;
}
// fixed-order combination of the elements held by lanes [0, count): e_0 o e_1 o ... o e_{count-1} in every lane;
// only ceil(log2 count) levels are executed
template <class O>
__device__ __forceinline__ typename O::E warp_reduce_n(typename O::E v, int count) {
  const int lane = threadIdx.x & 31;
#pragma unroll 1
  for (int d = 1; d < count; d <<= 1) {
    typename O::E o = elem_shfl_up<O>(v, d);
    if (lane >= d && lane < count) v = O::combine(o, v);
  }
  return elem_shfl<O>(v, count > 0 ? count - 1 : 0);
}

// one level of the tree for one chain
struct TreeSeg {
  const double *elems;
  const unsigned *flags;
  int start, count;  // entries [start, start + count) of the level, in order
};
// e = entry `pos` of the concatenation s0 | s1 | s2 (waits for its flag)
template <class O>
__device__ __forceinline__ typename O::E tree_fetch(const TreeSeg &s0, const TreeSeg &s1, const TreeSeg &s2, int pos, bool &ok) {
  const TreeSeg &s = pos < s0.count ? s0 : (pos < s0.count + s1.count ? s1 : s2);
  const int idx = s.start + (pos < s0.count ? pos : (pos < s0.count + s1.count ? pos - s0.count : pos - s0.count - s1.count));
  ok &= wait_flag(s.flags + idx);
  return elem_load_cg<O>(s.elems + (size_t)idx * O::kDoubles);
}
// the concatenation s0 | s1 | s2 combined in order, 64 entries per pass (two per lane)
template <class O>
__device__ __forceinline__ typename O::E tree_reduce(const TreeSeg &s0, const TreeSeg &s1, const TreeSeg &s2, bool &ok) {
  const int lane = threadIdx.x & 31;
  const int total = s0.count + s1.count + s2.count;
  typename O::E carry = O::identity();
#pragma unroll 1
  for (int base = 0; base < total; base += 64) {
    const int cnt = min(64, total - base);
    typename O::E mine = O::identity();
    if (cnt <= 32) {
      if (lane < cnt) mine = tree_fetch<O>(s0, s1, s2, base + lane, ok);
      mine = warp_reduce_n<O>(mine, cnt);
    } else {
      if (2 * lane < cnt) mine = tree_fetch<O>(s0, s1, s2, base + 2 * lane, ok);
      if (2 * lane + 1 < cnt) mine = O::combine(mine, tree_fetch<O>(s0, s1, s2, base + 2 * lane + 1, ok));
      mine = warp_reduce_n<O>(mine, (cnt + 1) / 2);
    }
    carry = base == 0 ? mine : O::combine(carry, mine);
  }
  return carry;
}

// scratch of one chain's tree (pointers already offset to the chain)
struct TreeView {
  double *e0, *e1, *e2;
  unsigned *f0, *f1, *f2;
  int n0;  // tiles of the chain
};
template <class O>
__device__ __forceinline__ void tree_publish(double *elems, unsigned *flags, int idx, const typename O::E &v) {
  if ((threadIdx.x & 31) == 0) {
    elem_store<O>(elems + (size_t)idx * O::kDoubles, v);
    __threadfence();
    st_release_u32(flags + idx, 1u);
  }
}
// exclusive prefix of tile `tile` over the earlier tiles of the chain; `own` = this tile's total (valid in every lane).
// Publishes the tile total and, from the last tile of a group, the group totals.
// (not inlined: called once per tile; its many live scan elements must not compete with the walks for registers)
template <class O>
__device__ __noinline__ typename O::E tree_exclusive(const TreeView &tv, int tile, const typename O::E &own, bool &ok) {
  tree_publish<O>(tv.e0, tv.f0, tile, own);
  const int d0 = tile & 31, g1 = tile >> 5, d1 = g1 & 31, g2 = g1 >> 5;
  const TreeSeg none{nullptr, nullptr, 0, 0};
  const TreeSeg tiles{tv.e0, tv.f0, g1 * 32, d0}, groups{tv.e1, tv.f1, g2 * 32, d1}, supers{tv.e2, tv.f2, 0, g2};
  if (d0 != 31) return tree_reduce<O>(supers, groups, tiles, ok);
  // last tile of a full group of 32
  const typename O::E a = tree_reduce<O>(tiles, none, none, ok);
  const typename O::E t1 = O::combine(a, own);
  if ((g1 + 1) * 32 < tv.n0) tree_publish<O>(tv.e1, tv.f1, g1, t1);
  if (d1 != 31) return O::combine(tree_reduce<O>(supers, groups, none, ok), a);
  const typename O::E b = tree_reduce<O>(groups, none, none, ok);
  if ((g2 + 1) * 1024 < tv.n0) tree_publish<O>(tv.e2, tv.f2, g2, O::combine(b, t1));
  return O::combine(O::combine(tree_reduce<O>(supers, none, none, ok), b), a);
}

// ---- one-pass forward scan ---------------------------------------------------
// One warp owns a tile of kTile = 32 * kRun cadences of one chain:
//   1. the tile of (t, diag, y) is read from HBM ONCE, coalesced, into shared memory (row pitch kRun + 1
//      doubles: lane r then walks row r conflict-free);
//   2. every lane walks its run from a known start (P = 0) -> one scan element per run; warp scan;
//   3. the tile total is published and the exclusive prefix over the earlier tiles of the chain comes from the
//      fixed tree above; applied to the stationary prior it gives the state entering the tile;
//   4. every lane applies its in-warp exclusive prefix and re-walks its run from shared memory with plain
//      Kalman steps, accumulating -1/2 (v^2/D + log D);
//   5. per-tile sums are folded in tile order by the last tile of the chain to finish (bit-reproducible).
constexpr int kTile = 32 * kRun;
constexpr int kPitch = kRun + 1;

// sum of log D over a run as the log of a product of mantissas plus an integer exponent sum: one DMUL and a
// few integer instructions per cadence instead of a log; mantissas are in [0.5, 1), so a run of <= 64 cadences
// cannot underflow, and the exponent sum is exact for ANY finite positive D (no D^4 overflow).
struct LogProd {
  double mant = 1.0, slow = 0.0;
  int expo = 0;
  __device__ __forceinline__ void mul(double D) {
    const int hi = __double2hiint(D);
    const int ef = (hi >> 20) & 0x7ff;
    if (hi > 0 && ef != 0 && ef != 0x7ff) {  // positive, normal
      expo += ef - 1022;
      mant *= __hiloint2double((hi & 0x800fffff) | 0x3fe00000, __double2loint(D));
    } else {
      slow += log(D);  // denormal, zero, negative, inf, nan: exact library semantics (NaN / -inf propagate)
    }
  }
  __device__ __forceinline__ double value() const { return log(mant) + slow + (double)expo * 0.69314718055994530941723212145818; }
};

// device scratch shared by the forward and the reverse tree: per level elements + flags, [B][n_level]
struct TreeBuffers {
  double *e0, *e1, *e2;
  unsigned *f0, *f1, *f2;
  int n0, n1, n2;
  __device__ __forceinline__ TreeView chain(int64_t b, int doubles) const {
    return TreeView{e0 + (size_t)b * n0 * doubles, e1 + (size_t)b * n1 * doubles, e2 + (size_t)b * n2 * doubles, f0 + (size_t)b * n0,
                    f1 + (size_t)b * n1, f2 + (size_t)b * n2, n0};
  }
};

#ifndef XO_GP_MINBLOCKS
#define XO_GP_MINBLOCKS 16  // resident warp-tiles per SM the forward scan is compiled for (build-time knob): 128 registers
#endif
template <int REC /* 0: none, 1: (v/D, K0, K1) per cadence for the smoother, 2: state entering every run of kRunB cadences */>
__global__ void __launch_bounds__(32, XO_GP_MINBLOCKS)
    gp_forward_kernel(const double *__restrict__ t, const double *__restrict__ y, const double *__restrict__ diag, int64_t N,
                      const double *__restrict__ S0, const double *__restrict__ w0, const double *__restrict__ Qf,
                      const double *__restrict__ jitter /* [B] added to diag, or null */, double eps,
                      unsigned *__restrict__ counters /* [0] ticket, [1 + b] tiles done of chain b */, const TreeBuffers tb,
                      double *__restrict__ cta_ll, GpStatus *__restrict__ status, double *__restrict__ loglike,
                      double *__restrict__ rec, int64_t Tb) {
  __shared__ double st[32 * kPitch], sd[32 * kPitch], sy[32 * kPitch];
  const int lane = threadIdx.x;
  unsigned ticket = 0u;
  if (lane == 0) ticket = atomicAdd(&counters[0], 1u);
  ticket = __shfl_sync(kFullGp, ticket, 0);
  const int tiles = tb.n0;
  const int64_t bch = ticket / (unsigned)tiles;
  const int tile = (int)(ticket % (unsigned)tiles);
  const Sho k = make_sho(S0[bch], w0[bch], Qf[bch], eps);
  const double jit = jitter ? jitter[bch] : 0.0;
  const int64_t base = (int64_t)tile * kTile;
  const int nvalid = (int)min((int64_t)kTile, N - base);
  const double *yb = y + bch * N;
  // 1. the tile, coalesced (consecutive lanes read consecutive cadences), once
#pragma unroll 4
  for (int idx = lane; idx < kTile; idx += 32) {
    const int at = (idx / kRun) * kPitch + idx % kRun;
    if (idx < nvalid) {
      st[at] = __ldg(t + base + idx);
      sd[at] = __ldg(diag + base + idx) + jit;
      sy[at] = __ldg(yb + base + idx);
    }
  }
  const double t_halo = __ldg(t + (base > 0 ? base - 1 : 0));  // the time stamp before the tile's first cadence
  __syncwarp();
  const int64_t s = base + (int64_t)lane * kRun;
  const int len = (int)max((int64_t)0, min((int64_t)kRun, N - s));
  const double *ts = st + lane * kPitch, *ds = sd + lane * kPitch, *ys = sy + lane * kPitch;
  const double t_before = lane > 0 ? st[(lane - 1) * kPitch + kRun - 1] : t_halo;
  // 2. the run's scan element, from a known start
  Agg v = agg_identity();
  if (len > 0) {
    double A00 = 1, A01 = 0, A10 = 0, A11 = 1, b0 = 0, b1 = 0, e0 = 0, e1 = 0, J00 = 0, J01 = 0, J11 = 0;
    double Z00 = k.a, Z01 = 0.0, Z11 = k.a * k.w2;  // P = 0  <=>  Z = Sigma_inf
    double tprev = t_before;
    ShoCache cc;
#pragma unroll 1
    for (int i = 0; i < len; ++i) {
      const double ti = ts[i];
      double f00, f01, f10, f11;
      sho_phi(k, ti - tprev, cc, f00, f01, f10, f11);
      tprev = ti;
      // A <- Phi A, b <- Phi b, Z <- Phi Z Phi^T
      double a00 = f00 * A00 + f01 * A10, a01 = f00 * A01 + f01 * A11;
      double a10 = f10 * A00 + f11 * A10, a11 = f10 * A01 + f11 * A11;
      double c0 = f00 * b0 + f01 * b1, c1 = f10 * b0 + f11 * b1;
      double z00 = f00 * Z00 + f01 * Z01, z01 = f00 * Z01 + f01 * Z11;
      double z10 = f10 * Z00 + f11 * Z01, z11 = f10 * Z01 + f11 * Z11;
      double p00 = z00 * f00 + z01 * f01, p01 = z00 * f10 + z01 * f11, p11 = z10 * f10 + z11 * f11;
      const double pm = k.a - p00;
      const double D = pm + ds[i];
      const double rD = xo::rcp_fast(D);
      const double K0 = pm * rD, K1 = -p01 * rD;
      const double vb = ys[i] - c0;
      J00 = fma(a00 * a00, rD, J00);
      J01 = fma(a00 * a01, rD, J01);
      J11 = fma(a01 * a01, rD, J11);
      e0 = fma(a00 * vb, rD, e0);
      e1 = fma(a01 * vb, rD, e1);
      A00 = a00 - K0 * a00;
      A01 = a01 - K0 * a01;
      A10 = a10 - K1 * a00;
      A11 = a11 - K1 * a01;
      b0 = fma(K0, vb, c0);
      b1 = fma(K1, vb, c1);
      Z00 = fma(D * K0, K0, p00);
      Z01 = fma(D * K0, K1, p01);
      Z11 = fma(D * K1, K1, p11);
    }
    v = Agg{A00, A01, A10, A11, b0, b1, k.a - Z00, -Z01, k.a * k.w2 - Z11, e0, e1, J00, J01, J11};
  }
  v = warp_inclusive_scan<AggOps>(v);
  // exclusive prefix of this lane's run inside the tile: the inclusive value of the previous lane
  const Agg pre = elem_shfl_up<AggOps>(v, 1);
  // 3. the state entering the tile
  bool ok = true;
  const Agg excl = tree_exclusive<AggOps>(tb.chain(bch, kAggDoubles), tile, elem_shfl<AggOps>(v, 31), ok);
  if (!__all_sync(kFullGp, ok) && lane == 0) atomicAdd(&status->wait_failures, 1ull);
  // 4. re-walk with plain Kalman steps
  double acc = 0.0;
  if (len > 0) {
    double m0 = 0.0, m1 = 0.0, P00 = k.a, P01 = 0.0, P11 = k.a * k.w2;  // the stationary prior
    agg_apply(excl, m0, m1, P00, P01, P11);               // ... entering the tile
    if (lane > 0) agg_apply(pre, m0, m1, P00, P01, P11);  // ... entering this lane's run
    double Z00 = k.a - P00, Z01 = -P01, Z11 = k.a * k.w2 - P11;  // Z = Sigma_inf - P
    double tprev = t_before;
    ShoCache cc;
    LogProd lp;
#pragma unroll 1
    for (int i = 0; i < len; ++i) {
      const double ti = ts[i];
      double f00, f01, f10, f11;
      sho_phi(k, ti - tprev, cc, f00, f01, f10, f11);
      tprev = ti;
      if (REC == 2 && (i % kRunB) == 0) {  // state entering every run of kRunB cadences: where the reverse pass restarts
        double *rb = rec + bch * 5 * Tb + (s + i) / kRunB;
        rb[0] = m0;
        rb[Tb] = m1;
        rb[2 * Tb] = Z00;
        rb[3 * Tb] = Z01;
        rb[4 * Tb] = Z11;
      }
      double c0 = f00 * m0 + f01 * m1, c1 = f10 * m0 + f11 * m1;
      double z00 = f00 * Z00 + f01 * Z01, z01 = f00 * Z01 + f01 * Z11;
      double z10 = f10 * Z00 + f11 * Z01, z11 = f10 * Z01 + f11 * Z11;
      double p00 = z00 * f00 + z01 * f01, p01 = z00 * f10 + z01 * f11, p11 = z10 * f10 + z11 * f11;
      const double pm = k.a - p00;
      const double D = pm + ds[i];
      const double rD = xo::rcp_fast(D);
      const double K0 = pm * rD, K1 = -p01 * rD;
      const double vv = ys[i] - c0;
      acc = fma(vv * vv, rD, acc);
      lp.mul(D);
      if (REC == 1) {  // what the backward (disturbance-smoother) pass needs
        double *rb = rec + bch * 3 * N + s + i;
        rb[0] = vv * rD;
        rb[N] = K0;
        rb[2 * N] = K1;
      }
      m0 = fma(K0, vv, c0);
      m1 = fma(K1, vv, c1);
      Z00 = fma(D * K0, K0, p00);
      Z01 = fma(D * K0, K1, p01);
      Z11 = fma(D * K1, K1, p11);
    }
    acc += lp.value();
    if (!isfinite(acc)) atomicAdd(&status->bad_lanes, 1ull);
  }
  // 5. tile sum in a fixed order; the last tile of the chain folds the per-tile sums in tile order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFullGp, acc, o);
  int is_last = 0;
  if (lane == 0) {
    cta_ll[(size_t)bch * tiles + tile] = acc;
    __threadfence();
    is_last = atomicAdd(&counters[1 + bch], 1u) == (unsigned)(tiles - 1);
  }
  if (__shfl_sync(kFullGp, is_last, 0)) {
    __threadfence();
    double tot = 0.0;
    for (int i = lane; i < tiles; i += 32) tot += __ldcg(cta_ll + (size_t)bch * tiles + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(kFullGp, tot, o);
    if (lane == 0) loglike[bch] = -0.5 * tot - 0.5 * (double)N * 1.8378770664093454835606594728112;  // log(2 pi)
  }
}

// ---------------------------------------------------------------------------
// alpha = K^-1 y by the disturbance smoother (Durbin & Koopman 2012, sec. 4.5), run as an
// affine scan in reverse time: with w_k = v_k / D_k, K_k the Kalman gain and Phi_{k+1} the next
// transition,   R_{k-1} = (I - K_k h)^T Phi_{k+1}^T R_k + h^T w_k,  R_{N-1} = 0,
//               alpha_k = w_k - (Phi_{k+1} K_k) . R_k.
// Same three phases as the forward pass, thread g owning run T-1-g so the scan runs right-to-left;
// the scan element is an affine map (A 2x2, b 2), combine = composition.
// ---------------------------------------------------------------------------
struct Aff {
  double a00, a01, a10, a11, b0, b1;
};
__device__ __forceinline__ Aff aff_identity() { return Aff{1, 0, 0, 1, 0, 0}; }
// cur after prev:  x -> A_cur (A_prev x + b_prev) + b_cur
__device__ __forceinline__ Aff aff_after(const Aff &prev, const Aff &cur) {
  Aff o;
  o.a00 = cur.a00 * prev.a00 + cur.a01 * prev.a10;
  o.a01 = cur.a00 * prev.a01 + cur.a01 * prev.a11;
  o.a10 = cur.a10 * prev.a00 + cur.a11 * prev.a10;
  o.a11 = cur.a10 * prev.a01 + cur.a11 * prev.a11;
  o.b0 = cur.a00 * prev.b0 + cur.a01 * prev.b1 + cur.b0;
  o.b1 = cur.a10 * prev.b0 + cur.a11 * prev.b1 + cur.b1;
  return o;
}
__device__ __forceinline__ Aff aff_shfl_up(const Aff &a, int d) {
  Aff o;
  const double *src = &a.a00;
  double *dst = &o.a00;
#pragma unroll
  for (int k = 0; k < 6; ++k) dst[k] = __shfl_up_sync(kFullGp, src[k], d);
  return o;
}
__device__ Aff aff_cta_inclusive_scan(Aff v, Aff *warp_tot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    Aff o = aff_shfl_up(v, d);
    if (lane >= d) v = aff_after(o, v);
  }
  if (lane == 31) warp_tot[warp] = v;
  __syncthreads();
  if (warp > 0) {
    Aff pre = warp_tot[0];
    for (int w = 1; w < warp; ++w) pre = aff_after(pre, warp_tot[w]);
    v = aff_after(pre, v);
  }
  return v;
}

// one backward step k of run walking: given R (= R_k) returns alpha_k and updates R to R_{k-1}
__device__ __forceinline__ double gp_bwd_step(const Sho &k, const double *__restrict__ t, const double *__restrict__ rb,
                                              int64_t N, int64_t i, ShoCache &cc, double &R0, double &R1) {
  const double w = __ldg(rb + i), K0 = __ldg(rb + N + i), K1 = __ldg(rb + 2 * N + i);
  double f00 = 1.0, f01 = 0.0, f10 = 0.0, f11 = 1.0;
  if (i + 1 < N) sho_phi(k, __ldg(t + i + 1) - __ldg(t + i), cc, f00, f01, f10, f11);
  // q = Phi_{k+1}^T R
  const double q0 = f00 * R0 + f10 * R1, q1 = f01 * R0 + f11 * R1;
  const double kq = K0 * q0 + K1 * q1;  // (Phi K) . R = K . (Phi^T R)
  const double alpha = w - kq;
  R0 = q0 - kq + w;  // (I - K h)^T q + h^T w,  h = (1, 0)
  R1 = q1;
  return alpha;
}

__global__ void __launch_bounds__(kThreadsGp) gp_bwd_aggregate_kernel(const double *__restrict__ t, int64_t N,
                                                                      const double *__restrict__ S0,
                                                                      const double *__restrict__ w0,
                                                                      const double *__restrict__ Qf, double eps,
                                                                      int64_t T, const double *__restrict__ rec,
                                                                      double *__restrict__ bpref,
                                                                      double *__restrict__ bcta) {
  __shared__ Aff warp_tot[kWarpsGp];
  const int64_t bch = blockIdx.y;
  const Sho k = make_sho(S0[bch], w0[bch], Qf[bch], eps);
  const int64_t g = (int64_t)blockIdx.x * kThreadsGp + threadIdx.x;
  const int64_t run = T - 1 - g;  // reversed ownership
  Aff v = aff_identity();
  if (run >= 0) {
    const int64_t s = run * kRunB, e = min(s + kRunB, N);
    const double *rb = rec + bch * 3 * N;
    // the run's map R_{e-1} -> R_{s-1}: push the two basis vectors and the origin through it
    double o0 = 0.0, o1 = 0.0, x0 = 1.0, x1 = 0.0, y0 = 0.0, y1 = 1.0;
    ShoCache cc;
    for (int64_t i = e - 1; i >= s; --i) {
      const double K0 = __ldg(rb + N + i), K1 = __ldg(rb + 2 * N + i), w = __ldg(rb + i);
      double f00 = 1.0, f01 = 0.0, f10 = 0.0, f11 = 1.0;
      if (i + 1 < N) sho_phi(k, __ldg(t + i + 1) - __ldg(t + i), cc, f00, f01, f10, f11);
      auto lin = [&](double &r0, double &r1, double add) {
        const double q0 = f00 * r0 + f10 * r1, q1 = f01 * r0 + f11 * r1;
        r0 = q0 - (K0 * q0 + K1 * q1) + add;
        r1 = q1;
      };
      lin(o0, o1, w);
      lin(x0, x1, w);
      lin(y0, y1, w);
    }
    v = Aff{x0 - o0, y0 - o0, x1 - o1, y1 - o1, o0, o1};
  }
  v = aff_cta_inclusive_scan(v, warp_tot);
  if (g < T) {
    double *o = bpref + (bch * T + g) * 6;
    o[0] = v.a00; o[1] = v.a01; o[2] = v.a10; o[3] = v.a11; o[4] = v.b0; o[5] = v.b1;
  }
  if (threadIdx.x == kThreadsGp - 1) {
    double *o = bcta + (bch * gridDim.x + blockIdx.x) * 6;
    o[0] = v.a00; o[1] = v.a01; o[2] = v.a10; o[3] = v.a11; o[4] = v.b0; o[5] = v.b1;
  }
}

// R entering each CTA (from the right): exclusive scan of the CTA totals applied to R = 0
__global__ void __launch_bounds__(kThreadsGp) gp_bwd_cta_scan_kernel(const double *__restrict__ bcta, int nctas,
                                                                     double *__restrict__ bstate) {
  __shared__ Aff warp_tot[kWarpsGp];
  __shared__ Aff lastlane[kWarpsGp];
  __shared__ Aff carry_s;
  const int64_t bch = blockIdx.x;
  Aff carry = aff_identity();
  for (int base = 0; base < nctas; base += kThreadsGp) {
    const int idx = base + threadIdx.x;
    Aff v = aff_identity();
    if (idx < nctas) {
      const double *p = bcta + (bch * nctas + idx) * 6;
      v = Aff{p[0], p[1], p[2], p[3], p[4], p[5]};
    }
    v = aff_cta_inclusive_scan(v, warp_tot);
    Aff incl = aff_after(carry, v);
    Aff prev = aff_shfl_up(incl, 1);
    if ((threadIdx.x & 31) == 31) lastlane[threadIdx.x >> 5] = incl;
    __syncthreads();
    Aff excl = ((threadIdx.x & 31) == 0) ? (threadIdx.x == 0 ? carry : lastlane[(threadIdx.x >> 5) - 1]) : prev;
    if (idx < nctas) {
      double *o = bstate + (bch * nctas + idx) * 2;
      o[0] = excl.b0;  // the map applied to R = 0
      o[1] = excl.b1;
    }
    if (threadIdx.x == kThreadsGp - 1) carry_s = incl;
    __syncthreads();
    carry = carry_s;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreadsGp) gp_bwd_solve_kernel(const double *__restrict__ t, int64_t N,
                                                                  const double *__restrict__ S0,
                                                                  const double *__restrict__ w0,
                                                                  const double *__restrict__ Qf, double eps, int64_t T,
                                                                  const double *__restrict__ rec,
                                                                  const double *__restrict__ bpref,
                                                                  const double *__restrict__ bstate,
                                                                  double *__restrict__ alpha) {
  const int64_t bch = blockIdx.y;
  const Sho k = make_sho(S0[bch], w0[bch], Qf[bch], eps);
  const int64_t g = (int64_t)blockIdx.x * kThreadsGp + threadIdx.x;
  const int64_t run = T - 1 - g;
  if (run < 0) return;
  const int64_t s = run * kRunB, e = min(s + kRunB, N);
  const double *st = bstate + (bch * gridDim.x + blockIdx.x) * 2;
  double R0 = st[0], R1 = st[1];
  if (threadIdx.x > 0) {  // apply the in-CTA exclusive prefix (inclusive of the previous thread)
    const double *p = bpref + (bch * T + g - 1) * 6;
    const double n0 = p[0] * R0 + p[1] * R1 + p[4], n1 = p[2] * R0 + p[3] * R1 + p[5];
    R0 = n0;
    R1 = n1;
  }
  const double *rb = rec + bch * 3 * N;
  double *ab = alpha + bch * N;
  ShoCache cc;
  for (int64_t i = e - 1; i >= s; --i) ab[i] = gp_bwd_step(k, t, rb, N, i, cc, R0, R1);
}

// ---------------------------------------------------------------------------
// Reverse-mode gradient of the log-likelihood w.r.t. (S0, w0, Q), diag[N] and y[N] -- what
// celerite2's own reverse pass provides, here as a second parallel scan.
//
// Forward step i maps the state s = (m, Z) to
//   cvec = Phi m,  Pz = Phi Z Phi^T,  k = (a - Pz00, -Pz01),  D = k0 + diag_i,  v = y_i - cvec0,
//   l_i = -1/2 (v^2/D + log D),       m' = cvec + k v/D,       Z' = Pz + k k^T / D.
// Its adjoint  lambda_{i-1} = J_i^T lambda_i + d l_i / d s_{i-1}  is AFFINE in the 5-vector
// lambda = (mbar0, mbar1, Zbar00, Zbar01, Zbar11), with block-triangular J^T (mbar does not depend
// on Zbar), so it runs as a reverse affine scan exactly like the disturbance smoother above:
// thread g owns run T-1-g, composes its run's map (24 doubles), in-CTA shuffle scan, one CTA over
// the CTA totals, then every thread re-walks its run backwards with the now known incoming lambda
// and accumulates the parameter cotangents.  The forward quantities of a step are recomputed from
// the state entering every run of kRunB cadences, which the forward scan leaves behind.
// ---------------------------------------------------------------------------
constexpr int kAdjDoubles = 24;
struct Adj {
  double A[4];  // mbar_in = A mbar_out + om
  double B[6];  // Zbar_in = B mbar_out + C Zbar_out + oz   (B 3x2, C 3x3, row-major)
  double C[9];
  double om[2], oz[3];
};
__device__ __forceinline__ Adj adj_identity() {
  Adj a;
  double *d = a.A;
#pragma unroll
  for (int i = 0; i < kAdjDoubles; ++i) d[i] = 0.0;
  a.A[0] = a.A[3] = 1.0;
  a.C[0] = a.C[4] = a.C[8] = 1.0;
  return a;
}
// cur after prev:  lambda -> cur(prev(lambda))
__device__ __forceinline__ Adj adj_after(const Adj &p, const Adj &c) {
  Adj o;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
#pragma unroll
    for (int q = 0; q < 2; ++q) o.A[2 * r + q] = c.A[2 * r] * p.A[q] + c.A[2 * r + 1] * p.A[2 + q];
    o.om[r] = c.A[2 * r] * p.om[0] + c.A[2 * r + 1] * p.om[1] + c.om[r];
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int q = 0; q < 2; ++q)
      o.B[2 * r + q] = c.B[2 * r] * p.A[q] + c.B[2 * r + 1] * p.A[2 + q] + c.C[3 * r] * p.B[q] + c.C[3 * r + 1] * p.B[2 + q] +
                       c.C[3 * r + 2] * p.B[4 + q];
#pragma unroll
    for (int q = 0; q < 3; ++q) o.C[3 * r + q] = c.C[3 * r] * p.C[q] + c.C[3 * r + 1] * p.C[3 + q] + c.C[3 * r + 2] * p.C[6 + q];
    o.oz[r] = c.B[2 * r] * p.om[0] + c.B[2 * r + 1] * p.om[1] + c.C[3 * r] * p.oz[0] + c.C[3 * r + 1] * p.oz[1] +
              c.C[3 * r + 2] * p.oz[2] + c.oz[r];
  }
  return o;
}
__device__ __forceinline__ void adj_apply(const Adj &a, double (&L)[5]) {
  const double m0 = a.A[0] * L[0] + a.A[1] * L[1] + a.om[0], m1 = a.A[2] * L[0] + a.A[3] * L[1] + a.om[1];
  double z[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    z[r] = a.B[2 * r] * L[0] + a.B[2 * r + 1] * L[1] + a.C[3 * r] * L[2] + a.C[3 * r + 1] * L[3] + a.C[3 * r + 2] * L[4] + a.oz[r];
  L[0] = m0;
  L[1] = m1;
  L[2] = z[0];
  L[3] = z[1];
  L[4] = z[2];
}

// forward quantities of step i, recomputed from the recorded state entering it
struct StepFwd {
  double f00, f01, f10, f11, Cf, Sf, dt;
  double m0, m1;
  double w00, w01, w10, w11;  // W = Phi Z
  double k0, k1, rD, w;       // k, 1/D, v/D
};
__device__ __forceinline__ StepFwd gp_step_fwd(const Sho &k, double dt, double yi, double di, double m0, double m1, double Z00,
                                               double Z01, double Z11, ShoCache &cc) {
  StepFwd s;
  s.dt = dt;
  sho_cs(k, s.dt, cc, s.Cf, s.Sf);
  s.f00 = fma(k.c, s.Sf, s.Cf);
  s.f01 = s.Sf;
  s.f10 = -k.w2 * s.Sf;
  s.f11 = fma(-k.c, s.Sf, s.Cf);
  s.m0 = m0;
  s.m1 = m1;
  s.w00 = s.f00 * Z00 + s.f01 * Z01;
  s.w01 = s.f00 * Z01 + s.f01 * Z11;
  s.w10 = s.f10 * Z00 + s.f11 * Z01;
  s.w11 = s.f10 * Z01 + s.f11 * Z11;
  const double p00 = s.w00 * s.f00 + s.w01 * s.f01, p01 = s.w00 * s.f10 + s.w01 * s.f11;
  s.k0 = k.a - p00;
  s.k1 = -p01;
  s.rD = xo::rcp_fast(s.k0 + di);
  s.w = (yi - (s.f00 * s.m0 + s.f01 * s.m1)) * s.rD;
  return s;
}

// lambda_i -> lambda_{i-1}.  kForce: include d l_i / d(v, D) (the affine part); kAcc: also return
// acc = (vbar, Dbar, abar, Phibar00, Phibar01, Phibar10, Phibar11) of this step.
template <bool kForce, bool kAcc>
__device__ __forceinline__ void gp_step_adj(const StepFwd &s, double (&L)[5], double *acc) {
  const double mb0 = L[0], mb1 = L[1], zb00 = L[2], zb01 = L[3], zb11 = L[4];
  const double km = s.k0 * mb0 + s.k1 * mb1;
  const double q = zb00 * s.k0 * s.k0 + zb01 * s.k0 * s.k1 + zb11 * s.k1 * s.k1;
  double vbar = km * s.rD;
  double Dbar = -(km * s.w + q * s.rD) * s.rD;
  if (kForce) {
    vbar -= s.w;
    Dbar += 0.5 * (s.w * s.w - s.rD);
  }
  const double kb0 = mb0 * s.w + (2.0 * zb00 * s.k0 + zb01 * s.k1) * s.rD;
  const double kb1 = mb1 * s.w + (zb01 * s.k0 + 2.0 * zb11 * s.k1) * s.rD;
  const double cb0 = mb0 - vbar, cb1 = mb1;
  const double pb00 = zb00 - kb0 - Dbar, hb = 0.5 * (zb01 - kb1), pb11 = zb11;
  // Zbar = Phi^T Sbar Phi, Sbar = [[pb00, hb], [hb, pb11]]
  const double u00 = pb00 * s.f00 + hb * s.f10, u01 = pb00 * s.f01 + hb * s.f11;
  const double u10 = hb * s.f00 + pb11 * s.f10, u11 = hb * s.f01 + pb11 * s.f11;
  L[2] = s.f00 * u00 + s.f10 * u10;
  L[3] = 2.0 * (s.f00 * u01 + s.f10 * u11);
  L[4] = s.f01 * u01 + s.f11 * u11;
  L[0] = s.f00 * cb0 + s.f10 * cb1;
  L[1] = s.f01 * cb0 + s.f11 * cb1;
  if (kAcc) {
    acc[0] = vbar;
    acc[1] = Dbar;
    acc[2] = kb0 + Dbar;
    // Phibar = cbar m^T + 2 Sbar W
    acc[3] = cb0 * s.m0 + 2.0 * (pb00 * s.w00 + hb * s.w10);
    acc[4] = cb0 * s.m1 + 2.0 * (pb00 * s.w01 + hb * s.w11);
    acc[5] = cb1 * s.m0 + 2.0 * (hb * s.w00 + pb11 * s.w10);
    acc[6] = cb1 * s.m1 + 2.0 * (hb * s.w01 + pb11 * s.w11);
  }
}

struct AdjOps {
  using E = Adj;
  static constexpr int kDoubles = kAdjDoubles;
  static __device__ __forceinline__ E identity() { return adj_identity(); }
  static __device__ __forceinline__ E combine(const E &earlier, const E &later) { return adj_after(earlier, later); }
  static __device__ __forceinline__ double *data(E &e) { return e.A; }
  static __device__ __forceinline__ const double *data(const E &e) { return e.A; }
};

// ---- one-pass reverse (adjoint) scan -------------------------------------------
// Same structure as the forward scan, run right-to-left: ticket -> (chain, reverse tile), lane l owns time-run
// 31 - l of a tile of kTileB = 32 x 8 cadences, so lane order = scan order.  The forward pass left only the state
// entering every run of 8 cadences (5 doubles per run, 5 B per cadence instead of 40); each lane re-derives the
// states inside its run into shared memory (lane-contiguous, conflict-free) and walks the run backwards twice:
// once pushing the origin and the five basis vectors (the run's affine map), once -- after the warp scan and the
// tree look-back over earlier tiles -- with the known incoming lambda, accumulating the hyper-parameter cotangents
// and emitting d/dy, d/d diag through shared memory (coalesced stores).  The last tile of a chain to finish folds
// the per-tile sums in tile order and chains (a, c, h, w2) -> (S0, w0, Q).
constexpr int kTileB = 32 * kRunB;
constexpr int kPitchB = kRunB + 1;

__global__ void __launch_bounds__(32, 12)
    gp_backward_kernel(const double *__restrict__ t, const double *__restrict__ y, const double *__restrict__ diag, int64_t N,
                       const double *__restrict__ S0, const double *__restrict__ w0, const double *__restrict__ Qf,
                       const double *__restrict__ jitter, double eps, int64_t Tb,
                       const double *__restrict__ rs /* [B][5][Tb] run-start states */, unsigned *__restrict__ counters,
                       const TreeBuffers tb, double *__restrict__ gpart /* [B][tiles][6] */, GpStatus *__restrict__ status,
                       double *__restrict__ gy, double *__restrict__ gdiag, double *__restrict__ ghyper /* [B][3] */,
                       double *__restrict__ gsum /* [B][2] or null */) {
  __shared__ double st[32 * kPitchB];
  __shared__ double sd[32 * kPitchB];        // diag in, d loglike / d diag out
  __shared__ double sy[32 * kPitchB];        // y in, d loglike / d y out
  __shared__ double sst[5][kRunB][32];       // state entering every step of every run
  const int lane = threadIdx.x;
  unsigned ticket = 0u;
  if (lane == 0) ticket = atomicAdd(&counters[0], 1u);
  ticket = __shfl_sync(kFullGp, ticket, 0);
  const int tiles = tb.n0;
  const int64_t bch = ticket / (unsigned)tiles;
  const int rtile = (int)(ticket % (unsigned)tiles);  // 0 = the LAST tile in time
  const int ttile = tiles - 1 - rtile;
  const Sho k = make_sho(S0[bch], w0[bch], Qf[bch], eps);
  const double jit = jitter ? jitter[bch] : 0.0;
  const int64_t base = (int64_t)ttile * kTileB;
  const int nvalid = (int)min((int64_t)kTileB, N - base);
  const double *yb = y + bch * N;
#pragma unroll 4
  for (int idx = lane; idx < kTileB; idx += 32) {
    const int at = (idx / kRunB) * kPitchB + idx % kRunB;
    if (idx < nvalid) {
      st[at] = __ldg(t + base + idx);
      sd[at] = __ldg(diag + base + idx) + jit;
      sy[at] = __ldg(yb + base + idx);
    }
  }
  const double t_halo = __ldg(t + (base > 0 ? base - 1 : 0));
  __syncwarp();
  const int r = 31 - lane;  // time-run inside the tile
  const int64_t s = base + (int64_t)r * kRunB;
  const int len = (int)max((int64_t)0, min((int64_t)kRunB, N - s));
  double *ts = st + r * kPitchB, *ds = sd + r * kPitchB, *ys = sy + r * kPitchB;
  const double t_before = r > 0 ? st[(r - 1) * kPitchB + kRunB - 1] : t_halo;
  // states entering every step of the run, re-derived from the run-start state the forward pass left
  if (len > 0) {
    const int64_t run = s / kRunB;
    const double *rp = rs + (size_t)bch * 5 * Tb + run;
    double m0 = __ldg(rp), m1 = __ldg(rp + Tb), Z00 = __ldg(rp + 2 * Tb), Z01 = __ldg(rp + 3 * Tb), Z11 = __ldg(rp + 4 * Tb);
    double tprev = t_before;
    ShoCache cc;
#pragma unroll 1
    for (int i = 0; i < len; ++i) {
      sst[0][i][lane] = m0;
      sst[1][i][lane] = m1;
      sst[2][i][lane] = Z00;
      sst[3][i][lane] = Z01;
      sst[4][i][lane] = Z11;
      if (i + 1 == len) break;
      const double ti = ts[i];
      double f00, f01, f10, f11;
      sho_phi(k, ti - tprev, cc, f00, f01, f10, f11);
      tprev = ti;
      double c0 = f00 * m0 + f01 * m1, c1 = f10 * m0 + f11 * m1;
      double z00 = f00 * Z00 + f01 * Z01, z01 = f00 * Z01 + f01 * Z11;
      double z10 = f10 * Z00 + f11 * Z01, z11 = f10 * Z01 + f11 * Z11;
      double p00 = z00 * f00 + z01 * f01, p01 = z00 * f10 + z01 * f11, p11 = z10 * f10 + z11 * f11;
      const double pm = k.a - p00;
      const double D = pm + ds[i];
      const double rD = xo::rcp_fast(D);
      const double K0 = pm * rD, K1 = -p01 * rD;
      const double vv = ys[i] - c0;
      m0 = fma(K0, vv, c0);
      m1 = fma(K1, vv, c1);
      Z00 = fma(D * K0, K0, p00);
      Z01 = fma(D * K0, K1, p01);
      Z11 = fma(D * K1, K1, p11);
    }
  }
  auto step_fwd = [&](int i, ShoCache &cc) {
    const double tp = i > 0 ? ts[i - 1] : t_before;
    return gp_step_fwd(k, ts[i] - tp, ys[i], ds[i], sst[0][i][lane], sst[1][i][lane], sst[2][i][lane], sst[3][i][lane],
                       sst[4][i][lane], cc);
  };
  // the run's map lambda_{e-1} -> lambda_{s-1}: push the origin (with the forcing) and the five basis vectors through it
  Adj v = adj_identity();
  if (len > 0) {
    double O[5] = {0, 0, 0, 0, 0}, X0[5] = {1, 0, 0, 0, 0}, X1[5] = {0, 1, 0, 0, 0}, X2[5] = {0, 0, 1, 0, 0},
           X3[5] = {0, 0, 0, 1, 0}, X4[5] = {0, 0, 0, 0, 1};
    ShoCache cc;
#pragma unroll 1
    for (int i = len - 1; i >= 0; --i) {
      const StepFwd f = step_fwd(i, cc);
      gp_step_adj<true, false>(f, O, nullptr);
      gp_step_adj<false, false>(f, X0, nullptr);
      gp_step_adj<false, false>(f, X1, nullptr);
      gp_step_adj<false, false>(f, X2, nullptr);
      gp_step_adj<false, false>(f, X3, nullptr);
      gp_step_adj<false, false>(f, X4, nullptr);
    }
    v.A[0] = X0[0]; v.A[1] = X1[0]; v.A[2] = X0[1]; v.A[3] = X1[1];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      v.B[2 * q] = X0[2 + q];
      v.B[2 * q + 1] = X1[2 + q];
      v.C[3 * q] = X2[2 + q];
      v.C[3 * q + 1] = X3[2 + q];
      v.C[3 * q + 2] = X4[2 + q];
      v.oz[q] = O[2 + q];
    }
    v.om[0] = O[0];
    v.om[1] = O[1];
  }
  v = warp_inclusive_scan<AdjOps>(v);
  const Adj pre = elem_shfl_up<AdjOps>(v, 1);
  bool ok = true;
  const Adj excl = tree_exclusive<AdjOps>(tb.chain(bch, kAdjDoubles), rtile, elem_shfl<AdjOps>(v, 31), ok);
  if (!__all_sync(kFullGp, ok) && lane == 0) atomicAdd(&status->wait_failures, 1ull);
  double ga = 0.0, gc = 0.0, gh = 0.0, gw2 = 0.0;  // cotangents of a, c, h and w2 = c^2 +- h^2
  double sy_ = 0.0, sd_ = 0.0;                      // sums of d/dy and d/d diag over the run
  if (len > 0) {
    // the tree's map applied to lambda = 0 beyond the last cadence, then the in-warp prefix
    double L[5] = {excl.om[0], excl.om[1], excl.oz[0], excl.oz[1], excl.oz[2]};
    if (lane > 0) adj_apply(pre, L);
    ShoCache cc;
#pragma unroll 1
    for (int i = len - 1; i >= 0; --i) {
      const StepFwd f = step_fwd(i, cc);
      double acc[7];
      gp_step_adj<true, true>(f, L, acc);
      ys[i] = acc[0];  // d / d y_i     (y_i, diag_i are dead from here on: every step reads its own only)
      ds[i] = acc[1];  // d / d diag_i
      sy_ += acc[0];
      sd_ += acc[1];
      ga += acc[2];
      // Phi = [[Cf + c Sf, Sf], [-w2 Sf, Cf - c Sf]]
      const double d0011 = acc[3] - acc[6];
      const double Cb = acc[3] + acc[6];
      const double Sb = k.c * d0011 + acc[4] - k.w2 * acc[5];
      gc += f.Sf * d0011 - f.dt * (f.Cf * Cb + f.Sf * Sb);
      gw2 -= f.Sf * acc[5];
      const double dCdh = (k.under ? -f.dt : f.dt) * k.h * f.Sf;
      gh += dCdh * Cb + (f.dt * f.Cf - f.Sf) / k.h * Sb;
    }
    if (!isfinite(ga + gc + gh + gw2)) atomicAdd(&status->bad_lanes, 1ull);
  }
  double vals[6] = {ga, gc, gh, gw2, sy_, sd_};
#pragma unroll
  for (int j = 0; j < 6; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vals[j] += __shfl_xor_sync(kFullGp, vals[j], o);
  }
  __syncwarp();
  // per-cadence gradients out, coalesced
  if (gy || gdiag) {
#pragma unroll 4
    for (int idx = lane; idx < kTileB; idx += 32) {
      if (idx < nvalid) {
        const int at = (idx / kRunB) * kPitchB + idx % kRunB;
        if (gy) gy[bch * N + base + idx] = sy[at];
        if (gdiag) gdiag[bch * N + base + idx] = sd[at];
      }
    }
  }
  int is_last = 0;
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 6; ++j) gpart[((size_t)bch * tiles + rtile) * 6 + j] = vals[j];
    __threadfence();
    is_last = atomicAdd(&counters[1 + bch], 1u) == (unsigned)(tiles - 1);
  }
  if (__shfl_sync(kFullGp, is_last, 0)) {
    // fold the per-tile cotangents of (a, c, h, w2) in tile order and chain them to (S0, w0, Q):
    //   a = S0 w0 Q,  c = w0 / (2Q),  h = c f,  f = sqrt(max(+-(4Q^2 - 1), eps)),  w2 = c^2 +- h^2
    __threadfence();
    double fa = 0.0, fc = 0.0, fh = 0.0, fw2 = 0.0, fy = 0.0, fd = 0.0;
    for (int i = lane; i < tiles; i += 32) {
      const double *p = gpart + ((size_t)bch * tiles + i) * 6;
      fa += __ldcg(p);
      fc += __ldcg(p + 1);
      fh += __ldcg(p + 2);
      fw2 += __ldcg(p + 3);
      fy += __ldcg(p + 4);
      fd += __ldcg(p + 5);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      fa += __shfl_xor_sync(kFullGp, fa, o);
      fc += __shfl_xor_sync(kFullGp, fc, o);
      fh += __shfl_xor_sync(kFullGp, fh, o);
      fw2 += __shfl_xor_sync(kFullGp, fw2, o);
      fy += __shfl_xor_sync(kFullGp, fy, o);
      fd += __shfl_xor_sync(kFullGp, fd, o);
    }
    if (lane == 0 && gsum) {
      gsum[2 * bch] = fy;
      gsum[2 * bch + 1] = fd;
    }
    if (lane == 0) {
      const double sgn = k.under ? 1.0 : -1.0;
      fc += 2.0 * k.c * fw2;
      fh += sgn * 2.0 * k.h * fw2;
      const double s0 = S0[bch], w = w0[bch], Q = Qf[bch];
      const double arg = sgn * (4.0 * Q * Q - 1.0);
      const double f = sqrt(fmax(arg, eps));
      const double dfdQ = arg > eps ? sgn * 4.0 * Q / f : 0.0;
      const double dcdQ = -k.c / Q;
      ghyper[3 * bch + 0] = fa * w * Q;
      ghyper[3 * bch + 1] = fa * s0 * Q + fc * (0.5 / Q) + fh * (0.5 / Q) * f;
      ghyper[3 * bch + 2] = fa * s0 * w + fc * dcdQ + fh * (dcdQ * f + k.c * dfdQ);
    }
  }
}

struct TreeOffsets {  // byte offsets of one tree's scratch inside the workspace
  int n0, n1, n2;
  size_t off_counters, off_f0, off_f1, off_f2, off_e0, off_e1, off_e2;
};
struct GpLayout {
  TreeOffsets fwd, bwd;  // forward: tiles of kTile cadences; reverse (adjoint): tiles of kTileB cadences
  int64_t Tb;            // runs of kRunB cadences per chain
  int nctas_b;           // smoother pass (three-kernel form): CTAs per chain
  size_t zero_bytes;     // status word + counters + flags at the head of the workspace: cleared before every call
  size_t off_ll, off_rs, off_gpart;
  size_t off_rec, off_bpref, off_bcta, off_bstate, bytes;
};
inline GpLayout gp_layout(int64_t N, int64_t B, bool solve = false, bool grad = false) {
  GpLayout L;
  L.Tb = (N + kRunB - 1) / kRunB;
  L.nctas_b = (int)((L.Tb + kThreadsGp - 1) / kThreadsGp);
  size_t o = 256;  // GpStatus
  auto take_bytes = [&](size_t n) {
    size_t at = o;
    o += (n + 255) & ~(size_t)255;
    return at;
  };
  auto take = [&](size_t n) { return take_bytes(n * sizeof(double)); };
  auto tree_counts = [&](TreeOffsets &T, int64_t tile) {
    T.n0 = (int)((N + tile - 1) / tile);
    T.n1 = (T.n0 + 31) / 32;
    T.n2 = (T.n1 + 31) / 32;
  };
  auto tree_flags = [&](TreeOffsets &T) {
    T.off_counters = take_bytes((size_t)(1 + B) * sizeof(unsigned));
    T.off_f0 = take_bytes((size_t)B * T.n0 * sizeof(unsigned));
    T.off_f1 = take_bytes((size_t)B * T.n1 * sizeof(unsigned));
    T.off_f2 = take_bytes((size_t)B * T.n2 * sizeof(unsigned));
  };
  auto tree_elems = [&](TreeOffsets &T, int doubles) {
    T.off_e0 = take((size_t)B * T.n0 * doubles);
    T.off_e1 = take((size_t)B * T.n1 * doubles);
    T.off_e2 = take((size_t)B * T.n2 * doubles);
  };
  tree_counts(L.fwd, kTile);
  tree_counts(L.bwd, kTileB);
  tree_flags(L.fwd);
  if (grad) tree_flags(L.bwd);
  L.zero_bytes = o;
  tree_elems(L.fwd, kAggDoubles);
  L.off_ll = take((size_t)B * L.fwd.n0);
  L.off_rec = L.off_bpref = L.off_bcta = L.off_bstate = 0;
  if (solve) {
    L.off_rec = take((size_t)B * 3 * N);
    L.off_bpref = take((size_t)B * L.Tb * 6);
    L.off_bcta = take((size_t)B * L.nctas_b * 6);
    L.off_bstate = take((size_t)B * L.nctas_b * 2);
  }
  L.off_rs = L.off_gpart = 0;
  if (grad) {
    L.off_rs = take((size_t)B * 5 * L.Tb);
    tree_elems(L.bwd, kAdjDoubles);
    L.off_gpart = take((size_t)B * L.bwd.n0 * 6);
  }
  L.bytes = o;
  return L;
}
inline TreeBuffers tree_buffers(char *ws, const TreeOffsets &T) {
  auto d = [&](size_t off) { return reinterpret_cast<double *>(ws + off); };
  auto u = [&](size_t off) { return reinterpret_cast<unsigned *>(ws + off); };
  return TreeBuffers{d(T.off_e0), d(T.off_e1), d(T.off_e2), u(T.off_f0), u(T.off_f1), u(T.off_f2), T.n0, T.n1, T.n2};
}

// the one-pass forward scan: clear status/counters/flags (one stream-ordered memset), one kernel
template <int REC>
int gp_forward(const double *t, const double *y, const double *diag, int64_t N, const double *S0, const double *w0, const double *Q,
               const double *jitter, int64_t B, double eps, double *loglike, double *rec, char *ws, const GpLayout &L, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(ws, 0, L.zero_bytes, st);
  if (e != cudaSuccess) return -(int)e;
  xo::launch(gp_forward_kernel<REC>, dim3((unsigned)((size_t)B * L.fwd.n0)), dim3(32), 0, st, t, y, diag, N, S0, w0, Q, jitter, eps,
             reinterpret_cast<unsigned *>(ws + L.fwd.off_counters), tree_buffers(ws, L.fwd), reinterpret_cast<double *>(ws + L.off_ll),
             reinterpret_cast<GpStatus *>(ws), loglike, rec, L.Tb);
  return 0;
}

}  // namespace

namespace xo {

// xo_celerite_sho_grad plus a per-chain jitter term added to diag and the per-chain sums gsum[B][2] =
// (sum_n d/dy_n, sum_n d/d diag_n): what the joint transit + GP model needs (d/d mean = -gsum[0],
// d/d jitter = gsum[1]).  Two kernels and one memset, all on `stream`.
int celerite_sho_grad_impl(const double *t, const double *y, const double *diag, int64_t N, const double *S0, const double *w0,
                           const double *Q, const double *jitter, int64_t B, double eps, double *loglike, double *ghyper,
                           double *gy, double *gdiag, double *gsum, void *workspace, size_t workspace_bytes, void *stream) {
  if (N < 0 || B < 0) return XO_EINVAL;
  if (B == 0) return 0;
  if (N == 0) return XO_EINVAL;
  if (B > 65535 || (double)B * (double)((N + kTileB - 1) / kTileB) > 2.0e9) return XO_ERANGE;
  if (!t || !y || !diag || !S0 || !w0 || !Q || !loglike || !ghyper || !workspace) return XO_EINVAL;
  GpLayout L = gp_layout(N, B, false, true);
  if (workspace_bytes < L.bytes) return XO_EWORKSPACE;
  char *ws = static_cast<char *>(workspace);
  auto at = [&](size_t off) { return reinterpret_cast<double *>(ws + off); };
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  XO_NVTX_SCOPE("xo_celerite_sho_grad");
  int rc = gp_forward<2>(t, y, diag, N, S0, w0, Q, jitter, B, eps, loglike, at(L.off_rs), ws, L, st);
  if (rc) return rc;
  xo::launch(gp_backward_kernel, dim3((unsigned)((size_t)B * L.bwd.n0)), dim3(32), 0, st, t, y, diag, N, S0, w0, Q, jitter, eps, L.Tb,
             (const double *)at(L.off_rs), reinterpret_cast<unsigned *>(ws + L.bwd.off_counters), tree_buffers(ws, L.bwd), at(L.off_gpart),
             reinterpret_cast<GpStatus *>(ws), gy, gdiag, ghyper, gsum);
  return xo::launch_status();
}

}  // namespace xo

extern "C" {

size_t xo_celerite_sho_workspace(int64_t N, int64_t B) {
  if (N <= 0 || B <= 0) return 0;
  return gp_layout(N, B).bytes;
}

int xo_celerite_sho_loglike(const double *t, const double *y, const double *diag, int64_t N, const double *S0,
                            const double *w0, const double *Q, int64_t B, double eps, double *loglike, void *workspace,
                            size_t workspace_bytes, void *stream) {
  if (N < 0 || B < 0) return XO_EINVAL;
  if (B == 0) return 0;
  if (N == 0) return XO_EINVAL;
  if ((double)B * (double)((N + kTileB - 1) / kTileB) > 2.0e9) return XO_ERANGE;
  if (!t || !y || !diag || !S0 || !w0 || !Q || !loglike || !workspace) return XO_EINVAL;
  XO_NVTX_SCOPE("xo_celerite_sho_loglike");
  GpLayout L = gp_layout(N, B);
  if (workspace_bytes < L.bytes) return XO_EWORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = gp_forward<0>(t, y, diag, N, S0, w0, Q, nullptr, B, eps, loglike, nullptr, static_cast<char *>(workspace), L, st);
  return rc ? rc : xo::launch_status();
}

size_t xo_celerite_sho_solve_workspace(int64_t N, int64_t B) {
  if (N <= 0 || B <= 0) return 0;
  return gp_layout(N, B, true).bytes;
}

// log-likelihood AND alpha = K^-1 y (so that d loglike / d y = -alpha): forward scan, then the
// disturbance smoother as a reverse affine scan.  alpha[B][N].
int xo_celerite_sho_solve(const double *t, const double *y, const double *diag, int64_t N, const double *S0,
                          const double *w0, const double *Q, int64_t B, double eps, double *loglike, double *alpha,
                          void *workspace, size_t workspace_bytes, void *stream) {
  if (N < 0 || B < 0) return XO_EINVAL;
  if (B == 0) return 0;
  if (N == 0) return XO_EINVAL;
  if (B > 65535 || (double)B * (double)((N + kTileB - 1) / kTileB) > 2.0e9) return XO_ERANGE;
  if (!t || !y || !diag || !S0 || !w0 || !Q || !loglike || !alpha || !workspace) return XO_EINVAL;
  GpLayout L = gp_layout(N, B, true);
  if (workspace_bytes < L.bytes) return XO_EWORKSPACE;
  char *ws = static_cast<char *>(workspace);
  auto at = [&](size_t off) { return reinterpret_cast<double *>(ws + off); };
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  XO_NVTX_SCOPE("xo_celerite_sho_solve");
  int rc = gp_forward<1>(t, y, diag, N, S0, w0, Q, nullptr, B, eps, loglike, at(L.off_rec), ws, L, st);
  if (rc) return rc;
  dim3 gridb(L.nctas_b, (unsigned)B);
  xo::launch(gp_bwd_aggregate_kernel, gridb, kThreadsGp, 0, st, t, N, S0, w0, Q, eps, L.Tb, at(L.off_rec), at(L.off_bpref),
                                                       at(L.off_bcta));
  xo::launch(gp_bwd_cta_scan_kernel, (unsigned)B, kThreadsGp, 0, st, at(L.off_bcta), L.nctas_b, at(L.off_bstate));
  xo::launch(gp_bwd_solve_kernel, gridb, kThreadsGp, 0, st, t, N, S0, w0, Q, eps, L.Tb, at(L.off_rec), at(L.off_bpref),
                                                   at(L.off_bstate), alpha);
  return xo::launch_status();
}

size_t xo_celerite_sho_grad_workspace(int64_t N, int64_t B) {
  if (N <= 0 || B <= 0) return 0;
  return gp_layout(N, B, false, true).bytes;
}

// log-likelihood and its exact reverse-mode gradient: ghyper[B][3] = d/d(S0, w0, Q), and (each
// optional) gy[B][N] = d/dy (= -K^-1 y), gdiag[B][N] = d/d diag.
int xo_celerite_sho_grad(const double *t, const double *y, const double *diag, int64_t N, const double *S0,
                         const double *w0, const double *Q, int64_t B, double eps, double *loglike, double *ghyper,
                         double *gy, double *gdiag, void *workspace, size_t workspace_bytes, void *stream) {
  return xo::celerite_sho_grad_impl(t, y, diag, N, S0, w0, Q, nullptr, B, eps, loglike, ghyper, gy, gdiag, nullptr, workspace,
                                    workspace_bytes, stream);
}

}  // extern "C"
