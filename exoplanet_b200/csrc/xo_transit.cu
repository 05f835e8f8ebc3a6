// xo_transit.cu -- the fused L3 path over a (chains x planets x cadences) batch.
//
// Replaces, in one launch each, the graph PyTensor builds from
//   /root/reference/src/exoplanet/light_curves/limb_dark.py:163-232 (get_light_curve)
//   /root/reference/src/exoplanet/orbits/keplerian.py:324-334, 380-409 (anomaly, position)
//   /root/reference/src/exoplanet/orbits/keplerian.py:729-775 (in_transit window)
// plus the Gaussian likelihood and the whole reverse pass.
//
// Data layout in HBM: t, y, ivar are [N] and shared by all chains (L2-resident after
// the first chain touches them); per-chain state is the element block [C][P][16],
// cl [C][3], mean [C]; outputs are O(#params) per chain.
//
// Design of the logp+grad kernel (the dominant kernel):
//  * one CTA = 4 warps working on one chain and a contiguous part of its cadences; which chain and part comes from
//    a table every CTA derives from per-chain work estimates (weighted_partition), so that a launch of whole waves
//    keeps every CTA slot busy with the same amount of work; the chain's element block, limb-darkening vector and
//    exposure stencil are staged once in shared memory;
//  * in-transit (cadence, planet) samples are evaluated 32 at a time with all lanes busy on the expensive part
//    (Kepler solve -> projection -> limb-darkening integral -> analytic VJP).  With the data moments, sorted t and the
//    time index the CTA enumerates the transit epochs that touch its cadences (one per thread: window -> cadence
//    range by table look-up), a block scan numbers the samples, and the warps take them 32 at a time; no time stamp
//    is read, nothing is queued.  Otherwise each warp steps through its cadences (with moments: 32 steps tested per
//    ballot, only hit steps loaded; without: every cadence read and the out-of-transit Gaussian terms summed on the
//    way), appends in-window cadences to per-warp, per-planet queues and drains a queue whenever it holds >= 32;
//  * per-lane gradient sums live in shared memory and are added up over the warp in a fixed order once per warp
//    (or when the drained planet changes), so results are bit-reproducible run to run; a second tiny kernel (or the
//    model epilogue) folds the per-warp partials;
//  * launched as a programmatic dependent of the model prologue: everything before pdl_wait() (accumulator
//    zeroing) overlaps the prologue's tail.
#include <cuda_runtime.h>

#include <cstdlib>

#include "xo_common.cuh"
#include "xo_math.cuh"
#include "xo_prologue.cuh"
#include "xo_transit.cuh"

namespace {

#ifndef XO_TRANSIT_WARPS
#define XO_TRANSIT_WARPS 4  // warps per CTA of the logp kernel (build-time knob for occupancy experiments)
#endif
constexpr int kWarps = XO_TRANSIT_WARPS;
constexpr int kThreads = kWarps * 32;
using xo::kNBase;
using xo::Split;
constexpr unsigned kFull = 0xffffffffu;
// How the drain of the logp kernel reads its chain's elements (see lc_point): as volatile shared memory, i.e. re-loaded at
// every use (a broadcast LDS) instead of held in registers across the whole sample evaluation.  That is what lets the kernel
// fit 128 registers with 36 B of spills (plain pointer: 196 B) and run 4 CTAs per SM: 54.5 -> 52.9 us per 128-chain step,
// 4.21e6 -> 4.46e6 evals/s at 1024 chains (profiles/r02_tuning_sweep.txt).  -DXO_ELEM_PLAIN restores the plain pointer.
#ifdef XO_ELEM_PLAIN
using DrainElemPtr = const double *;
#else
using DrainElemPtr = const volatile double *;
#endif

struct ChainSmem {
  double elem[XO_MAX_PLANETS][XO_NEL];
  double rperiod[XO_MAX_PLANETS];
  double wlo[XO_MAX_PLANETS], whi[XO_MAX_PLANETS];  // window in units of the period (phase); widened by ttv_dmax for the sweep
  double ttv_dmax[XO_MAX_PLANETS];                  // TTV: largest departure of a transit time from t0 + j P (0 without a table)
  double cl[3];
  double mean;
  double dtk[XO_MAX_STENCIL];
  double wk[XO_MAX_STENCIL];
};

// ---- transit-time tables (TTVOrbit, /root/reference/src/exoplanet/orbits/ttv.py:149-187) -------------------------------
// index of the transit nearest to t: the number of mid-points between consecutive transit times that lie below t
// (ttv.py:159-178: searchsorted over the bin edges; the two outermost edges select the same values as their neighbours)
__device__ __forceinline__ int ttv_index(const double *__restrict__ tt, int n, double t) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (0.5 * (__ldg(tt + mid) + __ldg(tt + mid + 1)) < t)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}
// centre of "epoch" k of a table: the table entry inside it, the first / last transit continued with period P outside
__device__ __forceinline__ double ttv_centre(const double *__restrict__ tt, int n, double per, long long k) {
  if (k < 0) return fma((double)k, per, __ldg(tt));
  if (k >= n) return fma((double)(k - n + 1), per, __ldg(tt + n - 1));
  return __ldg(tt + k);
}
struct TtvView {  // the table of (chain c, planet p), or n = 0
  const double *tt;
  int n;
};
__device__ __forceinline__ TtvView ttv_view(const double *times, const int *count, int nt, int64_t c, int P, int p) {
  return times ? TtvView{times + ((size_t)c * P + p) * nt, count[p]} : TtvView{nullptr, 0};
}
// time shift that maps a cadence at t onto the element block's own ephemeris: t - (tt[j] - t0)
__device__ __forceinline__ double ttv_shift(const TtvView &v, double t0, double t, int *j_out = nullptr) {
  if (v.n <= 0) return 0.0;
  const int j = ttv_index(v.tt, v.n, t);
  if (j_out) *j_out = j;
  return __ldg(v.tt + j) - t0;
}
// largest |tt[j] - (t0 + j P)| of a table: how far the sweep's periodic window test must be widened
__device__ __forceinline__ double ttv_departure(const TtvView &v, double t0, double per) {
  double d = 0.0;
  for (int j = 0; j < v.n; ++j) d = fmax(d, fabs(__ldg(v.tt + j) - fma((double)j, per, t0)));
  return d;
}

__device__ __forceinline__ void stage_derived(ChainSmem &sm, int P) {
  if (threadIdx.x < P) {
    const double rp = 1.0 / sm.elem[threadIdx.x][XO_EL_PERIOD];
    sm.rperiod[threadIdx.x] = rp;
    sm.wlo[threadIdx.x] = sm.elem[threadIdx.x][XO_EL_TSTART] * rp;
    sm.whi[threadIdx.x] = sm.elem[threadIdx.x][XO_EL_TEND] * rp;
    sm.ttv_dmax[threadIdx.x] = 0.0;
  }
  __syncthreads();
}
// with a transit-time table the sweep's periodic test runs on the linear ephemeris (t0, P) with the window widened by the
// largest departure of the table from it: a superset of the in-transit cadences (the extra ones evaluate to exactly 0)
__device__ __forceinline__ void stage_ttv(ChainSmem &sm, const double *times, const int *count, int nt, int64_t c, int P) {
  if (!times) return;
  if (threadIdx.x < P) {
    const int p = threadIdx.x;
    const double d = ttv_departure(ttv_view(times, count, nt, c, P, p), sm.elem[p][XO_EL_T0], sm.elem[p][XO_EL_PERIOD]);
    sm.ttv_dmax[p] = d;
    sm.wlo[p] -= d * sm.rperiod[p];
    sm.whi[p] += d * sm.rperiod[p];
  }
  __syncthreads();
}

__device__ __forceinline__ void stage_chain(ChainSmem &sm, const double *elem, const double *cl, const double *mean,
                                            const double *dtk, const double *wk, int64_t c, int P, int K) {
  for (int i = threadIdx.x; i < P * XO_NEL; i += blockDim.x) sm.elem[i / XO_NEL][i % XO_NEL] = elem[c * P * XO_NEL + i];
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    sm.dtk[i] = dtk[i];
    sm.wk[i] = wk[i];
  }
  if (threadIdx.x < 3) sm.cl[threadIdx.x] = cl[3 * c + threadIdx.x];
  if (threadIdx.x == 3) sm.mean = mean ? mean[c] : 0.0;
  __syncthreads();
  stage_derived(sm, P);
}

// single-launch model mode: the CTA derives its chain's element block from the sampler-side
// parameters itself (warp 0, two lanes per planet) instead of reading one a prologue kernel wrote
__device__ __forceinline__ void stage_chain_from_params(ChainSmem &sm, const xo::ModelArgs &ma, const double *mean,
                                                        const double *dtk, const double *wk, int64_t c, int P, int K) {
  if (threadIdx.x < 32) {
    const int p = threadIdx.x >> 1;
    const bool active = p < P;
    const int pp = active ? p : 0;
    xo::prologue_pair(ma.par + (c * P + pp) * XO_NPAR, ma.star + c * XO_NSTAR, ma.texp, ma.use_in_transit, threadIdx.x & 1, active,
                      sm.elem[pp]);
    if (threadIdx.x == 0) xo::limb_coeffs(ma.star[c * XO_NSTAR + 2], ma.star[c * XO_NSTAR + 3], sm.cl);
  } else {
    for (int i = threadIdx.x - 32; i < K; i += blockDim.x - 32) {
      sm.dtk[i] = dtk[i];
      sm.wk[i] = wk[i];
    }
    if (threadIdx.x == 32) sm.mean = mean[c];
  }
  __syncthreads();
  stage_derived(sm, P);
}

// window test of the sweep: phase = frac((t - t0)/P) in [-1/2, 1/2] against [tstart, tend]/P.
// Same set as keplerian.py:731,769 up to rounding at the window edges (where lc = 0 anyway).
__device__ __forceinline__ bool in_window_phase(const ChainSmem &sm, int p, double t) {
  const double u = (t - sm.elem[p][XO_EL_T0]) * sm.rperiod[p];
  const double ph = u - rint(u);
  return (ph >= sm.wlo[p]) & (ph <= sm.whi[p]);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// ---------------------------------------------------------------------------
// light curve forward: lc[C][N][P]
// ---------------------------------------------------------------------------
template <bool kDelay>
__global__ void __launch_bounds__(256) transit_lc_fwd_kernel(const double *__restrict__ t, int64_t N,
                                                             const double *__restrict__ elem,
                                                             const double *__restrict__ cl, int P,
                                                             const double *__restrict__ dtk,
                                                             const double *__restrict__ wk, int K,
                                                             const double *__restrict__ ttv_times,
                                                             const int *__restrict__ ttv_count, int ttv_nt,
                                                             const double *__restrict__ texp_vec, double *__restrict__ lc) {
  __shared__ ChainSmem sm;
  const int64_t c = blockIdx.y;
  stage_chain(sm, elem, cl, nullptr, dtk, wk, c, P, K);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    const double t_i = __ldg(t + i);
    // per-cadence exposure times (limb_dark.py:199-205): dtk then holds the unit stencil, scaled here
    const double tx = texp_vec ? __ldg(texp_vec + i) : 1.0;
    for (int p = 0; p < P; ++p) {
      const double tt = t_i - ttv_shift(ttv_view(ttv_times, ttv_count, ttv_nt, c, P, p), sm.elem[p][XO_EL_T0], t_i);
      double acc = 0.0;
      if (xo::in_window(sm.elem[p], sm.rperiod[p], tt)) {
        for (int k = 0; k < K; ++k)
          acc = fma(sm.wk[k], xo::lc_point<false, 1, kDelay>(sm.elem[p], sm.cl, tt + __dmul_rn(tx, sm.dtk[k]), 0.0, nullptr, nullptr), acc);
      }
      __stcs(lc + (c * N + i) * P + p, acc);
    }
  }
}

// residual of the light curve model: resid[C][N] = y - mean - sum_p lc  (the input of a noise model, e.g. the
// celerite scan of the joint transit + GP likelihood); same evaluation as transit_lc_fwd_kernel
__global__ void __launch_bounds__(256) transit_resid_kernel(const double *__restrict__ t, const double *__restrict__ y, int64_t N,
                                                            const double *__restrict__ elem, const double *__restrict__ cl,
                                                            const double *__restrict__ mean, int P,
                                                            const double *__restrict__ dtk, const double *__restrict__ wk, int K,
                                                            const double *__restrict__ ttv_times,
                                                            const int *__restrict__ ttv_count, int ttv_nt,
                                                            double *__restrict__ resid) {
  __shared__ ChainSmem sm;
  const int64_t c = blockIdx.y;
  stage_chain(sm, elem, cl, mean, dtk, wk, c, P, K);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    const double t_i = __ldg(t + i);
    double model = sm.mean;
    for (int p = 0; p < P; ++p) {
      const double tt = t_i - ttv_shift(ttv_view(ttv_times, ttv_count, ttv_nt, c, P, p), sm.elem[p][XO_EL_T0], t_i);
      if (xo::in_window(sm.elem[p], sm.rperiod[p], tt)) {
        for (int k = 0; k < K; ++k)
          model = fma(sm.wk[k], xo::lc_point<false>(sm.elem[p], sm.cl, tt + sm.dtk[k], 0.0, nullptr, nullptr), model);
      }
    }
    resid[c * N + i] = __ldg(y + i) - model;
  }
}

// ---------------------------------------------------------------------------
// time -> cadence index (data only, built once per data set by xo_transit_data_index, stored behind the
// moments in the same block): table[j] = number of cadences with t < t[0] + j w, j = 0..nb, w = (t[N-1] - t[0]) / nb.
// With it the sweep of the logp kernel does not look at the time stamps at all: for every transit epoch that
// touches a warp's cadence range it gets the index range of the in-transit window from two table look-ups and a
// short binary search inside one bin (lanes work on 32 epochs at a time).
// ---------------------------------------------------------------------------
struct TimeIndex {
  const uint32_t *table;
  int nb;
  double tmin, inv_w;
};
// first index i in [0, N] with t[i] >= x (kUpper: t[i] > x); t sorted ascending
template <bool kUpper>
__device__ __forceinline__ uint32_t index_bound(const double *__restrict__ t, uint32_t N, const TimeIndex &ix, double x) {
  double u = (x - ix.tmin) * ix.inv_w;
  u = u > 0.0 ? u : 0.0;  // also NaN -> 0
  u = fmin(u, (double)(ix.nb - 1));
  const int j = (int)u;
  uint32_t lo = __ldg(ix.table + j), hi = __ldg(ix.table + j + 1);
  auto before = [&](double v) { return kUpper ? (v <= x) : (v < x); };  // t[i] lies before the bound
  // the bin was picked in floating point: make sure the answer is inside [lo, hi], else search the whole array
  if (lo > 0u && !before(__ldg(t + lo - 1))) lo = 0u;
  if (hi < N && before(__ldg(t + hi))) hi = N;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (before(__ldg(t + mid)))
      lo = mid + 1u;
    else
      hi = mid;
  }
  return lo;
}

__global__ void transit_index_kernel(const double *__restrict__ t, int64_t N, double *__restrict__ moments, int nb) {
  uint32_t *table = reinterpret_cast<uint32_t *>(moments + XO_MOMENTS_LEN);
  const double tmin = __ldg(t), tmax = __ldg(t + N - 1);
  const double w = (tmax - tmin) / (double)nb;
  const bool usable = moments[4] != 0.0 && w > 0.0 && isfinite(w);  // sorted (transit_moments_kernel ran before), non-degenerate
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0) {
    moments[5] = usable ? (double)nb : 0.0;
    moments[6] = usable ? 1.0 / w : 0.0;
  }
  if (j > nb || !usable) return;
  uint32_t lo = 0u, hi = (uint32_t)N;
  if (j < nb) {
    const double x = fma((double)j, w, tmin);
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (__ldg(t + mid) < x)
        lo = mid + 1u;
      else
        hi = mid;
    }
  } else {
    lo = (uint32_t)N;
  }
  table[j] = lo;
}

inline int index_bins(int64_t N) {
  int64_t nb = N / 2;
  if (nb < 1) nb = 1;
  if (nb > (1 << 22)) nb = 1 << 22;
  return (int)nb;
}

// ---------------------------------------------------------------------------
// logp + grad (see header comment).  kMinBlocks trades registers for resident warps.
// ---------------------------------------------------------------------------
// 32-cadence groups per sweep step: with the data moments the sweep is sparse (only steps that
// can contain an in-transit cadence are visited), and a fine step wastes the least work around
// each transit; without them every cadence is read (t, y, ivar) and 4 loads in flight per array
// hide the latency.
constexpr int sweep_unroll(bool moments) { return (moments || kWarps > 4) ? 1 : 4; }
constexpr int kQueueMax = 32 * (4 + 1);  // 31 left over + kUnroll * 32 new entries

// CTA-level state of the indexed producer: one batch of kThreads transit epochs of one planet (thread i: epoch kcur + i),
// the cadence range of each (rlo, count) and the exclusive prefix of the counts -- entry e of the batch is cadence
// rlo[r] + e - roff[r] with r the last range whose offset is <= e (drains look it up; nothing is queued).
struct CtaIndexState {
  uint32_t rlo[kThreads], roff[kThreads];
  alignas(16) uint32_t wtot[kWarps];  // own 16 bytes: the compiler reads the four totals with one vector load
  long long kfirst[XO_MAX_PLANETS], kend[XO_MAX_PLANETS];  // per planet: first / last epoch that can touch this CTA's cadences
};

// Work partition of the launch (ma.weights != nullptr): the G CTAs of the grid are dealt to the chains in proportion to their
// expected number of in-transit cadences (weights[c][p], written by whoever derived the element blocks), every chain getting
// at least one, such that the heaviest CTA is as light as possible (weighted_partition).  Every CTA derives the same table
// from the same integers (block reductions + one block scan over the chains) and picks its own (chain, part) from it.  Without weights:
// chain = blockIdx.y, part = blockIdx.x of gridDim.x.
struct Partition {
  int64_t c;       // chain
  uint32_t part, nparts;
  int64_t row0;    // first row of this CTA in `partial` (one row per warp)
  bool idle;
};

constexpr int kPartK = 4;  // chains per thread the table is built for: C <= kPartK * kThreads (transit_weighted())
__device__ __forceinline__ Partition weighted_partition(const uint32_t *__restrict__ w, int C, int P, int2 *__restrict__ rows,
                                                        Partition *sh, unsigned long long *red) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t G = gridDim.x * gridDim.y, b = blockIdx.y * gridDim.x + blockIdx.x;
  // thread i owns chains i, i + kThreads, ... (their weights stay in registers for the whole table computation)
  uint32_t wv[kPartK];
  unsigned long long mine = 0ull;
#pragma unroll
  for (int j = 0; j < kPartK; ++j) {
    const int c = tid + j * kThreads;
    uint32_t wc = 0u;
    if (c < C)
      for (int p = 0; p < P; ++p) wc += __ldg(w + (size_t)c * P + p);
    wv[j] = wc;  // 0: no such chain
    mine += wc;
  }
  // total weight (integers: the same in every CTA whatever the order of the additions)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(kFull, mine, o);
  if (lane == 0) red[warp] = mine;
  if (tid == 0) sh->idle = true;
  __syncthreads();
  unsigned long long total = 0ull;
  for (int x = 0; x < kWarps; ++x) total += red[x];
  // CTAs per chain: n_c = ceil(w_c r) with the largest r for which they add up to at most G, i.e. the apportionment with
  // the smallest heaviest CTA (max w_c / n_c).  r = (G - C) / total is always feasible (sum of ceilings <= G - C + C),
  // r = G / total in general is not: five bisection steps in between, each one block-wide integer sum -- every CTA of the
  // launch walks the same steps on the same integers and ends up with the same table.
  double r_lo = total ? (double)(G - (uint32_t)C) / (double)total * (1.0 - 0x1p-40) : 0.0;
  double r_hi = total ? (double)G / (double)total : 0.0;
  auto ctas_of = [&](uint32_t wc, double r) {  // ceil(wc r), at least 1; 0 for an empty slot
    const double x = (double)wc * r;
    const uint32_t f = (uint32_t)x;
    return wc ? max(1u, f + ((double)f < x ? 1u : 0u)) : 0u;
  };
  uint32_t *red32 = reinterpret_cast<uint32_t *>(red);  // [2][kWarps], alternating: one barrier per step
  __syncthreads();                                      // red[] (64-bit view) has been read
#pragma unroll 1
  for (int it = 0; it < 5; ++it) {
    const double r = 0.5 * (r_lo + r_hi);
    uint32_t cnt = 0u;
#pragma unroll
    for (int j = 0; j < kPartK; ++j) cnt += ctas_of(wv[j], r);
    cnt = __reduce_add_sync(kFull, cnt);
    uint32_t *buf = red32 + (it & 1) * kWarps;
    if (lane == 0) buf[warp] = cnt;
    __syncthreads();
    uint32_t all = 0u;
    for (int x = 0; x < kWarps; ++x) all += buf[x];
    if (all <= G) r_lo = r; else r_hi = r;
  }
  // chain order of the table: c = i + j kThreads sorted by (j, i), i.e. ascending c
  uint32_t nj[kPartK];
#pragma unroll
  for (int j = 0; j < kPartK; ++j) nj[j] = ctas_of(wv[j], r_lo);
  uint32_t cum_prev = 0u;  // CTAs of all chains of the slices before j
#pragma unroll
  for (int j = 0; j < kPartK; ++j) {
    if (j * kThreads >= C) break;
    uint32_t inc = nj[j];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(kFull, inc, d);
      if (lane >= d) inc += v;
    }
    __syncthreads();  // the previous use of the buffers is over
    if (lane == 31) red32[warp] = inc;
    __syncthreads();
    uint32_t cum = cum_prev + inc - nj[j], slice = 0u;
    for (int x = 0; x < kWarps; ++x) {
      if (x < warp) cum += red32[x];
      slice += red32[x];
    }
    const int c = tid + j * kThreads;
    if (c < C && b >= cum && b < cum + nj[j]) {
      sh->c = c;
      sh->part = b - cum;
      sh->nparts = nj[j];
      sh->row0 = (int64_t)b * kWarps;  // the chain's rows are those of its CTAs cum .. cum + n - 1
      sh->idle = false;
      if (rows && b == cum) rows[c] = make_int2((int)(cum * kWarps), (int)(nj[j] * kWarps));
    }
    cum_prev += slice;
  }
  __syncthreads();
  return *sh;
}

// kVjp = false: Gaussian log-likelihood (y, ivar, mean) and its gradient.
// kVjp = true : reverse pass of the light curve alone: `y` is the cotangent bar_lc[C][N][P]
//               (ivar, mean unused), outputs are gel/gcl only.
// kMoments   : the out-of-transit part of the Gaussian sums comes from data-only moments
//              (moments[4] = y_ref, S1 = sum ivar, Sy = sum ivar (y - y_ref), Syy = sum ivar (y - y_ref)^2,
//              computed once per data set by transit_moments_kernel); the sweep then reads t only.
// kModel     : single-launch model evaluation: elements from parameters at CTA start (ma.par/star),
//              and the last CTA of each chain folds the partials and runs the reverse prologue.
// kDelay     : light-travel delay in the sample evaluation (element XO_EL_RSTAR > 0); sin i then carries a gradient
// kTtv       : transit-time tables (ma.ttv_*): compiled out of the ordinary kernels
template <int kMinBlocks, bool kVjp, bool kMoments, bool kModel, bool kDelay = false, bool kTtv = false>
__global__ void __launch_bounds__(kThreads, kMinBlocks) transit_logp_grad_kernel(
    const double *__restrict__ t, const double *__restrict__ y, const double *__restrict__ ivar, int64_t N,
    const double *__restrict__ elem, const double *__restrict__ cl, const double *__restrict__ mean, int P,
    const double *__restrict__ dtk, const double *__restrict__ wk, int K, int64_t L, int slots,
    const double *__restrict__ moments, double *__restrict__ partial, const xo::ModelArgs ma) {
  constexpr bool kBase = !kVjp && !kMoments;  // accumulate the per-cadence Gaussian terms in the sweep
  __shared__ ChainSmem sm;
  constexpr int kUnroll = sweep_unroll(kMoments);
  constexpr int kQueue = 32 * (kUnroll + 1);
  static_assert(kQueue <= kQueueMax, "queue");
  __shared__ uint32_t queue[kWarps][XO_MAX_PLANETS][kQueue];
  __shared__ double gacc[kWarps][XO_MAX_PLANETS][XO_NGRAD];
  __shared__ double lacc[XO_NGRAD + 3][kThreads];  // per-thread gradient accumulators of one drain
  __shared__ double pacc[XO_NGRAD][kThreads];      // per-thread cotangent-scaled sums of the current planet
  __shared__ CtaIndexState cix;
  __shared__ Partition part_sh;
  __shared__ unsigned long long part_red[kWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool weighted = !kModel && ma.weights != nullptr;
  // Everything up to pdl_wait() is independent of the kernel that precedes this one in the stream
  // (the model prologue writing elem / cl): when launched as its programmatic dependent this part
  // overlaps the prologue's tail -- accumulators zeroed, this warp's time stamps pulled into L2.
  xo::pdl_launch_dependents();
  for (int i = lane; i < XO_MAX_PLANETS * XO_NGRAD; i += 32) (&gacc[warp][0][0])[i] = 0.0;
#pragma unroll
  for (int j = 0; j < XO_NGRAD + 3; ++j) lacc[j][threadIdx.x] = 0.0;
#pragma unroll
  for (int j = 0; j < XO_NGRAD; ++j) pacc[j][threadIdx.x] = 0.0;
  // (data only, so legal before the wait) Pull the light curve and the time index into L2 while the preceding kernel is
  // still deriving the elements: every line of them is touched by some chain of the launch (ncu: 2.7 MB of DRAM reads per
  // launch = t, y, ivar and the table once), and what would be a chain of dependent cold misses at the start of each CTA
  // (window -> table entry -> time stamp -> sample) becomes L2 hits.  The launch's threads share the lines evenly.
  {
    const int64_t gthreads = (int64_t)gridDim.x * gridDim.y * kThreads;
    const int64_t gtid = ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * kThreads + threadIdx.x;
    const int64_t nl = (N * 8 + 127) / 128;  // 128-byte lines per array
    auto pull = [&](const double *base, int64_t lines) {
      for (int64_t i = gtid; i < lines; i += gthreads) xo::prefetch_l2(reinterpret_cast<const char *>(base) + i * 128);
    };
    pull(t, nl);
    if (!kVjp) {
      pull(y, nl);
      pull(ivar, nl);
    }
    if (kMoments) pull(moments, (XO_MOMENTS_LEN * 8 + ((int64_t)moments[5] + 1) * 4 + 127) / 128);
  }
  xo::pdl_wait();
  // which chain, and which part of its cadences, this CTA works on
  Partition pt;
  if (weighted) {
    pt = weighted_partition(ma.weights, ma.nchains, P, ma.rows, &part_sh, part_red);
    if (pt.idle) return;  // more CTAs in the grid than the table deals out
  } else {
    // heaviest chains first (ma.order, optional): CTAs are dispatched in blockIdx order, so the long
    // ones start early and the short ones fill the tail of the launch
    pt.c = ma.order ? ma.order[blockIdx.y] : blockIdx.y;
    pt.part = blockIdx.x;
    pt.nparts = gridDim.x;
    pt.row0 = pt.c * slots + (int64_t)blockIdx.x * kWarps;
    pt.idle = false;
  }
  const int64_t c = pt.c;
  // cadences per warp: the launch's L, or this chain's own split over its nparts CTAs
  const int64_t Lw = weighted ? ((N + (int64_t)pt.nparts * kWarps - 1) / ((int64_t)pt.nparts * kWarps) + 31) / 32 * 32 : L;
  const int slot = (int)pt.part * kWarps + warp;
  const uint32_t start = (uint32_t)min((int64_t)slot * Lw, N);
  const uint32_t end = (uint32_t)min((int64_t)slot * Lw + Lw, N);
  const uint32_t cstart = (uint32_t)min((int64_t)pt.part * kWarps * Lw, N);  // ... and of the whole CTA
  const uint32_t cend = (uint32_t)min(((int64_t)pt.part + 1) * kWarps * Lw, N);
  if (kModel && ma.stage_from_params)
    stage_chain_from_params(sm, ma, mean, dtk, wk, c, P, K);
  else
    stage_chain(sm, elem, cl, mean, dtk, wk, c, P, K);
  if (kTtv) stage_ttv(sm, ma.ttv_times, ma.ttv_count, ma.ttv_nt, c, P);
  double *const gl = &lacc[0][threadIdx.x];          // g[j]  at gl[j * kThreads]
  double *const gsl = &lacc[XO_NGRAD][threadIdx.x];  // gs[j] at gsl[j * kThreads]
  double *const pa = &pacc[0][threadIdx.x];          // sum over this lane's samples of scale * g[j]
  int pcur = -1;                                     // planet whose sums pa[] currently holds (warp-uniform)
  // Sum over the warp's 32 lanes of each of `rows` rows of a [rows][kThreads] shared array (this
  // warp's 32 columns): lane j < rows adds up row j in a fixed, lane-rotated order (rotation keeps
  // the 8-byte reads of different rows on different banks) -- ~1/5 of the instructions of one
  // shuffle tree per row, and just as deterministic.
  auto row_sum = [&](const double *rows_base) {  // rows_base = &arr[0][warp * 32]; returns row `lane`'s sum
    const double *r = rows_base + lane * kThreads;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int l = 0; l < 32; l += 4) {
      s0 += r[(l + lane) & 31];
      s1 += r[(l + 1 + lane) & 31];
      s2 += r[(l + 2 + lane) & 31];
      s3 += r[(l + 3 + lane) & 31];
    }
    return (s0 + s1) + (s2 + s3);
  };
  // lane sums -> gacc[warp][p]; runs when the drained planet changes and once at the end, not
  // once per drain
  auto flush = [&](int p) {
    __syncwarp();
    if (lane < XO_NGRAD && lane != 1 && (kDelay || lane != 7)) {  // d/dtref == d/dt0; sin i carries a gradient only through the light delay
      const double v = row_sum(&pacc[0][warp * 32]);
      gacc[warp][p][lane] += v;
      if (lane == 0) gacc[warp][p][1] += v;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < XO_NGRAD; ++j) pa[j * kThreads] = 0.0;
  };

  const int nacc = kNBase + P * XO_NGRAD;
  double lp = 0.0, gm = 0.0, gc0 = 0.0, gc1 = 0.0, gc2 = 0.0;
  int cnt = 0;  // lane p holds the fill level of queue p
  const double mean_c = sm.mean;
  const uint32_t lt_mask = (1u << lane) - 1u;

  bool indexed = false;  // this CTA enumerates transit epochs through the time index (set below, before the first drain)
  // entry e of the current batch of the indexed producer -> cadence (see CtaIndexState)
  auto batch_entry = [&](uint32_t e) {
    int lo = 0;
#pragma unroll
    for (int h = kThreads >> 1; h > 0; h >>= 1)
      if (cix.roff[lo + h] <= e) lo += h;
    return cix.rlo[lo] + (e - cix.roff[lo]);
  };
  // evaluate m (<= 32) entries of planet p: src[off .. off + m) of a queue, or (src == nullptr) entries off .. off + m of
  // the indexed producer's batch
  auto drain = [&](int p, const uint32_t *src, int off, int m) {
    if (p != pcur) {
      if (pcur >= 0) flush(pcur);
      pcur = p;
    }
    double g0s = 0.0;  // TTV: this lane's cotangent-scaled d lc / d (transit time)
    int jself = -1;    // ... and the transit it belongs to
    if (lane < m) {
      double scale;  // cotangent multiplying this lane's accumulated d lc / d elements
      const uint32_t idx = src ? src[off + lane] : batch_entry((uint32_t)(off + lane));
      const double tt = __ldg(t + idx);
      // issued now so that their latency hides behind the sample evaluation
      const double yy = kVjp ? (ma.vjp_shared ? -__ldg(y + (int64_t)c * N + idx) : __ldg(y + ((int64_t)c * N + idx) * P + p)) : __ldg(y + idx);
      const double iv = kVjp ? 0.0 : __ldg(ivar + idx);
      const double tx = (kTtv && ma.texp_vec) ? __ldg(ma.texp_vec + idx) : 1.0;  // per-cadence exposure time (unit stencil in dtk)
      double ltot = 0.0;
      bool owner = true;  // the lowest in-window planet owns the per-cadence terms
      // one call site for the sample evaluation: planet p with its gradient weights, any other
      // planet that is also in transit at this cadence with weight 0 (value only)
      for (int q = 0; q < P; ++q) {
        const bool self = (q == p);
        if (!self && kVjp) continue;
        // with a transit-time table the cadence is measured from its nearest transit (ttv.py:175-187)
        int jq = -1;
        const double ttq = kTtv ? tt - ttv_shift(ttv_view(ma.ttv_times, ma.ttv_count, ma.ttv_nt, c, P, q), sm.elem[q][XO_EL_T0], tt, &jq) : tt;
        // "is this cadence also in planet q's queue?" must be the producer's own predicate, or the per-cadence terms are
        // owned twice / by nobody: the stepping producer tests the (TTV: widened) phase window, the indexed one the exact
        // window around the nearest transit
        if (!self && !((kTtv && indexed) ? xo::in_window(sm.elem[q], sm.rperiod[q], ttq) : in_window_phase(sm, q, tt))) continue;
        if (q < p) owner = false;
        if (self) jself = jq;
        for (int k = 0; k < K; ++k) {
          const double w = sm.wk[k];
          ltot = fma(w, xo::lc_point<true, kThreads, kDelay, DrainElemPtr>(sm.elem[q], sm.cl, kTtv ? ttq + __dmul_rn(tx, sm.dtk[k]) : ttq + sm.dtk[k],
                                                             self ? w : 0.0, gl, gsl), ltot);
        }
      }
      if (kVjp) {
        scale = yy;  // the cotangent bar_lc[c][idx][p]
      } else {
        const double r0 = yy - mean_c;
        scale = iv * (r0 - ltot);
        if (owner) {  // -1/2 iv (res^2 - r0^2) and iv (res - r0), without cancellation
          lp = fma(0.5 * iv * ltot, 2.0 * r0 - ltot, lp);
          gm = fma(-iv, ltot, gm);
        }
      }
      g0s = scale * gl[0];
      gc0 = fma(scale, gsl[0], gc0);
      gc1 = fma(scale, gsl[kThreads], gc1);
      gc2 = fma(scale, gsl[2 * kThreads], gc2);
      gsl[0] = gsl[kThreads] = gsl[2 * kThreads] = 0.0;
#pragma unroll
      for (int j = 0; j < XO_NGRAD; ++j) {
        if (j == 1 || (j == 7 && !kDelay)) continue;
        pa[j * kThreads] = fma(scale, gl[j * kThreads], pa[j * kThreads]);
        gl[j * kThreads] = 0.0;
      }
      gl[kThreads] = 0.0;
    }
    if (kTtv && ma.ttv_grad) {
      // d / d (transit time): the entries of a drain are in cadence order, so the lanes of one transit are adjacent --
      // segmented sum in lane order, one atomic per transit and drain
      double v = g0s;
      const int key = jself;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double o = __shfl_up_sync(kFull, v, d);
        const int ko = __shfl_up_sync(kFull, key, d);
        if (lane >= d && ko == key) v += o;
      }
      const int kn = __shfl_down_sync(kFull, key, 1);
      if (key >= 0 && (lane == 31 || kn != key)) atomicAdd(ma.ttv_grad + ((size_t)c * P + p) * ma.ttv_nt + key, v);
    }
  };

  // The sweep is a producer / consumer loop written as ONE loop, so that the expensive drain body exists once in
  // the binary.  Two producers:
  //  * indexed (data moments + time index, sorted t): no time stamp is read and nothing is queued.  The CTA works as one:
  //    per planet, the transit epochs that can touch the CTA's cadence range are taken kThreads at a time (one per thread:
  //    window bounds -> cadence range by table look-up), a block scan turns the range lengths into offsets, and the
  //    warps drain the batch's entries 32 at a time, round robin -- every drain but the batch's last is full.
  //  * stepping (per warp): steps of 32*kUnroll cadences; for sorted t with moments, steps are taken 32 at a time (lane l:
  //    can step l contain an in-transit cadence of any planet?  one ballot gives the group's hit mask, only set
  //    bits are visited); otherwise every cadence is read and the out-of-transit Gaussian terms accumulated.  In-window
  //    cadences go to per-warp, per-planet queues that are drained whenever they hold >= 32 entries; a final pass
  //    (last = true) pools what the CTA's warps are left with.
  const bool sorted = kMoments && moments[4] != 0.0;  // t non-decreasing: steps can be skipped wholesale
  const int nbins = kMoments ? (int)moments[5] : 0;   // > 0: the block carries the time index
  const double *tp = t + lane;
  constexpr uint32_t kStep = 32 * kUnroll;
  uint32_t gnext = start, gcur = start, hitmask = 0u;
  __shared__ int wfill[kWarps][XO_MAX_PLANETS];  // final pass: what each warp is left with, per planet
  // ... pooled over the CTA: behind warp 0's (by then < 32) left-overs when the queue is long enough, else in its own array
  constexpr bool kPoolInQueue = kQueue >= 32 + kWarps * 32;
  __shared__ uint32_t pool_own[kPoolInQueue ? 1 : XO_MAX_PLANETS][kPoolInQueue ? 1 : kWarps * 32];
  auto pool = [&](int p) -> uint32_t * { return kPoolInQueue ? &queue[0][p][32] : pool_own[kPoolInQueue ? 0 : p]; };
  // epoch range of every planet over this CTA's cadences (thread p: planet p).  A planet with more epochs than cadences in
  // the range (or a non-finite / non-positive period) cannot be enumerated: the CTA then takes the stepping producer.
  __shared__ int ix_p;                  // indexed producer: current planet, next / last epoch of it
  __shared__ long long ix_kcur, ix_klast;
  if (kMoments && nbins > 0) {
    bool dense = false;
    if (threadIdx.x < P) {
      long long ka = 1, kb = 0;
      if (cstart < cend) {
        const int q = threadIdx.x;
        const double per = sm.elem[q][XO_EL_PERIOD], t0p = sm.elem[q][XO_EL_T0];
        const double tA = __ldg(t + cstart), tB = __ldg(t + cend - 1);
        const double dm = kTtv ? sm.ttv_dmax[q] : 0.0;  // 0 without a transit-time table
        const double fa = floor((tA - t0p - dm) * sm.rperiod[q] - 0.5) - 1.0, fb = ceil((tB - t0p + dm) * sm.rperiod[q] + 0.5) + 1.0;
        dense = !(per > 0.0) || !(fb - fa <= (double)(cend - cstart) + 4.0);
        if (!dense) {
          ka = (long long)fa;
          kb = (long long)fb;
        }
      }
      cix.kfirst[threadIdx.x] = ka;
      cix.kend[threadIdx.x] = kb;
    }
    if (threadIdx.x == 0) {
      ix_p = -1;
      ix_kcur = 1;
      ix_klast = 0;
    }
    // CTA-uniform: the drain's "is this cadence in planet q's queue too" predicate depends on the producer
    indexed = __syncthreads_and(!dense) != 0;
  }
#pragma unroll 1
  for (;;) {
    bool last = false;
    int pb = 0, pe = P;  // planets the consumer looks at this iteration
    uint32_t btot = 0u;  // indexed producer: entries in the current batch
    if (kMoments && indexed) {
      // ---------------- indexed producer (the whole CTA in step) ----------------
      int p = ix_p;
      long long kcur = ix_kcur, klast = ix_klast;
      while (kcur > klast) {  // this planet is exhausted: on to the next one
        if (++p >= P) break;
        kcur = cix.kfirst[p];
        klast = cix.kend[p];
      }
      if (p >= P) {
        last = true;
      } else {
        uint32_t lo = 0u, rc = 0u;
        const long long k = kcur + threadIdx.x;
        if (k <= klast) {
          const TimeIndex ix{reinterpret_cast<const uint32_t *>(moments + XO_MOMENTS_LEN), nbins, __ldg(t), moments[6]};
          const double per = sm.elem[p][XO_EL_PERIOD], t0p = sm.elem[p][XO_EL_T0], hp = 0.5 * per;
          const double ts = sm.elem[p][XO_EL_TSTART], te = sm.elem[p][XO_EL_TEND];
          const TtvView tv = kTtv ? ttv_view(ma.ttv_times, ma.ttv_count, ma.ttv_nt, c, P, p) : TtvView{nullptr, 0};
          uint32_t i0, i1;
          if (kTtv && tv.n > 0) {
            // epoch k of a transit-time table: the cadences nearest to transit k (between the mid-points to its
            // neighbours: (m_{k-1}, m_k], ttv.py:159-178) that fall inside the window around it
            const double ck = ttv_centre(tv.tt, tv.n, per, k);
            const double mlo = 0.5 * (ttv_centre(tv.tt, tv.n, per, k - 1) + ck), mhi = 0.5 * (ck + ttv_centre(tv.tt, tv.n, per, k + 1));
            i0 = (ck + ts > mlo) ? index_bound<false>(t, (uint32_t)N, ix, ck + ts) : index_bound<true>(t, (uint32_t)N, ix, mlo);
            i1 = (ck + te < mhi) ? index_bound<true>(t, (uint32_t)N, ix, ck + te) : index_bound<true>(t, (uint32_t)N, ix, mhi);
          } else {
            const double tc = fma((double)k, per, t0p);
            // the window clipped to the epoch's own period; the clip boundary B_k = t0 + (k + 1/2) P is computed by the
            // same expression on both sides, so consecutive epochs neither overlap nor leave a gap
            i0 = (ts > -hp) ? index_bound<false>(t, (uint32_t)N, ix, tc + ts)
                            : index_bound<false>(t, (uint32_t)N, ix, fma((double)k - 0.5, per, t0p));
            i1 = (te < hp) ? index_bound<true>(t, (uint32_t)N, ix, tc + te)
                           : index_bound<false>(t, (uint32_t)N, ix, fma((double)k + 0.5, per, t0p));
          }
          lo = max(i0, cstart);
          const uint32_t hi = min(i1, cend);
          rc = hi > lo ? hi - lo : 0u;
        }
        uint32_t inc = rc;  // inclusive scan of the range lengths over the warp ...
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t v = __shfl_up_sync(kFull, inc, d);
          if (lane >= d) inc += v;
        }
        __syncthreads();  // every warp is done with the previous batch (and has read ix_*)
        if (lane == 31) cix.wtot[warp] = inc;
        if (threadIdx.x == 0) {
          ix_p = p;
          ix_kcur = kcur + kThreads;
          ix_klast = klast;
        }
        __syncthreads();
        uint32_t base = 0u;  // ... and over the CTA
        for (int w = 0; w < kWarps; ++w) {
          if (w < warp) base += cix.wtot[w];
          btot += cix.wtot[w];
        }
        cix.rlo[threadIdx.x] = lo;
        cix.roff[threadIdx.x] = base + inc - rc;
        __syncthreads();
        pb = p;
        pe = p + 1;
      }
    } else {
      // ---------------- stepping producer ----------------
      uint32_t base = 0u;
      bool have = true;
      if (hitmask == 0u) {
        if (gnext >= end) {
          last = true;
        } else {
          const uint32_t sb = gnext + lane * kStep;
          bool hit = sb < end;
          if (sorted && hit) {
            const double tA = __ldg(t + sb), tB = __ldg(t + min(sb + kStep, end) - 1);
            hit = false;
            for (int p = 0; p < P; ++p) {
              const double rpp = sm.rperiod[p];
              const double uA = (tA - sm.elem[p][XO_EL_T0]) * rpp;
              const double a = uA - rint(uA), b = a + (tB - tA) * rpp;  // phase interval, a in [-1/2, 1/2]
              const double lo = sm.wlo[p] - 1e-9, hi = sm.whi[p] + 1e-9;
              hit |= (b - a >= 1.0) | ((a <= hi) & (b >= lo)) | ((a <= hi + 1.0) & (b >= lo + 1.0)) | ((a <= hi - 1.0) & (b >= lo - 1.0));
            }
          }
          hitmask = __ballot_sync(kFull, hit);
          gcur = gnext;
          gnext = (end - gnext > 32u * kStep) ? gnext + 32u * kStep : end;
          have = hitmask != 0u;
        }
      }
      if (!last && have) {
        double tt[kUnroll], yy[kUnroll], iv[kUnroll];
        uint32_t valid = 0u;  // bit u: cadence base + 32 u + lane exists
        base = gcur + (uint32_t)(__ffs(hitmask) - 1) * kStep;
        hitmask &= hitmask - 1u;
        if (kUnroll == 1 && hitmask != 0u && (lane & 15) == 0) {  // the next hit step's two lines on their way to L1
          const uint32_t bn = gcur + (uint32_t)(__ffs(hitmask) - 1) * kStep + lane;
          if (bn < end) xo::prefetch_l1(t + bn);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {  // kUnroll coalesced loads in flight together
          const bool v = base + 32 * u + lane < end;
          valid |= v ? (1u << u) : 0u;
          tt[u] = v ? __ldg(tp + base + 32 * u) : 0.0;
        }
        if (kBase) {
#pragma unroll
          for (int u = 0; u < kUnroll; ++u) {
            const uint32_t i = base + 32 * u + lane;
            yy[u] = (valid >> u) & 1u ? __ldg(y + i) : 0.0;
            iv[u] = (valid >> u) & 1u ? __ldg(ivar + i) : 0.0;
          }
#pragma unroll
          for (int u = 0; u < kUnroll; ++u) {
            const double r0 = yy[u] - mean_c;
            const double wr = iv[u] * r0;
            lp = fma(-0.5 * wr, r0, lp);
            gm += wr;
          }
        }
        for (int p = 0; p < P; ++p) {
          // window constants of planet p in registers for the whole step
          const double t0p = sm.elem[p][XO_EL_T0], rpp = sm.rperiod[p], lo = sm.wlo[p], hi = sm.whi[p];
          uint32_t inw = 0u;
#pragma unroll
          for (int u = 0; u < kUnroll; ++u) {
            const double ph0 = (tt[u] - t0p) * rpp;
            const double ph = ph0 - rint(ph0);
            inw |= ((ph >= lo) & (ph <= hi)) ? (1u << u) : 0u;
          }
          inw &= valid;
          if (!__any_sync(kFull, inw != 0u)) continue;  // nothing of this planet in this step
          int fill = __shfl_sync(kFull, cnt, p);
#pragma unroll
          for (int u = 0; u < kUnroll; ++u) {
            const bool in = (inw >> u) & 1u;
            const uint32_t mask = __ballot_sync(kFull, in);
            if (in) queue[warp][p][fill + __popc(mask & lt_mask)] = base + 32 * u + lane;
            fill += __popc(mask);
          }
          if (lane == p) cnt = fill;
        }
        __syncwarp();
      } else if (!last) {
        continue;  // a group of 32 steps without a hit
      }
    }
    // ---------------- consumer: the one drain call site ----------------
    // Indexed batches: entries 32 w, 32 (w + kWarps), ... of the batch go to warp w.  Stepping producer, ordinary
    // iterations: a queue is drained 32 entries at a time as soon as it holds that many; its final pass: the < 32
    // entries each warp is left with are pooled over the CTA's warps (warp order, so still deterministic) and drained
    // in chunks of 32 dealt round-robin to the warps -- up to kWarps partial drains become total/32 full ones.
    const bool coop = kMoments && indexed;
    if (coop && last) break;  // batches leave nothing behind
    if (last) {
      if (lane < P) wfill[warp][lane] = cnt;
      __syncthreads();
      for (int p = 0; p < P; ++p) {
        int before = 0;
        for (int w = 0; w < warp; ++w) before += wfill[w][p];
        const int mine = wfill[warp][p];
        if (lane < mine) pool(p)[before + lane] = queue[warp][p][lane];
      }
      __syncthreads();
    }
    for (int p = pb; p < pe; ++p) {
      int fill, off, stride;
      const uint32_t *src;
      if (coop) {
        fill = (int)btot;
        off = 32 * warp;
        stride = 32 * kWarps;
        src = nullptr;
      } else if (last) {
        fill = 0;
        for (int w = 0; w < kWarps; ++w) fill += wfill[w][p];
        off = 32 * warp;
        stride = 32 * kWarps;
        src = pool(p);
      } else {
        fill = __shfl_sync(kFull, cnt, p);
        if (fill < 32) continue;
        off = 0;
        stride = 32;
        src = queue[warp][p];
      }
#pragma unroll 1
      while ((coop || last) ? off < fill : fill - off >= 32) {
        drain(p, src, off, min(32, fill - off));
        off += stride;
      }
      if (!last && !coop) {  // move the < 32 left-over entries to the front
        const int rem = fill - off;
        const uint32_t v = lane < rem ? queue[warp][p][off + lane] : 0u;
        __syncwarp();
        if (lane < rem) queue[warp][p][lane] = v;
        __syncwarp();
        if (lane == p) cnt = rem;
      }
    }
    if (last) break;
  }
  if (pcur >= 0) flush(pcur);
  if (kMoments && slot == 0 && lane == 0) {  // out-of-transit Gaussian terms from the data moments
    const double m = mean_c - moments[0];
    lp += -0.5 * (moments[3] - 2.0 * m * moments[2] + m * m * moments[1]);
    gm += moments[2] - m * moments[1];
  }
  if (kVjp && ma.vjp_gsum && slot == 0 && lane == 0) gm = -ma.vjp_gsum[2 * c];  // d / d mean of a likelihood of the residuals
  // the five per-warp scalars through the same row sums (lacc's rows are free by now)
  gl[0] = lp;
  gl[kThreads] = gm;
  gl[2 * kThreads] = gc0;
  gl[3 * kThreads] = gc1;
  gl[4 * kThreads] = gc2;
  __syncwarp();
  double *out = partial + (pt.row0 + warp) * nacc;
  if (lane < kNBase) out[lane] = row_sum(&lacc[0][warp * 32]);
  for (int i = lane; i < P * XO_NGRAD; i += 32) out[kNBase + i] = gacc[warp][i / XO_NGRAD][i % XO_NGRAD];
  if (kModel) {
    // last CTA of the chain to finish folds all slots and runs the reverse pass of the prologue
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ma.counter + c, 1u) == pt.nparts - 1) ? 1 : 0;
    __syncthreads();
    if (is_last && warp == 0) {
      __threadfence();
      xo::epilogue_chain<true>(ma.par + c * P * XO_NPAR, ma.star + c * XO_NSTAR, partial + (int64_t)c * slots * nacc, slots, P,
                               ma.lognorm, ma.logp + c, ma.gpar + c * P * XO_NPAR, ma.gstar + c * XO_NSTAR, ma.gmean + c);
      if (lane == 0) ma.counter[c] = 0u;  // leave the counters zeroed for the next launch
    }
  }
}

// data-only moments of the Gaussian likelihood in ONE pass over (t, y, ivar), many CTAs:
// out = (y_ref, S1, Sy, Syy, sorted) with y_ref a reference value inside the data's range (the
// mean of 32 strided samples, the same in every CTA) so that the shifted sums
// Sy = sum ivar (y - y_ref), Syy = sum ivar (y - y_ref)^2 carry no cancellation, and sorted = 1
// when t is non-decreasing (lets the sweep skip whole 128-cadence steps with one interval test).
// Per-CTA partials go to out[8 + 4 b]; the last CTA to arrive (ticket in out[7]) folds them in
// CTA order, so the result does not depend on the arrival order (deterministic).
constexpr int kMomThreads = 256;
constexpr int kMomMaxCtas = (XO_MOMENTS_LEN - 8) / 4;
__global__ void __launch_bounds__(kMomThreads) transit_moments_kernel(const double *__restrict__ t, const double *__restrict__ y,
                                                                      const double *__restrict__ ivar, int64_t N,
                                                                      double *__restrict__ out, int nb) {
  __shared__ double red[4][kMomThreads / 32];
  __shared__ double yref_s;
  __shared__ int is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    const double v = warp_sum(__ldg(y + (int64_t)lane * N / 32));
    if (lane == 0) yref_s = v * (1.0 / 32.0);
  }
  __syncthreads();
  const double yref = yref_s;
  double s1 = 0.0, sy = 0.0, syy = 0.0, bad = 0.0;
  // nb > 0: the time index (see TimeIndex below) is built in the same pass -- table[j] = first cadence with t >= edge_j =
  // tmin + j w: the thread holding (t_i, t_{i+1}) owns the edges in (t_i, t_{i+1}] (-> i + 1), thread 0 also those <= t_0
  // (-> 0), the last one those > t_{N-1} (-> N).  Meaningless when t is not sorted; the fold below then marks it unusable.
  const double tmin = __ldg(t), bw = nb > 0 ? (__ldg(t + N - 1) - tmin) / (double)nb : 0.0;
  const bool build = nb > 0 && bw > 0.0 && isfinite(bw);
  const double inv_bw = build ? 1.0 / bw : 0.0;
  uint32_t *table = reinterpret_cast<uint32_t *>(out + XO_MOMENTS_LEN);
  auto edges_le = [&](double x) {  // number of edges j in [0, nb] with edge_j <= x (edge_j as index_bound computes it)
    const double u = (x - tmin) * inv_bw;
    int e = !(u >= 0.0) ? 0 : (u >= (double)nb ? nb + 1 : (int)u + 1);
    while (e <= nb && fma((double)e, bw, tmin) <= x) ++e;
    while (e > 0 && fma((double)(e - 1), bw, tmin) > x) --e;
    return e;
  };
  const int64_t stride = (int64_t)gridDim.x * kMomThreads;
  for (int64_t i = (int64_t)blockIdx.x * kMomThreads + threadIdx.x; i < N; i += stride) {
    const double w = __ldg(ivar + i), d = __ldg(y + i) - yref;
    s1 += w;
    sy = fma(w, d, sy);
    syy = fma(w * d, d, syy);
    const double ti = __ldg(t + i), tn = i + 1 < N ? __ldg(t + i + 1) : ti;
    if (!(tn >= ti)) bad += 1.0;
    if (build && tn >= ti) {
      const int e0 = min(edges_le(ti), nb);  // the last edge (= t_{N-1} up to rounding) always maps to N, as in transit_index_kernel
      if (i == 0)
        for (int j = 0; j < e0; ++j) table[j] = 0u;
      if (i + 1 < N) {
        const int e1 = min(edges_le(tn), nb);
        for (int j = e0; j < e1; ++j) table[j] = (uint32_t)(i + 1);
      } else {
        for (int j = e0; j <= nb; ++j) table[j] = (uint32_t)N;
      }
    }
  }
  double v4[4] = {s1, sy, syy, bad};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double v = warp_sum(v4[j]);
    if (lane == 0) red[j][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double a = 0.0;
    for (int w = 0; w < kMomThreads / 32; ++w) a += red[threadIdx.x][w];
    out[8 + 4 * blockIdx.x + threadIdx.x] = a;
  }
  __threadfence();
  __syncthreads();
  unsigned *ticket = reinterpret_cast<unsigned *>(out + 7);
  if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last || warp != 0) return;
  __threadfence();
  double f4[4] = {0.0, 0.0, 0.0, 0.0};
  for (int b = lane; b < (int)gridDim.x; b += 32) {
#pragma unroll
    for (int j = 0; j < 4; ++j) f4[j] += __ldcg(out + 8 + 4 * b + j);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) f4[j] = warp_sum(f4[j]);
  if (lane == 0) {
    out[0] = yref;
    out[1] = f4[0];
    out[2] = f4[1];
    out[3] = f4[2];
    out[4] = f4[3] == 0.0 ? 1.0 : 0.0;
    out[5] = (build && f4[3] == 0.0) ? (double)nb : 0.0;  // the sweep only trusts the table of a sorted series
    out[6] = (build && f4[3] == 0.0) ? inv_bw : 0.0;
    *ticket = 0u;
    reinterpret_cast<unsigned *>(out + 7)[1] = 0u;
  }
}

// rank chains by expected in-transit work, sum_p (window width / period), heaviest first;
// order[rank] = chain.  O(C^2) comparisons, C <= 65535 and typically <= 1024.
__global__ void chain_order_kernel(const double *__restrict__ elem, int64_t C, int P, int *__restrict__ order) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  auto weight = [&](int64_t k) {
    double w = 0.0;
    for (int p = 0; p < P; ++p) {
      const double *e = elem + (k * P + p) * XO_NEL;
      const double f = (e[XO_EL_TEND] - e[XO_EL_TSTART]) / e[XO_EL_PERIOD];
      w += (f == f && f < 1.0) ? f : 1.0;  // infinite window = every cadence
    }
    return w;
  };
  const double wc = weight(c);
  int rank = 0;
  for (int64_t k = 0; k < C; ++k) {
    const double wk = weight(k);
    rank += (wk > wc) || (wk == wc && k < c);
  }
  order[rank] = (int)c;
}

// fold per-warp partials in slot order: out[c][j] = sum_s partial[c][s][j]
__global__ void fold_partials_kernel(const double *__restrict__ partial, int slots, int nacc, int64_t C, int P,
                                     double lognorm, double *__restrict__ logp, double *__restrict__ gel,
                                     double *__restrict__ gcl, double *__restrict__ gmean) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * nacc) return;
  const int64_t c = idx / nacc;
  const int j = (int)(idx % nacc);
  const double *src = partial + c * slots * nacc + j;
  double acc = 0.0;
  for (int s = 0; s < slots; ++s) acc += src[(int64_t)s * nacc];
  if (j == 0)
    logp[c] = acc + lognorm;
  else if (j == 1)
    gmean[c] = acc;
  else if (j < kNBase)
    gcl[3 * c + (j - 2)] = acc;
  else
    gel[c * P * XO_NGRAD + (j - kNBase)] = acc;
}

// ---------------------------------------------------------------------------
// orbit_eval: relative position / R_star and RV shape, thread per (c, n, p)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) orbit_eval_kernel(const double *__restrict__ t, int64_t N,
                                                         const double *__restrict__ elem, int P,
                                                         double *__restrict__ X, double *__restrict__ Y,
                                                         double *__restrict__ Z, double *__restrict__ rv) {
  __shared__ double el[XO_MAX_PLANETS][XO_NEL];
  const int64_t c = blockIdx.y;
  for (int i = threadIdx.x; i < P * XO_NEL; i += blockDim.x) el[i / XO_NEL][i % XO_NEL] = elem[c * P * XO_NEL + i];
  __syncthreads();
  const int64_t total = N * P;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += stride) {
    const int64_t i = j / P;
    const int p = (int)(j % P);
    const double *e = el[p];
    double sf, cf, tm;
    xo::true_anomaly(e, __ldg(t + i), sf, cf, tm);
    const double ecc = e[XO_EL_ECC], cw = e[XO_EL_COSW], sw = e[XO_EL_SINW];
    const double rr = e[XO_EL_AOR] * (1.0 - ecc) * (1.0 + ecc) / fma(ecc, cf, 1.0);
    const double cwf = cw * cf - sw * sf, swf = sw * cf + cw * sf;
    const int64_t o = c * total + j;
    if (X) __stcs(X + o, -rr * cwf);
    if (Y) __stcs(Y + o, -rr * e[XO_EL_COSI] * swf);
    if (Z) __stcs(Z + o, rr * e[XO_EL_SINI] * swf);
    if (rv) __stcs(rv + o, e[XO_EL_KRV] * fma(ecc, cw, cwf));
  }
}

// ---------------------------------------------------------------------------
// RV logp + grad.  blockDim = (64, P): thread (x, p) owns planet p of cadence x of each
// 64-cadence tile, so the Kepler solution stays in registers between the forward sum
// over planets (through shared memory) and the gradient accumulation.
// ---------------------------------------------------------------------------
constexpr int kRvTile = 64;
// kCap: cap registers at 64 (two 512-thread CTAs, i.e. three 320-thread CTAs at P = 5, per SM)
template <bool kCap>
__global__ void __launch_bounds__(kRvTile *XO_MAX_PLANETS, kCap ? 2 : 1) rv_logp_grad_kernel(const double *__restrict__ t, const double *__restrict__ y,
                                    const double *__restrict__ ivar, int64_t N, const double *__restrict__ elem,
                                    const double *__restrict__ gamma, int P, int64_t L, int slots,
                                    double *__restrict__ partial) {
  __shared__ double el[XO_MAX_PLANETS][XO_NEL];
  __shared__ double contrib[2][XO_MAX_PLANETS][kRvTile];
  __shared__ double red[XO_MAX_PLANETS][kRvTile / 32][XO_NGRAD_RV + 2];
  const int64_t c = blockIdx.y;
  const int x = threadIdx.x, p = threadIdx.y;
  const int tid = p * kRvTile + x;
  for (int i = tid; i < P * XO_NEL; i += kRvTile * P) el[i / XO_NEL][i % XO_NEL] = elem[c * P * XO_NEL + i];
  __syncthreads();
  const double *e = el[p];
  const double nn = e[XO_EL_N], ecc = e[XO_EL_ECC], cw = e[XO_EL_COSW], sw = e[XO_EL_SINW], Kp = e[XO_EL_KRV];
  const double gam = gamma[c];
  double acc[XO_NGRAD_RV];
#pragma unroll
  for (int j = 0; j < XO_NGRAD_RV; ++j) acc[j] = 0.0;
  double lp = 0.0, gg = 0.0;
  const int64_t start = (int64_t)blockIdx.x * L, end = min(start + L, N);
  int buf = 0;
  for (int64_t base = start; base < end; base += kRvTile, buf ^= 1) {
    const int64_t i = base + x;
    const bool valid = i < end;
    double sf = 0.0, cf = 1.0, tm = 0.0, shape = 0.0, cwf = 0.0, swf = 0.0;
    if (valid) {
      xo::true_anomaly(e, __ldg(t + i), sf, cf, tm);
      cwf = cw * cf - sw * sf;
      swf = sw * cf + cw * sf;
      shape = fma(ecc, cw, cwf);
    }
    contrib[buf][p][x] = Kp * shape;
    __syncthreads();
    if (valid) {
      double model = gam;
      for (int q = 0; q < P; ++q) model += contrib[buf][q][x];
      const double res = __ldg(y + i) - model;
      const double wres = __ldg(ivar + i) * res;
      if (p == 0) {
        lp = fma(-0.5 * res, wres, lp);
        gg += wres;
      }
      double dfdM, dfde;
      xo::kepler_jac(sf, cf, ecc, dfdM, dfde);
      const double gM = -Kp * swf * dfdM * wres;
      acc[0] = fma(gM, -nn, acc[0]);
      acc[2] = fma(gM, tm, acc[2]);
      acc[3] = fma(wres * Kp, fma(-swf, dfde, cw), acc[3]);
      acc[4] = fma(wres * Kp, cf + ecc, acc[4]);
      acc[5] = fma(wres * Kp, -sf, acc[5]);
      acc[6] = fma(wres, shape, acc[6]);
    }
  }
  acc[1] = acc[0];
  // block reduce per planet row (2 warps per row), fixed order
  const int wrow = x >> 5, lane = x & 31;
#pragma unroll
  for (int j = 0; j < XO_NGRAD_RV; ++j) {
    double v = warp_sum(acc[j]);
    if (lane == 0) red[p][wrow][j] = v;
  }
  lp = warp_sum(lp);
  gg = warp_sum(gg);
  if (lane == 0) {
    red[p][wrow][XO_NGRAD_RV] = lp;
    red[p][wrow][XO_NGRAD_RV + 1] = gg;
  }
  __syncthreads();
  const int nacc = 2 + P * XO_NGRAD_RV;
  double *out = partial + ((int64_t)c * slots + blockIdx.x) * nacc;
  if (x < XO_NGRAD_RV) out[2 + p * XO_NGRAD_RV + x] = red[p][0][x] + red[p][1][x];
  if (p == 0 && x == 0) {
    out[0] = red[0][0][XO_NGRAD_RV] + red[0][1][XO_NGRAD_RV];
    out[1] = red[0][0][XO_NGRAD_RV + 1] + red[0][1][XO_NGRAD_RV + 1];
  }
}

__global__ void fold_rv_kernel(const double *__restrict__ partial, int slots, int nacc, int64_t C, int P, double lognorm,
                               double *__restrict__ logp, double *__restrict__ grv, double *__restrict__ ggamma) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * nacc) return;
  const int64_t c = idx / nacc;
  const int j = (int)(idx % nacc);
  const double *src = partial + c * slots * nacc + j;
  double acc = 0.0;
  for (int s = 0; s < slots; ++s) acc += src[(int64_t)s * nacc];
  if (j == 0)
    logp[c] = acc + lognorm;
  else if (j == 1)
    ggamma[c] = acc;
  else
    grv[c * P * XO_NGRAD_RV + (j - 2)] = acc;
}

inline Split split_rv(int64_t N, int64_t C) {
  const int sms = xo::sm_count();
  int64_t target = ((int64_t)sms * 6 + C - 1) / C;
  int64_t maxc = (N + 4 * kRvTile - 1) / (4 * kRvTile);
  int64_t s = target < 1 ? 1 : (target > maxc ? maxc : target);
  if (s < 1) s = 1;
  int64_t L = (N + s - 1) / s;
  L = ((L + kRvTile - 1) / kRvTile) * kRvTile;
  return Split{(int)s, L};
}

}  // namespace

namespace xo {

// tuning knobs (read once): XO_TRANSIT_MINBLOCKS in {2,3,4} picks the register cap of the
// logp kernel (254 / 168 / 128 registers), XO_TRANSIT_WARPS_PER_SM the launch width.
static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}
int transit_min_blocks() {
  static int v = env_int("XO_TRANSIT_MINBLOCKS", 4);  // 128 registers, 4 CTAs per SM (see DrainElemPtr); 3 = 168 registers
  return v;
}
static int transit_warps_per_sm() {
  static int v = env_int("XO_TRANSIT_WARPS_PER_SM", 20);  // re-measured in round 2 (profiles/r02_tuning_sweep.txt): 20 beats 24 by 2 %
  return v;
}

Split split_transit(int64_t N, int64_t C) {
  const int sms = sm_count();
  // aim at ~20 warps per SM over the whole launch, at least 256 cadences per warp
  int64_t target = ((int64_t)sms * transit_warps_per_sm() + C - 1) / C;
  // at least two CTAs per chain while the weighted partition can still deal them out unevenly: with one CTA per chain the
  // heaviest chain alone sets the length of the launch's tail (1024 chains: 3.2e6 -> 4.2e6 evals/s, profiles/r02_tuning_sweep.txt)
  static const int min_ctas = [] {
    const char *v = getenv("XO_TRANSIT_MIN_CTAS");
    return v && *v ? atoi(v) : 2;
  }();
  if (C > 1 && C <= kMaxWeighted && target < (int64_t)min_ctas * kWarps) target = (int64_t)min_ctas * kWarps;
  int64_t maxw = (N + 255) / 256;
  int64_t w = target < 1 ? 1 : (target > maxw ? maxw : target);
  if (w < 1) w = 1;
  w = ((w + kWarps - 1) / kWarps) * kWarps;  // whole CTAs
  int64_t L = (N + w - 1) / w;
  L = ((L + 31) / 32) * 32;
  return Split{(int)w, L};
}

// Weighted partition only while there are fewer chains than resident CTA slots: then every chain needs several CTAs to fill
// the device and how many each gets decides the balance (128 chains: 61.8 -> 56.7 us per step).  With more chains than slots
// the hardware's dynamic CTA dispatch evens the load out by itself and the table only costs (1024 chains: 4.2e6 evals/s with
// two CTAs per chain, 3.7e6 weighted; profiles/r02_tuning_sweep.txt).
static int64_t transit_wave() { return (int64_t)sm_count() * transit_min_blocks(); }  // resident CTAs (4 per SM at 128 registers)

bool transit_weighted(int64_t C, int64_t P) {
  static const int mode = env_int("XO_TRANSIT_WEIGHTED", -1);  // 0: never, 1: whenever the table fits, default: see above
  if (mode == 0 || C <= 1 || C * P > kMaxWeighted || C > kPartK * kThreads) return false;
  return mode == 1 || C <= transit_wave();
}

// CTAs of a weighted launch: whole waves of the device -- every CTA slot busy from the start, and (equal work per CTA)
// until the end; one wave while that leaves at least two CTAs per chain on average
int64_t transit_weighted_grid(int64_t N, int64_t C) {
  const Split sp = split_transit(N, C);
  const int64_t g0 = (int64_t)(sp.slots / kWarps) * C;  // what the uniform split would launch
  const int64_t wave = transit_wave();
  if (g0 <= wave) return g0;  // short series: the launch does not fill the device anyway
  static const int waves_env = env_int("XO_TRANSIT_WAVES", 0);
  int64_t waves = waves_env > 0 ? waves_env : (2 * C + wave - 1) / wave;
  if (waves * wave < C) waves = (C + wave - 1) / wave;
  return waves * wave;
}

int64_t transit_weighted_rows(int64_t N, int64_t C) { return transit_weighted_grid(N, C) * kWarps; }

int check_batch(int64_t N, int64_t C, int64_t P) {
  if (N < 0 || C < 0 || P < 0) return XO_EINVAL;
  if (P > XO_MAX_PLANETS) return XO_ERANGE;
  if (C > 65535) return XO_ERANGE;  // gridDim.y
  if (N >= ((int64_t)1 << 32)) return XO_ERANGE;
  return 0;
}

void launch_transit_logp_grad(const double *t, const double *y, const double *ivar, int64_t N, const double *elem,
                              const double *cl, const double *mean, int64_t C, int P, const double *dtk,
                              const double *wk, int K, Split sp, const double *moments, double *partial,
                              cudaStream_t st, const ModelArgs *model, bool pdl) {
  const ModelArgs ma = model ? *model : ModelArgs{};
  const dim3 grid = ma.weights ? dim3((unsigned)ma.grid_ctas) : dim3(sp.slots / kWarps, (unsigned)C);
#define XO_LAUNCH(MB, VJP, MOM, MODEL)                                                                                    \
  xo::launch_pdl(transit_logp_grad_kernel<MB, VJP, MOM, MODEL>, grid, dim3(kThreads), st, pdl, t, y, ivar, N, elem, cl, mean, P, \
                 dtk, wk, K, sp.L, sp.slots, moments, partial, ma)
#define XO_LAUNCH_TTV(VJP, MOM)                                                                                                    \
  xo::launch_pdl(transit_logp_grad_kernel<3, VJP, MOM, false, false, true>, grid, dim3(kThreads), st, pdl, t, y, ivar, N, elem, cl, mean, \
                 P, dtk, wk, K, sp.L, sp.slots, moments, partial, ma)
  if (ma.ttv_times || ma.texp_vec) {  // TTVOrbit, per-cadence exposure times: element-level launches only
    if (ivar == nullptr) XO_LAUNCH_TTV(true, false);
    else if (moments) XO_LAUNCH_TTV(false, true);
    else XO_LAUNCH_TTV(false, false);
    return;
  }
#undef XO_LAUNCH_TTV
  if (ivar == nullptr) {  // light-curve VJP mode: y = bar_lc[C][N][P] (or -ybar[C][N] with model->vjp_shared)
    if (ma.light_delay)
      xo::launch_pdl(transit_logp_grad_kernel<3, true, false, false, true>, grid, dim3(kThreads), st, pdl, t, y, ivar, N, elem, cl, mean, P,
                     dtk, wk, K, sp.L, sp.slots, moments, partial, ma);
    else
      XO_LAUNCH(3, true, false, false);
    return;
  }
  const int mb = transit_min_blocks();
  if (model && !model->elements_only) {
    if (moments) XO_LAUNCH(3, false, true, true);
    else XO_LAUNCH(3, false, false, true);
    return;
  }
  if (moments) {
    if (mb == 2) XO_LAUNCH(2, false, true, false);
    else if (mb == 4) XO_LAUNCH(4, false, true, false);
    else if (mb == 5) XO_LAUNCH(5, false, true, false);
    else XO_LAUNCH(3, false, true, false);
  } else {
    if (mb == 2) XO_LAUNCH(2, false, false, false);
    else if (mb == 4) XO_LAUNCH(4, false, false, false);
    else XO_LAUNCH(3, false, false, false);
  }
#undef XO_LAUNCH
}

void launch_transit_lc_fwd(const double *t, int64_t N, const double *elem, const double *cl, int64_t C, int P, const double *dtk,
                           const double *wk, int K, const double *ttv_times, const int *ttv_count, int ttv_nt, double *lc,
                           cudaStream_t st, bool light_delay, const double *texp_vec) {
  const int sms = sm_count();
  int64_t per_chain = (N + 255) / 256;
  int64_t want = ((int64_t)sms * 8 + C - 1) / C;
  int gx = (int)(per_chain < want ? per_chain : want);
  if (gx < 1) gx = 1;
  if (light_delay)
    xo::launch(transit_lc_fwd_kernel<true>, dim3(gx, (unsigned)C), dim3(256), 0, st, t, N, elem, cl, P, dtk, wk, K, ttv_times, ttv_count,
               ttv_nt, texp_vec, lc);
  else
    xo::launch(transit_lc_fwd_kernel<false>, dim3(gx, (unsigned)C), dim3(256), 0, st, t, N, elem, cl, P, dtk, wk, K, ttv_times, ttv_count,
               ttv_nt, texp_vec, lc);
}

void launch_transit_resid(const double *t, const double *y, int64_t N, const double *elem, const double *cl, const double *mean,
                          int64_t C, int P, const double *dtk, const double *wk, int K, double *resid, cudaStream_t st,
                          const double *ttv_times, const int *ttv_count, int ttv_nt) {
  const int sms = sm_count();
  int64_t per_chain = (N + 255) / 256;
  int64_t want = ((int64_t)sms * 8 + C - 1) / C;
  int gx = (int)(per_chain < want ? per_chain : want);
  if (gx < 1) gx = 1;
  xo::launch(transit_resid_kernel, dim3(gx, (unsigned)C), dim3(256), 0, st, t, y, N, elem, cl, mean, P, dtk, wk, K, ttv_times, ttv_count,
             ttv_nt, resid);
}

void launch_chain_order(const double *elem, int64_t C, int P, int *order, cudaStream_t st) {
  xo::launch(chain_order_kernel, (unsigned)((C + 127) / 128), 128, 0, st, elem, C, P, order);
}

}  // namespace xo

using xo::check_batch;
using xo::split_transit;

extern "C" {

int xo_transit_lc_fwd(const double *t, int64_t N, const double *elem, const double *cl, int64_t C, int64_t P,
                      const double *dtk, const double *wk, int K, double *lc, void *stream) {
  XO_NVTX_SCOPE("xo_transit_lc_fwd");
  if (int rc = check_batch(N, C, P)) return rc;
  if (K < 1 || K > XO_MAX_STENCIL) return XO_ERANGE;
  if (N == 0 || C == 0 || P == 0) return 0;
  if (!t || !elem || !cl || !dtk || !wk || !lc) return XO_EINVAL;
  xo::launch_transit_lc_fwd(t, N, elem, cl, C, (int)P, dtk, wk, K, nullptr, nullptr, 0, lc, static_cast<cudaStream_t>(stream));
  return xo::launch_status();
}

// ---- the same three entry points for a TTVOrbit (/root/reference/src/exoplanet/orbits/ttv.py:35-187) ----
int xo_transit_lc_fwd_ex(const double *t, int64_t N, const double *elem, const double *cl, int64_t C, int64_t P,
                         const double *dtk, const double *wk, int K, const double *ttv_times, const int *ttv_count, int ttv_nt,
                         int light_delay, const double *texp_vec, double *lc, void *stream) {
  XO_NVTX_SCOPE("xo_transit_lc_fwd_ex");
  if (int rc = check_batch(N, C, P)) return rc;
  if (K < 1 || K > XO_MAX_STENCIL) return XO_ERANGE;
  if (N == 0 || C == 0 || P == 0) return 0;
  if (!t || !elem || !cl || !dtk || !wk || !lc) return XO_EINVAL;
  if (ttv_times && (!ttv_count || ttv_nt < 1)) return XO_EINVAL;
  if (ttv_times && light_delay) return XO_EINVAL;  // the reference has no such combination either (ttv.py has no delayed path)
  xo::launch_transit_lc_fwd(t, N, elem, cl, C, (int)P, dtk, wk, K, ttv_times, ttv_count, ttv_nt, lc, static_cast<cudaStream_t>(stream),
                            light_delay != 0, texp_vec);
  return xo::launch_status();
}

size_t xo_transit_logp_grad_workspace(int64_t N, int64_t C, int64_t P) {
  if (N <= 0 || C <= 0 || P <= 0) return 0;
  Split sp = split_transit(N, C);
  return sizeof(double) * (size_t)C * (size_t)sp.slots * (size_t)(kNBase + P * XO_NGRAD);
}

// cadences per thread of the moments pass (XO_MOMENTS_PER_THREAD): fewer = more CTAs in flight for a 2.4 MB problem
static int64_t moments_per_thread() {
  static const int v = [] {
    const char *e = getenv("XO_MOMENTS_PER_THREAD");
    const int x = e && *e ? atoi(e) : 4;
    return x < 1 ? 1 : x;
  }();
  return v;
}

int xo_transit_data_moments(const double *t, const double *y, const double *ivar, int64_t N, double *moments,
                            void *stream) {
  XO_NVTX_SCOPE("xo_transit_data_moments");
  if (N <= 0 || !t || !y || !ivar || !moments) return XO_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(moments + 7, 0, sizeof(double), st);  // the arrival ticket
  if (e != cudaSuccess) return -(int)e;
  const int64_t want = (N + moments_per_thread() * kMomThreads - 1) / (moments_per_thread() * kMomThreads);
  const int ctas = (int)(want < 1 ? 1 : want > kMomMaxCtas ? kMomMaxCtas : want);
  xo::launch(transit_moments_kernel, ctas, kMomThreads, 0, st, t, y, ivar, N, moments, 0);
  return xo::launch_status();
}

// moments and time index in ONE pass over the data (what xo_transit_data_moments + xo_transit_data_index give in two)
int xo_transit_data_moments_index(const double *t, const double *y, const double *ivar, int64_t N, double *moments, size_t moments_len,
                                  void *stream) {
  XO_NVTX_SCOPE("xo_transit_data_moments_index");
  if (N <= 0 || !t || !y || !ivar || !moments) return XO_EINVAL;
  if (N >= ((int64_t)1 << 32)) return XO_ERANGE;
  if (moments_len < xo_transit_moments_len(N)) return XO_EWORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(moments + 7, 0, sizeof(double), st);  // the arrival ticket
  if (e != cudaSuccess) return -(int)e;
  const int64_t want = (N + moments_per_thread() * kMomThreads - 1) / (moments_per_thread() * kMomThreads);
  const int ctas = (int)(want < 1 ? 1 : want > kMomMaxCtas ? kMomMaxCtas : want);
  xo::launch(transit_moments_kernel, ctas, kMomThreads, 0, st, t, y, ivar, N, moments, index_bins(N));
  return xo::launch_status();
}

size_t xo_transit_moments_len(int64_t N) {
  if (N <= 0) return XO_MOMENTS_LEN;
  return (size_t)XO_MOMENTS_LEN + ((size_t)index_bins(N) + 2) / 2;  // the table holds nb + 1 32-bit entries
}

int xo_transit_data_index(const double *t, int64_t N, double *moments, size_t moments_len, void *stream) {
  XO_NVTX_SCOPE("xo_transit_data_index");
  if (N <= 0 || !t || !moments) return XO_EINVAL;
  if (N >= ((int64_t)1 << 32)) return XO_ERANGE;
  if (moments_len < xo_transit_moments_len(N)) return XO_EWORKSPACE;
  const int nb = index_bins(N);
  xo::launch(transit_index_kernel, (unsigned)((nb + 1 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream), t, N, moments, nb);
  return xo::launch_status();
}

int xo_transit_logp_grad(const double *t, const double *y, const double *ivar, int64_t N, const double *elem,
                         const double *cl, const double *mean, int64_t C, int64_t P, const double *dtk, const double *wk,
                         int K, double lognorm, const double *moments, double *logp, double *gel, double *gcl,
                         double *gmean, void *workspace, size_t workspace_bytes, void *stream) {
  XO_NVTX_SCOPE("xo_transit_logp_grad");
  if (int rc = check_batch(N, C, P)) return rc;
  if (K < 1 || K > XO_MAX_STENCIL) return XO_ERANGE;
  if (C == 0) return 0;
  if (N == 0 || P == 0) return XO_EINVAL;
  if (!t || !y || !ivar || !elem || !cl || !mean || !dtk || !wk || !logp || !gel || !gcl || !gmean || !workspace)
    return XO_EINVAL;
  if (workspace_bytes < xo_transit_logp_grad_workspace(N, C, P)) return XO_EWORKSPACE;
  Split sp = split_transit(N, C);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double *partial = static_cast<double *>(workspace);
  xo::launch_transit_logp_grad(t, y, ivar, N, elem, cl, mean, C, (int)P, dtk, wk, K, sp, moments, partial, st, nullptr);
  const int nacc = kNBase + (int)P * XO_NGRAD;
  const int64_t tot = C * nacc;
  xo::launch(fold_partials_kernel, (unsigned)((tot + 127) / 128), 128, 0, st, partial, sp.slots, nacc, C, (int)P, lognorm, logp,
                                                                      gel, gcl, gmean);
  return xo::launch_status();
}

int xo_transit_lc_vjp(const double *t, int64_t N, const double *elem, const double *cl, int64_t C, int64_t P,
                      const double *dtk, const double *wk, int K, const double *bar_lc, double *gel, double *gcl,
                      void *workspace, size_t workspace_bytes, void *stream) {
  XO_NVTX_SCOPE("xo_transit_lc_vjp");
  if (int rc = check_batch(N, C, P)) return rc;
  if (K < 1 || K > XO_MAX_STENCIL) return XO_ERANGE;
  if (C == 0) return 0;
  if (N == 0 || P == 0) return XO_EINVAL;
  if (!t || !elem || !cl || !dtk || !wk || !bar_lc || !gel || !gcl || !workspace) return XO_EINVAL;
  if (workspace_bytes < xo_transit_logp_grad_workspace(N, C, P) + sizeof(double) * 2 * (size_t)C) return XO_EWORKSPACE;
  Split sp = split_transit(N, C);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double *partial = static_cast<double *>(workspace);
  double *scratch = partial + (size_t)C * sp.slots * (kNBase + P * XO_NGRAD);  // logp / gmean of the fold (unused)
  xo::launch_transit_logp_grad(t, bar_lc, nullptr, N, elem, cl, nullptr, C, (int)P, dtk, wk, K, sp, nullptr, partial, st, nullptr);
  const int nacc = kNBase + (int)P * XO_NGRAD;
  const int64_t tot = C * nacc;
  xo::launch(fold_partials_kernel, (unsigned)((tot + 127) / 128), 128, 0, st, partial, sp.slots, nacc, C, (int)P, 0.0, scratch, gel,
                                                                      gcl, scratch + C);
  return xo::launch_status();
}

int xo_transit_logp_grad_ttv(const double *t, const double *y, const double *ivar, int64_t N, const double *elem,
                             const double *cl, const double *mean, int64_t C, int64_t P, const double *dtk, const double *wk,
                             int K, double lognorm, const double *moments, const double *ttv_times, const int *ttv_count,
                             int ttv_nt, double *logp, double *gel, double *gcl, double *gmean, double *gttv, void *workspace,
                             size_t workspace_bytes, void *stream) {
  XO_NVTX_SCOPE("xo_transit_logp_grad_ttv");
  if (int rc = check_batch(N, C, P)) return rc;
  if (K < 1 || K > XO_MAX_STENCIL) return XO_ERANGE;
  if (C == 0) return 0;
  if (N == 0 || P == 0) return XO_EINVAL;
  if (!t || !y || !ivar || !elem || !cl || !mean || !dtk || !wk || !logp || !gel || !gcl || !gmean || !gttv || !workspace ||
      !ttv_times || !ttv_count || ttv_nt < 1)
    return XO_EINVAL;
  if (workspace_bytes < xo_transit_logp_grad_workspace(N, C, P)) return XO_EWORKSPACE;
  Split sp = split_transit(N, C);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(gttv, 0, sizeof(double) * (size_t)C * P * ttv_nt, st);
  if (e != cudaSuccess) return -(int)e;
  double *partial = static_cast<double *>(workspace);
  xo::ModelArgs ma;
  ma.elements_only = 1;
  ma.ttv_times = ttv_times;
  ma.ttv_count = ttv_count;
  ma.ttv_nt = ttv_nt;
  ma.ttv_grad = gttv;
  xo::launch_transit_logp_grad(t, y, ivar, N, elem, cl, mean, C, (int)P, dtk, wk, K, sp, moments, partial, st, &ma);
  const int nacc = kNBase + (int)P * XO_NGRAD;
  const int64_t tot = C * nacc;
  xo::launch(fold_partials_kernel, (unsigned)((tot + 127) / 128), 128, 0, st, partial, sp.slots, nacc, C, (int)P, lognorm, logp, gel, gcl,
             gmean);
  return xo::launch_status();
}

int xo_transit_lc_vjp_ex(const double *t, int64_t N, const double *elem, const double *cl, int64_t C, int64_t P,
                         const double *dtk, const double *wk, int K, const double *ttv_times, const int *ttv_count, int ttv_nt,
                         int light_delay, const double *texp_vec, const double *bar_lc, double *gel, double *gcl, double *gttv,
                         void *workspace, size_t workspace_bytes, void *stream) {
  XO_NVTX_SCOPE("xo_transit_lc_vjp_ex");
  if (int rc = check_batch(N, C, P)) return rc;
  if (K < 1 || K > XO_MAX_STENCIL) return XO_ERANGE;
  if (C == 0) return 0;
  if (N == 0 || P == 0) return XO_EINVAL;
  if (!t || !elem || !cl || !dtk || !wk || !bar_lc || !gel || !gcl || !workspace) return XO_EINVAL;
  if (ttv_times && (!ttv_count || ttv_nt < 1 || !gttv || light_delay)) return XO_EINVAL;
  if (texp_vec && light_delay) return XO_EINVAL;  // the delayed reverse pass has no per-cadence stencil (use the scalar form)
  if (workspace_bytes < xo_transit_logp_grad_workspace(N, C, P) + sizeof(double) * 2 * (size_t)C) return XO_EWORKSPACE;
  Split sp = split_transit(N, C);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ttv_times) {
    cudaError_t e = cudaMemsetAsync(gttv, 0, sizeof(double) * (size_t)C * P * ttv_nt, st);
    if (e != cudaSuccess) return -(int)e;
  }
  double *partial = static_cast<double *>(workspace);
  double *scratch = partial + (size_t)C * sp.slots * (kNBase + P * XO_NGRAD);
  xo::ModelArgs ma;
  ma.elements_only = 1;
  ma.ttv_times = ttv_times;
  ma.ttv_count = ttv_count;
  ma.ttv_nt = ttv_nt;
  ma.ttv_grad = ttv_times ? gttv : nullptr;
  ma.light_delay = light_delay != 0;
  ma.texp_vec = texp_vec;
  xo::launch_transit_logp_grad(t, bar_lc, nullptr, N, elem, cl, nullptr, C, (int)P, dtk, wk, K, sp, nullptr, partial, st, &ma);
  const int nacc = kNBase + (int)P * XO_NGRAD;
  const int64_t tot = C * nacc;
  xo::launch(fold_partials_kernel, (unsigned)((tot + 127) / 128), 128, 0, st, partial, sp.slots, nacc, C, (int)P, 0.0, scratch, gel, gcl,
             scratch + C);
  return xo::launch_status();
}

int xo_orbit_eval(const double *t, int64_t N, const double *elem, int64_t C, int64_t P, double *X, double *Y, double *Z,
                  double *rv, void *stream) {
  XO_NVTX_SCOPE("xo_orbit_eval");
  if (int rc = check_batch(N, C, P)) return rc;
  if (N == 0 || C == 0 || P == 0) return 0;
  if (!t || !elem) return XO_EINVAL;
  const int sms = xo::sm_count();
  int64_t per_chain = (N * P + 255) / 256;
  int64_t want = ((int64_t)sms * 8 + C - 1) / C;
  int gx = (int)(per_chain < want ? per_chain : want);
  if (gx < 1) gx = 1;
  dim3 grid(gx, (unsigned)C);
  xo::launch(orbit_eval_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), t, N, elem, (int)P, X, Y, Z, rv);
  return xo::launch_status();
}

size_t xo_rv_logp_grad_workspace(int64_t N, int64_t C, int64_t P) {
  if (N <= 0 || C <= 0 || P <= 0) return 0;
  Split sp = split_rv(N, C);
  return sizeof(double) * (size_t)C * (size_t)sp.slots * (size_t)(2 + P * XO_NGRAD_RV);
}

int xo_rv_logp_grad(const double *t, const double *y, const double *ivar, int64_t N, const double *elem,
                    const double *gamma, int64_t C, int64_t P, double lognorm, double *logp, double *grv, double *ggamma,
                    void *workspace, size_t workspace_bytes, void *stream) {
  XO_NVTX_SCOPE("xo_rv_logp_grad");
  if (int rc = check_batch(N, C, P)) return rc;
  if (C == 0) return 0;
  if (N == 0 || P == 0) return XO_EINVAL;
  if (!t || !y || !ivar || !elem || !gamma || !logp || !grv || !ggamma || !workspace) return XO_EINVAL;
  if (workspace_bytes < xo_rv_logp_grad_workspace(N, C, P)) return XO_EWORKSPACE;
  Split sp = split_rv(N, C);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double *partial = static_cast<double *>(workspace);
  dim3 grid(sp.slots, (unsigned)C), block(kRvTile, (unsigned)P);
  static const bool cap = [] {  // XO_RV_REGCAP=0: uncapped build (66 registers, 2 CTAs of 320 threads per SM)
    const char *v = getenv("XO_RV_REGCAP");
    return !(v && *v == '0');
  }();
  if (cap)
    xo::launch(rv_logp_grad_kernel<true>, grid, block, 0, st, t, y, ivar, N, elem, gamma, (int)P, sp.L, sp.slots, partial);
  else
    xo::launch(rv_logp_grad_kernel<false>, grid, block, 0, st, t, y, ivar, N, elem, gamma, (int)P, sp.L, sp.slots, partial);
  const int nacc = 2 + (int)P * XO_NGRAD_RV;
  const int64_t tot = C * nacc;
  xo::launch(fold_rv_kernel, (unsigned)((tot + 127) / 128), 128, 0, st, partial, sp.slots, nacc, C, (int)P, lognorm, logp, grv,
                                                                ggamma);
  return xo::launch_status();
}

}  // extern "C"
