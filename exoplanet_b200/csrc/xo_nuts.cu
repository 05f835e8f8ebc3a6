// xo_nuts.cu -- the tree bookkeeping of a lock-step No-U-Turn sampler as four kernels.
//
// What it replaces: PyMC's NUTS, the sampler the reference's models are run with
// (/root/reference/docs/tutorials/data-and-models.md:505-509), one chain per OS process
// (/root/reference/docs/user/multiprocessing.rst).  exoplanet_b200/sampling.py::BatchedNUTS advances C chains together:
// multinomial NUTS with the generalised U-turn criterion (Hoffman & Gelman 2014; Betancourt 2017), doubling j integrating
// 2^j leapfrog steps from one end of every chain's tree.  Written with tensor operations that is ~50 tiny launches per leaf
// around the model evaluation; here one leaf is
//     xo_nuts_leaf_pre   half kick + drift                                   (C x D threads)
//     <logp + gradient at the new point: any callback; the fused transit posterior is 5 launches>
//     xo_nuts_leaf_post  half kick, energy error, divergence, acceptance statistic, progressive multinomial sampling
//                        inside the new half, momentum sums, U-turn checkpoints / checks        (one warp per chain)
// and one doubling adds xo_nuts_start (direction, start from the chosen edge) and xo_nuts_merge (merge the half into the
// tree, whole-tree U-turn check, who keeps growing).  The kernels are deterministic functions of the same inputs as the
// tensor code -- the uniform random numbers are drawn by the caller -- so the two implementations can be compared
// state for state (tests/test_gpu_sampling.py).
//
// State: one block of doubles (flags are 0.0 / 1.0), laid out by xo_nuts_state_layout():
//   per chain  [C]      h0 log_w lp_prop active depth diverged acc_sum n_leaves dh_max sgn sub_ok s_logw s_lp go_right
//   per coord  [C][D]   ql pl gl qr pr gr q_prop g_prop p_sum q p g s_q s_g s_psum
//   checkpoints [max_depth + 1][C][D]   ck_p ck_s
#include <cuda_runtime.h>

#include "xo_common.cuh"

namespace {

constexpr unsigned kFullN = 0xffffffffu;
enum Scalar { H0, LOG_W, LP_PROP, ACTIVE, DEPTH, DIVERGED, ACC_SUM, N_LEAVES, DH_MAX, SGN, SUB_OK, S_LOGW, S_LP, GO_RIGHT, kNScalar };
enum Vector { QL, PL, GL, QR, PR, GR, Q_PROP, G_PROP, P_SUM, Q, P, G, S_Q, S_G, S_PSUM, kNVector };

struct Layout {
  int64_t C;
  int D, max_depth;
  __host__ __device__ int64_t scalar(int k) const { return (int64_t)k * C; }
  __host__ __device__ int64_t vector(int k) const { return (int64_t)kNScalar * C + (int64_t)k * C * D; }
  __host__ __device__ int64_t ck_p(int level) const { return (int64_t)kNScalar * C + (int64_t)kNVector * C * D + (int64_t)level * C * D; }
  __host__ __device__ int64_t ck_s(int level) const { return ck_p(max_depth + 1) + (int64_t)level * C * D; }
  __host__ __device__ int64_t total() const { return ck_s(max_depth + 1); }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullN, v, o);
  return v;
}
__device__ __forceinline__ double logaddexp(double a, double b) {
  const double m = fmax(a, b);
  if (m == -INFINITY) return -INFINITY;
  return m + log1p(exp(-fabs(a - b)));
}
// generalised U-turn criterion on a (sub-)tree with end momenta pl, pr and momentum sum ps (this lane's coordinates)
__device__ __forceinline__ bool turning(const double *im, const double *pl, const double *pr, const double *ps, int D, int lane) {
  double a = 0.0, b = 0.0;
  for (int j = lane; j < D; j += 32) {
    const double adj = ps[j] - 0.5 * (pl[j] + pr[j]);
    a += im[j] * pl[j] * adj;
    b += im[j] * pr[j] * adj;
  }
  a = warp_sum(a);
  b = warp_sum(b);
  return a <= 0.0 || b <= 0.0;
}

// ---- a doubling starts: direction per chain, the new half grows out of the chosen edge --------------------------------
__global__ void nuts_start_kernel(double *__restrict__ st, Layout L, const double *__restrict__ u_dir, const double *__restrict__ logp_cur) {
  const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= L.C) return;
  const int lane = threadIdx.x & 31, D = L.D;
  const bool right = u_dir[c] < 0.5;
  const int64_t o = c * D;
  const double *qe = st + L.vector(right ? QR : QL) + o, *pe = st + L.vector(right ? PR : PL) + o, *ge = st + L.vector(right ? GR : GL) + o;
  for (int j = lane; j < D; j += 32) {
    st[L.vector(Q) + o + j] = qe[j];
    st[L.vector(P) + o + j] = pe[j];
    st[L.vector(G) + o + j] = ge[j];
    st[L.vector(S_Q) + o + j] = qe[j];
    st[L.vector(S_G) + o + j] = ge[j];
    st[L.vector(S_PSUM) + o + j] = 0.0;
  }
  if (lane == 0) {
    st[L.scalar(GO_RIGHT) + c] = right ? 1.0 : 0.0;
    st[L.scalar(SGN) + c] = right ? 1.0 : -1.0;
    st[L.scalar(SUB_OK) + c] = st[L.scalar(ACTIVE) + c];
    st[L.scalar(S_LOGW) + c] = -INFINITY;
    st[L.scalar(S_LP) + c] = logp_cur[c];
  }
}

// ---- one leaf, before the model evaluation: p += sgn eps/2 g;  q += sgn eps M^-1 p -----------------------------------
__global__ void nuts_leaf_pre_kernel(double *__restrict__ st, Layout L, const double *__restrict__ eps, const double *__restrict__ im) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L.C * L.D) return;
  const int64_t c = i / L.D;
  const double se = st[L.scalar(SGN) + c] * eps[c];
  const double p = st[L.vector(P) + i] + 0.5 * se * st[L.vector(G) + i];
  st[L.vector(P) + i] = p;
  st[L.vector(Q) + i] += se * im[i] * p;
}

// ---- one leaf, after it: lp[C], g[C][D] are logp and its gradient at the new q ------------------------------------------
__global__ void nuts_leaf_post_kernel(double *__restrict__ st, Layout L, const double *__restrict__ eps, const double *__restrict__ im,
                                      const double *__restrict__ lp, const double *__restrict__ g, const double *__restrict__ u_take,
                                      int n) {
  const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= L.C) return;
  const int lane = threadIdx.x & 31, D = L.D;
  const int64_t o = c * D;
  double *p = st + L.vector(P) + o, *gs = st + L.vector(G) + o, *q = st + L.vector(Q) + o;
  const double *imc = im + o, *gn = g + o;
  const double se = st[L.scalar(SGN) + c] * eps[c];
  double kin = 0.0;
  for (int j = lane; j < D; j += 32) {
    const double pj = p[j] + 0.5 * se * gn[j];
    p[j] = pj;
    gs[j] = gn[j];
    kin += pj * pj * imc[j];
  }
  kin = warp_sum(kin);
  const double lpc = lp[c];
  const double h = -lpc + 0.5 * kin;
  double dh = h - st[L.scalar(H0) + c];
  if (!isfinite(dh)) dh = INFINITY;
  const bool active = st[L.scalar(ACTIVE) + c] != 0.0;
  bool sub_ok = st[L.scalar(SUB_OK) + c] != 0.0;
  const double dh_max = st[L.scalar(DH_MAX) + c];
  const bool div_now = sub_ok && dh > 1000.0;
  sub_ok = sub_ok && !div_now;
  const bool sub_ok_leaf = sub_ok;  // ... as the sampling inside the half sees it (the U-turn checks below come after)
  const double w_new = sub_ok ? -h : -INFINITY;
  const double s_logw = st[L.scalar(S_LOGW) + c];
  const double s_logw_new = logaddexp(s_logw, w_new);
  const bool take = sub_ok && log(u_take[c]) < w_new - s_logw_new;
  double *s_q = st + L.vector(S_Q) + o, *s_g = st + L.vector(S_G) + o, *s_psum = st + L.vector(S_PSUM) + o;
  for (int j = lane; j < D; j += 32) {
    if (take) {
      s_q[j] = q[j];
      s_g[j] = gn[j];
    }
    if (sub_ok) s_psum[j] += p[j];
  }
  __syncwarp();
  // U-turn checks inside the half: checkpoints at even leaves, checks of every completed sub-tree at odd ones
  const int idx_max = __popc((unsigned)n >> 1);
  if ((n & 1) == 0) {
    double *cp = st + L.ck_p(idx_max) + o, *cs = st + L.ck_s(idx_max) + o;
    for (int j = lane; j < D; j += 32) {
      cp[j] = p[j];
      cs[j] = s_psum[j];
    }
  } else {
    int ones = 0;
    for (unsigned m = (unsigned)n; m & 1u; m >>= 1) ++ones;
    const int idx_min = idx_max - ones + 1;
    for (int i = idx_max; i >= idx_min; --i) {
      const double *cp = st + L.ck_p(i) + o, *cs = st + L.ck_s(i) + o;
      double a = 0.0, b = 0.0;  // turning(ck_p[i], p, s_psum - ck_s[i] + ck_p[i])
      for (int j = lane; j < D; j += 32) {
        const double sub = s_psum[j] - cs[j] + cp[j];
        const double adj = sub - 0.5 * (cp[j] + p[j]);
        a += imc[j] * cp[j] * adj;
        b += imc[j] * p[j] * adj;
      }
      a = warp_sum(a);
      b = warp_sum(b);
      if (a <= 0.0 || b <= 0.0) sub_ok = false;
    }
  }
  if (lane == 0) {
    if (active && fabs(dh) > fabs(dh_max) && isfinite(dh)) st[L.scalar(DH_MAX) + c] = dh;
    if (div_now) st[L.scalar(DIVERGED) + c] = 1.0;
    if (active) {  // acceptance statistic over every leaf that was actually visited
      st[L.scalar(ACC_SUM) + c] += exp(fmin(-dh, 0.0));
      st[L.scalar(N_LEAVES) + c] += 1.0;
    }
    if (take) st[L.scalar(S_LP) + c] = lpc;
    if (sub_ok_leaf) st[L.scalar(S_LOGW) + c] = s_logw_new;
    st[L.scalar(SUB_OK) + c] = sub_ok ? 1.0 : 0.0;
  }
}

// ---- a doubling ends: merge the half into the tree where it is valid, check the whole tree, who keeps growing ----------
__global__ void nuts_merge_kernel(double *__restrict__ st, Layout L, const double *__restrict__ im, const double *__restrict__ u_merge) {
  const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= L.C) return;
  const int lane = threadIdx.x & 31, D = L.D;
  const int64_t o = c * D;
  const bool active = st[L.scalar(ACTIVE) + c] != 0.0, sub_ok = st[L.scalar(SUB_OK) + c] != 0.0;
  const bool right = st[L.scalar(GO_RIGHT) + c] != 0.0;
  const bool merged = active && sub_ok;
  const double s_logw = st[L.scalar(S_LOGW) + c], log_w = st[L.scalar(LOG_W) + c];
  const bool take = merged && log(u_merge[c]) < s_logw - log_w;
  const double *q = st + L.vector(Q) + o, *p = st + L.vector(P) + o, *g = st + L.vector(G) + o;
  double *qe = st + L.vector(right ? QR : QL) + o, *pe = st + L.vector(right ? PR : PL) + o, *ge = st + L.vector(right ? GR : GL) + o;
  double *p_sum = st + L.vector(P_SUM) + o;
  for (int j = lane; j < D; j += 32) {
    if (take) {
      st[L.vector(Q_PROP) + o + j] = st[L.vector(S_Q) + o + j];
      st[L.vector(G_PROP) + o + j] = st[L.vector(S_G) + o + j];
    }
    if (merged) {
      p_sum[j] += st[L.vector(S_PSUM) + o + j];
      qe[j] = q[j];
      pe[j] = p[j];
      ge[j] = g[j];
    }
  }
  __syncwarp();
  const bool turned = turning(im + o, st + L.vector(PL) + o, st + L.vector(PR) + o, p_sum, D, lane);
  if (lane == 0) {
    if (take) st[L.scalar(LP_PROP) + c] = st[L.scalar(S_LP) + c];
    if (merged) st[L.scalar(LOG_W) + c] = logaddexp(log_w, s_logw);
    if (active) st[L.scalar(DEPTH) + c] += 1.0;
    st[L.scalar(ACTIVE) + c] = (merged && !turned) ? 1.0 : 0.0;
  }
}

inline int check(int64_t C, int D, int max_depth, const void *st) {
  if (C < 0 || D < 1 || max_depth < 1 || max_depth > 20) return XO_EINVAL;
  if (!st) return XO_EINVAL;
  return 0;
}

}  // namespace

extern "C" {

// number of doubles of the state block, and the offsets (in doubles) of its kNScalar + kNVector + 2 fields in the order
// of the header comment (ck_p, ck_s last): offsets[0..30]
size_t xo_nuts_state_layout(int64_t C, int D, int max_depth, int64_t *offsets) {
  if (C < 0 || D < 1 || max_depth < 1) return 0;
  const Layout L{C, D, max_depth};
  if (offsets) {
    for (int k = 0; k < kNScalar; ++k) offsets[k] = L.scalar(k);
    for (int k = 0; k < kNVector; ++k) offsets[kNScalar + k] = L.vector(k);
    offsets[kNScalar + kNVector] = L.ck_p(0);
    offsets[kNScalar + kNVector + 1] = L.ck_s(0);
  }
  return (size_t)L.total();
}

int xo_nuts_start(double *state, int64_t C, int D, int max_depth, const double *u_dir, const double *logp_cur, void *stream) {
  XO_NVTX_SCOPE("xo_nuts_start");
  if (int rc = check(C, D, max_depth, state)) return rc;
  if (C == 0) return 0;
  if (!u_dir || !logp_cur) return XO_EINVAL;
  xo::launch(nuts_start_kernel, (unsigned)((C + 3) / 4), 128, 0, static_cast<cudaStream_t>(stream), state, Layout{C, D, max_depth}, u_dir,
             logp_cur);
  return xo::launch_status();
}

int xo_nuts_leaf_pre(double *state, int64_t C, int D, int max_depth, const double *eps, const double *inv_mass, void *stream) {
  XO_NVTX_SCOPE("xo_nuts_leaf_pre");
  if (int rc = check(C, D, max_depth, state)) return rc;
  if (C == 0) return 0;
  if (!eps || !inv_mass) return XO_EINVAL;
  xo::launch(nuts_leaf_pre_kernel, (unsigned)((C * D + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream), state, Layout{C, D, max_depth},
             eps, inv_mass);
  return xo::launch_status();
}

int xo_nuts_leaf_post(double *state, int64_t C, int D, int max_depth, const double *eps, const double *inv_mass, const double *logp,
                      const double *grad, const double *u_take, int leaf, void *stream) {
  XO_NVTX_SCOPE("xo_nuts_leaf_post");
  if (int rc = check(C, D, max_depth, state)) return rc;
  if (C == 0) return 0;
  if (!eps || !inv_mass || !logp || !grad || !u_take || leaf < 0 || leaf >= (1 << max_depth)) return XO_EINVAL;
  xo::launch(nuts_leaf_post_kernel, (unsigned)((C + 3) / 4), 128, 0, static_cast<cudaStream_t>(stream), state, Layout{C, D, max_depth}, eps,
             inv_mass, logp, grad, u_take, leaf);
  return xo::launch_status();
}

int xo_nuts_merge(double *state, int64_t C, int D, int max_depth, const double *inv_mass, const double *u_merge, void *stream) {
  XO_NVTX_SCOPE("xo_nuts_merge");
  if (int rc = check(C, D, max_depth, state)) return rc;
  if (C == 0) return 0;
  if (!inv_mass || !u_merge) return XO_EINVAL;
  xo::launch(nuts_merge_kernel, (unsigned)((C + 3) / 4), 128, 0, static_cast<cudaStream_t>(stream), state, Layout{C, D, max_depth}, inv_mass,
             u_merge);
  return xo::launch_status();
}

}  // extern "C"
