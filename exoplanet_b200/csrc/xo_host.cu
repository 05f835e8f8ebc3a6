// xo_host.cu -- host-buffer entry points and library metadata.
//
// These have the shape of exoplanet-core's pybind11 `driver` (flat contiguous fp64 host
// arrays filled in place), i.e. what PyTensor's Op.perform would hand over at
//   /root/reference/src/exoplanet/orbits/keplerian.py:333, :744-753
//   /root/reference/src/exoplanet/light_curves/limb_dark.py:24
// Each call: stream-ordered scratch (cudaMallocAsync) -> H2D -> kernel -> D2H -> sync -> free.
// Nothing persists between calls; no CPU arithmetic path exists in this library.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>

#include "xo_common.cuh"
#include "xo_transit.cuh"

namespace {

// Per host thread and device: one non-blocking stream, created on first use, and the
// device's default memory pool told to keep freed blocks (release threshold = max) so that
// cudaMallocAsync/cudaFreeAsync in a sampling loop do not go back to the OS every call.
// This is the only state the library keeps; it never affects results.
struct DeviceScope {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  cudaStream_t st = nullptr;
  cudaStream_t side = nullptr;            // second stream: parameter-only work that overlaps the data upload
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  explicit DeviceScope(int device) {
    struct Cached {
      cudaStream_t st = nullptr, side = nullptr;
      cudaEvent_t fork = nullptr, join = nullptr;
    };
    static thread_local Cached cache[64];
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess) err = cudaSetDevice(device);
    if (err != cudaSuccess) return;
    if (device >= 0 && device < 64 && cache[device].st) {
      st = cache[device].st;
      side = cache[device].side;
      ev_fork = cache[device].fork;
      ev_join = cache[device].join;
      return;
    }
    err = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming);
    if (err != cudaSuccess) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    if (device >= 0 && device < 64) {
      cache[device].st = st;
      cache[device].side = side;
      cache[device].fork = ev_fork;
      cache[device].join = ev_join;
    }
  }
  ~DeviceScope() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Per-thread pinned staging buffers: the small per-call arrays (parameters in, results out) are
// packed through them so that each direction is ONE DMA instead of 4-5 tiny ones, and pageable
// caller memory never meets cudaMemcpyAsync.  Grown on demand, kept for the thread's lifetime.
struct HostStage {
  char *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = n < 65536 ? 65536 : n;
    cudaError_t e = cudaHostAlloc((void **)&p, want, cudaHostAllocDefault);
    if (e == cudaSuccess) cap = want;
    return e;
  }
};
static thread_local HostStage g_stage_in, g_stage_out;

// a bump allocator over ONE stream-ordered allocation
struct Arena {
  char *base = nullptr;
  size_t off = 0, cap = 0;
  cudaStream_t st;
  cudaStream_t also = nullptr;  // a second stream that touched the arena: the free is ordered after it too
  cudaEvent_t also_ev = nullptr;
  cudaError_t err = cudaSuccess;
  Arena(size_t bytes, cudaStream_t s) : cap(bytes), st(s) { err = cudaMallocAsync((void **)&base, bytes ? bytes : 256, s); }
  ~Arena() {
    if (!base) return;
    if (also && also_ev && cudaEventRecord(also_ev, also) == cudaSuccess) cudaStreamWaitEvent(st, also_ev, 0);
    cudaFreeAsync(base, st);
  }
  template <typename T>
  T *take(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~(size_t)255;
    T *p = reinterpret_cast<T *>(base + off);
    off += bytes;
    return p;
  }
};
inline size_t pad(size_t n, size_t sz = 8) { return (n * sz + 255) & ~(size_t)255; }

// Once a host entry point has enqueued asynchronous copies that touch the CALLER's buffers it must not return before they
// have completed, also on an error path (the caller is free to release the buffers as soon as the call returns).
struct StreamGuard {
  cudaStream_t a, b;
  bool armed = true;
  explicit StreamGuard(cudaStream_t s, cudaStream_t side = nullptr) : a(s), b(side) {}
  void disarm() { armed = false; }
  ~StreamGuard() {
    if (!armed) return;
    if (b) cudaStreamSynchronize(b);
    cudaStreamSynchronize(a);
  }
};

#define XO_TRY(expr)                         \
  do {                                       \
    cudaError_t _e = (expr);                 \
    if (_e != cudaSuccess) return -(int)_e;  \
  } while (0)
#define XO_RC(expr)        \
  do {                     \
    int _rc = (expr);      \
    if (_rc != 0) return _rc; \
  } while (0)

inline cudaError_t h2d(void *d, const void *h, size_t bytes, cudaStream_t st) {
  return cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st);
}
inline cudaError_t d2h(void *h, const void *d, size_t bytes, cudaStream_t st) {
  return cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st);
}

}  // namespace

extern "C" {

const char *xo_version(void) { return "xo_b200 0.1.0 (sm_100a)"; }
int xo_abi_version(void) { return XO_ABI_VERSION; }

int xo_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

const char *xo_strerror(int status) {
  if (status == 0) return "ok";
  if (status == XO_EINVAL) return "xo_b200: invalid argument (null pointer or negative size)";
  if (status == XO_ERANGE) return "xo_b200: size above a compiled limit (planets, stencil, chains or cadences)";
  if (status == XO_EWORKSPACE) return "xo_b200: workspace too small";
  if (status < 0) return cudaGetErrorString((cudaError_t)(-status));
  return "xo_b200: unknown status";
}

int xo_kepler_host(const double *M, const double *e, double *sinf, double *cosf, int64_t n, int device) {
  XO_NVTX_SCOPE("xo_kepler_host");
  if (n < 0) return XO_EINVAL;
  if (n == 0) return 0;
  if (!M || !e || !sinf || !cosf) return XO_EINVAL;
  DeviceScope ds(device);
  XO_TRY(ds.err);
  StreamGuard guard(ds.st, ds.side);  // declared before the arena: the stream-ordered free is enqueued first, then the guard waits
  const size_t b = (size_t)n * 8;
  Arena ar(4 * pad(n), ds.st);
  XO_TRY(ar.err);
  double *dM = ar.take<double>(n), *de = ar.take<double>(n), *ds_ = ar.take<double>(n), *dc = ar.take<double>(n);
  XO_TRY(h2d(dM, M, b, ds.st));
  XO_TRY(h2d(de, e, b, ds.st));
  XO_RC(xo_kepler_fwd(dM, de, ds_, dc, n, ds.st));
  XO_TRY(d2h(sinf, ds_, b, ds.st));
  XO_TRY(d2h(cosf, dc, b, ds.st));
  XO_TRY(cudaStreamSynchronize(ds.st));
  guard.disarm();
  return 0;
}

int xo_qsv_host(const double *b, const double *r, double *s, double *dsdb, double *dsdr, int64_t n, int device) {
  XO_NVTX_SCOPE("xo_qsv_host");
  if (n < 0) return XO_EINVAL;
  if (n == 0) return 0;
  if (!b || !r || !s || ((dsdb == nullptr) != (dsdr == nullptr))) return XO_EINVAL;
  DeviceScope ds(device);
  XO_TRY(ds.err);
  StreamGuard guard(ds.st, ds.side);  // declared before the arena: the stream-ordered free is enqueued first, then the guard waits
  const size_t nb = (size_t)n * 8;
  Arena ar(2 * pad(n) + 3 * pad(3 * n), ds.st);
  XO_TRY(ar.err);
  double *db = ar.take<double>(n), *dr = ar.take<double>(n), *dS = ar.take<double>(3 * n);
  double *dB = dsdb ? ar.take<double>(3 * n) : nullptr, *dR = dsdb ? ar.take<double>(3 * n) : nullptr;
  XO_TRY(h2d(db, b, nb, ds.st));
  XO_TRY(h2d(dr, r, nb, ds.st));
  XO_RC(xo_qsv_fwd(db, dr, dS, dB, dR, n, ds.st));
  XO_TRY(d2h(s, dS, 3 * nb, ds.st));
  if (dsdb) {
    XO_TRY(d2h(dsdb, dB, 3 * nb, ds.st));
    XO_TRY(d2h(dsdr, dR, 3 * nb, ds.st));
  }
  XO_TRY(cudaStreamSynchronize(ds.st));
  guard.disarm();
  return 0;
}

int xo_contact_points_host(const double *a, const double *e, const double *cosw, const double *sinw, const double *cosi,
                           const double *sini, const double *L, double *M_left, double *M_right, int32_t *flag,
                           int64_t n, int device) {
  XO_NVTX_SCOPE("xo_contact_points_host");
  if (n < 0) return XO_EINVAL;
  if (n == 0) return 0;
  if (!a || !e || !cosw || !sinw || !cosi || !sini || !L || !M_left || !M_right || !flag) return XO_EINVAL;
  DeviceScope ds(device);
  XO_TRY(ds.err);
  StreamGuard guard(ds.st, ds.side);  // declared before the arena: the stream-ordered free is enqueued first, then the guard waits
  const size_t nb = (size_t)n * 8;
  Arena ar(9 * pad(n) + pad(n, 4), ds.st);
  XO_TRY(ar.err);
  const double *src[7] = {a, e, cosw, sinw, cosi, sini, L};
  double *din[7];
  for (int k = 0; k < 7; ++k) {
    din[k] = ar.take<double>(n);
    XO_TRY(h2d(din[k], src[k], nb, ds.st));
  }
  double *dl = ar.take<double>(n), *dr = ar.take<double>(n);
  int32_t *df = ar.take<int32_t>(n);
  XO_RC(xo_contact_points(din[0], din[1], din[2], din[3], din[4], din[5], din[6], dl, dr, df, n, ds.st));
  XO_TRY(d2h(M_left, dl, nb, ds.st));
  XO_TRY(d2h(M_right, dr, nb, ds.st));
  XO_TRY(d2h(flag, df, (size_t)n * 4, ds.st));
  XO_TRY(cudaStreamSynchronize(ds.st));
  guard.disarm();
  return 0;
}

int xo_transit_lc_host(const double *t, int64_t N, const double *elem, const double *cl, int64_t C, int64_t P,
                       const double *dtk, const double *wk, int K, double *lc, int device) {
  XO_NVTX_SCOPE("xo_transit_lc_host");
  if (N < 0 || C < 0 || P < 0 || K < 1) return XO_EINVAL;
  if (N == 0 || C == 0 || P == 0) return 0;
  if (!t || !elem || !cl || !dtk || !wk || !lc) return XO_EINVAL;
  DeviceScope ds(device);
  XO_TRY(ds.err);
  StreamGuard guard(ds.st, ds.side);  // declared before the arena: the stream-ordered free is enqueued first, then the guard waits
  Arena ar(pad(N) + pad(C * P * XO_NEL) + pad(3 * C) + 2 * pad(K) + pad(C * N * P), ds.st);
  XO_TRY(ar.err);
  double *dt = ar.take<double>(N), *del = ar.take<double>(C * P * XO_NEL), *dcl = ar.take<double>(3 * C);
  double *dd = ar.take<double>(K), *dw = ar.take<double>(K), *dlc = ar.take<double>(C * N * P);
  XO_TRY(h2d(dt, t, (size_t)N * 8, ds.st));
  XO_TRY(h2d(del, elem, (size_t)C * P * XO_NEL * 8, ds.st));
  XO_TRY(h2d(dcl, cl, (size_t)C * 3 * 8, ds.st));
  XO_TRY(h2d(dd, dtk, (size_t)K * 8, ds.st));
  XO_TRY(h2d(dw, wk, (size_t)K * 8, ds.st));
  XO_RC(xo_transit_lc_fwd(dt, N, del, dcl, C, P, dd, dw, K, dlc, ds.st));
  XO_TRY(d2h(lc, dlc, (size_t)C * N * P * 8, ds.st));
  XO_TRY(cudaStreamSynchronize(ds.st));
  guard.disarm();
  return 0;
}

int xo_transit_logp_grad_host(const double *t, const double *y, const double *ivar, int64_t N, const double *elem,
                              const double *cl, const double *mean, int64_t C, int64_t P, const double *dtk,
                              const double *wk, int K, double lognorm, double *logp, double *gel, double *gcl,
                              double *gmean, int device) {
  XO_NVTX_SCOPE("xo_transit_logp_grad_host");
  if (N <= 0 || C < 0 || P <= 0 || K < 1) return XO_EINVAL;
  if (C == 0) return 0;
  if (!t || !y || !ivar || !elem || !cl || !mean || !dtk || !wk || !logp || !gel || !gcl || !gmean) return XO_EINVAL;
  DeviceScope ds(device);
  XO_TRY(ds.err);
  StreamGuard guard(ds.st, ds.side);  // declared before the arena: the stream-ordered free is enqueued first, then the guard waits
  const size_t ws = xo_transit_logp_grad_workspace(N, C, P);
  const size_t mom_len = xo_transit_moments_len(N);  // moments + time index
  Arena ar(3 * pad(N) + pad(C * P * XO_NEL) + pad(3 * C) + pad(C) + 2 * pad(K) + pad(C) + pad(C * P * XO_NGRAD) +
               pad(3 * C) + pad(C) + pad(mom_len) + pad(ws, 1),
           ds.st);
  XO_TRY(ar.err);
  double *dt = ar.take<double>(N), *dy = ar.take<double>(N), *dv = ar.take<double>(N);
  double *del = ar.take<double>(C * P * XO_NEL), *dcl = ar.take<double>(3 * C), *dm = ar.take<double>(C);
  double *dd = ar.take<double>(K), *dw = ar.take<double>(K);
  double *dlp = ar.take<double>(C), *dge = ar.take<double>(C * P * XO_NGRAD), *dgc = ar.take<double>(3 * C),
         *dgm = ar.take<double>(C);
  double *dmom = ar.take<double>(mom_len);
  void *dws = ar.take<char>(ws);
  XO_TRY(h2d(dt, t, (size_t)N * 8, ds.st));
  XO_TRY(h2d(dy, y, (size_t)N * 8, ds.st));
  XO_TRY(h2d(dv, ivar, (size_t)N * 8, ds.st));
  XO_TRY(h2d(del, elem, (size_t)C * P * XO_NEL * 8, ds.st));
  XO_TRY(h2d(dcl, cl, (size_t)C * 3 * 8, ds.st));
  XO_TRY(h2d(dm, mean, (size_t)C * 8, ds.st));
  XO_TRY(h2d(dd, dtk, (size_t)K * 8, ds.st));
  XO_TRY(h2d(dw, wk, (size_t)K * 8, ds.st));
  // data moments + time index (one pass, a few microseconds): the sweep then reads no time stamps
  XO_RC(xo_transit_data_moments_index(dt, dy, dv, N, dmom, mom_len, ds.st));
  XO_RC(xo_transit_logp_grad(dt, dy, dv, N, del, dcl, dm, C, P, dd, dw, K, lognorm, dmom, dlp, dge, dgc, dgm, dws, ws, ds.st));
  XO_TRY(d2h(logp, dlp, (size_t)C * 8, ds.st));
  XO_TRY(d2h(gel, dge, (size_t)C * P * XO_NGRAD * 8, ds.st));
  XO_TRY(d2h(gcl, dgc, (size_t)C * 3 * 8, ds.st));
  XO_TRY(d2h(gmean, dgm, (size_t)C * 8, ds.st));
  XO_TRY(cudaStreamSynchronize(ds.st));
  guard.disarm();
  return 0;
}

int xo_transit_model_logp_grad_host(const double *t, const double *y, const double *ivar, int64_t N, const double *par,
                                    const double *star, const double *mean, int64_t C, int64_t P, const double *dtk,
                                    const double *wk, int K, double texp, int use_in_transit, double lognorm,
                                    double *logp, double *gpar, double *gstar, double *gmean, int device) {
  XO_NVTX_SCOPE("xo_transit_model_logp_grad_host");
  if (N <= 0 || C < 0 || P <= 0 || K < 1) return XO_EINVAL;
  if (C == 0) return 0;
  if (!t || !y || !ivar || !par || !star || !mean || !dtk || !wk || !logp || !gpar || !gstar || !gmean) return XO_EINVAL;
  DeviceScope ds(device);
  XO_TRY(ds.err);
  StreamGuard guard(ds.st, ds.side);  // declared before the arena: the stream-ordered free is enqueued first, then the guard waits
  const size_t ws = xo_transit_model_workspace(N, C, P);
  const size_t nin = (size_t)C * P * XO_NPAR + (size_t)C * XO_NSTAR + (size_t)C + 2 * (size_t)K;
  const size_t nout = (size_t)C + (size_t)C * P * XO_NPAR + (size_t)C * XO_NSTAR + (size_t)C;
  const size_t mom_len = xo_transit_moments_len(N);  // moments + time index
  Arena ar(pad(3 * N) + pad(nin) + pad(nout) + pad(mom_len) + pad(ws, 1), ds.st);
  XO_TRY(ar.err);
  double *dt = ar.take<double>(3 * N), *dy = dt + N, *dv = dy + N;  // one block: one DMA when the caller's arrays are one block too
  double *din = ar.take<double>(nin), *dout = ar.take<double>(nout);
  double *dmom = ar.take<double>(mom_len);
  void *dws = ar.take<char>(ws);
  double *dp = din, *dst = dp + C * P * XO_NPAR, *dm = dst + C * XO_NSTAR, *dd = dm + C, *dw = dd + K;
  double *dlp = dout, *dgp = dlp + C, *dgs = dgp + C * P * XO_NPAR, *dgm = dgs + C * XO_NSTAR;
  XO_TRY(g_stage_in.ensure(nin * 8));
  XO_TRY(g_stage_out.ensure(nout * 8));
  const bool split = xo::model_split_supported();
  if (split) {
    ar.also = ds.side;
    ar.also_ev = ds.ev_join;
    XO_TRY(cudaEventRecord(ds.ev_fork, ds.st));  // the arena exists from here on in stream order
  }
  // the light curve straight from the caller's (ideally pinned) memory, enqueued first: ONE DMA when t, y, ivar are
  // consecutive slices of one host block (y == t + N, ivar == y + N), else three
  if (y == t + N && ivar == y + N) {
    XO_TRY(h2d(dt, t, (size_t)3 * N * 8, ds.st));
  } else {
    XO_TRY(h2d(dt, t, (size_t)N * 8, ds.st));
    XO_TRY(h2d(dy, y, (size_t)N * 8, ds.st));
    XO_TRY(h2d(dv, ivar, (size_t)N * 8, ds.st));
  }
  cudaStream_t ps = split ? ds.side : ds.st;  // the parameters and the prologue overlap that upload
  if (split) XO_TRY(cudaStreamWaitEvent(ps, ds.ev_fork, 0));
  {  // the parameters: packed through pinned staging -> one DMA
    double *hin = reinterpret_cast<double *>(g_stage_in.p);
    size_t o = 0;
    auto put = [&](const double *src, size_t n) {
      memcpy(hin + o, src, n * 8);
      o += n;
    };
    put(par, (size_t)C * P * XO_NPAR);
    put(star, (size_t)C * XO_NSTAR);
    put(mean, (size_t)C);
    put(dtk, (size_t)K);
    put(wk, (size_t)K);
    XO_TRY(h2d(din, hin, nin * 8, ps));
  }
  if (split) {
    XO_RC(xo::model_logp_grad_phases(dt, dy, dv, N, dp, dst, dm, C, P, dd, dw, K, texp, use_in_transit, lognorm, dmom, dlp, dgp,
                                     dgs, dgm, dws, ws, ps, 1));
    XO_TRY(cudaEventRecord(ds.ev_join, ps));
  }
  // data moments and the time index in one pass (a few microseconds): the sweep then reads no time stamp at all and the out-of-
  // transit Gaussian terms are closed-form
  XO_RC(xo_transit_data_moments_index(dt, dy, dv, N, dmom, mom_len, ds.st));
  if (split) {
    XO_TRY(cudaStreamWaitEvent(ds.st, ds.ev_join, 0));
    ar.also = nullptr;  // joined: the main stream is ordered after everything the side stream did
  }
  XO_RC(xo::model_logp_grad_phases(dt, dy, dv, N, dp, dst, dm, C, P, dd, dw, K, texp, use_in_transit, lognorm, dmom, dlp, dgp, dgs,
                                   dgm, dws, ws, ds.st, split ? 2 : 3));
  double *hout = reinterpret_cast<double *>(g_stage_out.p);
  XO_TRY(d2h(hout, dout, nout * 8, ds.st));
  XO_TRY(cudaStreamSynchronize(ds.st));
  size_t o = 0;
  auto get = [&](double *dst_, size_t n) {
    memcpy(dst_, hout + o, n * 8);
    o += n;
  };
  get(logp, (size_t)C);
  get(gpar, (size_t)C * P * XO_NPAR);
  get(gstar, (size_t)C * XO_NSTAR);
  get(gmean, (size_t)C);
  return 0;
}

// ---- resident data set: t, y, ivar (+ their moments) uploaded once, owned by the caller ----
struct xo_transit_data {
  int device = 0;
  int64_t N = 0;
  double *t = nullptr, *y = nullptr, *ivar = nullptr, *moments = nullptr;
  char *scratch = nullptr;  // parameter staging + workspace, grown on demand
  size_t scratch_bytes = 0;
  cudaStream_t st = nullptr;
};

int xo_transit_data_create(const double *t, const double *y, const double *ivar, int64_t N, int device,
                           xo_transit_data **out) {
  XO_NVTX_SCOPE("xo_transit_data_create");
  if (!t || !y || !ivar || N <= 0 || !out) return XO_EINVAL;
  int prev = 0;
  XO_TRY(cudaGetDevice(&prev));
  XO_TRY(cudaSetDevice(device));
  xo_transit_data *d = new (std::nothrow) xo_transit_data();
  if (!d) return XO_EINVAL;
  d->device = device;
  d->N = N;
  cudaError_t e = cudaStreamCreateWithFlags(&d->st, cudaStreamNonBlocking);
  const size_t nb = (size_t)N * 8;
  const size_t mom_len = xo_transit_moments_len(N);
  if (e == cudaSuccess) e = cudaMalloc((void **)&d->t, 3 * nb + mom_len * 8);
  if (e == cudaSuccess) {
    d->y = d->t + N;
    d->ivar = d->y + N;
    d->moments = d->ivar + N;
    e = cudaMemcpyAsync(d->t, t, nb, cudaMemcpyHostToDevice, d->st);
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(d->y, y, nb, cudaMemcpyHostToDevice, d->st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d->ivar, ivar, nb, cudaMemcpyHostToDevice, d->st);
  int rc = e == cudaSuccess ? xo_transit_data_moments_index(d->t, d->y, d->ivar, N, d->moments, mom_len, d->st) : -(int)e;
  if (rc == 0) rc = xo::cuda_status(cudaStreamSynchronize(d->st));
  cudaSetDevice(prev);
  if (rc != 0) {
    xo_transit_data_destroy(d);
    return rc;
  }
  *out = d;
  return 0;
}

int xo_transit_data_destroy(xo_transit_data *d) {
  XO_NVTX_SCOPE("xo_transit_data_destroy");
  if (!d) return 0;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(d->device);
  if (d->t) cudaFree(d->t);
  if (d->scratch) cudaFree(d->scratch);
  if (d->st) cudaStreamDestroy(d->st);
  cudaSetDevice(prev);
  delete d;
  return 0;
}

// parameters and results are HOST buffers; the light curve stays on the device
int xo_transit_model_logp_grad_resident(xo_transit_data *d, const double *par, const double *star, const double *mean,
                                        int64_t C, int64_t P, const double *dtk, const double *wk, int K, double texp,
                                        int use_in_transit, double lognorm, int use_moments, double *logp, double *gpar,
                                        double *gstar, double *gmean) {
  XO_NVTX_SCOPE("xo_transit_model_logp_grad_resident");
  if (!d || C < 0 || P <= 0 || K < 1) return XO_EINVAL;
  if (C == 0) return 0;
  if (!par || !star || !mean || !dtk || !wk || !logp || !gpar || !gstar || !gmean) return XO_EINVAL;
  int prev = 0;
  XO_TRY(cudaGetDevice(&prev));
  XO_TRY(cudaSetDevice(d->device));
  const size_t ws = xo_transit_model_workspace(d->N, C, P);
  const size_t nin = (size_t)C * P * XO_NPAR + (size_t)C * XO_NSTAR + (size_t)C + 2 * (size_t)K;
  const size_t nout = (size_t)C + (size_t)C * P * XO_NPAR + (size_t)C * XO_NSTAR + (size_t)C;
  const size_t need = pad(nin) + pad(nout) + pad(ws, 1);
  if (need > d->scratch_bytes) {
    if (d->scratch) cudaFree(d->scratch);
    d->scratch = nullptr;
    d->scratch_bytes = 0;
    cudaError_t e = cudaMalloc((void **)&d->scratch, need);
    if (e != cudaSuccess) {
      cudaSetDevice(prev);
      return -(int)e;
    }
    d->scratch_bytes = need;
  }
  double *din = reinterpret_cast<double *>(d->scratch);
  double *dout = reinterpret_cast<double *>(d->scratch + pad(nin));
  void *dws = d->scratch + pad(nin) + pad(nout);
  double *dp = din, *dst = dp + C * P * XO_NPAR, *dm = dst + C * XO_NSTAR, *dd = dm + C, *dw = dd + K;
  double *dlp = dout, *dgp = dlp + C, *dgs = dgp + C * P * XO_NPAR, *dgm = dgs + C * XO_NSTAR;
  cudaStream_t st = d->st;
  int rc = xo::cuda_status(g_stage_in.ensure(nin * 8));
  if (rc == 0) rc = xo::cuda_status(g_stage_out.ensure(nout * 8));
  if (rc == 0) {
    // pack [par | star | mean | dtk | wk] exactly as the device block is laid out -> one H2D
    double *hin = reinterpret_cast<double *>(g_stage_in.p);
    size_t o = 0;
    auto put = [&](const double *src, size_t n) {
      memcpy(hin + o, src, n * 8);
      o += n;
    };
    put(par, (size_t)C * P * XO_NPAR);
    put(star, (size_t)C * XO_NSTAR);
    put(mean, (size_t)C);
    put(dtk, (size_t)K);
    put(wk, (size_t)K);
    rc = xo::cuda_status(cudaMemcpyAsync(din, hin, nin * 8, cudaMemcpyHostToDevice, st));
  }
  if (rc == 0)
    rc = xo_transit_model_logp_grad(d->t, d->y, d->ivar, d->N, dp, dst, dm, C, P, dd, dw, K, texp, use_in_transit, lognorm,
                                    use_moments ? d->moments : nullptr, dlp, dgp, dgs, dgm, dws, ws, st);
  double *hout = reinterpret_cast<double *>(g_stage_out.p);
  if (rc == 0) rc = xo::cuda_status(cudaMemcpyAsync(hout, dout, nout * 8, cudaMemcpyDeviceToHost, st));
  if (rc == 0) rc = xo::cuda_status(cudaStreamSynchronize(st));
  if (rc == 0) {  // [logp | gpar | gstar | gmean]
    size_t o = 0;
    auto get = [&](double *dst_, size_t n) {
      memcpy(dst_, hout + o, n * 8);
      o += n;
    };
    get(logp, (size_t)C);
    get(gpar, (size_t)C * P * XO_NPAR);
    get(gstar, (size_t)C * XO_NSTAR);
    get(gmean, (size_t)C);
  }
  cudaSetDevice(prev);
  return rc;
}

}  // extern "C"
