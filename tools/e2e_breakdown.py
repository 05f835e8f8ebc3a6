"""Where the time of one host-buffer model evaluation goes (run on the GPU box):
H2D of the light curve, the moments kernel, the three model kernels, D2H, host overhead.

    python tools/e2e_breakdown.py [reps]
"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exoplanet_b200 import _lib, fused, synthetic  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
lib = _lib.require_cuda()
N, Cn, P = 100000, 128, 1
t = synthetic.cadence_times(N)
y = synthetic.noise(N)
p = synthetic.transit_params(Cn, P, ecc=0.5)


def pinned(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).pin_memory()


def ptr(x):
    return C.c_void_p(x.data_ptr())


par = np.stack([p["period"], p["t0"], p["ecc"], p["omega"], p["b"], p["ror"], p["m_planet"], p["K"]], axis=-1)
star = np.stack([np.ones(Cn), np.ones(Cn), p["u1"], p["u2"]], axis=-1)
h = dict(t=pinned(t), y=pinned(y), ivar=pinned(np.full(N, 1 / 5e-4 ** 2)), par=pinned(par), star=pinned(star), mean=pinned(p["mean"]),
         dtk=pinned(np.zeros(1)), wk=pinned(np.ones(1)))
o = dict(logp=pinned(np.empty(Cn)), gpar=pinned(np.empty((Cn, P, 8))), gstar=pinned(np.empty((Cn, 4))), gmean=pinned(np.empty(Cn)))


def host_call():
    rc = lib.xo_transit_model_logp_grad_host(ptr(h["t"]), ptr(h["y"]), ptr(h["ivar"]), N, ptr(h["par"]), ptr(h["star"]), ptr(h["mean"]), Cn,
                                             P, ptr(h["dtk"]), ptr(h["wk"]), 1, 0.0, 1, 0.0, ptr(o["logp"]), ptr(o["gpar"]), ptr(o["gstar"]),
                                             ptr(o["gmean"]), 0)
    assert rc == 0, rc


def stats(name, fn, n=reps):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e6
    print("%-46s min %7.1f  median %7.1f  mean %7.1f  p90 %7.1f  max %8.1f us" % (name, ts.min(), np.median(ts), ts.mean(),
                                                                                   np.percentile(ts, 90), ts.max()))


d3 = torch.empty(3 * N, dtype=torch.float64, device="cuda")
big = torch.cat([h["t"], h["y"], h["ivar"]]).pin_memory()


def h2d_three():
    d3[:N].copy_(h["t"], non_blocking=True)
    d3[N:2 * N].copy_(h["y"], non_blocking=True)
    d3[2 * N:].copy_(h["ivar"], non_blocking=True)
    torch.cuda.synchronize()


def h2d_one():
    d3.copy_(big, non_blocking=True)
    torch.cuda.synchronize()


small = torch.empty(8, dtype=torch.float64, device="cuda")
hsmall = pinned(np.zeros(8))


def h2d_tiny():
    small.copy_(hsmall, non_blocking=True)
    torch.cuda.synchronize()


stats("H2D 3 x 0.8 MB pinned + sync", h2d_three)
stats("H2D 1 x 2.4 MB pinned + sync", h2d_one)
stats("H2D 64 B + sync (latency floor)", h2d_tiny)
data = fused.TransitData(t, y, np.full(N, 5e-4), device="cuda")


def mom():
    data._moments = None
    data.moments()
    torch.cuda.synchronize()


stats("moments kernel (python wrapper) + sync", mom)
handle = C.c_void_p()
assert lib.xo_transit_data_create(ptr(h["t"]), ptr(h["y"]), ptr(h["ivar"]), N, 0, C.byref(handle)) == 0


def resident():
    rc = lib.xo_transit_model_logp_grad_resident(handle, ptr(h["par"]), ptr(h["star"]), ptr(h["mean"]), Cn, P, ptr(h["dtk"]), ptr(h["wk"]), 1,
                                                 0.0, 1, 0.0, 1, ptr(o["logp"]), ptr(o["gpar"]), ptr(o["gstar"]), ptr(o["gmean"]))
    assert rc == 0, rc


stats("resident call (params in, grads out)", resident)
stats("host call (everything in, grads out)", host_call)
lib.xo_transit_data_destroy(handle)
