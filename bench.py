#!/usr/bin/env python
"""bench.py -- model+grad evaluations/s of an N=1e5-cadence eccentric (e=0.5) transit
(BASELINE.json configs[1]) on 1..8 B200, next to the CPU restatement of the reference
path timed on the same box (`cpu_baseline`, and `--impl reference`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One *step* = one model+grad evaluation of every chain of the batch (default 128 chains
per GPU -- BASELINE configs[4]'s per-GPU share -- all seeing the same N=1e5 light curve):
params -> element blocks -> fused Kepler/projection/limb-darkening logp + full gradient.
`value` = evaluations/s with inputs resident in HBM, device-timed (CUDA events per step,
L2 flushed between steps, max over ranks).  `e2e` = the same evaluations through the
host-buffer C-ABI call (host arrays in, host arrays out, H2D/D2H inside the timed region).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CADENCES = 100000
CHAINS_PER_GPU = 128
SIGMA = 5e-4
ALG_BYTES_PER_CHAIN_CADENCE = 24  # t, y, sigma (SURVEY 8d / BASELINE.md section 5)
KEPLER_BYTES_PER_ELEMENT = 32  # M, e in; sin f, cos f out


def _ncu_traffic(summary, kernel_tag):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu summary
    (profiles/*.txt written by tools/ncu_summary.py from one `ncu --set full` capture), or None."""
    path = os.path.join(ROOT, "profiles", summary)
    if not os.path.exists(path):
        return None
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, ok = 0.0, False
    for line in open(path):
        f = line.split()
        if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and f[2] in mult:
            tot += float(f[1]) * mult[f[2]]
            ok = True
    return tot if ok else None


def _ncu_metric(summary, name):
    """one metric value per launch from a committed ncu summary (profiles/*.txt), or None"""
    path = os.path.join(ROOT, "profiles", summary)
    if not os.path.exists(path):
        return None
    for line in open(path):
        f = line.split()
        if len(f) >= 2 and f[0] == name:
            try:
                return float(f[1].replace(",", ""))
            except ValueError:
                return None
    return None


def _first_profile(*names):
    for n in names:
        if os.path.exists(os.path.join(ROOT, "profiles", n)):
            return n
    return names[-1]


def probe_reference_tier():
    """every run starts by checking whether the reference's own CPU path is importable
    (a driver-provided install under baseline/_ref, or site-packages).  Tier P = the reference itself (PyTensor +
    exoplanet-core); tier R = the C restatement of the same mathematics (oracle/)."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    code = ("import sys; sys.path.insert(0, %r); import importlib; miss = []\n"
            "for m in ('pytensor', 'pymc', 'exoplanet_core', 'exoplanet'):\n"
            "    try: importlib.import_module(m)\n"
            "    except Exception as e: miss.append(m)\n"
            "print(','.join(miss))") % ref_dir
    try:
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
        missing = [m for m in out.stdout.strip().split(",") if m]
    except Exception:
        missing = ["probe failed"]
    return {"tier": "R (C restatement of the reference mathematics, oracle/xo_oracle.c -- NOT the reference binary)" if missing
            else "P (reference PyTensor CPU path)", "not_importable": missing, "baseline_ref_dir_present": os.path.isdir(ref_dir)}


WORKLOAD = "configs[1]: eccentric (e=0.5) single-planet Kepler solve + LimbDarkLightCurve, N=1e5 cadences, fp64"


def bench_config(chains, gpus):
    """the `config` object BOTH arms print (identical keys and values for the same --gpus / --chains)"""
    return {"workload": WORKLOAD, "chains_per_gpu": chains, "chains_total": chains * gpus, "cadences": N_CADENCES,
            "use_in_transit": True, "parallelism": "chain-sharded, no collective",
            "l2": "flushed between timed steps (192 MiB memset) on the GPU arm"}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        if os.environ.get("XO_BENCH_SAMPLER", "1") == "0":  # diagnostic: does the NVML polling perturb the host-side numbers?
            return
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons = [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    out["sm_max_mhz"] = float(f[2])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def build_problem(C, seed_offset=0):
    """numpy description of the workload (shared by both arms)."""
    from exoplanet_b200 import synthetic

    t = synthetic.cadence_times(N_CADENCES)
    p = synthetic.transit_params(C, 1, seed=synthetic.SEED + 17 * seed_offset, ecc=0.5)
    truth = synthetic.transit_params(1, 1, seed=synthetic.SEED - 1, ecc=0.5)
    return t, p, truth


def cpu_port_setup(t, p, truth):
    """Element blocks / data through the CPU oracle (used ONLY by the baseline legs)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import xo_oracle as O

    el_truth = O.elements_from_params(truth["period"], truth["t0"], truth["ecc"], truth["omega"], truth["b"], truth["ror"])
    cl_truth = O.get_cl(truth["u1"], truth["u2"])
    dtk, wk = O.stencil(None)
    from exoplanet_b200 import synthetic

    y = O.transit_lc(t, el_truth, cl_truth, dtk, wk)[0, :, 0] + synthetic.noise(N_CADENCES, SIGMA)
    return O, y, dtk, wk


def _cpu_quota():
    """CPUs the cgroup lets this container use (CFS quota / period), or None when unlimited / unknown"""
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        return None if q == "max" else float(q) / float(per)
    except (OSError, ValueError):
        pass
    try:
        q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        per = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        return None if q <= 0 else q / per
    except (OSError, ValueError):
        return None


def _host_state(load_before=None):
    """what else competes for the host's cores while the CPU arm is timed (the GPU boxes are shared): load averages, the
    cgroup's CPU quota and the CPUs this process may run on"""
    out = {"loadavg_after": [round(x, 1) for x in os.getloadavg()], "cpu_quota": _cpu_quota()}
    if load_before is not None:
        out["loadavg_before"] = [round(x, 1) for x in load_before]
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            out["cgroup_cpu_quota"] = open(path).read().strip()
            break
        except OSError:
            pass
    return out


def arm_reference(args):
    """The reference's own CPU implementation of the path is exoplanet-core (not in this image) -> the
    CPU restatement (oracle port) on all host cores.  One step = `inner` passes over the same 128-chain batch the
    GPU arm evaluates per step (inner is calibrated in the warm-up so that a step lasts >= ~0.1 s: 20 steps are
    >= 2 s of timed work); every pass derives the element blocks + contact points in C and evaluates logp + the full
    gradient, OpenMP over chains (dynamic schedule), threads pinned (OMP_PROC_BIND=close, OMP_PLACES=cores, set in
    main() before libgomp loads).  value = evaluations / MEDIAN step time; mean and best are reported beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tier = probe_reference_tier()
    load0 = os.getloadavg()
    cpus_allowed = len(os.sched_getaffinity(0))  # before libgomp pins this thread to one core
    nc = args.chains
    t, p, truth = build_problem(nc)
    O, y, dtk, wk = cpu_port_setup(t, p, truth)
    threads = O.max_threads()
    O.set_threads(threads)
    ivar = np.full(t.size, 1.0 / SIGMA ** 2)
    cl = O.get_cl(p["u1"], p["u2"])

    def one_pass(sel=slice(None), use_in_transit=True):
        # prologue (contact points, elements) + logp + grad, as one reference evaluation would
        e2 = O.elements_from_params_c(p["period"][sel], p["t0"][sel], p["ecc"][sel], p["omega"][sel], p["b"][sel], p["ror"][sel],
                                      m_planet=p["m_planet"][sel], use_in_transit=use_in_transit)
        return O.transit_logp_grad(t, y, ivar, e2, cl[sel], p["mean"][sel], dtk, wk)

    one_pass()
    c0 = time.perf_counter()
    one_pass()
    dt1 = time.perf_counter() - c0
    inner = max(1, int(np.ceil(0.1 / max(dt1, 1e-6))))

    def step():
        for _ in range(inner):
            one_pass()

    for _ in range(args.warmup):
        step()
    per = []
    for _ in range(args.steps):
        s0 = time.perf_counter()
        step()
        per.append(time.perf_counter() - s0)
    per = np.array(per)
    med, best, mean = float(np.median(per)), float(per.min()), float(per.mean())
    val = nc * inner / med
    # BASELINE.md section 4 legs: one thread; and use_in_transit=False (every cadence solved) on all threads
    O.set_threads(1)
    k1 = min(nc, 8)
    one_pass(slice(0, k1))
    b1 = []
    for _ in range(3):
        s0 = time.perf_counter()
        one_pass(slice(0, k1))
        b1.append(time.perf_counter() - s0)
    O.set_threads(threads)
    ka = min(nc, max(threads, 1))
    ba = []
    for _ in range(2):
        s0 = time.perf_counter()
        one_pass(slice(0, ka), use_in_transit=False)
        ba.append(time.perf_counter() - s0)
    line = {
        "impl": "reference", "metric": "model+grad evals/sec, N=1e5-cadence e=0.5 transit", "value": val, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * med,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(nc, args.gpus),
        "cpu_baseline": {"value": val, "unit": "evals/s", "cores": threads, "kind": "port",
                         "sample": "%d passes over %d chains x N=1e5 per step (OpenMP over chains x 8 cadence blocks), %d steps, %.2f s timed; "
                                   "value from the median step" % (inner, nc, args.steps, float(per.sum())),
                         "value_mean": nc * inner / mean, "value_best": nc * inner / best,
                         "step_time_spread": float((per.max() - per.min()) / med),
                         "one_thread": {"value": k1 / min(b1), "unit": "evals/s", "sample": "%d chains, best of 3" % k1},
                         "use_in_transit_false": {"value": ka / min(ba), "unit": "evals/s", "cores": threads,
                                                  "sample": "%d chains, every cadence solved, best of 2" % ka},
                         "omp": {k: os.environ.get(k) for k in ("OMP_PROC_BIND", "OMP_PLACES", "OMP_NUM_THREADS")},
                         "host_cpus": os.cpu_count(), "host": dict(_host_state(load0), cpus_allowed=cpus_allowed)},
        "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_tier": tier,
        "note": "CPU restatement of the reference mathematics (oracle/xo_oracle.c); exoplanet-core/PyTensor are not installable here",
    }
    print(json.dumps(line))


def _time_steps(fn, steps, flush):
    import torch

    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        ev.append((a, b))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / steps


def extra_configs(lib, dev, args):
    """configs[2]: 5-planet joint transit + RV model, N=1e5, batched VJP through all ops;
    configs[3]: celerite SHOTerm GP log-likelihood, N=1e6."""
    import torch

    from exoplanet_b200 import fused, gp, synthetic

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import xo_oracle as O

    O._skip_cpu = bool(args.no_cpu)  # no CPU timing legs inside the extras either

    def d(x):
        return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=dev)

    flush = torch.empty(192 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    out = {}
    only = set(args.extra.split(",")) if args.extra else None
    want = lambda name: only is None or name in only  # noqa: E731
    N = N_CADENCES
    t = synthetic.cadence_times(N)
    y = synthetic.noise(N, SIGMA)
    if want("c0"):
        _extra_config0(lib, dev, d, flush, out, O)
    if want("c2"):
        _extra_config2(lib, dev, d, flush, out, O, t, y)
    if want("gp"):
        _extra_gp(lib, dev, d, flush, out, O)
    if want("c4"):
        _extra_1024(lib, dev, d, flush, out, O, t, y)
    if want("hmc"):
        _extra_hmc(lib, dev, d, flush, out, O, t, y)
    return out


def _extra_config0(lib, dev, d, flush, out, O):
    """configs[0] (BASELINE: PyMC logp+dlogp on the reference's PyTensor ops, N=1e3, circular) cannot be executed in
    this image (no PyTensor / exoplanet-core; probe_reference_tier() says so in `run.reference_tier`).  SURVEY.md 8(d)
    "Config 1 caveat": the same problem through the CPU restatement stands in for it -- circular orbits bypass the Kepler
    solver (keplerian.py:331-332); quad_solution_vector and the closed-form in_transit branch (:733-741) are exercised."""
    import time

    import torch

    from exoplanet_b200 import fused, synthetic

    N = 1000
    t = np.arange(N, dtype=np.float64) * 0.02
    res = {}
    entry = {"unit": "model+grad evals/s", "cadences": N, "note": "SUBSTITUTE for configs[0]: the reference's PyMC/PyTensor path is not runnable "
             "here; CPU figure = the oracle port on the same N=1e3 circular problem, one thread per evaluation as the reference's ops are"}
    for Cn in (1, 128, 4096):
        p = synthetic.transit_params(Cn, 1, seed=synthetic.SEED + 3, ecc=0.0)
        p["period"] = np.clip(p["period"], 3.0, 15.0)  # at least one transit inside the 20-day series
        truth = {k: (v[:1] if hasattr(v, "shape") else v) for k, v in p.items()}
        el_t = O.elements_from_params(truth["period"], truth["t0"], truth["ecc"], truth["omega"], truth["b"], truth["ror"])
        dtk, wk = O.stencil(None)
        y = O.transit_lc(t, el_t, O.get_cl(truth["u1"], truth["u2"]), dtk, wk)[0, :, 0] + synthetic.noise(N, SIGMA)
        data = fused.TransitData(t, y, np.full(N, SIGMA), device=dev)
        model = fused.TransitModel(data)
        par = fused.pack_params(d(p["period"]), d(p["t0"]), d(p["ecc"]), d(p["omega"]), d(p["b"]), d(p["ror"]), d(p["m_planet"]), d(p["K"]))
        one = torch.ones(Cn, dtype=torch.float64, device=dev)
        star = fused.pack_star(one, one, d(p["u1"]), d(p["u2"]))
        mean = d(p["mean"])

        def step():
            res["o"] = model.logp_grad(par, star, mean)

        ms = _time_steps(step, 20, flush)
        sel = np.unique(np.linspace(0, Cn - 1, 8).astype(int))
        el_o = O.elements_from_params(p["period"][sel], p["t0"][sel], p["ecc"][sel], p["omega"][sel], p["b"][sel], p["ror"][sel],
                                      m_planet=p["m_planet"][sel])
        args_o = (t, y, np.full(N, 1 / SIGMA ** 2), el_o, O.get_cl(p["u1"][sel], p["u2"][sel]), p["mean"][sel], np.zeros(1), np.ones(1))
        lo, geo, _, gmo = O.transit_logp_grad(*args_o, lognorm=data.lognorm)
        entry["%d chains" % Cn] = {
            "value": Cn / (ms * 1e-3), "ms_per_step": ms,
            "parity_logp_rel_err": float(np.max(np.abs(res["o"][0].cpu().numpy()[sel] - lo) / np.abs(lo))),
            "parity_gmean_rel_err": float(np.max(np.abs(res["o"][3].cpu().numpy()[sel] - gmo) / np.abs(gmo).max()))}
        if Cn == 128 and not getattr(O, "_skip_cpu", False):
            O.set_threads(1)
            reps, t0 = 0, time.perf_counter()
            while time.perf_counter() - t0 < 2.0:
                O.transit_logp_grad(*args_o, lognorm=data.lognorm)
                reps += 1
            dt = time.perf_counter() - t0
            O.set_threads(os.cpu_count() or 1)
            entry["cpu_port"] = {"cores": 1, "value": reps * len(sel) / dt, "unit": "model+grad evals/s", "kind": "port",
                                 "sample": "%d passes over %d chains, %.1f s" % (reps, len(sel), dt)}
    entry["value"] = entry["128 chains"]["value"]
    out["configs[0] substitute: circular orbit, N=1e3 (Kepler solver bypassed)"] = entry


def _extra_config2(lib, dev, d, flush, out, O, t, y):
    import torch

    from exoplanet_b200 import fused, synthetic

    # ---- configs[2] -----------------------------------------------------------------------------
    Cn, P, N = 128, 5, N_CADENCES
    p = synthetic.transit_params(Cn, P, seed=synthetic.SEED + 5)
    data = fused.TransitData(t, y, np.full(N, SIGMA), device=dev)
    rv_obs = 3.0 * np.random.default_rng(7).normal(size=N)
    rvdata = fused.TransitData(t, rv_obs, np.full(N, 3.0), device=dev)
    par = fused.pack_params(d(p["period"]), d(p["t0"]), d(p["ecc"]), d(p["omega"]), d(p["b"]), d(p["ror"]), d(p["m_planet"]), d(p["K"]))
    star = fused.pack_star(torch.ones(Cn, dtype=torch.float64, device=dev), torch.ones(Cn, dtype=torch.float64, device=dev), d(p["u1"]),
                           d(p["u2"]))
    mean, gamma = d(p["mean"]), d(p["gamma"])
    model = fused.TransitModel(data)
    el_o = O.elements_from_params(p["period"], p["t0"], p["ecc"], p["omega"], p["b"], p["ror"], m_planet=p["m_planet"], K=p["K"])
    elem = d(el_o)
    res = {}

    def step():
        res["lc"] = model.logp_grad(par, star, mean)
        res["rv"] = fused.rv_logp_grad(rvdata, elem, gamma)

    ms = _time_steps(step, 10, flush)
    sel = [0, Cn - 1]
    lo = O.transit_logp_grad(t, y, np.full(N, 1 / SIGMA ** 2), el_o[sel], O.get_cl(p["u1"][sel], p["u2"][sel]), p["mean"][sel], np.zeros(1),
                             np.ones(1), lognorm=data.lognorm)[0]
    lr, grvo, _ = O.rv_logp_grad(t, rv_obs, np.full(N, 1 / 9.0), el_o[sel], p["gamma"][sel], lognorm=rvdata.lognorm)
    err = max(float(np.max(np.abs(res["lc"][0].cpu().numpy()[sel] - lo) / np.abs(lo))),
              float(np.max(np.abs(res["rv"][0].cpu().numpy()[sel] - lr) / np.abs(lr))))
    gerr = float(np.max(np.abs(res["rv"][1].cpu().numpy()[sel] - grvo) / (np.abs(grvo).max(axis=(0, 1), keepdims=True) + 1e-300)))
    out["configs[2] 5-planet transit+RV, N=1e5"] = {"value": Cn / (ms * 1e-3), "unit": "model+grad evals/s", "ms_per_step": ms, "chains": Cn,
                                                    "parity_logp_rel_err": err, "parity_rv_grad_rel_err": gerr,
                                                    "parity_note": "2 sampled chains here; all gradients of 8 chains in tests/test_gpu_configs.py"}


def _extra_gp(lib, dev, d, flush, out, O):
    import torch

    from exoplanet_b200 import gp, synthetic

    # ---- configs[3] -----------------------------------------------------------------------------
    Ng = 1000000
    pr = synthetic.gp_problem(Ng)
    tg, yg, dg = d(pr["t"]), d(pr["y"]), d(pr["diag"])
    lo = O.celerite_sho_loglike(pr["t"], pr["y"], pr["diag"], pr["S0"], pr["w0"], pr["Q"])
    for B in (1, 16):
        yb = yg[None, :].repeat(B, 1).contiguous()
        ws = torch.empty(lib.xo_celerite_sho_workspace(Ng, B) // 8 + 1, dtype=torch.float64, device=dev)
        r = {}

        hyp = [d(np.full(B, pr[k_])) for k_ in ("S0", "w0", "Q")]  # device tensors: no per-call scalar uploads in the timed region

        def gstep():
            r["ll"] = gp.sho_loglike(tg, yb, dg, *hyp, workspace=ws)

        ms = _time_steps(gstep, 10, flush)
        out["configs[3] celerite SHO loglike, N=1e6, %d chain(s)" % B] = {
            "value": B * Ng / (ms * 1e-3), "unit": "cadences/s", "ms_per_step": ms, "gb_per_s": 24.0 * B * Ng / (ms * 1e-3) / 1e9,
            "parity_rel_err": float(abs(r["ll"][0].item() - lo) / abs(lo))}
        del ws
        wg = torch.empty(lib.xo_celerite_sho_grad_workspace(Ng, B) // 8 + 1, dtype=torch.float64, device=dev)

        def ggrad():
            r["g"] = gp.sho_grad(tg, yb, dg, *hyp, workspace=wg)

        ms = _time_steps(ggrad, 10, flush)
        out["configs[3] celerite SHO loglike + exact gradient (S0, w0, Q, y, diag), N=1e6, %d chain(s)" % B] = {
            "value": B * Ng / (ms * 1e-3), "unit": "cadences/s", "ms_per_step": ms,
            "loglike_rel_err": float(abs(r["g"][0][0].item() - lo) / abs(lo))}
        del wg
    # gradient parity of the 16-chain result (every chain sees the same data) against the O(N) C adjoint
    _, gho, gyo, gdo = O.celerite_sho_grad_c(pr["t"], pr["y"], pr["diag"], pr["S0"], pr["w0"], pr["Q"])
    gh, gy_, gd_ = r["g"][1][-1].cpu().numpy(), r["g"][2][-1].cpu().numpy(), r["g"][3][-1].cpu().numpy()
    out["configs[3] celerite SHO loglike + exact gradient (S0, w0, Q, y, diag), N=1e6, 16 chain(s)"].update(
        parity_hyper_rel_err=float(np.max(np.abs(gh - gho) / np.abs(gho))), parity_gy_rel_err=float(np.max(np.abs(gy_ - gyo)) / np.abs(gyo).max()),
        parity_gdiag_rel_err=float(np.max(np.abs(gd_ - gdo)) / np.abs(gdo).max()))
    # joint model of configs[3]: transit mean + SHO GP noise at N=1e6: logL and its gradient w.r.t. every parameter through
    # ONE C call (xo_transit_gp_logp_grad: prologue -> residual -> GP scan + adjoint -> light-curve VJP -> epilogue)
    from exoplanet_b200 import fused

    for Cj in (1, 16):
        pj = synthetic.transit_params(Cj, 1, seed=synthetic.SEED + 21, ecc_max=0.4)
        par = fused.pack_params(d(pj["period"]), d(pj["t0"]), d(pj["ecc"]), d(pj["omega"]), d(pj["b"]), d(pj["ror"]))
        one = torch.ones(Cj, dtype=torch.float64, device=dev)
        star = fused.pack_star(one, one, d(pj["u1"]), d(pj["u2"]))
        hyp = [d(np.full(Cj, pr[k])) for k in ("S0", "w0", "Q")]
        jm = fused.TransitGPModel(pr["t"], pr["y"], pr["diag"], device=dev)
        res = {}

        def jstep():
            res["o"] = jm.logp_grad(par, star, d(pj["mean"]), *hyp)

        ms = _time_steps(jstep, 10, flush)
        # parity of chain 0: the oracle's light curve -> residual -> sequential celerite adjoint
        el0 = O.elements_from_params(pj["period"][:1], pj["t0"][:1], pj["ecc"][:1], pj["omega"][:1], pj["b"][:1], pj["ror"][:1])
        mu0 = O.transit_lc(pr["t"], el0, O.get_cl(pj["u1"][:1], pj["u2"][:1]), np.zeros(1), np.ones(1))[0, :, 0]
        lo0, gh0, gy0, _ = O.celerite_sho_grad_c(pr["t"], pr["y"] - pj["mean"][0] - mu0, pr["diag"], pr["S0"], pr["w0"], pr["Q"])
        o = res["o"]
        out["configs[3] transit + SHO GP, N=1e6, %d chain(s): logL + d/d(all orbital, stellar, mean, S0, w0, Q) in one C call" % Cj] = {
            "value": Cj / (ms * 1e-3), "unit": "model+grad evals/s", "ms_per_step": ms, "chains": Cj, "status_word": list(jm.status()),
            "parity_loglike_rel_err": float(abs(o["loglike"][0].item() - lo0) / abs(lo0)),
            "parity_hyper_rel_err": float(np.max(np.abs(o["ghyper"][0].cpu().numpy() - gh0) / np.abs(gh0))),
            "parity_gmean_rel_err": float(abs(o["gmean"][0].item() + gy0.sum()) / np.abs(gy0).sum())}


def _extra_1024(lib, dev, d, flush, out, O, t, y):
    import torch

    from exoplanet_b200 import fused, synthetic

    N = N_CADENCES
    # ---- configs[4]'s chain count on ONE GPU: 1024 chains per step (the fixed ~30 us of prologue,
    # epilogue and launches is amortised over 8x the work of the headline step)
    pb = synthetic.transit_params(1024, 1, seed=synthetic.SEED + 11, ecc=0.5)
    parb = fused.pack_params(d(pb["period"]), d(pb["t0"]), d(pb["ecc"]), d(pb["omega"]), d(pb["b"]), d(pb["ror"]), d(pb["m_planet"]), d(pb["K"]))
    one = torch.ones(1024, dtype=torch.float64, device=dev)
    starb = fused.pack_star(one, one, d(pb["u1"]), d(pb["u2"]))
    meanb = d(pb["mean"])
    big = fused.TransitModel(fused.TransitData(t, y, np.full(N, SIGMA), device=dev))
    res = {}

    def bstep():
        res["o"] = big.logp_grad(parb, starb, meanb)

    msb = _time_steps(bstep, 10, flush)
    sel = np.linspace(0, 1023, 8).astype(int)
    el_o = O.elements_from_params(pb["period"][sel], pb["t0"][sel], pb["ecc"][sel], pb["omega"][sel], pb["b"][sel], pb["ror"][sel],
                                  m_planet=pb["m_planet"][sel])
    lo, _, _, gmo = O.transit_logp_grad(t, y, np.full(N, 1 / SIGMA ** 2), el_o, O.get_cl(pb["u1"][sel], pb["u2"][sel]), pb["mean"][sel], np.zeros(1),
                                        np.ones(1), lognorm=big.data.lognorm)
    out["configs[1] with 1024 chains on one GPU (N=1e5, e=0.5)"] = {
        "value": 1024 / (msb * 1e-3), "unit": "model+grad evals/s", "ms_per_step": msb, "chains": 1024,
        "parity_logp_rel_err": float(np.max(np.abs(res["o"][0].cpu().numpy()[sel] - lo) / np.abs(lo))),
        "parity_gmean_rel_err": float(np.max(np.abs(res["o"][3].cpu().numpy()[sel] - gmo) / np.abs(gmo).max())),
        "parity_note": "8 sampled chains against the oracle; every parameter gradient of 16 chains in tests/test_gpu_configs.py"}


def _extra_hmc(lib, dev, d, flush, out, O, t, y):
    import torch

    from exoplanet_b200 import fused, synthetic

    N = N_CADENCES
    # ---- the lock-step HMC driver on configs[1] (128 chains x N=1e5), one CUDA graph per transition
    from exoplanet_b200.sampling import BatchedHMC, TransitPosterior

    p1 = synthetic.transit_params(128, 1, seed=synthetic.SEED + 9, ecc=0.5)
    q_init = TransitPosterior.pack(p1["period"], p1["t0"], p1["ecc"], p1["omega"], p1["b"], p1["ror"], p1["u1"], p1["u2"], p1["mean"]).to(dev)
    sig = torch.tensor([1e-3, 0.02, 0.3, 0.3, 0.3, 0.3, 1.5, 1.5, 1e-3], dtype=torch.float64, device=dev)
    post = TransitPosterior(fused.TransitModel(fused.TransitData(t, y, np.full(N, SIGMA), device=dev)), 1, q_init, sig)
    for mode in ("one C call per transition (xo_hmc_transition)", "CUDA graph replay", "eager launches"):
        use_graph = mode == "CUDA graph replay"
        hmc = BatchedHMC(post, q_init, step_size=1e-6, n_leapfrog=8, seed=1, use_graph=use_graph, fused=None if mode.startswith("one C") else False)
        for _ in range(3):
            hmc.transition()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            hmc.transition()
        b.record()
        torch.cuda.synchronize()
        msj = a.elapsed_time(b) / 20
        out["HMC driver, configs[1] x 128 chains, 8 leapfrog steps per transition, %s" % mode] = {
            "value": 128 * 8 / (msj * 1e-3), "unit": "leapfrog model+grad evals/s (incl. parameter transforms, their chain rule and the priors)",
            "ms_per_transition": msj, "chain_transitions_per_s": 128 / (msj * 1e-3), "accept_mean": float(hmc.accept_prob.mean().item())}
    # ---- the NUTS driver on the same posterior: doublings as CUDA graphs against one launch per tensor op
    from exoplanet_b200.sampling import BatchedNUTS

    for mode, use_graph, fused_tree in (("tree kernels (xo_nuts_*), one CUDA graph per doubling", True, True),
                                        ("tensor operations, one CUDA graph per doubling", True, False),
                                        ("tensor operations, eager launches", False, False)):
        nuts = BatchedNUTS(post, q_init, step_size=1e-6, max_depth=4, seed=2, use_graph=use_graph, fused_tree=fused_tree)
        for _ in range(2):
            nuts.transition()
        torch.cuda.synchronize()
        n0 = nuts.n_leapfrog
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            nuts.transition()
        b.record()
        torch.cuda.synchronize()
        leaves = nuts.n_leapfrog - n0
        msl = a.elapsed_time(b) / max(leaves, 1)
        out["NUTS driver, configs[1] x 128 chains, max_depth 4, %s" % mode] = {
            "value": 128 / (msl * 1e-3), "unit": "leapfrog model+grad evals/s (tree building, U-turn checks and multinomial sampling included)",
            "ms_per_leapfrog": msl, "leapfrogs_per_transition": leaves / 5, "depth_mean": float(nuts.depth.to(torch.float64).mean().item())}


def arm_ours(args):
    import torch

    from exoplanet_b200 import _lib, elements, fused, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # one process per GPU, bound to the GPU's NUMA node before any pinned buffer is allocated
    from exoplanet_b200 import parallel as xo_parallel

    bound_cpus, prev_affinity = xo_parallel.bind_to_device_numa(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        # keep stdout to the one JSON line: NCCL prints its version banner to stdout at NCCL_DEBUG=VERSION/WARN
        # (and honours NCCL_DEBUG_FILE only above VERSION), so raise the level to WARN, point the log at
        # stderr, and park fd 1 on stderr while the communicator is created
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    lib = _lib.require_cuda()

    C_loc = args.chains
    t_np, p, truth = build_problem(C_loc, seed_offset=0 if args.same_seed else rank)

    def d(x):
        return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=dev)

    # data: truth light curve (GPU kernels) + seeded noise, resident in HBM
    el_truth = elements.elements_from_params(d(truth["period"]), d(truth["t0"]), d(truth["ecc"]), d(truth["omega"]), d(truth["b"]),
                                             d(truth["ror"]))
    cl_truth = fused.limb_dark_coeffs(d(truth["u1"]), d(truth["u2"]))
    t_dev = d(t_np)
    y_np = fused.light_curve(t_dev, el_truth, cl_truth)[0, :, 0].cpu().numpy() + synthetic.noise(N_CADENCES, SIGMA)
    data = fused.TransitData(t_np, y_np, np.full(N_CADENCES, SIGMA), device=dev)
    # sampler-side parameters resident in HBM: par[C][1][8], star[C][4], mean[C]
    par = fused.pack_params(d(p["period"]), d(p["t0"]), d(p["ecc"]), d(p["omega"]), d(p["b"]), d(p["ror"]), d(p["m_planet"]), d(p["K"]))
    star = fused.pack_star(torch.ones(C_loc, dtype=torch.float64, device=dev), torch.ones(C_loc, dtype=torch.float64, device=dev),
                           d(p["u1"]), d(p["u2"]))
    mean = d(p["mean"])
    stencil = fused.exposure_stencil(None)
    dtk, wk = d(stencil[0]), d(stencil[1])
    P = 1
    ws_bytes = lib.xo_transit_model_workspace(N_CADENCES, C_loc, P)
    ws = torch.empty(ws_bytes // 8 + 1, dtype=torch.float64, device=dev)
    logp = torch.empty(C_loc, dtype=torch.float64, device=dev)
    gpar = torch.empty_like(par)
    gstar = torch.empty_like(star)
    gmean = torch.empty(C_loc, dtype=torch.float64, device=dev)
    ptr = lambda x: C.c_void_p(x.data_ptr())
    stream = torch.cuda.current_stream()

    # data-only moments of (y, ivar), computed once per data set like the upload itself
    mom_ptr = None if args.full_sweep else ptr(data.moments())

    def step_device():
        # prologue (elements, contact points, limb-darkening vector) -> fused logp+grad -> reverse prologue
        rc = lib.xo_transit_model_logp_grad(ptr(data.t), ptr(data.y), ptr(data.ivar), N_CADENCES, ptr(par), ptr(star), ptr(mean), C_loc, P,
                                            ptr(dtk), ptr(wk), 1, 0.0, 1, data.lognorm, mom_ptr, ptr(logp), ptr(gpar), ptr(gstar),
                                            ptr(gmean), ptr(ws), ws.numel() * 8, C.c_void_p(stream.cuda_stream))
        if rc != 0:
            raise RuntimeError(lib.xo_strerror(rc).decode())

    flush = torch.empty(192 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    # measured fp64 ceiling of THIS box (register-resident DFMA loop on every SM, CUDA events on the launch stream)
    dfma = C.c_double(0.0)
    dfma_ms = C.c_double(0.0)
    rc = lib.xo_diag_dfma_peak(5, C.byref(dfma), C.byref(dfma_ms), C.c_void_p(stream.cuda_stream))
    if rc != 0:
        raise RuntimeError(lib.xo_strerror(rc).decode())
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    launches0 = lib.xo_diag_launch_count()
    for a, b in ev:
        flush.zero_()  # L2 flush between timed iterations (outside the events)
        a.record()
        step_device()
        b.record()
    barrier()
    gpu_launches = lib.xo_diag_launch_count() - launches0  # counted by the library at every launch site
    wall = time.perf_counter() - wall0
    ms_dev = sum(a.elapsed_time(b) for a, b in ev)
    # the same step with the Gaussian terms of EVERY cadence summed in the sweep (what the CPU arm does): no data moments
    saved_mom, mom_ptr = mom_ptr, None
    for _ in range(3):
        step_device()
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev2:
        flush.zero_()
        a.record()
        step_device()
        b.record()
    barrier()
    ms_full = sum(a.elapsed_time(b) for a, b in ev2)
    mom_ptr = saved_mom
    # ---- e2e: PINNED host buffers through the C ABI; everything (t, y, ivar, params) goes
    # host->device and (logp, grads) device->host inside the timed region, every step -----------
    def pinned(a):
        return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).pin_memory()

    # t, y, ivar as consecutive slices of ONE pinned block: the host entry point then uploads them with one DMA
    lc_block = pinned(np.concatenate([t_np, y_np, np.full(N_CADENCES, 1 / SIGMA ** 2)]))
    h = dict(t=lc_block[:N_CADENCES], y=lc_block[N_CADENCES:2 * N_CADENCES], ivar=lc_block[2 * N_CADENCES:], par=pinned(par.cpu().numpy()),
             star=pinned(star.cpu().numpy()), mean=pinned(p["mean"]), dtk=pinned(stencil[0]), wk=pinned(stencil[1]))
    o = dict(logp=pinned(np.empty(C_loc)), gpar=pinned(np.empty((C_loc, P, 8))), gstar=pinned(np.empty((C_loc, 4))),
             gmean=pinned(np.empty(C_loc)))

    def step_host():
        rc = lib.xo_transit_model_logp_grad_host(ptr(h["t"]), ptr(h["y"]), ptr(h["ivar"]), N_CADENCES, ptr(h["par"]), ptr(h["star"]),
                                                 ptr(h["mean"]), C_loc, P, ptr(h["dtk"]), ptr(h["wk"]), 1, 0.0, 1, data.lognorm,
                                                 ptr(o["logp"]), ptr(o["gpar"]), ptr(o["gstar"]), ptr(o["gmean"]), local)
        if rc != 0:
            raise RuntimeError(lib.xo_strerror(rc).decode())

    # warm-up: >= W calls AND >= 0.15 s of sustained host<->device traffic -- the PCIe link / copy
    # engines reach full speed only after tens of ms of activity (tools/numa_probe.py: the same
    # 3 x 0.8 MB copy takes 185 us cold, 65 us warm), and an MCMC run is sustained traffic
    def pcie_link():
        """'gen, width, gen max' of the GPU's PCIe link right now (it is renegotiated down when idle)"""
        try:
            q = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max",
                                "--format=csv,noheader"], capture_output=True, text=True, timeout=10).stdout.strip()
            return q
        except Exception:
            return None

    def link_below_max(q):
        try:
            gen, _, gmax = [int(x) for x in q.split(",")]
            return gen < gmax
        except Exception:
            return False

    link_before = pcie_link()
    # ... and until the per-call time has settled (the median of the last 20 calls within 5 % of the 20 before), at most 2 s
    w0, nw, hist = time.perf_counter(), 0, []
    while True:
        s0 = time.perf_counter()
        step_host()
        hist.append(time.perf_counter() - s0)
        nw += 1
        el = time.perf_counter() - w0
        if nw >= max(3, args.warmup) and el >= 0.4:
            settled = len(hist) >= 40 and abs(np.median(hist[-20:]) / np.median(hist[-40:-20]) - 1.0) < 0.05
            if settled or el > 2.0:
                break
    # ... and until the PCIe link runs at its full generation (an idle GPU's link drops to a lower speed and comes back
    # under traffic; a sampler keeps it busy for minutes): keep calling, look again every 200 calls, at most 3 s more
    link_mid, extra = pcie_link(), 0
    l0 = time.perf_counter()
    while link_below_max(link_mid) and time.perf_counter() - l0 < 3.0:
        for _ in range(200):
            s0 = time.perf_counter()
            step_host()
            hist.append(time.perf_counter() - s0)
        extra += 200
        link_mid = pcie_link()
    e2e_warm = {"calls": nw + extra, "seconds": time.perf_counter() - w0, "first_calls_us": [round(x * 1e6) for x in hist[:3]],
                "last_calls_us_median": round(float(np.median(hist[-20:])) * 1e6), "pcie_link_before": link_before,
                "pcie_link_after_warmup": link_mid, "calls_waiting_for_the_link": extra}
    barrier()
    # K synchronous calls, timed as a block; blocks are repeated for at least 1 s (and at least 5 of them) and the MEDIAN
    # block is reported: the host<->device path of a fresh process has slow plateaus of a few hundred milliseconds (the same
    # call takes 105 or 250 us; tools/e2e_repeat.sh, profiles/r02_e2e_repeat.txt), longer than 5 blocks of 20 calls
    e2e_blocks = []
    e2e_t0 = time.perf_counter()
    while len(e2e_blocks) < 5 or (time.perf_counter() - e2e_t0 < 1.0 and len(e2e_blocks) < 2000):
        e2e_steps = []
        e0 = time.perf_counter()
        for _ in range(args.steps):
            s0 = time.perf_counter()
            step_host()  # synchronous: returns when logp and the gradients are in host memory
            e2e_steps.append(time.perf_counter() - s0)
        torch.cuda.synchronize()
        e2e_blocks.append(time.perf_counter() - e0)
        if os.environ.get("XO_BENCH_VERBOSE"):
            sys.stderr.write("e2e per-step us: " + " ".join("%.0f" % (x * 1e6) for x in e2e_steps) + "\n")
    e2e_s = float(np.median(e2e_blocks))
    # ---- e2e with the light curve resident (uploaded once, as a compiled model holds its
    # constants): per step only the parameters go in and logp + gradients come back ----------
    handle = C.c_void_p()
    rc = lib.xo_transit_data_create(ptr(h["t"]), ptr(h["y"]), ptr(h["ivar"]), N_CADENCES, local, C.byref(handle))
    if rc != 0:
        raise RuntimeError(lib.xo_strerror(rc).decode())

    def step_resident():
        rc = lib.xo_transit_model_logp_grad_resident(handle, ptr(h["par"]), ptr(h["star"]), ptr(h["mean"]), C_loc, P, ptr(h["dtk"]),
                                                     ptr(h["wk"]), 1, 0.0, 1, data.lognorm, 0 if args.full_sweep else 1, ptr(o["logp"]),
                                                     ptr(o["gpar"]), ptr(o["gstar"]), ptr(o["gmean"]))
        if rc != 0:
            raise RuntimeError(lib.xo_strerror(rc).decode())

    w0, nw = time.perf_counter(), 0
    while nw < max(3, args.warmup) or time.perf_counter() - w0 < 0.05:
        step_resident()
        nw += 1
    barrier()
    res_blocks = []
    for _rep in range(5):
        res_steps = []
        r0 = time.perf_counter()
        for _ in range(args.steps):
            s0 = time.perf_counter()
            step_resident()
            res_steps.append(time.perf_counter() - s0)
        res_blocks.append(time.perf_counter() - r0)
        if os.environ.get("XO_BENCH_VERBOSE"):
            sys.stderr.write("resident per-step us: " + " ".join("%.0f" % (x * 1e6) for x in res_steps) + "\n")
    res_s = float(np.median(res_blocks))
    lib.xo_transit_data_destroy(handle)
    clocks = sampler.stop() if rank == 0 else None
    h2d = sum(h[k].numel() * 8 for k in h)
    d2h = sum(o[k].numel() * 8 for k in o)

    # ---- Kepler micro-benchmark (rank 0 numbers reported; > L2 working set) ---------------
    kep = None
    if not args.no_kepler:
        n = 1 << 26
        g = torch.Generator(device=dev).manual_seed(1234)
        M = (torch.rand(n, dtype=torch.float64, device=dev, generator=g) * 2 - 1) * (2000 * np.pi)
        kep = {}
        for name, e in (("e=0.5", torch.full((n,), 0.5, dtype=torch.float64, device=dev)),
                        ("e~U(0,0.95)", torch.rand(n, dtype=torch.float64, device=dev, generator=g) * 0.95)):
            sf = torch.empty_like(M)
            cf = torch.empty_like(M)
            for _ in range(3):
                lib.xo_kepler_fwd(ptr(M), ptr(e), ptr(sf), ptr(cf), n, C.c_void_p(stream.cuda_stream))
            torch.cuda.synchronize()
            evs = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                lib.xo_kepler_fwd(ptr(M), ptr(e), ptr(sf), ptr(cf), n, C.c_void_p(stream.cuda_stream))
                b.record()
                evs.append((a, b))
            torch.cuda.synchronize()
            ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
            kep[name] = {"cadences_per_s": n / (ms * 1e-3), "gb_per_s": n * KEPLER_BYTES_PER_ELEMENT / (ms * 1e-3) / 1e9,
                         "ms": ms, "n": n}
            del sf, cf, e
        # the other standalone ops at > L2 working sets (algorithmic bytes per element in BASELINE.md section 5)
        n2 = 1 << 24
        bb = torch.rand(n2, dtype=torch.float64, device=dev, generator=g) * 1.25
        rr = torch.rand(n2, dtype=torch.float64, device=dev, generator=g) * 0.18 + 0.02
        s3 = torch.empty((n2, 3), dtype=torch.float64, device=dev)
        db3, dr3 = torch.empty_like(s3), torch.empty_like(s3)
        e1 = torch.rand(n2, dtype=torch.float64, device=dev, generator=g) * 0.9
        a1, a2, o1, o2 = M[:n2].sin(), M[:n2].cos(), torch.empty(n2, dtype=torch.float64, device=dev), torch.empty(n2, dtype=torch.float64, device=dev)
        sp = C.c_void_p(stream.cuda_stream)
        micro = {
            "qsv_fwd+jac (88 B/pt)": (lambda: lib.xo_qsv_fwd(ptr(bb), ptr(rr), ptr(s3), ptr(db3), ptr(dr3), n2, sp), 88),
            "qsv_fwd (40 B/pt)": (lambda: lib.xo_qsv_fwd(ptr(bb), ptr(rr), ptr(s3), None, None, n2, sp), 40),
            "qsv_vjp (88 B/pt)": (lambda: lib.xo_qsv_vjp(ptr(s3), ptr(db3), ptr(dr3), ptr(o1), ptr(o2), n2, sp), 88),
            "kepler_vjp (56 B/elt)": (lambda: lib.xo_kepler_vjp(ptr(a1), ptr(a2), ptr(e1), ptr(bb), ptr(rr), ptr(o1), ptr(o2), n2, sp), 56),
        }
        for name, (fn, nbytes) in micro.items():
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            evs = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                evs.append((a, b))
            torch.cuda.synchronize()
            ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
            kep[name] = {"elements_per_s": n2 / (ms * 1e-3), "gb_per_s": n2 * nbytes / (ms * 1e-3) / 1e9, "ms": ms, "n": n2}
        del M, bb, rr, s3, db3, dr3, e1, a1, a2, o1, o2

    # ---- the other BASELINE configs, measured briefly on one GPU (parity-checked against the oracle) ----
    other = None
    if rank == 0 and world == 1 and not args.no_extra:
        other = extra_configs(lib, dev, args)

    tm = torch.tensor([ms_dev, e2e_s * 1e3, res_s * 1e3, ms_full], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms_dev, e2e_ms, res_ms, ms_full = tm[0].item(), tm[1].item(), tm[2].item(), tm[3].item()
    total_evals = C_loc * world * args.steps
    value = total_evals / (ms_dev * 1e-3)
    e2e = total_evals / (e2e_ms * 1e-3)
    peak, peak_src = _peaks()
    ach = ALG_BYTES_PER_CHAIN_CADENCE * C_loc * N_CADENCES / (ms_dev / args.steps * 1e-3) / 1e9  # per GPU, per launch

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        # parity gate before reporting: the timed GPU result against the oracle on a few chains
        O, y_o, dtk_o, wk_o = cpu_port_setup(t_np, p, truth)
        sel = [0, C_loc // 2, C_loc - 1]
        el_o = O.elements_from_params(p["period"][sel], p["t0"][sel], p["ecc"][sel], p["omega"][sel], p["b"][sel], p["ror"][sel],
                                      m_planet=p["m_planet"][sel])
        lo = O.transit_logp_grad(t_np, y_np, np.full(N_CADENCES, 1 / SIGMA ** 2), el_o, O.get_cl(p["u1"][sel], p["u2"][sel]),
                                 p["mean"][sel], dtk_o, wk_o, lognorm=data.lognorm)[0]
        rel = float(np.max(np.abs(logp.cpu().numpy()[sel] - lo) / np.abs(lo)))
        if rel > 1e-10:
            raise SystemExit("parity gate failed: logp rel err %g" % rel)
        # the CPU restatement, timed in a clean process on all host cores (~10-20 s)
        def all_cores():  # the child must not inherit this rank's NUMA binding
            if prev_affinity:
                os.sched_setaffinity(0, prev_affinity)

        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(args.steps), "--warmup",
                              str(args.warmup), "--chains", str(C_loc)], capture_output=True, text=True, timeout=900, preexec_fn=all_cores)
        try:
            ref = json.loads(out.stdout.strip().splitlines()[-1])
            cpu = ref["cpu_baseline"]
            cpu["parity_logp_rel_err"] = rel
            cpu["reference_tier"] = ref.get("reference_tier")
        except Exception:
            cpu = {"value": None, "unit": "evals/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: " + out.stderr[-200:]}

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    prof = _first_profile("r02_transit.txt", "r01_transit.txt")
    fp64_inst = _ncu_metric(prof, "smsp__inst_executed_pipe_fp64.sum")  # warp-level fp64-pipe instructions per launch
    fp64_pct = _ncu_metric(prof, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")
    step_s = ms_dev / args.steps * 1e-3
    peak_tf = 2.0 * dfma.value / 1e12
    if fp64_inst is not None and C_loc == CHAINS_PER_GPU:
        ach_tf = 2.0 * 32.0 * fp64_inst / step_s / 1e12
        how = ("fp64-pipe warp instructions per launch (ncu smsp__inst_executed_pipe_fp64.sum, profiles/%s; the workload is deterministic) "
               "x 32 lanes x 2 flop / the live CUDA-event duration of the step" % prof)
    else:
        ach_tf, how = None, "no fp64 instruction count committed for this chain count"
    line = {
        "metric": "model+grad evals/sec, N=1e5-cadence e=0.5 transit", "value": value, "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(C_loc, world),
        "run": {"out_of_transit_sums": "full sweep over y, ivar" if args.full_sweep else "data moments (t-only sweep)",
                "host_affinity": ("process bound to the %d CPUs of the GPU's NUMA node" % len(bound_cpus)) if bound_cpus else "unbound",
                "step": "params -> elements+contact points -> fused logp+grad -> reverse prologue, 3 kernel launches chained by programmatic "
                        "dependent launch (XO_MODEL_LAUNCHES=1|2 fuse them; slower), params resident in HBM",
                "seeds": "rank r draws its chains with seed SEED + 17 r (heterogeneous ranks); --same-seed gives every rank rank 0's chains"
                         if not args.same_seed else "every rank evaluates rank 0's chains (identical per-rank work)"},
        "value_full_sweep": {"value": total_evals / (ms_full * 1e-3), "unit": "evals/s", "ms_per_step": ms_full / args.steps,
                             "note": "same step with moments = NULL: y and ivar of every cadence are read and summed in the sweep, as the CPU arm does"},
        "e2e": {"value": e2e, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "xo_transit_model_logp_grad_host: t,y,ivar (one pinned block, one DMA), params from pinned host memory every call; "
                       "logp+grads back to host",
                "timing": "median of %d blocks of `steps` synchronous calls (>= 1 s of them)" % len(e2e_blocks),
                "blocks_quartiles_evals_per_s": [float(C_loc * args.steps / np.percentile(e2e_blocks, q)) for q in (75, 50, 25)], "warmup": e2e_warm, "pcie_link_after": pcie_link(),
                "blocks_evals_per_s": [float(C_loc * args.steps / b) for b in e2e_blocks[:5]]},
        "e2e_resident_data": {"value": total_evals / (res_ms * 1e-3), "unit": "evals/s",
                              "h2d_bytes_per_step": sum(h[k].numel() * 8 for k in ("par", "star", "mean", "dtk", "wk")), "d2h_bytes_per_step": d2h,
                              "api": "xo_transit_model_logp_grad_resident: light curve uploaded once (xo_transit_data_create); params in, logp+grads out per call"},
        "gpu_launches": int(gpu_launches),
        "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": (ach_tf / peak_tf) if ach_tf else None,
                     "traffic": _ncu_traffic(prof, "transit"), "traffic_unit": "bytes of DRAM per launch (ncu, profiles/%s)" % prof,
                     "kernel": "transit_logp_grad_kernel",
                     "peak_source": "measured in this run: xo_diag_dfma_peak (register-resident DFMA loop, %d CTAs x 256 threads, 8 chains per thread, "
                                    "best of 5, %.3f ms) = %.3e DFMA/s" % (4 * torch.cuda.get_device_properties(dev).multi_processor_count, dfma_ms.value, dfma.value),
                     "achieved_how": how, "ncu_pipe_fp64_pct_in_capture": fp64_pct,
                     "duration_ms": ms_dev / args.steps,
                     "duration_note": "CUDA events around the whole 3-kernel step (prologue + this kernel + epilogue, chained by programmatic "
                                      "dependent launch, so the kernel cannot be bracketed alone without breaking the overlap): a lower bound on "
                                      "the kernel's own rate",
                     "note": "the kernel is fp64-pipe / dependent-issue bound, not HBM bound: an N=1e5 light curve is L2-resident and shared by every "
                             "chain (BASELINE.md section 5), so the nominal-bytes figure is kept separately in roofline_hbm_nominal"},
        "roofline_hbm_nominal": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                 "algorithmic_bytes_per_launch": ALG_BYTES_PER_CHAIN_CADENCE * C_loc * N_CADENCES, "peak_source": peak_src,
                                 "note": "nominal 24 B per (chain,cadence) / step duration; real DRAM traffic is `roofline.traffic`"},
        "clocks": clocks, "wall_s": wall,
    }
    if kep is not None:
        line["kepler"] = kep
        k = kep["e~U(0,0.95)"]
        line["kepler_roofline"] = {"bound": "hbm", "achieved": k["gb_per_s"], "peak": peak, "unit": "GB/s", "frac": k["gb_per_s"] / peak,
                                   "traffic": _ncu_traffic(_first_profile("r02_kepler.txt", "r01_kepler.txt"), "kepler"), "algorithmic_bytes_per_launch": k["n"] * KEPLER_BYTES_PER_ELEMENT,
                                   "kernel": "kepler_fwd_kernel", "bytes_per_element": KEPLER_BYTES_PER_ELEMENT}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if other is not None:
        line["other_configs"] = other
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=CHAINS_PER_GPU, help="chains per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-kepler", action="store_true", help="skip the Kepler micro-benchmark")
    ap.add_argument("--no-extra", action="store_true", help="skip the brief configs[2] / configs[3] measurements")
    ap.add_argument("--extra", default="", help="comma list restricting the other-config measurements: c0,c2,gp,c4,hmc (default: all)")
    ap.add_argument("--same-seed", action="store_true", help="multi-GPU: every rank evaluates the same chains (identical per-rank work)")
    ap.add_argument("--full-sweep", action="store_true",
                    help="sum the Gaussian terms of every cadence in the sweep instead of using the data moments")
    args = ap.parse_args()
    if args.impl == "reference":
        # pin the OpenMP team before libgomp is loaded (the oracle is dlopen'ed later): unpinned threads were the
        # main source of the 2x run-to-run spread of this arm in round 1
        # torchrun exports OMP_NUM_THREADS=1 to every rank unless the user set it; this arm is "the reference's CPU path with all
        # the host threads it can use", run by rank 0 alone: drop that default so the OpenMP team spans the machine
        if ("TORCHELASTIC_RUN_ID" in os.environ or "LOCAL_RANK" in os.environ) and os.environ.get("OMP_NUM_THREADS") == "1":
            del os.environ["OMP_NUM_THREADS"]
        quota = _cpu_quota()
        if quota is not None and quota < (os.cpu_count() or 1) / 2 and "OMP_NUM_THREADS" not in os.environ:
            # a container with a CFS quota below the core count (the GPU boxes of this pool: 16 CPUs of 128): a team larger than
            # the quota is throttled by the scheduler in the middle of its parallel regions (and spinning idle threads burn the
            # quota) -- the 2x spread and the 9x-over-one-thread "scaling" of the 64-thread runs.  One thread per quota CPU,
            # placed by the kernel (the first 16 cores are not ours alone).
            os.environ["OMP_NUM_THREADS"] = str(max(1, int(quota)))
            os.environ.setdefault("OMP_PROC_BIND", "false")
        else:
            os.environ.setdefault("OMP_PROC_BIND", "close")
            os.environ.setdefault("OMP_PLACES", "cores")
        os.environ.setdefault("OMP_WAIT_POLICY", "active")
        arm_reference(args)
    else:
        arm_ours(args)


if __name__ == "__main__":
    main()
