"""Device-resident batched Hamiltonian Monte Carlo over many chains in lock-step: what
turns model+gradient evaluations per second into samples per second.

PyMC's NUTS (the sampler of /root/reference/docs/tutorials/data-and-models.md:458-509) advances ONE
chain per process and crosses the Python/C boundary at every leapfrog step.  Here every chain of the
batch takes the same number of leapfrog steps per transition, all state (positions, momenta, step
sizes, mass matrices, accept decisions) lives on the device, nothing synchronises with the host, and one
whole transition -- L fused logp+grad launches plus the elementwise updates around them -- is captured
in a CUDA graph and replayed.

``BatchedHMC`` is plain HMC (fixed number of leapfrog steps) with per-chain dual-averaging step sizes
and a diagonal mass matrix estimated in warm-up.  ``BatchedNUTS`` adds the No-U-Turn criterion: the
trajectory of every chain doubles until it turns back on itself (multinomial sampling over the tree,
generalised U-turn test on every completed subtree, as in Stan / PyMC), still in lock-step -- all chains
take the 2^j leapfrog steps of doubling j together, chains whose tree has already ended are masked, and
the host looks at the device once per doubling ("is any chain still growing?").
"""
import math

import torch

__all__ = ["BatchedHMC", "BatchedNUTS", "TransitPosterior"]


class TransitPosterior:
    """log posterior and its gradient for a batch of chains, in an unconstrained vector q[C, D]:

        per planet   log period, t0, sqrt(e) cos w, sqrt(e) sin w, b, log(r/R*)
        per chain    logit q1, logit q2 (Kipping 2013 limb darkening, distributions.py:98-102), mean flux

    with independent Normal priors on the unconstrained coordinates (prior_mu / prior_sigma, [D] or
    [C, D]).  The likelihood and its parameter gradients come from the fused kernels
    (fused.TransitModel); the change of variables is a handful of elementwise torch ops whose
    chain rule torch.autograd applies.  r_star = m_star = 1."""

    NPL = 6

    def __init__(self, model, n_planets, prior_mu, prior_sigma):
        self.model = model
        self.P = int(n_planets)
        self.D = self.NPL * self.P + 3
        dev = model.data.device
        self.mu = torch.as_tensor(prior_mu, dtype=torch.float64, device=dev)
        self.sigma = torch.as_tensor(prior_sigma, dtype=torch.float64, device=dev)

    @staticmethod
    def pack(period, t0, ecc, omega, b, ror, u1, u2, mean):
        """numpy/torch parameters ([C, P] and [C]) -> q[C, D]"""
        as_t = lambda x: torch.as_tensor(x, dtype=torch.float64)  # noqa: E731
        period, t0, ecc, omega, b, ror = [as_t(x) for x in (period, t0, ecc, omega, b, ror)]
        u1, u2, mean = [as_t(x) for x in (u1, u2, mean)]
        se = torch.sqrt(ecc)
        pl = torch.stack([torch.log(period), t0, se * torch.cos(omega), se * torch.sin(omega), b, torch.log(ror)], dim=-1)
        q1 = (u1 + u2) ** 2
        q2 = 0.5 * u1 / (u1 + u2)
        logit = lambda x: torch.log(x) - torch.log1p(-x)  # noqa: E731
        return torch.cat([pl.reshape(pl.shape[0], -1), torch.stack([logit(q1), logit(q2), mean], dim=-1)], dim=-1)

    def unpack(self, q):
        """q[C, D] -> (par[C, P, 8], star[C, 4], mean[C]) as differentiable torch expressions"""
        C = q.shape[0]
        pl = q[:, : self.NPL * self.P].reshape(C, self.P, self.NPL)
        period = torch.exp(pl[..., 0])
        h, k = pl[..., 2], pl[..., 3]
        ecc = h * h + k * k
        omega = torch.atan2(k, h)
        zero = torch.zeros_like(period)
        par = torch.stack([period, pl[..., 1], ecc, omega, pl[..., 4], torch.exp(pl[..., 5]), zero, zero], dim=-1)
        q1 = torch.sigmoid(q[:, -3])
        q2 = torch.sigmoid(q[:, -2])
        sq = torch.sqrt(q1)
        one = torch.ones_like(q1)
        star = torch.stack([one, one, 2.0 * sq * q2, sq * (1.0 - 2.0 * q2)], dim=-1)
        return par, star, q[:, -1]

    def _prior_vectors(self, C):
        """(prior_mu, prior_sigma, chain stride): [D] vectors shared by the chains (stride 0) or [C, D] rows (stride D)"""
        D = self.D
        if self.mu.dim() > 1 or self.sigma.dim() > 1:
            return self.mu.expand(C, D).contiguous(), self.sigma.expand(C, D).contiguous(), D
        return self.mu.expand(D).contiguous(), self.sigma.expand(D).contiguous(), 0

    def fused_args(self, C):
        """what the C-ABI posterior / HMC entry points need besides the state; None when this posterior cannot take the
        fused path (cadence-parallel model, CPU data)"""
        import ctypes as C_

        from . import _lib

        m = self.model
        if m.cadence_group is not None or not m.data.t.is_cuda:
            return None
        pv = self._prior_vectors(C)
        lib = _lib.require_cuda()
        d = m.data
        need = lib.xo_transit_posterior_workspace(d.n, C, self.P)
        ws = getattr(self, "_fused_ws", None)
        if ws is None or ws.numel() * 8 < need:
            ws = self._fused_ws = torch.empty(need // 8 + 1, dtype=torch.float64, device=d.device)
        ptr = lambda x: C_.c_void_p(x.data_ptr()) if x is not None else None  # noqa: E731
        mom = d.moments() if m.use_moments else None
        self._keep = (pv, mom)
        return dict(lib=lib, head=(ptr(d.t), ptr(d.y), ptr(d.ivar), d.n, ptr(mom)), prior=(ptr(pv[0]), ptr(pv[1]), pv[2]),
                    tail=(ptr(m.dtk), ptr(m.wk), m.dtk.numel(), m.texp, int(m.use_in_transit), d.lognorm), ws=ws)

    def __call__(self, q):
        """(log posterior[C], gradient[C, D]); no host synchronisation.  One C call (xo_transit_posterior_logp_grad:
        unpack kernel, fused model evaluation, chain-rule + prior kernel); ``autograd_call`` is the same mathematics
        through torch autograd."""
        import ctypes as C_

        from . import _lib

        q = q.detach().contiguous()
        fa = self.fused_args(q.shape[0]) if q.is_cuda else None
        if fa is not None:
            C = q.shape[0]
            if q.dtype != torch.float64 or tuple(q.shape) != (C, self.D):
                raise ValueError("q must be float64 [C, %d]" % self.D)
            lp = torch.empty(C, dtype=torch.float64, device=q.device)
            g = torch.empty_like(q)
            ptr = lambda x: C_.c_void_p(x.data_ptr())  # noqa: E731
            with torch.cuda.device(q.device):
                _lib.check(fa["lib"].xo_transit_posterior_logp_grad(*fa["head"], ptr(q), C, self.P, *fa["prior"], *fa["tail"], ptr(lp), ptr(g),
                                                                    ptr(fa["ws"]), fa["ws"].numel() * 8, _lib.stream_ptr()),
                           "transit_posterior_logp_grad")
            return lp, g
        return self.autograd_call(q)

    def autograd_call(self, q):
        """the same posterior through torch autograd (reference implementation of the fused path; per-chain priors)"""
        q = q.detach().requires_grad_(True)
        with torch.enable_grad():
            par, star, mean = self.unpack(q)
            z = (q - self.mu) / self.sigma
            lp = self.model(par, star, mean) - 0.5 * (z * z).sum(dim=-1)
            (g,) = torch.autograd.grad(lp.sum(), q)
        return lp.detach(), g


class BatchedHMC:
    """HMC over C chains in lock-step.  ``logp_grad(q[C, D]) -> (logp[C], grad[C, D])`` must not
    synchronise with the host (TransitPosterior, or any torch expression).

    state (device tensors): q, logp, grad, step size eps[C], inverse diagonal mass inv_mass[C, D]."""

    def __init__(self, logp_grad, q0, step_size=1e-2, n_leapfrog=16, seed=0, target_accept=0.8, use_graph=True, fused=None):
        self.f = logp_grad
        self.q = q0.detach().clone().contiguous()
        dev = self.q.device
        C, D = self.q.shape
        self.L = int(n_leapfrog)
        self.eps = torch.full((C,), float(step_size), dtype=torch.float64, device=dev)
        self.inv_mass = torch.ones((C, D), dtype=torch.float64, device=dev)
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(int(seed))
        self.logp, self.grad = self.f(self.q)
        self.logp, self.grad = self.logp.clone(), self.grad.clone()
        self.accept_prob = torch.zeros(C, dtype=torch.float64, device=dev)
        self.energy_error = torch.zeros(C, dtype=torch.float64, device=dev)
        self.target = float(target_accept)
        self.n_evals = 0
        self._graph = None
        self._use_graph = bool(use_graph) and dev.type == "cuda"
        # a TransitPosterior with chain-independent priors advances through ONE C call per transition
        # (xo_hmc_transition: momentum draw, L x (model evaluation + chain rule + kick + drift), accept / reject)
        self._fused = None
        self._seed = int(seed)
        self._counter = 0
        if fused is not False and dev.type == "cuda" and isinstance(logp_grad, TransitPosterior) and type(self) is BatchedHMC:
            self._fused = logp_grad.fused_args(C)
        # dual averaging (Hoffman & Gelman 2014, algorithm 5), one instance per chain
        self._mu = torch.log(10.0 * self.eps)
        self._log_eps_bar = torch.zeros_like(self.eps)
        self._hbar = torch.zeros_like(self.eps)
        self._m = 0

    # one transition for every chain; everything in place so that it can be captured once and replayed
    def _transition(self):
        q, eps, im = self.q, self.eps[:, None], self.inv_mass
        p = torch.randn(q.shape, dtype=q.dtype, device=q.device, generator=self.gen) / torch.sqrt(im)
        h0 = -self.logp + 0.5 * (p * p * im).sum(dim=-1)
        qn, g, lp = q, self.grad, self.logp
        p = p + 0.5 * eps * g
        for i in range(self.L):
            qn = qn + eps * im * p
            lp, g = self.f(qn)
            p = p + (eps if i + 1 < self.L else 0.5 * eps) * g
        h1 = -lp + 0.5 * (p * p * im).sum(dim=-1)
        dh = h0 - h1
        dh = torch.where(torch.isfinite(dh), dh, torch.full_like(dh, -float("inf")))
        a = torch.exp(torch.clamp(dh, max=0.0))
        u = torch.rand(a.shape, dtype=a.dtype, device=a.device, generator=self.gen)
        acc = u < a
        self.q.copy_(torch.where(acc[:, None], qn, q))
        self.logp.copy_(torch.where(acc, lp, self.logp))
        self.grad.copy_(torch.where(acc[:, None], g, self.grad))
        self.accept_prob.copy_(a)
        self.energy_error.copy_(-dh)

    def _fused_transition(self):
        import ctypes as C_

        from . import _lib

        fa = self._fused
        ptr = lambda x: C_.c_void_p(x.data_ptr())  # noqa: E731
        C = self.q.shape[0]
        with torch.cuda.device(self.q.device):
            _lib.check(fa["lib"].xo_hmc_transition(*fa["head"], C, self.f.P, *fa["prior"], *fa["tail"], ptr(self.q), ptr(self.logp),
                                                   ptr(self.grad), ptr(self.eps), ptr(self.inv_mass), self.L, self._seed, self._counter,
                                                   ptr(self.accept_prob), ptr(self.energy_error), ptr(fa["ws"]), fa["ws"].numel() * 8,
                                                   _lib.stream_ptr()), "hmc_transition")
        self._counter += 1
        self.n_evals += self.L

    def transition(self):
        if self._fused is not None:
            self._fused_transition()
            return
        if not self._use_graph:
            self._transition()
        else:
            if self._graph is None:
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):  # warm-up on a side stream, as graph capture requires
                    self._transition()
                torch.cuda.current_stream().wait_stream(s)
                self._graph = torch.cuda.CUDAGraph()
                self._graph.register_generator_state(self.gen)
                with torch.cuda.graph(self._graph):
                    self._transition()
                # the side-stream run above was a real transition (state and generator advanced): it IS this call's
                # transition; the captured graph is replayed from the next call on
                self.n_evals += self.L
                return
            self._graph.replay()
        self.n_evals += self.L

    def _adapt_step(self, gamma=0.05, t0=10.0, kappa=0.75):
        self._m += 1
        m = self._m
        w = 1.0 / (m + t0)
        self._hbar.mul_(1.0 - w).add_(w * (self.target - self.accept_prob))
        log_eps = self._mu - math.sqrt(m) / gamma * self._hbar
        eta = m ** (-kappa)
        self._log_eps_bar.mul_(1.0 - eta).add_(eta * log_eps)
        self.eps.copy_(torch.exp(log_eps))

    def warmup(self, n, adapt_mass=True):
        """n transitions of step-size adaptation; half way through, the diagonal mass matrix is set from
        the variance of the positions seen since the first quarter, pooled over the chains (they all
        target the same posterior, so C chains x k draws estimate it far better than k draws of one
        chain), and the step-size search restarts."""
        half = n // 2 if adapt_mass else -1
        s1 = torch.zeros_like(self.q)
        s2 = torch.zeros_like(self.q)
        k = 0
        for it in range(n):
            self.transition()
            self._adapt_step()
            if adapt_mass and it >= n // 4 and it < half:
                s1 += self.q
                s2 += self.q * self.q
                k += 1
            if it + 1 == half and k > 4:
                nk = k * self.q.shape[0]
                m1 = s1.sum(dim=0, keepdim=True) / nk
                var = (s2.sum(dim=0, keepdim=True) / nk - m1 * m1) * (nk / (nk - 1.0))
                var = ((nk / (nk + 5.0)) * var + 1e-3 * (5.0 / (nk + 5.0))).expand_as(self.q)  # Stan's shrinkage
                # rescale the step so that eps^2 * inv_mass keeps its order of magnitude
                scale = torch.sqrt(self.inv_mass.mean(dim=-1) / var.mean(dim=-1))
                self.inv_mass.copy_(var)
                self.eps.mul_(scale)
                self._mu = torch.log(10.0 * self.eps)
                self._log_eps_bar.zero_()
                self._hbar.zero_()
                self._m = 0
        if self._m > 0:
            self.eps.copy_(torch.exp(self._log_eps_bar))

    def sample(self, n):
        """n transitions; returns dict(q[n, C, D], logp[n, C], accept_prob[n, C], energy_error[n, C]) on the device."""
        C, D = self.q.shape
        dev = self.q.device
        out = dict(q=torch.empty((n, C, D), dtype=torch.float64, device=dev), logp=torch.empty((n, C), dtype=torch.float64, device=dev),
                   accept_prob=torch.empty((n, C), dtype=torch.float64, device=dev),
                   energy_error=torch.empty((n, C), dtype=torch.float64, device=dev))
        for i in range(n):
            self.transition()
            out["q"][i].copy_(self.q)
            out["logp"][i].copy_(self.logp)
            out["accept_prob"][i].copy_(self.accept_prob)
            out["energy_error"][i].copy_(self.energy_error)
        return out


def _popcount(n):
    return bin(n).count("1")


def _trailing_ones(n):
    k = 0
    while n & 1:
        n >>= 1
        k += 1
    return k


class BatchedNUTS(BatchedHMC):
    """No-U-Turn sampler over C chains in lock-step (Hoffman & Gelman 2014; multinomial variant with
    the generalised U-turn criterion of Betancourt 2017, the scheme of Stan and of PyMC's NUTS that the
    reference's tutorials run, docs/tutorials/data-and-models.md:458-509).

    One transition: every chain draws a momentum and grows a trajectory by doubling.  Doubling j picks a
    direction per chain and integrates 2^j leapfrog steps from that end of the tree; the new half is
    checked for U-turns on each of its completed sub-trees with the O(log) checkpoint scheme (momentum
    and momentum-sum remembered at even leaves), merged into the tree by biased progressive sampling, and
    the merged tree is checked once more.  A chain stops when it turns, diverges (energy error > 1000) or
    reaches ``max_depth``; all chains integrate together, stopped ones are masked.  Step sizes adapt on
    the mean Metropolis acceptance over the leaves of the last tree (dual averaging), the diagonal mass
    matrix on the pooled variance of the warm-up draws -- both inherited from BatchedHMC.

    Extra state after a transition: ``depth[C]`` (doublings used), ``diverging[C]``, ``n_leapfrog``."""

    _SCALARS = ("h0", "log_w", "lp_prop", "active", "depth", "diverged", "acc_sum", "n_leaves", "dh_max", "sgn", "sub_ok", "s_logw", "s_lp",
                "go_right")
    _VECTORS = ("ql", "pl", "gl", "qr", "pr", "gr", "q_prop", "g_prop", "p_sum", "q", "p", "g", "s_q", "s_g", "s_psum")

    def __init__(self, logp_grad, q0, step_size=1e-2, max_depth=8, seed=0, target_accept=0.8, use_graph=True, fused_tree=None):
        super().__init__(logp_grad, q0, step_size=step_size, n_leapfrog=1, seed=seed, target_accept=target_accept, use_graph=False)
        self.max_depth = int(max_depth)
        C, D = self.q.shape
        dev, dt = self.q.device, self.q.dtype
        self.depth = torch.zeros(C, dtype=torch.int64, device=dev)
        self.diverging = torch.zeros(C, dtype=torch.bool, device=dev)
        self.n_leapfrog = 0
        # The tree lives in persistent buffers that the pieces of a transition (begin, doubling j, end) update in place, so
        # that each piece is ONE CUDA graph (captured the first time it runs, replayed afterwards): a doubling of depth j is
        # 2^j model evaluations and ~50 small tensor kernels per leaf, launch-bound when issued one by one.  The host looks at
        # the device once per doubling ("is any chain still growing?"), exactly as without graphs.
        self._nuts_graph = bool(use_graph) and dev.type == "cuda"
        self._pieces = {}
        # fused_tree (default on CUDA, float64): the tree bookkeeping of a doubling runs in the library's NUTS kernels
        # (csrc/xo_nuts.cu: two launches per leaf instead of ~50 tensor operations); the tensor code below stays as the
        # CPU path and as the implementation the kernels are checked against (same uniforms -> same states).
        self._fused_tree = (dev.type == "cuda" and dt == torch.float64) if fused_tree is None else bool(fused_tree)
        z = lambda *shape, dtype=dt: torch.zeros(shape, dtype=dtype, device=dev)  # noqa: E731
        if self._fused_tree:
            import ctypes as C_

            from . import _lib

            lib = _lib.require_cuda()
            offs = (C_.c_int64 * 31)()
            n = lib.xo_nuts_state_layout(C, D, self.max_depth, offs)
            self._state = z(n)
            self._lib = lib
            view = lambda o, *shape: self._state[o:o + math.prod(shape)].view(*shape)  # noqa: E731
            self._t = {k: view(offs[i], C) for i, k in enumerate(self._SCALARS)}
            self._t.update({k: view(offs[len(self._SCALARS) + i], C, D) for i, k in enumerate(self._VECTORS)})
            self._t["ck_p"] = view(offs[29], self.max_depth + 1, C, D)
            self._t["ck_s"] = view(offs[30], self.max_depth + 1, C, D)
        else:
            self._t = dict(h0=z(C), ql=z(C, D), pl=z(C, D), gl=z(C, D), qr=z(C, D), pr=z(C, D), gr=z(C, D), q_prop=z(C, D), lp_prop=z(C),
                           g_prop=z(C, D), p_sum=z(C, D), log_w=z(C), active=z(C, dtype=torch.bool), depth=z(C, dtype=torch.int64),
                           diverged=z(C, dtype=torch.bool), acc_sum=z(C), n_leaves=z(C), dh_max=z(C),
                           ck_p=z(self.max_depth + 1, C, D), ck_s=z(self.max_depth + 1, C, D))

    @staticmethod
    def _turning(im, p_left, p_right, p_sum):
        adj = p_sum - 0.5 * (p_left + p_right)
        return ((im * p_left * adj).sum(-1) <= 0.0) | ((im * p_right * adj).sum(-1) <= 0.0)

    # ---- the three pieces of a transition; all state in self._t, updated in place ----
    def _begin(self):
        T, im = self._t, self.inv_mass
        p0 = torch.randn(self.q.shape, dtype=self.q.dtype, device=self.q.device, generator=self.gen) / torch.sqrt(im)
        T["h0"].copy_(-self.logp + 0.5 * (p0 * p0 * im).sum(-1))
        # the tree: both edges, the current proposal, sum of momenta, log of the summed weights exp(-H)
        for k in ("ql", "qr", "q_prop"):
            T[k].copy_(self.q)
        for k in ("pl", "pr", "p_sum"):
            T[k].copy_(p0)
        for k in ("gl", "gr", "g_prop"):
            T[k].copy_(self.grad)
        T["lp_prop"].copy_(self.logp)
        T["log_w"].copy_(-T["h0"])  # weight of the initial point
        T["active"].fill_(True)
        T["diverged"].fill_(False)
        for k in ("depth", "acc_sum", "n_leaves", "dh_max"):
            T[k].zero_()

    def _doubling_fused(self, j):
        """one doubling through the library's NUTS kernels; the uniforms are drawn here, in the order the tensor code draws them"""
        import ctypes as C_

        from . import _lib

        T, lib = self._t, self._lib
        C, D = self.q.shape
        dev, dt = self.q.device, self.q.dtype
        ptr = lambda x: C_.c_void_p(x.data_ptr())  # noqa: E731
        st, md, stream = ptr(self._state), self.max_depth, _lib.stream_ptr()
        eps, im = self.eps.contiguous(), self.inv_mass.contiguous()
        rand = lambda: torch.rand(C, dtype=dt, device=dev, generator=self.gen)  # noqa: E731
        with torch.cuda.device(dev):
            u = rand()
            _lib.check(lib.xo_nuts_start(st, C, D, md, ptr(u), ptr(self.logp), stream), "nuts_start")
            for n in range(2 ** j):
                _lib.check(lib.xo_nuts_leaf_pre(st, C, D, md, ptr(eps), ptr(im), stream), "nuts_leaf_pre")
                lp, g = self.f(T["q"])
                lp, g, u = lp.contiguous(), g.contiguous(), rand()
                _lib.check(lib.xo_nuts_leaf_post(st, C, D, md, ptr(eps), ptr(im), ptr(lp), ptr(g), ptr(u), n, stream), "nuts_leaf_post")
            u = rand()
            _lib.check(lib.xo_nuts_merge(st, C, D, md, ptr(im), ptr(u), stream), "nuts_merge")

    def _doubling(self, j):
        if self._fused_tree:
            return self._doubling_fused(j)
        T, im = self._t, self.inv_mass
        C, D = self.q.shape
        dev, dt = self.q.device, self.q.dtype
        eps = self.eps[:, None]
        h0, active = T["h0"], T["active"].clone()
        ck_p, ck_s = T["ck_p"], T["ck_s"]
        neg_inf = torch.full((C,), -float("inf"), dtype=dt, device=dev)
        dh_max, diverged = T["dh_max"].clone(), T["diverged"].clone()
        acc_sum, n_leaves = T["acc_sum"].clone(), T["n_leaves"].clone()
        go_right = torch.rand(C, dtype=dt, device=dev, generator=self.gen) < 0.5
        sgn = torch.where(go_right, 1.0, -1.0).to(dt)[:, None]
        # start from the edge the new half grows out of
        q = torch.where(go_right[:, None], T["qr"], T["ql"])
        p = torch.where(go_right[:, None], T["pr"], T["pl"])
        g = torch.where(go_right[:, None], T["gr"], T["gl"])
        sub_ok = active.clone()  # still building a valid (non-turning, non-diverging) half
        s_psum = torch.zeros_like(p)
        s_logw = neg_inf.clone()
        s_q, s_lp, s_g = q.clone(), self.logp.clone(), g.clone()
        for n in range(2 ** j):
            p = p + 0.5 * sgn * eps * g
            q = q + sgn * eps * im * p
            lp, g = self.f(q)
            p = p + 0.5 * sgn * eps * g
            h = -lp + 0.5 * (p * p * im).sum(-1)
            dh = h - h0
            dh = torch.where(torch.isfinite(dh), dh, torch.full_like(dh, float("inf")))
            dh_max = torch.where(active & (dh.abs() > dh_max.abs()) & torch.isfinite(dh), dh, dh_max)
            div_now = sub_ok & (dh > 1000.0)
            diverged = diverged | div_now
            sub_ok = sub_ok & ~div_now
            # acceptance statistic over every leaf that was actually visited
            acc_sum = acc_sum + torch.where(active, torch.exp(torch.clamp(-dh, max=0.0)), torch.zeros_like(dh))
            n_leaves = n_leaves + active.to(dt)
            # progressive (uniform-by-weight) sampling of the proposal inside this half
            w_new = torch.where(sub_ok, -h, neg_inf)
            s_logw_new = torch.logaddexp(s_logw, w_new)
            take = sub_ok & (torch.log(torch.rand(C, dtype=dt, device=dev, generator=self.gen)) < w_new - s_logw_new)
            s_q = torch.where(take[:, None], q, s_q)
            s_lp = torch.where(take, lp, s_lp)
            s_g = torch.where(take[:, None], g, s_g)
            s_logw = torch.where(sub_ok, s_logw_new, s_logw)
            s_psum = torch.where(sub_ok[:, None], s_psum + p, s_psum)
            # U-turn checks inside the half: checkpoints at even leaves, checks at odd ones
            idx_max = _popcount(n >> 1)
            if n % 2 == 0:
                ck_p[idx_max].copy_(p)
                ck_s[idx_max].copy_(s_psum)
            else:
                idx_min = idx_max - _trailing_ones(n) + 1
                for i in range(idx_max, idx_min - 1, -1):
                    sub_sum = s_psum - ck_s[i] + ck_p[i]
                    sub_ok = sub_ok & ~self._turning(im, ck_p[i], p, sub_sum)
        # merge the half into the tree (only where it is valid)
        merged = active & sub_ok
        take = merged & (torch.log(torch.rand(C, dtype=dt, device=dev, generator=self.gen)) < s_logw - T["log_w"])
        T["q_prop"].copy_(torch.where(take[:, None], s_q, T["q_prop"]))
        T["lp_prop"].copy_(torch.where(take, s_lp, T["lp_prop"]))
        T["g_prop"].copy_(torch.where(take[:, None], s_g, T["g_prop"]))
        T["log_w"].copy_(torch.where(merged, torch.logaddexp(T["log_w"], s_logw), T["log_w"]))
        T["p_sum"].copy_(torch.where(merged[:, None], T["p_sum"] + s_psum, T["p_sum"]))
        mr, ml = (merged & go_right)[:, None], (merged & ~go_right)[:, None]
        for k, v in (("q", q), ("p", p), ("g", g)):
            T[k + "r"].copy_(torch.where(mr, v, T[k + "r"]))
            T[k + "l"].copy_(torch.where(ml, v, T[k + "l"]))
        T["depth"].add_(active.to(torch.int64))
        T["dh_max"].copy_(dh_max)
        T["diverged"].copy_(diverged)
        T["acc_sum"].copy_(acc_sum)
        T["n_leaves"].copy_(n_leaves)
        # the whole tree must not have turned either
        turned = self._turning(im, T["pl"], T["pr"], T["p_sum"])
        T["active"].copy_(merged & ~turned)

    def _end(self):
        T = self._t
        self.q.copy_(T["q_prop"])
        self.logp.copy_(T["lp_prop"])
        self.grad.copy_(T["g_prop"])
        self.accept_prob.copy_(T["acc_sum"] / torch.clamp(T["n_leaves"], min=1.0))
        self.energy_error.copy_(T["dh_max"])  # largest energy error met along the trajectory
        self.depth.copy_(T["depth"])          # (the fused state keeps counts and flags as doubles: copy_ converts)
        self.diverging.copy_(T["diverged"] != 0)

    def _run(self, key, fn, *args):
        """run one piece: eagerly, or as its CUDA graph (first call: a real run on a side stream, then the capture)"""
        if not self._nuts_graph:
            fn(*args)
            return
        g = self._pieces.get(key)
        if g is not None:
            g.replay()
            return
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):  # this run IS the piece's execution for this call; the graph serves the later ones
            fn(*args)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        g.register_generator_state(self.gen)
        with torch.cuda.graph(g):
            fn(*args)
        self._pieces[key] = g

    def _transition(self):
        self._run("begin", self._begin)
        for j in range(self.max_depth):
            self._run(("doubling", j), self._doubling, j)
            self.n_leapfrog += 2 ** j
            if j + 1 < self.max_depth and not bool(self._t["active"].any()):  # the one look at the device per doubling
                break
        self._run("end", self._end)
        self.n_evals = self.n_leapfrog - self.L  # transition() adds L (= 1): n_evals counts model evaluations

