"""Sampler adaptor: hands whole walker / live-point batches to the GPU.

The reference runs bilby's emcee wrapper, which calls `log_likelihood()` once per walker
per step from Python (simulate.py:191-192).  Here a sampler sees ONE call per step:

* `BoxPrior`            batched log-prior for the `.inits` kinds Uniform / Gaussian / Sine plus
                        the sin(incl) limits constraint (priors.py:344-377), numpy or torch;
* `log_prob_fn`         a vectorised log-posterior `(nwalkers, ndim) -> (nwalkers,)` for
                        `emcee.EnsembleSampler(..., vectorize=True)`;
* `BatchPool`           a `pool.map` shim for samplers that evaluate lists of points through a
                        pool (dynesty `queue_size`, bilby `npool`): the whole list is one batch;
* `EnsembleSampler`     a small built-in affine-invariant stretch-move sampler (Goodman & Weare
                        2010; the algorithm emcee implements) that keeps walkers on the device,
                        so BASELINE config 2 (4096 walkers) runs without emcee, which is not
                        installed in this image.
"""
import math
import numpy as np


def _is_torch(x):
    return type(x).__module__.startswith('torch')


class BoxPrior(object):
    """Independent priors from an `.inits` limits dict {name: [a, b, kind]} (priors.py:356-366):
    Uniform(a, b), Gaussian(mu=a, sigma=b), Sine(a, b) [density ~ sin x on (a, b)], plus the
    optional {name: [lo, hi]} limits on sin(incl) (bilby Constraint, priors.py:351-354)."""

    def __init__(self, limits, constraints=None, keys=None):
        self.keys = list(limits.keys()) if keys is None else list(keys)
        kinds, a, b = [], [], []
        for k in self.keys:
            lo, hi, kind = limits[k][0], limits[k][1], limits[k][2]
            if kind not in ('Uniform', 'Gaussian', 'Sine'):
                raise ValueError('unknown prior kind %r for %s' % (kind, k))
            kinds.append(kind)
            a.append(float(lo))
            b.append(float(hi))
        self.kinds = kinds
        self.a = np.array(a)
        self.b = np.array(b)
        silc = {} if constraints is None else constraints.get('sin_incl_limits_constraints', {})
        self.sin_limits = [(self.keys.index(k), float(v[0]), float(v[1])) for k, v in silc.items()]
        self._const = 0.0
        for kind, lo, hi in zip(kinds, a, b):
            if kind == 'Uniform':
                self._const += -math.log(hi - lo)
            elif kind == 'Gaussian':
                self._const += -math.log(hi) - 0.5 * math.log(2.0 * math.pi)
            else:
                self._const += -math.log(math.cos(lo) - math.cos(hi))
        self._dev = {}

    @property
    def ndim(self):
        return len(self.keys)

    def sample(self, n, seed=0):
        rng = np.random.default_rng(seed)
        out = np.empty((n, self.ndim))
        for j, kind in enumerate(self.kinds):
            if kind == 'Uniform':
                out[:, j] = rng.uniform(self.a[j], self.b[j], n)
            elif kind == 'Gaussian':
                out[:, j] = rng.normal(self.a[j], self.b[j], n)
            else:
                u = rng.random(n)
                out[:, j] = np.arccos(np.cos(self.a[j]) - u * (np.cos(self.a[j]) - np.cos(self.b[j])))
        return out

    def prior_transform(self, u):
        """Unit hypercube -> parameter space, (ndim,) or (N, ndim), numpy or torch: what nested samplers
        (dynesty, bilby's dynesty wrapper) ask a prior for.  Uniform: a + (b - a) u; Gaussian: inverse normal
        CDF; Sine: arccos(cos a - u (cos a - cos b)).  The sin(incl) limits are a constraint, not a
        transform: points outside them get log-likelihood -inf from `log_prob_fn` / `stn_logpost`."""
        if _is_torch(u):
            import torch
            x = u.clone().to(torch.float64)
            cols = x if x.dim() == 2 else x[None, :]
            for j, kind in enumerate(self.kinds):
                a, b = float(self.a[j]), float(self.b[j])
                if kind == 'Uniform':
                    cols[:, j] = a + (b - a) * cols[:, j]
                elif kind == 'Gaussian':
                    cols[:, j] = a + b * math.sqrt(2.0) * torch.erfinv(2.0 * cols[:, j] - 1.0)
                else:
                    cols[:, j] = torch.arccos(math.cos(a) - cols[:, j] * (math.cos(a) - math.cos(b)))
            return x
        from scipy.special import erfinv
        x = np.array(u, dtype=np.float64, copy=True)
        cols = x if x.ndim == 2 else x[None, :]
        for j, kind in enumerate(self.kinds):
            a, b = self.a[j], self.b[j]
            if kind == 'Uniform':
                cols[:, j] = a + (b - a) * cols[:, j]
            elif kind == 'Gaussian':
                cols[:, j] = a + b * math.sqrt(2.0) * erfinv(2.0 * cols[:, j] - 1.0)
            else:
                cols[:, j] = np.arccos(np.cos(a) - cols[:, j] * (np.cos(a) - np.cos(b)))
        return x

    def log_prob(self, theta):
        """(N, ndim) -> (N,) log prior density (−inf outside the support)."""
        if _is_torch(theta):
            return self._log_prob_torch(theta)
        theta = np.asarray(theta, dtype=np.float64)
        lp = np.full(theta.shape[0], self._const)
        for j, kind in enumerate(self.kinds):
            x = theta[:, j]
            if kind == 'Gaussian':
                lp += -0.5 * ((x - self.a[j]) / self.b[j]) ** 2
            else:
                inside = (x >= self.a[j]) & (x <= self.b[j])
                if kind == 'Sine':
                    with np.errstate(divide='ignore', invalid='ignore'):
                        lp += np.where(inside, np.log(np.where(inside, np.sin(x), 1.0)), 0.0)
                lp = np.where(inside, lp, -np.inf)
        for j, lo, hi in self.sin_limits:
            s = np.sin(theta[:, j])
            lp = np.where((s >= lo) & (s <= hi), lp, -np.inf)
        return lp

    def _log_prob_torch(self, theta):
        import torch
        key = str(theta.device)
        if key not in self._dev:
            self._dev[key] = (torch.as_tensor(self.a, device=theta.device), torch.as_tensor(self.b, device=theta.device))
        a, b = self._dev[key]
        ninf = torch.tensor(-math.inf, dtype=theta.dtype, device=theta.device)
        lp = torch.full((theta.shape[0],), self._const, dtype=theta.dtype, device=theta.device)
        for j, kind in enumerate(self.kinds):
            x = theta[:, j]
            if kind == 'Gaussian':
                lp = lp - 0.5 * ((x - a[j]) / b[j]) ** 2
            else:
                inside = (x >= a[j]) & (x <= b[j])
                if kind == 'Sine':
                    lp = lp + torch.where(inside, torch.log(torch.where(inside, torch.sin(x), torch.ones_like(x))),
                                          torch.zeros_like(x))
                lp = torch.where(inside, lp, ninf)
        for j, lo, hi in self.sin_limits:
            s = torch.sin(theta[:, j])
            lp = torch.where((s >= lo) & (s <= hi), lp, ninf)
        return lp


def log_prob_fn(likelihood, prior):
    """Vectorised log-posterior for emcee (vectorize=True) and the built-in sampler.  Points
    outside the prior support are not sent through the likelihood's arithmetic result."""
    keys = prior.keys

    def fn(theta):
        lp = prior.log_prob(theta)
        ll = likelihood.log_likelihood_batch(theta, keys=keys)
        if _is_torch(theta):
            import torch
            return torch.where(torch.isfinite(lp), lp + ll, lp)
        return np.where(np.isfinite(lp), lp + ll, lp)

    return fn


class BatchPool(object):
    """`pool.map(fn, list_of_points)` -> one batched GPU call, for samplers that evaluate several
    points per iteration through a pool (dynesty `pool=..., queue_size=K`; bilby `npool`).

        pool = BatchPool(like, prior.keys)
        sampler = dynesty.NestedSampler(pool.loglike, prior_transform, ndim, pool=pool, queue_size=256)

    Samplers send MORE than likelihood calls through `pool.map` (dynesty maps its proposal /
    evolution functions and the prior transform too), so only calls whose `fn` is the likelihood are
    batched: `pool.loglike` itself, a callable registered with `register(fn)`, or a wrapper object
    that holds one of those in a `loglikelihood`, `func`, `fn` or `__wrapped__` attribute (how
    dynesty and functools wrap it).  Every other `fn` is applied point by point, as a serial pool would."""

    _WRAPPER_ATTRS = ('loglikelihood', 'func', 'fn', '__wrapped__')

    def __init__(self, likelihood, keys, size=1024):
        self.like = likelihood
        self.keys = list(keys)
        self.size = int(size)
        self._known = []

        def loglike(theta):
            theta = np.asarray(theta, dtype=np.float64)
            if theta.ndim == 1:
                return float(self.like.log_likelihood_batch(theta[None, :], keys=self.keys)[0])
            return self.like.log_likelihood_batch(theta, keys=self.keys)

        self.loglike = loglike
        self.register(loglike)

    def register(self, fn):
        """Declare `fn` (the callable handed to the sampler as its log-likelihood) as batchable."""
        self._known.append(fn)
        return fn

    def _is_loglike(self, fn, depth=0):
        if any(fn is k for k in self._known):
            return True
        if depth >= 3:
            return False
        for attr in self._WRAPPER_ATTRS:
            inner = getattr(fn, attr, None)
            if inner is not None and inner is not fn and self._is_loglike(inner, depth + 1):
                return True
        return False

    def map(self, fn, points):
        pts = list(points)
        if len(pts) == 0:
            return []
        if not self._is_loglike(fn):
            return [fn(p) for p in pts]
        theta = np.asarray(pts, dtype=np.float64).reshape(len(pts), -1)
        return list(self.like.log_likelihood_batch(theta, keys=self.keys))

    def close(self):
        pass

    def join(self):
        pass


class EnsembleSampler(object):
    """Affine-invariant ensemble sampler with the stretch move, red/blue split so that each
    half-step is ONE batched log-probability call.  State lives in torch tensors on
    `device` (CUDA for the product; CPU tensors work too and are what the CPU tests use).

        s = EnsembleSampler(nwalkers, ndim, log_prob, device='cuda:0')
        s.run(p0, nsteps);  chain = s.chain  # (nsteps, nwalkers, ndim)
    """

    def __init__(self, nwalkers, ndim, log_prob, a=2.0, device='cpu', seed=0):
        import torch
        if nwalkers % 2 or nwalkers < 2 * ndim:
            raise ValueError('nwalkers must be even and at least 2*ndim')
        self.torch = torch
        self.nwalkers, self.ndim, self.a = nwalkers, ndim, float(a)
        self.log_prob = log_prob
        self.device = torch.device(device)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(seed)
        self.chain = None
        self.lnprob = None
        self.naccepted = torch.zeros((), dtype=torch.int64, device=self.device)
        self.nproposed = 0
        self.ncalls = 0

    def _call(self, x):
        self.ncalls += 1
        out = self.log_prob(x)
        if not _is_torch(out):
            out = self.torch.as_tensor(np.asarray(out), device=self.device)
        return out.to(self.torch.float64)

    def run(self, p0, nsteps, store=True):
        torch = self.torch
        x = torch.as_tensor(np.asarray(p0) if not _is_torch(p0) else p0).to(self.device, torch.float64).clone()
        lp = self._call(x)
        if not bool(torch.isfinite(lp).all()):
            raise ValueError('initial walkers must have finite log-probability')
        half = self.nwalkers // 2
        if store:
            self.chain = torch.empty((nsteps, self.nwalkers, self.ndim), dtype=torch.float64, device=self.device)
            self.lnprob = torch.empty((nsteps, self.nwalkers), dtype=torch.float64, device=self.device)
        sets = (slice(0, half), slice(half, self.nwalkers))
        for t in range(nsteps):
            for k in (0, 1):
                S, C = sets[k], sets[1 - k]
                u = torch.rand(half, generator=self.gen, device=self.device, dtype=torch.float64)
                z = ((self.a - 1.0) * u + 1.0) ** 2 / self.a                      # g(z) ~ 1/sqrt(z) on [1/a, a]
                j = torch.randint(0, half, (half,), generator=self.gen, device=self.device)
                y = x[C][j] + z[:, None] * (x[S] - x[C][j])
                lpy = self._call(y)
                lnq = (self.ndim - 1.0) * torch.log(z) + lpy - lp[S]
                acc = torch.log(torch.rand(half, generator=self.gen, device=self.device, dtype=torch.float64)) < lnq
                acc = acc & torch.isfinite(lpy)
                x[S] = torch.where(acc[:, None], y, x[S])
                lp[S] = torch.where(acc, lpy, lp[S])
                self.naccepted += acc.sum()
                self.nproposed += half
            if store:
                self.chain[t] = x
                self.lnprob[t] = lp
        self.last = (x, lp)
        return x, lp

    @property
    def acceptance_fraction(self):
        return float(self.naccepted) / max(self.nproposed, 1)


class FusedEnsembleSampler(object):
    """The same stretch move entirely inside the library (stn_ensemble_run): walkers, proposals and
    log-probabilities never leave the GPU, no torch arithmetic and no Python runs in the loop.

        s = FusedEnsembleSampler(likelihood, prior, nwalkers, seed=1)
        s.run(p0, nsteps, thin=10);  s.chain  # (nsteps // thin, nwalkers, ndim) CUDA tensor

    engine='auto' picks per problem: short series (the reference's real-data regime, ~10 epochs per
    source) run as ONE persistent cooperative kernel -- propose, log prior + log L inline, accept, a
    grid-wide barrier between the two half-steps, `nsteps` steps per launch; long series use the
    throughput kernel, three launches per half-step issued from C.  'persistent' / 'launches' force one.
    Random numbers are counter-based (Philox keyed by `seed`): a run is reproducible, independent of
    launch geometry and of the engine.  Columns of p0 / chain follow prior.keys.

    A likelihood built with a device LIST shares the ensemble among those GPUs (devices=None: the
    likelihood's list; stn_multi_ensemble_run): every GPU keeps a full copy of the walkers, evaluates its
    slice of each half-step's proposals with the throughput kernel and its accept kernel stores the
    accepted walkers into all copies over NVLink.  Same chain, bit for bit, as on one GPU.  Worth it for
    long series x large ensembles (>= ~1 ms of likelihood work per half-step); short series are
    latency-bound and run best as the one-GPU persistent kernel (devices=[d]).
    """

    ENGINES = {'auto': 0, 'persistent': 1, 'launches': 2}

    def __init__(self, likelihood, prior, nwalkers, a=2.0, seed=0, mode=None, plan=None, engine='auto',
                 devices=None):
        import torch
        from . import _lib
        ndim = prior.ndim
        if nwalkers % 2 or nwalkers < 2 * ndim:
            raise ValueError('nwalkers must be even and at least 2*ndim')
        self.torch = torch
        self.lib = _lib.load()
        self.check = _lib.check
        self.like, self.prior = likelihood, prior
        self.nwalkers, self.ndim, self.a, self.seed = nwalkers, ndim, float(a), int(seed)
        self.device = torch.device('cuda', likelihood.device)
        self.devices = list(likelihood.devices) if devices is None else [int(d) for d in devices]
        self.ctx = likelihood._context(prior.keys)
        if self.ctx.prior is not prior:
            self.ctx.set_prior(prior)
        self.multi = None
        if len(self.devices) > 1:
            if self.devices != list(likelihood.devices):
                raise ValueError('devices must be the likelihood\'s device list %r' % (likelihood.devices,))
            self.multi = likelihood._multi(prior.keys)
            if self.multi.prior is not prior:
                self.multi.set_prior(prior)
        self.mode = likelihood.mode if mode is None else _lib.MODES[mode]
        self.engine = self.ENGINES[engine]
        # The persistent kernel sums in the order of plan lanes = 1 | splits = 1; the launch engines (one or
        # several GPUs) default to the plan of the whole half-ensemble, which keeps all SMs busy.  Give
        # plan=make_plan(1, 1) to make every engine produce the same chain bit for bit.
        persistent = self.multi is None and (engine == 'persistent' or (
            engine == 'auto' and self.mode == _lib.MODE_FAST and ndim <= 64 and self.ctx.problem.evals_per_sample <= 512))
        if plan is None:
            self.plan = _lib.make_plan(1, 1) if persistent else self.ctx.auto_plan(nwalkers // 2)
        else:
            self.plan = int(plan)
        self.step = 0
        self.chain = None
        self.lnprob = None
        self.nproposed = 0
        self._nacc = torch.zeros(1, dtype=torch.int64, device=self.device)

    def run(self, p0, nsteps, thin=1, store=True):
        torch = self.torch
        from . import _lib
        W, P = self.nwalkers, self.ndim
        dev = self.device
        x = torch.as_tensor(np.asarray(p0) if not _is_torch(p0) else p0).to(dev, torch.float64).contiguous().clone()
        if tuple(x.shape) != (W, P):
            raise ValueError('p0 must be (nwalkers, ndim)')
        lp = torch.empty(W, dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        self.ctx.logpost_ptr(x.data_ptr(), W, P, _lib.LAYOUT_AOS, self.mode, self.plan, lp.data_ptr(), stream)
        if not bool(torch.isfinite(lp).all()):
            raise ValueError('initial walkers must have finite log-probability')
        nkeep = nsteps // thin if store else 0
        chain_ptr = lnprob_ptr = None
        if store:
            self.chain = torch.empty((nkeep, W, P), dtype=torch.float64, device=dev)
            self.lnprob = torch.empty((nkeep, W), dtype=torch.float64, device=dev)
            if nkeep:
                chain_ptr, lnprob_ptr = self.chain.data_ptr(), self.lnprob.data_ptr()
        if self.multi is not None:
            torch.cuda.synchronize(dev)
            xs = [x] + [x.to(torch.device('cuda', d)) for d in self.devices[1:]]
            lps = [lp] + [lp.to(torch.device('cuda', d)) for d in self.devices[1:]]
            for d in self.devices:
                torch.cuda.synchronize(d)
            self._nacc_multi = getattr(self, '_nacc_multi', 0) + self.multi.ensemble_run(
                [t.data_ptr() for t in xs], [t.data_ptr() for t in lps], W, P, self.a, self.seed, self.step, nsteps, thin,
                chain_ptr, lnprob_ptr, self.mode, self.plan)
        else:
            self.check(self.lib.stn_ensemble_run(self.ctx.handle, x.data_ptr(), lp.data_ptr(), W, P, self.a, self.seed,
                                                 self.step, nsteps, thin, chain_ptr, lnprob_ptr, self._nacc.data_ptr(),
                                                 self.mode, self.plan, self.engine, stream))
        self.step += nsteps
        self.nproposed += W * nsteps
        self.last = (x, lp)
        return x, lp

    @property
    def acceptance_fraction(self):
        return (float(self._nacc.item()) + getattr(self, '_nacc_multi', 0)) / max(self.nproposed, 1)
