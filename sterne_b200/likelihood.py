"""Drop-in `Gaussianlikelihood` and `positions` backed by the sm_100a kernels.

Same class name, constructor arguments, `.parameters` dict and no-argument
`log_likelihood()` as the reference (simulate.py:325-384), plus the batched entry
point `log_likelihood_batch(theta)` for vectorised samplers.  `positions(...)` keeps
the reference signature (positions.py:13).
"""
import collections
import numpy as np

from . import _lib
from .constants import SENTINEL, ROOTS
from .problem import CompiledProblem, DeviceContext, MultiDeviceContext
from .shares import filter_dictionary_of_parameter_with_index

try:                                            # bilby is optional: samplers need it, the math does not
    from bilby import Likelihood as _LikelihoodBase
except Exception:                               # pragma: no cover - bilby absent in this image
    class _LikelihoodBase(object):
        """Minimal stand-in with bilby.Likelihood's attributes."""

        def __init__(self, parameters=None):
            self.parameters = parameters if parameters is not None else {}
            self._meta_data = None

        def log_likelihood(self):
            raise NotImplementedError

        def noise_log_likelihood(self):
            return np.nan

        def log_likelihood_ratio(self):
            return self.log_likelihood() - self.noise_log_likelihood()


_default_earth_provider = None


def set_default_earth_provider(provider):
    """Earth-ephemeris provider used when none is passed explicitly (ephemeris.py)."""
    global _default_earth_provider
    _default_earth_provider = provider


def _resolve_provider(provider):
    if provider is not None:
        return provider
    if _default_earth_provider is not None:
        return _default_earth_provider
    from .ephemeris import default_provider
    return default_provider()


def _is_torch_tensor(x):
    return type(x).__module__.startswith('torch') and hasattr(x, 'data_ptr')


class Gaussianlikelihood(_LikelihoodBase):
    """Sum over sources of -1/2 sum((obs - model)/sigma)^2 - sum(log sigma), plus the
    optional a1dot and sin(incl) Gaussian terms (simulate.py:362-384), evaluated on a B200.

    Positional arguments are the reference's (simulate.py:326).  `positions` is accepted
    for signature compatibility and ignored: the model is the device kernel.
    Keyword-only extras: earth_provider / earth_xyz (ephemeris.py), device, mode
    ('fast' | 'strict'), plan (0 = automatic; see stn_loglike in include/sterne_b200.h).
    `device` is a CUDA ordinal or a list of them: with a list, host (numpy) batches are cut into
    contiguous slices over all listed GPUs from this one process (stn_multi_loglike_host) and a
    vector's log L has the same bits as on one GPU; everything else runs on the first device.
    """

    def __init__(self, refepoch, list_of_dict_timing, list_of_dict_VLBI, shares, positions=None,
                 DoD_additional_constraints=None, a1dot_constraints=False, *, earth_provider=None,
                 earth_xyz=None, device=0, mode='fast', plan=0):
        self.refepoch = refepoch
        self.LoD_VLBI = list_of_dict_VLBI
        self.LoD_timing = list_of_dict_timing
        self.positions = positions
        self.shares = shares
        self.number_of_pmparins = len(self.LoD_VLBI)
        if DoD_additional_constraints is None:
            DoD_additional_constraints = {'sin_incl_Gaussian_constraints': {}, 'sin_incl_limits_constraints': {}}
        self.DoD_additional_constraints = DoD_additional_constraints
        self.dict_sin_incl_Gaussian_constraints = DoD_additional_constraints['sin_incl_Gaussian_constraints']
        self.a1dot_constraints = a1dot_constraints
        if isinstance(device, (list, tuple)):
            self.devices = [int(d) for d in device]
            if not self.devices:
                raise ValueError('device list is empty')
        else:
            self.devices = [int(device)]
        self.device = self.devices[0]
        self._multis = {}
        self.mode = _lib.MODES[mode]
        self.plan = int(plan)
        needs_px = any(int(v) >= 0 for v in shares[7])
        self._provider = _resolve_provider(earth_provider) if (needs_px and earth_xyz is None) else earth_provider
        self._earth_xyz = earth_xyz
        self._contexts = {}
        self._one = None
        ctx = self._context(None)
        super().__init__(dict.fromkeys(ctx.problem.default_names))

    # ------------------------------------------------------------------ contexts
    def _context(self, keys):
        """Device context whose column order is `keys` (default: shares order).  Samplers
        order columns by their prior keys (.inits line order), hence the cache."""
        k = None if keys is None else tuple(keys)
        ctx = self._contexts.get(k)
        if ctx is None:
            problem = CompiledProblem(self.refepoch, self.LoD_timing, self.LoD_VLBI, self.shares,
                                      earth_provider=self._provider,
                                      DoD_additional_constraints=self.DoD_additional_constraints,
                                      a1dot_constraints=self.a1dot_constraints, names=keys,
                                      earth_xyz=self._earth_xyz)
            if self._earth_xyz is None:
                self._earth_xyz = [src['earth_xyz'] for src in problem.sources]   # evaluate the ephemeris once
            ctx = DeviceContext(problem, self.device)
            self._contexts[k] = ctx
            if k is None:
                self._contexts[tuple(problem.default_names)] = ctx
        return ctx

    def _multi(self, keys):
        """Multi-device context (one StnMulti) for column order `keys`; only with a device list."""
        ctx = self._context(keys)
        k = tuple(ctx.problem.names)
        m = self._multis.get(k)
        if m is None:
            m = self._multis[k] = MultiDeviceContext(ctx.problem, self.devices)
        return m

    @property
    def parameter_names(self):
        return list(self._context(None).problem.default_names)

    @property
    def evals_per_sample(self):
        return self._context(None).problem.evals_per_sample

    # ------------------------------------------------------------ reference API
    def log_likelihood(self, parameters=None):
        """simulate.py:362: one parameter vector taken from self.parameters (or, for bilby
        versions that pass it, from the `parameters` argument).  A batch of one."""
        one = self._one
        if one is None:
            ctx = self._context(None)
            names = tuple(ctx.problem.default_names)
            theta = np.empty((1, len(names)), dtype=np.float64)
            out = np.empty(1, dtype=np.float64)
            one = self._one = (ctx, names, theta, out, theta.ctypes.data, out.ctypes.data)
        ctx, names, theta, out, tp, op = one
        src = self.parameters if parameters is None else parameters
        row = theta[0]
        for i, k in enumerate(names):
            row[i] = src[k]
        ctx.loglike_host_ptr(tp, 1, len(names), _lib.LAYOUT_AOS, self.mode, self.plan, op)
        return float(out[0])

    # ------------------------------------------------------------- batched API
    def log_likelihood_batch(self, theta, keys=None, out=None, layout='aos', mode=None, plan=None):
        """log-L of N parameter vectors.

        theta : (N, P) float64, numpy array or CUDA torch tensor (layout='aos'), or
                (P, N) with layout='soa'.  Columns follow list(self.parameters) unless
                `keys` names them.  Returns (N,) of the same kind (numpy / torch).
        """
        ctx = self._context(keys)
        P = ctx.problem.n_params
        mode = self.mode if mode is None else _lib.MODES[mode]
        plan = self.plan if plan is None else int(plan)
        lay = _lib.LAYOUT_AOS if layout == 'aos' else _lib.LAYOUT_SOA
        if _is_torch_tensor(theta):
            import torch
            if not theta.is_cuda:
                return torch.from_numpy(self.log_likelihood_batch(theta.numpy(), keys, None, layout, mode, plan))
            if theta.dtype != torch.float64 or theta.dim() != 2:
                raise TypeError('theta must be a 2-D float64 tensor')
            if theta.device.index != self.device:
                raise ValueError('theta lives on cuda:%s but this likelihood was built for cuda:%d'
                                 % (theta.device.index, self.device))
            if theta.stride(1) != 1:
                theta = theta.contiguous()
            n, ld = self._shape(theta.shape, theta.stride(0), lay, P)
            if out is None:
                out = torch.empty(n, dtype=torch.float64, device=theta.device)
            elif (not _is_torch_tensor(out) or out.dtype != torch.float64 or out.device != theta.device
                  or out.numel() != n or not out.is_contiguous()):
                raise TypeError('out must be a contiguous float64 tensor of %d elements on %s' % (n, theta.device))
            stream = torch.cuda.current_stream(theta.device).cuda_stream
            ctx.loglike_ptr(theta.data_ptr(), n, ld, lay, mode, plan, out.data_ptr(), stream)
            return out
        theta = np.asarray(theta, dtype=np.float64)
        if theta.ndim != 2:
            raise TypeError('theta must be 2-D')
        if theta.strides[1] != 8 or theta.strides[0] % 8 != 0 or theta.strides[0] < 0:
            theta = np.ascontiguousarray(theta)
        n, ld = self._shape(theta.shape, theta.strides[0] // 8, lay, P)
        if out is None:
            out = np.empty(n, dtype=np.float64)
        elif (not isinstance(out, np.ndarray) or out.dtype != np.float64 or out.shape != (n,)
              or not out.flags['C_CONTIGUOUS'] or not out.flags['WRITEABLE']):
            raise TypeError('out must be a writable C-contiguous float64 array of shape (%d,)' % n)
        target = self._multi(keys) if len(self.devices) > 1 else ctx
        target.loglike_host_ptr(theta.ctypes.data, n, ld, lay, mode, plan, out.ctypes.data)
        return out

    def log_likelihood_sharded(self, thetas, out=None, keys=None, mode=None, plan=None):
        """Device-resident slices, one (n_g, P) float64 CUDA tensor per listed device (in device
        order; empty slices allowed), evaluated where they live; the (sum n_g,) result is gathered
        with peer copies into `out` (default: a new tensor on the first device).  The plan comes
        from the total batch size, so the bits equal those of one GPU evaluating the whole batch."""
        import torch
        if len(thetas) != len(self.devices):
            raise ValueError('one tensor per device is needed (%d devices)' % len(self.devices))
        multi = self._multi(keys)
        P = multi.problem.n_params
        mode = self.mode if mode is None else _lib.MODES[mode]
        plan = self.plan if plan is None else int(plan)
        ptrs, counts, keep = [], [], []
        for t, d in zip(thetas, self.devices):
            if (not _is_torch_tensor(t) or not t.is_cuda or t.device.index != d or t.dtype != torch.float64
                    or t.dim() != 2 or t.shape[1] != P):
                raise TypeError('slice for cuda:%d must be an (n, %d) float64 tensor on that device' % (d, P))
            t = t.contiguous()
            keep.append(t)
            ptrs.append(t.data_ptr() if t.shape[0] else 0)
            counts.append(int(t.shape[0]))
        total = sum(counts)
        if out is None:
            out = torch.empty(total, dtype=torch.float64, device=torch.device('cuda', self.devices[0]))
        elif (not _is_torch_tensor(out) or not out.is_cuda or out.dtype != torch.float64 or out.numel() != total
              or not out.is_contiguous()):
            raise TypeError('out must be a contiguous float64 CUDA tensor of %d elements' % total)
        for d in set(self.devices) | {out.device.index}:
            torch.cuda.synchronize(d)                     # producers of the slices / earlier users of out
        multi.loglike_dev_ptrs(ptrs, counts, P, _lib.LAYOUT_AOS, mode, plan, out.data_ptr(), out.device.index)
        return out

    def log_posterior_batch(self, theta, prior, out=None, mode=None, plan=None):
        """log prior + log L of N vectors in ONE launch (stn_logpost); -inf outside the prior's
        support.  theta: (N, P) float64 CUDA tensor whose columns follow prior.keys
        (sampler.BoxPrior).  Returns a CUDA tensor."""
        import torch
        ctx = self._context(prior.keys)
        if ctx.prior is not prior:
            ctx.set_prior(prior)
        if not _is_torch_tensor(theta) or not theta.is_cuda:
            raise TypeError('log_posterior_batch takes a CUDA tensor (samplers on the device); use '
                            'sampler.log_prob_fn for host arrays')
        if theta.dtype != torch.float64 or theta.dim() != 2:
            raise TypeError('theta must be a 2-D float64 tensor')
        if theta.stride(1) != 1:
            theta = theta.contiguous()
        mode = self.mode if mode is None else _lib.MODES[mode]
        plan = self.plan if plan is None else int(plan)
        n, ld = self._shape(theta.shape, theta.stride(0), _lib.LAYOUT_AOS, ctx.problem.n_params)
        if out is None:
            out = torch.empty(n, dtype=torch.float64, device=theta.device)
        elif (not _is_torch_tensor(out) or out.dtype != torch.float64 or out.device != theta.device
              or out.numel() != n or not out.is_contiguous()):
            raise TypeError('out must be a contiguous float64 tensor of %d elements on %s' % (n, theta.device))
        stream = torch.cuda.current_stream(theta.device).cuda_stream
        ctx.logpost_ptr(theta.data_ptr(), n, ld, _lib.LAYOUT_AOS, mode, plan, out.data_ptr(), stream)
        return out

    def chi_square_batch(self, theta, keys=None, mode=None, plan=None):
        """sum((obs - model)/sigma)^2 over all sources and epochs for N vectors (simulate.py:297)."""
        import torch
        ctx = self._context(keys)
        dev = torch.device('cuda', self.device)
        was_numpy = not _is_torch_tensor(theta)
        th = torch.as_tensor(np.asarray(theta, dtype=np.float64) if was_numpy else theta).to(dev, torch.float64).contiguous()
        n, ld = self._shape(th.shape, th.stride(0), _lib.LAYOUT_AOS, ctx.problem.n_params)
        out = torch.empty(n, dtype=torch.float64, device=dev)
        mode = self.mode if mode is None else _lib.MODES[mode]
        plan = self.plan if plan is None else int(plan)
        ctx.chisq_ptr(th.data_ptr(), n, ld, _lib.LAYOUT_AOS, mode, plan, out.data_ptr(),
                      torch.cuda.current_stream(dev).cuda_stream)
        return out.cpu().numpy() if was_numpy else out

    def calculate_reduced_chi_square(self, dict_median):
        """(chi_sq, reduced chi_sq) at one parameter dict, as simulate.calculate_reduced_chi_square
        (simulate.py:290-301): DoF = 2 * total epochs - number of parameters."""
        names = self.parameter_names
        theta = np.array([[float(dict_median[k]) for k in names]], dtype=np.float64)
        chi_sq = float(self.chi_square_batch(theta)[0])
        n_obs = 2 * sum(len(v['epochs']) for v in self.LoD_VLBI)
        return chi_sq, chi_sq / (n_obs - len(dict_median))

    @staticmethod
    def _shape(shape, stride0, lay, P):
        if lay == _lib.LAYOUT_AOS:
            if shape[1] != P:
                raise ValueError('theta has %d columns, the model has %d parameters' % (shape[1], P))
            return int(shape[0]), int(stride0) if shape[0] > 1 else max(int(stride0), P)
        if shape[0] != P:
            raise ValueError('theta has %d rows, the model has %d parameters' % (shape[0], P))
        return int(shape[1]), int(stride0) if shape[0] > 1 else max(int(stride0), int(shape[1]))

    def model_positions(self, theta, source, keys=None, mode=None, offsets=False):
        """Model RA/Dec of one source for N parameter vectors: (N, 2E) = [ra..., dec...] rad
        (positions.py:13-30); with offsets=True also the (N, 2E) offsets in mas."""
        import torch
        ctx = self._context(keys)
        mode = self.mode if mode is None else _lib.MODES[mode]
        dev = torch.device('cuda', self.device)
        was_numpy = not _is_torch_tensor(theta)
        th = torch.as_tensor(np.asarray(theta, dtype=np.float64) if was_numpy else theta)
        th = th.to(dev, torch.float64).contiguous()
        n, P = th.shape
        if P != ctx.problem.n_params:
            raise ValueError('theta has %d columns, the model has %d parameters' % (P, ctx.problem.n_params))
        E = ctx.problem.sources[source]['n_epochs']
        radec = torch.empty((n, 2 * E), dtype=torch.float64, device=dev)
        off = torch.empty((n, 2 * E), dtype=torch.float64, device=dev) if offsets else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        ctx.model_ptr(source, th.data_ptr(), n, P, _lib.LAYOUT_AOS, mode, radec.data_ptr(),
                      off.data_ptr() if offsets else None, stream)
        if was_numpy:
            radec = radec.cpu().numpy()
            off = off.cpu().numpy() if offsets else None
        return (radec, off) if offsets else radec

    # bilby pickles the likelihood (resume files, npool > 1): device handles are dropped and the
    # contexts are rebuilt lazily in whichever process touches the object next
    def __getstate__(self):
        state = dict(self.__dict__)
        state['_contexts'] = {}
        state['_multis'] = {}
        state['_one'] = None
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)

    def close(self):
        seen = set()
        for ctx in self._contexts.values():
            if id(ctx) not in seen:
                seen.add(id(ctx))
                ctx.close()
        self._contexts = {}
        self._one = None
        for m in self._multis.values():
            m.close()
        self._multis = {}


# ------------------------------------------------------------------------------------
# positions(): the reference's stateless model function (positions.py:13-30)
# ------------------------------------------------------------------------------------
_positions_cache = collections.OrderedDict()
_POSITIONS_CACHE_SIZE = 16


def positions(refepoch, epochs, dict_timing, parameter_filter_index, dict_parameters, *,
              earth_provider=None, device=0, mode='strict'):
    """Model positions [ra(E)..., dec(E)...] in rad for one source.  The parameters of
    source `parameter_filter_index` are picked from `dict_parameters` exactly as the
    reference does (missing roots = -999).  Device contexts are cached per
    (refepoch, epochs, parfile) so repeated calls only launch the model kernel."""
    from .readers import plain_timing
    dict_timing = plain_timing(dict_timing)
    epochs = np.ascontiguousarray(epochs, dtype=np.float64)
    FP = filter_dictionary_of_parameter_with_index(dict_parameters, parameter_filter_index)
    keys = sorted(FP.keys())
    vals = [float(FP[k]) for k in keys]          # dec, efac, efad, incl, mu_a, mu_d, om_asc, px, ra
    present = [v != SENTINEL for v in vals]
    tkey = tuple(sorted((k, v) for k, v in dict_timing.items())) if dict_timing != {} else ()
    # the provider object itself is part of the key (kept alive by the cache), not its id()
    key = (float(refepoch), epochs.tobytes(), tkey, tuple(present), int(device), earth_provider)
    like = _positions_cache.get(key)
    if like is None:
        E = epochs.shape[0]
        shares = [[0] if on else [-1] for on in present]
        shares[1], shares[2] = [-1], [-1]        # EFAC plays no role in the model
        vlbi = {'epochs': epochs, 'radecs': np.zeros(2 * E), 'errs': np.ones(2 * E)}
        like = Gaussianlikelihood(refepoch, [dict_timing], [vlbi], shares, None, None,
                                  earth_provider=earth_provider, device=device, mode=mode)
        _positions_cache[key] = like
        while len(_positions_cache) > _POSITIONS_CACHE_SIZE:
            _positions_cache.popitem(last=False)[1].close()
    else:
        _positions_cache.move_to_end(key)
    by_root = {_root_of(k): v for k, v in zip(keys, vals)}
    theta = np.array([[by_root[_root_of(n)] for n in like.parameter_names]], dtype=np.float64)
    return like.model_positions(theta, 0, mode=mode)[0]


def positions_batch(refepoch, epochs, dict_timing, parameter_filter_index, names, samples, *,
                    earth_provider=None, device=0, mode='fast'):
    """positions() for MANY parameter vectors at once: the model curves the reference's sky plots
    draw by calling positions() once per posterior draw on a dense epoch grid
    (plot/sky_position_evolution.py:104-129, positions.py:240-264).

    names / samples: column names and (n, len(names)) values (e.g. posterior draws); the
    parameters of source `parameter_filter_index` are routed exactly as positions() routes a
    dict.  Returns (n, 2E) = [ra(E)..., dec(E)...] in rad."""
    from .readers import plain_timing
    from .shares import column_table
    samples = np.ascontiguousarray(samples, dtype=np.float64)
    names = list(names)
    epochs = np.ascontiguousarray(epochs, dtype=np.float64)
    from .shares import parameter_name_to_pmparin_indice
    n_sources = 1 + max([parameter_filter_index] + [i for n in names for i in parameter_name_to_pmparin_indice(n)[0]])
    idx = column_table(names, n_sources)[parameter_filter_index]
    present = [bool(c >= 0) for c in idx]
    shares = [[0] if on else [-1] for on in present]
    shares[1], shares[2] = [-1], [-1]
    E = epochs.shape[0]
    vlbi = {'epochs': epochs, 'radecs': np.zeros(2 * E), 'errs': np.ones(2 * E)}
    like = Gaussianlikelihood(refepoch, [plain_timing(dict_timing)], [vlbi], shares, None, None,
                              earth_provider=earth_provider, device=device, mode=mode)
    try:
        cols = [int(idx[ROOTS.index(_root_of(n))]) for n in like.parameter_names]
        return like.model_positions(samples[:, cols], 0, mode=mode)
    finally:
        like.close()


def _root_of(name):
    return '_'.join(t for t in name.split('_') if not t.isdigit())
