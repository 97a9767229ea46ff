"""Problem compiler: turns the reference's Python inputs (list_of_dict_VLBI,
list_of_dict_timing, shares, constraints -- simulate.py:326-358) into the flat epoch
tables, routing table and scalars the device needs (include/sterne_b200.h).

Everything here is parameter-independent and runs once per problem on the host.
"""
import ctypes
import numpy as np

from . import _lib, constants
from .constants import ROOTS, D2YR, DEG2RAD, C_M_S
from .ephemeris import earth_vectors
from .orbit import orbit_terms
from .readers import plain_timing
from .shares import get_parameters_from_shares, column_table

R_DEC, R_EFAC, R_EFAD, R_INCL, R_MU_A, R_MU_D, R_OM_ASC, R_PX, R_RA = range(9)


def parse_a1dot_constraints(a1dot_constraints):
    """simulate.py:387-399: entries of length 2 are (mu, sigma) in lt-s/s."""
    mus, sigmas = [], []
    for c in a1dot_constraints:
        if len(c) == 2:
            mus.append(float(c[0]))
            sigmas.append(float(c[1]))
    return mus, sigmas


class CompiledProblem(object):
    """Host image of one StnProblem.  Holds the numpy arrays alive for as long as
    the ctypes view is in use."""

    def __init__(self, refepoch, list_of_dict_timing, list_of_dict_VLBI, shares, earth_provider=None,
                 DoD_additional_constraints=None, a1dot_constraints=False, names=None, earth_xyz=None):
        S = len(list_of_dict_VLBI)
        if len(list_of_dict_timing) != S:
            raise ValueError('Unequal number of parfiles provided for pmparins.')
        if any(len(row) != S for row in shares):
            raise ValueError('every row of shares must have one entry per pmparin (%d)' % S)
        self.refepoch = refepoch
        self.n_sources = S
        list_of_dict_timing = [plain_timing(t) for t in list_of_dict_timing]
        self.default_names = list(get_parameters_from_shares(shares))
        self.names = list(self.default_names) if names is None else list(names)
        if sorted(self.names) != sorted(self.default_names):
            raise ValueError('keys %r do not match the parameters defined by shares %r'
                             % (self.names, self.default_names))
        self.n_params = len(self.names)
        self.idx = column_table(self.names, S)
        self.sources = []
        for s in range(S):
            self.sources.append(self._compile_source(s, list_of_dict_VLBI[s], list_of_dict_timing[s],
                                                     earth_provider, None if earth_xyz is None else earth_xyz[s]))
        self._compile_constraints(list_of_dict_timing, DoD_additional_constraints, a1dot_constraints)

    # ------------------------------------------------------------------ sources
    def _compile_source(self, s, vlbi, timing, earth_provider, xyz):
        idx = self.idx[s]
        epochs = np.asarray(vlbi['epochs'], dtype=np.float64)
        E = epochs.shape[0]
        radecs = np.asarray(vlbi['radecs'], dtype=np.float64)
        errs = np.asarray(vlbi['errs'], dtype=np.float64)
        if radecs.shape[0] != 2 * E or errs.shape[0] != 2 * E:
            raise ValueError('source %d: radecs/errs must hold 2*len(epochs) values' % s)
        if xyz is None:
            if idx[R_PX] >= 0:
                if earth_provider is None:
                    raise ValueError('source %d fits a parallax: an earth_provider (or earth_xyz) is required' % s)
                xyz = earth_vectors(earth_provider, epochs)
            else:
                xyz = np.zeros((E, 3))
        xyz = np.asarray(xyz, dtype=np.float64).reshape(E, 3)
        efac_on = bool(idx[R_EFAC] >= 0)
        binary = (timing != {})
        fast = np.zeros((E, _lib.FAST_ROW), dtype=np.float64)
        strict = np.zeros((E, _lib.STRICT_ROW), dtype=np.float64)
        dT = (epochs - self.refepoch) * D2YR                       # positions.py:108
        for tab in (fast, strict):
            tab[:, 0] = dT
            tab[:, 1:4] = xyz
            tab[:, 4] = radecs[:E]
            tab[:, 5] = radecs[E:]
        neg_sum_log = 0.0
        if efac_on:
            rnd = np.asarray(vlbi['errs_random'], dtype=np.float64)   # KeyError as in simulate.py:317
            sys_ = np.asarray(vlbi['errs_sys'], dtype=np.float64)
            fast[:, 6] = rnd[:E] ** 2
            fast[:, 7] = rnd[E:] ** 2
            fast[:, 8] = sys_[:E] ** 2
            fast[:, 9] = sys_[E:] ** 2
            strict[:, 8], strict[:, 9] = rnd[:E], rnd[E:]
            strict[:, 10], strict[:, 11] = sys_[:E], sys_[E:]
            # a sampled efac of exactly -999 falls back to 'errs' (simulate.py:318-320)
            errs_off = (errs ** 2) ** 0.5
            strict[:, 6], strict[:, 7] = errs_off[:E], errs_off[E:]
        else:
            errs_new = (errs ** 2) ** 0.5                            # simulate.py:319-320
            fast[:, 6] = 1.0 / errs_new[:E]
            fast[:, 7] = 1.0 / errs_new[E:]
            strict[:, 6], strict[:, 7] = errs_new[:E], errs_new[E:]
            neg_sum_log = float(-1 * np.sum(np.log(errs_new)))       # simulate.py:371
        cos_decj, a1_lts = 1.0, 0.0
        if binary:
            terms = orbit_terms(epochs, timing)
            fast[:, 10] = terms[:, 0] * terms[:, 1]
            fast[:, 11] = terms[:, 0] * terms[:, 2]
            strict[:, 12:15] = terms
            cos_decj = float(np.cos(timing['decj'] * DEG2RAD))       # reflex_motion.py:212
            a1_lts = float(timing['a1'] / C_M_S)                     # kopeikin_effects.py:61
        return dict(n_epochs=E, fast=np.ascontiguousarray(fast), strict=np.ascontiguousarray(strict),
                    col=idx.copy(), efac_on=efac_on, binary=binary, cos_decj=cos_decj,
                    neg_sum_log_err=neg_sum_log, a1_lts=a1_lts, earth_xyz=xyz)

    # -------------------------------------------------------------- constraints
    def _compile_constraints(self, list_of_dict_timing, DoD, a1dot_constraints):
        self.a1_source = np.zeros(0, dtype=np.int32)
        self.a1_mu = np.zeros(0, dtype=np.float64)
        self.a1_sigma = np.zeros(0, dtype=np.float64)
        if a1dot_constraints is not False and a1dot_constraints is not None:
            mus, sigmas = parse_a1dot_constraints(a1dot_constraints)
            if len(mus) != 0:
                binaries = [i for i, t in enumerate(list_of_dict_timing) if t != {}]
                if len(binaries) == 0:
                    raise ValueError("The list_of_modeled_a1dot is empty. Make sure not all parfiles are ''.")
                # modeled[nb] - mus[nm] follows numpy broadcasting in the reference (simulate.py:376)
                nb, nm = len(binaries), len(mus)
                if nb != nm and nb != 1 and nm != 1:
                    raise ValueError('operands could not be broadcast together with shapes (%d,) (%d,)' % (nb, nm))
                K = max(nb, nm)
                self.a1_source = np.array([binaries[k if nb > 1 else 0] for k in range(K)], dtype=np.int32)
                self.a1_mu = np.array([mus[k if nm > 1 else 0] for k in range(K)], dtype=np.float64)
                self.a1_sigma = np.array([sigmas[k if nm > 1 else 0] for k in range(K)], dtype=np.float64)
        gauss = {} if DoD is None else DoD.get('sin_incl_Gaussian_constraints', {})
        cols, mus, sigmas = [], [], []
        for name, (mu, sigma) in gauss.items():
            if name not in self.names:
                raise KeyError(name)
            cols.append(self.names.index(name))
            mus.append(float(mu))
            sigmas.append(float(sigma))
        self.sini_col = np.array(cols, dtype=np.int32)
        self.sini_mu = np.array(mus, dtype=np.float64)
        self.sini_sigma = np.array(sigmas, dtype=np.float64)

    # ------------------------------------------------------------------- ctypes
    def as_ctypes(self):
        """(StnProblem, keepalive) -- keepalive must outlive the stn_create call."""
        arr = (_lib.StnSource * self.n_sources)()
        for s, src in enumerate(self.sources):
            a = arr[s]
            a.n_epochs = src['n_epochs']
            a.fast_rows = src['fast'].ctypes.data_as(_lib.c_double_p)
            a.strict_rows = src['strict'].ctypes.data_as(_lib.c_double_p)
            for r in range(len(ROOTS)):
                a.col[r] = int(src['col'][r])
            a.efac_on = int(src['efac_on'])
            a.binary = int(src['binary'])
            a.cos_decj = src['cos_decj']
            a.neg_sum_log_err = src['neg_sum_log_err']
            a.a1_lts = src['a1_lts']
        pb = _lib.StnProblem()
        pb.n_sources = self.n_sources
        pb.n_params = self.n_params
        pb.sources = arr
        pb.n_a1dot = len(self.a1_source)
        pb.a1dot_source = self.a1_source.ctypes.data_as(_lib.c_int32_p)
        pb.a1dot_mu = self.a1_mu.ctypes.data_as(_lib.c_double_p)
        pb.a1dot_sigma = self.a1_sigma.ctypes.data_as(_lib.c_double_p)
        pb.n_sini = len(self.sini_col)
        pb.sini_col = self.sini_col.ctypes.data_as(_lib.c_int32_p)
        pb.sini_mu = self.sini_mu.ctypes.data_as(_lib.c_double_p)
        pb.sini_sigma = self.sini_sigma.ctypes.data_as(_lib.c_double_p)
        # the factors this host resolved (astropy's own when astropy is importable, constants.py)
        pb.mas2rad, pb.deg2rad, pb.masyr2rads = constants.MAS2RAD, constants.DEG2RAD, constants.MASYR2RADS
        return pb, (arr, self)

    @property
    def evals_per_sample(self):
        """One eval = one (sample, source, epoch) (BASELINE.json metric)."""
        return int(sum(src['n_epochs'] for src in self.sources))


def prior_as_ctypes(prior, names):
    """(StnPrior, keepalive) of a sampler.BoxPrior whose keys are the context's column order."""
    if list(prior.keys) != list(names):
        raise ValueError('prior keys %r differ from the context column order %r' % (prior.keys, names))
    kind = np.array([_lib.PRIOR_KINDS[k] for k in prior.kinds], dtype=np.int32)
    a = np.ascontiguousarray(prior.a, dtype=np.float64)
    b = np.ascontiguousarray(prior.b, dtype=np.float64)
    col = np.array([c for c, _, _ in prior.sin_limits], dtype=np.int32)
    lo = np.array([l for _, l, _ in prior.sin_limits], dtype=np.float64)
    hi = np.array([h for _, _, h in prior.sin_limits], dtype=np.float64)
    pr = _lib.StnPrior()
    pr.n_params = len(kind)
    pr.kind = kind.ctypes.data_as(_lib.c_int32_p)
    pr.a = a.ctypes.data_as(_lib.c_double_p)
    pr.b = b.ctypes.data_as(_lib.c_double_p)
    pr.n_sin_limits = len(col)
    pr.sin_col = col.ctypes.data_as(_lib.c_int32_p)
    pr.sin_lo = lo.ctypes.data_as(_lib.c_double_p)
    pr.sin_hi = hi.ctypes.data_as(_lib.c_double_p)
    return pr, (kind, a, b, col, lo, hi)


class DeviceContext(object):
    """Owns one StnCtx (device copies of a CompiledProblem)."""

    def __init__(self, problem, device=0):
        self.lib = _lib.load()
        self.problem = problem
        self.device = int(device)
        pb, keep = problem.as_ctypes()
        handle = ctypes.c_void_p()
        _lib.check(self.lib.stn_create(ctypes.byref(pb), self.device, ctypes.byref(handle)))
        del keep
        self.handle = handle
        self.prior = None

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.stn_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- raw entry points (pointers are integers: tensor.data_ptr() / ndarray.ctypes.data)
    def loglike_ptr(self, theta_ptr, n, ld, layout, mode, plan, out_ptr, stream=0):
        _lib.check(self.lib.stn_loglike(self.handle, theta_ptr, n, ld, layout, mode, plan, out_ptr, stream))

    def loglike_scatter_ptr(self, theta_ptr, n, ld, layout, mode, plan, out_ptrs, stream=0):
        """stn_loglike_scatter: the result is stored to every buffer of out_ptrs (the first one on
        this device, the others typically peer GPUs' buffers)."""
        k = len(out_ptrs)
        arr = (ctypes.c_void_p * k)(*[ctypes.c_void_p(int(p)) for p in out_ptrs])
        _lib.check(self.lib.stn_loglike_scatter(self.handle, theta_ptr, n, ld, layout, mode, plan, arr, k, stream))

    def set_prior(self, prior):
        """prior: sampler.BoxPrior whose keys are this context's column order."""
        pr, keep = prior_as_ctypes(prior, self.problem.names)
        _lib.check(self.lib.stn_set_prior(self.handle, ctypes.byref(pr)))
        del keep
        self.prior = prior

    def logpost_ptr(self, theta_ptr, n, ld, layout, mode, plan, out_ptr, stream=0):
        _lib.check(self.lib.stn_logpost(self.handle, theta_ptr, n, ld, layout, mode, plan, out_ptr, stream))

    def chisq_ptr(self, theta_ptr, n, ld, layout, mode, plan, out_ptr, stream=0):
        _lib.check(self.lib.stn_chisq(self.handle, theta_ptr, n, ld, layout, mode, plan, out_ptr, stream))

    def loglike_host_ptr(self, theta_ptr, n, ld, layout, mode, plan, out_ptr):
        _lib.check(self.lib.stn_loglike_host(self.handle, theta_ptr, n, ld, layout, mode, plan, out_ptr))

    def model_ptr(self, source, theta_ptr, n, ld, layout, mode, radec_ptr, off_ptr, stream=0):
        _lib.check(self.lib.stn_model(self.handle, source, theta_ptr, n, ld, layout, mode, radec_ptr,
                                      off_ptr, stream))

    def host_hint(self, concurrent_devices):
        """How many devices copy from this host at once (one process per GPU: the world size)."""
        _lib.check(self.lib.stn_host_hint(self.handle, int(concurrent_devices)))

    def auto_plan(self, n):
        """Execution plan the library would choose for a batch of n (lanes | splits << 8)."""
        return int(self.lib.stn_auto_plan(self.handle, n))

    def launch_count(self):
        return int(self.lib.stn_launch_count(self.handle))

    def loglike_timed(self, theta_ptr, n, ld, layout, mode, plan, out_ptr, steps):
        ms = ctypes.c_double()
        _lib.check(self.lib.stn_loglike_timed(self.handle, theta_ptr, n, ld, layout, mode, plan, out_ptr,
                                              steps, ctypes.byref(ms)))
        return ms.value


class MultiDeviceContext(object):
    """Owns one StnMulti: the same CompiledProblem on several GPUs of this box, driven from this
    one process (include/sterne_b200.h, stn_multi_*).  Host batches are cut into contiguous slices,
    every device copies its slice of the result straight into the caller's array."""

    def __init__(self, problem, devices):
        self.lib = _lib.load()
        self.problem = problem
        self.devices = [int(d) for d in devices]
        if not self.devices:
            raise ValueError('at least one device is needed')
        pb, keep = problem.as_ctypes()
        handle = ctypes.c_void_p()
        arr = (ctypes.c_int32 * len(self.devices))(*self.devices)
        _lib.check(self.lib.stn_multi_create(ctypes.byref(pb), arr, len(self.devices), ctypes.byref(handle)))
        del keep
        self.handle = handle
        self.prior = None

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.stn_multi_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return len(self.devices)

    def auto_plan(self, n):
        """Plan for a batch of n, always taken from the WHOLE batch (device 0's context decides)."""
        return int(self.lib.stn_auto_plan(self.lib.stn_multi_ctx(self.handle, 0), n))

    def launch_count(self):
        return int(self.lib.stn_multi_launch_count(self.handle))

    def slice(self, n, part, n_parts=None):
        lo, hi = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(self.lib.stn_multi_slice(n, len(self.devices) if n_parts is None else n_parts, part,
                                            ctypes.byref(lo), ctypes.byref(hi)))
        return lo.value, hi.value

    def loglike_host_ptr(self, theta_ptr, n, ld, layout, mode, plan, out_ptr):
        _lib.check(self.lib.stn_multi_loglike_host(self.handle, theta_ptr, n, ld, layout, mode, plan, out_ptr))

    def set_prior(self, prior):
        pr, keep = prior_as_ctypes(prior, self.problem.names)
        _lib.check(self.lib.stn_multi_set_prior(self.handle, ctypes.byref(pr)))
        del keep
        self.prior = prior

    def ensemble_run(self, x_ptrs, lp_ptrs, n_walkers, ld, a, seed, step0, n_steps, thin, chain_ptr, lnprob_ptr,
                     mode, plan):
        """stn_multi_ensemble_run; returns the number of acceptances of this call."""
        G = len(self.devices)
        xs = (ctypes.c_void_p * G)(*[ctypes.c_void_p(int(p)) for p in x_ptrs])
        ls = (ctypes.c_void_p * G)(*[ctypes.c_void_p(int(p)) for p in lp_ptrs])
        nacc = ctypes.c_int64(0)
        _lib.check(self.lib.stn_multi_ensemble_run(self.handle, xs, ls, n_walkers, ld, a, seed, step0, n_steps, thin,
                                                   chain_ptr, lnprob_ptr, ctypes.byref(nacc), mode, plan))
        return nacc.value

    def loglike_dev_ptrs(self, theta_ptrs, counts, ld, layout, mode, plan, out_ptr, out_device):
        G = len(self.devices)
        if len(theta_ptrs) != G or len(counts) != G:
            raise ValueError('one slice per device is needed (%d devices)' % G)
        tp = (ctypes.c_void_p * G)(*[ctypes.c_void_p(int(p) if p else 0) for p in theta_ptrs])
        cn = (ctypes.c_int64 * G)(*[int(c) for c in counts])
        _lib.check(self.lib.stn_multi_loglike_dev(self.handle, tp, cn, ld, layout, mode, plan, out_ptr,
                                                  int(out_device)))


def host_register(array):
    """Pins a numpy array in place (cudaHostRegister) so that every GPU can DMA from / into it;
    returns a handle whose .release() (or garbage collection) unpins it.  Worth it for buffers a
    sampler reuses: a pageable buffer is staged through the library's pinned ring instead."""
    lib = _lib.load()

    class _Pinned(object):
        def __init__(self, a):
            self.array = a
            self.ptr = a.ctypes.data
            _lib.check(lib.stn_host_register(self.ptr, a.nbytes))

        def release(self):
            if self.ptr:
                lib.stn_host_unregister(self.ptr)
                self.ptr = 0

        def __del__(self):
            try:
                self.release()
            except Exception:
                pass

    return _Pinned(array)


def fp64_peak(device=0, iters=1 << 16, reps=5):
    """Measured DFMA peak of `device` in TFLOP/s (the FP64 roofline denominator)."""
    lib = _lib.load()
    tf, ms = ctypes.c_double(), ctypes.c_double()
    _lib.check(lib.stn_fp64_peak(device, iters, reps, ctypes.byref(tf), ctypes.byref(ms)))
    return tf.value, ms.value
