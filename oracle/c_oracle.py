"""ctypes front end of the C oracle (oracle/c/sterne_oracle.c) -- TEST / BENCH INFRASTRUCTURE ONLY.

Builds the per-epoch input rows with the numpy oracle's own functions (no product code) and
evaluates batches on all host cores.  Used by tests as a second, independent CPU checker at
batch sizes the numpy restatement is too slow for, and by bench.py as the CPU baseline.
"""
import ctypes
import os
import subprocess

import numpy as np

from oracle import sterne_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_build', 'libsterne_oracle.so')


def build(force=False):
    src = os.path.join(HERE, 'c', 'sterne_oracle.c')
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(['make', '-s', '-B', '-C', os.path.join(HERE, 'c')])
    return LIB


class OrcSource(ctypes.Structure):
    _fields_ = [('n_epochs', ctypes.c_int64), ('rows', ctypes.POINTER(ctypes.c_double)),
                ('col', ctypes.c_int32 * 9), ('efac_on', ctypes.c_int32), ('binary', ctypes.c_int32),
                ('cos_decj', ctypes.c_double), ('a1_lts', ctypes.c_double)]


class COracle(object):
    def __init__(self, refepoch, LoD_timing, LoD_VLBI, shares, earth_xyz, DoD_additional_constraints=None,
                 a1dot_constraints=False):
        self.lib = ctypes.CDLL(build())
        self.lib.orc_loglike_batch.restype = ctypes.c_int
        self.names, idx = O.column_table(shares)
        S = len(LoD_VLBI)
        self.src = (OrcSource * S)()
        self._keep = []
        for s in range(S):
            v, t = LoD_VLBI[s], LoD_timing[s]
            E = len(v['epochs'])
            rows = np.zeros((E, 16))
            rows[:, 0] = (np.asarray(v['epochs'], dtype=np.float64) - refepoch) * O.D2YR     # positions.py:108
            rows[:, 1:4] = np.asarray(earth_xyz[s]).reshape(E, 3)
            rows[:, 4], rows[:, 5] = v['radecs'][:E], v['radecs'][E:]
            rows[:, 6], rows[:, 7] = v['errs'][:E], v['errs'][E:]
            efac_on = idx[s][1] >= 0
            if efac_on:
                rows[:, 8], rows[:, 9] = v['errs_random'][:E], v['errs_random'][E:]
                rows[:, 10], rows[:, 11] = v['errs_sys'][:E], v['errs_sys'][E:]
            if t != {}:
                for i, ep in enumerate(v['epochs']):
                    rows[i, 12:15] = O.reflex_orbit_terms(ep, t)
            rows = np.ascontiguousarray(rows)
            self._keep.append(rows)
            q = self.src[s]
            q.n_epochs = E
            q.rows = rows.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
            for r in range(9):
                q.col[r] = int(idx[s][r])
            q.efac_on = int(efac_on)
            q.binary = int(t != {})
            q.cos_decj = float(np.cos(t['decj'] * O.DEG2RAD)) if t != {} else 1.0
            q.a1_lts = float(t['a1'] / O.C_M_S) if t != {} else 0.0
        self.S = S
        self.P = len(self.names)
        self.a1_src = np.zeros(0, dtype=np.int32)
        self.a1_mu = self.a1_sig = np.zeros(0)
        if a1dot_constraints is not False and a1dot_constraints is not None:
            on, mus, sig = O.parse_a1dot_constraints(a1dot_constraints)
            if on:
                bins = [i for i, t in enumerate(LoD_timing) if t != {}]
                nb, nm = len(bins), len(mus)
                K = max(nb, nm)                                      # numpy broadcasting of (nb,) - (nm,)
                self.a1_src = np.array([bins[k if nb > 1 else 0] for k in range(K)], dtype=np.int32)
                self.a1_mu = np.array([mus[k if nm > 1 else 0] for k in range(K)], dtype=np.float64)
                self.a1_sig = np.array([sig[k if nm > 1 else 0] for k in range(K)], dtype=np.float64)
        gauss = {} if DoD_additional_constraints is None else DoD_additional_constraints['sin_incl_Gaussian_constraints']
        self.sini_col = np.array([self.names.index(k) for k in gauss], dtype=np.int32)
        self.sini_mu = np.array([float(v[0]) for v in gauss.values()], dtype=np.float64)
        self.sini_sig = np.array([float(v[1]) for v in gauss.values()], dtype=np.float64)

    def log_likelihood_batch(self, theta, nthreads=None):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        N, P = theta.shape
        assert P == self.P
        out = np.empty(N)
        dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
        rc = self.lib.orc_loglike_batch(self.src, self.S, dp(theta), ctypes.c_int64(N), ctypes.c_int64(P),
                                        len(self.a1_src), ip(self.a1_src), dp(self.a1_mu), dp(self.a1_sig),
                                        len(self.sini_col), ip(self.sini_col), dp(self.sini_mu), dp(self.sini_sig),
                                        dp(out), int(nthreads or os.cpu_count() or 1))
        if rc != 0:
            raise MemoryError('C oracle could not allocate scratch')
        return out


def from_case(case):
    return COracle(case.refepoch, case.LoD_timing, case.LoD_VLBI, case.shares, case.earth_xyz,
                   case.DoD_additional_constraints, case.a1dot_constraints)
