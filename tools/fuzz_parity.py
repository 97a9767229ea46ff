#!/usr/bin/env python
"""Randomised parity sweep: random problem shapes / shares / sentinels / constraints / plans,
GPU (fast and strict) against the C oracle.  usage: python tools/fuzz_parity.py [n_problems] [seed]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import c_oracle
from sterne_b200 import synthetic, _lib
from sterne_b200.shares import get_parameters_from_shares, parameter_name_to_pmparin_indice


def random_case(rng):
    S = int(rng.integers(1, 7))
    T = synthetic.TRUTH
    b = dict(synthetic.PARFILE_CFG3)
    b.update(pb=float(rng.uniform(1, 400)), ecc=float(rng.uniform(0, 0.7)), a1=float(rng.uniform(1, 80)),
             t0=56100.25, om=float(rng.uniform(0, 360)))
    if rng.random() < 0.5:
        b.update(omdot=0.0123)
    if rng.random() < 0.3:
        b.update(pbdot=1.2e-11, a1dot=3.4e-14)
    timing = synthetic.timing_dict(**b)
    efac = rng.random() < 0.6
    refepoch = 57365.0
    V, X, TM = [], [], []
    binary = [bool(rng.random() < 0.4) for _ in range(S)]
    for s in range(S):
        E = int(rng.choice([0, 1, 2, 3, 5, 17, 127, 128, 129, 300]))
        tm = timing if binary[s] else {}
        v, x = synthetic.make_source(rng, refepoch, 57000 + rng.uniform(0, 1500, E), T['ra'] + 3e-6 * s,
                                     T['dec'] - 2e-6 * s, T['mu_a'], T['mu_d'], T['px'], timing=tm or None,
                                     incl=T['incl'], om_asc=T['om_asc'], random_fraction=0.6 if efac else None)
        V.append(v); X.append(x); TM.append(tm)

    def row(kind):
        # random sharing pattern; kind decides how often a parameter is switched off
        r = []
        for s in range(S):
            off = {'always': 0.0, 'rare': 0.1, 'often': 0.5}[kind]
            r.append(-1 if rng.random() < off else int(rng.integers(0, max(1, S // 2 + 1))))
        return r
    shares = [row('rare'), row('often') if efac else [-1] * S, row('often') if efac else [-1] * S,
              [(-1 if (not binary[s] or rng.random() < 0.2) else 0) for s in range(S)], row('rare'), row('rare'),
              [(-1 if (not binary[s] or rng.random() < 0.2) else 0) for s in range(S)], row('rare'), row('rare')]
    names = list(get_parameters_from_shares(shares))
    if not names:
        return None
    truth = {}
    for k in names:
        root = parameter_name_to_pmparin_indice(k)[1]
        truth[k] = T[root]
    DoD = {'sin_incl_Gaussian_constraints': {}, 'sin_incl_limits_constraints': {}}
    incls = [k for k in names if k.startswith('incl')]
    if incls and rng.random() < 0.5:
        DoD['sin_incl_Gaussian_constraints'][incls[0]] = [0.9, 0.05]
    a1 = False
    nb = sum(1 for t in TM if t != {})
    if nb and rng.random() < 0.5:
        a1 = [[1.1e-14, 4e-15] if t != {} else [] for t in TM]
    return synthetic.SyntheticCase('fuzz', refepoch, TM, V, shares, X, truth, synthetic._box_for(names, truth), DoD, a1)


def main():
    n_problems = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    worst = 0.0
    done = 0
    while done < n_problems:
        case = random_case(rng)
        if case is None:
            continue
        N = int(rng.choice([1, 7, 33, 257, 1000]))
        theta = case.sample_theta(N, seed=int(rng.integers(1 << 30)))
        if rng.random() < 0.3:                      # value sentinels
            j = int(rng.integers(case.n_params))
            theta[::3, j] = -999.0
        want = c_oracle.from_case(case).log_likelihood_batch(theta)
        scale = max(1.0, float(sum(abs(np.sum(np.log(v['errs']))) for v in case.LoD_VLBI if len(v['errs']))))
        nchunks = sum(max(1, -(-len(v['epochs']) // 128)) for v in case.LoD_VLBI)
        for mode in ('fast', 'strict'):
            like = case.likelihood(mode=mode)
            plans = [0] if mode == 'strict' else [0, _lib.make_plan(int(rng.choice([1, 2, 4, 8, 16, 32])),
                                                                   int(rng.integers(1, nchunks + 1)))]
            for plan in plans:
                for layout in ('aos', 'soa'):
                    th = theta if layout == 'aos' else np.ascontiguousarray(theta.T)
                    got = like.log_likelihood_batch(th, layout=layout, plan=plan)
                    ok = np.isfinite(want)
                    if (np.isfinite(got) != ok).any():
                        print('FINITE MISMATCH', done, mode, plan, layout, case.shares)
                        return 1
                    err = (np.abs(got[ok] - want[ok]) / np.maximum(np.abs(want[ok]), scale)).max() if ok.any() else 0.0
                    worst = max(worst, err)
                    if err > 1e-10:
                        print('FAIL problem %d mode %s plan %d layout %s err %.3e shares %r epochs %r' %
                              (done, mode, plan, layout, err, case.shares, [len(v['epochs']) for v in case.LoD_VLBI]))
                        return 1
            like.close()
        done += 1
    print('fuzz ok: %d problems, worst scaled relative error %.3e' % (done, worst))
    return 0


if __name__ == '__main__':
    sys.exit(main())
