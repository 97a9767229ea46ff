#!/usr/bin/env python
"""Generates tests/golden/golden_reference.json by running the UNMODIFIED reference
(/root/reference/src/sterne) on seeded synthetic inputs.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

The reference's third-party imports (astropy, bilby, novas, psrqpy) are absent from
the image; oracle/ref_shims/ provides stand-ins (see its README).  Two adaptations,
both outside the reference's source:
  * numpy >= 2 removed the alias `np.mat` used at reflex_motion.py:202-208; it is
    restored as `np.mat = np.asmatrix` (what the alias always was);
  * the novas stand-in returns Earth vectors from sterne_b200's analytic orbit; the
    vectors are stored in the fixture, so tests never need the provider.
Everything recorded under "ref" is an output of reference code: readers, parameter
naming, positions(), adjust_errs_with_efac(), reflex_motion(), solve_u(),
calculate_a1dot_pm() and Gaussianlikelihood.log_likelihood().
"""
import io
import json
import os
import sys
import tempfile
import contextlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'ref_shims'))
sys.path.insert(0, '/root/reference/src')
os.environ.setdefault('TEMPO2', '/nonexistent-tempo2')

import numpy as np                                              # noqa: E402

np.mat = np.asmatrix

import warnings                                                 # noqa: E402
warnings.simplefilter('ignore')

from novas.compat import solsys                                 # noqa: E402  (stand-in)
from sterne import simulate as ref_simulate                     # noqa: E402  (REFERENCE)
from sterne import priors as ref_priors                         # noqa: E402
from sterne.model import positions as ref_positions             # noqa: E402
from sterne.model import reflex_motion as ref_reflex            # noqa: E402
from sterne.model import kopeikin_effects as ref_kopeikin       # noqa: E402

from sterne_b200 import synthetic                               # noqa: E402  (input generator only)
from sterne_b200.ephemeris import AnalyticEarthProvider         # noqa: E402

_analytic = AnalyticEarthProvider()
solsys.set_provider(lambda jd: _analytic(np.array([jd]))[0])


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def plain_timing(d):
    """astropy-unit dict (stand-in Quantities) -> plain floats in the Quantity's own unit."""
    return {k: float(v) for k, v in d.items()}


def sentinel_case():
    """2 sources; source 1 has px / mu_a / incl / om_asc switched off (mu_a = -999 is then
    USED AS A NUMBER by the reference, positions.py:59,110)."""
    rng = np.random.default_rng(11)
    T = synthetic.TRUTH
    b = dict(synthetic.PARFILE_CFG3)
    b.update(pb=175.46, ecc=0.3, a1=55.33, t0=56100.25, om=50.7)
    timing = synthetic.timing_dict(**b)
    refepoch = 57365.0
    v0, x0 = synthetic.make_source(rng, refepoch, 57000 + rng.uniform(0, 730, 7), T['ra'], T['dec'], T['mu_a'],
                                   T['mu_d'], T['px'], timing=timing, incl=T['incl'], om_asc=T['om_asc'])
    v1, x1 = synthetic.make_source(rng, refepoch, 57000 + rng.uniform(0, 730, 5), T['ra'] + 1e-6, T['dec'] + 1e-6,
                                   0.0, T['mu_d'], 0.0)
    shares = [[0, 1], [-1, -1], [-1, -1], [0, -1], [0, -1], [0, 0], [0, -1], [0, -1], [0, 1]]
    names = list(synthetic.get_parameters_from_shares(shares))
    truth = {}
    for k in names:
        srcs, root = synthetic.parameter_name_to_pmparin_indice(k)
        truth[k] = T[root] + (1e-6 if (root in ('ra', 'dec') and srcs == [1]) else 0.0)
    return synthetic.SyntheticCase('sentinels', refepoch, [timing, {}], [v0, v1], shares, [x0, x1], truth,
                                   synthetic._box_for(names, truth), parfile_text={0: synthetic.parfile_text(**b)})


def efad_case():
    case = synthetic.config2()
    shares = [list(r) for r in case.shares]
    shares[2] = [0]
    names = list(synthetic.get_parameters_from_shares(shares))
    truth = dict(case.truth)
    truth['efad_0'] = synthetic.TRUTH['efad']
    return synthetic.SyntheticCase('config2_efad', case.refepoch, case.LoD_timing, case.LoD_VLBI, shares,
                                   case.earth_xyz, truth, synthetic._box_for(names, truth))


def run_case(case, tmp, n_theta=4):
    files = synthetic.write_case_files(case, os.path.join(tmp, case.name))
    texts = {}
    for p in files['pmparins'] + (files['preliminaries'] or []) + [q for q in files['parfiles'] if q] + [files['inits']]:
        texts[os.path.basename(p)] = open(p).read()

    # ---- reference readers
    LoD_VLBI = quiet(ref_simulate.create_list_of_dict_VLBI, files['pmparins'], files['preliminaries'])
    LoD_timing = quiet(ref_simulate.create_list_of_dict_timing, files['parfiles'])
    limits, DoD = ref_priors.read_inits(files['inits'])
    shares = [list(r) for r in case.shares]
    names = list(ref_priors.get_parameters_from_shares(shares).keys())

    # Earth vectors at the epochs the reference will ask for (JD = MJD + 2400000.5)
    earth = [np.array([solsys.solarsystem(t + 2400000.5, 3, 0)[0] for t in d['epochs']]) for d in LoD_VLBI]

    like = quiet(ref_simulate.Gaussianlikelihood, case.refepoch, LoD_timing, LoD_VLBI, shares,
                 ref_positions.positions, DoD, case.a1dot_constraints)
    thetas = np.vstack([case.truth_vector()[None, :], case.sample_theta(n_theta - 1, seed=100)])
    # the truth vector was built for synthetic names; order is the same (same shares)
    assert names == case.names

    logl, models, errs_new, a1dots = [], [], [], []
    for th in thetas:
        params = {k: float(v) for k, v in zip(names, th)}
        like.parameters.update(params)
        logl.append(float(like.log_likelihood()))
        models.append([ref_positions.positions(case.refepoch, LoD_VLBI[i]['epochs'], LoD_timing[i], i,
                                               dict(params)).tolist() for i in range(len(LoD_VLBI))])
        errs_new.append([ref_simulate.adjust_errs_with_efac(LoD_VLBI[i], dict(params), i).tolist()
                         for i in range(len(LoD_VLBI))])
        if any(t != {} for t in LoD_timing):
            a1dots.append(ref_kopeikin.calculate_a1dot_pm(LoD_timing, dict(params)).tolist())

    reflex = []
    for i, t in enumerate(LoD_timing):
        if t != {}:
            p = dict(zip(names, thetas[1]))
            FP = ref_positions.filter_dictionary_of_parameter_with_index(p, i)
            Ps = sorted(FP.keys())
            for ep in LoD_VLBI[i]['epochs'][:4]:
                out = ref_reflex.reflex_motion(float(ep), t, float(FP[Ps[3]]), float(FP[Ps[6]]), float(FP[Ps[7]]))
                reflex.append({'source': i, 'epoch': float(ep), 'incl': float(FP[Ps[3]]), 'om_asc': float(FP[Ps[6]]),
                               'px': float(FP[Ps[7]]), 'out': [float(out[0]), float(out[1])]})

    return {
        'name': case.name, 'refepoch': case.refepoch, 'shares': shares, 'files': texts,
        'file_roles': {'pmparins': [os.path.basename(p) for p in files['pmparins']],
                       'preliminaries': None if files['preliminaries'] is None else
                       [os.path.basename(p) for p in files['preliminaries']],
                       'parfiles': [os.path.basename(p) if p else '' for p in files['parfiles']],
                       'inits': os.path.basename(files['inits'])},
        'a1dot_constraints': case.a1dot_constraints,
        'earth_xyz': [e.tolist() for e in earth],
        'theta': thetas.tolist(),
        'ref': {
            'names': names,
            'VLBI': [{k: np.asarray(v).tolist() for k, v in d.items()} for d in LoD_VLBI],
            'timing': [plain_timing(t) for t in LoD_timing],
            'limits': limits, 'DoD': DoD,
            'log_likelihood': logl, 'positions': models, 'errs_new': errs_new, 'a1dot_pm': a1dots,
            'reflex_motion': reflex,
        },
    }


def posterior_case(tmp):
    """A synthetic posterior for config3 'dots' -> the reference's bayesian_estimates.dat text."""
    from sterne import others as ref_others
    case = synthetic.config3('dots')
    files = synthetic.write_case_files(case, os.path.join(tmp, 'post'))
    LoD_VLBI = quiet(ref_simulate.create_list_of_dict_VLBI, files['pmparins'], files['preliminaries'])
    LoD_timing = quiet(ref_simulate.create_list_of_dict_timing, files['parfiles'])
    rng = np.random.default_rng(42)
    n = 1501
    names = case.names
    sig = {'ra': 3e-10, 'dec': 4e-10, 'mu_a': 0.05, 'mu_d': 0.06, 'px': 0.04, 'incl': 0.15, 'om_asc': 25.0}
    cols = []
    for k in names:
        root = synthetic.parameter_name_to_pmparin_indice(k)[1]
        mu = 350.0 if root == 'om_asc' else case.truth[k]          # straddles the 0/360 wrap
        x = mu + sig[root] * rng.standard_normal(n)
        cols.append(x % 360.0 if root == 'om_asc' else x)
    cols[names.index('mu_d_0')] += 0.5 * (cols[names.index('mu_a_0')] - case.truth['mu_a_0'])   # a correlation
    table = np.stack(cols + [rng.standard_normal(n), rng.standard_normal(n)], axis=1)
    samplefile = os.path.join(tmp, 'post', 'posterior_samples.dat')
    with open(samplefile, 'w') as f:
        f.write(' '.join(names + ['log_likelihood', 'log_prior']) + '\n')
        for row in table:
            f.write(' '.join(repr(float(v)) for v in row) + '\n')
    quiet(ref_simulate.make_a_summary_of_bayesian_inference, samplefile, case.refepoch, LoD_VLBI, LoD_timing)
    text = open(samplefile.replace('posterior_samples', 'bayesian_estimates')).read()
    est = []
    for seed, m in ((1, 400), (2, 401), (3, 1000)):
        x = np.random.default_rng(seed).standard_normal(m) * 3 + 1
        ang = (np.random.default_rng(seed).standard_normal(m) * 30 + 355) % 360
        est.append({'x': x.tolist(), 'ang': ang.tolist(), 'median': float(ref_others.sample2median(x)),
                    'range1': [float(v) for v in ref_others.sample2median_range(x, 1)],
                    'range2': [float(v) for v in ref_others.sample2median_range(x, 2)],
                    'periodic': [float(v) for v in ref_others.periodic_sample2estimate(ang)]})
    dms = [{'deg': d, 'dms': ref_others.deg2dms(float(d))} for d in (12.3456789, -0.5, 359.9999999, 0.0, -45.25)]
    return {'case': 'config3_dots', 'posterior_samples': open(samplefile).read(), 'bayesian_estimates': text,
            'estimators': est, 'deg2dms': dms,
            'earth_xyz': [np.array([solsys.solarsystem(t + 2400000.5, 3, 0)[0] for t in d['epochs']]).tolist()
                          for d in LoD_VLBI]}


def main():
    out = {'generator': 'tests/golden/make_golden.py', 'reference': 'dingswin/sterne 1.0.2 (/root/reference)',
           'note': 'outputs of the unmodified reference run under oracle/ref_shims stand-ins', 'cases': [],
           'solve_u': [], 'shares_names': [], 'dms2deg': [], 'decyear2mjd': []}
    with tempfile.TemporaryDirectory() as tmp:
        for case in (synthetic.config1(), synthetic.config2(), efad_case(), synthetic.config3('plain'),
                     synthetic.config3('wide'), synthetic.config3('dots'), synthetic.config4(), sentinel_case(),
                     synthetic.config3('efac'), synthetic.config3('efac_wide')):
            out['cases'].append(run_case(case, tmp))
            print('case %-16s logL(truth) = %.6f' % (case.name, out['cases'][-1]['ref']['log_likelihood'][0]))
    for e, c in [(0.0, 1.0), (1e-6, 0.3), (0.0794, 12.75), (0.3, -7.1), (0.6, 2.0), (0.9, 1e-3), (0.9, 3.0), (0.3, 500.123)]:
        u, it = ref_reflex.solve_u(e, c)
        out['solve_u'].append({'e': e, 'c': c, 'u': float(u), 'iterations': int(it)})
    for shares in ([[0, 0, 1, 1], [0, 1, 2, 2], [0, 0, 0, 0], [0, 0, 1, 1], [0, 0, 1, 1], [0, 0, 1, 1], [0, 0, 1, 1],
                    [0, 0, 0, 0], [0, 1, 2, 3]],
                   [[1, 2, 3, 4], [-1, -1, 0, 0], [-1, -1, -1, -1], [1, 1, 2, 2], [1, 1, 2, 2], [1, 1, 2, 2],
                    [-1, -1, 0, 0], [1, 1, 1, 1], [1, 2, 3, 4]],
                   [[0], [-1], [-1], [-1], [0], [0], [-1], [0], [0]],
                   [list(range(12)), [0] * 12, [-1] * 12, [-1] * 12, [0] * 12, [0] * 12, [-1] * 12, [0] * 6 + [1] * 6,
                    list(range(12))]):
        out['shares_names'].append({'shares': shares, 'names': list(ref_priors.get_parameters_from_shares(shares))})
    from sterne import others as ref_others
    for s in ['12:34:56.789', '-00:00:01.5', '+21:34:59', '-45:10:00.25', '00:00:00', '23 59 59.999', '-0:30:00']:
        out['dms2deg'].append({'s': s, 'deg': float(ref_others.dms2deg(s))})
    for y in [57123.456, 10000.5, 2015.0, 2016.5, 2020.123456, 1999.999]:
        out['decyear2mjd'].append({'epoch': y, 'mjd': float(ref_simulate.decyear2mjd(y)),
                                   'pinned': bool(y > 10000)})   # decimal-year branch is the Time stand-in
    with tempfile.TemporaryDirectory() as tmp:
        post = posterior_case(tmp)
    with open(os.path.join(HERE, 'golden_posterior_summary.json'), 'w') as f:
        json.dump(post, f)
    print('wrote golden_posterior_summary.json')
    # read_parfile's substring keyword matching (reflex_motion.py:59-75): ECCDOT overwrites ECC,
    # a KOM line matches 'OM ', EPS1/EPS2 are ignored, later lines win
    quirky = ('PSRJ            J1234+5678\n'
              'DECJ            -12:34:56.789               1  0.002\n'
              'PB              12.3456789               1  0.0000001\n'
              'ECC             0.1234               1  0.0001\n'
              'ECCDOT          1.5e-15               1  1e-16\n'
              'A1              9.8765               1  0.0001\n'
              'A1DOT           2.5e-14               1  1e-15\n'
              'T0              55555.5               1  0.01\n'
              'OM              123.4               1  0.1\n'
              'KOM             45.0               1  1.0\n'
              'OMDOT           0.0123               1  0.001\n'
              'PBDOT           1.2e-12               1  1e-13\n'
              'OM_ASC          77.7               1  2.0\n'
              'SINI            0.95               1  0.01\n')
    with tempfile.TemporaryDirectory() as tmp:
        qp = os.path.join(tmp, 'quirky.par')
        open(qp, 'w').write(quirky)
        out['parfile_quirks'] = {'text': quirky, 'ref': plain_timing(ref_reflex.read_parfile(qp))}
    path = os.path.join(HERE, 'golden_reference.json')
    with open(path, 'w') as f:
        json.dump(out, f)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
