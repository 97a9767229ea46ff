"""Synthetic workloads: the five BASELINE.json configurations.

There is no public data set for this path (the reference ships no example files), so
tests and bench.py use seeded synthetic astrometry: truth parameters -> model
positions -> Gaussian noise.  Cases can be kept in memory (dicts, as
create_list_of_dict_VLBI returns them) or written out as .pmpar.in / .preliminary /
.par / .inits files to exercise the readers.
"""
import os
import numpy as np

from .constants import D2YR, MAS2RAD, DEG2RAD, C_M_S
from .ephemeris import AnalyticEarthProvider, earth_vectors
from .orbit import orbit_terms
from .shares import get_parameters_from_shares, parameter_name_to_pmparin_indice


class SyntheticCase(object):
    def __init__(self, name, refepoch, LoD_timing, LoD_VLBI, shares, earth_xyz, truth, box,
                 DoD_additional_constraints=None, a1dot_constraints=False, parfile_text=None):
        self.name = name
        self.refepoch = refepoch
        self.LoD_timing = LoD_timing
        self.LoD_VLBI = LoD_VLBI
        self.shares = shares
        self.earth_xyz = earth_xyz
        self.names = list(get_parameters_from_shares(shares))
        self.truth = truth                      # {name: value}
        self.box = box                          # {name: (lo, hi)}
        self.DoD_additional_constraints = DoD_additional_constraints or {
            'sin_incl_Gaussian_constraints': {}, 'sin_incl_limits_constraints': {}}
        self.a1dot_constraints = a1dot_constraints
        self.parfile_text = parfile_text or {}

    @property
    def n_params(self):
        return len(self.names)

    @property
    def evals_per_sample(self):
        return int(sum(len(v['epochs']) for v in self.LoD_VLBI))

    def truth_vector(self):
        return np.array([self.truth[k] for k in self.names], dtype=np.float64)

    def sample_theta(self, N, seed, scale=1.0):
        """(N, P) uniform in the prior box (shrunk around its centre by `scale`)."""
        rng = np.random.default_rng(seed)
        lo = np.array([self.box[k][0] for k in self.names])
        hi = np.array([self.box[k][1] for k in self.names])
        mid, half = 0.5 * (lo + hi), 0.5 * (hi - lo) * scale
        return mid + (2.0 * rng.random((N, len(self.names))) - 1.0) * half

    def likelihood(self, **kwargs):
        from .likelihood import Gaussianlikelihood
        return Gaussianlikelihood(self.refepoch, self.LoD_timing, self.LoD_VLBI, self.shares, None,
                                  self.DoD_additional_constraints, self.a1dot_constraints,
                                  earth_xyz=self.earth_xyz, **kwargs)


# --------------------------------------------------------------------------------
# truth model used ONLY to synthesise observations (not the product path)
# --------------------------------------------------------------------------------
def _truth_positions(refepoch, epochs, xyz, ra, dec, mu_a, mu_d, px, timing=None, incl=None, om_asc=None):
    dT = (epochs - refepoch) * D2YR
    o_ra = dT * mu_a / np.cos(dec) + px * (xyz[:, 0] * np.sin(ra) - xyz[:, 1] * np.cos(ra)) / np.cos(dec)
    o_dec = dT * mu_d + px * (xyz[:, 0] * np.cos(ra) * np.sin(dec) + xyz[:, 1] * np.sin(ra) * np.sin(dec)
                             - xyz[:, 2] * np.cos(dec))
    if timing:
        t = orbit_terms(epochs, timing)
        off = t[:, 0] * px
        Om = om_asc * DEG2RAD
        o_ra += off * (np.sin(Om) * t[:, 1] + np.cos(Om) * np.cos(incl) * t[:, 2]) / np.cos(timing['decj'] * DEG2RAD)
        o_dec += off * (np.cos(Om) * t[:, 1] - np.sin(Om) * np.cos(incl) * t[:, 2])
    return ra + o_ra * MAS2RAD, dec + o_dec * MAS2RAD


def make_source(rng, refepoch, epochs, ra, dec, mu_a, mu_d, px, sigma_mas=(0.05, 0.3), timing=None,
                incl=None, om_asc=None, random_fraction=None, provider=None):
    """One source's VLBI dict (+ Earth vectors).  random_fraction: sigma_random / sigma
    (gives 'errs_random' / 'errs_sys' as a .preliminary file would)."""
    provider = provider or AnalyticEarthProvider()
    epochs = np.sort(np.asarray(epochs, dtype=np.float64))
    E = len(epochs)
    xyz = earth_vectors(provider, epochs)
    m_ra, m_dec = _truth_positions(refepoch, epochs, xyz, ra, dec, mu_a, mu_d, px, timing, incl, om_asc)
    s_ra = rng.uniform(sigma_mas[0], sigma_mas[1], E) * MAS2RAD
    s_dec = rng.uniform(sigma_mas[0], sigma_mas[1], E) * MAS2RAD
    # errRA in pmpar.in is seconds of time without cos(dec): sigma_ra here is already the
    # uncertainty of the RA coordinate itself (simulate.py:474-475)
    obs_ra = m_ra + rng.standard_normal(E) * s_ra
    obs_dec = m_dec + rng.standard_normal(E) * s_dec
    vlbi = {'epochs': epochs, 'radecs': np.concatenate([obs_ra, obs_dec]),
            'errs': np.concatenate([s_ra, s_dec])}
    if random_fraction is not None:
        vlbi['errs_random'] = vlbi['errs'] * random_fraction
        vlbi['errs_sys'] = (vlbi['errs'] ** 2 - vlbi['errs_random'] ** 2) ** 0.5
    return vlbi, xyz


def _box_for(names, truth):
    box = {}
    for k in names:
        root = parameter_name_to_pmparin_indice(k)[1]
        v = truth[k]
        if root in ('ra', 'dec'):
            box[k] = (v - 3.0 * MAS2RAD, v + 3.0 * MAS2RAD)
        elif root in ('mu_a', 'mu_d'):
            box[k] = (v - 2.0, v + 2.0)
        elif root == 'px':
            box[k] = (max(v - 1.0, 0.05), v + 1.0)
        elif root == 'efac':
            box[k] = (0.0, 15.0)
        elif root == 'efad':
            box[k] = (0.0, 20.0)
        elif root == 'incl':
            box[k] = (0.05, 3.09)
        elif root == 'om_asc':
            box[k] = (0.0, 360.0)
    return box


TRUTH = dict(ra=5.1, dec=0.37, mu_a=-3.2, mu_d=7.9, px=1.25, incl=1.2, om_asc=135.0, efac=1.3, efad=1.1)

PARFILE_CFG3 = dict(pb=1.5334, ecc=0.0794, a1=3.7294, t0=53155.9, om=1.35, decj='+21:34:59')


def timing_dict(pb, ecc, a1, t0, om, decj, omdot=None, pbdot=None, a1dot=None):
    """Plain-float orbital dict as readers.read_parfile returns it."""
    from .readers import dms_str2deg
    d = {'pb': pb, 'ecc': ecc, 'a1': a1 * C_M_S, 't0': t0, 'om': om,
         'decj': dms_str2deg(decj) if isinstance(decj, str) else decj}
    if omdot is not None:
        d['omdot'] = omdot
    if pbdot is not None:
        d['pbdot'] = pbdot
    if a1dot is not None:
        d['a1dot'] = a1dot * C_M_S
    return d


def parfile_text(pb, ecc, a1, t0, om, decj, omdot=None, pbdot=None, a1dot=None, name='J0000+0000'):
    """PSRCAT-style text: keyword and value at least 8 spaces apart (reflex_motion.py:65)."""
    lines = ['PSRJ            %s' % name,
             'RAJ             19:28:27.2               1  0.003',
             'DECJ            %s               1  0.004' % decj,
             'PB              %r               1  0.0000000002' % pb,
             'ECC             %r               1  0.000002' % ecc,
             'A1              %r               1  0.000005' % a1,
             'T0              %r               1  0.00003' % t0,
             'OM              %r               1  0.007' % om]
    if omdot is not None:
        lines.append('OMDOT           %r               1  0.001' % omdot)
    if pbdot is not None:
        lines.append('PBDOT           %r               1  1e-14' % pbdot)
    if a1dot is not None:
        lines.append('A1DOT           %r               1  1e-16' % a1dot)
    return '\n'.join(lines) + '\n'


def _single(name, seed, with_efac=False, binary=None, n_epochs=10, gaussian_sini=None, a1dot=None):
    rng = np.random.default_rng(seed)
    refepoch = 57365.0
    epochs = 57000.0 + rng.uniform(0.0, 730.0, n_epochs)
    timing = timing_dict(**binary) if binary else {}
    vlbi, xyz = make_source(rng, refepoch, epochs, TRUTH['ra'], TRUTH['dec'], TRUTH['mu_a'], TRUTH['mu_d'],
                            TRUTH['px'], timing=timing or None, incl=TRUTH['incl'], om_asc=TRUTH['om_asc'],
                            random_fraction=0.6 if with_efac else None)
    shares = [[0], [0] if with_efac else [-1], [-1], [0] if binary else [-1], [0], [0],
              [0] if binary else [-1], [0], [0]]
    names = list(get_parameters_from_shares(shares))
    truth = {k: TRUTH[parameter_name_to_pmparin_indice(k)[1]] for k in names}
    DoD = {'sin_incl_Gaussian_constraints': {}, 'sin_incl_limits_constraints': {}}
    if gaussian_sini is not None:
        DoD['sin_incl_Gaussian_constraints']['incl_0'] = list(gaussian_sini)
    return SyntheticCase(name, refepoch, [timing], [vlbi], shares, [xyz], truth, _box_for(names, truth), DoD,
                         a1dot if a1dot is not None else False,
                         {0: parfile_text(**binary)} if binary else {})


def config1():
    """single pulsar, ~10 epochs, 5 parameters (ra, dec, px, mu_a, mu_d)."""
    return _single('config1', 1)


def config2():
    """config 1 + .preliminary errors -> EFAC, 6 parameters; batch of 4096 walkers."""
    return _single('config2', 1, with_efac=True)


def config3(variant='efac'):
    """binary pulsar, 8 parameters incl. Omega_asc and inclination (Tempo2 reflex).

    'efac' is BASELINE config 3 as written: config 2's six parameters (EFAC on) plus incl_0 and
    om_asc_0, P = 8, parfile of SURVEY 8(d).  'plain' / 'wide' / 'dots' are the same source
    without EFAC (P = 7): the survey's orbit, a wide orbit, and OMDOT/PBDOT/A1DOT + constraints."""
    b = dict(PARFILE_CFG3)
    kw = {}
    if variant == 'efac':
        kw = dict(with_efac=True)
    elif variant == 'efac_wide':   # EFAC + a reflex term that is a visible fraction of a mas
        b.update(pb=175.46, ecc=0.3, a1=55.33, t0=56100.25, om=50.7)
        kw = dict(with_efac=True)
    elif variant == 'wide':          # a wide orbit so the reflex term is a visible fraction of a mas
        b.update(pb=175.46, ecc=0.3, a1=55.33, t0=56100.25, om=50.7)
    elif variant == 'dots':
        b.update(pb=175.46, ecc=0.3, a1=55.33, t0=56100.25, om=50.7, omdot=0.0123, pbdot=1.2e-11, a1dot=3.4e-14)
        kw = dict(gaussian_sini=(0.93, 0.02), a1dot=[[1.1e-14, 4e-15]])
    return _single('config3_' + variant, 3, binary=b, **kw)


def config4(n_epochs=10):
    """joint fit of one pulsar measured against 4 in-beam calibrators + the pulsar's own
    series: 5 sources, shared px / mu_a / mu_d / efac / incl / om_asc, distinct ra / dec."""
    rng = np.random.default_rng(4)
    S = 5
    refepoch = 57365.0
    b = dict(PARFILE_CFG3)
    b.update(pb=175.46, ecc=0.3, a1=55.33, t0=56100.25, om=50.7)
    timing = timing_dict(**b)
    LoD_VLBI, xyzs, LoD_timing = [], [], []
    truth_pos = []
    for s in range(S):
        ra = TRUTH['ra'] + s * 2.0e-6
        dec = TRUTH['dec'] - s * 1.5e-6
        truth_pos.append((ra, dec))
        epochs = 57000.0 + rng.uniform(0.0, 730.0, n_epochs + s)      # ragged epoch counts
        vlbi, xyz = make_source(rng, refepoch, epochs, ra, dec, TRUTH['mu_a'], TRUTH['mu_d'], TRUTH['px'],
                                timing=timing, incl=TRUTH['incl'], om_asc=TRUTH['om_asc'], random_fraction=0.6)
        LoD_VLBI.append(vlbi)
        xyzs.append(xyz)
        LoD_timing.append(dict(timing))
    shares = [list(range(S)), [0] * S, [-1] * S, [0] * S, [0] * S, [0] * S, [0] * S, [0] * S, list(range(S))]
    names = list(get_parameters_from_shares(shares))
    truth = {}
    for k in names:
        srcs, root = parameter_name_to_pmparin_indice(k)
        if root == 'ra':
            truth[k] = truth_pos[srcs[0]][0]
        elif root == 'dec':
            truth[k] = truth_pos[srcs[0]][1]
        else:
            truth[k] = TRUTH[root]
    return SyntheticCase('config4', refepoch, LoD_timing, LoD_VLBI, shares, xyzs, truth, _box_for(names, truth),
                         parfile_text={s: parfile_text(**b) for s in range(S)})


def config5(variant='C', n_sources=8, n_epochs=1024, seed=5):
    """throughput sweep: 8 sources x 1024 epochs, shared px / mu_a / mu_d, distinct ra / dec
    (A, P=19); + one shared EFAC (B, P=20); + source 0 binary with incl / om_asc (C, P=22)."""
    if variant not in ('A', 'B', 'C'):
        raise ValueError(variant)
    rng = np.random.default_rng(seed)
    S = n_sources
    refepoch = 57000.0 + 1826.0
    b = dict(PARFILE_CFG3)
    b.update(pb=175.46, ecc=0.3, a1=55.33, t0=56100.25, om=50.7, omdot=0.0123)
    timing0 = timing_dict(**b) if variant == 'C' else {}
    LoD_VLBI, xyzs, LoD_timing, truth_pos = [], [], [], []
    for s in range(S):
        ra = TRUTH['ra'] + s * 2.0e-6
        dec = TRUTH['dec'] - s * 1.5e-6
        truth_pos.append((ra, dec))
        epochs = 57000.0 + rng.uniform(0.0, 3652.0, n_epochs)
        tm = timing0 if (s == 0 and variant == 'C') else {}
        vlbi, xyz = make_source(rng, refepoch, epochs, ra, dec, TRUTH['mu_a'], TRUTH['mu_d'], TRUTH['px'],
                                timing=tm or None, incl=TRUTH['incl'], om_asc=TRUTH['om_asc'],
                                random_fraction=0.6 if variant in ('B', 'C') else None)
        LoD_VLBI.append(vlbi)
        xyzs.append(xyz)
        LoD_timing.append(tm)
    efac_row = [0] * S if variant in ('B', 'C') else [-1] * S
    bin_row = [0] + [-1] * (S - 1) if variant == 'C' else [-1] * S
    shares = [list(range(S)), efac_row, [-1] * S, list(bin_row), [0] * S, [0] * S, list(bin_row), [0] * S,
              list(range(S))]
    names = list(get_parameters_from_shares(shares))
    truth = {}
    for k in names:
        srcs, root = parameter_name_to_pmparin_indice(k)
        if root == 'ra':
            truth[k] = truth_pos[srcs[0]][0]
        elif root == 'dec':
            truth[k] = truth_pos[srcs[0]][1]
        else:
            truth[k] = TRUTH[root]
    box = _box_for(names, truth)
    for k in names:                                   # a tighter EFAC box keeps chi^2 in a sane range
        if parameter_name_to_pmparin_indice(k)[1] == 'efac':
            box[k] = (0.5, 3.0)
    return SyntheticCase('config5' + variant, refepoch, LoD_timing, LoD_VLBI, shares, xyzs, truth, box)


def device_theta(case, N, seed, device, sigma_frac=0.25):
    """(N, P) parameter vectors generated ON the device: truth + N(0, sigma), sigma a
    fraction of the prior half-width (an (N, P) host array at
    N = 2^22 would be 738 MB of H2D traffic)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    truth = torch.tensor(case.truth_vector(), dtype=torch.float64, device=device)
    half = torch.tensor([0.5 * (case.box[k][1] - case.box[k][0]) for k in case.names], dtype=torch.float64,
                        device=device)
    lo = torch.tensor([case.box[k][0] for k in case.names], dtype=torch.float64, device=device)
    hi = torch.tensor([case.box[k][1] for k in case.names], dtype=torch.float64, device=device)
    z = torch.randn((N, len(case.names)), dtype=torch.float64, device=device, generator=g)
    return torch.minimum(torch.maximum(truth + z * (half * sigma_frac), lo), hi)


# --------------------------------------------------------------------------------
# file writers (formats of SURVEY 3.4)
# --------------------------------------------------------------------------------
def _sexagesimal(value, n_dec):
    sign = '-' if value < 0 else ''
    value = abs(value)
    d = int(value)
    m = int((value - d) * 60)
    s = ((value - d) * 60 - m) * 60
    s = round(s, n_dec)
    if s >= 60:
        s -= 60
        m += 1
    if m >= 60:
        m -= 60
        d += 1
    return '%s%02d:%02d:%0*.*f' % (sign, d, m, n_dec + 3, n_dec, s)


def write_pmparin(path, epochs, radecs, errs):
    E = len(epochs)
    with open(path, 'w') as f:
        f.write('### pmpar format\n#name = synthetic\n#epoch = MJD\n')
        for i in range(E):
            ra_h = radecs[i] / (15 * np.pi / 180.)
            dec_d = radecs[E + i] / (np.pi / 180.)
            era_s = errs[i] / (15 * np.pi / 180. / 3600.)
            edec_as = errs[E + i] / (np.pi / 180. / 3600.)
            f.write('%.6f %s %.12g %s %.12g\n' % (epochs[i], _sexagesimal(ra_h, 10), era_s,
                                                  _sexagesimal(dec_d, 9), edec_as))


def write_case_files(case, directory):
    """Writes <i>.pmpar.in [.preliminary], <i>.par and case.inits; returns their paths."""
    os.makedirs(directory, exist_ok=True)
    pmparins, prelims, parfiles = [], [], []
    for i, vlbi in enumerate(case.LoD_VLBI):
        p = os.path.join(directory, 'src%d.pmpar.in' % i)
        write_pmparin(p, vlbi['epochs'], vlbi['radecs'], vlbi['errs'])
        pmparins.append(p)
        if 'errs_random' in vlbi:
            q = p + '.preliminary'
            write_pmparin(q, vlbi['epochs'], vlbi['radecs'], vlbi['errs_random'])
            prelims.append(q)
        if i in case.parfile_text:
            r = os.path.join(directory, 'src%d.par' % i)
            with open(r, 'w') as f:
                f.write(case.parfile_text[i])
            parfiles.append(r)
        else:
            parfiles.append('')
    inits = os.path.join(directory, case.name + '.inits')
    with open(inits, 'w') as f:
        f.write('#Prior info at MJD %f.\n' % case.refepoch)
        for k in case.names:
            lo, hi = case.box[k]
            kind = 'Sine' if 'incl' in k else 'Uniform'
            extra = ''
            g = case.DoD_additional_constraints['sin_incl_Gaussian_constraints'].get(k)
            if g is not None:
                extra = ',%r,%r,Sine_Gaussian' % (g[0], g[1])
            f.write('%s: %.17g,%.17g,%s%s\n' % (k, lo, hi, kind, extra))
    return dict(pmparins=pmparins, preliminaries=prelims or None, parfiles=parfiles, inits=inits)
