"""
CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The shipped package (sterne_b200/) never does.

This is a plain-float / numpy restatement of the astrometric log-likelihood hot
path of dingswin/sterne.  Each function cites the reference file:line it follows
(paths relative to /root/reference/src/sterne/).  The restatement keeps the
reference's OPERATION ORDER, because `obs - (ra + off*K)` cancels catastrophically
and the reference's rounding is the specification.

Parity pinning
--------------
The reference has no tests, fixtures or golden vectors of its own, and its
third-party dependencies (astropy, novas + DE421, bilby, psrqpy) are absent from
this image.  The oracle is pinned in two ways:
  1. tests/golden/make_golden.py imports the UNMODIFIED reference modules from
     /root/reference/src with small stand-in modules for the absent third-party
     packages (oracle/ref_shims/) and records the reference's own outputs of
     positions(), adjust_errs_with_efac(), Gaussianlikelihood.log_likelihood(),
     reflex_motion(), solve_u(), calculate_a1dot_pm(), get_parameters_from_shares(),
     readpmparin(), read_parfile(), read_inits() as fixtures; tests compare this
     oracle with them.
  2. closed-form known-answer tests and an mpmath arbiter (tests/test_oracle_kat.py).
What stays UNPINNED (third-party arithmetic that cannot run here): the exact
float values of astropy's unit-conversion factors, of novas/DE421 Earth positions
(treated as an input table), and astropy's decimalyear->MJD conversion.

Units: ra/dec rad, mu mas/yr, px mas, incl rad, om_asc deg, epochs MJD.
"""
import math
import numpy as np

# ---------------------------------------------------------------------------
# unit factors (astropy.units definitions; see header: values unpinned to the ulp)
# ---------------------------------------------------------------------------
D2YR = 86400.0 / 31557600.0            # (u.d).to(u.yr)      positions.py:108
MAS2RAD = (math.pi / 180.0) / 3600.0 * 0.001   # (u.mas).to(u.rad)  positions.py:119-120
DEG2RAD = math.pi / 180.0              # deg -> rad (np.sin of a deg Quantity)
RAD2DEG = 180.0 / math.pi              # rad -> deg (deg + rad Quantity addition)
C_M_S = 299792458.0                    # astropy.constants.c
M2AU = 1.0 / 149597870700.0            # (m).to(AU)          reflex_motion.py:200
D2S = 86400.0                          # d -> s
YR2D = 365.25
MASYR2RADS = MAS2RAD / (YR2D * D2S)    # (mas/yr).to(rad/s)  kopeikin_effects.py:45


def use_astropy():
    """The reference computes these factors with astropy at run time; when astropy is importable
    the oracle takes astropy's own values (last-bit agreement), as sterne_b200/constants.py does."""
    global D2YR, MAS2RAD, DEG2RAD, RAD2DEG, C_M_S, M2AU, D2S, YR2D, MASYR2RADS
    try:
        import astropy.units as u
        import astropy.constants as c
        vals = dict(D2YR=float((u.d).to(u.yr)), MAS2RAD=float((u.mas).to(u.rad)), DEG2RAD=float((u.deg).to(u.rad)),
                    RAD2DEG=float((u.rad).to(u.deg)), C_M_S=float(c.c.to(u.m / u.s).value), M2AU=float((u.m).to(u.AU)),
                    D2S=float((u.d).to(u.s)), YR2D=float((u.yr).to(u.d)),
                    MASYR2RADS=float((u.mas / u.yr).to(u.rad / u.s)))
    except Exception:
        return False
    g = globals()
    if any(not (v > 0 and abs(v / g[k] - 1.0) < 1e-12) for k, v in vals.items()):
        return False
    g.update(vals)
    return True


use_astropy()

ROOTS = ['dec', 'efac', 'efad', 'incl', 'mu_a', 'mu_d', 'om_asc', 'px', 'ra']


# ---------------------------------------------------------------------------
# priors.py:268-302  parameter names from `shares`
# ---------------------------------------------------------------------------
def group_elements_by_same_values(alist):
    """priors.py:279-302"""
    out = []
    alist = np.array(alist)
    grouped = []
    for i in range(len(alist)):
        if (i not in grouped) and (alist[i] >= 0):
            each_group = np.where(alist == alist[i])[0].tolist()
            grouped += each_group
            each_group.sort()
            out.append('_' + '_'.join(str(e) for e in each_group))
    return out


def get_parameters_from_shares(shares):
    """priors.py:268-277"""
    parameters = {}
    for i in range(len(ROOTS)):
        for string in group_elements_by_same_values(shares[i]):
            parameters[ROOTS[i] + string] = None
    return parameters


# ---------------------------------------------------------------------------
# positions.py:32-60  parameter -> source routing
# ---------------------------------------------------------------------------
def auto_fill_disabled_parameters(filtered):
    """positions.py:52-60 (substring test on the root, as in the reference)"""
    for root in ROOTS:
        if not any(root in parameter for parameter in filtered.keys()):
            filtered[root] = -999
    return filtered


def filter_dictionary_of_parameter_with_index(dict_of_parameters, filter_index):
    """positions.py:32-51"""
    filtered = dict_of_parameters.copy()
    for parameter in dict_of_parameters.keys():
        if not str(filter_index) in parameter.split('_'):
            del filtered[parameter]
    return auto_fill_disabled_parameters(filtered)


# ---------------------------------------------------------------------------
# reflex_motion.py:106-126  Kepler fixed point (tolerance 1e-5 is a reference quirk)
# ---------------------------------------------------------------------------
def solve_u(e, c, precision=1e-5):
    x = c / (1 - e)
    x1 = float('inf')
    iterations = 0
    while abs(x - x1) > precision:
        x1 = x
        x = e * np.sin(x1) + c
        iterations += 1
    return x, iterations


def orbital_dict(pb, ecc, a1, t0, om, decj_deg, omdot=None, pbdot=None, a1dot=None,
                 om_asc=None, sini=None):
    """Plain-float image of reflex_motion.read_parfile's dict (reflex_motion.py:76-93).

    pb [d], a1 stored in METRES (A1[lt-s]*c, :77), t0 [d], om [deg], omdot [deg/yr],
    a1dot stored in m/s (A1DOT*c, :89), decj [deg].  Optional keys absent when None.
    """
    d = {'pb': pb, 'ecc': ecc, 'a1': a1 * C_M_S, 't0': t0, 'om': om, 'decj': decj_deg}
    if omdot is not None:
        d['omdot'] = omdot
    if pbdot is not None:
        d['pbdot'] = pbdot
    if a1dot is not None:
        d['a1dot'] = a1dot * C_M_S
    if om_asc is not None:
        d['om_asc'] = om_asc
    if sini is not None:
        d['sini'] = sini
    return d


def reflex_orbit_terms(epoch, DoP):
    """Parameter-independent part of reflex_motion (reflex_motion.py:168-200).

    Returns (b_AU, theta_rad_argument_in_deg) -> (b_AU, cos(theta), sin(theta)).
    Unit conversions astropy performs implicitly are written out:
      :187  n [rad/d]
      :189  M = n*(t-T0), not reduced mod 2pi
      :191  ((1+e)/(1-e))**2  (reference quirk; not a square root)
      :192-193  k = omdot/n has unit (deg/yr)/(rad/d); k*A_u is deg*d/yr, converted
                to deg (x 1/365.25) when added to omega0 [deg]
      :194  theta = omega[deg] + A_u[rad -> deg]
      :196  a1[m] + adot[m/s]*(t-T0)[d -> s]
      :200  to AU
    """
    e, T0, Pb0, omega0, a0, = DoP['ecc'], DoP['t0'], DoP['pb'], DoP['om'], DoP['a1']
    omdot = DoP.get('omdot', 0.0)
    Pbdot = DoP.get('pbdot', 0)
    adot = DoP.get('a1dot', None)
    n = 2 * np.pi / Pb0 - np.pi * Pbdot * (epoch - T0) / (Pb0 ** 2)
    u1 = solve_u(e, n * (epoch - T0))[0]
    A_u = 2 * np.arctan(((1 + e) / (1 - e)) ** 2 * np.tan(u1 / 2))
    k = omdot / n
    omega = omega0 + (k * A_u) * D2YR              # deg*d/yr -> deg (default omdot = 0: no-op)
    theta_deg = omega + A_u * RAD2DEG
    if adot is not None:
        a1 = a0 + (adot * (epoch - T0)) * D2S
    else:
        a1 = a0
    b_abs = a1 * (1 - e * np.cos(u1))
    b_AU = b_abs * M2AU
    theta = theta_deg * DEG2RAD
    return b_AU, np.cos(theta), np.sin(theta)


def reflex_motion(epoch, DoP, incl, Om_asc, px):
    """reflex_motion.py:128-214.  Om_asc in deg, incl in rad, px in mas -> [dRA, dDEC] mas."""
    b_AU, cos_th, sin_th = reflex_orbit_terms(epoch, DoP)
    offset = b_AU * px
    Om = Om_asc * DEG2RAD
    sO, cO = np.sin(Om), np.cos(Om)
    si, ci = np.sin(incl), np.cos(incl)
    matr1 = np.asmatrix([[sO, -cO, 0],
                         [cO, sO, 0],
                         [0, 0, 1]])
    matr2 = np.asmatrix([[1, 0, 0],
                         [0, -ci, -si],
                         [0, si, -ci]])
    matr3 = np.asmatrix([[offset * cos_th],
                         [offset * sin_th],
                         [0]])
    b = matr1 * matr2 * matr3
    dRA = b.item(0, 0) / np.cos(DoP['decj'] * DEG2RAD)
    dDEC = b.item(1, 0)
    return np.array([dRA, dDEC])


# ---------------------------------------------------------------------------
# positions.py:266-310  parallax (Earth vector is an INPUT here; novas/DE421 absent)
# ---------------------------------------------------------------------------
def parallax_related_position_offset_from_the_barycentric_frame(XYZ, ra, dec, px):
    """positions.py:305-310 with (X, Y, Z) = solarsystem(epoch+2400000.5, 3, 0)[0] supplied."""
    X, Y, Z = XYZ
    dRA = px * (X * np.sin(ra) - Y * np.cos(ra)) / np.cos(dec)
    dDEC = px * (X * np.cos(ra) * np.sin(dec) + Y * np.sin(ra) * np.sin(dec) - Z * np.cos(dec))
    return np.array([dRA, dDEC])


def position_with_offset(refepoch, epoch, XYZ, dec_rad, incl, mu_a, mu_d, om_asc, px, ra_rad,
                         dict_of_timing_parameters={}):
    """positions.py:62-121; also returns the offset in mas (the 1e-12 mas check is on it)."""
    dT = (epoch - refepoch) * D2YR
    dRA = dT * mu_a / np.cos(dec_rad)
    dDEC = dT * mu_d
    offset = np.array([dRA, dDEC], dtype=np.float64)
    if px != -999:
        offset += parallax_related_position_offset_from_the_barycentric_frame(XYZ, ra_rad, dec_rad, px)
        if (incl != -999) and (om_asc != -999) and (dict_of_timing_parameters != {}):
            offset += reflex_motion(epoch, dict_of_timing_parameters, incl, om_asc, px)
    ra_rad = ra_rad + offset[0] * MAS2RAD
    dec_rad = dec_rad + offset[1] * MAS2RAD
    return ra_rad, dec_rad, offset[0], offset[1]


def position(refepoch, epoch, XYZ, dec_rad, incl, mu_a, mu_d, om_asc, px, ra_rad,
             dict_of_timing_parameters={}):
    return position_with_offset(refepoch, epoch, XYZ, dec_rad, incl, mu_a, mu_d, om_asc, px,
                                ra_rad, dict_of_timing_parameters)[:2]


def positions(refepoch, epochs, earth_xyz, dict_timing, parameter_filter_index, dict_parameters,
              return_offsets=False):
    """positions.py:13-30.  earth_xyz: (E, 3) AU.  Positional pick from the SORTED key list."""
    FP = filter_dictionary_of_parameter_with_index(dict_parameters, parameter_filter_index)
    Ps = sorted(FP.keys())
    ra_models, dec_models, ora, odec = [], [], [], []
    for i in range(len(epochs)):
        r, d, o0, o1 = position_with_offset(refepoch, epochs[i], earth_xyz[i], FP[Ps[0]], FP[Ps[3]],
                                            FP[Ps[4]], FP[Ps[5]], FP[Ps[6]], FP[Ps[7]], FP[Ps[8]],
                                            dict_timing)
        ra_models.append(r); dec_models.append(d); ora.append(o0); odec.append(o1)
    model = np.concatenate([np.array(ra_models, dtype=np.float64), np.array(dec_models, dtype=np.float64)])
    if return_offsets:
        return model, np.concatenate([np.array(ora, dtype=np.float64), np.array(odec, dtype=np.float64)])
    return model


# ---------------------------------------------------------------------------
# simulate.py:303-320  EFAC / EFAD
# ---------------------------------------------------------------------------
def adjust_errs_with_efac(VLBI_dict, parameters_dict, parameter_filter_index):
    FP = filter_dictionary_of_parameter_with_index(parameters_dict, parameter_filter_index)
    Ps = sorted(FP.keys())
    efac = FP[Ps[1]]
    efad = FP[Ps[2]]
    N = int(len(VLBI_dict['errs']) / 2)
    if efad != -999:
        efads = np.concatenate([np.ones(N), efad * np.ones(N)])
    else:
        efads = np.ones(2 * N)
    if efac != -999:
        errs_new_sq = (VLBI_dict['errs_random']) ** 2 + (efac * efads * VLBI_dict['errs_sys']) ** 2
    else:
        errs_new_sq = VLBI_dict['errs'] ** 2
    return errs_new_sq ** 0.5


# ---------------------------------------------------------------------------
# kopeikin_effects.py:11-68
# ---------------------------------------------------------------------------
def a1dot_pm_formalisim(a1, inc, mu_a, mu_d, om_asc):
    """kopeikin_effects.py:11-45; a1 lt-s, inc rad, mu mas/yr, om_asc deg -> lt-s/s."""
    mu = (mu_a ** 2 + mu_d ** 2) ** 0.5
    theta_mu = np.arctan2(mu_a, mu_d)
    a1dot_pm = mu * np.sin(theta_mu - om_asc * DEG2RAD) / np.tan(inc)
    a1dot_pm = a1dot_pm * a1
    return 1e0 * (a1dot_pm * MASYR2RADS)


def calculate_a1dot_pm(list_of_dict_timing, parameters_dict):
    """kopeikin_effects.py:47-68"""
    out = []
    for i in range(len(list_of_dict_timing)):
        if list_of_dict_timing[i] != {}:
            FP = filter_dictionary_of_parameter_with_index(parameters_dict, i)
            Ps = sorted(FP.keys())
            a1 = list_of_dict_timing[i]['a1'] / C_M_S       # (a1[m] / c / s).value
            out.append(a1dot_pm_formalisim(a1, FP[Ps[3]], FP[Ps[4]], FP[Ps[5]], FP[Ps[6]]))
    if len(out) == 0:
        raise SystemExit("The list_of_modeled_a1dot is empty. Make sure not all parfiles are ''.")
    return np.array(out, dtype=np.float64)


def parse_a1dot_constraints(a1dot_constraints):
    """simulate.py:387-399"""
    mus, sigmas = [], []
    for c in a1dot_constraints:
        if len(c) == 2:
            mus.append(c[0]); sigmas.append(c[1])
    return len(mus) != 0, np.array(mus, dtype=np.float64), np.array(sigmas, dtype=np.float64)


# ---------------------------------------------------------------------------
# simulate.py:325-384  the likelihood
# ---------------------------------------------------------------------------
class Gaussianlikelihood(object):
    """simulate.py:325-384 with the Earth vectors supplied per source (list of (E,3) arrays)."""

    def __init__(self, refepoch, list_of_dict_timing, list_of_dict_VLBI, shares, earth_xyz,
                 DoD_additional_constraints=None, a1dot_constraints=False):
        self.refepoch = refepoch
        self.LoD_VLBI = list_of_dict_VLBI
        self.LoD_timing = list_of_dict_timing
        self.earth_xyz = earth_xyz
        self.shares = shares
        self.number_of_pmparins = len(self.LoD_VLBI)
        if DoD_additional_constraints is None:
            DoD_additional_constraints = {'sin_incl_Gaussian_constraints': {},
                                          'sin_incl_limits_constraints': {}}
        self.dict_sin_incl_Gaussian_constraints = DoD_additional_constraints['sin_incl_Gaussian_constraints']
        if a1dot_constraints != False:
            self.a1dot_constraints, self.a1dot_mus, self.a1dot_sigmas = parse_a1dot_constraints(a1dot_constraints)
        else:
            self.a1dot_constraints = False
        self.parameters = get_parameters_from_shares(self.shares)

    def log_likelihood(self):
        """simulate.py:362-384"""
        log_p = 0
        for i in range(self.number_of_pmparins):
            res = self.LoD_VLBI[i]['radecs'] - positions(self.refepoch, self.LoD_VLBI[i]['epochs'],
                                                          self.earth_xyz[i], self.LoD_timing[i], i,
                                                          self.parameters)
            errs_new = adjust_errs_with_efac(self.LoD_VLBI[i], self.parameters, i)
            log_p += -0.5 * np.sum((res / errs_new) ** 2)
            log_p += -1 * np.sum(np.log(errs_new))
        if self.a1dot_constraints:
            modeled_a1dots = calculate_a1dot_pm(self.LoD_timing, self.parameters)
            res_a1dots = modeled_a1dots - self.a1dot_mus
            log_p += -0.5 * np.sum((res_a1dots / self.a1dot_sigmas) ** 2)
        if len(self.dict_sin_incl_Gaussian_constraints) != 0:
            for parameter in self.dict_sin_incl_Gaussian_constraints:
                sin_incl_mu, sin_incl_sigma = self.dict_sin_incl_Gaussian_constraints[parameter]
                res = sin_incl_mu - np.sin(self.parameters[parameter])
                log_p += -0.5 * (res / sin_incl_sigma) ** 2
        return log_p

    # -- convenience for tests / CPU baseline: loop the reference-shaped path over rows
    def log_likelihood_rows(self, theta, keys=None):
        keys = list(self.parameters) if keys is None else list(keys)
        out = np.empty(len(theta), dtype=np.float64)
        for n, row in enumerate(theta):
            for k, v in zip(keys, row):
                self.parameters[k] = float(v)
            out[n] = self.log_likelihood()
        return out


# ---------------------------------------------------------------------------
# Vectorised (over N parameter vectors) restatement: same operations in the same
# order, elementwise over N -- used for parity at N in the thousands and as the
# numpy CPU baseline.  Reductions over the 2E terms use np.sum along the last axis
# of an (N, 2E) array, as the reference's np.sum over a (2E,) vector does per row.
# ---------------------------------------------------------------------------
def column_table(shares):
    """idx[s][r] = column of root r for source s in get_parameters_from_shares order, or -1."""
    names = list(get_parameters_from_shares(shares))
    S = len(shares[0])
    idx = -np.ones((S, len(ROOTS)), dtype=np.int64)
    for c, name in enumerate(names):
        toks = name.split('_')
        srcs = [int(t) for t in toks if t.isdigit()]
        root = '_'.join(t for t in toks if not t.isdigit())
        for s in srcs:
            idx[s, ROOTS.index(root)] = c
    return names, idx


def _col(theta, c):
    if c < 0:
        return np.full(theta.shape[0], -999.0)
    return theta[:, c]


def batch_model(refepoch, VLBI, earth_xyz, timing, idx_row, theta):
    """Vectorised positions() for one source: returns (model (N,2E) rad, offsets (N,2E) mas)."""
    dec = _col(theta, idx_row[0]); incl = _col(theta, idx_row[3]); mu_a = _col(theta, idx_row[4])
    mu_d = _col(theta, idx_row[5]); om_asc = _col(theta, idx_row[6]); px = _col(theta, idx_row[7])
    ra = _col(theta, idx_row[8])
    epochs = np.asarray(VLBI['epochs'], dtype=np.float64)
    E = len(epochs)
    N = theta.shape[0]
    dT = (epochs - refepoch) * D2YR                               # (E,)
    cosd = np.cos(dec)[:, None]; sind = np.sin(dec)[:, None]
    cosa = np.cos(ra)[:, None]; sina = np.sin(ra)[:, None]
    dRA = dT[None, :] * mu_a[:, None] / cosd
    dDEC = dT[None, :] * mu_d[:, None]
    X = earth_xyz[:, 0][None, :]; Y = earth_xyz[:, 1][None, :]; Z = earth_xyz[:, 2][None, :]
    pxc = px[:, None]
    pRA = pxc * (X * sina - Y * cosa) / cosd
    pDEC = pxc * (X * cosa * sind + Y * sina * sind - Z * cosd)
    px_on = (px != -999)[:, None]
    oRA = np.where(px_on, dRA + pRA, dRA)
    oDEC = np.where(px_on, dDEC + pDEC, dDEC)
    if timing != {}:
        terms = [reflex_orbit_terms(t, timing) for t in epochs]
        bAU = np.array([t[0] for t in terms]); cth = np.array([t[1] for t in terms]); sth = np.array([t[2] for t in terms])
        offset = bAU[None, :] * pxc
        Om = om_asc * DEG2RAD
        sO = np.sin(Om)[:, None]; cO = np.cos(Om)[:, None]
        ci = np.cos(incl)[:, None]
        m0 = offset * cth[None, :]; m1 = offset * sth[None, :]
        b0 = sO * m0 + (cO * ci) * m1          # (matr1*matr2)[0,:] . matr3
        b1 = cO * m0 + (-(sO * ci)) * m1       # (matr1*matr2)[1,:] . matr3 ; sO*(-ci) = -(sO*ci)
        rRA = b0 / np.cos(timing['decj'] * DEG2RAD)
        rDEC = b1
        rf_on = px_on & ((incl != -999) & (om_asc != -999))[:, None]
        oRA = np.where(rf_on, oRA + rRA, oRA)
        oDEC = np.where(rf_on, oDEC + rDEC, oDEC)
    m_ra = ra[:, None] + oRA * MAS2RAD
    m_dec = dec[:, None] + oDEC * MAS2RAD
    return np.concatenate([m_ra, m_dec], axis=1), np.concatenate([oRA, oDEC], axis=1)


def batch_errs(VLBI, idx_row, theta):
    """Vectorised adjust_errs_with_efac -> (N, 2E)."""
    efac = _col(theta, idx_row[1]); efad = _col(theta, idx_row[2])
    N = theta.shape[0]
    E = int(len(VLBI['errs']) / 2)
    efads = np.ones((N, 2 * E))
    efads[:, E:] = np.where(efad != -999, efad, 1.0)[:, None] * np.ones(E)[None, :]
    out = np.empty((N, 2 * E))
    on = efac != -999
    if on.any():
        sq = (VLBI['errs_random']) ** 2 + (efac[on, None] * efads[on] * VLBI['errs_sys']) ** 2
        out[on] = sq ** 0.5
    if (~on).any():
        out[~on] = (VLBI['errs'] ** 2) ** 0.5
    return out


def batch_log_likelihood(refepoch, LoD_timing, LoD_VLBI, shares, earth_xyz, theta,
                         DoD_additional_constraints=None, a1dot_constraints=False, chunk=4096):
    """Vectorised Gaussianlikelihood.log_likelihood over rows of theta (N, P),
    columns in get_parameters_from_shares(shares) order."""
    names, idx = column_table(shares)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    N = theta.shape[0]
    out = np.zeros(N)
    gauss = {} if DoD_additional_constraints is None else DoD_additional_constraints['sin_incl_Gaussian_constraints']
    if a1dot_constraints != False:
        a1on, a1mus, a1sig = parse_a1dot_constraints(a1dot_constraints)
    else:
        a1on = False
    for lo in range(0, N, chunk):
        th = theta[lo:lo + chunk]
        log_p = np.zeros(th.shape[0])
        for i in range(len(LoD_VLBI)):
            model, _ = batch_model(refepoch, LoD_VLBI[i], earth_xyz[i], LoD_timing[i], idx[i], th)
            res = LoD_VLBI[i]['radecs'][None, :] - model
            errs_new = batch_errs(LoD_VLBI[i], idx[i], th)
            log_p += -0.5 * np.sum((res / errs_new) ** 2, axis=1)
            log_p += -1 * np.sum(np.log(errs_new), axis=1)
        if a1on:
            cols = []
            for i in range(len(LoD_timing)):
                if LoD_timing[i] != {}:
                    a1 = LoD_timing[i]['a1'] / C_M_S
                    cols.append(a1dot_pm_formalisim(a1, _col(th, idx[i][3]), _col(th, idx[i][4]),
                                                    _col(th, idx[i][5]), _col(th, idx[i][6])))
            modeled = np.stack(cols, axis=1)
            log_p += -0.5 * np.sum(((modeled - a1mus[None, :]) / a1sig[None, :]) ** 2, axis=1)
        for parameter in gauss:
            mu, sg = gauss[parameter]
            res = mu - np.sin(th[:, names.index(parameter)])
            log_p += -0.5 * (res / sg) ** 2
        out[lo:lo + chunk] = log_p
    return out


# ---------------------------------------------------------------------------
# Input readers (simulate.py:459-497, 211-241; reflex_motion.py:56-104; priors.py:305-341;
# others.py:35-70)
# ---------------------------------------------------------------------------
def dms_str2deg(string):
    """others.py:35-61"""
    if ':' in string:
        a = string.split(':')
    else:
        a = string.split(' ')
    sign1 = 1
    if a[0].strip().startswith('-'):
        sign1 = -1
    b = [float(item) for item in a]
    if b[0] < 0:
        sign = -1
    elif b[0] == 0 and sign1 == -1:
        sign = -1
    else:
        sign = 1
    b[0] = abs(b[0])
    i = len(b)
    degree = b[i - 1]
    while i != 1:
        degree /= 60
        i -= 1
        degree += b[i - 1]
    degree *= sign
    return degree


def _mjd_of_jan1(year):
    y = year - 1
    return 365 * y + y // 4 - y // 100 + y // 400 - 678575   # MJD of Jan 1, `year` (Gregorian)


def decyear2mjd(epoch):
    """simulate.py:488-497.  The decimal-year branch calls astropy Time(format='decimalyear')
    (absent here, UNPINNED): MJD(Jan 1 of year) + fraction * days_in_year, UTC quasi-JD."""
    if epoch > 10000:
        return epoch
    year = int(math.floor(epoch))
    start = _mjd_of_jan1(year)
    ndays = _mjd_of_jan1(year + 1) - start
    return float(start + (epoch - year) * ndays)


def readpmparin(pmparin):
    """simulate.py:459-486 -> dict of sorted arrays (the astropy Table is replaced by a dict)."""
    rows = []
    for line in open(pmparin).readlines():
        if line.count(':') == 4 and (not line.strip().startswith('#')):
            epoch, RA, errRA, DEC, errDEC = line.strip().split(' ')
            epoch = decyear2mjd(float(epoch.strip()))
            DEC = dms_str2deg(DEC.strip())
            DEC *= np.pi / 180.
            RA = dms_str2deg(RA.strip())
            RA *= 15 * np.pi / 180.
            errRA = float(errRA.strip())
            errRA *= 15 * np.pi / 180. / 3600.
            errDEC = float(errDEC.strip())
            errDEC *= np.pi / 180. / 3600
            rows.append((epoch, RA, errRA, DEC, errDEC))
    arr = np.array(rows, dtype=np.float64).reshape(-1, 5)
    order = np.argsort(arr[:, 0], kind='stable')
    arr = arr[order]
    return {'epoch': arr[:, 0].copy(), 'RA': arr[:, 1].copy(), 'errRA': arr[:, 2].copy(),
            'DEC': arr[:, 3].copy(), 'errDEC': arr[:, 4].copy()}


def create_list_of_dict_VLBI(pmparins, pmparin_preliminaries=None):
    """simulate.py:211-241"""
    out = []
    for pmparin in pmparins:
        t = readpmparin(pmparin)
        out.append({'epochs': np.array(t['epoch']),
                    'radecs': np.concatenate([t['RA'], t['DEC']]),
                    'errs': np.concatenate([t['errRA'], t['errDEC']])})
    if pmparin_preliminaries is not None:
        if len(pmparin_preliminaries) != len(pmparins):
            raise SystemExit(1)
        for i in range(len(pmparins)):
            t = readpmparin(pmparin_preliminaries[i])
            if not (np.array(t['epoch']) == out[i]['epochs']).all():
                raise SystemExit(1)
            errs_random = np.concatenate([t['errRA'], t['errDEC']])
            out[i]['errs_random'] = errs_random
            errs = out[i]['errs']
            out[i]['errs_sys'] = (errs ** 2 - errs_random ** 2) ** 0.5
    return out


def read_parfile(parfile):
    """reflex_motion.py:56-104 -> plain-float dict (see orbital_dict for the unit convention).
    Substring keyword match and the split on exactly 8 spaces are reference behaviour."""
    lines = open(parfile).readlines()
    keywords_needed = ['DECJ', 'PB ', 'ECC', 'A1 ', 'T0', 'OM ', 'OMDOT', 'OM_ASC', 'PBDOT', 'A1DOT', 'SINI']
    parameters = [kw.strip().lower() for kw in keywords_needed]
    d = {}
    for line in lines:
        for i in range(len(keywords_needed)):
            if keywords_needed[i] in line:
                alist = [element.strip() for element in line.split(' ' * 8)]
                alist = [x for x in alist if x != '']
                if parameters[i] == 'decj':
                    d[parameters[i]] = alist[1]
                else:
                    d[parameters[i]] = float(alist[1])
    d['pb'] *= 1.0
    d['a1'] *= C_M_S
    d['t0'] *= 1.0
    d['om'] *= 1.0
    if 'a1dot' in d:
        d['a1dot'] *= C_M_S
    d['decj'] = dms_str2deg(d['decj'])     # KeyError where the reference would query psrqpy
    return d


def create_list_of_dict_timing(parfiles):
    """simulate.py:200-209"""
    return [read_parfile(p) if p != '' else {} for p in parfiles]


def read_inits(initsfile):
    """priors.py:305-341"""
    lines = open(initsfile).readlines()
    dict_limits = {}
    DoD = {'sin_incl_Gaussian_constraints': {}, 'sin_incl_limits_constraints': {}}
    for line in lines:
        if not line.startswith('#'):
            for keyword in ['ra', 'dec', 'mu_a', 'mu_d', 'px', 'incl', 'om_asc', 'efac', 'efad']:
                if keyword in line:
                    parameter = line.split(':')[0].strip()
                    limits = [limit.strip() for limit in line.split(':')[-1].strip().split(',')]
                    limits[0] = float(limits[0])
                    limits[1] = float(limits[1])
                    dict_limits[parameter] = limits[:3]
                    try:
                        limits[3] = float(limits[3])
                        limits[4] = float(limits[4])
                        if (limits[5] == 'Sine_Gaussian') and ('incl' in parameter):
                            DoD['sin_incl_Gaussian_constraints'][parameter] = limits[3:5]
                        elif (limits[5] == 'Sine_limits') and ('incl' in parameter):
                            DoD['sin_incl_limits_constraints'][parameter] = limits[3:5]
                    except IndexError:
                        pass
    return dict_limits, DoD
