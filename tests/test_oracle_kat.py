"""Closed-form known-answer tests and an mpmath arbiter for the CPU oracle.
The reference has no tests of its own; these pin the formulas independently of any recorded run."""
import math

import mpmath as mp
import numpy as np

from oracle import sterne_oracle as O
from sterne_b200 import synthetic


def test_proper_motion_only_one_year():
    # refepoch 57000, epoch 57365.25 -> dT = 1 yr exactly; ra = dec = 0 -> cos(dec) = 1
    r, d, o0, o1 = O.position_with_offset(57000.0, 57365.25, (0.3, 0.4, 0.5), 0.0, -999, 1.0, 2.0, -999, -999, 0.0)
    assert (o0, o1) == (1.0, 2.0)
    assert r == 1.0 * O.MAS2RAD and d == 2.0 * O.MAS2RAD


def test_parallax_closed_forms():
    # (X,Y,Z) = (1,0,0), ra = pi/2, dec = 0, px = 1 -> dRA = sin(pi/2) = 1 mas, dDEC = cos(pi/2)*0 - 0 = 0
    out = O.parallax_related_position_offset_from_the_barycentric_frame((1.0, 0.0, 0.0), math.pi / 2, 0.0, 1.0)
    assert out[0] == 1.0 and out[1] == 0.0
    # (0,0,1), dec = 0 -> dDEC = -1 mas
    out = O.parallax_related_position_offset_from_the_barycentric_frame((0.0, 0.0, 1.0), 0.3, 0.0, 1.0)
    assert out[0] == 0.0 and out[1] == -1.0
    # px = -999 switches parallax off: only proper motion remains
    _, _, o0, o1 = O.position_with_offset(57000.0, 57365.25, (1.0, 1.0, 1.0), 0.0, -999, 1.0, 2.0, -999, -999, 0.0)
    assert (o0, o1) == (1.0, 2.0)


def test_reflex_closed_form_circular_orbit():
    # e = 0, t = T0 -> u = 0, A_u = 0, theta = omega0 = 0: b = a1 along the line of nodes.
    # Omega = 90 deg: b0 = offset*sin(90) = offset, b1 = offset*cos(90) ~ 0
    t = O.orbital_dict(pb=10.0, ecc=0.0, a1=10.0, t0=55000.0, om=0.0, decj_deg=0.0)
    out = O.reflex_motion(55000.0, t, 0.7, 90.0, 2.0)
    offset = (10.0 * O.C_M_S) * O.M2AU * 2.0
    assert out[0] == offset and abs(out[1]) < 1e-16 * offset * 2
    # a quarter of an orbit later theta = 90 deg: the offset turns into the cos(incl) direction
    out = O.reflex_motion(55002.5, t, 0.0, 0.0, 2.0)      # Omega = 0, i = 0: b0 = cosO*cosi*sin(th) = offset
    assert abs(out[0] - offset) < 1e-12 * offset and abs(out[1]) < 1e-12 * offset


def test_efac_off_sum_log_sigma_is_bitwise_log_errs():
    rng = np.random.default_rng(0)
    errs = rng.uniform(1e-10, 1e-9, 200)
    d = {'errs': errs}
    out = O.adjust_errs_with_efac(d, {'ra_0': 1.0}, 0)
    assert np.array_equal(out, errs)                       # sqrt(x*x) == x


def test_efac_efad_formula():
    d = {'errs': np.array([5.0, 5.0]), 'errs_random': np.array([3.0, 3.0]), 'errs_sys': np.array([4.0, 4.0])}
    out = O.adjust_errs_with_efac(d, {'efac_0': 2.0, 'efad_0': 0.5}, 0)
    assert out[0] == math.sqrt(9.0 + 64.0) and out[1] == math.sqrt(9.0 + 16.0)


def test_log_likelihood_has_no_2pi_term():
    # one epoch, model == observation: log L = -sum(log sigma) exactly
    v = {'epochs': np.array([57000.0]), 'radecs': np.array([1.0, 0.5]), 'errs': np.array([2e-9, 4e-9])}
    like = O.Gaussianlikelihood(57000.0, [{}], [v], [[0], [-1], [-1], [-1], [0], [0], [-1], [-1], [0]],
                                [np.zeros((1, 3))])
    like.parameters.update({'dec_0': 0.5, 'mu_a_0': 3.0, 'mu_d_0': -2.0, 'ra_0': 1.0})
    assert like.log_likelihood() == -(math.log(2e-9) + math.log(4e-9))


def _mp_loglike(case, theta):
    """The same formulas in 50-digit arithmetic (single non-binary source, EFAC off)."""
    mp.mp.dps = 50
    v = case.LoD_VLBI[0]
    E = len(v['epochs'])
    names = case.names
    p = {k: mp.mpf(float(x)) for k, x in zip(names, theta)}
    K = mp.mpf(O.MAS2RAD)
    logp = mp.mpf(0)
    for i in range(E):
        dT = (mp.mpf(float(v['epochs'][i])) - mp.mpf(case.refepoch)) * mp.mpf(O.D2YR)
        X, Y, Z = [mp.mpf(float(c)) for c in case.earth_xyz[0][i]]
        ra, dec = p['ra_0'], p['dec_0']
        oRA = dT * p['mu_a_0'] / mp.cos(dec) + p['px_0'] * (X * mp.sin(ra) - Y * mp.cos(ra)) / mp.cos(dec)
        oDEC = dT * p['mu_d_0'] + p['px_0'] * (X * mp.cos(ra) * mp.sin(dec) + Y * mp.sin(ra) * mp.sin(dec) - Z * mp.cos(dec))
        for obs, ref, off, err in ((v['radecs'][i], ra, oRA, v['errs'][i]), (v['radecs'][E + i], dec, oDEC, v['errs'][E + i])):
            res = mp.mpf(float(obs)) - (ref + off * K)
            logp += -mp.mpf('0.5') * (res / mp.mpf(float(err))) ** 2 - mp.log(mp.mpf(float(err)))
    return logp


def test_mpmath_arbiter_shows_the_references_own_rounding_floor():
    """The reference's float64 log-L for a 10-epoch problem sits ~1e-9 (relative) from the exact
    value of its own formula, because model = ra + off*K rounds at ulp(ra) ~ 1.8e-7 mas.  This
    is why parity is defined against the reference's ROUNDING ORDER (tolerance 1e-10), not
    against the exact value, and why the fast kernel keeps `obs - (ref + off)`."""
    case = synthetic.config1()
    thetas = case.sample_theta(24, seed=9)
    oracle = O.batch_log_likelihood(case.refepoch, case.LoD_timing, case.LoD_VLBI, case.shares, case.earth_xyz, thetas)
    exact = np.array([float(_mp_loglike(case, th)) for th in thetas])
    rel = np.abs(oracle - exact) / np.abs(exact)
    assert rel.max() < 1e-6                       # the formula is right ...
    assert rel.max() > 1e-11                      # ... and float64 rounding noise is far above 1e-16
    # a mathematically equal rearrangement, (obs - ra) - off*K, is CLOSER to exact than the reference
    v = case.LoD_VLBI[0]
    E = len(v['epochs'])
    names, idx = O.column_table(case.shares)
    model, off = O.batch_model(case.refepoch, v, case.earth_xyz[0], {}, idx[0], thetas)
    ref0 = np.concatenate([np.repeat(thetas[:, idx[0][8]][:, None], E, 1), np.repeat(thetas[:, idx[0][0]][:, None], E, 1)], 1)
    res_alt = (v['radecs'][None, :] - ref0) - off * O.MAS2RAD
    alt = -0.5 * np.sum((res_alt / v['errs']) ** 2, axis=1) - np.sum(np.log(v['errs']))
    assert (np.abs(alt - exact) / np.abs(exact)).max() < rel.max()
