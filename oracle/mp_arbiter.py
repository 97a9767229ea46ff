"""
HIGH-PRECISION ARBITER -- TEST INFRASTRUCTURE ONLY (same rules as sterne_oracle.py: only
tests/, tools/ parity reports, __graft_entry__.smoke() and bench.py's checker legs import it).

Evaluates the reference's log-likelihood formula (simulate.py:362-384 with
positions.py:62-121,266-310, reflex_motion.py:201-214, simulate.py:303-320,
kopeikin_effects.py:11-45) in 50-digit arithmetic on the SAME float64 inputs the float64
implementations see: epoch tables, Earth vectors, the per-epoch orbit terms
(b_AU, cos theta, sin theta -- parameter independent, taken from the float64 oracle) and the
unit factors.  What it arbitrates is therefore only the rounding of the per-vector arithmetic.

Use: where the CUDA result and the float64 oracle disagree by more than the 1e-10 contract,
`exact_log_likelihood` says how far each of them is from the exact value of the formula; an
exceedance is "explained" when both are rounding noise of the same size around it.
"""
import mpmath as mp
import numpy as np

from . import sterne_oracle as O

DPS = 50


def _mpf(x):
    return mp.mpf(float(x))


def exact_log_likelihood(case, theta_row, names=None):
    """Exact (50-digit) value of the reference's log L for one parameter vector of a
    sterne_b200.synthetic.SyntheticCase-like object (refepoch, LoD_timing, LoD_VLBI, shares,
    earth_xyz, DoD_additional_constraints, a1dot_constraints)."""
    mp.mp.dps = DPS
    cols, idx = O.column_table(case.shares)
    if names is not None:
        assert list(names) == list(cols)
    th = [_mpf(v) for v in theta_row]
    K = _mpf(O.MAS2RAD)
    D2YR = _mpf(O.D2YR)
    DEG = _mpf(O.DEG2RAD)
    sentinel = mp.mpf(-999)

    def col(c):
        return th[c] if c >= 0 else sentinel

    logp = mp.mpf(0)
    for s, vlbi in enumerate(case.LoD_VLBI):
        dec, efac, efad, incl, mu_a, mu_d, om_asc, px, ra = [col(c) for c in idx[s]]
        timing = case.LoD_timing[s]
        E = len(vlbi['epochs'])
        px_on = px != sentinel
        rf_on = px_on and incl != sentinel and om_asc != sentinel and timing != {}
        sd, cd, sa, ca = mp.sin(dec), mp.cos(dec), mp.sin(ra), mp.cos(ra)
        if rf_on:
            sO, cO, ci = mp.sin(om_asc * DEG), mp.cos(om_asc * DEG), mp.cos(incl)
            cdj = mp.cos(_mpf(timing['decj']) * DEG)
        efac_on = efac != sentinel                       # simulate.py:316
        for i in range(E):
            dT = (_mpf(vlbi['epochs'][i]) - _mpf(case.refepoch)) * D2YR
            X, Y, Z = [_mpf(c) for c in case.earth_xyz[s][i]]
            oRA = dT * mu_a / cd
            oDEC = dT * mu_d
            if px_on:
                oRA += px * (X * sa - Y * ca) / cd
                oDEC += px * (X * ca * sd + Y * sa * sd - Z * cd)
                if rf_on:
                    bAU, cth, sth = [_mpf(t) for t in O.reflex_orbit_terms(float(vlbi['epochs'][i]), timing)]
                    off = bAU * px
                    m0, m1 = off * cth, off * sth
                    oRA += (sO * m0 + cO * ci * m1) / cdj
                    oDEC += cO * m0 - sO * ci * m1
            for half, ref, off in ((0, ra, oRA), (1, dec, oDEC)):
                j = half * E + i
                res = _mpf(vlbi['radecs'][j]) - (ref + off * K)
                if efac_on:
                    f = efac * (efad if (half == 1 and efad != sentinel) else 1)
                    var = _mpf(vlbi['errs_random'][j]) ** 2 + (f * _mpf(vlbi['errs_sys'][j])) ** 2
                else:
                    var = _mpf(vlbi['errs'][j]) ** 2
                logp += -mp.mpf('0.5') * res * res / var - mp.log(var) / 2
    a1dot = getattr(case, 'a1dot_constraints', False)
    if a1dot is not False and a1dot is not None:
        on, mus, sigmas = O.parse_a1dot_constraints(a1dot)
        if on:
            modeled = []
            for s, timing in enumerate(case.LoD_timing):
                if timing != {}:
                    dec, efac, efad, incl, mu_a, mu_d, om_asc, px, ra = [col(c) for c in idx[s]]
                    a1 = _mpf(timing['a1']) / _mpf(O.C_M_S)
                    mu = mp.sqrt(mu_a ** 2 + mu_d ** 2)
                    v = mu * mp.sin(mp.atan2(mu_a, mu_d) - om_asc * DEG) / mp.tan(incl) * a1 * _mpf(O.MASYR2RADS)
                    modeled.append(v)
            nb, nm = len(modeled), len(mus)
            for k in range(max(nb, nm)):
                z = (modeled[k if nb > 1 else 0] - _mpf(mus[k if nm > 1 else 0])) / _mpf(sigmas[k if nm > 1 else 0])
                logp += -mp.mpf('0.5') * z * z
    DoD = getattr(case, 'DoD_additional_constraints', None)
    if DoD:
        for name, (mu, sg) in DoD.get('sin_incl_Gaussian_constraints', {}).items():
            z = (_mpf(mu) - mp.sin(th[cols.index(name)])) / _mpf(sg)
            logp += -mp.mpf('0.5') * z * z
    return logp


def parity_stats(case, theta, got, want, rtol=1e-10, floor=None, max_arbitrated=64, noise_sample=32, seed=0):
    """Pure-relative parity of `got` (CUDA) against `want` (float64 oracle), with every
    exceedance of `rtol` arbitrated in 50-digit arithmetic.

    Returns a dict: max / count of |got - want| / |want| (no floor), and for the exceedances
    (at most `max_arbitrated`, the worst ones) the distance of BOTH results from the exact value.
    An exceedance is EXPLAINED when
      * |want| is below `floor` (log L = -sum ln sigma - chi^2/2 cancels there: the summed terms are
        of size `floor`, so the relative error of the sum is not bounded by that of its terms), and
      * the CUDA result is rounding noise of the same size as the oracle's own: its distance from the
        exact value is at most 4x the larger of the oracle's distance for that vector and the 90th
        percentile of the oracle's distance over a random sample of the batch."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
        rel = np.abs(got - want) / np.abs(want)
    rel = np.where(got == want, 0.0, rel)
    bad = np.flatnonzero(~(rel <= rtol))
    out = {'n': int(len(got)), 'rtol': rtol, 'floor': floor, 'max_rel': float(np.nanmax(rel)) if len(rel) else 0.0,
           'median_rel': float(np.nanmedian(rel)) if len(rel) else 0.0, 'count_above': int(len(bad)),
           'max_abs': float(np.abs(got - want).max()) if len(got) else 0.0,
           'abs_logl_range': [float(np.abs(want).min()), float(np.abs(want).max())] if len(want) else None,
           'exceedances': [], 'explained': True}
    if len(bad) == 0:
        return out
    rng = np.random.default_rng(seed)
    pick = rng.choice(len(got), size=min(noise_sample, len(got)), replace=False)
    noise = np.array([abs(float(exact_log_likelihood(case, theta[k])) - want[k]) for k in pick])
    noise90 = float(np.quantile(noise, 0.9))
    out['oracle_abs_err_p90'] = noise90
    worst = bad[np.argsort(-rel[bad])][:max_arbitrated]
    for k in worst:
        exact = float(exact_log_likelihood(case, theta[k]))
        e_gpu, e_orc = abs(got[k] - exact), abs(want[k] - exact)
        ok = (floor is not None and abs(want[k]) < floor) and e_gpu <= 4.0 * max(e_orc, noise90)
        out['exceedances'].append({'index': int(k), 'logl': float(want[k]), 'rel': float(rel[k]),
                                   'abs_diff': float(abs(got[k] - want[k])), 'gpu_abs_err_vs_exact': float(e_gpu),
                                   'oracle_abs_err_vs_exact': float(e_orc), 'explained': bool(ok)})
        out['explained'] = out['explained'] and bool(ok)
    out['arbitrated'] = len(worst)
    return out
