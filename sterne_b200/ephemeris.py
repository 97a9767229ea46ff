"""Earth-position providers for the parallax term.

The reference calls novas' solarsystem(MJD + 2400000.5, 3, 0) on JPL DE421 once per
epoch per likelihood call (positions.py:299-305).  The Earth vector does not depend
on the sampled parameters, so here it is evaluated ONCE per epoch on the host and
uploaded with the epoch table.  A provider is any callable
    provider(jd_tdb: np.ndarray[E]) -> np.ndarray[E, 3]   (AU, ICRS, w.r.t. the SSB).
"""
import math
import os
import numpy as np

from .constants import MJD2JD, DEG2RAD


class NovasDE421Provider(object):
    """The reference's own source of Earth vectors: novas + $TEMPO2's DE421 file
    (positions.py:299-305).  Usable only where novas and TEMPO2 are installed."""

    def __init__(self, ephemeris_path=None):
        import novas.compat.solsys as solsys                # noqa: deferred, optional
        from novas.compat.eph_manager import ephem_open
        if ephemeris_path is None:
            tempo2_dir = os.getenv('TEMPO2')
            if tempo2_dir is None:
                raise RuntimeError('TEMPO2 not installed or its environment variable unset')
            ephemeris_path = os.path.join(tempo2_dir, 'T2runtime/ephemeris/DE421.1950.2050')
        ephem_open(ephemeris_path)
        self._solsys = solsys

    def __call__(self, jd):
        return np.array([self._solsys.solarsystem(float(t), 3, 0)[0] for t in np.atleast_1d(jd)],
                        dtype=np.float64).reshape(-1, 3)


class TableProvider(object):
    """Earth vectors supplied by the caller for exactly the epochs of the data
    (e.g. exported once from novas/DE421, jplephem or tempo2)."""

    def __init__(self, jd, xyz):
        self.jd = np.asarray(jd, dtype=np.float64).ravel()
        self.xyz = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
        if self.jd.shape[0] != self.xyz.shape[0]:
            raise ValueError('jd and xyz lengths differ')
        self._map = {float(t): i for i, t in enumerate(self.jd)}

    def __call__(self, jd):
        try:
            rows = [self._map[float(t)] for t in np.atleast_1d(jd)]
        except KeyError as err:
            raise KeyError('no Earth vector tabulated for JD %r' % (err.args[0],))
        return self.xyz[rows]


class AnalyticEarthProvider(object):
    """Keplerian Earth-Moon-barycentre orbit (Standish's J2000 mean elements with
    secular rates), rotated to the equatorial frame.  Accurate to ~1e-4 AU: meant for
    SYNTHETIC workloads and tests only, not for science; real data should use
    NovasDE421Provider or TableProvider."""

    A0, A1 = 1.00000261, 0.00000562
    E0, E1 = 0.01671123, -0.00004392
    I0, I1 = -0.00001531, -0.01294668
    L0, L1 = 100.46457166, 35999.37244981
    W0, W1 = 102.93768193, 0.32327364
    OBLIQUITY_DEG = 23.43928

    def __call__(self, jd):
        jd = np.atleast_1d(np.asarray(jd, dtype=np.float64))
        T = (jd - 2451545.0) / 36525.0
        a = self.A0 + self.A1 * T
        e = self.E0 + self.E1 * T
        inc = (self.I0 + self.I1 * T) * DEG2RAD
        L = self.L0 + self.L1 * T
        varpi = self.W0 + self.W1 * T
        M = np.remainder(L - varpi + 180.0, 360.0) - 180.0
        M = M * DEG2RAD
        w = varpi * DEG2RAD                       # node longitude is 0 for the EMB
        Ecc = M + e * np.sin(M)
        for _ in range(8):
            Ecc = Ecc - (Ecc - e * np.sin(Ecc) - M) / (1.0 - e * np.cos(Ecc))
        xp = a * (np.cos(Ecc) - e)
        yp = a * np.sqrt(1.0 - e * e) * np.sin(Ecc)
        cw, sw, ci, si = np.cos(w), np.sin(w), np.cos(inc), np.sin(inc)
        x_ecl = cw * xp - sw * yp
        y_ecl = (sw * xp + cw * yp) * ci
        z_ecl = (sw * xp + cw * yp) * si
        eps = self.OBLIQUITY_DEG * DEG2RAD
        ce, se = math.cos(eps), math.sin(eps)
        return np.stack([x_ecl, ce * y_ecl - se * z_ecl, se * y_ecl + ce * z_ecl], axis=1)


def tempo2_ephemeris_path():
    tempo2_dir = os.getenv('TEMPO2')
    if tempo2_dir is None:
        return None
    path = os.path.join(tempo2_dir, 'T2runtime/ephemeris/DE421.1950.2050')
    return path if os.path.exists(path) else None


def default_provider():
    """The reference's own source first (novas + $TEMPO2's DE421 file, positions.py:299-305);
    if novas is not installed but the DE421 file is there, this package's native reader of
    the same file (jpleph.JPLEphemerisProvider); otherwise raises -- callers that want the
    analytic orbit must ask for it."""
    try:
        return NovasDE421Provider()
    except Exception as err:
        path = tempo2_ephemeris_path()
        if path is not None:
            from .jpleph import JPLEphemerisProvider
            return JPLEphemerisProvider(path)
        raise RuntimeError(
            'no Earth ephemeris available: novas/DE421 could not be opened (%s). Pass '
            'earth_provider=TableProvider(...) with vectors for your epochs, or '
            'AnalyticEarthProvider() for synthetic data.' % (err,))


def earth_vectors(provider, epochs_mjd):
    """(E, 3) Earth vectors at JD = MJD + 2400000.5, the argument the reference passes
    (positions.py:305; the MJD is used as TDB without a time-scale conversion)."""
    jd = np.asarray(epochs_mjd, dtype=np.float64) + MJD2JD
    xyz = np.asarray(provider(jd), dtype=np.float64).reshape(-1, 3)
    if xyz.shape[0] != jd.shape[0]:
        raise ValueError('ephemeris provider returned %d vectors for %d epochs' % (xyz.shape[0], jd.shape[0]))
    return xyz
