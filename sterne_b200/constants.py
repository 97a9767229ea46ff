"""Unit factors used by the host precompute and handed to the kernels (StnProblem.mas2rad ...).

The reference obtains these from astropy.units / astropy.constants at run time
(positions.py:108,119-120; reflex_motion.py:76-93,169-213; kopeikin_effects.py:37-45).
astropy is NOT a dependency of this package: the factors are written out below from astropy's
unit definitions.  When astropy IS importable (a user's environment that also runs the
reference), its own values are taken instead, so that every factor agrees with the reference's
to the last bit -- e.g. whether (u.mas).to(u.rad) rounds to pi/648000000 or to
(pi/180)/3600*0.001 (1 ulp apart) is decided by astropy, not by this file.  `SOURCE` says which.
No astropy exists in the build image or on the GPU boxes, so the astropy branch is exercised by
tests only through a stand-in (tests/test_host.py) -- agreement with a real astropy is UNVERIFIED.
"""
import math

SENTINEL = -999.0                       # positions.py:59 "parameter switched off"

D2YR = 86400.0 / 31557600.0             # day -> Julian year
MAS2RAD = (math.pi / 180.0) / 3600.0 * 0.001
DEG2RAD = math.pi / 180.0
RAD2DEG = 180.0 / math.pi
C_M_S = 299792458.0                     # speed of light, m/s
AU_M = 149597870700.0                   # astronomical unit, m
M2AU = 1.0 / AU_M
D2S = 86400.0
YR2D = 365.25
MASYR2RADS = MAS2RAD / (YR2D * D2S)     # mas/yr -> rad/s
MJD2JD = 2400000.5                      # positions.py:305
SOURCE = 'built-in (astropy unit definitions written out)'


def factors_from_astropy(u, c):
    """The same factors computed by astropy itself, exactly as the reference's expressions do."""
    return {
        'D2YR': float((u.d).to(u.yr)),                            # positions.py:108
        'MAS2RAD': float((u.mas).to(u.rad)),                      # positions.py:119-120
        'DEG2RAD': float((u.deg).to(u.rad)),
        'RAD2DEG': float((u.rad).to(u.deg)),
        'C_M_S': float(c.c.to(u.m / u.s).value),                  # reflex_motion.py:80
        'AU_M': float((u.AU).to(u.m)),
        'M2AU': float((u.m).to(u.AU)),                            # reflex_motion.py:200
        'D2S': float((u.d).to(u.s)),
        'YR2D': float((u.yr).to(u.d)),
        'MASYR2RADS': float((u.mas / u.yr).to(u.rad / u.s)),      # kopeikin_effects.py:45
    }


def use_astropy(u=None, c=None):
    """Replaces the module-level factors by astropy's (called at import when astropy is present).
    Returns True when it did."""
    try:
        if u is None:
            import astropy.units as u
        if c is None:
            import astropy.constants as c
        vals = factors_from_astropy(u, c)
    except Exception:
        return False
    g = globals()
    for k, v in vals.items():
        if not (v > 0.0 and abs(v / g[k] - 1.0) < 1e-12):         # a stand-in or broken install: keep ours
            return False
    g.update(vals)
    g['SOURCE'] = 'astropy %s' % getattr(__import__('astropy'), '__version__', '?') if u.__name__.startswith('astropy') else 'injected'
    return True


use_astropy()

ROOTS = ('dec', 'efac', 'efad', 'incl', 'mu_a', 'mu_d', 'om_asc', 'px', 'ra')   # priors.py:270
