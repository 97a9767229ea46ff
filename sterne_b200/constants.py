"""Unit factors used by the host precompute and mirrored in csrc/sterne_b200.cu.

The reference obtains these from astropy.units at run time (positions.py:108,119-120;
reflex_motion.py:76-93,169-213; kopeikin_effects.py:37-45).  astropy is not a
dependency here; the factors are written out from astropy's unit definitions.
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

ROOTS = ('dec', 'efac', 'efad', 'incl', 'mu_a', 'mu_d', 'om_asc', 'px', 'ra')   # priors.py:270
