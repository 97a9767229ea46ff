"""Stand-in for astropy.constants (see ../README.md)."""
from . import units as u
c = u.Quantity(299792458.0, u.m / u.s)
