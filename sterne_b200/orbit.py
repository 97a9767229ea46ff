"""Parameter-independent part of the binary reflex motion, evaluated once per epoch
on the host (north_star: "orbital phase from the parfile's PB/T0/ECC/OM/A1 (Kepler
solve)").  Follows reflex_motion.py:106-126,168-200 including its two quirks:
the fixed-point Kepler loop stops at |du| <= 1e-5 rad, and the true anomaly uses
((1+e)/(1-e))**2 rather than the square root.  The (Omega_asc, incl, px)-dependent
rotation (reflex_motion.py:201-214) runs on the device.
"""
import numpy as np

from .constants import D2YR, RAD2DEG, DEG2RAD, D2S, M2AU


def solve_u(e, c, precision=1e-5):
    """Eccentric anomaly by fixed-point iteration u <- e sin u + c, first guess
    c/(1-e), stop when successive iterates differ by <= precision
    (reflex_motion.py:106-126).  Returns (u, iterations)."""
    u = c / (1 - e)
    prev = float('inf')
    iterations = 0
    while abs(u - prev) > precision:
        prev = u
        u = e * np.sin(prev) + c
        iterations += 1
    return u, iterations


def orbit_terms(epochs_mjd, timing):
    """Per-epoch (b_AU, cos(theta), sin(theta)) for a parfile dict (readers.read_parfile).

    The unit conversions astropy does implicitly in the reference are explicit here:
    mean motion in rad/d; OMDOT [deg/yr] / n [rad/d] * A_u [rad] -> deg via 1/365.25;
    theta = omega [deg] + A_u [rad -> deg]; A1DOT*c [m/s] * dt [d -> s]; metres -> AU.
    """
    ecc, t0, pb, om0, a0 = timing['ecc'], timing['t0'], timing['pb'], timing['om'], timing['a1']
    omdot = timing.get('omdot', 0.0)
    pbdot = timing.get('pbdot', 0)
    adot = timing.get('a1dot', None)
    epochs = np.asarray(epochs_mjd, dtype=np.float64)
    out = np.empty((epochs.shape[0], 3), dtype=np.float64)
    for i, t in enumerate(epochs):
        dt = t - t0
        n = 2 * np.pi / pb - np.pi * pbdot * dt / (pb ** 2)                 # :187
        u1 = solve_u(ecc, n * dt)[0]                                         # :189
        A_u = 2 * np.arctan(((1 + ecc) / (1 - ecc)) ** 2 * np.tan(u1 / 2))   # :191
        omega = om0 + ((omdot / n) * A_u) * D2YR                             # :192-193
        theta = (omega + A_u * RAD2DEG) * DEG2RAD                            # :194, cos/sin of a deg angle
        a1 = a0 + (adot * dt) * D2S if adot is not None else a0              # :195-198
        out[i, 0] = (a1 * (1 - ecc * np.cos(u1))) * M2AU                     # :199-200
        out[i, 1] = np.cos(theta)
        out[i, 2] = np.sin(theta)
    return out
