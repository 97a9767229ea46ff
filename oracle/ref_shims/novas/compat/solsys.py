"""Stand-in for novas.compat.solsys: Earth vector from an installed provider (see ../../README.md)."""
_provider = None


def set_provider(fn):
    """fn(jd_tdb) -> (X, Y, Z) in AU, ICRS, w.r.t. the solar-system barycentre."""
    global _provider
    _provider = fn


def solarsystem(tjd, body, origin):
    if body != 3 or origin != 0:
        raise NotImplementedError('stand-in handles Earth w.r.t. SSB only')
    xyz = _provider(tjd)
    return (tuple(float(v) for v in xyz), (0.0, 0.0, 0.0))
