"""Stand-in for novas.compat.eph_manager (see ../../README.md)."""
opened = []


def ephem_open(path=None):
    opened.append(path)
    return (2433264.5, 2469808.5, 421)
