"""Stand-in for astropy.time.Time: only format='decimalyear' -> .mjd (see ../README.md)."""
import math


def _mjd_jan1(year):
    y = year - 1
    return 365 * y + y // 4 - y // 100 + y // 400 - 678575


class Time(object):
    def __init__(self, val, format='mjd'):
        if format == 'decimalyear':
            year = int(math.floor(val))
            start = _mjd_jan1(year)
            self.mjd = float(start + (val - year) * (_mjd_jan1(year + 1) - start))
        elif format == 'mjd':
            self.mjd = float(val)
        else:
            raise NotImplementedError(format)
