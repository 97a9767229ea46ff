"""Input readers: pmpar.in (+ .preliminary), PSRCAT-style .par, .inits.

Same file formats and the same numbers as the reference's readers
(simulate.py:200-241,459-497; reflex_motion.py:56-104; priors.py:305-341;
others.py:35-70).  astropy is not required: tables are plain dicts of numpy arrays,
units are plain floats (see orbital-dict convention below).  Where the reference
prints and calls sys.exit(), these raise ValueError / KeyError.
"""
import math
import numpy as np

from .constants import C_M_S


def dms_str2deg(string):
    """'dd:mm:ss.s' (or blank separated) -> dd.ddd, sign taken from the leading field
    ('-00:...' stays negative).  Same arithmetic order as others.py:35-61 so the
    result is bit-identical: ((ss/60 + mm)/60 + dd) * sign."""
    fields = string.split(':') if ':' in string else string.split(' ')
    negative = fields[0].strip().startswith('-')
    vals = [float(f) for f in fields]
    sign = -1 if (vals[0] < 0 or (vals[0] == 0 and negative)) else 1
    vals[0] = abs(vals[0])
    acc = vals[-1]
    for v in reversed(vals[:-1]):
        acc /= 60
        acc += v
    return acc * sign


def dms2deg(array):
    """others.py:62-70"""
    if isinstance(array, str):
        return dms_str2deg(array)
    return np.array([dms_str2deg(s) for s in array], dtype=np.float64)


def _mjd_of_jan1(year):
    y = year - 1
    return 365 * y + y // 4 - y // 100 + y // 400 - 678575


def decyear2mjd(epoch):
    """simulate.py:488-497: values > 10000 are already MJD.  Decimal years follow
    astropy's 'decimalyear' definition: MJD(1 Jan) + fraction * days in that year."""
    if epoch > 10000:
        return epoch
    year = int(math.floor(epoch))
    start = _mjd_of_jan1(year)
    return float(start + (epoch - year) * (_mjd_of_jan1(year + 1) - start))


def readpmparin(pmparin):
    """simulate.py:459-486.  A data line is a non-comment line with exactly four ':';
    fields are separated by single spaces.  Returns columns sorted by epoch:
    epoch [MJD], RA [rad], errRA [rad, no cos(dec)], DEC [rad], errDEC [rad]."""
    rows = []
    with open(pmparin) as f:
        for line in f:
            if line.count(':') != 4 or line.strip().startswith('#'):
                continue
            fields = line.strip().split(' ')
            if len(fields) != 5:
                raise ValueError('%s: expected "epoch RA errRA DEC errDEC" separated by single spaces: %r'
                                 % (pmparin, line))
            epoch, RA, errRA, DEC, errDEC = fields
            dec = dms_str2deg(DEC.strip())
            dec *= np.pi / 180.
            ra = dms_str2deg(RA.strip())
            ra *= 15 * np.pi / 180.
            era = float(errRA.strip())
            era *= 15 * np.pi / 180. / 3600.
            edec = float(errDEC.strip())
            edec *= np.pi / 180. / 3600
            rows.append((decyear2mjd(float(epoch.strip())), ra, era, dec, edec))
    table = np.array(rows, dtype=np.float64).reshape(-1, 5)
    table = table[np.argsort(table[:, 0], kind='stable')]
    return {name: np.ascontiguousarray(table[:, i])
            for i, name in enumerate(('epoch', 'RA', 'errRA', 'DEC', 'errDEC'))}


def create_list_of_dict_VLBI(pmparins, pmparin_preliminaries=None):
    """simulate.py:211-241: per source {'epochs', 'radecs' = [RA..., DEC...],
    'errs' = [errRA..., errDEC...]} and, with .preliminary files, 'errs_random' and
    'errs_sys' = sqrt(errs^2 - errs_random^2)."""
    out = []
    for path in pmparins:
        t = readpmparin(path)
        out.append({'epochs': t['epoch'].copy(),
                    'radecs': np.concatenate([t['RA'], t['DEC']]),
                    'errs': np.concatenate([t['errRA'], t['errDEC']])})
    if pmparin_preliminaries is not None:
        if len(pmparin_preliminaries) != len(pmparins):
            raise ValueError('The number of pmpar.in.preliminary files has to match that of pmpar.in files.')
        for d, path in zip(out, pmparin_preliminaries):
            t = readpmparin(path)
            if t['epoch'].shape != d['epochs'].shape or not (t['epoch'] == d['epochs']).all():
                raise ValueError('Epochs of the pmpar.in.preliminary files should match those of the pmpar.in files.')
            d['errs_random'] = np.concatenate([t['errRA'], t['errDEC']])
            d['errs_sys'] = (d['errs'] ** 2 - d['errs_random'] ** 2) ** 0.5
    return out


_PAR_KEYWORDS = ('DECJ', 'PB ', 'ECC', 'A1 ', 'T0', 'OM ', 'OMDOT', 'OM_ASC', 'PBDOT', 'A1DOT', 'SINI')


def read_parfile(parfile):
    """reflex_motion.py:56-104 -> plain-float orbital dict:
        pb [d], ecc, a1 [m] (= A1[lt-s]*c), t0 [MJD], om [deg], decj [deg],
        optional omdot [deg/yr], pbdot, a1dot [m/s] (= A1DOT*c), om_asc [deg], sini.
    Keyword matching is by substring and values are split on runs of exactly 8 spaces,
    as in the reference.  DECJ is required here (the reference falls back to a network
    query through psrqpy)."""
    d = {}
    with open(parfile) as f:
        for line in f:
            for kw in _PAR_KEYWORDS:
                if kw in line:
                    fields = [x.strip() for x in line.split(' ' * 8)]
                    fields = [x for x in fields if x != '']
                    name = kw.strip().lower()
                    d[name] = fields[1] if name == 'decj' else float(fields[1])
    for key in ('pb', 'a1', 't0', 'om'):
        if key not in d:
            raise KeyError(key)
    if 'decj' not in d:
        raise KeyError("'decj' is essential but is missing from %s" % parfile)
    d['a1'] *= C_M_S
    if 'a1dot' in d:
        d['a1dot'] *= C_M_S
    d['decj'] = dms_str2deg(d['decj'])
    return d


def create_list_of_dict_timing(parfiles):
    """simulate.py:200-209: '' -> {} (no reflex motion for that source)."""
    return [read_parfile(p) if p != '' else {} for p in parfiles]


_INITS_KEYWORDS = ('ra', 'dec', 'mu_a', 'mu_d', 'px', 'incl', 'om_asc', 'efac', 'efad')


def read_inits(initsfile):
    """priors.py:305-341 -> (limits {name: [a, b, kind]}, constraints dict of two dicts).
    Line format: `name: a,b,Kind[,c,d,Sine_Gaussian|Sine_limits]`."""
    limits_by_name = {}
    constraints = {'sin_incl_Gaussian_constraints': {}, 'sin_incl_limits_constraints': {}}
    with open(initsfile) as f:
        for line in f:
            if line.startswith('#') or not any(kw in line for kw in _INITS_KEYWORDS):
                continue
            name = line.split(':')[0].strip()
            fields = [x.strip() for x in line.split(':')[-1].strip().split(',')]
            fields[0] = float(fields[0])
            fields[1] = float(fields[1])
            limits_by_name[name] = fields[:3]
            if len(fields) >= 6 and 'incl' in name:
                pair = [float(fields[3]), float(fields[4])]
                if fields[5] == 'Sine_Gaussian':
                    constraints['sin_incl_Gaussian_constraints'][name] = pair
                elif fields[5] == 'Sine_limits':
                    constraints['sin_incl_limits_constraints'][name] = pair
    return limits_by_name, constraints


def plain_timing(dict_timing):
    """Orbital dict with plain floats.  Accepts what the reference's read_parfile returns
    (astropy Quantities: value taken in the Quantity's own unit, which the reference fixes
    at reflex_motion.py:76-93 -- pb d, a1 m, t0 d, om deg, omdot deg/yr, a1dot m/s, decj deg)
    as well as this package's plain-float dicts."""
    out = {}
    for k, v in dict_timing.items():
        out[k] = float(getattr(v, 'value', v))
    return out
