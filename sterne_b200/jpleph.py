"""Native reader of JPL binary planetary ephemerides (DE405 / DE421 ... as shipped in
$TEMPO2/T2runtime/ephemeris/DE421.1950.2050) and Chebyshev evaluator of the Earth's
barycentric position.

Replaces, for the parallax term, the reference's per-epoch-per-call
    ephem_open($TEMPO2/T2runtime/ephemeris/DE421.1950.2050)
    solarsystem(MJD + 2400000.5, 3, 0)[0]                  (positions.py:299-305, novas C library)
with one vectorised evaluation for all epochs, so that neither novas nor TEMPO2's code is
needed -- only the ephemeris file.

STATUS: the image holds no JPL ephemeris file, so this reader is validated against files
written by `write_jpl_file` below (Chebyshev coefficients fitted to a known orbit), both in a
minimal two-body layout and in DE405 / DE421's REAL record layout -- the published GROUP 1050
pointer table with all 13 items, NCOEFF = 1018 doubles per record, 32-day records starting at
JD 2433264.5, Earth-Moon barycentre at offset 231 (13 coefficients x 2 sub-intervals), Moon at
441 (13 x 8), every other body's slots filled with a poison value -- in both byte orders.
Record addressing at the real offsets, sub-interval selection, Chebyshev evaluation, byte
order and the Earth = EMB - Moon/(1 + EMRAT) combination are tested that way.  Agreement with
a REAL `DE421.1950.2050` file remains UNVERIFIED: none exists in this image or on the GPU boxes.

File layout (JPL export format; every record is NCOEFF doubles):
  record 0  title 3x84 bytes | constant names 400x6 bytes | SS[3] = start JD, end JD, days per record
            | NCON int32 | AU (km) | EMRAT | IPT[12][3] int32 | DENUM int32 | LPT[3] int32
  record 1  constant values
  record 2+ data: [JD start, JD end, coefficients...]; body i uses IPT[i] = (1-based offset,
            coefficients per component, sub-intervals per record); positions in km.
Bodies: 0 Mercury .. 2 Earth-Moon barycentre .. 9 Moon (geocentric), 10 Sun, 11 nutations (2 components).
"""
import struct
import numpy as np

_NAMES_END = 252 + 2400
_EMB, _MOON = 2, 9


class JPLEphemerisProvider(object):
    """Earth w.r.t. the solar-system barycentre, ICRS, in AU: provider(jd_tdb[E]) -> (E, 3)."""

    def __init__(self, path):
        with open(path, 'rb') as f:
            raw = f.read()
        self.path = path
        self.order = self._detect_order(raw)
        o = self.order
        self.ss = struct.unpack_from(o + '3d', raw, _NAMES_END)
        (self.ncon,) = struct.unpack_from(o + 'i', raw, _NAMES_END + 24)
        self.au_km, self.emrat = struct.unpack_from(o + '2d', raw, _NAMES_END + 28)
        ipt = struct.unpack_from(o + '36i', raw, _NAMES_END + 44)
        (self.denum,) = struct.unpack_from(o + 'i', raw, _NAMES_END + 188)
        lpt = struct.unpack_from(o + '3i', raw, _NAMES_END + 192)
        self.ipt = [tuple(ipt[3 * i:3 * i + 3]) for i in range(12)] + [tuple(lpt)]
        ncoeff = 2
        for i, (_, ncf, na) in enumerate(self.ipt):
            ncoeff += ncf * na * (2 if i == 11 else 3)
        self.ncoeff = ncoeff
        nrec = len(raw) // (8 * ncoeff)
        if nrec < 3:
            raise ValueError('%s: too short for a JPL ephemeris with %d coefficients per record' % (path, ncoeff))
        data = np.frombuffer(raw, dtype=np.dtype(o + 'f8'), count=nrec * ncoeff).reshape(nrec, ncoeff)
        self.records = np.ascontiguousarray(data[2:]).astype(np.float64)
        if not (abs(self.records[0, 0] - self.ss[0]) < 1e-6 and self.records[0, 1] - self.records[0, 0] == self.ss[2]):
            raise ValueError('%s: first data record does not start at SS[0]; not a JPL export-format file' % path)

    @staticmethod
    def _detect_order(raw):
        for o in ('<', '>'):
            ss = struct.unpack_from(o + '3d', raw, _NAMES_END)
            if 1.0 <= ss[2] <= 400.0 and 1.0e5 < ss[0] < ss[1] < 1.0e7:
                return o
        raise ValueError('not a JPL binary ephemeris (SS header implausible in either byte order)')

    def _body(self, body, jd):
        """(E, 3) position in km of a header body at TDB Julian dates jd."""
        off, ncf, na = self.ipt[body]
        start, end, step = self.ss
        if (jd < start).any() or (jd > end).any():
            raise ValueError('epoch outside the ephemeris span JD %.1f - %.1f' % (start, end))
        nr = np.minimum(((jd - start) / step).astype(np.int64), self.records.shape[0] - 1)
        rec = self.records[nr]                                        # (E, ncoeff)
        t1 = (jd - rec[:, 0]) / step                                  # fraction of the record, [0, 1]
        temp = na * t1
        sub = np.minimum(temp.astype(np.int64), na - 1)               # sub-interval
        tc = 2.0 * (temp - sub) - 1.0                                 # Chebyshev time, [-1, 1]
        # Chebyshev polynomials T_0..T_{ncf-1}(tc)
        T = np.empty((jd.shape[0], ncf))
        T[:, 0] = 1.0
        if ncf > 1:
            T[:, 1] = tc
        for j in range(2, ncf):
            T[:, j] = 2.0 * tc * T[:, j - 1] - T[:, j - 2]
        out = np.empty((jd.shape[0], 3))
        base = (off - 1) + sub * (3 * ncf)
        cols = np.arange(ncf)
        for k in range(3):
            idx = (base + k * ncf)[:, None] + cols[None, :]
            out[:, k] = np.sum(np.take_along_axis(rec, idx, axis=1) * T, axis=1)
        return out

    def __call__(self, jd):
        jd = np.atleast_1d(np.asarray(jd, dtype=np.float64))
        emb = self._body(_EMB, jd)
        moon = self._body(_MOON, jd)
        return (emb - moon / (1.0 + self.emrat)) / self.au_km


# GROUP 1050 of the DE405 / DE421 headers: (offset, coefficients per component, sub-intervals) of
# Mercury, Venus, EMB, Mars, Jupiter, Saturn, Uranus, Neptune, Pluto, Moon, Sun, nutations, librations
DE421_GROUP_1050 = [(3, 14, 4), (171, 10, 2), (231, 13, 2), (309, 11, 1), (342, 8, 1), (366, 7, 1), (387, 6, 1),
                    (405, 6, 1), (423, 6, 1), (441, 13, 8), (753, 11, 2), (819, 10, 4), (899, 10, 4)]
DE421_NCOEFF = 1018
DE421_JD_START, DE421_JD_END = 2433264.5, 2469808.5


def write_jpl_file(path, emb_km, moon_km, jd_start, jd_end, step=32.0, ncf=(13, 13), na=(2, 8), au_km=149597870.7,
                   emrat=81.30056, denum=421, byteorder='<', layout='minimal', poison=1.0e30):
    """Writes a JPL export-format file holding the Earth-Moon barycentre and the geocentric
    Moon, Chebyshev-fitted to the callables emb_km(jd[M]) / moon_km(jd[M]) -> (M, 3) km.
    layout='minimal': only those two bodies.  layout='de421': DE405 / DE421's real record layout
    (DE421_GROUP_1050, 1018 doubles per record; ncf / na are ignored), the slots of every other
    body filled with `poison` so that a wrong offset cannot go unnoticed.
    For tests and for exporting other ephemerides into the format this package reads."""
    from numpy.polynomial import chebyshev as C
    if layout == 'de421':
        ipt = list(DE421_GROUP_1050[:12])
        lpt = DE421_GROUP_1050[12]
        ncoeff = DE421_NCOEFF
        off = lpt[0]
    elif layout == 'minimal':
        ipt = [(3, 0, 0)] * 12
        off = 3
        ipt[_EMB] = (off, ncf[0], na[0])
        off += 3 * ncf[0] * na[0]
        ipt[_MOON] = (off, ncf[1], na[1])
        off += 3 * ncf[1] * na[1]
        ipt = [(off, 0, 0) if p[1] == 0 else p for p in ipt]
        lpt = (off, 0, 0)
        ncoeff = off - 1
    else:
        raise ValueError('layout must be "minimal" or "de421"')
    nrec = int(round((jd_end - jd_start) / step))
    o = byteorder
    header = bytearray(8 * ncoeff)
    header[0:84] = b'synthetic JPL-format ephemeris (sterne_b200.jpleph.write_jpl_file)'.ljust(84)
    struct.pack_into(o + '3d', header, _NAMES_END, jd_start, jd_start + nrec * step, step)
    struct.pack_into(o + 'i', header, _NAMES_END + 24, 0)
    struct.pack_into(o + '2d', header, _NAMES_END + 28, au_km, emrat)
    struct.pack_into(o + '36i', header, _NAMES_END + 44, *[v for p in ipt for v in p])
    struct.pack_into(o + 'i', header, _NAMES_END + 188, denum)
    struct.pack_into(o + '3i', header, _NAMES_END + 192, *lpt)
    if 8 * ncoeff < _NAMES_END + 204:
        raise ValueError('records too short to hold the header: use more coefficients')
    recs = np.full((nrec, ncoeff), poison if layout == 'de421' else 0.0)
    for r in range(nrec):
        t0 = jd_start + r * step
        recs[r, 0], recs[r, 1] = t0, t0 + step
        for (o1, n, a), fn in ((ipt[_EMB], emb_km), (ipt[_MOON], moon_km)):
            for s in range(a):
                lo, hi = t0 + s * step / a, t0 + (s + 1) * step / a
                nodes = np.cos(np.pi * (np.arange(4 * n) + 0.5) / (4 * n))        # oversampled Chebyshev nodes
                xyz = fn(0.5 * (lo + hi) + 0.5 * (hi - lo) * nodes)
                for k in range(3):
                    recs[r, (o1 - 1) + (s * 3 + k) * n:(o1 - 1) + (s * 3 + k + 1) * n] = C.chebfit(nodes, xyz[:, k], n - 1)
    with open(path, 'wb') as f:
        f.write(bytes(header))
        f.write(bytes(8 * ncoeff))                                                # record 1: constants (none)
        f.write(recs.astype(np.dtype(o + 'f8')).tobytes())
    return ncoeff
