"""Native JPL-ephemeris reader (sterne_b200/jpleph.py) against files written in the same
export layout from a known orbit.  No real DE421 file exists in this image (see module docstring)."""
import os

import numpy as np
import pytest

import sterne_b200 as sb
from sterne_b200.jpleph import JPLEphemerisProvider, write_jpl_file

AU_KM = 149597870.7
EMRAT = 81.30056


def _emb_km(jd):
    return sb.AnalyticEarthProvider()(jd) * AU_KM


def _moon_km(jd):
    ph = 2 * np.pi * (jd - 2451550.1) / 27.321661
    r = 384400.0 * (1 + 0.0549 * np.cos(ph * 0.99))
    inc = np.radians(23.4 + 5.1)
    return np.stack([r * np.cos(ph), r * np.sin(ph) * np.cos(inc), r * np.sin(ph) * np.sin(inc)], axis=1)


@pytest.fixture(scope='module', params=['<', '>'])
def eph_file(request, tmp_path_factory):
    path = str(tmp_path_factory.mktemp('eph') / ('synth_%s.eph' % ('le' if request.param == '<' else 'be')))
    ncoeff = write_jpl_file(path, _emb_km, _moon_km, 2457000.5, 2457000.5 + 32 * 40, byteorder=request.param)
    return path, ncoeff


def test_reader_reproduces_the_fitted_orbit(eph_file):
    path, ncoeff = eph_file
    prov = JPLEphemerisProvider(path)
    assert prov.ncoeff == ncoeff and prov.ss[2] == 32.0 and prov.denum == 421
    assert abs(prov.emrat - EMRAT) < 1e-12 and prov.au_km == AU_KM
    rng = np.random.default_rng(0)
    jd = np.concatenate([2457000.5 + rng.uniform(0, 32 * 40, 3000),
                         [2457000.5, 2457000.5 + 32 * 40, 2457032.5, 2457016.5, 2457000.5 + 4.0]])   # span ends, record and sub-interval edges
    got = prov(jd)
    want = (_emb_km(jd) - _moon_km(jd) / (1 + EMRAT)) / AU_KM
    assert np.abs(got - want).max() < 2e-11          # AU: ~3 m, the Chebyshev fit error of this synthetic file
    assert got.shape == (len(jd), 3)
    with pytest.raises(ValueError):
        prov(np.array([2456000.0]))


def test_provider_plugs_into_the_problem_compiler(eph_file):
    path, _ = eph_file
    prov = JPLEphemerisProvider(path)
    epochs = 57000.0 + np.linspace(1.0, 1200.0, 17)
    xyz = sb.earth_vectors(prov, epochs)             # JD = MJD + 2400000.5, as positions.py:305
    want = (_emb_km(epochs + 2400000.5) - _moon_km(epochs + 2400000.5) / (1 + EMRAT)) / AU_KM
    assert np.abs(xyz - want).max() < 2e-11


def test_default_provider_uses_tempo2_file_when_novas_is_absent(eph_file, monkeypatch, tmp_path):
    path, _ = eph_file
    d = tmp_path / 'T2runtime' / 'ephemeris'
    d.mkdir(parents=True)
    os.symlink(path, str(d / 'DE421.1950.2050'))
    monkeypatch.setenv('TEMPO2', str(tmp_path))
    from sterne_b200 import ephemeris
    prov = ephemeris.default_provider()
    assert isinstance(prov, JPLEphemerisProvider)
    monkeypatch.delenv('TEMPO2')
    with pytest.raises(RuntimeError):
        ephemeris.default_provider()


def test_rejects_non_ephemeris_files(tmp_path):
    p = tmp_path / 'junk.bin'
    p.write_bytes(b'\0' * 10000)
    with pytest.raises(ValueError):
        JPLEphemerisProvider(str(p))


@pytest.mark.parametrize('order', ['<', '>'])
def test_real_de421_record_layout(order, tmp_path):
    """The published DE405 / DE421 GROUP 1050 table: 13 items, NCOEFF 1018, 32-day records from
    JD 2433264.5, EMB at offset 231 (13 x 2), Moon at 441 (13 x 8); all other bodies poisoned."""
    from sterne_b200.jpleph import DE421_GROUP_1050, DE421_NCOEFF, DE421_JD_START
    path = str(tmp_path / 'DE421.layout')
    n_rec = 24
    ncoeff = write_jpl_file(path, _emb_km, _moon_km, DE421_JD_START, DE421_JD_START + 32.0 * n_rec, byteorder=order,
                            layout='de421')
    assert ncoeff == DE421_NCOEFF == 1018
    assert os.path.getsize(path) == (2 + n_rec) * 1018 * 8
    prov = JPLEphemerisProvider(path)
    assert prov.order == order and prov.ncoeff == 1018 and prov.denum == 421
    assert prov.ipt == DE421_GROUP_1050
    assert prov.ipt[2] == (231, 13, 2) and prov.ipt[9] == (441, 13, 8) and prov.ipt[12] == (899, 10, 4)
    assert prov.ss == (DE421_JD_START, DE421_JD_START + 32.0 * n_rec, 32.0)
    rng = np.random.default_rng(1)
    jd = np.concatenate([DE421_JD_START + rng.uniform(0, 32.0 * n_rec, 2000),
                         DE421_JD_START + np.array([0.0, 16.0, 4.0, 32.0, 32.0 * n_rec, 32.0 * n_rec - 1e-9])])
    got = prov(jd)
    want = (_emb_km(jd) - _moon_km(jd) / (1 + EMRAT)) / AU_KM
    assert np.abs(got - want).max() < 2e-11          # a read from any other body's slots would return ~1e30 / AU
    # MJD 33282.0 = JD 2433282.5 is the first epoch of the file name's "1950" (1950-01-01); it lies in record 0
    assert np.abs(prov(np.array([2433282.5])) - (_emb_km(np.array([2433282.5])) - _moon_km(np.array([2433282.5])) / (1 + EMRAT)) / AU_KM).max() < 2e-11
