"""CPU-side tests of the product's host logic (no GPU): readers, parameter naming and
routing, orbit precompute, problem compiler, and the C-ABI library's exported symbols."""
import os
import re

import numpy as np
import pytest

import sterne_b200 as sb
from sterne_b200 import _lib, orbit, problem, shares as sh, synthetic
from oracle import sterne_oracle as O
from conftest import ROOT


def test_readers_match_reference(golden, golden_dir):
    for case in golden['cases']:
        f = golden_dir[case['name']]
        VLBI = sb.create_list_of_dict_VLBI(f['pmparins'], f['preliminaries'])
        timing = sb.create_list_of_dict_timing(f['parfiles'])
        limits, DoD = sb.read_inits(f['inits'])
        for mine, ref in zip(VLBI, case['ref']['VLBI']):
            assert set(mine) == set(ref)
            for k in ref:
                assert np.array_equal(mine[k], np.array(ref[k])), (case['name'], k)
        for mine, ref in zip(timing, case['ref']['timing']):
            assert mine == ref, case['name']
        assert limits == case['ref']['limits'] and DoD == case['ref']['DoD']


def test_dms_decyear_names(golden):
    for item in golden['dms2deg']:
        assert sb.dms2deg(item['s']) == item['deg']
    for item in golden['decyear2mjd']:
        assert sb.decyear2mjd(item['epoch']) == item['mjd']
    for item in golden['shares_names']:
        assert list(sb.get_parameters_from_shares(item['shares'])) == item['names']


def test_reader_error_behaviour(tmp_path):
    p = tmp_path / 'x.pmpar.in'
    p.write_text('57000.0 12:00:00.0  0.0001 30:00:00.0 0.001\n')      # double space: reference raises too
    with pytest.raises(ValueError):
        sb.readpmparin(str(p))
    q = tmp_path / 'x.par'
    q.write_text('PB              1.5\nECC             0.1\n')
    with pytest.raises(KeyError):
        sb.read_parfile(str(q))
    with pytest.raises(ValueError):
        sb.get_parameters_from_shares([[0]] * 8)


def test_column_table_equals_reference_routing(golden):
    """column_table() == filter_dictionary_of_parameter_with_index + sorted positional pick."""
    for item in golden['shares_names']:
        names = item['names']
        S = len(item['shares'][0])
        idx = sh.column_table(names, S)
        values = {k: float(i + 1) for i, k in enumerate(names)}
        for s in range(S):
            FP = O.filter_dictionary_of_parameter_with_index(values, s)
            Ps = sorted(FP.keys())
            for r in range(9):
                want = FP[Ps[r]]
                got = -999 if idx[s, r] < 0 else values[names[idx[s, r]]]
                assert got == want, (names, s, r)
            mine = sh.filter_dictionary_of_parameter_with_index(values, s)
            assert mine == FP


def test_orbit_terms_equal_oracle():
    for variant in ('plain', 'wide', 'dots'):
        case = synthetic.config3(variant)
        t = case.LoD_timing[0]
        ep = case.LoD_VLBI[0]['epochs']
        got = orbit.orbit_terms(ep, t)
        for i, e in enumerate(ep):
            want = O.reflex_orbit_terms(e, t)
            assert tuple(got[i]) == tuple(float(w) for w in want)
    assert orbit.solve_u(0.3, 500.123) == O.solve_u(0.3, 500.123)


def test_problem_compiler_tables():
    case = synthetic.config4()
    pb = problem.CompiledProblem(case.refepoch, case.LoD_timing, case.LoD_VLBI, case.shares,
                                 earth_xyz=case.earth_xyz)
    assert pb.n_params == 16 and pb.n_sources == 5
    assert pb.evals_per_sample == sum(10 + s for s in range(5))
    for s, src in enumerate(pb.sources):
        E = src['n_epochs']
        v = case.LoD_VLBI[s]
        assert src['efac_on'] and src['binary']
        assert np.array_equal(src['fast'][:, 4], v['radecs'][:E])
        assert np.array_equal(src['fast'][:, 6], v['errs_random'][:E] ** 2)
        assert np.array_equal(src['strict'][:, 11], v['errs_sys'][E:])
        assert np.array_equal(src['fast'][:, 1:4], case.earth_xyz[s])
    # column permutation through `keys`
    keys = list(reversed(pb.names))
    pb2 = problem.CompiledProblem(case.refepoch, case.LoD_timing, case.LoD_VLBI, case.shares,
                                  earth_xyz=case.earth_xyz, names=keys)
    assert all(keys[pb2.idx[s, r]] == pb.names[pb.idx[s, r]] for s in range(5) for r in range(9) if pb.idx[s, r] >= 0)
    with pytest.raises(ValueError):
        problem.CompiledProblem(case.refepoch, case.LoD_timing, case.LoD_VLBI, case.shares,
                                earth_xyz=case.earth_xyz, names=keys[:-1])
    # EFAC requested but no .preliminary errors: KeyError as in simulate.py:317
    v = [{k: d[k] for k in ('epochs', 'radecs', 'errs')} for d in case.LoD_VLBI]
    with pytest.raises(KeyError):
        problem.CompiledProblem(case.refepoch, case.LoD_timing, v, case.shares, earth_xyz=case.earth_xyz)


def test_a1dot_broadcast_rules():
    case = synthetic.config4()
    mk = lambda a: problem.CompiledProblem(case.refepoch, case.LoD_timing, case.LoD_VLBI, case.shares,
                                           earth_xyz=case.earth_xyz, a1dot_constraints=a)
    p = mk([[1e-14, 2e-15], [], [], [], []])
    assert list(p.a1_source) == [0, 1, 2, 3, 4] and list(p.a1_mu) == [1e-14] * 5     # (5,) - (1,)
    p = mk([[1e-14, 2e-15]] * 5)
    assert list(p.a1_source) == [0, 1, 2, 3, 4]
    with pytest.raises(ValueError):
        mk([[1e-14, 2e-15], [1e-14, 2e-15], [], [], []])                             # (5,) - (2,)
    assert len(mk([[], [], [], [], []]).a1_source) == 0


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'sterne_b200.h')).read()
    declared = set(re.findall(r'\b(stn_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations found'
    lib = _lib.load()
    assert declared == {name for name, _, _ in _lib.SYMBOLS}
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.stn_version() == _lib.ABI_VERSION == int(re.search(r'#define STN_ABI_VERSION (\d+)', header).group(1))


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    case = synthetic.config1()
    with pytest.raises(_lib.StnError) as err:
        case.likelihood()
    assert 'no CPU path' in str(err.value) or 'CUDA' in str(err.value)


def test_analytic_earth_orbit_is_sane():
    prov = sb.AnalyticEarthProvider()
    jd = 2451545.0 + np.linspace(0, 3652, 500)
    xyz = prov(jd)
    r = np.linalg.norm(xyz, axis=1)
    assert r.min() > 0.98 and r.max() < 1.02
    # obliquity: z/y amplitude ratio ~ tan(23.44 deg)
    assert abs(np.abs(xyz[:, 2]).max() / np.abs(xyz[:, 1]).max() - np.tan(np.radians(23.43928))) < 5e-3
    # around the March equinox the Earth is near -x (Sun at RA 0 as seen from the Earth)
    k = np.argmin(np.abs(jd - 2451623.8))
    assert xyz[k, 0] < -0.98


def test_files_round_trip_to_memory_case(tmp_path):
    case = synthetic.config4()
    f = synthetic.write_case_files(case, str(tmp_path))
    VLBI = sb.create_list_of_dict_VLBI(f['pmparins'], f['preliminaries'])
    for a, b in zip(VLBI, case.LoD_VLBI):
        assert np.allclose(a['epochs'], b['epochs'], atol=1e-6, rtol=0)
        assert np.abs(a['radecs'] - b['radecs']).max() < 1e-13         # 1e-10 s of time
        assert np.allclose(a['errs'], b['errs'], rtol=1e-10)
    timing = sb.create_list_of_dict_timing(f['parfiles'])
    assert timing[0]['pb'] == case.LoD_timing[0]['pb'] and timing[0]['a1'] == case.LoD_timing[0]['a1']


def test_parfile_substring_quirks_match_reference(golden, tmp_path):
    """reflex_motion.py:59-75 matches keywords by substring: an ECCDOT line overwrites ECC, a KOM line
    sets OM.  Both the product reader and the oracle reader must reproduce that."""
    q = golden['parfile_quirks']
    p = tmp_path / 'quirky.par'
    p.write_text(q['text'])
    assert sb.read_parfile(str(p)) == q['ref']
    assert O.read_parfile(str(p)) == q['ref']
    assert q['ref']['ecc'] == 1.5e-15 and q['ref']['om'] == 45.0


def test_unit_factors_are_injectable(monkeypatch):
    """The reference takes its unit factors from astropy at run time; this package takes astropy's own
    values when astropy is importable and hands them to the library (StnProblem.mas2rad ...).  No astropy
    exists here, so a stand-in whose mas->rad is the OTHER candidate double (pi/648000000, 1 ulp from the
    built-in (pi/180)/3600*0.001) is injected."""
    import math
    import types
    from sterne_b200 import constants, problem
    other = math.pi / 648000000.0
    assert other != constants.MAS2RAD and abs(other / constants.MAS2RAD - 1) < 1e-15

    class Unit(object):
        def __init__(self, name, scale):
            self.name, self.scale = name, scale

        def to(self, target):
            table = {('d', 'yr'): constants.D2YR, ('mas', 'rad'): other, ('deg', 'rad'): constants.DEG2RAD,
                     ('rad', 'deg'): constants.RAD2DEG, ('AU', 'm'): constants.AU_M, ('m', 'AU'): constants.M2AU,
                     ('d', 's'): constants.D2S, ('yr', 'd'): constants.YR2D,
                     ('mas/yr', 'rad/s'): other / (constants.YR2D * constants.D2S), ('c', 'm/s'): constants.C_M_S}
            v = table[(self.name, target.name)]
            return types.SimpleNamespace(value=v) if self.name == 'c' else v

        def __truediv__(self, o):
            return Unit(self.name + '/' + o.name, None)

    u = types.SimpleNamespace(__name__='fake_units', **{n: Unit(n, None) for n in ('d', 'yr', 'mas', 'rad', 'deg', 'AU', 'm', 's')})
    c = types.SimpleNamespace(c=Unit('c', None))
    saved = {k: getattr(constants, k) for k in ('D2YR', 'MAS2RAD', 'DEG2RAD', 'RAD2DEG', 'C_M_S', 'AU_M', 'M2AU', 'D2S',
                                                'YR2D', 'MASYR2RADS', 'SOURCE')}
    try:
        assert constants.use_astropy(u, c)
        assert constants.MAS2RAD == other and constants.SOURCE == 'injected'
        case = synthetic.config1()
        pb, keep = problem.CompiledProblem(case.refepoch, case.LoD_timing, case.LoD_VLBI, case.shares,
                                           earth_xyz=case.earth_xyz).as_ctypes()
        assert pb.mas2rad == other and pb.deg2rad == constants.DEG2RAD
        assert pb.masyr2rads == other / (constants.YR2D * constants.D2S)
    finally:
        for k, v in saved.items():
            setattr(constants, k, v)
    # a factor that is off by more than rounding (a broken or foreign 'astropy') is refused
    bad = types.SimpleNamespace(__name__='fake_units', **{n: Unit(n, None) for n in ('d', 'yr', 'mas', 'rad', 'deg', 'AU', 'm', 's')})
    bad.mas = types.SimpleNamespace(name='mas', to=lambda t: 1.0, __truediv__=None)
    assert not constants.use_astropy(bad, c)
    assert constants.MAS2RAD == saved['MAS2RAD']


def test_routing_properties_on_random_shares():
    """hypothesis: for ANY shares matrix (1..13 sources, values -1..4 per row) the names equal the oracle's
    restatement of priors.get_parameters_from_shares, every name parses back to its sources and root,
    and the integer column table routes exactly the value the reference's dict filter + sorted positional
    pick would hand to position() -- including the -999 fill for absent roots and two-digit source indices."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=150, deadline=None)
    @given(st.integers(min_value=1, max_value=13).flatmap(
        lambda S: st.lists(st.lists(st.integers(min_value=-1, max_value=4), min_size=S, max_size=S),
                           min_size=9, max_size=9)))
    def check(shares):
        S = len(shares[0])
        names = list(sb.get_parameters_from_shares(shares))
        assert names == list(O.get_parameters_from_shares(shares))
        assert len(set(names)) == len(names)
        for root_i, row in enumerate(shares):                      # every present (source, root) is named exactly once
            for s, v in enumerate(row):
                owners = [n for n in names if sh.parameter_name_to_pmparin_indice(n) ==
                          (sh.parameter_name_to_pmparin_indice(n)[0], sb.ROOTS[root_i])
                          and s in sh.parameter_name_to_pmparin_indice(n)[0]]
                assert len(owners) == (1 if v >= 0 else 0)
        if not names:
            return
        idx = sh.column_table(names, S)
        values = {k: 10.0 + i for i, k in enumerate(names)}
        for s in range(S):
            FP = O.filter_dictionary_of_parameter_with_index(values, s)
            Ps = sorted(FP.keys())
            assert len(Ps) == 9
            for r in range(9):
                got = -999 if idx[s, r] < 0 else values[names[idx[s, r]]]
                assert got == FP[Ps[r]], (shares, s, r)
            assert sh.filter_dictionary_of_parameter_with_index(values, s) == FP
        # sources that share a value share the column
        for r, row in enumerate(shares):
            for a in range(S):
                for b in range(S):
                    if row[a] >= 0 and row[a] == row[b]:
                        assert idx[a, r] == idx[b, r] >= 0

    check()


def test_c_slices_equal_python_shard_bounds():
    """stn_multi_slice (one process, several GPUs) and sharding.shard_bounds (one process per GPU) cut a batch
    the same way -- the C function is pure arithmetic and runs without a GPU."""
    import ctypes
    from sterne_b200.sharding import shard_bounds
    lib = _lib.load()
    lo, hi = ctypes.c_int64(), ctypes.c_int64()
    rng = np.random.default_rng(0)
    for _ in range(300):
        n = int(rng.integers(0, 1 << 23))
        g = int(rng.integers(1, 9))
        prev = 0
        for part in range(g):
            assert lib.stn_multi_slice(n, g, part, ctypes.byref(lo), ctypes.byref(hi)) == 0
            assert (lo.value, hi.value) == shard_bounds(n, g, part)
            assert lo.value == prev
            prev = hi.value
        assert prev == n
    assert lib.stn_multi_slice(10, 0, 0, ctypes.byref(lo), ctypes.byref(hi)) != 0
    assert lib.stn_multi_slice(10, 2, 2, ctypes.byref(lo), ctypes.byref(hi)) != 0
    assert b'slice' in lib.stn_last_error()
