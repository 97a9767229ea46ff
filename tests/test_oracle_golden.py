"""Pins the CPU oracle (oracle/sterne_oracle.py) against outputs of the UNMODIFIED
reference recorded in tests/golden/golden_reference.json (see make_golden.py)."""
import numpy as np

from oracle import sterne_oracle as O
from conftest import rel_err


def _inputs(case, files):
    VLBI = O.create_list_of_dict_VLBI(files['pmparins'], files['preliminaries'])
    timing = O.create_list_of_dict_timing(files['parfiles'])
    limits, DoD = O.read_inits(files['inits'])
    earth = [np.array(e) for e in case['earth_xyz']]
    return VLBI, timing, limits, DoD, earth


def test_readers_match_reference(golden, golden_dir):
    for case in golden['cases']:
        VLBI, timing, limits, DoD, _ = _inputs(case, golden_dir[case['name']])
        for mine, ref in zip(VLBI, case['ref']['VLBI']):
            assert set(mine) == set(ref)
            for k in ref:
                assert np.array_equal(mine[k], np.array(ref[k])), (case['name'], k)
        for mine, ref in zip(timing, case['ref']['timing']):
            assert set(mine) == set(ref), case['name']
            for k in ref:
                assert mine[k] == ref[k], (case['name'], k)
        assert limits == case['ref']['limits']
        assert DoD == case['ref']['DoD']


def test_parameter_names_match_reference(golden):
    for item in golden['shares_names']:
        assert list(O.get_parameters_from_shares(item['shares'])) == item['names']
    for case in golden['cases']:
        assert list(O.get_parameters_from_shares(case['shares'])) == case['ref']['names']


def test_solve_u_matches_reference(golden):
    for item in golden['solve_u']:
        u, it = O.solve_u(item['e'], item['c'])
        assert u == item['u'] and it == item['iterations']


def test_dms_and_decyear(golden):
    for item in golden['dms2deg']:
        assert O.dms_str2deg(item['s']) == item['deg']
    for item in golden['decyear2mjd']:
        assert O.decyear2mjd(item['epoch']) == item['mjd']


def test_reflex_motion_matches_reference(golden, golden_dir):
    n = 0
    for case in golden['cases']:
        _, timing, _, _, _ = _inputs(case, golden_dir[case['name']])
        for item in case['ref']['reflex_motion']:
            out = O.reflex_motion(item['epoch'], timing[item['source']], item['incl'], item['om_asc'], item['px'])
            # bit-for-bit: same operations in the same order
            assert out[0] == item['out'][0] and out[1] == item['out'][1], (case['name'], item)
            n += 1
    assert n >= 12


def test_positions_and_errs_match_reference(golden, golden_dir):
    for case in golden['cases']:
        VLBI, timing, _, _, earth = _inputs(case, golden_dir[case['name']])
        names = case['ref']['names']
        for t, th in enumerate(case['theta']):
            params = dict(zip(names, th))
            for i in range(len(VLBI)):
                model = O.positions(case['refepoch'], VLBI[i]['epochs'], earth[i], timing[i], i, params)
                assert np.array_equal(model, np.array(case['ref']['positions'][t][i])), (case['name'], t, i)
                errs = O.adjust_errs_with_efac(VLBI[i], params, i)
                assert np.array_equal(errs, np.array(case['ref']['errs_new'][t][i])), (case['name'], t, i)


def test_log_likelihood_matches_reference(golden, golden_dir):
    for case in golden['cases']:
        VLBI, timing, _, DoD, earth = _inputs(case, golden_dir[case['name']])
        like = O.Gaussianlikelihood(case['refepoch'], timing, VLBI, case['shares'], earth, DoD,
                                    case['a1dot_constraints'])
        got = like.log_likelihood_rows(np.array(case['theta']))
        ref = np.array(case['ref']['log_likelihood'])
        assert np.array_equal(got, ref), (case['name'], got - ref)
        if case['ref']['a1dot_pm']:
            for t, th in enumerate(case['theta']):
                a = O.calculate_a1dot_pm(timing, dict(zip(case['ref']['names'], th)))
                assert np.array_equal(a, np.array(case['ref']['a1dot_pm'][t]))


def test_vectorised_oracle_equals_scalar_oracle(golden, golden_dir):
    """The numpy-vectorised restatement (used at N in the thousands and as the CPU
    baseline) agrees with the reference-shaped scalar path to the last few ulps."""
    for case in golden['cases']:
        VLBI, timing, _, DoD, earth = _inputs(case, golden_dir[case['name']])
        theta = np.array(case['theta'])
        vec = O.batch_log_likelihood(case['refepoch'], timing, VLBI, case['shares'], earth, theta, DoD,
                                     case['a1dot_constraints'])
        ref = np.array(case['ref']['log_likelihood'])
        assert rel_err(vec, ref).max() < 1e-13, (case['name'], rel_err(vec, ref))
