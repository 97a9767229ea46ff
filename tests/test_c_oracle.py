"""The C restatement of the oracle (oracle/c/sterne_oracle.c) against the numpy oracle and the
reference's recorded outputs.  Two independent restatements agreeing bit for bit (same libm /
numpy sin-cos on this host) guard against a slip in either."""
import numpy as np
import pytest

from oracle import c_oracle, sterne_oracle as O
from sterne_b200 import synthetic
from conftest import rel_err


def _golden_inputs(case, files):
    VLBI = O.create_list_of_dict_VLBI(files['pmparins'], files['preliminaries'])
    timing = O.create_list_of_dict_timing(files['parfiles'])
    _, DoD = O.read_inits(files['inits'])
    return VLBI, timing, DoD, [np.array(e) for e in case['earth_xyz']]


def test_c_oracle_reproduces_recorded_reference_values(golden, golden_dir):
    for case in golden['cases']:
        VLBI, timing, DoD, earth = _golden_inputs(case, golden_dir[case['name']])
        co = c_oracle.COracle(case['refepoch'], timing, VLBI, case['shares'], earth, DoD, case['a1dot_constraints'])
        got = co.log_likelihood_batch(np.array(case['theta']))
        ref = np.array(case['ref']['log_likelihood'])
        assert rel_err(got, ref).max() < 1e-13, (case['name'], got, ref)


@pytest.mark.parametrize('mk', [synthetic.config2, lambda: synthetic.config3('dots'), synthetic.config4,
                                lambda: synthetic.config5('C')])
def test_c_oracle_equals_numpy_oracle(mk):
    case = mk()
    theta = case.sample_theta(96, seed=17)
    theta[::7, case.names.index('px_0') if 'px_0' in case.names else 0] = -999.0 if 'px_0' in case.names else theta[::7, 0]
    a = c_oracle.from_case(case).log_likelihood_batch(theta)
    b = O.batch_log_likelihood(case.refepoch, case.LoD_timing, case.LoD_VLBI, case.shares, case.earth_xyz, theta,
                               case.DoD_additional_constraints, case.a1dot_constraints)
    assert rel_err(a, b).max() < 1e-13
    for nt in (1, 3):
        assert np.array_equal(c_oracle.from_case(case).log_likelihood_batch(theta, nthreads=nt), a)
