"""Posterior summaries (sterne_b200/summary.py) against the reference's own output recorded in
tests/golden/golden_posterior_summary.json (make_golden.py: unmodified others.py / simulate.py)."""
import json
import os

import numpy as np
import pytest

from sterne_b200 import summary, synthetic
from conftest import ROOT


@pytest.fixture(scope='module')
def post():
    with open(os.path.join(ROOT, 'tests', 'golden', 'golden_posterior_summary.json')) as f:
        return json.load(f)


def test_estimators_match_reference(post):
    for e in post['estimators']:
        x, ang = np.array(e['x']), np.array(e['ang'])
        assert summary.sample2median(x) == e['median']
        assert list(summary.sample2median_range(x, 1)) == e['range1']
        assert list(summary.sample2median_range(x, 2)) == e['range2']
        assert list(summary.periodic_sample2estimate(ang)) == e['periodic']
    for d in post['deg2dms']:
        assert summary.deg_float2dms(d['deg']) == d['dms']


def test_brief_summary_text_is_the_references(post, tmp_path):
    p = tmp_path / 'posterior_samples.dat'
    p.write_text(post['posterior_samples'])
    names, data = summary.read_posterior_samples(str(p))
    _, text = summary.brief_summary(names[:-2], data[:, :-2])
    assert post['bayesian_estimates'].startswith(text)          # everything up to the chi-square lines


@pytest.mark.gpu
def test_full_summary_file_matches_reference(post, tmp_path):
    p = tmp_path / 'posterior_samples.dat'
    p.write_text(post['posterior_samples'])
    import sterne_b200 as sb
    case = synthetic.config3('dots')
    files = synthetic.write_case_files(case, str(tmp_path / 'inputs'))      # the files the reference read
    VLBI = sb.create_list_of_dict_VLBI(files['pmparins'], files['preliminaries'])
    timing = sb.create_list_of_dict_timing(files['parfiles'])
    like = sb.Gaussianlikelihood(case.refepoch, timing, VLBI, case.shares, None, None,
                                 earth_xyz=[np.array(e) for e in post['earth_xyz']], mode='strict')
    med, chi, rchi, out = summary.make_a_summary_of_bayesian_inference(str(p), like)
    assert open(out).read() == post['bayesian_estimates']
    like.close()
