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


@pytest.mark.gpu
def test_device_summary_equals_host_summary():
    """brief_summary on a CUDA chain (device sort / order statistics / correlations / periodic estimate)
    writes the same text as the host path -- on the reference's recorded posterior and on a large chain."""
    import torch
    from sterne_b200 import summary
    names = ['dec_0', 'incl_0', 'mu_a_0', 'mu_d_0', 'om_asc_0', 'px_0', 'ra_0']
    rng = np.random.default_rng(7)
    for n in (1501, 200000):
        cols = [0.37 + 4e-10 * rng.standard_normal(n), 1.2 + 0.15 * rng.standard_normal(n),
                -3.2 + 0.05 * rng.standard_normal(n), 7.9 + 0.06 * rng.standard_normal(n),
                (350.0 + 25.0 * rng.standard_normal(n)) % 360.0, 1.25 + 0.04 * rng.standard_normal(n),
                5.1 + 3e-10 * rng.standard_normal(n)]
        cols[3] = cols[3] + 0.5 * (cols[2] + 3.2)
        host = np.stack(cols, axis=1)
        med_h, text_h = summary.brief_summary(names, host)
        med_d, text_d = summary.brief_summary(names, torch.from_numpy(host).cuda())
        assert text_d == text_h, n
        assert med_d == med_h
