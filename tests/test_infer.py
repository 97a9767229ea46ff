"""End-to-end: input files -> priors -> fused GPU sampler -> posterior_samples.dat -> bayesian_estimates.dat."""
import numpy as np
import pytest

import sterne_b200 as sb
from sterne_b200 import infer, synthetic


def test_argument_splitting_follows_the_reference():
    assert infer.split_args('a.pmpar.in', 'a.par', 'b.pmpar.in', '') == (['a.pmpar.in', 'b.pmpar.in'], ['a.par', ''])
    with pytest.raises(ValueError):
        infer.split_args('a.pmpar.in', 'a.par', 'b.pmpar.in')
    assert infer.default_shares(2, False)[1] == [-1, -1] and infer.default_shares(2, True)[7] == [0, 0]


@pytest.mark.gpu
def test_inference_from_files_recovers_truth(tmp_path):
    case = synthetic.config2()                                   # EFAC, 6 parameters
    files = synthetic.write_case_files(case, str(tmp_path / 'in'))
    prov = sb.AnalyticEarthProvider()                            # the provider the synthetic data were made with
    res = infer.run(case.refepoch, files['inits'], files['pmparins'][0], files['parfiles'][0],
                    shares=case.shares, pmparin_preliminaries=files['preliminaries'], nwalkers=512,
                    iterations=1200, outdir=str(tmp_path / 'out'), earth_provider=prov, seed=3)
    names, data = sb.summary.read_posterior_samples(res['samplefile'])
    assert names == res['keys'] + ['log_likelihood', 'log_prior'] and data.shape == (res['n_samples'], 8)
    assert 0.1 < res['acceptance_fraction'] < 0.9
    for k in case.names:
        if k.startswith('efac'):
            continue
        col = data[:, names.index(k)]
        assert abs(np.median(col) - case.truth[k]) < 4.0 * col.std(), k
    text = open(res['estimates']).read()
    assert 'reduced chi-square' in text and 'r__' in text and text.startswith('#Medians')
    assert res['rchsq'] > 0
