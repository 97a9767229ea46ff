"""Sampler adaptor: batched priors, pool shim and the built-in stretch-move sampler (CPU),
plus end-to-end sampling of BASELINE configs 1 and 2 on the GPU."""
import numpy as np
import pytest
import torch

from sterne_b200 import synthetic
from sterne_b200.sampler import BoxPrior, BatchPool, EnsembleSampler, log_prob_fn

LIMITS = {'px_0': [0.5, 2.5, 'Uniform'], 'mu_a_0': [-3.0, 0.7, 'Gaussian'], 'incl_0': [0.2, 2.9, 'Sine']}


def test_box_prior_numpy_torch_and_normalisation():
    pr = BoxPrior(LIMITS, {'sin_incl_limits_constraints': {}})
    th = pr.sample(2000, seed=1)
    th[:5, 0] = 3.0                                   # outside Uniform support
    th[5:9, 2] = 3.0                                  # outside Sine support
    a = pr.log_prob(th)
    b = pr.log_prob(torch.from_numpy(th)).numpy()
    assert np.array_equal(np.isfinite(a), np.isfinite(b))
    assert np.allclose(a[np.isfinite(a)], b[np.isfinite(b)], rtol=1e-13)
    assert np.isinf(a[:9]).all() and np.isfinite(a[9:]).all()
    # each 1-D density integrates to 1
    x = np.linspace(0.2, 2.9, 20001)
    grid = np.stack([np.full_like(x, 1.0), np.full_like(x, -3.0), x], axis=1)
    dens = np.exp(pr.log_prob(grid) - pr.log_prob(np.array([[1.0, -3.0, 1.0]])) + np.log(np.sin(1.0))
                  - np.log(np.cos(0.2) - np.cos(2.9)))
    assert abs(np.trapezoid(dens, x) - 1.0) < 1e-6
    # sin(incl) limits constraint
    pc = BoxPrior(LIMITS, {'sin_incl_limits_constraints': {'incl_0': [0.5, 0.9]}})
    pts = np.array([[1.0, -3.0, 0.6], [1.0, -3.0, 1.5]])     # sin = 0.565 (in), 0.997 (out)
    assert np.isfinite(pc.log_prob(pts)[0]) and np.isinf(pc.log_prob(pts)[1])


def test_prior_transform_maps_the_unit_cube_onto_the_prior():
    """prior_transform (what dynesty asks for): uniform numbers in -> samples of the prior out, numpy == torch,
    1-D points and batches; checked through each marginal's CDF."""
    pr = BoxPrior(LIMITS, {'sin_incl_limits_constraints': {}})
    rng = np.random.default_rng(3)
    u = rng.random((20000, 3))
    x = pr.prior_transform(u)
    xt = pr.prior_transform(torch.from_numpy(u)).numpy()
    assert np.allclose(x, xt, rtol=1e-12, atol=1e-12)
    assert np.array_equal(pr.prior_transform(u[7]), x[7]) and u[7, 0] != x[7, 0]           # input left untouched
    assert np.isfinite(pr.log_prob(x)).all()                                                # inside the support
    # Uniform(0.5, 2.5), Gaussian(-3, 0.7), Sine(0.2, 2.9): empirical CDFs at a few points
    assert abs((x[:, 0] < 1.0).mean() - 0.25) < 0.01
    assert abs((x[:, 1] < -3.0).mean() - 0.5) < 0.01 and abs(x[:, 1].std() - 0.7) < 0.02
    cdf_sine = lambda t: (np.cos(0.2) - np.cos(t)) / (np.cos(0.2) - np.cos(2.9))
    for t in (0.8, 1.5, 2.2):
        assert abs((x[:, 2] < t).mean() - cdf_sine(t)) < 0.01
    # monotone in every coordinate (a valid inverse CDF)
    g = np.linspace(0.001, 0.999, 50)
    grid = pr.prior_transform(np.stack([g, g, g], axis=1))
    assert (np.diff(grid, axis=0) > 0).all()


def test_batch_pool_is_one_batched_call():
    calls = []

    class Like(object):
        def log_likelihood_batch(self, theta, keys=None):
            calls.append(theta.shape)
            return theta.sum(axis=1)

    pool = BatchPool(Like(), ['a', 'b'])
    pts = [np.array([1.0, 2.0]), np.array([3.0, 4.0]), np.array([5.0, 6.0])]
    assert pool.map(pool.loglike, pts) == [3.0, 7.0, 11.0] and calls == [(3, 2)]
    assert pool.map(pool.loglike, []) == []
    assert pool.loglike(np.array([1.0, 2.0])) == 3.0             # a single point outside map()

    # a sampler's own wrapper around the likelihood (dynesty keeps it in .loglikelihood) is recognised
    class Wrapped(object):
        def __init__(self, f):
            self.loglikelihood = f

        def __call__(self, x):
            return self.loglikelihood(x)

    del calls[:]
    assert pool.map(Wrapped(pool.loglike), pts) == [3.0, 7.0, 11.0] and calls == [(3, 2)]
    user_fn = pool.register(lambda x: -1.0)                     # explicitly registered callable
    assert pool.map(user_fn, pts) == [3.0, 7.0, 11.0]
    # anything else a sampler maps (prior transforms, proposal / evolution functions with tuple
    # arguments) is applied point by point and never reaches the likelihood
    del calls[:]
    assert pool.map(lambda u: 2.0 * u, [np.array([0.5, 1.0])])[0].tolist() == [1.0, 2.0]
    assert pool.map(lambda args: args[0] + args[1]['k'], [(1, {'k': 2}), (3, {'k': 4})]) == [3, 7]
    assert calls == []


def test_stretch_move_recovers_a_gaussian():
    mean = torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64)
    L = torch.tensor([[1.0, 0, 0], [0.6, 0.5, 0], [-0.3, 0.2, 2.0]], dtype=torch.float64)
    cov = L @ L.T
    prec = torch.linalg.inv(cov)
    log_prob = lambda x: -0.5 * torch.einsum('ni,ij,nj->n', x - mean, prec, x - mean)
    s = EnsembleSampler(256, 3, log_prob, seed=3)
    p0 = mean + 0.1 * torch.randn(256, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    s.run(p0, 600)
    flat = s.chain[200:].reshape(-1, 3)
    assert 0.2 < s.acceptance_fraction < 0.9
    assert (flat.mean(0) - mean).abs().max() < 0.08
    assert (torch.cov(flat.T) - cov).abs().max() < 0.25
    assert s.ncalls == 1 + 2 * 600                  # one batched call per half-step


@pytest.mark.gpu
def test_config1_posterior_recovers_truth_on_gpu():
    case = synthetic.config1()
    like = case.likelihood()
    limits = {k: [case.box[k][0], case.box[k][1], 'Uniform'] for k in case.names}
    prior = BoxPrior(limits)
    fn = log_prob_fn(like, prior)
    nw = 512
    s = EnsembleSampler(nw, case.n_params, fn, device='cuda:0', seed=1)
    p0 = torch.from_numpy(case.sample_theta(nw, seed=5, scale=0.05)).cuda()
    s.run(p0, 400)
    flat = s.chain[150:].reshape(-1, case.n_params).cpu().numpy()
    truth = case.truth_vector()
    z = (flat.mean(0) - truth) / flat.std(0)
    assert np.abs(z).max() < 4.0, z                  # truth within a few posterior sigma
    assert 0.1 < s.acceptance_fraction < 0.9
    # the chain's log-probabilities are what the oracle says they are
    from oracle import sterne_oracle as O
    x, lp = s.last
    want = O.batch_log_likelihood(case.refepoch, case.LoD_timing, case.LoD_VLBI, case.shares, case.earth_xyz,
                                  x.cpu().numpy()) + prior.log_prob(x.cpu().numpy())
    assert np.abs(lp.cpu().numpy() - want).max() < 1e-7
    like.close()


@pytest.mark.gpu
def test_config2_4096_walkers_step_on_gpu():
    """BASELINE config 2: EFAC, 6 parameters, a vectorised batch of 4096 walkers."""
    case = synthetic.config2()
    like = case.likelihood()
    limits = {k: [case.box[k][0], case.box[k][1], 'Uniform'] for k in case.names}
    keys = list(reversed(case.names))                # prior (.inits) order differs from shares order
    prior = BoxPrior(limits, keys=keys)
    fn = log_prob_fn(like, prior)
    s = EnsembleSampler(4096, case.n_params, fn, device='cuda:0', seed=2)
    p0 = torch.from_numpy(case.sample_theta(4096, seed=6, scale=0.2)[:, ::-1].copy()).cuda()
    s.run(p0, 50)
    assert 0.05 < s.acceptance_fraction < 0.95
    assert s.ncalls == 101
    like.close()


@pytest.mark.gpu
def test_fused_log_posterior_equals_prior_plus_likelihood():
    """stn_logpost (one launch) == BoxPrior.log_prob + log_likelihood_batch, -inf mask included."""
    case = synthetic.config3('dots')
    like = case.likelihood()
    keys = list(reversed(case.names))
    limits = {}
    for k in case.names:
        lo, hi = case.box[k]
        if k.startswith('incl'):
            limits[k] = [0.1, 3.0, 'Sine']
        elif k.startswith('mu_a'):
            limits[k] = [case.truth[k], 1.5, 'Gaussian']
        else:
            limits[k] = [lo, hi, 'Uniform']
    prior = BoxPrior(limits, {'sin_incl_limits_constraints': {'incl_0': [0.3, 0.999]}}, keys=keys)
    theta = case.sample_theta(2048, seed=12, scale=1.15)[:, ::-1].copy()      # ~25 % fall outside the box
    th = torch.from_numpy(theta).cuda()
    fused = like.log_posterior_batch(th, prior).cpu().numpy()
    lp = prior.log_prob(theta)
    ll = like.log_likelihood_batch(theta, keys=keys)
    want = np.where(np.isfinite(lp), lp + ll, -np.inf)
    assert np.array_equal(np.isfinite(fused), np.isfinite(want))
    assert 0.05 < np.isinf(fused).mean() < 0.95
    m = np.isfinite(want)
    assert (np.abs(fused[m] - want[m]) / np.maximum(np.abs(want[m]), 400.0)).max() < 1e-13
    # the built-in sampler driven by the fused entry point
    s = EnsembleSampler(256, case.n_params, lambda x: like.log_posterior_batch(x, prior), device='cuda:0', seed=4)
    half = np.array([0.5 * (case.box[k][1] - case.box[k][0]) for k in case.names])
    p0 = case.truth_vector() + 0.02 * half * np.random.default_rng(13).standard_normal((256, case.n_params))
    s.run(torch.from_numpy(p0[:, ::-1].copy()).cuda(), 30)
    assert 0.02 < s.acceptance_fraction < 0.98
    with pytest.raises(TypeError):
        like.log_posterior_batch(theta, prior)
    like.close()


@pytest.mark.gpu
def test_fused_sampler_matches_torch_sampler_statistically():
    """stn_ensemble_run: same posterior as the torch-op sampler and the truth inside it; reproducible
    for a given seed; the persistent-kernel engine and the three-launches engine give the SAME chain."""
    from sterne_b200.sampler import FusedEnsembleSampler
    case = synthetic.config2()
    like = case.likelihood()
    limits = {k: [case.box[k][0], case.box[k][1], 'Uniform'] for k in case.names}
    prior = BoxPrior(limits)
    nw = 1024
    half = np.array([0.5 * (case.box[k][1] - case.box[k][0]) for k in case.names])
    p0 = case.truth_vector() + 0.02 * half * np.random.default_rng(3).standard_normal((nw, case.n_params))
    p0[:, case.names.index('efac_0')] = np.abs(p0[:, case.names.index('efac_0')]) + 0.5
    f = FusedEnsembleSampler(like, prior, nw, seed=11)
    f.run(p0, 600, thin=2)
    flat_f = f.chain[100:].reshape(-1, case.n_params).cpu().numpy()
    t = EnsembleSampler(nw, case.n_params, lambda x: like.log_posterior_batch(x, prior), device='cuda:0', seed=12)
    t.run(torch.from_numpy(p0).cuda(), 600)
    flat_t = t.chain[200:].reshape(-1, case.n_params).cpu().numpy()
    assert 0.1 < f.acceptance_fraction < 0.9 and abs(f.acceptance_fraction - t.acceptance_fraction) < 0.05
    sd = flat_t.std(0)
    assert (np.abs(flat_f.mean(0) - flat_t.mean(0)) / sd).max() < 0.1, (flat_f.mean(0), flat_t.mean(0))
    assert (np.abs(flat_f.std(0) / sd - 1.0)).max() < 0.1
    astrometric = [i for i, k in enumerate(case.names) if not k.startswith('efac')]
    assert (np.abs(flat_f.mean(0) - case.truth_vector()) / flat_f.std(0))[astrometric].max() < 4.0
    # log-probabilities carried by the sampler are the posterior's, bit for bit (plan lanes = 1, splits = 1)
    from sterne_b200._lib import make_plan
    x, lp = f.last
    assert torch.equal(like.log_posterior_batch(x, prior, plan=make_plan(1, 1)), lp)
    # reproducible, and independent of the engine: one persistent cooperative kernel == three launches per half-step
    g = FusedEnsembleSampler(like, prior, nw, seed=11, engine='persistent')
    g.run(p0, 50, thin=5)
    h = FusedEnsembleSampler(like, prior, nw, seed=11, engine='launches', plan=make_plan(1, 1))
    h.run(p0, 50, thin=5)
    assert torch.equal(g.last[0], h.last[0]) and torch.equal(g.last[1], h.last[1])
    assert torch.equal(g.chain, h.chain) and torch.equal(g.lnprob, h.lnprob)
    assert g.acceptance_fraction == h.acceptance_fraction
    # continuing a run == one longer run (the step counter carries the random stream)
    k = FusedEnsembleSampler(like, prior, nw, seed=11)
    k.run(p0, 20, store=False)
    k.run(k.last[0], 30, store=False)
    assert torch.equal(k.last[0], g.last[0])
    # odd walker counts per warp / block and a binary problem with priors of every kind
    case3 = synthetic.config3('dots')
    like3 = case3.likelihood()
    lim3 = {}
    for kname in case3.names:
        lo, hi = case3.box[kname]
        lim3[kname] = [0.1, 3.0, 'Sine'] if kname.startswith('incl') else \
            [case3.truth[kname], 1.5, 'Gaussian'] if kname.startswith('mu_a') else [lo, hi, 'Uniform']
    prior3 = BoxPrior(lim3, {'sin_incl_limits_constraints': {'incl_0': [0.3, 0.999]}})
    nw3 = 166
    half3 = np.array([0.5 * (case3.box[kname][1] - case3.box[kname][0]) for kname in case3.names])
    p3 = case3.truth_vector() + 0.02 * half3 * np.random.default_rng(5).standard_normal((nw3, case3.n_params))
    a3 = FusedEnsembleSampler(like3, prior3, nw3, seed=3, engine='persistent')
    a3.run(p3, 40, thin=1)
    b3 = FusedEnsembleSampler(like3, prior3, nw3, seed=3, engine='launches', plan=make_plan(1, 1))
    b3.run(p3, 40, thin=1)
    assert torch.equal(a3.chain, b3.chain) and torch.equal(a3.lnprob, b3.lnprob)
    assert 0.02 < a3.acceptance_fraction < 0.98
    like3.close()
    # joint fits: the persistent kernel spreads the SOURCES of a walker over lanes of a warp (5 sources -> 8 lanes,
    # 3 -> 4, 2 -> 2) and still sums them in source order: same chain as the launch engine
    for case4, nw4 in ((synthetic.config4(), 250), (synthetic.config5('C', n_sources=3, n_epochs=40), 130),
                       (synthetic.config5('B', n_sources=2, n_epochs=30), 64)):
        like4 = case4.likelihood()
        prior4 = BoxPrior({kname: [case4.box[kname][0], case4.box[kname][1], 'Uniform'] for kname in case4.names})
        half4 = np.array([0.5 * (case4.box[kname][1] - case4.box[kname][0]) for kname in case4.names])
        p4 = case4.truth_vector() + 0.02 * half4 * np.random.default_rng(6).standard_normal((nw4, case4.n_params))
        a4 = FusedEnsembleSampler(like4, prior4, nw4, seed=8, engine='persistent')
        a4.run(p4, 30, thin=3)
        b4 = FusedEnsembleSampler(like4, prior4, nw4, seed=8, engine='launches', plan=make_plan(1, 1))
        b4.run(p4, 30, thin=3)
        assert torch.equal(a4.chain, b4.chain) and torch.equal(a4.lnprob, b4.lnprob), case4.name
        assert a4.acceptance_fraction == b4.acceptance_fraction
        like4.close()
    like.close()


@pytest.mark.gpu
def test_ensemble_shared_by_several_devices_gives_the_same_chain():
    """stn_multi_ensemble_run: the walkers of one ensemble are shared among several contexts / GPUs (every
    device proposes and evaluates its slice of a half-step, its accept kernel writes into all copies);
    the chain is bit-identical to the one-GPU run.  [0, 0] exercises it on a single-GPU box."""
    from sterne_b200.sampler import FusedEnsembleSampler
    case = synthetic.config5('C', n_epochs=200)
    limits = {k: [case.box[k][0], case.box[k][1], 'Uniform'] for k in case.names}
    prior = BoxPrior(limits)
    nw = 1000
    half = np.array([0.5 * (case.box[k][1] - case.box[k][0]) for k in case.names])
    p0 = case.truth_vector() + 0.02 * half * np.random.default_rng(3).standard_normal((nw, case.n_params))
    one = case.likelihood(device=0)
    ref = FusedEnsembleSampler(one, prior, nw, seed=5, engine='launches')
    ref.run(p0, 12, thin=3)
    lists = [[0, 0], [0, 0, 0]]
    if torch.cuda.device_count() >= 2:
        lists.append(list(range(min(torch.cuda.device_count(), 8))))
    for devices in lists:
        like = case.likelihood(device=devices)
        s = FusedEnsembleSampler(like, prior, nw, seed=5)
        s.run(p0, 12, thin=3)
        assert torch.equal(s.chain, ref.chain) and torch.equal(s.lnprob, ref.lnprob), devices
        assert torch.equal(s.last[0], ref.last[0]) and torch.equal(s.last[1], ref.last[1])
        assert s.acceptance_fraction == ref.acceptance_fraction
        s.run(s.last[0], 5, store=False)                       # continuing works (step counter, fresh copies)
        like.close()
    ref.run(ref.last[0], 5, store=False)
    assert torch.equal(s.last[0], ref.last[0])
    one.close()
