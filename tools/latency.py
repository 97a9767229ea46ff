#!/usr/bin/env python
"""Small-batch latency of the drop-in API (BASELINE configs 1-4 are latency-, not throughput-bound).
Prints one JSON object; the CPU column is the oracle's reference-shaped scalar path (checker only)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sterne_b200 import synthetic
from sterne_b200.sampler import BoxPrior, EnsembleSampler, FusedEnsembleSampler
from oracle import sterne_oracle as O

def timeit(fn, n):
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n

res = {}
for name, mk in (('config1', synthetic.config1), ('config2', synthetic.config2), ('config3', lambda: synthetic.config3('dots')),
                 ('config4', synthetic.config4)):
    case = mk()
    like = case.likelihood()
    th1 = case.sample_theta(1, seed=1)
    params = dict(zip(case.names, th1[0]))
    def one():
        like.parameters.update(params)
        return like.log_likelihood()
    th4k = case.sample_theta(4096, seed=2)
    d4k = torch.from_numpy(th4k).cuda()
    out = torch.empty(4096, dtype=torch.float64, device='cuda')
    r = {'evals_per_sample': case.evals_per_sample, 'P': case.n_params}
    r['dict_api_us_per_call'] = 1e6 * timeit(one, 2000)
    r['batch4096_host_us'] = 1e6 * timeit(lambda: like.log_likelihood_batch(th4k), 500)
    r['batch4096_device_us'] = 1e6 * timeit(lambda: like.log_likelihood_batch(d4k, out=out), 2000)
    ol = O.Gaussianlikelihood(case.refepoch, case.LoD_timing, case.LoD_VLBI, case.shares, case.earth_xyz,
                              case.DoD_additional_constraints, case.a1dot_constraints)
    t0 = time.perf_counter(); ol.log_likelihood_rows(case.sample_theta(50, seed=3)); dt = (time.perf_counter() - t0) / 50
    r['cpu_oracle_scalar_us_per_call'] = 1e6 * dt
    limits = {k: [case.box[k][0], case.box[k][1], 'Uniform'] for k in case.names}
    prior = BoxPrior(limits)
    s = EnsembleSampler(4096, case.n_params, lambda x: like.log_posterior_batch(x, prior), device='cuda:0', seed=1)
    half = np.array([0.5 * (case.box[k][1] - case.box[k][0]) for k in case.names])
    p0 = case.truth_vector() + 0.02 * half * np.random.default_rng(1).standard_normal((4096, case.n_params))
    s.run(torch.from_numpy(p0).cuda(), 20, store=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s.run(s.last[0], 200, store=False)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    r['sampler_4096_walkers_steps_per_s'] = 200 / dt
    r['sampler_loglike_evals_per_s'] = 200 * 4096 * case.evals_per_sample / dt
    for engine in ('persistent', 'launches'):
        f = FusedEnsembleSampler(like, prior, 4096, seed=1, engine=engine)
        f.run(p0, 50, store=False)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        f.run(f.last[0], 4000, store=False)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        r['fused_sampler_%s_4096_walkers_steps_per_s' % engine] = 4000 / dt
        r['fused_sampler_%s_loglike_evals_per_s' % engine] = 4000 * 4096 * case.evals_per_sample / dt
    res[name] = r
    like.close()
print(json.dumps(res))
