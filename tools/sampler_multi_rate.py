#!/usr/bin/env python
"""Ensemble steps/s when ONE process shares a large ensemble among 1, 2, ... GPUs (stn_multi_ensemble_run)
on the throughput problem (config 5 C: 8 sources x 1024 epochs per walker)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sterne_b200 import synthetic
from sterne_b200.sampler import BoxPrior, FusedEnsembleSampler
case = synthetic.config5('C')
prior = BoxPrior({k: [case.box[k][0], case.box[k][1], 'Uniform'] for k in case.names})
half = np.array([0.5 * (case.box[k][1] - case.box[k][0]) for k in case.names])
res = {}
ng = torch.cuda.device_count()
for W in (1 << 17, 1 << 19):
    p0 = case.truth_vector() + 0.02 * half * np.random.default_rng(1).standard_normal((W, case.n_params))
    base = None
    for G in [g for g in (1, 2, 4, 8) if g <= ng]:
        like = case.likelihood(device=list(range(G)) if G > 1 else 0)
        s = FusedEnsembleSampler(like, prior, W, seed=1, engine='launches')
        s.run(p0, 2, store=False)
        for d in range(G):
            torch.cuda.synchronize(d)
        n = 40
        t0 = time.perf_counter()
        s.run(s.last[0], n, store=False)
        for d in range(G):
            torch.cuda.synchronize(d)
        dt = time.perf_counter() - t0
        last = s.last[0].cpu().numpy()
        if base is None:
            base = last
        res['W%d/G%d' % (W, G)] = {'steps_per_s': n / dt, 'evals_per_s': n * W * case.evals_per_sample / dt,
                                   'same_walkers_as_1gpu': bool(np.array_equal(last, base))}
        like.close()
print(json.dumps(res))
