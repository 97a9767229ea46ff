#!/usr/bin/env python
"""Repeats the paths that involve host threads, events and barriers many times (flakiness / hang check)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sterne_b200 import synthetic
from sterne_b200.sampler import BoxPrior, FusedEnsembleSampler
ng = torch.cuda.device_count()
devs = list(range(ng)) if ng > 1 else [0, 0, 0]
res = {}
# 1. one process, several contexts, odd batch sizes, pageable and pinned, many repetitions
case = synthetic.config5('C', n_epochs=128)
single = case.likelihood(device=0)
multi = case.likelihood(device=devs)
rng = np.random.default_rng(0)
t0 = time.time(); bad = 0
for it in range(40):
    n = int(rng.integers(1, 400000))
    th = case.sample_theta(n, seed=it, scale=0.3)
    a = single.log_likelihood_batch(th)
    b = multi.log_likelihood_batch(th)
    bad += int(not np.array_equal(a, b))
res['multi_host_mismatches'] = bad; res['multi_host_s'] = time.time() - t0
# 2. the persistent sampler: long run, repeated launches
c2 = synthetic.config2(); l2 = c2.likelihood(device=0)
prior = BoxPrior({k: [c2.box[k][0], c2.box[k][1], 'Uniform'] for k in c2.names})
half = np.array([0.5 * (c2.box[k][1] - c2.box[k][0]) for k in c2.names])
p0 = c2.truth_vector() + 0.02 * half * np.random.default_rng(1).standard_normal((4096, c2.n_params))
s = FusedEnsembleSampler(l2, prior, 4096, seed=3)
t0 = time.time()
x = p0
for rep in range(20):
    x, lp = s.run(x, 5000, store=False)
torch.cuda.synchronize()
res['persistent_100k_steps_s'] = time.time() - t0; res['acc'] = s.acceptance_fraction
# 3. the multi-device sampler, many short runs
prior5 = BoxPrior({k: [case.box[k][0], case.box[k][1], 'Uniform'] for k in case.names})
h5 = np.array([0.5 * (case.box[k][1] - case.box[k][0]) for k in case.names])
p5 = case.truth_vector() + 0.02 * h5 * np.random.default_rng(2).standard_normal((3000, case.n_params))
ref = FusedEnsembleSampler(single, prior5, 3000, seed=9, engine='launches'); ref.run(p5, 60, store=False)
t0 = time.time(); bad = 0
for rep in range(5):
    m = FusedEnsembleSampler(multi, prior5, 3000, seed=9); m.run(p5, 60, store=False)
    bad += int(not torch.equal(m.last[0], ref.last[0]))
res['multi_sampler_mismatches'] = bad; res['multi_sampler_s'] = time.time() - t0
print(json.dumps(res))
