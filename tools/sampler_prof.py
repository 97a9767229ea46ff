#!/usr/bin/env python
"""One short run of the persistent sampler kernel (for an ncu capture)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sterne_b200 import synthetic
from sterne_b200.sampler import BoxPrior, FusedEnsembleSampler
case = synthetic.config2(); like = case.likelihood()
prior = BoxPrior({k: [case.box[k][0], case.box[k][1], 'Uniform'] for k in case.names})
half = np.array([0.5 * (case.box[k][1] - case.box[k][0]) for k in case.names])
p0 = case.truth_vector() + 0.02 * half * np.random.default_rng(1).standard_normal((4096, case.n_params))
f = FusedEnsembleSampler(like, prior, 4096, seed=1, engine='persistent')
f.run(p0, 300, store=False)
torch.cuda.synchronize()
print('acceptance', f.acceptance_fraction)
