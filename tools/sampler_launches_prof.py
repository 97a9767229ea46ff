import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from sterne_b200 import synthetic
from sterne_b200.sampler import BoxPrior, FusedEnsembleSampler
case = synthetic.config5('C')
prior = BoxPrior({k: [case.box[k][0], case.box[k][1], 'Uniform'] for k in case.names})
half = np.array([0.5 * (case.box[k][1] - case.box[k][0]) for k in case.names])
W = 1 << 19
p0 = case.truth_vector() + 0.02 * half * np.random.default_rng(1).standard_normal((W, case.n_params))
like = case.likelihood(device=0)
s = FusedEnsembleSampler(like, prior, W, seed=1, engine='launches')
s.run(p0, 3, store=False)
torch.cuda.synchronize()
print('plan', s.plan)
