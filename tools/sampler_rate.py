import sys, os, time, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from sterne_b200 import synthetic
from sterne_b200.sampler import BoxPrior, FusedEnsembleSampler
res = {}
for name, mk in (('config1', synthetic.config1), ('config2', synthetic.config2), ('config4', synthetic.config4)):
    case = mk(); like = case.likelihood()
    limits = {k: [case.box[k][0], case.box[k][1], 'Uniform'] for k in case.names}
    prior = BoxPrior(limits)
    half = np.array([0.5 * (case.box[k][1] - case.box[k][0]) for k in case.names])
    for W in (1024, 4096, 8192, 16384):
        p0 = case.truth_vector() + 0.02 * half * np.random.default_rng(1).standard_normal((W, case.n_params))
        f = FusedEnsembleSampler(like, prior, W, seed=1, engine='persistent')
        f.run(p0, 50, store=False)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        f.run(f.last[0], 4000, store=False)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        res['%s/W%d' % (name, W)] = round(4000 / dt)
    like.close()
print(json.dumps({'lib': os.environ.get('STERNE_B200_LIB', 'default'), 'steps_per_s': res}))
