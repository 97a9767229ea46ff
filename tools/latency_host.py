import os, sys, time, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from sterne_b200 import synthetic
res = {}
for name, mk in (('config2', synthetic.config2), ('config4', synthetic.config4)):
    case = mk(); like = case.likelihood()
    for n in (1024, 4096, 16384):
        th = case.sample_theta(n, seed=2)
        for _ in range(20): like.log_likelihood_batch(th)
        t0 = time.perf_counter()
        for _ in range(500): like.log_likelihood_batch(th)
        res['%s/%d' % (name, n)] = round(1e6 * (time.perf_counter() - t0) / 500, 1)
    like.close()
print(os.environ.get('STN_SMALL_ROWS'), res)
