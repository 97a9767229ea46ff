#!/usr/bin/env python
"""Host-path (stn_loglike_host) timing for one batch size under the current chunk schedule
(STN_HOST_FIRST_LOG2 / STN_HOST_GROWTH env).  usage: python tools/e2e_tune.py --log2n 19 [--plan 257]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sterne_b200 import synthetic
ap = argparse.ArgumentParser()
ap.add_argument('--log2n', type=int, default=19)
ap.add_argument('--plan', type=int, default=257)
ap.add_argument('--steps', type=int, default=20)
a = ap.parse_args()
case = synthetic.config5('C'); like = case.likelihood()
n = 1 << a.log2n
dev = torch.device('cuda', 0)
theta = synthetic.device_theta(case, n, 1234, dev)
th = torch.empty((n, case.n_params), dtype=torch.float64, pin_memory=True); th.copy_(theta)
out = torch.empty(n, dtype=torch.float64, pin_memory=True)
thn, outn = th.numpy(), out.numpy()
o = torch.empty(n, dtype=torch.float64, device=dev)
for _ in range(3):
    like.log_likelihood_batch(theta, out=o, plan=a.plan)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    like.log_likelihood_batch(theta, out=o, plan=a.plan)
e1.record(); torch.cuda.synchronize()
kms = e0.elapsed_time(e1) / a.steps
for _ in range(3):
    like.log_likelihood_batch(thn, out=outn, plan=a.plan)
t0 = time.perf_counter()
for _ in range(a.steps):
    like.log_likelihood_batch(thn, out=outn, plan=a.plan)
ems = (time.perf_counter() - t0) * 1e3 / a.steps
print(json.dumps({'log2n': a.log2n, 'first_log2': os.environ.get('STN_HOST_FIRST_LOG2'), 'growth': os.environ.get('STN_HOST_GROWTH'),
                  'kernel_ms': round(kms, 3), 'e2e_ms': round(ems, 3), 'ratio': round(ems / kms, 4),
                  'same': bool(np.array_equal(outn, o.cpu().numpy()))}))
