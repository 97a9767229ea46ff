#!/usr/bin/env python
"""Kernel-only timing of config 5 variants (tuning aid; bench.py is the contract bench).
usage: [STERNE_B200_LIB=path] python tools/kbench.py [--log2n 22] [--steps 5] [--variants ABC]"""
import argparse, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sterne_b200 import synthetic, fp64_peak

ap = argparse.ArgumentParser()
ap.add_argument('--log2n', type=float, default=22)
ap.add_argument('--steps', type=int, default=5)
ap.add_argument('--variants', default='ABC')
ap.add_argument('--plan', type=int, default=0)
ap.add_argument('--check', action='store_true')
a = ap.parse_args()
dev = torch.device('cuda', 0)
N = int(round(2 ** a.log2n))
flops = {'A': 30.0, 'B': 42.0, 'C': 43.5}
res = {'lib': os.environ.get('STERNE_B200_LIB', 'default'), 'log2n': a.log2n}
for v in a.variants:
    case = synthetic.config5(v)
    like = case.likelihood()
    th = synthetic.device_theta(case, N, 1234, dev)
    out = torch.empty(N, dtype=torch.float64, device=dev)
    for _ in range(3):
        like.log_likelihood_batch(th, out=out, plan=a.plan)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        like.log_likelihood_batch(th, out=out, plan=a.plan)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    ev = N * case.evals_per_sample / (ms * 1e-3)
    res[v] = {'ms': round(ms, 3), 'evals_per_s': ev, 'tflops': ev * flops[v] / 1e12}
    if a.check:
        n = min(N, 4096)
        strict = like.log_likelihood_batch(th[:n].contiguous(), mode='strict')
        rel = (out[:n] - strict).abs() / strict.abs()
        k = int(rel.argmax())
        res[v]['max_rel_vs_strict'] = rel.max().item()
        res[v]['worst'] = {'logl': strict[k].item(), 'abs_err': (out[k] - strict[k]).abs().item()}
        res[v]['max_abs_err'] = (out[:n] - strict).abs().max().item()
        res[v]['logl_range'] = [strict.min().item(), strict.max().item()]
    like.close()
pk, ms = fp64_peak(0, iters=1 << 17, reps=5)
res['dfma_peak_tflops'] = pk
print(json.dumps(res))
