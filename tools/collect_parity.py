#!/usr/bin/env python
"""Assembles the per-test parity records the GPU tests leave under gpurun_out/parity/ (tests/conftest.py,
record_parity) into one committed file.  usage: python tools/collect_parity.py profiles/r02_parity.json"""
import glob, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r02_parity.json')
recs = {}
for f in sorted(glob.glob(os.path.join(ROOT, 'gpurun_out', 'parity', '*.json'))):
    d = json.load(open(f))
    recs[d['tag']] = d['stats']
logl = {k: v for k, v in recs.items() if not k.startswith('offsets_mas/')}
summary = {
    'contract': '1e-10 relative on log L (pure |got - want| / |want|, no floor), 1e-12 mas on model offsets',
    'checker': 'oracle/sterne_oracle.py (float64, pinned to the unmodified reference); exceedances arbitrated by '
               'oracle/mp_arbiter.py (50-digit evaluation of the same formula on the same float64 inputs)',
    'logl_cases': len(logl),
    'logl_vectors': sum(v['n'] for v in logl.values()),
    'logl_vectors_above_1e-10': sum(v['count_above'] for v in logl.values()),
    'logl_all_exceedances_explained': all(v['explained'] for v in logl.values()),
    'logl_worst_pure_relative': max((v['max_rel'] for v in logl.values()), default=0.0),
    'offsets_worst_abs_mas': max((max(v['max_abs_mas'].values()) for k, v in recs.items() if k.startswith('offsets_mas/')), default=0.0),
}
json.dump({'summary': summary, 'records': recs}, open(dst, 'w'), indent=1)
print(json.dumps(summary, indent=1))
