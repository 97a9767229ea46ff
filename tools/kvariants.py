#!/usr/bin/env python
"""Runs tools/kbench.py once per library in build_variants/ (tuning aid: same source, different
build-time knobs) and collects the results.  usage: python tools/kvariants.py [--variants C] [--out file]"""
import argparse, glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument('--variants', default='C')
ap.add_argument('--steps', type=int, default=5)
ap.add_argument('--log2n', type=float, default=22)
ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'kvariants.json'))
ap.add_argument('--only', default='')
ap.add_argument('--plan', type=int, default=0)
a = ap.parse_args()
libs = sorted(glob.glob(os.path.join(ROOT, 'build_variants', 'lib_*.so')))
if a.only:
    libs = [l for l in libs if any(o in os.path.basename(l) for o in a.only.split(','))]
res = []
for lib in libs:
    env = dict(os.environ, STERNE_B200_LIB=lib)
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'kbench.py'), '--variants', a.variants, '--steps',
                        str(a.steps), '--log2n', str(a.log2n), '--plan', str(a.plan), '--check'], env=env, capture_output=True, text=True)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ''
    try:
        d = json.loads(line)
    except Exception:
        d = {'error': (r.stderr or r.stdout)[-800:]}
    d['lib'] = os.path.basename(lib)
    res.append(d)
    brief = {k: (v['ms'], '%.2e' % v.get('max_rel_vs_strict', -1)) for k, v in d.items() if isinstance(v, dict) and 'ms' in v}
    print(d['lib'], brief, d.get('error', ''), flush=True)
os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump(res, open(a.out, 'w'), indent=1)
