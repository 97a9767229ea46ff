"""bench.py's reference arm runs on the host cores only, so its JSON contract can be checked without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '2',
                          '--warmup', '1', '--cpu-rows', '256'], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip().startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['higher_is_better'] is True and d['unit'] == 'evals/s'
    assert d['metric'].startswith('log-L evals/sec') and d['dtype'] == 'f64' and d['data'] == 'synthetic'
    assert d['steps'] == 2 and d['n_gpus'] == 1 and d['vs_baseline'] is None and d['gpu_launches'] == 0
    assert d['value'] > 0 and d['ms_per_step'] > 0
    assert 'workload' in d['config'] and 'model' not in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and '256 parameter vectors' in cb['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 'evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', LOCAL_RANK='1', WORLD_SIZE='2')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1'],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ''


def test_ours_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip('GPU present')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--no-cpu-baseline', '--steps', '1'],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and 'no CPU path' in (out.stderr + out.stdout)
