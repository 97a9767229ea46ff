import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden', 'golden_reference.json')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu under gpurun)')
    config.addinivalue_line('filterwarnings', 'ignore::PendingDeprecationWarning')


@pytest.fixture(scope='session')
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


def materialise(case, directory):
    """Writes the fixture's input files; returns role -> paths (as simulate() would get them)."""
    os.makedirs(directory, exist_ok=True)
    for name, text in case['files'].items():
        with open(os.path.join(directory, name), 'w') as f:
            f.write(text)
    roles = case['file_roles']
    j = lambda n: os.path.join(directory, n) if n else ''
    return {'pmparins': [j(n) for n in roles['pmparins']],
            'preliminaries': None if roles['preliminaries'] is None else [j(n) for n in roles['preliminaries']],
            'parfiles': [j(n) for n in roles['parfiles']],
            'inits': j(roles['inits'])}


@pytest.fixture(scope='session')
def golden_dir(tmp_path_factory, golden):
    base = tmp_path_factory.mktemp('golden_inputs')
    return {c['name']: materialise(c, str(base / c['name'])) for c in golden['cases']}


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
