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


def logl_scale(list_of_dict_VLBI):
    """Magnitude of the log-determinant part of log-L: sum over sources of |sum(log errs)|.

    log L = -sum(log sigma) - chi^2/2 is a difference of two large positive numbers and
    crosses zero inside any prior box, where a purely relative error is meaningless.  The
    1e-10 tolerance is therefore applied relative to max(|log L|, this scale), the size of
    the terms that were summed (the reference's own rounding error scales the same way)."""
    return float(sum(abs(np.sum(np.log(np.asarray(v['errs'])))) for v in list_of_dict_VLBI))


def logl_err(got, want, scale):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    return np.abs(got - want) / np.maximum(np.abs(want), scale)


PARITY_DIR = os.path.join(ROOT, 'gpurun_out', 'parity')


def record_parity(tag, stats):
    """Parity statistics measured by the GPU tests, one JSON file per tag under gpurun_out/parity/
    (scratch; tools/collect_parity.py assembles them into profiles/rNN_parity.json)."""
    try:
        os.makedirs(PARITY_DIR, exist_ok=True)
        with open(os.path.join(PARITY_DIR, tag.replace('/', '__') + '.json'), 'w') as f:
            json.dump({'tag': tag, 'stats': stats}, f)
    except OSError:
        pass
