"""Randomised parity sweep (tools/fuzz_parity.py): random problem shapes, sharing patterns, switched-off
parameters, value sentinels, constraints, plans and layouts; GPU fast and strict vs the C oracle."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.gpu
def test_random_problems_match_the_c_oracle():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'fuzz_parity.py'), '40', '7'],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert 'fuzz ok: 40 problems' in out.stdout
