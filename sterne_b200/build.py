"""Builds libsterne_b200.so (the C-ABI library, include/sterne_b200.h) in-tree with nvcc for sm_100a."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, 'csrc', 'sterne_b200.cu')
LIB = os.path.join(HERE, 'libsterne_b200.so')

NVCC_FLAGS = [
    '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
    '-Xcompiler', '-fPIC', '-shared', '-I', os.path.join(ROOT, 'include'),
]


def nvcc_path():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found; sterne_b200 has no CPU path and cannot be built without it')


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    csrc = os.path.dirname(SRC)
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [os.path.join(ROOT, 'include', 'sterne_b200.h'),
                                                                os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False, extra=(), out=None):
    """out: alternative output path (tuning experiments); default is the in-tree library."""
    if out is None and not force and not needs_build():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + list(extra) + ['-o', out or LIB, SRC]
    if verbose:
        print(' '.join(cmd))
    subprocess.check_call(cmd)
    return out or LIB


DEBUG_LIB = os.path.join(ROOT, 'build_variants', 'libsterne_b200_debug.so')


def build_debug_library(verbose=False):
    """The bounds-asserting build (-DSTN_DEBUG, see csrc/sterne_b200.cu): select it with
    STERNE_B200_LIB=build_variants/libsterne_b200_debug.so."""
    os.makedirs(os.path.dirname(DEBUG_LIB), exist_ok=True)
    return build_library(force=True, verbose=verbose, extra=['-DSTN_DEBUG', '-lcuda'], out=DEBUG_LIB)


if __name__ == '__main__':
    if '--debug' in sys.argv:
        print(build_debug_library(verbose=True))
    else:
        build_library(force='--force' in sys.argv, verbose=True,
                      extra=['-Xptxas', '-v'] if '--ptxas-v' in sys.argv else [])
        print(LIB)
