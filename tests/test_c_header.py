"""The drop-in boundary is a C ABI: include/sterne_b200.h must compile as plain C99 and a C program must be
able to link and call the library without any Python, torch or CUDA headers."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

C_PROGRAM = r'''
#include <stdio.h>
#include <string.h>
#include "sterne_b200.h"

int main(void) {
    int64_t lo = -1, hi = -1;
    if (stn_version() != STN_ABI_VERSION) return 10;
    if (stn_multi_slice(10, 3, 1, &lo, &hi) != STN_OK || lo != 4 || hi != 7) return 11;
    if (stn_multi_slice(10, 0, 0, &lo, &hi) != STN_ERR_INVALID) return 12;
    if (strstr(stn_last_error(), "slice") == NULL) return 13;
    /* a one-source, one-epoch problem: without an sm_100 GPU stn_create must say so (no CPU path) */
    double fast[STN_FAST_ROW] = {0}, strict[STN_STRICT_ROW] = {0};
    fast[6] = fast[7] = 1.0; strict[6] = strict[7] = 1.0;
    StnSource src;
    memset(&src, 0, sizeof src);
    src.n_epochs = 1; src.fast_rows = fast; src.strict_rows = strict;
    for (int r = 0; r < STN_NROOTS; ++r) src.col[r] = -1;
    src.col[STN_DEC] = 0; src.col[STN_RA] = 1; src.cos_decj = 1.0;
    StnProblem pb;
    memset(&pb, 0, sizeof pb);
    pb.n_sources = 1; pb.n_params = 2; pb.sources = &src;
    StnCtx* ctx = NULL;
    int rc = stn_create(&pb, 0, &ctx);
    printf("create rc=%d (%s)\n", rc, rc ? stn_last_error() : "ok");
    if (rc == STN_OK) {
        double theta[2] = {0.3, 1.0}, out = 0.0;
        if (stn_loglike_host(ctx, theta, 1, 2, STN_LAYOUT_AOS, STN_MODE_STRICT, 0, &out) != STN_OK) return 14;
        printf("logl=%.17g\n", out);
        stn_destroy(ctx);
        return 0;
    }
    return rc == STN_ERR_DEVICE ? 0 : 15;
}
'''


def test_header_is_c99_and_library_links_from_c(tmp_path):
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('no gcc')
    from sterne_b200 import _lib
    _lib.load()                                                     # builds the library if needed
    src = tmp_path / 'demo.c'
    src.write_text(C_PROGRAM)
    exe = tmp_path / 'demo'
    libdir = os.path.join(ROOT, 'sterne_b200')
    subprocess.check_call([gcc, '-std=c99', '-Wall', '-Wextra', '-Werror', '-pedantic', '-I', os.path.join(ROOT, 'include'),
                           str(src), '-o', str(exe), '-L', libdir, '-lsterne_b200', '-Wl,-rpath,' + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert 'create rc=' in out.stdout
    import torch
    if not torch.cuda.is_available():
        assert 'rc=-4' in out.stdout and 'no CPU path' in out.stdout
