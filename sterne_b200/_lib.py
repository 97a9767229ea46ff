"""ctypes binding of libsterne_b200.so (include/sterne_b200.h).  There is no CPU path:
if the library cannot be loaded every entry point raises."""
import ctypes
import os

from . import build as _build

NROOTS = 9
FAST_ROW = 12
STRICT_ROW = 16
LAYOUT_AOS, LAYOUT_SOA = 0, 1
MODE_FAST, MODE_STRICT = 0, 1
MODES = {'fast': MODE_FAST, 'strict': MODE_STRICT, MODE_FAST: MODE_FAST, MODE_STRICT: MODE_STRICT}

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)


class StnSource(ctypes.Structure):
    _fields_ = [('n_epochs', ctypes.c_int64),
                ('fast_rows', c_double_p),
                ('strict_rows', c_double_p),
                ('col', ctypes.c_int32 * NROOTS),
                ('efac_on', ctypes.c_int32),
                ('binary', ctypes.c_int32),
                ('cos_decj', ctypes.c_double),
                ('neg_sum_log_err', ctypes.c_double),
                ('a1_lts', ctypes.c_double)]


class StnPrior(ctypes.Structure):
    _fields_ = [('n_params', ctypes.c_int32),
                ('kind', c_int32_p),
                ('a', c_double_p),
                ('b', c_double_p),
                ('n_sin_limits', ctypes.c_int32),
                ('sin_col', c_int32_p),
                ('sin_lo', c_double_p),
                ('sin_hi', c_double_p)]


PRIOR_KINDS = {'Uniform': 1, 'Gaussian': 2, 'Sine': 3}


class StnProblem(ctypes.Structure):
    _fields_ = [('n_sources', ctypes.c_int32),
                ('n_params', ctypes.c_int32),
                ('sources', ctypes.POINTER(StnSource)),
                ('n_a1dot', ctypes.c_int32),
                ('a1dot_source', c_int32_p),
                ('a1dot_mu', c_double_p),
                ('a1dot_sigma', c_double_p),
                ('n_sini', ctypes.c_int32),
                ('sini_col', c_int32_p),
                ('sini_mu', c_double_p),
                ('sini_sigma', c_double_p),
                ('mas2rad', ctypes.c_double),
                ('deg2rad', ctypes.c_double),
                ('masyr2rads', ctypes.c_double)]


# every symbol include/sterne_b200.h declares: (name, restype, argtypes)
_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
SYMBOLS = [
    ('stn_version', _int, []),
    ('stn_last_error', ctypes.c_char_p, []),
    ('stn_create', _int, [ctypes.POINTER(StnProblem), _int, ctypes.POINTER(_vp)]),
    ('stn_destroy', _int, [_vp]),
    ('stn_loglike', _int, [_vp, _vp, _i64, _i64, _int, _int, _int, _vp, _vp]),
    ('stn_set_prior', _int, [_vp, ctypes.POINTER(StnPrior)]),
    ('stn_logpost', _int, [_vp, _vp, _i64, _i64, _int, _int, _int, _vp, _vp]),
    ('stn_chisq', _int, [_vp, _vp, _i64, _i64, _int, _int, _int, _vp, _vp]),
    ('stn_loglike_host', _int, [_vp, _vp, _i64, _i64, _int, _int, _int, _vp]),
    ('stn_model', _int, [_vp, _int, _vp, _i64, _i64, _int, _int, _vp, _vp, _vp]),
    ('stn_auto_plan', _int, [_vp, _i64]),
    ('stn_launch_count', _i64, [_vp]),
    ('stn_host_alloc', _int, [ctypes.POINTER(_vp), _i64]),
    ('stn_host_free', _int, [_vp]),
    ('stn_fp64_peak', _int, [_int, _int, _int, c_double_p, c_double_p]),
    ('stn_stretch_propose', _int, [_int, _vp, _i64, ctypes.c_int32, _i64, _int, ctypes.c_double, ctypes.c_uint64,
                                   ctypes.c_uint64, _vp, _vp, _vp]),
    ('stn_stretch_accept', _int, [_int, _vp, _vp, _i64, ctypes.c_int32, _i64, _int, ctypes.c_uint64, ctypes.c_uint64,
                                  _vp, _vp, _vp, _vp, _vp]),
    ('stn_loglike_timed', _int, [_vp, _vp, _i64, _i64, _int, _int, _int, _vp, _int, c_double_p]),
    ('stn_host_register', _int, [_vp, _i64]),
    ('stn_host_unregister', _int, [_vp]),
    ('stn_host_hint', _int, [_vp, _int]),
    ('stn_multi_create', _int, [ctypes.POINTER(StnProblem), c_int32_p, _int, ctypes.POINTER(_vp)]),
    ('stn_multi_destroy', _int, [_vp]),
    ('stn_multi_size', _int, [_vp]),
    ('stn_multi_ctx', _vp, [_vp, _int]),
    ('stn_multi_launch_count', _i64, [_vp]),
    ('stn_multi_slice', _int, [_i64, _int, _int, ctypes.POINTER(_i64), ctypes.POINTER(_i64)]),
    ('stn_multi_loglike_host', _int, [_vp, _vp, _i64, _i64, _int, _int, _int, _vp]),
    ('stn_multi_loglike_dev', _int, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_i64), _i64, _int, _int, _int, _vp,
                                     _int]),
    ('stn_loglike_scatter', _int, [_vp, _vp, _i64, _i64, _int, _int, _int, ctypes.POINTER(_vp), _int, _vp]),
    ('stn_dev_alloc', _int, [_int, ctypes.POINTER(_vp), _i64]),
    ('stn_dev_free', _int, [_int, _vp]),
    ('stn_ipc_export', _int, [_vp, _vp]),
    ('stn_ipc_open', _int, [_int, _vp, ctypes.POINTER(_vp)]),
    ('stn_ipc_close', _int, [_int, _vp]),
    ('stn_multi_set_prior', _int, [_vp, ctypes.POINTER(StnPrior)]),
    ('stn_multi_ensemble_run', _int, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i64, _i64, ctypes.c_double,
                                      ctypes.c_uint64, ctypes.c_uint64, _i64, _i64, _vp, _vp, ctypes.POINTER(_i64), _int,
                                      _int]),
    ('stn_ensemble_run', _int, [_vp, _vp, _vp, _i64, _i64, ctypes.c_double, ctypes.c_uint64, ctypes.c_uint64, _i64, _i64,
                                _vp, _vp, _vp, _int, _int, _int, _vp]),
]
ABI_VERSION = 2
IPC_HANDLE_BYTES = 64
MAX_OUTS = 8


class StnError(RuntimeError):
    pass


_lib = None


def library_path():
    """STERNE_B200_LIB selects an alternative build of the same source (tuning experiments)."""
    return os.environ.get('STERNE_B200_LIB') or _build.LIB


def load(build_if_missing=True):
    """Loads the shared library (building it with nvcc first when it is absent and a
    compiler is at hand).  Raises if that fails -- nothing falls back to the CPU."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if build_if_missing and path == _build.LIB and _build.needs_build():
        try:
            _build.build_library()
        except Exception as err:
            if not os.path.exists(path):
                raise StnError('libsterne_b200.so is missing and could not be built (%s); '
                               'sterne_b200 has no CPU fallback' % (err,))
    if not os.path.exists(path):
        raise StnError('%s not found; run `python -m sterne_b200.build` (no CPU fallback)' % path)
    lib = ctypes.CDLL(path)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)            # AttributeError if the ABI is incomplete
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.stn_version() != ABI_VERSION:
        raise StnError('%s has ABI version %d, this package binds version %d; rebuild with '
                       '`python -m sterne_b200.build --force`' % (path, lib.stn_version(), ABI_VERSION))
    _lib = lib
    return lib


def make_plan(lanes=1, splits=1):
    """plan = lanes | (splits << 8) (include/sterne_b200.h, stn_loglike)."""
    return int(lanes) | (int(splits) << 8)


def plan_parts(plan):
    return plan & 0xff, (plan >> 8) & 0xffff


def check(rc):
    if rc != 0:
        msg = load().stn_last_error()
        raise StnError('sterne_b200 error %d: %s' % (rc, msg.decode() if msg else '?'))
