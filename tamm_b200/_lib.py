"""ctypes binding of libtamm_b200.so (include/tamm_b200.h).  Fails loudly when the library is missing."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB_PATH = os.path.join(_HERE, 'libtamm_b200.so')
# TAMM_B200_LIB selects another build of the SAME library (A/B experiments, tools/ab_variant.sh) without ever replacing
# the in-tree file; `fingerprint()` says which file is loaded so that bench lines and test logs record it.
LIB_PATH = os.environ.get('TAMM_B200_LIB') or DEFAULT_LIB_PATH

_vp = ctypes.c_void_p
_i = ctypes.c_int
_d = ctypes.c_double

# name -> (restype, argtypes); must list every symbol declared in include/tamm_b200.h
SIGNATURES = {
    'tamm_version': (_i, []),
    'tamm_last_error': (ctypes.c_char_p, []),
    'tamm_check_device': (_i, [_i]),
    'tamm_fp64_probe': (_i, [_vp, _i, _i, _vp]),
    'tamm_layer_matrices': (_i, [_vp, _vp, _i, _i, _d, _d] + [_vp] * 10 + [_vp, _vp]),
    'tamm_layer_modes': (_i, [_vp, _vp, _i, _i, _d, _d, _vp, _vp, _vp]),
    'tamm_peer_push': (_i, [_vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_ulonglong), _i, _i, _i, _vp]),
    'tamm_stack_workspace_bytes': (ctypes.c_size_t, [_i, _i, _i, _i]),
    'tamm_suggest_chunk': (_i, [_i, _i, _i, _i]),
    'tamm_stack_urd': (_i, [_vp, _vp, _i, _i, _i, _i, _d, _d, _i, _i, _vp, _vp, _vp, _vp]),
    'tamm_stack_urd_fp': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    'tamm_export_jones': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    'tamm_returns': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, ctypes.POINTER(_d), _i] + [_vp] * 7 + [_vp]),
    'tamm_stack_returns': (_i, [_vp, _vp, _i, _i, _i, _i, _d, _d, ctypes.POINTER(_d), _i, _i] + [_vp] * 7 + [_vp, _vp]),
    'tamm_power_inplace': (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    'tamm_returns_partial': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, ctypes.POINTER(_d), _i] + [_vp] * 7 + [_vp]),
    'tamm_returns_finish': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
}

_lib = None


class TammLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  There is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TammLibraryError(
            'libtamm_b200.so not found at %s -- build it with `python -c "import __graft_entry__ as g; g.build()"` '
            '(nvcc, sm_100a).  tamm_b200 has no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def fingerprint():
    """(path, md5 of the file) of the library in use -- recorded in the bench JSON line."""
    import hashlib
    with open(LIB_PATH, 'rb') as f:
        return LIB_PATH, hashlib.md5(f.read()).hexdigest()


def check(rc, lib=None):
    if rc != 0:
        lib = lib or load()
        raise TammLibraryError(lib.tamm_last_error().decode('utf-8', 'replace'))
