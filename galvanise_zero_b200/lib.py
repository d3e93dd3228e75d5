"""ctypes loader for libgz_b200.so (include/gz_b200.h). There is no fallback: if the library is
missing or a call fails, an exception is raised."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'csrc', 'libgz_b200.so')


class GzError(RuntimeError):
    """Stands where the reference has ggpzero_interface.GGPZeroInterfaceError
    (src/cpp/ggpzero_interface.cpp:89-92)."""


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GzError('libgz_b200.so is not built (%s); run `python -m galvanise_zero_b200.build` '
                      '— there is no CPU fallback' % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i64, u64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int
    L.gz_last_error.restype = ctypes.c_char_p
    L.gz_version.restype = ci
    L.gz_propnet_create.argtypes = [vp, i64, ci, ctypes.POINTER(vp)]
    L.gz_propnet_destroy.argtypes = [vp]
    L.gz_propnet_info.argtypes = [vp, ci]
    L.gz_propnet_info.restype = i64
    L.gz_sm_eval.argtypes = [vp, i64, vp, vp, vp, vp, ci, vp]
    L.gz_sm_transition.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, ci, vp]
    L.gz_sm_playout_step.argtypes = [vp, i64, i64, u64, u64, vp, vp, vp, vp, vp, vp, ci, vp]
    L.gz_sm_transition_host.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, ci]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise GzError('libgz_b200: %s (status %d)' % (lib().gz_last_error().decode(), rc))


# info selectors (include/gz_b200.h)
INFO_N_BASES, INFO_N_ROLES, INFO_STATE_WORDS, INFO_LEGAL_WORDS, INFO_VAL_WORDS, \
    INFO_WARPS_PER_CTA, INFO_SMEM_BYTES, INFO_GRID, INFO_STAGED_MASK, INFO_N_INPUTS = range(10)
INFO_ROLE_SIZE0 = 16
INFO_LEGAL_WORD_OFF0 = 24
GOALS_ALWAYS = 0
GOALS_TERMINAL_ONLY = 1
