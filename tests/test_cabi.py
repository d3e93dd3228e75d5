"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/gz_b200.h
declares (no compute calls — there is no GPU here)."""
import ctypes
import os
import re

from galvanise_zero_b200 import build as gzbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'gz_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(gz_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_declared_symbols():
    path = gzbuild.build_extension()
    L = ctypes.CDLL(path)
    syms = declared_symbols()
    assert 'gz_sm_transition' in syms and 'gz_propnet_create' in syms
    for s in syms:
        assert hasattr(L, s), 'missing export: ' + s


def test_error_paths_without_gpu():
    from galvanise_zero_b200 import lib
    L = lib.lib()
    assert L.gz_version() >= 100
    h = ctypes.c_void_p()
    rc = L.gz_propnet_create(None, 0, 0, ctypes.byref(h))
    assert rc != 0 and h.value is None
    bad = (ctypes.c_int32 * 80)()
    rc = L.gz_propnet_create(bad, 80, 0, ctypes.byref(h))
    assert rc == -3 and b'blob' in L.gz_last_error()
