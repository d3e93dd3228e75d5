"""CPU: csrc/gz_libm.h (glibc logf / powf, libstdc++ generate_canonical / normal / gamma, restated so that the
device draws the reference's Dirichlet noise) against the host libraries themselves, bit for bit
(tests/hostsim/libmcheck.cpp)."""
import ctypes
import os
import subprocess

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'hostsim')


def _lib():
    subprocess.check_call(['make', '-s', '-C', HERE, 'liblibmcheck.so'])
    L = ctypes.CDLL(os.path.join(HERE, 'liblibmcheck.so'))
    cl, u64, u32, vp, ci = ctypes.c_long, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int
    L.libmcheck_logf.restype = cl
    L.libmcheck_logf.argtypes = [cl, u64, u32, u32, vp]
    L.libmcheck_powf.restype = cl
    L.libmcheck_powf.argtypes = [cl, u64, ci, vp]
    L.libmcheck_gamma.restype = cl
    L.libmcheck_gamma.argtypes = [ci, ci, u64, vp]
    L.libmcheck_canonical.restype = cl
    L.libmcheck_canonical.argtypes = [cl, u64]
    return L


def test_logf_is_glibc_logf():
    fb = (ctypes.c_uint32 * 2)()
    # 4 M random bit patterns (every exponent, both signs, NaNs) + every float in [0.5, 2)
    assert _lib().libmcheck_logf(4000000, 1, 0x3f000000, 0x40000000, fb) == 0, hex(fb[0])


def test_powf_is_glibc_powf():
    """glibc selects an FMA build of powf at run time on CPUs that have it; the restatement is the plain
    algorithm, so the two may differ in the last bit when the double result sits within 2^-53 of a float
    rounding boundary (observed: 1 in 4e7). Bar: at most 2 in 8e6."""
    fb = (ctypes.c_uint32 * 2)()
    assert _lib().libmcheck_powf(4000000, 2, 4000, fb) <= 2, (hex(fb[0]), hex(fb[1]))


def test_canonical_float_is_generate_canonical():
    assert _lib().libmcheck_canonical(2000000, 3) == 0


def test_gamma_is_std_gamma_distribution():
    """alpha = 10.83 / n for n = 2..400 children (both the alpha >= 1 and the pow-boosted alpha < 1 branch),
    n draws per distribution object, 10 seeds each: same floats, same number of generator draws."""
    nd = ctypes.c_long()
    assert _lib().libmcheck_gamma(400, 10, 4, ctypes.byref(nd)) == 0
    assert nd.value > 500000
