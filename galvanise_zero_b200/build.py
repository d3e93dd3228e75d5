"""In-tree build of libgz_b200.so (hand-written sm_100a CUDA + the C-ABI of include/gz_b200.h)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(CSRC, 'libgz_b200.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-fmad=false',
              '-Xcompiler', '-fPIC', '-shared']


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh', '.h'))]
    inc = os.path.join(HERE, '..', 'include')
    deps += [os.path.join(inc, f) for f in os.listdir(inc)]
    return any(os.path.getmtime(d) > t for d in deps)


MODULE = os.path.join(CSRC, 'ggpzero_interface.so')       # importable as `ggpzero_interface` with csrc on sys.path


def build_module():
    """The reference's extension module for CPython 3 (csrc/ggpzero_interface.cpp), linked against libgz_b200.so."""
    import sysconfig
    import numpy
    cuda = os.environ.get('CUDA_HOME', '/usr/local/cuda')
    cmd = [os.environ.get('CXX', 'g++'), '-O2', '-std=c++17', '-fPIC', '-shared', '-Wall', '-Wno-missing-field-initializers',
           '-I' + sysconfig.get_paths()['include'], '-I' + numpy.get_include(), '-I' + os.path.join(cuda, 'include'),
           '-o', MODULE, os.path.join(CSRC, 'ggpzero_interface.cpp'), '-L' + CSRC, '-l:libgz_b200.so',
           '-L' + os.path.join(cuda, 'lib64'), '-lcudart', '-Wl,-rpath,$ORIGIN', '-Wl,-rpath,' + os.path.join(cuda, 'lib64')]
    subprocess.check_call(cmd)
    return MODULE


def build_extension(force=False, verbose=False):
    if not force and not needs_build() and os.path.exists(MODULE) and \
            os.path.getmtime(MODULE) >= os.path.getmtime(os.path.join(CSRC, 'ggpzero_interface.cpp')):
        return LIB
    if force or needs_build():
        nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB] + sources()
        subprocess.check_call(cmd)
    build_module()
    return LIB


if __name__ == '__main__':
    build_extension(force='--force' in sys.argv, verbose='-v' in sys.argv)
    print(LIB)
