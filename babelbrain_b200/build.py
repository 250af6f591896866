"""Builds libbabelb200.so in-tree with nvcc for sm_100a (no torch extension machinery needed:
the boundary is a plain C ABI loaded with ctypes)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.environ.get('BB_LIB', os.path.join(HERE, 'libbabelb200.so'))   # BB_LIB / BB_NVCC_FLAGS: build experiment variants
SOURCES = ['fdtd.cu', 'rayleigh.cu', 'bhte.cu']
HEADERS = ['common.h', 'fdtd_cell.cuh', 'fdtd_kernels.cuh', 'fdtd_direct.cuh', 'fdtd_tma.cuh', 'fdtd_tma2.cuh', 'nccl_dyn.h',
           os.path.join('..', '..', 'include', 'babelb200.h')]


def nccl_include():
    try:
        import nvidia.nccl
        return os.path.join(list(nvidia.nccl.__path__)[0], 'include')
    except Exception:
        return '/usr/include'


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
           '-Xcompiler', '-fPIC', '-shared', '-I' + nccl_include(), '-o', LIB] + \
          [os.path.join(CSRC, s) for s in SOURCES] + ['-ldl'] + os.environ.get('BB_NVCC_FLAGS', '').split()
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
