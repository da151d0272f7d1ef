"""Builds the CUDA C-ABI library in-tree (nvcc cross-compiles for sm_100a
without a GPU):  python -m pauxy_b200.build"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(CSRC, 'libpauxy_b200.so')
SOURCES = ['pxb_api.cu']
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith('.cuh')) + [os.path.join('..', '..', 'include', 'pauxy_b200.h')]
# --split-compile=0: the optimisation phase of the single translation unit runs on every host core
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
              '-std=c++17', '-Xcompiler', '-fPIC', '-shared', '--split-compile=0']


def needs_build():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = SOURCES + HEADERS + [os.path.join('..', 'build.py')]
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    # PXB_EXTRA_NVCC_FLAGS: development builds only (e.g. -DPXB_EXPERIMENTS for the timing switches)
    extra = os.environ.get('PXB_EXTRA_NVCC_FLAGS', '').split()
    cmd = [nvcc] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libpauxy_b200.so")
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
