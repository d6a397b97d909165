"""In-tree build of libsynchrad_b200.so (nvcc, sm_100a only; cross-compiles without a GPU)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
OUT = os.path.join(CSRC, 'libsynchrad_b200.so')
SOURCES = ['srb_api.cu']
# every source under csrc/ is compiled into the one library (srb_api.cu includes the rest)
DEPS = sorted(f for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh'))) + [os.path.join('..', '..', 'include', 'synchrad_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC', '-diag-suppress', '177']


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', OUT] + SOURCES
    subprocess.check_call(cmd, cwd=CSRC)
    return OUT


if __name__ == '__main__':
    import sys
    build(force='--force' in sys.argv, verbose='-v' in sys.argv)
    print(OUT)
