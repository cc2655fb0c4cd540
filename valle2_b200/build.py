"""Build libvalle_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m valle2_b200.build [--force]

The shared library is a plain C-ABI (include/valle_b200.h); it links the CUDA runtime statically and
resolves the one driver symbol it needs (cuTensorMapEncodeTiled) at run time, so it loads on a
GPU-less host too (symbols only -- no compute without a device).
"""
from __future__ import annotations

import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
OBJDIR = os.path.join(HERE, 'build')
LIB = os.path.join(LIBDIR, 'libvalle_b200.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr',
]


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _newest(paths) -> float:
    return max(os.path.getmtime(p) for p in paths)


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    deps.append(os.path.join(os.path.dirname(HERE), 'include', 'valle_b200.h'))
    return _newest(deps) > os.path.getmtime(LIB)


def _compile(src: str) -> str:
    obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + '.o')
    cmd = ['nvcc', *NVCC_FLAGS, '-c', src, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(_compile, sources()))
    cmd = ['nvcc', '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB, *objs, '-cudart', 'static']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    if verbose:
        print('built', LIB)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True)
