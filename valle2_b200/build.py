"""Build libvalle_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m valle2_b200.build [--force]

The shared library is a plain C-ABI (include/valle_b200.h); it links the CUDA runtime statically and
resolves the one driver symbol it needs (cuTensorMapEncodeTiled) at run time, so it loads on a
GPU-less host too (symbols only -- no compute without a device).
"""
from __future__ import annotations

import concurrent.futures
import fcntl
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


HASH = os.path.join(LIBDIR, 'libvalle_b200.srchash')


def _source_hash() -> str:
    """Content hash of everything the library is built from (file times do not survive a copy of the tree to a GPU box)."""
    import hashlib
    h = hashlib.sha256()
    deps = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh'))
    deps.append(os.path.join(os.path.dirname(HERE), 'include', 'valle_b200.h'))
    for path in deps:
        h.update(os.path.basename(path).encode())
        with open(path, 'rb') as fh:
            h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(HASH):
        return True
    with open(HASH) as fh:
        return fh.read().strip() != _source_hash()


def _compile(src: str) -> str:
    obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + '.o')
    cmd = ['nvcc', *NVCC_FLAGS, '-c', src, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile + link if the library is missing or older than its sources.  Safe under concurrency (one process per GPU
    under torchrun, all importing at once): an exclusive file lock serialises the builders, the check is repeated under
    the lock, and the library is linked to a temporary name and renamed into place atomically."""
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    with open(os.path.join(LIBDIR, '.build.lock'), 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():        # another process built it while we waited
                return LIB
            with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
                objs = list(ex.map(_compile, sources()))
            tmp = f'{LIB}.tmp.{os.getpid()}'
            cmd = ['nvcc', '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', tmp, *objs, '-cudart', 'static']
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
            os.replace(tmp, LIB)
            with open(HASH + '.tmp', 'w') as fh:
                fh.write(_source_hash())
            os.replace(HASH + '.tmp', HASH)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    if verbose:
        print('built', LIB)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True)
