"""TEST / BASELINE INFRASTRUCTURE ONLY -- install the UNMODIFIED reference into oracle/_ref/ (git-ignored).

    python -m oracle.build_ref            (authoring container: needs /root/reference)

The reference is a pure-Python package (valle/, 1.3 kLoC); "building" it means installing it where the GPU box can import
it: `/root/reference` does not exist there, but `oracle/_ref/` travels with the repo snapshot (it is git-ignored, not
gpurun-ignored).  Recipe, in order:
  1. `pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy of /root/reference>` from a copy
     under /tmp (the source tree is read-only).  The reference declares a poetry-core build backend; when that backend is
     not in the offline wheelhouse this step fails and
  2. the package directory `valle/` is copied verbatim into oracle/_ref/valle -- the same files a wheel install would
     place there (there is no native code and no generated file in the reference).
`oracle/_ref/INSTALL.json` records the method and a SHA-256 per installed file, so that `bench.py --impl reference` and the
tests can state exactly what they executed.  Nothing under oracle/_ref is ever committed (reference sources stay out of the
repository's history), and the product path never imports it: only `bench.py --impl reference` / `cpu_baseline`, through
`oracle/ref_shims.py` (four import stubs for packages that are absent offline).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get('VALLE_REFERENCE_SRC', '/root/reference')
DEST = os.path.join(HERE, '_ref')


def _hashes(root: str) -> dict:
    out = {}
    for dp, _, files in os.walk(root):
        for f in sorted(files):
            if f.endswith('.py'):
                p = os.path.join(dp, f)
                with open(p, 'rb') as fh:
                    out[os.path.relpath(p, root)] = hashlib.sha256(fh.read()).hexdigest()
    return out


def installed() -> bool:
    return os.path.isfile(os.path.join(DEST, 'valle', 'models', 'valle_ar.py'))


def build(force: bool = False, verbose: bool = True) -> str | None:
    """Install the reference into oracle/_ref.  Returns the method used, or None when the source tree is absent."""
    if not os.path.isdir(os.path.join(REF_SRC, 'valle')):
        if verbose:
            print(f'oracle/_ref: reference tree {REF_SRC} not present (GPU box): using the prebuilt copy'
                  if installed() else f'oracle/_ref: reference tree {REF_SRC} not present and nothing installed')
        return None
    src_hash = _hashes(os.path.join(REF_SRC, 'valle'))
    meta_path = os.path.join(DEST, 'INSTALL.json')
    if installed() and not force and os.path.exists(meta_path):
        with open(meta_path) as fh:
            if json.load(fh).get('files') == {os.path.join('valle', k): v for k, v in src_hash.items()}:
                return 'cached'
    shutil.rmtree(DEST, ignore_errors=True)
    os.makedirs(DEST, exist_ok=True)
    method, log = None, ''
    with tempfile.TemporaryDirectory(prefix='valle_ref_') as tmp:
        copy = os.path.join(tmp, 'reference')
        shutil.copytree(REF_SRC, copy)
        cmd = [sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--no-deps', '--find-links',
               '/opt/wheelhouse', '--target', DEST, copy]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = (r.stdout + r.stderr)[-600:]
        if r.returncode == 0 and installed():
            method = 'pip install --no-index --no-build-isolation --no-deps --target oracle/_ref'
    if method is None:
        shutil.rmtree(DEST, ignore_errors=True)
        os.makedirs(DEST, exist_ok=True)
        shutil.copytree(os.path.join(REF_SRC, 'valle'), os.path.join(DEST, 'valle'))
        method = 'verbatim copy of the package directory valle/ (pip could not build: ' + log.strip().splitlines()[-1][:160] + ')' \
            if log.strip() else 'verbatim copy of the package directory valle/'
    files = {k: v for k, v in _hashes(DEST).items() if k.startswith('valle' + os.sep)}
    assert files == {os.path.join('valle', k): v for k, v in src_hash.items()}, 'installed files differ from the reference sources'
    with open(meta_path, 'w') as fh:
        json.dump({'method': method, 'source': REF_SRC, 'files': files}, fh, indent=1)
    if verbose:
        print('oracle/_ref:', method, f'({len(files)} files)')
    return method


if __name__ == '__main__':
    build(force='--force' in sys.argv)
