"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads without a GPU, and exports every
symbol that include/valle_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'valle_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(vb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from valle2_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/valle_b200.h but not exported'
    assert sorted(_lib.SIGNATURES) == declared, 'ctypes signature table out of sync with the header'
    assert lib.vb_version() >= 100
    assert isinstance(lib.vb_last_error_string(), bytes)


def test_pure_queries_need_no_gpu():
    from valle2_b200 import _lib
    lib = _lib.load()
    assert lib.vb_attn_decode_ws_bytes(32, 16, 4) > 32 * 16 * 4 * 64 * 4
    ns = lib.vb_linear_decode_splits(3072, 1024, 32)
    assert 1 <= ns <= 8
    assert lib.vb_linear_decode_splits_m(64, 3072, 1024, 32) == ns                 # up to 128 batch rows: the same
    assert 1 <= lib.vb_linear_decode_splits_m(256, 3072, 1024, 32) <= ns           # above: batch tiles, never more slices


def test_bad_arguments_are_reported_not_crashed():
    from valle2_b200 import _lib
    lib = _lib.load()
    rc = lib.vb_sample(None, 1, 0, 0, 1, 10, ctypes.c_float(1.0), 1, ctypes.c_float(1.0), None, 0, None, 0, None, None, None)
    assert rc == -1 and b'null' in lib.vb_last_error_string()


def test_no_cpu_fallback():
    import pytest
    import torch
    from valle2_b200 import ops
    with pytest.raises(Exception):
        ops.linear(torch.zeros(4, 64), torch.zeros(8, 64))


def test_header_is_plain_c_and_a_c_program_links_through_it(tmp_path):
    """The boundary is a C ABI: include/valle_b200.h must compile as C99 (no C++ in the signatures), and a C translation unit
    that includes it must link against libvalle_b200.so and call into it (pure queries and an argument check only: no GPU here)."""
    import os
    import shutil
    import subprocess
    import pytest
    from valle2_b200 import _lib
    if shutil.which('gcc') is None:
        pytest.skip('gcc not available')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / 'use_abi.c'
    src.write_text('''
#include <stdio.h>
#include <string.h>
#include "valle_b200.h"
int main(void) {
    int splits = vb_linear_decode_splits(3072, 1024, 32);
    int rc = vb_linear_argmax(NULL, 0, NULL, 0, NULL, NULL, 1, 1, 1, 2048, 1024, 1024, NULL);
    printf("%d %d %d %d\\n", vb_version(), splits, rc, (int)(strstr(vb_last_error_string(), "null") != NULL));
    return 0;
}
''')
    exe = tmp_path / 'use_abi'
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-pedantic', '-I', os.path.join(root, 'include'), str(src), '-o', str(exe),
                    '-L', libdir, '-lvalle_b200', f'-Wl,-rpath,{libdir}'], check=True, capture_output=True, text=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) >= 100 and 1 <= int(out[1]) <= 8 and int(out[2]) == -1 and int(out[3]) == 1


def test_a_missing_library_fails_loudly():
    """No CPU fallback: with the extension absent (VALLE_B200_LIB pointing at a file that does not exist) the first op raises --
    a fresh interpreter, so the already-loaded handle of this process is not in the way."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ('import sys; sys.path.insert(0, %r)\n'
            'from valle2_b200 import _lib\n'
            'try:\n'
            '    _lib.load()\n'
            'except OSError as e:\n'
            '    print("LOUD", type(e).__name__)\n'
            'else:\n'
            '    print("SILENT")\n') % root
    env = dict(os.environ, VALLE_B200_LIB='/nonexistent/libvalle_b200.so')
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300, env=env)
    assert 'LOUD' in out.stdout and 'SILENT' not in out.stdout, (out.stdout, out.stderr[-500:])
