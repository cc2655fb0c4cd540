"""ctypes binding of libvalle_b200.so (C ABI declared in include/valle_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'lib', 'libvalle_b200.so')

VB_F32, VB_BF16 = 0, 1
EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL = 0, 1, 2, 3
MASK_NONE, MASK_PREFIX_LM, MASK_EXPLICIT = 0, 1, 2
FLAG_LATE_TRIGGER, FLAG_PREFETCH_KV, FLAG_ATTN_SIMT, FLAG_ATTN_TICKET, FLAG_DG_GLOBAL = 1, 2, 4, 8, 16

_p, _i, _i64, _f, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint64

# symbol -> (restype, argtypes); must list every function of include/valle_b200.h
SIGNATURES = {
    'vb_version': (_i, []),
    'vb_last_error_string': (C.c_char_p, []),
    'vb_device_info': (_i, [_p, _p, _p, _p]),
    'vb_embed_sum_pe_norm': (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i, _i64, _i64, _p, _p, _f, _p, _i, _p]),
    'vb_embed_sum_pe': (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i, _i64, _i64, _p]),
    'vb_residual_layernorm': (_i, [_p, _p, _i, _i64, _p, _p, _p, _p, _i, _i64, _i, _f, _p]),
    'vb_reduce_bias_act': (_i, [_p, _i, _i64, _p, _i, _p, _i, _i64, _i, _p]),
    'vb_linear': (_i, [_p, _i, _i64, _p, _i, _i64, _p, _p, _i64, _p, _i, _i64, _i64, _i64, _i64, _i, _p]),
    'vb_linear_categorical': (_i, [_p, _i64, _p, _i64, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _f, _u64, _i, _p]),
    'vb_linear_argmax': (_i, [_p, _i64, _p, _i64, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _p]),
    'vb_linear_t': (_i, [_p, _i64, _i, _p, _i64, _i, _p, _p, _i64, _p, _i, _i64, _i64, _i64, _i64, _i, _p]),
    'vb_linear_decode_splits': (_i, [_i64, _i64, _i]),
    'vb_linear_decode_splits_m': (_i, [_i64, _i64, _i64, _i]),
    'vb_linear_decode': (_i, [_p, _i64, _p, _i64, _p, _i64, _i64, _i64, _i64, _i, _i, _p, _p]),
    'vb_linear_decode_rows_splits': (_i, [_i, _i64, _i]),
    'vb_linear_decode_rows_ln': (_i, [_p, _i64, _p, _p, _f, _p, _i64, _p, _p, _i, _i64, _i, _i64, _i64, _i, _i, _p]),
    'vb_linear_decode_rows': (_i, [_p, _i64, _p, _i64, _p, _p, _i, _i64, _i64, _i, _i64, _i64, _i, _i, _i, _p, _p]),
    'vb_decode_gemm_plan': (_i, [_i, _i64, _i64, _p, _p, _p]),
    'vb_decode_gemm': (_i, [_p, _i64, _p, _i64, _i, _i64, _i64, _i, _p, _p, _p, _i, _f, _p, _i64, _p, _i64, _p, _i64, _p, _p, _p,
                            _i, _p]),
    'vb_decode_gemm_set_debug': (_i, [_p]),
    'vb_ar_step_tail': (_i, [_p, _i64, _i, _f, _i, _f, _p, _p, _i, _p, _p, _p, _i64, _p, _p, _p, _i, _i, _p, _p, _i, _p, _p, _p,
                             _p]),
    'vb_kv_prefetch_l2': (_i, [_p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _p]),
    'vb_transpose': (_i, [_p, _i, _i64, _i64, _i64, _p, _i64, _p]),
    'vb_colsum': (_i, [_p, _i, _i64, _i, _i64, _p, _i, _f, _p]),
    'vb_colsum_blocks': (_i, [_p, _i, _i64, _i, _i64, _p, _i, _p]),
    'vb_gelu_fwd': (_i, [_p, _i, _p, _i64, _p]),
    'vb_gelu_bwd': (_i, [_p, _p, _i, _p, _i64, _p]),
    'vb_dropout': (_i, [_p, _i, _i64, _f, _u64, _u64, _p]),
    'vb_dropout_add': (_i, [_p, _p, _i, _i64, _f, _u64, _u64, _p]),
    'vb_layernorm_bwd_blocks': (_i, [_i64]),
    'vb_layernorm_bwd': (_i, [_p, _p, _p, _i, _p, _p, _p, _i64, _i, _f, _p]),
    'vb_attention_bwd': (_i, [_p, _p, _p, _p, _i, _p, _p, _i, _i, _i, _i, _i, _p, _p, _i, _p]),
    'vb_cross_entropy': (_i, [_p, _i64, _p, _i64, _i, _p, _p, _i64, _f, _p]),
    'vb_embed_bwd': (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i64, _i64, _p]),
    'vb_linear_decode_set_debug': (_i, [_p]),
    'vb_attn_decode_set_debug': (_i, [_p]),
    'vb_attention_prefill_set_debug': (_i, [_p]),
    'vb_residual_layernorm_set_debug': (_i, [_p]),
    'vb_linear_decode_rows_set_debug': (_i, [_p]),
    'vb_attention': (_i, [_p, _p, _p, _i, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _p, _i, _i64, _i64,
                          _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _i64, _i64, _i64, _p]),
    'vb_attention_prefill_tc': (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    'vb_kv_scatter_paged': (_i, [_p, _i, _p, _i, _p, _i, _p, _i, _i, _i, _i, _p]),
    'vb_attn_decode_ws_bytes': (_i64, [_i, _i, _i]),
    'vb_attn_decode_paged': (_i, [_p, _i, _i64, _p, _i, _p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p]),
    'vb_sample': (_i, [_p, _i, _i64, _i64, _i, _i, _f, _i, _f, _p, _u64, _p, _i, _p, _p, _p]),
    'vb_ar_bookkeeping': (_i, [_p, _p, _p, _p, _p, _i64, _p, _p, _p, _i, _i, _p]),
}


class VBError(RuntimeError):
    pass


_lib = None


def load(build_if_missing: bool = True):
    """Load (building first if the in-tree .so is absent or stale and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        try:
            from . import build as _build
            if _build.needs_build():
                _build.build()
        except Exception as e:  # nvcc missing etc.: fall through to the loud failure below
            if not os.path.exists(LIB_PATH):
                raise VBError(f'libvalle_b200.so is missing and could not be built: {e}') from e
    if not os.path.exists(LIB_PATH):
        raise VBError(f'{LIB_PATH} not found -- run `python -m valle2_b200.build` (there is no CPU fallback)')
    lib = C.CDLL(os.environ.get('VALLE_B200_LIB') or LIB_PATH)      # override: experiment builds (tools/build_variant.sh)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)            # AttributeError if the symbol is missing: loud by design
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


LAUNCHES = 0            # kernel-launching C calls made through check() (bench.py's gpu_launches claim; graph replays are
_NO_LAUNCH = {'vb_device_info', 'vb_decode_gemm_plan'}    # counted separately by the engine)


def check(rc: int, what: str = '') -> None:
    global LAUNCHES
    if what not in _NO_LAUNCH:
        LAUNCHES += 1
    if rc != 0:
        msg = load().vb_last_error_string().decode(errors='replace')
        raise VBError(f'{what or "libvalle_b200"} failed with status {rc}: {msg}')
