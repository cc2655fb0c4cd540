"""valle2_b200 -- B200-native (sm_100a) implementation of the Valle2 decoding hot path.

Host side: Python mirror of the reference's ``valle.models`` API (same module / class / method names
and ``state_dict`` keys).  Device side: hand-written CUDA kernels behind the C ABI in
``include/valle_b200.h`` (``valle2_b200/lib/libvalle_b200.so``).  No CPU fallback.
"""
import os

_PRECISION = os.environ.get('VALLE_B200_PRECISION', 'bf16')


def set_precision(mode: str) -> None:
    """'bf16' (tcgen05 GEMMs, bf16 KV cache, fp32 accumulate) or 'fp32' (validation mode, SIMT fp32)."""
    global _PRECISION
    if mode not in ('bf16', 'fp32'):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    _PRECISION = mode


def get_precision() -> str:
    return _PRECISION


from .config import ConfigValle  # noqa: E402

__all__ = ['ConfigValle', 'set_precision', 'get_precision']
