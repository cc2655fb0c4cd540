"""Model registry -- same names as the reference's ``valle/models/__init__.py``."""
from .valle_ar import ValleAR
from .valle_nar import ValleNAR


class _EncodecUnavailable:
    """EnCodec wrapper is outside the hot-path scope; the ``encodec`` package and its weights are
    not available offline.  Constructing it says so instead of failing at import time."""

    def __init__(self, *a, **k):
        raise RuntimeError('EncodecPip is out of scope for valle2_b200 (needs the `encodec` package + weights)')


try:  # pragma: no cover
    from encodec import EncodecModel  # noqa: F401
    EncodecPip = _EncodecUnavailable
except Exception:
    EncodecPip = _EncodecUnavailable

MODEL_DICT = {'EncodecPip': EncodecPip, 'ValleAR': ValleAR, 'ValleNAR': ValleNAR}


def get_model_class(model_name: str):
    return MODEL_DICT[model_name]


__all__ = ['EncodecPip', 'ValleAR', 'ValleNAR', 'MODEL_DICT', 'get_model_class']
