"""TEST INFRASTRUCTURE ONLY -- make the real reference importable offline.

The reference (``/root/reference``, read-only, Python) needs four packages that are absent here
(``lightning``, ``encodec``, ``coloredlogs``, ``torchaudio`` may be absent too) and one function
that transformers 5.x no longer ships (``transformers.generation.utils.top_k_top_p_filtering``,
pinned at 4.38.2 by the reference's poetry.lock:3600).  This module registers minimal stand-ins in
``sys.modules`` *before* ``valle`` is imported, then imports the unmodified reference modules.

The reference is executed from ``/root/reference`` where that exists (the authoring container) and otherwise from
``oracle/_ref`` (the unmodified package installed by ``oracle/build_ref.py``; git-ignored, shipped to the GPU box with the
repo snapshot).  Users: ``oracle/make_golden.py`` (golden vectors), ``bench.py --impl reference`` and its ``cpu_baseline``
leg (the executed reference timed on the host cores).  The product path never imports this module.
"""
from __future__ import annotations

import importlib
import logging
import os
import sys
import types

_INSTALLED = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')     # oracle/build_ref.py (travels to the GPU box)
REFERENCE_ROOT = os.environ.get('VALLE_REFERENCE_ROOT') or (
    '/root/reference' if os.path.isdir('/root/reference/valle/models') else _INSTALLED)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'valle', 'models'))


def _top_k_top_p_filtering(logits, top_k: int = 0, top_p: float = 1.0,
                           filter_value: float = -float('inf'), min_tokens_to_keep: int = 1):
    """transformers 4.38.2 ``top_k_top_p_filtering`` rebuilt from the warpers it wrapped."""
    from transformers.generation.logits_process import TopKLogitsWarper, TopPLogitsWarper

    if top_k > 0:
        logits = TopKLogitsWarper(top_k=top_k, filter_value=filter_value,
                                  min_tokens_to_keep=min_tokens_to_keep)(None, logits)
    if 0 <= top_p <= 1.0:
        logits = TopPLogitsWarper(top_p=top_p, filter_value=filter_value,
                                  min_tokens_to_keep=min_tokens_to_keep)(None, logits)
    return logits


def install_stubs() -> None:
    import torch
    import torch.nn as nn

    if 'lightning' not in sys.modules:
        lightning = types.ModuleType('lightning')

        class LightningModule(nn.Module):
            def log(self, *args, **kwargs):  # no-op logger hook
                return None

        lightning.LightningModule = LightningModule
        lightning.Trainer = object
        lp = types.ModuleType('lightning.pytorch')
        lp.loggers = types.SimpleNamespace(TensorBoardLogger=object)
        lp.seed_everything = lambda seed: torch.manual_seed(seed)
        lightning.pytorch = lp
        sys.modules['lightning'] = lightning
        sys.modules['lightning.pytorch'] = lp

    if 'encodec' not in sys.modules:
        encodec = types.ModuleType('encodec')

        class EncodecModel:  # never instantiated by the hot path
            @staticmethod
            def encodec_model_24khz():
                raise RuntimeError('encodec is not available offline')

        encodec.EncodecModel = EncodecModel
        sys.modules['encodec'] = encodec

    if 'coloredlogs' not in sys.modules:
        coloredlogs = types.ModuleType('coloredlogs')
        coloredlogs.ColoredFormatter = logging.Formatter
        sys.modules['coloredlogs'] = coloredlogs

    try:
        importlib.import_module('torchaudio')
    except Exception:  # pragma: no cover - depends on the image
        ta = types.ModuleType('torchaudio')
        ta.functional = types.SimpleNamespace(resample=None)
        ta.load = None
        sys.modules['torchaudio'] = ta

    import transformers.generation.utils as tgu

    if not hasattr(tgu, 'top_k_top_p_filtering'):
        tgu.top_k_top_p_filtering = _top_k_top_p_filtering


def import_reference():
    """Return the reference's ``valle`` package (unmodified source, executed from its own tree)."""
    if not reference_available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')
    import torch

    install_stubs()
    # Our own repo also ships a `valle` alias package; make sure the reference's wins here.
    for name in [m for m in sys.modules if m == 'valle' or m.startswith('valle.')]:
        del sys.modules[name]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        prec = torch.get_float32_matmul_precision()
        valle = importlib.import_module('valle')
        importlib.import_module('valle.config')
        importlib.import_module('valle.models')
        importlib.import_module('valle.models.modules')
        importlib.import_module('valle.models.utils')
        # valle/utils.py:11 flips the global matmul precision; the CPU oracle is true fp32.
        torch.set_float32_matmul_precision(prec)
    finally:
        sys.path.remove(REFERENCE_ROOT)
    return valle


def release_reference() -> None:
    """Drop the reference's modules so the repo's own `valle` alias can be imported again."""
    for name in [m for m in sys.modules if m == 'valle' or m.startswith('valle.')]:
        del sys.modules[name]
