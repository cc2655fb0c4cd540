"""Batch-dict contract of the training path: the step BEFORE ``training_step``.

Mirrors ``valle/collate.py`` of the reference: ``get_collate(model_name)`` returns a dataclass whose ``__call__`` turns a
list of dataset items ``{'codes': (Q, T) int64, 'tokens': (Tx,) int64}`` into the padded batch dict that
``ValleAR.training_step`` (valle_ar.py:43-60) / ``ValleNAR.training_step`` (valle_nar.py:53-67) read.

* ``ValleARCollate`` (collate.py:19-46): first codebook only; ``codes`` = BOS + codes, ``target`` = codes + EOS (the shift by
  one), zero padding (0 is a valid class: the loss averages over it, K-5), lengths int64, and the reference's assertion
  that every clip has more audio frames than phonemes.
* ``ValleNARCollate`` (collate.py:49-60, REPAIRED per SURVEY A-13): upstream pads the ``(Q, T)`` item tensors along dim 0 and
  takes ``len() == Q`` as the length, which raises for ragged ``T``.  Here items are transposed to ``(T, Q)`` first, so the
  batch is ``codes (B, T_max, Q)`` with ``codes_lens`` = frames -- the layout ``ValleNAR.training_step`` /
  ``_prepare_audio_codes`` index (valle_nar.py:81, :177-186).
Pure host logic (CPU tensors in, CPU tensors out), as upstream.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import Tensor

from .config import ConfigValle


def collate_list(x_list: list[Tensor]) -> tuple[Tensor, Tensor]:
    """Zero-pad a list of tensors along their first dim -> (padded (B, L_max, ...), lengths (B,) int64)  (collate.py:63-66)."""
    lens = torch.tensor([int(x.shape[0]) for x in x_list], dtype=torch.int64)
    l_max = int(lens.max()) if len(x_list) else 0
    out = x_list[0].new_zeros((len(x_list), l_max) + tuple(x_list[0].shape[1:])) if x_list else torch.zeros(0, 0)
    for i, x in enumerate(x_list):
        out[i, : x.shape[0]] = x
    return out, lens


@dataclass
class ValleARCollate:
    config: ConfigValle

    def __call__(self, batch: list[dict[str, Tensor]]) -> dict[str, Tensor]:
        bos, eos = self.config.bos_token, self.config.eos_token
        first = [item['codes'][0] for item in batch]                       # first codebook only
        codes, codes_lens = collate_list([torch.cat([c.new_full((1,), bos), c]) for c in first])
        target, _ = collate_list([torch.cat([c, c.new_full((1,), eos)]) for c in first])
        tokens, tokens_lens = collate_list([item['tokens'] for item in batch])
        assert (codes_lens > tokens_lens).all(), 'Codes length must be greater than tokens length.'
        return {'codes': codes, 'codes_lens': codes_lens, 'target': target, 'tokens': tokens, 'tokens_lens': tokens_lens}


@dataclass
class ValleNARCollate:
    config: ConfigValle

    def __call__(self, batch: list[dict[str, Tensor]]) -> dict[str, Tensor]:
        codes, codes_lens = collate_list([item['codes'].transpose(0, 1).contiguous() for item in batch])   # (T, Q) per item
        tokens, tokens_lens = collate_list([item['tokens'] for item in batch])
        assert (codes_lens > tokens_lens).all(), 'Codes length must be greater than tokens length.'
        return {'codes': codes, 'codes_lens': codes_lens, 'tokens': tokens, 'tokens_lens': tokens_lens}


def get_collate(model_name: str):
    return {'ValleAR': ValleARCollate, 'ValleNAR': ValleNARCollate}[model_name]
