"""Masks, sampling and beam selection -- mirror of the reference's ``valle/models/utils.py``.

``build_pad_mask`` / ``build_attn_mask`` are host-side index logic (they exist for API parity and for the
module-level materialised-mask path; the engines evaluate the same predicates inside the attention kernels from
``(x_len, kv_len)`` scalars).  ``topk_sampling`` runs the fused sampling kernel (csrc/sample.cu).
"""
from __future__ import annotations

import torch
from torch import Tensor

from .. import ops


def build_pad_mask(lens: Tensor, device) -> Tensor:
    """True marks padding: position >= length.  Shape (len(lens), max(lens)).  (utils.py:8-14)"""
    longest = int(lens.max().item())
    steps = torch.arange(longest, device=device)
    return steps[None, :] >= lens.to(steps.device)[:, None]


def build_attn_mask(x_len: int, y_len: int, device) -> Tensor:
    """Prefix-LM mask of ValleAR, True = masked, shape (x_len+y_len, x_len+y_len)  (utils.py:17-43):
    text rows see all text and no audio; audio rows see all text and audio up to themselves."""
    n = x_len + y_len
    row = torch.arange(n, device=device)[:, None]
    col = torch.arange(n, device=device)[None, :]
    return (col >= x_len) & ((row < x_len) | (col > row))


def topk_sampling(logits: Tensor, top_k: int = 50, tok_p: float = 1.0, temperature: float | None = 1.0,
                  uniforms: Tensor | None = None):
    """temperature -> top-k -> top-p -> draw -> log-prob of the draw under the filtered distribution
    (utils.py:46-68 with transformers 4.38.2 ``top_k_top_p_filtering``).  Returns ((B,1) int64, (B,) float).

    The draw is an inverse-CDF lookup with one uniform per row (``uniforms`` or ``torch.rand``), not
    ``torch.multinomial``: same distribution, different random stream.  ``top_k == 1`` is greedy (lowest index on ties).
    """
    assert logits.dim() == 2
    B, V = logits.shape
    lg = logits.detach().float().contiguous()
    if uniforms is None:
        uniforms = torch.rand(B, device=lg.device, dtype=torch.float32)
    tok = torch.empty(B, device=lg.device, dtype=torch.int32)
    logprob = torch.empty(B, device=lg.device, dtype=torch.float32)
    ops.sample(lg, 1, 0, V, B, V, temperature=1.0 if temperature is None else temperature, top_k=top_k, top_p=tok_p,
               out_tok=tok, out_logprob=logprob, uniforms=uniforms.float().contiguous())
    return tok.long().unsqueeze(1), logprob


def get_best_beam(x: Tensor, sum_logprobs: Tensor, stop_token: int, length_penalty: float = 1.0) -> Tensor:
    """Beam with the best length-normalised log-probability, stop tokens stripped  (utils.py:71-88).
    ``length`` counts every non-stop entry of the row (BOS and prompt included), as the reference does."""
    keep = x != stop_token
    score = sum_logprobs / keep.sum(dim=-1) ** length_penalty
    winner = int(torch.argmax(score))
    return x[winner][keep[winner]]
