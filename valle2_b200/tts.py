"""End-to-end codec-token synthesis for a batch of utterances (BASELINE config 4): AR first codebook -> NAR codebooks
2..Q.  This is the hand-off the reference leaves to the caller (valle_ar.py:173-180 returns (T,) codes,
valle_nar.py:107-130 consumes them as ``target_codes_first_layer``); EnCodec decoding of the codes is out of scope."""
from __future__ import annotations

import torch


@torch.inference_mode()
def synthesize_batch(ar, nar, prompt_tokens: torch.Tensor, prompt_codes: torch.Tensor, target_tokens: torch.Tensor, *,
                     max_new: int | None = None, greedy_nar: bool = True, seed: int = 0, ignore_eos: bool = False,
                     nar_chunk: int | None = None):
    """prompt_tokens (B,Tp), prompt_codes (B,Tc,Q), target_tokens (B,Tt) -> list of (T_b, Q) int64 code matrices.

    AR: one beam per utterance (``ValleAR.generate_batch``); an utterance ends at its first EOS.  NAR: all utterances
    in one batch (or in chunks of ``nar_chunk`` utterances: the stages are tensor-core bound from a few dozen sequences on, so
    chunking costs nothing and bounds the activation workspaces), ragged lengths handled by key masking (``target_lens``)."""
    dev = ar.device
    B, Tc, Q = prompt_codes.shape
    eos, bos = ar.eos_token, ar.bos_token
    tokens = torch.cat([prompt_tokens, target_tokens], dim=1).to(dev)
    codes = torch.cat([torch.full((B, 1), bos, device=dev, dtype=torch.long), prompt_codes[:, :, 0].to(dev).long()], dim=1)
    out, n = ar.generate_batch(tokens, codes, max_new=max_new, seed=seed, ignore_eos=ignore_eos)
    out = out.long()
    is_eos = out == eos
    any_eos = is_eos.any(dim=1)
    first = torch.where(any_eos, is_eos.float().argmax(dim=1), torch.full((B,), n, device=dev))
    if ignore_eos:              # fixed-length (benchmark) mode: an EOS id drawn on the way is just another code, all n frames count
        first = torch.full((B,), n, device=dev)
    lens = first.clamp(min=1)                      # NAR needs at least one frame per utterance
    T = int(lens.max())
    first_layer = out[:, :T].clone()
    first_layer[torch.arange(T, device=dev)[None, :] >= lens[:, None]] = 0
    first_layer.clamp_(max=nar.config.num_audio_tokens - 1)
    pt, pc, tt = prompt_tokens.to(dev), prompt_codes.to(dev), target_tokens.to(dev)
    step = B if not nar_chunk else max(1, int(nar_chunk))
    parts = []
    for b0 in range(0, B, step):
        sl = slice(b0, min(B, b0 + step))
        Tc_ = int(lens[sl].max())
        parts.append(nar.generate_batch(pt[sl], pc[sl], tt[sl], first_layer[sl, :Tc_].contiguous(), greedy=greedy_nar, seed=seed,
                                        target_lens=lens[sl]))
    if ignore_eos and all(p.shape[1] == n for p in parts):      # fixed-length job: no per-utterance host reads
        full = torch.cat(parts, 0)
        return [full[b] for b in range(B)]
    first_h = first.tolist()
    out_list = []
    for ci, part in enumerate(parts):
        for j in range(part.shape[0]):
            out_list.append(part[j, : int(first_h[ci * step + j])])
    return out_list
