"""ValleAR -- first-codebook autoregressive decoder.  Mirror of the reference's ``valle/models/valle_ar.py``:
same constructor, sub-module names (=> state_dict keys), ``training_step`` / ``generate`` /
``configure_optimizers`` signatures.  ``generate`` runs the batched CUDA engine (``engine.ARDecoder``): one prefill
over [text | BOS + prompt codes] with the prefix-LM mask evaluated in-kernel, then KV-cached decode steps replayed
from a CUDA graph with device-side sampling and EOS bookkeeping."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import optim

import valle2_b200

from .. import ops
from ..config import ConfigValle
from ..engine import ARDecoder, _i32
from ._base import BaseModule
from .modules import PositionalEncoding, TokenEmbedding, Transformer
from .utils import get_best_beam


class ValleAR(BaseModule):
    def __init__(self, config: ConfigValle):
        super().__init__()
        self.config = config
        self.tokens_emb = TokenEmbedding(config.vocab_size, config.d_model)
        self.audio_emb = TokenEmbedding(config.num_audio_tokens + 2, config.d_model)   # + EOS, BOS
        self.tokens_position_emb = PositionalEncoding(config.d_model)
        self.audio_position_emb = PositionalEncoding(config.d_model)
        self.transformer = Transformer(config)
        self.proj = nn.Linear(config.d_model, config.num_audio_tokens + 1, bias=False)
        self._engine_cache = None

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def eos_token(self):
        return self.config.num_audio_tokens

    @property
    def bos_token(self):
        return self.config.num_audio_tokens + 1

    # -- engine management ----------------------------------------------------------------------
    def _engine(self) -> ARDecoder:
        precision = valle2_b200.get_precision()
        stamp = (precision, str(self.device), tuple(p._version for p in self.parameters()),
                 tuple(p.data_ptr() for p in self.parameters()))
        if self._engine_cache is None or self._engine_cache[0] != stamp:
            self._engine_cache = (stamp, ARDecoder(self, precision))
        return self._engine_cache[1]

    # -- teacher-forced forward (valle_ar.py:43-90) ------------------------------------------------
    @torch.no_grad()
    def forward_logits(self, batch: dict[str, torch.Tensor]) -> torch.Tensor:
        """Teacher-forced logits (B, Ty, V+1) through the CUDA stack; layout and masks as valle_ar.py:61-83:
        text padded to max(tokens_lens) then audio; prefix-LM mask + key padding on the audio part only (K-4)."""
        eng = self._engine()
        dev = self.device
        tokens, codes = batch['tokens'].to(dev), batch['codes'].to(dev)
        tokens_lens, codes_lens = batch['tokens_lens'], batch['codes_lens']
        B = tokens.shape[0]
        Tx, Ty = int(tokens_lens.max()), int(codes_lens.max())
        tokens, codes = tokens[:, :Tx], codes[:, :Ty]
        S = Tx + Ty
        d = self.config.d_model
        x = torch.empty(B * S, d, device=dev, dtype=torch.float32)
        ops.embed_sum_pe(_i32(tokens, dev).view(B, Tx, 1), eng.tok_table, eng.pe_t, x, out_rows_per_batch=S)
        ops.embed_sum_pe(_i32(codes, dev).view(B, Ty, 1), eng.aud_table, eng.pe_a, x, out_rows_per_batch=S,
                         out_row_offset=Tx)
        xl = torch.full((B,), Tx, device=dev, dtype=torch.int32)
        kv_lens = (xl + _i32(codes_lens, dev)).contiguous()
        eng.runner.forward(x, B, S, mask_mode=ops.MASK_PREFIX_LM, x_lens=xl, kv_lens=kv_lens)
        rows = x.view(B, S, d)[:, Tx:].reshape(B * Ty, d)
        if eng.precision == 'bf16':
            hb = torch.empty(B * Ty, d, device=dev, dtype=torch.bfloat16)
            ops.residual_layernorm(rows, None, None, hb)
            logits = ops.linear(hb, eng.wproj, out_dtype=torch.float32)
        else:
            logits = ops.linear(rows, eng.wproj)
        return logits.view(B, Ty, -1)

    def training_step(self, batch: dict[str, torch.Tensor], **kwargs) -> torch.Tensor:
        """valle_ar.py:43-90: mean CE over all positions incl. padding (K-5).  Forward AND backward run on the CUDA stack
        (valle2_b200/train.py); the returned loss is attached to the parameters through one autograd Function, so
        ``loss.backward()`` / Lightning's optimisation loop work as with the reference."""
        from .. import train
        precision = valle2_b200.get_precision()
        drop = train.dropout_plan(self, kwargs.get('dropout_seed'))       # None in eval mode / at rate 0
        loss = train.step_loss(self, lambda: train.ar_loss_and_grads(self, batch, precision, drop))
        self.log('train/loss', loss)
        return loss

    # -- generation ---------------------------------------------------------------------------------
    @torch.inference_mode()
    def generate(self, prompt_tokens: torch.Tensor, prompt_codes: torch.Tensor,
                 target_tokens: torch.Tensor | None = None, *, uniforms: torch.Tensor | None = None,
                 seed: int | None = None, use_graph: bool = True) -> torch.Tensor:
        """First-codebook codes for one utterance, ``num_beams`` independent samples, best beam returned
        (valle_ar.py:92-180).  Extra keyword-only arguments (not in the reference): ``uniforms``
        (max_audio_len, num_beams) injected draws, ``seed`` for the in-kernel generator."""
        assert prompt_tokens.dim() == 1, 'Prompt tokens should be 1D tensor.'
        assert prompt_codes.dim() == 2, 'Prompt codes should be 2D tensor.'
        if target_tokens is not None:
            assert target_tokens.dim() == 1, 'Target tokens should be 1D tensor.'
        cfg, dev = self.config, self.device
        tokens = prompt_tokens if target_tokens is None else torch.cat((prompt_tokens, target_tokens), dim=0)
        first = prompt_codes[..., 0].to(dev)
        codes = torch.cat([torch.full((1,), self.bos_token, device=dev, dtype=first.dtype), first])
        prompt_len = codes.shape[0]
        nb = cfg.num_beams
        if seed is None:
            seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
        out, sum_logprobs, n = self._engine().generate(
            tokens.to(dev).unsqueeze(0).repeat(nb, 1), codes.unsqueeze(0).repeat(nb, 1), max_new=cfg.max_audio_len,
            top_k=cfg.top_k, top_p=cfg.tok_p, temperature=cfg.temperature, uniforms=uniforms, seed=seed,
            use_graph=use_graph)
        beams = torch.cat([codes.unsqueeze(0).repeat(nb, 1).long(), out.long()], dim=1)
        best = get_best_beam(beams, sum_logprobs, self.eos_token, cfg.length_penalty)
        best = best[prompt_len:]
        return best[best != self.eos_token]

    @torch.inference_mode()
    def generate_batch(self, tokens: torch.Tensor, codes: torch.Tensor, *, code_lens: torch.Tensor | None = None,
                       max_new: int | None = None, ignore_eos: bool = False, seed: int = 0, use_graph: bool = True):
        """Extension (no upstream equivalent): decode B utterances at once, one beam each.
        tokens (B,Tx) phonemes, codes (B,P) = BOS + first-codebook prompt.  Returns (codes_out (B,n), n)."""
        cfg = self.config
        out, _, n = self._engine().generate(tokens, codes, code_lens=code_lens, max_new=max_new or cfg.max_audio_len,
                                            top_k=cfg.top_k, top_p=cfg.tok_p, temperature=cfg.temperature, seed=seed,
                                            ignore_eos=ignore_eos, use_graph=use_graph)
        return out, n

    def configure_optimizers(self):
        optimizer = optim.AdamW(self.parameters(), lr=self.config.lr, betas=self.config.betas,
                                weight_decay=self.config.weight_decay, fused=True)
        scheduler = optim.lr_scheduler.CosineAnnealingWarmRestarts(optimizer, self.config.lr_warmup)
        return {'optimizer': optimizer, 'lr_scheduler': scheduler}
