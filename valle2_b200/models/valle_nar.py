"""ValleNAR -- codebooks 2..Q, full attention, stage-conditioned AdaptiveLayerNorm, summed codebook embeddings.
Mirror of the reference's ``valle/models/valle_nar.py`` (constructor, sub-module names => state_dict keys, method
signatures).  Upstream ``generate`` / ``training_step`` raise; the behaviour implemented here is the repaired one
documented in SURVEY Appendix A (A-1..A-9) and restated in ``oracle/valle_oracle.py``."""
from __future__ import annotations

import random

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import optim

import valle2_b200

from .. import ops
from ..config import ConfigValle
from ..engine import NARDecoder, _i32
from ._base import BaseModule
from .modules import PositionalEncoding, TokenEmbedding, Transformer


class ValleNAR(BaseModule):
    def __init__(self, config: ConfigValle):
        super().__init__()
        self.config = config
        self.eos_token = config.num_audio_tokens
        self.bos_token = config.num_audio_tokens + 1
        q = config.num_quantizers
        self.tokens_emb = TokenEmbedding(config.vocab_size, config.d_model)
        self.codes_embs = nn.ModuleList([TokenEmbedding(config.num_audio_tokens, config.d_model) for _ in range(q)])
        self.tokens_position_emb = PositionalEncoding(config.d_model)
        self.audio_position_emb = PositionalEncoding(config.d_model)
        self.stage_embs = nn.ModuleList([TokenEmbedding(1, config.d_model) for _ in range(q - 1)])
        self.transformer = Transformer(config)
        self.proj_layers = nn.ModuleList(
            [nn.Linear(config.d_model, config.num_audio_tokens, bias=False) for _ in range(q - 1)])
        self._engine_cache = None

    @property
    def device(self):
        return next(self.parameters()).device

    def _engine(self) -> NARDecoder:
        precision = valle2_b200.get_precision()
        stamp = (precision, str(self.device), tuple(p._version for p in self.parameters()),
                 tuple(p.data_ptr() for p in self.parameters()))
        if self._engine_cache is None or self._engine_cache[0] != stamp:
            self._engine_cache = (stamp, NARDecoder(self, precision))
        return self._engine_cache[1]

    # -- teacher-forced stage (valle_nar.py:53-105 with repairs A-1..A-3, keep A-4) -------------------
    @torch.no_grad()
    def forward_logits(self, batch: dict[str, torch.Tensor], layer: int):
        eng = self._engine()
        dev, cfg = self.device, self.config
        codes, tokens = batch['codes'].to(dev), batch['tokens'].to(dev)
        B, T, Q = codes.shape
        Tx = int(batch['tokens_lens'].max())
        tokens = tokens[:, :Tx]
        prefix_len = min(T // 3, 3 * cfg.quantization_factor)
        S, d = Tx + T, cfg.d_model
        x = torch.empty(B * S, d, device=dev, dtype=torch.float32)
        ops.embed_sum_pe(_i32(tokens, dev).view(B, Tx, 1), eng.tok_table, eng.pe_t, x, out_rows_per_batch=S)
        ops.embed_sum_pe(_i32(codes, dev), eng.code_tables, eng.pe_a, x, t_split=prefix_len, nq_a=Q, nq_b=layer,
                         out_rows_per_batch=S, out_row_offset=Tx)
        eng.runner.forward(x, B, S, mask_mode=ops.MASK_NONE, stage=layer - 1)     # padding ignored (A-4)
        rows = x.view(B, S, d)[:, Tx + prefix_len:].reshape(-1, d)
        if eng.precision == 'bf16':
            hb = torch.empty(rows.shape[0], d, device=dev, dtype=torch.bfloat16)
            ops.residual_layernorm(rows, None, None, hb)
            logits = ops.linear(hb, eng.wproj[layer - 1], out_dtype=torch.float32)
        else:
            logits = ops.linear(rows, eng.wproj[layer - 1])
        return logits.view(B, T - prefix_len, -1), prefix_len

    def training_step(self, batch: dict[str, torch.Tensor], **kwargs) -> torch.Tensor:
        """valle_nar.py:53-105 (repairs A-1..A-3, A-4 kept): random stage (:76), CE on that codebook.  Forward and backward
        on the CUDA stack (valle2_b200/train.py); ``loss.backward()`` fills the parameter gradients."""
        from .. import train
        layer = kwargs.get('layer') or random.randint(1, self.config.num_quantizers - 1)   # :76
        precision = valle2_b200.get_precision()
        drop = train.dropout_plan(self, kwargs.get('dropout_seed'))
        return train.step_loss(self, lambda: train.nar_loss_and_grads(self, batch, layer, precision, drop))

    @torch.inference_mode()
    def generate(self, prompt_tokens: torch.Tensor, prompt_codes: torch.Tensor, target_tokens: torch.Tensor,
                 target_codes_first_layer: torch.Tensor, *, greedy: bool = False, seed: int = 0) -> torch.Tensor:
        """Remaining codebooks for one utterance -> (output_len, quantization_layers) int64 (valle_nar.py:107-165).
        ``greedy=True`` takes the arg-max instead of the reference's Categorical draw (A-9)."""
        out = self._engine().generate(prompt_tokens.unsqueeze(0), prompt_codes.unsqueeze(0), target_tokens.unsqueeze(0),
                                      target_codes_first_layer.unsqueeze(0), greedy=greedy,
                                      temperature=self.config.temperature, seed=seed)
        return out[0]

    @torch.inference_mode()
    def generate_batch(self, prompt_tokens, prompt_codes, target_tokens, first_layer, *, greedy: bool = True,
                       seed: int = 0, use_tc_attention: bool | None = None, target_lens=None) -> torch.Tensor:
        """Extension: B utterances at once -> (B, T, Q); ragged targets via ``target_lens`` (B,)."""
        return self._engine().generate(prompt_tokens, prompt_codes, target_tokens, first_layer, greedy=greedy,
                                       temperature=self.config.temperature, seed=seed,
                                       use_tc_attention=use_tc_attention, target_lens=target_lens)

    def configure_optimizers(self):
        """Upstream ``ValleNAR`` has no ``configure_optimizers`` (SURVEY A-11: ``-m ValleNAR`` cannot train there); this is
        ``ValleAR.configure_optimizers`` (valle_ar.py:182-194) verbatim in behaviour: AdamW + CosineAnnealingWarmRestarts."""
        optimizer = optim.AdamW(self.parameters(), lr=self.config.lr, betas=self.config.betas,
                                weight_decay=self.config.weight_decay, fused=True)
        scheduler = optim.lr_scheduler.CosineAnnealingWarmRestarts(optimizer, self.config.lr_warmup)
        return {'optimizer': optimizer, 'lr_scheduler': scheduler}

    def _prepare_audio_codes(self, codes: torch.Tensor, nar_stage: int) -> tuple[torch.Tensor, int]:
        """(B, T, Q) codes -> (summed embeddings (B, T, d), prefix_len)  (valle_nar.py:167-188): the first
        ``prefix_len = min(T//3, 3*quantization_factor)`` frames sum all Q codebooks, the rest codebooks < nar_stage."""
        eng = self._engine()
        B, T, Q = codes.shape
        prefix_len = min(T // 3, 3 * self.config.quantization_factor)
        d = self.config.d_model
        out = torch.empty(B * T, d, device=self.device, dtype=torch.float32)
        zero_pe = torch.zeros(1, d, device=self.device, dtype=torch.float32)
        ops.embed_sum_pe(_i32(codes, self.device), eng.code_tables, zero_pe, out, t_split=prefix_len, nq_a=Q,
                         nq_b=nar_stage, out_rows_per_batch=T)
        return out.view(B, T, d), prefix_len
