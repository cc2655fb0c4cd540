"""Module-level mirror of the reference's ``valle/models/modules.py`` (same class names, constructor and
``forward`` signatures, attribute names and ``state_dict`` keys), computing through libvalle_b200.so.

Parameters are ordinary fp32 ``nn.Parameter`` s, so checkpoints interchange with the reference.  The forward passes
run the CUDA kernels in the precision selected by ``valle2_b200.set_precision`` ('bf16' default, 'fp32' validation);
inputs must live on a CUDA device -- there is no CPU path.  These module-level entry points keep the reference's
materialised-mask API (``merge_masks`` etc.); the fast generation paths in ``engine.py`` never materialise a mask.
These module-level forwards are INFERENCE entry points: they do not build an autograd graph (inputs are detached; an input
that requires grad raises instead of silently yielding partial gradients).  Training goes through
``ValleAR.training_step`` / ``ValleNAR.training_step``, whose forward AND backward run on the CUDA stack (valle2_b200/train.py).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

import valle2_b200

from .. import ops
from ..config import ConfigValle


def _inference_only(x: torch.Tensor, who: str) -> None:
    if torch.is_grad_enabled() and x.requires_grad:
        raise RuntimeError(f'{who}.forward is an inference entry point and does not back-propagate into its input; '
                           'use ValleAR.training_step / ValleNAR.training_step (hand-written backward) for gradients')


def _compute_dtype() -> torch.dtype:
    return torch.bfloat16 if valle2_b200.get_precision() == 'bf16' else torch.float32


class _CastCache:
    """bf16 copies of fp32 parameters, refreshed when the parameter is modified in place or replaced."""

    def __init__(self):
        self._store = {}

    def get(self, p: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
        t = p.detach()
        if t.dtype == dtype and t.is_contiguous():
            return t
        key = (id(p), dtype)
        hit = self._store.get(key)
        if hit is None or hit[0] != p._version or hit[1].device != p.device or hit[2] != p.data_ptr():
            hit = (p._version, t.to(dtype).contiguous(), p.data_ptr())
            self._store[key] = hit
        return hit[1]


_cache = _CastCache()


def _rows(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """(..., d) tensor -> contiguous (R, d) rows in the compute dtype (cast kernel, no torch math)."""
    R = x.numel() // x.shape[-1]
    x2 = x.reshape(R, x.shape[-1])
    if x2.dtype == dtype and x2.is_contiguous():
        return x2
    x2 = x2.contiguous()
    if x2.dtype == torch.float32:
        out = torch.empty_like(x2, dtype=dtype)
        ops.residual_layernorm(x2, None, None, out)
        return out
    return x2.to(dtype)


def _linear_nd(x: torch.Tensor, weight, bias, *, gelu=False) -> torch.Tensor:
    cd = _compute_dtype()
    xr = _rows(x, cd)
    b = None if bias is None else _cache.get(bias, torch.float32)
    y = ops.linear(xr, _cache.get(weight, cd), b, gelu=gelu, out_dtype=torch.float32)
    return y.view(*x.shape[:-1], weight.shape[0]).to(x.dtype)


class TokenEmbedding(nn.Module):
    """modules.py:11-37 -- table gather (dropout is identity at p=0 / eval)."""

    def __init__(self, vocab_size: int, dim_model: int, dropout: float = 0.0):
        super().__init__()
        self.vocab_size = vocab_size
        self.dim_model = dim_model
        self.dropout = nn.Dropout(p=dropout)
        self.word_embeddings = nn.Embedding(vocab_size, dim_model)

    @property
    def weight(self) -> torch.Tensor:
        return self.word_embeddings.weight

    def embedding(self, index: int) -> torch.Tensor:
        return self.word_embeddings.weight[index: index + 1]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        w = self.word_embeddings.weight
        ids = x.reshape(-1, 1, 1).to(torch.int32).contiguous()
        out = torch.empty(ids.shape[0], self.dim_model, device=w.device, dtype=torch.float32)
        zero_pe = torch.zeros(1, self.dim_model, device=w.device, dtype=torch.float32)
        ops.embed_sum_pe(ids, w.detach().float().unsqueeze(0).contiguous(), zero_pe, out, out_rows_per_batch=1)
        return self.dropout(out.view(*x.shape, self.dim_model))


class PositionalEncoding(nn.Module):
    """modules.py:40-80 -- fp32 sinusoidal table in the ``pe`` buffer, shape (max_len, 1, d_model)."""

    def __init__(self, d_model, dropout=0.1, max_len=5000):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        freq = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        table = torch.zeros(max_len, d_model)
        table[:, 0::2] = torch.sin(pos * freq)
        table[:, 1::2] = torch.cos(pos * freq)
        self.register_buffer('pe', table.unsqueeze(1))

    def forward(self, x):
        # x: (batch, seq, d).  A broadcast add of a constant table (memory-bound plumbing at module level; the
        # generation engines fuse it into the embedding kernel).
        return self.dropout(x + self.pe[: x.size(1), 0].unsqueeze(0).to(x.dtype))


class AdaptiveLayerNorm(nn.Module):
    """modules.py:83-99 -- ``w * LayerNorm(x) + b`` with ``[w, b] = Linear(d -> 2d)(embedding)``."""

    def __init__(self, d_model) -> None:
        super().__init__()
        self.project_layer = nn.Linear(d_model, 2 * d_model)
        self.norm = nn.LayerNorm(d_model)
        self.d_model = d_model
        self.eps = self.norm.eps

    def forward(self, x: torch.Tensor, embedding: torch.Tensor) -> torch.Tensor:
        d = self.d_model
        _inference_only(x, type(self).__name__)
        e = embedding.detach().float().reshape(-1, d).contiguous()
        assert e.shape[0] == 1, 'stage embedding must be (1, d_model) (valle_nar.py:35,95,153)'
        wb = ops.linear(e, _cache.get(self.project_layer.weight, torch.float32),
                        _cache.get(self.project_layer.bias, torch.float32))
        w, b = wb[0, :d], wb[0, d:]
        gamma = (w * self.norm.weight.detach()).contiguous()
        beta = (w * self.norm.bias.detach() + b).contiguous()
        xr = x.detach().float().reshape(-1, d).contiguous()
        y = torch.empty_like(xr)
        ops.residual_layernorm(xr, gamma, beta, y, eps=self.eps)
        return y.view(x.shape).to(x.dtype)


def _layer_norm(norm: nn.LayerNorm, x: torch.Tensor) -> torch.Tensor:
    d = x.shape[-1]
    xr = x.detach().float().reshape(-1, d).contiguous()
    y = torch.empty_like(xr)
    ops.residual_layernorm(xr, _cache.get(norm.weight, torch.float32), _cache.get(norm.bias, torch.float32), y,
                           eps=norm.eps)
    return y.view(x.shape).to(x.dtype)


class MultiHeadAttention(nn.Module):
    """modules.py:102-207."""

    def __init__(self, d_model: int, n_heads: int) -> None:
        super().__init__()
        assert d_model % n_heads == 0, 'd_model should be divisible by n_heads'
        self.d_model = d_model
        self.n_heads = n_heads
        self.head_dim = d_model // n_heads
        self.qkv = nn.Linear(d_model, 3 * d_model, bias=False)
        self.out = nn.Linear(d_model, d_model)

    def forward(self, x: torch.Tensor, *, attn_mask: torch.Tensor | None = None,
                padding_mask: torch.Tensor | None = None, kv_cache=None, use_cache: bool = False):
        B, n, d = x.shape
        H, Dh = self.n_heads, self.head_dim
        cd = _compute_dtype()
        _inference_only(x, type(self).__name__)
        xr = _rows(x.detach(), cd)
        qkv = ops.linear(xr, _cache.get(self.qkv.weight, cd)).view(B, n, 3, H, Dh)
        q = qkv[:, :, 0].permute(0, 2, 1, 3)
        k = qkv[:, :, 1].permute(0, 2, 1, 3)
        v = qkv[:, :, 2].permute(0, 2, 1, 3)
        kv = None
        if use_cache and kv_cache is not None:
            # module-level API keeps the reference's dense (B,H,T,Dh) cache tuple; the engines use the paged pool
            k = torch.cat([kv_cache[0].to(cd), k], dim=-2)
            v = torch.cat([kv_cache[1].to(cd), v], dim=-2)
        if use_cache:
            kv = (k.to(x.dtype).contiguous(), v.to(x.dtype).contiguous())
        o = torch.empty(B, n, d, device=x.device, dtype=cd)
        if attn_mask is not None:
            merged = self.merge_masks(B, attn_mask, padding_mask)
            assert merged is not None, 'attn_mask should not be None'
            mask_u8 = merged.to(torch.bool).to(torch.uint8).contiguous()
            ops.attention(q, k, v, o, mask_mode=ops.MASK_EXPLICIT, mask=mask_u8)
        else:
            ops.attention(q, k, v, o, mask_mode=ops.MASK_NONE)      # padding_mask alone is ignored (:160)
        out = ops.linear(o.view(B * n, d), _cache.get(self.out.weight, cd), _cache.get(self.out.bias, torch.float32),
                         out_dtype=torch.float32)
        return out.view(B, n, d).to(x.dtype), kv

    def merge_masks(self, batch_size: int, attn_mask: torch.Tensor | None,
                    key_padding_mask: torch.Tensor | None) -> torch.Tensor | None:
        """(seq,seq) or (B,seq,seq) mask (+ optional (B,seq) key padding mask) -> materialised (B,H,seq,seq)."""
        if attn_mask is None:
            return None
        if attn_mask.dim() == 3:
            merged = attn_mask[:, None]
        else:
            merged = attn_mask[None, None].expand(batch_size, self.n_heads, -1, -1).contiguous()
        if key_padding_mask is not None:
            kpm = key_padding_mask[:, None, None, :].expand(batch_size, self.n_heads, 1, -1)
            merged = merged + kpm
        return merged


class FeedForward(nn.Module):
    """modules.py:210-221 -- always erf-GELU (reference quirk K-1)."""

    def __init__(self, d_model: int, d_ff: int, dropout: float = 0.1) -> None:
        super().__init__()
        self.linear_1 = nn.Linear(d_model, d_ff)
        self.activation = nn.GELU()
        self.dropout = nn.Dropout(dropout)
        self.linear_2 = nn.Linear(d_ff, d_model)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        cd = _compute_dtype()
        _inference_only(x, type(self).__name__)
        xr = _rows(x.detach(), cd)
        h = ops.linear(xr, _cache.get(self.linear_1.weight, cd), _cache.get(self.linear_1.bias, torch.float32), gelu=True)
        y = ops.linear(h, _cache.get(self.linear_2.weight, cd), _cache.get(self.linear_2.bias, torch.float32),
                       out_dtype=torch.float32)
        return y.view(x.shape).to(x.dtype)


class EncoderLayer(nn.Module):
    """modules.py:224-294 -- pre-norm block."""

    def __init__(self, config: ConfigValle) -> None:
        super().__init__()
        self.config = config
        self.self_attn = MultiHeadAttention(config.d_model, config.n_heads)
        self.ffn = FeedForward(config.d_model, config.dim_feedforward, dropout=config.dropout)
        self.norm1 = self._get_norm()(config.d_model)
        self.norm2 = self._get_norm()(config.d_model)
        self.dropout1 = nn.Dropout(config.dropout)
        self.dropout2 = nn.Dropout(config.dropout)
        self.activation = self._get_activation()()

    def _apply_norm(self, norm, x, embedding):
        if self.config.norm == 'LayerNorm':
            return _layer_norm(norm, x)
        return norm(x, embedding=embedding)

    def forward(self, x: torch.Tensor, *, padding_mask=None, attn_mask=None, embedding=None, kv_cache=None,
                use_cache: bool = False):
        a, next_kv = self.self_attn(self._apply_norm(self.norm1, x, embedding), attn_mask=attn_mask,
                                    padding_mask=padding_mask, kv_cache=kv_cache, use_cache=use_cache)
        x = x + self.dropout1(a)
        x = x + self.dropout2(self.ffn(self._apply_norm(self.norm2, x, embedding)))
        return x, next_kv

    def _get_norm(self):
        return {'LayerNorm': nn.LayerNorm, 'AdaptiveLayerNorm': AdaptiveLayerNorm}[self.config.norm]

    def _get_activation(self):
        return {'relu': nn.ReLU, 'gelu': nn.GELU}[self.config.activation]


class Transformer(nn.Module):
    """modules.py:297-352 -- layer stack; with a cache only the last position is processed."""

    def __init__(self, hparams: ConfigValle) -> None:
        super().__init__()
        self.hparams = hparams
        self.layers = nn.ModuleList([EncoderLayer(hparams) for _ in range(hparams.num_layers)])

    def forward(self, x: torch.Tensor, *, padding_mask=None, attn_mask=None, embedding=None,
                kv_cache: tuple | None = None, use_cache: bool = False):
        new_kv: tuple = ()
        if use_cache and kv_cache is not None:
            x = x[:, -1:]
            attn_mask = None
        else:
            kv_cache = tuple([None] * self.hparams.num_layers)
        for layer, past in zip(self.layers, kv_cache):
            x, kv = layer(x, padding_mask=padding_mask, attn_mask=attn_mask, embedding=embedding, kv_cache=past,
                          use_cache=use_cache)
            if use_cache:
                new_kv = new_kv + (kv,)
        return x, new_kv
