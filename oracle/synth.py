"""TEST INFRASTRUCTURE ONLY -- deterministic synthetic weights and inputs.

Weights are generated per tensor from ``(seed, crc32(key))`` so that the reference model, the
oracle and the CUDA engine can all be loaded with bit-identical parameters without storing
megabytes of fixtures.  Scales follow torch's default initialisers (Linear: U(+-1/sqrt(fan_in)),
Embedding: N(0,1)); LayerNorm affine parameters are perturbed away from (1, 0) so that the affine
and the AdaLN fold are actually exercised.
"""
from __future__ import annotations

import math
import zlib

import torch

from .valle_oracle import OracleConfig, sinusoidal_pe


def _layer_shapes(cfg: OracleConfig, p: str) -> dict:
    d, f = cfg.d_model, cfg.dim_feedforward
    s = {
        p + 'self_attn.qkv.weight': (3 * d, d),
        p + 'self_attn.out.weight': (d, d),
        p + 'self_attn.out.bias': (d,),
        p + 'ffn.linear_1.weight': (f, d),
        p + 'ffn.linear_1.bias': (f,),
        p + 'ffn.linear_2.weight': (d, f),
        p + 'ffn.linear_2.bias': (d,),
    }
    for n in ('norm1.', 'norm2.'):
        if cfg.norm == 'LayerNorm':
            s[p + n + 'weight'] = (d,)
            s[p + n + 'bias'] = (d,)
        else:
            s[p + n + 'project_layer.weight'] = (2 * d, d)
            s[p + n + 'project_layer.bias'] = (2 * d,)
            s[p + n + 'norm.weight'] = (d,)
            s[p + n + 'norm.bias'] = (d,)
    return s


def ar_state_shapes(cfg) -> dict:
    """Key inventory of ``ValleAR`` (valle_ar.py:15-29)."""
    cfg = OracleConfig.from_any(cfg)
    d = cfg.d_model
    s = {
        'tokens_emb.word_embeddings.weight': (cfg.vocab_size, d),
        'audio_emb.word_embeddings.weight': (cfg.num_audio_tokens + 2, d),
        'tokens_position_emb.pe': (5000, 1, d),
        'audio_position_emb.pe': (5000, 1, d),
    }
    for i in range(cfg.num_layers):
        s.update(_layer_shapes(cfg, f'transformer.layers.{i}.'))
    s['proj.weight'] = (cfg.num_audio_tokens + 1, d)
    return s


def nar_state_shapes(cfg) -> dict:
    """Key inventory of ``ValleNAR`` (valle_nar.py:17-47)."""
    cfg = OracleConfig.from_any(cfg)
    d = cfg.d_model
    s = {'tokens_emb.word_embeddings.weight': (cfg.vocab_size, d)}
    for j in range(cfg.num_quantizers):
        s[f'codes_embs.{j}.word_embeddings.weight'] = (cfg.num_audio_tokens, d)
    s['tokens_position_emb.pe'] = (5000, 1, d)
    s['audio_position_emb.pe'] = (5000, 1, d)
    for j in range(cfg.num_quantizers - 1):
        s[f'stage_embs.{j}.word_embeddings.weight'] = (1, d)
    for i in range(cfg.num_layers):
        s.update(_layer_shapes(cfg, f'transformer.layers.{i}.'))
    for j in range(cfg.num_quantizers - 1):
        s[f'proj_layers.{j}.weight'] = (cfg.num_audio_tokens, d)
    return s


def synth_tensor(key: str, shape: tuple, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31))
    if key.endswith('.pe'):
        return sinusoidal_pe(shape[0], shape[-1]).reshape(shape)
    if 'word_embeddings' in key:
        return torch.randn(shape, generator=g)
    leaf = key.rsplit('.', 2)
    is_ln = ('norm' in leaf[-2]) and 'project_layer' not in key
    if is_ln and key.endswith('weight'):
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if is_ln and key.endswith('bias'):
        return 0.1 * torch.randn(shape, generator=g)
    if key.endswith('weight'):
        bound = 1.0 / math.sqrt(shape[-1])
        return (torch.rand(shape, generator=g) * 2 - 1) * bound
    # Linear bias: fan_in is not recoverable from the bias alone; use the layer width.
    bound = 1.0 / math.sqrt(shape[0])
    return (torch.rand(shape, generator=g) * 2 - 1) * bound


def synth_state_dict(shapes: dict, seed: int = 0) -> dict:
    return {k: synth_tensor(k, tuple(v), seed) for k, v in shapes.items()}


def tiny_config(norm: str, **kw) -> OracleConfig:
    """BASELINE config 1: 2 layers, d=256, 4 heads, F=1024."""
    base = dict(num_layers=2, d_model=256, n_heads=4, dim_feedforward=1024, norm=norm,
                max_audio_len=40, top_k=1, num_beams=1)
    base.update(kw)
    return OracleConfig(**base)


def large_config(norm: str, **kw) -> OracleConfig:
    """BASELINE configs 2-4: 12 layers, d=1024, 16 heads, F=4096."""
    base = dict(num_layers=12, d_model=1024, n_heads=16, dim_feedforward=4096, norm=norm,
                max_audio_len=750, top_k=1, num_beams=1)
    base.update(kw)
    return OracleConfig(**base)


def tiny_inputs(seed: int = 0) -> dict:
    """Config-1 inputs: 7 prompt phonemes, 9 prompt frames, 5 target phonemes, 11 frames."""
    g = torch.Generator().manual_seed(1234 + seed)
    return {
        'prompt_tokens': torch.randint(0, 256, (7,), generator=g),
        'prompt_codes': torch.randint(0, 1024, (9, 8), generator=g),
        'target_tokens': torch.randint(0, 256, (5,), generator=g),
        'first_layer': torch.randint(0, 1024, (11,), generator=g),
    }
