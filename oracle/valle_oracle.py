"""TEST INFRASTRUCTURE ONLY -- fp32 CPU restatement of the Valle2 hot path.

This file restates, function by function, what the reference computes on the path
``ValleAR.generate`` / ``ValleAR.training_step`` / ``ValleNAR.generate`` and the module stack under
them.  It is written against a plain ``state_dict`` (name -> fp32 CPU tensor, the reference's own
key names) and uses torch CPU tensors only as an array library (matmul, exp, sort).
It never imports the reference and never touches CUDA.

Every function cites the reference lines it follows (paths relative to ``/root/reference``).
Repairs of reference defects (the upstream NAR code raises) are marked ``REPAIR A-n`` with the
numbering of SURVEY Appendix A.

Parity status: pinned -- ``tests/test_oracle_golden.py`` checks every function here against
vectors produced by executing the real reference (``oracle/make_golden.py``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

Tensor = torch.Tensor


@dataclass
class OracleConfig:
    """The subset of ``valle/config.py:7-64`` the hot path reads."""
    vocab_size: int = 256
    num_audio_tokens: int = 1024
    num_quantizers: int = 8
    d_model: int = 256
    n_heads: int = 4
    dim_feedforward: int = 1024
    num_layers: int = 8
    norm: str = 'AdaptiveLayerNorm'
    max_audio_len: int = 1024
    num_beams: int = 4
    top_k: int = 50
    tok_p: float = 1.0
    temperature: float = 1.0
    length_penalty: float = 1.0
    sampling_rate: int = 16000
    polling_factor: int = 320

    @property
    def eos_token(self) -> int:  # config.py:87-89
        return self.num_audio_tokens

    @property
    def bos_token(self) -> int:  # config.py:83-85
        return self.num_audio_tokens + 1

    @property
    def quantization_factor(self) -> int:  # config.py:79-81
        return self.sampling_rate // self.polling_factor

    @classmethod
    def from_any(cls, cfg) -> 'OracleConfig':
        if isinstance(cfg, cls):
            return cfg
        names = cls.__dataclass_fields__.keys()
        return cls(**{k: getattr(cfg, k) for k in names if hasattr(cfg, k)})


# ----------------------------------------------------------------------------------------------
# building blocks (valle/models/modules.py)
# ----------------------------------------------------------------------------------------------

def sinusoidal_pe(max_len: int, d_model: int) -> Tensor:
    """modules.py:60-66 -- the fp32 ``pe`` buffer, returned as ``(max_len, d_model)``."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def embed(table: Tensor, ids: Tensor) -> Tensor:
    """modules.py:33-37 -- row gather (dropout p=0)."""
    return table[ids]


def add_pe(x: Tensor, pe: Tensor) -> Tensor:
    """modules.py:78-80 -- ``x + pe[:T]`` in eval mode (dropout inactive)."""
    return x + pe[: x.shape[1]].unsqueeze(0)


def layer_norm(x: Tensor, weight: Tensor, bias: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.LayerNorm as used at modules.py:89,234-235: biased variance, eps inside the sqrt."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * weight + bias


def adaptive_layer_norm(x: Tensor, embedding: Tensor, proj_w: Tensor, proj_b: Tensor,
                        norm_w: Tensor, norm_b: Tensor, eps: float = 1e-5) -> Tensor:
    """modules.py:93-99 -- ``w * LN_affine(x) + b`` with ``[w, b] = Linear(d->2d)(embedding)``."""
    d = x.shape[-1]
    wb = embedding @ proj_w.t() + proj_b
    w, b = wb[..., :d], wb[..., d:]
    return w * layer_norm(x, norm_w, norm_b, eps) + b


def gelu_erf(x: Tensor) -> Tensor:
    """modules.py:216 -- ``nn.GELU()`` default = exact erf form (config.activation is ignored)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def feed_forward(x: Tensor, sd: dict, p: str) -> Tensor:
    """modules.py:220-221 -- ``linear_2(gelu(linear_1(x)))``, dropout inactive."""
    h = x @ sd[p + 'linear_1.weight'].t() + sd[p + 'linear_1.bias']
    return gelu_erf(h) @ sd[p + 'linear_2.weight'].t() + sd[p + 'linear_2.bias']


def merge_masks(batch_size: int, n_heads: int, attn_mask: Tensor | None,
                key_padding_mask: Tensor | None) -> Tensor | None:
    """modules.py:175-207 -- expand to ``(B,H,S,S)`` and ADD the key-padding mask."""
    if attn_mask is None:
        return None
    if attn_mask.dim() == 3:
        merged = attn_mask.unsqueeze(1)
    else:
        merged = attn_mask.unsqueeze(0).unsqueeze(0).expand(batch_size, n_heads, -1, -1)
    if key_padding_mask is not None:
        kpm = key_padding_mask.unsqueeze(1).unsqueeze(1).expand(batch_size, n_heads, 1, -1)
        merged = merged + kpm
    return merged


def sdpa(q: Tensor, k: Tensor, v: Tensor, allowed: Tensor | None) -> Tensor:
    """modules.py:167 -- softmax(q k^T / sqrt(Dh) [masked_fill(~allowed, -inf)]) v."""
    scale = 1.0 / math.sqrt(q.shape[-1])
    s = (q @ k.transpose(-1, -2)) * scale
    if allowed is not None:
        s = s.masked_fill(~allowed, float('-inf'))
    p = torch.softmax(s, dim=-1)
    return p @ v


def multi_head_attention(x: Tensor, sd: dict, p: str, n_heads: int, *, attn_mask=None,
                         padding_mask=None, kv_cache=None, use_cache=False):
    """modules.py:117-173."""
    B, n, d = x.shape
    dh = d // n_heads
    qkv = x @ sd[p + 'qkv.weight'].t()                                   # :146 (no bias)
    q, k, v = qkv.chunk(3, dim=-1)
    q, k, v = (t.reshape(B, n, n_heads, dh).permute(0, 2, 1, 3) for t in (q, k, v))   # :147
    kv = None
    if use_cache and kv_cache is not None:                               # :151-155
        k = torch.cat([kv_cache[0], k], dim=-2)
        v = torch.cat([kv_cache[1], v], dim=-2)
    if use_cache:
        kv = (k, v)
    allowed = None
    if attn_mask is not None:                                            # :160-164
        merged = merge_masks(B, n_heads, attn_mask, padding_mask)
        allowed = ~merged.to(torch.bool)
    a = sdpa(q, k, v, allowed)                                           # :167
    a = a.permute(0, 2, 1, 3).reshape(B, n, d)                           # :170
    out = a @ sd[p + 'out.weight'].t() + sd[p + 'out.bias']              # :171
    return out, kv


def _norm(x: Tensor, sd: dict, p: str, norm: str, embedding: Tensor | None) -> Tensor:
    if norm == 'LayerNorm':
        return layer_norm(x, sd[p + 'weight'], sd[p + 'bias'])
    return adaptive_layer_norm(x, embedding, sd[p + 'project_layer.weight'],
                               sd[p + 'project_layer.bias'], sd[p + 'norm.weight'],
                               sd[p + 'norm.bias'])


def encoder_layer(x: Tensor, sd: dict, p: str, cfg: OracleConfig, *, padding_mask=None,
                  attn_mask=None, embedding=None, kv_cache=None, use_cache=False):
    """modules.py:240-280 -- pre-norm block, dropout inactive."""
    a, kv = multi_head_attention(_norm(x, sd, p + 'norm1.', cfg.norm, embedding), sd,
                                 p + 'self_attn.', cfg.n_heads, attn_mask=attn_mask,
                                 padding_mask=padding_mask, kv_cache=kv_cache, use_cache=use_cache)
    x = x + a
    x = x + feed_forward(_norm(x, sd, p + 'norm2.', cfg.norm, embedding), sd, p + 'ffn.')
    return x, kv


def transformer(x: Tensor, sd: dict, cfg: OracleConfig, *, prefix: str = 'transformer.',
                padding_mask=None, attn_mask=None, embedding=None, kv_cache=None, use_cache=False):
    """modules.py:305-352 -- with a cache only the last position is kept and the mask dropped."""
    new_kv: tuple = ()
    if use_cache and kv_cache is not None:                               # :336-338
        x = x[:, -1:]
        attn_mask = None
    else:
        kv_cache = tuple([None] * cfg.num_layers)
    for i, past in zip(range(cfg.num_layers), kv_cache):
        x, kv = encoder_layer(x, sd, f'{prefix}layers.{i}.', cfg, padding_mask=padding_mask,
                              attn_mask=attn_mask, embedding=embedding, kv_cache=past,
                              use_cache=use_cache)
        if use_cache:
            new_kv = new_kv + (kv,)
    return x, new_kv


# ----------------------------------------------------------------------------------------------
# masks / sampling (valle/models/utils.py + transformers 4.38.2 warpers)
# ----------------------------------------------------------------------------------------------

def build_pad_mask(lens: Tensor) -> Tensor:
    """utils.py:8-14 -- True = padding."""
    max_len = int(lens.max())
    return torch.arange(max_len).unsqueeze(0).expand(len(lens), -1) >= lens.unsqueeze(1)


def build_attn_mask(x_len: int, y_len: int) -> Tensor:
    """utils.py:17-43 -- prefix-LM mask, True = masked."""
    n = x_len + y_len
    m = torch.zeros(n, n, dtype=torch.bool)
    m[:x_len, x_len:] = True
    m[x_len:, x_len:] = torch.triu(torch.ones(y_len, y_len, dtype=torch.bool), diagonal=1)
    return m


def top_k_top_p_filter(logits: Tensor, top_k: int, top_p: float) -> Tensor:
    """transformers 4.38.2 ``top_k_top_p_filtering`` (called at utils.py:63).

    top-k: ``scores < kth_largest`` -> -inf (ties with the k-th value are kept).
    top-p (runs whenever 0 <= p <= 1): ascending sort, softmax, cumsum, drop ``cum <= 1-p``,
    always keep the last (largest) entry, scatter back.  Ties in the sort are ordered by index
    (stable) -- the reference leaves tie order to ``torch.sort``.
    """
    logits = logits.clone()
    V = logits.shape[-1]
    if top_k > 0:
        k = min(top_k, V)
        kth = torch.topk(logits, k)[0][..., -1, None]
        logits = logits.masked_fill(logits < kth, float('-inf'))
    if 0 <= top_p <= 1.0:
        sorted_logits, sorted_idx = torch.sort(logits, descending=False, stable=True)
        cum = torch.softmax(sorted_logits, dim=-1).cumsum(dim=-1)
        remove = cum <= (1 - top_p)
        remove[..., -1:] = False
        remove = remove.scatter(1, sorted_idx, remove)
        logits = logits.masked_fill(remove, float('-inf'))
    return logits


def topk_sampling(logits: Tensor, top_k: int = 50, tok_p: float = 1.0,
                  temperature: float | None = 1.0, uniforms: Tensor | None = None):
    """utils.py:46-68.

    The reference draws with ``torch.multinomial`` (:64), whose RNG stream cannot be reproduced
    outside torch.  This restatement takes the draw explicitly: with ``uniforms`` (one u in [0,1)
    per row) the sample is the first index whose inclusive cumulative probability exceeds u
    (index order); with ``top_k == 1`` the sample is the lowest-index maximum (what multinomial
    returns when a single entry survives); otherwise it falls back to ``torch.multinomial``.
    Returns ``(samples (B,1) int64, logprobs (B,))`` like the reference.
    """
    if temperature is not None:
        logits = logits / temperature                                    # :59-60
    logits = top_k_top_p_filter(logits, top_k, tok_p)                    # :63
    probs = torch.softmax(logits, dim=-1)
    if uniforms is not None:
        cdf = probs.cumsum(dim=-1)
        samples = (cdf <= uniforms.reshape(-1, 1)).sum(dim=-1, keepdim=True)
        samples = samples.clamp(max=logits.shape[-1] - 1)
        # never land on a filtered (p == 0) slot through rounding of the cdf tail
        last_valid = (probs > 0).float().cumsum(-1).argmax(-1, keepdim=True)
        samples = torch.minimum(samples, last_valid)
    elif top_k == 1:
        samples = logits.argmax(dim=-1, keepdim=True)
    else:
        samples = torch.multinomial(probs, num_samples=1)                # :64
    logprobs = torch.log_softmax(logits, dim=-1)                         # :65
    current = logprobs[torch.arange(logits.shape[0]), samples[:, 0]]     # :66
    return samples, current


def get_best_beam(x: Tensor, sum_logprobs: Tensor, stop_token: int, length_penalty: float = 1.0):
    """utils.py:71-88."""
    length = torch.sum(x != stop_token, dim=-1)
    avg = sum_logprobs / length ** length_penalty
    best = x[torch.argmax(avg), :]
    return best[best != stop_token]


# ----------------------------------------------------------------------------------------------
# ValleAR (valle/models/valle_ar.py)
# ----------------------------------------------------------------------------------------------

def _pe_for(sd: dict, key: str, d_model: int) -> Tensor:
    if key in sd:
        return sd[key].reshape(-1, d_model)          # stored as (5000,1,d)
    return sinusoidal_pe(5000, d_model)


def ar_teacher_forced(sd: dict, cfg, tokens: Tensor, codes: Tensor, tokens_lens: Tensor,
                      codes_lens: Tensor, target: Tensor | None = None):
    """valle_ar.py:43-90 -- returns ``(logits (B,Ty,V+1), loss|None)``; eval mode (K-3)."""
    cfg = OracleConfig.from_any(cfg)
    d = cfg.d_model
    x_tok = add_pe(embed(sd['tokens_emb.word_embeddings.weight'], tokens),
                   _pe_for(sd, 'tokens_position_emb.pe', d))              # :61-62
    x_aud = add_pe(embed(sd['audio_emb.word_embeddings.weight'], codes),
                   _pe_for(sd, 'audio_position_emb.pe', d))               # :65-66
    max_tx = int(tokens_lens.max())
    pad = build_pad_mask(codes_lens)                                     # :69-73 (K-4)
    pad = torch.cat([torch.zeros(pad.shape[0], max_tx, dtype=torch.bool), pad], dim=1)
    attn = build_attn_mask(max_tx, int(codes_lens.max()))                # :74
    h, _ = transformer(torch.cat((x_tok, x_aud), dim=1), sd, cfg, padding_mask=pad, attn_mask=attn)
    h = h[:, max_tx:]                                                    # :80
    logits = h @ sd['proj.weight'].t()                                   # :83
    loss = None
    if target is not None:                                               # :86 (K-5: no ignore_index)
        logp = torch.log_softmax(logits, dim=-1)
        loss = -logp.gather(-1, target.unsqueeze(-1)).mean()
    return logits, loss


def ar_generate(sd: dict, cfg, prompt_tokens: Tensor, prompt_codes: Tensor,
                target_tokens: Tensor | None = None, uniforms: Tensor | None = None,
                return_trace: bool = False, max_steps: int | None = None):
    """valle_ar.py:92-180 (KV-cached path; A-12: the only path that runs).

    ``uniforms``: optional ``(max_audio_len, num_beams)`` injected draws (see topk_sampling).
    ``return_trace``: also return per-step logits ``list[(B,V+1)]`` and the beam matrix.
    """
    cfg = OracleConfig.from_any(cfg)
    d = cfg.d_model
    assert prompt_tokens.dim() == 1 and prompt_codes.dim() == 2           # :109-112
    codes = torch.cat([torch.tensor([cfg.bos_token]), prompt_codes[:, 0]]).unsqueeze(0)   # :115-117
    prompt_len = codes.shape[1]
    tokens = prompt_tokens if target_tokens is None else torch.cat((prompt_tokens, target_tokens))
    tokens_len = tokens.shape[0]
    pe_t = _pe_for(sd, 'tokens_position_emb.pe', d)
    pe_a = _pe_for(sd, 'audio_position_emb.pe', d)
    x_tok = add_pe(embed(sd['tokens_emb.word_embeddings.weight'], tokens.unsqueeze(0)), pe_t)
    attn_mask = build_attn_mask(tokens_len, prompt_len)                  # :132
    kv_cache = None
    B = cfg.num_beams
    sum_logprobs = torch.zeros(B)
    x_tok = x_tok.repeat(B, 1, 1)                                        # :137-138
    codes = codes.repeat(B, 1)
    trace = []
    n_steps = cfg.max_audio_len if max_steps is None else max_steps
    for step in range(n_steps):                                          # :141
        x_aud = add_pe(embed(sd['audio_emb.word_embeddings.weight'], codes), pe_a)   # :143-144
        h, kv_cache = transformer(torch.cat([x_tok, x_aud], dim=1), sd, cfg, attn_mask=attn_mask,
                                  kv_cache=kv_cache, use_cache=True)     # :150-155
        logits = (h @ sd['proj.weight'].t())[:, -1]                      # :158
        if return_trace:
            trace.append(logits.clone())
        u = None if uniforms is None else uniforms[step]
        samples, logp = topk_sampling(logits, cfg.top_k, cfg.tok_p, cfg.temperature, u)
        last = codes[:, -1]
        sum_logprobs = sum_logprobs + logp * (last != cfg.eos_token)     # :167
        samples[last == cfg.eos_token] = cfg.eos_token                   # :168
        if bool((samples[:, -1] == cfg.eos_token).all()):                # :169-170
            break
        codes = torch.cat([codes, samples], dim=1)                       # :171
    out = get_best_beam(codes, sum_logprobs, cfg.eos_token, cfg.length_penalty)   # :174-176
    out = out[prompt_len:]
    out = out[out != cfg.eos_token]
    if return_trace:
        return out, trace, codes, sum_logprobs
    return out


# ----------------------------------------------------------------------------------------------
# ValleNAR (valle/models/valle_nar.py) -- upstream generate()/training_step() raise; this follows
# the source with the repairs of SURVEY Appendix A.
# ----------------------------------------------------------------------------------------------

def nar_generate(sd: dict, cfg, prompt_tokens: Tensor, prompt_codes: Tensor,
                 target_tokens: Tensor, target_codes_first_layer: Tensor, greedy: bool = True,
                 return_trace: bool = False):
    """valle_nar.py:107-165 -- returns ``(T, Q)`` int64.

    REPAIR A-5: float ``(.,d)`` accumulators (source: ``zeros_like`` of int tensors, :127-128).
    REPAIR A-6: stage n adds ``codes_embs[n-1](codes of stage n-1)`` (source :144 uses index n
                and the concatenated tensor; training :180-185 defines the intended tables).
    REPAIR A-7: unpack the ``(x, kv)`` tuple returned by Transformer (:152).
    REPAIR A-8: new codebooks are stacked on a new last axis -> ``(T,Q)`` (:163, docstring :124).
    EXTENSION A-9: ``greedy`` = argmax instead of ``Categorical(...).sample()`` (:160).
    """
    cfg = OracleConfig.from_any(cfg)
    d = cfg.d_model
    Tc, Q = prompt_codes.shape
    emb_prompt = torch.zeros(Tc, d)
    for j in range(Q):                                                   # :131-133
        emb_prompt = emb_prompt + sd[f'codes_embs.{j}.word_embeddings.weight'][prompt_codes[:, j]]
    tokens = torch.cat([prompt_tokens, target_tokens]).unsqueeze(0)      # :136
    Tx = tokens.shape[1]
    pe_t = _pe_for(sd, 'tokens_position_emb.pe', d)
    pe_a = _pe_for(sd, 'audio_position_emb.pe', d)
    x_tok = add_pe(embed(sd['tokens_emb.word_embeddings.weight'], tokens), pe_t)   # :138-139
    T = target_codes_first_layer.shape[0]
    emb_out = torch.zeros(T, d)
    columns = [target_codes_first_layer]
    trace = []
    for n in range(1, Q):                                                # :142
        emb_out = emb_out + sd[f'codes_embs.{n - 1}.word_embeddings.weight'][columns[n - 1]]
        x_aud = add_pe(torch.cat([emb_prompt, emb_out], dim=0).unsqueeze(0), pe_a)   # :145-148
        h, _ = transformer(torch.cat([x_tok, x_aud], dim=1), sd, cfg,
                           embedding=sd[f'stage_embs.{n - 1}.word_embeddings.weight'])   # :152-154
        logits = h[:, Tx + Tc:] @ sd[f'proj_layers.{n - 1}.weight'].t()   # :157
        if return_trace:
            trace.append(logits[0].clone())
        if greedy:
            sampled = logits[0].argmax(dim=-1)
        else:
            sampled = torch.distributions.Categorical(logits=logits[0] / cfg.temperature).sample()
        columns.append(sampled)
    out = torch.stack(columns, dim=-1)
    if return_trace:
        return out, trace
    return out


def nar_prepare_audio_codes(sd: dict, cfg, codes: Tensor, nar_stage: int):
    """valle_nar.py:167-188 -- ``(y_emb (B,T,d), prefix_len)``."""
    cfg = OracleConfig.from_any(cfg)
    _, T, Q = codes.shape
    prefix_len = min(T // 3, 3 * cfg.quantization_factor)                # :178
    tab = [sd[f'codes_embs.{j}.word_embeddings.weight'] for j in range(Q)]
    prompts = tab[0][codes[:, :prefix_len, 0]]
    emb = tab[0][codes[:, prefix_len:, 0]]
    for j in range(1, Q):                                                # :182-185
        prompts = prompts + tab[j][codes[:, :prefix_len, j]]
        if j < nar_stage:
            emb = emb + tab[j][codes[:, prefix_len:, j]]
    return torch.cat((prompts, emb), dim=1), prefix_len


def nar_teacher_forced(sd: dict, cfg, tokens: Tensor, codes: Tensor, tokens_lens: Tensor,
                       codes_lens: Tensor, layer: int):
    """valle_nar.py:53-105 for a fixed stage ``layer`` (the source draws it at :76).

    REPAIR A-1: ``batch['target']`` is not read (the collate never provides it, :68).
    REPAIR A-2: target = raw int codes ``codes[:, prefix_len:, layer]`` (source slices embeddings).
    REPAIR A-3: logits over ``z[:, Tx+prefix_len:]`` (source indexes a single position, :97).
    KEEP A-4: the padding mask is passed without an attn_mask and is therefore ignored.
    Returns ``(logits (B,T-prefix,V), loss)``.
    """
    cfg = OracleConfig.from_any(cfg)
    d = cfg.d_model
    max_tx = int(tokens_lens.max())
    x_tok = add_pe(embed(sd['tokens_emb.word_embeddings.weight'], tokens),
                   _pe_for(sd, 'tokens_position_emb.pe', d))
    y_emb, prefix_len = nar_prepare_audio_codes(sd, cfg, codes, layer)
    y_emb = add_pe(y_emb, _pe_for(sd, 'audio_position_emb.pe', d))
    target = codes[:, prefix_len:, layer]
    z, _ = transformer(torch.cat([x_tok, y_emb], dim=1), sd, cfg,
                       embedding=sd[f'stage_embs.{layer - 1}.word_embeddings.weight'])
    z = z[:, max_tx + prefix_len:]
    logits = z @ sd[f'proj_layers.{layer - 1}.weight'].t()
    logp = torch.log_softmax(logits, dim=-1)
    loss = -logp.gather(-1, target.unsqueeze(-1)).mean()
    return logits, loss
