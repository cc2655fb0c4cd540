"""Teacher-forced training step on the CUDA stack: forward with saved activations + hand-written backward.

``ValleAR.training_step`` (valle_ar.py:43-90) and ``ValleNAR.training_step`` (valle_nar.py:53-105, repaired per SURVEY
App. A) return a loss that Lightning back-propagates with torch autograd.  Here the whole step -- embeddings, the
pre-norm transformer stack, logits, mean cross entropy, and every gradient -- runs on the kernels of libvalle_b200.so;
``step_loss`` hands the result to autograd through one ``torch.autograd.Function`` whose inputs are the model's parameters,
so ``loss.backward()`` / an optimizer step work exactly as with the reference.

GEMMs: forward and backward run on the tcgen05 GEMM (bf16) or the SIMT fp32 GEMM (validation mode).  For y = x W^T:
    dgrad  dx (R,K) = dy (R,N) . W (N,K)        bf16: ``ops.linear_t(dy, W, w_t=True)`` -- W itself is the MN-major B operand
    wgrad  dW (N,K) = dy^T (N,R) . x (R,K)      bf16: ``ops.linear_t(dy, x, x_t=True, w_t=True)`` -- dy and x are read in place
so no operand is transposed in HBM (147 transpose launches, 9.6 ms of a 70 ms step, are gone); the fp32 validation mode keeps
the explicit transposes (``vb_transpose``) in front of ``ops.linear``.
R (= B*S rows) is padded to a multiple of 8 with zero rows (TMA pitch rule of the bf16 GEMM).
Everything else (LayerNorm / GELU / attention backward, cross entropy, embedding scatter, bias column sums) is csrc/train.cu.

Precision: 'bf16' = bf16 operands, fp32 accumulation, fp32 residual stream and fp32 gradients of the residual stream and
of all parameters; 'fp32' = validation mode (gradients within 1e-4 of torch autograd on the CPU oracle).
Dropout (``model.train()`` with ``config.dropout`` > 0, and the PositionalEncoding's hard-wired p = 0.1, SURVEY K-3): applied at
the reference's four sites -- after embedding + PE (modules.py:78), inside the FFN after GELU (:221), on the attention and FFN
outputs before the residual adds (dropout1 / dropout2, :277-278) -- with counter-based masks (``vb_dropout``): the mask is a
hash of (step seed, site, element), recomputed by the backward pass instead of being stored.  torch's own RNG stream cannot be
reproduced, so parity is checked by giving the executed reference the SAME masks (tests/test_gpu_training.py).
"""
from __future__ import annotations

import os

import torch

from . import ops, parallel
from .ops import MASK_NONE, MASK_PREFIX_LM


def _cd(precision: str) -> torch.dtype:
    return torch.bfloat16 if precision == 'bf16' else torch.float32


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def _i32(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.int32).contiguous()


def _ids_in_range(t: torch.Tensor, hi: int, what: str) -> None:
    """The reference's nn.Embedding / F.cross_entropy fail on an id outside [0, hi) (a device-side assert on CUDA); the
    kernels in csrc/train.cu clamp instead, which would hide a data bug -- so the same check is made here, asynchronously
    (two tiny launches, no host synchronisation; a violation surfaces as a CUDA device-side assert, like the reference's)."""
    torch._assert_async(((t >= 0) & (t < hi)).all(), f'{what}: id outside [0, {hi})')


class _Lin:
    """One nn.Linear in the compute dtype: W (N,K), its transpose for dgrad, fp32 bias."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor | None, cd: torch.dtype, pad_n: bool = False):
        w = weight.detach().to(cd)
        self.N = w.shape[0]
        if pad_n and self.N % 8:                      # logits head (V = 1025): zero rows so that the dgrad K is aligned
            w = torch.cat([w, torch.zeros(_pad8(self.N) - self.N, w.shape[1], device=w.device, dtype=cd)], 0)
        self.w = w.contiguous()
        self._wt = None
        self.b = None if bias is None else bias.detach().float().contiguous()

    @property
    def wt(self) -> torch.Tensor:
        """(K, Npad) copy for the dgrad of the fp32 validation mode; the bf16 path reads W itself as an MN-major operand."""
        if self._wt is None:
            self._wt = ops.transpose(self.w)
        return self._wt


def _linear_fwd(x: torch.Tensor, lin: _Lin, *, residual: torch.Tensor | None = None, out: torch.Tensor | None = None,
                out_dtype: torch.dtype | None = None) -> torch.Tensor:
    return ops.linear(x, lin.w, lin.b, residual=residual, out=out, out_dtype=out_dtype)


def _rows(Rp: int, n: int, R: int, device, dtype: torch.dtype) -> torch.Tensor:
    """Activation / gradient buffer of Rp rows of which R are live.  bf16: uninitialised, only the <= 7 pad rows are zeroed
    (the kernels write every live row; the GEMMs read rows [:R] or zero-fill past them) -- a full memset of these buffers
    was ~12 GB of HBM writes per step.  fp32 validation mode: zero-filled (its transposed wgrad copies read the pad rows)."""
    if dtype != torch.bfloat16:
        return torch.zeros(Rp, n, device=device, dtype=dtype)
    t = torch.empty(Rp, n, device=device, dtype=dtype)
    if Rp > R:
        t[R:].zero_()
    return t


SITE_PE = 1 << 20            # dropout site ids: PE = SITE_PE, layer li: 3 * li + {0: attention out, 1: FFN inner, 2: FFN out}


def dropout_plan(model, seed: int | None = None) -> dict | None:
    """Dropout of this training step, or None when nothing is dropped (eval mode or all rates zero).  Every call advances the
    model's step counter so that consecutive steps draw different masks; ``seed`` pins the masks (tests)."""
    if not model.training:
        return None
    p = float(model.config.dropout)
    pe_t, pe_a = float(model.tokens_position_emb.dropout.p), float(model.audio_position_emb.dropout.p)
    assert pe_t == pe_a, 'text and audio PositionalEncoding dropout rates differ'
    if p <= 0.0 and pe_t <= 0.0:
        return None
    if seed is None:
        step = getattr(model, '_dropout_step', 0)
        model._dropout_step = step + 1
        seed = (torch.initial_seed() * 1000003 + step) & ((1 << 63) - 1)
    return {'p': p, 'pe_p': pe_t, 'seed': int(seed)}


def zero_pe_dropout(model) -> None:
    """The PositionalEncoding dropout is hard-wired to p = 0.1 independently of config.dropout (modules.py:56, SURVEY K-3):
    parity runs against a dropout-free oracle switch it off with this."""
    model.tokens_position_emb.dropout.p = 0.0
    model.audio_position_emb.dropout.p = 0.0


def dropout_masks(plan: dict, n_layers: int, rows: int, d: int, F: int, device) -> dict:
    """The step's keep masks (scaled by 1 / (1 - p)) as dense fp32 tensors, generated by the same kernel -- for tests that feed
    them to the reference in place of nn.Dropout."""
    out = {'pe': ops.dropout_(torch.ones(rows, d, device=device), plan['pe_p'], plan['seed'], SITE_PE)}
    for li in range(n_layers):
        out[f'{li}.attn'] = ops.dropout_(torch.ones(rows, d, device=device), plan['p'], plan['seed'], 3 * li)
        out[f'{li}.ffn_inner'] = ops.dropout_(torch.ones(rows, F, device=device), plan['p'], plan['seed'], 3 * li + 1)
        out[f'{li}.ffn'] = ops.dropout_(torch.ones(rows, d, device=device), plan['p'], plan['seed'], 3 * li + 2)
    return out


def _mn_ok(*ts: torch.Tensor) -> bool:
    """The tcgen05 GEMM can read these row-major matrices as MN-major operands (csrc/gemm_tc.cu, vb_linear_t)."""
    return all(t.dtype == torch.bfloat16 and t.stride(0) % 8 == 0 for t in ts)


def _wgrad(dy: torch.Tensor, x: torch.Tensor, R: int) -> torch.Tensor:
    """dW (N,K) fp32 = dy^T . x over the first R rows of dy (Rp,N) and x (Rp,K).  bf16: both matrices are read in place as
    MN-major MMA operands; fp32 validation mode: transposed copies (zero rows past R) + the SIMT GEMM."""
    if _mn_ok(dy, x) and x.shape[1] > 128:
        return ops.linear_t(dy[:R], x[:R], x_t=True, w_t=True, out_dtype=torch.float32)
    return ops.linear(ops.transpose(dy), ops.transpose(x), out_dtype=torch.float32)


def _dgrad(dy: torch.Tensor, lin: _Lin, out: torch.Tensor) -> torch.Tensor:
    """dx (R,K) = dy (R,N) . W (N,K): W is the MN-major B operand (bf16) or its transposed copy (fp32 mode)."""
    if _mn_ok(dy, lin.w) and lin.w.shape[1] > 128:
        return ops.linear_t(dy, lin.w, w_t=True, out=out)
    return ops.linear(dy, lin.wt, out=out)


class StackTrainer:
    """Forward (saving what the backward needs) and backward of a ``Transformer`` stack over packed rows."""

    def __init__(self, transformer, n_heads: int, precision: str, norm: str, stage_embs=None):
        self.cd = _cd(precision)
        self.H = n_heads
        self.norm = norm
        # the tensor-core forward hands its log-sum-exp to the backward kernels (VALLE_B200_SAVE_LSE=0: recompute it there;
        # the SIMT backward of VALLE_B200_ATTN_BWD_SIMT=1 always recomputes)
        self.save_lse = (os.environ.get('VALLE_B200_SAVE_LSE', '1') != '0' and os.environ.get('VALLE_B200_ATTN_BWD_SIMT', '0') != '1')
        self.layers = []
        for layer in transformer.layers:
            a, f = layer.self_attn, layer.ffn
            L = {'qkv': _Lin(a.qkv.weight, None, self.cd), 'o': _Lin(a.out.weight, a.out.bias, self.cd),
                 'f1': _Lin(f.linear_1.weight, f.linear_1.bias, self.cd), 'f2': _Lin(f.linear_2.weight, f.linear_2.bias, self.cd)}
            for name in ('norm1', 'norm2'):
                nm = getattr(layer, name)
                if norm == 'LayerNorm':
                    L[name] = {'g': nm.weight.detach().float().contiguous(), 'b': nm.bias.detach().float().contiguous(), 'eps': nm.eps}
                else:
                    L[name] = {'mod': nm, 'eps': nm.eps}
            self.layers.append(L)
        self.stage_embs = stage_embs
        self.d = self.layers[0]['o'].N
        self.F = self.layers[0]['f1'].N

    # -- stage-conditioned AdaLN folded into one affine LayerNorm per (layer, norm): modules.py:93-99 -------------------
    def _affine(self, L: dict, name: str, stage: int):
        nrm = L[name]
        if self.norm == 'LayerNorm':
            return nrm['g'], nrm['b'], None
        nm = nrm['mod']
        d = self.d
        e = self.stage_embs[stage].detach().float().reshape(1, d).contiguous()
        wb = ops.linear(e, nm.project_layer.weight.detach().float().contiguous(), nm.project_layer.bias.detach().float().contiguous())
        w, b = wb[0, :d], wb[0, d:]
        g0, b0 = nm.norm.weight.detach().float(), nm.norm.bias.detach().float()
        return (w * g0).contiguous(), (w * b0 + b).contiguous(), {'w': w, 'e': e, 'g0': g0, 'b0': b0}

    def forward(self, x: torch.Tensor, B: int, S: int, *, mask_mode: int, x_lens, kv_lens, stage: int = 0, drop: dict | None = None):
        """x (Rp, d) fp32 residual stream (rows >= B*S are zero padding), updated in place.  Returns the cache.
        drop = dropout_plan(model): dropout1 / dropout2 / the FFN's inner dropout with rate drop['p'] (0: the fused epilogues)."""
        p_drop, seed = (float(drop['p']), drop['seed']) if drop else (0.0, 0)
        Rp, d = x.shape
        R = B * S
        cd, H, F = self.cd, self.H, self.F
        dev = x.device
        use_tc = cd == torch.bfloat16 and d // H == 64
        cache = []
        for li, L in enumerate(self.layers):
            c = {}
            g1, b1, c['fold1'] = self._affine(L, 'norm1', stage)
            c['g1'] = g1
            c['x_in'] = x.clone()
            c['h1'] = _rows(Rp, d, R, dev, cd)
            ops.residual_layernorm(x[:R], g1, b1, c['h1'][:R], eps=L['norm1']['eps'])
            c['qkv'] = _rows(Rp, 3 * d, R, dev, cd)
            _linear_fwd(c['h1'][:R], L['qkv'], out=c['qkv'][:R])
            c['o'] = _rows(Rp, d, R, dev, cd)
            # the tensor-core forward also leaves the rows' log-sum-exp for the backward pass (5.5 KB per head and sequence)
            c['lse'] = torch.empty(B, H, S, device=dev, dtype=torch.float32) if (use_tc and self.save_lse) else None
            ops.attention_packed(c['qkv'][:R], c['o'][:R], B, S, H, mask_mode=mask_mode, x_lens=x_lens, kv_lens=kv_lens, use_tc=use_tc,
                                 lse=c['lse'])
            if p_drop > 0.0:        # x = x + dropout1(attn): the GEMM's residual epilogue cannot drop, so the add is its own kernel
                t = _rows(Rp, d, R, dev, cd)
                _linear_fwd(c['o'][:R], L['o'], out=t[:R])
                ops.dropout_add_(x[:R], t[:R], p_drop, seed, 3 * li)
            else:
                _linear_fwd(c['o'][:R], L['o'], residual=x[:R], out=x[:R])
            g2, b2, c['fold2'] = self._affine(L, 'norm2', stage)
            c['g2'] = g2
            c['x_mid'] = x.clone()
            c['h2'] = _rows(Rp, d, R, dev, cd)
            ops.residual_layernorm(x[:R], g2, b2, c['h2'][:R], eps=L['norm2']['eps'])
            c['f_pre'] = _rows(Rp, F, R, dev, cd)
            _linear_fwd(c['h2'][:R], L['f1'], out=c['f_pre'][:R])
            c['f'] = _rows(Rp, F, R, dev, cd)
            ops.gelu_fwd(c['f_pre'], c['f'])
            if p_drop > 0.0:
                ops.dropout_(c['f'][:R], p_drop, seed, 3 * li + 1)          # the cached f is the dropped one: what linear_2 saw
                t = _rows(Rp, d, R, dev, cd)
                _linear_fwd(c['f'][:R], L['f2'], out=t[:R])
                ops.dropout_add_(x[:R], t[:R], p_drop, seed, 3 * li + 2)
            else:
                _linear_fwd(c['f'][:R], L['f2'], residual=x[:R], out=x[:R])
            cache.append(c)
        return cache

    def backward(self, dx: torch.Tensor, cache: list, B: int, S: int, *, mask_mode: int, x_lens, kv_lens,
                 drop: dict | None = None, on_layer=None) -> list[dict]:
        """dx (Rp, d) fp32: gradient wrt the stack output, turned in place into the gradient wrt the stack input.
        Returns per-layer parameter gradients (fp32).  drop: the plan the forward ran with (same masks, recomputed)."""
        p_drop, seed = (float(drop['p']), drop['seed']) if drop else (0.0, 0)
        Rp, d = dx.shape
        R = B * S
        cd, H = self.cd, self.H
        dev = dx.device
        grads = [None] * len(self.layers)
        dxb = _rows(Rp, d, R, dev, cd)
        for li in range(len(self.layers) - 1, -1, -1):
            L, c = self.layers[li], cache[li]
            g = {}
            # ---- x_out = x_mid + dropout2(f W2^T + b2) ----
            if p_drop > 0.0:        # gradient of the dropped branch = dx through the same mask; dx itself flows on unchanged
                dt = ops.dropout_(dx[:R].clone(), p_drop, seed, 3 * li + 2)
                ops.residual_layernorm(dt, None, None, dxb[:R])
                g['f2.b'] = ops.colsum(dt)
            else:
                ops.residual_layernorm(dx[:R], None, None, dxb[:R])                # cast of the residual gradient
                g['f2.b'] = ops.colsum(dx[:R])
            g['f2.w'] = _wgrad(dxb, c['f'], R)                                     # (d, F)
            df = _rows(Rp, self.F, R, dev, cd)
            _dgrad(dxb[:R], L['f2'], df[:R])                                       # (R, F)
            if p_drop > 0.0:
                ops.dropout_(df[:R], p_drop, seed, 3 * li + 1)                     # through the FFN's inner dropout
            dpre = ops.gelu_bwd(c['f_pre'], df, df)                                # in place
            g['f1.b'] = ops.colsum(dpre[:R])
            g['f1.w'] = _wgrad(dpre, c['h2'], R)                                   # (F, d)
            dh = _rows(Rp, d, R, dev, cd)
            _dgrad(dpre[:R], L['f1'], dh[:R])
            g['n2.g'], g['n2.b'] = ops.layernorm_bwd(c['x_mid'][:R], c['g2'], dh[:R], dx[:R], L['norm2']['eps'])
            # ---- x_mid = x_in + dropout1(o Wo^T + bo) ----
            if p_drop > 0.0:
                dt = ops.dropout_(dx[:R].clone(), p_drop, seed, 3 * li)
                ops.residual_layernorm(dt, None, None, dxb[:R])
                g['o.b'] = ops.colsum(dt)
            else:
                ops.residual_layernorm(dx[:R], None, None, dxb[:R])
                g['o.b'] = ops.colsum(dx[:R])
            g['o.w'] = _wgrad(dxb, c['o'], R)                                      # (d, d)
            do = _rows(Rp, d, R, dev, cd)
            _dgrad(dxb[:R], L['o'], do[:R])
            dqkv = _rows(Rp, 3 * d, R, dev, cd)
            ops.attention_bwd(c['qkv'][:R], c['o'][:R], do[:R], dqkv[:R], B, S, H, mask_mode=mask_mode, x_lens=x_lens, kv_lens=kv_lens,
                              lse=c.get('lse'))
            g['qkv.w'] = _wgrad(dqkv, c['h1'], R)                                  # (3d, d)
            _dgrad(dqkv[:R], L['qkv'], dh[:R])
            g['n1.g'], g['n1.b'] = ops.layernorm_bwd(c['x_in'][:R], c['g1'], dh[:R], dx[:R], L['norm1']['eps'])
            g['fold1'], g['fold2'] = c['fold1'], c['fold2']
            grads[li] = g
            cache[li] = None                                                        # release the layer's activations
            if on_layer is not None:
                on_layer(li, g)                                                     # e.g. start this layer's gradient all-reduce
        return grads


def _layer_param_grads(prefix: str, layer_mod, g: dict, norm: str, out: dict, stage_key: str | None):
    """Map one layer's gradient dict onto the reference's parameter names (state_dict keys)."""
    out[prefix + 'self_attn.qkv.weight'] = g['qkv.w']
    out[prefix + 'self_attn.out.weight'] = g['o.w']
    out[prefix + 'self_attn.out.bias'] = g['o.b']
    out[prefix + 'ffn.linear_1.weight'] = g['f1.w']
    out[prefix + 'ffn.linear_1.bias'] = g['f1.b']
    out[prefix + 'ffn.linear_2.weight'] = g['f2.w']
    out[prefix + 'ffn.linear_2.bias'] = g['f2.b']
    for name, key in (('norm1', 'n1'), ('norm2', 'n2')):
        dgp, dbp = g[key + '.g'], g[key + '.b']                   # gradients of the folded gamma', beta'
        if norm == 'LayerNorm':
            out[prefix + name + '.weight'] = dgp
            out[prefix + name + '.bias'] = dbp
            continue
        # AdaLN (modules.py:93-99): gamma' = w * g0, beta' = w * b0 + b with [w, b] = project_layer(e_stage).
        # The chain through these d-sized vectors is host-side glue on tiny tensors.
        fold = g['fold1' if name == 'norm1' else 'fold2']
        w, e, g0, b0 = fold['w'], fold['e'], fold['g0'], fold['b0']
        dw = dgp * g0 + dbp * b0
        dwb = torch.cat([dw, dbp])                                # gradient of the project_layer output (2d)
        nm = getattr(layer_mod, name)
        out[prefix + name + '.norm.weight'] = dgp * w
        out[prefix + name + '.norm.bias'] = dbp * w
        out[prefix + name + '.project_layer.weight'] = torch.outer(dwb, e[0])
        out[prefix + name + '.project_layer.bias'] = dwb
        de = dwb @ nm.project_layer.weight.detach().float()
        out[stage_key] = out.get(stage_key, 0) + de.reshape(1, -1)


@torch.no_grad()
def ar_loss_and_grads(model, batch: dict, precision: str, drop: dict | None = None):
    """ValleAR.training_step (valle_ar.py:43-90): returns (loss scalar tensor, {param name: grad})."""
    cfg, dev = model.config, model.device
    cd = _cd(precision)
    tokens, codes = batch['tokens'].to(dev), batch['codes'].to(dev)
    tokens_lens, codes_lens = batch['tokens_lens'], batch['codes_lens']
    B = tokens.shape[0]
    Tx, Ty = int(tokens_lens.max()), int(codes_lens.max())
    tokens, codes = tokens[:, :Tx], codes[:, :Ty]
    S, d, V = Tx + Ty, cfg.d_model, cfg.num_audio_tokens + 1
    R, Rp = B * S, _pad8(B * S)
    tr = StackTrainer(model.transformer, cfg.n_heads, precision, cfg.norm)
    tok_table = model.tokens_emb.weight.detach().float().unsqueeze(0).contiguous()
    aud_table = model.audio_emb.weight.detach().float().unsqueeze(0).contiguous()
    pe_t = model.tokens_position_emb.pe.detach().float().reshape(-1, d).contiguous()
    pe_a = model.audio_position_emb.pe.detach().float().reshape(-1, d).contiguous()
    tok_i, cod_i = _i32(tokens, dev).view(B, Tx, 1), _i32(codes, dev).view(B, Ty, 1)
    _ids_in_range(tok_i, tok_table.shape[1], 'ValleAR.training_step tokens')
    _ids_in_range(cod_i, aud_table.shape[1], 'ValleAR.training_step codes')
    x = torch.zeros(Rp, d, device=dev, dtype=torch.float32)
    ops.embed_sum_pe(tok_i, tok_table, pe_t, x, out_rows_per_batch=S)
    ops.embed_sum_pe(cod_i, aud_table, pe_a, x, out_rows_per_batch=S, out_row_offset=Tx)
    if drop and drop['pe_p'] > 0.0:
        ops.dropout_(x[:R], drop['pe_p'], drop['seed'], SITE_PE)                     # PositionalEncoding.dropout (modules.py:78)
    xl = torch.full((B,), Tx, device=dev, dtype=torch.int32)
    kv_lens = (xl + _i32(codes_lens, dev)).contiguous()
    cache = tr.forward(x, B, S, mask_mode=MASK_PREFIX_LM, x_lens=xl, kv_lens=kv_lens, drop=drop)
    # logits over the audio rows (valle_ar.py:80-83), mean CE over every position incl. padding (K-5)
    Ra, Rap = B * Ty, _pad8(B * Ty)
    rows = x[:R].view(B, S, d)[:, Tx:].reshape(Ra, d).contiguous()
    proj = _Lin(model.proj.weight, None, cd, pad_n=True)
    Vp = proj.w.shape[0]
    hb = torch.zeros(Rap, d, device=dev, dtype=cd)
    ops.residual_layernorm(rows, None, None, hb[:Ra])                              # no final norm (K-2): plain cast
    logits = torch.zeros(Rap, Vp, device=dev, dtype=torch.float32)
    ops.linear(hb[:Ra], proj.w, out=logits[:Ra])
    target = _i32(batch['target'].to(dev)[:, :Ty], dev).reshape(-1)
    _ids_in_range(target, V, 'ValleAR.training_step target')
    dlogits = torch.zeros(Rap, Vp, device=dev, dtype=torch.float32)
    loss_rows = ops.cross_entropy(logits[:Ra], target, V, dlogits=dlogits[:Ra], scale=1.0 / Ra)
    loss = ops.colsum(loss_rows.view(Ra, 1), scale=1.0 / Ra)[0]
    grads = {}
    dl = torch.zeros(Rap, Vp, device=dev, dtype=cd)
    ops.residual_layernorm(dlogits[:Ra], None, None, dl[:Ra])
    grads['proj.weight'] = _wgrad(dl, hb, Ra)[:V]                                  # (V, d)
    dh = torch.zeros(Rap, d, device=dev, dtype=cd)
    _dgrad(dl[:Ra], proj, dh[:Ra])
    dx = torch.zeros(Rp, d, device=dev, dtype=torch.float32)
    dx[:R].view(B, S, d)[:, Tx:] = dh[:Ra].view(B, Ty, d).float()                   # scatter into the audio rows (plumbing)
    reducer = parallel.active_reducer()

    def layer_done(li, g):
        out = {}
        _layer_param_grads(f'transformer.layers.{li}.', model.transformer.layers[li], g, cfg.norm, out, None)
        grads.update(out)
        if reducer is not None:
            reducer.submit(li, out)          # the layer's all-reduce starts while the layers below are differentiated

    if reducer is not None:
        reducer.begin()
    tr.backward(dx, cache, B, S, mask_mode=MASK_PREFIX_LM, x_lens=xl, kv_lens=kv_lens, drop=drop, on_layer=layer_done)
    if drop and drop['pe_p'] > 0.0:
        ops.dropout_(dx[:R], drop['pe_p'], drop['seed'], SITE_PE)
    gt = torch.zeros_like(tok_table)
    ga = torch.zeros_like(aud_table)
    ops.embed_bwd(tok_i, dx, gt, rows_per_batch=S)
    ops.embed_bwd(cod_i, dx, ga, rows_per_batch=S, row_offset=Tx)
    grads['tokens_emb.word_embeddings.weight'] = gt[0]
    grads['audio_emb.word_embeddings.weight'] = ga[0]
    if reducer is not None:
        reducer.submit(len(tr.layers), grads)      # the remaining parameters (embeddings, proj)
        grads = reducer.finish()
    return loss, grads


@torch.no_grad()
def nar_loss_and_grads(model, batch: dict, layer: int, precision: str, drop: dict | None = None):
    """ValleNAR.training_step for a fixed stage ``layer`` (valle_nar.py:53-105 with repairs A-1..A-3, A-4 kept)."""
    cfg, dev = model.config, model.device
    cd = _cd(precision)
    codes, tokens = batch['codes'].to(dev), batch['tokens'].to(dev)
    B, T, Q = codes.shape
    Tx = int(batch['tokens_lens'].max())
    tokens = tokens[:, :Tx]
    prefix_len = min(T // 3, 3 * cfg.quantization_factor)
    S, d, V = Tx + T, cfg.d_model, cfg.num_audio_tokens
    R, Rp = B * S, _pad8(B * S)
    stage_embs = [m.weight for m in model.stage_embs]
    tr = StackTrainer(model.transformer, cfg.n_heads, precision, cfg.norm, stage_embs)
    tok_table = model.tokens_emb.weight.detach().float().unsqueeze(0).contiguous()
    code_tables = torch.stack([m.weight.detach().float() for m in model.codes_embs]).contiguous()
    pe_t = model.tokens_position_emb.pe.detach().float().reshape(-1, d).contiguous()
    pe_a = model.audio_position_emb.pe.detach().float().reshape(-1, d).contiguous()
    tok_i, cod_i = _i32(tokens, dev).view(B, Tx, 1), _i32(codes, dev)
    _ids_in_range(tok_i, tok_table.shape[1], 'ValleNAR.training_step tokens')
    _ids_in_range(cod_i, V, 'ValleNAR.training_step codes')            # embedding ids and the stage's targets alike
    x = torch.zeros(Rp, d, device=dev, dtype=torch.float32)
    ops.embed_sum_pe(tok_i, tok_table, pe_t, x, out_rows_per_batch=S)
    ops.embed_sum_pe(cod_i, code_tables, pe_a, x, t_split=prefix_len, nq_a=Q, nq_b=layer, out_rows_per_batch=S, out_row_offset=Tx)
    if drop and drop['pe_p'] > 0.0:
        ops.dropout_(x[:R], drop['pe_p'], drop['seed'], SITE_PE)
    cache = tr.forward(x, B, S, mask_mode=MASK_NONE, x_lens=None, kv_lens=None, stage=layer - 1, drop=drop)   # padding ignored (A-4)
    Tt = T - prefix_len
    Ra, Rap = B * Tt, _pad8(B * Tt)
    rows = x[:R].view(B, S, d)[:, Tx + prefix_len:].reshape(Ra, d).contiguous()
    proj = _Lin(model.proj_layers[layer - 1].weight, None, cd, pad_n=True)
    Vp = proj.w.shape[0]
    hb = torch.zeros(Rap, d, device=dev, dtype=cd)
    ops.residual_layernorm(rows, None, None, hb[:Ra])
    logits = torch.zeros(Rap, Vp, device=dev, dtype=torch.float32)
    ops.linear(hb[:Ra], proj.w, out=logits[:Ra])
    target = _i32(codes[:, prefix_len:, layer], dev).reshape(-1)
    dlogits = torch.zeros(Rap, Vp, device=dev, dtype=torch.float32)
    loss_rows = ops.cross_entropy(logits[:Ra], target, V, dlogits=dlogits[:Ra], scale=1.0 / Ra)
    loss = ops.colsum(loss_rows.view(Ra, 1), scale=1.0 / Ra)[0]
    grads = {}
    dl = torch.zeros(Rap, Vp, device=dev, dtype=cd)
    ops.residual_layernorm(dlogits[:Ra], None, None, dl[:Ra])
    grads[f'proj_layers.{layer - 1}.weight'] = _wgrad(dl, hb, Ra)[:V]
    dh = torch.zeros(Rap, d, device=dev, dtype=cd)
    _dgrad(dl[:Ra], proj, dh[:Ra])
    dx = torch.zeros(Rp, d, device=dev, dtype=torch.float32)
    dx[:R].view(B, S, d)[:, Tx + prefix_len:] = dh[:Ra].view(B, Tt, d).float()
    stage_key = f'stage_embs.{layer - 1}.word_embeddings.weight'
    reducer = parallel.active_reducer()

    def layer_done(li, g):
        out = {}
        _layer_param_grads(f'transformer.layers.{li}.', model.transformer.layers[li], g, cfg.norm, out, stage_key)
        de = out.pop(stage_key, None)              # the stage embedding collects a term from every AdaLN: last bucket
        if de is not None:
            grads[stage_key] = grads.get(stage_key, 0) + de
        grads.update(out)
        if reducer is not None:
            reducer.submit(li, out)

    if reducer is not None:
        reducer.begin()
    tr.backward(dx, cache, B, S, mask_mode=MASK_NONE, x_lens=None, kv_lens=None, drop=drop, on_layer=layer_done)
    if drop and drop['pe_p'] > 0.0:
        ops.dropout_(dx[:R], drop['pe_p'], drop['seed'], SITE_PE)
    gt = torch.zeros_like(tok_table)
    gc = torch.zeros_like(code_tables)
    ops.embed_bwd(tok_i, dx, gt, rows_per_batch=S)
    ops.embed_bwd(cod_i, dx, gc, t_split=prefix_len, nq_a=Q, nq_b=layer, rows_per_batch=S, row_offset=Tx)
    grads['tokens_emb.word_embeddings.weight'] = gt[0]
    for q in range(Q):
        # a table contributes when its codebook is summed somewhere: all Q over the prompt prefix, the first `layer` over the rest
        if q < layer or prefix_len > 0:
            grads[f'codes_embs.{q}.word_embeddings.weight'] = gc[q]
    if reducer is not None:
        reducer.submit(len(tr.layers), grads)
        grads = reducer.finish()
    return loss, grads


class _StepFn(torch.autograd.Function):
    """loss = f(parameters): forward AND gradients are computed by the CUDA stack in ``forward``; ``backward`` hands the
    stored gradients (times the incoming scalar) to autograd, so ``loss.backward()`` fills ``param.grad``."""

    @staticmethod
    def forward(ctx, fn, names, *params):
        loss, grads = fn()
        ctx.grads = [grads.get(n) for n in names]
        ctx.shapes = [p.shape for p in params]
        return loss.clone()

    @staticmethod
    def backward(ctx, gout):
        outs = []
        for g, shp in zip(ctx.grads, ctx.shapes):
            outs.append(None if g is None else (g.reshape(shp) * gout))
        return (None, None, *outs)


def step_loss(model, fn):
    """Differentiable loss of one training step: ``fn() -> (loss, {name: grad})`` runs on the CUDA stack."""
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    names = [n for n, _ in named]
    return _StepFn.apply(fn, names, *[p for _, p in named])
