"""Thin torch-tensor wrappers over the C ABI (include/valle_b200.h).

PyTorch is used here only for device memory and streams; every computation below is a kernel of
libvalle_b200.so launched on ``torch.cuda.current_stream()``.  All wrappers raise on CPU tensors:
there is no fallback path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (FLAG_ATTN_SIMT, FLAG_ATTN_TICKET, FLAG_DG_GLOBAL, FLAG_LATE_TRIGGER, FLAG_PREFETCH_KV, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_NONE, MASK_EXPLICIT, MASK_NONE,
                   MASK_PREFIX_LM, VB_BF16, VB_F32, check)

__all__ = ['EPI_NONE', 'EPI_BIAS', 'EPI_BIAS_GELU', 'EPI_BIAS_RESIDUAL', 'MASK_NONE', 'MASK_PREFIX_LM',
           'MASK_EXPLICIT', 'VB_F32', 'VB_BF16']

PAGE = 64


def _L():
    return _lib.load()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _code(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return VB_F32
    if dtype == torch.bfloat16:
        return VB_BF16
    raise TypeError(f'unsupported dtype {dtype}; the CUDA path handles float32 and bfloat16')


def _ptr(t: torch.Tensor | None) -> int | None:
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.VBError('valle2_b200 has no CPU path: tensor is not on a CUDA device')
    return t.data_ptr()


def device_info() -> dict:
    sm, maj, mn, smem = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
    check(_L().vb_device_info(C.byref(sm), C.byref(maj), C.byref(mn), C.byref(smem)), 'vb_device_info')
    return {'sm_count': sm.value, 'cc': (maj.value, mn.value), 'smem_optin': smem.value}


def embed_sum_pe(ids: torch.Tensor, tables: torch.Tensor, pe: torch.Tensor, out: torch.Tensor, *,
                 t_split: int = 0, nq_a: int | None = None, nq_b: int | None = None, pos_offset: int = 0,
                 pos_b: torch.Tensor | None = None, out_rows_per_batch: int | None = None,
                 out_row_offset: int = 0, norm_y: torch.Tensor | None = None, gamma: torch.Tensor | None = None,
                 beta: torch.Tensor | None = None, eps: float = 1e-5) -> None:
    """ids int32 (B,T,Q); tables fp32 (Q,V,d); pe fp32 (max_len,d); out fp32 rows of d.  With ``norm_y`` (rows of d, bf16 or
    fp32, indexed like out) the same kernel also writes LayerNorm(out row; gamma, beta, eps) -- a plain cast when gamma is
    None -- from registers (vb_embed_sum_pe_norm, d in {256, 512, 1024})."""
    B, T, Q = ids.shape
    Qt, V, d = tables.shape
    assert Qt == Q and ids.dtype == torch.int32 and ids.is_contiguous() and tables.is_contiguous()
    assert tables.dtype == torch.float32 and pe.dtype == torch.float32 and out.dtype == torch.float32
    nq_b = Q if nq_b is None else nq_b
    nq_a = nq_b if nq_a is None else nq_a
    rows = T if out_rows_per_batch is None else out_rows_per_batch
    if norm_y is not None:
        assert norm_y.is_contiguous() and norm_y.shape[-1] == d and norm_y.numel() == out.numel()
        for t_ in (gamma, beta):
            assert t_ is None or (t_.dtype == torch.float32 and t_.is_contiguous() and t_.numel() == d)
        check(_L().vb_embed_sum_pe_norm(_ptr(ids), _ptr(tables), _ptr(pe), _ptr(out), B, T, Q, V, d, t_split, nq_a, nq_b,
                                        pos_offset, _ptr(pos_b), pe.shape[0], rows, out_row_offset, _ptr(gamma), _ptr(beta),
                                        eps, _ptr(norm_y), _code(norm_y.dtype), _stream()), 'vb_embed_sum_pe_norm')
        return
    check(_L().vb_embed_sum_pe(_ptr(ids), _ptr(tables), _ptr(pe), _ptr(out), B, T, Q, V, d, t_split, nq_a, nq_b,
                               pos_offset, _ptr(pos_b), pe.shape[0], rows, out_row_offset, _stream()),
          'vb_embed_sum_pe')


def residual_layernorm(x: torch.Tensor, gamma: torch.Tensor | None, beta: torch.Tensor | None,
                       y: torch.Tensor | None, *, part: torch.Tensor | None = None, n_part: int = 0,
                       part_stride: int = 0, bias: torch.Tensor | None = None, eps: float = 1e-5) -> None:
    """x fp32 (R,d) updated in place when n_part > 0; y = LN(x) (or a plain cast when gamma is None)."""
    assert x.dtype == torch.float32 and x.is_contiguous()
    R, d = x.shape
    check(_L().vb_residual_layernorm(_ptr(x), _ptr(part), n_part, part_stride, _ptr(bias), _ptr(gamma), _ptr(beta),
                                     _ptr(y), _code(y.dtype) if y is not None else VB_F32, R, d, eps, _stream()),
          'vb_residual_layernorm')


def reduce_bias_act(part: torch.Tensor, n_part: int, part_stride: int, bias: torch.Tensor | None, gelu: bool,
                    y: torch.Tensor) -> None:
    R, N = y.shape
    check(_L().vb_reduce_bias_act(_ptr(part), n_part, part_stride, _ptr(bias), int(gelu), _ptr(y), _code(y.dtype),
                                  R, N, _stream()), 'vb_reduce_bias_act')


def linear(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, *, gelu: bool = False,
           residual: torch.Tensor | None = None, out: torch.Tensor | None = None,
           out_dtype: torch.dtype | None = None) -> torch.Tensor:
    """y = epilogue(x @ w.T); x (M,K), w (N,K) same dtype (fp32 -> SIMT, bf16 -> tcgen05)."""
    assert x.dim() == 2 and w.dim() == 2 and x.shape[1] == w.shape[1], (x.shape, w.shape)
    assert x.stride(1) == 1 and w.stride(1) == 1
    M, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=out_dtype or x.dtype)
    assert out.shape == (M, N) and out.stride(1) == 1
    if residual is not None:
        assert bias is not None and not gelu and residual.dtype == torch.float32 and residual.stride(1) == 1
        epi = EPI_BIAS_RESIDUAL
    elif gelu:
        assert bias is not None
        epi = EPI_BIAS_GELU
    elif bias is not None:
        epi = EPI_BIAS
    else:
        epi = EPI_NONE
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
    check(_L().vb_linear(_ptr(x), _code(x.dtype), x.stride(0), _ptr(w), _code(w.dtype), w.stride(0), _ptr(bias),
                         _ptr(residual), residual.stride(0) if residual is not None else 0, _ptr(out),
                         _code(out.dtype), out.stride(0), M, N, K, epi, _stream()), 'vb_linear')
    return out


def linear_argmax_ok(M: int, N: int) -> bool:
    """Shapes the fused logits + greedy-pick GEMM takes (the CTA-pair kernel: M >= 1024, N > 128)."""
    return M >= 1024 and N > 128


def linear_argmax(x: torch.Tensor, w: torch.Tensor, keys: torch.Tensor, out_tok: torch.Tensor, *, rows_per_batch: int | None = None,
                  batch_stride: int | None = None, row_stride: int = 1, temperature: float | None = None, seed: int = 0,
                  step: int = 0) -> None:
    """out_tok[...] = argmax_n (x @ w.T)[m, n] (lowest index on ties) without materialising the logits (vb_linear_argmax);
    with ``temperature`` a draw from Categorical(softmax(logits / temperature)) instead (vb_linear_categorical: Gumbel-max in the
    same epilogue, noise a pure function of (seed, step, row, column)):
    x (M,K), w (N,K) bf16; keys (M,) int64 scratch, zero before the first call (the call leaves it zero again); out_tok int32,
    row m = b * rows_per_batch + t goes to element b * batch_stride + t * row_stride (default: dense (M,))."""
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.stride(1) == 1 and w.stride(1) == 1
    assert x.shape[1] == w.shape[1], (x.shape, w.shape)
    M, K = x.shape
    N = w.shape[0]
    assert keys.dtype == torch.int64 and keys.is_contiguous() and keys.numel() >= M and out_tok.dtype == torch.int32
    rpb = rows_per_batch or max(M, 1)
    bs = batch_stride if batch_stride is not None else rpb * row_stride
    if temperature is None:
        check(_L().vb_linear_argmax(_ptr(x), x.stride(0), _ptr(w), w.stride(0), _ptr(keys), _ptr(out_tok), rpb, bs, row_stride,
                                    M, N, K, _stream()), 'vb_linear_argmax')
    else:
        check(_L().vb_linear_categorical(_ptr(x), x.stride(0), _ptr(w), w.stride(0), _ptr(keys), _ptr(out_tok), rpb, bs, row_stride,
                                         M, N, K, float(temperature), int(seed) & ((1 << 64) - 1), int(step), _stream()),
              'vb_linear_categorical')


def linear_t(x: torch.Tensor, w: torch.Tensor, *, x_t: bool = False, w_t: bool = False, bias: torch.Tensor | None = None,
             out: torch.Tensor | None = None, out_dtype: torch.dtype | None = None) -> torch.Tensor:
    """y (M, N) = X @ W.T (+ bias) on the tcgen05 GEMM with operands optionally stored transposed (no transpose kernel):
    x_t: x is (K, M) instead of (M, K); w_t: w is (K, N) instead of (N, K).  bf16 operands, N > 128."""
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.stride(1) == 1 and w.stride(1) == 1
    K, M = (x.shape[0], x.shape[1]) if x_t else (x.shape[1], x.shape[0])
    Kw, N = (w.shape[0], w.shape[1]) if w_t else (w.shape[1], w.shape[0])
    assert K == Kw, (x.shape, w.shape, x_t, w_t)
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=out_dtype or x.dtype)
    assert out.shape == (M, N) and out.stride(1) == 1
    check(_L().vb_linear_t(_ptr(x), x.stride(0), int(x_t), _ptr(w), w.stride(0), int(w_t), _ptr(bias), None, 0, _ptr(out),
                           _code(out.dtype), out.stride(0), M, N, K, EPI_BIAS if bias is not None else EPI_NONE, _stream()),
          'vb_linear_t')
    return out


def linear_decode_splits(N: int, K: int, max_split: int, M: int | None = None) -> int:
    """Split-K slice count vb_linear_decode uses for an (M, K) x (N, K)^T product (M = None: any batch up to 128 rows)."""
    if M is None:
        return int(_L().vb_linear_decode_splits(N, K, max_split))
    return int(_L().vb_linear_decode_splits_m(M, N, K, max_split))


def linear_decode(x: torch.Tensor, w: torch.Tensor, part: torch.Tensor, part_stride: int, max_split: int,
                  flags: int = 0) -> int:
    """part[s][m][n] (fp32) = split-K slices of x @ w.T; x (M<=1024,K) bf16, w (N,K) bf16.  Returns n_split."""
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and part.dtype == torch.float32
    M, K = x.shape
    N = w.shape[0]
    ns = C.c_int()
    check(_L().vb_linear_decode(_ptr(x), x.stride(0), _ptr(w), w.stride(0), _ptr(part), part_stride, M, N, K,
                                max_split, flags, C.byref(ns), _stream()), 'vb_linear_decode')
    return ns.value


def linear_decode_rows_splits(K: int, want_split: int = 1, M: int = 32) -> int:
    """Split count vb_linear_decode_rows uses for this K and batch (0: K unsupported).  want_split=0: whole K in one CTA
    when the M activation rows fit in shared memory (K <= 4096)."""
    return int(_L().vb_linear_decode_rows_splits(M, K, want_split))


def _rows_epilogue(bias, gelu, residual, y):
    if residual:
        assert bias is not None and not gelu and y.dtype == torch.float32
        return EPI_BIAS_RESIDUAL
    if gelu:
        assert bias is not None
        return EPI_BIAS_GELU
    return EPI_BIAS if bias is not None else EPI_NONE


def linear_decode_rows(x: torch.Tensor, w: torch.Tensor, y: torch.Tensor, *, bias: torch.Tensor | None = None,
                       gelu: bool = False, residual: bool = False, want_split: int = 1, flags: int = 0) -> int:
    """Decode-shape linear layer, "rows" form (csrc/gemm_decode_mma.cu): x (M<=32, K) bf16, w (N, K) bf16.
    One slice: y (M, N) fp32/bf16 = epilogue(x @ w.T); residual=True: y (fp32) += x @ w.T + bias in place.
    Several slices (K > 1024 or want_split > 1): y (n_split, M, N) fp32 slices, no epilogue.  Returns n_split."""
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.stride(1) == 1 and w.stride(1) == 1
    M, K = x.shape
    N = w.shape[0]
    epi = _rows_epilogue(bias, gelu, residual, y)
    if y.dim() == 3:
        assert y.shape[1:] == (M, N) and y.dtype == torch.float32 and y.stride(2) == 1
        ldy, sstride = y.stride(1), y.stride(0)
        assert y.shape[0] >= linear_decode_rows_splits(K, want_split, M)
    else:
        assert y.shape == (M, N) and y.stride(1) == 1
        ldy, sstride = y.stride(0), 0
        assert linear_decode_rows_splits(K, want_split, M) == 1, 'split-K output needs a (n_split, M, N) fp32 buffer'
    ns = C.c_int()
    check(_L().vb_linear_decode_rows(_ptr(x), x.stride(0), _ptr(w), w.stride(0), _ptr(bias), _ptr(y), _code(y.dtype), ldy,
                                     sstride, M, N, K, epi, want_split, flags, C.byref(ns), _stream()), 'vb_linear_decode_rows')
    return ns.value


def linear_decode_rows_ln(x: torch.Tensor, w: torch.Tensor, y: torch.Tensor, *, gamma: torch.Tensor | None = None,
                          beta: torch.Tensor | None = None, eps: float = 1e-5, bias: torch.Tensor | None = None,
                          gelu: bool = False, residual: bool = False, flags: int = 0) -> torch.Tensor:
    """y (M<=8, N) = epilogue(LayerNorm(x; gamma, beta) @ w.T) with the LayerNorm computed on load inside the GEMM kernel;
    x (M, K) fp32 residual rows, K in {256, 512, 1024}; gamma=None: plain bf16 cast of x."""
    assert x.dtype == torch.float32 and w.dtype == torch.bfloat16 and x.stride(1) == 1 and w.stride(1) == 1
    M, K = x.shape
    N = w.shape[0]
    assert y.shape == (M, N) and y.stride(1) == 1
    epi = _rows_epilogue(bias, gelu, residual, y)
    check(_L().vb_linear_decode_rows_ln(_ptr(x), x.stride(0), _ptr(gamma), _ptr(beta), float(eps), _ptr(w), w.stride(0),
                                        _ptr(bias), _ptr(y), _code(y.dtype), y.stride(0), M, N, K, epi, flags, _stream()),
          'vb_linear_decode_rows_ln')
    return y


DG_PLAIN, DG_LN, DG_LN_GELU, DG_RESIDUAL = 0, 1, 2, 3


def decode_gemm_plan(M: int, N: int, K: int) -> dict:
    """Grid of vb_decode_gemm for this shape: {'tiles', 'n_split', 'ws_bytes'}; RESIDUAL launches write `tiles` statistics
    chunks per row."""
    t, ns, ws = C.c_int(), C.c_int(), C.c_int64()
    check(_L().vb_decode_gemm_plan(M, N, K, C.byref(t), C.byref(ns), C.byref(ws)), 'vb_decode_gemm_plan')
    return {'tiles': t.value, 'n_split': ns.value, 'ws_bytes': ws.value}


def decode_gemm(x: torch.Tensor, w: torch.Tensor, mode: int, *, ws: torch.Tensor, counters: torch.Tensor,
                bias: torch.Tensor | None = None, colsum: torch.Tensor | None = None, stats_in: torch.Tensor | None = None,
                n_chunks_in: int = 0, eps: float = 1e-5, y32: torch.Tensor | None = None, y16: torch.Tensor | None = None,
                xres: torch.Tensor | None = None, stats_out: torch.Tensor | None = None, flags: int = 0) -> None:
    """Decode-shape linear layer finished inside one launch (include/valle_b200.h, vb_decode_gemm): x (M<=256, K) bf16,
    w (N, K) bf16; mode DG_PLAIN / DG_LN / DG_LN_GELU / DG_RESIDUAL."""
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.stride(1) == 1 and w.stride(1) == 1
    M, K = x.shape
    N = w.shape[0]
    assert w.shape[1] == K and ws.dtype == torch.float32 and counters.dtype == torch.int32
    assert ws.numel() * 4 >= decode_gemm_plan(M, N, K)['ws_bytes'], 'exchange workspace too small'
    for t in (bias, colsum, stats_in, stats_out, y32, xres):
        assert t is None or (t.dtype == torch.float32 and t.is_cuda)
    assert y16 is None or y16.dtype == torch.bfloat16
    for t in (y32, y16, xres):
        assert t is None or (t.shape == (M, N) and t.stride(1) == 1)
    check(_L().vb_decode_gemm(_ptr(x), x.stride(0), _ptr(w), w.stride(0), M, N, K, mode, _ptr(bias), _ptr(colsum),
                              _ptr(stats_in), n_chunks_in, float(eps), _ptr(y32), y32.stride(0) if y32 is not None else 0,
                              _ptr(y16), y16.stride(0) if y16 is not None else 0, _ptr(xres),
                              xres.stride(0) if xres is not None else 0, _ptr(stats_out), _ptr(ws), _ptr(counters), flags,
                              _stream()), 'vb_decode_gemm')


def ar_step_tail(logits: torch.Tensor, V: int, *, temperature: float, top_k: int, top_p: float,
                 uniforms: torch.Tensor | None, seed: torch.Tensor, row_offset: int, last: torch.Tensor,
                 sum_logprobs: torch.Tensor, codes_out: torch.Tensor, seq_lens: torch.Tensor, audio_pos: torch.Tensor,
                 state: torch.Tensor, eos: int, table: torch.Tensor, pe: torch.Tensor, x: torch.Tensor, xb: torch.Tensor,
                 stats: torch.Tensor) -> None:
    """sample + bookkeeping + next step's input row in one launch (include/valle_b200.h, vb_ar_step_tail).
    logits fp32 (B, >=V); seed: int64 tensor [1] on the device; state int32 [4]; table fp32 (rows, d); pe fp32 (max_len, d)."""
    B, d = x.shape
    assert logits.dtype == torch.float32 and logits.stride(1) == 1 and seed.dtype == torch.int64 and state.numel() >= 4
    assert state.dtype == torch.int32 and table.dtype == torch.float32 and pe.dtype == torch.float32 and table.shape[1] == d
    assert x.dtype == torch.float32 and xb.dtype == torch.bfloat16 and x.is_contiguous() and xb.is_contiguous()
    assert stats.dtype == torch.float32 and stats.numel() >= 2 * B and table.is_contiguous() and pe.is_contiguous()
    check(_L().vb_ar_step_tail(_ptr(logits), logits.stride(0), V, float(temperature), int(top_k), float(top_p), _ptr(uniforms),
                               _ptr(seed), int(row_offset), _ptr(last), _ptr(sum_logprobs), _ptr(codes_out), codes_out.stride(0),
                               _ptr(seq_lens), _ptr(audio_pos), _ptr(state), B, eos, _ptr(table), _ptr(pe), d, _ptr(x), _ptr(xb),
                               _ptr(stats), _stream()), 'vb_ar_step_tail')


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, *, mask_mode: int = MASK_NONE,
              q_pos0: int = 0, x_lens: torch.Tensor | None = None, kv_lens: torch.Tensor | None = None,
              mask: torch.Tensor | None = None) -> torch.Tensor:
    """q (B,H,Sq,Dh), k/v (B,H,Sk,Dh) strided views with unit last stride; out (B,Sq,H*Dh).
    mask (explicit mode): uint8, broadcastable to (B,H,Sq,Sk), nonzero = masked."""
    B, H, Sq, Dh = q.shape
    Sk = k.shape[2]
    assert q.stride(3) == 1 and k.stride(3) == 1 and v.stride(3) == 1 and out.stride(2) == 1
    assert q.dtype == k.dtype == v.dtype
    m_sb = m_sh = m_sq = 0
    if mask_mode == MASK_EXPLICIT:
        assert mask is not None and mask.dtype == torch.uint8 and mask.dim() == 4 and mask.stride(3) == 1
        assert mask.shape[2] == Sq and mask.shape[3] == Sk
        m_sb = mask.stride(0) if mask.shape[0] > 1 else 0
        m_sh = mask.stride(1) if mask.shape[1] > 1 else 0
        m_sq = mask.stride(2)
    for t in (x_lens, kv_lens):
        assert t is None or (t.dtype == torch.int32 and t.is_contiguous())
    check(_L().vb_attention(_ptr(q), _ptr(k), _ptr(v), _code(q.dtype), q.stride(0), q.stride(1), q.stride(2),
                            k.stride(0), k.stride(1), k.stride(2), v.stride(0), v.stride(1), v.stride(2),
                            _ptr(out), _code(out.dtype), out.stride(0), out.stride(1), B, H, Sq, Sk, Dh, mask_mode,
                            q_pos0, _ptr(x_lens), _ptr(kv_lens), _ptr(mask), m_sb, m_sh, m_sq, _stream()),
          'vb_attention')
    return out


def attention_packed(qkv: torch.Tensor, out: torch.Tensor, B: int, S: int, H: int, *, mask_mode: int,
                     x_lens: torch.Tensor | None, kv_lens: torch.Tensor | None, use_tc: bool = False,
                     lse: torch.Tensor | None = None) -> torch.Tensor:
    """Self-attention over a packed (B*S, 3*H*Dh) qkv buffer; out (B*S, H*Dh).  lse (B, H, S) fp32, tensor-core path only:
    receives the rows' log-sum-exp for attention_bwd."""
    d3 = qkv.shape[1]
    d = d3 // 3
    Dh = d // H
    if use_tc:
        assert lse is None or (lse.dtype == torch.float32 and lse.is_contiguous() and lse.numel() == B * H * S)
        check(_L().vb_attention_prefill_tc(_ptr(qkv), _ptr(out), B, S, H, mask_mode, _ptr(x_lens), _ptr(kv_lens),
                                           _ptr(lse), _stream()), 'vb_attention_prefill_tc')
        return out
    assert lse is None, 'lse is produced by the tensor-core attention only'
    base = qkv.view(B, S, 3, H, Dh)
    q = base[:, :, 0].permute(0, 2, 1, 3)
    k = base[:, :, 1].permute(0, 2, 1, 3)
    v = base[:, :, 2].permute(0, 2, 1, 3)
    attention(q, k, v, out.view(B, S, d), mask_mode=mask_mode, x_lens=x_lens, kv_lens=kv_lens)
    return out


def kv_scatter_paged(qkv: torch.Tensor, pool: torch.Tensor, block_table: torch.Tensor, kv_lens: torch.Tensor,
                     B: int, S: int, H: int, Dh: int) -> None:
    check(_L().vb_kv_scatter_paged(_ptr(qkv), _code(qkv.dtype), _ptr(pool), _code(pool.dtype), _ptr(block_table),
                                   block_table.shape[1], _ptr(kv_lens), B, S, H, Dh, _stream()),
          'vb_kv_scatter_paged')


def attn_decode_ws_bytes(B: int, H: int, n_tsplit: int) -> int:
    return int(_L().vb_attn_decode_ws_bytes(B, H, n_tsplit))


def attn_decode_paged(qkv_part: torch.Tensor, n_part: int, part_stride: int, pool: torch.Tensor,
                      block_table: torch.Tensor, seq_lens: torch.Tensor, out: torch.Tensor, B: int, H: int, Dh: int,
                      n_tsplit: int, ws: torch.Tensor | None, flags: int = 0) -> None:
    check(_L().vb_attn_decode_paged(_ptr(qkv_part), n_part, part_stride, _ptr(pool), _code(pool.dtype),
                                    _ptr(block_table), block_table.shape[1], _ptr(seq_lens), _ptr(out),
                                    _code(out.dtype), B, H, Dh, n_tsplit, flags, _ptr(ws), _stream()),
          'vb_attn_decode_paged')


def kv_prefetch_l2(pool: torch.Tensor, block_table: torch.Tensor, seq_lens: torch.Tensor, B: int, H: int, Dh: int,
                   page_lo_pct: int = 0, page_hi_pct: int = 100) -> None:
    """Hint: pull the cached pages (a percentage range of every sequence's pages) of one layer's pool into L2."""
    check(_L().vb_kv_prefetch_l2(_ptr(pool), _code(pool.dtype), _ptr(block_table), block_table.shape[1], _ptr(seq_lens),
                                 B, H, Dh, page_lo_pct, page_hi_pct, _stream()), 'vb_kv_prefetch_l2')


def sample(logits_part: torch.Tensor, n_part: int, part_stride: int, row_stride: int, R: int, V: int, *,
           temperature: float, top_k: int, top_p: float, out_tok: torch.Tensor,
           out_logprob: torch.Tensor | None = None, uniforms: torch.Tensor | None = None, seed: int = 0,
           step_ptr: torch.Tensor | None = None, row_offset: int = 0) -> None:
    assert out_tok.dtype == torch.int32
    check(_L().vb_sample(_ptr(logits_part), n_part, part_stride, row_stride, R, V, float(temperature), int(top_k),
                         float(top_p), _ptr(uniforms), seed & (2 ** 64 - 1), _ptr(step_ptr), int(row_offset), _ptr(out_tok),
                         _ptr(out_logprob), _stream()), 'vb_sample')


def ar_bookkeeping(sample_tok: torch.Tensor, logprob: torch.Tensor, last: torch.Tensor, sum_logprobs: torch.Tensor,
                   codes_out: torch.Tensor, seq_lens: torch.Tensor, audio_pos: torch.Tensor, state: torch.Tensor,
                   eos: int) -> None:
    B = last.shape[0]
    check(_L().vb_ar_bookkeeping(_ptr(sample_tok), _ptr(logprob), _ptr(last), _ptr(sum_logprobs), _ptr(codes_out),
                                 codes_out.stride(0), _ptr(seq_lens), _ptr(audio_pos), _ptr(state), B, eos,
                                 _stream()), 'vb_ar_bookkeeping')


# ---- training step, backward pass (csrc/train.cu) -----------------------------------------------------------------------
def transpose(src: torch.Tensor, dst: torch.Tensor | None = None) -> torch.Tensor:
    """dst (cols, rows) = src (rows, cols)^T; both with unit inner stride."""
    rows, cols = src.shape
    if dst is None:
        dst = torch.empty(cols, rows, device=src.device, dtype=src.dtype)
    assert src.stride(1) == 1 and dst.stride(1) == 1 and dst.shape == (cols, rows) and dst.dtype == src.dtype
    check(_L().vb_transpose(_ptr(src), _code(src.dtype), rows, cols, src.stride(0), _ptr(dst), dst.stride(0), _stream()),
          'vb_transpose')
    return dst


def colsum(x: torch.Tensor, out: torch.Tensor | None = None, *, accumulate: bool = False, scale: float = 1.0) -> torch.Tensor:
    R, N = x.shape
    if out is None:
        out = torch.zeros(N, device=x.device, dtype=torch.float32)
    assert x.stride(1) == 1 and out.dtype == torch.float32 and out.numel() == N
    if R >= 4096:       # long and narrow: two deterministic levels so that the first one fills the GPU
        rb = int(min(64, max(2, -(-R // 512))))     # <= 512 rows per CTA (64 per warp): enough CTAs to keep HBM busy
        part = torch.empty(rb, N, device=x.device, dtype=torch.float32)
        check(_L().vb_colsum_blocks(_ptr(x), _code(x.dtype), R, N, x.stride(0), _ptr(part), rb, _stream()), 'vb_colsum_blocks')
        x, R = part, rb
    check(_L().vb_colsum(_ptr(x), _code(x.dtype), R, N, x.stride(0), _ptr(out), int(accumulate), float(scale), _stream()),
          'vb_colsum')
    return out


def gelu_fwd(pre: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    assert pre.is_contiguous() and y.is_contiguous() and pre.dtype == y.dtype and pre.numel() == y.numel()
    check(_L().vb_gelu_fwd(_ptr(pre), _code(pre.dtype), _ptr(y), pre.numel(), _stream()), 'vb_gelu_fwd')
    return y


def gelu_bwd(pre: torch.Tensor, dy: torch.Tensor, dpre: torch.Tensor) -> torch.Tensor:
    assert pre.is_contiguous() and dy.is_contiguous() and dpre.is_contiguous() and pre.dtype == dy.dtype == dpre.dtype
    check(_L().vb_gelu_bwd(_ptr(pre), _ptr(dy), _code(pre.dtype), _ptr(dpre), pre.numel(), _stream()), 'vb_gelu_bwd')
    return dpre


def dropout_(x: torch.Tensor, p: float, seed: int, site: int) -> torch.Tensor:
    """In place nn.Dropout(p) with the counter-based mask of (seed, site); the same call on a gradient is its backward."""
    assert x.is_contiguous()
    if p > 0.0:
        check(_L().vb_dropout(_ptr(x), _code(x.dtype), x.numel(), float(p), seed & (2 ** 64 - 1), site, _stream()), 'vb_dropout')
    return x


def dropout_add_(x: torch.Tensor, t: torch.Tensor, p: float, seed: int, site: int) -> torch.Tensor:
    """x (fp32) += dropout(t) with the mask of (seed, site)."""
    assert x.dtype == torch.float32 and x.is_contiguous() and t.is_contiguous() and x.numel() == t.numel()
    check(_L().vb_dropout_add(_ptr(x), _ptr(t), _code(t.dtype), x.numel(), float(p), seed & (2 ** 64 - 1), site, _stream()),
          'vb_dropout_add')
    return x


def layernorm_bwd(x: torch.Tensor, gamma: torch.Tensor | None, dy: torch.Tensor, dx: torch.Tensor, eps: float = 1e-5):
    """dx (fp32, R x d) += LN backward of dy; returns (dgamma, dbeta) fp32 [d] (None, None when gamma is None = plain cast)."""
    R, d = x.shape
    assert x.dtype == torch.float32 and dx.dtype == torch.float32 and x.is_contiguous() and dx.is_contiguous() and dy.is_contiguous()
    if gamma is None:
        check(_L().vb_layernorm_bwd(_ptr(x), None, _ptr(dy), _code(dy.dtype), _ptr(dx), None, None, R, d, eps, _stream()),
              'vb_layernorm_bwd')
        return None, None
    nb = int(_L().vb_layernorm_bwd_blocks(R))
    pg = torch.empty(nb, d, device=x.device, dtype=torch.float32)
    pb = torch.empty(nb, d, device=x.device, dtype=torch.float32)
    check(_L().vb_layernorm_bwd(_ptr(x), _ptr(gamma), _ptr(dy), _code(dy.dtype), _ptr(dx), _ptr(pg), _ptr(pb), R, d, eps,
                                _stream()), 'vb_layernorm_bwd')
    return colsum(pg), colsum(pb)


def attention_bwd(qkv: torch.Tensor, o: torch.Tensor, do: torch.Tensor, dqkv: torch.Tensor, B: int, S: int, H: int, *,
                  mask_mode: int, x_lens: torch.Tensor | None, kv_lens: torch.Tensor | None,
                  lse: torch.Tensor | None = None) -> torch.Tensor:
    """Packed rows: qkv / dqkv (B*S, 3*H*64), o / do (B*S, H*64), all the same dtype.  lse (B, H, S) fp32: the forward's
    log-sum-exp (attention_packed(..., lse=)); without it the backward recomputes it."""
    d = o.shape[1]
    assert qkv.dtype == o.dtype == do.dtype == dqkv.dtype and qkv.is_contiguous() and o.is_contiguous() and do.is_contiguous()
    assert dqkv.is_contiguous() and qkv.shape[1] == 3 * d
    lse_in = lse is not None
    if lse is None:
        lse = torch.empty(B, H, S, device=qkv.device, dtype=torch.float32)
    assert lse.dtype == torch.float32 and lse.is_contiguous() and lse.numel() == B * H * S
    delta = torch.empty(B, H, S, device=qkv.device, dtype=torch.float32)
    check(_L().vb_attention_bwd(_ptr(qkv), _ptr(o), _ptr(do), _ptr(dqkv), _code(qkv.dtype), _ptr(lse), _ptr(delta), B, S, H,
                                d // H, mask_mode, _ptr(x_lens), _ptr(kv_lens), int(lse_in), _stream()), 'vb_attention_bwd')
    return dqkv


def cross_entropy(logits: torch.Tensor, target: torch.Tensor, V: int, *, dlogits: torch.Tensor | None = None,
                  scale: float = 1.0) -> torch.Tensor:
    """Per-row losses (fp32 [R]) over the first V columns of logits (R, >=V); dlogits = (softmax - onehot) * scale."""
    R = logits.shape[0]
    assert logits.dtype == torch.float32 and logits.stride(1) == 1 and target.dtype == torch.int32 and target.numel() == R
    loss_rows = torch.empty(R, device=logits.device, dtype=torch.float32)
    check(_L().vb_cross_entropy(_ptr(logits), logits.stride(0), _ptr(target), R, V, _ptr(loss_rows), _ptr(dlogits),
                                dlogits.stride(0) if dlogits is not None else 0, float(scale), _stream()), 'vb_cross_entropy')
    return loss_rows


def embed_bwd(ids: torch.Tensor, dx: torch.Tensor, grad_tables: torch.Tensor, *, t_split: int = 0,
              nq_a: int | None = None, nq_b: int | None = None, rows_per_batch: int | None = None, row_offset: int = 0) -> None:
    """grad_tables (Q,V,d) fp32 += scatter of dx rows (transpose of embed_sum_pe)."""
    B, T, Q = ids.shape
    Qt, V, d = grad_tables.shape
    assert Qt == Q and ids.dtype == torch.int32 and ids.is_contiguous() and grad_tables.is_contiguous()
    assert dx.dtype == torch.float32 and grad_tables.dtype == torch.float32 and dx.is_contiguous()
    nq_b = Q if nq_b is None else nq_b
    nq_a = nq_b if nq_a is None else nq_a
    rows = T if rows_per_batch is None else rows_per_batch
    check(_L().vb_embed_bwd(_ptr(ids), _ptr(dx), _ptr(grad_tables), B, T, Q, V, d, t_split, nq_a, nq_b, rows, row_offset,
                            _stream()), 'vb_embed_bwd')
