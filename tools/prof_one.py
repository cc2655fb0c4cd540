#!/usr/bin/env python
"""Launch ONE kernel configuration a few times (no graphs) -- the target of `ncu --set full` captures.
usage: python tools/prof_one.py {attn_decode|gemm_decode|rows_decode [B]|gemm_large|gemm_argmax|attn_prefill}"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from valle2_b200 import ops  # noqa: E402

what = sys.argv[1]
d, F, H = 1024, 4096, 16
if what == 'attn_decode':            # BASELINE config 2 at the mean context: B=32, ctx=750, 12 distinct layer pools
    B, ctx = 32, 750
    max_pages = (ctx + 64) // 64 + 1
    pools = torch.randn(12, B * max_pages, 2, H, 64, 64, device='cuda').bfloat16()
    bt = torch.arange(B * max_pages, device='cuda', dtype=torch.int32).view(B, max_pages)
    seq = torch.full((B,), ctx, device='cuda', dtype=torch.int32)
    part = torch.randn(6, B, 3 * d, device='cuda')
    o = torch.empty(B, d, device='cuda', dtype=torch.bfloat16)
    ws = torch.zeros(ops.attn_decode_ws_bytes(B, H, 1) // 4 + 64, device='cuda', dtype=torch.int32)
    for li in range(12):
        ops.attn_decode_paged(part, 6, B * 3 * d, pools[li], bt, seq, o, B, H, 64, 1, ws)
elif what == 'gemm_decode':          # the four weight-streaming GEMMs of one decode layer at B = argv[2] (default 32)
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    h = torch.randn(B, d, device='cuda').bfloat16()
    f = torch.randn(B, F, device='cuda').bfloat16()
    for rep in range(3):
        for (N, K, x) in [(3 * d, d, h), (d, d, h), (F, d, h), (d, F, f)]:
            w = (torch.randn(N, K, device='cuda') * 0.02).bfloat16()
            part = torch.zeros(32, B, N, device='cuda')
            ops.linear_decode(x, w, part, B * N, 32)
elif what == 'rows_decode':          # the four GEMMs of the lean small-batch layer (csrc/gemm_decode_mma.cu), B = argv[2] (default 1)
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    x = torch.randn(B, d, device='cuda')
    o = torch.randn(B, d, device='cuda').bfloat16()
    f = torch.zeros(B, F, device='cuda', dtype=torch.bfloat16)
    qkv = torch.zeros(B, 3 * d, device='cuda')
    g_, b_ = torch.randn(d, device='cuda'), torch.randn(d, device='cuda')
    bo, b1 = torch.randn(d, device='cuda'), torch.randn(F, device='cuda')
    for rep in range(3):
        wq, wo = (torch.randn(3 * d, d, device='cuda') * 0.02).bfloat16(), (torch.randn(d, d, device='cuda') * 0.02).bfloat16()
        w1, w2 = (torch.randn(F, d, device='cuda') * 0.02).bfloat16(), (torch.randn(d, F, device='cuda') * 0.02).bfloat16()
        ops.linear_decode_rows_ln(x, wq, qkv, gamma=g_, beta=b_)
        ops.linear_decode_rows(o, wo, x, bias=bo, residual=True)
        ops.linear_decode_rows_ln(x, w1, f, gamma=g_, beta=b_, bias=b1, gelu=True)
        ops.linear_decode_rows(f, w2, x, bias=bo, residual=True, want_split=0)
elif what == 'gemm_large':           # NAR config 3 shapes (M = 64 x 900)
    M = 57600
    for (N, K, epi) in [(3072, 1024, 'none'), (1024, 1024, 'residual'), (4096, 1024, 'gelu'), (1024, 4096, 'residual')]:
        x = torch.randn(M, K, device='cuda').bfloat16()
        w = torch.randn(N, K, device='cuda').bfloat16()
        bias = torch.randn(N, device='cuda')
        if epi == 'residual':
            y = torch.randn(M, N, device='cuda')
            ops.linear(x, w, bias, residual=y, out=y)
        elif epi == 'gelu':
            ops.linear(x, w, bias, gelu=True)
        else:
            ops.linear(x, w)
elif what == 'gemm_argmax':          # a greedy NAR stage's logits projection + pick (batch 64, 750 target frames): unfused, then fused
    M, N, K = 48000, 1024, 1024
    x = torch.randn(M, K, device='cuda').bfloat16()
    w = (torch.randn(N, K, device='cuda') / 32).bfloat16()
    tok0 = torch.empty(M, device='cuda', dtype=torch.int32)
    tok1 = torch.empty(M, device='cuda', dtype=torch.int32)
    keys = torch.zeros(M, device='cuda', dtype=torch.int64)
    logits = ops.linear(x, w, out_dtype=torch.float32)                                        # gemm_tc_kernel<..., EPI_NONE>: 197 MB of logits out
    ops.sample(logits, 1, 0, N, M, N, temperature=1.0, top_k=1, top_p=1.0, out_tok=tok0)     # sample_kernel: 197 MB back in
    ops.linear_argmax(x, w, keys, tok1)                                                       # gemm_tc_kernel<..., EPI_ARGMAX> + argmax_unpack_kernel
    torch.cuda.synchronize()
    assert torch.equal(tok0, tok1)
elif what == 'attn_prefill':         # NAR config 3 attention: B=64, S=900, 16 heads
    B, S = 64, 900
    qkv = torch.randn(B * S, 3 * d, device='cuda').bfloat16()
    o = torch.empty(B * S, d, device='cuda', dtype=torch.bfloat16)
    for _ in range(2):
        ops.attention_packed(qkv, o, B, S, H, mask_mode=ops.MASK_NONE, x_lens=None, kv_lens=None, use_tc=True)
torch.cuda.synchronize()
print('done', what)
