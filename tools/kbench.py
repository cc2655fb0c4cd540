#!/usr/bin/env python
"""Kernel micro-benchmarks (CUDA events around CUDA-graph replays, so host launch cost is excluded).
Usage: python tools/kbench.py [gemm] [decode] [nar] ...   -> JSON lines to stdout."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from valle2_b200 import ops  # noqa: E402

PEAK_TF, PEAK_GB = 1383.8, 6542.4
try:
    pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))
    PEAK_TF, PEAK_GB = pk['bf16_tflops_sustained'], pk['hbm_gbs']
except Exception:
    pass


def time_graph(fn, reps=20, inner=1):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(inner):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * inner)   # ms per fn()


def bench_gemm():
    M = 57600
    for (N, K, epi) in [(3072, 1024, 'none'), (1024, 1024, 'residual'), (4096, 1024, 'gelu'), (1024, 4096, 'residual'), (1024, 1024, 'none')]:
        x = torch.randn(M, K, device='cuda').bfloat16()
        w = torch.randn(N, K, device='cuda').bfloat16()
        bias = torch.randn(N, device='cuda')
        if epi == 'residual':
            y = torch.randn(M, N, device='cuda')
            fn = lambda: ops.linear(x, w, bias, residual=y, out=y)
        elif epi == 'gelu':
            y = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
            fn = lambda: ops.linear(x, w, bias, gelu=True, out=y)
        else:
            y = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
            fn = lambda: ops.linear(x, w, out=y)
        ms = time_graph(fn, reps=10)
        tf = 2 * M * N * K / (ms * 1e-3) / 1e12
        ref = lambda: torch.matmul(x, w.t())
        ms_ref = time_graph(ref, reps=10)
        print(json.dumps({'bench': 'gemm', 'M': M, 'N': N, 'K': K, 'epi': epi, 'ms': ms, 'tflops': tf,
                          'frac_sustained_peak': tf / PEAK_TF, 'cublas_ms': ms_ref,
                          'cublas_tflops': 2 * M * N * K / (ms_ref * 1e-3) / 1e12}), flush=True)
        del x, w, y


def bench_decode():
    d, F, H = 1024, 4096, 16
    for B in (1, 32):
        h = torch.randn(B, d, device='cuda').bfloat16()
        f = torch.randn(B, F, device='cuda').bfloat16()
        for name, (N, K, xin) in {'qkv': (3 * d, d, h), 'o': (d, d, h), 'f1': (F, d, h), 'f2': (d, F, f), 'logits': (1025, d, h)}.items():
            ws = [(torch.randn(N, K, device='cuda') * 0.02).bfloat16() for _ in range(12)]   # 12 layers: no L2 reuse
            part = torch.zeros(32, B, N, device='cuda')
            ns = ops.linear_decode_splits(N, K, 32)

            def fn():
                for w in ws:
                    ops.linear_decode(xin, w, part, B * N, 32)
            ms = time_graph(fn, reps=10) / 12
            print(json.dumps({'bench': 'linear_decode', 'B': B, 'name': name, 'N': N, 'K': K, 'n_split': ns, 'us': ms * 1e3,
                              'gbs': N * K * 2 / (ms * 1e-3) / 1e9, 'frac_hbm': N * K * 2 / (ms * 1e-3) / 1e9 / PEAK_GB}), flush=True)
        # layernorm with partials
        x = torch.randn(B, d, device='cuda')
        g, b = torch.ones(d, device='cuda'), torch.zeros(d, device='cuda')
        y = torch.empty(B, d, device='cuda', dtype=torch.bfloat16)
        for npart in (0, 8, 16):
            part = torch.zeros(max(npart, 1), B, d, device='cuda')
            ms = time_graph(lambda: ops.residual_layernorm(x, g, b, y, part=part if npart else None, n_part=npart,
                                                           part_stride=B * d, bias=b if npart else None), reps=20, inner=10)
            print(json.dumps({'bench': 'residual_layernorm', 'B': B, 'n_part': npart, 'us': ms * 1e3}), flush=True)
        # attention decode over 12 distinct layer pools
        for ctx in (376, 750, 1126):
            max_pages = (ctx + 64) // 64 + 1
            pools = torch.randn(12, B * max_pages, 2, H, 64, 64, device='cuda').bfloat16()
            bt = torch.arange(B * max_pages, device='cuda', dtype=torch.int32).view(B, max_pages)
            seq = torch.full((B,), ctx, device='cuda', dtype=torch.int32)
            part = torch.randn(6, B, 3 * d, device='cuda')
            o = torch.empty(B, d, device='cuda', dtype=torch.bfloat16)
            for nts in sorted({1, 2, 4, 8, 16} & set(range(1, max_pages + 1))):
                if B * H * nts < 128 and nts < 8:
                    continue
                wsb = torch.zeros(ops.attn_decode_ws_bytes(B, H, nts) // 4 + 64, device='cuda', dtype=torch.int32)

                def fn():
                    for li in range(12):
                        ops.attn_decode_paged(part, 6, B * 3 * d, pools[li], bt, seq, o, B, H, 64, nts, wsb)
                ms = time_graph(fn, reps=10) / 12
                byts = B * (ctx + 1) * 2 * d * 2
                print(json.dumps({'bench': 'attn_decode', 'B': B, 'ctx': ctx, 'n_tsplit': nts, 'us': ms * 1e3,
                                  'gbs': byts / (ms * 1e-3) / 1e9, 'frac_hbm': byts / (ms * 1e-3) / 1e9 / PEAK_GB}), flush=True)
            del pools


def bench_attn_prefill():
    H, Dh = 16, 64
    d = H * Dh
    for (B, S, mode) in [(64, 900, ops.MASK_NONE), (32, 376, ops.MASK_PREFIX_LM)]:
        qkv = torch.randn(B * S, 3 * d, device='cuda').bfloat16()
        o = torch.empty(B * S, d, device='cuda', dtype=torch.bfloat16)
        xl = torch.full((B,), 150, device='cuda', dtype=torch.int32)
        kl = torch.full((B,), S, device='cuda', dtype=torch.int32)
        for use_tc in (True, False):
            try:
                ms = time_graph(lambda: ops.attention_packed(qkv, o, B, S, H, mask_mode=mode, x_lens=xl, kv_lens=kl, use_tc=use_tc), reps=3)
            except Exception as e:
                print(json.dumps({'bench': 'attn_prefill', 'tc': use_tc, 'error': str(e)[:200]}), flush=True)
                continue
            fl = 4 * B * H * S * S * Dh
            print(json.dumps({'bench': 'attn_prefill', 'B': B, 'S': S, 'mode': mode, 'tc': use_tc, 'ms': ms,
                              'tflops': fl / (ms * 1e-3) / 1e12, 'frac_sustained_peak': fl / (ms * 1e-3) / 1e12 / PEAK_TF}), flush=True)


if __name__ == '__main__':
    which = sys.argv[1:] or ['gemm', 'decode', 'attn']
    if 'gemm' in which:
        bench_gemm()
    if 'decode' in which:
        bench_decode()
    if 'attn' in which:
        bench_attn_prefill()
