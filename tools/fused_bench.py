"""Micro-benchmark of the decode GEMM variants: 24 back-to-back launches (distinct weights) captured in a CUDA graph.
    python tools/fused_bench.py"""
import sys

import torch

sys.path.insert(0, '.')
from valle2_b200 import ops  # noqa: E402

dev, bf = 'cuda', torch.bfloat16
torch.manual_seed(0)
B = 32


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def case(N, K, label):
    nl = 24
    ws = [(torch.randn(N, K, device=dev) / 32).to(bf) for _ in range(nl)]
    a = torch.randn(B, K, device=dev).to(bf)
    x32 = torch.randn(B, K, device=dev)
    g, be = torch.randn(K, device=dev), torch.randn(K, device=dev)
    bias = torch.randn(N, device=dev)
    y = torch.zeros(B, N, device=dev)
    yb = torch.zeros(B, N, device=dev, dtype=bf)
    ns = ops.linear_decode_splits(N, K, 32)
    part = torch.zeros(ns, B, N, device=dev)
    res = {}
    res['split-k slices'] = timed(lambda: [ops.linear_decode(a, w, part, B * N, 32) for w in ws]) / nl
    for c in (0, 1, 2, 4, 8, 16):
        try:
            res[f'fused bf16 c={c}'] = timed(lambda: [ops.linear_decode_fused(a, w, y, cluster_k=c) for w in ws]) / nl
        except Exception as e:
            res[f'fused bf16 c={c}'] = str(e)[-60:]
    c0 = ops.linear_decode_fused_cluster(N, K)
    res[f'fused bf16 resid c={c0}'] = timed(lambda: [ops.linear_decode_fused(a, w, y, bias=bias, residual=True) for w in ws]) / nl
    if K % 64 == 0:
        res[f'fused LN c={c0}'] = timed(lambda: [ops.linear_decode_fused(x32, w, y, gamma=g, beta=be) for w in ws]) / nl
        res[f'fused LN+gelu c={c0}'] = timed(lambda: [ops.linear_decode_fused(x32, w, yb, bias=bias, gelu=True, gamma=g, beta=be) for w in ws]) / nl
        res[f'fused cast c={c0}'] = timed(lambda: [ops.linear_decode_fused(x32, w, y) for w in ws]) / nl
    print(label, f'N={N} K={K} auto cluster {c0} splits {ns}')
    for k, v in res.items():
        print(f'   {k:28s} {v if isinstance(v, str) else round(v, 2)}')


case(3072, 1024, 'QKV')
case(1024, 1024, 'Wo')
case(4096, 1024, 'W1')
case(1024, 4096, 'W2')
case(1024, 512, 'small')
