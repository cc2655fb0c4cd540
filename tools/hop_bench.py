"""Cost of one dependent hop in a PDL chain inside a CUDA graph: a trivial kernel, the decode LN / GELU-reduce kernels and
the decode GEMMs at BASELINE shapes (B = argv[1], default 32).   python tools/hop_bench.py [B]"""
import sys

import torch

sys.path.insert(0, '.')
from valle2_b200 import ops  # noqa: E402

dev, bf = 'cuda', torch.bfloat16
B, d, F = (int(sys.argv[1]) if len(sys.argv) > 1 else 32), 1024, 4096


def timed(fn, n, reps=5):
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
    return e0.elapsed_time(e1) / reps * 1e3 / n


x1 = torch.randn(1, 64, device=dev)
y1 = torch.zeros(1, 64, device=dev, dtype=bf)
print('trivial kernel (LN of one 64-wide row)      %.2f us/hop' % timed(lambda: [ops.residual_layernorm(x1, None, None, y1) for _ in range(100)], 100))
x = torch.randn(B, d, device=dev)
h = torch.zeros(B, d, device=dev, dtype=bf)
g_, b_ = torch.randn(d, device=dev), torch.randn(d, device=dev)
bias = torch.randn(d, device=dev)
for ns in (0, 8, 16) if B <= 128 else (0, 8):
    part = torch.randn(max(ns, 1), B, d, device=dev) * 0.01
    t = timed(lambda: [ops.residual_layernorm(x, g_, b_, h, part=part if ns else None, n_part=ns, part_stride=B * d, bias=bias if ns else None) for _ in range(100)], 100)
    print('LN kernel B=%d d=1024 n_part=%-2d              %.2f us/hop' % (B, ns, t))
nf1 = ops.linear_decode_splits(F, d, 32, B)
pf1 = torch.randn(nf1, B, F, device=dev)
f = torch.zeros(B, F, device=dev, dtype=bf)
b1 = torch.randn(F, device=dev)
print('GELU-reduce kernel B=%d F=4096 n_part=%d      %.2f us/hop' % (B, nf1, timed(lambda: [ops.reduce_bias_act(pf1, nf1, B * F, b1, True, f) for _ in range(100)], 100)))
ws = {k: [(torch.randn(n, kk, device=dev) / 32).to(bf) for _ in range(12)] for k, (n, kk) in {'qkv': (3 * d, d), 'o': (d, d), 'f1': (F, d), 'f2': (d, F)}.items()}
a_d, a_f = torch.randn(B, d, device=dev).to(bf), torch.randn(B, F, device=dev).to(bf)
for k, (n, kk) in {'qkv': (3 * d, d), 'o': (d, d), 'f1': (F, d), 'f2': (d, F)}.items():
    ns = ops.linear_decode_splits(n, kk, 32, B)
    part = torch.zeros(ns, B, n, device=dev)
    a = a_f if kk == F else a_d
    t = timed(lambda: [ops.linear_decode(a, w, part, B * n, 32) for w in ws[k] for _ in range(2)], 24)
    print('GEMM %-4s N=%d K=%d splits %-2d             %.2f us/hop  (%.2f TB/s of weights)' % (k, n, kk, ns, t, n * kk * 2 / t / 1e6))
