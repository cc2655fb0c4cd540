"""One decoder layer's dependent chain WITHOUT attention (LN1 -> QKV -> out-proj -> LN2 -> FFN1 [-> GELU-reduce] -> FFN2), 12
layers with distinct weights inside one CUDA graph, for every mix of the two decode GEMM forms (s = tcgen05 split-K slices,
r = mma.sync rows with fused epilogues).   python tools/layer_chain.py [B]"""
import itertools
import sys

import torch

sys.path.insert(0, '.')
from valle2_b200 import ops  # noqa: E402

dev, bf = 'cuda', torch.bfloat16
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
d, F, L = 1024, 4096, 12
torch.manual_seed(0)
W = [{'qkv': (torch.randn(3 * d, d, device=dev) / 32).to(bf), 'o': (torch.randn(d, d, device=dev) / 32).to(bf),
      'f1': (torch.randn(F, d, device=dev) / 32).to(bf), 'f2': (torch.randn(d, F, device=dev) / 64).to(bf)} for _ in range(L)]
bo, b1, b2 = torch.randn(d, device=dev), torch.randn(F, device=dev), torch.randn(d, device=dev)
g_, be_ = torch.ones(d, device=dev), torch.zeros(d, device=dev)
x = torch.randn(B, d, device=dev)
h = torch.zeros(B, d, device=dev, dtype=bf)
o = torch.randn(B, d, device=dev).to(bf)
f = torch.zeros(B, F, device=dev, dtype=bf)
ns = {k: ops.linear_decode_splits(n, kk, 32) for k, (n, kk) in {'qkv': (3 * d, d), 'o': (d, d), 'f1': (F, d), 'f2': (d, F)}.items()}
P = {k: torch.zeros(ns[k], B, n, device=dev) for k, n in {'qkv': 3 * d, 'o': d, 'f1': F, 'f2': d}.items()}
r_qkv = torch.zeros(1, B, 3 * d, device=dev)
r_f2 = torch.zeros(4, B, d, device=dev)


def layer(w, mode, first):
    # LN1 (+ previous FFN2 slices)
    if first:
        ops.residual_layernorm(x, g_, be_, h)
    elif mode['f2'] == 's':
        ops.residual_layernorm(x, g_, be_, h, part=P['f2'], n_part=ns['f2'], part_stride=B * d, bias=b2)
    else:
        ops.residual_layernorm(x, g_, be_, h, part=r_f2, n_part=4, part_stride=B * d, bias=b2)
    if mode['qkv'] == 's':
        ops.linear_decode(h, w['qkv'], P['qkv'], B * 3 * d, 32)
    else:
        ops.linear_decode_rows(h, w['qkv'], r_qkv[0])
    # (attention would run here)
    if mode['o'] == 's':
        ops.linear_decode(o, w['o'], P['o'], B * d, 32)
        ops.residual_layernorm(x, g_, be_, h, part=P['o'], n_part=ns['o'], part_stride=B * d, bias=bo)
    else:
        ops.linear_decode_rows(o, w['o'], x, bias=bo, residual=True)
        ops.residual_layernorm(x, g_, be_, h)
    if mode['f1'] == 's':
        ops.linear_decode(h, w['f1'], P['f1'], B * F, 32)
        ops.reduce_bias_act(P['f1'], ns['f1'], B * F, b1, True, f)
    else:
        ops.linear_decode_rows(h, w['f1'], f, bias=b1, gelu=True)
    if mode['f2'] == 's':
        ops.linear_decode(f, w['f2'], P['f2'], B * d, 32)
    else:
        ops.linear_decode_rows(f, w['f2'], r_f2)


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


print(f'B={B}: us per layer (chain without attention)')
for combo in itertools.product('sr', repeat=4):
    mode = dict(zip(('qkv', 'o', 'f1', 'f2'), combo))
    x.normal_()
    t = timed(lambda: [layer(W[i], mode, i == 0) for i in range(L)]) / L
    print('  qkv=%s o=%s f1=%s f2=%s   %.2f' % (*combo, t))
