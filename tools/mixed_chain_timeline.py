"""The lean layer's four GEMMs as a mixed PDL chain (o -> FFN1+LN -> FFN2 whole K -> QKV+LN, repeated), each launch with its
own stamp buffer: does kernel n+1 start early when n is a DIFFERENT kernel?   python tools/mixed_chain_timeline.py [B] [same]"""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from valle2_b200 import _lib, ops  # noqa: E402

dev, bf = 'cuda', torch.bfloat16
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
same = len(sys.argv) > 2
d, F = 1024, 4096
lib = _lib.load()
x = torch.randn(B, d, device=dev)
o = torch.randn(B, d, device=dev).to(bf)
f = torch.zeros(B, F, device=dev, dtype=bf)
qkv = torch.zeros(B, 3 * d, device=dev)
g_, b_ = torch.randn(d, device=dev), torch.randn(d, device=dev)
bo, b1 = torch.randn(d, device=dev), torch.randn(F, device=dev)
W = [{'q': (torch.randn(3 * d, d, device=dev) / 32).to(bf), 'o': (torch.randn(d, d, device=dev) / 32).to(bf),
      '1': (torch.randn(F, d, device=dev) / 32).to(bf), '2': (torch.randn(d, F, device=dev) / 64).to(bf)} for _ in range(4)]
pool = [torch.zeros(512, 16, device=dev, dtype=torch.int64) for _ in range(64)]   # allocated outside the capture
bufs, names = [], []


def dbg(name):
    buf = pool[len(bufs)]
    bufs.append(buf)
    names.append(name)
    _lib.check(lib.vb_linear_decode_rows_set_debug(buf.data_ptr()), 'dbg')


def chain():
    for w in W:
        dbg('o'); ops.linear_decode_rows(o, w['o'], x, bias=bo, residual=True)
        if same:
            dbg('o'); ops.linear_decode_rows(o, w['o'], x, bias=bo, residual=True)
            dbg('o'); ops.linear_decode_rows(o, w['o'], x, bias=bo, residual=True)
            dbg('o'); ops.linear_decode_rows(o, w['o'], x, bias=bo, residual=True)
            continue
        dbg('f1+LN'); ops.linear_decode_rows_ln(x, w['1'], f, gamma=g_, beta=b_, bias=b1, gelu=True)
        dbg('f2 whole K'); ops.linear_decode_rows(f, w['2'], x, bias=bo, residual=True, want_split=0)
        dbg('qkv+LN'); ops.linear_decode_rows_ln(x, w['q'], qkv, gamma=g_, beta=b_)
    _lib.check(lib.vb_linear_decode_rows_set_debug(None), 'dbg')


chain()
torch.cuda.synchronize()
bufs.clear(); names.clear()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    chain()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
t0, prev_end = None, None
for nm, b in zip(names, bufs):
    t = b.cpu().numpy().astype(np.float64)
    t = t[t[:, 0] > 0]
    if t0 is None:
        t0 = t[:, 0].min()
    print(f'  {nm:12s} start {(t[:, 0].min() - t0) / 1e3:7.2f}  dep {(np.median(t[:, 2]) - t0) / 1e3:7.2f}  end {(t[:, 6].max() - t0) / 1e3:7.2f}'
          f'   dep - prev end {((np.median(t[:, 2]) - prev_end) / 1e3 if prev_end else 0):5.2f}')
    prev_end = t[:, 6].max()
