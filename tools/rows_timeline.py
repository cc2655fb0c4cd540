"""In-kernel timeline of the rows decode GEMM (csrc/gemm_decode_mma.cu) inside a PDL chain: stamps per CTA.
    python tools/rows_timeline.py"""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from valle2_b200 import _lib, ops  # noqa: E402

dev, bf = 'cuda', torch.bfloat16
B, d, F = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 1024, 4096      # rows kernels serve M <= 16

lib = _lib.load()
names = ['start', 'w requested', 'dep resolved', 'acts landed', 'mma done', 'summed', 'stored']
cases = {'qkv': (3 * d, d, 1), 'qkv/2': (3 * d, d, 2), 'o': (d, d, 1), 'f1': (F, d, 1), 'f2': (d, F, 1)}
if B <= 8:
    cases.update({'qkv+LN': (3 * d, d, -1), 'f1+LN': (F, d, -1), 'f2 whole K': (d, F, 0)})
g_, b_ = torch.randn(d, device=dev), torch.randn(d, device=dev)
for k, (n, kk, split) in cases.items():
    ws = [(torch.randn(n, kk, device=dev) / 32).to(bf) for _ in range(6)]
    a = torch.randn(B, kk, device=dev).to(bf)
    a32 = torch.randn(B, kk, device=dev)
    ns = ops.linear_decode_rows_splits(kk, max(split, 0), B)
    y = torch.zeros(ns, B, n, device=dev)
    dbgs = [torch.zeros(1024, 16, device=dev, dtype=torch.int64) for _ in range(6)]

    def chain():
        for w, dbg in zip(ws, dbgs):
            _lib.check(lib.vb_linear_decode_rows_set_debug(dbg.data_ptr()), 'dbg')
            if split < 0:
                ops.linear_decode_rows_ln(a32, w, y[0], gamma=g_, beta=b_)
            else:
                ops.linear_decode_rows(a, w, y if ns > 1 else y[0], want_split=split)
        _lib.check(lib.vb_linear_decode_rows_set_debug(None), 'dbg')

    chain()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        chain()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    T = [x.cpu().numpy().astype(np.float64) for x in dbgs]
    grid = int((T[3][:, 0] > 0).sum())
    print(f'== {k}: N={n} K={kk} splits {ns} grid {grid}')
    prev_end = T[2][:grid, 6].max()
    t = T[3][:grid]
    for i, nm in enumerate(names):
        col = t[:, i] - prev_end
        print(f'   {nm:13s} min {col.min() / 1e3:7.2f}  med {np.median(col) / 1e3:7.2f}  max {col.max() / 1e3:7.2f}  us after the previous kernel\'s last store')
    cyc = T[3][:grid, 8:16]
    for i in range(0, 6):
        dcy = cyc[:, i + 1] - cyc[:, i]
        print(f'   cycles {names[i]:13s} -> {names[i + 1]:13s} min {dcy.min():7.0f}  med {np.median(dcy):7.0f}  max {dcy.max():7.0f}')
    print(f'   kernel-to-kernel period {(T[4][:grid, 6].max() - T[3][:grid, 6].max()) / 1e3:.2f} us')
