"""In-kernel timeline of the decode GEMM (split-K slices) inside a PDL chain: %globaltimer stamps per CTA.
    python tools/gemm_timeline.py [B]"""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from valle2_b200 import _lib, ops  # noqa: E402

dev, bf = 'cuda', torch.bfloat16
B, d, F = (int(sys.argv[1]) if len(sys.argv) > 1 else 32), 1024, 4096
lib = _lib.load()
sm = ops.device_info()['sm_count']
names = ['prologue', 'w requested', 'dep resolved', '1st kblock', 'mma issued', 'acc done', 'acc in regs', 'epi done']
for k, (n, kk) in {'qkv': (3 * d, d), 'o': (d, d), 'f1': (F, d), 'f2': (d, F)}.items():
    ws = [(torch.randn(n, kk, device=dev) / 32).to(bf) for _ in range(6)]
    a = torch.randn(B, kk, device=dev).to(bf)
    ns = ops.linear_decode_splits(n, kk, 32, B)
    part = torch.zeros(ns, B, n, device=dev)
    dbgs = [torch.zeros(sm, 16, device=dev, dtype=torch.int64) for _ in range(6)]

    def chain():
        for w, dbg in zip(ws, dbgs):
            _lib.check(lib.vb_linear_decode_set_debug(dbg.data_ptr()), 'dbg')
            ops.linear_decode(a, w, part, B * n, 32)
        _lib.check(lib.vb_linear_decode_set_debug(None), 'dbg')

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
    # kernel j = 3 (middle of the chain); reference time = when kernel 2's last CTA finished its epilogue
    prev_end = T[2][:grid, 7].max()
    t = T[3][:grid]
    for i, nm in enumerate(names):
        col = t[:, i] - prev_end
        print(f'   {nm:13s} min {col.min() / 1e3:7.2f}  med {np.median(col) / 1e3:7.2f}  max {col.max() / 1e3:7.2f}  us after the previous kernel\'s last epilogue store')
    cyc = T[3][:grid, 8:16]
    for i in range(2, 7):
        dcy = cyc[:, i + 1] - cyc[:, i]
        print(f'   cycles {names[i]:13s} -> {names[i + 1]:13s} min {dcy.min():7.0f}  med {np.median(dcy):7.0f}  max {dcy.max():7.0f}')
    print(f'   kernel-to-kernel period {(T[4][:grid, 7].max() - T[3][:grid, 7].max()) / 1e3:.2f} us')
