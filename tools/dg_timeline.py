#!/usr/bin/env python
"""In-kernel timeline of the fused decode GEMMs (csrc/gemm_decode_tc.cu) inside a captured PDL chain.

    python tools/dg_timeline.py [B]        -> table to stdout

One decoder layer's four GEMMs (out-proj, FFN1, FFN2, QKV; full-size shapes) are captured `reps` times in one CUDA graph;
every launch stamps its own debug buffer with %globaltimer (vb_decode_gemm_set_debug): prologue done, weights requested,
dependency resolved, first k-block landed, MMAs issued, accumulator complete, [8] partial stored, [9] arrived on the
counter, [6] all splits arrived, [7] outputs stored.  Printed per GEMM kind, median over launches of:
  the LAST CTA's time of each event relative to the previous kernel's last 'outputs stored' (i.e. the critical path).
Also the chain's wall time per layer from CUDA events, with and without the stamps.
"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from valle2_b200 import _lib, ops  # noqa: E402

EVENTS = [(0, 'prologue done'), (1, 'weights requested'), (2, 'dependency resolved'), (3, 'first k-block landed'),
          (4, 'MMAs issued'), (5, 'accumulator complete'), (8, 'partial stored'), (9, 'arrived'), (6, 'all splits arrived'),
          (10, 'first item reduced'), (11, 'its operands ready'), (7, 'outputs stored')]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    xflags = ops.FLAG_DG_GLOBAL if (len(sys.argv) > 2 and sys.argv[2] == 'global') else 0
    print('exchange:', 'L2 buffer + counters' if xflags else 'thread-block cluster / DSMEM where it applies')
    reps = 6
    d, F = 1024, 4096
    torch.manual_seed(0)
    mk = lambda n, k: (torch.randn(n, k, device='cuda') / k ** 0.5).bfloat16()
    layers = [dict(wo=mk(d, d), w1=mk(F, d), w2=mk(d, F), wq=mk(3 * d, d)) for _ in range(reps)]
    vec = lambda n: torch.randn(n, device='cuda') * 0.1
    bo, b1, b2, bq, c1, cq = vec(d), vec(F), vec(d), vec(3 * d), vec(F), vec(3 * d)
    shapes = {'o': (d, d), 'f1': (F, d), 'f2': (d, F), 'qkv': (3 * d, d)}
    plans = {k: ops.decode_gemm_plan(B, n, kk) for k, (n, kk) in shapes.items()}
    print('plans:', {k: (v['tiles'], v['n_split']) for k, v in plans.items()})
    ws = torch.zeros(max(p['ws_bytes'] for p in plans.values()) // 4, device='cuda')
    ctr = torch.zeros(4, 256, dtype=torch.int32, device='cuda')
    stats = torch.zeros(B * 64 * 2, device='cuda')
    o = torch.randn(B, d, device='cuda').bfloat16()
    x, xb = torch.randn(B, d, device='cuda'), torch.zeros(B, d, device='cuda', dtype=torch.bfloat16)
    f = torch.zeros(B, F, device='cuda', dtype=torch.bfloat16)
    qkv = torch.zeros(B, 3 * d, device='cuda')
    L = _lib.load()
    kinds = ['o', 'f1', 'f2', 'qkv']
    dbg = {(r, k): torch.zeros(148 * 16, dtype=torch.int64, device='cuda') for r in range(reps) for k in kinds}

    def chain(stamp):
        for r, W in enumerate(layers):
            def set_dbg(k):
                L.vb_decode_gemm_set_debug(dbg[(r, k)].data_ptr() if stamp else None)
            set_dbg('o')
            ops.decode_gemm(o, W['wo'], ops.DG_RESIDUAL, ws=ws, counters=ctr[0], bias=bo, xres=x, y16=xb, stats_out=stats, flags=xflags)
            set_dbg('f1')
            ops.decode_gemm(xb, W['w1'], ops.DG_LN_GELU, ws=ws, counters=ctr[1], bias=b1, colsum=c1, stats_in=stats,
                            n_chunks_in=plans['o']['tiles'], y16=f, flags=xflags)
            set_dbg('f2')
            ops.decode_gemm(f, W['w2'], ops.DG_RESIDUAL, ws=ws, counters=ctr[2], bias=b2, xres=x, y16=xb, stats_out=stats, flags=xflags)
            set_dbg('qkv')
            ops.decode_gemm(xb, W['wq'], ops.DG_LN, ws=ws, counters=ctr[3], bias=bq, colsum=cq, stats_in=stats,
                            n_chunks_in=plans['f2']['tiles'], y32=qkv, flags=xflags)
        L.vb_decode_gemm_set_debug(None)

    def timed(stamp):
        chain(stamp)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            chain(stamp)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 10 / reps * 1e3

    print(f'B={B}: chain of 4 fused GEMMs per layer: {timed(False):.2f} us/layer without stamps, {timed(True):.2f} us with stamps')
    # analyse the last replay's stamps
    prev_end = None
    rows = {k: {e: [] for e, _ in EVENTS} for k in kinds}
    rows_first = {k: {e: [] for e, _ in EVENTS} for k in kinds}
    for r in range(reps):
        for k in kinds:
            n_cta = plans[k]['tiles'] * plans[k]['n_split']
            t = dbg[(r, k)].view(148, 16)[:n_cta].cpu()
            end = int(t[:, 7].max())
            if prev_end is not None and r >= 1:
                for e, _ in EVENTS:
                    col = t[:, e]
                    col = col[col > 0]
                    if len(col):
                        rows[k][e].append((int(col.max()) - prev_end) / 1e3)
                        rows_first[k][e].append((int(col.min()) - prev_end) / 1e3)
            prev_end = end
    print(f'{"event (us after the previous kernel stored its outputs)":58s}' + ''.join(f'{k:>16s}' for k in kinds))
    for e, name in EVENTS:
        line = f'{name:58s}'
        for k in kinds:
            if rows[k][e]:
                line += f'{statistics.median(rows_first[k][e]):7.2f}/{statistics.median(rows[k][e]):6.2f}  '
            else:
                line += ' ' * 16
        print(line)
    print('(first CTA / last CTA to reach the event; %globaltimer resolution is coarse: differences below ~0.3 us are noise)')


if __name__ == '__main__':
    main()
