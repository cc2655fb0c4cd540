#!/usr/bin/env python
"""How much faster is the paged decode attention when (part of) its KV pages sit in L2, and does an L2 prefetch issued
ahead of it (cp.async.bulk.prefetch.L2, vb_kv_prefetch_l2) put them there?

    python tools/l2_probe.py            -> JSON lines

(1) `resident`: the kernel over ONE layer pool launched 12 times back to back (pages stay in L2 when the pool fits)
    against 12 distinct pools (HBM), for B = 8 / 16 / 32 at ctx 750 (25 / 49 / 98 MB of KV per launch).
(2) `prefetch`: per layer  [prefetch pct% of the pool] -> [spin ~20 us: the GEMM chain's stand-in, HBM idle] -> [attention]
    against the same sequence without the prefetch; the spin alone is timed too so that it can be subtracted.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from valle2_b200 import ops  # noqa: E402

H, Dh, d, L = 16, 64, 1024, 12


def time_graph(fn, reps=20):
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
    return e0.elapsed_time(e1) / reps * 1e3      # us per fn()


def setup(B, ctx):
    max_pages = (ctx + 64) // 64 + 1
    pools = torch.randn(L, B * max_pages, 2, H, 64, Dh, device='cuda').bfloat16()
    bt = torch.arange(B * max_pages, device='cuda', dtype=torch.int32).view(B, max_pages)
    seq = torch.full((B,), ctx, device='cuda', dtype=torch.int32)
    part = torch.randn(6, B, 3 * d, device='cuda')
    o = torch.empty(B, d, device='cuda', dtype=torch.bfloat16)
    nts = max(1, min(8, -(-int(3.46 * 148) // (B * H))))
    ws = torch.zeros(ops.attn_decode_ws_bytes(B, H, nts) // 4 + 64, device='cuda', dtype=torch.int32)
    return pools, bt, seq, part, o, nts, ws


def main():
    ctx = 750
    spin_cycles = int(os.environ.get('SPIN_CYCLES', '38000'))
    for B in (8, 16, 32):
        pools, bt, seq, part, o, nts, ws = setup(B, ctx)
        mb = B * (ctx + 1) * 2 * d * 2 / 1e6

        def attn(li):
            ops.attn_decode_paged(part, 6, B * 3 * d, pools[li], bt, seq, o, B, H, Dh, nts, ws)

        t_hbm = time_graph(lambda: [attn(li) for li in range(L)]) / L
        t_l2 = time_graph(lambda: [attn(0) for _ in range(L)]) / L
        print(json.dumps({'probe': 'resident', 'B': B, 'kv_mb': round(mb, 1), 'n_tsplit': nts, 'us_12_pools': round(t_hbm, 2),
                          'us_same_pool': round(t_l2, 2), 'gbs_hbm': round(mb / t_hbm * 1e3, 0),
                          'gbs_same_pool': round(mb / t_l2 * 1e3, 0)}), flush=True)
        if B != 32:
            del pools
            continue
        t_spin = time_graph(lambda: [torch.cuda._sleep(spin_cycles) for _ in range(L)]) / L
        t_base = time_graph(lambda: [(torch.cuda._sleep(spin_cycles), attn(li)) for li in range(L)]) / L
        print(json.dumps({'probe': 'prefetch', 'pct': 0, 'us_spin': round(t_spin, 2), 'us_spin_attn': round(t_base, 2),
                          'us_attn': round(t_base - t_spin, 2)}), flush=True)
        for pct in (10, 25, 40, 50, 60, 75, 100):
            def seq_pf(li):
                ops.kv_prefetch_l2(pools[li], bt, seq, B, H, Dh, 0, pct)
                torch.cuda._sleep(spin_cycles)
                attn(li)
            t = time_graph(lambda: [seq_pf(li) for li in range(L)]) / L
            t_pf = time_graph(lambda: [ops.kv_prefetch_l2(pools[li], bt, seq, B, H, Dh, 0, pct) for li in range(L)]) / L
            print(json.dumps({'probe': 'prefetch', 'pct': pct, 'mb_prefetched': round(mb * pct / 100, 1), 'us_prefetch_kernel': round(t_pf, 2),
                              'us_total': round(t, 2), 'us_attn_after_prefetch': round(t - t_spin - t_pf, 2)}), flush=True)
        # the tail of the pages instead of the head (the kernel's ring fills from page 0 before q arrives anyway)
        for lo, hi in ((25, 75), (50, 100)):
            def seq_pf2(li):
                ops.kv_prefetch_l2(pools[li], bt, seq, B, H, Dh, lo, hi)
                torch.cuda._sleep(spin_cycles)
                attn(li)
            t = time_graph(lambda: [seq_pf2(li) for li in range(L)]) / L
            print(json.dumps({'probe': 'prefetch', 'range': [lo, hi], 'us_total': round(t, 2),
                              'us_attn_after_prefetch_incl_kernel': round(t - t_spin, 2)}), flush=True)


if __name__ == '__main__':
    main()
