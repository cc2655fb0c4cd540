"""Timeline of one captured decode step at small batch: every rows-GEMM launch writes its per-CTA stamps into its own
buffer (vb_linear_decode_rows_set_debug), so the gaps between consecutive GEMMs are the attention kernels in situ.
    python tools/step_timeline.py [B]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import valle2_b200  # noqa: E402
from valle2_b200 import _lib, ops  # noqa: E402
from valle2_b200.models import ValleAR  # noqa: E402
from bench import large_cfg, TX, P0  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
valle2_b200.set_precision('bf16')
dev = torch.device('cuda')
torch.manual_seed(0)
model = ValleAR(large_cfg('LayerNorm', '/tmp/vb_timeline')).eval().to(dev)
eng = model._engine()
lib = _lib.load()
samp = {'temperature': 1.0, 'top_k': 1, 'top_p': 1.0, 'seed': 0}
g_ = torch.Generator().manual_seed(1)
tokens = torch.randint(0, 256, (B, TX), generator=g_).to(dev)
codes = torch.cat([torch.full((B, 1), 1025), torch.randint(0, 1024, (B, P0 + 370), generator=g_)], 1).to(dev)
eng.prefill(tokens, codes, max_new=64)
eng.first_token(samp, None, -1)
eng.decode_step(samp, None, -1)
torch.cuda.synchronize()

# stamp buffers are allocated BEFORE the capture: an allocation inside it adds a fill kernel between two GEMMs, which is a
# normal (non-programmatic) node and serialises the chain -- the timeline would then show 2.3 us hops that do not exist
pool = [torch.zeros(512, 16, device=dev, dtype=torch.int64) for _ in range(80)]
bufs, names = [], []
real_rows, real_ln = ops.linear_decode_rows, ops.linear_decode_rows_ln


def wrap(fn, tag):
    def inner(x, w, y, **kw):
        buf = pool[len(bufs)]
        bufs.append(buf)
        names.append(f'{tag} N={w.shape[0]} K={w.shape[1]}')
        _lib.check(lib.vb_linear_decode_rows_set_debug(buf.data_ptr()), 'dbg')
        r = fn(x, w, y, **kw)
        _lib.check(lib.vb_linear_decode_rows_set_debug(None), 'dbg')
        return r
    return inner


apool = [torch.zeros(4096, 8, device=dev, dtype=torch.int64) for _ in range(16)]
abufs = []
real_attn = ops.attn_decode_paged


def attn_wrap(*a, **k):
    buf = apool[len(abufs)]
    abufs.append(buf)
    bufs.append(buf)
    names.append('attention')
    _lib.check(lib.vb_attn_decode_set_debug(buf.data_ptr()), 'dbg')
    r = real_attn(*a, **k)
    _lib.check(lib.vb_attn_decode_set_debug(None), 'dbg')
    return r


real_ld = ops.linear_decode
gpool = [torch.zeros(512, 16, device=dev, dtype=torch.int64) for _ in range(64)]
gused = []


def ld_wrap(x, w, part, part_stride, max_split, flags=0):
    buf = gpool[len(gused)]
    gused.append(buf)
    bufs.append(buf)
    names.append(f'tc-splitK N={w.shape[0]} K={w.shape[1]}')
    _lib.check(lib.vb_linear_decode_set_debug(buf.data_ptr()), 'dbg')
    r = real_ld(x, w, part, part_stride, max_split, flags)
    _lib.check(lib.vb_linear_decode_set_debug(None), 'dbg')
    return r


real_ln = ops.residual_layernorm
lpool = [torch.zeros(64, 4, device=dev, dtype=torch.int64) for _ in range(64)]
lused = []


def ln_wrap(x, *a, **k):
    if x.shape[0] > 64:
        return real_ln(x, *a, **k)
    buf = lpool[len(lused)]
    lused.append(buf)
    bufs.append(buf)
    names.append(f'LN n_part={k.get("n_part", 0)}')
    _lib.check(lib.vb_residual_layernorm_set_debug(buf.data_ptr()), 'dbg')
    r = real_ln(x, *a, **k)
    _lib.check(lib.vb_residual_layernorm_set_debug(None), 'dbg')
    return r


ops.residual_layernorm = ln_wrap
ops.linear_decode = ld_wrap
ops.attn_decode_paged = attn_wrap
ops.linear_decode_rows, ops.linear_decode_rows_ln = wrap(real_rows, 'rows'), wrap(real_ln, 'rows+LN')
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    eng.decode_step(samp, None, -1)
ops.linear_decode_rows, ops.linear_decode_rows_ln = real_rows, real_ln
ops.attn_decode_paged = real_attn
ops.linear_decode = real_ld
ops.residual_layernorm = real_ln
for _ in range(3):
    graph.replay()
torch.cuda.synchronize()
T = [b.cpu().numpy().astype(np.float64) for b in bufs]
t0 = None
prev_end = None
print(f'B={B}: kernel | start (first CTA) | dep resolved (median) | last store | gap since previous GEMM\'s last store  [us]')
ev = ['start', 'copies issued', 'dep resolved', 'q ready', '1st page', 'pages done', 'partial out', 'o written']
for nm, t in zip(names, T):
    if nm == 'attention':
        t = t[t[:, 0] > 0]
        if t0 is None:
            t0 = t[:, 0].min()
        parts = []
        for i, e in enumerate(ev):
            col = t[:, i][t[:, i] > 0]
            if len(col):
                parts.append(f'{e} {(np.median(col) - t0) / 1e3:.2f}' + (f'..{(col.max() - t0) / 1e3:.2f}' if i in (0, 5, 6, 7) else ''))
        print(f'  attention ({len(t)} CTAs)       ' + ' | '.join(parts))
        if len(abufs) and nm == 'attention' and not globals().get('_split_done'):
            globals()['_split_done'] = True
            ns = eng._state['subs'][0]['n_tsplit']
            tt = t.reshape(-1, ns, 8)
            for sp in range(ns):
                c = tt[:, sp]
                print(f'      split {sp}: copies issued {(np.median(c[:, 1]) - t0) / 1e3:6.2f}  1st page {(np.median(c[:, 4][c[:, 4] > 0]) - t0) / 1e3 if (c[:, 4] > 0).any() else float("nan"):6.2f}'
                      f'  pages done {(np.median(c[:, 5]) - t0) / 1e3:6.2f}  partial out {(np.median(c[:, 6]) - t0) / 1e3:6.2f}')
        prev_end = t[:, 7].max()
        continue
    if nm.startswith('LN'):
        t = t[t[:, 0] > 0]
        if t0 is None:
            t0 = t[:, 0].min()
        gap = (np.median(t[:, 1]) - prev_end) if prev_end is not None else 0.0
        print(f'  {nm:24s} start {(t[:, 0].min() - t0) / 1e3:8.2f}  dep {(np.median(t[:, 1]) - t0) / 1e3:8.2f}  loaded {(np.median(t[:, 2]) - t0) / 1e3:8.2f}..{(t[:, 2].max() - t0) / 1e3:.2f}  end {(t[:, 3].max() - t0) / 1e3:8.2f}   dep - prev end {gap / 1e3:6.2f}')
        prev_end = t[:, 3].max()
        continue
    if nm.startswith('tc-splitK'):      # stamps: 0 prologue, 1 weights requested, 2 dep resolved, 3 first k-block, 4 MMAs issued, 5 acc done, 6 acc in regs, 7 stores
        t = t[t[:, 0] > 0]
        if t0 is None:
            t0 = t[:, 0].min()
        dep, end = np.median(t[:, 2]) - t0, t[:, 7].max() - t0
        gap = (np.median(t[:, 2]) - prev_end) if prev_end is not None else 0.0
        print(f'  {nm:24s} start {(t[:, 0].min() - t0) / 1e3:8.2f}  dep {dep / 1e3:8.2f}  1st kblock {(np.median(t[:, 3]) - t0) / 1e3:8.2f}  acc done {(np.median(t[:, 5]) - t0) / 1e3:8.2f}  end {end / 1e3:8.2f}   dep - prev end {gap / 1e3:6.2f}   run {(end - dep) / 1e3:5.2f}')
        prev_end = t[:, 7].max()
        continue
    live = t[:, 0] > 0
    if not live.any():
        continue
    t = t[live]
    if t0 is None:
        t0 = t[:, 0].min()
    start, dep, end = t[:, 0].min() - t0, np.median(t[:, 2]) - t0, t[:, 6].max() - t0
    gap = (np.median(t[:, 2]) - prev_end) if prev_end is not None else 0.0
    print(f'  {nm:24s} start {start / 1e3:8.2f}  dep {dep / 1e3:8.2f}  end {end / 1e3:8.2f}   dep - prev end {gap / 1e3:6.2f}   run {(end - dep) / 1e3:5.2f}')
    prev_end = t[:, 6].max()
