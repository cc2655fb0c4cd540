"""Where a decode step's time goes, in situ: the captured step replayed (a) as is, (b) with the attention launches removed,
(c) with everything but attention removed (monkey-patched ops; outputs are garbage, only the timing is used).
    python tools/step_breakdown.py [B ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import valle2_b200  # noqa: E402
from valle2_b200 import ops  # noqa: E402
from valle2_b200.models import ValleAR  # noqa: E402
from bench import large_cfg, TX, P0  # noqa: E402

valle2_b200.set_precision('bf16')
dev = torch.device('cuda')
torch.manual_seed(0)
model = ValleAR(large_cfg('LayerNorm', '/tmp/vb_breakdown')).eval().to(dev)
eng = model._engine()
samp = {'temperature': 1.0, 'top_k': 1, 'top_p': 1.0, 'seed': 0}
real = {k: getattr(ops, k) for k in ('attn_decode_paged', 'linear_decode', 'linear_decode_rows', 'linear_decode_rows_ln',
                                     'residual_layernorm', 'reduce_bias_act')}


def timed_step(K=200):
    eng.decode_step(samp, None, -1)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        eng.decode_step(samp, None, -1)
    for _ in range(5):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(K):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K * 1e3


for B in [int(a) for a in sys.argv[1:]] or [1, 32]:
    g_ = torch.Generator().manual_seed(1)
    tokens = torch.randint(0, 256, (B, TX), generator=g_).to(dev)
    codes = torch.cat([torch.full((B, 1), 1025), torch.randint(0, 1024, (B, P0 + 370), generator=g_)], 1).to(dev)
    eng.prefill(tokens, codes, max_new=700)
    eng.first_token(samp, None, -1)
    full = timed_step()
    ops.attn_decode_paged = lambda *a, **k: None
    no_attn = timed_step()
    ops.attn_decode_paged = real['attn_decode_paged']
    for k in ('linear_decode', 'linear_decode_rows', 'linear_decode_rows_ln', 'residual_layernorm', 'reduce_bias_act'):
        setattr(ops, k, lambda *a, **kw: 1)
    only_attn = timed_step()
    for k, v in real.items():
        setattr(ops, k, v)
    print(f'B={B:3d}  step {full:7.1f} us | without attention {no_attn:7.1f} | attention only (+embed, sample) {only_attn:7.1f} '
          f'| sum of parts {no_attn + only_attn:7.1f}  launches {eng.launches_per_step()}')
    if os.environ.get('BREAKDOWN_FAMILIES'):       # leave out one kernel family of the chain at a time
        for fam in (('linear_decode',), ('residual_layernorm',), ('reduce_bias_act',), ('residual_layernorm', 'reduce_bias_act')):
            for k in fam:
                setattr(ops, k, lambda *a, **kw: 1)
            t = timed_step()
            for k, v in real.items():
                setattr(ops, k, v)
            print(f'        without {"+".join(fam):40s} {t:7.1f} us  (-{full - t:6.1f})')
