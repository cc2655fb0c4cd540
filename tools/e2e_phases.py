"""Wall-clock phases of ARDecoder.generate at the bench shape (B=32, 758 steps), five repetitions: which phase makes the
end-to-end time vary?   python tools/e2e_phases.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import valle2_b200  # noqa: E402
from valle2_b200.models import ValleAR  # noqa: E402
from bench import large_cfg, TX, P0  # noqa: E402

valle2_b200.set_precision('bf16')
dev = torch.device('cuda')
torch.manual_seed(0)
model = ValleAR(large_cfg('LayerNorm', '/tmp/vb_e2e')).eval().to(dev)
eng = model._engine()
B, steps = 32, 758
g = torch.Generator().manual_seed(100)
tokens_h = torch.randint(0, 256, (B, TX), generator=g).pin_memory()
codes_h = torch.cat([torch.full((B, 1), 1025), torch.randint(0, 1024, (B, P0 - 1), generator=g)], 1).pin_memory()
samp = {'temperature': 1.0, 'top_k': 1, 'top_p': 1.0, 'seed': 0}


def now():
    torch.cuda.synchronize()
    return time.perf_counter()


for rep in range(6):
    t = [now()]
    tok, cod = tokens_h.to(dev, non_blocking=True), codes_h.to(dev, non_blocking=True)
    st = eng.prefill(tok, cod, max_new=steps); t.append(now())
    eng.first_token(samp, None, -1)
    eng.decode_step(samp, None, -1); t.append(now())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        eng.decode_step(samp, None, -1)
    t.append(now())
    eng._graph = graph
    h0 = time.perf_counter()
    for _ in range(steps - 2):
        graph.replay()
    h1 = time.perf_counter()
    t.append(now())
    out = st['codes_out'].to('cpu'); t.append(now())
    d = [(b - a) * 1e3 for a, b in zip(t, t[1:])]
    print(f'rep {rep}: prefill {d[0]:7.1f} ms | first + warm step {d[1]:6.1f} | capture {d[2]:6.1f} | {steps - 2} replays {d[3]:7.1f} '
          f'(host issue {1e3 * (h1 - h0):7.1f}) | D2H {d[4]:5.1f} | total {sum(d):7.1f}')
