"""Wall time of the full-size training step (forward + backward), 8 repetitions after 2 warm-ups.  python tools/train_time.py"""
import sys
import tempfile
import time

import torch

sys.path.insert(0, '.')
import valle2_b200  # noqa: E402
from bench import large_cfg  # noqa: E402
from valle2_b200.models import ValleAR  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
valle2_b200.set_precision('bf16')
torch.manual_seed(0)
model = ValleAR(large_cfg('LayerNorm', tempfile.mkdtemp())).train().cuda()
g = torch.Generator().manual_seed(1)
Tx, Ty = 225, 1126
batch = {'tokens': torch.randint(0, 256, (B, Tx), generator=g), 'tokens_lens': torch.full((B,), Tx),
         'codes': torch.randint(0, 1024, (B, Ty), generator=g), 'codes_lens': torch.full((B,), Ty),
         'target': torch.randint(0, 1025, (B, Ty), generator=g)}
ts = []
for it in range(10):
    for p_ in model.parameters():
        p_.grad = None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = model.training_step(batch)
    h1 = time.perf_counter()
    loss.backward()
    h2 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    ts.append((e0.elapsed_time(e1), 1e3 * (h1 - t0), 1e3 * (h2 - h1)))
for i, (gpu, hf, hb) in enumerate(ts):
    print(f'iter {i}: device {gpu:7.1f} ms | host issue forward {hf:6.1f} ms, backward {hb:6.1f} ms')
