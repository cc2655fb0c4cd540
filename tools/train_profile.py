"""Where the time of one full-size training step goes (forward / backward, per kernel family).
    python tools/train_profile.py [batch]"""
import sys
import tempfile

import torch

sys.path.insert(0, '.')
import valle2_b200  # noqa: E402
from bench import large_cfg  # noqa: E402
from valle2_b200.models import ValleAR  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
valle2_b200.set_precision('bf16')
tmp = tempfile.mkdtemp()
torch.manual_seed(0)
model = ValleAR(large_cfg('LayerNorm', tmp)).train().cuda()
g = torch.Generator().manual_seed(1)
Tx, Ty = 225, 1126
batch = {'tokens': torch.randint(0, 256, (B, Tx), generator=g), 'tokens_lens': torch.full((B,), Tx),
         'codes': torch.randint(0, 1024, (B, Ty), generator=g), 'codes_lens': torch.full((B,), Ty),
         'target': torch.randint(0, 1025, (B, Ty), generator=g)}
for _ in range(2):
    loss = model.training_step(batch)
    loss.backward()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    loss = model.training_step(batch)
    loss.backward()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=70))
print('loss', float(loss), 'peak mem GB', torch.cuda.max_memory_allocated() / 1e9)
