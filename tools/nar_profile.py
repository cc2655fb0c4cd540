"""Kernel breakdown of NAR config 3 (B=64, S=900, 7 stages) in situ.   python tools/nar_profile.py
NAR_TT=750 gives the bench job's shape (S = 150 + 225 + 750 = 1125)."""
import os
import sys
import tempfile

import torch

sys.path.insert(0, '.')
import valle2_b200  # noqa: E402
from bench import large_cfg  # noqa: E402
from valle2_b200.models import ValleNAR  # noqa: E402

valle2_b200.set_precision('bf16')
dev = torch.device('cuda')
torch.manual_seed(1)
nar = ValleNAR(large_cfg('AdaptiveLayerNorm', tempfile.mkdtemp())).eval().to(dev)
g = torch.Generator().manual_seed(7)
Bn, Tc, Tt = 64, 225, int(os.environ.get('NAR_TT', '525'))
pt = torch.randint(0, 256, (Bn, 50), generator=g).to(dev)
tt = torch.randint(0, 256, (Bn, 100), generator=g).to(dev)
pc = torch.randint(0, 1024, (Bn, Tc, 8), generator=g).to(dev)
fl = torch.randint(0, 1024, (Bn, Tt), generator=g).to(dev)
nar.generate_batch(pt[:2], pc[:2], tt[:2], fl[:2])
nar.generate_batch(pt, pc, tt, fl)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
nar.generate_batch(pt, pc, tt, fl)
e1.record()
torch.cuda.synchronize()
print('7 stages: %.1f ms' % e0.elapsed_time(e1))
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    nar.generate_batch(pt, pc, tt, fl)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=14, max_name_column_width=60))
