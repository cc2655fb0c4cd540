#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (racecheck / synccheck / memcheck, one tool per gpurun call):
tiny model (2 layers, d = 256), bf16: AR decode on the lean path (B = 2), on the fused tcgen05 decode GEMMs with the cluster /
DSMEM exchange (B = 9) and on the split-K path (B = 40), a few steps each, NAR stages, one training step with dropout.
    compute-sanitizer --tool racecheck python tools/sanitize_target.py"""
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import valle2_b200  # noqa: E402
from valle2_b200.config import ConfigValle  # noqa: E402
from valle2_b200.models import ValleAR, ValleNAR  # noqa: E402


def main():
    tmp = tempfile.mkdtemp(prefix='valle_san_')
    kw = dict(num_layers=2, d_model=256, n_heads=4, dim_feedforward=1024, ckpt_path=os.path.join(tmp, 'c'), log_path=os.path.join(tmp, 'l'))
    valle2_b200.set_precision('bf16')
    torch.manual_seed(0)
    ar = ValleAR(ConfigValle(norm='LayerNorm', dropout=0.1, max_audio_len=6, top_k=1, **kw)).cuda().eval()
    nar = ValleNAR(ConfigValle(norm='AdaptiveLayerNorm', dropout=0.0, **kw)).cuda().eval()
    g = torch.Generator().manual_seed(1)
    for B in (2, 9, 40):
        tok = torch.randint(0, 256, (B, 12), generator=g).cuda()
        cod = torch.cat([torch.full((B, 1), ar.bos_token), torch.randint(0, 1024, (B, 70), generator=g)], 1).cuda()
        out, n = ar.generate_batch(tok, cod, max_new=5, ignore_eos=True, use_graph=False)
        print('AR decode', B, tuple(out.shape), n, flush=True)
    pt, tt = torch.randint(0, 256, (2, 5), generator=g).cuda(), torch.randint(0, 256, (2, 7), generator=g).cuda()
    pc, fl = torch.randint(0, 1024, (2, 9, 8), generator=g).cuda(), torch.randint(0, 1024, (2, 11), generator=g).cuda()
    print('NAR', tuple(nar.generate_batch(pt, pc, tt, fl).shape), flush=True)
    ar.train()
    batch = {'tokens': torch.randint(0, 256, (2, 6), generator=g), 'tokens_lens': torch.tensor([6, 4]),
             'codes': torch.randint(0, 1024, (2, 9), generator=g), 'codes_lens': torch.tensor([9, 7]),
             'target': torch.randint(0, 1025, (2, 9), generator=g)}
    loss = ar.training_step(batch)
    loss.backward()
    torch.cuda.synchronize()
    print('train loss', float(loss), flush=True)


if __name__ == '__main__':
    main()
