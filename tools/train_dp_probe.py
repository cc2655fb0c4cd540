#!/usr/bin/env python
"""Where the data-parallel training step spends its extra time (run under torchrun, N >= 2 GPUs):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_dp_probe.py
Prints ms per full-size AR training step (16 clips of 15 s per GPU, forward + backward): local, with the bucket copies only,
with the overlapped all-reduce, with a blocking all-reduce after the backward pass."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    rank, lr = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(lr)
    dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
    import valle2_b200
    from valle2_b200 import parallel
    from valle2_b200.models import ValleAR
    valle2_b200.set_precision('bf16')
    dev = torch.device('cuda', lr)
    torch.manual_seed(2)
    model = ValleAR(bench.large_cfg('LayerNorm', f'/tmp/dp_probe_{os.getpid()}')).train().to(dev)
    g = torch.Generator().manual_seed(11 + rank)
    Bt, Txt, Tyt = 16, 225, 1126
    batch = {'tokens': torch.randint(0, 256, (Bt, Txt), generator=g), 'tokens_lens': torch.full((Bt,), Txt),
             'codes': torch.randint(0, 1024, (Bt, Tyt), generator=g), 'codes_lens': torch.full((Bt,), Tyt),
             'target': torch.randint(0, 1025, (Bt, Tyt), generator=g)}
    red = parallel.GradReducer(model)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {}
    for label in ('local', 'copies_only', 'overlapped', 'after_backward'):
        red.no_comm = label == 'copies_only'
        reducer = None if label in ('local', 'after_backward') else red
        for it in range(6):
            if it == 3:
                dist.barrier(); torch.cuda.synchronize(); e0.record()
            for p in model.parameters():
                p.grad = None
            with parallel.reducing(reducer):
                loss = model.training_step(batch)
            loss.backward()
            if label == 'after_backward':
                parallel.allreduce_gradients(model)
        e1.record(); dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[label] = round(float(t.item()), 2)
    if rank == 0:
        print(json.dumps({'world': dist.get_world_size(), 'ms_per_step': out}))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
