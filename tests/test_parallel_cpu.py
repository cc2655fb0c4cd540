"""World-size-2 gloo test (CPU) of the N>1 path: utterance sharding + the single end-of-job all-gather must
reproduce the single-process result bit for bit, including ragged lengths and an odd item count."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _fake_decode(idx):
    """Deterministic stand-in for generate_batch: item i yields (i % 4) + 2 codes  i*100 + t."""
    lens = torch.tensor([(i % 4) + 2 for i in idx], dtype=torch.int32)
    T = int(lens.max()) if len(idx) else 0
    rows = torch.zeros(len(idx), T, dtype=torch.int32)
    for r, i in enumerate(idx):
        rows[r, : int(lens[r])] = torch.arange(int(lens[r]), dtype=torch.int32) + i * 100
    return rows, lens


def _worker(rank, world_size, port, n_items, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    from valle2_b200 import parallel
    rows, lens = parallel.generate_sharded(_fake_decode, n_items, pad_value=-1)
    torch.save((rows, lens), os.path.join(out_dir, f'r{rank}.pt'))
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize('n_items', [7, 1, 0])      # odd count; fewer items than ranks (rank 1's shard is empty); nothing at all
def test_sharded_generation_matches_single_process(tmp_path, n_items):
    from valle2_b200 import parallel
    assert parallel.shard_indices(7, 0, 2) == [0, 2, 4, 6] and parallel.shard_indices(7, 1, 2) == [1, 3, 5]
    assert parallel.shard_indices(1, 1, 2) == [] and parallel.shard_indices(0, 0, 2) == []
    ref_rows, ref_lens = _fake_decode(list(range(n_items)))
    single_rows, single_lens = parallel.generate_sharded(_fake_decode, n_items, pad_value=-1)
    assert torch.equal(single_lens, ref_lens)
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, n_items, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        rows, lens = torch.load(os.path.join(tmp_path, f'r{rank}.pt'))
        assert torch.equal(lens, ref_lens)
        assert rows.shape[0] == n_items
        for i in range(n_items):
            n = int(ref_lens[i])
            assert torch.equal(rows[i, :n], ref_rows[i, :n])
            assert (rows[i, n:] == -1).all()


def _grad_worker(rank, world_size, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    from valle2_b200 import parallel
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3), torch.nn.Linear(3, 2))
    x = torch.arange(20, dtype=torch.float32).view(4, 5) / 10
    part = parallel.shard_batch({'x': x, 'meta': 'kept'}, rank, world_size)
    assert part['meta'] == 'kept' and part['x'].shape[0] == 2
    # rank-local loss: the MEAN over the local rows; the last layer gets a gradient on rank 0 only
    h = model[1](model[0](part['x']))
    loss = h.pow(2).mean() + (model[2](h).sum() if rank == 0 else 0.0)
    loss.backward()
    n_buckets = parallel.allreduce_gradients(model, bucket_bytes=128)
    assert n_buckets >= 2
    torch.save({n: p.grad.clone() for n, p in model.named_parameters()}, os.path.join(out_dir, f'g{rank}.pt'))
    dist.destroy_process_group()


def test_gradient_allreduce_matches_full_batch(tmp_path):
    """World-size-2 gloo: per-rank gradients of the sharded batch, averaged by allreduce_gradients, equal the gradients of
    the mean loss over the whole batch (the data-parallel training step, BASELINE config 5)."""
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_grad_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3), torch.nn.Linear(3, 2))
    x = torch.arange(20, dtype=torch.float32).view(4, 5) / 10
    h0, h1 = model[1](model[0](x[:2])), model[1](model[0](x[2:]))
    loss = 0.5 * (h0.pow(2).mean() + model[2](h0).sum()) + 0.5 * h1.pow(2).mean()
    loss.backward()
    g0, g1 = torch.load(os.path.join(tmp_path, 'g0.pt')), torch.load(os.path.join(tmp_path, 'g1.pt'))
    for n, p in model.named_parameters():
        assert torch.allclose(g0[n], p.grad, atol=1e-6), n
        assert torch.equal(g0[n], g1[n]), n


class _Toy(torch.nn.Module):
    """Parameter names shaped like the decoders': transformer.layers.{i}.* + a few others (one never used)."""

    def __init__(self):
        super().__init__()
        self.emb = torch.nn.Linear(5, 6)
        self.transformer = torch.nn.Module()
        self.transformer.layers = torch.nn.ModuleList([torch.nn.Linear(6, 6) for _ in range(3)])
        self.proj = torch.nn.Linear(6, 2)
        self.unused = torch.nn.Linear(6, 2)          # no gradient on any rank
        self.rank0_only = torch.nn.Linear(6, 1)      # gradient on rank 0 only

    def forward(self, x, rank):
        h = self.emb(x)
        for layer in self.transformer.layers:
            h = torch.tanh(layer(h))
        out = self.proj(h).pow(2).mean()
        return out + (self.rank0_only(h).sum() if rank == 0 else 0.0)


def _reducer_worker(rank, world_size, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    from valle2_b200 import parallel
    torch.manual_seed(0)
    model = _Toy()
    x = torch.arange(20, dtype=torch.float32).view(4, 5) / 10
    part = parallel.shard_batch({'x': x}, rank, world_size)['x']
    loss = model(part, rank)
    names = [n for n, _ in model.named_parameters()]
    grads = dict(zip(names, torch.autograd.grad(loss, list(model.parameters()), allow_unused=True)))
    red = parallel.GradReducer(model)
    assert len(red.buckets) == 4 and [n for n, _ in red.buckets[1]] == ['transformer.layers.1.weight', 'transformer.layers.1.bias']
    with parallel.reducing(red) as active:
        assert parallel.active_reducer() is active
        red.begin()
        for li in (2, 1, 0):                         # the order the backward pass finishes the layers in
            red.submit(li, {k: v for k, v in grads.items() if k.startswith(f'transformer.layers.{li}.')})
        red.submit(3, grads)
        avg = red.finish()
    assert parallel.active_reducer() is None
    for n, p in model.named_parameters():
        p.grad = avg[n].clone()
    dropped = red.drop_unused(model)
    assert dropped == 2 and model.unused.weight.grad is None and model.rank0_only.weight.grad is not None
    torch.save({n: (None if p.grad is None else p.grad.clone()) for n, p in model.named_parameters()}, os.path.join(out_dir, f'r{rank}.pt'))
    dist.destroy_process_group()


def test_overlapped_grad_reducer_matches_full_batch(tmp_path):
    """World-size-2 gloo: GradReducer (per-layer buckets in one flat buffer, asynchronous all-reduce per bucket, has-gradient
    flags) gives every rank the gradient of the mean loss over the whole batch; a parameter without a gradient on every rank
    keeps grad = None, one with a gradient on one rank gets the average with zeros."""
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_reducer_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    torch.manual_seed(0)
    model = _Toy()
    x = torch.arange(20, dtype=torch.float32).view(4, 5) / 10
    loss = 0.5 * model(x[:2], 0) + 0.5 * model(x[2:], 1)
    loss.backward()
    g0, g1 = torch.load(os.path.join(tmp_path, 'r0.pt')), torch.load(os.path.join(tmp_path, 'r1.pt'))
    for n, p in model.named_parameters():
        if p.grad is None:
            assert g0[n] is None and g1[n] is None, n
            continue
        assert torch.allclose(g0[n], p.grad, atol=1e-6), n
        assert torch.equal(g0[n], g1[n]), n
