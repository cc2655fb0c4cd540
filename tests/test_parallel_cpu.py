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


def test_sharded_generation_matches_single_process(tmp_path):
    from valle2_b200 import parallel
    n_items = 7
    assert parallel.shard_indices(7, 0, 2) == [0, 2, 4, 6] and parallel.shard_indices(7, 1, 2) == [1, 3, 5]
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
