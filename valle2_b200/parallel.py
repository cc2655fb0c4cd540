"""Multi-GPU plumbing for the inference path: shard utterances over ranks, decode independently, gather once.

The path has no cross-utterance dependency, so the only collective is one ``all_gather`` of the padded
int32 codes and their lengths at the very end (NCCL over NVLink on GPUs; gloo in the CPU tests).  One process per GPU,
``torch.distributed`` already initialised by the launcher (torchrun)."""
from __future__ import annotations

from typing import Callable

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(n_items: int, rank: int, world_size: int) -> list[int]:
    """Round-robin utterance -> rank assignment (item i lives on rank i % world_size)."""
    return list(range(rank, n_items, world_size))


def gather_ragged(local_rows: torch.Tensor, local_lens: torch.Tensor, n_items: int, pad_value: int = 0):
    """All-gather variable-length integer rows.

    local_rows (n_local, T_local) int32 padded rows of this rank's shard (round-robin order), local_lens (n_local,).
    Returns (rows (n_items, T_max) int32, lens (n_items,) int32) in the ORIGINAL item order on every rank.
    """
    rank, ws = world()
    dev = local_rows.device
    n_local_max = (n_items + ws - 1) // ws
    t_max = torch.tensor([local_rows.shape[1] if local_rows.numel() else 0], device=dev, dtype=torch.int64)
    if ws > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    T = int(t_max.item())
    buf = torch.full((n_local_max, T), pad_value, device=dev, dtype=torch.int32)
    lens = torch.zeros(n_local_max, device=dev, dtype=torch.int32)
    n_local = local_rows.shape[0]
    if n_local:
        buf[:n_local, : local_rows.shape[1]] = local_rows.to(torch.int32)
        lens[:n_local] = local_lens.to(torch.int32)
    def repad(rows_, lens_):      # positions past each row's length always carry pad_value
        if rows_.shape[1]:
            mask = torch.arange(rows_.shape[1], device=dev)[None, :] >= lens_[:, None]
            rows_ = rows_.masked_fill(mask, pad_value)
        return rows_, lens_

    if ws == 1:
        return repad(buf[:n_items], lens[:n_items])
    all_rows = [torch.empty_like(buf) for _ in range(ws)]
    all_lens = [torch.empty_like(lens) for _ in range(ws)]
    dist.all_gather(all_rows, buf)
    dist.all_gather(all_lens, lens)
    rows = torch.full((n_items, T), pad_value, device=dev, dtype=torch.int32)
    out_lens = torch.zeros(n_items, device=dev, dtype=torch.int32)
    for r in range(ws):
        idx = shard_indices(n_items, r, ws)
        if idx:
            rows[idx] = all_rows[r][: len(idx)]
            out_lens[idx] = all_lens[r][: len(idx)]
    return repad(rows, out_lens)


def generate_sharded(decode_fn: Callable[[list[int]], tuple[torch.Tensor, torch.Tensor]], n_items: int,
                     pad_value: int = 0):
    """Run ``decode_fn(local_item_indices) -> (rows, lens)`` on this rank's shard and gather the results.

    With ``decode_fn = lambda idx: model.generate_batch(tokens[idx], codes[idx], ...)`` this is BASELINE config 4
    (full-batch TTS sharded over 2/4/8 GPUs)."""
    rank, ws = world()
    idx = shard_indices(n_items, rank, ws)
    rows, lens = decode_fn(idx)
    return gather_ragged(rows, lens, n_items, pad_value)


# ---- training: batch-sharded data parallelism (BASELINE config 5) -------------------------------------------
def allreduce_gradients(model: torch.nn.Module, bucket_bytes: int = 64 << 20) -> int:
    """Average ``param.grad`` over the ranks: the ONE exchange step of the batch-sharded training step.

    Gradients are flattened into buckets of ~``bucket_bytes`` (fp32) and summed with ``all_reduce`` (NCCL over
    NVLink/NVSwitch on GPUs -- bucket size picked for launch latency, not link count -- gloo in the CPU tests), then
    divided by the world size; parameters without a gradient on this rank (e.g. other NAR stages) contribute zeros so that
    every rank issues the same collectives.  Returns the number of buckets reduced."""
    rank, ws = world()
    params = [p for _, p in sorted(model.named_parameters(), key=lambda kv: kv[0]) if p.requires_grad]
    if ws == 1 or not params:
        return 0
    buckets, cur, cur_bytes = [], [], 0
    for p in params:
        cur.append(p)
        cur_bytes += p.numel() * 4
        if cur_bytes >= bucket_bytes:
            buckets.append(cur)
            cur, cur_bytes = [], 0
    if cur:
        buckets.append(cur)
    for bucket in buckets:
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(ws)
        off = 0
        for p in bucket:
            n = p.numel()
            g = flat[off:off + n].view_as(p).to(p.dtype)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n
    return len(buckets)


def shard_batch(batch: dict, rank: int, world_size: int) -> dict:
    """Contiguous per-rank slice of a teacher-forced batch dict (tensors whose first dim is the batch)."""
    B = next(v.shape[0] for v in batch.values() if torch.is_tensor(v) and v.dim() >= 1)
    lo, hi = (B * rank) // world_size, (B * (rank + 1)) // world_size
    return {k: (v[lo:hi] if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == B else v) for k, v in batch.items()}
