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
def _buckets_of(model: torch.nn.Module) -> list[list[tuple[str, torch.nn.Parameter]]]:
    """Parameters grouped the way the backward pass finishes them: one bucket per transformer layer (last layer first is the
    order they complete in), everything else (embeddings, projections, stage embeddings) in a final bucket."""
    layers: dict[int, list] = {}
    rest = []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        parts = name.split('.')
        if len(parts) > 2 and parts[0] == 'transformer' and parts[1] == 'layers' and parts[2].isdigit():
            layers.setdefault(int(parts[2]), []).append((name, p))
        else:
            rest.append((name, p))
    return [layers[k] for k in sorted(layers)] + [rest]


class GradReducer:
    """Gradient averaging of the batch-sharded training step that OVERLAPS the backward pass.

    One persistent flat fp32 buffer holds every parameter's gradient (no per-step ``torch.cat`` / copy-back); it is cut into
    one bucket per transformer layer + one for the remaining parameters.  The hand-written backward (valle2_b200/train.py)
    hands each layer's gradients over as soon as that layer is done (``submit``): they are copied into the bucket and the
    bucket's ``all_reduce`` is enqueued at once (``async_op``: NCCL runs it on its own stream, ordered after the copies), so the
    exchange of layer l travels over NVLink while layers l-1 .. 0 are still being differentiated.  ``finish`` waits for the
    outstanding reductions and returns views of the averaged gradients.  A has-gradient flag per parameter is summed along
    with the last bucket: a parameter that received no gradient on ANY rank (e.g. the projections of the NAR stages that were
    not drawn) keeps ``grad = None`` (``drop_unused``), exactly as in a single-process step."""

    def __init__(self, model: torch.nn.Module):
        self.rank, self.ws = world()
        self.buckets = _buckets_of(model)
        dev = next(model.parameters()).device
        self.names = [n for b in self.buckets for n, _ in b]
        n_total = sum(p.numel() for b in self.buckets for _, p in b)
        n_total = (n_total + 3) // 4 * 4            # the flag vector starts 16-byte aligned
        self.flat = torch.zeros(n_total + len(self.names), device=dev, dtype=torch.float32)
        self.views, self.ranges, self.bucket_of = {}, [], {}
        off = 0
        for bi, bucket in enumerate(self.buckets):
            start = off
            for name, p in bucket:
                self.views[name] = self.flat[off:off + p.numel()].view_as(p)
                self.bucket_of[name] = bi
                off += p.numel()
            self.ranges.append((start, off))
        self.flags = self.flat[n_total:]
        self.flag_index = {n: i for i, n in enumerate(self.names)}
        self.ranges[-1] = (self.ranges[-1][0], self.flat.numel())       # the flags travel with the last bucket
        self._work, self._submitted = [], set()
        self.avg_op = dist.ReduceOp.AVG if (self.ws > 1 and dist.get_backend() == 'nccl') else None
        self.no_comm = False                    # measurement aid (tools/train_dp_probe.py): bucket copies without the collectives

    def begin(self) -> None:
        self._has = [0.0] * len(self.names)
        self._work, self._submitted = [], set()

    def _reduce(self, bi: int) -> None:
        lo, hi = self.ranges[bi]
        if self.ws > 1 and not self.no_comm:
            self._work.append(dist.all_reduce(self.flat[lo:hi], op=self.avg_op or dist.ReduceOp.SUM, async_op=True))

    def submit(self, bucket: int, grads: dict) -> None:
        """Gradients of one finished bucket ({parameter name: tensor or None}); missing names count as 'no gradient'."""
        assert bucket not in self._submitted, f'bucket {bucket} submitted twice'
        dst, src = [], []
        for name, _ in self.buckets[bucket]:
            g = grads.get(name)
            if g is None:
                self.views[name].zero_()
            else:
                dst.append(self.views[name])
                src.append(g.reshape(self.views[name].shape))
                self._has[self.flag_index[name]] = 1.0
        if dst:
            torch._foreach_copy_(dst, src)          # one multi-tensor copy per bucket
        if bucket == len(self.buckets) - 1:         # the flags travel with the last bucket (one small H2D copy per step)
            self.flags.copy_(torch.tensor(self._has, dtype=torch.float32), non_blocking=True)
        self._submitted.add(bucket)
        self._reduce(bucket)

    def finish(self) -> dict:
        """Wait for every reduction; returns {name: averaged gradient (a view of the flat buffer)}."""
        assert len(self._submitted) == len(self.buckets), 'not every bucket was submitted'
        for w in self._work:
            w.wait()
        if self.ws > 1 and self.avg_op is None and not self.no_comm:    # gloo has no AVG: sum, then divide (the flags stay sums)
            self.flat[:self.flat.numel() - len(self.names)].div_(self.ws)
        self._work = []
        return dict(self.views)

    def drop_unused(self, model: torch.nn.Module) -> int:
        """After ``loss.backward()``: parameters whose has-gradient flag is zero on every rank get ``grad = None`` (one host
        read of the flag vector).  Returns how many were dropped."""
        flags = self.flags.tolist()
        dropped = 0
        for name, p in model.named_parameters():
            i = self.flag_index.get(name)
            if i is not None and flags[i] == 0.0 and p.grad is not None:
                p.grad = None
                dropped += 1
        return dropped


_ACTIVE: list = []


class reducing:
    """``with parallel.reducing(reducer): loss = model.training_step(batch)`` -- the training step's backward hands its
    gradients to ``reducer`` layer by layer (overlapped all-reduce); outside the context the step is purely local."""

    def __init__(self, reducer: GradReducer | None):
        self.reducer = reducer

    def __enter__(self):
        _ACTIVE.append(self.reducer)
        return self.reducer

    def __exit__(self, *exc):
        _ACTIVE.pop()
        return False


def active_reducer() -> GradReducer | None:
    return _ACTIVE[-1] if _ACTIVE else None


def allreduce_gradients(model: torch.nn.Module, bucket_bytes: int = 64 << 20) -> int:
    """Average ``param.grad`` over the ranks AFTER the backward pass (no overlap): used with gradient accumulation, where the
    exchange belongs to the last micro-batch only.  Gradients are packed into fp32 buckets of ~``bucket_bytes`` and summed
    with ``all_reduce``; a has-gradient flag per parameter travels along so that a parameter without a gradient on every rank
    keeps ``grad = None`` (single-process behaviour) while one that has a gradient on some ranks gets zeros from the others.
    Returns the number of buckets reduced."""
    rank, ws = world()
    params = [p for _, p in sorted(model.named_parameters(), key=lambda kv: kv[0]) if p.requires_grad]
    if ws == 1 or not params:
        return 0
    dev = params[0].device
    flags = torch.tensor([0.0 if p.grad is None else 1.0 for p in params], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.SUM)
    has = flags.tolist()
    buckets, cur, cur_bytes = [], [], 0
    for p, h in zip(params, has):
        if h == 0.0:
            continue                               # no rank has a gradient: leave grad = None
        cur.append(p)
        cur_bytes += p.numel() * 4
        if cur_bytes >= bucket_bytes:
            buckets.append(cur)
            cur, cur_bytes = [], 0
    if cur:
        buckets.append(cur)
    for bucket in buckets:
        flat = torch.empty(sum(p.numel() for p in bucket), device=dev, dtype=torch.float32)
        off = 0
        for p in bucket:
            n = p.numel()
            if p.grad is None:
                flat[off:off + n].zero_()
            else:
                flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(ws)
        off = 0
        for p in bucket:
            n = p.numel()
            g = flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = g.to(p.dtype).clone()
            else:
                p.grad.copy_(g)
            off += n
    return len(buckets)


def shard_batch(batch: dict, rank: int, world_size: int) -> dict:
    """Contiguous per-rank slice of a teacher-forced batch dict (tensors whose first dim is the batch)."""
    B = next(v.shape[0] for v in batch.values() if torch.is_tensor(v) and v.dim() >= 1)
    lo, hi = (B * rank) // world_size, (B * (rank + 1)) // world_size
    return {k: (v[lo:hi] if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == B else v) for k, v in batch.items()}
