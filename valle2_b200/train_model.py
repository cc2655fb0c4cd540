"""Training driver: the caller of ``training_step``, mirroring ``valle/train_model.py:13-44``.

    python -m valle.train_model -c cfg.json -m ValleAR                      # reference CLI (``args.config``: upstream reads a
                                                                            # non-existent ``args.hparams`` -- defect A-11)
    torchrun --nproc-per-node 8 -m valle.train_model -c cfg.json -m ValleAR --synthetic 64

``train(hparams_fp, model_name)`` keeps the reference's signature and order of operations: config from JSON, seed, model from
``get_model_class``, data loaders, optimisation.  What differs is what sits underneath:

* the reference hands the loop to ``lightning.Trainer``; Lightning is not part of this image, so the loop is written out
  (``fit``): ``configure_optimizers()`` -> for every batch ``training_step`` -> ``loss.backward()`` -> gradient averaging
  across ranks (the ONE exchange step of the batch-sharded step, NCCL over NVLink; overlapped with the backward pass by
  ``parallel.GradReducer``) -> ``clip_grad_norm_(gradient_clip_val)`` -> AdamW step, scheduler once per epoch (Lightning's default
  interval), ``grad_accum`` micro-batches per step, checkpoints in Lightning's file format (``valle2_b200/checkpoint.py``).  With Lightning
  installed the models remain ``LightningModule``s and ``L.Trainer.fit`` works on them unchanged;
* the reference's data pipeline (``valle/data.py``: HF ``datasets`` + EnCodec + g2p) is out of scope.  ``train`` takes
  any iterable of item lists through ``loaders=``; ``--synthetic N`` builds N random items of the wire format
  ``{'codes': (Q, T) int64, 'tokens': (Tx,) int64}`` and batches them with the collate of ``valle.collate`` -- the step before
  the hot path -- so that the whole chain collate -> training_step -> backward -> all-reduce -> optimizer runs end to end.
"""
from __future__ import annotations

import argparse
import time
from pathlib import Path

import torch

from . import parallel
from .collate import get_collate
from .config import ConfigValle
from .models import get_model_class


def synthetic_items(config: ConfigValle, n_items: int, seed: int, *, min_frames: int = 60, max_frames: int = 150):
    """Random dataset items in the reference's wire format (collate.py:24-31): more frames than phonemes."""
    g = torch.Generator().manual_seed(seed)
    items = []
    for _ in range(n_items):
        T = int(torch.randint(min_frames, max_frames + 1, (1,), generator=g))
        Tx = int(torch.randint(max(2, T // 6), max(3, T // 3), (1,), generator=g))
        items.append({'codes': torch.randint(0, config.num_audio_tokens, (config.num_quantizers, T), generator=g),
                      'tokens': torch.randint(1, config.vocab_size, (Tx,), generator=g)})
    return items


def batches(items: list, batch_size: int, collate_fn, rank: int, world_size: int):
    """Global batches of ``batch_size * world_size`` items; every rank collates its contiguous share (batch-sharded data
    parallelism)."""
    per_step = batch_size * world_size
    for i in range(0, len(items) - per_step + 1, per_step):
        yield collate_fn(items[i + rank * batch_size: i + (rank + 1) * batch_size])


def seed_everything(seed: int) -> None:
    """``lightning.seed_everything``: torch, Python ``random`` (the NAR stage draw, valle_nar.py:76) and numpy -- identical on
    every rank, so that all ranks of a step train the same NAR stage (and therefore the same set of parameters)."""
    import random

    import numpy as np
    random.seed(seed)
    np.random.seed(seed % (2 ** 32))
    torch.manual_seed(seed)


def fit(model, loader, config: ConfigValle, *, max_steps: int | None = None, log=print, ckpt_every: int = 0,
        resume: Path | None = None, model_name: str | None = None) -> list[float]:
    """The optimisation loop ``L.Trainer(max_steps, gradient_clip_val, accumulate_grad_batches).fit`` would run:
    * the LR scheduler advances once per EPOCH (one pass over ``loader()``) -- Lightning's default interval for the
      ``{'optimizer', 'lr_scheduler'}`` dict the reference returns (valle_ar.py:182-194);
    * data parallelism: with ``grad_accum == 1`` the gradient exchange overlaps the backward pass (``parallel.GradReducer``:
      per-layer buckets in a persistent flat buffer, all-reduce enqueued as each layer's backward finishes); with accumulation
      the exchange follows the last micro-batch (``parallel.allreduce_gradients``);
    * checkpoints in Lightning's file format under ``config.ckpt_path`` every ``ckpt_every`` optimizer steps and at the end
      (rank 0); ``resume`` restores model, optimizer, scheduler and the step counter."""
    from .checkpoint import load_checkpoint, save_checkpoint
    rank, world_size = parallel.world()
    opt = model.configure_optimizers()
    optimizer, scheduler = opt['optimizer'], opt.get('lr_scheduler')
    max_steps = config.max_steps if max_steps is None else max_steps
    accum = max(1, int(config.grad_accum))
    step0, epoch = 0, 0
    if resume is not None:
        meta = load_checkpoint(resume, model, optimizer, scheduler)
        step0, epoch = int(meta.get('global_step', 0)), int(meta.get('epoch', 0))
        if rank == 0:
            log(f'resumed from {resume} at step {step0}, epoch {epoch}')
    reducer = parallel.GradReducer(model) if (world_size > 1 and accum == 1) else None
    name = model_name or type(model).__name__

    def ckpt(step):
        if rank == 0:
            p = save_checkpoint(Path(config.ckpt_path) / f'{name}-step{step}.ckpt', model, optimizer, scheduler, global_step=step, epoch=epoch)
            log(f'checkpoint {p}')

    losses, micro, t0 = [], 0, time.time()
    t_prev = t0
    model.train()
    optimizer.zero_grad(set_to_none=True)
    while step0 + len(losses) < max_steps:
        progressed = False
        for batch in loader():
            progressed = True
            with parallel.reducing(reducer):
                loss = model.training_step(batch)
            (loss / accum).backward()
            micro += 1
            if micro % accum:
                continue
            if reducer is not None:
                reducer.drop_unused(model)
            elif world_size > 1:
                parallel.allreduce_gradients(model)
            if config.gradient_clip_val:
                torch.nn.utils.clip_grad_norm_(model.parameters(), config.gradient_clip_val)
            optimizer.step()
            optimizer.zero_grad(set_to_none=True)
            losses.append(float(loss.detach()))                      # the step's one host sync
            step = step0 + len(losses)
            t_now = time.time()
            if rank == 0 and (step % max(1, int(config.log_every_n_steps)) == 0 or step == max_steps):
                log(f'step {step}  loss {losses[-1]:.4f}  this step {(t_now - t_prev) * 1e3:.1f} ms  '
                    f'(mean {(t_now - t0) / len(losses) * 1e3:.1f} ms/step)')
            t_prev = t_now
            if ckpt_every and step % ckpt_every == 0 and step < max_steps:
                ckpt(step)
            if step >= max_steps:
                break
        if not progressed:
            raise ValueError('the training loader yielded no batch')
        epoch += 1
        if scheduler is not None:
            scheduler.step()                                         # Lightning: interval = 'epoch'
    if ckpt_every:
        ckpt(step0 + len(losses))
    return losses


def train(hparams_fp: Path, model_name: str, *, loaders=None, synthetic: int = 0, max_steps: int | None = None,
          device: str | None = None, log=print, frames: tuple[int, int] = (60, 150), ckpt_every: int = 0,
          resume: Path | None = None) -> list[float]:
    """``valle/train_model.py:13-36`` with the loop written out.  ``loaders`` = a callable returning an iterable of batch
    dicts (e.g. a ``DataLoader`` with ``collate_fn=get_collate(model_name)(config)``); ``synthetic`` > 0 builds one."""
    config = ConfigValle.from_json(hparams_fp)
    rank, world_size = parallel.world()
    seed_everything(config.seed)                                     # lightning.seed_everything(config.seed)
    model = get_model_class(model_name)(config)
    if device is None:
        device = f'cuda:{rank % max(1, torch.cuda.device_count())}'
    model = model.to(device)
    if loaders is None:
        if synthetic <= 0:
            raise RuntimeError('valle.data (HF datasets + EnCodec + g2p) is out of scope of valle2_b200: pass loaders= or --synthetic N')
        items = synthetic_items(config, synthetic, config.seed, min_frames=frames[0], max_frames=frames[1])
        collate_fn = get_collate(model_name)(config)
        loaders = lambda: batches(items, config.batch_size, collate_fn, rank, world_size)   # noqa: E731
    if rank == 0:
        log(f'Training model {model_name} on {world_size} rank(s), {sum(p.numel() for p in model.parameters()) / 1e6:.1f} M parameters')
    return fit(model, loaders, config, max_steps=max_steps, log=log, ckpt_every=ckpt_every, resume=resume, model_name=model_name)


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('-c', '--config', type=Path, required=True)
    parser.add_argument('-m', '--model', type=str, choices=['ValleAR', 'ValleNAR'], required=True)
    parser.add_argument('--synthetic', type=int, default=0, help='train on N random items instead of valle.data')
    parser.add_argument('--max-steps', type=int, default=None)
    parser.add_argument('--frames', type=int, nargs=2, default=(60, 150), help='frame range of the synthetic clips (75 frames = 1 s)')
    parser.add_argument('--ckpt-every', type=int, default=0, help='write a checkpoint under config.ckpt_path every N optimizer steps (and at the end)')
    parser.add_argument('--resume', type=Path, default=None, help='checkpoint to resume from')
    args = parser.parse_args(argv)
    import os
    started = False
    if int(os.environ.get('WORLD_SIZE', '1')) > 1 and not torch.distributed.is_initialized():
        if torch.cuda.is_available():
            torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        torch.distributed.init_process_group('nccl' if torch.cuda.is_available() else 'gloo')
        started = True
    try:
        train(args.config, args.model, synthetic=args.synthetic, max_steps=args.max_steps,       # upstream: args.hparams (A-11)
              frames=tuple(args.frames), ckpt_every=args.ckpt_every, resume=args.resume)
    finally:
        if started:
            torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
