"""Checkpoint save / load with the reference's parameter names (SURVEY 5 "Checkpoint / resume", 8b state_dict keys).

The reference never writes a checkpoint itself: Lightning's default ``ModelCheckpoint`` does, under the logger directory, and
``config.ckpt_path`` is created but unused (config.py:74-75).  The file format here is Lightning's so that checkpoints
interchange in BOTH directions: a dict with ``state_dict`` (the model's reference-keyed ``state_dict()``), ``optimizer_states``,
``lr_schedulers``, ``global_step``, ``epoch`` and ``hyper_parameters`` -- ``ValleAR.load_state_dict(ckpt['state_dict'])`` works on
either side because the module tree (and therefore every key and shape) is the reference's.
"""
from __future__ import annotations

import dataclasses
import os
from pathlib import Path

import torch


def save_checkpoint(path, model, optimizer=None, scheduler=None, *, global_step: int = 0, epoch: int = 0, extra: dict | None = None) -> Path:
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    cfg = getattr(model, 'config', None)
    hp = {}
    if dataclasses.is_dataclass(cfg):
        hp = {k: (str(v) if isinstance(v, Path) else v) for k, v in dataclasses.asdict(cfg).items()}
    ckpt = {
        'state_dict': {k: v.detach().cpu() for k, v in model.state_dict().items()},
        'optimizer_states': [optimizer.state_dict()] if optimizer is not None else [],
        'lr_schedulers': [scheduler.state_dict()] if scheduler is not None else [],
        'global_step': int(global_step), 'epoch': int(epoch), 'hyper_parameters': hp,
        'model_class': type(model).__name__,
    }
    if extra:
        ckpt.update(extra)
    tmp = path.with_suffix(path.suffix + f'.tmp{os.getpid()}')
    torch.save(ckpt, tmp)
    os.replace(tmp, path)                       # atomic: a crash never leaves a truncated checkpoint under the final name
    return path


def load_checkpoint(path, model, optimizer=None, scheduler=None, *, strict: bool = True, map_location='cpu') -> dict:
    """Restores model (and optimizer / scheduler when given) in place; returns the remaining metadata."""
    ckpt = torch.load(Path(path), map_location=map_location, weights_only=False)
    state = ckpt['state_dict'] if 'state_dict' in ckpt else ckpt        # a bare state_dict file loads too
    model.load_state_dict(state, strict=strict)
    if optimizer is not None and ckpt.get('optimizer_states'):
        optimizer.load_state_dict(ckpt['optimizer_states'][0])
    if scheduler is not None and ckpt.get('lr_schedulers'):
        scheduler.load_state_dict(ckpt['lr_schedulers'][0])
    return {k: v for k, v in ckpt.items() if k not in ('state_dict', 'optimizer_states', 'lr_schedulers')}
