"""Hyper-parameter container -- same field names, defaults and derived properties as the reference's
``valle/config.py:7-99`` so that reference callers (train_model.py:14-16) construct it unchanged."""
from __future__ import annotations

import dataclasses
import json
import pathlib


@dataclasses.dataclass
class ConfigValle:
    # data
    dataset: str = 'keithito/lj_speech'
    num_workers: int = 4
    # input features
    vocab_size: int = 256
    num_audio_tokens: int = 1024
    num_quantizers: int = 8
    sampling_rate: int = 16000
    polling_factor: int = 320
    # model
    d_model: int = 256
    n_heads: int = 4
    dim_feedforward: int = 1024
    dropout: float = 0.1
    activation: str = 'relu'            # accepted for compatibility; the FFN is always erf-GELU (modules.py:216)
    num_layers: int = 8
    norm: str = 'AdaptiveLayerNorm'
    # optimiser
    lr: float = 1e-4
    lr_warmup: int = 1000
    betas: tuple = (0.9, 0.98)
    weight_decay: float = 0.1
    use_fused_adam: bool = True
    gradient_clip_val: float = 1.0
    grad_accum: int = 1
    # generation
    max_audio_len: int = 1024
    num_beams: int = 4
    use_kv_cache: bool = True
    top_k: int = 50
    tok_p: float = 1.0
    temperature: float = 1.0
    length_penalty: float = 1.0
    # training
    seed: int = 42
    batch_size: int = 4
    valid_batch_size: int = 1
    max_steps: int = 1000
    log_every_n_steps: int = 100
    ckpt_path: pathlib.Path = pathlib.Path('models/checkpoints')
    log_path: pathlib.Path = pathlib.Path('models/logs')

    def __post_init__(self):
        if self.dataset is None:
            raise ValueError('Dataset must be provided')
        if self.norm not in ('AdaptiveLayerNorm', 'LayerNorm'):
            raise ValueError('Normalization layer must be AdaptiveLayerNorm or LayerNorm')
        if self.activation not in ('relu', 'gelu'):
            raise ValueError('Activation function must be relu or gelu')
        # the reference creates both directories on construction (config.py:74-77); kept for parity
        for name in ('ckpt_path', 'log_path'):
            path = pathlib.Path(getattr(self, name))
            path.mkdir(parents=True, exist_ok=True)
            setattr(self, name, path)

    @property
    def quantization_factor(self) -> int:
        return self.sampling_rate // self.polling_factor

    @property
    def bos_token(self) -> int:
        return self.num_audio_tokens + 1

    @property
    def eos_token(self) -> int:
        return self.num_audio_tokens

    @classmethod
    def from_dict(cls, hparams: dict) -> 'ConfigValle':
        return cls(**hparams)

    @classmethod
    def from_json(cls, path) -> 'ConfigValle':
        with open(path, encoding='utf-8') as fh:
            return cls.from_dict(json.load(fh))
