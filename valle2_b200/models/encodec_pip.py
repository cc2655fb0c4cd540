"""EnCodec boundary -- mirror of the reference's ``valle/models/encodec_pip.py:6-131``: same class name, methods,
argument shapes and assertion messages.  The codec itself (a convolutional model from the un-vendored ``encodec`` pip package,
24 kHz, 8 codebooks at 6 kbps, hop 320 samples) is outside the hot path: this class is the WIRE-FORMAT adapter between it and
the decoders -- ``(Q, T)`` int64 codes on the codec side, ``(T, Q)`` on the model side (``ValleAR.generate`` /
``ValleNAR.generate`` take prompt codes as ``(T, Q)``, ``valle_ar.py:100-104``; the collate transposes too, ``collate.py``).

``EncodecPip()`` loads ``encodec.EncodecModel.encodec_model_24khz()`` exactly like the reference when the package (and its
weights) are present; offline it raises with that explanation.  ``EncodecPip(model=...)`` takes any object with the
``EncodecModel`` interface used here (``sample_rate``, ``encode``, ``decode``, ``encoder``, ``set_target_bandwidth``), which is
how the shape tests of the reference (``tests/test_encodec_pip.py``) run without the package (``tests/test_host_logic.py``).
"""
from __future__ import annotations

import torch


class EncodecPip:
    """Encodec model for audio coding and decoding (reference: encodec_pip.py:6-16)."""

    N_Q = 8                 # codebooks at the 6 kbps target bandwidth
    HOP = 320               # samples per frame at 24 kHz (75 frames/s)

    def __init__(self, model=None):
        if model is None:
            try:
                from encodec import EncodecModel
            except Exception as e:  # the package is not part of the offline image
                raise RuntimeError('EncodecPip needs the `encodec` package and its 24 kHz weights (not available offline); '
                                   'pass model= to use another EncodecModel-compatible codec') from e
            model = EncodecModel.encodec_model_24khz()
        self.model = model
        self.model.set_target_bandwidth(6.0)

    @property
    def sampling_rate(self) -> int:
        return self.model.sample_rate

    # -- the reference's methods (encodec_pip.py:23-131) ---------------------------------------------------------------
    @torch.inference_mode()
    def encode(self, audio: torch.Tensor) -> torch.Tensor:
        """1D audio [T] -> codes [N_Q, T]."""
        assert audio.dim() == 1, f'Expected 1D audio tensor, got {audio.dim()}D'
        frames = self.model.encode(audio.reshape(1, 1, -1))
        return torch.cat([enc[0] for enc in frames], dim=-1)[0]

    @torch.inference_mode()
    def batch_encode(self, audios: torch.Tensor) -> torch.Tensor:
        """2D audio [B, T] -> codes [B, N_Q, T]."""
        assert audios.dim() == 2, f'Expected 2D audio tensor, got {audios.dim()}D'
        frames = self.model.encode(audios.unsqueeze(1))
        return torch.cat([enc[0] for enc in frames], dim=-1)

    @torch.inference_mode()
    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        """codes [N_Q, T] -> 1D audio [T]."""
        assert codes.dim() == 2, f'Expected 2D codes tensor, got {codes.dim()}D'
        return self.model.decode([(codes.unsqueeze(0), None)]).reshape(-1)

    @torch.inference_mode()
    def batch_decode(self, codes: torch.Tensor) -> torch.Tensor:
        """codes [B, N_Q, T] -> 2D audio [B, T]."""
        assert codes.dim() == 3, f'Expected 3D codes tensor, got {codes.dim()}D'
        return self.model.decode([(codes, None)]).squeeze(1)

    @torch.inference_mode()
    def encode_decode(self, audio: torch.Tensor) -> torch.Tensor:
        return self.decode(self.encode(audio))

    @torch.inference_mode()
    def get_embedding(self, audio: torch.Tensor) -> torch.Tensor:
        """1D audio [T] -> encoder embedding [C, T]."""
        assert audio.dim() == 1, f'Expected 1D audio tensor, got {audio.dim()}D'
        return self.model.encoder(audio.reshape(1, 1, -1))[0]

    @torch.inference_mode()
    def batch_get_embedding(self, audios: torch.Tensor) -> torch.Tensor:
        """2D audio [B, T] -> encoder embedding [B, C, T]."""
        assert audios.dim() == 2, f'Expected 2D audio tensor, got {audios.dim()}D'
        return self.model.encoder(audios.unsqueeze(1))

    # -- wire format between the codec and the decoders ------------------------------------------------------------------
    @staticmethod
    def to_model_layout(codes: torch.Tensor) -> torch.Tensor:
        """Codec layout ``(Q, T)`` / ``(B, Q, T)`` -> decoder layout ``(T, Q)`` / ``(B, T, Q)`` int64 (prompt_codes of
        ``ValleAR.generate`` / ``ValleNAR.generate`` / ``tts.synthesize_batch``)."""
        assert codes.dim() in (2, 3), f'Expected (Q, T) or (B, Q, T) codes, got {codes.dim()}D'
        return codes.transpose(-1, -2).contiguous().long()

    @staticmethod
    def to_codec_layout(codes: torch.Tensor) -> torch.Tensor:
        """Decoder output ``(T, Q)`` / ``(B, T, Q)`` -> codec layout ``(Q, T)`` / ``(B, Q, T)`` int64 (input of ``decode``)."""
        assert codes.dim() in (2, 3), f'Expected (T, Q) or (B, T, Q) codes, got {codes.dim()}D'
        return codes.transpose(-1, -2).contiguous().long()

    def prompt_from_audio(self, audio: torch.Tensor) -> torch.Tensor:
        """1D prompt audio -> ``(T, Q)`` prompt codes for the decoders."""
        return self.to_model_layout(self.encode(audio))

    def audio_from_codes(self, codes_tq: torch.Tensor) -> torch.Tensor:
        """``(T, Q)`` codes produced by ``ValleNAR.generate`` / ``tts.synthesize_batch`` -> 1D audio."""
        return self.decode(self.to_codec_layout(codes_tq))
