"""LightningModule when lightning is installed (the reference's base class, valle_ar.py:14), else nn.Module
with a no-op ``log`` so the models stay constructible offline."""
import torch.nn as nn

try:  # pragma: no cover - lightning is absent in the build image
    import lightning as L
    BaseModule = L.LightningModule
except Exception:
    class BaseModule(nn.Module):
        def log(self, *args, **kwargs):
            return None
