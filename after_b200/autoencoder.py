"""Reference-shaped codec surface: ``AutoEncoder.encode`` / ``decode``
(after/autoencoder/networks/SimpleNetsStream.py:918-954, as exported by
after_scripts/export_autoencoder.py:251-265 -- ``encode`` returns ``z`` only)."""
from __future__ import annotations

from .engine import Engine


class AutoEncoder:

    def __init__(self, engine: Engine):
        if not engine.has_codec:
            raise RuntimeError("engine was created without autoencoder weights")
        self.engine = engine

    @property
    def ratio(self) -> int:
        return self.engine.ae_ratio

    def encode(self, x):
        """(B, 1, samples) -> (B, z_channels, samples / ratio)"""
        return self.engine.ae_encode(x)

    def decode(self, z):
        """(B, z_channels, T) -> (B, 1, T * ratio)"""
        return self.engine.ae_decode(z)

    def forward(self, x):
        return self.decode(self.encode(x))

    __call__ = forward

    def eval(self):
        return self
