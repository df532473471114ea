"""Reference-shaped model-call surface for the diffusion side (after/diffusion/model.py,
after/diffusion/networks/transformerv2.py, encoder.py): same class and method names, same
argument meaning, torch tensors on ``cuda:k`` in and out.  All arithmetic happens in
libafter_b200 (no torch ops on the data path, no fallback).
"""
from __future__ import annotations

import torch

from . import _lib as L
from .engine import Engine


class DenoiserV2:
    """``net(x, time=, cond=, time_cond=, cache_index=0)`` + ``roll_cache`` -- transformerv2.py:514-543.

    Like the reference module, which path runs is a construction-time property: an engine created with
    ``max_cache_size == 0`` is the offline model (``cache_index`` is accepted and, as in the reference, irrelevant); with
    ``max_cache_size > 0`` every forward attends to -- and ``roll_cache`` extends -- the KV history of ``cache_index``."""

    def __init__(self, engine: Engine):
        if not engine.has_denoiser:
            raise RuntimeError("engine was created without denoiser weights")
        self.engine = engine

    def __call__(self, x, time, cond, time_cond, cache_index: int = 0):
        return self.forward(x, time=time, cond=cond, time_cond=time_cond, cache_index=cache_index)

    def forward(self, x, time, cond, time_cond, cache_index: int = 0):
        return self.engine.denoiser_forward(x, time, cond, time_cond, cache_index if self.engine.streaming else None)

    def roll_cache(self, size: int, cache_index: int = 0):
        """transformerv2.py:514-515.  The offline model keeps no cache (the reference's ``max_cache_size = 0`` modules
        would fail here; exported offline models never call it): nothing to roll."""
        if self.engine.streaming:
            self.engine.roll_cache(size, cache_index)

    def reset_cache(self):
        if self.engine.streaming:
            self.engine.reset_cache()

    def eval(self):
        return self


class UNET1D:
    """``net(x, time=, time_cond=, cond=)`` of the conv denoiser -- unet1d.py:376-429.  An engine created with ``unet`` /
    ``unet_state`` runs ``RectifiedFlow.sample`` / ``model_forward`` over it."""

    def __init__(self, engine: Engine):
        if not engine.has_unet:
            raise RuntimeError("engine was created without UNET1D weights")
        self.engine = engine

    def __call__(self, x, time=None, time_cond=None, cond=None, cache_index: int = 0):
        return self.forward(x, time=time, time_cond=time_cond, cond=cond)

    def forward(self, x, time=None, time_cond=None, cond=None):
        return self.engine.unet_forward(x, time, cond, time_cond)

    def eval(self):
        return self


class Encoder1D:
    """Structure encoder ``encoder_time(z)`` -- encoder.py:273-298."""

    def __init__(self, engine: Engine):
        if not engine.has_structure:
            raise RuntimeError("engine was created without structure-encoder weights")
        self.engine = engine

    def __call__(self, z):
        return self.forward(z)

    def forward(self, z):
        return self.engine.structure_encode(z)

    def forward_stream(self, z, slot: int = 0):
        """encoder.py:300-322 under ``cc.use_cached_conv(True)`` (after_scripts/export.py:14-17): every conv carries its left
        context across calls.  Needs an engine created with ``stream_slots > 0``; an offline engine has no such state and,
        like the reference module built without cached convs, processes the buffer on its own."""
        if self.engine.stream_slots > 0:
            return self.engine.structure_encode_stream(slot, z)
        return self.engine.structure_encode(z)

    def eval(self):
        return self


class ECAPATDNN:
    """Timbre encoder ``encoder(z)`` -- ecapa_encoder.py:567-624: (B, C, T) -> (B, zt)."""

    def __init__(self, engine: Engine):
        if not engine.has_timbre:
            raise RuntimeError("engine was created without timbre-encoder weights")
        self.engine = engine

    def __call__(self, z):
        return self.forward(z)

    def forward(self, z):
        return self.engine.timbre_encode(z)

    forward_stream = forward  # identical arithmetic in the reference (ecapa_encoder.py:626-666)

    def eval(self):
        return self


class RectifiedFlow:
    """``RectifiedFlow(net=, sr=, encoder=, encoder_time=, emb_model=, drop_value=)`` -- model.py:570, 721-785.

    ``cfg_variant`` / ``clamp`` select between the audio model's guidance layout (model.py:730-759) and the
    MIDI streamer's (export_midi.py:322-360)."""

    def __init__(self, net: DenoiserV2, sr: int = 44100, encoder=None, encoder_time=None, post_encoder=None,
                 classifier=None, emb_model=None, time_transform=None, drop_value: float = -4.0, drop_rate: float = 0.2,
                 device=None, cfg_variant: int = L.CFG_AUDIO, clamp: float = 0.01, **kwargs):
        self.net = net
        self.sr = sr
        self.encoder = encoder
        self.encoder_time = encoder_time
        self.post_encoder = post_encoder
        self.classifier = classifier
        self.emb_model = emb_model
        self.time_transform = time_transform
        if abs(drop_value - net.engine.cfg.drop_value) > 0:
            raise ValueError("drop_value differs from the one the engine was created with")
        self.drop_value = drop_value
        self.drop_rate = drop_rate
        self.cfg_variant = cfg_variant
        self.clamp = clamp

    @property
    def device(self):
        return self.net.engine.device

    def eval(self):
        return self

    def sample_prior(self, x0_shape):
        return torch.randn(x0_shape, device=self.device)  # model.py:141-142

    def model_forward(self, x, time, cond, time_cond, guidance_timbre: float, guidance_structure: float,
                      cache_index: int = 0):
        eng = self.net.engine
        return eng.model_forward(x, time, cond, time_cond, guidance_timbre, guidance_structure, self.cfg_variant, self.clamp,
                                 cache_index if eng.streaming else None)

    @torch.no_grad()
    def sample(self, x0, cond, time_cond, nb_steps: int, guidance_timbre: float = 1.0, guidance_structure: float = 1.0):
        return self.net.engine.sample(x0.to(self.device), cond, time_cond, nb_steps, guidance_timbre,
                                      guidance_structure, self.cfg_variant, self.clamp)
