"""``Streamer``: the nn_tilde-facing method / attribute surface of ``after_scripts/export.py:145-507``, on top of an
``Engine``.  Same method names, channel counts, ratios and attribute setters, so host code written against the exported
model (notebooks/audio_to_audio_demo.ipynb cell 9, ``nn~ <model> generate_timbre 8192``) runs against it.

Semantics.  On an engine created with ``stream_slots >= 2`` and ``max_cache_size > 0`` every stage carries the state the
exported model carries: the two codec copies (``emb_model_structure`` = slot 0, ``emb_model_timbre`` = slot 1,
``export.py:161-168``) run the streaming export of the codec (cached-conv encoder, CachedGroupNorm stream branch,
overlap-add decoder: ``export_autoencoder.py:16-153``), ``structure`` runs ``Encoder1D.forward_stream`` with cached convs,
the denoiser keeps one rolling KV history per diffusion step (``export.py:398-416``), and the timbre encoder sees the rolling
``previous_timbre`` latent buffer (``ECAPATDNN.forward_stream`` is stateless).  On an engine without streaming state the
corresponding stage processes each buffer on its own with the offline kernels (what the reference computes with
``cc.use_cached_conv(False)``).  A freshly created engine starts from zero state; the exported ``.ts`` starts from the state
its export script left behind after its silent test passes -- ``prime_like_export()`` replays those.
TorchScript serialisation (``export_to_ts``) and the export-time TRAINING of the latent-map projection are out of scope;
``latent2map`` / ``map2latent`` apply a given projection (``Engine(latent_map_state=...)``) or the identity.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L
from .engine import Engine


class Streamer:

    def __init__(self, engine: Engine, n_signal_timbre: int = 64, chunk_size: int = 4, latent_range: float = 1.0):
        if not (engine.has_denoiser and engine.has_codec and engine.has_structure and engine.has_timbre):
            raise RuntimeError("Streamer needs denoiser, autoencoder, structure-encoder and timbre-encoder weights")
        self.engine = engine
        self.chunk_size = chunk_size
        self.n_signal_timbre = n_signal_timbre
        self.latent_range = latent_range
        self.zs_channels = engine.cfg.tcond_dim
        self.zt_channels = engine.cfg.cond_dim
        self.ae_latents = engine.cfg.n_channels
        self.ae_ratio = engine.ae_ratio
        self.drop_value = engine.cfg.drop_value
        self.sr = 44100
        self.zt_buffer = n_signal_timbre * self.ae_ratio
        # attributes are 1-tuples with get_/set_ accessors (export.py:180-182, 330-355)
        self.nb_steps = (1, )
        self.guidance_timbre = (1.0, )
        self.guidance_structure = (1.0, )
        # rolling latent history fed to the timbre encoder (export.py:185-187, 418-429)
        self.previous_timbre = torch.zeros(4, self.ae_latents, n_signal_timbre, device=engine.device)
        # (in_channels, in_ratio, out_channels, out_ratio) per method (export.py:190-328)
        r = self.ae_ratio
        self.methods = {
            "forward": (2, 1, 1, 1),
            "structure": (1, 1, self.zs_channels, r),
            "timbre": (1, 1, self.zt_channels, r),
            "diffuse": (self.zs_channels + self.zt_channels, r, self.ae_latents, r),
            "generate": (self.zs_channels + self.zt_channels, r, 1, 1),
            "generate_timbre": (self.zt_channels + 1, 1, 1, 1),
            "decode": (self.ae_latents, r, 1, 1),
        }

    # ---- attributes ---------------------------------------------------------------------------
    def get_guidance_timbre(self) -> float:
        return self.guidance_timbre[0]

    def set_guidance_timbre(self, guidance_timbre: float) -> int:
        self.guidance_timbre = (float(guidance_timbre), )
        return 0

    def get_guidance_structure(self) -> float:
        return self.guidance_structure[0]

    def set_guidance_structure(self, guidance_structure: float) -> int:
        self.guidance_structure = (float(guidance_structure), )
        return 0

    def get_nb_steps(self) -> int:
        return self.nb_steps[0]

    def set_nb_steps(self, nb_steps: int) -> int:
        self.nb_steps = (int(nb_steps), )
        return 0

    # ---- methods --------------------------------------------------------------------------------
    def _check(self, name: str, x: torch.Tensor):
        cin = self.methods[name][0]
        if x.dim() != 3 or x.shape[1] != cin:
            raise ValueError(f"{name} expects (n_batch, {cin}, buffer), got {tuple(x.shape)}")
        return x.to(self.engine.device, torch.float32).contiguous()

    def sample(self, x_last, cond, time_cond):
        """export.py:398-416; the streamer clamps the guidance ratio at 0.1 (:389-390).  On an engine created with
        ``max_cache_size > 0`` Euler step i attends to, and then rolls, KV history i (the exported model's behaviour);
        otherwise the block is sampled with the offline kernels."""
        if self.engine.streaming:
            return self.engine.sample_stream(x_last, cond, time_cond, self.nb_steps[0], self.guidance_timbre[0],
                                             self.guidance_structure[0], cfg_variant=L.CFG_AUDIO, clamp=0.1)
        return self.engine.sample(x_last, cond, time_cond, self.nb_steps[0], self.guidance_timbre[0],
                                  self.guidance_structure[0], cfg_variant=L.CFG_AUDIO, clamp=0.1)

    # codec copies of the exported model (export.py:161-168): slot 0 = emb_model_structure, slot 1 = emb_model_timbre
    def _encode(self, slot: int, x):
        if self.engine.stream_slots > slot:
            return self.engine.ae_encode_stream(slot, x)
        return self.engine.ae_encode(x)

    def prime_like_export(self, n_batch: int = 1):
        """A loaded ``export_stream.ts`` does not start from zero state: the export scripts ran silence through it.
        Replayed here, per codec copy: the two 131072-sample (64-frame) silent ``encode`` passes of
        export_autoencoder.py (:49-51 constructor, :314-315 main) and the ``decode`` of the second one (:316), then the
        16384-sample ``encode`` of export.py:173-174 on the structure copy.  NOT replayed: the constructor's plain
        ``model.decode`` of the throw-away offline encoder's latents (:51), which only seeds the decoder's GroupNorm
        history -- that history is fully overwritten after 64 latent frames (12 s) of real audio."""
        eng = self.engine
        if eng.stream_slots < 2:
            raise RuntimeError("prime_like_export needs an engine with stream_slots >= 2")
        if (eng.cfg.stream_max_frames or 64) < 64:
            raise RuntimeError("prime_like_export replays 64-frame buffers: create the engine with stream_max_frames >= 64")
        silence = torch.zeros(n_batch, 1, 64 * self.ae_ratio, device=eng.device)
        for slot in (0, 1):
            eng.ae_encode_stream(slot, silence)
            z = eng.ae_encode_stream(slot, silence)
            eng.ae_decode_stream(slot, z)
        eng.ae_encode_stream(0, silence[..., :8 * self.ae_ratio].contiguous())

    def timbre(self, x):
        x = self._check("timbre", x)
        z = self._encode(1, x)
        n, t = z.shape[0], z.shape[-1]
        hist = torch.cat((self.previous_timbre[:n], z), -1)[..., t:]
        self.previous_timbre[:n] = hist
        zsem = self.engine.timbre_encode(self.previous_timbre[:n].contiguous()) / self.latent_range
        return zsem.unsqueeze(-1).repeat(1, 1, self.chunk_size)

    def structure(self, x):
        x = self._check("structure", x)
        z = self._encode(0, x)
        if self.engine.stream_slots > 0:
            return self.engine.structure_encode_stream(0, z)  # encoder_time.forward_stream (export.py:431-435)
        return self.engine.structure_encode(z)

    def diffuse(self, x, noise: Optional[torch.Tensor] = None):
        """(n, zs + zt, T) -> (n, latents, T).  As in the reference only batch row 0 is sampled and the result is
        repeated (export.py:437-449).  ``noise``: optional prior (n, latents, T); by default it is drawn from the global
        torch RNG on the host, like the reference's ``torch.randn``."""
        x = self._check("diffuse", x)
        n, T = x.shape[0], x.shape[-1]
        zsem = x[:, -self.zt_channels:].mean(-1) * self.latent_range
        time_cond = x[:, :self.zs_channels]
        if noise is None:
            noise = torch.randn(n, self.ae_latents, T)
        noise = noise.to(self.engine.device, torch.float32)
        out = self.sample(noise[:1].contiguous(), zsem[:1].contiguous(), time_cond[:1].contiguous())
        return out.repeat(n, 1, 1) if n > 1 else out

    def diffuse_timbre(self, x, noise: Optional[torch.Tensor] = None):
        x = self._check("generate_timbre", x)
        n = x.shape[0]
        zsem = x[:, 1:].mean(-1) * self.latent_range
        time_cond = self.structure(x[:, :1].contiguous())
        if noise is None:
            noise = torch.randn(n, self.ae_latents, time_cond.shape[-1])
        noise = noise.to(self.engine.device, torch.float32)
        out = self.sample(noise[:1].contiguous(), zsem[:1].contiguous(), time_cond[:1].contiguous())
        return out.repeat(n, 1, 1) if n > 1 else out

    def decode(self, x):
        x = self._check("decode", x)
        if self.engine.stream_slots > 0:
            return self.engine.ae_decode_stream(0, x)  # emb_model_structure.decode (export.py:451-455)
        return self.engine.ae_decode(x)

    def generate(self, x, noise: Optional[torch.Tensor] = None):
        return self.decode(self.diffuse(x, noise))

    def generate_timbre(self, x, noise: Optional[torch.Tensor] = None):
        return self.decode(self.diffuse_timbre(x, noise))

    def forward(self, x, noise: Optional[torch.Tensor] = None):
        x = self._check("forward", x)
        structure = self.structure(x[:, :1].contiguous())
        timbre = self.timbre(x[:, 1:].contiguous())
        if timbre.shape[-1] != structure.shape[-1]:  # chunk_size frames vs buffer / ratio frames
            timbre = timbre[..., :1].repeat(1, 1, structure.shape[-1])
        return self.decode(self.diffuse(torch.cat((structure, timbre), 1), noise))

    __call__ = forward

    def latent2map(self, x):
        """export.py:503-508: time average of the timbre latent -> ``project_model.encoder`` -> repeated over the buffer
        (the identity without a projection, as exported with ``--nolatent_project``)."""
        return self.engine.latent_map(x, 0)

    def map2latent(self, x):
        """export.py:496-501: time average of the 2-D map position -> ``project_model.decoder`` -> repeated."""
        return self.engine.latent_map(x, 1)


class MidiStreamer:
    """The MIDI variant of the exported model (``after_scripts/export_midi.py:150-450``): the structure condition is a
    piano roll rasterised from ``n_poly`` (pitch, velocity) signal pairs instead of an encoded audio stream, the CFG rows
    are (cond, roll) / (cond, drop) / (drop, drop) with ``f = g_structure / max(g_timbre, 0.1)`` (:322-360), and ``timbre``
    repeats the embedding over the encoded frames (:383-398).  Methods: ``timbre`` (1 -> zt), ``diffuse`` /
    ``generate`` (2 n_poly + zt -> latents / audio), ``decode``.  Same semantics notes as ``Streamer``."""

    def __init__(self, engine: Engine, n_poly: int = 4, n_signal_timbre: int = 64, chunk_size: int = 4, latent_range: float = 1.0):
        if not (engine.has_denoiser and engine.has_codec and engine.has_timbre):
            raise RuntimeError("MidiStreamer needs denoiser, autoencoder and timbre-encoder weights")
        if engine.cfg.tcond_dim != 128:
            raise RuntimeError("MidiStreamer needs a midi model (structure condition = 128-pitch piano roll)")
        self.engine = engine
        self.n_poly = n_poly
        self.chunk_size = chunk_size
        self.n_signal_timbre = n_signal_timbre
        self.latent_range = latent_range
        self.zt_channels = engine.cfg.cond_dim
        self.ae_latents = engine.cfg.n_channels
        self.ae_ratio = engine.ae_ratio
        self.drop_value = engine.cfg.drop_value
        self.sr = 44100
        self.zt_buffer = n_signal_timbre * self.ae_ratio
        self.nb_steps = (1, )
        self.guidance_timbre = (1.0, )
        self.guidance_structure = (1.0, )
        self.previous_timbre = torch.zeros(4, self.ae_latents, n_signal_timbre, device=engine.device)
        self.last_zsem = torch.zeros(4, self.zt_channels, device=engine.device)  # export_midi.py:193
        r = self.ae_ratio
        cin = 2 * n_poly + self.zt_channels
        self.methods = {
            "timbre": (1, 1, self.zt_channels, r),
            "generate": (cin, r, 1, 1),
            "diffuse": (cin, r, self.ae_latents, r),
            "decode": (self.ae_latents, r, 1, 1),
        }

    get_guidance_timbre = Streamer.get_guidance_timbre
    set_guidance_timbre = Streamer.set_guidance_timbre
    get_guidance_structure = Streamer.get_guidance_structure
    set_guidance_structure = Streamer.set_guidance_structure
    get_nb_steps = Streamer.get_nb_steps
    set_nb_steps = Streamer.set_nb_steps
    _check = Streamer._check

    def sample(self, x_last, cond, time_cond):
        """export_midi.py:362-381 (per-step KV caches when the engine streams)."""
        fn = self.engine.sample_stream if self.engine.streaming else self.engine.sample
        return fn(x_last, cond, time_cond, self.nb_steps[0], self.guidance_timbre[0], self.guidance_structure[0],
                  cfg_variant=L.CFG_MIDI, clamp=0.1)

    def timbre(self, x):
        """export_midi.py:383-398; the single codec copy ``emb_model_timbre`` (slot 0) streams when the engine has the state."""
        x = self._check("timbre", x)
        z = self.engine.ae_encode_stream(0, x) if self.engine.stream_slots > 0 else self.engine.ae_encode(x)
        n, t = z.shape[0], z.shape[-1]
        self.previous_timbre[:n] = torch.cat((self.previous_timbre[:n], z), -1)[..., t:]
        zsem = self.engine.timbre_encode(self.previous_timbre[:n].contiguous())
        return zsem.unsqueeze(-1).repeat(1, 1, t) / self.latent_range

    def piano_roll(self, notes: torch.Tensor) -> torch.Tensor:
        """(1, 2 n_poly, T) pitch / velocity signals -> (1, 128, T) roll, exactly as export_midi.py:408-415 fills it:
        voices in order (later voices overwrite), a frame is written when ITS velocity is > 0, the value is velocity / 128,
        and -- the reference indexes with the voice's whole pitch row, ``notes[:, 2 i].long()`` -- it is written at every
        pitch that voice takes anywhere in the buffer (one pitch for a held note)."""
        T = notes.shape[-1]
        roll = torch.zeros(1, 128, T, device=notes.device, dtype=torch.float32)
        for i in range(self.n_poly):
            pitch = notes[0, 2 * i].long()
            vel = notes[0, 2 * i + 1]
            on = vel > 0
            if bool(on.any()):
                rows = torch.unique(pitch)
                roll[0, rows[:, None], torch.nonzero(on)[:, 0][None, :]] = (vel[on] / 128)[None, :]
        return roll

    def diffuse(self, x, noise: Optional[torch.Tensor] = None):
        x = self._check("diffuse", x)
        n, T = x.shape[0], x.shape[-1]
        zsem = x[:, -self.zt_channels:].mean(-1) * self.latent_range
        time_cond = self.piano_roll(x[:1, :2 * self.n_poly])
        if noise is None:
            noise = torch.randn(n, self.ae_latents, T)
        noise = noise.to(self.engine.device, torch.float32)
        out = self.sample(noise[:1].contiguous(), zsem[:1].contiguous(), time_cond.contiguous())
        return out.repeat(n, 1, 1) if n > 1 else out

    def decode(self, x):
        x = self._check("decode", x)
        if self.engine.stream_slots > 0:
            return self.engine.ae_decode_stream(0, x)  # emb_model_timbre.decode (export_midi.py:426-429)
        return self.engine.ae_decode(x)

    def generate(self, x, noise: Optional[torch.Tensor] = None):
        return self.decode(self.diffuse(x, noise))

    def latent2map(self, x):
        """export.py:503-508: time average of the timbre latent -> ``project_model.encoder`` -> repeated over the buffer
        (the identity without a projection, as exported with ``--nolatent_project``)."""
        return self.engine.latent_map(x, 0)

    def map2latent(self, x):
        """export.py:496-501: time average of the 2-D map position -> ``project_model.decoder`` -> repeated."""
        return self.engine.latent_map(x, 1)
