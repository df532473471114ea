"""``Engine``: one libafter_b200 handle on one GPU (weights + workspace), fed from reference
``state_dict``s.  The classes in ``after_b200.diffusion`` / ``after_b200.autoencoder`` are thin,
reference-shaped views over an Engine; torch tensors appear only at this boundary (device
pointers + the current CUDA stream are handed to the C ABI).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib as L
from .config import AutoEncoderConfig, DenoiserConfig, EcapaConfig, Encoder1DConfig, ModelConfig


def _fill_config(model: Optional[ModelConfig], ae: Optional[AutoEncoderConfig], max_batch: int, max_steps: int,
                 seq_len: Optional[int], max_samples: int, use_structure: bool, use_timbre: bool = False,
                 max_cache_size: int = 0, unet=None, stream_slots: int = 0, stream_max_frames: int = 0,
                 stream_gn_frames: int = 0) -> L.AfterConfig:
    c = L.AfterConfig()
    c.abi_version = L.ABI_VERSION
    c.max_cache_size = max_cache_size
    c.stream_slots = stream_slots
    c.stream_max_frames = stream_max_frames
    c.stream_gn_frames = stream_gn_frames
    d: DenoiserConfig = model.denoiser if model is not None else DenoiserConfig()
    c.n_channels = d.n_channels
    c.seq_len = seq_len if seq_len is not None else d.seq_len
    c.embed_dim = d.embed_dim
    c.cond_dim = d.cond_dim
    c.noise_embed_dims = d.noise_embed_dims
    c.n_layers = d.n_layers
    c.mlp_multiplier = d.mlp_multiplier
    c.tcond_dim = d.tcond_dim
    c.local_attention_size = d.local_attention_size
    c.attention_chunk_size = d.attention_chunk_size
    c.drop_value = model.drop_value if model is not None else -4.0
    c.max_batch = max_batch
    c.max_steps = max_steps
    if ae is not None:
        n = len(ae.factors)
        if n > L.MAX_STAGES:
            raise ValueError("too many codec stages")
        c.ae_in_channels = ae.in_channels
        c.ae_channels = ae.channels
        c.ae_z_channels = ae.z_channels
        c.ae_pqmf_bands = ae.pqmf_bands
        c.ae_n_stages = n
        for i, m in enumerate(ae.multipliers):
            c.ae_multipliers[i] = m
        for i, m in enumerate(ae.decoder_multipliers):
            c.ae_dec_multipliers[i] = m
        for i, f in enumerate(ae.factors):
            c.ae_factors[i] = f
        for i, dl in enumerate(ae.dilations[:ae.num_blocks]):
            c.ae_dilations[i] = dl
        c.ae_num_blocks = ae.num_blocks
        c.ae_kernel_size = ae.kernel_size
        c.ae_use_loudness = int(ae.use_loudness)
        c.ae_max_samples = max_samples
    se: Optional[Encoder1DConfig] = model.structure_encoder if (model is not None and use_structure) else None
    if se is not None:
        if any(r != 1 for r in se.ratios):
            raise ValueError("structure encoder ratios other than 1 are not supported (no shipped config uses them)")
        c.se_in_size = se.in_size
        c.se_n_blocks = len(se.channels)
        for i, ch in enumerate(se.channels):
            c.se_channels[i] = ch
        c.se_kernel_size = se.kernel_size
        c.se_causal = int(se.causal)
        c.se_use_tanh = int(se.use_tanh)
    te: Optional[EcapaConfig] = model.timbre_encoder if (model is not None and use_timbre) else None
    if te is not None:
        n = len(te.channels)
        if n > L.MAX_STAGES:
            raise ValueError("too many timbre-encoder stages")
        c.te_in_size = te.in_size
        c.te_n_blocks = n
        for i in range(n):
            c.te_channels[i] = te.channels[i]
            c.te_kernel_sizes[i] = te.kernel_sizes[i]
            c.te_dilations[i] = te.dilations[i]
        c.te_res2net_scale = te.res2net_scale
        c.te_se_channels = te.se_channels
        c.te_attention_channels = te.attention_channels
        c.te_out_dim = te.out_dim
        c.te_global_context = int(te.global_context)
        c.te_use_tanh = int(te.use_tanh)
    if unet is not None:  # UNET1D as RectifiedFlow.net (unet1d.py:255-268)
        n = len(unet.channels)
        if n > L.MAX_STAGES:
            raise ValueError("too many UNET1D levels")
        c.un_in_size = unet.in_size
        c.un_out_size = unet.out_size or 0
        c.un_n_levels = n
        for i, ch in enumerate(unet.channels):
            c.un_channels[i] = ch
        for i, r in enumerate(list(unet.ratios)[:n - 1]):
            c.un_ratios[i] = r
        c.un_kernel_size = unet.kernel_size
        c.un_time_channels = unet.time_channels
        c.un_time_cond_in_channels = unet.time_cond_in_channels
        c.un_time_cond_channels = unet.time_cond_channels
        c.un_cond_channels = unet.cond_channels
        c.un_n_attn_layers = unet.n_attn_layers
        c.un_use_res_last = int(unet.use_res_last)
        if model is None:  # the sampler-facing dimensions follow the net
            c.n_channels = unet.in_size
            c.cond_dim = unet.cond_channels
            c.tcond_dim = unet.time_cond_in_channels
            if seq_len is None:
                raise ValueError("seq_len must be given for a UNET1D engine")
    return c


def _ptr(t):
    return t.data_ptr() if t is not None else None


class Engine:
    """Owns one ``after_handle``.  Not re-entrant: one Engine per (device, caller thread)."""

    def __init__(self,
                 model: Optional[ModelConfig] = None,
                 autoencoder: Optional[AutoEncoderConfig] = None,
                 denoiser_state: Optional[Dict[str, torch.Tensor]] = None,
                 autoencoder_state: Optional[Dict[str, torch.Tensor]] = None,
                 structure_state: Optional[Dict[str, torch.Tensor]] = None,
                 timbre_state: Optional[Dict[str, torch.Tensor]] = None,
                 precision: str = "fp32",
                 device: int = 0,
                 max_batch: int = 8,
                 max_steps: int = 50,
                 seq_len: Optional[int] = None,
                 max_samples: int = 524288,
                 max_cache_size: int = 0,
                 unet=None,
                 unet_state: Optional[Dict[str, torch.Tensor]] = None,
                 drop_value: Optional[float] = None,
                 stream_slots: int = 0,
                 stream_max_frames: int = 0,
                 stream_gn_frames: int = 0,
                 latent_map_state: Optional[Dict[str, torch.Tensor]] = None):
        """``latent_map_state``: ``SmallAutoencoder.state_dict()`` of the export-time 2-D timbre projection
        (after/diffusion/latent_plot.py:20-37) behind ``latent_map`` / ``Streamer.latent2map`` / ``map2latent``; without it
        those are the identity the reference exports with ``--nolatent_project``.
        ``max_cache_size`` > 0 enables the streaming denoiser (what ``after_scripts/export.py:74-79`` binds to
        LOCAL_ATTENTION_SIZE): one rolling KV history per ``cache_index`` in [0, max_steps).
        ``unet`` / ``unet_state``: a ``config.UNetConfig`` + ``UNET1D.state_dict()`` make the conv denoiser
        (after/diffusion/networks/unet1d.py) the engine's ``net`` instead of DenoiserV2: ``sample`` / ``model_forward`` then
        run over it, and ``unet_forward`` is ``UNET1D.forward``.
        ``stream_slots`` > 0 adds that many independent streaming states to the codec and the structure encoder -- what the
        exported ``export_stream.ts`` / ``Encoder1D.forward_stream`` keep in cached convolutions and CachedGroupNorm
        (the exported Streamer holds two codec copies: slot 0 = structure, slot 1 = timbre); ``stream_max_frames`` = latent
        frames per streaming call the state is sized for (default 64), ``stream_gn_frames`` = CachedGroupNorm's padding in
        latent frames (default 64 = the 131072-sample first call of export_autoencoder.py)."""
        self._lib = L.load()
        self._h = C.c_void_p()
        if precision not in L.PRECISIONS:
            raise ValueError(f"precision must be one of {list(L.PRECISIONS)}")
        self.precision = precision
        self.device = torch.device("cuda", device)
        self.model_cfg = model
        self.ae_cfg = autoencoder
        self.cfg = _fill_config(model, autoencoder if autoencoder_state is not None else None, max_batch, max_steps,
                                seq_len, max_samples, structure_state is not None, timbre_state is not None,
                                max_cache_size, unet if unet_state is not None else None, stream_slots, stream_max_frames,
                                stream_gn_frames)
        if drop_value is not None:
            self.cfg.drop_value = float(drop_value)
        self.unet_cfg = unet
        L.check(self._lib.after_create(C.byref(self.cfg), device, C.byref(self._h)), None, "after_create")
        try:
            if denoiser_state is not None:
                self._load(L.MODULE_DENOISER, denoiser_state)
            if autoencoder_state is not None:
                self._load(L.MODULE_AUTOENCODER, autoencoder_state)
            if structure_state is not None:
                self._load(L.MODULE_STRUCTURE_ENCODER, structure_state)
            if timbre_state is not None:
                self._load(L.MODULE_TIMBRE_ENCODER, timbre_state)
            if unet_state is not None:
                if unet is None:
                    raise ValueError("unet_state needs a UNetConfig (unet=)")
                if denoiser_state is not None:
                    raise ValueError("an engine carries one `net`: pass denoiser_state or unet_state, not both")
                self._load(L.MODULE_UNET, unet_state)
            if latent_map_state is not None:
                self._load(L.MODULE_LATENT_MAP, latent_map_state)
            if any(s is not None for s in (denoiser_state, autoencoder_state, structure_state, timbre_state, unet_state)):
                L.check(self._lib.after_finalize_weights(self._h, L.PRECISIONS[precision]), self._h,
                        "after_finalize_weights")
        except Exception:
            self.close()
            raise
        self.has_denoiser = denoiser_state is not None
        self.has_unet = unet_state is not None
        self.has_codec = autoencoder_state is not None
        self.has_structure = structure_state is not None
        self.has_timbre = timbre_state is not None

    @classmethod
    def from_run(cls, folder: str, step: Optional[int] = None, codec_ts: Optional[str] = None,
                 autoencoder_state: Optional[Dict[str, torch.Tensor]] = None,
                 autoencoder: Optional[AutoEncoderConfig] = None, **kwargs) -> "Engine":
        """Engine from a reference training run folder (``config.gin`` + ``checkpoint<step>_EMA.pt``), the way
        ``after_scripts/export.py:52-101`` instantiates the model; the codec comes from an exported ``.ts``
        (``codec_ts``) or an ``AutoEncoder`` state dict.  ``kwargs`` go to the constructor (precision, max_batch, ...)."""
        from . import checkpoint as ck
        run = ck.load_run(folder, step)
        if codec_ts is not None:
            autoencoder_state = ck.codec_state_from_torchscript(codec_ts)
        if autoencoder_state is not None and autoencoder is None:
            autoencoder = ck.autoencoder_config_from_state(autoencoder_state)
        return cls(model=run["model"], autoencoder=autoencoder, denoiser_state=run["denoiser_state"],
                   autoencoder_state=autoencoder_state, structure_state=run["structure_state"],
                   timbre_state=run["timbre_state"], **kwargs)

    # ------------------------------------------------------------------ plumbing
    def _load(self, module: int, sd: Dict[str, torch.Tensor]):
        for key, t in sd.items():
            t = t.detach().cpu().contiguous()
            if t.dtype == torch.float32:
                dt = L.DTYPE_F32
            elif t.dtype == torch.float64:
                dt = L.DTYPE_F64
            elif t.dtype == torch.int64:
                dt = L.DTYPE_I64
            else:
                t = t.float()
                dt = L.DTYPE_F32
            shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
            L.check(
                self._lib.after_load_tensor(self._h, module, key.encode(), C.c_void_p(t.data_ptr()), shape, t.dim(), dt),
                self._h, f"after_load_tensor({key})")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.after_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def _dev(self, t: torch.Tensor, name: str) -> torch.Tensor:
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor")
        if t.device != self.device:
            raise RuntimeError(f"{name} is on {t.device}, this engine lives on {self.device} (no implicit copies, no CPU path)")
        if t.dtype != torch.float32:
            t = t.float()
        return t.contiguous()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def launch_count(self) -> int:
        return int(self._lib.after_launch_count(self._h))

    @property
    def device_bytes(self) -> int:
        return int(self._lib.after_device_bytes(self._h))

    @property
    def ae_ratio(self) -> int:
        return int(self._lib.after_ae_ratio(self._h))

    # ------------------------------------------------------------------ compute entry points
    def denoiser_forward(self, x, time, cond, time_cond, cache_index: Optional[int] = None):
        """``cache_index`` None: offline forward; an int: streaming forward against that KV history."""
        x = self._dev(x, "x")
        N, _, T = x.shape
        time = self._dev(time, "time")
        if time.dim() > 1:
            time = time.reshape(N, -1)[:, 0]  # only [..., 0] is read (transformerv2.py:524-528)
        time = time.contiguous()
        cond = self._dev(cond, "cond")
        time_cond = self._dev(time_cond, "time_cond")
        self._check_cond(cond, time_cond, N, T)
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            if cache_index is None:
                L.check(
                    self._lib.after_denoiser_forward(self._h, x.data_ptr(), time.data_ptr(), cond.data_ptr(),
                                                     time_cond.data_ptr(), out.data_ptr(), N, T, self._stream()), self._h,
                    "after_denoiser_forward")
            else:
                L.check(
                    self._lib.after_denoiser_forward_cached(self._h, x.data_ptr(), time.data_ptr(), cond.data_ptr(),
                                                            time_cond.data_ptr(), out.data_ptr(), N, T, int(cache_index),
                                                            self._stream()), self._h, "after_denoiser_forward_cached")
        return out

    @property
    def streaming(self) -> bool:
        return self.cfg.max_cache_size > 0

    def roll_cache(self, size: int, cache_index: int = 0):
        with torch.cuda.device(self.device):
            L.check(self._lib.after_roll_cache(self._h, int(size), int(cache_index), self._stream()), self._h,
                    "after_roll_cache")

    def reset_cache(self):
        with torch.cuda.device(self.device):
            L.check(self._lib.after_reset_cache(self._h, self._stream()), self._h, "after_reset_cache")

    def sample_stream(self, x_last, cond, time_cond, nb_steps, guidance_timbre=1.0, guidance_structure=1.0,
                      cfg_variant=L.CFG_AUDIO, clamp=0.1):
        """One audio block of the exported ``Streamer.sample`` (export.py:398-416): Euler step i uses and rolls KV cache i."""
        x_last = self._dev(x_last, "x_last")
        B, _, T = x_last.shape
        cond = self._dev(cond, "cond")
        time_cond = self._dev(time_cond, "time_cond")
        self._check_cond(cond, time_cond, B, T)
        out = torch.empty_like(x_last)
        with torch.cuda.device(self.device):
            L.check(
                self._lib.after_sample_stream(self._h, x_last.data_ptr(), cond.data_ptr(), time_cond.data_ptr(), out.data_ptr(),
                                              B, T, int(nb_steps), float(guidance_timbre), float(guidance_structure),
                                              int(cfg_variant), float(clamp), self._stream()), self._h, "after_sample_stream")
        return out

    def _check_cond(self, cond, time_cond, B, T):
        if self.has_unet:
            return self._check_cond_unet(cond, time_cond, B, T)
        if tuple(cond.shape) != (B, self.cfg.cond_dim):
            raise ValueError(f"cond must be ({B}, {self.cfg.cond_dim}), got {tuple(cond.shape)}")
        if tuple(time_cond.shape) != (B, self.cfg.tcond_dim, T):
            raise ValueError(f"time_cond must be ({B}, {self.cfg.tcond_dim}, {T}), got {tuple(time_cond.shape)}")

    def _check_cond_unet(self, cond, time_cond, B, T):
        u = self.unet_cfg
        if u.cond_channels and (cond is None or tuple(cond.shape) != (B, u.cond_channels)):
            raise ValueError(f"cond must be ({B}, {u.cond_channels})")
        if u.time_cond_in_channels and (time_cond is None or tuple(time_cond.shape) != (B, u.time_cond_in_channels, T)):
            raise ValueError(f"time_cond must be ({B}, {u.time_cond_in_channels}, {T})")

    def _conds(self, cond, time_cond):
        if self.has_unet:
            u = self.unet_cfg
            return self._opt(cond, "cond", u.cond_channels > 0), self._opt(time_cond, "time_cond", u.time_cond_in_channels > 0)
        return self._dev(cond, "cond"), self._dev(time_cond, "time_cond")

    def _opt(self, t, name, want: bool):
        """Device tensor or None (a UNET1D without that condition takes no tensor for it)."""
        if not want or t is None or t.numel() == 0:
            return None
        return self._dev(t, name)

    def unet_forward(self, x, time, cond=None, time_cond=None):
        """``UNET1D.forward(x, time=, time_cond=, cond=)`` (unet1d.py:376-429): (N, in_size, T) -> (N, out_size, T)."""
        if not self.has_unet:
            raise RuntimeError("this engine has no UNET1D weights")
        u = self.unet_cfg
        x = self._dev(x, "x")
        N, _, T = x.shape
        time = self._dev(time, "time").reshape(N, -1)[:, 0].contiguous()
        cond = self._opt(cond, "cond", u.cond_channels > 0)
        time_cond = self._opt(time_cond, "time_cond", u.time_cond_in_channels > 0)
        self._check_cond_unet(cond, time_cond, N, T)
        out = torch.empty(N, u.out_size or u.in_size, T, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            L.check(
                self._lib.after_unet_forward(self._h, x.data_ptr(), time.data_ptr(), cond.data_ptr() if cond is not None else None,
                                             time_cond.data_ptr() if time_cond is not None else None, out.data_ptr(), N, T,
                                             self._stream()), self._h, "after_unet_forward")
        return out

    def model_forward(self, x, time, cond, time_cond, guidance_timbre, guidance_structure, cfg_variant=L.CFG_AUDIO,
                      clamp=0.01, cache_index: Optional[int] = None):
        x = self._dev(x, "x")
        B, _, T = x.shape
        time = self._dev(time, "time").reshape(B, -1)[:, 0].contiguous()
        cond, time_cond = self._conds(cond, time_cond)
        self._check_cond(cond, time_cond, B, T)
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            if cache_index is None:
                L.check(
                    self._lib.after_model_forward(self._h, x.data_ptr(), time.data_ptr(), _ptr(cond), _ptr(time_cond),
                                                  out.data_ptr(), B, T, float(guidance_timbre), float(guidance_structure),
                                                  int(cfg_variant), float(clamp), self._stream()), self._h,
                    "after_model_forward")
            else:
                L.check(
                    self._lib.after_model_forward_cached(self._h, x.data_ptr(), time.data_ptr(), cond.data_ptr(),
                                                         time_cond.data_ptr(), out.data_ptr(), B, T, float(guidance_timbre),
                                                         float(guidance_structure), int(cfg_variant), float(clamp),
                                                         int(cache_index), self._stream()), self._h,
                    "after_model_forward_cached")
        return out

    def sample(self, x0, cond, time_cond, nb_steps, guidance_timbre=1.0, guidance_structure=1.0, cfg_variant=L.CFG_AUDIO,
               clamp=0.01):
        x0 = self._dev(x0, "x0")
        B, _, T = x0.shape
        cond, time_cond = self._conds(cond, time_cond)
        self._check_cond(cond, time_cond, B, T)
        out = torch.empty_like(x0)
        with torch.cuda.device(self.device):
            L.check(
                self._lib.after_sample(self._h, x0.data_ptr(), _ptr(cond), _ptr(time_cond), out.data_ptr(), B, T,
                                       int(nb_steps), float(guidance_timbre), float(guidance_structure), int(cfg_variant),
                                       float(clamp), self._stream()), self._h, "after_sample")
        return out

    def sample_host(self, x0, cond, time_cond, out, nb_steps, guidance_timbre=1.0, guidance_structure=1.0,
                    cfg_variant=L.CFG_AUDIO, clamp=0.01):
        """HOST tensors in, HOST tensor out (``out`` is filled and returned); H2D/D2H happen inside the call."""
        for name, t in (("x0", x0), ("cond", cond), ("time_cond", time_cond), ("out", out)):
            if t.device.type != "cpu" or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError(f"{name} must be a contiguous fp32 CPU tensor")
        B, _, T = x0.shape
        self._check_cond(cond, time_cond, B, T)
        with torch.cuda.device(self.device):
            L.check(
                self._lib.after_sample_host(self._h, x0.data_ptr(), cond.data_ptr(), time_cond.data_ptr(), out.data_ptr(), B, T,
                                            int(nb_steps), float(guidance_timbre), float(guidance_structure),
                                            int(cfg_variant), float(clamp), self._stream()), self._h, "after_sample_host")
        return out

    def ae_encode(self, audio):
        audio = self._dev(audio, "audio")
        if audio.dim() != 3 or audio.shape[1] != 1:
            raise ValueError("audio must be (B, 1, samples)")
        B, _, S = audio.shape
        r = self.ae_ratio
        if r == 0:
            raise RuntimeError("this engine has no codec")
        if S % r:
            raise ValueError(f"samples ({S}) must be a multiple of the codec ratio ({r})")
        z = torch.empty(B, self.cfg.ae_z_channels, S // r, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            L.check(self._lib.after_ae_encode(self._h, audio.data_ptr(), z.data_ptr(), B, S, self._stream()), self._h,
                    "after_ae_encode")
        return z

    def ae_decode(self, z):
        z = self._dev(z, "z")
        B, Cz, T = z.shape
        if Cz != self.cfg.ae_z_channels:
            raise ValueError(f"z must have {self.cfg.ae_z_channels} channels")
        audio = torch.empty(B, 1, T * self.ae_ratio, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            L.check(self._lib.after_ae_decode(self._h, z.data_ptr(), audio.data_ptr(), B, T, self._stream()), self._h,
                    "after_ae_decode")
        return audio

    # ---- streaming codec / structure encoder (engine created with stream_slots > 0) --------------------------------
    @property
    def stream_slots(self) -> int:
        return int(self.cfg.stream_slots)

    def ae_encode_stream(self, slot: int, audio):
        """One buffer through the streaming export's ``encode`` (export_autoencoder.py AE_notcausal): (B,1,S) -> (B,Z,S/ratio)."""
        audio = self._dev(audio, "audio")
        B, _, S = audio.shape
        r = self.ae_ratio
        if r == 0 or S % r:
            raise ValueError(f"samples ({S}) must be a multiple of the codec ratio ({r})")
        z = torch.empty(B, self.cfg.ae_z_channels, S // r, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            L.check(self._lib.after_ae_encode_stream(self._h, int(slot), audio.data_ptr(), z.data_ptr(), B, S, self._stream()),
                    self._h, "after_ae_encode_stream")
        return z

    def ae_decode_stream(self, slot: int, z):
        """One buffer through the streaming export's overlap-add ``decode`` (export_autoencoder.py:128-153)."""
        z = self._dev(z, "z")
        B, Cz, T = z.shape
        if Cz != self.cfg.ae_z_channels:
            raise ValueError(f"z must have {self.cfg.ae_z_channels} channels")
        audio = torch.empty(B, 1, T * self.ae_ratio, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            L.check(self._lib.after_ae_decode_stream(self._h, int(slot), z.data_ptr(), audio.data_ptr(), B, T, self._stream()),
                    self._h, "after_ae_decode_stream")
        return audio

    def structure_encode_stream(self, slot: int, z):
        """``Encoder1D.forward_stream`` (encoder.py:300-322) with cached convolutions."""
        z = self._dev(z, "z")
        B, Cin, T = z.shape
        out = torch.empty(B, self.cfg.se_channels[self.cfg.se_n_blocks - 1], T, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            L.check(self._lib.after_structure_encode_stream(self._h, int(slot), z.data_ptr(), out.data_ptr(), B, T, self._stream()),
                    self._h, "after_structure_encode_stream")
        return out

    def stream_reset(self, slot: int):
        with torch.cuda.device(self.device):
            L.check(self._lib.after_stream_reset(self._h, int(slot), self._stream()), self._h, "after_stream_reset")

    def structure_encode(self, z):
        z = self._dev(z, "z")
        B, Cin, T = z.shape
        out = torch.empty(B, self.cfg.se_channels[self.cfg.se_n_blocks - 1], T, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            L.check(self._lib.after_structure_encode(self._h, z.data_ptr(), out.data_ptr(), B, T, self._stream()), self._h,
                    "after_structure_encode")
        return out

    def timbre_encode(self, z):
        z = self._dev(z, "z")
        B, Cin, T = z.shape
        if Cin != self.cfg.te_in_size:
            raise ValueError(f"z must have {self.cfg.te_in_size} channels")
        out = torch.empty(B, self.cfg.te_out_dim, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            L.check(self._lib.after_timbre_encode(self._h, z.data_ptr(), out.data_ptr(), B, T, self._stream()), self._h,
                    "after_timbre_encode")
        return out

    def latent_map(self, x, direction: int):
        """``Streamer.latent2map`` (direction 0) / ``map2latent`` (direction 1), after_scripts/export.py:494-508:
        (B, C_in, T) -> (B, C_out, T): time average -> projection MLP (or identity) -> repeated over T."""
        x = self._dev(x, "x")
        B, Cin, T = x.shape
        out = torch.empty(B, 64, T, device=self.device, dtype=torch.float32)
        c_out = C.c_int(0)
        with torch.cuda.device(self.device):
            L.check(self._lib.after_latent_map(self._h, int(direction), x.data_ptr(), out.data_ptr(), B, Cin, T, C.byref(c_out),
                                               self._stream()), self._h, "after_latent_map")
        return out.view(-1)[:B * c_out.value * T].view(B, c_out.value, T)

    def generate(self, audio_structure, audio_timbre, x0, nb_steps, guidance_timbre=1.0, guidance_structure=1.0):
        """Whole audio-to-audio chain on device tensors: (B,1,S) x2 + prior noise (B,C,S/ratio) -> audio (B,1,S)."""
        a_s, a_t, x0 = self._dev(audio_structure, "audio_structure"), self._dev(audio_timbre, "audio_timbre"), self._dev(x0, "x0")
        B, _, S = a_s.shape
        if a_t.shape != a_s.shape or tuple(x0.shape) != (B, self.cfg.n_channels, S // max(self.ae_ratio, 1)):
            raise ValueError("audio_structure / audio_timbre / x0 shapes disagree")
        out = torch.empty_like(a_s)
        with torch.cuda.device(self.device):
            L.check(self._lib.after_generate(self._h, a_s.data_ptr(), a_t.data_ptr(), x0.data_ptr(), out.data_ptr(), B, S,
                                             int(nb_steps), float(guidance_timbre), float(guidance_structure), self._stream()),
                    self._h, "after_generate")
        return out

    def generate_host(self, audio_structure, audio_timbre, x0, out, nb_steps, guidance_timbre=1.0, guidance_structure=1.0):
        """Same with HOST tensors (pinned for speed); ``out`` (B,1,S) is filled and returned."""
        for name, t in (("audio_structure", audio_structure), ("audio_timbre", audio_timbre), ("x0", x0), ("out", out)):
            if t.device.type != "cpu" or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError(f"{name} must be a contiguous fp32 CPU tensor")
        B, _, S = audio_structure.shape
        with torch.cuda.device(self.device):
            L.check(self._lib.after_generate_host(self._h, audio_structure.data_ptr(), audio_timbre.data_ptr(), x0.data_ptr(),
                                                  out.data_ptr(), B, S, int(nb_steps), float(guidance_timbre),
                                                  float(guidance_structure), self._stream()), self._h, "after_generate_host")
        return out

    def profile(self, on: bool):
        L.check(self._lib.after_profile_enable(self._h, int(on)), self._h, "after_profile_enable")

    def profile_read(self):
        """{class: {launches, ms, flops, bytes}} accumulated since ``profile(True)``."""
        out = {}
        for name, k in L.KERNEL_CLASSES.items():
            n, ms, fl, by = C.c_int64(), C.c_double(), C.c_double(), C.c_double()
            L.check(self._lib.after_profile_read(self._h, k, C.byref(n), C.byref(ms), C.byref(fl), C.byref(by)), self._h,
                    "after_profile_read")
            out[name] = {"launches": n.value, "ms": ms.value, "flops": fl.value, "bytes": by.value}
        return out

    def debug_gemm(self, A, W, bias=None, precision="fp32"):
        A = self._dev(A, "A")
        W = self._dev(W, "W")
        M, K = A.shape
        N = W.shape[0]
        out = torch.empty(M, N, device=self.device, dtype=torch.float32)
        b = self._dev(bias, "bias").data_ptr() if bias is not None else None
        with torch.cuda.device(self.device):
            L.check(
                self._lib.after_debug_gemm(self._h, A.data_ptr(), W.data_ptr(), b, out.data_ptr(), M, N, K,
                                           L.PRECISIONS[precision], self._stream()), self._h, "after_debug_gemm")
        return out
