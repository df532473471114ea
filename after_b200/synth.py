"""Synthetic (seeded) weights in the reference's ``state_dict`` key layout.

There is no network access to fetch AFTER checkpoints, so tests / smoke / bench use random
weights.  The dictionaries built here use exactly the key names and shapes a reference
checkpoint has (verified by ``tests/golden/make_golden.py`` which
``load_state_dict(strict=True)``-s them into the reference modules), so the same loading path
serves real checkpoints.  Every parameter that defaults to an identity in PyTorch (norm
affines, BatchNorm running statistics, Snake alpha/beta, weight-norm gains) is randomised so
that a wrong fold shows up as a parity failure.

Also holds the host-side design of the PQMF prototype filter (pqmf.py:58-92, 186-279 in the
reference): real checkpoints carry the filters as buffers, synthetic ones need to design them.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from .config import (AutoEncoderConfig, DenoiserConfig, EcapaConfig, Encoder1DConfig)

StateDict = Dict[str, torch.Tensor]


class _Rng:

    def __init__(self, seed: int):
        self.g = torch.Generator().manual_seed(seed)

    def uniform(self, shape, lo, hi):
        return torch.rand(shape, generator=self.g) * (hi - lo) + lo

    def normal(self, shape, mean=0.0, std=1.0):
        return torch.randn(shape, generator=self.g) * std + mean

    def linear(self, out_f, in_f, bias=True):
        bound = 1.0 / math.sqrt(in_f)
        w = self.uniform((out_f, in_f), -bound, bound)
        b = self.uniform((out_f, ), -bound, bound) if bias else None
        return w, b


# ----------------------------------------------------------------------------- denoiser
def denoiser_state_dict(cfg: DenoiserConfig, seed: int = 0) -> StateDict:
    r = _Rng(seed)
    D = cfg.embed_dim
    sd: StateDict = {}

    def put_linear(prefix, out_f, in_f, bias=True):
        w, b = r.linear(out_f, in_f, bias)
        sd[prefix + ".weight"] = w
        if bias:
            sd[prefix + ".bias"] = b

    put_linear("embedding.0", D, cfg.cond_dim + cfg.noise_embed_dims)
    put_linear("embedding.2", D, D)
    tb = "denoiser_trans_block."
    sd[tb + "precomputed_pos_enc"] = torch.arange(cfg.seq_len)
    put_linear(tb + "patchify_and_embed.1", D, cfg.n_channels)
    put_linear(tb + "patchify_and_embed_tcond.1", cfg.tcond_dim, cfg.tcond_dim)
    freqs = 1.0 / (cfg.rotary_theta**(torch.arange(0, cfg.rotary_dim, 2).float() / cfg.rotary_dim))
    sd[tb + "rotary_emb.freqs"] = freqs.clone()
    for i in range(cfg.n_layers):
        p = f"{tb}decoder_blocks.{i}."
        put_linear(p + "self_attention.qkv_linear", 3 * D, D, bias=False)
        sd[p + "self_attention.mha.rotary_emb.freqs"] = freqs.clone()
        sd[p + "self_attention.rotary_emb.freqs"] = freqs.clone()
        put_linear(p + "mlp.mlp.0", cfg.mlp_multiplier * D, D)
        put_linear(p + "mlp.mlp.2", D, cfg.mlp_multiplier * D)
        for n in ("norm1", "norm3"):
            sd[p + n + ".weight"] = r.normal((D, ), 1.0, 0.1)
            sd[p + n + ".bias"] = r.normal((D, ), 0.0, 0.1)
        put_linear(p + "linear", 2 * D, D)
        put_linear(p + "tcond_linear", 2 * D, cfg.tcond_dim)
    put_linear(tb + "out_proj.0", cfg.n_channels, D)
    return sd


# ----------------------------------------------------------------------------- UNET1D (oracle-only so far)
def unet_state_dict(cfg, seed: int = 0) -> StateDict:
    """Synthetic ``UNET1D.state_dict()`` in the reference key layout (unet1d.py:30-118, 122-253, 255-375)."""
    r = _Rng(seed)
    sd: StateDict = {}
    k = cfg.kernel_size
    n = len(cfg.channels)
    ratios = [1] + list(cfg.ratios)
    tcc, tci = cfg.time_cond_channels, cfg.time_cond_in_channels
    out_size = cfg.in_size if cfg.out_size is None else cfg.out_size

    def conv(prefix, co, ci, ks):
        bound = 1.0 / math.sqrt(ci * ks)
        sd[prefix + ".weight"] = r.uniform((co, ci, ks), -bound, bound)
        sd[prefix + ".bias"] = r.uniform((co, ), -bound, bound)

    def lin(prefix, o, i):
        w, b = r.linear(o, i)
        sd[prefix + ".weight"], sd[prefix + ".bias"] = w, b

    def gn(prefix, c):
        sd[prefix + ".weight"] = r.normal((c, ), 1.0, 0.1)
        sd[prefix + ".bias"] = r.normal((c, ), 0.0, 0.1)

    def conv_block(prefix, in_c, out_c, skip_c):
        cin = in_c + skip_c + tcc
        conv(prefix + ".conv1", out_c, cin, k)
        gn(prefix + ".gn1", cin)
        conv(prefix + ".conv2", out_c, out_c, k)
        gn(prefix + ".gn2", out_c)
        lin(prefix + ".time_mlp.0", 128, cfg.time_channels)
        lin(prefix + ".time_mlp.2", 2 * out_c, 128)
        if cfg.cond_channels > 0:
            lin(prefix + ".cond_mlp.0", 128, cfg.cond_channels)
            lin(prefix + ".cond_mlp.2", 2 * out_c, 128)
        if skip_c:
            conv(prefix + ".to_out", out_c, in_c, 1)

    def attn(prefix, c):
        gn(prefix + ".norm", c)
        conv(prefix + ".qkv_proj", 3 * c, c, 1)
        conv(prefix + ".out_proj", c, c, 1)

    if tcc:
        conv("cond_emb_time.0.0", tcc, tci, k)
        for i in range(n):
            conv(f"cond_emb_time.{i + 1}.0", tcc, tcc, k)
    in0 = cfg.in_size + (tci if not tcc else 0)
    ins = [in0] + list(cfg.channels[:-1])
    for i in range(n):
        p = f"down_layers.{i}"
        conv_block(p + ".conv", ins[i], ins[i], 0)
        if i >= 1 and i >= n - cfg.n_attn_layers:
            attn(p + ".self_attn", ins[i])
        conv(p + ".pool", cfg.channels[i], ins[i], k)
    conv_block("middle_block.conv", cfg.channels[-1], cfg.channels[-1], 0)
    if cfg.n_attn_layers > 0:
        attn("middle_block.self_attn", cfg.channels[-1])
    for i in range(1, n + 1):
        p = f"up_layers.{i - 1}"
        last = i == n
        ic = cfg.channels[n - i]
        oc = out_size if last else cfg.channels[n - i - 1]
        ratio = ratios[n - i]
        if ratio == 1:
            if ic != oc:
                conv(p + ".up", oc, ic, 3)
        else:
            conv(p + ".up.1", oc, ic, 3)
        conv_block(p + ".conv", oc, oc, in0 if last else oc)
        if not last and i <= cfg.n_attn_layers:
            attn(p + ".self_attn", oc)
    return sd


# ----------------------------------------------------------------------------- PQMF design
def _kaiser_lowpass(wc: float, atten: float, n_taps=None):
    from scipy.signal import firwin, kaiserord
    n_est, beta = kaiserord(atten, wc / np.pi)
    n_est = 2 * (n_est // 2) + 1
    n = n_taps if n_taps is not None else n_est
    return firwin(n, wc, window=("kaiser", beta), scale=False, fs=2 * np.pi)


def pqmf_prototype(atten: float, n_band: int) -> np.ndarray:
    """Near-perfect-reconstruction prototype low-pass (cut-off found with Nelder-Mead so that the
    autocorrelation of the filter vanishes at multiples of 2M; reference pqmf.py:58-92)."""
    from scipy.optimize import fmin

    def objective(w):
        w = float(np.atleast_1d(w)[0])
        h = _kaiser_lowpass(w, atten)
        g = np.convolve(h, h[::-1], "full")
        g = np.abs(g[g.shape[-1] // 2::2 * n_band][1:])
        return np.max(g)

    wc = fmin(objective, 1.0 / n_band, disp=0)[0]
    return _kaiser_lowpass(wc, atten)


def pqmf_filters(atten: float, n_band: int) -> StateDict:
    """``pqmf.*`` buffers of ``CachedPQMF`` (pqmf.py:186-279)."""
    h = torch.from_numpy(pqmf_prototype(atten, n_band)).float()
    n = h.shape[-1]
    k = torch.arange(n_band).reshape(-1, 1)
    t = torch.arange(-(n // 2), n // 2 + 1)
    phase = (-1.0)**k * math.pi / 4
    hk = 2 * h * torch.cos((2 * k + 1) * math.pi / (2 * n_band) * t + phase)
    # centre-pad to the next power of two
    n2 = 2**math.ceil(math.log2(hk.shape[-1]))
    pad = n2 - hk.shape[-1]
    hk = torch.nn.functional.pad(hk, (pad // 2, pad // 2 + pad % 2))
    fwd = torch.nn.functional.pad(hk, (0, 1)).unsqueeze(1)  # (M, 1, n2+1)
    hki = hk.flip(-1).reshape(n_band, n2 // n_band, n_band).permute(2, 0, 1)  # "c (t m) -> m c t"
    inv = torch.nn.functional.pad(hki, (0, 1)).contiguous()  # (M, M, n2/M+1)
    return {
        "pqmf.hk": hk.contiguous(),
        "pqmf.h": h,
        "pqmf.forward_conv.weight": fwd.contiguous(),
        "pqmf.inverse_conv.weight": inv,
    }


# ----------------------------------------------------------------------------- codec
def _wn_conv(sd: StateDict, r: _Rng, prefix: str, out_c: int, in_c: int, k: int, transposed=False):
    """weight-normed conv: ``weight_g`` / ``weight_v`` (+ bias).  For ConvTranspose1d the stored
    tensor is (in, out, k) and the norm runs over dims != 0, i.e. per *input* channel."""
    fan_in = in_c * k
    bound = 1.0 / math.sqrt(fan_in)
    shape = (in_c, out_c, k) if transposed else (out_c, in_c, k)
    v = r.uniform(shape, -bound, bound)
    g = v.flatten(1).norm(dim=1).reshape(-1, 1, 1) * r.uniform((shape[0], 1, 1), 0.8, 1.2)
    sd[prefix + ".weight_v"] = v
    sd[prefix + ".weight_g"] = g
    sd[prefix + ".bias"] = r.uniform((out_c, ), -bound, bound)


def _conv_block(sd, r, prefix, in_c, out_c, k):
    """``ConvBlock1d``: CachedGroupNorm -> SnakeBeta -> weight-normed conv."""
    sd[prefix + ".net.0.pad"] = torch.zeros(4, in_c, 1)
    sd[prefix + ".net.0.gn.weight"] = r.normal((in_c, ), 1.0, 0.1)
    sd[prefix + ".net.0.gn.bias"] = r.normal((in_c, ), 0.0, 0.1)
    sd[prefix + ".net.1.alpha"] = r.uniform((in_c, ), 0.5, 1.5)
    sd[prefix + ".net.1.beta"] = r.uniform((in_c, ), 0.5, 1.5)
    _wn_conv(sd, r, prefix + ".net.2", out_c, in_c, k)


def _resnet(sd, r, prefix, in_c, out_c, k):
    _conv_block(sd, r, prefix + ".net.branches.0.0", in_c, out_c, k)
    _conv_block(sd, r, prefix + ".net.branches.0.1", out_c, out_c, 1)
    if in_c != out_c:
        _wn_conv(sd, r, prefix + ".net.branches.1", out_c, in_c, 1)


def _snake(sd, r, prefix, c):
    sd[prefix + ".alpha"] = r.uniform((c, ), 0.5, 1.5)
    sd[prefix + ".beta"] = r.uniform((c, ), 0.5, 1.5)


def autoencoder_state_dict(cfg: AutoEncoderConfig, seed: int = 0) -> StateDict:
    r = _Rng(seed)
    sd: StateDict = {}
    sd.update(pqmf_filters(cfg.pqmf_attenuation, cfg.pqmf_bands))
    ks = cfg.kernel_size
    # encoder (SimpleNetsStream.py:400-459)
    ch = [cfg.channels * m for m in cfg.multipliers]
    _resnet(sd, r, "encoder.net.0", cfg.in_channels, ch[0], ks)
    n_stage = len(cfg.factors)
    for i in range(n_stage):
        p = f"encoder.net.{i + 1}"
        for j in range(cfg.num_blocks):
            _resnet(sd, r, f"{p}.net.{j}", ch[i], ch[i], ks)
        _snake(sd, r, f"{p}.net.{cfg.num_blocks}", ch[i])
        f = cfg.factors[i]
        _wn_conv(sd, r, f"{p}.net.{cfg.num_blocks + 1}", ch[i + 1], ch[i], 2 * f)
    _snake(sd, r, f"encoder.net.{n_stage + 1}", ch[-1])
    _wn_conv(sd, r, f"encoder.net.{n_stage + 2}", cfg.z_channels, ch[-1], 3)
    # decoder (SimpleNetsStream.py:552-651)
    dch = [cfg.channels * m for m in cfg.decoder_multipliers]
    dfac = cfg.factors[::-1]
    _wn_conv(sd, r, "decoder.net.0", dch[0], cfg.z_channels, ks)
    for i in range(n_stage):
        p = f"decoder.net.{i + 1}"
        _snake(sd, r, f"{p}.net.0", dch[i])
        _wn_conv(sd, r, f"{p}.net.1", dch[i + 1], dch[i], 2 * dfac[i], transposed=True)
        for j in range(cfg.num_blocks):
            _resnet(sd, r, f"{p}.net.{2 + j}", dch[i + 1], dch[i + 1], ks)
    out_c = cfg.in_channels * (2 if cfg.use_loudness else 1)
    _conv_block(sd, r, "decoder.synth.branches.0.net.0", dch[-1], out_c, ks)
    _conv_block(sd, r, "decoder.synth.branches.0.net.1", out_c, out_c, 1)
    return sd


# ----------------------------------------------------------------------------- Encoder1D
def _batchnorm(sd, r, prefix, c):
    sd[prefix + ".weight"] = r.normal((c, ), 1.0, 0.1)
    sd[prefix + ".bias"] = r.normal((c, ), 0.0, 0.1)
    sd[prefix + ".running_mean"] = r.uniform((c, ), -0.5, 0.5)
    sd[prefix + ".running_var"] = r.uniform((c, ), 0.5, 1.5)
    sd[prefix + ".num_batches_tracked"] = torch.tensor(0)


def _v2_conv_block(sd, r, prefix, in_c, out_c, k):
    """``V2ConvBlock1D`` (encoder.py:25-71).  ``gn1``/``gn2`` are the same modules as
    ``net.branches.0.{0,3}`` (registered twice), so both key families carry the same tensors."""
    _batchnorm(sd, r, prefix + ".gn1", in_c)
    _batchnorm(sd, r, prefix + ".gn2", out_c)
    for name in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
        sd[f"{prefix}.net.branches.0.0.{name}"] = sd[f"{prefix}.gn1.{name}"]
        sd[f"{prefix}.net.branches.0.3.{name}"] = sd[f"{prefix}.gn2.{name}"]
    _wn_conv(sd, r, prefix + ".net.branches.0.2", out_c, in_c, k)
    _wn_conv(sd, r, prefix + ".net.branches.0.6", out_c, out_c, k)


def encoder1d_state_dict(cfg: Encoder1DConfig, seed: int = 0) -> StateDict:
    r = _Rng(seed)
    sd: StateDict = {}
    ratios = [1] + list(cfg.ratios)
    ins = [cfg.in_size] + list(cfg.channels[:-1])
    for i, (cin, cout, ratio) in enumerate(zip(ins, cfg.channels, ratios)):
        _v2_conv_block(sd, r, f"net.{i}.net.0", cin, cin, cfg.kernel_size)
        kpool = 1 if ratio == 1 else 2 * ratio
        _wn_conv(sd, r, f"net.{i}.net.1", cout, cin, kpool)
    _v2_conv_block(sd, r, f"net.{len(cfg.channels)}", cfg.channels[-1], cfg.channels[-1],
                   cfg.kernel_size)
    return sd


# ----------------------------------------------------------------------------- ECAPA-TDNN
def _tdnn(sd, r, prefix, in_c, out_c, k):
    """``TDNNBlock``: plain (not weight-normed) conv -> ReLU -> BatchNorm1d (ecapa_encoder.py:85-138)."""
    bound = 1.0 / math.sqrt(in_c * k)
    sd[prefix + ".conv.conv.weight"] = r.uniform((out_c, in_c, k), -bound, bound)
    sd[prefix + ".conv.conv.bias"] = r.uniform((out_c, ), -bound, bound)
    _batchnorm(sd, r, prefix + ".norm", out_c)


def _plain_conv(sd, r, prefix, in_c, out_c, k=1):
    bound = 1.0 / math.sqrt(in_c * k)
    sd[prefix + ".weight"] = r.uniform((out_c, in_c, k), -bound, bound)
    sd[prefix + ".bias"] = r.uniform((out_c, ), -bound, bound)


def ecapa_state_dict(cfg: EcapaConfig, seed: int = 0) -> StateDict:
    """``ECAPATDNN`` (ecapa_encoder.py:458-566) with pooling, global context, groups == 1."""
    r = _Rng(seed)
    sd: StateDict = {}
    ch, ks = cfg.channels, cfg.kernel_sizes
    _tdnn(sd, r, "blocks.0", cfg.in_size, ch[0], ks[0])
    for i in range(1, len(ch) - 1):
        p = f"blocks.{i}"
        _tdnn(sd, r, p + ".tdnn1", ch[i - 1], ch[i], 1)
        sub = ch[i] // cfg.res2net_scale
        for j in range(cfg.res2net_scale - 1):
            _tdnn(sd, r, f"{p}.res2net_block.blocks.{j}", sub, sub, ks[i])
        _tdnn(sd, r, p + ".tdnn2", ch[i], ch[i], 1)
        _plain_conv(sd, r, p + ".se_block.conv1.conv", ch[i], cfg.se_channels)
        _plain_conv(sd, r, p + ".se_block.conv2.conv", cfg.se_channels, ch[i])
        if ch[i - 1] != ch[i]:
            _plain_conv(sd, r, p + ".shortcut.conv", ch[i - 1], ch[i])
    _tdnn(sd, r, "mfa", ch[-1], ch[-1], ks[-1])
    _tdnn(sd, r, "asp.tdnn", ch[-1] * (3 if cfg.global_context else 1), cfg.attention_channels, 1)
    _plain_conv(sd, r, "asp.conv.conv", cfg.attention_channels, ch[-1])
    _batchnorm(sd, r, "asp_bn", 2 * ch[-1])
    _plain_conv(sd, r, "fc.conv", 2 * ch[-1], cfg.out_dim)
    return sd


# ----------------------------------------------------------------------------- inputs
def synth_inputs_per_stream(batch: int, cfg: DenoiserConfig, seed: int = 1234, frames: int = None, first: int = 0):
    """Like ``synth_inputs`` but stream ``b`` is drawn from its own generator (seed + first + b): stream b's inputs do not
    depend on how many streams are generated with it, so a 1-GPU run (8 streams) and an 8-GPU run (64 streams) agree on the
    streams they share -- what lets ``bench.py`` publish a checksum that must not change with the GPU count."""
    parts = [synth_inputs(1, cfg, seed=seed + first + b, frames=frames) for b in range(batch)]
    return tuple(torch.cat([p[i] for p in parts]) for i in range(3))


def synth_inputs(batch: int, cfg: DenoiserConfig, seed: int = 1234, frames: int = None):
    """x0 / cond / time_cond the way SURVEY.md section 8d prescribes (host-generated, so that
    1-GPU and N-GPU runs see identical streams)."""
    T = cfg.seq_len if frames is None else frames
    g = torch.Generator().manual_seed(seed)
    x0 = torch.randn(batch, cfg.n_channels, T, generator=g)
    cond = torch.randn(batch, cfg.cond_dim, generator=g)
    if cfg.tcond_dim == 128:  # piano roll: <= 8 voices per frame, velocity/127 in (0, 1]
        tc = torch.zeros(batch, 128, T)
        for b in range(batch):
            pitches = torch.randint(21, 109, (8, ), generator=g)
            for p in pitches:
                on = torch.rand(T, generator=g) < 0.5
                vel = torch.randint(1, 128, (1, ), generator=g).float() / 127.0
                tc[b, p, on] = vel
    else:
        tc = torch.randn(batch, cfg.tcond_dim, T, generator=g)
    return x0, cond, tc


def synth_audio(batch: int, samples: int, seed: int = 7):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, 1, samples, generator=g) * 2 - 1
