"""CPU oracle of the STREAMING codec / structure-encoder path: what the exported models compute buffer by buffer.

TEST INFRASTRUCTURE ONLY (same rules as ``after_oracle.py``): nothing under ``after_b200/`` imports this module.

What is restated, and from where:

* ``after_scripts/export_autoencoder.py:16-153, 305-319`` -- the non-causal export ``AE_notcausal`` that a baseAE run
  produces as ``export_stream.ts``: ENCODER built under ``cc.use_cached_conv(True)``, DECODER built offline and run over
  ``[z_buffer ; z]`` with a linear cross-fade of ``n_fade`` latent frames, ``CachedGroupNorm.stream = True`` in both,
  and -- because only ``model.encoder`` is swapped (:311-312) -- an OFFLINE (per-buffer, zero-padded) PQMF in front of
  the cached encoder.
* ``after/autoencoder/networks/SimpleNetsStream.py:95-147`` -- ``CachedGroupNorm`` stream branch: statistics over
  ``[pad ; x]`` where ``pad`` holds the previous ``padding_size`` frames (zeros at first), ``padding_size`` = the layer's
  length at the export script's first call (131072 samples / 64 latent frames: ``gn_latent_frames``).
* ``after/diffusion/networks/encoder.py:300-322`` + ``after_scripts/export.py:14-17, 431-435`` -- ``Encoder1D.forward_stream``
  with cached causal convolutions.
* **cached_conv** (acids-ircam/cached_conv >= 2.5.0, ``requirements.txt:12``) is an un-vendored dependency, absent here:
  its published streaming algorithm is restated below (``CachedPadding1d``, ``CachedConv1d`` with its stride delay,
  ``AlignBranches``, the cumulative-delay bookkeeping of ``CachedSequential``).  PARITY UNPINNED against the package
  itself; the restatement is anchored on properties checked in ``tests/test_oracle_stream.py`` against the offline
  oracle (which IS pinned to reference fixtures): with causal padding the streamed output equals the offline output
  exactly; without GroupNorm (``use_norm = False``) the streamed centred-padding encoder equals the offline encoder
  delayed by the cumulative delay the reference's own constructors compute.

State lives in a plain dict (one per exported model copy), keyed by the reference module path.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import after_oracle as O

Tensor = torch.Tensor
State = Dict[str, Tensor]


# ----------------------------------------------------------------------------- cached_conv primitives
def cached_pad(st: State, key: str, x: Tensor, pad: int, crop: bool = False) -> Tensor:
    """``CachedPadding1d(pad, crop)``: prepend the cached last ``pad`` frames (zeros at first), remember the new last
    ``pad`` frames, optionally drop the last ``pad`` frames (a pure delay)."""
    if pad == 0:
        return x
    buf = st.get(key)
    if buf is None or buf.shape[0] != x.shape[0]:
        buf = x.new_zeros(x.shape[0], x.shape[1], pad)
    y = torch.cat([buf, x], -1)
    st[key] = y[..., -pad:].clone()
    return y[..., :-pad] if crop else y


def stride_delay(r_pad: int, cd: int, stride: int) -> int:
    """``CachedConv1d.__init__``: frames of extra delay that align a strided conv with the delay accumulated so far."""
    return (stride - ((r_pad + cd) % stride)) % stride


def cached_conv(st: State, key: str, x: Tensor, w: Tensor, b: Optional[Tensor], pad, stride: int = 1, dilation: int = 1,
                cd: int = 0):
    """``CachedConv1d.forward``: delay by the stride delay, prepend the cache of ``l + r`` frames, un-padded conv.
    Returns (y, cumulative delay after this layer)."""
    l, r = pad
    sdly = stride_delay(r, cd, stride)
    x = cached_pad(st, key + ".downsampling_delay", x, sdly, crop=True)
    x = cached_pad(st, key + ".cache", x, l + r)
    return F.conv1d(x, w, b, stride=stride, dilation=dilation), (r + sdly + cd) // stride


def stream_group_norm(st: State, key: str, x: Tensor, groups: int, w: Tensor, b: Tensor, pad_frames: int) -> Tensor:
    """``CachedGroupNorm.forward`` with ``stream = True`` (SimpleNetsStream.py:134-144)."""
    t = x.shape[-1]
    buf = st.get(key + ".pad")
    if buf is None or buf.shape[0] != x.shape[0]:
        buf = x.new_zeros(x.shape[0], x.shape[1], pad_frames)
    y = torch.cat([buf, x], -1)
    st[key + ".pad"] = y[..., -pad_frames:].clone()
    return F.group_norm(y, groups, w, b, eps=1e-5)[..., -t:]


# ----------------------------------------------------------------------------- streaming codec encoder
def _conv_block_stream(sd, st, prefix, x, dilation, groups, scale, gn_latent_frames, cd=0):
    C = x.shape[1]
    y = x
    if (prefix + ".net.0.gn.weight") in sd:
        y = stream_group_norm(st, prefix + ".net.0", x, min(C, groups), sd[prefix + ".net.0.gn.weight"],
                              sd[prefix + ".net.0.gn.bias"], gn_latent_frames * scale)
    y = O.snake_beta(y, sd[prefix + ".net.1.alpha"], sd[prefix + ".net.1.beta"])
    w = O.fold_weight_norm(sd, prefix + ".net.2")
    return cached_conv(st, prefix + ".net.2", y, w, sd[prefix + ".net.2.bias"], O.same_padding(w.shape[-1], dilation), 1,
                       dilation, cd)


def _resnet_stream(sd, st, prefix, x, dilation, groups, scale, gnf, cd):
    """ResnetBlock1d under cached_conv (SimpleNetsStream.py:197-254): AlignBranches(net, to_out, delays=[d, 0]) delays the
    skip branch by the delay d of block1's conv; the block adds d to the cumulative delay."""
    y, d = _conv_block_stream(sd, st, prefix + ".net.branches.0.0", x, dilation, groups, scale, gnf)  # block convs get cd = 0
    y, _ = _conv_block_stream(sd, st, prefix + ".net.branches.0.1", y, 1, 8, scale, gnf)
    xd = cached_pad(st, prefix + ".net.paddings.1", x, d, crop=True)
    if (prefix + ".net.branches.1.weight_v") in sd:
        w = O.fold_weight_norm(sd, prefix + ".net.branches.1")
        xd, _ = cached_conv(st, prefix + ".net.branches.1", xd, w, sd[prefix + ".net.branches.1.bias"], (0, 0))
    return y + xd, cd + d


def ae_encode_stream(sd, cfg, st: State, audio: Tensor, gn_latent_frames: int = 64) -> Tensor:
    """One buffer through ``AE_notcausal.encode`` of the streaming export: offline PQMF on the buffer, cached encoder."""
    sd = O._cast(sd, audio.dtype)
    x = O.pqmf_analysis(sd, audio) if cfg.pqmf_bands > 1 else audio
    scale = cfg.ratio // max(cfg.pqmf_bands, 1)  # frames of this layer per latent frame
    x, cd = _resnet_stream(sd, st, "encoder.net.0", x, 1, cfg.resnet_groups, scale, gn_latent_frames, 0)
    nb = cfg.num_blocks
    for i, f in enumerate(cfg.factors):
        p = f"encoder.net.{i + 1}"
        for j in range(nb):
            x, cd = _resnet_stream(sd, st, f"{p}.net.{j}", x, cfg.dilations[j], cfg.resnet_groups, scale, gn_latent_frames, cd)
        x = O.snake_beta(x, sd[f"{p}.net.{nb}.alpha"], sd[f"{p}.net.{nb}.beta"])
        w = O.fold_weight_norm(sd, f"{p}.net.{nb + 1}")
        x, cd = cached_conv(st, f"{p}.net.{nb + 1}", x, w, sd[f"{p}.net.{nb + 1}.bias"], O.same_padding(2 * f), f, 1, cd)
        scale //= f
    n = len(cfg.factors)
    x = O.snake_beta(x, sd[f"encoder.net.{n + 1}.alpha"], sd[f"encoder.net.{n + 1}.beta"])
    w = O.fold_weight_norm(sd, f"encoder.net.{n + 2}")
    x, cd = cached_conv(st, f"encoder.net.{n + 2}", x, w, sd[f"encoder.net.{n + 2}.bias"], O.same_padding(3), 1, 1, cd)
    st["__cumulative_delay__"] = torch.tensor(cd)
    return x


def encoder_cumulative_delay(cfg) -> int:
    """Latent frames by which the streamed encoder lags the offline one (the reference's ``cumulative_delay`` chain)."""
    cd = 1  # to_in: ResnetBlock1d, k = 3
    for f in cfg.factors:
        cd += sum(cfg.dilations[:cfg.num_blocks]) * (cfg.kernel_size - 1) // 2
        r = O.same_padding(2 * f)[1]
        cd = (r + stride_delay(r, cd, f) + cd) // f
    return cd + 1


# ----------------------------------------------------------------------------- streaming codec decoder (overlap-add)
def ae_decode_stream(sd, cfg, st: State, z: Tensor, gn_latent_frames: int = 64, n_fade: int = 4) -> Tensor:
    """``AE_notcausal.decode`` (export_autoencoder.py:128-153): offline decoder over [z_buffer ; z] with stream GroupNorm,
    linear cross-fade of the first n_fade latent frames with the tail kept from the previous call."""
    n = z.shape[0]
    r = cfg.ratio
    zb = st.get("z_buffer")
    if zb is None or zb.shape[0] != n:
        zb = z.new_zeros(n, z.shape[1], n_fade)
    zc = torch.cat([zb, z], -1)
    lat = zc.shape[-1]

    def norm(key, x, groups, w, b):
        return stream_group_norm(st, "decoder:" + key, x, groups, w, b, gn_latent_frames * (x.shape[-1] // lat))

    x = O.ae_decode(sd, cfg, zc, norm=norm)
    st["z_buffer"] = zc[..., -n_fade:].clone()
    ob = st.get("out_buffer")
    if ob is None or ob.shape[0] != n:
        ob = x.new_zeros(n, 1, r * n_fade)
    alpha = torch.linspace(0, 1, n_fade * r)[None, None, :].to(x)
    x = x.clone()
    x[..., :r * n_fade] = (1 - alpha) * ob + alpha * x[..., :r * n_fade]
    st["out_buffer"] = x[..., -r * n_fade:].clone()
    return x[..., :-r * n_fade]


# ----------------------------------------------------------------------------- streaming structure encoder
def _v2_conv_block_stream(sd, st, prefix, x, causal):
    """V2ConvBlock1D under cached_conv (encoder.py:25-71): conv1 -> conv2 chain their delays, the identity branch is
    delayed by the total (zero with causal padding)."""
    k = sd[prefix + ".net.branches.0.2.weight_v"].shape[-1]
    pad = O.same_padding(k, 1, causal)
    y = F.silu(O._bn_eval(sd, prefix + ".net.branches.0.0", x))
    y, cd = cached_conv(st, prefix + ".net.branches.0.2", y, O.fold_weight_norm(sd, prefix + ".net.branches.0.2"),
                        sd[prefix + ".net.branches.0.2.bias"], pad)
    y = F.silu(O._bn_eval(sd, prefix + ".net.branches.0.3", y))
    y, cd = cached_conv(st, prefix + ".net.branches.0.6", y, O.fold_weight_norm(sd, prefix + ".net.branches.0.6"),
                        sd[prefix + ".net.branches.0.6.bias"], pad, cd=cd)
    return y + cached_pad(st, prefix + ".net.paddings.1", x, cd, crop=True)


def encoder1d_forward_stream(sd, cfg, st: State, z: Tensor) -> Tensor:
    """``Encoder1D.forward_stream`` (encoder.py:300-322) with cached convolutions, all ratios == 1."""
    sd = O._cast(sd, z.dtype)
    x = z
    n = len(cfg.channels)
    assert all(r == 1 for r in cfg.ratios), "oracle covers the shipped (ratio 1) configuration"
    for i in range(n):
        x = _v2_conv_block_stream(sd, st, f"net.{i}.net.0", x, cfg.causal)
        x = O._wn_conv(sd, f"net.{i}.net.1", x)  # 1x1 pool: no state
    x = _v2_conv_block_stream(sd, st, f"net.{n}", x, cfg.causal)
    return torch.tanh(x) if cfg.use_tanh else x
