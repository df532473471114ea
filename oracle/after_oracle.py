"""CPU oracle for the AFTER sampling hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

A functional (state-dict driven) restatement, in plain PyTorch CPU ops, of what the
reference computes on the path SURVEY.md section 8 scopes: ``DenoiserV2.forward``,
``RectifiedFlow.model_forward`` / ``sample``, ``AutoEncoder.encode`` / ``decode`` and the
structure encoder ``Encoder1D``.  Every function cites the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` may import
this module; the product package ``after_b200`` never does (it fails loudly when its CUDA
library is missing).

Parity pinning: the reference ships no tests or golden vectors, so this
oracle is pinned against outputs of the UNMODIFIED reference modules run in the authoring
container -- ``tests/golden/make_golden.py`` (committed) imports them from /root/reference under
the shims in ``tests/golden/ref_shims.py``, loads the same synthetic state dicts, and stores
input/output fixtures under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this
file against those fixtures.

The restatement deliberately differs in form from the reference (no nn.Modules, no mask
tensors built by Python loops, weight-norm folded up front, the CFG batch built by indexing),
so that agreement with the fixtures is evidence about the algorithm, not about copied code.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


def _cast(sd: StateDict, dtype) -> StateDict:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


# =============================================================================
# Denoiser (after/diffusion/networks/transformerv2.py)
# =============================================================================
def fourier_features(t: Tensor, n_dims: int, factor: float = 100.0,
                     max_positions: float = 10000.0) -> Tensor:
    """[cos(u), sin(u)], u_k = factor * t * max_positions^(-k/(n/2)); transformerv2.py:31-43."""
    half = n_dims // 2
    k = torch.arange(half, dtype=torch.float32)
    freqs = ((1.0 / max_positions)**(k / half)).to(t.dtype)
    u = torch.outer(t.reshape(-1) * factor, freqs)
    return torch.cat([u.cos(), u.sin()], dim=1)


def band_allowed(seq_len: int, chunk: int, window: int) -> Tensor:
    """Boolean (L, L) matrix, True where query j may attend key p; transformerv2.py:62-96.

    allowed(j, p)  <=>  p in [c, min(c+chunk, L))  or  (p < c and p >= j - window + 1),
    with c = chunk * floor(j / chunk)  (SURVEY.md appendix A.2)."""
    j = torch.arange(seq_len).unsqueeze(1)
    p = torch.arange(seq_len).unsqueeze(0)
    c = (j // chunk) * chunk
    in_chunk = (p >= c) & (p < c + chunk)
    back = (p < c) & (p >= j - window + 1)
    return in_chunk | back


def rope_rotate(x: Tensor, rot_dim: int = 32, theta: float = 10000.0, offset: int = 0) -> Tensor:
    """Interleaved-pair RoPE on the first ``rot_dim`` features of (..., L, dh);
    rotary_embedding.py:143-173, 196-236, 321-362."""
    L = x.shape[-2]
    inv = 1.0 / (theta**(torch.arange(0, rot_dim, 2, dtype=torch.float32) / rot_dim))
    pos = torch.arange(offset, offset + L, dtype=torch.float32)
    ang = torch.outer(pos, inv).to(x.dtype)  # (L, rot/2), fp32 product like the reference
    cos, sin = ang.cos(), ang.sin()
    xe, xo = x[..., 0:rot_dim:2], x[..., 1:rot_dim:2]
    re = xe * cos - xo * sin
    ro = xo * cos + xe * sin
    rot = torch.stack([re, ro], dim=-1).flatten(-2)
    return torch.cat([rot, x[..., rot_dim:]], dim=-1)


class StreamCache:
    """Rolling per-(layer, diffusion step) key/value history of the streaming denoiser:
    ``MHAttention.k_cache / v_cache`` of shape (max_batch, max_steps, H, cache, dh), zero-initialised
    (so the first blocks really attend to all-zero keys/values), plus the un-rotated k/v of the last
    forward (``last_k / last_v``); transformerv2.py:143-188."""

    def __init__(self, cfg, cache_size: int, max_batch: int = 4, max_steps: int = 16, dtype=torch.float32):
        shape = (max_batch, max_steps, cfg.n_heads, cache_size, cfg.head_dim)
        self.cache_size = cache_size
        self.k = [torch.zeros(shape, dtype=dtype) for _ in range(cfg.n_layers)]
        self.v = [torch.zeros(shape, dtype=dtype) for _ in range(cfg.n_layers)]
        self.last_k = [None] * cfg.n_layers
        self.last_v = [None] * cfg.n_layers

    def roll(self, roll_size: int, cache_index: int) -> None:
        """``DenoiserV2.roll_cache``: append the first ``roll_size`` frames of the last block, keep the newest
        ``cache_size``; transformerv2.py:167-186."""
        for l in range(len(self.k)):
            lk, lv = self.last_k[l], self.last_v[l]
            n = lk.shape[0]
            k = torch.cat([self.k[l][:n, cache_index], lk[:, :, :roll_size]], dim=2)[:, :, -self.cache_size:]
            v = torch.cat([self.v[l][:n, cache_index], lv[:, :, :roll_size]], dim=2)[:, :, -self.cache_size:]
            self.k[l][:n, cache_index] = k
            self.v[l][:n, cache_index] = v


def denoiser_forward(sd: StateDict, cfg, x: Tensor, time: Tensor, cond: Tensor,
                     time_cond: Tensor, taps: Optional[dict] = None, cache: Optional[StreamCache] = None,
                     cache_index: int = 0) -> Tensor:
    """``DenoiserV2.forward``; transformerv2.py:517-543, 437-457, 340-362, 190-236.  x (N,C,T),
    time (N,)|(N,1,1)|(N,1,T), cond (N,zt), time_cond (N,zs,T) -> (N,C,T).  ``cache`` = None is the offline
    path (max_cache_size = 0); with a ``StreamCache`` the keys/values of the block are appended to the cached
    history of ``cache_index`` (keys are cached UN-rotated; queries are rotated at offset = history length,
    rotary_embedding.py:215-236) and the band mask is the last T rows of the mask over history + block."""
    dt = x.dtype
    sd = _cast(sd, dt)
    D, H, dh = cfg.embed_dim, cfg.n_heads, cfg.head_dim
    if time.dim() > 1:
        time = time[..., 0]  # transformerv2.py:524-528
    time = time.reshape(-1).to(dt)
    N, _, T = x.shape

    emb_in = torch.cat([
        fourier_features(time, cfg.noise_embed_dims, cfg.fourier_factor, cfg.fourier_max_positions),
        cond
    ], dim=-1)
    feat = F.linear(F.gelu(F.linear(emb_in, sd["embedding.0.weight"], sd["embedding.0.bias"])),
                    sd["embedding.2.weight"], sd["embedding.2.bias"])  # (N, D)

    tb = "denoiser_trans_block."
    h = F.gelu(
        F.linear(x.transpose(1, 2), sd[tb + "patchify_and_embed.1.weight"],
                 sd[tb + "patchify_and_embed.1.bias"]))  # (N, T, D)
    tc = F.gelu(
        F.linear(time_cond.transpose(1, 2), sd[tb + "patchify_and_embed_tcond.1.weight"],
                 sd[tb + "patchify_and_embed_tcond.1.bias"]))  # (N, T, zs)

    hist = cache.cache_size if cache is not None else 0
    allowed = band_allowed(hist + T, cfg.attention_chunk_size, cfg.local_attention_size)[hist:]
    bias = torch.zeros(T, hist + T, dtype=dt).masked_fill(~allowed, float("-inf"))
    if taps is not None:
        taps["features"] = feat
        taps["h0"] = h

    for i in range(cfg.n_layers):
        p = f"{tb}decoder_blocks.{i}."
        # AdaLN on the per-frame structure condition (transformerv2.py:345-349)
        a_t, b_t = F.linear(tc, sd[p + "tcond_linear.weight"], sd[p + "tcond_linear.bias"]).chunk(2, -1)
        h = F.layer_norm(h, (D, )) * (1 + a_t) + b_t
        # banded rotary self-attention, no output projection (transformerv2.py:351, 190-236, 267)
        y = F.layer_norm(h, (D, ), sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        q, k, v = F.linear(y, sd[p + "self_attention.qkv_linear.weight"]).chunk(3, dim=2)
        q, k, v = (z.reshape(N, T, H, dh).transpose(1, 2) for z in (q, k, v))
        if cache is not None:
            cache.last_k[i], cache.last_v[i] = k, v
            k = torch.cat([cache.k[i][:N, cache_index].to(dt), k], dim=2)
            v = torch.cat([cache.v[i][:N, cache_index].to(dt), v], dim=2)
        q = rope_rotate(q, cfg.rotary_dim, cfg.rotary_theta, offset=hist)
        k = rope_rotate(k, cfg.rotary_dim, cfg.rotary_theta)
        s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dh) + bias
        a = torch.matmul(torch.softmax(s, dim=-1), v)
        h = h + a.transpose(1, 2).reshape(N, T, D)
        # AdaLN on the global (time, timbre) features (transformerv2.py:354-358)
        a_c, b_c = F.linear(feat, sd[p + "linear.weight"], sd[p + "linear.bias"]).chunk(2, -1)
        h = F.layer_norm(h, (D, )) * (1 + a_c.unsqueeze(1)) + b_c.unsqueeze(1)
        # MLP (transformerv2.py:361, 275-280)
        y = F.layer_norm(h, (D, ), sd[p + "norm3.weight"], sd[p + "norm3.bias"])
        y = F.gelu(F.linear(y, sd[p + "mlp.mlp.0.weight"], sd[p + "mlp.mlp.0.bias"]))
        h = h + F.linear(y, sd[p + "mlp.mlp.2.weight"], sd[p + "mlp.mlp.2.bias"])
        if taps is not None:
            taps[f"h{i + 1}"] = h

    out = F.linear(h, sd[tb + "out_proj.0.weight"], sd[tb + "out_proj.0.bias"])
    return out.transpose(1, 2)


# =============================================================================
# Rectified-flow sampler (after/diffusion/model.py)
# =============================================================================
CFG_AUDIO = 0  # rows: (cond, tc) / (drop, tc) / (drop, drop); factor = g_t / max(g_s, clamp)
CFG_MIDI = 1  # rows: (cond, tc) / (cond, drop) / (drop, drop); factor = g_s / max(g_t, clamp)


def model_forward(sd, cfg, x, time, cond, time_cond, guidance_timbre: float,
                  guidance_structure: float, drop_value: float = -4.0, cfg_variant: int = CFG_AUDIO,
                  clamp: float = 0.01, cache: Optional[StreamCache] = None, cache_index: int = 0, net=None) -> Tensor:
    """3-way classifier-free-guidance evaluation; model.py:721-761 (audio variant) and
    after_scripts/export_midi.py:322-360 (midi variant, clamp 0.1).  ``net(x, t, cond, time_cond)`` is the velocity
    network (default: DenoiserV2 with ``sd`` / ``cfg``; ``unet_net(sd, cfg)`` gives the UNET1D one).  A condition the
    network does not take is passed as None."""
    B = x.shape[0]
    idx = torch.arange(B).repeat(3)
    drop_c = torch.full_like(cond, drop_value) if cond is not None else None
    drop_t = torch.full_like(time_cond, drop_value) if time_cond is not None else None
    cat = lambda parts: torch.cat(parts) if parts[0] is not None else None  # noqa: E731
    if cfg_variant == CFG_AUDIO:
        conds = cat([cond, drop_c, drop_c])
        tconds = cat([time_cond, time_cond, drop_t])
        g_first, g_second = guidance_timbre, guidance_structure
    else:
        conds = cat([cond, cond, drop_c])
        tconds = cat([time_cond, drop_t, drop_t])
        g_first, g_second = guidance_structure, guidance_timbre
    t = time.reshape(B, -1)[:, 0]
    if net is None:
        d = denoiser_forward(sd, cfg, x[idx], t[idx], conds, tconds, cache=cache, cache_index=cache_index)
    else:
        d = net(x[idx], t[idx], conds, tconds)
    d_full, d_mid, d_none = d[:B], d[B:2 * B], d[2 * B:]
    total = 0.5 * (guidance_structure + guidance_timbre)
    factor = g_first / max(g_second, clamp)
    return d_none + total * (d_mid + factor * (d_full - d_mid) - d_none)


@torch.no_grad()
def sample(sd, cfg, x0, cond, time_cond, nb_steps: int, guidance_timbre: float = 1.0,
           guidance_structure: float = 1.0, drop_value: float = -4.0, cfg_variant: int = CFG_AUDIO,
           clamp: float = 0.01, net=None) -> Tensor:
    """Fixed-step Euler integration of the velocity field on t_i = i / N; model.py:763-785."""
    x = x0
    B = x0.shape[0]
    ts = torch.linspace(0, 1, nb_steps + 1)[:-1]  # fp32 grid, exactly as the reference builds it
    dt = 1 / nb_steps
    for t in ts:
        tt = t.to(x.dtype).reshape(1).repeat(B)
        x = x + model_forward(sd, cfg, x, tt, cond, time_cond, guidance_timbre,
                              guidance_structure, drop_value, cfg_variant, clamp, net=net) * dt
    return x


def unet_net(sd, cfg):
    """The UNET1D velocity network as a ``net`` for ``model_forward`` / ``sample`` (RectifiedFlow binds either net)."""
    return lambda x, t, cond, time_cond: unet1d_forward(sd, cfg, x, t, cond, time_cond)


@torch.no_grad()
def sample_stream(sd, cfg, cache: StreamCache, x_last, cond, time_cond, nb_steps: int, guidance_timbre: float = 1.0,
                  guidance_structure: float = 1.0, drop_value: float = -4.0, cfg_variant: int = CFG_AUDIO,
                  clamp: float = 0.1) -> Tensor:
    """One audio block of the exported ``Streamer.sample``: Euler step i runs against KV cache i, which is then
    rolled by the block length; after_scripts/export.py:356-416 (guidance ratio clamped at 0.1, :389-390)."""
    x = x_last
    B = x.shape[0]
    ts = torch.linspace(0, 1, nb_steps + 1)[:-1]
    dt = 1 / nb_steps
    for i, t in enumerate(ts):
        tt = t.to(x.dtype).reshape(1).repeat(B)
        x = x + model_forward(sd, cfg, x, tt, cond, time_cond, guidance_timbre, guidance_structure, drop_value,
                              cfg_variant, clamp, cache=cache, cache_index=i) * dt
        cache.roll(x.shape[-1], i)
    return x


# =============================================================================
# Codec (after/autoencoder/networks/SimpleNetsStream.py, pqmf.py, core.py)
# =============================================================================
def fold_weight_norm(sd: StateDict, prefix: str) -> Tensor:
    """w = g * v / ||v||, norm over all dims but 0 (torch.nn.utils.weight_norm, dim=0);
    SimpleNetsStream.py:84-92."""
    v, g = sd[prefix + ".weight_v"], sd[prefix + ".weight_g"]
    n = v.flatten(1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
    return v * (g / n)


def same_padding(kernel: int, dilation: int = 1, causal: bool = False):
    """cached_conv.get_padding (acids-ircam/cached_conv >= 2.5.0, un-vendored; semantics restated
    from the reference call sites SimpleNetsStream.py:45,177,453,590 and SURVEY.md section 8c)."""
    if kernel == 1:
        return (0, 0)
    p = (kernel - 1) * dilation + 1
    if causal:
        return (p // 2 + (p - 1) // 2, 0)
    return ((p - 1) // 2, p // 2)


def snake_beta(x: Tensor, alpha: Tensor, beta: Tensor) -> Tensor:
    """x + sin^2(alpha x) / (beta + 1e-9), per channel; core.py:217-218, 250-258."""
    a = alpha.reshape(1, -1, 1)
    b = beta.reshape(1, -1, 1)
    return x + torch.sin(x * a)**2 / (b + 1e-9)


def _wn_conv(sd, prefix, x, stride=1, dilation=1, pad=None, causal=False):
    w = fold_weight_norm(sd, prefix)
    if pad is None:
        pad = same_padding(w.shape[-1], dilation, causal)
    return F.conv1d(F.pad(x, pad), w, sd[prefix + ".bias"], stride=stride, dilation=dilation)


def _conv_block(sd, prefix, x, dilation=1, groups=8, norm=None):
    """GroupNorm(min(C,8)) -> SnakeBeta -> conv; SimpleNetsStream.py:150-194.  ``norm(key, x, groups, weight, bias)``
    replaces the offline GroupNorm (the streaming oracle passes CachedGroupNorm's stream branch); a codec built with
    ``use_norm = False`` has no gn tensors and skips the norm (SimpleNetsStream.py:165-167)."""
    C = x.shape[1]
    y = x
    if (prefix + ".net.0.gn.weight") in sd:
        w, b = sd[prefix + ".net.0.gn.weight"], sd[prefix + ".net.0.gn.bias"]
        y = norm(prefix + ".net.0", x, min(C, groups), w, b) if norm is not None else F.group_norm(x, min(C, groups), w, b, eps=1e-5)
    y = snake_beta(y, sd[prefix + ".net.1.alpha"], sd[prefix + ".net.1.beta"])
    return _wn_conv(sd, prefix + ".net.2", y, dilation=dilation)


def _resnet(sd, prefix, x, dilation=1, groups=8, norm=None):
    """block2(block1(x)) + skip(x); SimpleNetsStream.py:197-254."""
    y = _conv_block(sd, prefix + ".net.branches.0.0", x, dilation, groups, norm)
    y = _conv_block(sd, prefix + ".net.branches.0.1", y, 1, 8, norm)
    skip_key = prefix + ".net.branches.1.weight_v"
    skip = _wn_conv(sd, prefix + ".net.branches.1", x) if skip_key in sd else x
    return y + skip


def pqmf_analysis(sd, x: Tensor) -> Tensor:
    """(B,1,S) -> (B,M,S/M): strided FIR bank then negate odd bands at even frames;
    pqmf.py:16-20, 263-271, 286-290."""
    w = sd["pqmf.forward_conv.weight"]
    M = w.shape[0]
    y = F.conv1d(F.pad(x, same_padding(w.shape[-1])), w, stride=M)
    sign = torch.ones(M, y.shape[-1], dtype=y.dtype)
    sign[1::2, ::2] = -1
    return y * sign


def pqmf_synthesis(sd, x: Tensor) -> Tensor:
    """(B,M,T) -> (B,1,M*T); pqmf.py:292-301."""
    w = sd["pqmf.inverse_conv.weight"]
    M = w.shape[0]
    sign = torch.ones(M, x.shape[-1], dtype=x.dtype)
    sign[1::2, ::2] = -1
    y = F.conv1d(F.pad(x * sign, same_padding(w.shape[-1])), w) * M
    y = y.flip(1)  # band order reversed
    return y.transpose(1, 2).reshape(x.shape[0], 1, -1)  # sample m of frame t -> t*M + m


def ae_encode(sd: StateDict, cfg, audio: Tensor) -> Tensor:
    """``AutoEncoder.encode`` (bottleneck = identity at inference); SimpleNetsStream.py:918-941,
    400-459, 301-341, 753-760."""
    sd = _cast(sd, audio.dtype)
    x = pqmf_analysis(sd, audio) if cfg.pqmf_bands > 1 else audio
    x = _resnet(sd, "encoder.net.0", x, 1, cfg.resnet_groups)
    nb = cfg.num_blocks
    for i, f in enumerate(cfg.factors):
        p = f"encoder.net.{i + 1}"
        for j in range(nb):
            x = _resnet(sd, f"{p}.net.{j}", x, cfg.dilations[j], cfg.resnet_groups)
        x = snake_beta(x, sd[f"{p}.net.{nb}.alpha"], sd[f"{p}.net.{nb}.beta"])
        x = _wn_conv(sd, f"{p}.net.{nb + 1}", x, stride=f, pad=same_padding(2 * f))
    n = len(cfg.factors)
    x = snake_beta(x, sd[f"encoder.net.{n + 1}.alpha"], sd[f"encoder.net.{n + 1}.beta"])
    return _wn_conv(sd, f"encoder.net.{n + 2}", x)


def ae_decode(sd: StateDict, cfg, z: Tensor, norm=None) -> Tensor:
    """``AutoEncoder.decode``; SimpleNetsStream.py:943-954, 552-651, 344-384, 51-70.  ``norm``: see ``_conv_block``."""
    sd = _cast(sd, z.dtype)
    x = _wn_conv(sd, "decoder.net.0", z)
    nb = cfg.num_blocks
    for i, f in enumerate(cfg.factors[::-1]):
        p = f"decoder.net.{i + 1}"
        x = snake_beta(x, sd[f"{p}.net.0.alpha"], sd[f"{p}.net.0.beta"])
        w = fold_weight_norm(sd, f"{p}.net.1")  # (in, out, k): norm per *input* channel
        x = F.conv_transpose1d(x, w, sd[f"{p}.net.1.bias"], stride=f, padding=f // 2)
        for j in range(nb):
            x = _resnet(sd, f"{p}.net.{2 + j}", x, cfg.dilations[j], cfg.resnet_groups, norm)
    x = _conv_block(sd, "decoder.synth.branches.0.net.0", x, 1, cfg.resnet_groups, norm)
    x = _conv_block(sd, "decoder.synth.branches.0.net.1", x, 1, 8, norm)
    if cfg.use_loudness:
        half = x.shape[1] // 2
        x = x[:, :half] * torch.sigmoid(x[:, half:])
    return pqmf_synthesis(sd, x) if cfg.pqmf_bands > 1 else x


# =============================================================================
# Structure encoder (after/diffusion/networks/encoder.py)
# =============================================================================
def _bn_eval(sd, prefix, x, eps=1e-5):
    s = sd[prefix + ".weight"] / torch.sqrt(sd[prefix + ".running_var"] + eps)
    return x * s.reshape(1, -1, 1) + (sd[prefix + ".bias"] - sd[prefix + ".running_mean"] * s).reshape(1, -1, 1)


def _v2_conv_block(sd, prefix, x, causal):
    """x + conv(SiLU(BN(conv(SiLU(BN(x)))))) (dropout off); encoder.py:25-71."""
    y = F.silu(_bn_eval(sd, prefix + ".net.branches.0.0", x))
    y = _wn_conv(sd, prefix + ".net.branches.0.2", y, causal=causal)
    y = F.silu(_bn_eval(sd, prefix + ".net.branches.0.3", y))
    y = _wn_conv(sd, prefix + ".net.branches.0.6", y, causal=causal)
    return x + y


def encoder1d_forward(sd: StateDict, cfg, z: Tensor) -> Tensor:
    """``Encoder1D.forward`` with all ratios == 1; encoder.py:273-298, 74-113, 116-237."""
    sd = _cast(sd, z.dtype)
    x = z
    n = len(cfg.channels)
    assert all(r == 1 for r in cfg.ratios), "oracle covers the shipped (ratio 1) configuration"
    for i in range(n):
        x = _v2_conv_block(sd, f"net.{i}.net.0", x, cfg.causal)
        x = _wn_conv(sd, f"net.{i}.net.1", x)
    x = _v2_conv_block(sd, f"net.{n}", x, cfg.causal)
    return torch.tanh(x) if cfg.use_tanh else x


# =============================================================================
# Timbre encoder (after/diffusion/networks/ecapa_encoder.py)
# =============================================================================
def _reflect_conv(sd, prefix, x, dilation=1):
    """``Conv1dSamePaddingReflect`` (stride 1): reflect-pad (k-1)*d/2 each side; ecapa_encoder.py:12-82."""
    w = sd[prefix + ".weight"]
    pad = (w.shape[-1] - 1) * dilation // 2
    if pad:
        x = F.pad(x, (pad, pad), mode="reflect")
    return F.conv1d(x, w, sd[prefix + ".bias"], dilation=dilation)


def _tdnn(sd, prefix, x, dilation=1):
    """conv -> ReLU -> BatchNorm(eval); ecapa_encoder.py:85-138."""
    return _bn_eval(sd, prefix + ".norm", torch.relu(_reflect_conv(sd, prefix + ".conv.conv", x, dilation)))


def _attentive_stats(x, w, eps=1e-12):
    mean = (w * x).sum(dim=2)
    std = torch.sqrt((w * (x - mean.unsqueeze(2))**2).sum(dim=2).clamp(eps))
    return mean, std


def ecapa_forward(sd: StateDict, cfg, z: Tensor) -> Tensor:
    """``ECAPATDNN.forward`` (pooling, global context, groups == 1); ecapa_encoder.py:567-624, 141-455.
    z (B, in_size, T) -> (B, out_dim)."""
    sd = _cast(sd, z.dtype)
    ch = cfg.channels
    x = _tdnn(sd, "blocks.0", z, cfg.dilations[0])
    feats = []
    for i in range(1, len(ch) - 1):
        p = f"blocks.{i}"
        res = _reflect_conv(sd, p + ".shortcut.conv", x) if (p + ".shortcut.conv.weight") in sd else x
        y = _tdnn(sd, p + ".tdnn1", x)
        parts = list(torch.chunk(y, cfg.res2net_scale, dim=1))  # Res2Net: chained sub-band TDNNs
        outs = [parts[0]]
        prev = None
        for j in range(cfg.res2net_scale - 1):
            inp = parts[j + 1] if j == 0 else parts[j + 1] + prev
            prev = _tdnn(sd, f"{p}.res2net_block.blocks.{j}", inp, cfg.dilations[i])
            outs.append(prev)
        y = _tdnn(sd, p + ".tdnn2", torch.cat(outs, dim=1))
        s = y.mean(dim=2, keepdim=True)  # squeeze-excitation
        s = torch.relu(_reflect_conv(sd, p + ".se_block.conv1.conv", s))
        s = torch.sigmoid(_reflect_conv(sd, p + ".se_block.conv2.conv", s))
        x = s * y + res
        feats.append(x)
    x = _tdnn(sd, "mfa", torch.cat(feats, dim=1), cfg.dilations[-1])
    # attentive statistics pooling
    T = x.shape[-1]
    if cfg.global_context:
        mean, std = _attentive_stats(x, torch.tensor(1.0 / T, dtype=x.dtype))
        a = torch.cat([x, mean.unsqueeze(2).expand(-1, -1, T), std.unsqueeze(2).expand(-1, -1, T)], dim=1)
    else:
        a = x
    a = _reflect_conv(sd, "asp.conv.conv", torch.tanh(_tdnn(sd, "asp.tdnn", a)))
    a = torch.softmax(a, dim=2)
    mean, std = _attentive_stats(x, a)
    v = _bn_eval(sd, "asp_bn", torch.cat([mean, std], dim=1).unsqueeze(2))
    out = _reflect_conv(sd, "fc.conv", v).squeeze(2)
    return torch.tanh(out) if cfg.use_tanh else out


# =============================================================================
# UNET1D conv denoiser (after/diffusion/networks/unet1d.py, blocks.py) -- SURVEY.md section 8f rank 3.
# No shipped config binds it and no CUDA path exists yet; the oracle is pinned so that one can be built against it.
# =============================================================================
def _spe(t: Tensor, dim: int, max_positions: float = 10000.0, scale: float = 32.0) -> Tensor:
    """[sin(w x), cos(w x)], x = scale * t, w_f = max_positions^(-2 f / dim); unet1d.py:7-25."""
    x = t.reshape(-1) * scale
    f = torch.arange(dim // 2, dtype=torch.float32)
    w = ((1.0 / max_positions)**(2 * f / dim)).to(x.dtype)
    a = x[:, None] * w[None, :]
    return torch.cat([a.sin(), a.cos()], dim=-1)


def _same_conv(sd, prefix, x, stride=1):
    """nn.Conv1d with padding='same' (stride 1) or padding = k // 2 (strided); unet1d.py:46-50, 149-158."""
    w = sd[prefix + ".weight"]
    return F.conv1d(x, w, sd[prefix + ".bias"], stride=stride, padding=w.shape[-1] // 2)


def _unet_conv_block(sd, prefix, x, time_emb, cond, skip=None, time_cond=None, res=True):
    """GN -> SiLU -> conv -> (x * t_mult + t_add) -> (x * c_mult + c_add) -> GN -> SiLU -> conv (+ to_out(x_in));
    unet1d.py:84-118.  The residual branch sees the input BEFORE skip / time_cond are concatenated."""
    x_in = x
    parts = [x] + ([skip] if skip is not None else []) + ([time_cond] if time_cond is not None else [])
    h = torch.cat(parts, dim=1)
    C = h.shape[1]
    h = F.silu(F.group_norm(h, min(16, C // 4), sd[prefix + ".gn1.weight"], sd[prefix + ".gn1.bias"]))
    h = _same_conv(sd, prefix + ".conv1", h)
    t = F.linear(F.silu(F.linear(time_emb, sd[prefix + ".time_mlp.0.weight"], sd[prefix + ".time_mlp.0.bias"])),
                 sd[prefix + ".time_mlp.2.weight"], sd[prefix + ".time_mlp.2.bias"])
    t_mult, t_add = t.chunk(2, dim=1)
    h = h * t_mult[:, :, None] + t_add[:, :, None]
    if (prefix + ".cond_mlp.0.weight") in sd:
        c = F.linear(F.silu(F.linear(cond, sd[prefix + ".cond_mlp.0.weight"], sd[prefix + ".cond_mlp.0.bias"])),
                     sd[prefix + ".cond_mlp.2.weight"], sd[prefix + ".cond_mlp.2.bias"])
        c_mult, c_add = c.chunk(2, dim=1)
        h = h * c_mult[:, :, None] + c_add[:, :, None]
    Co = h.shape[1]
    h = F.silu(F.group_norm(h, min(16, Co // 4), sd[prefix + ".gn2.weight"], sd[prefix + ".gn2.bias"]))
    h = _same_conv(sd, prefix + ".conv2", h)
    if not res:
        return h
    if (prefix + ".to_out.weight") in sd:
        x_in = _same_conv(sd, prefix + ".to_out", x_in)
    return h + x_in


def _self_attention_1d(sd, prefix, x, n_head):
    """GroupNorm(1) -> 1x1 qkv -> full softmax attention per head (scale d^-1/4 on q and k) -> 1x1 out + residual;
    blocks.py:201-243."""
    n, c, s = x.shape
    y = F.group_norm(x, 1, sd[prefix + ".norm.weight"], sd[prefix + ".norm.bias"])
    qkv = F.conv1d(y, sd[prefix + ".qkv_proj.weight"], sd[prefix + ".qkv_proj.bias"])
    qkv = qkv.reshape(n, n_head * 3, c // n_head, s).transpose(2, 3)
    q, k, v = qkv.chunk(3, dim=1)
    scale = k.shape[3]**-0.25
    att = torch.softmax((q * scale) @ (k.transpose(2, 3) * scale), dim=3)
    y = (att @ v).transpose(2, 3).reshape(n, c, s)
    return x + F.conv1d(y, sd[prefix + ".out_proj.weight"], sd[prefix + ".out_proj.bias"])


def unet1d_forward(sd: StateDict, cfg, x: Tensor, time: Tensor, cond: Optional[Tensor], time_cond: Optional[Tensor]) -> Tensor:
    """``UNET1D.forward``; unet1d.py:376-429 (both the per-scale time_cond embedding path and the concat path).
    x (N, in_size, T), time (N,), cond (N, cond_channels), time_cond (N, time_cond_in_channels, T) -> (N, out_size, T)."""
    sd = _cast(sd, x.dtype)
    n = len(cfg.channels)
    ratios = [1] + list(cfg.ratios)
    temb = _spe(time.to(x.dtype), cfg.time_channels)
    skips, tcs = [], []
    tc = time_cond
    if not cfg.time_cond_channels:
        if cfg.time_cond_in_channels:
            x = torch.cat([x, time_cond], dim=1)
        tc = None
    for i in range(n):
        if cfg.time_cond_channels:
            tc = F.silu(_same_conv(sd, f"cond_emb_time.{i}.0", tc, stride=1 if i == 0 else ratios[i - 1]))
        p = f"down_layers.{i}"
        skip = _unet_conv_block(sd, p + ".conv", x, temb, cond, time_cond=tc)
        if (p + ".self_attn.norm.weight") in sd:
            skip = _self_attention_1d(sd, p + ".self_attn", skip, 4)
        x = _same_conv(sd, p + ".pool", skip, stride=ratios[i])
        skips.append(skip)
        tcs.append(tc)
    if cfg.time_cond_channels:
        tc = F.silu(_same_conv(sd, f"cond_emb_time.{n}.0", tc, stride=ratios[n - 1]))
    x = _unet_conv_block(sd, "middle_block.conv", x, temb, cond, time_cond=tc)
    if "middle_block.self_attn.norm.weight" in sd:
        x = _self_attention_1d(sd, "middle_block.self_attn", x, cfg.channels[-1] // 32)
    for i in range(1, n + 1):
        p = f"up_layers.{i - 1}"
        skip, tc_i = skips.pop(), tcs.pop()
        ratio = ratios[n - i]
        if ratio != 1:
            x = F.interpolate(x, scale_factor=ratio, mode="nearest")
            x = _same_conv(sd, p + ".up.1", x)
        elif (p + ".up.weight") in sd:
            x = _same_conv(sd, p + ".up", x)
        x = _unet_conv_block(sd, p + ".conv", x, temb, cond, skip=skip, time_cond=tc_i,
                             res=(cfg.use_res_last if i == n else True))
        if (p + ".self_attn.norm.weight") in sd:
            x = _self_attention_1d(sd, p + ".self_attn", x, 4)
    return x


# ---------------------------------------------------------------------------------------------------------------------
# 2-D timbre map of the exported model: Streamer.latent2map / map2latent (after_scripts/export.py:494-508) over the
# export-time projection (after/diffusion/latent_plot.py:20-37: Linear-GELU-Linear-GELU-Linear).  Pinned against the
# reference module by tests/golden/latent_map.npz (tests/golden/make_golden_latent_map.py).
# ---------------------------------------------------------------------------------------------------------------------
def latent_map(sd, x, direction):
    """direction 0 = latent2map (``project_model.encoder``), 1 = map2latent (``project_model.decoder``); ``sd`` None is the
    identity projection of ``--nolatent_project`` (export.py:143).  x (B, C_in, T) -> (B, C_out, T)."""
    v = x.mean(-1)
    if sd is not None:
        half = "encoder" if direction == 0 else "decoder"
        for i in (0, 2, 4):
            v = F.linear(v, sd[f"{half}.{i}.weight"], sd[f"{half}.{i}.bias"])
            if i < 4:
                v = F.gelu(v)
    return v.unsqueeze(-1).repeat(1, 1, x.shape[-1])
