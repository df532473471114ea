"""Import the UNMODIFIED reference leaf modules from /root/reference under small shims.

Only used in the authoring container (by ``make_golden.py`` and by the optional
``reference``-marked tests) -- /root/reference does not exist on the GPU box, so nothing
on the ``-m gpu`` / smoke / bench path imports this file.

Shims:
  * ``gin``           -> ``configurable`` = identity decorator
  * ``cached_conv``   -> non-cached semantics of acids-ircam/cached_conv>=2.5.0
                         (``get_padding``, ``Conv1d`` with tuple padding, ``ConvTranspose1d``,
                         ``CachedSequential``, ``AlignBranches``)
  * ``einops_exts``, ``torch_ema``, ``nn_tilde`` -> import stubs
  * scipy>=1.13: ``scipy.signal.kaiser`` alias and ``firwin(nyq=)`` -> ``fs=2*nyq``
"""
import importlib
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("AFTER_REFERENCE", "/root/reference")

_PAD_MODE = ["centered"]


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "after"))


# ----------------------------------------------------------------------------- gin
def _make_gin():
    gin = types.ModuleType("gin")

    def configurable(*args, **kwargs):
        if len(args) == 1 and callable(args[0]) and not kwargs:
            return args[0]

        def deco(f):
            return f

        return deco

    gin.configurable = configurable
    gin.add_config_file_search_path = lambda *a, **k: None
    gin.parse_config_file = lambda *a, **k: None
    gin.REQUIRED = object()
    return gin


# ----------------------------------------------------------------------------- cached_conv
def _make_cached_conv():
    cc = types.ModuleType("cached_conv")
    convs = types.ModuleType("cached_conv.convs")

    def get_padding(kernel_size, stride=1, dilation=1, mode=None):
        mode = _PAD_MODE[0] if mode is None else mode
        if kernel_size == 1:
            return (0, 0)
        p = (kernel_size - 1) * dilation + 1
        if mode == "centered":
            return ((p - 1) // 2, p // 2)
        if mode == "causal":
            return (p // 2 + (p - 1) // 2, 0)
        raise ValueError(mode)

    class Conv1d(nn.Conv1d):

        def __init__(self, *args, **kwargs):
            padding = kwargs.pop("padding", 0)
            kwargs.pop("cumulative_delay", None)
            if isinstance(padding, int):
                padding = (padding, padding)
            self._pad = tuple(int(p) for p in padding)
            super().__init__(*args, **kwargs)
            self.cumulative_delay = 0

        def forward(self, x):
            x = nn.functional.pad(x, self._pad)
            return nn.functional.conv1d(x, self.weight, self.bias, self.stride, 0,
                                        self.dilation, self.groups)

    class ConvTranspose1d(nn.ConvTranspose1d):

        def __init__(self, *args, **kwargs):
            kwargs.pop("cumulative_delay", None)
            super().__init__(*args, **kwargs)
            self.cumulative_delay = 0

    class CachedSequential(nn.Sequential):

        def __init__(self, *args, **kwargs):
            super().__init__(*args)
            self.cumulative_delay = 0

    class AlignBranches(nn.Module):

        def __init__(self, *branches, delays=None, cumulative_delay=0, stride=1):
            super().__init__()
            self.branches = nn.ModuleList(branches)
            self.cumulative_delay = 0

        def forward(self, x):
            return [b(x) for b in self.branches]

    class Branches(AlignBranches):
        pass

    def use_cached_conv(state: bool):
        assert not state, "shim only implements the non-cached (offline) path"

    for mod in (cc, convs):
        mod.get_padding = get_padding
        mod.Conv1d = Conv1d
        mod.ConvTranspose1d = ConvTranspose1d
        mod.CachedSequential = CachedSequential
        mod.AlignBranches = AlignBranches
        mod.Branches = Branches
        mod.use_cached_conv = use_cached_conv
        mod.USE_BUFFER_CONV = False
    cc.convs = convs
    return cc, convs


class padding_mode:
    """Context manager standing in for gin's scoped ``convs.get_padding.mode`` binding."""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.prev = _PAD_MODE[0]
        _PAD_MODE[0] = self.mode

    def __exit__(self, *a):
        _PAD_MODE[0] = self.prev


_loaded = {}


def _stub(name):
    m = types.ModuleType(name)
    m.__path__ = []
    sys.modules[name] = m
    return m


def _load(modname, relpath):
    if modname in _loaded:
        return _loaded[modname]
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    _loaded[modname] = mod
    return mod


def install():
    """Install shims + namespace stubs; returns a namespace with the reference leaf modules."""
    if "ns" in _loaded:
        return _loaded["ns"]
    if not reference_available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")

    sys.modules.setdefault("gin", _make_gin())
    if "cached_conv" not in sys.modules:
        cc, convs = _make_cached_conv()
        sys.modules["cached_conv"] = cc
        sys.modules["cached_conv.convs"] = convs
    if "einops_exts" not in sys.modules:
        m = types.ModuleType("einops_exts")
        m.rearrange_many = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
        sys.modules["einops_exts"] = m
    if "torch_ema" not in sys.modules:
        m = types.ModuleType("torch_ema")
        m.ExponentialMovingAverage = object
        sys.modules["torch_ema"] = m

    # scipy drift (reference pins scipy==1.12): kaiser moved, firwin(nyq=) removed
    import scipy.signal
    import scipy.signal.windows
    if not hasattr(scipy.signal, "kaiser"):
        scipy.signal.kaiser = scipy.signal.windows.kaiser
    _firwin = scipy.signal.firwin
    if not getattr(_firwin, "_after_shim", False):

        def firwin(*args, nyq=None, **kwargs):
            if nyq is not None:
                kwargs["fs"] = 2.0 * nyq
            return _firwin(*args, **kwargs)

        firwin._after_shim = True
        scipy.signal.firwin = firwin

    for name in ("after", "after.diffusion", "after.diffusion.networks", "after.autoencoder",
                 "after.autoencoder.networks"):
        if name not in sys.modules:
            _stub(name)

    ns = types.SimpleNamespace()
    ns.core = _load("after.autoencoder.core", "after/autoencoder/core.py")
    ns.pqmf = _load("after.autoencoder.networks.pqmf", "after/autoencoder/networks/pqmf.py")
    ns.ae = _load("after.autoencoder.networks.SimpleNetsStream",
                  "after/autoencoder/networks/SimpleNetsStream.py")
    ns.rotary = _load("after.diffusion.networks.rotary_embedding",
                      "after/diffusion/networks/rotary_embedding.py")
    ns.transformerv2 = _load("after.diffusion.networks.transformerv2",
                             "after/diffusion/networks/transformerv2.py")
    ns.encoder = _load("after.diffusion.networks.encoder", "after/diffusion/networks/encoder.py")
    ns.ecapa = _load("after.diffusion.networks.ecapa_encoder",
                     "after/diffusion/networks/ecapa_encoder.py")
    ns.model = _load("after.diffusion.model", "after/diffusion/model.py")
    ns.padding_mode = padding_mode
    _loaded["ns"] = ns
    return ns


# ----------------------------------------------------------------------------- builders
DENOISER_CFG = {
    "tiny": dict(n_channels=64, seq_len=256, embed_dim=256, cond_dim=6, noise_embed_dims=64,
                 n_layers=6, mlp_multiplier=3, dropout=0.1, causal=True, tcond_dim=12,
                 pos_emb_type="rotary", local_attention_size=8, attention_chunk_size=4),
    "base": dict(n_channels=64, seq_len=256, embed_dim=512, cond_dim=6, noise_embed_dims=64,
                 n_layers=6, mlp_multiplier=3, dropout=0.1, causal=True, tcond_dim=12,
                 pos_emb_type="rotary", local_attention_size=8, attention_chunk_size=4),
    "midi": dict(n_channels=64, seq_len=256, embed_dim=512, cond_dim=6, noise_embed_dims=64,
                 n_layers=6, mlp_multiplier=3, dropout=0.1, causal=True, tcond_dim=128,
                 pos_emb_type="rotary", local_attention_size=16, attention_chunk_size=4),
}

ECAPA_CFG = {
    "tiny": dict(in_size=64, channels=[256, 256, 256, 512]),
    "base": dict(in_size=64, channels=[512, 512, 512, 1024]),
    "midi": dict(in_size=64, channels=[512, 512, 512, 1024]),
}
ENC1D_CFG = {
    "tiny": dict(in_size=64, channels=[64, 128, 256, 256, 12]),
    "base": dict(in_size=64, channels=[64, 128, 256, 512, 12]),
}

AE_CFG = dict(in_channels=16, channels=64, pqmf_bands=16, z_channels=64,
              multipliers=[1, 2, 4, 4, 8, 8], factors=[2, 2, 2, 4, 4], dilations=[1, 3, 9],
              kernel_size=3, use_norm=True, decoder_ratio=1.5, use_loudness=True,
              use_noise=False)


def perturb_(module: nn.Module, seed: int = 1):
    """Randomise parameters that default to identity so that every fold is exercised
    (SURVEY.md section 8d): BN running stats, Snake alpha/beta, GroupNorm/BN affine, weight_g."""
    g = torch.Generator().manual_seed(seed)

    def u(shape, lo, hi):
        return torch.rand(shape, generator=g) * (hi - lo) + lo

    with torch.no_grad():
        for name, m in module.named_modules():
            if isinstance(m, (nn.BatchNorm1d, )):
                m.running_mean.copy_(u(m.running_mean.shape, -0.5, 0.5))
                m.running_var.copy_(u(m.running_var.shape, 0.5, 1.5))
                if m.affine:
                    m.weight.copy_(1 + 0.1 * torch.randn(m.weight.shape, generator=g))
                    m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
            elif isinstance(m, nn.GroupNorm):
                m.weight.copy_(1 + 0.1 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
            elif isinstance(m, nn.LayerNorm) and m.elementwise_affine:
                m.weight.copy_(1 + 0.1 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
            elif type(m).__name__ == "SnakeBeta":
                m.alpha.copy_(u(m.alpha.shape, 0.5, 1.5))
                m.beta.copy_(u(m.beta.shape, 0.5, 1.5))
        for name, p in module.named_parameters():
            if name.endswith("weight_g"):
                p.mul_(u(p.shape, 0.8, 1.2))
    return module


def build_denoiser(name: str, seed: int = 0):
    ns = install()
    torch.manual_seed(seed)
    net = ns.transformerv2.DenoiserV2(**DENOISER_CFG[name])
    perturb_(net, seed + 1)
    return net.eval()


def build_autoencoder(seed: int = 0, **overrides):
    ns = install()
    torch.manual_seed(seed)
    cfg = dict(AE_CFG)
    cfg.update(overrides)
    ae = ns.ae.AutoEncoder(bottleneck=ns.ae.ReluBottleneck(sigma=0.01, scale=3), **cfg)
    perturb_(ae, seed + 1)
    return ae.eval()


def build_encoder1d(name: str, seed: int = 0, **overrides):
    ns = install()
    torch.manual_seed(seed)
    cfg = dict(ratios=[1, 1, 1, 1], kernel_size=5, use_tanh=False, average_out=False,
               upscale_out=False, spherical_normalization=False, vae_regularisation=False,
               ac_regularisation=True)
    cfg.update(ENC1D_CFG[name])
    cfg.update(overrides)
    cfg["channels"] = list(cfg["channels"])
    with padding_mode("causal"):
        enc = ns.encoder.Encoder1D(**cfg)
    perturb_(enc, seed + 1)
    return enc.eval()


def build_ecapa(name: str, seed: int = 0, **overrides):
    ns = install()
    torch.manual_seed(seed)
    cfg = dict(attention_channels=128, dilations=[1, 1, 1, 1], global_context=True,
               groups=[1, 1, 1, 1], kernel_sizes=[3, 3, 3, 3], out_dim=6, pooling=True,
               res2net_scale=8, se_channels=128, spherical_normalisation=False, use_tanh=False,
               regularisation="ac")
    cfg.update(ECAPA_CFG[name])
    cfg.update(overrides)
    enc = ns.ecapa.ECAPATDNN(**cfg)
    perturb_(enc, seed + 1)
    return enc.eval()


def build_rectified_flow(name: str, seed: int = 0):
    ns = install()
    net = build_denoiser(name, seed)
    enc = build_ecapa(name, seed + 10)
    enc_t = build_encoder1d(name, seed + 20) if name in ENC1D_CFG else None
    rf = ns.model.RectifiedFlow(net=net, sr=44100, encoder=enc, encoder_time=enc_t,
                                classifier=None, time_transform=None, drop_value=-4.0,
                                drop_rate=0.2, device="cpu")
    return rf.eval()
