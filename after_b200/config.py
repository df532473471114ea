"""Hyper-parameters of the shipped AFTER configurations, as plain dataclasses.

The reference binds these through gin files; the numbers below restate
``after/diffusion/configs/{tiny,base,midi}.gin`` and ``after/autoencoder/configs/baseAE.gin``
(reference line numbers in the field comments).  ``IN_SIZE`` = 64 latent channels and
``N_SIGNAL`` = 256 frames correspond to one 524288-sample chunk through the baseAE codec.
"""
from dataclasses import dataclass, field, asdict
from typing import List, Optional


@dataclass
class DenoiserConfig:
    """``DenoiserV2`` constructor arguments (transformerv2.py:463-476)."""
    n_channels: int = 64            # base.gin:66  (IN_SIZE)
    seq_len: int = 256              # base.gin:67  (N_SIGNAL)
    embed_dim: int = 512            # base.gin:68  (tiny.gin:68 -> 256)
    cond_dim: int = 6               # base.gin:69  (ZT_CHANNELS)
    noise_embed_dims: int = 64      # base.gin:70
    n_layers: int = 6               # base.gin:71
    mlp_multiplier: int = 3         # base.gin:72
    tcond_dim: int = 12             # base.gin:75  (midi.gin:13 -> 128)
    local_attention_size: int = 8   # base.gin:18  (midi.gin:20 -> 16)
    attention_chunk_size: int = 4   # base.gin:19
    head_dim: int = 64              # transformerv2.py:320  n_heads = embed_dim // 64
    rotary_dim: int = 32            # transformerv2.py:406  RotaryEmbedding(32)
    rotary_theta: float = 10000.0   # rotary_embedding.py:44
    fourier_factor: float = 100.0   # transformerv2.py:485
    fourier_max_positions: float = 10000.0  # transformerv2.py:484

    @property
    def n_heads(self) -> int:
        return self.embed_dim // self.head_dim


@dataclass
class AutoEncoderConfig:
    """``AutoEncoder`` constructor arguments (baseAE.gin:35-52, SimpleNetsStream.py:834-849)."""
    in_channels: int = 16
    channels: int = 64
    z_channels: int = 64
    pqmf_bands: int = 16
    pqmf_attenuation: int = 100     # SimpleNetsStream.py:855
    multipliers: List[int] = field(default_factory=lambda: [1, 2, 4, 4, 8, 8])
    factors: List[int] = field(default_factory=lambda: [2, 2, 2, 4, 4])
    dilations: List[int] = field(default_factory=lambda: [1, 3, 9])
    kernel_size: int = 3
    resnet_groups: int = 8
    decoder_ratio: float = 1.5
    use_loudness: bool = True
    num_blocks: int = 3             # SimpleNetsStream.py:861

    @property
    def ratio(self) -> int:
        r = self.pqmf_bands
        for f in self.factors:
            r *= f
        return r

    @property
    def decoder_multipliers(self) -> List[int]:
        return [int(m * self.decoder_ratio) for m in self.multipliers[::-1]]


@dataclass
class Encoder1DConfig:
    """Structure encoder ``Encoder1D`` (base.gin:44-55; causal padding base.gin:55)."""
    in_size: int = 64
    channels: List[int] = field(default_factory=lambda: [64, 128, 256, 512, 12])
    ratios: List[int] = field(default_factory=lambda: [1, 1, 1, 1])
    kernel_size: int = 5
    causal: bool = True
    use_tanh: bool = False


@dataclass
class EcapaConfig:
    """Timbre encoder ``ECAPATDNN`` (base.gin:27-41)."""
    in_size: int = 64
    channels: List[int] = field(default_factory=lambda: [512, 512, 512, 1024])
    kernel_sizes: List[int] = field(default_factory=lambda: [3, 3, 3, 3])
    dilations: List[int] = field(default_factory=lambda: [1, 1, 1, 1])
    attention_channels: int = 128
    res2net_scale: int = 8
    se_channels: int = 128
    out_dim: int = 6
    global_context: bool = True
    use_tanh: bool = False


@dataclass
class UNetConfig:
    """``UNET1D`` constructor arguments (after/diffusion/networks/unet1d.py:255-268) -- the conv denoiser no shipped config
    binds.  Only the oracle covers it so far."""
    in_size: int = 64
    out_size: Optional[int] = None
    channels: List[int] = field(default_factory=lambda: [128, 128, 256, 256])
    ratios: List[int] = field(default_factory=lambda: [2, 2, 2])  # len(channels) - 1 entries are used (a leading 1 is added)
    kernel_size: int = 5
    time_channels: int = 64
    time_cond_in_channels: int = 12
    time_cond_channels: int = 64
    cond_channels: int = 6
    n_attn_layers: int = 0
    use_res_last: bool = False


@dataclass
class ModelConfig:
    name: str
    denoiser: DenoiserConfig
    structure_encoder: Optional[Encoder1DConfig]
    timbre_encoder: EcapaConfig
    drop_value: float = -4.0        # base.gin:88
    sr: int = 44100


def tiny() -> ModelConfig:
    return ModelConfig(
        "tiny", DenoiserConfig(embed_dim=256),
        Encoder1DConfig(channels=[64, 128, 256, 256, 12]),
        EcapaConfig(channels=[256, 256, 256, 512]))


def base() -> ModelConfig:
    return ModelConfig("base", DenoiserConfig(), Encoder1DConfig(), EcapaConfig())


def midi() -> ModelConfig:
    return ModelConfig("midi", DenoiserConfig(tcond_dim=128, local_attention_size=16), None,
                       EcapaConfig())


def get_config(name: str) -> ModelConfig:
    return {"tiny": tiny, "base": base, "midi": midi}[name]()


def base_autoencoder() -> AutoEncoderConfig:
    return AutoEncoderConfig()


def small_autoencoder() -> AutoEncoderConfig:
    """A reduced codec (same topology, fewer channels/stages) for fast parity tests."""
    return AutoEncoderConfig(channels=16, z_channels=8, multipliers=[1, 2, 2], factors=[2, 4],
                             dilations=[1, 3, 9], decoder_ratio=1.5)


__all__ = [
    "DenoiserConfig", "AutoEncoderConfig", "Encoder1DConfig", "EcapaConfig", "ModelConfig", "UNetConfig",
    "tiny", "base", "midi", "get_config", "base_autoencoder", "small_autoencoder", "asdict"
]
