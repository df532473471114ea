"""Pin the CPU oracle (oracle/after_oracle.py) against fixtures produced by the UNMODIFIED
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from after_b200 import config, synth
from oracle import after_oracle as O


def rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / b.norm())


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("name", ["tiny", "base", "midi"])
def test_denoiser_forward_matches_reference(golden, name):
    g = golden(f"denoiser_{name}")
    cfg = config.get_config(name).denoiser
    sd = synth.denoiser_state_dict(cfg, int(g["weight_seed"]))
    taps = {}
    out = O.denoiser_forward(sd, cfg, T(g["x"]), T(g["time"]), T(g["cond"]), T(g["time_cond"]), taps)
    assert rel(taps["h1"], g["h1"]) < 2e-6
    assert rel(taps["h6"], g["h6"]) < 5e-6
    assert rel(out, g["out"]) < 5e-6
    # fp64 evaluation of the oracle: the fp32 reference sits within fp32 rounding noise of it
    out64 = O.denoiser_forward(sd, cfg, T(g["x"]).double(), T(g["time"]).double(),
                               T(g["cond"]).double(), T(g["time_cond"]).double())
    assert rel(g["out"], out64) < 2e-5


def test_band_mask_matches_reference(golden):
    g = golden("band_mask")
    for key, (L, w) in {"w8": (64, 8), "w16": (64, 16), "w8_len30": (30, 8)}.items():
        allowed = O.band_allowed(L, 4, w)
        assert torch.equal(~allowed, T(g[key]).bool())
    counts = O.band_allowed(64, 4, 8).sum(1).tolist()
    assert counts[:16] == [4, 4, 4, 4, 8, 8, 8, 8, 11, 10, 9, 8, 11, 10, 9, 8]


@pytest.mark.parametrize("name", ["tiny", "base"])
def test_model_forward_and_sample_match_reference(golden, name):
    g = golden(f"sample_{name}")
    cfg = config.get_config(name).denoiser
    sd = synth.denoiser_state_dict(cfg, int(g["weight_seed"]))
    x0, cond, tc = T(g["x0"]), T(g["cond"]), T(g["time_cond"])
    B = x0.shape[0]
    dx = O.model_forward(sd, cfg, x0, torch.full((B, ), float(g["t_model_forward"])), cond, tc,
                         float(g["guidance_timbre"]), float(g["guidance_structure"]))
    assert rel(dx, g["dx"]) < 1e-5
    out = O.sample(sd, cfg, x0, cond, tc, int(g["nb_steps"]), float(g["guidance_timbre"]),
                   float(g["guidance_structure"]))
    assert rel(out, g["out"]) < 1e-5


@pytest.mark.parametrize("tag", ["small", "base"])
def test_codec_matches_reference(golden, tag):
    g = golden(f"codec_{tag}")
    acfg = config.small_autoencoder() if tag == "small" else config.base_autoencoder()
    sd = synth.autoencoder_state_dict(acfg, int(g["weight_seed"]))
    audio = T(g["audio"])
    assert rel(O.pqmf_analysis(sd, audio), g["multiband"]) < 1e-6
    z = O.ae_encode(sd, acfg, audio)
    assert z.shape == g["z"].shape
    assert rel(z, g["z"]) < 2e-5
    y = O.ae_decode(sd, acfg, T(g["z_in"]))
    assert y.shape == g["decoded"].shape
    assert rel(y, g["decoded"]) < 2e-5
    assert rel(O.ae_decode(sd, acfg, T(g["z"])), g["reconstructed"]) < 2e-5


def test_pqmf_design_matches_reference(golden):
    g = golden("pqmf_16band_100dB")
    sd = synth.pqmf_filters(100, 16)
    np.testing.assert_allclose(sd["pqmf.h"].numpy(), g["h"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(sd["pqmf.forward_conv.weight"].numpy(), g["forward"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(sd["pqmf.inverse_conv.weight"].numpy(), g["inverse"], rtol=0, atol=1e-7)


@pytest.mark.parametrize("name", ["tiny", "base"])
def test_encoder1d_matches_reference(golden, name):
    g = golden(f"encoder1d_{name}")
    ecfg = config.get_config(name).structure_encoder
    sd = synth.encoder1d_state_dict(ecfg, int(g["weight_seed"]))
    out = O.encoder1d_forward(sd, ecfg, T(g["z"]))
    assert rel(out, g["out"]) < 1e-5


@pytest.mark.parametrize("name", ["tiny", "base"])
def test_ecapa_matches_reference(golden, name):
    g = golden(f"ecapa_{name}")
    ecfg = config.get_config(name).timbre_encoder
    sd = synth.ecapa_state_dict(ecfg, int(g["weight_seed"]))
    out = O.ecapa_forward(sd, ecfg, T(g["z"]))
    assert out.shape == g["out"].shape
    assert rel(out, g["out"]) < 1e-5


def test_codec_roundtrip_length():
    """The reference's own shape self-check (export_autoencoder.py:50-54): encode->decode keeps
    the length; 65536 samples -> z (1,64,32) -> 65536."""
    acfg = config.small_autoencoder()
    sd = synth.autoencoder_state_dict(acfg, 0)
    x = torch.zeros(3, 1, 8192)
    z = O.ae_encode(sd, acfg, x)
    assert z.shape == (3, acfg.z_channels, 8192 // acfg.ratio)
    assert O.ae_decode(sd, acfg, z).shape == x.shape


# ---------------------------------------------------------------------------------------------------------------
# streaming path (per-diffusion-step rolling KV caches), fixtures from tests/golden/make_golden_stream.py
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,name", [("tiny", "tiny"), ("tiny_t8", "tiny"), ("base", "base"), ("midi", "midi")])
def test_stream_denoiser_matches_reference(golden, tag, name):
    g = golden(f"stream_denoiser_{tag}")
    cfg = config.get_config(name).denoiser
    sd = synth.denoiser_state_dict(cfg, int(g["weight_seed"]))
    cache = O.StreamCache(cfg, int(g["cache_size"]))
    worst = 0.0
    for b in range(g["x"].shape[0]):
        ci = int(g["cache_index"][b])
        out = O.denoiser_forward(sd, cfg, T(g["x"][b]), T(g["time"][b]), T(g["cond"][b]), T(g["time_cond"][b]),
                                 cache=cache, cache_index=ci)
        cache.roll(int(g["roll"]), ci)
        worst = max(worst, rel(out, g["out"][b]))
    assert worst < 1e-5, worst
    # the history matters: the offline forward of the last block differs from its streaming output
    off = O.denoiser_forward(sd, cfg, T(g["x"][-1]), T(g["time"][-1]), T(g["cond"][-1]), T(g["time_cond"][-1]))
    assert rel(off, g["out"][-1]) > 1e-3


@pytest.mark.parametrize("name", ["tiny", "base"])
def test_stream_sample_matches_reference(golden, name):
    g = golden(f"stream_sample_{name}")
    cfg = config.get_config(name).denoiser
    sd = synth.denoiser_state_dict(cfg, int(g["weight_seed"]))
    cache = O.StreamCache(cfg, int(g["cache_size"]))
    for b in range(g["x0"].shape[0]):
        out = O.sample_stream(sd, cfg, cache, T(g["x0"][b]), T(g["cond"][b]), T(g["time_cond"][b]), int(g["nb_steps"]),
                              float(g["guidance_timbre"]), float(g["guidance_structure"]))
        assert rel(out, g["out"][b]) < 2e-5, (b, rel(out, g["out"][b]))


# ---------------------------------------------------------------------------------------------------------------
# UNET1D conv denoiser: oracle only so far, fixtures from tests/golden/make_golden_unet.py
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["attn", "concat"])
def test_unet1d_matches_reference(golden, tag):
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from unet_cases import CASES
    cfg, wseed = CASES[tag]
    g = golden(f"unet_{tag}")
    assert int(g["weight_seed"]) == wseed
    sd = synth.unet_state_dict(cfg, wseed)
    cond = T(g["cond"]) if cfg.cond_channels else None
    out = O.unet1d_forward(sd, cfg, T(g["x"]), T(g["time"]), cond, T(g["time_cond"]))
    assert out.shape == g["out"].shape
    assert rel(out, g["out"]) < 1e-5, rel(out, g["out"])
    out64 = O.unet1d_forward(sd, cfg, T(g["x"]).double(), T(g["time"]).double(), cond.double() if cond is not None else None,
                             T(g["time_cond"]).double())
    assert rel(g["out"], out64) < 2e-5


def test_latent_map_matches_reference(golden):
    """Streamer.latent2map / map2latent: the oracle restatement against the reference SmallAutoencoder's outputs."""
    g = golden("latent_map")
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd.")}
    l2m = O.latent_map(sd, torch.from_numpy(g["latents"]), 0)
    m2l = O.latent_map(sd, torch.from_numpy(g["maps"]), 1)
    assert torch.allclose(l2m, torch.from_numpy(g["latent2map"]), atol=1e-6, rtol=1e-5)
    assert torch.allclose(m2l, torch.from_numpy(g["map2latent"]), atol=1e-6, rtol=1e-5)
    ident = O.latent_map(None, torch.from_numpy(g["maps"]), 0)
    assert torch.allclose(ident, torch.from_numpy(g["maps"]).mean(-1, keepdim=True).expand_as(ident))
