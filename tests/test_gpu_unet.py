"""GPU parity of the UNET1D conv denoiser (unet1d.py:30-429, blocks.py:201-243) through the C ABI:
reference fixtures, the CPU oracle at a tensor-core-sized configuration, and RectifiedFlow.sample over it."""
import os
import sys

import numpy as np
import pytest
import torch

from after_b200 import config, synth

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

pytestmark = pytest.mark.gpu

TOL = {"fp32_simt": 2e-4, "fp32": 2e-4, "bf16": 5e-2}


def rel(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm())


def T(a):
    return torch.from_numpy(np.asarray(a))


def unet_engine(cfg, wseed, precision, max_batch, frames, max_steps=4):
    from after_b200.engine import Engine
    sd = synth.unet_state_dict(cfg, wseed)
    return Engine(unet=cfg, unet_state=sd, precision=precision, max_batch=max_batch, max_steps=max_steps, seq_len=frames), sd


@pytest.mark.parametrize("precision", ["fp32_simt", "fp32", "bf16"])
@pytest.mark.parametrize("tag", ["attn", "concat"])
def test_unet_forward_matches_reference(golden, tag, precision):
    """Fixtures minted from the unmodified reference UNET1D (tests/golden/make_golden_unet.py): per-scale time_cond path with
    SelfAttention1d, and the concat path with a ratio-1 stage and a residual last block."""
    from unet_cases import CASES
    from after_b200.diffusion import UNET1D
    cfg, wseed = CASES[tag]
    g = golden(f"unet_{tag}")
    x = T(g["x"])
    eng, _ = unet_engine(cfg, wseed, precision, x.shape[0], x.shape[-1])
    try:
        cond = T(g["cond"]).cuda() if cfg.cond_channels else None
        out = UNET1D(eng)(x.cuda(), time=T(g["time"]).cuda(), time_cond=T(g["time_cond"]).cuda(), cond=cond)
        assert out.shape == g["out"].shape
        e = rel(out, g["out"])
        print(f"unet_{tag} {precision}: {e:.2e}")
        assert e < TOL[precision]
    finally:
        eng.close()


BIG = config.UNetConfig(in_size=64, channels=[128, 128, 256, 256], ratios=[2, 2, 2], kernel_size=5, time_channels=64,
                        time_cond_in_channels=12, time_cond_channels=64, cond_channels=6, n_attn_layers=2)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_unet_latent_sized_forward_matches_oracle(precision):
    """A configuration at the sizes the AFTER latents have (64 channels, T = 256, tcgen05-sized layers: 320-channel
    concatenations, strided pools, folded upsample convs, attention at two levels), N = 3, oracle as checker."""
    from oracle import after_oracle as O
    N, frames = 3, 256
    eng, sd = unet_engine(BIG, 11, precision, N, frames)
    try:
        g = torch.Generator().manual_seed(5)
        x = torch.randn(N, 64, frames, generator=g)
        t = torch.rand(N, generator=g)
        cond = torch.randn(N, 6, generator=g)
        tc = torch.randn(N, 12, frames, generator=g)
        want = O.unet1d_forward(sd, BIG, x, t, cond, tc)
        got = eng.unet_forward(x.cuda(), t.cuda(), cond.cuda(), tc.cuda())
        e = rel(got, want)
        print(f"unet latent-sized {precision}: {e:.2e}")
        assert e < TOL[precision]
        # ragged (still divisible by the ratios) shorter input through the same engine
        x2, tc2 = x[:2, :, :72].contiguous(), tc[:2, :, :72].contiguous()
        e2 = rel(eng.unet_forward(x2.cuda(), t[:2].cuda(), cond[:2].cuda(), tc2.cuda()), O.unet1d_forward(sd, BIG, x2, t[:2], cond[:2], tc2))
        assert e2 < TOL[precision]
    finally:
        eng.close()


@pytest.mark.parametrize("variant,clamp", [(0, 0.01), (1, 0.1)])
def test_unet_sample_matches_oracle(variant, clamp):
    """RectifiedFlow.sample / model_forward (model.py:721-785) with UNET1D as the net: 3-way CFG + Euler, both layouts."""
    from oracle import after_oracle as O
    from after_b200.diffusion import UNET1D, RectifiedFlow
    cfg = config.UNetConfig(in_size=16, channels=[32, 64, 64], ratios=[2, 2], kernel_size=5, time_channels=32,
                            time_cond_in_channels=4, time_cond_channels=16, cond_channels=6, n_attn_layers=1)
    B, frames, steps = 2, 32, 4
    eng, sd = unet_engine(cfg, 21, "fp32", B, frames, max_steps=steps)
    try:
        g = torch.Generator().manual_seed(9)
        x0 = torch.randn(B, 16, frames, generator=g)
        cond = torch.randn(B, 6, generator=g)
        tc = torch.randn(B, 4, frames, generator=g)
        net = O.unet_net(sd, cfg)
        rf = RectifiedFlow(net=UNET1D(eng), sr=44100, drop_value=-4.0, cfg_variant=variant, clamp=clamp)
        t = torch.full((B, 1, 1), 0.3)
        want_dx = O.model_forward(sd, cfg, x0, t, cond, tc, 2.0, 1.5, cfg_variant=variant, clamp=clamp, net=net)
        got_dx = rf.model_forward(x0.cuda(), t.cuda(), cond.cuda(), tc.cuda(), 2.0, 1.5)
        assert rel(got_dx, want_dx) < 2e-4
        want = O.sample(sd, cfg, x0, cond, tc, steps, 2.0, 1.5, cfg_variant=variant, clamp=clamp, net=net)
        got = rf.sample(x0.cuda(), cond.cuda(), tc.cuda(), steps, 2.0, 1.5)
        e = rel(got, want)
        print(f"unet sample variant {variant}: {e:.2e}")
        assert e < 2e-4
        assert rel(rf.sample(x0.cuda(), cond.cuda(), tc.cuda(), steps, 2.0, 1.5), got) < 1e-6  # graph replay
    finally:
        eng.close()


def test_unet_errors_are_loud():
    cfg = config.UNetConfig(in_size=16, channels=[32, 64, 64], ratios=[2, 2], kernel_size=5, time_channels=32,
                            time_cond_in_channels=4, time_cond_channels=16, cond_channels=6, n_attn_layers=0)
    eng, sd = unet_engine(cfg, 3, "fp32", 2, 32)
    try:
        x = torch.randn(2, 16, 30).cuda()  # 30 is not a multiple of 4
        with pytest.raises(RuntimeError):
            eng.unet_forward(x, torch.rand(2).cuda(), torch.randn(2, 6).cuda(), torch.randn(2, 4, 30).cuda())
        with pytest.raises(ValueError):
            eng.unet_forward(torch.randn(2, 16, 32).cuda(), torch.rand(2).cuda(), None, torch.randn(2, 4, 32).cuda())
        with pytest.raises(RuntimeError):
            eng.denoiser_forward(torch.randn(2, 16, 32).cuda(), torch.rand(2).cuda(), torch.randn(2, 6).cuda(),
                                 torch.randn(2, 4, 32).cuda())  # no DenoiserV2 on this handle
    finally:
        eng.close()
