"""GPU parity of the CUDA codec (AutoEncoder.encode/decode) and structure encoder (Encoder1D) against the
reference fixtures and the CPU oracle, through the C ABI."""
import numpy as np
import pytest
import torch

from after_b200 import config, synth

pytestmark = pytest.mark.gpu

# bf16 (single-product) mode: one bf16 rounding (2^-9 relative) of every conv operand AND weight through 40 / 39 convs.
# These synthetic random weights amplify rounding ~65x (the fp32 3-product mode lands at 1.3e-4 .. 3e-4 from 4.5e-6 per
# conv), so single-product bf16 measures 6.5e-2 (encode) / 1.4e-1 (decode) here against the 3e-2 SURVEY.md section 4.1
# hoped for: the bf16 codec is gated at 2e-1 (1.5x the measured worst case) and is NOT the mode the parity claim of
# BASELINE configs[1] / [4] rests on -- that is fp32 mode, gated at 1e-3 at full size below.
TOL = {"fp32": 1e-3, "fp32_simt": 1e-3, "bf16": 2e-1}


def rel(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm())


def T(a):
    return torch.from_numpy(np.asarray(a))


def codec_engine(tag, wseed, precision, max_batch, max_samples):
    from after_b200.engine import Engine
    acfg = config.small_autoencoder() if tag == "small" else config.base_autoencoder()
    sd = synth.autoencoder_state_dict(acfg, wseed)
    return Engine(autoencoder=acfg, autoencoder_state=sd, precision=precision, max_batch=max_batch,
                  max_samples=max_samples), sd, acfg


@pytest.mark.parametrize("precision", ["fp32_simt", "fp32", "bf16"])
@pytest.mark.parametrize("tag", ["small", "base"])
def test_codec_matches_reference(golden, tag, precision):
    from after_b200.autoencoder import AutoEncoder
    g = golden(f"codec_{tag}")
    audio = T(g["audio"])
    eng, _, acfg = codec_engine(tag, int(g["weight_seed"]), precision, audio.shape[0], audio.shape[-1])
    try:
        ae = AutoEncoder(eng)
        assert ae.ratio == acfg.ratio
        z = ae.encode(audio.cuda())
        assert z.shape == g["z"].shape
        ez = rel(z, g["z"])
        y = ae.decode(T(g["z_in"]).cuda())
        assert y.shape == g["decoded"].shape
        ey = rel(y, g["decoded"])
        rec = ae.decode(T(g["z"]).cuda())
        er = rel(rec, g["reconstructed"])
        print(f"codec_{tag} {precision}: encode {ez:.2e} decode {ey:.2e} reconstruct {er:.2e}")
        assert ez < TOL[precision] and ey < TOL[precision] and er < TOL[precision]
        # replay of the captured graph: same kernels, but the GroupNorm sums are fp64 atomics whose order is not fixed,
        # so equality is asserted to rounding, not bitwise
        assert rel(ae.decode(T(g["z_in"]).cuda()), y) < 1e-6
    finally:
        eng.close()


@pytest.mark.parametrize("precision", ["fp32_simt", "fp32"])
def test_codec_ragged_and_batched_vs_oracle(precision):
    """Length that is not a multiple of any kernel tile (ratio * 5 frames), B = 3, oracle as checker."""
    from oracle import after_oracle as O
    acfg = config.base_autoencoder()
    samples = acfg.ratio * 5
    eng, sd, _ = codec_engine("base", 5, precision, 3, samples)
    try:
        audio = synth.synth_audio(3, samples, seed=11)
        z_ref = O.ae_encode(sd, acfg, audio)
        z = eng.ae_encode(audio.cuda())
        assert rel(z, z_ref) < 1e-3
        y_ref = O.ae_decode(sd, acfg, z_ref)
        y = eng.ae_decode(z_ref.cuda())
        assert rel(y, y_ref) < 1e-3
        # streams are independent: stream 1 alone gives the same audio
        y1 = eng.ae_decode(z_ref[1:2].cuda())
        assert rel(y1, y[1:2]) < 1e-5
    finally:
        eng.close()


@pytest.mark.parametrize("precision", ["fp32_simt", "fp32", "bf16"])
@pytest.mark.parametrize("name", ["tiny", "base"])
def test_structure_encoder_matches_reference(golden, name, precision):
    from after_b200.engine import Engine
    from after_b200.diffusion import Encoder1D
    g = golden(f"encoder1d_{name}")
    mc = config.get_config(name)
    sd = synth.encoder1d_state_dict(mc.structure_encoder, int(g["weight_seed"]))
    z = T(g["z"])
    eng = Engine(model=mc, structure_state=sd, precision=precision, max_batch=z.shape[0], seq_len=z.shape[-1])
    try:
        out = Encoder1D(eng)(z.cuda())
        assert out.shape == g["out"].shape
        e = rel(out, g["out"])
        print(f"encoder1d_{name} {precision}: {e:.2e}")
        assert e < (5e-2 if precision == "bf16" else 2e-4)
    finally:
        eng.close()


@pytest.mark.parametrize("name", ["tiny", "base"])
def test_timbre_encoder_matches_reference(golden, name):
    from after_b200.engine import Engine
    from after_b200.diffusion import ECAPATDNN
    from oracle import after_oracle as O
    g = golden(f"ecapa_{name}")
    mc = config.get_config(name)
    sd = synth.ecapa_state_dict(mc.timbre_encoder, int(g["weight_seed"]))
    z = T(g["z"])
    eng = Engine(model=mc, timbre_state=sd, precision="fp32", max_batch=4, seq_len=256)
    try:
        enc = ECAPATDNN(eng)
        out = enc(z.cuda())
        assert out.shape == g["out"].shape
        e = rel(out, g["out"])
        print(f"ecapa_{name}: {e:.2e}")
        assert e < 1e-4
        # full-length chunk (T = 256), B = 3, oracle as checker
        z2 = torch.randn(3, 64, 256, generator=torch.Generator().manual_seed(9))
        assert rel(enc(z2.cuda()), O.ecapa_forward(sd, mc.timbre_encoder, z2)) < 1e-4
    finally:
        eng.close()


def test_generate_chain_matches_oracle():
    """after_generate / after_generate_host (2x encode -> Encoder1D + ECAPA -> sample -> decode) vs the oracle chain."""
    from after_b200.engine import Engine
    from oracle import after_oracle as O
    mc = config.get_config("tiny")
    acfg = config.base_autoencoder()
    sds = dict(den=synth.denoiser_state_dict(mc.denoiser, 1), ae=synth.autoencoder_state_dict(acfg, 2),
               se=synth.encoder1d_state_dict(mc.structure_encoder, 3), te=synth.ecapa_state_dict(mc.timbre_encoder, 4))
    B, frames, steps = 2, 8, 3
    S = frames * acfg.ratio
    eng = Engine(model=mc, autoencoder=acfg, denoiser_state=sds["den"], autoencoder_state=sds["ae"], structure_state=sds["se"],
                 timbre_state=sds["te"], precision="fp32", max_batch=B, max_steps=steps, seq_len=frames, max_samples=S)
    try:
        a_s, a_t = synth.synth_audio(B, S, seed=21), synth.synth_audio(B, S, seed=22)
        x0 = torch.randn(B, 64, frames, generator=torch.Generator().manual_seed(23))
        z_s, z_t = O.ae_encode(sds["ae"], acfg, a_s), O.ae_encode(sds["ae"], acfg, a_t)
        tcond = O.encoder1d_forward(sds["se"], mc.structure_encoder, z_s)
        cond = O.ecapa_forward(sds["te"], mc.timbre_encoder, z_t)
        x = O.sample(sds["den"], mc.denoiser, x0, cond, tcond, steps, 2.0, 1.0)
        want = O.ae_decode(sds["ae"], acfg, x)
        got = eng.generate(a_s.cuda(), a_t.cuda(), x0.cuda(), steps, 2.0, 1.0)
        assert got.shape == want.shape
        e = rel(got, want)
        print(f"generate chain: {e:.2e}")
        assert e < 1e-3
        host_out = torch.empty(B, 1, S)
        eng.generate_host(a_s, a_t, x0, host_out, steps, 2.0, 1.0)
        assert torch.equal(host_out, got.cpu())
    finally:
        eng.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_codec_full_chunk_matches_oracle(precision):
    """BASELINE chunk size: B = 2 x 524288 samples -> z (2, 64, 256) -> 524288 samples, VALUES against the CPU oracle
    (AutoEncoder.encode / decode, SimpleNetsStream.py:918-954): GroupNorm over T = 32768 with cross-CTA fp64 atomics, the
    128-frame TMA tile tails and the 256-frame PQMF at the size the RTF numbers are quoted on.  Also the reference's own
    self-check (export_autoencoder.py:50-54): encode -> decode keeps the length."""
    from oracle import after_oracle as O
    eng, sd, acfg = codec_engine("base", 2, precision, 2, 524288)
    try:
        audio = synth.synth_audio(2, 524288, seed=3)
        z_ref = O.ae_encode(sd, acfg, audio)
        z = eng.ae_encode(audio.cuda())
        assert z.shape == (2, 64, 256) and torch.isfinite(z).all()
        ez = rel(z, z_ref)
        zin = torch.randn(2, 64, 256, generator=torch.Generator().manual_seed(17))
        y_ref = O.ae_decode(sd, acfg, zin)
        y = eng.ae_decode(zin.cuda())
        assert y.shape == audio.shape and torch.isfinite(y).all()
        ey = rel(y, y_ref)
        # encode -> decode round trip on the oracle's own latents (what the chain feeds the decoder)
        er = rel(eng.ae_decode(z_ref.cuda()), O.ae_decode(sd, acfg, z_ref))
        print(f"codec 2x524288 {precision}: encode {ez:.2e} decode {ey:.2e} decode(encode) {er:.2e}")
        assert ez < TOL[precision] and ey < TOL[precision] and er < TOL[precision]
    finally:
        eng.close()


def test_generate_chain_full_size_matches_oracle():
    """The whole audio-to-audio chain at BASELINE configs[1] arithmetic: base models, 524288-sample chunks (T = 256), 50
    Euler steps, CFG 2.0 / 1.0, B = 2 -- after_generate vs the oracle chain, on latents and on audio (north_star: 1e-3)."""
    from after_b200.engine import Engine
    from oracle import after_oracle as O
    mc = config.get_config("base")
    acfg = config.base_autoencoder()
    sds = dict(den=synth.denoiser_state_dict(mc.denoiser, 0), ae=synth.autoencoder_state_dict(acfg, 0),
               se=synth.encoder1d_state_dict(mc.structure_encoder, 0), te=synth.ecapa_state_dict(mc.timbre_encoder, 0))
    B, steps, S = 2, 50, 524288
    frames = S // acfg.ratio
    eng = Engine(model=mc, autoencoder=acfg, denoiser_state=sds["den"], autoencoder_state=sds["ae"], structure_state=sds["se"],
                 timbre_state=sds["te"], precision="fp32", max_batch=B, max_steps=steps, seq_len=frames, max_samples=S)
    try:
        a_s, a_t = synth.synth_audio(B, S, seed=41), synth.synth_audio(B, S, seed=42)
        x0 = torch.randn(B, 64, frames, generator=torch.Generator().manual_seed(43))
        z_s, z_t = O.ae_encode(sds["ae"], acfg, a_s), O.ae_encode(sds["ae"], acfg, a_t)
        tcond = O.encoder1d_forward(sds["se"], mc.structure_encoder, z_s)
        cond = O.ecapa_forward(sds["te"], mc.timbre_encoder, z_t)
        x = O.sample(sds["den"], mc.denoiser, x0, cond, tcond, steps, 2.0, 1.0)
        want = O.ae_decode(sds["ae"], acfg, x)
        # stage by stage on the GPU (same calls the chain makes), then the one-call chain
        g_zs, g_zt = eng.ae_encode(a_s.cuda()), eng.ae_encode(a_t.cuda())
        g_tc, g_c = eng.structure_encode(g_zs), eng.timbre_encode(g_zt)
        g_x = eng.sample(x0.cuda(), g_c, g_tc, steps, 2.0, 1.0)
        e_lat = rel(g_x, x)
        got = eng.generate(a_s.cuda(), a_t.cuda(), x0.cuda(), steps, 2.0, 1.0)
        e_audio = rel(got, want)
        print(f"chain base T=256 50 steps: z {rel(g_zs, z_s):.2e} time_cond {rel(g_tc, tcond):.2e} cond {rel(g_c, cond):.2e} "
              f"latents {e_lat:.2e} audio {e_audio:.2e}")
        assert e_lat < 1e-3 and e_audio < 1e-3
        assert rel(eng.ae_decode(g_x), got) < 1e-5  # the one-call chain is the same computation
    finally:
        eng.close()


def test_codec_single_frame():
    """Shortest legal input: one latent frame (samples = ratio)."""
    from oracle import after_oracle as O
    acfg = config.base_autoencoder()
    eng, sd, _ = codec_engine("base", 6, "fp32", 1, acfg.ratio)
    try:
        audio = synth.synth_audio(1, acfg.ratio, seed=13)
        z_ref = O.ae_encode(sd, acfg, audio)
        assert rel(eng.ae_encode(audio.cuda()), z_ref) < 1e-3
        assert rel(eng.ae_decode(z_ref.cuda()), O.ae_decode(sd, acfg, z_ref)) < 1e-3
        with pytest.raises(ValueError):
            eng.ae_encode(audio[..., :-1].cuda())  # not a multiple of the codec ratio
    finally:
        eng.close()


def test_streamer_surface_matches_oracle_chain():
    """nn_tilde-shaped ``Streamer`` (export.py:145-507): method shapes, attribute setters, rolling timbre buffer and
    ``forward`` = structure | timbre -> diffuse -> decode, checked against the oracle with the same injected noise."""
    from after_b200.engine import Engine
    from after_b200.streamer import Streamer
    from oracle import after_oracle as O
    mc = config.get_config("tiny")
    acfg = config.base_autoencoder()
    sds = dict(den=synth.denoiser_state_dict(mc.denoiser, 1), ae=synth.autoencoder_state_dict(acfg, 2),
               se=synth.encoder1d_state_dict(mc.structure_encoder, 3), te=synth.ecapa_state_dict(mc.timbre_encoder, 4))
    n_sig = 16
    eng = Engine(model=mc, autoencoder=acfg, denoiser_state=sds["den"], autoencoder_state=sds["ae"], structure_state=sds["se"],
                 timbre_state=sds["te"], precision="fp32", max_batch=4, max_steps=8, seq_len=n_sig, max_samples=n_sig * acfg.ratio)
    try:
        st = Streamer(eng, n_signal_timbre=n_sig, chunk_size=4)
        assert st.set_nb_steps(3) == 0 and st.get_nb_steps() == 3
        st.set_guidance_timbre(2.0); st.set_guidance_structure(1.0)
        assert st.get_guidance_timbre() == 2.0 and st.methods["diffuse"] == (12 + 6, acfg.ratio, 64, acfg.ratio)
        frames = 4
        audio = torch.cat([synth.synth_audio(2, frames * acfg.ratio, seed=31), synth.synth_audio(2, frames * acfg.ratio, seed=32)], 1)
        noise = torch.randn(2, 64, frames, generator=torch.Generator().manual_seed(33))
        got = st.forward(audio.cuda(), noise=noise)
        assert got.shape == (2, 1, frames * acfg.ratio)
        # oracle chain: timbre history = zeros rolled by the new latents; only row 0 is diffused, then repeated
        z_s = O.ae_encode(sds["ae"], acfg, audio[:, :1])
        z_t = O.ae_encode(sds["ae"], acfg, audio[:, 1:])
        hist = torch.cat([torch.zeros(2, 64, n_sig), z_t], -1)[..., frames:]
        cond = O.ecapa_forward(sds["te"], mc.timbre_encoder, hist)
        tcond = O.encoder1d_forward(sds["se"], mc.structure_encoder, z_s)
        x = O.sample(sds["den"], mc.denoiser, noise[:1], cond[:1], tcond[:1], 3, 2.0, 1.0, clamp=0.1)
        want = O.ae_decode(sds["ae"], acfg, x).repeat(2, 1, 1)
        e = rel(got, want)
        print(f"streamer.forward: {e:.2e}")
        assert e < 1e-3
        assert torch.equal(st.previous_timbre[:2].cpu()[..., -frames:], eng.ae_encode(audio[:, 1:].cuda()).cpu())
        with pytest.raises(ValueError):
            st.decode(torch.zeros(1, 3, 4))
    finally:
        eng.close()
