"""GPU parity of the STREAMING denoiser (per-diffusion-step rolling KV caches) through the
C ABI, against fixtures produced by the unmodified reference (tests/golden/make_golden_stream.py) and the CPU oracle.
Tolerances: fp32 modes 2e-4 relative L2 (north_star allows 1e-3), bf16 mode 5e-2."""
import numpy as np
import pytest
import torch

from after_b200 import config, synth

pytestmark = pytest.mark.gpu

TOL = {"fp32": 2e-4, "fp32_simt": 2e-4, "bf16": 5e-2}


def rel(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm())


def T(a):
    return torch.from_numpy(np.asarray(a))


def make_engine(name, wseed, precision, frames, cache, max_batch=1, max_steps=4):
    from after_b200.engine import Engine
    mc = config.get_config(name)
    sd = synth.denoiser_state_dict(mc.denoiser, wseed)
    return Engine(model=mc, denoiser_state=sd, precision=precision, max_batch=max_batch, max_steps=max_steps,
                  seq_len=frames, max_cache_size=cache), sd, mc


@pytest.mark.parametrize("precision", ["fp32_simt", "fp32", "bf16"])
@pytest.mark.parametrize("tag,name", [("tiny", "tiny"), ("tiny_t8", "tiny"), ("base", "base"), ("midi", "midi")])
def test_cached_forward_and_roll_match_reference(golden, tag, name, precision):
    """DenoiserV2.forward(cache_index) + roll_cache over consecutive blocks (transformerv2.py:167-236, 514-543)."""
    from after_b200.diffusion import DenoiserV2
    g = golden(f"stream_denoiser_{tag}")
    frames = g["x"].shape[-1]
    eng, _, _ = make_engine(name, int(g["weight_seed"]), precision, frames, int(g["cache_size"]))
    try:
        net = DenoiserV2(eng)
        worst = 0.0
        for b in range(g["x"].shape[0]):
            ci = int(g["cache_index"][b])
            n = g["x"].shape[1]
            out = net(T(g["x"][b]).cuda(), time=T(g["time"][b]).cuda().reshape(n, 1, 1), cond=T(g["cond"][b]).cuda(),
                      time_cond=T(g["time_cond"][b]).cuda(), cache_index=ci)
            net.roll_cache(int(g["roll"]), ci)
            worst = max(worst, rel(out, g["out"][b]))
        assert worst < TOL[precision], worst
        # reset_cache restores the initial (all-zero history) state: block 0 reproduces
        net.reset_cache()
        ci = int(g["cache_index"][0])
        n = g["x"].shape[1]
        out = net(T(g["x"][0]).cuda(), time=T(g["time"][0]).cuda().reshape(n, 1, 1), cond=T(g["cond"][0]).cuda(),
                  time_cond=T(g["time_cond"][0]).cuda(), cache_index=ci)
        assert rel(out, g["out"][0]) < TOL[precision]
    finally:
        eng.close()


@pytest.mark.parametrize("precision", ["fp32_simt", "fp32", "bf16"])
@pytest.mark.parametrize("name", ["tiny", "base"])
def test_sample_stream_matches_reference(golden, name, precision):
    """One call per audio block = the exported Streamer.sample loop (export.py:398-416), graph-replayed after block 0."""
    g = golden(f"stream_sample_{name}")
    frames = g["x0"].shape[-1]
    steps = int(g["nb_steps"])
    eng, _, _ = make_engine(name, int(g["weight_seed"]), precision, frames, int(g["cache_size"]), max_steps=steps)
    try:
        worst = 0.0
        for b in range(g["x0"].shape[0]):
            out = eng.sample_stream(T(g["x0"][b]).cuda(), T(g["cond"][b]).cuda(), T(g["time_cond"][b]).cuda(), steps,
                                    float(g["guidance_timbre"]), float(g["guidance_structure"]))
            worst = max(worst, rel(out, g["out"][b]))
        assert worst < TOL[precision], worst
    finally:
        eng.close()


def test_stepwise_model_forward_equals_fused_sample_stream(golden):
    """RectifiedFlow.model_forward(cache_index=i) + net.roll_cache, driven from Python exactly like export.py:398-416,
    gives bit-identical blocks to the single after_sample_stream call."""
    from after_b200.diffusion import DenoiserV2, RectifiedFlow
    g = golden("stream_sample_tiny")
    frames, steps = g["x0"].shape[-1], int(g["nb_steps"])
    g_t, g_s = float(g["guidance_timbre"]), float(g["guidance_structure"])
    eng_a, _, _ = make_engine("tiny", int(g["weight_seed"]), "fp32", frames, int(g["cache_size"]), max_steps=steps)
    eng_b, _, _ = make_engine("tiny", int(g["weight_seed"]), "fp32", frames, int(g["cache_size"]), max_steps=steps)
    try:
        rf = RectifiedFlow(net=DenoiserV2(eng_a), sr=44100, drop_value=-4.0, clamp=0.1)
        for b in range(3):
            x0, cond, tc = T(g["x0"][b]).cuda(), T(g["cond"][b]).cuda(), T(g["time_cond"][b]).cuda()
            x = x0
            t = torch.linspace(0, 1, steps + 1)
            for i, tv in enumerate(t[:-1]):
                x = x + rf.model_forward(x, tv.repeat(x.shape[0], 1, x.shape[-1]).cuda(), cond, tc, g_t, g_s,
                                         cache_index=i) * (1 / steps)
                rf.net.roll_cache(x.shape[-1], i)
            fused = eng_b.sample_stream(x0, cond, tc, steps, g_t, g_s)
            assert rel(x, g["out"][b]) < 2e-4
            assert rel(fused, x) < 1e-6
    finally:
        eng_a.close()
        eng_b.close()


def test_streaming_long_run_against_oracle():
    """40 consecutive 4-frame blocks, 2 steps, base config, B = 1 (3 CFG rows): histories roll many times over;
    every block is compared with the CPU oracle run on the same inputs."""
    from oracle import after_oracle as O
    mc = config.get_config("base")
    cfg = mc.denoiser
    eng, sd, _ = make_engine("base", 81, "fp32", 4, cfg.local_attention_size, max_steps=2)
    try:
        cache = O.StreamCache(cfg, cfg.local_attention_size)
        gen = torch.Generator().manual_seed(9)
        worst = 0.0
        for b in range(40):
            x0 = torch.randn(1, cfg.n_channels, 4, generator=gen)
            cond = torch.randn(1, cfg.cond_dim, generator=gen)
            tc = torch.randn(1, cfg.tcond_dim, 4, generator=gen)
            want = O.sample_stream(sd, cfg, cache, x0, cond, tc, 2, 2.0, 1.0)
            got = eng.sample_stream(x0.cuda(), cond.cuda(), tc.cuda(), 2, 2.0, 1.0)
            worst = max(worst, rel(got, want))
        assert worst < 2e-4, worst
    finally:
        eng.close()


def test_partial_roll_and_errors():
    """roll_size < block length keeps only the first frames (transformerv2.py:173-178); offline handles refuse cache calls."""
    from oracle import after_oracle as O
    from after_b200.diffusion import DenoiserV2
    cfg = config.get_config("tiny").denoiser
    eng, sd, _ = make_engine("tiny", 82, "fp32", 8, cfg.local_attention_size)
    try:
        net = DenoiserV2(eng)
        cache = O.StreamCache(cfg, cfg.local_attention_size)
        gen = torch.Generator().manual_seed(3)
        for b, roll in enumerate((4, 8, 3, 4)):
            x = torch.randn(3, cfg.n_channels, 8, generator=gen)
            t = torch.rand(3, generator=gen)
            cond = torch.randn(3, cfg.cond_dim, generator=gen)
            tc = torch.randn(3, cfg.tcond_dim, 8, generator=gen)
            want = O.denoiser_forward(sd, cfg, x, t, cond, tc, cache=cache, cache_index=1)
            cache.roll(roll, 1)
            got = net(x.cuda(), time=t.cuda(), cond=cond.cuda(), time_cond=tc.cuda(), cache_index=1)
            net.roll_cache(roll, 1)
            assert rel(got, want) < 2e-4, (b, rel(got, want))
        with pytest.raises(RuntimeError):
            eng.roll_cache(4, 99)
    finally:
        eng.close()
    from after_b200.engine import Engine
    mc = config.get_config("tiny")
    off = Engine(model=mc, denoiser_state=synth.denoiser_state_dict(mc.denoiser, 82), max_batch=1, max_steps=2, seq_len=8)
    try:
        with pytest.raises(RuntimeError):
            off.roll_cache(4, 0)
        with pytest.raises(RuntimeError):
            off.sample_stream(torch.zeros(1, 64, 4).cuda(), torch.zeros(1, 6).cuda(), torch.zeros(1, 12, 4).cuda(), 2)
    finally:
        off.close()


def test_streamer_runs_the_streaming_denoiser_over_consecutive_buffers():
    """nn_tilde-shaped ``Streamer.forward`` on an engine with ``max_cache_size = LOCAL_ATTENTION_SIZE``: three consecutive
    8192-sample buffers; the diffusion stage carries its per-step KV histories across calls exactly like the exported model
    (export.py:398-416) while codec / encoders run block-offline.  Checked against the oracle chain with injected noise."""
    from after_b200.engine import Engine
    from after_b200.streamer import Streamer
    from oracle import after_oracle as O
    mc = config.get_config("tiny")
    acfg = config.base_autoencoder()
    sds = dict(den=synth.denoiser_state_dict(mc.denoiser, 1), ae=synth.autoencoder_state_dict(acfg, 2),
               se=synth.encoder1d_state_dict(mc.structure_encoder, 3), te=synth.ecapa_state_dict(mc.timbre_encoder, 4))
    n_sig, frames, steps = 16, 4, 3
    eng = Engine(model=mc, autoencoder=acfg, denoiser_state=sds["den"], autoencoder_state=sds["ae"], structure_state=sds["se"],
                 timbre_state=sds["te"], precision="fp32", max_batch=2, max_steps=steps, seq_len=n_sig,
                 max_samples=n_sig * acfg.ratio, max_cache_size=mc.denoiser.local_attention_size)
    try:
        st = Streamer(eng, n_signal_timbre=n_sig, chunk_size=4)
        st.set_nb_steps(steps); st.set_guidance_timbre(2.0); st.set_guidance_structure(1.0)
        cache = O.StreamCache(mc.denoiser, mc.denoiser.local_attention_size)
        hist = torch.zeros(1, 64, n_sig)
        worst = 0.0
        for blk in range(3):
            audio = torch.cat([synth.synth_audio(1, frames * acfg.ratio, seed=40 + blk),
                               synth.synth_audio(1, frames * acfg.ratio, seed=50 + blk)], 1)
            noise = torch.randn(1, 64, frames, generator=torch.Generator().manual_seed(60 + blk))
            got = st.forward(audio.cuda(), noise=noise)
            z_s = O.ae_encode(sds["ae"], acfg, audio[:, :1])
            z_t = O.ae_encode(sds["ae"], acfg, audio[:, 1:])
            hist = torch.cat([hist, z_t], -1)[..., frames:]
            cond = O.ecapa_forward(sds["te"], mc.timbre_encoder, hist)
            tcond = O.encoder1d_forward(sds["se"], mc.structure_encoder, z_s)
            x = O.sample_stream(sds["den"], mc.denoiser, cache, noise, cond, tcond, steps, 2.0, 1.0, clamp=0.1)
            want = O.ae_decode(sds["ae"], acfg, x)
            worst = max(worst, rel(got, want))
        print(f"streaming streamer.forward over 3 buffers: {worst:.2e}")
        assert worst < 1e-3
    finally:
        eng.close()


def test_two_streams_stream_independently():
    """B = 2 streaming streams on one engine (6 CFG rows, own histories per row) give, stream by stream, exactly what two
    single-stream engines give -- the reference allows it too (max_batch_size = 4 rows, transformerv2.py:131)."""
    cfg = config.get_config("tiny").denoiser
    W = cfg.local_attention_size
    eng2, _, _ = make_engine("tiny", 83, "fp32", 4, W, max_batch=2, max_steps=2)
    engs = [make_engine("tiny", 83, "fp32", 4, W, max_batch=1, max_steps=2)[0] for _ in range(2)]
    try:
        gen = torch.Generator().manual_seed(17)
        for blk in range(4):
            x0 = torch.randn(2, cfg.n_channels, 4, generator=gen).cuda()
            cond = torch.randn(2, cfg.cond_dim, generator=gen).cuda()
            tc = torch.randn(2, cfg.tcond_dim, 4, generator=gen).cuda()
            both = eng2.sample_stream(x0, cond, tc, 2, 2.0, 1.0)
            for s in range(2):
                one = engs[s].sample_stream(x0[s:s + 1].contiguous(), cond[s:s + 1].contiguous(), tc[s:s + 1].contiguous(), 2, 2.0, 1.0)
                # 24 rows run on the tensor-core tiles (bf16x3), 12 rows on the exact-fp32 skinny linears: same values to
                # fp32-mode accuracy, not bitwise
                assert rel(both[s:s + 1], one) < 1e-4, (blk, s)
    finally:
        eng2.close()
        for e in engs:
            e.close()
