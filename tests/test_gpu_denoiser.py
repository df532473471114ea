"""GPU parity of the CUDA denoiser / sampler (through the C ABI) against the reference fixtures and
the CPU oracle.  Tolerances: fp32 modes 1e-3 relative L2 (north_star), in practice ~1e-5; bf16 mode 5e-2."""
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


@pytest.fixture(scope="module")
def blank_engine():
    from after_b200.engine import Engine
    eng = Engine()
    yield eng
    eng.close()


@pytest.mark.parametrize("precision,tol", [("fp32_simt", 2e-6), ("fp32", 2e-5), ("bf16", 1e-2)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 192, 512), (1000, 64, 1536), (77, 32, 128), (6144, 1536, 512)])
def test_gemm_matches_fp64(blank_engine, precision, tol, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K**0.5
    bias = torch.randn(N, generator=g)
    ref = A.double() @ W.double().T + bias.double()
    out = blank_engine.debug_gemm(A.cuda(), W.cuda(), bias.cuda(), precision)
    assert rel(out, ref) < tol


def make_engine(name, wseed, precision, frames, max_batch=8, max_steps=8):
    from after_b200.engine import Engine
    mc = config.get_config(name)
    sd = synth.denoiser_state_dict(mc.denoiser, wseed)
    return Engine(model=mc, denoiser_state=sd, precision=precision, max_batch=max_batch, max_steps=max_steps,
                  seq_len=frames), sd, mc


@pytest.mark.parametrize("precision", ["fp32_simt", "fp32", "bf16"])
@pytest.mark.parametrize("name", ["tiny", "base", "midi"])
def test_denoiser_forward_matches_reference(golden, name, precision):
    g = golden(f"denoiser_{name}")
    x = T(g["x"])
    eng, _, _ = make_engine(name, int(g["weight_seed"]), precision, x.shape[-1])
    try:
        out = eng.denoiser_forward(x.cuda(), T(g["time"]).cuda(), T(g["cond"]).cuda(), T(g["time_cond"]).cuda())
        assert out.shape == x.shape
        assert rel(out, g["out"]) < TOL[precision]
    finally:
        eng.close()


@pytest.mark.parametrize("precision", ["fp32_simt", "fp32", "bf16"])
@pytest.mark.parametrize("name", ["tiny", "base"])
def test_model_forward_and_sample_match_reference(golden, name, precision):
    from after_b200.diffusion import DenoiserV2, RectifiedFlow
    g = golden(f"sample_{name}")
    x0, cond, tc = T(g["x0"]).cuda(), T(g["cond"]).cuda(), T(g["time_cond"]).cuda()
    B = x0.shape[0]
    eng, _, _ = make_engine(name, int(g["weight_seed"]), precision, x0.shape[-1])
    try:
        rf = RectifiedFlow(net=DenoiserV2(eng), sr=44100, drop_value=-4.0)
        t = torch.full((B, 1, 1), float(g["t_model_forward"]), device="cuda")
        dx = rf.model_forward(x0, t, cond, tc, float(g["guidance_timbre"]), float(g["guidance_structure"]))
        assert rel(dx, g["dx"]) < TOL[precision]
        out = rf.sample(x0, cond, tc, int(g["nb_steps"]), float(g["guidance_timbre"]), float(g["guidance_structure"]))
        assert rel(out, g["out"]) < TOL[precision]
        # replaying the captured graph gives the same answer (bitwise: same kernels, same order)
        out2 = rf.sample(x0, cond, tc, int(g["nb_steps"]), float(g["guidance_timbre"]), float(g["guidance_structure"]))
        assert torch.equal(out, out2)
        # host-buffer entry point
        host_out = torch.empty(x0.shape)
        eng.sample_host(x0.cpu(), cond.cpu(), tc.cpu(), host_out, int(g["nb_steps"]), float(g["guidance_timbre"]),
                        float(g["guidance_structure"]))
        assert torch.equal(host_out, out.cpu())
    finally:
        eng.close()


@pytest.mark.parametrize("name,variant,clamp", [("tiny", 0, 0.01), ("midi", 1, 0.1), ("base", 0, 0.1)])
def test_sample_matches_oracle(name, variant, clamp):
    """Fresh seeded inputs, ragged frame count, both CFG layouts; checker = CPU oracle."""
    from oracle import after_oracle as O
    frames, B, steps = 46, 3, 3  # 46 % 4 != 0: ragged last attention chunk
    eng, sd, mc = make_engine(name, 77, "fp32", frames)
    try:
        x0, cond, tc = synth.synth_inputs(B, mc.denoiser, seed=5, frames=frames)
        want = O.sample(sd, mc.denoiser, x0, cond, tc, steps, 2.0, 1.5, cfg_variant=variant, clamp=clamp)
        got = eng.sample(x0.cuda(), cond.cuda(), tc.cuda(), steps, 2.0, 1.5, cfg_variant=variant, clamp=clamp)
        assert rel(got, want) < 2e-4
        # linearity in the guidance: g_t = g_s = 0 reduces to the unconditional field
        want0 = O.sample(sd, mc.denoiser, x0, cond, tc, 1, 0.0, 0.0, cfg_variant=variant, clamp=clamp)
        got0 = eng.sample(x0.cuda(), cond.cuda(), tc.cuda(), 1, 0.0, 0.0, cfg_variant=variant, clamp=clamp)
        assert rel(got0, want0) < 2e-4
    finally:
        eng.close()


def test_full_size_properties():
    """BASELINE configs[1] size (base, B=8, T=256): finite, deterministic, and batch rows independent
    (stream b of a B=8 run equals the same stream run alone)."""
    eng, sd, mc = make_engine("base", 3, "fp32", 256, max_batch=8, max_steps=4)
    try:
        x0, cond, tc = synth.synth_inputs(8, mc.denoiser, seed=9)
        x0, cond, tc = x0.cuda(), cond.cuda(), tc.cuda()
        out = eng.sample(x0, cond, tc, 4, 2.0, 1.0)
        assert torch.isfinite(out).all()
        solo = eng.sample(x0[5:6], cond[5:6], tc[5:6], 4, 2.0, 1.0)
        assert rel(out[5:6], solo) < 1e-5
    finally:
        eng.close()


def test_errors_are_loud():
    eng, sd, mc = make_engine("tiny", 1, "fp32", 32, max_batch=2, max_steps=4)
    try:
        x0, cond, tc = synth.synth_inputs(2, mc.denoiser, seed=1, frames=32)
        with pytest.raises(RuntimeError):
            eng.sample(x0, cond, tc, 2)  # CPU tensors: no implicit copies, no CPU path
        with pytest.raises(ValueError):
            eng.sample(x0.cuda(), cond.cuda()[:, :3], tc.cuda(), 2)
        with pytest.raises(RuntimeError):
            eng.sample(x0.cuda(), cond.cuda(), tc.cuda(), 99)  # > max_steps
        x3, c3, t3 = synth.synth_inputs(3, mc.denoiser, seed=1, frames=32)
        with pytest.raises(RuntimeError):
            eng.sample(x3.cuda(), c3.cuda(), t3.cuda(), 2)  # > max_batch
    finally:
        eng.close()


def test_baseline_config_parity_50_steps():
    """BASELINE configs[1] at FULL size: base, B = 8 streams (24 CFG rows), T = 256, 50 Euler steps, CFG 2.0/1.0, fp32 mode
    vs the CPU oracle on ALL 8 streams (model.py:763-785) -- tolerance 1e-3 (north_star), per stream and overall; the
    bf16 mode (BASELINE configs[2] arithmetic) is gated at 1e-2 on the same 50-step trajectory."""
    from oracle import after_oracle as O
    B = 8
    eng, sd, mc = make_engine("base", 0, "fp32", 256, max_batch=B, max_steps=50)
    try:
        x0, cond, tc = synth.synth_inputs(B, mc.denoiser, seed=1234)
        want = O.sample(sd, mc.denoiser, x0, cond, tc, 50, 2.0, 1.0)
        got = eng.sample(x0.cuda(), cond.cuda(), tc.cuda(), 50, 2.0, 1.0)
        e = rel(got, want)
        per = [rel(got[b], want[b]) for b in range(B)]
        print(f"base B=8 T=256 50 steps fp32 mode: rel-L2 {e:.2e} (worst stream {max(per):.2e})")
        assert e < 1e-3 and max(per) < 1e-3
    finally:
        eng.close()
    eng, sd, mc = make_engine("base", 0, "bf16", 256, max_batch=B, max_steps=50)
    try:
        got = eng.sample(x0.cuda(), cond.cuda(), tc.cuda(), 50, 2.0, 1.0)
        e = rel(got, want)
        per = [rel(got[b], want[b]) for b in range(B)]
        print(f"base B=8 T=256 50 steps bf16 mode: rel-L2 {e:.2e} (worst stream {max(per):.2e})")
        assert e < 1e-2 and max(per) < 2e-2
    finally:
        eng.close()


def test_midi_config_full_length():
    """BASELINE configs[3] arithmetic at full size: midi (zs = 128 piano roll, window 16), MIDI CFG layout
    (export_midi.py:322-360, clamp 0.1), T = 256, 50 Euler steps, guidance 2.0 / 3.0."""
    from oracle import after_oracle as O
    eng, sd, mc = make_engine("midi", 5, "fp32", 256, max_batch=2, max_steps=50)
    try:
        x0, cond, tc = synth.synth_inputs(2, mc.denoiser, seed=77)
        want = O.sample(sd, mc.denoiser, x0, cond, tc, 50, 2.0, 3.0, cfg_variant=1, clamp=0.1)
        got = eng.sample(x0.cuda(), cond.cuda(), tc.cuda(), 50, 2.0, 3.0, cfg_variant=1, clamp=0.1)
        e = rel(got, want)
        print(f"midi T=256 50 steps fp32 mode: rel-L2 {e:.2e}")
        assert e < 1e-3
    finally:
        eng.close()


@pytest.mark.parametrize("frames", [1, 2, 5])
def test_degenerate_lengths(frames):
    """Shortest sequences (a single frame, a ragged single chunk): every kernel's tail handling."""
    from oracle import after_oracle as O
    eng, sd, mc = make_engine("tiny", 9, "fp32", frames, max_batch=1, max_steps=2)
    try:
        x0, cond, tc = synth.synth_inputs(1, mc.denoiser, seed=3, frames=frames)
        want = O.sample(sd, mc.denoiser, x0, cond, tc, 2, 1.0, 1.0)
        got = eng.sample(x0.cuda(), cond.cuda(), tc.cuda(), 2, 1.0, 1.0)
        assert rel(got, want) < 2e-4
    finally:
        eng.close()


def test_engine_from_run_folder(tmp_path):
    """A reference-style run folder (operative config.gin + checkpoint<step>_EMA.pt with net./encoder./encoder_time. keys,
    model.py:144-176, 264-265) loads into an Engine and samples like the oracle with the same weights."""
    from after_b200.engine import Engine
    from oracle import after_oracle as O
    from test_checkpoint import OPERATIVE  # tests/ is on sys.path (rootdir-relative "prepend" import mode)
    mc = config.get_config("tiny")
    den = synth.denoiser_state_dict(mc.denoiser, 5)
    state = {"net." + k: v for k, v in den.items()}
    state.update({"encoder." + k: v for k, v in synth.ecapa_state_dict(mc.timbre_encoder, 6).items()})
    state.update({"encoder_time." + k: v for k, v in synth.encoder1d_state_dict(mc.structure_encoder, 7).items()})
    run = tmp_path / "run"
    run.mkdir()
    torch.save({"model_state": state, "opt_state": {}}, run / "checkpoint500_EMA.pt")
    (run / "config.gin").write_text(OPERATIVE)
    eng = Engine.from_run(str(run), precision="fp32", max_batch=2, max_steps=3, seq_len=16)
    try:
        assert eng.has_denoiser and eng.has_structure and eng.has_timbre and not eng.has_codec
        x0, cond, tc = synth.synth_inputs(2, mc.denoiser, seed=9, frames=16)
        got = eng.sample(x0.cuda(), cond.cuda(), tc.cuda(), 3, 2.0, 1.0)
        assert rel(got, O.sample(den, mc.denoiser, x0, cond, tc, 3, 2.0, 1.0)) < TOL["fp32"]
    finally:
        eng.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_stale_rows_behind_a_smaller_batch_do_not_leak(precision):
    """attn_chunk_group_kernel reads key / value rows past the last chunk unconditionally (masked afterwards).  Rows behind
    the live batch may hold anything -- here the non-finite q|k|v a NaN input left there -- and must not reach the result."""
    frames = 22  # ragged: the last chunk has 2 queries and reads 2 + 8 rows of the next sequence / the stale area
    eng, sd, mc = make_engine("base", 5, precision, frames, max_batch=2, max_steps=2)
    fresh, _, _ = make_engine("base", 5, precision, frames, max_batch=2, max_steps=2)
    try:
        x0, cond, tc = (t.cuda() for t in synth.synth_inputs(2, mc.denoiser, frames=frames))
        bad = torch.full_like(x0, float("nan"))
        poisoned = eng.sample(bad, cond, tc, 2, 2.0, 1.0)
        assert not torch.isfinite(poisoned).all()
        out = eng.sample(x0[:1].contiguous(), cond[:1].contiguous(), tc[:1].contiguous(), 2, 2.0, 1.0)
        ref = fresh.sample(x0[:1].contiguous(), cond[:1].contiguous(), tc[:1].contiguous(), 2, 2.0, 1.0)
        assert torch.isfinite(out).all()
        assert torch.equal(out, ref)
    finally:
        eng.close()
        fresh.close()
