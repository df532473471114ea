"""GPU runs of the boundary itself: the plain-C host of the ABI and checkpoint / TorchScript ingestion into an Engine."""
import os
import subprocess

import pytest
import torch

from after_b200 import config, synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm())


def test_plain_c_host_samples_on_the_gpu(tmp_path):
    """examples/c_host.c bound to libafter_b200.so with nothing but the header: its success path (create -> load ->
    finalize -> after_sample_host -> 'sample ok') on a real device."""
    from after_b200 import build
    build.build()
    libdir = os.path.join(ROOT, "after_b200", "lib")
    exe = tmp_path / "c_host"
    subprocess.check_call(["gcc", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_host.c"),
                           "-o", str(exe), "-L", libdir, "-lafter_b200", f"-Wl,-rpath,{libdir}"])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "sample ok" in r.stdout and "yes" in r.stdout


def test_engine_from_run_with_torchscript_codec(tmp_path):
    """f4 end to end: run folder (wrapped operative config.gin + EMA checkpoint) + a torch.jit.save'd codec export ->
    Engine.from_run -> encode / sample / decode match the oracle with the same weights."""
    from after_b200.engine import Engine
    from oracle import after_oracle as O
    from test_checkpoint import OPERATIVE
    from ts_helpers import save_codec_ts
    mc = config.get_config("tiny")
    acfg = config.small_autoencoder()
    den = synth.denoiser_state_dict(mc.denoiser, 5)
    ae = synth.autoencoder_state_dict(acfg, 8)
    state = {"net." + k: v for k, v in den.items()}
    state.update({"encoder." + k: v for k, v in synth.ecapa_state_dict(mc.timbre_encoder, 6).items()})
    state.update({"encoder_time." + k: v for k, v in synth.encoder1d_state_dict(mc.structure_encoder, 7).items()})
    run = tmp_path / "run"
    run.mkdir()
    torch.save({"model_state": state, "opt_state": {}}, run / "checkpoint500_EMA.pt")
    (run / "config.gin").write_text(OPERATIVE)
    ts = save_codec_ts(ae, str(tmp_path / "export.ts"), acfg.z_channels, acfg.ratio)
    frames = 16
    eng = Engine.from_run(str(run), codec_ts=ts, precision="fp32", max_batch=2, max_steps=3, seq_len=frames,
                          max_samples=frames * acfg.ratio)
    try:
        assert eng.has_codec and eng.ae_ratio == acfg.ratio
        audio = synth.synth_audio(2, frames * acfg.ratio, seed=5)
        z_ref = O.ae_encode(ae, acfg, audio)
        assert rel(eng.ae_encode(audio.cuda()), z_ref) < 1e-3
        assert rel(eng.ae_decode(z_ref.cuda()), O.ae_decode(ae, acfg, z_ref)) < 1e-3
        x0, cond, tc = synth.synth_inputs(2, mc.denoiser, seed=9, frames=frames)
        assert rel(eng.sample(x0.cuda(), cond.cuda(), tc.cuda(), 3, 2.0, 1.0), O.sample(den, mc.denoiser, x0, cond, tc, 3, 2.0, 1.0)) < 2e-4
    finally:
        eng.close()


def test_latent_map_matches_reference_fixture_and_identity(golden):
    """after_latent_map (Streamer.latent2map / map2latent, export.py:494-508) against the reference projection's outputs,
    and the identity projection of a handle without AFTER_MODULE_LATENT_MAP tensors."""
    import torch
    from after_b200 import config, synth
    from after_b200.engine import Engine
    g = golden("latent_map")
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd.")}
    mc = config.get_config("tiny")
    den = synth.denoiser_state_dict(mc.denoiser, 0)
    eng = Engine(model=mc, denoiser_state=den, precision="fp32", max_batch=1, max_steps=1, seq_len=8, latent_map_state=sd)
    plain = Engine(model=mc, denoiser_state=den, precision="fp32", max_batch=1, max_steps=1, seq_len=8)
    try:
        l2m = eng.latent_map(torch.from_numpy(g["latents"]).cuda(), 0)
        m2l = eng.latent_map(torch.from_numpy(g["maps"]).cuda(), 1)
        assert l2m.shape == g["latent2map"].shape and m2l.shape == g["map2latent"].shape
        assert torch.allclose(l2m.cpu(), torch.from_numpy(g["latent2map"]), atol=2e-6, rtol=1e-5)
        assert torch.allclose(m2l.cpu(), torch.from_numpy(g["map2latent"]), atol=2e-6, rtol=1e-5)
        x = torch.from_numpy(g["maps"]).cuda()
        ident = plain.latent_map(x, 1)
        assert torch.allclose(ident, x.mean(-1, keepdim=True).expand_as(x), atol=1e-6)
        with pytest.raises(Exception):
            eng.latent_map(torch.zeros(1, 5, 4).cuda(), 0)  # wrong channel count for the loaded projection
    finally:
        eng.close()
        plain.close()
