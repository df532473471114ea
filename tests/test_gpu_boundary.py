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
