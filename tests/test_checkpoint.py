"""Host-side ingestion of reference training runs (after_b200/checkpoint.py): gin subset parser, state-dict split,
codec topology inference.  CPU only."""
import os
from dataclasses import asdict

import pytest
import torch

from after_b200 import checkpoint as C
from after_b200 import config, synth

# an operative-config style file (one-line bindings, as gin.operative_config_str() writes next to the checkpoints,
# after/diffusion/model.py:264-265), written for this test
OPERATIVE = """
# Macros:
# ==============================================================================
ATTENTION_CHUNK_SIZE = 4
IN_SIZE = 64
LOCAL_ATTENTION_SIZE = 8
N_SIGNAL = 256
SR = 44100
ZS_CHANNELS = 12
ZT_CHANNELS = 6

# Parameters for Base:
# ==============================================================================
Base.drop_value = -4.0
Base.encoder = @encoder/ECAPATDNN()
Base.encoder_time = @encoder_time/Encoder1D()
Base.net = @DenoiserV2()
Base.sr = %SR

# Parameters for DenoiserV2:
# ==============================================================================
DenoiserV2.attention_chunk_size = %ATTENTION_CHUNK_SIZE
DenoiserV2.causal = True
DenoiserV2.cond_dim = %ZT_CHANNELS
DenoiserV2.embed_dim = 256
DenoiserV2.local_attention_size = %LOCAL_ATTENTION_SIZE
DenoiserV2.mlp_multiplier = 3
DenoiserV2.n_channels = %IN_SIZE
DenoiserV2.n_layers = 6
DenoiserV2.noise_embed_dims = 64
DenoiserV2.pos_emb_type = 'rotary'   # trailing comment
DenoiserV2.seq_len = %N_SIGNAL
DenoiserV2.tcond_dim = %ZS_CHANNELS

# Parameters for encoder/ECAPATDNN:
# ==============================================================================
encoder/ECAPATDNN.attention_channels = 128
encoder/ECAPATDNN.channels = [256, 256,
                              256, 512]
encoder/ECAPATDNN.in_size = %IN_SIZE
encoder/ECAPATDNN.out_dim = %ZT_CHANNELS

# Parameters for encoder_time/Encoder1D:
# ==============================================================================
encoder_time/Encoder1D.channels = [64, 128, 256, 256, %ZS_CHANNELS]
encoder_time/Encoder1D.in_size = %IN_SIZE
encoder_time/Encoder1D.ratios = [1, 1, 1, 1]
encoder_time/get_padding.mode = 'causal'

# Parameters for classifier/Encoder1D:
# ==============================================================================
classifier/Encoder1D.channels = [64, 64, 64, 64, %ZT_CHANNELS]
classifier/Encoder1D.ratios = [1, 2, 2, 2, 1]
"""


def test_operative_config_parses_to_tiny():
    mc = C.model_config_from_gin(OPERATIVE, name="tiny")
    assert asdict(mc) == asdict(config.get_config("tiny"))
    macros, b = C.parse_gin(OPERATIVE)
    assert macros["N_SIGNAL"] == 256 and b[("classifier", "Encoder1D")]["ratios"] == [1, 2, 2, 2, 1]
    assert isinstance(b[("", "Base")]["net"], str) and b[("", "Base")]["net"].startswith("@")


@pytest.mark.reference
@pytest.mark.parametrize("name", ["base", "tiny", "midi"])
def test_reference_gin_files_parse_to_shipped_configs(name):
    with open(f"/root/reference/after/diffusion/configs/{name}.gin") as fh:
        mc = C.model_config_from_gin(fh.read(), in_size=64, n_signal=256, name=name)
    assert asdict(mc) == asdict(config.get_config(name))


def test_load_run_round_trip(tmp_path):
    mc = config.get_config("tiny")
    den = synth.denoiser_state_dict(mc.denoiser, 5)
    tim = synth.ecapa_state_dict(mc.timbre_encoder, 6)
    stc = synth.encoder1d_state_dict(mc.structure_encoder, 7)
    state = {}
    state.update({"net." + k: v for k, v in den.items()})
    state.update({"encoder." + k: v for k, v in tim.items()})
    state.update({"encoder_time." + k: v for k, v in stc.items()})
    state["classifier.net.0.weight"] = torch.zeros(3)  # training-only module: ignored
    run = tmp_path / "run"
    run.mkdir()
    for step in (1000, 25000):  # the latest EMA checkpoint wins (export.py:52-60)
        torch.save({"model_state": state if step == 25000 else {}, "opt_state": {}}, run / f"checkpoint{step}_EMA.pt")
    # IN_SIZE is unknown to a static config: it is recovered from the checkpoint
    (run / "config.gin").write_text(OPERATIVE.replace("IN_SIZE = 64", "IN_SIZE = None"))
    got = C.load_run(str(run))
    assert got["checkpoint"].endswith("checkpoint25000_EMA.pt")
    assert asdict(got["model"]) == {**asdict(mc), "name": "run"}
    for name, want in (("denoiser_state", den), ("timbre_state", tim), ("structure_state", stc)):
        assert set(got[name]) == set(want)
        assert all(torch.equal(got[name][k], want[k]) for k in want)
    with pytest.raises(FileNotFoundError):
        C.load_run(str(run), step=7)


def test_codec_prefix_and_topology_inference():
    for acfg in (config.base_autoencoder(), config.small_autoencoder()):
        sd = synth.autoencoder_state_dict(acfg, 3)
        wrapped = {"model." + k: v for k, v in sd.items()}
        wrapped["latent_mean"] = torch.zeros(4)  # export wrappers carry extra buffers
        stripped = C.strip_codec_prefix(wrapped)
        assert set(stripped) == set(sd)
        assert asdict(C.autoencoder_config_from_state(stripped)) == asdict(acfg)


def test_unsupported_nets_are_loud():
    with pytest.raises(ValueError):
        C.model_config_from_gin("Denoiser.n_channels = 64\n")
    with pytest.raises(KeyError):
        C.parse_gin("DenoiserV2.n_channels = %UNDEFINED\n")


# What gin.operative_config_str() writes for base.gin under ``from __gin__ import dynamic_registration``
# (after/diffusion/model.py:264-265): fully qualified selectors, and every binding longer than 80 columns wrapped as
# ``selector = \`` with the value indented on the following line(s).
OPERATIVE_WRAPPED = """\
import after
import after.diffusion
import after.diffusion.networks.ecapa_encoder
import after.diffusion.networks.encoder
import after.diffusion.networks.transformerv2
import cached_conv.convs

# Macros:
# ==============================================================================
ATTENTION_CHUNK_SIZE = 4
IN_SIZE = 64
LOCAL_ATTENTION_SIZE = 8
N_SIGNAL = 256
SR = 44100
ZS_CHANNELS = 12
ZT_CHANNELS = 6

# Parameters for diffusion.model.Base:
# ==============================================================================
diffusion.model.Base.drop_value = -4.0
diffusion.model.Base.encoder = \\
    @encoder/diffusion.networks.ecapa_encoder.ECAPATDNN()
diffusion.model.Base.encoder_time = \\
    @encoder_time/diffusion.networks.encoder.Encoder1D()
diffusion.model.Base.net = @diffusion.networks.transformerv2.DenoiserV2()
diffusion.model.Base.sr = %SR

# Parameters for diffusion.networks.transformerv2.DenoiserV2:
# ==============================================================================
diffusion.networks.transformerv2.DenoiserV2.attention_chunk_size = \\
    %ATTENTION_CHUNK_SIZE
diffusion.networks.transformerv2.DenoiserV2.causal = True
diffusion.networks.transformerv2.DenoiserV2.cond_dim = %ZT_CHANNELS
diffusion.networks.transformerv2.DenoiserV2.dropout = 0.1
diffusion.networks.transformerv2.DenoiserV2.embed_dim = 512
diffusion.networks.transformerv2.DenoiserV2.local_attention_size = \\
    %LOCAL_ATTENTION_SIZE
diffusion.networks.transformerv2.DenoiserV2.mlp_multiplier = 3
diffusion.networks.transformerv2.DenoiserV2.n_channels = %IN_SIZE
diffusion.networks.transformerv2.DenoiserV2.n_layers = 6
diffusion.networks.transformerv2.DenoiserV2.noise_embed_dims = 64
diffusion.networks.transformerv2.DenoiserV2.pos_emb_type = 'rotary'
diffusion.networks.transformerv2.DenoiserV2.seq_len = %N_SIGNAL
diffusion.networks.transformerv2.DenoiserV2.tcond_dim = %ZS_CHANNELS

# Parameters for encoder/diffusion.networks.ecapa_encoder.ECAPATDNN:
# ==============================================================================
encoder/diffusion.networks.ecapa_encoder.ECAPATDNN.attention_channels = 128
encoder/diffusion.networks.ecapa_encoder.ECAPATDNN.channels = \\
    [512, 512, 512, 1024]
encoder/diffusion.networks.ecapa_encoder.ECAPATDNN.dilations = [1, 1, 1, 1]
encoder/diffusion.networks.ecapa_encoder.ECAPATDNN.global_context = True
encoder/diffusion.networks.ecapa_encoder.ECAPATDNN.in_size = %IN_SIZE
encoder/diffusion.networks.ecapa_encoder.ECAPATDNN.kernel_sizes = \\
    [3, 3, 3, 3]
encoder/diffusion.networks.ecapa_encoder.ECAPATDNN.out_dim = %ZT_CHANNELS
encoder/diffusion.networks.ecapa_encoder.ECAPATDNN.res2net_scale = 8
encoder/diffusion.networks.ecapa_encoder.ECAPATDNN.se_channels = 128
encoder/diffusion.networks.ecapa_encoder.ECAPATDNN.use_tanh = False

# Parameters for encoder_time/diffusion.networks.encoder.Encoder1D:
# ==============================================================================
encoder_time/diffusion.networks.encoder.Encoder1D.channels = \\
    [64, 128, 256, 512, %ZS_CHANNELS]
encoder_time/diffusion.networks.encoder.Encoder1D.in_size = %IN_SIZE
encoder_time/diffusion.networks.encoder.Encoder1D.kernel_size = 5
encoder_time/diffusion.networks.encoder.Encoder1D.ratios = [1, 1, 1, 1]
encoder_time/diffusion.networks.encoder.Encoder1D.use_tanh = False

# Parameters for encoder_time/convs.get_padding:
# ==============================================================================
encoder_time/convs.get_padding.mode = 'causal'
"""


def test_wrapped_operative_config_parses_to_base():
    """backslash-continued bindings + fully qualified selectors, as a real run folder holds them."""
    mc = C.model_config_from_gin(OPERATIVE_WRAPPED, name="base")
    assert asdict(mc) == asdict(config.get_config("base"))
    _, b = C.parse_gin(OPERATIVE_WRAPPED)
    assert b[("", "DenoiserV2")]["attention_chunk_size"] == 4
    assert str(b[("", "Base")]["encoder"]).startswith("@encoder/")
    assert b[("encoder", "ECAPATDNN")]["channels"] == [512, 512, 512, 1024]


def test_reference_defaults_are_not_assumed():
    """A config that does not bind pos_emb_type / causal asks for the reference defaults ('learnable', non-causal,
    transformerv2.py:463-476), which this path does not implement: rejected, not silently run as rotary/causal."""
    text = "\n".join(l for l in OPERATIVE.splitlines() if "pos_emb_type" not in l)
    with pytest.raises(ValueError):
        C.model_config_from_gin(text)
    text = "\n".join(l for l in OPERATIVE.splitlines() if "DenoiserV2.causal" not in l)
    with pytest.raises(ValueError):
        C.model_config_from_gin(text)


def test_codec_state_round_trips_through_a_real_torchscript_file(tmp_path):
    """f4: ``codec_state_from_torchscript`` on an actual ``torch.jit.save``d export-shaped module (wrapper prefix + extra
    buffers, export_autoencoder.py:16-66): every AutoEncoder tensor comes back bit-identical and the topology is inferred."""
    from ts_helpers import save_codec_ts
    acfg = config.small_autoencoder()
    sd = synth.autoencoder_state_dict(acfg, 3)
    path = save_codec_ts(sd, str(tmp_path / "export.ts"), acfg.z_channels, acfg.ratio)
    got = C.codec_state_from_torchscript(path)
    assert set(got) == set(sd)
    assert all(torch.equal(got[k], sd[k]) for k in sd)
    assert asdict(C.autoencoder_config_from_state(got)) == asdict(acfg)
