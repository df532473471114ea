"""Host logic of the nn_tilde-shaped streamers (after_b200/streamer.py) on CPU: a stand-in engine that answers the Engine
calls with the CPU oracle lets the method tables, buffers, attribute setters, piano-roll rasterisation and guidance layout be
checked without a GPU (the arithmetic itself is covered by the -m gpu parity tests)."""
import types

import pytest
import torch

from after_b200 import _lib as L
from after_b200 import config, synth
from after_b200.streamer import MidiStreamer, Streamer
from oracle import after_oracle as O
from oracle import after_oracle_stream as S


class OracleEngine:
    """Duck-typed ``Engine``: same attributes / methods the streamers use, computed by the oracle on CPU."""

    def __init__(self, den_cfg, acfg, se_cfg, te_cfg, streaming=False, drop_value=-4.0, stream_slots=0):
        self.den_cfg, self.acfg, self.se_cfg, self.te_cfg = den_cfg, acfg, se_cfg, te_cfg
        self.device = torch.device("cpu")
        self.cfg = types.SimpleNamespace(tcond_dim=den_cfg.tcond_dim, cond_dim=den_cfg.cond_dim, n_channels=den_cfg.n_channels,
                                         drop_value=drop_value, max_cache_size=den_cfg.local_attention_size if streaming else 0)
        self.ae_ratio = acfg.ratio
        self.stream_slots = stream_slots  # > 0: the codec copies / structure encoder carry the exported model's state
        self.codec_state = [dict() for _ in range(stream_slots)]
        self.enc_state = [dict() for _ in range(stream_slots)]
        self.stream_calls = []
        self.has_denoiser = self.has_codec = self.has_timbre = True
        self.has_structure = se_cfg is not None
        self.sd_den = synth.denoiser_state_dict(den_cfg, 1)
        self.sd_ae = synth.autoencoder_state_dict(acfg, 2)
        self.sd_se = synth.encoder1d_state_dict(se_cfg, 3) if se_cfg is not None else None
        self.sd_te = synth.ecapa_state_dict(te_cfg, 4)
        self.cache = O.StreamCache(den_cfg, den_cfg.local_attention_size) if streaming else None
        self.calls = []

    @property
    def streaming(self):
        return self.cache is not None

    def ae_encode(self, x):
        return O.ae_encode(self.sd_ae, self.acfg, x)

    def ae_decode(self, z):
        return O.ae_decode(self.sd_ae, self.acfg, z)

    def structure_encode(self, z):
        return O.encoder1d_forward(self.sd_se, self.se_cfg, z)

    def ae_encode_stream(self, slot, x):
        self.stream_calls.append(("encode", slot))
        return S.ae_encode_stream(self.sd_ae, self.acfg, self.codec_state[slot], x, gn_latent_frames=8)

    def ae_decode_stream(self, slot, z):
        self.stream_calls.append(("decode", slot))
        return S.ae_decode_stream(self.sd_ae, self.acfg, self.codec_state[slot], z, gn_latent_frames=8)

    def structure_encode_stream(self, slot, z):
        self.stream_calls.append(("structure", slot))
        return S.encoder1d_forward_stream(self.sd_se, self.se_cfg, self.enc_state[slot], z)

    def timbre_encode(self, z):
        return O.ecapa_forward(self.sd_te, self.te_cfg, z)

    def latent_map(self, x, direction):
        return O.latent_map(getattr(self, "sd_map", None), x, direction)

    def sample(self, x0, cond, tc, nb_steps, g_t=1.0, g_s=1.0, cfg_variant=L.CFG_AUDIO, clamp=0.01):
        self.calls.append(("sample", cfg_variant, clamp, nb_steps, g_t, g_s))
        return O.sample(self.sd_den, self.den_cfg, x0, cond, tc, nb_steps, g_t, g_s, cfg_variant=cfg_variant, clamp=clamp)

    def sample_stream(self, x0, cond, tc, nb_steps, g_t=1.0, g_s=1.0, cfg_variant=L.CFG_AUDIO, clamp=0.1):
        self.calls.append(("sample_stream", cfg_variant, clamp, nb_steps, g_t, g_s))
        return O.sample_stream(self.sd_den, self.den_cfg, self.cache, x0, cond, tc, nb_steps, g_t, g_s, cfg_variant=cfg_variant,
                               clamp=clamp)


ACFG = config.small_autoencoder()  # ratio 128, 8 latent channels


def small_engine(midi: bool, streaming: bool = False, stream_slots: int = 0):
    den = config.DenoiserConfig(n_channels=ACFG.z_channels, embed_dim=256, n_layers=2, tcond_dim=128 if midi else 12,
                                local_attention_size=16 if midi else 8)
    se = None if midi else config.Encoder1DConfig(in_size=ACFG.z_channels, channels=[16, 16, 12], ratios=[1, 1])
    te = config.EcapaConfig(in_size=ACFG.z_channels, channels=[32, 32, 32, 64], attention_channels=16, se_channels=16)
    return OracleEngine(den, ACFG, se, te, streaming, stream_slots=stream_slots)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def test_audio_streamer_methods_buffers_and_attributes():
    eng = small_engine(midi=False)
    st = Streamer(eng, n_signal_timbre=8, chunk_size=4)
    r = ACFG.ratio
    assert st.methods == {"forward": (2, 1, 1, 1), "structure": (1, 1, 12, r), "timbre": (1, 1, 6, r), "diffuse": (18, r, 8, r),
                          "generate": (18, r, 1, 1), "generate_timbre": (7, 1, 1, 1), "decode": (8, r, 1, 1)}
    assert st.set_nb_steps(2) == 0 and st.get_nb_steps() == 2 and st.nb_steps == (2, )
    st.set_guidance_timbre(2.5); st.set_guidance_structure(0.0)
    frames = 4
    g = torch.Generator().manual_seed(0)
    audio = torch.rand(2, 2, frames * r, generator=g) * 2 - 1
    noise = torch.randn(2, 8, frames, generator=g)
    out = st.forward(audio, noise=noise)
    assert out.shape == (2, 1, frames * r)
    assert torch.equal(out[0], out[1])  # only row 0 is diffused, then repeated (export.py:445-448)
    assert eng.calls == [("sample", L.CFG_AUDIO, 0.1, 2, 2.5, 0.0)]  # audio layout, the streamer's 0.1 clamp (export.py:389-390)
    # rolling timbre history: zeros shifted left by the new latents (export.py:418-429)
    z_t = eng.ae_encode(audio[:, 1:])
    assert torch.equal(st.previous_timbre[:2, :, -frames:], z_t) and float(st.previous_timbre[:2, :, :-frames].abs().max()) == 0.0
    # the whole chain equals the oracle chain on row 0
    cond = eng.timbre_encode(torch.cat([torch.zeros(1, 8, 8), z_t[:1]], -1)[..., frames:])
    tcond = eng.structure_encode(eng.ae_encode(audio[:1, :1]))
    want = eng.ae_decode(O.sample(eng.sd_den, eng.den_cfg, noise[:1], cond, tcond, 2, 2.5, 0.0, clamp=0.1))
    assert rel(out[:1], want) < 1e-5
    with pytest.raises(ValueError):
        st.diffuse(torch.zeros(1, 5, 4))
    # without a projection the exported model's latent2map / map2latent are the time average repeated (export.py:143, 494-508)
    m = torch.arange(16.0).view(1, 2, 8)
    assert torch.equal(st.latent2map(m), m.mean(-1, keepdim=True).expand(1, 2, 8))
    assert torch.equal(st.map2latent(m), st.latent2map(m))


def test_streaming_engine_routes_to_sample_stream():
    eng = small_engine(midi=False, streaming=True)
    st = Streamer(eng, n_signal_timbre=8)
    st.set_nb_steps(2)
    x = torch.randn(1, 18, 4, generator=torch.Generator().manual_seed(1))
    a = st.diffuse(x, noise=torch.zeros(1, 8, 4))
    b = st.diffuse(x, noise=torch.zeros(1, 8, 4))
    assert [c[0] for c in eng.calls] == ["sample_stream", "sample_stream"]
    assert rel(a, b) > 1e-4  # the second block attends to the first one's keys / values: state is carried


def test_streamer_routes_the_two_codec_copies_like_the_export():
    """export.py:161-168, 418-455: structure -> emb_model_structure.encode (slot 0) + encoder_time.forward_stream, timbre ->
    emb_model_timbre.encode (slot 1), decode -> emb_model_structure.decode (slot 0); state is carried from buffer to buffer."""
    eng = small_engine(midi=False, streaming=True, stream_slots=2)
    st = Streamer(eng, n_signal_timbre=8)
    st.set_nb_steps(1)
    r = ACFG.ratio
    g = torch.Generator().manual_seed(5)
    audio = torch.rand(1, 2, 4 * r, generator=g) * 2 - 1
    noise = torch.randn(1, 8, 4, generator=g)
    a = st.forward(audio, noise=noise)
    assert eng.stream_calls == [("encode", 0), ("structure", 0), ("encode", 1), ("decode", 0)]
    b = st.forward(audio, noise=noise)
    assert a.shape == b.shape == (1, 1, 4 * r)
    assert rel(a, b) > 1e-3  # same buffer twice gives different audio: every stage carries state
    midi = small_engine(midi=True, streaming=True, stream_slots=1)
    ms = MidiStreamer(midi, n_poly=2, n_signal_timbre=8)
    ms.timbre(torch.rand(1, 1, 4 * r, generator=g))
    ms.decode(torch.randn(1, 8, 4, generator=g))
    assert midi.stream_calls == [("encode", 0), ("decode", 0)]  # the MIDI export holds one codec copy (export_midi.py:165)


def reference_piano_roll(notes, n_poly):
    """Literal transcription of the loop in after_scripts/export_midi.py:408-415 (batch row 0)."""
    T = notes.shape[-1]
    time_cond = torch.zeros((1, 128, T))
    for i in range(n_poly):
        for j in range(T):
            if notes[0, 2 * i + 1, j] > 0:
                time_cond[:, notes[:, 2 * i].long(), j] = notes[:, 2 * i + 1, j] / 128
    return time_cond


def test_midi_streamer_piano_roll_and_guidance_layout():
    eng = small_engine(midi=True)
    st = MidiStreamer(eng, n_poly=3, n_signal_timbre=8)
    r = ACFG.ratio
    assert st.methods == {"timbre": (1, 1, 6, r), "generate": (12, r, 1, 1), "diffuse": (12, r, 8, r), "decode": (8, r, 1, 1)}
    T = 8
    g = torch.Generator().manual_seed(2)
    notes = torch.zeros(1, 6, T)
    notes[0, 0] = 60; notes[0, 1] = torch.tensor([0, 90, 90, 90, 0, 0, 70, 70.])         # held note with gaps
    notes[0, 2] = torch.tensor([64, 64, 64, 67, 67, 67, 67, 67.]); notes[0, 3, 2:6] = 100   # pitch changes inside the buffer
    notes[0, 4] = 60; notes[0, 5, 3] = 30                                                  # a later voice overwrites voice 0
    assert torch.equal(st.piano_roll(notes), reference_piano_roll(notes, 3))
    rnd = torch.cat([torch.randint(21, 109, (1, 1, T), generator=g).float() if c % 2 == 0 else
                     torch.rand(1, 1, T, generator=g) * 127 * (torch.rand(1, 1, T, generator=g) > 0.4) for c in range(6)], 1)
    assert torch.equal(st.piano_roll(rnd), reference_piano_roll(rnd, 3))
    # diffuse: zsem = mean of the last zt channels * latent_range, MIDI CFG layout with the 0.1 clamp (export_midi.py:322-360)
    st.set_nb_steps(2); st.set_guidance_timbre(1.5); st.set_guidance_structure(3.0)
    x = torch.cat([notes, torch.randn(1, 6, T, generator=g)], 1)
    noise = torch.randn(1, 8, T, generator=g)
    out = st.diffuse(x, noise=noise)
    assert eng.calls[-1] == ("sample", L.CFG_MIDI, 0.1, 2, 1.5, 3.0)
    want = O.sample(eng.sd_den, eng.den_cfg, noise, x[:, -6:].mean(-1), reference_piano_roll(notes, 3), 2, 1.5, 3.0,
                    cfg_variant=O.CFG_MIDI, clamp=0.1)
    assert rel(out, want) < 1e-6
    # timbre: embedding of the rolling history repeated over the encoded frames (export_midi.py:383-398)
    audio = torch.rand(2, 1, 4 * r, generator=g) * 2 - 1
    zs = st.timbre(audio)
    assert zs.shape == (2, 6, 4) and torch.equal(zs[..., 0], zs[..., 3])
    with pytest.raises(RuntimeError):
        MidiStreamer(small_engine(midi=False))
