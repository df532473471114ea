"""Golden fixtures for the STREAMING denoiser path (per-diffusion-step rolling KV caches), produced by the
UNMODIFIED reference.  Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden_stream.py

What the export scripts do with gin (``after_scripts/export.py:74-79``: bind
``transformerv2.MHAttention.max_cache_size = LOCAL_ATTENTION_SIZE`` before instantiating the model) is done here by
giving ``MHAttention.__init__`` that default while the reference ``DenoiserV2`` is constructed; the module code itself is
untouched.  The block loop is the one of the exported ``Streamer.sample`` (``export.py:398-416``): per diffusion step i,
``model_forward(..., cache_index=i)`` then ``net.roll_cache(T, i)``.
"""
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.filterwarnings("ignore")

import ref_shims as R  # noqa: E402
from after_b200 import config, synth  # noqa: E402
from make_golden import save  # noqa: E402


def ref_stream_denoiser(name, seed, cache_size):
    ns = R.install()
    MHA = ns.transformerv2.MHAttention
    orig = MHA.__init__

    def patched(self, *a, **k):
        k.setdefault("max_cache_size", cache_size)
        orig(self, *a, **k)

    MHA.__init__ = patched
    try:
        net = ns.transformerv2.DenoiserV2(**R.DENOISER_CFG[name]).eval()
    finally:
        MHA.__init__ = orig
    missing, unexpected = net.load_state_dict(synth.denoiser_state_dict(config.get_config(name).denoiser, seed), strict=False)
    assert not unexpected and all("cache" in m for m in missing), (missing, unexpected)
    return net


@torch.no_grad()
def main():
    ns = R.install()

    # ---- DenoiserV2.forward(cache_index) + roll_cache over consecutive blocks -------------------------------
    # (name, N, frames per block, roll size, blocks, cache indices used round-robin, weight seed)
    for tag, name, n, frames, roll, blocks, idxs, wseed in (("tiny", "tiny", 3, 4, 4, 5, (0, ), 61),
                                                            ("tiny_t8", "tiny", 2, 8, 8, 4, (0, 1), 62),
                                                            ("base", "base", 3, 4, 4, 4, (0, ), 63),
                                                            ("midi", "midi", 3, 4, 4, 6, (2, ), 64)):
        cfg = config.get_config(name).denoiser
        net = ref_stream_denoiser(name, wseed, cfg.local_attention_size)
        g = torch.Generator().manual_seed(500 + wseed)
        xs, ts, conds, tcs, outs, cis = [], [], [], [], [], []
        for b in range(blocks):
            for ci in idxs:
                x = torch.randn(n, cfg.n_channels, frames, generator=g)
                t = torch.rand(n, generator=g)
                cond = torch.randn(n, cfg.cond_dim, generator=g)
                tc = torch.randn(n, cfg.tcond_dim, frames, generator=g)
                y = net(x, time=t.reshape(n, 1, 1), cond=cond, time_cond=tc, cache_index=ci)
                net.roll_cache(roll, ci)
                xs.append(x); ts.append(t); conds.append(cond); tcs.append(tc); outs.append(y); cis.append(ci)
        save(f"stream_denoiser_{tag}", weight_seed=wseed, cache_size=cfg.local_attention_size, roll=roll,
             cache_index=torch.tensor(cis), x=torch.stack(xs), time=torch.stack(ts), cond=torch.stack(conds),
             time_cond=torch.stack(tcs), out=torch.stack(outs))

    # ---- the exported Streamer's sampling loop over consecutive audio blocks (export.py:356-416) ----------------
    for name, frames, steps, blocks, wseed in (("tiny", 4, 4, 6, 71), ("base", 4, 3, 4, 72)):
        cfg = config.get_config(name).denoiser
        net = ref_stream_denoiser(name, wseed, cfg.local_attention_size)
        rf = ns.model.RectifiedFlow(net=net, sr=44100, encoder=None, encoder_time=None, classifier=None,
                                    drop_value=-4.0, device="cpu").eval()
        g = torch.Generator().manual_seed(600 + wseed)
        g_t, g_s = 2.0, 1.0
        x0s, conds, tcs, outs = [], [], [], []
        for b in range(blocks):
            x0 = torch.randn(1, cfg.n_channels, frames, generator=g)
            cond = torch.randn(1, cfg.cond_dim, generator=g)
            tc = torch.randn(1, cfg.tcond_dim, frames, generator=g)
            x = x0
            t = torch.linspace(0, 1, steps + 1)
            dt = 1 / steps
            for i, tv in enumerate(t[:-1]):
                x = x + rf.model_forward(x=x, time=tv.repeat(x.shape[0], 1, x.shape[-1]), cond=cond, time_cond=tc,
                                         guidance_timbre=g_t, guidance_structure=g_s, cache_index=i) * dt
                net.roll_cache(x.shape[-1], i)
            x0s.append(x0); conds.append(cond); tcs.append(tc); outs.append(x)
        save(f"stream_sample_{name}", weight_seed=wseed, cache_size=cfg.local_attention_size, nb_steps=steps,
             guidance_timbre=g_t, guidance_structure=g_s, x0=torch.stack(x0s), cond=torch.stack(conds),
             time_cond=torch.stack(tcs), out=torch.stack(outs))


if __name__ == "__main__":
    main()
