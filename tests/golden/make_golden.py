"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case the synthetic state dict from ``after_b200.synth`` (seeded, reference key layout)
is loaded ``strict=True`` into the reference nn.Module imported from /root/reference under
the shims of ``ref_shims.py``; the module is run on seeded inputs on CPU in fp32 and inputs +
outputs are stored as ``<case>.npz``.  The fixtures are what pins ``oracle/after_oracle.py``
(``tests/test_oracle_golden.py``) and, through it and directly, the CUDA path.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.filterwarnings("ignore")

import ref_shims as R  # noqa: E402
from after_b200 import config, synth  # noqa: E402


def save(name, **arrays):
    out = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
           for k, v in arrays.items()}
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in out.items()),
          f"-> {os.path.getsize(path) / 1024:.0f} KiB")


def ref_denoiser(name, seed):
    ns = R.install()
    net = ns.transformerv2.DenoiserV2(**R.DENOISER_CFG[name]).eval()
    net.load_state_dict(synth.denoiser_state_dict(config.get_config(name).denoiser, seed), strict=True)
    return net


def inputs(cfg, n, frames, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, cfg.n_channels, frames, generator=g)
    t = torch.rand(n, generator=g)
    cond = torch.randn(n, cfg.cond_dim, generator=g)
    tc = torch.randn(n, cfg.tcond_dim, frames, generator=g)
    return x, t, cond, tc


@torch.no_grad()
def main():
    ns = R.install()

    # ---- denoiser forward: tiny / base / midi --------------------------------------
    for name, n, frames, wseed in (("tiny", 3, 32, 11), ("base", 2, 64, 12), ("midi", 2, 48, 13)):
        cfg = config.get_config(name).denoiser
        net = ref_denoiser(name, wseed)
        x, t, cond, tc = inputs(cfg, n, frames, 100 + wseed)
        taps = {}
        hooks = [
            blk.register_forward_hook(lambda m, i, o, k=f"h{j + 1}": taps.__setitem__(k, o))
            for j, blk in enumerate(net.denoiser_trans_block.decoder_blocks)
        ]
        y = net(x, time=t.reshape(n, 1, 1), cond=cond, time_cond=tc)
        for h in hooks:
            h.remove()
        save(f"denoiser_{name}", weight_seed=wseed, x=x, time=t, cond=cond, time_cond=tc, out=y,
             h1=taps["h1"], h6=taps["h6"])

    # ---- band mask (row counts are the parity trap of SURVEY.md A.2) -----------------
    m8 = ns.transformerv2.combined_sliding_chunkwise_mask(64, 4, 8)
    m16 = ns.transformerv2.combined_sliding_chunkwise_mask(64, 4, 16)
    m_ragged = ns.transformerv2.combined_sliding_chunkwise_mask(30, 4, 8)
    save("band_mask", w8=m8.to(torch.uint8), w16=m16.to(torch.uint8), w8_len30=m_ragged.to(torch.uint8))

    # ---- model_forward + sample (RectifiedFlow) -------------------------------------
    for name, B, frames, steps, wseed in (("tiny", 2, 32, 4, 21), ("base", 1, 32, 2, 22)):
        cfg = config.get_config(name).denoiser
        net = ref_denoiser(name, wseed)
        rf = ns.model.RectifiedFlow(net=net, sr=44100, encoder=None, encoder_time=None,
                                    classifier=None, drop_value=-4.0, device="cpu").eval()
        x0, _, cond, tc = inputs(cfg, B, frames, 200 + wseed)
        t = torch.full((B, 1, 1), 0.3)
        dx = rf.model_forward(x0, t, cond, tc, guidance_timbre=2.0, guidance_structure=1.0)
        out = rf.sample(x0, cond, tc, nb_steps=steps, guidance_timbre=2.0, guidance_structure=1.0)
        save(f"sample_{name}", weight_seed=wseed, x0=x0, cond=cond, time_cond=tc, nb_steps=steps,
             guidance_timbre=2.0, guidance_structure=1.0, t_model_forward=0.3, dx=dx, out=out)

    # ---- codec: reduced topology and the real baseAE ---------------------------------
    for tag, acfg, B, samples, wseed in (("small", config.small_autoencoder(), 2, 4096, 31),
                                         ("base", config.base_autoencoder(), 1, 16384, 32)):
        ae = ns.ae.AutoEncoder(bottleneck=ns.ae.ReluBottleneck(sigma=0.01, scale=3),
                               in_channels=acfg.in_channels, channels=acfg.channels,
                               z_channels=acfg.z_channels, pqmf_bands=acfg.pqmf_bands,
                               multipliers=acfg.multipliers, factors=acfg.factors,
                               dilations=acfg.dilations, kernel_size=acfg.kernel_size,
                               decoder_ratio=acfg.decoder_ratio, use_loudness=acfg.use_loudness,
                               use_norm=True, use_noise=False).eval()
        ae.load_state_dict(synth.autoencoder_state_dict(acfg, wseed), strict=True)
        audio = synth.synth_audio(B, samples, seed=7)
        z, _, = ae.encode(audio)
        multiband = ae.pqmf(audio)
        g = torch.Generator().manual_seed(5)
        z_in = torch.randn(z.shape, generator=g)
        y = ae.decode(z_in)
        rec = ae.decode(z)
        save(f"codec_{tag}", weight_seed=wseed, audio=audio, multiband=multiband, z=z, z_in=z_in,
             decoded=y, reconstructed=rec)

    # ---- PQMF filters as the reference designs them ------------------------------------
    pq = ns.pqmf.CachedPQMF(attenuation=100, n_band=16)
    save("pqmf_16band_100dB", h=pq.h, hk=pq.hk, forward=pq.forward_conv.weight,
         inverse=pq.inverse_conv.weight)

    # ---- structure encoder -----------------------------------------------------------------
    for name, wseed in (("tiny", 41), ("base", 42)):
        ecfg = config.get_config(name).structure_encoder
        enc = R.build_encoder1d(name)
        enc.load_state_dict(synth.encoder1d_state_dict(ecfg, wseed), strict=True)
        g = torch.Generator().manual_seed(300 + wseed)
        z = torch.randn(2, ecfg.in_size, 40, generator=g)
        save(f"encoder1d_{name}", weight_seed=wseed, z=z, out=enc(z))

    # ---- timbre encoder (ECAPA-TDNN) ------------------------------------------------------------
    for name, wseed in (("tiny", 51), ("base", 52)):
        ecfg = config.get_config(name).timbre_encoder
        enc = R.build_ecapa(name)
        enc.load_state_dict(synth.ecapa_state_dict(ecfg, wseed), strict=True)
        g = torch.Generator().manual_seed(400 + wseed)
        z = torch.randn(2, ecfg.in_size, 40, generator=g)
        save(f"ecapa_{name}", weight_seed=wseed, z=z, out=enc(z))


if __name__ == "__main__":
    main()
