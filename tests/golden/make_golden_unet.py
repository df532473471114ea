"""Golden fixtures for the UNET1D conv denoiser, produced by the UNMODIFIED reference
module (after/diffusion/networks/unet1d.py).  Authoring container only:

    python tests/golden/make_golden_unet.py
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

from unet_cases import CASES  # noqa: E402


@torch.no_grad()
def main():
    R.install()
    import importlib.util
    # unet1d.py imports .blocks (einops only): load both leaf modules under the namespace stubs
    for name, rel in (("after.diffusion.networks.blocks", "after/diffusion/networks/blocks.py"),
                      ("after.diffusion.networks.unet1d", "after/diffusion/networks/unet1d.py")):
        spec = importlib.util.spec_from_file_location(name, os.path.join(R.REF_ROOT, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    unet1d = sys.modules["after.diffusion.networks.unet1d"]
    for tag, (cfg, wseed) in CASES.items():
        net = unet1d.UNET1D(in_size=cfg.in_size, out_size=cfg.out_size, channels=list(cfg.channels), ratios=list(cfg.ratios),
                            kernel_size=cfg.kernel_size, time_channels=cfg.time_channels,
                            time_cond_in_channels=cfg.time_cond_in_channels, time_cond_channels=cfg.time_cond_channels,
                            cond_channels=cfg.cond_channels, n_attn_layers=cfg.n_attn_layers,
                            use_res_last=cfg.use_res_last).eval()
        net.load_state_dict(synth.unet_state_dict(cfg, wseed), strict=True)
        g = torch.Generator().manual_seed(700 + wseed)
        n, frames = 2, 32
        x = torch.randn(n, cfg.in_size, frames, generator=g)
        t = torch.rand(n, generator=g)
        cond = torch.randn(n, cfg.cond_channels, generator=g) if cfg.cond_channels else None
        tc = torch.randn(n, cfg.time_cond_in_channels, frames, generator=g)
        y = net(x, time=t, time_cond=tc, cond=cond)
        save(f"unet_{tag}", weight_seed=wseed, x=x, time=t, cond=cond if cond is not None else torch.zeros(n, 0), time_cond=tc, out=y)


if __name__ == "__main__":
    main()
