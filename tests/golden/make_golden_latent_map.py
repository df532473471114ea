"""Golden fixture for Streamer.latent2map / map2latent (after_scripts/export.py:494-508), produced with the UNMODIFIED
reference projection module (after/diffusion/latent_plot.py:20-37, SmallAutoencoder).  Authoring container only:

    python tests/golden/make_golden_latent_map.py

latent_plot.py imports plotting / sklearn packages at module level that the projection itself never touches; the ones
missing here are stubbed before the file is executed.  The two exported methods live inside a class that export.py
defines in its main(), so their four lines each are applied here verbatim to the reference module's encoder / decoder."""
import importlib.util
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from make_golden import save  # noqa: E402

REF_ROOT = os.environ.get("AFTER_REFERENCE", "/root/reference")


def _stub(name, **attrs):
    if name in sys.modules:
        return
    try:
        importlib.import_module(name)
        return
    except Exception:
        pass
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    if "." in name:
        parent, _, child = name.rpartition(".")
        _stub(parent)
        setattr(sys.modules[parent], child, m)


@torch.no_grad()
def main():
    for name, attrs in (("matplotlib", {}), ("matplotlib.pyplot", {}), ("matplotlib.gridspec", {}), ("matplotlib.patches", {}),
                        ("matplotlib.cm", {}), ("matplotlib.colors", {"to_rgb": None}), ("sklearn", {}),
                        ("sklearn.preprocessing", {"LabelEncoder": None}), ("sklearn.model_selection", {"train_test_split": None}),
                        ("scipy.ndimage", {"gaussian_filter": None}), ("tqdm", {"tqdm": None})):
        _stub(name, **attrs)
    if not hasattr(sys.modules["matplotlib"], "cm"):
        sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    spec = importlib.util.spec_from_file_location("after_ref_latent_plot", os.path.join(REF_ROOT, "after/diffusion/latent_plot.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(11)
    project_model = mod.SmallAutoencoder(input_dim=6, latent_dim=2).eval()
    g = torch.Generator().manual_seed(12)
    latents = torch.randn(3, 6, 9, generator=g)
    maps = torch.randn(3, 2, 9, generator=g)
    # export.py:503-508 (latent2map)
    tdim = latents.shape[-1]
    l2m = project_model.encoder(latents.mean(-1)).unsqueeze(-1).repeat((1, 1, tdim))
    # export.py:496-501 (map2latent)
    tdim = maps.shape[-1]
    m2l = project_model.decoder(maps.mean(-1)).unsqueeze(-1).repeat((1, 1, tdim))
    sd = {"sd." + k: v for k, v in project_model.state_dict().items()}
    save("latent_map", latents=latents, maps=maps, latent2map=l2m, map2latent=m2l, **sd)


if __name__ == "__main__":
    main()
