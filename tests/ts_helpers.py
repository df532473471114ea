"""Test helper: a TorchScript file shaped like the codec export of ``after_scripts/export_autoencoder.py`` (an nn_tilde
wrapper with the ``AutoEncoder`` under ``.model`` plus its own buffers), built from a state dict alone -- the reference
module classes are not importable on the GPU box.  Only the *state* matters to ``codec_state_from_torchscript``."""
import torch
import torch.nn as nn


class _Node(nn.Module):
    """Container without compute: scripted as an empty module that still carries its parameters / buffers."""

    def __init__(self):
        super().__init__()


def module_from_state(sd):
    root = _Node()
    for key, v in sd.items():
        parts = key.split(".")
        m = root
        for p in parts[:-1]:
            if not hasattr(m, p):
                m.add_module(p, _Node())
            m = getattr(m, p)
        if v.is_floating_point():
            m.register_parameter(parts[-1], nn.Parameter(v.clone(), requires_grad=False))
        else:
            m.register_buffer(parts[-1], v.clone())
    return root


class ExportWrapper(nn.Module):
    """``AE_notcausal``-shaped: model + overlap-add buffers (export_autoencoder.py:16-66)."""

    def __init__(self, model: nn.Module, latent_size: int, ratio: int, n_fade: int = 4):
        super().__init__()
        self.model = model
        self.register_buffer("out_buffer", torch.zeros(4, 1, ratio * n_fade))
        self.register_buffer("z_buffer", torch.zeros(4, latent_size, n_fade))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x


def save_codec_ts(sd, path, latent_size, ratio):
    ts = torch.jit.script(ExportWrapper(module_from_state(sd), latent_size, ratio))
    ts.save(path)
    return path
