"""Ingest a reference training run: ``config.gin`` + ``checkpoint<step>_EMA.pt`` -> what an ``Engine`` takes.

Mirrors what ``after_scripts/export.py:52-101`` does with gin and ``load_state_dict``: the run folder holds the operative
gin config written at training start (``after/diffusion/model.py:264-265``) and checkpoints
``{"model_state": RectifiedFlow.state_dict() minus emb_model.*, "opt_state": ...}`` (``model.py:144-176``).  The
``RectifiedFlow`` sub-modules are ``net`` (DenoiserV2), ``encoder`` (ECAPATDNN, timbre), ``encoder_time`` (Encoder1D,
structure) and ``classifier`` (training only), so the state dict is split on those prefixes.

gin itself is not a dependency: the few constructs the AFTER configs use are parsed here -- macros (``NAME = value``,
``%NAME`` references), bindings in one-line (``scope/mod.Class.param = value``, the operative-config form) and block
form (``scope/mod.Class:`` + indented ``param = value``), Python literals spanning several lines, and ``@Class()``
references (kept as strings).  Host-side only; nothing here touches the GPU.
"""
from __future__ import annotations

import ast
import glob
import os
import re
from typing import Any, Dict, Optional, Tuple

import torch

from .config import AutoEncoderConfig, DenoiserConfig, EcapaConfig, Encoder1DConfig, ModelConfig

Bindings = Dict[Tuple[str, str], Dict[str, Any]]  # (scope, configurable's last name) -> {param: value}


class _Ref(str):
    """``@configurable()`` / ``@scope/configurable`` reference, kept verbatim."""


def _strip_comment(line: str) -> str:
    out, quote = [], None
    for ch in line:
        if quote:
            if ch == quote:
                quote = None
        elif ch in "'\"":
            quote = ch
        elif ch == "#":
            break
        out.append(ch)
    return "".join(out).rstrip()


def _balanced(text: str) -> bool:
    depth, quote = 0, None
    for ch in text:
        if quote:
            if ch == quote:
                quote = None
        elif ch in "'\"":
            quote = ch
        elif ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
    return depth <= 0


def _value(text: str, macros: Dict[str, Any]) -> Any:
    text = text.strip()
    if text.startswith("@"):
        return _Ref(text)

    def sub(m):
        name = m.group(1)
        if name not in macros:
            raise KeyError(f"gin macro %{name} is not defined")
        return repr(macros[name])

    text = re.sub(r"%([A-Za-z_][A-Za-z0-9_.]*)", sub, text)
    if re.search(r"@[A-Za-z_]", text):  # references inside containers: keep the whole value as a string
        return _Ref(text)
    try:
        return ast.literal_eval(text)
    except (ValueError, SyntaxError) as e:
        raise ValueError(f"cannot parse gin value {text!r}") from e


def parse_gin(text: str, macros: Optional[Dict[str, Any]] = None) -> Tuple[Dict[str, Any], Bindings]:
    """Returns (macros, bindings).  ``macros`` passed in pre-define / override file macros that are ``None``
    (``IN_SIZE`` / ``N_SIGNAL`` are bound at run time by the reference, ``after_scripts/train.py:80-86``)."""
    overrides = dict(macros or {})
    macros = {}
    bindings: Bindings = {}
    block: Optional[Tuple[str, str]] = None
    # gin.operative_config_str() wraps every binding longer than 80 columns as ``selector = \`` + the value indented on
    # the next line(s) (after/diffusion/model.py:264-265 writes config.gin that way): join backslash continuations first.
    # The joined line keeps the indentation of its FIRST physical line, which is what decides block membership below.
    lines = []
    pending = None
    for raw in text.splitlines():
        body = _strip_comment(raw)
        if pending is not None:
            body = pending + " " + body.strip()
            pending = None
        if body.rstrip().endswith("\\"):
            pending = body.rstrip()[:-1].rstrip()
            continue
        lines.append(body)
    if pending is not None:
        lines.append(pending)
    i = 0
    while i < len(lines):
        raw = lines[i]
        i += 1
        line = _strip_comment(raw)
        if not line.strip():
            continue
        while not _balanced(line) and i < len(lines):  # literals spanning several lines
            line += " " + _strip_comment(lines[i]).strip()
            i += 1
        stripped = line.strip()
        if stripped.startswith(("import ", "from ", "include ")):
            block = None
            continue
        indented = raw[:1] in " \t"
        if not indented:
            block = None
        if stripped.endswith(":") and "=" not in stripped:  # block header  scope/module.Class:
            target = stripped[:-1].strip()
            scope, _, name = target.rpartition("/")
            block = (scope, name.split(".")[-1])
            bindings.setdefault(block, {})
            continue
        if "=" not in stripped:
            continue
        lhs, _, rhs = stripped.partition("=")
        lhs = lhs.strip()
        if indented and block is not None:
            bindings[block][lhs] = _value(rhs, macros)
            continue
        if re.fullmatch(r"[A-Za-z_][A-Za-z0-9_]*", lhs):  # macro
            v = _value(rhs, macros)
            if v is None and lhs in overrides:
                v = overrides[lhs]
            macros[lhs] = overrides.get(lhs, v) if v is None else v
            continue
        scope, _, dotted = lhs.rpartition("/")
        parts = dotted.split(".")
        if len(parts) < 2:
            continue
        key = (scope, parts[-2])
        bindings.setdefault(key, {})[parts[-1]] = _value(rhs, macros)
    for k, v in overrides.items():
        macros.setdefault(k, v)
    return macros, bindings


def _find(bindings: Bindings, name: str, scope: Optional[str] = None) -> Dict[str, Any]:
    hits = [(k, v) for k, v in bindings.items() if k[1] == name and (scope is None or k[0] == scope)]
    if not hits:
        return {}
    return hits[0][1]


def model_config_from_gin(text: str, in_size: Optional[int] = None, n_signal: Optional[int] = None,
                          name: str = "run") -> ModelConfig:
    """Build a ``ModelConfig`` from gin text (a run's operative ``config.gin`` or a static ``base.gin``-style file)."""
    pre = {}
    if in_size is not None:
        pre["IN_SIZE"] = in_size
    if n_signal is not None:
        pre["N_SIGNAL"] = n_signal
    macros, b = parse_gin(text, pre)
    d = _find(b, "DenoiserV2")
    if not d:
        raise ValueError("config binds no DenoiserV2 (the v1 Denoiser / UNET1D nets are out of scope)")
    # reference defaults (transformerv2.py:463-476) are pos_emb_type='learnable', causal=False: a config that does not bind
    # them asks for a model this path does not implement, so it is rejected rather than silently run as rotary/causal
    if d.get("pos_emb_type", "learnable") != "rotary" or not d.get("causal", False):
        raise ValueError("only the causal, rotary DenoiserV2 of the shipped configs is supported "
                         "(config must bind DenoiserV2.pos_emb_type = 'rotary' and DenoiserV2.causal = True)")
    base = DenoiserConfig()

    def pick(key, default):
        v = d.get(key, default)
        return default if v is None else v

    den = DenoiserConfig(
        n_channels=pick("n_channels", base.n_channels), seq_len=pick("seq_len", base.seq_len),
        embed_dim=pick("embed_dim", 256), cond_dim=pick("cond_dim", 64), noise_embed_dims=pick("noise_embed_dims", 128),
        n_layers=pick("n_layers", 6), mlp_multiplier=pick("mlp_multiplier", 2), tcond_dim=pick("tcond_dim", 0),
        local_attention_size=pick("local_attention_size", base.local_attention_size),
        attention_chunk_size=pick("attention_chunk_size", 4))
    e = _find(b, "ECAPATDNN")
    tim = EcapaConfig(
        in_size=e.get("in_size") or den.n_channels, channels=list(e.get("channels", EcapaConfig().channels)),
        kernel_sizes=list(e.get("kernel_sizes", [3, 3, 3, 3])), dilations=list(e.get("dilations", [1, 1, 1, 1])),
        attention_channels=e.get("attention_channels", 128), res2net_scale=e.get("res2net_scale", 8),
        se_channels=e.get("se_channels", 128), out_dim=e.get("out_dim", den.cond_dim),
        global_context=bool(e.get("global_context", True)), use_tanh=bool(e.get("use_tanh", False)))
    base_b = _find(b, "Base") or _find(b, "RectifiedFlow")
    se = None
    if base_b.get("encoder_time", "@") is not None:
        s = _find(b, "Encoder1D", "encoder_time")
        if s:
            pad = _find(b, "get_padding", "encoder_time").get("mode", "centered")
            se = Encoder1DConfig(in_size=s.get("in_size") or den.n_channels, channels=list(s["channels"]),
                                 ratios=list(s.get("ratios", [1, 1, 1, 1])), kernel_size=s.get("kernel_size", 5),
                                 causal=(pad == "causal"), use_tanh=bool(s.get("use_tanh", False)))
    return ModelConfig(name, den, se, tim, drop_value=float(base_b.get("drop_value", -4.0)), sr=int(macros.get("SR", 44100)))


_PREFIXES = {"denoiser_state": "net.", "timbre_state": "encoder.", "structure_state": "encoder_time."}


def split_model_state(model_state: Dict[str, torch.Tensor]) -> Dict[str, Optional[Dict[str, torch.Tensor]]]:
    """``RectifiedFlow.state_dict()`` -> per-module state dicts (keys as the sub-module itself names them)."""
    out: Dict[str, Optional[Dict[str, torch.Tensor]]] = {}
    for name, prefix in _PREFIXES.items():
        sd = {k[len(prefix):]: v for k, v in model_state.items() if k.startswith(prefix)}
        out[name] = sd or None
    return out


def find_checkpoint(folder: str, step: Optional[int] = None) -> str:
    """Latest (or given) ``checkpoint<step>_EMA.pt`` of a run, like ``after_scripts/export.py:52-66``."""
    if step is not None:
        path = os.path.join(folder, f"checkpoint{step}_EMA.pt")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        return path
    best = None
    for p in glob.glob(os.path.join(folder, "checkpoint*_EMA.pt")):
        m = re.search(r"checkpoint(\d+)_EMA\.pt$", p)
        if m and (best is None or int(m.group(1)) > best[0]):
            best = (int(m.group(1)), p)
    if best is None:
        raise FileNotFoundError(f"no checkpoint*_EMA.pt in {folder}")
    return best[1]


def load_run(folder: str, step: Optional[int] = None, in_size: Optional[int] = None, n_signal: Optional[int] = None) -> Dict[str, Any]:
    """Read a training run folder.  Returns ``{"model": ModelConfig, "denoiser_state", "timbre_state",
    "structure_state", "checkpoint"}`` -- pass the first four to ``Engine(...)`` (plus a codec, see below)."""
    ckpt = find_checkpoint(folder, step)
    state = torch.load(ckpt, map_location="cpu", weights_only=True)["model_state"]
    parts = split_model_state(state)
    if parts["denoiser_state"] is None:
        raise ValueError(f"{ckpt} holds no net.* tensors")
    if in_size is None:  # IN_SIZE is bound at run time from the codec; the checkpoint knows it
        w = parts["denoiser_state"].get("denoiser_trans_block.patchify_and_embed.1.weight")
        in_size = int(w.shape[1]) if w is not None else None
    with open(os.path.join(folder, "config.gin")) as fh:
        model = model_config_from_gin(fh.read(), in_size=in_size, n_signal=n_signal, name=os.path.basename(os.path.normpath(folder)))
    if model.structure_encoder is not None and parts["structure_state"] is None:
        raise ValueError("config binds encoder_time but the checkpoint has no encoder_time.* tensors")
    return {"model": model, "checkpoint": ckpt, **parts}


def codec_state_from_torchscript(path: str) -> Dict[str, torch.Tensor]:
    """State dict of the ``AutoEncoder`` inside an exported codec ``.ts`` (``after_scripts/export_autoencoder.py``): the
    wrapper's prefix in front of ``pqmf.`` / ``encoder.`` / ``decoder.`` is dropped."""
    sd = torch.jit.load(path, map_location="cpu").state_dict()
    return strip_codec_prefix(sd)


def strip_codec_prefix(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    prefix = None
    for k in sd:
        m = re.match(r"^(.*?)(encoder\.net\.0\.)", k)
        if m:
            prefix = m.group(1)
            break
    if prefix is None:
        raise ValueError("no AutoEncoder tensors (encoder.net.0.*) in this state dict")
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix) and k[len(prefix):].split(".")[0] in ("pqmf", "encoder", "decoder")}


def autoencoder_config_from_state(sd: Dict[str, torch.Tensor], dilations=(1, 3, 9)) -> AutoEncoderConfig:
    """Infer the codec topology from tensor shapes (key layout of SimpleNetsStream.py:400-459, 552-651).  Dilations leave
    no trace in the weights: pass them if they differ from baseAE.gin's [1, 3, 9]."""
    bands = int(sd["pqmf.forward_conv.weight"].shape[0]) if "pqmf.forward_conv.weight" in sd else 1
    stages = sorted({int(m.group(1)) for k in sd for m in [re.match(r"encoder\.net\.(\d+)\.net\.\d+\.", k)] if m})
    first = sd["encoder.net.0.net.branches.0.0.net.2.weight_v"]  # ResnetBlock(in_channels -> channels * multipliers[0])
    in_channels = int(first.shape[1])
    n = len(stages)
    nb = len({int(m.group(1)) for k in sd for m in [re.match(rf"encoder\.net\.{stages[0]}\.net\.(\d+)\.net\.branches", k)] if m})
    outs, factors = [int(first.shape[0])], []
    for i in stages:
        w = sd[f"encoder.net.{i}.net.{nb + 1}.weight_v"]  # strided conv: (C_out, C_in, 2 f)
        outs.append(int(w.shape[0]))
        factors.append(int(w.shape[2]) // 2)
    channels = outs[0]  # multipliers[0] == 1 in every shipped config
    multipliers = [o // channels for o in outs]
    z = int(sd[f"encoder.net.{n + 2}.weight_v"].shape[0])
    dec_in = int(sd["decoder.net.0.weight_v"].shape[0])
    ratio = dec_in / float(channels * multipliers[-1])
    last = [k for k in sd if k.startswith("decoder.synth.branches.0.net.1.net.2.weight_v")]
    use_loudness = bool(last) and int(sd[last[0]].shape[0]) == 2 * in_channels
    return AutoEncoderConfig(in_channels=in_channels, channels=channels, z_channels=z, pqmf_bands=bands, multipliers=multipliers,
                             factors=factors, dilations=list(dilations), kernel_size=int(first.shape[2]), decoder_ratio=ratio,
                             use_loudness=use_loudness, num_blocks=nb)
