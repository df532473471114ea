"""Sampler throughput across the BASELINE configs (tiny B=1 x10 steps, base B=8/64 x50, midi B=32 x50), one GPU."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from after_b200 import config, synth
from after_b200.engine import Engine

cases = [("tiny", 1, 10, "fp32", 0), ("base", 1, 50, "fp32", 0), ("base", 8, 50, "fp32", 0), ("base", 8, 50, "bf16", 0),
         ("base", 64, 50, "bf16", 0), ("base", 64, 50, "fp32", 0), ("midi", 8, 50, "fp32", 1), ("midi", 32, 50, "fp32", 1)]
for name, B, steps, prec, variant in cases:
    mc = config.get_config(name)
    sd = synth.denoiser_state_dict(mc.denoiser, 0)
    x0, cond, tc = (t.cuda() for t in synth.synth_inputs(B, mc.denoiser))
    eng = Engine(model=mc, denoiser_state=sd, precision=prec, max_batch=B, max_steps=steps)
    clamp = 0.1 if variant else 0.01
    for _ in range(2):
        out = eng.sample(x0, cond, tc, steps, 2.0, 3.0 if variant else 1.0, cfg_variant=variant, clamp=clamp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(3):
        e0.record()
        out = eng.sample(x0, cond, tc, steps, 2.0, 3.0 if variant else 1.0, cfg_variant=variant, clamp=clamp)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    print(json.dumps({"model": name, "streams": B, "nb_steps": steps, "precision": prec, "ms": ms, "steps_per_s": steps / ms * 1e3,
                      "sequence_steps_per_s": steps * B / ms * 1e3, "finite": bool(torch.isfinite(out).all()),
                      "device_MB": eng.device_bytes / 1e6}))
    eng.close()
