"""Latency of one streaming block (after_sample_stream: B = 1, 4 frames, nb_steps Euler steps through the per-step KV caches)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from after_b200 import config, synth
from after_b200.engine import Engine

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
name = sys.argv[3] if len(sys.argv) > 3 else "base"
mc = config.get_config(name)
sd = synth.denoiser_state_dict(mc.denoiser, 0)
eng = Engine(model=mc, denoiser_state=sd, precision=prec, max_batch=1, max_steps=steps, seq_len=4,
             max_cache_size=mc.denoiser.local_attention_size)
x0, cond, tc = (t.cuda() for t in synth.synth_inputs(1, mc.denoiser, frames=4))
for _ in range(5):
    eng.sample_stream(x0, cond, tc, steps, 2.0, 1.0)
torch.cuda.synchronize()
ts = []
for _ in range(30):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.sample_stream(x0, cond, tc, steps, 2.0, 1.0); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
print(f"{name} {prec}: streaming block of 4 frames, {steps} steps: median {ts[15]:.3f} ms, best {ts[0]:.3f} ms "
      f"({ts[15] / steps * 1e3:.0f} us per step; block = 185.8 ms of audio)")
eng.close()
