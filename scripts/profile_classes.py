"""Per-kernel-class device time of one sample() call (library profiler: CUDA events around each launch)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from after_b200 import config, synth
from after_b200.engine import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
prec = sys.argv[3] if len(sys.argv) > 3 else "fp32"
mc = config.get_config("base")
sd = synth.denoiser_state_dict(mc.denoiser, 0)
x0, cond, tc = (t.cuda() for t in synth.synth_inputs(B, mc.denoiser))
eng = Engine(model=mc, denoiser_state=sd, precision=prec, max_batch=B, max_steps=steps)
eng.sample(x0, cond, tc, steps, 2.0, 1.0)
eng.profile(True)
eng.sample(x0, cond, tc, steps, 2.0, 1.0)
for k, v in eng.profile_read().items():
    if v["launches"]:
        print(f"{k:14s} n={v['launches']:5d} total {v['ms']:8.2f} ms avg {v['ms'] / v['launches'] * 1e3:7.1f} us  "
              f"{v['flops'] / v['ms'] / 1e9 if v['ms'] else 0:8.1f} TFLOP/s  {v['bytes'] / v['ms'] / 1e6 if v['ms'] else 0:8.1f} GB/s")
eng.profile(False)
eng.close()
