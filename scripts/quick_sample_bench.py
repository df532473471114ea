"""Quick device-side timing of RectifiedFlow.sample (base, B streams, T=256)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from after_b200 import config, synth
from after_b200.engine import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
mc = config.get_config("base")
sd = synth.denoiser_state_dict(mc.denoiser, 0)
x0, cond, tc = (t.cuda() for t in synth.synth_inputs(B, mc.denoiser))
flop = 3 * B * 7.36e9 * steps
for prec in sys.argv[3:] or ["fp32", "bf16", "fp32_simt"]:
    eng = Engine(model=mc, denoiser_state=sd, precision=prec, max_batch=B, max_steps=steps)
    for _ in range(3):
        out = eng.sample(x0, cond, tc, steps, 2.0, 1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        e0.record()
        out = eng.sample(x0, cond, tc, steps, 2.0, 1.0)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    print(f"{prec}: B={B} steps={steps} best {ms:.2f} ms  median {sorted(ts)[2]:.2f} ms -> {steps / ms * 1e3:.1f} steps/s, "
          f"{flop / ms / 1e9:.1f} TFLOP/s algorithmic, launches/call={eng.launch_count}")
    eng.close()
