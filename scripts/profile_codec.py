"""Per-kernel-class device time of one encode + one decode (library profiler)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from after_b200 import config, synth
from after_b200.engine import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
acfg = config.base_autoencoder()
sd = synth.autoencoder_state_dict(acfg, 0)
audio = synth.synth_audio(B, 524288).cuda()
eng = Engine(autoencoder=acfg, autoencoder_state=sd, precision=prec, max_batch=B, max_samples=524288)
z = eng.ae_encode(audio)
y = eng.ae_decode(z)
for name, fn in (("encode", lambda: eng.ae_encode(audio)), ("decode", lambda: eng.ae_decode(z))):
    eng.profile(True)
    fn()
    print(name)
    for k, v in eng.profile_read().items():
        if v["launches"]:
            print(f"  {k:14s} n={v['launches']:5d} total {v['ms']:8.3f} ms avg {v['ms'] / v['launches'] * 1e3:7.1f} us  "
                  f"{v['flops'] / v['ms'] / 1e9 if v['ms'] else 0:8.1f} TFLOP/s  {v['bytes'] / v['ms'] / 1e6 if v['ms'] else 0:8.1f} GB/s")
    eng.profile(False)
eng.close()
