"""One AutoEncoder.encode + decode at B streams of 524288 samples (for ncu captures: every kernel launches once).
    python scripts/codec_once.py [B] [precision]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from after_b200 import config, synth
from after_b200.engine import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
acfg = config.base_autoencoder()
eng = Engine(autoencoder=acfg, autoencoder_state=synth.autoencoder_state_dict(acfg, 0), precision=prec, max_batch=B, max_samples=524288)
audio = synth.synth_audio(B, 524288).cuda()
z = eng.ae_encode(audio)
y = eng.ae_decode(z)
torch.cuda.synchronize()
print("ok", tuple(z.shape), tuple(y.shape), eng.launch_count)
eng.close()
