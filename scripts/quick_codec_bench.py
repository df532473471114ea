"""Quick device-side timing of AutoEncoder.encode / decode (baseAE, B streams of 524288 samples)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from after_b200 import config, synth
from after_b200.engine import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S = 524288
acfg = config.base_autoencoder()
sd = synth.autoencoder_state_dict(acfg, 0)
audio = synth.synth_audio(B, S).cuda()
for prec in sys.argv[2:] or ["fp32", "bf16"]:
    eng = Engine(autoencoder=acfg, autoencoder_state=sd, precision=prec, max_batch=B, max_samples=S)
    z = eng.ae_encode(audio)
    y = eng.ae_decode(z)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    te, td = [], []
    for _ in range(5):
        e0.record()
        z = eng.ae_encode(audio)
        e1.record()
        y = eng.ae_decode(z)
        e2.record()
        torch.cuda.synchronize()
        te.append(e0.elapsed_time(e1))
        td.append(e1.elapsed_time(e2))
    print(f"{prec}: B={B} encode best {min(te):.2f} ms ({B * 45.1e9 / min(te) / 1e9:.1f} TFLOP/s, {B * 0.475 / min(te) * 1e3:.0f} GB/s alg) "
          f"decode best {min(td):.2f} ms ({B * 95.3e9 / min(td) / 1e9:.1f} TFLOP/s, {B * 0.620 / min(td) * 1e3:.0f} GB/s alg) "
          f"launches so far {eng.launch_count}")
    eng.close()
