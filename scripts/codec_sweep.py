"""BASELINE configs[4]: AutoEncoder encode+decode only, sweep of streams per GPU, reporting achieved GB/s against the
algorithmic byte model of SURVEY.md section 8d (each conv reads its input once and writes its output once:
475 MB encode / 620 MB decode per 524288-sample chunk in fp32) and TFLOP/s (45.1 / 95.3 GFLOP per chunk)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from after_b200 import config, synth
from after_b200.engine import Engine

S = 524288
acfg = config.base_autoencoder()
sd = synth.autoencoder_state_dict(acfg, 0)
peak = 6558.1
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
rows = []
for prec in sys.argv[1:] or ["fp32", "bf16"]:
    for B in (1, 2, 4, 8, 16):
        eng = Engine(autoencoder=acfg, autoencoder_state=sd, precision=prec, max_batch=B, max_samples=S)
        audio = synth.synth_audio(B, S).cuda()
        z = eng.ae_encode(audio)
        eng.ae_decode(z)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        te, td = [], []
        for _ in range(5):
            ev[0].record(); z = eng.ae_encode(audio); ev[1].record(); y = eng.ae_decode(z); ev[2].record()
            torch.cuda.synchronize()
            te.append(ev[0].elapsed_time(ev[1])); td.append(ev[1].elapsed_time(ev[2]))
        e, d = min(te), min(td)
        row = {"precision": prec, "streams": B, "encode_ms": e, "decode_ms": d,
               "encode_GBps": B * 0.475 / e * 1e3, "decode_GBps": B * 0.620 / d * 1e3,
               "encode_frac_hbm": B * 0.475 / e * 1e3 / peak, "decode_frac_hbm": B * 0.620 / d * 1e3 / peak,
               "encode_TFLOPs": B * 45.1e9 / e / 1e9, "decode_TFLOPs": B * 95.3e9 / d / 1e9,
               "rtf_encode_decode": B * (S / 44100) / ((e + d) / 1e3)}
        rows.append(row)
        print(json.dumps(row))
        eng.close()
