"""Device-side timing of the codec (encode / decode) and of the full audio-to-audio chain (base, B streams, one chunk)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from after_b200 import config, synth
from after_b200.engine import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
CHUNK = 524288
mc, acfg = config.get_config("base"), config.base_autoencoder()
eng = Engine(model=mc, autoencoder=acfg, denoiser_state=synth.denoiser_state_dict(mc.denoiser, 0),
             autoencoder_state=synth.autoencoder_state_dict(acfg, 0), structure_state=synth.encoder1d_state_dict(mc.structure_encoder, 0),
             timbre_state=synth.ecapa_state_dict(mc.timbre_encoder, 0), precision=prec, max_batch=B, max_steps=50, max_samples=CHUNK)
a_s, a_t = synth.synth_audio(B, CHUNK, seed=7).cuda(), synth.synth_audio(B, CHUNK, seed=107).cuda()
x0 = synth.synth_inputs(B, mc.denoiser)[0].cuda()


def timeit(fn, n=7):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), out


enc_ms, z = timeit(lambda: eng.ae_encode(a_s))
dec_ms, y = timeit(lambda: eng.ae_decode(z))
st_ms, tcnd = timeit(lambda: eng.structure_encode(z))
tm_ms, cnd = timeit(lambda: eng.timbre_encode(z))
ch_ms, out = timeit(lambda: eng.generate(a_s, a_t, x0, 50, 2.0, 1.0), 5)
print(json.dumps({"streams": B, "precision": prec, "encode_ms": enc_ms, "decode_ms": dec_ms, "structure_ms": st_ms, "timbre_ms": tm_ms,
                  "chain_ms": ch_ms, "rtf": B * CHUNK / 44100 / (ch_ms / 1e3),
                  "checksum": [float(z.double().sum()), float(y.double().sum()), float(out.double().sum())]}))
eng.close()
