"""Latency of the whole exported Streamer.forward on one 8192-sample buffer (two streaming codec encodes, structure encoder with
cached convs, ECAPA on the rolling timbre buffer, the streamed sampler, overlap-add decode), and of its stages.
    python scripts/quick_streamer_bench.py [nb_steps] [precision] [model]"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from after_b200 import config, synth
from after_b200.engine import Engine
from after_b200.streamer import Streamer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
name = sys.argv[3] if len(sys.argv) > 3 else "base"
mc = config.get_config(name)
acfg = config.base_autoencoder()
eng = Engine(model=mc, autoencoder=acfg, denoiser_state=synth.denoiser_state_dict(mc.denoiser, 0),
             autoencoder_state=synth.autoencoder_state_dict(acfg, 0), structure_state=synth.encoder1d_state_dict(mc.structure_encoder, 0),
             timbre_state=synth.ecapa_state_dict(mc.timbre_encoder, 0), precision=prec, max_batch=1, max_steps=steps, seq_len=64,
             max_samples=64 * 2048, max_cache_size=mc.denoiser.local_attention_size, stream_slots=2, stream_max_frames=4)
st = Streamer(eng, n_signal_timbre=64, chunk_size=4)
st.set_nb_steps(steps); st.set_guidance_timbre(2.0); st.set_guidance_structure(1.0)
buf = torch.cat([synth.synth_audio(1, 8192, seed=3), synth.synth_audio(1, 8192, seed=4)], 1).cuda()
noise = torch.randn(1, 64, 4, device="cuda")
z = torch.randn(1, 64, 4, device="cuda")


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


l0 = eng.launch_count
st.forward(buf, noise=noise)
torch.cuda.synchronize()
print(f"{name} {prec} {steps} steps: kernels per Streamer.forward = {eng.launch_count - l0}")
print(f"  forward          {timeit(lambda: st.forward(buf, noise=noise)):.3f} ms  (buffer = 185.8 ms of audio)")
print(f"  ae_encode_stream {timeit(lambda: eng.ae_encode_stream(0, buf[:, :1].contiguous())):.3f} ms")
print(f"  structure_stream {timeit(lambda: eng.structure_encode_stream(0, z)):.3f} ms")
print(f"  timbre_encode    {timeit(lambda: eng.timbre_encode(st.previous_timbre[:1].contiguous())):.3f} ms")
print(f"  sample_stream    {timeit(lambda: eng.sample_stream(noise, torch.zeros(1, 6, device='cuda'), torch.zeros(1, 12, 4, device='cuda'), steps, 2.0, 1.0)):.3f} ms")
print(f"  ae_decode_stream {timeit(lambda: eng.ae_decode_stream(0, z)):.3f} ms")
eng.close()
