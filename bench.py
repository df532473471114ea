#!/usr/bin/env python
"""Benchmark of the AFTER latent-sampling hot path on B200 (contract: see the task brief / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp32|bf16] [--impl ours|reference]

Workload (BASELINE.json configs[1]): base audio-to-audio, batch = 8 streams per GPU, 50 Euler steps x 3-way CFG
over one 524288-sample chunk per stream (T = 256 latent frames), synthetic seeded weights and inputs.
One bench "step" = one ``RectifiedFlow.sample`` call (50 diffusion steps over the per-GPU batch).
``value`` = diffusion-steps/s summed over GPUs (each GPU integrates its own 8 streams: weak scaling), inputs
resident in HBM; ``e2e`` = the same through the host-buffer C-ABI entry (pinned host tensors, H2D + D2H inside);
``rtf`` = real-time factor of the full chain (2x encode, structure encoder, sample, decode).
``configs`` = the other BASELINE.json configurations at their per-GPU sizes, each with its own roofline block:
``config3_bf16`` (base, bf16, 8 streams/GPU), ``config4_midi`` (midi model + MIDI CFG layout, 8 streams/GPU),
``config5_codec`` (AutoEncoder encode + decode only, 1..16 streams/GPU, GB/s against the byte model of SURVEY.md 8d).
``checksum`` = CRC32 of the sampled latents per stream: streams are generated per-stream-seeded on the host, so the
CRCs of streams 0..7 must be identical for every --gpus N (N-independence of the shard + gather).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import zlib

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHUNK = 524288
SR = 44100
FLOP_PER_SEQ = {"base": 7.36e9, "tiny": 1.86e9, "midi": 7.75e9}
# codec byte / flop model per 524288-sample chunk: every conv reads its input once and writes its
# output once, everything else fused
AE_BYTES = {"encode": {"fp32": 475e6, "bf16": 238e6}, "decode": {"fp32": 620e6, "bf16": 310e6}}
AE_FLOPS = {"encode": 45.1e9, "decode": 95.3e9}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region.  The sampler is started BEFORE the warm-up (nvidia-smi
    takes a few hundred ms to produce its first row) and every row is stamped on arrival; ``mark()`` brackets the timed region
    and ``stop()`` reports the rows inside it (or, for a region shorter than one sampling period, the rows nearest to it)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

    def mark(self):
        if self.t0 is None:
            self.t0 = time.monotonic()
        else:
            self.t1 = time.monotonic()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t0 = self.t0 if self.t0 is not None else 0.0
        t1 = self.t1 if self.t1 is not None else time.monotonic()
        inside = [r for ts, r in self.rows if t0 - 0.03 <= ts <= t1 + 0.08]
        where = "timed region"
        if not inside and self.rows:  # region shorter than a sampling period: the three rows nearest to it
            mid = 0.5 * (t0 + t1)
            inside = [r for _, r in sorted(self.rows, key=lambda tr: abs(tr[0] - mid))[:3]]
            where = "nearest samples (timed region shorter than one sampling period)"
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": where}


def synth_setup(name, B_total):
    from after_b200 import config, synth
    mc = config.get_config(name)
    x0, cond, tc = synth.synth_inputs_per_stream(B_total, mc.denoiser, seed=1234)  # stream b independent of B_total
    return mc, x0, cond, tc


def cpu_reference_steps_per_s(name, B, nb_steps, repeats=1, threads=None):
    """The CPU restatement of the reference's sampler (oracle port, plain PyTorch CPU ops) on this host."""
    import torch
    from after_b200 import synth
    from oracle import after_oracle as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    mc, x0, cond, tc = synth_setup(name, B)
    sd = synth.denoiser_state_dict(mc.denoiser, 0)
    O.sample(sd, mc.denoiser, x0[:1], cond[:1], tc[:1], 1, 2.0, 1.0)  # warm the thread pool / allocator
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        O.sample(sd, mc.denoiser, x0, cond, tc, nb_steps, 2.0, 1.0)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return nb_steps / best, threads, best


def workload_config(args, world):
    B, NS = args.batch, args.nb_steps
    return {"workload": f"{args.model} audio-to-audio, batch={B}/GPU, {NS} Euler steps x 3-way CFG, 524288-sample chunk (T=256)",
            "batch_per_gpu": B, "global_batch": B * world, "nb_steps": NS, "frames": 256, "precision": args.precision,
            "parallelism": f"batch-shard x{world}, one all_gather of the latents" if world > 1 else "single GPU",
            "l2": "flushed (256 MiB write) before every timed step; per-step CUDA events",
            "one_bench_step": f"one RectifiedFlow.sample call = {NS} diffusion steps over {B} streams (x3 CFG rows)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_steps = 2  # bounded sample of the 50-step workload: every diffusion step is identical work
    import torch
    from after_b200 import synth
    from oracle import after_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    mc, x0, cond, tc = synth_setup(args.model, args.batch)
    sd = synth.denoiser_state_dict(mc.denoiser, 0)
    for _ in range(max(args.warmup, 1)):
        O.sample(sd, mc.denoiser, x0[:1], cond[:1], tc[:1], 1, 2.0, 1.0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.sample(sd, mc.denoiser, x0, cond, tc, sample_steps, 2.0, 1.0)
    dt = time.perf_counter() - t0
    v = sample_steps * args.steps / dt
    sample = f"{sample_steps} of {args.nb_steps} diffusion steps per bench step, B={args.batch}, T=256, all host threads"
    print(json.dumps({
        "impl": "reference", "metric": "diffusion-steps/sec", "value": v, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3 * args.nb_steps / sample_steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": workload_config(args, int(os.environ.get("WORLD_SIZE", "1"))),
        "cpu_baseline": {"value": v, "unit": "steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        # ms_per_step is scaled from the bounded sample to the full nb_steps (every Euler step is identical work);
        # value (steps/s) is a measured rate, not extrapolated
        "extrapolated": True, "measured_ms_per_step": dt / args.steps * 1e3, "measured_diffusion_steps_per_step": sample_steps,
    }))


def crc32(t):
    return zlib.crc32(t.detach().float().cpu().contiguous().numpy().tobytes()) & 0xffffffff


def codec_traffic():
    """DRAM bytes per launch of the codec kernel classes from the committed ncu --set full capture (profiles/), or None."""
    tr = os.path.join(ROOT, "profiles", "traffic_codec.json")
    if not os.path.exists(tr):
        return None
    with open(tr) as fh:
        return json.load(fh)


def gemm_roofline(eng, run, precision, pk, shape=None):
    """Per-class CUDA-event timing of one ``run()`` (graph bypassed) -> (roofline block of the dominant GEMM kernel, classes)."""
    eng.profile(True)
    run()
    prof = eng.profile_read()
    eng.profile(False)
    tc_mode = precision != "fp32_simt"
    dom = "mlp_fused" if (tc_mode and prof["mlp_fused"]["launches"]) else ("tap_gemm_tc" if tc_mode else "tap_gemm_simt")
    g = prof[dom]
    tot_prof = sum(v["ms"] for v in prof.values())
    if not g["launches"]:
        return None, prof
    ach = g["flops"] / (g["ms"] / 1e3) / 1e12
    issued = 3 if precision == "fp32" else 1
    kname = {"mlp_fused": "mlp_fused_tc2_kernel<256> (MLP up + down projection, one persistent CTA-pair tcgen05 launch)",
             "tap_gemm_tc": "tap_gemm_tc2_kernel (CTA-pair tcgen05 tap-GEMM)", "tap_gemm_simt": "tap_gemm_simt_kernel"}[dom]
    roof = {"bound": "tensor", "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
            "frac": ach / pk["bf16_sustained"], "traffic": None, "kernel": kname,
            "launches_per_sample": g["launches"], "avg_launch_us": g["ms"] * 1e3 / g["launches"],
            "flops_per_launch": g["flops"] / g["launches"], "share_of_profiled_ms": g["ms"] / tot_prof,
            "issued_mma_per_product": issued, "issued_frac": ach * issued / pk["bf16_sustained"],
            "peak_source": pk["source"] + ", bf16 sustained",
            "how": "algorithmic 2*M*(N0*K0 + N2*K2) flops per launch / mean CUDA-event duration of the launches of one "
                   "sample() call on the library's work stream (graph bypassed, PDL off between events)"}
    if dom == "mlp_fused" and shape is not None:
        # What actually bounds this kernel (profiles/EXPERIMENTS.md, in-kernel trace r02p_mlp_trace_*.txt): operand tiles
        # streamed from L2 into shared memory.  With 256 x 256 CTA-pair tiles the activations are re-read once per 256-wide
        # N tile and the weights once per 256-row block; bf16 hi + lo operands are 4 bytes per element (bf16 mode: 2).
        M, D, HID = shape
        rb, bpe = -(-M // 256), (4 if precision == "fp32" else 2)
        elems = (HID // 256) * M * D + rb * HID * D + (D // 256) * M * HID + rb * D * HID
        roof["l2_operand"] = {"bytes_per_launch": elems * bpe, "achieved_gbs": elems * bpe / (g["ms"] * 1e-3 / g["launches"]) / 1e9,
                              "cap_gbs_nominal": 12000.0,
                              "note": "L2 -> shared-memory operand traffic of the tile shape / mean launch duration; "
                                      "B300_MICROARCH.md puts the chip-wide L2 cap at ~6300 B/cycle (~12 TB/s); the in-kernel "
                                      "trace shows the mainloop waiting for operands, not for the tensor pipe or the epilogue"}
    if dom == "mlp_fused":
        q = prof["tap_gemm_tc"]
        if q["launches"]:
            roof["other_gemm_class"] = {"kernel": "tap_gemm_tc2_kernel (QKV + out-proj)", "launches": q["launches"],
                                        "achieved": q["flops"] / (q["ms"] / 1e3) / 1e12,
                                        "frac": q["flops"] / (q["ms"] / 1e3) / 1e12 / pk["bf16_sustained"]}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        with open(tr) as fh:
            tj = json.load(fh)
        roof["traffic"] = tj.get("tc::mlp_fused_tc2_kernel<256>_bytes_per_launch" if dom == "mlp_fused" else "tap_gemm_tc_kernel_bytes_per_launch")
        roof["traffic_source"] = tj.get("source")
    return roof, prof


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16", "fp32_simt"])
    ap.add_argument("--model", default="base", choices=["base", "tiny", "midi"])
    ap.add_argument("--batch", type=int, default=8, help="streams per GPU")
    ap.add_argument("--nb-steps", type=int, default=50, help="diffusion (Euler) steps per sample call")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-chain", action="store_true", help="skip the full-chain RTF leg")
    ap.add_argument("--no-stream", action="store_true", help="skip the streaming-block latency leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the legs for BASELINE configs 3/4/5 (bf16, midi, codec sweep)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from after_b200 import build, config, synth
    from after_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()

    B, NS = args.batch, args.nb_steps
    mc, x0_all, cond_all, tc_all = synth_setup(args.model, B * world)  # host-generated once: identical for any N
    from after_b200 import parallel
    x0_h, cond_h, tc_h = (t.pin_memory() for t in parallel.shard([x0_all, cond_all, tc_all], world, rank))
    acfg = config.base_autoencoder()
    den_sd = synth.denoiser_state_dict(mc.denoiser, 0)
    chain = not args.no_chain
    ae_sd = synth.autoencoder_state_dict(acfg, 0) if chain else None
    se_sd = synth.encoder1d_state_dict(mc.structure_encoder, 0) if (chain and mc.structure_encoder is not None) else None
    te_sd = synth.ecapa_state_dict(mc.timbre_encoder, 0) if chain else None
    eng = Engine(model=mc, autoencoder=acfg if chain else None, denoiser_state=den_sd, autoencoder_state=ae_sd,
                 structure_state=se_sd, timbre_state=te_sd, precision=args.precision, device=local, max_batch=B, max_steps=NS,
                 max_samples=CHUNK)
    x0, cond, tc = x0_h.to(dev), cond_h.to(dev), tc_h.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def one_step():
        out = eng.sample(x0, cond, tc, NS, 2.0, 1.0)
        return parallel.gather_streams(out, B * world)  # the single collective of the path (no-op at N = 1)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()

    def timed(fn, K):
        """K steps, L2 flushed before each, per-step CUDA events on the current stream; returns total ms (max over ranks)."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for a, b in evs:
            flush.fill_(1)
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        per = [a.elapsed_time(b) for a, b in evs]
        tot = torch.tensor([sum(per)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item()), per

    l0 = eng.launch_count
    clocks.mark()
    total_ms, per = timed(one_step, args.steps)
    clocks.mark()
    launches = eng.launch_count - l0
    clk = clocks.stop() if rank == 0 else None
    value = NS * args.steps * world / (total_ms / 1e3)

    # ---- end to end through the host-buffer C-ABI entry (H2D of x0/cond/time_cond, D2H of the latents inside) ----
    out_h = torch.empty_like(x0_h).pin_memory()

    def host_step():
        eng.sample_host(x0_h, cond_h, tc_h, out_h, NS, 2.0, 1.0)

    host_step()
    e2e_ms, _ = timed(host_step, args.steps)
    e2e = {"value": NS * args.steps * world / (e2e_ms / 1e3), "unit": "steps/s",
           "h2d_bytes_per_step": int(4 * (x0_h.numel() + cond_h.numel() + tc_h.numel())),
           "d2h_bytes_per_step": int(4 * out_h.numel())}

    # ---- full chain: real-time factor -----------------------------------------------------------------------------
    rtf = None
    if chain:
        audio_s = synth.synth_audio(B, CHUNK, seed=7 + rank)    # structure source
        audio_t = synth.synth_audio(B, CHUNK, seed=107 + rank)  # timbre source
        if se_sd is not None:
            a_s_h, a_t_h = audio_s.pin_memory(), audio_t.pin_memory()
            audio_out_h = torch.empty_like(a_s_h).pin_memory()
            a_s_d, a_t_d = a_s_h.to(dev), a_t_h.to(dev)

            def chain_step():  # device-resident inputs
                return eng.generate(a_s_d, a_t_d, x0, NS, 2.0, 1.0)

            def chain_host_step():  # host buffers: H2D of 2 x audio + noise, D2H of the audio, inside the call
                eng.generate_host(a_s_h, a_t_h, x0_h, audio_out_h, NS, 2.0, 1.0)

            chain_step()
            chain_ms, _ = timed(chain_step, args.steps)
            chain_host_step()
            chain_host_ms, _ = timed(chain_host_step, args.steps)
            audio_s_total = B * world * (CHUNK / SR) * args.steps
            rtf = {"value": audio_s_total / (chain_ms / 1e3), "unit": "x real time", "ms_per_chunk_batch": chain_ms / args.steps,
                   "e2e_value": audio_s_total / (chain_host_ms / 1e3), "e2e_ms_per_chunk_batch": chain_host_ms / args.steps,
                   "e2e_h2d_bytes_per_step": int(4 * (2 * a_s_h.numel() + x0_h.numel())), "e2e_d2h_bytes_per_step": int(4 * audio_out_h.numel()),
                   "chain": "after_generate: 2x AutoEncoder.encode + Encoder1D + ECAPATDNN + sample(50 steps, CFG) + AutoEncoder.decode"}

    # ---- streaming (live nn~ use): latency of one 4-frame block (8192 samples = 185.8 ms of audio) through the
    # per-diffusion-step KV caches, B = 1 stream (3 CFG rows), as the exported Streamer.sample runs it -----------------
    stream = None
    if rank == 0 and not args.no_stream:
        s_steps, s_frames = 8, 4
        s_eng = Engine(model=mc, denoiser_state=den_sd, precision=args.precision, device=local, max_batch=1, max_steps=s_steps,
                       seq_len=s_frames, max_cache_size=mc.denoiser.local_attention_size)
        sx, sc, st_ = x0[:1, :, :s_frames].contiguous(), cond[:1].contiguous(), tc[:1, :, :s_frames].contiguous()
        for _ in range(5):
            s_eng.sample_stream(sx, sc, st_, s_steps, 2.0, 1.0)
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
        for a, b in evs:
            a.record()
            s_eng.sample_stream(sx, sc, st_, s_steps, 2.0, 1.0)
            b.record()
        torch.cuda.synchronize()
        per_blk = sorted(a.elapsed_time(b) for a, b in evs)
        blk_audio_ms = s_frames * 2048 / SR * 1e3
        stream = {"block_frames": s_frames, "block_audio_ms": blk_audio_ms, "nb_steps": s_steps, "cache_frames": mc.denoiser.local_attention_size,
                  "median_ms_per_block": per_blk[len(per_blk) // 2], "p95_ms_per_block": per_blk[int(len(per_blk) * 0.95)],
                  "diffusion_steps_per_s": s_steps / (per_blk[len(per_blk) // 2] / 1e3),
                  "realtime_margin": blk_audio_ms / per_blk[len(per_blk) // 2],
                  "what": "after_sample_stream: one persistent kernel per block (all Euler steps, layers, CFG, roll_cache; "
                          "sampler only)"}
        s_eng.close()
        # the whole exported Streamer.forward on one 8192-sample buffer (export.py:398-455): two streaming codec encodes,
        # Encoder1D.forward_stream, ECAPA on the rolling timbre buffer, the streamed sampler, overlap-add decode
        if chain and se_sd is not None:
            from after_b200.streamer import Streamer
            f_eng = Engine(model=mc, autoencoder=acfg, denoiser_state=den_sd, autoencoder_state=ae_sd, structure_state=se_sd,
                           timbre_state=te_sd, precision=args.precision, device=local, max_batch=1, max_steps=s_steps, seq_len=64,
                           max_samples=64 * 2048, max_cache_size=mc.denoiser.local_attention_size, stream_slots=2,
                           stream_max_frames=s_frames)
            stm = Streamer(f_eng, n_signal_timbre=64, chunk_size=4)
            stm.set_nb_steps(s_steps); stm.set_guidance_timbre(2.0); stm.set_guidance_structure(1.0)
            buf = torch.cat([synth.synth_audio(1, s_frames * 2048, seed=3), synth.synth_audio(1, s_frames * 2048, seed=4)], 1).to(dev)
            noise = torch.randn(1, 64, s_frames, device=dev)
            for _ in range(5):
                stm.forward(buf, noise=noise)
            torch.cuda.synchronize()
            wall = []
            for _ in range(30):
                t0 = time.perf_counter()
                stm.forward(buf, noise=noise)
                torch.cuda.synchronize()
                wall.append((time.perf_counter() - t0) * 1e3)
            wall.sort()
            stream["streamer_forward"] = {
                "median_ms_per_buffer": wall[len(wall) // 2], "p95_ms_per_buffer": wall[int(len(wall) * 0.95)],
                "realtime_margin": blk_audio_ms / wall[len(wall) // 2],
                "what": "Streamer.forward on one 8192-sample buffer, host wall clock incl. launches and the final sync: 2 x "
                        "after_ae_encode_stream + after_structure_encode_stream + after_timbre_encode + after_sample_stream + "
                        "after_ae_decode_stream (device-resident buffers)"}
            f_eng.close()

    # ---- roofline of the dominant kernel (tcgen05 tap-GEMM), per-launch CUDA events on the launching stream ----------
    pk = peaks()
    roof, prof = (None, None)
    if rank == 0:
        roof, prof = gemm_roofline(eng, lambda: eng.sample(x0, cond, tc, NS, 2.0, 1.0), args.precision, pk,
                                   shape=(3 * B * x0.shape[-1], mc.denoiser.embed_dim, mc.denoiser.embed_dim * mc.denoiser.mlp_multiplier))

    # ---- output checksums: per-stream CRC32 of the sampled latents; streams 0..7 must not depend on --gpus ------------
    out_local = eng.sample(x0, cond, tc, NS, 2.0, 1.0)
    gathered = parallel.gather_streams(out_local, B * world)
    torch.cuda.synchronize()
    local_crcs = [crc32(out_local[i]) for i in range(out_local.shape[0])]
    if world > 1:
        all_crcs = [None] * world
        dist.all_gather_object(all_crcs, local_crcs)
        all_crcs = [c for part in all_crcs for c in part]
    else:
        all_crcs = local_crcs
    checksum = None
    if rank == 0:
        checksum = {"algo": "crc32 of the fp32 latents of each stream after the 50-step sample (host-seeded per stream)",
                    "streams_0_7": [f"{c:08x}" for c in all_crcs[:8]],
                    "all_streams": f"{zlib.crc32(' '.join(f'{c:08x}' for c in all_crcs).encode()) & 0xffffffff:08x}",
                    "gathered_equals_shards": all(crc32(gathered[i]) == all_crcs[i] for i in range(B * world)),
                    "n_streams": B * world}

    # ---- the other BASELINE.json configurations (per-GPU sizes), each with its own roofline block --------------------
    configs = None
    if not args.no_configs:
        configs = {}

        def sampler_leg(model_name, precision, variant, g_t, g_s, clamp, what):
            mcl, xa, ca, ta = synth_setup(model_name, B * world)
            xs, cs, ts = (t.to(dev) for t in parallel.shard([xa, ca, ta], world, rank))
            e = Engine(model=mcl, denoiser_state=synth.denoiser_state_dict(mcl.denoiser, 0), precision=precision, device=local,
                       max_batch=B, max_steps=NS)

            def run():
                return e.sample(xs, cs, ts, NS, g_t, g_s, cfg_variant=variant, clamp=clamp)

            def step():
                return parallel.gather_streams(run(), B * world)

            for _ in range(3):
                step()
            ms, _ = timed(step, args.steps)
            v = NS * args.steps * world / (ms / 1e3)
            leg = {"workload": what, "metric": "diffusion-steps/sec", "value": v, "unit": "steps/s", "ms_per_step": ms / args.steps,
                   "precision": precision, "batch_per_gpu": B, "global_batch": B * world, "nb_steps": NS,
                   "algorithmic_tflops": 3 * B * FLOP_PER_SEQ[model_name] * NS * args.steps * world / (ms / 1e3) / 1e12}
            if rank == 0:
                leg["roofline"], kp = gemm_roofline(e, run, precision, pk)
                leg["kernel_profile_ms"] = {k: round(v_["ms"], 3) for k, v_ in kp.items()}
            e.close()
            return leg

        if args.precision != "bf16" or args.model != "base":
            configs["config3_bf16"] = sampler_leg("base", "bf16", 0, 2.0, 1.0, 0.01,
                                                  f"BASELINE configs[2]: base audio-to-audio, bf16, batch={B}/GPU ({B * world} total), {NS} steps")
        if args.model != "midi":
            configs["config4_midi"] = sampler_leg("midi", args.precision, 1, 2.0, 3.0, 0.1,
                                                  f"BASELINE configs[3]: midi (128-pitch piano roll, window 16, MIDI CFG layout), "
                                                  f"batch={B}/GPU ({B * world} total), {NS} steps")

        # configs[4]: AutoEncoder encode + decode only, HBM GB/s sweep over the per-GPU batch
        def codec_leg(precision):
            a_sd = ae_sd if ae_sd is not None else synth.autoencoder_state_dict(acfg, 0)
            e = Engine(autoencoder=acfg, autoencoder_state=a_sd, precision=precision, device=local, max_batch=16, max_samples=CHUNK)
            rows = []
            bm = "bf16" if precision == "bf16" else "fp32"
            for b in (1, 2, 4, 8, 16):
                audio = synth.synth_audio(b, CHUNK, seed=7 + rank).to(dev)
                z = e.ae_encode(audio)
                for _ in range(3):
                    e.ae_encode(audio)
                    e.ae_decode(z)
                ms_e, _ = timed(lambda: e.ae_encode(audio), args.steps)
                ms_d, _ = timed(lambda: e.ae_decode(z), args.steps)
                te, td = ms_e / args.steps / 1e3, ms_d / args.steps / 1e3
                rows.append({"batch_per_gpu": b, "global_batch": b * world, "encode_ms": te * 1e3, "decode_ms": td * 1e3,
                             "encode_GBps": AE_BYTES["encode"][bm] * b / te / 1e9, "decode_GBps": AE_BYTES["decode"][bm] * b / td / 1e9,
                             "encode_frac_hbm": AE_BYTES["encode"][bm] * b / te / 1e9 / pk["hbm_gbs"],
                             "decode_frac_hbm": AE_BYTES["decode"][bm] * b / td / 1e9 / pk["hbm_gbs"],
                             "encode_TFLOPs": AE_FLOPS["encode"] * b / te / 1e12, "decode_TFLOPs": AE_FLOPS["decode"] * b / td / 1e12,
                             "chunks_per_s_all_gpus": b * world / (te + td), "rtf_encode_decode": b * world * (CHUNK / SR) / (te + td)})
            leg = {"workload": "BASELINE configs[4]: AutoEncoder encode + decode only, 524288-sample chunks, batch/GPU sweep",
                   "precision": precision, "byte_model": f"SURVEY.md 8d, {bm}: encode {AE_BYTES['encode'][bm] / 1e6:.0f} MB, decode "
                   f"{AE_BYTES['decode'][bm] / 1e6:.0f} MB per chunk (each conv reads its input once, writes its output once)",
                   "sweep": rows}
            if rank == 0:  # per-kernel-class device time of one encode + decode at the largest batch (graph bypassed)
                audio = synth.synth_audio(16, CHUNK, seed=7).to(dev)
                z = e.ae_encode(audio)
                e.profile(True)
                e.ae_encode(audio)
                e.ae_decode(z)
                kp = e.profile_read()
                e.profile(False)
                tot = sum(v_["ms"] for v_ in kp.values()) or 1.0
                g = kp["tap_gemm_tc"]
                last = rows[-1]
                leg["roofline"] = {
                    "bound": "hbm", "unit": "GB/s", "peak": pk["hbm_gbs"], "peak_source": pk["source"] + ", copy bandwidth",
                    "achieved": (AE_BYTES["encode"][bm] + AE_BYTES["decode"][bm]) * 16 / ((last["encode_ms"] + last["decode_ms"]) / 1e3) / 1e9,
                    "kernel": "whole encode + decode at 16 chunks/GPU against the byte model (conv tap-GEMMs dominate)",
                    "conv_class": {"launches": g["launches"], "ms": g["ms"], "share_of_profiled_ms": g["ms"] / tot,
                                   "GBps": g["bytes"] / (g["ms"] / 1e3) / 1e9 if g["ms"] else None,
                                   "TFLOPs": g["flops"] / (g["ms"] / 1e3) / 1e12 if g["ms"] else None},
                    "act_operand_class": {"launches": kp["act_operand"]["launches"], "ms": kp["act_operand"]["ms"],
                                          "share_of_profiled_ms": kp["act_operand"]["ms"] / tot,
                                          "GBps": kp["act_operand"]["bytes"] / (kp["act_operand"]["ms"] / 1e3) / 1e9 if kp["act_operand"]["ms"] else None},
                    "traffic": codec_traffic()}
                leg["roofline"]["frac"] = leg["roofline"]["achieved"] / pk["hbm_gbs"]
                leg["kernel_profile_ms"] = {k: round(v_["ms"], 3) for k, v_ in kp.items()}
            e.close()
            return leg

        configs["config5_codec"] = codec_leg(args.precision)
        if args.precision != "bf16":
            configs["config5_codec_bf16"] = codec_leg("bf16")

    # ---- CPU baseline: the oracle port of the reference sampler on this box's host cores (rank 0, N = 1 only) --------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v0, cores, _ = cpu_reference_steps_per_s(args.model, B, 2)          # probe, then ~12 s of CPU work
        cpu_steps = int(max(3, min(NS, round(12.0 * v0))))
        v, cores, secs = cpu_reference_steps_per_s(args.model, B, cpu_steps)
        cpu = {"value": v, "unit": "steps/s", "cores": cores, "kind": "port",
               "sample": f"{cpu_steps} of {NS} diffusion steps, B={B}, T=256 ({secs:.1f} s of CPU work)"}

    if rank == 0:
        dtype = {"fp32": "fp32 (bf16x3 split products on tcgen05, fp32 accumulate)", "bf16": "bf16 (fp32 accumulate)",
                 "fp32_simt": "fp32 (FFMA)"}[args.precision]
        line = {
            "metric": "diffusion-steps/sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": workload_config(args, world),
            "sequence_steps_per_s": value * B,
            "algorithmic_tflops": 3 * B * FLOP_PER_SEQ[args.model] * NS * args.steps * world / (total_ms / 1e3) / 1e12,
            "e2e": e2e, "rtf": rtf, "stream": stream, "roofline": roof, "cpu_baseline": cpu, "gpu_launches": int(launches), "clocks": clk,
            "checksum": checksum, "configs": configs,
            "kernel_profile_ms": {k: round(v["ms"], 3) for k, v in (prof or {}).items()},
        }
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
