"""Run under torchrun (N ranks): batch-sharded sample + single gather must be bit-identical to the 1-GPU run."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from after_b200 import config, parallel, synth
from after_b200.engine import Engine

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B_total, steps = 4 * world + 1, 4  # ragged on purpose
mc = config.get_config("base")
sd = synth.denoiser_state_dict(mc.denoiser, 0)
x0, cond, tc = synth.synth_inputs(B_total, mc.denoiser, seed=1234, frames=64)
eng = Engine(model=mc, denoiser_state=sd, precision="fp32", device=local, max_batch=B_total, max_steps=steps, seq_len=64)
lx, lc, lt = (t.to(dev) for t in parallel.shard([x0, cond, tc], world, rank))
out = eng.sample(lx, lc, lt, steps, 2.0, 1.0)
full = parallel.gather_streams(out, B_total)
if rank == 0:
    ref = eng.sample(x0.to(dev), cond.to(dev), tc.to(dev), steps, 2.0, 1.0)
    same = torch.equal(full, ref)
    print(f"world={world} streams={B_total}: gathered == single-GPU result bitwise: {same}; max|diff|={float((full - ref).abs().max()):.3e}")
    assert same
eng.close()
if world > 1:
    dist.destroy_process_group()
