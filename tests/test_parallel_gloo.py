"""world_size-2 (and ragged world_size-3) gloo runs of the batch-shard + single-gather driver on CPU.  The compute
on each rank is the CPU oracle (test infrastructure); what is checked is the host logic: shard bounds, host-seeded
inputs, gather order -> the gathered result is bit-identical to the single-process run."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from after_b200 import config, parallel, synth


def test_shard_bounds_cover_everything():
    for n in (1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_streams, q):
    from oracle import after_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mc = config.get_config("tiny")
        sd = synth.denoiser_state_dict(mc.denoiser, 5)
        x0, cond, tc = synth.synth_inputs(n_streams, mc.denoiser, seed=1234, frames=16)  # same on every rank
        lx, lc, lt = parallel.shard([x0, cond, tc], world, rank)
        if lx.shape[0]:
            local = O.sample(sd, mc.denoiser, lx, lc, lt, 2, 2.0, 1.0)
        else:
            local = lx
        full = parallel.gather_streams(local, n_streams)
        if rank == 0:
            q.put(full)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_streams", [(2, 4), (3, 4)])
def test_sharded_run_equals_single_process(world, n_streams):
    from oracle import after_oracle as O
    mc = config.get_config("tiny")
    sd = synth.denoiser_state_dict(mc.denoiser, 5)
    x0, cond, tc = synth.synth_inputs(n_streams, mc.denoiser, seed=1234, frames=16)
    # per-stream evaluation is the reference result for any sharding (rows of a batch are independent)
    want = torch.cat([O.sample(sd, mc.denoiser, x0[i:i + 1], cond[i:i + 1], tc[i:i + 1], 2, 2.0, 1.0) for i in range(n_streams)])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got.shape == want.shape
    assert float((got - want).abs().max()) < 1e-5
