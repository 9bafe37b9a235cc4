"""N > 1 host logic on CPU (gloo, world size 2): streams shard contiguously with no data-path
collective; every rank denoises only its own shard and the union equals the single-process result."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, frames, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rnnoise_b200
    from oracle.portbind import Port
    from rnnoise_b200.synth_pcm import stream_pcm
    first, count = rnnoise_b200.shard(total, world, rank)
    port_ = Port(os.path.join(ROOT, "tests", "golden", "models", "default.bin"))
    sums = torch.zeros(total, dtype=torch.float64)
    for s in range(first, first + count):          # the checker stands in for the GPU batch on CPU
        st = port_.create()
        pcm = stream_pcm(s, frames)
        acc = 0.0
        for f in range(frames):
            acc += float(np.abs(port_.process_frame(st, pcm[f], trace=False)["out"]).sum())
        sums[s] = acc
        port_.destroy(st)
    # the only cross-rank traffic is bookkeeping (like bench.py's max-over-ranks of the timing)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ranges = [None] * world
    dist.all_gather_object(ranges, (first, count))
    if rank == 0:
        q.put((sums.numpy(), ranges))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    total, frames, world = 5, 3, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    sums, ranges = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # shards are contiguous, disjoint and cover every stream
    assert ranges == [(0, 3), (3, 2)]
    sys.path.insert(0, ROOT)
    from oracle.portbind import Port
    from rnnoise_b200.synth_pcm import stream_pcm
    port_ = Port(os.path.join(ROOT, "tests", "golden", "models", "default.bin"))
    for s in range(total):
        st = port_.create()
        pcm = stream_pcm(s, frames)
        ref = sum(float(np.abs(port_.process_frame(st, pcm[f], trace=False)["out"]).sum()) for f in range(frames))
        assert abs(ref - sums[s]) < 1e-9, s


def test_shard_helper_covers_everything():
    import rnnoise_b200
    for total in (1, 7, 4096, 65536):
        for world in (1, 2, 3, 8):
            got = [rnnoise_b200.shard(total, world, r) for r in range(world)]
            assert got[0][0] == 0 and sum(c for _, c in got) == total
            assert all(got[i][0] + got[i][1] == got[i + 1][0] for i in range(world - 1))
