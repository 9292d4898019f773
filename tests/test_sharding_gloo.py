"""World-size-2 checks of the stream sharding and the final gather (gloo on CPU; NCCL on the GPU box)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neuralsteganography_b200.sharding import gather_ragged, shard_bounds, shard_list


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _worker(rank, world, port, n_streams, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        streams = [[s * 10 + k for k in range(s % 5)] for s in range(n_streams)]     # ragged "token lists"
        mine = shard_list(streams, rank, world)
        width = 6
        rows = torch.full((len(mine), width), -1, dtype=torch.int32)
        lens = torch.zeros(len(mine), dtype=torch.int32)
        for i, t in enumerate(mine):
            rows[i, : len(t)] = torch.tensor(t, dtype=torch.int32)
            lens[i] = len(t)
        got = gather_ragged(rows, lens, dst=0)
        if rank == 0:
            q.put(got == streams)
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_streams", [7, 8])
def test_gather_ragged_world2(n_streams):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_streams
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_streams, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
