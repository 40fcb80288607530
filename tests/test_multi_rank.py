"""N > 1 plumbing on CPU: world_size-2 gloo processes shard a job, 'process' their block and gather on rank 0."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gtcrn_micro_b200 import shard


def test_shard_plan_covers_everything_once():
    for n in (0, 1, 7, 8, 1024, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            plan = shard.shard_plan(n, world)
            assert plan[0][0] == 0 and plan[-1][1] == n
            assert all(a <= b for a, b in plan)
            assert all(plan[i][1] == plan[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in plan) <= (n + world - 1) // world
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def _worker(rank, world, port, n_items, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = shard.shard_range(n_items, rank, world)
        g = torch.Generator().manual_seed(0)
        wav = torch.randn(n_items, 64, generator=g)              # every rank derives the same job, touches only its block
        local = wav[a:b] * 2.0 + 1.0                              # stand-in for model.enhance(wav[a:b]) (no GPU here)
        full = shard.gather_to_rank0(local, n_items)
        slowest = shard.max_over_ranks(1.0 + rank)                # rank r "took" 1+r seconds
        assert slowest == float(world)
        if rank == 0:
            assert torch.equal(full, wav * 2.0 + 1.0)
            torch.save({"ok": True, "slowest": slowest, "shape": tuple(full.shape)}, out_path)
        else:
            assert full is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [10, 7, 1])
def test_two_rank_gloo_shard_and_gather(tmp_path, n_items):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, port, n_items, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["ok"] and res["slowest"] == 2.0 and res["shape"] == (n_items, 64)
