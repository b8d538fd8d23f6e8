"""world_size-2 (and 8) gloo runs on CPU: the host-side multi-image logic -- launcher environment, id broadcast,
message plans of the ghost exchange and of the distributed coarse FFT (the device work needs a GPU, see
tests/test_gpu_multi_image.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nn, nc, nnt, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cafproject_b200.cube import exchange_plan, image_grid
        from cafproject_b200.dist import broadcast_bytes, check_plans, env_rank
        assert env_rank() == (rank, world, rank)
        assert image_grid(world) == tuple(nn)
        # the 128-byte id travels from image 1 to everybody (a stand-in payload: NCCL itself needs a GPU)
        payload = bytes(range(128)) if rank == 0 else None
        assert broadcast_bytes(payload, 128) == bytes(range(128))
        mine = exchange_plan(nn, rank, nc, nnt)
        plans = [None] * world
        dist.all_gather_object(plans, mine)
        check_plans(plans)
        # ghost cells of one image = extended grid minus physical minus the self-aliased (periodic) part
        ncell = sum(g[5] for g in mine["ghost"])
        ext = [nc + 12 if n > 1 else nc for n in nn]
        assert ncell >= ext[0] * ext[1] * ext[2] - nc ** 3
        t = torch.tensor([ncell], dtype=torch.int64)
        dist.all_reduce(t)
        if rank == 0:
            q.put(int(t.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nn,nc,nnt", [((2, 1, 1), 24, 2), ((2, 2, 2), 24, 2)])
def test_plans_agree_across_ranks(nn, nc, nnt):
    world = nn[0] * nn[1] * nn[2]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nn, nc, nnt, q)) for r in range(world)]
    for p in procs: p.start()
    for p in procs: p.join(180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=5) > 0


def test_plan_single_image_has_no_messages():
    from cafproject_b200.cube import exchange_plan
    p = exchange_plan((1, 1, 1), 0, 32, 2)
    assert p["ghost"] == [] and p["force_send"] == []
