"""One-process-per-GPU plumbing on top of ``torch.distributed`` (used by bench.py and the multi-rank tests).

The coarray runtime of the reference (``this_image()``, ``sync all``, cafcube.f90:6-14) is replaced by: rank/world from
the launcher's environment, image grid from :func:`cafproject_b200.cube.image_grid`, and the library's own NCCL
communicator whose 128-byte id is made on image 1 and broadcast here.  torch is plumbing only.
"""
from __future__ import annotations

import os

import numpy as np


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0, device=None) -> bytes:
    """Broadcast ``nbytes`` bytes from ``src`` to every rank of the default process group (gloo or nccl)."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(buf, src=src)
    return bytes(buf.cpu().numpy().tobytes())


def shared_nccl_id(device=None) -> bytes:
    """ncclUniqueId of the library's communicator: made on rank 0, broadcast to all."""
    import torch.distributed as dist
    from .cube import nccl_unique_id
    mine = nccl_unique_id() if dist.get_rank() == 0 else None
    return broadcast_bytes(mine, 128, 0, device)


def check_plans(plans) -> None:
    """Pairwise consistency of the per-image message plans (:func:`cafproject_b200.cube.exchange_plan`): for every
    ordered pair (S, R) the sequence of messages S sends to R equals the sequence R expects from S -- the matching rule
    of grouped ncclSend/ncclRecv."""
    n = len(plans)
    for s in range(n):
        for r in range(n):
            sent = [g[5] for g in plans[s]["ghost"] if g[4] == r]       # S -> R: (.., src, dst, ncell, cell0)
            recv = [g[5] for g in plans[r]["ghost"] if g[3] == s]
            assert sent == recv, ("ghost", s, r, sent, recv)
            fs = [k for q, k in plans[s]["force_send"] if q == r]
            fr = [k for q, k in plans[r]["force_recv"] if q == s]
            assert fs == fr, ("force planes", s, r, fs, fr)
    nc2 = None
    for p in plans:
        tot = sum(k for _, k in p["force_recv"])
        nc2 = tot if nc2 is None else nc2
        assert tot == nc2                                               # every image receives nc+2 planes
