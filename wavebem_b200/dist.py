"""Row sharding across GPUs: one process per GPU, torch.distributed only as plumbing.

Rows of both matrices are partitioned in contiguous blocks (the same formula as
wbem_set_topology in csrc/api.cu); vectors are replicated.  The only data-path collective is
the all-gather of result rows inside libwbem (NCCL, wbem_comm_init); torch.distributed is
used to hand the NCCL unique id from rank 0 to the other ranks and for host-side gathers in
tests.
"""
from __future__ import annotations

import os

import numpy as np


def row_block(n: int, rank: int, world: int):
    """[row0, row1) of `rank`: blocks of ceil(n / world) rows, the tail block may be short/empty."""
    chunk = (n + world - 1) // world
    r0 = min(n, rank * chunk)
    return r0, min(n, r0 + chunk)


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0) -> bytes:
    """Broadcast a fixed-size byte string from `src` over the default process group."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def init_comm(ctx):
    """Create libwbem's NCCL communicator for `ctx` (no-op for world_size 1)."""
    import torch.distributed as dist
    if ctx.params.world_size <= 1:
        return
    uid = type(ctx).comm_unique_id() if dist.get_rank() == 0 else None
    uid = broadcast_bytes(uid, 128, src=0)
    ctx.comm_init(uid)


def init_peer_gather(ctx) -> bool:
    """After ctx.set_topology on every rank: exchange CUDA-IPC handles of the gather buffers so
    that the mat-vec kernel writes its rows straight into every peer (fused GEMV + all-gather).
    Returns False (and leaves the NCCL all-gather path active) when IPC is unavailable."""
    import torch
    import torch.distributed as dist
    if ctx.params.world_size <= 1:
        return False
    world = dist.get_world_size()
    ok = 1
    try:
        mine = ctx.ipc_export()
    except Exception:
        mine, ok = bytes(64), 0
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    flag = torch.tensor([ok], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        return False
    handles = b"".join(bytes(o.cpu().numpy().tobytes()) for o in out)
    try:
        ctx.ipc_import(handles)
        ok = 1
    except Exception:
        ok = 0
    flag = torch.tensor([ok], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        ctx.ipc_close()   # some rank failed: nobody may use the peer path
        return False
    return True


def allgather_rows(local: np.ndarray, n: int) -> np.ndarray:
    """Host-side gather of per-rank row blocks (tests / diagnostics): (row1-row0, ...) -> (n, ...)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    chunk = (n + world - 1) // world
    pad = np.zeros((chunk,) + local.shape[1:], dtype=local.dtype)
    pad[: local.shape[0]] = local
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.from_numpy(pad).to(dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return np.concatenate([o.cpu().numpy() for o in out], axis=0)[:n]
