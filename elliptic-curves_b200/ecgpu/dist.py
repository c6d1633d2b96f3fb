"""Multi-rank plumbing for the batched path: one process per GPU, `torch.distributed` (NCCL on GPUs, gloo in the
CPU tests) used ONLY to scatter inputs / gather outputs, and for config 5's single exchange step (gather one
96-byte partial point per rank).  There is no collective inside the arithmetic.

The `engine` argument is anything with the ecgpu.Engine methods used below (the CPU gloo tests pass a stub backed
by the oracle; production passes ecgpu.Engine).
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, world: int, rank: int):
    """Contiguous index range of `rank`: same rule as make_shards() in csrc/ecgpu.cu."""
    base, rem = divmod(n, world)
    cnt = base + (1 if rank < rem else 0)
    off = rank * base + min(rank, rem)
    return off, cnt


def _dist():
    import torch.distributed as dist

    return dist


def scatter_rows(rows, row_bytes: int, n: int, src: int = 0, device="cpu"):
    """Rank `src` holds `rows` (uint8, n*row_bytes); every rank receives its shard_range slice."""
    import torch

    dist = _dist()
    world, rank = dist.get_world_size(), dist.get_rank()
    off, cnt = shard_range(n, world, rank)
    # equal-size scatter needs padding to the largest shard
    maxcnt = shard_range(n, world, 0)[1]
    recv = torch.empty(maxcnt * row_bytes, dtype=torch.uint8, device=device)
    if rank == src:
        full = torch.as_tensor(np.ascontiguousarray(rows, dtype=np.uint8).reshape(-1)).to(device)
        chunks = []
        for r in range(world):
            o, c = shard_range(n, world, r)
            t = torch.zeros(maxcnt * row_bytes, dtype=torch.uint8, device=device)
            t[: c * row_bytes] = full[o * row_bytes:(o + c) * row_bytes]
            chunks.append(t)
        dist.scatter(recv, chunks, src=src)
    else:
        dist.scatter(recv, None, src=src)
    return recv[: cnt * row_bytes]


def gather_rows(local, row_bytes: int, n: int, dst: int = 0):
    """Inverse of scatter_rows: rank `dst` returns the concatenation (numpy uint8), others None."""
    import torch

    dist = _dist()
    world, rank = dist.get_world_size(), dist.get_rank()
    maxcnt = shard_range(n, world, 0)[1]
    send = torch.zeros(maxcnt * row_bytes, dtype=torch.uint8, device=local.device)
    send[: local.numel()] = local
    if rank == dst:
        bufs = [torch.empty_like(send) for _ in range(world)]
        dist.gather(send, bufs, dst=dst)
        out = []
        for r in range(world):
            _, c = shard_range(n, world, r)
            out.append(bufs[r][: c * row_bytes].cpu().numpy())
        return np.concatenate(out) if out else np.zeros(0, np.uint8)
    dist.gather(send, None, dst=dst)
    return None


def mul_batch_distributed(engine, curve, n: int, k=None, P_xy=None, src: int = 0):
    """configs 2-4 across ranks: scatter (k, P) from `src`, every rank multiplies its shard, gather the affine
    results back to `src`.  Returns (out_xy, out_inf) on `src`, (None, None) elsewhere."""
    k_loc = scatter_rows(k, 32, n, src)
    p_loc = scatter_rows(P_xy, 64, n, src)
    oxy, oinf = engine.mul_batch(curve, k_loc.cpu().numpy(), p_loc.cpu().numpy(), None)
    import torch

    g_xy = gather_rows(torch.as_tensor(np.ascontiguousarray(oxy).reshape(-1)), 64, n, src)
    g_inf = gather_rows(torch.as_tensor(np.ascontiguousarray(oinf).reshape(-1)), 1, n, src)
    return g_xy, g_inf


def lincomb_distributed(engine, curve, k_local, P_xy_local, P_inf_local=None, dst: int = 0):
    """config 5: every rank reduces ITS terms to one Jacobian point (96 B); one all_gather; rank `dst` adds the
    `world` partial points and normalises.  Returns (xy, inf) on `dst`, None elsewhere."""
    import torch

    dist = _dist()
    world, rank = dist.get_world_size(), dist.get_rank()
    part = engine.lincomb_partial(curve, k_local, P_xy_local, P_inf_local)
    send = torch.as_tensor(np.ascontiguousarray(part, dtype=np.uint8).reshape(-1))
    bufs = [torch.empty(96, dtype=torch.uint8) for _ in range(world)]
    if dist.get_backend() == "nccl":
        send = send.cuda()
        bufs = [b.cuda() for b in bufs]
    dist.all_gather(bufs, send)
    if rank != dst:
        return None
    allp = np.concatenate([b.cpu().numpy() for b in bufs])
    return engine.point_sum(curve, allp)
