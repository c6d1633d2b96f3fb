"""Multi-rank plumbing for the batched path: one process per GPU, `torch.distributed` used ONLY to scatter inputs /
gather outputs and for config 5's single exchange step (gather one 96-byte partial point per rank).  There is no
collective inside the arithmetic.

Backends
  nccl : everything stays on the device.  The source rank uploads the batch once, `dist.scatter` moves the slices over
         NVLink, every rank runs the kernels on the received device buffers through a device-pointer `ecgpu.Engine`
         (`mul_batch_ptr`, `lincomb_partial_ptr`, `point_sum_ptr`), `dist.gather` brings the affine results back as device
         tensors and the destination rank downloads them once.
  gloo : the CPU tests (world size 2, no GPU): same sharding and collectives on host tensors; `engine` is then anything
         with the numpy methods of ecgpu.Engine (the tests pass a stub backed by the oracle).
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


def collective_device():
    """Where tensors handed to the collectives must live: the current CUDA device under NCCL, the host under gloo."""
    import torch

    dist = _dist()
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def scatter_rows(rows, row_bytes: int, n: int, src: int = 0, device=None):
    """Rank `src` holds `rows` (uint8, n*row_bytes, numpy or tensor); every rank receives its shard_range slice as a
    tensor on `device` (default: collective_device())."""
    import torch

    dist = _dist()
    world, rank = dist.get_world_size(), dist.get_rank()
    device = collective_device() if device is None else torch.device(device)
    off, cnt = shard_range(n, world, rank)
    maxcnt = shard_range(n, world, 0)[1]  # equal-size scatter: slices are padded to the largest shard
    recv = torch.empty(maxcnt * row_bytes, dtype=torch.uint8, device=device)
    if rank == src:
        host = rows if isinstance(rows, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(rows, dtype=np.uint8).reshape(-1))
        full = torch.zeros(world * maxcnt * row_bytes, dtype=torch.uint8, device=device)
        if n % world == 0:
            full.copy_(host, non_blocking=True)  # one upload; the scatter list is views of it
        else:
            for r in range(world):
                o, c = shard_range(n, world, r)
                full[r * maxcnt * row_bytes: r * maxcnt * row_bytes + c * row_bytes].copy_(host[o * row_bytes:(o + c) * row_bytes], non_blocking=True)
        chunks = [full[r * maxcnt * row_bytes:(r + 1) * maxcnt * row_bytes] for r in range(world)]
        dist.scatter(recv, chunks, src=src)
    else:
        dist.scatter(recv, None, src=src)
    return recv[: cnt * row_bytes]


def gather_rows(local, row_bytes: int, n: int, dst: int = 0, out=None):
    """Inverse of scatter_rows: rank `dst` returns the concatenation as a numpy uint8 array (written into `out` if given,
    e.g. pinned memory), others None.  `local` is this rank's tensor (device under NCCL)."""
    import torch

    dist = _dist()
    world, rank = dist.get_world_size(), dist.get_rank()
    maxcnt = shard_range(n, world, 0)[1]
    if local.numel() == maxcnt * row_bytes:
        send = local
    else:
        send = torch.zeros(maxcnt * row_bytes, dtype=torch.uint8, device=local.device)
        send[: local.numel()] = local
    if rank == dst:
        full = torch.empty(world * maxcnt * row_bytes, dtype=torch.uint8, device=local.device)
        bufs = [full[r * maxcnt * row_bytes:(r + 1) * maxcnt * row_bytes] for r in range(world)]
        dist.gather(send, bufs, dst=dst)
        res = out if out is not None else np.empty(n * row_bytes, np.uint8)
        res_t = torch.from_numpy(res)
        if n % world == 0:
            res_t.copy_(full)
        else:
            for r in range(world):
                o, c = shard_range(n, world, r)
                res_t[o * row_bytes:(o + c) * row_bytes].copy_(bufs[r][: c * row_bytes])
        return res
    dist.gather(send, None, dst=dst)
    return None


def mul_batch_distributed(engine, curve, n: int, k=None, P_xy=None, src: int = 0, out_xy=None, out_inf=None):
    """configs 2-4 across ranks: scatter (k, P) from `src`, every rank multiplies its shard, gather the affine
    results back to `src`.  Returns (out_xy, out_inf) on `src`, (None, None) elsewhere.
    Under NCCL `engine` must be a device-pointer ecgpu.Engine on the current device; the slices never leave the GPU."""
    import torch

    dist = _dist()
    k_loc = scatter_rows(k, 32, n, src)
    p_loc = scatter_rows(P_xy, 64, n, src)
    cnt = k_loc.numel() // 32
    if dist.get_backend() == "nccl":
        oxy = torch.empty(64 * cnt, dtype=torch.uint8, device=k_loc.device)
        oinf = torch.empty(cnt, dtype=torch.uint8, device=k_loc.device)
        if cnt:
            engine.mul_batch_ptr(curve, cnt, k_loc.data_ptr(), p_loc.data_ptr(), 0, oxy.data_ptr(), oinf.data_ptr())
    else:
        hxy, hinf = engine.mul_batch(curve, k_loc.numpy(), p_loc.numpy(), None)
        oxy = torch.from_numpy(np.ascontiguousarray(hxy).reshape(-1))
        oinf = torch.from_numpy(np.ascontiguousarray(hinf).reshape(-1))
    g_xy = gather_rows(oxy, 64, n, src, out_xy)
    g_inf = gather_rows(oinf, 1, n, src, out_inf)
    return g_xy, g_inf


def lincomb_distributed(engine, curve, k_local, P_xy_local, P_inf_local=None, dst: int = 0):
    """config 5: every rank reduces ITS terms to one Jacobian point (96 B); one all_gather; rank `dst` adds the
    `world` partial points and normalises.  Returns (xy, inf) on `dst`, None elsewhere.
    NCCL: k_local / P_xy_local are device tensors (uint8), `engine` a device-pointer Engine; the partial points are summed
    where the all_gather left them.  gloo: numpy arrays and a host-pointer engine."""
    import torch

    dist = _dist()
    world, rank = dist.get_world_size(), dist.get_rank()
    if dist.get_backend() == "nccl":
        dev = collective_device()
        n = k_local.numel() // 32
        part = torch.empty(96, dtype=torch.uint8, device=dev)
        parts = torch.empty(96 * world, dtype=torch.uint8, device=dev)
        engine.lincomb_partial_ptr(curve, n, k_local.data_ptr(), P_xy_local.data_ptr(), P_inf_local.data_ptr() if P_inf_local is not None else 0,
                                   part.data_ptr())
        dist.all_gather_into_tensor(parts, part)
        if rank != dst:
            return None
        res = torch.empty(65, dtype=torch.uint8, device=dev)
        engine.point_sum_ptr(curve, world, parts.data_ptr(), res.data_ptr(), res.data_ptr() + 64)
        h = res.cpu().numpy()
        return h[:64].copy(), int(h[64])
    part = engine.lincomb_partial(curve, k_local, P_xy_local, P_inf_local)
    send = torch.from_numpy(np.ascontiguousarray(part, dtype=np.uint8).reshape(-1))
    bufs = [torch.empty(96, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(bufs, send)
    if rank != dst:
        return None
    allp = np.concatenate([b.numpy() for b in bufs])
    return engine.point_sum(curve, allp)
