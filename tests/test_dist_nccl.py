"""GPU, world_size 2 over NCCL (skipped on boxes with one GPU; `bench.py --gpus N` runs the same paths at N >= 2):
ecgpu/dist.py with device tensors end to end — scatter over NVLink, kernels on the received slices through the
device-pointer C ABI, gather of device outputs, and config 5's exchange step (all_gather of the 96-byte partial points,
device-side sum) — bit-exact against the oracle."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    for p in (ROOT, os.path.join(ROOT, "elliptic-curves_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    import ecgpu
    import ecref
    import pyref
    from ecgpu import dist as ecdist
    from helpers import pack_points, pack_scalars, random_points

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    eng = ecgpu.Engine([rank], device_ptrs=True)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    ok = True
    for curve in ("k256", "p256"):
        c = pyref.CURVES[curve]
        ks = [pyref.synth_scalar(c, 77, b"k", i) for i in range(n)]
        ks[0], ks[1], ks[n - 1] = 0, 1, c.n - 1
        base = random_points(c, 16, seed=5)
        Ps = [base[i % 16] for i in range(n)]
        k_all = pack_scalars(ks)
        xy_all, _ = pack_points(Ps)
        # configs 2-4 shape: rank 0 holds the batch, n not divisible by the world size
        g_xy, g_inf = ecdist.mul_batch_distributed(eng, curve, n, k_all if rank == 0 else None, xy_all if rank == 0 else None, src=0)
        if rank == 0:
            r_xy, r_inf = ecref.mul_batch(curve, k_all, xy_all, None, nthreads=4)
            ok = ok and np.array_equal(g_xy, r_xy.reshape(-1)) and np.array_equal(g_inf, r_inf)
        else:
            ok = ok and g_xy is None
        # config 5 shape: every rank owns a slice of the terms (bucket method: > 2^13 terms per rank)
        reps = 4
        off, cnt = ecdist.shard_range(n, world, rank)
        k_loc = np.tile(k_all[32 * off:32 * (off + cnt)], reps)
        p_loc = np.tile(xy_all[64 * off:64 * (off + cnt)], reps)
        res = ecdist.lincomb_distributed(eng, curve, torch.from_numpy(k_loc).cuda(), torch.from_numpy(p_loc).cuda(), None, dst=0)
        if rank == 0:
            e_xy, e_inf = ecref.lincomb(curve, np.tile(k_all, reps), np.tile(xy_all, reps), None, nthreads=4)
            ok = ok and np.array_equal(res[0], e_xy) and res[1] == e_inf
        else:
            ok = ok and res is None
    q.put((rank, bool(ok), eng.kernel_launches))
    dist.barrier()
    dist.destroy_process_group()


def test_nccl_scatter_compute_gather_and_exchange():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (covered by `bench.py --gpus N`, N >= 2: strong_scaling / config 5 records)")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world, n = 2, 9001
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(launches > 0 for _, _, launches in res)
