"""CPU-only, world_size 2 over gloo: the multi-rank host logic (sharding, scatter/gather, config-5 exchange step).
The arithmetic is stubbed with the oracle (test infrastructure) — what is under test is ecgpu/dist.py."""
import os
import random
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleEngine:
    """Stub with the Engine methods dist.py uses, backed by oracle/ecref (CPU)."""

    def mul_batch(self, curve, k, P_xy, P_inf=None):
        import ecref

        return ecref.mul_batch(curve, k, P_xy, P_inf, nthreads=2)

    def lincomb_partial(self, curve, k, P_xy, P_inf=None):
        import ecref

        n = np.asarray(k).size // 32
        out = np.zeros(96, np.uint8)
        if n == 0:
            out[63] = 1
            return out
        xy, inf = ecref.lincomb(curve, k, P_xy, P_inf, nthreads=2)
        if inf:
            out[63] = 1
        else:
            out[:64] = xy
            out[95] = 1  # affine point as Jacobian with Z = 1
        return out

    def point_sum(self, curve, xyz):
        import ecref
        import pyref

        c = pyref.CURVES[curve]
        acc = None
        xyz = np.asarray(xyz, np.uint8).reshape(-1, 96)
        for row in xyz:
            z = int.from_bytes(row[64:].tobytes(), "big")
            if z == 0:
                continue
            assert z == 1
            acc = pyref.add(c, acc, (int.from_bytes(row[:32].tobytes(), "big"), int.from_bytes(row[32:64].tobytes(), "big")))
        b, f = pyref.enc_point(acc)
        return np.frombuffer(b, np.uint8).copy(), f


def _worker(rank, world, port, n, q):
    for p in (ROOT, os.path.join(ROOT, "elliptic-curves_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist

    import pyref
    from ecgpu import dist as ed
    from helpers import pack_points, pack_scalars, random_points

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        c = pyref.K256
        rng = random.Random(5)
        ks = [rng.randrange(c.n) for _ in range(n)]
        base = random_points(c, 8, seed=3)
        Ps = [base[i % 8] for i in range(n)]
        xy, inf = pack_points(Ps)
        K = pack_scalars(ks)
        eng = OracleEngine()
        # configs 2-4 shape: scatter -> shard compute -> gather
        g_xy, g_inf = ed.mul_batch_distributed(eng, "k256", n, K if rank == 0 else None, xy if rank == 0 else None)
        # config 5 shape: local partial -> all_gather(96 B) -> rank-0 sum
        off, cnt = ed.shard_range(n, world, rank)
        res = ed.lincomb_distributed(eng, "k256", K[32 * off:32 * (off + cnt)], xy[64 * off:64 * (off + cnt)])
        if rank == 0:
            exp = [pyref.mul(c, k, P) for k, P in zip(ks, Ps)]
            from helpers import unpack_points

            ok1 = unpack_points(g_xy, g_inf) == exp
            ok2 = pyref.dec_point(res[0].tobytes(), res[1]) == pyref.lincomb(c, ks, Ps)
            q.put((ok1, ok2))
        else:
            assert g_xy is None and res is None
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    sys.path.insert(0, os.path.join(ROOT, "elliptic-curves_b200"))
    from ecgpu.dist import shard_range

    for n in (0, 1, 7, 8, 9, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                off, cnt = shard_range(n, world, r)
                cover += list(range(off, off + cnt)) if n < 100 else []
                assert cnt in (n // world, n // world + 1)
            if n < 100:
                assert cover == list(range(n))
            assert sum(shard_range(n, world, r)[1] for r in range(world)) == n


@pytest.mark.timeout(300)
def test_world2_gloo_scatter_gather_and_lincomb_exchange():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n = 37  # ragged: 19 + 18
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    assert q.get(timeout=5) == (True, True)
