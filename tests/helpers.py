"""Shared helpers for the parity tests (oracle side = test infrastructure)."""
import json
import os
import random

import numpy as np

import pyref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(curve):
    with open(os.path.join(GOLDEN, f"{curve}.json")) as f:
        return json.load(f)


def pack_scalars(ks):
    return np.frombuffer(b"".join(int(k).to_bytes(32, "big") for k in ks), dtype=np.uint8).copy()


def pack_points(Ps):
    xy = bytearray()
    inf = bytearray()
    for P in Ps:
        b, f = pyref.enc_point(P)
        xy += b
        inf.append(f)
    return np.frombuffer(bytes(xy), dtype=np.uint8).copy(), np.frombuffer(bytes(inf), dtype=np.uint8).copy()


def unpack_points(out_xy, out_inf):
    out_xy = np.asarray(out_xy, dtype=np.uint8).reshape(-1, 64)
    res = []
    for i in range(out_xy.shape[0]):
        res.append(pyref.dec_point(out_xy[i].tobytes(), int(out_inf[i])))
    return res


def random_points(c, n, seed):
    """Uniform group elements t*G, t from a seeded RNG (k256/tests/projective.rs:21-25 builds points the same way)."""
    rng = random.Random(seed)
    G = pyref.G(c)
    return [pyref.mul(c, rng.randrange(1, c.n), G) for _ in range(n)]


def edge_scalars(c):
    n = c.n
    return [0, 1, 2, 3, 4, 7, 8, 15, 16, 17, 31, 32, 2**32 - 1, 2**32, 2**64, 2**127, 2**128 - 1, 2**128, 2**128 + 1,
            2**129, 2**255, n - 1, n - 2, n - 3, (n - 1) // 2, (n + 1) // 2, n // 3, pyref.K256_LAMBDA % n,
            (n - pyref.K256_LAMBDA) % n, int("5" * 64, 16) % n, int("a" * 64, 16) % n, int("f" * 63, 16)]
