"""GPU parity: secp256k1 through the C ABI vs the big-integer oracle and the reference's golden vectors."""
import random

import numpy as np
import pytest

import pyref
from helpers import edge_scalars, golden, pack_points, pack_scalars, random_points, unpack_points

pytestmark = pytest.mark.gpu
C = pyref.K256


def test_field_ops_vs_bigint(engine):
    rng = random.Random(11)
    p = C.p
    edge = [0, 1, 2, p - 1, p - 2, (p + 1) // 2, 2**32 + 977, 2**32 + 976, 2**255, 2**128]
    a = edge + [rng.randrange(p) for _ in range(4000)]
    b = [rng.choice(edge) for _ in edge] + [rng.randrange(p) for _ in range(4000)]
    A, B = pack_scalars(a), pack_scalars(b)
    model = {
        "add": lambda x, y: (x + y) % p,
        "sub": lambda x, y: (x - y) % p,
        "mul": lambda x, y: x * y % p,
        "neg": lambda x, y: (-x) % p,
        "sqr": lambda x, y: x * x % p,
        "inv": lambda x, y: pow(x, -1, p) if x else 0,
    }
    for op, f in model.items():
        out = engine.field_op("k256", op, A, B if op in ("add", "sub", "mul") else None)
        got = [int.from_bytes(out[i].tobytes(), "big") for i in range(len(a))]
        exp = [f(x, y) for x, y in zip(a, b)]
        assert got == exp, op


def test_field_golden_doubling_chain(engine):
    # DBL_TEST_VECTORS (k256/src/test_vectors/field.rs:6): 2^i; check v[i] + v[i] == v[i+1] and v[i]*2 via mul
    dbl = [int(h, 16) for h in golden("k256")["field"]["dbl"]]
    A = pack_scalars(dbl[:-1])
    out = engine.field_op("k256", "add", A, A)
    got = [int.from_bytes(out[i].tobytes(), "big") for i in range(len(dbl) - 1)]
    assert got == dbl[1:]
    two = pack_scalars([2] * (len(dbl) - 1))
    out = engine.field_op("k256", "mul", A, two)
    got = [int.from_bytes(out[i].tobytes(), "big") for i in range(len(dbl) - 1)]
    assert got == dbl[1:]


def test_field_rejects_noncanonical(engine):
    import ecgpu

    with pytest.raises(ecgpu.NotOnCurveError):
        engine.field_op("k256", "add", pack_scalars([1, C.p]), pack_scalars([1, 1]))


def test_golden_mul_vectors(engine):
    g = golden("k256")["group"]
    ks = [v["k"] for v in g["add"]] + [int(v["k"], 16) for v in g["mul"]]
    exp = [(int(v["x"], 16), int(v["y"], 16)) for v in g["add"] + g["mul"]]
    G = pyref.G(C)
    xy, inf = pack_points([G] * len(ks))
    out_xy, out_inf = engine.mul_batch("k256", pack_scalars(ks), xy, inf)
    assert unpack_points(out_xy, out_inf) == exp


def test_ecdsa_keypair_and_bench_scalars(engine):
    g = golden("k256")
    G = pyref.G(C)
    ks = [int(v["d"], 16) for v in g["ecdsa"]["keypairs"]]
    exp = [(int(v["x"], 16), int(v["y"], 16)) for v in g["ecdsa"]["keypairs"]]
    bs = [int(v["k"], 16) for v in g["bench"]["scalars"]]
    ks += bs
    exp += [pyref.mul(C, k, G) for k in bs]
    xy, inf = pack_points([G] * len(ks))
    out_xy, out_inf = engine.mul_batch("k256", pack_scalars(ks), xy, inf)
    assert unpack_points(out_xy, out_inf) == exp


def test_random_pairs_vs_oracle(engine):
    rng = random.Random(2024)
    n = 300
    Ps = random_points(C, n, seed=5)
    ks = [rng.randrange(C.n) for _ in range(n)]
    xy, inf = pack_points(Ps)
    out_xy, out_inf = engine.mul_batch("k256", pack_scalars(ks), xy, None)
    got = unpack_points(out_xy, out_inf)
    exp = [pyref.mul(C, k, P) for k, P in zip(ks, Ps)]
    assert got == exp


def test_edge_scalars_and_identity_inputs(engine):
    ks = edge_scalars(C)
    Ps = random_points(C, len(ks), seed=9)
    # identity inputs mixed in (k256/src/arithmetic/mul.rs:356-365: k*O == O, 0*P == O)
    Ps[3] = None
    Ps[10] = None
    xy, inf = pack_points(Ps)
    out_xy, out_inf = engine.mul_batch("k256", pack_scalars(ks), xy, inf)
    got = unpack_points(out_xy, out_inf)
    exp = [pyref.mul(C, k, P) for k, P in zip(ks, Ps)]
    assert got == exp
    assert got[0] is None and int(out_inf[0]) == 1 and not out_xy[0].any()


def test_ragged_sizes(engine):
    rng = random.Random(77)
    G = pyref.G(C)
    for n in (1, 2, 31, 33, 127, 129, 385):
        ks = [rng.randrange(C.n) for _ in range(n)]
        xy, inf = pack_points([G] * n)
        out_xy, out_inf = engine.mul_batch("k256", pack_scalars(ks), xy, inf)
        got = unpack_points(out_xy, out_inf)
        # spot check first/last against oracle, all against k*G linearity: (k)G + (n-k)G = O
        assert got[0] == pyref.mul(C, ks[0], G) and got[-1] == pyref.mul(C, ks[-1], G)
    out_xy, out_inf = engine.mul_batch("k256", np.zeros(0, np.uint8), np.zeros(0, np.uint8), None)
    assert out_xy.shape[0] == 0


def test_rejects_bad_inputs(engine):
    import ecgpu

    G = pyref.G(C)
    xy, inf = pack_points([G, G, G])
    with pytest.raises(ecgpu.ScalarRangeError) as ei:
        engine.mul_batch("k256", pack_scalars([1, C.n, 5]), xy, inf)
    assert ei.value.index == 1
    bad = xy.copy()
    bad[2 * 64 + 63] ^= 1  # off-curve
    with pytest.raises(ecgpu.NotOnCurveError) as ei:
        engine.mul_batch("k256", pack_scalars([1, 2, 3]), bad, inf)
    assert ei.value.index == 2
    bad = xy.copy()
    bad[0:32] = 0xFF  # x >= p
    with pytest.raises(ecgpu.NotOnCurveError):
        engine.mul_batch("k256", pack_scalars([1, 2, 3]), bad, inf)


def test_batch_normalize(engine):
    rng = random.Random(3)
    Ps = random_points(C, 40, seed=21)
    p = C.p
    xyz = bytearray()
    exp = []
    for i, P in enumerate(Ps):
        if i % 7 == 3:
            xyz += (rng.randrange(p)).to_bytes(32, "big") + (rng.randrange(p)).to_bytes(32, "big") + bytes(32)
            exp.append(None)
            continue
        z = rng.randrange(1, p)
        xyz += (P[0] * z * z % p).to_bytes(32, "big") + (P[1] * z * z * z % p).to_bytes(32, "big") + z.to_bytes(32, "big")
        exp.append(P)
    out_xy, out_inf = engine.batch_normalize("k256", np.frombuffer(bytes(xyz), dtype=np.uint8))
    assert unpack_points(out_xy, out_inf) == exp


def test_large_batch_properties(engine):
    """2^16 pairs: results must satisfy (k*P) + ((n-k)*P) == O  and match the oracle on a sample."""
    rng = random.Random(99)
    n = 1 << 16
    base = random_points(C, 64, seed=33)
    Ps = [base[i % 64] for i in range(n)]
    ks = [rng.randrange(1, C.n) for _ in range(n // 2)]
    ks = ks + [C.n - k for k in ks]
    Ps = Ps[: n // 2] + Ps[: n // 2]
    xy, inf = pack_points(Ps)
    out_xy, out_inf = engine.mul_batch("k256", pack_scalars(ks), xy, None)
    out_xy = np.asarray(out_xy).reshape(n, 64)
    h = n // 2
    assert not out_inf.any()
    # x equal, y negated
    assert (out_xy[:h, :32] == out_xy[h:, :32]).all()
    ya = [int.from_bytes(out_xy[i, 32:].tobytes(), "big") for i in range(0, h, 997)]
    yb = [int.from_bytes(out_xy[h + i, 32:].tobytes(), "big") for i in range(0, h, 997)]
    assert all((a + b) % C.p == 0 for a, b in zip(ya, yb))
    for i in range(0, h, 4099):
        assert unpack_points(out_xy[i : i + 1], out_inf[i : i + 1])[0] == pyref.mul(C, ks[i], Ps[i])
