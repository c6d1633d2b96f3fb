"""GPU parity for NIST P-384 — the next curve through the same templates — through the C ABI
with 48-byte records, against the big-integer model pinned to the reference's own P-384 vectors
(p384/src/test_vectors/group.rs:8,175 -> tests/golden/p384.json, oracle/extract_golden.py)."""
import random

import numpy as np
import pytest

import ecgpu
import pyref
from helpers import golden

pytestmark = pytest.mark.gpu
C = pyref.P384
NB = 48


def ks_bytes(ks):
    return np.frombuffer(b"".join(int(k).to_bytes(NB, "big") for k in ks), np.uint8).copy()


def pts_bytes(Ps):
    xy = np.frombuffer(b"".join(pyref.enc_point(P, NB)[0] for P in Ps), np.uint8).copy()
    return xy, np.array([1 if P is None else 0 for P in Ps], np.uint8)


def unpack(xy, inf):
    xy = np.asarray(xy, np.uint8).reshape(-1, 2 * NB)
    return [pyref.dec_point(xy[i].tobytes(), int(inf[i]), NB) for i in range(xy.shape[0])]


def rand_points(n, seed):
    rng = random.Random(seed)
    base = [pyref.mul(C, rng.randrange(1, C.n), pyref.G(C)) for _ in range(16)]
    return [base[i % 16] for i in range(n)]


def test_golden_vectors(engine):
    g = golden("p384")
    G = pyref.G(C)
    ks = list(range(1, 21)) + [int(v["k"], 16) for v in g["group"]["mul"]]
    want = [(int(v["x"], 16), int(v["y"], 16)) for v in g["group"]["add"]] + [(int(v["x"], 16), int(v["y"], 16)) for v in g["group"]["mul"]]
    xy, inf = engine.mul_by_generator("p384", ks_bytes(ks))          # fixed-base table (24 sixteen-bit windows)
    assert unpack(xy, inf) == want
    pxy, pinf = pts_bytes([G] * len(ks))
    xy, inf = engine.mul_batch("p384", ks_bytes(ks), pxy, pinf)     # variable-base kernel
    assert unpack(xy, inf) == want


def test_var_base_random_edges_and_identities(engine):
    rng = random.Random(3840)
    n = 700
    ks = [rng.randrange(C.n) for _ in range(n)]
    ks[:8] = [0, 1, 2, C.n - 1, C.n - 2, 2**383, 2**192, (C.n - 1) // 2]
    Ps = rand_points(n, 1)
    Ps[5] = None
    Ps[6] = None
    pxy, pinf = pts_bytes(Ps)
    xy, inf = engine.mul_batch("p384", ks_bytes(ks), pxy, pinf)
    assert unpack(xy, inf) == [pyref.mul(C, k, P) if P is not None else None for k, P in zip(ks, Ps)]
    x, xinf = engine.mul_batch_x("p384", ks_bytes(ks), pxy, pinf)    # ECDH shape: x only
    assert np.array_equal(x, np.asarray(xy)[:, :NB]) and np.array_equal(xinf, inf)
    gxy, ginf = engine.mul_by_generator("p384", ks_bytes(ks))
    assert unpack(gxy, ginf) == [pyref.mul(C, k, pyref.G(C)) for k in ks]


def test_lincomb_and_point_sum(engine):
    rng = random.Random(11)
    for n in (0, 1, 2, 33, 257):
        ks = [rng.randrange(C.n) for _ in range(n)]
        Ps = rand_points(n, 2)
        if n > 2:
            Ps[1] = None
        pxy, pinf = pts_bytes(Ps)
        xy, inf = engine.lincomb("p384", ks_bytes(ks), pxy, pinf)
        want = pyref.lincomb(C, ks, [P for P in Ps]) if n else None
        acc = None
        for k, P in zip(ks, Ps):
            if P is not None:
                acc = pyref.add(C, acc, pyref.mul(C, k, P))
        assert pyref.dec_point(np.asarray(xy).tobytes(), inf, NB) == acc
    # config-5 shape: two partial sums (Jacobian) combined by point_sum
    ks = [rng.randrange(C.n) for _ in range(60)]
    Ps = rand_points(60, 4)
    pxy, pinf = pts_bytes(Ps)
    K = ks_bytes(ks)
    p1 = engine.lincomb_partial("p384", K[:NB * 25], pxy[:2 * NB * 25], pinf[:25])
    p2 = engine.lincomb_partial("p384", K[NB * 25:], pxy[2 * NB * 25:], pinf[25:])
    xy, inf = engine.point_sum("p384", np.concatenate([p1, p2]))
    acc = None
    for k, P in zip(ks, Ps):
        acc = pyref.add(C, acc, pyref.mul(C, k, P))
    assert pyref.dec_point(np.asarray(xy).tobytes(), inf, NB) == acc


def test_batch_normalize_both_projective_forms(engine):
    rng = random.Random(5)
    Ps = rand_points(200, 3)
    jac, hom, exp = [], [], []
    for P in Ps:
        z = rng.randrange(1, C.p)
        jac.append(((P[0] * z * z) % C.p).to_bytes(NB, "big") + ((P[1] * z * z * z) % C.p).to_bytes(NB, "big") + z.to_bytes(NB, "big"))
        hom.append(((P[0] * z) % C.p).to_bytes(NB, "big") + ((P[1] * z) % C.p).to_bytes(NB, "big") + z.to_bytes(NB, "big"))
        exp.append(P)
    ident = bytes(NB) + (1).to_bytes(NB, "big") + bytes(NB)
    jac.append(ident)
    hom.append(ident)
    exp.append(None)
    xy, inf = engine.batch_normalize("p384", np.frombuffer(b"".join(jac), np.uint8))
    assert unpack(xy, inf) == exp
    xy, inf = engine.batch_normalize_hom("p384", np.frombuffer(b"".join(hom), np.uint8))
    assert unpack(xy, inf) == exp


def test_field_ops(engine):
    p = C.p
    rng = random.Random(6)
    a = [0, 1, p - 1, p - 2, 2**383, 2**128, 2**96, 2**32] + [rng.randrange(p) for _ in range(1500)]
    b = [rng.randrange(p) for _ in a]
    A, B = ks_bytes(a), ks_bytes(b)
    model = {"add": lambda x, y: (x + y) % p, "sub": lambda x, y: (x - y) % p, "mul": lambda x, y: x * y % p,
             "neg": lambda x, y: (-x) % p, "sqr": lambda x, y: x * x % p, "inv": lambda x, y: pow(x, -1, p) if x else 0}
    for op, f in model.items():
        out = engine.field_op("p384", op, A, B if op in ("add", "sub", "mul") else None)
        got = [int.from_bytes(o.tobytes(), "big") for o in out]
        assert got == [f(x, y) for x, y in zip(a, b)], op
    with pytest.raises(ecgpu.NotOnCurveError):
        engine.field_op("p384", "add", ks_bytes([1, p]), ks_bytes([1, 1]))


def test_rejects_bad_inputs_and_unsupported_entries(engine):
    Ps = rand_points(40, 9)
    pxy, pinf = pts_bytes(Ps)
    ks = [5] * 40
    bad = list(ks)
    bad[17] = C.n
    with pytest.raises(ecgpu.ScalarRangeError) as ei:
        engine.mul_batch("p384", ks_bytes(bad), pxy, pinf)
    assert ei.value.index == 17
    off = pxy.copy()
    off[2 * NB * 9 + 2 * NB - 1] ^= 1
    with pytest.raises(ecgpu.NotOnCurveError) as ei:
        engine.mul_batch("p384", ks_bytes(ks), off, pinf)
    assert ei.value.index == 9
    # (SEC1 decompression, the field square root, a*G + b*P and ECDSA serve P-384 too: tests/test_sec1_ext.py, test_ecdsa_ext.py)


def test_large_batch_symmetry_and_sample(engine):
    """2^15 pairs (several pipelined chunks): k*P + (n-k)*P = O for every element, and a sample against the model."""
    n = 1 << 15
    rng = np.random.default_rng(384)
    K = rng.integers(0, 256, size=(n, NB), dtype=np.uint8)
    K[:, 0] &= 0x7F
    ks = [int.from_bytes(K[i].tobytes(), "big") for i in range(n)]
    Kneg = ks_bytes([C.n - k for k in ks])
    Ps = rand_points(n, 12)
    pxy, pinf = pts_bytes(Ps)
    xy, inf = engine.mul_batch("p384", K.reshape(-1), pxy, pinf)
    nxy, ninf = engine.mul_batch("p384", Kneg, pxy, pinf)
    xy, nxy = np.asarray(xy), np.asarray(nxy)
    assert not inf.any() and not ninf.any()
    assert np.array_equal(xy[:, :NB], nxy[:, :NB])                       # same x
    ysum = [(int.from_bytes(xy[i, NB:].tobytes(), "big") + int.from_bytes(nxy[i, NB:].tobytes(), "big")) % C.p for i in range(0, n, 97)]
    assert not any(ysum)                                                 # y + y' = p
    for i in range(0, n, 1021):
        assert pyref.dec_point(xy[i].tobytes(), 0, NB) == pyref.mul(C, ks[i], Ps[i])


def test_lincomb_bucket_method_p384(engine):
    """>= 2^13 terms take the bucket method (12-limb digits, 384-bit scalars); 16 distinct points repeated, so that the
    expected sum is 16 big-integer multiplications: sum_i k_i P_(i mod 16) = sum_j (sum_{i = j mod 16} k_i) P_j"""
    n = 9001
    rng = random.Random(99)
    ks = [rng.randrange(C.n) for _ in range(n)]
    Ps = rand_points(n, 21)
    pxy, pinf = pts_bytes(Ps)
    xy, inf = engine.lincomb("p384", ks_bytes(ks), pxy, pinf)
    acc = None
    for j in range(16):
        acc = pyref.add(C, acc, pyref.mul(C, sum(ks[j::16]) % C.n, Ps[j]))
    assert pyref.dec_point(np.asarray(xy).tobytes(), inf, NB) == acc
    # thousands of identical terms: the bucket method declines (device-side skew flag), the per-term path answers
    ks2 = [ks[0]] * n
    pxy2, pinf2 = pts_bytes([Ps[0]] * n)
    xy, inf = engine.lincomb("p384", ks_bytes(ks2), pxy2, pinf2)
    assert pyref.dec_point(np.asarray(xy).tobytes(), inf, NB) == pyref.mul(C, ks[0] * n % C.n, Ps[0])


def test_var_base_4096_pairs_vs_c_restatement(engine):
    """a batch large enough for several blocks per SM, every output against oracle/ecref_p384.c"""
    import ecref

    n = 4096
    rng = np.random.default_rng(4096)
    K = rng.integers(0, 256, size=(n, NB), dtype=np.uint8)
    K[:, 0] &= 0x7F
    pxy, pinf = pts_bytes(rand_points(n, 33))
    xy, inf = engine.mul_batch("p384", K.reshape(-1), pxy, pinf)
    r_xy, r_inf = ecref.mul_batch("p384", K.reshape(-1), pxy, pinf, nthreads=8)
    assert np.array_equal(np.asarray(xy).reshape(-1), r_xy.reshape(-1)) and np.array_equal(inf, r_inf)
    gxy, ginf = engine.mul_by_generator("p384", K.reshape(-1))
    r_xy, r_inf = ecref.mul_gen_batch("p384", K.reshape(-1), nthreads=8)
    assert np.array_equal(np.asarray(gxy).reshape(-1), r_xy.reshape(-1)) and np.array_equal(ginf, r_inf)
    lxy, linf = engine.lincomb("p384", K.reshape(-1), pxy, pinf)
    e_xy, e_inf = ecref.lincomb("p384", K.reshape(-1), pxy, pinf, nthreads=8)
    assert np.array_equal(np.asarray(lxy), e_xy) and linf == e_inf
