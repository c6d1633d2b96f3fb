"""GPU parity, through the C ABI, for the curves on the generic Montgomery field policy:
sm2, brainpoolP256r1 / t1, brainpoolP384r1 / t1, bign-curve256v1 (little-endian records), P-224, P-192, P-521 (66-byte records) — against the
C restatement of the reference's generic primeorder path (oracle/ecref_prime.c), the big-integer model, and the
reference's own vectors where the crate holds any (tests/golden/{p224,p192,p521,bignp256}.json)."""
import random

import numpy as np
import pytest

import ecgpu
import ecref
import pyref
from test_curves_ext import golden_points, pts, recs, unpack

pytestmark = pytest.mark.gpu
EXT = pyref.EXT_CURVES
IDS = sorted(EXT)


def rand_points(c, n, seed, distinct=16):
    rng = random.Random(seed)
    base = [pyref.mul(c, rng.randrange(1, c.n), pyref.G(c)) for _ in range(distinct)]
    return [base[i % distinct] for i in range(n)]


def rand_scalars(c, n, seed):
    """n uniformly random scalars below the order, as an (n, FB) byte array in the curve's byte order"""
    nb = pyref.fbytes(c)
    rng = np.random.default_rng(seed)
    K = rng.integers(0, 256, size=(n, nb), dtype=np.uint8)
    top = c.n >> (8 * (nb - 1))          # most significant byte of n: keep the scalar's top byte strictly below it
    msb = nb - 1 if c.le else 0
    K[:, msb] = K[:, msb] % max(top, 1)
    return K


@pytest.mark.parametrize("name", ["p224", "p192", "bignp256", "p521"])
def test_reference_vectors(engine, name):
    c = pyref.CURVES[name]
    vec = golden_points(name)
    ks = [k for k, _ in vec]
    want = [P for _, P in vec]
    xy, inf = engine.mul_by_generator(name, recs(c, ks))           # fixed-base table of the curve, built on the device
    assert unpack(c, xy, inf) == want
    pxy, pinf = pts(c, [pyref.G(c)] * len(ks))
    xy, inf = engine.mul_batch(name, recs(c, ks), pxy, pinf)       # variable-base kernel
    assert unpack(c, xy, inf) == want


@pytest.mark.parametrize("cid", IDS)
def test_var_base_fixed_base_and_x_only(engine, cid):
    c = EXT[cid]
    nb = pyref.fbytes(c)
    rng = random.Random(1000 + cid)
    n = 600
    ks = [rng.randrange(c.n) for _ in range(n)]
    ks[:10] = [0, 1, 2, c.n - 1, c.n - 2, 2**(c.n.bit_length() - 1), 2**64, (c.n - 1) // 2, 15, 16]
    Ps = rand_points(c, n, cid)
    Ps[5] = None
    Ps[6] = None
    K = recs(c, ks)
    pxy, pinf = pts(c, Ps)
    xy, inf = engine.mul_batch(c.name, K, pxy, pinf)
    r_xy, r_inf = ecref.mul_batch(c.name, K, pxy, pinf, nthreads=8)
    assert np.array_equal(np.asarray(xy).reshape(-1), r_xy.reshape(-1)) and np.array_equal(inf, r_inf)
    got = unpack(c, xy, inf)
    for i in list(range(12)) + [n - 1]:
        assert got[i] == (pyref.mul(c, ks[i], Ps[i]) if Ps[i] is not None else None)
    x, xinf = engine.mul_batch_x(c.name, K, pxy, pinf)              # ECDH shape: x only
    assert np.array_equal(x, np.asarray(xy).reshape(n, 2 * nb)[:, :nb]) and np.array_equal(xinf, inf)
    gxy, ginf = engine.mul_by_generator(c.name, K)
    r_xy, r_inf = ecref.mul_gen_batch(c.name, K, nthreads=8)
    assert np.array_equal(np.asarray(gxy).reshape(-1), r_xy.reshape(-1)) and np.array_equal(ginf, r_inf)
    assert unpack(c, gxy, ginf)[:4] == [pyref.mul(c, k, pyref.G(c)) for k in ks[:4]]


@pytest.mark.parametrize("cid", IDS)
def test_lincomb_point_sum_and_bucket_method(engine, cid):
    c = EXT[cid]
    nb = pyref.fbytes(c)
    rng = random.Random(2000 + cid)
    for n in (0, 1, 2, 33, 257):
        ks = [rng.randrange(c.n) for _ in range(n)]
        Ps = rand_points(c, n, 2)
        if n > 2:
            Ps[1] = None
        K = recs(c, ks) if n else np.zeros(0, np.uint8)
        pxy, pinf = pts(c, Ps) if n else (np.zeros(0, np.uint8), np.zeros(0, np.uint8))
        xy, inf = engine.lincomb(c.name, K, pxy, pinf)
        want = None
        for k, P in zip(ks, Ps):
            if P is not None:
                want = pyref.add(c, want, pyref.mul(c, k, P))
        assert unpack(c, xy, [inf]) == [want]
    # config-5 shape: two partial sums (Jacobian) combined by point_sum
    ks = [rng.randrange(c.n) for _ in range(60)]
    Ps = rand_points(c, 60, 4)
    K = recs(c, ks)
    pxy, pinf = pts(c, Ps)
    p1 = engine.lincomb_partial(c.name, K[:nb * 25], pxy[:2 * nb * 25], pinf[:25])
    p2 = engine.lincomb_partial(c.name, K[nb * 25:], pxy[2 * nb * 25:], pinf[25:])
    xy, inf = engine.point_sum(c.name, np.concatenate([p1, p2]))
    e_xy, e_inf = ecref.lincomb(c.name, K, pxy, pinf, nthreads=4)
    assert np.array_equal(np.asarray(xy), e_xy) and inf == e_inf
    # >= 2^13 terms: the bucket method (digits over the curve's scalar width), every term through the C restatement
    n = 9001
    K = rand_scalars(c, n, 3000 + cid).reshape(-1)
    Ps = rand_points(c, n, 21)
    Ps[77] = None
    pxy, pinf = pts(c, Ps)
    xy, inf = engine.lincomb(c.name, K, pxy, pinf)
    e_xy, e_inf = ecref.lincomb(c.name, K, pxy, pinf, nthreads=8)
    assert np.array_equal(np.asarray(xy), e_xy) and inf == e_inf


@pytest.mark.parametrize("cid", IDS)
def test_batch_normalize_and_field_ops(engine, cid):
    c = EXT[cid]
    p, nb = c.p, pyref.fbytes(c)
    rng = random.Random(3000 + cid)
    Ps = rand_points(c, 200, 3)
    jac, hom, exp = [], [], []
    for P in Ps:
        z = rng.randrange(1, p)
        jac.append(pyref.enc_fe(c, P[0] * z * z % p) + pyref.enc_fe(c, P[1] * z * z * z % p) + pyref.enc_fe(c, z))
        hom.append(pyref.enc_fe(c, P[0] * z % p) + pyref.enc_fe(c, P[1] * z % p) + pyref.enc_fe(c, z))
        exp.append(P)
    ident = pyref.enc_fe(c, 0) + pyref.enc_fe(c, 1) + pyref.enc_fe(c, 0)
    jac.append(ident)
    hom.append(ident)
    exp.append(None)
    xy, inf = engine.batch_normalize(c.name, np.frombuffer(b"".join(jac), np.uint8))
    assert unpack(c, xy, inf) == exp
    xy, inf = engine.batch_normalize_hom(c.name, np.frombuffer(b"".join(hom), np.uint8))
    assert unpack(c, xy, inf) == exp
    a = [0, 1, p - 1, p - 2, 2**(8 * nb - 9) % p, 2**64, 2**32] + [rng.randrange(p) for _ in range(700)]
    b = [rng.randrange(p) for _ in a]
    A, B = recs(c, a), recs(c, b)
    model = {"add": lambda x, y: (x + y) % p, "sub": lambda x, y: (x - y) % p, "mul": lambda x, y: x * y % p,
             "neg": lambda x, y: (-x) % p, "sqr": lambda x, y: x * x % p, "inv": lambda x, y: pow(x, -1, p) if x else 0}
    for op, f in model.items():
        out = engine.field_op(c.name, op, A, B if op in ("add", "sub", "mul") else None)
        got = [pyref.dec_fe(c, o.tobytes()) for o in out]
        assert got == [f(x, y) for x, y in zip(a, b)], op
    with pytest.raises(ecgpu.NotOnCurveError):
        engine.field_op(c.name, "add", recs(c, [1, p]), recs(c, [1, 1]))


@pytest.mark.parametrize("cid", IDS)
def test_rejects_bad_inputs_and_unsupported_entries(engine, cid):
    c = EXT[cid]
    nb = pyref.fbytes(c)
    Ps = rand_points(c, 40, 9)
    pxy, pinf = pts(c, Ps)
    ks = [5] * 40
    bad = list(ks)
    bad[17] = c.n
    with pytest.raises(ecgpu.ScalarRangeError) as ei:
        engine.mul_batch(c.name, recs(c, bad), pxy, pinf)
    assert ei.value.index == 17
    with pytest.raises(ecgpu.ScalarRangeError) as ei:
        engine.mul_by_generator(c.name, recs(c, bad))
    assert ei.value.index == 17
    off = pxy.copy()
    y9 = (Ps[9][1] + 1) % c.p
    off[2 * nb * 9 + nb:2 * nb * 10] = np.frombuffer(pyref.enc_fe(c, y9), np.uint8)
    with pytest.raises(ecgpu.NotOnCurveError) as ei:
        engine.mul_batch(c.name, recs(c, ks), off, pinf)
    assert ei.value.index == 9
    # (SEC1 decompression and the field square root: tests/test_sec1_ext.py; a*G + b*P and ECDSA: tests/test_ecdsa_ext.py)
    lib = engine.lib
    z = np.zeros(256, np.uint8)
    vp = lambda a: a.ctypes.data  # noqa: E731
    assert lib.ecg_mul_batch(engine._ctx, 12, 1, vp(z), vp(z), None, vp(z), vp(z)) == ecgpu.ECG_EINVAL   # unknown curve id


@pytest.mark.parametrize("cid", IDS)
def test_4096_pairs_vs_c_restatement(engine, cid):
    """a batch large enough for several blocks per SM (and two pipelined chunks), every output against oracle/ecref_prime.c"""
    c = EXT[cid]
    n = 4096
    K = rand_scalars(c, n, 4096 + cid).reshape(-1)
    pxy, pinf = pts(c, rand_points(c, n, 33))
    xy, inf = engine.mul_batch(c.name, K, pxy, pinf)
    r_xy, r_inf = ecref.mul_batch(c.name, K, pxy, pinf, nthreads=16)
    assert np.array_equal(np.asarray(xy).reshape(-1), r_xy.reshape(-1)) and np.array_equal(inf, r_inf)
    gxy, ginf = engine.mul_by_generator(c.name, K)
    r_xy, r_inf = ecref.mul_gen_batch(c.name, K, nthreads=16)
    assert np.array_equal(np.asarray(gxy).reshape(-1), r_xy.reshape(-1)) and np.array_equal(ginf, r_inf)


@pytest.mark.parametrize("cid", [4, 6, 9])
def test_large_batch_symmetry(engine, cid):
    """2^15 pairs (several pipelined chunks): k P and (n - k) P share x and have opposite y, for every element"""
    c = EXT[cid]
    nb = pyref.fbytes(c)
    n = 1 << 15
    K = rand_scalars(c, n, 5000 + cid)
    ks = [pyref.dec_fe(c, K[i].tobytes()) for i in range(n)]
    Kneg = recs(c, [(c.n - k) % c.n for k in ks])
    Ps = rand_points(c, n, 12)
    pxy, pinf = pts(c, Ps)
    xy, inf = engine.mul_batch(c.name, K.reshape(-1), pxy, pinf)
    nxy, ninf = engine.mul_batch(c.name, Kneg, pxy, pinf)
    xy, nxy = np.asarray(xy).reshape(n, 2 * nb), np.asarray(nxy).reshape(n, 2 * nb)
    assert np.array_equal(inf, ninf) and inf.sum() == sum(1 for k in ks if k == 0)
    assert np.array_equal(xy[:, :nb], nxy[:, :nb])
    ysum = [(pyref.dec_fe(c, xy[i, nb:].tobytes()) + pyref.dec_fe(c, nxy[i, nb:].tobytes())) % c.p for i in range(0, n, 97) if not inf[i]]
    assert not any(ysum)
    for i in range(0, n, 2047):
        assert unpack(c, xy[i], [inf[i]]) == [pyref.mul(c, ks[i], Ps[i])]
