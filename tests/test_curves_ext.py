"""CPU-only parity for the curves on the generic Montgomery field policy: sm2,
brainpoolP256r1/t1, brainpoolP384r1/t1, bign-curve256v1, P-224, P-192, P-521.

 * oracle pinning: the big-integer model (oracle/pyref.py) and the C restatement of the reference's generic primeorder
   path (oracle/ecref_prime.c) against the reference's own vectors where it holds any (p224 / p192 / p521 / bignp256
   src/test_vectors/group.rs -> tests/golden/*.json), against OpenSSL where it knows the curve, and against each other;
 * the device code itself, executed on the host (tests/sim): field policy (ecg_fe_mont.cuh), the general-a formulas,
   the variable-base / fixed-base / bucket-method kernel chains with the records in each curve's byte order."""
import ctypes
import os
import random

import numpy as np
import pytest

import ecref
import pyref
from helpers import golden

EXT = pyref.EXT_CURVES
IDS = sorted(EXT)


def rec(c, v):
    return pyref.enc_fe(c, v)


def recs(c, vs):
    return np.frombuffer(b"".join(rec(c, v) for v in vs), np.uint8).copy()


def pts(c, Ps):
    nb = pyref.fbytes(c)
    xy = np.frombuffer(b"".join((rec(c, P[0]) + rec(c, P[1])) if P is not None else bytes(2 * nb) for P in Ps), np.uint8).copy()
    return xy, np.array([1 if P is None else 0 for P in Ps], np.uint8)


def unpack(c, xy, inf):
    nb = pyref.fbytes(c)
    xy = np.asarray(xy, np.uint8).reshape(-1, 2 * nb)
    return [None if inf[i] else (pyref.dec_fe(c, xy[i, :nb].tobytes()), pyref.dec_fe(c, xy[i, nb:].tobytes())) for i in range(xy.shape[0])]


def golden_points(name):
    """[(k, (x, y))] from the reference's vector file of that curve (records in the curve's byte order)"""
    c = pyref.CURVES[name]
    g = golden(name)["group"]
    out = [(v["k"], (pyref.dec_fe(c, bytes.fromhex(v["x"])), pyref.dec_fe(c, bytes.fromhex(v["y"])))) for v in g["add"]]
    out += [(pyref.dec_fe(c, bytes.fromhex(v["k"])), (pyref.dec_fe(c, bytes.fromhex(v["x"])), pyref.dec_fe(c, bytes.fromhex(v["y"]))))
            for v in g["mul"]]
    return out


# ---------------------------------------------------------------------------------------------------------------------
# oracle pinning

@pytest.mark.parametrize("name", ["p224", "p192", "bignp256", "p521"])
def test_oracles_reproduce_the_reference_vectors(name):
    c = pyref.CURVES[name]
    vec = golden_points(name)
    assert len(vec) >= 40
    for k, P in vec:
        assert pyref.mul(c, k, pyref.G(c)) == P
    xy, inf = ecref.mul_gen_batch(name, recs(c, [k for k, _ in vec]), nthreads=4)
    assert unpack(c, xy, inf) == [P for _, P in vec]
    pxy, pinf = pts(c, [pyref.G(c)] * len(vec))
    xy, inf = ecref.mul_batch(name, recs(c, [k for k, _ in vec]), pxy, pinf, nthreads=4)
    assert unpack(c, xy, inf) == [P for _, P in vec]


@pytest.mark.parametrize("cid", IDS)
def test_curve_constants_and_c_restatement_vs_model(cid):
    c = EXT[cid]
    G = pyref.G(c)
    assert pyref.on_curve(c, G) and pyref.mul(c, c.n - 1, G) == pyref.neg(c, G)
    rng = random.Random(cid)
    ks = [0, 1, 2, 3, c.n - 1, c.n - 2, 7, 8, 9, 15, 16, 17, (c.n - 1) // 2] + [rng.randrange(c.n) for _ in range(20)]
    Ps = [pyref.mul(c, rng.randrange(1, c.n), G) for _ in ks]
    Ps[4] = None
    pxy, pinf = pts(c, Ps)
    xy, inf = ecref.mul_batch(c.name, recs(c, ks), pxy, pinf, nthreads=4)
    assert unpack(c, xy, inf) == [pyref.mul(c, k, P) if P is not None else None for k, P in zip(ks, Ps)]
    oxy, oinf = ecref.lincomb(c.name, recs(c, ks), pxy, pinf, nthreads=3)
    want = None
    for k, P in zip(ks, Ps):
        if P is not None:
            want = pyref.add(c, want, pyref.mul(c, k, P))
    assert unpack(c, oxy, [oinf]) == [want]
    with pytest.raises(ValueError):   # Scalar::from_repr rejects k >= n
        ecref.mul_gen_batch(c.name, recs(c, [c.n]))
    bad = list(Ps[0])
    bad[1] = (bad[1] + 1) % c.p
    with pytest.raises(ValueError):   # AffinePoint::from_coordinates rejects an off-curve point
        ecref.mul_batch(c.name, recs(c, [5]), *pts(c, [tuple(bad)]))


@pytest.mark.parametrize("name,ossl", [("bp256r1", "BrainpoolP256R1"), ("bp384r1", "BrainpoolP384R1"), ("p224", "SECP224R1"), ("p192", "SECP192R1"), ("p521", "SECP521R1")])
def test_model_agrees_with_openssl(name, ossl):
    ec = pytest.importorskip("cryptography.hazmat.primitives.asymmetric.ec")
    c = pyref.CURVES[name]
    rng = random.Random(77)
    try:
        for _ in range(6):
            k = rng.randrange(1, c.n)
            pub = ec.derive_private_key(k, getattr(ec, ossl)()).public_key().public_numbers()
            assert pyref.mul(c, k, pyref.G(c)) == (pub.x, pub.y)
    except Exception as e:  # a build of OpenSSL without the curve
        if "unsupported" in str(e).lower() or "UnsupportedAlgorithm" in type(e).__name__:
            pytest.skip(f"OpenSSL does not provide {ossl}")
        raise


# ---------------------------------------------------------------------------------------------------------------------
# the device code on the host

@pytest.fixture(scope="module")
def sim():
    import __graft_entry__ as ge
    ge.build()
    lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sim", "libecgsim.so"))
    return lib


def _p(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)


@pytest.mark.parametrize("cid", IDS)
def test_field_policy_on_host(sim, cid):
    c = EXT[cid]
    p, nb = c.p, pyref.fbytes(c)
    rng = random.Random(100 + cid)

    def op(o, a, b=0):
        out = ctypes.create_string_buffer(nb)
        assert sim.sim_ext_fe_op(cid, o, rec(c, a), rec(c, b), out) == 0
        return pyref.dec_fe(c, out.raw)

    top = 1 << (p.bit_length() - 1)
    # the last four: bit patterns around the fold / rotation boundaries of the Mersenne reduction (P-521: 23-bit rotation)
    vals = [0, 1, 2, 3, p - 1, p - 2, (p + 1) // 2, (p - 1) // 2, 2**(8 * nb - 8) % p, (1 << 23) - 1, p - (1 << 23), top, top - 1] + \
           [rng.randrange(p) for _ in range(60)]
    for a in vals:
        for b in rng.sample(vals, 5):
            assert op(0, a, b) == (a + b) % p
            assert op(1, a, b) == (a - b) % p
            assert op(2, a, b) == a * b % p
        assert op(3, a) == a * a % p
        assert op(4, a) == (-a) % p
        assert op(5, a) == a * pow(2, -1, p) % p
        assert op(6, a) == 3 * a % p and op(9, a) == 8 * a % p and op(8, a) == a
    for a in vals[:20]:
        assert op(7, a) == (pow(a, -1, p) if a else 0)
    out = ctypes.create_string_buffer(2 * nb)
    sim.sim_ext_generator(cid, out)
    assert (pyref.dec_fe(c, out.raw[:nb]), pyref.dec_fe(c, out.raw[nb:])) == pyref.G(c)


def _mul_batch(sim, cid, c, ks, Ps):
    n = len(ks)
    nb = pyref.fbytes(c)
    pxy, pinf = pts(c, Ps)
    K = recs(c, ks)
    oxy, oinf, st = np.zeros(2 * nb * n, np.uint8), np.zeros(n, np.uint8), np.zeros(2, np.uint32)
    sim.simk_mul_batch(cid, ctypes.c_size_t(n), _p(K), _p(pxy), _p(pinf), _p(oxy), _p(oinf), _p(st))
    return oxy, oinf, st


@pytest.mark.parametrize("cid", IDS)
def test_varbase_kernel_chain(sim, cid):
    """generic_varbase_kernel + normalize_kernel: reference vectors where the crate has them, edge scalars, identities,
    the input validation with the smallest offending index"""
    c = EXT[cid]
    G = pyref.G(c)
    rng = random.Random(200 + cid)
    ks = [0, 1, 2, 3, c.n - 1, c.n - 2, 15, 16, 17, 2**(c.n.bit_length() - 1), (c.n - 1) // 2, (c.n + 1) // 2] + [rng.randrange(c.n) for _ in range(30)]
    base = [pyref.mul(c, rng.randrange(1, c.n), G) for _ in range(5)] + [G]
    Ps = [base[i % 6] for i in range(len(ks))]
    Ps[7] = None
    want = [pyref.mul(c, k, P) if P is not None else None for k, P in zip(ks, Ps)]
    if c.name in ("p224", "p192", "bignp256", "p521"):
        for k, P in golden_points(c.name):
            ks.append(k)
            Ps.append(G)
            want.append(P)
    oxy, oinf, st = _mul_batch(sim, cid, c, ks, Ps)
    assert st[0] == 0
    assert unpack(c, oxy, oinf) == want
    # against the C restatement as well (different formulas, same affine values)
    pxy, pinf = pts(c, Ps)
    rxy, rinf = ecref.mul_batch(c.name, recs(c, ks), pxy, pinf, nthreads=4)
    assert np.array_equal(rxy.reshape(-1), oxy) and np.array_equal(rinf, oinf)
    # validation
    bad_k = list(ks)
    bad_k[9] = c.n
    bad_P = list(Ps)
    bp = list(bad_P[3])
    bp[0] = (bp[0] + 1) % c.p
    bad_P[3] = tuple(bp)
    _, _, st = _mul_batch(sim, cid, c, bad_k, Ps)
    assert st[0] == 1 and st[1] == 9
    _, _, st = _mul_batch(sim, cid, c, bad_k, bad_P)
    assert st[0] == 3 and st[1] == 3


_FB_CACHE = {}


def ext_fb_table(sim, cid):
    """the fixed-base table as ensure_fb_table (ecgpu.cu) builds it: entry (i, j) = (2j + 1) 2^(16 i) G, then 2^(32 NL) G, here
    computed by the C restatement and converted by affine_to_table_kernel; built once per test session (2^15 * 2 NL + 1 points)"""
    if cid not in _FB_CACHE:
        c = EXT[cid]
        nl = (pyref.fbytes(c) + 3) // 4
        ks = [((2 * j + 1) << (16 * i)) % c.n for i in range(2 * nl) for j in range(1 << 15)] + [(1 << (32 * nl)) % c.n]
        xy, inf = ecref.mul_gen_batch(c.name, recs(c, ks), nthreads=os.cpu_count() or 4)
        assert not inf.any()
        table = np.zeros(len(ks) * 2 * nl, np.uint32)
        flat = np.ascontiguousarray(xy).reshape(-1)
        sim.simk_affine_to_table(cid, ctypes.c_size_t(len(ks)), _p(flat), _p(table))
        _FB_CACHE[cid] = table
    return _FB_CACHE[cid]


@pytest.fixture(scope="module")
def fb_tables(sim):
    return lambda cid: ext_fb_table(sim, cid)


@pytest.mark.parametrize("cid", [3, 4, 6, 9])   # one a = -3, one general-a, the little-endian one, the 7-limb one
def test_fixedbase_kernel_chain(sim, fb_tables, cid):
    c = EXT[cid]
    nb = pyref.fbytes(c)
    table = fb_tables(cid)
    rng = random.Random(300 + cid)
    ks = [0, 1, 2, 3, c.n - 1, c.n - 2, 2**16 - 1, 2**16, 2**16 + 1, 2**32, 2**(8 * nb - 1) % c.n] + [rng.randrange(c.n) for _ in range(150)]
    if c.name in ("p224", "bignp256"):
        ks += [k for k, _ in golden_points(c.name)]
    n = len(ks)
    oxy, oinf, st = np.zeros(2 * nb * n, np.uint8), np.zeros(n, np.uint8), np.zeros(2, np.uint32)
    K = recs(c, ks)   # keep the array alive across the call: _p() only takes its address
    sim.simk_mul_gen_batch(cid, ctypes.c_size_t(n), _p(K), _p(table), _p(oxy), _p(oinf), _p(st))
    assert st[0] == 0
    rxy, rinf = ecref.mul_gen_batch(c.name, recs(c, ks), nthreads=4)
    assert np.array_equal(rxy.reshape(-1), oxy) and np.array_equal(rinf, oinf)
    got = unpack(c, oxy, oinf)
    for i in list(range(12)) + [n - 1]:
        assert got[i] == pyref.mul(c, ks[i], pyref.G(c))


@pytest.mark.parametrize("cid", [4, 5, 6, 8, 10, 11])
def test_lincomb_kernel_chain_per_term_and_bucket_method(sim, cid):
    c = EXT[cid]
    nb = pyref.fbytes(c)
    rng = random.Random(400 + cid)
    base = [pyref.mul(c, rng.randrange(1, c.n), pyref.G(c)) for _ in range(12)]

    def run(ks, Ps, msm_min):
        n = len(ks)
        pxy, pinf = pts(c, Ps)
        K = recs(c, ks)
        oxy, oinf, st, path = np.zeros(2 * nb, np.uint8), np.zeros(1, np.uint8), np.zeros(2, np.uint32), ctypes.c_int(0)
        sim.simk_lincomb(cid, ctypes.c_size_t(n), _p(K), _p(pxy), _p(pinf), ctypes.c_size_t(msm_min), _p(oxy), _p(oinf), _p(st),
                         ctypes.byref(path))
        assert st[0] == 0
        rxy, rinf = ecref.lincomb(c.name, recs(c, ks), pxy, pinf, nthreads=4)
        assert np.array_equal(rxy, oxy) and rinf == int(oinf[0])
        return unpack(c, oxy, oinf)[0], path.value

    for n, msm_min, want_path in ((0, 8192, 0), (1, 8192, 0), (37, 8192, 0), (300, 64, 1)):
        ks = [rng.randrange(c.n) for _ in range(n)]
        Ps = [base[rng.randrange(12)] for _ in range(n)]
        if n > 5:
            Ps[2] = None
            ks[4] = 0
            ks[5] = c.n - 1
        got, path = run(ks, Ps, msm_min)
        assert path == want_path
        if n <= 37:
            want = None
            for k, P in zip(ks, Ps):
                if P is not None:
                    want = pyref.add(c, want, pyref.mul(c, k, P))
            assert got == want
    # cancellation to the identity through the bucket method: sum k_i P + (n - sum k_i) P = O
    ks = [rng.randrange(c.n) for _ in range(199)]
    ks.append((-sum(ks)) % c.n)
    got, path = run(ks, [base[0]] * 200, 64)
    assert got is None
