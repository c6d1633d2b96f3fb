"""CPU-only: the CUDA *kernels* of ecg_kernels.cuh executed on the host, one simulated thread at a time (tests/sim/sim.cpp
supplies threadIdx/blockIdx and runs the same kernel sequences as ecgpu.cu), against the oracle and the reference's
golden vectors — including all of its Wycheproof ECDSA vectors.  This covers what `-m gpu` covers on the device
(validation flags and indices, exceptional additions, the verification front ends) in a container without a GPU.
tests/sim is test infrastructure; nothing here is reachable from the product."""
import ctypes
import hashlib
import json
import os
import random

import numpy as np
import pytest

import ecref
import pyref
from helpers import GOLDEN, edge_scalars, golden, pack_points, pack_scalars, random_points, unpack_points, wycheproof_cases
from test_sim import sim  # noqa: F401  (fixture: builds tests/sim/libecgsim.so when stale)

CID = {"k256": 0, "p256": 1}
FB_W, FB_WINDOWS = 16, 16
FB_ENTRIES = 1 << (FB_W - 1)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


@pytest.fixture(scope="module")
def fb_tables(sim):
    """The fixed-base table in the device layout (ecgpu.cu ensure_fb_table): entry (i, j) = (2j+1) * 2^(16 i) * G plus
    one entry 2^256 * G; the points come from the C oracle, the layout from affine_to_table_kernel."""
    out = {}
    for curve in ("k256", "p256"):
        c = pyref.CURVES[curve]
        ks = [((2 * j + 1) << (FB_W * i)) % c.n for i in range(FB_WINDOWS) for j in range(FB_ENTRIES)] + [(1 << 256) % c.n]
        K = np.frombuffer(b"".join(k.to_bytes(32, "big") for k in ks), np.uint8)
        xy, inf = ecref.mul_gen_batch(curve, K, nthreads=os.cpu_count() or 4)
        assert not np.asarray(inf).any()
        xy = np.ascontiguousarray(xy, np.uint8).reshape(-1)
        table = np.zeros(16 * len(ks), np.uint32)
        sim.simk_affine_to_table(CID[curve], ctypes.c_size_t(len(ks)), _p(xy), _p(table))
        out[curve] = table
    return out


def _mul_batch(sim, curve, K, xy, inf):
    n = K.size // 32
    oxy, oinf, st = np.zeros(64 * n, np.uint8), np.zeros(n, np.uint8), np.zeros(2, np.uint32)
    sim.simk_mul_batch(CID[curve], ctypes.c_size_t(n), _p(K), _p(xy), _p(inf), _p(oxy), _p(oinf), _p(st))
    return oxy.reshape(n, 64), oinf, st


@pytest.mark.parametrize("curve", ["k256", "p256"])
def test_varbase_kernel_golden_edges_and_validation(sim, curve):
    c = pyref.CURVES[curve]
    g = golden(curve)
    # the reference's MUL_TEST_VECTORS as k * G through the variable-base kernel
    ks = [int(v["k"], 16) for v in g["group"]["mul"]]
    xy, inf = pack_points([pyref.G(c)] * len(ks))
    oxy, oinf, st = _mul_batch(sim, curve, pack_scalars(ks), xy, inf)
    assert st[0] == 0
    for v, P in zip(g["group"]["mul"], unpack_points(oxy, oinf)):
        assert P == (int(v["x"], 16), int(v["y"], 16))
    # edge scalars x (random points, identity) against the big-integer model
    pts = random_points(c, 6, seed=4) + [None]
    ks = edge_scalars(c)
    rng = random.Random(5)
    ks += [rng.randrange(c.n) for _ in range(40)]
    Ps = [pts[i % len(pts)] for i in range(len(ks))]
    xy, inf = pack_points(Ps)
    oxy, oinf, st = _mul_batch(sim, curve, pack_scalars(ks), xy, inf)
    assert st[0] == 0
    assert unpack_points(oxy, oinf) == [pyref.mul(c, k, P) for k, P in zip(ks, Ps)]
    # validation: scalar >= n (flag 1) and a point off the curve (flag 2); status[1] = smallest offending index
    bad_k = pack_scalars(ks).copy()
    bad_k[32 * 7:32 * 8] = np.frombuffer(c.n.to_bytes(32, "big"), np.uint8)
    _, _, st = _mul_batch(sim, curve, bad_k, xy, inf)
    assert (st[0], st[1]) == (1, 7)
    bad_p = xy.copy().reshape(-1)
    bad_p[64 * 3 + 63] ^= 1
    _, _, st = _mul_batch(sim, curve, bad_k, bad_p, inf)
    assert (st[0], st[1]) == (3, 3)


@pytest.mark.parametrize("curve", ["k256", "p256"])
def test_fixedbase_and_mul_gen_add_kernels(sim, fb_tables, curve):
    c = pyref.CURVES[curve]
    g = golden(curve)
    table = fb_tables[curve]
    ks = [int(v["k"], 16) for v in g["group"]["mul"]] + edge_scalars(c)
    n = len(ks)
    oxy, oinf, st = np.zeros(64 * n, np.uint8), np.zeros(n, np.uint8), np.zeros(2, np.uint32)
    K = pack_scalars(ks)  # kept alive across the call (_p() only takes the address)
    sim.simk_mul_gen_batch(CID[curve], ctypes.c_size_t(n), _p(K), _p(table), _p(oxy), _p(oinf), _p(st))
    assert st[0] == 0
    got = unpack_points(oxy.reshape(n, 64), oinf)
    for v, P in zip(g["group"]["mul"], got):
        assert P == (int(v["x"], 16), int(v["y"], 16))
    assert got == [pyref.mul(c, k, pyref.G(c)) for k in ks]
    # a*G + b*P with the exceptional endings: a*G == b*P (doubling), a*G == -(b*P) (identity), a = 0, b = 0, P = O
    rng = random.Random(8)
    G = pyref.G(c)
    cases = []
    for _ in range(12):
        cases.append((rng.randrange(c.n), rng.randrange(c.n), pyref.mul(c, rng.randrange(1, c.n), G)))
    t = rng.randrange(1, c.n)
    P = pyref.mul(c, t, G)
    b = rng.randrange(1, c.n)
    cases += [(b * t % c.n, b, P), ((c.n - b * t) % c.n, b, P), (0, b, P), (b, 0, P), (b, t, None), (0, 0, P), (1, 1, G), (1, c.n - 1, G)]
    a_s = pack_scalars([x[0] for x in cases])
    b_s = pack_scalars([x[1] for x in cases])
    xy, inf = pack_points([x[2] for x in cases])
    n = len(cases)
    oxy, oinf = np.zeros(64 * n, np.uint8), np.zeros(n, np.uint8)
    sim.simk_mul_gen_add_batch(CID[curve], ctypes.c_size_t(n), _p(a_s), _p(b_s), _p(xy), _p(inf), _p(table), _p(oxy), _p(oinf), _p(st))
    assert st[0] == 0
    want = [pyref.add(c, pyref.mul(c, a, G), pyref.mul(c, b, P)) for a, b, P in cases]
    assert unpack_points(oxy.reshape(n, 64), oinf) == want
    assert want[13] is None and want[19] is None          # the cancelling cases really end at the identity


@pytest.mark.parametrize("curve", ["k256", "p256"])
def test_ecdsa_kernels_against_wycheproof_and_fips_vectors(sim, fb_tables, curve):
    """k256/src/ecdsa.rs:262-389 and p256/src/ecdsa.rs:166-169 (Wycheproof), */src/test_vectors/ecdsa.rs (FIPS / RFC6979)"""
    c = pyref.CURVES[curve]
    cases, rejected = wycheproof_cases(curve)
    assert not any(v["pass"] for v in rejected)            # a signature that does not parse is an expected failure
    assert len(cases) > 150 and sum(x[4] for x in cases) > 100
    for v in json.load(open(os.path.join(GOLDEN, f"{curve}_ecdsa.json")))["vectors"]:
        z = bytes.fromhex(v["m"])
        z = z if len(z) == 32 else hashlib.sha256(z).digest()
        r, s = int(v["r"], 16), int(v["s"], 16)
        q = (int(v["q_x"], 16), int(v["q_y"], 16))
        cases.append((z, r, s, q, pyref.ecdsa_verify(c, int.from_bytes(z, "big"), r, s, q)))
    low_s = 1 if curve == "k256" else 0                    # EcdsaCurve::NORMALIZE_S
    # raw range failures the parser would have stopped, to see the kernel's own checks: r = 0, s = 0, r = n, s = n
    z0, r0, s0, q0, _ = cases[0]
    extra = [(z0, 0, s0, q0), (z0, r0, 0, q0), (z0, c.n, s0, q0), (z0, r0, c.n, q0), (z0, r0, s0, (q0[0], q0[1] ^ 1))]
    n = len(cases) + len(extra)
    Z = np.frombuffer(b"".join(x[0] for x in cases) + b"".join(x[0] for x in extra), np.uint8)
    S = np.frombuffer(b"".join(x[1].to_bytes(32, "big") + x[2].to_bytes(32, "big") for x in cases + extra), np.uint8)
    Q = np.frombuffer(b"".join(x[3][0].to_bytes(32, "big") + x[3][1].to_bytes(32, "big") for x in cases + extra), np.uint8)
    valid = np.full(n, 7, np.uint8)
    sim.simk_ecdsa_verify_batch(CID[curve], ctypes.c_size_t(n), _p(Z), _p(S), _p(Q), low_s, _p(fb_tables[curve]), _p(valid))
    want = [int(x[4]) for x in cases] + [0] * len(extra)
    wrong = [i for i in range(n) if valid[i] != want[i]]
    assert not wrong, f"{curve}: kernel verdict differs from the reference's expectation at {wrong[:8]}"
    # the oracle agrees with every vector too (pins oracle/pyref.py ecdsa_verify to the Wycheproof set)
    for z, r, s, q, exp in cases:
        assert pyref.ecdsa_verify(c, int.from_bytes(z, "big"), r, s, q, low_s_only=bool(low_s)) == exp
    if curve == "k256":                                     # without normalize_s a high-s signature must be refused
        z, r, s, q, exp = next(x for x in cases if x[4])
        one = np.full(1, 7, np.uint8)
        hi = np.frombuffer(r.to_bytes(32, "big") + (c.n - s).to_bytes(32, "big"), np.uint8)
        zb = np.frombuffer(z, np.uint8)
        qb = np.frombuffer(q[0].to_bytes(32, "big") + q[1].to_bytes(32, "big"), np.uint8)
        sim.simk_ecdsa_verify_batch(0, ctypes.c_size_t(1), _p(zb), _p(hi), _p(qb), 1, _p(fb_tables[curve]), _p(one))
        assert one[0] == 0


def test_schnorr_kernels_against_bip340_vectors(sim, fb_tables):
    """k256/src/schnorr.rs BIP340 vectors (sign vectors verify; verify vectors give their stated verdict)"""
    vec = json.load(open(os.path.join(GOLDEN, "k256_bip340.json")))["vectors"]
    vec = [v for v in vec if len(bytes.fromhex(v["msg"])) == 32]
    n = len(vec)
    assert n >= 10
    PK = np.frombuffer(b"".join(bytes.fromhex(v["pk"]) for v in vec), np.uint8)
    M = np.frombuffer(b"".join(bytes.fromhex(v["msg"]) for v in vec), np.uint8)
    S = np.frombuffer(b"".join(bytes.fromhex(v["sig"]) for v in vec), np.uint8)
    valid = np.full(n, 7, np.uint8)
    sim.simk_schnorr_verify_batch(ctypes.c_size_t(n), _p(PK), _p(M), _p(S), _p(fb_tables["k256"]), _p(valid))
    assert [int(x) for x in valid] == [int(bool(v["valid"])) for v in vec]
    assert 0 in valid and 1 in valid


@pytest.mark.parametrize("curve", ["k256", "p256"])
def test_decompress_kernel(sim, curve):
    c = pyref.CURVES[curve]
    rng = random.Random(21)
    recs, want = [], []
    for i in range(60):
        x = pyref.G(c)[0] if i == 0 else rng.randrange(c.p)
        tag = 2 + (i & 1)
        recs.append(bytes([tag]) + x.to_bytes(32, "big"))
        rhs = (pow(x, 3, c.p) + c.a * x + c.b) % c.p
        y = pow(rhs, (c.p + 1) // 4, c.p)
        if y * y % c.p == rhs:
            want.append((x, y if (y & 1) == (tag & 1) else c.p - y))
        else:
            want.append(None)
    recs += [bytes(33), bytes([4]) + bytes(32), bytes([2]) + c.p.to_bytes(32, "big")]   # identity, bad tag, x >= p
    n = len(recs)
    R = np.frombuffer(b"".join(recs), np.uint8)
    oxy, oinf, valid = np.zeros(64 * n, np.uint8), np.zeros(n, np.uint8), np.full(n, 7, np.uint8)
    sim.simk_decompress_batch(CID[curve], ctypes.c_size_t(n), _p(R), _p(oxy), _p(oinf), _p(valid))
    for i, w in enumerate(want):
        assert bool(valid[i]) == (w is not None)
        if w:
            assert (int.from_bytes(oxy[64 * i:64 * i + 32].tobytes(), "big"), int.from_bytes(oxy[64 * i + 32:64 * i + 64].tobytes(), "big")) == w
    assert valid[60] == 1 and oinf[60] == 1 and valid[61] == 0 and valid[62] == 0
    assert sum(w is not None for w in want) > 15


@pytest.mark.parametrize("curve", ["k256", "p256"])
def test_lincomb_kernels_per_term_and_bucket_method(sim, curve):
    """ecg_lincomb's two kernel chains on the host (tests/sim/sim.cpp simk_lincomb mirrors lincomb_shard / msm_run):
    per-term kernel + tree sum, and the bucket method (prep, three-kernel scan, scatter, buckets, recursive weighted
    reduction, final) including its refusal of skewed inputs.  LinearCombination::lincomb, k256 mul.rs:66-175."""
    c = pyref.CURVES[curve]
    rng = random.Random(77)
    base = random_points(c, 24, seed=31) + [None]

    def run(ks, Ps, msm_min):
        n = len(ks)
        xy, inf = pack_points(Ps)
        oxy, oinf, stt = np.zeros(64, np.uint8), np.zeros(1, np.uint8), np.zeros(2, np.uint32)
        path = ctypes.c_int(-1)
        K = pack_scalars(ks)
        sim.simk_lincomb(CID[curve], ctypes.c_size_t(n), _p(K), _p(xy), _p(inf), ctypes.c_size_t(msm_min),
                         _p(oxy), _p(oinf), _p(stt), ctypes.byref(path))
        assert stt[0] == 0
        return pyref.dec_point(oxy.tobytes(), int(oinf[0])), path.value

    def want(ks, Ps):
        if len(ks) <= 40:
            return pyref.lincomb(c, ks, Ps)
        xy, inf = pack_points(Ps)
        rxy, rinf = ecref.lincomb(curve, pack_scalars(ks), xy, inf, nthreads=4)
        return pyref.dec_point(np.asarray(rxy).tobytes(), int(rinf))

    for n in (1, 2, 3, 33, 70):                                # per-term path, ragged tree sizes
        ks = [rng.randrange(c.n) for _ in range(n)]
        Ps = [base[rng.randrange(len(base))] for _ in range(n)]
        got, path = run(ks, Ps, 1 << 13)
        assert path == 0 and got == want(ks, Ps)
    for n in (300, 777):                                       # bucket method on distinct points
        ks = [rng.randrange(c.n) for _ in range(n)]
        ks[5], ks[6] = 0, c.n - 1
        pts = random_points(c, n - 1, seed=n) + [None]
        got, path = run(ks, pts, 64)
        assert path == 1 and got == want(ks, pts)
    # terms that cancel: k*P + (n-k)*P over the bucket path -> identity
    P = base[0]
    ks = [rng.randrange(1, c.n) for _ in range(100)]
    got, path = run(ks + [c.n - k for k in ks], [P] * 200, 64)
    assert path == 1 and got is None
    # the same point 6000 times with the same scalar: every digit lands in one bucket per window -> refused as skewed,
    # the per-term path answers
    k0 = rng.randrange(c.n)
    got, path = run([k0] * 6000, [P] * 6000, 64)
    assert path == 2 and got == pyref.mul(c, k0 * 6000 % c.n, P)
    got, path = run([], [], 64)
    assert got is None


@pytest.mark.parametrize("curve", ["k256", "p256"])
def test_varbase_kernel_scalar_sweeps(sim, curve):
    """Dense sweeps where the recoding and the exceptional additions bite: every scalar in [0, 1500) and (n-1500, n),
    neighbours of multiples of lambda (k256: one GLV half tiny or zero) and of powers of two, times G and a random
    point — against the C oracle (bit-exact) and, on a sample, the big-integer model."""
    c = pyref.CURVES[curve]
    rng = random.Random(101)
    ks = list(range(1500)) + [c.n - i for i in range(1, 1500)]
    for e in range(8, 257, 8):
        ks += [(1 << e) % c.n, ((1 << e) - 1) % c.n, ((1 << e) + 1) % c.n]
    if curve == "k256":
        lam = 0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72
        for m in range(1, 40):
            for d in (-2, -1, 0, 1, 2):
                ks += [(m * lam + d) % c.n, (c.n - m * lam + d) % c.n]
    P = pyref.mul(c, rng.randrange(1, c.n), pyref.G(c))
    for pt in (pyref.G(c), P):
        xy, inf = pack_points([pt] * len(ks))
        K = pack_scalars(ks)
        oxy, oinf, st = _mul_batch(sim, curve, K, xy, inf)
        assert st[0] == 0
        rxy, rinf = ecref.mul_batch(curve, K, xy, inf, nthreads=os.cpu_count() or 4)
        assert np.array_equal(oxy.reshape(-1), np.asarray(rxy).reshape(-1)) and np.array_equal(oinf, rinf)
        got = unpack_points(oxy, oinf)
        for i in list(range(0, 40)) + list(range(1500, 1540)) + list(range(len(ks) - 30, len(ks))):
            assert got[i] == pyref.mul(c, ks[i], pt)


@pytest.mark.parametrize("curve", ["k256", "p256"])
def test_fixedbase_kernel_window_patterns(sim, fb_tables, curve):
    """The fixed-base recoding uses signed odd 16-bit digits with a carry into the next window and one 2^256 entry:
    sweep every window through its boundary values (0000, 0001, 7fff, 8000, 8001, ffff) with zero / all-ones / random
    neighbours, plus small and near-order scalars, against the C oracle."""
    c = pyref.CURVES[curve]
    rng = random.Random(202)
    ks = list(range(300)) + [c.n - i for i in range(1, 300)]
    pats = (0x0000, 0x0001, 0x7FFF, 0x8000, 0x8001, 0xFFFF, 0xFFFE)
    for w in range(16):
        for v in pats:
            for fill in (0, (1 << 256) - 1, rng.getrandbits(256)):
                k = (fill & ~(0xFFFF << (16 * w))) | (v << (16 * w))
                ks.append(k % c.n)
    n = len(ks)
    K = pack_scalars(ks)
    oxy, oinf, st = np.zeros(64 * n, np.uint8), np.zeros(n, np.uint8), np.zeros(2, np.uint32)
    sim.simk_mul_gen_batch(CID[curve], ctypes.c_size_t(n), _p(K), _p(fb_tables[curve]), _p(oxy), _p(oinf), _p(st))
    assert st[0] == 0
    rxy, rinf = ecref.mul_gen_batch(curve, K, nthreads=os.cpu_count() or 4)
    assert np.array_equal(oxy, np.asarray(rxy).reshape(-1)) and np.array_equal(oinf, rinf)
    assert [int(x) for x in oinf] == [int(k == 0) for k in ks]   # the identity exactly where k = 0


@pytest.mark.parametrize("curve", ["k256", "p256"])
@pytest.mark.parametrize("bucket_k", [4, 8])
def test_lincomb_warp_balanced_bucket_kernel(sim, curve, bucket_k):
    """msm_bucket_sorted_kernel (ECG_MSM_BUCKETS_PER_THREAD = 4 | 8 in the product): a block orders its 128*K buckets by
    size in shared memory; the sums must land in the same slots, so the lincomb result is unchanged."""
    c = pyref.CURVES[curve]
    rng = random.Random(500 + bucket_k)
    for n in (300, 1100):
        ks = [rng.randrange(c.n) for _ in range(n)]
        ks[3] = 0
        # a few popular points make some buckets much larger than others
        pts = random_points(c, 40, seed=n + bucket_k) + [None]
        Ps = [pts[min(rng.randrange(60), 40)] for _ in range(n)]
        xy, inf = pack_points(Ps)
        K = pack_scalars(ks)
        oxy, oinf, stt = np.zeros(64, np.uint8), np.zeros(1, np.uint8), np.zeros(2, np.uint32)
        path = ctypes.c_int(-1)
        sim.simk_lincomb_k(CID[curve], ctypes.c_size_t(n), _p(K), _p(xy), _p(inf), ctypes.c_size_t(64), _p(oxy), _p(oinf), _p(stt),
                           ctypes.byref(path), bucket_k)
        assert stt[0] == 0 and path.value == 1
        rxy, rinf = ecref.lincomb(curve, K, xy, inf, nthreads=4)
        assert np.array_equal(oxy, np.asarray(rxy).reshape(-1)) and int(oinf[0]) == int(rinf)


def test_p384_varbase_kernel_chain(sim):
    """P-384 through the same kernels (generic_varbase_kernel<CurveP384> -> normalize_kernel<FpP384>), 48-byte records:
    the reference's MUL_TEST_VECTORS (p384/src/test_vectors/group.rs:175), edge scalars, identity inputs, validation."""
    c = pyref.P384
    g = golden("p384")
    rng = random.Random(384)
    ks = [int(v["k"], 16) for v in g["group"]["mul"]] + [0, 1, c.n - 1, 2**383, rng.randrange(c.n)]
    Ps = [pyref.G(c)] * len(g["group"]["mul"]) + [pyref.mul(c, 7, pyref.G(c))] * 3 + [None, pyref.mul(c, rng.randrange(1, c.n), pyref.G(c))]
    n = len(ks)
    K = np.frombuffer(b"".join(k.to_bytes(48, "big") for k in ks), np.uint8).copy()
    xy = np.frombuffer(b"".join(pyref.enc_point(P, 48)[0] for P in Ps), np.uint8).copy()
    inf = np.array([1 if P is None else 0 for P in Ps], np.uint8)
    oxy, oinf, st = np.zeros(96 * n, np.uint8), np.zeros(n, np.uint8), np.zeros(2, np.uint32)
    sim.simk_mul_batch(2, ctypes.c_size_t(n), _p(K), _p(xy), _p(inf), _p(oxy), _p(oinf), _p(st))
    assert st[0] == 0
    got = [pyref.dec_point(oxy[96 * i:96 * i + 96].tobytes(), int(oinf[i]), 48) for i in range(n)]
    assert got == [pyref.mul(c, k, P) if P is not None else None for k, P in zip(ks, Ps)]
    for i, v in enumerate(g["group"]["mul"]):
        assert got[i] == (int(v["x"], 16), int(v["y"], 16))
    # validation: scalar = n at index 3, off-curve point at index 1 -> flags, smallest index
    Kb = K.copy()
    Kb[48 * 3:48 * 4] = np.frombuffer(c.n.to_bytes(48, "big"), np.uint8)
    xyb = xy.copy()
    xyb[96 * 1 + 95] ^= 1
    sim.simk_mul_batch(2, ctypes.c_size_t(n), _p(Kb), _p(xyb), _p(inf), _p(oxy), _p(oinf), _p(st))
    assert st[0] == 3 and st[1] == 1


def test_p384_lincomb_bucket_method_chain(sim):
    """the bucket-method chain (prep -> scan -> scatter -> order -> buckets -> window reduction -> Horner) with 12-limb
    scalars and points: 384-bit digits, 48 windows at c = 8"""
    c = pyref.P384
    rng = random.Random(12)
    n = 300
    base = [pyref.mul(c, rng.randrange(1, c.n), pyref.G(c)) for _ in range(6)]
    ks = [rng.randrange(c.n) for _ in range(n)]
    ks[:4] = [0, 1, c.n - 1, 2**383]
    Ps = [base[i % 6] for i in range(n)]
    Ps[7] = None
    K = np.frombuffer(b"".join(k.to_bytes(48, "big") for k in ks), np.uint8).copy()
    xy = np.frombuffer(b"".join(pyref.enc_point(P, 48)[0] for P in Ps), np.uint8).copy()
    inf = np.array([1 if P is None else 0 for P in Ps], np.uint8)
    want = None
    for k, P in zip(ks, Ps):
        if P is not None:
            want = pyref.add(c, want, pyref.mul(c, k, P))
    for min_terms, path_want in ((64, 1), (10**6, 0)):  # bucket method / per-term path
        oxy, oinf, st, path = np.zeros(96, np.uint8), np.zeros(1, np.uint8), np.zeros(2, np.uint32), ctypes.c_int(-1)
        sim.simk_lincomb(2, ctypes.c_size_t(n), _p(K), _p(xy), _p(inf), ctypes.c_size_t(min_terms), _p(oxy), _p(oinf), _p(st), ctypes.byref(path))
        assert st[0] == 0 and path.value == path_want
        assert pyref.dec_point(oxy.tobytes(), int(oinf[0]), 48) == want
