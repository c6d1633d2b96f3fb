"""CPU-only: the device arithmetic headers compiled for the host (PTX carry flags emulated) vs the big-int model.
This exercises exactly the code the CUDA kernels run — field ops on the weakly-reduced range, GLV split,
signed-odd recoding, window tables, exceptional cases — without a GPU.  tests/sim is test infrastructure."""
import ctypes
import os
import random
import subprocess

import pytest

import pyref
from helpers import edge_scalars, golden, random_points

HERE = os.path.dirname(os.path.abspath(__file__))
SIM_SO = os.path.join(HERE, "sim", "libecgsim.so")


@pytest.fixture(scope="module")
def sim():
    src = os.path.join(HERE, "sim", "sim.cpp")
    csrc = os.path.join(HERE, "..", "elliptic-curves_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if not os.path.exists(SIM_SO) or any(os.path.getmtime(d) > os.path.getmtime(SIM_SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-shared", "-fPIC", "-pthread", "-o", SIM_SO, src])
    return ctypes.CDLL(SIM_SO)


def _feop(sim, curve, op, a, b=0):
    out = ctypes.create_string_buffer(32)
    getattr(sim, f"sim_{curve}_fe_op")(op, a.to_bytes(32, "big"), b.to_bytes(32, "big"), out)
    return int.from_bytes(out.raw, "big")


def _mont(sim, curve):
    """(R^-1, R^2) mod p when the curve's internal form is the Montgomery domain (raw values x stand for x/R), else (1, 1)"""
    p = pyref.CURVES[curve].p
    if curve == "p256" and sim.sim_p256_is_mont():
        return pow(2**256, -1, p), pow(2**256, 2, p)
    return 1, 1


@pytest.mark.parametrize("curve", ["k256", "p256"])
def test_field_ops_full_256bit_range(sim, curve):
    p = pyref.CURVES[curve].p
    ri, r2 = _mont(sim, curve)
    rng = random.Random(7)
    edge = [0, 1, 2, p - 1, p, p + 1, 2**256 - 1, 2**256 - 2, 2**256 - p, 2**256 - p - 1, p - 2, (p + 1) // 2,
            2**255, 2**128, 2**224, 2**192, 2**96, 2**224 - 1, 2**256 - 2**224]
    vals = edge + [rng.randrange(2**256) for _ in range(150)]
    for a in vals:
        for b in rng.sample(vals, 8) + edge[:8]:
            assert _feop(sim, curve, 0, a, b) == (a + b) % p
            assert _feop(sim, curve, 1, a, b) == (a - b) % p
            assert _feop(sim, curve, 2, a, b) == a * b * ri % p
        assert _feop(sim, curve, 3, a) == a * a * ri % p
        assert _feop(sim, curve, 4, a) == (-a) % p
        assert _feop(sim, curve, 5, a) == a * pow(2, -1, p) % p
        assert _feop(sim, curve, 6, a) == 3 * a % p
        assert _feop(sim, curve, 9, a) == 8 * a % p
        assert _feop(sim, curve, 8, a) == a % p
    for a in vals[:30]:
        assert _feop(sim, curve, 7, a) == (pow(a % p, -1, p) * r2 % p if a % p else 0)
    if curve == "p256":  # boundary conversions: canonical in, canonical out
        for a in [v % p for v in vals]:
            assert _feop(sim, curve, 10, a) == a
            b = rng.randrange(p)
            assert _feop(sim, curve, 11, a, b) == a * b % p
        for a in [v % p for v in vals[:12]]:
            assert _feop(sim, curve, 12, a) == (pow(a, -1, p) if a else 0)


def test_field_mul_extremes(sim):
    # all-ones style operands maximise every carry chain
    xs = [2**256 - 1, 2**256 - 2**32, 0xFFFFFFFF << 224, (2**256 - 1) ^ (0xFFFFFFFF << 96), int("f0" * 32, 16), int("0f" * 32, 16)]
    for a in xs:
        for b in xs:
            assert _feop(sim, "k256", 2, a, b) == a * b % pyref.K256.p
            assert _feop(sim, "p256", 2, a, b) == a * b * _mont(sim, "p256")[0] % pyref.P256.p


def test_field_mul_structured_stress(sim):
    """limbs drawn from {0, 1, ~0, ~0-1, random}: drives every carry / overflow-fold path of both reductions"""
    rng = random.Random(123)

    def structured():
        v = 0
        for i in range(8):
            c = rng.random()
            w = 0 if c < 0.3 else 0xFFFFFFFF if c < 0.6 else 1 if c < 0.65 else 0xFFFFFFFE if c < 0.7 else rng.getrandbits(32)
            v |= w << (32 * i)
        return v

    for curve in ("k256", "p256"):
        p = pyref.CURVES[curve].p
        ri = _mont(sim, curve)[0]
        for _ in range(3000):
            a, b = structured(), structured()
            assert _feop(sim, curve, 2, a, b) == a * b * ri % p
            assert _feop(sim, curve, 3, a) == a * a * ri % p


def test_karatsuba_multiplier_variant(sim):
    rng = random.Random(321)

    def structured():
        v = 0
        for i in range(8):
            c = rng.random()
            w = 0 if c < 0.3 else 0xFFFFFFFF if c < 0.6 else 1 if c < 0.65 else 0xFFFFFFFE if c < 0.7 else rng.getrandbits(32)
            v |= w << (32 * i)
        return v

    for ci, curve in enumerate(("k256", "p256")):
        p = pyref.CURVES[curve].p
        for i in range(4000):
            a, b = (structured(), structured()) if i % 2 else (rng.getrandbits(256), rng.getrandbits(256))
            out = ctypes.create_string_buffer(32)
            assert sim.sim_fe_mul_kara(ci, a.to_bytes(32, "big"), b.to_bytes(32, "big"), out) == 0
            assert int.from_bytes(out.raw, "big") == a * b % p


def test_glv_split_matches_reference_definition(sim):
    rng = random.Random(2)
    n = pyref.K256.n
    for k in [0, 1, 2, n - 1, n - 2, pyref.K256_LAMBDA] + [rng.randrange(n) for _ in range(1500)]:
        out = ctypes.create_string_buffer(36)
        assert sim.sim_k256_glv(k.to_bytes(32, "big"), out) == 0
        k1, k2 = pyref.glv_split(k)
        for off, kk in ((0, k1), (18, k2)):
            h = int.from_bytes(out.raw[off:off + 16], "little")
            neg, ev = out.raw[off + 16], out.raw[off + 17]
            m = abs(kk)
            assert m < 2**128 and neg == (1 if kk < 0 else 0) and ev == 1 - (m & 1) and h == (m + ev) >> 1


def _mul(sim, fn, c, k, P):
    xy, _ = pyref.enc_point(P)
    out = ctypes.create_string_buffer(64)
    inf = ctypes.create_string_buffer(1)
    getattr(sim, fn)(k.to_bytes(32, "big"), xy, out, inf)
    return pyref.dec_point(out.raw, inf.raw[0])


@pytest.mark.parametrize("curve,fn", [("k256", "sim_k256_mul"), ("p256", "sim_p256_mul"), ("k256", "sim_k256_mul_generic"),
                                      ("p256", "sim_p256_mul_3m5s"), ("k256", "sim_k256_mul_ptcalls")])
def test_scalar_mul_thread_routine(sim, curve, fn):
    c = pyref.CURVES[curve]
    G = pyref.G(c)
    rng = random.Random(13)
    g = golden(curve)["group"]
    for v in g["mul"]:
        assert _mul(sim, fn, c, int(v["k"], 16), G) == (int(v["x"], 16), int(v["y"], 16))
    for v in g["add"]:
        assert _mul(sim, fn, c, v["k"], G) == (int(v["x"], 16), int(v["y"], 16))
    Ps = random_points(c, 6, seed=3)
    for i, k in enumerate(edge_scalars(c) + [rng.randrange(c.n) for _ in range(25)]):
        P = Ps[i % 6]
        assert _mul(sim, fn, c, k, P) == pyref.mul(c, k, P), hex(k)


def test_msm_digit_recoding(sim):
    """bucket-method recoding: digits reconstruct the scalar, non-top digits in (-2^(c-1), 2^(c-1)], top fits its slots"""
    import ctypes as ct

    rng = random.Random(44)
    for nbits in (128, 256):
        for c in range(8, 17):
            W = (nbits + c - 1) // c
            for trial in range(60):
                m = rng.getrandbits(nbits) if trial > 6 else [0, 1, 2**nbits - 1, 2**(nbits - 1), 2**nbits, (1 << (c - 1)), (1 << c) - 1][trial]
                out = (ct.c_int32 * 40)()
                assert sim.sim_msm_recode(m.to_bytes(36, "little"), c, nbits, out) == W
                d = [out[i] for i in range(W)]
                assert sum(x << (c * i) for i, x in enumerate(d)) == m
                assert all(-(1 << (c - 1)) < x <= (1 << (c - 1)) for x in d[:-1])
                assert 0 <= d[-1] <= (1 << (c + 1)) + 1


def test_verify_front_end_pieces(sim):
    """SHA-256 challenge, lift_x and mod-n Montgomery arithmetic used by the signature-verification kernels"""
    rng = random.Random(9)
    for _ in range(50):
        r, pk, m = rng.randbytes(32), rng.randbytes(32), rng.randbytes(32)
        out = ctypes.create_string_buffer(32)
        sim.sim_bip340_challenge(r, pk, m, out)
        assert out.raw == pyref.tagged_hash(b"BIP0340/challenge", r + pk + m)
    hits = 0
    for i in range(60):
        x = rng.randrange(pyref.K256.p) if i else pyref.K256.gx
        out = ctypes.create_string_buffer(32)
        ok = sim.sim_k256_lift_x(x.to_bytes(32, "big"), out)
        exp = pyref.lift_x(x)
        assert bool(ok) == (exp is not None)
        if exp:
            hits += 1
            assert int.from_bytes(out.raw, "big") == exp[1]
    assert hits > 10
    assert sim.sim_k256_lift_x((pyref.K256.p + 1).to_bytes(32, "big"), ctypes.create_string_buffer(32)) == 0
    for ci, curve in enumerate(("k256", "p256")):
        n = pyref.CURVES[curve].n
        for a, b in [(1, 1), (n - 1, n - 1), (2, (n + 1) // 2), (n - 1, 2)] + [(rng.randrange(1, n), rng.randrange(1, n)) for _ in range(40)]:
            out = ctypes.create_string_buffer(32)
            sim.sim_fn_op(ci, 0, a.to_bytes(32, "big"), b.to_bytes(32, "big"), out)
            assert int.from_bytes(out.raw, "big") == a * b % n
        for a in [1, 2, n - 1] + [rng.randrange(1, n) for _ in range(5)]:
            out = ctypes.create_string_buffer(32)
            sim.sim_fn_op(ci, 1, a.to_bytes(32, "big"), (1).to_bytes(32, "big"), out)
            assert int.from_bytes(out.raw, "big") == pow(a, -1, n)


def test_sec1_decompress(sim):
    rng = random.Random(33)
    for ci, curve in enumerate(("k256", "p256")):
        c = pyref.CURVES[curve]
        p = c.p
        hits = 0
        for i in range(40):
            x = c.gx if i == 0 else rng.randrange(p)
            rhs = (pow(x, 3, p) + c.a * x + c.b) % p
            y = pow(rhs, (p + 1) // 4, p)
            exists = y * y % p == rhs
            for odd in (0, 1):
                out = ctypes.create_string_buffer(32)
                ok = sim.sim_sec1_decompress(ci, x.to_bytes(32, "big"), odd, out)
                assert bool(ok) == exists
                if exists:
                    hits += 1
                    yy = int.from_bytes(out.raw, "big")
                    assert yy % 2 == odd and yy in (y, p - y) and pyref.on_curve(c, (x, yy))
        assert hits > 10
        assert sim.sim_sec1_decompress(ci, p.to_bytes(32, "big"), 0, ctypes.create_string_buffer(32)) == 0


def test_on_curve_check(sim):
    for curve in ("k256", "p256"):
        c = pyref.CURVES[curve]
        P = pyref.mul(c, 12345, pyref.G(c))
        xy, _ = pyref.enc_point(P)
        f = getattr(sim, f"sim_{curve}_on_curve")
        assert f(xy) == 1
        bad = bytearray(xy)
        bad[40] ^= 4
        assert f(bytes(bad)) == 0


# ---- P-384: the same templates with a 12-limb field policy ----
def _feop384(sim, op, a, b=0):
    out = ctypes.create_string_buffer(48)
    sim.sim_p384_fe_op(op, a.to_bytes(48, "big"), b.to_bytes(48, "big"), out)
    return int.from_bytes(out.raw, "big")


def test_p384_field_ops_full_384bit_range(sim):
    p = pyref.P384.p
    rng = random.Random(384)
    edge = [0, 1, 2, p - 1, p, p + 1, 2**384 - 1, 2**384 - 2, 2**384 - p, 2**384 - p - 1, p - 2, (p + 1) // 2, 2**383, 2**128, 2**96,
            2**32, 2**128 + 2**96 - 2**32 + 1, 2**352, (2**384 - 1) ^ (0xFFFFFFFF << 128)]

    def structured():
        v = 0
        for i in range(12):
            c = rng.random()
            w = 0 if c < 0.3 else 0xFFFFFFFF if c < 0.6 else 1 if c < 0.65 else 0xFFFFFFFE if c < 0.7 else rng.getrandbits(32)
            v |= w << (32 * i)
        return v

    vals = edge + [rng.randrange(2**384) for _ in range(120)] + [structured() for _ in range(400)]
    for a in vals:
        for b in rng.sample(vals, 6) + edge[:6]:
            assert _feop384(sim, 0, a, b) == (a + b) % p
            assert _feop384(sim, 1, a, b) == (a - b) % p
            assert _feop384(sim, 2, a, b) == a * b % p
        assert _feop384(sim, 3, a) == a * a % p
        assert _feop384(sim, 4, a) == (-a) % p
        assert _feop384(sim, 5, a) == a * pow(2, -1, p) % p
        assert _feop384(sim, 6, a) == 3 * a % p
        assert _feop384(sim, 9, a) == 8 * a % p
        assert _feop384(sim, 8, a) == a % p
    for a in vals[:25]:
        assert _feop384(sim, 7, a) == (pow(a % p, -1, p) if a % p else 0)


def test_p384_mul_golden_and_edges(sim):
    """p384/src/test_vectors/group.rs:8,175 (k*G for k = 1..20, 32 multiplication vectors) + edge scalars, identity results"""
    c = pyref.P384
    g = golden("p384")
    G = pyref.G(c)

    def mul(k, P):
        xy, _ = pyref.enc_point(P, 48)
        out = ctypes.create_string_buffer(96)
        inf = ctypes.create_string_buffer(1)
        sim.sim_p384_mul(k.to_bytes(48, "big"), xy, out, inf)
        return pyref.dec_point(out.raw, inf.raw[0], 48)

    for v in g["group"]["add"]:
        assert mul(v["k"], G) == (int(v["x"], 16), int(v["y"], 16))
    for v in g["group"]["mul"]:
        assert mul(int(v["k"], 16), G) == (int(v["x"], 16), int(v["y"], 16))
    rng = random.Random(5)
    P = pyref.mul(c, rng.randrange(1, c.n), G)
    assert sim.sim_p384_on_curve(pyref.enc_point(P, 48)[0]) == 1
    assert sim.sim_p384_on_curve(pyref.enc_point((P[0], (P[1] + 1) % c.p), 48)[0]) == 0
    for k in [0, 1, 2, 3, 15, 16, 17, 2**128, 2**383, c.n - 1, c.n - 2, (c.n - 1) // 2] + [rng.randrange(c.n) for _ in range(12)]:
        assert mul(k, P) == pyref.mul(c, k, P), hex(k)
