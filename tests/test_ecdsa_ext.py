"""ECDSA verification and a*G + b*P for the curves beyond secp256k1 / P-256 (widened to every
curve with ECDSA in the reference: p192, p224, p384, p521, brainpoolP256r1/t1, brainpoolP384r1/t1 — */src/ecdsa.rs).

Pinned to the reference's own vectors: the FIPS 186-4 vectors of p192 / p224 / p384 / p521 (src/test_vectors/ecdsa.rs ->
tests/golden/*_ecdsa.json) and the Wycheproof files of p224 / p384 / p521 (src/test_vectors/data/wycheproof.blb ->
tests/golden/*_wycheproof.json), replayed the way the reference's harness does (DER parsing, the curve's digest,
bits2field); the brainpool curves against the big-integer model on signatures made here (with corruptions).
CPU: the kernels on the host for P-224 / P-192.  GPU (-m gpu): every curve through the C ABI."""
import ctypes
import json
import os
import random

import numpy as np
import pytest

import ecref
import pyref
from helpers import GOLDEN, wycheproof_cases
from test_curves_ext import ext_fb_table, pts, recs, unpack

HERE = os.path.dirname(os.path.abspath(__file__))
ECDSA_CURVES = ["p384", "bp256r1", "bp256t1", "bp384r1", "bp384t1", "p224", "p192", "p521"]


def _p(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)


def fips_cases(curve):
    """[(z bytes, r, s, Q, True)] from the reference's FIPS vectors: `m` is the prehash"""
    c = pyref.CURVES[curve]
    nb = pyref.fbytes(c)
    out = []
    for v in json.load(open(os.path.join(GOLDEN, f"{curve}_ecdsa.json")))["vectors"]:
        m = bytes.fromhex(v["m"])
        z = m[:nb] if len(m) >= nb else bytes(nb - len(m)) + m     # bits2field
        out.append((z, int(v["r"], 16), int(v["s"], 16), (int(v["q_x"], 16), int(v["q_y"], 16)), True))
        # the key pair is a k*G fixture too
        assert pyref.mul(c, int(v["d"], 16), pyref.G(c)) == out[-1][3]
    assert len(out) == 15
    return out


def made_cases(curve, count, seed):
    """signatures made with the model (valid), then one corruption each (invalid): (z, r, s, Q, expected)"""
    c = pyref.CURVES[curve]
    nb = pyref.fbytes(c)
    rng = random.Random(seed)
    out = []
    for i in range(count):
        d, k = rng.randrange(1, c.n), rng.randrange(1, c.n)
        z = rng.randrange(1 << (8 * nb))
        r, s = pyref.ecdsa_sign(c, d, z % c.n, k)
        if r == 0 or s == 0:
            continue
        Q = pyref.mul(c, d, pyref.G(c))
        zb = z.to_bytes(nb, "big")
        out.append((zb, r, s, Q, True))
        how = i % 4
        if how == 0:
            out.append((zb, r, (s + 1) % c.n or 1, Q, None))
        elif how == 1:
            out.append(((z ^ 1).to_bytes(nb, "big"), r, s, Q, None))
        elif how == 2:
            out.append((zb, r, s, pyref.mul(c, d + 1, pyref.G(c)), None))
        else:
            out.append((zb, 0, s, Q, False))                       # r = 0 is refused
    return [(z, r, s, Q, pyref.ecdsa_verify(c, int.from_bytes(z, "big"), r, s, Q) if e is None else e) for z, r, s, Q, e in out]


def pack_cases(c, cases):
    nb = pyref.fbytes(c)
    Z = np.frombuffer(b"".join(z for z, *_ in cases), np.uint8).copy()
    S = np.frombuffer(b"".join(r.to_bytes(nb, "big") + s.to_bytes(nb, "big") for _, r, s, _, _ in cases), np.uint8).copy()
    Q = np.frombuffer(b"".join(q[0].to_bytes(nb, "big") + q[1].to_bytes(nb, "big") for *_, q, _ in cases), np.uint8).copy()
    return Z, S, Q, [e for *_, e in cases]


@pytest.mark.parametrize("curve", ["p192", "p224", "p384", "p521"])
def test_model_on_reference_vectors(curve):
    c = pyref.CURVES[curve]
    for z, r, s, Q, exp in fips_cases(curve):
        assert pyref.ecdsa_verify(c, int.from_bytes(z, "big"), r, s, Q) == exp
    if curve != "p192":
        cases, rejected = wycheproof_cases(curve)
        assert len(cases) > 100 and all(not v["pass"] for v in rejected)
        for z, r, s, Q, exp in cases[::3 if curve == "p521" else 2]:      # big-integer model: a sample here, every record on the GPU
            assert pyref.ecdsa_verify(c, int.from_bytes(z, "big"), r, s, Q) == exp


@pytest.mark.parametrize("curve", ["p224", "p192"])
def test_kernels_on_host(curve):
    import __graft_entry__ as ge
    ge.build()
    sim = ctypes.CDLL(os.path.join(HERE, "sim", "libecgsim.so"))
    c = pyref.CURVES[curve]
    cid = pyref.CURVE_IDS[curve]
    nb = pyref.fbytes(c)
    table = ext_fb_table(sim, cid)
    cases = fips_cases(curve) + made_cases(curve, 12, 5)
    if curve == "p224":
        cases += wycheproof_cases(curve)[0][::3]
    Z, S, Q, exp = pack_cases(c, cases)
    n = len(cases)
    valid = np.full(n, 7, np.uint8)
    sim.simk_ecdsa_verify_generic(cid, ctypes.c_size_t(n), _p(Z), _p(S), _p(Q), 0, _p(table), _p(valid))
    assert [bool(v) for v in valid] == exp
    # a*G + b*P with the exceptional endings (b*P = -a*G, = a*G) and identities
    rng = random.Random(3)
    G = pyref.G(c)
    a_s = [rng.randrange(c.n) for _ in range(10)] + [5, 5, 0, 7]
    b_s = [rng.randrange(c.n) for _ in range(10)] + [c.n - 5, 5, 0, 0]
    Ps = [pyref.mul(c, rng.randrange(1, c.n), G) for _ in range(10)] + [G, G, G, None]
    pxy, pinf = pts(c, Ps)
    A, B = recs(c, a_s), recs(c, b_s)
    oxy, oinf, st = np.zeros(2 * nb * len(a_s), np.uint8), np.zeros(len(a_s), np.uint8), np.zeros(2, np.uint32)
    sim.simk_mul_gen_add_generic(cid, ctypes.c_size_t(len(a_s)), _p(A), _p(B), _p(pxy), _p(pinf), _p(table), _p(oxy), _p(oinf), _p(st))
    assert st[0] == 0
    want = [pyref.add(c, pyref.mul(c, a, G), pyref.mul(c, b, P) if P is not None else None) for a, b, P in zip(a_s, b_s, Ps)]
    assert unpack(c, oxy, oinf) == want


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("curve", ECDSA_CURVES)
def test_gpu_ecdsa_verify(engine, curve):
    import ecgpu

    c = pyref.CURVES[curve]
    cases = made_cases(curve, 40, 11)
    if curve in ("p192", "p224", "p384", "p521"):
        cases += fips_cases(curve)
    if curve in ("p224", "p384", "p521"):
        cases += wycheproof_cases(curve)[0]
    Z, S, Q, exp = pack_cases(c, cases)
    valid = engine.ecdsa_verify_batch(curve, Z, S, Q)
    wrong = [i for i, (v, e) in enumerate(zip(valid, exp)) if bool(v) != e]
    assert not wrong, f"{curve}: verdict differs from the reference's expectation at {wrong[:8]}"
    assert sum(exp) >= 40
    # an off-curve public key / out-of-range values are per-signature failures, not API errors
    nb = pyref.fbytes(c)
    z, r, s, Qp, _ = cases[0]
    bad_q = (Qp[0], (Qp[1] + 1) % c.p)
    Zb, Sb, Qb, _ = pack_cases(c, [(z, r, s, bad_q, False), (z, c.n, s, Qp, False), (z, r, c.n, Qp, False), (z, r, s, Qp, True)])
    assert list(engine.ecdsa_verify_batch(curve, Zb, Sb, Qb)) == [0, 0, 0, 1]
    for name in ("sm2", "bignp256"):       # SM2DSA / bign signatures are other schemes: loud refusal
        with pytest.raises(ecgpu.EcgError):
            engine.ecdsa_verify_batch(name, np.zeros(32, np.uint8), np.zeros(64, np.uint8), np.zeros(64, np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("curve", ECDSA_CURVES + ["sm2", "bignp256"])
def test_gpu_mul_gen_add(engine, curve):
    c = pyref.CURVES[curve]
    rng = random.Random(77)
    G = pyref.G(c)
    n = 300
    a_s = [rng.randrange(c.n) for _ in range(n)]
    b_s = [rng.randrange(c.n) for _ in range(n)]
    base = [pyref.mul(c, rng.randrange(1, c.n), G) for _ in range(8)]
    Ps = [base[i % 8] for i in range(n)]
    a_s[:4], b_s[:4], Ps[:4] = [5, 5, 0, 7], [c.n - 5, 5, 0, 0], [G, G, G, None]
    pxy, pinf = pts(c, Ps)
    xy, inf = engine.mul_by_generator_and_mul_add(curve, recs(c, a_s), recs(c, b_s), pxy, pinf)
    got = unpack(c, xy, inf)
    # a*G + b*P through two calls of the C restatement and one model addition per element
    g_xy, g_inf = ecref.mul_gen_batch(curve, recs(c, a_s), nthreads=8)
    p_xy, p_inf = ecref.mul_batch(curve, recs(c, b_s), pxy, pinf, nthreads=8)
    ga, pb = unpack(c, g_xy, g_inf), unpack(c, p_xy, p_inf)
    assert got == [pyref.add(c, x, y) for x, y in zip(ga, pb)]
