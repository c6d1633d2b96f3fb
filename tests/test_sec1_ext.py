"""SEC1 decompression and the field square root for the curves beyond secp256k1 / P-256 (SURVEY.md section 8(f) rank 2,
widened): AffinePoint::decompress (primeorder/src/affine.rs:179-198) needs sqrt(x^3 + a x + b); for every curve of the
reference with p = 3 (mod 4) — all but P-224 — that is one exponentiation by (p + 1) / 4 through the same field policy; P-224
(p - 1 = 2^96 (2^128 - 1)) goes through Tonelli-Shanks, and since decompress picks the root by parity the answer is unique.
bign-curve256v1 records carry x in its little-endian FieldBytes (from_repr inside decompress).
CPU: the kernels on the host; GPU: through the C ABI.  Expected values: the big-integer model."""
import ctypes
import json
import os
import random

import numpy as np
import pytest

import pyref
from test_curves_ext import recs

HERE = os.path.dirname(os.path.abspath(__file__))
CURVES = ["p384", "sm2", "bp256r1", "bp256t1", "bignp256", "bp384r1", "bp384t1", "p224", "p192", "p521"]
# p384/tests/affine.rs:21-24 (COMPRESSED_BASEPOINT); the identity is FB + 1 zero bytes (:71-77)
_X = json.load(open(os.path.join(HERE, "golden", "sig_extras.json")))     # <- oracle/extract_golden.py
P384_COMPRESSED_BASEPOINT = _X["p384_compressed_basepoint"]
assert _X["p384_uncompressed_basepoint"] == "04" + "%096x%096x" % pyref.G(pyref.CURVES["p384"])
SQRT_CURVES = [c for c in CURVES if c != "p224"]


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def cases(c, count, seed):
    """compressed records (tag || x big-endian) with the expected point / validity"""
    nb = pyref.fbytes(c)
    rng = random.Random(seed)
    recs_, want = [], []
    for i in range(count):
        P = pyref.mul(c, rng.randrange(1, c.n), pyref.G(c))
        recs_.append(bytes([2 + (P[1] & 1)]) + pyref.enc_fe(c, P[0]))
        want.append((P, True))
        recs_.append(bytes([3 - (P[1] & 1)]) + pyref.enc_fe(c, P[0]))          # the other root
        want.append(((P[0], c.p - P[1]), True))
    Gp = pyref.G(c)                                                                # p384/tests/affine.rs:21-45: 03 || G.x -> the generator
    recs_.append(bytes([2 + (Gp[1] & 1)]) + pyref.enc_fe(c, Gp[0]))
    want.append((Gp, True))
    if c.name == "p384":
        assert recs_[-1].hex() == P384_COMPRESSED_BASEPOINT
    x = 1
    while pow((x**3 + c.a * x + c.b) % c.p, (c.p - 1) // 2, c.p) == 1:           # an x that is not on the curve
        x += 1
    recs_ += [bytes([2]) + pyref.enc_fe(c, x), bytes([2]) + pyref.enc_fe(c, c.p), bytes([4]) + pyref.enc_fe(c, 5), bytes(nb + 1)]
    want += [(None, False), (None, False), (None, False), (None, True)]           # no root, x >= p, bad tag, the identity
    return recs_, want


def check(c, want, xy, inf, valid):
    nb = pyref.fbytes(c)
    xy = np.asarray(xy).reshape(-1, 2 * nb)
    for i, (P, ok) in enumerate(want):
        assert bool(valid[i]) == ok, i
        if ok and P is not None:
            assert not inf[i] and (pyref.dec_fe(c, xy[i, :nb].tobytes()), pyref.dec_fe(c, xy[i, nb:].tobytes())) == P
        elif ok:
            assert inf[i] and not xy[i].any()
        else:
            assert not xy[i].any()


def sqrt_cases(c, seed):
    rng = random.Random(seed)
    a = [0, 1, 4] + [rng.randrange(c.p) for _ in range(40)]
    want = []
    for v in a:
        r = pow(v, (c.p + 1) // 4, c.p)
        want.append(r if r * r % c.p == v else None)
    return a, want


@pytest.mark.parametrize("name", ["p384", "sm2", "bp256r1", "bignp256", "p224", "p192", "p521"])
def test_kernels_on_host(name):
    import __graft_entry__ as ge
    ge.build()
    sim = ctypes.CDLL(os.path.join(HERE, "sim", "libecgsim.so"))
    c = pyref.CURVES[name]
    cid = pyref.CURVE_IDS[name]
    nb = pyref.fbytes(c)
    recs_, want = cases(c, 6, 3)
    n = len(recs_)
    buf = np.frombuffer(b"".join(recs_), np.uint8).copy()
    oxy, oinf, valid = np.full(2 * nb * n, 9, np.uint8), np.full(n, 9, np.uint8), np.full(n, 9, np.uint8)
    sim.simk_decompress_generic(cid, ctypes.c_size_t(n), _p(buf), _p(oxy), _p(oinf), _p(valid))
    check(c, want, oxy, oinf, valid)
    if name == "p224":
        return
    a, wroot = sqrt_cases(c, 4)
    A = recs(c, a)
    out, ok, st = np.zeros(nb * len(a), np.uint8), np.zeros(len(a), np.uint8), np.zeros(2, np.uint32)
    sim.simk_field_sqrt_generic(cid, ctypes.c_size_t(len(a)), _p(A), _p(out), _p(ok), _p(st))
    got = [pyref.dec_fe(c, out[nb * i:nb * i + nb].tobytes()) if ok[i] else None for i in range(len(a))]
    assert got == wroot and st[0] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", CURVES)
def test_gpu_decompress_and_sqrt(engine, name):
    import ecgpu

    c = pyref.CURVES[name]
    nb = pyref.fbytes(c)
    recs_, want = cases(c, 150, 7)
    xy, inf, valid = engine.decompress_batch(name, np.frombuffer(b"".join(recs_), np.uint8))
    check(c, want, xy, inf, valid)
    if name == "p224":                         # which root FieldElement::sqrt returns there is the external bignum crate's choice
        with pytest.raises(ecgpu.EcgError):
            engine.field_sqrt("p224", np.zeros(28, np.uint8))
        return
    a, wroot = sqrt_cases(c, 8)
    out, ok = engine.field_sqrt(name, recs(c, a))
    assert [pyref.dec_fe(c, out[i].tobytes()) if ok[i] else None for i in range(len(a))] == wroot
