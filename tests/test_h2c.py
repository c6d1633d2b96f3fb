"""Hash to curve (RFC 9380) — the four Weierstrass suites of the reference:
secp256k1_XMD:SHA-256_SSWU_, P256_XMD:SHA-256_SSWU_, P384_XMD:SHA-384_SSWU_, P521_XMD:SHA-512_SSWU_ (RO and NU), plus hash_to_scalar.

CPU: the big-integer model (oracle/pyref.py) is pinned to the vectors the reference's own tests hold
(k256/src/arithmetic/hash2curve.rs:289-370, p256/src/arithmetic/hash2curve.rs:133-310, p384 / p521 likewise -> tests/golden/h2c.json, every
intermediate: u_0, u_1, Q_0, Q_1, P), and the device code (ecg_h2c.cuh) runs on the host against it.
GPU (-m gpu): the same through the C ABI."""
import ctypes
import hashlib
import json
import os
import random

import numpy as np
import pytest

import pyref

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "h2c.json")))
CID = {"k256": 0, "p256": 1, "p384": 2, "p521": 11}
SUITES = ["k256", "p256", "p384", "p521"]
HASH = {"k256": "sha256", "p256": "sha256", "p384": "sha384", "p521": "sha512"}
# lengths around the SHA-2 block boundaries (64 / 128 bytes) of b_0 = H(Z_pad || msg || len || 0 || DST') and long messages
LENGTHS = (0, 1, 3, 31, 32, 33, 54, 55, 56, 57, 63, 64, 65, 110, 111, 112, 113, 119, 120, 121, 127, 128, 129, 200, 255, 256, 257, 1000)


def messages(seed):
    rng = random.Random(seed)
    return [bytes(rng.randrange(256) for _ in range(L)) for L in LENGTHS]


def vec_points(curve):
    return [(int(v["p_x"], 16), int(v["p_y"], 16)) for v in G["suites"][curve]["vectors"]]


@pytest.mark.parametrize("curve", SUITES)
def test_model_reproduces_the_reference_vectors(curve):
    su = G["suites"][curve]
    dst = su["dst"].encode()
    c = pyref.CURVES[curve]
    assert len(su["vectors"]) == 5
    for v in su["vectors"]:
        msg = v["msg"].encode()
        u0, u1 = pyref._h2c_field(curve, msg, dst, 2, c.p)
        assert (u0, u1) == (int(v["u_0"], 16), int(v["u_1"], 16))
        assert pyref.map_to_curve(curve, u0) == (int(v["q0_x"], 16), int(v["q0_y"], 16))
        assert pyref.map_to_curve(curve, u1) == (int(v["q1_x"], 16), int(v["q1_y"], 16))
        P = pyref.hash_to_curve(curve, msg, dst)
        assert P == (int(v["p_x"], 16), int(v["p_y"], 16)) and pyref.on_curve(c, P)


def test_model_hash_to_scalar_voprf_vectors():
    """p256/src/arithmetic/hash2curve.rs:257-310 (VOPRF DeriveKeyPair): first non-zero hash_to_scalar over the counter"""
    for v in G["p256_hash_to_scalar_voprf"]:
        dst, ki, seed = bytes.fromhex(v["dst_hex"]), v["key_info"].encode(), bytes.fromhex(v["seed"])
        s = 0
        for ctr in range(256):
            s = pyref.hash_to_scalar("p256", seed + len(ki).to_bytes(2, "big") + ki + bytes([ctr]), dst)
            if s:
                break
        assert s == int(v["sk_sm"], 16)


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def _pack(msgs):
    offs = np.zeros(len(msgs) + 1, np.uint64)
    offs[1:] = np.cumsum([len(m) for m in msgs])
    return np.frombuffer(b"".join(msgs) + b"\0", np.uint8).copy(), offs


def _dst_prime(dst, curve="k256"):
    if len(dst) > 255:
        dst = hashlib.new(HASH[curve], b"H2C-OVERSIZE-DST-" + dst).digest()
    return np.frombuffer(dst + bytes([len(dst)]), np.uint8).copy()


def _dec(curve, xy, inf):
    nb = pyref.fbytes(pyref.CURVES[curve])
    return pyref.dec_point(xy.tobytes(), int(inf), nb)


@pytest.fixture(scope="module")
def sim():
    import __graft_entry__ as ge
    ge.build()
    return ctypes.CDLL(os.path.join(HERE, "sim", "libecgsim.so"))


def sim_h2c(sim, curve, msgs, dst, nu):
    n = len(msgs)
    data, offs = _pack(msgs)
    dp = _dst_prime(dst, curve)
    nb = pyref.fbytes(pyref.CURVES[curve])
    oxy, oinf = np.zeros(2 * nb * n, np.uint8), np.zeros(n, np.uint8)
    sim.simk_hash_to_curve(CID[curve], ctypes.c_size_t(n), _p(data), _p(offs), _p(dp), len(dp), nu, _p(oxy), _p(oinf))
    oxy = oxy.reshape(n, 2 * nb)
    return [_dec(curve, oxy[i], oinf[i]) for i in range(n)]


@pytest.mark.parametrize("curve", SUITES)
def test_kernels_on_host(sim, curve):
    su = G["suites"][curve]
    dst = su["dst"].encode()
    assert sim_h2c(sim, curve, [v["msg"].encode() for v in su["vectors"]], dst, 0) == vec_points(curve)
    msgs = messages(7) if curve in ("k256", "p256") else messages(7)[::3]   # the 12- and 17-limb fields are slow on the host
    for d in ((dst, b"x", b"Y" * 255, b"Z" * 256, b"W" * 300) if curve in ("k256", "p256") else (dst, b"Z" * 256)):
        assert sim_h2c(sim, curve, msgs, d, 0) == [pyref.hash_to_curve(curve, m, d) for m in msgs]
        assert sim_h2c(sim, curve, msgs, d, 1) == [pyref.encode_to_curve(curve, m, d) for m in msgs]
    n = len(msgs)
    nb = pyref.fbytes(pyref.CURVES[curve])
    data, offs = _pack(msgs)
    dp = _dst_prime(dst, curve)
    out = np.zeros(nb * n, np.uint8)
    sim.simk_hash_to_scalar(CID[curve], ctypes.c_size_t(n), _p(data), _p(offs), _p(dp), len(dp), _p(out))
    assert [int.from_bytes(out[nb * i:nb * i + nb].tobytes(), "big") for i in range(n)] == [pyref.hash_to_scalar(curve, m, dst) for m in msgs]


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("curve", SUITES)
def test_gpu_hash_to_curve(engine, curve):
    import ecgpu

    su = G["suites"][curve]
    dst = su["dst"].encode()
    xy, inf = engine.hash_to_curve(curve, [v["msg"].encode() for v in su["vectors"]], dst)
    assert [_dec(curve, xy[i], inf[i]) for i in range(5)] == vec_points(curve)
    rng = random.Random(11)
    msgs = messages(8) + [bytes(rng.randrange(256) for _ in range(rng.randrange(0, 300))) for _ in range(2000)]
    for d in (dst, b"x", b"Y" * 255, b"Z" * 256):
        xy, inf = engine.hash_to_curve(curve, msgs, d)
        nxy, ninf = engine.encode_to_curve(curve, msgs, d)
        assert not inf.any() and not ninf.any()
        for i in list(range(len(LENGTHS))) + [100, 1999, len(msgs) - 1]:
            assert _dec(curve, xy[i], 0) == pyref.hash_to_curve(curve, msgs[i], d)
            assert _dec(curve, nxy[i], 0) == pyref.encode_to_curve(curve, msgs[i], d)
        # every output is a point of the curve
        c = pyref.CURVES[curve]
        for i in range(0, len(msgs), 37):
            assert pyref.on_curve(c, _dec(curve, xy[i], 0)) and pyref.on_curve(c, _dec(curve, nxy[i], 0))
    sc = engine.hash_to_scalar(curve, msgs[:64], dst)
    assert [int.from_bytes(sc[i].tobytes(), "big") for i in range(64)] == [pyref.hash_to_scalar(curve, m, dst) for m in msgs[:64]]
    # empty batch, empty DST (ExpandMsgXmdError::EmptyDst), curves without a SHA-256 suite
    xy, inf = engine.hash_to_curve(curve, [], dst)
    assert xy.shape[0] == 0
    with pytest.raises(ecgpu.EcgError) as ei:
        engine.hash_to_curve(curve, [b"a"], b"")
    assert ei.value.code == ecgpu.ECG_EINVAL
    with pytest.raises(ecgpu.EcgError):
        engine.hash_to_curve("sm2", [b"a"], dst)      # the reference defines no suite for it


@pytest.mark.gpu
def test_gpu_voprf_derive_key_vectors(engine):
    for v in G["p256_hash_to_scalar_voprf"]:
        dst, ki, seed = bytes.fromhex(v["dst_hex"]), v["key_info"].encode(), bytes.fromhex(v["seed"])
        msgs = [seed + len(ki).to_bytes(2, "big") + ki + bytes([ctr]) for ctr in range(256)]
        sc = engine.hash_to_scalar("p256", msgs, dst)
        first = next(int.from_bytes(sc[i].tobytes(), "big") for i in range(256) if sc[i].any())
        assert first == int(v["sk_sm"], 16)
