"""CPU-only: pin the oracle (C restatement + big-int model) against the reference's golden vectors."""
import random

import numpy as np
import pytest

import ecref
import pyref
from helpers import edge_scalars, golden, pack_points, pack_scalars, random_points, unpack_points

CURVES = ["k256", "p256"]


@pytest.mark.parametrize("curve", CURVES)
def test_pyref_reproduces_golden_group_vectors(curve):
    c = pyref.CURVES[curve]
    g = golden(curve)["group"]
    G = pyref.G(c)
    acc = None
    for v in g["add"]:  # ADD_TEST_VECTORS: repeated addition of the generator
        acc = pyref.add(c, acc, G)
        assert acc == (int(v["x"], 16), int(v["y"], 16))
    for v in g["mul"]:  # MUL_TEST_VECTORS
        assert pyref.mul(c, int(v["k"], 16), G) == (int(v["x"], 16), int(v["y"], 16))
    for v in golden(curve)["ecdsa"]["keypairs"]:
        assert pyref.mul(c, int(v["d"], 16), G) == (int(v["x"], 16), int(v["y"], 16))


@pytest.mark.parametrize("curve", CURVES)
def test_pyref_matches_openssl(curve):
    ec = pytest.importorskip("cryptography.hazmat.primitives.asymmetric.ec")
    c = pyref.CURVES[curve]
    oc = ec.SECP256K1() if curve == "k256" else ec.SECP256R1()
    rng = random.Random(5)
    for _ in range(8):
        k = rng.randrange(1, c.n)
        pub = ec.derive_private_key(k, oc).public_key().public_numbers()
        assert pyref.mul(c, k, pyref.G(c)) == (pub.x, pub.y)


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("variant", [0, 1])
def test_ecref_golden_mul_vectors(curve, variant):
    c = pyref.CURVES[curve]
    g = golden(curve)["group"]
    ks = [v["k"] for v in g["add"]] + [int(v["k"], 16) for v in g["mul"]]
    exp = [(int(v["x"], 16), int(v["y"], 16)) for v in g["add"] + g["mul"]]
    xy, inf = pack_points([pyref.G(c)] * len(ks))
    oxy, oinf = ecref.mul_batch(curve, pack_scalars(ks), xy, inf, nthreads=2, variant=variant)
    assert unpack_points(oxy, oinf) == exp
    oxy, oinf = ecref.mul_gen_batch(curve, pack_scalars(ks), nthreads=2)
    assert unpack_points(oxy, oinf) == exp


@pytest.mark.parametrize("curve", CURVES)
def test_ecref_field_golden_and_bigint(curve):
    c = pyref.CURVES[curve]
    p = c.p
    dbl = [int(h, 16) for h in golden(curve)["field"]["dbl"]]
    A = pack_scalars(dbl[:-1])
    out = ecref.field_op(curve, 0, A, A)
    assert [int.from_bytes(o.tobytes(), "big") for o in out] == dbl[1:]
    rng = random.Random(1)
    a = [0, 1, p - 1, p - 2, 2**255 % p] + [rng.randrange(p) for _ in range(500)]
    b = [p - 1, 0, p - 1, 2, 3] + [rng.randrange(p) for _ in range(500)]
    A, B = pack_scalars(a), pack_scalars(b)
    model = {0: lambda x, y: (x + y) % p, 1: lambda x, y: (x - y) % p, 2: lambda x, y: (-x) % p,
             3: lambda x, y: x * y % p, 4: lambda x, y: x * x % p, 5: lambda x, y: pow(x, -1, p) if x else 0}
    for op, f in model.items():
        out = ecref.field_op(curve, op, A, B if op in (0, 1, 3) else None)
        assert [int.from_bytes(o.tobytes(), "big") for o in out] == [f(x, y) for x, y in zip(a, b)], op


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("variant", [0, 1])
def test_ecref_varbase_vs_bigint(curve, variant):
    c = pyref.CURVES[curve]
    rng = random.Random(42 + variant)
    ks = edge_scalars(c) + [rng.randrange(c.n) for _ in range(40)]
    Ps = random_points(c, len(ks), seed=17)
    Ps[2] = None
    xy, inf = pack_points(Ps)
    oxy, oinf = ecref.mul_batch(curve, pack_scalars(ks), xy, inf, nthreads=3, variant=variant)
    assert unpack_points(oxy, oinf) == [pyref.mul(c, k, P) for k, P in zip(ks, Ps)]


@pytest.mark.parametrize("curve", CURVES)
def test_ecref_lincomb_vs_bigint(curve):
    c = pyref.CURVES[curve]
    rng = random.Random(8)
    for n in (1, 2, 5, 300):
        base = random_points(c, min(n, 16), seed=n)
        Ps = [base[i % len(base)] for i in range(n)]
        ks = [rng.randrange(c.n) for _ in range(n)]
        xy, inf = pack_points(Ps)
        oxy, oinf = ecref.lincomb(curve, pack_scalars(ks), xy, inf, nthreads=2)
        assert pyref.dec_point(oxy.tobytes(), oinf) == pyref.lincomb(c, ks, Ps)


@pytest.mark.parametrize("curve", CURVES)
def test_ecref_mul_gen_add_vs_bigint(curve):
    c = pyref.CURVES[curve]
    rng = random.Random(77)
    n = 40
    Ps = random_points(c, n, seed=5)
    a = [rng.randrange(c.n) for _ in range(n)]
    b = [rng.randrange(c.n) for _ in range(n)]
    a[0], b[0] = 0, 0
    a[1], b[1] = 3, 0
    a[2], b[2] = 0, 9
    Ps[3] = None
    Ps[4] = pyref.G(c)
    b[4] = c.n - a[4]
    xy, inf = pack_points(Ps)
    oxy, oinf = ecref.mul_gen_add_batch(curve, pack_scalars(a), pack_scalars(b), xy, inf, nthreads=2)
    G = pyref.G(c)
    assert unpack_points(oxy, oinf) == [pyref.add(c, pyref.mul(c, x, G), pyref.mul(c, y, P)) for x, y, P in zip(a, b, Ps)]


def test_radix16_properties():
    # primeorder/src/tables/radix16.rs:110-172: digits in [-8, 8], reconstruct the scalar
    rng = random.Random(3)
    for _ in range(200):
        k = rng.randrange(2**256)
        d = ecref.radix16(k.to_bytes(32, "big"), 65)
        assert all(-8 <= int(x) <= 8 for x in d)
        assert sum(int(x) * 16**i for i, x in enumerate(d)) == k


def test_wnaf_properties():
    # wnaf/src/lib.rs:70-150: odd digits |d| < 2^(w-1), no two non-zeros within w positions, exact value
    rng = random.Random(4)
    for w in (2, 3, 4, 5, 6, 8):
        for nbytes in (16, 32):
            for _ in range(60):
                k = rng.randrange(2 ** (8 * nbytes))
                d = [int(x) for x in ecref.wnaf(k.to_bytes(nbytes, "little"), 8 * nbytes, w)]
                assert sum(x * 2**i for i, x in enumerate(d)) == k
                nz = [i for i, x in enumerate(d) if x]
                assert all(d[i] % 2 != 0 and abs(d[i]) < 2 ** (w - 1) for i in nz)
                assert all(b - a >= w for a, b in zip(nz, nz[1:]))


def test_glv_decomposition_matches_reference_definition():
    # k256/src/arithmetic/mul/glv.rs:149-156 + bound proof :43-146
    rng = random.Random(6)
    n = pyref.K256.n
    for k in [0, 1, n - 1, pyref.K256_LAMBDA] + [rng.randrange(n) for _ in range(300)]:
        r1, r2 = ecref.glv(k)
        assert (r1 + r2 * pyref.K256_LAMBDA - k) % n == 0
        assert min(r1, n - r1) < 2**128 and min(r2, n - r2) < 2**128
        k1, k2 = pyref.glv_split(k)
        assert (k1 % n, k2 % n) == (r1, r2)


def test_oracle_signature_verification_vs_reference_vectors():
    """BIP340 vectors (k256/src/schnorr.rs) incl. the 10 invalid cases; FIPS 186-4 ECDSA vectors
    ({k256,p256}/src/test_vectors/ecdsa.rs): signing reproduces (r, s), verification accepts, tampering rejects."""
    import json
    import os

    from helpers import GOLDEN

    v = json.load(open(os.path.join(GOLDEN, "k256_bip340.json")))["vectors"]
    assert len(v) == 15 and sum(x["valid"] for x in v) == 5
    for x in v:
        assert pyref.bip340_verify(bytes.fromhex(x["pk"]), bytes.fromhex(x["msg"]), bytes.fromhex(x["sig"])) == x["valid"], x["index"]
        if "sk" in x:
            pk, sig = pyref.bip340_sign(int(x["sk"], 16), bytes.fromhex(x["msg"]), bytes.fromhex(x["aux"]))
            assert pk.hex() == x["pk"] and sig.hex() == x["sig"]
    for curve in CURVES:
        c = pyref.CURVES[curve]
        for x in json.load(open(os.path.join(GOLDEN, f"{curve}_ecdsa.json")))["vectors"]:
            d, k, z, r, s = (int(x[t], 16) for t in ("d", "k", "m", "r", "s"))
            Q = (int(x["q_x"], 16), int(x["q_y"], 16))
            assert pyref.ecdsa_sign(c, d, z, k) == (r, s)
            assert pyref.ecdsa_verify(c, z, r, s, Q) and not pyref.ecdsa_verify(c, z ^ 1, r, s, Q)


def test_p384_restatement_pinned_to_reference_vectors_and_model():
    """oracle/ecref_p384.c (the reference's generic primeorder path over a 384-bit Montgomery field) against the reference's
    own P-384 vectors (p384/src/test_vectors/group.rs:8,175) and the big-integer model: k*G, k*P, lincomb, identities,
    rejected inputs."""
    import random

    import numpy as np

    c = pyref.P384
    g = golden("p384")
    G = pyref.G(c)
    ks = list(range(1, 21)) + [int(v["k"], 16) for v in g["group"]["mul"]]
    want = [(int(v["x"], 16), int(v["y"], 16)) for v in g["group"]["add"]] + [(int(v["x"], 16), int(v["y"], 16)) for v in g["group"]["mul"]]
    K = np.frombuffer(b"".join(k.to_bytes(48, "big") for k in ks), np.uint8)
    xy, inf = ecref.mul_gen_batch("p384", K, nthreads=4)
    assert [pyref.dec_point(xy[i].tobytes(), int(inf[i]), 48) for i in range(len(ks))] == want
    Gb = np.frombuffer(pyref.enc_point(G, 48)[0] * len(ks), np.uint8)
    xy, inf = ecref.mul_batch("p384", K, Gb, None, nthreads=4)
    assert [pyref.dec_point(xy[i].tobytes(), int(inf[i]), 48) for i in range(len(ks))] == want
    rng = random.Random(384)
    n = 60
    ks = [0, 1, c.n - 1, 2**383] + [rng.randrange(c.n) for _ in range(n - 4)]
    Ps = [pyref.mul(c, rng.randrange(1, c.n), G) for _ in range(n)]
    Ps[9] = None
    K = np.frombuffer(b"".join(k.to_bytes(48, "big") for k in ks), np.uint8)
    P = np.frombuffer(b"".join(pyref.enc_point(p, 48)[0] for p in Ps), np.uint8)
    I = np.array([1 if p is None else 0 for p in Ps], np.uint8)
    xy, inf = ecref.mul_batch("p384", K, P, I, nthreads=3)
    assert [pyref.dec_point(xy[i].tobytes(), int(inf[i]), 48) for i in range(n)] == [pyref.mul(c, k, p) if p is not None else None for k, p in zip(ks, Ps)]
    lxy, linf = ecref.lincomb("p384", K, P, I, nthreads=3)
    acc = None
    for k, p in zip(ks, Ps):
        if p is not None:
            acc = pyref.add(c, acc, pyref.mul(c, k, p))
    assert pyref.dec_point(lxy.tobytes(), linf, 48) == acc
    bad = K.copy()
    bad[48 * 5:48 * 6] = np.frombuffer(c.n.to_bytes(48, "big"), np.uint8)
    with pytest.raises(ValueError):
        ecref.mul_batch("p384", bad, P, I)
    off = P.copy()
    off[96 * 3 + 95] ^= 1
    with pytest.raises(ValueError):
        ecref.mul_batch("p384", K, off, I)
