"""Big-integer ground truth for the k256 / p256 scalar-multiplication hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (libecgpu.so, the
`ecgpu` package) may import this module; only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg do.

This is the *mathematical* definition the reference's results must satisfy
(affine chord-and-tangent arithmetic over exact Python integers).  The
algorithm-faithful restatement of the reference's limb code lives in
oracle/ecref.c; both are pinned against the reference's own golden vectors
(tests/golden/*.json, extracted by oracle/extract_golden.py from
k256/src/test_vectors/group.rs:9,96, p256/src/test_vectors/group.rs:8,95,
*/src/test_vectors/field.rs:6).

Curve constants as they appear in the reference:
  k256  p  k256/src/arithmetic/field.rs:42       n  k256/src/lib.rs:71
        G  k256/src/arithmetic/affine.rs:61-77   b=7  k256/src/arithmetic.rs:32-41
        beta  k256/src/arithmetic/projective.rs:32-37   lambda  k256/src/arithmetic/mul.rs:4-5
  p256  p  p256/src/arithmetic/field.rs:35       n  p256/src/lib.rs:60
        a=-3, b, G  p256/src/arithmetic.rs:53-74
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass


@dataclass(frozen=True)
class Curve:
    name: str
    p: int
    n: int
    a: int
    b: int
    gx: int
    gy: int


K256 = Curve(
    "k256",
    p=2**256 - 2**32 - 977,
    n=0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141,
    a=0,
    b=7,
    gx=0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798,
    gy=0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8,
)
P256 = Curve(
    "p256",
    p=0xFFFFFFFF00000001000000000000000000000000FFFFFFFFFFFFFFFFFFFFFFFF,
    n=0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551,
    a=-3,
    b=0x5AC635D8AA3A93E7B3EBBD55769886BC651D06B0CC53B0F63BCE3C3E27D2604B,
    gx=0x6B17D1F2E12C4247F8BCE6E563A440F277037D812DEB33A0F4A13945D898C296,
    gy=0x4FE342E2FE1A7F9B8EE7EB4A7C0F9E162BCE33576B315ECECBB6406837BF51F5,
)
# next curve through the same templates: p384/src/arithmetic.rs:53-74, p384/src/lib.rs:73,
# p384/src/arithmetic/field.rs:35
P384 = Curve(
    "p384",
    p=2**384 - 2**128 - 2**96 + 2**32 - 1,
    n=0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFC7634D81F4372DDF581A0DB248B0A77AECEC196ACCC52973,
    a=-3,
    b=0xB3312FA7E23EE7E4988E056BE3F82D19181D9C6EFE8141120314088F5013875AC656398D8A2ED19D2A85C8EDD3EC2AEF,
    gx=0xAA87CA22BE8B05378EB1C71EF320AD746E1D3B628BA79B9859F741E082542A385502F25DBF55296C3A545E3872760AB7,
    gy=0x3617DE4A96262C6F5D9E98BF9292DC29F8F41DBD289A147CE9DA3113B5F0B8C00A60B1CE1D7E819D7A431D7C90EA0E5F,
)
CURVES = {"k256": K256, "p256": P256, "p384": P384, 0: K256, 1: P256, 2: P384}


def fbytes(c: Curve) -> int:
    """bytes per field element / scalar in the canonical encoding (32, or 48 for P-384)"""
    return (c.p.bit_length() + 7) // 8

K256_BETA = 0x7AE96A2B657C07106E64479EAC3434E99CF0497512F58995C1396C28719501EE
K256_LAMBDA = 0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72

# A point is None (identity) or an (x, y) tuple of ints in [0, p).


def on_curve(c: Curve, P) -> bool:
    if P is None:
        return True
    x, y = P
    return 0 <= x < c.p and 0 <= y < c.p and (y * y - (x * x * x + c.a * x + c.b)) % c.p == 0


def neg(c: Curve, P):
    if P is None:
        return None
    return (P[0], (-P[1]) % c.p)


def add(c: Curve, P, Q):
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % c.p == 0:
            return None
        lam = (3 * x1 * x1 + c.a) * pow(2 * y1, -1, c.p) % c.p
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, c.p) % c.p
    x3 = (lam * lam - x1 - x2) % c.p
    y3 = (lam * (x1 - x3) - y1) % c.p
    return (x3, y3)


def mul(c: Curve, k: int, P):
    """k*P by plain double-and-add (k taken mod n)."""
    k %= c.n
    R = None
    Q = P
    while k:
        if k & 1:
            R = add(c, R, Q)
        Q = add(c, Q, Q)
        k >>= 1
    return R


def lincomb(c: Curve, ks, Ps):
    R = None
    for k, P in zip(ks, Ps):
        R = add(c, R, mul(c, k, P))
    return R


def G(c: Curve):
    return (c.gx, c.gy)


# ---- canonical encodings used at the C-ABI boundary ----

def enc_scalar(k: int, nb: int = 32) -> bytes:
    return k.to_bytes(nb, "big")


def enc_point(P, nb: int = 32) -> tuple[bytes, int]:
    """(x||y big-endian 2*nb bytes, inf flag). Identity = zero bytes + flag 1
    (AffinePoint::IDENTITY, k256/src/arithmetic/affine.rs:53-57).  nb = 32, or 48 for P-384."""
    if P is None:
        return bytes(2 * nb), 1
    return P[0].to_bytes(nb, "big") + P[1].to_bytes(nb, "big"), 0


def dec_point(xy: bytes, inf: int, nb: int = 32):
    if inf:
        return None
    return (int.from_bytes(xy[:nb], "big"), int.from_bytes(xy[nb:2 * nb], "big"))


# ---- deterministic synthetic inputs ----

def synth_scalar(c: Curve, seed: int, tag: bytes, i: int) -> int:
    h = hashlib.sha256(seed.to_bytes(8, "little") + tag + i.to_bytes(8, "little")).digest()
    return int.from_bytes(h, "big") % c.n


def glv_split(k: int):
    """Exact-integer GLV split used only to cross-check the device decomposition:
    returns (k1, k2) signed with k1 + k2*lambda == k (mod n)
    (constants: k256/src/arithmetic/mul.rs:7-35, mul/glv.rs:10-37,149-156)."""
    n = K256.n
    a1 = 0x3086D221A7D46BCDE86C90E49284EB15
    b1 = -0xE4437ED6010E88286F547FA90ABFE4C3
    a2 = 0x114CA50F7A8E2F3F657C1108D9D44CFD8
    b2 = a1
    g1 = 0x3086D221A7D46BCDE86C90E49284EB153DAA8A1471E8CA7FE893209A45DBB031
    g2 = 0xE4437ED6010E88286F547FA90ABFE4C4221208AC9DF506C61571B4AE8AC47F71
    c1 = (k * g1 + (1 << 383)) >> 384
    c2 = (k * g2 + (1 << 383)) >> 384
    k1 = k - c1 * a1 - c2 * a2
    k2 = -c1 * b1 - c2 * b2
    assert (k1 + k2 * K256_LAMBDA - k) % n == 0
    return k1, k2


# ---- signature verification (first widening step) ---------------------------------
# BIP340: k256/src/schnorr/verifying.rs:76-99 (verify_raw), :36-52 (from_bytes / lift_x), k256/src/schnorr.rs:221-227
# (tagged_hash).  ECDSA: SEC1 4.1.4 as implemented by the `ecdsa` crate the curve crates re-export
# (k256/src/ecdsa.rs:93-121); low-S rule = EcdsaCurve::NORMALIZE_S (k256/src/ecdsa.rs:104-106).

def tagged_hash(tag: bytes, data: bytes) -> bytes:
    t = hashlib.sha256(tag).digest()
    return hashlib.sha256(t + t + data).digest()


def lift_x(x: int):
    p = K256.p
    if x >= p:
        return None
    rhs = (pow(x, 3, p) + 7) % p
    y = pow(rhs, (p + 1) // 4, p)
    if y * y % p != rhs:
        return None
    return (x, y if y % 2 == 0 else p - y)


def bip340_verify(pk32: bytes, msg: bytes, sig64: bytes) -> bool:
    p, n = K256.p, K256.n
    P = lift_x(int.from_bytes(pk32, "big"))
    r = int.from_bytes(sig64[:32], "big")
    s = int.from_bytes(sig64[32:], "big")
    if P is None or r >= p or s >= n or s == 0:   # Signature::try_from: s is a NonZeroScalar
        return False
    e = int.from_bytes(tagged_hash(b"BIP0340/challenge", sig64[:32] + pk32 + msg), "big") % n
    R = add(K256, mul(K256, s, G(K256)), mul(K256, (n - e) % n, P))
    return R is not None and R[1] % 2 == 0 and R[0] == r


def bip340_sign(sk: int, msg: bytes, aux: bytes) -> tuple[bytes, bytes]:
    """BIP340 default signing (test-data generator only). Returns (pk32, sig64)."""
    n = K256.n
    P = mul(K256, sk, G(K256))
    d = sk if P[1] % 2 == 0 else n - sk
    pk = P[0].to_bytes(32, "big")
    t = (d ^ int.from_bytes(tagged_hash(b"BIP0340/aux", aux), "big")).to_bytes(32, "big")
    k0 = int.from_bytes(tagged_hash(b"BIP0340/nonce", t + pk + msg), "big") % n
    R = mul(K256, k0, G(K256))
    k = k0 if R[1] % 2 == 0 else n - k0
    e = int.from_bytes(tagged_hash(b"BIP0340/challenge", R[0].to_bytes(32, "big") + pk + msg), "big") % n
    return pk, R[0].to_bytes(32, "big") + ((k + e * d) % n).to_bytes(32, "big")


def ecdsa_verify(c: Curve, z: int, r: int, s: int, Q, low_s_only: bool = False) -> bool:
    n = c.n
    if not (0 < r < n and 0 < s < n) or Q is None or not on_curve(c, Q):
        return False
    if low_s_only and s > n // 2:
        return False
    w = pow(s, -1, n)
    R = add(c, mul(c, (z % n) * w % n, G(c)), mul(c, r * w % n, Q))
    return R is not None and R[0] % n == r


def ecdsa_sign(c: Curve, d: int, z: int, k: int):
    n = c.n
    R = mul(c, k, G(c))
    r = R[0] % n
    s = pow(k, -1, n) * (z + r * d) % n
    return r, s
