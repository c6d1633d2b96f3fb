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
    le: bool = False  # canonical records little-endian (bign-curve256v1 only)


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

# ---- the remaining prime-order curves of the reference, ids as in include/ecgpu.h ------------
# sm2/src/arithmetic.rs:44-67, sm2/src/arithmetic/field.rs:34, sm2/src/lib.rs:86
SM2 = Curve(
    "sm2",
    p=0xFFFFFFFEFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFF00000000FFFFFFFFFFFFFFFF,
    n=0xFFFFFFFEFFFFFFFFFFFFFFFFFFFFFFFF7203DF6B21C6052B53BBF40939D54123,
    a=-3,
    b=0x28E9FA9E9D9F5E344D5A9E4BCF6509A7F39789F515AB8F92DDBCBD414D940E93,
    gx=0x32C4AE2C1F1981195F9904466A39C9948FE30BBFF2660BE1715A4589334C74C7,
    gy=0xBC3736A2F4F6779C59BDCEE36B692153D0A9877CC62A474002DF32E52139F0A0,
)
_BP256_P = 0xA9FB57DBA1EEA9BC3E660A909D838D726E3BF623D52620282013481D1F6E5377
_BP256_N = 0xA9FB57DBA1EEA9BC3E660A909D838D718C397AA3B561A6F7901E0E82974856A7
# bp256/src/r1/arithmetic.rs:34-53 (general a), bp256/src/arithmetic/field.rs:53, bp256/src/lib.rs:70
BP256R1 = Curve(
    "bp256r1", p=_BP256_P, n=_BP256_N,
    a=0x7D5A0975FC2C3057EEF67530417AFFE7FB8055C126DC5C6CE94A4B44F330B5D9,
    b=0x26DC5C6CE94A4B44F330B5D9BBD77CBF958416295CF7E1CE6BCCDC18FF8C07B6,
    gx=0x8BD2AEB9CB7E57CB2C4B482FFC81B7AFB9DE27E1E3BD23C23A4453BD9ACE3262,
    gy=0x547EF835C3DAC4FD97F8461A14611DC9C27745132DED8E545C1D54C72F046997,
)
# bp256/src/t1/arithmetic.rs:34-51
BP256T1 = Curve(
    "bp256t1", p=_BP256_P, n=_BP256_N, a=-3,
    b=0x662C61C430D84EA4FE66A7733D0B76B7BF93EBC4AF2F49256AE58101FEE92B04,
    gx=0xA3E8EB3CC1CFE7B7732213B23A656149AFA142C47AAFBC2B79A191562E1305F4,
    gy=0x2D996C823439C56D7F7B22E14644417E69BCB6DE39D027001DABE8F35B25C9BE,
)
# bignp256/src/arithmetic.rs:38-57 (constants written little-endian there), bignp256/src/arithmetic/field.rs:59-65,
# bignp256/src/lib.rs:74,102: the one curve whose canonical records are little-endian
BIGNP256 = Curve(
    "bignp256",
    p=2**256 - 189,
    n=0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFD95C8ED60DFB4DFC7E5ABF99263D6607,
    a=2**256 - 189 - 3,
    b=0x77CE6C1515F3A8EDD2C13AABE4D8FBBE4CF55069978B9253B22E7D6BD69C03F1,
    gx=0,
    gy=0x6BF7FC3CFB16D69F5CE4C9A351D6835D78913966C408F6521E29CF1804516A93,
    le=True,
)
_BP384_P = 0x8CB91E82A3386D280F5D6F7E50E641DF152F7109ED5456B412B1DA197FB71123ACD3A729901D1A71874700133107EC53
_BP384_N = 0x8CB91E82A3386D280F5D6F7E50E641DF152F7109ED5456B31F166E6CAC0425A7CF3AB6AF6B7FC3103B883202E9046565
# bp384/src/r1/arithmetic.rs:34-53 (general a), bp384/src/arithmetic/field.rs:53, bp384/src/lib.rs:73
BP384R1 = Curve(
    "bp384r1", p=_BP384_P, n=_BP384_N,
    a=0x7BC382C63D8C150C3C72080ACE05AFA0C2BEA28E4FB22787139165EFBA91F90F8AA5814A503AD4EB04A8C7DD22CE2826,
    b=0x04A8C7DD22CE28268B39B55416F0447C2FB77DE107DCD2A62E880EA53EEB62D57CB4390295DBC9943AB78696FA504C11,
    gx=0x1D1C64F068CF45FFA2A63A81B7C13F6B8847A3E77EF14FE3DB7FCAFE0CBD10E8E826E03436D646AAEF87B2E247D4AF1E,
    gy=0x8ABE1D7520F9C2A45CB1EB8E95CFD55262B70B29FEEC5864E19C054FF99129280E4646217791811142820341263C5315,
)
# bp384/src/t1/arithmetic.rs:34-51
BP384T1 = Curve(
    "bp384t1", p=_BP384_P, n=_BP384_N, a=-3,
    b=0x7F519EADA7BDA81BD826DBA647910F8C4B9346ED8CCDC64E4B1ABD11756DCE1D2074AA263B88805CED70355A33B471EE,
    gx=0x18DE98B02DB9A306F2AFCD7235F72A819B80AB12EBD653172476FECD462AABFFC4FF191B946A5F54D8D0AA2F418808CC,
    gy=0x25AB056962D30651A114AFD2755AD336747F93475B7A1FCA3B88F2B6A208CCFE469408584DC2B2912675BF5B9E582928,
)
# p224/src/arithmetic.rs:40-56, p224/src/arithmetic/field.rs:54-60, p224/src/lib.rs:50-55
P224 = Curve(
    "p224",
    p=2**224 - 2**96 + 1,
    n=0xFFFFFFFFFFFFFFFFFFFFFFFFFFFF16A2E0B8F03E13DD29455C5C2A3D,
    a=-3,
    b=0xB4050A850C04B3ABF54132565044B0B7D7BFD8BA270B39432355FFB4,
    gx=0xB70E0CBD6BB4BF7F321390B94A03C1D356C21122343280D6115C1D21,
    gy=0xBD376388B5F723FB4C22DFE6CD4375A05A07476444D5819985007E34,
)
# p192/src/arithmetic.rs:38-54, p192/src/arithmetic/field.rs:54, p192/src/lib.rs:41
P192 = Curve(
    "p192",
    p=2**192 - 2**64 - 1,
    n=0xFFFFFFFFFFFFFFFFFFFFFFFF99DEF836146BC9B1B4D22831,
    a=-3,
    b=0x64210519E59C80E70FA7E9AB72243049FEB8DEECC146B9B1,
    gx=0x188DA80EB03090F67CBF20EB43A18800F4FF0AFD82FF1012,
    gy=0x07192B95FFC8DA78631011ED6B24CDD573F977A11E794811,
)
# p521/src/arithmetic.rs:45-90, p521/src/arithmetic/field.rs:68-80, p521/src/lib.rs:51-74 (66-byte records)
P521 = Curve(
    "p521",
    p=2**521 - 1,
    n=0x01FFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFA51868783BF2F966B7FCC0148F709A5D03BB5C9B8899C47AEBB6FB71E91386409,
    a=-3,
    b=0x0051953EB9618E1C9A1F929A21A0B68540EEA2DA725B99B315F3B8B489918EF109E156193951EC7E937B1652C0BD3BB1BF073573DF883D2C34F1EF451FD46B503F00,
    gx=0x00C6858E06B70404E9CD9E3ECB662395B4429C648139053FB521F828AF606B4D3DBAA14B5E77EFE75928FE1DC127A2FFA8DE3348B3C1856A429BF97E7E31C2E5BD66,
    gy=0x011839296A789A3BC0045C8A5FB42C7D1BD998F54449579B446817AFBD17273E662C97EE72995EF42640C550B9013FAD0761353C7086A272C24088BE94769FD16650,
)
EXT_CURVES = {3: SM2, 4: BP256R1, 5: BP256T1, 6: BIGNP256, 7: BP384R1, 8: BP384T1, 9: P224, 10: P192, 11: P521}
CURVE_IDS = {"k256": 0, "p256": 1, "p384": 2}
CURVE_IDS.update({c.name: i for i, c in EXT_CURVES.items()})
CURVES = {"k256": K256, "p256": P256, "p384": P384, 0: K256, 1: P256, 2: P384}
CURVES.update(EXT_CURVES)
CURVES.update({c.name: c for c in EXT_CURVES.values()})


def byteorder(c: Curve) -> str:
    return "little" if c.le else "big"


def enc_fe(c: Curve, v: int) -> bytes:
    """one canonical record (field element or scalar) in the curve's byte order"""
    return v.to_bytes(fbytes(c), byteorder(c))


def dec_fe(c: Curve, b: bytes) -> int:
    return int.from_bytes(b, byteorder(c))


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


def ecdsa_recover(c: Curve, z: int, r: int, s: int, recid: int, low_s_only: bool = False):
    """ecdsa_core::VerifyingKey::recover_from_prehash (k256/src/ecdsa.rs:45-88 documents it; vectors :182-262): the public key,
    or None.  recid bit 0 = y of R odd, bit 1 = x of R was r + n."""
    n, p = c.n, c.p
    if not (0 < r < n and 0 < s < n) or not 0 <= recid < 4:
        return None
    if low_s_only and s > n // 2:
        return None
    x = r + (n if recid & 2 else 0)
    if x >= 1 << (8 * fbytes(c)) or x >= p:
        return None
    rhs = (x * x * x + c.a * x + c.b) % p
    y = pow(rhs, (p + 1) // 4, p)
    if y * y % p != rhs:
        return None
    if (y & 1) != (recid & 1):
        y = p - y
    ri = pow(r, -1, n)
    Q = add(c, mul(c, (-(z % n) * ri) % n, G(c)), mul(c, s * ri % n, (x, y)))
    if Q is None or not ecdsa_verify(c, z, r, s, Q):
        return None
    return Q


# ---- SM2DSA (sm2/src/dsa/verifying.rs:138-175, signing.rs:213-260; Z_A: sm2/src/distid.rs:21-47) ------------------------
def sm2_hash_z(distid: bytes, Q) -> bytes:
    import hashlib
    c = CURVES["sm2"]
    h = hashlib.new("sm3")
    h.update((8 * len(distid)).to_bytes(2, "big") + distid)
    for v in (c.a % c.p, c.b, c.gx, c.gy, Q[0], Q[1]):
        h.update(v.to_bytes(32, "big"))
    return h.digest()


def sm2_hash_msg(distid: bytes, Q, msg: bytes) -> bytes:
    import hashlib
    return hashlib.new("sm3", sm2_hash_z(distid, Q) + msg).digest()


def sm2dsa_verify(e: int, r: int, s: int, Q) -> bool:
    c = CURVES["sm2"]
    n = c.n
    if not (0 < r < n and 0 < s < n) or Q is None or not on_curve(c, Q):
        return False
    t = (r + s) % n
    if t == 0:
        return False
    R = add(c, mul(c, s, G(c)), mul(c, t, Q))
    return R is not None and (e + R[0]) % n == r


def sm2dsa_sign(d: int, e: int, k: int):
    c = CURVES["sm2"]
    n = c.n
    r = (e + mul(c, k, G(c))[0]) % n
    s = pow(1 + d, -1, n) * (k - r * d) % n
    return r, s


# ---- hash to curve (RFC 9380), the suites the reference implements with SHA-256 -------------------------------------
# hash2curve/src/group_digest.rs:88-143 (hash_from_bytes / encode_from_bytes / hash_to_scalar),
# hash2curve/src/hash2field.rs + hash2field/expand_msg/xmd.rs (hash_to_field over expand_message_xmd),
# primeorder/src/osswu.rs:60-146 (simplified SWU for q = 3 mod 4), k256/src/arithmetic/hash2curve.rs:52-148
# (secp256k1: SSWU on the 3-isogenous curve E' + isogeny map, :169-258), p256/src/arithmetic/hash2curve.rs:44-75.
H2C_SUITES = {
    "k256": {"ro": b"secp256k1_XMD:SHA-256_SSWU_RO_", "nu": b"secp256k1_XMD:SHA-256_SSWU_NU_", "Z": -11,
             "A": 0x3F8731ABDD661ADCA08A5558F0F5D272E953D363CB6F0E5D405447C01A444533, "B": 1771},
    "p256": {"ro": b"P256_XMD:SHA-256_SSWU_RO_", "nu": b"P256_XMD:SHA-256_SSWU_NU_", "Z": -10, "A": -3, "B": P256.b},
    # p384/src/arithmetic/hash2curve.rs:13-75 (SHA-384, L = 72, Z = -12), p521/src/arithmetic/hash2curve.rs:13-76 (SHA-512, L = 98, Z = -4)
    "p384": {"ro": b"P384_XMD:SHA-384_SSWU_RO_", "nu": b"P384_XMD:SHA-384_SSWU_NU_", "Z": -12, "A": -3, "B": P384.b, "hash": "sha384", "L": 72},
    "p521": {"ro": b"P521_XMD:SHA-512_SSWU_RO_", "nu": b"P521_XMD:SHA-512_SSWU_NU_", "Z": -4, "A": -3, "hash": "sha512", "L": 98},
}
# 3-isogeny E' -> secp256k1 (RFC 9380 appendix E.1; k256/src/arithmetic/hash2curve.rs:170-239), coefficients of x^0..x^3
K256_ISO = {
    "xnum": [0x8E38E38E38E38E38E38E38E38E38E38E38E38E38E38E38E38E38E38DAAAAA8C7, 0x07D3D4C80BC321D5B9F315CEA7FD44C5D595D2FC0BF63B92DFFF1044F17C6581,
             0x534C328D23F234E6E2A413DECA25CAECE4506144037C40314ECBD0B53D9DD262, 0x8E38E38E38E38E38E38E38E38E38E38E38E38E38E38E38E38E38E38DAAAAA88C],
    "xden": [0xD35771193D94918A9CA34CCBB7B640DD86CD409542F8487D9FE6B745781EB49B, 0xEDADC6F64383DC1DF7C4B2D51B54225406D36B641F5E41BBC52A56612A8C6D14, 1],
    "ynum": [0x4BDA12F684BDA12F684BDA12F684BDA12F684BDA12F684BDA12F684B8E38E23C, 0xC75E0C32D5CB7C0FA9D0A54B12A0A6D5647AB046D686DA6FDFFC90FC201D71A3,
             0x29A6194691F91A73715209EF6512E576722830A201BE2018A765E85A9ECEE931, 0x2F684BDA12F684BDA12F684BDA12F684BDA12F684BDA12F684BDA12F38E38D84],
    "yden": [0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEFFFFF93B, 0x7A06534BB8BDB49FD5E9E6632722C2989467C1BFC8E8D978DFB425D2685C2573,
             0x6484AA716545CA2CF3A70C3FA8FE337E0A3D21162F0D6299A7BF8192BFD2A76F, 1],
}


def expand_message_xmd(msg: bytes, dst: bytes, len_in_bytes: int, hash_name: str = "sha256") -> bytes:
    """RFC 9380 section 5.3.1 (hash2curve/src/hash2field/expand_msg/xmd.rs:43-99) with SHA-256 / SHA-384 / SHA-512;
    oversize DSTs are hashed first (expand_msg.rs:15,76-95)."""
    H = lambda d: hashlib.new(hash_name, d).digest()  # noqa: E731
    b_in, s_in = hashlib.new(hash_name).digest_size, hashlib.new(hash_name).block_size
    if len(dst) > 255:
        dst = H(b"H2C-OVERSIZE-DST-" + dst)
    ell = (len_in_bytes + b_in - 1) // b_in
    assert 0 < len_in_bytes <= 65535 and ell <= 255
    dst_prime = dst + bytes([len(dst)])
    b0 = H(bytes(s_in) + msg + len_in_bytes.to_bytes(2, "big") + b"\x00" + dst_prime)
    b = [H(b0 + b"\x01" + dst_prime)]
    for i in range(2, ell + 1):
        b.append(H(bytes(x ^ y for x, y in zip(b0, b[-1])) + bytes([i]) + dst_prime))
    return b"".join(b)[:len_in_bytes]


def hash_to_field(msg: bytes, dst: bytes, count: int, modulus: int, L: int = 48, hash_name: str = "sha256"):
    u = expand_message_xmd(msg, dst, count * L, hash_name)
    return [int.from_bytes(u[L * i:L * (i + 1)], "big") % modulus for i in range(count)]


def _sqrt_3mod4(v: int, p: int):
    r = pow(v, (p + 1) // 4, p)
    return r if r * r % p == v % p else None


def sswu(u: int, p: int, A: int, B: int, Z: int):
    """simplified SWU, RFC 9380 section 6.6.2 (the definition; the reference's straight-line version gives the same point)"""
    A, B, Z = A % p, B % p, Z % p
    tv1 = (Z * Z % p * pow(u, 4, p) + Z * u * u) % p
    if tv1 == 0:
        x1 = B * pow(Z * A, -1, p) % p
    else:
        x1 = (-B) * pow(A, -1, p) % p * (1 + pow(tv1, -1, p)) % p
    gx1 = (pow(x1, 3, p) + A * x1 + B) % p
    x2 = Z * u * u % p * x1 % p
    gx2 = (pow(x2, 3, p) + A * x2 + B) % p
    y1 = _sqrt_3mod4(gx1, p)
    if y1 is not None:
        x, y = x1, y1
    else:
        x, y = x2, _sqrt_3mod4(gx2, p)
    if (u & 1) != (y & 1):
        y = p - y
    return x, y


def k256_iso_map(x: int, y: int):
    p = K256.p
    ev = lambda co: sum(c * pow(x, i, p) for i, c in enumerate(co)) % p  # noqa: E731
    xd, yd = ev(K256_ISO["xden"]), ev(K256_ISO["yden"])
    if xd == 0 or yd == 0:
        return None
    return ev(K256_ISO["xnum"]) * pow(xd, -1, p) % p, y * ev(K256_ISO["ynum"]) % p * pow(yd, -1, p) % p


def map_to_curve(curve: str, u: int):
    c = CURVES[curve]
    s = H2C_SUITES[curve]
    x, y = sswu(u, c.p, s["A"], s.get("B", c.b), s["Z"])
    return k256_iso_map(x, y) if curve == "k256" else (x, y)


def _h2c_field(curve: str, msg: bytes, dst: bytes, count: int, modulus: int):
    s = H2C_SUITES[curve]
    return hash_to_field(msg, dst, count, modulus, s.get("L", 48), s.get("hash", "sha256"))


def hash_to_curve(curve: str, msg: bytes, dst: bytes):
    """hash_from_bytes: two field elements, two maps, one addition (all four curves have cofactor 1)"""
    c = CURVES[curve]
    u0, u1 = _h2c_field(curve, msg, dst, 2, c.p)
    return add(c, map_to_curve(curve, u0), map_to_curve(curve, u1))


def encode_to_curve(curve: str, msg: bytes, dst: bytes):
    c = CURVES[curve]
    (u,) = _h2c_field(curve, msg, dst, 1, c.p)
    return map_to_curve(curve, u)


def hash_to_scalar(curve: str, msg: bytes, dst: bytes) -> int:
    return _h2c_field(curve, msg, dst, 1, CURVES[curve].n)[0]
