"""Shared helpers for the parity tests (oracle side = test infrastructure)."""
import json
import os
import random

import numpy as np

import pyref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(curve):
    with open(os.path.join(GOLDEN, f"{curve}.json")) as f:
        return json.load(f)


def pack_scalars(ks):
    return np.frombuffer(b"".join(int(k).to_bytes(32, "big") for k in ks), dtype=np.uint8).copy()


def pack_points(Ps):
    xy = bytearray()
    inf = bytearray()
    for P in Ps:
        b, f = pyref.enc_point(P)
        xy += b
        inf.append(f)
    return np.frombuffer(bytes(xy), dtype=np.uint8).copy(), np.frombuffer(bytes(inf), dtype=np.uint8).copy()


def unpack_points(out_xy, out_inf):
    out_xy = np.asarray(out_xy, dtype=np.uint8).reshape(-1, 64)
    res = []
    for i in range(out_xy.shape[0]):
        res.append(pyref.dec_point(out_xy[i].tobytes(), int(out_inf[i])))
    return res


def random_points(c, n, seed):
    """Uniform group elements t*G, t from a seeded RNG (k256/tests/projective.rs:21-25 builds points the same way)."""
    rng = random.Random(seed)
    G = pyref.G(c)
    return [pyref.mul(c, rng.randrange(1, c.n), G) for _ in range(n)]


def edge_scalars(c):
    n = c.n
    return [0, 1, 2, 3, 4, 7, 8, 15, 16, 17, 31, 32, 2**32 - 1, 2**32, 2**64, 2**127, 2**128 - 1, 2**128, 2**128 + 1,
            2**129, 2**255, n - 1, n - 2, n - 3, (n - 1) // 2, (n + 1) // 2, n // 3, pyref.K256_LAMBDA % n,
            (n - pyref.K256_LAMBDA) % n, int("5" * 64, 16) % n, int("a" * 64, 16) % n, int("f" * 63, 16)]


# ---- Wycheproof ECDSA vectors (tests/golden/*_wycheproof.json, from the reference's blobby files) ----

def parse_der_signature(sig: bytes, n: int, nb: int = 32):
    """Strict DER, as `ecdsa::der::Signature::from_der` + `Signature::from_scalars` accept it (the reference's
    Wycheproof harness, k256/src/ecdsa.rs:305-317): SEQUENCE { INTEGER r, INTEGER s }, minimal definite lengths,
    minimal non-negative INTEGERs of at most nb (the curve's field size) significant bytes, nothing trailing, 0 < r, s < n.
    -> (r, s) | None"""
    def length(buf, pos):
        if pos >= len(buf):
            return None
        b = buf[pos]
        if b < 0x80:
            return b, pos + 1
        k = b & 0x7F
        if k == 0 or k > 2 or pos + 1 + k > len(buf):
            return None
        v = int.from_bytes(buf[pos + 1:pos + 1 + k], "big")
        if v < 0x80 or (k == 2 and v < 0x100):
            return None  # not the minimal form
        return v, pos + 1 + k

    def integer(buf, pos):
        if pos >= len(buf) or buf[pos] != 0x02:
            return None
        got = length(buf, pos + 1)
        if got is None:
            return None
        ln, pos = got
        body = buf[pos:pos + ln]
        if ln == 0 or len(body) != ln or body[0] & 0x80:
            return None
        if ln > 1 and body[0] == 0 and not body[1] & 0x80:
            return None  # superfluous leading zero
        if len(body.lstrip(b"\x00")) > nb:
            return None
        return int.from_bytes(body, "big"), pos + ln

    if len(sig) < 2 or sig[0] != 0x30:
        return None
    got = length(sig, 1)
    if got is None:
        return None
    ln, pos = got
    if pos + ln != len(sig):
        return None
    a = integer(sig, pos)
    if a is None:
        return None
    b = integer(sig, a[1])
    if b is None or b[1] != len(sig):
        return None
    r, s = a[0], b[0]
    if not (0 < r < n and 0 < s < n):
        return None
    return r, s


def parse_p1363_signature(sig: bytes, n: int):
    """`Signature::from_slice`: exactly r || s, 32 bytes each, both in [1, n)."""
    if len(sig) != 64:
        return None
    r, s = int.from_bytes(sig[:32], "big"), int.from_bytes(sig[32:], "big")
    if not (0 < r < n and 0 < s < n):
        return None
    return r, s


def wycheproof_cases(curve: str):
    """-> (cases, rejected): cases = [(z32 bytes, r, s, (qx, qy), expected bool)], after the parsing and (k256 only)
    `normalize_s` the reference's harness applies before `verify`; rejected = vectors whose signature does not parse
    (each must be an expected failure)."""
    import hashlib
    c = pyref.CURVES[curve]
    nb = pyref.fbytes(c)
    # the curve's DigestPrimitive (p224/src/ecdsa.rs, p384/src/ecdsa.rs, p521/src/ecdsa.rs), then bits2field: the leftmost
    # field-size bytes of a longer digest, a shorter one right-aligned (the `ecdsa` crate's hazmat::bits2field)
    digest = {"k256": "sha256", "p256": "sha256", "p224": "sha224", "p384": "sha384", "p521": "sha512"}[curve]
    vec = json.load(open(os.path.join(GOLDEN, f"{curve}_wycheproof.json")))["vectors"]
    cases, rejected = [], []
    for v in vec:
        wx, wy = bytes.fromhex(v["wx"]), bytes.fromhex(v["wy"])
        assert not any(wx[:-nb]) and not any(wy[:-nb])          # element_from_padded_slice
        q = (int.from_bytes(wx, "big"), int.from_bytes(wy, "big"))
        sig = bytes.fromhex(v["sig"])
        rs = parse_der_signature(sig, c.n, nb) if v["fmt"] == "der" else parse_p1363_signature(sig, c.n)
        if rs is None:
            rejected.append(v)
            continue
        r, s = rs
        if curve == "k256" and s > c.n // 2:
            s = c.n - s                                          # Signature::normalize_s
        h = hashlib.new(digest, bytes.fromhex(v["msg"])).digest()
        z = h[:nb] if len(h) >= nb else bytes(nb - len(h)) + h
        cases.append((z, r, s, q, bool(v["pass"])))
    return cases, rejected
