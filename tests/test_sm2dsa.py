"""SM2DSA batch verification (SURVEY.md section 8(f) rank 1 widened to the sm2 crate's own signature scheme:
sm2/src/dsa/verifying.rs:138-175 — s*G + t*Q with t = r + s, verdict (e + x1) mod n == r).

Pinned to the one signature vector the reference holds (sm2/tests/sm2dsa.rs:16-34: an OpenSSL-made signature over "testing" with
the distinguishing identifier "example@rustcrypto.org"; Z_A and e through SM3 as sm2/src/distid.rs:21-47 builds them), then
signatures made here with the model, each with the corruptions the reference's proptests apply (sm2/tests/sm2dsa.rs:76-91,
tests/dsa_extended.rs:41-56: every signature byte flipped).  CPU: the kernels on the host.  GPU: through the C ABI."""
import ctypes
import hashlib
import json
import os
import random

import numpy as np
import pytest

import pyref
from test_curves_ext import ext_fb_table, recs

HERE = os.path.dirname(os.path.abspath(__file__))
C = pyref.CURVES["sm2"]
# tests/golden/sig_extras.json <- oracle/extract_golden.py <- sm2/tests/sm2dsa.rs:16-34 (PUBLIC_KEY, IDENTITY, MSG, SIG)
_X = json.load(open(os.path.join(HERE, "golden", "sig_extras.json")))["sm2dsa"]
assert _X["public_key"][:2] == "04"
REF_Q = (int(_X["public_key"][2:66], 16), int(_X["public_key"][66:], 16))
REF_ID, REF_MSG = _X["identity"].encode(), _X["msg"].encode()
REF_R, REF_S = int(_X["sig"][:64], 16), int(_X["sig"][64:], 16)

needs_sm3 = pytest.mark.skipif("sm3" not in hashlib.algorithms_available, reason="hashlib without SM3")


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def ref_case():
    e = pyref.sm2_hash_msg(REF_ID, REF_Q, REF_MSG)
    return (e, REF_R, REF_S, REF_Q, True)


def made_cases(count, seed):
    rng = random.Random(seed)
    n = C.n
    out = []
    for i in range(count):
        d, k = rng.randrange(1, n - 1), rng.randrange(1, n)
        e = rng.randrange(1 << 256)                       # any 32-byte digest; Scalar::reduce folds it
        r, s = pyref.sm2dsa_sign(d, e % n, k)
        Q = pyref.mul(C, d, pyref.G(C))
        eb = e.to_bytes(32, "big")
        if r == 0 or s == 0 or (r + k) % n == 0:
            continue
        out.append((eb, r, s, Q, True))
        how = i % 6
        if how == 0:
            out.append((eb, r, (s + 1) % n or 1, Q, None))
        elif how == 1:
            out.append(((e ^ 1).to_bytes(32, "big"), r, s, Q, None))
        elif how == 2:
            out.append((eb, r, s, pyref.mul(C, d + 1, pyref.G(C)), None))
        elif how == 3:
            out.append((eb, 0, s, Q, False))               # r = 0
        elif how == 4:
            out.append((eb, r, n - r, Q, False))           # t = r + s = 0 (mod n)
        else:
            out.append((eb, r, n, Q, False))               # s out of range
    # every byte of one signature flipped (tests/dsa_extended.rs:41-56)
    eb, r, s, Q, _ = out[0]
    sig = bytearray(r.to_bytes(32, "big") + s.to_bytes(32, "big"))
    for j in range(64):
        sig[j] ^= 1
        out.append((eb, int.from_bytes(sig[:32], "big"), int.from_bytes(sig[32:], "big"), Q, None))
        sig[j] ^= 1
    return [(eb, r, s, Q, pyref.sm2dsa_verify(int.from_bytes(eb, "big"), r, s, Q) if x is None else x) for eb, r, s, Q, x in out]


def pack(cases):
    E = np.frombuffer(b"".join(c[0] for c in cases), np.uint8).copy()
    S = np.frombuffer(b"".join((c[1] % (1 << 256)).to_bytes(32, "big") + (c[2] % (1 << 256)).to_bytes(32, "big") for c in cases), np.uint8).copy()
    Q = np.frombuffer(b"".join(c[3][0].to_bytes(32, "big") + c[3][1].to_bytes(32, "big") for c in cases), np.uint8).copy()
    return E, S, Q, [c[4] for c in cases]


@needs_sm3
def test_model_on_the_reference_vector():
    e, r, s, Q, _ = ref_case()
    assert pyref.sm2dsa_verify(int.from_bytes(e, "big"), r, s, Q)
    assert not pyref.sm2dsa_verify(int.from_bytes(e, "big") ^ 1, r, s, Q)
    assert not pyref.sm2dsa_verify(int.from_bytes(pyref.sm2_hash_msg(b"other@rustcrypto.org", Q, REF_MSG), "big"), r, s, Q)


def test_kernels_on_host():
    import __graft_entry__ as ge
    ge.build()
    sim = ctypes.CDLL(os.path.join(HERE, "sim", "libecgsim.so"))
    cid = pyref.CURVE_IDS["sm2"]
    table = ext_fb_table(sim, cid)
    cases = made_cases(8, 3)
    if "sm3" in hashlib.algorithms_available:
        cases.append(ref_case())
    E, S, Q, exp = pack(cases)
    valid = np.full(len(cases), 7, np.uint8)
    assert sim.simk_sm2dsa_verify(ctypes.c_size_t(len(cases)), _p(E), _p(S), _p(Q), _p(table), _p(valid)) == 0
    assert [bool(v) for v in valid] == exp and sum(exp) >= 8


@pytest.mark.gpu
def test_gpu_sm2dsa_verify(engine):
    cases = made_cases(300, 9)
    if "sm3" in hashlib.algorithms_available:
        cases.append(ref_case())
    bad_q = (cases[0][3][0], (cases[0][3][1] + 1) % C.p)
    cases.append((cases[0][0], cases[0][1], cases[0][2], bad_q, False))       # off-curve key: a per-signature failure
    E, S, Q, exp = pack(cases)
    valid = engine.sm2dsa_verify_batch(E, S, Q)
    wrong = [i for i, (v, x) in enumerate(zip(valid, exp)) if bool(v) != x]
    assert not wrong, f"verdict differs from the model at {wrong[:8]}"
    assert sum(exp) >= 300 and exp.count(False) >= 100
    assert engine.sm2dsa_verify_batch(np.zeros(0, np.uint8), np.zeros(0, np.uint8), np.zeros(0, np.uint8)).size == 0
