#!/usr/bin/env python3
"""Extract the reference's golden vectors into tests/golden/*.json.

Run once in the authoring container (needs /root/reference, which does NOT
exist on the GPU box): `python oracle/extract_golden.py`.  The JSON files it
writes are committed; tests read only those.

Sources (relative to /root/reference):
  k256/src/test_vectors/group.rs:9   ADD_TEST_VECTORS (k*G, k=1..20)
  k256/src/test_vectors/group.rs:96  MUL_TEST_VECTORS (k, x, y)
  p256/src/test_vectors/group.rs:8,95   same for P-256
  p384/src/test_vectors/group.rs:8,175  same for P-384 (48-byte values)
  {k256,p256}/src/test_vectors/field.rs:6  DBL_TEST_VECTORS (2^i, 32 B BE)
  {k256,p256}/src/test_vectors/ecdsa.rs    d -> (q_x, q_y) pairs (a k*G fixture each)
  {k256,p256}/benches/point.rs             the criterion bench scalars
  k256/src/ecdsa.rs:190-211,229-261        RECOVERY_TEST_VECTORS + the Ethereum example (public-key recovery)   -> sig_extras.json
  sm2/tests/sm2dsa.rs:16-34                the SM2DSA signature vector                                          -> sig_extras.json
  p384/tests/affine.rs:14-24               (UN)COMPRESSED_BASEPOINT                                             -> sig_extras.json
  (Wycheproof blobs, BIP340, FIPS ECDSA, hash2curve, the other curves' group vectors: see the functions below)
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

HEX = re.compile(r'hex!\(\s*"([0-9A-Fa-f]+)"\s*\)')


def hexes(text):
    return [h.lower() for h in HEX.findall(text)]


def split_consts(text):
    """Map const name -> text of its initialiser."""
    out = {}
    parts = re.split(r"pub const (\w+)", text)
    for i in range(1, len(parts), 2):
        out[parts[i]] = parts[i + 1]
    return out


def group_vectors(curve):
    txt = open(f"{REF}/{curve}/src/test_vectors/group.rs").read()
    c = split_consts(txt)
    add = hexes(c["ADD_TEST_VECTORS"])
    mul = hexes(c["MUL_TEST_VECTORS"]) if "MUL_TEST_VECTORS" in c else []
    assert len(add) % 2 == 0 and len(mul) % 3 == 0
    return {
        "source": f"{curve}/src/test_vectors/group.rs",
        "add": [{"k": i // 2 + 1, "x": add[i], "y": add[i + 1]} for i in range(0, len(add), 2)],
        "mul": [{"k": mul[i], "x": mul[i + 1], "y": mul[i + 2]} for i in range(0, len(mul), 3)],
    }


def field_vectors(curve):
    txt = open(f"{REF}/{curve}/src/test_vectors/field.rs").read()
    c = split_consts(txt)
    return {"source": f"{curve}/src/test_vectors/field.rs", "dbl": hexes(c["DBL_TEST_VECTORS"])}


def ecdsa_keypairs(curve):
    txt = open(f"{REF}/{curve}/src/test_vectors/ecdsa.rs").read()
    out = []
    for m in re.finditer(r'd:\s*&hex!\("([0-9a-fA-F]+)"\),\s*q_x:\s*&hex!\("([0-9a-fA-F]+)"\),\s*q_y:\s*&hex!\("([0-9a-fA-F]+)"\)', txt):
        out.append({"d": m.group(1).lower(), "x": m.group(2).lower(), "y": m.group(3).lower()})
    return {"source": f"{curve}/src/test_vectors/ecdsa.rs", "keypairs": out}


def bench_scalars(curve):
    txt = open(f"{REF}/{curve}/benches/point.rs").read()
    out = []
    for m in re.finditer(r"fn (test_scalar_\w)\(\) -> Scalar \{(.*?)\n\}", txt, re.S):
        body = m.group(2)
        hx = hexes(body)
        if hx:
            out.append({"name": m.group(1), "k": hx[0]})
        else:
            b = re.findall(r"0x([0-9a-fA-F]{2})", body)
            out.append({"name": m.group(1), "k": "".join(b).lower()})
    return {"source": f"{curve}/benches/point.rs", "scalars": out}


def bip340_vectors():
    """k256/src/schnorr.rs: BIP340_SIGN_VECTORS (all valid) and BIP340_VERIFY_VECTORS (valid flag)."""
    txt = open(f"{REF}/k256/src/schnorr.rs").read()
    out = []
    sign = txt[txt.index("const BIP340_SIGN_VECTORS"):txt.index("fn bip340_sign_vectors")]
    for m in re.finditer(r"SignVector\s*\{(.*?)\n        \},", sign, re.S):
        body = m.group(1)
        idx = int(re.search(r"index:\s*(\d+)", body).group(1))
        f = {k: re.sub(r"\s+", "", v) for k, v in re.findall(r'(\w+):\s*hex!\(\s*"([0-9A-Fa-f\s]+)"\s*\)', body)}
        out.append({"index": idx, "pk": f["public_key"].lower(), "msg": f["message"].lower(), "sig": f["signature"].lower(), "valid": True,
                    "sk": f["secret_key"].lower(), "aux": f["aux_rand"].lower()})
    ver = txt[txt.index("const BIP340_VERIFY_VECTORS"):txt.index("fn bip340_verify_vectors")]
    for m in re.finditer(r"VerifyVector\s*\{(.*?)\n        \},", ver, re.S):
        body = m.group(1)
        idx = int(re.search(r"index:\s*(\d+)", body).group(1))
        f = {k: re.sub(r"\s+", "", v) for k, v in re.findall(r'(\w+):\s*hex!\(\s*"([0-9A-Fa-f\s]+)"\s*\)', body)}
        valid = re.search(r"valid:\s*(true|false)", body).group(1) == "true"
        out.append({"index": idx, "pk": f["public_key"].lower(), "msg": f["message"].lower(), "sig": f["signature"].lower(), "valid": valid})
    return {"source": "k256/src/schnorr.rs (BIP340_SIGN_VECTORS, BIP340_VERIFY_VECTORS)", "vectors": out}


def ecdsa_full(curve):
    txt = open(f"{REF}/{curve}/src/test_vectors/ecdsa.rs").read()
    out = []
    for m in re.finditer(r"TestVector\s*\{(.*?)\}", txt, re.S):
        f = {k: v.lower() for k, v in re.findall(r'(\w+):\s*&hex!\(\s*"([0-9a-fA-F]+)"\s*\)', m.group(1))}
        if {"d", "q_x", "q_y", "k", "m", "r", "s"} <= set(f):
            out.append(f)
    return {"source": f"{curve}/src/test_vectors/ecdsa.rs", "vectors": out}


def _vlq(d, pos):
    """git-flavoured variable-length quantity used by the `blobby` 0.4 container"""
    b = d[pos]
    pos += 1
    val = b & 0x7F
    while b & 0x80:
        b = d[pos]
        pos += 1
        val = ((val + 1) << 7) + (b & 0x7F)
    return val, pos


def blobby(path):
    """blobby 0.4: VLQ blob count, VLQ size of the de-duplication table, its entries (VLQ length + bytes), then one
    VLQ per blob: odd = reference (n >> 1) into the table, even = (n >> 1) inline bytes."""
    d = open(path, "rb").read()
    total, pos = _vlq(d, 0)
    nd, pos = _vlq(d, pos)
    table = []
    for _ in range(nd):
        m, pos = _vlq(d, pos)
        table.append(d[pos:pos + m])
        pos += m
    blobs = []
    while pos < len(d):
        n, pos = _vlq(d, pos)
        if n & 1:
            blobs.append(table[n >> 1])
        else:
            blobs.append(d[pos:pos + (n >> 1)])
            pos += n >> 1
    assert pos == len(d) and len(blobs) == total and total % 5 == 0
    return blobs


def wycheproof(curve):
    """ECDSA verification vectors of the reference's Wycheproof tests (k256/src/ecdsa.rs:262-389,
    p256/src/ecdsa.rs:166-169): records of (wx, wy, msg, sig, pass); `fmt` = der | p1363."""
    out = []
    files = [("wycheproof.blb", "der")] + ([("wycheproof-p1316.blb", "p1363")] if curve == "k256" else [])
    for name, fmt in files:
        b = blobby(f"{REF}/{curve}/src/test_vectors/data/{name}")
        for i in range(0, len(b), 5):
            wx, wy, msg, sig, ok = b[i:i + 5]
            assert ok in (b"\x00", b"\x01")
            out.append({"wx": wx.hex(), "wy": wy.hex(), "msg": msg.hex(), "sig": sig.hex(), "pass": ok[0], "fmt": fmt})
    return {"source": f"{curve}/src/test_vectors/data/wycheproof*.blb", "vectors": out}


def h2c_vectors():
    """RFC 9380 vectors the reference tests hold (k256/src/arithmetic/hash2curve.rs:289-370, p256/src/arithmetic/
    hash2curve.rs:133-256) and the VOPRF hash_to_scalar vectors (p256/src/arithmetic/hash2curve.rs:257-310)."""
    out = {"source": "k256/src/arithmetic/hash2curve.rs, p256/src/arithmetic/hash2curve.rs (#[cfg(test)] vectors)", "suites": {}}
    for curve in ("k256", "p256", "p384", "p521"):
        txt = open(f"{REF}/{curve}/src/arithmetic/hash2curve.rs").read()
        dst = re.search(r'const DST: &\[u8\] = b"(QUUX[^"]+)"', txt).group(1)
        vecs = []
        for m in re.finditer(r'TestVector \{\s*msg: b"([^"]*)",(.*?)\},', txt, re.S):
            fields = dict(re.findall(r'(\w+): hex!\(\s*"([0-9a-f]+)"\s*\)', m.group(2)))
            if "p_x" in fields:
                vecs.append({"msg": m.group(1), **fields})
        out["suites"][curve] = {"dst": dst, "vectors": vecs}
    txt = open(f"{REF}/p256/src/arithmetic/hash2curve.rs").read()
    sc = []
    for m in re.finditer(r'dst: b"([^"]+)",\s*key_info: b"([^"]+)",\s*seed: &hex!\("([0-9a-f]+)"\),\s*sk_sm: &hex!\("([0-9a-f]+)"\)', txt):
        dst = m.group(1).encode().decode("unicode_escape").encode("latin1").hex()
        sc.append({"dst_hex": dst, "key_info": m.group(2), "seed": m.group(3), "sk_sm": m.group(4)})
    out["p256_hash_to_scalar_voprf"] = sc
    return out


def signature_extras():
    """Vectors of the callers widened in round 2:
    k256/src/ecdsa.rs:190-211 RECOVERY_TEST_VECTORS (pk, msg, sig, RecoveryId::new(is_y_odd, is_x_reduced)) and :229-261 the
    Ethereum end-to-end example (signing key, message, signature bytes, recovery id 0);
    sm2/tests/sm2dsa.rs:16-34 (PUBLIC_KEY, IDENTITY, MSG, SIG: an OpenSSL-made SM2DSA signature);
    p384/tests/affine.rs:14-24 (UNCOMPRESSED_BASEPOINT / COMPRESSED_BASEPOINT)."""
    def hx(t):
        return re.sub(r"\s+", "", t).lower()

    txt = open(f"{REF}/k256/src/ecdsa.rs").read()
    rec = []
    for m in re.finditer(r'RecoveryTestVector \{\s*pk: hex!\("([0-9a-fA-F]+)"\),\s*msg: b"([^"]*)",\s*sig: hex!\(\s*"([0-9a-fA-F\s]+)"\s*\),\s*'
                         r'recid: RecoveryId::new\((true|false), (true|false)\)', txt):
        rec.append({"pk": m.group(1).lower(), "msg": m.group(2), "sig": hx(m.group(3)),
                    "recid": (1 if m.group(4) == "true" else 0) | (2 if m.group(5) == "true" else 0)})
    assert len(rec) == 2
    eth = txt[txt.index("fn ethereum_end_to_end_example"):]
    key = re.search(r'SigningKey::from_bytes\(\s*&hex!\("([0-9a-f]+)"\)', eth).group(1)
    msg = re.search(r'let msg = hex!\(\s*"([0-9a-f]+)"\s*\)', eth).group(1)
    sig = re.search(r'sig\.to_bytes\(\)\.as_slice\(\),\s*&hex!\(\s*"([0-9a-f]+)"', eth).group(1)
    rid = int(re.search(r"RecoveryId::from_byte\((\d)\)", eth).group(1))
    sm = open(f"{REF}/sm2/tests/sm2dsa.rs").read()
    sm2 = {"public_key": re.search(r'PUBLIC_KEY: \[u8; 65\] = hex!\(\s*"([0-9A-Fa-f]+)"', sm).group(1).lower(),
           "identity": re.search(r'IDENTITY: &str = "([^"]+)"', sm).group(1),
           "msg": re.search(r'MSG: &\[u8\] = b"([^"]+)"', sm).group(1),
           "sig": "".join(h.lower() for h in re.findall(r'"([0-9a-f]{64})"', sm[sm.index("const SIG: [u8; 64]"):sm.index("fn verify_test_vector")]))}
    assert len(sm2["sig"]) == 128 and len(sm2["public_key"]) == 130
    af = open(f"{REF}/p384/tests/affine.rs").read()
    comp = hx(re.search(r'COMPRESSED_BASEPOINT: &\[u8\] = &hex!\(\s*"([0-9a-fA-F\s]+)"', af[af.index("const COMPRESSED_BASEPOINT"):]).group(1))
    unc = hx(re.search(r'UNCOMPRESSED_BASEPOINT: &\[u8\] = &hex!\(\s*"([0-9a-fA-F\s]+)"', af).group(1))
    assert len(comp) == 98 and len(unc) == 194
    return {"reference_commit": "739304e026fdf06cd1a31606e4db487d3f47c5ae",
            "k256_recovery": rec, "k256_ethereum": {"signing_key": key, "msg_hex": msg, "sig": sig, "recid": rid, "digest": "keccak256"},
            "sm2dsa": sm2, "p384_compressed_basepoint": comp, "p384_uncompressed_basepoint": unc}


def main():
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "sig_extras.json"), "w") as f:
        v = signature_extras()
        json.dump(v, f, indent=1)
        print("sig_extras", len(v["k256_recovery"]), "recovery vectors, ethereum example, sm2dsa vector, p384 basepoint encodings")
    for curve in ("k256", "p256", "p224", "p384", "p521"):
        with open(os.path.join(OUT, f"{curve}_wycheproof.json"), "w") as f:
            v = wycheproof(curve)
            json.dump(v, f, indent=0)
            print("wycheproof", curve, len(v["vectors"]), sum(x["pass"] for x in v["vectors"]), "valid")
    with open(os.path.join(OUT, "k256_bip340.json"), "w") as f:
        v = bip340_vectors()
        json.dump(v, f, indent=1)
        print("bip340", len(v["vectors"]))
    for curve in ("k256", "p256", "p192", "p224", "p384", "p521"):
        with open(os.path.join(OUT, f"{curve}_ecdsa.json"), "w") as f:
            v = ecdsa_full(curve)
            json.dump(v, f, indent=1)
            print("ecdsa", curve, len(v["vectors"]))
    for curve in ("k256", "p256"):
        data = {
            "curve": curve,
            "reference_commit": "739304e026fdf06cd1a31606e4db487d3f47c5ae",
            "group": group_vectors(curve),
            "field": field_vectors(curve),
            "ecdsa": ecdsa_keypairs(curve),
            "bench": bench_scalars(curve),
        }
        path = os.path.join(OUT, f"{curve}.json")
        with open(path, "w") as f:
            json.dump(data, f, indent=1)
        print(path, len(data["group"]["add"]), len(data["group"]["mul"]), len(data["field"]["dbl"]),
              len(data["ecdsa"]["keypairs"]), len(data["bench"]["scalars"]))
    # P-384: p384/src/test_vectors/group.rs:8,175 (ADD / MUL vectors, 48-byte coordinates),
    # p384/src/test_vectors/ecdsa.rs (d -> Q pairs); the crate has no field.rs vector file and no benches/point.rs scalars
    data = {"curve": "p384", "reference_commit": "739304e026fdf06cd1a31606e4db487d3f47c5ae", "group": group_vectors("p384"),
            "ecdsa": ecdsa_keypairs("p384")}
    path = os.path.join(OUT, "p384.json")
    with open(path, "w") as f:
        json.dump(data, f, indent=1)
    print(path, len(data["group"]["add"]), len(data["group"]["mul"]), len(data["ecdsa"]["keypairs"]))
    # the other prime-order curves with vector files: p224 / p192 (ADD + MUL, big-endian), bignp256 (ADD only, the hex
    # strings are the curve's little-endian records: bignp256/src/test_vectors/group.rs:8); sm2 and the brainpool crates
    # hold no group vectors (their parity is pinned to the big-integer model and OpenSSL)
    with open(os.path.join(OUT, "h2c.json"), "w") as f:
        v = h2c_vectors()
        json.dump(v, f, indent=1)
        print("h2c", {k: len(x["vectors"]) for k, x in v["suites"].items()}, len(v["p256_hash_to_scalar_voprf"]))
    for curve in ("p224", "p192", "bignp256", "p521"):
        data = {"curve": curve, "reference_commit": "739304e026fdf06cd1a31606e4db487d3f47c5ae", "group": group_vectors(curve),
                "little_endian": curve == "bignp256"}
        path = os.path.join(OUT, f"{curve}.json")
        with open(path, "w") as f:
            json.dump(data, f, indent=1)
        print(path, len(data["group"]["add"]), len(data["group"]["mul"]))


if __name__ == "__main__":
    main()
