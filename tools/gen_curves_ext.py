#!/usr/bin/env python3
"""Generate elliptic-curves_b200/csrc/ecg_curves_ext.cuh: parameter structs for the curves served by the generic
Montgomery field policy (ecg_fe_mont.cuh).  Inputs are the PUBLIC curve constants as they appear in the reference
(file:line given per curve); everything else (R mod p, R^2 mod p, -p^-1 mod 2^32, Montgomery forms of a, b, G) is
derived here with Python integers.  Run: python tools/gen_curves_ext.py   (the generated header is committed).
"""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "elliptic-curves_b200", "csrc", "ecg_curves_ext.cuh")


def le(hexstr):
    """bign-curve256v1 constants are written little-endian in the reference (from_hex_vartime with ByteOrder::LittleEndian)"""
    return int.from_bytes(bytes.fromhex(hexstr), "little")


# id, name (C++), enum, NL, p, n, a, b, gx, gy, little-endian records, source
CURVES = [
    dict(id=3, name="Sm2", enum="ECG_SM2", nl=8,
         p=0xFFFFFFFEFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFF00000000FFFFFFFFFFFFFFFF,
         n=0xFFFFFFFEFFFFFFFFFFFFFFFFFFFFFFFF7203DF6B21C6052B53BBF40939D54123, a=-3,
         b=0x28E9FA9E9D9F5E344D5A9E4BCF6509A7F39789F515AB8F92DDBCBD414D940E93,
         gx=0x32C4AE2C1F1981195F9904466A39C9948FE30BBFF2660BE1715A4589334C74C7,
         gy=0xBC3736A2F4F6779C59BDCEE36B692153D0A9877CC62A474002DF32E52139F0A0, le=False,
         src="sm2/src/arithmetic.rs:44-67, sm2/src/arithmetic/field.rs:34, sm2/src/lib.rs:86"),
    dict(id=4, name="Bp256r1", enum="ECG_BP256R1", nl=8, field="Bp256",
         p=0xA9FB57DBA1EEA9BC3E660A909D838D726E3BF623D52620282013481D1F6E5377,
         n=0xA9FB57DBA1EEA9BC3E660A909D838D718C397AA3B561A6F7901E0E82974856A7,
         a=0x7D5A0975FC2C3057EEF67530417AFFE7FB8055C126DC5C6CE94A4B44F330B5D9,
         b=0x26DC5C6CE94A4B44F330B5D9BBD77CBF958416295CF7E1CE6BCCDC18FF8C07B6,
         gx=0x8BD2AEB9CB7E57CB2C4B482FFC81B7AFB9DE27E1E3BD23C23A4453BD9ACE3262,
         gy=0x547EF835C3DAC4FD97F8461A14611DC9C27745132DED8E545C1D54C72F046997, le=False,
         src="bp256/src/r1/arithmetic.rs:34-53 (EquationAIsGeneric), bp256/src/arithmetic/field.rs:53, bp256/src/lib.rs:70"),
    dict(id=5, name="Bp256t1", enum="ECG_BP256T1", nl=8, field="Bp256",
         p=0xA9FB57DBA1EEA9BC3E660A909D838D726E3BF623D52620282013481D1F6E5377,
         n=0xA9FB57DBA1EEA9BC3E660A909D838D718C397AA3B561A6F7901E0E82974856A7, a=-3,
         b=0x662C61C430D84EA4FE66A7733D0B76B7BF93EBC4AF2F49256AE58101FEE92B04,
         gx=0xA3E8EB3CC1CFE7B7732213B23A656149AFA142C47AAFBC2B79A191562E1305F4,
         gy=0x2D996C823439C56D7F7B22E14644417E69BCB6DE39D027001DABE8F35B25C9BE, le=False,
         src="bp256/src/t1/arithmetic.rs:34-51"),
    dict(id=6, name="BignP256", enum="ECG_BIGNP256", nl=8,
         p=2**256 - 189,
         n=0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFD95C8ED60DFB4DFC7E5ABF99263D6607,
         a=le("40FFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFF"),
         b=le("F1039CD66B7D2EB253928B976950F54CBEFBD8E4AB3AC1D2EDA8F315156CCE77"),
         gx=0,
         gy=le("936A510418CF291E52F608C4663991785D83D651A3C9E45C9FD616FB3CFCF76B"), le=True,
         src="bignp256/src/arithmetic.rs:38-57 (EquationAIsGeneric; constants written little-endian), "
             "bignp256/src/arithmetic/field.rs:59-65, bignp256/src/lib.rs:74,102"),
    dict(id=7, name="Bp384r1", enum="ECG_BP384R1", nl=12, field="Bp384",
         p=0x8CB91E82A3386D280F5D6F7E50E641DF152F7109ED5456B412B1DA197FB71123ACD3A729901D1A71874700133107EC53,
         n=0x8CB91E82A3386D280F5D6F7E50E641DF152F7109ED5456B31F166E6CAC0425A7CF3AB6AF6B7FC3103B883202E9046565,
         a=0x7BC382C63D8C150C3C72080ACE05AFA0C2BEA28E4FB22787139165EFBA91F90F8AA5814A503AD4EB04A8C7DD22CE2826,
         b=0x04A8C7DD22CE28268B39B55416F0447C2FB77DE107DCD2A62E880EA53EEB62D57CB4390295DBC9943AB78696FA504C11,
         gx=0x1D1C64F068CF45FFA2A63A81B7C13F6B8847A3E77EF14FE3DB7FCAFE0CBD10E8E826E03436D646AAEF87B2E247D4AF1E,
         gy=0x8ABE1D7520F9C2A45CB1EB8E95CFD55262B70B29FEEC5864E19C054FF99129280E4646217791811142820341263C5315, le=False,
         src="bp384/src/r1/arithmetic.rs:34-53 (EquationAIsGeneric), bp384/src/arithmetic/field.rs:53, bp384/src/lib.rs:73"),
    dict(id=8, name="Bp384t1", enum="ECG_BP384T1", nl=12, field="Bp384",
         p=0x8CB91E82A3386D280F5D6F7E50E641DF152F7109ED5456B412B1DA197FB71123ACD3A729901D1A71874700133107EC53,
         n=0x8CB91E82A3386D280F5D6F7E50E641DF152F7109ED5456B31F166E6CAC0425A7CF3AB6AF6B7FC3103B883202E9046565, a=-3,
         b=0x7F519EADA7BDA81BD826DBA647910F8C4B9346ED8CCDC64E4B1ABD11756DCE1D2074AA263B88805CED70355A33B471EE,
         gx=0x18DE98B02DB9A306F2AFCD7235F72A819B80AB12EBD653172476FECD462AABFFC4FF191B946A5F54D8D0AA2F418808CC,
         gy=0x25AB056962D30651A114AFD2755AD336747F93475B7A1FCA3B88F2B6A208CCFE469408584DC2B2912675BF5B9E582928, le=False,
         src="bp384/src/t1/arithmetic.rs:34-51"),
    dict(id=9, name="P224", enum="ECG_NISTP224", nl=7,
         p=2**224 - 2**96 + 1,
         n=0xFFFFFFFFFFFFFFFFFFFFFFFFFFFF16A2E0B8F03E13DD29455C5C2A3D, a=-3,
         b=0xB4050A850C04B3ABF54132565044B0B7D7BFD8BA270B39432355FFB4,
         gx=0xB70E0CBD6BB4BF7F321390B94A03C1D356C21122343280D6115C1D21,
         gy=0xBD376388B5F723FB4C22DFE6CD4375A05A07476444D5819985007E34, le=False,
         src="p224/src/arithmetic.rs:40-56, p224/src/arithmetic/field.rs:54-60, p224/src/lib.rs:50-55"),
    dict(id=10, name="P192", enum="ECG_NISTP192", nl=6,
         p=2**192 - 2**64 - 1,
         n=0xFFFFFFFFFFFFFFFFFFFFFFFF99DEF836146BC9B1B4D22831, a=-3,
         b=0x64210519E59C80E70FA7E9AB72243049FEB8DEECC146B9B1,
         gx=0x188DA80EB03090F67CBF20EB43A18800F4FF0AFD82FF1012,
         gy=0x07192B95FFC8DA78631011ED6B24CDD573F977A11E794811, le=False,
         src="p192/src/arithmetic.rs:38-54, p192/src/arithmetic/field.rs:54, p192/src/lib.rs:41"),
    dict(id=11, name="P521", enum="ECG_NISTP521", nl=17, fb=66,
         p=2**521 - 1,
         n=0x01FFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFA51868783BF2F966B7FCC0148F709A5D03BB5C9B8899C47AEBB6FB71E91386409,
         a=-3,
         b=0x0051953EB9618E1C9A1F929A21A0B68540EEA2DA725B99B315F3B8B489918EF109E156193951EC7E937B1652C0BD3BB1BF073573DF883D2C34F1EF451FD46B503F00,
         gx=0x00C6858E06B70404E9CD9E3ECB662395B4429C648139053FB521F828AF606B4D3DBAA14B5E77EFE75928FE1DC127A2FFA8DE3348B3C1856A429BF97E7E31C2E5BD66,
         gy=0x011839296A789A3BC0045C8A5FB42C7D1BD998F54449579B446817AFBD17273E662C97EE72995EF42640C550B9013FAD0761353C7086A272C24088BE94769FD16650,
         le=False,
         src="p521/src/arithmetic.rs:45-90, p521/src/arithmetic/field.rs:68-80 (the reference's field is fiat-crypto's "
             "unsaturated Solinas form; the values are the same residues), p521/src/lib.rs:51-74; 66-byte records in 17 limbs"),
]


def limbs(v, nl):
    return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(nl)]


def carr(v, nl):
    return "{" + ", ".join("0x%08Xu" % x for x in limbs(v, nl)) + "}"


def accessor(name, v, nl):
    body = " : ".join("i == %d ? 0x%08Xu" % (i, x) for i, x in enumerate(limbs(v, nl)))
    return "  ECG_D static constexpr uint32_t %s(int i) { return %s : 0u; }\n" % (name, body)


def main():
    out = []
    out.append("// ecg_curves_ext.cuh — GENERATED by tools/gen_curves_ext.py; do not edit.\n")
    out.append("// Parameter structs for the curves served by the generic Montgomery field policy (ecg_fe_mont.cuh): the public\n")
    out.append("// constants of each curve as the reference states them (source lines per curve) and the Montgomery-domain values\n")
    out.append("// derived from them (R = 2^(32 NL)).\n#pragma once\n#include \"ecg_fe_mont.cuh\"\n#include \"ecg_point.cuh\"\n\n")
    out.append("#if defined(__CUDACC__)\n#define ECG_XCONSTANT static __device__ const\n#else\n#define ECG_XCONSTANT static const\n#endif\n\nnamespace ecg {\n\n")
    fields_done = set()
    host = []
    for c in CURVES:
        nl, p = c["nl"], c["p"]
        R = 1 << (32 * nl)
        assert p < R and p % 2 == 1
        a = c["a"] % p
        assert (c["gy"] ** 2 - (c["gx"] ** 3 + a * c["gx"] + c["b"])) % p == 0, c["name"]
        field = c.get("field", c["name"])
        # a == -3 (mod p) selects the a = -3 formulas; that includes bign-curve256v1, whose a the reference states as
        # the residue p - 3 and routes through EquationAIsGeneric (same group law, affine results identical)
        generic_a = (a != p - 3)
        out.append("// ---- %s: %s ----\n" % (c["name"], c["src"]))
        if field not in fields_done:
            fields_done.add(field)
            out.append("struct Mp%s {\n  static constexpr int NL = %d;\n  static constexpr int FB = %d;  // bytes per canonical record\n  static constexpr bool LE = %s;\n"
                       % (field, nl, c.get("fb", 4 * nl), "true" if c["le"] else "false"))
            out.append("  static constexpr uint32_t N0INV = 0x%08Xu;  // -p^-1 mod 2^32\n" % ((-pow(p, -1, 1 << 32)) % (1 << 32)))
            mers = p.bit_length() if p == (1 << p.bit_length()) - 1 else 0
            out.append("  static constexpr int MERSENNE = %d;  // B when p = 2^B - 1: the Montgomery reduction is a fold and a rotation (ecg_fe_mont.cuh)\n" % mers)
            out.append(accessor("P", p, nl))
            out.append(accessor("ONE", R % p, nl))
            out.append(accessor("R2", R * R % p, nl))
            out.append(accessor("PM2", p - 2, nl))
            if p % 4 == 3:
                out.append("  static constexpr bool HAS_SQRT_EXP = true;   // p = 3 (mod 4): sqrt(a) = a^((p+1)/4) when a is a square\n")
                out.append(accessor("PP14", (p + 1) // 4, nl))
            else:
                out.append("  static constexpr bool HAS_SQRT_EXP = false;  // p = 1 (mod 4): no single-exponentiation square root\n")
                out.append(accessor("PP14", 0, nl))
                # Tonelli-Shanks constants: p - 1 = 2^S * Q, Q odd; z = the smallest quadratic non-residue; ZQ = z^Q (Montgomery form)
                S_, Q_ = 0, p - 1
                while Q_ % 2 == 0:
                    S_, Q_ = S_ + 1, Q_ // 2
                z = 2
                while pow(z, (p - 1) // 2, p) != p - 1:
                    z += 1
                assert Q_ == (1 << Q_.bit_length()) - 1, "the addition chain in sqrt_candidate_ts assumes Q = 2^k - 1"
                out.append("  static constexpr int TS_S = %d;      // p - 1 = 2^%d * (2^%d - 1)\n  static constexpr int TS_QBITS = %d;\n" % (S_, S_, Q_.bit_length(), Q_.bit_length()))
                out.append(accessor("TS_ZQ", pow(z, Q_, p) * R % p, nl).replace("{ return", "{  // %d^Q * R mod p, %d the smallest non-residue\n    return" % (z, z)))
            out.append("};\n")
            out.append("ECG_XCONSTANT uint32_t %s_P[%d] = %s;\n" % (field.upper(), nl, carr(p, nl)))
        if generic_a:
            out.append("struct Fp%s : FpMontT<Mp%s> {\n  typedef Fp%s Inline;\n" % (c["name"], field, c["name"]))
            out.append("  ECG_D static void curve_a(FeT& a) {  // EQUATION_A in the Montgomery domain\n    const uint32_t t[%d] = %s;\n#pragma unroll\n    for (int i = 0; i < %d; i++) a.v[i] = t[i];\n  }\n};\n" % (nl, carr(a * R % p, nl), nl))
        else:
            out.append("typedef FpMontT<Mp%s> Fp%s;\n" % (field, c["name"]))
        out.append("ECG_XCONSTANT uint32_t %s_N[%d] = %s;\n" % (c["name"].upper(), nl, carr(c["n"], nl)))
        out.append("struct Curve%s {\n  typedef Fp%s F;\n  static constexpr int A_IS_MINUS3 = %d;\n" % (c["name"], c["name"], 2 if generic_a else 1))
        out.append("  ECG_D static const uint32_t* P() { return %s_P; }\n  ECG_D static const uint32_t* N() { return %s_N; }\n" % (field.upper(), c["name"].upper()))
        out.append("  ECG_D static void b_internal(FeN<%d>& b) {\n    const uint32_t t[%d] = %s;\n#pragma unroll\n    for (int i = 0; i < %d; i++) b.v[i] = t[i];\n  }\n" % (nl, nl, carr(c["b"] * R % p, nl), nl))
        out.append("  ECG_D static void generator(AffN<%d>& g) {\n    const uint32_t x[%d] = %s;\n    const uint32_t y[%d] = %s;\n#pragma unroll\n    for (int i = 0; i < %d; i++) {\n      g.x.v[i] = x[i];\n      g.y.v[i] = y[i];\n    }\n  }\n};\n\n"
                   % (nl, nl, carr(c["gx"] * R % p, nl), nl, carr(c["gy"] * R % p, nl), nl))
        fb = c.get("fb", 4 * nl)
        order = "little" if c["le"] else "big"
        gbytes = c["gx"].to_bytes(fb, order) + c["gy"].to_bytes(fb, order)
        host.append((c, gbytes))
    # scalar fields: the same Montgomery policy over the group order n (ECDSA verification: s^-1, u1, u2; hash_to_scalar)
    p384 = 2**384 - 2**128 - 2**96 + 2**32 - 1
    out.append("// (p + 1) / 4 of the hand-written P-384 field policy (square roots: SEC1 decompression)\nstruct SqrtExpP384 {\n  static constexpr bool HAS_SQRT_EXP = true;\n")
    out.append(accessor("PP14", (p384 + 1) // 4, 12))
    out.append("};\n")
    out.append("template <class C>\nstruct SqrtExp {\n  typedef typename C::F::Params T;\n};\ntemplate <>\nstruct SqrtExp<CurveP384> {\n  typedef SqrtExpP384 T;\n};\n")
    out.append("// ---- scalar fields: FpMontT over the group order (arithmetic mod n for ECDSA verification and hash_to_scalar) ----\n")
    out.append("template <class C>\nstruct ScalarField;\n")
    hot = [("K256", "CurveK256", 8, 32, 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141),
           ("P256", "CurveP256", 8, 32, 0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551),
           ("P384", "CurveP384", 12, 48, 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFC7634D81F4372DDF581A0DB248B0A77AECEC196ACCC52973)]
    allc = hot + [(c["name"], "Curve" + c["name"], c["nl"], c.get("fb", 4 * c["nl"]), c["n"]) for c in CURVES]
    done = set()
    for name, cv, nl, fb, n in allc:
        le = any(c["name"] == name and c["le"] for c in CURVES)
        key = (n, nl, le)
        R = 1 << (32 * nl)
        assert n % 2 == 1 and n < R
        out.append("struct Mn%s {\n  static constexpr int NL = %d;\n  static constexpr int FB = %d;\n  static constexpr bool LE = %s;\n" % (name, nl, fb, "true" if le else "false"))
        out.append("  static constexpr int MERSENNE = 0;\n")
        out.append("  static constexpr uint32_t N0INV = 0x%08Xu;  // -n^-1 mod 2^32\n" % ((-pow(n, -1, 1 << 32)) % (1 << 32)))
        out.append(accessor("P", n, nl))
        out.append(accessor("ONE", R % n, nl))
        out.append(accessor("R2", R * R % n, nl))
        out.append(accessor("PM2", n - 2, nl))
        out.append("};\ntemplate <>\nstruct ScalarField<%s> {\n  typedef FpMontT<Mn%s> T;\n};\n" % (cv, name))
    out.append("\n}  // namespace ecg\n\n")
    # host-side table: order limbs, generator bytes (ABI byte order), sizes
    out.append("// host-side descriptors (ecgpu.cu: sizes at the ABI, fixed-base table construction)\n")
    out.append("struct EcgExtCurveHost {\n  int id, nl, fb, le;  // limbs, bytes per record, little-endian records\n  uint32_t n[18];\n  uint8_t g[136];\n};\n")
    out.append("static const EcgExtCurveHost ECG_EXT_CURVES[] = {\n")
    for c, g in host:
        out.append("    {%d, %d, %d, %d, %s, {%s}},  // %s\n" % (c["id"], c["nl"], c.get("fb", 4 * c["nl"]), 1 if c["le"] else 0, carr(c["n"], c["nl"]),
                                                              ", ".join("0x%02X" % b for b in g), c["enum"]))
    out.append("};\nstatic const int ECG_EXT_CURVE_COUNT = %d;\n" % len(host))
    open(OUT, "w").write("".join(out))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
