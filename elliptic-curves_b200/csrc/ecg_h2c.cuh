// ecg_h2c.cuh — hash to curve (RFC 9380) for the two suites the reference implements with SHA-256:
//   secp256k1_XMD:SHA-256_SSWU_{RO,NU}_  (k256/src/arithmetic/hash2curve.rs:14-20,52-148: simplified SWU on the 3-isogenous
//                                          curve E', then the isogeny map :169-258)
//   P256_XMD:SHA-256_SSWU_{RO,NU}_       (p256/src/arithmetic/hash2curve.rs:13-75 over primeorder/src/osswu.rs:60-146)
// and the drivers hash2curve/src/group_digest.rs:88-143 (hash_from_bytes = two field elements, two maps, one addition;
// encode_from_bytes = one map; hash_to_scalar), hash2curve/src/hash2field.rs + hash2field/expand_msg/xmd.rs:43-99
// (hash_to_field over expand_message_xmd).  SURVEY.md section 8(f) rank 4 ("hash-to-curve front end").
//
// One thread per message: streaming SHA-256 over the message bytes (b_0), the chained blocks b_1..b_ell, the 48-byte
// reductions d0 * 2^192 + d1, the straight-line SSWU with ONE exponentiation per map (sqrt_ratio for q = 3 mod 4) and
// no inversion: the map leaves x as a fraction, which becomes the Z of a Jacobian point (secp256k1: the isogeny is
// evaluated on the homogenised polynomials), the two points are added with the library's Jacobian addition and the
// batch is normalised by normalize_kernel (one inversion per ~32 points).
#pragma once
#include "ecg_curves.cuh"
#include "ecg_verify.cuh"

namespace ecg {

// ---- streaming SHA-256 (byte-granular updates; compression from ecg_verify.cuh) --------------------------------
struct Sha256Stream {
  uint32_t st[8];
  uint32_t blk[16];
  uint32_t fill;   // bytes in blk
  uint64_t total;  // bytes absorbed
  ECG_D void init() {
    const uint32_t iv[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = iv[i];
#pragma unroll
    for (int i = 0; i < 16; i++) blk[i] = 0;
    fill = 0;
    total = 0;
  }
  ECG_D void flush_block() {
    sha256_compress(st, blk);
#pragma unroll
    for (int i = 0; i < 16; i++) blk[i] = 0;
    fill = 0;
  }
  ECG_D void put(uint32_t byte) {
    blk[fill >> 2] |= (byte & 0xFFu) << (24 - 8 * (fill & 3));
    fill++;
    total++;
    if (fill == 64) flush_block();
  }
  ECG_D void update(const uint8_t* p, size_t len) {
    for (size_t i = 0; i < len; i++) put(p[i]);
  }
  ECG_D void put_word_be(uint32_t w) {  // four bytes, most significant first
    put(w >> 24);
    put(w >> 16);
    put(w >> 8);
    put(w);
  }
  ECG_D void finish(uint32_t* digest /* 8 big-endian words */) {
    uint64_t bits = total * 8;
    put(0x80u);
    total--;  // padding is not message
    if (fill > 56) flush_block();
    blk[14] = (uint32_t)(bits >> 32);
    blk[15] = (uint32_t)bits;
    sha256_compress(st, blk);
#pragma unroll
    for (int i = 0; i < 8; i++) digest[i] = st[i];
  }
};

// uniform[0 .. 32*ELL) = expand_message_xmd(msg, DST, len_in_bytes) as big-endian words (RFC 9380 section 5.3.1,
// hash2curve/src/hash2field/expand_msg/xmd.rs:43-99).  dst_prime = DST || I2OSP(len(DST), 1) (an oversize DST is
// replaced by its hash on the host: expand_msg.rs:76-95).
template <int ELL>
ECG_D void expand_message_xmd_sha256(uint32_t* uniform, const uint8_t* msg, size_t msg_len, const uint8_t* dst_prime, uint32_t dst_prime_len,
                                     uint32_t len_in_bytes) {
  Sha256Stream h;
  uint32_t b0[8], bi[8];
  h.init();
  h.flush_block();  // Z_pad: one block of zero bytes
  h.total = 64;
  h.update(msg, msg_len);
  h.put(len_in_bytes >> 8);
  h.put(len_in_bytes);
  h.put(0);
  h.update(dst_prime, dst_prime_len);
  h.finish(b0);
#pragma unroll 1
  for (int i = 1; i <= ELL; i++) {
    h.init();
#pragma unroll
    for (int w = 0; w < 8; w++) h.put_word_be(i == 1 ? b0[w] : (b0[w] ^ bi[w]));
    h.put((uint32_t)i);
    h.update(dst_prime, dst_prime_len);
    h.finish(bi);
#pragma unroll
    for (int w = 0; w < 8; w++) uniform[8 * (i - 1) + w] = bi[w];
  }
}

// ---- per-suite constants (canonical integers; converted to the field's internal form where they are used) --------
// C2 = sqrt(-Z) as sqrt_ratio_3mod4 of primeorder/src/osswu.rs:60-88 uses it (P-256: sqrt(10), the reference's value;
// secp256k1: sqrt(11) — the reference's own k256 map is the older straight-line variant whose constant is sqrt(-Z^3),
// k256/src/arithmetic/hash2curve.rs:64-69; both produce the same point, the sign of y being fixed by sgn0 at the end).
template <class C>
struct H2cSuite;
template <>
struct H2cSuite<CurveK256> {
  static constexpr bool ISOGENY = true;
  ECG_D static void A(Fe& r) {
    const uint32_t t[8] = {0x1A444533u, 0x405447C0u, 0xCB6F0E5Du, 0xE953D363u, 0xF0F5D272u, 0xA08A5558u, 0xDD661ADCu, 0x3F8731ABu};
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
  }
  ECG_D static void B(Fe& r) {
    const uint32_t t[8] = {0x000006EBu, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u};
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
  }
  ECG_D static void Z(Fe& r) {
    const uint32_t t[8] = {0xFFFFFC24u, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
  }
  ECG_D static void C2(Fe& r) {
    const uint32_t t[8] = {0x303C4A59u, 0x286729C8u, 0xA74789DDu, 0xEC184F00u, 0x8F842AFEu, 0x7AD13FB3u, 0x724013E5u, 0x31FDF302u};
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
  }
  ECG_D static void F_2_192(Fe& r) {
    const uint32_t t[8] = {0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000001u, 0x00000000u};
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
  }
  ECG_D static uint32_t C1(int i) {  // (p - 3) / 4, the exponent of sqrt_ratio_3mod4
    const uint32_t t[8] = {0xBFFFFF0Bu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x3FFFFFFFu};
    return t[i];
  }
  ECG_D static void XNUM(Fe& r, int k) {  // coefficient of x^k
    const uint32_t t[4][8] = {{0xAAAAA8C7u, 0x8E38E38Du, 0xE38E38E3u, 0x38E38E38u, 0x8E38E38Eu, 0xE38E38E3u, 0x38E38E38u, 0x8E38E38Eu}, {0xF17C6581u, 0xDFFF1044u, 0x0BF63B92u, 0xD595D2FCu, 0xA7FD44C5u, 0xB9F315CEu, 0x0BC321D5u, 0x07D3D4C8u}, {0x3D9DD262u, 0x4ECBD0B5u, 0x037C4031u, 0xE4506144u, 0xCA25CAECu, 0xE2A413DEu, 0x23F234E6u, 0x534C328Du}, {0xAAAAA88Cu, 0x8E38E38Du, 0xE38E38E3u, 0x38E38E38u, 0x8E38E38Eu, 0xE38E38E3u, 0x38E38E38u, 0x8E38E38Eu}};
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[k][i];
  }
  ECG_D static void XDEN(Fe& r, int k) {  // coefficient of x^k
    const uint32_t t[3][8] = {{0x781EB49Bu, 0x9FE6B745u, 0x42F8487Du, 0x86CD4095u, 0xB7B640DDu, 0x9CA34CCBu, 0x3D94918Au, 0xD3577119u}, {0x2A8C6D14u, 0xC52A5661u, 0x1F5E41BBu, 0x06D36B64u, 0x1B542254u, 0xF7C4B2D5u, 0x4383DC1Du, 0xEDADC6F6u}, {0x00000001u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u}};
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[k][i];
  }
  ECG_D static void YNUM(Fe& r, int k) {  // coefficient of x^k
    const uint32_t t[4][8] = {{0x8E38E23Cu, 0xA12F684Bu, 0x12F684BDu, 0x2F684BDAu, 0xF684BDA1u, 0x684BDA12u, 0x84BDA12Fu, 0x4BDA12F6u}, {0x201D71A3u, 0xDFFC90FCu, 0xD686DA6Fu, 0x647AB046u, 0x12A0A6D5u, 0xA9D0A54Bu, 0xD5CB7C0Fu, 0xC75E0C32u}, {0x9ECEE931u, 0xA765E85Au, 0x01BE2018u, 0x722830A2u, 0x6512E576u, 0x715209EFu, 0x91F91A73u, 0x29A61946u}, {0x38E38D84u, 0x84BDA12Fu, 0x4BDA12F6u, 0xBDA12F68u, 0xDA12F684u, 0xA12F684Bu, 0x12F684BDu, 0x2F684BDAu}};
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[k][i];
  }
  ECG_D static void YDEN(Fe& r, int k) {  // coefficient of x^k
    const uint32_t t[4][8] = {{0xFFFFF93Bu, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, {0x685C2573u, 0xDFB425D2u, 0xC8E8D978u, 0x9467C1BFu, 0x2722C298u, 0xD5E9E663u, 0xB8BDB49Fu, 0x7A06534Bu}, {0xBFD2A76Fu, 0xA7BF8192u, 0x2F0D6299u, 0x0A3D2116u, 0xA8FE337Eu, 0xF3A70C3Fu, 0x6545CA2Cu, 0x6484AA71u}, {0x00000001u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u}};
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[k][i];
  }
};
template <>
struct H2cSuite<CurveP256> {
  static constexpr bool ISOGENY = false;
  ECG_D static void A(Fe& r) {
    const uint32_t t[8] = {0xFFFFFFFCu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000001u, 0xFFFFFFFFu};
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
  }
  ECG_D static void B(Fe& r) {
    const uint32_t t[8] = {0x27D2604Bu, 0x3BCE3C3Eu, 0xCC53B0F6u, 0x651D06B0u, 0x769886BCu, 0xB3EBBD55u, 0xAA3A93E7u, 0x5AC635D8u};
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
  }
  ECG_D static void Z(Fe& r) {
    const uint32_t t[8] = {0xFFFFFFF5u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000001u, 0xFFFFFFFFu};
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
  }
  ECG_D static void C2(Fe& r) {
    const uint32_t t[8] = {0xE433C47Fu, 0x2CCD3427u, 0x4C55D5B6u, 0x7B8D1FF8u, 0x5180AAB2u, 0xC978FC67u, 0xE1D89B99u, 0xDA538E3Bu};
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
  }
  ECG_D static void F_2_192(Fe& r) {
    const uint32_t t[8] = {0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000001u, 0x00000000u};
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
  }
  ECG_D static uint32_t C1(int i) {  // (p - 3) / 4, the exponent of sqrt_ratio_3mod4
    const uint32_t t[8] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0x3FFFFFFFu, 0x00000000u, 0x00000000u, 0x40000000u, 0xC0000000u, 0x3FFFFFFFu};
    return t[i];
  }
};

// r = a^((p-3)/4), fixed 4-bit windows over the public exponent (252 squarings + 63 + 14 multiplications)
template <class C>
ECG_D void h2c_pow_c1(Fe& r, const Fe& a) {
  typedef typename C::F F;
  Fe tab[16];
  F::set_one(tab[0]);
  tab[1] = a;
#pragma unroll 1
  for (int i = 2; i < 16; i++) F::mul(tab[i], tab[i - 1], a);
  Fe acc;
  F::set_one(acc);
#pragma unroll 1
  for (int w = 63; w >= 0; w--) {
    F::sqr_n(acc, acc, 4);
    uint32_t d = (H2cSuite<C>::C1(w >> 3) >> (4 * (w & 7))) & 15u;
    F::mul(acc, acc, tab[d]);
  }
  r = acc;
}

// sqrt_ratio for q = 3 (mod 4): primeorder/src/osswu.rs:60-88 (RFC 9380 F.2.1.2).  Returns is_square(u / v); y = sqrt(u / v) if
// it is one, sqrt(Z u / v) otherwise.
template <class C>
ECG_D bool h2c_sqrt_ratio(Fe& y, const Fe& u, const Fe& v, const Fe& c2) {
  typedef typename C::F F;
  Fe tv1, tv2, tv3, y1, y2;
  F::sqr(tv1, v);
  F::mul(tv2, u, v);
  F::mul(tv1, tv1, tv2);
  h2c_pow_c1<C>(y1, tv1);
  F::mul(y1, y1, tv2);
  F::mul(y2, y1, c2);
  F::sqr(tv3, y1);
  F::mul(tv3, tv3, v);
  F::sub(tv3, tv3, u);
  bool is_qr = F::is_zero(tv3);
#pragma unroll
  for (int i = 0; i < 8; i++) y.v[i] = is_qr ? y1.v[i] : y2.v[i];
  return is_qr;
}

// parity of the canonical representative (Sgn0: k256/src/arithmetic/hash2curve.rs:46-50, p256/.../hash2curve.rs:38-42)
template <class C>
ECG_D uint32_t h2c_sgn0(const Fe& a) {
  Fe t;
  C::F::to_canonical(t, a);
  return t.v[0] & 1u;
}

// simplified SWU, straight line (primeorder/src/osswu.rs:92-146, RFC 9380 F.2): the point (xn / xd, y) on the curve
// y^2 = x^3 + A x + B of the suite (secp256k1: the isogenous curve E').  u in internal form.
template <class C>
ECG_D void h2c_sswu(Fe& xn, Fe& xd, Fe& y, const Fe& u) {
  typedef typename C::F F;
  typedef H2cSuite<C> S;
  Fe A, B, Z, c2, one, tv1, tv2, tv3, tv4, tv5, tv6, x, y1, t;
  S::A(t);
  F::from_canonical(A, t);
  S::B(t);
  F::from_canonical(B, t);
  S::Z(t);
  F::from_canonical(Z, t);
  S::C2(t);
  F::from_canonical(c2, t);
  F::set_one(one);
  F::sqr(tv1, u);          // 1
  F::mul(tv1, Z, tv1);     // 2
  F::sqr(tv2, tv1);        // 3
  F::add(tv2, tv2, tv1);   // 4
  F::add(tv3, tv2, one);   // 5
  F::mul(tv3, B, tv3);     // 6
  F::neg(t, tv2);          // 7: tv4 = CMOV(Z, -tv2, tv2 != 0)
  bool z2 = F::is_zero(tv2);
#pragma unroll
  for (int i = 0; i < 8; i++) tv4.v[i] = z2 ? Z.v[i] : t.v[i];
  F::mul(tv4, A, tv4);     // 8
  F::sqr(tv2, tv3);        // 9
  F::sqr(tv6, tv4);        // 10
  F::mul(tv5, A, tv6);     // 11
  F::add(tv2, tv2, tv5);   // 12
  F::mul(tv2, tv2, tv3);   // 13
  F::mul(tv6, tv6, tv4);   // 14
  F::mul(tv5, B, tv6);     // 15
  F::add(tv2, tv2, tv5);   // 16
  F::mul(x, tv1, tv3);     // 17
  bool is_sq = h2c_sqrt_ratio<C>(y1, tv2, tv6, c2);  // 18
  F::mul(y, tv1, u);       // 19
  F::mul(y, y, y1);        // 20
#pragma unroll
  for (int i = 0; i < 8; i++) {  // 21, 22
    x.v[i] = is_sq ? tv3.v[i] : x.v[i];
    y.v[i] = is_sq ? y1.v[i] : y.v[i];
  }
  uint32_t flip = h2c_sgn0<C>(u) ^ h2c_sgn0<C>(y);  // 23, 24
  fe_cneg<F>(y, flip);
  xn = x;                  // 25: x = x / tv4, left as a fraction
  xd = tv4;
}

// map_to_curve as a Jacobian point of the target curve (no inversion).
//   P-256: (xn / xd, y) is the point: (X : Y : Z) = (xn xd : y xd^3 : xd).
//   secp256k1: the 3-isogeny E' -> E (k256/src/arithmetic/hash2curve.rs:169-258, RFC 9380 E.1) on x' = N / D:
//     x = XN / (D XD),  y = y' YN / YD  with the homogenised polynomials XN = sum k_i N^i D^(3-i), XD = N^2 + ..., so
//     (X : Y : Z) = (XN D XD YD^2 : y' YN D^3 XD^3 YD^2 : D XD YD); a vanishing denominator gives Z = 0, the identity
//     (RFC 9380 section 6.6.3: exceptional cases of the isogeny map to the identity).
template <class C>
ECG_D void h2c_map_to_curve(Jac& r, const Fe& u) {
  typedef typename C::F F;
  typedef H2cSuite<C> S;
  Fe N, D, y;
  h2c_sswu<C>(N, D, y, u);
  if constexpr (!S::ISOGENY) {
    Fe d2;
    F::mul(r.X, N, D);
    F::sqr(d2, D);
    F::mul(d2, d2, D);
    F::mul(r.Y, y, d2);
    r.Z = D;
  } else {
    Fe pw[2][4];  // pw[0][i] = N^i, pw[1][i] = D^i
    F::set_one(pw[0][0]);
    F::set_one(pw[1][0]);
    pw[0][1] = N;
    pw[1][1] = D;
    F::sqr(pw[0][2], N);
    F::sqr(pw[1][2], D);
    F::mul(pw[0][3], pw[0][2], N);
    F::mul(pw[1][3], pw[1][2], D);
    Fe XN, XD, YN, YD, k, t;
    F::set_zero(XN);
    F::set_zero(XD);
    F::set_zero(YN);
    F::set_zero(YD);
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
      Fe nd3;
      F::mul(nd3, pw[0][i], pw[1][3 - i]);  // N^i D^(3-i)
      S::XNUM(k, i);
      F::from_canonical(k, k);
      F::mul(t, k, nd3);
      F::add(XN, XN, t);
      S::YNUM(k, i);
      F::from_canonical(k, k);
      F::mul(t, k, nd3);
      F::add(YN, YN, t);
      S::YDEN(k, i);
      F::from_canonical(k, k);
      F::mul(t, k, nd3);
      F::add(YD, YD, t);
      if (i < 3) {
        Fe nd2;
        F::mul(nd2, pw[0][i], pw[1][2 - i]);  // N^i D^(2-i)
        S::XDEN(k, i);
        F::from_canonical(k, k);
        F::mul(t, k, nd2);
        F::add(XD, XD, t);
      }
    }
    Fe dxd, z, yd2, t3;
    F::mul(dxd, D, XD);       // D XD
    F::mul(z, dxd, YD);       // Z = D XD YD
    F::sqr(yd2, YD);
    F::mul(t, XN, dxd);
    F::mul(r.X, t, yd2);      // XN D XD YD^2
    F::sqr(t3, dxd);
    F::mul(t3, t3, dxd);      // (D XD)^3
    F::mul(t, y, YN);
    F::mul(t, t, t3);
    F::mul(r.Y, t, yd2);      // y' YN D^3 XD^3 YD^2
    r.Z = z;
  }
}

// 48 uniform bytes (12 big-endian words) -> field element d0 * 2^192 + d1 in internal form
// (Reduce<Array<u8, U48>> for FieldElement: k256/src/arithmetic/hash2curve.rs:22-44, p256/.../hash2curve.rs:21-36)
template <class C>
ECG_D void h2c_field_from_okm(Fe& r, const uint32_t* w /* 12 words, most significant first */) {
  typedef typename C::F F;
  Fe d0, d1, f;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    d0.v[i] = i < 6 ? w[5 - i] : 0u;
    d1.v[i] = i < 6 ? w[11 - i] : 0u;
  }
  F::from_canonical(d0, d0);
  F::from_canonical(d1, d1);
  H2cSuite<C>::F_2_192(f);
  F::from_canonical(f, f);
  F::mul(d0, d0, f);
  F::add(r, d0, d1);
}

}  // namespace ecg

// hash_from_bytes (NU = false) / encode_from_bytes (NU = true) for a batch of messages: message i is
// msgs[offsets[i] .. offsets[i+1]); the result is left as Jacobian SoA for normalize_kernel.
template <class C, bool NU>
ECG_KERNEL(128)
    h2c_kernel(const uint8_t* __restrict__ msgs, const uint64_t* __restrict__ offsets, uint64_t msgs_base, size_t n,
               const uint8_t* __restrict__ dst_prime, uint32_t dst_prime_len, uint32_t* __restrict__ jac) {
  typedef typename C::F F;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const uint8_t* m = msgs + (offsets[idx] - msgs_base);
  const size_t mlen = (size_t)(offsets[idx + 1] - offsets[idx]);
  constexpr int ELL = NU ? 2 : 3;
  uint32_t uniform[8 * ELL];
  ecg::expand_message_xmd_sha256<ELL>(uniform, m, mlen, dst_prime, dst_prime_len, NU ? 48u : 96u);
  ecg::Fe u0;
  ecg::h2c_field_from_okm<C>(u0, uniform);
  ecg::Jac q0;
  ecg::h2c_map_to_curve<C>(q0, u0);
  if (!NU) {
    ecg::Fe u1;
    ecg::Jac q1;
    ecg::h2c_field_from_okm<C>(u1, uniform + 12);
    ecg::h2c_map_to_curve<C>(q1, u1);
    ecg::jac_add<F, C::A_IS_MINUS3>(q0, q0, q1);  // both curves have cofactor 1: clear_cofactor is the identity map
  }
  soa_store<8>(jac, n, idx, q0.X.v, 0);
  soa_store<8>(jac, n, idx, q0.Y.v, 8);
  soa_store<8>(jac, n, idx, q0.Z.v, 16);
}

// hash_to_scalar (hash2curve/src/group_digest.rs:131-143 with L = 48: Reduce<Array<u8, U48>> for Scalar,
// k256/src/arithmetic/hash2curve.rs:151-166): out[i] = (d0 * 2^192 + d1) mod n as 32 big-endian bytes
template <class C>
ECG_KERNEL(128)
    h2s_kernel(const uint8_t* __restrict__ msgs, const uint64_t* __restrict__ offsets, uint64_t msgs_base, size_t n,
               const uint8_t* __restrict__ dst_prime, uint32_t dst_prime_len, uint8_t* __restrict__ out) {
  typedef ecg::FnMont<C> N;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const uint8_t* m = msgs + (offsets[idx] - msgs_base);
  const size_t mlen = (size_t)(offsets[idx + 1] - offsets[idx]);
  uint32_t w[16];
  ecg::expand_message_xmd_sha256<2>(w, m, mlen, dst_prime, dst_prime_len, 48u);
  uint32_t d0[8], d1[8], f[8], t[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    d0[i] = i < 6 ? w[5 - i] : 0u;
    d1[i] = i < 6 ? w[11 - i] : 0u;
    f[i] = i == 6 ? 1u : 0u;  // 2^192
  }
  N::to_mont(t, d0);
  N::mul(t, t, f);  // Montgomery form times plain value = plain product d0 * 2^192 mod n
  uint32_t c = ecg::add8(t, t, d1);
  N::cond_sub_n(t, c != 0 || N::ge_n(t));
  ecg::store_be32(out + 32 * idx, t);
}
