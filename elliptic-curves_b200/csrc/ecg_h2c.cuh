// ecg_h2c.cuh — hash to curve (RFC 9380) for the four Weierstrass suites the reference implements:
//   secp256k1_XMD:SHA-256_SSWU_{RO,NU}_  (k256/src/arithmetic/hash2curve.rs:14-20,52-148: simplified SWU on the 3-isogenous
//                                          curve E', then the isogeny map :169-258)
//   P256_XMD:SHA-256_SSWU_{RO,NU}_       (p256/src/arithmetic/hash2curve.rs:13-75 over primeorder/src/osswu.rs:60-146)
//   P384_XMD:SHA-384_SSWU_{RO,NU}_       (p384/src/arithmetic/hash2curve.rs:13-75, L = 72)
//   P521_XMD:SHA-512_SSWU_{RO,NU}_       (p521/src/arithmetic/hash2curve.rs:13-76, L = 98)
// and the drivers hash2curve/src/group_digest.rs:88-143 (hash_from_bytes = two field elements, two maps, one addition;
// encode_from_bytes = one map; hash_to_scalar), hash2curve/src/hash2field.rs + hash2field/expand_msg/xmd.rs:43-99
// (hash_to_field over expand_message_xmd).  SURVEY.md section 8(f) rank 4 ("hash-to-curve front end").
//
// One thread per message: streaming SHA-2 over the message bytes (b_0), the chained blocks b_1..b_ell, the L-byte
// reductions (24-byte chunks folded with 2^192, which is d0 * 2^192 + d1 for L = 48), the straight-line SSWU with ONE exponentiation per map (sqrt_ratio for q = 3 mod 4) and
// no inversion: the map leaves x as a fraction, which becomes the Z of a Jacobian point (secp256k1: the isogeny is
// evaluated on the homogenised polynomials), the two points are added with the library's Jacobian addition and the
// batch is normalised by normalize_kernel (one inversion per ~32 points).
#pragma once
#include "ecg_curves.cuh"
#include "ecg_verify.cuh"
#include "ecg_h2c_consts.cuh"

namespace ecg {

// ---- streaming SHA-256 (byte-granular updates; compression from ecg_verify.cuh) --------------------------------
struct Sha256Stream {
  static constexpr int B = 32, S = 64;  // digest / block bytes
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
  ECG_D void zero_block() {  // Z_pad: one block of zero bytes
    flush_block();
    total += 64;
  }
  ECG_D void finish(uint8_t* digest /* B bytes */) {
    uint64_t bits = total * 8;
    put(0x80u);
    if (fill > 56) flush_block();
    blk[14] = (uint32_t)(bits >> 32);
    blk[15] = (uint32_t)bits;
    sha256_compress(st, blk);
#pragma unroll
    for (int i = 0; i < 32; i++) digest[i] = (uint8_t)(st[i >> 2] >> (24 - 8 * (i & 3)));
  }
};

// ---- SHA-512 / SHA-384 ------------------------------------------------------------------------------------------
ECG_D uint64_t sha_rotr64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
ECG_D void sha512_compress(uint64_t* st, const uint64_t* block) {
  const uint64_t K[80] = {
      0x428a2f98d728ae22ull, 0x7137449123ef65cdull, 0xb5c0fbcfec4d3b2full, 0xe9b5dba58189dbbcull, 0x3956c25bf348b538ull, 0x59f111f1b605d019ull,
      0x923f82a4af194f9bull, 0xab1c5ed5da6d8118ull, 0xd807aa98a3030242ull, 0x12835b0145706fbeull, 0x243185be4ee4b28cull, 0x550c7dc3d5ffb4e2ull,
      0x72be5d74f27b896full, 0x80deb1fe3b1696b1ull, 0x9bdc06a725c71235ull, 0xc19bf174cf692694ull, 0xe49b69c19ef14ad2ull, 0xefbe4786384f25e3ull,
      0x0fc19dc68b8cd5b5ull, 0x240ca1cc77ac9c65ull, 0x2de92c6f592b0275ull, 0x4a7484aa6ea6e483ull, 0x5cb0a9dcbd41fbd4ull, 0x76f988da831153b5ull,
      0x983e5152ee66dfabull, 0xa831c66d2db43210ull, 0xb00327c898fb213full, 0xbf597fc7beef0ee4ull, 0xc6e00bf33da88fc2ull, 0xd5a79147930aa725ull,
      0x06ca6351e003826full, 0x142929670a0e6e70ull, 0x27b70a8546d22ffcull, 0x2e1b21385c26c926ull, 0x4d2c6dfc5ac42aedull, 0x53380d139d95b3dfull,
      0x650a73548baf63deull, 0x766a0abb3c77b2a8ull, 0x81c2c92e47edaee6ull, 0x92722c851482353bull, 0xa2bfe8a14cf10364ull, 0xa81a664bbc423001ull,
      0xc24b8b70d0f89791ull, 0xc76c51a30654be30ull, 0xd192e819d6ef5218ull, 0xd69906245565a910ull, 0xf40e35855771202aull, 0x106aa07032bbd1b8ull,
      0x19a4c116b8d2d0c8ull, 0x1e376c085141ab53ull, 0x2748774cdf8eeb99ull, 0x34b0bcb5e19b48a8ull, 0x391c0cb3c5c95a63ull, 0x4ed8aa4ae3418acbull,
      0x5b9cca4f7763e373ull, 0x682e6ff3d6b2b8a3ull, 0x748f82ee5defb2fcull, 0x78a5636f43172f60ull, 0x84c87814a1f0ab72ull, 0x8cc702081a6439ecull,
      0x90befffa23631e28ull, 0xa4506cebde82bde9ull, 0xbef9a3f7b2c67915ull, 0xc67178f2e372532bull, 0xca273eceea26619cull, 0xd186b8c721c0c207ull,
      0xeada7dd6cde0eb1eull, 0xf57d4f7fee6ed178ull, 0x06f067aa72176fbaull, 0x0a637dc5a2c898a6ull, 0x113f9804bef90daeull, 0x1b710b35131c471bull,
      0x28db77f523047d84ull, 0x32caab7b40c72493ull, 0x3c9ebe0a15c9bebcull, 0x431d67c49c100d4cull, 0x4cc5d4becb3e42b6ull, 0x597f299cfc657e2aull,
      0x5fcb6fab3ad6faecull, 0x6c44198c4a475817ull};
  uint64_t w[16];
#pragma unroll
  for (int i = 0; i < 16; i++) w[i] = block[i];
  uint64_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll 1
  for (int i = 0; i < 80; i++) {
    uint64_t wi;
    if (i < 16) {
      wi = w[i & 15];
    } else {
      uint64_t w15 = w[(i - 15) & 15], w2 = w[(i - 2) & 15];
      uint64_t s0 = sha_rotr64(w15, 1) ^ sha_rotr64(w15, 8) ^ (w15 >> 7);
      uint64_t s1 = sha_rotr64(w2, 19) ^ sha_rotr64(w2, 61) ^ (w2 >> 6);
      wi = w[i & 15] + s0 + w[(i - 7) & 15] + s1;
      w[i & 15] = wi;
    }
    uint64_t S1 = sha_rotr64(e, 14) ^ sha_rotr64(e, 18) ^ sha_rotr64(e, 41);
    uint64_t ch = (e & f) ^ (~e & g);
    uint64_t t1 = h + S1 + ch + K[i] + wi;
    uint64_t S0 = sha_rotr64(a, 28) ^ sha_rotr64(a, 34) ^ sha_rotr64(a, 39);
    uint64_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint64_t t2 = S0 + mj;
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}
template <bool IS384>
struct Sha512Stream {
  static constexpr int B = IS384 ? 48 : 64, S = 128;
  uint64_t st[8];
  uint64_t blk[16];
  uint32_t fill;
  uint64_t total;
  ECG_D void init() {
    const uint64_t iv512[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                               0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
    const uint64_t iv384[8] = {0xcbbb9d5dc1059ed8ull, 0x629a292a367cd507ull, 0x9159015a3070dd17ull, 0x152fecd8f70e5939ull,
                               0x67332667ffc00b31ull, 0x8eb44a8768581511ull, 0xdb0c2e0d64f98fa7ull, 0x47b5481dbefa4fa4ull};
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = IS384 ? iv384[i] : iv512[i];
#pragma unroll
    for (int i = 0; i < 16; i++) blk[i] = 0;
    fill = 0;
    total = 0;
  }
  ECG_D void flush_block() {
    sha512_compress(st, blk);
#pragma unroll
    for (int i = 0; i < 16; i++) blk[i] = 0;
    fill = 0;
  }
  ECG_D void put(uint32_t byte) {
    blk[fill >> 3] |= (uint64_t)(byte & 0xFFu) << (56 - 8 * (fill & 7));
    fill++;
    total++;
    if (fill == 128) flush_block();
  }
  ECG_D void update(const uint8_t* p, size_t len) {
    for (size_t i = 0; i < len; i++) put(p[i]);
  }
  ECG_D void zero_block() {
    flush_block();
    total += 128;
  }
  ECG_D void finish(uint8_t* digest /* B bytes */) {
    uint64_t bits = total * 8;
    put(0x80u);
    if (fill > 112) flush_block();
    blk[15] = bits;  // the 128-bit length: the high word stays zero
    sha512_compress(st, blk);
#pragma unroll
    for (int i = 0; i < B; i++) digest[i] = (uint8_t)(st[i >> 3] >> (56 - 8 * (i & 7)));
  }
};
template <int HASH>
struct H2cHash;
template <>
struct H2cHash<256> {
  typedef Sha256Stream T;
};
template <>
struct H2cHash<384> {
  typedef Sha512Stream<true> T;
};
template <>
struct H2cHash<512> {
  typedef Sha512Stream<false> T;
};

// uniform[0 .. B*ELL) = expand_message_xmd(msg, DST, len_in_bytes) (RFC 9380 section 5.3.1,
// hash2curve/src/hash2field/expand_msg/xmd.rs:43-99).  dst_prime = DST || I2OSP(len(DST), 1) (an oversize DST is
// replaced by its hash on the host: expand_msg.rs:76-95).  HS: the streaming hash (digest B bytes, block S bytes).
template <class HS, int ELL>
ECG_D void expand_message_xmd(uint8_t* uniform, const uint8_t* msg, size_t msg_len, const uint8_t* dst_prime, uint32_t dst_prime_len,
                              uint32_t len_in_bytes) {
  HS h;
  uint8_t b0[HS::B], bi[HS::B];
  h.init();
  h.zero_block();
  h.update(msg, msg_len);
  h.put(len_in_bytes >> 8);
  h.put(len_in_bytes);
  h.put(0);
  h.update(dst_prime, dst_prime_len);
  h.finish(b0);
#pragma unroll 1
  for (int i = 1; i <= ELL; i++) {
    h.init();
#pragma unroll 1
    for (int j = 0; j < HS::B; j++) h.put(i == 1 ? b0[j] : (uint32_t)(b0[j] ^ bi[j]));
    h.put((uint32_t)i);
    h.update(dst_prime, dst_prime_len);
    h.finish(bi);
#pragma unroll 1
    for (int j = 0; j < HS::B; j++) uniform[HS::B * (i - 1) + j] = bi[j];
  }
}



// r = a^((p-3)/4), fixed 4-bit windows over the public exponent (32 NL - 4 squarings + 8 NL + 14 multiplications)
template <class C>
ECG_D void h2c_pow_c1(typename C::F::FeT& r, const typename C::F::FeT& a) {
  typedef typename C::F F;
  typename F::FeT tab[16];
  F::set_one(tab[0]);
  tab[1] = a;
#pragma unroll 1
  for (int i = 2; i < 16; i++) F::mul(tab[i], tab[i - 1], a);
  typename F::FeT acc;
  F::set_one(acc);
#pragma unroll 1
  for (int w = 8 * F::NL - 1; w >= 0; w--) {
    F::sqr_n(acc, acc, 4);
    uint32_t d = (H2cSuite<C>::C1(w >> 3) >> (4 * (w & 7))) & 15u;
    F::mul(acc, acc, tab[d]);
  }
  r = acc;
}

// sqrt_ratio for q = 3 (mod 4): primeorder/src/osswu.rs:60-88 (RFC 9380 F.2.1.2).  Returns is_square(u / v); y = sqrt(u / v) if
// it is one, sqrt(Z u / v) otherwise.
template <class C>
ECG_D bool h2c_sqrt_ratio(typename C::F::FeT& y, const typename C::F::FeT& u, const typename C::F::FeT& v, const typename C::F::FeT& c2) {
  typedef typename C::F F;
  typename F::FeT tv1, tv2, tv3, y1, y2;
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
  for (int i = 0; i < F::NL; i++) y.v[i] = is_qr ? y1.v[i] : y2.v[i];
  return is_qr;
}

// parity of the canonical representative (Sgn0: k256/src/arithmetic/hash2curve.rs:46-50, p256/.../hash2curve.rs:38-42)
template <class C>
ECG_D uint32_t h2c_sgn0(const typename C::F::FeT& a) {
  typename C::F::FeT t;
  C::F::to_canonical(t, a);
  return t.v[0] & 1u;
}

// simplified SWU, straight line (primeorder/src/osswu.rs:92-146, RFC 9380 F.2): the point (xn / xd, y) on the curve
// y^2 = x^3 + A x + B of the suite (secp256k1: the isogenous curve E').  u in internal form.
template <class C>
ECG_D void h2c_sswu(typename C::F::FeT& xn, typename C::F::FeT& xd, typename C::F::FeT& y, const typename C::F::FeT& u) {
  typedef typename C::F F;
  typedef typename F::FeT Fe;
  typedef H2cSuite<C> S;
  constexpr int NL = F::NL;
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
  for (int i = 0; i < NL; i++) tv4.v[i] = z2 ? Z.v[i] : t.v[i];
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
  for (int i = 0; i < NL; i++) {  // 21, 22
    x.v[i] = is_sq ? tv3.v[i] : x.v[i];
    y.v[i] = is_sq ? y1.v[i] : y.v[i];
  }
  uint32_t flip = h2c_sgn0<C>(u) ^ h2c_sgn0<C>(y);  // 23, 24
  fe_cneg<F>(y, flip);
  xn = x;                  // 25: x = x / tv4, left as a fraction
  xd = tv4;
}

// map_to_curve as a Jacobian point of the target curve (no inversion).
//   P-256 / P-384 / P-521: (xn / xd, y) is the point: (X : Y : Z) = (xn xd : y xd^3 : xd).
//   secp256k1: the 3-isogeny E' -> E (k256/src/arithmetic/hash2curve.rs:169-258, RFC 9380 E.1) on x' = N / D:
//     x = XN / (D XD),  y = y' YN / YD  with the homogenised polynomials XN = sum k_i N^i D^(3-i), XD = N^2 + ..., so
//     (X : Y : Z) = (XN D XD YD^2 : y' YN D^3 XD^3 YD^2 : D XD YD); a vanishing denominator gives Z = 0, the identity
//     (RFC 9380 section 6.6.3: exceptional cases of the isogeny map to the identity).
template <class C>
ECG_D void h2c_map_to_curve(typename C::F::JacT& r, const typename C::F::FeT& u) {
  typedef typename C::F F;
  typedef typename F::FeT Fe;
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

// L uniform bytes (big-endian integer) -> its residue in the field policy FF (internal form): 24-byte chunks, most
// significant first, folded with 2^192 — for L = 48 this is the reference's d0 * 2^192 + d1
// (Reduce<Array<u8, U48>> for FieldElement, k256/src/arithmetic/hash2curve.rs:22-44; U72: p384/.../hash2curve.rs:21-38; U98: p521)
template <class FF, int L>
ECG_D void h2c_from_okm(typename FF::FeT& r, const uint8_t* okm, const typename FF::FeT& f_2_192 /* internal form */) {
  typedef typename FF::FeT Fe;
  constexpr int NL = FF::NL;
  constexpr int FIRST = (L % 24) ? (L % 24) : 24;
  int pos = 0;
  FF::set_zero(r);
#pragma unroll 1
  for (int c = 0; pos < L; c++) {
    const int len = c == 0 ? FIRST : 24;
    Fe d;
#pragma unroll
    for (int i = 0; i < NL; i++) d.v[i] = 0;
#pragma unroll 1
    for (int j = 0; j < len; j++) {  // byte of weight 256^(len - 1 - j)
      const int wgt = len - 1 - j;
      d.v[wgt >> 2] |= (uint32_t)okm[pos + j] << (8 * (wgt & 3));
    }
    FF::from_canonical(d, d);
    if (c > 0) FF::mul(r, r, f_2_192);
    FF::add(r, r, d);
    pos += len;
  }
}

}  // namespace ecg

// hash_from_bytes (NU = false) / encode_from_bytes (NU = true) for a batch of messages: message i is
// msgs[offsets[i] .. offsets[i+1]); the result is left as Jacobian SoA for normalize_kernel.
template <class C, bool NU>
ECG_KERNEL(128)
    h2c_kernel(const uint8_t* __restrict__ msgs, const uint64_t* __restrict__ offsets, uint64_t msgs_base, size_t n,
               const uint8_t* __restrict__ dst_prime, uint32_t dst_prime_len, uint32_t* __restrict__ jac) {
  typedef typename C::F F;
  typedef ecg::H2cSuite<C> S;
  typedef typename ecg::H2cHash<S::HASH>::T HS;
  constexpr int NL = F::NL, L = S::L, COUNT = NU ? 1 : 2, ELL = (COUNT * L + HS::B - 1) / HS::B;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const uint8_t* m = msgs + (offsets[idx] - msgs_base);
  const size_t mlen = (size_t)(offsets[idx + 1] - offsets[idx]);
  uint8_t uniform[ELL * HS::B];
  ecg::expand_message_xmd<HS, ELL>(uniform, m, mlen, dst_prime, dst_prime_len, (uint32_t)(COUNT * L));
  typename F::FeT f192, u0;
  S::F_2_192(f192);
  F::from_canonical(f192, f192);
  ecg::h2c_from_okm<F, L>(u0, uniform, f192);
  typename F::JacT q0;
  ecg::h2c_map_to_curve<C>(q0, u0);
  if (!NU) {
    typename F::FeT u1;
    typename F::JacT q1;
    ecg::h2c_from_okm<F, L>(u1, uniform + L, f192);
    ecg::h2c_map_to_curve<C>(q1, u1);
    ecg::jac_add<F, C::A_IS_MINUS3>(q0, q0, q1);  // all four curves have cofactor 1: clear_cofactor is the identity map
  }
  soa_store<NL>(jac, n, idx, q0.X.v, 0);
  soa_store<NL>(jac, n, idx, q0.Y.v, NL);
  soa_store<NL>(jac, n, idx, q0.Z.v, 2 * NL);
}

// hash_to_scalar (hash2curve/src/group_digest.rs:131-143: hash_to_field with the group order as modulus, L as for the field;
// Reduce<Array<u8, U48 / U72 / U98>> for Scalar, k256/src/arithmetic/hash2curve.rs:151-166, p384 / p521 likewise):
// out[i] = OS2IP(uniform) mod n as one canonical FB-byte record.  FN: the scalar field as a Montgomery policy over n.
template <class C, class FN>
ECG_KERNEL(128)
    h2s_kernel(const uint8_t* __restrict__ msgs, const uint64_t* __restrict__ offsets, uint64_t msgs_base, size_t n,
               const uint8_t* __restrict__ dst_prime, uint32_t dst_prime_len, uint8_t* __restrict__ out) {
  typedef ecg::H2cSuite<C> S;
  typedef typename ecg::H2cHash<S::HASH>::T HS;
  constexpr int L = S::L, ELL = (L + HS::B - 1) / HS::B;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const uint8_t* m = msgs + (offsets[idx] - msgs_base);
  const size_t mlen = (size_t)(offsets[idx + 1] - offsets[idx]);
  uint8_t uniform[ELL * HS::B];
  ecg::expand_message_xmd<HS, ELL>(uniform, m, mlen, dst_prime, dst_prime_len, (uint32_t)L);
  typename FN::FeT f192, r;
#pragma unroll
  for (int i = 0; i < FN::NL; i++) f192.v[i] = i == 6 ? 1u : 0u;  // 2^192 < n for every curve here
  FN::from_canonical(f192, f192);
  ecg::h2c_from_okm<FN, L>(r, uniform, f192);
  FN::to_canonical(r, r);
  ecg::store_fe<FN>(out + FN::FB * idx, r.v);
}
