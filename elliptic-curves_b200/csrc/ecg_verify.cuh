// ecg_verify.cuh — signature-verification front end on top of a*G + b*P.
//
//   BIP340 / Schnorr (secp256k1):  k256/src/schnorr/verifying.rs:76-99  (verify_raw):
//        e = tagged_hash("BIP0340/challenge", r || pk || m) mod n;  R = s*G + (-e)*P;  accept iff
//        R != O, y(R) even, x(R) == r.     P = lift_x(pk) (VerifyingKey::from_bytes, verifying.rs:36-52)
//   ECDSA (k256, p256): the verification primitive lives in the external `ecdsa` crate the curve crates re-export
//        (k256/src/ecdsa.rs:93-121, p256/src/ecdsa.rs) — the standard SEC1 4.1.4 algorithm:
//        w = s^-1 mod n; u1 = z*w; u2 = r*w; R = u1*G + u2*Q; accept iff R != O and x(R) mod n == r.
//
// This header holds the per-element pieces: SHA-256 compression (for the BIP340 challenge), the field square root
// (for lift_x), Montgomery arithmetic mod n (for w, u1, u2).  The heavy lifting stays in mul_gen_add_kernel.
#pragma once
#include "ecg_curves.cuh"

namespace ecg {

// ---- SHA-256 compression -------------------------------------------------------------------------------------
ECG_D uint32_t sha_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

ECG_D void sha256_compress(uint32_t* st, const uint32_t* block) {
  const uint32_t K[64] = {
      0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, 0xd807aa98u, 0x12835b01u,
      0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, 0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu,
      0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, 0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u,
      0x06ca6351u, 0x14292967u, 0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
      0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u, 0x19a4c116u, 0x1e376c08u,
      0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u, 0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u,
      0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; i++) w[i] = block[i];
  uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll 1
  for (int i = 0; i < 64; i++) {
    uint32_t wi;
    if (i < 16) {
      wi = w[i & 15];
    } else {
      uint32_t w15 = w[(i - 15) & 15], w2 = w[(i - 2) & 15];
      uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
      uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
      wi = w[i & 15] + s0 + w[(i - 7) & 15] + s1;
      w[i & 15] = wi;
    }
    uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);
    uint32_t t1 = h + S1 + ch + K[i] + wi;
    uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + mj;
    h = g;
    g = f;
    f = e;
    e = d + t1;
    d = c;
    c = b;
    b = a;
    a = t1 + t2;
  }
  st[0] += a;
  st[1] += b;
  st[2] += c;
  st[3] += d;
  st[4] += e;
  st[5] += f;
  st[6] += g;
  st[7] += h;
}

// e = SHA256(SHA256(tag) || SHA256(tag) || r || pk || m) as 8 LE limbs (big-endian digest -> integer), tag =
// "BIP0340/challenge" (k256/src/schnorr.rs:85,221-227), 32-byte message.  The first block only depends on the tag:
// its midstate is a constant.
ECG_D void bip340_challenge(uint32_t* e, const uint8_t* r32, const uint8_t* pk32, const uint8_t* m32) {
  uint32_t st[8] = {0x9cecba11u, 0x23925381u, 0x11679112u, 0xd1627e0fu, 0x97c87550u, 0x003cc765u, 0x90f61164u, 0x33e9b66au};
  uint32_t blk[16];
  const uint32_t* rw = reinterpret_cast<const uint32_t*>(r32);
  const uint32_t* pw = reinterpret_cast<const uint32_t*>(pk32);
  const uint32_t* mw = reinterpret_cast<const uint32_t*>(m32);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    blk[i] = bswap32(rw[i]);
    blk[8 + i] = bswap32(pw[i]);
  }
  sha256_compress(st, blk);
#pragma unroll
  for (int i = 0; i < 8; i++) blk[i] = bswap32(mw[i]);
  blk[8] = 0x80000000u;
#pragma unroll
  for (int i = 9; i < 15; i++) blk[i] = 0;
  blk[15] = (64 + 96) * 8;
  sha256_compress(st, blk);
#pragma unroll
  for (int i = 0; i < 8; i++) e[i] = st[7 - i];
}

// ---- arithmetic mod n (Montgomery, R = 2^256) -------------------------------------------------------------------
template <class C>
struct FnMont {
  // a >= n ?
  ECG_D static bool ge_n(const uint32_t* a) { return !lt8(a, C::N()); }
  ECG_D static void cond_sub_n(uint32_t* a, bool doit) {
    uint32_t t[8];
    sub8(t, a, C::N());
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = doit ? t[i] : a[i];
  }
  // r = a*b*R^-1 mod n   (a, b < n)
  ECG_D static void mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t t[17];
    mul8x8(t, a, b);
    t[16] = 0;
    const uint32_t* n = C::N();
#pragma unroll 1
    for (int i = 0; i < 8; i++) {
      uint32_t m = t[i] * C::N_PRIME;
      uint32_t c = 0;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        uint64_t v = (uint64_t)m * n[j] + t[i + j] + c;
        t[i + j] = (uint32_t)v;
        c = (uint32_t)(v >> 32);
      }
#pragma unroll 1
      for (int k = i + 8; k < 17; k++) {
        uint64_t v = (uint64_t)t[k] + c;
        t[k] = (uint32_t)v;
        c = (uint32_t)(v >> 32);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = t[8 + i];
    cond_sub_n(r, t[16] != 0 || ge_n(r));
  }
  ECG_D static void to_mont(uint32_t* r, const uint32_t* a) { mul(r, a, C::N_R2()); }
  ECG_D static void from_mont(uint32_t* r, const uint32_t* a) {
    uint32_t one[8] = {1, 0, 0, 0, 0, 0, 0, 0};
    mul(r, a, one);
  }
  // r = a^(n-2) (Montgomery domain in and out); plain square-and-multiply over the bits of n-2 (public exponent)
  ECG_D static void inv(uint32_t* r, const uint32_t* a) {
    uint32_t e[8], acc[8];
    const uint32_t* n = C::N();
    // e = n - 2 (n is odd and > 2: no borrow beyond limb 0)
#pragma unroll
    for (int i = 0; i < 8; i++) e[i] = n[i];
    e[0] -= 2;
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = C::N_ONE()[i];
#pragma unroll 1
    for (int bit = 255; bit >= 0; bit--) {
      mul(acc, acc, acc);
      if ((e[bit >> 5] >> (bit & 31)) & 1u) mul(acc, acc, a);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = acc[i];
  }
  ECG_D static bool is_zero(const uint32_t* a) { return (a[0] | a[1] | a[2] | a[3] | a[4] | a[5] | a[6] | a[7]) == 0; }
};

// ---- secp256k1 square root: a^((p+1)/4)  (FieldElement::sqrt, k256/src/arithmetic/field.rs:200-235) ---------------
template <class F>
ECG_D void k256_sqrt_candidate(Fe& r, const Fe& a) {
  Fe x2, x3, x6, x9, x11, x22, x44, x88, x176, x220, x223, t;
  F::sqr(x2, a);
  F::mul(x2, x2, a);
  F::sqr(x3, x2);
  F::mul(x3, x3, a);
  F::sqr_n(x6, x3, 3);
  F::mul(x6, x6, x3);
  F::sqr_n(x9, x6, 3);
  F::mul(x9, x9, x3);
  F::sqr_n(x11, x9, 2);
  F::mul(x11, x11, x2);
  F::sqr_n(x22, x11, 11);
  F::mul(x22, x22, x11);
  F::sqr_n(x44, x22, 22);
  F::mul(x44, x44, x22);
  F::sqr_n(x88, x44, 44);
  F::mul(x88, x88, x44);
  F::sqr_n(x176, x88, 88);
  F::mul(x176, x176, x88);
  F::sqr_n(x220, x176, 44);
  F::mul(x220, x220, x44);
  F::sqr_n(x223, x220, 3);
  F::mul(x223, x223, x3);
  F::sqr_n(t, x223, 23);
  F::mul(t, t, x22);
  F::sqr_n(t, t, 6);
  F::mul(t, t, x2);
  F::sqr_n(r, t, 2);
}

// NIST P-256 square root: a^((p+1)/4), (p+1)/4 = 2^254 - 2^222 + 2^190 + 2^94
// (FieldElement::sqrt, p256/src/arithmetic/field.rs:121-147)
template <class F>
ECG_D void p256_sqrt_candidate(Fe& r, const Fe& a) {
  Fe x2, x4, x8, x16, x32, t;
  F::sqr(x2, a);
  F::mul(x2, x2, a);
  F::sqr_n(x4, x2, 2);
  F::mul(x4, x4, x2);
  F::sqr_n(x8, x4, 4);
  F::mul(x8, x8, x4);
  F::sqr_n(x16, x8, 8);
  F::mul(x16, x16, x8);
  F::sqr_n(x32, x16, 16);
  F::mul(x32, x32, x16);
  F::sqr_n(t, x32, 32);
  F::mul(t, t, a);
  F::sqr_n(t, t, 96);
  F::mul(t, t, a);
  F::sqr_n(r, t, 94);
}

// SEC1 decompression (AffinePoint::decompress, primeorder/src/affine.rs:179-198; k256/src/arithmetic/affine.rs
// DecompressPoint): y = sqrt(x^3 + a x + b) with the requested parity, or false (x >= p / no square root).
template <class C>
ECG_D bool sec1_decompress(Aff& P, const uint32_t* x_le, uint32_t y_is_odd) {
  typedef typename C::F F;
  if (!lt8(x_le, C::P())) return false;
  Fe xc, x, rhs, t, b, y, chk;
#pragma unroll
  for (int i = 0; i < 8; i++) xc.v[i] = x_le[i];
  F::from_canonical(x, xc);
  F::sqr(rhs, x);
  F::mul(rhs, rhs, x);
  if (C::A_IS_MINUS3) {
    F::mul_small(t, x, 3);
    F::sub(rhs, rhs, t);
  }
  C::b_internal(b);
  F::add(rhs, rhs, b);
  if (C::A_IS_MINUS3)
    p256_sqrt_candidate<F>(y, rhs);
  else
    k256_sqrt_candidate<F>(y, rhs);
  F::sqr(chk, y);
  F::sub(chk, chk, rhs);
  if (!F::is_zero(chk)) return false;
  Fe yc;
  F::to_canonical(yc, y);
  if ((yc.v[0] & 1u) != (y_is_odd & 1u)) F::neg(y, y);
  P.x = x;
  P.y = y;
  return true;
}

// lift_x: the point with the given x and even y, or false (x >= p or x^3 + 7 not a square).
template <class F>
ECG_D bool k256_lift_x(Aff& P, const uint32_t* x_le) {
  if (!lt8(x_le, K256_P)) return false;
  Fe x, rhs, y, chk, seven;
#pragma unroll
  for (int i = 0; i < 8; i++) x.v[i] = x_le[i];
  F::sqr(rhs, x);
  F::mul(rhs, rhs, x);
  F::set_zero(seven);
  seven.v[0] = 7;
  F::add(rhs, rhs, seven);
  k256_sqrt_candidate<F>(y, rhs);
  F::sqr(chk, y);
  F::sub(chk, chk, rhs);
  if (!F::is_zero(chk)) return false;
  F::normalize(y, y);
  if (y.v[0] & 1u) {
    F::neg(y, y);
    F::normalize(y, y);
  }
  P.x = x;
  P.y = y;
  return true;
}

}  // namespace ecg
