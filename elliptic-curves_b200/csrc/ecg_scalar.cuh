// ecg_scalar.cuh — scalar-side preparation: range check, GLV split (secp256k1), signed-odd recoding.
//
// Replaces glv::decompose_scalar (k256/src/arithmetic/mul/glv.rs:149-156) with WideScalar::mul_shift_vartime
// (k256/src/arithmetic/scalar/wide64.rs:64-119) and the sign folding of k256/src/arithmetic/mul.rs:120-132;
// and replaces Radix16Decomposition (primeorder/src/tables/radix16.rs:31-55) / wnaf_form
// (wnaf/src/lib.rs:70-150) with a *signed odd* fixed-window recoding: every digit is in
// {+-1, +-3, ..., +-15}, so every window performs exactly one mixed addition (no zero digits, no
// lane divergence) against a table of the 8 odd multiples.
//
// The split is computed in plain integers instead of mod n:
//     c1 = round(k*g1 / 2^384),  c2 = round(k*g2 / 2^384)            (same g1, g2 as glv.rs:28-37)
//     k1 = k - c1*a1 - c2*a2,    k2 = c1*(-b1) - c2*b2               (mul.rs:7-35 notation)
// which are the reference's (r1, r2) lifted to (-2^128, 2^128) (glv.rs:43-146 proves the bound), so
// low 160 bits in two's complement carry them exactly.
#pragma once
#include "ecg_prim.cuh"

namespace ecg {

// r[0..NA+NB) = a * b, plain row-wise schoolbook on 64-bit temporaries (not on the hot path).
template <int NA, int NB>
ECG_D void mul_limbs(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#pragma unroll
  for (int i = 0; i < NA + NB; i++) r[i] = 0;
#pragma unroll
  for (int i = 0; i < NB; i++) {
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < NA; j++) {
      uint64_t t = (uint64_t)a[j] * b[i] + r[i + j] + c;
      r[i + j] = (uint32_t)t;
      c = (uint32_t)(t >> 32);
    }
    r[i + NA] = c;
  }
}

// a < m ? (N limbs)
template <int N>
ECG_D bool ltN(const uint32_t* a, const uint32_t* m) {
  uint32_t t = sub_cc(a[0], m[0]);
#pragma unroll
  for (int i = 1; i < N; i++) t = subc_cc(a[i], m[i]);
  (void)t;
  return subc(0, 0) != 0;  // borrow => a < m
}
ECG_D bool lt8(const uint32_t* a, const uint32_t* m) { return ltN<8>(a, m); }

struct GlvHalf {
  uint32_t h[4];   // (|k_i| made odd) >> 1 : 32 four-bit windows, MSB-first consumption
  uint32_t neg;    // 1 if k_i < 0 (fold the sign into the point, mul.rs:120-132)
  uint32_t even;   // 1 if |k_i| was even: the loop computes (|k_i|+1)*P, subtract P afterwards
};

// two's-complement 160-bit -> (magnitude, sign); returns false if |v| >= 2^129 (cannot happen for k < n)
// CT: the negation is computed for every input and selected by mask (ECG_FLAG_CONSTTIME kernels)
template <bool CT = false>
ECG_D bool glv_finish(GlvHalf& o, uint32_t* v /*5 limbs*/) {
  uint32_t s = v[4] >> 31;
  if (CT) {  // v = (v ^ m) + s with m = -s: two's-complement negation when s = 1, identity when s = 0
    const uint32_t m = 0u - s;
    v[0] = add_cc(v[0] ^ m, s);
#pragma unroll
    for (int i = 1; i < 4; i++) v[i] = addc_cc(v[i] ^ m, 0);
    v[4] = addc(v[4] ^ m, 0);
  } else if (s) {
    v[0] = sub_cc(0, v[0]);
#pragma unroll
    for (int i = 1; i < 4; i++) v[i] = subc_cc(0, v[i]);
    v[4] = subc(0, v[4]);
  }
  o.neg = s;
  o.even = (~v[0]) & 1u;
  // m = |v| + even  (odd), h = m >> 1
  v[0] = add_cc(v[0], o.even);
#pragma unroll
  for (int i = 1; i < 4; i++) v[i] = addc_cc(v[i], 0);
  v[4] = addc(v[4], 0);
  bool ok = v[4] <= 1u;
#pragma unroll
  for (int i = 0; i < 4; i++) o.h[i] = funnel_r(v[i], v[i + 1], 1);
  return ok;
}

// secp256k1 GLV split of k (8 LE limbs, k < n).
template <bool CT = false>
ECG_D bool glv_split_k256(GlvHalf& h1, GlvHalf& h2, const uint32_t* k) {
  const uint32_t G1[8] = {0x45DBB031u, 0xE893209Au, 0x71E8CA7Fu, 0x3DAA8A14u, 0x9284EB15u, 0xE86C90E4u, 0xA7D46BCDu, 0x3086D221u};
  const uint32_t G2[8] = {0x8AC47F71u, 0x1571B4AEu, 0x9DF506C6u, 0x221208ACu, 0x0ABFE4C4u, 0x6F547FA9u, 0x010E8828u, 0xE4437ED6u};
  const uint32_t A1[4] = {0x9284EB15u, 0xE86C90E4u, 0xA7D46BCDu, 0x3086D221u};                // a1 = b2
  const uint32_t MB1[4] = {0x0ABFE4C3u, 0x6F547FA9u, 0x010E8828u, 0xE4437ED6u};               // -b1
  const uint32_t A2[5] = {0x9D44CFD8u, 0x57C1108Du, 0xA8E2F3F6u, 0x14CA50F7u, 0x00000001u};   // a2
  uint32_t t[16], c1[4], c2[4];
  mul8x8(t, k, G1);
  {
    uint32_t rb = t[11] >> 31;
    c1[0] = add_cc(t[12], rb);
    c1[1] = addc_cc(t[13], 0);
    c1[2] = addc_cc(t[14], 0);
    c1[3] = addc(t[15], 0);
  }
  mul8x8(t, k, G2);
  {
    uint32_t rb = t[11] >> 31;
    c2[0] = add_cc(t[12], rb);
    c2[1] = addc_cc(t[13], 0);
    c2[2] = addc_cc(t[14], 0);
    c2[3] = addc(t[15], 0);
  }
  uint32_t p[9], q[9], v[5];
  // k2 = c1*(-b1) - c2*b2   (mod 2^160)
  mul_limbs<4, 4>(p, c1, MB1);
  mul_limbs<4, 4>(q, c2, A1);
  v[0] = sub_cc(p[0], q[0]);
#pragma unroll
  for (int i = 1; i < 4; i++) v[i] = subc_cc(p[i], q[i]);
  v[4] = subc(p[4], q[4]);
  bool ok2 = glv_finish<CT>(h2, v);
  // k1 = k - c1*a1 - c2*a2   (mod 2^160)
  mul_limbs<4, 4>(p, c1, A1);
  mul_limbs<5, 4>(q, A2, c2);
  v[0] = sub_cc(k[0], p[0]);
#pragma unroll
  for (int i = 1; i < 4; i++) v[i] = subc_cc(k[i], p[i]);
  v[4] = subc(k[4], p[4]);
  v[0] = sub_cc(v[0], q[0]);
#pragma unroll
  for (int i = 1; i < 4; i++) v[i] = subc_cc(v[i], q[i]);
  v[4] = subc(v[4], q[4]);
  bool ok1 = glv_finish<CT>(h1, v);
  return ok1 & ok2;
}

// Window consumption, most significant first: returns the 4-bit window n and shifts h left by 4.
// digit d = 2n - 15:  n >= 8 -> +(2(n-8)+1), table index n-8;  n < 8 -> -(2(7-n)+1), index 7-n.
ECG_D uint32_t next_window(uint32_t* h) {
  uint32_t n = h[3] >> 28;
  h[3] = (h[3] << 4) | (h[2] >> 28);
  h[2] = (h[2] << 4) | (h[1] >> 28);
  h[1] = (h[1] << 4) | (h[0] >> 28);
  h[0] = h[0] << 4;
  return n;
}

// 256-bit scalar, no endomorphism (P-256): k in [0, n) -> odd m = k or k+1 (k+1 <= n-1+1 < 2^256),
// windows of h = m >> 1 (255 bits -> 64 windows, top window < 8... see kernels), parity flag.
// (NL limbs: 8 for the 256-bit curves, 12 for P-384, whose order is also just below 2^384)
template <int NL>
struct FullRecodeN {
  uint32_t h[NL];
  uint32_t even;
};
typedef FullRecodeN<8> FullRecode;
template <int NL>
ECG_D void recode_full(FullRecodeN<NL>& o, const uint32_t* k) {
  uint32_t v[NL];
  o.even = (~k[0]) & 1u;
  v[0] = add_cc(k[0], o.even);
#pragma unroll
  for (int i = 1; i < NL; i++) v[i] = addc_cc(k[i], 0);
#pragma unroll
  for (int i = 0; i < NL - 1; i++) o.h[i] = funnel_r(v[i], v[i + 1], 1);
  o.h[NL - 1] = v[NL - 1] >> 1;
}
template <int NL>
ECG_D uint32_t next_windowN(uint32_t* h) {
  uint32_t n = h[NL - 1] >> 28;
#pragma unroll
  for (int i = NL - 1; i > 0; i--) h[i] = (h[i] << 4) | (h[i - 1] >> 28);
  h[0] = h[0] << 4;
  return n;
}
ECG_D uint32_t next_window8(uint32_t* h) { return next_windowN<8>(h); }

}  // namespace ecg
