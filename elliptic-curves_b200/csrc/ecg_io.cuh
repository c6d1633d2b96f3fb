// ecg_io.cuh — canonical 32-byte big-endian <-> 8 little-endian 32-bit limbs, and Jacobian -> affine.
// Encoding contract: Scalar::to_bytes (k256/src/arithmetic/scalar.rs:96-98), FieldElement::to_bytes
// (k256/src/arithmetic/field.rs:110-112), p256 ByteOrder::BigEndian (p256/src/arithmetic/field.rs:41).
#pragma once
#include "ecg_prim.cuh"
#include "ecg_point.cuh"

namespace ecg {

// p must be 4-byte aligned (all library buffers are 256-byte aligned device allocations)
template <int NL>
ECG_D void load_be(uint32_t* limbs, const uint8_t* p) {  // 4*NL big-endian bytes -> NL little-endian limbs
  const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
  for (int i = 0; i < NL; i++) limbs[NL - 1 - i] = bswap32(w[i]);
}
template <int NL>
ECG_D void store_be(uint8_t* p, const uint32_t* limbs) {
  uint32_t* w = reinterpret_cast<uint32_t*>(p);
#pragma unroll
  for (int i = 0; i < NL; i++) w[i] = bswap32(limbs[NL - 1 - i]);
}
// little-endian records (bign-curve256v1: primefield ByteOrder::LittleEndian, bignp256/src/arithmetic/field.rs:65)
template <int NL>
ECG_D void load_le(uint32_t* limbs, const uint8_t* p) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
  for (int i = 0; i < NL; i++) limbs[i] = w[i];
}
template <int NL>
ECG_D void store_le(uint8_t* p, const uint32_t* limbs) {
  uint32_t* w = reinterpret_cast<uint32_t*>(p);
#pragma unroll
  for (int i = 0; i < NL; i++) w[i] = limbs[i];
}
// one canonical record (field element or scalar) of the curve whose field policy is F, in the byte order the
// reference uses for that curve
// records whose length is not 4 NL (P-521: 66 big-endian bytes in 17 limbs) are not word-aligned: byte accesses
template <int NL, int FB>
ECG_D void load_be_bytes(uint32_t* limbs, const uint8_t* p) {
#pragma unroll
  for (int i = 0; i < NL; i++) {
    uint32_t v = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int pos = FB - 1 - (4 * i + j);  // byte of weight 256^(4 i + j)
      if (pos >= 0) v |= (uint32_t)p[pos] << (8 * j);
    }
    limbs[i] = v;
  }
}
template <int NL, int FB>
ECG_D void store_be_bytes(uint8_t* p, const uint32_t* limbs) {
#pragma unroll
  for (int i = 0; i < NL; i++) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int pos = FB - 1 - (4 * i + j);
      if (pos >= 0) p[pos] = (uint8_t)(limbs[i] >> (8 * j));
    }
  }
}
template <class F>
ECG_D void load_fe(uint32_t* limbs, const uint8_t* p) {
  if (F::FB != 4 * F::NL)
    load_be_bytes<F::NL, F::FB>(limbs, p);
  else if (F::LE)
    load_le<F::NL>(limbs, p);
  else
    load_be<F::NL>(limbs, p);
}
template <class F>
ECG_D void store_fe(uint8_t* p, const uint32_t* limbs) {
  if (F::FB != 4 * F::NL)
    store_be_bytes<F::NL, F::FB>(p, limbs);
  else if (F::LE)
    store_le<F::NL>(p, limbs);
  else
    store_be<F::NL>(p, limbs);
}
ECG_D void load_be32(uint32_t* limbs, const uint8_t* p) { load_be<8>(limbs, p); }
ECG_D void store_be32(uint8_t* p, const uint32_t* limbs) { store_be<8>(p, limbs); }

// one table entry = x[NL], y[NL] (internal form), read as 128-bit loads (64-bit loads when 2 NL words are not a
// multiple of four: the 7-limb field of P-224)
template <int NL>
ECG_DEV void load_aff_entry(AffN<NL>& e, const uint32_t* __restrict__ table, size_t point) {
  uint32_t w[2 * NL];
  if ((2 * NL) % 4 == 0) {
    const uint4* p = reinterpret_cast<const uint4*>(table + point * (2 * NL));
#pragma unroll
    for (int q = 0; q < NL / 2; q++) {
      uint4 v = __ldg(p + q);
      w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
    }
  } else {
    const uint2* p = reinterpret_cast<const uint2*>(table + point * (2 * NL));
#pragma unroll
    for (int q = 0; q < NL; q++) {
      uint2 v = __ldg(p + q);
      w[2 * q] = v.x; w[2 * q + 1] = v.y;
    }
  }
#pragma unroll
  for (int i = 0; i < NL; i++) {
    e.x.v[i] = w[i];
    e.y.v[i] = w[NL + i];
  }
}
// (X:Y:Z) with zinv = 1/Z  ->  canonical affine integers
template <class F>
ECG_D void jac_to_affine_canonical(typename F::FeT& x, typename F::FeT& y, const typename F::JacT& p, const typename F::FeT& zinv) {
  typename F::FeT z2, z3;
  F::sqr(z2, zinv);
  F::mul(z3, z2, zinv);
  F::mul(x, p.X, z2);
  F::mul(y, p.Y, z3);
  F::to_canonical(x, x);
  F::to_canonical(y, y);
}

}  // namespace ecg
