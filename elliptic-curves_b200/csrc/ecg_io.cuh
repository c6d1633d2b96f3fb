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
ECG_D void load_be32(uint32_t* limbs, const uint8_t* p) { load_be<8>(limbs, p); }
ECG_D void store_be32(uint8_t* p, const uint32_t* limbs) { store_be<8>(p, limbs); }

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
