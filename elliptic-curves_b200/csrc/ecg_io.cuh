// ecg_io.cuh — canonical 32-byte big-endian <-> 8 little-endian 32-bit limbs, and Jacobian -> affine.
// Encoding contract: Scalar::to_bytes (k256/src/arithmetic/scalar.rs:96-98), FieldElement::to_bytes
// (k256/src/arithmetic/field.rs:110-112), p256 ByteOrder::BigEndian (p256/src/arithmetic/field.rs:41).
#pragma once
#include "ecg_prim.cuh"
#include "ecg_point.cuh"

namespace ecg {

// p must be 4-byte aligned (all library buffers are 256-byte aligned device allocations)
ECG_D void load_be32(uint32_t* limbs, const uint8_t* p) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
  for (int i = 0; i < 8; i++) limbs[7 - i] = bswap32(w[i]);
}
ECG_D void store_be32(uint8_t* p, const uint32_t* limbs) {
  uint32_t* w = reinterpret_cast<uint32_t*>(p);
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = bswap32(limbs[7 - i]);
}

// (X:Y:Z) with zinv = 1/Z  ->  canonical affine integers
template <class F>
ECG_D void jac_to_affine_canonical(Fe& x, Fe& y, const Jac& p, const Fe& zinv) {
  Fe z2, z3;
  F::sqr(z2, zinv);
  F::mul(z3, z2, zinv);
  F::mul(x, p.X, z2);
  F::mul(y, p.Y, z3);
  F::to_canonical(x, x);
  F::to_canonical(y, y);
}

}  // namespace ecg
