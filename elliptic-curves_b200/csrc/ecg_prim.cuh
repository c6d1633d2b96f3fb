// ecg_prim.cuh — 32-bit carry-chain primitives for 256-bit modular arithmetic on sm_100a.
//
// Device build: thin wrappers over PTX add.cc / addc / sub.cc / subc / mad.lo.cc / madc.hi.cc.
// ptxas fuses every adjacent (mad[c].lo.cc, madc.hi[.cc]) pair on an aligned register pair into ONE
// `IMAD.WIDE.U32[.X] Rd, Pout, Ra, Rb, Rc, Pin` — a 32x32->64 multiply-accumulate with carry-in and
// carry-out through predicate registers (verified with cuobjdump -sass).
//
// Host build (no __CUDA_ARCH__): the same functions emulate the PTX carry flag with a thread-local
// variable.  That build exists ONLY so tests/sim can run the exact kernel arithmetic on the CPU in a
// container without a GPU; it is never linked into libecgpu.so's execution path.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ECG_HD __host__ __device__ __forceinline__
#define ECG_D __device__ __forceinline__
#else
#define ECG_HD inline
#define ECG_D inline
#endif

namespace ecg {

#if defined(__CUDA_ARCH__)

ECG_D uint32_t add_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
ECG_D uint32_t addc_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
ECG_D uint32_t addc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
ECG_D uint32_t sub_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
ECG_D uint32_t subc_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
ECG_D uint32_t subc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
// (lo,hi) = a*b
ECG_D void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
// (lo,hi) += a*b ; CF = carry out           (starts a chain)
ECG_D void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
               : "+r"(lo), "+r"(hi)
               : "r"(a), "r"(b));
}
// (lo,hi) += a*b + CF ; CF = carry out      (continues a chain)
ECG_D void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
               : "+r"(lo), "+r"(hi)
               : "r"(a), "r"(b));
}
// lo += lo32(a*b) + CF ; hi = hi32(a*b) + carry   (ends a chain on a fresh top limb; cannot overflow)
ECG_D void madc_wide_top(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, 0;"
               : "+r"(lo), "=r"(hi)
               : "r"(a), "r"(b));
}
// lo = lo32(a*b) + CF ; hi = hi32(a*b) + carry    (both limbs fresh)
ECG_D void madc_wide_new(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, 0;\n\tmadc.hi.u32 %1, %2, %3, 0;"
               : "=r"(lo), "=r"(hi)
               : "r"(a), "r"(b));
}
ECG_D uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_r(lo, hi, s); }
ECG_D uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

#else  // ---------------------------------------------------------------- host emulation (tests only)

static thread_local uint32_t g_cf = 0;

inline uint32_t add_cc(uint32_t a, uint32_t b) {
  uint64_t s = (uint64_t)a + b;
  g_cf = (uint32_t)(s >> 32);
  return (uint32_t)s;
}
inline uint32_t addc_cc(uint32_t a, uint32_t b) {
  uint64_t s = (uint64_t)a + b + g_cf;
  g_cf = (uint32_t)(s >> 32);
  return (uint32_t)s;
}
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + g_cf; }
inline uint32_t sub_cc(uint32_t a, uint32_t b) {
  uint64_t s = (uint64_t)a - b;
  g_cf = (uint32_t)(s >> 63);  // PTX: CF holds the borrow for sub.cc/subc
  return (uint32_t)s;
}
inline uint32_t subc_cc(uint32_t a, uint32_t b) {
  uint64_t s = (uint64_t)a - b - g_cf;
  g_cf = (uint32_t)(s >> 63);
  return (uint32_t)s;
}
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - g_cf; }
inline void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  uint64_t p = (uint64_t)a * b;
  lo = (uint32_t)p;
  hi = (uint32_t)(p >> 32);
}
inline void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  uint64_t p = (uint64_t)a * b;
  uint64_t s0 = (uint64_t)lo + (uint32_t)p;
  uint64_t s1 = (uint64_t)hi + (uint32_t)(p >> 32) + (s0 >> 32);
  lo = (uint32_t)s0;
  hi = (uint32_t)s1;
  g_cf = (uint32_t)(s1 >> 32);
}
inline void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  uint64_t p = (uint64_t)a * b;
  uint64_t s0 = (uint64_t)lo + (uint32_t)p + g_cf;
  uint64_t s1 = (uint64_t)hi + (uint32_t)(p >> 32) + (s0 >> 32);
  lo = (uint32_t)s0;
  hi = (uint32_t)s1;
  g_cf = (uint32_t)(s1 >> 32);
}
inline void madc_wide_top(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  uint64_t p = (uint64_t)a * b;
  uint64_t s0 = (uint64_t)lo + (uint32_t)p + g_cf;
  lo = (uint32_t)s0;
  hi = (uint32_t)(p >> 32) + (uint32_t)(s0 >> 32);
}
inline void madc_wide_new(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  uint64_t p = (uint64_t)a * b;
  uint64_t s0 = (uint64_t)(uint32_t)p + g_cf;
  lo = (uint32_t)s0;
  hi = (uint32_t)(p >> 32) + (uint32_t)(s0 >> 32);
}
inline uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {
  s &= 31;
  return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
inline uint32_t bswap32(uint32_t x) { return __builtin_bswap32(x); }

#endif

// ---- 8-limb helpers shared by both fields ---------------------------------------------------

struct Fe {
  uint32_t v[8];  // little-endian 32-bit limbs
};

// r = a + b, returns carry-out (0/1)
ECG_D uint32_t add8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  r[0] = add_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) r[i] = addc_cc(a[i], b[i]);
  return addc(0, 0);
}
// r = a - b, returns borrow (0/1)
ECG_D uint32_t sub8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  r[0] = sub_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) r[i] = subc_cc(a[i], b[i]);
  return 0u - subc(0, 0);
}

// 8x8 -> 16 limb schoolbook product, row-wise, with the "even/odd accumulator" layout: products whose
// low limb lands on an even position accumulate in E, odd positions in O (O[k] holds position k+1), so
// every 32x32 product is ONE aligned-pair IMAD.WIDE.U32.X and each row is two independent 4-long
// carry chains.  64 IMAD.WIDE + 7 ADDC + 15 merge adds.
ECG_D void mul8x8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t E[16], O[16];
  // row 0
#pragma unroll
  for (int m = 0; m < 4; m++) {
    mul_wide(E[2 * m], E[2 * m + 1], a[2 * m], b[0]);
    mul_wide(O[2 * m], O[2 * m + 1], a[2 * m + 1], b[0]);
  }
#pragma unroll
  for (int i = 1; i < 8; i++) {
    if (i & 1) {
      // a_even * b_i -> odd positions i+2m -> O[i-1+2m]; all eight limbs already live; carry -> O[i+7]
      mad_wide_cc(O[i - 1], O[i], a[0], b[i]);
#pragma unroll
      for (int m = 1; m < 4; m++) madc_wide_cc(O[i - 1 + 2 * m], O[i + 2 * m], a[2 * m], b[i]);
      O[i + 7] = addc(0, 0);
      // a_odd * b_i -> even positions i+1+2m -> E[i+1+2m]; top pair: E[i+7] live only when i>=3
      mad_wide_cc(E[i + 1], E[i + 2], a[1], b[i]);
#pragma unroll
      for (int m = 1; m < 3; m++) madc_wide_cc(E[i + 1 + 2 * m], E[i + 2 + 2 * m], a[2 * m + 1], b[i]);
      if (i == 1)
        madc_wide_new(E[i + 7], E[i + 8], a[7], b[i]);
      else
        madc_wide_top(E[i + 7], E[i + 8], a[7], b[i]);
    } else {
      // a_even * b_i -> even positions -> E[i+2m]; carry -> E[i+8]
      mad_wide_cc(E[i], E[i + 1], a[0], b[i]);
#pragma unroll
      for (int m = 1; m < 4; m++) madc_wide_cc(E[i + 2 * m], E[i + 1 + 2 * m], a[2 * m], b[i]);
      E[i + 8] = addc(0, 0);
      // a_odd * b_i -> odd positions i+1+2m -> O[i+2m]; O[i+6] live (carry limb), O[i+7] fresh
      mad_wide_cc(O[i], O[i + 1], a[1], b[i]);
#pragma unroll
      for (int m = 1; m < 3; m++) madc_wide_cc(O[i + 2 * m], O[i + 1 + 2 * m], a[2 * m + 1], b[i]);
      madc_wide_top(O[i + 6], O[i + 7], a[7], b[i]);
    }
  }
  // merge: r = E + (O << 32)
  r[0] = E[0];
  r[1] = add_cc(E[1], O[0]);
#pragma unroll
  for (int k = 2; k < 15; k++) r[k] = addc_cc(E[k], O[k - 1]);
  r[15] = addc(E[15], O[14]);
}

}  // namespace ecg
