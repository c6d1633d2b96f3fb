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
#define ECG_D __device__ __forceinline__
// kernels (ecg_kernels.cuh, ecg_msm.cuh): ECG_KERNEL(launch bounds...) name(args) { ... }
#define ECG_KERNEL(...) __global__ void __launch_bounds__(__VA_ARGS__)
#define ECG_DEV __device__ __forceinline__
#else
#define ECG_D inline
// host build of a kernel header: only tests/sim/sim.cpp (ECG_HOST_SIM), which runs a kernel body per simulated thread
#define ECG_KERNEL(...) static void
#define ECG_DEV inline
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
// lo += lo32(a*b) ; hi = hi32(a*b) + carry       (one-product chain on a live low limb and a fresh top limb)
ECG_D void mad_wide_top(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, 0;"
               : "+r"(lo), "=r"(hi)
               : "r"(a), "r"(b));
}
// lo = lo32(a*b) + CF ; hi = hi32(a*b) + carry    (both limbs fresh)
ECG_D void madc_wide_new(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, 0;\n\tmadc.hi.u32 %1, %2, %3, 0;"
               : "=r"(lo), "=r"(hi)
               : "r"(a), "r"(b));
}
// (c0, c1, c2) += a*b   (three-word running accumulator of a product-scanning column)
ECG_D void mad_acc3(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t a, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
               : "+r"(c0), "+r"(c1), "+r"(c2)
               : "r"(a), "r"(b));
}
// the same with early-clobber accumulators: a multiplicand may be a value the compiler keeps in the register of c0
// (the Montgomery factor m = c0 * n0inv when n0inv == 1: sm2, P-192), and c0 is written before a is read again
ECG_D void mad_acc3x(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t a, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
               : "+&r"(c0), "+&r"(c1), "+&r"(c2)
               : "r"(a), "r"(b));
}
ECG_D uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_r(lo, hi, s); }
ECG_D uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_l(lo, hi, s); }
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
inline void mad_wide_top(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  uint64_t p = (uint64_t)a * b;
  uint64_t s0 = (uint64_t)lo + (uint32_t)p;
  lo = (uint32_t)s0;
  hi = (uint32_t)(p >> 32) + (uint32_t)(s0 >> 32);
}
inline void mad_acc3(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t a, uint32_t b) {
  uint64_t p = (uint64_t)a * b;
  uint64_t s0 = (uint64_t)c0 + (uint32_t)p;
  uint64_t s1 = (uint64_t)c1 + (uint32_t)(p >> 32) + (s0 >> 32);
  c0 = (uint32_t)s0;
  c1 = (uint32_t)s1;
  c2 += (uint32_t)(s1 >> 32);
}
inline void mad_acc3x(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t a, uint32_t b) { mad_acc3(c0, c1, c2, a, b); }
inline uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {
  s &= 31;
  return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
inline uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t s) {  // high word of ((hi:lo) << s)
  s &= 31;
  return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
inline uint32_t bswap32(uint32_t x) { return __builtin_bswap32(x); }

#endif

// ---- 8-limb helpers shared by both fields ---------------------------------------------------

// ECG_FE_ALIGN = 16 lets local-memory copies of a field element move as two 128-bit accesses (only the kbench
// "operands through memory" experiment, OPT bit 8, sets it; the library keeps natural alignment)
#ifndef ECG_FE_ALIGN
#define ECG_FE_ALIGN 4
#endif
// NL 32-bit limbs, little-endian.  8 limbs: secp256k1 / P-256; 12 limbs: P-384.  Everything above the field layer is
// written against a field policy F (F::NL, F::FeT, F::JacT, F::AffT), so a curve is a policy plus constants.
template <int NL>
struct alignas(ECG_FE_ALIGN) FeN {
  uint32_t v[NL];
};
typedef FeN<8> Fe;
template <int NL>
struct JacN {  // Jacobian (X:Y:Z) = (X/Z^2, Y/Z^3); Z == 0 (mod p) is the identity
  FeN<NL> X, Y, Z;
};
template <int NL>
struct AffN {
  FeN<NL> x, y;
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

// N-limb versions (N compile-time): r = a + b / a - b with carry / borrow out
template <int N>
ECG_D uint32_t addN(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  r[0] = add_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < N; i++) r[i] = addc_cc(a[i], b[i]);
  return addc(0, 0);
}
template <int N>
ECG_D uint32_t subN(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  r[0] = sub_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < N; i++) r[i] = subc_cc(a[i], b[i]);
  return 0u - subc(0, 0);
}

// N x N -> 2N limb schoolbook product for any even N: the row structure of mul8x8 below with the bounds written in N
// (products whose low limb lands on an even position accumulate in E, odd positions in O; O[k] holds position k+1).
// N*N IMAD.WIDE.  Used for the 12-limb field (P-384); the 8-limb fields keep the hand-checked mul8x8.
template <int N>
ECG_D void mulNxN(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int H = N / 2;
  uint32_t E[2 * N], O[2 * N];
#pragma unroll
  for (int m = 0; m < H; m++) {
    mul_wide(E[2 * m], E[2 * m + 1], a[2 * m], b[0]);
    mul_wide(O[2 * m], O[2 * m + 1], a[2 * m + 1], b[0]);
  }
#pragma unroll
  for (int i = 1; i < N; i++) {
    if (i & 1) {
      mad_wide_cc(O[i - 1], O[i], a[0], b[i]);
#pragma unroll
      for (int m = 1; m < H; m++) madc_wide_cc(O[i - 1 + 2 * m], O[i + 2 * m], a[2 * m], b[i]);
      O[i + N - 1] = addc(0, 0);
      mad_wide_cc(E[i + 1], E[i + 2], a[1], b[i]);
#pragma unroll
      for (int m = 1; m < H - 1; m++) madc_wide_cc(E[i + 1 + 2 * m], E[i + 2 + 2 * m], a[2 * m + 1], b[i]);
      if (i == 1)
        madc_wide_new(E[i + N - 1], E[i + N], a[N - 1], b[i]);
      else
        madc_wide_top(E[i + N - 1], E[i + N], a[N - 1], b[i]);
    } else {
      mad_wide_cc(E[i], E[i + 1], a[0], b[i]);
#pragma unroll
      for (int m = 1; m < H; m++) madc_wide_cc(E[i + 2 * m], E[i + 1 + 2 * m], a[2 * m], b[i]);
      E[i + N] = addc(0, 0);
      mad_wide_cc(O[i], O[i + 1], a[1], b[i]);
#pragma unroll
      for (int m = 1; m < H - 1; m++) madc_wide_cc(O[i + 2 * m], O[i + 1 + 2 * m], a[2 * m + 1], b[i]);
      madc_wide_top(O[i + N - 2], O[i + N - 1], a[N - 1], b[i]);
    }
  }
  r[0] = E[0];
  r[1] = add_cc(E[1], O[0]);
#pragma unroll
  for (int k = 2; k < 2 * N - 1; k++) r[k] = addc_cc(E[k], O[k - 1]);
  r[2 * N - 1] = addc(E[2 * N - 1], O[2 * N - 2]);
}

// N-limb square -> 2N limbs for any N: the N(N-1)/2 cross products a_i*a_j (i < j) once, accumulated column by column
// in a three-word running sum (each product = one mad.lo / madc.hi pair: the same multiplier time as one IMAD.WIDE),
// doubled by a 1-bit funnel shift, and the N diagonal squares added with one IMAD.WIDE.X chain: N(N+1)/2 multiplier
// slots instead of N*N (12 limbs: 78 instead of 144).  The 8-limb fields keep the hand-scheduled sqr8 below.
template <int N>
ECG_D void sqrN(uint32_t* r, const uint32_t* a) {
  uint32_t S[2 * N];
  uint32_t c0 = 0, c1 = 0, c2 = 0;
  S[0] = 0;
#pragma unroll
  for (int k = 1; k <= 2 * N - 3; k++) {
#pragma unroll
    for (int i = 0; i < N; i++) {
      const int j = k - i;
      if (i < j && j < N) mad_acc3(c0, c1, c2, a[i], a[j]);
    }
    S[k] = c0;
    c0 = c1;
    c1 = c2;
    c2 = 0;
  }
  S[2 * N - 2] = c0;
  S[2 * N - 1] = c1;
  uint32_t T[2 * N];
#pragma unroll
  for (int k = 2 * N - 1; k >= 1; k--) T[k] = funnel_l(S[k - 1], S[k], 1);
  T[0] = 0;
  mad_wide_cc(T[0], T[1], a[0], a[0]);
#pragma unroll
  for (int i = 1; i < N; i++) madc_wide_cc(T[2 * i], T[2 * i + 1], a[i], a[i]);
#pragma unroll
  for (int k = 0; k < 2 * N; k++) r[k] = T[k];
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

// 4x4 -> 8 limb product in the even/odd pair layout (16 IMAD.WIDE + 3 ADDC + 7 merge adds).
ECG_D void mul4x4(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t E[8], O[7];
  mul_wide(E[0], E[1], a[0], b[0]);
  mul_wide(E[2], E[3], a[2], b[0]);
  mul_wide(O[0], O[1], a[1], b[0]);
  mul_wide(O[2], O[3], a[3], b[0]);
  // row 1: a_even*b1 -> odd positions 1,3 ; a_odd*b1 -> even positions 2,4
  mad_wide_cc(O[0], O[1], a[0], b[1]);
  madc_wide_cc(O[2], O[3], a[2], b[1]);
  O[4] = addc(0, 0);
  mad_wide_cc(E[2], E[3], a[1], b[1]);
  madc_wide_new(E[4], E[5], a[3], b[1]);
  // row 2: a_even*b2 -> even positions 2,4 ; a_odd*b2 -> odd positions 3,5
  mad_wide_cc(E[2], E[3], a[0], b[2]);
  madc_wide_cc(E[4], E[5], a[2], b[2]);
  E[6] = addc(0, 0);
  mad_wide_cc(O[2], O[3], a[1], b[2]);
  madc_wide_top(O[4], O[5], a[3], b[2]);
  // row 3: a_even*b3 -> odd positions 3,5 ; a_odd*b3 -> even positions 4,6
  mad_wide_cc(O[2], O[3], a[0], b[3]);
  madc_wide_cc(O[4], O[5], a[2], b[3]);
  O[6] = addc(0, 0);
  mad_wide_cc(E[4], E[5], a[1], b[3]);
  madc_wide_top(E[6], E[7], a[3], b[3]);
  r[0] = E[0];
  r[1] = add_cc(E[1], O[0]);
#pragma unroll
  for (int k = 2; k < 7; k++) r[k] = addc_cc(E[k], O[k - 1]);
  r[7] = addc(E[7], O[6]);
}

// 8x8 -> 16 limb product by one level of Karatsuba on 4-limb halves: 48 IMAD.WIDE instead of 64, paid for with
// ~60 more carry-chain adds (ALU pipe).  a = a0 + a1 B^4, b = b0 + b1 B^4:
//   z0 = a0 b0, z2 = a1 b1, zm = (a0 + a1)(b0 + b1), z1 = zm - z0 - z2,  r = z0 + z1 B^4 + z2 B^8.
ECG_D void mul8x8_kara(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t z0[8], z2[8], zm[9], sa[4], sb[4];
  mul4x4(z0, a, b);
  mul4x4(z2, a + 4, b + 4);
  sa[0] = add_cc(a[0], a[4]);
  sa[1] = addc_cc(a[1], a[5]);
  sa[2] = addc_cc(a[2], a[6]);
  sa[3] = addc_cc(a[3], a[7]);
  uint32_t ca = addc(0, 0);
  sb[0] = add_cc(b[0], b[4]);
  sb[1] = addc_cc(b[1], b[5]);
  sb[2] = addc_cc(b[2], b[6]);
  sb[3] = addc_cc(b[3], b[7]);
  uint32_t cb = addc(0, 0);
  mul4x4(zm, sa, sb);
  // (sa + ca B^4)(sb + cb B^4) = sa sb + (ca sb + cb sa) B^4 + ca cb B^8
  uint32_t ma = 0u - ca, mb = 0u - cb;
  zm[4] = add_cc(zm[4], sb[0] & ma);
  zm[5] = addc_cc(zm[5], sb[1] & ma);
  zm[6] = addc_cc(zm[6], sb[2] & ma);
  zm[7] = addc_cc(zm[7], sb[3] & ma);
  zm[8] = addc(ca & cb, 0);
  zm[4] = add_cc(zm[4], sa[0] & mb);
  zm[5] = addc_cc(zm[5], sa[1] & mb);
  zm[6] = addc_cc(zm[6], sa[2] & mb);
  zm[7] = addc_cc(zm[7], sa[3] & mb);
  zm[8] = addc(zm[8], 0);
  // z1 = zm - z0 - z2   (non-negative, < 2^258)
  zm[0] = sub_cc(zm[0], z0[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) zm[i] = subc_cc(zm[i], z0[i]);
  zm[8] = subc(zm[8], 0);
  zm[0] = sub_cc(zm[0], z2[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) zm[i] = subc_cc(zm[i], z2[i]);
  zm[8] = subc(zm[8], 0);
  // r = z0 + z1 B^4 + z2 B^8
#pragma unroll
  for (int i = 0; i < 4; i++) r[i] = z0[i];
  r[4] = add_cc(z0[4], zm[0]);
  r[5] = addc_cc(z0[5], zm[1]);
  r[6] = addc_cc(z0[6], zm[2]);
  r[7] = addc_cc(z0[7], zm[3]);
  r[8] = addc_cc(z2[0], zm[4]);
  r[9] = addc_cc(z2[1], zm[5]);
  r[10] = addc_cc(z2[2], zm[6]);
  r[11] = addc_cc(z2[3], zm[7]);
  r[12] = addc_cc(z2[4], zm[8]);
  r[13] = addc_cc(z2[5], 0);
  r[14] = addc_cc(z2[6], 0);
  r[15] = addc(z2[7], 0);
}

// 8-limb square -> 16 limbs: the 28 cross products a_i*a_j (i<j) once, in the same even/odd pair layout
// (rows are triangular, so chains shorten), doubled by a 1-bit funnel shift, and the 8 diagonal squares
// accumulated on top with one IMAD.WIDE.X chain.  36 IMAD.WIDE instead of 64.
ECG_D void sqr8(uint32_t* r, const uint32_t* a) {
  uint32_t E[14], O[14];
  // row 0: a0 * a[1..7]  (all limbs fresh, independent products)
  mul_wide(O[0], O[1], a[0], a[1]);
  mul_wide(E[2], E[3], a[0], a[2]);
  mul_wide(O[2], O[3], a[0], a[3]);
  mul_wide(E[4], E[5], a[0], a[4]);
  mul_wide(O[4], O[5], a[0], a[5]);
  mul_wide(E[6], E[7], a[0], a[6]);
  mul_wide(O[6], O[7], a[0], a[7]);
  // row 1: a1 * a[2..7]: odd positions 3,5,7 -> O[2..7]; even positions 4,6,8 -> E[4..9]
  mad_wide_cc(O[2], O[3], a[1], a[2]);
  madc_wide_cc(O[4], O[5], a[1], a[4]);
  madc_wide_cc(O[6], O[7], a[1], a[6]);
  O[8] = addc(0, 0);
  mad_wide_cc(E[4], E[5], a[1], a[3]);
  madc_wide_cc(E[6], E[7], a[1], a[5]);
  madc_wide_new(E[8], E[9], a[1], a[7]);
  // row 2: a2 * a[3..7]: odd positions 5,7,9 -> O[4..9]; even positions 6,8 -> E[6..9]
  mad_wide_cc(O[4], O[5], a[2], a[3]);
  madc_wide_cc(O[6], O[7], a[2], a[5]);
  madc_wide_top(O[8], O[9], a[2], a[7]);
  mad_wide_cc(E[6], E[7], a[2], a[4]);
  madc_wide_cc(E[8], E[9], a[2], a[6]);
  E[10] = addc(0, 0);
  // row 3: a3 * a[4..7]: odd positions 7,9 -> O[6..9]; even positions 8,10 -> E[8..11]
  mad_wide_cc(O[6], O[7], a[3], a[4]);
  madc_wide_cc(O[8], O[9], a[3], a[6]);
  O[10] = addc(0, 0);
  mad_wide_cc(E[8], E[9], a[3], a[5]);
  madc_wide_top(E[10], E[11], a[3], a[7]);
  // row 4: a4 * a[5..7]: odd positions 9,11 -> O[8..11]; even position 10 -> E[10..11]
  mad_wide_cc(O[8], O[9], a[4], a[5]);
  madc_wide_top(O[10], O[11], a[4], a[7]);
  mad_wide_cc(E[10], E[11], a[4], a[6]);
  E[12] = addc(0, 0);
  // row 5: a5 * a[6..7]: odd position 11 -> O[10..11]; even position 12 -> E[12..13]
  mad_wide_cc(O[10], O[11], a[5], a[6]);
  O[12] = addc(0, 0);
  mad_wide_top(E[12], E[13], a[5], a[7]);
  // row 6: a6 * a7: odd position 13 -> O[12..13]
  mad_wide_top(O[12], O[13], a[6], a[7]);
  // merge: S = E + (O << 32), limbs 1..14  (S[0] = 0, E[0] = E[1] = 0)
  uint32_t S[16];
  S[0] = 0;
  S[1] = O[0];
  S[2] = add_cc(E[2], O[1]);
#pragma unroll
  for (int k = 3; k < 14; k++) S[k] = addc_cc(E[k], O[k - 1]);
  S[14] = addc(0, O[13]);
  // double (1-bit left funnel) and add the diagonal squares with one carry chain of IMAD.WIDE.X
  uint32_t T[16];
  T[15] = S[14] >> 31;
#pragma unroll
  for (int k = 14; k >= 1; k--) T[k] = funnel_l(S[k - 1], S[k], 1);
  T[0] = 0;
  mad_wide_cc(T[0], T[1], a[0], a[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) madc_wide_cc(T[2 * i], T[2 * i + 1], a[i], a[i]);
#pragma unroll
  for (int k = 0; k < 16; k++) r[k] = T[k];
}

}  // namespace ecg
