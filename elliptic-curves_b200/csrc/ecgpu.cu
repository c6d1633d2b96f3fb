// ecgpu.cu — kernels and C ABI of libecgpu.so (see include/ecgpu.h for the contract).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
// There is no CPU fallback in this file: every compute entry launches sm_100a kernels or fails.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/ecgpu.h"
#include "ecg_curves.cuh"
#include "ecg_io.cuh"
#include "ecg_mul.cuh"
#include "ecg_msm.cuh"
#include "ecg_verify.cuh"

using namespace ecg;

// ------------------------------------------------------------------------------------------------
// error flags written by kernels into status[0]; status[1] = smallest offending index
#define ERRF_SCALAR 1u
#define ERRF_POINT 2u

__device__ __forceinline__ void report_error(uint32_t* status, uint32_t flag, size_t idx) {
  atomicOr(&status[0], flag);
  atomicMin(&status[1], (uint32_t)(idx > 0xFFFFFFFEull ? 0xFFFFFFFEull : idx));
}

// SoA word-major intermediate layout: word w of element idx at buf[w*n + idx] (coalesced per word)
template <int NW>
__device__ __forceinline__ void soa_store(uint32_t* buf, size_t n, size_t idx, const uint32_t* v, int w0) {
#pragma unroll
  for (int w = 0; w < NW; w++) buf[(size_t)(w0 + w) * n + idx] = v[w];
}
template <int NW>
__device__ __forceinline__ void soa_load(uint32_t* v, const uint32_t* buf, size_t n, size_t idx, int w0) {
#pragma unroll
  for (int w = 0; w < NW; w++) v[w] = buf[(size_t)(w0 + w) * n + idx];
}

// Load + validate one (scalar, point) pair.  Returns error flags (0 = fine).  On error / identity the
// caller still runs the arithmetic on a harmless substitute (k = 1, P = G) and forces Z = 0 afterwards so
// that warps stay converged.
template <class C>
__device__ __forceinline__ uint32_t load_pair(uint32_t* k, Aff& P, bool& inf, const uint8_t* kb,
                                              const uint8_t* pxy, const uint8_t* pinf, size_t idx) {
  typedef typename C::F F;
  uint32_t err = 0;
  load_be32(k, kb + 32 * idx);
  if (!lt8(k, C::N())) err |= ERRF_SCALAR;
  inf = pinf != nullptr && pinf[idx] != 0;
  Fe x, y;
  load_be32(x.v, pxy + 64 * idx);
  load_be32(y.v, pxy + 64 * idx + 32);
  if (!inf) {
    bool ok = lt8(x.v, C::P()) && lt8(y.v, C::P());
    F::from_canonical(P.x, x);
    F::from_canonical(P.y, y);
    if (ok) {
      Fe b;
      C::b_internal(b);
      ok = aff_on_curve<F, C::A_IS_MINUS3>(P, b);
    }
    if (!ok) err |= ERRF_POINT;
  }
  if (inf || err) {
    C::generator(P);
#pragma unroll
    for (int i = 0; i < 8; i++) k[i] = (i == 0);
  }
  return err;
}

// ------------------------------------------------------------------------------------------------
// Per-thread window tables live in global memory, one slot per block: word w of entry e of thread t at
// gtab[blockIdx*BLOCK*EW + (e*WPE + w)*BLOCK + t]  (EW = words per thread: 128 affine / 192 Jacobian).
// A warp's access to one (e, w) of differing e per lane touches 32 distinct 4-byte words spread over at most 8
// rows; the blocks resident at any time keep ~50 MB of tables live, which stays in the 126 MB L2.  Shared memory
// was the first home of these tables (512-768 B/thread capped occupancy at 8-12 warps/SM); moving them out lets
// registers set the occupancy (16-20 warps/SM) and measured +7 % (k256) / +16 % (P-256), tools/kbench.cu.
#define K_TAB_WORDS 128  /* 8 affine entries  x 16 words */
#define P_TAB_WORDS 192  /* 8 Jacobian entries x 24 words */

// secp256k1 variable-base: one pair per thread.
template <int BLOCK, int MINBLK>
__global__ void __launch_bounds__(BLOCK, MINBLK)
    k256_varbase_kernel(const uint8_t* __restrict__ kb, const uint8_t* __restrict__ pxy,
                        const uint8_t* __restrict__ pinf, size_t n, uint32_t* __restrict__ jac,
                        uint32_t* __restrict__ gtab, uint32_t* __restrict__ status, size_t base) {
  size_t idx = (size_t)blockIdx.x * BLOCK + threadIdx.x;
  if (idx >= n) return;
  uint32_t k[8];
  Aff P;
  bool inf;
  uint32_t err = load_pair<CurveK256>(k, P, inf, kb, pxy, pinf, idx);
  if (err) report_error(status, err, base + idx);
  TabRef tab{gtab + (size_t)blockIdx.x * BLOCK * K_TAB_WORDS + threadIdx.x, (uint32_t)BLOCK};
  Jac r;
  k256_mul_thread(r, k, P, tab);
  if (inf || err) FpK256::set_zero(r.Z);
  soa_store<8>(jac, n, idx, r.X.v, 0);
  soa_store<8>(jac, n, idx, r.Y.v, 8);
  soa_store<8>(jac, n, idx, r.Z.v, 16);
}

// ------------------------------------------------------------------------------------------------
// Generic prime-order curve (P-256) variable-base: Jacobian window table (768 B / thread).
template <class C, int BLOCK, int MINBLK>
__global__ void __launch_bounds__(BLOCK, MINBLK)
    generic_varbase_kernel(const uint8_t* __restrict__ kb, const uint8_t* __restrict__ pxy,
                           const uint8_t* __restrict__ pinf, size_t n, uint32_t* __restrict__ jac,
                           uint32_t* __restrict__ gtab, uint32_t* __restrict__ status, size_t base) {
  typedef typename C::F F;
  size_t idx = (size_t)blockIdx.x * BLOCK + threadIdx.x;
  if (idx >= n) return;
  uint32_t k[8];
  Aff P;
  bool inf;
  uint32_t err = load_pair<C>(k, P, inf, kb, pxy, pinf, idx);
  if (err) report_error(status, err, base + idx);
  TabRefJ tab{gtab + (size_t)blockIdx.x * BLOCK * P_TAB_WORDS + threadIdx.x, (uint32_t)BLOCK};
  Jac r;
  generic_mul_thread<F, C::A_IS_MINUS3>(r, k, P, tab);
  if (inf || err) F::set_zero(r.Z);
  soa_store<8>(jac, n, idx, r.X.v, 0);
  soa_store<8>(jac, n, idx, r.Y.v, 8);
  soa_store<8>(jac, n, idx, r.Z.v, 16);
}

// ------------------------------------------------------------------------------------------------
// Fixed-base k*G from a device-resident table of affine odd multiples.
//   table layout: window i (0..FB_WINDOWS-1), entry j (0..2^(FB_W-1)-1) = (2j+1) * 2^(FB_W*i) * G, 16 words
//   (x[8], y[8], internal form); one extra entry at the end = 2^256 * G (the recoding's implicit top digit).
// Replaces BasepointTable (primeorder/src/tables/basepoint.rs:41-125; k256/src/arithmetic/tables.rs:12-22):
// same idea (precomputed multiples of G, only additions at run time), sized for a 126 MB L2 instead of a
// 30 KiB L1: 16 sixteen-bit windows -> 17 mixed additions and no doubling per scalar.
#define FB_W 16
#define FB_WINDOWS 16
#define FB_ENTRIES (1u << (FB_W - 1))
#define FB_TABLE_POINTS ((size_t)FB_WINDOWS * FB_ENTRIES + 1)

__device__ __forceinline__ void fb_load_entry(Aff& e, const uint32_t* __restrict__ table, size_t point) {
  const uint4* p = reinterpret_cast<const uint4*>(table + point * 16);
  uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
  e.x.v[0] = a.x; e.x.v[1] = a.y; e.x.v[2] = a.z; e.x.v[3] = a.w;
  e.x.v[4] = b.x; e.x.v[5] = b.y; e.x.v[6] = b.z; e.x.v[7] = b.w;
  e.y.v[0] = c.x; e.y.v[1] = c.y; e.y.v[2] = c.z; e.y.v[3] = c.w;
  e.y.v[4] = d.x; e.y.v[5] = d.y; e.y.v[6] = d.z; e.y.v[7] = d.w;
}

// acc += k*G (acc Jacobian on the true curve; pass Z = 0 to start from the identity)
template <class C, bool FROM_IDENTITY>
__device__ __forceinline__ void fixedbase_accumulate(Jac& acc, const uint32_t* k, const uint32_t* __restrict__ table) {
  typedef typename C::F F;
  FullRecode rc;
  recode_full(rc, k);
  Aff e;
  fb_load_entry(e, table, (size_t)FB_WINDOWS * FB_ENTRIES);  // 2^256 * G
  if (FROM_IDENTITY) {
    acc.X = e.x;
    acc.Y = e.y;
    F::set_one(acc.Z);
  } else {
    jac_madd<F, C::A_IS_MINUS3>(acc, acc, e);
  }
#pragma unroll 1
  for (int i = 0; i < FB_WINDOWS; i++) {
    uint32_t w = rc.h[0] & 0xFFFFu;
#pragma unroll
    for (int j = 0; j < 7; j++) rc.h[j] = funnel_r(rc.h[j], rc.h[j + 1], 16);
    rc.h[7] >>= 16;
    uint32_t pos = w >> (FB_W - 1);
    uint32_t idx = pos ? (w & (FB_ENTRIES - 1)) : (FB_ENTRIES - 1 - w);
    fb_load_entry(e, table, (size_t)i * FB_ENTRIES + idx);
    fe_cneg<F>(e.y, pos ^ 1u);
    jac_madd<F, C::A_IS_MINUS3>(acc, acc, e);
  }
  // parity correction: subtract G if k was even
  fb_load_entry(e, table, 0);
  F::neg(e.y, e.y);
  Jac t;
  jac_madd<F, C::A_IS_MINUS3>(t, acc, e);
  jac_csel(acc, t, rc.even);
}

template <class C>
__global__ void __launch_bounds__(128, 4)
    fixedbase_kernel(const uint8_t* __restrict__ kb, size_t n, const uint32_t* __restrict__ table,
                     uint32_t* __restrict__ jac, uint32_t* __restrict__ status, size_t base) {
  typedef typename C::F F;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  uint32_t k[8];
  load_be32(k, kb + 32 * idx);
  bool bad = !lt8(k, C::N());
  if (bad) {
    report_error(status, ERRF_SCALAR, base + idx);
#pragma unroll
    for (int i = 0; i < 8; i++) k[i] = (i == 0);
  }
  Jac acc;
  fixedbase_accumulate<C, true>(acc, k, table);
  if (bad) F::set_zero(acc.Z);
  soa_store<8>(jac, n, idx, acc.X.v, 0);
  soa_store<8>(jac, n, idx, acc.Y.v, 8);
  soa_store<8>(jac, n, idx, acc.Z.v, 16);
}

// a*G + b*P : variable-base thread routine, then the fixed-base accumulation on the same accumulator.
// Replaces mul_by_generator_and_mul_add_vartime (k256/src/arithmetic/mul.rs:303-310, primeorder/src/mul_backend.rs:31-40).
template <class C, int BLOCK, int MINBLK, bool IS_K256>
__global__ void __launch_bounds__(BLOCK, MINBLK)
    mul_gen_add_kernel(const uint8_t* __restrict__ ab, const uint8_t* __restrict__ kb,
                       const uint8_t* __restrict__ pxy, const uint8_t* __restrict__ pinf, size_t n,
                       const uint32_t* __restrict__ table, uint32_t* __restrict__ jac, uint32_t* __restrict__ gtab,
                       uint32_t* __restrict__ status, size_t base) {
  typedef typename C::F F;
  size_t idx = (size_t)blockIdx.x * BLOCK + threadIdx.x;
  if (idx >= n) return;
  uint32_t k[8], a[8];
  Aff P;
  bool inf;
  uint32_t err = load_pair<C>(k, P, inf, kb, pxy, pinf, idx);
  load_be32(a, ab + 32 * idx);
  if (!lt8(a, C::N())) err |= ERRF_SCALAR;
  if (err) report_error(status, err, base + idx);
  Jac r;
  if (IS_K256) {
    TabRef tab{gtab + (size_t)blockIdx.x * BLOCK * K_TAB_WORDS + threadIdx.x, (uint32_t)BLOCK};
    k256_mul_thread(r, k, P, tab);
  } else {
    TabRefJ tab{gtab + (size_t)blockIdx.x * BLOCK * P_TAB_WORDS + threadIdx.x, (uint32_t)BLOCK};
    generic_mul_thread<F, C::A_IS_MINUS3>(r, k, P, tab);
  }
  if (inf) F::set_zero(r.Z);  // b * O = O, the sum is a*G
  if (err) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = (i == 0);
  }
  fixedbase_accumulate<C, false>(r, a, table);
  if (err) F::set_zero(r.Z);
  soa_store<8>(jac, n, idx, r.X.v, 0);
  soa_store<8>(jac, n, idx, r.Y.v, 8);
  soa_store<8>(jac, n, idx, r.Z.v, 16);
}

// ------------------------------------------------------------------------------------------------
// Signature verification front ends (ecg_verify.cuh).  Both reduce to a*G + b*P through mul_gen_add_kernel; the
// kernels here prepare (a, b, P) and judge the result.  Invalid encodings never raise an API error: they are
// marked not-ok, replaced by harmless operands (a = b = 1, P = G) so warps stay converged, and reported as
// valid[i] = 0 — the reference returns Err(Error) per signature, not a batch failure.
__device__ __forceinline__ void store_scalar_be(uint8_t* dst, const uint32_t* limbs) { store_be32(dst, limbs); }

// BIP340: pk (x only), 32-byte message, signature r || s.
__global__ void __launch_bounds__(128)
    schnorr_prep_kernel(const uint8_t* __restrict__ pk, const uint8_t* __restrict__ msg, const uint8_t* __restrict__ sig, size_t n,
                        uint8_t* __restrict__ pxy, uint8_t* __restrict__ a_out, uint8_t* __restrict__ b_out, uint8_t* __restrict__ ok_out) {
  typedef FpK256 F;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  uint32_t x[8], r[8], sv[8], e[8];
  load_be32(x, pk + 32 * idx);
  load_be32(r, sig + 64 * idx);
  load_be32(sv, sig + 64 * idx + 32);
  Aff P;
  bool ok = k256_lift_x<F>(P, x);                    // VerifyingKey::from_bytes (schnorr/verifying.rs:36-52)
  ok = ok && lt8(r, K256_P);                           // Signature::try_from: r is a field element,
  ok = ok && lt8(sv, K256_N) && !FnMont<CurveK256>::is_zero(sv);  //   s a non-zero scalar (schnorr.rs:150-170)
  bip340_challenge(e, sig + 64 * idx, pk + 32 * idx, msg + 32 * idx);
  if (!lt8(e, K256_N)) {  // Reduce<FieldBytes>: one conditional subtraction (2^256 < 2n)
    uint32_t t[8];
    sub8(t, e, K256_N);
#pragma unroll
    for (int i = 0; i < 8; i++) e[i] = t[i];
  }
  // b = -e mod n
  uint32_t ne[8];
  if (FnMont<CurveK256>::is_zero(e)) {
#pragma unroll
    for (int i = 0; i < 8; i++) ne[i] = 0;
  } else {
    sub8(ne, K256_N, e);
  }
  if (!ok) {
    CurveK256::generator(P);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      sv[i] = (i == 0);
      ne[i] = (i == 0);
    }
  }
  Fe cx, cy;
  F::to_canonical(cx, P.x);
  F::to_canonical(cy, P.y);
  store_be32(pxy + 64 * idx, cx.v);
  store_be32(pxy + 64 * idx + 32, cy.v);
  store_scalar_be(a_out + 32 * idx, sv);
  store_scalar_be(b_out + 32 * idx, ne);
  ok_out[idx] = ok ? 1 : 0;
}
__global__ void __launch_bounds__(256)
    schnorr_check_kernel(const uint8_t* __restrict__ sig, const uint8_t* __restrict__ rxy, const uint8_t* __restrict__ rinf,
                         const uint8_t* __restrict__ ok, size_t n, uint8_t* __restrict__ valid) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const uint32_t* r = reinterpret_cast<const uint32_t*>(sig + 64 * idx);
  const uint32_t* x = reinterpret_cast<const uint32_t*>(rxy + 64 * idx);
  bool same = true;
#pragma unroll
  for (int i = 0; i < 8; i++) same = same && (r[i] == x[i]);
  bool y_even = (rxy[64 * idx + 63] & 1u) == 0;
  valid[idx] = (ok[idx] && !rinf[idx] && y_even && same) ? 1 : 0;  // verifying.rs:94
}

// ECDSA: z (32-byte hash), signature r || s, public key Q (x || y).  One modular inversion per thread slice
// (Montgomery's trick over s_i, as in normalize_kernel); scr: 8*n words.
template <class C>
__global__ void __launch_bounds__(128)
    ecdsa_prep_kernel(const uint8_t* __restrict__ zb, const uint8_t* __restrict__ sig, const uint8_t* __restrict__ qxy, size_t n,
                      int low_s_only, uint32_t* __restrict__ scr, uint8_t* __restrict__ pxy, uint8_t* __restrict__ a_out,
                      uint8_t* __restrict__ b_out, uint8_t* __restrict__ ok_out) {
  typedef typename C::F F;
  typedef FnMont<C> N;
  size_t T = (size_t)gridDim.x * blockDim.x;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  uint32_t acc[8], sm[8];
#pragma unroll
  for (int i = 0; i < 8; i++) acc[i] = C::N_ONE()[i];
  size_t last = t;
  // forward: validate, prefix products of the (Montgomery-form) s_i
  for (size_t idx = t; idx < n; idx += T) {
    uint32_t r[8], sv[8];
    load_be32(r, sig + 64 * idx);
    load_be32(sv, sig + 64 * idx + 32);
    bool ok = lt8(r, C::N()) && !N::is_zero(r) && lt8(sv, C::N()) && !N::is_zero(sv);
    if (ok && low_s_only) {  // EcdsaCurve::NORMALIZE_S (k256/src/ecdsa.rs:104-106): reject s > n/2
      uint32_t twice[8];
      uint32_t c = add8(twice, sv, sv);
      ok = !c && lt8(twice, C::N());
    }
    Aff Q;
    Fe qx, qy;
    load_be32(qx.v, qxy + 64 * idx);
    load_be32(qy.v, qxy + 64 * idx + 32);
    bool qok = lt8(qx.v, C::P()) && lt8(qy.v, C::P());
    F::from_canonical(Q.x, qx);
    F::from_canonical(Q.y, qy);
    if (qok) {
      Fe b;
      C::b_internal(b);
      qok = aff_on_curve<F, C::A_IS_MINUS3>(Q, b);
    }
    ok = ok && qok;
    ok_out[idx] = ok ? 1 : 0;
    if (!ok) {
#pragma unroll
      for (int i = 0; i < 8; i++) sv[i] = (i == 0);
    }
    N::to_mont(sm, sv);
    soa_store<8>(scr, n, idx, acc, 0);
    N::mul(acc, acc, sm);
    last = idx;
  }
  uint32_t inv[8];
  N::inv(inv, acc);
  for (size_t idx = last;; idx -= T) {
    uint32_t r[8], sv[8], z[8], pre[8], w[8], u1[8], u2[8];
    bool ok = ok_out[idx] != 0;
    load_be32(r, sig + 64 * idx);
    load_be32(sv, sig + 64 * idx + 32);
    load_be32(z, zb + 32 * idx);
    if (!ok) {
#pragma unroll
      for (int i = 0; i < 8; i++) sv[i] = (i == 0);
    }
    N::to_mont(sm, sv);
    soa_load<8>(pre, scr, n, idx, 0);
    N::mul(w, inv, pre);   // w = s^-1 (Montgomery form)
    N::mul(inv, inv, sm);
    N::cond_sub_n(z, N::ge_n(z));  // bits2field + reduce for 256-bit curves
    // u1 = z*w, u2 = r*w: mont_mul(plain, mont) = plain product
    N::mul(u1, z, w);
    N::mul(u2, r, w);
    if (!ok) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        u1[i] = (i == 0);
        u2[i] = (i == 0);
      }
      Aff G;
      C::generator(G);
      Fe gx, gy;
      F::to_canonical(gx, G.x);
      F::to_canonical(gy, G.y);
      store_be32(pxy + 64 * idx, gx.v);
      store_be32(pxy + 64 * idx + 32, gy.v);
    } else {
#pragma unroll
      for (int i = 0; i < 16; i++) reinterpret_cast<uint32_t*>(pxy + 64 * idx)[i] = reinterpret_cast<const uint32_t*>(qxy + 64 * idx)[i];
    }
    store_scalar_be(a_out + 32 * idx, u1);
    store_scalar_be(b_out + 32 * idx, u2);
    if (idx < T) break;
  }
}
template <class C>
__global__ void __launch_bounds__(256)
    ecdsa_check_kernel(const uint8_t* __restrict__ sig, const uint8_t* __restrict__ rxy, const uint8_t* __restrict__ rinf,
                       const uint8_t* __restrict__ ok, size_t n, uint8_t* __restrict__ valid) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  uint32_t r[8], x[8];
  load_be32(r, sig + 64 * idx);
  load_be32(x, rxy + 64 * idx);
  FnMont<C>::cond_sub_n(x, FnMont<C>::ge_n(x));  // x(R) mod n  (p < 2n)
  bool same = true;
#pragma unroll
  for (int i = 0; i < 8; i++) same = same && (r[i] == x[i]);
  valid[idx] = (ok[idx] && !rinf[idx] && same) ? 1 : 0;
}

// SEC1 compressed points (33 bytes: 02/03 || x; 33 zero bytes = identity) -> affine x || y, identity flag, validity.
template <class C>
__global__ void __launch_bounds__(128)
    decompress_kernel(const uint8_t* __restrict__ sec1, size_t n, uint8_t* __restrict__ out_xy, uint8_t* __restrict__ out_inf,
                      uint8_t* __restrict__ valid) {
  typedef typename C::F F;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const uint8_t* rec = sec1 + 33 * idx;
  uint8_t tag = rec[0];
  uint32_t x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {  // unaligned big-endian words
    const uint8_t* b = rec + 1 + 4 * (7 - i);
    x[i] = ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
  }
  bool zero = (x[0] | x[1] | x[2] | x[3] | x[4] | x[5] | x[6] | x[7]) == 0;
  Aff P;
  bool ok = false, inf = false;
  if (tag == 0 && zero) {
    ok = inf = true;
  } else if (tag == 2 || tag == 3) {
    ok = sec1_decompress<C>(P, x, tag & 1u);
  }
  Fe cx, cy;
  if (ok && !inf) {
    F::to_canonical(cx, P.x);
    F::to_canonical(cy, P.y);
  } else {
    F::set_zero(cx);
    F::set_zero(cy);
  }
  store_be32(out_xy + 64 * idx, cx.v);
  store_be32(out_xy + 64 * idx + 32, cy.v);
  out_inf[idx] = inf ? 1 : 0;
  valid[idx] = ok ? 1 : 0;
}

// canonical affine big-endian bytes (n*64) -> table words (internal form); used once, when a table is built
template <class C>
__global__ void __launch_bounds__(256)
    affine_to_table_kernel(const uint8_t* __restrict__ xy, size_t n, uint32_t* __restrict__ table) {
  typedef typename C::F F;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  Fe x, y;
  load_be32(x.v, xy + 64 * idx);
  load_be32(y.v, xy + 64 * idx + 32);
  F::from_canonical(x, x);
  F::from_canonical(y, y);
#pragma unroll
  for (int w = 0; w < 8; w++) {
    table[idx * 16 + w] = x.v[w];
    table[idx * 16 + 8 + w] = y.v[w];
  }
}

// Sum of Jacobian points: thread t adds elements t, t+T, t+2T, ... of `in` (SoA, n_in) and writes partial t of
// `out` (SoA, n_out = T).  Applied repeatedly until one point is left (lincomb's final reduction).
template <class C>
__global__ void __launch_bounds__(128)
    jac_sum_kernel(const uint32_t* __restrict__ in, size_t n_in, uint32_t* __restrict__ out, size_t n_out) {
  typedef typename C::F F;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_out) return;
  Jac acc;
  F::set_zero(acc.X);
  F::set_one(acc.Y);
  F::set_zero(acc.Z);
  for (size_t idx = t; idx < n_in; idx += n_out) {
    Jac p;
    soa_load<8>(p.X.v, in, n_in, idx, 0);
    soa_load<8>(p.Y.v, in, n_in, idx, 8);
    soa_load<8>(p.Z.v, in, n_in, idx, 16);
    jac_add<F, C::A_IS_MINUS3>(acc, acc, p);
  }
  soa_store<8>(out, n_out, t, acc.X.v, 0);
  soa_store<8>(out, n_out, t, acc.Y.v, 8);
  soa_store<8>(out, n_out, t, acc.Z.v, 16);
}

// SoA internal Jacobian -> AoS canonical big-endian X||Y||Z (96 bytes per point)
template <class C>
__global__ void __launch_bounds__(128)
    export_jac_kernel(const uint32_t* __restrict__ jac, size_t n, uint8_t* __restrict__ xyz) {
  typedef typename C::F F;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
#pragma unroll 1
  for (int c = 0; c < 3; c++) {
    Fe v;
    soa_load<8>(v.v, jac, n, idx, 8 * c);
    F::to_canonical(v, v);
    store_be32(xyz + 96 * idx + 32 * c, v.v);
  }
}

// ------------------------------------------------------------------------------------------------
// Jacobian (SoA) -> canonical affine bytes with Montgomery's trick along each thread's strided slice:
// thread t owns elements t, t+T, t+2T, ... ; one field inversion per thread, 7 field multiplications per
// element.  Replaces batch_normalize / BatchInvert (k256/src/arithmetic/projective.rs:367-391,
// k256/src/arithmetic/field.rs:244-291).  scr: 8*n words of scratch (prefix products).
template <class F>
__global__ void __launch_bounds__(256)
    normalize_kernel(const uint32_t* __restrict__ jac, size_t n, uint32_t* __restrict__ scr,
                     uint8_t* __restrict__ out_xy, uint8_t* __restrict__ out_inf) {
  size_t T = (size_t)gridDim.x * blockDim.x;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  Fe acc, one;
  F::set_one(one);
  acc = one;
  size_t last = t;
  for (size_t idx = t; idx < n; idx += T) {
    Fe z;
    soa_load<8>(z.v, jac, n, idx, 16);
    if (F::is_zero(z)) z = one;
    soa_store<8>(scr, n, idx, acc.v, 0);
    F::mul(acc, acc, z);
    last = idx;
  }
  Fe inv;
  F::inv(inv, acc);
  for (size_t idx = last;; idx -= T) {
    Jac p;
    soa_load<8>(p.Z.v, jac, n, idx, 16);
    bool inf = F::is_zero(p.Z);
    if (inf) p.Z = one;
    Fe pre, zinv;
    soa_load<8>(pre.v, scr, n, idx, 0);
    F::mul(zinv, inv, pre);
    F::mul(inv, inv, p.Z);
    soa_load<8>(p.X.v, jac, n, idx, 0);
    soa_load<8>(p.Y.v, jac, n, idx, 8);
    Fe x, y;
    jac_to_affine_canonical<F>(x, y, p, zinv);
    if (inf) {
      F::set_zero(x);
      F::set_zero(y);
    }
    store_be32(out_xy + 64 * idx, x.v);
    store_be32(out_xy + 64 * idx + 32, y.v);
    out_inf[idx] = inf ? 1 : 0;
    if (idx < T) break;
  }
}

// AoS big-endian X||Y||Z (n*96 bytes, canonical) -> SoA internal form; validates coordinates < p.
template <class C>
__global__ void __launch_bounds__(256)
    import_jac_kernel(const uint8_t* __restrict__ xyz, size_t n, uint32_t* __restrict__ jac,
                      uint32_t* __restrict__ status, size_t base) {
  typedef typename C::F F;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
#pragma unroll 1
  for (int c = 0; c < 3; c++) {
    Fe v, w;
    load_be32(v.v, xyz + 96 * idx + 32 * c);
    if (!lt8(v.v, C::P())) report_error(status, ERRF_POINT, base + idx);
    F::from_canonical(w, v);
    soa_store<8>(jac, n, idx, w.v, 8 * c);
  }
}

// ------------------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(256)
    field_op_kernel(int op, size_t n, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                    uint8_t* __restrict__ out, uint32_t* __restrict__ status, size_t base) {
  typedef typename C::F F;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  Fe x, y, r;
  load_be32(x.v, a + 32 * idx);
  if (!lt8(x.v, C::P())) report_error(status, ERRF_POINT, base + idx);
  F::from_canonical(x, x);
  bool binary = (op == ECG_FOP_ADD || op == ECG_FOP_SUB || op == ECG_FOP_MUL);
  if (binary) {
    load_be32(y.v, b + 32 * idx);
    if (!lt8(y.v, C::P())) report_error(status, ERRF_POINT, base + idx);
    F::from_canonical(y, y);
  } else {
    y = x;
  }
  switch (op) {
    case ECG_FOP_ADD: F::add(r, x, y); break;
    case ECG_FOP_SUB: F::sub(r, x, y); break;
    case ECG_FOP_NEG: F::neg(r, x); break;
    case ECG_FOP_MUL: F::mul(r, x, y); break;
    case ECG_FOP_SQR: F::sqr(r, x); break;
    default: F::inv(r, x); break;
  }
  F::to_canonical(r, r);
  store_be32(out + 32 * idx, r.v);
}

// ------------------------------------------------------------------------------------------------
// integer-pipe microbenchmarks (roofline denominators)
__global__ void __launch_bounds__(256) mb_imad_wide_kernel(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3 + 1, b0 = seed ^ 0x9E3779B9u, b1 = b0 + blockIdx.x;
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = a0 + i;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    // 4 independent chains of 4 IMAD.WIDE.U32.X each = 16 per iteration, x4 unrolled = 64
#pragma unroll
    for (int u = 0; u < 4; u++) {
      mad_wide_cc(r[0], r[1], a0, b0);
      madc_wide_cc(r[2], r[3], a1, b0);
      madc_wide_cc(r[4], r[5], a0, b1);
      madc_wide_cc(r[6], r[7], a1, b1);
      mad_wide_cc(r[8], r[9], a1, b0);
      madc_wide_cc(r[10], r[11], a0, b1);
      madc_wide_cc(r[12], r[13], a1, b1);
      madc_wide_cc(r[14], r[15], a0, b0);
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s ^= r[i];
  if (s == 0x12345678u) out[0] = s;
}
__global__ void __launch_bounds__(256) mb_imad_kernel(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a = seed + threadIdx.x, b = seed ^ 0x9E3779B9u;
  uint32_t r[8];
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = a + i;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < 8; i++) r[i] = r[i] * a + b;
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= r[i];
  if (s == 0x12345678u) out[0] = s;
}
__global__ void __launch_bounds__(256) mb_iadd_kernel(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a = seed + threadIdx.x;
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = a + i;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      r[0] = add_cc(r[0], r[8]);
#pragma unroll
      for (int i = 1; i < 8; i++) r[i] = addc_cc(r[i], r[8 + i]);
      r[8] = add_cc(r[8], r[1]);
#pragma unroll
      for (int i = 1; i < 8; i++) r[8 + i] = addc_cc(r[8 + i], r[(i + 1) & 7]);
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s ^= r[i];
  if (s == 0x12345678u) out[0] = s;
}
template <class F>
__global__ void __launch_bounds__(256) mb_fmul_kernel(uint32_t* out, int iters, uint32_t seed) {
  Fe a, b;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a.v[i] = seed * (i + 1) + threadIdx.x;
    b.v[i] = (seed ^ 0x9E3779B9u) * (i + 3) + blockIdx.x;
  }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    F::mul(a, a, b);
    F::mul(b, b, a);
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= a.v[i] ^ b.v[i];
  if (s == 0x12345678u) out[0] = s;
}

// ------------------------------------------------------------------------------------------------
// host side
//
// A ctx owns, per device, two "lanes" (stream + grow-only device buffers + status words).  Device-pointer
// mode uses lane 0 only (optionally on the caller's stream).  Host-pointer mode cuts every per-element batch
// into chunks and alternates lanes, so the H2D copy of chunk c+1 and the D2H copy of chunk c-1 overlap the
// kernels of chunk c (PCIe is the only thing between the caller's buffers and the SMs).
enum { B_K = 0, B_P = 1, B_INF = 2, B_JAC = 3, B_SCR = 4, B_OUT = 5, B_OINF = 6, B_AUX = 7, B_A = 8, B_JAC2 = 9, B_TAB = 10, B_MSM = 11, B_FB1 = 12, B_FB2 = 13, B_X = 14, B_V1 = 15, B_V2 = 16, B_V3 = 17, B_V4 = 18, B_V5 = 19, B_V6 = 20, B_COUNT = 21 };
static const size_t HOST_CHUNK = (size_t)1 << 18;  // elements per pipelined chunk in host-pointer mode
static const size_t DEV_CHUNK = (size_t)1 << 22;   // device-pointer mode: bound the temporaries (tables 512-768 B/element)

struct Lane {
  cudaStream_t stream = nullptr;       // owned
  cudaStream_t user_stream = nullptr;  // optional override (lane 0, ecg_ctx_set_stream)
  bool use_user_stream = false;
  void* buf[B_COUNT] = {nullptr};
  size_t cap[B_COUNT] = {0};
  uint32_t* status = nullptr;    // 2 words: error flags, smallest offending index
  uint32_t* h_status = nullptr;  // pinned
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // bracket the dominant kernel of the last call (ecg_timing)
  bool ev_pending = false;
  bool used = false;  // touched by the current call
  cudaStream_t s() const { return use_user_stream ? user_stream : stream; }
};
struct DevState {
  int dev = 0;
  Lane lane[2];
  uint32_t* fb_table[2] = {nullptr, nullptr};  // per curve, built lazily (like the reference's LazyLock table)
  int sm_count = 148;
};

struct ecg_ctx {
  std::vector<DevState> devs;
  unsigned flags = 0;
  std::string err;
  size_t err_index = (size_t)-1;
  uint64_t launches = 0;
  bool timing = false;    // ecg_timing_enable
  double dom_ms_sum = 0;  // accumulated device time of the dominant kernel (max over devices per call)
  uint64_t dom_calls = 0;
  bool devptr() const { return (flags & ECG_FLAG_DEVICE_PTRS) != 0; }
};

#define CU_TRY(ctx, call)                                                                                    \
  do {                                                                                                       \
    cudaError_t e_ = (call);                                                                                 \
    if (e_ != cudaSuccess) {                                                                                 \
      char m_[512];                                                                                          \
      snprintf(m_, sizeof m_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);  \
      (ctx)->err = m_;                                                                                       \
      return e_ == cudaErrorMemoryAllocation ? ECG_ENOMEM : ECG_ECUDA;                                       \
    }                                                                                                        \
  } while (0)
#define ST_TRY(expr)               \
  do {                             \
    ecg_status st_ = (expr);       \
    if (st_ != ECG_OK) return st_; \
  } while (0)
#define LAUNCHED(ctx)                \
  do {                               \
    (ctx)->launches++;               \
    CU_TRY(ctx, cudaGetLastError()); \
  } while (0)
// CUDA events around the dominant kernel of a call, on the launching stream (bench.py's roofline numerator)
#define DOM_BEGIN(ctx, L)                                              \
  do {                                                                 \
    if ((ctx)->timing) CU_TRY(ctx, cudaEventRecord((L).ev0, (L).s())); \
  } while (0)
#define DOM_END(ctx, L)                                  \
  do {                                                   \
    if ((ctx)->timing) {                                 \
      CU_TRY(ctx, cudaEventRecord((L).ev1, (L).s()));    \
      (L).ev_pending = true;                             \
    }                                                    \
  } while (0)

static ecg_status ensure(ecg_ctx* ctx, Lane& L, int which, size_t bytes) {
  if (bytes <= L.cap[which]) return ECG_OK;
  if (L.buf[which]) {
    CU_TRY(ctx, cudaStreamSynchronize(L.s()));
    CU_TRY(ctx, cudaFree(L.buf[which]));
  }
  L.buf[which] = nullptr;
  L.cap[which] = 0;
  size_t want = bytes + bytes / 8 + 256;
  CU_TRY(ctx, cudaMalloc(&L.buf[which], want));
  L.cap[which] = want;
  return ECG_OK;
}

extern "C" const char* ecg_version(void) { return "ecgpu 0.3 (sm_100a)"; }

extern "C" ecg_status ecg_ctx_create(const int* device_ids, int n_devices, unsigned flags, ecg_ctx** out) {
  if (!out) return ECG_EINVAL;
  *out = nullptr;
  if (n_devices < 0 || n_devices > 64) return ECG_EINVAL;
  if ((flags & ECG_FLAG_DEVICE_PTRS) && n_devices > 1) return ECG_EINVAL;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return ECG_ECUDA;  // no CPU fallback, by design
  ecg_ctx* ctx = new ecg_ctx();
  ctx->flags = flags;
  int nd = n_devices > 0 ? n_devices : 1;
  ctx->devs.resize(nd);
  for (int i = 0; i < nd; i++) {
    DevState& d = ctx->devs[i];
    d.dev = (device_ids && n_devices > 0) ? device_ids[i] : 0;
    if (d.dev < 0 || d.dev >= count) {
      ecg_ctx_destroy(ctx);
      return ECG_EINVAL;
    }
    bool ok = cudaSetDevice(d.dev) == cudaSuccess;
    for (int l = 0; ok && l < 2; l++) {
      Lane& L = d.lane[l];
      ok = cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking) == cudaSuccess &&
           cudaMalloc((void**)&L.status, 8) == cudaSuccess && cudaMallocHost((void**)&L.h_status, 8) == cudaSuccess &&
           cudaEventCreate(&L.ev0) == cudaSuccess && cudaEventCreate(&L.ev1) == cudaSuccess;
    }
    if (!ok) {
      ecg_ctx_destroy(ctx);
      return ECG_ECUDA;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, d.dev) == cudaSuccess) d.sm_count = prop.multiProcessorCount;
  }
  *out = ctx;
  return ECG_OK;
}

extern "C" void ecg_ctx_destroy(ecg_ctx* ctx) {
  if (!ctx) return;
  for (DevState& d : ctx->devs) {
    cudaSetDevice(d.dev);
    for (int l = 0; l < 2; l++) {
      Lane& L = d.lane[l];
      if (L.stream) {
        cudaStreamSynchronize(L.stream);
        cudaStreamDestroy(L.stream);
      }
      for (int i = 0; i < B_COUNT; i++)
        if (L.buf[i]) cudaFree(L.buf[i]);
      if (L.status) cudaFree(L.status);
      if (L.h_status) cudaFreeHost(L.h_status);
      if (L.ev0) cudaEventDestroy(L.ev0);
      if (L.ev1) cudaEventDestroy(L.ev1);
    }
    for (int i = 0; i < 2; i++)
      if (d.fb_table[i]) cudaFree(d.fb_table[i]);
  }
  delete ctx;
}

extern "C" const char* ecg_last_error(const ecg_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
extern "C" size_t ecg_last_error_index(const ecg_ctx* ctx) { return ctx ? ctx->err_index : (size_t)-1; }
extern "C" uint64_t ecg_kernel_launches(const ecg_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" ecg_status ecg_timing_enable(ecg_ctx* ctx, int on) {
  if (!ctx) return ECG_EINVAL;
  ctx->timing = on != 0;
  ctx->dom_ms_sum = 0;
  ctx->dom_calls = 0;
  return ECG_OK;
}
extern "C" ecg_status ecg_timing_read(const ecg_ctx* ctx, double* dominant_kernel_ms_sum, uint64_t* calls) {
  if (!ctx || !dominant_kernel_ms_sum || !calls) return ECG_EINVAL;
  *dominant_kernel_ms_sum = ctx->dom_ms_sum;
  *calls = ctx->dom_calls;
  return ECG_OK;
}

extern "C" ecg_status ecg_ctx_set_stream(ecg_ctx* ctx, void* cuda_stream) {
  if (!ctx) return ECG_EINVAL;
  Lane& L = ctx->devs[0].lane[0];
  L.user_stream = (cudaStream_t)cuda_stream;
  L.use_user_stream = cuda_stream != nullptr;
  return ECG_OK;
}

// Shard [0,n) into contiguous per-device ranges.
struct Shard {
  size_t off, cnt;
};
static std::vector<Shard> make_shards(size_t n, size_t ndev) {
  std::vector<Shard> v(ndev);
  size_t base = n / ndev, rem = n % ndev, off = 0;
  for (size_t i = 0; i < ndev; i++) {
    size_t c = base + (i < rem ? 1 : 0);
    v[i] = {off, c};
    off += c;
  }
  return v;
}

// operands of one chunk as seen by the kernels
struct DevPtrs {
  const uint8_t *k = nullptr, *p = nullptr, *inf = nullptr, *a = nullptr;
  uint8_t *out = nullptr, *oinf = nullptr;
};

// Make elements [off, off+cnt) of `src` (stride bytes each) available on the lane: in device-pointer mode that is
// pointer arithmetic, in host mode an async H2D copy into the lane's slot.
static ecg_status stage_in(ecg_ctx* ctx, Lane& L, int slot, const uint8_t* src, size_t off, size_t cnt, size_t stride,
                           const uint8_t** dst) {
  if (!src) {
    *dst = nullptr;
    return ECG_OK;
  }
  if (ctx->devptr()) {
    *dst = src + off * stride;
    return ECG_OK;
  }
  ST_TRY(ensure(ctx, L, slot, cnt * stride));
  CU_TRY(ctx, cudaMemcpyAsync(L.buf[slot], src + off * stride, cnt * stride, cudaMemcpyHostToDevice, L.s()));
  *dst = (const uint8_t*)L.buf[slot];
  return ECG_OK;
}
static ecg_status stage_out(ecg_ctx* ctx, Lane& L, size_t off, size_t cnt, uint8_t* out, size_t ostride, uint8_t* oinf,
                            DevPtrs& dp) {
  if (ctx->devptr()) {
    dp.out = out + off * ostride;
    dp.oinf = oinf ? oinf + off : nullptr;
    if (!dp.oinf) {
      ST_TRY(ensure(ctx, L, B_OINF, cnt));
      dp.oinf = (uint8_t*)L.buf[B_OINF];
    }
    return ECG_OK;
  }
  ST_TRY(ensure(ctx, L, B_OUT, cnt * ostride));
  ST_TRY(ensure(ctx, L, B_OINF, cnt));
  dp.out = (uint8_t*)L.buf[B_OUT];
  dp.oinf = (uint8_t*)L.buf[B_OINF];
  return ECG_OK;
}
static ecg_status copy_back(ecg_ctx* ctx, Lane& L, size_t off, size_t cnt, uint8_t* out, size_t ostride, uint8_t* oinf,
                            const DevPtrs& dp) {
  if (ctx->devptr()) return ECG_OK;
  CU_TRY(ctx, cudaMemcpyAsync(out + off * ostride, dp.out, cnt * ostride, cudaMemcpyDeviceToHost, L.s()));
  if (oinf) CU_TRY(ctx, cudaMemcpyAsync(oinf + off, dp.oinf, cnt, cudaMemcpyDeviceToHost, L.s()));
  return ECG_OK;
}
static ecg_status begin_lane(ecg_ctx* ctx, Lane& L) {
  if (L.used) return ECG_OK;
  CU_TRY(ctx, cudaMemsetAsync(L.status, 0, 4, L.s()));
  CU_TRY(ctx, cudaMemsetAsync(L.status + 1, 0xFF, 4, L.s()));
  L.used = true;
  return ECG_OK;
}
// Wait for every lane touched by this call; fold validation status and kernel timing into the ctx.
static ecg_status finish(ecg_ctx* ctx) {
  ecg_status rc = ECG_OK;
  size_t first = (size_t)-1;
  float dom_ms = 0;
  bool any_timed = false;
  for (DevState& d : ctx->devs) {
    float dev_ms = 0;
    for (int l = 0; l < 2; l++) {
      Lane& L = d.lane[l];
      if (!L.used) continue;
      L.used = false;
      CU_TRY(ctx, cudaSetDevice(d.dev));
      CU_TRY(ctx, cudaMemcpyAsync(L.h_status, L.status, 8, cudaMemcpyDeviceToHost, L.s()));
      CU_TRY(ctx, cudaStreamSynchronize(L.s()));
      CU_TRY(ctx, cudaGetLastError());
      if (L.ev_pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, L.ev0, L.ev1) == cudaSuccess) dev_ms += ms;
        L.ev_pending = false;
        any_timed = true;
      }
      if (L.h_status[0] && (size_t)L.h_status[1] < first) {
        first = L.h_status[1];
        rc = (L.h_status[0] & ERRF_POINT) ? ECG_ENOT_ON_CURVE : ECG_ESCALAR_RANGE;
      }
    }
    if (dev_ms > dom_ms) dom_ms = dev_ms;
  }
  if (any_timed) {
    ctx->dom_ms_sum += dom_ms;
    ctx->dom_calls++;
  }
  if (rc != ECG_OK) {
    ctx->err_index = first;
    ctx->err = rc == ECG_ESCALAR_RANGE ? "scalar out of range (>= n)" : "point not on curve / coordinate >= p";
  }
  return rc;
}
// abandon a call after a host-side failure: drain what was enqueued so buffers can be reused
static ecg_status fail(ecg_ctx* ctx, ecg_status rc) {
  std::string saved = ctx->err;
  for (DevState& d : ctx->devs)
    for (int l = 0; l < 2; l++)
      if (d.lane[l].used) {
        cudaSetDevice(d.dev);
        cudaStreamSynchronize(d.lane[l].s());
        d.lane[l].used = false;
        d.lane[l].ev_pending = false;
      }
  ctx->err = saved;
  return rc;
}

static inline unsigned grid_for(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

template <class F>
static ecg_status launch_normalize(ecg_ctx* ctx, DevState& d, Lane& L, size_t n, const uint32_t* jac, uint8_t* out, uint8_t* oinf) {
  ST_TRY(ensure(ctx, L, B_SCR, n * 32));
  // ~32 elements per thread amortise the per-thread inversion, but never leave SMs idle for small batches
  size_t want_threads = std::max<size_t>((n + 31) / 32, std::min<size_t>(n, (size_t)d.sm_count * 256));
  normalize_kernel<F><<<grid_for(want_threads, 256), 256, 0, L.s()>>>(jac, n, (uint32_t*)L.buf[B_SCR], out, oinf);
  LAUNCHED(ctx);
  return ECG_OK;
}
static ecg_status launch_norm(ecg_ctx* ctx, DevState& d, Lane& L, ecg_curve curve, size_t n, const uint32_t* jac, uint8_t* out,
                              uint8_t* oinf) {
  return curve == ECG_SECP256K1 ? launch_normalize<FpK256>(ctx, d, L, n, jac, out, oinf)
                                : launch_normalize<FpP256>(ctx, d, L, n, jac, out, oinf);
}

// launch geometry of the variable-base kernels (registers set the occupancy; tables are in global memory)
static const int K_BLOCK = 128, K_MINBLK = 4;  // secp256k1: <= 128 registers -> 16 warps/SM (mul = call, sqr inlined: OPT 7)
static const int P_BLOCK = 128, P_MINBLK = 4;  // P-256   : <= 128 registers -> 16 warps/SM

// per-block window-table slots for a launch of n elements
static ecg_status ensure_tab(ecg_ctx* ctx, Lane& L, ecg_curve curve, size_t n) {
  size_t block = curve == ECG_SECP256K1 ? K_BLOCK : P_BLOCK;
  size_t words = curve == ECG_SECP256K1 ? K_TAB_WORDS : P_TAB_WORDS;
  size_t blocks = (n + block - 1) / block;
  return ensure(ctx, L, B_TAB, blocks * block * words * 4);
}

// k*P for one chunk -> Jacobian SoA in `jac`; `status` / `base` locate validation errors
static ecg_status launch_varbase(ecg_ctx* ctx, Lane& L, ecg_curve curve, size_t n, const DevPtrs& dp, uint32_t* jac,
                                 uint32_t* status, size_t base) {
  ST_TRY(ensure_tab(ctx, L, curve, n));
  uint32_t* gtab = (uint32_t*)L.buf[B_TAB];
  DOM_BEGIN(ctx, L);
  if (curve == ECG_SECP256K1)
    k256_varbase_kernel<K_BLOCK, K_MINBLK><<<grid_for(n, K_BLOCK), K_BLOCK, 0, L.s()>>>(dp.k, dp.p, dp.inf, n, jac, gtab, status, base);
  else
    generic_varbase_kernel<CurveP256, P_BLOCK, P_MINBLK><<<grid_for(n, P_BLOCK), P_BLOCK, 0, L.s()>>>(dp.k, dp.p, dp.inf, n, jac, gtab, status, base);
  LAUNCHED(ctx);
  DOM_END(ctx, L);
  return ECG_OK;
}

static bool curve_ok(ecg_curve c) { return c == ECG_SECP256K1 || c == ECG_NISTP256; }

// ---- fixed-base table ------------------------------------------------------------------------------
// Built on the device with the variable-base kernel itself: entry (i, j) = ((2j+1) << 16 i mod n) * G.
static void scalar_be_from_shifted(uint8_t* out, uint64_t odd, int shift_bits, const uint32_t* n_le) {
  // v = odd << shift_bits  (< 2^257), reduced once by n (v < 2n always holds: n > 2^255 for both curves)
  uint32_t v[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  int w = shift_bits / 32, b = shift_bits % 32;
  uint64_t lo = odd << b;  // odd < 2^17, b < 32
  v[w] = (uint32_t)lo;
  if (w + 1 < 10) v[w + 1] = (uint32_t)(lo >> 32);
  uint32_t nn[9];
  for (int i = 0; i < 8; i++) nn[i] = n_le[i];
  nn[8] = 0;
  bool ge = true;
  for (int i = 8; i >= 0; i--) {
    if (v[i] != nn[i]) {
      ge = v[i] > nn[i];
      break;
    }
  }
  if (ge) {
    uint64_t borrow = 0;
    for (int i = 0; i < 9; i++) {
      uint64_t t = (uint64_t)v[i] - nn[i] - borrow;
      v[i] = (uint32_t)t;
      borrow = (t >> 63) & 1;
    }
  }
  for (int i = 0; i < 8; i++) {
    out[31 - 4 * i] = (uint8_t)v[i];
    out[30 - 4 * i] = (uint8_t)(v[i] >> 8);
    out[29 - 4 * i] = (uint8_t)(v[i] >> 16);
    out[28 - 4 * i] = (uint8_t)(v[i] >> 24);
  }
}
static const uint32_t H_K256_N[8] = {0xD0364141u, 0xBFD25E8Cu, 0xAF48A03Bu, 0xBAAEDCE6u, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
static const uint32_t H_P256_N[8] = {0xFC632551u, 0xF3B9CAC2u, 0xA7179E84u, 0xBCE6FAADu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x00000000u, 0xFFFFFFFFu};
static const uint8_t H_K256_G[64] = {
    0x79, 0xBE, 0x66, 0x7E, 0xF9, 0xDC, 0xBB, 0xAC, 0x55, 0xA0, 0x62, 0x95, 0xCE, 0x87, 0x0B, 0x07, 0x02, 0x9B, 0xFC, 0xDB, 0x2D, 0xCE,
    0x28, 0xD9, 0x59, 0xF2, 0x81, 0x5B, 0x16, 0xF8, 0x17, 0x98, 0x48, 0x3A, 0xDA, 0x77, 0x26, 0xA3, 0xC4, 0x65, 0x5D, 0xA4, 0xFB, 0xFC,
    0x0E, 0x11, 0x08, 0xA8, 0xFD, 0x17, 0xB4, 0x48, 0xA6, 0x85, 0x54, 0x19, 0x9C, 0x47, 0xD0, 0x8F, 0xFB, 0x10, 0xD4, 0xB8};
static const uint8_t H_P256_G[64] = {
    0x6B, 0x17, 0xD1, 0xF2, 0xE1, 0x2C, 0x42, 0x47, 0xF8, 0xBC, 0xE6, 0xE5, 0x63, 0xA4, 0x40, 0xF2, 0x77, 0x03, 0x7D, 0x81, 0x2D, 0xEB,
    0x33, 0xA0, 0xF4, 0xA1, 0x39, 0x45, 0xD8, 0x98, 0xC2, 0x96, 0x4F, 0xE3, 0x42, 0xE2, 0xFE, 0x1A, 0x7F, 0x9B, 0x8E, 0xE7, 0xEB, 0x4A,
    0x7C, 0x0F, 0x9E, 0x16, 0x2B, 0xCE, 0x33, 0x57, 0x6B, 0x31, 0x5E, 0xCE, 0xCB, 0xB6, 0x40, 0x68, 0x37, 0xBF, 0x51, 0xF5};

static ecg_status ensure_fb_table(ecg_ctx* ctx, DevState& d, ecg_curve curve) {
  if (d.fb_table[curve]) return ECG_OK;
  Lane& L = d.lane[0];
  const size_t np = FB_TABLE_POINTS;
  std::vector<uint8_t> hk(np * 32), hp(np * 64);
  const uint32_t* n_le = curve == ECG_SECP256K1 ? H_K256_N : H_P256_N;
  const uint8_t* g = curve == ECG_SECP256K1 ? H_K256_G : H_P256_G;
  for (int i = 0; i < FB_WINDOWS; i++)
    for (uint32_t j = 0; j < FB_ENTRIES; j++) scalar_be_from_shifted(&hk[((size_t)i * FB_ENTRIES + j) * 32], 2ull * j + 1, FB_W * i, n_le);
  scalar_be_from_shifted(&hk[(np - 1) * 32], 1, 256, n_le);  // 2^256 mod n
  for (size_t i = 0; i < np; i++) memcpy(&hp[i * 64], g, 64);
  uint8_t *dk = nullptr, *dpnt = nullptr, *dxy = nullptr, *dinf = nullptr;
  uint32_t *jac = nullptr, *scr = nullptr, *table = nullptr, *st = nullptr;
  CU_TRY(ctx, cudaMalloc((void**)&dk, np * 32));
  CU_TRY(ctx, cudaMalloc((void**)&dpnt, np * 64));
  CU_TRY(ctx, cudaMalloc((void**)&dxy, np * 64));
  CU_TRY(ctx, cudaMalloc((void**)&dinf, np));
  CU_TRY(ctx, cudaMalloc((void**)&jac, np * 96));
  CU_TRY(ctx, cudaMalloc((void**)&scr, np * 32));
  CU_TRY(ctx, cudaMalloc((void**)&table, np * 64));
  CU_TRY(ctx, cudaMalloc((void**)&st, 8));  // private status: building the table must not disturb a caller's validation state
  CU_TRY(ctx, cudaMemsetAsync(st, 0, 8, L.s()));
  CU_TRY(ctx, cudaMemcpyAsync(dk, hk.data(), np * 32, cudaMemcpyHostToDevice, L.s()));
  CU_TRY(ctx, cudaMemcpyAsync(dpnt, hp.data(), np * 64, cudaMemcpyHostToDevice, L.s()));
  DevPtrs dp;
  dp.k = dk;
  dp.p = dpnt;
  bool saved_timing = ctx->timing;
  ctx->timing = false;
  ecg_status rc = launch_varbase(ctx, L, curve, np, dp, jac, st, 0);
  ctx->timing = saved_timing;
  if (rc != ECG_OK) return rc;
  size_t want_threads = std::max<size_t>((np + 31) / 32, std::min<size_t>(np, (size_t)d.sm_count * 256));
  if (curve == ECG_SECP256K1) {
    normalize_kernel<FpK256><<<grid_for(want_threads, 256), 256, 0, L.s()>>>(jac, np, scr, dxy, dinf);
    LAUNCHED(ctx);
    affine_to_table_kernel<CurveK256><<<grid_for(np, 256), 256, 0, L.s()>>>(dxy, np, table);
  } else {
    normalize_kernel<FpP256><<<grid_for(want_threads, 256), 256, 0, L.s()>>>(jac, np, scr, dxy, dinf);
    LAUNCHED(ctx);
    affine_to_table_kernel<CurveP256><<<grid_for(np, 256), 256, 0, L.s()>>>(dxy, np, table);
  }
  LAUNCHED(ctx);
  CU_TRY(ctx, cudaStreamSynchronize(L.s()));
  cudaFree(dk);
  cudaFree(dpnt);
  cudaFree(dxy);
  cudaFree(dinf);
  cudaFree(jac);
  cudaFree(scr);
  cudaFree(st);
  // lane 1 (and any caller stream) may use the table from now on: it was completed with a full synchronize
  d.fb_table[curve] = table;
  return ECG_OK;
}

// ---- per-element batch driver ------------------------------------------------------------------------
// What one chunk does is the only thing that differs between ecg_mul_batch, ecg_mul_gen_batch, ecg_mul_gen_add_batch,
// ecg_batch_normalize and ecg_field_op_batch:
struct BatchOp {
  enum Kind { MUL, MULGEN, MULGENADD, NORMALIZE, FIELD, SCHNORR, ECDSA, DECOMPRESS } kind;
  ecg_curve curve;
  int fop = 0;  // field op, or the ECDSA low-S flag
  const uint8_t *k = nullptr, *a = nullptr, *p = nullptr, *inf = nullptr;  // host or device, per ctx flags
  const uint8_t* x = nullptr;  // extra 64-byte-stride input (ECDSA public keys)
  size_t pstride = 64;
  uint8_t *out = nullptr, *oinf = nullptr;
  uint8_t* aux_out = nullptr;  // third output array (decompress: validity flags)
  size_t ostride = 64;
};

static ecg_status run_chunk(ecg_ctx* ctx, DevState& d, Lane& L, const BatchOp& op, size_t off, size_t cnt) {
  DevPtrs dp;
  ST_TRY(begin_lane(ctx, L));
  ST_TRY(stage_in(ctx, L, B_K, op.k, off, cnt, 32, &dp.k));
  ST_TRY(stage_in(ctx, L, B_A, op.a, off, cnt, 32, &dp.a));
  ST_TRY(stage_in(ctx, L, B_P, op.p, off, cnt, op.pstride, &dp.p));
  ST_TRY(stage_in(ctx, L, B_INF, op.inf, off, cnt, 1, &dp.inf));
  const uint8_t* dx = nullptr;
  ST_TRY(stage_in(ctx, L, B_X, op.x, off, cnt, 64, &dx));
  ST_TRY(stage_out(ctx, L, off, cnt, op.out, op.ostride, op.oinf, dp));
  if (op.kind == BatchOp::DECOMPRESS) {
    // out = xy (64 B), oinf = identity flags, valid flags go to a third host array staged through B_V4
    ST_TRY(ensure(ctx, L, B_V4, cnt));
    uint8_t* vvalid = (ctx->devptr() && op.aux_out) ? op.aux_out + off : (uint8_t*)L.buf[B_V4];
    if (op.curve == ECG_SECP256K1)
      decompress_kernel<CurveK256><<<grid_for(cnt, 128), 128, 0, L.s()>>>(dp.p, cnt, dp.out, dp.oinf, vvalid);
    else
      decompress_kernel<CurveP256><<<grid_for(cnt, 128), 128, 0, L.s()>>>(dp.p, cnt, dp.out, dp.oinf, vvalid);
    LAUNCHED(ctx);
    if (!ctx->devptr() && op.aux_out) CU_TRY(ctx, cudaMemcpyAsync(op.aux_out + off, vvalid, cnt, cudaMemcpyDeviceToHost, L.s()));
    return copy_back(ctx, L, off, cnt, op.out, op.ostride, op.oinf, dp);
  }
  if (op.kind == BatchOp::SCHNORR || op.kind == BatchOp::ECDSA) {
    // front end -> (a, b, P) -> a*G + b*P -> affine -> verdict, all on the device
    ST_TRY(ensure(ctx, L, B_V1, cnt * 64));
    ST_TRY(ensure(ctx, L, B_V2, cnt * 32));
    ST_TRY(ensure(ctx, L, B_V3, cnt * 32));
    ST_TRY(ensure(ctx, L, B_V4, cnt));
    ST_TRY(ensure(ctx, L, B_V5, cnt * 64));
    ST_TRY(ensure(ctx, L, B_V6, cnt));
    ST_TRY(ensure(ctx, L, B_JAC, cnt * 96));
    ST_TRY(ensure_tab(ctx, L, op.curve, cnt));
    uint8_t *vp = (uint8_t*)L.buf[B_V1], *va = (uint8_t*)L.buf[B_V2], *vb = (uint8_t*)L.buf[B_V3], *vok = (uint8_t*)L.buf[B_V4];
    uint8_t *vxy = (uint8_t*)L.buf[B_V5], *vinf = (uint8_t*)L.buf[B_V6];
    uint32_t* vj = (uint32_t*)L.buf[B_JAC];
    const bool k1c = op.curve == ECG_SECP256K1;
    if (op.kind == BatchOp::SCHNORR) {
      schnorr_prep_kernel<<<grid_for(cnt, 128), 128, 0, L.s()>>>(dp.k, dp.a, dp.p, cnt, vp, va, vb, vok);
    } else {
      ST_TRY(ensure(ctx, L, B_SCR, cnt * 32));
      size_t want_threads = std::max<size_t>((cnt + 31) / 32, std::min<size_t>(cnt, (size_t)d.sm_count * 128));
      if (k1c)
        ecdsa_prep_kernel<CurveK256><<<grid_for(want_threads, 128), 128, 0, L.s()>>>(dp.k, dp.p, dx, cnt, op.fop, (uint32_t*)L.buf[B_SCR], vp, va, vb, vok);
      else
        ecdsa_prep_kernel<CurveP256><<<grid_for(want_threads, 128), 128, 0, L.s()>>>(dp.k, dp.p, dx, cnt, op.fop, (uint32_t*)L.buf[B_SCR], vp, va, vb, vok);
    }
    LAUNCHED(ctx);
    DOM_BEGIN(ctx, L);
    if (k1c)
      mul_gen_add_kernel<CurveK256, K_BLOCK, K_MINBLK, true><<<grid_for(cnt, K_BLOCK), K_BLOCK, 0, L.s()>>>(
          va, vb, vp, nullptr, cnt, d.fb_table[op.curve], vj, (uint32_t*)L.buf[B_TAB], L.status, off);
    else
      mul_gen_add_kernel<CurveP256, P_BLOCK, P_MINBLK, false><<<grid_for(cnt, P_BLOCK), P_BLOCK, 0, L.s()>>>(
          va, vb, vp, nullptr, cnt, d.fb_table[op.curve], vj, (uint32_t*)L.buf[B_TAB], L.status, off);
    LAUNCHED(ctx);
    DOM_END(ctx, L);
    ST_TRY(launch_norm(ctx, d, L, op.curve, cnt, vj, vxy, vinf));
    if (op.kind == BatchOp::SCHNORR)
      schnorr_check_kernel<<<grid_for(cnt, 256), 256, 0, L.s()>>>(dp.p, vxy, vinf, vok, cnt, dp.out);
    else if (k1c)
      ecdsa_check_kernel<CurveK256><<<grid_for(cnt, 256), 256, 0, L.s()>>>(dp.p, vxy, vinf, vok, cnt, dp.out);
    else
      ecdsa_check_kernel<CurveP256><<<grid_for(cnt, 256), 256, 0, L.s()>>>(dp.p, vxy, vinf, vok, cnt, dp.out);
    LAUNCHED(ctx);
    return copy_back(ctx, L, off, cnt, op.out, op.ostride, nullptr, dp);
  }
  uint32_t* jac = nullptr;
  if (op.kind != BatchOp::FIELD) {
    ST_TRY(ensure(ctx, L, B_JAC, cnt * 96));
    jac = (uint32_t*)L.buf[B_JAC];
  }
  const bool k1 = op.curve == ECG_SECP256K1;
  switch (op.kind) {
    case BatchOp::MUL:
      ST_TRY(launch_varbase(ctx, L, op.curve, cnt, dp, jac, L.status, off));
      break;
    case BatchOp::MULGEN:
      DOM_BEGIN(ctx, L);
      if (k1)
        fixedbase_kernel<CurveK256><<<grid_for(cnt, 128), 128, 0, L.s()>>>(dp.k, cnt, d.fb_table[op.curve], jac, L.status, off);
      else
        fixedbase_kernel<CurveP256><<<grid_for(cnt, 128), 128, 0, L.s()>>>(dp.k, cnt, d.fb_table[op.curve], jac, L.status, off);
      LAUNCHED(ctx);
      DOM_END(ctx, L);
      break;
    case BatchOp::MULGENADD:
      ST_TRY(ensure_tab(ctx, L, op.curve, cnt));
      DOM_BEGIN(ctx, L);
      if (k1)
        mul_gen_add_kernel<CurveK256, K_BLOCK, K_MINBLK, true><<<grid_for(cnt, K_BLOCK), K_BLOCK, 0, L.s()>>>(
            dp.a, dp.k, dp.p, dp.inf, cnt, d.fb_table[op.curve], jac, (uint32_t*)L.buf[B_TAB], L.status, off);
      else
        mul_gen_add_kernel<CurveP256, P_BLOCK, P_MINBLK, false><<<grid_for(cnt, P_BLOCK), P_BLOCK, 0, L.s()>>>(
            dp.a, dp.k, dp.p, dp.inf, cnt, d.fb_table[op.curve], jac, (uint32_t*)L.buf[B_TAB], L.status, off);
      LAUNCHED(ctx);
      DOM_END(ctx, L);
      break;
    case BatchOp::NORMALIZE:
      if (k1)
        import_jac_kernel<CurveK256><<<grid_for(cnt, 256), 256, 0, L.s()>>>(dp.p, cnt, jac, L.status, off);
      else
        import_jac_kernel<CurveP256><<<grid_for(cnt, 256), 256, 0, L.s()>>>(dp.p, cnt, jac, L.status, off);
      LAUNCHED(ctx);
      break;
    case BatchOp::SCHNORR:
    case BatchOp::ECDSA:
    case BatchOp::DECOMPRESS:
      break;  // handled above
    case BatchOp::FIELD:
      if (k1)
        field_op_kernel<CurveK256><<<grid_for(cnt, 256), 256, 0, L.s()>>>(op.fop, cnt, dp.k, dp.a, dp.out, L.status, off);
      else
        field_op_kernel<CurveP256><<<grid_for(cnt, 256), 256, 0, L.s()>>>(op.fop, cnt, dp.k, dp.a, dp.out, L.status, off);
      LAUNCHED(ctx);
      break;
  }
  if (op.kind != BatchOp::FIELD) ST_TRY(launch_norm(ctx, d, L, op.curve, cnt, jac, dp.out, dp.oinf));
  return copy_back(ctx, L, off, cnt, op.out, op.ostride, op.oinf, dp);
}

static ecg_status run_batch(ecg_ctx* ctx, const BatchOp& op, size_t n) {
  std::vector<Shard> shards = make_shards(n, ctx->devs.size());
  bool need_table = op.kind == BatchOp::MULGEN || op.kind == BatchOp::MULGENADD || op.kind == BatchOp::SCHNORR || op.kind == BatchOp::ECDSA;
  for (size_t i = 0; i < ctx->devs.size(); i++) {
    if (shards[i].cnt == 0) continue;
    DevState& d = ctx->devs[i];
    CU_TRY(ctx, cudaSetDevice(d.dev));
    if (need_table) {
      ecg_status st = ensure_fb_table(ctx, d, op.curve);
      if (st != ECG_OK) return fail(ctx, st);
    }
  }
  if (ctx->devptr()) {
    DevState& d = ctx->devs[0];
    for (size_t lo = 0; lo < n; lo += DEV_CHUNK) {
      ecg_status st = run_chunk(ctx, d, d.lane[0], op, lo, std::min(DEV_CHUNK, n - lo));
      if (st != ECG_OK) return fail(ctx, st);
    }
    return finish(ctx);
  }
  // host mode: interleave chunks across devices and lanes so copies and kernels of different chunks overlap
  size_t maxchunks = 0;
  for (const Shard& sh : shards) maxchunks = std::max(maxchunks, (sh.cnt + HOST_CHUNK - 1) / HOST_CHUNK);
  for (size_t c = 0; c < maxchunks; c++) {
    for (size_t i = 0; i < ctx->devs.size(); i++) {
      const Shard& sh = shards[i];
      size_t lo = c * HOST_CHUNK;
      if (lo >= sh.cnt) continue;
      size_t cnt = std::min(HOST_CHUNK, sh.cnt - lo);
      DevState& d = ctx->devs[i];
      CU_TRY(ctx, cudaSetDevice(d.dev));
      ecg_status st = run_chunk(ctx, d, d.lane[c & 1], op, sh.off + lo, cnt);
      if (st != ECG_OK) return fail(ctx, st);
    }
  }
  return finish(ctx);
}

extern "C" ecg_status ecg_mul_batch(ecg_ctx* ctx, ecg_curve curve, size_t n, const uint8_t* k, const uint8_t* P_xy,
                                    const uint8_t* P_inf, uint8_t* out_xy, uint8_t* out_inf) {
  if (!ctx) return ECG_EINVAL;
  if (n == 0) return ECG_OK;
  if (!k || !P_xy || !out_xy || !curve_ok(curve)) {
    ctx->err = "ecg_mul_batch: null pointer or unknown curve";
    return ECG_EINVAL;
  }
  BatchOp op;
  op.kind = BatchOp::MUL;
  op.curve = curve;
  op.k = k;
  op.p = P_xy;
  op.inf = P_inf;
  op.out = out_xy;
  op.oinf = out_inf;
  return run_batch(ctx, op, n);
}

extern "C" ecg_status ecg_mul_gen_batch(ecg_ctx* ctx, ecg_curve curve, size_t n, const uint8_t* k, uint8_t* out_xy,
                                        uint8_t* out_inf) {
  if (!ctx) return ECG_EINVAL;
  if (n == 0) return ECG_OK;
  if (!k || !out_xy || !curve_ok(curve)) {
    ctx->err = "ecg_mul_gen_batch: null pointer or unknown curve";
    return ECG_EINVAL;
  }
  BatchOp op;
  op.kind = BatchOp::MULGEN;
  op.curve = curve;
  op.k = k;
  op.out = out_xy;
  op.oinf = out_inf;
  return run_batch(ctx, op, n);
}

extern "C" ecg_status ecg_mul_gen_add_batch(ecg_ctx* ctx, ecg_curve curve, size_t n, const uint8_t* a, const uint8_t* b,
                                            const uint8_t* P_xy, const uint8_t* P_inf, uint8_t* out_xy, uint8_t* out_inf) {
  if (!ctx) return ECG_EINVAL;
  if (n == 0) return ECG_OK;
  if (!a || !b || !P_xy || !out_xy || !curve_ok(curve)) {
    ctx->err = "ecg_mul_gen_add_batch: null pointer or unknown curve";
    return ECG_EINVAL;
  }
  BatchOp op;
  op.kind = BatchOp::MULGENADD;
  op.curve = curve;
  op.a = a;
  op.k = b;
  op.p = P_xy;
  op.inf = P_inf;
  op.out = out_xy;
  op.oinf = out_inf;
  return run_batch(ctx, op, n);
}

extern "C" ecg_status ecg_decompress_batch(ecg_ctx* ctx, ecg_curve curve, size_t n, const uint8_t* sec1_33, uint8_t* out_xy,
                                            uint8_t* out_inf, uint8_t* valid) {
  if (!ctx) return ECG_EINVAL;
  if (n == 0) return ECG_OK;
  if (!sec1_33 || !out_xy || !out_inf || !valid || !curve_ok(curve)) {
    ctx->err = "ecg_decompress_batch: null pointer or unknown curve";
    return ECG_EINVAL;
  }
  BatchOp op;
  op.kind = BatchOp::DECOMPRESS;
  op.curve = curve;
  op.p = sec1_33;
  op.pstride = 33;
  op.out = out_xy;
  op.oinf = out_inf;
  op.aux_out = valid;
  return run_batch(ctx, op, n);
}

extern "C" ecg_status ecg_schnorr_verify_batch(ecg_ctx* ctx, size_t n, const uint8_t* pk_x, const uint8_t* msg32, const uint8_t* sig64,
                                                uint8_t* valid) {
  if (!ctx) return ECG_EINVAL;
  if (n == 0) return ECG_OK;
  if (!pk_x || !msg32 || !sig64 || !valid) {
    ctx->err = "ecg_schnorr_verify_batch: null pointer";
    return ECG_EINVAL;
  }
  BatchOp op;
  op.kind = BatchOp::SCHNORR;
  op.curve = ECG_SECP256K1;
  op.k = pk_x;
  op.a = msg32;
  op.p = sig64;
  op.out = valid;
  op.ostride = 1;
  return run_batch(ctx, op, n);
}

extern "C" ecg_status ecg_ecdsa_verify_batch(ecg_ctx* ctx, ecg_curve curve, size_t n, const uint8_t* z32, const uint8_t* sig64,
                                              const uint8_t* Q_xy, int low_s_only, uint8_t* valid) {
  if (!ctx) return ECG_EINVAL;
  if (n == 0) return ECG_OK;
  if (!z32 || !sig64 || !Q_xy || !valid || !curve_ok(curve)) {
    ctx->err = "ecg_ecdsa_verify_batch: null pointer or unknown curve";
    return ECG_EINVAL;
  }
  BatchOp op;
  op.kind = BatchOp::ECDSA;
  op.curve = curve;
  op.fop = low_s_only ? 1 : 0;
  op.k = z32;
  op.p = sig64;
  op.x = Q_xy;
  op.out = valid;
  op.ostride = 1;
  return run_batch(ctx, op, n);
}

extern "C" ecg_status ecg_batch_normalize(ecg_ctx* ctx, ecg_curve curve, size_t n, const uint8_t* xyz, uint8_t* out_xy,
                                          uint8_t* out_inf) {
  if (!ctx) return ECG_EINVAL;
  if (n == 0) return ECG_OK;
  if (!xyz || !out_xy || !curve_ok(curve)) return ECG_EINVAL;
  BatchOp op;
  op.kind = BatchOp::NORMALIZE;
  op.curve = curve;
  op.p = xyz;
  op.pstride = 96;
  op.out = out_xy;
  op.oinf = out_inf;
  return run_batch(ctx, op, n);
}

extern "C" ecg_status ecg_field_op_batch(ecg_ctx* ctx, ecg_curve curve, int fop, size_t n, const uint8_t* a, const uint8_t* b,
                                         uint8_t* out) {
  if (!ctx) return ECG_EINVAL;
  if (n == 0) return ECG_OK;
  bool binary = (fop == ECG_FOP_ADD || fop == ECG_FOP_SUB || fop == ECG_FOP_MUL);
  if (!a || !out || (binary && !b) || fop < 0 || fop > ECG_FOP_INV || !curve_ok(curve)) return ECG_EINVAL;
  BatchOp op;
  op.kind = BatchOp::FIELD;
  op.curve = curve;
  op.fop = fop;
  op.k = a;
  op.a = binary ? b : nullptr;
  op.out = out;
  op.ostride = 32;
  return run_batch(ctx, op, n);
}

// ---- lincomb ----------------------------------------------------------------------------------------
// Reduce n Jacobian points (SoA in `a`) to one, ping-ponging between a and b; *result = buffer holding it.
template <class C>
static ecg_status reduce_points(ecg_ctx* ctx, Lane& L, uint32_t* a, uint32_t* b, size_t n, uint32_t** result) {
  while (n > 1) {
    size_t m = (n + 31) / 32;
    jac_sum_kernel<C><<<grid_for(m, 128), 128, 0, L.s()>>>(a, n, b, m);
    LAUNCHED(ctx);
    std::swap(a, b);
    n = m;
  }
  *result = a;
  return ECG_OK;
}
static ecg_status reduce_points_c(ecg_ctx* ctx, Lane& L, ecg_curve curve, uint32_t* a, uint32_t* b, size_t n, uint32_t** result) {
  return curve == ECG_SECP256K1 ? reduce_points<CurveK256>(ctx, L, a, b, n, result) : reduce_points<CurveP256>(ctx, L, a, b, n, result);
}

// ---- bucket-method lincomb (ecg_msm.cuh) ------------------------------------------------------------
static const size_t MSM_MIN_TERMS = (size_t)1 << 13;   // below this the per-term kernel + tree sum is faster
static const size_t MSM_MAX_TERMS_DEFAULT = (size_t)1 << 24;  // per call (32-bit list offsets); larger shards are cut
// ECG_MSM_MAX_TERMS (environment) lowers the piece size so that tests can exercise the multi-piece path cheaply
static size_t msm_max_terms() {
  const char* e = getenv("ECG_MSM_MAX_TERMS");
  if (e) {
    size_t v = (size_t)strtoull(e, nullptr, 10);
    if (v >= MSM_MIN_TERMS && v <= MSM_MAX_TERMS_DEFAULT) return v;
  }
  return MSM_MAX_TERMS_DEFAULT;
}

static MsmGeom msm_geometry(ecg_curve curve, size_t n) {
  MsmGeom g;
  bool glv = curve == ECG_SECP256K1;
  size_t nsub = glv ? 2 * n : n;
  int lg = 0;
  while (((size_t)1 << (lg + 1)) <= nsub) lg++;
  g.c = std::min(16, std::max(8, lg - 5));
  g.nbits = glv ? 128 : 256;
  g.W = (g.nbits + g.c - 1) / g.c;
  g.nbw = ((uint32_t)1 << (g.c + 1)) + 2;
  return g;
}

struct Carver {
  uint8_t* base;
  size_t off = 0;
  template <class T>
  T* take(size_t count) {
    off = (off + 255) & ~(size_t)255;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
};

// *result == nullptr on return means "input too skewed for the bucket method, use the per-term path".
template <class C, bool GLV>
static ecg_status msm_run(ecg_ctx* ctx, Lane& L, const DevPtrs& dp, size_t n, size_t base, const MsmGeom& g, uint32_t** result) {
  const size_t nsub = GLV ? 2 * n : n;
  const size_t nb = (size_t)g.W * g.nbw;
  // recursion geometry of the weighted reduction
  std::vector<size_t> lens, nchs;
  size_t len = g.nbw - 1;
  for (;;) {
    size_t nch = (len + MSM_CH - 1) / MSM_CH;
    lens.push_back(len);
    nchs.push_back(nch);
    if (nch == 1) break;
    len = nch;
  }
  const int levels = (int)lens.size();
  // carve the scratch arena (first pass sizes it, second pass hands out pointers)
  uint32_t *pts = nullptr, *count = nullptr, *cursor = nullptr, *offset = nullptr, *list = nullptr, *bkt = nullptr, *res = nullptr;
  uint32_t* blocksum = nullptr;
  int32_t* digits = nullptr;
  std::vector<uint32_t*> S(levels), X(levels);
  uint32_t* Rw = nullptr;
  for (int pass = 0; pass < 2; pass++) {
    Carver cv{pass ? (uint8_t*)L.buf[B_MSM] : nullptr};
    pts = cv.take<uint32_t>(nsub * 16);
    digits = cv.take<int32_t>(nsub * (size_t)g.W);
    count = cv.take<uint32_t>(2 * nb + 4);  // count | cursor | maxcnt, cleared together
    cursor = count ? count + nb + 1 : nullptr;
    blocksum = cv.take<uint32_t>((nb + MSM_SCAN_CHUNK - 1) / MSM_SCAN_CHUNK + 1);
    offset = cv.take<uint32_t>(nb + 1);
    list = cv.take<uint32_t>(nsub * (size_t)g.W);
    bkt = cv.take<uint32_t>(nb * 24);
    for (int l = 0; l < levels; l++) {
      S[l] = cv.take<uint32_t>((size_t)g.W * nchs[l] * 24);
      X[l] = cv.take<uint32_t>((size_t)g.W * nchs[l] * 24);
    }
    Rw = cv.take<uint32_t>((size_t)g.W * 24);
    res = cv.take<uint32_t>(24);
    if (pass == 0) ST_TRY(ensure(ctx, L, B_MSM, cv.off + 256));
  }
  uint32_t* maxcnt = count + 2 * nb + 3;
  CU_TRY(ctx, cudaMemsetAsync(count, 0, (2 * nb + 4) * 4, L.s()));
  msm_prep_kernel<C, GLV><<<grid_for(n, 128), 128, 0, L.s()>>>(dp.k, dp.p, dp.inf, n, g, pts, digits, count, L.status, base);
  LAUNCHED(ctx);
  {
    unsigned sb = (unsigned)((nb + MSM_SCAN_CHUNK - 1) / MSM_SCAN_CHUNK);
    msm_scan_partial_kernel<<<sb, MSM_SCAN_BLOCK, 0, L.s()>>>(count, nb, blocksum, maxcnt);
    LAUNCHED(ctx);
    msm_scan_top_kernel<<<1, 1024, 0, L.s()>>>(blocksum, sb, offset, nb);
    LAUNCHED(ctx);
    msm_scan_final_kernel<<<sb, MSM_SCAN_BLOCK, 0, L.s()>>>(count, nb, blocksum, offset);
    LAUNCHED(ctx);
  }
  {
    // one bucket thread adds its points serially: refuse pathologically skewed inputs (e.g. thousands of identical
    // terms) and let the caller use the per-term kernel, whose cost does not depend on the data
    CU_TRY(ctx, cudaMemcpyAsync(L.h_status, maxcnt, 4, cudaMemcpyDeviceToHost, L.s()));
    CU_TRY(ctx, cudaStreamSynchronize(L.s()));
    size_t avg = nsub / ((size_t)1 << (g.c - 1)) + 1;
    if ((size_t)L.h_status[0] > 4096 && (size_t)L.h_status[0] > 32 * avg) {
      *result = nullptr;
      return ECG_OK;
    }
  }
  msm_scatter_kernel<<<grid_for(nsub, 256), 256, 0, L.s()>>>(digits, nsub, g, offset, cursor, list);
  LAUNCHED(ctx);
  DOM_BEGIN(ctx, L);
  msm_bucket_kernel<C><<<grid_for(nb, 128), 128, 0, L.s()>>>(pts, list, offset, nb, bkt);
  LAUNCHED(ctx);
  DOM_END(ctx, L);
  // weighted reduction, level by level (ecg_msm.cuh)
  for (int l = 0; l < levels; l++) {
    const uint32_t* in = l == 0 ? bkt : S[l - 1];
    size_t n_in = l == 0 ? nb : (size_t)g.W * nchs[l - 1];
    size_t stride = l == 0 ? g.nbw : nchs[l - 1];
    size_t off = l == 0 ? 1 : 0;
    msm_wreduce_kernel<C><<<grid_for((size_t)g.W * nchs[l], 128), 128, 0, L.s()>>>(in, n_in, stride, off, lens[l], g.W, nchs[l],
                                                                                 l == 0 ? nullptr : X[l - 1], l, S[l], X[l]);
    LAUNCHED(ctx);
  }
  msm_final_kernel<C><<<1, 32, 0, L.s()>>>(X[levels - 1], S[levels - 1], g.W, g.c, levels - 1, Rw, res);
  LAUNCHED(ctx);
  *result = res;
  return ECG_OK;
}

// one shard -> one Jacobian point left in *result (SoA with n = 1, i.e. 24 consecutive words), on lane 0
static ecg_status lincomb_shard(ecg_ctx* ctx, DevState& d, ecg_curve curve, const Shard& sh, const uint8_t* k,
                                const uint8_t* P_xy, const uint8_t* P_inf, uint32_t** result) {
  Lane& L = d.lane[0];
  DevPtrs dp;
  ST_TRY(begin_lane(ctx, L));
  ST_TRY(stage_in(ctx, L, B_K, k, sh.off, sh.cnt, 32, &dp.k));
  ST_TRY(stage_in(ctx, L, B_P, P_xy, sh.off, sh.cnt, 64, &dp.p));
  ST_TRY(stage_in(ctx, L, B_INF, P_inf, sh.off, sh.cnt, 1, &dp.inf));
  if (sh.cnt >= MSM_MIN_TERMS) {
    // bucket method, in pieces of at most MSM_MAX_TERMS terms whose partial sums are added at the end
    const size_t MSM_MAX_TERMS = msm_max_terms();
    size_t pieces = (sh.cnt + MSM_MAX_TERMS - 1) / MSM_MAX_TERMS;
    ST_TRY(ensure(ctx, L, B_JAC, pieces * 96 + 96));
    ST_TRY(ensure(ctx, L, B_JAC2, pieces * 96 + 96));
    uint32_t* parts = (uint32_t*)L.buf[B_JAC];
    for (size_t pc = 0; pc < pieces; pc++) {
      size_t lo = pc * MSM_MAX_TERMS, cnt = std::min(MSM_MAX_TERMS, sh.cnt - lo);
      DevPtrs q;
      q.k = dp.k + 32 * lo;
      q.p = dp.p + 64 * lo;
      q.inf = dp.inf ? dp.inf + lo : nullptr;
      MsmGeom g = msm_geometry(curve, cnt);
      uint32_t* r1 = nullptr;
      if (curve == ECG_SECP256K1)
        ST_TRY((msm_run<CurveK256, true>(ctx, L, q, cnt, sh.off + lo, g, &r1)));
      else
        ST_TRY((msm_run<CurveP256, false>(ctx, L, q, cnt, sh.off + lo, g, &r1)));
      if (r1 == nullptr) {  // skewed input: per-term path for this piece
        ST_TRY(ensure(ctx, L, B_FB1, cnt * 96));
        ST_TRY(ensure(ctx, L, B_FB2, ((cnt + 31) / 32) * 96 + 256));
        ST_TRY(launch_varbase(ctx, L, curve, cnt, q, (uint32_t*)L.buf[B_FB1], L.status, sh.off + lo));
        ST_TRY(reduce_points_c(ctx, L, curve, (uint32_t*)L.buf[B_FB1], (uint32_t*)L.buf[B_FB2], cnt, &r1));
      }
      if (pieces == 1) {
        *result = r1;
        return ECG_OK;
      }
      // gather piece results into an SoA array of `pieces` points (24 strided 4-byte copies)
      for (int w = 0; w < 24; w++)
        CU_TRY(ctx, cudaMemcpyAsync(parts + (size_t)w * pieces + pc, r1 + w, 4, cudaMemcpyDeviceToDevice, L.s()));
    }
    return reduce_points_c(ctx, L, curve, parts, (uint32_t*)L.buf[B_JAC2], pieces, result);
  }
  ST_TRY(ensure(ctx, L, B_JAC, sh.cnt * 96));
  ST_TRY(ensure(ctx, L, B_JAC2, ((sh.cnt + 31) / 32) * 96 + 96));
  ST_TRY(launch_varbase(ctx, L, curve, sh.cnt, dp, (uint32_t*)L.buf[B_JAC], L.status, sh.off));
  return reduce_points_c(ctx, L, curve, (uint32_t*)L.buf[B_JAC], (uint32_t*)L.buf[B_JAC2], sh.cnt, result);
}

static ecg_status export_point(ecg_ctx* ctx, Lane& L, ecg_curve curve, const uint32_t* jac1, uint8_t* dev_xyz) {
  if (curve == ECG_SECP256K1)
    export_jac_kernel<CurveK256><<<1, 128, 0, L.s()>>>(jac1, 1, dev_xyz);
  else
    export_jac_kernel<CurveP256><<<1, 128, 0, L.s()>>>(jac1, 1, dev_xyz);
  LAUNCHED(ctx);
  return ECG_OK;
}

extern "C" ecg_status ecg_lincomb_partial(ecg_ctx* ctx, ecg_curve curve, size_t n, const uint8_t* k, const uint8_t* P_xy,
                                          const uint8_t* P_inf, uint8_t* out_xyz) {
  if (!ctx) return ECG_EINVAL;
  if (!out_xyz || !curve_ok(curve) || (n > 0 && (!k || !P_xy))) {
    ctx->err = "ecg_lincomb_partial: null pointer or unknown curve";
    return ECG_EINVAL;
  }
  if (ctx->devs.size() != 1) {
    ctx->err = "ecg_lincomb_partial: single-device ctx only (use ecg_lincomb for a multi-device ctx)";
    return ECG_EINVAL;
  }
  DevState& d = ctx->devs[0];
  Lane& L = d.lane[0];
  CU_TRY(ctx, cudaSetDevice(d.dev));
  if (n == 0) {  // empty sum = identity (0 : 1 : 0)
    uint8_t z[96];
    memset(z, 0, sizeof z);
    z[63] = 1;
    if (ctx->devptr())
      CU_TRY(ctx, cudaMemcpy(out_xyz, z, 96, cudaMemcpyHostToDevice));
    else
      memcpy(out_xyz, z, 96);
    return ECG_OK;
  }
  Shard sh = {0, n};
  uint32_t* res = nullptr;
  ecg_status st = lincomb_shard(ctx, d, curve, sh, k, P_xy, P_inf, &res);
  if (st != ECG_OK) return fail(ctx, st);
  uint8_t* dst = out_xyz;
  if (!ctx->devptr()) {
    if ((st = ensure(ctx, L, B_AUX, 256)) != ECG_OK) return fail(ctx, st);
    dst = (uint8_t*)L.buf[B_AUX];
  }
  if ((st = export_point(ctx, L, curve, res, dst)) != ECG_OK) return fail(ctx, st);
  if (!ctx->devptr()) CU_TRY(ctx, cudaMemcpyAsync(out_xyz, dst, 96, cudaMemcpyDeviceToHost, L.s()));
  return finish(ctx);
}

// m Jacobian points as host bytes -> affine sum (device 0 of the ctx, lane 0)
static ecg_status point_sum_host(ecg_ctx* ctx, ecg_curve curve, size_t m, const uint8_t* xyz, uint8_t* out_xy, uint8_t* out_inf) {
  DevState& d = ctx->devs[0];
  Lane& L = d.lane[0];
  CU_TRY(ctx, cudaSetDevice(d.dev));
  ST_TRY(begin_lane(ctx, L));
  ST_TRY(ensure(ctx, L, B_AUX, m * 96 + 256));
  ST_TRY(ensure(ctx, L, B_JAC, m * 96 + 96));
  ST_TRY(ensure(ctx, L, B_JAC2, ((m + 31) / 32) * 96 + 96));
  ST_TRY(ensure(ctx, L, B_OUT, 64));
  ST_TRY(ensure(ctx, L, B_OINF, 1));
  uint8_t* dxyz = (uint8_t*)L.buf[B_AUX];
  CU_TRY(ctx, cudaMemcpyAsync(dxyz, xyz, m * 96, cudaMemcpyHostToDevice, L.s()));
  uint32_t* jac = (uint32_t*)L.buf[B_JAC];
  if (curve == ECG_SECP256K1)
    import_jac_kernel<CurveK256><<<grid_for(m, 256), 256, 0, L.s()>>>(dxyz, m, jac, L.status, 0);
  else
    import_jac_kernel<CurveP256><<<grid_for(m, 256), 256, 0, L.s()>>>(dxyz, m, jac, L.status, 0);
  LAUNCHED(ctx);
  uint32_t* res = nullptr;
  ST_TRY(reduce_points_c(ctx, L, curve, jac, (uint32_t*)L.buf[B_JAC2], m, &res));
  ST_TRY(launch_norm(ctx, d, L, curve, 1, res, (uint8_t*)L.buf[B_OUT], (uint8_t*)L.buf[B_OINF]));
  CU_TRY(ctx, cudaMemcpyAsync(out_xy, L.buf[B_OUT], 64, cudaMemcpyDeviceToHost, L.s()));
  CU_TRY(ctx, cudaMemcpyAsync(out_inf, L.buf[B_OINF], 1, cudaMemcpyDeviceToHost, L.s()));
  return finish(ctx);
}

extern "C" ecg_status ecg_point_sum(ecg_ctx* ctx, ecg_curve curve, size_t m, const uint8_t* xyz, uint8_t* out_xy,
                                    uint8_t* out_inf) {
  if (!ctx) return ECG_EINVAL;
  if (!out_xy || !out_inf || !curve_ok(curve) || (m > 0 && !xyz)) return ECG_EINVAL;
  if (m == 0) {
    memset(out_xy, 0, 64);
    *out_inf = 1;
    return ECG_OK;
  }
  ecg_status st = point_sum_host(ctx, curve, m, xyz, out_xy, out_inf);
  return st == ECG_OK ? st : fail(ctx, st);
}

extern "C" ecg_status ecg_lincomb(ecg_ctx* ctx, ecg_curve curve, size_t n, const uint8_t* k, const uint8_t* P_xy,
                                  const uint8_t* P_inf, uint8_t* out_xy, uint8_t* out_inf) {
  if (!ctx) return ECG_EINVAL;
  if (!out_xy || !out_inf || !curve_ok(curve) || (n > 0 && (!k || !P_xy))) {
    ctx->err = "ecg_lincomb: null pointer or unknown curve";
    return ECG_EINVAL;
  }
  if (n == 0) {
    uint8_t z[65];
    memset(z, 0, sizeof z);
    if (ctx->devptr()) {
      CU_TRY(ctx, cudaMemcpy(out_xy, z, 64, cudaMemcpyHostToDevice));
      z[0] = 1;
      CU_TRY(ctx, cudaMemcpy(out_inf, z, 1, cudaMemcpyHostToDevice));
    } else {
      memcpy(out_xy, z, 64);
      *out_inf = 1;
    }
    return ECG_OK;
  }
  size_t nd = ctx->devs.size();
  std::vector<Shard> shards = make_shards(n, nd);
  ecg_status st;
  if (nd == 1) {
    DevState& d = ctx->devs[0];
    Lane& L = d.lane[0];
    CU_TRY(ctx, cudaSetDevice(d.dev));
    uint32_t* res = nullptr;
    if ((st = lincomb_shard(ctx, d, curve, shards[0], k, P_xy, P_inf, &res)) != ECG_OK) return fail(ctx, st);
    DevPtrs dp;
    if ((st = stage_out(ctx, L, 0, 1, out_xy, 64, out_inf, dp)) != ECG_OK) return fail(ctx, st);
    if ((st = launch_norm(ctx, d, L, curve, 1, res, dp.out, dp.oinf)) != ECG_OK) return fail(ctx, st);
    if ((st = copy_back(ctx, L, 0, 1, out_xy, 64, out_inf, dp)) != ECG_OK) return fail(ctx, st);
    return finish(ctx);
  }
  // several devices: one partial point per device, gathered through the host (96 B each), summed on device 0
  std::vector<uint8_t> partial(nd * 96, 0);
  for (size_t i = 0; i < nd; i++) {
    DevState& d = ctx->devs[i];
    Lane& L = d.lane[0];
    partial[i * 96 + 63] = 1;  // identity (0:1:0) for empty shards
    if (shards[i].cnt == 0) continue;
    CU_TRY(ctx, cudaSetDevice(d.dev));
    uint32_t* res = nullptr;
    if ((st = lincomb_shard(ctx, d, curve, shards[i], k, P_xy, P_inf, &res)) != ECG_OK) return fail(ctx, st);
    if ((st = ensure(ctx, L, B_AUX, 256)) != ECG_OK) return fail(ctx, st);
    if ((st = export_point(ctx, L, curve, res, (uint8_t*)L.buf[B_AUX])) != ECG_OK) return fail(ctx, st);
    CU_TRY(ctx, cudaMemcpyAsync(&partial[i * 96], L.buf[B_AUX], 96, cudaMemcpyDeviceToHost, L.s()));
  }
  if ((st = finish(ctx)) != ECG_OK) return st;
  st = point_sum_host(ctx, curve, nd, partial.data(), out_xy, out_inf);
  return st == ECG_OK ? st : fail(ctx, st);
}

extern "C" ecg_status ecg_microbench(ecg_ctx* ctx, int which, int iters, double* ops_per_s, double* elapsed_ms) {
  if (!ctx || !ops_per_s || iters <= 0) return ECG_EINVAL;
  DevState& d = ctx->devs[0];
  Lane& L = d.lane[0];
  CU_TRY(ctx, cudaSetDevice(d.dev));
  ST_TRY(ensure(ctx, L, B_AUX, 256));
  uint32_t* out = (uint32_t*)L.buf[B_AUX];
  unsigned blocks = (unsigned)d.sm_count * 8, threads = 256;
  double per_thread_iter = 0;
  cudaEvent_t e0, e1;
  CU_TRY(ctx, cudaEventCreate(&e0));
  CU_TRY(ctx, cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {  // rep 0 = warm-up
    CU_TRY(ctx, cudaEventRecord(e0, L.s()));
    switch (which) {
      case 0: mb_imad_wide_kernel<<<blocks, threads, 0, L.s()>>>(out, iters, 12345u + rep); per_thread_iter = 32; break;
      case 1: mb_imad_kernel<<<blocks, threads, 0, L.s()>>>(out, iters, 12345u + rep); per_thread_iter = 64; break;
      case 2: mb_iadd_kernel<<<blocks, threads, 0, L.s()>>>(out, iters, 12345u + rep); per_thread_iter = 64; break;
      case 3: mb_fmul_kernel<FpK256><<<blocks, threads, 0, L.s()>>>(out, iters, 12345u + rep); per_thread_iter = 2; break;
      case 4: mb_fmul_kernel<FpP256><<<blocks, threads, 0, L.s()>>>(out, iters, 12345u + rep); per_thread_iter = 2; break;
      default:
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return ECG_EINVAL;
    }
    ctx->launches++;
    CU_TRY(ctx, cudaEventRecord(e1, L.s()));
    CU_TRY(ctx, cudaEventSynchronize(e1));
    CU_TRY(ctx, cudaGetLastError());
    float ms = 0;
    CU_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ops_per_s = (double)blocks * threads * (double)iters * per_thread_iter / (best * 1e-3);
  if (elapsed_ms) *elapsed_ms = best;
  return ECG_OK;
}
