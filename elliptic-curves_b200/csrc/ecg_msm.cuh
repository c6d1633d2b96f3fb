// ecg_msm.cuh — bucket-method (Pippenger) multi-scalar multiplication kernels: sum_i k_i * P_i for large n.
//
// Replaces LinearCombination::lincomb / lincomb_vartime for big slices (k256/src/arithmetic/mul.rs:66-175,
// primeorder/src/projective.rs:480-557).  The reference walks all terms through shared doublings (Straus, 2 KiB of
// tables per term); with 2^21 terms per GPU the bucket method needs ~16 mixed additions per term instead of a full
// scalar multiplication:
//   1. msm_prep_kernel      validate, GLV-split (secp256k1), cut every half-scalar into c-bit signed digits, count
//                           how many points fall in each (window, |digit|) bucket             [atomics on counters]
//   2. msm_scan_kernel      exclusive prefix sum of the counters                              [one block]
//   3. msm_scatter_kernel   write each (point, sign) reference into its bucket's slice        [counting sort]
//   4. msm_bucket_kernel    one thread per bucket: gather its points, mixed-add them           [the hot kernel]
//   5. msm_wreduce_kernel   per window sum_j j*B_j by chunked running sums (recursive on the chunk totals)
//   6. msm_final_kernel     undo the recursion's bias, result = sum_w 2^(c w) R_w
// Everything is order-independent group arithmetic, so the (affine, canonical) result is bit-identical to the
// reference's whatever order the atomics produce.
#pragma once
#include "ecg_curves.cuh"
#include "ecg_io.cuh"
#include "ecg_mul.cuh"

namespace ecg {

struct MsmGeom {
  int c;          // window bits
  int W;          // windows per sub-scalar
  int nbits;      // bits per sub-scalar magnitude (128 with GLV, 256 without)
  uint32_t nbw;   // bucket slots per window: digits 0 .. 2^(c+1)+1 (slot 0 unused)
};

// bits [pos, pos+width) of a little-endian limb array of nm limbs (9 for the 256-bit curves, 13 for P-384), width <= 17
ECG_D uint32_t msm_bits(const uint32_t* m, int pos, int width, int nm = 9) {
  int w = pos >> 5, b = pos & 31;
  uint64_t lo = w < nm ? m[w] : 0;
  uint64_t hi = (w + 1 < nm) ? m[w + 1] : 0;
  uint64_t v = (lo | (hi << 32)) >> b;
  return (uint32_t)(v & ((1u << width) - 1u));
}

// Signed c-bit digits of a magnitude m (9 limbs, < 2^nbits (+1 bit of slack)): windows 0..W-2 in
// [-2^(c-1)+1, 2^(c-1)], the top window unsigned (absorbs the carry and any slack bit).  out[w] = digit.
ECG_D void msm_recode(int32_t* out, const uint32_t* m, const MsmGeom& g, int nm = 9) {
  uint32_t carry = 0;
  const uint32_t half = 1u << (g.c - 1);
  for (int w = 0; w < g.W - 1; w++) {
    uint32_t v = msm_bits(m, g.c * w, g.c, nm) + carry;
    if (v > half) {
      out[w] = (int32_t)v - (int32_t)(1u << g.c);
      carry = 1;
    } else {
      out[w] = (int32_t)v;
      carry = 0;
    }
  }
  int top = g.c * (g.W - 1);
  int width = g.nbits + 1 - top;  // one slack bit
  if (width > 17) width = 17;
  out[g.W - 1] = (int32_t)(msm_bits(m, top, width, nm) + carry);
}

#if defined(__CUDACC__) || defined(ECG_HOST_SIM)  // kernels: CUDA, or the host simulation of tests/sim

// pts: sub-point j at pts[j*16 .. j*16+15] (x[8], y[8], internal form).  digits: sub-term j, window w at
// digits[w * nsub + j] (signed, the sub-scalar's own sign already folded in; 0 = nothing to add).
template <class C, bool GLV>
ECG_KERNEL(128)
    msm_prep_kernel(const uint8_t* __restrict__ kb, const uint8_t* __restrict__ pxy, const uint8_t* __restrict__ pinf,
                    size_t n, MsmGeom g, uint32_t* __restrict__ pts, int32_t* __restrict__ digits,
                    uint32_t* __restrict__ count, uint32_t* __restrict__ status, size_t base) {
  typedef typename C::F F;
  typedef typename F::FeT Fe;
  constexpr int NL = F::NL, FB = F::FB;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const size_t nsub = GLV ? 2 * n : n;
  uint32_t k[NL], m[NL + 1];
  typename F::AffT P;
  bool skip = false;
  {
    uint32_t err = 0;
    load_fe<F>(k, kb + FB * idx);
    if (!ltN<NL>(k, C::N())) err |= 1u;
    bool inf = pinf != nullptr && pinf[idx] != 0;
    Fe x, y;
    load_fe<F>(x.v, pxy + 2 * FB * idx);
    load_fe<F>(y.v, pxy + 2 * FB * idx + FB);
    F::from_canonical(P.x, x);
    F::from_canonical(P.y, y);
    if (!inf) {
      bool ok = ltN<NL>(x.v, C::P()) && ltN<NL>(y.v, C::P());
      if (ok) {
        Fe b;
        C::b_internal(b);
        ok = aff_on_curve<F, C::A_IS_MINUS3>(P, b);
      }
      if (!ok) err |= 2u;
    }
    if (err) {
      atomicOr(&status[0], err);
      size_t gi = base + idx;
      atomicMin(&status[1], (uint32_t)(gi > 0xFFFFFFFEull ? 0xFFFFFFFEull : gi));
    }
    skip = inf || err != 0;
  }
  int32_t dg[4 * NL + 1];  // c >= 8 -> at most 4 NL windows
  const int nh = GLV ? 2 : 1;
  GlvHalf h1, h2;
  if (GLV) {
    // magnitudes are recovered from the signed-odd form kept by glv_split_k256: |k_i| = 2h + 1 - even
    glv_split_k256(h1, h2, k);
  }
  for (int half = 0; half < nh; half++) {
    size_t j = GLV ? 2 * idx + half : idx;
    uint32_t neg = 0;
#pragma unroll
    for (int i = 0; i < NL + 1; i++) m[i] = 0;
    if (GLV) {
      const GlvHalf& h = half ? h2 : h1;
      // m = 2h + 1 - even
      uint32_t t[5];
      t[0] = (h.h[0] << 1) | 1u;
      t[1] = (h.h[1] << 1) | (h.h[0] >> 31);
      t[2] = (h.h[2] << 1) | (h.h[1] >> 31);
      t[3] = (h.h[3] << 1) | (h.h[2] >> 31);
      t[4] = h.h[3] >> 31;
      m[0] = sub_cc(t[0], h.even);
      m[1] = subc_cc(t[1], 0);
      m[2] = subc_cc(t[2], 0);
      m[3] = subc_cc(t[3], 0);
      m[4] = subc(t[4], 0);
      neg = h.neg;
    } else {
#pragma unroll
      for (int i = 0; i < NL; i++) m[i] = k[i];
    }
    // sub-point: P or (beta x, y)
    Fe px = P.x;
    if constexpr (GLV) {  // secp256k1 only
      if (half) {
        Fe beta;
        k256_beta(beta);
        F::mul(px, px, beta);
      }
    }
#pragma unroll
    for (int w = 0; w < NL; w++) {
      pts[j * (2 * NL) + w] = px.v[w];
      pts[j * (2 * NL) + NL + w] = P.y.v[w];
    }
    msm_recode(dg, m, g, NL + 1);
    for (int w = 0; w < g.W; w++) {
      int32_t d = skip ? 0 : dg[w];
      if (neg) d = -d;
      digits[(size_t)w * nsub + j] = d;
      if (d != 0) {
        uint32_t a = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
        atomicAdd(&count[(size_t)w * g.nbw + a], 1u);
      }
    }
  }
}

// Exclusive scan of count[0..m) into offset[0..m] in three small kernels (4096 counters per block):
//   msm_scan_partial_kernel  per-block totals + the largest bucket population (*maxcnt; the host falls back to the
//                            per-term kernel when one bucket would serialise the accumulation)
//   msm_scan_top_kernel      exclusive scan of the block totals (one block), offset[m] = grand total
//   msm_scan_final_kernel    per-block scan + block offset
#define MSM_SCAN_BLOCK 256
#define MSM_SCAN_PER_THREAD 16
#define MSM_SCAN_CHUNK (MSM_SCAN_BLOCK * MSM_SCAN_PER_THREAD)
template <int ECG_ONCE = 0>  // a template only so that several translation units may define it
ECG_KERNEL(MSM_SCAN_BLOCK)
    msm_scan_partial_kernel(const uint32_t* __restrict__ count, size_t m, uint32_t* __restrict__ blocksum, uint32_t* __restrict__ maxcnt) {
  __shared__ uint32_t ssum[MSM_SCAN_BLOCK], smax[MSM_SCAN_BLOCK];
  size_t lo = (size_t)blockIdx.x * MSM_SCAN_CHUNK + (size_t)threadIdx.x * MSM_SCAN_PER_THREAD;
  uint32_t s = 0, mx = 0;
  for (int i = 0; i < MSM_SCAN_PER_THREAD; i++)
    if (lo + i < m) {
      uint32_t c = count[lo + i];
      s += c;
      mx = c > mx ? c : mx;
    }
  ssum[threadIdx.x] = s;
  smax[threadIdx.x] = mx;
  __syncthreads();
  for (int st = MSM_SCAN_BLOCK / 2; st > 0; st >>= 1) {
    if ((int)threadIdx.x < st) {
      ssum[threadIdx.x] += ssum[threadIdx.x + st];
      smax[threadIdx.x] = smax[threadIdx.x] > smax[threadIdx.x + st] ? smax[threadIdx.x] : smax[threadIdx.x + st];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    blocksum[blockIdx.x] = ssum[0];
    atomicMax(maxcnt, smax[0]);
  }
}
template <int ECG_ONCE = 0>  // a template only so that several translation units may define it
ECG_KERNEL(1024)
    msm_scan_top_kernel(uint32_t* __restrict__ blocksum, size_t nblocks, uint32_t* __restrict__ offset, size_t m) {
  // nblocks <= a few thousand: serial per-thread chunks + one serial pass over 1024 partials
  __shared__ uint32_t part[1024];
  size_t per = (nblocks + 1023) / 1024;
  size_t lo = threadIdx.x * per, hi = lo + per < nblocks ? lo + per : nblocks;
  uint32_t s = 0;
  for (size_t i = lo; i < hi; i++) s += blocksum[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int i = 0; i < 1024; i++) {
      uint32_t t = part[i];
      part[i] = run;
      run += t;
    }
    offset[m] = run;
  }
  __syncthreads();
  uint32_t run = part[threadIdx.x];
  for (size_t i = lo; i < hi; i++) {
    uint32_t t = blocksum[i];
    blocksum[i] = run;
    run += t;
  }
}
template <int ECG_ONCE = 0>  // a template only so that several translation units may define it
ECG_KERNEL(MSM_SCAN_BLOCK)
    msm_scan_final_kernel(const uint32_t* __restrict__ count, size_t m, const uint32_t* __restrict__ blockoff, uint32_t* __restrict__ offset) {
  __shared__ uint32_t ssum[MSM_SCAN_BLOCK];
  size_t lo = (size_t)blockIdx.x * MSM_SCAN_CHUNK + (size_t)threadIdx.x * MSM_SCAN_PER_THREAD;
  uint32_t c[MSM_SCAN_PER_THREAD];
  uint32_t s = 0;
  for (int i = 0; i < MSM_SCAN_PER_THREAD; i++) {
    c[i] = lo + i < m ? count[lo + i] : 0;
    s += c[i];
  }
  ssum[threadIdx.x] = s;
  __syncthreads();
  // Hillis-Steele inclusive scan over the 256 thread totals
  for (int st = 1; st < MSM_SCAN_BLOCK; st <<= 1) {
    uint32_t v = (int)threadIdx.x >= st ? ssum[threadIdx.x - st] : 0;
    __syncthreads();
    ssum[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = blockoff[blockIdx.x] + ssum[threadIdx.x] - s;
  for (int i = 0; i < MSM_SCAN_PER_THREAD; i++)
    if (lo + i < m) {
      offset[lo + i] = run;
      run += c[i];
    }
}

template <int ECG_ONCE = 0>  // a template only so that several translation units may define it
ECG_KERNEL(256)
    msm_scatter_kernel(const int32_t* __restrict__ digits, size_t nsub, MsmGeom g, const uint32_t* __restrict__ offset,
                       uint32_t* __restrict__ cursor, uint32_t* __restrict__ list) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nsub) return;
  for (int w = 0; w < g.W; w++) {
    int32_t d = digits[(size_t)w * nsub + j];
    if (d == 0) continue;
    uint32_t a = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
    size_t key = (size_t)w * g.nbw + a;
    uint32_t pos = atomicAdd(&cursor[key], 1u);
    list[offset[key] + pos] = ((uint32_t)j << 1) | (d < 0 ? 1u : 0u);
  }
}

template <int NL>
ECG_DEV void msm_load_point(AffN<NL>& e, const uint32_t* __restrict__ pts, uint32_t j) {
  load_aff_entry<NL>(e, pts, (size_t)j);  // ecg_io.cuh: 128-bit (or, for odd NL, 64-bit) read-only loads
}

// Bucket ids ordered by decreasing population: counting sort over MSM_ORDER_CLASSES size classes (sizes beyond the
// last class share it).  hist: MSM_ORDER_CLASSES + 1 counters, cleared by the host; class 0 = largest.
#define MSM_ORDER_CLASSES 1024
ECG_DEV uint32_t msm_size_class(uint32_t sz) { return (MSM_ORDER_CLASSES - 1) - (sz < MSM_ORDER_CLASSES - 1 ? sz : MSM_ORDER_CLASSES - 1); }
template <int ECG_ONCE = 0>  // a template only so that several translation units may define it
ECG_KERNEL(256)
    msm_order_hist_kernel(const uint32_t* __restrict__ offset, size_t nb, uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[MSM_ORDER_CLASSES];
  for (unsigned i = threadIdx.x; i < MSM_ORDER_CLASSES; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += stride) atomicAdd(&sh[msm_size_class(offset[b + 1] - offset[b])], 1u);
  __syncthreads();
  for (unsigned i = threadIdx.x; i < MSM_ORDER_CLASSES; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
// exclusive scan of the class counts (one block of MSM_ORDER_CLASSES threads), in place: hist[c] becomes the first
// position of class c
template <int ECG_ONCE = 0>  // a template only so that several translation units may define it
ECG_KERNEL(MSM_ORDER_CLASSES)
    msm_order_scan_kernel(uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[MSM_ORDER_CLASSES];
  unsigned t = threadIdx.x;
  uint32_t v = hist[t];
  sh[t] = v;
  __syncthreads();
  for (int st = 1; st < MSM_ORDER_CLASSES; st <<= 1) {
    uint32_t a = (int)t >= st ? sh[t - st] : 0;
    __syncthreads();
    sh[t] += a;
    __syncthreads();
  }
  hist[t] = sh[t] - v;
}
// every block takes a contiguous run of buckets, counts its classes in shared memory, reserves one range per class with
// a single global atomic, and writes its bucket ids there
template <int ECG_ONCE = 0>  // a template only so that several translation units may define it
ECG_KERNEL(256)
    msm_order_scatter_kernel(const uint32_t* __restrict__ offset, size_t nb, uint32_t* __restrict__ cursor, uint32_t* __restrict__ order) {
  __shared__ uint32_t cnt[MSM_ORDER_CLASSES], basep[MSM_ORDER_CLASSES];
  const size_t per_block = (nb + gridDim.x - 1) / gridDim.x;
  const size_t lo = (size_t)blockIdx.x * per_block, hi = lo + per_block < nb ? lo + per_block : nb;
  for (unsigned i = threadIdx.x; i < MSM_ORDER_CLASSES; i += blockDim.x) cnt[i] = 0;
  __syncthreads();
  for (size_t b = lo + threadIdx.x; b < hi; b += blockDim.x) atomicAdd(&cnt[msm_size_class(offset[b + 1] - offset[b])], 1u);
  __syncthreads();
  for (unsigned i = threadIdx.x; i < MSM_ORDER_CLASSES; i += blockDim.x) {
    uint32_t c = cnt[i];
    basep[i] = c ? atomicAdd(&cursor[i], c) : 0u;
    cnt[i] = 0;
  }
  __syncthreads();
  for (size_t b = lo + threadIdx.x; b < hi; b += blockDim.x) {
    uint32_t cls = msm_size_class(offset[b + 1] - offset[b]);
    uint32_t pos = basep[cls] + atomicAdd(&cnt[cls], 1u);
    order[pos] = (uint32_t)b;
  }
}

// Skew guard, decided on the device (no host round trip in the middle of a call): one bucket thread adds its points
// serially, so an input that piles thousands of terms into one bucket (e.g. many identical terms) would serialise the
// whole accumulation.  When the largest bucket population (written by msm_scan_partial_kernel) exceeds both limits the
// bucket kernels do nothing except raise MSM_SKEW_FLAG in status[0]; the host sees the flag when the call finishes and
// repeats the call on the per-term path, whose cost does not depend on the data.
#define MSM_SKEW_FLAG 4u
struct MsmSkew {
  const uint32_t* maxcnt;
  uint32_t limit_abs, limit_rel;
  uint32_t* status;
};
ECG_DEV bool msm_skewed(const MsmSkew& sk, bool reporter) {
  uint32_t m = *sk.maxcnt;
  if (m > sk.limit_abs && m > sk.limit_rel) {
    if (reporter) atomicOr(&sk.status[0], MSM_SKEW_FLAG);
    return true;
  }
  return false;
}

// One thread per bucket: B = sum of its (signed) points.  bkt: SoA Jacobian over nb = W*nbw buckets.
template <class C>
ECG_KERNEL(128, 4)
    msm_bucket_kernel(const uint32_t* __restrict__ pts, const uint32_t* __restrict__ list, const uint32_t* __restrict__ offset,
                      size_t nb, uint32_t* __restrict__ bkt, MsmSkew sk, const uint32_t* __restrict__ order) {
  typedef typename C::F F;
  typedef typename F::JacT Jac;
  typedef typename F::AffT Aff;
  constexpr int NL = F::NL;
  (void)NL;
  size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (msm_skewed(sk, b == 0)) return;
  if (b >= nb) return;
  // order (msm_order_* kernels): bucket ids by decreasing population, so that the 32 lanes of a warp get buckets of
  // (nearly) equal size — a warp otherwise waits for the largest of 32 Poisson-sized buckets (+19 % at mean 128) — and
  // the longest buckets start first
  if (order != nullptr) b = order[b];
  uint32_t lo = offset[b], hi = offset[b + 1];
  Jac acc;
  F::set_zero(acc.X);
  F::set_one(acc.Y);
  F::set_zero(acc.Z);
  if (lo < hi) {
    uint32_t ent = list[lo];
    Aff e, nx;
    msm_load_point<NL>(e, pts, ent >> 1);
    fe_cneg<F>(e.y, ent & 1u);
    acc.X = e.x;
    acc.Y = e.y;
    F::set_one(acc.Z);
    // software pipeline: the gather of point i+1 (a random 64-byte read, usually an L2 miss) is issued before the
    // mixed addition of point i
    uint32_t nent = lo + 1 < hi ? list[lo + 1] : 0;
    if (lo + 1 < hi) msm_load_point<NL>(nx, pts, nent >> 1);
#pragma unroll 1
    for (uint32_t i = lo + 1; i < hi; i++) {
      e = nx;
      ent = nent;
      if (i + 1 < hi) {
        nent = list[i + 1];
        msm_load_point<NL>(nx, pts, nent >> 1);
      }
      fe_cneg<F>(e.y, ent & 1u);
      jac_madd<F, C::A_IS_MINUS3>(acc, acc, e);
    }
  }
#pragma unroll
  for (int w = 0; w < NL; w++) {
    bkt[(size_t)w * nb + b] = acc.X.v[w];
    bkt[(size_t)(NL + w) * nb + b] = acc.Y.v[w];
    bkt[(size_t)(2 * NL + w) * nb + b] = acc.Z.v[w];
  }
}

// The sum of one bucket (list entries offset[b] .. offset[b+1]) -> bkt (shared by both bucket kernels).
template <class C>
ECG_DEV void msm_bucket_sum(const uint32_t* __restrict__ pts, const uint32_t* __restrict__ list, uint32_t lo, uint32_t hi, size_t b,
                            size_t nb, uint32_t* __restrict__ bkt) {
  typedef typename C::F F;
  typedef typename F::JacT Jac;
  typedef typename F::AffT Aff;
  constexpr int NL = F::NL;
  (void)NL;
  Jac acc;
  F::set_zero(acc.X);
  F::set_one(acc.Y);
  F::set_zero(acc.Z);
  if (lo < hi) {
    uint32_t ent = list[lo];
    Aff e, nx;
    msm_load_point<NL>(e, pts, ent >> 1);
    fe_cneg<F>(e.y, ent & 1u);
    acc.X = e.x;
    acc.Y = e.y;
    F::set_one(acc.Z);
    uint32_t nent = lo + 1 < hi ? list[lo + 1] : 0;
    if (lo + 1 < hi) msm_load_point<NL>(nx, pts, nent >> 1);
#pragma unroll 1
    for (uint32_t i = lo + 1; i < hi; i++) {
      e = nx;
      ent = nent;
      if (i + 1 < hi) {
        nent = list[i + 1];
        msm_load_point<NL>(nx, pts, nent >> 1);
      }
      fe_cneg<F>(e.y, ent & 1u);
      jac_madd<F, C::A_IS_MINUS3>(acc, acc, e);
    }
  }
#pragma unroll
  for (int w = 0; w < NL; w++) {
    bkt[(size_t)w * nb + b] = acc.X.v[w];
    bkt[(size_t)(NL + w) * nb + b] = acc.Y.v[w];
    bkt[(size_t)(2 * NL + w) * nb + b] = acc.Z.v[w];
  }
}

// Variant of msm_bucket_kernel that balances the lanes of a warp (ECG_MSM_BUCKETS_PER_THREAD=K, K > 1; off by
// default until measured).  Bucket sizes are Poisson-like (mean 32 at 2^21 terms), and a warp waits for its largest
// bucket: max over 32 lanes ~ mean + 2 sigma = +35 %.  Here a block of 128 threads owns 128*K consecutive buckets,
// orders them by size with a counting sort in shared memory (256 size classes, largest first) and processes them in K
// rounds, thread t taking the (r*128 + t)-th largest: the 32 lanes of a warp get neighbours in that order, i.e.
// buckets of (almost) equal size.  Results land in the same bkt slots, so everything downstream is unchanged.
#define MSM_BS_BLOCK 128
template <class C, int K>
ECG_KERNEL(MSM_BS_BLOCK, 4)
    msm_bucket_sorted_kernel(const uint32_t* __restrict__ pts, const uint32_t* __restrict__ list, const uint32_t* __restrict__ offset,
                             size_t nb, uint32_t* __restrict__ bkt, MsmSkew sk) {
  __shared__ uint32_t hist[256], start[256], part[MSM_BS_BLOCK];
  __shared__ uint16_t order[MSM_BS_BLOCK * K];
  const unsigned t = threadIdx.x;
  const size_t base = (size_t)blockIdx.x * (MSM_BS_BLOCK * K);
  if (msm_skewed(sk, blockIdx.x == 0 && t == 0)) return;  // block-uniform: every thread takes the same exit
  hist[t] = 0;
  hist[t + MSM_BS_BLOCK] = 0;
  __syncthreads();
  uint32_t cls[K];
#pragma unroll
  for (int i = 0; i < K; i++) {
    size_t b = base + (size_t)i * MSM_BS_BLOCK + t;
    uint32_t sz = b < nb ? offset[b + 1] - offset[b] : 0u;
    cls[i] = 255u - (sz < 255u ? sz : 255u);  // class 0 = largest
    atomicAdd(&hist[cls[i]], 1u);
  }
  __syncthreads();
  // exclusive scan of the 256 class counts: per-thread pairs, one serial pass over 128 partials
  uint32_t h0 = hist[2 * t], h1 = hist[2 * t + 1];
  part[t] = h0 + h1;
  __syncthreads();
  if (t == 0) {
    uint32_t run = 0;
    for (int i = 0; i < MSM_BS_BLOCK; i++) {
      uint32_t v = part[i];
      part[i] = run;
      run += v;
    }
  }
  __syncthreads();
  start[2 * t] = part[t];
  start[2 * t + 1] = part[t] + h0;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < K; i++) {
    uint32_t pos = atomicAdd(&start[cls[i]], 1u);
    order[pos] = (uint16_t)(i * MSM_BS_BLOCK + t);
  }
  __syncthreads();
#pragma unroll 1
  for (int r = 0; r < K; r++) {
    size_t b = base + order[r * MSM_BS_BLOCK + t];
    if (b < nb) msm_bucket_sum<C>(pts, list, offset[b], offset[b + 1], b, nb, bkt);
  }
}

template <int NL>
ECG_DEV void msm_jload(JacN<NL>& p, const uint32_t* __restrict__ a, size_t n, size_t i) {
#pragma unroll
  for (int w = 0; w < NL; w++) {
    p.X.v[w] = a[(size_t)w * n + i];
    p.Y.v[w] = a[(size_t)(NL + w) * n + i];
    p.Z.v[w] = a[(size_t)(2 * NL + w) * n + i];
  }
}
template <int NL>
ECG_DEV void msm_jstore(uint32_t* __restrict__ a, size_t n, size_t i, const JacN<NL>& p) {
#pragma unroll
  for (int w = 0; w < NL; w++) {
    a[(size_t)w * n + i] = p.X.v[w];
    a[(size_t)(NL + w) * n + i] = p.Y.v[w];
    a[(size_t)(2 * NL + w) * n + i] = p.Z.v[w];
  }
}

// Weighted reduction of every window's buckets, R_w = sum_{j>=1} j * B_{w,j}, by recursive chunking.
// Level l reads W rows of `len` points A_j (level 0: the buckets, weight j+1 for element j; level l>0: the chunk
// totals S of level l-1).  Thread (w, ch) walks its chunk of MSM_CH elements from the top with the running-sum trick
//     S  = sum A_j                      (chunk total   -> input of level l+1)
//     T  = sum (j - ch*CH + 1) A_j      (chunk-local weighted sum)
// and carries the plain sum of everything the lower levels produced:
//     X_l[ch] = sum_{ch' in chunk} X_{l-1}[ch']  +  (CH_0 ... CH_{l-1}) * T          (X_{-1} = nothing)
// At the last level (one chunk per row)  X_L = R_w + Btot * (CH_0 + CH_0 CH_1 + ... + CH_0 ... CH_{L-1}),  Btot = S_L, which
// msm_final_kernel undoes before the Horner combination of the windows.
// Chunk sizes: 16 at level 0 (one thread per chunk still gives tens of thousands of threads: throughput-bound), 4 at
// the levels above, which hold only a few thousand elements and are bound by the LENGTH of each thread's dependent
// chain (3 additions per element): 16+4+4+... keeps that chain at 12 additions per level instead of 48.
#define MSM_CH0 16
#define MSM_CH0_LOG2 4
#define MSM_CHU 4
#define MSM_CHU_LOG2 2
ECG_D int msm_ch(int level) { return level == 0 ? MSM_CH0 : MSM_CHU; }
// log2 of the product of the chunk sizes of the levels below `level` (the weight of one level-`level` element)
ECG_D int msm_scale_log2(int level) { return level == 0 ? 0 : MSM_CH0_LOG2 + MSM_CHU_LOG2 * (level - 1); }
template <class C>
ECG_KERNEL(128)
    msm_wreduce_kernel(const uint32_t* __restrict__ in, size_t n_in, size_t stride_in, size_t off, size_t len, size_t len_low, int W,
                       size_t nch, const uint32_t* __restrict__ Xprev, int level, uint32_t* __restrict__ outS, uint32_t* __restrict__ outX) {
  typedef typename C::F F;
  typedef typename F::JacT Jac;
  typedef typename F::AffT Aff;
  constexpr int NL = F::NL;
  (void)NL;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)W * nch) return;
  size_t w = t / nch, ch = t % nch;
  // only the top window (unsigned, absorbs the recoding carry and the slack bit) uses all `len` slots of its row; the
  // signed windows below it stop at len_low = 2^(c-1) at level 0: three quarters of the rows' slots are never
  // populated, and the chunks that cover only such slots contribute the identity without being read
  const size_t len_w = (w + 1 == (size_t)W) ? len : len_low;
  const size_t CHL = (size_t)msm_ch(level);
  size_t lo = ch * CHL, hi = lo + CHL < len_w ? lo + CHL : len_w;
  if (hi < lo) hi = lo;
  Jac S, T, X, p;
  F::set_zero(S.X);
  F::set_one(S.Y);
  F::set_zero(S.Z);
  T = S;
  X = S;
  for (size_t j = hi; j > lo; j--) {
    msm_jload(p, in, n_in, w * stride_in + off + (j - 1));
    jac_add<F, C::A_IS_MINUS3>(S, S, p);
    jac_add<F, C::A_IS_MINUS3>(T, T, S);
    if (Xprev != nullptr) {
      msm_jload(p, Xprev, n_in, w * stride_in + (j - 1));
      jac_add<F, C::A_IS_MINUS3>(X, X, p);
    }
  }
  for (int i = 0; i < msm_scale_log2(level); i++) jac_dbl<F, C::A_IS_MINUS3>(T, T);
  jac_add<F, C::A_IS_MINUS3>(X, X, T);
  size_t n_out = (size_t)W * nch;
  msm_jstore(outS, n_out, t, S);
  msm_jstore(outX, n_out, t, X);
}

// R_w = X_L[w] - k * Btot[w],  k = CH0 (1 + CHU (1 + CHU (...))) with L factors (one per level above 0);
// then out = sum_w 2^(c w) R_w  (Horner, thread 0).
#define MSM_FINAL_THREADS 64 /* one thread per window: W <= 48 (384-bit scalars at c = 8) */
template <class C>
ECG_KERNEL(MSM_FINAL_THREADS)
    msm_final_kernel(const uint32_t* __restrict__ XL, const uint32_t* __restrict__ SL, int W, int c, int levels,
                     uint32_t* __restrict__ R, uint32_t* __restrict__ out) {
  typedef typename C::F F;
  typedef typename F::JacT Jac;
  typedef typename F::AffT Aff;
  constexpr int NL = F::NL;
  (void)NL;
  int w = threadIdx.x;
  if (w < W) {
    Jac x, s, ks, t;
    msm_jload(x, XL, W, w);
    msm_jload(s, SL, W, w);
    // ks = k * s by Horner: innermost factor first (levels above 0 use CHU), CH0 last
    ks = s;
    for (int l = 1; l < levels; l++) {
      for (int i = 0; i < MSM_CHU_LOG2; i++) jac_dbl<F, C::A_IS_MINUS3>(ks, ks);
      jac_add<F, C::A_IS_MINUS3>(ks, ks, s);
    }
    for (int i = 0; i < MSM_CH0_LOG2; i++) jac_dbl<F, C::A_IS_MINUS3>(ks, ks);
    if (levels == 0) {
      F::set_zero(ks.X);
      F::set_one(ks.Y);
      F::set_zero(ks.Z);
    }
    F::neg(ks.Y, ks.Y);
    jac_add<F, C::A_IS_MINUS3>(t, x, ks);
    msm_jstore(R, W, w, t);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Jac acc, p;
    msm_jload(acc, R, W, W - 1);
    for (int ww = W - 2; ww >= 0; ww--) {
      for (int i = 0; i < c; i++) jac_dbl<F, C::A_IS_MINUS3>(acc, acc);
      msm_jload(p, R, W, ww);
      jac_add<F, C::A_IS_MINUS3>(acc, acc, p);
    }
    msm_jstore(out, 1, 0, acc);
  }
}

#endif  // __CUDACC__ || ECG_HOST_SIM

}  // namespace ecg
