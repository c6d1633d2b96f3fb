// ecgpu.cu — kernels and C ABI of libecgpu.so (see include/ecgpu.h for the contract).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
// There is no CPU fallback in this file: every compute entry launches sm_100a kernels or fails.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/ecgpu.h"
#include "ecg_curves.cuh"
#include "ecg_io.cuh"
#include "ecg_mul.cuh"

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
// secp256k1 variable-base: one pair per thread, window table in shared memory (512 B / thread).
template <int BLOCK, int MINBLK>
__global__ void __launch_bounds__(BLOCK, MINBLK)
    k256_varbase_kernel(const uint8_t* __restrict__ kb, const uint8_t* __restrict__ pxy,
                        const uint8_t* __restrict__ pinf, size_t n, uint32_t* __restrict__ jac,
                        uint32_t* __restrict__ status) {
  extern __shared__ uint32_t smem[];
  size_t idx = (size_t)blockIdx.x * BLOCK + threadIdx.x;
  if (idx >= n) return;
  uint32_t k[8];
  Aff P;
  bool inf;
  uint32_t err = load_pair<CurveK256>(k, P, inf, kb, pxy, pinf, idx);
  if (err) report_error(status, err, idx);
  TabRef tab{smem + threadIdx.x, (uint32_t)BLOCK};
  Jac r;
  k256_mul_thread(r, k, P, tab);
  if (inf || err) FpK256::set_zero(r.Z);
  soa_store<8>(jac, n, idx, r.X.v, 0);
  soa_store<8>(jac, n, idx, r.Y.v, 8);
  soa_store<8>(jac, n, idx, r.Z.v, 16);
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
                      uint32_t* __restrict__ status) {
  typedef typename C::F F;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
#pragma unroll 1
  for (int c = 0; c < 3; c++) {
    Fe v, w;
    load_be32(v.v, xyz + 96 * idx + 32 * c);
    if (!lt8(v.v, C::P())) report_error(status, ERRF_POINT, idx);
    F::from_canonical(w, v);
    soa_store<8>(jac, n, idx, w.v, 8 * c);
  }
}

// ------------------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(256)
    field_op_kernel(int op, size_t n, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                    uint8_t* __restrict__ out, uint32_t* __restrict__ status) {
  typedef typename C::F F;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  Fe x, y, r;
  load_be32(x.v, a + 32 * idx);
  if (!lt8(x.v, C::P())) report_error(status, ERRF_POINT, idx);
  F::from_canonical(x, x);
  bool binary = (op == ECG_FOP_ADD || op == ECG_FOP_SUB || op == ECG_FOP_MUL);
  if (binary) {
    load_be32(y.v, b + 32 * idx);
    if (!lt8(y.v, C::P())) report_error(status, ERRF_POINT, idx);
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
struct DevState {
  int dev = 0;
  cudaStream_t stream = nullptr;      // owned
  cudaStream_t user_stream = nullptr; // optional override
  bool use_user_stream = false;
  // grow-only device buffers
  void* buf[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t cap[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t* status = nullptr;  // 2 words
  uint32_t* h_status = nullptr;  // pinned
  int sm_count = 148;
  cudaStream_t s() const { return use_user_stream ? user_stream : stream; }
};
enum { B_K = 0, B_P = 1, B_INF = 2, B_JAC = 3, B_SCR = 4, B_OUT = 5, B_OINF = 6, B_AUX = 7 };

struct ecg_ctx {
  std::vector<DevState> devs;
  unsigned flags = 0;
  std::string err;
  size_t err_index = (size_t)-1;
  uint64_t launches = 0;
};

#define CU_TRY(ctx, call)                                                                   \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      char m_[512];                                                                         \
      snprintf(m_, sizeof m_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      (ctx)->err = m_;                                                                      \
      return e_ == cudaErrorMemoryAllocation ? ECG_ENOMEM : ECG_ECUDA;                      \
    }                                                                                       \
  } while (0)

static ecg_status ensure(ecg_ctx* ctx, DevState& d, int which, size_t bytes) {
  if (bytes <= d.cap[which]) return ECG_OK;
  if (d.buf[which]) CU_TRY(ctx, cudaFree(d.buf[which]));
  d.buf[which] = nullptr;
  d.cap[which] = 0;
  size_t want = bytes + bytes / 8 + 256;
  CU_TRY(ctx, cudaMalloc(&d.buf[which], want));
  d.cap[which] = want;
  return ECG_OK;
}

extern "C" const char* ecg_version(void) { return "ecgpu 0.1 (sm_100a)"; }

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
      delete ctx;
      return ECG_EINVAL;
    }
    if (cudaSetDevice(d.dev) != cudaSuccess || cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc((void**)&d.status, 8) != cudaSuccess || cudaMallocHost((void**)&d.h_status, 8) != cudaSuccess) {
      delete ctx;
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
    if (d.stream) {
      cudaStreamSynchronize(d.stream);
      cudaStreamDestroy(d.stream);
    }
    for (int i = 0; i < 8; i++)
      if (d.buf[i]) cudaFree(d.buf[i]);
    if (d.status) cudaFree(d.status);
    if (d.h_status) cudaFreeHost(d.h_status);
  }
  delete ctx;
}

extern "C" const char* ecg_last_error(const ecg_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
extern "C" size_t ecg_last_error_index(const ecg_ctx* ctx) { return ctx ? ctx->err_index : (size_t)-1; }
extern "C" uint64_t ecg_kernel_launches(const ecg_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" ecg_status ecg_ctx_set_stream(ecg_ctx* ctx, void* cuda_stream) {
  if (!ctx) return ECG_EINVAL;
  DevState& d = ctx->devs[0];
  d.user_stream = (cudaStream_t)cuda_stream;
  d.use_user_stream = cuda_stream != nullptr;
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

struct DevPtrs {
  const uint8_t *k = nullptr, *p = nullptr, *inf = nullptr, *a = nullptr;
  uint8_t *out = nullptr, *oinf = nullptr;
};

// Stage inputs (host mode: async H2D into ctx buffers; device mode: use the caller's pointers).
static ecg_status stage_in(ecg_ctx* ctx, DevState& d, const Shard& sh, const uint8_t* k, size_t kstride,
                           const uint8_t* p, size_t pstride, const uint8_t* inf, DevPtrs& dp) {
  bool devptr = ctx->flags & ECG_FLAG_DEVICE_PTRS;
  if (devptr) {
    dp.k = k ? k + sh.off * kstride : nullptr;
    dp.p = p ? p + sh.off * pstride : nullptr;
    dp.inf = inf ? inf + sh.off : nullptr;
    return ECG_OK;
  }
  ecg_status st;
  if (k) {
    if ((st = ensure(ctx, d, B_K, sh.cnt * kstride)) != ECG_OK) return st;
    CU_TRY(ctx, cudaMemcpyAsync(d.buf[B_K], k + sh.off * kstride, sh.cnt * kstride, cudaMemcpyHostToDevice, d.s()));
    dp.k = (const uint8_t*)d.buf[B_K];
  }
  if (p) {
    if ((st = ensure(ctx, d, B_P, sh.cnt * pstride)) != ECG_OK) return st;
    CU_TRY(ctx, cudaMemcpyAsync(d.buf[B_P], p + sh.off * pstride, sh.cnt * pstride, cudaMemcpyHostToDevice, d.s()));
    dp.p = (const uint8_t*)d.buf[B_P];
  }
  if (inf) {
    if ((st = ensure(ctx, d, B_INF, sh.cnt)) != ECG_OK) return st;
    CU_TRY(ctx, cudaMemcpyAsync(d.buf[B_INF], inf + sh.off, sh.cnt, cudaMemcpyHostToDevice, d.s()));
    dp.inf = (const uint8_t*)d.buf[B_INF];
  }
  return ECG_OK;
}
static ecg_status stage_out(ecg_ctx* ctx, DevState& d, const Shard& sh, uint8_t* out, size_t ostride,
                            uint8_t* oinf, DevPtrs& dp) {
  bool devptr = ctx->flags & ECG_FLAG_DEVICE_PTRS;
  if (devptr) {
    dp.out = out + sh.off * ostride;
    dp.oinf = oinf ? oinf + sh.off : nullptr;
    if (!dp.oinf) {
      ecg_status st = ensure(ctx, d, B_OINF, sh.cnt);
      if (st != ECG_OK) return st;
      dp.oinf = (uint8_t*)d.buf[B_OINF];
    }
    return ECG_OK;
  }
  ecg_status st;
  if ((st = ensure(ctx, d, B_OUT, sh.cnt * ostride)) != ECG_OK) return st;
  if ((st = ensure(ctx, d, B_OINF, sh.cnt)) != ECG_OK) return st;
  dp.out = (uint8_t*)d.buf[B_OUT];
  dp.oinf = (uint8_t*)d.buf[B_OINF];
  return ECG_OK;
}
static ecg_status copy_back(ecg_ctx* ctx, DevState& d, const Shard& sh, uint8_t* out, size_t ostride,
                            uint8_t* oinf, const DevPtrs& dp) {
  if (ctx->flags & ECG_FLAG_DEVICE_PTRS) return ECG_OK;
  CU_TRY(ctx, cudaMemcpyAsync(out + sh.off * ostride, dp.out, sh.cnt * ostride, cudaMemcpyDeviceToHost, d.s()));
  if (oinf) CU_TRY(ctx, cudaMemcpyAsync(oinf + sh.off, dp.oinf, sh.cnt, cudaMemcpyDeviceToHost, d.s()));
  return ECG_OK;
}
static ecg_status reset_status(ecg_ctx* ctx, DevState& d) {
  static const uint32_t init[2] = {0u, 0xFFFFFFFFu};
  CU_TRY(ctx, cudaMemcpyAsync(d.status, init, 8, cudaMemcpyHostToDevice, d.s()));
  return ECG_OK;
}
// Wait for every device, fold the validation status into a return code.
static ecg_status finish(ecg_ctx* ctx, const std::vector<Shard>& shards) {
  ecg_status rc = ECG_OK;
  size_t first = (size_t)-1;
  for (size_t i = 0; i < ctx->devs.size(); i++) {
    DevState& d = ctx->devs[i];
    if (shards[i].cnt == 0) continue;
    CU_TRY(ctx, cudaSetDevice(d.dev));
    CU_TRY(ctx, cudaMemcpyAsync(d.h_status, d.status, 8, cudaMemcpyDeviceToHost, d.s()));
    CU_TRY(ctx, cudaStreamSynchronize(d.s()));
    CU_TRY(ctx, cudaGetLastError());
    if (d.h_status[0]) {
      size_t idx = shards[i].off + d.h_status[1];
      if (idx < first) {
        first = idx;
        rc = (d.h_status[0] & ERRF_SCALAR) && !(d.h_status[0] & ERRF_POINT) ? ECG_ESCALAR_RANGE
             : (d.h_status[0] & ERRF_POINT) && !(d.h_status[0] & ERRF_SCALAR) ? ECG_ENOT_ON_CURVE
                                                                             : ECG_ESCALAR_RANGE;
        if ((d.h_status[0] & (ERRF_POINT | ERRF_SCALAR)) == (ERRF_POINT | ERRF_SCALAR)) rc = ECG_ENOT_ON_CURVE;
      }
    }
  }
  if (rc != ECG_OK) {
    ctx->err_index = first;
    ctx->err = rc == ECG_ESCALAR_RANGE ? "scalar out of range (>= n)" : "point not on curve / coordinate >= p";
  }
  return rc;
}

static inline unsigned grid_for(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

template <class F>
static ecg_status launch_normalize(ecg_ctx* ctx, DevState& d, size_t n, const uint32_t* jac, uint8_t* out, uint8_t* oinf) {
  ecg_status st = ensure(ctx, d, B_SCR, n * 32);
  if (st != ECG_OK) return st;
  // ~32 elements per thread amortise the per-thread inversion; never fewer threads than one per SM-warp slot
  size_t want_threads = std::max<size_t>((n + 31) / 32, std::min<size_t>(n, (size_t)d.sm_count * 256));
  unsigned blocks = grid_for(want_threads, 256);
  normalize_kernel<F><<<blocks, 256, 0, d.s()>>>(jac, n, (uint32_t*)d.buf[B_SCR], out, oinf);
  ctx->launches++;
  CU_TRY(ctx, cudaGetLastError());
  return ECG_OK;
}

static const int VB_BLOCK = 128, VB_MINBLK = 3;

static ecg_status mul_batch_dev(ecg_ctx* ctx, DevState& d, ecg_curve curve, size_t n, const DevPtrs& dp) {
  ecg_status st = ensure(ctx, d, B_JAC, n * 96);
  if (st != ECG_OK) return st;
  uint32_t* jac = (uint32_t*)d.buf[B_JAC];
  if (curve == ECG_SECP256K1) {
    size_t smem = (size_t)VB_BLOCK * 8 * 16 * 4;
    static bool attr_set[64] = {false};
    if (!attr_set[d.dev & 63]) {
      CU_TRY(ctx, cudaFuncSetAttribute(k256_varbase_kernel<VB_BLOCK, VB_MINBLK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set[d.dev & 63] = true;
    }
    k256_varbase_kernel<VB_BLOCK, VB_MINBLK><<<grid_for(n, VB_BLOCK), VB_BLOCK, smem, d.s()>>>(dp.k, dp.p, dp.inf, n, jac, d.status);
    ctx->launches++;
    CU_TRY(ctx, cudaGetLastError());
    return launch_normalize<FpK256>(ctx, d, n, jac, dp.out, dp.oinf);
  }
  ctx->err = "curve not implemented yet";
  return ECG_EINVAL;
}

extern "C" ecg_status ecg_mul_batch(ecg_ctx* ctx, ecg_curve curve, size_t n, const uint8_t* k, const uint8_t* P_xy,
                                    const uint8_t* P_inf, uint8_t* out_xy, uint8_t* out_inf) {
  if (!ctx) return ECG_EINVAL;
  if (n == 0) return ECG_OK;
  if (!k || !P_xy || !out_xy || (curve != ECG_SECP256K1 && curve != ECG_NISTP256)) {
    ctx->err = "ecg_mul_batch: null pointer or unknown curve";
    return ECG_EINVAL;
  }
  std::vector<Shard> shards = make_shards(n, ctx->devs.size());
  std::vector<DevPtrs> dps(ctx->devs.size());
  for (size_t i = 0; i < ctx->devs.size(); i++) {
    DevState& d = ctx->devs[i];
    const Shard& sh = shards[i];
    if (sh.cnt == 0) continue;
    CU_TRY(ctx, cudaSetDevice(d.dev));
    ecg_status st;
    if ((st = reset_status(ctx, d)) != ECG_OK) return st;
    if ((st = stage_in(ctx, d, sh, k, 32, P_xy, 64, P_inf, dps[i])) != ECG_OK) return st;
    if ((st = stage_out(ctx, d, sh, out_xy, 64, out_inf, dps[i])) != ECG_OK) return st;
    if ((st = mul_batch_dev(ctx, d, curve, sh.cnt, dps[i])) != ECG_OK) return st;
    if ((st = copy_back(ctx, d, sh, out_xy, 64, out_inf, dps[i])) != ECG_OK) return st;
  }
  return finish(ctx, shards);
}

extern "C" ecg_status ecg_batch_normalize(ecg_ctx* ctx, ecg_curve curve, size_t n, const uint8_t* xyz,
                                          uint8_t* out_xy, uint8_t* out_inf) {
  if (!ctx) return ECG_EINVAL;
  if (n == 0) return ECG_OK;
  if (!xyz || !out_xy || (curve != ECG_SECP256K1 && curve != ECG_NISTP256)) return ECG_EINVAL;
  std::vector<Shard> shards = make_shards(n, ctx->devs.size());
  std::vector<DevPtrs> dps(ctx->devs.size());
  for (size_t i = 0; i < ctx->devs.size(); i++) {
    DevState& d = ctx->devs[i];
    const Shard& sh = shards[i];
    if (sh.cnt == 0) continue;
    CU_TRY(ctx, cudaSetDevice(d.dev));
    ecg_status st;
    if ((st = reset_status(ctx, d)) != ECG_OK) return st;
    if ((st = stage_in(ctx, d, sh, nullptr, 0, xyz, 96, nullptr, dps[i])) != ECG_OK) return st;
    if ((st = stage_out(ctx, d, sh, out_xy, 64, out_inf, dps[i])) != ECG_OK) return st;
    if ((st = ensure(ctx, d, B_JAC, sh.cnt * 96)) != ECG_OK) return st;
    uint32_t* jac = (uint32_t*)d.buf[B_JAC];
    if (curve == ECG_SECP256K1) {
      import_jac_kernel<CurveK256><<<grid_for(sh.cnt, 256), 256, 0, d.s()>>>(dps[i].p, sh.cnt, jac, d.status);
      ctx->launches++;
      CU_TRY(ctx, cudaGetLastError());
      if ((st = launch_normalize<FpK256>(ctx, d, sh.cnt, jac, dps[i].out, dps[i].oinf)) != ECG_OK) return st;
    } else {
      ctx->err = "curve not implemented yet";
      return ECG_EINVAL;
    }
    if ((st = copy_back(ctx, d, sh, out_xy, 64, out_inf, dps[i])) != ECG_OK) return st;
  }
  return finish(ctx, shards);
}

extern "C" ecg_status ecg_field_op_batch(ecg_ctx* ctx, ecg_curve curve, int op, size_t n, const uint8_t* a,
                                         const uint8_t* b, uint8_t* out) {
  if (!ctx) return ECG_EINVAL;
  if (n == 0) return ECG_OK;
  bool binary = (op == ECG_FOP_ADD || op == ECG_FOP_SUB || op == ECG_FOP_MUL);
  if (!a || !out || (binary && !b) || op < 0 || op > ECG_FOP_INV || (curve != ECG_SECP256K1 && curve != ECG_NISTP256))
    return ECG_EINVAL;
  std::vector<Shard> shards = make_shards(n, ctx->devs.size());
  std::vector<DevPtrs> dps(ctx->devs.size());
  for (size_t i = 0; i < ctx->devs.size(); i++) {
    DevState& d = ctx->devs[i];
    const Shard& sh = shards[i];
    if (sh.cnt == 0) continue;
    CU_TRY(ctx, cudaSetDevice(d.dev));
    ecg_status st;
    if ((st = reset_status(ctx, d)) != ECG_OK) return st;
    // a -> B_K slot, b -> B_P slot (both 32-byte strides)
    if ((st = stage_in(ctx, d, sh, a, 32, binary ? b : nullptr, 32, nullptr, dps[i])) != ECG_OK) return st;
    if ((st = stage_out(ctx, d, sh, out, 32, nullptr, dps[i])) != ECG_OK) return st;
    if (curve == ECG_SECP256K1) {
      field_op_kernel<CurveK256><<<grid_for(sh.cnt, 256), 256, 0, d.s()>>>(op, sh.cnt, dps[i].k, dps[i].p, dps[i].out, d.status);
    } else {
      ctx->err = "curve not implemented yet";
      return ECG_EINVAL;
    }
    ctx->launches++;
    CU_TRY(ctx, cudaGetLastError());
    if ((st = copy_back(ctx, d, sh, out, 32, nullptr, dps[i])) != ECG_OK) return st;
  }
  return finish(ctx, shards);
}

extern "C" ecg_status ecg_microbench(ecg_ctx* ctx, int which, int iters, double* ops_per_s, double* elapsed_ms) {
  if (!ctx || !ops_per_s || iters <= 0) return ECG_EINVAL;
  DevState& d = ctx->devs[0];
  CU_TRY(ctx, cudaSetDevice(d.dev));
  ecg_status st = ensure(ctx, d, B_AUX, 256);
  if (st != ECG_OK) return st;
  uint32_t* out = (uint32_t*)d.buf[B_AUX];
  unsigned blocks = (unsigned)d.sm_count * 8, threads = 256;
  double per_thread_iter = 0;
  cudaEvent_t e0, e1;
  CU_TRY(ctx, cudaEventCreate(&e0));
  CU_TRY(ctx, cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {  // rep 0 = warm-up
    CU_TRY(ctx, cudaEventRecord(e0, d.s()));
    switch (which) {
      case 0: mb_imad_wide_kernel<<<blocks, threads, 0, d.s()>>>(out, iters, 12345u + rep); per_thread_iter = 32; break;
      case 1: mb_imad_kernel<<<blocks, threads, 0, d.s()>>>(out, iters, 12345u + rep); per_thread_iter = 64; break;
      case 2: mb_iadd_kernel<<<blocks, threads, 0, d.s()>>>(out, iters, 12345u + rep); per_thread_iter = 64; break;
      case 3: mb_fmul_kernel<FpK256><<<blocks, threads, 0, d.s()>>>(out, iters, 12345u + rep); per_thread_iter = 2; break;
      default:
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return ECG_EINVAL;
    }
    ctx->launches++;
    CU_TRY(ctx, cudaEventRecord(e1, d.s()));
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

// ---- not yet implemented entries (filled in below as the build widens) ------------------------------
extern "C" ecg_status ecg_mul_gen_batch(ecg_ctx* ctx, ecg_curve, size_t, const uint8_t*, uint8_t*, uint8_t*) {
  if (ctx) ctx->err = "ecg_mul_gen_batch: not implemented";
  return ECG_EINVAL;
}
extern "C" ecg_status ecg_lincomb(ecg_ctx* ctx, ecg_curve, size_t, const uint8_t*, const uint8_t*, const uint8_t*, uint8_t*, uint8_t*) {
  if (ctx) ctx->err = "ecg_lincomb: not implemented";
  return ECG_EINVAL;
}
extern "C" ecg_status ecg_lincomb_partial(ecg_ctx* ctx, ecg_curve, size_t, const uint8_t*, const uint8_t*, const uint8_t*, uint8_t*) {
  if (ctx) ctx->err = "ecg_lincomb_partial: not implemented";
  return ECG_EINVAL;
}
extern "C" ecg_status ecg_point_sum(ecg_ctx* ctx, ecg_curve, size_t, const uint8_t*, uint8_t*, uint8_t*) {
  if (ctx) ctx->err = "ecg_point_sum: not implemented";
  return ECG_EINVAL;
}
extern "C" ecg_status ecg_mul_gen_add_batch(ecg_ctx* ctx, ecg_curve, size_t, const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*, uint8_t*, uint8_t*) {
  if (ctx) ctx->err = "ecg_mul_gen_add_batch: not implemented";
  return ECG_EINVAL;
}
