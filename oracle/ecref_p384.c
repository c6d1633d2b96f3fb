/* ecref_p384.c — CPU restatement of the reference's NIST P-384 scalar-multiplication path.
 *
 * TEST INFRASTRUCTURE ONLY: the checker / CPU baseline for the P-384 widening row.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it; the product path never does.
 *
 * What it follows (paths relative to the reference checkout, RustCrypto/elliptic-curves @ 739304e):
 *   field      p384/src/arithmetic/field.rs:35-77: a Montgomery field over U384.  The reference takes the arithmetic either
 *              from fiat-crypto (external, not under /root/reference) or from primefield's generic Montgomery form
 *              (primefield/src/monty.rs:319-375 over crypto-bigint 0.7.5 ConstMontyForm, Cargo.lock:367-368, also
 *              external).  Both compute a*b*R^-1 mod p with R = 2^384 on fully reduced values; restated here as word-by-word
 *              (CIOS) Montgomery multiplication on 6 x 64-bit limbs.
 *   points     primeorder::ProjectivePoint<NistP384> with EquationAIsMinusThree (p384/src/arithmetic.rs:43-44):
 *              primeorder/src/point_arithmetic.rs:222-245 (add, RCB alg. 4), :289-318 (double, RCB alg. 6).
 *   mul        primeorder/src/projective.rs:133-137, 532-557: constant-time lincomb over signed radix-16 digits
 *              (primeorder/src/tables/radix16.rs:31-55: 2*48+1 = 97 digits) with LookupTable select
 *              (primeorder/src/tables/lookup.rs:30-65).
 *   generator  the default backend of p384 is mul_backend::VariableOnly (p384/src/arithmetic.rs:47-51), i.e.
 *              mul_by_generator(k) = GENERATOR * k through the same routine.
 *   to_affine  primeorder/src/projective.rs:101-113 (inversion restated as Fermat exponentiation: the value is unique).
 * Pinned by tests/test_oracle.py to the reference's own vectors (p384/src/test_vectors/group.rs:8,175) and to the
 * big-integer model. */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct {
  uint64_t w[6];
} fp384;

static const uint64_t P[6] = {0x00000000FFFFFFFFULL, 0xFFFFFFFF00000000ULL, 0xFFFFFFFFFFFFFFFEULL,
                              0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL};
static const uint64_t ORDER[6] = {0xECEC196ACCC52973ULL, 0x581A0DB248B0A77AULL, 0xC7634D81F4372DDFULL,
                                  0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL};
#define P_PRIME 0x0000000100000001ULL /* -p^-1 mod 2^64 */
static const fp384 R2 = {{0xFFFFFFFE00000001ULL, 0x0000000200000000ULL, 0xFFFFFFFE00000000ULL, 0x0000000200000000ULL, 1ULL, 0ULL}};
static const fp384 ONE = {{0xFFFFFFFF00000001ULL, 0x00000000FFFFFFFFULL, 1ULL, 0ULL, 0ULL, 0ULL}}; /* R mod p */
static const fp384 ZERO = {{0, 0, 0, 0, 0, 0}};

static int ge6(const uint64_t* a, const uint64_t* m) {
  for (int i = 5; i >= 0; i--)
    if (a[i] != m[i]) return a[i] > m[i];
  return 1;
}
static uint64_t sub6(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t borrow = 0;
  for (int i = 0; i < 6; i++) {
    u128 t = (u128)a[i] - b[i] - borrow;
    r[i] = (uint64_t)t;
    borrow = (uint64_t)(t >> 64) & 1;
  }
  return borrow;
}
static uint64_t add6(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t carry = 0;
  for (int i = 0; i < 6; i++) {
    u128 t = (u128)a[i] + b[i] + carry;
    r[i] = (uint64_t)t;
    carry = (uint64_t)(t >> 64);
  }
  return carry;
}
/* add_mod / sub_mod on fully reduced values (primefield monty add / sub) */
static void fp_add(fp384* r, const fp384* a, const fp384* b) {
  uint64_t s[6], t[6];
  uint64_t c = add6(s, a->w, b->w);
  uint64_t bw = sub6(t, s, P);
  memcpy(r->w, (c || !bw) ? t : s, sizeof t);
}
static void fp_sub(fp384* r, const fp384* a, const fp384* b) {
  uint64_t s[6], t[6];
  uint64_t bw = sub6(s, a->w, b->w);
  add6(t, s, P);
  memcpy(r->w, bw ? t : s, sizeof t);
}
static void fp_dbl(fp384* r, const fp384* a) { fp_add(r, a, a); }
/* Montgomery multiplication, coarsely integrated operand scanning: r = a*b*2^-384 mod p */
static void fp_mul(fp384* r, const fp384* a, const fp384* b) {
  uint64_t t[8] = {0};
  for (int i = 0; i < 6; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 6; j++) {
      u128 v = (u128)a->w[j] * b->w[i] + t[j] + c;
      t[j] = (uint64_t)v;
      c = (uint64_t)(v >> 64);
    }
    u128 v = (u128)t[6] + c;
    t[6] = (uint64_t)v;
    t[7] = (uint64_t)(v >> 64);
    uint64_t m = t[0] * P_PRIME;
    v = (u128)m * P[0] + t[0];
    c = (uint64_t)(v >> 64);
    for (int j = 1; j < 6; j++) {
      v = (u128)m * P[j] + t[j] + c;
      t[j - 1] = (uint64_t)v;
      c = (uint64_t)(v >> 64);
    }
    v = (u128)t[6] + c;
    t[5] = (uint64_t)v;
    t[6] = t[7] + (uint64_t)(v >> 64);
  }
  uint64_t s[6];
  uint64_t bw = sub6(s, t, P);
  memcpy(r->w, (t[6] || !bw) ? s : t, sizeof s);
}
static void fp_sqr(fp384* r, const fp384* a) { fp_mul(r, a, a); }
static int fp_is_zero(const fp384* a) { return (a->w[0] | a->w[1] | a->w[2] | a->w[3] | a->w[4] | a->w[5]) == 0; }
static void limbs_from_be(uint64_t* w, const uint8_t* b) {
  for (int i = 0; i < 6; i++) {
    uint64_t v = 0;
    for (int j = 0; j < 8; j++) v = (v << 8) | b[8 * (5 - i) + j];
    w[i] = v;
  }
}
static void fp_from_be(fp384* r, const uint8_t* b) { /* from_uint_unchecked: a * R^2 * R^-1 */
  fp384 t;
  limbs_from_be(t.w, b);
  fp_mul(r, &t, &R2);
}
static void fp_to_be(uint8_t* b, const fp384* a) { /* to_canonical: a * 1 * R^-1 */
  fp384 one = {{1, 0, 0, 0, 0, 0}}, t;
  fp_mul(&t, a, &one);
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 8; j++) b[8 * (5 - i) + j] = (uint8_t)(t.w[i] >> (56 - 8 * j));
}
/* a^(p-2) by square-and-multiply over the bits of p - 2 (public exponent) */
static void fp_inv(fp384* r, const fp384* a) {
  uint64_t e[6];
  memcpy(e, P, sizeof e);
  e[0] -= 2;
  fp384 acc = ONE;
  for (int bit = 383; bit >= 0; bit--) {
    fp_sqr(&acc, &acc);
    if ((e[bit >> 6] >> (bit & 63)) & 1) fp_mul(&acc, &acc, a);
  }
  *r = acc;
}

typedef struct {
  fp384 x, y, z;
} qpt;
static fp384 B_M, GX_M, GY_M; /* EQUATION_B, GENERATOR in Montgomery form (p384/src/arithmetic.rs:56-74) */
static qpt Q_GEN;

static void qpt_identity(qpt* r) {
  r->x = ZERO;
  r->y = ONE;
  r->z = ZERO;
}
/* point_arithmetic.rs:222-245 add_assign (RCB algorithm 4, a = -3) */
static void qpt_add(qpt* r, const qpt* p, const qpt* q) {
  fp384 xx, yy, zz, xy_pairs, yz_pairs, xz_pairs, t, u;
  fp_mul(&xx, &p->x, &q->x);
  fp_mul(&yy, &p->y, &q->y);
  fp_mul(&zz, &p->z, &q->z);
  fp_add(&t, &p->x, &p->y);
  fp_add(&u, &q->x, &q->y);
  fp_mul(&xy_pairs, &t, &u);
  fp_add(&t, &xx, &yy);
  fp_sub(&xy_pairs, &xy_pairs, &t);
  fp_add(&t, &p->y, &p->z);
  fp_add(&u, &q->y, &q->z);
  fp_mul(&yz_pairs, &t, &u);
  fp_add(&t, &yy, &zz);
  fp_sub(&yz_pairs, &yz_pairs, &t);
  fp_add(&t, &p->x, &p->z);
  fp_add(&u, &q->x, &q->z);
  fp_mul(&xz_pairs, &t, &u);
  fp_add(&t, &xx, &zz);
  fp_sub(&xz_pairs, &xz_pairs, &t);

  fp384 bzz_part, bzz3_part, yy_m_bzz3, yy_p_bzz3, zz3, bxz_part, bxz3_part, xx3_m_zz3;
  fp_mul(&t, &B_M, &zz);
  fp_sub(&bzz_part, &xz_pairs, &t);
  fp_dbl(&t, &bzz_part);
  fp_add(&bzz3_part, &t, &bzz_part);
  fp_sub(&yy_m_bzz3, &yy, &bzz3_part);
  fp_add(&yy_p_bzz3, &yy, &bzz3_part);
  fp_dbl(&t, &zz);
  fp_add(&zz3, &t, &zz);
  fp_mul(&t, &B_M, &xz_pairs);
  fp_add(&u, &zz3, &xx);
  fp_sub(&bxz_part, &t, &u);
  fp_dbl(&t, &bxz_part);
  fp_add(&bxz3_part, &t, &bxz_part);
  fp_dbl(&t, &xx);
  fp_add(&t, &t, &xx);
  fp_sub(&xx3_m_zz3, &t, &zz3);

  fp384 rx, ry, rz;
  fp_mul(&t, &yy_p_bzz3, &xy_pairs);
  fp_mul(&u, &yz_pairs, &bxz3_part);
  fp_sub(&rx, &t, &u);
  fp_mul(&t, &yy_p_bzz3, &yy_m_bzz3);
  fp_mul(&u, &xx3_m_zz3, &bxz3_part);
  fp_add(&ry, &t, &u);
  fp_mul(&t, &yy_m_bzz3, &yz_pairs);
  fp_mul(&u, &xy_pairs, &xx3_m_zz3);
  fp_add(&rz, &t, &u);
  r->x = rx;
  r->y = ry;
  r->z = rz;
}
/* point_arithmetic.rs:289-318 double_in_place (RCB algorithm 6, a = -3) */
static void qpt_dbl(qpt* r, const qpt* p) {
  fp384 xx, yy, zz, xy2, xz2, t, u;
  fp_sqr(&xx, &p->x);
  fp_sqr(&yy, &p->y);
  fp_sqr(&zz, &p->z);
  fp_mul(&xy2, &p->x, &p->y);
  fp_dbl(&xy2, &xy2);
  fp_mul(&xz2, &p->x, &p->z);
  fp_dbl(&xz2, &xz2);
  fp384 bzz_part, bzz3_part, yy_m_bzz3, yy_p_bzz3, y_frag, x_frag, zz3, bxz2_part, bxz6_part, xx3_m_zz3;
  fp_mul(&t, &B_M, &zz);
  fp_sub(&bzz_part, &t, &xz2);
  fp_dbl(&t, &bzz_part);
  fp_add(&bzz3_part, &t, &bzz_part);
  fp_sub(&yy_m_bzz3, &yy, &bzz3_part);
  fp_add(&yy_p_bzz3, &yy, &bzz3_part);
  fp_mul(&y_frag, &yy_p_bzz3, &yy_m_bzz3);
  fp_mul(&x_frag, &yy_m_bzz3, &xy2);
  fp_dbl(&t, &zz);
  fp_add(&zz3, &t, &zz);
  fp_mul(&t, &B_M, &xz2);
  fp_add(&u, &zz3, &xx);
  fp_sub(&bxz2_part, &t, &u);
  fp_dbl(&t, &bxz2_part);
  fp_add(&bxz6_part, &t, &bxz2_part);
  fp_dbl(&t, &xx);
  fp_add(&t, &t, &xx);
  fp_sub(&xx3_m_zz3, &t, &zz3);
  fp384 rx, ry, rz, yz2;
  fp_mul(&t, &xx3_m_zz3, &bxz6_part);
  fp_add(&ry, &y_frag, &t);
  fp_mul(&yz2, &p->y, &p->z);
  fp_dbl(&yz2, &yz2);
  fp_mul(&t, &bxz6_part, &yz2);
  fp_sub(&rx, &x_frag, &t);
  fp_mul(&t, &yz2, &yy);
  fp_dbl(&t, &t);
  fp_dbl(&rz, &t);
  r->x = rx;
  r->y = ry;
  r->z = rz;
}
static void qpt_neg(qpt* r, const qpt* p) {
  r->x = p->x;
  fp_sub(&r->y, &ZERO, &p->y);
  r->z = p->z;
}

/* radix16.rs:31-55 Radix16Decomposition::new — 97 digits from 48 big-endian bytes */
#define ND 97
static void radix16(int8_t* d, const uint8_t* be48) {
  memset(d, 0, ND);
  for (int i = 0; i < 48; i++) {
    uint8_t b = be48[47 - i];
    d[2 * i] = (int8_t)(b & 0xf);
    d[2 * i + 1] = (int8_t)((b >> 4) & 0xf);
  }
  for (int i = 0; i < ND - 1; i++) {
    int8_t carry = (int8_t)((d[i] + 8) >> 4);
    d[i] = (int8_t)(d[i] - (carry << 4));
    d[i + 1] = (int8_t)(d[i + 1] + carry);
  }
}
typedef struct {
  qpt t[8];
  int8_t d[ND];
} qterm;
/* lookup.rs:30-38 LookupTable::new */
static void lut_new(qpt* tab, const qpt* p) {
  tab[0] = *p;
  for (int j = 0; j < 7; j++) qpt_add(&tab[j + 1], p, &tab[j]);
}
/* lookup.rs:43-65 select */
static void lut_select(qpt* r, const qpt* tab, int8_t x) {
  int8_t xmask = (int8_t)(x >> 7);
  int8_t xabs = (int8_t)((x + xmask) ^ xmask);
  qpt t;
  qpt_identity(&t);
  for (int j = 1; j <= 8; j++)
    if (xabs == j) t = tab[j - 1];
  if (xmask & 1) {
    qpt n;
    qpt_neg(&n, &t);
    t = n;
  }
  *r = t;
}
/* projective.rs:532-557 lincomb (constant time): shared doublings, one table per term */
static void lincomb_ct(qpt* out, const qterm* T, size_t nterms) {
  qpt q, t;
  qpt_identity(&q);
  for (size_t c = 0; c < nterms; c++) {
    lut_select(&t, T[c].t, T[c].d[ND - 1]);
    qpt_add(&q, &q, &t);
  }
  for (int i = ND - 2; i >= 0; i--) {
    for (int j = 0; j < 4; j++) qpt_dbl(&q, &q);
    for (size_t c = 0; c < nterms; c++) {
      lut_select(&t, T[c].t, T[c].d[i]);
      qpt_add(&q, &q, &t);
    }
  }
  *out = q;
}

static int g_init = 0;
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static void hexbe(uint8_t* out, const char* h, int n) {
  for (int i = 0; i < n; i++) {
    unsigned v = 0;
    for (int j = 0; j < 2; j++) {
      char c = h[2 * i + j];
      v = v * 16 + (unsigned)(c <= '9' ? c - '0' : (c | 32) - 'a' + 10);
    }
    out[i] = (uint8_t)v;
  }
}
void ecref384_init(void) {
  pthread_mutex_lock(&g_lock);
  if (!g_init) {
    uint8_t b[48];
    hexbe(b, "b3312fa7e23ee7e4988e056be3f82d19181d9c6efe8141120314088f5013875ac656398d8a2ed19d2a85c8edd3ec2aef", 48);
    fp_from_be(&B_M, b);
    hexbe(b, "aa87ca22be8b05378eb1c71ef320ad746e1d3b628ba79b9859f741e082542a385502f25dbf55296c3a545e3872760ab7", 48);
    fp_from_be(&GX_M, b);
    hexbe(b, "3617de4a96262c6f5d9e98bf9292dc29f8f41dbd289a147ce9da3113b5f0b8c00a60b1ce1d7e819d7a431d7c90ea0e5f", 48);
    fp_from_be(&GY_M, b);
    Q_GEN.x = GX_M;
    Q_GEN.y = GY_M;
    Q_GEN.z = ONE;
    g_init = 1;
  }
  pthread_mutex_unlock(&g_lock);
}

/* decode + validate like Scalar::from_repr / AffinePoint::from_coordinates; 0 ok, 2 scalar range, 3 not on curve */
static int load_point(qpt* out, const uint8_t* xy, int inf) {
  if (inf) {
    qpt_identity(out);
    return 0;
  }
  uint64_t x[6], y[6];
  limbs_from_be(x, xy);
  limbs_from_be(y, xy + 48);
  if (ge6(x, P) || ge6(y, P)) return 3;
  fp_from_be(&out->x, xy);
  fp_from_be(&out->y, xy + 48);
  out->z = ONE;
  fp384 l, r, t;
  fp_sqr(&l, &out->y);
  fp_sqr(&r, &out->x);
  fp_mul(&r, &r, &out->x);
  fp_dbl(&t, &out->x);
  fp_add(&t, &t, &out->x);
  fp_sub(&r, &r, &t);
  fp_add(&r, &r, &B_M);
  fp_sub(&l, &l, &r);
  return fp_is_zero(&l) ? 0 : 3;
}
static void store_point(uint8_t* xy, uint8_t* inf, const qpt* p) { /* to_affine; identity = zero bytes + flag */
  if (fp_is_zero(&p->z)) {
    memset(xy, 0, 96);
    *inf = 1;
    return;
  }
  fp384 zi, x, y;
  fp_inv(&zi, &p->z);
  fp_mul(&x, &p->x, &zi);
  fp_mul(&y, &p->y, &zi);
  fp_to_be(xy, &x);
  fp_to_be(xy + 48, &y);
  *inf = 0;
}
static int scalar_ok(const uint8_t* k) {
  uint64_t w[6];
  limbs_from_be(w, k);
  return !ge6(w, ORDER);
}

enum { OP_MUL = 0, OP_MULGEN = 1, OP_LINCOMB = 2 };
#define LINCOMB_CHUNK 256
typedef struct {
  int op;
  size_t lo, hi;
  const uint8_t *k, *pxy, *pinf;
  uint8_t *oxy, *oinf;
  int err;
  size_t err_index;
  qpt acc;
} job;

static void* worker(void* arg) {
  job* j = (job*)arg;
  j->err = 0;
  if (j->op == OP_LINCOMB) {
    qpt_identity(&j->acc);
    qterm* T = (qterm*)malloc(sizeof(qterm) * LINCOMB_CHUNK);
    for (size_t base = j->lo; base < j->hi && !j->err; base += LINCOMB_CHUNK) {
      size_t cnt = j->hi - base < LINCOMB_CHUNK ? j->hi - base : LINCOMB_CHUNK;
      for (size_t c = 0; c < cnt; c++) {
        size_t i = base + c;
        qpt pt;
        int e = scalar_ok(j->k + 48 * i) ? load_point(&pt, j->pxy + 96 * i, j->pinf ? j->pinf[i] : 0) : 2;
        if (e) {
          j->err = e;
          j->err_index = i;
          break;
        }
        lut_new(T[c].t, &pt);
        radix16(T[c].d, j->k + 48 * i);
      }
      if (j->err) break;
      qpt part;
      lincomb_ct(&part, T, cnt);
      qpt_add(&j->acc, &j->acc, &part);
    }
    free(T);
    return NULL;
  }
  for (size_t i = j->lo; i < j->hi; i++) {
    qpt pt, r;
    int e = scalar_ok(j->k + 48 * i) ? 0 : 2;
    if (!e) {
      if (j->op == OP_MUL)
        e = load_point(&pt, j->pxy + 96 * i, j->pinf ? j->pinf[i] : 0);
      else
        pt = Q_GEN;
    }
    if (e) {
      j->err = e;
      j->err_index = i;
      return NULL;
    }
    qterm T;
    lut_new(T.t, &pt);
    radix16(T.d, j->k + 48 * i);
    lincomb_ct(&r, &T, 1);
    store_point(j->oxy + 96 * i, j->oinf + i, &r);
  }
  return NULL;
}
static int run_jobs(job* tmpl, size_t n, int nthreads, job** out_jobs) {
  ecref384_init();
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > n && n > 0) nthreads = (int)n;
  job* jobs = (job*)calloc((size_t)nthreads, sizeof(job));
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  size_t base = n / (size_t)nthreads, rem = n % (size_t)nthreads, off = 0;
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = *tmpl;
    jobs[t].lo = off;
    off += base + ((size_t)t < rem ? 1 : 0);
    jobs[t].hi = off;
  }
  for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, worker, &jobs[t]);
  worker(&jobs[0]);
  for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
  int err = 0;
  for (int t = 0; t < nthreads; t++)
    if (jobs[t].err && !err) err = jobs[t].err;
  free(th);
  if (out_jobs)
    *out_jobs = jobs;
  else
    free(jobs);
  return err;
}

/* out[i] = k[i] * P[i]   (48-byte scalars, 96-byte points, flag bytes; as include/ecgpu.h with ECG_NISTP384) */
int ecref384_mul_batch(size_t n, const uint8_t* k, const uint8_t* pxy, const uint8_t* pinf, uint8_t* oxy, uint8_t* oinf, int nthreads) {
  if (n == 0) return 0;
  job t;
  memset(&t, 0, sizeof t);
  t.op = OP_MUL;
  t.k = k;
  t.pxy = pxy;
  t.pinf = pinf;
  t.oxy = oxy;
  t.oinf = oinf;
  return run_jobs(&t, n, nthreads, NULL);
}
int ecref384_mul_gen_batch(size_t n, const uint8_t* k, uint8_t* oxy, uint8_t* oinf, int nthreads) {
  if (n == 0) return 0;
  job t;
  memset(&t, 0, sizeof t);
  t.op = OP_MULGEN;
  t.k = k;
  t.oxy = oxy;
  t.oinf = oinf;
  return run_jobs(&t, n, nthreads, NULL);
}
/* out = sum_i k[i] * P[i]: reference-style lincomb calls of LINCOMB_CHUNK terms, partial sums added */
int ecref384_lincomb(size_t n, const uint8_t* k, const uint8_t* pxy, const uint8_t* pinf, uint8_t* oxy, uint8_t* oinf, int nthreads) {
  ecref384_init();
  qpt acc;
  qpt_identity(&acc);
  if (n > 0) {
    job t, *jobs = NULL;
    memset(&t, 0, sizeof t);
    t.op = OP_LINCOMB;
    t.k = k;
    t.pxy = pxy;
    t.pinf = pinf;
    int nt = nthreads < 1 ? 1 : nthreads;
    if ((size_t)nt > n) nt = (int)n;
    int err = run_jobs(&t, n, nt, &jobs);
    if (err) {
      free(jobs);
      return err;
    }
    for (int i = 0; i < nt; i++) qpt_add(&acc, &acc, &jobs[i].acc);
    free(jobs);
  }
  store_point(oxy, oinf, &acc);
  return 0;
}
