/* oracle/ecref.c — CPU restatement of the reference's scalar-multiplication path (k256 + p256).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load libecref.so.  The product (libecgpu.so) never links or calls this.
 *
 * What it is: a plain-C (gcc, unsigned __int128) restatement of the algorithms in
 * RustCrypto/elliptic-curves @ 739304e, function by function, each citing the reference file:line it
 * follows.  The Rust crate itself cannot be built in this environment (no rustc/cargo, dependencies not
 * vendored), so parity is pinned instead by the reference's own golden vectors (tests/golden/{k256,p256}.json,
 * extracted from k256/src/test_vectors/group.rs:9,96, p256/src/test_vectors/group.rs:8,95,
 * {k256,p256}/src/test_vectors/field.rs:6) — see tests/test_oracle.py.
 *
 * Third-party pieces the reference takes from crypto-bigint 0.7.5 (Cargo.lock:367-368), restated from
 * their mathematical definition: U256::widening_mul (4x4 schoolbook), invert_odd_mod (here: Fermat
 * exponentiation — the inverse is unique, so the value is identical).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ================================================================================================
 * secp256k1 base field, 5 x 52-bit limbs, lazily reduced
 * (k256/src/arithmetic/field/field_5x52.rs)
 * ================================================================================================ */
typedef struct {
  uint64_t n[5];
} fe52;

#define M52 0xFFFFFFFFFFFFFULL
#define R52 0x1000003D10ULL /* 2^260 mod p  (field_5x52.rs:263) */

/* field_5x52.rs:63-74 from_u256_unchecked (bytes big-endian, field_5x52.rs:26-38) */
static void fe52_from_be(fe52* r, const uint8_t* b) {
  uint64_t w[4];
  for (int i = 0; i < 4; i++) {
    uint64_t v = 0;
    for (int j = 0; j < 8; j++) v = (v << 8) | b[(3 - i) * 8 + j];
    w[i] = v;
  }
  r->n[0] = w[0] & M52;
  r->n[1] = ((w[0] >> 52) | (w[1] << 12)) & M52;
  r->n[2] = ((w[1] >> 40) | (w[2] << 24)) & M52;
  r->n[3] = ((w[2] >> 28) | (w[3] << 36)) & M52;
  r->n[4] = w[3] >> 16;
}
/* field_5x52.rs:43-61 to_u256 / to_bytes; input must be normalized */
static void fe52_to_be(uint8_t* b, const fe52* a) {
  uint64_t w[4];
  w[0] = a->n[0] | (a->n[1] << 52);
  w[1] = (a->n[1] >> 12) | (a->n[2] << 40);
  w[2] = (a->n[2] >> 24) | (a->n[3] << 28);
  w[3] = (a->n[3] >> 36) | (a->n[4] << 16);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 8; j++) b[(3 - i) * 8 + j] = (uint8_t)(w[i] >> (56 - 8 * j));
}
/* field_5x52.rs:82-101 add_modulus_correction */
static void fe52_add_corr(fe52* r, uint64_t x) {
  uint64_t t0 = r->n[0] + x * 0x1000003D1ULL;
  uint64_t t1 = r->n[1] + (t0 >> 52);
  uint64_t t2 = r->n[2] + (t1 >> 52);
  uint64_t t3 = r->n[3] + (t2 >> 52);
  uint64_t t4 = r->n[4] + (t3 >> 52);
  r->n[0] = t0 & M52;
  r->n[1] = t1 & M52;
  r->n[2] = t2 & M52;
  r->n[3] = t3 & M52;
  r->n[4] = t4;
}
/* field_5x52.rs:122-133 normalize_weak */
static void fe52_normalize_weak(fe52* r) {
  uint64_t x = r->n[4] >> 48;
  r->n[4] &= 0x0FFFFFFFFFFFFULL;
  fe52_add_corr(r, x);
}
/* field_5x52.rs:112-119 get_overflow */
static int fe52_overflow(const fe52* a) {
  uint64_t m = a->n[1] & a->n[2] & a->n[3];
  return ((a->n[4] >> 48) != 0) | ((a->n[4] == 0x0FFFFFFFFFFFFULL) & (m == M52) & (a->n[0] >= 0xFFFFEFFFFFC2FULL));
}
/* field_5x52.rs:138-155 normalize */
static void fe52_normalize(fe52* r) {
  fe52_normalize_weak(r);
  if (fe52_overflow(r)) {
    fe52_add_corr(r, 1);
    r->n[4] &= 0x0FFFFFFFFFFFFULL;
  }
}
/* field_5x52.rs:158-172 normalizes_to_zero */
static int fe52_normalizes_to_zero(const fe52* a) {
  fe52 t = *a;
  fe52_normalize_weak(&t);
  uint64_t z0 = t.n[0] | t.n[1] | t.n[2] | t.n[3] | t.n[4];
  uint64_t z1 = (t.n[0] ^ 0x1000003D0ULL) & t.n[1] & t.n[2] & t.n[3] & (t.n[4] ^ 0xF000000000000ULL);
  return (z0 == 0) | (z1 == M52);
}
/* field_5x52.rs:203-211 negate(magnitude) */
static void fe52_negate(fe52* r, const fe52* a, unsigned mag) {
  uint64_t m = (uint64_t)mag + 1;
  r->n[0] = 0xFFFFEFFFFFC2FULL * 2 * m - a->n[0];
  r->n[1] = M52 * 2 * m - a->n[1];
  r->n[2] = M52 * 2 * m - a->n[2];
  r->n[3] = M52 * 2 * m - a->n[3];
  r->n[4] = 0x0FFFFFFFFFFFFULL * 2 * m - a->n[4];
}
/* field_5x52.rs:215-223 add */
static void fe52_add(fe52* r, const fe52* a, const fe52* b) {
  for (int i = 0; i < 5; i++) r->n[i] = a->n[i] + b->n[i];
}
/* field_5x52.rs:227-236 mul_single */
static void fe52_mul_single(fe52* r, const fe52* a, uint32_t k) {
  for (int i = 0; i < 5; i++) r->n[i] = a->n[i] * k;
}
/* field_5x52.rs:240-401 mul_inner: 25 64x64->128 products, high columns folded with R = 2^260 mod p
 * 52 bits at a time so every accumulator stays below 2^128 for input magnitudes up to 8. */
static void fe52_mul(fe52* r, const fe52* A, const fe52* B) {
  const uint64_t a0 = A->n[0], a1 = A->n[1], a2 = A->n[2], a3 = A->n[3], a4 = A->n[4];
  const uint64_t b0 = B->n[0], b1 = B->n[1], b2 = B->n[2], b3 = B->n[3], b4 = B->n[4];
#define P(x, y) ((u128)(x) * (y))
  u128 c, d;
  uint64_t t3, t4, tx, u0, r0, r1, r2;
  d = P(a0, b3) + P(a1, b2) + P(a2, b1) + P(a3, b0); /* column 3 */
  c = P(a4, b4);                                     /* column 8 */
  d += (u128)((uint64_t)c & M52) * R52;
  c >>= 52;
  t3 = (uint64_t)d & M52;
  d >>= 52;
  d += P(a0, b4) + P(a1, b3) + P(a2, b2) + P(a3, b1) + P(a4, b0); /* column 4 */
  d += (u128)(uint64_t)c * R52;
  t4 = (uint64_t)d & M52;
  d >>= 52;
  tx = t4 >> 48;
  t4 &= (M52 >> 4);
  c = P(a0, b0);                                      /* column 0 */
  d += P(a1, b4) + P(a2, b3) + P(a3, b2) + P(a4, b1); /* column 5 */
  u0 = (uint64_t)d & M52;
  d >>= 52;
  u0 = (u0 << 4) | tx;
  c += (u128)u0 * (R52 >> 4);
  r0 = (uint64_t)c & M52;
  c >>= 52;
  c += P(a0, b1) + P(a1, b0);            /* column 1 */
  d += P(a2, b4) + P(a3, b3) + P(a4, b2); /* column 6 */
  c += (u128)((uint64_t)d & M52) * R52;
  d >>= 52;
  r1 = (uint64_t)c & M52;
  c >>= 52;
  c += P(a0, b2) + P(a1, b1) + P(a2, b0); /* column 2 */
  d += P(a3, b4) + P(a4, b3);             /* column 7 */
  c += (u128)((uint64_t)d & M52) * R52;
  d >>= 52;
  r2 = (uint64_t)c & M52;
  c >>= 52;
  c += (u128)(uint64_t)d * R52 + t3;
  r->n[0] = r0;
  r->n[1] = r1;
  r->n[2] = r2;
  r->n[3] = (uint64_t)c & M52;
  c >>= 52;
  r->n[4] = (uint64_t)c + t4;
#undef P
}
/* field_5x52.rs:414-416 square = mul_inner(self, self) */
static void fe52_sqr(fe52* r, const fe52* a) { fe52_mul(r, a, a); }

static const fe52 FE52_ZERO = {{0, 0, 0, 0, 0}};
static const fe52 FE52_ONE = {{1, 0, 0, 0, 0}};

/* k256/src/arithmetic/field.rs:178-184 invert -> crypto-bigint invert_odd_mod (third-party);
 * restated as a^(p-2).  Returns 0 for a == 0 (the reference returns CtOption::none). */
static void fe52_sqrn(fe52* r, const fe52* a, int n) {
  *r = *a;
  for (int i = 0; i < n; i++) fe52_sqr(r, r);
}
static void fe52_inv(fe52* r, const fe52* a) {
  fe52 x2, x3, x6, x9, x11, x22, x44, x88, x176, x220, x223, t;
  fe52_sqr(&x2, a);
  fe52_mul(&x2, &x2, a);
  fe52_sqr(&x3, &x2);
  fe52_mul(&x3, &x3, a);
  fe52_sqrn(&x6, &x3, 3);
  fe52_mul(&x6, &x6, &x3);
  fe52_sqrn(&x9, &x6, 3);
  fe52_mul(&x9, &x9, &x3);
  fe52_sqrn(&x11, &x9, 2);
  fe52_mul(&x11, &x11, &x2);
  fe52_sqrn(&x22, &x11, 11);
  fe52_mul(&x22, &x22, &x11);
  fe52_sqrn(&x44, &x22, 22);
  fe52_mul(&x44, &x44, &x22);
  fe52_sqrn(&x88, &x44, 44);
  fe52_mul(&x88, &x88, &x44);
  fe52_sqrn(&x176, &x88, 88);
  fe52_mul(&x176, &x176, &x88);
  fe52_sqrn(&x220, &x176, 44);
  fe52_mul(&x220, &x220, &x44);
  fe52_sqrn(&x223, &x220, 3);
  fe52_mul(&x223, &x223, &x3);
  fe52_sqrn(&t, &x223, 23);
  fe52_mul(&t, &t, &x22);
  fe52_sqrn(&t, &t, 5);
  fe52_mul(&t, &t, a);
  fe52_sqrn(&t, &t, 3);
  fe52_mul(&t, &t, &x2);
  fe52_sqrn(&t, &t, 2);
  fe52_mul(r, &t, a);
}

/* ================================================================================================
 * secp256k1 scalars mod n, 4 x 64 limbs (k256/src/arithmetic/scalar.rs, scalar/wide64.rs)
 * ================================================================================================ */
typedef struct {
  uint64_t w[4];
} sc256;

static const uint64_t K256_N[4] = {0xBFD25E8CD0364141ULL, 0xBAAEDCE6AF48A03BULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL};
/* 2^256 - n  (wide64.rs:11 NEG_MODULUS) */
static const uint64_t K256_NC[3] = {0x402DA1732FC9BEBFULL, 0x4551231950B75FC4ULL, 1ULL};

static void sc_from_be(uint64_t* w, const uint8_t* b) {
  for (int i = 0; i < 4; i++) {
    uint64_t v = 0;
    for (int j = 0; j < 8; j++) v = (v << 8) | b[(3 - i) * 8 + j];
    w[i] = v;
  }
}
static void sc_to_be(uint8_t* b, const uint64_t* w) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 8; j++) b[(3 - i) * 8 + j] = (uint8_t)(w[i] >> (56 - 8 * j));
}
static int limbs_ge(const uint64_t* a, const uint64_t* m, int n) {
  for (int i = n - 1; i >= 0; i--) {
    if (a[i] > m[i]) return 1;
    if (a[i] < m[i]) return 0;
  }
  return 1;
}
static uint64_t limbs_sub(uint64_t* r, const uint64_t* a, const uint64_t* b, int n) {
  uint64_t borrow = 0;
  for (int i = 0; i < n; i++) {
    u128 t = (u128)a[i] - b[i] - borrow;
    r[i] = (uint64_t)t;
    borrow = (uint64_t)(t >> 64) & 1;
  }
  return borrow;
}
static uint64_t limbs_add(uint64_t* r, const uint64_t* a, const uint64_t* b, int n) {
  uint64_t carry = 0;
  for (int i = 0; i < n; i++) {
    u128 t = (u128)a[i] + b[i] + carry;
    r[i] = (uint64_t)t;
    carry = (uint64_t)(t >> 64);
  }
  return carry;
}
/* wide64.rs:23-59 mul_wide: 4x4 schoolbook into 8 limbs */
static void sc_mul_wide(uint64_t* l, const uint64_t* a, const uint64_t* b) {
  memset(l, 0, 8 * sizeof(uint64_t));
  for (int i = 0; i < 4; i++) {
    uint64_t carry = 0;
    for (int j = 0; j < 4; j++) {
      u128 t = (u128)a[i] * b[j] + l[i + j] + carry;
      l[i + j] = (uint64_t)t;
      carry = (uint64_t)(t >> 64);
    }
    l[i + 4] = carry;
  }
}
/* wide64.rs:121-212 reduce_impl: fold the high half with NEG_MODULUS = 2^256 - n (512 -> 385 -> 258 -> 256
 * bits), then one conditional subtraction.  Restated as a loop over the same identity 2^256 == NC (mod n). */
static void sc_reduce512(uint64_t* r, const uint64_t* l) {
  uint64_t cur[8];
  memcpy(cur, l, sizeof cur);
  for (;;) {
    int hi_nonzero = (cur[4] | cur[5] | cur[6] | cur[7]) != 0;
    if (!hi_nonzero) break;
    /* cur = lo + hi * NC */
    uint64_t prod[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
      uint64_t carry = 0;
      for (int j = 0; j < 3; j++) {
        u128 t = (u128)cur[4 + i] * K256_NC[j] + prod[i + j] + carry;
        prod[i + j] = (uint64_t)t;
        carry = (uint64_t)(t >> 64);
      }
      prod[i + 3] += carry; /* cannot overflow: hi*NC < 2^(256+129) */
    }
    uint64_t lo[8] = {cur[0], cur[1], cur[2], cur[3], 0, 0, 0, 0};
    limbs_add(cur, lo, prod, 8);
  }
  while (limbs_ge(cur, K256_N, 4)) limbs_sub(cur, cur, K256_N, 4);
  memcpy(r, cur, 4 * sizeof(uint64_t));
}
/* scalar.rs:120-122 mul */
static void sc_mul(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t l[8];
  sc_mul_wide(l, a, b);
  sc_reduce512(r, l);
}
/* scalar.rs:102-105 add (mod n) */
static void sc_add(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t t[5];
  t[4] = limbs_add(t, a, b, 4);
  uint64_t n5[5] = {K256_N[0], K256_N[1], K256_N[2], K256_N[3], 0};
  if (limbs_ge(t, n5, 5)) limbs_sub(t, t, n5, 5);
  memcpy(r, t, 4 * sizeof(uint64_t));
}
/* scalar.rs:107-110 negate */
static void sc_neg(uint64_t* r, const uint64_t* a) {
  if ((a[0] | a[1] | a[2] | a[3]) == 0) {
    memset(r, 0, 32);
    return;
  }
  limbs_sub(r, K256_N, a, 4);
}
/* scalar.rs:419-423 is_high: a > n/2 */
static int sc_is_high(const uint64_t* a) {
  static const uint64_t HALF[4] = {0xDFE92F46681B20A0ULL, 0x5D576E7357A4501DULL, 0xFFFFFFFFFFFFFFFFULL, 0x7FFFFFFFFFFFFFFFULL};
  /* a > floor(n/2) */
  for (int i = 3; i >= 0; i--) {
    if (a[i] > HALF[i]) return 1;
    if (a[i] < HALF[i]) return 0;
  }
  return 0;
}
/* wide64.rs:64-119 mul_shift_vartime(a, b, 384): (a*b) >> 384 rounded to nearest */
static void sc_mul_shift384(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t l[8];
  sc_mul_wide(l, a, b);
  uint64_t res[4] = {l[6], l[7], 0, 0};
  uint64_t one[4] = {(l[5] >> 63) & 1, 0, 0, 0};
  sc_add(r, res, one);
}
/* k256/src/arithmetic/mul/glv.rs:10-37 constants */
static const uint64_t GLV_MINUS_LAMBDA[4] = {0xE0CFC810B51283CFULL, 0xA880B9FC8EC739C2ULL, 0x5AD9E3FD77ED9BA4ULL, 0xAC9C52B33FA3CF1FULL};
static const uint64_t GLV_MINUS_B1[4] = {0x6F547FA90ABFE4C3ULL, 0xE4437ED6010E8828ULL, 0, 0};
static const uint64_t GLV_MINUS_B2[4] = {0xD765CDA83DB1562CULL, 0x8A280AC50774346DULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL};
static const uint64_t GLV_G1[4] = {0xE893209A45DBB031ULL, 0x3DAA8A1471E8CA7FULL, 0xE86C90E49284EB15ULL, 0x3086D221A7D46BCDULL};
static const uint64_t GLV_G2[4] = {0x1571B4AE8AC47F71ULL, 0x221208AC9DF506C6ULL, 0x6F547FA90ABFE4C4ULL, 0xE4437ED6010E8828ULL};
/* glv.rs:149-156 decompose_scalar */
static void glv_decompose(uint64_t* r1, uint64_t* r2, const uint64_t* k) {
  uint64_t c1[4], c2[4], t[4];
  sc_mul_shift384(c1, k, GLV_G1);
  sc_mul(c1, c1, GLV_MINUS_B1);
  sc_mul_shift384(c2, k, GLV_G2);
  sc_mul(c2, c2, GLV_MINUS_B2);
  sc_add(r2, c1, c2);
  sc_mul(t, r2, GLV_MINUS_LAMBDA);
  sc_add(r1, k, t);
}

/* ================================================================================================
 * secp256k1 points: homogeneous projective, RCB-2015 complete formulas for a = 0, b = 7
 * (k256/src/arithmetic/projective.rs)
 * ================================================================================================ */
typedef struct {
  fe52 x, y, z;
} kpt;
typedef struct {
  fe52 x, y;
  int inf;
} kaff;

#define KB 7u /* CURVE_EQUATION_B_SINGLE */

static const fe52 K_BETA = {{0x96C28719501EEULL, 0x7512F58995C13ULL, 0xC3434E99CF049ULL, 0x7106E64479EAULL, 0x7AE96A2B657CULL}};

static void kpt_identity(kpt* r) { /* projective.rs:49-53 */
  r->x = FE52_ZERO;
  r->y = FE52_ONE;
  r->z = FE52_ZERO;
}
/* projective.rs:79-85 neg */
static void kpt_neg(kpt* r, const kpt* p) {
  r->x = p->x;
  fe52_negate(&r->y, &p->y, 1);
  fe52_normalize_weak(&r->y);
  r->z = p->z;
}
/* projective.rs:96-131 add_assign (RCB algorithm 7) */
static void kpt_add(kpt* r, const kpt* p, const kpt* q) {
  fe52 xx, yy, zz, n_xx_yy, n_yy_zz, n_xx_zz, xy_pairs, yz_pairs, xz_pairs, t, u;
  fe52 bzz, bzz3, yy_m_bzz3, yy_p_bzz3, byz, byz3, xx3, bxx9;
  fe52_mul(&xx, &p->x, &q->x);
  fe52_mul(&yy, &p->y, &q->y);
  fe52_mul(&zz, &p->z, &q->z);
  fe52_add(&t, &xx, &yy);
  fe52_negate(&n_xx_yy, &t, 2);
  fe52_add(&t, &yy, &zz);
  fe52_negate(&n_yy_zz, &t, 2);
  fe52_add(&t, &xx, &zz);
  fe52_negate(&n_xx_zz, &t, 2);
  fe52_add(&t, &p->x, &p->y);
  fe52_add(&u, &q->x, &q->y);
  fe52_mul(&xy_pairs, &t, &u);
  fe52_add(&xy_pairs, &xy_pairs, &n_xx_yy);
  fe52_add(&t, &p->y, &p->z);
  fe52_add(&u, &q->y, &q->z);
  fe52_mul(&yz_pairs, &t, &u);
  fe52_add(&yz_pairs, &yz_pairs, &n_yy_zz);
  fe52_add(&t, &p->x, &p->z);
  fe52_add(&u, &q->x, &q->z);
  fe52_mul(&xz_pairs, &t, &u);
  fe52_add(&xz_pairs, &xz_pairs, &n_xx_zz);

  fe52_mul_single(&bzz, &zz, KB);
  fe52_add(&t, &bzz, &bzz);
  fe52_add(&bzz3, &t, &bzz);
  fe52_normalize_weak(&bzz3);
  fe52_negate(&t, &bzz3, 1);
  fe52_add(&yy_m_bzz3, &yy, &t);
  fe52_add(&yy_p_bzz3, &yy, &bzz3);

  fe52_mul_single(&byz, &yz_pairs, KB);
  fe52_normalize_weak(&byz);
  fe52_add(&t, &byz, &byz);
  fe52_add(&byz3, &t, &byz);
  fe52_normalize_weak(&byz3);

  fe52_add(&t, &xx, &xx);
  fe52_add(&xx3, &t, &xx);
  fe52_add(&t, &xx3, &xx3);
  fe52_add(&bxx9, &t, &xx3);
  fe52_normalize_weak(&bxx9);
  fe52_mul_single(&bxx9, &bxx9, KB);
  fe52_normalize_weak(&bxx9);

  fe52 rx, ry, rz;
  fe52_mul(&t, &xy_pairs, &yy_m_bzz3);
  fe52_mul(&u, &byz3, &xz_pairs);
  fe52_negate(&u, &u, 1);
  fe52_add(&rx, &t, &u);
  fe52_normalize_weak(&rx);
  fe52_mul(&t, &yy_p_bzz3, &yy_m_bzz3);
  fe52_mul(&u, &bxx9, &xz_pairs);
  fe52_add(&ry, &t, &u);
  fe52_normalize_weak(&ry);
  fe52_mul(&t, &yz_pairs, &yy_p_bzz3);
  fe52_mul(&u, &xx3, &xy_pairs);
  fe52_add(&rz, &t, &u);
  fe52_normalize_weak(&rz);
  r->x = rx;
  r->y = ry;
  r->z = rz;
}
/* projective.rs:142-176 add_assign_mixed (RCB algorithm 8) */
static void kpt_add_mixed(kpt* r, const kpt* p, const kaff* q) {
  if (q->inf) {
    *r = *p;
    return;
  }
  fe52 xx, yy, xy_pairs, yz_pairs, xz_pairs, t, u;
  fe52 bzz, bzz3, yy_m_bzz3, yy_p_bzz3, byz, byz3, xx3, bxx9;
  fe52_mul(&xx, &p->x, &q->x);
  fe52_mul(&yy, &p->y, &q->y);
  fe52_add(&t, &p->x, &p->y);
  fe52_add(&u, &q->x, &q->y);
  fe52_mul(&xy_pairs, &t, &u);
  fe52_add(&t, &xx, &yy);
  fe52_negate(&t, &t, 2);
  fe52_add(&xy_pairs, &xy_pairs, &t);
  fe52_mul(&yz_pairs, &q->y, &p->z);
  fe52_add(&yz_pairs, &yz_pairs, &p->y);
  fe52_mul(&xz_pairs, &q->x, &p->z);
  fe52_add(&xz_pairs, &xz_pairs, &p->x);

  fe52_mul_single(&bzz, &p->z, KB);
  fe52_add(&t, &bzz, &bzz);
  fe52_add(&bzz3, &t, &bzz);
  fe52_normalize_weak(&bzz3);
  fe52_negate(&t, &bzz3, 1);
  fe52_add(&yy_m_bzz3, &yy, &t);
  fe52_add(&yy_p_bzz3, &yy, &bzz3);

  fe52_mul_single(&byz, &yz_pairs, KB);
  fe52_normalize_weak(&byz);
  fe52_add(&t, &byz, &byz);
  fe52_add(&byz3, &t, &byz);
  fe52_normalize_weak(&byz3);

  fe52_add(&t, &xx, &xx);
  fe52_add(&xx3, &t, &xx);
  fe52_add(&t, &xx3, &xx3);
  fe52_add(&bxx9, &t, &xx3);
  fe52_normalize_weak(&bxx9);
  fe52_mul_single(&bxx9, &bxx9, KB);
  fe52_normalize_weak(&bxx9);

  fe52 rx, ry, rz;
  fe52_mul(&t, &xy_pairs, &yy_m_bzz3);
  fe52_mul(&u, &byz3, &xz_pairs);
  fe52_negate(&u, &u, 1);
  fe52_add(&rx, &t, &u);
  fe52_normalize_weak(&rx);
  fe52_mul(&t, &yy_p_bzz3, &yy_m_bzz3);
  fe52_mul(&u, &bxx9, &xz_pairs);
  fe52_add(&ry, &t, &u);
  fe52_normalize_weak(&ry);
  fe52_mul(&t, &yz_pairs, &yy_p_bzz3);
  fe52_mul(&u, &xx3, &xy_pairs);
  fe52_add(&rz, &t, &u);
  fe52_normalize_weak(&rz);
  r->x = rx;
  r->y = ry;
  r->z = rz;
}
/* projective.rs:189-217 double_in_place (RCB algorithm 9) */
static void kpt_dbl(kpt* r, const kpt* p) {
  fe52 yy, zz, xy2, bzz, bzz3, bzz9, yy_m_bzz9, yy_p_bzz3, yy_zz, yy_zz8, t, u;
  fe52_sqr(&yy, &p->y);
  fe52_sqr(&zz, &p->z);
  fe52_mul(&xy2, &p->x, &p->y);
  fe52_add(&xy2, &xy2, &xy2);
  fe52_mul_single(&bzz, &zz, KB);
  fe52_add(&t, &bzz, &bzz);
  fe52_add(&bzz3, &t, &bzz);
  fe52_normalize_weak(&bzz3);
  fe52_add(&t, &bzz3, &bzz3);
  fe52_add(&bzz9, &t, &bzz3);
  fe52_normalize_weak(&bzz9);
  fe52_negate(&t, &bzz9, 1);
  fe52_add(&yy_m_bzz9, &yy, &t);
  fe52_add(&yy_p_bzz3, &yy, &bzz3);
  fe52_mul(&yy_zz, &yy, &zz);
  fe52_mul_single(&yy_zz8, &yy_zz, 8);
  fe52_add(&t, &yy_zz8, &yy_zz8);
  fe52_add(&t, &t, &yy_zz8);
  fe52_normalize_weak(&t);
  fe52_mul_single(&t, &t, KB);
  fe52 rx, ry, rz;
  fe52_mul(&rx, &xy2, &yy_m_bzz9);
  fe52_mul(&u, &yy, &p->y);
  fe52_mul(&rz, &u, &p->z);
  fe52_mul_single(&rz, &rz, 8);
  fe52_normalize_weak(&rz);
  fe52_mul(&u, &yy_m_bzz9, &yy_p_bzz3);
  fe52_add(&ry, &u, &t);
  fe52_normalize_weak(&ry);
  r->x = rx;
  r->y = ry;
  r->z = rz;
}
/* projective.rs:241-247 endomorphism */
static void kpt_endo(kpt* r, const kpt* p) {
  fe52_mul(&r->x, &p->x, &K_BETA);
  r->y = p->y;
  r->z = p->z;
}
/* projective.rs:64-75 to_affine */
static void kpt_to_affine(kaff* r, const kpt* p) {
  if (fe52_normalizes_to_zero(&p->z)) {
    r->x = FE52_ZERO;
    r->y = FE52_ZERO;
    r->inf = 1;
    return;
  }
  fe52 zi;
  fe52_inv(&zi, &p->z);
  fe52_mul(&r->x, &p->x, &zi);
  fe52_mul(&r->y, &p->y, &zi);
  fe52_normalize(&r->x);
  fe52_normalize(&r->y);
  r->inf = 0;
}
static void kpt_from_affine(kpt* r, const kaff* a) { /* projective.rs:281-290 */
  if (a->inf) {
    kpt_identity(r);
    return;
  }
  r->x = a->x;
  r->y = a->y;
  r->z = FE52_ONE;
}

/* ================================================================================================
 * P-256 base field, Montgomery form, 4 x 64 limbs (p256/src/arithmetic/field.rs, field/field64.rs)
 * ================================================================================================ */
typedef struct {
  uint64_t w[4];
} fp256;
static const uint64_t P256_P[4] = {0xFFFFFFFFFFFFFFFFULL, 0x00000000FFFFFFFFULL, 0x0000000000000000ULL, 0xFFFFFFFF00000001ULL};
static const uint64_t P256_N[4] = {0xF3B9CAC2FC632551ULL, 0xBCE6FAADA7179E84ULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFF00000000ULL};
static const fp256 P256_R2 = {{0x0000000000000003ULL, 0xFFFFFFFBFFFFFFFFULL, 0xFFFFFFFFFFFFFFFEULL, 0x00000004FFFFFFFDULL}}; /* field.rs:181-186 */
static const fp256 P256_ONE = {{0x0000000000000001ULL, 0xFFFFFFFF00000000ULL, 0xFFFFFFFFFFFFFFFFULL, 0x00000000FFFFFFFEULL}}; /* R mod p */
static const fp256 P256_ZERO = {{0, 0, 0, 0}};

/* field64.rs:127-144 sub_inner: 5-limb subtract, add the modulus back on borrow */
static void fp_sub_inner(uint64_t* r, const uint64_t* l, const uint64_t* rr) {
  uint64_t w[5];
  uint64_t borrow = limbs_sub(w, l, rr, 5);
  uint64_t mask = 0 - borrow;
  uint64_t m[4] = {P256_P[0] & mask, P256_P[1] & mask, P256_P[2] & mask, P256_P[3] & mask};
  limbs_add(r, w, m, 4);
}
/* field64.rs:7-23 add */
static void fp_add(fp256* r, const fp256* a, const fp256* b) {
  uint64_t s[5];
  s[4] = limbs_add(s, a->w, b->w, 4);
  uint64_t m[5] = {P256_P[0], P256_P[1], P256_P[2], P256_P[3], 0};
  fp_sub_inner(r->w, s, m);
}
/* field64.rs:25-34 sub */
static void fp_sub(fp256* r, const fp256* a, const fp256* b) {
  uint64_t l[5] = {a->w[0], a->w[1], a->w[2], a->w[3], 0};
  uint64_t m[5] = {b->w[0], b->w[1], b->w[2], b->w[3], 0};
  fp_sub_inner(r->w, l, m);
}
/* a*b + c + d -> (lo, hi)   (crypto-bigint Limb::carrying_mul_add) */
static inline uint64_t mac(uint64_t a, uint64_t b, uint64_t c, uint64_t d, uint64_t* hi) {
  u128 t = (u128)a * b + c + d;
  *hi = (uint64_t)(t >> 64);
  return (uint64_t)t;
}
static inline uint64_t adc(uint64_t a, uint64_t b, uint64_t c, uint64_t* hi) {
  u128 t = (u128)a + b + c;
  *hi = (uint64_t)(t >> 64);
  return (uint64_t)t;
}
/* field64.rs:83-123 montgomery_reduce: word-by-word, p' = 1, p[0] = 2^64-1 (carry initialised to the
 * limb itself), p[2] = 0 (carry only) */
static void fp_mont_reduce(fp256* r, const uint64_t* a) {
  uint64_t a0 = a[0], a1 = a[1], a2 = a[2], a3 = a[3], a4 = a[4], a5 = a[5], a6 = a[6], a7 = a[7], a8;
  uint64_t carry, carry2;
  a1 = mac(a0, P256_P[1], a1, a0, &carry);
  a2 = adc(a2, 0, carry, &carry);
  a3 = mac(a0, P256_P[3], a3, carry, &carry);
  a4 = adc(a4, 0, carry, &carry2);

  a2 = mac(a1, P256_P[1], a2, a1, &carry);
  a3 = adc(a3, 0, carry, &carry);
  a4 = mac(a1, P256_P[3], a4, carry, &carry);
  a5 = adc(a5, carry2, carry, &carry2);

  a3 = mac(a2, P256_P[1], a3, a2, &carry);
  a4 = adc(a4, 0, carry, &carry);
  a5 = mac(a2, P256_P[3], a5, carry, &carry);
  a6 = adc(a6, carry2, carry, &carry2);

  a4 = mac(a3, P256_P[1], a4, a3, &carry);
  a5 = adc(a5, 0, carry, &carry);
  a6 = mac(a3, P256_P[3], a6, carry, &carry);
  a7 = adc(a7, carry2, carry, &a8);

  uint64_t l[5] = {a4, a5, a6, a7, a8};
  uint64_t m[5] = {P256_P[0], P256_P[1], P256_P[2], P256_P[3], 0};
  fp_sub_inner(r->w, l, m);
}
/* field.rs:99-102 multiply = U256::widening_mul + montgomery_reduce */
static void fp_mul(fp256* r, const fp256* a, const fp256* b) {
  uint64_t l[8];
  sc_mul_wide(l, a->w, b->w);
  fp_mont_reduce(r, l);
}
static void fp_sqr(fp256* r, const fp256* a) { fp_mul(r, a, a); } /* field.rs:105-108 */
static void fp_dbl(fp256* r, const fp256* a) { fp_add(r, a, a); }  /* field.rs:76-78 */
/* field.rs:59-64 from_uint_unchecked: a * R2 */
static void fp_from_be(fp256* r, const uint8_t* b) {
  fp256 t;
  sc_from_be(t.w, b);
  fp_mul(r, &t, &P256_R2);
}
/* field.rs:94-96 to_canonical = montgomery_reduce(a, 0) */
static void fp_to_be(uint8_t* b, const fp256* a) {
  uint64_t l[8] = {a->w[0], a->w[1], a->w[2], a->w[3], 0, 0, 0, 0};
  fp256 c;
  fp_mont_reduce(&c, l);
  sc_to_be(b, c.w);
}
static int fp_is_zero(const fp256* a) { return (a->w[0] | a->w[1] | a->w[2] | a->w[3]) == 0; }
static void fp_sqrn(fp256* r, const fp256* a, int n) {
  *r = *a;
  for (int i = 0; i < n; i++) fp_sqr(r, r);
}
/* field.rs:111-113 invert -> primefield/src/monty.rs:373-375 -> crypto-bigint (third-party): a^(p-2) */
static void fp_inv(fp256* r, const fp256* a) {
  fp256 x2, x3, x6, x12, x15, x30, x32, t;
  fp_sqr(&x2, a);
  fp_mul(&x2, &x2, a);
  fp_sqr(&x3, &x2);
  fp_mul(&x3, &x3, a);
  fp_sqrn(&x6, &x3, 3);
  fp_mul(&x6, &x6, &x3);
  fp_sqrn(&x12, &x6, 6);
  fp_mul(&x12, &x12, &x6);
  fp_sqrn(&x15, &x12, 3);
  fp_mul(&x15, &x15, &x3);
  fp_sqrn(&x30, &x15, 15);
  fp_mul(&x30, &x30, &x15);
  fp_sqrn(&x32, &x30, 2);
  fp_mul(&x32, &x32, &x2);
  /* p-2 = [32 ones][31 zeros][1][96 zeros][94 ones][0][1] */
  fp_sqrn(&t, &x32, 32);
  fp_mul(&t, &t, a);
  fp_sqrn(&t, &t, 128);
  fp_mul(&t, &t, &x32);
  fp_sqrn(&t, &t, 32);
  fp_mul(&t, &t, &x32);
  fp_sqrn(&t, &t, 30);
  fp_mul(&t, &t, &x30);
  fp_sqrn(&t, &t, 2);
  fp_mul(r, &t, a);
}

/* ================================================================================================
 * P-256 points: primeorder::ProjectivePoint<NistP256> with EquationAIsMinusThree
 * (primeorder/src/point_arithmetic.rs:212-319, p256/src/arithmetic.rs:43-75)
 * ================================================================================================ */
typedef struct {
  fp256 x, y, z;
} ppt;
typedef struct {
  fp256 x, y;
  int inf;
} paff;
static fp256 P256_B; /* EQUATION_B in Montgomery form, set by ecref_init */
static fp256 P256_GX, P256_GY;

static void ppt_identity(ppt* r) {
  r->x = P256_ZERO;
  r->y = P256_ONE;
  r->z = P256_ZERO;
}
/* point_arithmetic.rs:222-245 add_assign (RCB algorithm 4) */
static void ppt_add(ppt* r, const ppt* p, const ppt* q) {
  fp256 xx, yy, zz, xy_pairs, yz_pairs, xz_pairs, t, u;
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

  fp256 bzz_part, bzz3_part, yy_m_bzz3, yy_p_bzz3, zz3, bxz_part, bxz3_part, xx3_m_zz3;
  fp_mul(&t, &P256_B, &zz);
  fp_sub(&bzz_part, &xz_pairs, &t);
  fp_dbl(&t, &bzz_part);
  fp_add(&bzz3_part, &t, &bzz_part);
  fp_sub(&yy_m_bzz3, &yy, &bzz3_part);
  fp_add(&yy_p_bzz3, &yy, &bzz3_part);
  fp_dbl(&t, &zz);
  fp_add(&zz3, &t, &zz);
  fp_mul(&t, &P256_B, &xz_pairs);
  fp_add(&u, &zz3, &xx);
  fp_sub(&bxz_part, &t, &u);
  fp_dbl(&t, &bxz_part);
  fp_add(&bxz3_part, &t, &bxz_part);
  fp_dbl(&t, &xx);
  fp_add(&t, &t, &xx);
  fp_sub(&xx3_m_zz3, &t, &zz3);

  fp256 rx, ry, rz;
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
/* point_arithmetic.rs:289-318 double_in_place (RCB algorithm 6) */
static void ppt_dbl(ppt* r, const ppt* p) {
  fp256 xx, yy, zz, xy2, xz2, t, u;
  fp_sqr(&xx, &p->x);
  fp_sqr(&yy, &p->y);
  fp_sqr(&zz, &p->z);
  fp_mul(&xy2, &p->x, &p->y);
  fp_dbl(&xy2, &xy2);
  fp_mul(&xz2, &p->x, &p->z);
  fp_dbl(&xz2, &xz2);
  fp256 bzz_part, bzz3_part, yy_m_bzz3, yy_p_bzz3, y_frag, x_frag, zz3, bxz2_part, bxz6_part, xx3_m_zz3;
  fp_mul(&t, &P256_B, &zz);
  fp_sub(&bzz_part, &t, &xz2);
  fp_dbl(&t, &bzz_part);
  fp_add(&bzz3_part, &t, &bzz_part);
  fp_sub(&yy_m_bzz3, &yy, &bzz3_part);
  fp_add(&yy_p_bzz3, &yy, &bzz3_part);
  fp_mul(&y_frag, &yy_p_bzz3, &yy_m_bzz3);
  fp_mul(&x_frag, &yy_m_bzz3, &xy2);
  fp_dbl(&t, &zz);
  fp_add(&zz3, &t, &zz);
  fp_mul(&t, &P256_B, &xz2);
  fp_add(&u, &zz3, &xx);
  fp_sub(&bxz2_part, &t, &u);
  fp_dbl(&t, &bxz2_part);
  fp_add(&bxz6_part, &t, &bxz2_part);
  fp_dbl(&t, &xx);
  fp_add(&t, &t, &xx);
  fp_sub(&xx3_m_zz3, &t, &zz3);
  fp256 rx, ry, rz, yz2;
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
static void ppt_neg(ppt* r, const ppt* p) { /* primeorder/src/projective.rs Neg: (x, -y, z) */
  r->x = p->x;
  fp_sub(&r->y, &P256_ZERO, &p->y);
  r->z = p->z;
}
/* primeorder/src/projective.rs:101-113 to_affine */
static void ppt_to_affine(paff* r, const ppt* p) {
  if (fp_is_zero(&p->z)) {
    r->x = P256_ZERO;
    r->y = P256_ZERO;
    r->inf = 1;
    return;
  }
  fp256 zi;
  fp_inv(&zi, &p->z);
  fp_mul(&r->x, &p->x, &zi);
  fp_mul(&r->y, &p->y, &zi);
  r->inf = 0;
}
static void ppt_from_affine(ppt* r, const paff* a) {
  if (a->inf) {
    ppt_identity(r);
    return;
  }
  r->x = a->x;
  r->y = a->y;
  r->z = P256_ONE;
}

/* ================================================================================================
 * Generic drivers, instantiated twice by macro: radix-16 lookup tables, constant-time lincomb,
 * wNAF recoding + Straus, basepoint table.
 * ================================================================================================ */

/* primeorder/src/tables/radix16.rs:31-55 Radix16Decomposition::new — nd digits from 32 BE bytes */
static void radix16(int8_t* d, int nd, const uint8_t* be32) {
  memset(d, 0, (size_t)nd);
  for (int i = 0; i < (nd - 1) / 2; i++) {
    uint8_t b = be32[31 - i];
    d[2 * i] = (int8_t)(b & 0xf);
    d[2 * i + 1] = (int8_t)((b >> 4) & 0xf);
  }
  for (int i = 0; i < nd - 1; i++) {
    int8_t carry = (int8_t)((d[i] + 8) >> 4);
    d[i] = (int8_t)(d[i] - (carry << 4));
    d[i + 1] = (int8_t)(d[i + 1] + carry);
  }
}

/* wnaf/src/limb_buffer.rs:5-68 LimbBuffer + wnaf/src/lib.rs:70-150 wnaf_form.
 * c: little-endian bytes (len bytes), returns number of digits written. */
static uint64_t le_limb(const uint8_t* c, size_t len, size_t idx) {
  uint64_t v = 0;
  for (int j = 0; j < 8; j++) {
    size_t p = idx * 8 + (size_t)j;
    if (p < len) v |= (uint64_t)c[p] << (8 * j);
  }
  return v;
}
static int wnaf_form(int8_t* wnaf, const uint8_t* c, size_t len, size_t bit_len, int window) {
  uint64_t width = 1ULL << window, mask = width - 1;
  size_t pos = 0;
  uint64_t carry = 0;
  int cursor = 0;
  while (pos < bit_len) {
    size_t u64_idx = pos / 64, bit_idx = pos % 64;
    uint64_t cur = le_limb(c, len, u64_idx), next = le_limb(c, len, u64_idx + 1);
    uint64_t bit_buf = (bit_idx + (size_t)window < 64) ? (cur >> bit_idx) : ((cur >> bit_idx) | (next << (64 - bit_idx)));
    uint64_t window_val = carry + (bit_buf & mask);
    if ((window_val & 1) == 0) {
      wnaf[cursor++] = 0;
      pos += 1;
    } else {
      int8_t dg = (int8_t)window_val;
      if (window_val < width / 2) {
        carry = 0;
      } else {
        carry = 1;
        dg = (int8_t)((int64_t)window_val - (int64_t)width);
      }
      wnaf[cursor++] = dg;
      size_t max_pos = bit_len >= carry ? bit_len - (size_t)carry : 0;
      size_t skip = (size_t)window < (max_pos - pos) ? (size_t)window : (max_pos - pos);
      for (size_t s = 1; s < skip; s++) wnaf[cursor++] = 0;
      pos += skip;
    }
  }
  if (carry) wnaf[cursor++] = (int8_t)carry;
  return cursor;
}

#define WNAF_W 5        /* k256/src/arithmetic/mul.rs:58, primeorder/src/projective.rs:39 */
#define WNAF_TAB 8      /* 2^(w-2) odd multiples */

#define DEFINE_DRIVERS(PFX, PT, AFF, ADD, DBL, NEG, IDENT)                                                   \
  /* primeorder/src/tables/lookup.rs:30-38 LookupTable::new */                                               \
  static void PFX##_lut_new(PT* tab, const PT* p) {                                                          \
    tab[0] = *p;                                                                                             \
    for (int j = 0; j < 7; j++) ADD(&tab[j + 1], p, &tab[j]);                                                \
  }                                                                                                          \
  /* lookup.rs:43-65 select: |x| by scan + conditional negate */                                             \
  static void PFX##_lut_select(PT* r, const PT* tab, int8_t x) {                                             \
    int8_t xmask = (int8_t)(x >> 7);                                                                         \
    int8_t xabs = (int8_t)((x + xmask) ^ xmask);                                                             \
    PT t;                                                                                                    \
    IDENT(&t);                                                                                               \
    for (int j = 1; j <= 8; j++)                                                                             \
      if (xabs == j) t = tab[j - 1];                                                                         \
    if (xmask & 1) {                                                                                         \
      PT n;                                                                                                  \
      NEG(&n, &t);                                                                                           \
      t = n;                                                                                                 \
    }                                                                                                        \
    *r = t;                                                                                                  \
  }                                                                                                          \
  /* wnaf/src/lib.rs:55-65 wnaf_table */                                                                     \
  static void PFX##_wnaf_table(PT* tab, const PT* base) {                                                    \
    PT dbl, cur = *base;                                                                                     \
    DBL(&dbl, base);                                                                                         \
    for (int i = 0; i < WNAF_TAB; i++) {                                                                     \
      tab[i] = cur;                                                                                          \
      ADD(&cur, &cur, &dbl);                                                                                 \
    }                                                                                                        \
  }                                                                                                          \
  /* wnaf/src/lib.rs:157-194 wnaf_multi_exp (Straus, MSB first) */                                           \
  static void PFX##_wnaf_multi_exp(PT* out, const PT* tabs /*nt*WNAF_TAB*/, const int8_t* digits /*nt*stride*/, \
                                   const int* lens, int nt, int stride) {                                    \
    int maxlen = 0;                                                                                          \
    for (int t = 0; t < nt; t++)                                                                             \
      if (lens[t] > maxlen) maxlen = lens[t];                                                                \
    PT result;                                                                                               \
    IDENT(&result);                                                                                          \
    int found_one = 0;                                                                                       \
    for (int i = maxlen - 1; i >= 0; i--) {                                                                  \
      if (found_one) DBL(&result, &result);                                                                  \
      for (int t = 0; t < nt; t++) {                                                                         \
        int n = i < lens[t] ? digits[t * stride + i] : 0;                                                    \
        if (n != 0) {                                                                                        \
          found_one = 1;                                                                                     \
          if (n > 0) {                                                                                       \
            ADD(&result, &result, &tabs[t * WNAF_TAB + n / 2]);                                              \
          } else {                                                                                           \
            PT ng;                                                                                           \
            NEG(&ng, &tabs[t * WNAF_TAB + (-n) / 2]);                                                        \
            ADD(&result, &result, &ng);                                                                      \
          }                                                                                                  \
        }                                                                                                    \
      }                                                                                                      \
    }                                                                                                        \
    *out = result;                                                                                           \
  }                                                                                                          \
  /* primeorder/src/tables/basepoint.rs:41-75 init_table: 33 LookupTables of 2^(8i) G */                     \
  static PT PFX##_basetab[33][8];                                                                            \
  static void PFX##_basetab_init(const PT* g) {                                                              \
    PT gen = *g;                                                                                             \
    for (int i = 0; i < 33; i++) {                                                                           \
      PFX##_lut_new(PFX##_basetab[i], &gen);                                                                 \
      for (int j = 0; j < 8; j++) DBL(&gen, &gen);                                                           \
    }                                                                                                        \
  }                                                                                                          \
  /* basepoint.rs:82-98 BasepointTable::mul  ==  k256/src/arithmetic/mul.rs:180-203 mul_by_generator */      \
  static void PFX##_mul_gen(PT* out, const uint8_t* k_be) {                                                  \
    int8_t d[65];                                                                                            \
    radix16(d, 65, k_be);                                                                                    \
    PT acc, acc2, t;                                                                                         \
    PFX##_lut_select(&acc, PFX##_basetab[32], d[64]);                                                        \
    IDENT(&acc2);                                                                                            \
    for (int i = 31; i >= 0; i--) {                                                                          \
      PFX##_lut_select(&t, PFX##_basetab[i], d[2 * i + 1]);                                                  \
      ADD(&acc2, &acc2, &t);                                                                                 \
      PFX##_lut_select(&t, PFX##_basetab[i], d[2 * i]);                                                      \
      ADD(&acc, &acc, &t);                                                                                   \
    }                                                                                                        \
    for (int j = 0; j < 4; j++) DBL(&acc2, &acc2);                                                           \
    ADD(out, &acc, &acc2);                                                                                   \
  }

DEFINE_DRIVERS(k, kpt, kaff, kpt_add, kpt_dbl, kpt_neg, kpt_identity)
DEFINE_DRIVERS(p, ppt, paff, ppt_add, ppt_dbl, ppt_neg, ppt_identity)

/* ---- k256 drivers -------------------------------------------------------------------------------- */

/* k256/src/arithmetic/mul.rs:112-163 lincomb (constant-time): GLV split, sign folding, two radix-16
 * tables per term, shared doublings.  nterms terms; scratch allocated by the caller. */
typedef struct {
  kpt t1[8], t2[8];
  int8_t d1[33], d2[33];
} kterm;
static void k_prepare_term(kterm* T, const kpt* x, const uint8_t* k_be) {
  uint64_t k[4], r1[4], r2[4];
  sc_from_be(k, k_be);
  glv_decompose(r1, r2, k);
  kpt xb, t;
  kpt_endo(&xb, x);
  int s1 = sc_is_high(r1), s2 = sc_is_high(r2);
  if (s1) sc_neg(r1, r1);
  if (s2) sc_neg(r2, r2);
  if (s1) {
    kpt_neg(&t, x);
    k_lut_new(T->t1, &t);
  } else {
    k_lut_new(T->t1, x);
  }
  if (s2) {
    kpt_neg(&t, &xb);
    k_lut_new(T->t2, &t);
  } else {
    k_lut_new(T->t2, &xb);
  }
  uint8_t b[32];
  sc_to_be(b, r1);
  radix16(T->d1, 33, b);
  sc_to_be(b, r2);
  radix16(T->d2, 33, b);
}
static void k_lincomb_ct(kpt* out, const kterm* T, size_t nterms) {
  kpt acc, t;
  kpt_identity(&acc);
  for (size_t c = 0; c < nterms; c++) {
    k_lut_select(&t, T[c].t1, T[c].d1[32]);
    kpt_add(&acc, &acc, &t);
    k_lut_select(&t, T[c].t2, T[c].d2[32]);
    kpt_add(&acc, &acc, &t);
  }
  for (int i = 31; i >= 0; i--) {
    for (int j = 0; j < 4; j++) kpt_dbl(&acc, &acc);
    for (size_t c = 0; c < nterms; c++) {
      k_lut_select(&t, T[c].t1, T[c].d1[i]);
      kpt_add(&acc, &acc, &t);
      k_lut_select(&t, T[c].t2, T[c].d2[i]);
      kpt_add(&acc, &acc, &t);
    }
  }
  *out = acc;
}
/* mul.rs:236-238 mul = lincomb(&[(x,k)]) */
static void k_mul_ct(kpt* out, const kpt* x, const uint8_t* k_be) {
  kterm T;
  k_prepare_term(&T, x, k_be);
  k_lincomb_ct(out, &T, 1);
}
/* glv.rs:170-189 decompose_wnaf_into + mul.rs:242-247 mul_vartime / :167-175 lincomb_vartime_glv_wnaf */
typedef struct {
  kpt tab[2][WNAF_TAB];
  int8_t dg[2][130];
  int len[2];
} kwterm;
static void k_prepare_wterm(kwterm* T, const kpt* x, const uint8_t* k_be) {
  uint64_t k[4], r1[4], r2[4];
  sc_from_be(k, k_be);
  glv_decompose(r1, r2, k);
  int s1 = sc_is_high(r1), s2 = sc_is_high(r2);
  if (s1) sc_neg(r1, r1);
  if (s2) sc_neg(r2, r2);
  uint8_t le[16];
  for (int i = 0; i < 16; i++) le[i] = (uint8_t)(r1[i / 8] >> (8 * (i % 8)));
  T->len[0] = wnaf_form(T->dg[0], le, 16, 128, WNAF_W);
  for (int i = 0; i < 16; i++) le[i] = (uint8_t)(r2[i / 8] >> (8 * (i % 8)));
  T->len[1] = wnaf_form(T->dg[1], le, 16, 128, WNAF_W);
  kpt p1, pb, p2;
  if (s1)
    kpt_neg(&p1, x);
  else
    p1 = *x;
  kpt_endo(&pb, x);
  if (s2)
    kpt_neg(&p2, &pb);
  else
    p2 = pb;
  k_wnaf_table(T->tab[0], &p1);
  k_wnaf_table(T->tab[1], &p2);
}
static void k_mul_vartime(kpt* out, const kpt* x, const uint8_t* k_be) {
  kwterm T;
  k_prepare_wterm(&T, x, k_be);
  k_wnaf_multi_exp(out, &T.tab[0][0], &T.dg[0][0], T.len, 2, 130);
}

/* mul.rs:303-310 mul_by_generator_and_mul_add_vartime: a*G + b*P = lincomb_vartime_glv_wnaf of the two decomposed terms */
static kpt K_GEN;
static void k_mul_gen_add(kpt* out, const uint8_t* a_be, const uint8_t* b_be, const kpt* P) {
  kwterm T[2];
  k_prepare_wterm(&T[0], &K_GEN, a_be);
  k_prepare_wterm(&T[1], P, b_be);
  kpt tabs[4 * WNAF_TAB];
  int8_t dg[4 * 130];
  int lens[4];
  for (int t = 0; t < 2; t++)
    for (int h = 0; h < 2; h++) {
      memcpy(&tabs[(2 * t + h) * WNAF_TAB], T[t].tab[h], sizeof(kpt) * WNAF_TAB);
      memcpy(&dg[(2 * t + h) * 130], T[t].dg[h], 130);
      lens[2 * t + h] = T[t].len[h];
    }
  k_wnaf_multi_exp(out, tabs, dg, lens, 4, 130);
}

/* ---- p256 drivers (primeorder/src/projective.rs:133-144, :532-557) -------------------------------- */
typedef struct {
  ppt t[8];
  int8_t d[65];
} pterm;
static void p_prepare_term(pterm* T, const ppt* x, const uint8_t* k_be) {
  p_lut_new(T->t, x);
  radix16(T->d, 65, k_be);
}
static void p_lincomb_ct(ppt* out, const pterm* T, size_t nterms) {
  ppt q, t;
  ppt_identity(&q);
  for (size_t c = 0; c < nterms; c++) {
    p_lut_select(&t, T[c].t, T[c].d[64]);
    ppt_add(&q, &q, &t);
  }
  for (int i = 63; i >= 0; i--) {
    for (int j = 0; j < 4; j++) ppt_dbl(&q, &q);
    for (size_t c = 0; c < nterms; c++) {
      p_lut_select(&t, T[c].t, T[c].d[i]);
      ppt_add(&q, &q, &t);
    }
  }
  *out = q;
}
static void p_mul_ct(ppt* out, const ppt* x, const uint8_t* k_be) {
  pterm T;
  p_prepare_term(&T, x, k_be);
  p_lincomb_ct(out, &T, 1);
}
/* projective.rs:142-144 mul_vartime = WnafBase::new(self) * WnafScalar::new(k)  (257-digit wNAF-5) */
static void p_mul_vartime(ppt* out, const ppt* x, const uint8_t* k_be) {
  ppt tab[WNAF_TAB];
  int8_t dg[260];
  uint8_t le[32];
  for (int i = 0; i < 32; i++) le[i] = k_be[31 - i];
  int len = wnaf_form(dg, le, 32, 256, WNAF_W);
  p_wnaf_table(tab, x);
  p_wnaf_multi_exp(out, tab, dg, &len, 1, 260);
}

/* primeorder/src/mul_backend.rs:31-40 (default): lincomb_vartime(&[(G, a), (P, b)]) = 2-term wNAF-5 Straus
 * (primeorder/src/projective.rs:525-529) */
static ppt P_GEN;
static void p_mul_gen_add(ppt* out, const uint8_t* a_be, const uint8_t* b_be, const ppt* P) {
  ppt tabs[2 * WNAF_TAB];
  int8_t dg[2 * 260];
  int lens[2];
  uint8_t le[32];
  for (int i = 0; i < 32; i++) le[i] = a_be[31 - i];
  lens[0] = wnaf_form(&dg[0], le, 32, 256, WNAF_W);
  for (int i = 0; i < 32; i++) le[i] = b_be[31 - i];
  lens[1] = wnaf_form(&dg[260], le, 32, 256, WNAF_W);
  p_wnaf_table(&tabs[0], &P_GEN);
  p_wnaf_table(&tabs[WNAF_TAB], P);
  p_wnaf_multi_exp(out, tabs, dg, lens, 2, 260);
}

/* ================================================================================================
 * Exported batch API (ctypes).  curve: 0 = k256, 1 = p256.  Bytes as in include/ecgpu.h.
 * variant: 0 = constant-time `*` path (mul.rs:236-238 / projective.rs:133-137), 1 = mul_vartime.
 * ================================================================================================ */
static int g_init = 0;
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;

static void hex32(uint8_t* out, const char* h) {
  for (int i = 0; i < 32; i++) {
    unsigned v = 0;
    for (int j = 0; j < 2; j++) {
      char c = h[2 * i + j];
      v = v * 16 + (unsigned)(c <= '9' ? c - '0' : (c | 32) - 'a' + 10);
    }
    out[i] = (uint8_t)v;
  }
}
void ecref_init(void) {
  pthread_mutex_lock(&g_lock);
  if (!g_init) {
    uint8_t b[32];
    hex32(b, "5ac635d8aa3a93e7b3ebbd55769886bc651d06b0cc53b0f63bce3c3e27d2604b"); /* p256/src/arithmetic.rs:55-57 */
    fp_from_be(&P256_B, b);
    hex32(b, "6b17d1f2e12c4247f8bce6e563a440f277037d812deb33a0f4a13945d898c296"); /* :67-74 */
    fp_from_be(&P256_GX, b);
    hex32(b, "4fe342e2fe1a7f9b8ee7eb4a7c0f9e162bce33576b315ececbb6406837bf51f5");
    fp_from_be(&P256_GY, b);
    ppt pg = {P256_GX, P256_GY, P256_ONE};
    P_GEN = pg;
    p_basetab_init(&pg);
    kpt kg;
    hex32(b, "79be667ef9dcbbac55a06295ce870b07029bfcdb2dce28d959f2815b16f81798"); /* k256/src/arithmetic/affine.rs:61-77 */
    fe52_from_be(&kg.x, b);
    hex32(b, "483ada7726a3c4655da4fbfc0e1108a8fd17b448a68554199c47d08ffb10d4b8");
    fe52_from_be(&kg.y, b);
    kg.z = FE52_ONE;
    K_GEN = kg;
    k_basetab_init(&kg);
    g_init = 1;
  }
  pthread_mutex_unlock(&g_lock);
}

static int scalar_in_range(int curve, const uint8_t* k_be) {
  uint64_t k[4];
  sc_from_be(k, k_be);
  return !limbs_ge(k, curve == 0 ? K256_N : P256_N, 4);
}
/* AffinePoint::from_coordinates on-curve check (k256/src/arithmetic/affine.rs:134-147) */
static int k_load_point(kpt* out, const uint8_t* xy, int inf) {
  if (inf) {
    kpt_identity(out);
    return 1;
  }
  fe52 x, y, l, r, seven = {{7, 0, 0, 0, 0}};
  fe52_from_be(&x, xy);
  fe52_from_be(&y, xy + 32);
  if (fe52_overflow(&x) || fe52_overflow(&y)) return 0;
  fe52_sqr(&l, &y);
  fe52_sqr(&r, &x);
  fe52_mul(&r, &r, &x);
  fe52_add(&r, &r, &seven);
  fe52_negate(&r, &r, 1);
  fe52_add(&l, &l, &r);
  if (!fe52_normalizes_to_zero(&l)) return 0;
  out->x = x;
  out->y = y;
  out->z = FE52_ONE;
  return 1;
}
static int p_load_point(ppt* out, const uint8_t* xy, int inf) {
  if (inf) {
    ppt_identity(out);
    return 1;
  }
  uint64_t w[4];
  sc_from_be(w, xy);
  if (limbs_ge(w, P256_P, 4)) return 0;
  sc_from_be(w, xy + 32);
  if (limbs_ge(w, P256_P, 4)) return 0;
  fp256 x, y, l, r, t;
  fp_from_be(&x, xy);
  fp_from_be(&y, xy + 32);
  fp_sqr(&l, &y);
  fp_sqr(&r, &x);
  fp_mul(&r, &r, &x);
  fp_dbl(&t, &x);
  fp_add(&t, &t, &x);
  fp_sub(&r, &r, &t);
  fp_add(&r, &r, &P256_B);
  fp_sub(&l, &l, &r);
  if (!fp_is_zero(&l)) return 0;
  out->x = x;
  out->y = y;
  out->z = P256_ONE;
  return 1;
}
static void k_store_point(uint8_t* xy, uint8_t* inf, const kpt* p) {
  kaff a;
  kpt_to_affine(&a, p);
  if (a.inf) {
    memset(xy, 0, 64);
    *inf = 1;
    return;
  }
  fe52_to_be(xy, &a.x);
  fe52_to_be(xy + 32, &a.y);
  *inf = 0;
}
static void p_store_point(uint8_t* xy, uint8_t* inf, const ppt* p) {
  paff a;
  ppt_to_affine(&a, p);
  if (a.inf) {
    memset(xy, 0, 64);
    *inf = 1;
    return;
  }
  fp_to_be(xy, &a.x);
  fp_to_be(xy + 32, &a.y);
  *inf = 0;
}

typedef struct {
  int curve, variant, op;
  size_t lo, hi;
  const uint8_t *k, *pxy, *pinf, *a;
  uint8_t *oxy, *oinf;
  int err; /* 0 ok, 2 scalar range, 3 not on curve (ecg_status values) */
  size_t err_index;
  /* lincomb partial */
  kpt kacc;
  ppt pacc;
} job;

enum { OP_MUL = 0, OP_MULGEN = 1, OP_LINCOMB = 2, OP_MULGENADD = 3 };
#define LINCOMB_CHUNK 256 /* terms per reference-style lincomb call (the reference's single call needs ~2 KiB/term) */

static void* worker(void* arg) {
  job* j = (job*)arg;
  j->err = 0;
  if (j->op == OP_LINCOMB) {
    kpt_identity(&j->kacc);
    ppt_identity(&j->pacc);
    kterm* KT = j->curve == 0 ? (kterm*)malloc(sizeof(kterm) * LINCOMB_CHUNK) : NULL;
    pterm* PT_ = j->curve == 1 ? (pterm*)malloc(sizeof(pterm) * LINCOMB_CHUNK) : NULL;
    for (size_t base = j->lo; base < j->hi; base += LINCOMB_CHUNK) {
      size_t cnt = j->hi - base < LINCOMB_CHUNK ? j->hi - base : LINCOMB_CHUNK;
      for (size_t c = 0; c < cnt; c++) {
        size_t i = base + c;
        int inf = j->pinf ? j->pinf[i] : 0;
        if (!scalar_in_range(j->curve, j->k + 32 * i)) {
          j->err = 2;
          j->err_index = i;
          goto done;
        }
        if (j->curve == 0) {
          kpt P;
          if (!k_load_point(&P, j->pxy + 64 * i, inf)) {
            j->err = 3;
            j->err_index = i;
            goto done;
          }
          k_prepare_term(&KT[c], &P, j->k + 32 * i);
        } else {
          ppt P;
          if (!p_load_point(&P, j->pxy + 64 * i, inf)) {
            j->err = 3;
            j->err_index = i;
            goto done;
          }
          p_prepare_term(&PT_[c], &P, j->k + 32 * i);
        }
      }
      if (j->curve == 0) {
        kpt part;
        k_lincomb_ct(&part, KT, cnt);
        kpt_add(&j->kacc, &j->kacc, &part);
      } else {
        ppt part;
        p_lincomb_ct(&part, PT_, cnt);
        ppt_add(&j->pacc, &j->pacc, &part);
      }
    }
  done:
    free(KT);
    free(PT_);
    return NULL;
  }
  for (size_t i = j->lo; i < j->hi; i++) {
    if (!scalar_in_range(j->curve, j->k + 32 * i)) {
      j->err = 2;
      j->err_index = i;
      return NULL;
    }
    if (j->curve == 0) {
      kpt P, R;
      if (j->op == OP_MULGEN) {
        k_mul_gen(&R, j->k + 32 * i);
      } else {
        if (!k_load_point(&P, j->pxy + 64 * i, j->pinf ? j->pinf[i] : 0)) {
          j->err = 3;
          j->err_index = i;
          return NULL;
        }
        if (j->op == OP_MULGENADD) {
          if (!scalar_in_range(j->curve, j->a + 32 * i)) {
            j->err = 2;
            j->err_index = i;
            return NULL;
          }
          k_mul_gen_add(&R, j->a + 32 * i, j->k + 32 * i, &P);
        } else if (j->variant == 0)
          k_mul_ct(&R, &P, j->k + 32 * i);
        else
          k_mul_vartime(&R, &P, j->k + 32 * i);
      }
      k_store_point(j->oxy + 64 * i, j->oinf + i, &R);
    } else {
      ppt P, R;
      if (j->op == OP_MULGEN) {
        p_mul_gen(&R, j->k + 32 * i);
      } else {
        if (!p_load_point(&P, j->pxy + 64 * i, j->pinf ? j->pinf[i] : 0)) {
          j->err = 3;
          j->err_index = i;
          return NULL;
        }
        if (j->op == OP_MULGENADD) {
          if (!scalar_in_range(j->curve, j->a + 32 * i)) {
            j->err = 2;
            j->err_index = i;
            return NULL;
          }
          p_mul_gen_add(&R, j->a + 32 * i, j->k + 32 * i, &P);
        } else if (j->variant == 0)
          p_mul_ct(&R, &P, j->k + 32 * i);
        else
          p_mul_vartime(&R, &P, j->k + 32 * i);
      }
      p_store_point(j->oxy + 64 * i, j->oinf + i, &R);
    }
  }
  return NULL;
}

static int run_jobs(job* tmpl, size_t n, int nthreads, job** out_jobs) {
  ecref_init();
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

int ecref_mul_batch(int curve, size_t n, const uint8_t* k, const uint8_t* pxy, const uint8_t* pinf, uint8_t* oxy,
                    uint8_t* oinf, int nthreads, int variant) {
  if (n == 0) return 0;
  job t;
  memset(&t, 0, sizeof t);
  t.curve = curve;
  t.variant = variant;
  t.op = OP_MUL;
  t.k = k;
  t.pxy = pxy;
  t.pinf = pinf;
  t.oxy = oxy;
  t.oinf = oinf;
  return run_jobs(&t, n, nthreads, NULL);
}
/* out[i] = a[i]*G + b[i]*P[i]  (mul_by_generator_and_mul_add_vartime) */
int ecref_mul_gen_add_batch(int curve, size_t n, const uint8_t* a, const uint8_t* b, const uint8_t* pxy, const uint8_t* pinf,
                            uint8_t* oxy, uint8_t* oinf, int nthreads) {
  if (n == 0) return 0;
  job t;
  memset(&t, 0, sizeof t);
  t.curve = curve;
  t.op = OP_MULGENADD;
  t.a = a;
  t.k = b;
  t.pxy = pxy;
  t.pinf = pinf;
  t.oxy = oxy;
  t.oinf = oinf;
  return run_jobs(&t, n, nthreads, NULL);
}
int ecref_mul_gen_batch(int curve, size_t n, const uint8_t* k, uint8_t* oxy, uint8_t* oinf, int nthreads) {
  if (n == 0) return 0;
  job t;
  memset(&t, 0, sizeof t);
  t.curve = curve;
  t.op = OP_MULGEN;
  t.k = k;
  t.oxy = oxy;
  t.oinf = oinf;
  return run_jobs(&t, n, nthreads, NULL);
}
/* sum_i k_i P_i: per-thread chunked constant-time lincomb (chunks of LINCOMB_CHUNK terms), partial sums added */
int ecref_lincomb(int curve, size_t n, const uint8_t* k, const uint8_t* pxy, const uint8_t* pinf, uint8_t* oxy,
                  uint8_t* oinf, int nthreads) {
  ecref_init();
  if (n == 0) {
    memset(oxy, 0, 64);
    *oinf = 1;
    return 0;
  }
  job t, *jobs = NULL;
  memset(&t, 0, sizeof t);
  t.curve = curve;
  t.op = OP_LINCOMB;
  t.k = k;
  t.pxy = pxy;
  t.pinf = pinf;
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > n) nthreads = (int)n;
  int err = run_jobs(&t, n, nthreads, &jobs);
  if (!err) {
    if (curve == 0) {
      kpt acc;
      kpt_identity(&acc);
      for (int i = 0; i < nthreads; i++) kpt_add(&acc, &acc, &jobs[i].kacc);
      k_store_point(oxy, oinf, &acc);
    } else {
      ppt acc;
      ppt_identity(&acc);
      for (int i = 0; i < nthreads; i++) ppt_add(&acc, &acc, &jobs[i].pacc);
      p_store_point(oxy, oinf, &acc);
    }
  }
  free(jobs);
  return err;
}
/* op: 0 add 1 sub 2 neg 3 mul 4 sqr 5 inv  (canonical big-endian in/out; inputs must be < p) */
int ecref_field_op(int curve, int op, size_t n, const uint8_t* a, const uint8_t* b, uint8_t* out) {
  ecref_init();
  for (size_t i = 0; i < n; i++) {
    if (curve == 0) {
      fe52 x, y, r;
      fe52_from_be(&x, a + 32 * i);
      if (b)
        fe52_from_be(&y, b + 32 * i);
      else
        y = x;
      switch (op) {
        case 0: fe52_add(&r, &x, &y); break;
        case 1:
          fe52_negate(&r, &y, 1);
          fe52_add(&r, &r, &x);
          break;
        case 2: fe52_negate(&r, &x, 1); break;
        case 3: fe52_mul(&r, &x, &y); break;
        case 4: fe52_sqr(&r, &x); break;
        default: fe52_inv(&r, &x); break;
      }
      fe52_normalize(&r);
      fe52_to_be(out + 32 * i, &r);
    } else {
      fp256 x, y, r;
      fp_from_be(&x, a + 32 * i);
      if (b)
        fp_from_be(&y, b + 32 * i);
      else
        y = x;
      switch (op) {
        case 0: fp_add(&r, &x, &y); break;
        case 1: fp_sub(&r, &x, &y); break;
        case 2: fp_sub(&r, &P256_ZERO, &x); break;
        case 3: fp_mul(&r, &x, &y); break;
        case 4: fp_sqr(&r, &x); break;
        default: fp_inv(&r, &x); break;
      }
      fp_to_be(out + 32 * i, &r);
    }
  }
  return 0;
}
/* radix-16 / wNAF recoders exposed for the property tests (primeorder/src/tables/radix16.rs:110-172) */
void ecref_radix16(const uint8_t* k_be, int nd, int8_t* out) { radix16(out, nd, k_be); }
int ecref_wnaf(const uint8_t* le, size_t len, size_t bit_len, int window, int8_t* out) { return wnaf_form(out, le, len, bit_len, window); }
void ecref_glv(const uint8_t* k_be, uint8_t* r1_be, uint8_t* r2_be) {
  uint64_t k[4], r1[4], r2[4];
  sc_from_be(k, k_be);
  glv_decompose(r1, r2, k);
  sc_to_be(r1_be, r1);
  sc_to_be(r2_be, r2);
}
