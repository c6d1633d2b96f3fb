/* ecref_prime.c — CPU restatement of the reference's generic prime-order path for the curves it binds to primeorder
 * over a Montgomery field synthesised from the modulus: sm2, brainpoolP256r1/t1, brainpoolP384r1/t1, bign-curve256v1,
 * P-224, P-192, P-521.
 *
 * TEST INFRASTRUCTURE ONLY: the checker / CPU baseline for those curves.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load it; the product path never does.
 *
 * What it follows (paths relative to the reference checkout, RustCrypto/elliptic-curves @ 739304e):
 *   field      primefield::monty_field_params! / monty_field_element! (primefield/src/monty.rs:319-375) over crypto-bigint
 *              0.7.5 ConstMontyForm (Cargo.lock:367-368, not under /root/reference): a*b*R^-1 mod p on fully reduced values,
 *              R = 2^(64 limbs); e.g. sm2/src/arithmetic/field.rs:34-60, bp256/src/arithmetic/field.rs:53-75,
 *              bp384/src/arithmetic/field.rs:53-75, bignp256/src/arithmetic/field.rs:59-80, p224/src/arithmetic/field.rs:54-75
 *              (U256 on 64-bit targets), p192/src/arithmetic/field.rs:54-75.  Restated as word-by-word (CIOS) Montgomery
 *              multiplication on 64-bit limbs.
 *   points     primeorder::ProjectivePoint<C>: EquationAIsMinusThree (primeorder/src/point_arithmetic.rs:222-245 add, RCB
 *              alg. 4; :289-318 double, RCB alg. 6) for sm2, brainpoolP256t1, brainpoolP384t1, P-224, P-192;
 *              EquationAIsGeneric (:63-111 add, RCB alg. 1; :166-207 double, RCB alg. 3) for brainpoolP256r1,
 *              brainpoolP384r1 and bign-curve256v1 (bignp256/src/arithmetic.rs:39).
 *   mul        primeorder/src/projective.rs:133-137, 532-557: constant-time lincomb over signed radix-16 digits
 *              (primeorder/src/tables/radix16.rs:31-55) with LookupTable select (primeorder/src/tables/lookup.rs:30-65).
 *   generator  mul_backend::VariableOnly for all of these (e.g. bp256/src/r1/arithmetic.rs:36): mul_by_generator(k) =
 *              GENERATOR * k through the same routine.
 *   records    big-endian, except bign-curve256v1 (ByteOrder::LittleEndian, bignp256/src/arithmetic/field.rs:65,
 *              bignp256/src/lib.rs:102).
 * Pinned by tests/test_oracle.py to the reference's own vectors where it holds any for these curves
 * (bignp256 / p224 / p192 src/test_vectors/group.rs) and to the big-integer model (oracle/pyref.py) for all of them. */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
enum { OP_MUL = 0, OP_MULGEN = 1, OP_LINCOMB = 2 };
#define LINCOMB_CHUNK 256

#define X_CAT2(a, b) a##_##b
#define X_CAT(a, b) X_CAT2(a, b)
#define X(name) X_CAT(name, NL)

#define NL 3
#include "ecref_prime_impl.inc"
#undef NL
#define NL 4
#include "ecref_prime_impl.inc"
#undef NL
#define NL 6
#include "ecref_prime_impl.inc"
#undef NL
#define NL 9
#include "ecref_prime_impl.inc"
#undef NL

/* curve ids as in include/ecgpu.h */
static curve_4 C_SM2, C_BP256R1, C_BP256T1, C_BIGN, C_P224;
static curve_6 C_BP384R1, C_BP384T1;
static curve_3 C_P192;
static curve_9 C_P521;
static int g_init = 0;
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;

void ecrefp_init(void) {
  pthread_mutex_lock(&g_lock);
  if (!g_init) {
    curve_init_4(&C_SM2, 3, 32, 0, 0, "fffffffeffffffffffffffffffffffffffffffff00000000ffffffffffffffff",
                 "fffffffeffffffffffffffffffffffff7203df6b21c6052b53bbf40939d54123",
                 "fffffffeffffffffffffffffffffffffffffffff00000000fffffffffffffffc",
                 "28e9fa9e9d9f5e344d5a9e4bcf6509a7f39789f515ab8f92ddbcbd414d940e93",
                 "32c4ae2c1f1981195f9904466a39c9948fe30bbff2660be1715a4589334c74c7",
                 "bc3736a2f4f6779c59bdcee36b692153d0a9877cc62a474002df32e52139f0a0");
    const char* bp256p = "a9fb57dba1eea9bc3e660a909d838d726e3bf623d52620282013481d1f6e5377";
    const char* bp256n = "a9fb57dba1eea9bc3e660a909d838d718c397aa3b561a6f7901e0e82974856a7";
    curve_init_4(&C_BP256R1, 4, 32, 0, 0, bp256p, bp256n, "7d5a0975fc2c3057eef67530417affe7fb8055c126dc5c6ce94a4b44f330b5d9",
                 "26dc5c6ce94a4b44f330b5d9bbd77cbf958416295cf7e1ce6bccdc18ff8c07b6",
                 "8bd2aeb9cb7e57cb2c4b482ffc81b7afb9de27e1e3bd23c23a4453bd9ace3262",
                 "547ef835c3dac4fd97f8461a14611dc9c27745132ded8e545c1d54c72f046997");
    curve_init_4(&C_BP256T1, 5, 32, 0, 0, bp256p, bp256n, "a9fb57dba1eea9bc3e660a909d838d726e3bf623d52620282013481d1f6e5374",
                 "662c61c430d84ea4fe66a7733d0b76b7bf93ebc4af2f49256ae58101fee92b04",
                 "a3e8eb3cc1cfe7b7732213b23a656149afa142c47aafbc2b79a191562e1305f4",
                 "2d996c823439c56d7f7b22e14644417e69bcb6de39d027001dabe8f35b25c9be");
    /* bign-curve256v1: constants of bignp256/src/arithmetic.rs:42-56 read little-endian; EquationAIsGeneric (:39) */
    curve_init_4(&C_BIGN, 6, 32, 1, 1, "ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff43",
                 "ffffffffffffffffffffffffffffffffd95c8ed60dfb4dfc7e5abf99263d6607",
                 "ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff40",
                 "77ce6c1515f3a8edd2c13aabe4d8fbbe4cf55069978b9253b22e7d6bd69c03f1", "0",
                 "6bf7fc3cfb16d69f5ce4c9a351d6835d78913966c408f6521e29cf1804516a93");
    const char* bp384p = "8cb91e82a3386d280f5d6f7e50e641df152f7109ed5456b412b1da197fb71123acd3a729901d1a71874700133107ec53";
    const char* bp384n = "8cb91e82a3386d280f5d6f7e50e641df152f7109ed5456b31f166e6cac0425a7cf3ab6af6b7fc3103b883202e9046565";
    curve_init_6(&C_BP384R1, 7, 48, 0, 0, bp384p, bp384n,
                 "7bc382c63d8c150c3c72080ace05afa0c2bea28e4fb22787139165efba91f90f8aa5814a503ad4eb04a8c7dd22ce2826",
                 "04a8c7dd22ce28268b39b55416f0447c2fb77de107dcd2a62e880ea53eeb62d57cb4390295dbc9943ab78696fa504c11",
                 "1d1c64f068cf45ffa2a63a81b7c13f6b8847a3e77ef14fe3db7fcafe0cbd10e8e826e03436d646aaef87b2e247d4af1e",
                 "8abe1d7520f9c2a45cb1eb8e95cfd55262b70b29feec5864e19c054ff99129280e4646217791811142820341263c5315");
    curve_init_6(&C_BP384T1, 8, 48, 0, 0, bp384p, bp384n,
                 "8cb91e82a3386d280f5d6f7e50e641df152f7109ed5456b412b1da197fb71123acd3a729901d1a71874700133107ec50",
                 "7f519eada7bda81bd826dba647910f8c4b9346ed8ccdc64e4b1abd11756dce1d2074aa263b88805ced70355a33b471ee",
                 "18de98b02db9a306f2afcd7235f72a819b80ab12ebd653172476fecd462aabffc4ff191b946a5f54d8d0aa2f418808cc",
                 "25ab056962d30651a114afd2755ad336747f93475b7a1fca3b88f2b6a208ccfe469408584dc2b2912675bf5b9e582928");
    curve_init_4(&C_P224, 9, 28, 0, 0, "ffffffffffffffffffffffffffffffff000000000000000000000001",
                 "ffffffffffffffffffffffffffff16a2e0b8f03e13dd29455c5c2a3d", "fffffffffffffffffffffffffffffffefffffffffffffffffffffffe",
                 "b4050a850c04b3abf54132565044b0b7d7bfd8ba270b39432355ffb4", "b70e0cbd6bb4bf7f321390b94a03c1d356c21122343280d6115c1d21",
                 "bd376388b5f723fb4c22dfe6cd4375a05a07476444d5819985007e34");
    curve_init_3(&C_P192, 10, 24, 0, 0, "fffffffffffffffffffffffffffffffeffffffffffffffff", "ffffffffffffffffffffffff99def836146bc9b1b4d22831",
                 "fffffffffffffffffffffffffffffffefffffffffffffffc", "64210519e59c80e70fa7e9ab72243049feb8deecc146b9b1",
                 "188da80eb03090f67cbf20eb43a18800f4ff0afd82ff1012", "07192b95ffc8da78631011ed6b24cdd573f977a11e794811");
    /* P-521: p521/src/arithmetic.rs:45-90 (the reference's field there is fiat-crypto's unsaturated Solinas form, not under
     * /root/reference; same residues), 66-byte records, U576 on 64-bit targets (p521/src/lib.rs:47) */
    curve_init_9(&C_P521, 11, 66, 0, 0,
                 "1ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff",
                 "1fffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffa51868783bf2f966b7fcc0148f709a5d03bb5c9b8899c47aebb6fb71e91386409",
                 "1fffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffc",
                 "51953eb9618e1c9a1f929a21a0b68540eea2da725b99b315f3b8b489918ef109e156193951ec7e937b1652c0bd3bb1bf073573df883d2c34f1ef451fd46b503f00",
                 "c6858e06b70404e9cd9e3ecb662395b4429c648139053fb521f828af606b4d3dbaa14b5e77efe75928fe1dc127a2ffa8de3348b3c1856a429bf97e7e31c2e5bd66",
                 "11839296a789a3bc0045c8a5fb42c7d1bd998f54449579b446817afbd17273e662c97ee72995ef42640c550b9013fad0761353c7086a272c24088be94769fd16650");
    g_init = 1;
  }
  pthread_mutex_unlock(&g_lock);
}

static int dispatch(int curve, int op, size_t n, const uint8_t* k, const uint8_t* pxy, const uint8_t* pinf, uint8_t* oxy, uint8_t* oinf,
                    int nthreads) {
  ecrefp_init();
  switch (curve) {
    case 3: return run_4(&C_SM2, op, n, k, pxy, pinf, oxy, oinf, nthreads);
    case 4: return run_4(&C_BP256R1, op, n, k, pxy, pinf, oxy, oinf, nthreads);
    case 5: return run_4(&C_BP256T1, op, n, k, pxy, pinf, oxy, oinf, nthreads);
    case 6: return run_4(&C_BIGN, op, n, k, pxy, pinf, oxy, oinf, nthreads);
    case 7: return run_6(&C_BP384R1, op, n, k, pxy, pinf, oxy, oinf, nthreads);
    case 8: return run_6(&C_BP384T1, op, n, k, pxy, pinf, oxy, oinf, nthreads);
    case 9: return run_4(&C_P224, op, n, k, pxy, pinf, oxy, oinf, nthreads);
    case 10: return run_3(&C_P192, op, n, k, pxy, pinf, oxy, oinf, nthreads);
    case 11: return run_9(&C_P521, op, n, k, pxy, pinf, oxy, oinf, nthreads);
    default: return 1;
  }
}
/* out[i] = k[i] * P[i]; records as include/ecgpu.h states them for the curve; 0 ok, 1 unknown curve, 2 scalar >= n, 3 off curve */
int ecrefp_mul_batch(int curve, size_t n, const uint8_t* k, const uint8_t* pxy, const uint8_t* pinf, uint8_t* oxy, uint8_t* oinf,
                     int nthreads) {
  return dispatch(curve, OP_MUL, n, k, pxy, pinf, oxy, oinf, nthreads);
}
int ecrefp_mul_gen_batch(int curve, size_t n, const uint8_t* k, uint8_t* oxy, uint8_t* oinf, int nthreads) {
  return dispatch(curve, OP_MULGEN, n, k, NULL, NULL, oxy, oinf, nthreads);
}
/* out = sum_i k[i] * P[i]: reference-style lincomb calls of LINCOMB_CHUNK terms, partial sums added */
int ecrefp_lincomb(int curve, size_t n, const uint8_t* k, const uint8_t* pxy, const uint8_t* pinf, uint8_t* oxy, uint8_t* oinf,
                   int nthreads) {
  return dispatch(curve, OP_LINCOMB, n, k, pxy, pinf, oxy, oinf, nthreads);
}
