"""compute-sanitizer workload for the code added in the second half of round 2: the curves on the generic Montgomery field
(incl. P-521's byte-granular records and the 7-limb P-224), the bucket method over them, generic ECDSA / a*G + b*P, hash to
curve (four suites), the constant-time ctx.  Tiny sizes; results still checked against the oracle / model.
    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_ext.py"""
import os, random, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "elliptic-curves_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import ecgpu, pyref, ecref
from test_curves_ext import recs, pts, unpack

def main():
    eng = ecgpu.Engine()
    ct = ecgpu.Engine(consttime=True)
    rng = random.Random(9)
    names = sys.argv[1:] or ["sm2", "bp256r1", "bignp256", "bp384r1", "p224", "p192", "p521", "p384"]
    for name in names:
        c = pyref.CURVES[name]
        n = 70
        ks = [rng.randrange(c.n) for _ in range(n)]; ks[0] = 0; ks[1] = c.n - 1
        base = [pyref.mul(c, rng.randrange(1, c.n), pyref.G(c)) for _ in range(4)]
        Ps = [base[i % 4] for i in range(n)]; Ps[2] = None
        K = recs(c, ks); pxy, pinf = pts(c, Ps)
        xy, inf = eng.mul_batch(name, K, pxy, pinf)
        r_xy, r_inf = ecref.mul_batch(name, K, pxy, pinf, nthreads=8)
        assert np.array_equal(np.asarray(xy).reshape(-1), r_xy.reshape(-1)) and np.array_equal(inf, r_inf), (name, "mul")
        cxy, cinf = ct.mul_batch(name, K, pxy, pinf)
        assert np.array_equal(cxy, xy) and np.array_equal(cinf, inf), (name, "ct mul")
        gxy, ginf = ct.mul_by_generator(name, K)
        r_xy, r_inf = ecref.mul_gen_batch(name, K, nthreads=8)
        assert np.array_equal(np.asarray(gxy).reshape(-1), r_xy.reshape(-1)), (name, "ct gen")
        eng.mul_batch_x(name, K, pxy, pinf)
        if name != "p521":      # the fixed-base table of P-521 is 1.1 M variable-base multiplications: too slow under the tool
            gxy, ginf = eng.mul_by_generator(name, K)
            assert np.array_equal(np.asarray(gxy).reshape(-1), r_xy.reshape(-1)) and np.array_equal(ginf, r_inf), (name, "gen")
            a = recs(c, [rng.randrange(c.n) for _ in range(n)])
            eng.mul_by_generator_and_mul_add(name, a, K, pxy, pinf)
            if name not in ("sm2", "bignp256"):
                nb = pyref.fbytes(c)
                z = [rng.randrange(1 << (8 * nb)) for _ in range(16)]
                sig, q = [], []
                for zi in z:
                    d = rng.randrange(1, c.n)
                    r, s = pyref.ecdsa_sign(c, d, zi % c.n, rng.randrange(1, c.n))
                    Q = pyref.mul(c, d, pyref.G(c))
                    sig.append(r.to_bytes(nb, "big") + s.to_bytes(nb, "big")); q.append(Q[0].to_bytes(nb, "big") + Q[1].to_bytes(nb, "big"))
                v = eng.ecdsa_verify_batch(name, np.frombuffer(b"".join(zi.to_bytes(nb, "big") for zi in z), np.uint8),
                                           np.frombuffer(b"".join(sig), np.uint8), np.frombuffer(b"".join(q), np.uint8))
                assert v.all(), (name, "ecdsa")
        for m in (5, 8200):
            kk = recs(c, [rng.randrange(c.n) for _ in range(m)])
            pp, pi = pts(c, [Ps[i % n] for i in range(m)])
            lxy, linf = eng.lincomb(name, kk, pp, pi)
            rxy, rinf = ecref.lincomb(name, kk, pp, pi, nthreads=8)
            assert np.array_equal(np.asarray(lxy), rxy) and int(linf) == int(rinf), (name, "lincomb", m)
        part = eng.lincomb_partial(name, K, pxy, pinf)
        eng.point_sum(name, np.concatenate([part, part]))
        fa, fb2 = recs(c, [rng.randrange(c.p) for _ in range(n)]), recs(c, [rng.randrange(c.p) for _ in range(n)])
        for op in range(6):
            eng.field_op(name, op, fa, fb2)
        if name in ("p384", "p521"):
            msgs = [bytes(rng.randrange(256) for _ in range(L)) for L in (0, 1, 111, 112, 127, 128, 129, 300)]
            dst = b"QUUX-V01-CS02-sanitize"
            xy, inf = eng.hash_to_curve(name, msgs, dst)
            nb = pyref.fbytes(c)
            assert [pyref.dec_point(xy[i].tobytes(), int(inf[i]), nb) for i in range(len(msgs))] == [pyref.hash_to_curve(name, m, dst) for m in msgs]
            eng.encode_to_curve(name, msgs, dst)
            eng.hash_to_scalar(name, msgs, dst)
    for name in ("k256", "p256"):
        msgs = [bytes(rng.randrange(256) for _ in range(L)) for L in (0, 1, 55, 56, 63, 64, 65, 300)]
        dst = b"QUUX-V01-CS02-sanitize"
        xy, inf = eng.hash_to_curve(name, msgs, dst)
        assert [pyref.dec_point(xy[i].tobytes(), int(inf[i])) for i in range(len(msgs))] == [pyref.hash_to_curve(name, m, dst) for m in msgs]
        eng.encode_to_curve(name, msgs, b"Z" * 300)
        eng.hash_to_scalar(name, msgs, dst)
    eng.close(); ct.close()
    print("sanitize_ext workload OK")

if __name__ == "__main__":
    main()
