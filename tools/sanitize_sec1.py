"""compute-sanitizer workload for the last widening steps of round 2: generic SEC1 decompression (incl. P-224's Tonelli-Shanks and
bign's little-endian records: byte-granular, unaligned loads), the generic field square root, SM2DSA verification.  Tiny sizes;
results checked against the model (tests/test_sec1_ext.py, tests/test_sm2dsa.py hold the case builders).
    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_sec1.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "elliptic-curves_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import ecgpu, pyref
import test_sec1_ext as t1
import test_sm2dsa as t2
from test_curves_ext import recs

def main():
    eng = ecgpu.Engine()
    for name in t1.CURVES:
        c = pyref.CURVES[name]
        r, want = t1.cases(c, 20, 5)
        xy, inf, valid = eng.decompress_batch(name, np.frombuffer(b"".join(r), np.uint8))
        t1.check(c, want, xy, inf, valid)
        if name != "p224":
            a, wroot = t1.sqrt_cases(c, 6)
            out, ok = eng.field_sqrt(name, recs(c, a))
            assert [pyref.dec_fe(c, out[i].tobytes()) if ok[i] else None for i in range(len(a))] == wroot, name
    cases = t2.made_cases(20, 4)
    E, S, Q, exp = t2.pack(cases)
    assert [bool(v) for v in eng.sm2dsa_verify_batch(E, S, Q)] == exp
    eng.close()
    print("sanitize_sec1 workload OK")

if __name__ == "__main__":
    main()
