"""P-521 after the Mersenne reduction: two whole waves of variable-base multiplications (75 776 pairs on 148 SMs) through the C ABI,
EVERY output against oracle/ecref_prime.c, kernel time by the library's CUDA events; plus the fixed-base path and the bucket method."""
import json, os, random, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "elliptic-curves_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import ecgpu, pyref, ecref
from test_curves_ext import recs

def main():
    name = "p521"
    c = pyref.CURVES[name]
    eng = ecgpu.Engine()
    rng = random.Random(521)
    n = 2 * 148 * 2 * 128
    ks = [rng.randrange(c.n) for _ in range(n)]
    ts = [rng.randrange(1, c.n) for _ in range(n)]
    K, T = recs(c, ks), recs(c, ts)
    pxy, pinf = eng.mul_by_generator(name, T)
    pxy = np.ascontiguousarray(pxy).reshape(-1)
    t0 = time.perf_counter()
    g_xy, g_inf = ecref.mul_gen_batch(name, T, nthreads=os.cpu_count() or 8)
    cpu_s = time.perf_counter() - t0
    gen_ok = bool(np.array_equal(np.asarray(g_xy).reshape(-1), pxy) and not pinf.any())
    eng.mul_batch(name, K, pxy, None)
    eng.timing_enable(True)
    for _ in range(3):
        xy, inf = eng.mul_batch(name, K, pxy, None)
    ms, calls = eng.timing_read()
    eng.timing_enable(False)
    r_xy, r_inf = ecref.mul_batch(name, K, pxy, np.zeros(n, np.uint8), nthreads=os.cpu_count() or 8)
    var_ok = bool(np.array_equal(np.asarray(xy).reshape(-1), np.asarray(r_xy).reshape(-1)) and np.array_equal(inf, r_inf))
    m = 9001
    l_xy, l_inf = eng.lincomb(name, K[:66 * m], pxy[:132 * m], None)
    o_xy, o_inf = ecref.lincomb(name, K[:66 * m], pxy[:132 * m], np.zeros(m, np.uint8), nthreads=8)
    lin_ok = bool(np.array_equal(np.asarray(l_xy), o_xy) and int(l_inf) == int(o_inf))
    print(json.dumps({"curve": name, "pairs": n, "kernel_ms": ms / 3, "calls": calls, "value": n / (ms / 3 * 1e-3), "unit": "scalar-mults/s",
                      "bit_exact_var_base": var_ok, "bit_exact_fixed_base": gen_ok, "bit_exact_lincomb_9001": lin_ok,
                      "cpu_fixed_base_s": cpu_s, "before": {"kernel_ms": 44.52, "value": 1.702e6, "file": "profiles/r02_bench_n1_call21.json"}}))
    eng.close()
    sys.exit(0 if (var_ok and gen_ok and lin_ok) else 1)

if __name__ == "__main__":
    main()
