"""one variable-base call (2^16 pairs) per curve given on the command line, for `ncu -k regex:generic_varbase_kernel`"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "elliptic-curves_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import ecgpu, pyref
from test_gpu_curves_ext import rand_scalars
eng = ecgpu.Engine()
for name in sys.argv[1:]:
    c = pyref.CURVES[name]
    n = 1 << 16
    K = rand_scalars(c, n, 1).reshape(-1)
    T = rand_scalars(c, n, 2); T[:, 0 if c.le else -1] |= 1
    pxy, pinf = eng.mul_by_generator(name, np.ascontiguousarray(K))      # points k*G (table build + fixed-base kernel)
    for _ in range(3):
        eng.mul_batch(name, T.reshape(-1), pxy, None)
    if name in ("k256", "p256", "p384", "p521"):
        msgs = [bytes([i & 255, i >> 8]) * 16 for i in range(1 << 14)]
        for _ in range(2):
            eng.hash_to_curve(name, msgs, b"QUUX-V01-CS02-ncu")
print("done")
