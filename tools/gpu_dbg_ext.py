"""debug helper: var-base vs fixed-base on the generic-Montgomery curves, small batches"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "elliptic-curves_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import ecgpu, pyref
from test_curves_ext import recs, pts, unpack
eng = ecgpu.Engine()
for cid, c in sorted(pyref.EXT_CURVES.items()):
    ks = [1, 2, 3, 5, 65535, 65536, 65537, 2**32 + 1, c.n - 1, 0]
    G = pyref.G(c)
    want = [pyref.mul(c, k, G) for k in ks]
    pxy, pinf = pts(c, [G] * len(ks))
    try:
        xy, inf = eng.mul_batch(c.name, recs(c, ks), pxy, pinf)
        got = unpack(c, xy, inf)
        print(cid, c.name, "varbase", [g == w for g, w in zip(got, want)])
    except Exception as e:
        print(cid, c.name, "varbase EXC", e)
    try:
        oxy = np.full(2 * pyref.fbytes(c) * len(ks), 0xAB, np.uint8); oinf = np.full(len(ks), 0xCD, np.uint8)
        xy, inf = eng.mul_by_generator(c.name, recs(c, ks), oxy, oinf)
        got = unpack(c, xy, inf)
        print(cid, c.name, "fixedbase", [g == w for g, w in zip(got, want)], "inf", list(inf[:4]), "raw", xy.reshape(-1)[:6])
    except Exception as e:
        print(cid, c.name, "fixedbase EXC", e)
