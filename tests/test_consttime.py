"""ECG_FLAG_CONSTTIME (the reference's constant-time `Mul` / `lincomb`, k256/src/arithmetic/mul.rs:112-163,
LookupTable::select primeorder/src/tables/lookup.rs:43-65): masked window-table selects, branch-free GLV sign folding, k*G
through the variable-base routine, per-term lincomb.  The results must be the same points as the default (vartime) path."""
import ctypes
import os
import random

import numpy as np
import pytest

import ecref
import pyref
from test_curves_ext import pts, recs, unpack

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["k256", "p256", "p384", "bp256r1", "bignp256", "p224"]


def _p(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)


def _inputs(c, n, seed):
    rng = random.Random(seed)
    ks = [rng.randrange(c.n) for _ in range(n)]
    ks[:8] = [0, 1, 2, c.n - 1, c.n - 2, 15, 16, (c.n - 1) // 2]
    base = [pyref.mul(c, rng.randrange(1, c.n), pyref.G(c)) for _ in range(6)]
    Ps = [base[i % 6] for i in range(n)]
    Ps[9] = None
    return ks, Ps


@pytest.mark.parametrize("name", NAMES)
def test_consttime_kernels_on_host(name):
    import __graft_entry__ as ge
    ge.build()
    sim = ctypes.CDLL(os.path.join(HERE, "sim", "libecgsim.so"))
    c = pyref.CURVES[name]
    cid = pyref.CURVE_IDS[name]
    nb = pyref.fbytes(c)
    n = 40
    ks, Ps = _inputs(c, n, 17)
    K = recs(c, ks)
    pxy, pinf = pts(c, Ps)
    oxy, oinf, st = np.zeros(2 * nb * n, np.uint8), np.zeros(n, np.uint8), np.zeros(2, np.uint32)
    sim.simk_mul_batch_ct(cid, ctypes.c_size_t(n), _p(K), _p(pxy), _p(pinf), _p(oxy), _p(oinf), _p(st))
    assert st[0] == 0
    assert unpack(c, oxy, oinf) == [pyref.mul(c, k, P) if P is not None else None for k, P in zip(ks, Ps)]
    # pxy == NULL: k * G through the same routine
    sim.simk_mul_batch_ct(cid, ctypes.c_size_t(n), _p(K), None, None, _p(oxy), _p(oinf), _p(st))
    assert st[0] == 0
    assert unpack(c, oxy, oinf) == [pyref.mul(c, k, pyref.G(c)) for k in ks]
    bad = list(ks)
    bad[5] = c.n
    Kb = recs(c, bad)
    sim.simk_mul_batch_ct(cid, ctypes.c_size_t(n), _p(Kb), None, None, _p(oxy), _p(oinf), _p(st))
    assert st[0] == 1 and st[1] == 5


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_consttime_ctx_matches_default_path(name):
    import ecgpu

    c = pyref.CURVES[name]
    nb = pyref.fbytes(c)
    ct = ecgpu.Engine(consttime=True)
    vt = ecgpu.Engine()
    try:
        n = 3000
        ks, Ps = _inputs(c, n, 23)
        K = recs(c, ks)
        pxy, pinf = pts(c, Ps)
        a_xy, a_inf = ct.mul_batch(name, K, pxy, pinf)
        b_xy, b_inf = vt.mul_batch(name, K, pxy, pinf)
        assert np.array_equal(a_xy, b_xy) and np.array_equal(a_inf, b_inf)
        assert unpack(c, a_xy, a_inf)[:12] == [pyref.mul(c, k, P) if P is not None else None for k, P in zip(ks[:12], Ps[:12])]
        x, xinf = ct.mul_batch_x(name, K, pxy, pinf)
        assert np.array_equal(x, np.asarray(a_xy).reshape(n, 2 * nb)[:, :nb])
        g_xy, g_inf = ct.mul_by_generator(name, K)               # no fixed-base table in this mode
        h_xy, h_inf = vt.mul_by_generator(name, K)
        assert np.array_equal(g_xy, h_xy) and np.array_equal(g_inf, h_inf)
        # lincomb: 9001 terms would take the bucket method by default; the constant-time ctx sums per term
        m = 9001
        rng = random.Random(5)
        ks2 = [rng.randrange(c.n) for _ in range(m)]
        Ps2 = [Ps[i % 8] for i in range(m)]
        K2 = recs(c, ks2)
        pxy2, pinf2 = pts(c, Ps2)
        l1 = ct.lincomb(name, K2, pxy2, pinf2)
        l2 = vt.lincomb(name, K2, pxy2, pinf2)
        assert np.array_equal(l1[0], l2[0]) and l1[1] == l2[1]
        bad = list(ks[:50])
        bad[31] = c.n
        with pytest.raises(ecgpu.ScalarRangeError) as ei:
            ct.mul_by_generator(name, recs(c, bad))
        assert ei.value.index == 31
    finally:
        ct.close()
        vt.close()
