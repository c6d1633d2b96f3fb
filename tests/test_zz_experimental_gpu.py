"""GPU checks of code paths that are off by default (selected by environment variables, measured in a later round).
Kept in the last test file so that a failure here cannot hide anything in the main parity suite under `-x`."""
import random

import numpy as np
import pytest

import ecref
import pyref
from helpers import pack_points, pack_scalars, random_points

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("curve", ["k256", "p256"])
@pytest.mark.parametrize("bucket_k", ["4", "8"])
def test_lincomb_warp_balanced_bucket_kernel(engine, monkeypatch, curve, bucket_k):
    """ECG_MSM_BUCKETS_PER_THREAD = 4 | 8 -> msm_bucket_sorted_kernel; same result as the default kernel and the oracle.
    The host-executed version of the same kernel is covered by tests/test_sim_kernels.py."""
    c = pyref.CURVES[curve]
    rng = random.Random(900 + int(bucket_k))
    n = (1 << 14) + 77
    base = random_points(c, 48, seed=9) + [None]
    Ps = [base[min(rng.randrange(70), 48)] for _ in range(n)]
    ks = [rng.randrange(c.n) for _ in range(n)]
    xy, inf = pack_points(Ps)
    K = pack_scalars(ks)
    monkeypatch.delenv("ECG_MSM_BUCKETS_PER_THREAD", raising=False)
    d_xy, d_inf = engine.lincomb(curve, K, xy, inf)
    monkeypatch.setenv("ECG_MSM_BUCKETS_PER_THREAD", bucket_k)
    s_xy, s_inf = engine.lincomb(curve, K, xy, inf)
    assert np.array_equal(d_xy, s_xy) and d_inf == s_inf
    r_xy, r_inf = ecref.lincomb(curve, K, xy, inf, nthreads=8)
    assert np.array_equal(s_xy, r_xy) and s_inf == r_inf
