"""The reference's proptest properties (k256/tests/projective.rs:21-141, p256/tests/projective.rs:54-149) restated
with hypothesis, same generators: scalar() = 32 random bytes reduced mod n, projective() = mul_by_generator(scalar()).
Each property runs against two back ends through the same code: the CUDA kernels executed on the host
(`-m "not gpu"`, tests/sim) and the C ABI on the GPU (`-m gpu`)."""
import ctypes

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import pyref
from helpers import pack_points, pack_scalars, unpack_points
from test_sim import sim  # noqa: F401
from test_sim_kernels import CID, _p, fb_tables  # noqa: F401

CURVES = ["k256", "p256"]
SETTINGS = dict(max_examples=40, deadline=None, derandomize=True,
                suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
bytes32 = st.binary(min_size=32, max_size=32)
# proptest's any::<[u8; 32]>() is uniform; hypothesis shrinks towards zeros and likes boundary patterns, which is welcome
batch = st.lists(st.tuples(bytes32, bytes32, bytes32), min_size=1, max_size=5)


class SimBackend:
    """ecg_* semantics on top of the host-executed kernels (tests/sim/sim.cpp)."""

    def __init__(self, lib, tables):
        self.lib, self.tables = lib, tables

    def mul_batch(self, curve, K, xy, inf):
        n = K.size // 32
        oxy, oinf, stt = np.zeros(64 * n, np.uint8), np.zeros(n, np.uint8), np.zeros(2, np.uint32)
        self.lib.simk_mul_batch(CID[curve], ctypes.c_size_t(n), _p(K), _p(xy), _p(inf), _p(oxy), _p(oinf), _p(stt))
        assert stt[0] == 0
        return oxy.reshape(n, 64), oinf

    def mul_by_generator(self, curve, K):
        n = K.size // 32
        oxy, oinf, stt = np.zeros(64 * n, np.uint8), np.zeros(n, np.uint8), np.zeros(2, np.uint32)
        self.lib.simk_mul_gen_batch(CID[curve], ctypes.c_size_t(n), _p(K), _p(self.tables[curve]), _p(oxy), _p(oinf), _p(stt))
        assert stt[0] == 0
        return oxy.reshape(n, 64), oinf

    def mul_by_generator_and_mul_add(self, curve, A, B, xy, inf):
        n = A.size // 32
        oxy, oinf, stt = np.zeros(64 * n, np.uint8), np.zeros(n, np.uint8), np.zeros(2, np.uint32)
        self.lib.simk_mul_gen_add_batch(CID[curve], ctypes.c_size_t(n), _p(A), _p(B), _p(xy), _p(inf), _p(self.tables[curve]),
                                        _p(oxy), _p(oinf), _p(stt))
        assert stt[0] == 0
        return oxy.reshape(n, 64), oinf


    def lincomb(self, curve, K, xy, inf):
        n = K.size // 32
        oxy, oinf, stt = np.zeros(64, np.uint8), np.zeros(1, np.uint8), np.zeros(2, np.uint32)
        path = ctypes.c_int(-1)
        xyc = np.ascontiguousarray(xy)
        self.lib.simk_lincomb(CID[curve], ctypes.c_size_t(n), _p(K), _p(xyc), _p(inf), ctypes.c_size_t(1 << 13),
                              _p(oxy), _p(oinf), _p(stt), ctypes.byref(path))
        assert stt[0] == 0
        return oxy, int(oinf[0])


def _check(be, curve, rows, with_lincomb):
    c = pyref.CURVES[curve]
    G = pyref.G(c)
    s = [int.from_bytes(r[0], "big") % c.n for r in rows]          # scalar()
    t = [int.from_bytes(r[1], "big") % c.n for r in rows]          # projective() = G * scalar()
    a = [int.from_bytes(r[2], "big") % c.n for r in rows]
    S, T, A = pack_scalars(s), pack_scalars(t), pack_scalars(a)
    # projective(): mul_by_generator; property mul_by_generator == GENERATOR * s (fixed-base vs variable-base kernel)
    pxy, pinf = be.mul_by_generator(curve, T)
    gxy, ginf = pack_points([G] * len(rows))
    vxy, vinf = be.mul_batch(curve, T, gxy, ginf)
    assert np.array_equal(np.asarray(pxy), np.asarray(vxy)) and np.array_equal(pinf, vinf)
    P = unpack_points(pxy, pinf)
    assert P == [pyref.mul(c, ti, G) for ti in t]
    # p * s against the big-integer model (the reference compares its two implementations; here: kernel vs definition)
    mxy, minf = be.mul_batch(curve, S, np.asarray(pxy).reshape(-1), pinf)
    prod = unpack_points(mxy, minf)
    assert prod == [pyref.mul(c, si, Pi) for si, Pi in zip(s, P)]
    # mul_by_generator_and_mul_add_vartime == G * a + p * s
    axy, ainf = be.mul_by_generator_and_mul_add(curve, A, S, np.asarray(pxy).reshape(-1), pinf)
    assert unpack_points(axy, ainf) == [pyref.add(c, pyref.mul(c, ai, G), pr) for ai, pr in zip(a, prod)]
    if with_lincomb:
        # lincomb(&[(p1, s1), ...]) == p1 * s1 + p2 * s2 + ...
        lxy, linf = be.lincomb(curve, S, np.asarray(pxy).reshape(-1), pinf)
        want = None
        for pr in prod:
            want = pyref.add(c, want, pr)
        assert pyref.dec_point(np.asarray(lxy).tobytes(), int(linf)) == want


@pytest.mark.parametrize("curve", CURVES)
def test_properties_on_host_executed_kernels(sim, fb_tables, curve):
    be = SimBackend(sim, fb_tables)

    @settings(**SETTINGS)
    @given(rows=batch)
    def run(rows):
        _check(be, curve, rows, with_lincomb=True)

    run()


@pytest.mark.gpu
@pytest.mark.parametrize("curve", CURVES)
def test_properties_through_the_c_abi(engine, curve):
    @settings(**SETTINGS)
    @given(rows=batch)
    def run(rows):
        _check(engine, curve, rows, with_lincomb=True)

    run()
