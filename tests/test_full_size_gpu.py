"""Parity at BENCHMARK size: config C3 - 1024 ConeSOC blocks of dim 64, dense A 65536 x 16384 (4.3 GB in
f32, generated in HBM) - iterated K in {1, 10} times by the device (fused DenseOp + ProductCone, pair fusion + speculative
pairing on, as bench.py times it) and by the f64 oracle over the SAME A (regenerated bit-identically on the host by
oracle/native.c, 8.6 GB in f64; one stacked MatOp = one dgemv per op) with the SAME b, c.  Tolerances are the ones bench.py's
`parity` object uses (relative l_inf of x_hat, y_hat: 5e-6 at K = 1, 5e-5 at K = 10; residual triple to 2 digits)."""
import ctypes as C
import math

import numpy as np
import pytest

import helpers as H
from helpers import capi, SOC
from totsu_b200 import host
import cpu_workloads as W
import totsu_oracle as O

pytestmark = pytest.mark.gpu


def test_c3_full_size_iterates_match_oracle():
    try:
        import psutil
        if psutil.virtual_memory().available < 24e9:
            pytest.skip("the f64 oracle needs ~20 GB of host memory for C3's A")
    except ImportError:
        pass
    capi.init(0)
    L = capi.lib()
    dt = np.float32
    nblk, bdim, n = 1024, 64, 16384
    m = nblk * bdim
    blocks = [(SOC, bdim)] * nblk
    scale = dt(1.0 / math.sqrt(n))
    abuf = capi.Buf(dtype=dt, length=m * n)
    capi.check(L.tb_fill_uniform_f32(abuf.view(), m, n, 0, 0, scale))
    # b = A x0 + s0, c = -A^T y0 through the backend itself (like bench.py), so device and oracle share b and c exactly
    rng = np.random.default_rng(12345)
    sc = 1.0 / math.sqrt(m)
    x0 = rng.standard_normal(n) * sc

    def interior():
        v = rng.standard_normal((nblk, bdim))
        v[:, 0] = np.linalg.norm(v[:, 1:], axis=1) + 1.0
        return v.reshape(m) * sc
    s0, y0 = interior(), interior()
    b = s0.astype(dt); c = np.zeros(n, dtype=dt)
    hop = C.c_int64()
    capi.check(L.tb_denseop_create(capi.TB_F32, abuf.view(), m, n, 0, m, C.byref(hop)))
    bx, by, bb, bc = capi.Buf(x0.astype(dt), mutable=False), capi.Buf(y0.astype(dt), mutable=False), capi.Buf(b), capi.Buf(c)
    capi.check(L.tb_denseop_apply_f32(hop.value, 0, 1.0, bx.view(), 1.0, bb.view()))
    capi.check(L.tb_denseop_apply_f32(hop.value, 1, -1.0, by.view(), 0.0, bc.view()))
    for bf in (bx, by, bb, bc):
        bf.release()
    capi.check(L.tb_denseop_destroy(hop.value))
    assert np.abs(b).max() > 0 and np.abs(c).max() > 0
    # ---- device
    ks = [1, 10]
    s = host.Session.dense(dt, abuf.view(), m, n, c, b, blocks, fused_op=True, fused_cone=True)
    assert s.begin(max_iter=None, eps_acc=0.0, eps_inf=0.0, device_precond=True) == "None"
    dev, done = {}, 0
    p0 = capi.pairs_fused()
    for k in ks:
        s.step(k - done); done = k
        dev[k] = s.xy() + ((s.last.c0, s.last.c1, s.last.c2),)
    assert capi.pairs_fused() - p0 == 3 * max(ks)            # the timed configuration: every pair served by one read of A
    s.close()
    abuf.release()
    # ---- oracle over the same A
    a64 = W.fill_f64(m, n, 0, 0, scale, True)
    # spot-check that host and device generated the same matrix: A x0 + s0 in f64 equals the device's f32 b to f32 rounding
    assert H.rel_linf(b, a64 @ x0.astype(dt).astype(np.float64) + s0.astype(dt).astype(np.float64)) < 5e-6
    snaps, trace = W.iterates(O, W.DenseProblem(O, a64, b, c, [("soc", bdim)] * nblk), ks)
    tol = {1: 5e-6, 10: 5e-5}
    for k in ks:
        ex, ey = H.rel_linf(dev[k][0], snaps[k][0]), H.rel_linf(dev[k][1], snaps[k][1])
        assert ex <= tol[k] and ey <= tol[k], (k, ex, ey)
        for got, want in zip(dev[k][2], trace[k - 1][1:]):
            assert abs(got - want) <= 5e-3 * max(abs(want), 1e-3), (k, got, want)
