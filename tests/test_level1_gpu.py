"""GPU parity: LinAlg level-1 ops and the SliceLike coherence protocol, through the C ABI, against the oracle's
F64LAPACK (oracle/totsu_oracle.py) on the same seeded inputs.  Edge cases: empty vectors, ragged lengths,
odd element offsets (the solver's dual block starts at n+2m+1), strided abssum, scale(0) over NaN."""
import ctypes as C

import numpy as np
import pytest

from helpers import O, capi, rel_linf

pytestmark = pytest.mark.gpu

DTYPES = [np.float32, np.float64]
TOL = {np.float32: 2e-6, np.float64: 1e-13}
LENS = [0, 1, 2, 31, 33, 255, 1000, 4097, 147457]


@pytest.fixture(scope="module", autouse=True)
def _init():
    capi.init(0)
    yield


def _views(dt, n, rng, k=2, pad=3):
    """k vectors of length n inside one root buffer at odd offsets."""
    total = pad + k * (n + pad)
    host = rng.standard_normal(total).astype(dt)
    buf = capi.Buf(host, mutable=True)
    offs = [pad + i * (n + pad) for i in range(k)]
    return host, buf, offs


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("n", LENS)
def test_norm_copy_scale_add_adds_di(dt, n):
    rng = np.random.default_rng(n + 1)
    L = capi.lib()
    F = C.c_float if dt == np.float32 else C.c_double
    host, buf, (ox, oy, od) = _views(dt, n, rng, k=3)
    ref = host.astype(np.float64).copy()
    x, y, d = (buf.view(o, n) for o in (ox, oy, od))
    rx, ry, rd = (ref[o:o + n] for o in (ox, oy, od))

    out = F()
    capi.check(capi.fn("tb_norm", dt)(x, C.byref(out)))
    assert abs(out.value - O.F64LAPACK.norm(rx)) <= TOL[dt] * max(1.0, O.F64LAPACK.norm(rx))

    capi.check(capi.fn("tb_add", dt)(0.75, x, y)); O.F64LAPACK.add(0.75, rx, ry)
    capi.check(capi.fn("tb_scale", dt)(-1.5, x)); O.F64LAPACK.scale(-1.5, rx)
    capi.check(capi.fn("tb_adds", dt)(0.125, y)); O.F64LAPACK.adds(0.125, ry)
    capi.check(capi.fn("tb_transform_di", dt)(2.0, d, x, 0.5, y)); O.F64LAPACK.transform_di(2.0, rd, rx, 0.5, ry)
    capi.check(capi.fn("tb_transform_di", dt)(1.0, d, y, 1.0, x)); O.F64LAPACK.transform_di(1.0, rd, ry, 1.0, rx)
    capi.check(capi.fn("tb_copy", dt)(x, d)); O.F64LAPACK.copy(rx, rd)
    capi.check(capi.fn("tb_transform_di", dt)(-1.0, x, x, 0.0, y)); O.F64LAPACK.transform_di(-1.0, rx, rx, 0.0, ry)
    buf.release()                      # SliceLike::drop: device-newer ranges flow back into `host`
    assert rel_linf(host, ref) <= 20 * TOL[dt]
    # untouched padding stays bit-identical
    assert host[0] == ref[0].astype(dt)


@pytest.mark.parametrize("dt", DTYPES)
def test_abssum_strided(dt):
    rng = np.random.default_rng(5)
    F = C.c_float if dt == np.float32 else C.c_double
    for n, inc in [(0, 1), (1, 1), (1000, 1), (1000, 7), (1001, 7), (63 * 50, 63), (10, 0), (5, 9)]:
        host = rng.standard_normal(n + 5).astype(dt)
        buf = capi.Buf(host, mutable=False)
        out = F(123.0)
        capi.check(capi.fn("tb_abssum", dt)(buf.view(5 if n else 0, n), inc, C.byref(out)))
        want = O.F64LAPACK.abssum(host[5:5 + n].astype(np.float64) if n else np.zeros(0), inc)
        assert abs(out.value - want) <= TOL[dt] * max(1.0, want) * 10, (n, inc)
        buf.release()


@pytest.mark.parametrize("dt", DTYPES)
def test_scale_zero_is_a_fill(dt):
    host = np.full(100, np.nan, dtype=dt)
    buf = capi.Buf(host)
    capi.check(capi.fn("tb_scale", dt)(0.0, buf.view(10, 50)))
    # beta == 0 must not read y either
    ones = capi.Buf(np.ones(40, dtype=dt), mutable=False)
    capi.check(capi.fn("tb_transform_di", dt)(2.0, ones.view(), ones.view(), 0.0, buf.view(60, 40)))
    buf.release(); ones.release()
    assert np.all(host[10:60] == 0) and np.all(np.isnan(host[:10])) and np.all(host[60:] == 2.0)


@pytest.mark.parametrize("dt", DTYPES)
def test_coherence_get_set_host_views(dt):
    """slicelike.rs:47-69 semantics: get/set see device writes; get_mut makes the next device use re-upload."""
    L = capi.lib()
    F = C.c_float if dt == np.float32 else C.c_double
    host = np.arange(16, dtype=dt)
    buf = capi.Buf(host)
    v = buf.view()
    out = F()
    capi.check(capi.fn("tb_get1", dt)(v, 3, C.byref(out))); assert out.value == 3       # host copy current
    capi.check(capi.fn("tb_scale", dt)(2.0, v))                                          # device-newer
    capi.check(capi.fn("tb_get1", dt)(v, 3, C.byref(out))); assert out.value == 6
    capi.check(capi.fn("tb_set1", dt)(v, 4, 100.0))
    capi.check(capi.fn("tb_get1", dt)(v, 4, C.byref(out))); assert out.value == 100
    assert host[5] == 5                                                                  # not yet synced
    capi.check(L.tb_host_ref(buf.view(4, 4)))                                            # get_ref on a sub-slice
    assert list(host[4:8]) == [100, 10, 12, 14] and host[8] == 8
    capi.check(L.tb_host_mut(buf.view(8, 2)))                                            # get_mut: host becomes owner
    assert list(host[8:10]) == [16, 18]
    host[8] = -1.0
    capi.check(capi.fn("tb_adds", dt)(1.0, v))                                           # re-uploads [8,10) first
    buf.release()
    want = np.arange(16, dtype=np.float64) * 2 + 1
    want[4] = 101; want[8] = 0
    assert np.array_equal(host, want.astype(dt))

@pytest.mark.parametrize("dt", DTYPES)
def test_set_of_the_value_just_read_launches_nothing(dt):
    """The solver writes back what it has just read - tau := max(tau, 0), kappa := min(kappa, 0) (solver.rs:551-553, 566-568).
    A set of exactly the value both copies hold is a no-op (no launch); a different value, or an element whose device copy is
    newer than the host's, still goes through; +0.0 over -0.0 is a change."""
    L = capi.lib()
    F = C.c_float if dt == np.float32 else C.c_double
    host = np.arange(1, 9, dtype=dt)
    buf = capi.Buf(host)
    v = buf.view()
    out = F()
    get, set1 = capi.fn("tb_get1", dt), capi.fn("tb_set1", dt)
    capi.check(capi.fn("tb_scale", dt)(-3.0, v))                         # device-newer: [-3, -6, ...]
    capi.check(get(v, 2, C.byref(out))); assert out.value == -9         # round trip; both copies of element 2 agree now

    def launches():
        capi.check(L.tb_flush())
        n = C.c_uint64(); capi.check(L.tb_launch_count(C.byref(n)))
        return n.value
    l0 = launches()
    capi.check(set1(v, 2, -9.0))                                         # kappa := min(kappa, 0) with kappa < 0
    assert launches() == l0
    capi.check(set1(v, 2, 0.0))                                          # a real change
    l1 = launches()
    assert l1 == l0 + 1
    capi.check(get(v, 2, C.byref(out))); assert out.value == 0 and launches() == l1      # write-through: no round trip
    capi.check(set1(v, 2, -0.0))                                         # bitwise different from +0.0
    assert launches() == l1 + 1
    capi.check(set1(v, 3, -12.0))                                        # equal to the device value, but the host copy is stale: not skipped
    assert launches() == l1 + 2
    buf.release()
    want = np.arange(1, 9, dtype=np.float64) * -3
    want[2] = -0.0
    assert np.array_equal(host, want.astype(dt)) and np.signbit(host[2])


def test_length_mismatch_is_an_error():
    """f64lapack.rs:27,39,63-64 assert; the C ABI returns TB_ERR_ARG and the shim asserts."""
    a = capi.Buf(np.zeros(4, dtype=np.float32)); b = capi.Buf(np.zeros(5, dtype=np.float32))
    L = capi.lib()
    assert L.tb_copy_f32(a.view(), b.view()) == 2
    assert L.tb_add_f32(1.0, a.view(), b.view()) == 2
    assert L.tb_copy_f64(a.view(), a.view()) == 2      # dtype mismatch
    assert L.tb_scale_f32(1.0, capi.View(a.h, 2, 10)) == 2   # out of range
    a.release(); b.release()


@pytest.mark.parametrize("dt", DTYPES)
def test_recip_clamp(dt):
    host = np.array([0.0, 1e-20, 0.5, 4.0, -3.0], dtype=dt)
    buf = capi.Buf(host)
    capi.check(capi.fn("tb_recip_clamp", dt)(1e-12, buf.view()))
    buf.release()
    want = 1.0 / np.maximum(np.array([0.0, 1e-20, 0.5, 4.0, -3.0], dtype=dt), dt(1e-12))
    assert np.allclose(host, want, rtol=1e-6)


@pytest.mark.parametrize("dt", DTYPES)
def test_view_of_host_and_wrapper_refcount(dt):
    """The binding protocol of rust/totsu_b200 (B200Slice is the host sub-slice): sub-slices resolve to views by host
    address, every non-empty wrapper holds a reference on its root, and the root is flushed to the caller's slice and
    freed only when the last wrapper drops (slicelike.rs:18-19,37-46)."""
    L = capi.lib()
    work = np.arange(100, dtype=dt)
    other = np.zeros(10, dtype=dt)
    root = capi.Buf(work)                               # new_mut: refcount 1
    ob = capi.Buf(other)
    es = work.itemsize
    v = capi.View()
    capi.check(L.tb_view_of_host(capi.dtype_id(dt), C.c_void_p(work.ctypes.data + 20 * es), 30, C.byref(v)))
    assert (v.buf, v.off, v.len) == (root.h, 20, 30)
    capi.check(L.tb_view_of_host(capi.dtype_id(dt), C.c_void_p(other.ctypes.data), 10, C.byref(v)))
    assert (v.buf, v.off, v.len) == (ob.h, 0, 10)
    capi.check(L.tb_view_of_host(capi.dtype_id(dt), C.c_void_p(work.ctypes.data + 99 * es), 0, C.byref(v)))
    assert (v.buf, v.off, v.len) == (0, 0, 0)           # empty slice -> the empty view, accepted everywhere
    capi.check(capi.fn("tb_scale", dt)(2.0, v))
    foreign = np.zeros(4, dtype=dt)
    assert L.tb_view_of_host(capi.dtype_id(dt), C.c_void_p(foreign.ctypes.data), 4, C.byref(v)) != 0
    # split_mut(20) -> two non-empty children, device work on one child, children drop, then the root
    capi.check(L.tb_buf_retain(root.h, 2))
    capi.check(L.tb_view_of_host(capi.dtype_id(dt), C.c_void_p(work.ctypes.data + 20 * es), 80, C.byref(v)))
    capi.check(capi.fn("tb_scale", dt)(3.0, v))
    capi.check(L.tb_buf_release(root.h))                # child a
    capi.check(L.tb_buf_release(root.h))                # child b
    assert work[20] == 20                               # nothing flushed yet: the root wrapper is still alive
    size = C.c_size_t()
    capi.check(L.tb_buf_len(root.h, C.byref(size)))     # handle still valid
    root.release()                                      # last wrapper: flush + free
    assert np.array_equal(work[:20], np.arange(20, dtype=dt)) and np.array_equal(work[20:], 3 * np.arange(20, 100, dtype=dt))
    ob.release()
