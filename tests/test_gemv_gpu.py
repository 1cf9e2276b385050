"""GPU parity: LinAlgEx::transform_ge / transform_sp and the fused DenseOp, through the C ABI, against the oracle
(F64LAPACK.transform_ge / transform_sp = dgemv / dspmv semantics) on the same seeded inputs, on both kernel
paths (generic LDG and TMA streaming), plus size-independent properties at the full BASELINE sizes."""
import ctypes as C

import numpy as np
import pytest

from helpers import O, capi, synth, rel_linf, device_matrix

pytestmark = pytest.mark.gpu

DTYPES = [np.float32, np.float64]


@pytest.fixture(scope="module", autouse=True)
def _init():
    capi.init(0)
    yield
    capi.check(capi.lib().tb_set_gemv_path(0))


def _tol(dt, k):
    # |err| <= tol * sum_k |a||x| bound; accumulation in the working precision over k terms, split 4..64 ways
    return (3e-6 if dt == np.float32 else 1e-13)


def _run_ge(dt, m, n, transpose, alpha, beta, path, seed=0, nan_y=False):
    rng = np.random.default_rng(seed)
    L = capi.lib()
    a = np.asfortranarray(rng.standard_normal((m, n)).astype(dt))
    xl, yl = (m, n) if transpose else (n, m)
    # vectors live at odd offsets inside one work buffer, like the solver's sub-slices
    work = rng.standard_normal(3 + xl + 5 + yl + 2).astype(dt)
    if nan_y:
        work[3 + xl + 5: 3 + xl + 5 + yl] = np.nan
    ref = work.astype(np.float64).copy()
    wbuf = capi.Buf(work)
    abuf = capi.Buf(np.ascontiguousarray(a.reshape(-1, order="F")), mutable=False)
    xv, yv = wbuf.view(3, xl), wbuf.view(3 + xl + 5, yl)
    capi.check(L.tb_set_gemv_path(path))
    capi.check(capi.fn("tb_transform_ge", dt)(1 if transpose else 0, m, n, alpha, abuf.view(), xv, beta, yv))
    capi.check(L.tb_set_gemv_path(0))
    wbuf.release(); abuf.release()
    rx, ry = ref[3:3 + xl], ref[3 + xl + 5: 3 + xl + 5 + yl]
    O.F64LAPACK.transform_ge(transpose, m, n, alpha, a.astype(np.float64).reshape(-1, order="F"), rx, beta, ry)
    a64 = np.abs(a.astype(np.float64))
    bound = (a64.T @ np.abs(rx) if transpose else a64 @ np.abs(rx)) * abs(alpha) + np.abs(ry) + 1e-30
    got = work[3 + xl + 5: 3 + xl + 5 + yl].astype(np.float64)
    err = np.abs(got - ry) / bound
    assert np.all(np.isfinite(got))
    assert err.max() <= _tol(dt, xl), (m, n, transpose, path, err.max())
    # neighbours of y untouched
    assert work[3 + xl + 4] == ref[3 + xl + 4].astype(dt) and work[-1] == ref[-1].astype(dt)


GENERIC_SHAPES = [(1, 1), (3, 2), (80, 61), (63, 1000), (1, 777), (777, 1), (16384, 1), (1, 16384), (257, 129), (1000, 33)]
STREAM_SHAPES = [(256, 16), (1024, 64), (2048, 512), (1028, 100), (4100, 37), (512, 4099), (8192, 4096), (3076, 1030)]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", GENERIC_SHAPES)
@pytest.mark.parametrize("transpose", [False, True])
def test_transform_ge_generic(dt, shape, transpose):
    _run_ge(dt, shape[0], shape[1], transpose, 1.0, 0.0, path=1, seed=1, nan_y=True)
    _run_ge(dt, shape[0], shape[1], transpose, -0.7, 1.0, path=1, seed=2)
    _run_ge(dt, shape[0], shape[1], transpose, 0.3, -1.25, path=1, seed=3)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", STREAM_SHAPES)
@pytest.mark.parametrize("transpose", [False, True])
def test_transform_ge_tma_stream(dt, shape, transpose):
    _run_ge(dt, shape[0], shape[1], transpose, 1.0, 0.0, path=2, seed=4, nan_y=True)
    _run_ge(dt, shape[0], shape[1], transpose, -0.7, 1.0, path=2, seed=5)


@pytest.mark.parametrize("dt", DTYPES)
def test_transform_ge_auto_path_and_errors(dt):
    _run_ge(dt, 2048, 1024, False, 1.0, 0.0, path=0, seed=6)
    _run_ge(dt, 2048, 1024, True, 1.0, 0.0, path=0, seed=7)
    a = capi.Buf(np.zeros(12, dtype=dt), mutable=False)
    x = capi.Buf(np.zeros(4, dtype=dt)); y = capi.Buf(np.zeros(3, dtype=dt))
    f = capi.fn("tb_transform_ge", dt)
    assert f(0, 3, 4, 1.0, a.view(), x.view(), 0.0, y.view()) == 0
    assert f(1, 3, 4, 1.0, a.view(), x.view(), 0.0, y.view()) == 2      # f64lapack.rs:128-129
    assert f(0, 4, 4, 1.0, a.view(), x.view(), 0.0, y.view()) == 2      # f64lapack.rs:125
    for b in (a, x, y):
        b.release()


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("n", [1, 2, 5, 16, 17, 100, 511, 512, 513, 1024, 1025, 2050, 4099, 6000])
@pytest.mark.parametrize("warps", [8, 16])
def test_transform_sp(dt, n, warps):
    """n < 512 (and unaligned views): generic kernel; n >= 512: the TMA streaming kernel (sizes chosen so that the
    packed length leaves 0..3 trailing elements outside the last aligned 16-byte window, and chunks are ragged), with 8
    consumer warps (the default) and with 16 (two per column in the column pass).  Work units are dealt to the CTAs by a
    device-side queue; partials are indexed by unit, so the result does not depend on the dealing."""
    if n < 512 and warps == 16:
        pytest.skip("the generic kernel has no warp variants")
    capi.check(capi.lib().tb_set_spmv_warps(warps))
    try:
        _transform_sp_case(dt, n)
    finally:
        capi.check(capi.lib().tb_set_spmv_warps(8))


def _transform_sp_case(dt, n):
    rng = np.random.default_rng(n)
    sp = rng.standard_normal(n * (n + 1) // 2).astype(dt)
    x = rng.standard_normal(n).astype(dt)
    y = rng.standard_normal(n).astype(dt)
    for alpha, beta in [(1.0, 0.0), (-0.5, 1.0), (2.0, 0.25)]:
        yy = y.copy()
        if beta == 0.0:
            yy[:] = np.nan
        sb, xb, yb = capi.Buf(sp.copy(), mutable=False), capi.Buf(x.copy(), mutable=False), capi.Buf(yy)
        capi.check(capi.fn("tb_transform_sp", dt)(n, alpha, sb.view(), xb.view(), beta, yb.view()))
        for b in (sb, xb, yb):
            b.release()
        ry = y.astype(np.float64).copy()
        O.F64LAPACK.transform_sp(n, alpha, sp.astype(np.float64), x.astype(np.float64), beta, ry)
        scale = np.abs(ry).max() + np.abs(sp).max() * np.abs(x).sum()
        assert np.abs(yy - ry).max() <= (3e-6 if dt == np.float32 else 1e-13) * scale, (n, alpha, beta)


@pytest.mark.parametrize("dt", DTYPES)
def test_matop_sympack_unit_vectors(dt):
    """totsu_core/src/matop.rs:179-212 through the device transform_sp."""
    array = np.arange(1, 16, dtype=dt)
    ref = np.array([[1, 2, 4, 7, 11], [2, 3, 5, 8, 12], [4, 5, 6, 9, 13], [7, 8, 9, 10, 14], [11, 12, 13, 14, 15]], dtype=np.float64)
    sb = capi.Buf(array, mutable=False)
    for i in range(5):
        x = np.zeros(5, dtype=dt); x[i] = 1
        y = np.zeros(5, dtype=dt)
        xb, yb = capi.Buf(x, mutable=False), capi.Buf(y)
        capi.check(capi.fn("tb_transform_sp", dt)(5, 1.0, sb.view(), xb.view(), 0.0, yb.view()))
        xb.release(); yb.release()
        assert np.allclose(y, ref[i], atol=1e-3)
    sb.release()


@pytest.mark.parametrize("dt", DTYPES)
def test_fill_uniform_matches_numpy_twin_bit_exactly(dt):
    m, n, off, seed = 1000, 37, 123, 7
    scale = dt(1.0 / np.sqrt(n))
    buf = capi.Buf(dtype=dt, length=m * n)
    capi.check(capi.fn("tb_fill_uniform", dt)(buf.view(), m, n, off, seed, scale))
    got = buf.download().reshape((m, n), order="F")
    want = synth.uniform_matrix(m, n, seed, scale, row_offset=off, dtype=dt)
    buf.release()
    assert np.array_equal(got, want)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(130, 70), (2048, 640), (5000, 333)])
def test_denseop_apply_pair_absadd(dt, shape):
    m, n = shape
    L = capi.lib()
    rng = np.random.default_rng(m)
    a = np.asfortranarray(rng.standard_normal((m, n)).astype(dt))
    a64 = a.astype(np.float64)
    abuf, av = device_matrix(a)
    h = C.c_int64()
    capi.check(L.tb_denseop_create(capi.dtype_id(dt), av, m, n, 0, 0, C.byref(h)))
    xn, xt = rng.standard_normal(n).astype(dt), rng.standard_normal(m).astype(dt)
    yn, yt = rng.standard_normal(m).astype(dt), rng.standard_normal(n).astype(dt)
    tol = 3e-6 if dt == np.float32 else 1e-13
    for path in (1, 2):
        capi.check(L.tb_set_gemv_path(path))
        bxn, bxt = capi.Buf(xn.copy(), mutable=False), capi.Buf(xt.copy(), mutable=False)
        gyn, gyt = yn.copy(), yt.copy()
        byn, byt = capi.Buf(gyn), capi.Buf(gyt)
        capi.check(capi.fn("tb_denseop_apply_pair", dt)(h.value, 0.5, bxn.view(), 1.0, byn.view(), -2.0, bxt.view(), 0.0, byt.view()))
        # single ops on top
        capi.check(capi.fn("tb_denseop_apply", dt)(h.value, 0, 1.0, bxn.view(), 1.0, byn.view()))
        capi.check(capi.fn("tb_denseop_apply", dt)(h.value, 1, 1.0, bxt.view(), 1.0, byt.view()))
        for b in (bxn, bxt, byn, byt):
            b.release()
        wn = 1.5 * (a64 @ xn) + yn
        wt = -1.0 * (a64.T @ xt)
        bn = 1.5 * (np.abs(a64) @ np.abs(xn)) + np.abs(yn)
        bt = 3.0 * (np.abs(a64).T @ np.abs(xt))
        assert (np.abs(gyn - wn) / bn).max() <= tol, path
        assert (np.abs(gyt - wt) / bt).max() <= tol, path
    capi.check(L.tb_set_gemv_path(0))
    tau, sig = np.ones(n, dtype=dt), np.ones(m, dtype=dt)
    bt_, bs_ = capi.Buf(tau), capi.Buf(sig)
    capi.check(capi.fn("tb_denseop_absadd_cols", dt)(h.value, bt_.view()))
    capi.check(capi.fn("tb_denseop_absadd_rows", dt)(h.value, bs_.view()))
    bt_.release(); bs_.release()
    assert rel_linf(tau, 1 + np.abs(a64).sum(0)) <= 10 * tol
    assert rel_linf(sig, 1 + np.abs(a64).sum(1)) <= 10 * tol
    capi.check(L.tb_denseop_destroy(h.value))
    abuf.release()

@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("m", [7, 4097, 70000])
def test_column_operator_on_a_host_written_scalar(dt, m):
    """An m x 1 operator (the solver's b and c) applied to a 1-element slice the host has just written (`work_one = [1]`,
    solver.rs:590-596): the scalar travels by value - no upload kernel - and the result is bit-identical to the same apply with
    the scalar on the device; both through tb_denseop_apply and through MatOp's tb_transform_ge."""
    L = capi.lib()
    rng = np.random.default_rng(m)
    col = rng.standard_normal(m).astype(dt)
    y0 = rng.standard_normal(m).astype(dt)
    cbuf, cv = device_matrix(np.asfortranarray(col.reshape(m, 1)))
    h = C.c_int64()
    capi.check(L.tb_denseop_create(capi.dtype_id(dt), cv, m, 1, 0, 0, C.byref(h)))
    alpha, beta, sval = dt(-0.75), dt(0.5), dt(1.25)
    want = (alpha * (col * sval) + beta * y0).astype(dt)                      # the arithmetic of both forms

    def launches():
        capi.check(L.tb_flush())
        n = C.c_uint64(); capi.check(L.tb_launch_count(C.byref(n)))
        return n.value
    res = {}
    for where in ("host", "device"):
        for api in ("denseop", "transform_ge"):
            s1 = np.array([sval], dtype=dt)
            sb = capi.Buf(s1)
            if where == "device":
                capi.check(capi.fn("tb_scale", dt)(1.0, sb.view()))           # the device copy is now the newer one
            y = y0.copy()
            yb = capi.Buf(y)
            capi.check(capi.fn("tb_scale", dt)(1.0, yb.view()))               # y on the device already
            l0 = launches()
            if api == "denseop":
                capi.check(capi.fn("tb_denseop_apply", dt)(h.value, 0, alpha, sb.view(), beta, yb.view()))
            else:
                capi.check(capi.fn("tb_transform_ge", dt)(0, m, 1, alpha, cv, sb.view(), beta, yb.view()))
            n_l = launches() - l0
            yb.release(); sb.release()
            res[(where, api)] = (y, n_l)
            assert rel_linf(y, want.astype(np.float64)) <= (1e-6 if dt == np.float32 else 1e-15), (where, api)   # the device contracts to FMA
    for api in ("denseop", "transform_ge"):
        assert np.array_equal(res[("host", api)][0], res[("device", api)][0]), api        # same arithmetic, bit for bit
    # by value: one launch (the program with the apply); from a host-newer device scalar the upload kernel would come on top
    assert res[("host", "denseop")][1] == 1 and res[("host", "transform_ge")][1] == 1, {k: v[1] for k, v in res.items()}
    capi.check(L.tb_denseop_destroy(h.value))
    cbuf.release()


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(2048, 640), (5000, 333), (1028, 2050)])
@pytest.mark.parametrize("rows_first", [False, True])
def test_one_pass_absadd_and_invalidation(dt, shape, rows_first):
    """absadd_cols + absadd_rows of a streaming-size operator come from ONE read of A (stream_kernel<T,1,1,ABS>): whichever
    is asked for first computes both sums of |A|, the other is served from the kept sums (matop.rs:98-138 semantics:
    tau[c] += sum_r |A[r,c]|, sigma[r] += sum_c |A[r,c]|); a write into the matrix drops the kept sums."""
    m, n = shape
    L = capi.lib()
    rng = np.random.default_rng(m + n)
    a = np.asfortranarray(rng.standard_normal((m, n)).astype(dt))
    abuf, av = device_matrix(a)
    h = C.c_int64()
    capi.check(L.tb_denseop_create(capi.dtype_id(dt), av, m, n, 0, 0, C.byref(h)))
    capi.check(L.tb_set_gemv_path(2))
    tol = 3e-5 if dt == np.float32 else 1e-12

    def both(mat):
        tau, sig = rng.standard_normal(n).astype(dt), rng.standard_normal(m).astype(dt)
        t0, s0 = tau.astype(np.float64), sig.astype(np.float64)
        bt_, bs_ = capi.Buf(tau), capi.Buf(sig)
        l0 = capi.launch_count()
        calls = [("tb_denseop_absadd_rows", bs_), ("tb_denseop_absadd_cols", bt_)] if rows_first else [("tb_denseop_absadd_cols", bt_), ("tb_denseop_absadd_rows", bs_)]
        for name, bf in calls:
            capi.check(capi.fn(name, dt)(h.value, bf.view()))
        capi.check(L.tb_flush())
        launches = capi.launch_count() - l0
        bt_.release(); bs_.release()
        a64 = mat.astype(np.float64)
        assert rel_linf(tau, t0 + np.abs(a64).sum(0)) <= tol
        assert rel_linf(sig, s0 + np.abs(a64).sum(1)) <= tol
        return launches
    try:
        first = both(a)
        again = both(a)                       # served from the kept sums: no pass over A at all
        assert again < first
        a2 = np.asfortranarray(rng.standard_normal((m, n)).astype(dt))
        abuf.upload(a2.reshape(-1, order="F"))   # the matrix changes: the kept sums must not be reused
        both(a2)
    finally:
        capi.check(L.tb_set_gemv_path(0))
        capi.check(L.tb_denseop_destroy(h.value))
        abuf.release()


def test_full_size_properties_c3():
    """BASELINE config C3 (A 65536 x 16384, f32, generated in HBM): properties that need no CPU copy of A -
    adjoint identity <A x, y> = <x, A^T y>, linearity in x, fused pair == separate calls, run-to-run bit
    reproducibility, and a spot check of generated rows against the numpy twin."""
    dt = np.float32
    m, n = 65536, 16384
    L = capi.lib()
    rng = np.random.default_rng(3)
    abuf = capi.Buf(dtype=dt, length=m * n)
    scale = dt(1.0 / np.sqrt(n))
    capi.check(L.tb_fill_uniform_f32(abuf.view(), m, n, 0, 11, scale))
    h = C.c_int64()
    capi.check(L.tb_denseop_create(capi.TB_F32, abuf.view(), m, n, 0, 0, C.byref(h)))
    x1, x2 = rng.standard_normal(n).astype(dt), rng.standard_normal(n).astype(dt)
    yv = rng.standard_normal(m).astype(dt)

    def op(x):
        xb = capi.Buf(x.copy(), mutable=False); out = np.zeros(m, dtype=dt); ob = capi.Buf(out)
        capi.check(L.tb_denseop_apply_f32(h.value, 0, 1.0, xb.view(), 0.0, ob.view()))
        xb.release(); ob.release()
        return out

    def top(y):
        yb = capi.Buf(y.copy(), mutable=False); out = np.zeros(n, dtype=dt); ob = capi.Buf(out)
        capi.check(L.tb_denseop_apply_f32(h.value, 1, 1.0, yb.view(), 0.0, ob.view()))
        yb.release(); ob.release()
        return out

    ax1, ax2, aty = op(x1), op(x2), top(yv)
    assert np.array_equal(ax1, op(x1)) and np.array_equal(aty, top(yv))           # bit-reproducible
    lhs, rhs = np.dot(ax1.astype(np.float64), yv.astype(np.float64)), np.dot(x1.astype(np.float64), aty.astype(np.float64))
    assert abs(lhs - rhs) <= 1e-4 * (np.linalg.norm(ax1) * np.linalg.norm(yv))     # adjoint identity
    lin = op((x1 + 2 * x2).astype(dt))
    assert rel_linf(lin, ax1.astype(np.float64) + 2 * ax2.astype(np.float64)) <= 2e-5
    # fused pair equals the two separate passes bit for bit (same kernels' per-element summation order)
    xb, yb = capi.Buf(x1.copy(), mutable=False), capi.Buf(yv.copy(), mutable=False)
    pn, pt = np.zeros(m, dtype=dt), np.zeros(n, dtype=dt)
    bn, bt = capi.Buf(pn), capi.Buf(pt)
    capi.check(L.tb_denseop_apply_pair_f32(h.value, 1.0, xb.view(), 0.0, bn.view(), 1.0, yb.view(), 0.0, bt.view()))
    for b in (xb, yb, bn, bt):
        b.release()
    assert np.array_equal(pn, ax1) and np.array_equal(pt, aty)
    # spot check: rows of A against the numpy twin through e_j probes would be slow; check (A x)[rows] instead
    rows = np.array([0, 1, 1023, 1024, 40000, 65535])
    sub = synth.uniform_matrix(len(rows), n, 11, scale, dtype=dt, rows=rows).astype(np.float64)
    assert rel_linf(ax1[rows], sub @ x1.astype(np.float64)) <= 2e-5
    capi.check(L.tb_denseop_destroy(h.value))
    abuf.release()


def test_full_size_properties_c5_single_gpu():
    """BASELINE config C5's matrix (m = 262144, n = 65536, f32 = 68.7 GB) on ONE B200 (it fits the 180 GB of HBM;
    the 8-GPU row-sharded form is tests/dist_c5_worker.py): adjoint identity, the fused pair equals the separate
    passes bit for bit, run-to-run reproducibility, and sampled rows / columns against the numpy twin of the
    generator - the oracle cannot hold this matrix (137 GB in f64)."""
    import torch
    free, total = torch.cuda.mem_get_info(0)
    dt = np.float32
    m, n = 262144, 65536
    if free < m * n * 4 + (8 << 30):
        pytest.skip("needs 77 GB of free HBM")
    L = capi.lib()
    rng = np.random.default_rng(5)
    abuf = capi.Buf(dtype=dt, length=m * n)
    scale = dt(1.0 / np.sqrt(n))
    capi.check(L.tb_fill_uniform_f32(abuf.view(), m, n, 0, 5, scale))
    h = C.c_int64()
    capi.check(L.tb_denseop_create(capi.TB_F32, abuf.view(), m, n, 0, 0, C.byref(h)))
    x = rng.standard_normal(n).astype(dt); u = rng.standard_normal(m).astype(dt)
    xb, ub = capi.Buf(x.copy(), mutable=False), capi.Buf(u.copy(), mutable=False)
    ax, atu, ax2, atu2 = (np.zeros(m, dt), np.zeros(n, dt), np.zeros(m, dt), np.zeros(n, dt))
    b1, b2, b3, b4 = capi.Buf(ax), capi.Buf(atu), capi.Buf(ax2), capi.Buf(atu2)
    capi.check(L.tb_set_pair_fusion(0))
    try:
        capi.check(L.tb_denseop_apply_f32(h.value, 0, 1.0, xb.view(), 0.0, b1.view()))
        capi.check(L.tb_denseop_apply_f32(h.value, 1, 1.0, ub.view(), 0.0, b2.view()))
    finally:
        capi.check(L.tb_set_pair_fusion(1))
    p0 = capi.pairs_fused()
    capi.check(L.tb_denseop_apply_f32(h.value, 0, 1.0, xb.view(), 0.0, b3.view()))
    capi.check(L.tb_denseop_apply_f32(h.value, 1, 1.0, ub.view(), 0.0, b4.view()))
    for b in (b1, b2, b3, b4, xb, ub):
        b.release()
    assert capi.pairs_fused() == p0 + 1
    assert np.array_equal(ax, ax2) and np.array_equal(atu, atu2)
    lhs, rhs = np.dot(ax.astype(np.float64), u.astype(np.float64)), np.dot(x.astype(np.float64), atu.astype(np.float64))
    assert abs(lhs - rhs) <= 1e-4 * (np.linalg.norm(ax) * np.linalg.norm(u))
    rows = np.array([0, 1, 1023, 1024, 131071, 131072, 200001, 262143])
    sub = synth.uniform_matrix(len(rows), n, 5, scale, dtype=dt, rows=rows).astype(np.float64)
    assert rel_linf(ax[rows], sub @ x.astype(np.float64)) <= 5e-5
    cols = np.array([0, 7, 8, 32767, 65535])
    subc = synth.uniform_matrix(m, len(cols), 5, scale, dtype=dt, cols=cols).astype(np.float64)
    assert rel_linf(atu[cols], subc.T @ u.astype(np.float64)) <= 5e-5
    capi.check(L.tb_denseop_destroy(h.value))
    abuf.release()


@pytest.mark.parametrize("dt", DTYPES)
def test_transform_sp_unaligned_view_and_engines_agree(dt):
    """The packed matrix as a sub-view that is not 16-byte aligned falls back to the generic kernel; both kernels agree
    with each other (tb_set_gemv_path(1) forces the generic one)."""
    n = 1500
    rng = np.random.default_rng(3)
    sp = rng.standard_normal(n * (n + 1) // 2 + 1).astype(dt)
    x = rng.standard_normal(n).astype(dt)
    res = []
    for off, path in ((0, 0), (1, 0), (0, 1)):
        capi.check(capi.lib().tb_set_gemv_path(path))
        y = np.zeros(n, dtype=dt)
        sb, xb, yb = capi.Buf(sp.copy(), mutable=False), capi.Buf(x.copy(), mutable=False), capi.Buf(y)
        capi.check(capi.fn("tb_transform_sp", dt)(n, 1.0, sb.view(off, n * (n + 1) // 2), xb.view(), 0.0, yb.view()))
        for b in (sb, xb, yb):
            b.release()
        ry = np.zeros(n)
        O.F64LAPACK.transform_sp(n, 1.0, sp[off:off + n * (n + 1) // 2].astype(np.float64), x.astype(np.float64), 0.0, ry)
        scale = np.abs(ry).max() + np.abs(sp).max() * np.abs(x).sum()
        assert np.abs(y - ry).max() <= (3e-6 if dt == np.float32 else 1e-13) * scale, (off, path)
        res.append(y)
    capi.check(capi.lib().tb_set_gemv_path(0))
    assert np.abs(res[0].astype(np.float64) - res[2].astype(np.float64)).max() <= (1e-5 if dt == np.float32 else 1e-12) * np.abs(res[0]).max()
