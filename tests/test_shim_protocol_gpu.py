"""The Rust binding's call PROTOCOL, driven on the GPU (rust/totsu_b200/src/b200_slice.rs cannot be compiled in this
image): the binding's slice type is the host sub-slice itself, so every operand of every call is resolved with
tb_view_of_host, every split child takes a reference on its root (tb_buf_retain) and gives it back when it drops
(tb_view_of_host + tb_buf_release), and `ProductCone` holds one device-only PSD work buffer.  The C++ host mirror issues
exactly that sequence in "shim-protocol" mode (totsu_b200/host/linalg.hpp) and checks every lookup against the view it
carries.  Here: the protocol changes no number, and - because tb_view_of_host / tb_buf_retain / a refcount-only
tb_buf_release are pure bookkeeping - op/trans_op pairing, speculative pairing and PSD pairing still fire behind it."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
from helpers import capi, ZERO, RPOS, SOC, ROTSOC, PSD
from totsu_b200 import host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _init():
    capi.init(0)
    yield
    host.set_shim_protocol(False)


def _run(dt, blocks, n, seed, iters, protocol, fused=True, gemv_path=0):
    L = capi.lib()
    m = sum(l for _, l in blocks)
    a, b, c = H.make_instance(m, n, blocks, seed=seed, dtype=dt)
    abuf, av = H.device_matrix(a)
    host.set_shim_protocol(protocol)
    capi.check(L.tb_set_gemv_path(gemv_path))
    try:
        p0 = capi.pairs_fused()
        sp0 = [C.c_uint64() for _ in range(3)]
        capi.check(L.tb_spec_stats(*[C.byref(v) for v in sp0]))
        q0 = C.c_uint64(); capi.check(L.tb_psd_pairs(C.byref(q0)))
        s = host.Session.dense(dt, av, m, n, c, b, blocks, fused_op=fused, fused_cone=fused)
        assert s.begin(max_iter=None, eps_acc=0.0, eps_inf=0.0, device_precond=False) == "None"
        s.step(iters)
        xh, yh = s.xy()
        res = (s.last.c0, s.last.c1, s.last.c2)
        s.end()
        xs, ys = s.solution()
        s.close()
        sp1 = [C.c_uint64() for _ in range(3)]
        capi.check(L.tb_spec_stats(*[C.byref(v) for v in sp1]))
        q1 = C.c_uint64(); capi.check(L.tb_psd_pairs(C.byref(q1)))
        stats = {"pairs": capi.pairs_fused() - p0, "served": sp1[1].value - sp0[1].value, "psd_pairs": q1.value - q0.value}
    finally:
        host.set_shim_protocol(False)
        capi.check(L.tb_set_gemv_path(0))
        abuf.release()
    return xh, yh, res, xs, ys, stats


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_protocol_keeps_pair_fusion_and_speculation(dt):
    """Fused route (DenseOp + ProductCone) on a matrix the streaming kernel serves: with the binding's protocol every
    iteration still fuses its 3 op/trans_op pairs and the criteria_conv pair is still served from the speculated products
    (a draining tb_view_of_host made both impossible from Rust); iterates are bit-identical."""
    blocks, n = [(SOC, 64)] * 32 + [(RPOS, 512)], 1024
    iters = 25
    plain = _run(dt, blocks, n, 7, iters, False, gemv_path=2)
    proto = _run(dt, blocks, n, 7, iters, True, gemv_path=2)
    assert np.array_equal(plain[0], proto[0]) and np.array_equal(plain[1], proto[1]) and plain[2] == proto[2]
    assert np.array_equal(plain[3], proto[3]) and np.array_equal(plain[4], proto[4])      # the solution read out of `work`
    assert plain[5]["pairs"] == 3 * iters and proto[5]["pairs"] == 3 * iters
    assert proto[5]["served"] >= iters - 3 and proto[5]["served"] == plain[5]["served"]


def test_protocol_keeps_psd_pairing():
    """C4's shape in small (one ConePSD block, k = 64, f32): the two projections of an iteration still run as one batch
    (the Rust `ProductCone` used to wrap + release its work slice per `proj`, which un-parked the first one)."""
    blocks, n = [(PSD, 2080)], 40
    iters = 12
    plain = _run(np.float32, blocks, n, 3, iters, False)
    proto = _run(np.float32, blocks, n, 3, iters, True)
    assert np.array_equal(plain[0], proto[0]) and np.array_equal(plain[1], proto[1])
    assert plain[5]["psd_pairs"] == iters and proto[5]["psd_pairs"] == iters


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_protocol_on_stock_route(dt):
    """Stock route (MatOp + the reference's own cones: host loops, per-block splits, 1-element slices wrapped every
    iteration): the protocol's lookups, retains and releases resolve every operand to the right root - same numbers."""
    blocks, n = [(ROTSOC, 18), (SOC, 12), (SOC, 12), (RPOS, 40), (PSD, 21), (ZERO, 5)], 30
    plain = _run(dt, blocks, n, 11, 40, False, fused=False)
    proto = _run(dt, blocks, n, 11, 40, True, fused=False)
    assert np.array_equal(plain[0], proto[0]) and np.array_equal(plain[1], proto[1]) and plain[2] == proto[2]
    assert np.array_equal(plain[3], proto[3]) and np.array_equal(plain[4], proto[4])


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_protocol_front_end_qp(dt):
    """ProbQP through the protocol (the composite operators split x / y with splitm! on every call, qp.rs:98-140)."""
    n, m, p = 24, 20, 3
    rng = np.random.default_rng(4)
    g0 = rng.standard_normal((n, n))
    pm = g0 @ g0.T / n + 0.1 * np.eye(n)
    w, v = np.linalg.eigh(pm)
    psq = (v * np.sqrt(w)) @ v.T
    sym = np.array([psq[r, c] for c in range(n) for r in range(c + 1)]).astype(dt)
    gm = (rng.standard_normal((m, n)) / np.sqrt(n)).astype(dt)
    am = (rng.standard_normal((p, n)) / np.sqrt(n)).astype(dt)
    x0 = rng.standard_normal(n)
    h = (gm.astype(np.float64) @ x0 + np.abs(rng.standard_normal(m)) + 0.1).astype(dt)
    b = (am.astype(np.float64) @ x0).astype(dt)
    q = rng.standard_normal(n).astype(dt)
    out = {}
    try:
        for proto in (False, True):
            host.set_shim_protocol(proto)
            s = host.Session.qp(dt, sym, q, gm, h, am, b, 1e-12, p_is_sqrt=True)
            assert s.begin(max_iter=None, eps_acc=0.0, eps_inf=0.0) == "None"
            s.step(30)
            out[proto] = s.xy()
            s.close()
    finally:
        host.set_shim_protocol(False)
    assert np.array_equal(out[False][0], out[True][0]) and np.array_equal(out[False][1], out[True][1])


def test_wrapping_one_read_only_array_twice_shares_the_mirror():
    """ProbSOCP::problem wraps the same G_i twice at once (socp.rs:450,463): one mirror, refcounted."""
    L = capi.lib()
    arr = np.arange(64, dtype=np.float32)
    h1, h2 = C.c_int64(), C.c_int64()
    capi.check(L.tb_buf_wrap(capi.TB_F32, arr.ctypes.data_as(C.c_void_p), arr.size, 0, C.byref(h1)))
    capi.check(L.tb_buf_wrap(capi.TB_F32, arr.ctypes.data_as(C.c_void_p), arr.size, 0, C.byref(h2)))
    assert h1.value == h2.value
    v = capi.View()
    capi.check(L.tb_view_of_host(capi.TB_F32, C.c_void_p(arr.ctypes.data + 16), 8, C.byref(v)))
    assert (v.buf, v.off, v.len) == (h1.value, 4, 8)
    capi.check(L.tb_buf_release(h1.value))
    out = C.c_float()
    capi.check(L.tb_norm_f32(capi.View(h2.value, 0, 64), C.byref(out)))          # still alive after the first release
    assert abs(out.value - float(np.linalg.norm(arr))) < 1e-2
    capi.check(L.tb_buf_release(h2.value))
    assert L.tb_norm_f32(capi.View(h2.value, 0, 64), C.byref(out)) != 0            # gone after the second
    # a mutable wrap that partially overlaps a live one is refused (two device copies could not be kept coherent)
    big = np.zeros(32, dtype=np.float32)
    ha, hb = C.c_int64(), C.c_int64()
    capi.check(L.tb_buf_wrap(capi.TB_F32, big.ctypes.data_as(C.c_void_p), 16, 1, C.byref(ha)))
    assert L.tb_buf_wrap(capi.TB_F32, C.c_void_p(big.ctypes.data + 32), 16, 1, C.byref(hb)) != 0
    capi.check(L.tb_buf_release(ha.value))


def test_unsharded_denseop_must_cover_the_whole_matrix():
    L = capi.lib()
    buf = capi.Buf(dtype=np.float32, length=64 * 8)
    h = C.c_int64()
    assert L.tb_denseop_create(capi.TB_F32, buf.view(), 64, 8, 0, 128, C.byref(h)) != 0      # world == 1: no silent partial operator
    assert L.tb_denseop_create(capi.TB_F32, buf.view(), 64, 8, 8, 64, C.byref(h)) != 0
    buf.release()


def test_cone_proj_argument_errors_surface_at_the_call():
    """A projection that will be parked is validated first: TB_ERR_ARG comes back from tb_cone_proj itself (-> Err(()) ->
    ConeFailure, solver.rs:548-549), not from a later unrelated call."""
    L = capi.lib()
    blk = (capi.ConeBlock * 1)(capi.ConeBlock(PSD, 0, 2080))
    h = C.c_int64()
    capi.check(L.tb_cone_create(blk, 1, C.byref(h)))
    x = capi.Buf(dtype=np.float32, length=2080)
    short = capi.Buf(dtype=np.float32, length=100)
    assert L.tb_cone_proj_f32(h.value, 1, x.view(), 1e-12, short.view()) == 2       # TB_ERR_ARG: work shortage
    out = C.c_float()
    capi.check(L.tb_norm_f32(x.view(), C.byref(out)))                               # and nothing is left parked
    capi.check(L.tb_cone_destroy(h.value))
    x.release(); short.release()


def test_two_host_threads_share_the_backend():
    """The C ABI is serialised by one lock (cargo runs #[test]s on parallel threads): two threads solving at the
    same time get the answers a single thread gets."""
    import threading
    blocks, n = [(SOC, 16)] * 8 + [(ZERO, 8)], 48
    m = sum(l for _, l in blocks)
    a, b, c = H.make_instance(m, n, blocks, seed=2, dtype=np.float32)
    abuf, av = H.device_matrix(a)
    ref = None
    results, errors = {}, []

    def work(tag):
        try:
            s = host.Session.dense(np.float32, av, m, n, c, b, blocks, fused_op=True, fused_cone=True)
            assert s.begin(max_iter=None, eps_acc=0.0, eps_inf=0.0) == "None"
            for _ in range(60):
                s.step(1)
            results[tag] = s.xy()
            s.close()
        except Exception as e:       # noqa: BLE001
            errors.append(e)
    work("ref")
    ref = results["ref"]
    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    abuf.release()
    assert not errors, errors
    for i in range(2):      # interleaving changes which scalars ride on which round trip (prefetch.cu): equal to rounding, not bit for bit
        assert np.allclose(results[i][0], ref[0], rtol=1e-5, atol=1e-7) and np.allclose(results[i][1], ref[1], rtol=1e-5, atol=1e-7)
