"""GPU parity: batched product-cone projection (tb_cone_*), LinAlgEx::map_eig and the ConePSD projection,
through the C ABI, against the oracle's cones (oracle/totsu_oracle.py = cone_*.rs / f64lapack.rs:78-255)."""
import ctypes as C
import math

import numpy as np
import pytest

from helpers import O, capi, rel_linf, oracle_cone, svec, ZERO, RPOS, SOC, ROTSOC, PSD

pytestmark = pytest.mark.gpu
DTYPES = [np.float32, np.float64]


@pytest.fixture(scope="module", autouse=True)
def _init():
    capi.init(0)
    yield


def _cone(blocks):
    blk = (capi.ConeBlock * len(blocks))(*[capi.ConeBlock(t, 0, ln) for t, ln in blocks])
    h = C.c_int64()
    capi.check(capi.lib().tb_cone_create(blk, len(blocks), C.byref(h)))
    return h.value


def _psd_worklen(blocks):
    w = 0
    for t, ln in blocks:
        if t == PSD:
            k = int((math.sqrt(8 * ln + 1) - 1) / 2 + 0.5)
            w = max(w, 2 * k * k + k)
    return w


CASES = [
    [(SOC, 64)] * 40 + [(ZERO, 5)],
    [(ROTSOC, 10), (RPOS, 20), (ZERO, 4), (SOC, 1), (ROTSOC, 1), (SOC, 2), (ROTSOC, 2), (SOC, 0), (RPOS, 0)],
    [(RPOS, 10000), (ZERO, 333)],
    [(ROTSOC, 8194), (RPOS, 500), (ZERO, 100)],
    [(SOC, 5000), (SOC, 3), (SOC, 2049)],
    [(PSD, 3), (SOC, 5), (PSD, 55), (ZERO, 2)],
]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("dual", [False, True])
def test_product_cone_proj(dt, case, dual):
    blocks = CASES[case]
    m = sum(l for _, l in blocks)
    rng = np.random.default_rng(case)
    tol = 5e-6 if dt == np.float32 else 1e-12
    # three regimes per SOC block: inside, in the polar (-> 0), and the generic case
    for trial in range(3):
        x = rng.standard_normal(m).astype(dt)
        o = 0
        for t, ln in blocks:
            if t in (SOC, ROTSOC) and ln > 1 and trial < 2:
                if t == SOC:
                    x[o] = (1 if trial == 0 else -1) * (np.linalg.norm(x[o + 1:o + ln]) * 2 + 1)
                else:
                    big = np.linalg.norm(x[o + 2:o + ln]) * 2 + 1
                    x[o] = x[o + 1] = (1 if trial == 0 else -1) * big
            o += ln
        want = x.astype(np.float64).copy()
        oracle_cone(blocks).proj(dual, want)
        got = x.copy()
        h = _cone(blocks)
        xb = capi.Buf(got)
        wl = _psd_worklen(blocks)
        wb = capi.Buf(np.zeros(max(wl, 1), dtype=dt))
        capi.check(capi.fn("tb_cone_proj", dt)(h, 1 if dual else 0, xb.view(), 1e-12, wb.view() if wl else capi.View(0, 0, 0)))
        xb.release(); wb.release()
        capi.check(capi.lib().tb_cone_destroy(h))
        has_psd = any(t == PSD for t, _ in blocks)
        assert rel_linf(got, want) <= (tol * (50 if has_psd else 1)), (case, trial)
        if trial == 0 and not has_psd:
            # points inside a self-dual cone are fixed points (up to the rotation round trip for RotSOC)
            o = 0
            for t, ln in blocks:
                if t == SOC and ln > 1:
                    assert np.array_equal(got[o:o + ln], x[o:o + ln])
                o += ln


@pytest.mark.parametrize("dt", DTYPES)
def test_product_group_min(dt):
    blocks = CASES[1] + [(SOC, 3000), (PSD, 6)]
    m = sum(l for _, l in blocks)
    rng = np.random.default_rng(0)
    t = (np.abs(rng.standard_normal(m)) + 0.1).astype(dt)
    want = t.astype(np.float64).copy()

    def group(g):
        if g.size > 0:
            g[...] = g.min()
    oracle_cone(blocks).product_group(want, group)
    h = _cone(blocks)
    tb = capi.Buf(t)
    capi.check(capi.fn("tb_cone_group_min", dt)(h, tb.view()))
    tb.release()
    capi.check(capi.lib().tb_cone_destroy(h))
    assert np.array_equal(t.astype(np.float64), want.astype(dt).astype(np.float64))


@pytest.mark.parametrize("dt", DTYPES)
def test_cone_psd_unit(dt):
    """totsu_core/src/cone_psd.rs:89-110: proj of diag(5,-5) -> diag(5,0)."""
    x = np.array([5., 0., -5.], dtype=dt)
    xb, wb = capi.Buf(x), capi.Buf(np.zeros(10, dtype=dt))
    capi.check(capi.fn("tb_proj_psd", dt)(xb.view(), 1e-12, wb.view()))
    xb.release(); wb.release()
    assert np.allclose(x, [5., 0., 0.], atol=1e-6)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("k", [1, 2, 3, 7, 16, 33, 64, 129])
@pytest.mark.parametrize("path", [0, 1])
def test_proj_psd_vs_oracle(dt, k, path, request):
    """path 0: GEMM-only matrix-sign iteration (the default), path 1: Jacobi eigendecomposition; both against the
    oracle's dsyevr + dsyr restatement of f64lapack.rs:78-108."""
    capi.check(capi.lib().tb_set_psd_path(path))
    request.addfinalizer(lambda: capi.check(capi.lib().tb_set_psd_path(0)))
    rng = np.random.default_rng(k)
    g = rng.standard_normal((k, k))
    x = svec((g + g.T) / 2).astype(dt)
    want = x.astype(np.float64).copy()
    O.ConePSD(np.zeros(O.ConePSD.query_worklen(x.size)), 1e-12).proj(False, want)
    xb, wb = capi.Buf(x), capi.Buf(np.zeros(2 * k * k + k, dtype=dt))
    capi.check(capi.fn("tb_proj_psd", dt)(xb.view(), 1e-12, wb.view()))
    xb.release(); wb.release()
    tol = 2e-5 if dt == np.float32 else 1e-11
    assert np.abs(x - want).max() <= tol * max(1.0, np.abs(want).max()), k
    # idempotence and PSD-ness of the result
    again = x.copy()
    xb, wb = capi.Buf(again), capi.Buf(np.zeros(2 * k * k + k, dtype=dt))
    capi.check(capi.fn("tb_proj_psd", dt)(xb.view(), 1e-12, wb.view()))
    xb.release(); wb.release()
    assert np.abs(again - x).max() <= 5 * tol * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("k", [2, 5, 40])
def test_map_eig_sqrt_closure(dt, k):
    """MatBuild::set_sqrt (matbuild/mod.rs:220-241): map_eig with scale_diag=None and e -> sqrt(e)."""
    rng = np.random.default_rng(k)
    g = rng.standard_normal((k, k + 2))
    p = g @ g.T / k
    packed = np.array([p[r, c] for c in range(k) for r in range(c + 1)], dtype=dt)
    want = packed.astype(np.float64).copy()
    O.F64LAPACK.map_eig(want, None, 1e-12, np.zeros(O.F64LAPACK.map_eig_worklen(k)), lambda e: math.sqrt(e) if e > 0 else None)
    F = C.c_float if dt == np.float32 else C.c_double
    mb, wb = capi.Buf(packed), capi.Buf(np.zeros(2 * k * k + k, dtype=dt))
    eigs = (F * k)()
    capi.check(capi.fn("tb_map_eig_begin", dt)(mb.view(), 0, 1.0, 1e-12, wb.view(), eigs))
    ev = np.array(list(eigs), dtype=np.float64)
    ref_ev = np.linalg.eigvalsh(p)
    assert np.abs(np.sort(ev) - ref_ev).max() <= (3e-5 if dt == np.float32 else 1e-11) * max(1.0, ref_ev.max())
    keep = (C.c_uint8 * k)(*[1 if e > 0 else 0 for e in ev])
    new = (F * k)(*[math.sqrt(e) if e > 0 else 0.0 for e in ev])
    capi.check(capi.fn("tb_map_eig_finish", dt)(mb.view(), 0, 1.0, wb.view(), new, keep))
    mb.release(); wb.release()
    assert np.abs(packed - want).max() <= (2e-3 if dt == np.float32 else 1e-10) * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("k,kind", [(5, "full"), (40, "full"), (64, "full"), (256, "full"), (512, "full"), (512, "lowrank"), (512, "diag"), (2048, "full"), (96, "indefinite")])
def test_sqrt_psd_vs_oracle(dt, k, kind):
    """tb_sqrt_psd = MatBuild::set_sqrt (matbuild/mod.rs:220-241): map_eig(scale_diag = None, e -> sqrt(e)) of an upper-packed
    symmetric PSD matrix, GEMM-only (coupled Newton-Schulz on the tcgen05 / FP64 engine) for k >= 32, against the oracle's
    dsyevr route (f64lapack.rs:78-108).  k = 2048 is 2 orders of magnitude beyond what the Jacobi closure path could do in round 1;
    `lowrank` (rank k/4) and `indefinite` exercise the fallback to the eigendecomposition."""
    if k >= 2048 and dt == np.float64:
        pytest.skip("f64 at k = 2048 runs on the FP64 pipe: minutes")
    rng = np.random.default_rng(k + len(kind))
    if kind == "full":
        g = rng.standard_normal((k, k + 8))
        p = g @ g.T / k + 0.05 * np.eye(k)
    elif kind == "lowrank":
        g = rng.standard_normal((k, k // 4))
        p = g @ g.T / k
    elif kind == "diag":
        p = np.diag(rng.uniform(0.05, 1.0, k))                      # bench.py's C2 P
    else:
        g = rng.standard_normal((k, k))
        p = (g + g.T) / 2                                           # not PSD: the closure drops the negative eigenvalues
    p = p.astype(dt).astype(np.float64)
    packed = np.array([p[r, c] for c in range(k) for r in range(c + 1)], dtype=dt) if k <= 512 else p.T[np.tril_indices(k)].astype(dt)
    w, v = np.linalg.eigh(p)
    want_m = (v * np.sqrt(np.maximum(w, 0.0))) @ v.T                  # same as dsyevr(V, (0, inf]) + dsyr with sqrt(e)
    if k <= 64:
        want_o = packed.astype(np.float64).copy()
        O.F64LAPACK.map_eig(want_o, None, 1e-12, np.zeros(O.F64LAPACK.map_eig_worklen(k)), lambda e: math.sqrt(e) if e > 0 else None)
        assert np.abs(want_o - want_m.T[np.tril_indices(k)]).max() <= 1e-9 * np.abs(want_m).max()
    mb, wb = capi.Buf(packed), capi.Buf(dtype=dt, length=2 * k * k + k)
    capi.check(capi.fn("tb_sqrt_psd", dt)(mb.view(), 1e-12, wb.view()))
    route, iters = C.c_int(), C.c_int()
    capi.check(capi.lib().tb_sqrt_psd_info(C.byref(route), C.byref(iters)))
    mb.release(); wb.release()
    got = np.zeros((k, k))
    got.T[np.tril_indices(k)] = packed
    got = np.triu(got) + np.triu(got, 1).T
    if kind in ("full", "diag") and k >= 32:
        assert route.value == 1 and 3 <= iters.value <= 40, (route.value, iters.value)      # the GEMM-only route took it
    if kind == "indefinite":
        assert route.value == 2
    scale = np.abs(want_m).max()
    # f32 rounding of the GEMM chain grows with k (4e-5 at 2048); sqrt is ill-conditioned at 0: compare S (to sqrt of the working precision for rank-deficient P) and S^2 = P+ (tight)
    tol_s = {"full": 3e-5 * max(1, k // 512), "diag": 3e-5, "lowrank": 3e-3, "indefinite": 3e-3}[kind] if dt == np.float32 else \
            {"full": 1e-11, "diag": 1e-11, "lowrank": 1e-5, "indefinite": 1e-8}[kind]        # lowrank: eigenvalues below eps_zero = 1e-12 are dropped, sqrt(1e-12) = 1e-6 each
    assert np.abs(got - want_m).max() <= tol_s * scale, (np.abs(got - want_m).max() / scale, route.value, iters.value)
    p_plus = (v * np.maximum(w, 0.0)) @ v.T
    tol_p = 1e-4 if dt == np.float32 else 1e-10
    assert np.abs(got @ got - p_plus).max() <= tol_p * np.abs(p_plus).max()


def _spectrum_matrix(k, lam, seed):
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((k, k)))
    x = (q * lam) @ q.T
    return (x + x.T) / 2


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("name", ["logspaced", "lowrank", "psd_already", "neg_def", "near_boundary", "zero"])
def test_proj_psd_hard_spectra(dt, name):
    """Spectra that stress the sign iteration: eigenvalues spread over 12 decades on both sides of zero, rank
    deficiency, inputs already inside / entirely outside the cone, iterates hugging the boundary, and X = 0."""
    k = 96
    rng = np.random.default_rng(3)
    lam = {
        "logspaced": np.concatenate([np.logspace(-12, 0, k // 2), -np.logspace(-12, 0, k // 2)]),
        "lowrank": np.concatenate([np.ones(5), np.zeros(k - 5)]),
        "psd_already": np.abs(rng.standard_normal(k)),
        "neg_def": -np.abs(rng.standard_normal(k)) - 0.1,
        "near_boundary": np.concatenate([rng.uniform(0.5, 1, k // 2), rng.standard_normal(k // 2) * 1e-6]),
        "zero": np.zeros(k),
    }[name]
    x = svec(_spectrum_matrix(k, lam, 4)).astype(dt)
    want = x.astype(np.float64).copy()
    O.ConePSD(np.zeros(O.ConePSD.query_worklen(x.size)), 1e-12).proj(False, want)
    scale = max(np.abs(x).max(), 1e-30)
    xb, wb = capi.Buf(x), capi.Buf(np.zeros(2 * k * k + k, dtype=dt))
    capi.check(capi.fn("tb_proj_psd", dt)(xb.view(), 1e-12, wb.view()))
    xb.release(); wb.release()
    assert np.isfinite(x).all()
    tol = 2e-5 if dt == np.float32 else 1e-10
    assert np.abs(x - want).max() <= tol * scale, (name, np.abs(x - want).max() / scale)


@pytest.mark.parametrize("dt", DTYPES)
def test_proj_psd_c4_full_size(dt):
    """Config C4's block: one ConePSD of 512 x 512 (sk = 131 328), against the oracle and by properties:
    result is PSD, idempotent, and x - proj(x) is the projection of -x onto the cone reflected (Moreau)."""
    k = 512
    rng = np.random.default_rng(512)
    g = rng.standard_normal((k, k))
    x0 = svec((g + g.T) / 2).astype(dt)
    want = x0.astype(np.float64).copy()
    O.ConePSD(np.zeros(O.ConePSD.query_worklen(x0.size)), 1e-12).proj(False, want)
    x = x0.copy()
    xb, wb = capi.Buf(x), capi.Buf(np.zeros(2 * k * k + k, dtype=dt))
    capi.check(capi.fn("tb_proj_psd", dt)(xb.view(), 1e-12, wb.view()))
    xb.release(); wb.release()
    tol = 5e-5 if dt == np.float32 else 1e-10
    assert np.abs(x - want).max() <= tol * np.abs(want).max()
    # Moreau decomposition: x0 = proj_K(x0) - proj_K(-x0), and <proj_K(x0), proj_K(-x0)> = 0
    neg = (-x0).copy()
    nb, wb = capi.Buf(neg), capi.Buf(np.zeros(2 * k * k + k, dtype=dt))
    capi.check(capi.fn("tb_proj_psd", dt)(nb.view(), 1e-12, wb.view()))
    nb.release(); wb.release()
    assert np.abs((x.astype(np.float64) - neg.astype(np.float64)) - x0).max() <= 4 * tol * np.abs(x0).max()
    ip = float(np.dot(x.astype(np.float64), neg.astype(np.float64)))
    assert abs(ip) <= 50 * tol * float(np.dot(x.astype(np.float64), x.astype(np.float64)))

@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", [0, 1, 2])
def test_cone_proj_pair_is_one_launch_and_bit_identical(dt, case):
    """The two projections of an iteration (dual cone on y, primal cone on s: solver.rs:548-549) on a product cone WITHOUT PSD
    blocks: the first call is parked and both vectors are projected by one cone_kernel launch (tb_cone_pairs counts it, the
    launch counter sees one launch instead of two); bit-identical to the two separate launches (same code per block), equal to
    the oracle; an unpairable second call (overlapping vector) and a lone call are still served in program order."""
    import ctypes as C
    L = capi.lib()
    blocks = CASES[case]
    m = sum(l for _, l in blocks)
    rng = np.random.default_rng(100 + case)
    y0, s0 = rng.standard_normal(m).astype(dt), rng.standard_normal(m).astype(dt)
    want = []
    for v, dual in ((y0, True), (s0, False)):
        w = v.astype(np.float64).copy()
        oracle_cone(blocks).proj(dual, w)
        want.append(w)
    h = _cone(blocks)
    f = capi.fn("tb_cone_proj", dt)
    none = capi.View(0, 0, 0)
    res, launches = {}, {}
    for pairing in (1, 0):
        capi.check(L.tb_set_psd_pairing(pairing))
        buf = np.concatenate([y0, s0]).copy()
        vb = capi.Buf(buf)
        capi.check(L.tb_flush())
        n0, n1, l0, l1 = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        capi.check(L.tb_cone_pairs(C.byref(n0))); capi.check(L.tb_launch_count(C.byref(l0)))
        capi.check(f(h, 1, vb.view(0, m), 1e-12, none))
        capi.check(f(h, 0, vb.view(m, m), 1e-12, none))
        capi.check(L.tb_flush())
        capi.check(L.tb_cone_pairs(C.byref(n1))); capi.check(L.tb_launch_count(C.byref(l1)))
        vb.release()
        assert n1.value - n0.value == (1 if pairing else 0)
        launches[pairing] = l1.value - l0.value          # cone launches + whatever the uploads of the two views took
        res[pairing] = buf
        tol = 5e-6 if dt == np.float32 else 1e-12
        for i in range(2):
            assert rel_linf(buf[i * m:(i + 1) * m], want[i]) <= tol, (pairing, i)
    capi.check(L.tb_set_psd_pairing(1))
    assert np.array_equal(res[1], res[0])
    assert launches[1] == launches[0] - 1, launches
    # the second call overlaps the first one's vector: not pairable, both run in program order (proj of proj = proj on K, then K*)
    twice = y0.copy()
    vb = capi.Buf(twice)
    capi.check(f(h, 0, vb.view(), 1e-12, none))
    capi.check(f(h, 0, vb.view(), 1e-12, none))
    vb.release()
    w = y0.astype(np.float64).copy()
    oracle_cone(blocks).proj(False, w)
    oracle_cone(blocks).proj(False, w)
    assert rel_linf(twice, w) <= (5e-6 if dt == np.float32 else 1e-12)
    capi.check(L.tb_cone_destroy(h))


# ---- tcgen05 engine of the sign iteration (csrc/psd_tc.cu) -------------------------------------------------
def _sym32(rng, k):
    g = rng.standard_normal((k, k))
    return ((g + g.T) / 2).astype(np.float32)


@pytest.mark.parametrize("k", [64, 128, 132, 200, 256, 384, 512, 640])
@pytest.mark.parametrize("splitk", [0, 1, 2, 4, 8])
def test_symm_gemm_tcgen05_vs_numpy(k, splitk):
    """C = alpha*A*B + beta*D + gamma*I on the tensor cores (3xTF32, split-K over a cluster) against numpy f64:
    fp32-level accuracy (<= 2e-6 of max|C|; a single-pass TF32 product would sit near 1e-3), exactly symmetric
    output, and bit-identical results run to run (fixed reduction order, no atomics)."""
    L = capi.lib()
    rng = np.random.default_rng(1000 * k + splitk)
    a, b, d = _sym32(rng, k), _sym32(rng, k), _sym32(rng, k)
    bufs = [capi.Buf(m.reshape(-1, order="F").copy(), mutable=False) for m in (a, b, d)]
    alpha, beta, gamma = 0.75, -0.5, 1.25
    outs = []
    for _ in range(2):
        c = np.zeros(k * k, dtype=np.float32)
        cb = capi.Buf(c)
        capi.check(L.tb_symm_gemm_f32(k, alpha, bufs[0].view(), bufs[1].view(), beta, bufs[2].view(), gamma, cb.view(), 2, splitk))
        cb.release()
        outs.append(c.reshape(k, k, order="F").copy())
    for bf in bufs:
        bf.release()
    got = outs[0].astype(np.float64)
    full = alpha * (a.astype(np.float64) @ b.astype(np.float64)) + beta * d.astype(np.float64) + gamma * np.eye(k)
    want = np.triu(full) + np.triu(full, 1).T
    assert np.array_equal(outs[0], outs[0].T)
    assert np.array_equal(outs[0], outs[1])
    # the tensor core truncates (does not round) each fp32 accumulation, so the error grows with the K length one CTA
    # accumulates in TMEM: measured 1e-6 (K = 128) ... 4e-6 (K = 512, no split-K); numpy fp32 matmul is at 5e-7
    assert np.abs(got - want).max() <= 8e-6 * np.abs(want).max()


def test_symm_gemm_engines_agree_and_no_d_term():
    """No D term (d.len == 0), FP32-pipe engine vs tcgen05 engine on the same operands."""
    L = capi.lib()
    k = 256
    rng = np.random.default_rng(7)
    a, b = _sym32(rng, k), _sym32(rng, k)
    ab, bb = capi.Buf(a.reshape(-1, order="F").copy(), mutable=False), capi.Buf(b.reshape(-1, order="F").copy(), mutable=False)
    res = {}
    for engine in (1, 2):
        c = np.zeros(k * k, dtype=np.float32)
        cb = capi.Buf(c)
        capi.check(L.tb_symm_gemm_f32(k, 1.0, ab.view(), bb.view(), 0.0, capi.View(0, 0, 0), 0.0, cb.view(), engine, 0))
        cb.release()
        res[engine] = c.astype(np.float64)
    ab.release(); bb.release()
    assert np.abs(res[1] - res[2]).max() <= 3e-6 * np.abs(res[1]).max()
    # an unaligned size is refused by the tensor-core engine, loudly
    k2 = 30
    a2 = _sym32(rng, k2).reshape(-1).copy()
    x = capi.Buf(a2, mutable=False); y = capi.Buf(np.zeros(k2 * k2, dtype=np.float32))
    assert L.tb_symm_gemm_f32(k2, 1.0, x.view(), x.view(), 0.0, capi.View(0, 0, 0), 0.0, y.view(), 2, 0) != 0
    x.release(); y.release()


@pytest.mark.parametrize("k", [64, 128, 132, 256])
@pytest.mark.parametrize("path", [0, 2, 3])
def test_proj_psd_tensor_core_paths_vs_oracle(k, path, request):
    """ConePSD::proj in f32 through the tcgen05 sign iteration (0: split-K chosen, 3: no split-K) and the FP32-pipe
    one (2), each against the oracle's dsyevr + dsyr restatement (f64lapack.rs:78-108)."""
    capi.check(capi.lib().tb_set_psd_path(path))
    request.addfinalizer(lambda: capi.check(capi.lib().tb_set_psd_path(0)))
    rng = np.random.default_rng(k)
    g = rng.standard_normal((k, k))
    x = svec((g + g.T) / 2).astype(np.float32)
    want = x.astype(np.float64).copy()
    O.ConePSD(np.zeros(O.ConePSD.query_worklen(x.size)), 1e-12).proj(False, want)
    xb, wb = capi.Buf(x), capi.Buf(np.zeros(2 * k * k + k, dtype=np.float32))
    capi.check(capi.fn("tb_proj_psd", np.float32)(xb.view(), 1e-12, wb.view()))
    xb.release(); wb.release()
    assert np.abs(x - want).max() <= 2e-5 * max(1.0, np.abs(want).max()), (k, path)


@pytest.mark.parametrize("k", [64, 132])
def test_cone_proj_pairing_batches_the_two_psd_projections(k):
    """The solver's two projections per iteration (dual cone on y, primal cone on s: solver.rs:548-549) on a product cone
    with a PSD block: the first tb_cone_proj_f32 is parked, the second runs both as one batch (tb_psd_pairs counts it);
    results match the oracle and the unbatched run, and a lone projection still completes when anything else is called."""
    import ctypes as C
    L = capi.lib()
    sk = k * (k + 1) // 2
    blocks = [(RPOS, 5), (PSD, sk), (SOC, 7)]
    m = sum(l for _, l in blocks)
    rng = np.random.default_rng(k)

    def vec():
        g = rng.standard_normal((k, k))
        return np.concatenate([rng.standard_normal(5), svec((g + g.T) / 2), rng.standard_normal(7)]).astype(np.float32)
    y0, s0 = vec(), vec()
    want = []
    for v, dual in ((y0, True), (s0, False)):
        w = v.astype(np.float64).copy()
        oracle_cone(blocks).proj(dual, w)
        want.append(w)
    h = _cone(blocks)
    res = {}
    for pairing in (1, 0):
        capi.check(L.tb_set_psd_pairing(pairing))
        n0 = C.c_uint64(); capi.check(L.tb_psd_pairs(C.byref(n0)))
        buf = np.concatenate([y0, s0]).copy()
        vb, wb = capi.Buf(buf), capi.Buf(np.zeros(2 * k * k + k, dtype=np.float32))
        capi.check(L.tb_cone_proj_f32(h, 1, vb.view(0, m), 1e-12, wb.view()))
        capi.check(L.tb_cone_proj_f32(h, 0, vb.view(m, m), 1e-12, wb.view()))
        vb.release(); wb.release()
        n1 = C.c_uint64(); capi.check(L.tb_psd_pairs(C.byref(n1)))
        assert n1.value - n0.value == (1 if pairing else 0)
        res[pairing] = buf
        for i in range(2):
            got = buf[i * m:(i + 1) * m]
            assert np.abs(got - want[i]).max() <= 2e-5 * max(1.0, np.abs(want[i]).max()), (pairing, i)
    assert np.abs(res[1] - res[0]).max() <= 2e-5 * np.abs(res[0]).max()
    # a single parked projection is flushed by the next call (here: the release of its buffer)
    capi.check(L.tb_set_psd_pairing(1))
    lone = y0.copy()
    vb, wb = capi.Buf(lone), capi.Buf(np.zeros(2 * k * k + k, dtype=np.float32))
    capi.check(L.tb_cone_proj_f32(h, 1, vb.view(), 1e-12, wb.view()))
    vb.release(); wb.release()
    assert np.abs(lone - want[0]).max() <= 2e-5 * max(1.0, np.abs(want[0]).max())
    capi.check(L.tb_cone_destroy(h))
