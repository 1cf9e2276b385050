"""Shared test utilities: synthetic instances (SURVEY.md §8d recipe), the oracle's view of a dense conic
problem, and comparison helpers.  Imports the oracle - test infrastructure only."""
from __future__ import annotations

import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "oracle") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))

import totsu_oracle as O  # noqa: E402
from totsu_b200 import capi, synth  # noqa: E402

ZERO, RPOS, SOC, ROTSOC, PSD = capi.CONE_ZERO, capi.CONE_RPOS, capi.CONE_SOC, capi.CONE_ROTSOC, capi.CONE_PSD


def svec(mat):
    """Symmetric matrix -> packed upper triangle by columns with off-diagonals scaled by sqrt(2) (cone_psd.rs:18)."""
    k = mat.shape[0]
    out = []
    for c in range(k):
        for r in range(c + 1):
            out.append(mat[r, c] * (1.0 if r == c else math.sqrt(2.0)))
    return np.array(out)


def interior_point(blocks, rng, dual=False):
    """A point strictly inside K (or K* when dual) for a list of (type, len) blocks."""
    parts = []
    for t, ln in blocks:
        if ln == 0:
            parts.append(np.zeros(0))
        elif t == ZERO:
            parts.append(rng.standard_normal(ln) if dual else np.zeros(ln))
        elif t == RPOS:
            parts.append(np.abs(rng.standard_normal(ln)) + 0.1)
        elif t == SOC:
            v = rng.standard_normal(ln - 1)
            parts.append(np.concatenate([[np.linalg.norm(v) + 1.0], v]))
        elif t == ROTSOC:
            if ln == 1:
                parts.append(np.array([1.0]))
            else:
                v = rng.standard_normal(ln - 2)
                r = 1.0 + abs(rng.standard_normal())
                s = (np.dot(v, v) + 1.0) / (2.0 * r)          # 2 r s > ||v||^2
                parts.append(np.concatenate([[r, s], v]))
        elif t == PSD:
            k = int((math.sqrt(8 * ln + 1) - 1) / 2 + 0.5)
            g = rng.standard_normal((k, k))
            parts.append(svec(g.T @ g / k + np.eye(k)))
        else:
            raise ValueError(t)
    return np.concatenate(parts) if parts else np.zeros(0)


def make_instance(m, n, blocks, seed=0, dtype=np.float32):
    """Feasible-by-construction dense conic instance: A ~ U(-1,1)/sqrt(n) from the counter-based generator,
    b = A x0 + s0, c = -A^T y0 with s0 in int K, y0 in int K*.  Everything is rounded to `dtype` so the f64 oracle
    and the device see identical inputs.  Returns (A column-major (m,n) dtype, b, c)."""
    assert sum(ln for _, ln in blocks) == m
    rng = np.random.default_rng(seed + 12345)
    scale = np.dtype(dtype).type(1.0 / math.sqrt(n))
    a = synth.uniform_matrix(m, n, seed, scale, dtype=np.dtype(dtype).type)
    a64 = a.astype(np.float64)
    x0 = rng.standard_normal(n)
    s0 = interior_point(blocks, rng, dual=False)
    y0 = interior_point(blocks, rng, dual=True)
    b = (a64 @ x0 + s0).astype(dtype)
    c = (-(a64.T @ y0)).astype(dtype)
    return a, b, c


def oracle_cone(blocks, eps_zero=1e-12):
    out = []
    for t, ln in blocks:
        if t == ZERO:
            out.append((O.ConeZero(), ln))
        elif t == RPOS:
            out.append((O.ConeRPos(), ln))
        elif t == SOC:
            out.append((O.ConeSOC(), ln))
        elif t == ROTSOC:
            out.append((O.ConeRotSOC(), ln))
        elif t == PSD:
            out.append((O.ConePSD(np.zeros(O.ConePSD.query_worklen(ln)), eps_zero), ln))
    return O._ProductCone(out)


def oracle_dense_problem(a, b, c, blocks, eps_zero=1e-12):
    """(op_c, op_a, op_b, cone, work) in the oracle for one dense A, like totsu_core/tests/solver.rs builds by hand."""
    m, n = a.shape
    op_c = O.MatOp(O.MatType.General(n, 1), np.asarray(c, dtype=np.float64))
    op_a = O.MatOp(O.MatType.General(m, n), np.asarray(a, dtype=np.float64).reshape(-1, order="F"))
    op_b = O.MatOp(O.MatType.General(m, 1), np.asarray(b, dtype=np.float64))
    work = np.zeros(O.Solver.query_worklen((m, n)))
    return op_c, op_a, op_b, oracle_cone(blocks, eps_zero), work


def oracle_iterates(a, b, c, blocks, ks, eps_zero=1e-12):
    """x_hat, y_hat after K iterations for each K in ks (raw iterates, before any 1/tau scaling), + trace."""
    s = O.Solver()
    kmax = max(ks)
    s.par.max_iter = kmax + 2
    s.par.eps_acc = 0.0
    s.par.eps_inf = 0.0
    s.par.eps_zero = eps_zero
    s.snapshots = {k: None for k in ks}
    s.trace = []
    try:
        s.solve(oracle_dense_problem(a, b, c, blocks, eps_zero))
    except O.SolverError:
        pass
    return s.snapshots, s.trace


def rel_linf(got, want):
    want = np.asarray(want, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    den = max(np.abs(want).max(), 1e-300) if want.size else 1.0
    return float(np.abs(got - want).max() / den) if want.size else 0.0


def device_matrix(a):
    """Upload a column-major numpy matrix into a device-only backend buffer; returns (Buf, view)."""
    flat = np.ascontiguousarray(a.reshape(-1, order="F"))
    buf = capi.Buf(dtype=flat.dtype, length=flat.size)
    buf.upload(flat)
    return buf, buf.view()


def smoke_check(np_mod=np):
    """__graft_entry__.smoke(): a small fused SOCP run + one TMA matvec, each checked against the oracle."""
    from totsu_b200 import host
    import ctypes as C
    # 1) matvec through the streaming (TMA) kernel
    rng = np.random.default_rng(0)
    m, n = 2048, 256
    a = np.asfortranarray(rng.standard_normal((m, n)).astype(np.float32))
    x = rng.standard_normal(n).astype(np.float32)
    y = np.zeros(m, dtype=np.float32)
    abuf, av = device_matrix(a)
    xb, yb = capi.Buf(x.copy()), capi.Buf(y)
    capi.check(capi.lib().tb_set_gemv_path(2))
    capi.check(capi.lib().tb_transform_ge_f32(0, m, n, 1.0, av, xb.view(), 0.0, yb.view()))
    capi.check(capi.lib().tb_set_gemv_path(0))
    got = yb.download()
    want = a.astype(np.float64) @ x.astype(np.float64)
    err = rel_linf(got, want)
    assert err < 1e-5, "smoke matvec mismatch: %g" % err
    for bf in (abuf, xb, yb):
        bf.release()
    # 2) fused SOCP: 8 SOC blocks of dim 16 + 8 equality rows, 10 iterations vs the oracle
    blocks = [(SOC, 16)] * 8 + [(ZERO, 8)]
    m, n = 136, 48
    a, b, c = make_instance(m, n, blocks, seed=1, dtype=np.float32)
    snaps, _ = oracle_iterates(a, b, c, blocks, [10])
    abuf, av = device_matrix(a)
    s = host.Session.dense(np.float32, av, m, n, c, b, blocks, fused_op=True, fused_cone=True)
    assert s.begin(max_iter=None, eps_acc=0.0, eps_inf=0.0) == "None"
    s.step(10)
    xh, yh = s.xy()
    s.close()
    abuf.release()
    ex, ey = rel_linf(xh, snaps[10][0]), rel_linf(yh, snaps[10][1])
    assert ex < 1e-4 and ey < 1e-4, "smoke SOCP mismatch: %g %g" % (ex, ey)
