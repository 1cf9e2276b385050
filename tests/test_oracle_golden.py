"""Pins the CPU oracle (oracle/totsu_oracle.py) against every known answer the reference holds
for the hot path.  CPU only; no GPU, no /root/reference access at run time."""
import math
import os

import numpy as np
import pytest

import totsu_oracle as O
from totsu_oracle import (MatType, MatOp, MatBuild, Solver, SolverError, ConePSD, ConeRPos,
                          ProbLP, ProbQP, ProbQCQP, ProbSOCP, ProbSDP, F64LAPACK)

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _cortex_m_lp():
    # examples/nostd_cortex-m/src/main.rs:64-89
    op_c = MatOp(MatType.General(2, 1), [-1., 0.])
    op_a = MatOp(MatType.General(3, 2), [4., -1., -1., -1., 4., -1.])
    op_b = MatOp(MatType.General(3, 1), [6., 6., 1.])
    return op_c, op_a, op_b, ConeRPos()


def test_golden_trace_cortex_m_lp():
    """examples/nostd_cortex-m/log_qemu.txt:7-25 — every printed residual line, iteration 159, 16-digit solution."""
    op_c, op_a, op_b, cone = _cortex_m_lp()
    s = Solver().set_par(lambda p: (setattr(p, "max_iter", 100_000), setattr(p, "log_period", 10)))
    assert Solver.query_worklen(op_a.size()) == 48                     # log_qemu.txt:3
    work = np.zeros(48)
    x, y = s.solve((op_c, op_a, op_b, cone, work))
    want = [l.strip() for l in open(os.path.join(GOLDEN, "cortex_m_lp_trace.txt")) if l.strip() and not l.startswith("#")]
    assert s.log == want
    assert s.iters == 159
    assert x.tolist() == [1.9999994251590176, 2.0000004472430635]     # log_qemu.txt:25


def test_backend_conformance_sdp():
    """totsu_f64lapack/tests/solver.rs:15-56 (and totsu_core/tests/solver.rs:13-54): x[0] = -2 +- 1e-3."""
    op_c = MatOp(MatType.General(1, 1), [1.])
    op_a = MatOp(MatType.General(3, 1), [0., -1. * 1.41421356, -3.])
    op_b = MatOp(MatType.General(3, 1), [1., 0. * 1.41421356, 10.])
    s = Solver().set_par(lambda p: setattr(p, "max_iter", 100_000))
    cone_w = np.zeros(ConePSD.query_worklen(op_a.size()[0]))
    cone = ConePSD(cone_w, s.par.eps_zero)
    work = np.zeros(Solver.query_worklen(op_a.size()))
    x, _ = s.solve((op_c, op_a, op_b, cone, work))
    assert abs(x[0] - (-2.0)) <= 1e-3


def test_lp1_infeasible():
    """totsu/tests/lp.rs:12-45"""
    vec_c = MatBuild(MatType.General(1, 1)).iter_colmaj([1.])
    mat_g = MatBuild(MatType.General(2, 1)).iter_rowmaj([1., -1.])
    vec_h = MatBuild(MatType.General(2, 1)).iter_colmaj([-5., -10.])
    mat_a = MatBuild(MatType.General(0, 1))
    vec_b = MatBuild(MatType.General(0, 1))
    s = Solver().set_par(lambda p: setattr(p, "max_iter", 100_000))
    lp = ProbLP(vec_c, mat_g, vec_h, mat_a, vec_b)
    with pytest.raises(SolverError) as e:
        s.solve(lp.problem())
    assert e.value.kind == SolverError.Infeasible


def test_lp2_unbounded():
    """totsu/tests/lp.rs:49-82"""
    vec_c = MatBuild(MatType.General(1, 1)).iter_colmaj([1.])
    mat_g = MatBuild(MatType.General(2, 1)).iter_rowmaj([1., 1.])
    vec_h = MatBuild(MatType.General(2, 1)).iter_colmaj([5., 10.])
    mat_a = MatBuild(MatType.General(0, 1))
    vec_b = MatBuild(MatType.General(0, 1))
    s = Solver().set_par(lambda p: setattr(p, "max_iter", 100_000))
    lp = ProbLP(vec_c, mat_g, vec_h, mat_a, vec_b)
    with pytest.raises(SolverError) as e:
        s.solve(lp.problem())
    assert e.value.kind == SolverError.Unbounded


def test_qp1():
    """totsu/tests/qp.rs:13-49 and the crate doc-test totsu_f64lapack/src/lib.rs:31-78: x = [2, 0]."""
    n, m, p = 2, 1, 0
    sym_p = MatBuild(MatType.SymPack(n)); sym_p[(0, 0)] = 1.; sym_p[(1, 1)] = 1.
    vec_q = MatBuild(MatType.General(n, 1)); vec_q[(0, 0)] = 1.; vec_q[(1, 0)] = 2.
    mat_g = MatBuild(MatType.General(m, n)); mat_g[(0, 0)] = -1. / 2.; mat_g[(0, 1)] = -1. / 3.
    vec_h = MatBuild(MatType.General(m, 1)); vec_h[(0, 0)] = -1.
    mat_a = MatBuild(MatType.General(p, n)); vec_b = MatBuild(MatType.General(p, 1))
    s = Solver().set_par(lambda q: setattr(q, "max_iter", 100_000))
    qp = ProbQP(sym_p, vec_q, mat_g, vec_h, mat_a, vec_b, s.par.eps_zero)
    x, _ = s.solve(qp.problem())
    assert np.allclose(x[0:2], [2., 0.], atol=1e-3)


def test_qcqp1():
    """totsu/tests/qcqp.rs:13-48: x = [5, 4]."""
    n, m, p = 2, 1, 0
    syms_p = [MatBuild(MatType.SymPack(n)) for _ in range(m + 1)]
    syms_p[0][(0, 0)] = 1.; syms_p[0][(1, 1)] = 1.
    vecs_q = [MatBuild(MatType.General(n, 1)) for _ in range(m + 1)]
    vecs_q[0][(0, 0)] = -5.; vecs_q[0][(1, 0)] = -4.
    vecs_q[1][(0, 0)] = -1. / 2.; vecs_q[1][(1, 0)] = -1. / 3.
    scls_r = [0., 1.]
    mat_a = MatBuild(MatType.General(p, n)); vec_b = MatBuild(MatType.General(p, 1))
    s = Solver().set_par(lambda q: setattr(q, "max_iter", 100_000))
    qp = ProbQCQP(syms_p, vecs_q, scls_r, mat_a, vec_b, s.par.eps_zero)
    x, _ = s.solve(qp.problem())
    assert np.allclose(x[0:2], [5., 4.], atol=1e-3)


def test_socp1():
    """totsu/tests/socp.rs:13-47: x = [-1, -1]."""
    n, m, p, ni = 2, 1, 0, 2
    vec_f = MatBuild(MatType.General(n, 1)).by_fn(lambda r, c: 1.)
    mats_g = [MatBuild(MatType.General(ni, n))]
    mats_g[0][(0, 0)] = 1.; mats_g[0][(1, 1)] = 1.
    vecs_h = [MatBuild(MatType.General(ni, 1))]
    vecs_c = [MatBuild(MatType.General(n, 1))]
    scls_d = [math.sqrt(2.)]
    mat_a = MatBuild(MatType.General(p, n)); vec_b = MatBuild(MatType.General(p, 1))
    s = Solver()
    socp = ProbSOCP(vec_f, mats_g, vecs_h, vecs_c, scls_d, mat_a, vec_b)
    x, _ = s.solve(socp.problem())
    assert np.allclose(x, [-1., -1.], atol=1e-3)


def test_socp2_zero_row_block():
    """totsu/tests/socp.rs:51-94: x = [2, 0]; the first G block has 0 rows (matop.rs:80-85 short-circuit)."""
    n, m, p = 2, 2, 0
    vec_f = MatBuild(MatType.General(n, 1)).iter_colmaj([0., 1.])
    mats_g = [MatBuild(MatType.General(0, n)), MatBuild(MatType.General(1, n)).iter_rowmaj([-1.0, 0.0])]
    vecs_h = [MatBuild(MatType.General(0, 1)), MatBuild(MatType.General(1, 1)).iter_colmaj([2.])]
    vecs_c = [MatBuild(MatType.General(m, 1)).iter_colmaj([0., -1.0]),
              MatBuild(MatType.General(m, 1)).iter_colmaj([0., 1.0])]
    scls_d = [50., 0.]
    mat_a = MatBuild(MatType.General(p, n)); vec_b = MatBuild(MatType.General(p, 1))
    s = Solver().set_par(lambda q: setattr(q, "max_iter", 100_000))
    socp = ProbSOCP(vec_f, mats_g, vecs_h, vecs_c, scls_d, mat_a, vec_b)
    x, _ = s.solve(socp.problem())
    assert np.allclose(x, [2., 0.], atol=1e-3)


def test_sdp1():
    """totsu/tests/sdp.rs:13-51: x = [3, 4]."""
    n, p, k = 2, 0, 2
    vec_c = MatBuild(MatType.General(n, 1)).iter_colmaj([1., 1.])
    syms_f = [MatBuild(MatType.SymPack(k)) for _ in range(n + 1)]
    syms_f[0].iter_rowmaj([-1., 0., 0., 0.])
    syms_f[1].iter_rowmaj([0., 0., 0., -1.])
    syms_f[2].iter_rowmaj([3., 0., 0., 4.])
    mat_a = MatBuild(MatType.General(p, n)); vec_b = MatBuild(MatType.General(p, 1))
    s = Solver().set_par(lambda q: setattr(q, "max_iter", 100_000))
    sdp = ProbSDP(vec_c, syms_f, mat_a, vec_b, s.par.eps_zero)
    x, _ = s.solve(sdp.problem())
    assert np.allclose(x, [3., 4.], atol=1e-3)


def test_matop_sympack_vs_dense():
    """totsu_core/src/matop.rs:179-212"""
    array = [1., 2., 3., 4., 5., 6., 7., 8., 9., 10., 11., 12., 13., 14., 15.]
    ref = np.array([[1., 2., 4., 7., 11.], [2., 3., 5., 8., 12.], [4., 5., 6., 9., 13.],
                    [7., 8., 9., 10., 14.], [11., 12., 13., 14., 15.]])
    m = MatOp(MatType.SymPack(5), array)
    x = np.zeros(5); y = np.zeros(5)
    for i in range(5):
        x[i] = 1.
        m.op(1., x, 0., y)
        assert np.allclose(y, ref[i], atol=1e-3)
        x[i] = 0.


def test_cone_psd_unit():
    """totsu_core/src/cone_psd.rs:89-110: proj of diag(5,-5) -> diag(5,0)."""
    x = np.array([5., 0., -5.])
    assert ConePSD.query_worklen(x.size) <= 10
    c = ConePSD(np.zeros(10), 1e-12)
    c.proj(False, x)
    assert np.allclose(x, [5., 0., 0.], atol=1e-6)


def test_matbuild_scale_nondiag():
    """totsu/src/matbuild/mod.rs:304-333"""
    ref = [1., 2. * 1.4, 3., 4. * 1.4, 5. * 1.4, 6., 7. * 1.4, 8. * 1.4, 9. * 1.4, 10.,
           11. * 1.4, 12. * 1.4, 13. * 1.4, 14. * 1.4, 15.]
    array = [1., 0., 0., 0., 0., 2., 3., 0., 0., 0., 4., 5., 6., 0., 0., 7., 8., 9., 10., 0., 11., 12., 13., 14., 15.]
    m = MatBuild(MatType.SymPack(5)).iter_colmaj(array).scale_nondiag(1.4)
    assert np.allclose(m.array, ref, atol=1e-3)


def test_vec_to_mat_roundtrip():
    """totsu_f64lapack/src/f64lapack.rs:262-287"""
    ref_v = np.array([1. * 0.7, 2., 3. * 0.7, 4., 5., 6. * 0.7, 7., 8., 9., 10. * 0.7, 11., 12., 13., 14., 15. * 0.7])
    ref_m = np.array([1., 0., 0., 0., 0., 2., 3., 0., 0., 0., 4., 5., 6., 0., 0., 7., 8., 9., 10., 0., 11., 12., 13., 14., 15.])
    v = ref_v.copy()
    m = O.vec_to_mat(v, 5, math.sqrt(2.))
    assert np.allclose(m.reshape(-1, order="F"), ref_m, atol=0.5)
    O.mat_to_vec(m, v, math.sqrt(2.))
    assert np.allclose(v, ref_v, atol=1e-6)


def test_operator_doc_definitions():
    """Normative brute-force definitions of trans_op / absadd_cols / absadd_rows (operator.rs:40-154)
    checked on a composite front-end operator (the imgnr_udef/utils2 operator_ref property test)."""
    rng = np.random.default_rng(7)
    n, blocks, p = 5, [3, 0, 4], 2
    mats_g = [MatBuild(MatType.General(ni, n), rng.standard_normal(ni * n)) for ni in blocks]
    vecs_h = [MatBuild(MatType.General(ni, 1), rng.standard_normal(ni)) for ni in blocks]
    vecs_c = [MatBuild(MatType.General(n, 1), rng.standard_normal(n)) for _ in blocks]
    socp = ProbSOCP(MatBuild(MatType.General(n, 1), rng.standard_normal(n)), mats_g, vecs_h, vecs_c,
                    list(rng.standard_normal(len(blocks))), MatBuild(MatType.General(p, n), rng.standard_normal(p * n)),
                    MatBuild(MatType.General(p, 1), rng.standard_normal(p)))
    _, op_a, op_b, _, _ = socp.problem()
    for op in (op_a, op_b):
        m, nn = op.size()
        dense = np.zeros((m, nn))
        for j in range(nn):
            e = np.zeros(nn); e[j] = 1.
            op.op(1., e, 0., dense[:, j])
        x = rng.standard_normal(m); y0 = rng.standard_normal(nn)
        y = y0.copy(); op.trans_op(0.7, x, -0.3, y)
        assert np.allclose(y, 0.7 * dense.T @ x - 0.3 * y0, atol=1e-12)
        if op is op_a:   # OpB's absadd_rows adds scl_d unsigned-less (socp.rs:272), so only check A
            t = np.zeros(nn); op.absadd_cols(t); assert np.allclose(t, np.abs(dense).sum(0), atol=1e-12)
            sg = np.zeros(m); op.absadd_rows(sg); assert np.allclose(sg, np.abs(dense).sum(1), atol=1e-12)
