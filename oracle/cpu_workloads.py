"""TEST INFRASTRUCTURE ONLY: the benchmark's synthetic workloads built for the CPU checker / CPU baseline.

bench.py's `--impl reference` arm, its `cpu_baseline` leg and its `parity` leg, and the full-size parity test
(tests/test_full_size_gpu.py) need config C3's A (65536 x 16384; 8.6 GB in f64) on the host.  The counter-based generator of
totsu_b200/csrc/synth.cu is restated in oracle/native.c (pthreads) so the f64 oracle sees bit-identical inputs in seconds;
this module wraps it and builds the two CPU formulations that are timed / checked:

  * `socp_blocks_problem` - the reference's own route for a SOCP: `ProbSOCP` with one `MatOp` G_i (n_i x n) and one vector
    c_i per cone block, i.e. one skinny dgemv + one dot per block per op (totsu/src/problem/socp.rs:83-124, rows
    [-c_i^T; -G_i] per block :359-366);
  * `dense_problem`       - ONE stacked column-major A behind a single `MatOp` (one dgemv per op/trans_op, matop.rs:76-96)
    with a product cone: what the fused DenseOp + ProductCone route of the device computes, and the apples-to-apples CPU
    number next to it.
Nothing here is imported by the product (totsu_b200/)."""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def native_lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle_native.so")
        if not os.path.exists(path):
            raise RuntimeError(path + " is missing: run __graft_entry__.build() (make -C oracle)")
        lib = C.CDLL(path)
        lib.oracle_fill_uniform_f64.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint64, C.c_double, C.c_int, C.c_int]
        lib.oracle_fill_uniform_f64.restype = None
        lib.oracle_fill_uniform_row_f64.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint64, C.c_double, C.c_int]
        lib.oracle_fill_uniform_row_f64.restype = None
        _LIB = lib
    return _LIB


def fill_f64(n_row, n_col, row_offset, seed, scale, as_f32, threads=0, out=None):
    """Rows [row_offset, row_offset + n_row) of the synthetic matrix as a Fortran-ordered f64 array (values rounded like
    the f32 device path when as_f32)."""
    a = np.empty((n_row, n_col), dtype=np.float64, order="F") if out is None else out
    assert a.flags["F_CONTIGUOUS"] and a.shape == (n_row, n_col)
    if a.size:
        native_lib().oracle_fill_uniform_f64(a.ctypes.data, n_row, n_col, n_row, row_offset, seed, float(scale), 1 if as_f32 else 0, threads)
    return a


def fill_row_f64(n_col, row, seed, scale, as_f32):
    v = np.empty(n_col, dtype=np.float64)
    native_lib().oracle_fill_uniform_row_f64(v.ctypes.data, n_col, row, seed, float(scale), 1 if as_f32 else 0)
    return v


def socp_blocks_problem(O, n, n_blocks, bdim, seed, scale, as_f32, x0, s0, y0, round_rhs=np.float32):
    """`ProbSOCP` over the first `n_blocks` cone blocks of the stacked A (block i = rows [i*bdim, (i+1)*bdim): row 0 is
    -c_i^T, the rest -G_i; socp.rs:359-366), with b = A x0 + s0 -> (d_i, h_i) and f = c = -A^T y0 computed in f64 from the
    same rows and rounded to `round_rhs` (the element type the device holds them in)."""
    gs, hs, cs, ds = [], [], [], []
    c_acc = np.zeros(n)
    for i in range(n_blocks):
        r0 = i * bdim
        crow = fill_row_f64(n, r0, seed, scale, as_f32)                      # A[r0, :]      = -c_i^T
        g = fill_f64(bdim - 1, n, r0 + 1, seed, scale, as_f32, threads=1)     # A[r0+1.., :]  = -G_i
        b_blk = np.concatenate([[crow @ x0], g @ x0]) + s0[r0:r0 + bdim]
        c_acc += crow * y0[r0] + g.T @ y0[r0 + 1:r0 + bdim]
        b_blk = b_blk.astype(round_rhs).astype(np.float64)
        np.negative(crow, out=crow)
        np.negative(g, out=g)
        cs.append(O.MatBuild(O.MatType.General(n, 1), crow))
        mb = O.MatBuild(O.MatType.General(bdim - 1, n))
        mb.array = g.reshape(-1, order="F")                                  # no second copy of the block
        gs.append(mb)
        ds.append(float(b_blk[0]))
        hs.append(O.MatBuild(O.MatType.General(bdim - 1, 1), b_blk[1:]))
    f = (-c_acc).astype(round_rhs).astype(np.float64)
    return O.ProbSOCP(O.MatBuild(O.MatType.General(n, 1), f), gs, hs, cs, ds,
                      O.MatBuild(O.MatType.General(0, n)), O.MatBuild(O.MatType.General(0, 1)))


def oracle_cone(O, blocks, eps_zero=1e-12):
    """blocks: [(kind, len)] with kind in {"zero","rpos","soc","rotsoc","psd"}."""
    out = []
    for kind, ln in blocks:
        if kind == "zero":
            out.append((O.ConeZero(), ln))
        elif kind == "rpos":
            out.append((O.ConeRPos(), ln))
        elif kind == "soc":
            out.append((O.ConeSOC(), ln))
        elif kind == "rotsoc":
            out.append((O.ConeRotSOC(), ln))
        elif kind == "psd":
            out.append((O.ConePSD(np.zeros(O.ConePSD.query_worklen(ln)), eps_zero), ln))
        else:
            raise ValueError(kind)
    return O._ProductCone(out)


class DenseProblem:
    """(op_c, op_a, op_b, cone, work) over ONE stacked A, like totsu_core/tests/solver.rs builds a problem by hand."""

    def __init__(self, O, a, b, c, blocks, eps_zero=1e-12):
        self.O, self.a, self.b, self.c, self.blocks, self.eps_zero = O, a, np.asarray(b, dtype=np.float64), np.asarray(c, dtype=np.float64), blocks, eps_zero

    def problem(self):
        O = self.O
        m, n = self.a.shape
        op_c = O.MatOp(O.MatType.General(n, 1), self.c)
        op_a = O.MatOp(O.MatType.General(m, n), self.a.reshape(-1, order="F"))
        op_b = O.MatOp(O.MatType.General(m, 1), self.b)
        return op_c, op_a, op_b, oracle_cone(O, self.blocks, self.eps_zero), np.zeros(O.Solver.query_worklen((m, n)))


def rhs_from_rows(a, x0, s0, y0, round_rhs=np.float32):
    """b = A x0 + s0, c = -A^T y0 in f64, rounded to the device's element type."""
    b = (a @ x0 + s0).astype(round_rhs).astype(np.float64)
    c = (-(a.T @ y0)).astype(round_rhs).astype(np.float64)
    return b, c


def time_iterations(O, prob, steps, warmup):
    """Seconds per solver iteration (update_vecs + criteria_conv, solver.rs:382-386) of the oracle on `prob`, over
    `steps` iterations after `warmup` untimed ones; convergence tests disabled so every run does identical work."""
    s = O.Solver()
    s.par.max_iter = warmup + steps + 1
    s.par.eps_acc = 0.0
    s.par.eps_inf = 0.0
    stamps = []
    orig = s._update_vecs

    def timed_update(*args):
        stamps.append(time.perf_counter())
        return orig(*args)
    s._update_vecs = timed_update
    try:
        s.solve(prob.problem())
    except O.SolverError:
        pass
    stamps.append(time.perf_counter())
    t = stamps[warmup:warmup + steps + 1]
    return (t[-1] - t[0]) / max(1, len(t) - 1), len(t) - 1


def iterates(O, prob, ks, eps_zero=1e-12):
    """x_hat, y_hat after K iterations for each K in ks, plus the per-iteration residual trace."""
    s = O.Solver()
    s.par.max_iter = max(ks) + 2
    s.par.eps_acc = 0.0
    s.par.eps_inf = 0.0
    s.par.eps_zero = eps_zero
    s.snapshots = {k: None for k in ks}
    s.trace = []
    try:
        s.solve(prob.problem())
    except O.SolverError:
        pass
    return s.snapshots, s.trace
