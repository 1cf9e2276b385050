"""CPU oracle for the Totsu first-order conic iteration (TEST INFRASTRUCTURE ONLY).

This file is a numpy/scipy restatement of the reference's `totsu_f64lapack` path:
the solver loop, the self-dual embedding, `MatOp`, the five cones, the
`F64LAPACK` backend and the LP/QP/QCQP/SOCP/SDP front-ends.  It is the checker the
parity tests compare the CUDA path against.  Nothing in the product
(`totsu_b200/`) imports it; only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s CPU-baseline / `--impl reference` legs do.

The reference is Rust and cannot be compiled in this image (no rustc/cargo), and
the arithmetic of `totsu_f64lapack` lives in un-vendored crates
(cblas 0.4.0 / lapacke 0.5.0 over intel-mkl-src 0.8.1).  Here the same BLAS/LAPACK
routines are reached through numpy (OpenBLAS `dgemv`, `ddot`, ...) and
`scipy.linalg.lapack.dsyevr` (OpenBLAS 0.3.x instead of MKL).

Parity pinning (see tests/test_oracle_golden.py): reproduces
`examples/nostd_cortex-m/log_qemu.txt` (every printed residual, iteration 159 and
the 16-digit solution), the backend-conformance SDP (x = -2), and every known
answer of `totsu/tests/{lp,qp,qcqp,socp,sdp}.rs` including Infeasible/Unbounded.

Every function cites the reference file:line it follows (paths relative to
`solver_rust_conic/`).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
from scipy.linalg import lapack as _lapack

F = np.float64


# --------------------------------------------------------------------------------------
# Errors — totsu_core/src/solver/solver_error.rs:3-18
# --------------------------------------------------------------------------------------
class SolverError(Exception):
    Unbounded = "Unbounded"
    Infeasible = "Infeasible"
    ExcessIter = "ExcessIter"
    InvalidOp = "InvalidOp"
    WorkShortage = "WorkShortage"
    ConeFailure = "ConeFailure"

    def __init__(self, kind: str):
        super().__init__(kind)
        self.kind = kind


# --------------------------------------------------------------------------------------
# F64LAPACK — totsu_f64lapack/src/f64lapack.rs:15-255
# Vectors are numpy float64 views; all ops are in place like the cblas calls.
# --------------------------------------------------------------------------------------
class F64LAPACK:
    @staticmethod
    def norm(x: np.ndarray) -> float:                       # f64lapack.rs:20-23 (dnrm2)
        return float(np.sqrt(np.dot(x, x))) if x.size else 0.0

    @staticmethod
    def copy(x: np.ndarray, y: np.ndarray) -> None:          # :25-30 (dcopy)
        assert x.shape == y.shape
        y[...] = x

    @staticmethod
    def scale(alpha: float, x: np.ndarray) -> None:          # :32-35 (dscal)
        if alpha == 0.0:
            x[...] = 0.0            # the solver uses alpha=0 as a zero-fill (solver.rs:95,165,490)
        else:
            x *= alpha

    @staticmethod
    def add(alpha: float, x: np.ndarray, y: np.ndarray) -> None:   # :37-42 (daxpy)
        assert x.shape == y.shape
        y += alpha * x

    @staticmethod
    def adds(s: float, y: np.ndarray) -> None:               # :44-49 (daxpy incx=0)
        y += s

    @staticmethod
    def abssum(x: np.ndarray, incx: int) -> float:           # :51-59 (dasum, count=ceil(len/incx))
        if incx == 0:
            return 0.0
        return float(np.abs(x[::incx]).sum())

    @staticmethod
    def transform_di(alpha, mat, x, beta, y) -> None:        # :61-73 (dsbmv k=0)
        assert mat.shape == x.shape == y.shape
        if beta == 0.0:
            y[...] = alpha * (mat * x)
        else:
            y *= beta
            y += alpha * (mat * x)

    @staticmethod
    def transform_ge(transpose, n_row, n_col, alpha, mat, x, beta, y) -> None:   # :123-146 (dgemv ColumnMajor lda=n_row)
        assert mat.size == n_row * n_col
        a = mat.reshape((n_row, n_col), order="F")
        if transpose:
            assert x.size == n_row and y.size == n_col
            r = a.T @ x
        else:
            assert x.size == n_col and y.size == n_row
            r = a @ x
        if beta == 0.0:
            y[...] = alpha * r
        else:
            y *= beta
            y += alpha * r

    @staticmethod
    def transform_sp(n, alpha, mat, x, beta, y) -> None:     # :149-163 (dspmv Upper, packed by columns)
        assert mat.size == n * (n + 1) // 2 and x.size == n and y.size == n
        if n == 0:
            return
        if mat.dtype == np.float64 and x.dtype == np.float64 and y.dtype == np.float64:
            # the same BLAS routine the reference calls (cblas::dspmv, ColumnMajor, Upper): packed index c*(c+1)/2 + r
            from scipy.linalg.blas import dspmv
            r = dspmv(n, alpha, mat, np.ascontiguousarray(x), beta=beta, y=np.array(y, dtype=np.float64), lower=0, overwrite_y=1)
            y[...] = r
            return
        full = np.zeros((n, n), dtype=mat.dtype)
        # packed index c*(c+1)/2 + r (r <= c): iterate columns, rows within column
        full.T[np.tril_indices(n)] = mat           # full.T lower (c, r<=c) row-major == packed order
        full = np.triu(full) + np.triu(full, 1).T
        r = full @ x
        if beta == 0.0:
            y[...] = alpha * r
        else:
            y *= beta
            y += alpha * r

    @staticmethod
    def map_eig_worklen(n: int) -> int:                      # :165-170
        return n * n + n + n * n

    @staticmethod
    def map_eig(mat: np.ndarray, scale_diag: Optional[float], eps_zero: float,
                work: np.ndarray, fmap: Callable[[float], Optional[float]]) -> None:   # :172-190
        sn = mat.size
        n = (int(math.sqrt(8 * sn + 1)) - 1) // 2
        assert n * (n + 1) // 2 == sn
        assert work.size >= F64LAPACK.map_eig_worklen(n)
        a = vec_to_mat(mat, n, scale_diag)
        a = eig_func(a, n, eps_zero, fmap)
        mat_to_vec(a, mat, scale_diag)


def vec_to_mat(v: np.ndarray, n: int, scale: Optional[float]) -> np.ndarray:   # f64lapack.rs:195-224
    a = np.zeros((n, n), dtype=np.float64, order="F")
    pos = 0
    for c in range(n):                      # upper triangle packed by columns
        a[: c + 1, c] = v[pos: pos + c + 1]
        pos += c + 1
    if scale is not None:
        a[np.arange(n), np.arange(n)] *= scale      # dscal stride n+1
    return a


def mat_to_vec(a: np.ndarray, v: np.ndarray, scale: Optional[float]) -> None:  # f64lapack.rs:226-255
    n = a.shape[0]
    if scale is not None:
        a[np.arange(n), np.arange(n)] *= 1.0 / scale
    pos = 0
    for c in range(n):
        v[pos: pos + c + 1] = a[: c + 1, c]
        pos += c + 1


def eig_func(a: np.ndarray, n: int, eps_zero: float, func) -> np.ndarray:      # f64lapack.rs:78-108
    """dsyevr(jobz=V, range=V, uplo=U, vl=0, vu=+inf, abstol=eps_zero) then a := sum func(w_i) z_i z_i^T (upper)."""
    w, z, m, _isuppz, info = _lapack.dsyevr(a, compute_v=1, range="V", lower=0,
                                            vl=0.0, vu=np.inf, abstol=eps_zero)
    # NOTE: the reference discards dsyevr's return value (f64lapack.rs:86-91), and OpenBLAS' dsyevr does
    # return info=1 on some 2x2 inputs of the conformance SDP while still delivering the right eigenpair.
    eig_func.nonzero_info += int(info != 0)
    out = np.zeros((n, n), dtype=np.float64, order="F")
    es = []
    cols = []
    for i in range(int(m)):
        e = func(float(w[i]))
        if e is not None:
            es.append(e)
            cols.append(i)
    if cols:
        zc = z[:, cols]
        out[...] = np.triu((zc * np.asarray(es)) @ zc.T)      # the dsyr(Upper) loop, batched
    return out


eig_func.nonzero_info = 0


# --------------------------------------------------------------------------------------
# MatOp — totsu_core/src/matop.rs:9-175
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class MatType:
    kind: str
    a: int
    b: int = 0

    @staticmethod
    def General(n_row: int, n_col: int) -> "MatType":
        return MatType("General", n_row, n_col)

    @staticmethod
    def SymPack(n: int) -> "MatType":
        return MatType("SymPack", n, n)

    def len(self) -> int:                                    # matop.rs:24-30
        return self.a * self.b if self.kind == "General" else self.a * (self.a + 1) // 2

    def size(self) -> Tuple[int, int]:                       # matop.rs:33-39
        return (self.a, self.b)


class MatOp:
    L = F64LAPACK

    def __init__(self, typ: MatType, array):                 # matop.rs:64-74
        self.typ = typ
        self.array = np.ascontiguousarray(array, dtype=np.float64).reshape(-1)
        assert typ.len() == self.array.size

    def size(self):
        return self.typ.size()

    def _op_impl(self, transpose, alpha, x, beta, y):        # matop.rs:76-96
        L = self.L
        if self.typ.kind == "General":
            nr, nc = self.typ.a, self.typ.b
            if nr > 0 and nc > 0:
                L.transform_ge(transpose, nr, nc, alpha, self.array, x, beta, y)
            else:
                L.scale(beta, y)
        else:
            n = self.typ.a
            if n > 0:
                L.transform_sp(n, alpha, self.array, x, beta, y)
            else:
                L.scale(beta, y)

    def op(self, alpha, x, beta, y):
        self._op_impl(False, alpha, x, beta, y)

    def trans_op(self, alpha, x, beta, y):
        self._op_impl(True, alpha, x, beta, y)

    def _absadd_impl(self, colwise: bool, y: np.ndarray):    # matop.rs:98-138
        L = self.L
        if self.typ.kind == "General":
            nr, nc = self.typ.a, self.typ.b
            if nr == 0 or nc == 0:
                assert y.size == (nc if colwise else nr)
                return
            a = np.abs(self.array.reshape((nr, nc), order="F"))
            if colwise:
                assert nc == y.size
                y += a.sum(axis=0)            # one abssum(col, 1) per column
            else:
                assert nr == y.size
                y += a.sum(axis=1)            # one abssum(row, nr) per row
        else:
            n = self.typ.a
            assert n == y.size
            s = 0
            for c in range(n):
                col = self.array[s: s + c + 1]
                s += c + 1
                y[c] = L.abssum(col, 1) + y[c]
                y[:c] += np.abs(col[:c])

    def absadd_cols(self, tau):
        self._absadd_impl(True, tau)

    def absadd_rows(self, sigma):
        self._absadd_impl(False, sigma)


# --------------------------------------------------------------------------------------
# Cones — totsu_core/src/cone_{zero,rpos,soc,rotsoc,psd}.rs
# --------------------------------------------------------------------------------------
class ConeZero:
    L = F64LAPACK

    def proj(self, dual_cone: bool, x: np.ndarray) -> None:  # cone_zero.rs:38-44
        if not dual_cone:
            self.L.scale(0.0, x)

    def product_group(self, dp_tau, group):                  # cone_zero.rs:46-49
        pass


class ConeRPos:
    def proj(self, dual_cone: bool, x: np.ndarray) -> None:  # cone_rpos.rs:38-45
        np.maximum(x, 0.0, out=x)

    def product_group(self, dp_tau, group):
        pass


class ConeSOC:
    L = F64LAPACK

    def proj(self, dual_cone: bool, x: np.ndarray) -> None:  # cone_soc.rs:38-65
        L = self.L
        if x.size > 0:
            s = x[:1]
            v = x[1:]
            val_s = float(s[0])
            norm_v = L.norm(v)
            if norm_v <= -val_s:
                L.scale(0.0, v)
                s[0] = 0.0
            elif norm_v <= val_s:
                pass
            else:
                alpha = (1.0 + val_s / norm_v) / 2.0
                L.scale(alpha, v)
                s[0] = (norm_v + val_s) / 2.0

    def product_group(self, dp_tau, group):                  # cone_soc.rs:67-70
        group(dp_tau)


class ConeRotSOC:
    def __init__(self):
        self.soc = ConeSOC()

    def proj(self, dual_cone: bool, x: np.ndarray) -> None:  # cone_rotsoc.rs:38-65
        fsqrt2 = math.sqrt(2.0)
        if x.size > 0:
            if x.size == 1:
                x[0] = max(float(x[0]), 0.0)
            else:
                r, s = float(x[0]), float(x[1])
                x[0] = (r + s) / fsqrt2
                x[1] = (r - s) / fsqrt2
                self.soc.proj(dual_cone, x)
                r, s = float(x[0]), float(x[1])
                x[0] = (r + s) / fsqrt2
                x[1] = (r - s) / fsqrt2

    def product_group(self, dp_tau, group):
        group(dp_tau)


class ConePSD:
    L = F64LAPACK

    @classmethod
    def query_worklen(cls, nvars: int) -> int:               # cone_psd.rs:32-38
        n = (int(math.sqrt(8 * nvars + 1)) - 1) // 2
        assert n * (n + 1) // 2 == nvars
        return cls.L.map_eig_worklen(n)

    def __init__(self, work: np.ndarray, eps_zero: float):   # cone_psd.rs:40-46
        self.work = work
        self.eps_zero = eps_zero

    def proj(self, dual_cone: bool, x: np.ndarray) -> None:  # cone_psd.rs:56-79
        if self.work.size < self.query_worklen(x.size):
            raise SolverError(SolverError.ConeFailure)
        self.L.map_eig(x, math.sqrt(2.0), self.eps_zero, self.work,
                       lambda e: e if e > 0.0 else None)

    def product_group(self, dp_tau, group):                  # cone_psd.rs:81-84
        group(dp_tau)


# --------------------------------------------------------------------------------------
# Solver — totsu_core/src/solver/solver.rs
# --------------------------------------------------------------------------------------
@dataclass
class SolverParam:                                           # solver.rs:13-41
    max_iter: Optional[int] = None
    eps_acc: float = 1e-6
    eps_inf: float = 1e-6
    eps_zero: float = 1e-12
    log_period: int = 10_000


def fmt_e2(v: float) -> str:
    """Rust's `{:.2e}` (LowerExp, 2 decimals): '8.95e-1', '0.00e0' (solver.rs:391)."""
    if v == 0.0:
        return "0.00e0"
    if math.isinf(v):
        return "inf" if v > 0 else "-inf"
    if math.isnan(v):
        return "NaN"
    m, e = ("%.2e" % v).split("e")
    return "%se%d" % (m, int(e))


class SelfDualEmbed:                                         # solver.rs:45-184
    L = F64LAPACK

    def __init__(self, c, a, b, L=None):
        self.c, self.a, self.b = c, a, b
        if L is not None:
            self.L = L

    def fr_norm(self, op, work_v, work_t) -> float:          # solver.rs:85-107
        L = self.L
        assert work_v.size == op.size()[1] and work_t.size == op.size()[0]
        L.scale(0.0, work_v)
        sq_norm = 0.0
        for row in range(op.size()[1]):
            work_v[row] = 1.0
            op.op(1.0, work_v, 0.0, work_t)
            n = L.norm(work_t)
            sq_norm = sq_norm + n * n
            work_v[row] = 0.0
        return math.sqrt(sq_norm)

    def op(self, alpha, x, beta, y):                         # solver.rs:109-131
        L = self.L
        m, n = self.a.size()
        assert x.size == n + m + m + 1 and y.size == n + m + 1
        x_x, x_y, x_s, x_tau = x[:n], x[n:n + m], x[n + m:n + 2 * m], x[n + 2 * m:]
        y_n, y_m, y_1 = y[:n], y[n:n + m], y[n + m:]
        self.a.trans_op(alpha, x_y, beta, y_n)
        self.c.op(alpha, x_tau, 1.0, y_n)
        self.a.op(-alpha, x_x, beta, y_m)
        L.add(-alpha, x_s, y_m)
        self.b.op(alpha, x_tau, 1.0, y_m)
        self.c.trans_op(-alpha, x_x, beta, y_1)
        self.b.trans_op(-alpha, x_y, 1.0, y_1)

    def trans_op(self, alpha, x, beta, y):                   # solver.rs:133-157
        L = self.L
        m, n = self.a.size()
        assert x.size == n + m + 1 and y.size == n + m + m + 1
        x_n, x_m, x_1 = x[:n], x[n:n + m], x[n + m:]
        y_x, y_y, y_s, y_tau = y[:n], y[n:n + m], y[n + m:n + 2 * m], y[n + 2 * m:]
        self.a.trans_op(-alpha, x_m, beta, y_x)
        self.c.op(-alpha, x_1, 1.0, y_x)
        self.a.op(alpha, x_n, beta, y_y)
        self.b.op(-alpha, x_1, 1.0, y_y)
        L.scale(beta, y_s)
        L.add(-alpha, x_m, y_s)
        self.c.trans_op(alpha, x_n, beta, y_tau)
        self.b.trans_op(alpha, x_m, 1.0, y_tau)

    def abssum(self, tau, sigma):                            # solver.rs:159-183
        L = self.L
        m, n = self.a.size()
        L.scale(0.0, tau)
        tau_x, tau_y, tau_s, tau_tau = tau[:n], tau[n:n + m], tau[n + m:n + 2 * m], tau[n + 2 * m:]
        self.a.absadd_cols(tau_x)
        self.c.absadd_rows(tau_x)
        self.a.absadd_rows(tau_y)
        self.b.absadd_rows(tau_y)
        L.adds(1.0, tau_s)
        self.c.absadd_cols(tau_tau)
        self.b.absadd_cols(tau_tau)
        sigma_n, sigma_m, sigma_1 = sigma[:n], sigma[n:n + m], sigma[n + m:]
        L.copy(tau_x, sigma_n)
        L.copy(tau_y, sigma_m)
        L.add(1.0, tau_s, sigma_m)
        L.copy(tau_tau, sigma_1)


class Solver:
    """First-order conic solver (HSDE + preconditioned Pock-Chambolle). solver.rs:216-657."""
    L = F64LAPACK

    def __init__(self, L=None):
        self.par = SolverParam()
        self.log: List[str] = []          # the `log::debug!` lines of solver.rs:391,426
        self.trace: Optional[list] = None  # if a list: (i, cri_pri, cri_dual, cri_gap) every iteration
        self.iters = 0
        self.snapshots: Optional[dict] = None   # {iteration: (x copy, y copy)} if a dict with wanted keys
        if L is not None:
            self.L = L

    @staticmethod
    def query_worklen(op_a_size) -> int:                     # solver.rs:231-249
        m, n = op_a_size
        return (n + m + m + 1) * 4 + (n + m + 1) * 2

    def set_par(self, fn) -> "Solver":                       # solver.rs:265-270
        fn(self.par)
        return self

    def solve(self, prob):                                   # solver.rs:285-321
        op_c, op_a, op_b, cone, work = prob
        m, n = op_a.size()
        if tuple(op_c.size()) != (n, 1) or tuple(op_b.size()) != (m, 1):
            raise SolverError(SolverError.InvalidOp)
        if self.query_worklen((m, n)) > work.size:
            raise SolverError(SolverError.WorkShortage)
        self.op_k = SelfDualEmbed(op_c, op_a, op_b, self.L)
        self.cone = cone
        self._core_solve(work)
        return work[:n], work[n:n + m]

    # -- SolverCore::solve, solver.rs:340-457
    def _core_solve(self, work):
        L = self.L
        par = self.par
        m, n = self.op_k.a.size()
        norm_b, norm_c = self._calc_norms(work)
        lx, ly = n + 2 * m + 1, n + m + 1
        o = 0
        x = work[o:o + lx]; o += lx
        y = work[o:o + ly]; o += ly
        dp_tau = work[o:o + lx]; o += lx
        dp_sigma = work[o:o + ly]; o += ly
        tmpw = work[o:o + 2 * lx]
        self._init_vecs(x, y)
        self._calc_precond(dp_tau, dp_sigma)
        self.norm_b, self.norm_c = norm_b, norm_c

        i = 0
        while True:
            excess_iter = (i + 1 >= par.max_iter) if par.max_iter is not None else False
            log_trig = (i % par.log_period == 0) if par.log_period > 0 else False

            val_tau = self._update_vecs(x, y, dp_tau, dp_sigma, tmpw)
            if self.snapshots is not None and (i + 1) in self.snapshots:
                self.snapshots[i + 1] = (x.copy(), y.copy())

            if val_tau > par.eps_zero:
                cri_pri, cri_dual, cri_gap = self._criteria_conv(x, norm_c, norm_b, tmpw)
                if self.trace is not None:
                    self.trace.append((i, cri_pri, cri_dual, cri_gap))
                term_conv = (cri_pri <= par.eps_acc) and (cri_dual <= par.eps_acc) and (cri_gap <= par.eps_acc)
                if log_trig or excess_iter or term_conv:
                    self.log.append("%d: pri_dual_gap %s %s %s" % (i, fmt_e2(cri_pri), fmt_e2(cri_dual), fmt_e2(cri_gap)))
                if excess_iter or term_conv:
                    L.scale(1.0 / val_tau, x[:n])
                    L.scale(1.0 / val_tau, x[n:n + m])
                    self.iters = i
                    if term_conv:
                        return
                    raise SolverError(SolverError.ExcessIter)
            else:
                cri_unbdd, cri_infeas = self._criteria_inf(x, norm_c, norm_b, tmpw)
                if self.trace is not None:
                    self.trace.append((i, cri_unbdd, cri_infeas, float("nan")))
                term_unbdd = cri_unbdd <= par.eps_inf
                term_infeas = cri_infeas <= par.eps_inf
                if log_trig or excess_iter or term_unbdd or term_infeas:
                    self.log.append("%d: unbdd_infeas %s %s" % (i, fmt_e2(cri_unbdd), fmt_e2(cri_infeas)))
                if excess_iter or term_unbdd or term_infeas:
                    self.iters = i
                    if term_unbdd:
                        raise SolverError(SolverError.Unbounded)
                    elif term_infeas:
                        raise SolverError(SolverError.Infeasible)
                    raise SolverError(SolverError.ExcessIter)
            i += 1
            assert not excess_iter

    def _calc_norms(self, work):                             # solver.rs:460-481
        work1 = np.zeros(1)
        m = self.op_k.b.size()[0]
        norm_b = self.op_k.fr_norm(self.op_k.b, work1, work[:m])
        n = self.op_k.c.size()[0]
        norm_c = self.op_k.fr_norm(self.op_k.c, work1, work[:n])
        return norm_b, norm_c

    def _init_vecs(self, x, y):                              # solver.rs:483-494
        m, n = self.op_k.a.size()
        self.L.scale(0.0, x)
        self.L.scale(0.0, y)
        x[n + m + m] = 1.0

    def _calc_precond(self, dp_tau, dp_sigma):               # solver.rs:496-524
        m, n = self.op_k.a.size()
        eps_zero = self.par.eps_zero
        self.op_k.abssum(dp_tau, dp_sigma)
        dp_tau[...] = 1.0 / np.maximum(dp_tau, eps_zero)
        dp_sigma[...] = 1.0 / np.maximum(dp_sigma, eps_zero)

        def group(tau_group):
            if tau_group.size > 0:
                tau_group[...] = tau_group.min()
        self.cone.product_group(dp_tau[n:n + m], group)
        self.cone.product_group(dp_tau[n + m:n + 2 * m], group)

    def _update_vecs(self, x, y, dp_tau, dp_sigma, tmpw):    # solver.rs:526-571
        L = self.L
        m, n = self.op_k.a.size()
        lx = x.size
        rx, tx = tmpw[:lx], tmpw[lx:2 * lx]
        L.copy(x, rx)
        self.op_k.trans_op(-1.0, y, 0.0, tx)
        L.transform_di(1.0, dp_tau, tx, 1.0, x)
        x_y, x_s, x_tau = x[n:n + m], x[n + m:n + 2 * m], x[n + 2 * m:]
        try:
            self.cone.proj(True, x_y)
            self.cone.proj(False, x_s)
        except SolverError:
            raise
        val_tau = max(float(x_tau[0]), 0.0)
        x_tau[0] = val_tau
        L.add(-2.0, x, rx)
        ty = tx[:y.size]
        self.op_k.op(-1.0, rx, 0.0, ty)
        L.transform_di(1.0, dp_sigma, ty, 1.0, y)
        kappa = min(float(y[n + m]), 0.0)
        y[n + m] = kappa
        return val_tau

    def _criteria_conv(self, x, norm_c, norm_b, tmpw):       # solver.rs:573-612
        L = self.L
        m, n = self.op_k.a.size()
        x_x, x_y, x_s, x_tau = x[:n], x[n:n + m], x[n + m:n + 2 * m], x[n + 2 * m:]
        p, d = tmpw[:m], tmpw[m:m + n]
        val_tau = float(x_tau[0])
        assert val_tau > 0.0
        work_one = np.ones(1)
        L.copy(x_s, p)
        self.op_k.b.op(-1.0, work_one, 1.0 / val_tau, p)
        self.op_k.a.op(1.0 / val_tau, x_x, 1.0, p)
        self.op_k.c.op(1.0, work_one, 0.0, d)
        self.op_k.a.trans_op(1.0 / val_tau, x_y, 1.0, d)
        self.op_k.c.trans_op(1.0 / val_tau, x_x, 0.0, work_one)
        g_x = float(work_one[0])
        self.op_k.b.trans_op(1.0 / val_tau, x_y, 0.0, work_one)
        g_y = float(work_one[0])
        g = g_x + g_y
        cri_pri = L.norm(p) / (1.0 + norm_b)
        cri_dual = L.norm(d) / (1.0 + norm_c)
        cri_gap = abs(g) / (1.0 + abs(g_x) + abs(g_y))
        return cri_pri, cri_dual, cri_gap

    def _criteria_inf(self, x, norm_c, norm_b, tmpw):        # solver.rs:614-656
        L = self.L
        m, n = self.op_k.a.size()
        x_x, x_y, x_s = x[:n], x[n:n + m], x[n + m:n + 2 * m]
        p, d = tmpw[:m], tmpw[m:m + n]
        work_one = np.zeros(1)
        L.copy(x_s, p)
        self.op_k.a.op(1.0, x_x, 1.0, p)
        self.op_k.a.trans_op(1.0, x_y, 0.0, d)
        self.op_k.c.trans_op(-1.0, x_x, 0.0, work_one)
        m_cx = float(work_one[0])
        self.op_k.b.trans_op(-1.0, x_y, 0.0, work_one)
        m_by = float(work_one[0])
        eps_zero = self.par.eps_zero
        cri_unbdd = L.norm(p) * norm_c / m_cx if m_cx > eps_zero else math.inf
        cri_infeas = L.norm(d) * norm_b / m_by if m_by > eps_zero else math.inf
        return cri_unbdd, cri_infeas


# --------------------------------------------------------------------------------------
# MatBuild (only what the front-ends need) — totsu/src/matbuild/mod.rs
# --------------------------------------------------------------------------------------
class MatBuild:
    L = F64LAPACK

    def __init__(self, typ: MatType, array=None):            # matbuild/mod.rs:24-31
        self.typ = typ
        self.array = np.zeros(typ.len()) if array is None else np.array(array, dtype=np.float64).reshape(-1)
        assert self.array.size == typ.len()

    def size(self):
        return self.typ.size()

    def is_sympack(self):
        return self.typ.kind == "SymPack"

    def index(self, r, c):                                   # matbuild/mod.rs:249-272
        if self.typ.kind == "General":
            nr, nc = self.typ.a, self.typ.b
            assert r < nr and c < nc
            return c * nr + r
        n = self.typ.a
        assert r < n and c < n
        if r > c:
            r, c = c, r
        return c * (c + 1) // 2 + r

    def __getitem__(self, rc):
        return self.array[self.index(*rc)]

    def __setitem__(self, rc, v):
        self.array[self.index(*rc)] = v

    def by_fn(self, func):                                   # matbuild/mod.rs:52-77
        if self.typ.kind == "General":
            nr, nc = self.typ.a, self.typ.b
            for c in range(nc):
                for r in range(nr):
                    self[(r, c)] = func(r, c)
        else:
            n = self.typ.a
            for c in range(n):
                for r in range(c + 1):
                    self[(r, c)] = func(r, c)
        return self

    def iter_colmaj(self, it):                               # matbuild/mod.rs:80-103
        it = iter(it)
        nr, nc = self.typ.size()
        for c in range(nc):
            for r in range(nr):
                try:
                    self[(r, c)] = next(it)
                except StopIteration:
                    return self
        return self

    def iter_rowmaj(self, it):                               # matbuild/mod.rs:106-129
        it = iter(it)
        nr, nc = self.typ.size()
        for r in range(nr):
            for c in range(nc):
                try:
                    self[(r, c)] = next(it)
                except StopIteration:
                    return self
        return self

    def scale_nondiag(self, alpha):                          # matbuild/mod.rs:143-175 (SymPack arm)
        assert self.typ.kind == "SymPack"
        n = self.typ.a
        for c in range(n - 1):
            i = self.index(c, c)
            ii = self.index(c + 1, c + 1)
            self.array[i + 1: ii] *= alpha
        return self

    def reshape_colvec(self):                                # matbuild/mod.rs:182-192
        self.typ = MatType.General(self.array.size, 1)
        return self

    def sqrt(self, eps_zero):                                # matbuild/mod.rs:220-247
        assert self.typ.kind == "SymPack"
        n = self.typ.a
        work = np.zeros(self.L.map_eig_worklen(n))
        self.L.map_eig(self.array, None, eps_zero, work,
                       lambda e: math.sqrt(e) if e > 0.0 else None)
        return self

    def as_op(self) -> MatOp:                                # matbuild/mod.rs:47-50
        return MatOp(self.typ, self.array)

    def clone(self):
        return MatBuild(self.typ, self.array.copy())


# --------------------------------------------------------------------------------------
# Front-ends — totsu/src/problem/{lp,qp,qcqp,socp,sdp}.rs
# Each `problem()` returns (op_c, op_a, op_b, cone, work) exactly like the reference.
# --------------------------------------------------------------------------------------
class _VecOp:
    """Wrapper giving a MatOp the Operator surface (ProbLPOpC etc. just forward)."""
    def __init__(self, m: MatOp):
        self.m = m

    def size(self):
        n, one = self.m.size()
        assert one == 1
        return (n, 1)

    def op(self, a, x, b, y): self.m.op(a, x, b, y)
    def trans_op(self, a, x, b, y): self.m.trans_op(a, x, b, y)
    def absadd_cols(self, t): self.m.absadd_cols(t)
    def absadd_rows(self, s): self.m.absadd_rows(s)


class _StackOp:
    """Vertical stack of MatOps with the same column count: ProbLPOpA (lp.rs:67-115),
    ProbLPOpB (lp.rs:137-188), ProbSDPOpA (sdp.rs:66-114) share this shape."""
    L = F64LAPACK

    def __init__(self, mats: Sequence[MatOp], signs: Optional[Sequence[float]] = None):
        self.mats = list(mats)
        self.signs = list(signs) if signs is not None else [1.0] * len(self.mats)
        self.ncol = self.mats[0].size()[1]

    def size(self):
        return (sum(mt.size()[0] for mt in self.mats), self.ncol)

    def op(self, alpha, x, beta, y):
        o = 0
        for mt, sg in zip(self.mats, self.signs):
            r = mt.size()[0]
            mt.op(sg * alpha, x, beta, y[o:o + r])
            o += r

    def trans_op(self, alpha, x, beta, y):
        o = 0
        for k, (mt, sg) in enumerate(zip(self.mats, self.signs)):
            r = mt.size()[0]
            mt.trans_op(sg * alpha, x[o:o + r], beta if k == 0 else 1.0, y)
            o += r

    def absadd_cols(self, tau):
        for mt in self.mats:
            mt.absadd_cols(tau)

    def absadd_rows(self, sigma):
        o = 0
        for mt in self.mats:
            r = mt.size()[0]
            mt.absadd_rows(sigma[o:o + r])
            o += r


class _ProductCone:
    """Sequence of (cone, length) blocks: ProbLPCone lp.rs:190-219, ProbQPCone qp.rs:260-296,
    ProbQCQPCone qcqp.rs:303-350, ProbSOCPCone socp.rs:286-333, ProbSDPCone sdp.rs:188-220."""
    def __init__(self, blocks):
        self.blocks = blocks

    def proj(self, dual_cone, x):
        o = 0
        for cone, ln in self.blocks:
            cone.proj(dual_cone, x[o:o + ln])
            o += ln

    def product_group(self, dp_tau, group):
        o = 0
        for cone, ln in self.blocks:
            cone.product_group(dp_tau[o:o + ln], group)
            o += ln


class ProbLP:                                                # lp.rs:222-338
    def __init__(self, vec_c, mat_g, vec_h, mat_a, vec_b):
        n, m, p = vec_c.size()[0], vec_h.size()[0], vec_b.size()[0]
        assert vec_c.size() == (n, 1) and mat_g.size() == (m, n) and vec_h.size() == (m, 1)
        assert mat_a.size() == (p, n) and vec_b.size() == (p, 1)
        self.vec_c, self.mat_g, self.vec_h, self.mat_a, self.vec_b = vec_c, mat_g, vec_h, mat_a, vec_b

    def problem(self):
        m, p = self.vec_h.size()[0], self.vec_b.size()[0]
        op_c = _VecOp(self.vec_c.as_op())
        op_a = _StackOp([self.mat_g.as_op(), self.mat_a.as_op()])
        op_b = _StackOp([self.vec_h.as_op(), self.vec_b.as_op()])
        cone = _ProductCone([(ConeRPos(), m), (ConeZero(), p)])
        work = np.zeros(Solver.query_worklen(op_a.size()))
        return op_c, op_a, op_b, cone, work


class _QPOpC:                                                # qp.rs:10-61 / qcqp.rs:10-61
    L = F64LAPACK

    def __init__(self, n): self.n = n
    def size(self): return (self.n + 1, 1)

    def op(self, alpha, x, beta, y):
        n = self.n
        self.L.scale(beta, y[:n])
        self.L.scale(beta, y[n:n + 1])
        self.L.add(alpha, x, y[n:n + 1])

    def trans_op(self, alpha, x, beta, y):
        n = self.n
        self.L.scale(beta, y)
        self.L.add(alpha, x[n:n + 1], y)

    def absadd_cols(self, tau): tau[0] = tau[0] + 1.0
    def absadd_rows(self, sigma): sigma[self.n] = sigma[self.n] + 1.0


class _QPOpA:                                                # qp.rs:65-170
    L = F64LAPACK

    def __init__(self, sym_p_sqrt, vec_q, mat_g, mat_a):
        self.sym_p_sqrt, self.vec_q, self.mat_g, self.mat_a = sym_p_sqrt, vec_q, mat_g, mat_a

    def dim(self):
        n = self.sym_p_sqrt.size()[0]
        return n, self.mat_g.size()[0], self.mat_a.size()[0]

    def size(self):
        n, m, p = self.dim()
        return ((2 + n) + m + p, n + 1)

    def op(self, alpha, x, beta, y):
        L = self.L
        n, m, p = self.dim()
        x_n, x_t = x[:n], x[n:n + 1]
        y_r, y_s, y_n, y_m, y_p = y[0:1], y[1:2], y[2:2 + n], y[2 + n:2 + n + m], y[2 + n + m:]
        L.scale(beta, y_r)
        self.vec_q.trans_op(alpha, x_n, beta, y_s)
        L.add(-alpha, x_t, y_s)
        self.sym_p_sqrt.op(-alpha, x_n, beta, y_n)
        self.mat_g.op(alpha, x_n, beta, y_m)
        self.mat_a.op(alpha, x_n, beta, y_p)

    def trans_op(self, alpha, x, beta, y):
        L = self.L
        n, m, p = self.dim()
        x_s, x_n, x_m, x_p = x[1:2], x[2:2 + n], x[2 + n:2 + n + m], x[2 + n + m:]
        y_n, y_t = y[:n], y[n:n + 1]
        self.vec_q.op(alpha, x_s, beta, y_n)
        self.sym_p_sqrt.op(-alpha, x_n, 1.0, y_n)
        self.mat_g.trans_op(alpha, x_m, 1.0, y_n)
        self.mat_a.trans_op(alpha, x_p, 1.0, y_n)
        L.scale(beta, y_t)
        L.add(-alpha, x_s, y_t)

    def absadd_cols(self, tau):
        n, _, _ = self.dim()
        tau_n = tau[:n]
        self.vec_q.absadd_rows(tau_n)
        self.sym_p_sqrt.absadd_cols(tau_n)
        self.mat_g.absadd_cols(tau_n)
        self.mat_a.absadd_cols(tau_n)
        tau[n] = tau[n] + 1.0

    def absadd_rows(self, sigma):
        n, m, p = self.dim()
        self.vec_q.absadd_cols(sigma[1:2])
        sigma[1] = sigma[1] + 1.0
        self.sym_p_sqrt.absadd_rows(sigma[2:2 + n])
        self.mat_g.absadd_rows(sigma[2 + n:2 + n + m])
        self.mat_a.absadd_rows(sigma[2 + n + m:])


class _QPOpB:                                                # qp.rs:174-258
    L = F64LAPACK

    def __init__(self, n, vec_h, vec_b):
        self.n, self.vec_h, self.vec_b = n, vec_h, vec_b

    def dim(self):
        return self.n, self.vec_h.size()[0], self.vec_b.size()[0]

    def size(self):
        n, m, p = self.dim()
        return ((2 + n) + m + p, 1)

    def op(self, alpha, x, beta, y):
        L = self.L
        n, m, p = self.dim()
        y_r, y_sn, y_m, y_p = y[0:1], y[1:2 + n], y[2 + n:2 + n + m], y[2 + n + m:]
        L.scale(beta, y_r)
        L.add(alpha, x, y_r)
        L.scale(beta, y_sn)
        self.vec_h.op(alpha, x, beta, y_m)
        self.vec_b.op(alpha, x, beta, y_p)

    def trans_op(self, alpha, x, beta, y):
        L = self.L
        n, m, p = self.dim()
        x_r, x_m, x_p = x[0:1], x[2 + n:2 + n + m], x[2 + n + m:]
        self.vec_h.trans_op(alpha, x_m, beta, y)
        self.vec_b.trans_op(alpha, x_p, 1.0, y)
        L.add(alpha, x_r, y)

    def absadd_cols(self, tau):
        tau[0] = tau[0] + 1.0
        self.vec_h.absadd_cols(tau)
        self.vec_b.absadd_cols(tau)

    def absadd_rows(self, sigma):
        n, m, p = self.dim()
        sigma[0] = sigma[0] + 1.0
        self.vec_h.absadd_rows(sigma[2 + n:2 + n + m])
        self.vec_b.absadd_rows(sigma[2 + n + m:])


class ProbQP:                                                # qp.rs:300-437
    def __init__(self, sym_p, vec_q, mat_g, vec_h, mat_a, vec_b, eps_zero, p_is_sqrt=False):
        """p_is_sqrt: `sym_p` already holds P^(1/2) (bench shortcut for config C2's diagonal P, where the n = 8192
        eigendecomposition of qp.rs:386 `set_sqrt` would take minutes and is not part of the timed iteration)."""
        n, m, p = vec_q.size()[0], vec_h.size()[0], vec_b.size()[0]
        assert sym_p.is_sympack() and sym_p.size() == (n, n)
        assert mat_g.size() == (m, n) and mat_a.size() == (p, n)
        self.vec_q, self.mat_g, self.vec_h, self.mat_a, self.vec_b = vec_q, mat_g, vec_h, mat_a, vec_b
        self.sym_p_sqrt = sym_p if p_is_sqrt else sym_p.sqrt(eps_zero)

    def problem(self):
        n, m, p = self.vec_q.size()[0], self.vec_h.size()[0], self.vec_b.size()[0]
        op_c = _QPOpC(n)
        op_a = _QPOpA(self.sym_p_sqrt.as_op(), self.vec_q.as_op(), self.mat_g.as_op(), self.mat_a.as_op())
        op_b = _QPOpB(n, self.vec_h.as_op(), self.vec_b.as_op())
        cone = _ProductCone([(ConeRotSOC(), 2 + n), (ConeRPos(), m), (ConeZero(), p)])
        work = np.zeros(Solver.query_worklen(op_a.size()))
        return op_c, op_a, op_b, cone, work


class _QCQPOpA:                                              # qcqp.rs:65-193
    L = F64LAPACK

    def __init__(self, syms_p_sqrt, vecs_q, mat_a):
        self.syms_p_sqrt, self.vecs_q, self.mat_a = syms_p_sqrt, vecs_q, mat_a

    def dim(self):
        p, n = self.mat_a.size()
        return n, len(self.syms_p_sqrt), p

    def size(self):
        n, m1, p = self.dim()
        return (m1 * (2 + n) + p, n + 1)

    def op(self, alpha, x, beta, y):
        L = self.L
        n, m1, p = self.dim()
        x_n, x_t = x[:n], x[n:n + 1]
        for i, (sp, vq) in enumerate(zip(self.syms_p_sqrt, self.vecs_q)):
            o = i * (2 + n)
            y_r, y_s, y_n = y[o:o + 1], y[o + 1:o + 2], y[o + 2:o + 2 + n]
            L.scale(beta, y_r)
            vq.trans_op(alpha, x_n, beta, y_s)
            if i == 0:
                L.add(-alpha, x_t, y_s)
            sp.op(-alpha, x_n, beta, y_n)
        self.mat_a.op(alpha, x_n, beta, y[m1 * (2 + n):])

    def trans_op(self, alpha, x, beta, y):
        L = self.L
        n, m1, p = self.dim()
        y_n, y_t = y[:n], y[n:n + 1]
        L.scale(beta, y_n)
        L.scale(beta, y_t)
        for i, (sp, vq) in enumerate(zip(self.syms_p_sqrt, self.vecs_q)):
            o = i * (2 + n)
            x_s, x_n = x[o + 1:o + 2], x[o + 2:o + 2 + n]
            vq.op(alpha, x_s, 1.0, y_n)
            sp.op(-alpha, x_n, 1.0, y_n)
            if i == 0:
                L.add(-alpha, x_s, y_t)
        self.mat_a.trans_op(alpha, x[m1 * (2 + n):], 1.0, y_n)

    def absadd_cols(self, tau):
        n, m1, p = self.dim()
        tau_n = tau[:n]
        for vq in self.vecs_q:
            vq.absadd_rows(tau_n)
        for sp in self.syms_p_sqrt:
            sp.absadd_cols(tau_n)
        self.mat_a.absadd_cols(tau_n)
        tau[n] = tau[n] + 1.0

    def absadd_rows(self, sigma):
        n, m1, p = self.dim()
        for i, (sp, vq) in enumerate(zip(self.syms_p_sqrt, self.vecs_q)):
            o = i * (2 + n)
            vq.absadd_cols(sigma[o + 1:o + 2])
            if i == 0:
                sigma[o + 1] = sigma[o + 1] + 1.0
            sp.absadd_rows(sigma[o + 2:o + 2 + n])
        self.mat_a.absadd_rows(sigma[m1 * (2 + n):])


class _QCQPOpB:                                              # qcqp.rs:197-299
    L = F64LAPACK

    def __init__(self, n, scls_r, vec_b):
        self.n, self.scls_r, self.vec_b = n, np.asarray(scls_r, dtype=np.float64), vec_b
        self.abssum_scls_r = self.L.abssum(self.scls_r, 1)

    def dim(self):
        return self.n, len(self.scls_r), self.vec_b.size()[0]

    def size(self):
        n, m1, p = self.dim()
        return (m1 * (2 + n) + p, 1)

    def op(self, alpha, x, beta, y):
        L = self.L
        n, m1, p = self.dim()
        for i, r in enumerate(self.scls_r):
            o = i * (2 + n)
            y_r, y_s, y_n = y[o:o + 1], y[o + 1:o + 2], y[o + 2:o + 2 + n]
            L.scale(beta, y_r); L.add(alpha, x, y_r)
            L.scale(beta, y_s); L.add(-alpha * r, x, y_s)
            L.scale(beta, y_n)
        self.vec_b.op(alpha, x, beta, y[m1 * (2 + n):])

    def trans_op(self, alpha, x, beta, y):
        L = self.L
        n, m1, p = self.dim()
        L.scale(beta, y)
        for i, r in enumerate(self.scls_r):
            o = i * (2 + n)
            L.add(alpha, x[o:o + 1], y)
            L.add(-alpha * r, x[o + 1:o + 2], y)
        self.vec_b.trans_op(alpha, x[m1 * (2 + n):], 1.0, y)

    def absadd_cols(self, tau):
        n, m1, p = self.dim()
        tau[0] = tau[0] + float(m1) + self.abssum_scls_r
        self.vec_b.absadd_cols(tau)

    def absadd_rows(self, sigma):
        n, m1, p = self.dim()
        for i, r in enumerate(self.scls_r):
            o = i * (2 + n)
            sigma[o] = sigma[o] + 1.0
            sigma[o + 1] = sigma[o + 1] + abs(r)
        self.vec_b.absadd_rows(sigma[m1 * (2 + n):])


class ProbQCQP:                                              # qcqp.rs:354-481
    def __init__(self, syms_p, vecs_q, scls_r, mat_a, vec_b, eps_zero):
        p, n = mat_a.size()
        m1 = len(syms_p)
        assert len(vecs_q) == m1 and len(scls_r) == m1
        self.vecs_q, self.scls_r, self.mat_a, self.vec_b = vecs_q, list(scls_r), mat_a, vec_b
        self.syms_p_sqrt = [sp.sqrt(eps_zero) for sp in syms_p]

    def problem(self):
        p, n = self.mat_a.size()
        m1 = len(self.syms_p_sqrt)
        op_c = _QPOpC(n)
        op_a = _QCQPOpA([s.as_op() for s in self.syms_p_sqrt], [q.as_op() for q in self.vecs_q], self.mat_a.as_op())
        op_b = _QCQPOpB(n, self.scls_r, self.vec_b.as_op())
        cone = _ProductCone([(ConeRotSOC(), 2 + n)] * m1 + [(ConeZero(), p)])
        work = np.zeros(Solver.query_worklen(op_a.size()))
        return op_c, op_a, op_b, cone, work


class _SOCPOpA:                                              # socp.rs:47-165
    L = F64LAPACK

    def __init__(self, mats_g, vecs_c, mat_a):
        self.mats_g, self.vecs_c, self.mat_a = mats_g, vecs_c, mat_a

    def size(self):
        p, n = self.mat_a.size()
        return (sum(1 + g.size()[0] for g in self.mats_g) + p, n)

    def op(self, alpha, x, beta, y):
        done = 0
        for g, c in zip(self.mats_g, self.vecs_c):
            ni = g.size()[0]
            c.trans_op(-alpha, x, beta, y[done:done + 1])
            g.op(-alpha, x, beta, y[done + 1:done + 1 + ni])
            done += 1 + ni
        self.mat_a.op(alpha, x, beta, y[done:])

    def trans_op(self, alpha, x, beta, y):
        self.L.scale(beta, y)
        done = 0
        for g, c in zip(self.mats_g, self.vecs_c):
            ni = g.size()[0]
            c.op(-alpha, x[done:done + 1], 1.0, y)
            g.trans_op(-alpha, x[done + 1:done + 1 + ni], 1.0, y)
            done += 1 + ni
        self.mat_a.trans_op(alpha, x[done:], 1.0, y)

    def absadd_cols(self, tau):
        for c in self.vecs_c:
            c.absadd_rows(tau)
        for g in self.mats_g:
            g.absadd_cols(tau)
        self.mat_a.absadd_cols(tau)

    def absadd_rows(self, sigma):
        done = 0
        for g, c in zip(self.mats_g, self.vecs_c):
            ni = g.size()[0]
            c.absadd_cols(sigma[done:done + 1])
            g.absadd_rows(sigma[done + 1:done + 1 + ni])
            done += 1 + ni
        self.mat_a.absadd_rows(sigma[done:])


class _SOCPOpB:                                              # socp.rs:169-282
    L = F64LAPACK

    def __init__(self, vecs_h, scls_d, vec_b):
        self.vecs_h, self.scls_d, self.vec_b = vecs_h, np.asarray(scls_d, dtype=np.float64), vec_b
        self.abssum_scls_d = self.L.abssum(self.scls_d, 1)

    def size(self):
        p = self.vec_b.size()[0]
        return (sum(1 + h.size()[0] for h in self.vecs_h) + p, 1)

    def op(self, alpha, x, beta, y):
        L = self.L
        done = 0
        for h, d in zip(self.vecs_h, self.scls_d):
            ni = h.size()[0]
            y_1 = y[done:done + 1]
            L.scale(beta, y_1)
            L.add(alpha * d, x, y_1)
            h.op(alpha, x, beta, y[done + 1:done + 1 + ni])
            done += 1 + ni
        self.vec_b.op(alpha, x, beta, y[done:])

    def trans_op(self, alpha, x, beta, y):
        L = self.L
        L.scale(beta, y)
        done = 0
        for h, d in zip(self.vecs_h, self.scls_d):
            ni = h.size()[0]
            L.add(alpha * d, x[done:done + 1], y)
            h.trans_op(alpha, x[done + 1:done + 1 + ni], 1.0, y)
            done += 1 + ni
        self.vec_b.trans_op(alpha, x[done:], 1.0, y)

    def absadd_cols(self, tau):
        tau[0] = tau[0] + self.abssum_scls_d
        for h in self.vecs_h:
            h.absadd_cols(tau)
        self.vec_b.absadd_cols(tau)

    def absadd_rows(self, sigma):
        done = 0
        for h, d in zip(self.vecs_h, self.scls_d):
            ni = h.size()[0]
            sigma[done] = sigma[done] + d          # NOTE: the reference adds scl_d itself, not |scl_d| (socp.rs:272)
            h.absadd_rows(sigma[done + 1:done + 1 + ni])
            done += 1 + ni
        self.vec_b.absadd_rows(sigma[done:])


class ProbSOCP:                                              # socp.rs:337-472
    def __init__(self, vec_f, mats_g, vecs_h, vecs_c, scls_d, mat_a, vec_b):
        n, m, p = vec_f.size()[0], len(mats_g), vec_b.size()[0]
        assert len(vecs_h) == m and len(vecs_c) == m and len(scls_d) == m
        for i in range(m):
            ni = mats_g[i].size()[0]
            assert mats_g[i].size() == (ni, n) and vecs_h[i].size() == (ni, 1) and vecs_c[i].size() == (n, 1)
        assert mat_a.size() == (p, n) and vec_b.size() == (p, 1)
        self.vec_f, self.mats_g, self.vecs_h, self.vecs_c = vec_f, mats_g, vecs_h, vecs_c
        self.scls_d, self.mat_a, self.vec_b = list(scls_d), mat_a, vec_b

    def problem(self):
        p = self.vec_b.size()[0]
        op_c = _VecOp(self.vec_f.as_op())
        op_a = _SOCPOpA([g.as_op() for g in self.mats_g], [c.as_op() for c in self.vecs_c], self.mat_a.as_op())
        op_b = _SOCPOpB([h.as_op() for h in self.vecs_h], self.scls_d, self.vec_b.as_op())
        cone = _ProductCone([(ConeSOC(), 1 + g.size()[0]) for g in self.mats_g] + [(ConeZero(), p)])
        work = np.zeros(Solver.query_worklen(op_a.size()))
        return op_c, op_a, op_b, cone, work


class ProbSDP:                                               # sdp.rs:224-365
    def __init__(self, vec_c, syms_f, mat_a, vec_b, eps_zero):
        n, p = vec_c.size()[0], vec_b.size()[0]
        assert len(syms_f) == n + 1
        k = syms_f[0].size()[0]
        for s in syms_f:
            assert s.is_sympack() and s.size() == (k, k)
        fsqrt2 = math.sqrt(2.0)
        syms_f = [s.clone().scale_nondiag(fsqrt2).reshape_colvec() for s in syms_f]
        self.symvec_f_n = syms_f.pop()
        sk = self.symvec_f_n.size()[0]
        self.symmat_f = MatBuild(MatType.General(sk, n))
        for c in range(n):
            self.symmat_f.array[c * sk:(c + 1) * sk] = syms_f[c].array
        self.vec_c, self.mat_a, self.vec_b, self.eps_zero = vec_c, mat_a, vec_b, eps_zero

    def problem(self):
        p = self.vec_b.size()[0]
        sk = self.symvec_f_n.size()[0]
        op_c = _VecOp(self.vec_c.as_op())
        op_a = _StackOp([self.symmat_f.as_op(), self.mat_a.as_op()])
        op_b = _StackOp([self.symvec_f_n.as_op(), self.vec_b.as_op()], signs=[-1.0, 1.0])
        w_cone = np.zeros(ConePSD.query_worklen(sk))
        cone = _ProductCone([(ConePSD(w_cone, self.eps_zero), sk), (ConeZero(), p)])
        work = np.zeros(Solver.query_worklen(op_a.size()))
        return op_c, op_a, op_b, cone, work
