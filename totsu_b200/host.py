"""ctypes binding of libtotsu_b200_host.so (totsu_b200/host/driver.cpp): sessions that run the C++ mirror of
`totsu_core::solver::Solver` on the B200 backend.  Harness only - used by tests/ and bench.py."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi
from .capi import View, ConeBlock, TotsuB200Error, dtype_id

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libtotsu_b200_host.so")

ERRORS = ["None", "Unbounded", "Infeasible", "ExcessIter", "InvalidOp", "WorkShortage", "ConeFailure"]


class Param(C.Structure):
    _fields_ = [("max_iter", C.c_int64), ("eps_acc", C.c_double), ("eps_inf", C.c_double), ("eps_zero", C.c_double),
                ("log_period", C.c_uint64), ("device_precond", C.c_int32), ("reserved", C.c_int32)]


class Iter(C.Structure):
    _fields_ = [("i", C.c_uint64), ("conv_branch", C.c_int32), ("logged", C.c_int32),
                ("val_tau", C.c_double), ("c0", C.c_double), ("c1", C.c_double), ("c2", C.c_double)]


_hlib = None


def hlib():
    global _hlib
    if _hlib is None:
        capi.lib()   # libtotsu_b200.so first (RTLD_GLOBAL)
        if not os.path.exists(HOST_LIB_PATH):
            raise TotsuB200Error(f"{HOST_LIB_PATH} is missing: run __graft_entry__.build()")
        L = C.CDLL(HOST_LIB_PATH)
        vp, sz, i, d = C.c_void_p, C.c_size_t, C.c_int, C.c_double
        L.tbh_last_error.restype = C.c_char_p
        L.tbh_session_lp.restype = vp
        L.tbh_session_lp.argtypes = [i, sz, sz, sz, vp, vp, vp, vp, vp]
        L.tbh_session_qp.restype = vp
        L.tbh_session_qp.argtypes = [i, sz, sz, sz, vp, vp, vp, vp, vp, vp, d, i]
        L.tbh_session_qcqp.restype = vp
        L.tbh_session_qcqp.argtypes = [i, sz, sz, sz, vp, vp, vp, vp, vp, d]
        L.tbh_session_socp.restype = vp
        L.tbh_session_socp.argtypes = [i, sz, sz, C.POINTER(C.c_uint64), sz, vp, vp, vp, vp, vp, vp, vp]
        L.tbh_session_sdp.restype = vp
        L.tbh_session_sdp.argtypes = [i, sz, sz, sz, vp, vp, vp, vp, d]
        L.tbh_session_dense.restype = vp
        L.tbh_session_dense.argtypes = [i, sz, sz, View, sz, sz, vp, vp, C.POINTER(ConeBlock), sz, i, i, d]
        L.tbh_session_begin.argtypes = [vp, C.POINTER(Param)]
        L.tbh_session_step.argtypes = [vp, C.c_uint64, C.POINTER(Iter), C.POINTER(i)]
        L.tbh_session_run.argtypes = [vp, C.POINTER(Iter), sz, C.POINTER(sz), i, C.POINTER(Iter)]
        L.tbh_session_end.argtypes = [vp]
        L.tbh_session_xy.argtypes = [vp, vp, vp]
        L.tbh_session_solution.argtypes = [vp, vp, vp]
        L.tbh_session_dims.argtypes = [vp, C.POINTER(sz), C.POINTER(sz)]
        L.tbh_session_norms.argtypes = [vp, C.POINTER(d), C.POINTER(d)]
        L.tbh_session_destroy.argtypes = [vp]
        L.tbh_session_destroy.restype = None
        L.tbh_set_shim_protocol.argtypes = [i]
        L.tbh_set_shim_protocol.restype = None
        L.tbh_get_shim_protocol.restype = i
        L.tbh_set_map_eig_fast_paths.argtypes = [i]
        L.tbh_set_map_eig_fast_paths.restype = None
        _hlib = L
    return _hlib


def set_shim_protocol(on):
    """Drive the backend with the Rust binding's call protocol (see totsu_b200/host/linalg.hpp); set before creating sessions."""
    hlib().tbh_set_shim_protocol(1 if on else 0)


def set_map_eig_fast_paths(on):
    """B200::map_eig recognises the reference's two closures (ConePSD::proj, MatBuild::set_sqrt) and takes the GEMM-only paths."""
    hlib().tbh_set_map_eig_fast_paths(1 if on else 0)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def _arr(a, dt, order="C"):
    return np.ascontiguousarray(np.asarray(a, dtype=dt).reshape(-1, order=order))


def fmt_e2(v):
    """Rust's `{:.2e}` so traces diff against examples/nostd_cortex-m/log_qemu.txt."""
    if v == 0.0:
        return "0.00e0"
    m, e = ("%.2e" % v).split("e")
    return "%se%d" % (m, int(e))


class Session:
    """One `Solver::solve` call, steppable.  Mirrors `s.solve(prob.problem())` of the reference's tests."""

    def __init__(self, handle, dtype, keep=()):
        if not handle:
            raise TotsuB200Error("session creation failed: " + hlib().tbh_last_error().decode())
        self.h = handle
        self.dtype = np.dtype(dtype)
        self._keep = keep
        m, n = C.c_size_t(), C.c_size_t()
        hlib().tbh_session_dims(self.h, C.byref(m), C.byref(n))
        self.m, self.n = m.value, n.value
        self.last = Iter()

    # ---- constructors -------------------------------------------------------------------------------
    @staticmethod
    def lp(dtype, vec_c, mat_g, vec_h, mat_a=None, vec_b=None):
        dt = np.dtype(dtype)
        c = _arr(vec_c, dt); n = c.size
        h = _arr(vec_h, dt); m = h.size
        g = _arr(np.asarray(mat_g, dtype=dt).reshape(m, n), dt, "F")
        b = _arr(vec_b if vec_b is not None else [], dt); p = b.size
        a = _arr(np.asarray(mat_a if mat_a is not None else np.zeros((0, n)), dtype=dt).reshape(p, n), dt, "F")
        hd = hlib().tbh_session_lp(dtype_id(dt), n, m, p, _p(c), _p(g), _p(h), _p(a), _p(b))
        return Session(hd, dt, (c, g, h, a, b))

    @staticmethod
    def qp(dtype, sym_p_packed, vec_q, mat_g, vec_h, mat_a, vec_b, eps_zero, p_is_sqrt=False, col_major=False):
        """p_is_sqrt: `sym_p_packed` already holds P^(1/2) (qp.rs:386 set_sqrt skipped); col_major: mat_g / mat_a are
        given as flat column-major arrays (no transposing copy of a large G)."""
        dt = np.dtype(dtype)
        q = _arr(vec_q, dt); n = q.size
        h = _arr(vec_h, dt); m = h.size
        b = _arr(vec_b, dt); p = b.size
        sp = _arr(sym_p_packed, dt)
        if col_major:
            g, a = _arr(mat_g, dt), _arr(mat_a, dt)
            assert g.size == m * n and a.size == p * n
        else:
            g = _arr(np.asarray(mat_g, dtype=dt).reshape(m, n), dt, "F")
            a = _arr(np.asarray(mat_a, dtype=dt).reshape(p, n), dt, "F")
        hd = hlib().tbh_session_qp(dtype_id(dt), n, m, p, _p(sp), _p(q), _p(g), _p(h), _p(a), _p(b), eps_zero, 1 if p_is_sqrt else 0)
        return Session(hd, dt, (sp, q, g, h, a, b))

    @staticmethod
    def qcqp(dtype, syms_p_packed, vecs_q, scls_r, mat_a, vec_b, eps_zero):
        dt = np.dtype(dtype)
        vq = np.asarray(vecs_q, dtype=dt); m1, n = vq.shape
        sp = _arr(np.asarray(syms_p_packed, dtype=dt).reshape(m1, n * (n + 1) // 2), dt)
        vq = _arr(vq, dt)
        r = _arr(scls_r, dt)
        b = _arr(vec_b, dt); p = b.size
        a = _arr(np.asarray(mat_a, dtype=dt).reshape(p, n), dt, "F")
        hd = hlib().tbh_session_qcqp(dtype_id(dt), n, m1, p, _p(sp), _p(vq), _p(r), _p(a), _p(b), eps_zero)
        return Session(hd, dt, (sp, vq, r, a, b))

    @staticmethod
    def socp(dtype, vec_f, mats_g, vecs_h, vecs_c, scls_d, mat_a=None, vec_b=None):
        dt = np.dtype(dtype)
        f = _arr(vec_f, dt); n = f.size
        mats_g = [np.asarray(gm, dtype=dt) for gm in mats_g]     # each (ni, n)
        for gm in mats_g:
            assert gm.ndim == 2 and gm.shape[1] == n
        ni = np.array([gm.shape[0] for gm in mats_g], dtype=np.uint64)
        g = np.concatenate([_arr(gm, dt, "F") for gm in mats_g]) if len(mats_g) else np.zeros(0, dt)
        h = np.concatenate([_arr(hv, dt) for hv in vecs_h]) if len(vecs_h) else np.zeros(0, dt)
        c = np.concatenate([_arr(cv, dt) for cv in vecs_c]) if len(vecs_c) else np.zeros(0, dt)
        d = _arr(scls_d, dt)
        b = _arr(vec_b if vec_b is not None else [], dt); p = b.size
        a = _arr(np.asarray(mat_a if mat_a is not None else np.zeros((0, n)), dtype=dt).reshape(p, n), dt, "F")
        g = np.ascontiguousarray(g); h = np.ascontiguousarray(h); c = np.ascontiguousarray(c)
        hd = hlib().tbh_session_socp(dtype_id(dt), n, len(ni), ni.ctypes.data_as(C.POINTER(C.c_uint64)), p,
                                     _p(f), _p(g), _p(h), _p(c), _p(d), _p(a), _p(b))
        return Session(hd, dt, (f, ni, g, h, c, d, a, b))

    @staticmethod
    def sdp(dtype, vec_c, syms_f_packed, mat_a, vec_b, eps_zero):
        dt = np.dtype(dtype)
        c = _arr(vec_c, dt); n = c.size
        sf = np.asarray(syms_f_packed, dtype=dt); assert sf.shape[0] == n + 1
        sk = sf.shape[1]
        k = int((np.sqrt(8 * sk + 1) - 1) / 2 + 0.5)
        sf = _arr(sf, dt)
        b = _arr(vec_b, dt); p = b.size
        a = _arr(np.asarray(mat_a, dtype=dt).reshape(p, n), dt, "F")
        hd = hlib().tbh_session_sdp(dtype_id(dt), n, k, p, _p(c), _p(sf), _p(a), _p(b), eps_zero)
        return Session(hd, dt, (c, sf, a, b))

    @staticmethod
    def dense(dtype, a_view, m_local, n, vec_c, vec_b, blocks, fused_op=True, fused_cone=True, eps_zero=1e-12,
              row_offset=0, m_total=0, keep=()):
        """(op_c, op_a, op_b, cone) over one dense A held in a backend buffer (view `a_view`, column-major)."""
        dt = np.dtype(dtype)
        c = _arr(vec_c, dt)
        b = _arr(vec_b, dt)
        blk = (ConeBlock * len(blocks))(*[ConeBlock(t, 0, ln) for t, ln in blocks])
        hd = hlib().tbh_session_dense(dtype_id(dt), m_local, n, a_view, row_offset, m_total, _p(c), _p(b), blk, len(blocks),
                                      1 if fused_op else 0, 1 if fused_cone else 0, eps_zero)
        return Session(hd, dt, (c, b, blk) + tuple(keep))

    # ---- solver ---------------------------------------------------------------------------------------
    def _chk(self, st):
        if st < 0:
            raise TotsuB200Error("host driver: " + hlib().tbh_last_error().decode())
        return ERRORS[st]

    def begin(self, max_iter=None, eps_acc=1e-6, eps_inf=1e-6, eps_zero=1e-12, log_period=10_000, device_precond=False):
        p = Param(-1 if max_iter is None else int(max_iter), eps_acc, eps_inf, eps_zero, log_period, 1 if device_precond else 0, 0)
        return self._chk(hlib().tbh_session_begin(self.h, C.byref(p)))

    def step(self, k=1):
        done = C.c_int(0)
        st = self._chk(hlib().tbh_session_step(self.h, k, C.byref(self.last), C.byref(done)))
        return st, bool(done.value)

    def run(self, trace_cap=0, only_logged=True):
        """Iterate to termination; returns (status, [Iter...])."""
        buf = (Iter * trace_cap)() if trace_cap else None
        cnt = C.c_size_t(0)
        st = self._chk(hlib().tbh_session_run(self.h, buf, trace_cap, C.byref(cnt), 1 if only_logged else 0, C.byref(self.last)))
        return st, [buf[i] for i in range(cnt.value)] if trace_cap else []

    def solve(self, **par):
        """`Solver::solve`: returns (status, x, y)."""
        st = self.begin(**par)
        if st != "None":
            return st, None, None
        st, _ = self.run()
        self.end()
        x, y = self.solution()
        return st, x, y

    def end(self):
        self._chk(hlib().tbh_session_end(self.h))

    def xy(self):
        x = np.empty(self.n + 2 * self.m + 1, dtype=self.dtype)
        y = np.empty(self.n + self.m + 1, dtype=self.dtype)
        self._chk(hlib().tbh_session_xy(self.h, _p(x), _p(y)))
        return x, y

    def solution(self):
        x = np.empty(self.n, dtype=self.dtype)
        y = np.empty(self.m, dtype=self.dtype)
        self._chk(hlib().tbh_session_solution(self.h, _p(x), _p(y)))
        return x, y

    def norms(self):
        nb, nc = C.c_double(), C.c_double()
        hlib().tbh_session_norms(self.h, C.byref(nb), C.byref(nc))
        return nb.value, nc.value

    def close(self):
        if self.h:
            hlib().tbh_session_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def log_lines(iters):
    """Format trace records like solver.rs:391 / :426."""
    out = []
    for it in iters:
        if it.conv_branch:
            out.append("%d: pri_dual_gap %s %s %s" % (it.i, fmt_e2(it.c0), fmt_e2(it.c1), fmt_e2(it.c2)))
        else:
            out.append("%d: unbdd_infeas %s %s" % (it.i, fmt_e2(it.c0), fmt_e2(it.c1)))
    return out
