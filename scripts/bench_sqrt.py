#!/usr/bin/env python
"""MatBuild::set_sqrt at real size (totsu/src/matbuild/mod.rs:220-241, qp.rs:386): tb_sqrt_psd_f32 - the
GEMM-only coupled Newton-Schulz square root on the tcgen05 engine - on n x n PSD matrices up to C2's n = 8192, timed with the
upload of the packed matrix excluded, next to the oracle's LAPACK route (dsyevr + dsyr loop, f64lapack.rs:78-108) on the host
cores for the sizes where that finishes in seconds.  Prints one JSON object."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from totsu_b200 import capi  # noqa: E402


def main():
    capi.init(0)
    L = capi.lib()
    out = {"rows": []}
    rng = np.random.default_rng(0)
    for n, kind in ((512, "full"), (2048, "full"), (8192, "diag"), (8192, "full")):
        if kind == "diag":
            p = np.diag(rng.uniform(0.05, 1.0, n)).astype(np.float32)          # bench.py's C2 P
        else:
            g = rng.standard_normal((n, n + 8)).astype(np.float32)
            p = (g @ g.T) / np.float32(n) + np.float32(0.05) * np.eye(n, dtype=np.float32)
        packed = np.ascontiguousarray(p.T[np.tril_indices(n)].astype(np.float32))
        mb = capi.Buf(dtype=np.float32, length=packed.size)
        wb = capi.Buf(dtype=np.float32, length=1)                               # the GEMM-only route does not touch the caller's work area
        wv = capi.View(wb.h, 0, 0)
        times = []
        for rep in range(2):
            mb.upload(packed)
            capi.check(L.tb_device_sync())
            t0 = time.perf_counter()
            st = L.tb_sqrt_psd_f32(mb.view(), 1e-12, capi.View(wb.h, 0, 0) if False else _work_view(n, wb))
            capi.check(L.tb_device_sync())
            times.append(time.perf_counter() - t0)
            capi.check(st)
        route, iters = C.c_int(), C.c_int()
        capi.check(L.tb_sqrt_psd_info(C.byref(route), C.byref(iters)))
        got_packed = mb.download()
        # check S^2 = P on a sample of columns (full n^3 check is the CPU's job only for small n)
        s = np.zeros((n, n), dtype=np.float64)
        s.T[np.tril_indices(n)] = got_packed
        s = np.triu(s) + np.triu(s, 1).T
        cols = rng.choice(n, size=min(n, 64), replace=False)
        err = float(np.abs(s @ s[:, cols] - p.astype(np.float64)[:, cols]).max() / np.abs(p).max())
        row = {"n": n, "matrix": kind, "route": {1: "newton_schulz", 2: "eigendecomposition"}.get(route.value, "?"), "steps": iters.value, "gemms": 3 * iters.value,
               "seconds": min(times), "useful_tflops": (3 * iters.value * 2.0 * n ** 3) / min(times) / 1e12, "rel_err_S2_vs_P_sampled": err}
        if n <= 2048:
            import totsu_oracle as O
            v = packed.astype(np.float64)
            t0 = time.perf_counter()
            O.F64LAPACK.map_eig(v, None, 1e-12, np.zeros(O.F64LAPACK.map_eig_worklen(n)), lambda e: np.sqrt(e) if e > 0 else None)
            row["cpu_lapack_f64_seconds"] = time.perf_counter() - t0
            row["rel_err_vs_lapack"] = float(np.abs(got_packed - v).max() / np.abs(v).max())
        out["rows"].append(row)
        mb.release(); wb.release()
        for bf in _WORK:
            bf.release()
        _WORK.clear()
    print(json.dumps(out))


_WORK = []


def _work_view(n, _unused):
    bf = capi.Buf(dtype=np.float32, length=2 * n * n + n)
    _WORK.append(bf)
    return bf.view()


if __name__ == "__main__":
    main()
