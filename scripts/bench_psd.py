#!/usr/bin/env python
"""Config C4 micro-benchmark: ConePSD::proj of one 512 x 512 block (sk = 131 328) - projections/second of the two
device paths (matrix-sign iteration, Jacobi eigendecomposition) next to the oracle's LAPACK dsyevr + dsyr restatement
(f64lapack.rs:78-108) on the host cores.  Prints one JSON line."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import helpers as H  # noqa: E402
from totsu_b200 import capi  # noqa: E402


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    capi.init(0)
    L = capi.lib()
    stream = torch.cuda.ExternalStream(capi.stream_ptr())
    rng = np.random.default_rng(k)
    g = rng.standard_normal((k, k))
    x0 = H.svec((g + g.T) / 2)
    out = {"k": k, "reps": reps}
    jac = os.environ.get("PSD_BENCH_JACOBI", "1") == "1"
    for dt in (np.float32, np.float64):
        paths = [(0, "sign_tcgen05"), (3, "sign_tcgen05_nosplit"), (2, "sign_fp32pipe")] if dt == np.float32 else [(0, "sign")]
        if jac:
            paths.append((1, "jacobi"))
        for path, name in paths:
            n_rep = reps if path != 1 else max(2, reps // 10)
            capi.check(L.tb_set_psd_path(path))
            xb = capi.Buf(dtype=dt, length=x0.size)
            wb = capi.Buf(dtype=dt, length=2 * k * k + k)
            xb.upload(x0.astype(dt))
            capi.check(capi.fn("tb_proj_psd", dt)(xb.view(), 1e-12, wb.view()))    # warm-up
            capi.check(L.tb_device_sync())
            l0 = capi.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(n_rep):
                xb.upload(x0.astype(dt))
                capi.check(capi.fn("tb_proj_psd", dt)(xb.view(), 1e-12, wb.view()))
            e1.record(stream)
            capi.check(L.tb_device_sync())
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n_rep
            out["%s_%s" % (name, np.dtype(dt).name)] = {"ms_per_proj": ms, "proj_per_s": 1e3 / ms, "launches_per_proj": (capi.launch_count() - l0) / n_rep}
            xb.release(); wb.release()
    capi.check(L.tb_set_psd_path(0))
    # the GEMM alone: 3xTF32 tcgen05 kernel (split-K chosen / off) vs the FP32-pipe kernel; useful flops = 2 k^3
    if k % 4 == 0:
        mats = [capi.Buf(dtype=np.float32, length=k * k) for _ in range(3)]
        sym = ((g + g.T) / (2 * np.linalg.norm(g))).astype(np.float32).reshape(-1, order="F").copy()
        for mb in mats[:2]:
            mb.upload(sym)
        empty = capi.View(0, 0, 0)
        for label, engine, splitk in (("gemm_tcgen05", 2, 0), ("gemm_tcgen05_nosplit", 2, 1), ("gemm_fp32pipe", 1, 0)):
            n_g = 200
            for _ in range(5):
                capi.check(L.tb_symm_gemm_f32(k, 1.0, mats[0].view(), mats[1].view(), 0.0, empty, 0.0, mats[2].view(), engine, splitk))
            capi.check(L.tb_device_sync())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(n_g):
                capi.check(L.tb_symm_gemm_f32(k, 1.0, mats[0].view(), mats[1].view(), 0.0, empty, 0.0, mats[2].view(), engine, splitk))
            e1.record(stream)
            capi.check(L.tb_device_sync())
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / n_g * 1e3
            out[label] = {"us_per_gemm": us, "useful_tflops": 2.0 * k ** 3 / (us * 1e-6) / 1e12,
                          "tensor_tflops_issued": (3 * 2.0 * k ** 3 / (us * 1e-6) / 1e12) if engine == 2 else None}
        for mb in mats:
            mb.release()
    # CPU: the oracle's restatement of F64LAPACK::map_eig (dsyevr V/V/U (0, inf] + dsyr loop), all host threads
    import totsu_oracle as O
    cone = O.ConePSD(np.zeros(O.ConePSD.query_worklen(x0.size)), 1e-12)
    t0 = time.perf_counter()
    n_cpu = 3
    for _ in range(n_cpu):
        v = x0.copy()
        cone.proj(False, v)
    out["cpu_lapack_f64"] = {"ms_per_proj": (time.perf_counter() - t0) / n_cpu * 1e3, "cores": os.cpu_count()}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
