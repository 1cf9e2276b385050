#!/usr/bin/env python
"""Per-kernel micro-benchmarks against the measured HBM peak (every row of SURVEY.md 8a that is a device kernel):
transform_ge (streaming N / T / paired), transform_sp, level-1 passes, the batched cone projection, ConePSD::proj.
Prints one JSON object; sizes are the BASELINE configs'.  L2 is flushed between timed repetitions for the
bandwidth-bound kernels whose operands would otherwise fit in it."""
import ctypes as C
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from totsu_b200 import capi  # noqa: E402


def main():
    capi.init(0)
    L = capi.lib()
    stream = torch.cuda.ExternalStream(capi.stream_ptr())
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    dt = np.float32
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")          # 256 MB > 126 MB L2

    def timed(fn, reps, flush_l2):
        fn(); capi.check(L.tb_flush()); capi.check(L.tb_device_sync())
        tot = 0.0
        for _ in range(reps):
            if flush_l2:
                flush.zero_()
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            capi.check(L.tb_flush())
            e1.record(stream)
            capi.check(L.tb_device_sync()); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / reps

    out = {"peak_gbs": peak, "rows": []}

    def row(name, ms, nbytes, note=""):
        gbs = nbytes / (ms * 1e-3) / 1e9
        out["rows"].append({"kernel": name, "ms": ms, "algorithmic_bytes": nbytes, "gbs": gbs, "frac_of_measured_hbm_peak": gbs / peak, "note": note})

    # ---- transform_ge / DenseOp on C3's A
    m, n = 65536, 16384
    abuf = capi.Buf(dtype=dt, length=m * n)
    capi.check(capi.fn("tb_fill_uniform", dt)(abuf.view(), m, n, 0, 0, dt(1.0 / math.sqrt(n))))
    xn, ym = capi.Buf(dtype=dt, length=n), capi.Buf(dtype=dt, length=m)
    xm, yn = capi.Buf(dtype=dt, length=m), capi.Buf(dtype=dt, length=n)
    rng = np.random.default_rng(0)
    xn.upload(rng.standard_normal(n).astype(dt)); xm.upload(rng.standard_normal(m).astype(dt))
    ge = capi.fn("tb_transform_ge", dt)
    row("transform_ge N (stream_kernel<1,0>), A 65536x16384", timed(lambda: capi.check(ge(0, m, n, 1.0, abuf.view(), xn.view(), 0.0, ym.view())), 10, False), m * n * 4)
    row("transform_ge T (stream_kernel<0,1>), A 65536x16384", timed(lambda: capi.check(ge(1, m, n, 1.0, abuf.view(), xm.view(), 0.0, yn.view())), 10, False), m * n * 4)
    hop = C.c_int64()
    capi.check(L.tb_denseop_create(capi.dtype_id(dt), abuf.view(), m, n, 0, m, C.byref(hop)))
    pair = capi.fn("tb_denseop_apply_pair", dt)
    row("denseop pair (stream_kernel<1,1>): op + trans_op, one read", timed(lambda: capi.check(pair(hop.value, 1.0, xn.view(), 0.0, ym.view(), 1.0, xm.view(), 0.0, yn.view())), 10, False),
        2 * m * n * 4, "algorithmic bytes = two matvecs; streamed bytes are half")
    capi.check(L.tb_denseop_destroy(hop.value))
    for bf in (abuf, xn, ym, xm, yn):
        bf.release()

    # ---- transform_sp on C2's packed P^(1/2)
    n = 8192
    sp = capi.Buf(dtype=dt, length=n * (n + 1) // 2)
    capi.check(capi.fn("tb_fill_uniform", dt)(sp.view(), n * (n + 1) // 2, 1, 0, 1, dt(1.0 / math.sqrt(n))))
    x, y = capi.Buf(dtype=dt, length=n), capi.Buf(dtype=dt, length=n)
    x.upload(rng.standard_normal(n).astype(dt))
    spf = capi.fn("tb_transform_sp", dt)
    for warps in (8, 16):
        capi.check(L.tb_set_spmv_warps(warps))
        row("transform_sp (spmv_stream_kernel, %d consumer warps + finalize), n = 8192 packed" % warps,
            timed(lambda: capi.check(spf(n, 1.0, sp.view(), x.view(), 0.0, y.view())), 10, True), n * (n + 1) // 2 * 4, "134 MB: L2 flushed between repetitions")
    capi.check(L.tb_set_spmv_warps(8))
    for bf in (sp, x, y):
        bf.release()

    # ---- level-1 on a C5-size x_hat (589 825 elements) and a 16M-element vector
    for ln, tag in ((589825, "C5 x_hat"), (1 << 24, "16M")):
        a, b, d = (capi.Buf(dtype=dt, length=ln) for _ in range(3))
        a.upload(rng.standard_normal(ln).astype(dt)); b.upload(rng.standard_normal(ln).astype(dt)); d.upload(rng.standard_normal(ln).astype(dt))
        fl = ln * 4 > 100e6
        row("add (axpby_kernel), %s" % tag, timed(lambda: capi.check(capi.fn("tb_add", dt)(0.5, a.view(), b.view())), 10, fl), 3 * ln * 4)
        row("transform_di (diag_kernel), %s" % tag, timed(lambda: capi.check(capi.fn("tb_transform_di", dt)(1.0, d.view(), a.view(), 1.0, b.view())), 10, fl), 4 * ln * 4)
        nrm = C.c_float()
        row("norm (reduce_kernel + host-visible scalar), %s" % tag, timed(lambda: capi.check(capi.fn("tb_norm", dt)(a.view(), C.byref(nrm))), 10, fl), ln * 4)
        for bf in (a, b, d):
            bf.release()

    # ---- batched cone projection: C3's 1024 x SOC(64) and C5's RPos(262144)
    for blocks, tag in (([(capi.CONE_SOC, 64)] * 1024, "1024 x ConeSOC(64)"), ([(capi.CONE_RPOS, 262144)], "ConeRPos(262144)")):
        ln = sum(b for _, b in blocks)
        arr = (capi.ConeBlock * len(blocks))(*[capi.ConeBlock(t, 0, b) for t, b in blocks])
        h = C.c_int64()
        capi.check(L.tb_cone_create(arr, len(blocks), C.byref(h)))
        v = capi.Buf(dtype=dt, length=ln)
        v.upload(rng.standard_normal(ln).astype(dt))
        row("cone proj (cone_kernel), %s" % tag, timed(lambda: capi.check(capi.fn("tb_cone_proj", dt)(h.value, 0, v.view(), 1e-12, capi.View(0, 0, 0))), 10, False), 2 * ln * 4,
            "latency-bound: %d KB" % (ln * 4 // 1024))
        capi.check(L.tb_cone_destroy(h.value)); v.release()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
