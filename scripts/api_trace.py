#!/usr/bin/env python
"""Host-side view of one Solver::solve at config C3 through the C ABI: calls and wall seconds per entry point
(tb_set_api_trace), for begin() and for the iterations separately, with carried views and with the Rust binding's call
protocol (--shim-protocol).  Answers "which entry points does the host spend its time in"."""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C  # noqa: E402
from totsu_b200 import capi, host  # noqa: E402


def main():
    capi.init(0)
    L = capi.lib()
    dt = np.float32
    nblk, bdim, n = 1024, 64, 16384
    m = nblk * bdim
    abuf = capi.Buf(dtype=dt, length=m * n)
    capi.check(L.tb_fill_uniform_f32(abuf.view(), m, n, 0, 0, dt(1.0 / math.sqrt(n))))
    rng = np.random.default_rng(0)
    b = (rng.standard_normal(m) / math.sqrt(m)).astype(dt)
    c = (rng.standard_normal(n) / math.sqrt(m)).astype(dt)
    blocks = [(capi.CONE_SOC, bdim)] * nblk
    out = {}
    for proto in (0, 1):
        for dp in (0, 1):
            host.set_shim_protocol(bool(proto))
            s = host.Session.dense(dt, abuf.view(), m, n, c, b, blocks, fused_op=True, fused_cone=True)
            capi.check(L.tb_set_api_trace(1))
            assert s.begin(max_iter=None, eps_acc=0.0, eps_inf=0.0, device_precond=bool(dp)) == "None"
            t_begin = capi.api_trace_table()
            capi.check(L.tb_set_api_trace(1))
            s.step(50)
            capi.check(L.tb_flush())
            t_iter = capi.api_trace_table()
            capi.check(L.tb_set_api_trace(0))
            s.close()
            key = "shim_protocol=%d device_precond=%d" % (proto, dp)
            top = lambda t: sorted(((k, v[0], round(v[1] * 1e3, 3)) for k, v in t.items()), key=lambda r: -r[2])[:12]
            out[key] = {"begin_ms_total": round(sum(v[1] for v in t_begin.values()) * 1e3, 3), "begin_top": top(t_begin),
                        "iterations_50_ms_total": round(sum(v[1] for v in t_iter.values()) * 1e3, 3), "iterations_top": top(t_iter),
                        "calls_per_iteration": sum(v[0] for v in t_iter.values()) / 50}
    host.set_shim_protocol(False)
    abuf.release()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
