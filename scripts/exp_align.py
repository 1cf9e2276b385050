"""Experiment: does the source alignment of cp.async.bulk (16 B vs 128 B) change the streaming rate?  transform_ge N/T on a
matrix whose leading dimension is a multiple of 32 floats (every column 128-B aligned) vs a multiple of 4 only (16-B aligned)."""
import json, math, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from totsu_b200 import capi
capi.init(0); L = capi.lib(); dt = np.float32
stream = torch.cuda.ExternalStream(capi.stream_ptr())
ge = capi.fn("tb_transform_ge", dt)
out = []
for m in (65536, 65540, 65544, 65552, 65568):
    n = 16384
    a = capi.Buf(dtype=dt, length=m * n)
    capi.check(capi.fn("tb_fill_uniform", dt)(a.view(), m, n, 0, 0, dt(1.0 / math.sqrt(n))))
    xn, ym, xm, yn = capi.Buf(dtype=dt, length=n), capi.Buf(dtype=dt, length=m), capi.Buf(dtype=dt, length=m), capi.Buf(dtype=dt, length=n)
    xn.upload(np.ones(n, dtype=dt)); xm.upload(np.ones(m, dtype=dt))
    for tr, x, y in ((0, xn, ym), (1, xm, yn)):
        for _ in range(3):
            capi.check(ge(tr, m, n, 1.0, a.view(), x.view(), 0.0, y.view()))
        capi.check(L.tb_device_sync())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10):
            capi.check(ge(tr, m, n, 1.0, a.view(), x.view(), 0.0, y.view()))
        capi.check(L.tb_flush())
        e1.record(stream)
        capi.check(L.tb_device_sync()); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        out.append({"n_row": m, "row_bytes_mod_128": (m * 4) % 128, "trans": tr, "ms": ms, "gbs": m * n * 4 / ms / 1e6})
        print(out[-1], flush=True)
    for b in (a, xn, ym, xm, yn):
        b.release()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2e_exp_align.json"), "w"), indent=1)
