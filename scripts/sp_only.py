"""ncu target: transform_sp through the streaming kernel at n = 8192 / 16384, 16 and 8 consumer warps (scripts/gpu_r2_d.sh)."""
import sys, os, math, numpy as np
sys.path.insert(0, os.getcwd())
from totsu_b200 import capi
capi.init(0); L = capi.lib(); dt = np.float32
for n in (8192, 16384):
    sp = capi.Buf(dtype=dt, length=n*(n+1)//2)
    capi.check(capi.fn("tb_fill_uniform", dt)(sp.view(), n*(n+1)//2, 1, 0, 1, dt(0.01)))
    x, y = capi.Buf(dtype=dt, length=n), capi.Buf(dtype=dt, length=n)
    x.upload(np.ones(n, dtype=dt))
    for warps in (8, 16):
        capi.check(L.tb_set_spmv_warps(warps))
        for _ in range(1 if len(sys.argv) > 1 else 6):
            capi.check(capi.fn("tb_transform_sp", dt)(n, 1.0, sp.view(), x.view(), 0.0, y.view()))
    capi.check(L.tb_device_sync())
