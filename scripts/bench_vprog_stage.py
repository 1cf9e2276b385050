#!/usr/bin/env python
"""What does one barrier-separated stage of a cluster vector program cost?  A chain of 40 dependent copies x -> y -> x ...
(different bases: the recorder puts a cluster barrier before each) against a chain of 40 scales of x (same thread owns the
element: no barriers), for several vector lengths and both settings of the cluster / wide threshold."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from totsu_b200 import capi
capi.init(0); L = capi.lib(); dt = np.float32
stream = torch.cuda.ExternalStream(capi.stream_ptr())
copy, scale = capi.fn("tb_copy", dt), capi.fn("tb_scale", dt)
rows = []
for max_n in (49152, 262144):
    capi.check(L.tb_set_vprog_max_n(max_n))
    for n in (1024, 8193, 17412, 65536, 147457):
        x, y = capi.Buf(dtype=dt, length=n), capi.Buf(dtype=dt, length=n)
        x.upload(np.ones(n, dtype=dt))
        def chain_copy():
            for _ in range(20):
                capi.check(copy(x.view(), y.view())); capi.check(copy(y.view(), x.view()))
        def chain_scale():
            for _ in range(40):
                capi.check(scale(1.0001, x.view()))
        for name, fn in (("40 dependent copies (barrier or launch each)", chain_copy), ("40 scales of one vector (no hazards)", chain_scale)):
            import ctypes as C
            fn(); capi.check(L.tb_flush()); capi.check(L.tb_device_sync())
            l0, o0, l1, o1 = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
            capi.check(L.tb_vprog_stats(C.byref(l0), C.byref(o0)))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record(stream)
            for _ in range(reps):
                fn(); capi.check(L.tb_flush())
            e1.record(stream)
            capi.check(L.tb_device_sync()); torch.cuda.synchronize()
            capi.check(L.tb_vprog_stats(C.byref(l1), C.byref(o1)))
            us = e0.elapsed_time(e1) * 1e3 / reps
            rows.append({"vprog_max_n": max_n, "n": n, "chain": name, "us_per_chain": us, "us_per_op": us / 40, "launches_per_chain": (l1.value - l0.value) / reps})
            print(rows[-1], flush=True)
        x.release(); y.release()
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "r2l_vprog_stage.json"), "w"), indent=1)
