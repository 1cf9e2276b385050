#!/usr/bin/env python
"""Phase timeline of one tcgen05 GEMM launch inside a back-to-back stream of them (see tb_symm_gemm_trace_f32)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from totsu_b200 import capi  # noqa: E402

NAMES = ["entered", "prologue done", "dependency resolved", "first chunk staged", "producers done", "accumulator complete",
         "partials pushed", "cluster barrier passed", "results stored", "own rows reduced (warp 0)", "all warps reduced"]


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    capi.init(0)
    L = capi.lib()
    rng = np.random.default_rng(0)
    g = rng.standard_normal((k, k)).astype(np.float32)
    sym = ((g + g.T) / 2).reshape(-1).copy()
    a, b, c = capi.Buf(dtype=np.float32, length=k * k), capi.Buf(dtype=np.float32, length=k * k), capi.Buf(dtype=np.float32, length=k * k)
    a.upload(sym); b.upload(sym)
    for splitk in (8, 4, 1):
        for trial in range(3):
            st = (C.c_uint64 * 16)()
            capi.check(L.tb_symm_gemm_trace_f32(k, a.view(), b.view(), c.view(), splitk, 20, st))
            t = [int(v) for v in st[:11]]
            base = t[0]
            print("k=%d splitk=%d:" % (k, splitk), ", ".join("%s +%.2fus" % (NAMES[i], (t[i] - base) / 1e3) for i in (1, 2, 3, 4, 5, 6, 7, 9, 10, 8) if t[i]))


if __name__ == "__main__":
    main()
