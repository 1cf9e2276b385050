#!/usr/bin/env python
"""Launches, a few times each and at config-C3 sizes, the small kernels that had no `ncu --set full` capture in round 1
(VERDICT r1 item 7): reduce_kernel, finalize_kernel / finalize2_kernel, gemv_n_generic / gemv_t_generic, axpby / diag / copy /
fill kernels, cone_kernel, prefetch_reduce_kernel, vprog kernels.  Run under
    ncu --set full --clock-control none --import-source on -k regex:'<names>' -c <N> python scripts/ncu_targets.py
Vector programs are switched off for the level-1 part so that each command is its own kernel."""
import ctypes as C
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from totsu_b200 import capi, host  # noqa: E402


def main():
    capi.init(0)
    L = capi.lib()
    dt = np.float32
    rng = np.random.default_rng(0)
    n, m = 16384, 65536
    lx = n + 2 * m + 1                         # x_hat of C3: 147457 elements
    capi.check(L.tb_set_vprog(0))
    a, b, d = (capi.Buf(dtype=dt, length=lx) for _ in range(3))
    for bf in (a, b, d):
        bf.upload(rng.standard_normal(lx).astype(dt))
    out = C.c_float()
    for _ in range(3):
        capi.check(L.tb_copy_f32(a.view(), b.view()))                               # copy_kernel
        capi.check(L.tb_add_f32(-2.0, a.view(), b.view()))                          # axpby_kernel<1>
        capi.check(L.tb_transform_di_f32(1.0, d.view(), a.view(), 1.0, b.view()))   # diag_kernel<1>
        capi.check(L.tb_scale_f32(0.0, b.view(0, m)))                               # fill_kernel
        capi.check(L.tb_norm_f32(a.view(0, m), C.byref(out)))                       # reduce_kernel<0>, 65536 elements (|p|)
        capi.check(L.tb_norm_f32(a.view(), C.byref(out)))                           # reduce_kernel<0>, 147457 elements
    # the 63 x 16384 blocks of ProbSOCP (socp.rs:83-124): generic (LDG) matvecs, N and T
    g = capi.Buf(dtype=dt, length=63 * n)
    g.upload((rng.standard_normal(63 * n) / math.sqrt(n)).astype(dt))
    xn, y63 = capi.Buf(dtype=dt, length=n), capi.Buf(dtype=dt, length=63)
    xn.upload(rng.standard_normal(n).astype(dt)); y63.upload(rng.standard_normal(63).astype(dt))
    for _ in range(3):
        capi.check(L.tb_transform_ge_f32(0, 63, n, 1.0, g.view(), xn.view(), 0.0, y63.view()))     # gemv_n_generic
        capi.check(L.tb_transform_ge_f32(1, 63, n, 1.0, g.view(), y63.view(), 1.0, xn.view()))     # gemv_t_generic
    # a streaming pair on a 16384 x 4096 matrix with separate finalize kernels (vector programs off): finalize2_kernel
    mm, nn = 16384, 4096
    abuf = capi.Buf(dtype=dt, length=mm * nn)
    capi.check(L.tb_fill_uniform_f32(abuf.view(), mm, nn, 0, 0, dt(1.0 / math.sqrt(nn))))
    h = C.c_int64()
    capi.check(L.tb_denseop_create(capi.TB_F32, abuf.view(), mm, nn, 0, mm, C.byref(h)))
    x4, ym, xm, y4 = capi.Buf(dtype=dt, length=nn), capi.Buf(dtype=dt, length=mm), capi.Buf(dtype=dt, length=mm), capi.Buf(dtype=dt, length=nn)
    x4.upload(rng.standard_normal(nn).astype(dt)); xm.upload(rng.standard_normal(mm).astype(dt))
    for _ in range(3):
        capi.check(L.tb_denseop_apply_pair_f32(h.value, 1.0, x4.view(), 0.0, ym.view(), 1.0, xm.view(), 0.0, y4.view()))   # stream_kernel<1,1> + finalize2_kernel
        capi.check(L.tb_transform_ge_f32(0, mm, nn, 1.0, abuf.view(), x4.view(), 0.0, ym.view()))                         # stream_kernel<1,0> + finalize_kernel
    tau, sig = capi.Buf(dtype=dt, length=nn), capi.Buf(dtype=dt, length=mm)
    capi.check(L.tb_denseop_absadd_cols_f32(h.value, tau.view()))                   # stream_kernel<1,1,ABS>: |A|^T 1 and |A| 1 from one read
    capi.check(L.tb_denseop_absadd_rows_f32(h.value, sig.view()))
    capi.check(L.tb_denseop_destroy(h.value))
    # batched cone projection, C3's cone
    blocks = [(capi.CONE_SOC, 64)] * 1024
    arr = (capi.ConeBlock * len(blocks))(*[capi.ConeBlock(t, 0, ln) for t, ln in blocks])
    hc = C.c_int64()
    capi.check(L.tb_cone_create(arr, len(blocks), C.byref(hc)))
    v = capi.Buf(dtype=dt, length=m)
    v.upload(rng.standard_normal(m).astype(dt))
    for _ in range(3):
        capi.check(L.tb_cone_proj_f32(hc.value, 0, v.view(), 1e-12, capi.View(0, 0, 0)))   # cone_kernel<float,0>
    capi.check(L.tb_cone_destroy(hc.value))
    # vector programs + scalar prefetch inside a small solve (vprog_kernel, vprog_wide_kernel, prefetch_reduce_kernel)
    capi.check(L.tb_set_vprog(1))
    import helpers as H
    blocks = [(H.SOC, 64)] * 512 + [(H.RPOS, 20000)]
    mt = sum(l for _, l in blocks)
    am, bm, cm = H.make_instance(mt, 600, blocks, seed=5, dtype=dt)
    ab2, av2 = H.device_matrix(am)
    s = host.Session.dense(dt, av2, mt, 600, cm, bm, blocks, fused_op=True, fused_cone=True)
    assert s.begin(max_iter=None, eps_acc=0.0, eps_inf=0.0, device_precond=True) == "None"
    s.step(12)
    s.close()
    capi.check(L.tb_device_sync())
    print("ncu targets done")


if __name__ == "__main__":
    main()
