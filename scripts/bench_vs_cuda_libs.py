#!/usr/bin/env python
"""Same-box comparison with the library calls the reference's GPU backend makes ("beat cuBLAS / cuSOLVER as
called by F32CUDA, on the same box"):

  * transform_ge  : cublasSgemv_v2 N / T, column-major, lda = n_row            (totsu_f32cuda/src/f32cuda.rs:144-171)
                    vs this repo's stream_kernel (tb_transform_ge_f32) and the paired pass (tb_denseop_apply_pair_f32)
                    on C3's A (65536 x 16384) and C2's G (8192 x 8192);
  * ConePSD::proj : cusolverDnSsyevdx(jobz = V, range = V (0, +inf], uplo = U) + cublasSscal (zero A) + one cublasSsyr per
                    positive eigenvalue, after a D2H of the eigenvalues          (f32cuda.rs:223-303)
                    vs tb_proj_psd_f32 (GEMM-only sign iteration on tcgen05) on a random symmetric 512 x 512 (config C4).
The libraries are called through ctypes exactly as F32CUDA calls them (host-pointer scalars, default handles, bufferSize
re-queried per call like f32cuda.rs:247); the pack / unpack copies F32CUDA adds around the eigensolve (k launches each,
f32cuda.rs:316-324,361-369) are NOT charged to the library side.  Device memory for the library side is torch's.
Prints one JSON object; commit it under profiles/."""
import ctypes as C
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from totsu_b200 import capi  # noqa: E402


def load(names):
    for n in names:
        try:
            return C.CDLL(n)
        except OSError:
            continue
    raise OSError("cannot load any of %s" % (names,))


def main():
    capi.init(0)
    L = capi.lib()
    dev = torch.device("cuda", 0)
    lib_stream = torch.cuda.ExternalStream(capi.stream_ptr(), device=dev)
    cublas = load(["libcublas.so.12", "libcublas.so"])
    cusolver = load(["libcusolver.so.11", "libcusolver.so"])
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    cusolver.cusolverDnSsyevdx_bufferSize.argtypes = [vp, ci, ci, ci, ci, vp, ci, cf, cf, ci, ci, C.POINTER(ci), vp, C.POINTER(ci)]
    cusolver.cusolverDnSsyevdx.argtypes = [vp, ci, ci, ci, ci, vp, ci, cf, cf, ci, ci, C.POINTER(ci), vp, vp, ci, vp]
    cublas.cublasSgemv_v2.argtypes = [vp, ci, ci, ci, C.POINTER(cf), vp, ci, vp, ci, C.POINTER(cf), vp, ci]
    cublas.cublasSsyr_v2.argtypes = [vp, ci, ci, C.POINTER(cf), vp, ci, vp, ci]
    cublas.cublasSscal_v2.argtypes = [vp, ci, C.POINTER(cf), vp, ci]
    hb, hs = C.c_void_p(), C.c_void_p()
    assert cublas.cublasCreate_v2(C.byref(hb)) == 0
    assert cusolver.cusolverDnCreate(C.byref(hs)) == 0
    ts = torch.cuda.current_stream(dev)
    assert cublas.cublasSetStream_v2(hb, C.c_void_p(ts.cuda_stream)) == 0
    assert cusolver.cusolverDnSetStream(hs, C.c_void_p(ts.cuda_stream)) == 0
    out = {"peak_gbs": peak, "gemv": [], "psd": {}}

    def time_torch(fn, reps):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def time_lib(fn, reps):
        fn(); capi.check(L.tb_flush()); capi.check(L.tb_device_sync())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(lib_stream)
        for _ in range(reps):
            fn()
        capi.check(L.tb_flush())
        e1.record(lib_stream)
        capi.check(L.tb_device_sync()); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # ---- gemv: the matrices are larger than L2 (4.3 GB) or L2-sized (268 MB: flushed by alternating two copies)
    one, zero = C.c_float(1.0), C.c_float(0.0)
    for (m, n, tag) in ((65536, 16384, "C3 A"), (8192, 8192, "C2 G")):
        copies = 1 if m * n * 4 > 1e9 else 4            # rotate over 4 x 268 MB > 126 MB L2
        abufs = [capi.Buf(dtype=np.float32, length=m * n) for _ in range(copies)]
        for ab in abufs:
            capi.check(L.tb_fill_uniform_f32(ab.view(), m, n, 0, 0, np.float32(1.0 / math.sqrt(n))))
        # the same bytes for cuBLAS: copy our device buffer into a torch tensor (D2D through the host-free tb_download is not
        # available for 4 GB; regenerate with torch instead - values do not matter for timing)
        tas = [torch.empty(m * n, dtype=torch.float32, device=dev).uniform_(-1, 1) for _ in range(copies)]
        xn, xm = torch.randn(n, device=dev), torch.randn(m, device=dev)
        ym, yn = torch.zeros(m, device=dev), torch.zeros(n, device=dev)
        bxn, bxm, bym, byn = (capi.Buf(dtype=np.float32, length=k) for k in (n, m, m, n))
        bxn.upload(xn.cpu().numpy()); bxm.upload(xm.cpu().numpy())
        it = {"i": 0}

        def cublas_n():
            a = tas[it["i"] % copies]; it["i"] += 1
            assert cublas.cublasSgemv_v2(hb, 0, m, n, C.byref(one), C.c_void_p(a.data_ptr()), m, C.c_void_p(xn.data_ptr()), 1, C.byref(zero), C.c_void_p(ym.data_ptr()), 1) == 0

        def cublas_t():
            a = tas[it["i"] % copies]; it["i"] += 1
            assert cublas.cublasSgemv_v2(hb, 1, m, n, C.byref(one), C.c_void_p(a.data_ptr()), m, C.c_void_p(xm.data_ptr()), 1, C.byref(zero), C.c_void_p(yn.data_ptr()), 1) == 0

        def ours_n():
            a = abufs[it["i"] % copies]; it["i"] += 1
            capi.check(L.tb_transform_ge_f32(0, m, n, 1.0, a.view(), bxn.view(), 0.0, bym.view()))

        def ours_t():
            a = abufs[it["i"] % copies]; it["i"] += 1
            capi.check(L.tb_transform_ge_f32(1, m, n, 1.0, a.view(), bxm.view(), 0.0, byn.view()))
        hops = []
        for ab in abufs:
            h = C.c_int64()
            capi.check(L.tb_denseop_create(capi.TB_F32, ab.view(), m, n, 0, m, C.byref(h)))
            hops.append(h.value)

        def ours_pair():
            h = hops[it["i"] % copies]; it["i"] += 1
            capi.check(L.tb_denseop_apply_pair_f32(h, 1.0, bxn.view(), 0.0, bym.view(), 1.0, bxm.view(), 0.0, byn.view()))
        reps = 20
        nbytes = m * n * 4
        row = {"matrix": "%s %d x %d f32 (%.0f MB)" % (tag, m, n, nbytes / 1e6), "l2": "larger than L2" if copies == 1 else "rotating over %d copies (> L2)" % copies}
        for name, fn, timer, mult in (("cublasSgemv_v2 N", cublas_n, time_torch, 1), ("cublasSgemv_v2 T", cublas_t, time_torch, 1),
                                      ("tb_transform_ge N (stream_kernel<1,0>)", ours_n, time_lib, 1), ("tb_transform_ge T (stream_kernel<0,1>)", ours_t, time_lib, 1),
                                      ("tb_denseop_apply_pair (stream_kernel<1,1>: N and T from one read)", ours_pair, time_lib, 1)):
            ms = timer(fn, reps)
            row[name] = {"ms": ms, "gbs_streamed": nbytes * mult / (ms * 1e-3) / 1e9, "frac_of_measured_hbm_peak": nbytes * mult / (ms * 1e-3) / 1e9 / peak}
        row["op_plus_trans_op"] = {"cublas_ms": row["cublasSgemv_v2 N"]["ms"] + row["cublasSgemv_v2 T"]["ms"],
                                   "totsu_b200_pair_ms": row["tb_denseop_apply_pair (stream_kernel<1,1>: N and T from one read)"]["ms"]}
        row["op_plus_trans_op"]["speedup"] = row["op_plus_trans_op"]["cublas_ms"] / row["op_plus_trans_op"]["totsu_b200_pair_ms"]
        out["gemv"].append(row)
        for h in hops:
            capi.check(L.tb_denseop_destroy(h))
        for bf in abufs + [bxn, bxm, bym, byn]:
            bf.release()
        del tas
        torch.cuda.empty_cache()

    # ---- ConePSD::proj at k = 512: F32CUDA's eig_func (f32cuda.rs:242-303) vs tb_proj_psd_f32
    import helpers as H
    k = 512
    rng = np.random.default_rng(k)
    g = rng.standard_normal((k, k))
    sym = ((g + g.T) / 2).astype(np.float32)
    a0 = torch.tensor(sym, device=dev).t().contiguous()          # column-major k x k (symmetric anyway)
    a = torch.empty_like(a0)
    w = torch.empty(k, device=dev)
    z = torch.empty(k * k, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    meig, lwork = C.c_int(0), C.c_int(0)
    work = {"t": None}

    def f32cuda_eig_func():
        a.copy_(a0)
        # eig_func_worklen: bufferSize re-queried on every call (f32cuda.rs:247)
        assert cusolver.cusolverDnSsyevdx_bufferSize(hs, 1, 1003, 1, k, None, k, 0.0, float("inf"), 0, 0, None, None, C.byref(lwork)) == 0
        if work["t"] is None or work["t"].numel() < lwork.value:
            work["t"] = torch.empty(max(lwork.value, 1), device=dev)
        assert cusolver.cusolverDnSsyevdx(hs, 1, 1003, 1, k, C.c_void_p(a.data_ptr()), k, 0.0, float("inf"), 0, 0, C.byref(meig), C.c_void_p(w.data_ptr()),
                                          C.c_void_p(work["t"].data_ptr()), lwork.value, C.c_void_p(info.data_ptr())) == 0
        assert int(info.cpu()[0]) == 0                           # dev_info.copy_to (f32cuda.rs:268)
        z.copy_(a.reshape(-1))                                   # a -> z (f32cuda.rs:273)
        assert cublas.cublasSscal_v2(hb, k * k, C.byref(zero), C.c_void_p(a.data_ptr()), 1) == 0
        w_host = w.cpu().numpy()                                 # w.get_ref(): D2H of the eigenvalues (f32cuda.rs:287)
        for i in range(meig.value):
            e = C.c_float(float(w_host[i]))
            if e.value > 0.0:
                assert cublas.cublasSsyr_v2(hb, 1, k, C.byref(e), C.c_void_p(z.data_ptr() + 4 * i * k), 1, C.c_void_p(a.data_ptr()), k) == 0
    ms_lib = time_torch(f32cuda_eig_func, 5)
    proj_lib = np.triu(a.cpu().numpy().T)                        # upper triangle of the result
    x0 = H.svec((g + g.T) / 2).astype(np.float32)
    xb = capi.Buf(dtype=np.float32, length=x0.size)
    wb = capi.Buf(dtype=np.float32, length=2 * k * k + k)

    def ours():
        capi.check(L.tb_upload(xb.view(), x0.ctypes.data_as(C.c_void_p)))
        capi.check(L.tb_proj_psd_f32(xb.view(), 1e-12, wb.view()))
    ms_ours = time_lib(ours, 20)
    got = xb.download()
    # both against numpy f64
    ev, evec = np.linalg.eigh((g + g.T) / 2)
    want = (evec * np.maximum(ev, 0)) @ evec.T
    err_ours = float(np.abs(got - H.svec(want)).max() / np.abs(want).max())
    err_lib = float(np.abs(proj_lib - np.triu(want)).max() / np.abs(want).max())
    out["psd"] = {"k": k, "positive_eigenvalues": int(meig.value),
                  "f32cuda_route_ms": ms_lib, "f32cuda_route": "cusolverDnSsyevdx_bufferSize + cusolverDnSsyevdx(V, V, U, (0, inf]) + cublasSscal + D2H eigenvalues + %d x cublasSsyr" % meig.value,
                  "totsu_b200_ms": ms_ours, "totsu_b200_route": "tb_proj_psd_f32 (pack/unpack included; H2D of the 131328-float input included in both timings' setup)",
                  "speedup": ms_lib / ms_ours, "rel_err_vs_numpy_f64": {"f32cuda_route": err_lib, "totsu_b200": err_ours}}
    xb.release(); wb.release()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
