"""ctypes binding of libtotsu_b200.so (include/totsu_b200.h).

Thin on purpose: the product is the CUDA library and the C++ host layer (totsu_b200/host); Python only loads the
shared objects for the tests and bench.py.  Importing this module never touches a GPU; `init()` does, and fails
loudly when there is none (there is no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtotsu_b200.so")

TB_F32, TB_F64 = 0, 1
CONE_ZERO, CONE_RPOS, CONE_SOC, CONE_ROTSOC, CONE_PSD = 0, 1, 2, 3, 4
NCCL_ID_BYTES = 128


class View(C.Structure):
    _fields_ = [("buf", C.c_int64), ("off", C.c_size_t), ("len", C.c_size_t)]

    def split(self, mid):
        assert 0 <= mid <= self.len
        return View(self.buf, self.off, mid), View(self.buf, self.off + mid, self.len - mid)

    def sub(self, off, ln):
        assert off + ln <= self.len
        return View(self.buf, self.off + off, ln)


class ConeBlock(C.Structure):
    _fields_ = [("type", C.c_int32), ("reserved", C.c_int32), ("len", C.c_uint64)]


class TotsuB200Error(RuntimeError):
    pass


_lib = None

# every symbol include/totsu_b200.h declares: (name, restype, argtypes or None for per-dtype generation)
_F = {TB_F32: C.c_float, TB_F64: C.c_double}
_SUF = {TB_F32: "f32", TB_F64: "f64"}


def _signatures():
    sz, i, u64, vp = C.c_size_t, C.c_int, C.c_uint64, C.c_void_p
    H = C.c_int64
    sig = {
        "tb_init": (i, [i]), "tb_shutdown": (i, []), "tb_last_error": (C.c_char_p, []), "tb_device_sync": (i, []),
        "tb_get_stream": (i, [C.POINTER(vp)]), "tb_sm_count": (i, [C.POINTER(i)]),
        "tb_launch_count": (i, [C.POINTER(u64)]), "tb_set_gemv_path": (i, [i]), "tb_set_pdl": (i, [i]), "tb_set_spmv_warps": (i, [i]), "tb_set_psd_path": (i, [i]),
        "tb_set_pair_fusion": (i, [i]), "tb_pairs_fused": (i, [C.POINTER(u64)]),
        "tb_host_wait_stats": (i, [C.POINTER(C.c_double), C.POINTER(u64)]),
        "tb_set_psd_pairing": (i, [i]), "tb_psd_pairs": (i, [C.POINTER(u64)]), "tb_cone_pairs": (i, [C.POINTER(u64)]),
        "tb_set_speculation": (i, [i]), "tb_spec_stats": (i, [C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]),
        "tb_set_scalar_prefetch": (i, [i]), "tb_scalar_prefetch_stats": (i, [C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]),
        "tb_set_api_trace": (i, [i]), "tb_api_trace_dump": (i, [C.c_char_p, sz, C.POINTER(sz)]),
        "tb_timeline_begin": (i, [sz]), "tb_timeline_dump": (i, [C.c_char_p, sz, C.POINTER(sz)]),
        "tb_set_vprog": (i, [i]), "tb_set_vprog_max_n": (i, [sz]), "tb_vprog_stats": (i, [C.POINTER(u64), C.POINTER(u64)]), "tb_flush": (i, []),
        "tb_prof_enable": (i, [i]), "tb_prof_read": (i, [C.POINTER(u64), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "tb_prof_read_variants": (i, [C.POINTER(u64), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "tb_buf_wrap": (i, [i, vp, sz, i, C.POINTER(H)]), "tb_buf_alloc": (i, [i, sz, C.POINTER(H)]),
        "tb_buf_release": (i, [H]), "tb_buf_retain": (i, [H, i]), "tb_view_of_host": (i, [i, vp, sz, C.POINTER(View)]), "tb_buf_len": (i, [H, C.POINTER(sz)]),
        "tb_host_ref": (i, [View]), "tb_host_mut": (i, [View]),
        "tb_upload": (i, [View, vp]), "tb_download": (i, [View, vp]),
        "tb_map_eig_worklen": (sz, [sz]),
        "tb_sqrt_psd_info": (i, [C.POINTER(i), C.POINTER(i)]),
        "tb_symm_gemm_f32": (i, [sz, C.c_float, View, View, C.c_float, View, C.c_float, View, i, i]),
        "tb_symm_gemm_trace_f32": (i, [sz, View, View, View, i, i, C.POINTER(u64)]),
        "tb_denseop_create": (i, [i, View, sz, sz, sz, sz, C.POINTER(H)]), "tb_denseop_destroy": (i, [H]),
        "tb_cone_create": (i, [C.POINTER(ConeBlock), sz, C.POINTER(H)]), "tb_cone_destroy": (i, [H]),
        "tb_dist_unique_id": (i, [vp]), "tb_dist_init": (i, [i, i, vp]), "tb_dist_finalize": (i, []),
        "tb_dist_info": (i, [C.POINTER(i), C.POINTER(i)]), "tb_dist_p2p_enabled": (i, [C.POINTER(i)]), "tb_dist_exchanges": (i, [C.POINTER(u64)]),
    }
    for dt, F in _F.items():
        s = _SUF[dt]
        FP = C.POINTER(F)
        sig.update({
            f"tb_get1_{s}": (i, [View, sz, FP]), f"tb_set1_{s}": (i, [View, sz, F]),
            f"tb_norm_{s}": (i, [View, FP]), f"tb_copy_{s}": (i, [View, View]), f"tb_scale_{s}": (i, [F, View]),
            f"tb_add_{s}": (i, [F, View, View]), f"tb_adds_{s}": (i, [F, View]),
            f"tb_abssum_{s}": (i, [View, sz, FP]), f"tb_transform_di_{s}": (i, [F, View, View, F, View]),
            f"tb_transform_ge_{s}": (i, [i, sz, sz, F, View, View, F, View]),
            f"tb_transform_sp_{s}": (i, [sz, F, View, View, F, View]),
            f"tb_map_eig_begin_{s}": (i, [View, i, F, F, View, FP]),
            f"tb_map_eig_finish_{s}": (i, [View, i, F, View, FP, C.POINTER(C.c_uint8)]),
            f"tb_proj_psd_{s}": (i, [View, F, View]), f"tb_sqrt_psd_{s}": (i, [View, F, View]),
            f"tb_denseop_apply_{s}": (i, [H, i, F, View, F, View]),
            f"tb_denseop_apply_pair_{s}": (i, [H, F, View, F, View, F, View, F, View]),
            f"tb_denseop_absadd_cols_{s}": (i, [H, View]), f"tb_denseop_absadd_rows_{s}": (i, [H, View]),
            f"tb_cone_proj_{s}": (i, [H, i, View, F, View]), f"tb_cone_group_min_{s}": (i, [H, View]),
            f"tb_recip_clamp_{s}": (i, [F, View]),
            f"tb_fill_uniform_{s}": (i, [View, sz, sz, sz, u64, F]),
        })
    return sig


SIGNATURES = _signatures()


def lib():
    """Load libtotsu_b200.so (no GPU needed for loading)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TotsuB200Error(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                 "(totsu_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise TotsuB200Error("totsu_b200 error %d: %s" % (status, lib().tb_last_error().decode()))


def init(device=-1):
    check(lib().tb_init(device))


def dtype_id(dt):
    dt = np.dtype(dt)
    if dt == np.float32:
        return TB_F32
    if dt == np.float64:
        return TB_F64
    raise TypeError(dt)


def fn(name, dt):
    return getattr(lib(), "%s_%s" % (name, _SUF[dtype_id(dt)]))


class Buf:
    """A root buffer: SliceLike::new_ref / new_mut over a numpy array (kept alive here), or a device-only buffer."""

    def __init__(self, arr=None, mutable=True, dtype=None, length=None):
        L = lib()
        h = C.c_int64(0)
        if arr is not None:
            assert arr.flags["C_CONTIGUOUS"] and arr.ndim == 1
            self.arr = arr
            self.dtype = arr.dtype
            self.len = arr.size
            check(L.tb_buf_wrap(dtype_id(arr.dtype), arr.ctypes.data_as(C.c_void_p), arr.size, 1 if mutable else 0, C.byref(h)))
        else:
            self.arr = None
            self.dtype = np.dtype(dtype)
            self.len = int(length)
            check(L.tb_buf_alloc(dtype_id(dtype), self.len, C.byref(h)))
        self.h = h.value

    def view(self, off=0, ln=None):
        return View(self.h, off, self.len - off if ln is None else ln)

    def release(self):
        if self.h:
            check(lib().tb_buf_release(self.h))
            self.h = 0

    def upload(self, src, off=0):
        src = np.ascontiguousarray(src, dtype=self.dtype)
        check(lib().tb_upload(self.view(off, src.size), src.ctypes.data_as(C.c_void_p)))

    def download(self, off=0, ln=None):
        ln = self.len - off if ln is None else ln
        out = np.empty(ln, dtype=self.dtype)
        check(lib().tb_download(self.view(off, ln), out.ctypes.data_as(C.c_void_p)))
        return out


def api_trace_table():
    """{entry point: (calls, seconds)} collected since tb_set_api_trace(1)."""
    need = C.c_size_t()
    check(lib().tb_api_trace_dump(None, 0, C.byref(need)))
    buf = C.create_string_buffer(need.value + 4096)       # the dump call itself is traced: leave room for its own line
    check(lib().tb_api_trace_dump(buf, len(buf), C.byref(need)))
    out = {}
    for ln in buf.value.decode().splitlines():
        f = ln.split()
        if len(f) == 3:
            out[f[0]] = (int(f[1]), float(f[2]))
    return out


def stream_ptr():
    p = C.c_void_p()
    check(lib().tb_get_stream(C.byref(p)))
    return p.value


def launch_count():
    v = C.c_uint64()
    check(lib().tb_launch_count(C.byref(v)))
    return v.value


def pairs_fused():
    v = C.c_uint64()
    check(lib().tb_pairs_fused(C.byref(v)))
    return v.value


def sm_count():
    v = C.c_int()
    check(lib().tb_sm_count(C.byref(v)))
    return v.value
