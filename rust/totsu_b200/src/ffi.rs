//! Raw bindings of `include/totsu_b200.h` (the f32 half; the `_f64` entry points mirror these 1:1).
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

pub type tb_handle = i64;

#[repr(C)]
#[derive(Copy, Clone, Debug)]
pub struct tb_view {
    pub buf: tb_handle,
    pub off: usize,
    pub len: usize,
}

#[repr(C)]
#[derive(Copy, Clone, Debug)]
pub struct tb_cone_block {
    pub typ: i32,
    pub reserved: i32,
    pub len: u64,
}

pub const TB_OK: c_int = 0;
pub const TB_ERR_ARG: c_int = 2;
pub const TB_F32: c_int = 0;
pub const TB_CONE_ZERO: i32 = 0;
pub const TB_CONE_RPOS: i32 = 1;
pub const TB_CONE_SOC: i32 = 2;
pub const TB_CONE_ROTSOC: i32 = 3;
pub const TB_CONE_PSD: i32 = 4;

extern "C" {
    pub fn tb_init(device: c_int) -> c_int;
    pub fn tb_last_error() -> *const c_char;

    pub fn tb_buf_wrap(dtype: c_int, host: *mut c_void, len: usize, host_is_mut: c_int, out: *mut tb_handle) -> c_int;
    pub fn tb_buf_retain(buf: tb_handle, n: c_int) -> c_int;
    pub fn tb_buf_release(buf: tb_handle) -> c_int;
    pub fn tb_view_of_host(dtype: c_int, host: *const c_void, len: usize, out: *mut tb_view) -> c_int;
    pub fn tb_host_ref(v: tb_view) -> c_int;
    pub fn tb_host_mut(v: tb_view) -> c_int;
    pub fn tb_get1_f32(v: tb_view, idx: usize, out: *mut f32) -> c_int;
    pub fn tb_set1_f32(v: tb_view, idx: usize, val: f32) -> c_int;

    pub fn tb_norm_f32(x: tb_view, out: *mut f32) -> c_int;
    pub fn tb_copy_f32(x: tb_view, y: tb_view) -> c_int;
    pub fn tb_scale_f32(alpha: f32, x: tb_view) -> c_int;
    pub fn tb_add_f32(alpha: f32, x: tb_view, y: tb_view) -> c_int;
    pub fn tb_adds_f32(s: f32, y: tb_view) -> c_int;
    pub fn tb_abssum_f32(x: tb_view, incx: usize, out: *mut f32) -> c_int;
    pub fn tb_transform_di_f32(alpha: f32, mat: tb_view, x: tb_view, beta: f32, y: tb_view) -> c_int;

    pub fn tb_transform_ge_f32(transpose: c_int, n_row: usize, n_col: usize, alpha: f32, mat: tb_view, x: tb_view, beta: f32, y: tb_view) -> c_int;
    pub fn tb_transform_sp_f32(n: usize, alpha: f32, mat: tb_view, x: tb_view, beta: f32, y: tb_view) -> c_int;
    pub fn tb_map_eig_worklen(n: usize) -> usize;
    pub fn tb_map_eig_begin_f32(mat: tb_view, has_scale: c_int, scale_diag: f32, eps_zero: f32, work: tb_view, host_eigs: *mut f32) -> c_int;
    pub fn tb_map_eig_finish_f32(mat: tb_view, has_scale: c_int, scale_diag: f32, work: tb_view, new_eigs: *const f32, keep: *const u8) -> c_int;

    pub fn tb_denseop_create(dtype: c_int, mat: tb_view, n_row: usize, n_col: usize, row_offset: usize, n_row_total: usize, out: *mut tb_handle) -> c_int;
    pub fn tb_denseop_destroy(op: tb_handle) -> c_int;
    pub fn tb_denseop_apply_f32(op: tb_handle, transpose: c_int, alpha: f32, x: tb_view, beta: f32, y: tb_view) -> c_int;
    pub fn tb_denseop_absadd_cols_f32(op: tb_handle, tau: tb_view) -> c_int;
    pub fn tb_denseop_absadd_rows_f32(op: tb_handle, sigma: tb_view) -> c_int;

    pub fn tb_cone_create(blocks: *const tb_cone_block, n_blocks: usize, out: *mut tb_handle) -> c_int;
    pub fn tb_cone_destroy(cone: tb_handle) -> c_int;
    pub fn tb_cone_proj_f32(cone: tb_handle, dual_cone: c_int, x: tb_view, eps_zero: f32, psd_work: tb_view) -> c_int;
    pub fn tb_cone_group_min_f32(cone: tb_handle, dp_tau: tb_view) -> c_int;

    // switches / housekeeping: none is needed for correctness - every call that returns data to the host drains deferred work
    pub fn tb_flush() -> c_int;
    pub fn tb_device_sync() -> c_int;
    pub fn tb_set_pair_fusion(on: c_int) -> c_int;
    pub fn tb_set_speculation(on: c_int) -> c_int;
    pub fn tb_set_vprog(on: c_int) -> c_int;
}

/// The traits have no error channel, so a failed call panics - exactly like totsu_f32cuda asserts on every cuBLAS
/// status (totsu_f32cuda/src/f32cuda.rs:38).
#[inline]
pub fn check(st: c_int, what: &str) {
    if st != TB_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(tb_last_error()) }.to_string_lossy().into_owned();
        panic!("totsu_b200: {} failed with status {}: {}", what, st, msg);
    }
}

/// One process drives one GPU; `LinAlg` functions have no `self` (linalg.rs:22-67), so the context is process-global.
pub fn ensure_init() {
    use std::sync::Once;
    static INIT: Once = Once::new();
    INIT.call_once(|| {
        check(unsafe { tb_init(-1) }, "tb_init"); // -1: $LOCAL_RANK or device 0
        log::info!("totsu_b200: backend initialised");
    });
}
