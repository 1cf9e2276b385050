//! Raw bindings of `include/totsu_b200.h` (both element types) and [`Elem`], the per-precision dispatch the generic
//! `B200T<F>` / `B200SliceT<F>` / `DenseOpT<F>` / `ProductConeT<F>` types are written against.
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
pub const TB_F64: c_int = 1;
pub const TB_CONE_ZERO: i32 = 0;
pub const TB_CONE_RPOS: i32 = 1;
pub const TB_CONE_SOC: i32 = 2;
pub const TB_CONE_ROTSOC: i32 = 3;
pub const TB_CONE_PSD: i32 = 4;

macro_rules! typed_externs {
    ($F:ty, $get1:ident, $set1:ident, $norm:ident, $copy:ident, $scale:ident, $add:ident, $adds:ident, $abssum:ident, $di:ident,
     $ge:ident, $sp:ident, $eb:ident, $ef:ident, $apply:ident, $acols:ident, $arows:ident, $proj:ident, $gmin:ident, $ppsd:ident, $spsd:ident) => {
        extern "C" {
            pub fn $ppsd(x: tb_view, eps_zero: $F, work: tb_view) -> c_int;
            pub fn $spsd(mat: tb_view, eps_zero: $F, work: tb_view) -> c_int;
            pub fn $get1(v: tb_view, idx: usize, out: *mut $F) -> c_int;
            pub fn $set1(v: tb_view, idx: usize, val: $F) -> c_int;
            pub fn $norm(x: tb_view, out: *mut $F) -> c_int;
            pub fn $copy(x: tb_view, y: tb_view) -> c_int;
            pub fn $scale(alpha: $F, x: tb_view) -> c_int;
            pub fn $add(alpha: $F, x: tb_view, y: tb_view) -> c_int;
            pub fn $adds(s: $F, y: tb_view) -> c_int;
            pub fn $abssum(x: tb_view, incx: usize, out: *mut $F) -> c_int;
            pub fn $di(alpha: $F, mat: tb_view, x: tb_view, beta: $F, y: tb_view) -> c_int;
            pub fn $ge(transpose: c_int, n_row: usize, n_col: usize, alpha: $F, mat: tb_view, x: tb_view, beta: $F, y: tb_view) -> c_int;
            pub fn $sp(n: usize, alpha: $F, mat: tb_view, x: tb_view, beta: $F, y: tb_view) -> c_int;
            pub fn $eb(mat: tb_view, has_scale: c_int, scale_diag: $F, eps_zero: $F, work: tb_view, host_eigs: *mut $F) -> c_int;
            pub fn $ef(mat: tb_view, has_scale: c_int, scale_diag: $F, work: tb_view, new_eigs: *const $F, keep: *const u8) -> c_int;
            pub fn $apply(op: tb_handle, transpose: c_int, alpha: $F, x: tb_view, beta: $F, y: tb_view) -> c_int;
            pub fn $acols(op: tb_handle, tau: tb_view) -> c_int;
            pub fn $arows(op: tb_handle, sigma: tb_view) -> c_int;
            pub fn $proj(cone: tb_handle, dual_cone: c_int, x: tb_view, eps_zero: $F, psd_work: tb_view) -> c_int;
            pub fn $gmin(cone: tb_handle, dp_tau: tb_view) -> c_int;
        }
    };
}
typed_externs!(f32, tb_get1_f32, tb_set1_f32, tb_norm_f32, tb_copy_f32, tb_scale_f32, tb_add_f32, tb_adds_f32, tb_abssum_f32, tb_transform_di_f32,
               tb_transform_ge_f32, tb_transform_sp_f32, tb_map_eig_begin_f32, tb_map_eig_finish_f32, tb_denseop_apply_f32,
               tb_denseop_absadd_cols_f32, tb_denseop_absadd_rows_f32, tb_cone_proj_f32, tb_cone_group_min_f32, tb_proj_psd_f32, tb_sqrt_psd_f32);
typed_externs!(f64, tb_get1_f64, tb_set1_f64, tb_norm_f64, tb_copy_f64, tb_scale_f64, tb_add_f64, tb_adds_f64, tb_abssum_f64, tb_transform_di_f64,
               tb_transform_ge_f64, tb_transform_sp_f64, tb_map_eig_begin_f64, tb_map_eig_finish_f64, tb_denseop_apply_f64,
               tb_denseop_absadd_cols_f64, tb_denseop_absadd_rows_f64, tb_cone_proj_f64, tb_cone_group_min_f64, tb_proj_psd_f64, tb_sqrt_psd_f64);

extern "C" {
    pub fn tb_init(device: c_int) -> c_int;
    pub fn tb_last_error() -> *const c_char;

    pub fn tb_buf_wrap(dtype: c_int, host: *mut c_void, len: usize, host_is_mut: c_int, out: *mut tb_handle) -> c_int;
    pub fn tb_buf_alloc(dtype: c_int, len: usize, out: *mut tb_handle) -> c_int;
    pub fn tb_buf_retain(buf: tb_handle, n: c_int) -> c_int;
    pub fn tb_buf_release(buf: tb_handle) -> c_int;
    pub fn tb_view_of_host(dtype: c_int, host: *const c_void, len: usize, out: *mut tb_view) -> c_int;
    pub fn tb_host_ref(v: tb_view) -> c_int;
    pub fn tb_host_mut(v: tb_view) -> c_int;
    pub fn tb_map_eig_worklen(n: usize) -> usize;

    pub fn tb_denseop_create(dtype: c_int, mat: tb_view, n_row: usize, n_col: usize, row_offset: usize, n_row_total: usize, out: *mut tb_handle) -> c_int;
    pub fn tb_denseop_destroy(op: tb_handle) -> c_int;
    pub fn tb_cone_create(blocks: *const tb_cone_block, n_blocks: usize, out: *mut tb_handle) -> c_int;
    pub fn tb_cone_destroy(cone: tb_handle) -> c_int;

    // switches / housekeeping: none is needed for correctness - every call that returns data to the host drains deferred work
    pub fn tb_flush() -> c_int;
    pub fn tb_device_sync() -> c_int;
    pub fn tb_set_pair_fusion(on: c_int) -> c_int;
    pub fn tb_set_speculation(on: c_int) -> c_int;
    pub fn tb_set_vprog(on: c_int) -> c_int;
}

/// Element type of the backend: `f32` (`type F = f32`, the sibling of `F32CUDA`) or `f64` (bit-closer to `F64LAPACK`).
/// One method per suffixed C entry point.
pub trait Elem: num_traits::Float + Default + 'static {
    const DTYPE: c_int;
    unsafe fn get1(v: tb_view, idx: usize, out: *mut Self) -> c_int;
    unsafe fn set1(v: tb_view, idx: usize, val: Self) -> c_int;
    unsafe fn norm(x: tb_view, out: *mut Self) -> c_int;
    unsafe fn copy(x: tb_view, y: tb_view) -> c_int;
    unsafe fn scale(alpha: Self, x: tb_view) -> c_int;
    unsafe fn add(alpha: Self, x: tb_view, y: tb_view) -> c_int;
    unsafe fn adds(s: Self, y: tb_view) -> c_int;
    unsafe fn abssum(x: tb_view, incx: usize, out: *mut Self) -> c_int;
    unsafe fn transform_di(alpha: Self, mat: tb_view, x: tb_view, beta: Self, y: tb_view) -> c_int;
    unsafe fn transform_ge(transpose: c_int, n_row: usize, n_col: usize, alpha: Self, mat: tb_view, x: tb_view, beta: Self, y: tb_view) -> c_int;
    unsafe fn transform_sp(n: usize, alpha: Self, mat: tb_view, x: tb_view, beta: Self, y: tb_view) -> c_int;
    unsafe fn map_eig_begin(mat: tb_view, has_scale: c_int, scale_diag: Self, eps_zero: Self, work: tb_view, host_eigs: *mut Self) -> c_int;
    unsafe fn map_eig_finish(mat: tb_view, has_scale: c_int, scale_diag: Self, work: tb_view, new_eigs: *const Self, keep: *const u8) -> c_int;
    unsafe fn denseop_apply(op: tb_handle, transpose: c_int, alpha: Self, x: tb_view, beta: Self, y: tb_view) -> c_int;
    unsafe fn denseop_absadd_cols(op: tb_handle, tau: tb_view) -> c_int;
    unsafe fn denseop_absadd_rows(op: tb_handle, sigma: tb_view) -> c_int;
    unsafe fn cone_proj(cone: tb_handle, dual_cone: c_int, x: tb_view, eps_zero: Self, psd_work: tb_view) -> c_int;
    unsafe fn cone_group_min(cone: tb_handle, dp_tau: tb_view) -> c_int;
    unsafe fn proj_psd(x: tb_view, eps_zero: Self, work: tb_view) -> c_int;
    unsafe fn sqrt_psd(mat: tb_view, eps_zero: Self, work: tb_view) -> c_int;
}

macro_rules! impl_elem {
    ($F:ty, $DT:expr, $get1:ident, $set1:ident, $norm:ident, $copy:ident, $scale:ident, $add:ident, $adds:ident, $abssum:ident, $di:ident,
     $ge:ident, $sp:ident, $eb:ident, $ef:ident, $apply:ident, $acols:ident, $arows:ident, $proj:ident, $gmin:ident, $ppsd:ident, $spsd:ident) => {
        impl Elem for $F {
            unsafe fn proj_psd(x: tb_view, eps_zero: Self, work: tb_view) -> c_int { $ppsd(x, eps_zero, work) }
            unsafe fn sqrt_psd(mat: tb_view, eps_zero: Self, work: tb_view) -> c_int { $spsd(mat, eps_zero, work) }
            const DTYPE: c_int = $DT;
            unsafe fn get1(v: tb_view, idx: usize, out: *mut Self) -> c_int { $get1(v, idx, out) }
            unsafe fn set1(v: tb_view, idx: usize, val: Self) -> c_int { $set1(v, idx, val) }
            unsafe fn norm(x: tb_view, out: *mut Self) -> c_int { $norm(x, out) }
            unsafe fn copy(x: tb_view, y: tb_view) -> c_int { $copy(x, y) }
            unsafe fn scale(alpha: Self, x: tb_view) -> c_int { $scale(alpha, x) }
            unsafe fn add(alpha: Self, x: tb_view, y: tb_view) -> c_int { $add(alpha, x, y) }
            unsafe fn adds(s: Self, y: tb_view) -> c_int { $adds(s, y) }
            unsafe fn abssum(x: tb_view, incx: usize, out: *mut Self) -> c_int { $abssum(x, incx, out) }
            unsafe fn transform_di(alpha: Self, mat: tb_view, x: tb_view, beta: Self, y: tb_view) -> c_int { $di(alpha, mat, x, beta, y) }
            unsafe fn transform_ge(transpose: c_int, n_row: usize, n_col: usize, alpha: Self, mat: tb_view, x: tb_view, beta: Self, y: tb_view) -> c_int {
                $ge(transpose, n_row, n_col, alpha, mat, x, beta, y)
            }
            unsafe fn transform_sp(n: usize, alpha: Self, mat: tb_view, x: tb_view, beta: Self, y: tb_view) -> c_int { $sp(n, alpha, mat, x, beta, y) }
            unsafe fn map_eig_begin(mat: tb_view, has_scale: c_int, scale_diag: Self, eps_zero: Self, work: tb_view, host_eigs: *mut Self) -> c_int {
                $eb(mat, has_scale, scale_diag, eps_zero, work, host_eigs)
            }
            unsafe fn map_eig_finish(mat: tb_view, has_scale: c_int, scale_diag: Self, work: tb_view, new_eigs: *const Self, keep: *const u8) -> c_int {
                $ef(mat, has_scale, scale_diag, work, new_eigs, keep)
            }
            unsafe fn denseop_apply(op: tb_handle, transpose: c_int, alpha: Self, x: tb_view, beta: Self, y: tb_view) -> c_int { $apply(op, transpose, alpha, x, beta, y) }
            unsafe fn denseop_absadd_cols(op: tb_handle, tau: tb_view) -> c_int { $acols(op, tau) }
            unsafe fn denseop_absadd_rows(op: tb_handle, sigma: tb_view) -> c_int { $arows(op, sigma) }
            unsafe fn cone_proj(cone: tb_handle, dual_cone: c_int, x: tb_view, eps_zero: Self, psd_work: tb_view) -> c_int { $proj(cone, dual_cone, x, eps_zero, psd_work) }
            unsafe fn cone_group_min(cone: tb_handle, dp_tau: tb_view) -> c_int { $gmin(cone, dp_tau) }
        }
    };
}
impl_elem!(f32, TB_F32, tb_get1_f32, tb_set1_f32, tb_norm_f32, tb_copy_f32, tb_scale_f32, tb_add_f32, tb_adds_f32, tb_abssum_f32, tb_transform_di_f32,
           tb_transform_ge_f32, tb_transform_sp_f32, tb_map_eig_begin_f32, tb_map_eig_finish_f32, tb_denseop_apply_f32,
           tb_denseop_absadd_cols_f32, tb_denseop_absadd_rows_f32, tb_cone_proj_f32, tb_cone_group_min_f32, tb_proj_psd_f32, tb_sqrt_psd_f32);
impl_elem!(f64, TB_F64, tb_get1_f64, tb_set1_f64, tb_norm_f64, tb_copy_f64, tb_scale_f64, tb_add_f64, tb_adds_f64, tb_abssum_f64, tb_transform_di_f64,
           tb_transform_ge_f64, tb_transform_sp_f64, tb_map_eig_begin_f64, tb_map_eig_finish_f64, tb_denseop_apply_f64,
           tb_denseop_absadd_cols_f64, tb_denseop_absadd_rows_f64, tb_cone_proj_f64, tb_cone_group_min_f64, tb_proj_psd_f64, tb_sqrt_psd_f64);

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
/// The library serialises its entry points with one lock, so the safe wrappers may be used from several threads
/// (cargo runs `#[test]`s in parallel); the reference's backends keep thread-local managers instead (cuda_mgr.rs:113).
pub fn ensure_init() {
    use std::sync::Once;
    static INIT: Once = Once::new();
    INIT.call_once(|| {
        check(unsafe { tb_init(-1) }, "tb_init"); // -1: $LOCAL_RANK or device 0
        log::info!("totsu_b200: backend initialised");
    });
}
