//! [`B200`]: `LinAlg` (totsu_core/src/solver/linalg.rs:10-68) + `LinAlgEx` (totsu_core/src/linalg_ex.rs:7-66).
//! One C-ABI call per trait function; CPU twin: totsu_f64lapack/src/f64lapack.rs:15-191.

use crate::b200_slice::B200Slice;
use crate::ffi::*;
use totsu_core::solver::{LinAlg, SliceLike};
use totsu_core::LinAlgEx;

/// `f32`-specific [`LinAlgEx`] implementation on hand-written sm_100a kernels.
#[derive(Clone)]
pub struct B200;

impl LinAlg for B200 {
    type F = f32;
    type Sl = B200Slice;

    fn norm(x: &B200Slice) -> f32 {
        let mut out = 0f32;
        check(unsafe { tb_norm_f32(x.view(), &mut out) }, "tb_norm_f32");
        out
    }

    fn copy(x: &B200Slice, y: &mut B200Slice) {
        assert_eq!(x.len(), y.len());
        check(unsafe { tb_copy_f32(x.view(), y.view()) }, "tb_copy_f32");
    }

    fn scale(alpha: f32, x: &mut B200Slice) {
        check(unsafe { tb_scale_f32(alpha, x.view()) }, "tb_scale_f32");
    }

    fn add(alpha: f32, x: &B200Slice, y: &mut B200Slice) {
        assert_eq!(x.len(), y.len());
        check(unsafe { tb_add_f32(alpha, x.view(), y.view()) }, "tb_add_f32");
    }

    fn adds(s: f32, y: &mut B200Slice) {
        check(unsafe { tb_adds_f32(s, y.view()) }, "tb_adds_f32");
    }

    fn abssum(x: &B200Slice, incx: usize) -> f32 {
        let mut out = 0f32;
        check(unsafe { tb_abssum_f32(x.view(), incx, &mut out) }, "tb_abssum_f32");
        out
    }

    fn transform_di(alpha: f32, mat: &B200Slice, x: &B200Slice, beta: f32, y: &mut B200Slice) {
        assert_eq!(mat.len(), x.len());
        assert_eq!(mat.len(), y.len());
        check(unsafe { tb_transform_di_f32(alpha, mat.view(), x.view(), beta, y.view()) }, "tb_transform_di_f32");
    }
}

impl LinAlgEx for B200 {
    fn transform_ge(transpose: bool, n_row: usize, n_col: usize, alpha: f32, mat: &B200Slice, x: &B200Slice, beta: f32, y: &mut B200Slice) {
        assert_eq!(mat.len(), n_row * n_col);
        check(
            unsafe { tb_transform_ge_f32(transpose as i32, n_row, n_col, alpha, mat.view(), x.view(), beta, y.view()) },
            "tb_transform_ge_f32",
        );
    }

    fn transform_sp(n: usize, alpha: f32, mat: &B200Slice, x: &B200Slice, beta: f32, y: &mut B200Slice) {
        assert_eq!(mat.len(), n * (n + 1) / 2);
        check(unsafe { tb_transform_sp_f32(n, alpha, mat.view(), x.view(), beta, y.view()) }, "tb_transform_sp_f32");
    }

    fn map_eig_worklen(n: usize) -> usize {
        unsafe { tb_map_eig_worklen(n) }
    }

    fn map_eig<M>(mat: &mut B200Slice, scale_diag: Option<f32>, eps_zero: f32, work: &mut B200Slice, map: M)
    where
        M: Fn(f32) -> Option<f32>,
    {
        let sn = mat.len();
        let n = (((8 * sn + 1) as f64).sqrt() as usize - 1) / 2;
        assert_eq!(n * (n + 1) / 2, sn);
        assert!(work.len() >= Self::map_eig_worklen(n));
        let (has_scale, sd) = match scale_diag {
            Some(s) => (1, s),
            None => (0, 1f32),
        };
        // eigendecomposition on the device; only the n eigenvalues cross to the host for the closure
        let mut eigs = vec![0f32; n];
        check(
            unsafe { tb_map_eig_begin_f32(mat.view(), has_scale, sd, eps_zero, work.view(), eigs.as_mut_ptr()) },
            "tb_map_eig_begin_f32",
        );
        // dsyevr(range = V, (0, +inf]) hands only the positive eigenpairs to `map` (f64lapack.rs:86-107)
        let mut keep = vec![0u8; n];
        for i in 0..n {
            if eigs[i] > 0. {
                if let Some(e) = map(eigs[i]) {
                    eigs[i] = e;
                    keep[i] = 1;
                }
            }
        }
        check(
            unsafe { tb_map_eig_finish_f32(mat.view(), has_scale, sd, work.view(), eigs.as_ptr(), keep.as_ptr()) },
            "tb_map_eig_finish_f32",
        );
    }
}
