//! [`B200T`]: `LinAlg` (totsu_core/src/solver/linalg.rs:10-68) + `LinAlgEx` (totsu_core/src/linalg_ex.rs:7-66).
//! One C-ABI call per trait function; CPU twin: totsu_f64lapack/src/f64lapack.rs:15-191.

use crate::b200_slice::B200SliceT;
use crate::ffi::*;
use std::marker::PhantomData;
use totsu_core::solver::{LinAlg, SliceLike};
use totsu_core::LinAlgEx;

/// [`LinAlgEx`] implementation on hand-written sm_100a kernels, element type `F` = `f32` or `f64`.
#[derive(Clone)]
pub struct B200T<F: Elem> {
    _ph: PhantomData<F>,
}

/// `f32` backend - the drop-in sibling of `F32CUDA`.
pub type B200 = B200T<f32>;
/// `f64` backend - the device twin of `F64LAPACK` (the reference's default `eps_acc = 1e-6` converges in it).
pub type B200F64 = B200T<f64>;

impl<F: Elem> LinAlg for B200T<F> {
    type F = F;
    type Sl = B200SliceT<F>;

    fn norm(x: &Self::Sl) -> F {
        let mut out = F::default();
        check(unsafe { F::norm(x.view(), &mut out) }, "tb_norm");
        out
    }

    fn copy(x: &Self::Sl, y: &mut Self::Sl) {
        assert_eq!(x.len(), y.len());
        check(unsafe { F::copy(x.view(), y.view()) }, "tb_copy");
    }

    fn scale(alpha: F, x: &mut Self::Sl) {
        check(unsafe { F::scale(alpha, x.view()) }, "tb_scale");
    }

    fn add(alpha: F, x: &Self::Sl, y: &mut Self::Sl) {
        assert_eq!(x.len(), y.len());
        check(unsafe { F::add(alpha, x.view(), y.view()) }, "tb_add");
    }

    fn adds(s: F, y: &mut Self::Sl) {
        check(unsafe { F::adds(s, y.view()) }, "tb_adds");
    }

    fn abssum(x: &Self::Sl, incx: usize) -> F {
        let mut out = F::default();
        check(unsafe { F::abssum(x.view(), incx, &mut out) }, "tb_abssum");
        out
    }

    fn transform_di(alpha: F, mat: &Self::Sl, x: &Self::Sl, beta: F, y: &mut Self::Sl) {
        assert_eq!(mat.len(), x.len());
        assert_eq!(mat.len(), y.len());
        check(unsafe { F::transform_di(alpha, mat.view(), x.view(), beta, y.view()) }, "tb_transform_di");
    }
}

enum Closure {
    General,
    KeepPositive,
    Sqrt,
}

fn recognise<F: Elem, M: Fn(F) -> Option<F>>(map: &M) -> Closure {
    const PROBES: [f64; 14] = [1e-30, 3e-21, 1e-12, 7e-7, 1e-3, 0.25, 1.0, 2.0, 9.0, 1234.5, 1e6, 3e12, 1e20, 1e30];
    let (mut keep, mut root) = (true, true);
    for pd in PROBES.iter() {
        let e = F::from(*pd).unwrap();
        match map(e) {
            None => return Closure::General,
            Some(out) => {
                keep = keep && out == e;
                root = root && out == e.sqrt();
            }
        }
    }
    if keep {
        Closure::KeepPositive
    } else if root {
        Closure::Sqrt
    } else {
        Closure::General
    }
}

impl<F: Elem> LinAlgEx for B200T<F> {
    fn transform_ge(transpose: bool, n_row: usize, n_col: usize, alpha: F, mat: &Self::Sl, x: &Self::Sl, beta: F, y: &mut Self::Sl) {
        assert_eq!(mat.len(), n_row * n_col);
        check(
            unsafe { F::transform_ge(transpose as i32, n_row, n_col, alpha, mat.view(), x.view(), beta, y.view()) },
            "tb_transform_ge",
        );
    }

    fn transform_sp(n: usize, alpha: F, mat: &Self::Sl, x: &Self::Sl, beta: F, y: &mut Self::Sl) {
        assert_eq!(mat.len(), n * (n + 1) / 2);
        check(unsafe { F::transform_sp(n, alpha, mat.view(), x.view(), beta, y.view()) }, "tb_transform_sp");
    }

    fn map_eig_worklen(n: usize) -> usize {
        unsafe { tb_map_eig_worklen(n) }
    }

    fn map_eig<M>(mat: &mut Self::Sl, scale_diag: Option<F>, eps_zero: F, work: &mut Self::Sl, map: M)
    where
        M: Fn(F) -> Option<F>,
    {
        let sn = mat.len();
        let n = (((8 * sn + 1) as f64).sqrt() as usize - 1) / 2;
        assert_eq!(n * (n + 1) / 2, sn);
        assert!(work.len() >= Self::map_eig_worklen(n));
        let (has_scale, sd) = match scale_diag {
            Some(s) => (1, s),
            None => (0, F::one()),
        };
        // The two closures the reference itself passes are recognised by probing them on positive arguments spanning the
        // floating-point range (only eigenvalues > 0 ever reach the closure, f64lapack.rs:86-107) and served GEMM-only:
        //   e -> Some(e)       ConePSD::proj (cone_psd.rs:69-76, scale_diag = Some(sqrt 2))  -> tb_proj_psd (matrix-sign iteration)
        //   e -> Some(sqrt e)  MatBuild::set_sqrt (matbuild/mod.rs:231-238, scale_diag None) -> tb_sqrt_psd (coupled Newton-Schulz)
        match recognise(&map) {
            Closure::KeepPositive if scale_diag == Some((F::one() + F::one()).sqrt()) => {
                check(unsafe { F::proj_psd(mat.view(), eps_zero, work.view()) }, "tb_proj_psd");
                return;
            }
            Closure::Sqrt if scale_diag.is_none() => {
                check(unsafe { F::sqrt_psd(mat.view(), eps_zero, work.view()) }, "tb_sqrt_psd");
                return;
            }
            _ => {}
        }
        // eigendecomposition on the device; only the n eigenvalues cross to the host for the closure
        let mut eigs = vec![F::zero(); n];
        check(
            unsafe { F::map_eig_begin(mat.view(), has_scale, sd, eps_zero, work.view(), eigs.as_mut_ptr()) },
            "tb_map_eig_begin",
        );
        // dsyevr(range = V, (0, +inf]) hands only the positive eigenpairs to `map` (f64lapack.rs:86-107)
        let mut keep = vec![0u8; n];
        for i in 0..n {
            if eigs[i] > F::zero() {
                if let Some(e) = map(eigs[i]) {
                    eigs[i] = e;
                    keep[i] = 1;
                }
            }
        }
        check(
            unsafe { F::map_eig_finish(mat.view(), has_scale, sd, work.view(), eigs.as_ptr(), keep.as_ptr()) },
            "tb_map_eig_finish",
        );
    }
}
