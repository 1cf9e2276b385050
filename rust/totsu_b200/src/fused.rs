//! Device-resident `Operator` / `Cone` implementors handed to the *unmodified* `Solver::solve` - the extension route
//! examples/imgnr_udef/src/main.rs:62-67 demonstrates.  They remove the O(#blocks) launch pattern of the stock
//! front-ends (totsu/src/problem/socp.rs:83-124) and the host loops of `ConeRPos` / `calc_precond`.

use crate::b200::B200;
use crate::b200_slice::B200Slice;
use crate::ffi::*;
use totsu_core::solver::{Cone, Operator, SliceLike, SliceRef};

/// One stacked dense column-major `A` (optionally this rank's row shard of it) as an [`Operator`].
pub struct DenseOp<'a> {
    h: tb_handle,
    n_row_total: usize,
    n_col: usize,
    _array: SliceRef<'a, B200Slice>,
}

impl<'a> DenseOp<'a> {
    /// `array`: column-major `n_row x n_col` (lda = n_row), the rows `[row_offset, row_offset + n_row)` of an
    /// `n_row_total x n_col` matrix (`n_row_total = n_row` on one GPU).
    pub fn new(array: &'a [f32], n_row: usize, n_col: usize, row_offset: usize, n_row_total: usize) -> Self {
        assert_eq!(array.len(), n_row * n_col);
        let sl = B200Slice::new_ref(array); // upload point, like MatOp::new (matop.rs:66-74)
        let mut h: tb_handle = 0;
        check(
            unsafe { tb_denseop_create(TB_F32, sl.view(), n_row, n_col, row_offset, n_row_total, &mut h) },
            "tb_denseop_create",
        );
        DenseOp { h, n_row_total, n_col, _array: sl }
    }
}

impl<'a> Drop for DenseOp<'a> {
    fn drop(&mut self) {
        unsafe { tb_denseop_destroy(self.h) };
    }
}

impl<'a> Operator<B200> for DenseOp<'a> {
    fn size(&self) -> (usize, usize) {
        (self.n_row_total, self.n_col)
    }
    fn op(&self, alpha: f32, x: &B200Slice, beta: f32, y: &mut B200Slice) {
        check(unsafe { tb_denseop_apply_f32(self.h, 0, alpha, x.view(), beta, y.view()) }, "tb_denseop_apply_f32");
    }
    fn trans_op(&self, alpha: f32, x: &B200Slice, beta: f32, y: &mut B200Slice) {
        check(unsafe { tb_denseop_apply_f32(self.h, 1, alpha, x.view(), beta, y.view()) }, "tb_denseop_apply_f32");
    }
    fn absadd_cols(&self, tau: &mut B200Slice) {
        check(unsafe { tb_denseop_absadd_cols_f32(self.h, tau.view()) }, "tb_denseop_absadd_cols_f32");
    }
    fn absadd_rows(&self, sigma: &mut B200Slice) {
        check(unsafe { tb_denseop_absadd_rows_f32(self.h, sigma.view()) }, "tb_denseop_absadd_rows_f32");
    }
}

/// Zero / RPos / SOC / RotSOC / PSD blocks laid out back to back (the shape of `ProbSOCPCone`, socp.rs:296-313),
/// projected in ONE launch (+ one GEMM-only sign iteration per PSD block).
pub struct ProductCone {
    h: tb_handle,
    eps_zero: f32,
    psd_work: Vec<f32>,
}

impl ProductCone {
    pub fn new(blocks: &[(i32, usize)], eps_zero: f32) -> Self {
        let bl: Vec<tb_cone_block> = blocks.iter().map(|&(typ, len)| tb_cone_block { typ, reserved: 0, len: len as u64 }).collect();
        let mut h: tb_handle = 0;
        crate::ffi::ensure_init();
        check(unsafe { tb_cone_create(bl.as_ptr(), bl.len(), &mut h) }, "tb_cone_create");
        let mut wl = 0;
        for &(typ, len) in blocks {
            if typ == TB_CONE_PSD {
                let k = (((8 * len + 1) as f64).sqrt() as usize - 1) / 2;
                wl = wl.max(unsafe { tb_map_eig_worklen(k) });
            }
        }
        ProductCone { h, eps_zero, psd_work: vec![0.; wl] }
    }
}

impl Drop for ProductCone {
    fn drop(&mut self) {
        unsafe { tb_cone_destroy(self.h) };
    }
}

impl Cone<B200> for ProductCone {
    fn proj(&mut self, dual_cone: bool, x: &mut B200Slice) -> Result<(), ()> {
        let work = B200Slice::new_mut(&mut self.psd_work);
        let st = unsafe { tb_cone_proj_f32(self.h, dual_cone as i32, x.view(), self.eps_zero, work.view()) };
        if st == TB_ERR_ARG {
            return Err(()); // -> SolverError::ConeFailure (solver.rs:548-549)
        }
        check(st, "tb_cone_proj_f32");
        Ok(())
    }

    // The solver's `group` closure is the min-fill (solver.rs:509-518); it runs on the device for every block of
    // size > 1, which is what calling `group` once per such block would compute (cone.rs:20-29).
    fn product_group<G: Fn(&mut B200Slice) + Copy>(&self, dp_tau: &mut B200Slice, _group: G) {
        check(unsafe { tb_cone_group_min_f32(self.h, dp_tau.view()) }, "tb_cone_group_min_f32");
    }
}
