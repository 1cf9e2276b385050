//! Device-resident `Operator` / `Cone` implementors handed to the *unmodified* `Solver::solve` - the extension route
//! examples/imgnr_udef/src/main.rs:62-67 demonstrates.  They remove the O(#blocks) launch pattern of the stock
//! front-ends (totsu/src/problem/socp.rs:83-124) and the host loops of `ConeRPos` / `calc_precond`.

use crate::b200::B200T;
use crate::b200_slice::B200SliceT;
use crate::ffi::*;
use std::marker::PhantomData;
use totsu_core::solver::{Cone, Operator, SliceLike, SliceRef};

/// One stacked dense column-major `A` (optionally this rank's row shard of it) as an [`Operator`].
pub struct DenseOpT<'a, F: Elem> {
    h: tb_handle,
    n_row_total: usize,
    n_col: usize,
    _array: SliceRef<'a, B200SliceT<F>>,
}
pub type DenseOp<'a> = DenseOpT<'a, f32>;
pub type DenseOpF64<'a> = DenseOpT<'a, f64>;

impl<'a, F: Elem> DenseOpT<'a, F> {
    /// `array`: column-major `n_row x n_col` (lda = n_row), the rows `[row_offset, row_offset + n_row)` of an
    /// `n_row_total x n_col` matrix (`n_row_total = n_row` on one GPU).
    pub fn new(array: &'a [F], n_row: usize, n_col: usize, row_offset: usize, n_row_total: usize) -> Self {
        assert_eq!(array.len(), n_row * n_col);
        let sl = B200SliceT::<F>::new_ref(array); // upload point, like MatOp::new (matop.rs:66-74)
        let mut h: tb_handle = 0;
        check(
            unsafe { tb_denseop_create(F::DTYPE, sl.view(), n_row, n_col, row_offset, n_row_total, &mut h) },
            "tb_denseop_create",
        );
        DenseOpT { h, n_row_total, n_col, _array: sl }
    }
}

impl<'a, F: Elem> Drop for DenseOpT<'a, F> {
    fn drop(&mut self) {
        unsafe { tb_denseop_destroy(self.h) };
    }
}

impl<'a, F: Elem> Operator<B200T<F>> for DenseOpT<'a, F> {
    fn size(&self) -> (usize, usize) {
        (self.n_row_total, self.n_col)
    }
    fn op(&self, alpha: F, x: &B200SliceT<F>, beta: F, y: &mut B200SliceT<F>) {
        check(unsafe { F::denseop_apply(self.h, 0, alpha, x.view(), beta, y.view()) }, "tb_denseop_apply");
    }
    fn trans_op(&self, alpha: F, x: &B200SliceT<F>, beta: F, y: &mut B200SliceT<F>) {
        check(unsafe { F::denseop_apply(self.h, 1, alpha, x.view(), beta, y.view()) }, "tb_denseop_apply");
    }
    fn absadd_cols(&self, tau: &mut B200SliceT<F>) {
        check(unsafe { F::denseop_absadd_cols(self.h, tau.view()) }, "tb_denseop_absadd_cols");
    }
    fn absadd_rows(&self, sigma: &mut B200SliceT<F>) {
        check(unsafe { F::denseop_absadd_rows(self.h, sigma.view()) }, "tb_denseop_absadd_rows");
    }
}

/// Zero / RPos / SOC / RotSOC / PSD blocks laid out back to back (the shape of `ProbSOCPCone`, socp.rs:296-313),
/// projected in ONE launch (+ one GEMM-only sign iteration per PSD block).
///
/// The PSD work area (2k^2 + k elements, cone_psd.rs:32-38) is a device-only buffer allocated ONCE here: `proj` neither
/// wraps nor releases anything, so the first projection of an iteration stays parked until the second arrives and the two
/// run as one batch on the tensor cores (csrc/cone.cu "pairing of the two projections").
pub struct ProductConeT<F: Elem> {
    h: tb_handle,
    eps_zero: F,
    psd_work: tb_view,
    blocks: Vec<(i32, usize)>,
    _ph: PhantomData<F>,
}
pub type ProductCone = ProductConeT<f32>;
pub type ProductConeF64 = ProductConeT<f64>;

impl<F: Elem> ProductConeT<F> {
    pub fn new(blocks: &[(i32, usize)], eps_zero: F) -> Self {
        let bl: Vec<tb_cone_block> = blocks.iter().map(|&(typ, len)| tb_cone_block { typ, reserved: 0, len: len as u64 }).collect();
        let mut h: tb_handle = 0;
        ensure_init();
        check(unsafe { tb_cone_create(bl.as_ptr(), bl.len(), &mut h) }, "tb_cone_create");
        let mut wl = 0;
        for &(typ, len) in blocks {
            if typ == TB_CONE_PSD {
                let k = (((8 * len + 1) as f64).sqrt() as usize - 1) / 2;
                wl = wl.max(unsafe { tb_map_eig_worklen(k) });
            }
        }
        let mut psd_work = tb_view { buf: 0, off: 0, len: 0 };
        if wl > 0 {
            let mut wh: tb_handle = 0;
            check(unsafe { tb_buf_alloc(F::DTYPE, wl, &mut wh) }, "tb_buf_alloc");
            psd_work = tb_view { buf: wh, off: 0, len: wl };
        }
        ProductConeT { h, eps_zero, psd_work, blocks: blocks.to_vec(), _ph: PhantomData }
    }
}

impl<F: Elem> Drop for ProductConeT<F> {
    fn drop(&mut self) {
        unsafe {
            if self.psd_work.buf != 0 {
                tb_buf_release(self.psd_work.buf);
            }
            tb_cone_destroy(self.h)
        };
    }
}

impl<F: Elem> Cone<B200T<F>> for ProductConeT<F> {
    fn proj(&mut self, dual_cone: bool, x: &mut B200SliceT<F>) -> Result<(), ()> {
        // arguments are validated at submit time, before the projection may be parked: TB_ERR_ARG can only come from here
        let st = unsafe { F::cone_proj(self.h, dual_cone as i32, x.view(), self.eps_zero, self.psd_work) };
        if st == TB_ERR_ARG {
            return Err(()); // -> SolverError::ConeFailure (solver.rs:548-549)
        }
        check(st, "tb_cone_proj");
        Ok(())
    }

    // cone.rs:20-29: split dp_tau into the blocks and call `group` once per block of size > 1.  `Solver` passes the
    // min-fill closure (solver.rs:509-518), which `tb_cone_group_min` computes for every block in one launch; any OTHER
    // closure gets the trait's contract literally - one `group` call per block on its sub-slice.
    fn product_group<G: Fn(&mut B200SliceT<F>) + Copy>(&self, dp_tau: &mut B200SliceT<F>, group: G) {
        if is_min_fill(group) {
            check(unsafe { F::cone_group_min(self.h, dp_tau.view()) }, "tb_cone_group_min");
            return;
        }
        // same idiom as ProbSOCPCone::product_group (totsu/src/problem/socp.rs:315-331): re-split from the front per block
        let mut done = 0;
        for &(typ, len) in self.blocks.iter() {
            let (_t_done, mut spl) = dp_tau.split_mut(done);
            let (mut t_blk, _) = spl.split_mut(len);
            done += len;
            // ConeZero / ConeRPos do not group (cone_zero.rs:46-49, cone_rpos.rs:47-50)
            if len > 1 && typ != TB_CONE_ZERO && typ != TB_CONE_RPOS {
                group(&mut t_blk);
            }
        }
    }
}

/// Probe a `group` closure on a 3-element device slice: the solver's closure fills the group with its minimum.
fn is_min_fill<F: Elem, G: Fn(&mut B200SliceT<F>) + Copy>(group: G) -> bool {
    let three = F::one() + F::one() + F::one();
    let mut probe = [three, F::one(), F::one() + F::one()];
    {
        let mut sl = B200SliceT::<F>::new_mut(&mut probe);
        group(&mut sl);
    }
    probe[0] == F::one() && probe[1] == F::one() && probe[2] == F::one()
}
