//! [`B200SliceT`]: the `SliceLike` of the B200 backend (trait: totsu_core/src/solver/slicelike.rs:9-70; role of
//! totsu_f32cuda/src/f32cuda_slice.rs, redesigned).
//!
//! `B200SliceT<F>` is a transparent wrapper of the host slice `[F]` - a dynamically sized type, like the reference's own
//! `impl SliceLike for [F]` (slicelike.rs:191-250).  A `&B200SliceT<F>` therefore *is* the (pointer, length) of a host
//! sub-slice, and `split_ref`/`split_mut` are plain `split_at`s: no `Pin<Box<..>>`, no HashMap, no heap traffic per
//! `splitm!`.  The device mirror lives in libtotsu_b200.so, keyed by the host address range of the root slice that
//! `new_ref`/`new_mut` wrapped (`tb_buf_wrap`); any sub-slice resolves to a `(handle, offset, len)` view with
//! `tb_view_of_host` - a pure table lookup (ordered map by start address) that neither launches nor drains anything, so
//! the library's op/trans_op pairing and PSD pairing keep firing behind `splitm!`.  Host/device coherence is tracked
//! inside the library per element range.
//!
//! Lifetime protocol ("`drop` shall be called when the wrapper drops", slicelike.rs:18-19,42-46): every non-empty
//! wrapper holds one reference on its root (`tb_buf_wrap` = 1, each non-empty child of a split = `tb_buf_retain`);
//! `SliceLike::drop` gives it back (`tb_buf_release`).  The library flushes device-newer ranges to the caller's slice
//! and frees the mirror when the last wrapper is gone, which is when `Solver::solve` reads the solution out of
//! `work` (solver.rs:315-320).  The same call protocol is exercised on the GPU by the C++ host mirror in
//! "shim-protocol" mode (totsu_b200/host/linalg.hpp, tests/test_shim_protocol_gpu.py).

use crate::ffi::*;
use std::os::raw::c_void;
use totsu_core::solver::{SliceLike, SliceMut, SliceRef};

/// Slice of `F` with a device mirror on the B200; [`SliceLike`] implementation for [`crate::B200T`]`::Sl`.
#[repr(transparent)]
pub struct B200SliceT<F: Elem> {
    host: [F],
}

/// `f32` slice ([`crate::B200`]) and `f64` slice ([`crate::B200F64`]).
pub type B200Slice = B200SliceT<f32>;
pub type B200SliceF64 = B200SliceT<f64>;

impl<F: Elem> B200SliceT<F> {
    #[inline]
    fn from_host(s: &[F]) -> &B200SliceT<F> {
        // SAFETY: repr(transparent) over [F]
        unsafe { &*(s as *const [F] as *const B200SliceT<F>) }
    }
    #[inline]
    fn from_host_mut(s: &mut [F]) -> &mut B200SliceT<F> {
        unsafe { &mut *(s as *mut [F] as *mut B200SliceT<F>) }
    }

    /// The device view of this slice.
    #[inline]
    pub fn view(&self) -> tb_view {
        let mut v = tb_view { buf: 0, off: 0, len: 0 };
        check(
            unsafe { tb_view_of_host(F::DTYPE, self.host.as_ptr() as *const c_void, self.host.len(), &mut v) },
            "tb_view_of_host",
        );
        v
    }

    fn wrap(ptr: *mut F, len: usize, is_mut: bool) {
        ensure_init();
        if len > 0 {
            let mut h: tb_handle = 0;
            check(unsafe { tb_buf_wrap(F::DTYPE, ptr as *mut c_void, len, is_mut as i32, &mut h) }, "tb_buf_wrap");
        }
    }
}

impl<F: Elem> SliceLike for B200SliceT<F> {
    type F = F;

    fn new_ref(s: &[F]) -> SliceRef<'_, B200SliceT<F>> {
        B200SliceT::<F>::wrap(s.as_ptr() as *mut F, s.len(), false);
        unsafe { SliceRef::new(B200SliceT::from_host(s)) }
    }

    fn new_mut(s: &mut [F]) -> SliceMut<'_, B200SliceT<F>> {
        B200SliceT::<F>::wrap(s.as_mut_ptr(), s.len(), true);
        unsafe { SliceMut::new(B200SliceT::from_host_mut(s)) }
    }

    fn split_ref(&self, mid: usize) -> (SliceRef<'_, B200SliceT<F>>, SliceRef<'_, B200SliceT<F>>) {
        let (a, b) = self.host.split_at(mid);
        let n = (!a.is_empty()) as i32 + (!b.is_empty()) as i32;
        if n > 0 {
            check(unsafe { tb_buf_retain(self.view().buf, n) }, "tb_buf_retain");
        }
        unsafe { (SliceRef::new(B200SliceT::from_host(a)), SliceRef::new(B200SliceT::from_host(b))) }
    }

    fn split_mut(&mut self, mid: usize) -> (SliceMut<'_, B200SliceT<F>>, SliceMut<'_, B200SliceT<F>>) {
        let root = if self.host.is_empty() { 0 } else { self.view().buf };
        let (a, b) = self.host.split_at_mut(mid);
        let n = (!a.is_empty()) as i32 + (!b.is_empty()) as i32;
        if n > 0 {
            check(unsafe { tb_buf_retain(root, n) }, "tb_buf_retain");
        }
        unsafe { (SliceMut::new(B200SliceT::from_host_mut(a)), SliceMut::new(B200SliceT::from_host_mut(b))) }
    }

    fn drop(&self) {
        if !self.host.is_empty() {
            check(unsafe { tb_buf_release(self.view().buf) }, "tb_buf_release");
        }
    }

    fn len(&self) -> usize {
        self.host.len()
    }

    fn get_ref(&self) -> &[F] {
        check(unsafe { tb_host_ref(self.view()) }, "tb_host_ref");
        &self.host
    }

    fn get_mut(&mut self) -> &mut [F] {
        check(unsafe { tb_host_mut(self.view()) }, "tb_host_mut");
        &mut self.host
    }

    // One mailbox read / one kernel-argument write instead of two nested splits + get_ref/get_mut
    // (default impl: slicelike.rs:54-69).
    fn get(&self, idx: usize) -> F {
        let mut out = F::default();
        check(unsafe { F::get1(self.view(), idx, &mut out) }, "tb_get1");
        out
    }

    fn set(&mut self, idx: usize, val: F) {
        check(unsafe { F::set1(self.view(), idx, val) }, "tb_set1");
    }
}
