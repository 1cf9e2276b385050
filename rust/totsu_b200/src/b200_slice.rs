//! [`B200Slice`]: the `SliceLike` of the B200 backend (trait: totsu_core/src/solver/slicelike.rs:9-70; role of
//! totsu_f32cuda/src/f32cuda_slice.rs, redesigned).
//!
//! `B200Slice` is a transparent wrapper of the host slice `[f32]` - a dynamically sized type, like the reference's own
//! `impl SliceLike for [F]` (slicelike.rs:191-250).  A `&B200Slice` therefore *is* the (pointer, length) of a host
//! sub-slice, and `split_ref`/`split_mut` are plain `split_at`s: no `Pin<Box<..>>`, no HashMap, no heap traffic per
//! `splitm!`.  The device mirror lives in libtotsu_b200.so, keyed by the host address range of the root slice that
//! `new_ref`/`new_mut` wrapped (`tb_buf_wrap`); any sub-slice resolves to a `(handle, offset, len)` view with
//! `tb_view_of_host`.  Host/device coherence is tracked inside the library per element range.
//!
//! Lifetime protocol ("`drop` shall be called when the wrapper drops", slicelike.rs:18-19,42-46): every non-empty
//! wrapper holds one reference on its root (`tb_buf_wrap` = 1, each non-empty child of a split = `tb_buf_retain`);
//! `SliceLike::drop` gives it back (`tb_buf_release`).  The library flushes device-newer ranges to the caller's slice
//! and frees the mirror when the last wrapper is gone, which is when `Solver::solve` reads the solution out of
//! `work` (solver.rs:315-320).

use crate::ffi::*;
use std::os::raw::c_void;
use totsu_core::solver::{SliceLike, SliceMut, SliceRef};

/// `f32` slice with a device mirror on the B200, [`SliceLike`] implementation for [`crate::B200`]`::Sl`.
#[repr(transparent)]
pub struct B200Slice {
    host: [f32],
}

impl B200Slice {
    #[inline]
    fn from_host(s: &[f32]) -> &B200Slice {
        // SAFETY: repr(transparent) over [f32]
        unsafe { &*(s as *const [f32] as *const B200Slice) }
    }
    #[inline]
    fn from_host_mut(s: &mut [f32]) -> &mut B200Slice {
        unsafe { &mut *(s as *mut [f32] as *mut B200Slice) }
    }

    /// The device view of this slice.
    #[inline]
    pub fn view(&self) -> tb_view {
        let mut v = tb_view { buf: 0, off: 0, len: 0 };
        check(
            unsafe { tb_view_of_host(TB_F32, self.host.as_ptr() as *const c_void, self.host.len(), &mut v) },
            "tb_view_of_host",
        );
        v
    }

    fn wrap(ptr: *mut f32, len: usize, is_mut: bool) {
        ensure_init();
        if len > 0 {
            let mut h: tb_handle = 0;
            check(unsafe { tb_buf_wrap(TB_F32, ptr as *mut c_void, len, is_mut as i32, &mut h) }, "tb_buf_wrap");
        }
    }
}

impl SliceLike for B200Slice {
    type F = f32;

    fn new_ref(s: &[f32]) -> SliceRef<'_, B200Slice> {
        B200Slice::wrap(s.as_ptr() as *mut f32, s.len(), false);
        unsafe { SliceRef::new(B200Slice::from_host(s)) }
    }

    fn new_mut(s: &mut [f32]) -> SliceMut<'_, B200Slice> {
        B200Slice::wrap(s.as_mut_ptr(), s.len(), true);
        unsafe { SliceMut::new(B200Slice::from_host_mut(s)) }
    }

    fn split_ref(&self, mid: usize) -> (SliceRef<'_, B200Slice>, SliceRef<'_, B200Slice>) {
        let (a, b) = self.host.split_at(mid);
        let n = (!a.is_empty()) as i32 + (!b.is_empty()) as i32;
        if n > 0 {
            check(unsafe { tb_buf_retain(self.view().buf, n) }, "tb_buf_retain");
        }
        unsafe { (SliceRef::new(B200Slice::from_host(a)), SliceRef::new(B200Slice::from_host(b))) }
    }

    fn split_mut(&mut self, mid: usize) -> (SliceMut<'_, B200Slice>, SliceMut<'_, B200Slice>) {
        let root = if self.host.is_empty() { 0 } else { self.view().buf };
        let (a, b) = self.host.split_at_mut(mid);
        let n = (!a.is_empty()) as i32 + (!b.is_empty()) as i32;
        if n > 0 {
            check(unsafe { tb_buf_retain(root, n) }, "tb_buf_retain");
        }
        unsafe { (SliceMut::new(B200Slice::from_host_mut(a)), SliceMut::new(B200Slice::from_host_mut(b))) }
    }

    fn drop(&self) {
        if !self.host.is_empty() {
            check(unsafe { tb_buf_release(self.view().buf) }, "tb_buf_release");
        }
    }

    fn len(&self) -> usize {
        self.host.len()
    }

    fn get_ref(&self) -> &[f32] {
        check(unsafe { tb_host_ref(self.view()) }, "tb_host_ref");
        &self.host
    }

    fn get_mut(&mut self) -> &mut [f32] {
        check(unsafe { tb_host_mut(self.view()) }, "tb_host_mut");
        &mut self.host
    }

    // One 4-byte mailbox read / one kernel-argument write instead of two nested splits + get_ref/get_mut
    // (default impl: slicelike.rs:54-69).
    fn get(&self, idx: usize) -> f32 {
        let mut out = 0f32;
        check(unsafe { tb_get1_f32(self.view(), idx, &mut out) }, "tb_get1_f32");
        out
    }

    fn set(&mut self, idx: usize, val: f32) {
        check(unsafe { tb_set1_f32(self.view(), idx, val) }, "tb_set1_f32");
    }
}
