/*!
`totsu_b200` links [`totsu_core`](https://crates.io/crates/totsu_core) to hand-written sm_100a CUDA kernels
(`libtotsu_b200.so`, C ABI in `include/totsu_b200.h`) - a sibling of `totsu_f64lapack` and `totsu_f32cuda`.

* [`B200`] (`f32`) and [`B200F64`] (`f64`) implement `LinAlg` + `LinAlgEx`, [`B200Slice`] / [`B200SliceF64`] implement
  `SliceLike`: `Solver<B200>`, `MatOp<B200>`, `ConePSD<B200>`, `MatBuild<B200>` and the `ProbLP/QP/QCQP/SOCP/SDP<B200>`
  front-ends run unmodified.
* [`DenseOp`] and [`ProductCone`] (and their `F64` twins) are device-resident `Operator` / `Cone` implementors for large
  dense problems (one stacked `A`, a whole product cone per launch), passed to the same unmodified `Solver::solve`.

```no_run
use totsu::prelude::*;
use totsu::*;
use totsu_b200::B200;

type La = B200;
type AMatBuild = MatBuild<La>;
type AProbQP = ProbQP<La>;
type ASolver = Solver<La>;
// ... build the QP exactly as in totsu_f32cuda's crate example (totsu_f32cuda/src/lib.rs:31-76) ...
```

NOTE: this crate was written in an image without `rustc`/`cargo`; it is a reviewed-by-eye binding, compiled nowhere yet.
The same call PROTOCOL (one `tb_view_of_host` per operand per call, `tb_buf_retain` / `tb_buf_release` per split child) is
exercised on the GPU by the C++ host mirror in `totsu_b200/host/` in its "shim-protocol" mode.
*/

mod b200;
mod b200_slice;
pub mod ffi;
mod fused;

pub use b200::{B200, B200F64, B200T};
pub use b200_slice::{B200Slice, B200SliceF64, B200SliceT};
pub use ffi::{Elem, TB_CONE_PSD, TB_CONE_ROTSOC, TB_CONE_RPOS, TB_CONE_SOC, TB_CONE_ZERO};
pub use fused::{DenseOp, DenseOpF64, DenseOpT, ProductCone, ProductConeF64, ProductConeT};
