/*
 * totsu_b200.h — C ABI of libtotsu_b200.so, the B200 (sm_100a) linear-algebra / cone backend for
 * Totsu's first-order conic solver.
 *
 * This is the boundary a Rust FFI crate `totsu_b200` (a sibling of `totsu_f64lapack` / `totsu_f32cuda`)
 * binds: one entry point per function of the reference's plugin traits
 *
 *     LinAlg     solver_rust_conic/totsu_core/src/solver/linalg.rs:10-68
 *     LinAlgEx   solver_rust_conic/totsu_core/src/linalg_ex.rs:7-66
 *     SliceLike  solver_rust_conic/totsu_core/src/solver/slicelike.rs:9-70
 *     Operator   solver_rust_conic/totsu_core/src/solver/operator.rs:11-156   (fused dense operator)
 *     Cone       solver_rust_conic/totsu_core/src/solver/cone.rs:9-30         (batched product cone)
 *
 * Conventions
 *   - plain C: pointers, sizes, scalars by value; no C++/torch types.
 *   - every function returns 0 (TB_OK) or a TB_ERR_* code; tb_last_error() gives the message.  The traits
 *     have no error channel, so the Rust shim asserts on the status exactly like totsu_f32cuda does with
 *     cuBLAS statuses (totsu_f32cuda/src/f32cuda.rs:38).
 *   - `LinAlg` functions are associated functions without `self` (linalg.rs:22-67), so the backend state is a
 *     process-global context bound to ONE device (one process per GPU); not re-entrant, single host thread.
 *   - element type is chosen by the suffix: _f32 (type F = f32) or _f64 (type F = f64).
 *   - vectors/matrices are "views" = (buffer handle, offset, length) in elements, the image of a Rust
 *     sub-slice produced by SliceLike::split_ref/split_mut.  There is no CPU fallback anywhere.
 */
#ifndef TOTSU_B200_H
#define TOTSU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t tb_handle;                 /* > 0 when valid */

typedef struct tb_view {
    tb_handle buf;                         /* root buffer */
    size_t    off;                         /* offset in elements */
    size_t    len;                         /* length in elements */
} tb_view;

enum { TB_OK = 0, TB_ERR_CUDA = 1, TB_ERR_ARG = 2, TB_ERR_STATE = 3, TB_ERR_NCCL = 4, TB_ERR_UNSUPPORTED = 5 };
enum { TB_F32 = 0, TB_F64 = 1 };

/* ---- lifecycle (role of totsu_f32cuda/src/cuda_mgr.rs:12-162) ------------------------------------------ */
int         tb_init(int device);           /* device ordinal; -1 = $LOCAL_RANK or 0.  Idempotent. */
int         tb_shutdown(void);
const char* tb_last_error(void);
int         tb_device_sync(void);
int         tb_get_stream(void** cuda_stream_out);   /* the cudaStream_t every call is enqueued on (for event timing) */
int         tb_sm_count(int* out);
/* number of kernels this library has launched since tb_init (bench.py's gpu_launches claim) */
int         tb_launch_count(uint64_t* out);
/* CUDA-event timing of the dominant kernel (the streaming matvec) on the library's own stream, for bench.py's
 * roofline: enable, run, then read (launch count, summed kernel milliseconds, summed algorithmic bytes). */
int         tb_prof_enable(int on);
int         tb_prof_read(uint64_t* launches, double* total_ms, double* total_bytes);
/* the same records split by kernel variant: index NN * 3 + NT of stream_kernel<T, NN, NT> (NN / NT = number of A*x / A^T*x
 * products served by the one read of A: 3 = <1,0>, 1 = <0,1>, 4 = <1,1> pair, 8 = <2,2> pair + speculated pair); arrays of 9 */
int         tb_prof_read_variants(uint64_t* launches9, double* ms9, double* bytes9);
/* Programmatic dependent launch of the small dependent kernels of an iteration (vector programs, finalize, cone, peer
 * exchange, reductions, the streaming kernel's prologue): the next kernel becomes resident while its predecessor runs and
 * blocks in griddepcontrol.wait until it has completed - launch latency leaves the critical path.  Default on (TB_PDL=0 or
 * tb_set_pdl(0): ordinary launches; results are identical either way). */
int         tb_set_pdl(int on);
/* tuning knob for tests: 0 = auto, 1 = force the generic (LDG) matvec, 2 = force the TMA matvec where legal */
int         tb_set_gemv_path(int mode);
/* Lazy op/trans_op pairing (default on): tb_denseop_apply parks the call until the opposite-direction apply on the
 * same operator arrives (SelfDualEmbed::op / trans_op, solver.rs:128-131,150-153; criteria_conv :595-598) and then
 * serves both with ONE read of A; calls in between are queued behind it and every host-visible call drains the
 * queue first, so results are bit-identical to the unfused order.  0 restores launch-at-call behaviour. */
int         tb_set_pair_fusion(int on);
int         tb_pairs_fused(uint64_t* out);           /* pairs served by one pass since tb_init */
/* Small vector commands (level-1 ops, vector operators, finalize steps, single-element get/set) are recorded and run as
 * one launch per batch - a "vector program" (csrc/vprog.cu) - submitted when anything else uses the stream or a
 * host-visible result is needed.  tb_set_vprog(0) launches one kernel per command instead (A/B, debugging);
 * tb_flush submits whatever is pending (deferred dense applies and the recorded program) without waiting;
 * tb_vprog_stats: programs launched / micro-ops recorded since tb_init. */
int         tb_set_vprog(int on);
int         tb_vprog_stats(uint64_t* launches, uint64_t* ops);
/* longest vector (elements) a barrier-capable cluster program takes; longer ones go to barrier-free full-grid programs that are
 * cut at every hazard.  Default 49152 (env TB_VPROG_MAX_N); diagnostics / A-B only. */
int         tb_set_vprog_max_n(size_t n);
int         tb_flush(void);
/* Speculative pairing (csrc/gemv.cu): while one op/trans_op pair streams A, the two products of the pair that followed
 * it last time are computed from the same staged tiles and parked; when that pair arrives it is served by the finalize
 * step alone (one read of A fewer).  Any device write overlapping the speculated inputs drops the speculation, so results
 * are bit-identical with it on or off.  tb_spec_stats: speculative passes launched / pairs served / speculations dropped. */
int         tb_set_speculation(int on);
int         tb_spec_stats(uint64_t* launched, uint64_t* served, uint64_t* dropped);
/* Scalar prefetch (csrc/prefetch.cu): when a host-visible scalar has to be fetched from the device, the reductions that
 * followed it last time (dot products into 1-element views, sums of squares: g_x, g_y, |p|, |d| of criteria_conv,
 * solver.rs:599-608) are computed by one extra small kernel and ride on the same round trip; a device write overlapping their
 * inputs drops them.  6 host round trips per solver iteration become 3.  Results agree with the un-prefetched path to rounding
 * (double accumulation).  tb_scalar_prefetch_stats: prefetch kernels launched / requests served / prefetched values dropped. */
int         tb_set_scalar_prefetch(int on);
int         tb_scalar_prefetch_stats(uint64_t* launched, uint64_t* served, uint64_t* dropped);
/* diagnostics: host seconds spent waiting for host-visible scalars (and how many waits) since the last call; a host that
 * never waits is launch-bound, one that mostly waits is device-bound */
int         tb_host_wait_stats(double* seconds, uint64_t* waits);
/* Tracing: every entry point of this header runs inside an NVTX range carrying its own name (nsys / ncu
 * timelines; TB_NVTX=0 turns the ranges off).  tb_set_api_trace(1) additionally collects host-side call counts and wall
 * seconds per entry point; tb_api_trace_dump writes "name calls seconds\n" lines into buf (needed = bytes required). */
int         tb_set_api_trace(int on);
int         tb_api_trace_dump(char* buf, size_t cap, size_t* needed);
/* Device timeline (diagnostics): after tb_timeline_begin(max_events) every kernel launch of the library records a CUDA event
 * behind it; tb_timeline_dump writes "index file:line device_us host_us\n" per launch (device_us = completion time of that
 * kernel, host_us = when the launch call was made, both relative to the first launch) and stops recording.  Shows one solver
 * iteration kernel by kernel with its real gaps (PDL overlap, host round trips, peer waits) without a serialising profiler. */
int         tb_timeline_begin(size_t max_events);
int         tb_timeline_dump(char* buf, size_t cap, size_t* needed);

/* ---- buffers: the SliceLike role (slicelike.rs:23-69; totsu_f32cuda/src/f32cuda_slice.rs:215-309) ------- */
/* SliceLike::new_ref / new_mut: wrap caller-owned host memory with a device mirror.  The host slice stays
 * the caller's; it is made coherent on tb_host_ref/tb_host_mut and on tb_buf_release (slicelike.rs:18-19). */
int tb_buf_wrap(int dtype, void* host, size_t len, int host_is_mut, tb_handle* out);
/* device-only buffer (no host mirror), zero-filled: used for matrices generated in HBM */
int tb_buf_alloc(int dtype, size_t len, tb_handle* out);
/* SliceLike::drop of a root slice: flush device-newer ranges to the host slice (if mutable) and free */
int tb_buf_release(tb_handle buf);
int tb_buf_len(tb_handle buf, size_t* out);
/* For bindings whose slice type IS the host sub-slice (rust/totsu_b200: `B200Slice` is a transparent wrapper of
 * `[F]`, so SliceLike::split_ref/split_mut (slicelike.rs:37-40) cost no allocation): find the wrapped root that
 * contains host range [host, host + len) and return the view; a 0-length range yields the empty view {0,0,0}, which
 * every entry point accepts.  tb_buf_retain adds `n` live wrappers to a root (one per non-empty child of a split);
 * tb_buf_release removes one and only flushes + frees the root when none is left (slicelike.rs:18-19,42-46). */
int tb_view_of_host(int dtype, const void* host, size_t len, tb_view* out);
int tb_buf_retain(tb_handle buf, int n);
/* SliceLike::get_ref / get_mut: make the host range current (D2H of device-newer sub-ranges);
 * get_mut additionally marks the range host-newer so the next device use re-uploads it */
int tb_host_ref(tb_view v);
int tb_host_mut(tb_view v);
/* SliceLike::get / set (slicelike.rs:54-69) without the two nested splits */
int tb_get1_f32(tb_view v, size_t idx, float* out);
int tb_get1_f64(tb_view v, size_t idx, double* out);
int tb_set1_f32(tb_view v, size_t idx, float val);
int tb_set1_f64(tb_view v, size_t idx, double val);
/* explicit copies between caller memory and the device copy of a view (tests, device-only buffers) */
int tb_upload(tb_view v, const void* src);
int tb_download(tb_view v, void* dst);

/* ---- LinAlg (linalg.rs:22-67; CPU twin totsu_f64lapack/src/f64lapack.rs:20-73) -------------------------- */
int tb_norm_f32(tb_view x, float* out);                                   /* linalg.rs:27  dnrm2  */
int tb_norm_f64(tb_view x, double* out);
int tb_copy_f32(tb_view x, tb_view y);                                    /* linalg.rs:33  dcopy  */
int tb_copy_f64(tb_view x, tb_view y);
int tb_scale_f32(float alpha, tb_view x);                                 /* linalg.rs:39  dscal; alpha==0 is a zero-fill */
int tb_scale_f64(double alpha, tb_view x);
int tb_add_f32(float alpha, tb_view x, tb_view y);                        /* linalg.rs:46  daxpy  */
int tb_add_f64(double alpha, tb_view x, tb_view y);
int tb_adds_f32(float s, tb_view y);                                      /* linalg.rs:52  daxpy incx=0 */
int tb_adds_f64(double s, tb_view y);
int tb_abssum_f32(tb_view x, size_t incx, float* out);                    /* linalg.rs:58  dasum over ceil(len/incx) */
int tb_abssum_f64(tb_view x, size_t incx, double* out);
int tb_transform_di_f32(float alpha, tb_view mat, tb_view x, float beta, tb_view y);   /* linalg.rs:67  dsbmv k=0 */
int tb_transform_di_f64(double alpha, tb_view mat, tb_view x, double beta, tb_view y);

/* ---- LinAlgEx (linalg_ex.rs:23-65; CPU twin f64lapack.rs:120-191) --------------------------------------- */
/* y = alpha*G*x + beta*y  or  alpha*G^T*x + beta*y; G column-major n_row x n_col, lda = n_row (linalg_ex.rs:23) */
int tb_transform_ge_f32(int transpose, size_t n_row, size_t n_col, float alpha, tb_view mat, tb_view x, float beta, tb_view y);
int tb_transform_ge_f64(int transpose, size_t n_row, size_t n_col, double alpha, tb_view mat, tb_view x, double beta, tb_view y);
/* y = alpha*S*x + beta*y; S symmetric, upper triangle packed by columns (linalg_ex.rs:37) */
int tb_transform_sp_f32(size_t n, float alpha, tb_view mat, tb_view x, float beta, tb_view y);
int tb_transform_sp_f64(size_t n, double alpha, tb_view mat, tb_view x, double beta, tb_view y);
/* consumer warps of the streaming transform_sp kernel: 8 (default) or 16 (two warps per column in the column pass; measured slower) */
int tb_set_spmv_warps(int warps);
/* linalg_ex.rs:43 */
size_t tb_map_eig_worklen(size_t n);
/* linalg_ex.rs:64 map_eig, split around the host closure `map: Fn(F)->Option<F>`:
 *   begin : unpack `mat` (optionally scaling the diagonal), eigendecompose on the device, return all n
 *           eigenvalues in `host_eigs` (ascending order is NOT guaranteed);
 *   (host applies the closure to the eigenvalues in (0, +inf] like dsyevr(range=V, vl=0) does, f64lapack.rs:86-107)
 *   finish: mat := sum_i keep[i] ? new_eigs[i] * z_i z_i^T : 0, diagonal unscaled, repacked. */
int tb_map_eig_begin_f32(tb_view mat, int has_scale, float scale_diag, float eps_zero, tb_view work, float* host_eigs);
int tb_map_eig_begin_f64(tb_view mat, int has_scale, double scale_diag, double eps_zero, tb_view work, double* host_eigs);
int tb_map_eig_finish_f32(tb_view mat, int has_scale, float scale_diag, tb_view work, const float* new_eigs, const uint8_t* keep);
int tb_map_eig_finish_f64(tb_view mat, int has_scale, double scale_diag, tb_view work, const double* new_eigs, const uint8_t* keep);
/* ConePSD::proj fast path (cone_psd.rs:56-79): the closure `e > 0 ? Some(e) : None` applied on the device.
 * Default algorithm: GEMM-only matrix-sign iteration, proj = (X + X sign(X))/2, no host round trip;
 * tb_set_psd_path(1) selects the Jacobi eigendecomposition that tb_map_eig_begin/finish use (tests compare both). */
int tb_set_psd_path(int mode);
/* The GEMM every step of the sign iteration is made of: C = alpha*A*B + beta*D + gamma*I with A, B, D symmetric k x k
 * column-major (d.len == 0: no D term), C symmetric (upper triangle computed, mirrored).  engine 2 = tcgen05 3xTF32
 * tensor-core kernel with split-K over a thread-block cluster (splitk 0 = choose, 1, 2, 4 or 8; k % 4 == 0);
 * engine 1 = FP32-pipe kernel.  f32 only (f64 stays on the FP64 pipe).  Replaces the dsyr rank-1 loop of
 * f64lapack.rs:96-105 / the cublasSsyr loop of f32cuda.rs:316-324 inside ConePSD::proj; exported for the parity tests
 * and the tensor-pipe measurement of config C4.  tb_set_psd_path: 0 = sign iteration (tcgen05 for f32), 1 = Jacobi,
 * 2 = sign iteration on the FP32/FP64 pipes, 3 = tcgen05 without split-K. */
int tb_symm_gemm_f32(size_t k, float alpha, tb_view a, tb_view b, float beta, tb_view d, float gamma, tb_view c, int engine, int splitk);
/* Diagnostics: `reps` back-to-back tensor-core GEMMs c = a*b; the last one records %globaltimer stamps (ns) of CTA 0
 * at its phases into stamps_ns[0..8]: entered, prologue done, dependency resolved, first chunk staged, producers
 * done, accumulator complete, partials pushed, cluster barrier passed, results stored. */
int tb_symm_gemm_trace_f32(size_t k, tb_view a, tb_view b, tb_view c, int splitk, int reps, uint64_t* stamps_ns);
/* The two tb_cone_proj calls of one solver iteration (dual cone on y, primal cone on s: solver.rs:548-549) are batched: the
 * first is parked until the second arrives (or anything else is called); the Zero / RPos / SOC / RotSOC blocks of both vectors then
 * run as one launch and, in f32, every GEMM of the PSD sign iteration covers both problems.  tb_set_psd_pairing(0) runs each call
 * at once; tb_psd_pairs counts batched PSD block pairs, tb_cone_pairs the projection pairs served by one launch. */
int tb_set_psd_pairing(int on);
int tb_psd_pairs(uint64_t* out);
int tb_cone_pairs(uint64_t* out);
int tb_proj_psd_f32(tb_view x, float eps_zero, tb_view work);
/* MatBuild::set_sqrt (totsu/src/matbuild/mod.rs:220-241): mat := P^(1/2) for an upper-packed symmetric PSD P, i.e. map_eig with
 * scale_diag = None and the closure e -> sqrt(e) on the positive eigenvalues - GEMM-only (coupled Newton-Schulz on the
 * tcgen05 / FP64 engine of the ConePSD projection), with the Jacobi eigendecomposition as the fallback for k < 32 and for
 * matrices the iteration cannot resolve.  tb_sqrt_psd_info: route of the last call (1 = Newton-Schulz, 2 = eigendecomposition)
 * and its step count. */
int tb_sqrt_psd_f32(tb_view mat, float eps_zero, tb_view work);
int tb_sqrt_psd_f64(tb_view mat, double eps_zero, tb_view work);
int tb_sqrt_psd_info(int* route, int* iterations);
int tb_proj_psd_f64(tb_view x, double eps_zero, tb_view work);

/* ---- fused device-resident Operator (operator.rs:11-156) for one stacked dense A ------------------------ */
/* `mat` is column-major n_row x n_col (lda = n_row): the rows THIS rank owns.  With tb_dist_init'ed world
 * size G > 1 the operator is row-sharded: rows [row_offset, row_offset+n_row) of an n_row_total x n_col A;
 * op() all-gathers the y slices, trans_op() all-reduces the partial sums (north_star). */
int tb_denseop_create(int dtype, tb_view mat, size_t n_row, size_t n_col, size_t row_offset, size_t n_row_total, tb_handle* out);
int tb_denseop_destroy(tb_handle op);
/* Operator::op (transpose=0) / trans_op (transpose=1); x, y are full-length (replicated) vectors */
int tb_denseop_apply_f32(tb_handle op, int transpose, float alpha, tb_view x, float beta, tb_view y);
int tb_denseop_apply_f64(tb_handle op, int transpose, double alpha, tb_view x, double beta, tb_view y);
/* one pass over A for an op / trans_op pair that reads different inputs and writes different outputs:
 *   y_n = alpha_n*A*x_n + beta_n*y_n   and   y_t = alpha_t*A^T*x_t + beta_t*y_t   */
int tb_denseop_apply_pair_f32(tb_handle op, float alpha_n, tb_view x_n, float beta_n, tb_view y_n,
                              float alpha_t, tb_view x_t, float beta_t, tb_view y_t);
int tb_denseop_apply_pair_f64(tb_handle op, double alpha_n, tb_view x_n, double beta_n, tb_view y_n,
                              double alpha_t, tb_view x_t, double beta_t, tb_view y_t);
/* Operator::absadd_cols (tau[c] += sum_r |A[r,c]|) and absadd_rows (sigma[r] += sum_c |A[r,c]|) */
int tb_denseop_absadd_cols_f32(tb_handle op, tb_view tau);
int tb_denseop_absadd_cols_f64(tb_handle op, tb_view tau);
int tb_denseop_absadd_rows_f32(tb_handle op, tb_view sigma);
int tb_denseop_absadd_rows_f64(tb_handle op, tb_view sigma);

/* ---- batched product Cone (cone.rs:9-30; cone_{zero,rpos,soc,rotsoc}.rs) -------------------------------- */
enum { TB_CONE_ZERO = 0, TB_CONE_RPOS = 1, TB_CONE_SOC = 2, TB_CONE_ROTSOC = 3, TB_CONE_PSD = 4 };
typedef struct tb_cone_block { int32_t type; int32_t reserved; uint64_t len; } tb_cone_block;
/* blocks are laid out back to back over the m-vector, like ProbSOCPCone (socp.rs:296-313) */
int tb_cone_create(const tb_cone_block* blocks, size_t n_blocks, tb_handle* out);
int tb_cone_destroy(tb_handle cone);
/* Cone::proj for the whole product cone in one launch (+ one eigensolve per PSD block, using psd_work) */
int tb_cone_proj_f32(tb_handle cone, int dual_cone, tb_view x, float eps_zero, tb_view psd_work);
int tb_cone_proj_f64(tb_handle cone, int dual_cone, tb_view x, double eps_zero, tb_view psd_work);
/* Cone::product_group with the solver's `group` closure (min over each block of size>1; solver.rs:509-523) */
int tb_cone_group_min_f32(tb_handle cone, tb_view dp_tau);
int tb_cone_group_min_f64(tb_handle cone, tb_view dp_tau);

/* ---- vector helpers used by the solver's host loops (solver.rs:501-506, 551-552, 566-567) --------------- */
/* x[i] = 1 / max(x[i], eps)  — calc_precond's two host loops over get_mut() */
int tb_recip_clamp_f32(float eps, tb_view x);
int tb_recip_clamp_f64(double eps, tb_view x);

/* ---- synthetic instances (bench / tests): counter-based generator keyed (seed, global row, col) --------- */
/* mat (column-major n_row x n_col, lda = n_row) := scale * (2*u(seed,row_offset+r,c) - 1), u in [0,1) with 24 bits */
int tb_fill_uniform_f32(tb_view mat, size_t n_row, size_t n_col, size_t row_offset, uint64_t seed, float scale);
int tb_fill_uniform_f64(tb_view mat, size_t n_row, size_t n_col, size_t row_offset, uint64_t seed, double scale);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink 5 / NVSwitch -------------------------------------- */
#define TB_NCCL_ID_BYTES 128
int tb_dist_unique_id(void* id_out /* TB_NCCL_ID_BYTES */);   /* rank 0, then broadcast by the launcher */
int tb_dist_init(int rank, int world, const void* id);
int tb_dist_finalize(void);
int tb_dist_info(int* rank, int* world);
/* 1 when the sharded operator's all-gather / all-reduce run as peer stores fused into the matvec epilogue
 * (cudaIpc-mapped staging over NVLink; TB_P2P=0 forces the NCCL baseline), 0 when they go through NCCL */
int tb_dist_p2p_enabled(int* out);
/* peer exchanges (one push + one wait kernel each) since tb_dist_init: a solver iteration needs 2 - the gather of A x and the
 * reduce of A^T y of an op/trans_op pair travel together, and the speculated criteria_conv pair rides with the pair before it */
int tb_dist_exchanges(uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* TOTSU_B200_H */
