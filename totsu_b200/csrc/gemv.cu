// Dense matrix-vector products: LinAlgEx::transform_ge (totsu_core/src/linalg_ex.rs:23; CPU twin
// totsu_f64lapack/src/f64lapack.rs:123-146 = cblas dgemv ColumnMajor lda=n_row; superseded GPU call
// totsu_f32cuda/src/f32cuda.rs:144-171 = cublasSgemv_v2) and the fused device-resident dense Operator
// (operator.rs:11-156) that the throughput configurations hand to the solver.
//
// The matvec is HBM-bound: one read of A per call, 2*m*n*sizeof(F) bytes per op/trans_op pair.
//
// Fast path (`stream_kernel`): persistent, warp-specialised kernel, one CTA per SM.
//   * a producer warp streams column segments of A (TR rows x TC columns per stage, 32 KB) into a
//     6-stage shared-memory ring with TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP),
//     each lane issuing one 4 KB column segment; full/empty mbarriers per stage;
//   * 8 consumer warps read the staged tile with 128-bit LDS:
//       pass N (y = A x):   thread t owns rows 4t..4t+3 of the row chunk and keeps their partial sums in
//                           registers across the whole column range of the work unit;
//       pass T (y = A^T x): warp w owns column w of the tile, lanes stride down the rows against the x
//                           chunk held in registers, one shuffle reduction per column;
//     the two passes can run on the same staged tile, which is what tb_denseop_apply_pair uses to serve an
//     op and a trans_op with a single read of A;
//   * partial results go to a scratch buffer and a small finalize kernel adds them in a fixed order
//     (no atomics: results are bit-reproducible), applying alpha/beta.
// Generic path (`gemv_n_generic` / `gemv_t_generic`): plain coalesced LDG kernels for matrices whose
// leading dimension / base address is not 16-byte aligned (e.g. the 63 x n blocks of ProbSOCP) or that are
// too small to be worth a persistent launch.  Same two-stage deterministic reduction.
#include "common.cuh"
#include "vprog.cuh"
#include "ptx.cuh"

namespace tb {

// -------------------------------------------------------------------------------------------------------
// finalize: y[i] = alpha * sum_j part[j*ld + i] + beta*y[i]      (beta == 0: y is not read)
// -------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void finalize_kernel(const T* __restrict__ part, int nparts, size_t ld, size_t len, T alpha, T beta, T* y) {
    tbd::pdl_entry();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) {
        T s = T(0);
        for (int j = 0; j < nparts; ++j) s += part[(size_t)j * ld + i];
        T r = alpha * s;
        if (beta != T(0)) r += beta * y[i];
        y[i] = r;
    }
}

template <typename T> void l2_finalize(const T* part, int nparts, size_t ld, size_t len, T alpha, T beta, T* y) {
    if (len == 0) return;
    if (vp_enabled_wide(len)) { vp_finalize(DT<T>::id, part, nparts, ld, len, (double)alpha, (double)beta, y); return; }
    int g = (int)std::min<size_t>((len + 255) / 256, (size_t)ctx().sm_count * 8);
    launch_pdl(finalize_kernel<T>, dim3(g), dim3(256), 0, ctx().stream, part, nparts, ld, len, alpha, beta, y);
    TB_LAUNCH_CHECK();
}
template void l2_finalize<float>(const float*, int, size_t, size_t, float, float, float*);
template void l2_finalize<double>(const double*, int, size_t, size_t, double, double, double*);
template <typename T> static void finalize(const T* part, int nparts, size_t ld, size_t len, T alpha, T beta, T* y) {
    l2_finalize<T>(part, nparts, ld, len, alpha, beta, y);
}

// both outputs of a paired pass in one launch: indices [0, len_a) finalize y_a, [len_a, len_a + len_b) finalize y_b
template <typename T>
__global__ void finalize2_kernel(const T* __restrict__ part_a, int nparts_a, size_t ld_a, size_t len_a, T alpha_a, T beta_a, T* y_a,
                                 const T* __restrict__ part_b, int nparts_b, size_t ld_b, size_t len_b, T alpha_b, T beta_b, T* y_b) {
    tbd::pdl_entry();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < len_a + len_b; i += (size_t)gridDim.x * blockDim.x) {
        const bool first = i < len_a;
        const T* part = first ? part_a : part_b;
        const int nparts = first ? nparts_a : nparts_b;
        const size_t ld = first ? ld_a : ld_b;
        const size_t k = first ? i : i - len_a;
        T s = T(0);
        for (int j = 0; j < nparts; ++j) s += part[(size_t)j * ld + k];
        const T alpha = first ? alpha_a : alpha_b, beta = first ? beta_a : beta_b;
        T* y = first ? y_a : y_b;
        T r = alpha * s;
        if (beta != T(0)) r += beta * y[k];
        y[k] = r;
    }
}
template <typename T>
static void finalize2(const T* part_a, int nparts_a, size_t ld_a, size_t len_a, T alpha_a, T beta_a, T* y_a,
                      const T* part_b, int nparts_b, size_t ld_b, size_t len_b, T alpha_b, T beta_b, T* y_b) {
    const size_t len = len_a + len_b;
    if (len == 0) return;
    if (vp_enabled_wide(std::max(len_a, len_b))) {     // two independent micro-ops: no barrier between them
        if (len_a) vp_finalize(DT<T>::id, part_a, nparts_a, ld_a, len_a, (double)alpha_a, (double)beta_a, y_a);
        if (len_b) vp_finalize(DT<T>::id, part_b, nparts_b, ld_b, len_b, (double)alpha_b, (double)beta_b, y_b);
        return;
    }
    int g = (int)std::min<size_t>((len + 255) / 256, (size_t)ctx().sm_count * 8);
    launch_pdl(finalize2_kernel<T>, dim3(g), dim3(256), 0, ctx().stream, part_a, nparts_a, ld_a, len_a, alpha_a, beta_a, y_a, part_b, nparts_b, ld_b, len_b, alpha_b, beta_b, y_b);
    TB_LAUNCH_CHECK();
}

// -------------------------------------------------------------------------------------------------------
// generic kernels
// -------------------------------------------------------------------------------------------------------
// y-part[j][r] = sum_{c in split j} f(A[r,c]) * x[c];  thread per row.  ABS: f = |.| and x == 1 (absadd_rows).
template <typename T, bool ABS>
__global__ void gemv_n_generic(const T* __restrict__ A, size_t lda, size_t n_row, size_t n_col,
                               const T* __restrict__ x, T* __restrict__ out, size_t ld_out, size_t cols_per_split,
                               bool direct, T alpha, T beta) {
    tbd::pdl_entry();
    size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t c0 = (size_t)blockIdx.y * cols_per_split;
    size_t c1 = c0 + cols_per_split < n_col ? c0 + cols_per_split : n_col;
    if (r >= n_row) return;
    const T* a = A + r;
    T acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    size_t c = c0;
    for (; c + 4 <= c1; c += 4) {
        T a0 = a[(c + 0) * lda], a1 = a[(c + 1) * lda], a2 = a[(c + 2) * lda], a3 = a[(c + 3) * lda];
        if (ABS) {
            acc0 += fabs(a0); acc1 += fabs(a1); acc2 += fabs(a2); acc3 += fabs(a3);
        } else {
            acc0 += a0 * x[c + 0]; acc1 += a1 * x[c + 1]; acc2 += a2 * x[c + 2]; acc3 += a3 * x[c + 3];
        }
    }
    for (; c < c1; ++c) {
        T a0 = a[c * lda];
        acc0 += ABS ? fabs(a0) : a0 * x[c];
    }
    T s = (acc0 + acc1) + (acc2 + acc3);
    if (direct) {
        T rr = alpha * s;
        if (beta != T(0)) rr += beta * out[r];
        out[r] = rr;
    } else {
        out[(size_t)blockIdx.y * ld_out + r] = s;
    }
}

// out-part[j][c] = sum_{r in split j} f(A[r,c]) * x[r];  a CTA of 256 threads handles 8 columns.
template <typename T, bool ABS>
__global__ void gemv_t_generic(const T* __restrict__ A, size_t lda, size_t n_row, size_t n_col,
                               const T* __restrict__ x, T* __restrict__ out, size_t ld_out, size_t rows_per_split,
                               bool direct, T alpha, T beta, T* __restrict__ y_final, unsigned int* tickets) {
    constexpr int CW = 8;
    __shared__ T red[CW][8];
    __shared__ bool last;
    tbd::pdl_entry();
    size_t cb = (size_t)blockIdx.x * CW;
    size_t r0 = (size_t)blockIdx.y * rows_per_split;
    size_t r1 = r0 + rows_per_split < n_row ? r0 + rows_per_split : n_row;
    int ncol = (int)(n_col - cb < (size_t)CW ? n_col - cb : (size_t)CW);
    T acc[CW];
#pragma unroll
    for (int k = 0; k < CW; ++k) acc[k] = 0;
    const T* a = A + cb * lda;
    if (ncol == CW) {
        for (size_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
            T xv = ABS ? T(1) : x[r];
            T v[CW];
#pragma unroll
            for (int k = 0; k < CW; ++k) v[k] = a[(size_t)k * lda + r];
#pragma unroll
            for (int k = 0; k < CW; ++k) acc[k] += (ABS ? fabs(v[k]) : v[k]) * xv;
        }
    } else {
        for (size_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
            T xv = ABS ? T(1) : x[r];
            for (int k = 0; k < ncol; ++k) {
                T v = a[(size_t)k * lda + r];
                acc[k] += (ABS ? fabs(v) : v) * xv;
            }
        }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < CW; ++k) {
        T s = tbd::warp_sum(acc[k]);
        if (lane == 0) red[k][w] = s;
    }
    __syncthreads();
    if (threadIdx.x < ncol) {
        int k = threadIdx.x;
        T s = T(0);
        int nw = blockDim.x >> 5;
        for (int i = 0; i < nw; ++i) s += red[k][i];
        size_t c = cb + k;
        if (direct) {
            T rr = alpha * s;
            if (beta != T(0)) rr += beta * out[c];
            out[c] = rr;
        } else {
            out[(size_t)blockIdx.y * ld_out + c] = s;
        }
    }
    if (!direct && y_final != nullptr) {
        // fused finalize: the last row-split block of this column group adds the partials in split order
        // (deterministic) and applies alpha/beta
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            unsigned int t = atomicInc(&tickets[blockIdx.x], gridDim.y - 1);      // wraps back to 0
            last = (t == gridDim.y - 1);
        }
        __syncthreads();
        if (last && threadIdx.x < ncol) {
            __threadfence();
            const size_t c = cb + threadIdx.x;
            T s = T(0);
            for (unsigned int j = 0; j < gridDim.y; ++j) s += __ldcg(&out[(size_t)j * ld_out + c]);
            T rr = alpha * s;
            if (beta != T(0)) rr += beta * y_final[c];
            y_final[c] = rr;
        }
    }
}

// y = alpha * f(A) x + beta y  (N) on the generic path
template <typename T, bool ABS>
static void run_generic_n(const T* A, size_t lda, size_t n_row, size_t n_col, const T* x, T alpha, T beta, T* y) {
    Context& c = ctx();
    if (n_row == 0) return;
    if (!ABS && n_col == 1 && vp_enabled_wide(n_row)) {       // an n x 1 operator applied to a device scalar: y = alpha * A * x[0] + beta * y
        vp_axs(DT<T>::id, (double)alpha, A, x, (double)beta, y, n_row);
        return;
    }
    size_t row_blocks = (n_row + 127) / 128;
    size_t want = ((size_t)c.sm_count * 2 + row_blocks - 1) / row_blocks;
    size_t max_splits = std::max<size_t>(1, n_col / 64);
    size_t splits = std::max<size_t>(1, std::min(want, max_splits));
    splits = std::min<size_t>(splits, 65535);
    size_t cps = (n_col + splits - 1) / splits;
    splits = n_col == 0 ? 1 : (n_col + cps - 1) / cps;
    dim3 grid((unsigned)row_blocks, (unsigned)splits);
    if (splits == 1) {
        launch_pdl(gemv_n_generic<T, ABS>, grid, dim3(128), 0, c.stream, A, lda, n_row, n_col, x, y, (size_t)0, std::max<size_t>(cps, 1), true, alpha, beta);
        TB_LAUNCH_CHECK();
    } else {
        T* part = reinterpret_cast<T*>(scratch(splits * n_row * sizeof(T)));
        launch_pdl(gemv_n_generic<T, ABS>, grid, dim3(128), 0, c.stream, A, lda, n_row, n_col, x, part, n_row, cps, false, alpha, beta);
        TB_LAUNCH_CHECK();
        finalize<T>(part, (int)splits, n_row, n_row, alpha, beta, y);
    }
}

template <typename T, bool ABS>
static void run_generic_t(const T* A, size_t lda, size_t n_row, size_t n_col, const T* x, T alpha, T beta, T* y) {
    Context& c = ctx();
    if (n_col == 0) return;
    if (!ABS && n_col == 1 && n_row > 0 && vp_enabled_red(n_row)) {     // the transposed n x 1 operator: a dot product into y[0]
        vp_dot(DT<T>::id, (double)alpha, A, x, n_row, (double)beta, y);
        return;
    }
    size_t col_groups = (n_col + 7) / 8;
    size_t want = ((size_t)c.sm_count * 2 + col_groups - 1) / col_groups;
    size_t max_splits = std::max<size_t>(1, n_row / 1024);
    size_t splits = std::max<size_t>(1, std::min(want, max_splits));
    splits = std::min<size_t>(splits, 65535);
    size_t rps = (n_row + splits - 1) / splits;
    splits = n_row == 0 ? 1 : (n_row + rps - 1) / rps;
    TB_REQUIRE(col_groups <= 2147483647u, "too many columns");
    dim3 grid((unsigned)col_groups, (unsigned)splits);
    if (splits == 1) {
        launch_pdl(gemv_t_generic<T, ABS>, grid, dim3(256), 0, c.stream, A, lda, n_row, n_col, x, y, (size_t)0, std::max<size_t>(rps, 1), true, alpha, beta, (T*)nullptr, (unsigned int*)nullptr);
        TB_LAUNCH_CHECK();
    } else {
        T* part = reinterpret_cast<T*>(scratch(splits * n_col * sizeof(T)));
        const bool fused = col_groups <= (size_t)Context::kTicketPool;      // one ticket per column group
        launch_pdl(gemv_t_generic<T, ABS>, grid, dim3(256), 0, c.stream, A, lda, n_row, n_col, x, part, n_col, rps, false, alpha, beta,
                   fused ? y : (T*)nullptr, c.tickets + 64);
        TB_LAUNCH_CHECK();
        if (!fused) finalize<T>(part, (int)splits, n_col, n_col, alpha, beta, y);
    }
}

// -------------------------------------------------------------------------------------------------------
// TMA streaming kernel
// -------------------------------------------------------------------------------------------------------


template <typename T> struct StreamCfg;
template <> struct StreamCfg<float> {
    static constexpr int VEC = 4;
    static constexpr int STAGES = 6;
};
template <> struct StreamCfg<double> {
    static constexpr int VEC = 2;
    static constexpr int STAGES = 6;
};

constexpr int kConsumerWarps = 8;
constexpr int kConsumers = kConsumerWarps * 32;   // 256
constexpr int kTC = 8;                            // columns per tile == consumer warps
constexpr int kStreamThreads = kConsumers + 32;   // + producer warp
constexpr int kMaxUnitCols = 2048;                // x slice of a work unit staged in shared memory

template <typename T> struct alignas(16) Vec16;
template <> struct alignas(16) Vec16<float> { float v[4]; };
template <> struct alignas(16) Vec16<double> { double v[2]; };

struct StreamParams {
    const void* A;
    size_t lda, n_row, n_col;
    const void* x_n[2];  // length n_col  (N passes: y = A x), null when unused
    const void* x_t[2];  // length n_row  (T passes: y = A^T x)
    void* part_n[2];     // [n_splits][n_row] each
    void* part_t[2];     // [n_chunks][n_col] each
    int n_chunks;        // row chunks of TR rows
    int n_splits;        // column splits per row chunk
    long long n_tiles;   // ceil(n_col / kTC): split s covers column tiles [n_tiles*s/n_splits, n_tiles*(s+1)/n_splits)
    int n_units;         // n_chunks * n_splits
};

// NN / NT: how many N passes (y = A x) and T passes (y = A^T x) are served by this one read of A (0..2 each).  <1,1> is
// an op/trans_op pair; <2,2> adds the speculated products of the NEXT pair (see "speculative pairing" below).
// ABS: the products are taken with |A| and x == 1 (no x is read): one pass yields the row sums AND the column sums of |A|,
// i.e. Operator::absadd_rows and absadd_cols (matop.rs:98-138) from a single read of A.
template <typename T, int NN, int NT, bool ABS = false>
__global__ void __launch_bounds__(kStreamThreads, 1) stream_kernel(const StreamParams p) {
    using Cfg = StreamCfg<T>;
    constexpr int VEC = Cfg::VEC;
    constexpr int TR = kConsumers * VEC;              // rows per tile (1024 f32 / 512 f64): 4 KB per column segment
    constexpr int STAGES = Cfg::STAGES;
    constexpr int STAGE_ELEMS = TR * kTC;             // 32 KB
    constexpr int KROW = TR / (32 * VEC);             // 8 row groups per lane in pass T
    constexpr int NNX = NN > 0 ? NN : 1, NTX = NT > 0 ? NT : 1;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* stage_base = reinterpret_cast<T*>(smem_raw);
    T* xs = stage_base + (size_t)STAGES * STAGE_ELEMS;                                   // [2][kMaxUnitCols]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(xs + 2 * kMaxUnitCols);
    uint64_t* empty_bar = full_bar + STAGES;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const T* __restrict__ A = reinterpret_cast<const T*>(p.A);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], kConsumerWarps);
        }
        ptx::mbar_fence_init();
    }
    tbd::pdl_entry();          // barrier set-up above overlaps the tail of the previous kernel; no global access before this point
    __syncthreads();

    // contiguous range of work units for this CTA; unit u = chunk * n_splits + split
    const int u_begin = (int)(((long long)p.n_units * blockIdx.x) / gridDim.x);
    const int u_end = (int)(((long long)p.n_units * (blockIdx.x + 1)) / gridDim.x);

    if (warp == kConsumerWarps) {
        // ===================== producer warp =====================
        const uint64_t pol = ptx::policy_evict_first();
        int s = 0;
        uint32_t phase = 0;
        for (int u = u_begin; u < u_end; ++u) {
            const int chunk = u / p.n_splits, split = u - chunk * p.n_splits;
            const size_t row0 = (size_t)chunk * TR;
            const uint32_t rows = (uint32_t)(p.n_row - row0 < (size_t)TR ? p.n_row - row0 : (size_t)TR);
            const size_t c0 = (size_t)kTC * (size_t)((p.n_tiles * split) / p.n_splits);
            const size_t c1e = (size_t)kTC * (size_t)((p.n_tiles * (split + 1)) / p.n_splits);
            const size_t c1 = c1e < p.n_col ? c1e : p.n_col;
            for (size_t c = c0; c < c1; c += kTC) {
                const int ncols = (int)(c1 - c < (size_t)kTC ? c1 - c : (size_t)kTC);
                ptx::mbar_wait(&empty_bar[s], phase ^ 1);
                if (lane == 0) ptx::mbar_expect_tx(&full_bar[s], rows * (uint32_t)sizeof(T) * (uint32_t)ncols);
                __syncwarp();
                if (lane < ncols)
                    ptx::bulk_g2s(stage_base + (size_t)s * STAGE_ELEMS + (size_t)lane * TR, A + (c + lane) * p.lda + row0,
                                  rows * (uint32_t)sizeof(T), &full_bar[s], pol);
                if (++s == STAGES) { s = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================== consumer warps =====================
        int s = 0;
        uint32_t phase = 0;
        for (int u = u_begin; u < u_end; ++u) {
            const int chunk = u / p.n_splits, split = u - chunk * p.n_splits;
            const size_t row0 = (size_t)chunk * TR;
            const int rows = (int)(p.n_row - row0 < (size_t)TR ? p.n_row - row0 : (size_t)TR);
            const size_t c0 = (size_t)kTC * (size_t)((p.n_tiles * split) / p.n_splits);
            const size_t c1e = (size_t)kTC * (size_t)((p.n_tiles * (split + 1)) / p.n_splits);
            const size_t c1 = c1e < p.n_col ? c1e : p.n_col;
            const int ucols = (int)(c1 - c0);

            if (NN > 0 && !ABS) {
                // stage this unit's slice of x (previous unit's readers are done: barrier first)
                ptx::named_bar_sync(1, kConsumers);
#pragma unroll
                for (int q = 0; q < NNX; ++q) {
                    const T* __restrict__ x_n = reinterpret_cast<const T*>(p.x_n[q]);
                    for (int i = tid; i < ucols; i += kConsumers) xs[q * kMaxUnitCols + i] = x_n[c0 + i];
                }
                ptx::named_bar_sync(1, kConsumers);
            }
            // pass T: x chunk for the rows this lane strides over, rows beyond the matrix read as 0
            T xr[NTX][KROW * VEC];
            if (NT > 0 && !ABS) {
#pragma unroll
                for (int q = 0; q < NTX; ++q) {
                    const T* __restrict__ x_t = reinterpret_cast<const T*>(p.x_t[q]);
#pragma unroll
                    for (int k = 0; k < KROW; ++k)
#pragma unroll
                        for (int i = 0; i < VEC; ++i) {
                            int r = (lane + 32 * k) * VEC + i;
                            xr[q][k * VEC + i] = r < rows ? x_t[row0 + r] : T(0);
                        }
                }
            }
            T acc[NNX][VEC];
#pragma unroll
            for (int q = 0; q < NNX; ++q)
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[q][i] = T(0);
            const bool row_ok = tid * VEC < rows;       // rows is a multiple of VEC

            int cl = 0;   // column offset inside the unit
            for (size_t c = c0; c < c1; c += kTC, cl += kTC) {
                const int ncols = (int)(c1 - c < (size_t)kTC ? c1 - c : (size_t)kTC);
                ptx::mbar_wait(&full_bar[s], phase);
                const T* tile = stage_base + (size_t)s * STAGE_ELEMS;
                if (NN > 0 && row_ok) {
                    if (ncols == kTC) {
                        Vec16<T> v[kTC];
#pragma unroll
                        for (int j = 0; j < kTC; ++j) v[j] = *reinterpret_cast<const Vec16<T>*>(tile + (size_t)j * TR + tid * VEC);
#pragma unroll
                        for (int q = 0; q < NNX; ++q)
#pragma unroll
                            for (int j = 0; j < kTC; ++j) {
                                if (ABS) {
#pragma unroll
                                    for (int i = 0; i < VEC; ++i) acc[q][i] += fabs(v[j].v[i]);
                                } else {
                                    T xv = xs[q * kMaxUnitCols + cl + j];
#pragma unroll
                                    for (int i = 0; i < VEC; ++i) acc[q][i] += v[j].v[i] * xv;
                                }
                            }
                    } else {
                        for (int j = 0; j < ncols; ++j) {
                            Vec16<T> v = *reinterpret_cast<const Vec16<T>*>(tile + (size_t)j * TR + tid * VEC);
#pragma unroll
                            for (int q = 0; q < NNX; ++q) {
                                T xv = ABS ? T(1) : xs[q * kMaxUnitCols + cl + j];
#pragma unroll
                                for (int i = 0; i < VEC; ++i) acc[q][i] += (ABS ? fabs(v.v[i]) : v.v[i]) * xv;
                            }
                        }
                    }
                }
                if (NT > 0) {
                    if (warp < ncols) {
                        const T* col = tile + (size_t)warp * TR;
                        T sum0[NTX], sum1[NTX];
#pragma unroll
                        for (int q = 0; q < NTX; ++q) { sum0[q] = T(0); sum1[q] = T(0); }
                        if (rows == TR) {
                            // full chunk (all but the last row chunk): no per-row predicates, all loads issued up front
                            Vec16<T> v[KROW];
#pragma unroll
                            for (int k = 0; k < KROW; ++k) v[k] = *reinterpret_cast<const Vec16<T>*>(col + (size_t)(lane + 32 * k) * VEC);
#pragma unroll
                            for (int k = 0; k < KROW; ++k)
#pragma unroll
                                for (int q = 0; q < NTX; ++q)
#pragma unroll
                                    for (int i = 0; i < VEC; ++i) {
                                        const T t = ABS ? fabs(v[k].v[i]) : v[k].v[i] * xr[q][k * VEC + i];
                                        if ((k & 1) == 0) sum0[q] += t;
                                        else sum1[q] += t;
                                    }
                        } else {
#pragma unroll
                            for (int k = 0; k < KROW; ++k) {
                                if ((lane + 32 * k) * VEC < rows) {
                                    Vec16<T> v = *reinterpret_cast<const Vec16<T>*>(col + (size_t)(lane + 32 * k) * VEC);
#pragma unroll
                                    for (int q = 0; q < NTX; ++q)
#pragma unroll
                                        for (int i = 0; i < VEC; ++i) {
                                            const T t = ABS ? fabs(v.v[i]) : v.v[i] * xr[q][k * VEC + i];
                                            if ((k & 1) == 0) sum0[q] += t;
                                            else sum1[q] += t;
                                        }
                                }
                            }
                        }
#pragma unroll
                        for (int q = 0; q < NTX; ++q) {
                            T sum = tbd::warp_sum(sum0[q] + sum1[q]);
                            if (lane == 0) reinterpret_cast<T*>(p.part_t[q])[(size_t)chunk * p.n_col + c + warp] = sum;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&empty_bar[s]);
                if (++s == STAGES) { s = 0; phase ^= 1; }
            }
            if (NN > 0 && row_ok) {
#pragma unroll
                for (int q = 0; q < NNX; ++q) {
                    T* dst = reinterpret_cast<T*>(p.part_n[q]) + (size_t)split * p.n_row + row0 + (size_t)tid * VEC;
#pragma unroll
                    for (int i = 0; i < VEC; ++i) dst[i] = acc[q][i];
                }
            }
        }
    }
}

template <typename T> static size_t stream_smem_bytes() {
    using Cfg = StreamCfg<T>;
    size_t stage = (size_t)kConsumers * Cfg::VEC * kTC * sizeof(T);
    return stage * Cfg::STAGES + 2 * (size_t)kMaxUnitCols * sizeof(T) + 2 * Cfg::STAGES * sizeof(uint64_t) + 128;
}

template <typename T> static bool stream_eligible(const T* A, size_t lda, size_t n_row, size_t n_col) {
    constexpr size_t VEC = 16 / sizeof(T);
    if ((reinterpret_cast<uintptr_t>(A) & 15) != 0) return false;
    if (lda % VEC != 0 || n_row % VEC != 0) return false;
    if (n_row < 256 || n_col < 16) return false;
    int mode = ctx().gemv_mode;
    if (mode == 1) return false;
    if (mode == 2) return true;
    return n_row * n_col >= (size_t(1) << 20);
}

template <typename T, int NN, int NT, bool ABS = false> static void launch_stream(const StreamParams& p, int grid, size_t smem) {
    static bool attr_set = false;     // one flag per kernel instantiation
    if (!attr_set) {
        TB_CUDA(cudaFuncSetAttribute(stream_kernel<T, NN, NT, ABS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    Context& c = ctx();
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c.prof_on) {
        auto get_ev = [&]() {
            cudaEvent_t e;
            if (!c.prof_pool.empty()) { e = c.prof_pool.back(); c.prof_pool.pop_back(); }
            else TB_CUDA(cudaEventCreate(&e));
            return e;
        };
        e0 = get_ev(); e1 = get_ev();
        TB_CUDA(cudaEventRecord(e0, c.stream));
    }
    launch_pdl(stream_kernel<T, NN, NT, ABS>, dim3(grid), dim3(kStreamThreads), smem, c.stream, p);
    TB_LAUNCH_CHECK();
    if (c.prof_on) {
        TB_CUDA(cudaEventRecord(e1, c.stream));
        c.prof_recs.push_back({e0, e1, NN * 3 + NT, (double)p.n_row * (double)p.n_col * sizeof(T)});     // one read of A
    }
}

static size_t gcd_sz(size_t a, size_t b) { while (b) { size_t t = a % b; a = b; b = t; } return a; }

// Runs pass N and/or pass T over one read of A.  Outputs: y_n (len n_row), y_t (len n_col).
// `sharded`: A is this rank's row shard; y_n is then the BASE of the full-length vector (the local slice starts at
// rank*n_row) and the epilogues carry the collective: all-gather of the slices / all-reduce of the partial sums.
// Products computed ahead of their call in the same read of A (see "speculative pairing" further down): raw partials of
// A x_n and A^T x_t, finalized later by stream_finalize_partials when the call they belong to arrives.
template <typename T> struct SpecJob {
    const T* x_n;
    const T* x_t;
    T* part_n;            // out: [n_splits][n_row], allocated by run_stream
    T* part_t;            // out: [n_chunks][n_col]
    size_t n_splits, n_chunks;
    // sharded: when the speculated pair's raw products travelled in the carrying pair's exchange (dist.cu PairSeg) they are
    // already gathered / reduced over the ranks, un-scaled: raw_n [world * n_row], raw_t [n_col]
    size_t n_row_total;   // in: rows of the whole operator
    T* raw_n;
    T* raw_t;
    bool exchanged;
};
char* spec_buffer(size_t bytes);

template <typename T>
static void stream_finalize_partials(const T* part_n, size_t n_splits, const T* part_t, size_t n_chunks, size_t n_row, size_t n_col,
                                     T alpha_n, T beta_n, T* y_n, T alpha_t, T beta_t, T* y_t, bool sharded) {
    const bool do_n = y_n != nullptr, do_t = y_t != nullptr;
    if (sharded) {
        if (do_n && do_t) {
            dist_finalize_pair<T>(part_n, (int)n_splits, n_row, n_row, alpha_n, beta_n, y_n, part_t, (int)n_chunks, n_col, n_col, alpha_t, beta_t, y_t);
            return;
        }
        if (do_n) dist_finalize_gather<T>(part_n, (int)n_splits, n_row, n_row, alpha_n, beta_n, y_n);
        if (do_t) dist_finalize_reduce<T>(part_t, (int)n_chunks, n_col, n_col, alpha_t, beta_t, y_t);
        return;
    }
    if (do_n && do_t) {
        finalize2<T>(part_n, (int)n_splits, n_row, n_row, alpha_n, beta_n, y_n, part_t, (int)n_chunks, n_col, n_col, alpha_t, beta_t, y_t);
        return;
    }
    if (do_n) finalize<T>(part_n, (int)n_splits, n_row, n_row, alpha_n, beta_n, y_n);
    if (do_t) finalize<T>(part_t, (int)n_chunks, n_col, n_col, alpha_t, beta_t, y_t);
}

template <typename T>
static void run_stream(const T* A, size_t lda, size_t n_row, size_t n_col,
                       const T* x_n, T alpha_n, T beta_n, T* y_n,
                       const T* x_t, T alpha_t, T beta_t, T* y_t, bool sharded = false, SpecJob<T>* spec = nullptr, bool abs_ones = false) {
    Context& c = ctx();
    constexpr int VEC = StreamCfg<T>::VEC;
    constexpr size_t TR = (size_t)kConsumers * VEC;
    const bool do_n = y_n != nullptr, do_t = y_t != nullptr;
    const int n_cta = c.sm_count;
    const size_t n_chunks = (n_row + TR - 1) / TR;
    // column splits per row chunk: make n_chunks*n_splits a multiple of the CTA count (equal unit sizes ->
    // balanced persistent CTAs), keep every unit's x slice within the shared-memory window
    size_t base = (size_t)n_cta / gcd_sz(n_chunks, (size_t)n_cta);
    size_t min_splits = (n_col + kMaxUnitCols - 1) / kMaxUnitCols;
    size_t max_splits = std::max<size_t>(1, n_col / 32);
    size_t n_splits = base;
    while (n_splits < min_splits) n_splits += base;
    // prefer >= 4 units per CTA when the matrix is wide enough
    while (n_splits + base <= max_splits && n_chunks * n_splits < (size_t)4 * n_cta) n_splits += base;
    if (n_splits > max_splits) n_splits = std::max(min_splits, max_splits);
    const size_t n_tiles = (n_col + kTC - 1) / kTC;
    n_splits = std::max<size_t>(1, std::min(n_splits, n_tiles));
    TB_REQUIRE(((n_tiles + n_splits - 1) / n_splits) * kTC <= (size_t)kMaxUnitCols, "internal: unit too wide");

    size_t bytes_n = do_n ? n_splits * n_row * sizeof(T) : 0;
    size_t bytes_t = do_t ? n_chunks * n_col * sizeof(T) : 0;
    bytes_n = (bytes_n + 255) & ~size_t(255);
    char* sc = reinterpret_cast<char*>(scratch(bytes_n + bytes_t + 256));
    StreamParams p;
    p.A = A; p.lda = lda; p.n_row = n_row; p.n_col = n_col;
    p.x_n[0] = x_n; p.x_t[0] = x_t; p.x_n[1] = nullptr; p.x_t[1] = nullptr;
    p.part_n[0] = sc; p.part_t[0] = sc + bytes_n; p.part_n[1] = nullptr; p.part_t[1] = nullptr;
    p.n_chunks = (int)n_chunks; p.n_splits = (int)n_splits; p.n_tiles = (long long)n_tiles;
    p.n_units = (int)(n_chunks * n_splits);
    int grid = std::min(n_cta, p.n_units);
    size_t smem = stream_smem_bytes<T>();
    if (spec != nullptr && do_n && do_t) {
        // second N pass and second T pass on the same staged tiles: their partials outlive this call in the spec buffer
        size_t sb_n = n_splits * n_row * sizeof(T);
        sb_n = (sb_n + 255) & ~size_t(255);
        size_t sb_t = n_chunks * n_col * sizeof(T);
        sb_t = (sb_t + 255) & ~size_t(255);
        size_t sb_rn = sharded ? spec->n_row_total * sizeof(T) : 0;
        sb_rn = (sb_rn + 255) & ~size_t(255);
        char* sb = spec_buffer(sb_n + sb_t + sb_rn + (sharded ? n_col * sizeof(T) : 0));
        spec->part_n = reinterpret_cast<T*>(sb);
        spec->part_t = reinterpret_cast<T*>(sb + sb_n);
        spec->raw_n = reinterpret_cast<T*>(sb + sb_n + sb_t);
        spec->raw_t = reinterpret_cast<T*>(sb + sb_n + sb_t + sb_rn);
        spec->exchanged = false;
        spec->n_splits = n_splits; spec->n_chunks = n_chunks;
        p.x_n[1] = spec->x_n; p.x_t[1] = spec->x_t;
        p.part_n[1] = spec->part_n; p.part_t[1] = spec->part_t;
        launch_stream<T, 2, 2>(p, grid, smem);
        if (sharded && dist_finalize_pair_spec<T>(reinterpret_cast<const T*>(p.part_n[0]), (int)n_splits, n_row, n_row, alpha_n, beta_n, y_n,
                                                  reinterpret_cast<const T*>(p.part_t[0]), (int)n_chunks, n_col, n_col, alpha_t, beta_t, y_t,
                                                  spec->part_n, spec->part_t, spec->raw_n, spec->raw_t)) {
            spec->exchanged = true;       // one exchange carried both pairs: the speculated one will be served by a local axpby
            return;
        }
    } else if (abs_ones) {
        TB_REQUIRE(do_n && do_t, "internal: the |A| pass produces both sums");
        launch_stream<T, 1, 1, true>(p, grid, smem);
    } else if (do_n && do_t) launch_stream<T, 1, 1>(p, grid, smem);
    else if (do_n) launch_stream<T, 1, 0>(p, grid, smem);
    else launch_stream<T, 0, 1>(p, grid, smem);
    stream_finalize_partials<T>(reinterpret_cast<const T*>(p.part_n[0]), n_splits, reinterpret_cast<const T*>(p.part_t[0]), n_chunks, n_row, n_col,
                                alpha_n, beta_n, y_n, alpha_t, beta_t, y_t, sharded);
}

// y = alpha*op(A)*x + beta*y on device pointers (single GPU, no collectives)
template <typename T>
void gemv_dev(bool transpose, size_t n_row, size_t n_col, size_t lda, T alpha, const T* A, const T* x, T beta, T* y) {
    if (stream_eligible<T>(A, lda, n_row, n_col)) {
        if (!transpose) run_stream<T>(A, lda, n_row, n_col, x, alpha, beta, y, nullptr, T(0), T(0), nullptr);
        else run_stream<T>(A, lda, n_row, n_col, nullptr, T(0), T(0), nullptr, x, alpha, beta, y);
    } else {
        if (!transpose) run_generic_n<T, false>(A, lda, n_row, n_col, x, alpha, beta, y);
        else run_generic_t<T, false>(A, lda, n_row, n_col, x, alpha, beta, y);
    }
}

template <typename T>
void gemv_pair_dev(size_t n_row, size_t n_col, size_t lda, const T* A,
                   T alpha_n, const T* x_n, T beta_n, T* y_n, T alpha_t, const T* x_t, T beta_t, T* y_t, SpecJob<T>* spec = nullptr) {
    if (stream_eligible<T>(A, lda, n_row, n_col)) {
        run_stream<T>(A, lda, n_row, n_col, x_n, alpha_n, beta_n, y_n, x_t, alpha_t, beta_t, y_t, false, spec);
    } else {
        run_generic_n<T, false>(A, lda, n_row, n_col, x_n, alpha_n, beta_n, y_n);
        run_generic_t<T, false>(A, lda, n_row, n_col, x_t, alpha_t, beta_t, y_t);
    }
}

// -------------------------------------------------------------------------------------------------------
// LinAlgEx::transform_ge
// -------------------------------------------------------------------------------------------------------
template <typename T>
static void api_transform_ge(int transpose, size_t n_row, size_t n_col, T alpha, tb_view mat, tb_view x, T beta, tb_view y) {
    require_init();
    TB_REQUIRE(mat.len == n_row * n_col, "transform_ge: mat.len != n_row*n_col");       // f64lapack.rs:125
    if (transpose) {
        TB_REQUIRE(x.len == n_row && y.len == n_col, "transform_ge(T): vector length mismatch");   // :128-129
    } else {
        TB_REQUIRE(x.len == n_col && y.len == n_row, "transform_ge(N): vector length mismatch");   // :133-134
    }
    if (transpose && n_col == 1 && n_row > 0) {      // MatOp n x 1 transposed (ProbLPOpC / OpB etc.): a dot product into a 1-element view
        double v = 0.0;
        if (pf_try_dot(DT<T>::id, mat, x, y, (double)alpha, (double)beta, &v)) {
            set_scalar<T>(y, alpha * (T)v);
            return;
        }
    }
    if (!transpose && n_col == 1 && n_row > 0 && vp_enabled_wide(n_row)) {      // MatOp n x 1 applied to a host-current 1-element slice: see denseop_apply
        double sv = 0.0;
        if (host_scalar_if_current(x, DT<T>::id, &sv)) {
            const T* A0 = rptr<T>(mat);
            T* py0 = wptr<T>(y, beta == T(0));
            vp_axs_imm(DT<T>::id, (double)alpha, A0, sv, (double)beta, py0, n_row);
            return;
        }
    }
    const T* A = rptr<T>(mat);
    const T* px = rptr<T>(x);
    T* py = wptr<T>(y, beta == T(0));
    if (y.len == 0) return;
    if (n_row == 0 || n_col == 0) {          // dgemv quick return: y = beta*y
        l1_scale<T>(beta, py, y.len);
        return;
    }
    gemv_dev<T>(transpose != 0, n_row, n_col, n_row, alpha, A, px, beta, py);
}

// -------------------------------------------------------------------------------------------------------
// Fused dense Operator (row-sharded across ranks when tb_dist_init'ed)
// -------------------------------------------------------------------------------------------------------
struct DenseOp {
    int dtype;
    tb_view mat;
    size_t n_row, n_col;        // local shard
    size_t row_offset, n_row_total;
    void* tmp_n;                // n_col elements: local partial of A^T x before the all-reduce (sharded only)
    // one-pass absadd: row sums and column sums of |A| (this rank's rows) from ONE streaming read, kept until A changes
    void* abs_rows = nullptr;   // n_row elements
    void* abs_cols = nullptr;   // n_col elements
    bool abs_valid = false;
};

static DenseOp& get_op(tb_handle h) {
    Context& c = ctx();
    if (h <= 0 || (size_t)h > c.denseops.size() || c.denseops[(size_t)h - 1] == nullptr) fail(TB_ERR_ARG, "invalid denseop handle");
    return *c.denseops[(size_t)h - 1];
}

template <typename T> static void denseop_apply(tb_handle h, int transpose, T alpha, tb_view x, T beta, tb_view y) {
    require_init();
    DenseOp& op = get_op(h);
    Context& c = ctx();
    TB_REQUIRE(op.dtype == DT<T>::id, "denseop dtype mismatch");
    const size_t m = op.n_row_total, n = op.n_col;
    if (transpose) TB_REQUIRE(x.len == m && y.len == n, "denseop trans_op: vector length mismatch");
    else TB_REQUIRE(x.len == n && y.len == m, "denseop op: vector length mismatch");
    if (transpose && n == 1 && op.n_row == m) {      // an m x 1 operator transposed = a dot product into a 1-element view (c^T x, b^T y)
        double v = 0.0;
        if (pf_try_dot(DT<T>::id, op.mat, x, y, (double)alpha, (double)beta, &v)) {
            set_scalar<T>(y, alpha * (T)v);          // served from the prefetched product: same arithmetic as the combine step it replaces
            return;
        }
    }
    if (!transpose && n == 1 && op.n_row == m && m > 0 && vp_enabled_wide(m)) {
        // an m x 1 operator applied to a 1-element slice the host has just written (b and c on `work_one`, solver.rs:590-596): the
        // scalar travels by value - no upload of the element, no device read of it, same arithmetic as the device-scalar form
        double sv = 0.0;
        if (host_scalar_if_current(x, DT<T>::id, &sv)) {
            const T* A0 = rptr<T>(op.mat);
            T* py0 = wptr<T>(y, beta == T(0));
            vp_axs_imm(DT<T>::id, (double)alpha, A0, sv, (double)beta, py0, m);
            return;
        }
    }
    const T* A = rptr<T>(op.mat);
    const T* px = rptr<T>(x);
    T* py = wptr<T>(y, beta == T(0));
    const bool sharded = c.world > 1 && op.n_row != op.n_row_total;
    if (!sharded) {
        gemv_dev<T>(transpose != 0, op.n_row, n, op.n_row, alpha, A, px, beta, py);
        return;
    }
    const T* xl = px + op.row_offset;      // trans_op reads only the rows this rank owns
    if (stream_eligible<T>(A, op.n_row, op.n_row, n)) {
        // one streaming pass over the shard; the epilogue stores straight into the peers (dist.cu)
        if (!transpose) run_stream<T>(A, op.n_row, op.n_row, n, px, alpha, beta, py, nullptr, T(0), T(0), nullptr, true);
        else run_stream<T>(A, op.n_row, op.n_row, n, nullptr, T(0), T(0), nullptr, xl, alpha, beta, py, true);
        return;
    }
    if (!transpose) {
        // local slice of y, then all-gather the slices (equal shard sizes, rank-major)
        gemv_dev<T>(false, op.n_row, n, op.n_row, alpha, A, px, beta, py + op.row_offset);
        dist_allgather_inplace(py, op.n_row, DT<T>::id);
    } else {
        // partial A_loc^T x_loc -> all-reduce -> y = alpha*sum + beta*y
        T* tmp = reinterpret_cast<T*>(op.tmp_n);
        gemv_dev<T>(true, op.n_row, n, op.n_row, T(1), A, xl, T(0), tmp);
        dist_allreduce_sum(tmp, n, DT<T>::id);
        l1_axpby<T>(alpha, tmp, beta, py, n);
    }
}

// ---- speculative pairing ------------------------------------------------------------------------------------
// The solver's iteration is a chain of three op/trans_op pairs on A (SelfDualEmbed::trans_op, SelfDualEmbed::op,
// criteria_conv: solver.rs:146-153, 122-131, 594-598).  The inputs of the third pair (x_x, x_y: parts of x_hat) are already
// final when the second runs, so its two products can be computed from the SAME staged tiles: stream_kernel<T, 2, 2>
// runs two N passes and two T passes per read of A and parks the raw partials of the extra pair in a side buffer.  When
// that pair then arrives through the trait surface - alpha, beta and the output views only now known - it is served by the
// finalize step alone.  An iteration reads A twice instead of three times; results are bit-identical (same kernel code,
// same decomposition, same summation order).
//   * prediction: a first-order table "pair with inputs (x_n, x_t) was followed by the pair with inputs (x_n', x_t')",
//     learnt from the call stream itself - nothing about the solver is hard-wired;
//   * validity: every device write goes through dev_ptr(write) (and every host->device refresh through the same
//     function), which reports the written range to spec_note_write; a write that overlaps the speculated inputs drops
//     the speculation and marks that prediction as not speculable (the x_hat-dependent second pair is learnt that way
//     after one wasted attempt - wasted flops, never wasted bytes);
//   * tb_set_speculation(0) turns it off; tb_spec_stats reports served / dropped speculations.
struct SpecSig { tb_view xn, xt; };
static inline bool view_eq(const tb_view& a, const tb_view& b) { return a.buf == b.buf && a.off == b.off && a.len == b.len; }
static inline bool sig_eq(const SpecSig& a, const SpecSig& b) { return view_eq(a.xn, b.xn) && view_eq(a.xt, b.xt); }

struct SpecState {
    bool valid = false;
    tb_handle op = 0;
    int dtype = 0;
    SpecSig sig{};
    void* part_n = nullptr;
    void* part_t = nullptr;
    size_t n_splits = 0, n_chunks = 0;
    void* raw_n = nullptr;      // sharded + exchanged: gathered / reduced raw products (see SpecJob)
    void* raw_t = nullptr;
    bool exchanged = false;
    char* buf = nullptr;
    size_t buf_bytes = 0;
    bool have_last = false;
    tb_handle last_op = 0;
    SpecSig last{};
    std::vector<std::pair<SpecSig, SpecSig>> next_of;
    std::vector<SpecSig> bad;
};
static SpecState g_spec;

char* spec_buffer(size_t bytes) {
    SpecState& S = g_spec;
    if (bytes > S.buf_bytes) {
        TB_CUDA(cudaStreamSynchronize(ctx().stream));
        if (S.buf) TB_CUDA(cudaFree(S.buf));
        TB_CUDA(cudaMalloc(&S.buf, bytes));
        S.buf_bytes = bytes;
    }
    return S.buf;
}

static inline bool view_overlaps(const tb_view& v, tb_handle buf, size_t off, size_t len) {
    return v.buf == buf && v.len > 0 && len > 0 && v.off < off + len && off < v.off + v.len;
}

// called by dev_ptr for every range about to change on the device
void spec_note_write(tb_handle buf, size_t off, size_t len) {
    pf_note_write(buf, off, len);
    for (DenseOp* op : ctx().denseops)          // cached |A| sums die with any write into their matrix
        if (op != nullptr && op->abs_valid && view_overlaps(op->mat, buf, off, len)) op->abs_valid = false;
    SpecState& S = g_spec;
    if (!S.valid) return;
    if (view_overlaps(S.sig.xn, buf, off, len) || view_overlaps(S.sig.xt, buf, off, len)) {
        S.valid = false;
        S.bad.push_back(S.sig);
        ctx().spec_dropped += 1;
    }
}
// a buffer went away: forget everything that names it (handles are recycled)
void spec_note_release(tb_handle buf) {
    pf_note_release(buf);
    SpecState& S = g_spec;
    auto names = [&](const SpecSig& g) { return g.xn.buf == buf || g.xt.buf == buf; };
    if (S.valid && names(S.sig)) S.valid = false;
    if (S.have_last && names(S.last)) S.have_last = false;
    for (size_t i = 0; i < S.next_of.size();) {
        if (names(S.next_of[i].first) || names(S.next_of[i].second)) S.next_of.erase(S.next_of.begin() + (long)i);
        else ++i;
    }
    for (size_t i = 0; i < S.bad.size();) {
        if (names(S.bad[i])) S.bad.erase(S.bad.begin() + (long)i);
        else ++i;
    }
}
void spec_reset() {
    SpecState& S = g_spec;
    S.valid = false; S.have_last = false;
    S.next_of.clear(); S.bad.clear();
}

template <typename T>
static void denseop_apply_pair(tb_handle h, T alpha_n, tb_view x_n, T beta_n, tb_view y_n, T alpha_t, tb_view x_t, T beta_t, tb_view y_t) {
    require_init();
    DenseOp& op = get_op(h);
    Context& c = ctx();
    TB_REQUIRE(op.dtype == DT<T>::id, "denseop dtype mismatch");
    const size_t m = op.n_row_total, n = op.n_col;
    TB_REQUIRE(x_n.len == n && y_n.len == m && x_t.len == m && y_t.len == n, "denseop pair: vector length mismatch");
    const bool sharded = c.world > 1 && op.n_row != op.n_row_total;
    const T* A = rptr<T>(op.mat);
    const bool streams = stream_eligible<T>(A, op.n_row, op.n_row, n);
    SpecState& S = g_spec;
    const SpecSig cur{x_n, x_t};

    // ---- served from a speculation made during the previous pass over A?
    if (S.valid && S.op == h && S.dtype == DT<T>::id && sig_eq(S.sig, cur) && streams) {
        (void)rptr<T>(x_n); (void)rptr<T>(x_t);         // same coherence side effects as a real apply (no-ops when current)
        T* syn = wptr<T>(y_n, beta_n == T(0));
        T* syt = wptr<T>(y_t, beta_t == T(0));
        if (S.valid) {                                   // the output views may overlap the inputs: wptr above would have dropped it
            if (S.exchanged) {
                // the raw products crossed the ranks with the carrying pair: only alpha / beta are left to apply, locally
                l1_axpby<T>(alpha_n, reinterpret_cast<const T*>(S.raw_n), beta_n, syn, m);
                l1_axpby<T>(alpha_t, reinterpret_cast<const T*>(S.raw_t), beta_t, syt, n);
            } else {
                stream_finalize_partials<T>(reinterpret_cast<const T*>(S.part_n), S.n_splits, reinterpret_cast<const T*>(S.part_t), S.n_chunks,
                                            op.n_row, n, alpha_n, beta_n, syn, alpha_t, beta_t, syt, sharded);
            }
            S.valid = false;
            c.spec_served += 1;
            if (S.have_last && S.last_op == h) {
                bool known = false;
                for (auto& e : S.next_of) if (sig_eq(e.first, S.last)) { e.second = cur; known = true; }
                if (!known) S.next_of.push_back({S.last, cur});
            }
            S.last = cur; S.last_op = h; S.have_last = true;
            return;
        }
    }
    if (S.valid) { S.valid = false; c.spec_dropped += 1; }      // a different pair came: the parked products are stale

    // ---- learn the transition, then look the next pair up
    if (S.have_last && S.last_op == h) {
        bool known = false;
        for (auto& e : S.next_of) if (sig_eq(e.first, S.last)) { e.second = cur; known = true; }
        if (!known) {
            if (S.next_of.size() >= 16) S.next_of.erase(S.next_of.begin());
            S.next_of.push_back({S.last, cur});
        }
    }
    S.last = cur; S.last_op = h; S.have_last = true;
    const SpecSig* next = nullptr;
    if (c.speculation && streams) {
        for (auto& e : S.next_of) if (sig_eq(e.first, cur)) next = &e.second;
        if (next != nullptr) {
            for (auto& b : S.bad) if (sig_eq(b, *next)) { next = nullptr; break; }
        }
        if (next != nullptr && (sig_eq(*next, cur))) next = nullptr;
    }

    const T* pxn = rptr<T>(x_n);
    const T* pxt = rptr<T>(x_t);
    SpecJob<T> job{};
    SpecSig nsig{};
    if (next != nullptr) {
        nsig = *next;                                    // copy: the table may be edited below
        job.x_n = rptr<T>(nsig.xn);
        job.x_t = rptr<T>(nsig.xt);
        job.n_row_total = m;
        if (sharded) job.x_t += op.row_offset;
    }
    T* pyn = wptr<T>(y_n, beta_n == T(0));
    T* pyt = wptr<T>(y_t, beta_t == T(0));
    // outputs of THIS pair that overlap the speculated inputs would be read before they are written: do not speculate then
    bool spec_ok = next != nullptr;
    if (spec_ok) {
        const tb_view outs[2] = {y_n, y_t};
        for (const tb_view& o : outs)
            if (view_overlaps(nsig.xn, o.buf, o.off, o.len) || view_overlaps(nsig.xt, o.buf, o.off, o.len)) spec_ok = false;
    }
    if (sharded) {
        if (streams) {
            run_stream<T>(A, op.n_row, op.n_row, n, pxn, alpha_n, beta_n, pyn, pxt + op.row_offset, alpha_t, beta_t, pyt, true, spec_ok ? &job : nullptr);
        } else {
            denseop_apply<T>(h, 0, alpha_n, x_n, beta_n, y_n);
            denseop_apply<T>(h, 1, alpha_t, x_t, beta_t, y_t);
            return;
        }
    } else {
        gemv_pair_dev<T>(op.n_row, n, op.n_row, A, alpha_n, pxn, beta_n, pyn, alpha_t, pxt, beta_t, pyt, (spec_ok && streams) ? &job : nullptr);
    }
    if (spec_ok && streams) {
        S.valid = true; S.op = h; S.dtype = DT<T>::id; S.sig = nsig;
        S.part_n = job.part_n; S.part_t = job.part_t; S.n_splits = job.n_splits; S.n_chunks = job.n_chunks;
        S.raw_n = job.raw_n; S.raw_t = job.raw_t; S.exchanged = job.exchanged;
        c.spec_launched += 1;
    }
}

// tb_denseop_apply: park the call, or pair it with the parked one (see "lazy op/trans_op pairing", common.cuh)
template <typename T> static void denseop_submit(tb_handle h, int transpose, T alpha, tb_view x, T beta, tb_view y) {
    require_init();
    Context& c = ctx();
    DenseOp& op = get_op(h);
    TB_REQUIRE(op.dtype == DT<T>::id, "denseop dtype mismatch");
    const size_t m = op.n_row_total, n = op.n_col;
    if (transpose) TB_REQUIRE(x.len == m && y.len == n, "denseop trans_op: vector length mismatch");
    else TB_REQUIRE(x.len == n && y.len == m, "denseop op: vector length mismatch");
    if (!c.pair_fusion) {
        queue_drain();
        denseop_apply<T>(h, transpose, alpha, x, beta, y);
        return;
    }
    // only operators the streaming kernel serves are worth parking; the n x 1 / m x 1 operators that carry c and b
    // on the fused route are ordinary deferrable commands (they sit BETWEEN the two halves of a pair)
    const bool candidate = op.n_row >= 256 && op.n_col >= 16 && op.n_row * op.n_col >= (size_t(1) << 20);
    Cmd cmd;
    cmd.is_dense = true; cmd.op = h; cmd.trans = transpose ? 1 : 0; cmd.dtype = DT<T>::id;
    cmd.alpha = (double)alpha; cmd.beta = (double)beta; cmd.x = x; cmd.y = y;
    cmd.reads[cmd.n_reads++] = x;
    cmd.reads[cmd.n_reads++] = op.mat;
    if (beta != T(0)) cmd.reads[cmd.n_reads++] = y;
    cmd.writes[cmd.n_writes++] = y;
    cmd.run = [=] { denseop_apply<T>(h, transpose, alpha, x, beta, y); };
    if (!candidate) {
        if (c.queue.empty()) { cmd.run(); return; }
        cmd.is_dense = false;
        c.queue.push_back(std::move(cmd));
        if (c.queue.size() > 16) queue_drain();
        return;
    }
    if (!c.queue.empty()) {
        const Cmd& head = c.queue[0];
        bool fuse = head.is_dense && head.op == h && head.trans != cmd.trans && head.dtype == cmd.dtype && !cmds_conflict(head, cmd);
        if (fuse) {
            bool hoist_ok = true, sink_ok = true;
            for (size_t i = 1; i < c.queue.size(); ++i) {
                if (cmds_conflict(c.queue[i], cmd)) hoist_ok = false;
                if (cmds_conflict(c.queue[i], head)) sink_ok = false;
            }
            if (hoist_ok || sink_ok) {
                std::vector<Cmd> q;
                q.swap(c.queue);
                const Cmd& nn = q[0].trans ? cmd : q[0];      // the A*x half
                const Cmd& tt = q[0].trans ? q[0] : cmd;      // the A^T*y half
                auto pair = [&] {
                    denseop_apply_pair<T>(h, (T)nn.alpha, nn.x, (T)nn.beta, nn.y, (T)tt.alpha, tt.x, (T)tt.beta, tt.y);
                    c.pairs_fused += 1;
                };
                if (hoist_ok) pair();
                for (size_t i = 1; i < q.size(); ++i) q[i].run();
                if (!hoist_ok) pair();
                return;
            }
        }
        queue_drain();
    }
    c.queue.push_back(std::move(cmd));
}

// Operator::absadd_cols (tau[c] += sum_r |A[r,c]|) and absadd_rows (sigma[r] += sum_c |A[r,c]|), the two halves of
// SelfDualEmbed::abssum (solver.rs:159-183; per-column / per-row L::abssum loops in MatOp::absadd_impl, matop.rs:98-138).
// For a matrix the streaming kernel serves, whichever is asked for first runs stream_kernel<T,1,1,ABS> - ONE read of A that
// yields both the row sums and the column sums of |A| - and the other is served from the kept sums (C3: 0.66 ms for the
// pair instead of 2.5 + 0.7 ms through the generic kernels).  Any write into the matrix drops the kept sums.
template <typename T> static void denseop_absadd(tb_handle h, bool cols, tb_view v) {
    require_init();
    DenseOp& op = get_op(h);
    Context& c = ctx();
    TB_REQUIRE(op.dtype == DT<T>::id, "denseop dtype mismatch");
    const T* A = rptr<T>(op.mat);
    TB_REQUIRE(v.len == (cols ? op.n_col : op.n_row_total), cols ? "absadd_cols: length mismatch" : "absadd_rows: length mismatch");
    T* pv = wptr<T>(v);
    const bool sharded = c.world > 1 && op.n_row != op.n_row_total;
    if (stream_eligible<T>(A, op.n_row, op.n_row, op.n_col)) {
        if (!op.abs_rows) TB_CUDA(cudaMalloc(&op.abs_rows, op.n_row * sizeof(T)));
        if (!op.abs_cols) TB_CUDA(cudaMalloc(&op.abs_cols, op.n_col * sizeof(T)));
        T* ar = reinterpret_cast<T*>(op.abs_rows);
        T* ac = reinterpret_cast<T*>(op.abs_cols);
        if (!op.abs_valid) {
            run_stream<T>(A, op.n_row, op.n_row, op.n_col, nullptr, T(1), T(0), ar, nullptr, T(1), T(0), ac, false, nullptr, true);
            op.abs_valid = true;
        }
        if (cols) {
            if (!sharded) {
                l1_axpby<T>(T(1), ac, T(1), pv, op.n_col);
            } else {
                T* tmp = reinterpret_cast<T*>(op.tmp_n);
                l1_copy<T>(ac, tmp, op.n_col);
                dist_allreduce_sum(tmp, op.n_col, DT<T>::id);
                l1_axpby<T>(T(1), tmp, T(1), pv, op.n_col);
            }
        } else {
            if (!sharded) {
                l1_axpby<T>(T(1), ar, T(1), pv, op.n_row);
            } else {
                l1_axpby<T>(T(1), ar, T(1), pv + op.row_offset, op.n_row);
                dist_allgather_inplace(pv, op.n_row, DT<T>::id);
            }
        }
        return;
    }
    if (cols) {
        if (!sharded) {
            run_generic_t<T, true>(A, op.n_row, op.n_row, op.n_col, nullptr, T(1), T(1), pv);
        } else {
            T* tmp = reinterpret_cast<T*>(op.tmp_n);
            run_generic_t<T, true>(A, op.n_row, op.n_row, op.n_col, nullptr, T(1), T(0), tmp);
            dist_allreduce_sum(tmp, op.n_col, DT<T>::id);
            l1_axpby<T>(T(1), tmp, T(1), pv, op.n_col);
        }
    } else {
        if (!sharded) {
            run_generic_n<T, true>(A, op.n_row, op.n_row, op.n_col, nullptr, T(1), T(1), pv);
        } else {
            run_generic_n<T, true>(A, op.n_row, op.n_row, op.n_col, nullptr, T(1), T(1), pv + op.row_offset);
            dist_allgather_inplace(pv, op.n_row, DT<T>::id);
        }
    }
}

}  // namespace tb

using namespace tb;

extern "C" {

int tb_prof_enable(int on) {
    return api([&] {
        require_init();
        ctx().prof_on = on != 0;
    });
}
// reads and clears the records: per-variant sums into l9 / ms9 / b9 (index NN * 3 + NT), any of which may be null
static void prof_collect(uint64_t* l9, double* ms9, double* b9) {
    require_init();
    Context& c = ctx();
    TB_CUDA(cudaStreamSynchronize(c.stream));
    for (int i = 0; i < 9; ++i) {
        if (l9) l9[i] = 0;
        if (ms9) ms9[i] = 0.0;
        if (b9) b9[i] = 0.0;
    }
    for (auto& pr : c.prof_recs) {
        float ms = 0.f;
        TB_CUDA(cudaEventElapsedTime(&ms, pr.e0, pr.e1));
        if (l9) l9[pr.variant] += 1;
        if (ms9) ms9[pr.variant] += ms;
        if (b9) b9[pr.variant] += pr.bytes;
        c.prof_pool.push_back(pr.e0);
        c.prof_pool.push_back(pr.e1);
    }
    c.prof_recs.clear();
}
int tb_prof_read(uint64_t* launches, double* total_ms, double* total_bytes) {
    return api([&] {
        uint64_t l9[9]; double ms9[9], b9[9];
        prof_collect(l9, ms9, b9);
        *launches = 0; *total_ms = 0.0; *total_bytes = 0.0;
        for (int i = 0; i < 9; ++i) { *launches += l9[i]; *total_ms += ms9[i]; *total_bytes += b9[i]; }
    });
}
int tb_prof_read_variants(uint64_t* launches9, double* ms9, double* bytes9) {
    return api([&] { prof_collect(launches9, ms9, bytes9); });
}

int tb_transform_ge_f32(int tr, size_t nr, size_t nc, float a, tb_view m, tb_view x, float b, tb_view y) {
    return api_defer({m, x, y}, {y}, [=] { api_transform_ge<float>(tr, nr, nc, a, m, x, b, y); });
}
int tb_transform_ge_f64(int tr, size_t nr, size_t nc, double a, tb_view m, tb_view x, double b, tb_view y) {
    return api_defer({m, x, y}, {y}, [=] { api_transform_ge<double>(tr, nr, nc, a, m, x, b, y); });
}

int tb_denseop_create(int dtype, tb_view mat, size_t n_row, size_t n_col, size_t row_offset, size_t n_row_total, tb_handle* out) {
    return api([&] {
        require_init();
        TB_REQUIRE(dtype == TB_F32 || dtype == TB_F64, "bad dtype");
        TB_REQUIRE(mat.len == n_row * n_col, "denseop: mat.len != n_row*n_col");
        if (n_row_total == 0) n_row_total = n_row;
        TB_REQUIRE(row_offset + n_row <= n_row_total, "denseop: shard out of range");
        Context& c = ctx();
        if (c.world > 1 && n_row != n_row_total)
            TB_REQUIRE(n_row * (size_t)c.world == n_row_total && row_offset == n_row * (size_t)c.rank,
                       "denseop: row shards must be equal-sized and rank-ordered");
        else       // not sharded: the operator IS the whole matrix (a partial one would silently leave rows of y untouched)
            TB_REQUIRE(n_row == n_row_total && row_offset == 0, "denseop: a row shard needs tb_dist_init with world > 1");
        (void)dev_ptr(mat, dtype, false);
        void* tmp = nullptr;
        if (c.world > 1 && n_row != n_row_total) TB_CUDA(cudaMalloc(&tmp, std::max<size_t>(n_col, 1) * (dtype == TB_F32 ? 4 : 8)));
        DenseOp* op = new DenseOp();
        op->dtype = dtype; op->mat = mat; op->n_row = n_row; op->n_col = n_col; op->row_offset = row_offset; op->n_row_total = n_row_total; op->tmp_n = tmp;
        c.denseops.push_back(op);
        *out = (tb_handle)c.denseops.size();
    });
}
int tb_denseop_destroy(tb_handle h) {
    return api([&] {
        DenseOp& op = get_op(h);
        spec_reset();
        if (op.tmp_n || op.abs_rows || op.abs_cols) {
            cudaStreamSynchronize(ctx().stream);
            if (op.tmp_n) cudaFree(op.tmp_n);
            if (op.abs_rows) cudaFree(op.abs_rows);
            if (op.abs_cols) cudaFree(op.abs_cols);
        }
        delete &op;
        ctx().denseops[(size_t)h - 1] = nullptr;
    });
}
int tb_denseop_apply_f32(tb_handle op, int tr, float a, tb_view x, float b, tb_view y) { return api_raw([&] { denseop_submit<float>(op, tr, a, x, b, y); }); }
int tb_denseop_apply_f64(tb_handle op, int tr, double a, tb_view x, double b, tb_view y) { return api_raw([&] { denseop_submit<double>(op, tr, a, x, b, y); }); }
int tb_denseop_apply_pair_f32(tb_handle op, float an, tb_view xn, float bn, tb_view yn, float at, tb_view xt, float bt, tb_view yt) {
    return api([&] { denseop_apply_pair<float>(op, an, xn, bn, yn, at, xt, bt, yt); });
}
int tb_denseop_apply_pair_f64(tb_handle op, double an, tb_view xn, double bn, tb_view yn, double at, tb_view xt, double bt, tb_view yt) {
    return api([&] { denseop_apply_pair<double>(op, an, xn, bn, yn, at, xt, bt, yt); });
}
int tb_denseop_absadd_cols_f32(tb_handle op, tb_view tau) { return api([&] { denseop_absadd<float>(op, true, tau); }); }
int tb_denseop_absadd_cols_f64(tb_handle op, tb_view tau) { return api([&] { denseop_absadd<double>(op, true, tau); }); }
int tb_denseop_absadd_rows_f32(tb_handle op, tb_view s) { return api([&] { denseop_absadd<float>(op, false, s); }); }
int tb_denseop_absadd_rows_f64(tb_handle op, tb_view s) { return api([&] { denseop_absadd<double>(op, false, s); }); }
}
