// LinAlgEx::transform_sp (totsu_core/src/linalg_ex.rs:37; CPU twin totsu_f64lapack/src/f64lapack.rs:149-163 =
// cblas dspmv Upper; superseded GPU call totsu_f32cuda/src/f32cuda.rs:174-187 = cublasSspmv_v2):
//     y = alpha*S*x + beta*y,  S symmetric n x n, upper triangle packed by columns: S[r,c] (r<=c) at c(c+1)/2 + r
//     (index convention: totsu_core/src/floatgeneric.rs:193-201).
//
// One pass over the packed array.  The triangle is cut into column panels of PANEL columns; a CTA owns one
// panel x one row range.  For every packed element a = S[r,c] (r < c) it accumulates both contributions
//     y[r] += a*x[c]   (thread-owned row r, register accumulator across the panel's columns)
//     y[c] += a*x[r]   (per-thread per-column accumulators, block-reduced once per CTA)
// plus the diagonal.  Per-CTA partial vectors go to scratch and a finalize kernel adds them in fixed order.
#include "common.cuh"

namespace tb {

constexpr int SP_PANEL = 16;     // columns per CTA
constexpr int SP_THREADS = 256;
constexpr int SP_ROWS = 2048;    // rows per CTA (8 per thread)

// part_r[panel][r]  : row contributions of panel `panel` (only r < panel_end are written/read)
// part_c[chunk][c]  : column contributions of row chunk `chunk`
template <typename T>
__global__ void spmv_kernel(const T* __restrict__ S, size_t n, const T* __restrict__ x, T* __restrict__ part_r, T* __restrict__ part_c) {
    __shared__ T red[SP_PANEL][SP_THREADS / 32];
    const size_t panel = blockIdx.x, chunk = blockIdx.y;
    const size_t c0 = panel * SP_PANEL;
    const size_t c1 = c0 + SP_PANEL < n ? c0 + SP_PANEL : n;
    const size_t r0 = chunk * SP_ROWS;
    if (r0 >= c1) return;                 // strictly below the diagonal block: nothing stored (uniform per CTA)
    const size_t r1 = r0 + SP_ROWS < c1 ? r0 + SP_ROWS : c1;
    const int ncol = (int)(c1 - c0);

    T xc[SP_PANEL];
    size_t base[SP_PANEL];
#pragma unroll
    for (int k = 0; k < SP_PANEL; ++k) {
        size_t c = c0 + k;
        xc[k] = k < ncol ? x[c] : T(0);
        base[k] = c * (c + 1) / 2;
    }
    T cacc[SP_PANEL];
#pragma unroll
    for (int k = 0; k < SP_PANEL; ++k) cacc[k] = T(0);

    for (size_t r = r0 + threadIdx.x; r < r1; r += SP_THREADS) {
        const T xr = x[r];
        T racc = T(0);
#pragma unroll
        for (int k = 0; k < SP_PANEL; ++k) {
            const size_t c = c0 + k;
            if (k < ncol && r <= c) {
                const T a = S[base[k] + r];
                if (r < c) {
                    racc += a * xc[k];        // y[r] += S[r,c] x[c]
                    cacc[k] += a * xr;        // y[c] += S[r,c] x[r]
                } else {
                    cacc[k] += a * xr;        // diagonal, counted once
                }
            }
        }
        part_r[panel * n + r] = racc;
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < SP_PANEL; ++k) {
        T s = tbd::warp_sum(cacc[k]);
        if (lane == 0) red[k][w] = s;
    }
    __syncthreads();
    if ((int)threadIdx.x < ncol) {
        T s = T(0);
        for (int i = 0; i < SP_THREADS / 32; ++i) s += red[threadIdx.x][i];
        part_c[chunk * n + c0 + threadIdx.x] = s;
    }
}

// y[i] = alpha * ( sum_{panels p with panel_end > i} part_r[p][i]  +  sum_{chunks q with q*ROWS < panel_end(i)} part_c[q][i] ) + beta*y[i]
template <typename T>
__global__ void spmv_finalize(const T* __restrict__ part_r, const T* __restrict__ part_c, size_t n, T alpha, T beta, T* y) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t n_panels = (n + SP_PANEL - 1) / SP_PANEL;
    T s = T(0);
    // row contributions: panels whose columns reach past row i, i.e. panel >= i / PANEL
    for (size_t p = i / SP_PANEL; p < n_panels; ++p) s += part_r[p * n + i];
    // column contributions: column i belongs to panel i/PANEL with c1 = min(n, (i/PANEL+1)*PANEL); chunks with r0 < c1
    size_t c1 = (i / SP_PANEL + 1) * SP_PANEL;
    if (c1 > n) c1 = n;
    const size_t n_chunks = (c1 + SP_ROWS - 1) / SP_ROWS;
    T t = T(0);
    for (size_t q = 0; q < n_chunks; ++q) t += part_c[q * n + i];
    T r = alpha * (s + t);
    if (beta != T(0)) r += beta * y[i];
    y[i] = r;
}

template <typename T> static void api_transform_sp(size_t n, T alpha, tb_view mat, tb_view x, T beta, tb_view y) {
    require_init();
    TB_REQUIRE(mat.len == n * (n + 1) / 2, "transform_sp: mat.len != n(n+1)/2");     // f64lapack.rs:151
    TB_REQUIRE(x.len == n && y.len == n, "transform_sp: vector length mismatch");     // :153-154
    const T* S = rptr<T>(mat);
    const T* px = rptr<T>(x);
    T* py = wptr<T>(y, beta == T(0));
    if (n == 0) return;
    Context& c = ctx();
    const size_t n_panels = (n + SP_PANEL - 1) / SP_PANEL;
    const size_t n_chunks = (n + SP_ROWS - 1) / SP_ROWS;
    size_t bytes_r = n_panels * n * sizeof(T);
    bytes_r = (bytes_r + 255) & ~size_t(255);
    size_t bytes_c = n_chunks * n * sizeof(T);
    char* sc = reinterpret_cast<char*>(scratch(bytes_r + bytes_c));
    T* part_r = reinterpret_cast<T*>(sc);
    T* part_c = reinterpret_cast<T*>(sc + bytes_r);
    dim3 grid((unsigned)n_panels, (unsigned)n_chunks);
    spmv_kernel<T><<<grid, SP_THREADS, 0, c.stream>>>(S, n, px, part_r, part_c);
    TB_LAUNCH_CHECK();
    spmv_finalize<T><<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(part_r, part_c, n, alpha, beta, py);
    TB_LAUNCH_CHECK();
}

}  // namespace tb

using namespace tb;
extern "C" {
int tb_transform_sp_f32(size_t n, float a, tb_view m, tb_view x, float b, tb_view y) { return api_defer({m, x, y}, {y}, [=] { api_transform_sp<float>(n, a, m, x, b, y); }); }
int tb_transform_sp_f64(size_t n, double a, tb_view m, tb_view x, double b, tb_view y) { return api_defer({m, x, y}, {y}, [=] { api_transform_sp<double>(n, a, m, x, b, y); }); }
}
