// LinAlg level-1 operations (reference trait: totsu_core/src/solver/linalg.rs:22-67; CPU twin:
// totsu_f64lapack/src/f64lapack.rs:20-73; superseded cuBLAS calls: totsu_f32cuda/src/f32cuda.rs:27-136).
//
// All vectors here are sub-slices of the solver's `work` array at arbitrary element offsets (e.g. the dual
// block starts at n+2m+1), so the kernels use scalar coalesced accesses: the vectors are <= a few MB, live in
// L2 and these launches are latency-bound, not bandwidth-bound.  Reductions accumulate in double and are
// two-stage with a fixed summation order (bit-reproducible run to run); the last block to finish combines
// the per-block partials, so one launch yields the scalar.
#include "common.cuh"
#include "vprog.cuh"

namespace tb {

static constexpr int kThreads = 256;

static inline int grid_for(size_t n, int per_thread = 4) {
    size_t blocks = (n + (size_t)kThreads * per_thread - 1) / ((size_t)kThreads * per_thread);
    size_t cap = (size_t)ctx().sm_count * 8;
    return (int)std::max<size_t>(1, std::min(blocks, cap));
}

template <typename T> __global__ void fill_kernel(T* x, T v, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = v;
}
template <typename T> __global__ void scale_kernel(T alpha, T* x, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = alpha * x[i];
}
template <typename T> __global__ void copy_kernel(const T* __restrict__ x, T* __restrict__ y, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = x[i];
}
// y = alpha*x + beta*y ; BETA_MODE 0: beta==0 (y not read), 1: beta==1, 2: general
template <typename T, int BETA_MODE> __global__ void axpby_kernel(T alpha, const T* __restrict__ x, T beta, T* y, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        T r = alpha * x[i];
        if (BETA_MODE == 1) r += y[i];
        if (BETA_MODE == 2) r += beta * y[i];
        y[i] = r;
    }
}
template <typename T> __global__ void adds_kernel(T s, T* y, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] += s;
}
template <typename T, int BETA_MODE>
__global__ void diag_kernel(T alpha, const T* __restrict__ d, const T* __restrict__ x, T beta, T* y, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        T r = alpha * (d[i] * x[i]);
        if (BETA_MODE == 1) r += y[i];
        if (BETA_MODE == 2) r += beta * y[i];
        y[i] = r;
    }
}
template <typename T> __global__ void recip_clamp_kernel(T eps, T* x, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        T v = x[i];
        v = v > eps ? v : eps;        // f.max(eps_zero)
        x[i] = T(1) / v;              // .recip()
    }
}

// MODE 0: sum of squares, MODE 1: sum of |x[i*inc]|
template <typename T, int MODE>
__global__ void reduce_kernel(const T* __restrict__ x, size_t count, size_t inc, double* partials, unsigned int* ticket, double* out,
                              double* box, unsigned long long seq) {
    __shared__ double red[32];
    __shared__ bool last;
    tbd::pdl_entry();
    double acc = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        double v = (double)x[i * inc];
        acc += (MODE == 0) ? v * v : fabs(v);
    }
    double s = tbd::block_sum(acc, red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = s;
        __threadfence();
        unsigned int t = atomicInc(ticket, gridDim.x - 1);   // wraps back to 0 after the last block
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        double a = 0.0;
        for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) a += __ldcg(&partials[i]);
        double tot = tbd::block_sum(a, red);
        if (threadIdx.x == 0) {
            if (box != nullptr) tbd::box_post(box, tot, seq);     // host-visible: straight into mapped pinned memory
            else *out = tot;
        }
    }
}

template <typename T, int MODE> static double reduce_sync(const T* x, size_t count, size_t inc) {
    Context& c = ctx();
    if (count == 0) return 0.0;
    if (vp_enabled_red(count)) {      // the reduction closes the pending vector program and posts into the host box
        const double r = vp_reduce_to_host(DT<T>::id, MODE, x, count, inc);
        dist_check_fault();
        return r;
    }
    int g = grid_for(count, 8);
    double* partials = reinterpret_cast<double*>(scratch((size_t)g * sizeof(double)));
    const uint64_t seq = box_next();
    launch_pdl(reduce_kernel<T, MODE>, dim3(g), dim3(kThreads), 0, c.stream, x, count, inc, partials, c.tickets, (double*)nullptr, c.hostbox_dev, (unsigned long long)seq);
    TB_LAUNCH_CHECK();
    const double r = box_wait(seq);
    dist_check_fault();
    return r;
}

template <typename T> void l1_scale(T alpha, T* x, size_t n) {
    if (n == 0) return;
    if (vp_enabled_wide(n)) {
        if (alpha == T(0)) vp_fill(DT<T>::id, x, 0.0, n);
        else vp_scale(DT<T>::id, (double)alpha, x, n);
        return;
    }
    Context& c = ctx();
    if (alpha == T(0)) fill_kernel<T><<<grid_for(n), kThreads, 0, c.stream>>>(x, T(0), n);
    else scale_kernel<T><<<grid_for(n), kThreads, 0, c.stream>>>(alpha, x, n);
    TB_LAUNCH_CHECK();
}
template <typename T> void l1_copy(const T* x, T* y, size_t n) {
    if (n == 0) return;
    if (vp_enabled_wide(n)) { vp_copy(DT<T>::id, x, y, n); return; }
    copy_kernel<T><<<grid_for(n), kThreads, 0, ctx().stream>>>(x, y, n);
    TB_LAUNCH_CHECK();
}
template <typename T> void l1_axpby(T alpha, const T* x, T beta, T* y, size_t n) {
    if (n == 0) return;
    if (vp_enabled_wide(n)) { vp_axpby(DT<T>::id, (double)alpha, x, (double)beta, y, n); return; }
    Context& c = ctx();
    int g = grid_for(n);
    if (beta == T(0)) axpby_kernel<T, 0><<<g, kThreads, 0, c.stream>>>(alpha, x, beta, y, n);
    else if (beta == T(1)) axpby_kernel<T, 1><<<g, kThreads, 0, c.stream>>>(alpha, x, beta, y, n);
    else axpby_kernel<T, 2><<<g, kThreads, 0, c.stream>>>(alpha, x, beta, y, n);
    TB_LAUNCH_CHECK();
}
template <typename T> double l1_sumsq_sync(const T* x, size_t n) { return reduce_sync<T, 0>(x, n, 1); }
// sum of squares left in device memory (*out_dev), no host round trip
template <typename T> void l1_sumsq_async(const T* x, size_t n, double* out_dev) {
    Context& c = ctx();
    if (n == 0) { TB_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double), c.stream)); return; }
    int g = grid_for(n, 8);
    double* partials = reinterpret_cast<double*>(scratch((size_t)g * sizeof(double)));
    launch_pdl(reduce_kernel<T, 0>, dim3(g), dim3(kThreads), 0, c.stream, x, n, (size_t)1, partials, c.tickets, out_dev, (double*)nullptr, 0ull);
    TB_LAUNCH_CHECK();
}
template void l1_sumsq_async<float>(const float*, size_t, double*);
template void l1_sumsq_async<double>(const double*, size_t, double*);

template void l1_scale<float>(float, float*, size_t);
template void l1_scale<double>(double, double*, size_t);
template void l1_copy<float>(const float*, float*, size_t);
template void l1_copy<double>(const double*, double*, size_t);
template void l1_axpby<float>(float, const float*, float, float*, size_t);
template void l1_axpby<double>(double, const double*, double, double*, size_t);
template double l1_sumsq_sync<float>(const float*, size_t);
template double l1_sumsq_sync<double>(const double*, size_t);

// ---- API bodies ------------------------------------------------------------------------------------------
template <typename T> static void api_norm(tb_view x, T* out) {
    require_init();
    double ss = 0.0;
    if (pf_try_sumsq(DT<T>::id, x, &ss)) {          // rode on an earlier round trip (prefetch.cu)
        *out = (T)sqrt(ss);
        return;
    }
    const T* p = rptr<T>(x);
    *out = (T)sqrt(reduce_sync<T, 0>(p, x.len, 1));
}
template <typename T> static void api_copy(tb_view x, tb_view y) {
    require_init();
    TB_REQUIRE(x.len == y.len, "copy: length mismatch");          // f64lapack.rs:27
    const T* px = rptr<T>(x);
    T* py = wptr<T>(y, true);
    l1_copy<T>(px, py, x.len);
}
template <typename T> static void api_scale(T alpha, tb_view x) {
    require_init();
    T* p = wptr<T>(x, alpha == T(0));
    l1_scale<T>(alpha, p, x.len);
}
template <typename T> static void api_add(T alpha, tb_view x, tb_view y) {
    require_init();
    TB_REQUIRE(x.len == y.len, "add: length mismatch");           // f64lapack.rs:39
    const T* px = rptr<T>(x);
    T* py = wptr<T>(y);
    l1_axpby<T>(alpha, px, T(1), py, x.len);
}
template <typename T> static void api_adds(T s, tb_view y) {
    require_init();
    T* py = wptr<T>(y);
    if (y.len == 0) return;
    if (vp_enabled_wide(y.len)) { vp_adds(DT<T>::id, (double)s, py, y.len); return; }
    adds_kernel<T><<<grid_for(y.len), kThreads, 0, ctx().stream>>>(s, py, y.len);
    TB_LAUNCH_CHECK();
}
template <typename T> static void api_abssum(tb_view x, size_t incx, T* out) {
    require_init();
    if (incx == 0) { *out = T(0); return; }                      // f64lapack.rs:53-55
    const T* p = rptr<T>(x);
    size_t count = (x.len + (incx - 1)) / incx;                   // f64lapack.rs:57
    *out = (T)reduce_sync<T, 1>(p, count, incx);
}
template <typename T> static void api_transform_di(T alpha, tb_view mat, tb_view x, T beta, tb_view y) {
    require_init();
    TB_REQUIRE(mat.len == x.len && mat.len == y.len, "transform_di: length mismatch");   // f64lapack.rs:63-64
    const T* pd = rptr<T>(mat);
    const T* px = rptr<T>(x);
    T* py = wptr<T>(y, beta == T(0));
    size_t n = y.len;
    if (n == 0) return;
    if (vp_enabled_wide(n)) { vp_diag(DT<T>::id, (double)alpha, pd, px, (double)beta, py, n); return; }
    Context& c = ctx();
    int g = grid_for(n);
    if (beta == T(0)) diag_kernel<T, 0><<<g, kThreads, 0, c.stream>>>(alpha, pd, px, beta, py, n);
    else if (beta == T(1)) diag_kernel<T, 1><<<g, kThreads, 0, c.stream>>>(alpha, pd, px, beta, py, n);
    else diag_kernel<T, 2><<<g, kThreads, 0, c.stream>>>(alpha, pd, px, beta, py, n);
    TB_LAUNCH_CHECK();
}
template <typename T> static void api_recip_clamp(T eps, tb_view x) {
    require_init();
    T* p = wptr<T>(x);
    if (x.len == 0) return;
    recip_clamp_kernel<T><<<grid_for(x.len), kThreads, 0, ctx().stream>>>(eps, p, x.len);
    TB_LAUNCH_CHECK();
}

}  // namespace tb

using namespace tb;

extern "C" {
int tb_norm_f32(tb_view x, float* out) { return api([&] { api_norm<float>(x, out); }); }
int tb_norm_f64(tb_view x, double* out) { return api([&] { api_norm<double>(x, out); }); }
int tb_copy_f32(tb_view x, tb_view y) { return api_defer({x}, {y}, [=] { api_copy<float>(x, y); }); }
int tb_copy_f64(tb_view x, tb_view y) { return api_defer({x}, {y}, [=] { api_copy<double>(x, y); }); }
int tb_scale_f32(float a, tb_view x) { return api_defer({x}, {x}, [=] { api_scale<float>(a, x); }); }
int tb_scale_f64(double a, tb_view x) { return api_defer({x}, {x}, [=] { api_scale<double>(a, x); }); }
int tb_add_f32(float a, tb_view x, tb_view y) { return api_defer({x, y}, {y}, [=] { api_add<float>(a, x, y); }); }
int tb_add_f64(double a, tb_view x, tb_view y) { return api_defer({x, y}, {y}, [=] { api_add<double>(a, x, y); }); }
int tb_adds_f32(float s, tb_view y) { return api_defer({y}, {y}, [=] { api_adds<float>(s, y); }); }
int tb_adds_f64(double s, tb_view y) { return api_defer({y}, {y}, [=] { api_adds<double>(s, y); }); }
int tb_abssum_f32(tb_view x, size_t incx, float* out) { return api([&] { api_abssum<float>(x, incx, out); }); }
int tb_abssum_f64(tb_view x, size_t incx, double* out) { return api([&] { api_abssum<double>(x, incx, out); }); }
int tb_transform_di_f32(float a, tb_view m, tb_view x, float b, tb_view y) { return api_defer({m, x, y}, {y}, [=] { api_transform_di<float>(a, m, x, b, y); }); }
int tb_transform_di_f64(double a, tb_view m, tb_view x, double b, tb_view y) { return api_defer({m, x, y}, {y}, [=] { api_transform_di<double>(a, m, x, b, y); }); }
int tb_recip_clamp_f32(float eps, tb_view x) { return api([&] { api_recip_clamp<float>(eps, x); }); }
int tb_recip_clamp_f64(double eps, tb_view x) { return api([&] { api_recip_clamp<double>(eps, x); }); }
}
