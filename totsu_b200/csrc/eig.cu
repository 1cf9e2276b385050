// LinAlgEx::map_eig (totsu_core/src/linalg_ex.rs:43-65) and the ConePSD projection built on it
// (totsu_core/src/cone_psd.rs:56-79).  CPU twin: totsu_f64lapack/src/f64lapack.rs:78-108 (dsyevr V/V/U
// (0,+inf] + a dsyr loop), :172-190 (map_eig), :195-255 (vec_to_mat / mat_to_vec).  Superseded GPU path:
// totsu_f32cuda/src/f32cuda.rs:218-370 (cusolverDnSsyevdx + up to k cublasSsyr launches + 2k cublasScopy).
//
// Work layout (2k^2 + k elements, same as the reference): a[k*k] | w[k] | z[k*k], all column-major.
//
//   1. unpack: packed upper triangle -> full symmetric a, diagonal scaled (the sqrt(2) convention, cone_psd.rs:18)
//   2. eigendecomposition by parallel one-sided Jacobi on B = a + s*I, s >= ||a||_F >= -lambda_min, so B is
//      symmetric positive semi-definite and its singular pairs ARE its eigenpairs: rotations orthogonalise the
//      columns of G (initially B) while V accumulates them; at convergence G = V*diag(sigma), lambda = sigma - s.
//      k/2 disjoint column pairs per step (round-robin tournament), one CTA per pair, k-1 steps per sweep.
//   3. reconstruct: a := sum_i coef_i z_i z_i^T as a tiled GEMM (Z*diag(coef))*Z^T over the upper tiles
//   4. pack: upper triangle back into the vector, diagonal unscaled.
// Round 1 keeps the GEMM on the FP32/FP64 pipes; moving step 3 (and a blocked two-sided variant of step 2)
// onto tcgen05 is tracked in DESIGN.md.
#include "common.cuh"
#include <cmath>

namespace tb {

constexpr int EG_THREADS = 128;

template <typename T>
__global__ void unpack_kernel(const T* __restrict__ x, size_t k, int has_scale, T scale, T* __restrict__ a) {
    size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= k * k) return;
    size_t r = idx % k, c = idx / k;
    size_t rr = r <= c ? r : c, cc = r <= c ? c : r;
    T v = x[cc * (cc + 1) / 2 + rr];
    if (r == c && has_scale) v *= scale;
    a[idx] = v;
}

template <typename T>
__global__ void shift_init_kernel(T* __restrict__ a, T* __restrict__ z, size_t k, T shift) {
    size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= k * k) return;
    size_t r = idx % k, c = idx / k;
    if (r == c) { a[idx] += shift; z[idx] = T(1); }
    else z[idx] = T(0);
}

// one tournament step: CTA `i` orthogonalises column pair (p, q) of G and applies the same rotation to V
template <typename T>
__global__ void __launch_bounds__(EG_THREADS) jacobi_step_kernel(T* __restrict__ g, T* __restrict__ v, int k, int kp, int step, double tol, unsigned int* rotations) {
    __shared__ double red[3][EG_THREADS / 32];
    __shared__ double cs[2];
    const int i = blockIdx.x;
    int p, q;
    if (i == 0) { p = kp - 1; q = step; }
    else { p = (step + i) % (kp - 1); q = (step - i + (kp - 1)) % (kp - 1); }
    if (p >= k || q >= k) return;                  // dummy index of an odd-sized problem (uniform per CTA)
    if (p > q) { int t = p; p = q; q = t; }
    T* gp = g + (size_t)p * k;
    T* gq = g + (size_t)q * k;
    double al = 0.0, be = 0.0, ga = 0.0;
    for (int r = threadIdx.x; r < k; r += EG_THREADS) {
        double a = (double)gp[r], b = (double)gq[r];
        al += a * a; be += b * b; ga += a * b;
    }
    al = tbd::warp_sum(al); be = tbd::warp_sum(be); ga = tbd::warp_sum(ga);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { red[0][w] = al; red[1][w] = be; red[2][w] = ga; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double A = 0, B = 0, G = 0;
        for (int j = 0; j < EG_THREADS / 32; ++j) { A += red[0][j]; B += red[1][j]; G += red[2][j]; }
        double c = 1.0, s = 0.0;
        if (fabs(G) > tol * sqrt(A * B) && G != 0.0) {
            double zeta = (B - A) / (2.0 * G);
            double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            c = 1.0 / sqrt(1.0 + t * t);
            s = c * t;
            atomicAdd(rotations, 1u);
        }
        cs[0] = c; cs[1] = s;
    }
    __syncthreads();
    const double c = cs[0], s = cs[1];
    if (s == 0.0) return;
    T* vp = v + (size_t)p * k;
    T* vq = v + (size_t)q * k;
    for (int r = threadIdx.x; r < k; r += EG_THREADS) {
        double a = (double)gp[r], b = (double)gq[r];
        gp[r] = (T)(c * a - s * b);
        gq[r] = (T)(s * a + c * b);
        double x = (double)vp[r], y = (double)vq[r];
        vp[r] = (T)(c * x - s * y);
        vq[r] = (T)(s * x + c * y);
    }
}

// w[i] = ||g_i|| - shift
template <typename T>
__global__ void __launch_bounds__(EG_THREADS) eigvals_kernel(const T* __restrict__ g, int k, double shift, T* __restrict__ w) {
    __shared__ double red[32];
    const T* gi = g + (size_t)blockIdx.x * k;
    double acc = 0.0;
    for (int r = threadIdx.x; r < k; r += EG_THREADS) { double a = (double)gi[r]; acc += a * a; }
    double tot = tbd::block_sum(acc, red);
    if (threadIdx.x == 0) w[blockIdx.x] = (T)(sqrt(tot) - shift);
}

// coef[i] = w[i] > 0 ? w[i] : 0   (ConePSD's closure, cone_psd.rs:69-76)
template <typename T> __global__ void psd_coef_kernel(T* w, int k) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) { T e = w[i]; w[i] = e > T(0) ? e : T(0); }
}

// C(upper tiles) = Z * diag(coef) * Z^T ; Z, C column-major k x k.  64x64 tile per CTA, 4x4 per thread.
template <typename T>
__global__ void __launch_bounds__(256) zdzt_kernel(const T* __restrict__ Z, const T* __restrict__ coef, int k, T* __restrict__ C) {
    constexpr int TS = 64, KS = 16;
    const int bi = blockIdx.x, bj = blockIdx.y;
    if (bi > bj) return;                           // only tiles touching the upper triangle
    __shared__ T As[KS][TS + 4];
    __shared__ T Bs[KS][TS + 4];
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const int i0 = bi * TS, j0 = bj * TS;
    T acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = T(0);
    for (int l0 = 0; l0 < k; l0 += KS) {
        // 256 threads load 16 x 64 elements of each operand: thread -> (l = tid/16, 4 consecutive rows)
        {
            const int l = threadIdx.x / 16, r4 = (threadIdx.x % 16) * 4;
            const int gl = l0 + l;
            const T cf = gl < k ? coef[gl] : T(0);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int gi = i0 + r4 + u, gj = j0 + r4 + u;
                As[l][r4 + u] = (gl < k && gi < k) ? Z[(size_t)gl * k + gi] * cf : T(0);
                Bs[l][r4 + u] = (gl < k && gj < k) ? Z[(size_t)gl * k + gj] : T(0);
            }
        }
        __syncthreads();
#pragma unroll
        for (int l = 0; l < KS; ++l) {
            T a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { a[u] = As[l][ty * 4 + u]; b[u] = Bs[l][tx * 4 + u]; }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int w = 0; w < 4; ++w) acc[u][w] += a[u] * b[w];
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            int gi = i0 + ty * 4 + u, gj = j0 + tx * 4 + w;
            if (gi < k && gj < k) C[(size_t)gj * k + gi] = acc[u][w];
        }
}

template <typename T>
__global__ void pack_kernel(const T* __restrict__ a, size_t k, int has_scale, T inv_scale, T* __restrict__ x) {
    size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= k * k) return;
    size_t r = idx % k, c = idx / k;
    if (r > c) return;
    T v = a[idx];
    if (r == c && has_scale) v *= inv_scale;
    x[c * (c + 1) / 2 + r] = v;
}

static size_t tri_dim(size_t sn) {
    size_t n = (size_t)((std::sqrt((double)(8 * sn + 1)) - 1.0) / 2.0 + 0.5);
    while (n * (n + 1) / 2 > sn) --n;
    while ((n + 1) * (n + 2) / 2 <= sn) ++n;
    return n;
}

// Steps 1-2: leaves eigenvalues in w (device), eigenvectors in z; returns nothing.  Syncs the stream.
template <typename T> static void eig_decompose(const T* x, size_t k, bool has_scale, T scale, T* a, T* w, T* z) {
    Context& c = ctx();
    const size_t kk = k * k;
    const unsigned gkk = (unsigned)((kk + 255) / 256);
    unpack_kernel<T><<<gkk, 256, 0, c.stream>>>(x, k, has_scale ? 1 : 0, scale, a);
    TB_LAUNCH_CHECK();
    const double fro = std::sqrt(l1_sumsq_sync<T>(a, kk));
    TB_REQUIRE(std::isfinite(fro), "map_eig: matrix has non-finite entries");
    const double shift = fro * (1.0 + 1.0 / 64.0);
    shift_init_kernel<T><<<gkk, 256, 0, c.stream>>>(a, z, k, (T)shift);
    TB_LAUNCH_CHECK();
    // the shift actually applied, after rounding to T
    const double shift_applied = (double)(T)shift;
    if (k > 1 && fro > 0.0) {
        const int kp = (int)((k + 1) / 2 * 2);
        const double tol = 8.0 * (sizeof(T) == 4 ? 1.1920929e-7 : 2.220446049250313e-16);
        unsigned int* rot = c.tickets + 32;
        const int max_sweeps = 40;
        for (int sweep = 0; sweep < max_sweeps; ++sweep) {
            TB_CUDA(cudaMemsetAsync(rot, 0, sizeof(unsigned int), c.stream));
            for (int step = 0; step < kp - 1; ++step) {
                jacobi_step_kernel<T><<<kp / 2, EG_THREADS, 0, c.stream>>>(a, z, (int)k, kp, step, tol, rot);
            }
            TB_CUDA(cudaGetLastError());
            count_launch(kp - 1);
            TB_CUDA(cudaMemcpyAsync(c.mailbox_host, rot, sizeof(unsigned int), cudaMemcpyDeviceToHost, c.stream));
            TB_CUDA(cudaStreamSynchronize(c.stream));
            if (*reinterpret_cast<unsigned int*>(c.mailbox_host) == 0) break;
        }
    }
    eigvals_kernel<T><<<(unsigned)k, EG_THREADS, 0, c.stream>>>(a, (int)k, shift_applied, w);
    TB_LAUNCH_CHECK();
}

// Steps 3-4 with coefficients already in w (device)
template <typename T> static void eig_reconstruct(T* x, size_t k, bool has_scale, T scale, T* a, const T* w, const T* z) {
    Context& c = ctx();
    const size_t kk = k * k;
    const unsigned tiles = (unsigned)((k + 63) / 64);
    zdzt_kernel<T><<<dim3(tiles, tiles), 256, 0, c.stream>>>(z, w, (int)k, a);
    TB_LAUNCH_CHECK();
    pack_kernel<T><<<(unsigned)((kk + 255) / 256), 256, 0, c.stream>>>(a, k, has_scale ? 1 : 0, has_scale ? T(1) / scale : T(1), x);
    TB_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// ConePSD fast path: projection by the matrix sign function, GEMM-only, no host round trip.
//
//   proj_{S+}(X) = V max(L,0) V^T = (X + X sign(X)) / 2,       sign(X) = V sign(L) V^T.
//
// sign(X) comes from polynomial iterations on S0 = X/||X||_F (all iterates are polynomials in X: symmetric,
// commuting, so only the upper triangle is computed and mirrored - which also keeps the iterates exactly symmetric):
//   phase 1  S <- S (a I + b S^2 + c S^4), (a,b,c) = (3.4445,-4.7750,2.0315): lifts an eigenvalue of relative
//            size eps to O(1) in log(1/eps)/log(3.44) steps (3 GEMMs each);
//   phase 2  S <- S (3 I - S^2)/2  (Newton-Schulz, 2 GEMMs): quadratic convergence of every |lambda| in (0, sqrt 3) to 1.
// Eigenvalues below ~3.44^-n1 relative to ||X||_F stay unconverged; their contribution to the projection error is
// bounded by their own magnitude, which is below the working precision for the step counts chosen here
// (f32: 10 + 8 steps = 47 GEMMs, f64: 24 + 12 = 97).  For the reference this replaces dsyevr(V,V,U,(0,inf]) +
// a dsyr loop (f64lapack.rs:78-108): same mathematical result, verified against the oracle in the tests.
// The general closure path (tb_map_eig_begin/finish, e.g. MatBuild::set_sqrt) keeps the Jacobi eigensolver.
// ---------------------------------------------------------------------------------------------------------
constexpr int SG_TILE = 32;      // output tile
constexpr int SG_KC = 64;        // K chunk per stage: 4 thread groups x 16
constexpr int SG_THREADS = 256;

// C = alpha * (A * B) + beta * D + gamma * I for symmetric k x k column-major A, B, D; C symmetric.
// Since A = A^T:  C[i,j] = sum_l A[l*k + i] * B[l*k + j]  - both operands are read along contiguous columns.
// 256 threads = 4 groups (split-K inside the chunk) x 64 threads (8 x 8, 4 x 4 micro-tile each).
template <typename T>
__global__ void __launch_bounds__(SG_THREADS) symm_gemm_kernel(const T* __restrict__ A, const T* __restrict__ B, const T* __restrict__ D,
                                                                T* __restrict__ C, int k, T alpha, T beta, T gamma,
                                                                const double* __restrict__ inv_norm_sq) {
    const int bi = blockIdx.x, bj = blockIdx.y;
    if (bi > bj) return;
    constexpr int RS = SG_TILE + 1;                            // padded row stride of the reduction tiles
    __shared__ __align__(16) T smem[4 * SG_TILE * RS];         // As | Bs (2 * 64 * 32), later reused for the split-K reduction
    T (*As)[SG_TILE] = reinterpret_cast<T (*)[SG_TILE]>(smem);
    T (*Bs)[SG_TILE] = reinterpret_cast<T (*)[SG_TILE]>(smem + SG_KC * SG_TILE);
    const int tid = threadIdx.x;
    const int grp = tid >> 6, t64 = tid & 63;
    const int tx = t64 & 7, ty = t64 >> 3;
    const int i0 = bi * SG_TILE, j0 = bj * SG_TILE;
    // loader mapping: thread -> (row l = tid / 4 in 0..63, 8 consecutive columns (tid % 4) * 8)
    const int ll = tid >> 2, lc = (tid & 3) * 8;
    T acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = T(0);
    T ra[8], rb[8];
    auto gload = [&](int l0) {
        const int gl = l0 + ll;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int gi = i0 + lc + u, gj = j0 + lc + u;
            ra[u] = (gl < k && gi < k) ? A[(size_t)gl * k + gi] : T(0);
            rb[u] = (gl < k && gj < k) ? B[(size_t)gl * k + gj] : T(0);
        }
    };
    gload(0);
    for (int l0 = 0; l0 < k; l0 += SG_KC) {
#pragma unroll
        for (int u = 0; u < 8; ++u) { As[ll][lc + u] = ra[u]; Bs[ll][lc + u] = rb[u]; }
        __syncthreads();
        if (l0 + SG_KC < k) gload(l0 + SG_KC);          // prefetch the next chunk while this one is consumed
#pragma unroll
        for (int l = 0; l < 16; ++l) {
            const int lr = grp * 16 + l;
            T a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { a[u] = As[lr][ty * 4 + u]; b[u] = Bs[lr][tx * 4 + u]; }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int w = 0; w < 4; ++w) acc[u][w] += a[u] * b[w];
        }
        __syncthreads();
    }
    // split-K reduction through shared memory: red[grp][i][j], padded rows
    T* red = smem;
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int w = 0; w < 4; ++w) red[(grp * SG_TILE + ty * 4 + u) * RS + tx * 4 + w] = acc[u][w];
    __syncthreads();
    // optional scaling by 1/||X||_F taken from device memory (first step only)
    T al = alpha;
    if (inv_norm_sq != nullptr) {
        const double ss = *inv_norm_sq;
        al = ss > 0.0 ? (T)((double)alpha / ss) : T(0);
    }
    // each thread finalises 4 entries: (i = tid % 32, j = tid / 32 + 8 q)
    const int ii = tid & 31;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int jj = (tid >> 5) + 8 * q;
        T v = red[(0 * SG_TILE + ii) * RS + jj] + red[(1 * SG_TILE + ii) * RS + jj] +
              red[(2 * SG_TILE + ii) * RS + jj] + red[(3 * SG_TILE + ii) * RS + jj];
        const int gi = i0 + ii, gj = j0 + jj;
        // Only entries on or above the diagonal are kept and then mirrored, so every iterate is EXACTLY symmetric:
        // the kernel really computes A*B^T, and an antisymmetric rounding residue in S is amplified by the
        // polynomial iterations (x3.4 per phase-1 step) until it destroys the result.
        if (gi < k && gj < k && gi <= gj) {
            v = al * v;
            if (beta != T(0)) v += beta * D[(size_t)gj * k + gi];
            if (gi == gj) v += gamma;
            C[(size_t)gj * k + gi] = v;
            if (gi != gj) C[(size_t)gi * k + gj] = v;    // mirror
        }
    }
}

// engine: 0 = choose (tensor cores for f32 when the operands qualify), 1 = FP32/FP64-pipe kernel above, 2 = tcgen05 (psd_tc.cu)
template <typename T>
static void symm_gemm(const T* A, const T* B, const T* D, T* C, size_t k, T alpha, T beta, T gamma, const double* inv_norm_sq = nullptr,
                      int engine = 0, int splitk = 0) {
    if constexpr (sizeof(T) == 4) {
        if (engine == 2 || (engine == 0 && inv_norm_sq == nullptr && k >= 64 && symm_gemm_tc_usable(A, B, D, C, k))) {
            TB_REQUIRE(inv_norm_sq == nullptr, "symm_gemm: the tensor-core engine takes alpha by value");
            symm_gemm_tc(A, B, D, C, k, alpha, beta, gamma, splitk);
            return;
        }
    } else {
        TB_REQUIRE(engine != 2, "symm_gemm: the tensor-core engine is f32 only (no FP64 tensor path on sm_100a worth using)");
    }
    const unsigned nt = (unsigned)((k + SG_TILE - 1) / SG_TILE);
    symm_gemm_kernel<T><<<dim3(nt, nt), SG_THREADS, 0, ctx().stream>>>(A, B, D, C, (int)k, alpha, beta, gamma, inv_norm_sq);
    TB_LAUNCH_CHECK();
}

// s = x * rsqrt(sumsq)  (0 if the matrix is zero)
template <typename T> __global__ void scale_inv_norm_kernel(const T* __restrict__ x, T* __restrict__ s, size_t n, const double* __restrict__ sumsq) {
    const double ss = *sumsq;
    const T inv = ss > 0.0 ? (T)(1.0 / sqrt(ss)) : T(0);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s[i] = x[i] * inv;
}

static char* eig_scratch(size_t bytes) {
    Context& c = ctx();
    if (bytes > c.eig_scratch_bytes) {
        TB_CUDA(cudaStreamSynchronize(c.stream));
        if (c.eig_scratch) TB_CUDA(cudaFree(c.eig_scratch));
        TB_CUDA(cudaMalloc(&c.eig_scratch, bytes));
        c.eig_scratch_bytes = bytes;
    }
    return c.eig_scratch;
}

// f32 schedule of the sign iteration: n1 quintic steps (3 GEMMs each, slope 3.4445 at 0) then n2 Newton-Schulz steps (2 GEMMs
// each, slope 1.5).  An eigenvalue lambda is resolved once the accumulated growth has carried |lambda| / ||X||_F to ~1; one that is
// not contributes an error <= |lambda| to the projection, so a growth of ~1e6 puts the unresolved ones at the f32 rounding floor
// of the GEMM chain (2e-6 of max|X|).  The quintic leaves resolved eigenvalues in [0.68, 1.2]; Newton-Schulz takes that to 1 within
// 1e-9 in 5 steps.  9 + 6 (growth 7.8e5, 40 GEMMs) measures the same errors as round 1's 10 + 8 (47 GEMMs) on random, log-spaced
// (12 decades), low-rank, near-boundary, clustered and definite spectra at k = 96 / 512 and is 14 % faster
// (profiles/r02_psd_schedule.md).  TB_PSD_STEPS="n1,n2" overrides it (scripts/psd_schedule.py).
static void psd_schedule_f32(int& n1, int& n2) {
    static int s1 = -1, s2 = -1;
    if (s1 < 0) {
        s1 = 9; s2 = 6;
        if (const char* e = std::getenv("TB_PSD_STEPS")) {
            int a = 0, b = 0;
            if (std::sscanf(e, "%d,%d", &a, &b) == 2 && a >= 1 && a <= 40 && b >= 0 && b <= 40) { s1 = a; s2 = b; }
        }
    }
    n1 = s1; n2 = s2;
}

template <typename T> static void psd_project_sign(T* x, size_t k, T* work) {
    Context& c = ctx();
    const size_t kk = k * k;
    T* X = work;                 // unpacked, diagonal scaled by sqrt(2)
    T* S0 = work + kk + k;       // the reference's z area
    T* sc = reinterpret_cast<T*>(eig_scratch(3 * kk * sizeof(T)));
    T* S1 = sc; T* T2 = sc + kk; T* P = sc + 2 * kk;
    double* sumsq = c.mailbox_dev + 16;
    const T sq2 = (T)1.41421356237309504880;
    const unsigned gkk = (unsigned)((kk + 255) / 256);
    unpack_kernel<T><<<gkk, 256, 0, c.stream>>>(x, k, 1, sq2, X);
    TB_LAUNCH_CHECK();
    l1_sumsq_async<T>(X, kk, sumsq);
    scale_inv_norm_kernel<T><<<std::min<unsigned>(gkk, 4096u), 256, 0, c.stream>>>(X, S0, kk, sumsq);
    TB_LAUNCH_CHECK();
    int n1 = 24, n2 = 12;
    if (sizeof(T) == 4) psd_schedule_f32(n1, n2);
    const T qa = (T)3.4445, qb = (T)-4.7750, qc = (T)2.0315;
    T* S = S0; T* Sn = S1;
    // psd_mode 0: tensor cores where they apply (f32, k >= 64, k % 4 == 0); 2: FP32/FP64-pipe GEMM; 3: tensor cores without split-K
    const int eng = c.psd_mode == 2 ? 1 : 0, sk = c.psd_mode == 3 ? 1 : 0;
    for (int it = 0; it < n1; ++it) {
        symm_gemm<T>(S, S, nullptr, T2, k, T(1), T(0), T(0), nullptr, eng, sk);            // T2 = S^2
        symm_gemm<T>(T2, T2, T2, P, k, qc, qb, qa, nullptr, eng, sk);                     // P = c S^4 + b S^2 + a I
        symm_gemm<T>(S, P, nullptr, Sn, k, T(1), T(0), T(0), nullptr, eng, sk);           // S' = S P
        std::swap(S, Sn);
    }
    for (int it = 0; it < n2; ++it) {
        symm_gemm<T>(S, S, nullptr, P, k, T(-0.5), T(0), T(1.5), nullptr, eng, sk);       // P = 1.5 I - 0.5 S^2
        symm_gemm<T>(S, P, nullptr, Sn, k, T(1), T(0), T(0), nullptr, eng, sk);           // S' = S P
        std::swap(S, Sn);
    }
    symm_gemm<T>(X, S, X, P, k, T(0.5), T(0.5), T(0), nullptr, eng, sk);                  // proj = (X S + X) / 2
    pack_kernel<T><<<gkk, 256, 0, c.stream>>>(P, k, 1, T(1) / sq2, x);
    TB_LAUNCH_CHECK();
}

// Both ConePSD projections of one solver iteration (dual cone on y, primal cone on s: solver.rs:548-549) as one batch:
// every step of the sign iteration is ONE launch of the tcgen05 GEMM over both problems (blockIdx.y), which fills
// twice as many SMs.  Problem 0 uses the caller's work buffer like psd_project_sign, problem 1 lives in the library's
// scratch.  f32, k >= 64, k % 4 == 0, 16-byte aligned views (psd_pair_usable).
bool psd_pair_usable(size_t sn, const float* work) {
    const size_t k = tri_dim(sn);
    return ctx().psd_mode == 0 && k * (k + 1) / 2 == sn && k >= 64 && k % 4 == 0 && (reinterpret_cast<uintptr_t>(work) & 15u) == 0;
}

void psd_project_pair(float* x0, float* x1, size_t sn, float* work, size_t work_len) {
    Context& c = ctx();
    const size_t k = tri_dim(sn), kk = k * k;
    TB_REQUIRE(work_len >= 2 * kk + k, "ConePSD: work shortage");
    float* sc = reinterpret_cast<float*>(eig_scratch((8 * kk + k) * sizeof(float)));
    float* X[2] = {work, sc + 3 * kk};
    float* S0[2] = {work + kk + k, sc + 3 * kk + kk + k};
    float* S1[2] = {sc, sc + 5 * kk + k};
    float* T2[2] = {sc + kk, sc + 6 * kk + k};
    float* P[2] = {sc + 2 * kk, sc + 7 * kk + k};
    float* xs[2] = {x0, x1};
    const float sq2 = 1.41421356237309504880f;
    const unsigned gkk = (unsigned)((kk + 255) / 256);
    for (int i = 0; i < 2; ++i) {
        double* sumsq = c.mailbox_dev + 16 + i;
        unpack_kernel<float><<<gkk, 256, 0, c.stream>>>(xs[i], k, 1, sq2, X[i]);
        TB_LAUNCH_CHECK();
        l1_sumsq_async<float>(X[i], kk, sumsq);
        scale_inv_norm_kernel<float><<<std::min<unsigned>(gkk, 4096u), 256, 0, c.stream>>>(X[i], S0[i], kk, sumsq);
        TB_LAUNCH_CHECK();
    }
    int n1, n2;
    psd_schedule_f32(n1, n2);                                   // same schedule as psd_project_sign<float>
    const float qa = 3.4445f, qb = -4.7750f, qc = 2.0315f;
    const float* nul[2] = {nullptr, nullptr};
    float* S[2] = {S0[0], S0[1]};
    float* Sn[2] = {S1[0], S1[1]};
    for (int it = 0; it < n1; ++it) {
        symm_gemm_tc_pair(S, S, nul, T2, k, 1.f, 0.f, 0.f, 0);             // T2 = S^2
        symm_gemm_tc_pair(T2, T2, T2, P, k, qc, qb, qa, 0);               // P = c S^4 + b S^2 + a I
        symm_gemm_tc_pair(S, P, nul, Sn, k, 1.f, 0.f, 0.f, 0);            // S' = S P
        std::swap(S[0], Sn[0]); std::swap(S[1], Sn[1]);
    }
    for (int it = 0; it < n2; ++it) {
        symm_gemm_tc_pair(S, S, nul, P, k, -0.5f, 0.f, 1.5f, 0);          // P = 1.5 I - 0.5 S^2
        symm_gemm_tc_pair(S, P, nul, Sn, k, 1.f, 0.f, 0.f, 0);            // S' = S P
        std::swap(S[0], Sn[0]); std::swap(S[1], Sn[1]);
    }
    symm_gemm_tc_pair(X, S, X, P, k, 0.5f, 0.5f, 0.f, 0);                 // proj = (X S + X) / 2
    for (int i = 0; i < 2; ++i) {
        pack_kernel<float><<<gkk, 256, 0, c.stream>>>(P[i], k, 1, 1.f / sq2, xs[i]);
        TB_LAUNCH_CHECK();
    }
}

template <typename T> void psd_project(T* x, size_t sn, T eps_zero, T* work, size_t work_len) {
    (void)eps_zero;
    const size_t k = tri_dim(sn);
    TB_REQUIRE(k * (k + 1) / 2 == sn, "ConePSD: length is not a triangular number");       // cone_psd.rs:35
    TB_REQUIRE(work_len >= 2 * k * k + k, "ConePSD: work shortage");                        // cone_psd.rs:58-61
    if (k == 0) return;
    if (ctx().psd_mode != 1 && k > 1) {
        psd_project_sign<T>(x, k, work);
        return;
    }
    T* a = work;
    T* w = work + k * k;
    T* z = w + k;
    const T sq2 = (T)1.41421356237309504880;
    eig_decompose<T>(x, k, true, sq2, a, w, z);
    psd_coef_kernel<T><<<(unsigned)((k + 127) / 128), 128, 0, ctx().stream>>>(w, (int)k);
    TB_LAUNCH_CHECK();
    eig_reconstruct<T>(x, k, true, sq2, a, w, z);
}
template void psd_project<float>(float*, size_t, float, float*, size_t);
template void psd_project<double>(double*, size_t, double, double*, size_t);

template <typename T>
static void api_map_eig_begin(tb_view mat, int has_scale, T scale_diag, T eps_zero, tb_view work, T* host_eigs) {
    (void)eps_zero;
    require_init();
    const size_t k = tri_dim(mat.len);
    TB_REQUIRE(k * (k + 1) / 2 == mat.len, "map_eig: length is not a triangular number");   // f64lapack.rs:177-178
    TB_REQUIRE(work.len >= 2 * k * k + k, "map_eig: work shortage");                        // f64lapack.rs:180
    const T* x = rptr<T>(mat);
    T* wk = wptr<T>(work);
    if (k == 0) return;
    T* a = wk; T* w = wk + k * k; T* z = w + k;
    eig_decompose<T>(x, k, has_scale != 0, scale_diag, a, w, z);
    TB_CUDA(cudaMemcpyAsync(host_eigs, w, k * sizeof(T), cudaMemcpyDeviceToHost, ctx().stream));
    TB_CUDA(cudaStreamSynchronize(ctx().stream));
}

template <typename T>
static void api_map_eig_finish(tb_view mat, int has_scale, T scale_diag, tb_view work, const T* new_eigs, const uint8_t* keep) {
    require_init();
    const size_t k = tri_dim(mat.len);
    TB_REQUIRE(k * (k + 1) / 2 == mat.len, "map_eig: length is not a triangular number");
    TB_REQUIRE(work.len >= 2 * k * k + k, "map_eig: work shortage");
    T* x = wptr<T>(mat, true);
    T* wk = wptr<T>(work);
    if (k == 0) return;
    T* a = wk; T* w = wk + k * k; T* z = w + k;
    std::vector<T> coef(k);
    for (size_t i = 0; i < k; ++i) coef[i] = keep[i] ? new_eigs[i] : T(0);
    TB_CUDA(cudaMemcpyAsync(w, coef.data(), k * sizeof(T), cudaMemcpyHostToDevice, ctx().stream));
    TB_CUDA(cudaStreamSynchronize(ctx().stream));     // coef is a stack-owned staging vector
    eig_reconstruct<T>(x, k, has_scale != 0, scale_diag, a, w, z);
}

// ---------------------------------------------------------------------------------------------------------
// MatBuild::set_sqrt at real size (totsu/src/matbuild/mod.rs:220-241: map_eig with the closure e -> sqrt(e), called once
// per ProbQP / ProbQCQP construction, qp.rs:386, qcqp.rs:445-448): P^(1/2) of a symmetric PSD matrix WITHOUT an
// eigendecomposition - the coupled Newton-Schulz iteration, three symmetric GEMMs per step on the same engine as the
// ConePSD projection (tcgen05 3xTF32 for f32, FP64 pipe for f64):
//     Y0 = P / ||P||_F,  Z0 = I;     R = I - Z Y;   Y <- Y + Y R / 2;   Z <- Z + Z R / 2;      Y -> (P/||P||_F)^(1/2)
// (all iterates are polynomials in P: symmetric and commuting, so the symmetric-operand GEMM applies).  ||R||_F is read
// back after every step - one 8-byte round trip next to three k^3 GEMMs - and the loop ends when it is at rounding level
// or has stopped shrinking.  Returns false (matrix untouched) if the iteration does not converge - P indefinite or
// numerically singular beyond what the working precision resolves - so that the caller can take the Jacobi route.
// ---------------------------------------------------------------------------------------------------------
template <typename T> __global__ void identity_kernel(T* __restrict__ a, size_t k) {
    size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= k * k) return;
    a[idx] = (idx % k == idx / k) ? T(1) : T(0);
}
template <typename T> __global__ void scale_sqrt_norm_kernel(T* __restrict__ y, size_t n, const double* __restrict__ sumsq) {
    const T f = (T)sqrt(sqrt(*sumsq));          // sqrt(||P||_F)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] *= f;
}

template <typename T> static bool psd_sqrt_newton_schulz(T* x, size_t k, int* iters_out) {
    Context& c = ctx();
    const size_t kk = k * k;
    T* sc = reinterpret_cast<T*>(eig_scratch(6 * kk * sizeof(T)));
    T* X = sc; T* Y = sc + kk; T* Z = sc + 2 * kk; T* R = sc + 3 * kk; T* Yn = sc + 4 * kk; T* Zn = sc + 5 * kk;
    double* sumsq = c.mailbox_dev + 20;
    double* rsq = c.mailbox_dev + 21;
    const unsigned gkk = (unsigned)((kk + 255) / 256);
    unpack_kernel<T><<<gkk, 256, 0, c.stream>>>(x, k, 0, T(1), X);
    TB_LAUNCH_CHECK();
    l1_sumsq_async<T>(X, kk, sumsq);
    scale_inv_norm_kernel<T><<<std::min<unsigned>(gkk, 4096u), 256, 0, c.stream>>>(X, Y, kk, sumsq);
    TB_LAUNCH_CHECK();
    identity_kernel<T><<<gkk, 256, 0, c.stream>>>(Z, k);
    TB_LAUNCH_CHECK();
    double nrm2 = 0.0;
    TB_CUDA(cudaMemcpyAsync(&nrm2, sumsq, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    TB_CUDA(cudaStreamSynchronize(c.stream));
    if (!(nrm2 > 0.0)) { if (iters_out) *iters_out = 0; return std::isfinite(nrm2); }      // the zero matrix is its own square root
    const double floor_r = (sizeof(T) == 4 ? 2e-6 : 1e-14) * (double)k;        // ||R||_F at rounding level: ~eps * sqrt(k) per entry, k^2 entries
    const int max_it = 100;
    double r_prev = 1e300;
    bool ok = false;
    int it = 0;
    for (; it < max_it; ++it) {
        symm_gemm<T>(Z, Y, nullptr, R, k, T(-1), T(0), T(1));                  // R = I - Z Y
        l1_sumsq_async<T>(R, kk, rsq);
        double r2 = 0.0;
        TB_CUDA(cudaMemcpyAsync(&r2, rsq, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        symm_gemm<T>(Y, R, Y, Yn, k, T(0.5), T(1), T(0));                      // Y' = Y + Y R / 2   (enqueued while the host waits for r)
        symm_gemm<T>(Z, R, Z, Zn, k, T(0.5), T(1), T(0));                      // Z' = Z + Z R / 2
        TB_CUDA(cudaStreamSynchronize(c.stream));
        const double r = sqrt(r2);
        if (!std::isfinite(r) || r > 4.0 * (double)k) break;                   // diverging: not PSD
        std::swap(Y, Yn); std::swap(Z, Zn);
        if (r <= floor_r || (r_prev < 0.25 && r > 0.5 * r_prev)) { ok = true; ++it; break; }
        r_prev = r;
    }
    if (iters_out) *iters_out = it;
    if (!ok) return false;
    scale_sqrt_norm_kernel<T><<<std::min<unsigned>(gkk, 4096u), 256, 0, c.stream>>>(Y, kk, sumsq);
    TB_LAUNCH_CHECK();
    pack_kernel<T><<<gkk, 256, 0, c.stream>>>(Y, k, 0, T(1), x);
    TB_LAUNCH_CHECK();
    return true;
}

static int g_last_sqrt_iters = 0, g_last_sqrt_route = 0;     // diagnostics: tb_sqrt_psd_info

// map_eig(mat, None, eps_zero, work, |e| if e > 0 { Some(sqrt(e)) } else { None }) - exactly MatBuild::set_sqrt's call
template <typename T> static void api_sqrt_psd(tb_view mat, T eps_zero, tb_view work) {
    require_init();
    const size_t k = tri_dim(mat.len);
    TB_REQUIRE(k * (k + 1) / 2 == mat.len, "sqrt_psd: length is not a triangular number");
    TB_REQUIRE(work.len >= 2 * k * k + k, "sqrt_psd: work shortage");                      // matbuild/mod.rs:226-229
    if (k == 0) return;
    T* x = wptr<T>(mat);
    g_last_sqrt_route = 0;
    if (k >= 32 && ctx().psd_mode != 1 && psd_sqrt_newton_schulz<T>(x, k, &g_last_sqrt_iters)) {
        g_last_sqrt_route = 1;
        return;
    }
    // small matrices and anything the GEMM-only iteration cannot resolve: eigendecomposition, sqrt of the positive eigenvalues
    T* wk = wptr<T>(work);
    T* a = wk; T* w = wk + k * k; T* z = w + k;
    eig_decompose<T>(x, k, false, T(1), a, w, z);
    std::vector<T> ev(k);
    TB_CUDA(cudaMemcpyAsync(ev.data(), w, k * sizeof(T), cudaMemcpyDeviceToHost, ctx().stream));
    TB_CUDA(cudaStreamSynchronize(ctx().stream));
    for (size_t i = 0; i < k; ++i) ev[i] = ev[i] > T(0) ? (T)std::sqrt(ev[i]) : T(0);
    (void)eps_zero;
    TB_CUDA(cudaMemcpyAsync(w, ev.data(), k * sizeof(T), cudaMemcpyHostToDevice, ctx().stream));
    TB_CUDA(cudaStreamSynchronize(ctx().stream));
    eig_reconstruct<T>(x, k, false, T(1), a, w, z);
    g_last_sqrt_route = 2;
}

template <typename T> static void api_proj_psd(tb_view x, T eps_zero, tb_view work) {
    require_init();
    T* px = wptr<T>(x);
    T* wk = wptr<T>(work);
    psd_project<T>(px, x.len, eps_zero, wk, work.len);
}

}  // namespace tb

using namespace tb;
extern "C" {
int tb_set_psd_path(int mode) {
    return api([&] {
        TB_REQUIRE(mode >= 0 && mode <= 3, "mode must be 0 (matrix-sign iteration, tcgen05 GEMMs for f32), 1 (Jacobi eigendecomposition), "
                                           "2 (matrix-sign iteration on the FP32/FP64 pipes) or 3 (tcgen05 without split-K)");
        ctx().psd_mode = mode;
    });
}
size_t tb_map_eig_worklen(size_t n) { return 2 * n * n + n; }    // f64lapack.rs:165-170 (len_a + len_w + len_z)
int tb_map_eig_begin_f32(tb_view m, int hs, float sd, float ez, tb_view w, float* e) { return api([&] { api_map_eig_begin<float>(m, hs, sd, ez, w, e); }); }
int tb_map_eig_begin_f64(tb_view m, int hs, double sd, double ez, tb_view w, double* e) { return api([&] { api_map_eig_begin<double>(m, hs, sd, ez, w, e); }); }
int tb_map_eig_finish_f32(tb_view m, int hs, float sd, tb_view w, const float* ne, const uint8_t* k) { return api([&] { api_map_eig_finish<float>(m, hs, sd, w, ne, k); }); }
int tb_map_eig_finish_f64(tb_view m, int hs, double sd, tb_view w, const double* ne, const uint8_t* k) { return api([&] { api_map_eig_finish<double>(m, hs, sd, w, ne, k); }); }
int tb_symm_gemm_f32(size_t k, float alpha, tb_view a, tb_view b, float beta, tb_view d, float gamma, tb_view cv, int engine, int splitk) {
    return api([&] {
        require_init();
        TB_REQUIRE(a.len == k * k && b.len == k * k && cv.len == k * k && (d.len == 0 || d.len == k * k), "symm_gemm: every matrix is k x k");
        TB_REQUIRE(engine == 1 || engine == 2, "symm_gemm: engine is 1 (FP32 pipe) or 2 (tcgen05)");
        const float* pa = rptr<float>(a);
        const float* pb = rptr<float>(b);
        const float* pd = d.len ? rptr<float>(d) : nullptr;
        float* pc = wptr<float>(cv, true);
        symm_gemm<float>(pa, pb, pd, pc, k, alpha, beta, gamma, nullptr, engine, splitk);
    });
}
int tb_symm_gemm_trace_f32(size_t k, tb_view a, tb_view b, tb_view cv, int splitk, int reps, uint64_t* stamps_ns) {
    return api([&] {
        require_init();
        TB_REQUIRE(a.len == k * k && b.len == k * k && cv.len == k * k && reps >= 1, "symm_gemm_trace: every matrix is k x k");
        const float* pa = rptr<float>(a);
        const float* pb = rptr<float>(b);
        float* pc = wptr<float>(cv, true);
        unsigned long long* dev = nullptr;
        TB_CUDA(cudaMalloc(&dev, 16 * sizeof(unsigned long long)));
        TB_CUDA(cudaMemsetAsync(dev, 0, 16 * sizeof(unsigned long long), ctx().stream));
        for (int r = 0; r < reps; ++r) symm_gemm_tc(pa, pb, nullptr, pc, k, 1.f, 0.f, 0.f, splitk, r == reps - 1 ? dev : nullptr);
        TB_CUDA(cudaMemcpyAsync(stamps_ns, dev, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx().stream));
        TB_CUDA(cudaStreamSynchronize(ctx().stream));
        TB_CUDA(cudaFree(dev));
    });
}
int tb_sqrt_psd_f32(tb_view m, float ez, tb_view w) { return api([&] { api_sqrt_psd<float>(m, ez, w); }); }
int tb_sqrt_psd_f64(tb_view m, double ez, tb_view w) { return api([&] { api_sqrt_psd<double>(m, ez, w); }); }
int tb_sqrt_psd_info(int* route, int* iterations) {
    return api_raw([&] { *route = g_last_sqrt_route; *iterations = g_last_sqrt_iters; });
}
int tb_proj_psd_f32(tb_view x, float ez, tb_view w) { return api([&] { api_proj_psd<float>(x, ez, w); }); }
int tb_proj_psd_f64(tb_view x, double ez, tb_view w) { return api([&] { api_proj_psd<double>(x, ez, w); }); }
}
