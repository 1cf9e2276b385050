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
#include "ptx.cuh"
#include <map>

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


// ---------------------------------------------------------------------------------------------------------
// Streaming path: the packed triangle read once at HBM speed.
//
// Same machine as gemv.cu's stream_kernel (persistent CTAs, one producer warp feeding a 6-stage shared-memory ring with
// TMA bulk copies, 8 consumer warps), adapted to the packed layout: column c is the contiguous run S[0..c, c] at
// element c(c+1)/2.  The triangle is cut into row chunks of TR rows and column tiles of 8; a tile of chunk R holds, for
// each of its columns, the segment rows [R*TR, min((R+1)*TR, c+1)).  A segment starts at an arbitrary element, so the
// producer copies the enclosing 16-byte-aligned window (<= VEC-1 elements of slack on each side, still inside the array)
// and the consumers index it with the per-column shift.  Every staged element a = S[r,c] is used twice while it is in
// shared memory: y[r] += a x[c] (thread-owned rows, r < c) and y[c] += a x[r] (one warp per column, r <= c).
// Work units (chunk, tile range) are sized to ~4 per CTA and dealt to CTAs by cumulative bytes (the triangle makes
// chunks unequal).  Partials go to scratch; spmv_stream_finalize adds them in a fixed order: no atomics.
// The last total % VEC elements of the array cannot be fetched by an aligned window without reading past the end; the
// finalize kernel adds their (<= 3) contributions directly.
// ---------------------------------------------------------------------------------------------------------
constexpr int SPS_WARPS = 8;             // consumer warps of the base variant; the chunk height TR = 256 * VEC is fixed by it
constexpr int SPS_CONSUMERS = SPS_WARPS * 32;
constexpr int SPS_THREADS = SPS_CONSUMERS + 32;
constexpr int SPS_TC = 8;               // columns per tile
// H = 1 (default): 8 consumer warps (one per column in the column pass, 4 rows per thread in the row pass).  H = 2: 16 consumer
// warps - two warps share a column (upper / lower half of the chunk's rows) and a thread owns 2 rows: the same shared-memory
// reads spread over twice the warps.  Built to test the round-1 reading of the ncu capture ("consumers latency-bound, 2 warps per
// scheduler"); measured on B200 it is 7-19 % SLOWER (profiles/r02_spmv_summary.md): the consumers were never the limit, the
// static split of the triangle over the CTAs was (SMs idle 36 % of the kernel) - hence the unit queue below.
// tb_set_spmv_warps selects; the finalize kernel sums H column-pass partials per (chunk, column).
constexpr int SPS_MAX_UNIT_COLS = 2048;
constexpr size_t SPS_XRES_BYTES = 49152;     // x resident in shared memory up to this size

struct SpUnit { int chunk, split, tile0, tile1; };
static int g_spmv_halves = 1;            // consumer warps / 8 of the streaming kernel (tb_set_spmv_warps); 8 warps measured faster (below)

// Shift of column c's segment inside its 16-byte-aligned window: (c(c+1)/2 + row0) mod VEC.  row0 is a multiple of VEC and
// a tile starts at a multiple of 8, so the shift depends only on the column's slot j = c mod 8 in the tile:
// c(c+1)/2 mod 4 = {0,1,3,2,2,3,1,0}[c mod 8], c(c+1)/2 mod 2 = {0,1,1,0}[c mod 4].
template <int VEC> __host__ __device__ constexpr int sp_shift(int j) {
    return VEC == 4 ? ((j * (j + 1) / 2) & 3) : ((j * (j + 1) / 2) & 1);
}
struct SpParams {
    const void* S;
    size_t n, total;
    const void* x;
    void* part_n;            // [max_splits][n]
    void* part_t;            // [n_chunks][n]
    const SpUnit* units;     // heaviest first
    int n_units;
    int* next_unit;          // work-queue head: CTAs take units in order with atomicAdd; the finalize kernel resets it to 0
    int tail;                // trailing elements of the array left to the finalize kernel
};

// XRES: all of x lives in shared memory for the whole kernel (n * sizeof(T) <= 48 KB, one ring stage fewer); otherwise each
// work unit stages the x slice of its columns and re-reads the x of its rows from L2.
template <typename T, bool XRES, int H>
__global__ void __launch_bounds__(SPS_CONSUMERS * H + 32, 1) spmv_stream_kernel(const SpParams p) {
    constexpr int SPS_STAGES = XRES ? 5 : 6;
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr int TR = SPS_CONSUMERS * VEC;          // rows per chunk: 1024 (f32) / 512 (f64)
    constexpr int SLOT = TR + VEC;                   // staged elements per column: the aligned window may start VEC-1 early
    constexpr int STAGE_ELEMS = SPS_TC * SLOT;
    constexpr int CONS = SPS_CONSUMERS * H;          // consumer threads
    constexpr int NW = SPS_WARPS * H;                // consumer warps; warp NW is the producer
    constexpr int RPT = TR / CONS;                   // rows per thread in the row pass: thread t owns rows t + CONS * i
    constexpr int HR = TR / H;                       // rows per warp in the column pass: warp w takes rows [(w / 8) * HR, +HR) of column w % 8
    constexpr int KT = HR / 32;                      // rows per lane in the column pass

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* stage_base = reinterpret_cast<T*>(smem_raw);
    T* xs = stage_base + (size_t)SPS_STAGES * STAGE_ELEMS;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(xs + (XRES ? SPS_XRES_BYTES / sizeof(T) : (size_t)SPS_MAX_UNIT_COLS));
    uint64_t* empty_bar = full_bar + SPS_STAGES;
    int4* stage_unit = reinterpret_cast<int4*>(empty_bar + SPS_STAGES);      // descriptor of the unit the tile in each stage belongs to (16-byte aligned); chunk < 0: no more work

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const T* __restrict__ S = reinterpret_cast<const T*>(p.S);
    const size_t n = p.n;
    if (tid == 0) {
        for (int s = 0; s < SPS_STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], NW); }
        ptx::mbar_fence_init();
    }
    tbd::pdl_entry();          // barrier set-up overlaps the tail of the previous kernel; no global access before this point
    __syncthreads();
    const size_t last_lim = n - (size_t)p.tail;      // exclusive row limit of the last column
    if (XRES && warp < NW) {
        const T* __restrict__ xg = reinterpret_cast<const T*>(p.x);
        for (size_t i = tid; i < n; i += CONS) xs[i] = xg[i];
        ptx::named_bar_sync(1, CONS);
    }
    const int cw = warp % SPS_WARPS, ch = warp / SPS_WARPS;      // column-pass role: column slot and row half of this warp
    const int hb = ch * HR;                                      // first chunk-local row of this warp's half

    if (warp == NW) {
        // ===================== producer warp =====================
        const uint64_t pol = ptx::policy_evict_first();
        int s = 0;
        uint32_t phase = 0;
        // Units are taken from a global queue (atomicAdd), heaviest first: the triangle makes units unequal (tiles cut by the
        // diagonal cost ~3x a dense tile) and a static split left SMs idle 36 % of the kernel (ncu r02: sm__cycles_active avg
        // 143 k vs max 216 k at n = 16384).  The unit id travels to the consumers with the tile (stage_unit, published by the
        // arrive on full_bar); results do not depend on which CTA ran a unit: partials are indexed by (chunk, split) only.
        // The queue is read one unit ahead (the atomic for unit i+1 is in flight while unit i streams; reading further ahead would
        // let a CTA sit on units that idle CTAs could take at the end of the kernel).  The descriptor load at the start of a unit is
        // an L2 round trip on the producer only: the ring holds several tiles of slack.
        int pend = 0;                                     // lane 0: result of the atomic in flight
        if (lane == 0) pend = atomicAdd(p.next_unit, 1);
        for (;;) {
            const int u_cur = __shfl_sync(0xffffffffu, pend, 0);
            if (u_cur >= p.n_units) {
                ptx::mbar_wait(&empty_bar[s], phase ^ 1);
                if (lane == 0) { stage_unit[s] = make_int4(-1, 0, 0, 0); ptx::mbar_arrive(&full_bar[s]); }
                break;
            }
            const SpUnit un = p.units[u_cur];
            if (lane == 0) pend = atomicAdd(p.next_unit, 1);
            const size_t row0 = (size_t)un.chunk * TR;
            for (int t = un.tile0; t < un.tile1; ++t) {
                const size_t c = (size_t)t * SPS_TC + lane;
                uint32_t bytes = 0;
                const T* src = nullptr;
                if (lane < SPS_TC && c < n) {
                    const size_t rlim = c == n - 1 ? last_lim : c + 1;
                    if (rlim > row0) {
                        const size_t len = rlim - row0 < (size_t)TR ? rlim - row0 : (size_t)TR;
                        const size_t e = c * (c + 1) / 2 + row0;
                        const size_t sh = e % VEC;
                        src = S + (e - sh);
                        bytes = (uint32_t)(((sh + len + VEC - 1) / VEC) * VEC * sizeof(T));
                    }
                }
                const uint32_t tile_bytes = __reduce_add_sync(0xffffffffu, bytes);
                ptx::mbar_wait(&empty_bar[s], phase ^ 1);
                if (lane == 0) { stage_unit[s] = make_int4(un.chunk, un.split, un.tile0, un.tile1); ptx::mbar_expect_tx(&full_bar[s], tile_bytes); }
                __syncwarp();
                if (bytes) ptx::bulk_g2s(stage_base + (size_t)s * STAGE_ELEMS + (size_t)lane * SLOT, src, bytes, &full_bar[s], pol);
                if (++s == SPS_STAGES) { s = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================== consumer warps =====================
        const T* __restrict__ x = reinterpret_cast<const T*>(p.x);
        T* __restrict__ part_n = reinterpret_cast<T*>(p.part_n);
        T* __restrict__ part_t = reinterpret_cast<T*>(p.part_t);
        int s = 0;
        uint32_t phase = 0;
        for (;;) {
            // the first tile of the next unit (or the end mark) is already on its way: its stage names the unit
            ptx::mbar_wait(&full_bar[s], phase);
            const int4 ud = stage_unit[s];
            if (ud.x < 0) break;
            const SpUnit un{ud.x, ud.y, ud.z, ud.w};
            const size_t row0 = (size_t)un.chunk * TR;
            const size_t c0 = (size_t)un.tile0 * SPS_TC;
            const size_t c1 = (size_t)un.tile1 * SPS_TC < n ? (size_t)un.tile1 * SPS_TC : n;
            const int ucols = (int)(c1 - c0);
            if (!XRES) {
                // stage x[c] of this unit's columns (the previous unit's readers are done: barrier first)
                ptx::named_bar_sync(1, CONS);
                for (int i = tid; i < ucols; i += CONS) xs[i] = x[c0 + i];
                ptx::named_bar_sync(1, CONS);
            }
            const T* xcol = XRES ? xs + c0 : xs;      // x of the unit's columns, indexed from the unit's first column
            // column pass: x of the rows this lane strides over (rows hb + lane + 32 k of the chunk)
            T xr[KT];
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                const size_t r = row0 + hb + lane + 32 * k;
                xr[k] = r < n ? (XRES ? xs[r] : x[r]) : T(0);
            }
            // row pass: thread t owns rows t + CONS * i of the chunk (conflict-free scalar LDS)
            T acc[RPT];
#pragma unroll
            for (int i = 0; i < RPT; ++i) acc[i] = T(0);

            int cl = 0;
            for (int t = un.tile0; t < un.tile1; ++t, cl += SPS_TC) {
                const size_t ct = (size_t)t * SPS_TC;
                const int ncols = (int)(n - ct < (size_t)SPS_TC ? n - ct : (size_t)SPS_TC);
                ptx::mbar_wait(&full_bar[s], phase);
                const T* tile = stage_base + (size_t)s * STAGE_ELEMS;
                const long long d0 = (long long)ct - (long long)row0;           // local row of column ct's diagonal element
                if (d0 >= (long long)TR && ct + SPS_TC < n) {
                    // right of the chunk's diagonal block and without the last column: every staged row counts, no masks
#pragma unroll
                    for (int j = 0; j < SPS_TC; ++j) {
                        const T* col = tile + (size_t)j * SLOT + sp_shift<VEC>(j);
                        const T xv = xcol[cl + j];
#pragma unroll
                        for (int i = 0; i < RPT; ++i) acc[i] += col[tid + CONS * i] * xv;
                    }
                    const T* col = tile + (size_t)cw * SLOT + sp_shift<VEC>(cw) + hb;
                    T sum0 = T(0), sum1 = T(0), sum2 = T(0), sum3 = T(0);
#pragma unroll
                    for (int k = 0; k < KT; k += 4) {
                        sum0 += col[lane + 32 * k] * xr[k];
                        sum1 += col[lane + 32 * (k + 1)] * xr[k + 1];
                        sum2 += col[lane + 32 * (k + 2)] * xr[k + 2];
                        sum3 += col[lane + 32 * (k + 3)] * xr[k + 3];
                    }
                    const T sum = tbd::warp_sum((sum0 + sum1) + (sum2 + sum3));
                    if (lane == 0) part_t[((size_t)un.chunk * H + ch) * n + ct + cw] = sum;
                } else {
                    // per-column limits in chunk-local rows: rows < lim_t are staged (column pass), rows < lim_n also skip the
                    // diagonal (row pass); right of the chunk's diagonal block both are >= TR and mask nothing
    #pragma unroll
                    for (int j = 0; j < SPS_TC; ++j) {
                        if (j < ncols) {
                            const size_t c = ct + j;
                            long long lim_t = d0 + j + 1;
                            if (c == n - 1) lim_t = (long long)last_lim - (long long)row0;
                            long long lim_n64 = lim_t < d0 + j ? lim_t : d0 + j;
                            const int lim_n = lim_n64 > (long long)TR ? TR : (int)lim_n64;
                            const T* col = tile + (size_t)j * SLOT + sp_shift<VEC>(j);
                            const T xv = xcol[cl + j];
    #pragma unroll
                            for (int i = 0; i < RPT; ++i)
                                if (warp * 32 + CONS * i < lim_n) {       // warp-uniform skip of rows below the diagonal
                                    if (tid + CONS * i < lim_n) acc[i] += col[tid + CONS * i] * xv;
                                }
                        }
                    }
                    if (cw < ncols) {
                        const size_t c = ct + cw;
                        long long lim_t = d0 + cw + 1;
                        if (c == n - 1) lim_t = (long long)last_lim - (long long)row0;
                        const T* col = tile + (size_t)cw * SLOT + sp_shift<VEC>(cw) + hb;
                        T sum0 = T(0), sum1 = T(0), sum2 = T(0), sum3 = T(0);
                        const int lim = (lim_t > (long long)TR ? TR : (int)lim_t) - hb;      // rows of this warp's half that are staged
    #pragma unroll
                        for (int k = 0; k < KT; k += 4) {
                            if (32 * k >= lim) break;                 // warp-uniform: nothing staged below the diagonal
                            if (lane + 32 * k < lim) sum0 += col[lane + 32 * k] * xr[k];
                            if (lane + 32 * (k + 1) < lim) sum1 += col[lane + 32 * (k + 1)] * xr[k + 1];
                            if (lane + 32 * (k + 2) < lim) sum2 += col[lane + 32 * (k + 2)] * xr[k + 2];
                            if (lane + 32 * (k + 3) < lim) sum3 += col[lane + 32 * (k + 3)] * xr[k + 3];
                        }
                        const T sum = tbd::warp_sum((sum0 + sum1) + (sum2 + sum3));
                        if (lane == 0) part_t[((size_t)un.chunk * H + ch) * n + c] = sum;
                    }
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&empty_bar[s]);
                if (++s == SPS_STAGES) { s = 0; phase ^= 1; }
            }
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const size_t r = row0 + tid + CONS * i;
                if (r < n) part_n[(size_t)un.split * n + r] = acc[i];
            }
        }
    }
}

// y[i] = alpha * ( sum_{s < splits(chunk(i))} part_n[s][i] + sum_{R <= chunk(i)} part_t[R][i] + tail terms ) + beta * y[i]
// A CTA of 256 threads finishes 32 outputs: thread (w, lane) adds partials w, w + 8, ... of output lane, then the 8
// sub-sums are added in a fixed order.
template <typename T>
__global__ void __launch_bounds__(256) spmv_stream_finalize(const T* __restrict__ part_n, const T* __restrict__ part_t, const T* __restrict__ S,
                                                            const T* __restrict__ x, size_t n, size_t total, int tail, const int* __restrict__ chunk_splits,
                                                            int halves, T alpha, T beta, T* y, int* next_unit) {
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr size_t TR = (size_t)SPS_CONSUMERS * VEC;
    __shared__ T red[8][33];
    tbd::pdl_entry();
    if (blockIdx.x == 0 && threadIdx.x == 0) *next_unit = 0;          // the streaming kernel's work queue, for its next launch
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const size_t i = blockIdx.x * (size_t)32 + lane;
    T v = T(0);
    if (i < n) {
        const size_t chunk = i / TR;
        const size_t splits = (size_t)chunk_splits[chunk];
        // four independent loads in flight per thread (the partials are L2-resident; a serial chain would pay one L2 round trip each)
        size_t sidx = w;
        T v0 = T(0), v1 = T(0), v2 = T(0), v3 = T(0);
        for (; sidx + 24 < splits; sidx += 32) {
            v0 += part_n[sidx * n + i];
            v1 += part_n[(sidx + 8) * n + i];
            v2 += part_n[(sidx + 16) * n + i];
            v3 += part_n[(sidx + 24) * n + i];
        }
        for (; sidx < splits; sidx += 8) v0 += part_n[sidx * n + i];
        const size_t n_t = (chunk + 1) * (size_t)halves;
        size_t r = w;
        for (; r + 24 < n_t; r += 32) {
            v0 += part_t[r * n + i];
            v1 += part_t[(r + 8) * n + i];
            v2 += part_t[(r + 16) * n + i];
            v3 += part_t[(r + 24) * n + i];
        }
        for (; r < n_t; r += 8) v1 += part_t[r * n + i];
        v = (v0 + v1) + (v2 + v3);
    }
    red[w][lane] = v;
    __syncthreads();
    if (w == 0 && i < n) {
        T tot = red[0][lane];
#pragma unroll
        for (int k = 1; k < 8; ++k) tot += red[k][lane];
        if (tail > 0) {
            const size_t lo = n - (size_t)tail;             // rows [lo, n) of the last column
            const T* lastcol = S + (total - n);
            if (i >= lo && i + 1 < n) tot += lastcol[i] * x[n - 1];
            if (i + 1 == n) {
                T t = T(0);
                for (size_t r = lo; r < n; ++r) t += lastcol[r] * x[r];
                tot += t;
            }
        }
        T r = alpha * tot;
        if (beta != T(0)) r += beta * y[i];
        y[i] = r;
    }
}

struct SpPlan {
    SpUnit* units_dev = nullptr;
    int* next_unit_dev = nullptr;
    int* chunk_splits_dev = nullptr;      // [n_chunks]: units (= row-pass partials) per row chunk
    int grid = 0, n_units = 0, max_splits = 0, n_chunks = 0;
};

template <typename T> static const SpPlan& spmv_plan(size_t n) {
    static std::map<size_t, SpPlan> cache;
    auto it = cache.find(n);
    if (it != cache.end()) return it->second;
    constexpr size_t VEC = 16 / sizeof(T);
    constexpr size_t TR = (size_t)SPS_CONSUMERS * VEC;
    const size_t n_tiles = (n + SPS_TC - 1) / SPS_TC, n_chunks = (n + TR - 1) / TR;
    const size_t n_cta = (size_t)ctx().sm_count;
    // Consumer cost, not bytes, is what a tile takes: a tile cut by the diagonal (or holding the last column) runs the masked path,
    // measured ~3x the instructions of a dense tile (ncu, profiles/r01_spmv_stream_full.md).
    auto tile_weight = [&](size_t R, size_t t) {
        const size_t ct = t * SPS_TC, row0 = R * TR;
        return (ct >= row0 + TR && ct + SPS_TC < n) ? 1.0 : 3.0;
    };
    double w_total = 0.0;
    for (size_t R = 0; R < n_chunks; ++R)
        for (size_t t = R * (TR / SPS_TC); t < n_tiles; ++t) w_total += tile_weight(R, t);
    // Guided unit sizes: the first 3/4 of every chunk's weight is cut into units of 1/4 of a CTA's share (few units: each costs a
    // row-pass partial and a pipeline hand-over), the last quarter into units of 1/16 of a share, and the queue serves the big
    // ones first - the tail of the kernel is then at most 1/16 of a share (~6 %) long.
    const double share = w_total / (double)n_cta;
    const double w_big = std::max(share / 4.0, 8.0), w_small = std::max(share / 16.0, 4.0);       // >= 4 tiles (128 KB) per unit
    std::vector<SpUnit> units;
    std::vector<double> weight;
    std::vector<int> chunk_splits(n_chunks, 0);
    size_t max_splits = 0;
    for (size_t R = 0; R < n_chunks; ++R) {
        const size_t t_start = R * (TR / SPS_TC);
        double w_chunk = 0.0;
        for (size_t t = t_start; t < n_tiles; ++t) w_chunk += tile_weight(R, t);
        size_t split = 0, t0 = t_start;
        double done = 0.0;
        while (t0 < n_tiles) {
            const double target = done < 0.75 * w_chunk ? w_big : w_small;
            double w = 0.0;
            size_t t1 = t0;
            while (t1 < n_tiles && t1 - t0 < (size_t)(SPS_MAX_UNIT_COLS / SPS_TC) && (w < target || t1 == t0)) { w += tile_weight(R, t1); ++t1; }
            units.push_back(SpUnit{(int)R, (int)split, (int)t0, (int)t1});
            weight.push_back(w);
            done += w;
            t0 = t1;
            ++split;
        }
        chunk_splits[R] = (int)split;
        max_splits = std::max(max_splits, split);
    }
    const int grid = (int)std::min<size_t>(n_cta, units.size());
    // heaviest first (stable: equal weights keep the chunk-major order, which streams the array front to back)
    std::vector<size_t> order(units.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return weight[a] > weight[b]; });
    std::vector<SpUnit> sorted(units.size());
    for (size_t i = 0; i < order.size(); ++i) sorted[i] = units[order[i]];
    SpPlan pl;
    TB_CUDA(cudaMalloc(&pl.units_dev, sorted.size() * sizeof(SpUnit)));
    TB_CUDA(cudaMalloc(&pl.next_unit_dev, sizeof(int)));
    TB_CUDA(cudaMemcpy(pl.units_dev, sorted.data(), sorted.size() * sizeof(SpUnit), cudaMemcpyHostToDevice));
    TB_CUDA(cudaMemset(pl.next_unit_dev, 0, sizeof(int)));
    TB_CUDA(cudaMalloc(&pl.chunk_splits_dev, chunk_splits.size() * sizeof(int)));
    TB_CUDA(cudaMemcpy(pl.chunk_splits_dev, chunk_splits.data(), chunk_splits.size() * sizeof(int), cudaMemcpyHostToDevice));
    pl.n_units = (int)sorted.size();
    pl.grid = grid; pl.max_splits = (int)max_splits; pl.n_chunks = (int)n_chunks;
    return cache.emplace(n, pl).first->second;
}

template <typename T> static bool spmv_stream_eligible(const T* S, size_t n) {
    if (ctx().gemv_mode == 1) return false;                          // tb_set_gemv_path(1): generic kernels only
    return (reinterpret_cast<uintptr_t>(S) & 15) == 0 && n >= 512;
}

template <typename T> static void spmv_stream(size_t n, T alpha, const T* S, const T* x, T beta, T* y) {
    Context& c = ctx();
    constexpr size_t VEC = 16 / sizeof(T);
    constexpr size_t TR = (size_t)SPS_CONSUMERS * VEC;
    const SpPlan& pl = spmv_plan<T>(n);
    const size_t total = n * (n + 1) / 2;
    size_t bytes_n = (size_t)pl.max_splits * n * sizeof(T);
    bytes_n = (bytes_n + 255) & ~size_t(255);
    const int halves = g_spmv_halves;
    const size_t bytes_t = (size_t)pl.n_chunks * halves * n * sizeof(T);
    char* sc = reinterpret_cast<char*>(scratch(bytes_n + bytes_t));
    SpParams p;
    p.S = S; p.n = n; p.total = total; p.x = x;
    p.part_n = sc; p.part_t = sc + bytes_n;
    p.units = pl.units_dev; p.n_units = pl.n_units; p.next_unit = pl.next_unit_dev;
    p.tail = (int)(total % VEC);
    const bool xres = n * sizeof(T) <= SPS_XRES_BYTES;
    const size_t stages = xres ? 5 : 6;
    const size_t x_elems = xres ? SPS_XRES_BYTES / sizeof(T) : (size_t)SPS_MAX_UNIT_COLS;
    const size_t smem = stages * SPS_TC * (TR + VEC) * sizeof(T) + x_elems * sizeof(T) + 2 * stages * sizeof(uint64_t) + (stages + 1) * 16 + 128;
    auto launch = [&](auto kern, int threads) {
        static std::map<const void*, bool> attr_set;           // one flag per kernel instantiation
        const void* key = reinterpret_cast<const void*>(kern);
        if (!attr_set[key]) {
            TB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set[key] = true;
        }
        launch_pdl(kern, dim3(pl.grid), dim3(threads), smem, c.stream, p);
    };
    if (halves == 2) {
        if (xres) launch(spmv_stream_kernel<T, true, 2>, SPS_CONSUMERS * 2 + 32);
        else launch(spmv_stream_kernel<T, false, 2>, SPS_CONSUMERS * 2 + 32);
    } else {
        if (xres) launch(spmv_stream_kernel<T, true, 1>, SPS_THREADS);
        else launch(spmv_stream_kernel<T, false, 1>, SPS_THREADS);
    }
    TB_LAUNCH_CHECK();
    launch_pdl(spmv_stream_finalize<T>, dim3((unsigned)((n + 31) / 32)), dim3(256), 0, c.stream, reinterpret_cast<const T*>(p.part_n), reinterpret_cast<const T*>(p.part_t), S, x,
               n, total, p.tail, (const int*)pl.chunk_splits_dev, halves, alpha, beta, y, pl.next_unit_dev);
    TB_LAUNCH_CHECK();
}

template <typename T> static void api_transform_sp(size_t n, T alpha, tb_view mat, tb_view x, T beta, tb_view y) {
    require_init();
    TB_REQUIRE(mat.len == n * (n + 1) / 2, "transform_sp: mat.len != n(n+1)/2");     // f64lapack.rs:151
    TB_REQUIRE(x.len == n && y.len == n, "transform_sp: vector length mismatch");     // :153-154
    const T* S = rptr<T>(mat);
    const T* px = rptr<T>(x);
    T* py = wptr<T>(y, beta == T(0));
    if (n == 0) return;
    if (spmv_stream_eligible<T>(S, n)) { spmv_stream<T>(n, alpha, S, px, beta, py); return; }
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
int tb_set_spmv_warps(int warps) {
    return api([&] {
        TB_REQUIRE(warps == 8 || warps == 16, "spmv consumer warps: 8 or 16");
        g_spmv_halves = warps / 8;
    });
}
int tb_transform_sp_f32(size_t n, float a, tb_view m, tb_view x, float b, tb_view y) { return api_defer({m, x, y}, {y}, [=] { api_transform_sp<float>(n, a, m, x, b, y); }); }
int tb_transform_sp_f64(size_t n, double a, tb_view m, tb_view x, double b, tb_view y) { return api_defer({m, x, y}, {y}, [=] { api_transform_sp<double>(n, a, m, x, b, y); }); }
}
