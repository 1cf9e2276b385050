// Vector programs: the small vector commands of one solver iteration, run as ONE launch.
//
// Between two streaming matvecs the unmodified solver issues a dozen level-1 / vector-operator calls on vectors of
// <= a few hundred thousand elements (SelfDualEmbed::op / trans_op, update_vecs, criteria_conv: solver.rs:109-157,
// 526-612; LinAlg trait: linalg.rs:22-67).  As separate kernels each costs a launch (~3 us of host time, ~4-5 us of
// stream time) for ~1 us of work, and that floor (~0.2 ms per iteration) is what limits every configuration whose
// matrices are small or sharded 8 ways.  Here the launch sites record a micro-op instead (vp_push_*); the recorded
// program is submitted as one kernel the moment anything else touches the stream (Context::stream flushes on use) or
// a host-visible result is needed - in which case the LAST micro-op posts it into the mapped host box.
//
// Execution: ONE thread-block cluster of 8 CTAs x 1024 threads walks the op list in order, every thread owning the
// indices congruent to its cluster-wide id.  A hardware cluster barrier (barrier.cluster release/acquire: orders global
// memory inside the cluster, ~0.3 us) is placed only where the recorder found a data hazard (RAW / WAR / WAW on
// overlapping byte ranges) since the previous one - and not even then when both accesses are element-wise from the same
// base, because the same thread owns the element in both ops.  Inputs are read with ld.global.cg (L2): another CTA may
// have produced them in this launch.  Reductions (dot products, norms) are two micro-ops: per-CTA partials in double to
// a small global slot, barrier, then a fixed-order sum of the 8 partials - deterministic, no atomics.
//
// Programs whose vectors are all <= 48K elements run on the cluster.  8 SMs cannot move longer vectors faster than a
// full-grid kernel, and a software grid barrier across all SMs measured no cheaper than a launch boundary (~3 us in a
// dependent chain; profiles/r01_vprog_summary.md) - so a program that contains a longer vector is "wide": it runs as an
// ordinary grid of up to two CTAs per SM, holds no barrier at all (the recorder cuts it at every hazard instead) and
// no reductions; what it still saves is one launch per run of independent or same-thread-dependent element-wise ops
// (finalize + the c / b vector-operator updates + axpby chains of SelfDualEmbed::op / trans_op).  Measured: small SOCP (vectors of
// 20K) 0.276 -> 0.223 ms per iteration, C2 (43K) 0.498 -> 0.483 ms; C3 (147K) unchanged by construction.
#include "common.cuh"
#include "vprog.cuh"

namespace tb {

namespace {

constexpr int VP_CLUSTER = 8;
constexpr int VP_THREADS = 1024;
constexpr int VP_MAX_OPS = 44;
constexpr int VP_SLOTS = 32;            // reduction slots of VP_MAX_CTAS doubles each
constexpr int VP_MAX_CTAS = VP_CLUSTER;
// Longest vector run on one cluster of 8 SMs.  Measured on B200 (scripts/bench_vprog_stage.py, profiles/r02_vprog_threshold.md):
// a barrier-separated stage costs 1.3 us up to 64 K elements but 3.5 us at 147 K (8 SMs pull only ~0.4 TB/s out of L2), while a
// barrier-free wide program pays a launch boundary (3-5 us with PDL) at every hazard.  Raising the threshold to 256 K takes C3's
// iteration from 24 launches to 14 and makes it SLOWER (1.447 -> 1.540 ms; C4 0.929 -> 1.027): 48 K stays.
// tb_set_vprog_max_n / TB_VPROG_MAX_N change it (diagnostics).
size_t g_vp_max_n = 49152;
#define VP_MAX_N g_vp_max_n
// Reductions (dot products, norms) are different: one pass of reads, no stores - 8 SMs take a 64 K-element dot in ~1.7 us, where
// the dedicated two-stage kernel costs a launch plus ~8 us of its own latency chain.  They stay on the cluster up to 512 K elements
// and never make a program wide.
constexpr size_t VP_MAX_N_RED = size_t(1) << 19;
constexpr size_t VP_WIDE_MAX_N = size_t(1) << 22;    // beyond this a dedicated kernel with a larger grid is the better tool

enum : uint8_t {
    VOP_FILL = 1,        // y[i] = a
    VOP_SCALE,           // y[i] = a * y[i]
    VOP_COPY,            // y[i] = x[i]
    VOP_AXPBY,           // y[i] = a * x[i] (+ y[i] | + b * y[i])                   flags: beta mode 0 / 1 / 2
    VOP_ADDS,            // y[i] += a
    VOP_DIAG,            // y[i] = a * (d[i] * x[i]) (+ y[i] | + b * y[i])          p2 = d
    VOP_FINALIZE,        // y[i] = a * sum_j part[j * ld + i] (+ b * y[i])          x = part, aux = nparts, aux2 = ld
    VOP_AXS,             // y[i] = a * (x[i] * s[0]) (+ b * y[i])                   p2 = s (device scalar): an n x 1 operator's op
    VOP_SET1,            // y[0] = a
    VOP_PART_DOT,        // slot[rank] = sum x[i] * d[i] over this CTA's share      p2 = d, aux = slot
    VOP_PART_SUMSQ,      // slot[rank] = sum x[i*inc]^2                             aux = slot, aux2 = inc
    VOP_PART_ABSSUM,     // slot[rank] = sum |x[i*inc]|
    VOP_COMBINE_Y,       // y[0] = a * sum_r slot[r] (+ b * y[0])                   aux = slot
    VOP_COMBINE_PF,      // box[4 + aux2] = sum_r slot[r]  (scalar prefetch, prefetch.cu)       aux = slot, aux2 = prefetch slot
    VOP_PF_SEQ,          // box[8] = aux2 after a system fence: the prefetched values of this program are complete
    VOP_COMBINE_BOX,     // box <- sum_r slot[r]                                    aux = slot, seq in aux2
    VOP_FETCH_BOX,       // box <- x[0]                                             seq in aux2
};

struct MicroOp {
    uint8_t code, f64, mode, barrier;     // barrier: cluster barrier BEFORE this op
    uint32_t aux;
    const void* x;
    const void* p2;
    void* y;
    unsigned long long n, aux2;
    double a, b;
};
static_assert(sizeof(MicroOp) == 64, "MicroOp is 64 bytes");

struct Program {
    int n_ops;
    int pad;
    double* slots;
    double* box;
    MicroOp ops[VP_MAX_OPS];
};
static_assert(sizeof(Program) <= 4000, "the program travels as a kernel parameter");

__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

template <typename T> __device__ __forceinline__ T ldg_cg(const T* p) { return __ldcg(p); }

// Plain grid-stride loops: nvcc unrolls them 4x with the loads hoisted, which is all the memory-level parallelism these
// latency-bound passes need.
template <typename T> __device__ __forceinline__ void run_elementwise(const MicroOp& op, size_t gtid, size_t gstride) {
    // no __restrict__ / read-only hints: an input may have been written by an earlier op of this same launch
    const T* x = reinterpret_cast<const T*>(op.x);
    const T* d = reinterpret_cast<const T*>(op.p2);
    T* y = reinterpret_cast<T*>(op.y);
    const size_t n = op.n;
    const T a = (T)op.a, b = (T)op.b;
    switch (op.code) {
        case VOP_FILL:
            for (size_t i = gtid; i < n; i += gstride) y[i] = a;
            break;
        case VOP_SCALE:
            for (size_t i = gtid; i < n; i += gstride) y[i] = a * ldg_cg(y + i);
            break;
        case VOP_COPY:
            for (size_t i = gtid; i < n; i += gstride) y[i] = ldg_cg(x + i);
            break;
        case VOP_ADDS:
            for (size_t i = gtid; i < n; i += gstride) y[i] = ldg_cg(y + i) + a;
            break;
        case VOP_AXPBY:
            if (op.mode == 0) { for (size_t i = gtid; i < n; i += gstride) y[i] = a * ldg_cg(x + i); }
            else if (op.mode == 1) { for (size_t i = gtid; i < n; i += gstride) y[i] = a * ldg_cg(x + i) + ldg_cg(y + i); }
            else { for (size_t i = gtid; i < n; i += gstride) y[i] = a * ldg_cg(x + i) + b * ldg_cg(y + i); }
            break;
        case VOP_DIAG:
            if (op.mode == 0) { for (size_t i = gtid; i < n; i += gstride) y[i] = a * (ldg_cg(d + i) * ldg_cg(x + i)); }
            else if (op.mode == 1) { for (size_t i = gtid; i < n; i += gstride) y[i] = a * (ldg_cg(d + i) * ldg_cg(x + i)) + ldg_cg(y + i); }
            else { for (size_t i = gtid; i < n; i += gstride) y[i] = a * (ldg_cg(d + i) * ldg_cg(x + i)) + b * ldg_cg(y + i); }
            break;
        case VOP_FINALIZE: {
            const int nparts = (int)op.aux;
            const size_t ld = op.aux2;
            for (size_t i = gtid; i < n; i += gstride) {
                T s = T(0);
                for (int j = 0; j < nparts; ++j) s += ldg_cg(x + (size_t)j * ld + i);
                T r = a * s;
                if (op.mode != 0) r += b * ldg_cg(y + i);
                y[i] = r;
            }
            break;
        }
        case VOP_AXS: {
            const T sc = op.aux ? (T)__longlong_as_double((long long)op.aux2) : ldg_cg(d);      // aux = 1: the scalar travels by value
            if (op.mode == 0) { for (size_t i = gtid; i < n; i += gstride) y[i] = a * (ldg_cg(x + i) * sc); }
            else { for (size_t i = gtid; i < n; i += gstride) y[i] = a * (ldg_cg(x + i) * sc) + b * ldg_cg(y + i); }
            break;
        }
        default: break;
    }
}

template <typename T> __device__ __forceinline__ double partial_sum(const MicroOp& op, size_t gtid, size_t gstride) {
    const T* x = reinterpret_cast<const T*>(op.x);
    const T* d = reinterpret_cast<const T*>(op.p2);
    const size_t n = op.n;
    double acc = 0.0;
    if (op.code == VOP_PART_DOT) {
        for (size_t i = gtid; i < n; i += gstride) acc += (double)ldg_cg(x + i) * (double)ldg_cg(d + i);
    } else if (op.code == VOP_PART_SUMSQ) {
        const size_t inc = op.aux2;
        for (size_t i = gtid; i < n; i += gstride) { const double v = (double)ldg_cg(x + i * inc); acc += v * v; }
    } else {
        const size_t inc = op.aux2;
        for (size_t i = gtid; i < n; i += gstride) acc += fabs((double)ldg_cg(x + i * inc));
    }
    return acc;
}

// sum of the G per-CTA partials of a slot in a fixed order (lane-strided, then a shuffle tree); call from a full warp
__device__ __forceinline__ double combine_slot(const double* slot, int g) {
    const int lane = threadIdx.x & 31;
    double s = 0.0;
    for (int r = lane; r < g; r += 32) s += __ldcg(slot + r);
    return tbd::warp_sum(s);
}

// CLUSTER: one cluster of VP_CLUSTER CTAs with hardware barriers.  Otherwise ("wide" programs: some vector is longer than
// VP_MAX_N) an ordinary grid of up to two CTAs per SM and NO barriers: the recorder cuts a wide program at every hazard.
template <bool CLUSTER>
__device__ __forceinline__ void run_program(const Program& prog) {
    __shared__ double red[32];
    tbd::pdl_entry();
    const unsigned rank = CLUSTER ? cluster_ctarank() : blockIdx.x;
    const int g = CLUSTER ? VP_CLUSTER : (int)gridDim.x;
    const size_t gtid = (size_t)rank * VP_THREADS + threadIdx.x;
    const size_t gstride = (size_t)g * VP_THREADS;
    for (int k = 0; k < prog.n_ops; ++k) {
        const MicroOp& op = prog.ops[k];
        if (CLUSTER && op.barrier) cluster_barrier();
        switch (op.code) {
            case VOP_PART_DOT:
            case VOP_PART_SUMSQ:
            case VOP_PART_ABSSUM: {
                const double acc = op.f64 ? partial_sum<double>(op, gtid, gstride) : partial_sum<float>(op, gtid, gstride);
                const double tot = tbd::block_sum(acc, red);
                if (threadIdx.x == 0) prog.slots[(size_t)op.aux * VP_MAX_CTAS + rank] = tot;
                break;
            }
            case VOP_COMBINE_Y:
                if (rank == 0 && threadIdx.x < 32) {
                    const double s = combine_slot(prog.slots + (size_t)op.aux * VP_MAX_CTAS, g);
                    if (threadIdx.x == 0) {
                        if (op.f64) {
                            double* y = reinterpret_cast<double*>(op.y);
                            double r = op.a * s;
                            if (op.mode != 0) r += op.b * __ldcg(y);
                            y[0] = r;
                        } else {
                            float* y = reinterpret_cast<float*>(op.y);
                            float r = (float)op.a * (float)s;
                            if (op.mode != 0) r += (float)op.b * __ldcg(y);
                            y[0] = r;
                        }
                    }
                }
                break;
            case VOP_COMBINE_PF:
                if (rank == 0 && threadIdx.x < 32) {
                    const double s = combine_slot(prog.slots + (size_t)op.aux * VP_MAX_CTAS, g);
                    if (threadIdx.x == 0) *reinterpret_cast<volatile double*>(prog.box + 4 + op.aux2) = s;
                }
                break;
            case VOP_PF_SEQ:
                if (rank == 0 && threadIdx.x == 0) {
                    __threadfence_system();
                    *reinterpret_cast<volatile unsigned long long*>(prog.box + 8) = op.aux2;
                }
                break;
            case VOP_COMBINE_BOX:
                if (rank == 0 && threadIdx.x < 32) {
                    const double s = combine_slot(prog.slots + (size_t)op.aux * VP_MAX_CTAS, g);
                    if (threadIdx.x == 0) tbd::box_post(prog.box, s, op.aux2);
                }
                break;
            case VOP_FETCH_BOX:
                if (gtid == 0) {
                    const double v = op.f64 ? __ldcg(reinterpret_cast<const double*>(op.x)) : (double)__ldcg(reinterpret_cast<const float*>(op.x));
                    tbd::box_post(prog.box, v, op.aux2);
                }
                break;
            case VOP_SET1:
                if (gtid == 0) {
                    if (op.f64) reinterpret_cast<double*>(op.y)[0] = op.a;
                    else reinterpret_cast<float*>(op.y)[0] = (float)op.a;
                }
                break;
            default:
                if (op.f64) run_elementwise<double>(op, gtid, gstride);
                else run_elementwise<float>(op, gtid, gstride);
                break;
        }
    }
}

__global__ void __cluster_dims__(VP_CLUSTER, 1, 1) __launch_bounds__(VP_THREADS, 1) vprog_kernel(const __grid_constant__ Program prog) {
    run_program<true>(prog);
}
__global__ void __launch_bounds__(VP_THREADS, 2) vprog_wide_kernel(const __grid_constant__ Program prog) {
    run_program<false>(prog);
}

// ---- recorder ------------------------------------------------------------------------------------------
// tag: the op touches element i of this range from the thread that owns index i (identity mapping from `lo`); two such
// accesses with the same base are ordered by program order inside one thread and need no barrier
struct Range { uintptr_t lo, hi; bool tag; };

struct Recorder {
    Program prog{};
    std::vector<Range> reads, writes;       // since the last barrier
    int next_slot = 0;
    double* slots = nullptr;
    bool flushing = false;
    bool wide = false;            // some op is longer than VP_MAX_N: barrier-free, runs on the whole GPU
    bool has_barrier = false, has_reduction = false;
    size_t max_n = 0;
};
Recorder g_rec;

bool overlaps(const std::vector<Range>& v, Range r) {
    if (r.lo >= r.hi) return false;
    for (const Range& q : v)
        if (q.lo < r.hi && r.lo < q.hi && !(q.tag && r.tag && q.lo == r.lo)) return true;
    return false;
}

inline Range range_of(const void* p, size_t bytes) { return Range{reinterpret_cast<uintptr_t>(p), reinterpret_cast<uintptr_t>(p) + bytes, false}; }
inline Range elems_of(const void* p, size_t bytes) { return Range{reinterpret_cast<uintptr_t>(p), reinterpret_cast<uintptr_t>(p) + bytes, true}; }

// Append one micro-op; rd[] / wr[] are the byte ranges it reads / writes.
void push(MicroOp op, std::initializer_list<Range> rd, std::initializer_list<Range> wr, bool reduction = false, bool force_big = false) {
    Recorder& R = g_rec;
    if (R.prog.n_ops == VP_MAX_OPS) vp_flush();
    const bool big = !reduction && ((size_t)op.n > VP_MAX_N || force_big);
    if (R.prog.n_ops > 0) {
        // a wide program has neither barriers nor reductions (their partial slots are sized for the cluster)
        if (big && !R.wide && (R.has_barrier || R.has_reduction)) vp_flush();
        if (reduction && R.wide) vp_flush();
    }
    bool hazard = false;
    for (const Range& r : rd) hazard = hazard || overlaps(R.writes, r);
    for (const Range& w : wr) hazard = hazard || overlaps(R.writes, w) || overlaps(R.reads, w);
    if (hazard && R.prog.n_ops > 0 && (R.wide || big)) { vp_flush(); hazard = false; }      // cut instead of a barrier
    if (hazard) { R.reads.clear(); R.writes.clear(); }
    op.barrier = (hazard && R.prog.n_ops > 0) ? 1 : 0;
    if (big) R.wide = true;
    if (op.barrier) R.has_barrier = true;
    if (reduction) R.has_reduction = true;
    R.max_n = std::max<size_t>(R.max_n, (size_t)op.n);
    for (const Range& r : rd) if (r.lo < r.hi) R.reads.push_back(r);
    for (const Range& w : wr) if (w.lo < w.hi) R.writes.push_back(w);
    R.prog.ops[R.prog.n_ops++] = op;
}

int take_slot() {
    Recorder& R = g_rec;
    if (R.next_slot == VP_SLOTS) vp_flush();      // slots are recycled per program
    return R.next_slot++;
}
Range slot_range(int slot) { return range_of(g_rec.slots + (size_t)slot * VP_MAX_CTAS, VP_MAX_CTAS * sizeof(double)); }

}  // namespace

bool vp_enabled(size_t n) { return ctx().vprog && n <= VP_MAX_N; }
void vp_set_max_n(size_t n) { vp_flush(); g_vp_max_n = n; }
bool vp_enabled_red(size_t n) { return ctx().vprog && n <= VP_MAX_N_RED; }
// element-wise ops of any practical length can be recorded: beyond VP_MAX_N the program becomes a barrier-free "wide" one
bool vp_enabled_wide(size_t n) { return ctx().vprog && n <= VP_WIDE_MAX_N; }

void vp_flush() {
    Recorder& R = g_rec;
    if (R.prog.n_ops == 0 || R.flushing) return;
    R.flushing = true;
    Context& c = ctx();
    R.prog.slots = R.slots;
    R.prog.box = c.hostbox_dev;
    cudaError_t e = cudaSuccess;
    try {
        if (R.wide) {
            const int g = (int)std::max<size_t>(1, std::min<size_t>((R.max_n + VP_THREADS - 1) / VP_THREADS, (size_t)c.sm_count * 2));
            launch_pdl(vprog_wide_kernel, dim3(g), dim3(VP_THREADS), 0, c.stream.raw, R.prog);
            c.vprog_wide_launches += 1;
        } else {
            launch_pdl(vprog_kernel, dim3(VP_CLUSTER), dim3(VP_THREADS), 0, c.stream.raw, R.prog);
        }
    } catch (const Error&) {
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaErrorLaunchFailure;
    }
    if (g_timeline_on) timeline_mark(R.wide ? "vprog_wide#ops" : "vprog_cluster#ops", R.prog.n_ops);
    R.prog.n_ops = 0;
    R.next_slot = 0;
    R.wide = false; R.has_barrier = false; R.has_reduction = false; R.max_n = 0;
    R.reads.clear();
    R.writes.clear();
    R.flushing = false;
    c.launches += 1;
    c.vprog_launches += 1;
    if (e != cudaSuccess) fail(TB_ERR_CUDA, std::string("vector program launch: ") + cudaGetErrorString(e));
}

void vp_init() {
    if (g_rec.slots == nullptr) TB_CUDA(cudaMalloc(&g_rec.slots, (size_t)VP_SLOTS * VP_MAX_CTAS * sizeof(double)));
    g_rec.prog.n_ops = 0;
    g_rec.next_slot = 0;
    g_rec.reads.clear();
    g_rec.writes.clear();
}
void vp_shutdown() {
    g_rec.prog.n_ops = 0;
    if (g_rec.slots) { cudaFree(g_rec.slots); g_rec.slots = nullptr; }

}

static inline MicroOp mk(uint8_t code, int dtype) {
    MicroOp op{};
    op.code = code;
    op.f64 = dtype == TB_F64 ? 1 : 0;
    ctx().vprog_ops += 1;
    return op;
}
static inline size_t es(int dtype) { return dtype == TB_F64 ? 8 : 4; }
static inline uint8_t beta_mode(double b) { return b == 0.0 ? 0 : b == 1.0 ? 1 : 2; }

void vp_fill(int dtype, void* y, double v, size_t n) {
    MicroOp op = mk(VOP_FILL, dtype); op.y = y; op.a = v; op.n = n;
    push(op, {}, {elems_of(y, n * es(dtype))});
}
void vp_scale(int dtype, double a, void* y, size_t n) {
    MicroOp op = mk(VOP_SCALE, dtype); op.y = y; op.a = a; op.n = n;
    push(op, {elems_of(y, n * es(dtype))}, {elems_of(y, n * es(dtype))});
}
void vp_copy(int dtype, const void* x, void* y, size_t n) {
    MicroOp op = mk(VOP_COPY, dtype); op.x = x; op.y = y; op.n = n;
    push(op, {elems_of(x, n * es(dtype))}, {elems_of(y, n * es(dtype))});
}
void vp_axpby(int dtype, double a, const void* x, double b, void* y, size_t n) {
    MicroOp op = mk(VOP_AXPBY, dtype); op.x = x; op.y = y; op.a = a; op.b = b; op.n = n; op.mode = beta_mode(b);
    const Range ry = elems_of(y, n * es(dtype));
    if (op.mode == 0) push(op, {elems_of(x, n * es(dtype))}, {ry});
    else push(op, {elems_of(x, n * es(dtype)), ry}, {ry});
}
void vp_adds(int dtype, double s, void* y, size_t n) {
    MicroOp op = mk(VOP_ADDS, dtype); op.y = y; op.a = s; op.n = n;
    push(op, {elems_of(y, n * es(dtype))}, {elems_of(y, n * es(dtype))});
}
void vp_diag(int dtype, double a, const void* d, const void* x, double b, void* y, size_t n) {
    MicroOp op = mk(VOP_DIAG, dtype); op.x = x; op.p2 = d; op.y = y; op.a = a; op.b = b; op.n = n; op.mode = beta_mode(b);
    const size_t nb = n * es(dtype);
    const Range ry = elems_of(y, nb);
    if (op.mode == 0) push(op, {elems_of(x, nb), elems_of(d, nb)}, {ry});
    else push(op, {elems_of(x, nb), elems_of(d, nb), ry}, {ry});
}
void vp_finalize(int dtype, const void* part, int nparts, size_t ld, size_t len, double a, double b, void* y) {
    MicroOp op = mk(VOP_FINALIZE, dtype); op.x = part; op.y = y; op.a = a; op.b = b; op.n = len; op.aux = (uint32_t)nparts; op.aux2 = ld;
    op.mode = b == 0.0 ? 0 : 2;
    const Range rp = range_of(part, ((size_t)(nparts - 1) * ld + len) * es(dtype)), ry = elems_of(y, len * es(dtype));
    // What a finalize moves is nparts x len partials, not len: C2's 17 412 outputs sum 74 partials each - 5.2 MB, 13 us through the 8
    // SMs of the cluster (~0.4 TB/s) against ~2 us on the whole chip plus a launch boundary.  Beyond 512 K partial elements (2 MB: where the cluster time passes the cost of a launch boundary) the op
    // counts as long and goes to a wide program.  TB_VP_FINALIZE_WIDE (elements) moves the threshold (diagnostics).
    static const size_t wide_elems = [] { const char* e = std::getenv("TB_VP_FINALIZE_WIDE"); return e ? (size_t)std::atoll(e) : (size_t)524288; }();
    const bool long_op = (size_t)nparts * len > wide_elems;
    if (op.mode == 0) push(op, {rp}, {ry}, false, long_op);
    else push(op, {rp, ry}, {ry}, false, long_op);
}
void vp_axs(int dtype, double a, const void* x, const void* s, double b, void* y, size_t n) {
    MicroOp op = mk(VOP_AXS, dtype); op.x = x; op.p2 = s; op.y = y; op.a = a; op.b = b; op.n = n; op.mode = b == 0.0 ? 0 : 2;
    const Range ry = elems_of(y, n * es(dtype));
    if (op.mode == 0) push(op, {elems_of(x, n * es(dtype)), range_of(s, es(dtype))}, {ry});
    else push(op, {elems_of(x, n * es(dtype)), range_of(s, es(dtype)), ry}, {ry});
}
void vp_axs_imm(int dtype, double a, const void* x, double sval, double b, void* y, size_t n) {
    MicroOp op = mk(VOP_AXS, dtype); op.x = x; op.p2 = nullptr; op.y = y; op.a = a; op.b = b; op.n = n; op.mode = b == 0.0 ? 0 : 2;
    op.aux = 1;
    std::memcpy(&op.aux2, &sval, sizeof(double));
    const Range ry = elems_of(y, n * es(dtype));
    if (op.mode == 0) push(op, {elems_of(x, n * es(dtype))}, {ry});
    else push(op, {elems_of(x, n * es(dtype)), ry}, {ry});
}
void vp_set1(int dtype, void* y, double v) {
    MicroOp op = mk(VOP_SET1, dtype); op.y = y; op.a = v; op.n = 1;
    push(op, {}, {range_of(y, es(dtype))});
}
void vp_dot(int dtype, double a, const void* x, const void* d, size_t n, double b, void* y) {
    const int slot = take_slot();
    MicroOp p = mk(VOP_PART_DOT, dtype); p.x = x; p.p2 = d; p.n = n; p.aux = (uint32_t)slot;
    push(p, {range_of(x, n * es(dtype)), range_of(d, n * es(dtype))}, {slot_range(slot)}, true);
    MicroOp q = mk(VOP_COMBINE_Y, dtype); q.y = y; q.a = a; q.b = b; q.aux = (uint32_t)slot; q.mode = b == 0.0 ? 0 : 2;
    if (q.mode == 0) push(q, {slot_range(slot)}, {range_of(y, es(dtype))}, true);
    else push(q, {slot_range(slot), range_of(y, es(dtype))}, {range_of(y, es(dtype))}, true);
}
// sum of squares (mode 0) or of absolute values (mode 1) of x[i * inc], i < count, delivered to the host
double vp_reduce_to_host(int dtype, int mode, const void* x, size_t count, size_t inc) {
    const int slot = take_slot();
    MicroOp p = mk(mode == 0 ? VOP_PART_SUMSQ : VOP_PART_ABSSUM, dtype); p.x = x; p.n = count; p.aux = (uint32_t)slot; p.aux2 = inc;
    push(p, {range_of(x, ((count - 1) * inc + 1) * es(dtype))}, {slot_range(slot)}, true);
    const uint64_t seq = box_next();
    MicroOp q = mk(VOP_COMBINE_BOX, dtype); q.aux = (uint32_t)slot; q.aux2 = seq;
    push(q, {slot_range(slot)}, {}, true);
    pf_into_program();            // reductions predicted to be asked for next ride in this same launch (prefetch.cu)
    vp_flush();
    return box_wait(seq);
}
// scalar prefetch riding in the pending program: one reduction whose result goes to prefetch slot k of the host box
// (kind 1: sum of squares of a, 2: dot(a, b)); vp_pf_seq closes the set.  Accumulation in double, fixed order.
void vp_pf_reduce(int dtype, int kind, const void* a, const void* b, size_t n, int k) {
    const int slot = take_slot();
    MicroOp p = mk(kind == 2 ? VOP_PART_DOT : VOP_PART_SUMSQ, dtype); p.x = a; p.p2 = b; p.n = n; p.aux = (uint32_t)slot; p.aux2 = 1;
    if (kind == 2) push(p, {range_of(a, n * es(dtype)), range_of(b, n * es(dtype))}, {slot_range(slot)}, true);
    else push(p, {range_of(a, n * es(dtype))}, {slot_range(slot)}, true);
    MicroOp q = mk(VOP_COMBINE_PF, dtype); q.aux = (uint32_t)slot; q.aux2 = (unsigned long long)k; q.n = 1;
    push(q, {slot_range(slot)}, {range_of(ctx().hostbox_dev + 4 + k, sizeof(double))}, true);
}
void vp_pf_seq(unsigned long long seq) {
    MicroOp q = mk(VOP_PF_SEQ, TB_F64); q.aux2 = seq; q.n = 1;
    push(q, {}, {range_of(ctx().hostbox_dev + 8, sizeof(double))}, true);
}
double vp_fetch_to_host(int dtype, const void* x) {
    const uint64_t seq = box_next();
    MicroOp q = mk(VOP_FETCH_BOX, dtype); q.x = x; q.aux2 = seq;
    push(q, {range_of(x, es(dtype))}, {});
    pf_into_program();
    vp_flush();
    return box_wait(seq);
}

}  // namespace tb
