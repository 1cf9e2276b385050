// Shared internals of libtotsu_b200.so: the process-global context, buffer table with host/device
// coherence tracking, error plumbing and small device helpers.  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <functional>
#include <initializer_list>
#include <stdexcept>
#include <algorithm>
#include <utility>
#include <mutex>
#include <map>
#include "../../include/totsu_b200.h"

namespace tb {

struct Error {
    int code;
    std::string msg;
};

[[noreturn]] inline void fail(int code, const std::string& m) { throw Error{code, m}; }

#define TB_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            ::tb::fail(TB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +    \
                                        __FILE__ + ":" + std::to_string(__LINE__) + ")");          \
    } while (0)

#define TB_REQUIRE(cond, msg)                                                                      \
    do {                                                                                           \
        if (!(cond)) ::tb::fail(TB_ERR_ARG, std::string(msg) + " [" #cond "] (" + __FILE__ + ":" + \
                                                std::to_string(__LINE__) + ")");                   \
    } while (0)

// Sorted set of disjoint half-open element ranges.  Replaces F32CUDASlice's per-split 3-state machine
// (f32cuda_slice.rs:14-20,157-168): coherence is tracked per root buffer as "which ranges are newer on
// the host" and "which are newer on the device", so splitting a slice costs nothing.
class IntervalSet {
public:
    struct Iv { size_t a, b; };
    bool empty() const { return v_.empty(); }
    void clear() { v_.clear(); }
    void add(size_t a, size_t b) {
        if (a >= b) return;
        std::vector<Iv> out;
        out.reserve(v_.size() + 1);
        size_t i = 0;
        while (i < v_.size() && v_[i].b < a) out.push_back(v_[i++]);
        while (i < v_.size() && v_[i].a <= b) { a = std::min(a, v_[i].a); b = std::max(b, v_[i].b); ++i; }
        out.push_back({a, b});
        while (i < v_.size()) out.push_back(v_[i++]);
        v_.swap(out);
    }
    void sub(size_t a, size_t b) {
        if (a >= b || v_.empty()) return;
        std::vector<Iv> out;
        out.reserve(v_.size() + 1);
        for (const Iv& iv : v_) {
            if (iv.b <= a || iv.a >= b) { out.push_back(iv); continue; }
            if (iv.a < a) out.push_back({iv.a, a});
            if (iv.b > b) out.push_back({b, iv.b});
        }
        v_.swap(out);
    }
    bool intersects(size_t a, size_t b) const {
        for (const Iv& iv : v_) if (iv.a < b && a < iv.b) return true;
        return false;
    }
    template <typename F> void for_each_in(size_t a, size_t b, F f) const {
        for (const Iv& iv : v_) {
            size_t lo = std::max(a, iv.a), hi = std::min(b, iv.b);
            if (lo < hi) f(lo, hi);
        }
    }
    const std::vector<Iv>& ivs() const { return v_; }
private:
    std::vector<Iv> v_;
};

struct Buffer {
    bool alive = false;
    int dtype = TB_F32;
    size_t len = 0;
    size_t esize = 4;
    char* dev = nullptr;
    char* host = nullptr;      // caller-owned mirror, may be null (device-only)
    bool host_mut = false;
    int refs = 1;              // live wrappers (root + sub-slices) when the binding counts them: tb_buf_retain/release
    uint64_t gen = 0;          // creation order, newest wins an address lookup
    int small_slot = -1;       // >= 0: carved from the small-buffer slab, no cudaMalloc/cudaFree
    IntervalSet host_newer;    // host copy is newer than device
    IntervalSet dev_newer;     // device copy is newer than host
};

struct DenseOp;
struct ConeSet;

// ---- lazy op/trans_op pairing ----------------------------------------------------------------------------
// The solver calls A.trans_op and A.op back to back on independent vectors (SelfDualEmbed::op / trans_op,
// solver.rs:128-131,150-153; criteria_conv, solver.rs:595-598), each a full read of A.  Behind the unmodified
// trait surface the two calls arrive separately, so tb_denseop_apply does not launch at once: it parks the call
// as the head of a small command queue.  Deferrable calls that follow (the c/b vector products in between) are
// appended; when the opposite-direction apply on the same operator arrives and the data dependencies allow it to
// be hoisted up to the head (or the head to be sunk down to it), both run as ONE streaming pass over A
// (stream_kernel<T, true, true>).  Any call with a host-visible result, or one that does not fit the pattern,
// drains the queue in program order first.  Results are bit-identical to the unfused sequence.
struct Cmd {
    tb_view reads[4];
    tb_view writes[2];
    int n_reads = 0, n_writes = 0;
    std::function<void()> run;
    // dense-apply metadata (is_dense): enough to re-issue it as half of a pair
    bool is_dense = false;
    tb_handle op = 0;
    int trans = 0, dtype = 0;
    double alpha = 0.0, beta = 0.0;
    tb_view x{0, 0, 0}, y{0, 0, 0};
};

// The library's stream.  Reading it as a cudaStream_t (every kernel launch, copy, synchronisation, NCCL call) first
// submits the vector program recorded so far (vprog.cu), so recorded micro-ops can never be overtaken; `.raw` is the
// handle itself.
void vp_flush();
struct StreamRef {
    cudaStream_t raw = nullptr;
    operator cudaStream_t() const { vp_flush(); return raw; }
};

struct Context {
    bool inited = false;
    int device = 0;
    int sm_count = 148;
    StreamRef stream;
    bool vprog = true;                   // record small vector kernels into one launch (vprog.cu); tb_set_vprog
    uint64_t vprog_launches = 0, vprog_ops = 0, vprog_wide_launches = 0;
    std::vector<Buffer> bufs;            // handle = index + 1
    std::vector<int64_t> free_ids;
    // scratch for two-stage (deterministic) reductions / matvec partials
    char* scratch = nullptr;
    size_t scratch_bytes = 0;
    // scalar mailbox: pinned host memory for D2H of reduction results / get1
    double* mailbox_host = nullptr;
    double* mailbox_dev = nullptr;       // device-side staging for reduction results
    unsigned int* tickets = nullptr;     // "last block done" counters: [0,64) fixed roles, [64, 64+kTicketPool) per-group pool
    static constexpr int kTicketPool = 4096;
    // host box: 16 doubles of MAPPED pinned memory a kernel writes a host-visible scalar into (value at [0], then a
    // sequence number at [1]); the host spins on the sequence number instead of cudaMemcpy + cudaStreamSynchronize
    volatile double* hostbox = nullptr;
    double* hostbox_dev = nullptr;       // device alias of hostbox
    uint64_t box_seq = 0;
    double box_wait_s = 0.0;             // host time spent spinning on the box since the last tb_host_wait_stats
    uint64_t box_waits = 0;
    // slab for tiny buffers (the solver wraps a 1-element slice every iteration: solver.rs:590-591)
    char* small_slab = nullptr;
    std::vector<int> small_free;
    // Released device blocks kept for the next buffer of the same (256-byte rounded) size: every Solver::solve wraps its `work`
    // slice anew (solver.rs:315) and drops it at the end, and cudaMalloc / cudaFree are the two calls of that sequence that
    // serialise with everything else in the driver (measured: 1-250 ms stalls of begin / end when a monitoring tool polls the
    // GPU).  Blocks above kPoolMaxBlock (the matrices) are returned to the driver as before.
    static constexpr size_t kPoolMaxBlock = size_t(256) << 20, kPoolMaxBytes = size_t(1) << 30;
    std::multimap<size_t, char*> pool;
    size_t pool_bytes = 0;
    static constexpr size_t kSmallBytes = 256;
    static constexpr int kSmallSlots = 1024;
    uint64_t launches = 0;
    uint64_t buf_gen = 0;
    // wrapped host ranges by start address (tb_view_of_host): O(log #buffers) per operand instead of a scan of the table,
    // which matters on the stock front-end routes (ProbSOCP wraps one MatOp array per cone block)
    std::multimap<const char*, tb_handle> host_index;
    size_t host_max_bytes = 0;           // longest wrapped range so far: bounds the backward walk of a lookup
    int gemv_mode = 0;
    bool psd_pairing = true;             // park the first ConePSD projection of an iteration and batch it with the second (cone.cu)
    uint64_t psd_pairs = 0;
    uint64_t cone_pairs = 0;
    uint64_t sets_skipped = 0;        // tb_set1 calls that wrote back the value both copies already held          // pairs of projections served by one cone_kernel launch
    int psd_mode = 0;                    // 0: matrix-sign iteration (tcgen05 GEMMs for f32), 1: Jacobi eigendecomposition, 2: sign on FP32/FP64 pipes, 3: tcgen05 without split-K
    char* eig_scratch = nullptr;         // 3 k*k matrices for the sign iteration
    size_t eig_scratch_bytes = 0;
    // event-pair profiling of the streaming matvec
    bool prof_on = false;
    struct ProfRec { cudaEvent_t e0, e1; int variant; double bytes; };   // variant = NN * 3 + NT of stream_kernel<T, NN, NT>
    std::vector<ProfRec> prof_recs;      // recorded, not yet read
    std::vector<cudaEvent_t> prof_pool;
    std::vector<DenseOp*> denseops;
    std::vector<ConeSet*> cones;
    // deferred commands (see "lazy op/trans_op pairing" below); non-empty only while queue[0] is a dense apply
    // waiting for its partner
    std::vector<Cmd> queue;
    bool pair_fusion = true;
    uint64_t pairs_fused = 0;
    // speculative pairing (gemv.cu): the next pair's products computed in the current read of A
    bool speculation = true;
    uint64_t spec_launched = 0, spec_served = 0, spec_dropped = 0;
    // distributed
    int rank = 0, world = 1;
    void* nccl_comm = nullptr;
};

Context& ctx();
// One lock around every C-ABI entry point.  The reference keeps its managers in thread_local! (cuda_mgr.rs:113,
// f32cuda_slice.rs:89) because LinAlg has no `self`; here the context is process-global (one process drives one GPU), so
// calls from several host threads - e.g. cargo's parallel #[test]s - are serialised instead of racing on the buffer
// table, the deferred-command queue and the host box.  Recursive: an entry point may run deferred commands that re-enter.
std::recursive_mutex& api_mutex();
void set_last_error(const std::string& m);      // per host thread
void bind_thread();                             // the CUDA current device is per host thread: make this thread use the context's
void require_init();

Buffer& get_buf(tb_handle h);
// Host-visible scalar protocol: `uint64_t s = box_next();` -> launch a kernel that ends with box_post(hostbox_dev, v, s)
// in exactly one thread -> `double v = box_wait(s);`
uint64_t box_next();
double box_wait(uint64_t seq);
// Make [off, off+len) current on the device; if `write`, mark it device-newer.  Returns the device pointer
// of element `off`.  `full_overwrite` skips the upload of host-newer data the kernel will overwrite anyway.
char* dev_ptr(const tb_view& v, int dtype, bool write, bool full_overwrite = false);
void* scratch(size_t bytes);

template <typename T> struct DT;
template <> struct DT<float> { static constexpr int id = TB_F32; };
template <> struct DT<double> { static constexpr int id = TB_F64; };

// A 1-element operand whose host copy is current (the solver wraps `work_one = [1]` anew every iteration and applies the n x 1
// operators b and c to it, solver.rs:590-596): its value can travel by value instead of being uploaded and read back on the device.
bool host_scalar_if_current(const tb_view& x, int dtype, double* out);
template <typename T> inline const T* rptr(const tb_view& v) { return reinterpret_cast<const T*>(dev_ptr(v, DT<T>::id, false)); }
template <typename T> inline T* wptr(const tb_view& v, bool full_overwrite = false) { return reinterpret_cast<T*>(dev_ptr(v, DT<T>::id, true, full_overwrite)); }

inline void count_launch(int n = 1) { ctx().launches += (uint64_t)n; }

// Programmatic dependent launch for the chain of small dependent kernels that makes up a solver iteration between two
// streaming passes: a kernel launched through launch_pdl may become resident while its predecessor in the stream is still
// running and then blocks in tbd::pdl_entry() (griddepcontrol.wait) until that predecessor has completed and its writes are
// visible - the launch latency (2-3 us per dependent launch, measured on the PSD GEMM chain: 12.3 -> 9.1 us) is hidden
// behind the predecessor instead of being paid on the critical path.  EVERY kernel launched this way starts with
// tbd::pdl_entry() before it touches global memory (reads AND writes: the predecessor may still be reading what this kernel
// overwrites).  TB_PDL=0 / tb_set_pdl(0) launches normally (the entry instructions are then no-ops).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    int na = 0;
    if (pdl_enabled()) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        na = 1;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    TB_CUDA(cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...));
}
// Device timeline (tb_timeline_begin / tb_timeline_dump): with it on, every launch site records a CUDA event behind its kernel
// and the host time of the launch call, so one solver iteration can be laid out kernel by kernel with its real gaps - under the
// real overlap (PDL, host round trips, peer waits), which a serialising profiler cannot show.  Off: one predictable branch.
extern bool g_timeline_on;
void timeline_mark(const char* file, int line);
#define TB_LAUNCH_CHECK()                                                       \
    do {                                                                        \
        TB_CUDA(cudaGetLastError());                                            \
        ::tb::count_launch();                                                   \
        if (::tb::g_timeline_on) ::tb::timeline_mark(__FILE__, __LINE__);       \
    } while (0)

// A parked cone projection (cone.cu: the first of the two projections of an iteration waits for its partner) runs before
// anything else touches the library.
extern bool g_cone_pending;
void cone_flush_pending();

// API wrapper: translate internal exceptions to status codes.
// Every C-ABI entry point runs inside an ApiScope: an NVTX range named after the entry point (visible in nsys / ncu
// timelines; a no-op costing one predictable branch when no tool is attached and tracing is off) and, with
// tb_set_api_trace(1), host-side call counts and wall time per entry point (tb_api_trace_dump).
struct ApiScope {
    const char* name;
    bool on;
    double t0;
    explicit ApiScope(const char* n);
    ~ApiScope();
};
template <typename F> inline int api_keep_pending_impl(const char* name, F f) {
    std::lock_guard<std::recursive_mutex> lock(api_mutex());
    ApiScope scope(name);
    try {
        bind_thread();
        f();
        return TB_OK;
    } catch (const Error& e) {
        set_last_error(e.msg);
        return e.code;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return TB_ERR_STATE;
    }
}
template <typename F> inline int api_raw_impl(const char* name, F f) {
    return api_keep_pending_impl(name, [&] {
        if (g_cone_pending) cone_flush_pending();
        f();
    });
}
void queue_drain();     // run every deferred command in program order (context.cu)
// Default entry-point wrapper: anything deferred runs first, then the call itself - used by every function that
// returns a value to the host, manages buffers, or is not worth deferring.
template <typename F> inline int api_impl(const char* name, F f) {
    return api_raw_impl(name, [&] {
        if (!ctx().queue.empty()) queue_drain();
        f();
    });
}
// Deferrable entry point: runs at once unless a dense apply is parked, in which case it queues behind it.
// `f` must capture its arguments by value.
template <typename F> inline int api_defer_impl(const char* name, std::initializer_list<tb_view> reads, std::initializer_list<tb_view> writes, F f) {
    return api_raw_impl(name, [&] {
        Context& c = ctx();
        if (c.queue.empty()) { f(); return; }
        Cmd cmd;
        for (const tb_view& v : reads) cmd.reads[cmd.n_reads++] = v;
        for (const tb_view& v : writes) cmd.writes[cmd.n_writes++] = v;
        cmd.run = f;
        c.queue.push_back(std::move(cmd));
        if (c.queue.size() > 16) queue_drain();
    });
}
// the entry points call these through macros so that the scope carries the entry point's own name
#define api_keep_pending(...) ::tb::api_keep_pending_impl(__func__, __VA_ARGS__)
#define api_raw(...) ::tb::api_raw_impl(__func__, __VA_ARGS__)
#define api(...) ::tb::api_impl(__func__, __VA_ARGS__)
#define api_defer(...) ::tb::api_defer_impl(__func__, __VA_ARGS__)
bool cmds_conflict(const Cmd& a, const Cmd& b);
// speculative pairing hooks (gemv.cu): every range that changes on the device, every buffer that goes away
void spec_note_write(tb_handle buf, size_t off, size_t len);
void spec_note_release(tb_handle buf);
void spec_reset();
// scalar prefetch hooks (prefetch.cu): reductions that followed a host-visible request last time ride on its round trip
void pf_before_wait();
void pf_into_program();        // the same reductions recorded into the pending vector program instead (one launch fewer, no latency chain of their own)
void pf_note_write(tb_handle buf, size_t off, size_t len);
void pf_note_release(tb_handle buf);
void pf_reset();
bool pf_try_dot(int dtype, const tb_view& a, const tb_view& x, const tb_view& y, double alpha, double beta, double* value_out);
bool pf_try_sumsq(int dtype, const tb_view& a, double* value_out);
void pf_miss_fetch(int dtype, const tb_view& elem);
template <typename T> void set_scalar(const tb_view& one, T val);     // SliceLike::set on a 1-element view (context.cu)

// ---- level-1 internals reused across translation units -------------------------------------------------
template <typename T> void l1_scale(T alpha, T* x, size_t n);
template <typename T> void l1_axpby(T alpha, const T* x, T beta, T* y, size_t n);   // y = alpha*x + beta*y (beta==0: y not read)
template <typename T> void l1_copy(const T* x, T* y, size_t n);
template <typename T> double l1_sumsq_sync(const T* x, size_t n);                   // returns sum of squares (double), syncs
template <typename T> void l1_sumsq_async(const T* x, size_t n, double* out_dev);   // same, result stays on the device

// ---- level-2 internals ----------------------------------------------------------------------------------
// y[i] = alpha * sum_j part[j*ld + i] + beta*y[i]   (fixed summation order; beta == 0: y is not read)
template <typename T> void l2_finalize(const T* part, int nparts, size_t ld, size_t len, T alpha, T beta, T* y);

// ---- distributed internals (dist.cu) --------------------------------------------------------------------
void dist_allreduce_sum(void* buf, size_t count, int dtype);
void dist_allgather_inplace(void* base, size_t count_per_rank, int dtype);   // rank r's slice lives at base + r*count
// matvec epilogues fused with their collective (equal shard sizes, rank-major):
//   gather: y_base[rank*len_local + i] = alpha*sum_j part[j*ld+i] + beta*(old value), then all ranks hold all slices
//   reduce: y[i] = alpha * sum_ranks sum_j part[j*ld+i] + beta*y[i], summed in rank order on every rank
template <typename T> void dist_finalize_gather(const T* part, int nparts, size_t ld, size_t len_local, T alpha, T beta, T* y_base);
template <typename T> void dist_finalize_reduce(const T* part, int nparts, size_t ld, size_t n, T alpha, T beta, T* y);
template <typename T>
void dist_finalize_pair(const T* part_n, int nparts_n, size_t ld_n, size_t len_local, T alpha_n, T beta_n, T* y_base,
                        const T* part_t, int nparts_t, size_t ld_t, size_t n, T alpha_t, T beta_t, T* y_t);     // both in one exchange
// the same plus the raw products of a speculated pair in the same exchange; false (nothing done) when the peer path cannot take it
template <typename T>
bool dist_finalize_pair_spec(const T* part_n, int nparts_n, size_t ld_n, size_t len_local, T alpha_n, T beta_n, T* y_base,
                             const T* part_t, int nparts_t, size_t ld_t, size_t n, T alpha_t, T beta_t, T* y_t,
                             const T* spec_n, const T* spec_t, T* raw_n, T* raw_t);
void dist_check_fault();      // throws if a peer-exchange wait timed out

// ---- eig internals (cone.cu calls into eig.cu for PSD blocks) -------------------------------------------
template <typename T> void psd_project(T* x, size_t sn, T eps_zero, T* work, size_t work_len);
bool psd_pair_usable(size_t sn, const float* work);
void psd_project_pair(float* x0, float* x1, size_t sn, float* work, size_t work_len);     // two projections, every GEMM step one batched launch
// tcgen05 3xTF32 symmetric GEMM (psd_tc.cu): C = alpha*A*B + beta*D + gamma*I, all symmetric k x k column-major f32
bool symm_gemm_tc_usable(const float* A, const float* B, const float* D, const float* C, size_t k);
void symm_gemm_tc_pair(const float* const A[2], const float* const B[2], const float* const D[2], float* const C[2], size_t k,
                       float alpha, float beta, float gamma, int splitk);   // two independent products, one launch
void symm_gemm_tc(const float* A, const float* B, const float* D, float* C, size_t k, float alpha, float beta, float gamma, int splitk,
                  unsigned long long* trace = nullptr);   // trace: 16 device u64 phase stamps of CTA 0 (diagnostics)

}  // namespace tb

// ---- device helpers -------------------------------------------------------------------------------------
namespace tbd {

// first statement of every kernel launched through tb::launch_pdl: let the NEXT kernel in the stream become resident, then
// wait until the PREVIOUS one has completed and flushed (no-ops under a normal launch)
__device__ __forceinline__ void pdl_entry() {
    asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// publish a host-visible scalar: value first, then the sequence number the host spins on
__device__ __forceinline__ void box_post(double* box, double v, unsigned long long seq) {
    *reinterpret_cast<volatile double*>(box) = v;
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(box + 1) = seq;
}

// Deterministic block-wide sum; result valid in thread 0 (and broadcast through smem to all).
// `red` must hold >= 32 doubles.  blockDim.x must be a multiple of 32.
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = (lane < nw) ? red[lane] : 0.0;
        r = warp_sum(r);
        if (lane == 0) red[0] = r;
    }
    __syncthreads();
    return red[0];
}
__device__ __forceinline__ double block_min(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_min(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double r = (lane < nw) ? red[lane] : red[0];
        r = warp_min(r);
        if (lane == 0) red[0] = r;
    }
    __syncthreads();
    return red[0];
}

}  // namespace tbd
