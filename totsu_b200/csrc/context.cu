// Context, buffer table and host/device coherence: the SliceLike role of the backend
// (reference: solver_rust_conic/totsu_core/src/solver/slicelike.rs:9-70; prior art being replaced:
// totsu_f32cuda/src/f32cuda_slice.rs and cuda_mgr.rs).  B200-first redesign: one device allocation per root
// slice, range-based dirty tracking instead of a per-split state machine + HashMap, a slab for the tiny
// 1-element slices the solver wraps every iteration, and a pinned mailbox for host-visible scalars.
#include "common.cuh"
#include "vprog.cuh"
#include <cstdlib>
#include <chrono>
#include <map>
#include <nvtx3/nvToolsExt.h>

namespace tb {

static Context g_ctx;
Context& ctx() { return g_ctx; }

std::recursive_mutex& api_mutex() {
    static std::recursive_mutex m;
    return m;
}
static thread_local std::string t_last_error;
void set_last_error(const std::string& m) { t_last_error = m; }
// ---- per-entry-point NVTX ranges + optional host-side call statistics (tb_set_api_trace / tb_api_trace_dump) ----
static bool g_api_trace = false;
static bool g_nvtx = true;                       // nvtx3 is header-only: calls are no-ops unless a tool injected itself
struct ApiStat { uint64_t calls = 0; double seconds = 0.0; };
static std::map<std::string, ApiStat> g_api_stats;
static thread_local int t_api_depth = 0;
static inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
ApiScope::ApiScope(const char* n) : name(n), on(false), t0(0.0) {
    if (t_api_depth++ > 0) return;               // nested scopes (deferred commands re-entering) belong to the outer call
    if (g_nvtx) nvtxRangePushA(n);
    if (g_api_trace) { on = true; t0 = now_s(); }
}
ApiScope::~ApiScope() {
    if (--t_api_depth > 0) return;
    if (on) {
        ApiStat& st = g_api_stats[name];
        st.calls += 1;
        st.seconds += now_s() - t0;
    }
    if (g_nvtx) nvtxRangePop();
}
// ---- device timeline ------------------------------------------------------------------------------------------
bool g_timeline_on = false;
namespace {
struct TimelineEntry { const char* file; int line; cudaEvent_t ev; double host_s; };
std::vector<TimelineEntry> g_timeline;
std::vector<cudaEvent_t> g_timeline_pool;
size_t g_timeline_cap = 0;
}  // namespace
void timeline_mark(const char* file, int line) {
    if (g_timeline.size() >= g_timeline_cap) return;
    cudaEvent_t ev = g_timeline_pool[g_timeline.size()];
    if (cudaEventRecord(ev, g_ctx.stream.raw) != cudaSuccess) return;
    const char* base = file;
    for (const char* p = file; *p; ++p) if (*p == '/') base = p + 1;
    g_timeline.push_back(TimelineEntry{base, line, ev, std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count()});
}
static int g_pdl = -1;
bool pdl_enabled() {
    if (g_pdl < 0) {
        const char* e = std::getenv("TB_PDL");
        g_pdl = (e && std::atoi(e) == 0) ? 0 : 1;
    }
    return g_pdl != 0;
}
static thread_local int t_bound_device = -1;
void bind_thread() {
    if (g_ctx.inited && t_bound_device != g_ctx.device) {
        TB_CUDA(cudaSetDevice(g_ctx.device));
        t_bound_device = g_ctx.device;
    }
}

void require_init() {
    if (!g_ctx.inited) fail(TB_ERR_STATE, "tb_init has not been called");
}

Buffer& get_buf(tb_handle h) {
    Context& c = ctx();
    if (h <= 0 || (size_t)h > c.bufs.size() || !c.bufs[(size_t)h - 1].alive) fail(TB_ERR_ARG, "invalid buffer handle");
    return c.bufs[(size_t)h - 1];
}

static inline bool views_overlap(const tb_view& a, const tb_view& b) {
    return a.buf == b.buf && a.len > 0 && b.len > 0 && a.off < b.off + b.len && b.off < a.off + a.len;
}

// true if the two commands may not be reordered: one writes what the other reads or writes
bool cmds_conflict(const Cmd& a, const Cmd& b) {
    for (int i = 0; i < a.n_writes; ++i) {
        for (int j = 0; j < b.n_reads; ++j) if (views_overlap(a.writes[i], b.reads[j])) return true;
        for (int j = 0; j < b.n_writes; ++j) if (views_overlap(a.writes[i], b.writes[j])) return true;
    }
    for (int i = 0; i < b.n_writes; ++i)
        for (int j = 0; j < a.n_reads; ++j) if (views_overlap(b.writes[i], a.reads[j])) return true;
    return false;
}

void queue_drain() {
    Context& c = ctx();
    if (c.queue.empty()) return;
    std::vector<Cmd> q;
    q.swap(c.queue);                // a throwing command drops the rest: the error reaches the caller of this API call
    for (Cmd& cmd : q) cmd.run();
}

uint64_t box_next() { return ++ctx().box_seq; }

double box_wait(uint64_t seq) {
    Context& c = ctx();
    volatile unsigned long long* flag = reinterpret_cast<volatile unsigned long long*>(c.hostbox + 1);
    struct Timer {      // host time spent waiting for the device: tells a launch-bound host from a busy device (tb_host_wait_stats)
        Context& c; std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        ~Timer() { c.box_wait_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); c.box_waits += 1; }
    } timer{c};
    pf_before_wait();        // reductions predicted to be asked for next are enqueued right behind the kernel being waited for
    for (unsigned long long spins = 0;; ++spins) {
        if (*flag == seq) break;
        if ((spins & 0xFFFF) == 0xFFFF) {
            // every ~65k polls make sure the stream is still healthy: a faulted kernel never posts
            cudaError_t e = cudaStreamQuery(c.stream);
            if (e != cudaSuccess && e != cudaErrorNotReady) TB_CUDA(e);
            if (e == cudaSuccess && *flag != seq) fail(TB_ERR_STATE, "host box: stream drained without a post");
        }
    }
    return c.hostbox[0];
}

void* scratch(size_t bytes) {
    Context& c = ctx();
    if (bytes > c.scratch_bytes) {
        // grow; the stream is drained first so no in-flight kernel still reads the old block
        TB_CUDA(cudaStreamSynchronize(c.stream));
        if (c.scratch) TB_CUDA(cudaFree(c.scratch));
        size_t nb = std::max(bytes, c.scratch_bytes * 2);
        nb = (nb + 255) & ~size_t(255);
        TB_CUDA(cudaMalloc(&c.scratch, nb));
        c.scratch_bytes = nb;
    }
    return c.scratch;
}

// Tiny host->device updates travel as a kernel argument: fully asynchronous, no pageable-memory staging
// (a cudaMemcpyAsync from pageable memory may synchronise the stream).
struct SmallPayload { unsigned char b[Context::kSmallBytes]; };
__global__ void poke_kernel(unsigned char* dst, SmallPayload p, unsigned int nbytes) {
    tbd::pdl_entry();
    for (unsigned int i = threadIdx.x; i < nbytes; i += blockDim.x) dst[i] = p.b[i];
}

char* dev_ptr(const tb_view& v, int dtype, bool write, bool full_overwrite) {
    if (v.buf == 0 && v.len == 0) return nullptr;       // the empty view (tb_view_of_host of a 0-length slice)
    Buffer& b = get_buf(v.buf);
    if (b.dtype != dtype) fail(TB_ERR_ARG, "view element type does not match the function suffix");
    if (v.off > b.len || v.len > b.len - v.off) fail(TB_ERR_ARG, "view out of range");
    Context& c = ctx();
    if (!b.host_newer.empty() && v.len > 0) {
        if (!(write && full_overwrite)) {
            b.host_newer.for_each_in(v.off, v.off + v.len, [&](size_t lo, size_t hi) {
                spec_note_write(v.buf, lo, hi - lo);          // the device copy of [lo, hi) is about to change
                size_t nb = (hi - lo) * b.esize;
                if (nb <= Context::kSmallBytes) {
                    SmallPayload p;
                    std::memcpy(p.b, b.host + lo * b.esize, nb);
                    launch_pdl(poke_kernel, dim3(1), dim3(64), 0, c.stream, (unsigned char*)(b.dev + lo * b.esize), p, (unsigned int)nb);
                    TB_LAUNCH_CHECK();
                } else {
                    TB_CUDA(cudaMemcpyAsync(b.dev + lo * b.esize, b.host + lo * b.esize, nb, cudaMemcpyHostToDevice, c.stream));
                }
            });
        }
        b.host_newer.sub(v.off, v.off + v.len);
    }
    if (write && v.len > 0) spec_note_write(v.buf, v.off, v.len);
    if (write && b.host && b.host_mut && v.len > 0) b.dev_newer.add(v.off, v.off + v.len);
    return b.dev + v.off * b.esize;
}

bool host_scalar_if_current(const tb_view& x, int dtype, double* out) {
    if (x.len != 1 || x.buf == 0) return false;
    Buffer& b = get_buf(x.buf);
    if (b.dtype != dtype || b.host == nullptr || x.off >= b.len) return false;
    if (b.dev_newer.intersects(x.off, x.off + 1)) return false;
    *out = dtype == TB_F32 ? (double)reinterpret_cast<const float*>(b.host)[x.off] : reinterpret_cast<const double*>(b.host)[x.off];
    return true;
}

static tb_handle new_handle() {
    Context& c = ctx();
    if (!c.free_ids.empty()) {
        tb_handle h = c.free_ids.back();
        c.free_ids.pop_back();
        return h;
    }
    c.bufs.emplace_back();
    return (tb_handle)c.bufs.size();
}

static void alloc_dev(Buffer& b) {
    Context& c = ctx();
    size_t bytes = b.len * b.esize;
    if (bytes <= Context::kSmallBytes && !c.small_free.empty()) {
        b.small_slot = c.small_free.back();
        c.small_free.pop_back();
        b.dev = c.small_slab + (size_t)b.small_slot * Context::kSmallBytes;
    } else {
        b.small_slot = -1;
        const size_t rounded = (std::max<size_t>(bytes, 16) + 255) & ~size_t(255);
        auto it = c.pool.find(rounded);
        if (it != c.pool.end()) {
            b.dev = it->second;
            c.pool.erase(it);
            c.pool_bytes -= rounded;
        } else {
            TB_CUDA(cudaMalloc(&b.dev, rounded));
        }
    }
}

// give a device block back: to the pool when it is small enough to be worth keeping, else to the driver.  The caller has
// drained the stream (nothing in flight still touches the block).
static void free_dev(char* dev, size_t bytes) {
    Context& c = ctx();
    const size_t rounded = (std::max<size_t>(bytes, 16) + 255) & ~size_t(255);
    if (rounded <= Context::kPoolMaxBlock) {
        while (c.pool_bytes + rounded > Context::kPoolMaxBytes && !c.pool.empty()) {
            auto big = std::prev(c.pool.end());
            cudaFree(big->second);
            c.pool_bytes -= big->first;
            c.pool.erase(big);
        }
        c.pool.emplace(rounded, dev);
        c.pool_bytes += rounded;
        return;
    }
    TB_CUDA(cudaFree(dev));
}

static void host_sync_range(Buffer& b, size_t off, size_t len) {
    Context& c = ctx();
    bool any = false;
    b.dev_newer.for_each_in(off, off + len, [&](size_t lo, size_t hi) {
        TB_CUDA(cudaMemcpyAsync(b.host + lo * b.esize, b.dev + lo * b.esize, (hi - lo) * b.esize,
                                cudaMemcpyDeviceToHost, c.stream));
        any = true;
    });
    if (any) {
        TB_CUDA(cudaStreamSynchronize(c.stream));
        b.dev_newer.sub(off, off + len);
    }
}

template <typename T> __global__ void set1_kernel(T* p, T v) { tbd::pdl_entry(); *p = v; }
template <typename T> __global__ void fetch1_kernel(const T* p, double* box, unsigned long long seq) { tbd::pdl_entry(); tbd::box_post(box, (double)*p, seq); }

template <typename T> static void get1(const tb_view& v, size_t idx, T* out) {
    require_init();
    Buffer& b = get_buf(v.buf);
    TB_REQUIRE(b.dtype == DT<T>::id, "dtype mismatch");
    TB_REQUIRE(v.off <= b.len && v.len <= b.len - v.off && idx < v.len, "index out of range");
    size_t i = v.off + idx;
    Context& c = ctx();
    if (b.host && !b.dev_newer.intersects(i, i + 1)) {   // host copy is current: no device round trip
        *out = reinterpret_cast<const T*>(b.host)[i];
        return;
    }
    pf_miss_fetch(DT<T>::id, tb_view{v.buf, i, 1});
    if (vp_enabled()) {
        *out = (T)vp_fetch_to_host(DT<T>::id, b.dev + i * b.esize);      // last micro-op of the pending vector program
    } else {
        const uint64_t seq = box_next();
        launch_pdl(fetch1_kernel<T>, dim3(1), dim3(1), 0, c.stream, reinterpret_cast<const T*>(b.dev + i * b.esize), c.hostbox_dev, (unsigned long long)seq);
        TB_LAUNCH_CHECK();
        *out = (T)box_wait(seq);            // T -> double -> T is exact
    }
    if (b.host && b.host_mut) {      // like SliceLike::get -> get_ref on the 1-element split: host copy becomes current
        reinterpret_cast<T*>(b.host)[i] = *out;
        b.dev_newer.sub(i, i + 1);
    }
}

template <typename T> static void set1(const tb_view& v, size_t idx, T val) {
    require_init();
    TB_REQUIRE(idx < v.len, "index out of range");
    {
        // The solver writes back what it has just read: tau := max(tau, 0), kappa := min(kappa, 0) (solver.rs:551-553, 566-568).
        // When both copies of the element agree and already hold exactly this value the set changes nothing on either side:
        // no launch, and - nothing on the device being touched - no speculation or prefetch is invalidated.
        Buffer& b0 = get_buf(v.buf);
        TB_REQUIRE(b0.dtype == DT<T>::id, "dtype mismatch");
        TB_REQUIRE(v.off <= b0.len && v.len <= b0.len - v.off, "view out of range");
        const size_t i = v.off + idx;
        if (b0.host && b0.host_mut && !b0.dev_newer.intersects(i, i + 1) && !b0.host_newer.intersects(i, i + 1) &&
            std::memcmp(b0.host + i * b0.esize, &val, sizeof(T)) == 0) {
            ctx().sets_skipped += 1;
            return;
        }
    }
    tb_view one{v.buf, v.off + idx, 1};
    T* p = wptr<T>(one, true);
    if (vp_enabled()) {
        vp_set1(DT<T>::id, p, (double)val);
    } else {
        launch_pdl(set1_kernel<T>, dim3(1), dim3(1), 0, ctx().stream, p, val);
        TB_LAUNCH_CHECK();
    }
    Buffer& b = get_buf(v.buf);
    if (b.host && b.host_mut) {      // write-through: both copies agree, a following get() needs no device round trip
        reinterpret_cast<T*>(b.host)[v.off + idx] = val;
        b.dev_newer.sub(v.off + idx, v.off + idx + 1);
    }
}

template <typename T> void set_scalar(const tb_view& one, T val) { set1<T>(one, 0, val); }
template void set_scalar<float>(const tb_view&, float);
template void set_scalar<double>(const tb_view&, double);

}  // namespace tb

using namespace tb;

extern "C" {

int tb_init(int device) {
    return api([&] {
        Context& c = ctx();
        if (c.inited) return;
        if (device < 0) {
            const char* lr = std::getenv("LOCAL_RANK");
            device = lr ? std::atoi(lr) : 0;
        }
        int ndev = 0;
        TB_CUDA(cudaGetDeviceCount(&ndev));
        if (ndev <= 0) fail(TB_ERR_CUDA, "no CUDA device visible: totsu_b200 has no CPU fallback");
        if (device >= ndev) fail(TB_ERR_ARG, "device ordinal out of range");
        TB_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        TB_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) fail(TB_ERR_UNSUPPORTED, std::string("totsu_b200 is built for sm_100a (B200); found ") + prop.name);
        c.device = device;
        t_bound_device = device;
        c.sm_count = prop.multiProcessorCount;
        TB_CUDA(cudaStreamCreateWithFlags(&c.stream.raw, cudaStreamNonBlocking));
        vp_init();
        TB_CUDA(cudaHostAlloc(&c.mailbox_host, 64 * sizeof(double), cudaHostAllocDefault));
        TB_CUDA(cudaMalloc(&c.mailbox_dev, 64 * sizeof(double)));
        TB_CUDA(cudaMalloc(&c.tickets, (64 + Context::kTicketPool) * sizeof(unsigned int)));
        TB_CUDA(cudaMemset(c.tickets, 0, (64 + Context::kTicketPool) * sizeof(unsigned int)));
        {
            void* hb = nullptr;
            TB_CUDA(cudaHostAlloc(&hb, 16 * sizeof(double), cudaHostAllocMapped));
            std::memset(hb, 0, 16 * sizeof(double));
            c.hostbox = reinterpret_cast<volatile double*>(hb);
            void* hd = nullptr;
            TB_CUDA(cudaHostGetDevicePointer(&hd, hb, 0));
            c.hostbox_dev = reinterpret_cast<double*>(hd);
            c.box_seq = 0;
        }
        TB_CUDA(cudaMalloc(&c.small_slab, Context::kSmallBytes * Context::kSmallSlots));
        TB_CUDA(cudaMemset(c.small_slab, 0, Context::kSmallBytes * Context::kSmallSlots));
        c.small_free.clear();
        for (int i = Context::kSmallSlots - 1; i >= 0; --i) c.small_free.push_back(i);
        c.scratch = nullptr;
        c.scratch_bytes = 0;
        (void)scratch(size_t(8) << 20);
        c.launches = 0;
        if (const char* e = std::getenv("TB_NVTX")) g_nvtx = std::atoi(e) != 0;
        if (const char* e = std::getenv("TB_VPROG_MAX_N")) { const long v = std::atol(e); if (v >= 1024 && v <= (1l << 22)) vp_set_max_n((size_t)v); }
        c.inited = true;
    });
}

int tb_shutdown(void) {
    return api([&] {
        Context& c = ctx();
        if (!c.inited) return;
        c.queue.clear();
        pf_reset();
        cudaStreamSynchronize(c.stream);
        for (Buffer& b : c.bufs)
            if (b.alive && b.small_slot < 0 && b.dev) cudaFree(b.dev);
        c.bufs.clear();
        c.free_ids.clear();
        c.host_index.clear();
        for (auto& e : c.pool) cudaFree(e.second);
        c.pool.clear();
        c.pool_bytes = 0;
        if (c.scratch) cudaFree(c.scratch);
        c.scratch = nullptr;
        c.scratch_bytes = 0;
        cudaFreeHost(c.mailbox_host);
        cudaFreeHost((void*)c.hostbox);
        c.hostbox = nullptr; c.hostbox_dev = nullptr;
        cudaFree(c.mailbox_dev);
        cudaFree(c.tickets);
        cudaFree(c.small_slab);
        vp_shutdown();
        cudaStreamDestroy(c.stream.raw);
        c.stream.raw = nullptr;
        c.inited = false;
    });
}

const char* tb_last_error(void) { return t_last_error.c_str(); }

int tb_device_sync(void) {
    return api([&] {
        require_init();
        TB_CUDA(cudaStreamSynchronize(ctx().stream));
        dist_check_fault();
    });
}

int tb_get_stream(void** out) {
    return api([&] {
        require_init();
        *out = (void*)(cudaStream_t)ctx().stream;
    });
}

int tb_sm_count(int* out) {
    return api([&] {
        require_init();
        *out = ctx().sm_count;
    });
}

int tb_launch_count(uint64_t* out) {
    return api([&] { *out = ctx().launches; });
}

int tb_set_pair_fusion(int on) {
    return api([&] { ctx().pair_fusion = on != 0; });
}

int tb_set_vprog(int on) {
    return api([&] {
        require_init();
        vp_flush();
        ctx().vprog = on != 0;
    });
}
int tb_set_vprog_max_n(size_t n) {
    return api([&] {
        require_init();
        TB_REQUIRE(n >= 1024 && n <= (size_t(1) << 22), "vprog max n: 1024 .. 4M elements");
        vp_set_max_n(n);
    });
}
int tb_vprog_stats(uint64_t* launches, uint64_t* ops) {
    return api_raw([&] { *launches = ctx().vprog_launches; *ops = ctx().vprog_ops; });
}
int tb_flush(void) {
    return api([&] {
        require_init();
        vp_flush();
    });
}
int tb_set_speculation(int on) {
    return api([&] {
        require_init();
        ctx().speculation = on != 0;
        spec_reset();
    });
}
int tb_spec_stats(uint64_t* launched, uint64_t* served, uint64_t* dropped) {
    return api_raw([&] { *launched = ctx().spec_launched; *served = ctx().spec_served; *dropped = ctx().spec_dropped; });
}
int tb_host_wait_stats(double* seconds, uint64_t* waits) {
    return api_raw([&] {
        *seconds = ctx().box_wait_s; *waits = ctx().box_waits;
        ctx().box_wait_s = 0.0; ctx().box_waits = 0;
    });
}
int tb_timeline_begin(size_t max_events) {
    return api([&] {
        require_init();
        TB_CUDA(cudaStreamSynchronize(ctx().stream));
        g_timeline.clear();
        while (g_timeline_pool.size() < max_events) {
            cudaEvent_t ev;
            TB_CUDA(cudaEventCreate(&ev));
            g_timeline_pool.push_back(ev);
        }
        g_timeline_cap = max_events;
        g_timeline_on = max_events > 0;
    });
}
// "index file:line device_us host_us" per launch since tb_timeline_begin, both clocks relative to the first launch; the device
// time of a launch is the completion time of its kernel.  Stops recording.
int tb_timeline_dump(char* buf, size_t cap, size_t* needed) {
    return api([&] {
        require_init();
        g_timeline_on = false;
        TB_CUDA(cudaStreamSynchronize(ctx().stream));
        std::string out;
        char line[160];
        for (size_t i = 0; i < g_timeline.size(); ++i) {
            float ms = 0.f;
            TB_CUDA(cudaEventElapsedTime(&ms, g_timeline[0].ev, g_timeline[i].ev));
            std::snprintf(line, sizeof line, "%zu %s:%d %.3f %.3f\n", i, g_timeline[i].file, g_timeline[i].line, (double)ms * 1e3,
                          (g_timeline[i].host_s - g_timeline[0].host_s) * 1e6);
            out += line;
        }
        if (needed) *needed = out.size() + 1;
        if (buf && cap > 0) {
            const size_t nb = std::min(cap - 1, out.size());
            std::memcpy(buf, out.data(), nb);
            buf[nb] = 0;
        }
    });
}

int tb_set_api_trace(int on) {
    return api_keep_pending([&] {
        g_api_trace = on != 0;
        if (on) g_api_stats.clear();
    });
}
int tb_api_trace_dump(char* buf, size_t cap, size_t* needed) {
    return api_keep_pending([&] {
        std::string out;
        for (const auto& kv : g_api_stats) {
            char line[256];
            std::snprintf(line, sizeof line, "%s %llu %.9f\n", kv.first.c_str(), (unsigned long long)kv.second.calls, kv.second.seconds);
            out += line;
        }
        if (needed) *needed = out.size() + 1;
        if (buf && cap > 0) {
            const size_t n = std::min(cap - 1, out.size());
            std::memcpy(buf, out.data(), n);
            buf[n] = 0;
        }
    });
}
int tb_pairs_fused(uint64_t* out) {
    return api([&] { *out = ctx().pairs_fused; });
}

int tb_set_pdl(int on) {
    return api([&] {
        if (ctx().inited) TB_CUDA(cudaStreamSynchronize(ctx().stream));
        g_pdl = on != 0 ? 1 : 0;
    });
}

int tb_set_gemv_path(int mode) {
    return api([&] {
        TB_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0, 1 or 2");
        ctx().gemv_mode = mode;
    });
}

int tb_buf_wrap(int dtype, void* host, size_t len, int host_is_mut, tb_handle* out) {
    return api([&] {
        require_init();
        TB_REQUIRE(dtype == TB_F32 || dtype == TB_F64, "bad dtype");
        TB_REQUIRE(host != nullptr || len == 0, "null host pointer");
        if (!host_is_mut && len > 0) {
            // The reference may wrap the same read-only array twice at once (ProbSOCP::problem: mats_g[i].as_op() for op_a and
            // again for the cone's dimensions, socp.rs:450,463): share one mirror instead of uploading and storing it twice,
            // so that retain / release of either wrapper land on the same refcount.
            Context& c = ctx();
            for (size_t i = 0; i < c.bufs.size(); ++i) {
                Buffer& o = c.bufs[i];
                if (o.alive && !o.host_mut && o.host == (char*)host && o.len == len && o.dtype == dtype) {
                    o.refs += 1;
                    *out = (tb_handle)(i + 1);
                    return;
                }
            }
        }
        if (host_is_mut && len > 0) {
            // a mutable wrap must not partially overlap another live mirror: two device copies of one host range cannot be
            // kept coherent (an identical or enclosing range is what nested wrappers produce and resolves to the newest)
            Context& c = ctx();
            const char* lo = (const char*)host;
            const char* hi = lo + len * (dtype == TB_F32 ? 4 : 8);
            for (const Buffer& o : c.bufs) {
                if (!o.alive || !o.host || o.len == 0) continue;
                const char* olo = o.host;
                const char* ohi = o.host + o.len * o.esize;
                const bool overlap = lo < ohi && olo < hi;
                const bool nested = (olo <= lo && hi <= ohi) || (lo <= olo && ohi <= hi);
                if (overlap && !nested) fail(TB_ERR_ARG, "tb_buf_wrap: mutable host range partially overlaps another wrapped slice");
            }
        }
        tb_handle h = new_handle();
        Buffer& b = ctx().bufs[(size_t)h - 1];
        b = Buffer();
        b.dtype = dtype;
        b.esize = dtype == TB_F32 ? 4 : 8;
        b.len = len;
        b.host = (char*)host;
        b.host_mut = host_is_mut != 0;
        try {
            alloc_dev(b);
        } catch (...) {
            ctx().free_ids.push_back(h);
            throw;
        }
        b.alive = true;
        b.gen = ++ctx().buf_gen;
        if (len > 0) {
            ctx().host_index.emplace(b.host, h);
            ctx().host_max_bytes = std::max(ctx().host_max_bytes, len * b.esize);
        }
        if (len > 0) b.host_newer.add(0, len);   // uploaded on first device use (or right away for big read-only data)
        if (!b.host_mut && len * b.esize >= (size_t(1) << 20)) {
            tb_view v{h, 0, len};
            (void)dev_ptr(v, dtype, false);       // matrices: upload now, like MatOp::new -> new_ref does (matop.rs:66-74)
        }
        *out = h;
    });
}

int tb_buf_alloc(int dtype, size_t len, tb_handle* out) {
    return api([&] {
        require_init();
        TB_REQUIRE(dtype == TB_F32 || dtype == TB_F64, "bad dtype");
        tb_handle h = new_handle();
        Buffer& b = ctx().bufs[(size_t)h - 1];
        b = Buffer();
        b.dtype = dtype;
        b.esize = dtype == TB_F32 ? 4 : 8;
        b.len = len;
        try {
            alloc_dev(b);
        } catch (...) {
            ctx().free_ids.push_back(h);
            throw;
        }
        b.alive = true;
        b.gen = ++ctx().buf_gen;
        TB_CUDA(cudaMemsetAsync(b.dev, 0, std::max<size_t>(len * b.esize, 0), ctx().stream));
        *out = h;
    });
}

int tb_buf_release(tb_handle h) {
    // dropping one of several live wrappers (a split child of the binding) is pure bookkeeping: it must not run a parked
    // dense apply or cone projection, or op/trans_op pairing could never fire behind `splitm!` (slicelike.rs:162-201)
    return api_keep_pending([&] {
        require_init();
        Buffer& b0 = get_buf(h);
        if (b0.refs > 1) { b0.refs -= 1; return; }
        if (g_cone_pending) cone_flush_pending();
        if (!ctx().queue.empty()) queue_drain();
        Buffer& b = get_buf(h);
        Context& c = ctx();
        if (--b.refs > 0) return;
        spec_note_release(h);
        if (b.host && b.len > 0) {
            auto range = c.host_index.equal_range(b.host);
            for (auto it = range.first; it != range.second; ++it)
                if (it->second == h) { c.host_index.erase(it); break; }
        }
        if (b.host && b.host_mut) host_sync_range(b, 0, b.len);
        if (b.small_slot >= 0) {
            // slab slots are recycled without a device sync: every use is stream-ordered
            c.small_free.push_back(b.small_slot);
        } else if (b.dev) {
            TB_CUDA(cudaStreamSynchronize(c.stream));
            free_dev(b.dev, b.len * b.esize);
        }
        b = Buffer();
        c.free_ids.push_back(h);
    });
}

int tb_buf_retain(tb_handle h, int n) {
    return api_keep_pending([&] {         // bookkeeping only: nothing deferred has to run first
        require_init();
        TB_REQUIRE(n >= 0, "retain count must be >= 0");
        get_buf(h).refs += n;
    });
}

int tb_view_of_host(int dtype, const void* host, size_t len, tb_view* out) {
    // A pure table lookup: the binding resolves EVERY operand of EVERY call through it (rust/totsu_b200/src/b200_slice.rs
    // `view()`), also between the two halves of an op/trans_op pair and between the two cone projections of an iteration,
    // so it must neither drain the deferred-command queue nor run a parked projection.
    return api_keep_pending([&] {
        require_init();
        Context& c = ctx();
        if (len == 0) { *out = tb_view{0, 0, 0}; return; }
        const size_t es = dtype == TB_F32 ? 4 : 8;
        const char* lo = reinterpret_cast<const char*>(host);
        const char* hi = lo + len * es;
        auto covers = [&](const Buffer& b) {
            return b.alive && b.host && b.dtype == dtype && b.host <= lo && hi <= b.host + b.len * b.esize;
        };
        // nearest wrapped range starting at or below `lo` that covers [lo, hi); among wrappers of the same start address the
        // newest wins (a stack slot re-wrapped every iteration, solver.rs:590-591)
        tb_handle best = 0;
        uint64_t best_gen = 0;
        const char* best_start = nullptr;
        for (auto it = c.host_index.upper_bound(lo); it != c.host_index.begin();) {
            --it;
            if (best != 0 && it->first != best_start) break;
            if ((size_t)(lo - it->first) > c.host_max_bytes) break;
            const Buffer& b = c.bufs[(size_t)it->second - 1];
            if (covers(b) && b.gen > best_gen) { best = it->second; best_gen = b.gen; best_start = it->first; }
        }
        if (best == 0) fail(TB_ERR_ARG, "tb_view_of_host: the host range is not inside any wrapped slice");
        const Buffer& b = c.bufs[(size_t)best - 1];
        *out = tb_view{best, (size_t)(lo - b.host) / es, len};
    });
}

int tb_buf_len(tb_handle h, size_t* out) {
    return api([&] {
        require_init();
        *out = get_buf(h).len;
    });
}

int tb_host_ref(tb_view v) {
    return api([&] {
        require_init();
        if (v.buf == 0 && v.len == 0) return;
        Buffer& b = get_buf(v.buf);
        TB_REQUIRE(b.host != nullptr, "buffer has no host mirror");
        TB_REQUIRE(v.off <= b.len && v.len <= b.len - v.off, "view out of range");
        host_sync_range(b, v.off, v.len);
    });
}

int tb_host_mut(tb_view v) {
    return api([&] {
        require_init();
        if (v.buf == 0 && v.len == 0) return;
        Buffer& b = get_buf(v.buf);
        TB_REQUIRE(b.host != nullptr && b.host_mut, "buffer has no mutable host mirror");
        TB_REQUIRE(v.off <= b.len && v.len <= b.len - v.off, "view out of range");
        host_sync_range(b, v.off, v.len);
        b.host_newer.add(v.off, v.off + v.len);
    });
}

int tb_get1_f32(tb_view v, size_t idx, float* out) { return api([&] { get1<float>(v, idx, out); }); }
int tb_get1_f64(tb_view v, size_t idx, double* out) { return api([&] { get1<double>(v, idx, out); }); }
int tb_set1_f32(tb_view v, size_t idx, float val) { return api([&] { set1<float>(v, idx, val); }); }
int tb_set1_f64(tb_view v, size_t idx, double val) { return api([&] { set1<double>(v, idx, val); }); }

int tb_upload(tb_view v, const void* src) {
    return api([&] {
        require_init();
        Buffer& b = get_buf(v.buf);
        char* d = dev_ptr(v, b.dtype, true, true);
        if (v.len == 0) return;
        TB_CUDA(cudaMemcpyAsync(d, src, v.len * b.esize, cudaMemcpyHostToDevice, ctx().stream));
        TB_CUDA(cudaStreamSynchronize(ctx().stream));
    });
}

int tb_download(tb_view v, void* dst) {
    return api([&] {
        require_init();
        Buffer& b = get_buf(v.buf);
        char* d = dev_ptr(v, b.dtype, false);
        if (v.len == 0) return;
        TB_CUDA(cudaMemcpyAsync(dst, d, v.len * b.esize, cudaMemcpyDeviceToHost, ctx().stream));
        TB_CUDA(cudaStreamSynchronize(ctx().stream));
    });
}

}  // extern "C"
