// Scalar prefetch: fewer host round trips per solver iteration.
//
// The unmodified solver reads six scalars back per iteration - tau, kappa (update_vecs, solver.rs:551,566), g_x, g_y,
// |p|, |d| (criteria_conv, solver.rs:599-608) - and every one is a full host <-> device round trip: the host cannot issue
// the next call before the value has crossed PCIe.  Four of them are reductions over vectors that are ALREADY FINAL when an
// earlier scalar is fetched (g_x = c.x_x/tau and g_y = b.x_y/tau when kappa is read; |d| when |p| is read), so they can ride
// on that earlier round trip: when a host-visible request misses, one extra small kernel (prefetch_reduce_kernel: up to 4
// dot products / sums of squares, two-stage, fixed order, double accumulation) computes the reductions that FOLLOWED this
// request last time and posts them into the mapped host box next to the value being waited for.  When the following
// requests then arrive they are served from the box: no launch, no round trip.  Six round trips become three.
//
//   * nothing about the solver is hard-wired: the library learns "request R was followed by requests F1..F4" from the call
//     stream (a request = plain element fetch | sum of squares of a view | dot product into a 1-element view + its fetch);
//   * validity: every device write goes through dev_ptr(write) -> spec_note_write -> pf_note_write; a write that overlaps the
//     inputs of a prefetched reduction drops it; if F is then still requested, (R -> F) is marked as not worth prefetching
//     again (|p| at the time kappa is read is learnt that way after one wasted attempt); if F is simply not requested this
//     time round (the criteria_inf branch skips |d|), nothing is marked;
//   * alpha is applied at serve time with the arithmetic of the kernel it replaces; the reduction itself is accumulated in
//     double like the vector-program reductions (results agree with the un-prefetched path to rounding, not bit for bit);
//   * tb_set_scalar_prefetch(0) turns it off; tb_scalar_prefetch_stats reports kernels launched / requests served / dropped.
#include "common.cuh"
#include "vprog.cuh"
#include <chrono>
#include <cstdlib>

namespace tb {

namespace {

struct PfReq {
    int kind = 0;      // 0: fetch of one element (a); 1: sum of squares of a; 2: dot(a, b) scaled into the 1-element view y, then fetched
    int dtype = 0;
    tb_view a{0, 0, 0}, b{0, 0, 0}, y{0, 0, 0};
};
inline bool veq(const tb_view& p, const tb_view& q) { return p.buf == q.buf && p.off == q.off && p.len == q.len; }
inline bool req_eq(const PfReq& p, const PfReq& q) {
    if (p.kind != q.kind || p.dtype != q.dtype || !veq(p.a, q.a)) return false;
    return p.kind != 2 || (veq(p.b, q.b) && veq(p.y, q.y));
}
inline bool overlaps(const tb_view& v, tb_handle buf, size_t off, size_t len) {
    return v.buf == buf && v.len > 0 && len > 0 && v.off < off + len && off < v.off + v.len;
}

struct PfChain { PfReq trigger; std::vector<PfReq> followers; std::vector<int> unseen; };      // unseen[i]: rounds in a row follower i was not asked for
struct PfLive { PfReq req; int slot; bool valid; };

constexpr int kMaxJobs = 4;
constexpr int kBoxVal = 4;       // hostbox[4..7]: prefetched values; hostbox[8]: their sequence number
constexpr int kBoxSeq = 8;

struct PfState {
    bool on = true;
    std::vector<PfChain> table, open;
    std::vector<std::pair<PfReq, PfReq>> bad;         // (trigger, follower): the follower's inputs change between the two requests - never prefetch
    std::vector<std::pair<PfReq, PfReq>> suspect;     // a prefetched value was overwritten unread: bad only if the request then still comes
    std::vector<PfLive> live;
    PfReq live_trigger;
    uint64_t live_seq = 0;
    bool have_dot = false;
    PfReq dot;
    std::vector<PfReq> armed;
    PfReq armed_trigger;
    double* partials = nullptr;
    int partials_cap = 0;
    uint64_t seq_counter = 0;
    uint64_t launched = 0, served = 0, dropped = 0;
};
PfState g_pf;

// TB_PF_DEBUG=1: one stderr line per request / launch / drop (diagnostics)
bool pf_debug() {
    static const bool on = [] { const char* e = std::getenv("TB_PF_DEBUG"); return e && e[0] == '1'; }();
    return on;
}
void pf_log(const char* what, const PfReq& r, const char* extra = "") {
    if (!pf_debug()) return;
    std::fprintf(stderr, "[pf] %-8s kind=%d dt=%d a=(%lld,%zu,%zu) b=(%lld,%zu,%zu) y=(%lld,%zu,%zu) %s\n", what, r.kind, r.dtype, (long long)r.a.buf, r.a.off, r.a.len,
                 (long long)r.b.buf, r.b.off, r.b.len, (long long)r.y.buf, r.y.off, r.y.len, extra);
}

struct PfJob { const void* a; const void* b; unsigned long long n; int kind; int f64; };
struct PfJobs {
    int n_jobs;
    PfJob job[kMaxJobs];
    double* partials;
    unsigned int* ticket;
    double* box;
    unsigned long long seq;
};

template <typename T> __device__ __forceinline__ double pf_partial(const PfJob& j) {
    const T* a = reinterpret_cast<const T*>(j.a);
    const T* b = reinterpret_cast<const T*>(j.b);
    double acc = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if (j.kind == 2) {
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < j.n; i += stride) acc += (double)a[i] * (double)b[i];
    } else {
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < j.n; i += stride) { const double v = (double)a[i]; acc += v * v; }
    }
    return acc;
}

// up to 4 reductions in one launch: per-CTA partials in double, the last CTA (ticket) sums them in CTA order and posts all
// results, then the sequence number, into mapped host memory
__global__ void __launch_bounds__(256) prefetch_reduce_kernel(const __grid_constant__ PfJobs J) {
    __shared__ double red[32];
    __shared__ bool last;
    tbd::pdl_entry();
    for (int k = 0; k < J.n_jobs; ++k) {
        const double acc = J.job[k].f64 ? pf_partial<double>(J.job[k]) : pf_partial<float>(J.job[k]);
        const double s = tbd::block_sum(acc, red);
        if (threadIdx.x == 0) J.partials[(size_t)k * gridDim.x + blockIdx.x] = s;
    }
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicInc(J.ticket, gridDim.x - 1);      // wraps back to 0
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    for (int k = 0; k < J.n_jobs; ++k) {
        double a = 0.0;
        for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) a += __ldcg(&J.partials[(size_t)k * gridDim.x + i]);
        const double tot = tbd::block_sum(a, red);
        if (threadIdx.x == 0) *reinterpret_cast<volatile double*>(J.box + kBoxVal + k) = tot;
    }
    if (threadIdx.x == 0) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>(J.box + kBoxSeq) = J.seq;
    }
}

bool is_bad(const PfReq& trig, const PfReq& f) {
    for (const auto& pr : g_pf.bad)
        if (req_eq(pr.first, trig) && req_eq(pr.second, f)) return true;
    return false;
}

// A finished chain replaces what was known about its trigger - except that followers seen in earlier rounds and not in this one
// are kept for a few rounds: a solver iteration that takes the other branch (criteria_inf asks for |d| only when b.y > eps_zero)
// must not make the next ordinary iteration pay a round trip to re-learn |p| -> |d|.
void commit(PfChain&& ch) {
    PfState& S = g_pf;
    ch.unseen.assign(ch.followers.size(), 0);
    for (PfChain& t : S.table) {
        if (!req_eq(t.trigger, ch.trigger)) continue;
        for (size_t i = 0; i < t.followers.size(); ++i) {
            bool seen = false;
            for (const PfReq& f : ch.followers) seen = seen || req_eq(f, t.followers[i]);
            const int unseen = i < t.unseen.size() ? t.unseen[i] + 1 : 1;
            if (!seen && unseen < 8 && ch.followers.size() < (size_t)kMaxJobs) {
                ch.followers.push_back(t.followers[i]);
                ch.unseen.push_back(unseen);
            }
        }
        t.followers = std::move(ch.followers);
        t.unseen = std::move(ch.unseen);
        return;
    }
    if (S.table.size() >= 32) S.table.erase(S.table.begin());
    S.table.push_back(std::move(ch));
}

// one host-visible request (served from the box, or about to miss): learn the chains, and on a miss arm the followers
void on_request(const PfReq& r, bool was_served) {
    PfState& S = g_pf;
    pf_log(was_served ? "served" : "miss", r);
    // A prefetched value of r was overwritten before r was asked for, and now r IS asked for (from the device, or from a later
    // trigger's prefetch): its inputs change between that trigger and the request (|p| and |d| at the time kappa is read) -
    // structural, never prefetch it on that trigger again.  A suspect whose trigger comes round again without the request
    // having been made was only a branch not taken this time (criteria_inf asks for |d| only when b.y > eps_zero,
    // solver.rs:650-655) and stays prefetchable.
    for (size_t i = 0; i < S.suspect.size();) {
        if (req_eq(S.suspect[i].second, r)) {
            if (S.bad.size() >= 64) S.bad.erase(S.bad.begin());
            S.bad.push_back(S.suspect[i]);
            pf_log("bad", r);
            S.suspect.erase(S.suspect.begin() + (long)i);
        } else if (!was_served && req_eq(S.suspect[i].first, r)) {
            S.suspect.erase(S.suspect.begin() + (long)i);
        } else {
            ++i;
        }
    }
    if (r.kind != 0)
        for (PfChain& ch : S.open)
            if (ch.followers.size() < (size_t)kMaxJobs) ch.followers.push_back(r);
    if (r.kind == 0) {                          // a plain fetch ends every chain being learnt
        for (PfChain& ch : S.open) commit(std::move(ch));
        S.open.clear();
    }
    if (!was_served) {
        if (S.open.size() < 8) S.open.push_back(PfChain{r, {}, {}});
        S.armed.clear();
        for (const PfChain& t : S.table) {
            if (!req_eq(t.trigger, r)) continue;
            for (const PfReq& f : t.followers)
                if (!is_bad(r, f) && (int)S.armed.size() < kMaxJobs) S.armed.push_back(f);
        }
        S.armed_trigger = r;
    }
}

PfLive* find_live(const PfReq& r) {
    for (PfLive& e : g_pf.live)
        if (e.valid && req_eq(e.req, r)) return &e;
    return nullptr;
}

double live_value(const PfLive& e) {
    Context& c = ctx();
    volatile unsigned long long* flag = reinterpret_cast<volatile unsigned long long*>(c.hostbox + kBoxSeq);
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned long long spins = 0; *flag != g_pf.live_seq; ++spins) {
        if ((spins & 0xFFFF) == 0xFFFF) {
            cudaError_t err = cudaStreamQuery(c.stream.raw);
            if (err != cudaSuccess && err != cudaErrorNotReady) TB_CUDA(err);
            if (err == cudaSuccess && *flag != g_pf.live_seq) fail(TB_ERR_STATE, "scalar prefetch: stream drained without a post");
        }
    }
    c.box_wait_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return c.hostbox[kBoxVal + e.slot];
}

}  // namespace

// ---- hooks --------------------------------------------------------------------------------------------------
// start of box_wait: the trigger's own kernel is already enqueued; launch the armed reductions right behind it
void pf_before_wait() {
    PfState& S = g_pf;
    if (S.armed.empty()) return;
    std::vector<PfReq> want;
    want.swap(S.armed);
    if (!S.on) return;
    Context& c = ctx();
    PfJobs J{};
    std::vector<PfLive> live;
    size_t max_n = 0;
    for (const PfReq& f : want) {
        try {
            PfJob j{};
            j.kind = f.kind;
            j.f64 = f.dtype == TB_F64 ? 1 : 0;
            j.a = dev_ptr(f.a, f.dtype, false);
            if (f.kind == 2) {
                if (f.a.len != f.b.len) continue;
                j.b = dev_ptr(f.b, f.dtype, false);
            }
            j.n = f.a.len;
            if (j.n == 0) continue;
            pf_log("launch", f);
            live.push_back(PfLive{f, J.n_jobs, true});
            J.job[J.n_jobs++] = j;
            max_n = std::max<size_t>(max_n, j.n);
        } catch (const Error&) {
            pf_log("skip", f);
            continue;                     // a view learnt from an earlier buffer generation no longer resolves: skip it
        }
    }
    if (J.n_jobs == 0) return;
    const int g = (int)std::max<size_t>(1, std::min<size_t>((max_n + 2047) / 2048, (size_t)c.sm_count));
    if (S.partials_cap < g * kMaxJobs) {
        if (S.partials) { TB_CUDA(cudaStreamSynchronize(c.stream.raw)); TB_CUDA(cudaFree(S.partials)); }
        S.partials_cap = c.sm_count * kMaxJobs;
        TB_CUDA(cudaMalloc(&S.partials, (size_t)S.partials_cap * sizeof(double)));
    }
    J.partials = S.partials;
    J.ticket = c.tickets + 40;
    J.box = c.hostbox_dev;
    J.seq = ++S.seq_counter;
    launch_pdl(prefetch_reduce_kernel, dim3(g), dim3(256), 0, c.stream.raw, J);
    TB_LAUNCH_CHECK();
    S.live = std::move(live);
    S.live_trigger = S.armed_trigger;
    S.live_seq = J.seq;
    S.launched += 1;
}

// Called by the vector-program recorder just before it closes a program with a host-visible result (the trigger): the armed
// reductions become micro-ops of that same program - partial sums on the cluster, combined into the prefetch slots of the host
// box, sequence number last - so they cost neither a launch nor the two-stage kernel's latency chain, and the trigger's own value
// is posted before them.  Falls back to pf_before_wait's kernel (armed list left alone) when a vector is too long for the cluster.
void pf_into_program() {
    PfState& S = g_pf;
    static const bool ride = [] { const char* e = std::getenv("TB_PF_IN_PROGRAM"); return !(e && e[0] == '0'); }();      // A/B: 0 = always the separate kernel
    if (S.armed.empty() || !S.on || !ride) return;
    for (const PfReq& f : S.armed)
        if (!vp_enabled_red(f.a.len)) return;
    std::vector<PfReq> want;
    want.swap(S.armed);
    std::vector<PfLive> live;
    int k = 0;
    for (const PfReq& f : want) {
        if (k >= kMaxJobs) break;
        try {
            if (f.a.len == 0 || (f.kind == 2 && f.a.len != f.b.len)) continue;
            const void* a = dev_ptr(f.a, f.dtype, false);
            const void* b = f.kind == 2 ? dev_ptr(f.b, f.dtype, false) : nullptr;
            pf_log("ride", f);
            vp_pf_reduce(f.dtype, f.kind, a, b, f.a.len, k);
            live.push_back(PfLive{f, k, true});
            ++k;
        } catch (const Error&) {
            pf_log("skip", f);
            continue;
        }
    }
    if (k == 0) return;
    const uint64_t seq = ++S.seq_counter;
    vp_pf_seq(seq);
    S.live = std::move(live);
    S.live_trigger = S.armed_trigger;
    S.live_seq = seq;
    S.launched += 1;
}

void pf_note_write(tb_handle buf, size_t off, size_t len) {
    PfState& S = g_pf;
    for (PfLive& e : S.live) {
        if (!e.valid) continue;
        if (overlaps(e.req.a, buf, off, len) || (e.req.kind == 2 && overlaps(e.req.b, buf, off, len))) {
            e.valid = false;
            pf_log("drop", e.req);
            if (S.suspect.size() >= 64) S.suspect.erase(S.suspect.begin());
            S.suspect.push_back({S.live_trigger, e.req});
            S.dropped += 1;
        }
    }
}

void pf_note_release(tb_handle buf) {
    PfState& S = g_pf;
    for (PfLive& e : S.live)
        if (e.valid && (e.req.a.buf == buf || e.req.b.buf == buf || e.req.y.buf == buf)) { e.valid = false; pf_log("released", e.req); }
    if (S.have_dot && (S.dot.a.buf == buf || S.dot.b.buf == buf || S.dot.y.buf == buf)) S.have_dot = false;
}

void pf_reset() {
    PfState& S = g_pf;
    S.table.clear(); S.open.clear(); S.bad.clear(); S.suspect.clear(); S.live.clear(); S.armed.clear();
    S.have_dot = false;
}

// an n x 1 operator applied transposed into a 1-element view: y[0] = alpha * <a, x> + beta * y[0].  Returns true when the
// product was served from a prefetched value (y is then already written on host and device).
bool pf_try_dot(int dtype, const tb_view& a, const tb_view& x, const tb_view& y, double alpha, double beta, double* value_out) {
    PfState& S = g_pf;
    if (!S.on || y.len != 1 || a.len != x.len) return false;
    PfReq r;
    r.kind = 2; r.dtype = dtype; r.a = a; r.b = x; r.y = y;
    if (beta == 0.0) {
        if (PfLive* e = find_live(r)) {
            const double v = live_value(*e);
            e->valid = false;
            S.served += 1;
            S.have_dot = false;
            on_request(r, true);
            *value_out = v;
            return true;
        }
    }
    S.have_dot = beta == 0.0;
    S.dot = r;
    (void)alpha;
    return false;
}

// sum of squares of a view about to be reduced for the host: true + value when it was prefetched
bool pf_try_sumsq(int dtype, const tb_view& a, double* value_out) {
    PfState& S = g_pf;
    if (!S.on || a.len == 0) return false;
    PfReq r;
    r.kind = 1; r.dtype = dtype; r.a = a;
    if (PfLive* e = find_live(r)) {
        *value_out = live_value(*e);
        e->valid = false;
        S.served += 1;
        on_request(r, true);
        return true;
    }
    on_request(r, false);           // a miss: the caller reduces now; box_wait launches whatever followed last time
    return false;
}

// a single element is about to be fetched from the device (tb_get1 with a device-newer element)
void pf_miss_fetch(int dtype, const tb_view& elem) {
    PfState& S = g_pf;
    if (!S.on) return;
    PfReq r;
    if (S.have_dot && S.dot.dtype == dtype && veq(S.dot.y, elem)) {
        r = S.dot;                   // the element is the result of the dot just executed: a "dot + fetch" request
    } else {
        r.kind = 0; r.dtype = dtype; r.a = elem;
    }
    S.have_dot = false;
    on_request(r, false);
}

}  // namespace tb

using namespace tb;
extern "C" {
int tb_set_scalar_prefetch(int on) {
    return api([&] {
        require_init();
        g_pf.on = on != 0;
        pf_reset();
    });
}
int tb_scalar_prefetch_stats(uint64_t* launched, uint64_t* served, uint64_t* dropped) {
    return api_raw([&] { *launched = g_pf.launched; *served = g_pf.served; *dropped = g_pf.dropped; });
}
}
