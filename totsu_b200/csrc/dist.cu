// Multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch (north_star).
// The reference has no collectives at all (single device, totsu_f32cuda/src/cuda_mgr.rs:37-39).  NCCL is bound
// at run time with dlopen so the library has no link-time dependency on a particular libnccl: inside a
// torchrun-launched process this resolves to the libnccl.so.2 torch already loaded.
#include "common.cuh"
#include <cstdlib>
#include <dlfcn.h>
#include <nccl.h>

namespace tb {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
};
static NcclApi g_nccl;

static void load_nccl() {
    if (g_nccl.lib) return;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) fail(TB_ERR_NCCL, std::string("cannot dlopen libnccl.so.2: ") + dlerror());
    auto sym = [&](const char* s) {
        void* p = dlsym(g_nccl.lib, s);
        if (!p) fail(TB_ERR_NCCL, std::string("libnccl is missing symbol ") + s);
        return p;
    };
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(sym("ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(sym("ncclCommInitRank"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(sym("ncclCommDestroy"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(sym("ncclGetErrorString"));
    g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(sym("ncclAllReduce"));
    g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(sym("ncclAllGather"));
}

#define TB_NCCL(expr)                                                                                          \
    do {                                                                                                       \
        ncclResult_t _r = (expr);                                                                              \
        if (_r != ncclSuccess) fail(TB_ERR_NCCL, std::string(#expr) + ": " + g_nccl.GetErrorString(_r));       \
    } while (0)

// ---------------------------------------------------------------------------------------------------------
// Peer exchange: the two collectives of the row-sharded operator fused into the matvec epilogue over NVLink.
//
// Every rank owns one "region" (cudaMalloc'ed, exported with cudaIpcGetMemHandle, mapped by every peer):
//     [ flags: uint64[world] | pad to 4 KB | stage parity 0 | stage parity 1 ]
// All ranks issue the same sequence of collectives; collective number `seq` uses stage parity seq&1.
//   push  : the epilogue kernel that turns matvec partials into values STORES them straight into every peer's
//           stage (plain st.global on the peer-mapped address = NVLink writes), fences at system scope, and the
//           last CTA releases flag[rank] = seq in every peer's region;
//   wait  : a second kernel acquires flag[src] >= seq for every src and then either copies the gathered slices
//           into y (A*x: all-gather) or sums the `world` partial vectors in rank order and applies alpha/beta
//           (A^T*y: all-reduce).  The fixed order makes the result bit-identical on every rank and run to run.
// Double buffering is sufficient: a peer can only push collective s+2 after it has seen my flag for s+1, and my
// push of s+1 is stream-ordered after my wait of s, so nobody overwrites a stage that is still being read.
// A rank that never arrives would hang the spin loop, so the wait gives up after kSpinTimeoutNs and raises a
// fault word in mapped host memory that the next host-visible call turns into TB_ERR_NCCL.
// NCCL (ncclAllGather / ncclAllReduce) stays as the baseline path: TB_P2P=0, or messages larger than a stage.
// ---------------------------------------------------------------------------------------------------------
constexpr int kMaxWorld = 16;
constexpr size_t kFlagBytes = 4096;
constexpr unsigned long long kSpinTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

struct PeerPtrs {
    char* region[kMaxWorld];
    int world, rank;
    size_t stage_bytes;
};

struct PeerExchange {
    bool on = false;
    PeerPtrs pp{};
    char* local = nullptr;
    uint64_t seq = 0;
    int* fault_host = nullptr;      // mapped pinned
    int* fault_dev = nullptr;
    void* tmp = nullptr;            // NCCL-path staging for all-reduce partials
    size_t tmp_bytes = 0;
    uint64_t exchanges = 0;         // peer exchanges (push + wait) since tb_dist_init
};
static PeerExchange g_px;

__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
template <typename T> __device__ __forceinline__ T* stage_ptr(const PeerPtrs& pp, int p, uint64_t seq) {
    return reinterpret_cast<T*>(pp.region[p] + kFlagBytes + (size_t)(seq & 1) * pp.stage_bytes);
}

// MODE 0 (gather): element i of my slice -> my y_local[i] and every peer's stage[stage_off + i]
// MODE 1 (reduce): element i of my partial -> every rank's stage[rank*len + i] (mine included: one summation order)
// FROM_PARTS: value = sum_j src[j*ld + i] (the matvec partials, epilogue fusion), else value = src[i]
template <typename T, int MODE, bool FROM_PARTS>
__global__ void __launch_bounds__(256) peer_push_kernel(const PeerPtrs pp, uint64_t seq, const T* __restrict__ src, int nparts, size_t ld, size_t len,
                                                        T alpha, T beta, T* y_local, size_t stage_off, unsigned int* ticket) {
    __shared__ bool last;
    tbd::pdl_entry();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) {
        T v;
        if (FROM_PARTS) {
            T s = T(0);
            for (int j = 0; j < nparts; ++j) s += src[(size_t)j * ld + i];
            v = s;
        } else {
            v = src[i];
        }
        if (MODE == 0) {
            if (FROM_PARTS) {
                v = alpha * v;
                if (beta != T(0)) v += beta * y_local[i];
                y_local[i] = v;
            }
            for (int p = 0; p < pp.world; ++p)
                if (p != pp.rank) stage_ptr<T>(pp, p, seq)[stage_off + i] = v;
        } else {
            for (int p = 0; p < pp.world; ++p) stage_ptr<T>(pp, p, seq)[(size_t)pp.rank * len + i] = v;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicInc(ticket, gridDim.x - 1);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x < pp.world) {
        __threadfence_system();
        const int p = threadIdx.x;
        if (MODE == 1 || p != pp.rank) st_release_sys(reinterpret_cast<uint64_t*>(pp.region[p]) + pp.rank, seq);
    }
}

// MODE 0: y[r*count + i] = stage[r*count + i] for r != rank;  MODE 1: y[i] = alpha * sum_r stage[r*count + i] + beta*y[i]
template <typename T, int MODE>
__global__ void __launch_bounds__(256) peer_wait_kernel(const PeerPtrs pp, uint64_t seq, size_t count, T alpha, T beta, T* y, int* fault) {
    tbd::pdl_entry();
    if (threadIdx.x < pp.world && (MODE == 1 || (int)threadIdx.x != pp.rank)) {
        const uint64_t* flag = reinterpret_cast<const uint64_t*>(pp.region[pp.rank]) + threadIdx.x;
        if (ld_acquire_sys(flag) < seq) {
            const unsigned long long t0 = global_ns();
            while (ld_acquire_sys(flag) < seq) {
                if (global_ns() - t0 > kSpinTimeoutNs) { *fault = 1; break; }
            }
        }
    }
    __syncthreads();
    const T* st = stage_ptr<T>(pp, pp.rank, seq);
    if (MODE == 0) {
        const size_t total = count * (size_t)pp.world;
        const size_t lo = count * (size_t)pp.rank, hi = lo + count;
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
            if (i < lo || i >= hi) y[i] = __ldcg(st + i);
    } else {
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
            T s = T(0);
            for (int r = 0; r < pp.world; ++r) s += __ldcg(st + (size_t)r * count + i);
            T v = alpha * s;
            if (beta != T(0)) v += beta * y[i];
            y[i] = v;
        }
    }
}

// Both collectives of one op/trans_op pair in ONE exchange: the push kernel finalizes my slice of A x (alpha/beta applied,
// stored into my y and every peer's stage) AND stores my partial of A^T x into every rank's stage behind the gather area,
// then releases one flag per peer; the wait kernel acquires the flags once, copies the gathered slices and sums the reduce
// partials in rank order.  Halves the launches and the flag round trips of a pair (6 exchanges per solver iteration -> 3).
//
// A SECOND, speculated pair can ride in the same exchange (len_g2 / len_r2 > 0): the raw products of the pair that will be
// asked for next (gemv.cu "speculative pairing": its alpha, beta and output views are not known yet) are summed per rank,
// pushed behind the first pair's areas and left - gathered / reduced in rank order, un-scaled - in local buffers, so that
// when the pair arrives it is served by a local axpby: 3 exchanges per solver iteration -> 2.
// Stage layout: [world*len_g gathered | world*len_r partials | world*len_g2 raw gathered | world*len_r2 raw partials].
template <typename T> struct PairSeg {
    const T* part_g; int nparts_g; size_t ld_g, len_g; T alpha_g, beta_g; T* y_local;     // gather of the carrying pair
    const T* part_r; int nparts_r; size_t ld_r, len_r;                                    // reduce of the carrying pair
    const T* part_g2; int nparts_g2; size_t ld_g2, len_g2;                                // speculated pair: raw gather
    const T* part_r2; int nparts_r2; size_t ld_r2, len_r2;                                // speculated pair: raw reduce
};

template <typename T>
__global__ void __launch_bounds__(256) peer_push_pair_kernel(const PeerPtrs pp, uint64_t seq, const PairSeg<T> g, unsigned int* ticket) {
    __shared__ bool last;
    tbd::pdl_entry();
    const size_t goff = (size_t)pp.rank * g.len_g;
    const size_t rbase = (size_t)pp.world * g.len_g;
    const size_t g2base = rbase + (size_t)pp.world * g.len_r;
    const size_t r2base = g2base + (size_t)pp.world * g.len_g2;
    const size_t e1 = g.len_g, e2 = e1 + g.len_r, e3 = e2 + g.len_g2, e4 = e3 + g.len_r2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < e4; i += (size_t)gridDim.x * blockDim.x) {
        if (i < e1) {
            T s = T(0);
            for (int j = 0; j < g.nparts_g; ++j) s += g.part_g[(size_t)j * g.ld_g + i];
            T v = g.alpha_g * s;
            if (g.beta_g != T(0)) v += g.beta_g * g.y_local[i];
            g.y_local[i] = v;
            for (int p = 0; p < pp.world; ++p)
                if (p != pp.rank) stage_ptr<T>(pp, p, seq)[goff + i] = v;
        } else if (i < e2) {
            const size_t k = i - e1;
            T s = T(0);
            for (int j = 0; j < g.nparts_r; ++j) s += g.part_r[(size_t)j * g.ld_r + k];
            for (int p = 0; p < pp.world; ++p) stage_ptr<T>(pp, p, seq)[rbase + (size_t)pp.rank * g.len_r + k] = s;
        } else if (i < e3) {
            const size_t k = i - e2;
            T s = T(0);
            for (int j = 0; j < g.nparts_g2; ++j) s += g.part_g2[(size_t)j * g.ld_g2 + k];
            for (int p = 0; p < pp.world; ++p) stage_ptr<T>(pp, p, seq)[g2base + (size_t)pp.rank * g.len_g2 + k] = s;
        } else {
            const size_t k = i - e3;
            T s = T(0);
            for (int j = 0; j < g.nparts_r2; ++j) s += g.part_r2[(size_t)j * g.ld_r2 + k];
            for (int p = 0; p < pp.world; ++p) stage_ptr<T>(pp, p, seq)[r2base + (size_t)pp.rank * g.len_r2 + k] = s;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicInc(ticket, gridDim.x - 1);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x < pp.world) {
        __threadfence_system();
        const int p = threadIdx.x;
        st_release_sys(reinterpret_cast<uint64_t*>(pp.region[p]) + pp.rank, seq);      // my own flag too: the reduce reads my partial from my stage
    }
}

template <typename T>
__global__ void __launch_bounds__(256) peer_wait_pair_kernel(const PeerPtrs pp, uint64_t seq, size_t len_g, T* y_g_base,
                                                             size_t len_r, T alpha_r, T beta_r, T* y_r,
                                                             size_t len_g2, T* raw_g2, size_t len_r2, T* raw_r2, int* fault) {
    tbd::pdl_entry();
    if (threadIdx.x < pp.world) {
        const uint64_t* flag = reinterpret_cast<const uint64_t*>(pp.region[pp.rank]) + threadIdx.x;
        if (ld_acquire_sys(flag) < seq) {
            const unsigned long long t0 = global_ns();
            while (ld_acquire_sys(flag) < seq) {
                if (global_ns() - t0 > kSpinTimeoutNs) { *fault = 1; break; }
            }
        }
    }
    __syncthreads();
    const T* st = stage_ptr<T>(pp, pp.rank, seq);
    const size_t total_g = len_g * (size_t)pp.world;
    const size_t lo = len_g * (size_t)pp.rank, hi = lo + len_g;
    const size_t rbase = total_g;
    const size_t g2base = rbase + (size_t)pp.world * len_r;
    const size_t total_g2 = len_g2 * (size_t)pp.world;
    const size_t r2base = g2base + total_g2;
    const size_t e1 = total_g, e2 = e1 + len_r, e3 = e2 + total_g2, e4 = e3 + len_r2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < e4; i += (size_t)gridDim.x * blockDim.x) {
        if (i < e1) {
            if (i < lo || i >= hi) y_g_base[i] = __ldcg(st + i);
        } else if (i < e2) {
            const size_t k = i - e1;
            T s = T(0);
            for (int r = 0; r < pp.world; ++r) s += __ldcg(st + rbase + (size_t)r * len_r + k);
            T v = alpha_r * s;
            if (beta_r != T(0)) v += beta_r * y_r[k];
            y_r[k] = v;
        } else if (i < e3) {
            const size_t k = i - e2;
            raw_g2[k] = __ldcg(st + g2base + k);
        } else {
            const size_t k = i - e3;
            T s = T(0);
            for (int r = 0; r < pp.world; ++r) s += __ldcg(st + r2base + (size_t)r * len_r2 + k);
            raw_r2[k] = s;
        }
    }
}

static inline int px_grid(size_t len) {
    return (int)std::max<size_t>(1, std::min<size_t>((len + 255) / 256, (size_t)ctx().sm_count * 2));
}
static inline unsigned int* px_ticket() { return ctx().tickets + 48; }

void dist_check_fault() {
    if (g_px.fault_host && *g_px.fault_host) fail(TB_ERR_NCCL, "peer exchange: a rank did not arrive within the spin timeout");
}

static void* px_tmp(size_t bytes) {
    if (bytes > g_px.tmp_bytes) {
        TB_CUDA(cudaStreamSynchronize(ctx().stream));
        if (g_px.tmp) TB_CUDA(cudaFree(g_px.tmp));
        TB_CUDA(cudaMalloc(&g_px.tmp, bytes));
        g_px.tmp_bytes = bytes;
    }
    return g_px.tmp;
}

template <typename T> static bool px_fits(size_t elems) { return g_px.on && elems * sizeof(T) <= g_px.pp.stage_bytes; }

template <typename T> static void px_gather(const T* src, bool from_parts, int nparts, size_t ld, size_t len_local, T alpha, T beta, T* y_base) {
    Context& c = ctx();
    const uint64_t seq = ++g_px.seq;
    g_px.exchanges += 1;
    T* y_local = y_base + (size_t)c.rank * len_local;
    const size_t off = (size_t)c.rank * len_local;
    if (from_parts)
        launch_pdl(peer_push_kernel<T, 0, true>, dim3(px_grid(len_local)), dim3(256), 0, c.stream, g_px.pp, seq, src, nparts, ld, len_local, alpha, beta, y_local, off, px_ticket());
    else
        launch_pdl(peer_push_kernel<T, 0, false>, dim3(px_grid(len_local)), dim3(256), 0, c.stream, g_px.pp, seq, src, 0, (size_t)0, len_local, T(1), T(0), y_local, off, px_ticket());
    TB_LAUNCH_CHECK();
    launch_pdl(peer_wait_kernel<T, 0>, dim3(px_grid(len_local * c.world)), dim3(256), 0, c.stream, g_px.pp, seq, len_local, T(1), T(0), y_base, g_px.fault_dev);
    TB_LAUNCH_CHECK();
}

template <typename T> static void px_reduce(const T* src, bool from_parts, int nparts, size_t ld, size_t n, T alpha, T beta, T* y) {
    Context& c = ctx();
    const uint64_t seq = ++g_px.seq;
    g_px.exchanges += 1;
    if (from_parts)
        launch_pdl(peer_push_kernel<T, 1, true>, dim3(px_grid(n)), dim3(256), 0, c.stream, g_px.pp, seq, src, nparts, ld, n, T(1), T(0), (T*)nullptr, (size_t)0, px_ticket());
    else
        launch_pdl(peer_push_kernel<T, 1, false>, dim3(px_grid(n)), dim3(256), 0, c.stream, g_px.pp, seq, src, 0, (size_t)0, n, T(1), T(0), (T*)nullptr, (size_t)0, px_ticket());
    TB_LAUNCH_CHECK();
    launch_pdl(peer_wait_kernel<T, 1>, dim3(px_grid(n)), dim3(256), 0, c.stream, g_px.pp, seq, n, alpha, beta, y, g_px.fault_dev);
    TB_LAUNCH_CHECK();
}

template <typename T> static ncclDataType_t nccl_type() { return sizeof(T) == 4 ? ncclFloat32 : ncclFloat64; }

// y_base[rank*len .. +len) = alpha * sum_j part[j*ld + i] + beta * (same slice); then every rank holds all slices
template <typename T> void dist_finalize_gather(const T* part, int nparts, size_t ld, size_t len_local, T alpha, T beta, T* y_base) {
    Context& c = ctx();
    if (len_local == 0) return;
    if (c.world > 1 && px_fits<T>(len_local * (size_t)c.world)) {
        px_gather<T>(part, true, nparts, ld, len_local, alpha, beta, y_base);
        return;
    }
    T* y_local = y_base + (size_t)c.rank * len_local;
    l2_finalize<T>(part, nparts, ld, len_local, alpha, beta, y_local);
    if (c.world > 1) {
        TB_NCCL(g_nccl.AllGather(y_local, y_base, len_local, nccl_type<T>(), (ncclComm_t)c.nccl_comm, c.stream));
        count_launch();
    }
}

// y = alpha * sum_ranks sum_j part[j*ld + i] + beta*y
template <typename T> void dist_finalize_reduce(const T* part, int nparts, size_t ld, size_t n, T alpha, T beta, T* y) {
    Context& c = ctx();
    if (n == 0) return;
    if (c.world > 1 && px_fits<T>(n * (size_t)c.world)) {
        px_reduce<T>(part, true, nparts, ld, n, alpha, beta, y);
        return;
    }
    if (c.world <= 1) {
        l2_finalize<T>(part, nparts, ld, n, alpha, beta, y);
        return;
    }
    T* tmp = reinterpret_cast<T*>(px_tmp(n * sizeof(T)));
    l2_finalize<T>(part, nparts, ld, n, T(1), T(0), tmp);
    TB_NCCL(g_nccl.AllReduce(tmp, tmp, n, nccl_type<T>(), ncclSum, (ncclComm_t)c.nccl_comm, c.stream));
    count_launch();
    l1_axpby<T>(alpha, tmp, beta, y, n);
}
// the two epilogues of a pair; one exchange when the peer path can take both (TB_P2P_PAIR=0 keeps them separate)
template <typename T>
void dist_finalize_pair(const T* part_n, int nparts_n, size_t ld_n, size_t len_local, T alpha_n, T beta_n, T* y_base,
                        const T* part_t, int nparts_t, size_t ld_t, size_t n, T alpha_t, T beta_t, T* y_t) {
    Context& c = ctx();
    static const bool fused = [] { const char* e = getenv("TB_P2P_PAIR"); return !(e && e[0] == '0'); }();
    if (fused && c.world > 1 && len_local > 0 && n > 0 && px_fits<T>((len_local + n) * (size_t)c.world)) {
        const uint64_t seq = ++g_px.seq;
    g_px.exchanges += 1;
        T* y_local = y_base + (size_t)c.rank * len_local;
        PairSeg<T> g{part_n, nparts_n, ld_n, len_local, alpha_n, beta_n, y_local, part_t, nparts_t, ld_t, n, nullptr, 0, 0, 0, nullptr, 0, 0, 0};
        launch_pdl(peer_push_pair_kernel<T>, dim3(px_grid(len_local + n)), dim3(256), 0, c.stream, g_px.pp, seq, g, px_ticket());
        TB_LAUNCH_CHECK();
        launch_pdl(peer_wait_pair_kernel<T>, dim3(px_grid(len_local * c.world + n)), dim3(256), 0, c.stream, g_px.pp, seq, len_local, y_base, n, alpha_t, beta_t, y_t,
                   (size_t)0, (T*)nullptr, (size_t)0, (T*)nullptr, g_px.fault_dev);
        TB_LAUNCH_CHECK();
        return;
    }
    dist_finalize_gather<T>(part_n, nparts_n, ld_n, len_local, alpha_n, beta_n, y_base);
    dist_finalize_reduce<T>(part_t, nparts_t, ld_t, n, alpha_t, beta_t, y_t);
}
// The carrying pair's epilogues AND the speculated pair's raw products in ONE exchange (see PairSeg).  raw_n (world*len_local)
// and raw_t (n) receive the speculated pair's gathered / reduced sums, un-scaled.  Returns false - nothing done - when the peer
// path cannot take it (NCCL baseline, stage too small, TB_P2P_SPEC=0): the caller then finalizes the carrying pair on its own
// and exchanges the speculated one when it is asked for.
template <typename T>
bool dist_finalize_pair_spec(const T* part_n, int nparts_n, size_t ld_n, size_t len_local, T alpha_n, T beta_n, T* y_base,
                             const T* part_t, int nparts_t, size_t ld_t, size_t n, T alpha_t, T beta_t, T* y_t,
                             const T* spec_n, const T* spec_t, T* raw_n, T* raw_t) {
    Context& c = ctx();
    static const bool on = [] {
        const char* e = getenv("TB_P2P_SPEC");
        const char* f = getenv("TB_P2P_PAIR");
        return !(e && e[0] == '0') && !(f && f[0] == '0');
    }();
    if (!on || c.world <= 1 || len_local == 0 || n == 0 || !px_fits<T>(2 * (len_local + n) * (size_t)c.world)) return false;
    const uint64_t seq = ++g_px.seq;
    g_px.exchanges += 1;
    T* y_local = y_base + (size_t)c.rank * len_local;
    PairSeg<T> g{part_n, nparts_n, ld_n, len_local, alpha_n, beta_n, y_local, part_t, nparts_t, ld_t, n,
                 spec_n, nparts_n, ld_n, len_local, spec_t, nparts_t, ld_t, n};
    launch_pdl(peer_push_pair_kernel<T>, dim3(px_grid(2 * (len_local + n))), dim3(256), 0, c.stream, g_px.pp, seq, g, px_ticket());
    TB_LAUNCH_CHECK();
    launch_pdl(peer_wait_pair_kernel<T>, dim3(px_grid(2 * (len_local * c.world + n))), dim3(256), 0, c.stream, g_px.pp, seq, len_local, y_base, n, alpha_t, beta_t, y_t,
               len_local, raw_n, n, raw_t, g_px.fault_dev);
    TB_LAUNCH_CHECK();
    return true;
}
template bool dist_finalize_pair_spec<float>(const float*, int, size_t, size_t, float, float, float*, const float*, int, size_t, size_t, float, float, float*,
                                             const float*, const float*, float*, float*);
template bool dist_finalize_pair_spec<double>(const double*, int, size_t, size_t, double, double, double*, const double*, int, size_t, size_t, double, double, double*,
                                              const double*, const double*, double*, double*);
template void dist_finalize_pair<float>(const float*, int, size_t, size_t, float, float, float*, const float*, int, size_t, size_t, float, float, float*);
template void dist_finalize_pair<double>(const double*, int, size_t, size_t, double, double, double*, const double*, int, size_t, size_t, double, double, double*);
template void dist_finalize_gather<float>(const float*, int, size_t, size_t, float, float, float*);
template void dist_finalize_gather<double>(const double*, int, size_t, size_t, double, double, double*);
template void dist_finalize_reduce<float>(const float*, int, size_t, size_t, float, float, float*);
template void dist_finalize_reduce<double>(const double*, int, size_t, size_t, double, double, double*);

void dist_allreduce_sum(void* buf, size_t count, int dtype) {
    Context& c = ctx();
    if (c.world <= 1 || count == 0) return;
    if (dtype == TB_F32 && px_fits<float>(count * (size_t)c.world)) {
        px_reduce<float>(reinterpret_cast<const float*>(buf), false, 0, 0, count, 1.f, 0.f, reinterpret_cast<float*>(buf));
        return;
    }
    if (dtype == TB_F64 && px_fits<double>(count * (size_t)c.world)) {
        px_reduce<double>(reinterpret_cast<const double*>(buf), false, 0, 0, count, 1.0, 0.0, reinterpret_cast<double*>(buf));
        return;
    }
    TB_NCCL(g_nccl.AllReduce(buf, buf, count, dtype == TB_F32 ? ncclFloat32 : ncclFloat64, ncclSum, (ncclComm_t)c.nccl_comm, c.stream));
    count_launch();
}

void dist_allgather_inplace(void* base, size_t count_per_rank, int dtype) {
    Context& c = ctx();
    if (c.world <= 1 || count_per_rank == 0) return;
    if (dtype == TB_F32 && px_fits<float>(count_per_rank * (size_t)c.world)) {
        float* b = reinterpret_cast<float*>(base);
        px_gather<float>(b + (size_t)c.rank * count_per_rank, false, 0, 0, count_per_rank, 1.f, 0.f, b);
        return;
    }
    if (dtype == TB_F64 && px_fits<double>(count_per_rank * (size_t)c.world)) {
        double* b = reinterpret_cast<double*>(base);
        px_gather<double>(b + (size_t)c.rank * count_per_rank, false, 0, 0, count_per_rank, 1.0, 0.0, b);
        return;
    }
    const size_t es = dtype == TB_F32 ? 4 : 8;
    const char* send = reinterpret_cast<const char*>(base) + (size_t)c.rank * count_per_rank * es;
    TB_NCCL(g_nccl.AllGather(send, base, count_per_rank, dtype == TB_F32 ? ncclFloat32 : ncclFloat64, (ncclComm_t)c.nccl_comm, c.stream));
    count_launch();
}

// Map every peer's region.  Collective: all ranks call it, and all ranks agree on the outcome.
static void px_setup() {
    Context& c = ctx();
    const char* env = std::getenv("TB_P2P");
    int want = (env && std::atoi(env) == 0) ? 0 : 1;
    size_t stage_mb = 16;
    if (const char* e = std::getenv("TB_P2P_STAGE_MB")) stage_mb = (size_t)std::max(1, std::atoi(e));
    if (c.world > kMaxWorld) want = 0;
    g_px = PeerExchange();
    g_px.pp.world = c.world; g_px.pp.rank = c.rank; g_px.pp.stage_bytes = stage_mb << 20;
    const size_t region_bytes = kFlagBytes + 2 * g_px.pp.stage_bytes;
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof(mine));
    if (want) {
        if (cudaMalloc(&g_px.local, region_bytes) != cudaSuccess) { want = 0; cudaGetLastError(); g_px.local = nullptr; }
    }
    if (want) {
        TB_CUDA(cudaMemsetAsync(g_px.local, 0, region_bytes, c.stream));
        if (cudaIpcGetMemHandle(&mine, g_px.local) != cudaSuccess) { want = 0; cudaGetLastError(); }
    }
    // exchange (ok flag, handle) records with the NCCL communicator that already exists
    struct Rec { int ok; int pad; cudaIpcMemHandle_t h; };
    std::vector<Rec> all((size_t)c.world);
    Rec me; me.ok = want; me.pad = 0; me.h = mine;
    Rec* d = nullptr;
    TB_CUDA(cudaMalloc(&d, sizeof(Rec) * (size_t)c.world));
    TB_CUDA(cudaMemcpyAsync(d + c.rank, &me, sizeof(Rec), cudaMemcpyHostToDevice, c.stream));
    TB_NCCL(g_nccl.AllGather(d + c.rank, d, sizeof(Rec), ncclChar, (ncclComm_t)c.nccl_comm, c.stream));
    TB_CUDA(cudaMemcpyAsync(all.data(), d, sizeof(Rec) * (size_t)c.world, cudaMemcpyDeviceToHost, c.stream));
    TB_CUDA(cudaStreamSynchronize(c.stream));
    int ok = 1;
    for (const Rec& r : all) ok &= r.ok;
    if (ok) {
        for (int p = 0; p < c.world && ok; ++p) {
            if (p == c.rank) { g_px.pp.region[p] = g_px.local; continue; }
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[(size_t)p].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); }
            g_px.pp.region[p] = reinterpret_cast<char*>(ptr);
        }
    }
    // second round: did every rank manage to map every peer?
    int* dflag = reinterpret_cast<int*>(d);
    TB_CUDA(cudaMemcpyAsync(dflag, &ok, sizeof(int), cudaMemcpyHostToDevice, c.stream));
    TB_NCCL(g_nccl.AllReduce(dflag, dflag, 1, ncclInt32, ncclMin, (ncclComm_t)c.nccl_comm, c.stream));
    TB_CUDA(cudaMemcpyAsync(&ok, dflag, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    TB_CUDA(cudaStreamSynchronize(c.stream));
    TB_CUDA(cudaFree(d));
    if (ok) {
        TB_CUDA(cudaHostAlloc(&g_px.fault_host, sizeof(int), cudaHostAllocMapped));
        *g_px.fault_host = 0;
        TB_CUDA(cudaHostGetDevicePointer(&g_px.fault_dev, g_px.fault_host, 0));
        g_px.on = true;
    } else {
        for (int p = 0; p < c.world; ++p)
            if (p != c.rank && g_px.pp.region[p]) cudaIpcCloseMemHandle(g_px.pp.region[p]);
        if (g_px.local) cudaFree(g_px.local);
        g_px = PeerExchange();
    }
}

static void px_teardown() {
    Context& c = ctx();
    if (g_px.on) {
        // nobody may still be pushing into a region that is about to be unmapped: barrier first
        int* d = nullptr;
        if (cudaMalloc(&d, sizeof(int)) == cudaSuccess) {
            cudaMemsetAsync(d, 0, sizeof(int), c.stream);
            g_nccl.AllReduce(d, d, 1, ncclInt32, ncclSum, (ncclComm_t)c.nccl_comm, c.stream);
            cudaStreamSynchronize(c.stream);
            cudaFree(d);
        }
        for (int p = 0; p < c.world; ++p)
            if (p != c.rank && g_px.pp.region[p]) cudaIpcCloseMemHandle(g_px.pp.region[p]);
        cudaFree(g_px.local);
        cudaFreeHost(g_px.fault_host);
    }
    if (g_px.tmp) cudaFree(g_px.tmp);
    g_px = PeerExchange();
}

}  // namespace tb


using namespace tb;
extern "C" {

int tb_dist_unique_id(void* id_out) {
    return api([&] {
        static_assert(TB_NCCL_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
        load_nccl();
        ncclUniqueId id;
        TB_NCCL(g_nccl.GetUniqueId(&id));
        std::memcpy(id_out, &id, sizeof(id));
    });
}

int tb_dist_init(int rank, int world, const void* id_bytes) {
    return api([&] {
        require_init();
        Context& c = ctx();
        TB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank/world");
        if (c.nccl_comm) fail(TB_ERR_STATE, "tb_dist_init called twice");
        if (world == 1) { c.rank = 0; c.world = 1; return; }
        load_nccl();
        ncclUniqueId id;
        std::memcpy(&id, id_bytes, sizeof(id));
        ncclComm_t comm;
        TB_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
        c.nccl_comm = comm;
        c.rank = rank;
        c.world = world;
        px_setup();
    });
}

int tb_dist_finalize(void) {
    return api([&] {
        Context& c = ctx();
        if (c.nccl_comm) {
            cudaStreamSynchronize(c.stream);
            px_teardown();
            g_nccl.CommDestroy((ncclComm_t)c.nccl_comm);
            c.nccl_comm = nullptr;
        }
        c.rank = 0;
        c.world = 1;
    });
}

int tb_dist_p2p_enabled(int* out) {
    return api([&] { *out = g_px.on ? 1 : 0; });
}

int tb_dist_exchanges(uint64_t* out) {
    return api_raw([&] { *out = g_px.exchanges; });
}

int tb_dist_info(int* rank, int* world) {
    return api([&] {
        *rank = ctx().rank;
        *world = ctx().world;
    });
}
}
