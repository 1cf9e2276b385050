// Batched product-cone projection: Cone::proj / Cone::product_group (totsu_core/src/solver/cone.rs:9-30) for a
// product of ConeZero (cone_zero.rs:38-44), ConeRPos (cone_rpos.rs:38-45), ConeSOC (cone_soc.rs:38-65),
// ConeRotSOC (cone_rotsoc.rs:38-65) and ConePSD (cone_psd.rs:56-79) blocks laid out back to back, the shape of
// ProbLPCone / ProbQPCone / ProbQCQPCone / ProbSOCPCone / ProbSDPCone (totsu/src/problem/*.rs).
//
// The reference projects block by block from the host (one blocking norm + two scalar round trips per SOC
// block, a PCIe bounce of the whole vector for RPos).  Here the whole product cone is ONE launch: a device
// table of work items maps CTAs to (a) 8 small second-order blocks, one per warp, (b) one large second-order
// block per CTA, or (c) a slice of an element-wise (RPos / Zero) range.  PSD blocks are projected afterwards by
// the eigensolver in eig.cu.  Norms accumulate in double.
#include "common.cuh"

namespace tb {

constexpr int CN_THREADS = 256;
constexpr int CN_SMALL_MAX = 2048;       // second-order blocks up to this length are handled by one warp
constexpr int CN_ELEMS_PER_CTA = 4096;   // element-wise slice per CTA

struct ConeItem {
    int kind;            // 0: elementwise slice, 1: up to 8 small SOC-type blocks, 2: one large SOC-type block
    int type;            // kind 0: TB_CONE_ZERO/RPOS; kind 2: TB_CONE_SOC/ROTSOC
    unsigned long long off;   // kind 0/2: element offset;   kind 1: index of the first small block in `small`
    unsigned long long len;   // kind 0/2: element count;    kind 1: number of small blocks (<= 8)
};
struct SmallBlock {
    unsigned long long off;
    unsigned int len;
    int type;
};

struct ConeSet {
    std::vector<tb_cone_block> blocks;
    size_t total_len = 0;
    ConeItem* d_items = nullptr;
    SmallBlock* d_small = nullptr;
    int n_items = 0;
    bool has_psd = false;
    bool has_work = false;       // anything for the batched kernel to do
};

// --- second-order cone projection on a range handled by `nthreads` cooperating threads -------------------
// MODE 0: projection; MODE 1: group-min fill (product_group with the solver's min closure)
template <typename T, int MODE, bool WARP>
__device__ __forceinline__ void soc_block(T* x, unsigned long long len, int type, double* red) {
    const int lane = WARP ? (threadIdx.x & 31) : threadIdx.x;
    const int nth = WARP ? 32 : blockDim.x;
    if (len == 0) return;
    if (MODE == 1) {
        double mn = 1.0e300;
        for (unsigned long long i = lane; i < len; i += nth) mn = fmin(mn, (double)x[i]);
        mn = WARP ? tbd::warp_min(mn) : tbd::block_min(mn, red);
        for (unsigned long long i = lane; i < len; i += nth) x[i] = (T)mn;
        return;
    }
    if (type == TB_CONE_ROTSOC && len == 1) {              // cone_rotsoc.rs:46-49
        if (lane == 0) x[0] = x[0] > T(0) ? x[0] : T(0);
        return;
    }
    // s, and (for the rotated cone) the rotated second coordinate: cone_rotsoc.rs:51-54
    T x0 = x[0];
    T x1 = len > 1 ? x[1] : T(0);
    const T isq2 = (T)0.70710678118654752440;
    T s = x0, v1 = x1;
    if (type == TB_CONE_ROTSOC) {
        s = (x0 + x1) * isq2;
        v1 = (x0 - x1) * isq2;
    }
    // ||v||^2 with v = x[1..]  (cone_soc.rs:47-48), x[1] possibly replaced by its rotated value
    double acc = 0.0;
    for (unsigned long long i = 2 + lane; i < len; i += nth) { double v = (double)x[i]; acc += v * v; }
    double tot = WARP ? tbd::warp_sum(acc) : tbd::block_sum(acc, red);
    if (len > 1) tot += (double)v1 * (double)v1;
    const T norm_v = (T)sqrt(tot);
    T scale_v, new_s;
    if (norm_v <= -s) { scale_v = T(0); new_s = T(0); }                       // cone_soc.rs:50-53
    else if (norm_v <= s) { scale_v = T(1); new_s = s; }                      // :54-56
    else { scale_v = (T(1) + s / norm_v) / T(2); new_s = (norm_v + s) / T(2); }   // :57-61
    if (!WARP) __syncthreads();     // everybody has read x[0], x[1] before they are rewritten
    else __syncwarp();
    if (scale_v != T(1)) {
        if (scale_v == T(0)) { for (unsigned long long i = 2 + lane; i < len; i += nth) x[i] = T(0); }
        else { for (unsigned long long i = 2 + lane; i < len; i += nth) x[i] = scale_v * x[i]; }
    }
    if (lane == 0) {
        T nv1 = (scale_v == T(0)) ? T(0) : scale_v * v1;
        if (type == TB_CONE_ROTSOC) {                                          // rotate back: cone_rotsoc.rs:58-61
            x[0] = (new_s + nv1) * isq2;
            x[1] = (new_s - nv1) * isq2;
        } else {
            x[0] = new_s;
            if (len > 1) x[1] = nv1;
        }
    }
}

// blockIdx.y selects the vector: the two projections of one solver iteration (y block onto K*, s block onto K,
// solver.rs:548-549) run as ONE launch over the same item table (gridDim.y = 2); a single projection has gridDim.y = 1.
template <typename T, int MODE>
__global__ void __launch_bounds__(CN_THREADS) cone_kernel(T* x0, const ConeItem* __restrict__ items, const SmallBlock* __restrict__ small, int dual0,
                                                          T* x1 = nullptr, int dual1 = 0) {
    __shared__ double red[32];
    tbd::pdl_entry();
    T* x = blockIdx.y == 0 ? x0 : x1;
    const int dual_cone = blockIdx.y == 0 ? dual0 : dual1;
    const ConeItem it = items[blockIdx.x];
    if (it.kind == 0) {
        if (MODE == 1) return;                                 // Zero / RPos: product_group does nothing
        T* p = x + it.off;
        if (it.type == TB_CONE_RPOS) {                         // cone_rpos.rs:40-43
            for (unsigned long long i = threadIdx.x; i < it.len; i += blockDim.x) { T v = p[i]; p[i] = v > T(0) ? v : T(0); }
        } else if (!dual_cone) {                               // cone_zero.rs:40-42 (dual cone = free: untouched)
            for (unsigned long long i = threadIdx.x; i < it.len; i += blockDim.x) p[i] = T(0);
        }
    } else if (it.kind == 1) {
        const int w = threadIdx.x >> 5;
        if ((unsigned long long)w < it.len) {
            const SmallBlock b = small[it.off + w];
            soc_block<T, MODE, true>(x + b.off, b.len, b.type, red);
        }
    } else {
        soc_block<T, MODE, false>(x + it.off, it.len, it.type, red);
    }
}

static ConeSet& get_cone(tb_handle h) {
    Context& c = ctx();
    if (h <= 0 || (size_t)h > c.cones.size() || c.cones[(size_t)h - 1] == nullptr) fail(TB_ERR_ARG, "invalid cone handle");
    return *c.cones[(size_t)h - 1];
}

static void build_items(ConeSet& cs) {
    std::vector<ConeItem> items;
    std::vector<SmallBlock> small;
    size_t off = 0;
    size_t small_run_start = 0;
    auto flush_small = [&]() {
        for (size_t i = small_run_start; i < small.size(); i += 8) {
            ConeItem it;
            it.kind = 1; it.type = 0; it.off = i; it.len = std::min<size_t>(8, small.size() - i);
            items.push_back(it);
        }
        small_run_start = small.size();
    };
    for (const tb_cone_block& b : cs.blocks) {
        const size_t len = (size_t)b.len;
        switch (b.type) {
            case TB_CONE_ZERO:
            case TB_CONE_RPOS:
                for (size_t o = 0; o < len; o += CN_ELEMS_PER_CTA) {
                    ConeItem it;
                    it.kind = 0; it.type = b.type; it.off = off + o; it.len = std::min<size_t>(CN_ELEMS_PER_CTA, len - o);
                    items.push_back(it);
                }
                break;
            case TB_CONE_SOC:
            case TB_CONE_ROTSOC:
                if (len == 0) break;
                if (len <= CN_SMALL_MAX) {
                    small.push_back(SmallBlock{(unsigned long long)off, (unsigned int)len, b.type});
                } else {
                    ConeItem it;
                    it.kind = 2; it.type = b.type; it.off = off; it.len = len;
                    items.push_back(it);
                }
                break;
            case TB_CONE_PSD:
                cs.has_psd = true;
                break;
            default:
                fail(TB_ERR_ARG, "unknown cone block type");
        }
        off += len;
    }
    flush_small();
    cs.total_len = off;
    cs.n_items = (int)items.size();
    cs.has_work = !items.empty();
    if (!items.empty()) {
        TB_CUDA(cudaMalloc(&cs.d_items, items.size() * sizeof(ConeItem)));
        TB_CUDA(cudaMemcpy(cs.d_items, items.data(), items.size() * sizeof(ConeItem), cudaMemcpyHostToDevice));
    }
    if (!small.empty()) {
        TB_CUDA(cudaMalloc(&cs.d_small, small.size() * sizeof(SmallBlock)));
        TB_CUDA(cudaMemcpy(cs.d_small, small.data(), small.size() * sizeof(SmallBlock), cudaMemcpyHostToDevice));
    }
}

// Everything Cone::proj can reject is rejected HERE, at the call that the binding maps to Err(()) -> ConeFailure
// (solver.rs:548-549) - before a projection is parked, so that the error cannot surface on a later, unrelated call.
template <typename T> static void cone_validate(const ConeSet& cs, const tb_view& x, const tb_view& psd_work) {
    TB_REQUIRE(x.len == cs.total_len, "cone proj: vector length != sum of block lengths");
    if (x.len > 0) {
        const Buffer& bx = get_buf(x.buf);
        TB_REQUIRE(bx.dtype == DT<T>::id, "cone proj: view element type does not match the function suffix");
        TB_REQUIRE(x.off <= bx.len && x.len <= bx.len - x.off, "cone proj: view out of range");
    }
    if (!cs.has_psd) return;
    size_t need = 0;
    for (const tb_cone_block& b : cs.blocks) {
        if (b.type != TB_CONE_PSD || b.len == 0) continue;
        size_t k = 0;
        while ((k + 1) * (k + 2) / 2 <= (size_t)b.len) ++k;
        TB_REQUIRE(k * (k + 1) / 2 == (size_t)b.len, "cone proj: PSD block length is not a triangular number");     // cone_psd.rs:32-38
        need = std::max(need, (size_t)tb_map_eig_worklen(k));
    }
    TB_REQUIRE(psd_work.len >= need, "cone proj: PSD work slice too short (cone_psd.rs:60-63)");
    if (need > 0) {
        const Buffer& bw = get_buf(psd_work.buf);
        TB_REQUIRE(bw.dtype == DT<T>::id, "cone proj: work element type does not match the function suffix");
        TB_REQUIRE(psd_work.off <= bw.len && psd_work.len <= bw.len - psd_work.off, "cone proj: work view out of range");
    }
}

template <typename T> static void cone_proj(tb_handle h, int dual_cone, tb_view x, T eps_zero, tb_view psd_work) {
    require_init();
    ConeSet& cs = get_cone(h);
    cone_validate<T>(cs, x, psd_work);
    T* px = wptr<T>(x);
    Context& c = ctx();
    if (cs.has_work) {
        launch_pdl(cone_kernel<T, 0>, dim3(cs.n_items), dim3(CN_THREADS), 0, c.stream, px, (const ConeItem*)cs.d_items, (const SmallBlock*)cs.d_small, dual_cone,
                   (T*)nullptr, 0);
        TB_LAUNCH_CHECK();
    }
    if (cs.has_psd) {
        size_t off = 0;
        for (const tb_cone_block& b : cs.blocks) {
            if (b.type == TB_CONE_PSD && b.len > 0) {
                T* w = wptr<T>(psd_work);
                psd_project<T>(px + off, (size_t)b.len, eps_zero, w, psd_work.len);
            }
            off += (size_t)b.len;
        }
    }
}

// ---- pairing of the two projections of one iteration ----------------------------------------------------------
// The solver projects the y block onto K* and the s block onto K back to back (solver.rs:548-549).  The first call is parked;
// when the second arrives (same cone, same element type, disjoint vector, same work buffer) both run together: the
// Zero / RPos / SOC / RotSOC blocks of both vectors in ONE cone_kernel launch (gridDim.y = 2), the PSD blocks of both (f32) as
// ONE batch on the tensor cores (eig.cu:psd_project_pair).  Any other API call runs the parked projection first (api_raw checks
// g_cone_pending), so program order is preserved; arguments are validated before parking.  tb_set_psd_pairing(0) turns it off.
struct PendingProj {
    tb_handle h = 0;
    int dual = 0;
    int dtype = TB_F32;
    tb_view x{0, 0, 0}, w{0, 0, 0};
    double eps = 0.0;
};
static PendingProj g_pending;
bool g_cone_pending = false;

void cone_flush_pending() {
    if (!g_cone_pending) return;
    g_cone_pending = false;
    const PendingProj p = g_pending;
    if (p.dtype == TB_F32) cone_proj<float>(p.h, p.dual, p.x, (float)p.eps, p.w);
    else cone_proj<double>(p.h, p.dual, p.x, p.eps, p.w);
}

static void cone_flush_one(const PendingProj& p) {
    if (p.dtype == TB_F32) cone_proj<float>(p.h, p.dual, p.x, (float)p.eps, p.w);
    else cone_proj<double>(p.h, p.dual, p.x, p.eps, p.w);
}

static bool views_disjoint(const tb_view& a, const tb_view& b) {
    return a.buf != b.buf || a.off + a.len <= b.off || b.off + b.len <= a.off;
}

// both projections at once
template <typename T> static void cone_proj_pair(const PendingProj& p0, tb_handle h, int dual1, tb_view x1, T eps1, tb_view w) {
    ConeSet& cs = get_cone(h);
    Context& c = ctx();
    T* px0 = wptr<T>(p0.x);
    T* px1 = wptr<T>(x1);
    if (cs.has_work) {
        launch_pdl(cone_kernel<T, 0>, dim3(cs.n_items, 2), dim3(CN_THREADS), 0, c.stream, px0, (const ConeItem*)cs.d_items, (const SmallBlock*)cs.d_small, p0.dual,
                   px1, dual1);
        TB_LAUNCH_CHECK();
        c.cone_pairs += 1;
    }
    if (!cs.has_psd) return;
    T* pw = wptr<T>(w);
    size_t off = 0;
    for (const tb_cone_block& b : cs.blocks) {
        if (b.type == TB_CONE_PSD && b.len > 0) {
            bool batched = false;
            if constexpr (sizeof(T) == 4) {
                if (c.psd_mode == 0 && psd_pair_usable((size_t)b.len, pw) && ((reinterpret_cast<uintptr_t>(px0 + off) | reinterpret_cast<uintptr_t>(px1 + off)) & 3u) == 0) {
                    psd_project_pair(px0 + off, px1 + off, (size_t)b.len, pw, w.len);
                    c.psd_pairs += 1;
                    batched = true;
                }
            }
            if (!batched) {
                psd_project<T>(px0 + off, (size_t)b.len, (T)p0.eps, pw, w.len);
                psd_project<T>(px1 + off, (size_t)b.len, eps1, pw, w.len);
            }
        }
        off += (size_t)b.len;
    }
}

template <typename T> static void cone_proj_submit(tb_handle h, int dual, tb_view x, T eps, tb_view w) {
    require_init();
    if (!ctx().queue.empty()) queue_drain();
    ConeSet& cs = get_cone(h);
    cone_validate<T>(cs, x, w);
    if (g_cone_pending) {
        const PendingProj p0 = g_pending;
        const bool pairable = p0.h == h && p0.dtype == DT<T>::id && views_disjoint(p0.x, x) && p0.w.buf == w.buf && p0.w.off == w.off && p0.w.len == w.len &&
                              (w.len == 0 || (views_disjoint(p0.x, w) && views_disjoint(x, w)));
        g_cone_pending = false;
        if (pairable) {
            cone_proj_pair<T>(p0, h, dual, x, eps, w);
            return;
        }
        cone_flush_one(p0);
    }
    if ((cs.has_work || cs.has_psd) && ctx().psd_pairing) {
        g_pending.h = h; g_pending.dual = dual; g_pending.dtype = DT<T>::id; g_pending.x = x; g_pending.w = w; g_pending.eps = (double)eps;
        g_cone_pending = true;                 // parked: runs with its partner, or alone as soon as anything else is called
        return;
    }
    cone_proj<T>(h, dual, x, eps, w);
}

template <typename T> static void cone_group_min(tb_handle h, tb_view dp_tau) {
    require_init();
    ConeSet& cs = get_cone(h);
    TB_REQUIRE(dp_tau.len == cs.total_len, "cone group: vector length != sum of block lengths");
    T* px = wptr<T>(dp_tau);
    Context& c = ctx();
    if (cs.has_work) {
        cone_kernel<T, 1><<<cs.n_items, CN_THREADS, 0, c.stream>>>(px, cs.d_items, cs.d_small, 0);
        TB_LAUNCH_CHECK();
    }
    if (cs.has_psd) {
        // a PSD block is one group (cone_psd.rs:81-84): reuse the large-block path through a temporary item list
        size_t off = 0;
        std::vector<ConeItem> items;
        for (const tb_cone_block& b : cs.blocks) {
            if (b.type == TB_CONE_PSD && b.len > 0) {
                ConeItem it;
                it.kind = 2; it.type = TB_CONE_SOC; it.off = off; it.len = (size_t)b.len;
                items.push_back(it);
            }
            off += (size_t)b.len;
        }
        if (!items.empty()) {
            ConeItem* d = nullptr;
            TB_CUDA(cudaMalloc(&d, items.size() * sizeof(ConeItem)));
            TB_CUDA(cudaMemcpyAsync(d, items.data(), items.size() * sizeof(ConeItem), cudaMemcpyHostToDevice, c.stream));
            cone_kernel<T, 1><<<(unsigned)items.size(), CN_THREADS, 0, c.stream>>>(px, d, nullptr, 0);
            TB_LAUNCH_CHECK();
            TB_CUDA(cudaStreamSynchronize(c.stream));
            TB_CUDA(cudaFree(d));
        }
    }
}

}  // namespace tb

using namespace tb;
extern "C" {

int tb_cone_create(const tb_cone_block* blocks, size_t n_blocks, tb_handle* out) {
    return api([&] {
        require_init();
        ConeSet* cs = new ConeSet();
        cs->blocks.assign(blocks, blocks + n_blocks);
        try {
            build_items(*cs);
        } catch (...) {
            delete cs;
            throw;
        }
        ctx().cones.push_back(cs);
        *out = (tb_handle)ctx().cones.size();
    });
}
int tb_cone_destroy(tb_handle h) {
    return api([&] {
        ConeSet& cs = get_cone(h);
        cudaStreamSynchronize(ctx().stream);
        if (cs.d_items) cudaFree(cs.d_items);
        if (cs.d_small) cudaFree(cs.d_small);
        delete &cs;
        ctx().cones[(size_t)h - 1] = nullptr;
    });
}
int tb_cone_proj_f32(tb_handle cone, int dual, tb_view x, float eps, tb_view w) { return api_keep_pending([&] { cone_proj_submit<float>(cone, dual, x, eps, w); }); }
int tb_set_psd_pairing(int on) {
    return api([&] { ctx().psd_pairing = on != 0; });
}
int tb_psd_pairs(uint64_t* out) {
    return api_raw([&] { *out = ctx().psd_pairs; });
}
int tb_cone_pairs(uint64_t* out) {
    return api_raw([&] { *out = ctx().cone_pairs; });
}
int tb_cone_proj_f64(tb_handle cone, int dual, tb_view x, double eps, tb_view w) { return api_keep_pending([&] { cone_proj_submit<double>(cone, dual, x, eps, w); }); }
int tb_cone_group_min_f32(tb_handle cone, tb_view t) { return api([&] { cone_group_min<float>(cone, t); }); }
int tb_cone_group_min_f64(tb_handle cone, tb_view t) { return api([&] { cone_group_min<double>(cone, t); }); }
}
