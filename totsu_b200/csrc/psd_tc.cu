// Tensor-core engine of the ConePSD projection (totsu_core/src/cone_psd.rs:56-79; CPU twin
// totsu_f64lapack/src/f64lapack.rs:78-108).  eig.cu computes proj_{S+}(X) = (X + X sign(X)) / 2 with a GEMM-only
// polynomial iteration; every product in it is  C = alpha * A * B + beta * D + gamma * I  with A, B, D symmetric
// k x k (f32, column-major) and C symmetric.  This file is that GEMM on the 5th-generation tensor cores:
//
//   * tcgen05.mma.cta_group::1.kind::tf32, M = 128, N = 128, K = 8 per instruction, accumulator in TMEM (128 columns);
//   * fp32 accuracy from three TF32 products per K step ("3xTF32"):  A = Ah + Al, B = Bh + Bl with Ah = rna_tf32(A),
//     Al = A - Ah (exact in fp32)  =>  A*B ~= Ah*Bh + Ah*Bl + Al*Bh, dropped term Al*Bl <= 2^-22 |A||B|;
//   * both operands are K-major for free: B[l, j] is column j of the column-major array, and A[i, l] = A[l, i]
//     (A symmetric) is column i - so a tile row is a contiguous run of K in global memory;
//   * 8 producer warps read 128-byte row segments (L2-resident: the matrices are <= a few MB), split hi/lo in
//     registers and store both into the canonical no-swizzle K-major core-matrix layout (8 rows x 16 B contiguous,
//     LBO = 144 B along K - a 16-byte skew that makes coalesced global reads and conflict-free shared stores
//     compatible - SBO = 1152 B along M/N), mbarrier ring + one chunk prefetched in registers; one elected lane of
//     warp 8 issues the MMAs and releases stages with tcgen05.commit;
//   * split-K across a thread-block cluster (1,1,S), S <= 8: each CTA owns K/S of the reduction; from TMEM it PUSHES
//     (st.shared::cluster) the rows [z'*128/S, (z'+1)*128/S) of its partial tile into the inbox of CTA z', and after one
//     cluster barrier every CTA sums the S partials of its own rows from local shared memory in rank order (fixed
//     order: bit-reproducible, no atomics) and runs the epilogue for them - stores are fire-and-forget, nothing waits
//     on a remote load;
//   * launched with programmatic stream serialization: barrier init + TMEM allocation of GEMM n+1 overlap the tail of
//     GEMM n (griddepcontrol.wait before the first dependent access);
//   * epilogue keeps entries on or above the diagonal and mirrors them (through a padded smem transpose so both the
//     direct and the mirrored stores are coalesced): iterates stay EXACTLY symmetric, which the sign iteration needs
//     (eig.cu) and which makes the A-operand trick above legal for the next product.
//
// k = 512 (BASELINE config C4) gives only 10 upper 128 x 128 tiles, so the machine is far from full: split-K 8 brings
// it to 80 CTAs of 2 K-chunks each; the GEMM is then bounded by fixed latencies (launch, L2 round trip, cluster
// barrier, epilogue), not by the tensor pipe - DESIGN.md section 3.4 has the measured numbers.
#include "common.cuh"

namespace tb {
namespace tc {

constexpr int TM = 128, TN = 128, KC = 32;
constexpr int PRODUCER_WARPS = 8;
constexpr int THREADS = (PRODUCER_WARPS + 1) * 32;
// Core matrices (8 rows x 16 B = 128 B) are laid out 144 B apart along K: the 16-byte skew puts the eight 16-byte
// K-chunks of one row into eight different bank groups, so a quarter-warp can read one row's contiguous 128 B from
// global memory (fully coalesced: one L1 tag per quarter-warp) AND store them to shared memory conflict-free.
constexpr int LBO = 144;                          // next core matrix along K
constexpr int SBO = (KC / 4) * LBO;               // next 8-row group along M/N: 1152 B
constexpr int TILE_BYTES = (TM / 8) * SBO;        // one operand half (hi or lo) of one stage: 18 KB
constexpr int STAGE_BYTES = 4 * TILE_BYTES;       // A_hi | A_lo | B_hi | B_lo
constexpr int PAD = TM + 1;                       // padded leading dimension of the staging tiles
constexpr int TMEM_COLS = 128;
static_assert(TM == TN, "the staging tiles assume square tiles");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"((uint32_t)TMEM_COLS) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, both K-major, TF32 inputs, fp32 accumulation
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive columns: thread `lane` of the warp gets row (lane_base + lane), columns [col, col + 32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared-memory matrix descriptor, SWIZZLE_NONE, K-major (cute/arch/mma_sm100_desc.hpp SmemDescriptor layout):
// start address >> 4 at [0,14), leading byte offset >> 4 at [16,30), stride byte offset >> 4 at [32,46), version 1 at [46,48)
__host__ __device__ __forceinline__ uint64_t smem_desc_fields(uint32_t lbo, uint32_t sbo) {
    return ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint64_t smem_desc(uint64_t fields, uint32_t saddr) { return fields | (uint64_t)((saddr & 0x3FFFFu) >> 4); }
// instruction descriptor (InstrDescriptor): c_format F32 = 1 at [4,6), a/b_format TF32 = 2 at [7,10)/[10,13),
// a/b major K = 0 at 15/16, N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

// hi = x rounded to TF32 (nearest, ties away: add half an ulp of the 10-bit mantissa to the magnitude, clear the low 13
// bits - what cvt.rna.tf32.f32 computes, without the NaN/Inf branch ptxas emits for it), lo = x - hi (exact)
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ void split_tf32(const float4& v, float4& hi, float4& lo) {
    hi.x = tf32_hi(v.x); lo.x = v.x - hi.x;
    hi.y = tf32_hi(v.y); lo.y = v.y - hi.y;
    hi.z = tf32_hi(v.z); lo.z = v.z - hi.z;
    hi.w = tf32_hi(v.w); lo.w = v.w - hi.w;
}

__device__ __forceinline__ void st_dsmem(uint32_t remote_addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote_addr), "f"(v) : "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_addr, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
    return remote;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// programmatic dependent launch: everything before pdl_wait() overlaps the tail of the previous kernel in the stream
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// phase stamps of CTA 0 (diagnostics: tb_symm_gemm_trace); slot i of `trace` is written by one thread
#define TC_STAMP(i) do { if (trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (tid == 0)) trace[i] = gtimer(); } while (0)

template <int SPLITK> struct Cfg {
    static constexpr int STAGES = SPLITK > 1 ? 2 : 3;             // + one chunk in flight in registers
    static constexpr int ROWS = TM / SPLITK;                      // tile rows this CTA finishes after the split-K exchange
    static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
    static constexpr int INBOX_BYTES = SPLITK > 1 ? SPLITK * TN * ROWS * 4 : 0;   // [source rank][column][ROWS]: 64 KB
    static constexpr int SMEM = PIPE_BYTES + INBOX_BYTES + 128;
    static_assert(ROWS * PAD * 4 <= PIPE_BYTES, "the mirror staging tile reuses the pipeline buffers");
};

// Up to two independent products per launch (blockIdx.y): the two ConePSD projections of one solver iteration (dual cone
// on the y block, primal cone on the s block: solver.rs:548-549) run as one batch and fill twice as many SMs.
struct TcOperands {
    const float* A[2];
    const float* B[2];
    const float* D[2];
    float* C[2];
};

// C = alpha * (A * B) + beta * D + gamma * I; A, B, D symmetric k x k column-major, k % 4 == 0, 16-byte aligned.
// grid = (upper tile pairs, batch, SPLITK), cluster (1, 1, SPLITK).
template <int SPLITK>
__global__ void __launch_bounds__(THREADS, 1) symm_gemm_tc_kernel(const TcOperands ops, int k, float alpha, float beta, float gamma, uint64_t dfields,
                                                                  unsigned long long* __restrict__ trace) {
    const float* __restrict__ A = ops.A[blockIdx.y];
    const float* __restrict__ B = ops.B[blockIdx.y];
    const float* __restrict__ D = ops.D[blockIdx.y];
    float* __restrict__ C = ops.C[blockIdx.y];
    using G = Cfg<SPLITK>;
    constexpr int STAGES = G::STAGES, ROWS = G::ROWS;
    extern __shared__ __align__(1024) uint8_t smem[];
    float* inbox = reinterpret_cast<float*>(smem + G::PIPE_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + G::PIPE_BYTES + G::INBOX_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* accf = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accf + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    TC_STAMP(0);                           // kernel entered
    // decode the upper tile pair (bi <= bj) from blockIdx.x: row bi holds nt - bi tiles
    const int nt = (k + TM - 1) / TM;
    int bi = 0, rem = blockIdx.x;
    while (rem >= nt - bi) { rem -= nt - bi; ++bi; }
    const int bj = bi + rem;
    const int i0 = bi * TM, j0 = bj * TN;
    const int z = SPLITK > 1 ? (int)cluster_rank() : 0;
    const int nk = (k + KC - 1) / KC;
    const int c0 = (int)((long long)z * nk / SPLITK), c1 = (int)((long long)(z + 1) * nk / SPLITK);
    const int nmy = c1 - c0;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], PRODUCER_WARPS * 32); mbar_init(&empty[s], 1); }
        mbar_init(accf, 1);
        mbar_fence_init();
    }
    if (warp == PRODUCER_WARPS) tmem_alloc(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    TC_STAMP(1);                           // prologue done (barriers, TMEM)
    if (SPLITK > 1) cluster_arrive();      // "I have started": matched by cluster_wait() before the first remote store
    pdl_wait();                            // the producing kernel's stores are visible from here on
    pdl_launch_dependents();               // the next kernel may run its prologue while this one works
    TC_STAMP(2);                           // dependency resolved

    if (warp < PRODUCER_WARPS) {
        // ---- producers: global (L2) -> registers -> hi/lo split -> canonical K-major core matrices in smem
        // lane -> (row within a group of 4, 16-byte K chunk): a quarter-warp covers one row's 128 contiguous bytes
        const int r4 = lane >> 3, ck = lane & 7;
        auto load = [&](int chunk, float4* v) {
            const int kk = chunk * KC + ck * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int row = q * 32 + warp * 4 + r4;
                const int gi = i0 + row, gj = j0 + row;
                v[q] = (gi < k && kk < k) ? *reinterpret_cast<const float4*>(A + (size_t)gi * k + kk) : make_float4(0.f, 0.f, 0.f, 0.f);
                v[4 + q] = (gj < k && kk < k) ? *reinterpret_cast<const float4*>(B + (size_t)gj * k + kk) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        float4 cur[8], nxt[8];
        if (nmy > 0) load(c0, cur);
        for (int it = 0; it < nmy; ++it) {
            if (it + 1 < nmy) load(c0 + it + 1, nxt);            // next chunk's loads fly while this one is split and stored
            const int s = it % STAGES;
            const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
            mbar_wait(&empty[s], ph ^ 1u);
            uint8_t* st = smem + s * STAGE_BYTES;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int row = q * 32 + warp * 4 + r4;
                const int off = (row >> 3) * SBO + ck * LBO + (row & 7) * 16;
                float4 hi, lo;
                split_tf32(cur[q], hi, lo);
                *reinterpret_cast<float4*>(st + off) = hi;
                *reinterpret_cast<float4*>(st + TILE_BYTES + off) = lo;
                split_tf32(cur[4 + q], hi, lo);
                *reinterpret_cast<float4*>(st + 2 * TILE_BYTES + off) = hi;
                *reinterpret_cast<float4*>(st + 3 * TILE_BYTES + off) = lo;
            }
            fence_async_shared();          // generic-proxy stores -> visible to the tensor core's async-proxy reads
            mbar_arrive(&full[s]);
            if (it == 0) TC_STAMP(3);      // first chunk loaded, split and stored
#pragma unroll
            for (int q = 0; q < 8; ++q) cur[q] = nxt[q];
        }
    } else {
        // ---- MMA issuer: one lane, 3 TF32 products per K step of 8 (small terms first)
        for (int it = 0; it < nmy; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
                for (int ks = 0; ks < KC / 8; ++ks) {
                    const uint32_t o = (uint32_t)ks * 2u * LBO;
                    const uint64_t ah = smem_desc(dfields, sa + o), al = smem_desc(dfields, sa + TILE_BYTES + o);
                    const uint64_t bh = smem_desc(dfields, sa + 2 * TILE_BYTES + o), bl = smem_desc(dfields, sa + 3 * TILE_BYTES + o);
                    mma_tf32(tmem_base, al, bh, IDESC, (it | ks) != 0 ? 1u : 0u);
                    mma_tf32(tmem_base, ah, bl, IDESC, 1u);
                    mma_tf32(tmem_base, ah, bh, IDESC, 1u);
                }
                mma_commit(&empty[s]);                       // stage reusable once these MMAs have read it
                if (it == nmy - 1) mma_commit(accf);         // accumulator complete
            }
            __syncwarp();
        }
    }

    // ---- epilogue: TMEM -> registers (thread = tile row, registers = 32 consecutive columns)
    float* Rt = reinterpret_cast<float*>(smem);     // [row][PAD] finished values for the mirror pass (the pipeline is drained by then)
    if (SPLITK > 1) cluster_wait();                 // every CTA of the cluster runs: its inbox may be written
    if (warp < PRODUCER_WARPS) {
        const int quad = warp & 3, chalf = warp >> 2;
        const int row = quad * 32 + lane;
        TC_STAMP(4);                       // producers done
        if (nmy > 0) { mbar_wait(accf, 0u); tc_fence_after(); }
        TC_STAMP(5);                       // accumulator complete
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
            const int cbase = chalf * 64 + cc * 32;
            float v[32];
            if (nmy > 0) tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)cbase, v);
            else {
#pragma unroll
                for (int t = 0; t < 32; ++t) v[t] = 0.f;
            }
            if (SPLITK > 1) {
                // push this row's partial sums into the inbox of the CTA that finishes the row: slot [z][col][row % ROWS]
                const uint32_t la = smem_u32(&inbox[((size_t)z * TN + cbase) * ROWS + (row % ROWS)]);
                const uint32_t ra = map_to_rank(la, (uint32_t)(row / ROWS));
#pragma unroll
                for (int t = 0; t < 32; ++t) st_dsmem(ra + (uint32_t)(t * ROWS * 4), v[t]);
            } else {
                const int gi = i0 + row;
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    const int gj = j0 + cbase + t;
                    float val = alpha * v[t];
                    if (gi < k && gj < k && gi <= gj) {
                        if (beta != 0.f) val += beta * D[(size_t)gj * k + gi];
                        if (gi == gj) val += gamma;
                        C[(size_t)gj * k + gi] = val;
                    }
                    Rt[row * PAD + cbase + t] = val;
                }
            }
        }
    }
    tc_fence_before();
    if (SPLITK > 1) {
        TC_STAMP(6);                        // partial sums pushed
        cluster_arrive();
        cluster_wait();                     // all partial sums have landed; nothing remote is touched after this point
        TC_STAMP(7);
        if (warp < PRODUCER_WARPS) {
            for (int idx = tid; idx < TN * ROWS; idx += PRODUCER_WARPS * 32) {
                const int i = idx % ROWS, j = idx / ROWS;
                float acc = 0.f;
#pragma unroll
                for (int r = 0; r < SPLITK; ++r) acc += inbox[r * TN * ROWS + idx];       // rank order: bit-reproducible
                const int gi = i0 + z * ROWS + i, gj = j0 + j;
                float val = alpha * acc;
                if (gi < k && gj < k && gi <= gj) {
                    if (beta != 0.f) val += beta * D[(size_t)gj * k + gi];
                    if (gi == gj) val += gamma;
                    C[(size_t)gj * k + gi] = val;
                }
                Rt[i * PAD + j] = val;
            }
        }
        TC_STAMP(9);                        // own rows reduced and stored (warp 0)
    }
    __syncthreads();
    TC_STAMP(10);
    if (warp < PRODUCER_WARPS) {
        // mirror: rows of the finished block become columns of C below the diagonal (coalesced along j)
        for (int i = warp; i < ROWS; i += PRODUCER_WARPS) {
            const int gi = i0 + z * ROWS + i;
            for (int jj = lane; jj < TN; jj += 32) {
                const int gj = j0 + jj;
                if (gi < k && gj < k && gi < gj) C[(size_t)gi * k + gj] = Rt[i * PAD + jj];
            }
        }
    }
    __syncthreads();
    TC_STAMP(8);                            // results stored
    tc_fence_after();
    if (warp == PRODUCER_WARPS) tmem_dealloc(tmem_base);
}

static bool env_flag(const char* name, bool dflt) {
    const char* e = getenv(name);
    return e ? e[0] != '0' : dflt;
}

template <int SPLITK> static void launch(const TcOperands& ops, int batch, int k, float alpha, float beta, float gamma, unsigned long long* trace) {
    static bool configured = false;
    auto kern = symm_gemm_tc_kernel<SPLITK>;
    if (!configured) {
        TB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<SPLITK>::SMEM));
        configured = true;
    }
    const int nt = (k + TM - 1) / TM;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(nt * (nt + 1) / 2), (unsigned)batch, SPLITK);
    cfg.blockDim = dim3(THREADS, 1, 1);
    cfg.dynamicSmemBytes = Cfg<SPLITK>::SMEM;
    cfg.stream = ctx().stream;
    cudaLaunchAttribute attr[2];
    unsigned na = 0;
    if (SPLITK > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 1;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = SPLITK;
        ++na;
    }
    // TB_TC_PDL=0 turns programmatic dependent launch off (A/B measurement)
    static const bool pdl = env_flag("TB_TC_PDL", true);
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    // TB_TC_DESC_SWAP=1 (debug) exchanges the leading/stride byte offsets of the operand descriptors
    static const bool swap = env_flag("TB_TC_DESC_SWAP", false);
    const uint64_t dfields = swap ? smem_desc_fields(SBO, LBO) : smem_desc_fields(LBO, SBO);
    TB_CUDA(cudaLaunchKernelEx(&cfg, kern, ops, k, alpha, beta, gamma, dfields, trace));
    count_launch();
}

}  // namespace tc

bool symm_gemm_tc_usable(const float* A, const float* B, const float* D, const float* C, size_t k) {
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    return k >= 4 && k % 4 == 0 && k <= 32768 && al(A) && al(B) && al(C) && (D == nullptr || al(D));
}

static int choose_splitk(size_t k, int batch) {
    // enough CTAs to occupy the machine (one CTA per SM: never more CTAs than SMs), at least two K chunks per CTA
    const size_t nt = (k + tc::TM - 1) / tc::TM, pairs = nt * (nt + 1) / 2 * (size_t)batch, nk = (k + tc::KC - 1) / tc::KC;
    static const int cap = [] { const char* e = getenv("TB_TC_MAX_SPLITK"); return e ? atoi(e) : 8; }();
    int splitk = 1;
    while (splitk < cap && pairs * (size_t)splitk * 2 <= (size_t)ctx().sm_count && nk / (size_t)(splitk * 2) >= 2) splitk *= 2;
    return splitk;
}

static void dispatch(const tc::TcOperands& ops, int batch, size_t k, float alpha, float beta, float gamma, int splitk, unsigned long long* trace) {
    if (splitk == 0) splitk = choose_splitk(k, batch);
    switch (splitk) {
        case 1: tc::launch<1>(ops, batch, (int)k, alpha, beta, gamma, trace); break;
        case 2: tc::launch<2>(ops, batch, (int)k, alpha, beta, gamma, trace); break;
        case 4: tc::launch<4>(ops, batch, (int)k, alpha, beta, gamma, trace); break;
        case 8: tc::launch<8>(ops, batch, (int)k, alpha, beta, gamma, trace); break;
        default: fail(TB_ERR_ARG, "symm_gemm_tc: splitk must be 0, 1, 2, 4 or 8");
    }
}

// splitk: 0 = choose, else 1 / 2 / 4 / 8
void symm_gemm_tc(const float* A, const float* B, const float* D, float* C, size_t k, float alpha, float beta, float gamma, int splitk,
                  unsigned long long* trace) {
    TB_REQUIRE(symm_gemm_tc_usable(A, B, D, C, k), "tensor-core symmetric GEMM needs k % 4 == 0 and 16-byte aligned matrices");
    TB_REQUIRE(C != A && C != B && C != D, "symm_gemm: the output must not alias an input");
    tc::TcOperands ops{};
    ops.A[0] = A; ops.B[0] = B; ops.D[0] = D; ops.C[0] = C;
    dispatch(ops, 1, k, alpha, beta, gamma, splitk, trace);
}

// two independent products with the same scalars in one launch
void symm_gemm_tc_pair(const float* const A[2], const float* const B[2], const float* const D[2], float* const C[2], size_t k,
                       float alpha, float beta, float gamma, int splitk) {
    tc::TcOperands ops{};
    for (int i = 0; i < 2; ++i) {
        TB_REQUIRE(symm_gemm_tc_usable(A[i], B[i], D[i], C[i], k), "tensor-core symmetric GEMM needs k % 4 == 0 and 16-byte aligned matrices");
        TB_REQUIRE(C[i] != A[i] && C[i] != B[i] && C[i] != D[i], "symm_gemm: the output must not alias an input");
        ops.A[i] = A[i]; ops.B[i] = B[i]; ops.D[i] = D[i]; ops.C[i] = C[i];
    }
    dispatch(ops, 2, k, alpha, beta, gamma, splitk, nullptr);
}

}  // namespace tb
