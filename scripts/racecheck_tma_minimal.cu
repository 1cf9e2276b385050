// Minimal textbook TMA pipeline (one producer warp, four consumer warps, 2-stage full / empty mbarrier ring): the same
// synchronisation pattern as csrc/gemv.cu stream_kernel and csrc/spmv.cu, in 80 lines, with nothing else around it.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o /tmp/rc_min scripts/racecheck_tma_minimal.cu
//   compute-sanitizer --tool racecheck /tmp/rc_min
//
// Purpose: compute-sanitizer's racecheck reports a shared-memory hazard between the bulk-copy write (async proxy,
// cp.async.bulk ... mbarrier::complete_tx) and the consumers' reads on stream_kernel (profiles/r02_sanitizer_summary.md).
// The ordering there is full_bar (complete_tx -> consumer try_wait) for RAW and empty_bar (consumer arrive -> producer
// try_wait) for WAR - the pattern of every TMA pipeline.  If racecheck flags THIS kernel too, the report is a limitation of
// the tool's model of complete_tx barriers, not a property of stream_kernel.  The program checks its own result.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int STAGES = 2, TILE = 4096, CONSUMER_WARPS = 4;      // floats per tile

__global__ void __launch_bounds__(32 * (CONSUMER_WARPS + 1)) pipeline_sum(const float* __restrict__ x, int n_tiles, float* out) {
    __shared__ __align__(128) float stage[STAGES][TILE];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CONSUMER_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int s = 0;
    uint32_t phase = 0;
    if (warp == CONSUMER_WARPS) {
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            mbar_wait(&empty_bar[s], phase ^ 1);
            if (lane == 0) {
                mbar_expect_tx(&full_bar[s], TILE * sizeof(float));
                bulk_g2s(stage[s], x + (size_t)t * TILE, TILE * sizeof(float), &full_bar[s]);
            }
            if (++s == STAGES) { s = 0; phase ^= 1; }
        }
    } else {
        float acc = 0.f;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            mbar_wait(&full_bar[s], phase);
            for (int i = threadIdx.x; i < TILE; i += 32 * CONSUMER_WARPS) acc += stage[s][i];
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
            if (++s == STAGES) { s = 0; phase ^= 1; }
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) atomicAdd(out, acc);
    }
}

int main() {
    const int n_tiles = 64;
    float *x, *out;
    cudaMalloc(&x, (size_t)n_tiles * TILE * sizeof(float));
    cudaMalloc(&out, sizeof(float));
    cudaMemset(out, 0, sizeof(float));
    float* h = new float[(size_t)n_tiles * TILE];
    for (size_t i = 0; i < (size_t)n_tiles * TILE; ++i) h[i] = 1.0f;
    cudaMemcpy(x, h, (size_t)n_tiles * TILE * sizeof(float), cudaMemcpyHostToDevice);
    pipeline_sum<<<4, 32 * (CONSUMER_WARPS + 1)>>>(x, n_tiles, out);
    float r = 0.f;
    cudaError_t e = cudaMemcpy(&r, out, sizeof(float), cudaMemcpyDeviceToHost);
    std::printf("minimal TMA pipeline: sum = %.1f (want %.1f), cuda status %d\n", r, (float)n_tiles * TILE, (int)e);
    return (e == cudaSuccess && r == (float)n_tiles * TILE) ? 0 : 1;
}
