// Synthetic dense instances: a counter-based generator keyed on (seed, global row, column), so
// any row shard on any rank - and the numpy twin in totsu_b200/synth.py used by the CPU oracle - regenerates
// bit-identical values without ever materialising the matrix on the host (config C5's A is 68.7 GB).
#include "common.cuh"

namespace tb {

__host__ __device__ inline uint64_t mix64(uint64_t z) {          // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// u24 in [0, 2^24): value = scale * ((2*u24 + 1) / 2^24 - 1), exactly representable before the final multiply
template <typename T>
__global__ void fill_uniform_kernel(T* a, size_t n_row, size_t n_col, size_t row_offset, uint64_t seed, T scale) {
    const size_t total = n_row * n_col;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t r = idx % n_row, c = idx / n_row;
        const uint64_t key = mix64(seed * 0x9E3779B97F4A7C15ULL + (uint64_t)(row_offset + r)) ^ ((uint64_t)c * 0xD1B54A32D192ED03ULL);
        const uint32_t u24 = (uint32_t)(mix64(key) >> 40);
        const T base = (T)((double)(2 * (int64_t)u24 + 1) * (1.0 / 16777216.0) - 1.0);   // exact in f32 and f64
        a[idx] = scale * base;
    }
}

template <typename T> static void api_fill_uniform(tb_view mat, size_t n_row, size_t n_col, size_t row_offset, uint64_t seed, T scale) {
    require_init();
    TB_REQUIRE(mat.len == n_row * n_col, "fill_uniform: mat.len != n_row*n_col");
    T* a = wptr<T>(mat, true);
    if (mat.len == 0) return;
    int g = (int)std::min<size_t>((mat.len + 255) / 256, (size_t)ctx().sm_count * 32);
    fill_uniform_kernel<T><<<g, 256, 0, ctx().stream>>>(a, n_row, n_col, row_offset, seed, scale);
    TB_LAUNCH_CHECK();
}

}  // namespace tb

using namespace tb;
extern "C" {
int tb_fill_uniform_f32(tb_view m, size_t nr, size_t nc, size_t ro, uint64_t seed, float scale) { return api([&] { api_fill_uniform<float>(m, nr, nc, ro, seed, scale); }); }
int tb_fill_uniform_f64(tb_view m, size_t nr, size_t nc, size_t ro, uint64_t seed, double scale) { return api([&] { api_fill_uniform<double>(m, nr, nc, ro, seed, scale); }); }
}
