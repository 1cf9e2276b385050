/* oracle/native.c - TEST INFRASTRUCTURE ONLY (CPU side of the checker; never linked into the product).
 *
 * Host twin of the counter-based instance generator (totsu_b200/csrc/synth.cu, numpy twin totsu_b200/synth.py):
 *   A[r, c] = scale * base(seed, row_offset + r, c),  base = (2*u24 + 1)/2^24 - 1,  u24 = mix64(mix64(seed*G + row) ^ c*C) >> 40
 * so the f64 CPU oracle / CPU baseline sees bit-identical inputs at the benchmark's full size (config C3: 65536 x 16384,
 * 8.6 GB in f64) in seconds instead of the minutes the numpy twin takes.  This is data generation for the synthetic
 * workloads of SURVEY.md 8d, not part of the reference's algorithm (the reference has no generator: its
 * experimental/benchmark_lp/src/main.rs:13-74 draws from rand::thread_rng).
 *
 * Build: oracle/Makefile (gcc -O3 -pthread -shared -fPIC) -> oracle/liboracle_native.so
 */
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <unistd.h>

static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

struct fill_job {
    double* out;
    size_t n_row, c0, c1, ld, row_offset;
    uint64_t seed;
    double scale;
    int as_f32;
};

static void* fill_cols(void* arg) {
    const struct fill_job* j = (const struct fill_job*)arg;
    const float scale_f = (float)j->scale;
    for (size_t c = j->c0; c < j->c1; ++c) {
        double* col = j->out + c * j->ld;
        const uint64_t ck = (uint64_t)c * 0xD1B54A32D192ED03ULL;
        for (size_t r = 0; r < j->n_row; ++r) {
            const uint64_t key = mix64(j->seed * 0x9E3779B97F4A7C15ULL + (uint64_t)(j->row_offset + r)) ^ ck;
            const uint32_t u24 = (uint32_t)(mix64(key) >> 40);
            const double base = (double)(2 * (int64_t)u24 + 1) * (1.0 / 16777216.0) - 1.0;   /* exact in f32 and f64 */
            col[r] = (j->as_f32 ? (double)(scale_f * (float)base) : j->scale * base);
        }
    }
    return 0;
}

/* out: column-major n_row x n_col with leading dimension ld (>= n_row), f64.
 * as_f32 != 0: the value is computed like the f32 device path (float scale * float base) and widened, so the f64 oracle
 * and the f32 device see identical inputs.
 * Columns are dealt to `threads` pthreads (<= 0: all online cores). */
void oracle_fill_uniform_f64(double* out, size_t n_row, size_t n_col, size_t ld, size_t row_offset, uint64_t seed,
                             double scale, int as_f32, int threads) {
    enum { MAXT = 256 };
    if (threads <= 0) threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (threads > MAXT) threads = MAXT;
    if ((size_t)threads > n_col) threads = n_col ? (int)n_col : 1;
    if (n_row * n_col < (size_t)1 << 16) threads = 1;
    pthread_t tid[MAXT];
    struct fill_job job[MAXT];
    for (int t = 0; t < threads; ++t) {
        job[t] = (struct fill_job){out, n_row, n_col * (size_t)t / (size_t)threads, n_col * (size_t)(t + 1) / (size_t)threads, ld, row_offset, seed, scale, as_f32};
        if (t + 1 < threads) pthread_create(&tid[t], 0, fill_cols, &job[t]);
    }
    fill_cols(&job[threads - 1]);
    for (int t = 0; t + 1 < threads; ++t) pthread_join(tid[t], 0);
}

/* one row of the same matrix (the c_i^T rows of the SOCP stacking), length n_col */
void oracle_fill_uniform_row_f64(double* out, size_t n_col, size_t row, uint64_t seed, double scale, int as_f32) {
    const float scale_f = (float)scale;
    const uint64_t rk = mix64(seed * 0x9E3779B97F4A7C15ULL + (uint64_t)row);
    for (size_t c = 0; c < n_col; ++c) {
        const uint32_t u24 = (uint32_t)(mix64(rk ^ ((uint64_t)c * 0xD1B54A32D192ED03ULL)) >> 40);
        const double base = (double)(2 * (int64_t)u24 + 1) * (1.0 / 16777216.0) - 1.0;
        out[c] = (as_f32 ? (double)(scale_f * (float)base) : scale * base);
    }
}

int oracle_native_abi(void) { return 1; }
