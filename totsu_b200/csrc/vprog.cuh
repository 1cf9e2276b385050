// Recorder interface of the vector-program executor (vprog.cu).  Launch sites of the small vector kernels call these
// instead of launching when vp_enabled(n); pointers are device pointers, dtype is TB_F32 / TB_F64.
#pragma once
#include <cstddef>

namespace tb {

bool vp_enabled_wide(size_t n);   // element-wise ops: any length up to 4M elements (long ones make the program a barrier-free wide one)
bool vp_enabled(size_t n = 1);      // recording is on and a vector of n elements is short enough for the cluster executor
void vp_set_max_n(size_t n);     // longest vector a cluster program takes (flushes the pending program)
bool vp_enabled_red(size_t n);      // reductions stay on the cluster executor up to 512 K elements
void vp_pf_reduce(int dtype, int kind, const void* a, const void* b, size_t n, int k);   // scalar prefetch riding in the pending program
void vp_pf_seq(unsigned long long seq);
void vp_init();
void vp_shutdown();
void vp_fill(int dtype, void* y, double v, size_t n);
void vp_scale(int dtype, double a, void* y, size_t n);
void vp_copy(int dtype, const void* x, void* y, size_t n);
void vp_axpby(int dtype, double a, const void* x, double b, void* y, size_t n);                  // y = a x + b y
void vp_adds(int dtype, double s, void* y, size_t n);
void vp_diag(int dtype, double a, const void* d, const void* x, double b, void* y, size_t n);   // y = a (d .* x) + b y
void vp_finalize(int dtype, const void* part, int nparts, size_t ld, size_t len, double a, double b, void* y);
void vp_axs(int dtype, double a, const void* x, const void* s, double b, void* y, size_t n);    // y = a x s[0] + b y
void vp_axs_imm(int dtype, double a, const void* x, double s, double b, void* y, size_t n);   // the same with the scalar by value
void vp_set1(int dtype, void* y, double v);
void vp_dot(int dtype, double a, const void* x, const void* d, size_t n, double b, void* y);     // y[0] = a <x, d> + b y[0]
double vp_reduce_to_host(int dtype, int mode, const void* x, size_t count, size_t inc);          // mode 0: sum x^2, 1: sum |x|
double vp_fetch_to_host(int dtype, const void* x);

}  // namespace tb
