// Host-side mirror of the reference's plugin surface, written in C++ because the reference's own language
// (Rust) has no toolchain in this image.  It is what the Rust crate `totsu_b200` does, 1:1:
//
//   B200<F>      : LinAlg  (totsu_core/src/solver/linalg.rs:10-68)  +  LinAlgEx (totsu_core/src/linalg_ex.rs:7-66)
//   Slice<F>     : SliceLike (totsu_core/src/solver/slicelike.rs:9-70), the role of F32CUDASlice
//
// Every method forwards to one C-ABI entry point of libtotsu_b200.so and asserts on its status - the traits
// have no error channel, exactly like totsu_f32cuda asserts on cuBLAS statuses (f32cuda.rs:38).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "../../include/totsu_b200.h"

namespace totsu_b200 {

struct BackendError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline void tb_check(int status, const char* what) {
    if (status != TB_OK) throw BackendError(std::string(what) + ": " + tb_last_error());
}
#define TBH_CALL(expr) ::totsu_b200::tb_check((expr), #expr)

// per-precision dispatch to the suffixed C functions
template <typename F> struct Abi;
#define TBH_ABI(F_, S_)                                                                                                   \
    template <> struct Abi<F_> {                                                                                          \
        static constexpr int dtype = (sizeof(F_) == 4 ? TB_F32 : TB_F64);                                                 \
        static int get1(tb_view v, size_t i, F_* o) { return tb_get1_##S_(v, i, o); }                                      \
        static int set1(tb_view v, size_t i, F_ x) { return tb_set1_##S_(v, i, x); }                                       \
        static int norm(tb_view x, F_* o) { return tb_norm_##S_(x, o); }                                                   \
        static int copy(tb_view x, tb_view y) { return tb_copy_##S_(x, y); }                                               \
        static int scale(F_ a, tb_view x) { return tb_scale_##S_(a, x); }                                                  \
        static int add(F_ a, tb_view x, tb_view y) { return tb_add_##S_(a, x, y); }                                        \
        static int adds(F_ s, tb_view y) { return tb_adds_##S_(s, y); }                                                    \
        static int abssum(tb_view x, size_t inc, F_* o) { return tb_abssum_##S_(x, inc, o); }                              \
        static int transform_di(F_ a, tb_view m, tb_view x, F_ b, tb_view y) { return tb_transform_di_##S_(a, m, x, b, y); } \
        static int transform_ge(int t, size_t nr, size_t nc, F_ a, tb_view m, tb_view x, F_ b, tb_view y) {                \
            return tb_transform_ge_##S_(t, nr, nc, a, m, x, b, y);                                                         \
        }                                                                                                                  \
        static int transform_sp(size_t n, F_ a, tb_view m, tb_view x, F_ b, tb_view y) { return tb_transform_sp_##S_(n, a, m, x, b, y); } \
        static int map_eig_begin(tb_view m, int hs, F_ sd, F_ ez, tb_view w, F_* e) { return tb_map_eig_begin_##S_(m, hs, sd, ez, w, e); } \
        static int map_eig_finish(tb_view m, int hs, F_ sd, tb_view w, const F_* ne, const uint8_t* k) {                   \
            return tb_map_eig_finish_##S_(m, hs, sd, w, ne, k);                                                            \
        }                                                                                                                  \
        static int proj_psd(tb_view x, F_ ez, tb_view w) { return tb_proj_psd_##S_(x, ez, w); }                            \
        static int sqrt_psd(tb_view m, F_ ez, tb_view w) { return tb_sqrt_psd_##S_(m, ez, w); }                            \
        static int denseop_apply(tb_handle h, int t, F_ a, tb_view x, F_ b, tb_view y) { return tb_denseop_apply_##S_(h, t, a, x, b, y); } \
        static int denseop_apply_pair(tb_handle h, F_ an, tb_view xn, F_ bn, tb_view yn, F_ at, tb_view xt, F_ bt, tb_view yt) { \
            return tb_denseop_apply_pair_##S_(h, an, xn, bn, yn, at, xt, bt, yt);                                          \
        }                                                                                                                  \
        static int denseop_absadd_cols(tb_handle h, tb_view t) { return tb_denseop_absadd_cols_##S_(h, t); }               \
        static int denseop_absadd_rows(tb_handle h, tb_view s) { return tb_denseop_absadd_rows_##S_(h, s); }               \
        static int cone_proj(tb_handle h, int d, tb_view x, F_ ez, tb_view w) { return tb_cone_proj_##S_(h, d, x, ez, w); } \
        static int cone_group_min(tb_handle h, tb_view t) { return tb_cone_group_min_##S_(h, t); }                         \
        static int recip_clamp(F_ e, tb_view x) { return tb_recip_clamp_##S_(e, x); }                                      \
    };
TBH_ABI(float, f32)
TBH_ABI(double, f64)
#undef TBH_ABI

// ---------------------------------------------------------------------------------------------------------
// "Shim-protocol" mode.  The Rust binding's slice type IS the host sub-slice (rust/totsu_b200/src/b200_slice.rs: a
// transparent wrapper of `[F]`), so it carries no handle: every operand of every call is resolved with
// tb_view_of_host, every split child takes a reference on its root (tb_buf_retain) and gives it back when it drops
// (tb_view_of_host + tb_buf_release).  This mirror normally carries (handle, offset, length) views instead, which skips
// those calls.  With shim_protocol() on it issues exactly the binding's call sequence - and checks that each lookup
// resolves to the view it carries - so that what is tested and measured here transfers to `Solver<B200>`.
// ---------------------------------------------------------------------------------------------------------
inline bool& shim_protocol() {
    static bool on = false;
    return on;
}
// map_eig: recognise the reference's own two closures and take the GEMM-only paths (default on; off = always the general route)
inline bool& map_eig_fast_paths() {
    static bool on = true;
    return on;
}

// ---------------------------------------------------------------------------------------------------------
// Slice<F>: SliceLike.  A root slice (new_ref / new_mut) owns a device mirror of caller memory and restores
// host coherence when dropped (slicelike.rs:18-19); split slices are plain (handle, offset, length) views.
// ---------------------------------------------------------------------------------------------------------
template <typename F> class Slice {
public:
    Slice() = default;
    Slice(const Slice&) = delete;
    Slice& operator=(const Slice&) = delete;
    Slice(Slice&& o) noexcept { *this = std::move(o); }
    Slice& operator=(Slice&& o) noexcept {
        if (this != &o) {
            drop();
            v_ = o.v_; root_ = o.root_; host_ = o.host_; mut_ = o.mut_;
            child_ = o.child_; rest_host_ = o.rest_host_; rest_len_ = o.rest_len_;
            o.root_ = false; o.v_ = tb_view{0, 0, 0}; o.child_ = false; o.rest_len_ = 0;
        }
        return *this;
    }
    ~Slice() { drop(); }

    static Slice new_ref(const F* s, size_t len) {              // slicelike.rs:23
        Slice r;
        tb_handle h = 0;
        TBH_CALL(tb_buf_wrap(Abi<F>::dtype, const_cast<F*>(s), len, 0, &h));
        r.v_ = tb_view{h, 0, len}; r.root_ = true; r.host_ = const_cast<F*>(s); r.mut_ = false;
        return r;
    }
    static Slice new_mut(F* s, size_t len) {                    // slicelike.rs:27
        Slice r;
        tb_handle h = 0;
        TBH_CALL(tb_buf_wrap(Abi<F>::dtype, s, len, 1, &h));
        r.v_ = tb_view{h, 0, len}; r.root_ = true; r.host_ = s; r.mut_ = true;
        return r;
    }
    // view over a device-only buffer (matrices generated in HBM); not owning
    static Slice from_view(tb_view v) {
        Slice r;
        r.v_ = v; r.root_ = false; r.host_ = nullptr; r.mut_ = false;
        return r;
    }

    // split_ref / split_mut (slicelike.rs:31-37)
    std::pair<Slice, Slice> split(size_t mid) const {
        if (mid > v_.len) throw BackendError("split: mid > len");
        Slice a = raw_sub(0, mid), b = raw_sub(mid, v_.len - mid);
        if (shim_protocol() && host_) {                          // b200_slice.rs split_ref / split_mut
            const int n = (mid > 0) + (v_.len - mid > 0);
            if (n > 0) TBH_CALL(tb_buf_retain(view().buf, n));
            a.child_ = mid > 0;
            b.child_ = v_.len - mid > 0;
        }
        return {std::move(a), std::move(b)};
    }
    // One step of the splitm! / splitm_mut! macros (slicelike.rs:162-201): `(part; len)` at the front of what is left of
    // the parent.  In shim-protocol mode this is ONE split of the binding: its two children are `part` and the remainder,
    // both released when `part` goes out of scope (the macro keeps the remainder alive exactly as long).
    Slice sub(size_t off, size_t len) const {
        Slice r = raw_sub(off, len);
        if (shim_protocol() && host_) {
            const size_t rest = v_.len - (off + len);
            const int n = (len > 0) + (rest > 0);
            if (n > 0) TBH_CALL(tb_buf_retain(view().buf, n));
            r.child_ = len > 0;
            r.rest_host_ = host_ + off + len;
            r.rest_len_ = rest;
        }
        return r;
    }
    // SliceLike::drop (slicelike.rs:41): only the root has anything to reconcile
    void drop() {
        if (root_ && v_.buf > 0) {
            tb_buf_release(shim_protocol() && host_ && v_.len > 0 ? view().buf : v_.buf);
            root_ = false;
            v_ = tb_view{0, 0, 0};
        }
        if (child_) {                                            // b200_slice.rs drop(): resolve, then release
            child_ = false;
            tb_view cv{0, 0, 0};
            if (tb_view_of_host(Abi<F>::dtype, host_, v_.len, &cv) == TB_OK) tb_buf_release(cv.buf);
        }
        if (rest_len_ > 0) {
            tb_view rv{0, 0, 0};
            if (tb_view_of_host(Abi<F>::dtype, rest_host_, rest_len_, &rv) == TB_OK) tb_buf_release(rv.buf);
            rest_len_ = 0;
        }
    }
    size_t len() const { return v_.len; }                        // slicelike.rs:44
    const F* get_ref() const {                                   // slicelike.rs:47
        if (!host_) throw BackendError("get_ref on a device-only slice");
        TBH_CALL(tb_host_ref(view()));
        return host_;
    }
    F* get_mut() {                                               // slicelike.rs:50
        if (!host_ || !mut_) throw BackendError("get_mut on a read-only slice");
        TBH_CALL(tb_host_mut(view()));
        return host_;
    }
    F get(size_t idx) const {                                    // slicelike.rs:54-59
        F out;
        TBH_CALL(Abi<F>::get1(view(), idx, &out));
        return out;
    }
    void set(size_t idx, F val) {                                // slicelike.rs:63-68
        TBH_CALL(Abi<F>::set1(view(), idx, val));
    }
    // the device view of this slice; in shim-protocol mode resolved from the host address like B200Slice::view()
    tb_view view() const {
        if (shim_protocol() && host_ && v_.len > 0) {
            tb_view v{0, 0, 0};
            TBH_CALL(tb_view_of_host(Abi<F>::dtype, host_, v_.len, &v));
            if (v.buf != v_.buf || v.off != v_.off || v.len != v_.len) throw BackendError("shim protocol: tb_view_of_host resolved to a different view");
            return v;
        }
        return v_;
    }

private:
    Slice raw_sub(size_t off, size_t len) const {
        if (off > v_.len || len > v_.len - off) throw BackendError("sub-slice out of range");
        Slice r;
        r.v_ = tb_view{v_.buf, v_.off + off, len};
        r.root_ = false;
        r.host_ = host_ ? host_ + off : nullptr;
        r.mut_ = mut_;
        return r;
    }
    tb_view v_{0, 0, 0};
    bool root_ = false;
    F* host_ = nullptr;
    bool mut_ = false;
    // shim-protocol bookkeeping: this wrapper holds one reference on its root (child_), and one more for the remainder
    // sibling of the split that produced it (rest_host_, rest_len_)
    bool child_ = false;
    F* rest_host_ = nullptr;
    size_t rest_len_ = 0;
};

// ---------------------------------------------------------------------------------------------------------
// B200<F>: LinAlg + LinAlgEx as associated (static) functions, like the trait.
// ---------------------------------------------------------------------------------------------------------
template <typename F> struct B200 {
    using Sl = Slice<F>;

    static F norm(const Sl& x) {                                                     // linalg.rs:27
        F r;
        TBH_CALL(Abi<F>::norm(x.view(), &r));
        return r;
    }
    static void copy(const Sl& x, Sl& y) { TBH_CALL(Abi<F>::copy(x.view(), y.view())); }              // linalg.rs:33
    static void scale(F alpha, Sl& x) { TBH_CALL(Abi<F>::scale(alpha, x.view())); }                   // linalg.rs:39
    static void add(F alpha, const Sl& x, Sl& y) { TBH_CALL(Abi<F>::add(alpha, x.view(), y.view())); } // linalg.rs:46
    static void adds(F s, Sl& y) { TBH_CALL(Abi<F>::adds(s, y.view())); }                             // linalg.rs:52
    static F abssum(const Sl& x, size_t incx) {                                                       // linalg.rs:58
        F r;
        TBH_CALL(Abi<F>::abssum(x.view(), incx, &r));
        return r;
    }
    static void transform_di(F alpha, const Sl& mat, const Sl& x, F beta, Sl& y) {                    // linalg.rs:67
        TBH_CALL(Abi<F>::transform_di(alpha, mat.view(), x.view(), beta, y.view()));
    }
    static void transform_ge(bool transpose, size_t n_row, size_t n_col, F alpha, const Sl& mat, const Sl& x, F beta, Sl& y) {  // linalg_ex.rs:23
        TBH_CALL(Abi<F>::transform_ge(transpose ? 1 : 0, n_row, n_col, alpha, mat.view(), x.view(), beta, y.view()));
    }
    static void transform_sp(size_t n, F alpha, const Sl& mat, const Sl& x, F beta, Sl& y) {          // linalg_ex.rs:37
        TBH_CALL(Abi<F>::transform_sp(n, alpha, mat.view(), x.view(), beta, y.view()));
    }
    static size_t map_eig_worklen(size_t n) { return tb_map_eig_worklen(n); }                          // linalg_ex.rs:43

    // linalg_ex.rs:64.  `map(e, out)` returns false for None, or true with the replacement eigenvalue in `out`.
    // Like dsyevr(range=V, vl=0, vu=+inf) in the CPU twin (f64lapack.rs:86-91) only eigenvalues > 0 reach the closure.
    // The two closures the reference itself passes to map_eig are recognised by probing them on a fixed set of positive
    // arguments spanning the floating-point range (only eigenvalues > 0 ever reach the closure, f64lapack.rs:86-107):
    //   e -> Some(e)        ConePSD::proj (cone_psd.rs:69-76), with scale_diag = Some(sqrt 2)  -> tb_proj_psd (GEMM-only sign iteration)
    //   e -> Some(sqrt(e))  MatBuild::set_sqrt (matbuild/mod.rs:231-238), scale_diag = None    -> tb_sqrt_psd (GEMM-only Newton-Schulz)
    // so that the UNMODIFIED ConePSD / ProbQP / ProbQCQP get the fast paths through the trait surface; any other closure
    // takes the general route below (eigendecomposition on the device, eigenvalues through the closure on the host).
    enum class Closure { General, KeepPositive, Sqrt };
    static Closure recognise(const std::function<bool(F, F&)>& map) {
        static const double probes[] = {1e-30, 3e-21, 1e-12, 7e-7, 1e-3, 0.25, 1.0, 2.0, 9.0, 1234.5, 1e6, 3e12, 1e20, 1e30};
        bool keep = true, root = true;
        for (double pd : probes) {
            const F e = (F)pd;
            F out = F(0);
            if (!map(e, out)) return Closure::General;
            keep = keep && out == e;
            root = root && out == (F)std::sqrt(e);
        }
        return keep ? Closure::KeepPositive : root ? Closure::Sqrt : Closure::General;
    }
    static void map_eig(Sl& mat, bool has_scale, F scale_diag, F eps_zero, Sl& work, const std::function<bool(F, F&)>& map) {
        const size_t sn = mat.len();
        size_t n = 0;
        while ((n + 1) * (n + 2) / 2 <= sn) ++n;
        if (n * (n + 1) / 2 != sn) throw BackendError("map_eig: length is not a triangular number");
        if (map_eig_fast_paths()) {
            const Closure cl = recognise(map);
            if (cl == Closure::KeepPositive && has_scale && scale_diag == (F)std::sqrt(F(2))) {
                TBH_CALL(Abi<F>::proj_psd(mat.view(), eps_zero, work.view()));
                return;
            }
            if (cl == Closure::Sqrt && !has_scale) {
                TBH_CALL(Abi<F>::sqrt_psd(mat.view(), eps_zero, work.view()));
                return;
            }
        }
        std::vector<F> eigs(n), neweigs(n);
        std::vector<uint8_t> keep(n);
        TBH_CALL(Abi<F>::map_eig_begin(mat.view(), has_scale ? 1 : 0, scale_diag, eps_zero, work.view(), eigs.data()));
        for (size_t i = 0; i < n; ++i) {
            F out = F(0);
            bool k = eigs[i] > F(0) && map(eigs[i], out);
            keep[i] = k ? 1 : 0;
            neweigs[i] = k ? out : F(0);
        }
        TBH_CALL(Abi<F>::map_eig_finish(mat.view(), has_scale ? 1 : 0, scale_diag, work.view(), neweigs.data(), keep.data()));
    }
};

}  // namespace totsu_b200
