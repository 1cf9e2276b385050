// C++ mirror of the solver-side code that calls the backend, so the hot path can be driven and measured in
// this image (no rustc): the Operator / Cone traits, MatOp, the five cones and the first-order solver loop.
// In a Rust build these are the reference's own, unmodified sources; here every function restates the cited
// lines 1:1 (same call order, same scalars crossing the host boundary) so that measured iterations/second
// transfer to `totsu_core::solver::Solver<B200>`.
//
//   Operator     totsu_core/src/solver/operator.rs:11-156
//   Cone         totsu_core/src/solver/cone.rs:9-30
//   MatOp        totsu_core/src/matop.rs:9-175
//   ConeZero/RPos/SOC/RotSOC/PSD   totsu_core/src/cone_{zero,rpos,soc,rotsoc,psd}.rs
//   Solver       totsu_core/src/solver/solver.rs:13-657
#pragma once
#include <cmath>
#include <limits>
#include <memory>
#include <optional>
#include "linalg.hpp"

namespace totsu_b200 {

enum class SolverError { None = 0, Unbounded, Infeasible, ExcessIter, InvalidOp, WorkShortage, ConeFailure };   // solver_error.rs:3-18

template <typename F> struct Operator {                          // operator.rs:11-156
    using Sl = Slice<F>;
    virtual ~Operator() = default;
    virtual std::pair<size_t, size_t> size() const = 0;
    virtual void op(F alpha, const Sl& x, F beta, Sl& y) const = 0;
    virtual void trans_op(F alpha, const Sl& x, F beta, Sl& y) const = 0;
    virtual void absadd_cols(Sl& tau) const = 0;
    virtual void absadd_rows(Sl& sigma) const = 0;
};

template <typename F> struct Cone {                              // cone.rs:9-30
    using Sl = Slice<F>;
    using Group = std::function<void(Sl&)>;
    virtual ~Cone() = default;
    virtual bool proj(bool dual_cone, Sl& x) = 0;                // false <=> Err(())
    virtual void product_group(Sl& dp_tau, const Group& group) const = 0;
};

// ---------------------------------------------------------------------------------------------------------
struct MatType {                                                 // matop.rs:9-40
    enum Kind { General, SymPack } kind;
    size_t a, b;
    static MatType general(size_t n_row, size_t n_col) { return MatType{General, n_row, n_col}; }
    static MatType sympack(size_t n) { return MatType{SymPack, n, n}; }
    size_t len() const { return kind == General ? a * b : a * (a + 1) / 2; }
    std::pair<size_t, size_t> size() const { return {a, b}; }
};

template <typename F> class MatOp : public Operator<F> {         // matop.rs:44-175
public:
    using L = B200<F>;
    using Sl = Slice<F>;
    MatOp(MatType typ, const F* array, size_t len) : typ_(typ), array_(Sl::new_ref(array, len)) {   // matop.rs:64-74 (upload point)
        if (typ.len() != len) throw BackendError("MatOp: typ.len() != array.len()");
    }
    // matrix already resident in a device buffer
    MatOp(MatType typ, tb_view dev) : typ_(typ), array_(Sl::from_view(dev)) {
        if (typ.len() != dev.len) throw BackendError("MatOp: typ.len() != array.len()");
    }
    std::pair<size_t, size_t> size() const override { return typ_.size(); }
    void op(F alpha, const Sl& x, F beta, Sl& y) const override { op_impl(false, alpha, x, beta, y); }
    void trans_op(F alpha, const Sl& x, F beta, Sl& y) const override { op_impl(true, alpha, x, beta, y); }
    void absadd_cols(Sl& tau) const override { absadd_impl(true, tau); }
    void absadd_rows(Sl& sigma) const override { absadd_impl(false, sigma); }
    const Sl& array() const { return array_; }

private:
    void op_impl(bool transpose, F alpha, const Sl& x, F beta, Sl& y) const {       // matop.rs:76-96
        if (typ_.kind == MatType::General) {
            if (typ_.a > 0 && typ_.b > 0) L::transform_ge(transpose, typ_.a, typ_.b, alpha, array_, x, beta, y);
            else L::scale(beta, y);
        } else {
            if (typ_.a > 0) L::transform_sp(typ_.a, alpha, array_, x, beta, y);
            else L::scale(beta, y);
        }
    }
    void absadd_impl(bool colwise, Sl& y) const {                                   // matop.rs:98-138
        if (typ_.kind == MatType::General) {
            const size_t nr = typ_.a, nc = typ_.b;
            if (colwise) {
                if (nc != y.len()) throw BackendError("absadd_cols: length mismatch");
                F* ym = y.get_mut();
                for (size_t i = 0; i < nc; ++i) {
                    Sl col = array_.sub(i * nr, nr);
                    ym[i] = L::abssum(col, 1) + ym[i];
                }
            } else {
                if (nr != y.len()) throw BackendError("absadd_rows: length mismatch");
                F* ym = y.get_mut();
                for (size_t i = 0; i < nr; ++i) {
                    Sl row = array_.sub(i, nr * nc - i);
                    ym[i] = L::abssum(row, nr) + ym[i];
                }
            }
        } else {
            const size_t n = typ_.a;
            if (n != y.len()) throw BackendError("absadd: length mismatch");
            F* ym = y.get_mut();
            size_t sum = 0;
            for (size_t c = 0; c < n; ++c) {
                Sl col = array_.sub(sum, c + 1);
                sum += c + 1;
                ym[c] = L::abssum(col, 1) + ym[c];
                const F* cr = col.get_ref();
                for (size_t i = 0; i < c; ++i) ym[i] = ym[i] + std::fabs(cr[i]);
            }
        }
    }
    MatType typ_;
    Sl array_;
};

// ---------------------------------------------------------------------------------------------------------
template <typename F> struct ConeZero : Cone<F> {                // cone_zero.rs:38-49
    using Sl = Slice<F>;
    bool proj(bool dual_cone, Sl& x) override {
        if (!dual_cone) B200<F>::scale(F(0), x);
        return true;
    }
    void product_group(Sl&, const typename Cone<F>::Group&) const override {}
};

template <typename F> struct ConeRPos : Cone<F> {                // cone_rpos.rs:38-50 (host loop via get_mut)
    using Sl = Slice<F>;
    bool proj(bool, Sl& x) override {
        F* xm = x.get_mut();
        for (size_t i = 0; i < x.len(); ++i) xm[i] = xm[i] > F(0) ? xm[i] : F(0);
        return true;
    }
    void product_group(Sl&, const typename Cone<F>::Group&) const override {}
};

template <typename F> struct ConeSOC : Cone<F> {                 // cone_soc.rs:38-70
    using Sl = Slice<F>;
    bool proj(bool, Sl& x) override {
        using L = B200<F>;
        if (x.len() > 0) {
            auto sv = x.split(1);
            Sl& s = sv.first;
            Sl& v = sv.second;
            const F val_s = s.get(0);
            const F norm_v = L::norm(v);
            if (norm_v <= -val_s) {
                L::scale(F(0), v);
                s.set(0, F(0));
            } else if (norm_v <= val_s) {
                // as they are
            } else {
                const F alpha = (F(1) + val_s / norm_v) / F(2);
                L::scale(alpha, v);
                s.set(0, (norm_v + val_s) / F(2));
            }
        }
        return true;
    }
    void product_group(Sl& dp_tau, const typename Cone<F>::Group& group) const override { group(dp_tau); }
};

template <typename F> struct ConeRotSOC : Cone<F> {              // cone_rotsoc.rs:38-70
    using Sl = Slice<F>;
    ConeSOC<F> soc;
    bool proj(bool dual_cone, Sl& x) override {
        const F fsqrt2 = std::sqrt(F(2));
        if (x.len() > 0) {
            if (x.len() == 1) {
                const F r = x.get(0);
                x.set(0, r > F(0) ? r : F(0));
            } else {
                F r = x.get(0), s = x.get(1);
                x.set(0, (r + s) / fsqrt2);
                x.set(1, (r - s) / fsqrt2);
                if (!soc.proj(dual_cone, x)) return false;
                r = x.get(0); s = x.get(1);
                x.set(0, (r + s) / fsqrt2);
                x.set(1, (r - s) / fsqrt2);
            }
        }
        return true;
    }
    void product_group(Sl& dp_tau, const typename Cone<F>::Group& group) const override { group(dp_tau); }
};

template <typename F> class ConePSD : public Cone<F> {           // cone_psd.rs:9-85
public:
    using Sl = Slice<F>;
    static size_t query_worklen(size_t nvars) {                  // cone_psd.rs:32-38
        size_t n = 0;
        while ((n + 1) * (n + 2) / 2 <= nvars) ++n;
        if (n * (n + 1) / 2 != nvars) throw BackendError("ConePSD: nvars is not a triangular number");
        return B200<F>::map_eig_worklen(n);
    }
    ConePSD(F* work, size_t work_len, F eps_zero) : work_(Sl::new_mut(work, work_len)), eps_zero_(eps_zero) {}
    bool proj(bool, Sl& x) override {                            // cone_psd.rs:56-79
        if (work_.len() < query_worklen(x.len())) return false;
        const F fsqrt2 = std::sqrt(F(2));
        B200<F>::map_eig(x, true, fsqrt2, eps_zero_, work_, [](F e, F& out) {
            if (e > F(0)) { out = e; return true; }
            return false;
        });
        return true;
    }
    void product_group(Sl& dp_tau, const typename Cone<F>::Group& group) const override { group(dp_tau); }

private:
    Sl work_;
    F eps_zero_;
};

// ---------------------------------------------------------------------------------------------------------
template <typename F> struct SolverParam {                       // solver.rs:13-41
    std::optional<size_t> max_iter;
    F eps_acc = F(1e-6), eps_inf = F(1e-6), eps_zero = F(1e-12);
    size_t log_period = 10000;
};

template <typename F> struct IterInfo {
    size_t i = 0;
    bool conv_branch = true;         // true: criteria_conv ran (tau > eps_zero), false: criteria_inf
    F val_tau = 0, c0 = 0, c1 = 0, c2 = 0;   // (cri_pri, cri_dual, cri_gap) or (cri_unbdd, cri_infeas, -)
};

template <typename F> class SelfDualEmbed {                      // solver.rs:45-184
public:
    using L = B200<F>;
    using Sl = Slice<F>;
    SelfDualEmbed(const Operator<F>& c, const Operator<F>& a, const Operator<F>& b) : c_(c), a_(a), b_(b) {}
    const Operator<F>& c() const { return c_; }
    const Operator<F>& a() const { return a_; }
    const Operator<F>& b() const { return b_; }

    static F fr_norm(const Operator<F>& op, Sl& work_v, Sl& work_t) {          // solver.rs:85-107
        if (work_v.len() != op.size().second || work_t.len() != op.size().first) throw BackendError("fr_norm: size mismatch");
        L::scale(F(0), work_v);
        F sq_norm = F(0);
        for (size_t row = 0; row < op.size().second; ++row) {
            work_v.set(row, F(1));
            op.op(F(1), work_v, F(0), work_t);
            const F n = L::norm(work_t);
            sq_norm = sq_norm + n * n;
            work_v.set(row, F(0));
        }
        return std::sqrt(sq_norm);
    }

    void op(F alpha, const Sl& x, F beta, Sl& y) const {                         // solver.rs:109-131
        const auto [m, n] = a_.size();
        if (x.len() != n + m + m + 1 || y.len() != n + m + 1) throw BackendError("SelfDualEmbed::op: size mismatch");
        Sl x_x = x.sub(0, n), x_y = x.sub(n, m), x_s = x.sub(n + m, m), x_tau = x.sub(n + 2 * m, 1);
        Sl y_n = y.sub(0, n), y_m = y.sub(n, m), y_1 = y.sub(n + m, 1);
        a_.trans_op(alpha, x_y, beta, y_n);
        c_.op(alpha, x_tau, F(1), y_n);
        a_.op(-alpha, x_x, beta, y_m);
        L::add(-alpha, x_s, y_m);
        b_.op(alpha, x_tau, F(1), y_m);
        c_.trans_op(-alpha, x_x, beta, y_1);
        b_.trans_op(-alpha, x_y, F(1), y_1);
    }

    void trans_op(F alpha, const Sl& x, F beta, Sl& y) const {                   // solver.rs:133-157
        const auto [m, n] = a_.size();
        if (x.len() != n + m + 1 || y.len() != n + m + m + 1) throw BackendError("SelfDualEmbed::trans_op: size mismatch");
        Sl x_n = x.sub(0, n), x_m = x.sub(n, m), x_1 = x.sub(n + m, 1);
        Sl y_x = y.sub(0, n), y_y = y.sub(n, m), y_s = y.sub(n + m, m), y_tau = y.sub(n + 2 * m, 1);
        a_.trans_op(-alpha, x_m, beta, y_x);
        c_.op(-alpha, x_1, F(1), y_x);
        a_.op(alpha, x_n, beta, y_y);
        b_.op(-alpha, x_1, F(1), y_y);
        L::scale(beta, y_s);
        L::add(-alpha, x_m, y_s);
        c_.trans_op(alpha, x_n, beta, y_tau);
        b_.trans_op(alpha, x_m, F(1), y_tau);
    }

    void abssum(Sl& tau, Sl& sigma) const {                                      // solver.rs:159-183
        const auto [m, n] = a_.size();
        L::scale(F(0), tau);
        Sl tau_x = tau.sub(0, n), tau_y = tau.sub(n, m), tau_s = tau.sub(n + m, m), tau_tau = tau.sub(n + 2 * m, 1);
        a_.absadd_cols(tau_x);
        c_.absadd_rows(tau_x);
        a_.absadd_rows(tau_y);
        b_.absadd_rows(tau_y);
        L::adds(F(1), tau_s);
        c_.absadd_cols(tau_tau);
        b_.absadd_cols(tau_tau);
        Sl sigma_n = sigma.sub(0, n), sigma_m = sigma.sub(n, m), sigma_1 = sigma.sub(n + m, 1);
        L::copy(tau_x, sigma_n);
        L::copy(tau_y, sigma_m);
        L::add(F(1), tau_s, sigma_m);
        L::copy(tau_tau, sigma_1);
    }

private:
    const Operator<F>& c_;
    const Operator<F>& a_;
    const Operator<F>& b_;
};

// The solver.  `solve()` is solver.rs:285-321 + SolverCore::solve (solver.rs:340-457); `begin()/step()` expose
// the same loop body one iteration at a time so bench.py can time exactly K iterations.
template <typename F> class Solver {
public:
    using L = B200<F>;
    using Sl = Slice<F>;
    using Trace = std::function<void(const IterInfo<F>&, bool logged)>;

    SolverParam<F> par;
    Trace trace;                     // called once per iteration (the log::debug!/trace! lines of solver.rs:391-394,426-429)
    bool device_precond = false;     // calc_precond's two host loops (solver.rs:501-506) done by one kernel each

    static size_t query_worklen(std::pair<size_t, size_t> op_a_size) {           // solver.rs:231-249
        const size_t m = op_a_size.first, n = op_a_size.second;
        return (n + m + m + 1) * 4 + (n + m + 1) * 2;
    }

    // solver.rs:285-321.  On success (or ExcessIter) work[0..n] = x, work[n..n+m] = y after the work slice drops.
    SolverError solve(const Operator<F>& op_c, const Operator<F>& op_a, const Operator<F>& op_b, Cone<F>& cone, F* work, size_t work_len) {
        SolverError e = begin(op_c, op_a, op_b, cone, work, work_len);
        if (e != SolverError::None) return e;
        bool done = false;
        while (!done) e = step(done);
        end();
        return e;
    }

    SolverError begin(const Operator<F>& op_c, const Operator<F>& op_a, const Operator<F>& op_b, Cone<F>& cone, F* work, size_t work_len) {
        const auto [m, n] = op_a.size();
        if (op_c.size() != std::make_pair(n, size_t(1)) || op_b.size() != std::make_pair(m, size_t(1))) return SolverError::InvalidOp;   // solver.rs:292-295
        if (query_worklen({m, n}) > work_len) return SolverError::WorkShortage;                                                      // solver.rs:297-300
        m_ = m; n_ = n;
        op_k_.reset(new SelfDualEmbed<F>(op_c, op_a, op_b));
        cone_ = &cone;
        work_ = Sl::new_mut(work, work_len);                                     // solver.rs:315
        // ---- SolverCore::solve prologue, solver.rs:342-359
        calc_norms();
        const size_t lx = n + m + m + 1, ly = n + m + 1;
        size_t o = 0;
        x_ = work_.sub(o, lx); o += lx;
        y_ = work_.sub(o, ly); o += ly;
        dp_tau_ = work_.sub(o, lx); o += lx;
        dp_sigma_ = work_.sub(o, ly); o += ly;
        tmpw_ = work_.sub(o, 2 * lx);
        init_vecs();
        calc_precond();
        i_ = 0;
        return SolverError::None;
    }

    // one pass of the loop body, solver.rs:364-456
    SolverError step(bool& done) {
        done = false;
        const bool excess_iter = par.max_iter ? (i_ + 1 >= *par.max_iter) : false;
        const bool log_trig = par.log_period > 0 ? (i_ % par.log_period == 0) : false;
        F val_tau;
        if (!update_vecs(val_tau)) { done = true; return SolverError::ConeFailure; }
        IterInfo<F> info;
        info.i = i_; info.val_tau = val_tau;
        if (val_tau > par.eps_zero) {
            F cri_pri, cri_dual, cri_gap;
            criteria_conv(cri_pri, cri_dual, cri_gap);
            const bool term_conv = (cri_pri <= par.eps_acc) && (cri_dual <= par.eps_acc) && (cri_gap <= par.eps_acc);
            info.conv_branch = true; info.c0 = cri_pri; info.c1 = cri_dual; info.c2 = cri_gap;
            if (trace) trace(info, log_trig || excess_iter || term_conv);
            last_ = info;
            if (excess_iter || term_conv) {
                Sl x_x_ast = x_.sub(0, n_), x_y_ast = x_.sub(n_, m_);
                L::scale(F(1) / val_tau, x_x_ast);
                L::scale(F(1) / val_tau, x_y_ast);
                done = true;
                return term_conv ? SolverError::None : SolverError::ExcessIter;
            }
        } else {
            F cri_unbdd, cri_infeas;
            criteria_inf(cri_unbdd, cri_infeas);
            const bool term_unbdd = cri_unbdd <= par.eps_inf, term_infeas = cri_infeas <= par.eps_inf;
            info.conv_branch = false; info.c0 = cri_unbdd; info.c1 = cri_infeas;
            if (trace) trace(info, log_trig || excess_iter || term_unbdd || term_infeas);
            last_ = info;
            if (excess_iter || term_unbdd || term_infeas) {
                done = true;
                if (term_unbdd) return SolverError::Unbounded;
                if (term_infeas) return SolverError::Infeasible;
                return SolverError::ExcessIter;
            }
        }
        ++i_;
        return SolverError::None;
    }

    // drop the work slice: device-newer ranges flow back to the caller's `work` (solver.rs:317-320)
    void end() {
        x_ = Sl(); y_ = Sl(); dp_tau_ = Sl(); dp_sigma_ = Sl(); tmpw_ = Sl();
        work_ = Sl();
        op_k_.reset();
    }

    size_t iterations() const { return i_; }
    const IterInfo<F>& last() const { return last_; }
    const Sl& x() const { return x_; }
    const Sl& y() const { return y_; }
    F norm_b() const { return norm_b_; }
    F norm_c() const { return norm_c_; }

private:
    void calc_norms() {                                                          // solver.rs:460-481
        F work1[1] = {F(0)};
        Sl work_one = Sl::new_mut(work1, 1);
        {
            const size_t m = op_k_->b().size().first;
            Sl t = work_.sub(0, m);
            norm_b_ = SelfDualEmbed<F>::fr_norm(op_k_->b(), work_one, t);
        }
        {
            const size_t n = op_k_->c().size().first;
            Sl t = work_.sub(0, n);
            norm_c_ = SelfDualEmbed<F>::fr_norm(op_k_->c(), work_one, t);
        }
    }

    void init_vecs() {                                                           // solver.rs:483-494
        L::scale(F(0), x_);
        L::scale(F(0), y_);
        x_.set(n_ + m_ + m_, F(1));
    }

    void calc_precond() {                                                        // solver.rs:496-524
        op_k_->abssum(dp_tau_, dp_sigma_);
        if (device_precond) {
            TBH_CALL(Abi<F>::recip_clamp(par.eps_zero, dp_tau_.view()));
            TBH_CALL(Abi<F>::recip_clamp(par.eps_zero, dp_sigma_.view()));
        } else {
            F* t = dp_tau_.get_mut();
            for (size_t k = 0; k < dp_tau_.len(); ++k) t[k] = F(1) / (t[k] > par.eps_zero ? t[k] : par.eps_zero);
            F* s = dp_sigma_.get_mut();
            for (size_t k = 0; k < dp_sigma_.len(); ++k) s[k] = F(1) / (s[k] > par.eps_zero ? s[k] : par.eps_zero);
        }
        // grouping dependent on cone
        typename Cone<F>::Group group = [](Sl& tau_group) {
            if (tau_group.len() > 0) {
                F* tg = tau_group.get_mut();
                F min_t = tg[0];
                for (size_t k = 0; k < tau_group.len(); ++k) min_t = tg[k] < min_t ? tg[k] : min_t;
                for (size_t k = 0; k < tau_group.len(); ++k) tg[k] = min_t;
            }
        };
        Sl dpt_dual_cone = dp_tau_.sub(n_, m_), dpt_cone = dp_tau_.sub(n_ + m_, m_);
        cone_->product_group(dpt_dual_cone, group);
        cone_->product_group(dpt_cone, group);
    }

    bool update_vecs(F& val_tau) {                                               // solver.rs:526-571
        const size_t lx = x_.len();
        Sl rx = tmpw_.sub(0, lx), tx = tmpw_.sub(lx, lx);
        L::copy(x_, rx);
        op_k_->trans_op(F(-1), y_, F(0), tx);
        L::transform_di(F(1), dp_tau_, tx, F(1), x_);
        {
            Sl x_y = x_.sub(n_, m_), x_s = x_.sub(n_ + m_, m_), x_tau = x_.sub(n_ + 2 * m_, 1);
            if (!cone_->proj(true, x_y)) return false;
            if (!cone_->proj(false, x_s)) return false;
            const F t = x_tau.get(0);
            val_tau = t > F(0) ? t : F(0);
            x_tau.set(0, val_tau);
        }
        L::add(F(-2), x_, rx);
        {
            Sl ty = tx.sub(0, y_.len());
            op_k_->op(F(-1), rx, F(0), ty);
            L::transform_di(F(1), dp_sigma_, ty, F(1), y_);
        }
        {
            Sl y_1 = y_.sub(n_ + m_, 1);
            const F k = y_1.get(0);
            y_1.set(0, k < F(0) ? k : F(0));
        }
        return true;
    }

    void criteria_conv(F& cri_pri, F& cri_dual, F& cri_gap) {                    // solver.rs:573-612
        Sl x_x = x_.sub(0, n_), x_y = x_.sub(n_, m_), x_s = x_.sub(n_ + m_, m_), x_tau = x_.sub(n_ + 2 * m_, 1);
        Sl p = tmpw_.sub(0, m_), d = tmpw_.sub(m_, n_);
        const F val_tau = x_tau.get(0);
        F work1[1] = {F(1)};
        Sl work_one = Sl::new_mut(work1, 1);
        L::copy(x_s, p);
        op_k_->b().op(F(-1), work_one, F(1) / val_tau, p);
        op_k_->a().op(F(1) / val_tau, x_x, F(1), p);
        op_k_->c().op(F(1), work_one, F(0), d);
        op_k_->a().trans_op(F(1) / val_tau, x_y, F(1), d);
        op_k_->c().trans_op(F(1) / val_tau, x_x, F(0), work_one);
        const F g_x = work_one.get(0);
        op_k_->b().trans_op(F(1) / val_tau, x_y, F(0), work_one);
        const F g_y = work_one.get(0);
        const F g = g_x + g_y;
        cri_pri = L::norm(p) / (F(1) + norm_b_);
        cri_dual = L::norm(d) / (F(1) + norm_c_);
        cri_gap = std::fabs(g) / (F(1) + std::fabs(g_x) + std::fabs(g_y));
    }

    void criteria_inf(F& cri_unbdd, F& cri_infeas) {                             // solver.rs:614-656
        Sl x_x = x_.sub(0, n_), x_y = x_.sub(n_, m_), x_s = x_.sub(n_ + m_, m_);
        Sl p = tmpw_.sub(0, m_), d = tmpw_.sub(m_, n_);
        F work1[1] = {F(0)};
        Sl work_one = Sl::new_mut(work1, 1);
        L::copy(x_s, p);
        op_k_->a().op(F(1), x_x, F(1), p);
        op_k_->a().trans_op(F(1), x_y, F(0), d);
        op_k_->c().trans_op(F(-1), x_x, F(0), work_one);
        const F m_cx = work_one.get(0);
        op_k_->b().trans_op(F(-1), x_y, F(0), work_one);
        const F m_by = work_one.get(0);
        const F finf = std::numeric_limits<F>::infinity();
        cri_unbdd = m_cx > par.eps_zero ? L::norm(p) * norm_c_ / m_cx : finf;
        cri_infeas = m_by > par.eps_zero ? L::norm(d) * norm_b_ / m_by : finf;
    }

    size_t m_ = 0, n_ = 0, i_ = 0;
    std::unique_ptr<SelfDualEmbed<F>> op_k_;
    Cone<F>* cone_ = nullptr;
    Sl work_, x_, y_, dp_tau_, dp_sigma_, tmpw_;
    F norm_b_ = 0, norm_c_ = 0;
    IterInfo<F> last_;
};

}  // namespace totsu_b200
