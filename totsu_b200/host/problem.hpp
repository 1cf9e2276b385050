// C++ mirror of the reference's problem front-ends (callers of the hot path):
//   MatBuild   totsu/src/matbuild/mod.rs:16-272
//   ProbLP     totsu/src/problem/lp.rs      ProbQP   totsu/src/problem/qp.rs     ProbQCQP  totsu/src/problem/qcqp.rs
//   ProbSOCP   totsu/src/problem/socp.rs    ProbSDP  totsu/src/problem/sdp.rs
// They stack per-block MatOps with sub-slices and L::scale / L::add exactly like the reference, so that the stock
// (un-fused) path through the backend - one transform_ge per block - is exercised and parity-checked.
#pragma once
#include "solver.hpp"

namespace totsu_b200 {

template <typename F> class MatBuild {                           // matbuild/mod.rs:16-272
public:
    using L = B200<F>;
    MatBuild() : typ_(MatType::general(0, 0)) {}
    explicit MatBuild(MatType typ) : typ_(typ), array_(typ.len(), F(0)) {}
    MatBuild(MatType typ, const F* data) : typ_(typ), array_(data, data + typ.len()) {}
    std::pair<size_t, size_t> size() const { return typ_.size(); }
    bool is_sympack() const { return typ_.kind == MatType::SymPack; }
    MatType typ() const { return typ_; }
    std::vector<F>& array() { return array_; }
    const std::vector<F>& array() const { return array_; }
    size_t index(size_t r, size_t c) const {                     // matbuild/mod.rs:249-272
        if (typ_.kind == MatType::General) return c * typ_.a + r;
        if (r > c) std::swap(r, c);
        return c * (c + 1) / 2 + r;
    }
    F& at(size_t r, size_t c) { return array_[index(r, c)]; }
    std::unique_ptr<MatOp<F>> as_op() const {                    // matbuild/mod.rs:47-50
        return std::unique_ptr<MatOp<F>>(new MatOp<F>(typ_, array_.data(), array_.size()));
    }
    void set_scale_nondiag(F alpha) {                            // matbuild/mod.rs:143-175 (SymPack arm)
        if (!is_sympack()) throw BackendError("set_scale_nondiag: SymPack only");
        const size_t n = typ_.a;
        for (size_t c = 0; c + 1 < n; ++c) {
            size_t i = index(c, c), ii = index(c + 1, c + 1);
            if (ii - i - 1 == 0) continue;
            Slice<F> s = Slice<F>::new_mut(array_.data() + i + 1, ii - i - 1);
            L::scale(alpha, s);
        }
    }
    void set_reshape_colvec() { typ_ = MatType::general(array_.size(), 1); }     // matbuild/mod.rs:182-186
    void set_sqrt(F eps_zero) {                                  // matbuild/mod.rs:220-241
        if (!is_sympack()) throw BackendError("set_sqrt: SymPack only");
        const size_t n = typ_.a;
        std::vector<F> work_vec(L::map_eig_worklen(n), F(0));
        Slice<F> work = Slice<F>::new_mut(work_vec.data(), work_vec.size());
        Slice<F> a = Slice<F>::new_mut(array_.data(), array_.size());
        L::map_eig(a, false, F(1), eps_zero, work, [](F e, F& out) {
            if (e > F(0)) { out = std::sqrt(e); return true; }
            return false;
        });
    }

private:
    MatType typ_;
    std::vector<F> array_;
};

// ---------------------------------------------------------------------------------------------------------
// A problem owns its operators, cone and solver work memory: `(op_c, op_a, op_b, cone, work)` of problem().
template <typename F> struct Problem {
    // work memory first: it must outlive the cone / operators whose root slices flush into it when dropped
    std::vector<F> w_solver;
    std::vector<F> w_cone;
    std::unique_ptr<Operator<F>> op_c, op_a, op_b;
    std::unique_ptr<Cone<F>> cone;
    virtual ~Problem() = default;
};

template <typename F> struct VecOp : Operator<F> {               // ProbLPOpC lp.rs:10-46 (and SOCP/SDP OpC)
    using Sl = Slice<F>;
    std::unique_ptr<MatOp<F>> v;
    explicit VecOp(std::unique_ptr<MatOp<F>> m) : v(std::move(m)) {}
    std::pair<size_t, size_t> size() const override { return {v->size().first, 1}; }
    void op(F a, const Sl& x, F b, Sl& y) const override { v->op(a, x, b, y); }
    void trans_op(F a, const Sl& x, F b, Sl& y) const override { v->trans_op(a, x, b, y); }
    void absadd_cols(Sl& t) const override { v->absadd_cols(t); }
    void absadd_rows(Sl& s) const override { v->absadd_rows(s); }
};

// vertical stack of MatOps sharing the column count: ProbLPOpA lp.rs:49-116, ProbLPOpB lp.rs:120-186,
// ProbSDPOpA sdp.rs:49-114, ProbSDPOpB sdp.rs:118-184 (first block negated)
template <typename F> struct StackOp : Operator<F> {
    using Sl = Slice<F>;
    std::vector<std::unique_ptr<MatOp<F>>> mats;
    std::vector<F> signs;
    size_t ncol = 0;
    std::pair<size_t, size_t> size() const override {
        size_t r = 0;
        for (auto& m : mats) r += m->size().first;
        return {r, ncol};
    }
    void op(F alpha, const Sl& x, F beta, Sl& y) const override {
        size_t o = 0;
        for (size_t k = 0; k < mats.size(); ++k) {
            size_t r = mats[k]->size().first;
            Sl yk = y.sub(o, r);
            mats[k]->op(signs[k] * alpha, x, beta, yk);
            o += r;
        }
    }
    void trans_op(F alpha, const Sl& x, F beta, Sl& y) const override {
        size_t o = 0;
        for (size_t k = 0; k < mats.size(); ++k) {
            size_t r = mats[k]->size().first;
            Sl xk = x.sub(o, r);
            mats[k]->trans_op(signs[k] * alpha, xk, k == 0 ? beta : F(1), y);
            o += r;
        }
    }
    void absadd_cols(Sl& tau) const override {
        for (auto& m : mats) m->absadd_cols(tau);
    }
    void absadd_rows(Sl& sigma) const override {
        size_t o = 0;
        for (auto& m : mats) {
            size_t r = m->size().first;
            Sl sk = sigma.sub(o, r);
            m->absadd_rows(sk);
            o += r;
        }
    }
};

// sequence of (cone, length): ProbLPCone lp.rs:190-219, ProbQPCone qp.rs:260-296, ProbQCQPCone qcqp.rs:303-350,
// ProbSOCPCone socp.rs:286-333, ProbSDPCone sdp.rs:188-220
template <typename F> struct SeqCone : Cone<F> {
    using Sl = Slice<F>;
    std::vector<std::pair<std::shared_ptr<Cone<F>>, size_t>> blocks;
    bool proj(bool dual, Sl& x) override {
        size_t o = 0;
        for (auto& b : blocks) {
            Sl xb = x.sub(o, b.second);
            if (!b.first->proj(dual, xb)) return false;
            o += b.second;
        }
        return true;
    }
    void product_group(Sl& t, const typename Cone<F>::Group& g) const override {
        size_t o = 0;
        for (auto& b : blocks) {
            Sl tb = t.sub(o, b.second);
            b.first->product_group(tb, g);
            o += b.second;
        }
    }
};

// ---- LP ---------------------------------------------------------------------------------------------------
template <typename F> struct ProbLP : Problem<F> {               // lp.rs:222-338
    MatBuild<F> vec_c, mat_g, vec_h, mat_a, vec_b;
    ProbLP(MatBuild<F> c, MatBuild<F> g, MatBuild<F> h, MatBuild<F> a, MatBuild<F> b)
        : vec_c(std::move(c)), mat_g(std::move(g)), vec_h(std::move(h)), mat_a(std::move(a)), vec_b(std::move(b)) {
        const size_t n = vec_c.size().first, m = vec_h.size().first, p = vec_b.size().first;
        if (mat_g.size() != std::make_pair(m, n) || mat_a.size() != std::make_pair(p, n)) throw BackendError("ProbLP: size mismatch");
    }
    void problem() {
        const size_t m = vec_h.size().first, p = vec_b.size().first;
        this->op_c.reset(new VecOp<F>(vec_c.as_op()));
        auto* a = new StackOp<F>();
        a->ncol = vec_c.size().first;
        a->mats.push_back(mat_g.as_op()); a->mats.push_back(mat_a.as_op());
        a->signs = {F(1), F(1)};
        this->op_a.reset(a);
        auto* b = new StackOp<F>();
        b->ncol = 1;
        b->mats.push_back(vec_h.as_op()); b->mats.push_back(vec_b.as_op());
        b->signs = {F(1), F(1)};
        this->op_b.reset(b);
        auto* cone = new SeqCone<F>();
        cone->blocks.push_back({std::make_shared<ConeRPos<F>>(), m});
        cone->blocks.push_back({std::make_shared<ConeZero<F>>(), p});
        this->cone.reset(cone);
        this->w_solver.assign(Solver<F>::query_worklen(this->op_a->size()), F(0));
    }
};

// ---- QP / QCQP shared OpC -----------------------------------------------------------------------------------
template <typename F> struct QPOpC : Operator<F> {               // qp.rs:10-61, qcqp.rs:10-61
    using Sl = Slice<F>;
    using L = B200<F>;
    size_t n;
    explicit QPOpC(size_t n_) : n(n_) {}
    std::pair<size_t, size_t> size() const override { return {n + 1, 1}; }
    void op(F alpha, const Sl& x, F beta, Sl& y) const override {
        Sl y_n = y.sub(0, n), y_t = y.sub(n, 1);
        L::scale(beta, y_n);
        L::scale(beta, y_t);
        L::add(alpha, x, y_t);
    }
    void trans_op(F alpha, const Sl& x, F beta, Sl& y) const override {
        Sl x_t = x.sub(n, 1);
        L::scale(beta, y);
        L::add(alpha, x_t, y);
    }
    void absadd_cols(Sl& tau) const override { tau.set(0, tau.get(0) + F(1)); }
    void absadd_rows(Sl& sigma) const override { sigma.set(n, sigma.get(n) + F(1)); }
};

template <typename F> struct QPOpA : Operator<F> {               // qp.rs:65-170
    using Sl = Slice<F>;
    using L = B200<F>;
    std::unique_ptr<MatOp<F>> sym_p_sqrt, vec_q, mat_g, mat_a;
    void dim(size_t& n, size_t& m, size_t& p) const { n = sym_p_sqrt->size().first; m = mat_g->size().first; p = mat_a->size().first; }
    std::pair<size_t, size_t> size() const override {
        size_t n, m, p; dim(n, m, p);
        return {(2 + n) + m + p, n + 1};
    }
    void op(F alpha, const Sl& x, F beta, Sl& y) const override {
        size_t n, m, p; dim(n, m, p);
        Sl x_n = x.sub(0, n), x_t = x.sub(n, 1);
        Sl y_r = y.sub(0, 1), y_s = y.sub(1, 1), y_n = y.sub(2, n), y_m = y.sub(2 + n, m), y_p = y.sub(2 + n + m, p);
        L::scale(beta, y_r);
        vec_q->trans_op(alpha, x_n, beta, y_s);
        L::add(-alpha, x_t, y_s);
        sym_p_sqrt->op(-alpha, x_n, beta, y_n);
        mat_g->op(alpha, x_n, beta, y_m);
        mat_a->op(alpha, x_n, beta, y_p);
    }
    void trans_op(F alpha, const Sl& x, F beta, Sl& y) const override {
        size_t n, m, p; dim(n, m, p);
        Sl x_s = x.sub(1, 1), x_n = x.sub(2, n), x_m = x.sub(2 + n, m), x_p = x.sub(2 + n + m, p);
        Sl y_n = y.sub(0, n), y_t = y.sub(n, 1);
        vec_q->op(alpha, x_s, beta, y_n);
        sym_p_sqrt->op(-alpha, x_n, F(1), y_n);
        mat_g->trans_op(alpha, x_m, F(1), y_n);
        mat_a->trans_op(alpha, x_p, F(1), y_n);
        L::scale(beta, y_t);
        L::add(-alpha, x_s, y_t);
    }
    void absadd_cols(Sl& tau) const override {
        size_t n, m, p; dim(n, m, p);
        Sl tau_n = tau.sub(0, n), tau_t = tau.sub(n, 1);
        vec_q->absadd_rows(tau_n);
        sym_p_sqrt->absadd_cols(tau_n);
        mat_g->absadd_cols(tau_n);
        mat_a->absadd_cols(tau_n);
        tau_t.set(0, tau_t.get(0) + F(1));
    }
    void absadd_rows(Sl& sigma) const override {
        size_t n, m, p; dim(n, m, p);
        Sl s_s = sigma.sub(1, 1), s_n = sigma.sub(2, n), s_m = sigma.sub(2 + n, m), s_p = sigma.sub(2 + n + m, p);
        vec_q->absadd_cols(s_s);
        s_s.set(0, s_s.get(0) + F(1));
        sym_p_sqrt->absadd_rows(s_n);
        mat_g->absadd_rows(s_m);
        mat_a->absadd_rows(s_p);
    }
};

template <typename F> struct QPOpB : Operator<F> {               // qp.rs:174-258
    using Sl = Slice<F>;
    using L = B200<F>;
    size_t n;
    std::unique_ptr<MatOp<F>> vec_h, vec_b;
    std::pair<size_t, size_t> size() const override { return {(2 + n) + vec_h->size().first + vec_b->size().first, 1}; }
    void op(F alpha, const Sl& x, F beta, Sl& y) const override {
        const size_t m = vec_h->size().first, p = vec_b->size().first;
        Sl y_r = y.sub(0, 1), y_sn = y.sub(1, 1 + n), y_m = y.sub(2 + n, m), y_p = y.sub(2 + n + m, p);
        L::scale(beta, y_r);
        L::add(alpha, x, y_r);
        L::scale(beta, y_sn);
        vec_h->op(alpha, x, beta, y_m);
        vec_b->op(alpha, x, beta, y_p);
    }
    void trans_op(F alpha, const Sl& x, F beta, Sl& y) const override {
        const size_t m = vec_h->size().first, p = vec_b->size().first;
        Sl x_r = x.sub(0, 1), x_m = x.sub(2 + n, m), x_p = x.sub(2 + n + m, p);
        vec_h->trans_op(alpha, x_m, beta, y);
        vec_b->trans_op(alpha, x_p, F(1), y);
        L::add(alpha, x_r, y);
    }
    void absadd_cols(Sl& tau) const override {
        tau.set(0, tau.get(0) + F(1));
        vec_h->absadd_cols(tau);
        vec_b->absadd_cols(tau);
    }
    void absadd_rows(Sl& sigma) const override {
        const size_t m = vec_h->size().first, p = vec_b->size().first;
        Sl s_r = sigma.sub(0, 1), s_m = sigma.sub(2 + n, m), s_p = sigma.sub(2 + n + m, p);
        s_r.set(0, s_r.get(0) + F(1));
        vec_h->absadd_rows(s_m);
        vec_b->absadd_rows(s_p);
    }
};

template <typename F> struct ProbQP : Problem<F> {               // qp.rs:300-437
    MatBuild<F> vec_q, mat_g, vec_h, mat_a, vec_b, sym_p_sqrt;
    // p_is_sqrt: sym_p already holds P^(1/2) (bench shortcut for config C2's diagonal P; qp.rs:386 is skipped)
    ProbQP(MatBuild<F> sym_p, MatBuild<F> q, MatBuild<F> g, MatBuild<F> h, MatBuild<F> a, MatBuild<F> b, F eps_zero, bool p_is_sqrt = false)
        : vec_q(std::move(q)), mat_g(std::move(g)), vec_h(std::move(h)), mat_a(std::move(a)), vec_b(std::move(b)), sym_p_sqrt(std::move(sym_p)) {
        if (!sym_p_sqrt.is_sympack()) throw BackendError("ProbQP: sym_p must be SymPack");
        if (!p_is_sqrt) sym_p_sqrt.set_sqrt(eps_zero);           // qp.rs:386
    }
    void problem() {
        const size_t n = vec_q.size().first, m = vec_h.size().first, p = vec_b.size().first;
        this->op_c.reset(new QPOpC<F>(n));
        auto* a = new QPOpA<F>();
        a->sym_p_sqrt = sym_p_sqrt.as_op(); a->vec_q = vec_q.as_op(); a->mat_g = mat_g.as_op(); a->mat_a = mat_a.as_op();
        this->op_a.reset(a);
        auto* b = new QPOpB<F>();
        b->n = n; b->vec_h = vec_h.as_op(); b->vec_b = vec_b.as_op();
        this->op_b.reset(b);
        auto* cone = new SeqCone<F>();
        cone->blocks.push_back({std::make_shared<ConeRotSOC<F>>(), 2 + n});
        cone->blocks.push_back({std::make_shared<ConeRPos<F>>(), m});
        cone->blocks.push_back({std::make_shared<ConeZero<F>>(), p});
        this->cone.reset(cone);
        this->w_solver.assign(Solver<F>::query_worklen(this->op_a->size()), F(0));
    }
};

// ---- QCQP ---------------------------------------------------------------------------------------------------
template <typename F> struct QCQPOpA : Operator<F> {             // qcqp.rs:65-193
    using Sl = Slice<F>;
    using L = B200<F>;
    std::vector<std::unique_ptr<MatOp<F>>> syms_p_sqrt, vecs_q;
    std::unique_ptr<MatOp<F>> mat_a;
    void dim(size_t& n, size_t& m1, size_t& p) const { p = mat_a->size().first; n = mat_a->size().second; m1 = syms_p_sqrt.size(); }
    std::pair<size_t, size_t> size() const override {
        size_t n, m1, p; dim(n, m1, p);
        return {m1 * (2 + n) + p, n + 1};
    }
    void op(F alpha, const Sl& x, F beta, Sl& y) const override {
        size_t n, m1, p; dim(n, m1, p);
        Sl x_n = x.sub(0, n), x_t = x.sub(n, 1);
        for (size_t i = 0; i < m1; ++i) {
            const size_t o = i * (2 + n);
            Sl y_r = y.sub(o, 1), y_s = y.sub(o + 1, 1), y_n = y.sub(o + 2, n);
            L::scale(beta, y_r);
            vecs_q[i]->trans_op(alpha, x_n, beta, y_s);
            if (i == 0) L::add(-alpha, x_t, y_s);
            syms_p_sqrt[i]->op(-alpha, x_n, beta, y_n);
        }
        Sl y_p = y.sub(m1 * (2 + n), p);
        mat_a->op(alpha, x_n, beta, y_p);
    }
    void trans_op(F alpha, const Sl& x, F beta, Sl& y) const override {
        size_t n, m1, p; dim(n, m1, p);
        Sl y_n = y.sub(0, n), y_t = y.sub(n, 1);
        L::scale(beta, y_n);
        L::scale(beta, y_t);
        for (size_t i = 0; i < m1; ++i) {
            const size_t o = i * (2 + n);
            Sl x_s = x.sub(o + 1, 1), x_n = x.sub(o + 2, n);
            vecs_q[i]->op(alpha, x_s, F(1), y_n);
            syms_p_sqrt[i]->op(-alpha, x_n, F(1), y_n);
            if (i == 0) L::add(-alpha, x_s, y_t);
        }
        Sl x_p = x.sub(m1 * (2 + n), p);
        mat_a->trans_op(alpha, x_p, F(1), y_n);
    }
    void absadd_cols(Sl& tau) const override {
        size_t n, m1, p; dim(n, m1, p);
        Sl tau_n = tau.sub(0, n), tau_t = tau.sub(n, 1);
        for (auto& q : vecs_q) q->absadd_rows(tau_n);
        for (auto& s : syms_p_sqrt) s->absadd_cols(tau_n);
        mat_a->absadd_cols(tau_n);
        tau_t.set(0, tau_t.get(0) + F(1));
    }
    void absadd_rows(Sl& sigma) const override {
        size_t n, m1, p; dim(n, m1, p);
        for (size_t i = 0; i < m1; ++i) {
            const size_t o = i * (2 + n);
            Sl s_s = sigma.sub(o + 1, 1), s_n = sigma.sub(o + 2, n);
            vecs_q[i]->absadd_cols(s_s);
            if (i == 0) s_s.set(0, s_s.get(0) + F(1));
            syms_p_sqrt[i]->absadd_rows(s_n);
        }
        Sl s_p = sigma.sub(m1 * (2 + n), p);
        mat_a->absadd_rows(s_p);
    }
};

template <typename F> struct QCQPOpB : Operator<F> {             // qcqp.rs:197-299
    using Sl = Slice<F>;
    using L = B200<F>;
    size_t n;
    std::vector<F> scls_r;
    F abssum_scls_r = 0;
    std::unique_ptr<MatOp<F>> vec_b;
    std::pair<size_t, size_t> size() const override { return {scls_r.size() * (2 + n) + vec_b->size().first, 1}; }
    void op(F alpha, const Sl& x, F beta, Sl& y) const override {
        const size_t m1 = scls_r.size(), p = vec_b->size().first;
        for (size_t i = 0; i < m1; ++i) {
            const size_t o = i * (2 + n);
            Sl y_r = y.sub(o, 1), y_s = y.sub(o + 1, 1), y_n = y.sub(o + 2, n);
            L::scale(beta, y_r); L::add(alpha, x, y_r);
            L::scale(beta, y_s); L::add(-alpha * scls_r[i], x, y_s);
            L::scale(beta, y_n);
        }
        Sl y_p = y.sub(m1 * (2 + n), p);
        vec_b->op(alpha, x, beta, y_p);
    }
    void trans_op(F alpha, const Sl& x, F beta, Sl& y) const override {
        const size_t m1 = scls_r.size(), p = vec_b->size().first;
        L::scale(beta, y);
        for (size_t i = 0; i < m1; ++i) {
            const size_t o = i * (2 + n);
            Sl x_r = x.sub(o, 1), x_s = x.sub(o + 1, 1);
            L::add(alpha, x_r, y);
            L::add(-alpha * scls_r[i], x_s, y);
        }
        Sl x_p = x.sub(m1 * (2 + n), p);
        vec_b->trans_op(alpha, x_p, F(1), y);
    }
    void absadd_cols(Sl& tau) const override {
        tau.set(0, tau.get(0) + F(scls_r.size()) + abssum_scls_r);
        vec_b->absadd_cols(tau);
    }
    void absadd_rows(Sl& sigma) const override {
        const size_t m1 = scls_r.size(), p = vec_b->size().first;
        for (size_t i = 0; i < m1; ++i) {
            const size_t o = i * (2 + n);
            sigma.set(o, sigma.get(o) + F(1));
            sigma.set(o + 1, sigma.get(o + 1) + std::fabs(scls_r[i]));
        }
        Sl s_p = sigma.sub(m1 * (2 + n), p);
        vec_b->absadd_rows(s_p);
    }
};

template <typename F> struct ProbQCQP : Problem<F> {             // qcqp.rs:354-481
    std::vector<MatBuild<F>> vecs_q, syms_p_sqrt;
    std::vector<F> scls_r;
    MatBuild<F> mat_a, vec_b;
    ProbQCQP(std::vector<MatBuild<F>> syms_p, std::vector<MatBuild<F>> q, std::vector<F> r, MatBuild<F> a, MatBuild<F> b, F eps_zero)
        : vecs_q(std::move(q)), syms_p_sqrt(std::move(syms_p)), scls_r(std::move(r)), mat_a(std::move(a)), vec_b(std::move(b)) {
        for (auto& s : syms_p_sqrt) s.set_sqrt(eps_zero);        // qcqp.rs:445-448
    }
    void problem() {
        const size_t p = mat_a.size().first, n = mat_a.size().second, m1 = syms_p_sqrt.size();
        this->op_c.reset(new QPOpC<F>(n));
        auto* a = new QCQPOpA<F>();
        for (auto& s : syms_p_sqrt) a->syms_p_sqrt.push_back(s.as_op());
        for (auto& q : vecs_q) a->vecs_q.push_back(q.as_op());
        a->mat_a = mat_a.as_op();
        this->op_a.reset(a);
        auto* b = new QCQPOpB<F>();
        b->n = n; b->scls_r = scls_r; b->vec_b = vec_b.as_op();
        {
            Slice<F> r = Slice<F>::new_ref(scls_r.data(), scls_r.size());
            b->abssum_scls_r = B200<F>::abssum(r, 1);            // qcqp.rs:463
        }
        this->op_b.reset(b);
        auto* cone = new SeqCone<F>();
        auto rot = std::make_shared<ConeRotSOC<F>>();
        for (size_t i = 0; i < m1; ++i) cone->blocks.push_back({rot, 2 + n});
        cone->blocks.push_back({std::make_shared<ConeZero<F>>(), p});
        this->cone.reset(cone);
        this->w_solver.assign(Solver<F>::query_worklen(this->op_a->size()), F(0));
    }
};

// ---- SOCP ---------------------------------------------------------------------------------------------------
template <typename F> struct SOCPOpA : Operator<F> {             // socp.rs:47-165
    using Sl = Slice<F>;
    using L = B200<F>;
    std::vector<std::unique_ptr<MatOp<F>>> mats_g, vecs_c;
    std::unique_ptr<MatOp<F>> mat_a;
    std::pair<size_t, size_t> size() const override {
        size_t s = 0;
        for (auto& g : mats_g) s += 1 + g->size().first;
        return {s + mat_a->size().first, mat_a->size().second};
    }
    void op(F alpha, const Sl& x, F beta, Sl& y) const override {
        size_t done = 0;
        for (size_t i = 0; i < mats_g.size(); ++i) {
            const size_t ni = mats_g[i]->size().first;
            Sl y_1 = y.sub(done, 1), y_ni = y.sub(done + 1, ni);
            done += 1 + ni;
            vecs_c[i]->trans_op(-alpha, x, beta, y_1);
            mats_g[i]->op(-alpha, x, beta, y_ni);
        }
        Sl y_p = y.sub(done, mat_a->size().first);
        mat_a->op(alpha, x, beta, y_p);
    }
    void trans_op(F alpha, const Sl& x, F beta, Sl& y) const override {
        L::scale(beta, y);
        size_t done = 0;
        for (size_t i = 0; i < mats_g.size(); ++i) {
            const size_t ni = mats_g[i]->size().first;
            Sl x_1 = x.sub(done, 1), x_ni = x.sub(done + 1, ni);
            done += 1 + ni;
            vecs_c[i]->op(-alpha, x_1, F(1), y);
            mats_g[i]->trans_op(-alpha, x_ni, F(1), y);
        }
        Sl x_p = x.sub(done, mat_a->size().first);
        mat_a->trans_op(alpha, x_p, F(1), y);
    }
    void absadd_cols(Sl& tau) const override {
        for (auto& c : vecs_c) c->absadd_rows(tau);
        for (auto& g : mats_g) g->absadd_cols(tau);
        mat_a->absadd_cols(tau);
    }
    void absadd_rows(Sl& sigma) const override {
        size_t done = 0;
        for (size_t i = 0; i < mats_g.size(); ++i) {
            const size_t ni = mats_g[i]->size().first;
            Sl s_1 = sigma.sub(done, 1), s_ni = sigma.sub(done + 1, ni);
            done += 1 + ni;
            vecs_c[i]->absadd_cols(s_1);
            mats_g[i]->absadd_rows(s_ni);
        }
        Sl s_p = sigma.sub(done, mat_a->size().first);
        mat_a->absadd_rows(s_p);
    }
};

template <typename F> struct SOCPOpB : Operator<F> {             // socp.rs:169-282
    using Sl = Slice<F>;
    using L = B200<F>;
    std::vector<std::unique_ptr<MatOp<F>>> vecs_h;
    std::vector<F> scls_d;
    F abssum_scls_d = 0;
    std::unique_ptr<MatOp<F>> vec_b;
    std::pair<size_t, size_t> size() const override {
        size_t s = 0;
        for (auto& h : vecs_h) s += 1 + h->size().first;
        return {s + vec_b->size().first, 1};
    }
    void op(F alpha, const Sl& x, F beta, Sl& y) const override {
        size_t done = 0;
        for (size_t i = 0; i < vecs_h.size(); ++i) {
            const size_t ni = vecs_h[i]->size().first;
            Sl y_1 = y.sub(done, 1), y_ni = y.sub(done + 1, ni);
            done += 1 + ni;
            L::scale(beta, y_1);
            L::add(alpha * scls_d[i], x, y_1);
            vecs_h[i]->op(alpha, x, beta, y_ni);
        }
        Sl y_p = y.sub(done, vec_b->size().first);
        vec_b->op(alpha, x, beta, y_p);
    }
    void trans_op(F alpha, const Sl& x, F beta, Sl& y) const override {
        L::scale(beta, y);
        size_t done = 0;
        for (size_t i = 0; i < vecs_h.size(); ++i) {
            const size_t ni = vecs_h[i]->size().first;
            Sl x_1 = x.sub(done, 1), x_ni = x.sub(done + 1, ni);
            done += 1 + ni;
            L::add(alpha * scls_d[i], x_1, y);
            vecs_h[i]->trans_op(alpha, x_ni, F(1), y);
        }
        Sl x_p = x.sub(done, vec_b->size().first);
        vec_b->trans_op(alpha, x_p, F(1), y);
    }
    void absadd_cols(Sl& tau) const override {
        tau.set(0, tau.get(0) + abssum_scls_d);
        for (auto& h : vecs_h) h->absadd_cols(tau);
        vec_b->absadd_cols(tau);
    }
    void absadd_rows(Sl& sigma) const override {
        size_t done = 0;
        for (size_t i = 0; i < vecs_h.size(); ++i) {
            const size_t ni = vecs_h[i]->size().first;
            Sl s_1 = sigma.sub(done, 1), s_ni = sigma.sub(done + 1, ni);
            done += 1 + ni;
            s_1.set(0, s_1.get(0) + scls_d[i]);                  // socp.rs:272 (no abs in the reference)
            vecs_h[i]->absadd_rows(s_ni);
        }
        Sl s_p = sigma.sub(done, vec_b->size().first);
        vec_b->absadd_rows(s_p);
    }
};

template <typename F> struct ProbSOCP : Problem<F> {             // socp.rs:337-472
    MatBuild<F> vec_f;
    std::vector<MatBuild<F>> mats_g, vecs_h, vecs_c;
    std::vector<F> scls_d;
    MatBuild<F> mat_a, vec_b;
    ProbSOCP(MatBuild<F> f, std::vector<MatBuild<F>> g, std::vector<MatBuild<F>> h, std::vector<MatBuild<F>> c, std::vector<F> d,
             MatBuild<F> a, MatBuild<F> b)
        : vec_f(std::move(f)), mats_g(std::move(g)), vecs_h(std::move(h)), vecs_c(std::move(c)), scls_d(std::move(d)),
          mat_a(std::move(a)), vec_b(std::move(b)) {
        if (vecs_h.size() != mats_g.size() || vecs_c.size() != mats_g.size() || scls_d.size() != mats_g.size())
            throw BackendError("ProbSOCP: block count mismatch");
    }
    void problem() {
        const size_t p = vec_b.size().first;
        this->op_c.reset(new VecOp<F>(vec_f.as_op()));
        auto* a = new SOCPOpA<F>();
        for (auto& g : mats_g) a->mats_g.push_back(g.as_op());
        for (auto& c : vecs_c) a->vecs_c.push_back(c.as_op());
        a->mat_a = mat_a.as_op();
        this->op_a.reset(a);
        auto* b = new SOCPOpB<F>();
        for (auto& h : vecs_h) b->vecs_h.push_back(h.as_op());
        b->scls_d = scls_d;
        {
            Slice<F> d = Slice<F>::new_ref(scls_d.data(), scls_d.size());
            b->abssum_scls_d = B200<F>::abssum(d, 1);            // socp.rs:457
        }
        b->vec_b = vec_b.as_op();
        this->op_b.reset(b);
        auto* cone = new SeqCone<F>();
        auto soc = std::make_shared<ConeSOC<F>>();
        for (auto& g : mats_g) cone->blocks.push_back({soc, 1 + g.size().first});
        cone->blocks.push_back({std::make_shared<ConeZero<F>>(), p});
        this->cone.reset(cone);
        this->w_solver.assign(Solver<F>::query_worklen(this->op_a->size()), F(0));
    }
};

// ---- SDP ----------------------------------------------------------------------------------------------------
template <typename F> struct ProbSDP : Problem<F> {              // sdp.rs:224-365
    MatBuild<F> vec_c, mat_a, vec_b, symmat_f, symvec_f_n;
    F eps_zero;
    ProbSDP(MatBuild<F> c, std::vector<MatBuild<F>> syms_f, MatBuild<F> a, MatBuild<F> b, F eps)
        : vec_c(std::move(c)), mat_a(std::move(a)), vec_b(std::move(b)), eps_zero(eps) {
        const size_t n = vec_c.size().first;
        if (syms_f.size() != n + 1) throw BackendError("ProbSDP: need n+1 symmetric matrices");
        const F fsqrt2 = std::sqrt(F(2));
        for (auto& s : syms_f) {
            s.set_scale_nondiag(fsqrt2);                         // sdp.rs:305-308
            s.set_reshape_colvec();
        }
        symvec_f_n = std::move(syms_f.back());
        syms_f.pop_back();
        const size_t sk = symvec_f_n.size().first;
        symmat_f = MatBuild<F>(MatType::general(sk, n));
        for (size_t c2 = 0; c2 < n; ++c2)
            for (size_t r = 0; r < sk; ++r) symmat_f.at(r, c2) = syms_f[c2].array()[r];
    }
    void problem() {
        const size_t p = vec_b.size().first, sk = symvec_f_n.size().first;
        this->op_c.reset(new VecOp<F>(vec_c.as_op()));
        auto* a = new StackOp<F>();
        a->ncol = vec_c.size().first;
        a->mats.push_back(symmat_f.as_op()); a->mats.push_back(mat_a.as_op());
        a->signs = {F(1), F(1)};
        this->op_a.reset(a);
        auto* b = new StackOp<F>();
        b->ncol = 1;
        b->mats.push_back(symvec_f_n.as_op()); b->mats.push_back(vec_b.as_op());
        b->signs = {F(-1), F(1)};
        this->op_b.reset(b);
        this->w_cone.assign(ConePSD<F>::query_worklen(sk), F(0));
        auto* cone = new SeqCone<F>();
        cone->blocks.push_back({std::make_shared<ConePSD<F>>(this->w_cone.data(), this->w_cone.size(), eps_zero), sk});
        cone->blocks.push_back({std::make_shared<ConeZero<F>>(), p});
        this->cone.reset(cone);
        this->w_solver.assign(Solver<F>::query_worklen(this->op_a->size()), F(0));
    }
};

}  // namespace totsu_b200
