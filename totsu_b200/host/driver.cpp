// libtotsu_b200_host.so: flat C entry points over the C++ host layer so that the Python tests and bench.py can
// build the reference's problems (ProbLP/QP/QCQP/SOCP/SDP or raw Operator/Cone tuples), run the solver loop
// on the B200 backend and read iterates back.  This is harness plumbing, not part of the backend ABI
// (include/totsu_b200.h is); everything numerical happens in libtotsu_b200.so.
#include <cstring>
#include <memory>
#include <string>
#include "fused.hpp"
#include "problem.hpp"

using namespace totsu_b200;

extern "C" {

struct tbh_param {
    int64_t max_iter;      // < 0: None
    double eps_acc, eps_inf, eps_zero;
    uint64_t log_period;
    int32_t device_precond;
    int32_t reserved;
};

struct tbh_iter {
    uint64_t i;
    int32_t conv_branch;
    int32_t logged;
    double val_tau, c0, c1, c2;
};

}  // extern "C"

namespace {

std::string g_err;

struct SessionBase {
    virtual ~SessionBase() = default;
    virtual int begin(const tbh_param& p) = 0;
    virtual int step(uint64_t k, tbh_iter* last, int* done) = 0;
    virtual int run(tbh_iter* trace, size_t cap, size_t* n_trace, int only_logged, tbh_iter* last) = 0;
    virtual void end() = 0;
    virtual void xy(void* x_hat, void* y_hat) = 0;
    virtual void solution(void* x, void* y) = 0;
    virtual void dims(size_t* m, size_t* n) = 0;
    virtual void norms(double* nb, double* nc) = 0;
};

template <typename F> struct Session : SessionBase {
    // storage for raw/dense sessions (declared first: outlives the operators wrapping it)
    std::vector<F> c_host, b_host;
    Slice<F> c_dev, b_dev;                  // fused path: c and b wrapped once, served by DenseOp (n x 1, m x 1)
    std::unique_ptr<Problem<F>> prob;       // owns operators, cone, work
    Solver<F> solver;
    std::vector<tbh_iter>* trace_sink = nullptr;
    int only_logged = 0;
    size_t m = 0, n = 0;
    bool begun = false;

    void set_dims() {
        auto sz = prob->op_a->size();
        m = sz.first; n = sz.second;
    }
    int begin(const tbh_param& p) override {
        solver.par.max_iter = p.max_iter >= 0 ? std::optional<size_t>((size_t)p.max_iter) : std::nullopt;
        solver.par.eps_acc = (F)p.eps_acc; solver.par.eps_inf = (F)p.eps_inf; solver.par.eps_zero = (F)p.eps_zero;
        solver.par.log_period = (size_t)p.log_period;
        solver.device_precond = p.device_precond != 0;
        solver.trace = [this](const IterInfo<F>& it, bool logged) {
            last_ = tbh_iter{(uint64_t)it.i, it.conv_branch ? 1 : 0, logged ? 1 : 0, (double)it.val_tau, (double)it.c0, (double)it.c1, (double)it.c2};
            if (trace_sink && (logged || !only_logged)) trace_sink->push_back(last_);
        };
        SolverError e = solver.begin(*prob->op_c, *prob->op_a, *prob->op_b, *prob->cone, prob->w_solver.data(), prob->w_solver.size());
        begun = e == SolverError::None;
        return (int)e;
    }
    int step(uint64_t k, tbh_iter* last, int* done) override {
        SolverError e = SolverError::None;
        bool d = false;
        for (uint64_t j = 0; j < k && !d; ++j) e = solver.step(d);
        if (last) *last = last_;
        if (done) *done = d ? 1 : 0;
        return (int)e;
    }
    int run(tbh_iter* trace, size_t cap, size_t* n_trace, int only_logged_, tbh_iter* last) override {
        std::vector<tbh_iter> sink;
        trace_sink = trace ? &sink : nullptr;
        only_logged = only_logged_;
        SolverError e = SolverError::None;
        bool d = false;
        while (!d) e = solver.step(d);
        trace_sink = nullptr;
        if (trace) {
            size_t cnt = std::min(cap, sink.size());
            // keep the tail if the buffer is too small
            std::memcpy(trace, sink.data() + (sink.size() - cnt), cnt * sizeof(tbh_iter));
            if (n_trace) *n_trace = cnt;
        }
        if (last) *last = last_;
        return (int)e;
    }
    void end() override {
        if (begun) solver.end();
        begun = false;
    }
    void xy(void* x_hat, void* y_hat) override {
        TBH_CALL(tb_download(solver.x().view(), x_hat));
        TBH_CALL(tb_download(solver.y().view(), y_hat));
    }
    void solution(void* x, void* y) override {
        std::memcpy(x, prob->w_solver.data(), n * sizeof(F));
        std::memcpy(y, prob->w_solver.data() + n, m * sizeof(F));
    }
    void dims(size_t* m_, size_t* n_) override { *m_ = m; *n_ = n; }
    void norms(double* nb, double* nc) override { *nb = (double)solver.norm_b(); *nc = (double)solver.norm_c(); }
    tbh_iter last_{};
};

template <typename F> MatBuild<F> mb_general(size_t r, size_t c, const void* data) {
    if (r * c == 0) return MatBuild<F>(MatType::general(r, c));
    return MatBuild<F>(MatType::general(r, c), reinterpret_cast<const F*>(data));
}
template <typename F> MatBuild<F> mb_sym(size_t n, const void* data) {
    if (n == 0) return MatBuild<F>(MatType::sympack(0));
    return MatBuild<F>(MatType::sympack(n), reinterpret_cast<const F*>(data));
}

template <typename Fn> void* guarded_new(Fn fn) {
    try {
        return fn();
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
template <typename Fn> int guarded(Fn fn) {
    try {
        return fn();
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

template <typename F>
SessionBase* make_lp(size_t n, size_t m, size_t p, const void* c, const void* g, const void* h, const void* a, const void* b) {
    auto* pr = new ProbLP<F>(mb_general<F>(n, 1, c), mb_general<F>(m, n, g), mb_general<F>(m, 1, h), mb_general<F>(p, n, a), mb_general<F>(p, 1, b));
    auto* s = new Session<F>();
    s->prob.reset(pr);
    pr->problem();
    s->set_dims();
    return s;
}
template <typename F>
SessionBase* make_qp(size_t n, size_t m, size_t p, const void* sym_p, const void* q, const void* g, const void* h, const void* a, const void* b, double eps_zero,
                     int p_is_sqrt) {
    auto* pr = new ProbQP<F>(mb_sym<F>(n, sym_p), mb_general<F>(n, 1, q), mb_general<F>(m, n, g), mb_general<F>(m, 1, h),
                             mb_general<F>(p, n, a), mb_general<F>(p, 1, b), (F)eps_zero, p_is_sqrt != 0);
    auto* s = new Session<F>();
    s->prob.reset(pr);
    pr->problem();
    s->set_dims();
    return s;
}
template <typename F>
SessionBase* make_qcqp(size_t n, size_t m1, size_t p, const void* syms_p, const void* vecs_q, const void* scls_r, const void* a, const void* b, double eps_zero) {
    const size_t sn = n * (n + 1) / 2;
    std::vector<MatBuild<F>> sp, vq;
    std::vector<F> r(m1);
    for (size_t i = 0; i < m1; ++i) {
        sp.push_back(mb_sym<F>(n, reinterpret_cast<const F*>(syms_p) + i * sn));
        vq.push_back(mb_general<F>(n, 1, reinterpret_cast<const F*>(vecs_q) + i * n));
        r[i] = reinterpret_cast<const F*>(scls_r)[i];
    }
    auto* pr = new ProbQCQP<F>(std::move(sp), std::move(vq), std::move(r), mb_general<F>(p, n, a), mb_general<F>(p, 1, b), (F)eps_zero);
    auto* s = new Session<F>();
    s->prob.reset(pr);
    pr->problem();
    s->set_dims();
    return s;
}
// blocks: ni[i] rows each; mats_g concatenated (each ni x n column-major), vecs_h concatenated, vecs_c n each
template <typename F>
SessionBase* make_socp(size_t n, size_t nblk, const uint64_t* ni, size_t p, const void* f, const void* mats_g, const void* vecs_h,
                       const void* vecs_c, const void* scls_d, const void* a, const void* b) {
    std::vector<MatBuild<F>> g, h, c;
    std::vector<F> d(nblk);
    size_t og = 0, oh = 0;
    for (size_t i = 0; i < nblk; ++i) {
        g.push_back(mb_general<F>((size_t)ni[i], n, reinterpret_cast<const F*>(mats_g) + og));
        h.push_back(mb_general<F>((size_t)ni[i], 1, reinterpret_cast<const F*>(vecs_h) + oh));
        c.push_back(mb_general<F>(n, 1, reinterpret_cast<const F*>(vecs_c) + i * n));
        d[i] = reinterpret_cast<const F*>(scls_d)[i];
        og += (size_t)ni[i] * n;
        oh += (size_t)ni[i];
    }
    auto* pr = new ProbSOCP<F>(mb_general<F>(n, 1, f), std::move(g), std::move(h), std::move(c), std::move(d), mb_general<F>(p, n, a), mb_general<F>(p, 1, b));
    auto* s = new Session<F>();
    s->prob.reset(pr);
    pr->problem();
    s->set_dims();
    return s;
}
template <typename F>
SessionBase* make_sdp(size_t n, size_t k, size_t p, const void* c, const void* syms_f, const void* a, const void* b, double eps_zero) {
    const size_t sk = k * (k + 1) / 2;
    std::vector<MatBuild<F>> sf;
    for (size_t i = 0; i <= n; ++i) sf.push_back(mb_sym<F>(k, reinterpret_cast<const F*>(syms_f) + i * sk));
    auto* pr = new ProbSDP<F>(mb_general<F>(n, 1, c), std::move(sf), mb_general<F>(p, n, a), mb_general<F>(p, 1, b), (F)eps_zero);
    auto* s = new Session<F>();
    s->prob.reset(pr);
    pr->problem();
    s->set_dims();
    return s;
}

// raw (op_c, op_a, op_b, cone) tuple over one dense A, like totsu_core/tests/solver.rs builds by hand.
// fused_op: DenseOp instead of MatOp; fused_cone: ProductCone instead of a sequence of stock cones.
template <typename F>
SessionBase* make_dense(size_t m_local, size_t n, tb_view a_view, size_t row_offset, size_t m_total, const void* c, const void* b,
                        const tb_cone_block* blocks, size_t nblk, int fused_op, int fused_cone, double eps_zero) {
    auto* s = new Session<F>();
    auto* pr = new Problem<F>();
    s->prob.reset(pr);
    if (m_total == 0) m_total = m_local;
    s->c_host.assign(reinterpret_cast<const F*>(c), reinterpret_cast<const F*>(c) + n);
    s->b_host.assign(reinterpret_cast<const F*>(b), reinterpret_cast<const F*>(b) + m_total);
    if (fused_op) {
        // c and b as device-resident n x 1 / m x 1 dense operators: their absadd_* run as kernels instead of one
        // blocking strided abssum per element through the stock MatOp (matop.rs:105-116)
        s->c_dev = Slice<F>::new_ref(s->c_host.data(), n);
        s->b_dev = Slice<F>::new_ref(s->b_host.data(), m_total);
        pr->op_c.reset(new DenseOp<F>(s->c_dev.view(), n, 1));
        pr->op_b.reset(new DenseOp<F>(s->b_dev.view(), m_total, 1));
        pr->op_a.reset(new DenseOp<F>(a_view, m_local, n, row_offset, m_total));
    } else {
        pr->op_c.reset(new VecOp<F>(std::unique_ptr<MatOp<F>>(new MatOp<F>(MatType::general(n, 1), s->c_host.data(), n))));
        pr->op_b.reset(new VecOp<F>(std::unique_ptr<MatOp<F>>(new MatOp<F>(MatType::general(m_total, 1), s->b_host.data(), m_total))));
        if (m_local != m_total) throw BackendError("stock MatOp path cannot be row-sharded");
        pr->op_a.reset(new MatOp<F>(MatType::general(m_local, n), a_view));
    }
    std::vector<tb_cone_block> blk(blocks, blocks + nblk);
    if (fused_cone) {
        pr->cone.reset(new ProductCone<F>(blk, (F)eps_zero));
    } else {
        auto* cone = new SeqCone<F>();
        size_t max_psd = 0;
        for (auto& bk : blk)
            if (bk.type == TB_CONE_PSD) max_psd = std::max<size_t>(max_psd, ConePSD<F>::query_worklen((size_t)bk.len));
        pr->w_cone.assign(max_psd, F(0));
        for (auto& bk : blk) {
            std::shared_ptr<Cone<F>> cn;
            switch (bk.type) {
                case TB_CONE_ZERO: cn = std::make_shared<ConeZero<F>>(); break;
                case TB_CONE_RPOS: cn = std::make_shared<ConeRPos<F>>(); break;
                case TB_CONE_SOC: cn = std::make_shared<ConeSOC<F>>(); break;
                case TB_CONE_ROTSOC: cn = std::make_shared<ConeRotSOC<F>>(); break;
                case TB_CONE_PSD: cn = std::make_shared<ConePSD<F>>(pr->w_cone.data(), pr->w_cone.size(), (F)eps_zero); break;
                default: throw BackendError("unknown cone type");
            }
            cone->blocks.push_back({cn, (size_t)bk.len});
        }
        pr->cone.reset(cone);
    }
    pr->w_solver.assign(Solver<F>::query_worklen({m_total, n}), F(0));
    s->m = m_total; s->n = n;
    return s;
}

}  // namespace

#define DISPATCH(dtype, fn, ...) ((dtype) == TB_F32 ? fn<float>(__VA_ARGS__) : fn<double>(__VA_ARGS__))

extern "C" {

const char* tbh_last_error(void) { return g_err.c_str(); }

// 1: issue the Rust binding's call protocol (tb_view_of_host per operand, tb_buf_retain / tb_buf_release per split child;
// host/linalg.hpp "shim-protocol mode"); 0: carry (handle, offset, length) views.  Set before sessions are created.
void tbh_set_shim_protocol(int on) { shim_protocol() = on != 0; }
int tbh_get_shim_protocol(void) { return shim_protocol() ? 1 : 0; }
// 0: B200::map_eig always takes the general closure route (device eigendecomposition + host closure); 1 (default): the two
// closures the reference itself uses are recognised and served by tb_proj_psd / tb_sqrt_psd
void tbh_set_map_eig_fast_paths(int on) { map_eig_fast_paths() = on != 0; }

void* tbh_session_lp(int dtype, size_t n, size_t m, size_t p, const void* c, const void* g, const void* h, const void* a, const void* b) {
    return guarded_new([&] { return (void*)DISPATCH(dtype, make_lp, n, m, p, c, g, h, a, b); });
}
void* tbh_session_qp(int dtype, size_t n, size_t m, size_t p, const void* sym_p, const void* q, const void* g, const void* h, const void* a, const void* b, double eps_zero,
                     int p_is_sqrt) {
    return guarded_new([&] { return (void*)DISPATCH(dtype, make_qp, n, m, p, sym_p, q, g, h, a, b, eps_zero, p_is_sqrt); });
}
void* tbh_session_qcqp(int dtype, size_t n, size_t m1, size_t p, const void* syms_p, const void* vecs_q, const void* scls_r, const void* a, const void* b, double eps_zero) {
    return guarded_new([&] { return (void*)DISPATCH(dtype, make_qcqp, n, m1, p, syms_p, vecs_q, scls_r, a, b, eps_zero); });
}
void* tbh_session_socp(int dtype, size_t n, size_t nblk, const uint64_t* ni, size_t p, const void* f, const void* mats_g, const void* vecs_h,
                       const void* vecs_c, const void* scls_d, const void* a, const void* b) {
    return guarded_new([&] { return (void*)DISPATCH(dtype, make_socp, n, nblk, ni, p, f, mats_g, vecs_h, vecs_c, scls_d, a, b); });
}
void* tbh_session_sdp(int dtype, size_t n, size_t k, size_t p, const void* c, const void* syms_f, const void* a, const void* b, double eps_zero) {
    return guarded_new([&] { return (void*)DISPATCH(dtype, make_sdp, n, k, p, c, syms_f, a, b, eps_zero); });
}
void* tbh_session_dense(int dtype, size_t m_local, size_t n, tb_view a_view, size_t row_offset, size_t m_total, const void* c, const void* b,
                        const tb_cone_block* blocks, size_t nblk, int fused_op, int fused_cone, double eps_zero) {
    return guarded_new([&] { return (void*)DISPATCH(dtype, make_dense, m_local, n, a_view, row_offset, m_total, c, b, blocks, nblk, fused_op, fused_cone, eps_zero); });
}

int tbh_session_begin(void* s, const tbh_param* p) { return guarded([&] { return ((SessionBase*)s)->begin(*p); }); }
int tbh_session_step(void* s, uint64_t k, tbh_iter* last, int* done) { return guarded([&] { return ((SessionBase*)s)->step(k, last, done); }); }
int tbh_session_run(void* s, tbh_iter* trace, size_t cap, size_t* n_trace, int only_logged, tbh_iter* last) {
    return guarded([&] { return ((SessionBase*)s)->run(trace, cap, n_trace, only_logged, last); });
}
int tbh_session_end(void* s) { return guarded([&] { ((SessionBase*)s)->end(); return 0; }); }
int tbh_session_xy(void* s, void* x_hat, void* y_hat) { return guarded([&] { ((SessionBase*)s)->xy(x_hat, y_hat); return 0; }); }
int tbh_session_solution(void* s, void* x, void* y) { return guarded([&] { ((SessionBase*)s)->solution(x, y); return 0; }); }
int tbh_session_dims(void* s, size_t* m, size_t* n) { return guarded([&] { ((SessionBase*)s)->dims(m, n); return 0; }); }
int tbh_session_norms(void* s, double* nb, double* nc) { return guarded([&] { ((SessionBase*)s)->norms(nb, nc); return 0; }); }
void tbh_session_destroy(void* s) {
    try {
        SessionBase* b = (SessionBase*)s;
        b->end();
        delete b;
    } catch (const std::exception& e) {
        g_err = e.what();
    }
}

}  // extern "C"
