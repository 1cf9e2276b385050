// Device-resident fused Operator / Cone structs of the totsu_b200 crate: they plug into
// the UNMODIFIED solver through the public traits, the extension path examples/imgnr_udef demonstrates
// (examples/imgnr_udef/src/main.rs:62-67), and remove the O(#blocks) launch pattern of the stock front-ends
// (e.g. totsu/src/problem/socp.rs:83-124: one gemv + one dot per cone block).
//
//   DenseOp<F>      one stacked column-major A (row-sharded across ranks) as an Operator  -> tb_denseop_*
//   ProductCone<F>  Zero / RPos / SOC / RotSOC / PSD blocks back to back as a Cone         -> tb_cone_*
#pragma once
#include "solver.hpp"

namespace totsu_b200 {

template <typename F> class DenseOp : public Operator<F> {
public:
    using Sl = Slice<F>;
    // `mat`: view of the local rows (column-major n_row_local x n_col, lda = n_row_local)
    DenseOp(tb_view mat, size_t n_row_local, size_t n_col, size_t row_offset = 0, size_t n_row_total = 0) {
        n_row_total_ = n_row_total ? n_row_total : n_row_local;
        n_col_ = n_col;
        TBH_CALL(tb_denseop_create(Abi<F>::dtype, mat, n_row_local, n_col, row_offset, n_row_total_, &h_));
    }
    ~DenseOp() override {
        if (h_) tb_denseop_destroy(h_);
    }
    DenseOp(const DenseOp&) = delete;
    DenseOp& operator=(const DenseOp&) = delete;

    std::pair<size_t, size_t> size() const override { return {n_row_total_, n_col_}; }
    void op(F alpha, const Sl& x, F beta, Sl& y) const override { TBH_CALL(Abi<F>::denseop_apply(h_, 0, alpha, x.view(), beta, y.view())); }
    void trans_op(F alpha, const Sl& x, F beta, Sl& y) const override { TBH_CALL(Abi<F>::denseop_apply(h_, 1, alpha, x.view(), beta, y.view())); }
    void absadd_cols(Sl& tau) const override { TBH_CALL(Abi<F>::denseop_absadd_cols(h_, tau.view())); }
    void absadd_rows(Sl& sigma) const override { TBH_CALL(Abi<F>::denseop_absadd_rows(h_, sigma.view())); }
    // one read of A for an op / trans_op pair on independent vectors
    void op_pair(F alpha_n, const Sl& x_n, F beta_n, Sl& y_n, F alpha_t, const Sl& x_t, F beta_t, Sl& y_t) const {
        TBH_CALL(Abi<F>::denseop_apply_pair(h_, alpha_n, x_n.view(), beta_n, y_n.view(), alpha_t, x_t.view(), beta_t, y_t.view()));
    }
    tb_handle handle() const { return h_; }

private:
    tb_handle h_ = 0;
    size_t n_row_total_ = 0, n_col_ = 0;
};

template <typename F> class ProductCone : public Cone<F> {
public:
    using Sl = Slice<F>;
    ProductCone(const std::vector<tb_cone_block>& blocks, F eps_zero) : blocks_(blocks), eps_zero_(eps_zero) {
        TBH_CALL(tb_cone_create(blocks_.data(), blocks_.size(), &h_));
        size_t wl = 0;
        for (const auto& b : blocks_) {
            if (b.type == TB_CONE_PSD) {
                size_t k = 0;
                while ((k + 1) * (k + 2) / 2 <= b.len) ++k;
                wl = std::max(wl, (size_t)tb_map_eig_worklen(k));
            }
        }
        if (wl > 0) {
            psd_work_host_.assign(wl, F(0));
            psd_work_ = Sl::new_mut(psd_work_host_.data(), wl);
        }
    }
    ~ProductCone() override {
        psd_work_.drop();
        if (h_) tb_cone_destroy(h_);
    }
    bool proj(bool dual_cone, Sl& x) override {
        int st = Abi<F>::cone_proj(h_, dual_cone ? 1 : 0, x.view(), eps_zero_, psd_work_.view());
        if (st == TB_ERR_ARG) return false;       // work shortage / malformed block: Err(()) -> ConeFailure
        tb_check(st, "tb_cone_proj");
        return true;
    }
    // the solver's `group` closure is the min-fill (solver.rs:509-518); it runs on the device for every block
    void product_group(Sl& dp_tau, const typename Cone<F>::Group&) const override {
        TBH_CALL(Abi<F>::cone_group_min(h_, dp_tau.view()));
    }

private:
    std::vector<tb_cone_block> blocks_;
    F eps_zero_;
    tb_handle h_ = 0;
    std::vector<F> psd_work_host_;
    Sl psd_work_;
};

}  // namespace totsu_b200
