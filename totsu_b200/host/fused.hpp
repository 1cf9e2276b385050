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
    // The PSD work area (2k^2 + k elements, cone_psd.rs:32-38) is a device-only buffer allocated ONCE here, exactly like
    // rust/totsu_b200/src/fused.rs: proj() neither wraps nor releases anything, so the first projection of an iteration
    // stays parked until the second arrives (csrc/cone.cu "pairing of the two projections").
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
            tb_handle wh = 0;
            TBH_CALL(tb_buf_alloc(Abi<F>::dtype, wl, &wh));
            psd_work_ = tb_view{wh, 0, wl};
        }
    }
    ~ProductCone() override {
        if (psd_work_.buf) tb_buf_release(psd_work_.buf);
        if (h_) tb_cone_destroy(h_);
    }
    ProductCone(const ProductCone&) = delete;
    ProductCone& operator=(const ProductCone&) = delete;
    bool proj(bool dual_cone, Sl& x) override {
        // arguments are validated at submit time (before the call may be parked), so TB_ERR_ARG can only come from here
        int st = Abi<F>::cone_proj(h_, dual_cone ? 1 : 0, x.view(), eps_zero_, psd_work_);
        if (st == TB_ERR_ARG) return false;       // work shortage / malformed block: Err(()) -> ConeFailure
        tb_check(st, "tb_cone_proj");
        return true;
    }
    // cone.rs:20-29: split dp_tau into the blocks and call `group` once per block of size > 1.  The solver passes the
    // min-fill closure (solver.rs:509-518), which tb_cone_group_min computes for every block in one launch; any OTHER
    // closure gets the trait's contract literally - one `group` call per block on its sub-slice.
    void product_group(Sl& dp_tau, const typename Cone<F>::Group& group) const override {
        if (is_min_fill(group)) {
            TBH_CALL(Abi<F>::cone_group_min(h_, dp_tau.view()));
            return;
        }
        size_t done = 0;
        for (const auto& b : blocks_) {
            const size_t len = (size_t)b.len;
            Sl t_blk = dp_tau.sub(done, len);
            done += len;
            if (len > 1 && b.type != TB_CONE_ZERO && b.type != TB_CONE_RPOS) group(t_blk);      // cone_zero.rs:46-49, cone_rpos.rs:47-50
        }
    }

private:
    // probe the closure on a 3-element slice: the solver's closure fills the group with its minimum
    static bool is_min_fill(const typename Cone<F>::Group& group) {
        F probe[3] = {F(3), F(1), F(2)};
        {
            Sl sl = Sl::new_mut(probe, 3);
            group(sl);
        }
        return probe[0] == F(1) && probe[1] == F(1) && probe[2] == F(1);
    }
    std::vector<tb_cone_block> blocks_;
    F eps_zero_;
    tb_handle h_ = 0;
    tb_view psd_work_{0, 0, 0};
};

}  // namespace totsu_b200
