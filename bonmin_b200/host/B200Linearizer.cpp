#include "B200Linearizer.hpp"

namespace Bonmin {

void B200Linearizer::get_oas_batch(OsiCuts &cs, const double *X, int B, bool getObj, bool global) const {
  if (!model_ || B <= 0) return;
  OaRows r;
  last_rc_ = oa_(B, X, model_->x_l_.data(), model_->x_u_.data(), getObj ? 1 : 0, tiny_, very_tiny_, rhs_relax_, infty_, r);
  if (last_rc_) return;
  const int slots = B * r.rows_per_point;
  for (int k = 0; k < slots; ++k) {
    if (r.src_row[k] < 0) continue;                 // row not generated (linear row, cleaned away, not violated)
    const int a = r.row_start[k], b = r.row_start[k + 1];
    OsiRowCut cut;
    cut.setRow(b - a, r.col.data() + a, r.val.data() + a);
    cut.setLb(r.lb[k]);
    cut.setUb(r.ub[k]);
    if (global) cut.setGloballyValid();             // BonOsiTMINLPInterface.cpp:2040-2043
    cs.insert(cut);
  }
}

void B200Linearizer::get_refined_oa(OsiCuts &cs) const {
  if (!model_ || refine_X_.empty()) return;
  get_oas_batch(cs, refine_X_.data(), (int)(refine_X_.size() / model_->n), false, true);
}

void B200Linearizer::extract(OsiSolverInterface *si, const double *x, bool getObj) {
  if (!model_ || !si) return;
  LpRelaxation lp;
  last_rc_ = extract_(x, model_->x_l_.data(), model_->x_u_.data(), getObj ? 1 : 0, tiny_, very_tiny_, rhs_relax_, infty_, lp);
  if (last_rc_) return;
  const int nrows = (int)lp.row_lb.size();
  si->loadProblem(lp.n_cols, nrows, lp.row_start.data(), lp.col.data(), lp.val.data(), lp.col_lb.data(), lp.col_ub.data(),
                  lp.obj.data(), lp.row_lb.data(), lp.row_ub.data());
  for (int j = 0; j < lp.n_cols; ++j)
    if (lp.is_integer[j]) si->setInteger(j);        // BonOsiTMINLPInterface.cpp:2525-2531
  si->setObjOffset(lp.obj_offset);
}

}  // namespace Bonmin
