// The OA side of the boundary: a TMINLP2OsiLP lineariser (src/Interfaces/BonTMINLP2OsiLP.hpp:24-120,
// installed with BabSetupBase::set_linearizer, src/Algorithms/BonBabSetupBase.hpp:330-332) whose get_oas / extract are
// served by the batched engine.  Osi types are absent from this image: OsiRowCut / OsiCuts / OsiSolverInterface below
// keep the member names the generators use (setRow, setLb, setUb, setGloballyValid, insert, sizeRowCuts, rowCut,
// loadProblem, setInteger) and nothing else.
#pragma once
#include <functional>
#include <vector>

#include "bonmin_shim.hpp"

namespace Bonmin {

class OsiRowCut {
 public:
  void setRow(int size, const int *colIndices, const double *elements) { idx_.assign(colIndices, colIndices + size); el_.assign(elements, elements + size); }
  void setLb(double v) { lb_ = v; }
  void setUb(double v) { ub_ = v; }
  void setGloballyValid(bool t = true) { global_ = t; }
  double lb() const { return lb_; }
  double ub() const { return ub_; }
  bool globallyValid() const { return global_; }
  const std::vector<int> &indices() const { return idx_; }
  const std::vector<double> &elements() const { return el_; }

 private:
  std::vector<int> idx_;
  std::vector<double> el_;
  double lb_ = 0, ub_ = 0;
  bool global_ = false;
};

class OsiCuts {
 public:
  void insert(const OsiRowCut &c) { cuts_.push_back(c); }
  int sizeRowCuts() const { return (int)cuts_.size(); }
  const OsiRowCut &rowCut(int i) const { return cuts_[i]; }

 private:
  std::vector<OsiRowCut> cuts_;
};

class OsiSolverInterface {   // the LP the relaxation is loaded into (CSR by rows)
 public:
  void loadProblem(int ncols, int nrows, const int *start, const int *index, const double *value, const double *collb,
                   const double *colub, const double *obj, const double *rowlb, const double *rowub) {
    ncols_ = ncols; nrows_ = nrows;
    start_.assign(start, start + nrows + 1); index_.assign(index, index + start[nrows]); value_.assign(value, value + start[nrows]);
    collb_.assign(collb, collb + ncols); colub_.assign(colub, colub + ncols); obj_.assign(obj, obj + ncols);
    rowlb_.assign(rowlb, rowlb + nrows); rowub_.assign(rowub, rowub + nrows);
    integer_.assign(ncols, 0);
  }
  void setInteger(int i) { integer_[i] = 1; }
  void setObjOffset(double v) { obj_offset_ = v; }
  int getNumCols() const { return ncols_; }
  int getNumRows() const { return nrows_; }
  int ncols_ = 0, nrows_ = 0;
  std::vector<int> start_, index_, integer_;
  std::vector<double> value_, collb_, colub_, obj_, rowlb_, rowub_;
  double obj_offset_ = 0;
};

class TMINLP2OsiLP {   // BonTMINLP2OsiLP.hpp:24-120
 public:
  virtual ~TMINLP2OsiLP() {}
  virtual TMINLP2OsiLP *clone() const = 0;
  void set_tols(double tiny, double very_tiny, double rhs_relax, double infty) { tiny_ = tiny; very_tiny_ = very_tiny; rhs_relax_ = rhs_relax; infty_ = infty; }
  void set_model(TMINLP2TNLP *model) { model_ = model; }
  virtual void extract(OsiSolverInterface *si, const double *x, bool getObj) = 0;
  virtual void get_refined_oa(OsiCuts &cs) const = 0;
  virtual void get_oas(OsiCuts &cs, const double *x, bool getObj, bool global) const = 0;

 protected:
  double tiny_ = -0., very_tiny_ = -0., rhs_relax_ = 0, infty_ = 0;
  TMINLP2TNLP *model_ = nullptr;
};

// B x getOuterApproximation: same argument meaning as bnlp_oa_capacity + bnlp_oa_batch (include/bnlp.h); the provider
// sizes and fills the CSR arrays
struct OaRows { int rows_per_point = 0; std::vector<int> row_start, col, src_row; std::vector<double> val, lb, ub; };
using OaBatchFn = std::function<int(int B, const double *X, const double *XL, const double *XU, int getObj, double tiny,
                                    double very_tiny, double rhs_relax, double infty, OaRows &out)>;
// extractLinearRelaxation + addObjectiveFunction: same meaning as bnlp_extract_linear_relaxation
struct LpRelaxation { int n_cols = 0, has_alpha = 0; double obj_offset = 0; std::vector<int> row_start, col, is_integer;
                      std::vector<double> val, row_lb, row_ub, col_lb, col_ub, obj; };
using ExtractFn = std::function<int(const double *x, const double *xl, const double *xu, int getObj, double tiny, double very_tiny,
                                    double rhs_relax, double infty, LpRelaxation &out)>;

class B200Linearizer : public TMINLP2OsiLP {
 public:
  B200Linearizer(OaBatchFn oa, ExtractFn ex) : oa_(std::move(oa)), extract_(std::move(ex)) {}
  TMINLP2OsiLP *clone() const override { return new B200Linearizer(*this); }
  void extract(OsiSolverInterface *si, const double *x, bool getObj) override;
  void get_oas(OsiCuts &cs, const double *x, bool getObj, bool global) const override { get_oas_batch(cs, x, 1, getObj, global); }
  // the reference's "refined" OA linearises at a grid of points built from the variable bounds; with a batched engine
  // that is get_oas_batch on those points — the caller supplies them through set_refinement_points
  void get_refined_oa(OsiCuts &cs) const override;
  void set_refinement_points(const std::vector<double> &X) { refine_X_ = X; }
  // many points in one engine call (a MILP solution pool, the points of an ECP round)
  void get_oas_batch(OsiCuts &cs, const double *X, int B, bool getObj, bool global) const;
  int lastError() const { return last_rc_; }

 private:
  OaBatchFn oa_;
  ExtractFn extract_;
  std::vector<double> refine_X_;
  mutable int last_rc_ = 0;
};

}  // namespace Bonmin
