// ORACLE (test infrastructure, never shipped, never on the product path).
//
// CPU restatement of the per-block scratch object the physics modules read and write:
//   Workset<EvalT>                                   src/tools/workset.hpp:21-586
//   field registry addSolutionField                  src/tools/workset.cpp:344-436
//   SolutionField label parsing                      src/tools/fields.hpp:44-128
//   resetResidual                                    src/tools/workset.cpp:497-528
//   computeSolnSteadySeeded                          src/tools/workset.cpp:864-901
//   computeSolnTransientSeeded (seedwhat 0/1/2/3)    src/tools/workset.cpp:600-834
//   evaluateSolutionField (dof-ascending sum, dof 0 assigned first)   src/tools/workset.cpp:978-1111
//   getSolutionField / lazily allocated zero fields  src/tools/workset.cpp:1537-1650
//   getElementSize / getSideElementSize              src/tools/workset.cpp:2699-2733
// and of the group storage it points at (Group / BoundaryGroup / GroupMetaData,
// src/tools/group.hpp:205-267, group.cpp:134-250).
#pragma once
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "function_manager.hpp"
#include "mesh.hpp"

namespace oracle {

struct View4 {  // (elem, dof, pt, comp) row-major, non-owning
  const double* p = nullptr;
  int n1 = 0, n2 = 0, n3 = 0;
  double operator()(int e, int d, int q, int c) const { return p[(((size_t)e * n1 + d) * n2 + q) * n3 + c]; }
  int extent1() const { return n1; }
  int extent2() const { return n2; }
  int extent3() const { return n3; }
};

struct Group {  // volume group: <= workset-size consecutive elements (assemblyManager_groups.hpp:83-426)
  int numElem = 0;
  std::vector<int> elem_ids;
  std::vector<int> LIDs;                       // (E, ndof)
  std::vector<double> nodes;                   // (E, nverts, dim)
  std::vector<double> wts;                     // (E, nqp)
  std::vector<double> ip[3];                   // (E, nqp)
  std::vector<std::vector<double>> basis, basis_grad, basis_curl, basis_div;  // per basis
  // boundary groups only
  int sideset = -1, local_side = -1;
  std::string sidename;
  std::vector<double> normals[3];              // (E, nqp_side)
};

struct TimeData {  // Butcher tableau + BDF weights (solverManager_setup.hpp:181-480)
  bool isTransient = false;
  double time = 0.0, deltat = 1.0;
  int stage = 0;
  std::vector<double> butcher_A;  // (s,s)
  std::vector<double> butcher_b, butcher_c;
  std::vector<double> BDF_wts;
  int nstages() const { return (int)butcher_b.size(); }
};

template <class EvalT>
struct SolutionField {
  std::string expression;
  int var = 0;
  std::string derivative_type;
  int component = 0;
  bool is_updated = false, is_initialized = false;
  View2<EvalT> data;
  SolutionField(const std::string& e, int v) : expression(e), var(v) {  // fields.hpp:44-128
    if (e.find("[x]") != std::string::npos) component = 0;
    if (e.find("[y]") != std::string::npos) component = 1;
    if (e.find("[z]") != std::string::npos) component = 2;
    if (e.find("grad") != std::string::npos) derivative_type = "grad";
    if (e.find("div") != std::string::npos) derivative_type = "div";
    if (e.find("curl") != std::string::npos) derivative_type = "curl";
    if (e.find("_t") != std::string::npos) derivative_type = "time";
  }
};

struct ScalarField {
  std::string expression;
  View2<double> data;
};

template <class EvalT>
class Workset {
 public:
  int maxElem = 0, numElem = 0, dimension = 0, numip = 0, numsideip = 0;
  std::vector<VarInfo> vars;
  std::vector<std::vector<int>> offsets;
  std::vector<int> usebasis;
  int maxRes = 0;

  // handles into the current group's storage (updateWorkset, assemblyManager_workset.hpp:963-1049)
  std::vector<View4> basis, basis_grad, basis_curl, basis_div;
  std::vector<View4> basis_side, basis_grad_side;
  const double* wts_p = nullptr;
  const double* wts_side_p = nullptr;
  double wts(int e, int q) const { return wts_p[(size_t)e * numip + q]; }
  double wts_side(int e, int q) const { return wts_side_p[(size_t)e * numsideip + q]; }

  View2<EvalT> res;
  std::vector<View2<EvalT>> sol_vals, sol_dot_vals;  // per variable (E, dof)
  std::vector<SolutionField<EvalT>> soln_fields, side_soln_fields;
  std::vector<ScalarField> scalar_fields, side_scalar_fields;

  bool isOnSide = false, isTransient = false;
  bool isAdjoint = false;   // Workset::isAdjoint (updateWorksetAdjoint, assemblyManager_workset.hpp:310-350)
  double time = 0.0, deltat = 1.0, alpha = 1.0;
  int current_stage = 0;
  TimeData td;
  std::string sidename;
  int currentside = 0;
  std::vector<std::vector<std::string>> var_bcs;  // [var][sideset]

  Workset() {}
  Workset(int maxElem_, int dim, int numip_, int numsideip_, const DofMap& dm)
      : maxElem(maxElem_), numElem(maxElem_), dimension(dim), numip(numip_), numsideip(numsideip_) {
    vars = dm.vars;
    offsets = dm.offsets;
    int maxdof = 0;
    for (auto& v : vars) { usebasis.push_back(v.basis); maxdof = std::max(maxdof, dm.bases[v.basis].card); }
    maxRes = (int)vars.size() * maxdof;  // workset.cpp:152-182
    res = View2<EvalT>(maxElem, maxRes);
    for (auto& v : vars) {
      sol_vals.push_back(View2<EvalT>(maxElem, dm.bases[v.basis].card));
      sol_dot_vals.push_back(View2<EvalT>(maxElem, dm.bases[v.basis].card));
    }
    for (size_t v = 0; v < vars.size(); ++v) addSolutionField(vars[v].name, (int)v, vars[v].btype);
    for (const char* s : {"x", "y", "z"}) {  // addScalarFields: ip coordinates
      scalar_fields.push_back({s, View2<double>(maxElem, numip)});
      side_scalar_fields.push_back({s, View2<double>(maxElem, numsideip)});
    }
    for (const char* s : {"n[x]", "n[y]", "n[z]"}) side_scalar_fields.push_back({s, View2<double>(maxElem, numsideip)});
    basis.resize(dm.bases.size()); basis_grad.resize(dm.bases.size()); basis_curl.resize(dm.bases.size());
    basis_div.resize(dm.bases.size()); basis_side.resize(dm.bases.size()); basis_grad_side.resize(dm.bases.size());
  }

  void addSolutionField(const std::string& var, int vi, const std::string& btype) {  // workset.cpp:344-436
    auto add = [&](std::vector<SolutionField<EvalT>>& list, const std::string& e) { list.push_back(SolutionField<EvalT>(e, vi)); };
    if (btype.substr(0, 5) == "HGRAD") {
      for (auto* l : {&soln_fields, &side_soln_fields}) {
        add(*l, var); add(*l, "grad(" + var + ")[x]"); add(*l, "grad(" + var + ")[y]"); add(*l, "grad(" + var + ")[z]");
      }
      add(soln_fields, var + "_t");
    } else if (btype.substr(0, 4) == "HDIV") {
      for (auto* l : {&soln_fields, &side_soln_fields}) {
        add(*l, var + "[x]"); add(*l, var + "[y]"); add(*l, var + "[z]"); add(*l, "div(" + var + ")");
      }
      add(soln_fields, var + "_t[x]"); add(soln_fields, var + "_t[y]"); add(soln_fields, var + "_t[z]");
    } else if (btype.substr(0, 4) == "HVOL") {
      add(soln_fields, var); add(soln_fields, var + "_t"); add(side_soln_fields, var);
    } else if (btype.substr(0, 5) == "HCURL") {
      for (auto* l : {&soln_fields, &side_soln_fields}) {
        add(*l, var + "[x]"); add(*l, var + "[y]"); add(*l, var + "[z]");
        add(*l, "curl(" + var + ")[x]"); add(*l, "curl(" + var + ")[y]"); add(*l, "curl(" + var + ")[z]");
      }
      add(soln_fields, var + "_t[x]"); add(soln_fields, var + "_t[y]"); add(soln_fields, var + "_t[z]");
    }
  }

  void reset() {  // Workset::reset: mark fields stale, zero residual
    for (auto& f : soln_fields) f.is_updated = false;
    for (auto& f : side_soln_fields) f.is_updated = false;
    resetResidual();
  }
  void resetResidual() {  // workset.cpp:497-528
    for (auto& x : res.a) x = EvalT(0.0);
  }

  // ---- seeding ------------------------------------------------------------------------
  // u: gathered (E, nvar, maxdof) values
  void computeSolnSteadySeeded(const std::vector<double>& u, int maxdof, int seedwhat) {  // workset.cpp:864-901
    for (size_t var = 0; var < vars.size(); ++var) {
      auto& u_AD = sol_vals[var];
      for (int elem = 0; elem < numElem; ++elem)
        for (int dof = 0; dof < u_AD.n1; ++dof) {
          const double val = u[((size_t)elem * vars.size() + var) * maxdof + dof];
          if (seedwhat == 1) u_AD(elem, dof) = ADTraits<EvalT>::seed(offsets[var][dof], val);
          else u_AD(elem, dof) = EvalT(val);
        }
    }
  }
  // u_prev (E,nvar,maxdof,nsteps), u_stage (E,nvar,maxdof,nstages)
  void computeSolnTransientSeeded(const std::vector<double>& u, const std::vector<double>& u_prev, const std::vector<double>& u_stage,
                                  int maxdof, int nsteps, int nstages, int seedwhat, int index = 0) {  // workset.cpp:600-834
    // seedwhat 1: the stage solution; 2: previous step `index` (compute_previous_jac, :669-726); 3: previous stage `index` (:727-785)
    const double dt = deltat;
    const int stage = current_stage;
    const int ns = td.nstages();
    auto b_A = [&](int i, int j) { return td.butcher_A[(size_t)i * ns + j]; };
    for (size_t var = 0; var < vars.size(); ++var) {
      auto& u_AD = sol_vals[var];
      auto& u_dot_AD = sol_dot_vals[var];
      for (int elem = 0; elem < numElem; ++elem) {
        const double alpha_u = b_A(stage, stage) / td.butcher_b[stage];
        const double timewt = 1.0 / dt / td.butcher_b[stage];
        const double alpha_t = td.BDF_wts[0] * timewt;
        for (int dof = 0; dof < u_AD.n1; ++dof) {
          const size_t base = ((size_t)elem * vars.size() + var) * maxdof + dof;
          auto cu_prev = [&](int s) { return u_prev[base * nsteps + s]; };
          auto cu_stage = [&](int s) { return u_stage[base * nstages + s]; };
          if (seedwhat == 2 || seedwhat == 3) {
            const double stageval = u[base];
            EvalT u_prev_val = EvalT(cu_prev(0));
            if (seedwhat == 2 && index == 0) u_prev_val = ADTraits<EvalT>::seed(offsets[var][dof], cu_prev(0));
            EvalT beta_u = (1.0 - alpha_u) * u_prev_val;
            for (int s = 0; s < stage; s++) {
              EvalT u_stage_val = EvalT(cu_stage(s));
              if (seedwhat == 3 && index == s) u_stage_val = ADTraits<EvalT>::seed(offsets[var][dof], cu_stage(s));
              beta_u += b_A(stage, s) / td.butcher_b[s] * (u_stage_val - u_prev_val);
            }
            u_AD(elem, dof) = alpha_u * stageval + beta_u;
            EvalT beta_t = EvalT(0.0);
            for (size_t s = 1; s < td.BDF_wts.size(); s++) {
              EvalT pv = EvalT(cu_prev((int)s - 1));
              if (seedwhat == 2 && index == (int)s - 1) pv = ADTraits<EvalT>::seed(offsets[var][dof], cu_prev((int)s - 1));
              beta_t += td.BDF_wts[s] * pv;
            }
            beta_t *= timewt;
            u_dot_AD(elem, dof) = alpha_t * stageval + beta_t;
            continue;
          }
          EvalT stageval = (seedwhat == 1) ? ADTraits<EvalT>::seed(offsets[var][dof], u[base]) : EvalT(u[base]);
          double beta_u = (1.0 - alpha_u) * cu_prev(0);
          for (int s = 0; s < stage; s++) beta_u += b_A(stage, s) / td.butcher_b[s] * (cu_stage(s) - cu_prev(0));
          u_AD(elem, dof) = alpha_u * stageval + beta_u;
          double beta_t = 0.0;
          for (size_t s = 1; s < td.BDF_wts.size(); s++) beta_t += td.BDF_wts[s] * cu_prev((int)s - 1);
          beta_t *= timewt;
          u_dot_AD(elem, dof) = alpha_t * stageval + beta_t;
        }
      }
    }
  }

  // ---- fields -------------------------------------------------------------------------
  int findField(const std::string& label, bool side) const {
    const auto& list = side ? side_soln_fields : soln_fields;
    for (size_t i = 0; i < list.size(); ++i) if (list[i].expression == label) return (int)i;
    return -1;
  }
  int findScalarField(const std::string& label, bool side) const {
    const auto& list = side ? side_scalar_fields : scalar_fields;
    for (size_t i = 0; i < list.size(); ++i) if (list[i].expression == label) return (int)i;
    return -1;
  }
  void checkAllocation(SolutionField<EvalT>& f, bool side) {  // workset.cpp:1632-1650 (zero-initialised on first use)
    if (!f.is_initialized) { f.data = View2<EvalT>(maxElem, side ? numsideip : numip); f.is_initialized = true; }
  }
  View2<EvalT>& getSolutionField(const std::string& label) {  // workset.cpp:1537-1625
    const bool side = isOnSide;
    int i = findField(label, side);
    if (i < 0) throw std::runtime_error("Error: could not find a solution field named " + label);
    return getSolutionFieldByIndex(i, side);
  }
  View2<EvalT>& getSolutionFieldByIndex(int i, bool side) {
    auto& f = side ? side_soln_fields[i] : soln_fields[i];
    checkAllocation(f, side);
    if (!f.is_updated) evaluateSolutionField(f, side);
    return f.data;
  }
  View2<double>& getScalarField(const std::string& label) {
    int i = findScalarField(label, isOnSide);
    if (i < 0) throw std::runtime_error("Error: could not find a scalar field named " + label);
    return isOnSide ? side_scalar_fields[i].data : scalar_fields[i].data;
  }

  void evaluateSolutionField(SolutionField<EvalT>& f, bool side) {  // workset.cpp:978-1111
    bool proceed = true;
    if (f.derivative_type == "time") {
      if (!isTransient) proceed = false;
      else if (side) proceed = false;
    }
    if (!proceed) return;  // field keeps its zero initial content (steady *_t terms are exact zeros)
    const View2<EvalT>& solvals = (f.derivative_type == "time") ? sol_dot_vals[f.var] : sol_vals[f.var];
    const int b = usebasis[f.var];
    auto& fd = f.data;
    if (f.derivative_type == "div") {
      const View4& sb = basis_div[b];
      for (int elem = 0; elem < numElem; ++elem) {
        for (int pt = 0; pt < sb.extent2(); ++pt) fd(elem, pt) = solvals(elem, 0) * sb(elem, 0, pt, 0);
        for (int dof = 1; dof < sb.extent1(); ++dof)
          for (int pt = 0; pt < sb.extent2(); ++pt) fd(elem, pt) += solvals(elem, dof) * sb(elem, dof, pt, 0);
      }
    } else {
      const View4* cb;
      if (f.derivative_type == "grad") cb = side ? &basis_grad_side[b] : &basis_grad[b];
      else if (f.derivative_type == "curl") cb = &basis_curl[b];
      else cb = side ? &basis_side[b] : &basis[b];
      if (f.component >= cb->extent3()) {
        for (auto& x : fd.a) x = EvalT(0.0);
        f.is_updated = true;
        return;
      }
      const int c = f.component;
      for (int elem = 0; elem < numElem; ++elem) {
        for (int pt = 0; pt < cb->extent2(); ++pt) fd(elem, pt) = solvals(elem, 0) * (*cb)(elem, 0, pt, c);
        for (int dof = 1; dof < cb->extent1(); ++dof)
          for (int pt = 0; pt < cb->extent2(); ++pt) fd(elem, pt) += solvals(elem, dof) * (*cb)(elem, dof, pt, c);
      }
    }
    f.is_updated = true;
  }

  std::vector<double> getElementSize() const {  // workset.cpp:2699-2712
    std::vector<double> h(numElem);
    for (int e = 0; e < numElem; ++e) {
      double vol = 0.0;
      for (int i = 0; i < numip; ++i) vol += wts(e, i);
      h[e] = std::pow(vol, 1.0 / (double)dimension);
    }
    return h;
  }
  std::vector<double> getSideElementSize() const {  // workset.cpp:2718-2733
    std::vector<double> h(numElem);
    for (int e = 0; e < numElem; ++e) {
      double vol = 0.0;
      for (int i = 0; i < numsideip; ++i) vol += wts_side(e, i);
      h[e] = std::pow(vol, 1.0 / ((double)dimension - 1.0));
    }
    return h;
  }

  // hooks the function manager uses to find and pull workset data
  void connect(FunctionManager<EvalT>& fm) {
    fm.hooks.find_soln_field = [this](const std::string& e, const std::string& loc) { return findField(e, loc == "side ip"); };
    fm.hooks.find_scalar_field = [this](const std::string& e, const std::string& loc) { return findScalarField(e, loc == "side ip"); };
    fm.hooks.get_soln_field = [this](int i, const std::string& loc) { return &getSolutionFieldByIndex(i, loc == "side ip"); };
    fm.hooks.get_scalar_field = [this](int i, const std::string& loc) { return loc == "side ip" ? &side_scalar_fields[i].data : &scalar_fields[i].data; };
    fm.hooks.get_time = [this]() { return time; };
    fm.hooks.is_on_side = [this]() { return isOnSide; };
  }
};

}  // namespace oracle
