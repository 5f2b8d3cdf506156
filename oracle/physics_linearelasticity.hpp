// ORACLE (test infrastructure, never shipped, never on the product path).
//
// CPU restatement of the linear elasticity module:
//   ctor / variable list / modelparams   src/physics/linearelasticity.cpp:19-59
//   defineFunctions                      src/physics/linearelasticity.cpp:64-86   (mu defaults to 0.5)
//   volumeResidual                       src/physics/linearelasticity.cpp:91-238  (2-D: pt-outer, 3-D: dof-outer)
//   boundaryResidual                     src/physics/linearelasticity.cpp:243-674 (Neumann traction, Nitsche weak Dirichlet)
//   computeStress                        src/physics/linearelasticity.cpp:915-1273
// Restated: 2-D and 3-D, Lame form, `incplanestress`.  Not restated (no configuration in BASELINE.json uses
// them): crystal elasticity, the thermo-elastic (variable T) and Biot (variable p) couplings, 1-D, "interface" sides.
#pragma once
#include "physics_base.hpp"

namespace oracle {

template <class EvalT>
class linearelasticity : public PhysicsBase<EvalT> {
 public:
  using PhysicsBase<EvalT>::wkset;
  using PhysicsBase<EvalT>::functionManager;
  int spaceDim = 3;
  int dx_num = -1, dy_num = -1, dz_num = -1;
  bool incplanestress = false;
  double modelparams[5] = {1.0, 10.0, 0.0, 0.0, 1.0e-6};
  std::vector<EvalT> stress;  // (elem, pt, d, d)
  int stress_np = 0;
  EvalT& S(int e, int k, int a, int b) { return stress[(((size_t)e * stress_np + k) * spaceDim + a) * spaceDim + b]; }

  linearelasticity(const Settings& settings, int dim) : spaceDim(dim) {
    this->label = "linearelasticity";
    if (dim < 2) throw std::runtime_error("oracle: 1-D linear elasticity is not restated");
    this->myvars = {"dx", "dy"};
    this->mybasistypes = {"HGRAD", "HGRAD"};
    if (dim == 3) { this->myvars.push_back("dz"); this->mybasistypes.push_back("HGRAD"); }
    if (settings.getb("use crystal elasticity", false) || settings.getb("Biot", false))
      throw std::runtime_error("oracle: crystal elasticity / Biot coupling are not restated");
    incplanestress = settings.getb("incplanestress", false);
    modelparams[0] = settings.getd("form_param", 1.0);
    modelparams[1] = settings.getd("penalty", 10.0);
  }

  void defineFunctions(const Settings& fs, FunctionManager<EvalT>* fm) override {
    functionManager = fm;
    fm->addFunction("lambda", fs.get("lambda", "1.0"), "ip");
    fm->addFunction("mu", fs.get("mu", "0.5"), "ip");
    fm->addFunction("source dx", fs.get("source dx", "0.0"), "ip");
    fm->addFunction("source dy", fs.get("source dy", "0.0"), "ip");
    fm->addFunction("source dz", fs.get("source dz", "0.0"), "ip");
    fm->addFunction("lambda", fs.get("lambda", "1.0"), "side ip");
    fm->addFunction("mu", fs.get("mu", "0.5"), "side ip");
  }

  void setWorkset(Workset<EvalT>* w) override {
    wkset = w;
    dx_num = this->findVar("dx"); dy_num = this->findVar("dy"); dz_num = this->findVar("dz");
  }

  void computeStress(Vista<EvalT>& lambda, Vista<EvalT>& mu, bool onside) {
    const int np = onside ? wkset->numsideip : wkset->numip;
    stress_np = np;
    stress.assign((size_t)wkset->maxElem * np * spaceDim * spaceDim, EvalT(0.0));
    if (spaceDim == 2) {
      auto& ddx_dx = wkset->getSolutionField("grad(dx)[x]");
      auto& ddx_dy = wkset->getSolutionField("grad(dx)[y]");
      auto& ddy_dx = wkset->getSolutionField("grad(dy)[x]");
      auto& ddy_dy = wkset->getSolutionField("grad(dy)[y]");
      for (int e = 0; e < wkset->numElem; ++e)
        for (int k = 0; k < np; ++k) {
          if (incplanestress) {
            S(e, k, 0, 0) = 4.0 * mu(e, k) * ddx_dx(e, k) + 2.0 * mu(e, k) * ddy_dy(e, k);
            S(e, k, 0, 1) = mu(e, k) * (ddx_dy(e, k) + ddy_dx(e, k));
            S(e, k, 1, 0) = mu(e, k) * (ddx_dy(e, k) + ddy_dx(e, k));
            S(e, k, 1, 1) = 4.0 * mu(e, k) * ddy_dy(e, k) + 2.0 * mu(e, k) * ddx_dx(e, k);
          } else {
            S(e, k, 0, 0) = (2.0 * mu(e, k) + lambda(e, k)) * ddx_dx(e, k) + lambda(e, k) * ddy_dy(e, k);
            S(e, k, 0, 1) = mu(e, k) * (ddx_dy(e, k) + ddy_dx(e, k));
            S(e, k, 1, 0) = mu(e, k) * (ddx_dy(e, k) + ddy_dx(e, k));
            S(e, k, 1, 1) = (2.0 * mu(e, k) + lambda(e, k)) * ddy_dy(e, k) + lambda(e, k) * ddx_dx(e, k);
          }
        }
    } else {
      auto& ddx_dx = wkset->getSolutionField("grad(dx)[x]");
      auto& ddx_dy = wkset->getSolutionField("grad(dx)[y]");
      auto& ddx_dz = wkset->getSolutionField("grad(dx)[z]");
      auto& ddy_dx = wkset->getSolutionField("grad(dy)[x]");
      auto& ddy_dy = wkset->getSolutionField("grad(dy)[y]");
      auto& ddy_dz = wkset->getSolutionField("grad(dy)[z]");
      auto& ddz_dx = wkset->getSolutionField("grad(dz)[x]");
      auto& ddz_dy = wkset->getSolutionField("grad(dz)[y]");
      auto& ddz_dz = wkset->getSolutionField("grad(dz)[z]");
      for (int e = 0; e < wkset->numElem; ++e)
        for (int k = 0; k < np; ++k) {
          S(e, k, 0, 0) = (2.0 * mu(e, k) + lambda(e, k)) * ddx_dx(e, k) + lambda(e, k) * (ddy_dy(e, k) + ddz_dz(e, k));
          S(e, k, 0, 1) = mu(e, k) * (ddx_dy(e, k) + ddy_dx(e, k));
          S(e, k, 0, 2) = mu(e, k) * (ddx_dz(e, k) + ddz_dx(e, k));
          S(e, k, 1, 0) = mu(e, k) * (ddx_dy(e, k) + ddy_dx(e, k));
          S(e, k, 1, 1) = (2.0 * mu(e, k) + lambda(e, k)) * ddy_dy(e, k) + lambda(e, k) * (ddx_dx(e, k) + ddz_dz(e, k));
          S(e, k, 1, 2) = mu(e, k) * (ddy_dz(e, k) + ddz_dy(e, k));
          S(e, k, 2, 0) = mu(e, k) * (ddx_dz(e, k) + ddz_dx(e, k));
          S(e, k, 2, 1) = mu(e, k) * (ddy_dz(e, k) + ddz_dy(e, k));
          S(e, k, 2, 2) = (2.0 * mu(e, k) + lambda(e, k)) * ddz_dz(e, k) + lambda(e, k) * (ddx_dx(e, k) + ddy_dy(e, k));
        }
    }
  }

  void volumeResidual() override {
    Vista<EvalT> lambda, mu, source[3];
    source[0] = functionManager->evaluate("source dx", "ip");
    source[1] = functionManager->evaluate("source dy", "ip");
    if (spaceDim > 2) source[2] = functionManager->evaluate("source dz", "ip");
    lambda = functionManager->evaluate("lambda", "ip");
    mu = functionManager->evaluate("mu", "ip");
    computeStress(lambda, mu, false);
    auto& res = wkset->res;
    const int vnum[3] = {dx_num, dy_num, dz_num};
    for (int d = 0; d < spaceDim; ++d) {
      const int b = wkset->usebasis[vnum[d]];
      const View4& basis = wkset->basis[b];
      const View4& basis_grad = wkset->basis_grad[b];
      const auto& off = wkset->offsets[vnum[d]];
      for (int elem = 0; elem < wkset->numElem; ++elem) {
        if (spaceDim == 2) {
          for (int pt = 0; pt < basis.extent2(); ++pt)
            for (int dof = 0; dof < basis.extent1(); ++dof)
              res(elem, off[dof]) += (S(elem, pt, d, 0) * basis_grad(elem, dof, pt, 0) + S(elem, pt, d, 1) * basis_grad(elem, dof, pt, 1) -
                                      source[d](elem, pt) * basis(elem, dof, pt, 0)) * wkset->wts(elem, pt);
        } else {
          for (int dof = 0; dof < basis.extent1(); ++dof)
            for (int pt = 0; pt < basis.extent2(); ++pt)
              res(elem, off[dof]) += (S(elem, pt, d, 0) * basis_grad(elem, dof, pt, 0) + S(elem, pt, d, 1) * basis_grad(elem, dof, pt, 1) +
                                      S(elem, pt, d, 2) * basis_grad(elem, dof, pt, 2) - source[d](elem, pt) * basis(elem, dof, pt, 0)) * wkset->wts(elem, pt);
        }
      }
    }
  }

  void boundaryResidual() override {
    const int cside = wkset->currentside;
    const int vnum[3] = {dx_num, dy_num, dz_num};
    const char* vname[3] = {"dx", "dy", "dz"};
    std::string sidetype[3] = {"Dirichlet", "Dirichlet", "Dirichlet"};
    for (int d = 0; d < spaceDim; ++d) sidetype[d] = wkset->var_bcs[vnum[d]][cside];
    if (sidetype[0] == "Dirichlet" && sidetype[1] == "Dirichlet" && sidetype[2] == "Dirichlet") return;
    Vista<EvalT> source[3], lambda_side, mu_side;
    for (int d = 0; d < spaceDim; ++d) {
      if (sidetype[d] == "Neumann") source[d] = functionManager->evaluate(std::string("Neumann ") + vname[d] + " " + wkset->sidename, "side ip");
      else if (sidetype[d] == "weak Dirichlet") source[d] = functionManager->evaluate(std::string("Dirichlet ") + vname[d] + " " + wkset->sidename, "side ip");
    }
    lambda_side = functionManager->evaluate("lambda", "side ip");
    mu_side = functionManager->evaluate("mu", "side ip");
    auto h = wkset->getSideElementSize();
    auto& res = wkset->res;
    computeStress(lambda_side, mu_side, true);
    View2<double>* n[3] = {&wkset->getScalarField("n[x]"), &wkset->getScalarField("n[y]"), &wkset->getScalarField("n[z]")};
    View2<EvalT>* disp[3] = {nullptr, nullptr, nullptr};
    for (int d = 0; d < spaceDim; ++d) disp[d] = &wkset->getSolutionField(vname[d]);
    for (int d = 0; d < spaceDim; ++d) {
      const int b = wkset->usebasis[vnum[d]];
      const View4& basis = wkset->basis_side[b];
      const View4& basis_grad = wkset->basis_grad_side[b];
      const auto& off = wkset->offsets[vnum[d]];
      if (sidetype[d] == "Neumann") {
        for (int e = 0; e < wkset->numElem; ++e)
          for (int k = 0; k < basis.extent2(); ++k)
            for (int i = 0; i < basis.extent1(); ++i) res(e, off[i]) += (-source[d](e, k) * basis(e, i, k, 0)) * wkset->wts_side(e, k);
      } else if (sidetype[d] == "weak Dirichlet") {
        for (int e = 0; e < wkset->numElem; ++e)
          for (int k = 0; k < basis.extent2(); ++k) {
            const EvalT lam = lambda_side(e, k), mu = mu_side(e, k);
            EvalT penalty = modelparams[1] * (lam + 2.0 * mu) / h[e];
            EvalT delta[3] = {EvalT(0.0), EvalT(0.0), EvalT(0.0)};
            for (int c = 0; c < spaceDim; ++c) delta[c] = (*disp[c])(e, k) - source[c](e, k);
            const double nn[3] = {(*n[0])(e, k), (*n[1])(e, k), spaceDim > 2 ? (*n[2])(e, k) : 0.0};
            // b_c = (C : (delta (x) n))_{d c}: row d of the "adjoint-consistency" flux (linearelasticity.cpp:373-375, 503-509)
            EvalT bb[3] = {EvalT(0.0), EvalT(0.0), EvalT(0.0)};
            if (spaceDim == 2) {
              if (d == 0) {
                bb[0] = (lam + 2.0 * mu) * delta[0] * nn[0] + lam * delta[1] * nn[1];
                bb[1] = mu * delta[1] * nn[0] + mu * delta[0] * nn[1];
              } else {
                bb[0] = mu * delta[1] * nn[0] + mu * delta[0] * nn[1];
                bb[1] = lam * delta[0] * nn[0] + (lam + 2.0 * mu) * delta[1] * nn[1];
              }
            } else {
              if (d == 0) {
                bb[0] = (lam + 2.0 * mu) * delta[0] * nn[0] + lam * delta[1] * nn[1] + lam * delta[2] * nn[2];
                bb[1] = mu * delta[1] * nn[0] + mu * delta[0] * nn[1];
                bb[2] = mu * delta[2] * nn[0] + mu * delta[0] * nn[2];
              } else if (d == 1) {
                bb[0] = mu * delta[1] * nn[0] + mu * delta[0] * nn[1];
                bb[1] = lam * delta[0] * nn[0] + (lam + 2.0 * mu) * delta[1] * nn[1] + lam * delta[2] * nn[2];
                bb[2] = mu * delta[2] * nn[1] + mu * delta[1] * nn[2];
              } else {
                bb[0] = mu * delta[2] * nn[0] + mu * delta[0] * nn[2];
                bb[1] = mu * delta[2] * nn[1] + mu * delta[1] * nn[2];
                bb[2] = lam * delta[0] * nn[0] + lam * delta[1] * nn[1] + (lam + 2.0 * mu) * delta[2] * nn[2];
              }
            }
            for (int i = 0; i < basis.extent1(); ++i) {
              if (spaceDim == 2)
                res(e, off[i]) += ((-S(e, k, d, 0) * nn[0] - S(e, k, d, 1) * nn[1]) * basis(e, i, k, 0) + penalty * delta[d] * basis(e, i, k, 0) -
                                   modelparams[0] * (bb[0] * basis_grad(e, i, k, 0) + bb[1] * basis_grad(e, i, k, 1))) * wkset->wts_side(e, k);
              else
                res(e, off[i]) += ((-S(e, k, d, 0) * nn[0] - S(e, k, d, 1) * nn[1] - S(e, k, d, 2) * nn[2]) * basis(e, i, k, 0) + penalty * delta[d] * basis(e, i, k, 0) -
                                   modelparams[0] * (bb[0] * basis_grad(e, i, k, 0) + bb[1] * basis_grad(e, i, k, 1) + bb[2] * basis_grad(e, i, k, 2))) * wkset->wts_side(e, k);
            }
          }
      }
    }
  }
};

}  // namespace oracle
