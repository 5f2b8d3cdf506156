// ORACLE (test infrastructure, never shipped, never on the product path).
//
// CPU restatement of the Maxwell module (3-D mixed form, E in HCURL, B in HDIV):
//   ctor ("active variables", "use leap frog")   src/physics/maxwell.cpp:18-58
//   defineFunctions                               src/physics/maxwell.cpp:63-78
//   volumeResidual 3-D                            src/physics/maxwell.cpp:138-209 (B equation), :262-303 (E equation)
//       B:  (B_t + curl E) . psi                  leap-frog: curl E only in stage 0
//       E:  (n^2 E_t + (sigma E + J)/eps) . phi - B/(mu eps) . curl phi      [SURVEY 8(g) g12]; leap-frog: stage 1 only
//   boundaryResidual 3-D (ABC on "Neumann" B sides, gamma = -0.9944)   src/physics/maxwell.cpp:313-403
// Not restated: the 2-D (HVOL B) branch.
#pragma once
#include "physics_base.hpp"

namespace oracle {

template <class EvalT>
class maxwell : public PhysicsBase<EvalT> {
 public:
  using PhysicsBase<EvalT>::wkset;
  using PhysicsBase<EvalT>::functionManager;
  int spaceDim = 3, Enum = -1, Bnum = -1;
  bool include_Beqn = true, include_Eeqn = true, useLeapFrog = false;

  maxwell(const Settings& settings, int dim) : spaceDim(dim) {
    this->label = "maxwell";
    if (dim != 3) throw std::runtime_error("oracle: only the 3-D Maxwell branch is restated");
    if (settings.has("active variables")) {
      const std::string a = settings.get("active variables", "");
      include_Eeqn = a.find("E") != std::string::npos;
      include_Beqn = a.find("B") != std::string::npos;
    }
    if (include_Eeqn) { this->myvars.push_back("E"); this->mybasistypes.push_back("HCURL"); }
    if (include_Beqn) { this->myvars.push_back("B"); this->mybasistypes.push_back("HDIV"); }
    useLeapFrog = settings.getb("use leap frog", false);
  }

  void defineFunctions(const Settings& fs, FunctionManager<EvalT>* fm) override {
    functionManager = fm;
    fm->addFunction("current x", fs.get("current x", "0.0"), "ip");
    fm->addFunction("current y", fs.get("current y", "0.0"), "ip");
    fm->addFunction("current z", fs.get("current z", "0.0"), "ip");
    fm->addFunction("mu", fs.get("permeability", "1.0"), "ip");
    fm->addFunction("refractive index", fs.get("refractive index", "1.0"), "ip");
    fm->addFunction("epsilon", fs.get("permittivity", "1.0"), "ip");
    fm->addFunction("sigma", fs.get("conductivity", "0.0"), "ip");
    fm->addFunction("epsilon", fs.get("permittivity", "1.0"), "side ip");
  }

  void setWorkset(Workset<EvalT>* w) override {
    wkset = w;
    Enum = this->findVar("E"); Bnum = this->findVar("B");
  }

  void volumeResidual() override {
    const int stage = wkset->current_stage;
    auto& res = wkset->res;
    if (include_Beqn) {
      const View4& basis = wkset->basis[wkset->usebasis[Bnum]];
      const auto& off = wkset->offsets[Bnum];
      auto& dBx_dt = wkset->getSolutionField("B_t[x]");
      auto& dBy_dt = wkset->getSolutionField("B_t[y]");
      auto& dBz_dt = wkset->getSolutionField("B_t[z]");
      const bool with_curl = !useLeapFrog || stage == 0;
      if (with_curl) {
        auto& curlE_x = wkset->getSolutionField("curl(E)[x]");
        auto& curlE_y = wkset->getSolutionField("curl(E)[y]");
        auto& curlE_z = wkset->getSolutionField("curl(E)[z]");
        for (int elem = 0; elem < wkset->numElem; ++elem)
          for (int pt = 0; pt < basis.extent2(); ++pt) {
            EvalT f0 = (dBx_dt(elem, pt) + curlE_x(elem, pt)) * wkset->wts(elem, pt);
            EvalT f1 = (dBy_dt(elem, pt) + curlE_y(elem, pt)) * wkset->wts(elem, pt);
            EvalT f2 = (dBz_dt(elem, pt) + curlE_z(elem, pt)) * wkset->wts(elem, pt);
            for (int dof = 0; dof < basis.extent1(); ++dof) {
              res(elem, off[dof]) += f0 * basis(elem, dof, pt, 0);
              res(elem, off[dof]) += f1 * basis(elem, dof, pt, 1);
              res(elem, off[dof]) += f2 * basis(elem, dof, pt, 2);
            }
          }
      } else {
        for (int elem = 0; elem < wkset->numElem; ++elem)
          for (int pt = 0; pt < basis.extent2(); ++pt) {
            EvalT f0 = dBx_dt(elem, pt) * wkset->wts(elem, pt);
            EvalT f1 = dBy_dt(elem, pt) * wkset->wts(elem, pt);
            EvalT f2 = dBz_dt(elem, pt) * wkset->wts(elem, pt);
            for (int dof = 0; dof < basis.extent1(); ++dof) {
              res(elem, off[dof]) += f0 * basis(elem, dof, pt, 0);
              res(elem, off[dof]) += f1 * basis(elem, dof, pt, 1);
              res(elem, off[dof]) += f2 * basis(elem, dof, pt, 2);
            }
          }
      }
    }
    if (include_Eeqn) {
      Vista<EvalT> mu, epsilon, sigma, rindex, current_x, current_y, current_z;
      current_x = functionManager->evaluate("current x", "ip");
      current_y = functionManager->evaluate("current y", "ip");
      current_z = functionManager->evaluate("current z", "ip");
      mu = functionManager->evaluate("mu", "ip");
      epsilon = functionManager->evaluate("epsilon", "ip");
      rindex = functionManager->evaluate("refractive index", "ip");
      sigma = functionManager->evaluate("sigma", "ip");
      if (!useLeapFrog || stage == 1) {
        const int eb = wkset->usebasis[Enum];
        const View4& basis = wkset->basis[eb];
        const View4& basis_curl = wkset->basis_curl[eb];
        auto& dEx_dt = wkset->getSolutionField("E_t[x]");
        auto& dEy_dt = wkset->getSolutionField("E_t[y]");
        auto& dEz_dt = wkset->getSolutionField("E_t[z]");
        auto& Bx = wkset->getSolutionField("B[x]");
        auto& By = wkset->getSolutionField("B[y]");
        auto& Bz = wkset->getSolutionField("B[z]");
        auto& Ex = wkset->getSolutionField("E[x]");
        auto& Ey = wkset->getSolutionField("E[y]");
        auto& Ez = wkset->getSolutionField("E[z]");
        const auto& off = wkset->offsets[Enum];
        for (int elem = 0; elem < wkset->numElem; ++elem)
          for (int pt = 0; pt < basis.extent2(); ++pt) {
            const double w = wkset->wts(elem, pt);
            EvalT eps = epsilon(elem, pt);
            EvalT f0 = (rindex(elem, pt) * rindex(elem, pt) * dEx_dt(elem, pt) + 1.0 / eps * (sigma(elem, pt) * Ex(elem, pt) + current_x(elem, pt))) * w;
            EvalT f1 = (rindex(elem, pt) * rindex(elem, pt) * dEy_dt(elem, pt) + 1.0 / eps * (sigma(elem, pt) * Ey(elem, pt) + current_y(elem, pt))) * w;
            EvalT f2 = (rindex(elem, pt) * rindex(elem, pt) * dEz_dt(elem, pt) + 1.0 / eps * (sigma(elem, pt) * Ez(elem, pt) + current_z(elem, pt))) * w;
            EvalT c0 = -1.0 / mu(elem, pt) * 1.0 / eps * Bx(elem, pt) * w;
            EvalT c1 = -1.0 / mu(elem, pt) * 1.0 / eps * By(elem, pt) * w;
            EvalT c2 = -1.0 / mu(elem, pt) * 1.0 / eps * Bz(elem, pt) * w;
            for (int dof = 0; dof < basis.extent1(); ++dof) {
              res(elem, off[dof]) += f0 * basis(elem, dof, pt, 0) + c0 * basis_curl(elem, dof, pt, 0);
              res(elem, off[dof]) += f1 * basis(elem, dof, pt, 1) + c1 * basis_curl(elem, dof, pt, 1);
              res(elem, off[dof]) += f2 * basis(elem, dof, pt, 2) + c2 * basis_curl(elem, dof, pt, 2);
            }
          }
      }
    }
  }

  void boundaryResidual() override {
    const int cside = wkset->currentside;
    const double gamma = -0.9944;
    if (!(include_Beqn && wkset->var_bcs[Bnum][cside] == "Neumann")) return;  // "really ABC"
    auto& nx = wkset->getScalarField("n[x]");
    auto& ny = wkset->getScalarField("n[y]");
    auto& nz = wkset->getScalarField("n[z]");
    auto& Ex = wkset->getSolutionField("E[x]");
    auto& Ey = wkset->getSolutionField("E[y]");
    auto& Ez = wkset->getSolutionField("E[z]");
    const auto& off = wkset->offsets[Bnum];
    const View4& basis = wkset->basis_side[wkset->usebasis[Bnum]];
    auto& res = wkset->res;
    for (int elem = 0; elem < wkset->numElem; ++elem)
      for (int pt = 0; pt < basis.extent2(); ++pt) {
        const double w = wkset->wts_side(elem, pt);
        EvalT nce_x = ny(elem, pt) * Ez(elem, pt) - nz(elem, pt) * Ey(elem, pt);
        EvalT nce_y = nz(elem, pt) * Ex(elem, pt) - nx(elem, pt) * Ez(elem, pt);
        EvalT nce_z = nx(elem, pt) * Ey(elem, pt) - ny(elem, pt) * Ex(elem, pt);
        EvalT c0 = -(1.0 + gamma) * (ny(elem, pt) * nce_z - nz(elem, pt) * nce_y) * w;
        EvalT c1 = -(1.0 + gamma) * (nz(elem, pt) * nce_x - nx(elem, pt) * nce_z) * w;
        EvalT c2 = -(1.0 + gamma) * (nx(elem, pt) * nce_y - ny(elem, pt) * nce_x) * w;
        for (int dof = 0; dof < basis.extent1(); ++dof)
          res(elem, off[dof]) += c0 * basis(elem, dof, pt, 0) + c1 * basis(elem, dof, pt, 1) + c2 * basis(elem, dof, pt, 2);
      }
  }
};

}  // namespace oracle
