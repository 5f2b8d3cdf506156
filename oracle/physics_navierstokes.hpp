// ORACLE (test infrastructure, never shipped, never on the product path).
//
// CPU restatement of the incompressible Navier-Stokes module (equal-order HGRAD velocity/pressure):
//   ctor / variable order ux, pr, uy, uz   src/physics/navierstokes.cpp:18-59
//   defineFunctions                        src/physics/navierstokes.cpp:64-76
//   volumeResidual  2-D                    src/physics/navierstokes.cpp:254-490
//                   3-D                    src/physics/navierstokes.cpp:493-848
//       [SURVEY 8(g) g1] the 3-D "Uz equation" block takes `off` from uy_num (navierstokes.cpp:688), so the
//       z-momentum Galerkin and SUPG terms land in the uy rows and the uz rows get no volume contribution.
//       `ns3d_uz_rows` = "reference" (default) reproduces this, "corrected" uses the uz offsets.
//   boundaryResidual (Neumann only)        src/physics/navierstokes.cpp:855-1013
//   computeTau                             src/physics/navierstokes.cpp:1053-1081   (g2: nvel branch at 1e-12)
// Not restated: 1-D, the energy coupling (variable "e").
#pragma once
#include "physics_base.hpp"

namespace oracle {

template <class EvalT>
class navierstokes : public PhysicsBase<EvalT> {
 public:
  using PhysicsBase<EvalT>::wkset;
  using PhysicsBase<EvalT>::functionManager;
  int spaceDim = 3;
  int ux_num = -1, pr_num = -1, uy_num = -1, uz_num = -1;
  bool useSUPG = false, usePSPG = false, uz_rows_reference = true;

  navierstokes(const Settings& settings, int dim) : spaceDim(dim) {
    this->label = "navierstokes";
    if (dim < 2) throw std::runtime_error("oracle: 1-D Navier-Stokes is not restated");
    this->myvars = {"ux", "pr", "uy"};
    this->mybasistypes = {"HGRAD", "HGRAD", "HGRAD"};
    if (dim == 3) { this->myvars.push_back("uz"); this->mybasistypes.push_back("HGRAD"); }
    useSUPG = settings.getb("useSUPG", false);
    usePSPG = settings.getb("usePSPG", false);
    uz_rows_reference = settings.get("ns3d_uz_rows", "reference") != "corrected";
  }

  void defineFunctions(const Settings& fs, FunctionManager<EvalT>* fm) override {
    functionManager = fm;
    fm->addFunction("source ux", fs.get("source ux", "0.0"), "ip");
    fm->addFunction("source pr", fs.get("source pr", "0.0"), "ip");
    fm->addFunction("source uy", fs.get("source uy", "0.0"), "ip");
    fm->addFunction("source uz", fs.get("source uz", "0.0"), "ip");
    fm->addFunction("density", fs.get("density", "1.0"), "ip");
    fm->addFunction("viscosity", fs.get("viscosity", "1.0"), "ip");
  }

  void setWorkset(Workset<EvalT>* w) override {
    wkset = w;
    ux_num = this->findVar("ux"); pr_num = this->findVar("pr"); uy_num = this->findVar("uy"); uz_num = this->findVar("uz");
  }

  EvalT computeTau(const EvalT& localdiff, const EvalT& xvl, const EvalT& yvl, const EvalT& zvl, double h, int dim, double dt, bool isTransient) const {
    using std::sqrt;
    const double C1 = 4.0, C2 = 2.0, C3 = isTransient ? 2.0 : 0.0;
    EvalT nvel = EvalT(0.0);
    if (dim == 2) nvel = xvl * xvl + yvl * yvl;
    else nvel = xvl * xvl + yvl * yvl + zvl * zvl;
    if (nvel > 1E-12) nvel = sqrt(nvel);
    EvalT tau = (C1 * localdiff / h / h) * (C1 * localdiff / h / h) + (C2 * nvel / h) * (C2 * nvel / h) + EvalT((C3 / dt) * (C3 / dt));
    tau = 1. / sqrt(tau);
    return tau;
  }

  void volumeResidual() override {
    const double dt = wkset->deltat;
    const bool isTransient = wkset->isTransient;
    Vista<EvalT> dens, visc, source[3], source_pr;
    source[0] = functionManager->evaluate("source ux", "ip");
    source_pr = functionManager->evaluate("source pr", "ip");
    source[1] = functionManager->evaluate("source uy", "ip");
    if (spaceDim > 2) source[2] = functionManager->evaluate("source uz", "ip");
    dens = functionManager->evaluate("density", "ip");
    visc = functionManager->evaluate("viscosity", "ip");
    auto& res = wkset->res;
    const int D = spaceDim;
    const int vnum[3] = {ux_num, uy_num, uz_num};
    const char* vn[3] = {"ux", "uy", "uz"};
    const char* cn[3] = {"[x]", "[y]", "[z]"};
    View2<EvalT>* u[3] = {nullptr, nullptr, nullptr};
    View2<EvalT>* ut[3] = {nullptr, nullptr, nullptr};
    View2<EvalT>* gu[3][3];
    View2<EvalT>* gp[3] = {nullptr, nullptr, nullptr};
    for (int a = 0; a < D; ++a) {
      u[a] = &wkset->getSolutionField(vn[a]);
      ut[a] = &wkset->getSolutionField(std::string(vn[a]) + "_t");
      for (int b = 0; b < D; ++b) gu[a][b] = &wkset->getSolutionField(std::string("grad(") + vn[a] + ")" + cn[b]);
      gp[a] = &wkset->getSolutionField(std::string("grad(pr)") + cn[a]);
    }
    auto& pr = wkset->getSolutionField("pr");
    std::vector<double> h;
    if (useSUPG || usePSPG) h = wkset->getElementSize();
    const EvalT zero = EvalT(0.0);

    // momentum equations
    for (int d = 0; d < D; ++d) {
      const int b = wkset->usebasis[vnum[d]];
      const View4& basis = wkset->basis[b];
      const View4& basis_grad = wkset->basis_grad[b];
      int offvar = vnum[d];
      if (D == 3 && d == 2 && uz_rows_reference) offvar = uy_num;  // navierstokes.cpp:688
      const auto& off = wkset->offsets[offvar];
      for (int elem = 0; elem < wkset->numElem; ++elem)
        for (int pt = 0; pt < basis.extent2(); ++pt) {
          const double w = wkset->wts(elem, pt);
          EvalT Fc[3];
          for (int c = 0; c < D; ++c) {
            Fc[c] = visc(elem, pt) * (*gu[d][c])(elem, pt);
            if (c == d) Fc[c] = Fc[c] - pr(elem, pt);
            Fc[c] *= w;
          }
          EvalT F = (*ut[d])(elem, pt) + (*u[0])(elem, pt) * (*gu[d][0])(elem, pt) + (*u[1])(elem, pt) * (*gu[d][1])(elem, pt);
          if (D == 3) F = F + (*u[2])(elem, pt) * (*gu[d][2])(elem, pt);
          F = F - source[d](elem, pt);
          F *= dens(elem, pt) * w;
          for (int dof = 0; dof < basis.extent1(); ++dof) {
            if (D == 2) res(elem, off[dof]) += Fc[0] * basis_grad(elem, dof, pt, 0) + Fc[1] * basis_grad(elem, dof, pt, 1) + F * basis(elem, dof, pt, 0);
            else res(elem, off[dof]) += Fc[0] * basis_grad(elem, dof, pt, 0) + Fc[1] * basis_grad(elem, dof, pt, 1) + Fc[2] * basis_grad(elem, dof, pt, 2) + F * basis(elem, dof, pt, 0);
          }
        }
      if (useSUPG) {
        for (int elem = 0; elem < wkset->numElem; ++elem)
          for (int pt = 0; pt < basis.extent2(); ++pt) {
            const double w = wkset->wts(elem, pt);
            EvalT tau = computeTau(visc(elem, pt), (*u[0])(elem, pt), (*u[1])(elem, pt), D == 3 ? (*u[2])(elem, pt) : zero, h[elem], D, dt, isTransient);
            EvalT conv = (*u[0])(elem, pt) * (*gu[d][0])(elem, pt) + (*u[1])(elem, pt) * (*gu[d][1])(elem, pt);
            if (D == 3) conv = conv + (*u[2])(elem, pt) * (*gu[d][2])(elem, pt);
            EvalT stabres = dens(elem, pt) * (*ut[d])(elem, pt) + dens(elem, pt) * conv + (*gp[d])(elem, pt) - dens(elem, pt) * source[d](elem, pt);
            EvalT Sc[3];
            for (int c = 0; c < D; ++c) Sc[c] = tau * stabres * (*u[c])(elem, pt) * w;
            for (int dof = 0; dof < basis.extent1(); ++dof) {
              if (D == 2) res(elem, off[dof]) += Sc[0] * basis_grad(elem, dof, pt, 0) + Sc[1] * basis_grad(elem, dof, pt, 1);
              else res(elem, off[dof]) += Sc[0] * basis_grad(elem, dof, pt, 0) + Sc[1] * basis_grad(elem, dof, pt, 1) + Sc[2] * basis_grad(elem, dof, pt, 2);
            }
          }
      }
    }
    // pressure (continuity) equation
    {
      const int b = wkset->usebasis[pr_num];
      const View4& basis = wkset->basis[b];
      const View4& basis_grad = wkset->basis_grad[b];
      const auto& off = wkset->offsets[pr_num];
      for (int elem = 0; elem < wkset->numElem; ++elem)
        for (int pt = 0; pt < basis.extent2(); ++pt) {
          EvalT divu = (*gu[0][0])(elem, pt) + (*gu[1][1])(elem, pt);
          if (D == 3) divu = divu + (*gu[2][2])(elem, pt);
          divu = divu * wkset->wts(elem, pt);
          for (int dof = 0; dof < basis.extent1(); ++dof) res(elem, off[dof]) += divu * basis(elem, dof, pt, 0);
        }
      if (usePSPG) {
        for (int elem = 0; elem < wkset->numElem; ++elem)
          for (int pt = 0; pt < basis.extent2(); ++pt) {
            const double w = wkset->wts(elem, pt);
            EvalT tau = computeTau(visc(elem, pt), (*u[0])(elem, pt), (*u[1])(elem, pt), D == 3 ? (*u[2])(elem, pt) : zero, h[elem], D, dt, isTransient);
            EvalT Sc[3];
            for (int d = 0; d < D; ++d) {
              EvalT conv = (*u[0])(elem, pt) * (*gu[d][0])(elem, pt) + (*u[1])(elem, pt) * (*gu[d][1])(elem, pt);
              if (D == 3) conv = conv + (*u[2])(elem, pt) * (*gu[d][2])(elem, pt);
              Sc[d] = dens(elem, pt) * (*ut[d])(elem, pt) + dens(elem, pt) * conv + (*gp[d])(elem, pt) - dens(elem, pt) * source[d](elem, pt);
              Sc[d] *= tau * w / dens(elem, pt);
            }
            for (int dof = 0; dof < basis.extent1(); ++dof) {
              if (D == 2) res(elem, off[dof]) += Sc[0] * basis_grad(elem, dof, pt, 0) + Sc[1] * basis_grad(elem, dof, pt, 1);
              else res(elem, off[dof]) += Sc[0] * basis_grad(elem, dof, pt, 0) + Sc[1] * basis_grad(elem, dof, pt, 1) + Sc[2] * basis_grad(elem, dof, pt, 2);
            }
          }
      }
    }
  }

  void boundaryResidual() override {
    const int cside = wkset->currentside;
    const int vnum[3] = {ux_num, uy_num, uz_num};
    const char* vn[3] = {"ux", "uy", "uz"};
    auto& res = wkset->res;
    for (int d = 0; d < spaceDim; ++d) {
      if (wkset->var_bcs[vnum[d]][cside] != "Neumann") continue;
      Vista<EvalT> src = functionManager->evaluate(std::string("Neumann ") + vn[d] + " " + wkset->sidename, "side ip");
      const View4& basis = wkset->basis_side[wkset->usebasis[vnum[d]]];
      const auto& off = wkset->offsets[vnum[d]];
      for (int e = 0; e < wkset->numElem; ++e)
        for (int k = 0; k < basis.extent2(); ++k)
          for (int i = 0; i < basis.extent1(); ++i) res(e, off[i]) += (-src(e, k) * basis(e, i, k, 0)) * wkset->wts_side(e, k);
    }
  }
};

}  // namespace oracle
