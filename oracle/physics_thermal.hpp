// ORACLE (test infrastructure, never shipped, never on the product path).
//
// CPU restatement of the thermal module:
//   ctor / variable list          src/physics/thermal.cpp:17-41
//   defineFunctions               src/physics/thermal.cpp:47-65
//   volumeResidual                src/physics/thermal.cpp:70-165   (dof-outer, qp-inner, separate += per term)
//   boundaryResidual              src/physics/thermal.cpp:171-281  (Neumann; weak Dirichlet = Nitsche, epen = 10)
#pragma once
#include "physics_base.hpp"

namespace oracle {

template <class EvalT>
class thermal : public PhysicsBase<EvalT> {
 public:
  using PhysicsBase<EvalT>::wkset;
  using PhysicsBase<EvalT>::functionManager;
  int T_num = -1, T_basis_num = -1;
  bool have_nsvel = false;
  double formparam = 1.0;
  bool have_advection = false;

  thermal(const Settings& settings) {
    this->label = "thermal";
    this->myvars.push_back("T");
    this->mybasistypes.push_back("HGRAD");
    formparam = settings.getd("form_param", 1.0);
    have_advection = settings.getb("include advection", false);
  }

  void defineFunctions(const Settings& fs, FunctionManager<EvalT>* fm) override {
    functionManager = fm;
    fm->addFunction("thermal source", fs.get("thermal source", "0.0"), "ip");
    fm->addFunction("thermal diffusion", fs.get("thermal diffusion", "1.0"), "ip");
    fm->addFunction("specific heat", fs.get("specific heat", "1.0"), "ip");
    fm->addFunction("density", fs.get("density", "1.0"), "ip");
    fm->addFunction("bx", fs.get("advection x", "0.0"), "ip");
    fm->addFunction("by", fs.get("advection y", "0.0"), "ip");
    fm->addFunction("bz", fs.get("advection z", "0.0"), "ip");
    fm->addFunction("thermal diffusion", fs.get("thermal diffusion", "1.0"), "side ip");
    fm->addFunction("robin alpha", fs.get("robin alpha", "0.0"), "side ip");
  }

  void setWorkset(Workset<EvalT>* w) override {
    wkset = w;
    T_num = this->findVar("T");
    T_basis_num = wkset->usebasis[T_num];
    have_nsvel = this->findVar("ux") >= 0;   // thermal.cpp:362-379: a Navier-Stokes velocity in the block advects the temperature
  }

  void volumeResidual() override {
    const int spaceDim = wkset->dimension;
    const View4& basis = wkset->basis[T_basis_num];
    const View4& basis_grad = wkset->basis_grad[T_basis_num];
    Vista<EvalT> source, diff, cp, rho, bx, by, bz;
    source = functionManager->evaluate("thermal source", "ip");
    diff = functionManager->evaluate("thermal diffusion", "ip");
    cp = functionManager->evaluate("specific heat", "ip");
    rho = functionManager->evaluate("density", "ip");
    if (have_advection) {
      bx = functionManager->evaluate("bx", "ip");
      if (spaceDim > 1) by = functionManager->evaluate("by", "ip");
      if (spaceDim > 2) bz = functionManager->evaluate("bz", "ip");
    }
    auto& res = wkset->res;
    auto& dTdt = wkset->getSolutionField("T_t");
    const auto& off = wkset->offsets[T_num];
    auto& dTdx = wkset->getSolutionField("grad(T)[x]");
    auto& dTdy = wkset->getSolutionField("grad(T)[y]");
    auto& dTdz = wkset->getSolutionField("grad(T)[z]");
    for (int elem = 0; elem < wkset->numElem; ++elem) {
      for (int dof = 0; dof < basis.extent1(); ++dof) {
        for (int pt = 0; pt < basis.extent2(); ++pt) {
          const double w = wkset->wts(elem, pt);
          res(elem, off[dof]) += (rho(elem, pt) * cp(elem, pt) * dTdt(elem, pt) - source(elem, pt)) * w * basis(elem, dof, pt, 0);
          res(elem, off[dof]) += diff(elem, pt) * dTdx(elem, pt) * w * basis_grad(elem, dof, pt, 0);
          if (spaceDim > 1) res(elem, off[dof]) += diff(elem, pt) * dTdy(elem, pt) * w * basis_grad(elem, dof, pt, 1);
          if (spaceDim > 2) res(elem, off[dof]) += diff(elem, pt) * dTdz(elem, pt) * w * basis_grad(elem, dof, pt, 2);
          if (have_nsvel) {   // thermal.cpp:139-149
            auto& Ux = wkset->getSolutionField("ux");
            if (spaceDim == 1) res(elem, off[dof]) += Ux(elem, pt) * dTdx(elem, pt) * w * basis(elem, dof, pt, 0);
            else if (spaceDim == 2) res(elem, off[dof]) += (Ux(elem, pt) * dTdx(elem, pt) + wkset->getSolutionField("uy")(elem, pt) * dTdy(elem, pt)) * w * basis(elem, dof, pt, 0);
            else res(elem, off[dof]) += (Ux(elem, pt) * dTdx(elem, pt) + wkset->getSolutionField("uy")(elem, pt) * dTdy(elem, pt) + wkset->getSolutionField("uz")(elem, pt) * dTdz(elem, pt)) * w * basis(elem, dof, pt, 0);
          }
          if (have_advection) {
            if (spaceDim == 1) res(elem, off[dof]) += bx(elem, pt) * dTdx(elem, pt) * w * basis(elem, dof, pt, 0);
            else if (spaceDim == 2) res(elem, off[dof]) += (bx(elem, pt) * dTdx(elem, pt) + by(elem, pt) * dTdy(elem, pt)) * w * basis(elem, dof, pt, 0);
            else res(elem, off[dof]) += (bx(elem, pt) * dTdx(elem, pt) + by(elem, pt) * dTdy(elem, pt) + bz(elem, pt) * dTdz(elem, pt)) * w * basis(elem, dof, pt, 0);
          }
        }
      }
    }
  }

  void boundaryResidual() override {
    const int cside = wkset->currentside;
    const std::string bctype = wkset->var_bcs[T_num][cside];
    const View4& basis = wkset->basis_side[T_basis_num];
    const View4& basis_grad = wkset->basis_grad_side[T_basis_num];
    Vista<EvalT> nsource, diff_side, robin_alpha;
    if (bctype == "weak Dirichlet") nsource = functionManager->evaluate("Dirichlet T " + wkset->sidename, "side ip");
    else if (bctype == "Neumann") nsource = functionManager->evaluate("Neumann T " + wkset->sidename, "side ip");
    diff_side = functionManager->evaluate("thermal diffusion", "side ip");
    robin_alpha = functionManager->evaluate("robin alpha", "side ip");
    const double sf = wkset->isAdjoint ? 1.0 : formparam;   // thermal.cpp:197-201
    auto h = wkset->getSideElementSize();
    auto& res = wkset->res;
    const auto& off = wkset->offsets[T_num];
    const int dim = wkset->dimension;
    if (bctype == "Neumann") {
      for (int elem = 0; elem < wkset->numElem; ++elem)
        for (int dof = 0; dof < basis.extent1(); ++dof)
          for (int pt = 0; pt < basis.extent2(); ++pt)
            res(elem, off[dof]) += -nsource(elem, pt) * wkset->wts_side(elem, pt) * basis(elem, dof, pt, 0);
    } else if (bctype == "weak Dirichlet") {
      auto& T = wkset->getSolutionField("T");
      auto& dTdx = wkset->getSolutionField("grad(T)[x]");
      auto& dTdy = wkset->getSolutionField("grad(T)[y]");
      auto& dTdz = wkset->getSolutionField("grad(T)[z]");
      auto& nx = wkset->getScalarField("n[x]");
      auto& ny = wkset->getScalarField("n[y]");
      auto& nz = wkset->getScalarField("n[z]");
      const Vista<EvalT>& bdata = nsource;
      const double epen = 10.0;
      for (int elem = 0; elem < wkset->numElem; ++elem)
        for (int dof = 0; dof < basis.extent1(); ++dof)
          for (int pt = 0; pt < basis.extent2(); ++pt) {
            const double w = wkset->wts_side(elem, pt);
            res(elem, off[dof]) += epen / h[elem] * diff_side(elem, pt) * (T(elem, pt) - bdata(elem, pt)) * w * basis(elem, dof, pt, 0);
            if (dim == 2) {
              res(elem, off[dof]) += -diff_side(elem, pt) * (dTdx(elem, pt) * nx(elem, pt) + dTdy(elem, pt) * ny(elem, pt)) * w * basis(elem, dof, pt, 0);
              res(elem, off[dof]) += -sf * diff_side(elem, pt) * (T(elem, pt) - bdata(elem, pt)) * w *
                                     (basis_grad(elem, dof, pt, 0) * nx(elem, pt) + basis_grad(elem, dof, pt, 1) * ny(elem, pt));
            } else {
              res(elem, off[dof]) += -diff_side(elem, pt) * (dTdx(elem, pt) * nx(elem, pt) + dTdy(elem, pt) * ny(elem, pt) + dTdz(elem, pt) * nz(elem, pt)) * w * basis(elem, dof, pt, 0);
              res(elem, off[dof]) += -sf * diff_side(elem, pt) * (T(elem, pt) - bdata(elem, pt)) * w *
                                     (basis_grad(elem, dof, pt, 0) * nx(elem, pt) + basis_grad(elem, dof, pt, 1) * ny(elem, pt) + basis_grad(elem, dof, pt, 2) * nz(elem, pt));
            }
          }
    }
  }
};

}  // namespace oracle
