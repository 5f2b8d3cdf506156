// ORACLE (test infrastructure, never shipped, never on the product path).
//
// By-name module factory, restating PhysicsImporter<EvalT>::import
// (src/physics/physicsImporter.cpp:64-281) for the modules on the benchmarked path.
#pragma once
#include "assembly.hpp"
#include "physics_linearelasticity.hpp"
#include "physics_maxwell.hpp"
#include "physics_navierstokes.hpp"
#include "physics_thermal.hpp"

namespace oracle {

template <class EvalT>
std::unique_ptr<PhysicsBase<EvalT>> import_physics(const std::string& name, const Settings& modset, int dim) {
  if (name == "thermal") return std::unique_ptr<PhysicsBase<EvalT>>(new thermal<EvalT>(modset));
  if (name == "linearelasticity" || name == "linear elasticity") return std::unique_ptr<PhysicsBase<EvalT>>(new linearelasticity<EvalT>(modset, dim));
  if (name == "navier stokes" || name == "Navier Stokes") return std::unique_ptr<PhysicsBase<EvalT>>(new navierstokes<EvalT>(modset, dim));
  if (name == "maxwell") return std::unique_ptr<PhysicsBase<EvalT>>(new maxwell<EvalT>(modset, dim));
  throw std::runtime_error("oracle: physics module not restated: " + name);
}

}  // namespace oracle
