// ORACLE (test infrastructure, never shipped, never on the product path).
//
// By-name module factory, restating PhysicsImporter<EvalT>::import
// (src/physics/physicsImporter.cpp:64-281) for the modules on the benchmarked path.
#pragma once
#include "assembly.hpp"
#include "physics_thermal.hpp"

namespace oracle {

template <class EvalT>
std::unique_ptr<PhysicsBase<EvalT>> import_physics(const std::string& name, const Settings& modset, int /*dim*/) {
  if (name == "thermal") return std::unique_ptr<PhysicsBase<EvalT>>(new thermal<EvalT>(modset));
  throw std::runtime_error("oracle: physics module not restated: " + name);
}

}  // namespace oracle
