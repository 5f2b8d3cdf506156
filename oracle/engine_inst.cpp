// ORACLE (test infrastructure, never shipped, never on the product path).
// Explicit instantiation of the assembly engine for one evaluation type (-DORACLE_T=...),
// mirroring the reference's per-type explicit instantiations (e.g. src/physics/thermal.cpp tail).
#define ORACLE_INSTANTIATE
#include "assembly.hpp"
#include "physics_all.hpp"
namespace oracle {
template std::unique_ptr<EngineBase> make_engine<ORACLE_T>(AssemblyManager&);
}
