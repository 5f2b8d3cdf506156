// Boundary-group pass of the thermal module (Neumann and weak-Dirichlet/Nitsche sides).
// Reference: assemblyManager_jacres.hpp:485-603 (boundary group loop), updateWorksetBoundary
// assemblyManager_workset.hpp:680-746, side data discretizationInterface_integration.hpp:592-774,
// thermal::boundaryResidual src/physics/thermal.cpp:171-281.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "plan.hpp"
#include "expr.hpp"
#include "kernel_abi.h"

namespace mrhyde_b200 {

constexpr int BND_MAXQ = 4;  // side cubature points (2-point Gauss on an edge / 2x2 on a face)
constexpr int BND_MAXV = 8;

struct BoundaryGroupHost {
  int sideset = 0, local_side = 0, nqp = 0;
  std::vector<int32_t> elem_ids;
  std::vector<double> pts, wts, val, grad;
  std::vector<std::vector<double>> bval, bgrad;   // per basis: values [card][nqp][vdim], gradients [card][nqp][dim] (HGRAD)
  double tu[3] = {0, 0, 0}, tv[3] = {0, 0, 0};
  std::string bctype = "none";
  ExprProgram data;
};

struct BoundaryGroupDev {  // one (sideset, local side) family, global memory
  int32_t bctype;          // 1 Neumann, 2 weak Dirichlet
  int32_t nqp;
  double wts[BND_MAXQ];
  double tu[3], tv[3];
  double gN[BND_MAXQ][BND_MAXV];        // geometry shape values at the side points
  double gdN[BND_MAXQ][BND_MAXV][3];
  double phi[BND_MAXQ][BND_MAXV];       // HGRAD values / reference gradients at the side points
  double dphi[BND_MAXQ][BND_MAXV][3];
  ExprProgram data;                     // Dirichlet / Neumann data at side ip
};

struct BoundarySetup {
  int dim = 3, nv = 8;
  double formparam = 1.0;
  ExprProgram diffusion;
};

struct BoundaryColour {
  int32_t first = 0, count = 0;  // range in the item arrays
};

struct BoundaryPlan {
  int dim = 3;
  double formparam = 1.0;
  ExprProgram diffusion;
  std::vector<BoundaryColour> groups;   // one launch per colour
  int32_t* d_item_elem = nullptr;       // [n_items]
  int32_t* d_item_group = nullptr;      // [n_items]
  BoundaryGroupDev* d_groups = nullptr;
  const double* vx = nullptr; const double* vy = nullptr; const double* vz = nullptr;
  const int32_t* conn = nullptr; const int32_t* lids = nullptr;
  ~BoundaryPlan();
};

void build_boundary_plan(const BoundarySetup& bs, const std::vector<BoundaryGroupHost>& groups, const MeshGraph& m, BoundaryPlan& out, size_t* dev_bytes);
void launch_boundary(const BoundaryPlan& B, const double* sol, const TimeDev& td, const GraphDev& G, const OutDev& O, void* stream, bool adjoint = false);

}  // namespace mrhyde_b200
