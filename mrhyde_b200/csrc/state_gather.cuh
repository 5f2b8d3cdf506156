// Gather of the element state with the transient stage/BDF combination folded in.
//   performGather / performGather4D          assemblyManager_gather.hpp:181-291
//   computeSolnSteadySeeded                  workset.cpp:864-901   (u = sol, du/ddof = 1)
//   computeSolnTransientSeeded (seedwhat 1)  workset.cpp:600-834
//       u   = alpha_u u_s + (1-alpha_u) u_prev0 + sum_{s'<s} A(s,s')/b(s') (u_stage[s'] - u_prev0)
//       u_t = alpha_t u_s + (sum_{k>=1} BDF(k) u_prev[k-1]) / (dt b(s))
#pragma once
#include "thermal.cuh"

namespace mrhyde_b200 {

__device__ __forceinline__ void gather_dof(const double* __restrict__ sol, const TimeDev& td, int lid, double& u, double& ut) {
  const double s = __ldg(sol + lid);
  u = s; ut = 0.0;
  if (td.transient) {
    const double p0 = __ldg(td.prev[0] + lid);
    double bu = td.one_minus_alpha_u * p0;
    for (int k = 0; k < td.nstage_lo; ++k) bu += td.stage_w[k] * (__ldg(td.stg[k] + lid) - p0);
    u = td.alpha_u * s + bu;
    double bt = td.bdf[1] * p0;
    for (int k = 2; k <= td.nprev; ++k) bt += td.bdf[k] * __ldg(td.prev[k - 1] + lid);
    bt *= td.timewt;
    ut = td.alpha_t * s + bt;
  }
}

}  // namespace mrhyde_b200
