// Device views of the plan and the shared phase-2 "pull scatter" used by every volume kernel.
#pragma once
#include <cstdint>

#include "plan.hpp"

namespace mrhyde_b200 {

struct PatchDev {
  const int32_t* patch_elem_ptr;
  const int32_t* patch_elems;
  const int32_t* patch_row_ptr;
  const int32_t* patch_rows;
  const int32_t* patch_tmpl;
  const TemplateHeader* tmpl;
  const uint16_t* slot_row;
  const uint16_t* slot_k;
  const uint32_t* cptr;
  const uint16_t* csrc;
};

struct GraphDev {
  const int64_t* rowptr;
  const int32_t* colind;
  const uint8_t* fixed;
};

struct OutDev {
  double* res;       // may be null (compute_residual = 0)
  double* jac;       // may be null (compute_jacobian = 0)
  int accumulate;    // 1: += into caller-zeroed arrays (reference contract); 0: overwrite
};

// Phase 2: every CSR slot (and the residual entry) of the rows this patch owns sums its staged
// contributions in ascending element order and is written once.
//   res(row) (+)= -sum r_e[i]           (assemblyManager_scatter.hpp:227, sign convention -F)
//   J(row, col) (+)= sum dF_i/du_j      (:261-271)
//   fixed rows are skipped (:208, :253); in overwrite mode they receive the dofConstraints result
//   directly: J(d,d) = 1, rest of the row 0, res(d) = 0 (assemblyManager_constraints.hpp:125-138).
__device__ __forceinline__ void pull_scatter(const PatchDev& D, const GraphDev& G, const OutDev& O, const double* stage, int patch) {
  const TemplateHeader T = D.tmpl[D.patch_tmpl[patch]];
  const int32_t* rows = D.patch_rows + D.patch_row_ptr[patch];
  const uint16_t* srow = D.slot_row + T.off_slot;
  const uint16_t* sk = D.slot_k + T.off_slot;
  const uint32_t* cp = D.cptr + T.off_cptr;
  const uint16_t* cs = D.csrc + T.off_csrc;
  for (int s = threadIdx.x; s < T.n_slots; s += blockDim.x) {
    const uint32_t k = sk[s];
    const bool is_res = (k == SLOT_RES);
    if (is_res ? (O.res == nullptr) : (O.jac == nullptr)) continue;
    const int32_t row = rows[srow[s]];
    const uint32_t c0 = cp[s], c1 = cp[s + 1];
    double acc = 0.0;
    for (uint32_t c = c0; c < c1; ++c) acc += stage[cs[c]];
    const bool fixed = G.fixed[row] != 0;
    if (is_res) {
      if (!fixed) { if (O.accumulate) O.res[row] += -acc; else O.res[row] = -acc; }
      else if (!O.accumulate) O.res[row] = 0.0;
    } else {
      const int64_t p = G.rowptr[row] + k;
      if (!fixed) { if (O.accumulate) O.jac[p] += acc; else O.jac[p] = acc; }
      else if (!O.accumulate) O.jac[p] = (G.colind[p] == row) ? 1.0 : 0.0;
    }
  }
}

}  // namespace mrhyde_b200
