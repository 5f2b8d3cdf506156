// Cell-geometry helpers shared by the kernels: Jacobian determinant / inverse
// (CellTools::setJacobianDet / setJacobianInv as called at discretizationInterface_basis.hpp:407-413)
// and the upper-triangle index of a symmetric local matrix.
#pragma once

namespace mrhyde_b200 {

template <int NV>
__host__ __device__ constexpr int tri(int i, int j) { return i * NV - (i * (i - 1)) / 2 + (j - i); }

template <int DIM>
__device__ __forceinline__ double det_inverse(const double (&J)[DIM][DIM], double (&Ji)[DIM][DIM]) {
  if constexpr (DIM == 2) {
    const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double id = 1.0 / det;
    Ji[0][0] = J[1][1] * id; Ji[0][1] = -J[0][1] * id; Ji[1][0] = -J[1][0] * id; Ji[1][1] = J[0][0] * id;
    return det;
  } else {
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    const double id = 1.0 / det;
    Ji[0][0] = c00 * id; Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id; Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    Ji[1][0] = c01 * id; Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id; Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    Ji[2][0] = c02 * id; Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id; Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
    return det;
  }
}

}  // namespace mrhyde_b200
