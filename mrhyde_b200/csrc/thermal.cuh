// Parameter block of the thermal Q1 volume kernel (thermal.cu).
#pragma once
#include <cstdint>

#include "device_plan.cuh"
#include "expr.hpp"

namespace mrhyde_b200 {

template <int DIM>
struct Q1Shape {
  static constexpr int NV = 1 << DIM;            // vertices == HGRAD C1 dofs
  static constexpr int NQ = 1 << DIM;            // 2-point Gauss per direction
  static constexpr int NT = NV * (NV + 1) / 2;   // upper triangle of the local matrix
  static constexpr int NG = DIM * (DIM + 1) / 2; // symmetric metric tensor entries
  static constexpr int STAGE = NT + NV;          // staged doubles per element: J upper triangle, then residual
};

constexpr int MAX_PREV = 4;   // BDF order <= 4 previous steps
constexpr int MAX_STAGE = 4;  // Butcher stages

struct TimeDev {  // computeSolnTransientSeeded coefficients (workset.cpp:600-834)
  int transient;
  int nprev, nstage_lo;          // previous steps used, stages below the current one
  double alpha_u, alpha_t;       // du/d(dof), du_t/d(dof)
  double one_minus_alpha_u;
  double timewt;                 // 1 / (dt b_s)
  double bdf[MAX_PREV + 1];      // BDF weights 1..nprev (index 0 unused)
  double stage_w[MAX_STAGE];     // A(s,s') / b(s')
  double time;                   // stage time
  const double* prev[MAX_PREV];
  const double* stg[MAX_STAGE];
};

template <int DIM>
struct ThermalTables {
  typedef Q1Shape<DIM> S;
  double gN[S::NQ][S::NV];          // geometry (Hex8/Quad4) shape values at the cubature points
  double gdN[S::NQ][S::NV][DIM];    // and reference gradients
  double phi[S::NQ][S::NV];         // HGRAD basis values  (setReferenceBasisData)
  double dphi[S::NQ][S::NV][DIM];   // HGRAD reference gradients
  double qw[S::NQ];
  double qpt[S::NQ][DIM];
  // constant-coefficient tables on parallelepipeds:
  //   K_ij = kappa |det| sum_{a<=b} G_ab Stab[ab][ij],  G = J^-1 J^-T
  //   M_ij = rho cp |det| Mtab[ij]
  double Stab[S::NG][S::NT];
  double Mtab[S::NT];
};

template <int DIM>
struct ThermalParams {
  ThermalTables<DIM> tab;
  ExprProgram source, diffusion, specific_heat, density;
  TimeDev td;
  int all_const;          // diffusion, specific heat, density are constants
  // mesh
  const double* vx; const double* vy; const double* vz;
  const int32_t* conn;    // [nelem][NV]
  const int32_t* lids;    // [nelem][NV]
  const uint8_t* affine;  // [nelem]
  const double* sol;
  PatchDev patches;
  GraphDev graph;
  OutDev out;
};

void launch_thermal_q1_2d(const ThermalParams<2>& P, int n_patches, int threads, size_t smem, void* stream);
void launch_thermal_q1_3d(const ThermalParams<3>& P, int n_patches, int threads, size_t smem, void* stream);
int thermal_q1_max_smem();

}  // namespace mrhyde_b200
