// Thermal module, HGRAD Q1 on Quad4 / Hex8: fused gather + geometry + push-forward + quadrature
// physics + derivative lanes + scatter, one CTA per patch of rows.
//
// Replaces, for one call of assembleJacRes (assemblyManager_jacres.hpp:336-603):
//   performGather                      assemblyManager_gather.hpp:181-234
//   computeSoln{Steady,Transient}Seeded workset.cpp:864-901, 600-834
//   evaluateSolutionField (T_t, grad(T)[x|y|z])  workset.cpp:978-1111
//   FunctionManager::evaluate ("thermal source", "thermal diffusion", "specific heat", "density")
//   thermal::volumeResidual            src/physics/thermal.cpp:70-165
//       res_i += (rho cp T_t - f) w phi_i + kappa grad(T).grad(phi_i) w
//   fused scatter + dofConstraints     assemblyManager_scatter.hpp:162-278, constraints.hpp:241-270
// The residual is linear in the seeded state, so the AD derivative lanes collapse to
//   dres_i/du_j = alpha_t rho cp w phi_i phi_j + alpha_u kappa w grad(phi_i).grad(phi_j)
// which is accumulated directly (36 upper-triangle entries in registers per element).
// Physical basis tables are never stored: the Jacobian and the HGRAD push-forward
// (discretizationInterface_basis.hpp:407-470) are recomputed from the vertex coordinates.
#include <cuda_runtime.h>

#include "expr_device.cuh"
#include "geometry.cuh"
#include "state_gather.cuh"
#include "thermal.cuh"

namespace mrhyde_b200 {

template <int DIM>
__device__ __forceinline__ void thermal_element(const ThermalParams<DIM>& P, const int e, const int le, const int n_pe, double* __restrict__ stage) {
  typedef Q1Shape<DIM> S;
  constexpr int NV = S::NV, NQ = S::NQ, NT = S::NT, NG = S::NG;
  const ThermalTables<DIM>& tab = P.tab;
  const TimeDev& td = P.td;

  int cn[NV], ld[NV];
  {
    const int4* c4 = reinterpret_cast<const int4*>(P.conn + (size_t)e * NV);
    const int4* l4 = reinterpret_cast<const int4*>(P.lids + (size_t)e * NV);
#pragma unroll
    for (int k = 0; k < NV / 4; ++k) {
      const int4 a = __ldg(c4 + k), b = __ldg(l4 + k);
      cn[4 * k] = a.x; cn[4 * k + 1] = a.y; cn[4 * k + 2] = a.z; cn[4 * k + 3] = a.w;
      ld[4 * k] = b.x; ld[4 * k + 1] = b.y; ld[4 * k + 2] = b.z; ld[4 * k + 3] = b.w;
    }
  }
  const bool affine = P.affine[e] != 0;

  // ---- gather + transient combination (u, u_t per dof)
  double u[NV], ut[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) gather_dof(P.sol, td, ld[i], u[i], ut[i]);

  double K[NT], r[NV];
#pragma unroll
  for (int t = 0; t < NT; ++t) K[t] = 0.0;
#pragma unroll
  for (int i = 0; i < NV; ++i) r[i] = 0.0;

  if (affine && P.all_const) {
    // ================= parallelepiped, constant coefficients: table path =================
    double X0[DIM], J[DIM][DIM], Ji[DIM][DIM];
    {
      const double* vc[3] = {P.vx, P.vy, P.vz};
      constexpr int nb[3] = {1, 3, 4};  // +xi, +eta, +zeta neighbours of vertex 0 (Shards order)
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        X0[d] = __ldg(vc[d] + cn[0]);
#pragma unroll
        for (int a = 0; a < DIM; ++a) J[d][a] = 0.5 * (__ldg(vc[d] + cn[nb[a]]) - X0[d]);
      }
    }
    const double det = det_inverse<DIM>(J, Ji);
    const double adet = fabs(det);
    double G[NG];
    {
      int g = 0;
#pragma unroll
      for (int a = 0; a < DIM; ++a) {
        double s = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) s += Ji[a][d] * Ji[a][d];
        G[g++] = s;
      }
#pragma unroll
      for (int a = 0; a < DIM; ++a)
#pragma unroll
        for (int b = a + 1; b < DIM; ++b) {
          double s = 0.0;
#pragma unroll
          for (int d = 0; d < DIM; ++d) s += Ji[a][d] * Ji[b][d];
          G[g++] = s;
        }
    }
    const double kd = P.diffusion.cval * adet;
#pragma unroll
    for (int g = 0; g < NG; ++g) G[g] *= kd;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      double s = 0.0;
#pragma unroll
      for (int g = 0; g < NG; ++g) s += G[g] * tab.Stab[g][t];
      K[t] = s;
    }
    // r = K u (symmetric), then the mass part
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = i; j < NV; ++j) {
        const double k = K[tri<NV>(i, j)];
        r[i] += k * u[j];
        if (j != i) r[j] += k * u[i];
      }
    if (td.transient) {
      const double md = P.density.cval * P.specific_heat.cval * adet;
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = i; j < NV; ++j) {
          const double mm = md * tab.Mtab[tri<NV>(i, j)];
          r[i] += mm * ut[j];
          if (j != i) r[j] += mm * ut[i];
          K[tri<NV>(i, j)] = td.alpha_u * K[tri<NV>(i, j)] + td.alpha_t * mm;
        }
    }
    // source
    if (P.source.is_const) {
      const double f = P.source.cval * adet;
#pragma unroll
      for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int i = 0; i < NV; ++i) r[i] -= f * tab.qw[q] * tab.phi[q][i];
    } else {
#pragma unroll 1
      for (int q = 0; q < NQ; ++q) {
        ExprVars in;
#pragma unroll
        for (int d = 0; d < 3; ++d) in.v[d] = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          double x = X0[d];
#pragma unroll
          for (int a = 0; a < DIM; ++a) x += J[d][a] * (tab.qpt[q][a] + 1.0);
          in.v[d] = x;
        }
        in.v[3] = td.time; in.v[4] = in.v[5] = in.v[6] = 0.0;
        const double fw = expr_eval_program(P.source, in) * tab.qw[q] * adet;
#pragma unroll
        for (int i = 0; i < NV; ++i) r[i] -= fw * tab.phi[q][i];
      }
    }
  } else {
    // ================= general path: per-point Jacobian and coefficients =================
    double X[NV][DIM];
    {
      const double* vc[3] = {P.vx, P.vy, P.vz};
#pragma unroll
      for (int n = 0; n < NV; ++n)
#pragma unroll
        for (int d = 0; d < DIM; ++d) X[n][d] = __ldg(vc[d] + cn[n]);
    }
#pragma unroll 1
    for (int q = 0; q < NQ; ++q) {
      double J[DIM][DIM], Ji[DIM][DIM];
      ExprVars in;
#pragma unroll
      for (int d = 0; d < 3; ++d) in.v[d] = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        double x = 0.0;
#pragma unroll
        for (int a = 0; a < DIM; ++a) J[d][a] = 0.0;
#pragma unroll
        for (int n = 0; n < NV; ++n) {
          x += tab.gN[q][n] * X[n][d];
#pragma unroll
          for (int a = 0; a < DIM; ++a) J[d][a] += X[n][d] * tab.gdN[q][n][a];
        }
        in.v[d] = x;
      }
      in.v[3] = td.time; in.v[4] = in.v[5] = in.v[6] = 0.0;
      const double det = det_inverse<DIM>(J, Ji);
      const double wd = fabs(det) * tab.qw[q];
      double g[NV][DIM];
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          double s = 0.0;
#pragma unroll
          for (int a = 0; a < DIM; ++a) s += Ji[a][d] * tab.dphi[q][i][a];
          g[i][d] = s;
        }
      const double kap = expr_eval(P.diffusion, in);
      const double f = expr_eval(P.source, in);
      double gT[DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NV; ++j) s += u[j] * g[j][d];
        gT[d] = s * kap * wd;
      }
      double lin = -f * wd, mw = 0.0;
      if (td.transient) {
        const double rc = expr_eval(P.density, in) * expr_eval(P.specific_heat, in);
        double Tt = 0.0;
#pragma unroll
        for (int j = 0; j < NV; ++j) Tt += ut[j] * tab.phi[q][j];
        lin += rc * Tt * wd;
        mw = td.alpha_t * rc * wd;
      }
      const double kw = td.alpha_u * kap * wd;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        double s = lin * tab.phi[q][i];
#pragma unroll
        for (int d = 0; d < DIM; ++d) s += gT[d] * g[i][d];
        r[i] += s;
        double gi[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) gi[d] = kw * g[i][d];
        const double mi = mw * tab.phi[q][i];
#pragma unroll
        for (int j = i; j < NV; ++j) {
          double k = mi * tab.phi[q][j];
#pragma unroll
          for (int d = 0; d < DIM; ++d) k += gi[d] * g[j][d];
          K[tri<NV>(i, j)] += k;
        }
      }
    }
  }

  // ---- stage: [entry][local element] keeps the writes bank-conflict free
#pragma unroll
  for (int t = 0; t < NT; ++t) stage[t * n_pe + le] = K[t];
#pragma unroll
  for (int i = 0; i < NV; ++i) stage[(NT + i) * n_pe + le] = r[i];
}

template <int DIM>
__global__ void __launch_bounds__(256) thermal_q1_volume_kernel(const __grid_constant__ ThermalParams<DIM> P) {
  extern __shared__ double stage[];
  const int patch = blockIdx.x;
  const int e0 = P.patches.patch_elem_ptr[patch];
  const int n_pe = P.patches.patch_elem_ptr[patch + 1] - e0;
  for (int le = threadIdx.x; le < n_pe; le += blockDim.x) {
    const int e = P.patches.patch_elems[e0 + le];
    thermal_element<DIM>(P, e, le, n_pe, stage);
  }
  __syncthreads();
  pull_scatter(P.patches, P.graph, P.out, stage, patch);
}

template <int DIM>
static void launch_impl(const ThermalParams<DIM>& P, int n_patches, int threads, size_t smem, void* stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(thermal_q1_volume_kernel<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, thermal_q1_max_smem());
    attr_set = true;
  }
  thermal_q1_volume_kernel<DIM><<<n_patches, threads, smem, (cudaStream_t)stream>>>(P);
}

void launch_thermal_q1_2d(const ThermalParams<2>& P, int n_patches, int threads, size_t smem, void* stream) { launch_impl<2>(P, n_patches, threads, smem, stream); }
void launch_thermal_q1_3d(const ThermalParams<3>& P, int n_patches, int threads, size_t smem, void* stream) { launch_impl<3>(P, n_patches, threads, smem, stream); }
int thermal_q1_max_smem() { return 200 * 1024; }

}  // namespace mrhyde_b200
