// Volume assembly kernel of the thermal module, HGRAD Q1 on Quad4 / Hex8: one CTA walks one CHAIN of
// the sweep plan (plan.hpp).  Per step: (1) every thread computes one element -- gather, cell geometry,
// push-forward, coefficient functions, quadrature, derivative lanes -- and stages its local matrix +
// residual in the shared-memory ring; (2) every warp finalises rows: each lane sums the staged
// contributions of one CSR entry in ascending element order and stores it once.
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
// which is accumulated directly (upper triangle).  Physical basis tables are never stored: the Jacobian and
// the HGRAD push-forward (discretizationInterface_basis.hpp:407-470) are recomputed from the vertices.
//
// This file is compiled twice: ahead of time by nvcc (coefficient functions run through the bytecode
// interpreter, tables come from the kernel parameters) and, per plan, by NVRTC with MRH_JIT defined
// (jit.cpp): then the plan's `Functions:` expressions are generated C++ (mrh_fn_*), the reference tables are
// constexpr arrays (jit_tab::*) and the quadrature loops are fully unrolled so the compiler folds zeros and
// shares sub-expressions between quadrature points.
#pragma once
#ifndef __CUDACC_RTC__
#include "kernel_abi.h"
#endif

namespace mrhyde_b200 {

#ifdef MRH_JIT
#define MRH_TRANSIENT(td) (MRH_JIT_TRANSIENT != 0)
#define MRH_TAB(f) jit_tab::f
#define MRH_UNROLL_Q _Pragma("unroll")
#else
#define MRH_TRANSIENT(td) ((td).transient != 0)
#define MRH_TAB(f) P.tab.f
#define MRH_UNROLL_Q _Pragma("unroll 1")
#endif

template <int NV>
__host__ __device__ constexpr int tri(int i, int j) { return i * NV - (i * (i - 1)) / 2 + (j - i); }

// CellTools::setJacobianDet / setJacobianInv as called at discretizationInterface_basis.hpp:407-413
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

// ---------------------------------------------------------------------------------------------------------
// Coefficient functions at a point
// ---------------------------------------------------------------------------------------------------------
#ifndef MRH_JIT
struct ExprVars {  // workset scalar fields at one point: x y z t n[x] n[y] n[z]
  double v[7];
};

__device__ __forceinline__ double expr_var(const ExprVars& in, int i) {
  double r = in.v[0];
  r = (i == 1) ? in.v[1] : r;
  r = (i == 2) ? in.v[2] : r;
  r = (i == 3) ? in.v[3] : r;
  r = (i == 4) ? in.v[4] : r;
  r = (i == 5) ? in.v[5] : r;
  r = (i == 6) ? in.v[6] : r;
  return r;
}

// One thread evaluates one point; the program lives in kernel-parameter constant memory so the opcode fetch
// is a uniform constant load and the switch is a warp-uniform branch.
static __device__ __noinline__ double expr_eval_program(const ExprProgram& p, const ExprVars& in) {
  double st[EXPR_MAXSTACK];
  int sp = 0;        // number of values below the top-of-stack register
  double a = 0.0;    // top of stack
  const int n = p.n;
  for (int i = 0; i < n; ++i) {
    const double c = p.c[i];
    switch (p.op[i]) {
      case OP_PUSHC: st[sp & (EXPR_MAXSTACK - 1)] = a; ++sp; a = c; break;
      case OP_PUSHV: st[sp & (EXPR_MAXSTACK - 1)] = a; ++sp; a = expr_var(in, (int)c); break;
      case OP_ADD: --sp; a = st[sp & (EXPR_MAXSTACK - 1)] + a; break;
      case OP_SUB: --sp; a = st[sp & (EXPR_MAXSTACK - 1)] + (-a); break;
      case OP_MUL: --sp; a = st[sp & (EXPR_MAXSTACK - 1)] * a; break;
      case OP_DIV: --sp; a = st[sp & (EXPR_MAXSTACK - 1)] / a; break;
      case OP_POW: --sp; a = pow(st[sp & (EXPR_MAXSTACK - 1)], a); break;
      case OP_LT: --sp; a = st[sp & (EXPR_MAXSTACK - 1)] < a ? 1.0 : 0.0; break;
      case OP_LTE: --sp; a = st[sp & (EXPR_MAXSTACK - 1)] <= a ? 1.0 : 0.0; break;
      case OP_GT: --sp; a = st[sp & (EXPR_MAXSTACK - 1)] > a ? 1.0 : 0.0; break;
      case OP_GTE: --sp; a = st[sp & (EXPR_MAXSTACK - 1)] >= a ? 1.0 : 0.0; break;
      case OP_MAX: { --sp; const double l = st[sp & (EXPR_MAXSTACK - 1)]; a = a > l ? a : l; } break;
      case OP_MIN: { --sp; const double l = st[sp & (EXPR_MAXSTACK - 1)]; a = a < l ? a : l; } break;
      case OP_MEAN: --sp; a = 0.5 * st[sp & (EXPR_MAXSTACK - 1)] + 0.5 * a; break;
      case OP_ADDC: a = a + c; break;
      case OP_SUBC: a = a + (-c); break;
      case OP_MULC: a = a * c; break;
      case OP_DIVC: a = a / c; break;
      case OP_POWC: a = pow(a, c); break;
      case OP_ADDV: a = a + expr_var(in, (int)c); break;
      case OP_SUBV: a = a + (-expr_var(in, (int)c)); break;
      case OP_MULV: a = a * expr_var(in, (int)c); break;
      case OP_DIVV: a = a / expr_var(in, (int)c); break;
      case OP_SIN: a = sin(a); break;
      case OP_COS: a = cos(a); break;
      case OP_TAN: a = tan(a); break;
      case OP_EXP: a = exp(a); break;
      case OP_LOG: a = log(a); break;
      case OP_ABS: a = a < 0.0 ? -a : a; break;
      case OP_SQRT: a = a <= 0.0 ? 0.0 : sqrt(a); break;
      case OP_SINH: a = sinh(a); break;
      case OP_COSH: a = cosh(a); break;
      default: break;
    }
  }
  return a;
}

__device__ __forceinline__ double expr_eval(const ExprProgram& p, const ExprVars& in) {
  if (p.is_const) return p.cval;
  return expr_eval_program(p, in);
}
#endif  // !MRH_JIT

// thermal::defineFunctions names (thermal.cpp:47-65)
enum ThermalFn { FN_SOURCE = 0, FN_DIFFUSION, FN_SPECIFIC_HEAT, FN_DENSITY };

template <int DIM, int FN>
__device__ __forceinline__ double thermal_fn(const ThermalParams<DIM>& P, const double (&x)[3], double t) {
#ifdef MRH_JIT
  if constexpr (FN == FN_SOURCE) return mrh_fn_source(x[0], x[1], x[2], t);
  else if constexpr (FN == FN_DIFFUSION) return mrh_fn_diffusion(x[0], x[1], x[2], t);
  else if constexpr (FN == FN_SPECIFIC_HEAT) return mrh_fn_specific_heat(x[0], x[1], x[2], t);
  else return mrh_fn_density(x[0], x[1], x[2], t);
#else
  ExprVars in;
  in.v[0] = x[0]; in.v[1] = x[1]; in.v[2] = x[2]; in.v[3] = t; in.v[4] = in.v[5] = in.v[6] = 0.0;
  if constexpr (FN == FN_SOURCE) return expr_eval(P.source, in);
  else if constexpr (FN == FN_DIFFUSION) return expr_eval(P.diffusion, in);
  else if constexpr (FN == FN_SPECIFIC_HEAT) return expr_eval(P.specific_heat, in);
  else return expr_eval(P.density, in);
#endif
}

#ifdef MRH_JIT
#define MRH_ALL_CONST (MRH_JIT_ALL_CONST != 0)
#define MRH_SOURCE_CONST (MRH_JIT_SOURCE_CONST != 0)
#else
#define MRH_ALL_CONST (P.all_const != 0)
#define MRH_SOURCE_CONST (P.source.is_const != 0)
#endif

// ---------------------------------------------------------------------------------------------------------
// Gather of the element state with the transient stage/BDF combination folded in.
//   performGather / performGather4D          assemblyManager_gather.hpp:181-291
//   computeSolnSteadySeeded                  workset.cpp:864-901   (u = sol, du/ddof = 1)
//   computeSolnTransientSeeded (seedwhat 1)  workset.cpp:600-834
//       u   = alpha_u u_s + (1-alpha_u) u_prev0 + sum_{s'<s} A(s,s')/b(s') (u_stage[s'] - u_prev0)
//       u_t = alpha_t u_s + (sum_{k>=1} BDF(k) u_prev[k-1]) / (dt b(s))
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gather_dof(const double* __restrict__ sol, const TimeDev& td, int lid, double& u, double& ut) {
  const double s = __ldg(sol + lid);
  u = s; ut = 0.0;
  if (MRH_TRANSIENT(td)) {
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

// ---------------------------------------------------------------------------------------------------------
// Parallelepiped cell with constant coefficients: the local matrix is a linear combination of reference
// tables, K = kappa |det| sum_ab G_ab Stab_ab with G = J^-1 J^-T (BOX: J diagonal, only the G_aa terms exist).
// Each entry is staged as soon as it is final so that only r, u and G stay live in registers.
// ---------------------------------------------------------------------------------------------------------
template <int DIM, bool BOX>
__device__ __forceinline__ void thermal_affine(const ThermalParams<DIM>& P, const int (&cn)[1 << DIM], const double (&u)[1 << DIM],
                                               const double (&ut)[1 << DIM], double (&r)[1 << DIM], const int cap, double* __restrict__ st) {
  typedef Q1Shape<DIM> S;
  constexpr int NV = S::NV, NQ = S::NQ, NG = S::NG;
  constexpr int NGU = BOX ? DIM : NG;
  const TimeDev& td = P.td;
  constexpr int nb[3] = {1, 3, 4};  // +xi, +eta, +zeta neighbours of vertex 0 (Shards order)
  const double* vc[3] = {P.vx, P.vy, P.vz};
  double X0[DIM], J[DIM][DIM], G[NG];
  double adet;
#pragma unroll
  for (int d = 0; d < DIM; ++d) X0[d] = __ldg(vc[d] + cn[0]);
  const double xzero[3] = {0.0, 0.0, 0.0};
  const double kap = thermal_fn<DIM, FN_DIFFUSION>(P, xzero, td.time);
  if constexpr (BOX) {
    double h[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) h[d] = 0.5 * (__ldg(vc[d] + cn[nb[d]]) - X0[d]);
    double det = h[0];
#pragma unroll
    for (int d = 1; d < DIM; ++d) det *= h[d];
    const double inv = 1.0 / det;
    adet = fabs(det);
    const double kd = kap * adet;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      double ih = inv;  // 1/h_d = (product of the other h) / det
#pragma unroll
      for (int o = 0; o < DIM; ++o) if (o != d) ih *= h[o];
      G[d] = kd * ih * ih;
    }
#pragma unroll
    for (int d = 0; d < DIM; ++d)
#pragma unroll
      for (int a = 0; a < DIM; ++a) J[d][a] = (a == d) ? h[d] : 0.0;
  } else {
    double Ji[DIM][DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
#pragma unroll
      for (int a = 0; a < DIM; ++a) J[d][a] = 0.5 * (__ldg(vc[d] + cn[nb[a]]) - X0[d]);
    const double det = det_inverse<DIM>(J, Ji);
    adet = fabs(det);
    const double kd = kap * adet;
    int g = 0;
#pragma unroll
    for (int a = 0; a < DIM; ++a) {
      double s = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) s += Ji[a][d] * Ji[a][d];
      G[g++] = s * kd;
    }
#pragma unroll
    for (int a = 0; a < DIM; ++a)
#pragma unroll
      for (int b = a + 1; b < DIM; ++b) {
        double s = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) s += Ji[a][d] * Ji[b][d];
        G[g++] = s * kd;
      }
  }
  double md = 0.0;
  if (MRH_TRANSIENT(td)) md = thermal_fn<DIM, FN_DENSITY>(P, xzero, td.time) * thermal_fn<DIM, FN_SPECIFIC_HEAT>(P, xzero, td.time) * adet;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = i; j < NV; ++j) {
      const int t = tri<NV>(i, j);
      double k = 0.0;
#pragma unroll
      for (int g = 0; g < NGU; ++g) k += G[g] * MRH_TAB(Stab)[g][t];
      r[i] += k * u[j];
      if (j != i) r[j] += k * u[i];
      if (MRH_TRANSIENT(td)) {
        const double mm = md * MRH_TAB(Mtab)[t];
        r[i] += mm * ut[j];
        if (j != i) r[j] += mm * ut[i];
        k = td.alpha_u * k + td.alpha_t * mm;
      }
      st[t * cap] = k;
    }
  // source
  if (MRH_SOURCE_CONST) {
    const double f = thermal_fn<DIM, FN_SOURCE>(P, xzero, td.time) * adet;
#pragma unroll
    for (int i = 0; i < NV; ++i) r[i] -= f * MRH_TAB(Ltab)[i];
  }
#ifdef MRH_JIT
  else if constexpr (BOX) {
    // tensor-product points of an axis-aligned box: coordinates take MRH_NQAd distinct values per axis, and the generated
    // mrh_fn_source_box evaluates every one-coordinate sub-expression once per distinct value
    constexpr int nqa[3] = {MRH_NQA0, MRH_NQA1, MRH_NQA2};
    double xa[3][NQ];
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int i = 0; i < NQ; ++i) xa[d][i] = (d < DIM && i < nqa[d]) ? X0[d < DIM ? d : 0] + J[d < DIM ? d : 0][d < DIM ? d : 0] * (jit_tab::qax[d][i] + 1.0) : 0.0;
    double f[NQ];
    mrh_fn_source_box(xa[0], xa[1], xa[2], td.time, f);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const double fw = f[q] * MRH_TAB(qw)[q] * adet;
#pragma unroll
      for (int i = 0; i < NV; ++i) r[i] -= fw * MRH_TAB(phi)[q][i];
    }
  }
#endif
  else {
    MRH_UNROLL_Q
    for (int q = 0; q < NQ; ++q) {
      double x[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        if constexpr (BOX) x[d] = X0[d] + J[d][d] * (MRH_TAB(qpt)[q][d] + 1.0);
        else {
          double s = X0[d];
#pragma unroll
          for (int a = 0; a < DIM; ++a) s += J[d][a] * (MRH_TAB(qpt)[q][a] + 1.0);
          x[d] = s;
        }
      }
      const double fw = thermal_fn<DIM, FN_SOURCE>(P, x, td.time) * MRH_TAB(qw)[q] * adet;
#pragma unroll
      for (int i = 0; i < NV; ++i) r[i] -= fw * MRH_TAB(phi)[q][i];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Phase 1: one element -> staged upper triangle of dF/du and residual F
//   st points at this element's column of the ring slot; entry t lives at st[t * cap]
// ---------------------------------------------------------------------------------------------------------
template <int DIM>
__device__ __forceinline__ void thermal_element(const ThermalParams<DIM>& P, const int e, const int cap, double* __restrict__ st) {
  typedef Q1Shape<DIM> S;
  constexpr int NV = S::NV, NQ = S::NQ, NT = S::NT;
  const TimeDev& td = P.td;
  const double* vc[3] = {P.vx, P.vy, P.vz};

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
  const int ecls = P.eclass[e];

  // ---- gather + transient combination (u, u_t per dof)
  double u[NV], ut[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) gather_dof(P.sol, td, ld[i], u[i], ut[i]);

  double r[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) r[i] = 0.0;

  if (ecls != 0 && MRH_ALL_CONST) {
    if (ecls == 2) thermal_affine<DIM, true>(P, cn, u, ut, r, cap, st);
    else thermal_affine<DIM, false>(P, cn, u, ut, r, cap, st);
  } else {
    // ================= general path: per-point Jacobian and coefficients =================
    double K[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) K[t] = 0.0;
    double X[NV][DIM];
#pragma unroll
    for (int n = 0; n < NV; ++n)
#pragma unroll
      for (int d = 0; d < DIM; ++d) X[n][d] = __ldg(vc[d] + cn[n]);
    MRH_UNROLL_Q
    for (int q = 0; q < NQ; ++q) {
      double J[DIM][DIM], Ji[DIM][DIM];
      double x[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        double xs = 0.0;
#pragma unroll
        for (int a = 0; a < DIM; ++a) J[d][a] = 0.0;
#pragma unroll
        for (int n = 0; n < NV; ++n) {
          xs += MRH_TAB(gN)[q][n] * X[n][d];
#pragma unroll
          for (int a = 0; a < DIM; ++a) J[d][a] += X[n][d] * MRH_TAB(gdN)[q][n][a];
        }
        x[d] = xs;
      }
      const double det = det_inverse<DIM>(J, Ji);
      const double wd = fabs(det) * MRH_TAB(qw)[q];
      double g[NV][DIM];
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          double s = 0.0;
#pragma unroll
          for (int a = 0; a < DIM; ++a) s += Ji[a][d] * MRH_TAB(dphi)[q][i][a];
          g[i][d] = s;
        }
      const double kap = thermal_fn<DIM, FN_DIFFUSION>(P, x, td.time);
      const double f = thermal_fn<DIM, FN_SOURCE>(P, x, td.time);
      double gT[DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NV; ++j) s += u[j] * g[j][d];
        gT[d] = s * kap * wd;
      }
      double lin = -f * wd, mw = 0.0;
      if (MRH_TRANSIENT(td)) {
        const double rc = thermal_fn<DIM, FN_DENSITY>(P, x, td.time) * thermal_fn<DIM, FN_SPECIFIC_HEAT>(P, x, td.time);
        double Tt = 0.0;
#pragma unroll
        for (int j = 0; j < NV; ++j) Tt += ut[j] * MRH_TAB(phi)[q][j];
        lin += rc * Tt * wd;
        mw = td.alpha_t * rc * wd;
      }
      const double kw = td.alpha_u * kap * wd;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        double s = lin * MRH_TAB(phi)[q][i];
#pragma unroll
        for (int d = 0; d < DIM; ++d) s += gT[d] * g[i][d];
        r[i] += s;
        double gi[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) gi[d] = kw * g[i][d];
        const double mi = mw * MRH_TAB(phi)[q][i];
#pragma unroll
        for (int j = i; j < NV; ++j) {
          double k = mi * MRH_TAB(phi)[q][j];
#pragma unroll
          for (int d = 0; d < DIM; ++d) k += gi[d] * g[j][d];
          K[tri<NV>(i, j)] += k;
        }
      }
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) st[t * cap] = K[t];
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) st[(NT + i) * cap] = r[i];
}

// ---------------------------------------------------------------------------------------------------------
// Phase 2: rows completed by this step.  One warp per row; each lane owns one lane item of the row's pattern
// (up to 4 staged values of one CSR entry), partial sums of an entry are combined with shuffles, and the
// first lane of each entry stores it.  Row records and CSR row offsets of the step are staged in shared memory
// during phase 1; the lane items of the next row are fetched while the current row is summed.
//   res(row) (+)= -sum r_e[i]           (assemblyManager_scatter.hpp:227, sign convention -F)
//   J(row, col) (+)= sum dF_i/du_j      (:261-271)
//   fixed rows are skipped (:208, :253); in overwrite mode they receive the dofConstraints result
//   directly: J(d,d) = 1, rest of the row 0, res(d) = 0 (assemblyManager_constraints.hpp:125-138).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double lds_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

struct RowItems {  // one lane's item of a row chunk
  unsigned meta;
  uint4 src;
};

__device__ __forceinline__ RowItems load_items(const ChainDev& C, const uint4* __restrict__ isrc, const int4& rr, const int it0, const int lane) {
  RowItems R;
  R.meta = 0u;
  R.src = make_uint4(SRC_NONE, SRC_NONE, SRC_NONE, SRC_NONE);
  const int it = it0 + lane;
  if (it < (int)((unsigned)rr.w & 0xFFFFu) && !(((unsigned)rr.w >> 16) & ROW_FIXED)) {
    R.meta = __ldg(C.item_meta + rr.y + it);
    R.src = __ldg(isrc + rr.y + it);
  }
  return R;
}

template <bool HAS_RES, bool HAS_JAC, bool ACC>
__device__ __forceinline__ void pull_rows(const ChainDev& C, const OutDev& O, const int n_rows, const int parity, const unsigned ring_s,
                                          const int4* __restrict__ tab_rec, const int64_t* __restrict__ tab_base) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const uint4* __restrict__ isrc = reinterpret_cast<const uint4*>(parity ? C.item_src1 : C.item_src0);
  int lr = warp;
  if (lr >= n_rows) return;
  int4 rr = tab_rec[lr];
  int64_t base = tab_base[lr];
  RowItems cur = load_items(C, isrc, rr, 0, lane);
  for (;;) {
    const int lr_next = lr + nwarps;
    const bool has_next = lr_next < n_rows;
    int4 rr_n = rr;
    int64_t base_n = base;
    RowItems nxt = cur;
    if (has_next) {
      rr_n = tab_rec[lr_next];
      base_n = tab_base[lr_next];
      nxt = load_items(C, isrc, rr_n, 0, lane);
    }
    // ---- current row
    const int row = rr.x;
    const unsigned n_items = (unsigned)rr.w & 0xFFFFu;
    if (((unsigned)rr.w >> 16) & ROW_FIXED) {
      if (!ACC) {
        const unsigned diag_k = (unsigned)rr.z >> 16;
        const int len = (int)n_items;  // fixed rows carry their CSR length here
        if (HAS_JAC) for (int k = lane; k < len; k += 32) O.jac[base + k] = ((unsigned)k == diag_k) ? 1.0 : 0.0;
        if (HAS_RES && lane == 0) O.res[row] = 0.0;
      }
    } else {
      const unsigned rbase = ring_s + ((unsigned)rr.z & 0xFFFFu) * 8u;
      for (unsigned it0 = 0;;) {
        double acc = 0.0;
        if (cur.src.x != SRC_NONE) acc = lds_f64(rbase + cur.src.x);
        if (cur.src.y != SRC_NONE) acc += lds_f64(rbase + cur.src.y);
        if (cur.src.z != SRC_NONE) acc += lds_f64(rbase + cur.src.z);
        if (cur.src.w != SRC_NONE) acc += lds_f64(rbase + cur.src.w);
        double v = __shfl_down_sync(0xffffffffu, acc, 1);
        if (cur.meta & ITEM_ADD1) acc += v;
        if (C.need_add2) {
          v = __shfl_down_sync(0xffffffffu, acc, 2);
          if (cur.meta & ITEM_ADD2) acc += v;
        }
        const bool is_res = (cur.meta & ITEM_RES) != 0u;
        if ((cur.meta & ITEM_HEAD) && (is_res ? HAS_RES : HAS_JAC)) {
          double* p = is_res ? (O.res + row) : (O.jac + base + (cur.meta & 0xFFFFu));
          double val = is_res ? -acc : acc;
          if (ACC) val += *p;
          *p = val;
        }
        it0 += 32u;
        if (it0 >= n_items) break;
        cur = load_items(C, isrc, rr, (int)it0, lane);
      }
    }
    if (!has_next) break;
    lr = lr_next; rr = rr_n; base = base_n; cur = nxt;
  }
}

__device__ __forceinline__ void pull_dispatch(const ChainDev& C, const OutDev& O, const int n_rows, const int parity, const unsigned ring_s,
                                              const int4* tab_rec, const int64_t* tab_base) {
  const int mode = (O.res ? 1 : 0) | (O.jac ? 2 : 0) | (O.accumulate ? 4 : 0);
  switch (mode) {
    case 1: pull_rows<true, false, false>(C, O, n_rows, parity, ring_s, tab_rec, tab_base); break;
    case 2: pull_rows<false, true, false>(C, O, n_rows, parity, ring_s, tab_rec, tab_base); break;
    case 3: pull_rows<true, true, false>(C, O, n_rows, parity, ring_s, tab_rec, tab_base); break;
    case 5: pull_rows<true, false, true>(C, O, n_rows, parity, ring_s, tab_rec, tab_base); break;
    case 6: pull_rows<false, true, true>(C, O, n_rows, parity, ring_s, tab_rec, tab_base); break;
    case 7: pull_rows<true, true, true>(C, O, n_rows, parity, ring_s, tab_rec, tab_base); break;
    default: break;
  }
}

// Row record of a step -> shared-memory row table entry.  Fixed rows get their CSR length in place of n_items.
__device__ __forceinline__ void fetch_row(const ChainDev& C, const GraphDev& G, const int idx, int4& rec, int64_t& base) {
  rec = __ldg(reinterpret_cast<const int4*>(C.rows + idx));
  base = __ldg(G.rowptr + rec.x);
  if (((unsigned)rec.w >> 16) & ROW_FIXED) {
    const int len = (int)(__ldg(G.rowptr + rec.x + 1) - base);
    rec.w = (int)(((unsigned)rec.w & 0xFFFF0000u) | (unsigned)len);
  }
}

template <int DIM>
__device__ __forceinline__ void thermal_chain(const ThermalParams<DIM>& P) {
  extern __shared__ double ring[];
  typedef Q1Shape<DIM> S;
  const ChainDev& C = P.chains;
  const int s0 = __ldg(C.chain_step_ptr + blockIdx.x), s1 = __ldg(C.chain_step_ptr + blockIdx.x + 1);
  const int cap = C.cap;
  const int slot_doubles = cap * S::STAGE;
  const int row_tab = C.row_tab;
  int64_t* tab_base = reinterpret_cast<int64_t*>(ring + 2 * slot_doubles);
  int4* tab_rec = reinterpret_cast<int4*>(tab_base + row_tab);
  const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
  const int tid = threadIdx.x, nth = blockDim.x;
  for (int s = s0; s < s1; ++s) {
    const int4 sr = __ldg(reinterpret_cast<const int4*>(C.steps + s));
    const int elem_begin = sr.x, n_elem = sr.y, row_begin = sr.z, n_rows = sr.w;
    const int parity = (s - s0) & 1;
    double* slot = ring + parity * slot_doubles;
    // row-table prefetch (first pass): issued before the element work so its latency is hidden
    const int pass0 = n_rows < row_tab ? n_rows : row_tab;
    int4 rec0 = make_int4(0, 0, 0, 0);
    int64_t base0 = 0;
    if (tid < pass0) fetch_row(C, P.graph, row_begin + tid, rec0, base0);
    for (int le = tid; le < n_elem; le += nth)
      thermal_element<DIM>(P, __ldg(C.step_elems + elem_begin + le), cap, slot + le);
    if (tid < pass0) { tab_rec[tid] = rec0; tab_base[tid] = base0; }
    for (int i = tid + nth; i < pass0; i += nth) { int4 r; int64_t b; fetch_row(C, P.graph, row_begin + i, r, b); tab_rec[i] = r; tab_base[i] = b; }
    __syncthreads();
    pull_dispatch(C, P.out, pass0, parity, ring_s, tab_rec, tab_base);
    for (int done = pass0; done < n_rows; done += row_tab) {  // steps with more rows than the table holds
      __syncthreads();
      const int n = (n_rows - done) < row_tab ? (n_rows - done) : row_tab;
      for (int i = tid; i < n; i += nth) { int4 r; int64_t b; fetch_row(C, P.graph, row_begin + done + i, r, b); tab_rec[i] = r; tab_base[i] = b; }
      __syncthreads();
      pull_dispatch(C, P.out, n, parity, ring_s, tab_rec, tab_base);
    }
    __syncthreads();  // the next step overwrites the slot this pull read as "previous", and the row table
  }
}

#ifndef MRH_THREADS
#define MRH_THREADS 256
#endif
#ifndef MRH_MIN_BLOCKS
#define MRH_MIN_BLOCKS 2
#endif

#if defined(MRH_JIT) || defined(MRH_DEFINE_KERNELS)
#if !defined(MRH_JIT) || MRH_JIT_DIM == 2
extern "C" __global__ void __launch_bounds__(MRH_THREADS, MRH_MIN_BLOCKS) mrh_thermal_q1_2d(const __grid_constant__ ThermalParams<2> P) { thermal_chain<2>(P); }
#endif
#if !defined(MRH_JIT) || MRH_JIT_DIM == 3
extern "C" __global__ void __launch_bounds__(MRH_THREADS, MRH_MIN_BLOCKS) mrh_thermal_q1_3d(const __grid_constant__ ThermalParams<3> P) { thermal_chain<3>(P); }
#endif
#endif

}  // namespace mrhyde_b200
