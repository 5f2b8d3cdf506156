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
#define MRH_TAB(f) jit_tab::f      /* constexpr copy: folds into immediates / index arithmetic */
#define MRH_CTAB(f) jit_ctab::f    /* __constant__ copy: the value is a constant-bank operand of the FMA */
#ifdef MRH_JIT_LITERAL_TABLES
#define MRH_LTAB(f) jit_tab::f     /* literal: equal table entries (the tables are snapped) share one register */
#else
#define MRH_LTAB(f) jit_ctab::f
#endif
#define MRH_UNROLL_Q _Pragma("unroll")
#else
#define MRH_TRANSIENT(td) ((td).transient != 0)
#define MRH_TAB(f) P.tab.f
#define MRH_CTAB(f) P.tab.f
#define MRH_LTAB(f) P.tab.f
#define MRH_UNROLL_Q _Pragma("unroll 1")
#endif

// CLASS ring (plan-specialised builds, MRH_JIT_CLASS_NC): on a mesh of axis-aligned boxes with constant coefficients the
// local matrix K_e = sum_d G_d Stab[d] (+ md Mtab) takes only NC distinct values -- the classes of upper-triangle entries
// whose table columns coincide (8 on hexahedra, 4 on quadrilaterals: K_e(i,j) depends on which axes i and j differ in).
// A step then stages NC + NV doubles per element instead of NT + NV, with the plan's descriptors built on the class map
// (jit_tab::cls), so the pull is unchanged; the smaller ring lets more CTAs share an SM.
#ifdef MRH_JIT_CLASS_NC
#define MRH_STAGE_K(NT) MRH_JIT_CLASS_NC
#else
#define MRH_STAGE_K(NT) (NT)
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

#if defined(MRH_JIT) && !MRH_JIT_HAS_GENERAL && !defined(MRH_JIT_LATE_STAGE1)
#define MRH_EARLY_STAGE1 1
#endif
#ifdef MRH_JIT
#define MRH_HAS_BOX (MRH_JIT_HAS_BOX != 0)
#define MRH_HAS_AFFINE (MRH_JIT_HAS_AFFINE != 0)
#define MRH_HAS_GENERAL (MRH_JIT_HAS_GENERAL != 0)
#define MRH_ALL_CONST (MRH_JIT_ALL_CONST != 0)
#define MRH_SOURCE_CONST (MRH_JIT_SOURCE_CONST != 0)
#else
#define MRH_HAS_BOX true
#define MRH_HAS_AFFINE true
#define MRH_HAS_GENERAL true
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
// Inputs of one element, fetched one sweep step ahead (stage 1: connectivity and LIDs, stage 2: state and the
// vertices that span a parallelepiped) so that the global-memory latency is hidden behind the previous step's work.
template <int DIM>
struct ElemPre {
  int ecls;                       // 0 general cell, 1 parallelepiped, 2 axis-aligned box
  int cn[1 << DIM], ld[1 << DIM];
  double u[1 << DIM], ut[1 << DIM];
  double xv[DIM + 1][DIM];        // vertex 0 and its +xi, +eta, +zeta neighbours (Shards vertices 1, 3, 4)
};

#if defined(MRH_JIT)
#ifndef MRH_SRC_CACHE_N
#define MRH_SRC_CACHE_N 0
#endif
#ifndef MRH_SRC_SHARED_N
#define MRH_SRC_SHARED_N 0
#endif
#define MRH_SRC_REUSE (MRH_SRC_CACHE_N + MRH_SRC_SHARED_N > 0)
// Distinct quadrature-point coordinates per axis of an axis-aligned box (tensor-product points): the one place that computes them, so that
// the thread that fills the CTA-shared sub-expression values sees bitwise the coordinates every other thread would use.
template <int DIM>
__device__ __forceinline__ void box_axis_points(const ElemPre<DIM>& E, double (&xa)[3][1 << DIM]) {
  constexpr int NQ = 1 << DIM;
  constexpr int nqa[3] = {MRH_NQA0, MRH_NQA1, MRH_NQA2};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const int dd = d < DIM ? d : 0;
    const double x0 = E.xv[0][dd];
    const double h = 0.5 * (E.xv[dd + 1][dd] - x0);
#pragma unroll
    for (int i = 0; i < NQ; ++i) xa[d][i] = (d < DIM && i < nqa[d]) ? x0 + h * (jit_tab::qax[d][i] + 1.0) : 0.0;
  }
}
#endif

// kreg: class-ring builds of the register-staged pipeline (MRH_JIT_PIPE) return the class values here instead of storing them
template <int DIM, bool BOX>
__device__ __forceinline__ void thermal_affine(const ThermalParams<DIM>& P, const ElemPre<DIM>& E, double (&r)[1 << DIM], const int cap,
                                               double* __restrict__ st, double* __restrict__ kreg = nullptr, double* src_cache = nullptr, const int src_reuse = 0,
                                               const double* src_shared = nullptr) {
  const double (&u)[1 << DIM] = E.u;
  const double (&ut)[1 << DIM] = E.ut;
  typedef Q1Shape<DIM> S;
  constexpr int NV = S::NV, NQ = S::NQ, NG = S::NG;
  constexpr int NGU = BOX ? DIM : NG;
  const TimeDev& td = P.td;
  double X0[DIM], J[DIM][DIM], G[NG];
  double adet;
#pragma unroll
  for (int d = 0; d < DIM; ++d) X0[d] = E.xv[0][d];
  const double xzero[3] = {0.0, 0.0, 0.0};
  const double kap = thermal_fn<DIM, FN_DIFFUSION>(P, xzero, td.time);
  if constexpr (BOX) {
    double h[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) h[d] = 0.5 * (E.xv[d + 1][d] - X0[d]);
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
      for (int a = 0; a < DIM; ++a) J[d][a] = 0.5 * (E.xv[a + 1][d] - X0[d]);
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
#ifdef MRH_JIT_CLASS_NC
  if constexpr (BOX) {   // class-ring plans hold axis-aligned boxes only: the sheared instantiation is never executed
    constexpr int NC = MRH_JIT_CLASS_NC;
    double kc[NC], mc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      double k = 0.0;
#pragma unroll
      for (int g = 0; g < NGU; ++g) k += G[g] * MRH_LTAB(Stab)[g][jit_tab::rep[c]];
      kc[c] = k; mc[c] = 0.0;
      if (MRH_TRANSIENT(td)) { mc[c] = md * MRH_LTAB(Mtab)[jit_tab::rep[c]]; k = td.seed_u * k + td.seed_t * mc[c]; }
#ifdef MRH_JIT_PIPE
      kreg[c] = k;
#else
      st[c * cap] = k;
#endif
    }
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = jit_tab::cls[tri<NV>(i < j ? i : j, i < j ? j : i)];
        r[i] += kc[c] * u[j];
        if (MRH_TRANSIENT(td)) r[i] += mc[c] * ut[j];
      }
  }
#else
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = i; j < NV; ++j) {
      const int t = tri<NV>(i, j);
      double k = 0.0;
#pragma unroll
      for (int g = 0; g < NGU; ++g) k += G[g] * MRH_CTAB(Stab)[g][t];
      r[i] += k * u[j];
      if (j != i) r[j] += k * u[i];
      if (MRH_TRANSIENT(td)) {
        const double mm = md * MRH_CTAB(Mtab)[t];
        r[i] += mm * ut[j];
        if (j != i) r[j] += mm * ut[i];
        k = td.seed_u * k + td.seed_t * mm;
      }
      st[t * cap] = k;
    }
#endif
  // source
  if (MRH_SOURCE_CONST) {
    const double f = thermal_fn<DIM, FN_SOURCE>(P, xzero, td.time) * adet;
#pragma unroll
    for (int i = 0; i < NV; ++i) r[i] -= f * MRH_CTAB(Ltab)[i];
  }
#ifdef MRH_JIT
  else if constexpr (BOX) {
    // tensor-product points of an axis-aligned box: coordinates take MRH_NQAd distinct values per axis, and the generated
    // mrh_fn_source_box evaluates every one-coordinate sub-expression once per distinct value
    double xa[3][NQ];
    box_axis_points<DIM>(E, xa);
    double f[NQ];
#if MRH_SRC_REUSE
    // column cache: the one-coordinate sub-expressions of the axes the chain does not move along keep their values from step to step;
    // those of an axis along which all elements of a step agree come from shared memory (one warp evaluated them during the previous pull)
    double fresh[MRH_SRC_CACHE_N > 0 ? MRH_SRC_CACHE_N : 1];
#if MRH_SRC_SHARED_N > 0
    mrh_fn_source_box(xa[0], xa[1], xa[2], td.time, f, src_cache ? src_cache : fresh, src_cache ? src_reuse : 0, src_shared);
#else
    mrh_fn_source_box(xa[0], xa[1], xa[2], td.time, f, src_cache ? src_cache : fresh, src_cache ? src_reuse : 0);
#endif
#else
    mrh_fn_source_box(xa[0], xa[1], xa[2], td.time, f);
#endif
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const double fw = f[q] * MRH_LTAB(qw)[q] * adet;
#pragma unroll
      for (int i = 0; i < NV; ++i) r[i] -= fw * MRH_LTAB(phi)[q][i];
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
      const double fw = thermal_fn<DIM, FN_SOURCE>(P, x, td.time) * MRH_CTAB(qw)[q] * adet;
#pragma unroll
      for (int i = 0; i < NV; ++i) r[i] -= fw * MRH_CTAB(phi)[q][i];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Phase 1: one element -> staged upper triangle of dF/du and residual F
//   st points at this element's column of the ring slot; entry t lives at st[t * cap]
// ---------------------------------------------------------------------------------------------------------
// k: position of the element in the plan's step order
template <int DIM>
__device__ __forceinline__ void elem_stage1(const ThermalParams<DIM>& P, const int k, ElemPre<DIM>& E) {
  constexpr int NV = 1 << DIM;
  const int4* c4 = reinterpret_cast<const int4*>(P.chains.step_conn + (size_t)k * NV);
#ifdef MRH_JIT_LIDS_ARE_CONN
  // dof ids equal vertex ids on every element of the plan (one scalar HGRAD-C1 field numbered like the nodes): one stream
#pragma unroll
  for (int k = 0; k < NV / 4; ++k) {
    const int4 a = __ldg(c4 + k);
    E.cn[4 * k] = a.x; E.cn[4 * k + 1] = a.y; E.cn[4 * k + 2] = a.z; E.cn[4 * k + 3] = a.w;
    E.ld[4 * k] = a.x; E.ld[4 * k + 1] = a.y; E.ld[4 * k + 2] = a.z; E.ld[4 * k + 3] = a.w;
  }
#else
  const int4* l4 = reinterpret_cast<const int4*>(P.chains.step_lids + (size_t)k * NV);
#pragma unroll
  for (int k = 0; k < NV / 4; ++k) {
    const int4 a = __ldg(c4 + k), b = __ldg(l4 + k);
    E.cn[4 * k] = a.x; E.cn[4 * k + 1] = a.y; E.cn[4 * k + 2] = a.z; E.cn[4 * k + 3] = a.w;
    E.ld[4 * k] = b.x; E.ld[4 * k + 1] = b.y; E.ld[4 * k + 2] = b.z; E.ld[4 * k + 3] = b.w;
  }
#endif
#ifdef MRH_JIT_ONLY_ECLASS
  E.ecls = MRH_JIT_ONLY_ECLASS;
#else
  E.ecls = P.chains.step_eclass[k];
#endif
}

template <int DIM>
__device__ __forceinline__ void elem_stage2(const ThermalParams<DIM>& P, ElemPre<DIM>& E) {
  constexpr int NV = 1 << DIM;
  constexpr int nb[4] = {0, 1, 3, 4};  // vertex 0 and its +xi, +eta, +zeta neighbours (Shards order)
  const double* vc[3] = {P.vx, P.vy, P.vz};
#pragma unroll
  for (int i = 0; i < NV; ++i) gather_dof(P.sol, P.td, E.ld[i], E.u[i], E.ut[i]);
#pragma unroll
  for (int v = 0; v <= DIM; ++v)
#pragma unroll
    for (int d = 0; d < DIM; ++d) E.xv[v][d] = __ldg(vc[d] + E.cn[nb[v]]);
}

// L2 prefetch of what elem_stage2 will load: issued as soon as the next step's connectivity is in registers (before the pull), so the
// loads after the pull find their lines in L2 instead of paying the DRAM latency at the top of the next step.  No registers are held.
template <int DIM>
__device__ __forceinline__ void elem_prefetch2(const ThermalParams<DIM>& P, const ElemPre<DIM>& E) {
  constexpr int NV = 1 << DIM;
  constexpr int nb[4] = {0, 1, 3, 4};
  const double* vc[3] = {P.vx, P.vy, P.vz};
#if defined(MRH_JIT_PREFETCH2) && MRH_JIT_PREFETCH2 == 2
  // lean: the lines of the element's first vertex / dof only -- the other seven are some neighbour's first one on all but the rim of a column
  constexpr int NP = 1, VP = 0;
#else
  constexpr int NP = NV, VP = DIM;
#endif
#pragma unroll
  for (int i = 0; i < NP; ++i) asm volatile("prefetch.global.L2 [%0];" :: "l"(P.sol + E.ld[i]));
  if (MRH_TRANSIENT(P.td)) {
#pragma unroll
    for (int i = 0; i < NP; ++i) asm volatile("prefetch.global.L2 [%0];" :: "l"(P.td.prev[0] + E.ld[i]));
  }
#pragma unroll
  for (int v = 0; v <= VP; ++v)
#pragma unroll
    for (int d = 0; d < DIM; ++d)
      if (v == 0 || MRH_HAS_GENERAL || MRH_HAS_AFFINE || d == v - 1) asm volatile("prefetch.global.L2 [%0];" :: "l"(vc[d] + E.cn[nb[v]]));
}

template <int DIM>
__device__ __forceinline__ void thermal_element(const ThermalParams<DIM>& P, const ElemPre<DIM>& E, const int cap, double* __restrict__ st,
                                                double* src_cache = nullptr, const int src_reuse = 0, const double* src_shared = nullptr) {
  typedef Q1Shape<DIM> S;
  constexpr int NV = S::NV, NQ = S::NQ, NT = S::NT;
  const TimeDev& td = P.td;
  const double* vc[3] = {P.vx, P.vy, P.vz};
  const int (&cn)[NV] = E.cn;
  const double (&u)[NV] = E.u;
  const double (&ut)[NV] = E.ut;
  const int ecls = E.ecls;

  double r[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) r[i] = 0.0;

  // MRH_HAS_*: cell classes present in the plan's mesh (the specialised build drops the paths it cannot take)
  if (ecls != 0 && MRH_ALL_CONST && (MRH_HAS_BOX || MRH_HAS_AFFINE)) {
    if (MRH_HAS_BOX && (ecls == 2 || !MRH_HAS_AFFINE)) thermal_affine<DIM, true>(P, E, r, cap, st, nullptr, src_cache, src_reuse, src_shared);
    else if (MRH_HAS_AFFINE) thermal_affine<DIM, false>(P, E, r, cap, st);
  } else if (MRH_HAS_GENERAL) {
    // ================= general path: per-point Jacobian and coefficients =================
    // the local matrix is accumulated in this element's own ring column, point by point, rather than in 36
    // registers: the general path must not set the register footprint of the whole kernel
    double X[NV][DIM];
#pragma unroll
    for (int n = 0; n < NV; ++n)
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        // vertices 0, 1, 3, 4 came with the prefetch
        if (n == 0) X[n][d] = E.xv[0][d];
        else if (n == 1) X[n][d] = E.xv[1][d];
        else if (n == 3) X[n][d] = E.xv[2][d];
        else if (n == 4 && DIM == 3) X[n][d] = E.xv[DIM][d];
        else X[n][d] = __ldg(vc[d] + cn[n]);
      }
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
      const double wd = fabs(det) * MRH_CTAB(qw)[q];
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
        for (int j = 0; j < NV; ++j) Tt += ut[j] * MRH_CTAB(phi)[q][j];
        lin += rc * Tt * wd;
        mw = td.seed_t * rc * wd;
      }
      const double kw = td.seed_u * kap * wd;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        double s = lin * MRH_CTAB(phi)[q][i];
#pragma unroll
        for (int d = 0; d < DIM; ++d) s += gT[d] * g[i][d];
        r[i] += s;
        double gi[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) gi[d] = kw * g[i][d];
        const double mi = mw * MRH_CTAB(phi)[q][i];
#pragma unroll
        for (int j = i; j < NV; ++j) {
          double k = mi * MRH_CTAB(phi)[q][j];
#pragma unroll
          for (int d = 0; d < DIM; ++d) k += gi[d] * g[j][d];
          double* kp = st + tri<NV>(i, j) * cap;
          if (q == 0) *kp = k; else *kp += k;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) st[(MRH_STAGE_K(NT) + i) * cap] = r[i];
}

#ifdef MRH_JIT_PIPE
// Class ring, register-staged: the NC class values and the NV residual entries of one element (plans of axis-aligned boxes only)
template <int DIM>
__device__ __forceinline__ void thermal_element_class(const ThermalParams<DIM>& P, const ElemPre<DIM>& E, double (&out)[MRH_JIT_CLASS_NC + (1 << DIM)]) {
  constexpr int NV = 1 << DIM, NC = MRH_JIT_CLASS_NC;
  double r[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) r[i] = 0.0;
  thermal_affine<DIM, true>(P, E, r, 0, nullptr, out);
#pragma unroll
  for (int i = 0; i < NV; ++i) out[NC + i] = r[i];
}
#endif

// ---------------------------------------------------------------------------------------------------------
// Phase 2: rows completed by this step.  A warp takes a BATCH of up to 32 rows that share one gather pattern,
// one row per lane: all lanes walk the same slot descriptors (warp-uniform loads and branches) and differ only
// in their ring anchor, so the shared-memory reads of a warp are consecutive.  Each CSR entry is the sum of its
// staged contributions in ascending element order.  Results go through a small per-warp transpose buffer so
// that the global stores cover 32 contiguous bytes per row.
//   res(row) (+)= -sum r_e[i]           (assemblyManager_scatter.hpp:227, sign convention -F)
//   J(row, col) (+)= sum dF_i/du_j      (:261-271)
//   fixed rows are skipped (:208, :253); in overwrite mode they receive the dofConstraints result
//   directly: J(d,d) = 1, rest of the row 0, res(d) = 0 (assemblyManager_constraints.hpp:125-138).
// ---------------------------------------------------------------------------------------------------------
#ifdef MRH_JIT_CONST_DESC
#define MRH_DESC_LOAD(p) (*(p))
#else
#define MRH_DESC_LOAD(p) __ldg(p)
#endif
constexpr int PULL_CHUNK = 4;                       // CSR entries per transpose round
constexpr int PULL_PITCH = PULL_CHUNK + 1;          // doubles per lane in the transpose buffer (padding: no bank conflicts)
#ifdef MRH_JIT_ROWBUF
// generated pull code stages whole rows (32 row offsets, then 32 rows of MRH_JIT_ROWBUF doubles) and writes them out as
// one flat, coalesced stream (mrh_flush_rows in the generated prelude)
constexpr int PULL_WARP_DOUBLES = (32 + 32 * MRH_JIT_ROWBUF) > 32 * PULL_PITCH ? (32 + 32 * MRH_JIT_ROWBUF) : 32 * PULL_PITCH;
#else
constexpr int PULL_WARP_DOUBLES = 32 * PULL_PITCH;
#endif
#ifdef MRH_JIT_PULL
static_assert(PULL_CHUNK == 4 && PULL_PITCH == 5, "generated pull code assumes 4-entry chunks with pitch 5");
#endif

__device__ __forceinline__ double lds_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

// sum of one slot for this lane's row
__device__ __forceinline__ double slot_sum(const uint4* __restrict__ desc, const int k, const unsigned rbase) {
  const uint4 a = MRH_DESC_LOAD(desc + 2 * k);
  double acc = 0.0;
  if (a.x != SRC_NONE) acc = lds_f64(rbase + a.x);
  if (a.y != SRC_NONE) acc += lds_f64(rbase + a.y);
  if (a.z != SRC_NONE) acc += lds_f64(rbase + a.z);
  if (a.w != SRC_NONE) {
    acc += lds_f64(rbase + a.w);
    const uint4 b = MRH_DESC_LOAD(desc + 2 * k + 1);
    if (b.x != SRC_NONE) acc += lds_f64(rbase + b.x);
    if (b.y != SRC_NONE) acc += lds_f64(rbase + b.y);
    if (b.z != SRC_NONE) acc += lds_f64(rbase + b.z);
    if (b.w != SRC_NONE) acc += lds_f64(rbase + b.w);
  }
  return acc;
}

// Loads of the software pipeline are issued through volatile asm so that the compiler keeps them where they are
// written (early), instead of sinking them next to their first use.
__device__ __forceinline__ int4 ldg_pinned_v4(const void* p) {
  int4 v;
  asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

struct BatchRegs {  // what a lane holds of its batch: header (uniform), its row record and CSR offset
  int4 hdr;         // row_begin, desc_begin, n_rows | n_slots << 16, flags
  int2 rec;         // row, anchor | aux << 16
  int64_t base;
};

// two independent loads (the row table is addressed by batch index, not through the header)
__device__ __forceinline__ BatchRegs fetch_batch(const ChainDev& C, const GraphDev& G, const int batch, const int lane) {
  BatchRegs R;
  R.hdr = ldg_pinned_v4(C.batches + batch);
  const int4 rw = ldg_pinned_v4(C.rows + (size_t)batch * 32 + lane);   // idle lanes read zero padding (they never store)
  R.rec = make_int2(rw.x, rw.y);
  R.base = (int64_t)(((uint64_t)(uint32_t)rw.w << 32) | (uint64_t)(uint32_t)rw.z);
  return R;
}

#ifdef MRH_JIT_METRIC
// ---------------------------------------------------------------------------------------------------------
// METRIC ring (kernel_abi.h; plan-specialised builds only).  When every cell of the plan is a parallelepiped and
// diffusion / specific heat / density are constants, the local matrix is K_e = sum_g G_g(e) Stab[g] (+ md_e Mtab):
// phase 1 stages G (3 doubles on boxes, 6 on sheared cells), the load vector and the element state instead of
// the 36 + 8 doubles of the full local system, and the pull evaluates
//   J(row, k)  = alpha_u sum_e sum_g G_g(e) Stab[g][t_e] + alpha_t sum_e md_e Mtab[t_e]
//   res(row)   = -( sum_k (K_k u_k + M_k ut_k) - sum_e load_e[i_e] )
// per CSR entry (same thermal::volumeResidual integrand, thermal.cpp:70-165, summed per row instead of per element;
// the scatter / dofConstraints rules are those of pull_batch).  Entry m of element column c in ring slot s sits at
// double (m * 2 + s) * cap + c, so sources are addressed by their element column alone and the smaller ring lets
// more CTAs share an SM.
// ---------------------------------------------------------------------------------------------------------
template <int DIM>
struct MetricLayout {
  typedef Q1Shape<DIM> S;
  static constexpr int NGU = MRH_JIT_METRIC_NG;                      // DIM when every cell is an axis-aligned box, else NG
  static constexpr int MD = NGU;                                     // rho cp |det| (transient builds only)
  static constexpr int B0 = NGU + (MRH_JIT_TRANSIENT ? 1 : 0);       // load vector
  static constexpr int U0 = B0 + S::NV;                              // element state
  static constexpr int UT0 = U0 + S::NV;                             // time derivative (transient builds only)
  static constexpr int STAGE = UT0 + (MRH_JIT_TRANSIENT ? S::NV : 0);
  static constexpr int ES = 2 * MRH_JIT_CAP;                         // doubles between consecutive entries of a column
  static constexpr unsigned ESB = 8u * ES;
};

template <int DIM, bool BOX>
__device__ __forceinline__ void metric_cell(const ThermalParams<DIM>& P, const ElemPre<DIM>& E, double* __restrict__ st, double* src_cache = nullptr,
                                            const int src_reuse = 0) {
  typedef Q1Shape<DIM> S;
  typedef MetricLayout<DIM> L;
  constexpr int NV = S::NV, NQ = S::NQ, NG = S::NG;
  const TimeDev& td = P.td;
  double X0[DIM], J[DIM][DIM], G[NG];
  double adet;
#pragma unroll
  for (int d = 0; d < DIM; ++d) X0[d] = E.xv[0][d];
#pragma unroll
  for (int g = 0; g < NG; ++g) G[g] = 0.0;
  const double xzero[3] = {0.0, 0.0, 0.0};
  const double kap = thermal_fn<DIM, FN_DIFFUSION>(P, xzero, td.time);
  if constexpr (BOX) {
    double h[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) h[d] = 0.5 * (E.xv[d + 1][d] - X0[d]);
    double det = h[0];
#pragma unroll
    for (int d = 1; d < DIM; ++d) det *= h[d];
    const double inv = 1.0 / det;
    adet = fabs(det);
    const double kd = kap * adet;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      double ih = inv;
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
      for (int a = 0; a < DIM; ++a) J[d][a] = 0.5 * (E.xv[a + 1][d] - X0[d]);
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
#pragma unroll
  for (int g = 0; g < L::NGU; ++g) st[g * L::ES] = G[g];
  if (MRH_JIT_TRANSIENT)
    st[L::MD * L::ES] = thermal_fn<DIM, FN_DENSITY>(P, xzero, td.time) * thermal_fn<DIM, FN_SPECIFIC_HEAT>(P, xzero, td.time) * adet;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    st[(L::U0 + j) * L::ES] = E.u[j];
    if (MRH_JIT_TRANSIENT) st[(L::UT0 + j) * L::ES] = E.ut[j];
  }
  // load vector  sum_q f(x_q) w_q |det| phi_i(q)
  double fl[NV];
  if (MRH_SOURCE_CONST) {
    const double f = thermal_fn<DIM, FN_SOURCE>(P, xzero, td.time) * adet;
#pragma unroll
    for (int i = 0; i < NV; ++i) fl[i] = f * MRH_CTAB(Ltab)[i];
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) fl[i] = 0.0;
    if constexpr (BOX) {
      double xa[3][NQ];
      box_axis_points<DIM>(E, xa);
      double f[NQ];
#if MRH_SRC_REUSE
      double fresh[MRH_SRC_CACHE_N > 0 ? MRH_SRC_CACHE_N : 1];
#if MRH_SRC_SHARED_N > 0
      mrh_fn_source_box(xa[0], xa[1], xa[2], td.time, f, src_cache ? src_cache : fresh, src_cache ? (src_reuse & 7) : 0, nullptr);   // no CTA-shared values in the metric build
#else
      mrh_fn_source_box(xa[0], xa[1], xa[2], td.time, f, src_cache ? src_cache : fresh, src_cache ? src_reuse : 0);
#endif
#else
      mrh_fn_source_box(xa[0], xa[1], xa[2], td.time, f);
#endif
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const double fw = f[q] * MRH_LTAB(qw)[q] * adet;
#pragma unroll
        for (int i = 0; i < NV; ++i) fl[i] += fw * MRH_LTAB(phi)[q][i];
      }
    } else {
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        double x[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          double sx = X0[d];
#pragma unroll
          for (int a = 0; a < DIM; ++a) sx += J[d][a] * (MRH_TAB(qpt)[q][a] + 1.0);
          x[d] = sx;
        }
        const double fw = thermal_fn<DIM, FN_SOURCE>(P, x, td.time) * MRH_CTAB(qw)[q] * adet;
#pragma unroll
        for (int i = 0; i < NV; ++i) fl[i] += fw * MRH_CTAB(phi)[q][i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) st[(L::B0 + i) * L::ES] = fl[i];
}

template <int DIM>
__device__ __forceinline__ void thermal_element_metric(const ThermalParams<DIM>& P, const ElemPre<DIM>& E, double* __restrict__ st, double* src_cache = nullptr,
                                                       const int src_reuse = 0) {
  if (MRH_HAS_BOX && (E.ecls == 2 || !MRH_HAS_AFFINE)) metric_cell<DIM, true>(P, E, st, src_cache, src_reuse);
  else metric_cell<DIM, false>(P, E, st);
}

#ifdef MRH_JIT_CONST_MDESC
#define MRH_MDESC_LOAD(p) (*(p))
#else
#define MRH_MDESC_LOAD(p) __ldg(p)
#endif

// one CSR entry of this lane's row from the metric source words: jv = Jacobian value, rc = its share of the residual
template <int DIM, bool HAS_RES>
__device__ __forceinline__ void metric_slot(const uint4* __restrict__ desc, const int k, const unsigned rbase, const double au, const double at,
                                            double& jv, double& rc) {
  typedef MetricLayout<DIM> L;
  const uint4 a = MRH_MDESC_LOAD(desc + 2 * k);
  uint4 b = make_uint4(SRC_NONE, SRC_NONE, SRC_NONE, SRC_NONE);
  if (a.w != SRC_NONE) b = MRH_MDESC_LOAD(desc + 2 * k + 1);
  const unsigned w[SLOT_SRCS] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  double kp = 0.0, mp = 0.0;
#pragma unroll
  for (int z = 0; z < SLOT_SRCS; ++z)
    if (w[z] != SRC_NONE) {
      const unsigned col = rbase + (w[z] & MSRC_OFF_MASK);
      const int t = (int)((w[z] >> MSRC_T_SHIFT) & 63u);
#pragma unroll
      for (int g = 0; g < L::NGU; ++g) kp = fma(lds_f64(col + g * L::ESB), jit_ctab::Stab[g][t], kp);
      if (MRH_JIT_TRANSIENT) mp = fma(lds_f64(col + L::MD * L::ESB), jit_ctab::Mtab[t], mp);
    }
  jv = MRH_JIT_TRANSIENT ? au * kp + at * mp : kp;
  rc = 0.0;
  if (HAS_RES && a.x != SRC_NONE) {
    const unsigned col0 = rbase + (a.x & MSRC_OFF_MASK), j0 = (a.x >> MSRC_J_SHIFT) & 7u;
    rc = kp * lds_f64(col0 + (L::U0 + j0) * L::ESB);
    if (MRH_JIT_TRANSIENT) rc = fma(mp, lds_f64(col0 + (L::UT0 + j0) * L::ESB), rc);
  }
}

template <int DIM, bool HAS_RES, bool HAS_JAC, bool ACC>
__device__ __forceinline__ void pull_rows_metric(const int desc_begin, const int n_jac, const int parity, const unsigned rbase, const ChainDev& C,
                                                 double* __restrict__ wbuf, const int lane, const int rsub, const int kk_st, double* const (&pj)[4],
                                                 const bool (&rv)[4], double* pres, const bool active, const double au, const double at) {
  typedef MetricLayout<DIM> L;
#ifdef MRH_JIT_CONST_MDESC
  const uint4* __restrict__ desc = (parity ? mrh_mdesc1 : mrh_mdesc0) + 2 * desc_begin;
#else
  const uint4* __restrict__ desc = reinterpret_cast<const uint4*>(parity ? C.mdesc1 : C.mdesc0) + 2 * (size_t)desc_begin;
#endif
  double racc = 0.0;
  for (int k0 = 0; k0 < n_jac; k0 += PULL_CHUNK) {
#pragma unroll
    for (int kk = 0; kk < PULL_CHUNK; ++kk)
      if (k0 + kk < n_jac) {
        double jv, rc;
        metric_slot<DIM, HAS_RES>(desc, k0 + kk, rbase, au, at, jv, rc);
        if (HAS_JAC) wbuf[lane * PULL_PITCH + kk] = jv;
        racc += rc;
      }
    if (HAS_JAC) {
      __syncwarp();
      if (k0 + kk_st < n_jac) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (rv[j]) {
            double* p = pj[j] + k0;
            double v = wbuf[(rsub + 8 * j) * PULL_PITCH + kk_st];
            if (ACC) v += *p;
            *p = v;
          }
      }
      __syncwarp();
    }
  }
  if (HAS_RES) {
    const uint4 a = MRH_MDESC_LOAD(desc + 2 * n_jac);
    uint4 b = make_uint4(SRC_NONE, SRC_NONE, SRC_NONE, SRC_NONE);
    if (a.w != SRC_NONE) b = MRH_MDESC_LOAD(desc + 2 * n_jac + 1);
    const unsigned w[SLOT_SRCS] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    double bs = 0.0;
#pragma unroll
    for (int z = 0; z < SLOT_SRCS; ++z)
      if (w[z] != SRC_NONE) bs += lds_f64(rbase + (w[z] & MSRC_OFF_MASK) + (L::B0 + ((w[z] >> MSRC_I_SHIFT) & 7u)) * L::ESB);
    if (active) {
      double v = bs - racc;
      if (ACC) v += *pres;
      *pres = v;
    }
  }
}
#endif  // MRH_JIT_METRIC

template <int MDIM, bool HAS_RES, bool HAS_JAC, bool ACC>   // MDIM: 0 = full local systems in the ring, else the metric ring of that dimension
__device__ __forceinline__ void pull_batch(const BatchRegs& R, const ChainDev& C, const GraphDev& G, const OutDev& O, const int parity,
                                           const unsigned ring_s, double* __restrict__ wbuf, const int lane, const double au, const double at, const PushDev* X) {
  const int n_rows = R.hdr.z & 0xFFFF, n_slots = (int)((unsigned)R.hdr.z >> 16);
  const bool active = lane < n_rows;
  const int row = R.rec.x;
  if ((unsigned)R.hdr.w & BATCH_FIXED) {
    if (!ACC && active) {
      const int len = (int)(__ldg(G.rowptr + row + 1) - R.base);
      const int diag_k = (int)((unsigned)R.rec.y >> 16);
      if (HAS_JAC) for (int k = 0; k < len; ++k) O.jac[R.base + k] = (k == diag_k && O.diag_one) ? 1.0 : 0.0;
      if (HAS_RES) O.res[row] = 0.0;
    }
    return;
  }
#ifdef MRH_JIT_CONST_DESC
  const uint4* __restrict__ desc = (parity ? mrh_desc1 : mrh_desc0) + 2 * R.hdr.y;
#else
  const uint4* __restrict__ desc = reinterpret_cast<const uint4*>(parity ? C.desc1 : C.desc0) + 2 * (size_t)R.hdr.y;
#endif
  const unsigned rbase = ring_s + ((unsigned)R.rec.y & 0xFFFFu) * 8u;
  const int n_jac = n_slots - 1;
#ifdef MRH_JIT_ROWBUF
  // straight-line code generated for the plan's most frequent patterns: ring offsets (and, on the metric ring, table entries
  // and state slots) are immediates; rows leave through the per-warp row buffer as one coalesced stream
  {
    double* const pres2 = HAS_RES ? (O.res + row) : nullptr;
#ifdef MRH_JIT_METRIC
    if (MDIM != 0 && mrh_pull_metric_special<HAS_RES, HAS_JAC, ACC>(R.hdr.y, parity, rbase, wbuf, lane, n_rows, R.base, O.jac, pres2, active, au, at, X)) return;
#else
    if (mrh_pull_special<HAS_RES, HAS_JAC, ACC>(R.hdr.y, parity, rbase, wbuf, lane, n_rows, R.base, O.jac, pres2, active, X)) return;
#endif
  }
#if MRH_JIT_FLUSH == 2
  mrh_bulk_wait_read();   // the generic path below reuses the row buffer a bulk copy of this warp may still be reading
  __syncwarp();
#endif
#endif
  // store side of the transpose: this lane writes entry (k0 + kk_st) of rows rsub + 8 j
  const int rsub = lane >> 2, kk_st = lane & 3;
  double* pj[4] = {nullptr, nullptr, nullptr, nullptr};
  bool rv[4] = {false, false, false, false};
  if (HAS_JAC) {
    int64_t* wbase = reinterpret_cast<int64_t*>(wbuf);   // the row offsets are exchanged through the transpose buffer itself
    wbase[lane] = R.base;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) { pj[j] = O.jac + wbase[rsub + 8 * j] + kk_st; rv[j] = (rsub + 8 * j) < n_rows; }
    __syncwarp();
  }
  double* pres = HAS_RES ? (O.res + row) : nullptr;
#ifdef MRH_JIT_METRIC
  if constexpr (MDIM != 0) {
    pull_rows_metric<MDIM, HAS_RES, HAS_JAC, ACC>(R.hdr.y, n_jac, parity, rbase, C, wbuf, lane, rsub, kk_st, pj, rv, pres, active, au, at);
    return;
  }
#endif
  if (HAS_JAC) {
    for (int k0 = 0; k0 < n_jac; k0 += PULL_CHUNK) {
#pragma unroll
      for (int kk = 0; kk < PULL_CHUNK; ++kk)
        if (k0 + kk < n_jac) wbuf[lane * PULL_PITCH + kk] = slot_sum(desc, k0 + kk, rbase);
      __syncwarp();
      if (k0 + kk_st < n_jac) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (rv[j]) {
            double* p = pj[j] + k0;
            double v = wbuf[(rsub + 8 * j) * PULL_PITCH + kk_st];
            if (ACC) v += *p;
            *p = v;
          }
      }
      __syncwarp();
    }
  }
  if (HAS_RES) {
    const double acc = slot_sum(desc, n_jac, rbase);
    if (active) {
      double v = -acc;
      if (ACC) v += *pres;
      *pres = v;
    }
  }
}

template <int MDIM, bool HAS_RES, bool HAS_JAC, bool ACC>
__device__ __forceinline__ void pull_step(BatchRegs R, const ChainDev& C, const GraphDev& G, const OutDev& O, const int batch_begin, const int n_batches,
                                          const int parity, const unsigned ring_s, double* __restrict__ wbuf, const double au, const double at, const PushDev* X) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int b = warp; b < n_batches; b += nwarps) {
    const int bn = b + nwarps;
    BatchRegs N = R;
    if (bn < n_batches) N = fetch_batch(C, G, batch_begin + bn, lane);   // next batch of this warp: in flight while this one is summed
    pull_batch<MDIM, HAS_RES, HAS_JAC, ACC>(R, C, G, O, parity, ring_s, wbuf, lane, au, at, X);
    R = N;
  }
}

template <int DIM>
__device__ __forceinline__ void thermal_chain(const ThermalParams<DIM>& P) {
  extern __shared__ double ring[];
  typedef Q1Shape<DIM> S;
  const ChainDev& C = P.chains;
  const int chain = (int)blockIdx.x + C.chain_offset;
  const int s0 = __ldg(C.chain_step_ptr + chain), s1 = __ldg(C.chain_step_ptr + chain + 1);
  const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef MRH_JIT_METRIC
  constexpr int MDIM = DIM;
  constexpr int cap = MRH_JIT_CAP;
  constexpr int slot_doubles = cap;                      // interleaved slots: slot s starts at double s * cap
  double* wbuf = ring + 2 * cap * MetricLayout<DIM>::STAGE + warp * PULL_WARP_DOUBLES;
#else
  constexpr int MDIM = 0;
#ifdef MRH_JIT_CAP
  constexpr int cap = MRH_JIT_CAP;
#else
  const int cap = C.cap;
#endif
  const int slot_doubles = cap * (MRH_STAGE_K(S::NT) + S::NV);
  double* wbuf = ring + ((2 * slot_doubles + 1) & ~1) + warp * PULL_WARP_DOUBLES;   // 16-byte aligned (bulk copies read it)
#endif
  const double au = P.td.seed_u, at = P.td.seed_t;   // derivative seeds: J = au dR/du + at dR/du_t
#ifdef MRH_JIT_MODE
  constexpr int mode = MRH_JIT_MODE;   // output mode of this build: 1 res | 2 jac | 4 accumulate (the host picks the build)
#else
  const int mode = (P.out.res ? 1 : 0) | (P.out.jac ? 2 : 0) | (P.out.accumulate ? 4 : 0);
#endif
  // software pipeline over the steps.  While step s is computed and summed the inputs of step s+1 are in flight:
  //   top of step s   : step record of s+2 and this warp's batch of step s (addresses depend on step records only)
  //   after compute s : connectivity / LIDs of s+1 (step order: no element-id indirection)
  //   after pull s    : state and vertices of s+1 (need the LIDs / connectivity requested before the pull)
  // so no load is issued right behind the load that produces its address.
#if defined(MRH_JIT_STAGGER_NS) && MRH_JIT_STAGGER_NS > 0
  // CTAs of the first wave start in step with each other and, since every step costs about the same, stay so: all SMs then
  // compute, read and write at the same moments.  Offsetting the co-resident CTAs of an SM (and neighbouring SMs) by a
  // fraction of a step lets element work, shared-memory sums and the store stream of different CTAs overlap.
  if (blockIdx.x < MRH_JIT_STAGGER_SLOTS) {
    const unsigned d = ((blockIdx.x / MRH_JIT_STAGGER_SMS) % MRH_JIT_STAGGER_BLOCKS) * (MRH_JIT_STAGGER_NS / MRH_JIT_STAGGER_BLOCKS) +
                       ((blockIdx.x % MRH_JIT_STAGGER_SMS) % 4u) * (MRH_JIT_STAGGER_NS / (4u * MRH_JIT_STAGGER_BLOCKS));
    const long long t0 = clock64();
    const long long ticks = (long long)d * 2;   // ~2 GHz SM clock: the offset need not be exact
    while (clock64() - t0 < ticks) {}
  }
#endif
#ifdef MRH_JIT_PUSH
  const bool pushing = P.push.enabled && chain < P.push.n_push_chains;
  if (pushing && tid == 0) {   // the slab of this parity was read by the owner two calls ago (practically never waits)
    unsigned long long a;
    do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(a) : "l"(P.push.local_ack) : "memory"); } while (a + 2ull < P.push.epoch);
  }
#endif
#ifdef MRH_JIT_PIPE
  // Register-staged pipeline (class ring).  Per warp and step s:   P(s) = pull of the rows step s completes,  E(s+1) = element work of
  // the next step with its NC + NV staged values kept in REGISTERS;  then  barrier | store E(s+1) into the slot step s-1 occupied |
  // barrier.  P(s) and E(s+1) are independent -- P reads the slots of steps s and s-1, E reads global memory only -- so a warp runs
  // them back to back without a barrier in between, and even / odd warps run them in opposite order: at any time some warps of the
  // CTA are in the FP64-heavy element code and the others in the shared-memory / store-heavy pull, instead of all warps being in
  // the same phase (the two-barrier loop below).  Inputs: connectivity two steps ahead, state / vertices one step ahead.
  {
    constexpr int NST = MRH_JIT_CLASS_NC + S::NV;
    int4 sr = __ldg(reinterpret_cast<const int4*>(C.steps + s0));
    int4 sr1 = sr, sr2 = sr;
    if (s0 + 1 < s1) sr1 = __ldg(reinterpret_cast<const int4*>(C.steps + s0 + 1));
    if (s0 + 2 < s1) sr2 = __ldg(reinterpret_cast<const int4*>(C.steps + s0 + 2));
    ElemPre<DIM> E;
    double stg[NST];
    if (tid < sr.y) {
      elem_stage1<DIM>(P, sr.x + tid, E); elem_stage2<DIM>(P, E);
      thermal_element_class<DIM>(P, E, stg);
#pragma unroll
      for (int m = 0; m < NST; ++m) ring[m * cap + tid] = stg[m];
    }
    if (s0 + 1 < s1 && tid < sr1.y) { elem_stage1<DIM>(P, sr1.x + tid, E); elem_stage2<DIM>(P, E); }
    ElemPre<DIM> Nx;   // connectivity of step s + 2
    if (s0 + 2 < s1 && tid < sr2.y) elem_stage1<DIM>(P, sr2.x + tid, Nx);
    BatchRegs Rn;      // this warp's first batch of the next step
    Rn.hdr = make_int4(0, 0, 0, 0); Rn.rec = make_int2(0, 0); Rn.base = 0;
    if (warp < sr.w) Rn = fetch_batch(C, P.graph, sr.z + warp, lane);
    __syncthreads();
    for (int s = s0; s < s1; ++s) {
      const int batch_begin = sr.z, n_batches = sr.w;
      const int parity = (s - s0) & 1;
      const bool more = (s + 1 < s1) && (tid < sr1.y);        // this thread has an element in step s + 1 (inputs in E)
      const bool more2 = (s + 2 < s1) && (tid < sr2.y);       // ... and in step s + 2 (connectivity in Nx)
      int4 sr3 = sr2;
      if (s + 3 < s1) sr3 = ldg_pinned_v4(C.steps + s + 3);
      BatchRegs R = Rn;   // requested one step ahead
      if (s + 1 < s1) { Rn.hdr = make_int4(0, 0, 0, 0); Rn.rec = make_int2(0, 0); Rn.base = 0; if (warp < sr1.w) Rn = fetch_batch(C, P.graph, sr1.z + warp, lane); }
      auto pull = [&]() {
        switch (mode) {
          case 1: pull_step<MDIM, true, false, false>(R, C, P.graph, P.out, batch_begin, n_batches, parity, ring_s, wbuf, au, at, &P.push); break;
          case 2: pull_step<MDIM, false, true, false>(R, C, P.graph, P.out, batch_begin, n_batches, parity, ring_s, wbuf, au, at, &P.push); break;
          case 3: pull_step<MDIM, true, true, false>(R, C, P.graph, P.out, batch_begin, n_batches, parity, ring_s, wbuf, au, at, &P.push); break;
          case 5: pull_step<MDIM, true, false, true>(R, C, P.graph, P.out, batch_begin, n_batches, parity, ring_s, wbuf, au, at, &P.push); break;
          case 6: pull_step<MDIM, false, true, true>(R, C, P.graph, P.out, batch_begin, n_batches, parity, ring_s, wbuf, au, at, &P.push); break;
          case 7: pull_step<MDIM, true, true, true>(R, C, P.graph, P.out, batch_begin, n_batches, parity, ring_s, wbuf, au, at, &P.push); break;
          default: break;
        }
      };
      auto element = [&]() {
        if (more) thermal_element_class<DIM>(P, E, stg);
        if (more2) {   // the next step's connectivity has arrived: request its state / vertices, then the connectivity after it
          E.ecls = Nx.ecls;
#pragma unroll
          for (int i = 0; i < (1 << DIM); ++i) { E.cn[i] = Nx.cn[i]; E.ld[i] = Nx.ld[i]; }
          elem_stage2<DIM>(P, E);
        }
        if ((s + 3 < s1) && (tid < sr3.y)) elem_stage1<DIM>(P, sr3.x + tid, Nx);
      };
#if MRH_JIT_PIPE == 2
      // one copy of each body in the instruction stream: the order is a run-time property of the warp
#pragma unroll 1
      for (int h = 0; h < 2; ++h) { if (((warp ^ h) & 1) == 0) pull(); else element(); }
#else
      pull(); element();
#endif
      __syncthreads();   // every pull of step s is done: the slot of step s - 1 is free
      if (more) {
        double* slot = ring + (parity ^ 1) * slot_doubles;
#pragma unroll
        for (int m = 0; m < NST; ++m) slot[m * cap + tid] = stg[m];
      }
      sr = sr1; sr1 = sr2; sr2 = sr3;
      __syncthreads();   // step s + 1 is staged
    }
  }
#else
  int4 sr = __ldg(reinterpret_cast<const int4*>(C.steps + s0));
  int4 sr_next = sr;
  if (s0 + 1 < s1) sr_next = __ldg(reinterpret_cast<const int4*>(C.steps + s0 + 1));
  ElemPre<DIM> E;
  if (tid < sr.y) { elem_stage1<DIM>(P, sr.x + tid, E); elem_stage2<DIM>(P, E); }
#if MRH_SRC_REUSE
  double src_cache[MRH_SRC_CACHE_N > 0 ? MRH_SRC_CACHE_N : 1];
#pragma unroll
  for (int i = 0; i < (MRH_SRC_CACHE_N > 0 ? MRH_SRC_CACHE_N : 1); ++i) src_cache[i] = 0.0;
  const int chain_flags = C.chain_invariant ? (int)C.chain_invariant[chain] : 0;
  const int chain_inv = chain_flags & 7;
#if MRH_SRC_SHARED_N > 0
  // sub-expression values every element of a step shares: the last warp evaluates those of step s + 1 once it is through with the pull of
  // step s (interior steps leave it without a batch) and the barrier that ends the step publishes them
  const double* src_shared = wbuf - warp * PULL_WARP_DOUBLES + (MRH_THREADS / 32) * PULL_WARP_DOUBLES;
  const bool chain_shares = ((chain_flags >> 4) & MRH_SRC_SHARED_AXES) == MRH_SRC_SHARED_AXES;
  bool shared_ready = false;
#endif
#endif
  for (int s = s0; s < s1; ++s) {
    const int n_elem = sr.y, batch_begin = sr.z, n_batches = sr.w;
    const int parity = (s - s0) & 1;
    double* slot = ring + parity * slot_doubles;
    // unconditional (clamped) load: a predicated one needs a select right behind it, which waits for the load
    const int4 sr_next2 = ldg_pinned_v4(C.steps + min(s + 2, s1 - 1));
    // this warp's first batch of the step
    BatchRegs R;
    R.hdr = make_int4(0, 0, 0, 0); R.rec = make_int2(0, 0); R.base = 0;
    if (warp < n_batches) R = fetch_batch(C, P.graph, batch_begin + warp, lane);
    const bool more = (s + 1 < s1) && (tid < sr_next.y);
#ifdef MRH_EARLY_STAGE1
    // builds without general cells no longer need this step's connectivity: the next step's connectivity / LIDs are requested
    // before the element work, so they have arrived by the pull and never share a scoreboard with its first loads
    ElemPre<DIM> Nx;
    if (more) elem_stage1<DIM>(P, sr_next.x + tid, Nx);
#endif
#if defined(MRH_DEBUG_SKIP) && (MRH_DEBUG_SKIP & 2)
    if (tid < n_elem) slot[tid] = E.u[0] + E.xv[0][0];   // timing experiment: no element work
#elif defined(MRH_JIT_METRIC)
#if MRH_SRC_REUSE
    if (tid < n_elem) thermal_element_metric<DIM>(P, E, slot + tid, src_cache, s > s0 ? chain_inv : 0);
#else
    if (tid < n_elem) thermal_element_metric<DIM>(P, E, slot + tid);
#endif
#else
#if MRH_SRC_REUSE
#if MRH_SRC_SHARED_N > 0
    if (tid < n_elem) thermal_element<DIM>(P, E, cap, slot + tid, src_cache, (s > s0 ? chain_inv : 0) | (shared_ready ? (MRH_SRC_SHARED_AXES << 4) : 0), src_shared);
#else
    if (tid < n_elem) thermal_element<DIM>(P, E, cap, slot + tid, src_cache, s > s0 ? chain_inv : 0);
#endif
#else
    if (tid < n_elem) thermal_element<DIM>(P, E, cap, slot + tid);
#endif
#endif
#ifndef MRH_EARLY_STAGE1
    if (more) elem_stage1<DIM>(P, sr_next.x + tid, E);
#endif
#if defined(MRH_JIT_PREFETCH2) && defined(MRH_EARLY_STAGE1)
    if (more) elem_prefetch2<DIM>(P, Nx);  // the next step's connectivity arrived during the element work: nothing waits here
#endif
    __syncthreads();
#if defined(MRH_JIT_PREFETCH2) && !defined(MRH_EARLY_STAGE1)
    if (more) elem_prefetch2<DIM>(P, E);   // E holds the next step's connectivity (requested before the barrier)
#endif
#if defined(MRH_EARLY_STAGE1) && defined(MRH_JIT_EARLY_STAGE2)
    // the next step's state and vertices are requested before the pull (their addresses arrived during the element work),
    // so no global-memory latency is left exposed between two steps; costs the registers that hold them across the pull
    if (more) {
      E.ecls = Nx.ecls;
#pragma unroll
      for (int i = 0; i < (1 << DIM); ++i) { E.cn[i] = Nx.cn[i]; E.ld[i] = Nx.ld[i]; }
      elem_stage2<DIM>(P, E);
    }
#endif
    switch (mode) {
      case 1: pull_step<MDIM, true, false, false>(R, C, P.graph, P.out, batch_begin, n_batches, parity, ring_s, wbuf, au, at, &P.push); break;
      case 2: pull_step<MDIM, false, true, false>(R, C, P.graph, P.out, batch_begin, n_batches, parity, ring_s, wbuf, au, at, &P.push); break;
      case 3: pull_step<MDIM, true, true, false>(R, C, P.graph, P.out, batch_begin, n_batches, parity, ring_s, wbuf, au, at, &P.push); break;
      case 5: pull_step<MDIM, true, false, true>(R, C, P.graph, P.out, batch_begin, n_batches, parity, ring_s, wbuf, au, at, &P.push); break;
      case 6: pull_step<MDIM, false, true, true>(R, C, P.graph, P.out, batch_begin, n_batches, parity, ring_s, wbuf, au, at, &P.push); break;
      case 7: pull_step<MDIM, true, true, true>(R, C, P.graph, P.out, batch_begin, n_batches, parity, ring_s, wbuf, au, at, &P.push); break;
      default: break;
    }
#if defined(MRH_EARLY_STAGE1) && defined(MRH_JIT_EARLY_STAGE2)
    // requested before the pull
#else
#ifdef MRH_EARLY_STAGE1
    if (more) {
      E.ecls = Nx.ecls;
#pragma unroll
      for (int i = 0; i < (1 << DIM); ++i) { E.cn[i] = Nx.cn[i]; E.ld[i] = Nx.ld[i]; }
    }
#endif
    if (more) elem_stage2<DIM>(P, E);
#endif
#if defined(MRH_SRC_SHARED_N) && MRH_SRC_SHARED_N > 0 && !defined(MRH_JIT_METRIC)
    {
      // the last warp's first element of step s + 1 stands for all of them (the plan checked the coordinates bitwise)
      const bool shared_next = chain_shares && (s + 1 < s1) && ((MRH_THREADS / 32 - 1) * 32 < sr_next.y);
      if (shared_next && tid == (MRH_THREADS / 32 - 1) * 32) {
        double xa[3][1 << DIM], vals[MRH_SRC_SHARED_N];
        box_axis_points<DIM>(E, xa);
        mrh_fn_source_box_shared(xa[0], xa[1], xa[2], P.td.time, vals);
        double* dst = const_cast<double*>(src_shared);
#pragma unroll
        for (int i = 0; i < MRH_SRC_SHARED_N; ++i) dst[i] = vals[i];
      }
      shared_ready = shared_next;
    }
#endif
#ifdef MRH_JIT_PREFETCH_META
    // the record streams of step s + 2 (connectivity, row and batch records: each byte is read once, so every load of them is a DRAM
    // miss) are requested into L2 one whole step ahead, one 128-byte line per thread: no register, no scoreboard
    if (s + 2 < s1) {
      // only lines that hold bytes of the step: first line = the one of its first byte, count from the line of its last byte
      const unsigned long long c0 = reinterpret_cast<unsigned long long>(P.chains.step_conn + (size_t)sr_next2.x * (1 << DIM));
      const int cl = (int)((((c0 + (unsigned long long)(sr_next2.y * (1 << DIM) * 4 - 1)) >> 7) - (c0 >> 7)) + 1ull);
      const char* cb = reinterpret_cast<const char*>(c0 & ~127ull);
      const char* rb = reinterpret_cast<const char*>(C.rows + (size_t)sr_next2.z * 32);
      const int rl = sr_next2.w * 4;   // 32 records of 16 bytes per batch
      const char* p = nullptr;
      if (tid < cl) p = cb + tid * 128;
      else if (tid < cl + rl) p = rb + (tid - cl) * 128;
      else if (tid == cl + rl) p = reinterpret_cast<const char*>(C.batches + sr_next2.z);
#ifndef MRH_JIT_LIDS_ARE_CONN
      else if (tid - cl - rl - 1 < cl) p = cb + (reinterpret_cast<const char*>(P.chains.step_lids) - reinterpret_cast<const char*>(P.chains.step_conn)) + (tid - cl - rl - 1) * 128;
#endif
      if (p) asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
    }
#endif
    sr = sr_next; sr_next = sr_next2;
    __syncthreads();  // the next step overwrites the slot this pull read as "previous"
  }
#endif  // MRH_JIT_PIPE
#if defined(MRH_JIT_ROWBUF) && MRH_JIT_FLUSH == 2
  mrh_bulk_wait_read();   // shared memory must outlive the bulk copies that read it
#endif
#ifdef MRH_JIT_PUSH
  if (pushing) {
    // every row this chain pushed has landed in the owner's slab; the last push chain of the grid raises the owner's flag
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      __threadfence_system();
      if (atomicAdd(P.push.counter, 1u) == (unsigned)P.push.n_push_chains - 1u) {
        *P.push.counter = 0u;
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(P.push.remote_shift), "l"((unsigned long long)P.push.shift) : "memory");
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(P.push.remote_arrive), "l"(P.push.epoch) : "memory");
      }
    }
  }
#endif
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
