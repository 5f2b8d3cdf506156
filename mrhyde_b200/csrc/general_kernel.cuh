// General element kernel: residual + Jacobian of one batch of elements (or boundary sides) for any of the restated
// physics modules, with forward-mode AD derivative components across the threads of an element.
//
// Reference path replaced (one launch instead of ~20 per group of 100 elements):
//   performGather                       assemblyManager_gather.hpp:181-234
//   computeSoln{Steady,Transient}Seeded workset.cpp:864-901, 600-834
//   getPhysicalVolumetricBasis          discretizationInterface_basis.hpp:382-611   (push-forward, recomputed here)
//   getPhysicalIntegrationData          discretizationInterface_integration.hpp:218-262, side data :592-774
//   evaluateSolutionField               workset.cpp:978-1111
//   FunctionManager::evaluate           functionManager_evaluate.hpp:14-229
//   <module>::volumeResidual / boundaryResidual   (general_physics.cuh cites each)
// The fused scatter (assemblyManager_scatter.hpp:162-278) becomes `pull_rows_kernel` in general.cu.
//
// Mapping.  A CTA owns EPB elements; thread (e, g) owns K derivative components (columns) of element e.  Stages,
// separated by __syncthreads, all operands in shared memory:
//   S0 gather dofs (+ transient combination) and vertex coordinates
//   S1 cell Jacobian, inverse, determinant, weights, physical points (side: normals and side weights) per (e, q)
//   S2 push-forward of the reference basis tables: PB[b][i][q][k]   k = (val, grad xyz) | (val xyz, curl xyz) | (val xyz, div)
//   S3 solution fields F[v][k] = sum_i u_i PB, time-derivative fields, coefficient functions (bytecode) per (e, q)
//   S4 weak-form coefficients Cf[v][k] of the module at each point: values (double) by thread (e, q); derivative
//      components (dual numbers seeded with PB of the thread's columns) by thread (e, g), followed by the test-function
//      loop  dF_{(v,i)}/du_col += sum_k dCf[v][k] PB[b(v)][i][q][k]  with the N accumulators in registers
//   S5 residual rows F_{(v,i)} = sum_q sum_k Cf[v][k] PB
// Output: element matrices / vectors in a scratch array (row-major per element); the pull kernel sums them into the CSR
// values in ascending element order (the serial reference's order) without atomics.
//
// This header also compiles with a host compiler (MRH_HOST_EMULATION): the stage functions are plain functions of
// (block, thread) and general_emulate.cpp runs them in loops -- a debugging aid for the kernel logic on machines
// without a GPU, reachable only through mrhyde_b200_plan_debug_emulate on host-only plans; assemble_* never uses it.
#pragma once
#include "kernel_abi.h"

#if defined(__CUDACC__)
#define MRH_HD __device__ __forceinline__
#define MRH_CE __host__ __device__ constexpr
#define MRH_LDG(p) __ldg(p)
#else
#define MRH_HD inline
#define MRH_CE constexpr
#define MRH_LDG(p) (*(p))
#include <cmath>
#endif

namespace mrhyde_b200 {

// ---- dual number with K derivative components -----------------------------------------------------------
template <int K>
struct Dual {
  double v;
  double d[K];
  MRH_HD Dual() {}
  MRH_HD Dual(double x) : v(x) {
#pragma unroll
    for (int k = 0; k < K; ++k) d[k] = 0.0;
  }
};
#define MRH_DUAL_LOOP _Pragma("unroll") for (int k = 0; k < K; ++k)
template <int K> MRH_HD Dual<K> operator-(const Dual<K>& a) { Dual<K> r; r.v = -a.v; MRH_DUAL_LOOP r.d[k] = -a.d[k]; return r; }
template <int K> MRH_HD Dual<K> operator+(const Dual<K>& a, const Dual<K>& b) { Dual<K> r; r.v = a.v + b.v; MRH_DUAL_LOOP r.d[k] = a.d[k] + b.d[k]; return r; }
template <int K> MRH_HD Dual<K> operator-(const Dual<K>& a, const Dual<K>& b) { Dual<K> r; r.v = a.v - b.v; MRH_DUAL_LOOP r.d[k] = a.d[k] - b.d[k]; return r; }
template <int K> MRH_HD Dual<K> operator*(const Dual<K>& a, const Dual<K>& b) { Dual<K> r; r.v = a.v * b.v; MRH_DUAL_LOOP r.d[k] = a.d[k] * b.v + a.v * b.d[k]; return r; }
template <int K> MRH_HD Dual<K> operator/(const Dual<K>& a, const Dual<K>& b) {
  Dual<K> r; const double ib = 1.0 / b.v; r.v = a.v * ib; MRH_DUAL_LOOP r.d[k] = (a.d[k] - r.v * b.d[k]) * ib; return r;
}
template <int K> MRH_HD Dual<K> operator+(const Dual<K>& a, double b) { Dual<K> r = a; r.v += b; return r; }
template <int K> MRH_HD Dual<K> operator+(double a, const Dual<K>& b) { Dual<K> r = b; r.v += a; return r; }
template <int K> MRH_HD Dual<K> operator-(const Dual<K>& a, double b) { Dual<K> r = a; r.v -= b; return r; }
template <int K> MRH_HD Dual<K> operator-(double a, const Dual<K>& b) { Dual<K> r; r.v = a - b.v; MRH_DUAL_LOOP r.d[k] = -b.d[k]; return r; }
template <int K> MRH_HD Dual<K> operator*(const Dual<K>& a, double b) { Dual<K> r; r.v = a.v * b; MRH_DUAL_LOOP r.d[k] = a.d[k] * b; return r; }
template <int K> MRH_HD Dual<K> operator*(double a, const Dual<K>& b) { Dual<K> r; r.v = a * b.v; MRH_DUAL_LOOP r.d[k] = a * b.d[k]; return r; }
template <int K> MRH_HD Dual<K> operator/(const Dual<K>& a, double b) { const double ib = 1.0 / b; Dual<K> r; r.v = a.v * ib; MRH_DUAL_LOOP r.d[k] = a.d[k] * ib; return r; }
template <int K> MRH_HD Dual<K> operator/(double a, const Dual<K>& b) { Dual<K> r; const double ib = 1.0 / b.v; r.v = a * ib; MRH_DUAL_LOOP r.d[k] = -r.v * b.d[k] * ib; return r; }
// 1/sqrt(x) and sqrt(x) = x * (1/sqrt(x)) share one reciprocal square root; the derivative factors need no division
MRH_HD double mrh_rsqrt(double a) {
#if defined(__CUDA_ARCH__)
  return rsqrt(a);
#else
  return 1.0 / sqrt(a);
#endif
}
template <int K> MRH_HD Dual<K> mrh_rsqrt(const Dual<K>& a) { Dual<K> r; r.v = mrh_rsqrt(a.v); const double g = -0.5 * r.v * r.v * r.v; MRH_DUAL_LOOP r.d[k] = g * a.d[k]; return r; }
template <int K> MRH_HD Dual<K> mrh_sqrt(const Dual<K>& a) { Dual<K> r; const double rs = mrh_rsqrt(a.v); r.v = a.v * rs; const double g = 0.5 * rs; MRH_DUAL_LOOP r.d[k] = g * a.d[k]; return r; }
MRH_HD double mrh_sqrt(double a) { return a * mrh_rsqrt(a); }
template <int K> MRH_HD double mrh_val(const Dual<K>& a) { return a.v; }
MRH_HD double mrh_val(double a) { return a; }

// NN consecutive doubles from 16-byte aligned shared memory (NN even): LDS.128 on the device
template <int NN>
MRH_HD void mrh_ldn(const double* __restrict__ p, double* __restrict__ o) {
#if defined(__CUDA_ARCH__)
  const double2* p2 = reinterpret_cast<const double2*>(p);
#pragma unroll
  for (int i = 0; i < NN / 2; ++i) { const double2 t = p2[i]; o[2 * i] = t.x; o[2 * i + 1] = t.y; }
#else
  for (int i = 0; i < NN; ++i) o[i] = p[i];
#endif
}

// ---- launch parameters --------------------------------------------------------------------------------------
constexpr int GEN_MAXVARS = 5;   // navier stokes + thermal in 3-D: ux, pr, uy, uz, T
constexpr int GEN_MAXFN = 20;   // two-module blocks: 14 module functions + 4 boundary data
constexpr int GEN_MAXDOF = 96;
enum GenBasisType : int32_t { BT_HGRAD = 0, BT_HCURL = 1, BT_HDIV = 2, BT_HVOL = 3 };
enum GenBcType : int32_t { BC_NONE = 0, BC_DIRICHLET = 1, BC_WEAK_DIRICHLET = 2, BC_NEUMANN = 3 };

struct GenFnRec {   // one coefficient function: constant or a bytecode range in GenParams::fn_op / fn_c
  int32_t begin, n, is_const, pad;
  double cval;
};

struct GenOpts {    // module flags by their YAML key
  double form_param;     // "form_param"
  double penalty;        // linearelasticity "penalty"
  int32_t have_advection;  // thermal "include advection"
  int32_t useSUPG, usePSPG;
  int32_t uz_reference;    // navierstokes.cpp:688 defect reproduced (1) or corrected (0)
  int32_t incplanestress;
  int32_t leapfrog;
};

struct GenParams {
  // mesh (global memory)
  const double* vx; const double* vy; const double* vz;
  const int32_t* conn;      // [n_elem][nverts]
  const int32_t* lids;      // [n_elem][N]
  const int8_t* orient;     // [n_elem][N] or null
  const double* sol;
  TimeDev td;
  // work items: volume -> elements [item_begin, item_end); side -> items[item_begin .. item_end) hold element ids
  const int32_t* items;     // null for the volume launch
  int64_t item_begin, item_end;
  int64_t inst_base;        // scratch instance index of item 0 (volume: 0; side groups: n_elem + offset)
  // tables (global memory): geometry shape functions and reference bases at this launch's points
  const double* geo_N;      // [nq][nverts]
  const double* geo_dN;     // [nq][nverts][dim]
  const double* ref_tab;    // per basis [card][nq][ncb], concatenated in basis order
  const double* qwts;       // [nq] reference weights
  double tan_u[3], tan_v[3];  // side launches: reference side tangents
  // dof bookkeeping: element-local dof of (var, basis function)
  int16_t off[GEN_MAXVARS][GEN_MAXDOF];
  // functions
  GenFnRec fn[GEN_MAXFN];
  const uint8_t* fn_op;
  const double* fn_c;
  GenOpts opt;
  // side launches: boundary condition of each variable on this sideset and the index of its data function
  int32_t bc_type[GEN_MAXVARS];
  int32_t bc_fn[GEN_MAXVARS];
  // outputs (scratch)
  double* elem_jac;         // [n_inst][N][N]   may be null
  double* elem_res;         // [n_inst][N]      may be null
  int32_t epb;              // elements per CTA
  // mass mode (getWeightedMass, assemblyManager_mass.hpp:1065-1146): the module's point function is replaced by
  // Cf[v][value components] = mass_wts[v] * F[v] * w, so the "Jacobian" is the weighted mass matrix; 2: projection of the
  // initial conditions (gen_initial_point, residual stage only)
  int32_t mass_mode;
  double mass_wts[GEN_MAXVARS];
  // 1: the scratch element matrices / vectors are indexed variable-major, (v, i) -> v * card + i, instead of by the element-local
  // dof off[v][i] (device plans of the tensor-core contraction: a fragment row then is a contiguous run in global memory; the
  // pull's contribution list and position tables are permuted to match at upload, general.cu)
  int32_t var_major;
  int32_t adjoint;    // useadjoint: element matrices are stored transposed (updateJac, assemblyManager_jacres.hpp:1459-1475)
  int32_t tensor;     // host replay only: 1 = replay the tensor-core build's stages (S4d / S4m), 0 = the derivative-lane build's (S4b)
  int32_t fn_state;   // some coefficient function of this launch reads a solution field (GenFnRec::pad marks which)
};

// what the physics sees at one point
struct QpCtx {
  double x, y, z, t, w, h, ih, dt;   // ih = 1 / h
  double n[3];
  const double* fn;         // function values at this point
  const double* dfn;        // derivative components of the function values (state-dependent coefficients), [function][K]; may be null
  int32_t transient, stage;
  const int32_t* bc_type;
};

// Coefficient function i at the point, as the module's point function sees it.  STATE = false (no `Functions:` entry of the plan
// reads a solution field): a plain double, exactly as before.  STATE = true: the value type of the evaluation, i.e. for Dual<K> the
// value together with its derivative components -- the reference carries Sacado types through FunctionManager::evaluate for the
// same reason (functionManager_evaluate.hpp:59-229, the is_AD_ branches).
template <bool STATE>
MRH_HD double fn_get(const QpCtx& c, int i, const double*) { return c.fn[i]; }
template <bool STATE, int K>
MRH_HD auto fn_get(const QpCtx& c, int i, const Dual<K>*) {
  if constexpr (STATE) {
    Dual<K> r;
    r.v = c.fn[i];
#pragma unroll
    for (int k = 0; k < K; ++k) r.d[k] = c.dfn ? c.dfn[i * K + k] : 0.0;
    return r;
  } else {
    return c.fn[i];
  }
}
#define MRH_FN(i) fn_get<STATE>(c, (i), (const T*)nullptr)

// ---- bytecode evaluation (same op set and order as volume_kernel.cuh / expr.cpp) ----------------------------
// instructions [i0, i1) on an empty stack; `other` (may be null) fills the variables of another point of the element: it serves the
// element reductions, whose argument is re-run at every point (no nesting: the inner call passes null)
struct GenNoPoints { static constexpr bool kHasPoints = false; MRH_HD void operator()(int, double*) const {} };
template <int NXV, class Other>
MRH_HD double gen_expr_range(const uint8_t* __restrict__ ops, const double* __restrict__ cs, int i0, int i1, const double* __restrict__ var, int npts, const Other& other) {
  double st[16];
  int sp = 0;
  double a = 0.0;
  for (int i = i0; i < i1; ++i) {
    const double c = MRH_LDG(cs + i);
    const int opc = MRH_LDG(ops + i);
    if (opc >= OP_EBEGIN) {
      if (opc == OP_EBEGIN) continue;
      // functionManager_evaluate.hpp:413-460, literally: emax / emin keep the LAST value that beats the value at point 0; emean adds the
      // value at point 0 twice
      if constexpr (Other::kHasPoints) {   // (the argument's own evaluation never reduces again: no recursion in device code)
        const int b0 = i - (int)c + 1;
        double t0 = 0.0, r = 0.0;
        for (int qq = 0; qq < npts; ++qq) {
          double vq[NXV] = {0.0};
          other(qq, vq);
          const double t = gen_expr_range<NXV>(ops, cs, b0, i, vq, 0, GenNoPoints());
          if (qq == 0) { t0 = t; r = opc == OP_EMEAN ? t / (double)npts : t; }
          if (opc == OP_EMEAN) r += t / (double)npts;
          else if (opc == OP_EMAX ? (t > t0) : (t < t0)) r = t;
        }
        a = r;
      }
      continue;
    }
    switch (opc) {
      case OP_PUSHC: st[sp & 15] = a; ++sp; a = c; break;
      case OP_PUSHV: st[sp & 15] = a; ++sp; a = var[(int)c]; break;
      case OP_ADD: --sp; a = st[sp & 15] + a; break;
      case OP_SUB: --sp; a = st[sp & 15] + (-a); break;
      case OP_MUL: --sp; a = st[sp & 15] * a; break;
      case OP_DIV: --sp; a = st[sp & 15] / a; break;
      case OP_POW: --sp; a = pow(st[sp & 15], a); break;
      case OP_LT: --sp; a = st[sp & 15] < a ? 1.0 : 0.0; break;
      case OP_LTE: --sp; a = st[sp & 15] <= a ? 1.0 : 0.0; break;
      case OP_GT: --sp; a = st[sp & 15] > a ? 1.0 : 0.0; break;
      case OP_GTE: --sp; a = st[sp & 15] >= a ? 1.0 : 0.0; break;
      case OP_MAX: { --sp; const double l = st[sp & 15]; a = a > l ? a : l; } break;
      case OP_MIN: { --sp; const double l = st[sp & 15]; a = a < l ? a : l; } break;
      case OP_MEAN: --sp; a = 0.5 * st[sp & 15] + 0.5 * a; break;
      case OP_ADDC: a = a + c; break;
      case OP_SUBC: a = a + (-c); break;
      case OP_MULC: a = a * c; break;
      case OP_DIVC: a = a / c; break;
      case OP_POWC: a = pow(a, c); break;
      case OP_ADDV: a = a + var[(int)c]; break;
      case OP_SUBV: a = a + (-var[(int)c]); break;
      case OP_MULV: a = a * var[(int)c]; break;
      case OP_DIVV: a = a / var[(int)c]; break;
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
MRH_HD double gen_expr_eval(const GenFnRec& f, const uint8_t* __restrict__ ops, const double* __restrict__ cs, const double* __restrict__ var) {
  if (f.is_const) return f.cval;
  return gen_expr_range<7>(ops, cs, f.begin, f.begin + f.n, var, 0, GenNoPoints());
}

// The same program on (value, one derivative component): var / dvar hold the variables and their derivative components.  Comparisons,
// abs at 0 and max / min ties take the branch of the values (their derivative is that of the selected operand), as Sacado does.
MRH_HD void gen_expr_eval_dual(const GenFnRec& f, const uint8_t* __restrict__ ops, const double* __restrict__ cs, const double* __restrict__ var,
                               const double* __restrict__ dvar, double& val, double& der) {
  if (f.is_const) { val = f.cval; der = 0.0; return; }
  double st[16], sd[16];
  int sp = 0;
  double a = 0.0, da = 0.0;
  for (int i = f.begin; i < f.begin + f.n; ++i) {
    const double c = MRH_LDG(cs + i);
    const int op = MRH_LDG(ops + i);
    switch (op) {
      case OP_PUSHC: st[sp & 15] = a; sd[sp & 15] = da; ++sp; a = c; da = 0.0; break;
      case OP_PUSHV: st[sp & 15] = a; sd[sp & 15] = da; ++sp; a = var[(int)c]; da = dvar[(int)c]; break;
      case OP_ADD: --sp; a = st[sp & 15] + a; da = sd[sp & 15] + da; break;
      case OP_SUB: --sp; a = st[sp & 15] + (-a); da = sd[sp & 15] - da; break;
      case OP_MUL: { --sp; const double l = st[sp & 15], dl = sd[sp & 15]; da = dl * a + l * da; a = l * a; } break;
      case OP_DIV: { --sp; const double l = st[sp & 15], dl = sd[sp & 15]; const double q = l / a; da = (dl - q * da) / a; a = q; } break;
      case OP_POW: { --sp; const double l = st[sp & 15], dl = sd[sp & 15]; const double p = pow(l, a);
                     da = (a != 0.0 ? a * pow(l, a - 1.0) * dl : 0.0) + (da != 0.0 ? p * log(l) * da : 0.0); a = p; } break;
      case OP_LT: --sp; a = st[sp & 15] < a ? 1.0 : 0.0; da = 0.0; break;
      case OP_LTE: --sp; a = st[sp & 15] <= a ? 1.0 : 0.0; da = 0.0; break;
      case OP_GT: --sp; a = st[sp & 15] > a ? 1.0 : 0.0; da = 0.0; break;
      case OP_GTE: --sp; a = st[sp & 15] >= a ? 1.0 : 0.0; da = 0.0; break;
      case OP_MAX: { --sp; const double l = st[sp & 15]; if (!(a > l)) { a = l; da = sd[sp & 15]; } } break;
      case OP_MIN: { --sp; const double l = st[sp & 15]; if (!(a < l)) { a = l; da = sd[sp & 15]; } } break;
      case OP_MEAN: --sp; a = 0.5 * st[sp & 15] + 0.5 * a; da = 0.5 * sd[sp & 15] + 0.5 * da; break;
      case OP_ADDC: a = a + c; break;
      case OP_SUBC: a = a + (-c); break;
      case OP_MULC: a = a * c; da = da * c; break;
      case OP_DIVC: a = a / c; da = da / c; break;
      case OP_POWC: da = c != 0.0 ? c * pow(a, c - 1.0) * da : 0.0; a = pow(a, c); break;
      case OP_ADDV: a = a + var[(int)c]; da = da + dvar[(int)c]; break;
      case OP_SUBV: a = a + (-var[(int)c]); da = da - dvar[(int)c]; break;
      case OP_MULV: { const double r = var[(int)c]; da = da * r + a * dvar[(int)c]; a = a * r; } break;
      case OP_DIVV: { const double r = var[(int)c]; const double q = a / r; da = (da - q * dvar[(int)c]) / r; a = q; } break;
      case OP_SIN: da = cos(a) * da; a = sin(a); break;
      case OP_COS: da = -sin(a) * da; a = cos(a); break;
      case OP_TAN: { const double tt = tan(a); da = (1.0 + tt * tt) * da; a = tt; } break;
      case OP_EXP: a = exp(a); da = a * da; break;
      case OP_LOG: da = da / a; a = log(a); break;
      case OP_ABS: if (a < 0.0) { a = -a; da = -da; } break;
      case OP_SQRT: if (a <= 0.0) { a = 0.0; da = 0.0; } else { a = sqrt(a); da = 0.5 * da / a; } break;
      case OP_SINH: da = cosh(a) * da; a = sinh(a); break;
      case OP_COSH: da = sinh(a) * da; a = cosh(a); break;
      default: break;
    }
  }
  val = a; der = da;
}

// computeSolnTransientSeeded folded into the gather (same formulas as volume_kernel.cuh: gather_dof)
MRH_HD void gen_gather_dof(const double* __restrict__ sol, const TimeDev& td, int lid, double& u, double& ut) {
  const double s = MRH_LDG(sol + lid);
  u = s; ut = 0.0;
  if (td.transient) {
    const double p0 = MRH_LDG(td.prev[0] + lid);
    double bu = td.one_minus_alpha_u * p0;
    for (int k = 0; k < td.nstage_lo; ++k) bu += td.stage_w[k] * (MRH_LDG(td.stg[k] + lid) - p0);
    u = td.alpha_u * s + bu;
    double bt = td.bdf[1] * p0;
    for (int k = 2; k <= td.nprev; ++k) bt += td.bdf[k] * MRH_LDG(td.prev[k - 1] + lid);
    bt *= td.timewt;
    ut = td.alpha_t * s + bt;
  }
}

// weighted mass "physics": block-diagonal in the variables, value components only
template <class Phys, class T>
MRH_HD void gen_mass_point(const QpCtx& c, const double* mw, const T (&F)[Phys::NVAR][Phys::NC], T (&Cf)[Phys::NVAR][Phys::NC]) {
#pragma unroll
  for (int v = 0; v < Phys::NVAR; ++v)
#pragma unroll
    for (int k = 0; k < Phys::NC; ++k)
      if (k < Phys::nval(Phys::var_basis(v))) Cf[v][k] = (mw[v] * c.w) * F[v][k];
}

// setInitial, right-hand side of the L2 projection (assemblyManager_initial.hpp:260-311): the point function is replaced by
// Cf[v][value components] = initial_v[component](x) * w; in this mode the function slots hold the "initial <var>[...]" functions
// in (variable, component) order instead of the module's coefficient functions
template <class Phys, class T>
MRH_HD void gen_initial_point(const QpCtx& c, T (&Cf)[Phys::NVAR][Phys::NC]) {
  int slot = 0;
#pragma unroll
  for (int v = 0; v < Phys::NVAR; ++v)
#pragma unroll
    for (int k = 0; k < Phys::NC; ++k)
      if (k < Phys::nval(Phys::var_basis(v))) Cf[v][k] = c.fn[slot++] * c.w;
}

// ---- shared-memory layout of one element (offsets in doubles; every block is a multiple of 2 doubles) ---------
template <class Phys, int NQ, bool WITH_D = false>   // WITH_D: the tensor-core build's D region (requested for single-basis layouts only)
struct GenLayout {
  static constexpr int DIM = Phys::DIM, NV = 1 << DIM, NVAR = Phys::NVAR, NC = Phys::NC, NFN = Phys::NFN + (Phys::NVAR > 4 ? Phys::NVAR : 4);   // function slots per point: the module's functions + boundary data per variable
  static constexpr int N = Phys::N;
  static constexpr int GEO = 28;   // w, x y z, Jinv[9], J[9], det, pad, n[3], pad
  static MRH_CE int even(int x) { return (x + 1) & ~1; }
  // PB[b][i][q][k]: one basis function is a ROW of prs(b) doubles; the single-basis layouts pad the row by two doubles so that the
  // eight rows a tensor-core fragment load touches fall into distinct shared-memory banks (an unpadded row is a multiple of 128 bytes)
  static MRH_CE int prs(int b) { return NQ * Phys::ncb(b) + (Phys::NBASIS == 1 ? 2 : 0); }
  static MRH_HD int prs_rt(int b) { return NQ * Phys::ncb_rt(b) + (Phys::NBASIS == 1 ? 2 : 0); }
  static MRH_CE int pb_size() { int s = 0; for (int b = 0; b < Phys::NBASIS; ++b) s += Phys::card(b) * prs(b); return s; }
  static MRH_CE int pb_off(int b) { int s = 0; for (int c = 0; c < b; ++c) s += Phys::card(c) * prs(c); return s; }
  static MRH_CE int rt_off(int b) { int s = 0; for (int c = 0; c < b; ++c) s += Phys::card(c) * NQ * Phys::ncb(c); return s; }   // reference tables (global memory, unpadded)
  static constexpr int U = 0;
  static constexpr int UT = U + even(N);
  static constexpr int VX = UT + even(N);
  static constexpr int G = VX + even(NV * 3);
  static constexpr int H = G + NQ * GEO;
  static constexpr int PB = H + 2;
  static constexpr int FV = PB + even(pb_size());
  static constexpr int FT = FV + NQ * NVAR * NC;
  static constexpr int FN = FT + NQ * NVAR * NC;
  static constexpr int CV = FN + even(NQ * NFN);
  // tensor-core contraction (single-basis HGRAD modules): D_q = d Cf / d F at every point, [q][(v,k)][(w,l)]
  static constexpr bool TC_CAPABLE = (Phys::NBASIS == 1);
  static constexpr bool TC = TC_CAPABLE && WITH_D;
  static constexpr int NCV = NVAR * NC;
  static constexpr int DM = CV + NQ * NVAR * NC;
  static constexpr int DS = NCV + 2;   // row stride of D (padded: the four rows a fragment load touches sit in distinct banks)
  static constexpr int SIZE = DM + (TC ? NQ * NCV * DS : 0);
};

#if defined(__CUDA_ARCH__)
// FP64 tensor-core tile: C (8x8) += A (8x4, row) * B (4x8, col).  Lane l holds A[l/4][l%4], B[l%4][l/4], C[l/4][2 (l%4) + {0,1}].
__device__ __forceinline__ void mrh_dmma(double (&c)[2], const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
#endif

// ---- the stages ----------------------------------------------------------------------------------------------
// STATEK: build for plans whose coefficient functions read solution fields (their values are evaluated after the fields and
// differentiated in the Jacobian stages); plans without such functions run the STATEK = false build and pay nothing for the feature
template <class Phys, int NQ, int K, bool SIDE, bool TCK = false, bool STATEK = false>
struct GenBlock {
  typedef GenLayout<Phys, NQ, TCK> L;
  static constexpr int DIM = Phys::DIM, NV = L::NV, N = L::N, NVAR = L::NVAR, NC = L::NC;
  static constexpr int TPE = N / K;   // threads per element in the derivative stage
  static_assert(N % K == 0, "K must divide the element dof count");

  MRH_HD static int64_t item_of(const GenParams& P, int blk, int el) { return P.item_begin + (int64_t)blk * P.epb + el; }
  MRH_HD static int32_t elem_of(const GenParams& P, int64_t item) { return SIDE ? MRH_LDG(P.items + item) : (int32_t)item; }

  // S0: gather
  MRH_HD static void s0(const GenParams& P, double* sm, int blk, int idx) {
    if (idx < P.epb * N) {
      const int el = idx / N, c = idx % N;
      const int64_t item = item_of(P, blk, el);
      double u = 0.0, ut = 0.0;
      if (item < P.item_end) gen_gather_dof(P.sol, P.td, MRH_LDG(P.lids + (int64_t)elem_of(P, item) * N + c), u, ut);
      sm[el * L::SIZE + L::U + c] = u;
      sm[el * L::SIZE + L::UT + c] = ut;
    }
    if (idx < P.epb * NV) {
      const int el = idx / NV, n = idx % NV;
      const int64_t item = item_of(P, blk, el);
      double x = 0.0, y = 0.0, z = 0.0;
      if (item < P.item_end) {
        const int32_t v = MRH_LDG(P.conn + (int64_t)elem_of(P, item) * NV + n);
        x = MRH_LDG(P.vx + v); y = MRH_LDG(P.vy + v);
        if (DIM == 3) z = MRH_LDG(P.vz + v);
      } else {   // padding element: unit reference cell keeps every stage finite
        x = (n & 1) ^ ((n >> 1) & 1) ? 1.0 : 0.0; y = (n >> 1) & 1 ? 1.0 : 0.0; z = (n >> 2) & 1 ? 1.0 : 0.0;
      }
      double* vxs = sm + el * L::SIZE + L::VX;
      vxs[n * 3] = x; vxs[n * 3 + 1] = y; vxs[n * 3 + 2] = z;
    }
  }

  // S1: geometry at (element, point)
  MRH_HD static void s1(const GenParams& P, double* sm, int /*blk*/, int idx) {
    if (idx >= P.epb * NQ) return;
    const int el = idx / NQ, q = idx % NQ;
    const double* vxs = sm + el * L::SIZE + L::VX;
    double* g = sm + el * L::SIZE + L::G + q * L::GEO;
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, x[3] = {0, 0, 0};
    for (int n = 0; n < NV; ++n) {
      const double Nn = MRH_LDG(P.geo_N + q * NV + n);
      for (int i = 0; i < DIM; ++i) {
        x[i] += vxs[n * 3 + i] * Nn;
        for (int j = 0; j < DIM; ++j) J[i][j] += vxs[n * 3 + i] * MRH_LDG(P.geo_dN + (q * NV + n) * DIM + j);
      }
    }
    double Ji[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, det;
    if (DIM == 2) {
      det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
      Ji[0][0] = J[1][1] / det; Ji[0][1] = -J[0][1] / det; Ji[1][0] = -J[1][0] / det; Ji[1][1] = J[0][0] / det;
    } else {
      const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2], c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
      det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
      Ji[0][0] = c00 / det; Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det; Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
      Ji[1][0] = c01 / det; Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det; Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
      Ji[2][0] = c02 / det; Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det; Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
    }
    double w, nrm[3] = {0, 0, 0};
    const double wr = MRH_LDG(P.qwts + q);
    if (!SIDE) {
      w = (det < 0.0 ? -det : det) * wr;
    } else if (DIM == 2) {   // getPhysicalBoundaryIntegrationData: tangent -> rotated normal, measure = |tangent|
      double t[2] = {0, 0};
      for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) t[i] += J[i][j] * P.tan_u[j];
      nrm[0] = t[1]; nrm[1] = -t[0];
      const double len = sqrt(t[0] * t[0] + t[1] * t[1]);
      w = len * wr;
      nrm[0] /= len; nrm[1] /= len;
    } else {
      double tu[3] = {0, 0, 0}, tv[3] = {0, 0, 0};
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { tu[i] += J[i][j] * P.tan_u[j]; tv[i] += J[i][j] * P.tan_v[j]; }
      nrm[0] = tu[1] * tv[2] - tu[2] * tv[1];
      nrm[1] = tu[2] * tv[0] - tu[0] * tv[2];
      nrm[2] = tu[0] * tv[1] - tu[1] * tv[0];
      const double len = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
      w = len * wr;
      nrm[0] /= len; nrm[1] /= len; nrm[2] /= len;
    }
    g[0] = w; g[1] = x[0]; g[2] = x[1]; g[3] = x[2];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { g[4 + i * 3 + j] = Ji[i][j]; g[13 + i * 3 + j] = J[i][j]; }
    g[22] = det; g[24] = nrm[0]; g[25] = nrm[1]; g[26] = nrm[2];
  }

  // S1b: element size h = (sum_q w)^(1/dim) (getElementSize, workset.cpp:2699-2712; side: ^(1/(dim-1)), :2718-2733)
  MRH_HD static void s1b(const GenParams& P, double* sm, int /*blk*/, int idx) {
    if (idx >= P.epb) return;
    double vol = 0.0;
    for (int q = 0; q < NQ; ++q) vol += sm[idx * L::SIZE + L::G + q * L::GEO];
    const double h = pow(vol, 1.0 / (SIDE ? (double)DIM - 1.0 : (double)DIM));
    sm[idx * L::SIZE + L::H] = h;
    sm[idx * L::SIZE + L::H + 1] = 1.0 / h;
  }

  // S2: push-forward of basis b, function i, at point q
  template <int B>
  MRH_HD static void s2_basis(const GenParams& P, double* sm, int blk, int idx) {
    constexpr int CARD = Phys::card(B), NCB = Phys::ncb(B), BT = Phys::btype(B);
    if (idx >= P.epb * CARD * NQ) return;
    const int el = idx / (CARD * NQ), r = idx % (CARD * NQ), i = r / NQ, q = r % NQ;
    const double* g = sm + el * L::SIZE + L::G + q * L::GEO;
    const double* rt = P.ref_tab + L::rt_off(B) + (i * NQ + q) * NCB;
    double* out = sm + el * L::SIZE + L::PB + L::pb_off(B) + i * L::prs(B) + q * NCB;
    if (BT == BT_HGRAD) {
      out[0] = MRH_LDG(rt);
      for (int d = 0; d < 3; ++d) {
        double s = 0.0;
        for (int k = 0; k < DIM; ++k) s += g[4 + k * 3 + d] * MRH_LDG(rt + 1 + k);   // Jinv^T grad_ref
        out[1 + d] = (d < DIM) ? s : 0.0;
      }
    } else {
      double sg = 1.0;
      if (P.orient) {
        const int64_t item = item_of(P, blk, el);
        if (item < P.item_end) sg = (double)MRH_LDG(P.orient + (int64_t)elem_of(P, item) * N + P.off[Phys::first_var(B)][i]);
      }
      const double det = g[22];
      if (BT == BT_HCURL) {
        for (int d = 0; d < 3; ++d) {
          double s = 0.0, c = 0.0;
          for (int k = 0; k < 3; ++k) { s += g[4 + k * 3 + d] * MRH_LDG(rt + k); c += g[13 + d * 3 + k] * MRH_LDG(rt + 3 + k); }
          out[d] = sg * s; out[3 + d] = sg * c / det;
        }
      } else {  // HDIV
        for (int d = 0; d < 3; ++d) {
          double s = 0.0;
          for (int k = 0; k < 3; ++k) s += g[13 + d * 3 + k] * MRH_LDG(rt + k);
          out[d] = sg * s / det;
        }
        out[3] = sg * MRH_LDG(rt + 3) / det;
      }
    }
  }
  MRH_HD static void s2(const GenParams& P, double* sm, int blk, int idx) {
    s2_basis<0>(P, sm, blk, idx);
    if (Phys::NBASIS > 1) s2_basis<(Phys::NBASIS > 1 ? 1 : 0)>(P, sm, blk, idx);
  }

  // S3: fields and functions at (element, point)
  template <int V>
  MRH_HD static void s3_var(const GenParams& P, double* sme, int q) {
    constexpr int B = Phys::var_basis(V), CARD = Phys::card(B), NCB = Phys::ncb(B), NVAL = Phys::nval(B);
    double f[NC], ft[NC];
    for (int k = 0; k < NC; ++k) { f[k] = 0.0; ft[k] = 0.0; }
    const double* pb = sme + L::PB + L::pb_off(B) + q * NCB;
    for (int i = 0; i < CARD; ++i) {
      const int c = P.off[V][i];
      const double u = sme[L::U + c], ut = sme[L::UT + c];
      for (int k = 0; k < NCB; ++k) f[k] += u * pb[i * L::prs(B) + k];
      for (int k = 0; k < NVAL; ++k) ft[k] += ut * pb[i * L::prs(B) + k];
    }
    for (int k = 0; k < NC; ++k) { sme[L::FV + (q * NVAR + V) * NC + k] = f[k]; sme[L::FT + (q * NVAR + V) * NC + k] = ft[k]; }
  }
  // S3: work item (t, el, q) = fields of variable t at the point
  static constexpr int S3_KINDS = NVAR;
  MRH_HD static void s3(const GenParams& P, double* sm, int /*blk*/, int idx) {
    const int neq = P.epb * NQ;
    if (idx >= neq * S3_KINDS) return;
    const int t = idx / neq, eq = idx % neq, el = eq / NQ, q = eq % NQ;
    double* sme = sm + el * L::SIZE;
    switch (t) {
      case 0: s3_var<0>(P, sme, q); break;
      case 1: s3_var<(NVAR > 1 ? 1 : 0)>(P, sme, q); break;
      case 2: s3_var<(NVAR > 2 ? 2 : 0)>(P, sme, q); break;
      case 3: s3_var<(NVAR > 3 ? 3 : 0)>(P, sme, q); break;
      default: s3_var<(NVAR > 4 ? 4 : 0)>(P, sme, q); break;
    }
    static_assert(NVAR <= 5, "per-variable dispatch is written out for up to five variables");
  }
  // variables of the expression evaluator at a point: x y z t n[x] n[y] n[z], then (plans with state-dependent coefficients) the
  // solution fields F[v][k] and their time derivatives Ft[v][k] (slots of FunctionSet::set_solution_slots, abi.cu)
  static constexpr int NXV = EXPR_STATE0 + (STATEK ? 2 * NVAR * NC : 0);
  MRH_HD static void expr_vars(const GenParams& P, const double* sme, int q, double (&var)[NXV]) {
    const double* g = sme + L::G + q * L::GEO;
    var[0] = g[1]; var[1] = g[2]; var[2] = g[3]; var[3] = P.td.time; var[4] = g[24]; var[5] = g[25]; var[6] = g[26];
    if constexpr (STATEK) {
      for (int j = 0; j < NVAR * NC; ++j) {
        var[EXPR_STATE0 + j] = sme[L::FV + q * NVAR * NC + j];
        var[EXPR_STATE0 + NVAR * NC + j] = P.td.transient ? sme[L::FT + q * NVAR * NC + j] : 0.0;
      }
    }
  }
  struct OtherPoints {   // variables of point qq of the same element / side (element reductions)
    static constexpr bool kHasPoints = true;
    const GenParams* P; const double* sme;
    MRH_HD void operator()(int qq, double* vq) const { double t[NXV]; expr_vars(*P, sme, qq, t); for (int j = 0; j < NXV; ++j) vq[j] = t[j]; }
  };
  MRH_HD static const GenFnRec* fn_rec(const GenParams& P, int f) {
    if (f < Phys::NFN) return &P.fn[f];
    if (SIDE && P.bc_fn[f - Phys::NFN] >= 0) return &P.fn[P.bc_fn[f - Phys::NFN]];
    return nullptr;
  }
  // S3f (after the fields): work item (f, el, q) = coefficient function f at the point (the last NVAR are the boundary data of each
  // variable on this sideset)
  static constexpr int S3F_KINDS = Phys::NFN + NVAR;
  MRH_HD static void s3f(const GenParams& P, double* sm, int /*blk*/, int idx) {
    const int neq = P.epb * NQ;
    if (idx >= neq * S3F_KINDS) return;
    const int f = idx / neq, eq = idx % neq, el = eq / NQ, q = eq % NQ;
    double* sme = sm + el * L::SIZE;
    double var[NXV];
    expr_vars(P, sme, q, var);
    const GenFnRec* fr = fn_rec(P, f);
    double val = 0.0;
    if (fr) {
      if (fr->is_const) val = fr->cval;
      else if (fr->pad & 2) {   // element reduction inside: its argument is evaluated at every point of this element / side
        val = gen_expr_range<NXV>(P.fn_op, P.fn_c, fr->begin, fr->begin + fr->n, var, NQ, OtherPoints{&P, sme});
      } else val = gen_expr_range<NXV>(P.fn_op, P.fn_c, fr->begin, fr->begin + fr->n, var, 0, GenNoPoints());
    }
    sme[L::FN + q * L::NFN + f] = val;
  }
  // derivative components of the state-dependent coefficient functions at a point for the seeds dvar (others: 0)
  MRH_HD static void fn_derivatives(const GenParams& P, const double* sme, int q, const double (&var)[NXV], const double (&dvar)[NXV], double* dfn, int stride, int kk) {
    for (int f = 0; f < S3F_KINDS; ++f) {
      const GenFnRec* fr = fn_rec(P, f);
      double v = 0.0, d = 0.0;
      if (fr && (fr->pad & 1)) gen_expr_eval_dual(*fr, P.fn_op, P.fn_c, var, dvar, v, d);
      dfn[f * stride + kk] = d;
    }
  }

  MRH_HD static void make_ctx(const GenParams& P, const double* sme, int q, QpCtx& c) {
    const double* g = sme + L::G + q * L::GEO;
    c.w = g[0]; c.x = g[1]; c.y = g[2]; c.z = g[3]; c.t = P.td.time; c.h = sme[L::H]; c.ih = sme[L::H + 1];
    c.dt = P.td.deltat;
    c.n[0] = g[24]; c.n[1] = g[25]; c.n[2] = g[26];
    c.fn = sme + L::FN + q * L::NFN;
    c.dfn = nullptr;
    c.transient = P.td.transient; c.stage = P.td.nstage_lo;
    c.bc_type = P.bc_type;
  }

  // The module's point function on Dual<KK> when coefficient functions read solution fields: the derivative components of those
  // functions follow from the seeds of F / Ft (bytecode evaluated on (value, derivative) pairs).  It rebuilds the seeded fields itself
  // (seed_var < 0: direction seed_idx of the tensor-core build; else the KK columns i0.. of variable seed_var of the derivative-lane
  // build); only the STATEK builds of the kernel contain it.
  template <int KK>
  MRH_HD static void point_state(const GenParams& P, const double* sme, int q, int seed_var, int seed_idx, Dual<KK> (&Cf)[NVAR][NC]) {
    QpCtx c;
    make_ctx(P, sme, q, c);
    const bool transient = P.td.transient != 0;
    const double au = transient ? P.td.seed_u : 1.0, at = transient ? P.td.seed_t : 0.0;   // derivative seeds (TimeDev)
    Dual<KK> F[NVAR][NC], Ft[NVAR][NC];
    for (int v = 0; v < NVAR; ++v)
      for (int k = 0; k < NC; ++k) {
        F[v][k] = Dual<KK>(sme[L::FV + (q * NVAR + v) * NC + k]);
        Ft[v][k] = Dual<KK>(transient ? sme[L::FT + (q * NVAR + v) * NC + k] : 0.0);
        Cf[v][k] = Dual<KK>(0.0);
      }
    if (seed_var < 0) {
      const int v = seed_idx / NC, k = seed_idx % NC;
      F[v][k].d[0] = au;
      if (k < Phys::nval(Phys::var_basis_rt(v))) Ft[v][k].d[0] = at;
    } else {
      const int wb = Phys::var_basis_rt(seed_var), ncb = Phys::ncb_rt(wb), nval = Phys::nval(wb);
      const double* pbw = sme + L::PB + L::pb_off(wb) + seed_idx * L::prs_rt(wb);
      for (int kk = 0; kk < KK; ++kk)
        for (int k = 0; k < ncb; ++k) {
          const double sd = pbw[kk * L::prs_rt(wb) + q * ncb + k];
          F[seed_var][k].d[kk] = au * sd;
          if (k < nval) Ft[seed_var][k].d[kk] = at * sd;
        }
    }
    double var[NXV], dvar[NXV], dfn[S3F_KINDS * KK];
    expr_vars(P, sme, q, var);
    for (int kk = 0; kk < KK; ++kk) {
      for (int j = 0; j < EXPR_STATE0; ++j) dvar[j] = 0.0;
      for (int v = 0; v < NVAR; ++v)
        for (int k = 0; k < NC; ++k) {
          dvar[EXPR_STATE0 + v * NC + k] = F[v][k].d[kk];
          dvar[EXPR_STATE0 + NVAR * NC + v * NC + k] = Ft[v][k].d[kk];
        }
      fn_derivatives(P, sme, q, var, dvar, dfn, KK, kk);
    }
    c.dfn = dfn;
    if (SIDE) Phys::template boundary<Dual<KK>, true>(c, P.opt, F, Ft, Cf);
    else Phys::template volume<Dual<KK>, true>(c, P.opt, F, Ft, Cf);
  }

  // S4a: coefficient values
  MRH_HD static void s4a(const GenParams& P, double* sm, int /*blk*/, int idx) {
    if (idx >= P.epb * NQ) return;
    const int el = idx / NQ, q = idx % NQ;
    double* sme = sm + el * L::SIZE;
    QpCtx c;
    make_ctx(P, sme, q, c);
    double F[NVAR][NC], Ft[NVAR][NC], Cf[NVAR][NC];
    for (int v = 0; v < NVAR; ++v)
      for (int k = 0; k < NC; ++k) { F[v][k] = sme[L::FV + (q * NVAR + v) * NC + k]; Ft[v][k] = sme[L::FT + (q * NVAR + v) * NC + k]; Cf[v][k] = 0.0; }
    if (P.mass_mode == 2) gen_initial_point<Phys, double>(c, Cf);
    else if (P.mass_mode) gen_mass_point<Phys, double>(c, P.mass_wts, F, Cf);
    else if (SIDE) Phys::template boundary<double, false>(c, P.opt, F, Ft, Cf);
    else Phys::template volume<double, false>(c, P.opt, F, Ft, Cf);
    for (int v = 0; v < NVAR; ++v)
      for (int k = 0; k < NC; ++k) sme[L::CV + (q * NVAR + v) * NC + k] = Cf[v][k];
  }

  // S4b: derivative components.  Thread (el, g) owns columns (wv, i0 .. i0+K-1); the variable wv is a run-time value so
  // that the lanes of a warp -- which hold different variables of one element -- execute one instruction stream.
  template <int V>
  MRH_HD static void store_var(const GenParams& P, double* out, int col, const double (&acc)[N][K], int kk) {
    constexpr int CARD = Phys::card(Phys::var_basis(V)), R0 = Phys::row0(V);
#pragma unroll
    for (int i = 0; i < CARD; ++i) out[P.adjoint ? col * N + (int)P.off[V][i] : (int)P.off[V][i] * N + col] = acc[R0 + i][kk];
  }
  MRH_HD static void store_rows(const GenParams& P, double* out, int col, const double (&acc)[N][K], int kk) {
    store_var<0>(P, out, col, acc, kk);
    if (NVAR > 1) store_var<(NVAR > 1 ? 1 : 0)>(P, out, col, acc, kk);
    if (NVAR > 2) store_var<(NVAR > 2 ? 2 : 0)>(P, out, col, acc, kk);
    if (NVAR > 3) store_var<(NVAR > 3 ? 3 : 0)>(P, out, col, acc, kk);
    if (NVAR > 4) store_var<(NVAR > 4 ? 4 : 0)>(P, out, col, acc, kk);
  }
  template <int V>
  MRH_HD static void rows_var(const double* __restrict__ pbase, int q, const Dual<K> (&Cf)[NVAR][NC], double (&acc)[N][K]) {
    constexpr int B = Phys::var_basis(V), CARD = Phys::card(B), NCB = Phys::ncb(B), R0 = Phys::row0(V);
    const double* pbq = pbase + L::pb_off(B) + q * NCB;
#pragma unroll
    for (int i = 0; i < CARD; ++i) {
      double b[NCB];
      mrh_ldn<NCB>(pbq + i * L::prs(B), b);
#pragma unroll
      for (int k = 0; k < NCB; ++k) {
#pragma unroll
        for (int kk = 0; kk < K; ++kk) acc[R0 + i][kk] += Cf[V][k].d[kk] * b[k];
      }
    }
  }
  MRH_HD static void s4b_rows_all(const double* __restrict__ pbase, int q, const Dual<K> (&Cf)[NVAR][NC], double (&acc)[N][K]) {
    rows_var<0>(pbase, q, Cf, acc);
    if (NVAR > 1) rows_var<(NVAR > 1 ? 1 : 0)>(pbase, q, Cf, acc);
    if (NVAR > 2) rows_var<(NVAR > 2 ? 2 : 0)>(pbase, q, Cf, acc);
    if (NVAR > 3) rows_var<(NVAR > 3 ? 3 : 0)>(pbase, q, Cf, acc);
    if (NVAR > 4) rows_var<(NVAR > 4 ? 4 : 0)>(pbase, q, Cf, acc);
  }
  MRH_HD static void s4b(const GenParams& P, double* sm, int blk, int idx) {
    if (idx >= P.epb * TPE) return;
    const int el = idx / TPE, g = idx % TPE;
    // groups are numbered variable-major: variable v owns card(v)/K consecutive groups
    int wv = 0, g0 = 0;
    for (; wv < NVAR - 1; ++wv) {
      const int ng = Phys::card_of_var(wv) / K;
      if (g < g0 + ng) break;
      g0 += ng;
    }
    const int i0 = (g - g0) * K;
    const int wb = Phys::var_basis_rt(wv), ncb = Phys::ncb_rt(wb), nval = Phys::nval(wb);
    const int64_t item = item_of(P, blk, el);
    const double* sme = sm + el * L::SIZE;
    const double* pbw = sme + L::PB + L::pb_off(wb) + i0 * L::prs_rt(wb);
    double acc[N][K];
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
      for (int kk = 0; kk < K; ++kk) acc[r][kk] = 0.0;
    const bool transient = P.td.transient != 0;
    const double au = transient ? P.td.seed_u : 1.0, at = transient ? P.td.seed_t : 0.0;   // derivative seeds (TimeDev)
    for (int q = 0; q < NQ; ++q) {
      QpCtx c;
      make_ctx(P, sme, q, c);
      Dual<K> F[NVAR][NC], Ft[NVAR][NC], Cf[NVAR][NC];
      {
        double fv[NVAR * NC], ft[NVAR * NC];
        mrh_ldn<NVAR * NC>(sme + L::FV + q * NVAR * NC, fv);
        if (transient) mrh_ldn<NVAR * NC>(sme + L::FT + q * NVAR * NC, ft);
#pragma unroll
        for (int v = 0; v < NVAR; ++v)
#pragma unroll
          for (int k = 0; k < NC; ++k) {
            F[v][k].v = fv[v * NC + k];
            Ft[v][k].v = transient ? ft[v * NC + k] : 0.0;
            Cf[v][k] = Dual<K>(0.0);
          }
      }
#pragma unroll
      for (int kk = 0; kk < K; ++kk) {
        double sd[NC];
        const double* pb = pbw + kk * L::prs_rt(wb) + q * ncb;
        if (Phys::NBASIS == 1) mrh_ldn<NC>(pb, sd);
        else {
#pragma unroll
          for (int k = 0; k < NC; ++k) sd[k] = k < ncb ? pb[k] : 0.0;
        }
#pragma unroll
        for (int v = 0; v < NVAR; ++v)
#pragma unroll
          for (int k = 0; k < NC; ++k) {
            F[v][k].d[kk] = (v == wv) ? au * sd[k] : 0.0;
            Ft[v][k].d[kk] = (v == wv && k < nval) ? at * sd[k] : 0.0;
          }
      }
      if (P.mass_mode) gen_mass_point<Phys, Dual<K>>(c, P.mass_wts, F, Cf);
      else if (STATEK) point_state<K>(P, sme, q, wv, i0, Cf);   // coefficients that read solution fields (STATEK builds only)
      else if (SIDE) Phys::template boundary<Dual<K>, false>(c, P.opt, F, Ft, Cf);
      else Phys::template volume<Dual<K>, false>(c, P.opt, F, Ft, Cf);
      // test-function loop: rows in (variable, basis function) order
      s4b_rows_all(sme + L::PB, q, Cf, acc);
    }
    if (item < P.item_end && P.elem_jac) {
      double* out = P.elem_jac + (P.inst_base + (item - P.item_begin)) * (int64_t)(N * N);
#pragma unroll
      for (int kk = 0; kk < K; ++kk) store_rows(P, out, (int)P.off[wv][i0 + kk], acc, kk);
    }
  }

  // ---- tensor-core form of the Jacobian (single-basis HGRAD modules; north_star: "fp64 DMMA for the per-element B^T D B") ----
  // The AD derivative lanes run over FIELD DIRECTIONS instead of element dofs: one evaluation of the module's point function
  // on Dual<1> per (point, direction (w,l)) gives column (w,l) of  D_q = alpha_u dCf/dF + alpha_t dCf/dF_t  (NVAR NC columns
  // instead of N), and the element matrix is the contraction
  //     J[(v,i),(w,j)] = sum_q sum_k sum_l PB[i][q][k] D_q[(v,k),(w,l)] PB[j][q][l]
  // done per variable pair (v,w) as a GEMM on the FP64 tensor cores: A[i][(q,k)] = PB (already laid out that way in shared
  // memory), B[(q,k)][j] = sum_l D_q[(v,k),(w,l)] PB[j][q][l] built in registers, one mma.m8n8k4 k-step per point.
  static constexpr bool TC = L::TC;
  static constexpr int NCV = L::NCV, CARD = Phys::card(0), IT = (CARD + 7) / 8;
  MRH_HD static int row_index(const GenParams& P, int v, int i) { return P.var_major ? v * CARD + i : (int)P.off[v][i]; }

  // S4d: column `dir` of D at (element, point)
  MRH_HD static void s4d(const GenParams& P, double* sm, int /*blk*/, int idx) {
    if (idx >= P.epb * NQ * NCV) return;
    const int el = idx / (NQ * NCV), r = idx % (NQ * NCV), q = r / NCV, dir = r % NCV;
    double* sme = sm + el * L::SIZE;
    QpCtx c;
    make_ctx(P, sme, q, c);
    const bool transient = P.td.transient != 0;
    const double au = transient ? P.td.seed_u : 1.0, at = transient ? P.td.seed_t : 0.0;   // derivative seeds (TimeDev)
    Dual<1> F[NVAR][NC], Ft[NVAR][NC], Cf[NVAR][NC];
#pragma unroll
    for (int v = 0; v < NVAR; ++v)
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        const bool mine = (v * NC + k) == dir;
        F[v][k].v = sme[L::FV + (q * NVAR + v) * NC + k];
        F[v][k].d[0] = mine ? au : 0.0;
        Ft[v][k].v = transient ? sme[L::FT + (q * NVAR + v) * NC + k] : 0.0;
        Ft[v][k].d[0] = (mine && k < Phys::nval(0)) ? at : 0.0;
        Cf[v][k] = Dual<1>(0.0);
      }
    if (P.mass_mode) gen_mass_point<Phys, Dual<1>>(c, P.mass_wts, F, Cf);
    else if (STATEK) point_state<1>(P, sme, q, -1, dir, Cf);
    else if (SIDE) Phys::template boundary<Dual<1>, false>(c, P.opt, F, Ft, Cf);
    else Phys::template volume<Dual<1>, false>(c, P.opt, F, Ft, Cf);
    double* D = sme + L::DM + q * NCV * L::DS;
#pragma unroll
    for (int v = 0; v < NVAR; ++v)
#pragma unroll
      for (int k = 0; k < NC; ++k) D[(v * NC + k) * L::DS + dir] = Cf[v][k].d[0];
  }

#if defined(__CUDA_ARCH__)
  // S4m: one warp, one (element, v, w) block of the element matrix -- or, for bases of more than 24 functions, one half of its
  // column tiles (JSPLIT = 2), so that a warp's accumulators stay within 32 registers and twice as many warps share the work
  static constexpr int JSPLIT = IT >= 4 ? 2 : 1, JT = IT / JSPLIT;
  static constexpr int S4M_ITEMS = NVAR * NVAR * JSPLIT;   // per element
  __device__ __forceinline__ static void s4m_warp(const GenParams& P, const double* sm, int blk, int item, int lane) {
    const int el = item / S4M_ITEMS, r = item % S4M_ITEMS, vw = r / JSPLIT, jh = r % JSPLIT, v = vw / NVAR, w = vw % NVAR;
    const double* sme = sm + el * L::SIZE;
    const double* pb = sme + L::PB;
    const int g = lane >> 2, t = lane & 3;
    const double* D = sme + L::DM + (v * NC + t) * L::DS + w * NC;
    double acc[IT][JT][2];
    int ro[IT];
#pragma unroll
    for (int a = 0; a < IT; ++a) {
      ro[a] = ((a * 8 + g) < CARD ? (a * 8 + g) : (CARD - 1)) * L::prs(0);   // rows beyond the basis repeat the last one (their results are dropped)
#pragma unroll
      for (int b = 0; b < JT; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }
    }
#pragma unroll 2
    for (int q = 0; q < NQ; ++q) {
      double d[NC], fa[IT], fb[JT];
      mrh_ldn<NC>(D + q * NCV * L::DS, d);
#pragma unroll
      for (int a = 0; a < IT; ++a) fa[a] = pb[ro[a] + q * NC + t];     // A fragment: entry t of basis row a * 8 + g
#pragma unroll
      for (int b = 0; b < JT; ++b) {                                  // B fragment: (D_q^{vw} PB_j)[t] for basis column (jh JT + b) * 8 + g
        double pr[NC];
        mrh_ldn<NC>(pb + ro[jh * JT + b] + q * NC, pr);
        fb[b] = d[0] * pr[0] + d[1] * pr[1] + d[2] * pr[2] + d[3] * pr[3];
      }
#pragma unroll
      for (int a = 0; a < IT; ++a)
#pragma unroll
        for (int b = 0; b < JT; ++b) mrh_dmma(acc[a][b], fa[a], fb[b]);
    }
    const int64_t it = item_of(P, blk, el);
    if (it < P.item_end && P.elem_jac) {
      double* out = P.elem_jac + (P.inst_base + (it - P.item_begin)) * (int64_t)(N * N);
#pragma unroll
      for (int a = 0; a < IT; ++a) {
        const int i = a * 8 + g;
        if (i < CARD) {
          const int ri = row_index(P, v, i);
#pragma unroll
          for (int b = 0; b < JT; ++b)
#pragma unroll
            for (int x = 0; x < 2; ++x) {
              const int j = (jh * JT + b) * 8 + 2 * t + x;
              if (j < CARD) { const int cj = row_index(P, w, j); out[P.adjoint ? (int64_t)cj * N + ri : (int64_t)ri * N + cj] = acc[a][b][x]; }
            }
        }
      }
    }
  }
#endif
  // host replay of S4m (one work item = one (element, v, w) block, plain loops)
  MRH_HD static void s4m_item(const GenParams& P, const double* sm, int blk, int item) {
    if (item >= P.epb * NVAR * NVAR) return;
    const int el = item / (NVAR * NVAR), vw = item % (NVAR * NVAR), v = vw / NVAR, w = vw % NVAR;
    const int64_t it = item_of(P, blk, el);
    if (it >= P.item_end || !P.elem_jac) return;
    const double* sme = sm + el * L::SIZE;
    const double* pb = sme + L::PB;
    double* out = P.elem_jac + (P.inst_base + (it - P.item_begin)) * (int64_t)(N * N);
    for (int i = 0; i < CARD; ++i)
      for (int j = 0; j < CARD; ++j) {
        double s = 0.0;
        for (int q = 0; q < NQ; ++q) {
          const double* D = sme + L::DM + q * NCV * L::DS;
          for (int k = 0; k < NC; ++k) {
            double b = 0.0;
            for (int l = 0; l < NC; ++l) b += D[(v * NC + k) * L::DS + w * NC + l] * pb[j * L::prs(0) + q * NC + l];
            s += pb[i * L::prs(0) + q * NC + k] * b;
          }
        }
        const int ri = row_index(P, v, i), cj = row_index(P, w, j);
        out[P.adjoint ? (int64_t)cj * N + ri : (int64_t)ri * N + cj] = s;
      }
  }

  // S5: residual rows
  MRH_HD static void s5(const GenParams& P, double* sm, int blk, int idx) {
    if (idx >= P.epb * N || !P.elem_res) return;
    const int el = idx / N, r = idx % N;
    const int64_t item = item_of(P, blk, el);
    if (item >= P.item_end) return;
    // r is numbered variable-major (v, i); the element-local dof is off[v][i]
    int v = 0, r0 = 0;
    for (; v < NVAR - 1; ++v) {
      if (r < r0 + Phys::card_of_var(v)) break;
      r0 += Phys::card_of_var(v);
    }
    const int i = r - r0;
    const int b = Phys::var_basis_rt(v), ncb = Phys::ncb_rt(b);
    const double* sme = sm + el * L::SIZE;
    const double* pb = sme + L::PB + L::pb_off(b) + i * L::prs_rt(b);
    double s = 0.0;
    for (int q = 0; q < NQ; ++q)
      for (int k = 0; k < ncb; ++k) s += sme[L::CV + (q * NVAR + v) * NC + k] * pb[q * ncb + k];
    P.elem_res[(P.inst_base + (item - P.item_begin)) * (int64_t)N + (P.var_major ? r : (int)P.off[v][i])] = s;
  }
};

#if defined(__CUDACC__)
// MAXT / MINB: launch bounds chosen per instantiation (general_dispatch.hpp) so that the register allocation leaves
// MINB CTAs of MAXT threads resident per SM
// TCK: Jacobian by field-direction derivatives + tensor-core contraction (S4d / S4m; single-basis layouts only), else by one
// derivative lane per element dof (S4b)
template <class Phys, int NQ, int K, bool SIDE, int MAXT, int MINB, bool TCK, bool STATEK>
__global__ void __launch_bounds__(MAXT, MINB) gen_element_kernel(const __grid_constant__ GenParams P) {
  extern __shared__ __align__(16) double gen_smem[];
  typedef GenBlock<Phys, NQ, K, SIDE, TCK, STATEK> Bk;
  const int blk = blockIdx.x, T = blockDim.x, tid = threadIdx.x;
  typedef GenLayout<Phys, NQ, TCK> L;
  for (int i = tid; i < P.epb * (L::N > L::NV ? L::N : L::NV); i += T) Bk::s0(P, gen_smem, blk, i);
  __syncthreads();
  for (int i = tid; i < P.epb * NQ; i += T) Bk::s1(P, gen_smem, blk, i);
  __syncthreads();
  for (int i = tid; i < P.epb; i += T) Bk::s1b(P, gen_smem, blk, i);
  for (int i = tid; i < P.epb * Phys::max_card() * NQ; i += T) Bk::s2(P, gen_smem, blk, i);
  __syncthreads();
  for (int i = tid; i < P.epb * NQ * Bk::S3_KINDS; i += T) Bk::s3(P, gen_smem, blk, i);
  if (STATEK) __syncthreads();   // coefficient functions that read solution fields wait for the fields
  for (int i = tid; i < P.epb * NQ * Bk::S3F_KINDS; i += T) Bk::s3f(P, gen_smem, blk, i);
  __syncthreads();
  for (int i = tid; i < P.epb * NQ; i += T) Bk::s4a(P, gen_smem, blk, i);
  if constexpr (TCK && Bk::TC) {
    if (P.elem_jac)
      for (int i = tid; i < P.epb * NQ * Bk::NCV; i += T) Bk::s4d(P, gen_smem, blk, i);
    __syncthreads();
    if (P.elem_jac)
      for (int item = tid >> 5; item < P.epb * Bk::S4M_ITEMS; item += T >> 5) Bk::s4m_warp(P, gen_smem, blk, item, tid & 31);
  } else {
    if (P.elem_jac)
      for (int i = tid; i < P.epb * Bk::TPE; i += T) Bk::s4b(P, gen_smem, blk, i);
    __syncthreads();
  }
  for (int i = tid; i < P.epb * L::N; i += T) Bk::s5(P, gen_smem, blk, i);
}
#endif

}  // namespace mrhyde_b200
