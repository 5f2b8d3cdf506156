// C ABI of the assembly library (include/mrhyde_b200.h): plan object, option/function registry,
// upload of the plan to the device, kernel dispatch.  No CPU fallback exists: if no CUDA device or
// no kernel matches the described block the call fails with an error code.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/mrhyde_b200.h"
#include <cstdio>
#include <cstdlib>
#include <set>

#include "boundary.cuh"
#include "expr.hpp"
#include "general.hpp"
#include "halo.hpp"
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3 (the injection library is looked up at run time; no-ops without a profiler)
#include "plan.hpp"
#include "volume_kernel.cuh"
#include "volume_launch.hpp"
#include "build/jit_sources.h"

using namespace mrhyde_b200;

namespace {

thread_local std::string g_error;

struct AbiError {
  int code;
  std::string msg;
};
[[noreturn]] void fail(int code, const std::string& msg) { throw AbiError{code, msg}; }

#define CUDA_OK(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (expr);                                                                            \
    if (e_ != cudaSuccess) fail(MRHYDE_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

#define ABI_BEGIN try {
#define ABI_END                                                                  \
  return MRHYDE_B200_OK;                                                         \
  }                                                                              \
  catch (const AbiError& e) { g_error = e.msg; return e.code; }                  \
  catch (const ExprError& e) { g_error = e.what(); return e.code; }              \
  catch (const std::exception& e) { g_error = e.what(); return MRHYDE_B200_ERR_INVALID; } \
  catch (...) { g_error = "unknown failure"; return MRHYDE_B200_ERR_INVALID; }

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  ~DevBuf() { if (p) cudaFree(p); }
  void upload(const std::vector<T>& h, size_t* total) {
    if (p) { cudaFree(p); p = nullptr; }
    n = h.size();
    const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    CUDA_OK(cudaMalloc(&p, bytes));
    if (n) CUDA_OK(cudaMemcpy(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice));
    if (total) *total += bytes;
  }
  void resize(size_t count, size_t* total) {
    if (count <= n && p) return;
    if (p) { cudaFree(p); p = nullptr; }
    n = count;
    CUDA_OK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    if (total) *total += n * sizeof(T);
  }
};

struct BasisCopy {
  std::string type;
  int order = 1, card = 0;
  std::vector<double> val, grad, curl, div;
};

struct BCEntry {
  std::string type = "none", expr = "0.0";
};

}  // namespace

struct mrhyde_b200_plan {
  int device = 0;
  // description
  std::string physics;
  int dim = 3, nvars = 0, ndof_elem = 0, max_card = 0, nqp = 0;
  std::vector<std::string> var_names;
  std::vector<int> var_basis;
  std::vector<BasisCopy> bases;
  std::vector<int32_t> offsets;
  std::vector<double> qp_pts, qp_wts;
  // registry
  std::map<std::string, std::string> functions, options;
  std::vector<std::string> side_names;
  std::map<std::pair<std::string, std::string>, BCEntry> bcs;  // (var, side)
  std::vector<BoundaryGroupHost> bgroups;
  // mesh / graph
  MeshGraph mesh;
  bool have_mesh = false, have_graph = false, finalized = false;
  ChainPlan cp;
  // device copies
  size_t dev_bytes = 0;
  DevBuf<double> d_vx, d_vy, d_vz;
  DevBuf<int32_t> d_conn, d_lids, d_colind, d_chain_step_ptr, d_step_elems, d_step_conn, d_step_lids, d_orphans;
  DevBuf<int64_t> d_rowptr, d_fixed_diag;
  DevBuf<int64_t> d_point_rows;       // point constraints: (first entry, end, diagonal entry) per constrained row
  std::vector<int32_t> point_dofs;
  bool point_on_ghost = false;        // a ghost row is point-constrained: its identity row must be what the halo sum sends (no in-kernel push)
  DevBuf<uint8_t> d_fixed, d_eclass, d_step_eclass, d_chain_invariant;
  DevBuf<StepRec> d_steps;
  DevBuf<BatchRec> d_batches;
  DevBuf<RowRec> d_rows;
  DevBuf<uint32_t> d_desc0, d_desc1, d_mdesc0, d_mdesc1;
  // plan-specialised (NVRTC) volume kernel; falls back to the ahead-of-time kernel when absent
  std::map<int, std::unique_ptr<JitKernel>> jit;   // specialised builds keyed by transient * 8 + output mode, built on first use
  bool use_jit = false;
  std::string jit_note;   // why the plan is not specialised (empty when it is)
  std::string jit_source; // translation unit handed to NVRTC
  // host-buffer entry point scratch
  DevBuf<double> h_sol, h_res, h_jac;
  std::vector<DevBuf<double>> h_prev, h_stage;
  // kernels
  ThermalParams<2> th2;
  ThermalParams<3> th3;
  BoundaryPlan boundary;
  int threads = 256;
  int row_tab = 0;
  int jit_min_blocks = 1;
  size_t smem = 0;
  bool suppress_overlap = false;
  bool push_ok = false;            // the specialised kernels can push ghost rows into the owner's slab themselves (in-kernel halo push)
  int64_t pushed_assembles = 0;    // assemble calls that did
  int64_t overlapped_assembles = 0;   // assemble calls that started the halo exchange after the ghost-row chains
  int metric_ng = 0;               // > 0: the specialised kernels use the metric ring with this many metric entries per element
  int class_nc = 0;                // > 0: class ring (boxes + constant coefficients): distinct local-matrix values staged per element
  std::vector<int32_t> class_of_t, class_rep;   // upper-triangle entry -> class, class -> representative entry
  size_t smem_metric[2] = {0, 0};  // dynamic shared memory of the steady / transient metric builds
  int64_t n_affine = 0, n_box = 0;
  int launches_per_assemble = 0;   // kernels launched by the last assemble call
  bool accumulate = true;
  int stage_len = 0;
  std::vector<uint16_t> kmap, rmap;
  // timing of the volume kernel
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
  size_t ev_next = 0, ev_used = 0;
  double ev_ms = 0.0;
  int64_t ev_count = 0;
  // halo
  std::unique_ptr<HaloExchange> halo;
  // general path (any module / basis): element kernel + deterministic pull (general.hpp)
  bool use_general = false;
  GeneralPlanHost gen;
  GeneralPlanDev* gen_dev = nullptr;
  const GenDeviceKernels* gen_kernels = nullptr;
  const GenHostKernels* gen_host = nullptr;

  ~mrhyde_b200_plan() {
    if (gen_dev) gen_free(gen_dev);
    for (auto& p : ev) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  }
};

namespace {

// NVTX ranges named like the reference's Teuchos timers (assemblyManager.hpp:2183-2198, linearAlgebraInterface.hpp:640-650), so that a
// timeline of a MrHyDE run with this library reads like the reference's `print timers` table
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

const char* kKnownOptions[] = {"accumulate", "use strong DBCs", "assemble boundary terms", "assemble volume terms", "form_param", "include advection",
                               "ns3d_uz_rows", "useSUPG", "usePSPG", "penalty", "incplanestress", "kernel", "batch elems", "elements per cta", "column elements", "min chains", "min segment levels", "cta slots", "sweep axis", "threads", "jit", "use leap frog", "ring", "pull patterns", "min blocks", "max registers", "pull group", "debug skip", "max blocks", "stage1", "stagger ns", "flush", "flush unroll", "stage2", "tables", "overlap halo", "halo transport", "pipeline", "store hint", "prefetch", "jacobian", "halo push", "lump mass", "prefetch records", "column cache", "fix zero rows", "scratch GB", "debug transient", "debug mode", nullptr};

std::string opt(const mrhyde_b200_plan* P, const std::string& key, const std::string& def) {
  auto it = P->options.find(key);
  return it == P->options.end() ? def : it->second;
}
bool opt_bool(const mrhyde_b200_plan* P, const std::string& key, bool def) {
  const std::string v = opt(P, key, def ? "true" : "false");
  return v == "true" || v == "True" || v == "1";
}
// Axes whose one-coordinate sub-expressions stay the same along most chains of the plan (an extruded column changes the sweep coordinate only):
// the generated source function keeps their values per thread ("column cache", default on).
static bool all_boxes(const mrhyde_b200_plan* P) {   // every cell of the plan is an axis-aligned box (class 2)
  for (uint8_t c : P->cp.step_eclass) if (c != 2) return false;
  return !P->cp.step_eclass.empty();
}
int column_cache_axes(const mrhyde_b200_plan* P) {
  const std::string mode = opt(P, "column cache", "true");   // true | false | registers (no CTA-shared values)
  if ((mode != "registers" && !opt_bool(P, "column cache", true)) || P->cp.chain_invariant.empty()) return 0;
  int axes = 0;
  for (int d = 0; d < P->dim; ++d) {
    size_t n = 0, u = 0;
    for (uint8_t f : P->cp.chain_invariant) { n += (f >> d) & 1; u += (f >> (4 + d)) & 1; }
    if (2 * n > P->cp.chain_invariant.size()) axes |= 1 << d;
    // bits 4-6: not step-invariant, but the same for all elements of a step on most chains; only plans made of axis-aligned boxes
    else if (mode != "registers" && 2 * u > P->cp.chain_invariant.size() && all_boxes(P)) axes |= 16 << d;
  }
  return axes;
}
int prefetch_mode(const mrhyde_b200_plan* P) {   // "prefetch": true | false | lean (first vertex / dof of an element only)
  if (opt(P, "prefetch", "true") == "lean") return 2;
  return opt_bool(P, "prefetch", true) ? 1 : 0;
}

// Hex8 / Quad4 nodal shape functions in Shards vertex order (CellTools::setJacobian uses the cell's
// HGRAD C1 basis; discretizationInterface_basis.hpp:244-252, 407-413)
void geom_shape(int dim, const double* xi, double* N, double* dN) {
  static const double sg[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};
  const int nv = 1 << dim;
  for (int n = 0; n < nv; ++n) {
    double f[3] = {1, 1, 1}, df[3] = {0, 0, 0};
    for (int d = 0; d < dim; ++d) { f[d] = 0.5 * (1.0 + sg[n][d] * xi[d]); df[d] = 0.5 * sg[n][d]; }
    N[n] = f[0] * f[1] * (dim == 3 ? f[2] : 1.0);
    for (int d = 0; d < dim; ++d) {
      double g = df[d];
      for (int o = 0; o < dim; ++o) if (o != d) g *= f[o];
      dN[n * dim + d] = g;
    }
  }
}

template <int DIM>
void fill_thermal_tables(const mrhyde_b200_plan* P, ThermalTables<DIM>& T) {
  typedef Q1Shape<DIM> S;
  const BasisCopy& B = P->bases[P->var_basis[0]];
  std::memset(&T, 0, sizeof(T));
  for (int q = 0; q < S::NQ; ++q) {
    double N[8], dN[24];
    geom_shape(DIM, &P->qp_pts[(size_t)q * DIM], N, dN);
    T.qw[q] = P->qp_wts[q];
    for (int d = 0; d < DIM; ++d) T.qpt[q][d] = P->qp_pts[(size_t)q * DIM + d];
    for (int n = 0; n < S::NV; ++n) {
      T.gN[q][n] = N[n];
      T.phi[q][n] = B.val[(size_t)n * S::NQ + q];
      for (int d = 0; d < DIM; ++d) {
        T.gdN[q][n][d] = dN[n * DIM + d];
        T.dphi[q][n][d] = B.grad[((size_t)n * S::NQ + q) * DIM + d];
      }
    }
  }
  int pa[S::NG], pb[S::NG], g = 0;
  for (int a = 0; a < DIM; ++a) { pa[g] = a; pb[g] = a; ++g; }
  for (int a = 0; a < DIM; ++a) for (int b = a + 1; b < DIM; ++b) { pa[g] = a; pb[g] = b; ++g; }
  for (int i = 0; i < S::NV; ++i)
    for (int j = i; j < S::NV; ++j) {
      const int t = i * S::NV - (i * (i - 1)) / 2 + (j - i);
      double m = 0.0;
      for (int q = 0; q < S::NQ; ++q) m += T.qw[q] * T.phi[q][i] * T.phi[q][j];
      if (j == i) { double l = 0.0; for (int q = 0; q < S::NQ; ++q) l += T.qw[q] * T.phi[q][i]; T.Ltab[i] = l; }
      T.Mtab[t] = m;
      for (int k = 0; k < S::NG; ++k) {
        double s = 0.0;
        for (int q = 0; q < S::NQ; ++q) {
          if (pa[k] == pb[k]) s += T.qw[q] * T.dphi[q][i][pa[k]] * T.dphi[q][j][pa[k]];
          else s += T.qw[q] * (T.dphi[q][i][pa[k]] * T.dphi[q][j][pb[k]] + T.dphi[q][i][pb[k]] * T.dphi[q][j][pa[k]]);
        }
        T.Stab[k][t] = s;
      }
    }
  // The table entries are small rationals evaluated by quadrature: entries that agree to rounding are made identical and
  // entries that are zero to rounding exactly zero, so that equal coefficients share a register in the generated kernels
  // and the box path drops only exact zeros.
  auto snap = [](double* v, size_t n) {
    double vmax = 0.0;
    for (size_t i = 0; i < n; ++i) vmax = std::max(vmax, std::fabs(v[i]));
    for (size_t i = 0; i < n; ++i) {
      if (std::fabs(v[i]) <= 1e-13 * vmax) { v[i] = 0.0; continue; }
      for (size_t j = 0; j < i; ++j) {
        if (std::fabs(v[i] - v[j]) <= 1e-13 * std::fabs(v[j])) { v[i] = v[j]; break; }
        if (std::fabs(v[i] + v[j]) <= 1e-13 * std::fabs(v[j])) { v[i] = -v[j]; break; }
      }
    }
  };
  snap(&T.Stab[0][0], (size_t)S::NG * S::NT);
  snap(&T.Mtab[0], S::NT);
  snap(&T.Ltab[0], S::NV);
  snap(&T.phi[0][0], (size_t)S::NQ * S::NV);
  snap(&T.qw[0], S::NQ);
}

// ---- source of the plan-specialised kernel: prelude (constants + generated coefficient functions) + embedded headers
std::string hexd(double v) {
  char b[64];
  std::snprintf(b, sizeof(b), "%a", v);
  return b;
}
void emit_values(std::string& o, const double* v, size_t n) {
  for (size_t i = 0; i < n; ++i) { o += hexd(v[i]); o += (i + 1 < n) ? "," : ""; }
}
std::string pull_codegen(const ChainPlan& cp, int max_patterns, int group, int flush_mode);
std::string pull_codegen_metric(const ChainPlan& cp, int max_patterns, int nv, int ng_all, int ngu, const double* Stab, const double* Mtab, int group, int flush_mode);
template <int DIM>
std::string thermal_jit_source(const ThermalTables<DIM>& T, const FunctionSet& fs, int all_const, int source_const, const ChainPlan& cp,
                               const int64_t (&n_class)[3], int metric_ng, int max_patterns, int pull_group,
                               const std::vector<int32_t>& class_of_t, const std::vector<int32_t>& class_rep, int debug_skip, bool late_stage1, int flush_mode, int flush_unroll, bool early_stage2, bool literal_tables, bool lids_are_conn, int pipe, int store_hint, int prefetch2, bool push, bool prefetch_meta, int cache_axes) {
  typedef Q1Shape<DIM> S;
  std::string o;
  o += "// generated by mrhyde_b200 (abi.cu: thermal_jit_source)\n";
  o += "#define MRH_JIT_DIM " + std::to_string(DIM) + "\n";
  o += "#define MRH_JIT_TRANSIENT 0  /*@transient@*/\n";
  o += "#define MRH_JIT_MODE 7  /*@mode@*/\n";
  // cell classes of this mesh; when coefficients are not all constant every cell takes the general path
  o += "#define MRH_JIT_HAS_GENERAL " + std::to_string((n_class[0] > 0 || !all_const) ? 1 : 0) + "\n";
  o += "#define MRH_JIT_HAS_AFFINE " + std::to_string((n_class[1] > 0 && all_const) ? 1 : 0) + "\n";
  o += "#define MRH_JIT_HAS_BOX " + std::to_string((n_class[2] > 0 && all_const) ? 1 : 0) + "\n";
  o += "#define MRH_JIT_ALL_CONST " + std::to_string(all_const) + "\n";
  o += "#define MRH_JIT_SOURCE_CONST " + std::to_string(source_const) + "\n";
  if (late_stage1) o += "#define MRH_JIT_LATE_STAGE1 1\n";
  if (push) o += "#define MRH_JIT_PUSH 1   /* multi-rank plan: ghost rows are also stored into the owner's receive slab (kernel_abi.h: PushDev) */\n";
  if (prefetch2) o += "#define MRH_JIT_PREFETCH2 " + std::to_string(prefetch2) + "   /* L2 prefetch of the next step's state / vertices before the pull (2: first vertex only) */\n";
  if (early_stage2) o += "#define MRH_JIT_EARLY_STAGE2 1\n";
  if (prefetch_meta) o += "#define MRH_JIT_PREFETCH_META 1   /* L2 prefetch of the record streams of step s + 2 */\n";
  if (literal_tables) o += "#define MRH_JIT_LITERAL_TABLES 1\n";
  o += "/*@stagger@*/\n";
  o += "#define MRH_JIT_FLUSH_UNROLL " + std::to_string(flush_unroll) + "\n";
  o += "#define MRH_JIT_FLUSH " + std::to_string(flush_mode) + "   /* 2: bulk-copy engine (cp.async.bulk shared -> global per run of CSR-contiguous rows), 1: one store instruction per row, 0: flat stream over the batch */\n";
  if (lids_are_conn) o += "#define MRH_JIT_LIDS_ARE_CONN 1   /* dof ids == vertex ids for every element: one connectivity stream */\n";
  {
    int present = 0, only = -1;
    for (int c = 0; c < 3; ++c) if (n_class[c] > 0) { ++present; only = c; }
    if (present == 1) o += "#define MRH_JIT_ONLY_ECLASS " + std::to_string(only) + "   /* every cell of the mesh has this class: no per-element class load */\n";
  }
  o += "#define MRH_DEBUG_SKIP " + std::to_string(debug_skip) + "   /* timing experiments only: 1 no global stores, 2 no element work, 3 neither */\n";
  if (!class_rep.empty()) o += "#define MRH_JIT_CLASS_NC " + std::to_string(class_rep.size()) + "\n";   // class ring (volume_kernel.cuh)
  if (!class_rep.empty() && pipe > 0) o += "#define MRH_JIT_PIPE " + std::to_string(pipe) + "   /* register-staged pipeline: 2 = even / odd warps run pull and element work in opposite order, 1 = same order */\n";
  if (metric_ng > 0) {   // metric ring (volume_kernel.cuh): parallelepiped cells + constant coefficients only
    o += "#define MRH_JIT_METRIC 1\n#define MRH_JIT_METRIC_NG " + std::to_string(metric_ng) + "\n";
  }
  o += "#define MRH_JIT_CAP " + std::to_string(cp.cap) + "   /* ring slot capacity: staging offsets become immediates */\n";
  o += "#define MRH_JIT_STORE_HINT " + std::to_string(store_hint) + "   /* 1: finished rows leave with an L2 evict-first policy (state / vertices stay L2-resident) */\n";
  o += kKernelAbiSrc;
  o += "\nnamespace mrhyde_b200 {\n";
  o += "__device__ __forceinline__ double mrh_abs(double a) { return a < 0.0 ? -a : a; }\n";
  o += "__device__ __forceinline__ double mrh_sqrt(double a) { return a <= 0.0 ? 0.0 : sqrt(a); }\n";
  o += "__device__ __forceinline__ double mrh_max(double a, double b) { return b > a ? b : a; }\n";
  o += "__device__ __forceinline__ double mrh_min(double a, double b) { return b < a ? b : a; }\n";
  // sin / cos without the math library's coefficient table in global memory: Cody-Waite reduction by pi/2 (three FMA
  // terms, exact for |x| < 1e5) and the fdlibm k_sin / k_cos minimax polynomials on [-pi/4, pi/4] (< 1 ulp); larger or
  // non-finite arguments take the library routine
  o += R"MRH(
// coefficients live in the constant bank so that each one is a direct operand of its FMA
__constant__ double mrh_sck[16] = {0x1.45f306dc9c883p-1, 0x1.921fb54442d18p+0, 0x1.1a62633145c07p-54, -0x1.f1976b7ed8fbcp-110,
  -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07, 2.48015872894767294178e-05,
  -1.38888888888741095749e-03, 4.16666666666666019037e-02,
  1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06, -1.98412698298579493134e-04,
  8.33333333332248946124e-03, -1.66666666666666324348e-01};
__device__ __forceinline__ double mrh_sincos(double x, int shift) {
  if (!(fabs(x) < 1.0e5)) return shift ? cos(x) : sin(x);
  const double q = rint(x * mrh_sck[0]);
  double r = fma(-q, mrh_sck[1], x);
  r = fma(-q, mrh_sck[2], r);
  r = fma(-q, mrh_sck[3], r);
  const int n = (int)q + shift;
  const double z = r * r;
  double v;
  if (n & 1) {
    double p = fma(z, mrh_sck[4], mrh_sck[5]);
    p = fma(z, p, mrh_sck[6]);
    p = fma(z, p, mrh_sck[7]);
    p = fma(z, p, mrh_sck[8]);
    p = fma(z, p, mrh_sck[9]);
    v = 1.0 - (0.5 * z - z * (z * p));
  } else {
    double p = fma(z, mrh_sck[10], mrh_sck[11]);
    p = fma(z, p, mrh_sck[12]);
    p = fma(z, p, mrh_sck[13]);
    p = fma(z, p, mrh_sck[14]);
    p = fma(z, p, mrh_sck[15]);
    v = fma(z * r, p, r);
  }
  return (n & 2) ? -v : v;
}
__device__ __forceinline__ double mrh_sin(double x) { return mrh_sincos(x, 0); }
__device__ __forceinline__ double mrh_cos(double x) { return mrh_sincos(x, 1); }
)MRH";
  const char* fn[4][2] = {{"mrh_fn_source", "thermal source"}, {"mrh_fn_diffusion", "thermal diffusion"}, {"mrh_fn_specific_heat", "specific heat"}, {"mrh_fn_density", "density"}};
  for (auto& f : fn)
    o += std::string("__device__ __forceinline__ double ") + f[0] + "(double x, double y, double z, double t) { return " + fs.codegen(f[1]) + "; }\n";
  // distinct cubature coordinates per axis: sub-expressions of one coordinate are evaluated once per distinct value
  int nqa[3] = {1, 1, 1}, qidx[S::NQ * 3];
  double qax[3][S::NQ];
  for (int d = 0; d < 3; ++d) for (int q = 0; q < S::NQ; ++q) { qax[d][q] = 0.0; qidx[q * 3 + d] = 0; }
  for (int d = 0; d < DIM; ++d) {
    nqa[d] = 0;
    for (int q = 0; q < S::NQ; ++q) {
      int f = -1;
      for (int i = 0; i < nqa[d]; ++i) if (qax[d][i] == T.qpt[q][d]) f = i;
      if (f < 0) { f = nqa[d]++; qax[d][f] = T.qpt[q][d]; }
      qidx[q * 3 + d] = f;
    }
  }
  int cache_n = 0, shared_n = 0;
  const int shared_axes = (cache_axes >> 4) & 7;   // packed by column_cache_axes: bits 0-2 per-thread cache, bits 4-6 CTA-shared values
  cache_axes &= 7;
  {
    std::string gen = fs.codegen_tensor("thermal source", "mrh_fn_source_box", S::NQ, nqa, qidx, cache_axes, &cache_n, shared_axes, &shared_n);
    if (shared_n > 8) gen = fs.codegen_tensor("thermal source", "mrh_fn_source_box", S::NQ, nqa, qidx, cache_axes, &cache_n, 0, &shared_n);   // 64 bytes are reserved
    o += gen;
  }
  o += "#define MRH_SRC_SHARED_N " + std::to_string(shared_n) + "   /* sub-expression values all elements of a step share (shared memory, axes mask " + std::to_string(shared_n > 0 ? shared_axes : 0) + ") */\n";
  o += "#define MRH_SRC_SHARED_AXES " + std::to_string(shared_n > 0 ? shared_axes : 0) + "\n";
  o += "#define MRH_SRC_CACHE_N " + std::to_string(cache_n) + "   /* one-coordinate sub-expression values of the source a thread keeps along its chain (axes mask " + std::to_string(cache_axes) + ") */\n";
  o += "#define MRH_NQA0 " + std::to_string(nqa[0]) + "\n#define MRH_NQA1 " + std::to_string(nqa[1]) + "\n#define MRH_NQA2 " + std::to_string(nqa[2]) + "\n";
  o += "namespace jit_tab {\n";
  auto arr = [&](const char* decl, const double* v, size_t n) { o += std::string("constexpr double ") + decl + " = {"; emit_values(o, v, n); o += "};\n"; };
  const std::string nq = std::to_string(S::NQ), nv = std::to_string(S::NV), dm = std::to_string(DIM), nt = std::to_string(S::NT), ng = std::to_string(S::NG);
  arr(("gN[" + nq + "][" + nv + "]").c_str(), &T.gN[0][0], S::NQ * S::NV);
  arr(("gdN[" + nq + "][" + nv + "][" + dm + "]").c_str(), &T.gdN[0][0][0], S::NQ * S::NV * DIM);
  arr(("phi[" + nq + "][" + nv + "]").c_str(), &T.phi[0][0], S::NQ * S::NV);
  arr(("dphi[" + nq + "][" + nv + "][" + dm + "]").c_str(), &T.dphi[0][0][0], S::NQ * S::NV * DIM);
  arr(("qw[" + nq + "]").c_str(), &T.qw[0], S::NQ);
  arr(("qpt[" + nq + "][" + dm + "]").c_str(), &T.qpt[0][0], S::NQ * DIM);
  arr(("Stab[" + ng + "][" + nt + "]").c_str(), &T.Stab[0][0], S::NG * S::NT);
  arr(("Mtab[" + nt + "]").c_str(), &T.Mtab[0], S::NT);
  arr(("Ltab[" + nv + "]").c_str(), &T.Ltab[0], S::NV);
  arr(("qax[3][" + nq + "]").c_str(), &qax[0][0], 3 * S::NQ);
  if (!class_rep.empty()) {
    auto iarr = [&](const std::string& decl, const std::vector<int32_t>& v) {
      o += "constexpr int " + decl + "[" + std::to_string(v.size()) + "] = {";
      for (size_t i = 0; i < v.size(); ++i) o += std::to_string(v[i]) + (i + 1 < v.size() ? "," : "");
      o += "};\n";
    };
    iarr("cls", class_of_t); iarr("rep", class_rep);
  }
  o += "}  // namespace jit_tab\nnamespace jit_ctab {\n";
  auto carr = [&](const char* decl, const double* v, size_t n) { o += std::string("__constant__ double ") + decl + " = {"; emit_values(o, v, n); o += "};\n"; };
  carr(("phi[" + nq + "][" + nv + "]").c_str(), &T.phi[0][0], S::NQ * S::NV);
  carr(("qw[" + nq + "]").c_str(), &T.qw[0], S::NQ);
  carr(("Stab[" + ng + "][" + nt + "]").c_str(), &T.Stab[0][0], S::NG * S::NT);
  carr(("Mtab[" + nt + "]").c_str(), &T.Mtab[0], S::NT);
  carr(("Ltab[" + nv + "]").c_str(), &T.Ltab[0], S::NV);
  o += "}  // namespace jit_ctab\n";
  // gather patterns of the plan as constant data (warp-uniform reads in the pull phase hit the constant cache)
  const size_t desc_words = cp.desc[0].size();
  if (desc_words > 0 && desc_words * 4 * 2 <= 40 * 1024) {
    o += metric_ng > 0 ? "#define MRH_JIT_CONST_MDESC 1\n" : "#define MRH_JIT_CONST_DESC 1\n";
    for (int par = 0; par < 2; ++par) {
      const std::vector<uint32_t>& D = metric_ng > 0 ? cp.mdesc[par] : cp.desc[par];
      o += std::string("__constant__ uint4 ") + (metric_ng > 0 ? "mrh_mdesc" : "mrh_desc") + std::to_string(par) + "[" + std::to_string(desc_words / 4) + "] = {";
      char b[64];
      for (size_t q = 0; q < desc_words / 4; ++q) {
        std::snprintf(b, sizeof(b), "{%uu,%uu,%uu,%uu}%s", D[4 * q], D[4 * q + 1], D[4 * q + 2], D[4 * q + 3], q + 1 < desc_words / 4 ? "," : "");
        o += b;
      }
      o += "};\n";
    }
  }
  if (metric_ng > 0) o += pull_codegen_metric(cp, max_patterns, S::NV, S::NG, metric_ng, &T.Stab[0][0], &T.Mtab[0], pull_group, flush_mode);
  else o += pull_codegen(cp, max_patterns, pull_group, flush_mode);
  o += "}  // namespace mrhyde_b200\n";
  o += kVolumeKernelSrc;
  return o;
}

// ---- patterns that get generated pull code: the most frequent ones, by rows ------------------------------------------
struct SpecialPattern { int32_t desc_begin; int n_slots; };
std::vector<SpecialPattern> special_patterns(const ChainPlan& cp, int max_patterns) {
  std::map<int32_t, int64_t> freq;   // desc_begin -> rows
  for (const BatchRec& B : cp.batches) if (!(B.flags & BATCH_FIXED)) freq[B.desc_begin] += B.n_rows;
  std::vector<std::pair<int64_t, int32_t>> order;
  for (auto& kv : freq) order.push_back({kv.second, kv.first});
  std::sort(order.rbegin(), order.rend());
  std::vector<SpecialPattern> out;
  if (max_patterns <= 0) return out;
  auto slots_of = [&](int32_t db) { int n = 0; for (const PatternRec& PR : cp.patterns) if (PR.desc_begin == db) n = PR.n_slots; return n; };
  for (auto& pr : order) {
    if ((int)out.size() >= max_patterns) break;
    const int n_slots = slots_of(pr.second);
    if (n_slots < 2 || n_slots > 65) continue;   // rows of up to 64 entries go through the per-warp row buffer
    out.push_back({pr.second, n_slots});
  }
  for (int32_t db : cp.ghost_patterns) {   // patterns of ghost rows (in-kernel halo push), whatever their frequency
    bool have = false;
    for (const SpecialPattern& sp : out) if (sp.desc_begin == db) have = true;
    const int n_slots = slots_of(db);
    if (!have && n_slots >= 2 && n_slots <= 65) out.push_back({db, n_slots});
  }
  return out;
}
int row_pitch(int n_jac) { return n_jac | 1; }   // odd pitch: a lane per row writes the row buffer without bank conflicts
// doubles per row of the per-warp row buffer (MRH_JIT_ROWBUF); 0 when no pattern is specialised
// flush_mode 2 (bulk copies): rows sit back to back (pitch = row length) from the start of the warp's buffer (the 32 row offsets of
// the other modes are not stored), and every run of CSR-contiguous rows may be shifted by up to 2 * 31 + 1 doubles so that its
// shared-memory address has the 16-byte phase of its global address: 32 NJ + 63 <= 32 + 32 (NJ + 1) doubles
int row_buffer_pitch(const ChainPlan& cp, int max_patterns, int flush_mode) {
  int pitch = 0;
  for (const SpecialPattern& sp : special_patterns(cp, max_patterns))
    pitch = std::max(pitch, flush_mode == 2 ? sp.n_slots - 1 + 1 : row_pitch(sp.n_slots - 1));
  return pitch;
}
// Row flush shared by the generated pull code: the warp's rows sit in rowb (row r at r * PITCH), their CSR offsets in wbase.
// One store instruction writes one row (lane = entry), so a row leaves as one contiguous piece of NJ doubles whatever the
// neighbouring rows are, and every shared-memory offset is an immediate.
const char* kFlushRowsSrc = R"MRH(
#if MRH_JIT_FLUSH == 1
template <int NJ, int PITCH, bool ACC>
__device__ __forceinline__ void mrh_flush_rows(const double* rowb, const long long* wbase, double* __restrict__ jac, const int n_rows, const int lane) {
  __syncwarp();
  if (!(MRH_DEBUG_SKIP & 1)) {
    if (n_rows == 32) {   // full batch: no per-row tests
#pragma unroll MRH_JIT_FLUSH_UNROLL
      for (int r = 0; r < 32; ++r) {
#pragma unroll
        for (int c = 0; c < NJ; c += 32) {
          const int k = c + lane;
          if (k < NJ) { double v = rowb[r * PITCH + k]; double* p = jac + wbase[r] + k; if (ACC) v += *p; *p = v; }
        }
      }
    } else {
      for (int r = 0; r < n_rows; ++r) {
        for (int k = lane; k < NJ; k += 32) { double v = rowb[r * PITCH + k]; double* p = jac + wbase[r] + k; if (ACC) v += *p; *p = v; }
      }
    }
  }
  __syncwarp();
}
#else
// flat variant: position p = lane + 32 i of the concatenated rows is written by lane `lane` in round i, so rows that are
// neighbours in the CSR array -- a batch holds rows in ascending order -- leave as 256-byte contiguous stores
template <int NJ, int PITCH, bool ACC>
__device__ __forceinline__ void mrh_flush_rows(const double* rowb, const long long* wbase, double* __restrict__ jac, const int n_rows, const int lane) {
  __syncwarp();
  if (!(MRH_DEBUG_SKIP & 1)) {
    int row = lane / NJ, k = lane % NJ;
#pragma unroll 3
    for (int i = 0; i < NJ; ++i) {
      if (row < n_rows) {
        double v = rowb[row * PITCH + k];
        double* p = jac + wbase[row] + k;
        if (ACC) v += *p;
        *p = v;
      }
      row += 32 / NJ; k += 32 % NJ;
      if (k >= NJ) { k -= NJ; ++row; }
    }
  }
  __syncwarp();
}
#endif
#if MRH_JIT_FLUSH == 2
// Bulk-copy flush.  Rows of a batch that follow each other in the CSR value array form a RUN; the lanes of a run park their rows back
// to back (pitch NJ), the run shifted by 0..1 doubles so that shared and global address agree modulo 16 bytes, and the first lane of
// the run hands the whole run to the bulk-copy engine (cp.async.bulk shared::cta -> global; an odd first / last double is peeled off
// with a plain store).  The warp no longer reads the buffer back or issues one store per row: the LSU pipe only sees the parking
// stores.  Accumulate mode uses the reducing form (cp.reduce.async.bulk .add.f64): every entry receives exactly one addend per call.
struct MrhRun { double* rl; unsigned starts; bool start; };
__device__ __forceinline__ void mrh_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
template <int NJ>
__device__ __forceinline__ MrhRun mrh_run_begin(double* rowb, const double* jac, const long long base, const int n_rows, const int lane, const PushDev* X) {
  mrh_bulk_wait_read();   // the previous batch of this warp has left the buffer (the issuing lanes wait; a no-op for the others)
  __syncwarp();
  const long long prev = __shfl_up_sync(0xffffffffu, base, 1);
  MrhRun R;
  R.start = lane < n_rows && (lane == 0 || base != prev + NJ);
#ifdef MRH_JIT_PUSH
  if (X->enabled) R.start = R.start || (lane < n_rows && ((base >= X->ghost_base) != (prev >= X->ghost_base)));   // a run is owned rows or ghost rows, never both
#endif
  R.starts = __ballot_sync(0xffffffffu, R.start);
  const int ridx = __popc(R.starts & (0xffffffffu >> (31 - lane))) - 1;   // runs that begin at or below this lane, minus one
  const int e = (int)(((long long)(reinterpret_cast<unsigned long long>(jac + base) >> 3) - (long long)(lane * NJ)) & 1);   // constant within a run
  // idle lanes (they park garbage sums) sit one run further up so that they cannot touch the last row of the last run
  R.rl = rowb + lane * NJ + 2 * (ridx + (lane >= n_rows ? 1 : 0)) + e;
  return R;
}
template <int NJ, bool ACC>
__device__ __forceinline__ void mrh_run_flush(const MrhRun& R, const long long base, double* __restrict__ jac, const int n_rows, const int lane, const PushDev* X) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores of this lane -> visible to the bulk-copy engine
  __syncwarp();
  if (R.start && !(MRH_DEBUG_SKIP & 1)) {
    const unsigned later = R.starts & ~(0xffffffffu >> (31 - lane));
    const int end = later ? (__ffs((int)later) - 1) : n_rows;
    int cnt = (end - lane) * NJ;
    long long g = base;
    const double* s = R.rl;
    if ((reinterpret_cast<unsigned long long>(jac + g) >> 3) & 1) { double v = s[0]; if (ACC) v += jac[g]; jac[g] = v; ++g; ++s; --cnt; }
    if (cnt & 1) { double v = s[cnt - 1]; if (ACC) v += jac[g + cnt - 1]; jac[g + cnt - 1] = v; --cnt; }
    if (cnt > 0) {
      const unsigned sa = (unsigned)__cvta_generic_to_shared(s);
#if MRH_JIT_STORE_HINT
      unsigned long long pol;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
      if (ACC) asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.L2::cache_hint.add.f64 [%0], [%1], %2, %3;" :: "l"(jac + g), "r"(sa), "r"(cnt * 8), "l"(pol) : "memory");
      else asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" :: "l"(jac + g), "r"(sa), "r"(cnt * 8), "l"(pol) : "memory");
#else
      if (ACC) asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" :: "l"(jac + g), "r"(sa), "r"(cnt * 8) : "memory");
      else asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(jac + g), "r"(sa), "r"(cnt * 8) : "memory");
#endif
    }
#ifdef MRH_JIT_PUSH
    if (X->enabled && base >= X->ghost_base) {
      // ghost rows: the same run once more, into the owner's receive slab over NVLink (plain copies: the slab holds this call's
      // values; its 16-byte phase equals the local array's, PushDev::shift)
      int c2 = (end - lane) * NJ;
      long long g2 = base - X->ghost_base;
      const double* s2 = R.rl;
      double* rj = X->remote_jac;
      if ((reinterpret_cast<unsigned long long>(rj + g2) >> 3) & 1) { rj[g2] = s2[0]; ++g2; ++s2; --c2; }
      if (c2 & 1) { rj[g2 + c2 - 1] = s2[c2 - 1]; --c2; }
      if (c2 > 0) {
        const unsigned sa2 = (unsigned)__cvta_generic_to_shared(s2);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(rj + g2), "r"(sa2), "r"(c2 * 8) : "memory");
      }
    }
#endif
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}
#ifdef MRH_JIT_PUSH
// residual entry of a ghost row: also into the owner's slab
__device__ __forceinline__ void mrh_push_res(const PushDev* X, const double* pres, const double v) {
  if (X->enabled) {
    const long long row = pres - X->res_base;
    if (row >= X->n_owned) X->remote_res[row - X->n_owned] = v;
  }
}
#define MRH_PUSH_RES(pres, v) mrh_push_res(X, (pres), (v))
#else
#define MRH_PUSH_RES(pres, v)
#endif
#endif
)MRH";

// ---- straight-line pull code for the plan's most frequent gather patterns (jit only) ------------------------------
// For a pattern the slot descriptors are plan constants, so the generated code has no descriptor loads, no "unused"
// tests and no address arithmetic: every staged value is one LDS with an immediate offset from the row's ring anchor.
// Sums keep the ascending element order of the generic loop (slot_sum), so both paths give identical bits.
std::string pull_codegen(const ChainPlan& cp, int max_patterns, int group, int flush_mode) {
  const bool bulk = flush_mode == 2;
  group = std::max(4, (group / 4) * 4);   // sums formed before they are parked in the row buffer: loads in flight vs registers
  const std::vector<SpecialPattern> sel = special_patterns(cp, max_patterns);
  std::string o;
  if (sel.empty()) return o;
  o += "#define MRH_JIT_PULL 1\n#define MRH_JIT_ROWBUF " + std::to_string(row_buffer_pitch(cp, max_patterns, flush_mode)) + "\n";
  o += "__device__ __forceinline__ double mrh_lds_at(unsigned a) { double v; asm volatile(\"ld.shared.f64 %0, [%1];\" : \"=d\"(v) : \"r\"(a)); return v; }\n";
  o += "#define MRH_L(off) mrh_lds_at(rbase + (off##u))\n";
  o += kFlushRowsSrc;
  o += "template <bool HAS_RES, bool HAS_JAC, bool ACC>\n__device__ __forceinline__ bool mrh_pull_special(const int desc_begin, const int parity, const unsigned rbase, "
       "double* __restrict__ wbuf, const int lane, const int n_rows, const long long base, double* __restrict__ jac, double* pres, const bool active, const PushDev* X) {\n";
  o += "  long long* const wbase = reinterpret_cast<long long*>(wbuf);\n  double* const rowb = wbuf + 32;\n";
  o += "  switch (desc_begin * 2 + parity) {\n";
  for (const SpecialPattern& sp : sel) {
    const int32_t db = sp.desc_begin;
    const int n_jac = sp.n_slots - 1, pitch = row_pitch(n_jac);
    for (int par = 0; par < 2; ++par) {
      auto sum_expr = [&](int k) {
        std::string e;
        int cnt = 0;
        for (int z = 0; z < SLOT_SRCS; ++z) {
          const uint32_t src = cp.desc[par][((size_t)db + (size_t)k) * SLOT_SRCS + (size_t)z];
          if (src == SRC_NONE) continue;
          const std::string t = "MRH_L(" + std::to_string(src) + ")";
          e = cnt == 0 ? t : "(" + e + " + " + t + ")";
          ++cnt;
        }
        return cnt ? e : std::string("0.0");
      };
      o += "    case " + std::to_string(db * 2 + par) + ": {\n      if (HAS_JAC) {\n";
      if (bulk) o += "        const MrhRun run = mrh_run_begin<" + std::to_string(n_jac) + ">(wbuf, jac, base, n_rows, lane, X);\n";
      else o += "        wbase[lane] = base;\n";
      // a group of sums first (the loads are independent and can be in flight together), then they are parked in the row buffer
      for (int g0 = 0; g0 < n_jac; g0 += group) {
        for (int k = g0; k < g0 + group && k < n_jac; ++k) o += "        const double a" + std::to_string(k) + " = " + sum_expr(k) + ";\n";
        for (int k = g0; k < g0 + group && k < n_jac; ++k)
          o += bulk ? "        run.rl[" + std::to_string(k) + "] = a" + std::to_string(k) + ";\n"
                    : "        rowb[lane * " + std::to_string(pitch) + " + " + std::to_string(k) + "] = a" + std::to_string(k) + ";\n";
      }
      if (bulk) o += "        mrh_run_flush<" + std::to_string(n_jac) + ", ACC>(run, base, jac, n_rows, lane, X);\n";
      else o += "        mrh_flush_rows<" + std::to_string(n_jac) + ", " + std::to_string(pitch) + ", ACC>(rowb, wbase, jac, n_rows, lane);\n";
      o += "      }\n      if (HAS_RES) { const double acc = " + sum_expr(n_jac) + "; if (active) { double v = -acc; if (ACC) v += *pres; *pres = v; " + (bulk ? "MRH_PUSH_RES(pres, v); " : "") + "} }\n";
      o += "      return true;\n    }\n";
    }
  }
  o += "    default: return false;\n  }\n}\n";
  return o;
}

// ---- the same for the metric ring: a CSR entry is  sum_e sum_g G_g(e) Stab[g][t_e]  with the element columns, table entries
// and state slots of the pattern as immediates.  The metric entries of the pattern's element columns are loaded once per row.
std::string pull_codegen_metric(const ChainPlan& cp, int max_patterns, int nv, int ng_all, int ngu, const double* Stab, const double* Mtab, int group, int flush_mode) {
  const bool bulk = flush_mode == 2;
  group = std::max(4, (group / 4) * 4);   // entries summed before their transpose rounds start (instruction-level parallelism vs registers)
  const std::vector<SpecialPattern> sel = special_patterns(cp, max_patterns);
  const int nt = nv * (nv + 1) / 2;
  (void)ng_all;
  std::string o;
  o += "#define MRH_JIT_PULL_METRIC 1\n#define MRH_JIT_ROWBUF " + std::to_string(std::max(1, row_buffer_pitch(cp, max_patterns, flush_mode))) + "\n";
  o += kFlushRowsSrc;
  o += "#define MRH_ESB (16u * MRH_JIT_CAP)\n";
  o += "#define MRH_MD MRH_JIT_METRIC_NG\n#define MRH_B0 (MRH_JIT_METRIC_NG + MRH_JIT_TRANSIENT)\n";
  o += "#define MRH_U0 (MRH_B0 + " + std::to_string(nv) + ")\n#define MRH_UT0 (MRH_U0 + " + std::to_string(nv) + ")\n";
  o += "__device__ __forceinline__ double mrh_mlds(unsigned a) { double v; asm volatile(\"ld.shared.f64 %0, [%1];\" : \"=d\"(v) : \"r\"(a)); return v; }\n";
  o += "#define MRH_ML(off, m) mrh_mlds(rbase + (off) + (m) * MRH_ESB)\n";
  o += "template <bool HAS_RES, bool HAS_JAC, bool ACC>\n__device__ __forceinline__ bool mrh_pull_metric_special(const int desc_begin, const int parity, const unsigned rbase, "
       "double* __restrict__ wbuf, const int lane, const int n_rows, const long long base, double* __restrict__ jac, double* pres, const bool active, "
       "const double au, const double at, const PushDev* X) {\n";
  o += "  long long* const wbase = reinterpret_cast<long long*>(wbuf);\n  double* const rowb = wbuf + 32;\n";
  o += "  switch (desc_begin * 2 + parity) {\n";
  for (const SpecialPattern& sp : sel) {
    const int32_t db = sp.desc_begin;
    const int n_slots = sp.n_slots, n_jac = n_slots - 1, pitch = row_pitch(n_jac);
    for (int par = 0; par < 2; ++par) {
      auto word = [&](int k, int z) { return cp.mdesc[par][((size_t)db + (size_t)k) * SLOT_SRCS + (size_t)z]; };
      // element columns of the pattern in order of first use
      std::vector<uint32_t> cols;
      auto col_of = [&](uint32_t off) {
        for (size_t c = 0; c < cols.size(); ++c) if (cols[c] == off) return (int)c;
        cols.push_back(off);
        return (int)cols.size() - 1;
      };
      for (int k = 0; k < n_slots; ++k) for (int z = 0; z < SLOT_SRCS; ++z) if (word(k, z) != SRC_NONE) col_of(word(k, z) & MSRC_OFF_MASK);
      o += "    case " + std::to_string(db * 2 + par) + ": {\n";
      for (size_t c = 0; c < cols.size(); ++c) {
        for (int g = 0; g < ngu; ++g)
          o += "      const double g" + std::to_string(c) + "_" + std::to_string(g) + " = MRH_ML(" + std::to_string(cols[c]) + "u, " + std::to_string(g) + ");\n";
        o += "#if MRH_JIT_TRANSIENT\n      const double m" + std::to_string(c) + " = MRH_ML(" + std::to_string(cols[c]) + "u, MRH_MD);\n#endif\n";
      }
      if (bulk) o += "      MrhRun run; run.rl = rowb; run.starts = 0u; run.start = false;\n      if (HAS_JAC) run = mrh_run_begin<" + std::to_string(n_jac) + ">(wbuf, jac, base, n_rows, lane, X);\n      double racc = 0.0;\n";
      else o += "      if (HAS_JAC) wbase[lane] = base;\n      double racc = 0.0;\n";
      for (int g0 = 0; g0 < n_jac; g0 += group) {
        for (int k = g0; k < g0 + group && k < n_jac; ++k) {
          const std::string a = "a" + std::to_string(k), mm = "mm" + std::to_string(k);
          std::string ea, em;
          bool first_a = true, first_m = true;
          uint32_t w0 = SRC_NONE;
          for (int z = 0; z < SLOT_SRCS; ++z) {
            const uint32_t w = word(k, z);
            if (w == SRC_NONE) continue;
            if (w0 == SRC_NONE) w0 = w;
            const int c = col_of(w & MSRC_OFF_MASK), t = (int)((w >> MSRC_T_SHIFT) & 63u);
            for (int g = 0; g < ngu; ++g) {
              if (Stab[(size_t)g * nt + t] == 0.0) continue;   // exact zeros of the reference tables add nothing
              const std::string G = "g" + std::to_string(c) + "_" + std::to_string(g), Sx = hexd(Stab[(size_t)g * nt + t]);   // literal: equal coefficients share a register
              if (first_a) { ea += "      double " + a + " = " + G + " * " + Sx + ";\n"; first_a = false; }
              else ea += "      " + a + " = fma(" + G + ", " + Sx + ", " + a + ");\n";
            }
            const std::string Mx = hexd(Mtab[t]);
            if (first_m) { em += "      double " + mm + " = m" + std::to_string(c) + " * " + Mx + ";\n"; first_m = false; }
            else em += "      " + mm + " = fma(m" + std::to_string(c) + ", " + Mx + ", " + mm + ");\n";
          }
          if (first_a) ea += "      double " + a + " = 0.0;\n";
          if (first_m) em += "      double " + mm + " = 0.0;\n";
          o += ea;
          o += "#if MRH_JIT_TRANSIENT\n" + em + "#endif\n";
          if (w0 != SRC_NONE) {
            const std::string off = std::to_string(w0 & MSRC_OFF_MASK) + "u", j0 = std::to_string((w0 >> MSRC_J_SHIFT) & 7u);
            o += "      if (HAS_RES) {\n        racc = fma(" + a + ", MRH_ML(" + off + ", MRH_U0 + " + j0 + "), racc);\n";
            o += "#if MRH_JIT_TRANSIENT\n        racc = fma(" + mm + ", MRH_ML(" + off + ", MRH_UT0 + " + j0 + "), racc);\n#endif\n      }\n";
          }
          o += "#if MRH_JIT_TRANSIENT\n      " + a + " = fma(au, " + a + ", at * " + mm + ");\n#endif\n";
        }
        for (int k = g0; k < g0 + group && k < n_jac; ++k)
          o += bulk ? "      if (HAS_JAC) run.rl[" + std::to_string(k) + "] = a" + std::to_string(k) + ";\n"
                    : "      if (HAS_JAC) rowb[lane * " + std::to_string(pitch) + " + " + std::to_string(k) + "] = a" + std::to_string(k) + ";\n";
      }
      if (bulk) o += "      if (HAS_JAC) mrh_run_flush<" + std::to_string(n_jac) + ", ACC>(run, base, jac, n_rows, lane, X);\n";
      else o += "      if (HAS_JAC) mrh_flush_rows<" + std::to_string(n_jac) + ", " + std::to_string(pitch) + ", ACC>(rowb, wbase, jac, n_rows, lane);\n";
      o += "      if (HAS_RES) {\n        double bs = 0.0;\n";
      for (int z = 0; z < SLOT_SRCS; ++z) {
        const uint32_t w = word(n_jac, z);
        if (w == SRC_NONE) continue;
        o += "        bs += MRH_ML(" + std::to_string(w & MSRC_OFF_MASK) + "u, MRH_B0 + " + std::to_string((w >> MSRC_I_SHIFT) & 7u) + ");\n";
      }
      o += std::string("        if (active) { double v = bs - racc; if (ACC) v += *pres; *pres = v; ") + (bulk ? "MRH_PUSH_RES(pres, v); " : "") + "}\n      }\n";
      o += "      return true;\n    }\n";
    }
  }
  o += "    default: return false;\n  }\n}\n";
  return o;
}


// ---- module tables of the general path ----------------------------------------------------------------------------
// YAML key and default of every coefficient function a module registers in defineFunctions, in the index order the
// device physics reads them (general_physics.cuh): thermal.cpp:47-65, linearelasticity.cpp:64-86,
// navierstokes.cpp:64-76, maxwell.cpp:63-78 (there the YAML keys permeability / permittivity / conductivity feed the
// functions mu / epsilon / sigma).
struct ModuleFn { const char* key; const char* def; };
static std::string canonical_module(std::string name) {
  while (!name.empty() && name.front() == ' ') name.erase(name.begin());
  while (!name.empty() && name.back() == ' ') name.pop_back();
  if (name == "linear elasticity") return "linearelasticity";
  if (name == "Navier Stokes" || name == "navierstokes") return "navier stokes";
  return name;
}
// "modules: a, b" (the reference's comma separated list, physicsInterface) -> "a+b": the name of a two-module block kernel
std::string canonical_physics(const std::string& name) {
  std::string out;
  size_t b = 0;
  while (b <= name.size()) {
    size_t e = name.find(',', b);
    if (e == std::string::npos) e = name.size();
    const std::string m = canonical_module(name.substr(b, e - b));
    if (!m.empty()) out += (out.empty() ? "" : "+") + m;
    b = e + 1;
  }
  return out;
}
const ModuleFn* module_functions(const std::string& phys) {
  static const ModuleFn thermal[] = {{"thermal source", "0.0"}, {"thermal diffusion", "1.0"}, {"specific heat", "1.0"}, {"density", "1.0"},
                                     {"advection x", "0.0"}, {"advection y", "0.0"}, {"advection z", "0.0"}, {"robin alpha", "0.0"}, {nullptr, nullptr}};
  static const ModuleFn le[] = {{"lambda", "1.0"}, {"mu", "0.5"}, {"source dx", "0.0"}, {"source dy", "0.0"}, {"source dz", "0.0"}, {nullptr, nullptr}};
  static const ModuleFn ns[] = {{"source ux", "0.0"}, {"source pr", "0.0"}, {"source uy", "0.0"}, {"source uz", "0.0"}, {"density", "1.0"}, {"viscosity", "1.0"}, {nullptr, nullptr}};
  static const ModuleFn mx[] = {{"current x", "0.0"}, {"current y", "0.0"}, {"current z", "0.0"}, {"permeability", "1.0"}, {"refractive index", "1.0"},
                                {"permittivity", "1.0"}, {"conductivity", "0.0"}, {nullptr, nullptr}};
  // two-module blocks: the functions of the first module, then those of the second (general_physics.cuh); "density" of thermal and of
  // navier stokes is one and the same `Functions:` entry, as in the reference's per-block function manager
  static const ModuleFn th_le[] = {{"thermal source", "0.0"}, {"thermal diffusion", "1.0"}, {"specific heat", "1.0"}, {"density", "1.0"},
                                   {"advection x", "0.0"}, {"advection y", "0.0"}, {"advection z", "0.0"}, {"robin alpha", "0.0"},
                                   {"lambda", "1.0"}, {"mu", "0.5"}, {"source dx", "0.0"}, {"source dy", "0.0"}, {"source dz", "0.0"}, {nullptr, nullptr}};
  static const ModuleFn ns_th[] = {{"source ux", "0.0"}, {"source pr", "0.0"}, {"source uy", "0.0"}, {"source uz", "0.0"}, {"density", "1.0"}, {"viscosity", "1.0"},
                                   {"thermal source", "0.0"}, {"thermal diffusion", "1.0"}, {"specific heat", "1.0"}, {"density", "1.0"},
                                   {"advection x", "0.0"}, {"advection y", "0.0"}, {"advection z", "0.0"}, {"robin alpha", "0.0"}, {nullptr, nullptr}};
  if (phys == "thermal") return thermal;
  if (phys == "linearelasticity") return le;
  if (phys == "navier stokes") return ns;
  if (phys == "maxwell") return mx;
  if (phys == "thermal+linearelasticity") return th_le;
  if (phys == "navier stokes+thermal") return ns_th;
  return nullptr;
}
std::vector<std::string> module_variables(const std::string& phys, int dim) {
  if (phys == "thermal") return {"T"};
  if (phys == "linearelasticity") return dim == 2 ? std::vector<std::string>{"dx", "dy"} : std::vector<std::string>{"dx", "dy", "dz"};
  if (phys == "navier stokes") return dim == 2 ? std::vector<std::string>{"ux", "pr", "uy"} : std::vector<std::string>{"ux", "pr", "uy", "uz"};
  if (phys == "maxwell") return {"E", "B"};
  const size_t plus = phys.find('+');
  if (plus != std::string::npos) {   // two-module block: variables module by module
    std::vector<std::string> a = module_variables(phys.substr(0, plus), dim), b = module_variables(phys.substr(plus + 1), dim);
    a.insert(a.end(), b.begin(), b.end());
    return a;
  }
  return {};
}

std::vector<std::string> initial_function_names(const mrhyde_b200_plan* P, int v) {
  const std::string& t = P->bases[(size_t)P->var_basis[(size_t)v]].type;
  const std::string base = "initial " + P->var_names[(size_t)v];
  if (t == "HCURL" || t == "HDIV") {
    std::vector<std::string> out;
    static const char* comp[3] = {"[x]", "[y]", "[z]"};
    for (int d = 0; d < P->dim; ++d) out.push_back(base + comp[d]);
    return out;
  }
  return {base};
}

FunctionSet make_function_set(const mrhyde_b200_plan* P, bool side, bool state_slots = false) {
  FunctionSet fs;
  // module defaults (thermal::defineFunctions, thermal.cpp:47-65), then user overrides
  for (const ModuleFn* f = module_functions(canonical_physics(P->physics)); f && f->key; ++f) fs.set(f->key, f->def);
  // "initial <var>" (HGRAD / HVOL) or "initial <var>[x|y|z]" (HCURL / HDIV): variables the Initial conditions sublist does not
  // name start from 0.0 (physicsInterface_functions.hpp:154-226)
  for (size_t v = 0; v < P->var_names.size(); ++v)
    for (const std::string& nm : initial_function_names(P, (int)v)) fs.set(nm, "0.0");
  for (auto& kv : P->functions) fs.set(kv.first, kv.second);
  std::vector<std::string> sol;
  static const char* comps[3] = {"[x]", "[y]", "[z]"};
  for (auto& v : P->var_names) {
    sol.push_back(v); sol.push_back(v + "_t");
    for (int d = 0; d < 3; ++d) {
      sol.push_back("grad(" + v + ")" + comps[d]);
      sol.push_back(v + comps[d]); sol.push_back(v + "_t" + comps[d]);
      sol.push_back("curl(" + v + ")" + comps[d]);
    }
    sol.push_back("div(" + v + ")");
  }
  fs.set_solution_fields(sol);
  if (state_slots) {
    // general path: the evaluator reads solution fields from slots v * NC + k (fields F[v][k] of general_physics.cuh), time
    // derivatives of the values NVAR * NC further on.  HGRAD: value, grad x y z; HCURL: value x y z, curl x y z; HDIV: value x y z, div
    std::map<std::string, int> slots;
    const int NC = canonical_physics(P->physics) == "maxwell" ? 6 : 4, NCV = NC * (int)P->var_names.size();
    for (size_t v = 0; v < P->var_names.size(); ++v) {
      const std::string& nm = P->var_names[v];
      const std::string& bt = P->bases[(size_t)P->var_basis[v]].type;
      const int base = (int)v * NC;
      if (bt == "HGRAD") {
        slots[nm] = base; slots[nm + "_t"] = NCV + base;
        for (int d = 0; d < 3; ++d) slots["grad(" + nm + ")" + comps[d]] = base + 1 + d;
      } else if (bt == "HCURL" || bt == "HDIV") {
        for (int d = 0; d < 3; ++d) { slots[nm + comps[d]] = base + d; slots[nm + "_t" + comps[d]] = NCV + base + d; }
        if (bt == "HCURL") for (int d = 0; d < 3; ++d) slots["curl(" + nm + ")" + comps[d]] = base + 3 + d;
        else slots["div(" + nm + ")"] = base + 3;
      }
    }
    fs.set_solution_slots(slots);
  }
  if (side) fs.set_scalar_fields({"x", "y", "z", "", "n[x]", "n[y]", "n[z]"});
  else fs.set_scalar_fields({"x", "y", "z"});
  // boundary data functions at side ip: "Dirichlet <var> <side>" / "Neumann <var> <side>"
  if (side)
    for (auto& kv : P->bcs) {
      const BCEntry& b = kv.second;
      if (b.type == "weak Dirichlet" || b.type == "Dirichlet") fs.set("Dirichlet " + kv.first.first + " " + kv.first.second, b.expr);
      else if (b.type == "Neumann") fs.set("Neumann " + kv.first.first + " " + kv.first.second, b.expr);
    }
  return fs;
}

__global__ void eval_points_kernel(const __grid_constant__ ExprProgram prog, const double* __restrict__ xyz, double time, int64_t n, double* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  ExprVars in;
  in.v[0] = xyz[3 * i]; in.v[1] = xyz[3 * i + 1]; in.v[2] = xyz[3 * i + 2]; in.v[3] = time; in.v[4] = in.v[5] = in.v[6] = 0.0;
  out[i] = expr_eval(prog, in);
}

__global__ void orphan_rows_kernel(const int32_t* __restrict__ rows, int n, GraphDev G, OutDev O) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t r = rows[i];
  if (O.res) O.res[r] = 0.0;
  if (O.jac)
    for (int64_t p = G.rowptr[r]; p < G.rowptr[r + 1]; ++p) O.jac[p] = (G.fixed[r] && G.colind[p] == r && O.diag_one) ? 1.0 : 0.0;
}

// dofConstraints -> setJacobianConstraints: J(d,d) = 1 on strong-Dirichlet dofs (replaceLocalValues)
__global__ void fixed_diag_kernel(const int64_t* __restrict__ diag, int n, double* __restrict__ jac) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && diag[i] >= 0) jac[diag[i]] = 1.0;
}

// dofConstraints, point constraints: the whole row becomes the identity row (one warp per row)
__global__ void point_rows_kernel(const int64_t* __restrict__ rows, int n, double* __restrict__ jac) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n) return;
  const int64_t b = rows[3 * w], e = rows[3 * w + 1], d = rows[3 * w + 2];
  for (int64_t p = b + lane; p < e; p += 32) jac[p] = (p == d) ? 1.0 : 0.0;
}

static int point_rows(mrhyde_b200_plan* P, double* jac, cudaStream_t st);
// Solver: fix zero rows (assemblyManager_jacres.hpp:609-626): rows with sum |J(row, :)| < 1e-14 get J(row,row) = 1; one warp per row
__global__ void fix_zero_rows_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colind, int64_t nrows, double* __restrict__ jac) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nrows) return;
  const int64_t b = rowptr[row], e = rowptr[row + 1];
  double s = 0.0;
  for (int64_t p = b + lane; p < e; p += 32) s += fabs(jac[p]);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (s < 1.0e-14)
    for (int64_t p = b + lane; p < e; p += 32) if (colind[p] == row) jac[p] = 1.0;
}

static int point_constraints(mrhyde_b200_plan* P, double* jac, cudaStream_t st) {
  int launched = point_rows(P, jac, st);   // second loop of dofConstraints
  if (jac && opt_bool(P, "fix zero rows", false) && P->mesh.nrows > 0) {
    // after dofConstraints, as in the reference; the row sums of an entry order other than the reference's differ in the last bits, which
    // only matters for rows within rounding of the 1e-14 threshold
    const int64_t n = P->mesh.nrows;
    fix_zero_rows_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(P->d_rowptr.p, P->d_colind.p, n, jac);
    CUDA_OK(cudaGetLastError());
    ++launched;
  }
  return launched;
}

static int point_rows(mrhyde_b200_plan* P, double* jac, cudaStream_t st) {
  const int n = (int)(P->d_point_rows.n / 3);
  if (n == 0 || !jac) return 0;
  point_rows_kernel<<<(n * 32 + 255) / 256, 256, 0, st>>>(P->d_point_rows.p, n, jac);
  CUDA_OK(cudaGetLastError());
  return 1;
}

void fill_time(const mrhyde_b200_time* t, TimeDev& td, bool device_ptrs_ok) {
  std::memset(&td, 0, sizeof(td));
  td.alpha_u = 1.0;
  td.seed_u = 1.0;
  td.deltat = 1.0;
  if (!t) return;
  td.time = t->time;
  if (t->nstages <= 0) {        // steady evaluation at a given time
    if (t->seed_what > 1) fail(MRHYDE_B200_ERR_INVALID, "time: seed_what 2 / 3 (previous step / stage) needs a transient evaluation");
    return;
  }
  if (t->stage < 0 || t->stage >= t->nstages || t->nstages > MAX_STAGE) fail(MRHYDE_B200_ERR_INVALID, "time: stage index / stage count out of range");
  if (t->nbdf < 2 || t->nbdf - 1 > MAX_PREV) fail(MRHYDE_B200_ERR_INVALID, "time: unsupported number of BDF weights");
  if (!t->butcher_A || !t->butcher_b || !t->butcher_c || !t->bdf_wts || !t->sol_prev) fail(MRHYDE_B200_ERR_INVALID, "time: missing tables");
  const int s = t->stage, ns = t->nstages;
  td.transient = 1;
  td.deltat = t->deltat;
  td.alpha_u = t->butcher_A[s * ns + s] / t->butcher_b[s];
  td.one_minus_alpha_u = 1.0 - td.alpha_u;
  td.timewt = 1.0 / t->deltat / t->butcher_b[s];
  td.alpha_t = t->bdf_wts[0] * td.timewt;
  td.nprev = t->nbdf - 1;
  for (int k = 1; k < t->nbdf; ++k) td.bdf[k] = t->bdf_wts[k];
  td.nstage_lo = s;
  for (int k = 0; k < s; ++k) td.stage_w[k] = t->butcher_A[s * ns + k] / t->butcher_b[k];
  td.time = t->time + t->butcher_c[s] * t->deltat;
  // derivative seeds (computeSolnTransientSeeded, workset.cpp:622-785): the chain-rule factors of the seeded vector in u and u_t, accumulated
  // in the order Sacado accumulates them
  td.seed_u = td.alpha_u; td.seed_t = td.alpha_t;
  if (t->seed_what == 2) {        // previous step seed_index
    if (t->seed_index < 0 || t->seed_index >= td.nprev) fail(MRHYDE_B200_ERR_INVALID, "time: seed_index is not a previous step of this BDF formula");
    td.seed_u = 0.0;
    if (t->seed_index == 0) { td.seed_u = td.one_minus_alpha_u; for (int k = 0; k < s; ++k) td.seed_u += td.stage_w[k] * -1.0; }
    td.seed_t = td.bdf[t->seed_index + 1] * td.timewt;
  } else if (t->seed_what == 3) { // previous stage seed_index
    if (t->seed_index < 0 || t->seed_index >= s) fail(MRHYDE_B200_ERR_INVALID, "time: seed_index is not a stage below the current one");
    td.seed_u = td.stage_w[t->seed_index];
    td.seed_t = 0.0;
  } else if (t->seed_what != 0 && t->seed_what != 1) {
    fail(MRHYDE_B200_ERR_INVALID, "time: seed_what must be 0 / 1 (stage solution), 2 (previous step) or 3 (previous stage)");
  }
  if (device_ptrs_ok) {
    for (int k = 0; k < td.nprev; ++k) { td.prev[k] = t->sol_prev[k]; if (!td.prev[k]) fail(MRHYDE_B200_ERR_INVALID, "time: null sol_prev vector"); }
    if (s > 0 && !t->sol_stage) fail(MRHYDE_B200_ERR_INVALID, "time: missing sol_stage");
    for (int k = 0; k < s; ++k) { td.stg[k] = t->sol_stage[k]; if (!td.stg[k]) fail(MRHYDE_B200_ERR_INVALID, "time: null sol_stage vector"); }
  }
}

constexpr size_t EV_RING = 64;
void fold_oldest(mrhyde_b200_plan* P) {
  auto& pr = P->ev[P->ev_next];
  CUDA_OK(cudaEventSynchronize(pr.second));
  float ms = 0.f;
  CUDA_OK(cudaEventElapsedTime(&ms, pr.first, pr.second));
  P->ev_ms += ms; ++P->ev_count;
  P->ev_next = (P->ev_next + 1) % EV_RING; --P->ev_used;
}
void record_begin(mrhyde_b200_plan* P, cudaStream_t st, size_t& slot) {
  if (P->ev.empty())
    for (size_t i = 0; i < EV_RING; ++i) {
      cudaEvent_t a, b;
      CUDA_OK(cudaEventCreate(&a)); CUDA_OK(cudaEventCreate(&b));
      P->ev.push_back({a, b});
    }
  if (P->ev_used == EV_RING) fold_oldest(P);
  slot = (P->ev_next + P->ev_used) % EV_RING;  // ev_next = oldest pending pair
  CUDA_OK(cudaEventRecord(P->ev[slot].first, st));
}
void record_end(mrhyde_b200_plan* P, cudaStream_t st, size_t slot) {
  CUDA_OK(cudaEventRecord(P->ev[slot].second, st));
  ++P->ev_used;
}

// dynamic shared memory and launch-bound CTAs per SM of a specialised build (the metric ring is smaller in steady builds)
size_t variant_smem(const mrhyde_b200_plan* P, bool transient) { return P->metric_ng > 0 ? P->smem_metric[transient ? 1 : 0] : P->smem; }
int variant_min_blocks(const mrhyde_b200_plan* P, size_t smem) {
  const int want = std::stoi(opt(P, "min blocks", "0"));
  if (want > 0) return want;
  // the kernels want ~128 registers: more than "max blocks" (3) CTAs of 160 threads per SM would force spills
  const int cap_regs = std::max(1, std::stoi(opt(P, "max blocks", "3")));
  return std::max(1, std::min(std::min(cap_regs, 2048 / P->threads), (int)((228 * 1024) / (smem + 1024))));
}

// Specialised kernel for (steady | transient, output mode); compiled by NVRTC on first use.
// the plan's translation unit with the markers of one build replaced: steady | transient, output mode, start-up stagger
std::string specialise_source(const mrhyde_b200_plan* P, bool transient, int mode, std::string& log) {
  std::string src = P->jit_source;
  auto swap_define = [&](const std::string& mark, const std::string& with) {
    const size_t at = src.find(mark);
    if (at == std::string::npos) return false;
    src.replace(at, mark.size(), with);
    return true;
  };
  if (!swap_define("#define MRH_JIT_TRANSIENT 0  /*@transient@*/", std::string("#define MRH_JIT_TRANSIENT ") + (transient ? "1" : "0")) ||
      !swap_define("#define MRH_JIT_MODE 7  /*@mode@*/", "#define MRH_JIT_MODE " + std::to_string(mode))) {
    log = "specialisation markers missing from the generated source";
    return std::string();
  }
  const int stagger = std::max(0, std::stoi(opt(P, "stagger ns", "0")));
  int n_sm = 148;
  if (P->device >= 0) cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, P->device);
  const int blocks = variant_min_blocks(P, variant_smem(P, transient));
  swap_define("/*@stagger@*/", "#define MRH_JIT_STAGGER_NS " + std::to_string(stagger) + "u\n#define MRH_JIT_STAGGER_SMS " + std::to_string(n_sm) +
                               "u\n#define MRH_JIT_STAGGER_BLOCKS " + std::to_string(blocks) + "u\n#define MRH_JIT_STAGGER_SLOTS " + std::to_string(n_sm * blocks) + "u");
  return src;
}

const JitKernel* jit_variant(mrhyde_b200_plan* P, bool transient, int mode, std::string& log) {
  const int key = (transient ? 8 : 0) + mode;
  auto it = P->jit.find(key);
  if (it != P->jit.end()) return it->second.get();
  const std::string src = specialise_source(P, transient, mode, log);
  if (src.empty()) return nullptr;
  std::unique_ptr<JitKernel> k(new JitKernel());
  const size_t smem = variant_smem(P, transient);
  if (!k->build(src, P->dim == 3 ? "mrh_thermal_q1_3d" : "mrh_thermal_q1_2d", P->threads, variant_min_blocks(P, smem), smem, log, std::stoi(opt(P, "max registers", "0")))) return nullptr;
  const JitKernel* raw = k.get();
  P->jit[key] = std::move(k);
  return raw;
}

void do_assemble(mrhyde_b200_plan* P, const double* sol, const TimeDev& td, bool want_jac, bool want_res, double* res, double* jac, cudaStream_t st, bool adjoint = false) {
  if (!P->finalized) fail(MRHYDE_B200_ERR_STATE, "assemble called before mrhyde_b200_plan_finalize");
  if (P->device == -1) fail(MRHYDE_B200_ERR_STATE, "host-only analysis plan (device = -1) cannot assemble: there is no CPU path");
  if (!sol) fail(MRHYDE_B200_ERR_INVALID, "sol is null");
  if (want_res && !res) fail(MRHYDE_B200_ERR_INVALID, "res is null");
  if (want_jac && !jac) fail(MRHYDE_B200_ERR_INVALID, "jac_values is null");
  CUDA_OK(cudaSetDevice(P->device));
  NvtxRange total(want_jac ? (want_res ? "MrHyDE::AssemblyManager::computeJacRes() - total assembly" : "MrHyDE::AssemblyManager::computeJac() - Jacobian assembly")
                           : "MrHyDE::AssemblyManager::computeRes() - residual assembly");
  OutDev out;
  out.res = want_res ? res : nullptr;
  out.jac = want_jac ? jac : nullptr;
  out.accumulate = P->accumulate ? 1 : 0;
  out.diag_one = opt_bool(P, "use strong DBCs", true) ? 1 : 0;   // the reference calls setJacobianConstraints only then (assemblyManager_constraints.hpp:250)
  GraphDev G{P->d_rowptr.p, P->d_colind.p, P->d_fixed.p};
  const bool volume = opt_bool(P, "assemble volume terms", true);
  int launched = 0;
  if (P->use_general) {
    const bool bnd = opt_bool(P, "assemble boundary terms", true);
    if (!P->accumulate && !volume) fail(MRHYDE_B200_ERR_UNSUPPORTED, "accumulate=false needs the volume pass (it defines every entry)");
    size_t slot = 0;
    record_begin(P, st, slot);
    GenLaunchStats stats;
    NvtxRange physics("MrHyDE::AssemblyManager::computeJacRes() - physics evaluation");   // gather + physics + boundary + scatter, fused
    const char* err = gen_assemble(P->gen_dev, P->gen, P->gen_kernels, P->d_vx.p, P->d_vy.p, P->d_vz.p, P->d_conn.p, P->d_lids.p, G, out, sol, td, volume, bnd, st, &stats, adjoint);
    if (err) fail(MRHYDE_B200_ERR_CUDA, std::string("general assembly launch: ") + err);
    record_end(P, st, slot);
    launched = stats.launches;
    if (want_jac && P->accumulate && P->d_fixed_diag.n > 0 && opt_bool(P, "use strong DBCs", true)) {
      NvtxRange dbc("MrHyDE::AssemblyManager::dofConstraints()");
      const int n = (int)P->d_fixed_diag.n;
      fixed_diag_kernel<<<(n + 255) / 256, 256, 0, st>>>(P->d_fixed_diag.p, n, jac);
      ++launched;
      CUDA_OK(cudaGetLastError());
    }
    if (want_jac) launched += point_constraints(P, jac, st);
    P->launches_per_assemble = launched;
    return;
  }
  if (volume) {
    NvtxRange physics("MrHyDE::AssemblyManager::computeJacRes() - physics evaluation");   // gather + physics + scatter + constraints, one launch
    size_t slot = 0;
    record_begin(P, st, slot);
    const void* params;
    if (P->dim == 3) { P->th3.sol = sol; P->th3.td = td; P->th3.out = out; params = &P->th3; }
    else { P->th2.sol = sol; P->th2.td = td; P->th2.out = out; params = &P->th2; }
    // in-kernel halo push: the assembly kernel stores ghost rows into the owner's slab as it completes them
    PushDev push;
    std::memset(&push, 0, sizeof(push));
    if (P->push_ok && !P->point_on_ghost && P->use_jit && P->halo && !P->suppress_overlap && want_jac && want_res && !adjoint && P->boundary.groups.empty() && P->cp.orphan_rows.empty() &&
        !opt_bool(P, "overlap halo", false)) {
      if (P->halo->push_params(res, jac, P->mesh.nowned, P->mesh.rowptr[(size_t)P->mesh.nowned], P->cp.n_early_chains, push)) ++P->pushed_assembles;
    }
    if (P->dim == 3) P->th3.push = push; else P->th2.push = push;
    const JitKernel* jk = nullptr;
    if (P->use_jit) {
      const int mode = (out.res ? 1 : 0) | (out.jac ? 2 : 0) | (out.accumulate ? 4 : 0);
      std::string log;
      jk = jit_variant(P, td.transient != 0, mode, log);
      if (!jk) fail(MRHYDE_B200_ERR_CUDA, "jit: kernel build failed: " + log);
    }
    // Multi-rank plans with option "overlap halo": the chains that complete the ghost rows run first (plan.cpp gives them the
    // lowest ids); their rows are final after that launch -- every row is written by exactly one chain -- so the exchange with
    // the owners starts there and runs beside the rest of the assembly.  mrhyde_b200_halo_sum then only waits and adds.
    const int n_early = P->cp.n_early_chains;
    const bool overlap = !P->suppress_overlap && opt_bool(P, "overlap halo", false) && P->halo && P->halo->can_start() && want_jac && want_res &&
                         n_early > 0 && n_early < P->cp.n_chains && P->boundary.groups.empty() && P->cp.orphan_rows.empty();
    if (P->halo) P->halo->drain(st);
    auto launch_range = [&](int first, int count) -> const char* {
      if (P->dim == 3) P->th3.chains.chain_offset = first; else P->th2.chains.chain_offset = first;
      return P->use_jit ? jk->launch(params, count, P->threads, variant_smem(P, td.transient != 0), st)
                        : launch_thermal_q1_aot(P->dim, params, count, P->threads, P->smem, st);
    };
    const char* lerr = nullptr;
    if (overlap) {
      lerr = launch_range(0, n_early);
      if (!lerr) {
        std::string herr;
        if (!P->halo->start(res, jac, st, herr)) fail(MRHYDE_B200_ERR_NCCL, herr);
        ++P->overlapped_assembles;
        lerr = launch_range(n_early, P->cp.n_chains - n_early);
        ++launched;
      }
    } else {
      lerr = launch_range(0, P->cp.n_chains);
    }
    if (lerr) fail(MRHYDE_B200_ERR_CUDA, std::string("volume kernel launch: ") + lerr);
    record_end(P, st, slot);
    ++launched;
    CUDA_OK(cudaGetLastError());
    if (!P->accumulate && !P->cp.orphan_rows.empty()) {
      const int n = (int)P->cp.orphan_rows.size();
      orphan_rows_kernel<<<(n + 127) / 128, 128, 0, st>>>(P->d_orphans.p, n, G, out);
      ++launched;
    }
  } else if (!P->accumulate) {
    fail(MRHYDE_B200_ERR_UNSUPPORTED, "accumulate=false needs the volume pass (it defines every entry)");
  }
  if (!P->boundary.groups.empty() && opt_bool(P, "assemble boundary terms", true)) {
    NvtxRange boundary("MrHyDE::AssemblyManager::computeJacRes() - boundary evaluation");
    OutDev bout = out;
    bout.accumulate = 1;  // boundary groups always add on top of the volume result
    launch_boundary(P->boundary, sol, td, G, bout, st, adjoint);
    launched += (int)P->boundary.groups.size();
    CUDA_OK(cudaGetLastError());
  }
  if (want_jac && P->accumulate && P->d_fixed_diag.n > 0 && opt_bool(P, "use strong DBCs", true)) {
    NvtxRange dbc("MrHyDE::AssemblyManager::dofConstraints()");
    const int n = (int)P->d_fixed_diag.n;
    fixed_diag_kernel<<<(n + 255) / 256, 256, 0, st>>>(P->d_fixed_diag.p, n, jac);
    ++launched;
    CUDA_OK(cudaGetLastError());
  }
  if (want_jac) launched += point_constraints(P, jac, st);
  P->launches_per_assemble = launched;
}


// ---- general path set-up --------------------------------------------------------------------------------------
int bc_code(const std::string& t) {
  if (t == "Dirichlet") return BC_DIRICHLET;
  if (t == "weak Dirichlet") return BC_WEAK_DIRICHLET;
  if (t == "Neumann") return BC_NEUMANN;
  return BC_NONE;
}

// reference tables of kernel basis kb in the uniform layout [card][nq][ncb] (general_physics.cuh)
void append_ref_tab(const BasisCopy& B, const std::vector<double>& val, const std::vector<double>& grad, const std::vector<double>& curl,
                    const std::vector<double>& div, int dim, int nq, int ncb, std::vector<double>& out) {
  for (int i = 0; i < B.card; ++i)
    for (int q = 0; q < nq; ++q) {
      double e[6] = {0, 0, 0, 0, 0, 0};
      if (B.type == "HGRAD") {
        e[0] = val[(size_t)i * nq + q];
        for (int d = 0; d < dim; ++d) e[1 + d] = grad.empty() ? 0.0 : grad[((size_t)i * nq + q) * dim + d];
      } else if (B.type == "HCURL") {
        for (int d = 0; d < dim; ++d) { e[d] = val[((size_t)i * nq + q) * dim + d]; e[3 + d] = curl.empty() ? 0.0 : curl[((size_t)i * nq + q) * dim + d]; }
      } else {  // HDIV
        for (int d = 0; d < dim; ++d) e[d] = val[((size_t)i * nq + q) * dim + d];
        e[3] = div.empty() ? 0.0 : div[(size_t)i * nq + q];
      }
      for (int k = 0; k < ncb; ++k) out.push_back(e[k]);
    }
}

void compile_functions(const FunctionSet& fs, const std::vector<std::string>& names, GenFnRec* rec, std::vector<uint8_t>& ops, std::vector<double>& cs) {
  for (size_t f = 0; f < names.size(); ++f) {
    GenFnRec r;
    r.begin = 0; r.n = 0; r.is_const = 1; r.pad = 0; r.cval = 0.0;
    if (!names[f].empty()) {
      const LongProgram lp = fs.compile_long(names[f]);
      r.is_const = lp.is_const ? 1 : 0; r.cval = lp.cval;
      r.pad = (lp.uses_state ? 1 : 0) | (lp.uses_reduction ? 2 : 0);   // bit 0: reads a solution field (evaluated after the fields, differentiated in the Jacobian stages); bit 1: element reduction
      if (!lp.is_const) {
        r.begin = (int32_t)ops.size(); r.n = (int32_t)lp.op.size();
        ops.insert(ops.end(), lp.op.begin(), lp.op.end());
        cs.insert(cs.end(), lp.c.begin(), lp.c.end());
      }
    }
    rec[f] = r;
  }
}

void finalize_general(mrhyde_b200_plan* P, const std::string& phys) {
  const bool host_only = (P->device == -1);
  MeshGraph& M = P->mesh;
  GeneralPlanHost& H = P->gen;
  const int dim = P->dim;
  const int order = P->bases[(size_t)P->var_basis[0]].order;
  const int nqs = P->bgroups.empty() ? 0 : P->bgroups[0].nqp;
  P->gen_host = gen_find_host(phys, dim, order, P->nqp, nqs);   // null unless the test-only emulator library is registered
  P->gen_kernels = gen_find_device(phys, dim, order, P->nqp, nqs);   // the instantiation table (no CUDA call: host-only plans read its sizes too)
  if (!P->gen_kernels)
    fail(MRHYDE_B200_ERR_UNSUPPORTED, "no kernel for physics '" + P->physics + "' dim " + std::to_string(dim) + " order " + std::to_string(order) + " with " +
                                          std::to_string(P->nqp) + " volume / " + std::to_string(nqs) + " side points; built: " + gen_supported_list());
  const GenKernelInfo& I = P->gen_kernels->info;
  H.info = I;
  // ---- the block must be the module's own variable / basis layout
  const std::vector<std::string> want = module_variables(phys, dim);
  if ((int)want.size() != P->nvars || I.nvars != P->nvars) fail(MRHYDE_B200_ERR_UNSUPPORTED, "general path: the block's variables are not the module's (" + phys + ")");
  for (int v = 0; v < P->nvars; ++v) if (P->var_names[(size_t)v] != want[(size_t)v]) fail(MRHYDE_B200_ERR_UNSUPPORTED, "general path: variable order must be the module's: expected '" + want[(size_t)v] + "', got '" + P->var_names[(size_t)v] + "'");
  if (P->ndof_elem != I.N) fail(MRHYDE_B200_ERR_INVALID, "general path: ndof_elem does not match the module's basis cardinalities");
  static const char* tname[4] = {"HGRAD", "HCURL", "HDIV", "HVOL"};
  std::vector<int> kb_desc((size_t)I.nbasis, -1);   // kernel basis -> descriptor basis
  for (int v = 0; v < P->nvars; ++v) {
    const int kb = I.var_basis[v];
    const BasisCopy& B = P->bases[(size_t)P->var_basis[(size_t)v]];
    const int bt = phys == "maxwell" ? (kb == 0 ? BT_HCURL : BT_HDIV) : BT_HGRAD;
    if (B.type != tname[bt] || B.card != I.card[kb] || B.order != order) fail(MRHYDE_B200_ERR_UNSUPPORTED, "general path: variable '" + P->var_names[(size_t)v] + "' is not on the basis the module expects");
    if (kb_desc[(size_t)kb] < 0) kb_desc[(size_t)kb] = P->var_basis[(size_t)v];
    else if (kb_desc[(size_t)kb] != P->var_basis[(size_t)v]) fail(MRHYDE_B200_ERR_UNSUPPORTED, "general path: variables of one basis type must share one basis table");
  }
  std::memset(H.off, 0, sizeof(H.off));
  {
    std::vector<int> seen((size_t)I.N, 0);
    for (int v = 0; v < P->nvars; ++v)
      for (int i = 0; i < I.card[I.var_basis[v]]; ++i) {
        const int32_t o = P->offsets[(size_t)v * P->max_card + i];
        if (o < 0 || o >= I.N || seen[(size_t)o]++) fail(MRHYDE_B200_ERR_INVALID, "general path: offsets are not a permutation of the element dofs");
        H.off[v][i] = (int16_t)o;
      }
  }
  // ---- tables at the volume points
  auto geo_tables = [&](const std::vector<double>& pts, int nq, std::vector<double>& gN, std::vector<double>& gdN) {
    const int nv = 1 << dim;
    gN.assign((size_t)nq * nv, 0.0); gdN.assign((size_t)nq * nv * dim, 0.0);
    for (int q = 0; q < nq; ++q) geom_shape(dim, &pts[(size_t)q * dim], &gN[(size_t)q * nv], &gdN[(size_t)q * nv * dim]);
  };
  geo_tables(P->qp_pts, P->nqp, H.geo_N, H.geo_dN);
  H.qwts = P->qp_wts;
  H.ref_tab.clear();
  for (int kb = 0; kb < I.nbasis; ++kb) {
    const BasisCopy& B = P->bases[(size_t)kb_desc[(size_t)kb]];
    if (B.type == "HGRAD" && B.grad.empty()) fail(MRHYDE_B200_ERR_INVALID, "general path: HGRAD basis without reference gradients");
    if (B.type == "HCURL" && B.curl.empty()) fail(MRHYDE_B200_ERR_INVALID, "general path: HCURL basis without reference curls");
    if (B.type == "HDIV" && B.div.empty()) fail(MRHYDE_B200_ERR_INVALID, "general path: HDIV basis without reference divergences");
    append_ref_tab(B, B.val, B.grad, B.curl, B.div, dim, P->nqp, I.ncb[kb], H.ref_tab);
  }
  // ---- options
  std::memset(&H.opt, 0, sizeof(H.opt));
  H.opt.form_param = std::stod(opt(P, "form_param", "1.0"));
  H.opt.penalty = std::stod(opt(P, "penalty", "10.0"));
  H.opt.have_advection = opt_bool(P, "include advection", false) ? 1 : 0;
  H.opt.useSUPG = opt_bool(P, "useSUPG", false) ? 1 : 0;
  H.opt.usePSPG = opt_bool(P, "usePSPG", false) ? 1 : 0;
  H.opt.uz_reference = opt(P, "ns3d_uz_rows", "reference") == "corrected" ? 0 : 1;
  H.opt.incplanestress = opt_bool(P, "incplanestress", false) ? 1 : 0;
  H.opt.leapfrog = opt_bool(P, "use leap frog", false) ? 1 : 0;
  // ---- coefficient functions at the volume points
  std::vector<std::string> fnames;
  for (const ModuleFn* f = module_functions(phys); f->key; ++f) fnames.push_back(f->key);
  if ((int)fnames.size() != I.nfn) fail(MRHYDE_B200_ERR_INVALID, "general path: module function table out of step with the kernel");
  H.fn_op.clear(); H.fn_c.clear();
  {
    FunctionSet fs = make_function_set(P, false, true);
    std::vector<std::string> names = fnames;
    names.resize(GEN_MAXFN);
    compile_functions(fs, names, H.fn, H.fn_op, H.fn_c);
    // initial conditions in (variable, component) order: the function slots of the projection mode (gen_initial_point)
    std::vector<std::string> inames;
    for (int v = 0; v < P->nvars; ++v) for (const std::string& nm : initial_function_names(P, v)) inames.push_back(nm);
    if ((int)inames.size() > I.nfn) fail(MRHYDE_B200_ERR_UNSUPPORTED, "general path: more initial-condition components than function slots");
    inames.resize(GEN_MAXFN);
    compile_functions(fs, inames, H.init_fn, H.fn_op, H.fn_c);
  }
  // ---- boundary families
  H.sides.clear();
  int64_t inst = M.nelem;
  if (!P->bgroups.empty()) {
    FunctionSet fss = make_function_set(P, true, true);
    for (auto& g : P->bgroups) {
      GenSideFamily S;
      S.sideset = g.sideset; S.local_side = g.local_side; S.nqs = g.nqp;
      if (g.nqp != I.nqs) fail(MRHYDE_B200_ERR_UNSUPPORTED, "general path: boundary groups must all use the side rule the kernel was built for");
      S.items = g.elem_ids;
      for (int32_t e : S.items) if (e < 0 || e >= M.nelem) fail(MRHYDE_B200_ERR_INVALID, "boundary group: element id out of range");
      for (int d = 0; d < 3; ++d) { S.tan_u[d] = g.tu[d]; S.tan_v[d] = g.tv[d]; }
      const std::string& sname = P->side_names[(size_t)g.sideset];
      std::vector<std::string> names = fnames;
      names.resize(GEN_MAXFN);
      for (int v = 0; v < P->nvars; ++v) {
        auto it = P->bcs.find({P->var_names[(size_t)v], sname});
        const std::string t = it == P->bcs.end() ? "none" : it->second.type;
        S.bc_type[v] = bc_code(t);
        if (S.bc_type[v] == BC_WEAK_DIRICHLET) { names[(size_t)I.nfn + v] = "Dirichlet " + P->var_names[(size_t)v] + " " + sname; S.bc_fn[v] = I.nfn + v; S.active = true; }
        else if (S.bc_type[v] == BC_NEUMANN) { names[(size_t)I.nfn + v] = "Neumann " + P->var_names[(size_t)v] + " " + sname; S.bc_fn[v] = I.nfn + v; S.active = true; }
      }
      compile_functions(fss, names, S.fn, H.fn_op, H.fn_c);
      geo_tables(g.pts, g.nqp, S.geo_N, S.geo_dN);
      S.qwts = g.wts;
      for (int kb = 0; kb < I.nbasis; ++kb) {
        const size_t db = (size_t)kb_desc[(size_t)kb];
        const BasisCopy& B = P->bases[db];
        if (db >= g.bval.size()) fail(MRHYDE_B200_ERR_INVALID, "boundary group: side basis tables missing");
        static const std::vector<double> none;
        append_ref_tab(B, g.bval[db], g.bgrad[db], none, none, dim, g.nqp, I.ncb[kb], S.ref_tab);
      }
      if (S.active) { S.inst_base = inst; inst += (int64_t)S.items.size(); }
      H.sides.push_back(std::move(S));
    }
  }
  // ---- pull schedule
  // one batch by default: measured on the B200, splitting the element range so that a batch of element matrices stays
  // L2-resident costs more in launch tails than the pull saves in DRAM reads (profiles/r01_general_bench.jsonl)
  const int64_t batch_elems = std::stoll(opt(P, "batch elems", "-1"));
  // element scratch budget: option "scratch GB" or a quarter of the free device memory; a plan that fits runs one batch, otherwise
  // the element range is cut into batches whose scratch ring fits (plan_stat "general_scratch_bytes", "general_batches")
  int64_t budget = (int64_t)(std::stod(opt(P, "scratch GB", "0")) * 1e9);
  if (budget <= 0 && !host_only) {
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) budget = (int64_t)(fr / 4);
  }
  try {
    gen_build_pull(M, I.N, H.sides, batch_elems, budget, H);
  } catch (const std::exception& e) {
    fail(MRHYDE_B200_ERR_INVALID, e.what());
  }
  H.epb_override = std::stoi(opt(P, "elements per cta", "0"));
  H.lump_mass = opt_bool(P, "lump mass", false);
  {
    // which build of the element kernel assembles Jacobians: tensor = field-direction derivatives + FP64 tensor-core contraction
    // (single-basis HGRAD modules), lanes = one derivative lane per element dof (every module)
    const std::string jm = opt(P, "jacobian", "auto");
    if (jm != "auto" && jm != "tensor" && jm != "lanes") fail(MRHYDE_B200_ERR_INVALID, "option jacobian must be auto|tensor|lanes");
    if (jm == "tensor" && !I.tensor) fail(MRHYDE_B200_ERR_UNSUPPORTED, "jacobian=tensor: the module's basis layout has no tensor-core build");
    // auto = lanes: measured on the B200 (profiles/r02_s7_general_tensor_vs_lanes.log), the FP64 tensor-core contraction does not beat
    // the derivative lanes -- DMMA.8x8x4 runs at the FP64 FMA rate, the 27 -> 32 padding of hex-Q2 costs 40 % more FMAs, and the
    // fragment set-up of 8 x 8 tiles outweighs the saved issue slots on Q1
    H.use_tensor = I.tensor && jm == "tensor";
  }
  P->use_general = true;
  P->launches_per_assemble = 2 * (int)H.batches.size();
  for (auto& S : H.sides) if (S.active && !S.items.empty()) ++P->launches_per_assemble;
  if (host_only) { P->finalized = true; return; }
  // ---- upload
  size_t* tot = &P->dev_bytes;
  P->d_vx.upload(M.vcoord[0], tot); P->d_vy.upload(M.vcoord[1], tot); P->d_vz.upload(M.vcoord[2], tot);
  P->d_conn.upload(M.conn, tot); P->d_lids.upload(M.lids, tot);
  P->d_rowptr.upload(M.rowptr, tot); P->d_colind.upload(M.colind, tot); P->d_fixed.upload(M.fixed, tot);
  {
    std::vector<int64_t> diag;
    for (int64_t r = 0; r < M.nrows; ++r) {
      if (!M.fixed[(size_t)r] || r >= M.nowned) continue;
      int64_t pos = -1;
      for (int64_t p = M.rowptr[(size_t)r]; p < M.rowptr[(size_t)r + 1]; ++p) if (M.colind[(size_t)p] == r) pos = p;
      diag.push_back(pos);
    }
    P->d_fixed_diag.upload(diag, tot);
    if (diag.empty()) P->d_fixed_diag.n = 0;
  }
  std::string err;
  P->gen_dev = gen_upload(H, M, tot, err);
  if (!P->gen_dev) fail(MRHYDE_B200_ERR_CUDA, err);
  if (P->d_fixed_diag.n > 0) P->launches_per_assemble += 1;
  P->finalized = true;
}

}  // namespace

// Host replay of the metric ring: element metrics by the formulas of metric_cell (volume_kernel.cuh), then the plan's
// metric source words (host_apply_metric_plan).  Checks the plan analysis on machines without a GPU; never reached by assemble_*.
template <int DIM>
static void metric_host_elements(const mrhyde_b200_plan* P, const ThermalParams<DIM>& TH, const double* sol, const TimeDev& td, std::vector<double>& out) {
  typedef Q1Shape<DIM> S;
  constexpr int NV = S::NV, NQ = S::NQ, NG = S::NG, SL = NG + 1 + 3 * NV;
  const MeshGraph& M = P->mesh;
  static const int nb[4] = {0, 1, 3, 4};
  auto fn = [&](const ExprProgram& p, const double* x) {
    if (p.is_const) return p.cval;
    const double v[7] = {x[0], x[1], x[2], td.time, 0.0, 0.0, 0.0};
    return FunctionSet::eval_host(p, v);
  };
  const double xz[3] = {0, 0, 0};
  const double kap = fn(TH.diffusion, xz), rc = fn(TH.density, xz) * fn(TH.specific_heat, xz);
  out.assign((size_t)M.nelem * SL, 0.0);
  for (int64_t e = 0; e < M.nelem; ++e) {
    double* o = &out[(size_t)e * SL];
    double X[DIM + 1][DIM], J[DIM][DIM], Ji[DIM][DIM];
    for (int v = 0; v <= DIM; ++v) for (int d = 0; d < DIM; ++d) X[v][d] = M.vcoord[d][(size_t)M.conn[(size_t)e * NV + nb[v]]];
    for (int d = 0; d < DIM; ++d) for (int a = 0; a < DIM; ++a) J[d][a] = 0.5 * (X[a + 1][d] - X[0][d]);
    double det;
    if (DIM == 2) {
      det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
      Ji[0][0] = J[1][1] / det; Ji[0][1] = -J[0][1] / det; Ji[1][0] = -J[1][0] / det; Ji[1][1] = J[0][0] / det;
    } else {
      double c[3][3];
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b)
        c[a][b] = J[(a + 1) % 3][(b + 1) % 3] * J[(a + 2) % 3][(b + 2) % 3] - J[(a + 1) % 3][(b + 2) % 3] * J[(a + 2) % 3][(b + 1) % 3];
      det = J[0][0] * c[0][0] + J[0][1] * c[0][1] + J[0][2] * c[0][2];
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) Ji[a][b] = c[b][a] / det;
    }
    const double adet = std::fabs(det), kd = kap * adet;
    int g = 0;
    for (int a = 0; a < DIM; ++a) { double t = 0; for (int d = 0; d < DIM; ++d) t += Ji[a][d] * Ji[a][d]; o[g++] = t * kd; }
    for (int a = 0; a < DIM; ++a) for (int b = a + 1; b < DIM; ++b) { double t = 0; for (int d = 0; d < DIM; ++d) t += Ji[a][d] * Ji[b][d]; o[g++] = t * kd; }
    o[NG] = rc * adet;
    for (int q = 0; q < NQ; ++q) {
      double x[3] = {0, 0, 0};
      for (int d = 0; d < DIM; ++d) { x[d] = X[0][d]; for (int a = 0; a < DIM; ++a) x[d] += J[d][a] * (TH.tab.qpt[q][a] + 1.0); }
      const double fw = fn(TH.source, x) * TH.tab.qw[q] * adet;
      for (int i = 0; i < NV; ++i) o[NG + 1 + i] += fw * TH.tab.phi[q][i];
    }
    for (int j = 0; j < NV; ++j) {
      const int32_t lid = M.lids[(size_t)e * NV + j];
      double u = sol[lid], ut = 0.0;
      if (td.transient) {
        const double p0 = td.prev[0][lid];
        double bu = td.one_minus_alpha_u * p0;
        for (int k = 0; k < td.nstage_lo; ++k) bu += td.stage_w[k] * (td.stg[k][lid] - p0);
        u = td.alpha_u * sol[lid] + bu;
        double bt = td.bdf[1] * p0;
        for (int k = 2; k <= td.nprev; ++k) bt += td.bdf[k] * td.prev[k - 1][lid];
        ut = td.alpha_t * sol[lid] + bt * td.timewt;
      }
      o[NG + 1 + NV + j] = u; o[NG + 1 + 2 * NV + j] = ut;
    }
  }
}

static const GenHostKernels* need_emulator(mrhyde_b200_plan* P) {
  if (!P->gen_host) {
    const GenKernelInfo& I = P->gen.info;
    P->gen_host = gen_find_host(I.physics, I.dim, I.order, I.nq, I.nqs);
  }
  if (!P->gen_host) fail(MRHYDE_B200_ERR_STATE, "debug_emulate: the test-only library libmrhyde_b200_emulate.so is not registered (mrhyde_b200_debug_set_emulator)");
  return P->gen_host;
}

extern "C" {

int mrhyde_b200_debug_set_emulator(void* lookup) {
  gen_set_emulator(reinterpret_cast<GenEmulatorLookup>(lookup));
  return MRHYDE_B200_OK;
}

const char* mrhyde_b200_version(void) { return "mrhyde_b200 0.1 (sm_100a)"; }
const char* mrhyde_b200_last_error(void) { return g_error.c_str(); }

int mrhyde_b200_plan_create(mrhyde_b200_plan** plan, const mrhyde_b200_desc* d, int device) {
  ABI_BEGIN
  if (!plan || !d) fail(MRHYDE_B200_ERR_INVALID, "plan_create: null argument");
  *plan = nullptr;
  if (!d->physics || !d->var_names || !d->var_basis || !d->bases || !d->offsets || !d->qp_pts || !d->qp_wts) fail(MRHYDE_B200_ERR_INVALID, "plan_create: null field in descriptor");
  if (d->dim != 2 && d->dim != 3) fail(MRHYDE_B200_ERR_UNSUPPORTED, "plan_create: only Quadrilateral_4 (dim 2) and Hexahedron_8 (dim 3) cells");
  if (d->nvars < 1 || d->nbases < 1 || d->nqp < 1 || d->ndof_elem < 1 || d->max_card < 1) fail(MRHYDE_B200_ERR_INVALID, "plan_create: non-positive size");
  if (device != -1) {  // -1: host-only analysis plan (cannot assemble)
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device || device < 0)
      fail(MRHYDE_B200_ERR_CUDA, "plan_create: CUDA device " + std::to_string(device) + " not available (this library has no CPU path)");
  }
  std::unique_ptr<mrhyde_b200_plan> P(new mrhyde_b200_plan());
  P->device = device;
  P->physics = d->physics;
  P->dim = d->dim; P->nvars = d->nvars; P->ndof_elem = d->ndof_elem; P->max_card = d->max_card; P->nqp = d->nqp;
  for (int v = 0; v < d->nvars; ++v) {
    P->var_names.push_back(d->var_names[v]);
    if (d->var_basis[v] < 0 || d->var_basis[v] >= d->nbases) fail(MRHYDE_B200_ERR_INVALID, "plan_create: var_basis out of range");
    P->var_basis.push_back(d->var_basis[v]);
  }
  P->offsets.assign(d->offsets, d->offsets + (size_t)d->nvars * d->max_card);
  P->qp_pts.assign(d->qp_pts, d->qp_pts + (size_t)d->nqp * d->dim);
  P->qp_wts.assign(d->qp_wts, d->qp_wts + d->nqp);
  for (int b = 0; b < d->nbases; ++b) {
    const mrhyde_b200_basis& in = d->bases[b];
    if (!in.type || !in.val || in.card < 1) fail(MRHYDE_B200_ERR_INVALID, "plan_create: incomplete basis table");
    BasisCopy B;
    B.type = in.type; B.order = in.order; B.card = in.card;
    const int vdim = (B.type == "HCURL" || B.type == "HDIV") ? d->dim : 1;
    B.val.assign(in.val, in.val + (size_t)in.card * d->nqp * vdim);
    if (in.grad) B.grad.assign(in.grad, in.grad + (size_t)in.card * d->nqp * d->dim);
    if (in.curl) B.curl.assign(in.curl, in.curl + (size_t)in.card * d->nqp * d->dim);
    if (in.div) B.div.assign(in.div, in.div + (size_t)in.card * d->nqp);
    P->bases.push_back(std::move(B));
  }
  P->mesh.dim = d->dim; P->mesh.nverts = 1 << d->dim; P->mesh.ndof = d->ndof_elem;
  *plan = P.release();
  ABI_END
}

void mrhyde_b200_plan_destroy(mrhyde_b200_plan* plan) { delete plan; }

int mrhyde_b200_plan_set_function(mrhyde_b200_plan* P, const char* name, const char* expression) {
  ABI_BEGIN
  if (!P || !name || !expression) fail(MRHYDE_B200_ERR_INVALID, "set_function: null argument");
  if (P->finalized) fail(MRHYDE_B200_ERR_STATE, "set_function after finalize");
  P->functions[name] = expression;
  ABI_END
}

int mrhyde_b200_plan_set_option(mrhyde_b200_plan* P, const char* key, const char* value) {
  ABI_BEGIN
  if (key) {
    // experiment keys that compile parts of the kernel out or force a build variant: results may be wrong, so a production host
    // cannot set them by accident -- they need MRHYDE_B200_DEBUG_OPTIONS=1 in the environment
    const std::string k(key);
    if (k == "debug skip" || k == "debug transient" || k == "debug mode") {
      const char* e = getenv("MRHYDE_B200_DEBUG_OPTIONS");
      if (!e || std::string(e) != "1") fail(MRHYDE_B200_ERR_INVALID, "option '" + k + "' is a kernel-debugging key: set MRHYDE_B200_DEBUG_OPTIONS=1 to allow it");
    }
  }
  if (!P || !key || !value) fail(MRHYDE_B200_ERR_INVALID, "set_option: null argument");
  bool known = false;
  for (const char** k = kKnownOptions; *k; ++k) if (std::string(*k) == key) known = true;
  if (!known) fail(MRHYDE_B200_ERR_INVALID, std::string("set_option: unknown key '") + key + "'");
  const std::string k(key), v(value);
  if (P->finalized && k != "accumulate" && k != "assemble boundary terms" && k != "assemble volume terms" && k != "debug transient" && k != "debug mode") fail(MRHYDE_B200_ERR_STATE, "set_option: '" + k + "' must be set before finalize");
  if (k == "accumulate") {
    if (v != "true" && v != "false") fail(MRHYDE_B200_ERR_INVALID, "set_option: accumulate must be true|false");
    P->accumulate = (v == "true");
  }
  if (k == "ns3d_uz_rows" && v != "reference" && v != "corrected") fail(MRHYDE_B200_ERR_INVALID, "set_option: ns3d_uz_rows must be reference|corrected");
  P->options[k] = v;
  ABI_END
}

int mrhyde_b200_plan_set_mesh(mrhyde_b200_plan* P, int64_t n_elem, const double* elem_nodes, const int32_t* lids, const int8_t* orient_sign) {
  ABI_BEGIN
  if (!P || !elem_nodes || !lids) fail(MRHYDE_B200_ERR_INVALID, "set_mesh: null argument");
  if (n_elem < 1 || n_elem > 0x7fffffff) fail(MRHYDE_B200_ERR_INVALID, "set_mesh: element count out of range");
  if (P->finalized) fail(MRHYDE_B200_ERR_STATE, "set_mesh after finalize");
  P->mesh.set_elem_nodes(n_elem, elem_nodes);
  P->mesh.lids.assign(lids, lids + (size_t)n_elem * P->ndof_elem);
  if (orient_sign) P->mesh.orient.assign(orient_sign, orient_sign + (size_t)n_elem * P->ndof_elem); else P->mesh.orient.clear();
  P->have_mesh = true;
  ABI_END
}

int mrhyde_b200_plan_set_mesh_indexed(mrhyde_b200_plan* P, int64_t n_verts, const double* vc, int64_t n_elem, const int32_t* conn,
                                      const int32_t* lids, const int8_t* orient_sign) {
  ABI_BEGIN
  if (!P || !vc || !conn || !lids) fail(MRHYDE_B200_ERR_INVALID, "set_mesh_indexed: null argument");
  if (n_elem < 1 || n_elem > 0x7fffffff || n_verts < 1 || n_verts > 0x7fffffff) fail(MRHYDE_B200_ERR_INVALID, "set_mesh_indexed: count out of range");
  if (P->finalized) fail(MRHYDE_B200_ERR_STATE, "set_mesh after finalize");
  MeshGraph& M = P->mesh;
  M.nelem = n_elem; M.nvert = n_verts;
  for (int d = 0; d < 3; ++d) M.vcoord[d].assign((size_t)n_verts, 0.0);
  for (int64_t v = 0; v < n_verts; ++v) for (int d = 0; d < P->dim; ++d) M.vcoord[d][(size_t)v] = vc[v * P->dim + d];
  M.conn.assign(conn, conn + (size_t)n_elem * M.nverts);
  for (int32_t c : M.conn) if (c < 0 || c >= n_verts) fail(MRHYDE_B200_ERR_INVALID, "set_mesh_indexed: connectivity out of range");
  M.lids.assign(lids, lids + (size_t)n_elem * P->ndof_elem);
  if (orient_sign) M.orient.assign(orient_sign, orient_sign + (size_t)n_elem * P->ndof_elem); else M.orient.clear();
  P->have_mesh = true;
  ABI_END
}

int mrhyde_b200_plan_set_graph(mrhyde_b200_plan* P, int64_t n_rows, int64_t n_owned, const int64_t* row_map, const int32_t* entries, const uint8_t* is_fixed) {
  ABI_BEGIN
  if (!P || !row_map || !entries) fail(MRHYDE_B200_ERR_INVALID, "set_graph: null argument");
  if (n_rows < 1 || n_rows > 0x7fffffff || n_owned < 0 || n_owned > n_rows) fail(MRHYDE_B200_ERR_INVALID, "set_graph: row counts out of range");
  if (P->finalized) fail(MRHYDE_B200_ERR_STATE, "set_graph after finalize");
  MeshGraph& M = P->mesh;
  M.nrows = n_rows; M.nowned = n_owned;
  M.rowptr.assign(row_map, row_map + n_rows + 1);
  if (M.rowptr[0] != 0) fail(MRHYDE_B200_ERR_INVALID, "set_graph: row_map[0] != 0");
  for (int64_t r = 0; r < n_rows; ++r) if (M.rowptr[(size_t)r + 1] < M.rowptr[(size_t)r]) fail(MRHYDE_B200_ERR_INVALID, "set_graph: row_map not monotone");
  M.nnz = M.rowptr[(size_t)n_rows];
  M.colind.assign(entries, entries + M.nnz);
  for (int64_t r = 0; r < n_rows; ++r)
    for (int64_t p = M.rowptr[(size_t)r] + 1; p < M.rowptr[(size_t)r + 1]; ++p)
      if (M.colind[(size_t)p] <= M.colind[(size_t)p - 1]) fail(MRHYDE_B200_ERR_INVALID, "set_graph: column indices must be strictly ascending within a row (fillComplete'd graph)");
  if (is_fixed) M.fixed.assign(is_fixed, is_fixed + n_rows); else M.fixed.assign((size_t)n_rows, 0);
  P->have_graph = true;
  ABI_END
}

int mrhyde_b200_plan_set_sidesets(mrhyde_b200_plan* P, int32_t n_sides, const char* const* side_names) {
  ABI_BEGIN
  if (!P || (n_sides > 0 && !side_names)) fail(MRHYDE_B200_ERR_INVALID, "set_sidesets: null argument");
  P->side_names.clear();
  for (int s = 0; s < n_sides; ++s) P->side_names.push_back(side_names[s]);
  ABI_END
}

int mrhyde_b200_plan_set_bc(mrhyde_b200_plan* P, const char* var, const char* side, const char* type, const char* expression) {
  ABI_BEGIN
  if (!P || !var || !side || !type) fail(MRHYDE_B200_ERR_INVALID, "set_bc: null argument");
  const std::string t(type);
  if (t != "none" && t != "Dirichlet" && t != "weak Dirichlet" && t != "Neumann") fail(MRHYDE_B200_ERR_INVALID, "set_bc: unknown type '" + t + "'");
  bool okv = false, oks = false;
  for (auto& v : P->var_names) okv = okv || v == var;
  for (auto& s : P->side_names) oks = oks || s == side;
  if (!okv) fail(MRHYDE_B200_ERR_INVALID, std::string("set_bc: unknown variable '") + var + "'");
  if (!oks) fail(MRHYDE_B200_ERR_INVALID, std::string("set_bc: unknown sideset '") + side + "' (call set_sidesets first)");
  BCEntry b; b.type = t; b.expr = expression ? expression : "0.0";
  P->bcs[{var, side}] = b;
  ABI_END
}

int mrhyde_b200_plan_add_boundary_group(mrhyde_b200_plan* P, const mrhyde_b200_boundary_group* bg) {
  ABI_BEGIN
  if (!P || !bg) fail(MRHYDE_B200_ERR_INVALID, "add_boundary_group: null argument");
  if (P->finalized) fail(MRHYDE_B200_ERR_STATE, "add_boundary_group after finalize");
  if (bg->sideset < 0 || bg->sideset >= (int)P->side_names.size()) fail(MRHYDE_B200_ERR_INVALID, "add_boundary_group: sideset out of range");
  if (bg->n_elem < 0 || (bg->n_elem > 0 && !bg->elem_ids) || bg->nqp_side < 1 || !bg->side_pts || !bg->side_wts || !bg->tangent_u || !bg->side_bases)
    fail(MRHYDE_B200_ERR_INVALID, "add_boundary_group: incomplete group");
  BoundaryGroupHost H;
  H.sideset = bg->sideset; H.local_side = bg->local_side; H.nqp = bg->nqp_side;
  H.elem_ids.assign(bg->elem_ids, bg->elem_ids + bg->n_elem);
  H.pts.assign(bg->side_pts, bg->side_pts + (size_t)bg->nqp_side * P->dim);
  H.wts.assign(bg->side_wts, bg->side_wts + bg->nqp_side);
  for (int d = 0; d < 3; ++d) { H.tu[d] = bg->tangent_u[d]; H.tv[d] = bg->tangent_v ? bg->tangent_v[d] : 0.0; }
  const mrhyde_b200_basis& sb = bg->side_bases[P->var_basis[0]];
  if (!sb.val) fail(MRHYDE_B200_ERR_INVALID, "add_boundary_group: missing side basis values");
  H.val.assign(sb.val, sb.val + (size_t)sb.card * bg->nqp_side);
  if (sb.grad) H.grad.assign(sb.grad, sb.grad + (size_t)sb.card * bg->nqp_side * P->dim);
  for (size_t b = 0; b < P->bases.size(); ++b) {   // every basis at the side points (general path)
    const mrhyde_b200_basis& in = bg->side_bases[b];
    const BasisCopy& B = P->bases[b];
    if (!in.val || in.card != B.card) fail(MRHYDE_B200_ERR_INVALID, "add_boundary_group: side basis table missing or of the wrong cardinality");
    const int vdim = (B.type == "HCURL" || B.type == "HDIV") ? P->dim : 1;
    H.bval.emplace_back(in.val, in.val + (size_t)in.card * bg->nqp_side * vdim);
    if (in.grad) H.bgrad.emplace_back(in.grad, in.grad + (size_t)in.card * bg->nqp_side * P->dim); else H.bgrad.emplace_back();
  }
  P->bgroups.push_back(std::move(H));
  ABI_END
}

int mrhyde_b200_plan_finalize(mrhyde_b200_plan* P) {
  ABI_BEGIN
  if (!P) fail(MRHYDE_B200_ERR_INVALID, "finalize: null plan");
  if (P->finalized) fail(MRHYDE_B200_ERR_STATE, "finalize called twice");
  if (!P->have_mesh || !P->have_graph) fail(MRHYDE_B200_ERR_STATE, "finalize: set_mesh and set_graph must be called first");
  const bool host_only = (P->device == -1);
  if (!host_only) CUDA_OK(cudaSetDevice(P->device));
  MeshGraph& M = P->mesh;
  for (int32_t l : M.lids) if (l < 0 || l >= M.nrows) fail(MRHYDE_B200_ERR_INVALID, "finalize: LID outside the graph's rows");

  const std::string phys = canonical_physics(P->physics);
  if (!module_functions(phys)) fail(MRHYDE_B200_ERR_UNSUPPORTED, "physics module '" + P->physics + "' has no device kernel in this build");
  // Kernel choice: the sweep kernel (volume_kernel.cuh) covers thermal on HGRAD order 1 with the 2-point Gauss rule and no
  // advection; everything else -- and option kernel=general -- takes the general element kernel + pull (general.hpp).
  const int NV = 1 << P->dim;
  const BasisCopy& B = P->bases[P->var_basis[0]];
  {
    const std::string want = opt(P, "kernel", "auto");
    if (want != "auto" && want != "general" && want != "sweep") fail(MRHYDE_B200_ERR_INVALID, "option kernel must be auto|general|sweep");
    bool sweep_ok = phys == "thermal" && P->nvars == 1 && B.type == "HGRAD" && B.order == 1 && B.card == NV && P->ndof_elem == NV && P->nqp == NV && !B.grad.empty() &&
                    !opt_bool(P, "include advection", false);
    for (int i = 0; sweep_ok && i < NV; ++i) if (P->offsets[i] != i) sweep_ok = false;
    if (opt_bool(P, "lump mass", false)) sweep_ok = false;   // the column redirect of the lumped scatter lives in the general path's pull
    if (sweep_ok) {
      // a coefficient that reads the solution (e.g. thermal diffusion: 1.0+T*T) needs derivative lanes through the function
      // evaluation (functionManager_evaluate.hpp:59-229): the sweep kernel's collapsed Jacobian does not apply -> general path
      FunctionSet probe = make_function_set(P, false, true);
      for (const char* nm : {"thermal source", "thermal diffusion", "specific heat", "density"})
        { const LongProgram lp = probe.compile_long(nm); if (lp.uses_state || lp.uses_reduction) sweep_ok = false; }   // element reductions (emax / emin / emean) too
    }
    if (want == "sweep" && !sweep_ok) fail(MRHYDE_B200_ERR_UNSUPPORTED, "kernel=sweep: the sweep kernel covers thermal, HGRAD order 1, 2-point Gauss rule, no advection only");
    if (want == "general" || !sweep_ok) { finalize_general(P, phys); return MRHYDE_B200_OK; }
  }

  FunctionSet fs = make_function_set(P, false);
  ExprProgram src = fs.compile("thermal source"), dif = fs.compile("thermal diffusion"), cp = fs.compile("specific heat"), rho = fs.compile("density");

  M.classify_cells();
  P->n_affine = 0; P->n_box = 0;
  for (uint8_t a : M.eclass) { P->n_affine += (a != 0); P->n_box += (a == 2); }

  // reference tables first: the ring layout depends on them
  auto fill_host = [&](auto& th) {
    th.source = src; th.diffusion = dif; th.specific_heat = cp; th.density = rho;
    th.all_const = (dif.is_const && cp.is_const && rho.is_const) ? 1 : 0;
  };
  if (P->dim == 3) { fill_thermal_tables<3>(P, P->th3.tab); fill_host(P->th3); }
  else { fill_thermal_tables<2>(P, P->th2.tab); fill_host(P->th2); }
  // Ring layout (option ring = auto | class | metric | full; volume_kernel.cuh).  The compressed layouts exist in the
  // plan-specialised build only, so they need NVRTC:
  //   class   axis-aligned boxes throughout + constant diffusion / specific heat / density: the local matrix takes NC distinct
  //           values (classes of upper-triangle entries with identical table columns), NC + NV doubles per element
  //   metric  parallelepipeds throughout + constant coefficients: scaled metric + load vector + state, combined in the pull
  //   full    upper triangle of the local matrix + residual (any cell, any coefficient; the ahead-of-time kernel's layout)
  const int NT = NV * (NV + 1) / 2;
  const std::string ring = opt(P, "ring", "auto");
  if (ring != "auto" && ring != "class" && ring != "metric" && ring != "full") fail(MRHYDE_B200_ERR_INVALID, "option ring must be auto|class|metric|full");
  std::string nvrtc_why;
  const bool all_const = dif.is_const && cp.is_const && rho.is_const;
  const bool jit_possible = opt(P, "jit", "auto") != "false" && nvrtc_available(nvrtc_why);
  const bool class_ok = jit_possible && all_const && P->n_box == M.nelem;
  const bool metric_ok = jit_possible && all_const && P->n_affine == M.nelem;
  if (ring == "class" && !class_ok) fail(MRHYDE_B200_ERR_UNSUPPORTED, "ring=class needs axis-aligned box cells throughout, constant diffusion / specific heat / density and the plan-specialised build");
  if (ring == "metric" && !metric_ok) fail(MRHYDE_B200_ERR_UNSUPPORTED, "ring=metric needs parallelepiped cells throughout, constant diffusion / specific heat / density and the plan-specialised build");
  const bool use_class = class_ok && (ring == "auto" || ring == "class");
  const bool use_metric = !use_class && metric_ok && (ring == "auto" || ring == "metric");
  P->class_nc = 0; P->class_of_t.clear(); P->class_rep.clear();
  if (use_class) {
    const double* St = P->dim == 3 ? &P->th3.tab.Stab[0][0] : &P->th2.tab.Stab[0][0];
    const double* Mt = P->dim == 3 ? &P->th3.tab.Mtab[0] : &P->th2.tab.Mtab[0];
    P->class_of_t.assign((size_t)NT, -1);
    for (int t = 0; t < NT; ++t) {
      for (size_t c = 0; c < P->class_rep.size() && P->class_of_t[(size_t)t] < 0; ++c) {
        const int r = P->class_rep[c];
        bool same = Mt[t] == Mt[r];
        for (int g = 0; g < P->dim && same; ++g) same = St[(size_t)g * NT + t] == St[(size_t)g * NT + r];   // tables are snapped: exact compare
        if (same) P->class_of_t[(size_t)t] = (int32_t)c;
      }
      if (P->class_of_t[(size_t)t] < 0) { P->class_of_t[(size_t)t] = (int32_t)P->class_rep.size(); P->class_rep.push_back(t); }
    }
    P->class_nc = (int)P->class_rep.size();
  }
  // staged vector per element: the local Jacobian (upper triangle, or one value per class), then the residual
  const int NK = use_class ? P->class_nc : NT, STAGE = NK + NV;
  std::vector<uint16_t> kmap((size_t)NV * NV), rmap((size_t)NV);
  for (int i = 0; i < NV; ++i) {
    rmap[(size_t)i] = (uint16_t)(NK + i);
    for (int j = 0; j < NV; ++j) {
      const int a = std::min(i, j), b = std::max(i, j), t = a * NV - (a * (a - 1)) / 2 + (b - a);
      kmap[(size_t)i * NV + j] = (uint16_t)(use_class ? P->class_of_t[(size_t)t] : t);
    }
  }
  // how the generated pull code writes finished rows: bulk (cp.async.bulk per run of CSR-contiguous rows, default) | row | flat
  const std::string flush_opt = opt(P, "flush", "bulk");
  if (flush_opt != "bulk" && flush_opt != "row" && flush_opt != "flat") fail(MRHYDE_B200_ERR_INVALID, "option flush must be bulk|row|flat");
  const int flush_mode = flush_opt == "bulk" ? 2 : (flush_opt == "row" ? 1 : 0);
  // one scalar field numbered like the vertices (the reference's Q1 thermal blocks): dof ids == vertex ids, one connectivity stream
  const bool lids_are_conn = M.lids.size() == M.conn.size() && M.lids == M.conn;
  ChainOptions co;
  co.column_elems = std::stoi(opt(P, "column elements", "128"));
  co.min_chains = std::stoi(opt(P, "min chains", "592"));
  co.sweep_axis = std::stoi(opt(P, "sweep axis", "-1"));
  co.cta_slots = std::max(0, std::stoi(opt(P, "cta slots", "0")));   // 0: from the device's SM count and the CTAs per SM the build allows
  co.max_blocks_per_sm = std::max(1, std::stoi(opt(P, "max blocks", "3")));
  if (!host_only) {
    int n_sm = 0;
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, P->device) == cudaSuccess && n_sm > 0) co.n_sm = n_sm;
  }
  {
    // shared memory of the build that will run: ring entries per element and the per-warp row buffer (longest free row)
    int64_t max_len = 0;
    for (int64_t r = 0; r < M.nrows; ++r) if (!M.fixed[(size_t)r]) max_len = std::max(max_len, M.rowptr[(size_t)r + 1] - M.rowptr[(size_t)r]);
    co.ring_stage_len = use_metric ? (P->dim + 2 * NV) : STAGE;
    co.warp_buffer_bytes = jit_possible ? std::max<size_t>(PULL_WARP_DOUBLES * 8, (size_t)(32 + 32 * (flush_mode == 2 ? std::min<int64_t>(max_len, 64) + 1 : (std::min<int64_t>(max_len, 64) | 1))) * 8) : (size_t)PULL_WARP_DOUBLES * 8;
  }
  co.min_segment_levels = std::max(1, std::stoi(opt(P, "min segment levels", "8")));
  if (co.column_elems < 1 || co.min_chains < 1) fail(MRHYDE_B200_ERR_INVALID, "options 'column elements' and 'min chains' must be positive");
  build_chain_plan(M, kmap, rmap, STAGE, co, P->cp);
  if (M.nowned > 0 && M.nowned < M.nrows && jit_possible && flush_mode == 2 && opt_bool(P, "halo push", true)) {
    // patterns of the batches that hold ghost rows: at most 8 get generated pull code on top of the most frequent ones
    std::set<int32_t> gp;
    for (size_t b = 0; b < P->cp.batches.size(); ++b) {
      const BatchRec& B = P->cp.batches[b];
      if (B.flags & BATCH_FIXED) continue;
      for (int l = 0; l < (int)B.n_rows; ++l) if (P->cp.rows[b * 32 + (size_t)l].row >= M.nowned) { gp.insert(B.desc_begin); break; }
    }
    if (gp.size() <= 8) P->cp.ghost_patterns.assign(gp.begin(), gp.end());
  }
  {
    const int want = std::stoi(opt(P, "threads", "0"));
    int th = want > 0 ? want : std::max(128, std::min(256, ((P->cp.cap + 31) / 32) * 32));
    if (th < 32 || th > 256 || th % 32) fail(MRHYDE_B200_ERR_INVALID, "option threads must be a multiple of 32 in [32,256]");
    if (th < P->cp.cap) th = ((P->cp.cap + 31) / 32) * 32;   // one thread per element of a sweep step
    if (th > 256) fail(MRHYDE_B200_ERR_UNSUPPORTED, "sweep steps larger than 256 elements are not supported (lower 'column elements')");
    P->threads = th;
  }
  // ring (2 slots) + one transpose buffer per warp
  // per-warp transpose buffer; the specialised builds park whole rows there (generated pull code, row_buffer_pitch)
  const int max_patterns = std::max(0, std::stoi(opt(P, "pull patterns", "3")));
  const size_t warp_doubles = std::max<size_t>(PULL_WARP_DOUBLES, jit_possible ? 32 + 32 * (size_t)row_buffer_pitch(P->cp, max_patterns, flush_mode) : 0);
  P->smem = (size_t)(2 * P->cp.slot_bytes()) + 8 /* the row buffers start 16-byte aligned */ + (size_t)(P->threads / 32) * warp_doubles * sizeof(double)
            + 64 /* CTA-shared sub-expression values of the source (up to 8 doubles, behind the row buffers) */;

  P->stage_len = STAGE;
  P->kmap = kmap; P->rmap = rmap;
  {
    // In-kernel halo push: possible when every batch that holds ghost rows is written by the generated bulk-flush code (one of
    // the specialised patterns) or is a batch of strong-Dirichlet rows (those stay zero in the owner's slab), and the chains that
    // complete ghost rows are the first n_early_chains of the plan.  The transport is checked per call (HaloExchange::push_params).
    P->push_ok = false;
    if (jit_possible && flush_mode == 2 && !use_metric && M.nowned > 0 && M.nowned < M.nrows && P->cp.n_early_chains > 0 && opt_bool(P, "halo push", true)) {
      std::set<int32_t> special;
      for (const SpecialPattern& sp : special_patterns(P->cp, max_patterns)) special.insert(sp.desc_begin);
      bool ok = true;
      for (size_t b = 0; b < P->cp.batches.size() && ok; ++b) {
        const BatchRec& B = P->cp.batches[b];
        bool ghost = false;
        for (int l = 0; l < (int)B.n_rows; ++l) if (P->cp.rows[b * 32 + (size_t)l].row >= M.nowned) ghost = true;
        if (ghost && !(B.flags & BATCH_FIXED) && !special.count(B.desc_begin)) ok = false;
      }
      P->push_ok = ok;
    }
  }
  {
    const int64_t n_class[3] = {M.nelem - P->n_affine, P->n_affine - P->n_box, P->n_box};
    P->metric_ng = (use_metric && !P->cp.mdesc[0].empty()) ? (n_class[1] == 0 ? P->dim : P->dim * (P->dim + 1) / 2) : 0;
    if (ring == "metric" && P->metric_ng == 0) fail(MRHYDE_B200_ERR_UNSUPPORTED, "ring=metric: the plan has no metric source words (sweep steps larger than 256 elements)");
    for (int tr = 0; tr < 2; ++tr)
      P->smem_metric[tr] = (size_t)(2 * P->cp.cap) * (size_t)(P->metric_ng + tr + NV * (2 + tr)) * sizeof(double) + (size_t)(P->threads / 32) * warp_doubles * sizeof(double);
    const int pull_group = std::stoi(opt(P, "pull group", "8"));
    P->jit_source = P->dim == 3 ? thermal_jit_source<3>(P->th3.tab, fs, P->th3.all_const, src.is_const, P->cp, n_class, P->metric_ng, max_patterns, pull_group, P->class_of_t, P->class_rep, std::stoi(opt(P, "debug skip", "0")), opt(P, "stage1", "late") != "early", flush_mode, std::max(1, std::stoi(opt(P, "flush unroll", "8"))), opt(P, "stage2", "late") == "early", opt(P, "tables", "constant") == "literal", lids_are_conn, std::max(0, std::min(2, std::stoi(opt(P, "pipeline", "0")))), opt_bool(P, "store hint", false) ? 1 : 0, prefetch_mode(P), P->push_ok, opt_bool(P, "prefetch records", true), column_cache_axes(P))
                                : thermal_jit_source<2>(P->th2.tab, fs, P->th2.all_const, src.is_const, P->cp, n_class, P->metric_ng, max_patterns, pull_group, P->class_of_t, P->class_rep, std::stoi(opt(P, "debug skip", "0")), opt(P, "stage1", "late") != "early", flush_mode, std::max(1, std::stoi(opt(P, "flush unroll", "8"))), opt(P, "stage2", "late") == "early", opt(P, "tables", "constant") == "literal", lids_are_conn, std::max(0, std::min(2, std::stoi(opt(P, "pipeline", "0")))), opt_bool(P, "store hint", false) ? 1 : 0, prefetch_mode(P), P->push_ok, opt_bool(P, "prefetch records", true), column_cache_axes(P));
  }
  if (host_only) {
    // boundary groups still get their expressions compiled so that set-up errors surface
    if (!P->bgroups.empty()) {
      FunctionSet fss = make_function_set(P, true);
      for (auto& g : P->bgroups) {
        const std::string& sname = P->side_names[(size_t)g.sideset];
        auto it = P->bcs.find({P->var_names[0], sname});
        g.bctype = (it == P->bcs.end()) ? "none" : it->second.type;
        if (g.bctype == "weak Dirichlet") g.data = fss.compile("Dirichlet " + P->var_names[0] + " " + sname);
        else if (g.bctype == "Neumann") g.data = fss.compile("Neumann " + P->var_names[0] + " " + sname);
      }
    }
    P->launches_per_assemble = 1;
    P->finalized = true;
    return MRHYDE_B200_OK;
  }
  // ---- upload
  size_t* tot = &P->dev_bytes;
  const ChainPlan& CP = P->cp;
  P->d_vx.upload(M.vcoord[0], tot); P->d_vy.upload(M.vcoord[1], tot); P->d_vz.upload(M.vcoord[2], tot);
  P->d_conn.upload(M.conn, tot); P->d_lids.upload(M.lids, tot);
  P->d_rowptr.upload(M.rowptr, tot); P->d_colind.upload(M.colind, tot); P->d_fixed.upload(M.fixed, tot); P->d_eclass.upload(M.eclass, tot);
  P->d_chain_step_ptr.upload(CP.chain_step_ptr, tot); P->d_steps.upload(CP.steps, tot); P->d_step_elems.upload(CP.step_elems, tot);
  P->d_batches.upload(CP.batches, tot); P->d_rows.upload(CP.rows, tot);
  P->d_step_conn.upload(CP.step_conn, tot); P->d_step_eclass.upload(CP.step_eclass, tot);
  P->d_chain_invariant.upload(CP.chain_invariant, tot);
  if (!lids_are_conn) P->d_step_lids.upload(CP.step_lids, tot);   // else the kernels read the one connectivity stream for both
  P->d_desc0.upload(CP.desc[0], tot); P->d_desc1.upload(CP.desc[1], tot);
  if (P->metric_ng > 0) { P->d_mdesc0.upload(CP.mdesc[0], tot); P->d_mdesc1.upload(CP.mdesc[1], tot); }
  P->d_orphans.upload(CP.orphan_rows, tot);
  {
    std::vector<int64_t> diag;
    for (int64_t r = 0; r < M.nrows; ++r) {
      if (!M.fixed[(size_t)r] || r >= M.nowned) continue;   // ghost copies of fixed rows stay zero (see plan.cpp)
      int64_t pos = -1;
      for (int64_t p = M.rowptr[(size_t)r]; p < M.rowptr[(size_t)r + 1]; ++p) if (M.colind[(size_t)p] == r) pos = p;
      diag.push_back(pos);
    }
    P->d_fixed_diag.upload(diag, tot);
    if (diag.empty()) P->d_fixed_diag.n = 0;
  }
  ChainDev D{};
  D.chain_step_ptr = P->d_chain_step_ptr.p; D.steps = P->d_steps.p; D.step_elems = P->d_step_elems.p;
  D.batches = P->d_batches.p; D.rows = P->d_rows.p;
  D.step_conn = P->d_step_conn.p; D.step_lids = lids_are_conn ? P->d_step_conn.p : P->d_step_lids.p; D.step_eclass = P->d_step_eclass.p;
  D.desc0 = reinterpret_cast<const SrcQuad*>(P->d_desc0.p); D.desc1 = reinterpret_cast<const SrcQuad*>(P->d_desc1.p);
  D.mdesc0 = reinterpret_cast<const SrcQuad*>(P->d_mdesc0.p); D.mdesc1 = reinterpret_cast<const SrcQuad*>(P->d_mdesc1.p);
  D.cap = CP.cap;
  D.chain_invariant = CP.chain_invariant.empty() ? nullptr : P->d_chain_invariant.p;
  GraphDev G{P->d_rowptr.p, P->d_colind.p, P->d_fixed.p};
  auto fill_common = [&](auto& th) {
    th.vx = P->d_vx.p; th.vy = P->d_vy.p; th.vz = P->d_vz.p;
    th.conn = P->d_conn.p; th.lids = P->d_lids.p; th.eclass = P->d_eclass.p;
    th.chains = D; th.graph = G;
  };
  if (P->dim == 3) fill_common(P->th3); else fill_common(P->th2);
  P->launches_per_assemble = 1;

  P->jit_min_blocks = std::max(1, std::min(std::min(8, 2048 / P->threads), (int)((228 * 1024) / (P->smem + 1024))));
  // ---- plan-specialised kernel (NVRTC): expressions, tables and block size become compile-time constants
  {
    const std::string want = opt(P, "jit", "auto");
    if (want != "auto" && want != "true" && want != "false") fail(MRHYDE_B200_ERR_INVALID, "option jit must be auto|true|false");
    P->use_jit = false;
    if (want == "false") P->jit_note = "disabled by option";
    else {
      std::string log;
      // build the variant the first assemble call will most likely use (steady, residual + Jacobian, current accumulate mode)
      if (jit_variant(P, false, 3 | (P->accumulate ? 4 : 0), log)) P->use_jit = true;
      else if (want == "true" || P->class_nc > 0 || opt(P, "ring", "auto") == "metric")   // a class-ring plan has no ahead-of-time kernel
        fail(MRHYDE_B200_ERR_CUDA, "the plan could not be specialised (option ring=full selects the layout of the ahead-of-time kernel): " + log);
      else { P->jit_note = log; P->metric_ng = 0; }   // the ahead-of-time kernel keeps full local systems in the ring
    }
  }

  // ---- boundary groups (Neumann / weak Dirichlet); strong-Dirichlet and "none" sides add nothing
  if (!P->bgroups.empty()) {
    FunctionSet fss = make_function_set(P, true);
    BoundarySetup bs;
    bs.dim = P->dim; bs.nv = NV;
    bs.formparam = std::stod(opt(P, "form_param", "1.0"));
    for (auto& g : P->bgroups) {
      const std::string& sname = P->side_names[(size_t)g.sideset];
      auto it = P->bcs.find({P->var_names[0], sname});
      g.bctype = (it == P->bcs.end()) ? "none" : it->second.type;
      if (g.bctype == "weak Dirichlet") g.data = fss.compile("Dirichlet " + P->var_names[0] + " " + sname);
      else if (g.bctype == "Neumann") g.data = fss.compile("Neumann " + P->var_names[0] + " " + sname);
    }
    bs.diffusion = fss.compile("thermal diffusion");
    build_boundary_plan(bs, P->bgroups, M, P->boundary, tot);
    P->boundary.vx = P->d_vx.p; P->boundary.vy = P->d_vy.p; P->boundary.vz = P->d_vz.p;
    P->boundary.conn = P->d_conn.p; P->boundary.lids = P->d_lids.p;
    P->launches_per_assemble += (int)P->boundary.groups.size();
  }
  if (P->d_fixed_diag.n > 0) P->launches_per_assemble += 1;
  P->finalized = true;
  ABI_END
}

int mrhyde_b200_assemble_jacres(mrhyde_b200_plan* P, const double* sol, const mrhyde_b200_time* t, int compute_jacobian, int compute_residual,
                                double* res, double* jac_values, void* stream) {
  ABI_BEGIN
  if (!P) fail(MRHYDE_B200_ERR_INVALID, "assemble_jacres: null plan");
  if (!compute_jacobian && !compute_residual) fail(MRHYDE_B200_ERR_INVALID, "assemble_jacres: nothing requested");
  TimeDev td;
  fill_time(t, td, true);
  do_assemble(P, sol, td, compute_jacobian != 0, compute_residual != 0, res, jac_values, (cudaStream_t)stream);
  ABI_END
}

int mrhyde_b200_assemble_jacres_adjoint(mrhyde_b200_plan* P, const double* sol, const mrhyde_b200_time* t, double* res, double* jac_values, void* stream) {
  ABI_BEGIN
  if (!P) fail(MRHYDE_B200_ERR_INVALID, "assemble_jacres_adjoint: null plan");
  TimeDev td;
  fill_time(t, td, true);
  // the sweep kernel's local matrices are symmetric (thermal without advection and without state-dependent coefficients): the
  // transposed fill changes nothing there; its boundary kernel and the general path transpose explicitly
  do_assemble(P, sol, td, true, true, res, jac_values, (cudaStream_t)stream, true);
  ABI_END
}

int mrhyde_b200_plan_warmup(mrhyde_b200_plan* P, int transient, int compute_jacobian, int compute_residual) {
  ABI_BEGIN
  if (!P) fail(MRHYDE_B200_ERR_INVALID, "plan_warmup: null plan");
  if (!P->finalized) fail(MRHYDE_B200_ERR_STATE, "plan_warmup called before mrhyde_b200_plan_finalize");
  if (P->device == -1) fail(MRHYDE_B200_ERR_STATE, "host-only analysis plan (device = -1) has no kernels to build");
  if (!compute_jacobian && !compute_residual) fail(MRHYDE_B200_ERR_INVALID, "plan_warmup: nothing requested");
  if (P->use_jit && !P->use_general) {
    CUDA_OK(cudaSetDevice(P->device));
    const int mode = (compute_residual ? 1 : 0) | (compute_jacobian ? 2 : 0) | (P->accumulate ? 4 : 0);
    std::string log;
    if (!jit_variant(P, transient != 0, mode, log)) fail(MRHYDE_B200_ERR_CUDA, "jit: kernel build failed: " + log);
  }
  ABI_END
}

int mrhyde_b200_assemble_res(mrhyde_b200_plan* P, const double* sol, const mrhyde_b200_time* t, double* res, void* stream) {
  ABI_BEGIN
  if (!P) fail(MRHYDE_B200_ERR_INVALID, "assemble_res: null plan");
  TimeDev td;
  fill_time(t, td, true);
  do_assemble(P, sol, td, false, true, res, nullptr, (cudaStream_t)stream);
  ABI_END
}

int mrhyde_b200_assemble_jacres_host(mrhyde_b200_plan* P, const double* sol, const mrhyde_b200_time* t, int compute_jacobian, int compute_residual,
                                     double* res, double* jac_values) {
  ABI_BEGIN
  if (!P) fail(MRHYDE_B200_ERR_INVALID, "assemble_jacres_host: null plan");
  if (!P->finalized) fail(MRHYDE_B200_ERR_STATE, "assemble called before mrhyde_b200_plan_finalize");
  if (P->device == -1) fail(MRHYDE_B200_ERR_STATE, "host-only analysis plan (device = -1) cannot assemble: there is no CPU path");
  if (!sol || (compute_residual && !res) || (compute_jacobian && !jac_values)) fail(MRHYDE_B200_ERR_INVALID, "assemble_jacres_host: null buffer");
  CUDA_OK(cudaSetDevice(P->device));
  const size_t nr = (size_t)P->mesh.nrows, nnz = (size_t)P->mesh.nnz;
  cudaStream_t st = 0;
  P->h_sol.resize(nr, nullptr);
  CUDA_OK(cudaMemcpyAsync(P->h_sol.p, sol, nr * sizeof(double), cudaMemcpyHostToDevice, st));
  TimeDev td;
  fill_time(t, td, false);
  if (td.transient) {
    if ((int)P->h_prev.size() < td.nprev) P->h_prev.resize((size_t)td.nprev);
    if ((int)P->h_stage.size() < td.nstage_lo) P->h_stage.resize((size_t)td.nstage_lo);
    for (int k = 0; k < td.nprev; ++k) {
      P->h_prev[(size_t)k].resize(nr, nullptr);
      CUDA_OK(cudaMemcpyAsync(P->h_prev[(size_t)k].p, t->sol_prev[k], nr * sizeof(double), cudaMemcpyHostToDevice, st));
      td.prev[k] = P->h_prev[(size_t)k].p;
    }
    for (int k = 0; k < td.nstage_lo; ++k) {
      P->h_stage[(size_t)k].resize(nr, nullptr);
      CUDA_OK(cudaMemcpyAsync(P->h_stage[(size_t)k].p, t->sol_stage[k], nr * sizeof(double), cudaMemcpyHostToDevice, st));
      td.stg[k] = P->h_stage[(size_t)k].p;
    }
  }
  if (compute_residual) {
    P->h_res.resize(nr, nullptr);
    if (P->accumulate) CUDA_OK(cudaMemcpyAsync(P->h_res.p, res, nr * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  if (compute_jacobian) {
    P->h_jac.resize(nnz, nullptr);
    if (P->accumulate) CUDA_OK(cudaMemcpyAsync(P->h_jac.p, jac_values, nnz * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  P->suppress_overlap = true;   // host buffers: there is no halo_sum on the library's scratch arrays
  try { do_assemble(P, P->h_sol.p, td, compute_jacobian != 0, compute_residual != 0, P->h_res.p, P->h_jac.p, st); }
  catch (...) { P->suppress_overlap = false; throw; }
  P->suppress_overlap = false;
  if (compute_residual) CUDA_OK(cudaMemcpyAsync(res, P->h_res.p, nr * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (compute_jacobian) CUDA_OK(cudaMemcpyAsync(jac_values, P->h_jac.p, nnz * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  ABI_END
}

int mrhyde_b200_comm_unique_id(uint8_t* id128) {
  ABI_BEGIN
  if (!id128) fail(MRHYDE_B200_ERR_INVALID, "comm_unique_id: null argument");
  std::string err;
  if (!halo_unique_id(id128, err)) fail(MRHYDE_B200_ERR_NCCL, err);
  ABI_END
}

int mrhyde_b200_plan_comm_init(mrhyde_b200_plan* P, const uint8_t* id128, int rank, int nranks) {
  ABI_BEGIN
  if (!P || !id128) fail(MRHYDE_B200_ERR_INVALID, "comm_init: null argument");
  CUDA_OK(cudaSetDevice(P->device));
  std::string err;
  P->halo.reset(new HaloExchange());
  if (!P->halo->init(id128, rank, nranks, err)) { P->halo.reset(); fail(MRHYDE_B200_ERR_NCCL, err); }
  ABI_END
}

int mrhyde_b200_plan_set_halo(mrhyde_b200_plan* P, int64_t n_cols, const int64_t* col_gids) {
  ABI_BEGIN
  if (!P || !col_gids) fail(MRHYDE_B200_ERR_INVALID, "set_halo: null argument");
  if (!P->halo) fail(MRHYDE_B200_ERR_STATE, "set_halo: call plan_comm_init first");
  if (!P->have_graph) fail(MRHYDE_B200_ERR_STATE, "set_halo: call set_graph first");
  if (n_cols < P->mesh.nrows) fail(MRHYDE_B200_ERR_INVALID, "set_halo: n_cols must be >= n_rows");
  for (int32_t c : P->mesh.colind) if (c < 0 || c >= n_cols) fail(MRHYDE_B200_ERR_INVALID, "set_halo: a column index of the graph is outside [0, n_cols)");
  CUDA_OK(cudaSetDevice(P->device));
  std::string err;
  // "overlap halo" starts the exchange of some ranks with NCCL send/recv from inside assemble: every rank must then use that transport
  P->halo->set_transport(opt_bool(P, "overlap halo", false) ? std::string("nccl") : opt(P, "halo transport", "auto"));
  if (!P->halo->setup(P->mesh.nrows, P->mesh.nowned, n_cols, col_gids, P->mesh.rowptr.data(), P->mesh.colind.data(), err)) fail(MRHYDE_B200_ERR_NCCL, err);
  ABI_END
}

int mrhyde_b200_halo_sum(mrhyde_b200_plan* P, double* res, double* jac_values, void* stream) {
  ABI_BEGIN
  if (!P) fail(MRHYDE_B200_ERR_INVALID, "halo_sum: null plan");
  if (!P->halo || !P->halo->ready()) fail(MRHYDE_B200_ERR_STATE, "halo_sum: call plan_comm_init and plan_set_halo first");
  CUDA_OK(cudaSetDevice(P->device));
  NvtxRange exp("MrHyDE::LinearAlgebraInterface::export*()");
  std::string err;
  if (!P->halo->sum(res, jac_values, (cudaStream_t)stream, err)) fail(MRHYDE_B200_ERR_NCCL, err);
  ABI_END
}

int mrhyde_b200_plan_set_point_dofs(mrhyde_b200_plan* P, int64_t n, const int32_t* lids) {
  ABI_BEGIN
  if (!P || n < 0 || (n > 0 && !lids)) fail(MRHYDE_B200_ERR_INVALID, "plan_set_point_dofs: null argument");
  if (!P->finalized) fail(MRHYDE_B200_ERR_STATE, "plan_set_point_dofs called before mrhyde_b200_plan_finalize");
  const MeshGraph& M = P->mesh;
  std::vector<int64_t> rows;
  bool ghost = false;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t r = lids[i];
    if (r < 0 || r >= M.nrows) fail(MRHYDE_B200_ERR_INVALID, "plan_set_point_dofs: row id out of range");
    int64_t d = -1;
    for (int64_t p = M.rowptr[(size_t)r]; p < M.rowptr[(size_t)r + 1]; ++p) if (M.colind[(size_t)p] == r) d = p;
    rows.push_back(M.rowptr[(size_t)r]); rows.push_back(M.rowptr[(size_t)r + 1]); rows.push_back(d);
    if (M.nowned > 0 && r >= M.nowned) ghost = true;
  }
  P->point_dofs.assign(lids, lids + n);
  P->point_on_ghost = ghost;
  if (P->device >= 0) {
    CUDA_OK(cudaSetDevice(P->device));
    CUDA_OK(cudaDeviceSynchronize());   // an assembly in flight may still read the old list
    size_t dummy = 0;
    P->d_point_rows.upload(rows, &dummy);
    if (rows.empty()) P->d_point_rows.n = 0;
  }
  ABI_END
}

int mrhyde_b200_plan_owned_extent(mrhyde_b200_plan* P, int64_t* n_owned_rows, int64_t* nnz_owned) {
  ABI_BEGIN
  if (!P || !n_owned_rows || !nnz_owned) fail(MRHYDE_B200_ERR_INVALID, "plan_owned_extent: null argument");
  if (!P->have_graph) fail(MRHYDE_B200_ERR_STATE, "plan_owned_extent: call set_graph first");
  *n_owned_rows = P->mesh.nowned > 0 ? P->mesh.nowned : P->mesh.nrows;
  *nnz_owned = P->mesh.rowptr[(size_t)*n_owned_rows];
  ABI_END
}

int mrhyde_b200_plan_stat(mrhyde_b200_plan* P, const char* key, int64_t* value) {
  ABI_BEGIN
  if (!P || !key || !value) fail(MRHYDE_B200_ERR_INVALID, "plan_stat: null argument");
  const std::string k(key);
  if (k == "n_chains") *value = P->cp.n_chains;
  else if (k == "n_columns") *value = P->cp.n_columns;
  else if (k == "n_segments") *value = P->cp.n_segments;
  else if (k == "n_levels") *value = P->cp.n_levels;
  else if (k == "n_steps") *value = (int64_t)P->cp.steps.size();
  else if (k == "n_patterns") *value = (int64_t)P->cp.patterns.size();
  else if (k == "n_pattern_slots") *value = (int64_t)(P->cp.desc[0].size() / SLOT_SRCS);
  else if (k == "n_batches") *value = (int64_t)P->cp.batches.size();
  else if (k == "max_batches_per_step") *value = P->cp.max_batches_step;
  else if (k == "ring_capacity") *value = P->cp.cap;
  else if (k == "max_rows_per_step") *value = P->cp.max_rows_step;
  else if (k == "kernel_launches_per_assemble") *value = P->launches_per_assemble;
  else if (k == "halo_launches_per_sum") *value = P->halo ? P->halo->launches_per_sum() : 0;
  else if (k == "pushed_assembles") *value = P->pushed_assembles;
  else if (k == "halo_p2p") *value = (P->halo && P->halo->p2p()) ? 1 : 0;
  else if (k == "smem_bytes") *value = (int64_t)variant_smem(P, false);
  else if (k == "metric_ring") *value = P->metric_ng;
  else if (k == "overlapped_assembles") *value = P->overlapped_assembles;
  else if (k == "n_early_chains") *value = P->cp.n_early_chains;
  else if (k == "column_cache_axes") *value = column_cache_axes(P) & 7;
  else if (k == "step_shared_axes") *value = (column_cache_axes(P) >> 4) & 7;   // axes whose sub-expressions one warp evaluates for the whole CTA   // axes whose one-coordinate source sub-expressions a thread keeps along its chain
  else if (k == "plan_hash") {   // FNV-1a over every array of the sweep plan: two builds that agree here launch the same schedule
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; } };
    const ChainPlan& c = P->cp;
    auto vec = [&](const auto& v) { if (!v.empty()) mix(v.data(), v.size() * sizeof(v[0])); };
    vec(c.chain_step_ptr); vec(c.steps); vec(c.step_elems); vec(c.step_conn); vec(c.step_lids); vec(c.step_eclass); vec(c.batches); vec(c.rows);
    vec(c.desc[0]); vec(c.desc[1]); vec(c.mdesc[0]); vec(c.mdesc[1]); vec(c.orphan_rows); vec(c.ghost_patterns); vec(c.chain_invariant);
    mix(&c.cap, sizeof(c.cap)); mix(&c.n_early_chains, sizeof(c.n_early_chains));
    *value = (int64_t)(h >> 1);
  }
  else if (k == "n_invariant_chains") { int64_t n = 0; for (uint8_t f : P->cp.chain_invariant) n += f != 0; *value = n; }
  else if (k == "class_ring") *value = P->class_nc;
  else if (k == "stage_len") *value = P->stage_len;
  else if (k == "threads_per_block") *value = P->threads;
  else if (k == "n_elem") *value = P->mesh.nelem;
  else if (k == "n_elem_with_halo") *value = P->cp.n_elem_with_halo;
  else if (k == "n_rows") *value = P->mesh.nrows;
  else if (k == "nnz") *value = P->mesh.nnz;
  else if (k == "n_verts") *value = P->mesh.nvert;
  else if (k == "plan_device_bytes") *value = (int64_t)P->dev_bytes;
  else if (k == "n_affine") *value = P->n_affine;
  else if (k == "n_box") *value = P->n_box;
  else if (k == "n_orphan_rows") *value = (int64_t)P->cp.orphan_rows.size();
  else if (k == "jit") *value = P->use_jit ? 1 : 0;
  else if (k == "jit_registers") { *value = 0; for (auto& kv : P->jit) *value = std::max<int64_t>(*value, kv.second->regs()); }
  else if (k == "jit_variants") *value = (int64_t)P->jit.size();
  else if (k == "general") *value = P->use_general ? 1 : 0;
  else if (k == "general_batches") *value = (int64_t)P->gen.batches.size();
  else if (k == "general_instances") *value = P->gen.n_inst;
  else if (k == "general_scratch_bytes") *value = gen_scratch_instances(P->gen) * (int64_t)P->gen.info.N * (P->gen.info.N + 1) * 8;
  else if (k == "general_batches") *value = (int64_t)P->gen.batches.size();
  else if (k == "general_pos_tables") *value = P->gen.info.N > 0 ? (int64_t)(P->gen.pos_tab.size() / ((size_t)P->gen.info.N * P->gen.info.N)) : 0;
  else fail(MRHYDE_B200_ERR_INVALID, "plan_stat: unknown key '" + k + "'");
  ABI_END
}

int mrhyde_b200_plan_kernel_time(mrhyde_b200_plan* P, int reset, double* avg_ms, int64_t* n_launches) {
  ABI_BEGIN
  if (!P) fail(MRHYDE_B200_ERR_INVALID, "kernel_time: null plan");
  while (P->ev_used > 0) fold_oldest(P);
  if (avg_ms) *avg_ms = P->ev_count ? P->ev_ms / (double)P->ev_count : 0.0;
  if (n_launches) *n_launches = P->ev_count;
  if (reset) { P->ev_ms = 0.0; P->ev_count = 0; }
  ABI_END
}

int mrhyde_b200_plan_eval_function(mrhyde_b200_plan* P, const char* name, int64_t npts, const double* xyz, double time, double* out) {
  ABI_BEGIN
  if (!P || !name || !xyz || !out || npts < 0) fail(MRHYDE_B200_ERR_INVALID, "eval_function: bad argument");
  if (P->device == -1) fail(MRHYDE_B200_ERR_STATE, "eval_function needs a device plan (use mrhyde_b200_expr_eval_host for host checks)");
  CUDA_OK(cudaSetDevice(P->device));
  FunctionSet fs = make_function_set(P, false);
  if (!fs.has(name)) fail(MRHYDE_B200_ERR_INVALID, std::string("eval_function: function not registered: ") + name);
  const ExprProgram prog = fs.compile(name);
  if (npts == 0) return MRHYDE_B200_OK;
  DevBuf<double> dx, dout;
  dx.resize((size_t)npts * 3, nullptr); dout.resize((size_t)npts, nullptr);
  CUDA_OK(cudaMemcpy(dx.p, xyz, (size_t)npts * 3 * sizeof(double), cudaMemcpyHostToDevice));
  eval_points_kernel<<<(unsigned)((npts + 127) / 128), 128>>>(prog, dx.p, time, npts, dout.p);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpy(out, dout.p, (size_t)npts * sizeof(double), cudaMemcpyDeviceToHost));
  ABI_END
}

static FunctionSet adhoc_set(int32_t n, const char* const* names, const char* const* exprs) {
  FunctionSet fs;
  for (int i = 0; i < n; ++i) {
    if (!names[i] || !exprs[i]) fail(MRHYDE_B200_ERR_INVALID, "null function name / expression");
    fs.set(names[i], exprs[i]);
  }
  fs.set_scalar_fields({"x", "y", "z", "", "n[x]", "n[y]", "n[z]"});
  return fs;
}

int mrhyde_b200_expr_disassemble(int32_t n, const char* const* names, const char* const* exprs, const char* which, char* out, size_t out_cap) {
  ABI_BEGIN
  if (n < 0 || (n > 0 && (!names || !exprs)) || !which || !out || out_cap == 0) fail(MRHYDE_B200_ERR_INVALID, "expr_disassemble: bad argument");
  FunctionSet fs = adhoc_set(n, names, exprs);
  const std::string text = FunctionSet::disassemble(fs.compile(which));
  if (text.size() + 1 > out_cap) fail(MRHYDE_B200_ERR_INVALID, "expr_disassemble: output buffer too small");
  std::memcpy(out, text.c_str(), text.size() + 1);
  ABI_END
}

int mrhyde_b200_expr_eval_host(int32_t n, const char* const* names, const char* const* exprs, const char* which, int64_t npts, const double* vars7, double* out) {
  ABI_BEGIN
  if (n < 0 || (n > 0 && (!names || !exprs)) || !which || npts < 0 || (npts > 0 && (!vars7 || !out))) fail(MRHYDE_B200_ERR_INVALID, "expr_eval_host: bad argument");
  FunctionSet fs = adhoc_set(n, names, exprs);
  const ExprProgram p = fs.compile(which);
  for (int64_t i = 0; i < npts; ++i) out[i] = p.is_const ? p.cval : FunctionSet::eval_host(p, vars7 + 7 * i);
  ABI_END
}

int mrhyde_b200_plan_debug_jit(mrhyde_b200_plan* P, const char* source_path, const char* cubin_path, char* log, size_t log_cap) {
  ABI_BEGIN
  if (!P) fail(MRHYDE_B200_ERR_INVALID, "debug_jit: null plan");
  if (!P->finalized) fail(MRHYDE_B200_ERR_STATE, "debug_jit before finalize");
  if (source_path) {
    FILE* f = std::fopen(source_path, "wb");
    if (!f) fail(MRHYDE_B200_ERR_INVALID, "debug_jit: cannot write the source file");
    std::fwrite(P->jit_source.data(), 1, P->jit_source.size(), f);
    std::fclose(f);
  }
  std::string cubin, text;
  // which build: options "debug transient" (0 | 1) and "debug mode" (1 res | 2 jac | 4 accumulate, summed); default steady, mode 7
  const bool transient = opt(P, "debug transient", "0") == "1";
  const int mode = std::stoi(opt(P, "debug mode", "7"));
  if (mode < 1 || mode > 7 || (mode & 3) == 0) fail(MRHYDE_B200_ERR_INVALID, "debug_jit: option 'debug mode' must have the residual and/or Jacobian bit set");
  const std::string variant = specialise_source(P, transient, mode, text);
  if (variant.empty()) fail(MRHYDE_B200_ERR_CUDA, "debug_jit: " + text);
  if (source_path) {   // the specialised text replaces the template written above
    FILE* f = std::fopen(source_path, "wb");
    if (f) { std::fwrite(variant.data(), 1, variant.size(), f); std::fclose(f); }
  }
  const int min_blocks = variant_min_blocks(P, variant_smem(P, transient));
  const bool ok = nvrtc_compile(variant, P->threads, min_blocks, cubin, text, std::stoi(opt(P, "max registers", "0")));
  if (log && log_cap) { std::snprintf(log, log_cap, "%s", text.c_str()); }
  if (!ok) fail(MRHYDE_B200_ERR_CUDA, "debug_jit: " + text);
  if (cubin_path) {
    FILE* f = std::fopen(cubin_path, "wb");
    if (!f) fail(MRHYDE_B200_ERR_INVALID, "debug_jit: cannot write the cubin file");
    std::fwrite(cubin.data(), 1, cubin.size(), f);
    std::fclose(f);
  }
  ABI_END
}

int mrhyde_b200_plan_debug_emulate(mrhyde_b200_plan* P, const double* sol, const mrhyde_b200_time* t, int compute_jacobian, int compute_residual,
                                   double* res, double* jac) {
  ABI_BEGIN
  if (!P || !sol) fail(MRHYDE_B200_ERR_INVALID, "debug_emulate: null argument");
  if (!P->finalized || !P->use_general) fail(MRHYDE_B200_ERR_STATE, "debug_emulate: needs a finalized plan on the general path (option kernel=general)");
  if (P->device != -1) fail(MRHYDE_B200_ERR_STATE, "debug_emulate: only host-only analysis plans (device = -1) replay the kernel stages on the host; device plans assemble on the GPU");
  TimeDev td;
  fill_time(t, td, true);   // host pointers: the replay reads them on the host
  const GeneralPlanHost& H = P->gen;
  const GenKernelInfo& I = H.info;
  const MeshGraph& M = P->mesh;
  const size_t N = (size_t)I.N;
  std::vector<double> ej(compute_jacobian ? (size_t)H.n_inst * N * N : 0, 0.0), er(compute_residual ? (size_t)H.n_inst * N : 0, 0.0);
  GenParams Q;
  std::memset(&Q, 0, sizeof(Q));
  Q.tensor = H.use_tensor ? 1 : 0;
  Q.vx = M.vcoord[0].data(); Q.vy = M.vcoord[1].data(); Q.vz = M.vcoord[2].data(); Q.conn = M.conn.data(); Q.lids = M.lids.data();
  Q.orient = M.orient.empty() ? nullptr : M.orient.data();
  Q.sol = sol; Q.td = td;
  std::memcpy(Q.off, H.off, sizeof(Q.off));
  Q.fn_op = H.fn_op.data(); Q.fn_c = H.fn_c.data(); Q.opt = H.opt;
  Q.elem_jac = compute_jacobian ? ej.data() : nullptr;
  Q.elem_res = compute_residual ? er.data() : nullptr;
  if (opt_bool(P, "assemble volume terms", true)) {
    Q.epb = 3;   // deliberately not a divisor of most meshes: exercises the padding elements
    Q.items = nullptr; Q.item_begin = 0; Q.item_end = M.nelem; Q.inst_base = 0;
    Q.geo_N = H.geo_N.data(); Q.geo_dN = H.geo_dN.data(); Q.ref_tab = H.ref_tab.data(); Q.qwts = H.qwts.data();
    std::memcpy(Q.fn, H.fn, sizeof(Q.fn));
    Q.fn_state = 0; for (int f = 0; f < GEN_MAXFN; ++f) if ((Q.fn[f].pad & 1) && !Q.fn[f].is_const) Q.fn_state = 1;
    for (int v = 0; v < GEN_MAXVARS; ++v) { Q.bc_type[v] = 0; Q.bc_fn[v] = -1; }
    need_emulator(P)->emulate(false, Q, (int)((M.nelem + Q.epb - 1) / Q.epb));
  }
  if (opt_bool(P, "assemble boundary terms", true))
    for (auto& S : H.sides) {
      if (!S.active || S.items.empty()) continue;
      Q.epb = 2;
      Q.items = S.items.data(); Q.item_begin = 0; Q.item_end = (int64_t)S.items.size(); Q.inst_base = S.inst_base;
      Q.geo_N = S.geo_N.data(); Q.geo_dN = S.geo_dN.data(); Q.ref_tab = S.ref_tab.data(); Q.qwts = S.qwts.data();
      for (int d = 0; d < 3; ++d) { Q.tan_u[d] = S.tan_u[d]; Q.tan_v[d] = S.tan_v[d]; }
      for (int v = 0; v < GEN_MAXVARS; ++v) { Q.bc_type[v] = S.bc_type[v]; Q.bc_fn[v] = S.bc_fn[v]; }
      std::memcpy(Q.fn, S.fn, sizeof(Q.fn));
      Q.fn_state = 0; for (int f = 0; f < GEN_MAXFN; ++f) if ((Q.fn[f].pad & 1) && !Q.fn[f].is_const) Q.fn_state = 1;
      need_emulator(P)->emulate(true, Q, (int)((Q.item_end + Q.epb - 1) / Q.epb));
    }
  gen_pull_host(H, M, compute_jacobian ? ej.data() : nullptr, compute_residual ? er.data() : nullptr, P->accumulate, compute_residual ? res : nullptr,
                compute_jacobian ? jac : nullptr);
  if (compute_jacobian && jac && P->accumulate && opt_bool(P, "use strong DBCs", true))
    for (int64_t r = 0; r < M.nowned; ++r)
      if (M.fixed[(size_t)r])
        for (int64_t p = M.rowptr[(size_t)r]; p < M.rowptr[(size_t)r + 1]; ++p) if (M.colind[(size_t)p] == r) jac[p] = 1.0;
  if (compute_jacobian && jac)
    for (int32_t r : P->point_dofs)
      for (int64_t p = M.rowptr[(size_t)r]; p < M.rowptr[(size_t)r + 1]; ++p) jac[p] = (M.colind[(size_t)p] == r) ? 1.0 : 0.0;
  if (compute_jacobian && jac && opt_bool(P, "fix zero rows", false))
    for (int64_t r = 0; r < M.nrows; ++r) {
      double s = 0.0;
      for (int64_t p = M.rowptr[(size_t)r]; p < M.rowptr[(size_t)r + 1]; ++p) s += std::fabs(jac[p]);
      if (s < 1.0e-14) for (int64_t p = M.rowptr[(size_t)r]; p < M.rowptr[(size_t)r + 1]; ++p) if (M.colind[(size_t)p] == r) jac[p] = 1.0;
    }
  ABI_END
}

// shared by assemble_mass / debug_emulate_mass: argument checks
static void check_mass_plan(mrhyde_b200_plan* P, const double* mass_wts) {
  if (!P || !mass_wts) fail(MRHYDE_B200_ERR_INVALID, "assemble_mass: null argument");
  if (!P->finalized) fail(MRHYDE_B200_ERR_STATE, "assemble_mass called before mrhyde_b200_plan_finalize");
  if (!P->use_general)
    fail(MRHYDE_B200_ERR_UNSUPPORTED, "assemble_mass needs a plan on the general path (set option kernel=general for thermal HGRAD-1 blocks)");
}

int mrhyde_b200_assemble_mass(mrhyde_b200_plan* P, const double* mass_wts, int lump, double* mass_values, double* diag, void* stream) {
  ABI_BEGIN
  check_mass_plan(P, mass_wts);
  if (P->device == -1) fail(MRHYDE_B200_ERR_STATE, "host-only analysis plan (device = -1) cannot assemble: there is no CPU path");
  if (!mass_values && !diag) fail(MRHYDE_B200_ERR_INVALID, "assemble_mass: both outputs are null");
  CUDA_OK(cudaSetDevice(P->device));
  GraphDev G{P->d_rowptr.p, P->d_colind.p, P->d_fixed.p};
  GenLaunchStats stats;
  const char* err = gen_assemble_mass(P->gen_dev, P->gen, P->gen_kernels, P->d_vx.p, P->d_vy.p, P->d_vz.p, P->d_conn.p, P->d_lids.p, G, mass_wts, lump != 0,
                                      P->accumulate, mass_values, diag, (cudaStream_t)stream, &stats);
  if (err) fail(MRHYDE_B200_ERR_CUDA, std::string("mass assembly launch: ") + err);
  P->launches_per_assemble = stats.launches;
  ABI_END
}

int mrhyde_b200_apply_mass(mrhyde_b200_plan* P, const double* mass_wts, const double* x, double* y, void* stream) {
  ABI_BEGIN
  check_mass_plan(P, mass_wts);
  if (P->device == -1) fail(MRHYDE_B200_ERR_STATE, "host-only analysis plan (device = -1) cannot assemble: there is no CPU path");
  if (!x || !y) fail(MRHYDE_B200_ERR_INVALID, "apply_mass: null vector");
  CUDA_OK(cudaSetDevice(P->device));
  GraphDev G{P->d_rowptr.p, P->d_colind.p, P->d_fixed.p};
  GenLaunchStats stats;
  const char* err = gen_apply_mass(P->gen_dev, P->gen, P->gen_kernels, P->d_vx.p, P->d_vy.p, P->d_vz.p, P->d_conn.p, P->d_lids.p, G, mass_wts, P->accumulate,
                                   x, y, (cudaStream_t)stream, &stats);
  if (err) fail(MRHYDE_B200_ERR_CUDA, std::string("mass apply launch: ") + err);
  P->launches_per_assemble = stats.launches;
  ABI_END
}

int mrhyde_b200_project_initial(mrhyde_b200_plan* P, double time, double* rhs, void* stream) {
  ABI_BEGIN
  const double ones[GEN_MAXVARS] = {1.0, 1.0, 1.0, 1.0, 1.0};
  check_mass_plan(P, ones);
  if (P->device == -1) fail(MRHYDE_B200_ERR_STATE, "host-only analysis plan (device = -1) cannot assemble: there is no CPU path");
  if (!rhs) fail(MRHYDE_B200_ERR_INVALID, "project_initial: null vector");
  CUDA_OK(cudaSetDevice(P->device));
  GraphDev G{P->d_rowptr.p, P->d_colind.p, P->d_fixed.p};
  GenLaunchStats stats;
  const char* err = gen_project_initial(P->gen_dev, P->gen, P->gen_kernels, P->d_vx.p, P->d_vy.p, P->d_vz.p, P->d_conn.p, P->d_lids.p, G, time, P->accumulate,
                                        rhs, (cudaStream_t)stream, &stats);
  if (err) fail(MRHYDE_B200_ERR_CUDA, std::string("initial projection launch: ") + err);
  P->launches_per_assemble = stats.launches;
  ABI_END
}

int mrhyde_b200_set_initial(mrhyde_b200_plan* P, double time, double* rhs, double* mass_values, void* stream) {
  ABI_BEGIN
  if (!P || !rhs || !mass_values) fail(MRHYDE_B200_ERR_INVALID, "set_initial: null argument");
  if (P->finalized && opt_bool(P, "lump mass", false))
    fail(MRHYDE_B200_ERR_UNSUPPORTED, "set_initial: the lumped insertion of setInitial (assemblyManager_initial.hpp:100-105) is not built");
  const double ones[GEN_MAXVARS] = {1.0, 1.0, 1.0, 1.0, 1.0};   // getMass: unit weights
  int rc = mrhyde_b200_project_initial(P, time, rhs, stream);
  if (rc != MRHYDE_B200_OK) return rc;
  int launches = P->launches_per_assemble;
  rc = mrhyde_b200_assemble_mass(P, ones, 0, mass_values, nullptr, stream);
  if (rc != MRHYDE_B200_OK) return rc;
  launches += P->launches_per_assemble;
  // setInitial fixes empty rows of the mass matrix unconditionally (`bool fix_zero_rows = true`, assemblyManager_initial.hpp:48, 114-131)
  const int64_t n = P->mesh.nrows;
  if (n > 0) {
    fix_zero_rows_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P->d_rowptr.p, P->d_colind.p, n, mass_values);
    CUDA_OK(cudaGetLastError());
    ++launches;
  }
  P->launches_per_assemble = launches;
  ABI_END
}

int mrhyde_b200_plan_debug_emulate_initial(mrhyde_b200_plan* P, double time, double* rhs) {
  ABI_BEGIN
  const double ones[GEN_MAXVARS] = {1.0, 1.0, 1.0, 1.0, 1.0};
  check_mass_plan(P, ones);
  if (P->device != -1) fail(MRHYDE_B200_ERR_STATE, "debug_emulate_initial: only host-only analysis plans (device = -1) replay the kernel stages on the host");
  if (!rhs) fail(MRHYDE_B200_ERR_INVALID, "debug_emulate_initial: null vector");
  const GeneralPlanHost& H = P->gen;
  const GenKernelInfo& I = H.info;
  const MeshGraph& M = P->mesh;
  std::vector<double> er((size_t)H.n_inst * (size_t)I.N, 0.0), zero((size_t)M.nrows, 0.0);
  GenParams Q;
  std::memset(&Q, 0, sizeof(Q));
  Q.tensor = H.use_tensor ? 1 : 0;
  Q.vx = M.vcoord[0].data(); Q.vy = M.vcoord[1].data(); Q.vz = M.vcoord[2].data(); Q.conn = M.conn.data(); Q.lids = M.lids.data();
  Q.orient = M.orient.empty() ? nullptr : M.orient.data();
  Q.sol = zero.data();
  Q.td.alpha_u = 1.0; Q.td.seed_u = 1.0; Q.td.deltat = 1.0; Q.td.time = time;
  std::memcpy(Q.off, H.off, sizeof(Q.off));
  std::memcpy(Q.fn, H.init_fn, sizeof(Q.fn));
  Q.fn_op = H.fn_op.data(); Q.fn_c = H.fn_c.data(); Q.opt = H.opt;
  for (int v = 0; v < GEN_MAXVARS; ++v) { Q.bc_type[v] = 0; Q.bc_fn[v] = -1; Q.mass_wts[v] = 1.0; }
  Q.elem_jac = nullptr; Q.elem_res = er.data();
  Q.mass_mode = 2;
  Q.epb = 3;
  Q.items = nullptr; Q.item_begin = 0; Q.item_end = M.nelem; Q.inst_base = 0;
  Q.geo_N = H.geo_N.data(); Q.geo_dN = H.geo_dN.data(); Q.ref_tab = H.ref_tab.data(); Q.qwts = H.qwts.data();
  need_emulator(P)->emulate(false, Q, (int)((M.nelem + Q.epb - 1) / Q.epb));
  gen_pull_apply_host(H, er.data(), P->accumulate, rhs);
  ABI_END
}

int mrhyde_b200_plan_debug_emulate_mass(mrhyde_b200_plan* P, const double* mass_wts, int lump, double* mass_values, double* diag) {
  ABI_BEGIN
  check_mass_plan(P, mass_wts);
  if (P->device != -1) fail(MRHYDE_B200_ERR_STATE, "debug_emulate_mass: only host-only analysis plans (device = -1) replay the kernel stages on the host");
  const GeneralPlanHost& H = P->gen;
  const GenKernelInfo& I = H.info;
  const MeshGraph& M = P->mesh;
  const size_t N = (size_t)I.N;
  std::vector<double> ej((size_t)H.n_inst * N * N, 0.0), zero((size_t)M.nrows, 0.0);
  GenParams Q;
  std::memset(&Q, 0, sizeof(Q));
  Q.tensor = H.use_tensor ? 1 : 0;
  Q.vx = M.vcoord[0].data(); Q.vy = M.vcoord[1].data(); Q.vz = M.vcoord[2].data(); Q.conn = M.conn.data(); Q.lids = M.lids.data();
  Q.orient = M.orient.empty() ? nullptr : M.orient.data();
  Q.sol = zero.data();
  Q.td.alpha_u = 1.0; Q.td.seed_u = 1.0; Q.td.deltat = 1.0;
  std::memcpy(Q.off, H.off, sizeof(Q.off));
  std::memcpy(Q.fn, H.fn, sizeof(Q.fn));
  Q.fn_op = H.fn_op.data(); Q.fn_c = H.fn_c.data(); Q.opt = H.opt;
  for (int v = 0; v < GEN_MAXVARS; ++v) { Q.bc_type[v] = 0; Q.bc_fn[v] = -1; }
  Q.elem_jac = ej.data(); Q.elem_res = nullptr;
  Q.mass_mode = 1;
  for (int v = 0; v < I.nvars; ++v) Q.mass_wts[v] = mass_wts[v];
  Q.epb = 3;
  Q.items = nullptr; Q.item_begin = 0; Q.item_end = M.nelem; Q.inst_base = 0;
  Q.geo_N = H.geo_N.data(); Q.geo_dN = H.geo_dN.data(); Q.ref_tab = H.ref_tab.data(); Q.qwts = H.qwts.data();
  need_emulator(P)->emulate(false, Q, (int)((M.nelem + Q.epb - 1) / Q.epb));
  gen_pull_mass_host(H, M, ej.data(), P->accumulate, lump != 0, mass_values, diag);
  ABI_END
}

int mrhyde_b200_plan_debug_emulate_apply_mass(mrhyde_b200_plan* P, const double* mass_wts, const double* x, double* y) {
  ABI_BEGIN
  check_mass_plan(P, mass_wts);
  if (P->device != -1) fail(MRHYDE_B200_ERR_STATE, "debug_emulate_apply_mass: only host-only analysis plans (device = -1) replay the kernel stages on the host");
  if (!x || !y) fail(MRHYDE_B200_ERR_INVALID, "debug_emulate_apply_mass: null vector");
  const GeneralPlanHost& H = P->gen;
  const GenKernelInfo& I = H.info;
  const MeshGraph& M = P->mesh;
  std::vector<double> er((size_t)H.n_inst * (size_t)I.N, 0.0);
  GenParams Q;
  std::memset(&Q, 0, sizeof(Q));
  Q.tensor = H.use_tensor ? 1 : 0;
  Q.vx = M.vcoord[0].data(); Q.vy = M.vcoord[1].data(); Q.vz = M.vcoord[2].data(); Q.conn = M.conn.data(); Q.lids = M.lids.data();
  Q.orient = M.orient.empty() ? nullptr : M.orient.data();
  Q.sol = x;
  Q.td.alpha_u = 1.0; Q.td.seed_u = 1.0; Q.td.deltat = 1.0;
  std::memcpy(Q.off, H.off, sizeof(Q.off));
  std::memcpy(Q.fn, H.fn, sizeof(Q.fn));
  Q.fn_op = H.fn_op.data(); Q.fn_c = H.fn_c.data(); Q.opt = H.opt;
  for (int v = 0; v < GEN_MAXVARS; ++v) { Q.bc_type[v] = 0; Q.bc_fn[v] = -1; }
  Q.elem_jac = nullptr; Q.elem_res = er.data();
  Q.mass_mode = 1;
  for (int v = 0; v < I.nvars; ++v) Q.mass_wts[v] = mass_wts[v];
  Q.epb = 3;
  Q.items = nullptr; Q.item_begin = 0; Q.item_end = M.nelem; Q.inst_base = 0;
  Q.geo_N = H.geo_N.data(); Q.geo_dN = H.geo_dN.data(); Q.ref_tab = H.ref_tab.data(); Q.qwts = H.qwts.data();
  need_emulator(P)->emulate(false, Q, (int)((M.nelem + Q.epb - 1) / Q.epb));
  gen_pull_apply_host(H, er.data(), P->accumulate, y);
  ABI_END
}

int mrhyde_b200_plan_debug_metric_host(mrhyde_b200_plan* P, const double* sol, const mrhyde_b200_time* t, int accumulate, double* res, double* jac) {
  ABI_BEGIN
  if (!P || !sol) fail(MRHYDE_B200_ERR_INVALID, "debug_metric_host: null argument");
  if (!P->finalized || P->use_general) fail(MRHYDE_B200_ERR_STATE, "debug_metric_host: needs a finalized plan on the sweep kernel");
  if (P->metric_ng <= 0) fail(MRHYDE_B200_ERR_STATE, "debug_metric_host: this plan does not use the metric ring (general cells, non-constant coefficients or ring=full)");
  TimeDev td;
  fill_time(t, td, true);   // host pointers: the replay reads them on the host
  std::vector<double> met;
  const int ng = P->dim * (P->dim + 1) / 2;
  if (P->dim == 3) { metric_host_elements<3>(P, P->th3, sol, td, met); host_apply_metric_plan(P->mesh, P->cp, met.data(), ng, &P->th3.tab.Stab[0][0], &P->th3.tab.Mtab[0], td.seed_u, td.transient ? td.seed_t : 0.0, accumulate != 0, res, jac); }
  else { metric_host_elements<2>(P, P->th2, sol, td, met); host_apply_metric_plan(P->mesh, P->cp, met.data(), ng, &P->th2.tab.Stab[0][0], &P->th2.tab.Mtab[0], td.seed_u, td.transient ? td.seed_t : 0.0, accumulate != 0, res, jac); }
  ABI_END
}

int mrhyde_b200_plan_debug_chain_rows(mrhyde_b200_plan* P, int32_t chain_begin, int32_t chain_end, uint8_t* mask /*[n_rows]*/) {
  ABI_BEGIN
  if (!P || !mask) fail(MRHYDE_B200_ERR_INVALID, "debug_chain_rows: null argument");
  if (!P->finalized || P->use_general) fail(MRHYDE_B200_ERR_STATE, "debug_chain_rows: needs a finalized plan on the sweep kernel");
  const ChainPlan& cp = P->cp;
  if (chain_begin < 0 || chain_end > cp.n_chains || chain_begin > chain_end) fail(MRHYDE_B200_ERR_INVALID, "debug_chain_rows: chain range out of bounds");
  for (int32_t c = chain_begin; c < chain_end; ++c)
    for (int32_t st = cp.chain_step_ptr[(size_t)c]; st < cp.chain_step_ptr[(size_t)c + 1]; ++st) {
      const StepRec& S = cp.steps[(size_t)st];
      for (int32_t b = 0; b < S.n_batches; ++b) {
        const BatchRec& B = cp.batches[(size_t)(S.batch_begin + b)];
        for (int l = 0; l < (int)B.n_rows; ++l) mask[cp.rows[(size_t)(S.batch_begin + b) * 32 + (size_t)l].row] = 1;
      }
    }
  ABI_END
}

// Host replay of the class ring: element metrics and load vectors by the kernel's formulas (metric_host_elements), one local-matrix
// value per class + the residual as thermal_affine stages them, then the plan's scatter programs (host_apply_chain_plan).
int mrhyde_b200_plan_debug_class_host(mrhyde_b200_plan* P, const double* sol, const mrhyde_b200_time* t, int accumulate, double* res, double* jac) {
  ABI_BEGIN
  if (!P || !sol) fail(MRHYDE_B200_ERR_INVALID, "debug_class_host: null argument");
  if (!P->finalized || P->use_general) fail(MRHYDE_B200_ERR_STATE, "debug_class_host: needs a finalized plan on the sweep kernel");
  if (P->class_nc <= 0) fail(MRHYDE_B200_ERR_STATE, "debug_class_host: this plan does not use the class ring (needs box cells, constant coefficients, ring=auto|class)");
  TimeDev td;
  fill_time(t, td, true);   // host pointers: the replay reads them on the host
  const int dim = P->dim, NV = 1 << dim, NT = NV * (NV + 1) / 2, NG = dim * (dim + 1) / 2, SLM = NG + 1 + 3 * NV, NC = P->class_nc, SL = NC + NV;
  std::vector<double> met;
  if (dim == 3) metric_host_elements<3>(P, P->th3, sol, td, met); else metric_host_elements<2>(P, P->th2, sol, td, met);
  const double* St = dim == 3 ? &P->th3.tab.Stab[0][0] : &P->th2.tab.Stab[0][0];
  const double* Mt = dim == 3 ? &P->th3.tab.Mtab[0] : &P->th2.tab.Mtab[0];
  const MeshGraph& M = P->mesh;
  std::vector<double> stage((size_t)M.nelem * SL, 0.0);
  for (int64_t e = 0; e < M.nelem; ++e) {
    const double* m = &met[(size_t)e * SLM];
    double* o = &stage[(size_t)e * SL];
    std::vector<double> kc((size_t)NC), mc((size_t)NC, 0.0);
    for (int c = 0; c < NC; ++c) {
      const int rep = P->class_rep[(size_t)c];
      double k = 0.0;
      for (int g = 0; g < dim; ++g) k += m[g] * St[(size_t)g * NT + rep];   // boxes: the diagonal metric entries only
      kc[(size_t)c] = k;
      if (td.transient) { mc[(size_t)c] = m[NG] * Mt[rep]; k = td.seed_u * k + td.seed_t * mc[(size_t)c]; }
      o[c] = k;
    }
    for (int i = 0; i < NV; ++i) {
      double r = 0.0;
      for (int j = 0; j < NV; ++j) {
        const int a = std::min(i, j), b = std::max(i, j), c = P->class_of_t[(size_t)(a * NV - (a * (a - 1)) / 2 + (b - a))];
        r += kc[(size_t)c] * m[NG + 1 + NV + j];
        if (td.transient) r += mc[(size_t)c] * m[NG + 1 + 2 * NV + j];
      }
      o[NC + i] = r - m[NG + 1 + i];
    }
  }
  host_apply_chain_plan(M, P->cp, stage.data(), accumulate != 0, res, jac);
  ABI_END
}

int mrhyde_b200_plan_debug_stage_map(mrhyde_b200_plan* P, int32_t* kmap /*[ndof*ndof]*/, int32_t* rmap /*[ndof]*/) {
  ABI_BEGIN
  if (!P || !kmap || !rmap) fail(MRHYDE_B200_ERR_INVALID, "debug_stage_map: null argument");
  if (!P->finalized || P->use_general) fail(MRHYDE_B200_ERR_STATE, "debug_stage_map: needs a finalized plan on the sweep kernel");
  for (size_t i = 0; i < P->kmap.size(); ++i) kmap[i] = P->kmap[i];
  for (size_t i = 0; i < P->rmap.size(); ++i) rmap[i] = P->rmap[i];
  ABI_END
}

int mrhyde_b200_plan_debug_scatter_host(mrhyde_b200_plan* P, const double* stage, int64_t stage_len, int accumulate, double* res, double* jac) {
  ABI_BEGIN
  if (!P || !stage) fail(MRHYDE_B200_ERR_INVALID, "debug_scatter_host: null argument");
  if (!P->finalized) fail(MRHYDE_B200_ERR_STATE, "debug_scatter_host before finalize");
  if (stage_len != P->stage_len) fail(MRHYDE_B200_ERR_INVALID, "debug_scatter_host: stage_len must be " + std::to_string(P->stage_len));
  host_apply_chain_plan(P->mesh, P->cp, stage, accumulate != 0, res, jac);
  ABI_END
}

}  // extern "C"
