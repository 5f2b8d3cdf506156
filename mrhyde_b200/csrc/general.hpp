// Host side of the general assembly path (any restated module, any supported basis): plan-time analysis for the
// deterministic pull (the fused scatter of assemblyManager_scatter.hpp:162-278 turned around), the table of compiled
// kernel instantiations, and the per-call launch sequence.
//
//   element kernel   general_kernel.cuh   -> scratch element matrices / vectors (one instance per element, then one per
//                                            boundary (element, side) in boundary-group order: the reference's order of
//                                            contributions, assemblyManager_jacres.hpp:336-603)
//   pull kernel      general.cu           -> every CSR row sums its instances in ascending order and is written once
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "general_kernel.cuh"
#include "plan.hpp"

namespace mrhyde_b200 {

// One compiled instantiation <Phys, NQ, NQS, K>.
struct GenKernelInfo {
  const char* physics;
  int dim, order, nq, nqs;
  int N, nvars, nbasis, nfn, K;
  int tpe;                                      // threads per element in the derivative stage
  int tensor;                                   // 1: Jacobian by field-direction derivatives + FP64 tensor-core contraction (general_kernel.cuh, S4d / S4m)
  int max_threads, min_blocks;                  // launch bounds of the element kernel (derivative-lane build)
  int tc_max_threads, tc_min_blocks;            // launch bounds of the tensor-core build
  int smem_doubles_volume, smem_doubles_side;   // per element (derivative-lane build)
  int tc_smem_doubles_volume, tc_smem_doubles_side;   // per element (tensor-core build: + the D region)
  int card[2], ncb[2];                          // per basis
  int var_basis[GEN_MAXVARS];
};
struct GenDeviceKernels {
  GenKernelInfo info;
  // returns nullptr on success, else a static error string
  // [state]: 0 = no coefficient function reads a solution field, 1 = the build that evaluates / differentiates such functions
  const char* (*launch[2])(bool side, const GenParams& P, int nblocks, int threads, size_t smem, void* stream);
  const char* (*launch_tc[2])(bool side, const GenParams& P, int nblocks, int threads, size_t smem, void* stream);   // null: no tensor-core build
};
struct GenHostKernels {
  GenKernelInfo info;
  void (*emulate)(bool side, const GenParams& P, int nblocks);
};
const GenDeviceKernels* gen_find_device(const std::string& physics, int dim, int order, int nq, int nqs);
// host replay of the kernel stages: lives in the test-only library libmrhyde_b200_emulate.so and is registered at run time
// (mrhyde_b200_debug_set_emulator); null in a process that has not loaded it
typedef const void* (*GenEmulatorLookup)(const char* physics, int dim, int order, int nq, int nqs);
void gen_set_emulator(GenEmulatorLookup fn);
const GenHostKernels* gen_find_host(const std::string& physics, int dim, int order, int nq, int nqs);
std::string gen_supported_list();

struct GenSideFamily {      // one boundary group: (sideset, local side) with its side cubature and tables
  int sideset = 0, local_side = 0, nqs = 0;
  std::vector<int32_t> items;          // element ids
  int64_t inst_base = 0;               // scratch instance of items[0]
  std::vector<double> geo_N, geo_dN, ref_tab, qwts;
  double tan_u[3] = {0, 0, 0}, tan_v[3] = {0, 0, 0};
  int32_t bc_type[GEN_MAXVARS] = {0, 0, 0, 0, 0};
  int32_t bc_fn[GEN_MAXVARS] = {-1, -1, -1, -1, -1};
  GenFnRec fn[GEN_MAXFN];              // module functions at side ip, then the boundary data of each variable
  bool active = false;                 // some variable has a Neumann / weak Dirichlet condition here
};

struct GenBatch {
  int64_t elem_begin = 0, elem_end = 0;   // element kernel range
  int64_t row_begin = 0, row_end = 0;     // rows (in row_order) that are complete after this batch
};

struct GeneralPlanHost {
  GenKernelInfo info;
  int64_t n_elem = 0, n_inst = 0, n_rows = 0, n_owned = 0;
  int32_t max_row_len = 0;
  int epb_override = 0;                  // option "elements per cta" (0 = automatic)
  bool lump_mass = false;                // Solver: lump mass (fused scatter's column redirect to the diagonal)
  bool use_tensor = false;               // option "jacobian" = auto | tensor | lanes: which build of the element kernel assembles Jacobians
  // pull schedule
  std::vector<int32_t> row_order;        // rows sorted by completion batch
  std::vector<int64_t> contrib_ptr;      // [n_rows+1] in row_order order
  std::vector<int32_t> contrib;          // inst * N + local row, ascending instance
  std::vector<uint16_t> pos_tab;         // [n_pos_patterns][N][N]: position of column LID(c) inside row LID(i); instances with
  std::vector<int32_t> pos_id;           // identical tables share one (pos_id[inst]): a handful on structured meshes, L2-resident
  // element scratch: volume instances live in a RING of `scratch_cap` instances (slot = inst % scratch_cap, a multiple of the batch
  // size), chosen at plan time so that an instance is overwritten only after every row it feeds has been pulled; boundary-side
  // instances follow the ring (slot = scratch_cap + inst - n_elem).  One batch: scratch_cap = n_elem.
  int64_t scratch_cap = 0;
  std::vector<GenBatch> batches;         // volume batches; rows completed by side instances sit in the last batch
  // launch tables (volume)
  std::vector<double> geo_N, geo_dN, ref_tab, qwts;
  std::vector<uint8_t> fn_op;
  std::vector<double> fn_c;
  GenFnRec fn[GEN_MAXFN];
  GenFnRec init_fn[GEN_MAXFN];           // "initial <var>[...]" functions in (variable, component) order (setInitial)
  int16_t off[GEN_MAXVARS][GEN_MAXDOF];
  GenOpts opt;
  std::vector<GenSideFamily> sides;
};

// builds contrib / pos / row_order / batches from the mesh graph and the side families' items
// batch_elems <= 0: chosen from scratch_budget_bytes (the largest batch whose ring fits; one batch if everything fits)
void gen_build_pull(const MeshGraph& m, int N, const std::vector<GenSideFamily>& sides, int64_t batch_elems, int64_t scratch_budget_bytes, GeneralPlanHost& out);
inline int64_t gen_slot(const GeneralPlanHost& H, int64_t inst) { return inst < H.n_elem ? inst % H.scratch_cap : H.scratch_cap + (inst - H.n_elem); }
inline int64_t gen_scratch_instances(const GeneralPlanHost& H) { return H.scratch_cap + (H.n_inst - H.n_elem); }

// host replay of the pull (plan verification and the emulation hook)
void gen_pull_host(const GeneralPlanHost& H, const MeshGraph& m, const double* elem_jac, const double* elem_res, bool accumulate,
                   double* res, double* jac);

// weighted-mass variant of the pull: no fixed-row skipping, boundary instances ignored, diagonal vector = Jacobi diagonal or lumped
void gen_pull_mass_host(const GeneralPlanHost& H, const MeshGraph& m, const double* elem_jac, bool accumulate, bool lump, double* mass, double* diag);

// applyMassMatrixFree: y (+)= sum over volume instances of the element vectors (no sign flip, no fixed-dof handling)
void gen_pull_apply_host(const GeneralPlanHost& H, const double* elem_res, bool accumulate, double* y);

// ---- device side (general.cu) ---------------------------------------------------------------------------------
struct GeneralPlanDev;   // device buffers
struct GenLaunchStats { int launches = 0; };
GeneralPlanDev* gen_upload(const GeneralPlanHost& H, const MeshGraph& m, size_t* dev_bytes, std::string& err);
void gen_free(GeneralPlanDev* D);
// the whole assemble call: element kernels + pull per batch.  Returns nullptr or an error string.
const char* gen_assemble(GeneralPlanDev* D, const GeneralPlanHost& H, const GenDeviceKernels* kd, const double* vx, const double* vy, const double* vz,
                         const int32_t* conn, const int32_t* lids, const GraphDev& G, const OutDev& O, const double* sol, const TimeDev& td,
                         bool volume, bool boundary, void* stream, GenLaunchStats* stats, bool adjoint = false);
// getWeightedMass: element kernel in mass mode over the volume elements + mass pull.  mass / diag may be null.
const char* gen_assemble_mass(GeneralPlanDev* D, const GeneralPlanHost& H, const GenDeviceKernels* kd, const double* vx, const double* vy, const double* vz,
                              const int32_t* conn, const int32_t* lids, const GraphDev& G, const double* mass_wts, bool lump, bool accumulate,
                              double* mass, double* diag, void* stream, GenLaunchStats* stats);
// setInitial, projection right-hand side: rhs (+)= sum_e sum_q initial(x_q) phi_i w (element kernel in initial mode, residual stage only)
const char* gen_project_initial(GeneralPlanDev* D, const GeneralPlanHost& H, const GenDeviceKernels* kd, const double* vx, const double* vy, const double* vz,
                                const int32_t* conn, const int32_t* lids, const GraphDev& G, double time, bool accumulate, double* rhs, void* stream,
                                GenLaunchStats* stats);
// applyMassMatrixFree: y (+)= M x without forming M (element kernel in mass mode, residual stage only)
const char* gen_apply_mass(GeneralPlanDev* D, const GeneralPlanHost& H, const GenDeviceKernels* kd, const double* vx, const double* vy, const double* vz,
                           const int32_t* conn, const int32_t* lids, const GraphDev& G, const double* mass_wts, bool accumulate, const double* x, double* y,
                           void* stream, GenLaunchStats* stats);

}  // namespace mrhyde_b200
