/* mrhyde_b200.h -- C ABI of the B200-native element residual/Jacobian assembly.
 *
 * This library replaces the body of MrHyDE's
 *     AssemblyManager<Node>::assembleJacRes<EvalT>(set, ..., block)
 *         src/managers/assembly/assemblyManager_jacres.hpp:119-630   (group loop :336-603,
 *         gather gather.hpp:181-234, updateWorkset workset.hpp:963-1049, physics
 *         volumeResidual/boundaryResidual physicsInterface_residual.hpp:13-216, fused scatter
 *         scatter.hpp:162-278, dofConstraints constraints.hpp:241-270)
 * and of assembleRes (jacres.hpp:668-875) with hand-written sm_100a kernels.  The reference host
 * code (DOF managers, Tpetra graph/vectors, solvers, YAML) stays as it is; it hands this library
 * the raw arrays those objects already hold.  INTEGRATION.md shows the call sites.
 *
 * Conventions
 *   - every entry point returns an int status (MRHYDE_B200_OK == 0); no C++ exception crosses the
 *     boundary; mrhyde_b200_last_error() returns the message of the last failure on this thread.
 *   - "host" pointers are read during the call and never retained unless stated; "device" pointers
 *     are CUDA device pointers on the plan's device.
 *   - LO = int32_t (preferences.hpp:39-97), row_map = int64_t (KokkosSparse size_type), fp64 only.
 *   - there is no CPU fallback: every assemble call launches the CUDA kernels or fails.
 *   - a plan is not thread-safe (the reference is not re-entrant either: shared workset scratch,
 *     assemblyManager_jacres.hpp:334).
 */
#ifndef MRHYDE_B200_H
#define MRHYDE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRHYDE_B200_OK 0
#define MRHYDE_B200_ERR_INVALID 1     /* bad argument / inconsistent sizes                      */
#define MRHYDE_B200_ERR_UNSUPPORTED 2 /* valid reference input this build has no kernel for       */
#define MRHYDE_B200_ERR_PARSE 3       /* expression could not be decomposed (reference wording)   */
#define MRHYDE_B200_ERR_CUDA 4        /* CUDA runtime / launch failure                             */
#define MRHYDE_B200_ERR_STATE 5       /* call order violated (e.g. assemble before finalize)       */
#define MRHYDE_B200_ERR_NCCL 6

typedef struct mrhyde_b200_plan mrhyde_b200_plan; /* opaque, library-owned */

/* Reference basis tabulated at the reference cubature points -- exactly what
 * DiscretizationInterface::setReferenceBasisData stores (discretizationInterface_basis.hpp:12-133),
 * so a MrHyDE host passes Intrepid2's own numbers.  Layouts are row-major. */
typedef struct {
  const char* type;      /* "HGRAD" | "HCURL" | "HDIV" | "HVOL"                                   */
  int32_t order;
  int32_t card;          /* basis cardinality                                                    */
  const double* val;     /* [card][nqp][vdim]  vdim = 1 (HGRAD/HVOL) or dim (HCURL/HDIV)          */
  const double* grad;    /* [card][nqp][dim]   HGRAD, else NULL                                  */
  const double* curl;    /* [card][nqp][dim]   HCURL (3-D), else NULL                            */
  const double* div;     /* [card][nqp]        HDIV, else NULL                                   */
} mrhyde_b200_basis;

/* One block of one physics set: what AssemblyManager/PhysicsInterface know after construction. */
typedef struct {
  const char* physics;          /* module name as in PhysicsImporter::import (physicsImporter.cpp:64-281):
                                   "thermal" | "linearelasticity" | "navierstokes" | "maxwell"  */
  int32_t dim;                  /* 2 | 3 ; cell topology is Quadrilateral_4 / Hexahedron_8       */
  int32_t nvars;
  const char* const* var_names; /* module variable order, e.g. {"ux","pr","uy","uz"}             */
  const int32_t* var_basis;     /* [nvars] index into bases[]                                    */
  int32_t nbases;
  const mrhyde_b200_basis* bases;
  int32_t ndof_elem;            /* LIDs per element                                              */
  const int32_t* offsets;       /* [nvars][max_card], wkset->offsets (getGIDFieldOffsets), -1 padded */
  int32_t max_card;
  int32_t nqp;                  /* volume cubature (DefaultCubatureFactory, _construct.hpp:103-112) */
  const double* qp_pts;         /* [nqp][dim] reference points                                   */
  const double* qp_wts;         /* [nqp]                                                         */
} mrhyde_b200_desc;

/* Time-integration data of one assembleJacRes call (updateWorksetTime / computeSolnTransientSeeded,
 * workset.cpp:600-834; Butcher and BDF tables solverManager_setup.hpp:181-480).  NULL = steady. */
typedef struct {
  double time;                  /* t_n; stage time is time + butcher_c[stage]*deltat             */
  double deltat;
  int32_t stage;
  int32_t nstages;
  const double* butcher_A;      /* host [nstages][nstages]                                       */
  const double* butcher_b;      /* host [nstages]                                                */
  const double* butcher_c;      /* host [nstages]                                                */
  int32_t nbdf;                 /* number of BDF weights (order+1)                               */
  const double* bdf_wts;        /* host [nbdf]                                                   */
  const double* const* sol_prev;  /* host array of [nbdf-1] device vectors u_{n}, u_{n-1}, ...    */
  const double* const* sol_stage; /* host array of [nstages] device vectors (stages < stage read) */
  /* What the Jacobian is the derivative with respect to (seedwhat / seedindex of updateWorkset, assemblyManager_jacres.hpp:176-190;
   * computeSolnTransientSeeded, workset.cpp:622-785): 0 or 1 = the stage solution `sol` (assembleJacRes' default), 2 = previous step
   * seed_index (compute_previous_jac: sol_prev[seed_index]), 3 = previous stage seed_index (sol_stage[seed_index]).  The residual and
   * the evaluation point are the same in all modes. */
  int32_t seed_what;
  int32_t seed_index;
} mrhyde_b200_time;

/* Side data of one boundary group family (BoundaryGroup, assemblyManager_groups.hpp; reference
 * side cubature + side-tabulated bases as produced by getPhysicalBoundaryIntegrationData,
 * discretizationInterface_integration.hpp:592-774). */
typedef struct {
  int32_t sideset;              /* index into the side-name list given to set_sidesets           */
  int32_t local_side;           /* Shards local side id of the cell                              */
  int32_t n_elem;
  const int32_t* elem_ids;      /* host [n_elem] element indices (as in set_mesh)                */
  int32_t nqp_side;
  const double* side_pts;       /* host [nqp_side][dim] side points in the CELL reference frame  */
  const double* side_wts;       /* host [nqp_side]                                               */
  const double* tangent_u;      /* host [3] reference side tangent(s) (getReferenceFaceTangents) */
  const double* tangent_v;      /* host [3]                                                      */
  const mrhyde_b200_basis* side_bases; /* [nbases] tables at side_pts (val/grad)                 */
} mrhyde_b200_boundary_group;

const char* mrhyde_b200_version(void);
const char* mrhyde_b200_last_error(void);

/* ---- plan construction (host side, once per block/set) -------------------------------------- */
int mrhyde_b200_plan_create(mrhyde_b200_plan** plan, const mrhyde_b200_desc* desc, int device);
void mrhyde_b200_plan_destroy(mrhyde_b200_plan* plan);

/* `Functions:` sublist entry, reference grammar (functionManager_create.hpp:78-540, interpreter.cpp).
 * Module defaults (e.g. thermal.cpp:47-65) are registered by the library; user entries override. */
int mrhyde_b200_plan_set_function(mrhyde_b200_plan* plan, const char* name, const char* expression);

/* Module / solver flags by their YAML key ("useSUPG", "usePSPG", "form_param", "include advection",
 * "use strong DBCs", "assemble boundary terms", ...) plus library keys:
 *   "ns3d_uz_rows" = "reference" | "corrected"   (navierstokes.cpp:688)
 *   "accumulate"   = "true" (reference contract: sum into caller-zeroed res/J) | "false" (overwrite)
 *   "jit"          = "auto" (default: specialise the kernel with NVRTC when available) | "true" | "false"
 *   "kernel"       = "auto" (default: the sweep kernel where it applies -- thermal, HGRAD order 1 -- else the general
 *                    element kernel + pull) | "general" | "sweep"
 *   "batch elems"  = elements per launch of the general path (default -1: one launch over all elements; N > 0: rows are
 *                    pulled as soon as the batch that completes them has been computed)
 *   "penalty", "incplanestress"   linearelasticity module keys
 *   "ring"         = "auto" (default) | "class" | "metric" | "full": what a sweep step stages per element;
 *                    asking for a layout the mesh / coefficients do not allow is MRHYDE_B200_ERR_UNSUPPORTED
 *   "column elements", "min chains", "min segment levels", "sweep axis", "threads", "cta slots", "max blocks", "min blocks",
 *   "max registers", "pull patterns", "pull group", "flush", "flush unroll", "stage1", "stage2", "stagger ns"
 *                    sweep-plan / specialised-build tuning (defaults are the measured best)
 *   "lump mass"      = Solver: lump mass -- the fused scatter adds every entry of a row to its diagonal (assemblyManager_scatter.hpp:263-268;
 *                      general path; a thermal plan that asks for it leaves the sweep kernel)
 *   "fix zero rows"  = Solver: fix zero rows -- after the constraints, rows with sum |J(row,:)| < 1e-14 get a unit diagonal (assemblyManager_jacres.hpp:609-626)
 *   "jacobian"       = "auto" | "lanes" | "tensor": derivative-lane or FP64 tensor-core (mma.m8n8k4.f64) build of the general path's Jacobian stage
 *   "scratch GB"     = budget of the general path's element scratch ring (default: a quarter of the free device memory)
 *   "halo transport" = "auto" | "p2p" | "nccl";  "halo push" = in-kernel push of ghost rows into the owner's slab (default true);  "overlap halo"
 *   "prefetch" = true | false | lean, "prefetch records", "column cache" = true | false | registers, "pipeline", "store hint"
 *                    further sweep-kernel build options (each was measured; defaults are the measured best)
 *   "debug skip", "debug transient", "debug mode"   kernel-debugging keys, accepted only with MRHYDE_B200_DEBUG_OPTIONS=1 in the environment
 *                    ("debug skip" compiles parts of the specialised kernel out: results are then wrong)
 * Unknown keys are an error, never silently ignored. */
int mrhyde_b200_plan_set_option(mrhyde_b200_plan* plan, const char* key, const char* value);

/* Element data of the block: Group::nodes / Group::LIDs for every element, group order
 * (assemblyManager_groups.hpp:413-426).  elem_nodes[n_elem][nverts][dim], lids[n_elem][ndof_elem],
 * orient_sign[n_elem][ndof_elem] (+1/-1 per dof after Intrepid2 orientation; NULL = all +1). */
int mrhyde_b200_plan_set_mesh(mrhyde_b200_plan* plan, int64_t n_elem, const double* elem_nodes,
                              const int32_t* lids, const int8_t* orient_sign);
/* Same, from a vertex table + connectivity (skips the coordinate de-duplication). */
int mrhyde_b200_plan_set_mesh_indexed(mrhyde_b200_plan* plan, int64_t n_verts, const double* vert_coords /*[n_verts][dim]*/,
                                      int64_t n_elem, const int32_t* conn /*[n_elem][nverts]*/,
                                      const int32_t* lids, const int8_t* orient_sign);

/* The overlapped CrsGraph of J (J->getLocalMatrixDevice().graph, jacres.hpp:145-151) and the
 * isFixedDOF mask (assemblyManager_constraints.hpp:13-92).  n_owned <= n_rows: rows [n_owned,n_rows)
 * are ghost rows (overlapped_map = owned then ghosted, discretizationInterface_dof.hpp:129-137). */
int mrhyde_b200_plan_set_graph(mrhyde_b200_plan* plan, int64_t n_rows, int64_t n_owned, const int64_t* row_map,
                               const int32_t* entries, const uint8_t* is_fixed_dof);

/* Boundary conditions: side names, then per (variable, sideset) a type
 * "none" | "Dirichlet" | "weak Dirichlet" | "Neumann" and the data expression. */
int mrhyde_b200_plan_set_sidesets(mrhyde_b200_plan* plan, int32_t n_sides, const char* const* side_names);
int mrhyde_b200_plan_set_bc(mrhyde_b200_plan* plan, const char* var, const char* side, const char* type, const char* expression);
int mrhyde_b200_plan_add_boundary_group(mrhyde_b200_plan* plan, const mrhyde_b200_boundary_group* bg);

/* Builds patches, scatter programs and compiled expressions and uploads them. */
int mrhyde_b200_plan_finalize(mrhyde_b200_plan* plan);

/* ---- the hot path ---------------------------------------------------------------------------- */
/* assembleJacRes (jacres.hpp:119-630): res += -F(u) on free rows, jac_values += dF/du on free rows,
 * J(d,d) = 1 on strong-Dirichlet rows when compute_jacobian (dofConstraints).  With option
 * accumulate=false the outputs are overwritten instead (caller need not zero them).
 * sol/res: [n_rows]; jac_values: [nnz] in the order of `entries`.  `stream` is a cudaStream_t. */
int mrhyde_b200_assemble_jacres(mrhyde_b200_plan* plan, const double* sol, const mrhyde_b200_time* t,
                                int compute_jacobian, int compute_residual, double* res, double* jac_values,
                                void* stream);
/* assembleRes (jacres.hpp:668-875): residual only (the ScalarT workset path). */
/* assembleJacRes(..., useadjoint = true, ...) (assemblyManager_jacres.hpp:119-630 with the useadjoint branches: updateJac :1459-1475,
 * updateJacBoundary :1047-1062): the residual is the forward one, every local Jacobian is filled TRANSPOSED before the scatter --
 * entry (row, col) receives d res(col) / d u(row), strong-Dirichlet rows are skipped as rows of the transposed matrix -- and thermal's
 * weak-Dirichlet sides use sf = 1 (thermal.cpp:197-201).  Seeding modes for discretized-parameter sensitivities (seedwhat = 3) and the
 * previous-step Jacobians (seedwhat = 2) are not built. */
int mrhyde_b200_assemble_jacres_adjoint(mrhyde_b200_plan* plan, const double* sol, const mrhyde_b200_time* t, double* res, double* jac_values, void* stream);
/* Builds the plan-specialised kernel variant an upcoming call will need -- (steady | transient stage) x (residual, Jacobian, both) in
 * the plan's output mode -- so that no assemble call compiles (NVRTC) on the hot path and build failures surface at set-up.
 * finalize pre-builds the steady residual + Jacobian variant only.  No-op for plans on the ahead-of-time kernels. */
int mrhyde_b200_plan_warmup(mrhyde_b200_plan* plan, int transient, int compute_jacobian, int compute_residual);
int mrhyde_b200_assemble_res(mrhyde_b200_plan* plan, const double* sol, const mrhyde_b200_time* t, double* res,
                             void* stream);
/* Same as assemble_jacres but sol/res/jac_values (and t->sol_prev/sol_stage) are HOST buffers:
 * copies in, assembles, copies out, synchronises. */
int mrhyde_b200_assemble_jacres_host(mrhyde_b200_plan* plan, const double* sol, const mrhyde_b200_time* t,
                                     int compute_jacobian, int compute_residual, double* res, double* jac_values);

/* getWeightedMass (assemblyManager_mass.hpp:13-275; element matrices :1065-1146), SURVEY 8(f) rank 1: the per-variable mass
 * blocks  M[(n,i),(n,j)] = mass_wts[n] * sum_q (phi_i . phi_j) w  summed into mass_values (order of the graph's entries;
 * entries that couple different variables receive nothing) and the diagonal vector the reference builds beside it:
 * lump = 0 -> the Jacobi diagonal M(r,r); lump = 1 -> the lumped mass, sum over the row's elements of sum_j |M_e(r,j)|.
 * isFixedDOF is not consulted (the reference does not either).  mass_wts: host [nvars] (physics->mass_wts[set][block]);
 * mass_values [nnz] / diag [n_rows]: device pointers, either may be NULL; both follow the plan's accumulate option.
 * Needs a plan on the general path (every module except thermal HGRAD-1, or option kernel=general). */
int mrhyde_b200_assemble_mass(mrhyde_b200_plan* plan, const double* mass_wts, int lump, double* mass_values, double* diag,
                              void* stream);

/* applyMassMatrixFree (assemblyManager_mass.hpp:555-800): y (+)= M x with the same weighted per-variable mass blocks, without
 * forming M (the element kernel's residual stage on x, then a row sum).  x, y: device [n_rows]; y follows the accumulate option
 * (the reference adds into y). */
int mrhyde_b200_apply_mass(mrhyde_b200_plan* plan, const double* mass_wts, const double* x, double* y, void* stream);
/* setInitial, right-hand side of the L2 projection of the initial conditions (assemblyManager_initial.hpp:36-76 with
 * getInitial(project = true), :260-311): rhs(row) (+)= sum_e sum_q initial_var(x_q) phi_i(x_q) w_q, vector-valued for HCURL / HDIV
 * variables.  The `Initial conditions` entries are passed through plan_set_function as "initial <var>" (HGRAD / HVOL) or
 * "initial <var>[x]" / "[y]" / "[z]" (HCURL / HDIV) -- the names the reference registers, physicsInterface_functions.hpp:154-226;
 * unnamed variables start from 0.0.  The mass matrix of the projection is mrhyde_b200_assemble_mass with unit weights.  No
 * isFixedDOF check (the reference has none here); rhs: device [n_rows], follows the accumulate option; time = initial time. */
int mrhyde_b200_project_initial(mrhyde_b200_plan* plan, double time, double* rhs, void* stream);
/* setInitial as a whole (assemblyManager_initial.hpp:36-133): rhs as project_initial, mass_values = getMass (unit weights) summed into the
 * graph's entries, then the routine's own fix_zero_rows loop (always on there: rows with sum |M(row,:)| < 1e-14 get a unit diagonal).
 * Both outputs follow the accumulate option.  The lumped insertion (Solver: lump mass) is not built: ERR_UNSUPPORTED. */
int mrhyde_b200_set_initial(mrhyde_b200_plan* plan, double time, double* rhs, double* mass_values, void* stream);

/* ---- multi-GPU: the Tpetra Export(overlapped -> owned, ADD) replacement ---------------------------
 * (linearAlgebraInterface_matrix.hpp:233-237, _vector.hpp:56-66).  Ghost rows of this rank are summed
 * into the owning rank's rows in fixed neighbour-rank order. */
int mrhyde_b200_comm_unique_id(uint8_t* id128);                      /* ncclGetUniqueId, 128 bytes */
int mrhyde_b200_plan_comm_init(mrhyde_b200_plan* plan, const uint8_t* id128, int rank, int nranks);
/* Global ids of this rank's local column indices: entries [0,n_rows) are the rows (owned then ghost, the
 * overlapped map), entries [n_rows,n_cols) are column-only ghosts -- remote columns that owned interface rows
 * need so that the summed row is complete (the column map of the owned Tpetra matrix J that
 * exportMatrixFromOverlapped fills).  n_cols >= n_rows.  Collective over the communicator. */
int mrhyde_b200_plan_set_halo(mrhyde_b200_plan* plan, int64_t n_cols, const int64_t* col_gids /*host [n_cols]*/);
/* Deviation on strong-Dirichlet rows that sit on a partition interface: the library writes J(d,d) = 1 on the OWNED copy only, ghost
 * copies of fixed rows stay zero, so after halo_sum the diagonal is 1.  The reference calls replaceLocalValues on every overlapped
 * dbc dof (assemblyManager_constraints.hpp:125-138), and its Export(ADD) therefore leaves the sharing multiplicity (2, 4, ...) on such
 * diagonals.  Residual entries of fixed rows are 0 in both, so the Newton update of those dofs is 0 either way.  Point-constrained
 * rows (plan_set_point_dofs) DO follow the reference: every rank that holds the row sets the identity row before the sum. */
int mrhyde_b200_halo_sum(mrhyde_b200_plan* plan, double* res, double* jac_values, void* stream);
/* The OWNED matrix after halo_sum, i.e. what the reference obtains with fillComplete(J_over) -> Export(ADD) -> fillComplete(J)
 * (solverManager_solvers.hpp:463-466, linearAlgebraInterface_matrix.hpp:233-237), needs no compaction pass here: owned rows come
 * first in the overlapped numbering, so the owned CSR matrix is the PREFIX of the arrays the caller already holds --
 *   rows [0, n_owned_rows), row_map[0 .. n_owned_rows], entries / jac_values [0, nnz_owned) --
 * with column ids in the local numbering of set_halo's col_gids (owned, ghost rows, column-only ghosts = the Tpetra column map of J).
 * Ghost rows [n_owned_rows, n_rows) hold this rank's partial sums only and are not part of the owned matrix. */
int mrhyde_b200_plan_owned_extent(mrhyde_b200_plan* plan, int64_t* n_owned_rows, int64_t* nnz_owned);

/* Point constraints (disc->point_dofs, discretizationInterface_dof.hpp:365-628): dofConstraints replaces the whole Jacobian row of every
 * such dof by the identity row after the assembly, whatever the output mode; the residual entry is left as assembled
 * (setJacobianConstraints(J, dofs, block, ...), assemblyManager_constraints.hpp:97-116, 261-266).  lids: local row ids (owned or ghost;
 * like the reference, every rank that holds the row sets it before the export).  n = 0 clears the list.  Call after plan_finalize. */
int mrhyde_b200_plan_set_point_dofs(mrhyde_b200_plan* plan, int64_t n, const int32_t* lids /*host [n]*/);

/* ---- introspection (tests, bench) ----------------------------------------------------------- */
/* keys: "n_chains" "n_columns" "n_segments" "n_levels" "n_steps" "n_patterns" "n_pattern_slots" "n_batches" "max_batches_per_step" "ring_capacity"
 *       "max_rows_per_step" "kernel_launches_per_assemble" "halo_launches_per_sum" "smem_bytes" "threads_per_block"
 *       "n_elem" "n_elem_with_halo" "n_rows" "nnz" "n_verts" "plan_device_bytes" "n_affine" "n_box"
 *       "n_orphan_rows" "jit" (1 = plan-specialised NVRTC kernel in use) "jit_registers"                        */
int mrhyde_b200_plan_stat(mrhyde_b200_plan* plan, const char* key, int64_t* value);
/* Average device time (ms, CUDA events on the launch stream) of the volume kernel over the
 * assemble calls since the last reset; used by bench.py for the roofline figure. */
int mrhyde_b200_plan_kernel_time(mrhyde_b200_plan* plan, int reset, double* avg_ms, int64_t* n_launches);
/* Evaluates a registered function at points (host), through the device expression kernel. */
int mrhyde_b200_plan_eval_function(mrhyde_b200_plan* plan, const char* name, int64_t npts, const double* xyz /*host [npts][3]*/,
                                   double time, double* out /*host [npts]*/);

/* ---- host-side analysis hooks (no GPU needed; used by the CPU test-suite) ---------------------------
 * A plan created with device = -1 is "host only": set-up and finalize run the full plan analysis
 * (expression compilation, affine classification, patches, scatter programs) but nothing is
 * uploaded and every assemble call fails with MRHYDE_B200_ERR_STATE. */
/* Compiles `which` out of n (name, expression) pairs and writes the flattened program as text. */
int mrhyde_b200_expr_disassemble(int32_t n, const char* const* names, const char* const* exprs, const char* which,
                                 char* out, size_t out_cap);
/* Runs the same flattened program on the host at vars = {x,y,z,t,n[x],n[y],n[z]} (checks of the
 * compiler against the reference grammar; the kernels never call this). */
int mrhyde_b200_expr_eval_host(int32_t n, const char* const* names, const char* const* exprs, const char* which,
                               int64_t npts, const double* vars7, double* out);
/* Generates the translation unit NVRTC compiles for this plan (the volume kernel with the plan's expressions,
 * tables and block size as constants), compiles it for sm_100a -- no device needed -- and optionally writes
 * the source / cubin to files (inspection with cuobjdump -sass).  log receives the compiler output. */
int mrhyde_b200_plan_debug_jit(mrhyde_b200_plan* plan, const char* source_path, const char* cubin_path, char* log, size_t log_cap);
/* Replays the metric ring of the sweep kernel on the host (element metrics by the kernel's formulas, then the plan's metric
 * source words) for plans that use it (plan_stat "metric_ring" > 0): checks the plan analysis against the oracle on
 * machines without a GPU.  All pointers are host pointers; no assemble_* entry point can reach this. */
int mrhyde_b200_plan_debug_metric_host(mrhyde_b200_plan* plan, const double* sol, const mrhyde_b200_time* t, int accumulate,
                                       double* res, double* jac_values);
/* The same for plans on the class ring (plan_stat "class_ring" > 0): one local-matrix value per class + the residual per element, by
 * the kernel's formulas, then the plan's scatter programs. */
int mrhyde_b200_plan_debug_class_host(mrhyde_b200_plan* plan, const double* sol, const mrhyde_b200_time* t, int accumulate,
                                      double* res, double* jac_values);
/* Replays the general path's kernel stage functions (general_kernel.cuh, compiled by the host compiler) and the pull on
 * the host for a host-only plan (device = -1) built with option kernel=general: a debugging aid that lets the kernel
 * logic be checked against the oracle on machines without a GPU.  Fails with ERR_STATE on device plans; the assemble
 * entry points never use it.  sol / res / jac_values (and t's vectors) are host buffers. */
/* Registers the host replay of the general kernel's stage functions.  It lives in the TEST-ONLY library libmrhyde_b200_emulate.so
 * (symbol mrhyde_b200_emulator_lookup), not in this one: without it the mrhyde_b200_plan_debug_emulate* calls fail with ERR_STATE. */
int mrhyde_b200_debug_set_emulator(void* lookup);
int mrhyde_b200_plan_debug_emulate(mrhyde_b200_plan* plan, const double* sol, const mrhyde_b200_time* t, int compute_jacobian,
                                   int compute_residual, double* res, double* jac_values);
/* Mass-matrix counterpart of mrhyde_b200_plan_debug_emulate (host-only plans; host buffers). */
int mrhyde_b200_plan_debug_emulate_mass(mrhyde_b200_plan* plan, const double* mass_wts, int lump, double* mass_values, double* diag);
int mrhyde_b200_plan_debug_emulate_apply_mass(mrhyde_b200_plan* plan, const double* mass_wts, const double* x, double* y);
int mrhyde_b200_plan_debug_emulate_initial(mrhyde_b200_plan* plan, double time, double* rhs);
/* Applies the plan's scatter programs on the host to caller-supplied staged element vectors
 * stage[n_elem][stage_len] (local Jacobian entries then residual entries), with the
 * same ordering and fixed-row rules as the device pull-scatter.  Verifies plan logic only. */
int mrhyde_b200_plan_debug_scatter_host(mrhyde_b200_plan* plan, const double* stage, int64_t stage_len, int accumulate,
                                        double* res, double* jac_values);
/* Layout of the staged element vector of a sweep-kernel plan (plan_stat "stage_len" doubles per element): kmap[i*ndof+j] =
 * index of local-matrix entry (i,j) -- the upper-triangle index, or the entry's class on plans with the class ring
 * (plan_stat "class_ring" > 0) -- and rmap[i] = index of residual entry i. */
int mrhyde_b200_plan_debug_stage_map(mrhyde_b200_plan* plan, int32_t* kmap, int32_t* rmap);
/* mask[r] = 1 for every row the sweep chains [chain_begin, chain_end) write.  Multi-rank plans number the chains that complete
 * ghost rows first (plan_stat "n_early_chains"): option "overlap halo" launches those, starts the exchange, then launches the rest. */
int mrhyde_b200_plan_debug_chain_rows(mrhyde_b200_plan* plan, int32_t chain_begin, int32_t chain_end, uint8_t* mask);

#ifdef __cplusplus
}
#endif
#endif /* MRHYDE_B200_H */
