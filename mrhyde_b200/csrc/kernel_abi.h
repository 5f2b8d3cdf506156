// Plain structs shared by the host plan code and the device kernels.  This header is compiled three ways:
// by g++ (plan.cpp, expr.cpp), by nvcc (the ahead-of-time kernels) and by NVRTC (the plan-specialised
// kernels, jit.cpp), so it uses nothing but fixed-width integers.
#pragma once
#ifdef __CUDACC_RTC__
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#else
#include <cstddef>
#include <cstdint>
#endif

namespace mrhyde_b200 {

// ---- sweep plan records (plan.hpp explains the scheme) ------------------------------------------------
struct StepRec {      // one level of one chain
  int32_t elem_begin, n_elem;     // into step_elems (global element ids, ascending); n_elem <= cap
  int32_t batch_begin, n_batches; // into batches: the rows that are complete after this step
};
// Rows of a step are grouped into BATCHES of up to 32 rows that share one gather pattern; a warp takes a batch with one
// row per lane, so every lane walks the same descriptor list and differs only in its ring anchor and output offset.
struct BatchRec {     // 16 bytes, read as one int4; batch b owns rows[32 b .. 32 b + n_rows)
  int32_t row_begin;  // = 32 * (index of this batch)
  int32_t desc_begin; // first slot descriptor of the pattern
  uint16_t n_rows;    // 1..32
  uint16_t n_slots;   // CSR entries of a row + 1: the last slot is the residual entry
  uint32_t flags;     // bit 0: strong-Dirichlet rows (isFixedDOF): no pattern, rows may differ in length
};
struct RowRec {       // 16 bytes, read as one int4; stored 32 per batch so that its address needs no batch header
  int32_t row;        // local row id (LID)
  uint16_t anchor;    // ring-slot element index the pattern offsets are relative to
  uint16_t aux;       // fixed rows: position of the diagonal entry inside the CSR row (0xFFFF if absent)
  int64_t base;       // row_map[row]: offset of the row in the CSR value array (copied from the graph at plan time)
};
struct PatternRec {   // host-side bookkeeping of the de-duplicated patterns
  int32_t desc_begin, n_slots;
};
// A slot descriptor lists up to 8 staged doubles (byte offsets into the ring relative to the row's anchor,
// 0xFFFFFFFF = unused, ascending element order); desc[parity] is valid when the current step writes ring slot `parity`.
constexpr uint32_t SRC_NONE = 0xFFFFFFFFu;
constexpr uint32_t BATCH_FIXED = 1u;
constexpr int SLOT_SRCS = 8;

struct SrcQuad { uint32_t x, y, z, w; };  // half a slot descriptor (read as uint4 on the device)

// METRIC ring (volume_kernel.cuh, MRH_JIT_METRIC): on parallelepiped cells with constant coefficients the local matrix is
// a fixed linear combination of reference tables, K_e = sum_g G_g(e) Stab[g], so a step stages only the element's scaled
// metric G (+ load vector and state) and the pull evaluates the combination per CSR entry.  The two ring slots are
// interleaved entry by entry -- entry m of element column c in slot s sits at double (m * 2 + s) * cap + c -- so a source
// is addressed by its element column alone.  A metric source word packs
//   bits  0..11  byte offset of the element column relative to the row's anchor: ((rel ^ parity) * cap + le - anchor) * 8
//   bits 12..17  upper-triangle index of the local entry (i, j)     bits 18..20  local column j     bits 21..23  local row i
// with the same slot / source order as the plain descriptors (mdesc[parity] parallels desc[parity]).
constexpr uint32_t MSRC_OFF_MASK = 0xFFFu;
constexpr int MSRC_T_SHIFT = 12, MSRC_J_SHIFT = 18, MSRC_I_SHIFT = 21;

struct ChainDev {
  const int32_t* chain_step_ptr;   // [n_chains+1]
  const StepRec* steps;
  const int32_t* step_elems;       // global element id of every step element (diagnostics)
  const int32_t* step_conn;        // [n_elem_with_halo][nverts]  connectivity in step order: the kernel reads element inputs
  const int32_t* step_lids;        // [n_elem_with_halo][ndof]    with one coalesced load level instead of id -> conn -> data
  const uint8_t* step_eclass;      // [n_elem_with_halo]
  const BatchRec* batches;
  const RowRec* rows;
  const SrcQuad* desc0;            // parity 0 / 1 descriptor tables, 2 SrcQuad per slot
  const SrcQuad* desc1;
  const SrcQuad* mdesc0;           // metric-ring source words (same indexing as desc0 / desc1)
  const SrcQuad* mdesc1;
  int32_t cap;                     // ring slot capacity (elements)
  int32_t chain_offset;            // first chain of this launch (a multi-rank assemble launches the ghost-row chains first)
  const uint8_t* chain_invariant;  // [n_chains] bit d: along the whole chain, thread t's element keeps the axis-d coordinates of its box
                                   // (vertex 0 and the +axis neighbour, bitwise): one-coordinate sub-expressions of axis d are step-invariant
                                   // bit 4 + d: all elements of a step share their axis-d interval: such sub-expressions are the same for the whole CTA
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
  int diag_one;      // overwrite mode: strong-Dirichlet rows get J(d,d) = 1 (setJacobianConstraints, only with `use strong DBCs`) or stay zero rows
};

// ---- in-kernel halo push (multi-rank plans on the p2p transport, halo.hpp) ------------------------------------------
// The chains that complete ghost rows (the first n_push_chains of the plan) write those rows a second time, straight into the owning
// rank's receive slab over NVLink (bulk copies for the matrix rows, plain stores for the residual entries), and the last of them to
// finish raises the owner's arrival flag: the exchange of Tpetra's Export(ADD) rides on the assembly kernel, and
// mrhyde_b200_halo_sum is left with waiting for the flag and adding the slab.
struct PushDev {
  double* remote_jac;                 // owner's slab, matrix part (already shifted to the 16-byte phase of the local array): index = CSR offset - ghost_base
  double* remote_res;                 // owner's slab, residual part: index = row - n_owned
  const double* res_base;             // this call's residual array (to recover the row from a residual pointer)
  unsigned long long* remote_arrive;  // owner's arrive[this rank]
  unsigned long long* remote_shift;   // owner's shift[this rank]: 0 | 1 doubles between the slab's matrix part and the first value
  const unsigned long long* local_ack;  // my ack[owner]: the call whose slab the owner has finished reading
  unsigned* counter;                  // push chains that have finished (reset by the last one)
  unsigned long long epoch;           // call counter of the halo sum this push belongs to (slab parity = epoch & 1)
  int64_t ghost_base;                 // row_map[n_owned]: CSR offset of the first ghost row
  int32_t n_owned, n_push_chains, shift, enabled;
};

// ---- flattened expression programs (expr.hpp) -------------------------------------------------------------
enum ExprOp : uint8_t {
  OP_END = 0,
  OP_PUSHC,   // push constant
  OP_PUSHV,   // push variable (index in c: 0 x, 1 y, 2 z, 3 t, 4.. extra inputs)
  OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW, OP_LT, OP_LTE, OP_GT, OP_GTE, OP_MAX, OP_MIN, OP_MEAN,   // binary: a = a op b
  OP_ADDC, OP_SUBC, OP_MULC, OP_DIVC, OP_POWC,        // binary with constant right operand
  OP_ADDV, OP_SUBV, OP_MULV, OP_DIVV,                 // binary with variable right operand
  OP_SIN, OP_COS, OP_TAN, OP_EXP, OP_LOG, OP_ABS, OP_SQRT, OP_SINH, OP_COSH,  // unary on top of stack
  // element reductions (functionManager_evaluate.hpp:413-460): the argument -- the instructions between OP_EBEGIN and the closing op, whose
  // constant holds the index of its OP_EBEGIN -- is evaluated at EVERY point of the element / side and the reduced value replaces the top of
  // the stack at all of them.  General path only (long programs).
  OP_EBEGIN, OP_EMAX, OP_EMIN, OP_EMEAN,
};

constexpr int EXPR_MAXOPS = 56;
constexpr int EXPR_MAXSTACK = 8;
constexpr int EXPR_NVARS = 10;  // x y z t n[x] n[y] n[z] + spare
constexpr int EXPR_STATE0 = 7;  // general path: variable EXPR_STATE0 + s is solution-field slot s (fields F[v][k], then their time derivatives)

struct ExprProgram {  // POD, copied into kernel parameters
  int32_t n = 0;
  int32_t is_const = 1;
  double cval = 0.0;
  uint8_t op[EXPR_MAXOPS] = {0};
  double c[EXPR_MAXOPS] = {0};
};

// ---- time integration coefficients (computeSolnTransientSeeded, workset.cpp:600-834) ------------------
constexpr int MAX_PREV = 4;   // BDF order <= 4 previous steps
constexpr int MAX_STAGE = 4;  // Butcher stages

struct TimeDev {
  int transient;
  int nprev, nstage_lo;          // previous steps used, stages below the current one
  double alpha_u, alpha_t;       // weights of the stage solution in the evaluation point: u = alpha_u s + ..., u_t = alpha_t s + ...
  double seed_u, seed_t;         // du/d(seeded dof), du_t/d(seeded dof): (alpha_u, alpha_t) when the stage solution is seeded, the chain-rule
                                 // factors of a previous step / stage otherwise (mrhyde_b200_time::seed_what)
  double one_minus_alpha_u;
  double timewt;                 // 1 / (dt b_s)
  double bdf[MAX_PREV + 1];      // BDF weights 1..nprev (index 0 unused)
  double stage_w[MAX_STAGE];     // A(s,s') / b(s')
  double time;                   // stage time
  const double* prev[MAX_PREV];
  const double* stg[MAX_STAGE];
  double deltat;                 // time step (1 when steady)
};

// ---- thermal, HGRAD Q1 ----------------------------------------------------------------------------------
template <int DIM>
struct Q1Shape {
  static constexpr int NV = 1 << DIM;            // vertices == HGRAD C1 dofs
  static constexpr int NQ = 1 << DIM;            // 2-point Gauss per direction
  static constexpr int NT = NV * (NV + 1) / 2;   // upper triangle of the local matrix
  static constexpr int NG = DIM * (DIM + 1) / 2; // symmetric metric tensor entries
  static constexpr int STAGE = NT + NV;          // staged doubles per element: J upper triangle, then residual
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
  //   K_ij = kappa |det| sum_{a<=b} G_ab Stab[ab][ij],  G = J^-1 J^-T   (ab order: 00 11 22 01 02 12)
  //   M_ij = rho cp |det| Mtab[ij],   load_i = f |det| Ltab[i] for a constant source
  double Stab[S::NG][S::NT];
  double Mtab[S::NT];
  double Ltab[S::NV];
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
  const uint8_t* eclass;  // [nelem] 0 general, 1 parallelepiped, 2 axis-aligned box
  const double* sol;
  ChainDev chains;
  GraphDev graph;
  OutDev out;
  PushDev push;
};

}  // namespace mrhyde_b200
