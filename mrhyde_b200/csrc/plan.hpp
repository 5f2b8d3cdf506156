// Plan-time (host) analysis for the owner-computes, sweep-ordered assembly kernels.
//
// The reference scatters each element's local matrix into the CSR matrix with a per-entry linear
// search and atomics (assemblyManager_scatter.hpp:162-278), one small group of elements at a time.
// Here the scatter is turned around at plan time:
//   * elements are layered by a breadth-first sweep (two elements that share a dof are at most one
//     LEVEL apart) and cut into COLUMNS across the sweep; a CHAIN is one column over a range of levels;
//   * one CTA walks a chain level by level (a STEP): it computes the local matrices of the step's
//     elements -- its own column plus the one-element halo ring around it -- into a two-slot ring
//     buffer in shared memory, then every CSR slot (and residual entry) of the rows that became
//     complete sums its contributions, which by construction sit in the current or the previous
//     ring slot, in ascending element order -- the serial reference's order -- and
//     is written exactly once with a plain store.  No atomics, no colour passes, reproducible.
//   * the per-row gather lists are stored as PATTERNS relative to a per-row anchor, so a structured
//     mesh needs a handful of patterns that stay cache resident.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "kernel_abi.h"

namespace mrhyde_b200 {

struct MeshGraph {  // host inputs of one block
  int dim = 3, nverts = 8, ndof = 8;
  int64_t nelem = 0, nvert = 0, nrows = 0, nowned = 0, nnz = 0;
  std::vector<double> vcoord[3];   // SoA vertex coordinates
  std::vector<int32_t> conn;       // [nelem][nverts]
  std::vector<int32_t> lids;       // [nelem][ndof]
  std::vector<int8_t> orient;      // [nelem][ndof] or empty
  std::vector<int64_t> rowptr;     // [nrows+1]
  std::vector<int32_t> colind;     // [nnz]
  std::vector<uint8_t> fixed;      // [nrows]
  std::vector<uint8_t> eclass;     // [nelem] 0 = general cell, 1 = parallelepiped (constant Jacobian within 2e-14),
                                   //         2 = parallelepiped whose Jacobian is diagonal (axis-aligned box)
  void classify_cells();
  // de-duplicates per-element coordinates into the vertex table + connectivity
  void set_elem_nodes(int64_t n_elem, const double* elem_nodes);
};

struct ChainPlan {
  int32_t n_chains = 0, n_levels = 0, n_columns = 0, n_segments = 0;
  int32_t n_early_chains = 0;  // chains [0, n_early_chains) complete every ghost row of a multi-rank plan (0: none / single rank)
  int32_t cap = 0;             // ring slot capacity in elements (max elements of any step)
  int32_t stage_len = 0;       // staged doubles per element
  int32_t max_rows_step = 0, max_batches_step = 0;
  int64_t n_elem_with_halo = 0;
  std::vector<int32_t> chain_step_ptr;   // [n_chains+1]
  std::vector<StepRec> steps;
  std::vector<int32_t> step_elems;
  std::vector<int32_t> step_conn, step_lids;   // element inputs in step order
  std::vector<uint8_t> step_eclass;
  std::vector<BatchRec> batches;
  std::vector<RowRec> rows;
  std::vector<PatternRec> patterns;
  std::vector<uint32_t> desc[2];         // SLOT_SRCS per slot
  std::vector<uint32_t> mdesc[2];        // metric-ring source words, parallel to desc (kernel_abi.h); empty when ndof > 8
  std::vector<uint8_t> chain_invariant;  // [n_chains] bit d: every step of the chain has the same element count and element t of every step spans the
                                         // same axis-d interval (bitwise equal coordinates of vertex 0 and its +d neighbour) as element t of the first step
                                         // bit 4 + d: within every step of the chain all elements span one and the same axis-d interval
  std::vector<int32_t> orphan_rows;      // rows no element touches
  std::vector<int32_t> ghost_patterns;   // desc_begin of the patterns of batches that hold ghost rows (multi-rank plans): they always get
                                         // generated pull code, so that the in-kernel halo push covers every ghost row
  int64_t slot_bytes() const { return (int64_t)cap * stage_len * 8; }
};

struct ChainOptions {
  int sweep_axis = -1;          // -1: last axis
  int column_elems = 128;       // target owned elements per level and column
  int min_chains = 592;         // aim for at least this many chains (4 per SM) by cutting the sweep into segments
  int min_segment_levels = 8;
  int cta_slots = 0;            // CTAs resident on the device at once (SMs x CTAs per SM): the chain count is tuned to fill whole waves;
                                // 0: derived from the fields below once the ring capacity is known
  int n_sm = 148;               // multiprocessors of the device
  int max_blocks_per_sm = 3;    // register-limited CTAs per SM of the volume kernel (128 registers x 160 threads)
  int ring_stage_len = 0;       // doubles per element in the ring of the build that will run (0: stage_len)
  size_t warp_buffer_bytes = 1280;  // per-warp transpose / row buffer
  size_t smem_budget = 108 * 1024;  // ring (2 slots) must fit here; the row table (24 B per row of a step) sits behind it
};

// kmap[i*ndof + j] = index of local-matrix entry (i,j) in the staged per-element vector,
// rmap[i] = index of residual entry i; stage_len = staged doubles per element.
void build_chain_plan(const MeshGraph& m, const std::vector<uint16_t>& kmap, const std::vector<uint16_t>& rmap,
                      int stage_len, const ChainOptions& opt, ChainPlan& out);

// Applies the plan on the host to caller-supplied staged element vectors stage[nelem][stage_len], walking
// chains, ring slots, patterns and lane items exactly like the device pull phase (plan verification).
void host_apply_chain_plan(const MeshGraph& m, const ChainPlan& cp, const double* stage, bool accumulate, double* res, double* jac);

// Host replay of the METRIC pull (kernel_abi.h): `metric` holds, per element, [ng scaled metric entries | md | load[nd] | u[nd] | ut[nd]]
// (stride ng + 1 + 3 nd); Stab[ng][nt], Mtab[nt] are the reference tables.  Walks chains, interleaved ring slots and mdesc words
// exactly like the device code:  J_k = alpha_u sum G Stab + alpha_t sum md Mtab,  res = -(sum_k (K_k u_k + M_k ut_k) - sum load).
void host_apply_metric_plan(const MeshGraph& m, const ChainPlan& cp, const double* metric, int ng, const double* Stab, const double* Mtab,
                            double alpha_u, double alpha_t, bool accumulate, double* res, double* jac);

}  // namespace mrhyde_b200
