// Plan-time (host) analysis for the owner-computes assembly kernels.
//
// The reference scatters each element's local matrix into the CSR matrix with a per-entry linear
// search and atomics (assemblyManager_scatter.hpp:162-278).  Here the scatter is turned around at
// plan time: rows are grouped into spatially compact PATCHES; one CTA computes every element
// that touches its rows (elements on patch borders are recomputed by the neighbouring patches),
// stages the local matrices in shared memory, and then each CSR slot of an owned row sums its
// contributions in ascending element order -- the serial reference's order -- and is
// written exactly once with a plain store.  No atomics, no colour passes, run-to-run reproducible.
//
// The per-slot contribution lists ("scatter program") are identical for all interior patches of a
// structured mesh, so programs are de-duplicated into TEMPLATES that stay L2-resident.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

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
  std::vector<uint8_t> affine;     // [nelem] 1 = constant Jacobian (parallelepiped) within 2e-14
  void classify_affine();
  // de-duplicates per-element coordinates into the vertex table + connectivity
  void set_elem_nodes(int64_t n_elem, const double* elem_nodes);
};

constexpr uint16_t SLOT_RES = 0xFFFF;  // slot_k value of the residual slot of a row

struct TemplateHeader {  // one scatter program
  int32_t n_pe;        // elements of the patch, halo included
  int32_t n_rows;      // owned rows
  int32_t n_slots;     // sum(rowlen + 1) over owned rows
  int32_t pad;
  int64_t off_slot;    // offset into slot_row / slot_k / cptr (cptr has n_slots + 1 entries per template)
  int64_t off_cptr;
  int64_t off_csrc;    // offset into csrc
};

struct PatchPlan {
  int32_t n_patches = 0;
  int32_t chunk = 0;                 // target elements per patch before the halo
  int32_t max_pe = 0, max_slots = 0, max_rows = 0;
  int64_t n_elem_with_halo = 0;
  std::vector<int32_t> patch_elem_ptr, patch_elems;  // [n_patches+1], global element ids
  std::vector<int32_t> patch_row_ptr, patch_rows;    // [n_patches+1], owned rows (ascending)
  std::vector<int32_t> patch_tmpl;                   // [n_patches]
  std::vector<TemplateHeader> tmpl;
  std::vector<uint16_t> slot_row, slot_k;
  std::vector<uint32_t> cptr;
  std::vector<uint16_t> csrc;                        // staged index = entry * n_pe + local element
  std::vector<int32_t> orphan_rows;                  // rows no element touches
};

// kmap[i*ndof + j] = index of local-matrix entry (i,j) in the staged per-element vector,
// rmap[i] = index of residual entry i; stage_len = staged doubles per element.
void build_patch_plan(const MeshGraph& m, const std::vector<uint16_t>& kmap, const std::vector<uint16_t>& rmap,
                      int stage_len, int chunk_target, size_t smem_budget_bytes, PatchPlan& out);

}  // namespace mrhyde_b200
