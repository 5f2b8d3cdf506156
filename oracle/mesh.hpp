// ORACLE (test infrastructure, never shipped, never on the product path).
//
// Synthetic inline rectangle / brick mesh, DOF numbering and CSR graph:
//   node / cell numbering  = SimpleMeshManager_Brick (nodes x fastest, Hex8 connectivity)
//                            src/tools/simplemeshmanager.hpp:1460-1512 (rectangle: same with k dropped)
//   DOF numbering          : Panzer's DOFManager is third-party and not visible, so this is a
//                            documented choice: per basis, entity-major with the variables that share
//                            the basis interleaved per entity in module order; bases concatenated in
//                            order of first appearance.  offsets(var,dof) follows the same interleaving
//                            (what getGIDFieldOffsets yields for a nodal field pattern).
//   CSR pattern            = union over elements of gids x gids, columns ascending per row
//                            src/interfaces/linear_algebra/linearAlgebraInterface_construct.hpp:213-265
//   strong-Dirichlet dofs  = side-closure dofs of each Dirichlet variable
//                            src/interfaces/discretization/discretizationInterface_dof.hpp:643-740
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "discretization.hpp"

namespace oracle {

struct BrickMesh {
  int dim = 3;
  int n[3] = {1, 1, 1};
  double lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
  bool periodic[3] = {false, false, false};  // "Periodic BCs" of the mesh sublist (panzer_stk periodic matchers): dof lattices wrap
  CellTopo topo;
  int num_nodes = 0, num_elems = 0;
  std::vector<double> nodes;  // (num_nodes, dim)
  std::vector<int> conn;      // (num_elems, nverts)
  std::vector<std::string> side_names;  // left right bottom top back front

  void build() {
    topo = make_topo(dim);
    if (dim == 2) n[2] = 0;
    const int nnx = n[0] + 1, nny = n[1] + 1, nnz = (dim == 3) ? n[2] + 1 : 1;
    num_nodes = nnx * nny * nnz;
    num_elems = n[0] * n[1] * (dim == 3 ? n[2] : 1);
    nodes.assign((size_t)num_nodes * dim, 0.0);
    const double dx = (hi[0] - lo[0]) / n[0], dy = (hi[1] - lo[1]) / n[1], dz = (dim == 3) ? (hi[2] - lo[2]) / n[2] : 0.0;
    int ct = 0;
    for (int k = 0; k < nnz; ++k)
      for (int j = 0; j < nny; ++j)
        for (int i = 0; i < nnx; ++i) {
          nodes[(size_t)ct * dim + 0] = lo[0] + i * dx;
          nodes[(size_t)ct * dim + 1] = lo[1] + j * dy;
          if (dim == 3) nodes[(size_t)ct * dim + 2] = lo[2] + k * dz;
          ++ct;
        }
    const int nv = topo.nverts;
    conn.assign((size_t)num_elems * nv, 0);
    const int nxy = nnx * nny;
    ct = 0;
    for (int k = 0; k < (dim == 3 ? n[2] : 1); ++k)
      for (int j = 0; j < n[1]; ++j)
        for (int i = 0; i < n[0]; ++i) {
          int* c = &conn[(size_t)ct * nv];
          c[0] = k * nxy + j * nnx + i;
          c[1] = k * nxy + j * nnx + (i + 1);
          c[2] = k * nxy + (j + 1) * nnx + (i + 1);
          c[3] = k * nxy + (j + 1) * nnx + i;
          if (dim == 3) for (int a = 0; a < 4; ++a) c[4 + a] = c[a] + nxy;
          ++ct;
        }
    side_names = {"left", "right", "bottom", "top"};
    if (dim == 3) { side_names.push_back("back"); side_names.push_back("front"); }
  }
  void elem_ijk(int e, int ijk[3]) const {
    ijk[0] = e % n[0]; ijk[1] = (e / n[0]) % n[1]; ijk[2] = (dim == 3) ? e / (n[0] * n[1]) : 0;
  }
  // sideset index (0..2*dim-1: -x +x -y +y -z +z) -> Shards local side id of the cell
  int local_side(int sideset) const {
    if (dim == 2) { static const int m[4] = {3, 1, 0, 2}; return m[sideset]; }
    static const int m[6] = {3, 1, 0, 2, 4, 5};
    return m[sideset];
  }
};

struct VarInfo {
  std::string name, btype;
  int order = 1;
  int basis = 0;  // index into unique basis list
  int rank = 0;   // rank among the variables sharing that basis
};

struct DofMap {
  std::vector<VarInfo> vars;
  std::vector<Basis> bases;
  std::vector<int> nvb;        // variables per basis
  std::vector<int64_t> gbase;  // first global dof of each basis block
  std::vector<int> lbase;      // first element-local dof of each basis block
  std::vector<std::vector<int>> offsets;  // [var][dof]
  int ndof_elem = 0;
  int64_t num_dofs = 0;
  std::vector<int> lids;       // (num_elems, ndof_elem)
  std::vector<int8_t> on_side; // (num_dofs, 2*dim) side-closure membership

  // entity lattice helpers ---------------------------------------------------------
  static int64_t hgrad_entities(const BrickMesh& m, int p) {
    int64_t c = 1;
    for (int d = 0; d < m.dim; ++d) c *= (int64_t)p * m.n[d] + (m.periodic[d] ? 0 : 1);
    return c;
  }

  void build(const BrickMesh& m, const std::vector<VarInfo>& vars_in) {
    vars = vars_in;
    bases.clear(); nvb.clear();
    for (auto& v : vars) {
      int found = -1;
      for (size_t b = 0; b < bases.size(); ++b) if (bases[b].type == v.btype && bases[b].order == v.order) found = (int)b;
      if (found < 0) { bases.push_back(make_basis(v.btype, v.order, m.dim)); nvb.push_back(0); found = (int)bases.size() - 1; }
      v.basis = found; v.rank = nvb[found]++;
    }
    gbase.assign(bases.size(), 0); lbase.assign(bases.size(), 0);
    const int nx = m.n[0], ny = m.n[1], nz = (m.dim == 3) ? m.n[2] : 0;
    // lattice points per direction: n+1, or n when the direction is periodic (the last plane is the first)
    const int px = nx + (m.periodic[0] ? 0 : 1), py = ny + (m.periodic[1] ? 0 : 1), pz = nz + (m.periodic[2] ? 0 : 1);
    std::vector<int64_t> nent(bases.size(), 0);
    int64_t g = 0; int l = 0;
    for (size_t b = 0; b < bases.size(); ++b) {
      const Basis& B = bases[b];
      if (B.type == "HGRAD") nent[b] = hgrad_entities(m, B.order);
      else if (B.type == "HVOL") nent[b] = m.num_elems;
      else if (B.type == "HCURL") nent[b] = (int64_t)nx * py * pz + (int64_t)px * ny * pz + (int64_t)px * py * nz;
      else if (B.type == "HDIV") nent[b] = (int64_t)px * ny * nz + (int64_t)nx * py * nz + (int64_t)nx * ny * pz;
      gbase[b] = g; lbase[b] = l;
      g += nent[b] * nvb[b]; l += B.card * nvb[b];
    }
    num_dofs = g; ndof_elem = l;
    offsets.assign(vars.size(), {});
    for (size_t v = 0; v < vars.size(); ++v) {
      const int b = vars[v].basis;
      offsets[v].resize(bases[b].card);
      for (int d = 0; d < bases[b].card; ++d) offsets[v][d] = lbase[b] + d * nvb[b] + vars[v].rank;
    }
    lids.assign((size_t)m.num_elems * ndof_elem, 0);
    on_side.assign((size_t)num_dofs * 2 * m.dim, 0);
    for (int e = 0; e < m.num_elems; ++e) {
      int eijk[3];
      m.elem_ijk(e, eijk);
      for (size_t b = 0; b < bases.size(); ++b) {
        const Basis& B = bases[b];
        for (int d = 0; d < B.card; ++d) {
          int64_t ent = 0;
          int8_t side[6] = {0, 0, 0, 0, 0, 0};
          if (B.type == "HGRAD") {
            int t[3];
            hgrad_ordinal_to_ijk(B, d, t);
            const int p = B.order;
            int64_t L[3], N[3];
            for (int c = 0; c < 3; ++c) {
              N[c] = (c < m.dim) ? (int64_t)p * m.n[c] + (m.periodic[c] ? 0 : 1) : 1;
              L[c] = (c < m.dim) ? ((int64_t)p * eijk[c] + t[c]) % N[c] : 0;
            }
            ent = L[0] + N[0] * (L[1] + N[1] * L[2]);
            for (int c = 0; c < m.dim; ++c) if (!m.periodic[c]) { side[2 * c] = (L[c] == 0); side[2 * c + 1] = (L[c] == N[c] - 1); }
          } else if (B.type == "HVOL") {
            ent = e;
          } else if (B.type == "HCURL") {
            const int64_t nxe = (int64_t)nx * py * pz, nye = (int64_t)px * ny * pz;
            int i, j, k;
            if (d < 4) {  // x-directed: d = j + 2k
              j = eijk[1] + (d % 2); k = eijk[2] + (d / 2); i = eijk[0];
              side[2] = (j == 0); side[3] = (j == ny); side[4] = (k == 0); side[5] = (k == nz);
              ent = i + (int64_t)nx * ((j % py) + (int64_t)py * (k % pz));
            } else if (d < 8) {  // y-directed: d-4 = i + 2k
              i = eijk[0] + ((d - 4) % 2); k = eijk[2] + ((d - 4) / 2); j = eijk[1];
              side[0] = (i == 0); side[1] = (i == nx); side[4] = (k == 0); side[5] = (k == nz);
              ent = nxe + (i % px) + (int64_t)px * (j + (int64_t)ny * (k % pz));
            } else {  // z-directed: d-8 = i + 2j
              i = eijk[0] + ((d - 8) % 2); j = eijk[1] + ((d - 8) / 2); k = eijk[2];
              side[0] = (i == 0); side[1] = (i == nx); side[2] = (j == 0); side[3] = (j == ny);
              ent = nxe + nye + (i % px) + (int64_t)px * ((j % py) + (int64_t)py * k);
            }
            for (int c = 0; c < 3; ++c) if (m.periodic[c]) side[2 * c] = side[2 * c + 1] = 0;
          } else if (B.type == "HDIV") {
            const int64_t nxf = (int64_t)px * ny * nz, nyf = (int64_t)nx * py * nz;
            const int dir = d / 2, s = d % 2;
            const int i = eijk[0] + (dir == 0 ? s : 0), j = eijk[1] + (dir == 1 ? s : 0), k = eijk[2] + (dir == 2 ? s : 0);
            if (dir == 0) { ent = (i % px) + (int64_t)px * (j + (int64_t)ny * k); side[0] = (i == 0); side[1] = (i == nx); }
            else if (dir == 1) { ent = nxf + i + (int64_t)nx * ((j % py) + (int64_t)py * k); side[2] = (j == 0); side[3] = (j == ny); }
            else { ent = nxf + nyf + i + (int64_t)nx * (j + (int64_t)ny * (k % pz)); side[4] = (k == 0); side[5] = (k == nz); }
            for (int c = 0; c < 3; ++c) if (m.periodic[c]) side[2 * c] = side[2 * c + 1] = 0;
          }
          for (int r = 0; r < nvb[b]; ++r) {
            const int64_t gid = gbase[b] + ent * nvb[b] + r;
            lids[(size_t)e * ndof_elem + lbase[b] + d * nvb[b] + r] = (int)gid;
            for (int s = 0; s < 2 * m.dim; ++s) if (side[s]) on_side[(size_t)gid * 2 * m.dim + s] = 1;
          }
        }
      }
    }
  }
  // global dof -> variable index (dofs of a basis block are interleaved by rank)
  int var_of_dof(int64_t gid) const {
    for (size_t b = 0; b < bases.size(); ++b) {
      const int64_t end = (b + 1 < bases.size()) ? gbase[b + 1] : num_dofs;
      if (gid >= gbase[b] && gid < end) {
        const int r = (int)((gid - gbase[b]) % nvb[b]);
        for (size_t v = 0; v < vars.size(); ++v) if (vars[v].basis == (int)b && vars[v].rank == r) return (int)v;
      }
    }
    return -1;
  }
};

struct CsrGraph {
  int64_t nrows = 0;
  std::vector<int64_t> rowptr;
  std::vector<int> colind;
  void build(int64_t nrows_, const std::vector<int>& lids, int num_elems, int ndof) {
    nrows = nrows_;
    std::vector<std::vector<int>> adj(nrows);  // row -> elements
    for (int e = 0; e < num_elems; ++e)
      for (int i = 0; i < ndof; ++i) adj[lids[(size_t)e * ndof + i]].push_back(e);
    rowptr.assign(nrows + 1, 0);
    colind.clear();
    std::vector<int> tmp;
    for (int64_t r = 0; r < nrows; ++r) {
      tmp.clear();
      for (int e : adj[r]) for (int j = 0; j < ndof; ++j) tmp.push_back(lids[(size_t)e * ndof + j]);
      std::sort(tmp.begin(), tmp.end());
      tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
      colind.insert(colind.end(), tmp.begin(), tmp.end());
      rowptr[r + 1] = (int64_t)colind.size();
    }
  }
};

}  // namespace oracle
