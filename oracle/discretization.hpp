// ORACLE (test infrastructure, never shipped, never on the product path).
//
// CPU restatement of the discretization services the hot path consumes
// (src/interfaces/discretization/*).  The arithmetic itself lives in Intrepid2/Shards
// (Trilinos, un-vendored, version unpinned by the reference: CMakeLists.txt:20,40-45), so
// this file restates the *published* definitions behind the reference's call sites:
//   basis factory  getBasis                discretizationInterface_basis.hpp:219-346
//     HGRAD deg 1  -> Basis_HGRAD_{QUAD,HEX}_C1_FEM  (Shards vertex order)
//     HGRAD deg>1  -> Basis_HGRAD_{QUAD,HEX}_Cn_FEM(deg, EQUISPACED)  (tensor order, x fastest)
//     HCURL / HDIV -> Basis_{HCURL,HDIV}_HEX_In_FEM(1, EQUISPACED)     (tensor order: x-, y-, z-directed)
//   quadrature     DefaultCubatureFactory(cellTopo, degree), degree default 2*maxorder
//                                          discretizationInterface_construct.hpp:103-112
//   reference tabulation setReferenceBasisData     _basis.hpp:12-133
//   push-forward   getPhysicalVolumetricBasis      _basis.hpp:382-611
//        J(c,p,i,j) = dx_i/dxi_j (CellTools::setJacobian), wts = |det J| w_ref (computeCellMeasure),
//        HGRAD grad  = J^{-T} grad_ref,  HCURL val = J^{-T} v_ref,  HCURL curl = J c_ref / det,
//        HDIV  val   = J v_ref / det,    HDIV div  = div_ref / det
//   volume ip/wts  getPhysicalIntegrationData      _integration.hpp:218-262
//   side ip/wts/normals getPhysicalBoundaryIntegrationData  _integration.hpp:592-774
// Conventions chosen where Intrepid2 is not visible (documented, "parity unpinned" at entry level):
//   * tensor Gauss points ordered x fastest; line Gauss points ascending;
//   * orientation: lowest-order HCURL/HDIV dofs carry a sign per (element, dof) supplied by the
//     mesh (all +1 on lexicographically numbered bricks, where every edge/face already points
//     from the lower to the higher global vertex id).
#pragma once
#include <array>
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

namespace oracle {

struct CellTopo {
  int dim = 0;
  int nverts = 0;
  std::vector<std::array<double, 3>> ref_verts;
  std::vector<std::vector<int>> side_nodes;  // Shards side -> vertex list
  std::vector<std::vector<int>> edge_nodes;  // Shards edge -> vertex pair (3-D only)
};

inline CellTopo make_topo(int dim) {
  CellTopo t;
  t.dim = dim;
  if (dim == 2) {  // shards::Quadrilateral<4>
    t.nverts = 4;
    t.ref_verts = {{-1, -1, 0}, {1, -1, 0}, {1, 1, 0}, {-1, 1, 0}};
    t.side_nodes = {{0, 1}, {1, 2}, {2, 3}, {3, 0}};
    t.edge_nodes = t.side_nodes;
  } else if (dim == 3) {  // shards::Hexahedron<8>
    t.nverts = 8;
    t.ref_verts = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};
    t.side_nodes = {{0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {0, 4, 7, 3}, {0, 3, 2, 1}, {4, 5, 6, 7}};
    t.edge_nodes = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {0, 4}, {1, 5}, {2, 6}, {3, 7}, {4, 5}, {5, 6}, {6, 7}, {7, 4}};
  } else {
    throw std::runtime_error("oracle: only 2-D quad and 3-D hex cells are restated");
  }
  return t;
}

// ---- Gauss-Legendre line rule exact for polynomial degree `degree` (ceil((degree+1)/2) points)
inline void gauss_line(int degree, std::vector<double>& x, std::vector<double>& w) {
  int n = (degree + 2) / 2;
  if (n < 1) n = 1;
  x.assign(n, 0.0); w.assign(n, 0.0);
  for (int i = 0; i < n; ++i) {  // Newton on P_n, roots ascending
    double z = -std::cos(M_PI * (i + 0.75) / (n + 0.5));
    double pp = 1.0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 0; j < n; ++j) { double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0); }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      double dz = p1 / pp;
      z -= dz;
      if (std::fabs(dz) < 1e-16) break;
    }
    {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 0; j < n; ++j) { double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0); }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
    }
    x[i] = z;
    w[i] = 2.0 / ((1.0 - z * z) * pp * pp);
  }
  for (int i = 0; i < n / 2; ++i) {  // symmetrise exactly
    double a = 0.5 * (x[n - 1 - i] - x[i]);
    x[i] = -a; x[n - 1 - i] = a;
    double b = 0.5 * (w[i] + w[n - 1 - i]);
    w[i] = b; w[n - 1 - i] = b;
  }
  if (n % 2 == 1) x[n / 2] = 0.0;
  if (n == 2) { x[0] = -1.0 / std::sqrt(3.0); x[1] = 1.0 / std::sqrt(3.0); w[0] = w[1] = 1.0; }
}

struct Cubature {
  int dim = 0, n = 0;
  std::vector<double> pts;  // (n, dim)
  std::vector<double> wts;  // (n)
};

inline Cubature tensor_gauss(int dim, int degree) {
  std::vector<double> x, w;
  gauss_line(degree, x, w);
  const int m = (int)x.size();
  Cubature c;
  c.dim = dim;
  c.n = 1;
  for (int d = 0; d < dim; ++d) c.n *= m;
  c.pts.assign((size_t)c.n * dim, 0.0);
  c.wts.assign(c.n, 0.0);
  for (int q = 0; q < c.n; ++q) {
    int r = q;
    double wt = 1.0;
    for (int d = 0; d < dim; ++d) { int i = r % m; r /= m; c.pts[(size_t)q * dim + d] = x[i]; wt *= w[i]; }
    c.wts[q] = wt;
  }
  return c;
}

// ---------------------------------------------------------------------------------------
// Reference bases
// ---------------------------------------------------------------------------------------
struct Basis {
  std::string type;  // HGRAD | HCURL | HDIV | HVOL
  int order = 1;
  int dim = 0;
  int card = 0;
  int vdim = 1;  // components of a basis value
};

inline Basis make_basis(const std::string& type, int order, int dim) {
  Basis b;
  b.type = type; b.order = order; b.dim = dim;
  if (type == "HGRAD") { b.card = 1; for (int d = 0; d < dim; ++d) b.card *= (order + 1); b.vdim = 1; }
  else if (type == "HVOL") { b.card = 1; b.vdim = 1; }
  else if (type == "HCURL" && dim == 3 && order == 1) { b.card = 12; b.vdim = 3; }
  else if (type == "HDIV" && dim == 3 && order == 1) { b.card = 6; b.vdim = 3; }
  else throw std::runtime_error("oracle: basis " + type + " order " + std::to_string(order) + " not restated");
  return b;
}

inline void lagrange1d(int p, double x, std::vector<double>& l, std::vector<double>& dl) {  // equispaced closed nodes
  l.assign(p + 1, 0.0); dl.assign(p + 1, 0.0);
  std::vector<double> xn(p + 1);
  for (int a = 0; a <= p; ++a) xn[a] = -1.0 + 2.0 * a / p;
  for (int a = 0; a <= p; ++a) {
    double v = 1.0;
    for (int b = 0; b <= p; ++b) if (b != a) v *= (x - xn[b]) / (xn[a] - xn[b]);
    l[a] = v;
    double dv = 0.0;
    for (int c = 0; c <= p; ++c) if (c != a) {
      double t = 1.0 / (xn[a] - xn[c]);
      for (int b = 0; b <= p; ++b) if (b != a && b != c) t *= (x - xn[b]) / (xn[a] - xn[b]);
      dv += t;
    }
    dl[a] = dv;
  }
}

// ordinal -> tensor index of the HGRAD basis (C1 uses the Shards vertex order, Cn the tensor order)
inline void hgrad_ordinal_to_ijk(const Basis& b, int ord, int ijk[3]) {
  ijk[0] = ijk[1] = ijk[2] = 0;
  if (b.order == 1) {
    static const int q4[4][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}};
    if (b.dim == 2) { ijk[0] = q4[ord][0]; ijk[1] = q4[ord][1]; }
    else { ijk[0] = q4[ord % 4][0]; ijk[1] = q4[ord % 4][1]; ijk[2] = ord / 4; }
  } else {
    const int n = b.order + 1;
    ijk[0] = ord % n; ijk[1] = (ord / n) % n; ijk[2] = (b.dim == 3) ? ord / (n * n) : 0;
  }
}

struct RefBasisTab {      // setReferenceBasisData (_basis.hpp:12-133): values at a point set
  int card = 0, npts = 0, vdim = 1, dim = 0;
  std::vector<double> val;   // (card, npts, vdim)
  std::vector<double> grad;  // (card, npts, dim)      HGRAD
  std::vector<double> curl;  // (card, npts, dim)      HCURL
  std::vector<double> div;   // (card, npts)           HDIV
};

inline RefBasisTab tabulate(const Basis& b, const double* pts, int npts) {
  RefBasisTab t;
  t.card = b.card; t.npts = npts; t.vdim = b.vdim; t.dim = b.dim;
  const int dim = b.dim;
  t.val.assign((size_t)b.card * npts * b.vdim, 0.0);
  if (b.type == "HGRAD") {
    t.grad.assign((size_t)b.card * npts * dim, 0.0);
    std::vector<double> l[3], dl[3];
    for (int p = 0; p < npts; ++p) {
      for (int d = 0; d < dim; ++d) lagrange1d(b.order, pts[(size_t)p * dim + d], l[d], dl[d]);
      for (int f = 0; f < b.card; ++f) {
        int ijk[3];
        hgrad_ordinal_to_ijk(b, f, ijk);
        double v = 1.0;
        for (int d = 0; d < dim; ++d) v *= l[d][ijk[d]];
        t.val[(size_t)f * npts + p] = v;
        for (int d = 0; d < dim; ++d) {
          double g = 1.0;
          for (int e = 0; e < dim; ++e) g *= (e == d) ? dl[e][ijk[e]] : l[e][ijk[e]];
          t.grad[((size_t)f * npts + p) * dim + d] = g;
        }
      }
    }
  } else if (b.type == "HVOL") {
    for (int p = 0; p < npts; ++p) t.val[p] = 1.0;
  } else if (b.type == "HCURL") {
    t.curl.assign((size_t)b.card * npts * 3, 0.0);
    for (int p = 0; p < npts; ++p) {
      const double x = pts[p * 3], y = pts[p * 3 + 1], z = pts[p * 3 + 2];
      const double lx[2] = {0.5 * (1 - x), 0.5 * (1 + x)}, ly[2] = {0.5 * (1 - y), 0.5 * (1 + y)}, lz[2] = {0.5 * (1 - z), 0.5 * (1 + z)};
      const double dlv[2] = {-0.5, 0.5};
      int f = 0;
      for (int k = 0; k < 2; ++k) for (int j = 0; j < 2; ++j, ++f) {  // x-directed: (ly_j lz_k, 0, 0)
        t.val[((size_t)f * npts + p) * 3 + 0] = ly[j] * lz[k];
        t.curl[((size_t)f * npts + p) * 3 + 1] = ly[j] * dlv[k];
        t.curl[((size_t)f * npts + p) * 3 + 2] = -dlv[j] * lz[k];
      }
      for (int k = 0; k < 2; ++k) for (int i = 0; i < 2; ++i, ++f) {  // y-directed: (0, lx_i lz_k, 0)
        t.val[((size_t)f * npts + p) * 3 + 1] = lx[i] * lz[k];
        t.curl[((size_t)f * npts + p) * 3 + 0] = -lx[i] * dlv[k];
        t.curl[((size_t)f * npts + p) * 3 + 2] = dlv[i] * lz[k];
      }
      for (int j = 0; j < 2; ++j) for (int i = 0; i < 2; ++i, ++f) {  // z-directed: (0, 0, lx_i ly_j)
        t.val[((size_t)f * npts + p) * 3 + 2] = lx[i] * ly[j];
        t.curl[((size_t)f * npts + p) * 3 + 0] = lx[i] * dlv[j];
        t.curl[((size_t)f * npts + p) * 3 + 1] = -dlv[i] * ly[j];
      }
    }
  } else if (b.type == "HDIV") {
    t.div.assign((size_t)b.card * npts, 0.0);
    for (int p = 0; p < npts; ++p) {
      const double c[3] = {pts[p * 3], pts[p * 3 + 1], pts[p * 3 + 2]};
      int f = 0;
      for (int d = 0; d < 3; ++d) for (int i = 0; i < 2; ++i, ++f) {  // d-normal: l_i(x_d) e_d
        t.val[((size_t)f * npts + p) * 3 + d] = (i == 0) ? 0.5 * (1 - c[d]) : 0.5 * (1 + c[d]);
        t.div[(size_t)f * npts + p] = (i == 0) ? -0.5 : 0.5;
      }
    }
  }
  return t;
}

// HGRAD C1 nodal shape functions of the cell geometry (always Quad4/Hex8: _basis.hpp:244-252)
inline void geom_shape(int dim, const double* xi, double* N, double* dN /*(nverts,dim)*/) {
  static const Basis q1_2 = make_basis("HGRAD", 1, 2), q1_3 = make_basis("HGRAD", 1, 3);
  RefBasisTab t = tabulate(dim == 2 ? q1_2 : q1_3, xi, 1);
  for (int f = 0; f < t.card; ++f) {
    N[f] = t.val[f];
    for (int d = 0; d < dim; ++d) dN[f * dim + d] = t.grad[(size_t)f * dim + d];
  }
}

inline double det_inv(int dim, const double* J, double* Ji) {  // CellTools::setJacobianDet / setJacobianInv
  if (dim == 2) {
    const double det = J[0] * J[3] - J[1] * J[2];
    Ji[0] = J[3] / det; Ji[1] = -J[1] / det; Ji[2] = -J[2] / det; Ji[3] = J[0] / det;
    return det;
  }
  const double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
  const double det = J[0] * c00 + J[1] * c01 + J[2] * c02;
  Ji[0] = c00 / det; Ji[1] = (J[2] * J[7] - J[1] * J[8]) / det; Ji[2] = (J[1] * J[5] - J[2] * J[4]) / det;
  Ji[3] = c01 / det; Ji[4] = (J[0] * J[8] - J[2] * J[6]) / det; Ji[5] = (J[2] * J[3] - J[0] * J[5]) / det;
  Ji[6] = c02 / det; Ji[7] = (J[1] * J[6] - J[0] * J[7]) / det; Ji[8] = (J[0] * J[4] - J[1] * J[3]) / det;
  return det;
}

// Physical data of one element at one reference point set
struct ElemGeom {
  std::vector<double> ip;    // (npts, dim)
  std::vector<double> J;     // (npts, dim, dim)
  std::vector<double> Jinv;  // (npts, dim, dim)
  std::vector<double> det;   // (npts)
};

inline void element_geometry(const CellTopo& topo, const double* nodes /*(nverts,dim)*/, const double* pts, int npts, ElemGeom& g) {
  const int dim = topo.dim, nv = topo.nverts;
  g.ip.assign((size_t)npts * dim, 0.0);
  g.J.assign((size_t)npts * dim * dim, 0.0);
  g.Jinv.assign((size_t)npts * dim * dim, 0.0);
  g.det.assign(npts, 0.0);
  double N[8], dN[24];
  for (int p = 0; p < npts; ++p) {
    geom_shape(dim, pts + (size_t)p * dim, N, dN);
    double* J = &g.J[(size_t)p * dim * dim];
    for (int n = 0; n < nv; ++n)
      for (int i = 0; i < dim; ++i) {
        g.ip[(size_t)p * dim + i] += nodes[n * dim + i] * N[n];
        for (int j = 0; j < dim; ++j) J[i * dim + j] += nodes[n * dim + i] * dN[n * dim + j];
      }
    g.det[p] = det_inv(dim, J, &g.Jinv[(size_t)p * dim * dim]);
  }
}

// Push one reference table forward to one element (values laid out (dof, pt, comp) like basis(elem,dof,pt,comp))
struct PhysBasis {
  std::vector<double> val, grad, curl, div;
};

inline void push_forward(const Basis& b, const RefBasisTab& r, const ElemGeom& g, const double* sign /*(card) or null*/, PhysBasis& out) {
  const int dim = b.dim, np = r.npts, nb = r.card;
  out.val.assign((size_t)nb * np * b.vdim, 0.0);
  out.grad.clear(); out.curl.clear(); out.div.clear();
  if (b.type == "HGRAD") {
    out.grad.assign((size_t)nb * np * dim, 0.0);
    out.val = r.val;
    for (int f = 0; f < nb; ++f)
      for (int p = 0; p < np; ++p) {
        const double* Ji = &g.Jinv[(size_t)p * dim * dim];
        for (int d = 0; d < dim; ++d) {
          double s = 0.0;
          for (int k = 0; k < dim; ++k) s += Ji[k * dim + d] * r.grad[((size_t)f * np + p) * dim + k];
          out.grad[((size_t)f * np + p) * dim + d] = s;
        }
      }
  } else if (b.type == "HVOL") {
    out.val = r.val;
  } else if (b.type == "HCURL") {
    out.curl.assign((size_t)nb * np * dim, 0.0);
    for (int f = 0; f < nb; ++f) {
      const double sg = sign ? sign[f] : 1.0;
      for (int p = 0; p < np; ++p) {
        const double* Ji = &g.Jinv[(size_t)p * dim * dim];
        const double* J = &g.J[(size_t)p * dim * dim];
        for (int d = 0; d < dim; ++d) {
          double s = 0.0, c = 0.0;
          for (int k = 0; k < dim; ++k) {
            s += Ji[k * dim + d] * r.val[((size_t)f * np + p) * dim + k];
            c += J[d * dim + k] * r.curl[((size_t)f * np + p) * dim + k];
          }
          out.val[((size_t)f * np + p) * dim + d] = sg * s;
          out.curl[((size_t)f * np + p) * dim + d] = sg * c / g.det[p];
        }
      }
    }
  } else if (b.type == "HDIV") {
    out.div.assign((size_t)nb * np, 0.0);
    for (int f = 0; f < nb; ++f) {
      const double sg = sign ? sign[f] : 1.0;
      for (int p = 0; p < np; ++p) {
        const double* J = &g.J[(size_t)p * dim * dim];
        for (int d = 0; d < dim; ++d) {
          double s = 0.0;
          for (int k = 0; k < dim; ++k) s += J[d * dim + k] * r.val[((size_t)f * np + p) * dim + k];
          out.val[((size_t)f * np + p) * dim + d] = sg * s / g.det[p];
        }
        out.div[(size_t)f * np + p] = sg * r.div[(size_t)f * np + p] / g.det[p];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Side quadrature: reference side rule mapped into the cell through the Shards subcell
// parametrisation (CellTools::mapToReferenceSubcell), reference side tangents
// (getReferenceEdgeTangent / getReferenceFaceTangents)
// ---------------------------------------------------------------------------------------
struct SideRule {
  int n = 0;
  std::vector<double> pts;   // (n, dim) in the cell's reference frame
  std::vector<double> wts;   // (n)
  double tanU[3] = {0, 0, 0}, tanV[3] = {0, 0, 0};
};

inline SideRule make_side_rule(const CellTopo& topo, int side, int degree) {
  SideRule s;
  const int dim = topo.dim;
  Cubature c = tensor_gauss(dim - 1, degree);
  s.n = c.n;
  s.wts = c.wts;
  s.pts.assign((size_t)c.n * dim, 0.0);
  const auto& sn = topo.side_nodes[side];
  if (dim == 2) {
    const auto& a = topo.ref_verts[sn[0]];
    const auto& b = topo.ref_verts[sn[1]];
    for (int d = 0; d < 2; ++d) s.tanU[d] = 0.5 * (b[d] - a[d]);
    for (int q = 0; q < c.n; ++q) {
      const double u = c.pts[q];
      for (int d = 0; d < 2; ++d) s.pts[(size_t)q * 2 + d] = 0.5 * (1 - u) * a[d] + 0.5 * (1 + u) * b[d];
    }
  } else {
    const auto& a = topo.ref_verts[sn[0]];
    const auto& b = topo.ref_verts[sn[1]];
    const auto& cc = topo.ref_verts[sn[2]];
    const auto& d4 = topo.ref_verts[sn[3]];
    for (int d = 0; d < 3; ++d) {
      s.tanU[d] = 0.25 * (-a[d] + b[d] + cc[d] - d4[d]);
      s.tanV[d] = 0.25 * (-a[d] - b[d] + cc[d] + d4[d]);
    }
    for (int q = 0; q < c.n; ++q) {
      const double u = c.pts[q * 2], v = c.pts[q * 2 + 1];
      for (int d = 0; d < 3; ++d)
        s.pts[(size_t)q * 3 + d] = 0.25 * ((1 - u) * (1 - v) * a[d] + (1 + u) * (1 - v) * b[d] + (1 + u) * (1 + v) * cc[d] + (1 - u) * (1 + v) * d4[d]);
    }
  }
  return s;
}

// side weights and unit normals of one element (getPhysicalBoundaryIntegrationData :652-774)
inline void side_measure(const CellTopo& topo, const SideRule& s, const ElemGeom& g, std::vector<double>& wts, std::vector<double>& normals /*(n,dim)*/) {
  const int dim = topo.dim;
  wts.assign(s.n, 0.0);
  normals.assign((size_t)s.n * dim, 0.0);
  for (int p = 0; p < s.n; ++p) {
    const double* J = &g.J[(size_t)p * dim * dim];
    double nrm[3] = {0, 0, 0};
    if (dim == 2) {
      double t[2] = {0, 0};
      for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) t[i] += J[i * 2 + j] * s.tanU[j];
      nrm[0] = t[1]; nrm[1] = -t[0];  // rotation [[0,1],[-1,0]]
      wts[p] = std::sqrt(t[0] * t[0] + t[1] * t[1]) * s.wts[p];
    } else {
      double tu[3] = {0, 0, 0}, tv[3] = {0, 0, 0};
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { tu[i] += J[i * 3 + j] * s.tanU[j]; tv[i] += J[i * 3 + j] * s.tanV[j]; }
      nrm[0] = tu[1] * tv[2] - tu[2] * tv[1];
      nrm[1] = tu[2] * tv[0] - tu[0] * tv[2];
      nrm[2] = tu[0] * tv[1] - tu[1] * tv[0];
      wts[p] = std::sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]) * s.wts[p];
    }
    double len = 0.0;
    for (int d = 0; d < dim; ++d) len += nrm[d] * nrm[d];
    len = std::sqrt(len);
    for (int d = 0; d < dim; ++d) normals[(size_t)p * dim + d] = nrm[d] * (1.0 / len);
  }
}

}  // namespace oracle
