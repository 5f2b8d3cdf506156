// Boundary-group kernels of the thermal module (see boundary.cuh for the reference map).
// One thread per (element, side) item; items are coloured at plan time so that the items of one
// launch never share a row, which makes the plain `+=` into res / J race-free and the summation
// order (colour by colour) fixed from run to run.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <stdexcept>

#include "boundary.cuh"
#include "volume_kernel.cuh"

namespace mrhyde_b200 {

struct BoundaryParams {
  const int32_t* item_elem;
  const int32_t* item_group;
  const BoundaryGroupDev* groups;
  int32_t first, count;
  double formparam;
  int32_t adjoint;   // useadjoint: local matrices are filled transposed and sf = 1 (thermal.cpp:197-201, updateJacBoundary :1047-1062)
  ExprProgram diffusion;
  TimeDev td;
  const double* vx; const double* vy; const double* vz;
  const int32_t* conn; const int32_t* lids;
  const double* sol;
  GraphDev graph;
  OutDev out;
};

template <int DIM>
__global__ void __launch_bounds__(64) thermal_boundary_kernel(const __grid_constant__ BoundaryParams P) {
  constexpr int NV = 1 << DIM;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P.count) return;
  const int e = P.item_elem[P.first + k];
  const BoundaryGroupDev& g = P.groups[P.item_group[P.first + k]];
  int cn[NV], ld[NV];
  double X[NV][DIM], u[NV], ut[NV];
  const double* vc[3] = {P.vx, P.vy, P.vz};
  for (int n = 0; n < NV; ++n) {
    cn[n] = P.conn[(size_t)e * NV + n];
    ld[n] = P.lids[(size_t)e * NV + n];
    for (int d = 0; d < DIM; ++d) X[n][d] = vc[d][cn[n]];
    gather_dof(P.sol, P.td, ld[n], u[n], ut[n]);
  }
  // ---- side geometry: weights, unit normals, points (getPhysicalBoundaryIntegrationData)
  double ws[BND_MAXQ], nrm[BND_MAXQ][3], xq[BND_MAXQ][3], Ji[BND_MAXQ][DIM][DIM];
  double vol = 0.0;
  for (int q = 0; q < g.nqp; ++q) {
    double J[DIM][DIM];
    for (int d = 0; d < 3; ++d) { xq[q][d] = 0.0; nrm[q][d] = 0.0; }
    for (int d = 0; d < DIM; ++d) {
      for (int a = 0; a < DIM; ++a) J[d][a] = 0.0;
      for (int n = 0; n < NV; ++n) {
        xq[q][d] += g.gN[q][n] * X[n][d];
        for (int a = 0; a < DIM; ++a) J[d][a] += X[n][d] * g.gdN[q][n][a];
      }
    }
    double len;
    if constexpr (DIM == 2) {
      double t[2] = {0, 0};
      for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) t[i] += J[i][j] * g.tu[j];
      nrm[q][0] = t[1]; nrm[q][1] = -t[0];
      len = sqrt(t[0] * t[0] + t[1] * t[1]);
    } else {
      double tu[3] = {0, 0, 0}, tv[3] = {0, 0, 0};
      for (int i = 0; i < DIM; ++i) for (int j = 0; j < DIM; ++j) { tu[i] += J[i][j] * g.tu[j]; tv[i] += J[i][j] * g.tv[j]; }
      nrm[q][0] = tu[1] * tv[2] - tu[2] * tv[1];
      nrm[q][1] = tu[2] * tv[0] - tu[0] * tv[2];
      nrm[q][2] = tu[0] * tv[1] - tu[1] * tv[0];
      len = sqrt(nrm[q][0] * nrm[q][0] + nrm[q][1] * nrm[q][1] + nrm[q][2] * nrm[q][2]);
    }
    ws[q] = len * g.wts[q];
    const double il = 1.0 / len;
    for (int d = 0; d < DIM; ++d) nrm[q][d] *= il;
    vol += ws[q];
    det_inverse<DIM>(J, Ji[q]);  // for the HGRAD gradient push-forward
  }
  const double h = pow(vol, 1.0 / ((double)DIM - 1.0));  // Workset::getSideElementSize, workset.cpp:2719-2733

  double r[NV], K[NV][NV];
  for (int i = 0; i < NV; ++i) { r[i] = 0.0; for (int j = 0; j < NV; ++j) K[i][j] = 0.0; }
  const double epen = 10.0, sf = P.adjoint ? 1.0 : P.formparam;
  for (int q = 0; q < g.nqp; ++q) {
    ExprVars in;
    in.v[0] = xq[q][0]; in.v[1] = xq[q][1]; in.v[2] = xq[q][2]; in.v[3] = P.td.time;
    in.v[4] = nrm[q][0]; in.v[5] = nrm[q][1]; in.v[6] = nrm[q][2];
    const double data = expr_eval(g.data, in);
    if (g.bctype == 1) {  // Neumann: res_i += -g w phi_i
      for (int i = 0; i < NV; ++i) r[i] += -data * ws[q] * g.phi[q][i];
      continue;
    }
    const double kap = expr_eval(P.diffusion, in);
    double gn[NV];  // grad(phi_i) . n
    double T = 0.0, dTn = 0.0;
    for (int i = 0; i < NV; ++i) {
      double s = 0.0;
      for (int d = 0; d < DIM; ++d) {
        double gd = 0.0;
        for (int a = 0; a < DIM; ++a) gd += Ji[q][a][d] * g.dphi[q][i][a];
        s += gd * nrm[q][d];
      }
      gn[i] = s;
      T += u[i] * g.phi[q][i];
      dTn += u[i] * s;
    }
    const double w = ws[q];
    for (int i = 0; i < NV; ++i) {
      const double ph = g.phi[q][i];
      r[i] += epen / h * kap * (T - data) * w * ph;
      r[i] += -kap * dTn * w * ph;
      r[i] += -sf * kap * (T - data) * w * gn[i];
      for (int j = 0; j < NV; ++j)
        K[i][j] += P.td.seed_u * (epen / h * kap * g.phi[q][j] * w * ph - kap * gn[j] * w * ph - sf * kap * g.phi[q][j] * w * gn[i]);
    }
  }
  // ---- scatter (items of one launch never share a row)
  for (int i = 0; i < NV; ++i) {
    const int row = ld[i];
    if (P.graph.fixed[row]) continue;
    if (P.out.res) P.out.res[row] += -r[i];
    if (P.out.jac && g.bctype == 2) {
      const int64_t rs = P.graph.rowptr[row], re = P.graph.rowptr[row + 1];
      for (int j = 0; j < NV; ++j) {
        int64_t lo = rs, hi = re - 1;
        const int col = ld[j];
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (P.graph.colind[mid] < col) lo = mid + 1; else hi = mid; }
        P.out.jac[lo] += P.adjoint ? K[j][i] : K[i][j];
      }
    }
  }
}

BoundaryPlan::~BoundaryPlan() {
  if (d_item_elem) cudaFree(d_item_elem);
  if (d_item_group) cudaFree(d_item_group);
  if (d_groups) cudaFree(d_groups);
}

static void shape_at(int dim, const double* xi, double* N, double* dN) {
  static const double sg[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};
  const int nv = 1 << dim;
  for (int n = 0; n < nv; ++n) {
    double f[3] = {1, 1, 1}, df[3] = {0, 0, 0};
    for (int d = 0; d < dim; ++d) { f[d] = 0.5 * (1.0 + sg[n][d] * xi[d]); df[d] = 0.5 * sg[n][d]; }
    N[n] = f[0] * f[1] * (dim == 3 ? f[2] : 1.0);
    for (int d = 0; d < dim; ++d) {
      double g = df[d];
      for (int o = 0; o < dim; ++o) if (o != d) g *= f[o];
      dN[n * 3 + d] = g;
    }
  }
}

void build_boundary_plan(const BoundarySetup& bs, const std::vector<BoundaryGroupHost>& groups, const MeshGraph& m, BoundaryPlan& out, size_t* dev_bytes) {
  out.dim = bs.dim; out.formparam = bs.formparam; out.diffusion = bs.diffusion;
  const int nv = bs.nv, dim = bs.dim;
  std::vector<BoundaryGroupDev> gd;
  std::vector<int32_t> item_elem, item_group;
  for (const auto& g : groups) {
    int code = 0;
    if (g.bctype == "Neumann") code = 1;
    else if (g.bctype == "weak Dirichlet") code = 2;
    if (code == 0 || g.elem_ids.empty()) continue;  // strong Dirichlet / none: boundaryResidual adds nothing
    if (g.nqp > BND_MAXQ) throw std::runtime_error("boundary: side cubature has more points than this build supports");
    if (code == 2 && g.grad.empty()) throw std::runtime_error("boundary: weak Dirichlet needs side-tabulated basis gradients");
    BoundaryGroupDev D;
    std::memset(&D, 0, sizeof(D));
    D.bctype = code; D.nqp = g.nqp;
    for (int d = 0; d < 3; ++d) { D.tu[d] = g.tu[d]; D.tv[d] = g.tv[d]; }
    for (int q = 0; q < g.nqp; ++q) {
      D.wts[q] = g.wts[(size_t)q];
      double N[8], dN[24];
      shape_at(dim, &g.pts[(size_t)q * dim], N, dN);
      for (int n = 0; n < nv; ++n) {
        D.gN[q][n] = N[n];
        D.phi[q][n] = g.val[(size_t)n * g.nqp + q];
        for (int d = 0; d < dim; ++d) {
          D.gdN[q][n][d] = dN[n * 3 + d];
          if (!g.grad.empty()) D.dphi[q][n][d] = g.grad[((size_t)n * g.nqp + q) * dim + d];
        }
      }
    }
    D.data = g.data;
    for (int32_t e : g.elem_ids) {
      if (e < 0 || e >= m.nelem) throw std::runtime_error("boundary: element id out of range");
      item_elem.push_back(e); item_group.push_back((int32_t)gd.size());
    }
    gd.push_back(D);
  }
  out.groups.clear();
  if (item_elem.empty()) return;
  // greedy colouring on shared rows
  const size_t ni = item_elem.size();
  std::vector<uint64_t> used((size_t)m.nrows, 0);
  std::vector<int> colour(ni, 0);
  int ncol = 0;
  for (size_t k = 0; k < ni; ++k) {
    uint64_t mask = 0;
    for (int i = 0; i < m.ndof; ++i) mask |= used[(size_t)m.lids[(size_t)item_elem[k] * m.ndof + i]];
    int c = 0;
    while (c < 63 && (mask >> c) & 1) ++c;
    if (c >= 63) throw std::runtime_error("boundary: colouring needs more than 63 colours");
    colour[k] = c; ncol = std::max(ncol, c + 1);
    for (int i = 0; i < m.ndof; ++i) used[(size_t)m.lids[(size_t)item_elem[k] * m.ndof + i]] |= (1ull << c);
  }
  std::vector<int32_t> se, sg;
  for (int c = 0; c < ncol; ++c) {
    BoundaryColour bc;
    bc.first = (int32_t)se.size();
    for (size_t k = 0; k < ni; ++k) if (colour[k] == c) { se.push_back(item_elem[k]); sg.push_back(item_group[k]); }
    bc.count = (int32_t)se.size() - bc.first;
    out.groups.push_back(bc);
  }
  auto up = [&](auto** dst, const auto& h) {
    typedef typename std::remove_reference<decltype(h[0])>::type T;
    if (cudaMalloc((void**)dst, h.size() * sizeof(T)) != cudaSuccess) throw std::runtime_error("boundary: cudaMalloc failed");
    if (cudaMemcpy(*dst, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) throw std::runtime_error("boundary: cudaMemcpy failed");
    if (dev_bytes) *dev_bytes += h.size() * sizeof(T);
  };
  up(&out.d_item_elem, se); up(&out.d_item_group, sg); up(&out.d_groups, gd);
}

void launch_boundary(const BoundaryPlan& B, const double* sol, const TimeDev& td, const GraphDev& G, const OutDev& O, void* stream, bool adjoint) {
  BoundaryParams P;
  P.adjoint = adjoint ? 1 : 0;
  P.item_elem = B.d_item_elem; P.item_group = B.d_item_group; P.groups = B.d_groups;
  P.formparam = B.formparam; P.diffusion = B.diffusion; P.td = td;
  P.vx = B.vx; P.vy = B.vy; P.vz = B.vz; P.conn = B.conn; P.lids = B.lids; P.sol = sol; P.graph = G; P.out = O;
  for (const auto& c : B.groups) {
    P.first = c.first; P.count = c.count;
    const int blocks = (c.count + 63) / 64;
    if (B.dim == 3) thermal_boundary_kernel<3><<<blocks, 64, 0, (cudaStream_t)stream>>>(P);
    else thermal_boundary_kernel<2><<<blocks, 64, 0, (cudaStream_t)stream>>>(P);
  }
}

}  // namespace mrhyde_b200
