// A host program in the reference's language (C++) that drives the C ABI directly -- no Python, no torch: the calls a MrHyDE
// AssemblyManager would make for one thermal quad-Q1 block (INTEGRATION.md shows the same calls inside the reference's classes).
//
//   g++ -std=c++17 -I include examples/host_assemble.cpp -L mrhyde_b200 -lmrhyde_b200 -Wl,-rpath,$PWD/mrhyde_b200 -o host_assemble
//   ./host_assemble [nx] [device]        device = -1 (default): host-only analysis plan, no GPU needed; >= 0: assemble on that GPU
//
// It builds an nx x nx quad mesh of the unit square, tabulates the Q1 basis at the 2 x 2 Gauss points (what
// DiscretizationInterface::setReferenceBasisData stores), hands graph, mesh, functions and boundary conditions to the plan and,
// on a GPU, assembles residual and Jacobian of the reference's 2D_verification problem through mrhyde_b200_assemble_jacres_host.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>

#include "mrhyde_b200.h"

static void check(int rc, const char* what) {
  if (rc != MRHYDE_B200_OK) {
    std::fprintf(stderr, "%s failed (%d): %s\n", what, rc, mrhyde_b200_last_error());
    std::exit(1);
  }
}

int main(int argc, char** argv) {
  const int nx = argc > 1 ? std::atoi(argv[1]) : 8;
  const int device = argc > 2 ? std::atoi(argv[2]) : -1;
  const int nn = nx + 1, n_nodes = nn * nn, n_elem = nx * nx;

  // ---- reference tables: Q1 on [-1,1]^2 (Shards vertex order), 2 x 2 Gauss rule
  const double g = 1.0 / std::sqrt(3.0);
  const double qp[4][2] = {{-g, -g}, {g, -g}, {-g, g}, {g, g}};
  const double qw[4] = {1.0, 1.0, 1.0, 1.0};
  const double vx[4] = {-1, 1, 1, -1}, vy[4] = {-1, -1, 1, 1};
  std::vector<double> val(4 * 4), grad(4 * 4 * 2);
  for (int i = 0; i < 4; ++i)
    for (int q = 0; q < 4; ++q) {
      val[i * 4 + q] = 0.25 * (1 + vx[i] * qp[q][0]) * (1 + vy[i] * qp[q][1]);
      grad[(i * 4 + q) * 2 + 0] = 0.25 * vx[i] * (1 + vy[i] * qp[q][1]);
      grad[(i * 4 + q) * 2 + 1] = 0.25 * vy[i] * (1 + vx[i] * qp[q][0]);
    }
  mrhyde_b200_basis basis = {"HGRAD", 1, 4, val.data(), grad.data(), nullptr, nullptr};
  const char* var_names[1] = {"T"};
  const int32_t var_basis[1] = {0};
  const int32_t offsets[4] = {0, 1, 2, 3};
  mrhyde_b200_desc d;
  d.physics = "thermal"; d.dim = 2; d.nvars = 1; d.var_names = var_names; d.var_basis = var_basis;
  d.nbases = 1; d.bases = &basis; d.ndof_elem = 4; d.offsets = offsets; d.max_card = 4;
  d.nqp = 4; d.qp_pts = &qp[0][0]; d.qp_wts = qw;

  mrhyde_b200_plan* plan = nullptr;
  check(mrhyde_b200_plan_create(&plan, &d, device), "plan_create");
  check(mrhyde_b200_plan_set_function(plan, "thermal source", "8*(pi*pi)*sin(2*pi*x)*sin(2*pi*y)"), "set_function");
  check(mrhyde_b200_plan_set_option(plan, "accumulate", "false"), "set_option");

  // ---- mesh: vertex table + connectivity; one dof per node, numbered like the nodes
  std::vector<double> coords(2 * (size_t)n_nodes);
  for (int j = 0; j < nn; ++j)
    for (int i = 0; i < nn; ++i) { coords[2 * (j * nn + i)] = (double)i / nx; coords[2 * (j * nn + i) + 1] = (double)j / nx; }
  std::vector<int32_t> conn(4 * (size_t)n_elem);
  for (int j = 0; j < nx; ++j)
    for (int i = 0; i < nx; ++i) {
      int32_t* c = &conn[4 * (size_t)(j * nx + i)];
      c[0] = j * nn + i; c[1] = c[0] + 1; c[2] = c[1] + nn; c[3] = c[0] + nn;
    }
  check(mrhyde_b200_plan_set_mesh_indexed(plan, n_nodes, coords.data(), n_elem, conn.data(), conn.data(), nullptr), "set_mesh_indexed");

  // ---- graph of J (node-to-node through elements, sorted columns) and the strong-Dirichlet mask (all boundaries)
  std::vector<std::set<int32_t>> adj((size_t)n_nodes);
  for (int e = 0; e < n_elem; ++e)
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) adj[(size_t)conn[4 * e + a]].insert(conn[4 * e + b]);
  std::vector<int64_t> row_map((size_t)n_nodes + 1, 0);
  std::vector<int32_t> entries;
  std::vector<uint8_t> fixed((size_t)n_nodes, 0);
  for (int r = 0; r < n_nodes; ++r) {
    entries.insert(entries.end(), adj[(size_t)r].begin(), adj[(size_t)r].end());
    row_map[(size_t)r + 1] = (int64_t)entries.size();
    const int i = r % nn, j = r / nn;
    fixed[(size_t)r] = (i == 0 || j == 0 || i == nx || j == nx) ? 1 : 0;
  }
  check(mrhyde_b200_plan_set_graph(plan, n_nodes, n_nodes, row_map.data(), entries.data(), fixed.data()), "set_graph");
  check(mrhyde_b200_plan_finalize(plan), "plan_finalize");

  int64_t v = 0;
  check(mrhyde_b200_plan_stat(plan, "n_chains", &v), "plan_stat");
  std::printf("mrhyde_b200 %s: %d elements, %d rows, %lld non-zeros, sweep plan of %lld chains\n", mrhyde_b200_version(), n_elem, n_nodes,
              (long long)entries.size(), (long long)v);

  if (device >= 0) {   // host buffers in, host buffers out (copies inside the call)
    std::vector<double> sol((size_t)n_nodes, 0.0), res((size_t)n_nodes, 0.0), jac(entries.size(), 0.0);
    check(mrhyde_b200_assemble_jacres_host(plan, sol.data(), nullptr, 1, 1, res.data(), jac.data()), "assemble_jacres_host");
    double r2 = 0.0, trace = 0.0;
    for (double x : res) r2 += x * x;
    for (int r = 0; r < n_nodes; ++r)
      for (int64_t p = row_map[(size_t)r]; p < row_map[(size_t)r + 1]; ++p) if (entries[(size_t)p] == r) trace += jac[(size_t)p];
    std::printf("|res|_2 = %.12e   trace(J) = %.12e\n", std::sqrt(r2), trace);
  } else {
    std::printf("host-only plan (device = -1): analysis done, no assembly without a GPU\n");
  }
  mrhyde_b200_plan_destroy(plan);
  return 0;
}
