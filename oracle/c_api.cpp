// ORACLE (test infrastructure, never shipped, never on the product path).
//
// ctypes-facing C entry points of the CPU restatement.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "assembly.hpp"
#include "physics_all.hpp"

using namespace oracle;

static thread_local std::string g_err;

struct OracleHandle {
  std::unique_ptr<AssemblyManager> am;
  std::string scratch;
};

#define ORACLE_TRY try {
#define ORACLE_CATCH(ret) } catch (const std::exception& e) { g_err = e.what(); return ret; }

extern "C" {

const char* oracle_last_error() { return g_err.c_str(); }

// config: "key\tvalue\n" lines with '/'-separated sublist paths (flattened input YAML)
void* oracle_create(const char* config) {
  ORACLE_TRY
  Settings s;
  std::istringstream in(config);
  std::string line;
  while (std::getline(in, line)) {
    size_t t = line.find('\t');
    if (t == std::string::npos) continue;
    s.kv[line.substr(0, t)] = line.substr(t + 1);
  }
  auto* h = new OracleHandle();
  h->am.reset(new AssemblyManager(s));
  return h;
  ORACLE_CATCH(nullptr)
}

void oracle_destroy(void* hv) { delete (OracleHandle*)hv; }

// sizes: [dim, num_nodes, num_elems, nverts, ndof_elem, num_dofs, nnz, nqp, nqp_side, num_groups, num_bgroups, type_AD, nvars, nbases, workset]
int oracle_sizes(void* hv, int64_t* out) {
  auto& am = *((OracleHandle*)hv)->am;
  out[0] = am.mesh.dim; out[1] = am.mesh.num_nodes; out[2] = am.mesh.num_elems; out[3] = am.mesh.topo.nverts;
  out[4] = am.dofs.ndof_elem; out[5] = am.dofs.num_dofs; out[6] = (int64_t)am.graph.colind.size(); out[7] = am.cub.n;
  out[8] = am.side_rules.empty() ? 0 : am.side_rules[0].n; out[9] = (int64_t)am.groups.size(); out[10] = (int64_t)am.boundary_groups.size();
  out[11] = am.type_AD; out[12] = (int64_t)am.dofs.vars.size(); out[13] = (int64_t)am.dofs.bases.size(); out[14] = am.workset_size;
  return 0;
}

int oracle_get_mesh(void* hv, double* nodes, int32_t* conn, int32_t* lids) {
  auto& am = *((OracleHandle*)hv)->am;
  if (nodes) std::memcpy(nodes, am.mesh.nodes.data(), am.mesh.nodes.size() * sizeof(double));
  if (conn) std::memcpy(conn, am.mesh.conn.data(), am.mesh.conn.size() * sizeof(int));
  if (lids) std::memcpy(lids, am.dofs.lids.data(), am.dofs.lids.size() * sizeof(int));
  return 0;
}

int oracle_get_graph(void* hv, int64_t* rowptr, int32_t* colind, uint8_t* is_fixed) {
  auto& am = *((OracleHandle*)hv)->am;
  if (rowptr) std::memcpy(rowptr, am.graph.rowptr.data(), am.graph.rowptr.size() * sizeof(int64_t));
  if (colind) std::memcpy(colind, am.graph.colind.data(), am.graph.colind.size() * sizeof(int));
  if (is_fixed) std::memcpy(is_fixed, am.isFixedDOF.data(), am.isFixedDOF.size());
  return 0;
}

// offsets flattened [var][maxdof] (-1 padded), numdof[var], basis index [var]
int oracle_get_offsets(void* hv, int32_t* offsets, int32_t* numdof, int32_t* usebasis, int32_t* maxdof) {
  auto& am = *((OracleHandle*)hv)->am;
  *maxdof = am.maxdof;
  for (size_t v = 0; v < am.dofs.vars.size(); ++v) {
    numdof[v] = (int)am.dofs.offsets[v].size();
    usebasis[v] = am.dofs.vars[v].basis;
    for (int d = 0; d < am.maxdof; ++d) offsets[v * am.maxdof + d] = d < numdof[v] ? am.dofs.offsets[v][d] : -1;
  }
  return 0;
}

// reference quadrature and basis tables (what setReferenceBasisData produces)
int oracle_get_quadrature(void* hv, double* pts, double* wts) {
  auto& am = *((OracleHandle*)hv)->am;
  std::memcpy(pts, am.cub.pts.data(), am.cub.pts.size() * sizeof(double));
  std::memcpy(wts, am.cub.wts.data(), am.cub.wts.size() * sizeof(double));
  return 0;
}
// basis b: sizes out = [card, vdim, has_grad, has_curl, has_div]; arrays may be NULL
int oracle_get_ref_basis(void* hv, int b, int64_t* sizes, double* val, double* grad, double* curl, double* div) {
  auto& am = *((OracleHandle*)hv)->am;
  const RefBasisTab& t = am.ref_basis[b];
  sizes[0] = t.card; sizes[1] = t.vdim; sizes[2] = !t.grad.empty(); sizes[3] = !t.curl.empty(); sizes[4] = !t.div.empty();
  if (val) std::memcpy(val, t.val.data(), t.val.size() * sizeof(double));
  if (grad && !t.grad.empty()) std::memcpy(grad, t.grad.data(), t.grad.size() * sizeof(double));
  if (curl && !t.curl.empty()) std::memcpy(curl, t.curl.data(), t.curl.size() * sizeof(double));
  if (div && !t.div.empty()) std::memcpy(div, t.div.data(), t.div.size() * sizeof(double));
  return 0;
}
// side rule of Shards local side `side`: pts (n,dim) in the cell frame, wts (n), tanU/tanV (3 each)
int oracle_get_side_rule(void* hv, int side, double* pts, double* wts, double* tanU, double* tanV) {
  auto& am = *((OracleHandle*)hv)->am;
  const SideRule& s = am.side_rules[side];
  std::memcpy(pts, s.pts.data(), s.pts.size() * sizeof(double));
  std::memcpy(wts, s.wts.data(), s.wts.size() * sizeof(double));
  std::memcpy(tanU, s.tanU, 3 * sizeof(double));
  std::memcpy(tanV, s.tanV, 3 * sizeof(double));
  return 0;
}
int oracle_get_ref_basis_side(void* hv, int side, int b, double* val, double* grad) {
  auto& am = *((OracleHandle*)hv)->am;
  const RefBasisTab& t = am.ref_basis_side[side][b];
  if (val) std::memcpy(val, t.val.data(), t.val.size() * sizeof(double));
  if (grad && !t.grad.empty()) std::memcpy(grad, t.grad.data(), t.grad.size() * sizeof(double));
  return 0;
}
// boundary group g: out = [numElem, sideset, local_side]; elem ids copied if non-NULL
int oracle_get_bgroup(void* hv, int g, int64_t* out, int32_t* elem_ids) {
  auto& am = *((OracleHandle*)hv)->am;
  const Group& G = am.boundary_groups[g];
  out[0] = G.numElem; out[1] = G.sideset; out[2] = G.local_side;
  if (elem_ids) std::memcpy(elem_ids, G.elem_ids.data(), G.elem_ids.size() * sizeof(int));
  return 0;
}
// BC type code per (var, sideset): 0 none, 1 strong Dirichlet, 2 weak Dirichlet, 3 Neumann
int oracle_get_bcs(void* hv, int32_t* codes) {
  auto& am = *((OracleHandle*)hv)->am;
  const size_t ns = am.mesh.side_names.size();
  for (size_t v = 0; v < am.bcs.size(); ++v)
    for (size_t s = 0; s < ns; ++s) {
      const std::string& t = am.bcs[v][s].type;
      codes[v * ns + s] = t == "Dirichlet" ? 1 : t == "weak Dirichlet" ? 2 : t == "Neumann" ? 3 : 0;
    }
  return 0;
}

// string-valued metadata: "modules", "var_names", "basis_types", "basis_orders" (comma separated),
// "bc_expr <var> <side>" (the boundary data expression)
const char* oracle_get_string(void* hv, const char* key) {
  auto* h = (OracleHandle*)hv;
  auto& am = *h->am;
  const std::string k(key);
  std::string out;
  auto join = [&](const std::vector<std::string>& v) { std::string s; for (size_t i = 0; i < v.size(); ++i) { if (i) s += ","; s += v[i]; } return s; };
  if (k == "modules") out = join(am.modules);
  else if (k == "var_names") { std::vector<std::string> v; for (auto& x : am.dofs.vars) v.push_back(x.name); out = join(v); }
  else if (k == "basis_types") { std::vector<std::string> v; for (auto& b : am.dofs.bases) v.push_back(b.type); out = join(v); }
  else if (k == "basis_orders") { std::vector<std::string> v; for (auto& b : am.dofs.bases) v.push_back(std::to_string(b.order)); out = join(v); }
  else if (k.compare(0, 8, "bc_expr ") == 0) {
    int v = 0, s = 0;
    if (std::sscanf(key + 8, "%d %d", &v, &s) == 2 && v >= 0 && v < (int)am.bcs.size() && s >= 0 && s < (int)am.bcs[v].size()) out = am.bcs[v][s].expr;
  }
  h->scratch = out;
  return h->scratch.c_str();
}

int oracle_set_time(void* hv, int isTransient, double time, double deltat, int stage, int nstages, const double* A, const double* b,
                    const double* c, int nbdf, const double* bdf) {
  auto& am = *((OracleHandle*)hv)->am;
  am.td.isTransient = isTransient != 0;
  am.td.time = time; am.td.deltat = deltat; am.td.stage = stage;
  am.td.butcher_A.assign(A, A + (size_t)nstages * nstages);
  am.td.butcher_b.assign(b, b + nstages);
  am.td.butcher_c.assign(c, c + nstages);
  am.td.BDF_wts.assign(bdf, bdf + nbdf);
  return 0;
}

int oracle_set_seeding(void* hv, int seedwhat, int seedindex) {
  ((OracleHandle*)hv)->am->seedwhat = seedwhat;
  ((OracleHandle*)hv)->am->seedindex = seedindex;
  return 0;
}

int oracle_set_point_dofs(void* hv, int n, const int* dofs) {
  ((OracleHandle*)hv)->am->point_dofs.assign(dofs, dofs + n);
  return 0;
}

int oracle_set_adjoint(void* hv, int useadjoint) {
  ((OracleHandle*)hv)->am->useadjoint = useadjoint != 0;
  return 0;
}

int oracle_assemble_jacres(void* hv, const double* sol, const double* const* sol_prev, const double* const* sol_stage,
                           int compute_jacobian, double* res, double* Jvals) {
  ORACLE_TRY
  ((OracleHandle*)hv)->am->assembleJacRes(sol, sol_prev, sol_stage, compute_jacobian != 0, res, Jvals);
  return 0;
  ORACLE_CATCH(1)
}

int oracle_apply_mass(void* hv, const double* masswts, const double* x, double* y) {
  ORACLE_TRY
  ((OracleHandle*)hv)->am->applyMassMatrixFree(masswts, x, y);
  return 0;
  ORACLE_CATCH(1)
}

int oracle_project_initial(void* hv, double* rhs) {
  ORACLE_TRY
  auto* h = (OracleHandle*)hv;
  h->am->eng_scalar->projectInitial(*h->am, rhs);
  return 0;
  ORACLE_CATCH(1)
}

int oracle_weighted_mass(void* hv, const double* masswts, int lump, double* Mvals, double* diag) {
  ORACLE_TRY
  ((OracleHandle*)hv)->am->weightedMass(masswts, lump != 0, Mvals, diag);
  return 0;
  ORACLE_CATCH(1)
}

int oracle_assemble_res(void* hv, const double* sol, const double* const* sol_prev, const double* const* sol_stage, double* res) {
  ORACLE_TRY
  ((OracleHandle*)hv)->am->assembleRes(sol, sol_prev, sol_stage, res);
  return 0;
  ORACLE_CATCH(1)
}

// (numElem of group, npts) values of a registered function / solution field, and the group's weights
int oracle_eval_function(void* hv, const char* name, const char* loc, int grp, double* out) {
  ORACLE_TRY
  auto* h = (OracleHandle*)hv;
  h->am->eng_scalar->evalFunction(*h->am, name, loc, grp, out);
  return 0;
  ORACLE_CATCH(1)
}
int oracle_eval_field(void* hv, const char* label, const double* sol, int grp, double* out) {
  ORACLE_TRY
  auto* h = (OracleHandle*)hv;
  h->am->eng_scalar->evalField(*h->am, label, sol, grp, out);
  return 0;
  ORACLE_CATCH(1)
}
int oracle_group_info(void* hv, int grp, int boundary, int64_t* numElem, double* wts, double* ipx, double* ipy, double* ipz) {
  auto& am = *((OracleHandle*)hv)->am;
  const Group& g = boundary ? am.boundary_groups[grp] : am.groups[grp];
  *numElem = g.numElem;
  if (wts) std::memcpy(wts, g.wts.data(), g.wts.size() * sizeof(double));
  double* ips[3] = {ipx, ipy, ipz};
  for (int d = 0; d < am.mesh.dim; ++d) if (ips[d]) std::memcpy(ips[d], g.ip[d].data(), g.ip[d].size() * sizeof(double));
  return 0;
}
const char* oracle_print_tree(void* hv, const char* name, const char* loc) {
  ORACLE_TRY
  auto* h = (OracleHandle*)hv;
  h->scratch = h->am->eng_scalar->printTree(name, loc);
  return h->scratch.c_str();
  ORACLE_CATCH(nullptr)
}

}  // extern "C"
