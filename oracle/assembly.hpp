// ORACLE (test infrastructure, never shipped, never on the product path).
//
// CPU restatement of the reference's element assembly, structured like the reference
// (groups of `workset size` elements, stored physical basis tables, one pass per field and
// per physics kernel, fused scatter with linear column search):
//   AssemblyManager ctor: groups, AD width, constraints      assemblyManager_construct.hpp:120-175
//   createGroups (sequential chunks; boundary groups per (sideset, local side))
//                                                            assemblyManager_groups.hpp:13-468
//   allocateGroupStorage -> Group::computeBasis               group.cpp:134-250
//   createConstraints (isFixedDOF mask)                      assemblyManager_constraints.hpp:13-92
//   assembleJacRes<EvalT> (group loop, boundary loop, dofConstraints)
//                                                            assemblyManager_jacres.hpp:119-630
//   assembleRes (ScalarT, seedwhat 0)                        assemblyManager_jacres.hpp:668-875
//   performGather                                            assemblyManager_gather.hpp:181-234
//   updateWorkset / updateWorksetBoundary                    assemblyManager_workset.hpp:963-1049, 680-746
//   fused scatter<EvalT>                                     assemblyManager_scatter.hpp:162-278
//   dofConstraints -> setJacobianConstraints                 assemblyManager_constraints.hpp:125-138, 241-270
//   PhysicsInterface::volumeResidual / boundaryResidual      physicsInterface_residual.hpp:13-216
//   PhysicsImporter (by-name module factory)                 physicsImporter.cpp:64-281
#pragma once
#include <chrono>
#include <cstring>
#include <memory>
#include <sstream>

#include "physics_base.hpp"

namespace oracle {

typedef SFad<2> AD2;
typedef SFad<4> AD4;
typedef SFad<8> AD8;
typedef SFad<16> AD16;
typedef SFad<18> AD18;
typedef SFad<24> AD24;
typedef SFad<32> AD32;
#ifndef ORACLE_MAXDERIVS
#define ORACLE_MAXDERIVS 81  // reference default is 64; hex-Q2 elasticity needs >= 81 (discretizationInterface_dof.hpp:96)
#endif
typedef SFad<ORACLE_MAXDERIVS> AD;

template <class EvalT>
std::unique_ptr<PhysicsBase<EvalT>> import_physics(const std::string& name, const Settings& modset, int dim);

struct BCSpec {  // one (variable, sideset) boundary condition
  std::string type;  // "Dirichlet" (strong) | "weak Dirichlet" | "Neumann" | "none"
  std::string expr;
};

class AssemblyManager;

struct EngineBase {
  virtual ~EngineBase() {}
  virtual void assemble(AssemblyManager& am, const double* sol, const double* const* sol_prev, const double* const* sol_stage,
                        bool compute_jacobian, bool seed, double* res, double* Jvals) = 0;
  virtual std::string printTree(const std::string& name, const std::string& loc) = 0;
  virtual void evalFunction(AssemblyManager& am, const std::string& name, const std::string& loc, int grp, double* out) = 0;
  virtual void evalField(AssemblyManager& am, const std::string& label, const double* sol, int grp, double* out) = 0;
  virtual void projectInitial(AssemblyManager& am, double* rhs) = 0;
};

class AssemblyManager {
 public:
  Settings settings;
  BrickMesh mesh;
  DofMap dofs;
  CsrGraph graph;
  Cubature cub;
  int quadorder = 2, side_quadorder = 2;
  std::vector<SideRule> side_rules;                        // per Shards local side
  std::vector<RefBasisTab> ref_basis;                      // per basis, at volume ip
  std::vector<std::vector<RefBasisTab>> ref_basis_side;    // [local side][basis]
  int workset_size = 100;
  std::vector<Group> groups, boundary_groups;
  std::vector<uint8_t> isFixedDOF;
  std::vector<int> dbc_dofs;
  std::vector<double> dbc_vals;
  std::vector<std::vector<BCSpec>> bcs;                    // [var][sideset]
  std::vector<std::string> modules;
  int type_AD = 0, maxdof = 0;
  bool assemble_volume_terms = true, assemble_boundary_terms = true, use_strong_DBCs = true;
  bool lump_mass = false;    // Solver: lump mass (assemblyManager_construct.hpp:36): the fused scatter sends every entry of a row to its diagonal
  std::vector<int> point_dofs;   // disc->point_dofs (local ids): dofConstraints replaces their whole Jacobian row by the identity row
  int seedwhat = 1, seedindex = 0;   // compute_previous_jac: seedwhat = 2, seedindex = stepindex (assemblyManager_jacres.hpp:176-190); 3: a previous stage
  bool fix_zero_rows = false;   // Solver: fix zero rows (assemblyManager_construct.hpp:33)
  bool useadjoint = false;   // assembleJacRes(..., useadjoint, ...): transposed local Jacobians (updateJac, assemblyManager_jacres.hpp:1459-1475)
  TimeData td;
  std::unique_ptr<EngineBase> eng_scalar, eng_ad;
  std::vector<int8_t> orient_sign;  // (num_elems, ndof_elem), +1 on lexicographic bricks

  explicit AssemblyManager(const Settings& s);
  void assembleJacRes(const double* sol, const double* const* sol_prev, const double* const* sol_stage,
                      bool compute_jacobian, double* res, double* Jvals) {
    eng_ad->assemble(*this, sol, sol_prev, sol_stage, compute_jacobian, true, res, Jvals);
  }
  void assembleRes(const double* sol, const double* const* sol_prev, const double* const* sol_stage, double* res) {
    eng_scalar->assemble(*this, sol, sol_prev, sol_stage, false, false, res, nullptr);
  }
  void computeGroupBasis(Group& g, bool boundary);
  // getWeightedMass (assemblyManager_mass.hpp:13-275; element matrices :1065-1146): per-variable mass blocks
  //   localmass(e, off(i), off(j)) += basis(e,i,k,d) basis(e,j,k,d) wts(e,k) mwt[n]
  // summed into the CSR values with sumIntoValues (no isFixedDOF check), and the diagonal vector: Jacobi
  // (localmass(e,row,row)) or lumped (sum_k |localmass(e,row,col_k)|, the LA-accessible branch :126-148)
  // applyMassMatrixFree (assemblyManager_mass.hpp:555-800, the !storeMass branch): y(indi) += massval x(indj) per variable block
  void applyMassMatrixFree(const double* masswts, const double* x, double* y) const {
    const int ndofE = dofs.ndof_elem;
    for (const Group& g : groups)
      for (int e = 0; e < g.numElem; ++e) {
        const int* LIDs = &g.LIDs[(size_t)e * ndofE];
        for (size_t n = 0; n < dofs.vars.size(); ++n) {
          const int b = dofs.vars[n].basis;
          const Basis& B = dofs.bases[b];
          const double* cb = &g.basis[b][(size_t)e * B.card * cub.n * B.vdim];
          const auto& off = dofs.offsets[n];
          for (int i = 0; i < B.card; ++i)
            for (int j = 0; j < B.card; ++j) {
              double massval = 0.0;
              for (int k = 0; k < cub.n; ++k)
                for (int d = 0; d < B.vdim; ++d)
                  massval += cb[((size_t)i * cub.n + k) * B.vdim + d] * cb[((size_t)j * cub.n + k) * B.vdim + d] * g.wts[(size_t)e * cub.n + k] * masswts[n];
              y[LIDs[off[i]]] += massval * x[LIDs[off[j]]];
            }
        }
      }
  }
  void weightedMass(const double* masswts, bool lump, double* Mvals, double* diag) const {
    const int ndofE = dofs.ndof_elem;
    std::vector<double> lm((size_t)ndofE * ndofE);
    for (const Group& g : groups) {
      for (int e = 0; e < g.numElem; ++e) {
        std::fill(lm.begin(), lm.end(), 0.0);
        for (size_t n = 0; n < dofs.vars.size(); ++n) {
          const int b = dofs.vars[n].basis;
          const Basis& B = dofs.bases[b];
          const double* cb = &g.basis[b][(size_t)e * B.card * cub.n * B.vdim];
          const auto& off = dofs.offsets[n];
          for (int i = 0; i < B.card; ++i)
            for (int j = 0; j < B.card; ++j)
              for (int k = 0; k < cub.n; ++k)
                for (int d = 0; d < B.vdim; ++d)
                  lm[(size_t)off[i] * ndofE + off[j]] += cb[((size_t)i * cub.n + k) * B.vdim + d] * cb[((size_t)j * cub.n + k) * B.vdim + d] * g.wts[(size_t)e * cub.n + k] * masswts[n];
        }
        const int* LIDs = &g.LIDs[(size_t)e * ndofE];
        for (size_t n = 0; n < dofs.vars.size(); ++n) {
          const auto& off = dofs.offsets[n];
          for (size_t j = 0; j < off.size(); ++j) {
            const int row = off[j], rowIndex = LIDs[row];
            if (diag) {
              double val = 0.0;
              if (!lump) val = lm[(size_t)row * ndofE + row];
              else for (size_t k = 0; k < off.size(); ++k) val += std::fabs(lm[(size_t)row * ndofE + off[k]]);
              diag[rowIndex] += val;
            }
            if (Mvals)
              for (size_t k = 0; k < off.size(); ++k) {
                const int col = LIDs[off[k]];
                for (int64_t p = graph.rowptr[rowIndex]; p < graph.rowptr[rowIndex + 1]; ++p)
                  if (graph.colind[p] == col) { Mvals[p] += lm[(size_t)row * ndofE + off[k]]; break; }
              }
          }
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------------------
template <class EvalT>
struct Engine : EngineBase {
  Workset<EvalT> wkset;
  FunctionManager<EvalT> fm;
  std::vector<std::unique_ptr<PhysicsBase<EvalT>>> mods;
  std::vector<double> gsol, gsol_prev, gsol_stage;  // GroupMetaData::sol / sol_prev / sol_stage

  explicit Engine(AssemblyManager& am) {
    const int nsideip = am.side_rules.empty() ? 1 : am.side_rules[0].n;
    wkset = Workset<EvalT>(am.workset_size, am.mesh.dim, am.cub.n, nsideip, am.dofs);
    fm = FunctionManager<EvalT>(am.workset_size, am.cub.n, nsideip);
    wkset.connect(fm);
    wkset.var_bcs.assign(am.dofs.vars.size(), std::vector<std::string>(am.mesh.side_names.size(), "none"));
    for (size_t v = 0; v < am.bcs.size(); ++v)
      for (size_t s = 0; s < am.bcs[v].size(); ++s) wkset.var_bcs[v][s] = am.bcs[v][s].type;
    // user functions first, then module defaults (createFunctions, assemblyManager_functions.hpp)
    Settings fs;
    for (auto& p : am.settings.sub("Functions/")) { fs.kv[p.first] = p.second; }
    for (auto& p : fs.kv) {
      fm.addFunction(p.first, p.second, "ip");
      fm.addFunction(p.first, p.second, "side ip");
    }
    for (auto& name : am.modules) {
      Settings ms;
      for (auto& p : am.settings.sub("Physics/")) ms.kv[p.first] = p.second;
      mods.push_back(import_physics<EvalT>(name, ms, am.mesh.dim));
      mods.back()->defineFunctions(fs, &fm);
      mods.back()->setWorkset(&wkset);
    }
    // boundary data functions: "Dirichlet <var> <side>" / "Neumann <var> <side>" at side ip
    for (size_t v = 0; v < am.bcs.size(); ++v)
      for (size_t s = 0; s < am.bcs[v].size(); ++s) {
        const BCSpec& b = am.bcs[v][s];
        if (b.type == "weak Dirichlet" || b.type == "Dirichlet")
          fm.addFunction("Dirichlet " + am.dofs.vars[v].name + " " + am.mesh.side_names[s], b.expr, "side ip");
        else if (b.type == "Neumann")
          fm.addFunction("Neumann " + am.dofs.vars[v].name + " " + am.mesh.side_names[s], b.expr, "side ip");
      }
    // initial conditions (physicsInterface_functions.hpp:154-226): "initial <var>" for HGRAD / HVOL variables, "initial <var>[x|y|z]"
    // for HCURL / HDIV ones; variables the sublist does not name start from 0.0
    for (size_t v = 0; v < am.dofs.vars.size(); ++v) {
      const std::string& var = am.dofs.vars[v].name;
      const std::string bt = am.dofs.bases[am.dofs.vars[v].basis].type;
      if (bt == "HCURL" || bt == "HDIV") {
        for (const char* c : {"[x]", "[y]", "[z]"}) fm.addFunction("initial " + var + c, am.settings.get("Physics/Initial conditions/" + var + c, "0.0"), "ip");
      } else {
        fm.addFunction("initial " + var, am.settings.get("Physics/Initial conditions/" + var, "0.0"), "ip");
      }
    }
    // true solutions (postprocess) so the gold L2 errors can be reproduced
    for (auto& p : am.settings.sub("Postprocess/True solutions/")) {
      fm.addFunction("true " + p.first, p.second, "ip");
      fm.addFunction("true " + p.first, p.second, "side ip");
    }
    gsol.assign((size_t)am.workset_size * am.dofs.vars.size() * am.maxdof, 0.0);
  }

  void performGather(AssemblyManager& am, const Group& g, const double* vec, std::vector<double>& data, int stride, int slot) {  // gather.hpp:181-234
    const int nvar = (int)am.dofs.vars.size(), ndof = am.dofs.ndof_elem;
    for (int elem = 0; elem < g.numElem; ++elem)
      for (int var = 0; var < nvar; ++var)
        for (int dof = 0; dof < (int)am.dofs.offsets[var].size(); ++dof)
          data[(((size_t)elem * nvar + var) * am.maxdof + dof) * stride + slot] = vec[g.LIDs[(size_t)elem * ndof + am.dofs.offsets[var][dof]]];
  }

  void pointAtGroup(AssemblyManager& am, const Group& g, bool boundary) {
    wkset.numElem = g.numElem;
    const int nb = (int)am.dofs.bases.size();
    const int np = boundary ? wkset.numsideip : wkset.numip;
    for (int b = 0; b < nb; ++b) {
      const Basis& B = am.dofs.bases[b];
      View4 v{g.basis[b].data(), B.card, np, B.vdim};
      View4 gr{g.basis_grad[b].empty() ? nullptr : g.basis_grad[b].data(), B.card, np, g.basis_grad[b].empty() ? 0 : B.dim};
      if (boundary) { wkset.basis_side[b] = v; wkset.basis_grad_side[b] = gr; }
      else {
        wkset.basis[b] = v; wkset.basis_grad[b] = gr;
        wkset.basis_curl[b] = View4{g.basis_curl[b].empty() ? nullptr : g.basis_curl[b].data(), B.card, np, g.basis_curl[b].empty() ? 0 : B.dim};
        wkset.basis_div[b] = View4{g.basis_div[b].empty() ? nullptr : g.basis_div[b].data(), B.card, np, 1};
      }
    }
    auto& sf = boundary ? wkset.side_scalar_fields : wkset.scalar_fields;
    for (int d = 0; d < 3; ++d)
      for (int e = 0; e < g.numElem; ++e)
        for (int q = 0; q < np; ++q) sf[d].data(e, q) = (d < am.mesh.dim) ? g.ip[d][(size_t)e * np + q] : 0.0;
    if (boundary) {
      wkset.wts_side_p = g.wts.data();
      for (int d = 0; d < 3; ++d)
        for (int e = 0; e < g.numElem; ++e)
          for (int q = 0; q < np; ++q) sf[3 + d].data(e, q) = (d < am.mesh.dim) ? g.normals[d][(size_t)e * np + q] : 0.0;
    } else {
      wkset.wts_p = g.wts.data();
    }
  }

  void seed(AssemblyManager& am, bool doseed) {
    if (wkset.isTransient)
      wkset.computeSolnTransientSeeded(gsol, gsol_prev, gsol_stage, am.maxdof, std::max(1, (int)am.td.BDF_wts.size() - 1),
                                       std::max(1, am.td.nstages()), doseed ? am.seedwhat : 0, am.seedindex);
    else
      wkset.computeSolnSteadySeeded(gsol, am.maxdof, doseed ? 1 : 0);
  }

  void gatherAll(AssemblyManager& am, const Group& g, const double* sol, const double* const* sol_prev, const double* const* sol_stage) {
    performGather(am, g, sol, gsol, 1, 0);
    if (wkset.isTransient) {
      const int nsteps = std::max(1, (int)am.td.BDF_wts.size() - 1), nstages = std::max(1, am.td.nstages());
      const size_t n = (size_t)am.workset_size * am.dofs.vars.size() * am.maxdof;
      gsol_prev.assign(n * nsteps, 0.0);
      gsol_stage.assign(n * nstages, 0.0);
      for (int s = 0; s < nsteps; ++s) performGather(am, g, sol_prev[s], gsol_prev, nsteps, s);
      for (int s = 0; s < nstages; ++s) performGather(am, g, sol_stage[s], gsol_stage, nstages, s);
    }
  }

  // fused scatter (assemblyManager_scatter.hpp:162-278), serial build: plain +=, linear column search
  void scatter(AssemblyManager& am, const Group& g, bool compute_jacobian, double* res_view, double* Jvals) {
    const auto& offsets = wkset.offsets;
    const int ndofE = am.dofs.ndof_elem;
    auto& res = wkset.res;
    for (int elem = 0; elem < g.numElem; ++elem) {
      const int* LIDs = &g.LIDs[(size_t)elem * ndofE];
      for (size_t n = 0; n < offsets.size(); ++n)
        for (size_t j = 0; j < offsets[n].size(); ++j) {
          const int row = offsets[n][j];
          const int rowIndex = LIDs[row];
          if (!am.isFixedDOF[rowIndex]) res_view[rowIndex] += -ADTraits<EvalT>::val(res(elem, row));
        }
      if (compute_jacobian && ADTraits<EvalT>::size > 0) {
        int cols[ORACLE_MAXDERIVS];
        double vals[ORACLE_MAXDERIVS];
        for (size_t n = 0; n < offsets.size(); ++n)
          for (size_t j = 0; j < offsets[n].size(); ++j) {
            const int row = offsets[n][j];
            const int rowIndex = LIDs[row];
            if (am.isFixedDOF[rowIndex]) continue;
            for (size_t m = 0; m < offsets.size(); ++m)
              for (size_t k = 0; k < offsets[m].size(); ++k) {
                const int col = offsets[m][k];
                // adjoint: local_J(elem, offsets(m,k), offsets(n,j)) += res(elem, offsets(n,j)).dx(offsets(m,k)), i.e. the entry in
                // (row, col) of the scattered matrix is d res(col) / d u(row)   (updateJac / updateJacBoundary, useadjoint branch)
                vals[col] = am.useadjoint ? ADTraits<EvalT>::dx(res(elem, col), row) : ADTraits<EvalT>::dx(res(elem, row), col);
                cols[col] = am.lump_mass ? rowIndex : LIDs[col];   // assemblyManager_scatter.hpp:263-268
              }
            // KokkosSparse::CrsMatrix::sumIntoValues(row, cols, n, vals, is_sorted=false): linear search per entry
            const int64_t rs = am.graph.rowptr[rowIndex], re = am.graph.rowptr[rowIndex + 1];
            for (int c = 0; c < ndofE; ++c)
              for (int64_t p = rs; p < re; ++p)
                if (am.graph.colind[p] == cols[c]) { Jvals[p] += vals[c]; break; }
          }
      }
    }
  }

  void assemble(AssemblyManager& am, const double* sol, const double* const* sol_prev, const double* const* sol_stage,
                bool compute_jacobian, bool doseed, double* res, double* Jvals) override {
    // updateWorksetTime (assemblyManager_workset.hpp): t = current stage time, alpha = 1/dt
    wkset.isTransient = am.td.isTransient;
    wkset.isAdjoint = am.useadjoint;
    wkset.td = am.td;
    wkset.deltat = am.td.deltat;
    wkset.alpha = 1.0 / am.td.deltat;
    wkset.current_stage = am.td.stage;
    wkset.time = am.td.isTransient && !am.td.butcher_c.empty() ? am.td.time + am.td.butcher_c[am.td.stage] * am.td.deltat : am.td.time;
    if (am.assemble_volume_terms) {
      wkset.isOnSide = false;
      for (size_t grp = 0; grp < am.groups.size(); ++grp) {
        const Group& g = am.groups[grp];
        gatherAll(am, g, sol, sol_prev, sol_stage);
        pointAtGroup(am, g, false);
        wkset.reset();
        seed(am, doseed);
        for (auto& m : mods) m->volumeResidual();
        scatter(am, g, compute_jacobian, res, Jvals);
      }
    }
    if (am.assemble_boundary_terms) {
      wkset.isOnSide = true;
      for (size_t grp = 0; grp < am.boundary_groups.size(); ++grp) {
        const Group& g = am.boundary_groups[grp];
        if (g.numElem == 0) continue;
        gatherAll(am, g, sol, sol_prev, sol_stage);
        pointAtGroup(am, g, true);
        wkset.reset();
        wkset.currentside = g.sideset;
        wkset.sidename = g.sidename;
        seed(am, doseed);
        for (auto& m : mods) m->boundaryResidual();
        scatter(am, g, compute_jacobian, res, Jvals);
      }
      wkset.isOnSide = false;
    }
    // dofConstraints -> setJacobianConstraints: J(d,d) = 1 on strong-Dirichlet dofs (replaceLocalValues)
    if (compute_jacobian && Jvals && am.use_strong_DBCs) {
      for (int d : am.dbc_dofs) {
        for (int64_t p = am.graph.rowptr[d]; p < am.graph.rowptr[d + 1]; ++p)
          if (am.graph.colind[p] == d) Jvals[p] = 1.0;
      }
    }
    // point constraints: setJacobianConstraints(J, fixedDOFs, block, ...) sets ALL entries of the row to 0 and the diagonal to 1
    // (assemblyManager_constraints.hpp:97-116, 261-266); the residual entry stays what the assembly made it
    if (compute_jacobian && Jvals) {
      for (int d : am.point_dofs)
        for (int64_t p = am.graph.rowptr[d]; p < am.graph.rowptr[d + 1]; ++p) Jvals[p] = (am.graph.colind[p] == d) ? 1.0 : 0.0;
    }
    // fix_zero_rows (assemblyManager_jacres.hpp:609-626): a row whose entries sum to less than 1e-14 in absolute value gets a unit diagonal
    if (am.fix_zero_rows && Jvals) {
      for (int64_t row = 0; row < am.dofs.num_dofs; ++row) {
        double abssum = 0.0;
        for (int64_t p = am.graph.rowptr[row]; p < am.graph.rowptr[row + 1]; ++p) abssum += std::abs(Jvals[p]);
        if (abssum < 1.0e-14)
          for (int64_t p = am.graph.rowptr[row]; p < am.graph.rowptr[row + 1]; ++p) if (am.graph.colind[p] == row) Jvals[p] = 1.0;
      }
    }
  }

  std::string printTree(const std::string& name, const std::string& loc) override { return fm.printTree(name, loc); }

  void evalFunction(AssemblyManager& am, const std::string& name, const std::string& loc, int grp, double* out) override {
    const bool side = (loc == "side ip");
    const Group& g = side ? am.boundary_groups[grp] : am.groups[grp];
    wkset.isOnSide = side;
    // postprocess evaluates true solutions at the current (post-step) time: am.td.time when steady, else stage time
    wkset.time = am.td.isTransient && !am.td.butcher_c.empty() ? am.td.time + am.td.butcher_c[am.td.stage] * am.td.deltat : am.td.time;
    pointAtGroup(am, g, side);
    wkset.reset();
    Vista<EvalT> v = fm.evaluate(name, loc);
    const int np = side ? wkset.numsideip : wkset.numip;
    for (int e = 0; e < g.numElem; ++e)
      for (int q = 0; q < np; ++q) out[(size_t)e * np + q] = ADTraits<EvalT>::val(v(e, q));
    wkset.isOnSide = false;
  }

  // setInitial, right-hand side of the L2 projection (assemblyManager_initial.hpp:36-76 with getInitial(project = true), :260-311, and
  // PhysicsInterface::getInitial, physicsInterface_initial.hpp:30-95):
  //   rhs(LID(e, off(dof))) += sum_pt initial_var(e, pt[, dim]) basis(e, dof, pt[, dim]) wts(e, pt)      (no isFixedDOF check)
  void projectInitial(AssemblyManager& am, double* rhs) override {
    const int ndofE = am.dofs.ndof_elem, np = wkset.numip, dim = am.mesh.dim;
    wkset.isOnSide = false;
    wkset.time = am.td.time;
    static const char* comp[3] = {"[x]", "[y]", "[z]"};
    for (const Group& g : am.groups) {
      pointAtGroup(am, g, false);
      wkset.reset();
      for (size_t n = 0; n < am.dofs.vars.size(); ++n) {
        const int b = am.dofs.vars[n].basis;
        const Basis& B = am.dofs.bases[b];
        const auto& off = am.dofs.offsets[n];
        const double* cb = &g.basis[b][0];
        const bool vec = (B.type == "HCURL" || B.type == "HDIV");
        for (int d = 0; d < (vec ? dim : 1); ++d) {
          Vista<EvalT> v = fm.evaluate("initial " + am.dofs.vars[n].name + (vec ? comp[d] : ""), "ip");
          for (int e = 0; e < g.numElem; ++e)
            for (int dof = 0; dof < B.card; ++dof)
              for (int pt = 0; pt < np; ++pt)
                rhs[g.LIDs[(size_t)e * ndofE + off[dof]]] += ADTraits<EvalT>::val(v(e, pt)) * cb[(((size_t)e * B.card + dof) * np + pt) * B.vdim + d] * g.wts[(size_t)e * np + pt];
        }
      }
    }
  }

  void evalField(AssemblyManager& am, const std::string& label, const double* sol, int grp, double* out) override {
    const Group& g = am.groups[grp];
    wkset.isOnSide = false;
    const bool tr = wkset.isTransient;
    wkset.isTransient = false;
    performGather(am, g, sol, gsol, 1, 0);
    pointAtGroup(am, g, false);
    wkset.reset();
    wkset.computeSolnSteadySeeded(gsol, am.maxdof, 0);
    auto& f = wkset.getSolutionField(label);
    for (int e = 0; e < g.numElem; ++e)
      for (int q = 0; q < wkset.numip; ++q) out[(size_t)e * wkset.numip + q] = ADTraits<EvalT>::val(f(e, q));
    wkset.isTransient = tr;
  }
};

// ---------------------------------------------------------------------------------------
inline std::vector<std::string> split_list(const std::string& s) {
  std::vector<std::string> out;
  std::string cur;
  for (char c : s) {
    if (c == ',' ) { if (!cur.empty()) out.push_back(cur); cur.clear(); }
    else cur += c;
  }
  if (!cur.empty()) out.push_back(cur);
  for (auto& t : out) {
    size_t a = t.find_first_not_of(' '), b = t.find_last_not_of(' ');
    t = (a == std::string::npos) ? "" : t.substr(a, b - a + 1);
  }
  return out;
}

inline void AssemblyManager::computeGroupBasis(Group& g, bool boundary) {  // Group::computeBasis, group.cpp:134-250
  const int dim = mesh.dim, nv = mesh.topo.nverts;
  const int nb = (int)dofs.bases.size();
  const double* pts = boundary ? side_rules[g.local_side].pts.data() : cub.pts.data();
  const int np = boundary ? side_rules[g.local_side].n : cub.n;
  g.wts.assign((size_t)g.numElem * np, 0.0);
  for (int d = 0; d < dim; ++d) g.ip[d].assign((size_t)g.numElem * np, 0.0);
  if (boundary) for (int d = 0; d < dim; ++d) g.normals[d].assign((size_t)g.numElem * np, 0.0);
  g.basis.assign(nb, {}); g.basis_grad.assign(nb, {}); g.basis_curl.assign(nb, {}); g.basis_div.assign(nb, {});
  for (int b = 0; b < nb; ++b) {
    const Basis& B = dofs.bases[b];
    g.basis[b].assign((size_t)g.numElem * B.card * np * B.vdim, 0.0);
    if (B.type == "HGRAD") g.basis_grad[b].assign((size_t)g.numElem * B.card * np * dim, 0.0);
    if (B.type == "HCURL" && !boundary) g.basis_curl[b].assign((size_t)g.numElem * B.card * np * dim, 0.0);
    if (B.type == "HDIV" && !boundary) g.basis_div[b].assign((size_t)g.numElem * B.card * np, 0.0);
  }
  ElemGeom eg;
  PhysBasis pb;
  std::vector<double> sw, sn;
  std::vector<double> sign(dofs.ndof_elem, 1.0);
  for (int e = 0; e < g.numElem; ++e) {
    const double* nodes = &g.nodes[(size_t)e * nv * dim];
    element_geometry(mesh.topo, nodes, pts, np, eg);
    for (int q = 0; q < np; ++q) {
      for (int d = 0; d < dim; ++d) g.ip[d][(size_t)e * np + q] = eg.ip[(size_t)q * dim + d];
    }
    if (boundary) {
      side_measure(mesh.topo, side_rules[g.local_side], eg, sw, sn);
      for (int q = 0; q < np; ++q) {
        g.wts[(size_t)e * np + q] = sw[q];
        for (int d = 0; d < dim; ++d) g.normals[d][(size_t)e * np + q] = sn[(size_t)q * dim + d];
      }
    } else {
      for (int q = 0; q < np; ++q) g.wts[(size_t)e * np + q] = std::fabs(eg.det[q]) * cub.wts[q];
    }
    for (int b = 0; b < nb; ++b) {
      const Basis& B = dofs.bases[b];
      const RefBasisTab& rt = boundary ? ref_basis_side[g.local_side][b] : ref_basis[b];
      const double* sg = nullptr;
      if (B.type == "HCURL" || B.type == "HDIV") {
        // per-dof orientation sign of the first variable on this basis
        for (int d = 0; d < B.card; ++d) sign[d] = (double)orient_sign[(size_t)g.elem_ids[e] * dofs.ndof_elem + dofs.lbase[b] + d * dofs.nvb[b]];
        sg = sign.data();
      }
      push_forward(B, rt, eg, sg, pb);
      std::memcpy(&g.basis[b][(size_t)e * B.card * np * B.vdim], pb.val.data(), pb.val.size() * sizeof(double));
      if (!pb.grad.empty()) std::memcpy(&g.basis_grad[b][(size_t)e * B.card * np * dim], pb.grad.data(), pb.grad.size() * sizeof(double));
      if (!pb.curl.empty() && !boundary) std::memcpy(&g.basis_curl[b][(size_t)e * B.card * np * dim], pb.curl.data(), pb.curl.size() * sizeof(double));
      if (!pb.div.empty() && !boundary) std::memcpy(&g.basis_div[b][(size_t)e * B.card * np], pb.div.data(), pb.div.size() * sizeof(double));
    }
  }
}

template <class EvalT>
std::unique_ptr<EngineBase> make_engine(AssemblyManager& am) { return std::unique_ptr<EngineBase>(new Engine<EvalT>(am)); }
#ifndef ORACLE_INSTANTIATE  // one translation unit per AD width (oracle/engine_inst.cpp) keeps the build parallel
extern template std::unique_ptr<EngineBase> make_engine<double>(AssemblyManager&);
extern template std::unique_ptr<EngineBase> make_engine<AD2>(AssemblyManager&);
extern template std::unique_ptr<EngineBase> make_engine<AD4>(AssemblyManager&);
extern template std::unique_ptr<EngineBase> make_engine<AD8>(AssemblyManager&);
extern template std::unique_ptr<EngineBase> make_engine<AD16>(AssemblyManager&);
extern template std::unique_ptr<EngineBase> make_engine<AD18>(AssemblyManager&);
extern template std::unique_ptr<EngineBase> make_engine<AD24>(AssemblyManager&);
extern template std::unique_ptr<EngineBase> make_engine<AD32>(AssemblyManager&);
extern template std::unique_ptr<EngineBase> make_engine<AD>(AssemblyManager&);
#endif

inline AssemblyManager::AssemblyManager(const Settings& s) : settings(s) {
  // ---- mesh (Mesh sublist; SimpleMeshManager numbering)
  mesh.dim = s.geti("Mesh/dimension", 3);
  mesh.n[0] = s.geti("Mesh/NX", 1); mesh.n[1] = s.geti("Mesh/NY", 1); mesh.n[2] = s.geti("Mesh/NZ", 1);
  mesh.lo[0] = s.getd("Mesh/xmin", 0.0); mesh.hi[0] = s.getd("Mesh/xmax", 1.0);
  mesh.lo[1] = s.getd("Mesh/ymin", 0.0); mesh.hi[1] = s.getd("Mesh/ymax", 1.0);
  mesh.lo[2] = s.getd("Mesh/zmin", 0.0); mesh.hi[2] = s.getd("Mesh/zmax", 1.0);
  // "Periodic BCs" (panzer_stk periodic matchers, meshInterface_construct.hpp): 'xz-all ...: top;bottom' pairs the y sides,
  // 'yz-all ...: left;right' the x sides, 'xy-all ...: back;front' the z sides
  for (auto& p : s.sub("Mesh/Periodic BCs/")) {
    if (p.first == "Count") continue;
    if (p.second.find("yz-all") != std::string::npos) mesh.periodic[0] = true;
    else if (p.second.find("xz-all") != std::string::npos) mesh.periodic[1] = true;
    else if (p.second.find("xy-all") != std::string::npos) mesh.periodic[2] = true;
    else throw std::runtime_error("oracle: periodic condition not restated: " + p.second);
  }
  mesh.build();
  // optional smooth vertex perturbation (synthetic non-affine meshes for parity tests; boundary nodes stay put)
  const double pert = s.getd("Mesh/perturb", 0.0);
  if (pert != 0.0) {
    for (int n = 0; n < mesh.num_nodes; ++n) {
      double* x = &mesh.nodes[(size_t)n * mesh.dim];
      double bub = 1.0;
      for (int d = 0; d < mesh.dim; ++d) { const double t = (x[d] - mesh.lo[d]) / (mesh.hi[d] - mesh.lo[d]); bub *= std::sin(M_PI * t); }
      const double x0 = x[0], y0 = x[1], z0 = (mesh.dim == 3) ? x[2] : 0.0;
      x[0] += pert * bub * std::sin(3.0 * y0 + 1.0 + 2.0 * z0);
      x[1] += pert * bub * std::cos(2.0 * x0 + 0.5 + 3.0 * z0);
      if (mesh.dim == 3) x[2] += pert * bub * std::sin(2.5 * x0 + 1.5 * y0 + 0.3);
    }
  }

  // optional affine shear of the whole brick (synthetic parallelepiped cells with a full Jacobian for parity tests)
  const double shear = s.getd("Mesh/shear", 0.0);
  if (shear != 0.0) {
    for (int n = 0; n < mesh.num_nodes; ++n) {
      double* x = &mesh.nodes[(size_t)n * mesh.dim];
      const double y0 = x[1], z0 = (mesh.dim == 3) ? x[2] : 0.0;
      x[0] += shear * y0 + 0.5 * shear * z0;
      x[1] += 0.7 * shear * z0;
    }
  }

  // ---- physics modules -> variables (PhysicsInterface ctor)
  modules = split_list(s.get("Physics/modules", "thermal"));
  std::vector<VarInfo> vars;
  for (auto& name : modules) {
    Settings ms;
    for (auto& p : s.sub("Physics/")) ms.kv[p.first] = p.second;
    auto m = import_physics<double>(name, ms, mesh.dim);
    for (size_t i = 0; i < m->myvars.size(); ++i) {
      VarInfo v;
      v.name = m->myvars[i]; v.btype = m->mybasistypes[i];
      v.order = s.geti("Discretization/order/" + v.name, 1);
      vars.push_back(v);
    }
  }
  dofs.build(mesh, vars);
  maxdof = 0;
  for (auto& b : dofs.bases) maxdof = std::max(maxdof, b.card);
  orient_sign.assign((size_t)mesh.num_elems * dofs.ndof_elem, 1);
  graph.build(dofs.num_dofs, dofs.lids, mesh.num_elems, dofs.ndof_elem);

  // ---- quadrature (discretizationInterface_construct.hpp:103-150)
  int mxorder = 0;
  for (auto& v : dofs.vars) mxorder = std::max(mxorder, v.order);
  quadorder = s.geti("Discretization/quadrature", 2 * mxorder);
  side_quadorder = s.geti("Discretization/side quadrature", 2 * mxorder);
  cub = tensor_gauss(mesh.dim, quadorder);
  for (size_t sd = 0; sd < mesh.topo.side_nodes.size(); ++sd) side_rules.push_back(make_side_rule(mesh.topo, (int)sd, side_quadorder));
  for (auto& B : dofs.bases) ref_basis.push_back(tabulate(B, cub.pts.data(), cub.n));
  ref_basis_side.resize(side_rules.size());
  for (size_t sd = 0; sd < side_rules.size(); ++sd)
    for (auto& B : dofs.bases) ref_basis_side[sd].push_back(tabulate(B, side_rules[sd].pts.data(), side_rules[sd].n));

  // ---- solver / assembly flags
  workset_size = s.geti("Solver/workset size", 100);
  use_strong_DBCs = s.getb("Solver/use strong DBCs", true);
  lump_mass = s.getb("Solver/lump mass", false);
  fix_zero_rows = s.getb("Solver/fix zero rows", false);
  assemble_volume_terms = s.getb("Physics/assemble volume terms", true);
  assemble_boundary_terms = s.getb("Physics/assemble boundary terms", true);

  // ---- boundary conditions (PhysicsInterface: Dirichlet / Neumann conditions per variable and sideset)
  const int nsides = (int)mesh.side_names.size();
  bcs.assign(dofs.vars.size(), std::vector<BCSpec>(nsides, BCSpec{"none", "0.0"}));
  for (size_t v = 0; v < dofs.vars.size(); ++v) {
    const std::string& vn = dofs.vars[v].name;
    for (const char* kind : {"Dirichlet", "Neumann"}) {
      const std::string pre = std::string("Physics/") + kind + " conditions/" + vn + "/";
      for (auto& p : s.sub(pre)) {
        for (int sd = 0; sd < nsides; ++sd) {
          if (p.first == "all boundaries" || p.first == mesh.side_names[sd]) {
            BCSpec b;
            b.expr = p.second;
            if (std::string(kind) == "Dirichlet") b.type = use_strong_DBCs ? "Dirichlet" : "weak Dirichlet";
            else b.type = "Neumann";
            bcs[v][sd] = b;
          }
        }
      }
    }
  }
  // createConstraints: fixed-dof mask from the strong-Dirichlet side closures
  isFixedDOF.assign(dofs.num_dofs, 0);
  for (int64_t d = 0; d < dofs.num_dofs; ++d) {
    const int v = dofs.var_of_dof(d);
    for (int sd = 0; sd < nsides; ++sd)
      if (dofs.on_side[(size_t)d * nsides + sd] && bcs[v][sd].type == "Dirichlet") isFixedDOF[d] = 1;
    if (isFixedDOF[d]) dbc_dofs.push_back((int)d);
  }

  // ---- AD width (assemblyManager_construct.hpp:127-171)
  const int max_ndr = dofs.ndof_elem;
  if (max_ndr <= 2) type_AD = 2;
  else if (max_ndr <= 4) type_AD = 4;
  else if (max_ndr <= 8) type_AD = 8;
  else if (max_ndr <= 16) type_AD = 16;
  else if (max_ndr <= 18) type_AD = 18;
  else if (max_ndr <= 24) type_AD = 24;
  else if (max_ndr <= 32) type_AD = 32;
  else type_AD = -1;
  if (!s.getb("Solver/enable autotune", true)) type_AD = -1;
  if (max_ndr > ORACLE_MAXDERIVS) throw std::runtime_error("oracle: element dofs exceed MAXDERIVS");

  // ---- groups (createGroups: sequential chunks of `workset size`)
  const int nv = mesh.topo.nverts, dim = mesh.dim, ndofE = dofs.ndof_elem;
  auto fill_group = [&](Group& g, const std::vector<int>& ids) {
    g.numElem = (int)ids.size();
    g.elem_ids = ids;
    g.LIDs.resize((size_t)g.numElem * ndofE);
    g.nodes.resize((size_t)g.numElem * nv * dim);
    for (int e = 0; e < g.numElem; ++e) {
      std::memcpy(&g.LIDs[(size_t)e * ndofE], &dofs.lids[(size_t)ids[e] * ndofE], ndofE * sizeof(int));
      for (int n = 0; n < nv; ++n)
        for (int d = 0; d < dim; ++d) g.nodes[((size_t)e * nv + n) * dim + d] = mesh.nodes[(size_t)mesh.conn[(size_t)ids[e] * nv + n] * dim + d];
    }
  };
  for (int prog = 0; prog < mesh.num_elems; prog += workset_size) {
    std::vector<int> ids;
    for (int e = prog; e < std::min(prog + workset_size, mesh.num_elems); ++e) ids.push_back(e);
    groups.emplace_back();
    fill_group(groups.back(), ids);
    computeGroupBasis(groups.back(), false);
  }
  if (assemble_boundary_terms) {
    for (int sd = 0; sd < nsides; ++sd) {
      std::vector<int> ids;
      const int axis = sd / 2, hiside = sd % 2;
      for (int e = 0; e < mesh.num_elems; ++e) {
        int ijk[3];
        mesh.elem_ijk(e, ijk);
        if (ijk[axis] == (hiside ? mesh.n[axis] - 1 : 0)) ids.push_back(e);
      }
      for (size_t prog = 0; prog < ids.size(); prog += workset_size) {
        std::vector<int> chunk(ids.begin() + prog, ids.begin() + std::min(prog + (size_t)workset_size, ids.size()));
        boundary_groups.emplace_back();
        Group& g = boundary_groups.back();
        g.sideset = sd; g.local_side = mesh.local_side(sd); g.sidename = mesh.side_names[sd];
        fill_group(g, chunk);
        computeGroupBasis(g, true);
      }
    }
  }

  // ---- engines: ScalarT worksets always, one AD width chosen by type_AD (importPhysicsAD)
  eng_scalar = make_engine<double>(*this);
  switch (type_AD) {
    case 2: eng_ad = make_engine<AD2>(*this); break;
    case 4: eng_ad = make_engine<AD4>(*this); break;
    case 8: eng_ad = make_engine<AD8>(*this); break;
    case 16: eng_ad = make_engine<AD16>(*this); break;
    case 18: eng_ad = make_engine<AD18>(*this); break;
    case 24: eng_ad = make_engine<AD24>(*this); break;
    case 32: eng_ad = make_engine<AD32>(*this); break;
    default: eng_ad = make_engine<AD>(*this); break;
  }
}

}  // namespace oracle
