// Run-time expression compiler: `Functions:` strings -> flat bytecode evaluated inside the
// assembly kernels at each quadrature point.
//
// Grammar and evaluation order are the reference's (FunctionManager + Interpreter):
//   leaf classification order   src/managers/function/functionManager_create.hpp:78-540
//   operator recognition        src/tools/interpreter.cpp:360-446  (op(arg) / op(arg1,arg2))
//   splitting                   src/tools/interpreter.cpp:63-352   (+,- first; then * / < > <= >=; then ^;
//                                                                   then enclosing parentheses)
//   evaluation                  src/managers/function/functionManager_evaluate.hpp:59-229
//       a branch is "dep0, then op_k applied with dep_k, left to right"; a-b is a += -b;
//       a leading '-' becomes "0.0-..."; sqrt(x<=0) = 0; lt/gt/... yield 1.0/0.0
//   constant branches are folded at set-up, as the reference does (_create.hpp:519-534).
// Instead of one kernel launch per binary op over (elem,pt) arrays, the tree is flattened into a
// small stack program that each thread runs in registers.
#pragma once
#include <cstdint>
#include <functional>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "kernel_abi.h"

namespace mrhyde_b200 {

struct ExprError : std::runtime_error {
  int code;
  ExprError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// Variable-length program for the general kernels (bytecode kept in global memory): same ops, no length limit, stack <= 16.
struct LongProgram {
  bool is_const = false;
  bool uses_reduction = false;   // contains emax / emin / emean: evaluated over all points of the element (general path only)
  bool uses_state = false;   // reads a solution field (variable slots >= EXPR_STATE0): its value depends on the state
  double cval = 0.0;
  std::vector<uint8_t> op;
  std::vector<double> c;
};

// A named set of functions at one location ("ip" / "side ip"), i.e. one reference Forest.
class FunctionSet {
 public:
  void set(const std::string& name, const std::string& expr) { funcs_[name] = expr; }
  bool has(const std::string& name) const { return funcs_.count(name) > 0; }
  // names of fields the workset would provide (solution fields); referencing one is "unsupported"
  void set_solution_fields(const std::vector<std::string>& f) { soln_fields_.assign(f.begin(), f.end()); }
  void set_scalar_fields(const std::vector<std::string>& f) { scalar_fields_ = f; }  // index = variable slot
  // solution fields the evaluator can read (general path): name -> slot; the leaf becomes variable EXPR_STATE0 + slot.  Fields that are
  // listed by set_solution_fields but have no slot stay "unsupported" (the sweep kernel's fixed-size programs)
  void set_solution_slots(const std::map<std::string, int>& m) { soln_slots_ = m; }
  ExprProgram compile(const std::string& name) const;
  LongProgram compile_long(const std::string& name) const;
  // The same tree as a C++ expression in x, y, z, t (and nx, ny, nz on sides) for the plan-specialised (NVRTC)
  // kernels: every reference op is applied in the reference's left-to-right order; constants are hex floats.
  std::string codegen(const std::string& name) const;
  // Definition of `void fname(const double* xs, const double* ys, const double* zs, double t, double* out)` that evaluates
  // the function at the nq points (xs[qidx[q][0]], ys[qidx[q][1]], zs[qidx[q][2]]) of a tensor-product point set with
  // nqa[a] distinct coordinates per axis: sub-trees that depend on one coordinate only are evaluated once per distinct
  // value (they are loop invariants of the reference's point loop); every value equals the pointwise evaluation bit for bit.
  // cache_axes: bit a set = one-coordinate sub-expressions of axis a are kept in a caller-owned array and reused when the caller says the
  // coordinate values did not change (extra parameters `double* cache, int reuse`); *cache_n receives the array length
  // shared_axes: sub-expressions of these axes may instead be read from a caller-supplied array (`const double* shv`, bit 4 + a of `reuse`) that
  // the companion function <fname>_shared(xs, ys, zs, t, vals) fills; *shared_n receives its length
  std::string codegen_tensor(const std::string& name, const std::string& fname, int nq, const int nqa[3], const int* qidx, int cache_axes = 0, int* cache_n = nullptr,
                             int shared_axes = 0, int* shared_n = nullptr) const;
  // human-readable flattened program (tests)
  static std::string disassemble(const ExprProgram& p);
  // reference-style host evaluation of a program (used for constant folding checks in tests)
  static double eval_host(const ExprProgram& p, const double* vars);

 private:
  struct Node {
    enum Kind { CONST, VAR, CHAIN } kind = CONST;
    double value = 0.0;
    int var = 0;
    std::vector<std::pair<std::string, int>> deps;  // (op, node index)
  };
  int build(const std::string& expr, std::vector<Node>& nodes, std::set<std::string>& active) const;
  static bool fold(std::vector<Node>& nodes, int idx);
  static void emit(const std::vector<Node>& nodes, int idx, const std::function<void(uint8_t, double)>& push_op, int& depth, int& maxdepth);
  static std::string gen(const std::vector<Node>& nodes, int idx);
  static std::string gen_chain(const Node& n, const std::function<std::string(int)>& child);
  std::map<std::string, std::string> funcs_;
  std::vector<std::string> soln_fields_;
  std::map<std::string, int> soln_slots_;
  std::vector<std::string> scalar_fields_ = {"x", "y", "z"};
};

}  // namespace mrhyde_b200
