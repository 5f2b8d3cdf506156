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
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

namespace mrhyde_b200 {

enum ExprOp : uint8_t {
  OP_END = 0,
  OP_PUSHC,   // push constant
  OP_PUSHV,   // push variable (index in c: 0 x, 1 y, 2 z, 3 t, 4.. extra inputs)
  OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW, OP_LT, OP_LTE, OP_GT, OP_GTE, OP_MAX, OP_MIN, OP_MEAN,   // binary: a = a op b
  OP_ADDC, OP_SUBC, OP_MULC, OP_DIVC, OP_POWC,        // binary with constant right operand
  OP_ADDV, OP_SUBV, OP_MULV, OP_DIVV,                 // binary with variable right operand
  OP_SIN, OP_COS, OP_TAN, OP_EXP, OP_LOG, OP_ABS, OP_SQRT, OP_SINH, OP_COSH,  // unary on top of stack
};

constexpr int EXPR_MAXOPS = 56;
constexpr int EXPR_MAXSTACK = 8;
constexpr int EXPR_NVARS = 10;  // x y z t n[x] n[y] n[z] + spare

struct ExprProgram {  // POD, copied into kernel parameters
  int32_t n = 0;
  int32_t is_const = 1;
  double cval = 0.0;
  uint8_t op[EXPR_MAXOPS] = {0};
  double c[EXPR_MAXOPS] = {0};
};

struct ExprError : std::runtime_error {
  int code;
  ExprError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// A named set of functions at one location ("ip" / "side ip"), i.e. one reference Forest.
class FunctionSet {
 public:
  void set(const std::string& name, const std::string& expr) { funcs_[name] = expr; }
  bool has(const std::string& name) const { return funcs_.count(name) > 0; }
  // names of fields the workset would provide (solution fields); referencing one is "unsupported"
  void set_solution_fields(const std::vector<std::string>& f) { soln_fields_.assign(f.begin(), f.end()); }
  void set_scalar_fields(const std::vector<std::string>& f) { scalar_fields_ = f; }  // index = variable slot
  ExprProgram compile(const std::string& name) const;
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
  static void emit(const std::vector<Node>& nodes, int idx, ExprProgram& p, int& depth, int& maxdepth);
  std::map<std::string, std::string> funcs_;
  std::vector<std::string> soln_fields_;
  std::vector<std::string> scalar_fields_ = {"x", "y", "z"};
};

}  // namespace mrhyde_b200
