// Host side of the expression compiler (see expr.hpp for the reference map).
#include "expr.hpp"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <functional>
#include <sstream>

#include "../../include/mrhyde_b200.h"

namespace mrhyde_b200 {

namespace {

const double kPi = 3.141592653589793238463;  // src/preferences.hpp:70

// Interpreter::isScalar (interpreter.cpp:17-57): digits, one '.', one 'e'/'E', sign first or after the exponent
bool looks_scalar(const std::string& s) {
  bool isnum = true;
  int ndots = 0, nexp = 0;
  for (size_t k = 0; k < s.size(); ++k) {
    const char ch = s[k];
    if (std::isdigit((unsigned char)ch)) continue;
    if (ch == '.') { if (ndots++ > 0) isnum = false; }
    else if (ch == 'e' || ch == 'E') { if (nexp++ > 0) isnum = false; }
    else if (ch == '+' || ch == '-') { if (k > 0 && !(s[k - 1] == 'e' || s[k - 1] == 'E')) isnum = false; }
    else isnum = false;
  }
  return isnum;
}

const char* const kOps[] = {"sin", "cos", "exp", "log", "tan", "abs", "max", "min", "mean", "emax", "emin", "emean", "sqrt", "sinh", "cosh"};

// op(arg) recognition, interpreter.cpp:360-446
bool match_operator(const std::string& s, std::string& oper, std::string& arg) {
  for (const char* opname : kOps) {
    const std::string op(opname);
    const size_t L = op.size();
    if (s.size() < L || s.compare(0, L, op) != 0) continue;
    if (s.size() <= L + 1 || s[L] != '(' || s.back() != ')') continue;
    bool inner_close = false;
    for (size_t j = L + 1; j + 1 < s.size(); ++j) if (s[j] == ')') inner_close = true;
    if (inner_close) continue;
    oper = op;
    arg = s.substr(L + 1, s.size() - L - 2);
    return true;
  }
  return false;
}

typedef std::vector<std::pair<std::string, std::string>> Pieces;  // (op, sub-expression)

// Interpreter::split (interpreter.cpp:63-352)
Pieces split_expression(std::string s) {
  Pieces out;
  if (!s.empty() && s[0] == '-') s = "0.0" + s;
  if (s.empty()) return out;
  if (s.size() == 1) { out.push_back({"", s}); return out; }
  size_t num_pm = 0, num_mdp = 0, num_pow = 0;
  int paren = 0;
  for (char ch : s) {
    if (ch == '(') ++paren;
    else if (ch == ')') --paren;
    else if (paren == 0) {
      if (ch == '+' || ch == '-') ++num_pm;
      if (ch == '*' || ch == '/' || ch == '<' || ch == '>') ++num_mdp;
      if (ch == '^') ++num_pow;
    }
  }
  if (paren > 0) throw ExprError(MRHYDE_B200_ERR_PARSE, "Error: MrHyDE found an unclosed parenthesis in: " + s);
  if (paren < 0) throw ExprError(MRHYDE_B200_ERR_PARSE, "Error: MrHyDE found an extra parenthesis in: " + s);
  std::string cur, curop;
  if (num_pm > 0) {
    for (size_t i = 0; i < s.size(); ++i) {
      const char ch = s[i];
      if (ch == ' ' || ch == '=') {}
      else if (ch == '(') { ++paren; cur += ch; }
      else if (ch == ')') { --paren; cur += ch; }
      else if (paren == 0 && (ch == '+' || ch == '-') && !cur.empty()) {
        out.push_back({curop, cur});
        cur.clear();
        curop = (ch == '+') ? "plus" : "minus";
      } else cur += ch;
    }
    if (!cur.empty()) out.push_back({curop, cur});
  } else if (num_mdp > 0) {
    for (size_t i = 0; i < s.size(); ++i) {
      const char ch = s[i];
      if (ch == ' ') {}
      else if (ch == '(') { ++paren; cur += ch; }
      else if (ch == ')') { --paren; cur += ch; }
      else if (paren == 0 && (ch == '*' || ch == '/' || ch == '<' || ch == '>')) {
        out.push_back({curop, cur});
        cur.clear();
        if (ch == '*') curop = "times";
        else if (ch == '/') curop = "divide";
        else {
          const bool eq = (i + 1 < s.size() && s[i + 1] == '=');
          curop = (ch == '<') ? (eq ? "lte" : "lt") : (eq ? "gte" : "gt");
          if (eq) ++i;
        }
      } else cur += ch;
    }
    out.push_back({curop, cur});
  } else if (num_pow > 0) {
    for (size_t i = 0; i < s.size(); ++i) {
      const char ch = s[i];
      if (ch == '(') { ++paren; cur += ch; }
      else if (ch == ')') { --paren; cur += ch; }
      else if (paren == 0 && ch == '^') { out.push_back({curop, cur}); cur.clear(); curop = "power"; }
      else cur += ch;
    }
    out.push_back({curop, cur});
  } else if (s.front() == '(' && s.back() == ')') {
    out.push_back({"", s.substr(1, s.size() - 2)});
  } else {
    // name(args) where isOperator refused the string because args contain parentheses (interpreter.cpp:316-341):
    // the reference makes `name` the op of a single dependency, e.g. exp(-(a+b)^2) -> exp applied to -(a+b)^2
    size_t pindex = std::string::npos;
    for (size_t k = 1; k + 1 < s.size(); ++k) if (s[k] == '(') { pindex = k; break; }
    if (pindex != std::string::npos && s.back() == ')') {
      const std::string name = s.substr(0, pindex);
      static const char* const unary[] = {"sin", "cos", "tan", "exp", "log", "abs", "sqrt", "sinh", "cosh"};
      bool known = false;
      for (const char* u : unary) known = known || name == u;
      if (!known) throw ExprError(MRHYDE_B200_ERR_PARSE, "Error: MrHyDE was not able to decompose or find: " + s);
      out.push_back({name, s.substr(pindex + 1, s.size() - pindex - 2)});
    }
  }
  return out;
}

bool is_reduction(const std::string& op) { return op == "emax" || op == "emin" || op == "emean"; }

double apply_binary(const std::string& op, double a, double b) {  // functionManager_evaluate.hpp:237-560
  if (op == "plus") return a + b;
  if (op == "minus") return a + (-b);
  if (op == "times") return a * b;
  if (op == "divide") return a / b;
  if (op == "power") return std::pow(a, b);
  if (op == "lt") return a < b ? 1.0 : 0.0;
  if (op == "lte") return a <= b ? 1.0 : 0.0;
  if (op == "gt") return a > b ? 1.0 : 0.0;
  if (op == "gte") return a >= b ? 1.0 : 0.0;
  if (op == "max") return b > a ? b : a;
  if (op == "min") return b < a ? b : a;
  if (op == "mean") return 0.5 * a + 0.5 * b;
  throw ExprError(MRHYDE_B200_ERR_UNSUPPORTED, "expression operator not supported on the device path: " + op);
}
double apply_unary(const std::string& op, double a) {
  if (op == "sin") return std::sin(a);
  if (op == "cos") return std::cos(a);
  if (op == "tan") return std::tan(a);
  if (op == "exp") return std::exp(a);
  if (op == "log") return std::log(a);
  if (op == "abs") return a < 0.0 ? -a : a;
  if (op == "sqrt") return a <= 0.0 ? 0.0 : std::sqrt(a);
  if (op == "sinh") return std::sinh(a);
  if (op == "cosh") return std::cosh(a);
  if (is_reduction(op)) return a;   // scalar forms: a constant argument is its own max / min / mean (functionManager_evaluate.hpp:1311-1319)
  throw ExprError(MRHYDE_B200_ERR_UNSUPPORTED, "expression operator not supported on the device path: " + op);
}
bool is_unary(const std::string& op) {
  return is_reduction(op) || op == "sin" || op == "cos" || op == "tan" || op == "exp" || op == "log" || op == "abs" || op == "sqrt" || op == "sinh" || op == "cosh";
}
uint8_t binary_code(const std::string& op) {
  if (op == "plus") return OP_ADD;
  if (op == "minus") return OP_SUB;
  if (op == "times") return OP_MUL;
  if (op == "divide") return OP_DIV;
  if (op == "power") return OP_POW;
  if (op == "lt") return OP_LT;
  if (op == "lte") return OP_LTE;
  if (op == "gt") return OP_GT;
  if (op == "gte") return OP_GTE;
  if (op == "max") return OP_MAX;
  if (op == "min") return OP_MIN;
  if (op == "mean") return OP_MEAN;
  throw ExprError(MRHYDE_B200_ERR_UNSUPPORTED, "expression operator not supported on the device path: " + op);
}
uint8_t unary_code(const std::string& op) {
  if (op == "sin") return OP_SIN;
  if (op == "cos") return OP_COS;
  if (op == "tan") return OP_TAN;
  if (op == "exp") return OP_EXP;
  if (op == "log") return OP_LOG;
  if (op == "abs") return OP_ABS;
  if (op == "sqrt") return OP_SQRT;
  if (op == "sinh") return OP_SINH;
  if (op == "emax") return OP_EMAX;
  if (op == "emin") return OP_EMIN;
  if (op == "emean") return OP_EMEAN;
  return OP_COSH;
}

}  // namespace

int FunctionSet::build(const std::string& expr, std::vector<Node>& nodes, std::set<std::string>& active) const {
  const int me = (int)nodes.size();
  nodes.emplace_back();
  // 1. solution fields of the workset (AD data): a coefficient that depends on the state
  for (auto& f : soln_fields_)
    if (expr == f) {
      auto slot = soln_slots_.find(expr);
      if (slot == soln_slots_.end())
        throw ExprError(MRHYDE_B200_ERR_UNSUPPORTED, "solution-dependent coefficient '" + expr + "' has no device kernel in this build");
      nodes[me].kind = Node::VAR; nodes[me].var = EXPR_STATE0 + slot->second;   // functionManager_create.hpp: a workset solution field (AD data)
      return me;
    }
  // 2. scalar fields of the workset (x, y, z; n[x].. on sides)
  for (size_t j = 0; j < scalar_fields_.size(); ++j)
    if (!scalar_fields_[j].empty() && expr == scalar_fields_[j]) { nodes[me].kind = Node::VAR; nodes[me].var = (int)j; return me; }
  // 3. another function of the forest
  auto it = funcs_.find(expr);
  if (it != funcs_.end()) {
    if (active.count(expr)) throw ExprError(MRHYDE_B200_ERR_PARSE, "Error: MrHyDE detected a cyclic graph in: " + expr);
    active.insert(expr);
    const int sub = build(it->second, nodes, active);
    active.erase(expr);
    nodes[me].kind = Node::CHAIN;
    nodes[me].deps.push_back({"", sub});
    return me;
  }
  // 4. plain number
  if (!expr.empty() && looks_scalar(expr)) {
    try { nodes[me].value = std::stod(expr); } catch (...) { throw ExprError(MRHYDE_B200_ERR_PARSE, "Error: MrHyDE was not able to decompose or find: " + expr); }
    nodes[me].kind = Node::CONST;
    return me;
  }
  // 5. known variables
  if (expr == "t") { nodes[me].kind = Node::VAR; nodes[me].var = 3; return me; }
  if (expr == "pi") { nodes[me].kind = Node::CONST; nodes[me].value = kPi; return me; }
  // 6. op(arg) / op(arg1,arg2)
  std::string oper, arg;
  if (match_operator(expr, oper, arg)) {
    if (arg.empty()) throw ExprError(MRHYDE_B200_ERR_PARSE, "Error: MrHyDE was not able to decompose or find: " + expr);
    size_t comma = std::string::npos;
    for (size_t i = 0; i + 1 < arg.size(); ++i) if (arg[i] == ',') comma = i;
    nodes[me].kind = Node::CHAIN;
    if (comma != std::string::npos) {
      const int a = build(arg.substr(0, comma), nodes, active);
      const int b = build(arg.substr(comma + 1), nodes, active);
      nodes[me].deps.push_back({"", a});
      nodes[me].deps.push_back({oper, b});
    } else {
      const int a = build(arg, nodes, active);
      nodes[me].deps.push_back({oper, a});
    }
    return me;
  }
  // 7. split
  Pieces pieces = split_expression(expr);
  if (pieces.empty()) throw ExprError(MRHYDE_B200_ERR_PARSE, "Error: MrHyDE was not able to decompose or find: " + expr);
  if (pieces.size() == 1 && pieces[0].first.empty() && pieces[0].second == expr)
    throw ExprError(MRHYDE_B200_ERR_PARSE, "Error: MrHyDE was not able to decompose or find: " + expr);
  nodes[me].kind = Node::CHAIN;
  for (auto& pc : pieces) {
    const int sub = build(pc.second, nodes, active);
    nodes[me].deps.push_back({pc.first, sub});
  }
  return me;
}

// Folds branches whose dependencies are all constant (reference: evaluated once at set-up).
bool FunctionSet::fold(std::vector<Node>& nodes, int idx) {
  Node& n = nodes[idx];
  if (n.kind == Node::CONST) return true;
  if (n.kind == Node::VAR) return false;
  bool all = true;
  for (auto& d : n.deps) all = fold(nodes, d.second) && all;
  if (!all) return false;
  double acc = 0.0;
  for (size_t k = 0; k < n.deps.size(); ++k) {
    const double v = nodes[n.deps[k].second].value;
    const std::string& op = n.deps[k].first;
    if (op.empty()) acc = v;
    else if (is_unary(op)) acc = apply_unary(op, v);
    else acc = apply_binary(op, acc, v);
  }
  n.kind = Node::CONST;
  n.value = acc;
  n.deps.clear();
  return true;
}

void FunctionSet::emit(const std::vector<Node>& nodes, int idx, const std::function<void(uint8_t, double)>& push_op, int& depth, int& maxdepth) {
  const Node& n = nodes[idx];
  if (n.kind == Node::CONST) { push_op(OP_PUSHC, n.value); maxdepth = std::max(maxdepth, ++depth); return; }
  if (n.kind == Node::VAR) { push_op(OP_PUSHV, (double)n.var); maxdepth = std::max(maxdepth, ++depth); return; }
  for (size_t k = 0; k < n.deps.size(); ++k) {
    const std::string& op = n.deps[k].first;
    const Node& d = nodes[n.deps[k].second];
    if (op.empty()) {
      // "data = dep": the splitter only produces this for the first dependency
      if (k > 0) throw ExprError(MRHYDE_B200_ERR_PARSE, "Error: assignment in a chained position");
      emit(nodes, n.deps[k].second, push_op, depth, maxdepth);
    } else if (is_unary(op)) {
      if (k > 0) throw ExprError(MRHYDE_B200_ERR_PARSE, "Error: unary operator in a chained position");
      if (is_reduction(op)) {
        // the argument is bracketed so that the evaluator can re-run it at the other points of the element; the closing op carries the
        // distance back to its OP_EBEGIN (relative: programs are concatenated into one table)
        int count = 0;
        auto counting = [&](uint8_t o, double c) { ++count; push_op(o, c); };
        push_op(OP_EBEGIN, 0.0);
        emit(nodes, n.deps[k].second, counting, depth, maxdepth);
        push_op(unary_code(op), (double)(count + 1));
      } else {
        emit(nodes, n.deps[k].second, push_op, depth, maxdepth);
        push_op(unary_code(op), 0.0);
      }
    } else {
      if (k == 0) throw ExprError(MRHYDE_B200_ERR_PARSE, "Error: binary operator without a left operand");
      const uint8_t code = binary_code(op);
      if (d.kind == Node::CONST && code >= OP_ADD && code <= OP_POW) push_op((uint8_t)(OP_ADDC + (code - OP_ADD)), d.value);
      else if (d.kind == Node::VAR && code >= OP_ADD && code <= OP_DIV) push_op((uint8_t)(OP_ADDV + (code - OP_ADD)), (double)d.var);
      else {
        emit(nodes, n.deps[k].second, push_op, depth, maxdepth);
        push_op(code, 0.0);
        --depth;
      }
    }
  }
}

ExprProgram FunctionSet::compile(const std::string& name) const {
  auto it = funcs_.find(name);
  if (it == funcs_.end()) throw ExprError(MRHYDE_B200_ERR_INVALID, "function not registered: " + name);
  std::vector<Node> nodes;
  std::set<std::string> active;
  active.insert(name);
  const int root = build(it->second, nodes, active);
  ExprProgram p;
  if (fold(nodes, root)) {
    p.is_const = 1;
    p.cval = nodes[root].value;
    p.n = 1;
    p.op[0] = OP_PUSHC; p.c[0] = p.cval; p.op[1] = OP_END;
    return p;
  }
  p.is_const = 0;
  int depth = 0, maxdepth = 0;
  auto push_op = [&](uint8_t op, double c) {
    if (p.n >= EXPR_MAXOPS - 1) throw ExprError(MRHYDE_B200_ERR_UNSUPPORTED, "expression too long for the device evaluator");
    p.op[p.n] = op; p.c[p.n] = c; ++p.n;
  };
  emit(nodes, root, push_op, depth, maxdepth);
  for (int i = 0; i < p.n; ++i)
    if (p.op[i] == OP_EBEGIN) throw ExprError(MRHYDE_B200_ERR_UNSUPPORTED, "element reduction (emax / emin / emean) in '" + name + "' needs the general path");
  for (int i = 0; i < p.n; ++i)
    if ((p.op[i] == OP_PUSHV || (p.op[i] >= OP_ADDV && p.op[i] <= OP_DIVV)) && (int)p.c[i] >= EXPR_STATE0)
      throw ExprError(MRHYDE_B200_ERR_UNSUPPORTED, "solution-dependent coefficient '" + name + "' needs the general path");
  p.op[p.n] = OP_END;
  if (maxdepth > EXPR_MAXSTACK) throw ExprError(MRHYDE_B200_ERR_UNSUPPORTED, "expression nests too deeply for the device evaluator: " + name);
  return p;
}


// ---- C++ code generation (plan-specialised kernels) ---------------------------------------------------------
namespace {
std::string hexfloat(double v) {
  char buf[64];
  if (v != v) return "(0.0/0.0)";
  if (v == HUGE_VAL) return "(1.0/0.0)";
  if (v == -HUGE_VAL) return "(-1.0/0.0)";
  std::snprintf(buf, sizeof(buf), "%a", v);
  return std::string("(") + buf + ")";
}
}  // namespace

// Combines the generated code of a CHAIN node's dependencies in the reference's left-to-right order.
std::string FunctionSet::gen_chain(const Node& n, const std::function<std::string(int)>& child) {
  std::string acc;
  for (size_t k = 0; k < n.deps.size(); ++k) {
    const std::string& op = n.deps[k].first;
    const std::string d = child(n.deps[k].second);
    if (op.empty()) {
      if (k > 0) throw ExprError(MRHYDE_B200_ERR_PARSE, "Error: assignment in a chained position");
      acc = d;
    } else if (is_unary(op)) {
      if (k > 0) throw ExprError(MRHYDE_B200_ERR_PARSE, "Error: unary operator in a chained position");
      if (is_reduction(op)) throw ExprError(MRHYDE_B200_ERR_UNSUPPORTED, "element reduction (" + op + ") needs the general path");
      if (op == "abs") acc = "mrh_abs(" + d + ")";
      else if (op == "sqrt") acc = "mrh_sqrt(" + d + ")";
      else if (op == "sin" || op == "cos") acc = "mrh_" + op + "(" + d + ")";   // table-free kernels (jit prelude)
      else acc = op + "(" + d + ")";
    } else {
      if (k == 0) throw ExprError(MRHYDE_B200_ERR_PARSE, "Error: binary operator without a left operand");
      if (op == "plus") acc = "(" + acc + " + " + d + ")";
      else if (op == "minus") acc = "(" + acc + " + (-" + d + "))";
      else if (op == "times") acc = "(" + acc + " * " + d + ")";
      else if (op == "divide") acc = "(" + acc + " / " + d + ")";
      else if (op == "power") acc = "pow(" + acc + ", " + d + ")";
      else if (op == "lt") acc = "((" + acc + " < " + d + ") ? 1.0 : 0.0)";
      else if (op == "lte") acc = "((" + acc + " <= " + d + ") ? 1.0 : 0.0)";
      else if (op == "gt") acc = "((" + acc + " > " + d + ") ? 1.0 : 0.0)";
      else if (op == "gte") acc = "((" + acc + " >= " + d + ") ? 1.0 : 0.0)";
      else if (op == "max") acc = "mrh_max(" + acc + ", " + d + ")";
      else if (op == "min") acc = "mrh_min(" + acc + ", " + d + ")";
      else if (op == "mean") acc = "(0.5 * " + acc + " + 0.5 * " + d + ")";
      else throw ExprError(MRHYDE_B200_ERR_UNSUPPORTED, "expression operator not supported on the device path: " + op);
    }
  }
  return acc;
}

std::string FunctionSet::gen(const std::vector<Node>& nodes, int idx) {
  static const char* vars[] = {"x", "y", "z", "t", "nx", "ny", "nz"};
  const Node& n = nodes[idx];
  if (n.kind == Node::CONST) return hexfloat(n.value);
  if (n.kind == Node::VAR) {
    if (n.var < 0 || n.var > 6) throw ExprError(MRHYDE_B200_ERR_UNSUPPORTED, "expression variable has no generated-code name");
    return vars[n.var];
  }
  return gen_chain(n, [&](int c) { return gen(nodes, c); });
}

std::string FunctionSet::codegen(const std::string& name) const {
  auto it = funcs_.find(name);
  if (it == funcs_.end()) throw ExprError(MRHYDE_B200_ERR_INVALID, "function not registered: " + name);
  std::vector<Node> nodes;
  std::set<std::string> active;
  active.insert(name);
  const int root = build(it->second, nodes, active);
  fold(nodes, root);
  return gen(nodes, root);
}

std::string FunctionSet::codegen_tensor(const std::string& name, const std::string& fname, int nq, const int nqa[3], const int* qidx, int cache_axes, int* cache_n, int shared_axes,
                                        int* shared_n) const {
  auto it = funcs_.find(name);
  if (it == funcs_.end()) throw ExprError(MRHYDE_B200_ERR_INVALID, "function not registered: " + name);
  std::vector<Node> nodes;
  std::set<std::string> active;
  active.insert(name);
  const int root = build(it->second, nodes, active);
  fold(nodes, root);
  std::function<int(int)> mask = [&](int idx) -> int {
    const Node& n = nodes[idx];
    if (n.kind == Node::CONST) return 0;
    if (n.kind == Node::VAR) return n.var >= 0 && n.var <= 3 ? (1 << n.var) : 16;
    int m = 0;
    for (auto& d : n.deps) m |= mask(d.second);
    return m;
  };
  std::string decls;
  int ntemp = 0, ncache = 0, nshared = 0;
  std::string shared_body;   // body of <fname>_shared
  static const char* axis_arr[3] = {"xs", "ys", "zs"};
  static const char* axis_var[3] = {"x", "y", "z"};
  static const char* axis_tok[3] = {"@x", "@y", "@z"};
  std::function<std::string(int)> gen_t = [&](int idx) -> std::string {
    const Node& n = nodes[idx];
    if (n.kind == Node::CONST) return hexfloat(n.value);
    const int m = mask(idx);
    if (m & 16) throw ExprError(MRHYDE_B200_ERR_UNSUPPORTED, "volume coefficient depends on a side-only field");
    if (n.kind == Node::VAR) {
      if (n.var == 3) return "t";
      return std::string(axis_arr[n.var]) + "[" + axis_tok[n.var] + "]";
    }
    for (int a = 0; a < 3; ++a)
      if (m == (1 << a)) {  // depends on one coordinate only: evaluate once per distinct coordinate value
        const std::string h = "h" + std::to_string(ntemp++);
        decls += "  double " + h + "[" + std::to_string(std::max(1, nqa[a])) + "];\n";
        const std::string loop = "_Pragma(\"unroll\") for (int i = 0; i < " + std::to_string(nqa[a]) + "; ++i) ";
        const std::string eval = "const double " + std::string(axis_var[a]) + " = " + axis_arr[a] + "[i]; " + h + "[i] = " + gen(nodes, idx) + ";";
        if ((cache_axes >> a) & 1) {
          const std::string off = std::to_string(ncache);
          ncache += nqa[a];
          decls += "  if ((reuse >> " + std::to_string(a) + ") & 1) { " + loop + h + "[i] = cache[" + off + " + i]; }\n";
          decls += "  else { " + loop + "{ " + eval + " cache[" + off + " + i] = " + h + "[i]; } }\n";
        } else if ((shared_axes >> a) & 1) {
          const std::string off = std::to_string(nshared);
          nshared += nqa[a];
          decls += "  if ((reuse >> " + std::to_string(4 + a) + ") & 1) { " + loop + h + "[i] = shv[" + off + " + i]; }\n";
          decls += "  else { " + loop + "{ " + eval + " } }\n";
          shared_body += "  { double " + h + "[" + std::to_string(std::max(1, nqa[a])) + "]; " + loop + "{ " + eval + " vals[" + off + " + i] = " + h + "[i]; } }\n";
        } else {
          decls += "  " + loop + "{ " + eval + " }\n";
        }
        return h + "[" + axis_tok[a] + "]";
      }
    if (m == 8) {
      const std::string h = "h" + std::to_string(ntemp++);
      decls += "  const double " + h + " = " + gen(nodes, idx) + ";\n";
      return h;
    }
    return gen_chain(n, gen_t);
  };
  const std::string body = gen_t(root);
  if (cache_n) *cache_n = ncache;
  if (shared_n) *shared_n = nshared;
  std::string o;
  if (nshared > 0) o += "__device__ __forceinline__ void " + fname + "_shared(const double* xs, const double* ys, const double* zs, double t, double* vals) {\n" + shared_body + "}\n";
  o += "__device__ __forceinline__ void " + fname + "(const double* xs, const double* ys, const double* zs, double t, double* out" +
       (ncache + nshared > 0 ? ", double* cache, const int reuse" : "") + (nshared > 0 ? ", const double* shv" : "") + ") {\n";
  o += decls;
  for (int q = 0; q < nq; ++q) {
    std::string e = body;
    for (int a = 0; a < 3; ++a) {
      const std::string tok = axis_tok[a], rep = std::to_string(qidx[q * 3 + a]);
      size_t p = 0;
      while ((p = e.find(tok, p)) != std::string::npos) { e.replace(p, tok.size(), rep); p += rep.size(); }
    }
    o += "  out[" + std::to_string(q) + "] = " + e + ";\n";
  }
  o += "}\n";
  return o;
}

LongProgram FunctionSet::compile_long(const std::string& name) const {
  auto it = funcs_.find(name);
  if (it == funcs_.end()) throw ExprError(MRHYDE_B200_ERR_INVALID, "function not registered: " + name);
  std::vector<Node> nodes;
  std::set<std::string> active;
  active.insert(name);
  const int root = build(it->second, nodes, active);
  LongProgram p;
  if (fold(nodes, root)) {
    p.is_const = true;
    p.cval = nodes[root].value;
    return p;
  }
  int depth = 0, maxdepth = 0;
  auto push_op = [&](uint8_t op, double c) { p.op.push_back(op); p.c.push_back(c); };
  emit(nodes, root, push_op, depth, maxdepth);
  if (maxdepth > 16) throw ExprError(MRHYDE_B200_ERR_UNSUPPORTED, "expression nests too deeply for the device evaluator: " + name);
  int open = 0;
  for (size_t i = 0; i < p.op.size(); ++i) {
    if ((p.op[i] == OP_PUSHV || (p.op[i] >= OP_ADDV && p.op[i] <= OP_DIVV)) && (int)p.c[i] >= EXPR_STATE0) p.uses_state = true;
    if (p.op[i] == OP_EBEGIN) { p.uses_reduction = true; if (++open > 1) throw ExprError(MRHYDE_B200_ERR_UNSUPPORTED, "nested element reductions in '" + name + "'"); }
    if (p.op[i] == OP_EMAX || p.op[i] == OP_EMIN || p.op[i] == OP_EMEAN) --open;
  }
  if (p.uses_reduction && p.uses_state)
    throw ExprError(MRHYDE_B200_ERR_UNSUPPORTED, "'" + name + "': an element reduction over a solution-dependent expression has no device kernel in this build");
  return p;
}

double FunctionSet::eval_host(const ExprProgram& p, const double* vars) {
  double st[EXPR_MAXSTACK + 1];
  int sp = -1;
  for (int i = 0; i < p.n; ++i) {
    const double c = p.c[i];
    switch (p.op[i]) {
      case OP_PUSHC: st[++sp] = c; break;
      case OP_PUSHV: st[++sp] = vars[(int)c]; break;
      case OP_ADD: st[sp - 1] = st[sp - 1] + st[sp]; --sp; break;
      case OP_SUB: st[sp - 1] = st[sp - 1] + (-st[sp]); --sp; break;
      case OP_MUL: st[sp - 1] = st[sp - 1] * st[sp]; --sp; break;
      case OP_DIV: st[sp - 1] = st[sp - 1] / st[sp]; --sp; break;
      case OP_POW: st[sp - 1] = std::pow(st[sp - 1], st[sp]); --sp; break;
      case OP_LT: st[sp - 1] = st[sp - 1] < st[sp] ? 1.0 : 0.0; --sp; break;
      case OP_LTE: st[sp - 1] = st[sp - 1] <= st[sp] ? 1.0 : 0.0; --sp; break;
      case OP_GT: st[sp - 1] = st[sp - 1] > st[sp] ? 1.0 : 0.0; --sp; break;
      case OP_GTE: st[sp - 1] = st[sp - 1] >= st[sp] ? 1.0 : 0.0; --sp; break;
      case OP_MAX: st[sp - 1] = st[sp] > st[sp - 1] ? st[sp] : st[sp - 1]; --sp; break;
      case OP_MIN: st[sp - 1] = st[sp] < st[sp - 1] ? st[sp] : st[sp - 1]; --sp; break;
      case OP_MEAN: st[sp - 1] = 0.5 * st[sp - 1] + 0.5 * st[sp]; --sp; break;
      case OP_ADDC: st[sp] = st[sp] + c; break;
      case OP_SUBC: st[sp] = st[sp] + (-c); break;
      case OP_MULC: st[sp] = st[sp] * c; break;
      case OP_DIVC: st[sp] = st[sp] / c; break;
      case OP_POWC: st[sp] = std::pow(st[sp], c); break;
      case OP_ADDV: st[sp] = st[sp] + vars[(int)c]; break;
      case OP_SUBV: st[sp] = st[sp] + (-vars[(int)c]); break;
      case OP_MULV: st[sp] = st[sp] * vars[(int)c]; break;
      case OP_DIVV: st[sp] = st[sp] / vars[(int)c]; break;
      case OP_SIN: st[sp] = std::sin(st[sp]); break;
      case OP_COS: st[sp] = std::cos(st[sp]); break;
      case OP_TAN: st[sp] = std::tan(st[sp]); break;
      case OP_EXP: st[sp] = std::exp(st[sp]); break;
      case OP_LOG: st[sp] = std::log(st[sp]); break;
      case OP_ABS: st[sp] = st[sp] < 0.0 ? -st[sp] : st[sp]; break;
      case OP_SQRT: st[sp] = st[sp] <= 0.0 ? 0.0 : std::sqrt(st[sp]); break;
      case OP_SINH: st[sp] = std::sinh(st[sp]); break;
      case OP_COSH: st[sp] = std::cosh(st[sp]); break;
      default: break;
    }
  }
  return sp >= 0 ? st[sp] : 0.0;
}

std::string FunctionSet::disassemble(const ExprProgram& p) {
  static const char* names[] = {"end", "pushc", "pushv", "add", "sub", "mul", "div", "pow", "lt", "lte", "gt", "gte", "max", "min", "mean",
                                "addc", "subc", "mulc", "divc", "powc", "addv", "subv", "mulv", "divv",
                                "sin", "cos", "tan", "exp", "log", "abs", "sqrt", "sinh", "cosh", "ebegin", "emax", "emin", "emean"};
  std::ostringstream os;
  os.precision(17);
  if (p.is_const) { os << "const " << p.cval; return os.str(); }
  for (int i = 0; i < p.n; ++i) {
    os << names[p.op[i]];
    const uint8_t o = p.op[i];
    if (o == OP_PUSHC || (o >= OP_ADDC && o <= OP_POWC)) os << " " << p.c[i];
    if (o == OP_PUSHV || (o >= OP_ADDV && o <= OP_DIVV)) os << " v" << (int)p.c[i];
    os << "; ";
  }
  return os.str();
}

}  // namespace mrhyde_b200
