// ORACLE (test infrastructure, never shipped, never on the product path).
//
// CPU restatement of the reference's run-time expression engine:
//   Interpreter::isScalar / split / isOperator      src/tools/interpreter.cpp:17-57, 63-352, 360-446
//   Branch / Tree / Forest                          src/tools/dag.hpp
//   FunctionManager::decomposeFunctions             src/managers/function/functionManager_create.hpp:78-540
//   FunctionManager::checkDepDataType               src/managers/function/functionManager_create.hpp:583-619
//   FunctionManager::evaluate (recursive)           src/managers/function/functionManager_evaluate.hpp:14-229
//   evaluateOp{VToV,SToV,SToS}                      src/managers/function/functionManager_evaluate.hpp:237-1353
//   Vista (view-or-constant result handle)          src/tools/vista.hpp:21-132
// Evaluation order is the reference's: each branch is "dep0, then op_k applied with dep_k
// left to right", `a-b` is `a += -b`, a leading '-' becomes "0.0-...", one pass over the
// whole (elem,pt) array per op.
#pragma once
#include <algorithm>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "sfad.hpp"

namespace oracle {

#define ORACLE_PI 3.141592653589793238463  // src/preferences.hpp:70

template <class T>
struct View2 {  // (dim0, dim1) row-major array, zero-initialised like Kokkos::View
  int n0 = 0, n1 = 0;
  std::vector<T> a;
  View2() {}
  View2(int n0_, int n1_) : n0(n0_), n1(n1_), a((size_t)n0_ * n1_, T(0.0)) {}
  T& operator()(int i, int j) { return a[(size_t)i * n1 + j]; }
  const T& operator()(int i, int j) const { return a[(size_t)i * n1 + j]; }
};

// ---------------------------------------------------------------------------------------
// Vista: what FunctionManager::evaluate hands to the physics kernels (vista.hpp:92-106)
// ---------------------------------------------------------------------------------------
template <class EvalT>
struct Vista {
  bool is_view = false, is_AD = false;
  const View2<EvalT>* vdata = nullptr;
  const View2<double>* vdata_sc = nullptr;
  EvalT sdata = EvalT(0.0);
  EvalT operator()(int e, int pt) const {
    if (is_view) {
      if (is_AD) return (*vdata)(e, pt);
      return EvalT((*vdata_sc)(e, pt));
    }
    return sdata;
  }
};

template <class EvalT>
struct Branch {
  std::string expression;
  bool is_leaf = false, is_decomposed = false, is_func = false, is_view = false, is_AD = false;
  bool is_constant = false, is_workset_data = false, is_time = false, currently_checking = false;
  int func_index = 0, workset_data_index = 0;
  std::vector<int> dep_list;
  std::vector<std::string> dep_ops;
  double data_Sc = 0.0;
  EvalT data = EvalT(0.0);
  View2<EvalT>* viewdata = nullptr;      // may alias workset / other tree storage
  View2<double>* viewdata_Sc = nullptr;
  View2<EvalT> own;
  View2<double> own_Sc;
  Branch() {}
  explicit Branch(const std::string& e) : expression(e) {}
};

template <class EvalT>
struct Tree {
  std::string name, expression;
  std::vector<Branch<EvalT>> branches;
  Vista<EvalT> vista;
};

template <class EvalT>
struct Forest {
  std::string location;
  int dim0 = 1, dim1 = 1;
  std::vector<Tree<EvalT>> trees;
};

// ---------------------------------------------------------------------------------------
// Interpreter (interpreter.cpp)
// ---------------------------------------------------------------------------------------
inline bool interp_isScalar(const std::string& s) {  // interpreter.cpp:17-57
  bool isnum = true;
  int numdots = 0, nume = 0;
  for (size_t k = 0; k < s.length(); k++) {
    if (!isdigit((unsigned char)s[k])) {
      if (s[k] == '.') {
        if (numdots == 0) numdots += 1; else isnum = false;
      } else if (s[k] == 'e' || s[k] == 'E') {
        if (nume == 0) nume += 1; else isnum = false;
      } else if (s[k] == '+' || s[k] == '-') {
        if (k > 0) {
          if (!(s[k - 1] == 'e' || s[k - 1] == 'E')) isnum = false;
        }
      } else {
        isnum = false;
      }
    }
  }
  return isnum;
}

template <class EvalT>
void interp_add_dep(std::vector<Branch<EvalT>>& branches, size_t index, const std::string& expr, const std::string& op) {
  branches.push_back(Branch<EvalT>(expr));
  branches[index].dep_list.push_back((int)branches.size() - 1);
  branches[index].dep_ops.push_back(op);
}

template <class EvalT>
void interp_split(std::vector<Branch<EvalT>>& branches, size_t index) {  // interpreter.cpp:63-352
  std::string s = branches[index].expression;
  if (s.length() > 0 && s[0] == '-') s = "0.0" + s;

  if (s.length() == 0) {
  } else if (s.length() == 1) {
    interp_add_dep(branches, index, std::string(1, s[0]), "");
  } else {
    size_t num_pm = 0, num_mdp = 0, num_pow = 0;
    int paren = 0;
    for (size_t i = 0; i < s.length(); i++) {
      if (s[i] == '(') paren += 1;
      else if (s[i] == ')') paren += -1;
      else if (paren == 0) {
        if (s[i] == '+' || s[i] == '-') num_pm += 1;
        if (s[i] == '*' || s[i] == '/' || s[i] == '<' || s[i] == '>') num_mdp += 1;
        if (s[i] == '^') num_pow += 1;
      }
    }
    paren = 0;
    std::string currbranch = "", currop = "";
    if (num_pm > 0) {
      for (size_t i = 0; i < s.length(); i++) {
        if (s[i] == ' ') {
        } else if (s[i] == '=') {
        } else if (s[i] == '(') { paren += 1; currbranch += s[i]; }
        else if (s[i] == ')') { paren += -1; currbranch += s[i]; }
        else if (paren == 0 && s[i] == '+' && currbranch.length() > 0) {
          interp_add_dep(branches, index, currbranch, currop);
          currbranch = ""; currop = "plus";
        } else if (paren == 0 && s[i] == '-' && currbranch.length() > 0) {
          interp_add_dep(branches, index, currbranch, currop);
          currbranch = ""; currop = "minus";
        } else {
          currbranch += s[i];
        }
        if (i == s.length() - 1 && currbranch.length() > 0) {
          interp_add_dep(branches, index, currbranch, currop);
        }
      }
    } else if (num_mdp > 0) {
      for (size_t i = 0; i < s.length(); i++) {
        if (s[i] == ' ') {
        } else if (s[i] == '(') { paren += 1; currbranch += s[i]; }
        else if (s[i] == ')') { paren += -1; currbranch += s[i]; }
        else if (paren == 0 && s[i] == '*') {
          interp_add_dep(branches, index, currbranch, currop);
          currbranch = ""; currop = "times";
        } else if (paren == 0 && s[i] == '/') {
          interp_add_dep(branches, index, currbranch, currop);
          currbranch = ""; currop = "divide";
        } else if (paren == 0 && s[i] == '<') {
          interp_add_dep(branches, index, currbranch, currop);
          currbranch = "";
          if (i + 1 < s.length() && s[i + 1] == '=') { currop = "lte"; ++i; } else currop = "lt";
        } else if (paren == 0 && s[i] == '>') {
          interp_add_dep(branches, index, currbranch, currop);
          currbranch = "";
          if (i + 1 < s.length() && s[i + 1] == '=') { currop = "gte"; ++i; } else currop = "gt";
        } else {
          currbranch += s[i];
        }
        if (i == s.length() - 1) {
          interp_add_dep(branches, index, currbranch, currop);
        }
      }
    } else if (num_pow > 0) {
      for (size_t i = 0; i < s.length(); i++) {
        if (s[i] == '(') { paren += 1; currbranch += s[i]; }
        else if (s[i] == ')') { paren += -1; currbranch += s[i]; }
        else if (paren == 0 && s[i] == '^') {
          interp_add_dep(branches, index, currbranch, currop);
          currbranch = ""; currop = "power";
        } else {
          currbranch += s[i];
        }
        if (i == s.length() - 1) {
          interp_add_dep(branches, index, currbranch, currop);
        }
      }
    } else {
      if (s[0] == '(' && s[s.length() - 1] == ')') {
        interp_add_dep(branches, index, s.substr(1, s.length() - 2), currop);
      } else {
        bool foundparen = false;
        size_t pindex = 0;
        for (size_t k = 1; k + 1 < s.length(); k++) {
          if (s[k] == '(' && !foundparen) { foundparen = true; pindex = k; }
        }
        if (foundparen && s[s.length() - 1] == ')') {
          interp_add_dep(branches, index, s.substr(pindex + 1, s.length() - pindex - 2), s.substr(0, pindex));
        }
      }
    }
    if (paren > 0) throw std::runtime_error("Error: MrHyDE found an unclosed parenthesis in: " + s);
    if (paren < 0) throw std::runtime_error("Error: MrHyDE found an extra parenthesis in: " + s);
  }
}

template <class EvalT>
bool interp_isOperator(std::vector<Branch<EvalT>>& branches, size_t index, const std::vector<std::string>& ops) {  // :360-446
  std::string s = branches[index].expression;
  std::string oper, argument;
  bool found = false;
  size_t k = 0;
  while (!found && k < ops.size()) {
    size_t L = ops[k].length();
    if (s.length() >= L) {
      bool iseq = true;
      for (size_t j = 0; j < L; j++) if (s[j] != ops[k][j]) iseq = false;
      if (iseq) {
        if (s.length() <= L || s[L] != '(' || s[s.length() - 1] != ')') iseq = false;
        if (iseq) {
          for (size_t j = L + 1; j + 1 < s.length(); j++) {
            if (s[j] == ')') iseq = false;
          }
        }
      }
      if (iseq) {
        found = true;
        oper = ops[k];
        argument = s.substr(L + 1, s.length() - L - 2);
      }
    }
    k += 1;
  }
  if (found) {
    bool has_comma = false;
    size_t comma_ind = 0;
    for (size_t i = 0; i + 1 < argument.length(); ++i) {
      if (argument[i] == ',') { has_comma = true; comma_ind = i; }
    }
    if (has_comma) {
      interp_add_dep(branches, index, argument.substr(0, comma_ind), "");
      interp_add_dep(branches, index, argument.substr(comma_ind + 1), oper);
      branches[index].is_decomposed = true;
    } else {
      interp_add_dep(branches, index, argument, oper);
      branches[index].is_decomposed = true;
    }
  }
  return found;
}

// ---------------------------------------------------------------------------------------
// Pointwise op semantics shared by VToV / SToV / SToS (functionManager_evaluate.hpp:237-1353)
// ---------------------------------------------------------------------------------------
template <class D, class S>
inline void fm_apply_op(D& data, const S& t, const std::string& op) {
  using std::sin; using std::cos; using std::tan; using std::exp; using std::log;
  using std::sqrt; using std::sinh; using std::cosh; using std::pow;
  if (op == "") data = D(t);
  else if (op == "plus") data += t;
  else if (op == "minus") data += -t;
  else if (op == "times") data *= t;
  else if (op == "divide") data /= t;
  else if (op == "power") data = pow(data, D(t));
  else if (op == "sin") data = D(sin(t));
  else if (op == "cos") data = D(cos(t));
  else if (op == "tan") data = D(tan(t));
  else if (op == "exp") data = D(exp(t));
  else if (op == "log") data = D(log(t));
  else if (op == "sinh") data = D(sinh(t));
  else if (op == "cosh") data = D(cosh(t));
  else if (op == "abs") { if (t < 0.0) data = D(-t); else data = D(t); }
  else if (op == "max") { if (t > data) data = D(t); }
  else if (op == "min") { if (t < data) data = D(t); }
  else if (op == "mean") data = 0.5 * data + D(0.5 * t);
  else if (op == "lt") { if (data < t) data = D(1.0); else data = D(0.0); }
  else if (op == "lte") { if (data <= t) data = D(1.0); else data = D(0.0); }
  else if (op == "gt") { if (data > t) data = D(1.0); else data = D(0.0); }
  else if (op == "gte") { if (data >= t) data = D(1.0); else data = D(0.0); }
  else if (op == "sqrt") { if (t <= 0.0) data = D(0.0); else data = D(sqrt(t)); }
  else if (op == "emax" || op == "emin" || op == "emean") data = D(t);  // scalar forms (:1311-1319)
  else throw std::runtime_error("Error: unknown operator in function manager: " + op);
}

// ---------------------------------------------------------------------------------------
// FunctionManager
// ---------------------------------------------------------------------------------------
template <class EvalT>
class FunctionManager {
 public:
  // Hooks into the workset (decomposeFunctions looks leaves up in the workset's field lists,
  // functionManager_create.hpp:107-225, and evaluate() pulls their data, _evaluate.hpp:63-110)
  struct WorksetHooks {
    // returns index or -1; location is "ip" or "side ip"
    std::function<int(const std::string& expr, const std::string& loc)> find_soln_field;
    std::function<int(const std::string& expr, const std::string& loc)> find_scalar_field;
    std::function<View2<EvalT>*(int idx, const std::string& loc)> get_soln_field;     // evaluates if stale
    std::function<View2<double>*(int idx, const std::string& loc)> get_scalar_field;
    std::function<double()> get_time;
    std::function<bool()> is_on_side;
  } hooks;

  std::vector<Forest<EvalT>> forests;
  std::vector<std::string> known_vars = {"x", "y", "z", "u", "v", "w", "t", "pi", "h"};
  std::vector<std::string> known_ops = {"sin", "cos", "exp", "log", "tan", "abs", "max", "min", "mean",
                                        "emax", "emin", "emean", "sqrt", "sinh", "cosh"};
  bool decomposed = false;

  FunctionManager() {}
  FunctionManager(int num_elem, int num_ip, int num_ip_side) {  // functionManager_construct.hpp:25-40
    auto mk = [&](const char* loc, int d0, int d1) { Forest<EvalT> f; f.location = loc; f.dim0 = d0; f.dim1 = d1; forests.push_back(f); };
    mk("ip", num_elem, num_ip);
    mk("side ip", num_elem, num_ip_side);
    mk("point", 1, 1);
  }

  // functionManager.cpp addFunction: spaces are not stripped here; a repeated name at the same
  // location keeps the first registration (user "Functions:" entries are registered before module defaults)
  int addFunction(const std::string& fname, const std::string& expression, const std::string& location) {
    for (auto& f : forests) {
      if (f.location == location) {
        for (auto& t : f.trees) if (t.name == fname) return 0;
        Tree<EvalT> t;
        t.name = fname; t.expression = expression;
        t.branches.push_back(Branch<EvalT>(expression));
        f.trees.push_back(t);
        decomposed = false;
        return 1;
      }
    }
    throw std::runtime_error("Error: function manager has no location " + location);
  }

  void decomposeFunctions() {  // functionManager_create.hpp:78-540
    for (size_t fiter = 0; fiter < forests.size(); fiter++) {
      Forest<EvalT>& F = forests[fiter];
      const int maxiter = 100;
      for (size_t titer = 0; titer < F.trees.size(); titer++) {
        bool done = false;
        int iter = 0;
        while (!done && iter < maxiter) {
          iter++;
          size_t Nbranches = F.trees[titer].branches.size();
          for (size_t k = 0; k < Nbranches; k++) {
            auto& B = [&]() -> Branch<EvalT>& { return F.trees[titer].branches[k]; }();
            bool decompose = !(B.is_leaf || B.is_decomposed);
            std::string expr = B.expression;
            if (decompose && (F.location == "ip" || F.location == "side ip")) {  // AD data stored in the workset
              int j = hooks.find_soln_field ? hooks.find_soln_field(expr, F.location) : -1;
              if (j >= 0) {
                decompose = false;
                B.is_leaf = B.is_decomposed = B.is_view = B.is_AD = B.is_workset_data = true;
                B.workset_data_index = j;
              }
            }
            if (decompose && (F.location == "ip" || F.location == "side ip")) {  // Scalar data stored in the workset
              int j = hooks.find_scalar_field ? hooks.find_scalar_field(expr, F.location) : -1;
              if (j >= 0) {
                decompose = false;
                B.is_leaf = B.is_decomposed = B.is_view = B.is_workset_data = true;
                B.is_AD = false;
                B.workset_data_index = j;
              }
            }
            // (scalar / discretized parameters: out of scope, path runs with no active parameters)
            if (decompose) {  // another function at this location
              for (size_t j = 0; j < F.trees.size(); j++) {
                if (expr == F.trees[j].name) {
                  F.trees[titer].branches[k].is_decomposed = true;
                  F.trees[titer].branches[k].is_func = true;
                  F.trees[titer].branches[k].func_index = (int)j;
                  decompose = false;
                }
              }
            }
            if (decompose) {  // simple scalar
              if (interp_isScalar(expr)) {
                auto& b = F.trees[titer].branches[k];
                b.is_leaf = b.is_decomposed = b.is_constant = true;
                b.data_Sc = std::stod(expr);
                decompose = false;
              }
            }
            if (decompose) {  // known variables t / pi
              for (size_t j = 0; j < known_vars.size(); j++) {
                if (expr == known_vars[j]) {
                  auto& b = F.trees[titer].branches[k];
                  decompose = false;
                  b.is_leaf = b.is_decomposed = true;
                  if (known_vars[j] == "t") { b.is_time = true; b.data_Sc = hooks.get_time ? hooks.get_time() : 0.0; }
                  else if (known_vars[j] == "pi") { b.is_constant = true; b.data_Sc = ORACLE_PI; }
                }
              }
            }
            if (decompose) {
              if (interp_isOperator(F.trees[titer].branches, k, known_ops)) decompose = false;
            }
            if (decompose) {
              size_t cnumb = F.trees[titer].branches.size();
              interp_split(F.trees[titer].branches, k);
              F.trees[titer].branches[k].is_decomposed = true;
              if (cnumb == F.trees[titer].branches.size()) {
                const std::string& e2 = F.trees[titer].branches[k].expression;
                if (!(e2 == "n[x]" || e2 == "n[y]" || e2 == "n[z]" || e2 == "t[x]" || e2 == "t[y]" || e2 == "t[z]"))
                  throw std::runtime_error("Error: MrHyDE was not able to decompose or find: " + e2);
              }
            }
          }
          bool isdone = true;
          for (auto& b : F.trees[titer].branches) if (!b.is_leaf && !b.is_decomposed) isdone = false;
          done = isdone;
        }
        if (!done && iter >= maxiter) throw std::runtime_error("Error: MrHyDE was not able to decompose " + F.trees[titer].name);
      }
    }
    for (auto& F : forests) {
      for (size_t k = 0; k < F.trees.size(); k++) {
        for (size_t j = 0; j < F.trees[k].branches.size(); j++) {
          bool isConst = true, isView = false, isAD = false;
          checkDepDataType(F, (int)k, (int)j, isConst, isView, isAD);
          auto& b = F.trees[k].branches[j];
          b.is_constant = isConst; b.is_view = isView; b.is_AD = isAD;
          if (isView && !b.is_workset_data) {
            if (isAD) { b.own = View2<EvalT>(F.dim0, F.dim1); b.viewdata = &b.own; }
            else { b.own_Sc = View2<double>(F.dim0, F.dim1); b.viewdata_Sc = &b.own_Sc; }
          }
        }
      }
    }
    decomposed = true;
    for (size_t f = 0; f < forests.size(); ++f) {
      for (size_t k = 0; k < forests[f].trees.size(); k++) {
        for (size_t j = 0; j < forests[f].trees[k].branches.size(); j++) {
          auto& b = forests[f].trees[k].branches[j];
          if (b.is_constant && !b.is_leaf) evaluate(f, k, j);
        }
      }
    }
  }

  void checkDepDataType(Forest<EvalT>& F, int tindex, int bindex, bool& isConst, bool& isView, bool& isAD) {  // :583-619
    auto& b = F.trees[tindex].branches[bindex];
    if (b.currently_checking) throw std::runtime_error("Error: MrHyDE detected a cyclic graph in: " + b.expression);
    b.currently_checking = true;
    if (b.is_leaf) {
      if (!b.is_constant) isConst = false;
      if (b.is_view) isView = true;
      if (b.is_AD) isAD = true;
    } else if (b.is_func) {
      checkDepDataType(F, b.func_index, 0, isConst, isView, isAD);
    } else {
      for (size_t k = 0; k < b.dep_list.size(); k++) checkDepDataType(F, tindex, b.dep_list[k], isConst, isView, isAD);
    }
    F.trees[tindex].branches[bindex].currently_checking = false;
  }

  Vista<EvalT> evaluate(const std::string& fname, const std::string& location) {  // _evaluate.hpp:14-52
    for (size_t fiter = 0; fiter < forests.size(); fiter++) {
      if (forests[fiter].location != location) continue;
      for (size_t titer = 0; titer < forests[fiter].trees.size(); titer++) {
        if (fname != forests[fiter].trees[titer].name) continue;
        if (!decomposed) decomposeFunctions();
        auto& b0 = forests[fiter].trees[titer].branches[0];
        if (!b0.is_constant) evaluate(fiter, titer, 0);
        Vista<EvalT> v;
        v.is_view = b0.is_view; v.is_AD = b0.is_AD;
        if (b0.is_view) { v.vdata = b0.viewdata; v.vdata_sc = b0.viewdata_Sc; }
        else v.sdata = b0.is_AD ? b0.data : EvalT(b0.data_Sc);
        return v;
      }
    }
    throw std::runtime_error("Error: function manager could not evaluate: " + fname + " at " + location);
  }

  bool isConstant(const std::string& fname, const std::string& location) {
    for (auto& F : forests) if (F.location == location)
      for (auto& t : F.trees) if (t.name == fname) { if (!decomposed) decomposeFunctions(); return t.branches[0].is_constant; }
    return false;
  }

  // Parse-tree dump in the shape the reference prints with verbosity 100 (regression/functions/Valid)
  std::string printTree(const std::string& fname, const std::string& location) {
    std::string out;
    for (auto& F : forests) if (F.location == location)
      for (auto& t : F.trees) if (t.name == fname) {
        if (!decomposed) decomposeFunctions();
        for (size_t j = 0; j < t.branches.size(); ++j) {
          out += std::to_string(j) + ":" + t.branches[j].expression + "|";
          for (size_t k = 0; k < t.branches[j].dep_list.size(); ++k)
            out += t.branches[j].dep_ops[k] + ">" + std::to_string(t.branches[j].dep_list[k]) + ",";
          out += "\n";
        }
      }
    return out;
  }

  void evaluate(size_t findex, size_t tindex, size_t bindex) {  // _evaluate.hpp:59-229
    Forest<EvalT>& F = forests[findex];
    auto& B = F.trees[tindex].branches[bindex];
    if (B.is_leaf) {
      if (B.is_workset_data) {
        const std::string loc = (hooks.is_on_side && hooks.is_on_side()) ? "side ip" : "ip";
        if (B.is_AD) B.viewdata = hooks.get_soln_field(B.workset_data_index, loc);
        else B.viewdata_Sc = hooks.get_scalar_field(B.workset_data_index, loc);
      } else if (B.is_time) {
        B.data_Sc = hooks.get_time();
      }
    } else if (B.is_func) {
      int fi = B.func_index;
      evaluate(findex, fi, 0);
      auto& S = F.trees[fi].branches[0];
      auto& B2 = F.trees[tindex].branches[bindex];
      if (B2.is_AD) { if (B2.is_view) B2.viewdata = S.viewdata; else B2.data = S.data; }
      else { if (B2.is_view) B2.viewdata_Sc = S.viewdata_Sc; else B2.data_Sc = S.data_Sc; }
    } else {
      const bool isAD = B.is_AD, isView = B.is_view;
      const size_t ndeps = B.dep_list.size();
      for (size_t k = 0; k < ndeps; k++) {
        int dep = F.trees[tindex].branches[bindex].dep_list[k];
        evaluate(findex, tindex, dep);
        auto& P = F.trees[tindex].branches[bindex];
        auto& T = F.trees[tindex].branches[dep];
        const std::string& op = P.dep_ops[k];
        if (isView) {
          if (T.is_view) {
            if (isAD) { if (T.is_AD) opVToV(*P.viewdata, *T.viewdata, op); else opVToV(*P.viewdata, *T.viewdata_Sc, op); }
            else if (!T.is_AD) opVToV(*P.viewdata_Sc, *T.viewdata_Sc, op);
          } else {
            if (isAD) { if (T.is_AD) opSToV(*P.viewdata, T.data, op); else opSToV(*P.viewdata, T.data_Sc, op); }
            else if (!T.is_AD) opSToV(*P.viewdata_Sc, T.data_Sc, op);
          }
        } else if (!T.is_view) {
          if (isAD) { if (T.is_AD) fm_apply_op(P.data, T.data, op); else fm_apply_op(P.data, T.data_Sc, op); }
          else if (!T.is_AD) fm_apply_op(P.data_Sc, T.data_Sc, op);
        }
      }
    }
  }

 private:
  template <class D, class S>
  void opVToV(View2<D>& data, const View2<S>& t, const std::string& op) {  // :237-560
    const int dim0 = std::min(data.n0, t.n0), dim1 = std::min(data.n1, t.n1);
    if (op == "emax" || op == "emin") {
      const bool mx = (op == "emax");
      for (int e = 0; e < dim0; ++e) {
        data(e, 0) = D(t(e, 0));
        for (int n = 0; n < dim1; ++n) if (mx ? (t(e, n) > t(e, 0)) : (t(e, n) < t(e, 0))) data(e, 0) = D(t(e, n));
        for (int n = 0; n < dim1; ++n) data(e, n) = data(e, 0);
      }
      return;
    }
    if (op == "emean") {
      for (int e = 0; e < dim0; ++e) {
        const double scale = (double)dim1;
        data(e, 0) = D(t(e, 0) / scale);
        for (int n = 0; n < dim1; ++n) data(e, 0) += D(t(e, n) / scale);
        for (int n = 0; n < dim1; ++n) data(e, n) = data(e, 0);
      }
      return;
    }
    for (int e = 0; e < dim0; ++e)
      for (int pt = 0; pt < dim1; ++pt) fm_apply_op(data(e, pt), t(e, pt), op);
  }
  template <class D, class S>
  void opSToV(View2<D>& data, const S& t, const std::string& op) {  // :891-1240
    if (op == "power") {
      if constexpr (std::is_arithmetic<S>::value) {  // small integer exponent: repeated multiply (:952-983)
        double pscalar = (double)t, rounded = std::round(pscalar);
        if (rounded >= 2.0 && rounded <= 4.0 && std::abs(pscalar - rounded) < 1.0e-12) {
          const int n = (int)rounded;
          for (int e = 0; e < data.n0; ++e)
            for (int pt = 0; pt < data.n1; ++pt) {
              D base = data(e, pt), acc = base;
              for (int j = 1; j < n; ++j) acc = acc * base;
              data(e, pt) = acc;
            }
          return;
        }
      }
    }
    for (int e = 0; e < data.n0; ++e)
      for (int pt = 0; pt < data.n1; ++pt) fm_apply_op(data(e, pt), t, op);
  }
};

}  // namespace oracle
