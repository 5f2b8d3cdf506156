// Device evaluator of the flattened expression programs (expr.hpp).  One thread evaluates one
// point; the program lives in kernel-parameter constant memory so the opcode fetch is a uniform
// constant load and the switch is a warp-uniform branch.
#pragma once
#include "expr.hpp"

namespace mrhyde_b200 {

struct ExprVars {  // workset scalar fields at one point: x y z t n[x] n[y] n[z]
  double v[7];
};

__device__ __forceinline__ double expr_var(const ExprVars& in, int i) {
  double r = in.v[0];
  r = (i == 1) ? in.v[1] : r;
  r = (i == 2) ? in.v[2] : r;
  r = (i == 3) ? in.v[3] : r;
  r = (i == 4) ? in.v[4] : r;
  r = (i == 5) ? in.v[5] : r;
  r = (i == 6) ? in.v[6] : r;
  return r;
}

static __device__ __noinline__ double expr_eval_program(const ExprProgram& p, const ExprVars& in) {
  double st[EXPR_MAXSTACK];
  int sp = 0;        // number of values below the top-of-stack register
  double a = 0.0;    // top of stack
  const int n = p.n;
  for (int i = 0; i < n; ++i) {
    const double c = p.c[i];
    switch (p.op[i]) {
      case OP_PUSHC: st[sp & (EXPR_MAXSTACK - 1)] = a; ++sp; a = c; break;
      case OP_PUSHV: st[sp & (EXPR_MAXSTACK - 1)] = a; ++sp; a = expr_var(in, (int)c); break;
      case OP_ADD: --sp; a = st[sp & (EXPR_MAXSTACK - 1)] + a; break;
      case OP_SUB: --sp; a = st[sp & (EXPR_MAXSTACK - 1)] + (-a); break;
      case OP_MUL: --sp; a = st[sp & (EXPR_MAXSTACK - 1)] * a; break;
      case OP_DIV: --sp; a = st[sp & (EXPR_MAXSTACK - 1)] / a; break;
      case OP_POW: --sp; a = pow(st[sp & (EXPR_MAXSTACK - 1)], a); break;
      case OP_LT: --sp; a = st[sp & (EXPR_MAXSTACK - 1)] < a ? 1.0 : 0.0; break;
      case OP_LTE: --sp; a = st[sp & (EXPR_MAXSTACK - 1)] <= a ? 1.0 : 0.0; break;
      case OP_GT: --sp; a = st[sp & (EXPR_MAXSTACK - 1)] > a ? 1.0 : 0.0; break;
      case OP_GTE: --sp; a = st[sp & (EXPR_MAXSTACK - 1)] >= a ? 1.0 : 0.0; break;
      case OP_MAX: { --sp; const double l = st[sp & (EXPR_MAXSTACK - 1)]; a = a > l ? a : l; } break;
      case OP_MIN: { --sp; const double l = st[sp & (EXPR_MAXSTACK - 1)]; a = a < l ? a : l; } break;
      case OP_MEAN: --sp; a = 0.5 * st[sp & (EXPR_MAXSTACK - 1)] + 0.5 * a; break;
      case OP_ADDC: a = a + c; break;
      case OP_SUBC: a = a + (-c); break;
      case OP_MULC: a = a * c; break;
      case OP_DIVC: a = a / c; break;
      case OP_POWC: a = pow(a, c); break;
      case OP_ADDV: a = a + expr_var(in, (int)c); break;
      case OP_SUBV: a = a + (-expr_var(in, (int)c)); break;
      case OP_MULV: a = a * expr_var(in, (int)c); break;
      case OP_DIVV: a = a / expr_var(in, (int)c); break;
      case OP_SIN: a = sin(a); break;
      case OP_COS: a = cos(a); break;
      case OP_TAN: a = tan(a); break;
      case OP_EXP: a = exp(a); break;
      case OP_LOG: a = log(a); break;
      case OP_ABS: a = a < 0.0 ? -a : a; break;
      case OP_SQRT: a = a <= 0.0 ? 0.0 : sqrt(a); break;
      case OP_SINH: a = sinh(a); break;
      case OP_COSH: a = cosh(a); break;
      default: break;
    }
  }
  return a;
}

__device__ __forceinline__ double expr_eval(const ExprProgram& p, const ExprVars& in) {
  if (p.is_const) return p.cval;
  return expr_eval_program(p, in);
}

}  // namespace mrhyde_b200
