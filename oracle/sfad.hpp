// ORACLE (test infrastructure, never shipped, never on the product path).
//
// Minimal forward-mode AD value type with a compile-time number of derivative
// components.  Restates the subset of Sacado::Fad::SFad<double,N> semantics the
// reference hot path relies on (Sacado itself is an un-vendored Trilinos package,
// version unpinned by the reference):
//   * typedefs AD2..AD32 / AD (SFad<ScalarT,MAXDERIVS>)      src/preferences.hpp:74-97
//   * seeding ctor  EvalT(size, index, value)               src/tools/workset.cpp:644,884
//   * .val() / .fastAccessDx(i) in the scatter              src/managers/assembly/assemblyManager_scatter.hpp:213-261
//   * comparisons act on the value only (used by the function manager's lt/gt/abs/max ops,
//     src/managers/function/functionManager_evaluate.hpp:369-520)
#pragma once
#include <cmath>

namespace oracle {

template <int N>
struct SFad {
  double v;
  double d[N];

  SFad() : v(0.0) { for (int i = 0; i < N; ++i) d[i] = 0.0; }
  SFad(double x) : v(x) { for (int i = 0; i < N; ++i) d[i] = 0.0; }
  // SFad(sz, i, x): value x, unit derivative in slot i
  SFad(int /*sz*/, int i, double x) : v(x) { for (int k = 0; k < N; ++k) d[k] = 0.0; d[i] = 1.0; }

  static constexpr int size() { return N; }
  double val() const { return v; }
  double fastAccessDx(int i) const { return d[i]; }
  double& fastAccessDx(int i) { return d[i]; }

  SFad& operator+=(const SFad& b) { v += b.v; for (int i = 0; i < N; ++i) d[i] += b.d[i]; return *this; }
  SFad& operator-=(const SFad& b) { v -= b.v; for (int i = 0; i < N; ++i) d[i] -= b.d[i]; return *this; }
  SFad& operator*=(const SFad& b) { for (int i = 0; i < N; ++i) d[i] = d[i] * b.v + v * b.d[i]; v *= b.v; return *this; }
  SFad& operator/=(const SFad& b) {
    for (int i = 0; i < N; ++i) d[i] = (d[i] * b.v - v * b.d[i]) / (b.v * b.v);
    v /= b.v; return *this;
  }
  SFad& operator+=(double b) { v += b; return *this; }
  SFad& operator-=(double b) { v -= b; return *this; }
  SFad& operator*=(double b) { v *= b; for (int i = 0; i < N; ++i) d[i] *= b; return *this; }
  SFad& operator/=(double b) { v /= b; for (int i = 0; i < N; ++i) d[i] /= b; return *this; }
};

template <int N> inline SFad<N> operator-(const SFad<N>& a) { SFad<N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <int N> inline SFad<N> operator+(const SFad<N>& a, const SFad<N>& b) { SFad<N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> inline SFad<N> operator-(const SFad<N>& a, const SFad<N>& b) { SFad<N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> inline SFad<N> operator*(const SFad<N>& a, const SFad<N>& b) { SFad<N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N> inline SFad<N> operator/(const SFad<N>& a, const SFad<N>& b) {
  SFad<N> r; r.v = a.v / b.v; for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] * b.v - a.v * b.d[i]) / (b.v * b.v); return r;
}
template <int N> inline SFad<N> operator+(const SFad<N>& a, double b) { SFad<N> r = a; r.v += b; return r; }
template <int N> inline SFad<N> operator+(double a, const SFad<N>& b) { SFad<N> r = b; r.v += a; return r; }
template <int N> inline SFad<N> operator-(const SFad<N>& a, double b) { SFad<N> r = a; r.v -= b; return r; }
template <int N> inline SFad<N> operator-(double a, const SFad<N>& b) { SFad<N> r; r.v = a - b.v; for (int i = 0; i < N; ++i) r.d[i] = -b.d[i]; return r; }
template <int N> inline SFad<N> operator*(const SFad<N>& a, double b) { SFad<N> r; r.v = a.v * b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b; return r; }
template <int N> inline SFad<N> operator*(double a, const SFad<N>& b) { SFad<N> r; r.v = a * b.v; for (int i = 0; i < N; ++i) r.d[i] = a * b.d[i]; return r; }
template <int N> inline SFad<N> operator/(const SFad<N>& a, double b) { SFad<N> r; r.v = a.v / b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / b; return r; }
template <int N> inline SFad<N> operator/(double a, const SFad<N>& b) {
  SFad<N> r; r.v = a / b.v; for (int i = 0; i < N; ++i) r.d[i] = -a * b.d[i] / (b.v * b.v); return r;
}

#define ORACLE_CMP(OP) \
  template <int N> inline bool operator OP(const SFad<N>& a, const SFad<N>& b) { return a.v OP b.v; } \
  template <int N> inline bool operator OP(const SFad<N>& a, double b) { return a.v OP b; } \
  template <int N> inline bool operator OP(double a, const SFad<N>& b) { return a OP b.v; }
ORACLE_CMP(<) ORACLE_CMP(>) ORACLE_CMP(<=) ORACLE_CMP(>=)
#undef ORACLE_CMP

#define ORACLE_UNARY(NAME, VAL, DER) \
  template <int N> inline SFad<N> NAME(const SFad<N>& a) { \
    SFad<N> r; const double x = a.v; r.v = VAL; const double g = DER; \
    for (int i = 0; i < N; ++i) { r.d[i] = g * a.d[i]; } \
    return r; }
ORACLE_UNARY(sin, std::sin(x), std::cos(x))
ORACLE_UNARY(cos, std::cos(x), -std::sin(x))
ORACLE_UNARY(tan, std::tan(x), 1.0 + std::tan(x) * std::tan(x))
ORACLE_UNARY(exp, std::exp(x), std::exp(x))
ORACLE_UNARY(log, std::log(x), 1.0 / x)
ORACLE_UNARY(sqrt, std::sqrt(x), 0.5 / std::sqrt(x))
ORACLE_UNARY(sinh, std::sinh(x), std::cosh(x))
ORACLE_UNARY(cosh, std::cosh(x), std::sinh(x))
#undef ORACLE_UNARY

template <int N> inline SFad<N> pow(const SFad<N>& a, const SFad<N>& b) {
  SFad<N> r; r.v = std::pow(a.v, b.v);
  // d(a^b) = a^b (b' ln a + b a'/a); Sacado guards a == 0 by returning zero derivatives
  if (a.v == 0.0) return r;
  const double la = std::log(a.v);
  for (int i = 0; i < N; ++i) {
    double t = b.v * a.d[i] / a.v;
    if (b.d[i] != 0.0) t += b.d[i] * la;
    r.d[i] = r.v * t;
  }
  return r;
}
template <int N> inline SFad<N> pow(const SFad<N>& a, double b) {
  SFad<N> r; r.v = std::pow(a.v, b);
  if (a.v == 0.0) return r;
  const double g = b * r.v / a.v;
  for (int i = 0; i < N; ++i) r.d[i] = g * a.d[i];
  return r;
}

// ---- uniform access for EvalT = double | SFad<N> (the reference instantiates every
// class for ScalarT and the AD types alike, physicsBase.hpp / workset.cpp explicit instantiations)
template <class T> struct ADTraits;
template <> struct ADTraits<double> {
  static constexpr int size = 0;
  static double val(const double& x) { return x; }
  static double dx(const double&, int) { return 0.0; }
  static double seed(int, double x) { return x; }
};
template <int N> struct ADTraits<SFad<N>> {
  static constexpr int size = N;
  static double val(const SFad<N>& x) { return x.v; }
  static double dx(const SFad<N>& x, int i) { return x.d[i]; }
  static SFad<N> seed(int i, double x) { return SFad<N>(N, i, x); }
};

}  // namespace oracle
