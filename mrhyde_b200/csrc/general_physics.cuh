// Weak-form coefficients of the restated physics modules at one quadrature point, written once for T = double
// (residual values) and T = Dual<K> (derivative components) -- the analogue of the reference's EvalT template.
//
// Every module returns, per variable v, the coefficient vector Cf[v][k] that multiplies the test function's
// physical basis entries PB[b(v)][i][q][k]:
//     HGRAD  k = 0: phi_i          k = 1..3: d phi_i / dx,dy,dz
//     HCURL  k = 0..2: phi_i (vector)   k = 3..5: curl phi_i
//     HDIV   k = 0..2: psi_i (vector)   k = 3: div psi_i
// so that  res(e, off(v,i)) += sum_q sum_k Cf[v][k] PB[..][k]   reproduces the module's volumeResidual /
// boundaryResidual loops (weights are folded into Cf as in the reference).  Fields use the same layout:
// F[v][k] = sum_i u_i PB[..][k]  (value, gradient | value, curl | value, div), Ft[v][k] the time derivative of the values.
//
//   thermal           src/physics/thermal.cpp:70-165 (volume), :171-281 (boundary)
//   linearelasticity  src/physics/linearelasticity.cpp:91-238, :243-674, computeStress :915-1273
//   navierstokes      src/physics/navierstokes.cpp:81-848, :855-1013, computeTau :1053-1081
//   maxwell           src/physics/maxwell.cpp:79-308, :313-403
#pragma once
#include "general_kernel.cuh"

namespace mrhyde_b200 {

MRH_CE int ipow(int b, int e) { int r = 1; for (int i = 0; i < e; ++i) r *= b; return r; }

// shared by the single-basis HGRAD modules
template <int DIM_, int P_, int NVAR_>
struct HGradSystem {
  static constexpr int DIM = DIM_, ORDER = P_, NVAR = NVAR_, NBASIS = 1, NC = 4;
  static constexpr int CARD = ipow(P_ + 1, DIM_);
  static constexpr int N = NVAR_ * CARD;
  static MRH_CE int card(int) { return CARD; }
  static MRH_CE int ncb(int) { return 4; }
  static MRH_CE int nval(int) { return 1; }
  static MRH_CE int btype(int) { return BT_HGRAD; }
  static MRH_CE int var_basis(int) { return 0; }
  static MRH_CE int var_basis_rt(int) { return 0; }
  static MRH_CE int ncb_rt(int) { return 4; }
  static MRH_CE int first_var(int) { return 0; }
  static MRH_CE int card_of_var(int) { return CARD; }
  static MRH_CE int row0(int v) { return v * CARD; }
  static MRH_CE int max_card() { return CARD; }
};

// ---------------------------------------------------------------------------------------------------------
// thermal: functions 0 source, 1 diffusion, 2 specific heat, 3 density, 4..6 advection x,y,z, 7 robin alpha (unused by
// the reference's residual, thermal.cpp:181), bdata at NFN + v
// ---------------------------------------------------------------------------------------------------------
template <int DIM_, int P_>
struct ThermalPhys : HGradSystem<DIM_, P_, 1> {
  static constexpr int NFN = 8;
  template <class T, bool STATE = false, int FN0 = 0, int BD0 = NFN, int V0 = 0>
  MRH_HD static void volume(const QpCtx& c, const GenOpts& o, const T (&F)[1][4], const T (&Ft)[1][4], T (&Cf)[1][4]) {
    const auto source = MRH_FN(FN0 + 0), diff = MRH_FN(FN0 + 1), cp = MRH_FN(FN0 + 2), rho = MRH_FN(FN0 + 3);
    Cf[0][0] = (rho * cp * Ft[0][0] - source) * c.w;
    for (int d = 0; d < DIM_; ++d) Cf[0][1 + d] = diff * F[0][1 + d] * c.w;
    if (o.have_advection) {
      T adv = MRH_FN(FN0 + 4) * F[0][1];
      if (DIM_ > 1) adv = adv + MRH_FN(FN0 + 5) * F[0][2];
      if (DIM_ > 2) adv = adv + MRH_FN(FN0 + 6) * F[0][3];
      Cf[0][0] = Cf[0][0] + adv * c.w;
    }
  }
  template <class T, bool STATE = false, int FN0 = 0, int BD0 = NFN, int V0 = 0>
  MRH_HD static void boundary(const QpCtx& c, const GenOpts& o, const T (&F)[1][4], const T (&)[1][4], T (&Cf)[1][4]) {
    const auto diff = MRH_FN(FN0 + 1), bdata = MRH_FN(BD0 + 0);
    if (c.bc_type[V0] == BC_NEUMANN) {
      Cf[0][0] = T(0.0) - bdata * c.w;
    } else if (c.bc_type[V0] == BC_WEAK_DIRICHLET) {
      const double epen = 10.0;
      T flux = F[0][1] * c.n[0] + F[0][2] * c.n[1];
      if (DIM_ > 2) flux = flux + F[0][3] * c.n[2];
      const T jump = F[0][0] - bdata;
      Cf[0][0] = (epen * c.ih * diff * jump - diff * flux) * c.w;
      for (int d = 0; d < DIM_; ++d) Cf[0][1 + d] = -o.form_param * diff * jump * c.w * c.n[d];
    }
  }
};

// ---------------------------------------------------------------------------------------------------------
// linear elasticity: functions 0 lambda, 1 mu, 2..4 source dx,dy,dz; variables dx, dy(, dz)
// ---------------------------------------------------------------------------------------------------------
template <int DIM_, int P_>
struct ElasticityPhys : HGradSystem<DIM_, P_, DIM_> {
  static constexpr int NFN = 5;
  static constexpr int NVAR = DIM_;
  template <class T, bool STATE = false, int FN0 = 0, int BD0 = NFN, int V0 = 0>
  MRH_HD static void stress(const QpCtx& c, const GenOpts& o, const T (&F)[NVAR][4], T (&s)[3][3]) {
    const auto lambda = MRH_FN(FN0 + 0), mu = MRH_FN(FN0 + 1);
    // computeStress (linearelasticity.cpp:1158-1240)
    if (DIM_ == 2) {
      if (o.incplanestress) {
        s[0][0] = 4.0 * mu * F[0][1] + 2.0 * mu * F[1][2];
        s[1][1] = 4.0 * mu * F[1][2] + 2.0 * mu * F[0][1];
      } else {
        s[0][0] = (2.0 * mu + lambda) * F[0][1] + lambda * F[1][2];
        s[1][1] = (2.0 * mu + lambda) * F[1][2] + lambda * F[0][1];
      }
      s[0][1] = mu * (F[0][2] + F[1][1]);
      s[1][0] = s[0][1];
    } else {
      const int X = 0, Y = 1, Z = DIM_ > 2 ? 2 : 0;
      s[0][0] = (2.0 * mu + lambda) * F[X][1] + lambda * (F[Y][2] + F[Z][3]);
      s[0][1] = mu * (F[X][2] + F[Y][1]);
      s[0][2] = mu * (F[X][3] + F[Z][1]);
      s[1][0] = s[0][1];
      s[1][1] = (2.0 * mu + lambda) * F[Y][2] + lambda * (F[X][1] + F[Z][3]);
      s[1][2] = mu * (F[Y][3] + F[Z][2]);
      s[2][0] = s[0][2];
      s[2][1] = s[1][2];
      s[2][2] = (2.0 * mu + lambda) * F[Z][3] + lambda * (F[X][1] + F[Y][2]);
    }
  }
  template <class T, bool STATE = false, int FN0 = 0, int BD0 = NFN, int V0 = 0>
  MRH_HD static void volume(const QpCtx& c, const GenOpts& o, const T (&F)[NVAR][4], const T (&)[NVAR][4], T (&Cf)[NVAR][4]) {
    T s[3][3];
    stress<T, STATE, FN0, BD0, V0>(c, o, F, s);
    for (int d = 0; d < DIM_; ++d) {
      Cf[d][0] = T(0.0) - MRH_FN(FN0 + 2 + d) * c.w;
      for (int e = 0; e < DIM_; ++e) Cf[d][1 + e] = s[d][e] * c.w;
    }
  }
  template <class T, bool STATE = false, int FN0 = 0, int BD0 = NFN, int V0 = 0>
  MRH_HD static void boundary(const QpCtx& c, const GenOpts& o, const T (&F)[NVAR][4], const T (&)[NVAR][4], T (&Cf)[NVAR][4]) {
    const auto lam = MRH_FN(FN0 + 0), mu = MRH_FN(FN0 + 1);
    bool any = false;
    for (int d = 0; d < DIM_; ++d) any = any || c.bc_type[V0 + d] == BC_WEAK_DIRICHLET;
    T s[3][3], delta[3];
    if (any) {
      stress<T, STATE, FN0, BD0, V0>(c, o, F, s);
      // data of a variable that is neither Neumann nor weak Dirichlet on this side is an unset Vista in the reference
      // (linearelasticity.cpp:264-283); it reads as 0 here
      for (int d = 0; d < DIM_; ++d) delta[d] = F[d][0] - MRH_FN(BD0 + d);
    }
    for (int d = 0; d < DIM_; ++d) {
      if (c.bc_type[V0 + d] == BC_NEUMANN) {
        Cf[d][0] = T(0.0) - MRH_FN(BD0 + d) * c.w;
      } else if (c.bc_type[V0 + d] == BC_WEAK_DIRICHLET) {
        const auto penalty = o.penalty * (lam + 2.0 * mu) * c.ih;
        T trac = s[d][0] * c.n[0] + s[d][1] * c.n[1];
        if (DIM_ > 2) trac = trac + s[d][2] * c.n[2];
        Cf[d][0] = (penalty * delta[d] - trac) * c.w;
        // bb_e = (C : (delta (x) n))_{d e}   (linearelasticity.cpp:373-375, 421-423, 503-509, 556-558, 609-611)
        for (int e = 0; e < DIM_; ++e) {
          T bb;
          if (e == d) {
            bb = (lam + 2.0 * mu) * delta[d] * c.n[d];
            for (int f = 0; f < DIM_; ++f) if (f != d) bb = bb + lam * delta[f] * c.n[f];
          } else {
            bb = mu * delta[e] * c.n[d] + mu * delta[d] * c.n[e];
          }
          Cf[d][1 + e] = -o.form_param * bb * c.w;
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------------------
// Navier-Stokes: variables ux, pr, uy(, uz) in that order (navierstokes.cpp:27-43); functions 0 source ux, 1 source pr,
// 2 source uy, 3 source uz, 4 density, 5 viscosity
// ---------------------------------------------------------------------------------------------------------
template <int DIM_, int P_>
struct NavierStokesPhys : HGradSystem<DIM_, P_, DIM_ + 1> {
  static constexpr int NFN = 6;
  static constexpr int NVAR = DIM_ + 1;
  MRH_HD static int vel(int d) { return d == 0 ? 0 : d + 1; }   // ux 0, uy 2, uz 3; pr 1
  MRH_HD static int src(int d) { return d == 0 ? 0 : d + 1; }   // source ux 0, uy 2, uz 3
  // computeTau (navierstokes.cpp:1053-1081): tau = [(C1 nu/h^2)^2 + (C2 |u|/h)^2 + (C3/dt)^2]^(-1/2); divisions by h, dt are
  // multiplications by reciprocals and the square roots share one rsqrt (a few ulp, far inside the 1e-12 tolerance)
  template <class T, class V>
  MRH_HD static T compute_tau(const V visc, const T (&u)[3], double ih, double dt, int transient) {
    const double C1 = 4.0, C2 = 2.0, C3 = transient ? 2.0 : 0.0;
    T nvel = u[0] * u[0] + u[1] * u[1];
    if (DIM_ > 2) nvel = nvel + u[2] * u[2];
    if (mrh_val(nvel) > 1E-12) nvel = mrh_sqrt(nvel);   // below the threshold the squared speed is used
    const auto t1 = (C1 * ih * ih) * visc;
    const double t3 = transient ? C3 / dt : 0.0;
    const T nv = (C2 * ih) * nvel;
    return mrh_rsqrt(nv * nv + (t1 * t1 + t3 * t3));
  }
  template <class T, bool STATE = false, int FN0 = 0, int BD0 = NFN, int V0 = 0>
  MRH_HD static void volume(const QpCtx& c, const GenOpts& o, const T (&F)[NVAR][4], const T (&Ft)[NVAR][4], T (&Cf)[NVAR][4]) {
    const auto dens = MRH_FN(FN0 + 4), visc = MRH_FN(FN0 + 5);
    T u[3];
    for (int d = 0; d < 3; ++d) u[d] = d < DIM_ ? F[vel(d < DIM_ ? d : 0)][0] : T(0.0);
    const T pr = F[1][0];
    T tau = T(0.0);
    if (o.useSUPG || o.usePSPG) tau = compute_tau(visc, u, c.ih, c.dt, c.transient);
    T divu = T(0.0);
    const auto wdens = dens * c.w;
    auto pspg_w = c.w / dens;
    if (!o.usePSPG) pspg_w = pspg_w * 0.0;
#pragma unroll
    for (int d = 0; d < DIM_; ++d) {
      const int v = vel(d);
      T conv = u[0] * F[v][1] + u[1] * F[v][2];
      if (DIM_ > 2) conv = conv + u[2] * F[v][3];
      T co[4];
      co[0] = (Ft[v][0] + conv - MRH_FN(FN0 + src(d))) * wdens;
      for (int e = 0; e < DIM_; ++e) {
        T Fe = visc * F[v][1 + e];
        if (e == d) Fe = Fe - pr;
        co[1 + e] = Fe * c.w;
      }
      T stabres = T(0.0);
      if (o.useSUPG || o.usePSPG) stabres = dens * Ft[v][0] + dens * conv + F[1][1 + d] - dens * MRH_FN(FN0 + src(d));
      if (o.useSUPG) for (int e = 0; e < DIM_; ++e) co[1 + e] = co[1 + e] + tau * stabres * u[e] * c.w;
      // [g1] navierstokes.cpp:688: the 3-D z-momentum block writes through the uy offsets
      if (DIM_ == 3 && d == 2 && o.uz_reference) {
        for (int k = 0; k <= DIM_; ++k) Cf[2][k] = Cf[2][k] + co[k];
      } else {
        for (int k = 0; k <= DIM_; ++k) Cf[v][k] = Cf[v][k] + co[k];
      }
      if (o.usePSPG) Cf[1][1 + d] = stabres * (tau * pspg_w);
      divu = divu + F[v][1 + d];
    }
    Cf[1][0] = divu * c.w;
  }
  template <class T, bool STATE = false, int FN0 = 0, int BD0 = NFN, int V0 = 0>
  MRH_HD static void boundary(const QpCtx& c, const GenOpts&, const T (&)[NVAR][4], const T (&)[NVAR][4], T (&Cf)[NVAR][4]) {
    for (int d = 0; d < DIM_; ++d)
      if (c.bc_type[V0 + vel(d)] == BC_NEUMANN) Cf[vel(d)][0] = T(0.0) - MRH_FN(BD0 + vel(d)) * c.w;
  }
};

// ---------------------------------------------------------------------------------------------------------
// Blocks with two modules ("modules: thermal, linearelasticity", "modules: navier stokes, thermal"): the reference runs the modules one
// after the other on the block's workset (volumeResidual / boundaryResidual of every module, assemblyManager_jacres.hpp:362-368,
// 470-480), each writing the residual rows of its own variables; variables and coefficient functions are listed module by module,
// boundary data after all functions (index NFN + variable).  Couplings a module switches on when it finds another module's variable:
//   thermal: have_nsvel (thermal.cpp:117-149, 379) -- the temperature is advected by the Navier-Stokes velocity.
// (linearelasticity's thermo-elastic terms and navierstokes' energy term key on a variable named "e"; the thermal module's variable is
//  "T", so neither is active in these blocks -- linearelasticity.cpp:905, navierstokes.cpp:1031-1045.)
// ---------------------------------------------------------------------------------------------------------
template <int NA, int NB, class T, int NV>
MRH_HD const T (&sub_fields(const T (&F)[NV][4]))[NB][4] { return *reinterpret_cast<const T (*)[NB][4]>(&F[NA]); }
template <int NA, int NB, class T, int NV>
MRH_HD T (&sub_coefs(T (&F)[NV][4]))[NB][4] { return *reinterpret_cast<T (*)[NB][4]>(&F[NA]); }

template <int DIM_, int P_>
struct ThermalElasticityPhys : HGradSystem<DIM_, P_, 1 + DIM_> {   // variables T, dx, dy(, dz)
  typedef ThermalPhys<DIM_, P_> A;
  typedef ElasticityPhys<DIM_, P_> B;
  static constexpr int NVAR = 1 + DIM_, NFN = A::NFN + B::NFN;
  template <class T, bool STATE = false>
  MRH_HD static void volume(const QpCtx& c, const GenOpts& o, const T (&F)[NVAR][4], const T (&Ft)[NVAR][4], T (&Cf)[NVAR][4]) {
    A::template volume<T, STATE, 0, NFN, 0>(c, o, sub_fields<0, 1>(F), sub_fields<0, 1>(Ft), sub_coefs<0, 1>(Cf));
    B::template volume<T, STATE, A::NFN, NFN + 1, 1>(c, o, sub_fields<1, DIM_>(F), sub_fields<1, DIM_>(Ft), sub_coefs<1, DIM_>(Cf));
  }
  template <class T, bool STATE = false>
  MRH_HD static void boundary(const QpCtx& c, const GenOpts& o, const T (&F)[NVAR][4], const T (&Ft)[NVAR][4], T (&Cf)[NVAR][4]) {
    A::template boundary<T, STATE, 0, NFN, 0>(c, o, sub_fields<0, 1>(F), sub_fields<0, 1>(Ft), sub_coefs<0, 1>(Cf));
    B::template boundary<T, STATE, A::NFN, NFN + 1, 1>(c, o, sub_fields<1, DIM_>(F), sub_fields<1, DIM_>(Ft), sub_coefs<1, DIM_>(Cf));
  }
};

template <int DIM_, int P_>
struct NavierStokesThermalPhys : HGradSystem<DIM_, P_, DIM_ + 2> {   // variables ux, pr, uy(, uz), T
  typedef NavierStokesPhys<DIM_, P_> A;
  typedef ThermalPhys<DIM_, P_> B;
  static constexpr int NA = DIM_ + 1, NVAR = DIM_ + 2, NFN = A::NFN + B::NFN;
  template <class T, bool STATE = false>
  MRH_HD static void volume(const QpCtx& c, const GenOpts& o, const T (&F)[NVAR][4], const T (&Ft)[NVAR][4], T (&Cf)[NVAR][4]) {
    A::template volume<T, STATE, 0, NFN, 0>(c, o, sub_fields<0, NA>(F), sub_fields<0, NA>(Ft), sub_coefs<0, NA>(Cf));
    B::template volume<T, STATE, A::NFN, NFN + NA, NA>(c, o, sub_fields<NA, 1>(F), sub_fields<NA, 1>(Ft), sub_coefs<NA, 1>(Cf));
    // have_nsvel (thermal.cpp:139-149): + (u . grad T) v
    T adv = F[A::vel(0)][0] * F[NA][1];
    if (DIM_ > 1) adv = adv + F[A::vel(DIM_ > 1 ? 1 : 0)][0] * F[NA][2];
    if (DIM_ > 2) adv = adv + F[A::vel(DIM_ > 2 ? 2 : 0)][0] * F[NA][3];
    Cf[NA][0] = Cf[NA][0] + adv * c.w;
  }
  template <class T, bool STATE = false>
  MRH_HD static void boundary(const QpCtx& c, const GenOpts& o, const T (&F)[NVAR][4], const T (&Ft)[NVAR][4], T (&Cf)[NVAR][4]) {
    A::template boundary<T, STATE, 0, NFN, 0>(c, o, sub_fields<0, NA>(F), sub_fields<0, NA>(Ft), sub_coefs<0, NA>(Cf));
    B::template boundary<T, STATE, A::NFN, NFN + NA, NA>(c, o, sub_fields<NA, 1>(F), sub_fields<NA, 1>(Ft), sub_coefs<NA, 1>(Cf));
  }
};

// ---------------------------------------------------------------------------------------------------------
// Maxwell 3-D: variables E (HCURL, lowest order, 12 dofs), B (HDIV, lowest order, 6 dofs); functions 0..2 current x,y,z,
// 3 mu, 4 refractive index, 5 epsilon, 6 sigma
// ---------------------------------------------------------------------------------------------------------
struct MaxwellPhys {
  static constexpr int DIM = 3, NVAR = 2, NBASIS = 2, NC = 6, N = 18, NFN = 7;
  static MRH_CE int card(int b) { return b == 0 ? 12 : 6; }
  static MRH_CE int ncb(int b) { return b == 0 ? 6 : 4; }
  static MRH_CE int nval(int) { return 3; }
  static MRH_CE int btype(int b) { return b == 0 ? BT_HCURL : BT_HDIV; }
  static MRH_CE int var_basis(int v) { return v; }
  static MRH_CE int var_basis_rt(int v) { return v; }
  static MRH_CE int ncb_rt(int b) { return b == 0 ? 6 : 4; }
  static MRH_CE int first_var(int b) { return b; }
  static MRH_CE int card_of_var(int v) { return v == 0 ? 12 : 6; }
  static MRH_CE int row0(int v) { return v == 0 ? 0 : 12; }
  static MRH_CE int max_card() { return 12; }
  template <class T, bool STATE = false>
  MRH_HD static void volume(const QpCtx& c, const GenOpts& o, const T (&F)[2][6], const T (&Ft)[2][6], T (&Cf)[2][6]) {
    // B equation: (B_t + curl E) . psi; leap-frog keeps curl E in stage 0 only (maxwell.cpp:138-209)
    const bool with_curl = !o.leapfrog || c.stage == 0;
    for (int d = 0; d < 3; ++d) Cf[1][d] = with_curl ? (Ft[1][d] + F[0][3 + d]) * c.w : Ft[1][d] * c.w;
    // E equation (maxwell.cpp:262-303): (n^2 E_t + (sigma E + J)/eps) . phi - B/(mu eps) . curl phi
    if (!o.leapfrog || c.stage == 1) {
      const auto mu = MRH_FN(3), rindex = MRH_FN(4), eps = MRH_FN(5), sigma = MRH_FN(6);
      const auto ieps = 1.0 / eps, cb = (-1.0 / mu * 1.0 / eps) * c.w, n2 = rindex * rindex;
      for (int d = 0; d < 3; ++d) {
        Cf[0][d] = (n2 * Ft[0][d] + ieps * (sigma * F[0][d] + MRH_FN(d))) * c.w;
        Cf[0][3 + d] = cb * F[1][d];
      }
    }
  }
  template <class T, bool STATE = false>
  MRH_HD static void boundary(const QpCtx& c, const GenOpts&, const T (&F)[2][6], const T (&)[2][6], T (&Cf)[2][6]) {
    if (c.bc_type[1] != BC_NEUMANN) return;   // "really ABC" (maxwell.cpp:341)
    const double gamma = -0.9944;
    const double* n = c.n;
    const T nce[3] = {n[1] * F[0][2] - n[2] * F[0][1], n[2] * F[0][0] - n[0] * F[0][2], n[0] * F[0][1] - n[1] * F[0][0]};
    Cf[1][0] = -(1.0 + gamma) * (n[1] * nce[2] - n[2] * nce[1]) * c.w;
    Cf[1][1] = -(1.0 + gamma) * (n[2] * nce[0] - n[0] * nce[2]) * c.w;
    Cf[1][2] = -(1.0 + gamma) * (n[0] * nce[1] - n[1] * nce[0]) * c.w;
  }
};

}  // namespace mrhyde_b200
