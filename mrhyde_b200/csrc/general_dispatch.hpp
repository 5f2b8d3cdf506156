// The compiled instantiations of the general element kernel: (module, dimension, basis order, volume points, side points).
// Included by general.cu (device launchers) and general_emulate.cpp (host replay of the same stage functions).
#pragma once
#include "general_physics.cuh"

namespace mrhyde_b200 {
typedef ThermalPhys<2, 1> GenTh21;
typedef ThermalPhys<3, 1> GenTh31;
typedef ThermalPhys<2, 2> GenTh22;
typedef ThermalPhys<3, 2> GenTh32;
typedef ElasticityPhys<2, 1> GenLe21;
typedef ElasticityPhys<3, 1> GenLe31;
typedef ElasticityPhys<2, 2> GenLe22;
typedef ElasticityPhys<3, 2> GenLe32;
typedef NavierStokesPhys<2, 1> GenNs21;
typedef NavierStokesPhys<3, 1> GenNs31;
typedef ThermalElasticityPhys<2, 1> GenThLe21;
typedef ThermalElasticityPhys<3, 1> GenThLe31;
typedef NavierStokesThermalPhys<2, 1> GenNsTh21;
typedef NavierStokesThermalPhys<3, 1> GenNsTh31;
}  // namespace mrhyde_b200

// X(physics name, dim, order, NQ, NQS, K, Phys, MAXT, MINB, MAXT_L, MINB_L): launch bounds (threads per CTA at most, CTAs per SM at
// least) of the tensor-core build (S4d / S4m: CTAs hold as many elements as the shared memory of MINB resident CTAs allows and MAXT
// threads whatever the element size, one warp per (element, variable pair) block) and of the derivative-lane build (S4b: one thread
// per element dof).  The two-basis Maxwell layout has the lane build only.
// four parts of roughly equal compile time, one translation unit each (general_inst0.cu .. general_inst3.cu)
#define MRH_GEN_LIST_0(X)                                                   \
  X("thermal", 2, 1, 4, 2, 1, GenTh21, 256, 3, 256, 3)                      \
  X("thermal", 3, 1, 8, 4, 1, GenTh31, 256, 3, 256, 3)                      \
  X("thermal", 2, 2, 9, 3, 1, GenTh22, 256, 3, 256, 2)                      \
  X("thermal", 3, 2, 27, 9, 1, GenTh32, 128, 4, 128, 3)                     \
  X("maxwell", 3, 1, 8, 4, 1, MaxwellPhys, 128, 3, 128, 3)
#define MRH_GEN_LIST_1(X)                                                   \
  X("linearelasticity", 2, 1, 4, 2, 1, GenLe21, 256, 3, 256, 3)             \
  X("linearelasticity", 3, 1, 8, 4, 1, GenLe31, 256, 3, 256, 3)             \
  X("linearelasticity", 2, 2, 9, 3, 1, GenLe22, 256, 3, 128, 4)             \
  X("navier stokes", 2, 1, 4, 2, 1, GenNs21, 256, 3, 256, 2)
#define MRH_GEN_LIST_2(X)                                                   \
  X("linearelasticity", 3, 2, 27, 9, 1, GenLe32, 576, 1, 96, 2)             \
  X("navier stokes", 3, 1, 8, 4, 1, GenNs31, 256, 2, 128, 3)
#define MRH_GEN_LIST_3(X)                                                   \
  X("thermal+linearelasticity", 2, 1, 4, 2, 1, GenThLe21, 256, 3, 256, 3)   \
  X("thermal+linearelasticity", 3, 1, 8, 4, 1, GenThLe31, 256, 2, 256, 2)   \
  X("navier stokes+thermal", 2, 1, 4, 2, 1, GenNsTh21, 256, 2, 256, 2)      \
  X("navier stokes+thermal", 3, 1, 8, 4, 1, GenNsTh31, 256, 2, 128, 3)
#define MRH_GEN_LIST(X) MRH_GEN_LIST_0(X) MRH_GEN_LIST_1(X) MRH_GEN_LIST_2(X) MRH_GEN_LIST_3(X)

namespace mrhyde_b200 {
template <class Phys, int NQ, int NQS, int K>
inline GenKernelInfo gen_make_info(const char* physics, int dim, int order, int maxt, int minb, int maxt_l, int minb_l) {
  GenKernelInfo I;
  I.physics = physics; I.dim = dim; I.order = order; I.nq = NQ; I.nqs = NQS;
  I.tc_max_threads = maxt; I.tc_min_blocks = minb;
  I.max_threads = maxt_l; I.min_blocks = minb_l;
  I.N = Phys::N; I.nvars = Phys::NVAR; I.nbasis = Phys::NBASIS; I.nfn = Phys::NFN; I.K = K; I.tpe = GenBlock<Phys, NQ, K, false>::TPE; I.tensor = GenLayout<Phys, NQ>::TC_CAPABLE ? 1 : 0;
  I.smem_doubles_volume = GenLayout<Phys, NQ, false>::SIZE;
  I.smem_doubles_side = GenLayout<Phys, NQS, false>::SIZE;
  I.tc_smem_doubles_volume = GenLayout<Phys, NQ, true>::SIZE;
  I.tc_smem_doubles_side = GenLayout<Phys, NQS, true>::SIZE;
  for (int b = 0; b < 2; ++b) { I.card[b] = b < Phys::NBASIS ? Phys::card(b) : 0; I.ncb[b] = b < Phys::NBASIS ? Phys::ncb(b) : 0; }
  for (int v = 0; v < GEN_MAXVARS; ++v) I.var_basis[v] = v < Phys::NVAR ? Phys::var_basis(v) : 0;
  return I;
}
}  // namespace mrhyde_b200
