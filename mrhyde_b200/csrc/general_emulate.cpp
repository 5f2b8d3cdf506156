// Host replay of the general element kernel's stage functions (general_kernel.cuh compiled by the host compiler).
// TEST-ONLY LIBRARY: this file is NOT part of libmrhyde_b200.so.  It builds into libmrhyde_b200_emulate.so, which tests load next to
// the product library and register with mrhyde_b200_debug_set_emulator; the mrhyde_b200_plan_debug_emulate* entry points (host-only
// plans, device = -1) then replay the kernel stages on the CPU.  mrhyde_b200_assemble_* never reaches it -- there is no CPU fallback.
#include <cstring>
#include <vector>

#include "general.hpp"
#include "general_dispatch.hpp"

namespace mrhyde_b200 {

namespace {

template <class Phys, int NQ, int K, bool SIDE, bool TCK, bool STATEK>
void emulate_blocks(const GenParams& P, int nblocks) {
  typedef GenBlock<Phys, NQ, K, SIDE, TCK, STATEK> Bk;
  typedef GenLayout<Phys, NQ, TCK> L;
  std::vector<double> sm((size_t)P.epb * L::SIZE);
  for (int blk = 0; blk < nblocks; ++blk) {
    std::fill(sm.begin(), sm.end(), 0.0);
    const int n0 = P.epb * (L::N > L::NV ? L::N : L::NV);
    for (int i = 0; i < n0; ++i) Bk::s0(P, sm.data(), blk, i);
    for (int i = 0; i < P.epb * NQ; ++i) Bk::s1(P, sm.data(), blk, i);
    for (int i = 0; i < P.epb; ++i) Bk::s1b(P, sm.data(), blk, i);
    for (int i = 0; i < P.epb * Phys::max_card() * NQ; ++i) Bk::s2(P, sm.data(), blk, i);
    for (int i = 0; i < P.epb * NQ * Bk::S3_KINDS; ++i) Bk::s3(P, sm.data(), blk, i);
    for (int i = 0; i < P.epb * NQ * Bk::S3F_KINDS; ++i) Bk::s3f(P, sm.data(), blk, i);
    for (int i = 0; i < P.epb * NQ; ++i) Bk::s4a(P, sm.data(), blk, i);
    if (P.elem_jac) {
      if constexpr (Bk::TC) {   // field-direction derivatives + contraction (the device runs the contraction on the FP64 tensor cores)
        for (int i = 0; i < P.epb * NQ * Bk::NCV; ++i) Bk::s4d(P, sm.data(), blk, i);
        for (int i = 0; i < P.epb * L::NVAR * L::NVAR; ++i) Bk::s4m_item(P, sm.data(), blk, i);
      } else {
        for (int i = 0; i < P.epb * Bk::TPE; ++i) Bk::s4b(P, sm.data(), blk, i);
      }
    }
    for (int i = 0; i < P.epb * L::N; ++i) Bk::s5(P, sm.data(), blk, i);
  }
}

template <class Phys, int NQ, int NQS, int K>
void emulate_entry(bool side, const GenParams& P, int nblocks) {
  // replay the stages of the build the plan would launch: option jacobian (tensor | lanes) x state-dependent coefficients (or not)
  const bool tc = P.tensor && GenLayout<Phys, NQ>::TC_CAPABLE, st = P.fn_state != 0;
  if (tc) {
    if constexpr (GenLayout<Phys, NQ>::TC_CAPABLE) {
      if (st) { if (side) emulate_blocks<Phys, NQS, K, true, true, true>(P, nblocks); else emulate_blocks<Phys, NQ, K, false, true, true>(P, nblocks); }
      else { if (side) emulate_blocks<Phys, NQS, K, true, true, false>(P, nblocks); else emulate_blocks<Phys, NQ, K, false, true, false>(P, nblocks); }
    }
    return;
  }
  if (st) { if (side) emulate_blocks<Phys, NQS, K, true, false, true>(P, nblocks); else emulate_blocks<Phys, NQ, K, false, false, true>(P, nblocks); }
  else { if (side) emulate_blocks<Phys, NQS, K, true, false, false>(P, nblocks); else emulate_blocks<Phys, NQ, K, false, false, false>(P, nblocks); }
}

}  // namespace

std::vector<GenHostKernels>& host_table() {
  static std::vector<GenHostKernels> T;
  if (T.empty()) {
#define X(NAME, DIM, ORDER, NQ, NQS, K, PHYS, MAXT, MINB, MAXT_L, MINB_L) \
    T.push_back(GenHostKernels{gen_make_info<PHYS, NQ, NQS, K>(NAME, DIM, ORDER, MAXT, MINB, MAXT_L, MINB_L), &emulate_entry<PHYS, NQ, NQS, K>});
    MRH_GEN_LIST(X)
#undef X
  }
  return T;
}

}  // namespace mrhyde_b200

// what mrhyde_b200_debug_set_emulator takes: (physics, dim, order, volume points, side points) -> const GenHostKernels* or null
extern "C" const void* mrhyde_b200_emulator_lookup(const char* physics, int dim, int order, int nq, int nqs) {
  for (auto& k : mrhyde_b200::host_table())
    if (std::string(physics) == k.info.physics && dim == k.info.dim && order == k.info.order && nq == k.info.nq && (nqs == 0 || nqs == k.info.nqs)) return &k;
  return nullptr;
}
