// Launchers of the general element kernel's instantiations.  The instantiation list (general_dispatch.hpp) is compiled in four
// translation units (general_inst0.cu .. general_inst3.cu) so that the build runs them in parallel; general.cu collects the table.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "general.hpp"
#include "general_dispatch.hpp"

namespace mrhyde_b200 {

template <class Phys, int NQ, int NQS, int K, int MAXT, int MINB, bool TCK, bool STATEK>
const char* launch_entry(bool side, const GenParams& P, int nblocks, int threads, size_t smem, void* stream) {
  static size_t attr[2][64] = {};   // opt-in shared memory already granted, per (volume | side, device)
  int devid = 0;
  cudaGetDevice(&devid);
  devid &= 63;
  if (threads > MAXT) return "general element kernel: more threads per CTA than the instantiation's launch bounds";
  const void* fn = side ? (const void*)gen_element_kernel<Phys, NQS, K, true, MAXT, MINB, TCK, STATEK> : (const void*)gen_element_kernel<Phys, NQ, K, false, MAXT, MINB, TCK, STATEK>;
  if (smem > 48 * 1024 && smem > attr[side ? 1 : 0][devid]) {
    const cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    attr[side ? 1 : 0][devid] = smem;
  }
  if (side) gen_element_kernel<Phys, NQS, K, true, MAXT, MINB, TCK, STATEK><<<nblocks, threads, smem, (cudaStream_t)stream>>>(P);
  else gen_element_kernel<Phys, NQ, K, false, MAXT, MINB, TCK, STATEK><<<nblocks, threads, smem, (cudaStream_t)stream>>>(P);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

template <class Phys, int NQ, int NQS, int K, int MAXT, int MINB, int MAXT_L, int MINB_L>
GenDeviceKernels make_device_entry(const char* name, int dim, int order) {
  GenDeviceKernels k;
  k.info = gen_make_info<Phys, NQ, NQS, K>(name, dim, order, MAXT, MINB, MAXT_L, MINB_L);
  k.launch[0] = &launch_entry<Phys, NQ, NQS, K, MAXT_L, MINB_L, false, false>;
  k.launch[1] = &launch_entry<Phys, NQ, NQS, K, MAXT_L, MINB_L, false, true>;
  k.launch_tc[0] = k.launch_tc[1] = nullptr;
  if constexpr (GenLayout<Phys, NQ>::TC_CAPABLE) {
    k.launch_tc[0] = &launch_entry<Phys, NQ, NQS, K, MAXT, MINB, true, false>;
    k.launch_tc[1] = &launch_entry<Phys, NQ, NQS, K, MAXT, MINB, true, true>;
  }
  return k;
}


#define MRH_GEN_PART(FN, LIST)                                                                                                          \
  void FN(std::vector<GenDeviceKernels>& T) {                                                                                           \
    LIST(MRH_GEN_PART_ENTRY)                                                                                                            \
  }
#define MRH_GEN_PART_ENTRY(NAME, DIM, ORDER, NQ, NQS, K, PHYS, MAXT, MINB, MAXT_L, MINB_L) \
  T.push_back(make_device_entry<PHYS, NQ, NQS, K, MAXT, MINB, MAXT_L, MINB_L>(NAME, DIM, ORDER));

}  // namespace mrhyde_b200
