// Ahead-of-time build of the thermal Q1 volume kernels (volume_kernel.cuh) with the bytecode expression
// interpreter: used when a plan is not specialised through NVRTC (option "jit" = "false", or NVRTC missing).
#include <cuda_runtime.h>

#define MRH_DEFINE_KERNELS
#include "volume_kernel.cuh"
#include "volume_launch.hpp"

namespace mrhyde_b200 {

const char* launch_thermal_q1_aot(int dim, const void* params, int n_chains, int threads, size_t smem, void* stream) {
  static size_t attr_smem[2] = {0, 0};
  const void* fn = dim == 3 ? (const void*)mrh_thermal_q1_3d : (const void*)mrh_thermal_q1_2d;
  if (threads > MRH_THREADS) return "thermal kernel: more threads per block than the ahead-of-time build allows";
  if (smem > attr_smem[dim - 2]) {
    const cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    attr_smem[dim - 2] = smem;
  }
  if (dim == 3) mrh_thermal_q1_3d<<<n_chains, threads, smem, (cudaStream_t)stream>>>(*(const ThermalParams<3>*)params);
  else mrh_thermal_q1_2d<<<n_chains, threads, smem, (cudaStream_t)stream>>>(*(const ThermalParams<2>*)params);
  return nullptr;
}

}  // namespace mrhyde_b200
