// Plan-specialised kernels: the volume kernel source (volume_kernel.cuh + kernel_abi.h, embedded in this library at
// build time) is compiled per plan by NVRTC for sm_100a with the plan's `Functions:` expressions, reference tables
// and block size as compile-time constants, then loaded with cudaLibraryLoadData.  NVRTC is resolved with dlopen
// so the library loads (and the ahead-of-time kernels work) on a box without it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <cstring>
#include <vector>

#include "volume_launch.hpp"

namespace mrhyde_b200 {

namespace {

struct NvrtcApi {
  void* handle = nullptr;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  const char* (*GetErrorString)(nvrtcResult) = nullptr;
  bool ok = false;
  std::string why;
};

NvrtcApi& api() {
  static NvrtcApi A;
  if (A.handle || !A.why.empty()) return A;
  const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
  std::string errs;
  for (const char* n : names) {
    A.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (A.handle) break;
    errs += std::string(dlerror()) + "; ";
  }
  if (!A.handle) { A.why = "cannot load NVRTC: " + errs; return A; }
  auto sym = [&](const char* n) { void* p = dlsym(A.handle, n); if (!p && A.why.empty()) A.why = std::string("NVRTC lacks ") + n; return p; };
  A.CreateProgram = (decltype(A.CreateProgram))sym("nvrtcCreateProgram");
  A.DestroyProgram = (decltype(A.DestroyProgram))sym("nvrtcDestroyProgram");
  A.CompileProgram = (decltype(A.CompileProgram))sym("nvrtcCompileProgram");
  A.GetCUBINSize = (decltype(A.GetCUBINSize))sym("nvrtcGetCUBINSize");
  A.GetCUBIN = (decltype(A.GetCUBIN))sym("nvrtcGetCUBIN");
  A.GetProgramLogSize = (decltype(A.GetProgramLogSize))sym("nvrtcGetProgramLogSize");
  A.GetProgramLog = (decltype(A.GetProgramLog))sym("nvrtcGetProgramLog");
  A.GetErrorString = (decltype(A.GetErrorString))sym("nvrtcGetErrorString");
  A.ok = A.why.empty();
  return A;
}

}  // namespace

bool nvrtc_available(std::string& why) {
  NvrtcApi& A = api();
  why = A.why;
  return A.ok;
}

bool nvrtc_compile(const std::string& source, int threads, int min_blocks, std::string& cubin, std::string& log, int max_regs) {
  NvrtcApi& A = api();
  if (!A.ok) { log = A.why; return false; }
  nvrtcProgram prog;
  nvrtcResult r = A.CreateProgram(&prog, source.c_str(), "mrhyde_b200_volume_kernel.cu", 0, nullptr, nullptr);
  if (r != NVRTC_SUCCESS) { log = std::string("nvrtcCreateProgram: ") + A.GetErrorString(r); return false; }
  const std::string dthreads = "-DMRH_THREADS=" + std::to_string(threads), dblocks = "-DMRH_MIN_BLOCKS=" + std::to_string(min_blocks);
  const std::string dregs = "--maxrregcount=" + std::to_string(max_regs);
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "-DMRH_JIT=1", dthreads.c_str(), dblocks.c_str(), dregs.c_str()};
  r = A.CompileProgram(prog, (int)(sizeof(opts) / sizeof(opts[0])) - (max_regs > 0 ? 0 : 1), opts);
  size_t n = 0;
  if (A.GetProgramLogSize(prog, &n) == NVRTC_SUCCESS && n > 1) { log.resize(n); A.GetProgramLog(prog, &log[0]); }
  if (r != NVRTC_SUCCESS) { log = std::string("NVRTC compile failed (") + A.GetErrorString(r) + "):\n" + log; A.DestroyProgram(&prog); return false; }
  if (A.GetCUBINSize(prog, &n) != NVRTC_SUCCESS || n == 0) { log = "NVRTC produced no cubin"; A.DestroyProgram(&prog); return false; }
  cubin.resize(n);
  A.GetCUBIN(prog, &cubin[0]);
  A.DestroyProgram(&prog);
  return true;
}

JitKernel::~JitKernel() {
  if (library_) cudaLibraryUnload((cudaLibrary_t)library_);
}

bool JitKernel::build(const std::string& source, const std::string& entry, int threads, int min_blocks, size_t smem, std::string& log, int max_regs) {
  if (!nvrtc_compile(source, threads, min_blocks, cubin_, log, max_regs)) return false;
  cudaLibrary_t lib = nullptr;
  cudaError_t e = cudaLibraryLoadData(&lib, cubin_.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e != cudaSuccess) { log = std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e); return false; }
  cudaKernel_t k = nullptr;
  e = cudaLibraryGetKernel(&k, lib, entry.c_str());
  if (e != cudaSuccess) { log = std::string("cudaLibraryGetKernel(") + entry + "): " + cudaGetErrorString(e); cudaLibraryUnload(lib); return false; }
  e = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { log = std::string("cudaFuncSetAttribute(max dynamic shared memory): ") + cudaGetErrorString(e); cudaLibraryUnload(lib); return false; }
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, (const void*)k) == cudaSuccess) regs_ = fa.numRegs;
  library_ = lib; kernel_ = k;
  return true;
}

const char* JitKernel::launch(const void* params, int grid, int threads, size_t smem, void* stream) const {
  void* args[] = {const_cast<void*>(params)};
  const cudaError_t e = cudaLaunchKernel((const void*)kernel_, dim3((unsigned)grid), dim3((unsigned)threads), args, smem, (cudaStream_t)stream);
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace mrhyde_b200
