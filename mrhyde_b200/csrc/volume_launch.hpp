// Launchers of the volume kernel: the ahead-of-time build (thermal.cu) and the plan-specialised NVRTC build (jit.cu).
#pragma once
#include <cstddef>
#include <string>

namespace mrhyde_b200 {

// returns nullptr on success, else a static error string
const char* launch_thermal_q1_aot(int dim, const void* params, int n_chains, int threads, size_t smem, void* stream);

// One NVRTC-compiled kernel, loaded into the current device's context.
class JitKernel {
 public:
  ~JitKernel();
  // source: complete CUDA C++ translation unit; entry: extern "C" kernel name.  On failure returns false and
  // fills `log` (compiler output or the loader error).
  bool build(const std::string& source, const std::string& entry, int threads, int min_blocks, size_t smem, std::string& log, int max_regs = 0);
  bool ready() const { return kernel_ != nullptr; }
  const char* launch(const void* params, int grid, int threads, size_t smem, void* stream) const;
  int regs() const { return regs_; }
  const std::string& cubin() const { return cubin_; }

 private:
  void* library_ = nullptr;  // cudaLibrary_t
  void* kernel_ = nullptr;   // cudaKernel_t
  int regs_ = 0;
  std::string cubin_;
};

bool nvrtc_available(std::string& why);
// NVRTC only (no device needed): used by the CPU test-suite to check that generated sources compile for sm_100a.
bool nvrtc_compile(const std::string& source, int threads, int min_blocks, std::string& cubin, std::string& log, int max_regs = 0);

}  // namespace mrhyde_b200
