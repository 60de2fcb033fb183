// Kernel launch helper: programmatic dependent launch (HML_PDL=0 turns it off everywhere), see modarith.cuh.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace hml {

inline int pdl_enabled() {
  static const int v = [] {
    const char *e = getenv("HML_PDL");
    return e && atoi(e) == 0 ? 0 : 1;
  }();
  return v;
}

// cudaFuncSetAttribute is per device: true exactly once per (call site, device), so a process that drives several GPUs
// opts every one of them into the large dynamic shared-memory carve-out
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return true;
    const bool f = !done[dev];
    done[dev] = true;
    return f;
  }
};

template <class... KArgs, class... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled();
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace hml
