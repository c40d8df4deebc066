// Programmatic dependent launch (PDL) for every kernel of the library.
//
// The per-scale work of the pyramid is a long chain of short dependent kernels: at the coarse scales a
// kernel runs for a few microseconds and the gap between two dependent launches — grid scheduling, the prologue of the
// next kernel (barrier init, TMEM allocation, descriptor prefetch) — is as long as the kernel itself.  Launched with
// cudaLaunchAttributeProgrammaticStreamSerialization a kernel's CTAs are scheduled as soon as every CTA of its
// predecessor in the stream has passed pdl_grid_sync(); they run their prologue and then block in griddepcontrol.wait
// until the predecessor has COMPLETED and its memory is visible.  So:
//   * device side: everything that touches global memory comes after pdl_grid_sync(); only shared-memory / TMEM set-up
//     may precede it;
//   * pdl_grid_sync() = wait, then launch_dependents: the run-ahead is one kernel deep (kernel N+2 is scheduled once
//     kernel N+1 has passed its wait, i.e. once kernel N is complete);
//   * a kernel launched without the attribute (HPVG_PDL=0) executes both instructions as no-ops.
// Stream capture turns the attribute into programmatic edges of the CUDA graph.
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>
#include <utility>

namespace hpvg {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_grid_sync() {
  pdl_wait();
  pdl_launch_dependents();
}

// -1: not decided yet (environment HPVG_PDL, default on); 0 / 1: off / on (hpvg_set_pdl).  Defined in api.cu.
extern int g_pdl_mode;
inline bool pdl_enabled() {
  if (g_pdl_mode < 0) {
    const char* e = std::getenv("HPVG_PDL");
    g_pdl_mode = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl_mode != 0;
}

// launch(kernel, grid, block, dynamic smem, stream, args...) — the <<<>>> of this library
template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

}  // namespace hpvg
