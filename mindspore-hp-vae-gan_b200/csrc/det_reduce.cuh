// Deterministic cross-block reduction ("the last block to finish sums every block's partial in BLOCK ORDER").
// Replaces fp64 atomicAdd accumulation, whose result depends on the order in which blocks happen to arrive (differences
// at the 1e-16 level that Adam then amplifies wherever a gradient element is tiny): with this, the BatchNorm statistics,
// the BatchNorm-backward sums and the bias-gradient column sums are bitwise reproducible from run to run.
#pragma once
#include <cuda_runtime.h>

namespace hpvg {

constexpr int DET_MAX_BLOCKS = 148 * 8;
struct DetScratch {
  double* partials;        // [DET_MAX_BLOCKS][128], per stream (csrc/api.cu StreamCtx)
  unsigned int* counter;   // zero before the launch, left at zero
};

// Called by the same 128 threads (t = 0..127, `active`) of EVERY block of the grid, convergently.
//   NAMED = false: the whole block calls it (other threads pass active = false); synchronises with __syncthreads
//   NAMED = true : exactly 128 threads call it; synchronises with the named barrier 1
// value: this block's partial of statistic t.  out[t] (t < 64: out_lo, else out_hi) (+)= sum over blocks, in block order.
template <bool NAMED>
__device__ __forceinline__ void det_reduce_128(double value, int t, bool active, unsigned int block_id,
                                               unsigned int nblocks, DetScratch s, double* out_lo, double* out_hi,
                                               bool accumulate) {
  __shared__ unsigned int s_last;
  if (active) s.partials[static_cast<size_t>(block_id) * 128 + t] = value;
  __threadfence();
  if constexpr (NAMED) asm volatile("bar.sync 1, 128;" ::: "memory");
  else __syncthreads();
  if (active && t == 0) s_last = (atomicAdd(s.counter, 1u) == nblocks - 1u) ? 1u : 0u;
  if constexpr (NAMED) asm volatile("bar.sync 1, 128;" ::: "memory");
  else __syncthreads();
  if (s_last && active) {
    __threadfence();
    // four interleaved chains (rows b % 4), combined in a fixed order: the loads of a chain step are independent, so
    // the tail of a launch-latency-bound kernel is nblocks / 4 dependent adds instead of nblocks dependent L2 loads
    const double* p = s.partials + t;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    unsigned int b = 0;
    for (; b + 4 <= nblocks; b += 4) {
      a0 += __ldcg(p + static_cast<size_t>(b) * 128);
      a1 += __ldcg(p + static_cast<size_t>(b + 1) * 128);
      a2 += __ldcg(p + static_cast<size_t>(b + 2) * 128);
      a3 += __ldcg(p + static_cast<size_t>(b + 3) * 128);
    }
    for (; b < nblocks; ++b) a0 += __ldcg(p + static_cast<size_t>(b) * 128);
    const double acc = (a0 + a1) + (a2 + a3);
    double* o = (t < 64) ? out_lo + t : out_hi + (t - 64);
    *o = accumulate ? *o + acc : acc;
    if (t == 0) *s.counter = 0u;
  }
}

// Scalar twin: every block contributes one value (thread 0); the last block writes out = scale * sum (fixed order).
__device__ __forceinline__ void det_reduce_scalar(float value, unsigned int block_id, unsigned int nblocks, DetScratch s,
                                                  float scale, float* out) {
  __shared__ unsigned int s_last1;
  if (threadIdx.x == 0) {
    s.partials[block_id] = static_cast<double>(value);
    __threadfence();
    s_last1 = (atomicAdd(s.counter, 1u) == nblocks - 1u) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last1) {
    // the whole (last) block: thread i sums partials i, i + blockDim, ... ; a fixed-order tree over shared memory
    // finishes — same result whatever block happens to be last
    __shared__ double s_tree[256];
    __threadfence();
    double acc = 0.0;
    for (unsigned int b = threadIdx.x; b < nblocks; b += blockDim.x) acc += __ldcg(s.partials + b);
    if (threadIdx.x < 256) s_tree[threadIdx.x] = acc;
    __syncthreads();
    for (unsigned int w = 128; w >= 1; w >>= 1) {
      if (threadIdx.x < w && threadIdx.x + w < blockDim.x) s_tree[threadIdx.x] += s_tree[threadIdx.x + w];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      *out = static_cast<float>(s_tree[0]) * scale;
      *s.counter = 0u;
    }
  }
}

}  // namespace hpvg
