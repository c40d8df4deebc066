// Deterministic cross-block reduction ("the last block to finish sums every block's partial in BLOCK ORDER").
// Replaces fp64 atomicAdd accumulation, whose result depends on the order in which blocks happen to arrive (differences
// at the 1e-16 level that Adam then amplifies wherever a gradient element is tiny): with this, the BatchNorm statistics,
// the BatchNorm-backward sums and the bias-gradient column sums are bitwise reproducible from run to run.
#pragma once
#include <cuda_runtime.h>

namespace hpvg {

constexpr int DET_MAX_BLOCKS = 148 * 8;   // capacity of the scratch
// grid cap of the streaming reductions: the last block's final pass is one L2 round trip per 32 blocks, and 4 resident
// blocks per SM with 4-8 loads in flight per thread already saturate HBM
constexpr int DET_STREAM_BLOCKS = 148 * 4;   // (and >= 4 voxel steps per block: a 7 752-voxel tensor runs 61 blocks, a 2-step final pass)
struct DetScratch {
  double* partials;        // [DET_MAX_BLOCKS][128], per stream (csrc/api.cu StreamCtx)
  unsigned int* counter;   // zero before the launch, left at zero
};

// Sum of partials[b][col] over b = first, first + step, ... < nblocks, taken 16 rows at a time: the 16 loads of a step are
// independent (one L2 round trip per step instead of one per row) and the 16 chains are combined in a fixed tree.
__device__ __forceinline__ double det_column_sum(const double* col, unsigned int nblocks, unsigned int first_chunk,
                                                 unsigned int chunk_step) {
  double a[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = 0.0;
  for (unsigned int b = first_chunk * 16u; b < nblocks; b += chunk_step * 16u) {
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (b + k < nblocks) a[k] += __ldcg(col + static_cast<size_t>(b + k) * 128);
  }
#pragma unroll
  for (int w = 8; w >= 1; w >>= 1) {
#pragma unroll
    for (int k = 0; k < w; ++k) a[k] += a[k + w];
  }
  return a[0];
}

// Called by the same 128 threads (t = 0..127, `active`) of EVERY block of the grid, convergently.
//   NAMED = false: the whole block (>= 256 threads) calls it (other threads pass active = false); synchronises with
//                  __syncthreads; threads 128..255 help with the final pass
//   NAMED = true : exactly 128 threads call it; synchronises with the named barrier 1
// value: this block's partial of statistic t.  out[t] (t < 64: out_lo, else out_hi) (+)= sum over blocks, in a fixed order
// (whichever block happens to be last): bitwise reproducible.
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
  if (!s_last) return;
  __threadfence();
  if constexpr (NAMED) {
    const double acc = det_column_sum(s.partials + t, nblocks, 0u, 1u);
    double* o = (t < 64) ? out_lo + t : out_hi + (t - 64);
    *o = accumulate ? *o + acc : acc;
    if (t == 0) *s.counter = 0u;
  } else {
    // columns 0..127 x two interleaved halves of the 16-row chunks (threads 0..127 / 128..255)
    __shared__ double s_half[128];
    const unsigned int tid = threadIdx.x;
    double acc = 0.0;
    if (tid < 256) acc = det_column_sum(s.partials + (tid & 127u), nblocks, tid >> 7, 2u);
    if (tid >= 128 && tid < 256) s_half[tid - 128] = acc;
    __syncthreads();
    if (tid < 128) {
      acc += s_half[tid];
      double* o = (tid < 64) ? out_lo + tid : out_hi + (tid - 64);
      *o = accumulate ? *o + acc : acc;
      if (tid == 0) *s.counter = 0u;
    }
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
