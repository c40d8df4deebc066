// fp32 channels-last twins of the bandwidth-bound kernels that touch activations (tf32 precision mode: the tensors the
// kind::tf32 convolutions read and write).  Same operations, same reference call sites as their bf16 versions in
// elementwise.cu: training-mode BatchNorm3d + LeakyReLU (networks_3d.py:52-53) forward / backward, LeakyReLU backward,
// bias-gradient column sums, and the NCDHW <-> channels-last layout change at the API edge.
// Every value that a convolution or weight-gradient kernel will read as an operand is rounded to tf32 (round to nearest)
// when it is stored: the tensor core ignores the low 13 mantissa bits, i.e. it would otherwise truncate.
#include <cuda_runtime.h>
#include <cstdint>

#include "det_reduce.cuh"
#include "launch.cuh"
#include "elementwise.h"

namespace hpvg {

namespace {

__device__ __forceinline__ float rtf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__device__ __forceinline__ float4 rtf32(float4 v) { return make_float4(rtf32(v.x), rtf32(v.y), rtf32(v.z), rtf32(v.w)); }

inline int grid_for(long long n, int block, int cap = 148 * 16) {
  long long b = (n + block - 1) / block;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

// ----------------------------------------------------------------------------------------------- pack / unpack
// Narrow source (C <= 4: clips, 3-channel gradients) -> one 16-byte voxel (channels >= C zero).
__global__ void pack_cl_f32_skinny_kernel(const float* __restrict__ x, int C, long long sp, long long voxels,
                                          float* __restrict__ y, int c_pitch, int c_off, int groups) {
  pdl_grid_sync();
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= voxels * groups) return;
  const long long v = gid / groups;
  const int g = static_cast<int>(gid - v * groups);
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  if (g == 0) {
    const long long n = v / sp, s = v - n * sp;
    float f[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) f[e] = (e < C) ? rtf32(__ldg(x + (n * C + e) * sp + s)) : 0.f;
    o = make_float4(f[0], f[1], f[2], f[3]);
  }
  *reinterpret_cast<float4*>(y + v * c_pitch + c_off + g * 4) = o;
}

// Wide tensors: a 32-voxel x 32-channel tile through shared memory so both sides are coalesced (fp32 NCDHW side: 128-byte
// runs of one channel; channels-last side: 128-byte runs of one voxel).
__global__ void __launch_bounds__(256)
pack_cl_f32_tiled_kernel(const float* __restrict__ x, int C, long long sp, long long voxels, float* __restrict__ y,
                         int c_pitch, int c_off, int c_fill /* channels [c_off, c_off + c_fill) are written */) {
  pdl_grid_sync();
  __shared__ float tile[32][33];
  const int cb = blockIdx.y * 32;
  const long long v0 = static_cast<long long>(blockIdx.x) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  {
    const long long v = v0 + tx;
    if (v < voxels) {
      const long long n = v / sp, s = v - n * sp;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = cb + ty + 8 * j;
        tile[ty + 8 * j][tx] = (c < C) ? __ldg(x + (n * C + c) * sp + s) : 0.f;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long v = v0 + ty + 8 * j;
    const int c = cb + tx;
    if (v < voxels && c < c_fill) y[v * c_pitch + c_off + c] = rtf32(tile[tx][ty + 8 * j]);
  }
}

__global__ void __launch_bounds__(256)
unpack_cl_f32_tiled_kernel(const float* __restrict__ x, int C, long long sp, long long voxels, int c_pitch, int c_off,
                           float* __restrict__ y) {
  pdl_grid_sync();
  __shared__ float tile[32][33];
  const int cb = blockIdx.y * 32;
  const long long v0 = static_cast<long long>(blockIdx.x) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long v = v0 + ty + 8 * j;
    const int c = cb + tx;
    if (v < voxels && c < C) tile[ty + 8 * j][tx] = x[v * c_pitch + c_off + c];
  }
  __syncthreads();
  const long long v = v0 + tx;
  if (v >= voxels) return;
  const long long n = v / sp, s = v - n * sp;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = cb + ty + 8 * j;
    if (c < C) y[(n * C + c) * sp + s] = tile[tx][ty + 8 * j];
  }
}

// ----------------------------------------------------------------------------------------------- BatchNorm (train)
// y: (voxels, 64) fp32.  thread -> one float4 channel group; 16 groups per voxel; block = 256 threads = 16 voxels/iter.
// sums[0][c] += sum y, sums[1][c] += sum y^2
__global__ void bn_stats_cl_f32_kernel(const float* __restrict__ y, long long voxels, double* __restrict__ sum,
                                       double* __restrict__ sumsq, const DetScratch det) {
  pdl_grid_sync();
  const int g = threadIdx.x & 15, vl = threadIdx.x >> 4;
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long v = static_cast<long long>(blockIdx.x) * 16 + vl; v < voxels;
       v += static_cast<long long>(gridDim.x) * 16) {
    const float4 f = *reinterpret_cast<const float4*>(y + v * 64 + g * 4);
    a[0] += f.x; a[1] += f.y; a[2] += f.z; a[3] += f.w;
    b[0] = fmaf(f.x, f.x, b[0]); b[1] = fmaf(f.y, f.y, b[1]); b[2] = fmaf(f.z, f.z, b[2]); b[3] = fmaf(f.w, f.w, b[3]);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    a[e] += __shfl_xor_sync(0xffffffffu, a[e], 16);
    b[e] += __shfl_xor_sync(0xffffffffu, b[e], 16);
  }
  __shared__ float red[8][2][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < 16) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      red[warp][0][lane * 4 + e] = a[e];
      red[warp][1][lane * 4 + e] = b[e];
    }
  }
  __syncthreads();
  double acc = 0.0;
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) acc += static_cast<double>(red[wv][which][c]);
  }
  det_reduce_128<false>(acc, threadIdx.x, threadIdx.x < 128, blockIdx.x, gridDim.x, det, sum, sumsq, false);
}

__device__ __forceinline__ float4 affine_act(float4 f, const float* sc, const float* sh, int c, int act) {
  float v[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[e] = fmaf(v[e], sc[c + e], sh[c + e]);
    if (act == 1) v[e] = v[e] > 0.f ? v[e] : 0.2f * v[e];
    v[e] = rtf32(v[e]);
  }
  return make_float4(v[0], v[1], v[2], v[3]);
}

__global__ void bn_apply_cl_f32_kernel(const float* __restrict__ y, long long groups /*voxels*16*/,
                                       const float* __restrict__ scale, const float* __restrict__ shift, int act,
                                       float* __restrict__ x) {
  pdl_grid_sync();
  __shared__ float sc[64], sh[64];
  if (threadIdx.x < 64) {
    sc[threadIdx.x] = scale[threadIdx.x];
    sh[threadIdx.x] = shift[threadIdx.x];
  }
  __syncthreads();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < groups;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i & 15) * 4;
    const float4 f = *reinterpret_cast<const float4*>(y + i * 4);
    *reinterpret_cast<float4*>(x + i * 4) = affine_act(f, sc, sh, c, act);
  }
}

// One-pass training-mode BatchNorm (statistics from the producing conv's epilogue); see bn_train_apply_cl_kernel.
__global__ void bn_train_apply_cl_f32_kernel(const float* __restrict__ y, long long groups /*voxels*16*/,
                                             const double* __restrict__ sums /*[2][64]*/, double count,
                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                             float eps, float momentum, float* __restrict__ mm,
                                             float* __restrict__ mv, float* __restrict__ saved, int act,
                                             float* __restrict__ x, const float* __restrict__ center) {
  pdl_grid_sync();
  __shared__ float sc[64], sh[64];
  if (threadIdx.x < 64) {
    const int c = threadIdx.x;
    const double mean = sums[c] / count;
    double var = sums[64 + c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    const float s = gamma[c] * invstd;
    const float b = beta[c] - static_cast<float>(mean) * s;
    sc[c] = s;
    sh[c] = b;
    if (blockIdx.x == 0) {
      if (saved) {
        saved[c] = s;
        saved[64 + c] = b;
        saved[128 + c] = static_cast<float>(mean);
        saved[192 + c] = invstd;
      }
      if (mm) mm[c] = momentum * mm[c] + (1.f - momentum) * (static_cast<float>(mean) + (center ? center[c] : 0.f));
      if (mv) mv[c] = momentum * mv[c] + (1.f - momentum) * static_cast<float>(var);
    }
  }
  __syncthreads();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < groups;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i & 15) * 4;
    const float4 f = *reinterpret_cast<const float4*>(y + i * 4);
    *reinterpret_cast<float4*>(x + i * 4) = affine_act(f, sc, sh, c, act);
  }
}

// ----------------------------------------------------------------------------------------------- backward pieces
__global__ void lrelu_bwd_cl_f32_kernel(const float* __restrict__ ga, const float* __restrict__ a, long long groups,
                                        float* __restrict__ gz) {
  pdl_grid_sync();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < groups;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 g = *reinterpret_cast<const float4*>(ga + i * 4);
    const float4 av = *reinterpret_cast<const float4*>(a + i * 4);
    g.x = av.x > 0.f ? g.x : 0.2f * g.x;
    g.y = av.y > 0.f ? g.y : 0.2f * g.y;
    g.z = av.z > 0.f ? g.z : 0.2f * g.z;
    g.w = av.w > 0.f ? g.w : 0.2f * g.w;
    *reinterpret_cast<float4*>(gz + i * 4) = rtf32(g);
  }
}

// BatchNorm(train)+LeakyReLU backward, pass 1: sums[0][c] = sum gz, sums[1][c] = sum gz*xhat
__global__ void bn_bwd_reduce_cl_f32_kernel(const float* __restrict__ ga, const float* __restrict__ y,
                                            long long voxels, const float* __restrict__ saved /*[4][64]*/, int act,
                                            double* __restrict__ sums, const DetScratch det) {
  pdl_grid_sync();
  const int g = threadIdx.x & 15, vl = threadIdx.x >> 4;
  float sc[4], sh[4], mu[4], is[4], s0[4], s1[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    sc[e] = saved[g * 4 + e];
    sh[e] = saved[64 + g * 4 + e];
    mu[e] = saved[128 + g * 4 + e];
    is[e] = saved[192 + g * 4 + e];
    s0[e] = s1[e] = 0.f;
  }
  for (long long v = static_cast<long long>(blockIdx.x) * 16 + vl; v < voxels;
       v += static_cast<long long>(gridDim.x) * 16) {
    const float4 g4 = *reinterpret_cast<const float4*>(ga + v * 64 + g * 4);
    const float4 y4 = *reinterpret_cast<const float4*>(y + v * 64 + g * 4);
    const float gv4[4] = {g4.x, g4.y, g4.z, g4.w}, yv4[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float gv = gv4[e];
      if (act == 1 && fmaf(yv4[e], sc[e], sh[e]) <= 0.f) gv *= 0.2f;
      s0[e] += gv;
      s1[e] = fmaf(gv, (yv4[e] - mu[e]) * is[e], s1[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    s0[e] += __shfl_xor_sync(0xffffffffu, s0[e], 16);
    s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], 16);
  }
  __shared__ float red[8][2][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < 16) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      red[warp][0][lane * 4 + e] = s0[e];
      red[warp][1][lane * 4 + e] = s1[e];
    }
  }
  __syncthreads();
  double acc = 0.0;
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) acc += static_cast<double>(red[wv][which][c]);
  }
  det_reduce_128<false>(acc, threadIdx.x, threadIdx.x < 128, blockIdx.x, gridDim.x, det, sums, sums + 64, false);
}

// fp32 -> tf32 with STOCHASTIC rounding (13 hashed dither bits, keyed by the element index: reproducible): the stored
// gradient stays unbiased, so the per-channel mean corrections below survive (see bn_bwd_apply_cl_kernel in
// elementwise.cu for the measurement that motivated this in the bf16 path).
__device__ __forceinline__ float tf32_sr(float v, uint32_t key) {
  key ^= key >> 16;
  key *= 0x7feb352dU;
  key ^= key >> 15;
  key *= 0x846ca68bU;
  key ^= key >> 16;
  return __uint_as_float((__float_as_uint(v) + (key & 0x1FFFu)) & 0xFFFFE000u);
}

// pass 2: gy = gamma*invstd*(gz - mean(gz) - xhat*mean(gz*xhat))
__global__ void bn_bwd_apply_cl_f32_kernel(const float* __restrict__ ga, const float* __restrict__ y, long long groups,
                                           const float* __restrict__ saved, int act, const double* __restrict__ sums,
                                           double inv_count, float* __restrict__ gy, float* __restrict__ dgamma,
                                           float* __restrict__ dbeta, int accumulate) {
  pdl_grid_sync();
  __shared__ float sc[64], sh[64], mu[64], is[64], m0[64], m1[64];
  if (threadIdx.x < 64) {
    const int c = threadIdx.x;
    sc[c] = saved[c];
    sh[c] = saved[64 + c];
    mu[c] = saved[128 + c];
    is[c] = saved[192 + c];
    m0[c] = static_cast<float>(sums[c] * inv_count);
    m1[c] = static_cast<float>(sums[64 + c] * inv_count);
    if (blockIdx.x == 0) {
      if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + static_cast<float>(sums[c]);
      if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + static_cast<float>(sums[64 + c]);
    }
  }
  __syncthreads();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < groups;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cb = static_cast<int>(i & 15) * 4;
    const float4 g4 = *reinterpret_cast<const float4*>(ga + i * 4);
    const float4 y4 = *reinterpret_cast<const float4*>(y + i * 4);
    const float gv4[4] = {g4.x, g4.y, g4.z, g4.w}, yv4[4] = {y4.x, y4.y, y4.z, y4.w};
    float r[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = cb + e;
      float gv = gv4[e];
      if (act == 1 && fmaf(yv4[e], sc[c], sh[c]) <= 0.f) gv *= 0.2f;
      const float xh = (yv4[e] - mu[c]) * is[c];
      r[e] = tf32_sr(sc[c] * (gv - m0[c] - xh * m1[c]), static_cast<uint32_t>(i) * 4u + e);
    }
    *reinterpret_cast<float4*>(gy + i * 4) = make_float4(r[0], r[1], r[2], r[3]);
  }
}

__global__ void d2f_f32_kernel(const double* __restrict__ in, int n, float scale, int accumulate,
                               float* __restrict__ out) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (accumulate ? out[i] : 0.f) + scale * static_cast<float>(in[i]);
}

}  // namespace

#define LAUNCH_CHECK()                          \
  do {                                          \
    cudaError_t e_ = cudaGetLastError();        \
    if (e_ != cudaSuccess) return e_;           \
  } while (0)

cudaError_t ew_pack_cl_f32(const float* x, int N, int C, long long sp, float* y, int c_pitch, int c_off,
                           int c_zero_to, cudaStream_t st) {
  const long long voxels = static_cast<long long>(N) * sp;
  const int fill = (c_zero_to - c_off > C) ? c_zero_to - c_off : C;   // channels written (zero beyond C)
  if (C <= 4) {
    const int groups = (fill + 3) >> 2;
    const long long total = voxels * groups;
    launch(pack_cl_f32_skinny_kernel, static_cast<unsigned>((total + 255) / 256), 256, 0, st, x, C, sp, voxels, y, c_pitch,
                                                                                         c_off, groups);
  } else {
    if ((voxels + 31) / 32 >= (1LL << 31)) return cudaErrorInvalidValue;
    launch(pack_cl_f32_tiled_kernel, dim3(static_cast<unsigned>((voxels + 31) / 32), (fill + 31) / 32), 256, 0, st, 
        x, C, sp, voxels, y, c_pitch, c_off, fill);
  }
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_unpack_cl_f32(const float* x, int N, int C, long long sp, int c_pitch, int c_off, float* y,
                             cudaStream_t st) {
  const long long voxels = static_cast<long long>(N) * sp;
  if ((voxels + 31) / 32 >= (1LL << 31)) return cudaErrorInvalidValue;
  launch(unpack_cl_f32_tiled_kernel, dim3(static_cast<unsigned>((voxels + 31) / 32), (C + 31) / 32), 256, 0, st, 
      x, C, sp, voxels, c_pitch, c_off, y);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_bn_stats_cl_f32(const float* y, long long voxels, double* sum, double* sumsq, DetScratch det,
                               cudaStream_t st) {
  launch(bn_stats_cl_f32_kernel, grid_for(voxels, 64, DET_STREAM_BLOCKS), 256, 0, st, y, voxels, sum, sumsq, det);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_bn_apply_cl_f32(const float* y, long long voxels, const float* scale, const float* shift, int act,
                               float* x, cudaStream_t st) {
  launch(bn_apply_cl_f32_kernel, grid_for(voxels * 16, 256), 256, 0, st, y, voxels * 16, scale, shift, act, x);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_bn_train_apply_cl_f32(const float* y, long long voxels, const double* sums, const float* gamma,
                                     const float* beta, float eps, float momentum, float* mm, float* mv, float* saved,
                                     int act, float* x, const float* center, cudaStream_t st) {
  launch(bn_train_apply_cl_f32_kernel, grid_for(voxels * 16, 256), 256, 0, st, 
      y, voxels * 16, sums, static_cast<double>(voxels), gamma, beta, eps, momentum, mm, mv, saved, act, x, center);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_lrelu_bwd_cl_f32(const float* ga, const float* a, long long elems, float* gz, cudaStream_t st) {
  launch(lrelu_bwd_cl_f32_kernel, grid_for(elems / 4, 256), 256, 0, st, ga, a, elems / 4, gz);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_bn_bwd_cl_f32(const float* ga, const float* y, long long voxels, const float* saved, int act,
                             double* sums, DetScratch det, float* gy, float* dgamma, float* dbeta, int accumulate,
                             cudaStream_t st) {
  launch(bn_bwd_reduce_cl_f32_kernel, grid_for(voxels, 64, DET_STREAM_BLOCKS), 256, 0, st, ga, y, voxels, saved, act, sums, det);
  LAUNCH_CHECK();
  launch(bn_bwd_apply_cl_f32_kernel, grid_for(voxels * 16, 256), 256, 0, st, ga, y, voxels * 16, saved, act, sums,
                                                                         1.0 / static_cast<double>(voxels), gy, dgamma,
                                                                         dbeta, accumulate);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_colsum_cl_f32(const float* g, long long voxels, double* scratch, DetScratch det, float* out,
                             int accumulate, cudaStream_t st) {
  cudaError_t e = ew_bn_stats_cl_f32(g, voxels, scratch, scratch + 64, det, st);
  if (e != cudaSuccess) return e;
  launch(d2f_f32_kernel, 1, 64, 0, st, scratch, 64, 1.f, accumulate, out);
  LAUNCH_CHECK();
  return cudaSuccess;
}

}  // namespace hpvg
