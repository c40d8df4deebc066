// Weight gradient of the 3x3x3 / stride 1 / pad 1 convolution on tcgen05 kind::tf32 (sm_100a), fp32 channels-last
// operands — the fp32-accurate twin of conv3d_wgrad.cu (same reference call sites: MindSpore autodiff of nn.Conv3d's
// weight, networks_3d.py:48-50 under TrainOneStepCell, train_video.py:110-113; the WGAN-GP double backward,
// losses.py:47-52):
//     dW[co][ci][dt][dh][dw] = sum_v gy[v][co] * x[v + (dt-1, dh-1, dw-1)][ci]
//
// GEMM view per tap: D[ci][co] += X_tap^T[ci][k] * GY[k][co], K = voxels.  Both operands are read straight from the
// channels-last fp32 tensors as MN-major UMMA operands.  MN-major tf32 has ONE legal shared-memory layout, the 128-byte
// swizzle with 32-byte atoms (TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B <-> descriptor layout type 1): a 128-byte row =
// 32 channels of one voxel, a K atom = 4 voxels, K = 8 per MMA.  A 64-channel tensor is therefore two 32-channel
// half tiles:
//   * CTA c owns (temporal tap dt, input-channel half): 9 taps x 32 ci x 64 co.  The three dw taps of one dh are stacked
//     along M through the descriptor's leading-dimension offset (atom i = the x row shifted by i voxels; the 4th atom of
//     the M = 128 MMA is a don't-care whose accumulator rows are never read) -> 3 accumulators (dh), 3 MMAs per K-step.
//   * gy's two 32-channel halves sit LBO apart and form the N = 64 operand.
// No masking anywhere: TMA zero-fills out-of-bounds x (the conv padding), out-of-bounds gy and the missing channels of
// narrow tensors (4-channel block inputs / tail gradients).  Partials [groups][27][ci][co] are reduced in fixed order
// (deterministic) by wgrad_reduce_kernel (conv3d_wgrad.cu).
#include <cuda.h>
#include <cuda_runtime.h>

#include "conv3d_umma.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace hpvg {

// defined in conv3d_wgrad.cu
void wgrad_reduce_launch(const float* partial, int groups, float* dw, int w_cin, int kt, int co_off, int co_n,
                         int ci_off, int ci_n, int accumulate, float scale, cudaStream_t stream);

namespace {

constexpr int WT_NH = 3;                                  // gy rows per tile
constexpr int WT_WS = 64;                                 // gy voxels per row per tile
constexpr int WT_XP = WT_WS + 2;                          // x row pitch (voxels)
constexpr int WT_X_BYTES = (WT_NH + 2) * WT_XP * 128;     // 42240: one 32-channel half of the haloed x tile
constexpr int WT_X_STRIDE = 43008;                        // 1024-aligned (the M = 128 MMA's 4th atom reads 128 B past the tile)
constexpr int WT_GH_BYTES = WT_NH * WT_WS * 128;          // 24576: one 32-channel half of the gy tile
constexpr int WT_STAGE = WT_X_STRIDE + 2 * WT_GH_BYTES;   // 92160
constexpr int WT_STAGES = 2;
constexpr int WT_THREADS = 192;
constexpr int WT_SMEM = 1024 + WT_STAGES * WT_STAGE + 64;
constexpr uint32_t WT_LAYOUT = 1;                         // SWIZZLE_128B_BASE32B

struct WgradTParams {
  int N, T, H, W;
  int h_blocks, w_segs, n_tiles, groups;
  int halves;       // input-channel halves handled (2, or 1 when x is a narrow tensor)
  int gy_halves;    // 32-channel halves of gy loaded (2, or 1 when gy is narrow)
  int ncols;        // MMA N: 64, or 8 for a narrow gy
  float* partial;   // [groups][27][64 ci][64 co]
};

__global__ void __launch_bounds__(WT_THREADS, 1)
conv3d_wgrad_tf32_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_gy,
                         const __grid_constant__ WgradTParams p) {
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t base_u32 = smem_u32(smem_dyn);
  uint8_t* sm = smem_dyn + (((base_u32 + 1023u) & ~1023u) - base_u32);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + WT_STAGES * WT_STAGE);
  uint64_t* full = bars;                // [2]
  uint64_t* empty = bars + WT_STAGES;   // [2]
  uint64_t* done = empty + WT_STAGES;   // [1]
  uint32_t* tmem_ptr_sm = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dt = blockIdx.x % 3;
  const int half = (blockIdx.x / 3) % p.halves;
  const int g = blockIdx.x / (3 * p.halves);

  if (threadIdx.x == 0) {
    for (int i = 0; i < WT_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_sm, 256);
  pdl_grid_sync();   // launch.cuh: global memory only from here on
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_sm;

  const int per_plane = p.h_blocks * p.w_segs;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmap_x);
      tma_prefetch_desc(&tmap_gy);
      uint32_t j = 0;
      for (int tile = g; tile < p.n_tiles; tile += p.groups) {
        const int nt = tile / per_plane, rem = tile - nt * per_plane;
        const int n = nt / p.T, t = nt - n * p.T;
        const int t_in = t + dt - 1;
        if (t_in < 0 || t_in >= p.T) continue;
        const int hb = rem / p.w_segs, ws = rem - hb * p.w_segs;
        const int h0 = hb * WT_NH, w0 = ws * WT_WS;
        const uint32_t s = j % WT_STAGES, ph = (j / WT_STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_expect_tx(&full[s], WT_X_BYTES + p.gy_halves * WT_GH_BYTES);
        uint8_t* st = sm + s * WT_STAGE;
        tma_load_5d(st, &tmap_x, &full[s], half * 32, w0 - 1, h0 - 1, t_in, n);
        for (int gh = 0; gh < p.gy_halves; ++gh)
          tma_load_5d(st + WT_X_STRIDE + gh * WT_GH_BYTES, &tmap_gy, &full[s], gh * 32, w0, h0, t, n);
        ++j;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_tf32(128, p.ncols, 1, 1);
      uint32_t j = 0;
      uint32_t accum = 0;
      for (int tile = g; tile < p.n_tiles; tile += p.groups) {
        const int nt = tile / per_plane;
        const int t = nt % p.T;
        const int t_in = t + dt - 1;
        if (t_in < 0 || t_in >= p.T) continue;
        // K-steps (8 voxels each) that contain at least one real voxel: the rest of a ragged last segment is zero fill
        const int w0 = ((tile - nt * per_plane) % p.w_segs) * WT_WS;
        const int nks = (p.W - w0 >= WT_WS) ? WT_WS / 8 : (p.W - w0 + 7) / 8;
        const uint32_t s = j % WT_STAGES, ph = (j / WT_STAGES) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t xs = smem_u32(sm + s * WT_STAGE);
        const uint32_t gs = xs + WT_X_STRIDE;
        const uint64_t x0 = make_smem_desc(xs, 128, 512, WT_LAYOUT);   // descriptors: one add per MMA (ptx.cuh)
        const uint64_t g0 = make_smem_desc(gs, WT_GH_BYTES, 512, WT_LAYOUT);
#pragma unroll 1
        for (int hh = 0; hh < WT_NH; ++hh) {
          const uint64_t xh = desc_add_lo(x0, static_cast<uint32_t>(hh * WT_XP * 128) >> 4);
          const uint64_t gh = desc_add_lo(g0, static_cast<uint32_t>(hh * WT_WS * 128) >> 4);
#pragma unroll
          for (int ks = 0; ks < WT_WS / 8; ++ks) {
            if (ks >= nks) break;
            // B: gy row hh, voxels [8ks, 8ks+8): N = 64 = two 32-channel atoms WT_GH_BYTES apart, K = two 4-voxel atoms
            const uint64_t bd = desc_add_lo(gh, desc_lo_delta(ks * 8 * 128));
            // A: x rows hh + dh, the same voxels shifted by dw = 0..3 (atom stride = one voxel); accumulator dh
#pragma unroll
            for (int dh = 0; dh < 3; ++dh)
              umma_tf32(tmem_base + dh * 64, desc_add_lo(xh, desc_lo_delta((dh * WT_XP + ks * 8) * 128)), bd, idesc, accum);
            accum = 1;
          }
        }
        umma_commit(&empty[s]);
        ++j;
      }
      umma_commit(done);
      if (accum == 0) *tmem_ptr_sm = 0xFFFFFFFFu;   // no tile at all for this CTA: accumulators were never written
    }
    __syncwarp();
  }
  // ---------------------------------------------------------------- readout: TMEM -> partial[g][tap][ci][co]
  mbar_wait(done, 0);
  tc_fence_after();
  __syncthreads();
  const bool empty_cta = (*tmem_ptr_sm == 0xFFFFFFFFu);
  if (warp >= 2) {
    const int quad = warp & 3;            // TMEM lanes quad*32 .. +31  <->  M atom `quad` = tap dw
    if (quad < 3) {
      const int dw = quad, ci = half * 32 + lane;
      float* pg = p.partial + static_cast<size_t>(g) * 27 * 4096;
#pragma unroll 1
      for (int dh = 0; dh < 3; ++dh) {
        const int tap = dt * 9 + dh * 3 + dw;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + dh * 64;
        uint32_t r0[32], r1[32];
        if (!empty_cta) {
          tmem_ld32(taddr, r0);
          tmem_ld32(taddr + 32, r1);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r0[i] = r1[i] = 0u;
        }
        float4* dst = reinterpret_cast<float4*>(pg + (static_cast<size_t>(tap) * 64 + ci) * 64);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          dst[c] = make_float4(__uint_as_float(r0[4 * c]), __uint_as_float(r0[4 * c + 1]),
                               __uint_as_float(r0[4 * c + 2]), __uint_as_float(r0[4 * c + 3]));
#pragma unroll
        for (int c = 0; c < 8; ++c)
          dst[8 + c] = make_float4(__uint_as_float(r1[4 * c]), __uint_as_float(r1[4 * c + 1]),
                                   __uint_as_float(r1[4 * c + 2]), __uint_as_float(r1[4 * c + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn wt_get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

bool wt_make_map(EncodeTiledFn enc, CUtensorMap* m, const void* base, int pitch, int N, int T, int H, int W, int bw,
                 int bh) {
  // a tensor of >= 64 channels: the 64-channel slice at `base` (the CTA picks a 32-channel half through the channel
  // coordinate); a narrow tensor (pitch < 32): all of its channels, the rest of the 32-wide box is TMA zero fill
  cuuint64_t gd[5] = {(cuuint64_t)(pitch < 64 ? pitch : 64), (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)N};
  const cuuint64_t vox = static_cast<cuuint64_t>(pitch) * 4;
  cuuint64_t gs[4] = {vox, vox * W, vox * W * H, vox * W * H * T};
  cuuint32_t bx[5] = {32, (cuuint32_t)bw, (cuuint32_t)bh, 1, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(base), gd, gs, bx, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

size_t conv3d_wgrad_tf32_workspace_bytes(int sm_count) {
  const int groups = sm_count / 3;
  return static_cast<size_t>(groups) * 27 * 4096 * sizeof(float);
}

// x: fp32 cl (N,T,H,W,x_pitch) [64 channels starting at x], gy: fp32 cl (N,T,H,W,gy_pitch) [64 channels at gy]; a pitch
// below 32 (multiple of 4) means "all channels of a narrow tensor, the rest zero".
const char* conv3d_wgrad_tf32_launch(const void* x, int x_pitch, const void* gy, int gy_pitch, int N, int T, int H,
                                     int W, float* dw, int w_cin, int kt, int co_off, int co_n, int ci_off, int ci_n,
                                     int accumulate, float scale, float* workspace, int sm_count,
                                     cudaStream_t stream) {
  EncodeTiledFn enc = wt_get_encode();
  if (!enc) return "cuTensorMapEncodeTiled entry point not available";
  CUtensorMap mx, mg;
  if (!wt_make_map(enc, &mx, x, x_pitch, N, T, H, W, WT_XP, WT_NH + 2)) return "tensor map (x) failed";
  if (!wt_make_map(enc, &mg, gy, gy_pitch, N, T, H, W, WT_WS, WT_NH)) return "tensor map (gy) failed";
  WgradTParams p;
  p.N = N; p.T = T; p.H = H; p.W = W;
  p.h_blocks = (H + WT_NH - 1) / WT_NH;
  p.w_segs = (W + WT_WS - 1) / WT_WS;
  p.n_tiles = N * T * p.h_blocks * p.w_segs;
  p.halves = (x_pitch < 64 || ci_n <= 32) ? 1 : 2;
  p.gy_halves = (gy_pitch < 64 || co_n <= 32) ? 1 : 2;
  p.ncols = (co_n <= 8) ? 8 : (co_n <= 32 ? 32 : 64);
  int groups = sm_count / (3 * p.halves);
  if (groups > p.n_tiles) groups = p.n_tiles;
  if (groups < 1) groups = 1;
  p.groups = groups;
  p.partial = workspace;
  static bool configured = false;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(conv3d_wgrad_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    configured = true;
  }
  launch(conv3d_wgrad_tf32_kernel, 3 * p.halves * groups, WT_THREADS, WT_SMEM, stream, mx, mg, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cudaGetErrorString(e);
  wgrad_reduce_launch(workspace, groups, dw, w_cin, kt, co_off, co_n, ci_off, ci_n, accumulate, scale, stream);
  e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace hpvg
