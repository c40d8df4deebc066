// Weight gradient of the 3x3x3 / stride 1 / pad 1 convolution on tcgen05 (sm_100a), 64 -> 64 channel blocks:
//     dW[co][ci][dt][dh][dw] = sum_v gy[v][co] * x[v + (dt-1, dh-1, dw-1)][ci]
// (what MindSpore autodiff produces for nn.Conv3d's weight, reference networks_3d.py:48-50 under
// TrainOneStepCell, train_video.py:110-113; the same kernel serves the WGAN-GP double-backward).
//
// GEMM view per tap: D[ci][co] += X_tap^T[ci][k] * GY[k][co] with K = voxels.  Both operands are read straight
// from the channels-last bf16 tensors as MN-major / 128B-swizzled UMMA operands: one voxel = one 128-byte row =
// 64 channels, K runs over 16 consecutive voxels of an image row.  The tap shift is a shifted start address of
// the SAME x tile (rows +dh, voxels +dw); two dh taps are stacked into one M=128 MMA through the descriptor's
// leading-dimension offset (second 64-channel block = the x row below), the third uses an M=64 MMA.
//
// Work split: CTA c handles temporal tap dt = c % 3 (its 9 taps = 6 TMEM accumulators, 384 columns, live for the
// whole kernel) and every (c/3 + i*G)-th tile (4 image rows x 64 voxels of gy, haloed 6 x 66 tile of x).
// No masking is needed anywhere: TMA zero-fills out-of-bounds x (the conv padding) and out-of-bounds gy.
// Partials [G][27][ci][co] are reduced (fixed order -> deterministic) and transposed by a second kernel.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "conv3d_umma.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace hpvg {

namespace {

#ifndef HPVG_WG_NH
#define HPVG_WG_NH 4
#endif
#ifndef HPVG_WG_STAGES
#define HPVG_WG_STAGES 2
#endif
constexpr int WG_NH = HPVG_WG_NH;        // gy rows per tile
constexpr int WG_WS = 64;                // gy voxels per row per tile
constexpr int WG_XP = WG_WS + 2;         // x row pitch (voxels)
constexpr int WG_X_BYTES = (WG_NH + 2) * WG_XP * 128;
constexpr int WG_X_STRIDE = (WG_X_BYTES + 1023) & ~1023;   // 1024-aligned
constexpr int WG_GY_BYTES = WG_NH * WG_WS * 128;
constexpr int WG_STAGE = WG_X_STRIDE + WG_GY_BYTES;
constexpr int WG_STAGES = HPVG_WG_STAGES;
static_assert(1024 + WG_STAGES * WG_STAGE + 64 <= 227 * 1024, "wgrad ring exceeds the shared memory of an SM");
constexpr int WG_THREADS = 192;
constexpr int WG_SMEM = 1024 + WG_STAGES * WG_STAGE + 64;

struct WgradParams {
  int N, T, H, W;
  int h_blocks, w_segs, n_tiles, groups;
  int ncols;        // MMA N: 64, or 16 when gy is a narrow tensor (tail convs: co_n <= 16) — a quarter of the MMA time
  float* partial;   // [groups][27][64 ci][64 co]
};

__global__ void __launch_bounds__(WG_THREADS, 1)
conv3d_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_gy,
                    const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t base_u32 = smem_u32(smem_dyn);
  uint8_t* sm = smem_dyn + (((base_u32 + 1023u) & ~1023u) - base_u32);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + WG_STAGES * WG_STAGE);
  uint64_t* full = bars;               // [2]
  uint64_t* empty = bars + WG_STAGES;  // [2]
  uint64_t* done = empty + WG_STAGES;  // [1]
  uint32_t* tmem_ptr_sm = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dt = blockIdx.x % 3;
  const int g = blockIdx.x / 3;

  if (threadIdx.x == 0) {
    for (int i = 0; i < WG_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_sm, 512);
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
        const int h0 = hb * WG_NH, w0 = ws * WG_WS;
        const uint32_t s = j % WG_STAGES, ph = (j / WG_STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_expect_tx(&full[s], WG_X_BYTES + WG_GY_BYTES);
        uint8_t* st = sm + s * WG_STAGE;
        tma_load_5d(st, &tmap_x, &full[s], 0, w0 - 1, h0 - 1, t_in, n);
        tma_load_5d(st + WG_X_STRIDE, &tmap_gy, &full[s], 0, w0, h0, t, n);
        ++j;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc128 = make_idesc_bf16(128, p.ncols, 1, 1);
      const uint32_t idesc64 = make_idesc_bf16(64, p.ncols, 1, 1);
      uint32_t j = 0;
      uint32_t accum = 0;
      for (int tile = g; tile < p.n_tiles; tile += p.groups) {
        const int nt = tile / per_plane;
        const int t = nt % p.T;
        const int t_in = t + dt - 1;
        if (t_in < 0 || t_in >= p.T) continue;
        // K-steps (16 voxels each) that contain at least one real voxel: the rest of a ragged last segment is TMA
        // zero fill and contributes nothing (W = 257: 4.06 instead of 5 segments' worth of MMAs)
        const int w0 = ((tile - nt * per_plane) % p.w_segs) * WG_WS;
        const int nks = (p.W - w0 >= WG_WS) ? WG_WS / 16 : (p.W - w0 + 15) / 16;
        const uint32_t s = j % WG_STAGES, ph = (j / WG_STAGES) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t xs = smem_u32(sm + s * WG_STAGE);
        const uint32_t gs = xs + WG_X_STRIDE;
        const uint64_t x0 = make_smem_desc(xs, 0, 1024, 2);    // descriptors: one add per MMA (ptx.cuh); LBO per tap pair
        const uint64_t g0 = make_smem_desc(gs, 16, 1024, 2);
#pragma unroll 1
        for (int hh = 0; hh < WG_NH; ++hh) {
          const uint64_t xh = desc_add_lo(x0, static_cast<uint32_t>(hh * WG_XP * 128) >> 4);
          const uint64_t gh = desc_add_lo(g0, static_cast<uint32_t>(hh * WG_WS * 128) >> 4);
#pragma unroll
          for (int ks = 0; ks < WG_WS / 16; ++ks) {
            if (ks >= nks) break;
            const uint64_t bd = desc_add_lo(gh, desc_lo_delta(ks * 16 * 128));
            // nine taps in five MMAs: any two taps whose A tiles differ by a CONSTANT byte offset stack along M through
            // the descriptor's leading-dimension offset (second 64-channel block = first + LBO).
            //   acc 0..2: (dh 0, dw) + (dh 1, dw)   LBO = one x row      acc 3: (dh 2, dw 0) + (dh 2, dw 1)   LBO = one voxel
            //   acc 4   : (dh 2, dw 2) alone (M = 64)
#pragma unroll
            for (int dw = 0; dw < 3; ++dw)
              umma_bf16(tmem_base + dw * 64, desc_add_lo(xh, desc_lo_delta((ks * 16 + dw) * 128, WG_XP * 128)), bd,
                        idesc128, accum);
            constexpr uint32_t ROW2 = 2 * WG_XP * 128;
            umma_bf16(tmem_base + 3 * 64, desc_add_lo(xh, desc_lo_delta(ROW2 + ks * 16 * 128, 128)), bd, idesc128, accum);
            umma_bf16(tmem_base + 4 * 64, desc_add_lo(xh, desc_lo_delta(ROW2 + (ks * 16 + 2) * 128, WG_XP * 128)), bd,
                      idesc64, accum);
            accum = 1;
          }
        }
        umma_commit(&empty[s]);
        ++j;
      }
      umma_commit(done);
      // no tile at all for this CTA (tiny inputs): accumulators were never written -> flag through `accum`
      if (accum == 0) *tmem_ptr_sm = 0xFFFFFFFFu;
    }
    __syncwarp();
  }
  // ---------------------------------------------------------------- readout: TMEM -> partial[g][tap][ci][co]
  mbar_wait(done, 0);
  tc_fence_after();
  __syncthreads();
  const bool empty_cta = (*tmem_ptr_sm == 0xFFFFFFFFu);
  if (warp >= 2) {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;   // TMEM lane
    float* pg = p.partial + static_cast<size_t>(g) * 27 * 4096;
#pragma unroll 1
    for (int a = 0; a < 5; ++a) {
      // accumulator a -> (dh, dw) of this TMEM lane (see the MMA issue loop)
      int dh, dw, ci;
      bool valid = true;
      if (a < 3) {
        dh = row >> 6;
        dw = a;
        ci = row & 63;
      } else if (a == 3) {
        dh = 2;
        dw = row >> 6;
        ci = row & 63;
      } else {   // M = 64 tile: row i sits in lane (i/16)*32 + i%16
        dh = 2;
        dw = 2;
        valid = (row & 31) < 16;
        ci = (row >> 5) * 16 + (row & 15);
      }
      const int tap = dt * 9 + dh * 3 + dw;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + a * 64;
      uint32_t r0[32], r1[32];
      if (!empty_cta) {
        tmem_ld32(taddr, r0);
        tmem_ld32(taddr + 32, r1);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r0[i] = r1[i] = 0u;
      }
      if (valid) {
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
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// dW[(co+co_off)][(ci+ci_off)][tap] (+)= sum_g partial[g][tap][ci][co]    (tap stride = 1; layout (Cout, Cin, kt,3,3))
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int groups, float* __restrict__ dw, int w_cin,
                                    int kt, int co_off, int co_n, int ci_off, int ci_n, int accumulate, float scale) {
  pdl_grid_sync();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // over 27*64*64, co fastest (coalesced partial reads)
  if (idx >= 27 * 4096) return;
  const int co = idx & 63, ci = (idx >> 6) & 63, tap = idx >> 12;
  if (co >= co_n || ci >= ci_n) return;   // zero-padded channels of skinny layers
  float acc = 0.f;
  for (int gg = 0; gg < groups; ++gg) acc += partial[static_cast<size_t>(gg) * 27 * 4096 + idx];
  int tap_o = tap;
  if (kt == 1) {
    if (tap / 9 != 1) return;   // 2-D filter: only the centre temporal tap exists
    tap_o = tap - 9;
  }
  const size_t o = (static_cast<size_t>(co + co_off) * w_cin + (ci + ci_off)) * (9 * kt) + tap_o;
  dw[o] = (accumulate ? dw[o] : 0.f) + acc * scale;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn wg_get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

bool make_map(EncodeTiledFn enc, CUtensorMap* m, const void* base, int pitch, int N, int T, int H, int W, int bw,
              int bh) {
  // a narrow tensor (pitch < 64: the 8-channel block inputs / tail gradients) is mapped with its real channel count:
  // the box still spans 64 channels and TMA zero-fills the out-of-bounds ones, so no zero-padded 64-channel copy of a
  // 3-channel tensor ever exists in HBM (it used to cost a pack kernel and 8x the operand traffic)
  cuuint64_t gd[5] = {(cuuint64_t)(pitch < 64 ? pitch : 64), (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)N};
  const cuuint64_t vox = static_cast<cuuint64_t>(pitch) * 2;
  cuuint64_t gs[4] = {vox, vox * W, vox * W * H, vox * W * H * T};
  cuuint32_t bx[5] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gd, gs, bx, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

void wgrad_reduce_launch(const float* partial, int groups, float* dw, int w_cin, int kt, int co_off, int co_n,
                         int ci_off, int ci_n, int accumulate, float scale, cudaStream_t stream) {
  launch(wgrad_reduce_kernel, (27 * 4096 + 255) / 256, 256, 0, stream, partial, groups, dw, w_cin, kt, co_off, co_n, ci_off,
                                                                  ci_n, accumulate, scale);
}

size_t conv3d_wgrad_workspace_bytes(int sm_count) {
  const int groups = sm_count / 3;
  return static_cast<size_t>(groups) * 27 * 4096 * sizeof(float);
}

// x: bf16 cl (N,T,H,W,x_pitch) [64 channels starting at x], gy: bf16 cl (N,T,H,W,gy_pitch) [64 channels at gy];
// a pitch below 64 (multiple of 8) means "all channels of a narrow tensor, the rest zero"
// dw: fp32 (w_cout, w_cin, kt, 3, 3); the 64x64 block at (co_off, ci_off) is written (or accumulated into).
const char* conv3d_wgrad_launch(const void* x, int x_pitch, const void* gy, int gy_pitch, int N, int T, int H, int W,
                                float* dw, int w_cin, int kt, int co_off, int co_n, int ci_off, int ci_n, int accumulate,
                                float scale, float* workspace, int sm_count, cudaStream_t stream) {
  EncodeTiledFn enc = wg_get_encode();
  if (!enc) return "cuTensorMapEncodeTiled entry point not available";
  CUtensorMap mx, mg;
  if (!make_map(enc, &mx, x, x_pitch, N, T, H, W, WG_XP, WG_NH + 2)) return "tensor map (x) failed";
  if (!make_map(enc, &mg, gy, gy_pitch, N, T, H, W, WG_WS, WG_NH)) return "tensor map (gy) failed";
  WgradParams p;
  p.N = N; p.T = T; p.H = H; p.W = W;
  p.h_blocks = (H + WG_NH - 1) / WG_NH;
  p.w_segs = (W + WG_WS - 1) / WG_WS;
  p.n_tiles = N * T * p.h_blocks * p.w_segs;
  int groups = sm_count / 3;
  if (groups > p.n_tiles) groups = p.n_tiles;
  if (groups < 1) groups = 1;
  p.groups = groups;
  p.ncols = (co_n <= 16) ? 16 : 64;      // accumulator columns >= ncols are never written and never read back
  p.partial = workspace;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv3d_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    configured = true;
  }
  launch(conv3d_wgrad_kernel, 3 * groups, WG_THREADS, WG_SMEM, stream, mx, mg, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cudaGetErrorString(e);
  launch(wgrad_reduce_kernel, (27 * 4096 + 255) / 256, 256, 0, stream, workspace, groups, dw, w_cin, kt, co_off, co_n,
                                                                  ci_off, ci_n, accumulate, scale);
  e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace hpvg
