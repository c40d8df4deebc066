// 3x3x3 / stride 1 / zero-pad 1 convolution 64 -> (<= 3) channels: the tail convs of every decoder / body block
// (reference networks_3d.py:380,399: Conv3d(N, nc_im)), the discriminator tail (networks_3d.py:186: Conv3d(N, 1)) and
// the data gradient of the 3 -> 64 head convs.
//
// Why a separate kernel: as an ordinary implicit GEMM (conv3d_umma_kernel<CONV_MODE_64_16>) this layer issues 27 taps
// x 4 K-steps of M=256 x N=16 MMAs per output plane, and every one of them re-reads its 128-voxel x 16-channel A tile
// from shared memory: the tensor core is starved by the A-operand read (measured: same time as the 64 -> 64 layer for
// 1/21 of the FLOPs).  Here the nine IN-PLANE taps are moved into the N dimension instead:
//     P[v'][(dh,dw,co)] = sum_dt sum_ci x[t+dt-1][v'][ci] * w[dt][dh][dw][ci][co]      for every voxel v' of the HALOED plane
// is ONE accumulator (N = 27 -> 32 columns, the three temporal taps accumulate in TMEM), so the activations are read
// 12 x per plane instead of 108 x, and the epilogue finishes the convolution with a 9-tap shifted gather
//     out[h][w][co] = sum_{dh,dw} P[(h+dh, w+dw)][(dh,dw,co)]
// through shared memory.  DRAM traffic is unchanged (each input plane is fetched once per tile by TMA, zero fill =
// padding); the kernel is now HBM/L2 bound like the layer should be.
//
// One CTA per SM (cta_group::1), 16h x 8w output tile walking along T.  Haloed plane = 18 x 10 = 180 voxels = 180 rows
// of 128 B (64 bf16 channels, 128B swizzle): rows 0..127 -> MMA M=128, rows 128..191 -> MMA M=64 (rows >= 180 are
// never gathered).  Warp roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..5 / 6..9 epilogue of the
// even / odd output planes (accumulator stage and partial-sum buffer 0 / 1).  The kernel runs at the pace of its
// epilogue (24 small MMAs per plane), and one group alone is a single warp per SM sub-partition walking a dependent
// chain TMEM load -> shared store -> barrier -> 27 shared loads -> global store; the second group overlaps it.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "conv3d_umma.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace hpvg {

namespace {

constexpr int TT_W = 8, TT_H = 16;
constexpr int TB_W = TT_W + 2, TB_H = TT_H + 2;
constexpr int T_ROWS = TB_W * TB_H;               // 180 haloed voxels
constexpr int T_PLANE_BYTES = T_ROWS * 128;       // 23040
constexpr int T_SLOT_STRIDE = 192 * 128;          // 24576: the M=64 MMA reads rows 128..191
constexpr int T_SLOTS = 6;
constexpr int T_W_BYTES = 3 * 32 * 128;           // [dt][n = (dh*3+dw)*3+co (32 rows)][64 ci]  = 12288
constexpr int T_PS = 27;                          // floats per voxel in the partial-sum buffer (odd: conflict-free)
constexpr int T_PBUF_FLOATS = T_ROWS * T_PS;      // 4860
constexpr int T_THREADS = 320;                      // TMA warp, MMA warp, 2 groups of 4 epilogue warps
constexpr int T_TMEM_COLS = 128;                  // 2 buffers x (32 cols M=128 tile + 32 cols M=64 tile)
constexpr int T_SMEM = 1024 + T_W_BYTES + T_SLOTS * T_SLOT_STRIDE + 2 * T_PBUF_FLOATS * 4 + 64 + (2 * T_SLOTS + 5) * 8 + 16;

struct TailParams {
  int N, T, H, W;
  int w_tiles, h_tiles, n_units;
  const uint8_t* wimg;
  const float* scale;
  const float* shift;
  int act;
  float* out;            // fp32 NCDHW, cout_real channels
  int cout_real;
  const float* addend;   // optional fp32 NCDHW residual (added after scale/shift, before the activation)
};

// TF32: the same kernel over fp32 channels-last activations, 32 input channels per launch (a 128-byte row = 32 tf32
// channels; K = 8 per MMA): a 64 -> 3 tail is two launches, the second adding the first's output through `addend`.
template <bool TF32>
__global__ void __launch_bounds__(T_THREADS, 1)
conv3d_tail_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ TailParams p) {
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t base_u32 = smem_u32(smem_dyn);
  uint8_t* sm = smem_dyn + (((base_u32 + 1023u) & ~1023u) - base_u32);
  uint8_t* w_sm = sm;                                        // 12 KB (1024-aligned)
  uint8_t* planes = sm + T_W_BYTES;                          // ring of haloed planes
  float* pbuf = reinterpret_cast<float*>(planes + T_SLOTS * T_SLOT_STRIDE);   // [2][180][27]
  float* scale_sm = pbuf + 2 * T_PBUF_FLOATS;                // [4]
  float* shift_sm = scale_sm + 4;                            // [4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(shift_sm + 4 + 8);
  uint64_t* a_full = bars;                 // [SLOTS]
  uint64_t* a_empty = a_full + T_SLOTS;    // [SLOTS]
  uint64_t* acc_full = a_empty + T_SLOTS;  // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]  4 arrivals (one per epilogue warp)
  uint64_t* w_full = acc_empty + 2;        // [1]
  uint32_t* tmem_ptr_sm = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.T;

  if (threadIdx.x == 0) {
    for (int i = 0; i < T_SLOTS; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    mbar_init(w_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_sm, T_TMEM_COLS);
  pdl_grid_sync();   // launch.cuh: global memory only from here on
  if (threadIdx.x < 4) {
    scale_sm[threadIdx.x] = threadIdx.x < p.cout_real ? p.scale[threadIdx.x] : 0.f;
    shift_sm[threadIdx.x] = threadIdx.x < p.cout_real ? p.shift[threadIdx.x] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_sm;
  const int per_n = p.w_tiles * p.h_tiles;

  if (warp == 0) {
    // ======================================================================================= TMA producer
    if (elect_one()) {
      tma_prefetch_desc(&tmap_in);
      mbar_expect_tx(w_full, T_W_BYTES);
      bulk_load(w_sm, p.wimg, T_W_BYTES, w_full);
      uint32_t j = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        const int n = u / per_n, rem = u - n * per_n;
        const int h0 = (rem / p.w_tiles) * TT_H, w0 = (rem % p.w_tiles) * TT_W;
        for (int t = 0; t < T; ++t, ++j) {
          const uint32_t slot = j % T_SLOTS, ph = (j / T_SLOTS) & 1u;
          mbar_wait(&a_empty[slot], ph ^ 1u);
          mbar_expect_tx(&a_full[slot], T_PLANE_BYTES);
          tma_load_5d(planes + slot * T_SLOT_STRIDE, &tmap_in, &a_full[slot], 0, w0 - 1, h0 - 1, t, n);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================================================================================= MMA issuer
    // The WHOLE warp walks the loop, so counters, barrier addresses and descriptors are warp-uniform and live in the
    // uniform datapath; one elected lane (always the same one: tcgen05.commit tracks the MMAs of its own thread)
    // issues.  Ring positions rotate (no modulo), descriptors are one add away from two built before the loop.
    {
      mbar_wait(w_full, 0);
      const uint32_t idesc128 = make_idesc<TF32>(128, 32);
      const uint32_t idesc64 = make_idesc<TF32>(64, 32);
      const uint64_t a_desc0 = make_smem_desc(smem_u32(planes), 16, 1024, 2);   // slot 0; one add per slot / MMA
      const uint64_t b_desc0 = make_smem_desc(smem_u32(w_sm), 16, 1024, 2);
      constexpr uint32_t SLOT16 = T_SLOT_STRIDE >> 4;
      uint32_t q = 0;
      const uint32_t my_units = (static_cast<uint32_t>(p.n_units) - blockIdx.x + gridDim.x - 1u) / gridDim.x;
      const uint32_t q_end = my_units * static_cast<uint32_t>(T);
      // ring position of plane pl-1 / pl / pl+1 and the phase parity of plane pl+1's slot
      uint32_t s_prev = 0, s_cur = 0, s_next = 0, ph_next = 0;
      auto advance = [&]() {   // (s_prev, s_cur, s_next) <- (s_cur, s_next, s_next + 1)
        s_prev = s_cur;
        s_cur = s_next;
        if (++s_next == T_SLOTS) {
          s_next = 0;
          ph_next ^= 1u;
        }
      };
      // the tensor pipe idles while this warp waits between two planes (conv3d_umma.cu): the waits for the plane
      // the dt = 2 taps read and for the NEXT plane's accumulator stage sit behind the MMAs of dt = 0, 1
      for (uint32_t u = 0; u < my_units; ++u) {
        for (int pl = 0; pl < T; ++pl, ++q) {
          const uint32_t ab = q & 1u;
          const uint32_t d1 = tmem_base + ab * 64, d2 = d1 + 32;
          auto issue_dt = [&](int dt, uint32_t slot, uint32_t accum) {
            const uint64_t a0 = desc_add_lo(a_desc0, slot * SLOT16);
            const uint64_t b0 = desc_add_lo(b_desc0, desc_lo_delta(dt * (32 * 128)));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t bd = desc_add_lo(b0, desc_lo_delta(k * 32));
              umma_ss<TF32>(d1, desc_add_lo(a0, desc_lo_delta(k * 32)), bd, idesc128, accum);
              umma_ss<TF32>(d2, desc_add_lo(a0, desc_lo_delta(128 * 128 + k * 32)), bd, idesc64, accum);
              accum = 1;
            }
          };
          if (pl == 0) {   // a new unit: its plane 0 is the next ring position
            mbar_wait(&a_full[s_next], ph_next);
            advance();     // s_cur = plane 0, s_next = plane 1
          }
          if (q < 2) mbar_wait(&acc_empty[ab], 1u);
          tc_fence_after();
          if (elect_one()) {
            if (pl >= 1) issue_dt(0, s_prev, 0);
            issue_dt(1, s_cur, pl >= 1 ? 1u : 0u);
          }
          __syncwarp();
          if (pl + 1 < T) mbar_wait(&a_full[s_next], ph_next);
          if (q + 1 < q_end && q + 1 >= 2) mbar_wait(&acc_empty[ab ^ 1u], (((q + 1) >> 1) & 1u) ^ 1u);
          tc_fence_after();
          if (elect_one()) {
            if (pl + 1 < T) issue_dt(2, s_next, 1);
            umma_commit(&acc_full[ab]);
            if (pl >= 1) umma_commit(&a_empty[s_prev]);
            if (pl == T - 1) umma_commit(&a_empty[s_cur]);
          }
          __syncwarp();
          if (pl + 1 < T) advance();
        }
      }
    }
  } else {
    // ======================================================================================= epilogue (2 x 4 warps)
    const int quad = warp & 3;
    const uint32_t grp = static_cast<uint32_t>(warp - 2) >> 2;
    const int row = quad * 32 + lane;          // TMEM lane == P row (M=128 tile) == output voxel of the tile
    const int hh = row >> 3, ww = row & 7;
    // M=64 tile: P row 128 + i sits in lane (i/16)*32 + i%16
    const bool has2 = lane < 16 && (128 + quad * 16 + lane) < T_ROWS;
    const int row2 = 128 + quad * 16 + lane;
    const size_t plane_sz = static_cast<size_t>(p.H) * p.W;
    const size_t chan_sz = plane_sz * T;
    uint32_t q = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int n = u / per_n, rem = u - n * per_n;
      const int h = (rem / p.w_tiles) * TT_H + hh, w = (rem % p.w_tiles) * TT_W + ww;
      const bool inb = (h < p.H) && (w < p.W);
      for (int pl = 0; pl < T; ++pl, ++q) {
        const uint32_t ab = q & 1u;
        if (ab != grp) continue;
        float* pb = pbuf + ab * T_PBUF_FLOATS;
        mbar_wait(&acc_full[ab], (q >> 1) & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + ab * 64;
        uint32_t r1[32], r2[32];
        tmem_ld32(taddr, r1);
        tmem_ld32(taddr + 32, r2);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_relaxed(&acc_empty[ab]);
        {
          float* dst = pb + row * T_PS;
#pragma unroll
          for (int i = 0; i < T_PS; ++i) dst[i] = __uint_as_float(r1[i]);
          if (has2) {
            float* dst2 = pb + row2 * T_PS;
#pragma unroll
            for (int i = 0; i < T_PS; ++i) dst2[i] = __uint_as_float(r2[i]);
          }
        }
        // all 180 partial-sum rows of this plane are in shared memory (one named barrier per group)
        asm volatile("bar.sync %0, 128;" ::"r"(1u + grp) : "memory");
        float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int s = 0; s < 9; ++s) {
          const float* src = pb + ((hh + s / 3) * TB_W + (ww + s % 3)) * T_PS + s * 3;
          acc[0] += src[0];
          acc[1] += src[1];
          acc[2] += src[2];
        }
        if (inb) {
          const size_t sp = (static_cast<size_t>(pl) * p.H + h) * p.W + w;
          float v[3];
          size_t idx[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            idx[c] = (static_cast<size_t>(n) * p.cout_real + c) * chan_sz + sp;
            v[c] = fmaf(acc[c], scale_sm[c], shift_sm[c]);
            if (p.addend && c < p.cout_real) v[c] += p.addend[idx[c]];
          }
          if (p.act == CONV_ACT_TANH) {
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = tanhf(v[c]);
          } else if (p.act == CONV_ACT_LRELU) {
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = v[c] > 0.f ? v[c] : 0.2f * v[c];
          }
#pragma unroll
          for (int c = 0; c < 3; ++c)
            if (c < p.cout_real) p.out[idx[c]] = v[c];
        }
        asm volatile("bar.sync %0, 128;" ::"r"(3u + grp) : "memory");   // the gather is done before the buffer is rewritten
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, T_TMEM_COLS);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tail_get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

}  // namespace

int conv3d_tail_wimg_bytes() { return T_W_BYTES; }

const char* conv3d_tail_launch(const ConvLaunch& L, int sm_count, cudaStream_t stream) {
  EncodeTiledFn enc = tail_get_encode();
  if (!enc) return "cuTensorMapEncodeTiled entry point not available";
  const bool tf32 = (L.mode == CONV_MODE_T32_T);
  const int cbox = tf32 ? 32 : 64, esz = tf32 ? 4 : 2;
  if (L.in_pitch < cbox || ((L.in_pitch * esz) & 15)) return "input pitch must be a multiple of 16 bytes and >= the box";
  if ((reinterpret_cast<uintptr_t>(L.in) & 15) != 0) return "input not 16-byte aligned";
  if (L.out_mode != CONV_OUT_F32_NCDHW || L.cout_real < 1 || L.cout_real > 3) return "tail conv: 1..3 fp32 NCDHW outputs";
  if (L.act == CONV_ACT_LRELU_MASK) return "tail conv: LRELU_MASK is a bf16-output epilogue";
  CUtensorMap tmap;
  cuuint64_t gd[5] = {static_cast<cuuint64_t>(cbox), static_cast<cuuint64_t>(L.W), static_cast<cuuint64_t>(L.H),
                      static_cast<cuuint64_t>(L.T), static_cast<cuuint64_t>(L.N)};
  const cuuint64_t vox = static_cast<cuuint64_t>(L.in_pitch) * esz;
  cuuint64_t gs[4] = {vox, vox * L.W, vox * L.W * L.H, vox * L.W * L.H * L.T};
  cuuint32_t bx[5] = {static_cast<cuuint32_t>(cbox), TB_W, TB_H, 1, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  if (enc(&tmap, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(L.in), gd, gs, bx, es,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return "cuTensorMapEncodeTiled failed";
  TailParams prm;
  prm.N = L.N; prm.T = L.T; prm.H = L.H; prm.W = L.W;
  prm.w_tiles = (L.W + TT_W - 1) / TT_W;
  prm.h_tiles = (L.H + TT_H - 1) / TT_H;
  prm.n_units = L.N * prm.w_tiles * prm.h_tiles;
  prm.wimg = static_cast<const uint8_t*>(L.wimg);
  prm.scale = L.scale; prm.shift = L.shift; prm.act = L.act;
  prm.out = static_cast<float*>(L.out);
  prm.cout_real = L.cout_real;
  prm.addend = L.addend;
  if (prm.n_units < 1) return nullptr;
  static bool configured = false;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(conv3d_tail_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv3d_tail_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T_SMEM);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    configured = true;
  }
  const int grid = prm.n_units < sm_count ? prm.n_units : sm_count;
  if (tf32) launch(conv3d_tail_kernel<true>, grid, T_THREADS, T_SMEM, stream, tmap, prm);
  else launch(conv3d_tail_kernel<false>, grid, T_THREADS, T_SMEM, stream, tmap, prm);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace hpvg
