// 3x3x3 / stride 1 / zero-pad 1 convolution as an implicit GEMM on the 5th-gen tensor cores (tcgen05, sm_100a).
//
// Replaces the library call behind `nn.Conv3d` in reference ConvBlock3D (src/modules/networks_3d.py:48-50), the
// decoder/body tail convs (networks_3d.py:380,399), `SpectualNormConv3d.conv3d` (src/tools/spectral_norm.py:152)
// and, with repacked weights, their data-gradient.  Conv2d (networks_2d.py) is the T=1 case.
//
// Design:
//   * activations live in HBM channels-last (N,T,H,W,C) bf16; one TMA box = one haloed (18h x 10w x C) input plane;
//     out-of-bounds box elements are zero-filled by TMA, which IS the conv zero padding.
//   * a CTA PAIR (cluster of 2, tcgen05 cta_group::2) owns two vertically adjacent 8w x 16h output tiles and walks
//     along T.  M = 256 rows (128 voxels per CTA), N = Cout, K = 27 taps x Cin.  The A operand of tap (dt,dh,dw) is
//     the SAME shared-memory plane addressed through a shifted matrix descriptor (start += (dh*10+dw) rows) — no
//     im2col copy exists anywhere.  Input planes sit in a ring: plane t is loaded once per strip and used by the
//     three output planes t-1, t, t+1.
//   * the whole filter bank (27 taps) stays resident in shared memory, split across the pair (each CTA holds half
//     of Cout), so the steady state moves only activations: ~1.4 x 128 B per voxel from L2.
//   * FP32 accumulators in TMEM, double buffered: the epilogue of plane p (TMEM -> registers -> per-channel
//     scale/shift (+bias, folded BN, 1/sigma of spectral norm) -> LeakyReLU/tanh -> bf16/f32 store) overlaps the
//     MMAs of plane p+1.
//   * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA only) + TMEM owner, warps 2..5 = epilogue
//     (the head variant has a second epilogue group, warps 6..9, one group per accumulator stage: Cfg::EPI_GROUPS).
//   * the MMA-issuing thread is on the critical path (the tensor pipe runs only a few MMAs behind it): descriptors are
//     one add away from two built per plane, and the mbarrier waits sit behind already queued MMAs (see the issuer).
//   * every launch is a programmatic dependent launch (launch.cuh): barrier init and TMEM allocation overlap the tail
//     of the previous kernel of the stream; global memory is touched only after pdl_grid_sync().
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "conv3d_umma.h"
#include "det_reduce.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace hpvg {

namespace {

constexpr int TILE_W = 8;
constexpr int TILE_H = 16;
constexpr int BOX_W = TILE_W + 2;
constexpr int BOX_H = TILE_H + 2;

template <int MODE>
struct Cfg;
// MODE 0: Cin 64 -> Cout 64 (SW128 planes, 4 K-steps per tap)
template <>
struct Cfg<CONV_MODE_64_64> {
  static constexpr bool TF32 = false, HEAD = false;
  static constexpr int CIN = 64, NOUT = 64, ROWS_PER_CTA = 32;
  static constexpr int PLANE_BYTES = BOX_H * BOX_W * 128, SLOT_STRIDE = 23552, SLOTS = 4;
  static constexpr int W_BYTES = 27 * ROWS_PER_CTA * 128, DT_BYTES = 9 * ROWS_PER_CTA * 128;
  static constexpr int ACC_STRIDE = 64, TMEM_COLS = 128;
  static constexpr int EPI_GROUPS = 1;   // groups of 4 epilogue warps (the epilogue hides under the 108 MMAs of a plane)
  static constexpr int STAGES = 1;   // 16 KB output staging tiles for the TMA-store epilogue (what the 227 KB leave)
};
// MODE 1: Cin 64 -> Cout <= 16 (tail convs 64->3 / 64->1)
template <>
struct Cfg<CONV_MODE_64_16> {
  static constexpr bool TF32 = false, HEAD = false;
  static constexpr int CIN = 64, NOUT = 16, ROWS_PER_CTA = 8;
  static constexpr int PLANE_BYTES = BOX_H * BOX_W * 128, SLOT_STRIDE = 23552, SLOTS = 6;
  static constexpr int W_BYTES = 27 * ROWS_PER_CTA * 128, DT_BYTES = 9 * ROWS_PER_CTA * 128;
  static constexpr int ACC_STRIDE = 32, TMEM_COLS = 64;
  static constexpr int EPI_GROUPS = 1;
  static constexpr int STAGES = 0;
};
// MODE 2: Cin <= 8 -> Cout 64 (head convs 3->64); no-swizzle planes of 16 B per voxel, taps paired through LBO
template <>
struct Cfg<CONV_MODE_8_64> {
  static constexpr bool TF32 = false, HEAD = true;
  static constexpr int CIN = 8, NOUT = 64, ROWS_PER_CTA = 32;
  static constexpr int PLANE_BYTES = BOX_H * BOX_W * 16, SLOT_STRIDE = 3072, SLOTS = 8;
  static constexpr int W_BYTES = 3 * 5 * 1024, DT_BYTES = 5 * 1024;
  static constexpr int ACC_STRIDE = 64, TMEM_COLS = 128;
  // The head conv issues 15 MMAs per output plane against 108 for the 64 -> 64 layer, so it runs at the pace of its
  // epilogue: one warp per SM sub-partition walking ld -> 64 x (affine, activation, pack) -> 8 stores per plane at an
  // issue rate of 0.22 (profiles/r2_ncu_head_conv_store_experiment.csv).  Two groups of 4 epilogue warps, one per
  // accumulator stage, put two warps on every sub-partition and overlap one plane's TMEM load with the other's stores.
  static constexpr int EPI_GROUPS = 2;
  // One staging tile per epilogue group.  (With ONE group the staged TMA store was 14 % slower than direct stores,
  // 0.359 vs 0.316 ms at 8 x 13x192x257, although the L2 write requests dropped 8x: the lone group waited on its own
  // store's read-out.  With direct stores the kernel sits at 79 % L1TEX throughput — 32 partial-line streams per store.)
  static constexpr int STAGES = 2;
};
// kind::tf32 twins: fp32 operands, the SAME byte layouts with half the channels per row (a 128-byte row = 32 tf32
// channels, K = 8 per MMA = the same 32 bytes; a 16-byte head voxel = 4 fp32 channels).  A 64 -> 64 layer at tf32 is two
// launches of the 32 -> 64 variant (second launch adds the first's raw fp32 partial sums): exactly the half-rate of the
// tf32 tensor pipe, and the full 27-tap filter bank of 32 input channels still fits the shared memory next to the ring.
template <>
struct Cfg<CONV_MODE_T32_64> : Cfg<CONV_MODE_64_64> {
  static constexpr bool TF32 = true;
  static constexpr int CIN = 32;
};
template <>
struct Cfg<CONV_MODE_T4_64> : Cfg<CONV_MODE_8_64> {
  static constexpr bool TF32 = true;
  static constexpr int CIN = 4;
};

template <int MODE>
constexpr int num_threads() { return 64 + 128 * Cfg<MODE>::EPI_GROUPS; }   // TMA warp, MMA warp, epilogue warps

constexpr int STAGE_BYTES = TILE_H * TILE_W * 128;   // one CTA's output tile: 128 voxels x 64 bf16 channels

struct Unit {
  int n, h0, w0;
};
__device__ __forceinline__ Unit decode_unit(int u, const ConvParams& p, uint32_t rank) {
  Unit r;
  const int per_n = p.w_tiles * p.h_pairs;
  r.n = u / per_n;
  const int rem = u - r.n * per_n;
  const int hp = rem / p.w_tiles;
  const int wt = rem - hp * p.w_tiles;
  r.h0 = (2 * hp + static_cast<int>(rank)) * TILE_H;
  r.w0 = wt * TILE_W;
  return r;
}

// Work partition: the (unit, output plane) sequence is cut into n_pairs contiguous ranges of equal length (+-1), so a
// CTA pair owns a partial strip at either end of its range and whole strips in between.  Splitting strips along T
// costs at most two re-loaded halo planes per pair and removes the wave quantisation of whole-strip scheduling
// (198 strips on 74 pairs: 3 waves for 2.68 waves of work at the finest scale, batch 1).
struct Item {
  int u, t0, t1;   // strip, output planes [t0, t1)
};
__device__ __forceinline__ void work_range(int pair, int n_pairs, const ConvParams& p, int& g0, int& g1) {
  const long long P = static_cast<long long>(p.n_units) * p.T;
  g0 = static_cast<int>(P * pair / n_pairs);
  g1 = static_cast<int>(P * (pair + 1) / n_pairs);
}
__device__ __forceinline__ Item next_item(int g, int g1, int T) {
  Item it;
  it.u = g / T;
  it.t0 = g - it.u * T;
  it.t1 = min(T, it.t0 + (g1 - g));
  return it;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == CONV_ACT_LRELU) return v > 0.f ? v : 0.2f * v;
  if (act == CONV_ACT_TANH) return tanhf(v);
  return v;
}

// One accumulator row (64 fp32 channels in r0|r1) -> affine -> activation -> 128 bytes of bf16.  The activation and
// the presence of an addend are compile-time so the 64-element body is branch-free straight-line code (the runtime
// switch sits outside, once per plane).  KEEP: write the bf16-rounded results back into r0/r1 as fp32 bit patterns
// (input of the fused BatchNorm statistics).
template <int ACT, bool ADD, bool KEEP>
__device__ __forceinline__ void epilogue_bf16_row(uint32_t (&r0)[32], uint32_t (&r1)[32], const float* scale_sm,
                                                  const float* shift_sm, const float* __restrict__ add,
                                                  __nv_bfloat16* __restrict__ dst,
                                                  const __nv_bfloat16* __restrict__ mask = nullptr, int swz = 0) {
  // swz: 0 for a global row; (row & 7) for a row of the 128B-swizzled shared-memory staging tile of the TMA store
  const float4* sc4 = reinterpret_cast<const float4*>(scale_sm);
  const float4* sh4 = reinterpret_cast<const float4*>(shift_sm);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t* r = half ? r1 : r0;
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
      const int cb = half * 32 + c8 * 8;
      const float4 sa = sc4[cb / 4], sb = sc4[cb / 4 + 1], ha = sh4[cb / 4], hb = sh4[cb / 4 + 1];
      const float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
      const float sh[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
      float ad[8];
      if constexpr (ADD) {
        const float4 a0 = *reinterpret_cast<const float4*>(add + cb), a1 = *reinterpret_cast<const float4*>(add + cb + 4);
        ad[0] = a0.x; ad[1] = a0.y; ad[2] = a0.z; ad[3] = a0.w;
        ad[4] = a1.x; ad[5] = a1.y; ad[6] = a1.z; ad[7] = a1.w;
      }
      float mk[8];
      if constexpr (ACT == CONV_ACT_LRELU_MASK) {   // LeakyReLU'(stored activation): sign(activation) == sign(pre-act)
        const uint4 m4 = *reinterpret_cast<const uint4*>(mask + cb);
        const uint32_t mw[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          mk[2 * e2] = __uint_as_float(mw[e2] << 16);
          mk[2 * e2 + 1] = __uint_as_float(mw[e2] & 0xFFFF0000u);
        }
      }
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float a = __uint_as_float(r[c8 * 8 + e]);
        if constexpr (ADD) a += ad[e];
        a = fmaf(a, sc[e], sh[e]);
        if constexpr (ACT == CONV_ACT_LRELU) a = fmaxf(a, 0.2f * a);   // == a > 0 ? a : 0.2a
        if constexpr (ACT == CONV_ACT_TANH) a = tanhf(a);
        if constexpr (ACT == CONV_ACT_LRELU_MASK) a = mk[e] > 0.f ? a : 0.2f * a;
        v[e] = a;
      }
      uint4 pk;
      pk.x = pack_bf16x2(v[0], v[1]);
      pk.y = pack_bf16x2(v[2], v[3]);
      pk.z = pack_bf16x2(v[4], v[5]);
      pk.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(dst + ((((cb >> 3) ^ swz)) << 3)) = pk;
      if constexpr (KEEP) {
        const uint32_t pw[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          r[c8 * 8 + 2 * e2] = pw[e2] << 16;
          r[c8 * 8 + 2 * e2 + 1] = pw[e2] & 0xFFFF0000u;
        }
      }
    }
  }
}

template <bool KEEP>
__device__ __forceinline__ void epilogue_bf16_dispatch(int act, uint32_t (&r0)[32], uint32_t (&r1)[32],
                                                       const float* scale_sm, const float* shift_sm,
                                                       const float* __restrict__ add, __nv_bfloat16* __restrict__ dst,
                                                       const __nv_bfloat16* __restrict__ mask, int swz = 0) {
  if (act == CONV_ACT_LRELU_MASK) {
    if (add) epilogue_bf16_row<CONV_ACT_LRELU_MASK, true, KEEP>(r0, r1, scale_sm, shift_sm, add, dst, mask, swz);
    else epilogue_bf16_row<CONV_ACT_LRELU_MASK, false, KEEP>(r0, r1, scale_sm, shift_sm, add, dst, mask, swz);
    return;
  }
  if (add) {   // split-Cin accumulation (128 -> 64 layers): rare
    if (act == CONV_ACT_LRELU) epilogue_bf16_row<CONV_ACT_LRELU, true, KEEP>(r0, r1, scale_sm, shift_sm, add, dst, nullptr, swz);
    else if (act == CONV_ACT_TANH) epilogue_bf16_row<CONV_ACT_TANH, true, KEEP>(r0, r1, scale_sm, shift_sm, add, dst, nullptr, swz);
    else epilogue_bf16_row<CONV_ACT_NONE, true, KEEP>(r0, r1, scale_sm, shift_sm, add, dst, nullptr, swz);
  } else {
    if (act == CONV_ACT_LRELU) epilogue_bf16_row<CONV_ACT_LRELU, false, KEEP>(r0, r1, scale_sm, shift_sm, add, dst, nullptr, swz);
    else if (act == CONV_ACT_TANH) epilogue_bf16_row<CONV_ACT_TANH, false, KEEP>(r0, r1, scale_sm, shift_sm, add, dst, nullptr, swz);
    else epilogue_bf16_row<CONV_ACT_NONE, false, KEEP>(r0, r1, scale_sm, shift_sm, add, dst, nullptr, swz);
  }
}

// fp32 channels-last twin (tf32 modes): 64 fp32 channels = 256 bytes per voxel, values rounded to tf32 (round to nearest)
// because the consumer's tensor core would otherwise truncate them (except the statistics variant, see below).
template <int ACT, bool ADD, bool KEEP>
__device__ __forceinline__ void epilogue_f32_row(uint32_t (&r0)[32], uint32_t (&r1)[32], const float* scale_sm,
                                                 const float* shift_sm, const float* __restrict__ add,
                                                 float* __restrict__ dst) {
  const float4* sc4 = reinterpret_cast<const float4*>(scale_sm);
  const float4* sh4 = reinterpret_cast<const float4*>(shift_sm);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t* r = half ? r1 : r0;
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      const int cb = half * 32 + c4 * 4;
      const float4 sa = sc4[cb / 4], ha = sh4[cb / 4];
      const float sc[4] = {sa.x, sa.y, sa.z, sa.w};
      const float sh[4] = {ha.x, ha.y, ha.z, ha.w};
      float ad[4] = {0.f, 0.f, 0.f, 0.f};
      if constexpr (ADD) {
        const float4 a0 = *reinterpret_cast<const float4*>(add + cb);
        ad[0] = a0.x; ad[1] = a0.y; ad[2] = a0.z; ad[3] = a0.w;
      }
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float a = __uint_as_float(r[c4 * 4 + e]);
        if constexpr (ADD) a += ad[e];
        a = fmaf(a, sc[e], sh[e]);
        if constexpr (ACT == CONV_ACT_LRELU) a = fmaxf(a, 0.2f * a);
        if constexpr (ACT == CONV_ACT_TANH) a = tanhf(a);
        // KEEP = the pre-BatchNorm output of a training-mode layer: read by the normalise / backward passes only, never by
        // a tensor core, and its SIGN after the affine decides the LeakyReLU mask — stored unrounded
        v[e] = KEEP ? a : round_tf32(a);
      }
      *reinterpret_cast<float4*>(dst + cb) = make_float4(v[0], v[1], v[2], v[3]);
      if constexpr (KEEP) {
#pragma unroll
        for (int e = 0; e < 4; ++e) r[c4 * 4 + e] = __float_as_uint(v[e]);
      }
    }
  }
}

template <bool KEEP>
__device__ __forceinline__ void epilogue_f32_dispatch(int act, uint32_t (&r0)[32], uint32_t (&r1)[32],
                                                      const float* scale_sm, const float* shift_sm,
                                                      const float* __restrict__ add, float* __restrict__ dst) {
  if (add) {
    if (act == CONV_ACT_LRELU) epilogue_f32_row<CONV_ACT_LRELU, true, KEEP>(r0, r1, scale_sm, shift_sm, add, dst);
    else if (act == CONV_ACT_TANH) epilogue_f32_row<CONV_ACT_TANH, true, KEEP>(r0, r1, scale_sm, shift_sm, add, dst);
    else epilogue_f32_row<CONV_ACT_NONE, true, KEEP>(r0, r1, scale_sm, shift_sm, add, dst);
  } else {
    if (act == CONV_ACT_LRELU) epilogue_f32_row<CONV_ACT_LRELU, false, KEEP>(r0, r1, scale_sm, shift_sm, add, dst);
    else if (act == CONV_ACT_TANH) epilogue_f32_row<CONV_ACT_TANH, false, KEEP>(r0, r1, scale_sm, shift_sm, add, dst);
    else epilogue_f32_row<CONV_ACT_NONE, false, KEEP>(r0, r1, scale_sm, shift_sm, add, dst);
  }
}

template <int MODE, bool STATS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(num_threads<MODE>(), 1)
conv3d_umma_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_out,
                   const __grid_constant__ ConvParams p) {
  using C = Cfg<MODE>;
  extern __shared__ uint8_t smem_dyn[];
  // carve (identical in both CTAs of the pair): [weights | plane ring | scale | shift | barriers | tmem ptr]
  const uint32_t base_u32 = smem_u32(smem_dyn);
  uint8_t* sm = smem_dyn + (((base_u32 + 1023u) & ~1023u) - base_u32);
  uint8_t* w_sm = sm;
  uint8_t* planes = sm + ((C::W_BYTES + 1023) & ~1023);
  uint8_t* stage = planes + C::SLOTS * C::SLOT_STRIDE;     // [STAGES][128 rows][128 B], 1024-byte aligned, 128B-swizzled
  float* scale_sm = reinterpret_cast<float*>(stage + C::STAGES * STAGE_BYTES);
  float* shift_sm = scale_sm + 64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(shift_sm + 64);
  uint64_t* a_full = bars;                    // [SLOTS]  (used in the leader CTA; tx from both CTAs)
  uint64_t* a_empty = a_full + C::SLOTS;      // [SLOTS]  (each CTA its own; multicast commit)
  uint64_t* acc_full = a_empty + C::SLOTS;    // [2]      (each CTA its own; multicast commit)
  uint64_t* acc_empty = acc_full + 2;         // [2]      (leader's is used; 8 arrivals = 4 warps x 2 CTAs)
  uint64_t* w_full = acc_empty + 2;           // [3]      this CTA's filter taps of temporal offset dt have landed
  uint64_t* w_peer = w_full + 3;              // [3]      (leader's is used) the peer's taps of dt have landed
  uint32_t* tmem_ptr_sm = reinterpret_cast<uint32_t*>(w_peer + 3);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int n_pairs = gridDim.x >> 1;
  const int T = p.T;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::SLOTS; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 8);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_peer[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_ptr_sm, C::TMEM_COLS);
  // barriers and TMEM are set up while the previous kernel of the stream drains (launch.cuh); global memory from here on
  pdl_grid_sync();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_sm;
  // Order in which the temporal thirds of the filter bank are fetched = the order this pair's FIRST output plane consumes
  // them: dt = 1, 2, 0 when it is the first plane of a strip (no dt = 0 tap), else dt = 0, 1, 2.  (Measured with
  // clock64 stamps at 4x38x51, one plane per pair: the three pairs out of four that start mid-strip waited for the
  // whole 110 KB bank before their first MMA.)
  int w_first = 1;
  {
    int g0w, g1w;
    work_range(pair, n_pairs, p, g0w, g1w);
    if (g1w > g0w && next_item(g0w, g1w, T).t0 > 0) w_first = 0;
  }

  if (warp == 0) {
    // =========================================================================== TMA producer (both CTAs)
    if (elect_one()) {
      tma_prefetch_desc(&tmap_in);
      // resident filter bank: this CTA's half of Cout, all taps — one barrier per temporal offset, so the first MMAs
      // start after a third of the bank.  The first third goes out now, the other two behind the first two input planes
      // (what the first MMAs need is one third of the bank and two planes, not the whole bank).
      const uint8_t* wsrc = p.wimg + static_cast<size_t>(rank) * C::W_BYTES;
      constexpr int CHUNK = (C::DT_BYTES % 3 == 0 && C::DT_BYTES / 3 >= 1024) ? C::DT_BYTES / 3 : C::DT_BYTES;
      static_assert(C::W_BYTES == 3 * C::DT_BYTES && C::DT_BYTES % CHUNK == 0 && CHUNK % 16 == 0, "filter bank layout");
      auto load_bank_third = [&](int i) {
        const int dt = (w_first + i) % 3;
        mbar_expect_tx(&w_full[dt], C::DT_BYTES);
        for (int off = dt * C::DT_BYTES; off < (dt + 1) * C::DT_BYTES; off += CHUNK)
          bulk_load(w_sm + off, wsrc + off, CHUNK, &w_full[dt]);
      };
      load_bank_third(0);
      bool bank_complete = false;
      uint32_t leader_full[C::SLOTS];
#pragma unroll
      for (int i = 0; i < C::SLOTS; ++i) leader_full[i] = map_to_cta(smem_u32(&a_full[i]), 0);
      uint32_t j = 0;
      int g0, g1;
      work_range(pair, n_pairs, p, g0, g1);
      for (int g = g0; g < g1;) {
        const Item it = next_item(g, g1, T);
        g += it.t1 - it.t0;
        const Unit un = decode_unit(it.u, p, rank);
        const int tlo = max(it.t0 - 1, 0), thi = min(it.t1, T - 1);   // input planes incl. the temporal halo
        for (int t = tlo; t <= thi; ++t, ++j) {
          const uint32_t slot = j % C::SLOTS;
          const uint32_t ph = (j / C::SLOTS) & 1u;
          mbar_wait(&a_empty[slot], ph ^ 1u);
          if (rank == 0) mbar_expect_tx(&a_full[slot], 2u * C::PLANE_BYTES);
          uint32_t dst_bar = leader_full[0];
#pragma unroll
          for (int i = 1; i < C::SLOTS; ++i) dst_bar = (slot == static_cast<uint32_t>(i)) ? leader_full[i] : dst_bar;
          if (C::HEAD && p.in_merged)   // (C, W) merged into one 160-byte box row
            tma_load_5d_pair(planes + slot * C::SLOT_STRIDE, &tmap_in, dst_bar, (un.w0 - 1) * C::CIN, un.h0 - 1, t, un.n, 0);
          else
            tma_load_5d_pair(planes + slot * C::SLOT_STRIDE, &tmap_in, dst_bar, 0, un.w0 - 1, un.h0 - 1, t, un.n);
          if (!bank_complete && j >= 1) {
            load_bank_third(1);
            load_bank_third(2);
            bank_complete = true;
          }
        }
      }
      if (!bank_complete) {   // (a pair with at most one input plane; every third is waited for before the kernel ends)
        load_bank_third(1);
        load_bank_third(2);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================================================================== MMA issuer (leader CTA only)
    if (rank == 1 && elect_one()) {
      // tell the leader as each third of this CTA's half of the filter bank becomes resident
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int dt = (w_first + i) % 3;   // the order of arrival (see w_first)
        mbar_wait(&w_full[dt], 0);
        mbar_arrive_cluster(map_to_cta(smem_u32(&w_peer[dt]), 0));
      }
    }
    if (rank == 0 && elect_one()) {
      uint32_t w_ready = 0;                      // bit dt: the taps of temporal offset dt are resident in both CTAs
      const uint32_t idesc = make_idesc<C::TF32>(256, C::NOUT);
      const uint32_t w_addr = smem_u32(w_sm);
      const uint32_t planes_addr = smem_u32(planes);
      uint32_t j0 = 0, q = 0;
      int g0, g1;
      work_range(pair, n_pairs, p, g0, g1);
      const uint32_t q_end = static_cast<uint32_t>(g1 - g0);   // output planes of this pair
      // The tensor pipe runs only a few MMAs behind this thread, so whatever the thread does BETWEEN the last MMA of
      // one plane and the first MMA of the next is idle time of the pipe (measured: ~1100 cycles per plane when every
      // wait sat there — a third of the 64 -> 64 layer's time and two thirds of the head conv's).  Hence the order:
      //   taps of dt = 0, 1 (planes pl-1, pl: waited for during the previous plane)
      //   -> wait for plane pl+1 AND for the next plane's accumulator stage, behind the MMAs just queued
      //   -> taps of dt = 2 -> commits.
      for (int g = g0; g < g1;) {
        const Item it = next_item(g, g1, T);
        g += it.t1 - it.t0;
        const int tlo = max(it.t0 - 1, 0), thi = min(it.t1, T - 1);
        int arrived = tlo;                       // input planes [tlo, arrived) of this item have been waited for
        auto plane_slot = [&](int t) { return (j0 + static_cast<uint32_t>(t - tlo)) % C::SLOTS; };
        auto wait_planes = [&](int need) {
          for (; arrived <= need; ++arrived) {
            const uint32_t jj = j0 + static_cast<uint32_t>(arrived - tlo);
            mbar_wait(&a_full[jj % C::SLOTS], (jj / C::SLOTS) & 1u);
          }
        };
        for (int pl = it.t0; pl < it.t1; ++pl, ++q) {
          const uint32_t ab = q & 1u;
          const uint32_t d_tmem = tmem_base + ab * C::ACC_STRIDE;
          uint32_t accum = 0;
          auto issue_dt = [&](int dt) {
            if (!(w_ready & (1u << dt))) {
              mbar_wait(&w_full[dt], 0);
              mbar_wait(&w_peer[dt], 0);
              tc_fence_after();
              w_ready |= 1u << dt;
            }
            const uint32_t a_base = planes_addr + plane_slot(pl + dt - 1) * C::SLOT_STRIDE;
            const uint32_t b_base = w_addr + dt * C::DT_BYTES;
            if constexpr (C::HEAD) {
              // 9 in-plane taps, 8 channels (16 B) each -> 5 MMAs of K=16: (tap0 | zero-weight dummy), (1|2) ... (7|8)
              // (tf32: 4 channels per 16-byte voxel, K = 8: the same two 16-byte K groups)
              const uint64_t a0 = make_smem_desc(a_base, 0, BOX_W * 16, 0);   // LBO chosen per tap pair below
              const uint64_t b0 = make_smem_desc(b_base, C::ROWS_PER_CTA * 16, 128, 0);
#pragma unroll
              for (int s = 0; s < 5; ++s) {
                const int ta = (s == 0) ? 0 : 2 * s - 1;
                const int tb = (s == 0) ? 1 : 2 * s;
                const uint32_t offa = ((ta / 3) * BOX_W + (ta % 3)) * 16;
                const uint32_t offb = ((tb / 3) * BOX_W + (tb % 3)) * 16;
                umma_ss_pair<C::TF32>(d_tmem, desc_add_lo(a0, desc_lo_delta(offa, offb - offa)),
                                      desc_add_lo(b0, desc_lo_delta(s * 1024)), idesc, accum);
                accum = 1;
              }
            } else {
              const uint64_t a0 = make_smem_desc(a_base, 16, BOX_W * 128, 2);
              const uint64_t b0 = make_smem_desc(b_base, 16, 1024, 2);
#pragma unroll
              for (int s = 0; s < 9; ++s) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_ss_pair<C::TF32>(d_tmem, desc_add_lo(a0, desc_lo_delta(((s / 3) * BOX_W + (s % 3)) * 128 + k * 32)),
                                        desc_add_lo(b0, desc_lo_delta(s * (C::ROWS_PER_CTA * 128) + k * 32)), idesc,
                                        accum);
                  accum = 1;
                }
              }
            }
          };
          // (steady state: nothing to wait for here — see below; the first plane of an item waits for its own planes,
          // the first two planes of the pair find their accumulator stages free)
          wait_planes(pl);
          if (q < 2) mbar_wait(&acc_empty[ab], 1u);
          tc_fence_after();
          if (pl - 1 >= 0) issue_dt(0);
          issue_dt(1);
          if (pl + 1 < T) wait_planes(pl + 1);
          if (q + 1 < q_end && q + 1 >= 2) mbar_wait(&acc_empty[ab ^ 1u], (((q + 1) >> 1) & 1u) ^ 1u);
          tc_fence_after();
          if (pl + 1 < T) issue_dt(2);
          umma_commit_pair(&acc_full[ab], 3);
          // plane pl-1 has fed its last output plane; the item's last output also frees planes pl and pl+1
          if (pl - 1 >= tlo) umma_commit_pair(&a_empty[plane_slot(pl - 1)], 3);
          if (pl == it.t1 - 1) {
            umma_commit_pair(&a_empty[plane_slot(pl)], 3);
            if (pl + 1 <= thi) umma_commit_pair(&a_empty[plane_slot(pl + 1)], 3);
          }
        }
        j0 += static_cast<uint32_t>(thi - tlo + 1);
      }
      // never leave with a bulk copy into this CTA's shared memory still in flight (T == 1 uses only dt == 1)
#pragma unroll
      for (int dt = 0; dt < 3; ++dt)
        if (!(w_ready & (1u << dt))) mbar_wait(&w_full[dt], 0);
    }
    __syncwarp();
  } else {
    // =========================================================================== epilogue (both CTAs, 4 warps per group)
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const uint32_t grp = static_cast<uint32_t>(warp - 2) >> 2;   // EPI_GROUPS == 2: group g owns accumulator stage g
    const int row = quad * 32 + lane;          // accumulator row == voxel of this CTA's tile
    const int hh = row >> 3, ww = row & 7;
    uint32_t leader_empty[2];
    leader_empty[0] = map_to_cta(smem_u32(&acc_empty[0]), 0);
    leader_empty[1] = map_to_cta(smem_u32(&acc_empty[1]), 0);
    // epilogue vectors: fetched by the warps that use them, off the critical path of the prologue (they have nothing to
    // do until the first accumulator is complete)
    if (threadIdx.x - 64 < 64) {
      const int c = threadIdx.x - 64;
      scale_sm[c] = c < C::NOUT ? p.scale[c] : 0.f;
      shift_sm[c] = c < C::NOUT ? p.shift[c] : 0.f;
    }
    asm volatile("bar.sync 6, %0;" ::"n"(128 * C::EPI_GROUPS) : "memory");
    uint32_t q = 0;
    // ---- TMA-store epilogue (bf16 channels-last output): the four epilogue warps write their 128 rows (128 B each) into
    // a 128B-swizzled shared-memory tile — conflict-free 16-byte stores — and ONE bulk tensor store moves the tile out:
    // full 128-byte lines instead of 32 partial-line store streams per warp instruction, and the tile's out-of-range
    // rows / columns are clipped by the tensor map instead of being predicated per voxel.
    const bool tma_out = C::STAGES > 0 && !C::TF32 && p.tma_out != 0;
    uint32_t stage_n = 0;                      // tiles staged so far by this CTA
    // EPI_GROUPS == 2: group g stages into tile g and synchronises on its own named barrier
    const bool store_thread = (warp == 2 + 4 * static_cast<int>(grp) && lane == 0);
    const uint32_t stage_bar = 4u + grp;
    auto stage_ptr = [&]() {
      return stage + (C::EPI_GROUPS == 2 ? grp : (C::STAGES > 1 ? (stage_n & 1u) : 0u)) * STAGE_BYTES;
    };
    auto stage_row = [&]() { return reinterpret_cast<__nv_bfloat16*>(stage_ptr() + row * 128); };
    auto stage_row_begin = [&]() {
      // the tile about to be overwritten must have been read by the bulk store issued STAGES tiles ago
      if (store_thread) {
        if constexpr (C::STAGES > 1 && C::EPI_GROUPS == 1) tma_store_wait_read<1>();
        else tma_store_wait_read<0>();
      }
      asm volatile("bar.sync %0, 128;" ::"r"(stage_bar) : "memory");
    };
    auto stage_row_end = [&](const Unit& u, int plane) {
      fence_proxy_async();                     // generic-proxy writes -> visible to the async proxy (TMA)
      asm volatile("bar.sync %0, 128;" ::"r"(stage_bar) : "memory");
      if (store_thread) {
        tma_store_5d(&tmap_out, stage_ptr(), 0, u.w0, u.h0, plane, u.n);
        tma_store_commit();
      }
      ++stage_n;
    };
    float st_s0 = 0.f, st_s1 = 0.f, st_q0 = 0.f, st_q1 = 0.f;   // BatchNorm partial sums of channels 2*lane, 2*lane+1
    int g0, g1;
    work_range(pair, n_pairs, p, g0, g1);
    for (int g = g0; g < g1;) {
      const Item it = next_item(g, g1, T);
      g += it.t1 - it.t0;
      const Unit un = decode_unit(it.u, p, rank);
      const int h = un.h0 + hh, w = un.w0 + ww;
      const bool inb = (h < p.H) && (w < p.W);
      for (int pl = it.t0; pl < it.t1; ++pl, ++q) {
        const uint32_t ab = q & 1u;
        if (C::EPI_GROUPS == 2 && ab != grp) continue;
        mbar_wait(&acc_full[ab], (q >> 1) & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + ab * C::ACC_STRIDE;
        const size_t vox = ((static_cast<size_t>(un.n) * T + pl) * p.H + h) * p.W + w;
        if constexpr (C::NOUT == 64) {
          uint32_t r0[32], r1[32];
          tmem_ld32(taddr, r0);
          tmem_ld32(taddr + 32, r1);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(leader_empty[ab]);
          if constexpr (!STATS) {
            if (inb) {
              if (p.out_mode == CONV_OUT_F32_RAW) {
                float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + vox * 64);
                if constexpr (C::TF32) {
                  if (p.addend) {   // running fp32 partial sum of a split-Cin layer (may alias the output)
                    const float4* ad = reinterpret_cast<const float4*>(p.addend + vox * 64);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                      const float4 a0 = ad[c], a1 = ad[8 + c];
                      r0[4 * c] = __float_as_uint(__uint_as_float(r0[4 * c]) + a0.x);
                      r0[4 * c + 1] = __float_as_uint(__uint_as_float(r0[4 * c + 1]) + a0.y);
                      r0[4 * c + 2] = __float_as_uint(__uint_as_float(r0[4 * c + 2]) + a0.z);
                      r0[4 * c + 3] = __float_as_uint(__uint_as_float(r0[4 * c + 3]) + a0.w);
                      r1[4 * c] = __float_as_uint(__uint_as_float(r1[4 * c]) + a1.x);
                      r1[4 * c + 1] = __float_as_uint(__uint_as_float(r1[4 * c + 1]) + a1.y);
                      r1[4 * c + 2] = __float_as_uint(__uint_as_float(r1[4 * c + 2]) + a1.z);
                      r1[4 * c + 3] = __float_as_uint(__uint_as_float(r1[4 * c + 3]) + a1.w);
                    }
                  }
                }
#pragma unroll
                for (int c = 0; c < 8; ++c)
                  dst[c] = make_float4(__uint_as_float(r0[4 * c]), __uint_as_float(r0[4 * c + 1]),
                                       __uint_as_float(r0[4 * c + 2]), __uint_as_float(r0[4 * c + 3]));
#pragma unroll
                for (int c = 0; c < 8; ++c)
                  dst[8 + c] = make_float4(__uint_as_float(r1[4 * c]), __uint_as_float(r1[4 * c + 1]),
                                           __uint_as_float(r1[4 * c + 2]), __uint_as_float(r1[4 * c + 3]));
              } else if constexpr (C::TF32) {
                const float* add = p.addend ? p.addend + vox * 64 : nullptr;
                float* dst = static_cast<float*>(p.out) + vox * p.out_pitch + p.out_coff;
                epilogue_f32_dispatch<false>(p.act, r0, r1, scale_sm, shift_sm, add, dst);
              } else if (!tma_out) {
                const float* add = p.addend ? p.addend + vox * 64 : nullptr;
                __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.out) + vox * p.out_pitch + p.out_coff;
                const __nv_bfloat16* mk =
                    p.mask ? static_cast<const __nv_bfloat16*>(p.mask) + vox * p.mask_pitch : nullptr;
                epilogue_bf16_dispatch<false>(p.act, r0, r1, scale_sm, shift_sm, add, dst, mk);
              }
            }
            if constexpr (C::STAGES > 0 && !C::TF32) {
              if (tma_out) {
                stage_row_begin();
                if (inb) {
                  const float* add = p.addend ? p.addend + vox * 64 : nullptr;
                  const __nv_bfloat16* mk =
                      p.mask ? static_cast<const __nv_bfloat16*>(p.mask) + vox * p.mask_pitch : nullptr;
                  epilogue_bf16_dispatch<false>(p.act, r0, r1, scale_sm, shift_sm, add, stage_row(), mk, row & 7);
                }
                stage_row_end(un, pl);
              }
            }
          } else {
            // bf16 output + BatchNorm batch statistics (sum, sum of squares of the values AS STORED, i.e. bf16-rounded)
            if constexpr (C::STAGES > 0 && !C::TF32) {
              if (tma_out) stage_row_begin();      // (a CTA barrier: outside the per-voxel `inb` divergence)
            }
            if (inb) {
              const float* add = p.addend ? p.addend + vox * 64 : nullptr;
              if constexpr (C::TF32) {
                float* dst = static_cast<float*>(p.out) + vox * p.out_pitch + p.out_coff;
                epilogue_f32_dispatch<true>(p.act, r0, r1, scale_sm, shift_sm, add, dst);
              } else {
                const __nv_bfloat16* mk =
                    p.mask ? static_cast<const __nv_bfloat16*>(p.mask) + vox * p.mask_pitch : nullptr;
                if constexpr (C::STAGES > 0) {
                  if (tma_out) {
                    epilogue_bf16_dispatch<true>(p.act, r0, r1, scale_sm, shift_sm, add, stage_row(), mk, row & 7);
                  } else {
                    __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.out) + vox * p.out_pitch + p.out_coff;
                    epilogue_bf16_dispatch<true>(p.act, r0, r1, scale_sm, shift_sm, add, dst, mk);
                  }
                } else {
                  __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.out) + vox * p.out_pitch + p.out_coff;
                  epilogue_bf16_dispatch<true>(p.act, r0, r1, scale_sm, shift_sm, add, dst, mk);
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) r0[i] = r1[i] = 0u;
            }
            if constexpr (C::STAGES > 0 && !C::TF32) {
              if (tma_out) stage_row_end(un, pl);
            }
            // per-channel sums over the 32 voxels of this warp: a butterfly that halves the channel set at every
            // step (62 shuffles per statistic instead of 320); lane L ends up with channels 2L, 2L+1.
            float s[32], sq[32];
            const bool up16 = lane & 16;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float lo = __uint_as_float(r0[i]), hi = __uint_as_float(r1[i]);
              const float keep = up16 ? hi : lo, send = up16 ? lo : hi;
              const float got = __shfl_xor_sync(0xffffffffu, send, 16);
              const float gq = __shfl_xor_sync(0xffffffffu, send * send, 16);
              s[i] = keep + got;
              sq[i] = fmaf(keep, keep, gq);
            }
#define HPVG_BFLY(M_, N_)                                                              \
  {                                                                                    \
    const bool upm = lane & (M_);                                                      \
    _Pragma("unroll") for (int i = 0; i < (N_); ++i) {                                 \
      const float ks = upm ? s[(N_) + i] : s[i], ss = upm ? s[i] : s[(N_) + i];        \
      const float kq = upm ? sq[(N_) + i] : sq[i], sx = upm ? sq[i] : sq[(N_) + i];    \
      s[i] = ks + __shfl_xor_sync(0xffffffffu, ss, (M_));                              \
      sq[i] = kq + __shfl_xor_sync(0xffffffffu, sx, (M_));                             \
    }                                                                                  \
  }
            HPVG_BFLY(8, 16)
            HPVG_BFLY(4, 8)
            HPVG_BFLY(2, 4)
            HPVG_BFLY(1, 2)
#undef HPVG_BFLY
            st_s0 += s[0];
            st_s1 += s[1];
            st_q0 += sq[0];
            st_q1 += sq[1];
          }
        } else {
          uint32_t r[16];
          tmem_ld16(taddr, r);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(leader_empty[ab]);
          if (inb) {
            // fp32 NCDHW, cout_real channels: out[n][c][t][h][w] = act(acc*scale + shift (+ residual))
            const size_t plane_sz = static_cast<size_t>(p.H) * p.W;
            const size_t sp = (static_cast<size_t>(pl) * p.H + h) * p.W + w;
            const size_t chan_sz = plane_sz * T;
            float* out = static_cast<float*>(p.out);
            float v[4];
            size_t idx[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              idx[c] = (static_cast<size_t>(un.n) * p.cout_real + c) * chan_sz + sp;
              v[c] = fmaf(__uint_as_float(r[c]), scale_sm[c], shift_sm[c]);
              if (p.addend && c < p.cout_real) v[c] += p.addend[idx[c]];
            }
            if (p.act == CONV_ACT_TANH) {
#pragma unroll
              for (int c = 0; c < 4; ++c) v[c] = tanhf(v[c]);
            } else if (p.act == CONV_ACT_LRELU) {
#pragma unroll
              for (int c = 0; c < 4; ++c) v[c] = v[c] > 0.f ? v[c] : 0.2f * v[c];
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (c < p.cout_real) out[idx[c]] = v[c];
          }
        }
      }
    }
    if (tma_out && store_thread) tma_store_wait_all<0>();   // the staged tiles have reached global memory
    if constexpr (STATS) {
      {
        // all MMAs of this pair have completed (the last acc_full fired), so the plane ring is free: use it to combine
        // the 4 epilogue warps, then one fp64 atomic per channel and statistic per CTA.
        float* red = reinterpret_cast<float*>(planes);   // [epilogue warps][2 stats][64 ch]
        const int ew = warp - 2;
        // (two groups: the other group's last plane may still be feeding the tensor pipe from the ring)
        if constexpr (C::EPI_GROUPS == 2) asm volatile("bar.sync 3, 256;" ::: "memory");
        red[(ew * 2 + 0) * 64 + 2 * lane] = st_s0;
        red[(ew * 2 + 0) * 64 + 2 * lane + 1] = st_s1;
        red[(ew * 2 + 1) * 64 + 2 * lane] = st_q0;
        red[(ew * 2 + 1) * 64 + 2 * lane + 1] = st_q1;
        if constexpr (C::EPI_GROUPS == 2) asm volatile("bar.sync 3, 256;" ::: "memory");
        else asm volatile("bar.sync 1, 128;" ::: "memory");
        if (ew < 4) {
          const int t = threadIdx.x - 64;   // 0..127 : statistic (t >> 6), channel (t & 63)
          float tot = red[t] + red[128 + t] + red[256 + t] + red[384 + t];
          if constexpr (C::EPI_GROUPS == 2) tot += red[512 + t] + red[640 + t] + red[768 + t] + red[896 + t];
          // CTA partials are summed in CTA order by whichever CTA finishes last: bitwise reproducible statistics
          det_reduce_128<true>(static_cast<double>(tot), t, true, blockIdx.x, gridDim.x, p.det, p.stats, p.stats + 64,
                               true);
        }
      }
    }
  }
  // =========================================================================== teardown
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
}

template <int MODE>
constexpr int smem_bytes_for() {
  using C = Cfg<MODE>;
  static_assert((C::SLOTS * C::SLOT_STRIDE) % 1024 == 0, "the staging tiles must stay 1024-byte aligned");
  return 1024 + ((C::W_BYTES + 1023) & ~1023) + C::SLOTS * C::SLOT_STRIDE + C::STAGES * STAGE_BYTES + 512 +
         (2 * C::SLOTS + 10) * 8 + 16;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

template <int MODE, bool STATS>
cudaError_t launch_mode(const CUtensorMap& tmap, const CUtensorMap& tmap_out, const ConvParams& prm, int n_pairs,
                        cudaStream_t stream) {
  static bool configured = false;
  constexpr int smem = smem_bytes_for<MODE>();
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(conv3d_umma_kernel<MODE, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  launch(conv3d_umma_kernel<MODE, STATS>, dim3(2 * n_pairs), dim3(num_threads<MODE>()), smem, stream, tmap, tmap_out, prm);
  return cudaGetLastError();
}

}  // namespace

const char* conv3d_umma_launch(const ConvLaunch& L, cudaStream_t stream) {
  if (L.mode == CONV_MODE_64_T || L.mode == CONV_MODE_T32_T) return conv3d_tail_launch(L, 2 * L.max_pairs, stream);
  EncodeTiledFn enc = get_encode();
  if (!enc) return "cuTensorMapEncodeTiled entry point not available";
  const bool tf32 = conv_mode_is_tf32(L.mode);
  const bool head = (L.mode == CONV_MODE_8_64 || L.mode == CONV_MODE_T4_64);
  const int esz = tf32 ? 4 : 2;                      // bytes per activation element
  const int cin_pitch = L.in_pitch;                  // channels per voxel row in the input tensor
  const int cin_box = head ? (tf32 ? 4 : 8) : (tf32 ? 32 : 64);   // one 16-byte (head) or 128-byte box row
  if (cin_pitch < cin_box || ((cin_pitch * esz) & 15)) return "input pitch must be a multiple of 16 bytes and >= the box";
  if ((reinterpret_cast<uintptr_t>(L.in) & 15) != 0) return "input not 16-byte aligned";
  CUtensorMap tmap;
  cuuint64_t gd[5] = {static_cast<cuuint64_t>(cin_box), static_cast<cuuint64_t>(L.W), static_cast<cuuint64_t>(L.H),
                      static_cast<cuuint64_t>(L.T), static_cast<cuuint64_t>(L.N)};
  const cuuint64_t vox = static_cast<cuuint64_t>(cin_pitch) * esz;
  cuuint64_t gs[4] = {vox, vox * L.W, vox * L.W * L.H, vox * L.W * L.H * L.T};
  cuuint32_t bx[5] = {static_cast<cuuint32_t>(cin_box), BOX_W, BOX_H, 1, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  // head convs on a densely packed 8-channel tensor: voxels of an image row are contiguous, so (C, W) merge into one
  // dimension and every box row becomes ONE 160-byte request instead of ten 16-byte ones
  static const bool no_merge = getenv("HPVG_NO_MERGE") != nullptr;
  const bool merged = (head && cin_pitch == cin_box && !no_merge);
  if (merged) {
    gd[0] = static_cast<cuuint64_t>(cin_box) * L.W; gd[1] = L.H; gd[2] = L.T; gd[3] = L.N; gd[4] = 1;
    gs[0] = vox * L.W; gs[1] = vox * L.W * L.H; gs[2] = vox * L.W * L.H * L.T; gs[3] = vox * L.W * L.H * L.T * L.N;
    bx[0] = cin_box * BOX_W; bx[1] = BOX_H; bx[2] = 1; bx[3] = 1; bx[4] = 1;
  }
  CUresult r = enc(&tmap, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5,
                   const_cast<void*>(L.in), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   head ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed";

  if (L.act == CONV_ACT_LRELU_MASK && (tf32 || !L.mask || L.out_mode != CONV_OUT_BF16_NDHWC || (L.mask_pitch & 7) ||
                                       (reinterpret_cast<uintptr_t>(L.mask) & 15)))
    return "LRELU_MASK needs a 16-byte aligned bf16 mask tensor and the bf16 output mode";
  ConvParams prm;
  prm.N = L.N;
  prm.T = L.T;
  prm.H = L.H;
  prm.W = L.W;
  prm.w_tiles = (L.W + TILE_W - 1) / TILE_W;
  const int h_tiles = (L.H + TILE_H - 1) / TILE_H;
  prm.h_pairs = (h_tiles + 1) / 2;
  prm.n_units = L.N * prm.w_tiles * prm.h_pairs;
  prm.wimg = static_cast<const uint8_t*>(L.wimg);
  prm.scale = L.scale;
  prm.shift = L.shift;
  prm.act = L.act;
  prm.out_mode = L.out_mode;
  prm.out = L.out;
  prm.out_pitch = L.out_pitch;
  prm.out_coff = L.out_coff;
  prm.cout_real = L.cout_real;
  prm.addend = L.addend;
  prm.stats = L.stats;
  prm.det = L.det;
  if (L.stats && (!L.det.partials || !L.det.counter)) return "fused statistics need the stream's reduction scratch";
  prm.mask = L.mask;
  // output tensor map of the TMA-store epilogue: (C = 64 channels at out_coff, W, H, T, N) over the channels-last bf16
  // output, box = one CTA tile (64 x TILE_W x TILE_H), 128B swizzle; rows / columns past H / W are clipped by the store
  CUtensorMap tmap_out = tmap;
  prm.tma_out = 0;
  static const bool no_tma_store = getenv("HPVG_NO_TMA_STORE") != nullptr;
  if (!tf32 && !no_tma_store && L.out_mode == CONV_OUT_BF16_NDHWC &&
      (L.mode == CONV_MODE_64_64 || L.mode == CONV_MODE_8_64) &&
      (L.out_pitch & 7) == 0 && (L.out_coff & 7) == 0 && (reinterpret_cast<uintptr_t>(L.out) & 15) == 0) {
    const cuuint64_t ovox = static_cast<cuuint64_t>(L.out_pitch) * 2;
    cuuint64_t ogd[5] = {64, static_cast<cuuint64_t>(L.W), static_cast<cuuint64_t>(L.H), static_cast<cuuint64_t>(L.T),
                         static_cast<cuuint64_t>(L.N)};
    cuuint64_t ogs[4] = {ovox, ovox * L.W, ovox * L.W * L.H, ovox * L.W * L.H * L.T};
    cuuint32_t obx[5] = {64, TILE_W, TILE_H, 1, 1};
    void* obase = static_cast<char*>(L.out) + static_cast<size_t>(L.out_coff) * 2;
    CUresult ro = enc(&tmap_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, obase, ogd, ogs, obx, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (ro != CUDA_SUCCESS) return "cuTensorMapEncodeTiled (output) failed";
    prm.tma_out = 1;
  }
  prm.mask_pitch = L.mask_pitch;
  prm.in_merged = merged ? 1 : 0;
  const long long out_planes = static_cast<long long>(prm.n_units) * L.T;   // work items of one plane each
  if (out_planes >= (1LL << 31) / (L.max_pairs + 1)) return "problem too large for the int work partition";
  int n_pairs = out_planes < L.max_pairs ? static_cast<int>(out_planes) : L.max_pairs;
  if (n_pairs < 1) return nullptr;
  cudaError_t e;
  if (tf32) {
    if (L.out_mode != CONV_OUT_F32_NDHWC && L.out_mode != CONV_OUT_F32_RAW) return "bad out_mode for a tf32 conv";
    if (L.out_mode == CONV_OUT_F32_NDHWC && ((L.out_pitch & 3) || (L.out_coff & 3) ||
                                             (reinterpret_cast<uintptr_t>(L.out) & 15)))
      return "fp32 channels-last output: pitch / offset must be multiples of 4 channels, pointer 16-byte aligned";
    if (L.addend && (reinterpret_cast<uintptr_t>(L.addend) & 15)) return "addend not 16-byte aligned";
    if (L.stats && L.out_mode != CONV_OUT_F32_NDHWC) return "fused statistics need the channels-last output mode";
  }
  switch (L.mode) {
    case CONV_MODE_64_64:
      if (L.out_mode != CONV_OUT_BF16_NDHWC && L.out_mode != CONV_OUT_F32_RAW) return "bad out_mode for 64->64";
      if (L.stats && L.out_mode != CONV_OUT_BF16_NDHWC) return "fused statistics need the bf16 output mode";
      e = L.stats ? launch_mode<CONV_MODE_64_64, true>(tmap, tmap_out, prm, n_pairs, stream)
                  : launch_mode<CONV_MODE_64_64, false>(tmap, tmap_out, prm, n_pairs, stream);
      break;
    case CONV_MODE_64_16:
      if (L.out_mode != CONV_OUT_F32_NCDHW || L.cout_real > 4) return "bad out_mode for 64->16";
      if (L.stats) return "fused statistics need a 64-channel output";
      e = launch_mode<CONV_MODE_64_16, false>(tmap, tmap_out, prm, n_pairs, stream);
      break;
    case CONV_MODE_8_64:
      if (L.out_mode != CONV_OUT_BF16_NDHWC && L.out_mode != CONV_OUT_F32_RAW) return "bad out_mode for 8->64";
      if (L.stats && L.out_mode != CONV_OUT_BF16_NDHWC) return "fused statistics need the bf16 output mode";
      e = L.stats ? launch_mode<CONV_MODE_8_64, true>(tmap, tmap_out, prm, n_pairs, stream)
                  : launch_mode<CONV_MODE_8_64, false>(tmap, tmap_out, prm, n_pairs, stream);
      break;
    case CONV_MODE_T32_64:
      e = L.stats ? launch_mode<CONV_MODE_T32_64, true>(tmap, tmap_out, prm, n_pairs, stream)
                  : launch_mode<CONV_MODE_T32_64, false>(tmap, tmap_out, prm, n_pairs, stream);
      break;
    case CONV_MODE_T4_64:
      e = L.stats ? launch_mode<CONV_MODE_T4_64, true>(tmap, tmap_out, prm, n_pairs, stream)
                  : launch_mode<CONV_MODE_T4_64, false>(tmap, tmap_out, prm, n_pairs, stream);
      break;
    default:
      return "unknown conv mode";
  }
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

int conv3d_umma_wimg_bytes(int mode) {
  switch (mode) {
    case CONV_MODE_64_64: return 2 * Cfg<CONV_MODE_64_64>::W_BYTES;
    case CONV_MODE_64_16: return 2 * Cfg<CONV_MODE_64_16>::W_BYTES;
    case CONV_MODE_8_64: return 2 * Cfg<CONV_MODE_8_64>::W_BYTES;
    case CONV_MODE_64_T: return conv3d_tail_wimg_bytes();
    case CONV_MODE_T32_64: return 2 * Cfg<CONV_MODE_T32_64>::W_BYTES;
    case CONV_MODE_T4_64: return 2 * Cfg<CONV_MODE_T4_64>::W_BYTES;
    case CONV_MODE_T32_T: return conv3d_tail_wimg_bytes();
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Filter-bank packing: fp32 (Cout, Cin, 3,3,3) [or (Cout,Cin,3,3) with kt==1] -> the per-CTA shared-memory image the
// conv kernel bulk-copies.  transpose_flip=1 packs the data-gradient filter (Cin<->Cout swapped, taps mirrored).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pack_weights_element(const float* __restrict__ w, int cout, int cin, int kt, int mode,
                                                     int transpose_flip, int cin_off, int cout_off, int w_cout,
                                                     int w_cin, __nv_bfloat16* __restrict__ img, int idx) {
  // one thread per bf16 element of the image
  int rank, tap, row, ci;
  float val = 0.f;
  bool valid = true;
  if (mode == CONV_MODE_64_T) {
    // [dt 3][row n = (dh*3+dw)*3 + co (32 rows, 27 used)][64 ci], 128B-swizzled rows; no CTA split
    const int dt = idx / (32 * 64);
    const int rem = idx - dt * 32 * 64;
    const int n = rem / 64, within = rem - n * 64;
    const int chunk = (within >> 3) ^ (n & 7);
    ci = chunk * 8 + (within & 7);
    const int s = n / 3, co3 = n - s * 3;
    valid = n < 27;
    {
      const int dh = s / 3, dw = s - dh * 3;
      int dtt = dt;
      if (kt == 1) {
        valid = valid && (dt == 1);
        dtt = 0;
      }
      if (valid && co3 < cout && ci < cin) {
        if (!transpose_flip) {
          const size_t o =
              ((((static_cast<size_t>(co3 + cout_off) * w_cin) + (ci + cin_off)) * kt + dtt) * 3 + dh) * 3 + dw;
          val = w[o];
        } else {
          const int fdt = (kt == 1) ? 0 : 2 - dtt;
          const size_t o =
              ((((static_cast<size_t>(ci + cin_off) * w_cin) + (co3 + cout_off)) * kt + fdt) * 3 + (2 - dh)) * 3 + (2 - dw);
          val = w[o];
        }
      }
    }
    img[idx] = __float2bfloat16_rn(val);
    return;
  }
  if (mode == CONV_MODE_64_64 || mode == CONV_MODE_64_16) {
    const int rows = (mode == CONV_MODE_64_64) ? 32 : 8;
    const int per_rank = 27 * rows * 64;
    rank = idx / per_rank;
    int rem = idx - rank * per_rank;
    tap = rem / (rows * 64);
    rem -= tap * rows * 64;
    row = rem / 64;
    const int within = rem - row * 64;          // element position inside the swizzled 128 B row
    const int chunk_sw = within >> 3, e = within & 7;
    const int chunk = chunk_sw ^ (row & 7);     // un-swizzle: stored position -> logical chunk
    ci = chunk * 8 + e;
  } else {
    // [rank][dt 3][step 5][kg 2][row 32][8]
    const int per_rank = 3 * 5 * 2 * 32 * 8;
    rank = idx / per_rank;
    int rem = idx - rank * per_rank;
    const int dt = rem / (5 * 512);
    rem -= dt * 5 * 512;
    const int step = rem / 512;
    rem -= step * 512;
    const int kg = rem / 256;
    rem -= kg * 256;
    row = rem / 8;
    ci = rem & 7;
    int s;  // in-plane tap index 0..8
    if (step == 0) {
      s = 0;
      valid = (kg == 0);
    } else {
      s = 2 * step - 1 + kg;
    }
    tap = dt * 9 + s;
  }
  const int rows_per = (mode == CONV_MODE_64_16) ? 8 : 32;
  const int co = rank * rows_per + row;
  int dt = tap / 9, dh = (tap / 3) % 3, dw = tap % 3;
  if (kt == 1) {  // 2-D filter: only the centre temporal tap is populated
    valid = valid && (dt == 1);
    dt = 0;
  }
  if (valid && co < cout && ci < cin) {
    // logical conv: out[co] += w_eff[co][ci][dt][dh][dw] * in[ci]
    if (!transpose_flip) {
      const size_t o = ((((static_cast<size_t>(co + cout_off) * w_cin) + (ci + cin_off)) * kt + dt) * 3 + dh) * 3 + dw;
      val = w[o];
    } else {
      // data gradient: w_eff[co=ci_fwd][ci=co_fwd][taps mirrored]
      const int fdt = (kt == 1) ? 0 : 2 - dt;
      const size_t o =
          ((((static_cast<size_t>(ci + cin_off) * w_cin) + (co + cout_off)) * kt + fdt) * 3 + (2 - dh)) * 3 + (2 - dw);
      val = w[o];
    }
  }
  (void)w_cout;
  img[idx] = __float2bfloat16_rn(val);
}

__global__ void pack_weights_kernel(const float* __restrict__ w, int cout, int cin, int kt, int mode,
                                    int transpose_flip, int cin_off, int cout_off, int w_cout, int w_cin,
                                    __nv_bfloat16* __restrict__ img, int total) {
  pdl_grid_sync();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < total)
    pack_weights_element(w, cout, cin, kt, mode, transpose_flip, cin_off, cout_off, w_cout, w_cin, img, idx);
}

// Many banks (and epilogue-vector pairs) in ONE launch: at the coarse scales of the pyramid an iteration repacks ~45
// filter banks of its trainable layers, each a 3-4 us launch on the critical path of a launch-latency-bound step.
// blockIdx.y selects the table entry; mode < 0 marks an "affine" entry: img = fp32 [2][64] = (1 | 1/sigma, bias).
__global__ void pack_weights_multi_kernel(const __grid_constant__ PackTable tab) {
  pdl_grid_sync();
  const PackEntry& e = tab.e[blockIdx.y];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= e.total) return;
  if (e.mode < 0) {
    float* out = static_cast<float*>(e.img);
    const int c = idx & 63;
    out[idx] = (idx < 64) ? 1.0f : ((c < e.cout && e.w) ? e.w[c] : 0.f);
    return;
  }
  pack_weights_element(e.w, e.cout, e.cin, e.kt, e.mode, e.flip, e.cin_off, e.cout_off, e.w_cout, e.w_cin,
                       static_cast<__nv_bfloat16*>(e.img), idx);
}

// The same images for the kind::tf32 variants: fp32 elements rounded to tf32, 32 input channels per 128-byte row
// (16-byte swizzle chunk = 4 channels), 4 channels per 16-byte head voxel.  One thread per fp32 element.
__global__ void pack_weights_tf32_kernel(const float* __restrict__ w, int cout, int cin, int kt, int mode,
                                         int transpose_flip, int cin_off, int cout_off, int w_cin,
                                         float* __restrict__ img, int total) {
  pdl_grid_sync();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int co, ci, dt, dh, dw;
  bool valid = true;
  if (mode == CONV_MODE_T32_T) {
    // [dt 3][row n = (dh*3+dw)*3 + co (32 rows, 27 used)][32 ci], 128B-swizzled rows
    dt = idx / (32 * 32);
    const int rem = idx - dt * 32 * 32;
    const int n = rem / 32, within = rem - n * 32;
    const int chunk = (within >> 2) ^ (n & 7);
    ci = chunk * 4 + (within & 3);
    const int s = n / 3;
    co = n - s * 3;
    valid = n < 27;
    dh = s / 3;
    dw = s - dh * 3;
  } else if (mode == CONV_MODE_T32_64) {
    // [rank 2][tap 27][row 32][32 ci], 128B-swizzled rows
    const int per_rank = 27 * 32 * 32;
    const int rank = idx / per_rank;
    int rem = idx - rank * per_rank;
    const int tap = rem / (32 * 32);
    rem -= tap * 32 * 32;
    const int row = rem / 32, within = rem - row * 32;
    const int chunk = (within >> 2) ^ (row & 7);
    ci = chunk * 4 + (within & 3);
    co = rank * 32 + row;
    dt = tap / 9; dh = (tap / 3) % 3; dw = tap % 3;
  } else {
    // [rank][dt 3][step 5][kg 2][row 32][4]
    const int per_rank = 3 * 5 * 2 * 32 * 4;
    const int rank = idx / per_rank;
    int rem = idx - rank * per_rank;
    dt = rem / (5 * 256);
    rem -= dt * 5 * 256;
    const int step = rem / 256;
    rem -= step * 256;
    const int kg = rem / 128;
    rem -= kg * 128;
    const int row = rem / 4;
    ci = rem & 3;
    int s;
    if (step == 0) {
      s = 0;
      valid = (kg == 0);
    } else {
      s = 2 * step - 1 + kg;
    }
    co = rank * 32 + row;
    dh = s / 3; dw = s % 3;
  }
  if (kt == 1) {  // 2-D filter: only the centre temporal tap is populated
    valid = valid && (dt == 1);
    dt = 0;
  }
  float val = 0.f;
  if (valid && co < cout && ci < cin) {
    if (!transpose_flip) {
      const size_t o = ((((static_cast<size_t>(co + cout_off) * w_cin) + (ci + cin_off)) * kt + dt) * 3 + dh) * 3 + dw;
      val = w[o];
    } else {
      const int fdt = (kt == 1) ? 0 : 2 - dt;
      const size_t o =
          ((((static_cast<size_t>(ci + cin_off) * w_cin) + (co + cout_off)) * kt + fdt) * 3 + (2 - dh)) * 3 + (2 - dw);
      val = w[o];
    }
  }
  img[idx] = round_tf32(val);
}

const char* conv3d_pack_weights(const float* w, int w_cout, int w_cin, int kt, int mode, int transpose_flip,
                                int cout_off, int cout, int cin_off, int cin, void* img, cudaStream_t stream) {
  if (kt != 1 && kt != 3) return "kernel depth must be 1 or 3";
  if (conv_mode_is_tf32(mode)) {
    const int n = conv3d_umma_wimg_bytes(mode) / 4;
    launch(pack_weights_tf32_kernel, (n + 255) / 256, 256, 0, stream, w, cout, cin, kt, mode, transpose_flip, cin_off,
                                                                   cout_off, w_cin, static_cast<float*>(img), n);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
  }
  const int total = conv3d_umma_wimg_bytes(mode) / 2;
  if (total == 0) return "unknown conv mode";
  launch(pack_weights_kernel, (total + 255) / 256, 256, 0, stream, w, cout, cin, kt, mode, transpose_flip, cin_off,
                                                                cout_off, w_cout, w_cin,
                                                                static_cast<__nv_bfloat16*>(img), total);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* conv3d_pack_weights_multi(const PackEntry* entries, int n, cudaStream_t stream) {
  for (int base = 0; base < n; base += PACK_MAX_ENTRIES) {
    PackTable tab;
    const int cnt = (n - base < PACK_MAX_ENTRIES) ? n - base : PACK_MAX_ENTRIES;
    int max_total = 0;
    for (int i = 0; i < cnt; ++i) {
      tab.e[i] = entries[base + i];
      PackEntry& e = tab.e[i];
      if (e.mode < 0) {
        e.total = 128;
      } else {
        if (conv_mode_is_tf32(e.mode)) return "pack_weights_multi: bf16 kernel variants only";
        if (e.kt != 1 && e.kt != 3) return "kernel depth must be 1 or 3";
        e.total = conv3d_umma_wimg_bytes(e.mode) / 2;
        if (e.total == 0) return "unknown conv mode";
      }
      max_total = e.total > max_total ? e.total : max_total;
    }
    launch(pack_weights_multi_kernel, dim3((max_total + 255) / 256, cnt), 256, 0, stream, tab);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cudaGetErrorString(e);
  }
  return nullptr;
}

}  // namespace hpvg
