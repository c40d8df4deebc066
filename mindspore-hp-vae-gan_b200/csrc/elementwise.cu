// Bandwidth-bound kernels of the HP-VAE-GAN hot path: layout changes at the API edge, linear resize (fwd/bwd),
// the fused "upsample + noise + pack" block input stage, training-mode BatchNorm, spectral-norm power iteration,
// losses, reparameterisation and the multi-tensor clip+Adam step.  All plain SIMT, coalesced, 16-byte vector
// accesses where the layout allows it.  Reference call sites are cited per kernel.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "det_reduce.cuh"
#include "launch.cuh"
#include "elementwise.h"

namespace hpvg {

namespace {

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// one 16-byte block-input voxel: 8 bf16 channels, or (tf32 mode) 4 fp32 channels rounded to tf32
__device__ __forceinline__ uint4 xin_voxel(const float (&v)[8], int f32) {
  if (f32) {
    uint4 r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r.x) : "f"(v[0]));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r.y) : "f"(v[1]));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r.z) : "f"(v[2]));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r.w) : "f"(v[3]));
    return r;
  }
  return make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
}

// ----------------------------------------------------------------------------------------------- pack / unpack
// thread = (voxel, 8-channel group); consecutive threads walk consecutive voxels so the fp32 side is coalesced.
__global__ void pack_cl_kernel(const float* __restrict__ x, int C, long long sp /*T*H*W*/, long long voxels,
                               __nv_bfloat16* __restrict__ y, int c_pitch, int c_off, int groups) {
  pdl_grid_sync();
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= voxels * groups) return;
  const int g = static_cast<int>(gid / voxels);
  const long long v = gid - static_cast<long long>(g) * voxels;
  const long long n = v / sp, s = v - n * sp;
  float f[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = g * 8 + e;
    f[e] = (c < C) ? x[(n * C + c) * sp + s] : 0.f;
  }
  uint4 pk = make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
  *reinterpret_cast<uint4*>(y + v * c_pitch + c_off + g * 8) = pk;
}

// Skinny source (C <= 8: the 3-channel clips / gradients) zero-padded to a wide channels-last tensor: consecutive
// threads walk the 16-byte groups of ONE voxel, so a warp writes 512 contiguous bytes (the voxel-fastest mapping above
// would write 16 bytes out of every 128); only group 0 reads anything.
__global__ void pack_cl_skinny_kernel(const float* __restrict__ x, int C, long long sp, long long voxels,
                                      __nv_bfloat16* __restrict__ y, int c_pitch, int c_off, int groups) {
  pdl_grid_sync();
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= voxels * groups) return;
  const long long v = gid / groups;
  const int g = static_cast<int>(gid - v * groups);
  uint4 pk = make_uint4(0u, 0u, 0u, 0u);
  if (g == 0) {
    const long long n = v / sp, s = v - n * sp;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = (e < C) ? __ldg(x + (n * C + e) * sp + s) : 0.f;
    pk = make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
  }
  *reinterpret_cast<uint4*>(y + v * c_pitch + c_off + g * 8) = pk;
}

__global__ void unpack_cl_kernel(const __nv_bfloat16* __restrict__ x, int C, long long sp, long long voxels,
                                 int c_pitch, int c_off, float* __restrict__ y, int groups) {
  pdl_grid_sync();
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= voxels * groups) return;
  const int g = static_cast<int>(gid / voxels);
  const long long v = gid - static_cast<long long>(g) * voxels;
  const long long n = v / sp, s = v - n * sp;
  const uint4 pk = *reinterpret_cast<const uint4*>(x + v * c_pitch + c_off + g * 8);
  const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
  for (int e2 = 0; e2 < 4; ++e2) {
    const float2 f = unpack2(w[e2]);
    const int c = g * 8 + 2 * e2;
    if (c < C) y[(n * C + c) * sp + s] = f.x;
    if (c + 1 < C) y[(n * C + c + 1) * sp + s] = f.y;
  }
}

// Wide tensors (C > 8): a 64-voxel x 64-channel tile goes through shared memory so that BOTH sides are coalesced —
// the fp32 NCDHW side as 256-byte runs of one channel, the channels-last side as whole 128-byte voxel rows.  The
// direct mapping above writes 16 bytes out of every 128 per warp store (32 % of the copy bandwidth).  Tile rows are
// 33 words apart: conflict-free for the channel-pair writes (bank = v + pair) and for the 16-byte row reads
// (bank = v + 4*chunk + k).
constexpr int PK_VOX = 64, PK_PITCH = 33;
__global__ void __launch_bounds__(256)
pack_cl_tiled_kernel(const float* __restrict__ x, int C, long long sp, long long voxels, __nv_bfloat16* __restrict__ y,
                     int c_pitch, int c_off, int groups) {
  pdl_grid_sync();
  __shared__ uint32_t tile[PK_VOX * PK_PITCH];
  const int cb = blockIdx.y * 64;                      // first channel of this block's 64-channel slab
  const long long v0 = static_cast<long long>(blockIdx.x) * PK_VOX;
  {
    const int vl = threadIdx.x & 63, p = threadIdx.x >> 6;
    const long long v = v0 + vl;
    if (v < voxels) {
      const long long n = v / sp, s = v - n * sp;
      const float* src = x + (n * C) * sp + s;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = cb + (p * 8 + j) * 2;
        const float f0 = (c < C) ? __ldg(src + static_cast<long long>(c) * sp) : 0.f;
        const float f1 = (c + 1 < C) ? __ldg(src + static_cast<long long>(c + 1) * sp) : 0.f;
        tile[vl * PK_PITCH + p * 8 + j] = pack2(f0, f1);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int idx = threadIdx.x + k * 256;
    const int vl = idx >> 3, ch = idx & 7;
    const long long v = v0 + vl;
    if (v < voxels && (cb >> 3) + ch < groups) {
      const uint32_t* t = tile + vl * PK_PITCH + ch * 4;
      *reinterpret_cast<uint4*>(y + v * c_pitch + c_off + cb + ch * 8) = make_uint4(t[0], t[1], t[2], t[3]);
    }
  }
}

__global__ void __launch_bounds__(256)
unpack_cl_tiled_kernel(const __nv_bfloat16* __restrict__ x, int C, long long sp, long long voxels, int c_pitch, int c_off,
                       float* __restrict__ y, int groups) {
  pdl_grid_sync();
  __shared__ uint32_t tile[PK_VOX * PK_PITCH];
  const int cb = blockIdx.y * 64;
  const long long v0 = static_cast<long long>(blockIdx.x) * PK_VOX;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int idx = threadIdx.x + k * 256;
    const int vl = idx >> 3, ch = idx & 7;
    const long long v = v0 + vl;
    if (v < voxels && (cb >> 3) + ch < groups) {
      const uint4 pk = *reinterpret_cast<const uint4*>(x + v * c_pitch + c_off + cb + ch * 8);
      uint32_t* t = tile + vl * PK_PITCH + ch * 4;
      t[0] = pk.x; t[1] = pk.y; t[2] = pk.z; t[3] = pk.w;
    }
  }
  __syncthreads();
  const int vl = threadIdx.x & 63, p = threadIdx.x >> 6;
  const long long v = v0 + vl;
  if (v >= voxels) return;
  const long long n = v / sp, s = v - n * sp;
  float* dst = y + (n * C) * sp + s;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cb + (p * 8 + j) * 2;
    if (c < C) {
      const float2 f = unpack2(tile[vl * PK_PITCH + p * 8 + j]);
      dst[static_cast<long long>(c) * sp] = f.x;
      if (c + 1 < C) dst[static_cast<long long>(c + 1) * sp] = f.y;
    }
  }
}

// ----------------------------------------------------------------------------------------------- linear resize
// Source index / weight rule of UpsampleTrilinear3D / ResizeBilinear (ATen rule; trilinear.py:222-233 KAT).
// IEEE fp32, no FMA contraction, identical on host (linear_tap_host) and device -> bit-exact tables.
struct Tap {
  int i0, i1;
  float l0, l1;
};
__host__ __device__ __forceinline__ Tap linear_tap(int o, int n_in, float scale, int align_corners) {
  float r;
#ifdef __CUDA_ARCH__
  if (align_corners) {
    r = __fmul_rn(scale, static_cast<float>(o));
  } else {
    r = __fsub_rn(__fmul_rn(scale, __fadd_rn(static_cast<float>(o), 0.5f)), 0.5f);
    r = fmaxf(r, 0.f);
  }
#else
  if (align_corners) {
    volatile float m = scale * static_cast<float>(o);
    r = m;
  } else {
    volatile float a = static_cast<float>(o) + 0.5f;
    volatile float m = scale * a;
    volatile float s = m - 0.5f;
    r = s < 0.f ? 0.f : s;
  }
#endif
  Tap t;
  t.i0 = static_cast<int>(r);
  if (t.i0 > n_in - 1) t.i0 = n_in - 1;
  t.i1 = t.i0 + (t.i0 < n_in - 1 ? 1 : 0);
#ifdef __CUDA_ARCH__
  t.l1 = __fsub_rn(r, static_cast<float>(t.i0));
  t.l0 = __fsub_rn(1.0f, t.l1);
#else
  volatile float l1 = r - static_cast<float>(t.i0);
  volatile float l0 = 1.0f - l1;
  t.l1 = l1;
  t.l0 = l0;
#endif
  return t;
}

__device__ __forceinline__ float lerp_rn(float l0, float a, float l1, float b) {
  return __fadd_rn(__fmul_rn(l0, a), __fmul_rn(l1, b));
}

__device__ __forceinline__ float trilerp(const float* __restrict__ xc, int Hi, int Wi, const Tap& tt, const Tap& th,
                                         const Tap& tw) {
  const long long pw = static_cast<long long>(Hi) * Wi;
  const float* p00 = xc + tt.i0 * pw + static_cast<long long>(th.i0) * Wi;
  const float* p01 = xc + tt.i0 * pw + static_cast<long long>(th.i1) * Wi;
  const float* p10 = xc + tt.i1 * pw + static_cast<long long>(th.i0) * Wi;
  const float* p11 = xc + tt.i1 * pw + static_cast<long long>(th.i1) * Wi;
  const float a00 = lerp_rn(tw.l0, __ldg(p00 + tw.i0), tw.l1, __ldg(p00 + tw.i1));
  const float a01 = lerp_rn(tw.l0, __ldg(p01 + tw.i0), tw.l1, __ldg(p01 + tw.i1));
  const float a10 = lerp_rn(tw.l0, __ldg(p10 + tw.i0), tw.l1, __ldg(p10 + tw.i1));
  const float a11 = lerp_rn(tw.l0, __ldg(p11 + tw.i0), tw.l1, __ldg(p11 + tw.i1));
  const float b0 = lerp_rn(th.l0, a00, th.l1, a01);
  const float b1 = lerp_rn(th.l0, a10, th.l1, a11);
  return lerp_rn(tt.l0, b0, tt.l1, b1);
}

struct ResizeGeom {
  int Ti, Hi, Wi, To, Ho, Wo;
  float st, sh, sw;
  int align;
};

__global__ void resize3d_fwd_kernel(const float* __restrict__ x, long long NC, ResizeGeom g, float* __restrict__ y) {
  pdl_grid_sync();
  const long long total = NC * g.To * g.Ho * g.Wo;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(idx % g.Wo);
    long long r = idx / g.Wo;
    const int h = static_cast<int>(r % g.Ho);
    r /= g.Ho;
    const int t = static_cast<int>(r % g.To);
    const long long nc = r / g.To;
    const Tap tt = linear_tap(t, g.Ti, g.st, g.align), th = linear_tap(h, g.Hi, g.sh, g.align),
              tw = linear_tap(w, g.Wi, g.sw, g.align);
    y[idx] = trilerp(x + nc * g.Ti * g.Hi * g.Wi, g.Hi, g.Wi, tt, th, tw);
  }
}

// adjoint as a gather: each input element visits the (few) output positions whose taps touch it.
__device__ __forceinline__ void cand_range(int i, int n_out, float scale, int align, int& lo, int& hi) {
  if (scale <= 0.f) {
    lo = 0;
    hi = n_out - 1;
    return;
  }
  const float inv = 1.0f / scale;
  const float off = align ? 0.f : 0.5f;
  lo = static_cast<int>(floorf((static_cast<float>(i) - 1.f + off) * inv - off)) - 1;
  hi = static_cast<int>(ceilf((static_cast<float>(i) + 1.f + off) * inv - off)) + 1;
  if (lo < 0) lo = 0;
  if (hi > n_out - 1) hi = n_out - 1;
}

__global__ void resize3d_bwd_kernel(const float* __restrict__ gy, long long NC, ResizeGeom g,
                                    float* __restrict__ gx) {
  pdl_grid_sync();
  const long long total = NC * g.Ti * g.Hi * g.Wi;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(idx % g.Wi);
    long long r = idx / g.Wi;
    const int h = static_cast<int>(r % g.Hi);
    r /= g.Hi;
    const int t = static_cast<int>(r % g.Ti);
    const long long nc = r / g.Ti;
    int tlo, thi, hlo, hhi, wlo, whi;
    cand_range(t, g.To, g.st, g.align, tlo, thi);
    cand_range(h, g.Ho, g.sh, g.align, hlo, hhi);
    cand_range(w, g.Wo, g.sw, g.align, wlo, whi);
    const float* gyc = gy + nc * g.To * g.Ho * g.Wo;
    float acc = 0.f;
    for (int ot = tlo; ot <= thi; ++ot) {
      const Tap a = linear_tap(ot, g.Ti, g.st, g.align);
      const float ct = (a.i0 == t ? a.l0 : 0.f) + (a.i1 == t ? a.l1 : 0.f);
      if (ct == 0.f) continue;
      for (int oh = hlo; oh <= hhi; ++oh) {
        const Tap b = linear_tap(oh, g.Hi, g.sh, g.align);
        const float chh = (b.i0 == h ? b.l0 : 0.f) + (b.i1 == h ? b.l1 : 0.f);
        if (chh == 0.f) continue;
        const float cth = ct * chh;
        const float* row = gyc + (static_cast<long long>(ot) * g.Ho + oh) * g.Wo;
        for (int ow = wlo; ow <= whi; ++ow) {
          const Tap c = linear_tap(ow, g.Wi, g.sw, g.align);
          const float cw = (c.i0 == w ? c.l0 : 0.f) + (c.i1 == w ? c.l1 : 0.f);
          if (cw != 0.f) acc = fmaf(cth * cw, __ldg(row + ow), acc);
        }
      }
    }
    gx[idx] = acc;
  }
}

__global__ void linear_taps_kernel(int n_in, int n_out, float scale, int align, int* i0, int* i1, float* l0,
                                   float* l1) {
  pdl_grid_sync();
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  const Tap t = linear_tap(o, n_in, scale, align);
  i0[o] = t.i0;
  i1[o] = t.i1;
  l0[o] = t.l0;
  l1[o] = t.l1;
}

// ----------------------------------------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u01(uint32_t x) { return (static_cast<float>(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }
// Box-Muller straight on the SFU (MUFU.LG2 / SQRT / SIN / COS: abs error ~1e-6, irrelevant for a noise source).  The
// IEEE sqrtf() alone costs ~10 instructions plus a divergent slow-path guard, and the fused block-input kernel is
// instruction-issue bound by exactly this code.
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  float lg, r, s, c;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u01(a)));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(lg * -1.3862943611198906f));   // sqrt(-2 ln u)
  const float ang = 6.283185307179586f * u01(b);
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(ang));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(ang));
  z0 = r * c;
  z1 = r * s;
}

// ----------------------------------------------------------------------------------------------- tiled resize
// Shared-memory-staged, separable versions of the three resize kernels (the generic ones above remain as fallback for
// shapes whose tile does not fit).  One CTA = one (n, c) slice x one band of rows x ALL frames.
//   forward : stage the W-interpolated source rows of the band once (Ti x n_hs x Wo floats), then every output voxel
//             is 4 shared-memory reads + an H and a T interpolation — same operation order as trilerp() => same bits.
//   backward: accumulate ct*ch*gy into b[ti][hs][wo] (frame planes owned by thread rows, columns by thread columns,
//             fixed loop order => deterministic, no atomics), then contract the W axis with an inverse tap table.
// Thread layout (nx, ny): nx = ceil(Wo / cpt) columns (cpt = columns per thread, so odd widths like 257 do not waste a
// whole pass), ny thread rows split the image rows.  Tap tables live in shared memory.
struct TiledGeom {
  ResizeGeom g;
  int band;       // output rows per CTA (forward) / source rows per CTA (backward); power of two
  int band_log2;
  int n_hs;       // forward: staged source rows per frame (upper bound)
  int nx, ny, cpt;
};
struct TapRow {   // one interpolation tap pair, pre-multiplied row offsets into the staged tile
  int o0, o1;
  float l0, l1;
};

__device__ __forceinline__ void build_w_taps(const ResizeGeom& g, int* tw_i0, float* tw_l1, int tid, int nthreads) {
  for (int wo = tid; wo < g.Wo; wo += nthreads) {
    const Tap t = linear_tap(wo, g.Wi, g.sw, g.align);
    tw_i0[wo] = t.i0;
    tw_l1[wo] = t.l1;
  }
}

// stage a[ti][r][wo] = W-lerp of source row (hs_lo + r) of frame ti, for r < n_rows.  Column-outer (the W tap of a
// column is read once), four rows per step with all eight loads issued before the arithmetic (memory-level parallelism).
__device__ __forceinline__ void stage_w_lerp(const float* __restrict__ xc, const TiledGeom& tg, int hs_lo, int n_rows,
                                             const int* tw_i0, const float* tw_l1, float* a) {
  const ResizeGeom& g = tg.g;
  const int rows = g.Ti * n_rows;
  const int frame_skip = (g.Hi - n_rows) * g.Wi;      // source offset jump between the last band row of a frame and
  for (int wo = threadIdx.x; wo < g.Wo; wo += tg.nx) {   // the first of the next
    const int i0 = tw_i0[wo];
    const int i1 = i0 + (i0 < g.Wi - 1 ? 1 : 0);
    const float l1 = tw_l1[wo];
    const float l0 = __fsub_rn(1.0f, l1);
    for (int r0 = threadIdx.y; r0 < rows; r0 += 4 * tg.ny) {
      float va[4], vb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u * tg.ny;
        if (r < rows) {
          const int ti = r / n_rows;
          const float* src = xc + static_cast<long long>(hs_lo + r) * g.Wi + static_cast<long long>(ti) * frame_skip;
          va[u] = __ldg(src + i0);
          vb[u] = __ldg(src + i1);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u * tg.ny;
        if (r < rows) a[static_cast<long long>(r) * g.Wo + wo] = lerp_rn(l0, va[u], l1, vb[u]);
      }
    }
  }
}

// H stage: hb[ti][hl][wo] = lerp_h(a[ti][h0(hl)][wo], a[ti][h1(hl)][wo]) for the rows of the band
__device__ __forceinline__ void interp_h(const TiledGeom& tg, int band_rows, int n_rows, const float* a,
                                         const TapRow* htab, float* hb) {
  const ResizeGeom& g = tg.g;
  const int rows = g.Ti * band_rows;
  for (int r = threadIdx.y; r < rows; r += tg.ny) {
    const int ti = r / band_rows, hl = r - ti * band_rows;
    const TapRow th = htab[hl];
    const float* r0 = a + ti * n_rows * g.Wo + th.o0;
    const float* r1 = a + ti * n_rows * g.Wo + th.o1;
    float* dst = hb + r * g.Wo;
    for (int wo = threadIdx.x; wo < g.Wo; wo += tg.nx) dst[wo] = lerp_rn(th.l0, r0[wo], th.l1, r1[wo]);
  }
}

// T stage: the band's rows are contiguous both in hb (per frame) and in y (per frame): a pure streaming loop
__device__ __forceinline__ void interp_t(const TiledGeom& tg, int band_rows, int ho0, const float* hb,
                                         const TapRow* ttab, float* __restrict__ yc, int tid, int nth) {
  const ResizeGeom& g = tg.g;
  const int n = band_rows * g.Wo, fs = band_rows * g.Wo;
  float* dst = yc + static_cast<long long>(ho0) * g.Wo;
  const long long plane = static_cast<long long>(g.Ho) * g.Wo;
  for (int to = 0; to < g.To; ++to, dst += plane) {
    const TapRow tt = ttab[to];                 // o0 / o1 = source frame indices here
    const float* f0 = hb + tt.o0 * fs;
    const float* f1 = hb + tt.o1 * fs;
    for (int i = tid; i < n; i += nth) dst[i] = lerp_rn(tt.l0, f0[i], tt.l1, f1[i]);
  }
}

struct FwdTile {
  float* a;
  float* hb;
  int* tw_i0;
  float* tw_l1;
  TapRow* ttab;
  TapRow* htab;
  int ho0, ho1, hs_lo, n_rows;
};
__device__ __forceinline__ FwdTile carve_fwd(float* sm, const TiledGeom& tg) {
  const ResizeGeom& g = tg.g;
  FwdTile f;
  f.ttab = reinterpret_cast<TapRow*>(sm);                               // 16-byte aligned first
  f.htab = f.ttab + g.To;
  f.a = reinterpret_cast<float*>(f.htab + tg.band);
  f.hb = f.a + g.Ti * tg.n_hs * g.Wo;
  f.tw_i0 = reinterpret_cast<int*>(f.hb + g.Ti * tg.band * g.Wo);
  f.tw_l1 = reinterpret_cast<float*>(f.tw_i0 + g.Wo);
  f.ho0 = blockIdx.x * tg.band;
  f.ho1 = min(f.ho0 + tg.band, g.Ho);
  f.hs_lo = linear_tap(f.ho0, g.Hi, g.sh, g.align).i0;
  f.n_rows = linear_tap(f.ho1 - 1, g.Hi, g.sh, g.align).i1 - f.hs_lo + 1;   // <= tg.n_hs by construction
  return f;
}

// T / H tap tables of one band: ttab = source frame indices, htab = row offsets (floats) inside one staged frame
__device__ __forceinline__ void build_th_taps(const TiledGeom& tg, const FwdTile& f, int tid, int nthreads) {
  const ResizeGeom& g = tg.g;
  for (int i = tid; i < g.To + (f.ho1 - f.ho0); i += nthreads) {
    if (i < g.To) {
      const Tap t = linear_tap(i, g.Ti, g.st, g.align);
      f.ttab[i] = TapRow{t.i0, t.i1, t.l0, t.l1};
    } else {
      const int hl = i - g.To;
      const Tap t = linear_tap(f.ho0 + hl, g.Hi, g.sh, g.align);
      f.htab[hl] = TapRow{(t.i0 - f.hs_lo) * g.Wo, (t.i1 - f.hs_lo) * g.Wo, t.l0, t.l1};
    }
  }
}

__global__ void resize3d_fwd_tiled_kernel(const float* __restrict__ x, const TiledGeom tg, float* __restrict__ y) {
  pdl_grid_sync();
  extern __shared__ __align__(16) float rs_sm[];
  const ResizeGeom& g = tg.g;
  const FwdTile f = carve_fwd(rs_sm, tg);
  const int tid = threadIdx.y * tg.nx + threadIdx.x, nth = tg.nx * tg.ny;
  const long long nc = blockIdx.y;
  build_w_taps(g, f.tw_i0, f.tw_l1, tid, nth);
  build_th_taps(tg, f, tid, nth);
  __syncthreads();
  stage_w_lerp(x + nc * g.Ti * g.Hi * g.Wi, tg, f.hs_lo, f.n_rows, f.tw_i0, f.tw_l1, f.a);
  __syncthreads();
  interp_h(tg, f.ho1 - f.ho0, f.n_rows, f.a, f.htab, f.hb);
  __syncthreads();
  interp_t(tg, f.ho1 - f.ho0, f.ho0, f.hb, f.ttab, y + nc * g.To * g.Ho * g.Wo, tid, nth);
}

// fused block input stage, tiled: per channel W / H / T stages -> up ; then x_in = bf16(up + noise*amp) for the voxels
// this thread just wrote (same thread <-> same voxel in the T stage and in the pack loop).
__global__ void upsample_noise_pack_tiled_kernel(const float* __restrict__ x, int C, const TiledGeom tg,
                                                 const float* __restrict__ noise, float amp, unsigned long long seed,
                                                 unsigned long long sample_base,
                                                 const unsigned long long* __restrict__ d_sample_offset,
                                                 float* __restrict__ up, __nv_bfloat16* __restrict__ xin, int xin_f32) {
  pdl_grid_sync();
  extern __shared__ __align__(16) float rs_sm[];
  const ResizeGeom& g = tg.g;
  if (d_sample_offset) sample_base += *d_sample_offset;
  const FwdTile f = carve_fwd(rs_sm, tg);
  const int tid = threadIdx.y * tg.nx + threadIdx.x, nth = tg.nx * tg.ny;
  const long long n = blockIdx.y;
  const long long spo = static_cast<long long>(g.To) * g.Ho * g.Wo;
  const long long spi = static_cast<long long>(g.Ti) * g.Hi * g.Wi;
  const int band_rows = f.ho1 - f.ho0;
  build_w_taps(g, f.tw_i0, f.tw_l1, tid, nth);
  build_th_taps(tg, f, tid, nth);
  for (int c = 0; c < C; ++c) {
    __syncthreads();
    stage_w_lerp(x + (n * C + c) * spi, tg, f.hs_lo, f.n_rows, f.tw_i0, f.tw_l1, f.a);
    __syncthreads();
    interp_h(tg, band_rows, f.n_rows, f.a, f.htab, f.hb);
    __syncthreads();
    interp_t(tg, band_rows, f.ho0, f.hb, f.ttab, up + (n * C + c) * spo, tid, nth);
  }
  // pack: x_in[n][to][ho][wo][0..7] = bf16(up + noise*amp), zero padded to 8 channels.  SAME (frame, element) -> thread
  // mapping as interp_t(), so every thread re-reads only values it stored itself.
  const unsigned long long sample = sample_base + static_cast<unsigned long long>(n);
  const int nel = band_rows * g.Wo;
  const long long plane = static_cast<long long>(g.Ho) * g.Wo;
  for (int to = 0; to < g.To; ++to) {
    const long long base = to * plane + static_cast<long long>(f.ho0) * g.Wo;
    for (int i = tid; i < nel; i += nth) {
      const long long sidx = base + i;
      float z[4] = {0.f, 0.f, 0.f, 0.f};
      if (!noise && seed != 0ull) {
        const uint4 rnd = philox4x32_10(make_uint4(static_cast<uint32_t>(sidx), static_cast<uint32_t>(sidx >> 32),
                                                   static_cast<uint32_t>(sample), static_cast<uint32_t>(sample >> 32)),
                                        make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
        box_muller(rnd.x, rnd.y, z[0], z[1]);
        box_muller(rnd.z, rnd.w, z[2], z[3]);
      }
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int c = 0; c < C; ++c) {
        const long long o = (n * C + c) * spo + sidx;
        const float u = up[o];
        const float nz = noise ? noise[o] : z[c];
        v[c] = fmaf(nz, amp, u);
      }
      *reinterpret_cast<uint4*>(xin + (n * spo + sidx) * 8) = xin_voxel(v, xin_f32);
    }
  }
}

// ----------------------------------------------------------------------------------------------- column-walk resize
// Register-only forward for Ti <= 8 (every pyramid level): a thread owns one output column (ho, wo), loads the 4 source
// values of every source frame up front (4*Ti independent loads, served by L1/L2 — neighbouring threads share them),
// forms the W- then H-interpolated value hv[ti] of each frame ONCE, and walks along T_out writing ~To/Ti outputs per
// frame.  Same W -> H -> T operation order as trilerp() => identical bits.  No shared-memory tile and no barrier.
constexpr int CW_MAX_TI = 8;
constexpr int CW_THREADS = 256;

// T-axis walk of one launch, built on the HOST with linear_tap() (bit-identical to the device function, §3.4 of
// DESIGN.md) and passed as a kernel parameter: the constant bank serves block-uniform reads for free, so the kernels
// carry no shared-memory tap table, no barrier and no per-output tap arithmetic.
constexpr int CW_MAX_PER_FRAME = 4;
struct TWalk {
  // outputs whose lower tap is source frame f, in order: lambda pairs at COMPILE-TIME constant-bank offsets (the
  // unrolled walk reads them as immediate c[][] operands — no table load, no loop counter, no index arithmetic)
  float l0[CW_MAX_TI][CW_MAX_PER_FRAME], l1[CW_MAX_TI][CW_MAX_PER_FRAME];
  int n[CW_MAX_TI];            // how many outputs frame f feeds as the lower tap (<= CW_MAX_PER_FRAME)
  unsigned col_magic;          // ceil(2^32 / Wo): ho = umulhi(col, magic), exact while Ho*Wo*Wo < 2^32 (0: Wo == 1)
};

template <int C, int TI>
__device__ __forceinline__ void colwalk_load(const float* __restrict__ xs /* block-uniform slice base */, unsigned spi,
                                             const ResizeGeom& g, const Tap& th, const Tap& tw,
                                             float (&hv)[C][TI + 1]) {
  // two row pointers per channel; the W neighbour is the +4-byte immediate of the same address register (the last
  // column, where i1 == i0, re-uses its own value instead), frames are one multiply-add-wide away: ~1 integer
  // instruction per load instead of 4
  const unsigned fsz4 = g.Hi * g.Wi * 4u;   // frame stride in bytes (< 2^31, checked by the launcher)
  const unsigned dw = (tw.i1 - tw.i0) * 4u;  // 0 on the last column (i1 == i0), else 4 bytes
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const char* p00 = reinterpret_cast<const char*>(xs + (c * spi + th.i0 * g.Wi + tw.i0));
    const char* p10 = reinterpret_cast<const char*>(xs + (c * spi + th.i1 * g.Wi + tw.i0));
    const char* p01 = p00 + dw;
    const char* p11 = p10 + dw;
    float v[TI][4];
#pragma unroll
    for (int f = 0; f < TI; ++f) {   // all 4*TI loads are independent: issued back to back, one round trip
      const unsigned long long off = static_cast<unsigned long long>(f) * fsz4;
      v[f][0] = __ldg(reinterpret_cast<const float*>(p00 + off));
      v[f][1] = __ldg(reinterpret_cast<const float*>(p01 + off));
      v[f][2] = __ldg(reinterpret_cast<const float*>(p10 + off));
      v[f][3] = __ldg(reinterpret_cast<const float*>(p11 + off));
    }
#pragma unroll
    for (int f = 0; f < TI; ++f) {
      const float a0 = lerp_rn(tw.l0, v[f][0], tw.l1, v[f][1]);
      const float a1 = lerp_rn(tw.l0, v[f][2], tw.l1, v[f][3]);
      hv[c][f] = lerp_rn(th.l0, a0, th.l1, a1);
    }
    hv[c][TI] = hv[c][TI - 1];   // i1 == i0 on the last source frame
  }
}

__device__ __forceinline__ void colwalk_split(unsigned col, const ResizeGeom& g, const TWalk& w, unsigned& ho,
                                              unsigned& wo) {
  ho = w.col_magic ? __umulhi(col, w.col_magic) : col;
  wo = col - ho * g.Wo;
}

template <int TI>
__global__ void __launch_bounds__(CW_THREADS)
resize3d_fwd_colwalk_kernel(const float* __restrict__ x, const ResizeGeom g, const __grid_constant__ TWalk w,
                            float* __restrict__ y) {
  pdl_grid_sync();
  const unsigned plane = g.Ho * g.Wo;
  const unsigned col = blockIdx.x * CW_THREADS + threadIdx.x;
  if (col >= plane) return;
  unsigned ho, wo;
  colwalk_split(col, g, w, ho, wo);
  const long long nc = blockIdx.y;
  const Tap th = linear_tap(ho, g.Hi, g.sh, g.align), tw = linear_tap(wo, g.Wi, g.sw, g.align);
  const unsigned spi = TI * g.Hi * g.Wi;
  float hv[1][TI + 1];
  colwalk_load<1, TI>(x + nc * spi, spi, g, th, tw, hv);
  char* __restrict__ dst = reinterpret_cast<char*>(y + (nc * g.To * plane + col));
  const unsigned plane4 = plane * 4u;
  unsigned to = 0;                       // block-uniform: lives in a uniform register
#pragma unroll
  for (int f = 0; f < TI; ++f) {
    const int n = w.n[f];
    const float a = hv[0][f], b = hv[0][f + 1];
#pragma unroll
    for (int k = 0; k < CW_MAX_PER_FRAME; ++k) {
      if (k >= n) break;
      *reinterpret_cast<float*>(dst + static_cast<unsigned long long>(to) * plane4) =
          lerp_rn(w.l0[f][k], a, w.l1[f][k], b);
      ++to;
    }
  }
}

template <int C, int TI>
__global__ void __launch_bounds__(CW_THREADS)
upsample_noise_pack_colwalk_kernel(const float* __restrict__ x, const ResizeGeom g, const __grid_constant__ TWalk w,
                                   const float* __restrict__ noise, float amp, unsigned long long seed,
                                   unsigned long long sample_base,
                                   const unsigned long long* __restrict__ d_sample_offset, float* __restrict__ up,
                                   __nv_bfloat16* __restrict__ xin, int xin_f32) {
  pdl_grid_sync();
  const unsigned plane = g.Ho * g.Wo;
  const unsigned col = blockIdx.x * CW_THREADS + threadIdx.x;
  if (col >= plane) return;
  if (d_sample_offset) sample_base += *d_sample_offset;
  unsigned ho, wo;
  colwalk_split(col, g, w, ho, wo);
  const long long n = blockIdx.y;
  const Tap th = linear_tap(ho, g.Hi, g.sh, g.align), tw = linear_tap(wo, g.Wi, g.sw, g.align);
  const unsigned spi = TI * g.Hi * g.Wi;
  const unsigned spo = g.To * plane;                 // C * spo < 2^31 (checked by the launcher)
  float hv[C][TI + 1];
  colwalk_load<C, TI>(x + n * C * static_cast<long long>(spi), spi, g, th, tw, hv);
  const unsigned long long sample = sample_base + static_cast<unsigned long long>(n);
  const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  const bool draw = !noise && seed != 0ull;
  unsigned sidx = col;                                 // spatial index inside the sample
  float* __restrict__ up_p = up + (n * C * spo + col);
  const float* __restrict__ nz_p = noise ? noise + (n * C * spo + col) : nullptr;
  uint4* __restrict__ xin_p = reinterpret_cast<uint4*>(xin) + (n * spo + col);
#pragma unroll
  for (int f = 0; f < TI; ++f) {
    const int nf = w.n[f];
#pragma unroll
    for (int k = 0; k < CW_MAX_PER_FRAME; ++k) {
      if (k >= nf) break;
      const float l0 = w.l0[f][k], l1 = w.l1[f][k];
      float z[4] = {0.f, 0.f, 0.f, 0.f};
      if (draw) {
        const uint4 rnd = philox4x32_10(make_uint4(sidx, 0u, static_cast<uint32_t>(sample),
                                                   static_cast<uint32_t>(sample >> 32)), key);
        box_muller(rnd.x, rnd.y, z[0], z[1]);
        box_muller(rnd.z, rnd.w, z[2], z[3]);
      } else if (nz_p) {
#pragma unroll
        for (int c = 0; c < C; ++c) z[c] = __ldg(nz_p + c * spo);
        nz_p += plane;
      }
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float u = lerp_rn(l0, hv[c][f], l1, hv[c][f + 1]);
        up_p[c * spo] = u;
        v[c] = fmaf(z[c], amp, u);
      }
      *xin_p = xin_voxel(v, xin_f32);
      up_p += plane;
      xin_p += plane;
      sidx += plane;
    }
  }
}

// Backward (adjoint) for moderate up-sampling (at most 4 output rows / columns tap one source row / column, i.e. ratio
// <= 1.5 in H and W — every pyramid level).  InvTap: for one source row (column) the run of output rows (columns) that
// tap it and their coefficients — fixed along T.  The kernel that uses them (resize3d_bwd_colwalk_kernel) is below.
struct InvTap {
  int lo, n;        // contributing outputs [lo, lo+n)
  float c[4];       // their coefficients
};
__host__ __device__ __forceinline__ int first_output_at_or_above(int i, int n_in, int n_out, float scale, int align) {
  // min{o : i0(o) >= i}; i0 is non-decreasing in o
  if (i <= 0) return 0;
  if (scale <= 0.f) return n_out;
  const float off = align ? 0.f : 0.5f;
  int o = static_cast<int>(floorf((static_cast<float>(i) + off) / scale - off)) - 1;
  if (o < 0) o = 0;
  while (o < n_out && linear_tap(o, n_in, scale, align).i0 < i) ++o;
  return o;
}
__host__ __device__ __forceinline__ InvTap make_inv_tap(int i, int n_in, int n_out, float scale, int align) {
  InvTap e;
  e.lo = first_output_at_or_above(i - 1, n_in, n_out, scale, align);
  const int hi = (i + 1 > n_in - 1 + 1) ? n_out : first_output_at_or_above(i + 1, n_in, n_out, scale, align);
  e.n = hi - e.lo;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int k = 0; k < 4; ++k) {
    e.c[k] = 0.f;
    if (k < e.n) {
      const Tap t = linear_tap(e.lo + k, n_in, scale, align);
      e.c[k] = (t.i0 == i ? t.l0 : 0.f) + (t.i1 == i ? t.l1 : 0.f);
    }
  }
  return e;
}

constexpr int BW_TX = 32, BW_TY = 8;
// Adjoint in two register-resident stages per block (source tile 32 x 8, all frames):
//   1. T first: every gy column of the tile's window (rows x cols, halo included) is walked along T_out by one thread —
//      coalesced, all of a column's loads issued back to back (predicated, no control flow) — and reduced straight to
//      the TI source frames with the constant-bank walk; the TI values go to shared memory [TI][rows][cols].
//   2. H / W: each source voxel's thread gathers its <= NH x NW window per source frame from shared memory.
// Reducing T before H / W runs the 3 x 3 gather TI (7) instead of To (13) times and never stages raw gy in shared
// memory; fixed summation order => deterministic, no atomics.
struct BwdWin {
  int rows, cols;     // window extent (max over blocks, computed on the host with the same tap functions)
};
template <int NH, int NW, int TI, int NK>   // bounds: contributing rows / columns, source frames, outputs per frame
__global__ void __launch_bounds__(BW_TX * BW_TY)
resize3d_bwd_colwalk_kernel(const float* __restrict__ gy, const ResizeGeom g, const __grid_constant__ TWalk w,
                            const BwdWin win, float* __restrict__ gx) {
  pdl_grid_sync();
  extern __shared__ __align__(16) float bw_sm[];          // [TI][win.rows][win.cols]
  __shared__ InvTap htab[BW_TY], wtab[BW_TX];
  const int tid = threadIdx.y * BW_TX + threadIdx.x;
  const int hs0 = blockIdx.y * BW_TY, wi0 = blockIdx.x * BW_TX;
  if (tid < BW_TY) {
    htab[tid] = make_inv_tap(min(hs0 + tid, g.Hi - 1), g.Hi, g.Ho, g.sh, g.align);
  } else if (tid >= 32 && tid < 32 + BW_TX) {
    const int j = tid - 32;
    wtab[j] = make_inv_tap(min(wi0 + j, g.Wi - 1), g.Wi, g.Wo, g.sw, g.align);
  }
  __syncthreads();
  const long long nc = blockIdx.z;
  const unsigned plane_o = g.Ho * g.Wo, plane_i = g.Hi * g.Wi;
  const int ho_lo = htab[0].lo, wo_lo = wtab[0].lo;
  const int rows = min(htab[BW_TY - 1].lo + htab[BW_TY - 1].n, g.Ho) - ho_lo;      // <= win.rows
  const int cols = min(wtab[BW_TX - 1].lo + wtab[BW_TX - 1].n, g.Wo) - wo_lo;      // <= win.cols
  const int fstride = win.rows * win.cols;
  {
    const char* src = reinterpret_cast<const char*>(gy + (nc * g.To * plane_o + ho_lo * g.Wo + wo_lo));
    const unsigned plane4 = plane_o * 4u;
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < rows; r += BW_TY) {
      for (int c = lane; c < cols; c += 32) {
        const char* p = src + static_cast<unsigned>(r * g.Wo + c) * 4u;
        float gv[TI][NK];
        unsigned to = 0;                                   // block-uniform
#pragma unroll
        for (int f = 0; f < TI; ++f) {
          const int n = w.n[f];
#pragma unroll
          for (int k = 0; k < NK; ++k)
            gv[f][k] = (k < n) ? __ldg(reinterpret_cast<const float*>(p + static_cast<unsigned long long>(to + k) * plane4))
                               : 0.f;
          to += n;
        }
        float acc[TI];
#pragma unroll
        for (int f = 0; f < TI; ++f) acc[f] = 0.f;
#pragma unroll
        for (int f = 0; f < TI; ++f) {
#pragma unroll
          for (int k = 0; k < NK; ++k) {                   // lambdas past n[f] are zero (make_twalk)
            acc[f] = fmaf(w.l0[f][k], gv[f][k], acc[f]);
            if (f == TI - 1) acc[f] = fmaf(w.l1[f][k], gv[f][k], acc[f]);
            else acc[f + 1] = fmaf(w.l1[f][k], gv[f][k], acc[f + 1]);
          }
        }
        float* d = bw_sm + r * win.cols + c;
#pragma unroll
        for (int f = 0; f < TI; ++f) d[f * fstride] = acc[f];
      }
    }
  }
  __syncthreads();
  const int hs = hs0 + threadIdx.y, wi = wi0 + threadIdx.x;
  if (hs >= g.Hi || wi >= g.Wi) return;
  const InvTap eh = htab[threadIdx.y], ew = wtab[threadIdx.x];
  const float* wp = bw_sm + (eh.lo - ho_lo) * win.cols + (ew.lo - wo_lo);
  int off[NH][NW];    // positions past the contributing range alias (0, .) / (., 0) and meet a zero coefficient
#pragma unroll
  for (int k = 0; k < NH; ++k)
#pragma unroll
    for (int j = 0; j < NW; ++j) off[k][j] = (k < eh.n ? k : 0) * win.cols + (j < ew.n ? j : 0);
  float* dst = gx + (nc * TI * plane_i + hs * g.Wi + wi);
#pragma unroll
  for (int f = 0; f < TI; ++f) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NH; ++k) {
      float r = 0.f;
#pragma unroll
      for (int j = 0; j < NW; ++j) r = fmaf(ew.c[j], wp[off[k][j]], r);
      s = fmaf(eh.c[k], r, s);
    }
    wp += fstride;
    dst[static_cast<size_t>(f) * plane_i] = s;
  }
}

// backward: a thread owns (source row hs, column wo) and walks along T_out with register accumulators for the (at
// most two) source frames currently being fed — no shared-memory read-modify-write, fixed order => deterministic.
__global__ void resize3d_bwd_tiled_kernel(const float* __restrict__ gy, const TiledGeom tg, float* __restrict__ gx) {
  pdl_grid_sync();
  extern __shared__ __align__(16) float rs_sm[];
  const ResizeGeom& g = tg.g;
  const int hs0 = blockIdx.x * tg.band;
  const int hs1 = min(hs0 + tg.band, g.Hi);
  const int n_rows = hs1 - hs0;
  TapRow* htab = reinterpret_cast<TapRow*>(rs_sm);                     // [Ho]: (i0, i1, l0, l1) of every output row
  TapRow* ttab = htab + g.Ho;                                          // [To]
  float* b = reinterpret_cast<float*>(ttab + g.To);                    // [Ti][band][Wo]
  int* tw_i0 = reinterpret_cast<int*>(b + static_cast<long long>(g.Ti) * tg.band * g.Wo);
  float* tw_l1 = reinterpret_cast<float*>(tw_i0 + g.Wo);
  int* first = reinterpret_cast<int*>(tw_l1 + g.Wo);                   // [Wi + 2]: first[i] = min{wo : i0[wo] >= i}
  int* hfirst = first + g.Wi + 2;                                      // [Hi + 2]: first[i] = min{ho : i0[ho] >= i}
  const int tid = threadIdx.y * tg.nx + threadIdx.x, nth = tg.nx * tg.ny;
  const long long nc = blockIdx.y;
  build_w_taps(g, tw_i0, tw_l1, tid, nth);
  for (int i = tid; i < g.Ho + g.To; i += nth) {
    if (i < g.Ho) {
      const Tap t = linear_tap(i, g.Hi, g.sh, g.align);
      htab[i] = TapRow{t.i0, t.i1, t.l0, t.l1};
    } else {
      const Tap t = linear_tap(i - g.Ho, g.Ti, g.st, g.align);
      ttab[i - g.Ho] = TapRow{t.i0, t.i1, t.l0, t.l1};
    }
  }
  __syncthreads();
  for (int wo = tid; wo < g.Wo; wo += nth) {                           // inverse W table
    const int cur = tw_i0[wo];
    const int lo = (wo == 0) ? 0 : tw_i0[wo - 1] + 1;
    for (int i = lo; i <= cur; ++i) first[i] = wo;
    if (wo == g.Wo - 1)
      for (int i = cur + 1; i <= g.Wi + 1; ++i) first[i] = g.Wo;
  }
  for (int ho = tid; ho < g.Ho; ho += nth) {                           // inverse H table
    const int cur = htab[ho].o0;
    const int lo = (ho == 0) ? 0 : htab[ho - 1].o0 + 1;
    for (int i = lo; i <= cur; ++i) hfirst[i] = ho;
    if (ho == g.Ho - 1)
      for (int i = cur + 1; i <= g.Hi + 1; ++i) hfirst[i] = g.Ho;
  }
  __syncthreads();
  const float* gyc = gy + nc * g.To * g.Ho * g.Wo;
  const long long plane = static_cast<long long>(g.Ho) * g.Wo;
  const long long fs = static_cast<long long>(tg.band) * g.Wo;         // frame stride inside b
  // stage 1: b[ti][hs][wo] = sum_to ct(to -> ti) * sum_ho ch(ho -> hs) * gy[to][ho][wo]
  for (int hl = threadIdx.y; hl < n_rows; hl += tg.ny) {
    const int hs = hs0 + hl;
    const int h_lo = hfirst[hs > 0 ? hs - 1 : 0], h_hi = hfirst[hs + 1];   // candidates: i0(ho) in {hs-1, hs}
    for (int wo = threadIdx.x; wo < g.Wo; wo += tg.nx) {
      float* bcol = b + static_cast<long long>(hl) * g.Wo + wo;
      int f_lo = 0;
      float acc_lo = 0.f, acc_hi = 0.f;
      const float* src = gyc + wo + static_cast<long long>(h_lo) * g.Wo;
      const int nh = h_hi - h_lo;
      float chv[4] = {0.f, 0.f, 0.f, 0.f};   // H coefficients of the candidate rows (constant along T)
      if (nh <= 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < nh) {
            const TapRow th = htab[h_lo + k];
            chv[k] = (th.o0 == hs ? th.l0 : 0.f) + (th.o1 == hs ? th.l1 : 0.f);
          }
      }
      float cur[4] = {0.f, 0.f, 0.f, 0.f};
      if (nh <= 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < nh) cur[k] = __ldg(src + static_cast<long long>(k) * g.Wo);
      }
      for (int to = 0; to < g.To; ++to, src += plane) {
        float sh = 0.f;
        if (nh <= 4) {
          float nxt[4] = {0.f, 0.f, 0.f, 0.f};
          if (to + 1 < g.To) {       // issue the next frame's loads before consuming this frame's
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < nh) nxt[k] = __ldg(src + plane + static_cast<long long>(k) * g.Wo);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) sh = fmaf(chv[k], cur[k], sh);
#pragma unroll
          for (int k = 0; k < 4; ++k) cur[k] = nxt[k];
        } else {
          for (int k = 0; k < nh; ++k) {
            const TapRow th = htab[h_lo + k];
            const float ch = (th.o0 == hs ? th.l0 : 0.f) + (th.o1 == hs ? th.l1 : 0.f);
            sh = fmaf(ch, __ldg(src + static_cast<long long>(k) * g.Wo), sh);
          }
        }
        const TapRow tt = ttab[to];
        while (f_lo < tt.o0) {        // frames below i0(to) are complete (i0 is non-decreasing in `to`)
          bcol[f_lo * fs] = acc_lo;
          acc_lo = acc_hi;
          acc_hi = 0.f;
          ++f_lo;
        }
        acc_lo = fmaf(tt.l0, sh, acc_lo);
        if (tt.o1 == f_lo) acc_lo = fmaf(tt.l1, sh, acc_lo);
        else acc_hi = fmaf(tt.l1, sh, acc_hi);
      }
      for (; f_lo < g.Ti; ++f_lo) {
        bcol[f_lo * fs] = acc_lo;
        acc_lo = acc_hi;
        acc_hi = 0.f;
      }
    }
  }
  __syncthreads();
  // stage 2: gx[ti][hs][wi] = sum_wo cw(wo -> wi) * b[ti][hs][wo]
  float* gxc = gx + nc * g.Ti * g.Hi * g.Wi;
  for (int r = threadIdx.y; r < g.Ti * n_rows; r += tg.ny) {
    const int ti = r / n_rows, hl = r - ti * n_rows;
    const float* brow = b + (static_cast<long long>(ti) * tg.band + hl) * g.Wo;
    float* dst = gxc + (static_cast<long long>(ti) * g.Hi + (hs0 + hl)) * g.Wi;
    for (int wi = threadIdx.x; wi < g.Wi; wi += tg.nx) {
      const int lo = first[wi > 0 ? wi - 1 : 0], hi = first[wi + 1];
      float acc = 0.f;
      for (int wo = lo; wo < hi; ++wo) {
        const int i0 = tw_i0[wo];
        const int i1 = i0 + (i0 < g.Wi - 1 ? 1 : 0);
        const float l1 = tw_l1[wo];
        const float cw = (i0 == wi ? __fsub_rn(1.0f, l1) : 0.f) + (i1 == wi ? l1 : 0.f);
        acc = fmaf(cw, brow[wo], acc);
      }
      dst[wi] = acc;
    }
  }
}

// ----------------------------------------------------------------------------------------------- clip from frames
// Device-side restatement of the reference's data path between the decoder and the network:
// generate_frames.py:42-46 (cv2.resize(rgb, INTER_LINEAR) on uint8) -> video.py:52-59 (frames[idx : idx+lcm+1 : every],
// float32 / 255) -> video.py:75-86 (optional horizontal flip, Normalize(mean .5, std .5), (C, T, H, W)).
// The resize is OpenCV's 8-bit fixed-point bilinear, restated exactly: half-pixel centres, float fractions from a
// double coordinate, 11-bit coefficients (cvRound(f * 2048), saturated to short), x indices clamped WITH their
// fraction zeroed, y rows clamped WITHOUT touching the fraction, int horizontal pass, vertical pass
// ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2 >> 2; an exact 2x decimation is OpenCV's 2x2 box mean.
struct FrameGeom {
  int Hs, Ws, H, W, T, start, every, hflip, bgr;
  double sx, sy;      // Ws / W, Hs / H
};
__device__ __forceinline__ int sat_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }
__device__ __forceinline__ float cv_coord(int d, double scale, int& s) {
  const float f = static_cast<float>(__dsub_rn(__dmul_rn(__dadd_rn(static_cast<double>(d), 0.5), scale), 0.5));
  const float fl = floorf(f);
  s = static_cast<int>(fl);
  return __fsub_rn(f, fl);
}
constexpr int FC_ROWS = 4;      // output rows per block: a thread keeps its column coefficients for FC_ROWS x T pixels
__global__ void __launch_bounds__(256)
frames_to_clip_kernel(const uint8_t* __restrict__ frames, const FrameGeom g, float* __restrict__ clip) {
  pdl_grid_sync();
  // ((v / 255) - 0.5) / 0.5 for the 256 possible pixel values, each with the reference's correctly rounded fp32
  // operations, once per block: the per-pixel epilogue becomes three shared-memory lookups instead of six divisions
  __shared__ float lut[256];
  __shared__ int yrow[FC_ROWS][4];                     // r0 * Ws * 3, r1 * Ws * 3, b0, b1 of the block's rows
  lut[threadIdx.x] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(threadIdx.x), 255.f), 0.5f), 0.5f);
  const int h0 = blockIdx.y * FC_ROWS;
  const int mode = (g.Hs == g.H && g.Ws == g.W) ? 0 : (g.Hs == 2 * g.H && g.Ws == 2 * g.W) ? 1 : 2;
  if (mode == 2 && threadIdx.x < FC_ROWS && h0 + threadIdx.x < g.H) {
    int sy;
    const float fy = cv_coord(h0 + threadIdx.x, g.sy, sy);
    yrow[threadIdx.x][0] = min(max(sy, 0), g.Hs - 1) * g.Ws * 3;
    yrow[threadIdx.x][1] = min(max(sy + 1, 0), g.Hs - 1) * g.Ws * 3;
    yrow[threadIdx.x][2] = sat_short(__float2int_rn(__fmul_rn(__fsub_rn(1.f, fy), 2048.f)));
    yrow[threadIdx.x][3] = sat_short(__float2int_rn(__fmul_rn(fy, 2048.f)));
  }
  __syncthreads();
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= g.W) return;
  int x0 = 0, x1 = 0, a0 = 0, a1 = 0;                  // column taps: the same for every row and frame
  if (mode == 2) {
    int sx;
    float fx = cv_coord(w, g.sx, sx);
    if (sx < 0) { fx = 0.f; sx = 0; }
    if (sx >= g.Ws - 1) { fx = 0.f; sx = g.Ws - 1; }
    a0 = sat_short(__float2int_rn(__fmul_rn(__fsub_rn(1.f, fx), 2048.f)));
    a1 = sat_short(__float2int_rn(__fmul_rn(fx, 2048.f)));
    x0 = sx * 3;
    x1 = min(sx + 1, g.Ws - 1) * 3;
  }
  const long long plane = static_cast<long long>(g.H) * g.W;
  const size_t fsz = static_cast<size_t>(g.Hs) * g.Ws * 3;
  const int wo = g.hflip ? g.W - 1 - w : w;
  const int c0 = g.bgr ? 2 : 0, cstep = g.bgr ? -1 : 1;   // decoder order -> RGB (cv2.COLOR_BGR2RGB, generate_frames.py:42)
  const int rows = min(FC_ROWS, g.H - h0);
  for (int t = 0; t < g.T; ++t) {
    const uint8_t* f = frames + static_cast<size_t>(g.start + t * g.every) * fsz;
    for (int r = 0; r < rows; ++r) {
      const int h = h0 + r;
      int v[3];
      if (mode == 0) {
        const uint8_t* p = f + (static_cast<size_t>(h) * g.Ws + w) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = p[c];
      } else if (mode == 1) {
        const uint8_t* p0 = f + (static_cast<size_t>(2 * h) * g.Ws + 2 * w) * 3;
        const uint8_t* p1 = p0 + static_cast<size_t>(g.Ws) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = (p0[c] + p0[3 + c] + p1[c] + p1[3 + c] + 2) >> 2;
      } else {
        const uint8_t* p0 = f + yrow[r][0];
        const uint8_t* p1 = f + yrow[r][1];
        const int b0 = yrow[r][2], b1 = yrow[r][3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int S0 = p0[x0 + c] * a0 + p0[x1 + c] * a1;
          const int S1 = p1[x0 + c] * a0 + p1[x1 + c] * a1;
          const int o = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
          v[c] = min(max(o, 0), 255);
        }
      }
      float* dst = clip + (static_cast<long long>(t) * g.H + h) * g.W + wo;
#pragma unroll
      for (int c = 0; c < 3; ++c) dst[c * g.T * plane] = lut[v[c0 + cstep * c]];
    }
  }
}

// z[i] ~ N(0,1): Philox4x32-10 keyed by `seed`, counter = (element/4, offset [+ *d_offset]) — the device stand-in for
// the reference's host numpy draws (images.py:17-21, networks_3d.py:28-34) when the step is replayed as a CUDA graph.
__global__ void randn_kernel(float* __restrict__ z, long long n, unsigned long long seed, unsigned long long offset,
                             const unsigned long long* __restrict__ d_offset) {
  pdl_grid_sync();
  if (d_offset) offset += *d_offset;
  const long long n4 = (n + 3) >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 rnd = philox4x32_10(make_uint4(static_cast<uint32_t>(i), static_cast<uint32_t>(i >> 32),
                                               static_cast<uint32_t>(offset), static_cast<uint32_t>(offset >> 32)),
                                    make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
    float v[4];
    box_muller(rnd.x, rnd.y, v[0], v[1]);
    box_muller(rnd.z, rnd.w, v[2], v[3]);
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (4 * i + e < n) z[4 * i + e] = v[e];
  }
}
__global__ void counter_add_kernel(unsigned long long* c, unsigned long long inc) {
  pdl_grid_sync(); *c += inc; }

// block input stage (networks_3d.py:440-446): up = resize(x_prev) ; x_in = up + noise*amp ; C <= 4
__global__ void upsample_noise_pack_kernel(const float* __restrict__ x, int N, int C, ResizeGeom g,
                                           const float* __restrict__ noise, float amp, unsigned long long seed,
                                           unsigned long long sample_base,
                                           const unsigned long long* __restrict__ d_sample_offset,
                                           float* __restrict__ up, __nv_bfloat16* __restrict__ xin, int xin_f32) {
  pdl_grid_sync();
  if (d_sample_offset) sample_base += *d_sample_offset;   // device-resident draw counter (CUDA-graph replays)
  const long long spo = static_cast<long long>(g.To) * g.Ho * g.Wo;
  const long long spi = static_cast<long long>(g.Ti) * g.Hi * g.Wi;
  const long long total = static_cast<long long>(N) * spo;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = idx / spo, s = idx - n * spo;
    const int w = static_cast<int>(s % g.Wo);
    const long long r = s / g.Wo;
    const int h = static_cast<int>(r % g.Ho);
    const int t = static_cast<int>(r / g.Ho);
    const Tap tt = linear_tap(t, g.Ti, g.st, g.align), th = linear_tap(h, g.Hi, g.sh, g.align),
              tw = linear_tap(w, g.Wi, g.sw, g.align);
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    if (seed != 0ull) {
      const unsigned long long sample = sample_base + static_cast<unsigned long long>(n);
      const uint4 rnd = philox4x32_10(make_uint4(static_cast<uint32_t>(s), static_cast<uint32_t>(s >> 32),
                                                 static_cast<uint32_t>(sample), static_cast<uint32_t>(sample >> 32)),
                                      make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
      box_muller(rnd.x, rnd.y, z[0], z[1]);
      box_muller(rnd.z, rnd.w, z[2], z[3]);
    }
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < C; ++c) {
      const float u = trilerp(x + (n * C + c) * spi, g.Hi, g.Wi, tt, th, tw);
      const long long o = (n * C + c) * spo + s;
      up[o] = u;
      float nz = 0.f;
      if (noise) nz = noise[o];
      else if (seed != 0ull) nz = z[c];
      v[c] = fmaf(nz, amp, u);
    }
    *reinterpret_cast<uint4*>(xin + idx * 8) = xin_voxel(v, xin_f32);
  }
}

// ----------------------------------------------------------------------------------------------- BatchNorm (train)
// y: (voxels, 64) bf16.  thread -> one 16 B channel group; 8 groups per voxel; block = 256 threads = 32 voxels/iter.
__global__ void bn_stats_cl_kernel(const __nv_bfloat16* __restrict__ y, long long voxels, double* __restrict__ sum,
                                   double* __restrict__ sumsq, const DetScratch det) {
  pdl_grid_sync();
  const int g = threadIdx.x & 7;         // channel group
  const int vl = threadIdx.x >> 3;       // voxel lane 0..31
  float a[8], b[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = b[e] = 0.f;
  auto one = [&](const uint4& pk) {
    const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
    for (int e2 = 0; e2 < 4; ++e2) {
      const float2 f = unpack2(w[e2]);
      a[2 * e2] += f.x;
      a[2 * e2 + 1] += f.y;
      b[2 * e2] = fmaf(f.x, f.x, b[2 * e2]);
      b[2 * e2 + 1] = fmaf(f.y, f.y, b[2 * e2 + 1]);
    }
  };
  // four voxels in flight per thread (one 16-byte load each), accumulated in the order of a one-voxel walk
  const long long vs = static_cast<long long>(gridDim.x) * 32;
  long long v = static_cast<long long>(blockIdx.x) * 32 + vl;
  for (; v + 3 * vs < voxels; v += 4 * vs) {
    const uint4 p0 = *reinterpret_cast<const uint4*>(y + v * 64 + g * 8);
    const uint4 p1 = *reinterpret_cast<const uint4*>(y + (v + vs) * 64 + g * 8);
    const uint4 p2 = *reinterpret_cast<const uint4*>(y + (v + 2 * vs) * 64 + g * 8);
    const uint4 p3 = *reinterpret_cast<const uint4*>(y + (v + 3 * vs) * 64 + g * 8);
    one(p0);
    one(p1);
    one(p2);
    one(p3);
  }
  for (; v < voxels; v += vs) one(*reinterpret_cast<const uint4*>(y + v * 64 + g * 8));
  // reduce over the 32 voxel lanes that share a channel group: lanes with equal (threadIdx.x & 7)
  // within a warp: 4 voxel lanes per group -> shuffle over xor 8, 16 ; then across 8 warps via smem
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    a[e] += __shfl_xor_sync(0xffffffffu, a[e], 8);
    a[e] += __shfl_xor_sync(0xffffffffu, a[e], 16);
    b[e] += __shfl_xor_sync(0xffffffffu, b[e], 8);
    b[e] += __shfl_xor_sync(0xffffffffu, b[e], 16);
  }
  __shared__ float red[8][2][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < 8) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      red[warp][0][lane * 8 + e] = a[e];
      red[warp][1][lane * 8 + e] = b[e];
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

__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* __restrict__ mm, float* __restrict__ mv,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_o,
                                   float* __restrict__ invstd_o) {
  pdl_grid_sync();
  const int c = threadIdx.x;
  if (c >= 64) return;
  const double mean = sum[c] / count;
  double var = sumsq[c] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - static_cast<float>(mean) * sc;
  if (mean_o) mean_o[c] = static_cast<float>(mean);
  if (invstd_o) invstd_o[c] = invstd;
  if (mm) mm[c] = momentum * mm[c] + (1.f - momentum) * static_cast<float>(mean);
  if (mv) mv[c] = momentum * mv[c] + (1.f - momentum) * static_cast<float>(var);
}

__global__ void bn_apply_cl_kernel(const __nv_bfloat16* __restrict__ y, long long groups /*voxels*8*/,
                                   const float* __restrict__ scale, const float* __restrict__ shift, int act,
                                   __nv_bfloat16* __restrict__ x) {
  pdl_grid_sync();
  __shared__ float sc[64], sh[64];
  if (threadIdx.x < 64) {
    sc[threadIdx.x] = scale[threadIdx.x];
    sh[threadIdx.x] = shift[threadIdx.x];
  }
  __syncthreads();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < groups;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i & 7);
    const uint4 pk = *reinterpret_cast<const uint4*>(y + i * 8);
    const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
    uint32_t o[4];
#pragma unroll
    for (int e2 = 0; e2 < 4; ++e2) {
      float2 f = unpack2(w[e2]);
      const int c = g * 8 + 2 * e2;
      f.x = fmaf(f.x, sc[c], sh[c]);
      f.y = fmaf(f.y, sc[c + 1], sh[c + 1]);
      if (act == 1) {
        f.x = f.x > 0.f ? f.x : 0.2f * f.x;
        f.y = f.y > 0.f ? f.y : 0.2f * f.y;
      }
      o[e2] = pack2(f.x, f.y);
    }
    *reinterpret_cast<uint4*>(x + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// Training-mode BatchNorm in ONE pass over the data: the per-channel sums come from the producing conv's epilogue
// (conv3d_umma.cu), every block redoes the 64-channel finalisation (cheaper than a separate launch), block 0 also
// writes the saved (scale, shift, mean, invstd) vectors for the backward and updates the moving statistics.
__global__ void bn_train_apply_cl_kernel(const __nv_bfloat16* __restrict__ y, long long groups /*voxels*8*/,
                                         const double* __restrict__ sums /*[2][64]*/, double count,
                                         const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                         float momentum, float* __restrict__ mm, float* __restrict__ mv,
                                         float* __restrict__ saved /*[4][64], nullable*/, int act,
                                         __nv_bfloat16* __restrict__ x, const float* __restrict__ center) {
  pdl_grid_sync();
  // center (nullable): the per-channel offset the producing conv subtracted from y before storing it (kink-centred
  // storage, see bn_center_multi_kernel).  Everything here and in the backward works in the centred frame — BatchNorm
  // is shift invariant — except the MOVING mean, which tracks the true mean = centred mean + center.
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
  // a thread's channel group is fixed (block size and grid stride are multiples of 8 groups): its 8 (scale, shift)
  // pairs live in registers instead of 16 shared-memory reads per 16-byte group; two groups in flight per thread
  const int g = threadIdx.x & 7;
  float csc[8], csh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    csc[e] = sc[g * 8 + e];
    csh[e] = sh[g * 8 + e];
  }
  auto one = [&](long long i, const uint4& pk) {
    const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
    uint32_t o[4];
#pragma unroll
    for (int e2 = 0; e2 < 4; ++e2) {
      float2 f = unpack2(w[e2]);
      f.x = fmaf(f.x, csc[2 * e2], csh[2 * e2]);
      f.y = fmaf(f.y, csc[2 * e2 + 1], csh[2 * e2 + 1]);
      if (act == 1) {
        f.x = f.x > 0.f ? f.x : 0.2f * f.x;
        f.y = f.y > 0.f ? f.y : 0.2f * f.y;
      }
      o[e2] = pack2(f.x, f.y);
    }
    *reinterpret_cast<uint4*>(x + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  };
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + stride < groups; i += 2 * stride) {
    const uint4 p0 = *reinterpret_cast<const uint4*>(y + i * 8);
    const uint4 p1 = *reinterpret_cast<const uint4*>(y + (i + stride) * 8);
    one(i, p0);
    one(i + stride, p1);
  }
  if (i < groups) one(i, *reinterpret_cast<const uint4*>(y + i * 8));
}

// Deferred moving-statistics update for many BatchNorm layers in one launch (block b = layer b): when several
// training-mode forwards of the same network run CONCURRENTLY (hpvg/train.py:GraphedIteration) their read-modify-write
// of the shared moving statistics is taken out of the forwards and replayed afterwards, forward by forward, in the
// reference's order.  saved = (scale, shift, mean, invstd) of bn_train_apply_cl; var = 1/invstd^2 - eps.
__global__ void bn_moving_update_multi_kernel(const BnMovingTable tab, float eps, float momentum) {
  pdl_grid_sync();
  const int b = blockIdx.x, c = threadIdx.x;
  const float* sv = tab.saved[b];
  const float mean = sv[128 + c] + (tab.center[b] ? tab.center[b][c] : 0.f), invstd = sv[192 + c];
  const float var = fmaxf(1.0f / (invstd * invstd) - eps, 0.f);
  tab.mm[b][c] = momentum * tab.mm[b][c] + (1.f - momentum) * mean;
  tab.mv[b][c] = momentum * tab.mv[b][c] + (1.f - momentum) * var;
}

// Kink-centred storage of the pre-BatchNorm activation (bf16 mode).  A training-mode layer stores y = conv(x) + bias in
// bf16 and its backward recomputes the LeakyReLU mask from that stored y; bf16 rounding of y moves ~1e-3 of the
// pre-activations across the kink relative to fp32 arithmetic (the rounding error is relative to |y|, and the kink of
// channel c sits at y = mean_c - beta_c / (gamma_c * invstd_c), usually far from 0).  Storing y - center_c with
// center_c = the ESTIMATED kink (moving statistics and the current gamma / beta: all known before the conv runs) puts
// bf16's finest resolution exactly where the mask is decided.  BatchNorm is shift invariant, so the statistics, the
// normalise pass and the backward simply run in the centred frame; only the moving mean needs center added back.
// One block per layer: center[c] and the conv's epilogue vectors aff = (1, bias - center).
__global__ void bn_center_multi_kernel(const BnCenterTable tab, float eps) {
  pdl_grid_sync();
  const int b = blockIdx.x, c = threadIdx.x;
  const float g = tab.gamma[b][c];
  float cen = tab.mm[b][c];
  if (fabsf(g) > 1e-3f) cen -= tab.beta[b][c] * sqrtf(tab.mv[b][c] + eps) / g;
  cen = __bfloat162float(__float2bfloat16_rn(cen));     // a bf16 value: exactly representable in every frame
  tab.center[b][c] = cen;
  tab.aff[b][c] = 1.0f;
  tab.aff[b][64 + c] = tab.bias[b][c] - cen;
}

// ----------------------------------------------------------------------------------------------- spectral norm
// One CTA.  W: (cout, k) fp32 row-major (the (Cout, Cin*27) view of spectral_norm.py:146).
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

__device__ void sn_power_iter_body(const float* __restrict__ w, int cout, int k, float* __restrict__ u,
                                   float* __restrict__ v, float* __restrict__ sigma, float* __restrict__ inv_sigma,
                                   const float* __restrict__ bias, float* __restrict__ aff /*[2][64] nullable*/,
                                   float* __restrict__ u_copy, float* __restrict__ v_copy, float* sm, float* red) {
  float* su = sm;            // [cout]
  float* sv = su + cout;     // [k]
  float* swv = sv + k;       // [cout]
  for (int i = threadIdx.x; i < cout; i += blockDim.x) su[i] = u[i];
  __syncthreads();
  // v = l2normalize(W^T u)
  float part = 0.f;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < cout; ++i) acc = fmaf(w[static_cast<size_t>(i) * k + j], su[i], acc);
    sv[j] = acc;
    part = fmaf(acc, acc, part);
  }
  float nrm = block_sum(part, red);
  float inv = rsqrtf(fmaxf(nrm, 1e-12f));
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    sv[j] *= inv;
    v[j] = sv[j];
    if (v_copy) v_copy[j] = sv[j];
  }
  __syncthreads();
  // Wv (one warp per output row, strided over rows)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int i = warp; i < cout; i += nw) {
    float acc = 0.f;
    for (int j = lane; j < k; j += 32) acc = fmaf(w[static_cast<size_t>(i) * k + j], sv[j], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) swv[i] = acc;
  }
  __syncthreads();
  part = 0.f;
  for (int i = threadIdx.x; i < cout; i += blockDim.x) part = fmaf(swv[i], swv[i], part);
  nrm = block_sum(part, red);
  inv = rsqrtf(fmaxf(nrm, 1e-12f));
  // u = l2normalize(Wv) ; sigma = u^T (W v)
  part = 0.f;
  for (int i = threadIdx.x; i < cout; i += blockDim.x) {
    const float un = swv[i] * inv;
    u[i] = un;
    if (u_copy) u_copy[i] = un;
    part = fmaf(un, swv[i], part);
  }
  const float sg = block_sum(part, red);
  if (threadIdx.x == 0) {
    sigma[0] = sg;
    inv_sigma[0] = 1.0f / sg;
  }
  if (aff && threadIdx.x < 64) {   // conv epilogue vectors: scale = 1/sigma, shift = bias
    aff[threadIdx.x] = 1.0f / sg;
    aff[64 + threadIdx.x] = (bias && threadIdx.x < cout) ? bias[threadIdx.x] : 0.f;
  }
}

__global__ void sn_power_iter_kernel(const float* __restrict__ w, int cout, int k, float* __restrict__ u,
                                     float* __restrict__ v, float* __restrict__ sigma,
                                     float* __restrict__ inv_sigma) {
  pdl_grid_sync();
  extern __shared__ float sm[];
  __shared__ float red[32];
  sn_power_iter_body(w, cout, k, u, v, sigma, inv_sigma, nullptr, nullptr, nullptr, nullptr, sm, red);
}

// all spectrally normalised layers of a network in one launch: block b = layer b
__global__ void sn_power_iter_multi_kernel(const SnTable tab) {
  pdl_grid_sync();
  extern __shared__ float sm[];
  __shared__ float red[32];
  const int b = blockIdx.x;
  sn_power_iter_body(tab.w[b], tab.cout[b], tab.k[b], tab.u[b], tab.v[b], tab.sigma[b], tab.sigma[b] + 1, tab.bias[b],
                     tab.aff[b], tab.u_copy[b], tab.v_copy[b], sm, red);
}

// ----------------------------------------------------------------------------------------------- epilogue vectors
__global__ void bn_fold_eval_kernel(const float* gamma, const float* beta, const float* mean, const float* var,
                                    float eps, const float* bias, int C, float* scale, float* shift) {
  pdl_grid_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] / sqrtf(var[c] + eps);
  scale[c] = sc;
  shift[c] = (bias[c] - mean[c]) * sc + beta[c];
}
__global__ void affine_from_bias_kernel(const float* bias, const float* inv_sigma, int C, float* scale, float* shift) {
  pdl_grid_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  scale[c] = inv_sigma ? inv_sigma[0] : 1.0f;
  shift[c] = bias ? bias[c] : 0.f;
}

// ----------------------------------------------------------------------------------------------- reductions
template <int OP>  // 0: sum (a-b)^2 ; 1: sum a ; 2: sum -0.5(1+lv-mu^2-exp(lv))
__global__ void reduce_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float inv_n,
                              float* __restrict__ out, const DetScratch det) {
  pdl_grid_sync();
  __shared__ float red[32];
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (OP == 0) {
      const float d = a[i] - b[i];
      acc = fmaf(d, d, acc);
    } else if (OP == 1) {
      acc += a[i];
    } else {
      const float mu = a[i], lv = b[i];
      acc += -0.5f * (1.f + lv - mu * mu - expf(lv));
    }
  }
  const float t = block_sum(acc, red);
  det_reduce_scalar(t, blockIdx.x, gridDim.x, det, inv_n, out);
}

__global__ void reparam_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                               const float* __restrict__ eps, long long n, float* __restrict__ z) {
  pdl_grid_sync();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    z[i] = fmaf(eps[i], expf(0.5f * lv[i]), mu[i]);
}

// backward of z = eps*exp(0.5*logvar) + mu (networks_3d.py:415-417): gmu += gz ; glogvar += gz*eps*0.5*exp(0.5*logvar)
__global__ void reparam_bwd_kernel(const float* __restrict__ gz, const float* __restrict__ eps,
                                   const float* __restrict__ lv, long long n, float* __restrict__ gmu,
                                   float* __restrict__ glv) {
  pdl_grid_sync();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float g = gz[i];
    gmu[i] += g;
    glv[i] = fmaf(g * eps[i], 0.5f * expf(0.5f * lv[i]), glv[i]);
  }
}

// ----------------------------------------------------------------------------------------------- clip + Adam
// grid (blocks, tensors); 16-byte vector accesses on the aligned body, scalar tail.  cudaMalloc'd tensors are
// 256-byte aligned; a misaligned view falls back to the scalar loop (vec = 0).
__global__ void adam_norm_kernel(const AdamTable tab, float* __restrict__ norms) {
  pdl_grid_sync();
  __shared__ float red[32];
  const int t = blockIdx.y;
  const float* __restrict__ g = tab.g[t];
  const long long n = tab.n[t];
  const bool vec = (reinterpret_cast<uintptr_t>(g) & 15) == 0;
  const long long n4 = vec ? n >> 2 : 0;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nth = static_cast<long long>(gridDim.x) * blockDim.x;
  float acc = 0.f;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = tid; i < n4; i += nth) {
    const float4 v = g4[i];
    acc = fmaf(v.x, v.x, acc);
    acc = fmaf(v.y, v.y, acc);
    acc = fmaf(v.z, v.z, acc);
    acc = fmaf(v.w, v.w, acc);
  }
  for (long long i = (n4 << 2) + tid; i < n; i += nth) acc = fmaf(g[i], g[i], acc);
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) norms[t * gridDim.x + blockIdx.x] = s;     // block partial; summed in block order by the apply kernel
}

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float coef, float beta1, float beta2,
                                         float eps, float lr_t) {
  const float gi = g * coef;
  m = m + (gi - m) * (1.f - beta1);
  v = v + (gi * gi - v) * (1.f - beta2);
  p = p - lr_t * m / (sqrtf(v) + eps);
}

__global__ void adam_apply_kernel(const AdamTable tab, const float* __restrict__ norms, float beta1, float beta2,
                                  float eps, float bc /* sqrt(1-b2^t)/(1-b1^t) */, float clip,
                                  const unsigned long long* __restrict__ d_step, const float* __restrict__ d_hyper,
                                  int norm_blocks) {
  pdl_grid_sync();
  if (d_hyper) {   // every hyper-parameter lives on the device (ops.Custom binding): lr, beta1, beta2, eps, clip, step
    beta1 = d_hyper[1];
    beta2 = d_hyper[2];
    eps = d_hyper[3];
    clip = d_hyper[4];
    const double t = static_cast<double>(d_hyper[5]);
    bc = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), t)) / (1.0 - pow(static_cast<double>(beta1), t)));
  } else if (d_step) {   // 1-based step lives on the device (CUDA-graph replays): same formula as the host path, in fp64
    const double t = static_cast<double>(*d_step);
    bc = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), t)) / (1.0 - pow(static_cast<double>(beta1), t)));
  }
  const int t = blockIdx.y;
  float* __restrict__ p = tab.p[t];
  const float* __restrict__ g = tab.g[t];
  float* __restrict__ m = tab.m[t];
  float* __restrict__ v = tab.v[t];
  const long long n = tab.n[t];
  // (the table is a kernel PARAMETER: it is only ever read — a dynamic write would force a per-thread local copy of it)
  const float lr_t = (d_hyper ? d_hyper[0] : tab.lr[t]) * bc;
  float coef = 1.0f;
  if (clip > 0.f) {   // ClipByNorm: g*c / max(||g||, c); ||g||^2 = the norm kernel's block partials in block order
    float nrm2 = 0.f;
    for (int b = 0; b < norm_blocks; ++b) nrm2 += norms[t * norm_blocks + b];
    coef = clip / fmaxf(sqrtf(nrm2), clip);
  }
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const long long n4 = vec ? n >> 2 : 0;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nth = static_cast<long long>(gridDim.x) * blockDim.x;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (long long i = tid; i < n4; i += nth) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = g4[i];
    adam_one(pp.x, gg.x, mm.x, vv.x, coef, beta1, beta2, eps, lr_t);
    adam_one(pp.y, gg.y, mm.y, vv.y, coef, beta1, beta2, eps, lr_t);
    adam_one(pp.z, gg.z, mm.z, vv.z, coef, beta1, beta2, eps, lr_t);
    adam_one(pp.w, gg.w, mm.w, vv.w, coef, beta1, beta2, eps, lr_t);
    m4[i] = mm;
    v4[i] = vv;
    p4[i] = pp;
  }
  for (long long i = (n4 << 2) + tid; i < n; i += nth) adam_one(p[i], g[i], m[i], v[i], coef, beta1, beta2, eps, lr_t);
}

// ----------------------------------------------------------------------------------------------- backward pieces
// gz = ga * lrelu'(a)  (a = the stored activation; sign(a) == sign(pre-activation))          bf16 cl, any width
__global__ void lrelu_bwd_cl_kernel(const __nv_bfloat16* __restrict__ ga, const __nv_bfloat16* __restrict__ a,
                                    long long groups, __nv_bfloat16* __restrict__ gz) {
  pdl_grid_sync();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < groups;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 g4 = *reinterpret_cast<const uint4*>(ga + i * 8);
    const uint4 a4 = *reinterpret_cast<const uint4*>(a + i * 8);
    const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w}, aw[4] = {a4.x, a4.y, a4.z, a4.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 g = unpack2(gw[e]);
      const float2 av = unpack2(aw[e]);
      g.x = av.x > 0.f ? g.x : 0.2f * g.x;
      g.y = av.y > 0.f ? g.y : 0.2f * g.y;
      o[e] = pack2(g.x, g.y);
    }
    *reinterpret_cast<uint4*>(gz + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// BatchNorm(train)+LeakyReLU backward, pass 1: sums[0][c] = sum gz, sums[1][c] = sum gz*xhat
//   z = y*scale+shift (mask), xhat = (y-mean)*invstd, gz = ga*lrelu'(z)
__global__ void bn_bwd_reduce_cl_kernel(const __nv_bfloat16* __restrict__ ga, const __nv_bfloat16* __restrict__ y,
                                        long long voxels, const float* __restrict__ saved /*[4][64]*/, int act,
                                        double* __restrict__ sums, const DetScratch det) {
  pdl_grid_sync();
  const int g = threadIdx.x & 7, vl = threadIdx.x >> 3;
  float sc[8], sh[8], mu[8], is[8], s0[8], s1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sc[e] = saved[g * 8 + e];
    sh[e] = saved[64 + g * 8 + e];
    mu[e] = saved[128 + g * 8 + e];
    is[e] = saved[192 + g * 8 + e];
    s0[e] = s1[e] = 0.f;
  }
  auto one = [&](const uint4& g4, const uint4& y4) {
    const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w}, yw[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
    for (int e2 = 0; e2 < 4; ++e2) {
      const float2 gg = unpack2(gw[e2]), yy = unpack2(yw[e2]);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int e = 2 * e2 + h;
        const float yv = h ? yy.y : yy.x;
        float gv = h ? gg.y : gg.x;
        if (act == 1 && fmaf(yv, sc[e], sh[e]) <= 0.f) gv *= 0.2f;
        s0[e] += gv;
        s1[e] = fmaf(gv, (yv - mu[e]) * is[e], s1[e]);
      }
    }
  };
  // two voxels (four 16-byte loads) in flight per thread; accumulated in the same order as a one-voxel walk
  const long long vs = static_cast<long long>(gridDim.x) * 32;
  long long v = static_cast<long long>(blockIdx.x) * 32 + vl;
  for (; v + vs < voxels; v += 2 * vs) {
    const uint4 ga0 = *reinterpret_cast<const uint4*>(ga + v * 64 + g * 8);
    const uint4 y0 = *reinterpret_cast<const uint4*>(y + v * 64 + g * 8);
    const uint4 ga1 = *reinterpret_cast<const uint4*>(ga + (v + vs) * 64 + g * 8);
    const uint4 y1 = *reinterpret_cast<const uint4*>(y + (v + vs) * 64 + g * 8);
    one(ga0, y0);
    one(ga1, y1);
  }
  if (v < voxels)
    one(*reinterpret_cast<const uint4*>(ga + v * 64 + g * 8), *reinterpret_cast<const uint4*>(y + v * 64 + g * 8));
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    s0[e] += __shfl_xor_sync(0xffffffffu, s0[e], 8);
    s0[e] += __shfl_xor_sync(0xffffffffu, s0[e], 16);
    s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], 8);
    s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], 16);
  }
  __shared__ float red[8][2][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < 8) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      red[warp][0][lane * 8 + e] = s0[e];
      red[warp][1][lane * 8 + e] = s1[e];
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

// fp32 pair -> bf16 pair with STOCHASTIC rounding (16 hashed dither bits per value, keyed by the element index, so the
// result is reproducible).  E[round(v)] == v for every v, whatever its position on the bf16 grid.
__device__ __forceinline__ uint32_t pack2_sr(float lo, float hi, uint32_t key) {
  key ^= key >> 16;            // lowbias32
  key *= 0x7feb352dU;
  key ^= key >> 15;
  key *= 0x846ca68bU;
  key ^= key >> 16;
  const uint32_t a = __float_as_uint(lo) + (key & 0xFFFFu);
  const uint32_t b = __float_as_uint(hi) + (key >> 16);
  return (a >> 16) | (b & 0xFFFF0000u);
}

// pass 2: gy = gamma*invstd*(gz - mean(gz) - xhat*mean(gz*xhat))   [scale == gamma*invstd]
// The two mean terms are ~1e-3 of |gz|, gz itself is an exact product of bf16 values, and when gamma*invstd sits near a
// power of two scale*gz lands (almost) on the bf16 grid: round-to-nearest then snaps most elements back to it and the
// mean corrections are systematically lost — gy keeps a per-channel DC of ~scale*mean(gz) which the weight gradient
// sum_v gy[v]*x[v+tap] amplifies by N*mean(x) (measured at 13x192x257: dW off by 1.4-4 % while every other gradient
// is within 3e-3).  Stochastic rounding keeps the stored gy unbiased: sum_v gy == 0 up to sqrt(N) noise.
__global__ void __launch_bounds__(256, 3)
bn_bwd_apply_cl_kernel(const __nv_bfloat16* __restrict__ ga, const __nv_bfloat16* __restrict__ y, long long groups,
                       const float* __restrict__ saved, int act, const double* __restrict__ sums, double inv_count,
                       __nv_bfloat16* __restrict__ gy, float* __restrict__ dgamma, float* __restrict__ dbeta,
                       int accumulate) {
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
    if (blockIdx.x == 0) {   // dbeta = sum gz, dgamma = sum gz*xhat: written here instead of by two more launches
      if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + static_cast<float>(sums[c]);
      if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + static_cast<float>(sums[64 + c]);
    }
  }
  __syncthreads();
  // A thread's channel group never changes (block size and grid stride are multiples of 8 groups), so its 8 channels'
  // coefficients live in registers: read from shared memory per element they were 48 wavefronts per 16-byte group
  // against 3 global accesses, and the kernel sat at 91 % L1/shared throughput (ncu) and 59 % of the copy bandwidth.
  const int g = threadIdx.x & 7;
  float csc[8], csh[8], cmu[8], ck1[8], cm0[8];   // ck1 = inv_std * mean(gz * xhat): one register array fewer
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    csc[e] = sc[g * 8 + e];
    csh[e] = sh[g * 8 + e];
    cmu[e] = mu[g * 8 + e];
    ck1[e] = is[g * 8 + e] * m1[g * 8 + e];
    cm0[e] = m0[g * 8 + e];
  }
  auto one = [&](long long i, const uint4& g4, const uint4& y4) {
    const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w}, yw[4] = {y4.x, y4.y, y4.z, y4.w};
    uint32_t o[4];
#pragma unroll
    for (int e2 = 0; e2 < 4; ++e2) {
      const float2 gg = unpack2(gw[e2]), yy = unpack2(yw[e2]);
      float r[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int e = 2 * e2 + h;
        const float yv = h ? yy.y : yy.x;
        float gv = h ? gg.y : gg.x;
        if (act == 1 && fmaf(yv, csc[e], csh[e]) <= 0.f) gv *= 0.2f;
        r[h] = csc[e] * (gv - cm0[e] - (yv - cmu[e]) * ck1[e]);
      }
      o[e2] = pack2_sr(r[0], r[1], static_cast<uint32_t>(i) * 4u + e2);
    }
    *reinterpret_cast<uint4*>(gy + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  };
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + stride < groups; i += 2 * stride) {   // two groups in flight per thread
    const uint4 ga0 = *reinterpret_cast<const uint4*>(ga + i * 8);
    const uint4 y0 = *reinterpret_cast<const uint4*>(y + i * 8);
    const uint4 ga1 = *reinterpret_cast<const uint4*>(ga + (i + stride) * 8);
    const uint4 y1 = *reinterpret_cast<const uint4*>(y + (i + stride) * 8);
    one(i, ga0, y0);
    one(i + stride, ga1, y1);
  }
  if (i < groups) one(i, *reinterpret_cast<const uint4*>(ga + i * 8), *reinterpret_cast<const uint4*>(y + i * 8));
}

// out[c] (+)= scale * in[c]  (double -> float), used for dgamma / dbeta / bias gradients
__global__ void d2f_kernel(const double* __restrict__ in, int n, float scale, int accumulate, float* __restrict__ out) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (accumulate ? out[i] : 0.f) + scale * static_cast<float>(in[i]);
}

// g (+)= coef*(a - b)                                               (MSE gradient, coef = weight*2/n)
__global__ void diff_scale_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float coef,
                                  int accumulate, float* __restrict__ g) {
  pdl_grid_sync();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    g[i] = (accumulate ? g[i] : 0.f) + coef * (a[i] - b[i]);
}
// gpre = g*(1 - out^2)
__global__ void tanh_bwd_kernel(const float* __restrict__ g, const float* __restrict__ out, long long n,
                                float* __restrict__ gpre) {
  pdl_grid_sync();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    gpre[i] = g[i] * (1.f - out[i] * out[i]);
}
// y = a*x + b*y
__global__ void axpby_kernel(float a, const float* __restrict__ x, float b, float* __restrict__ y, long long n) {
  pdl_grid_sync();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    y[i] = a * x[i] + (b == 0.f ? 0.f : b * y[i]);
}
// dst[i] = src[i*stride + offset]   (e.g. one tap of a (Cout,Cin,27) weight-gradient tensor)
__global__ void gather_strided_kernel(const float* __restrict__ src, long long n, long long stride, long long offset,
                                      float* __restrict__ dst) {
  pdl_grid_sync();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] = src[i * stride + offset];
}
__global__ void fill_kernel(float* __restrict__ y, float v, long long n) {
  pdl_grid_sync();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    y[i] = v;
}
// out[c] (+)= sum over n, spatial of g[n][c][s]                       (bias gradient of the fp32 NCDHW tails)
// grid (blocks, C): fp32 block partials -> fp64 atomics into scratch[c]; a d2f pass writes the result.
__global__ void channel_sum_ncdhw_kernel(const float* __restrict__ g, int N, int C, long long sp,
                                         double* __restrict__ scratch) {
  pdl_grid_sync();
  __shared__ float red[32];
  const int c = blockIdx.y;
  float acc = 0.f;
  for (int n = 0; n < N; ++n) {
    const float* p = g + (static_cast<long long>(n) * C + c) * sp;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < sp;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
      acc += p[i];
  }
  const float t = block_sum(acc, red);
  if (threadIdx.x == 0) scratch[static_cast<size_t>(c) * gridDim.x + blockIdx.x] = static_cast<double>(t);
}
// out[c] (+)= sum of the nblk block partials of channel c, in block order (deterministic)
__global__ void channel_sum_final_kernel(const double* __restrict__ partials, int nblk, int C, int accumulate,
                                         float* __restrict__ out) {
  pdl_grid_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double acc = 0.0;
  for (int b = 0; b < nblk; ++b) acc += partials[static_cast<size_t>(c) * nblk + b];
  out[c] = (accumulate ? out[c] : 0.f) + static_cast<float>(acc);
}
// KL gradient (losses.py:5-7): d/dmu = coef*mu ; d/dlogvar = coef*0.5*(exp(lv)-1) ; coef = kl_weight/n
__global__ void kl_grad_kernel(const float* __restrict__ mu, const float* __restrict__ lv, long long n, float coef,
                               float* __restrict__ gmu, float* __restrict__ glv) {
  pdl_grid_sync();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    gmu[i] = coef * mu[i];
    glv[i] = coef * 0.5f * (expf(lv[i]) - 1.f);
  }
}
// spectral-norm chain rule (u, v constants): gW (+)= (G - <G, W/sigma> u v^T) / sigma
// pass 1: SN_GRAD_BLOCKS partial dot products <G, W> ; pass 2: every block sums them in fixed order and applies.
constexpr int SN_GRAD_BLOCKS = 64;
__global__ void sn_grad_dot_kernel(const float* __restrict__ G, const float* __restrict__ w, int n,
                                   float* __restrict__ partial) {
  pdl_grid_sync();
  __shared__ float red[32];
  float part = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    part = fmaf(G[i], w[i], part);
  const float t = block_sum(part, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
__global__ void sn_grad_apply_kernel(const float* __restrict__ G, const float* __restrict__ partial,
                                     const float* __restrict__ u, const float* __restrict__ v,
                                     const float* __restrict__ sigma, int cout, int k, int accumulate,
                                     float* __restrict__ gw) {
  pdl_grid_sync();
  const float sg = sigma[0];
  float dot = 0.f;
  for (int i = 0; i < SN_GRAD_BLOCKS; ++i) dot += partial[i];
  dot /= sg;   // <G, W/sigma>
  const int n = cout * k;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = i / k, c = i - r * k;
    const float val = (G[i] - dot * u[r] * v[c]) / sg;
    gw[i] = (accumulate ? gw[i] : 0.f) + val;
  }
}
// WGAN-GP (losses.py:47-52): xhat = alpha*real + (1-alpha)*fake
__global__ void lerp_kernel(const float* __restrict__ a, const float* __restrict__ b, float alpha, long long n,
                            float* __restrict__ out) {
  pdl_grid_sync();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = alpha * a[i] + (1.f - alpha) * b[i];
}
// per voxel: nrm = ||g[:, v]||_2 over C channels ; gp += lambda*(nrm-1)^2/V ; G[c] = lambda*2*(nrm-1)/nrm*g[c]/V
__global__ void gp_grad_kernel(const float* __restrict__ g, int N, int C, long long sp, float lambda,
                               float* __restrict__ Gout, float* __restrict__ gp, const DetScratch det) {
  pdl_grid_sync();
  __shared__ float red[32];
  const long long V = static_cast<long long>(N) * sp;
  const float invV = 1.0f / static_cast<float>(V);
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < V;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = i / sp, s = i - n * sp;
    float ss = 0.f;
    for (int c = 0; c < C; ++c) {
      const float x = g[(n * C + c) * sp + s];
      ss = fmaf(x, x, ss);
    }
    const float nrm = sqrtf(ss);
    acc += (nrm - 1.f) * (nrm - 1.f);
    const float coef = nrm > 0.f ? lambda * 2.f * (nrm - 1.f) / nrm * invV : 0.f;
    for (int c = 0; c < C; ++c) Gout[(n * C + c) * sp + s] = coef * g[(n * C + c) * sp + s];
  }
  const float t = block_sum(acc, red);
  det_reduce_scalar(t, blockIdx.x, gridDim.x, det, lambda * invV, gp);
}

inline int grid_for(long long n, int block, int cap = 148 * 16) {
  long long b = (n + block - 1) / block;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace

// =============================================================================================== host wrappers
#define LAUNCH_CHECK()                          \
  do {                                          \
    cudaError_t e_ = cudaGetLastError();        \
    if (e_ != cudaSuccess) return e_;           \
  } while (0)

cudaError_t ew_pack_cl(const float* x, int N, int C, long long sp, __nv_bfloat16* y, int c_pitch, int c_off,
                       int c_zero_to, cudaStream_t st) {
  const long long voxels = static_cast<long long>(N) * sp;
  const int groups = (c_zero_to - c_off > C ? c_zero_to - c_off : C) + 7 >> 3;
  const long long total = voxels * groups;
  if (C <= 8 && groups > 1)
    launch(pack_cl_skinny_kernel, static_cast<unsigned>((total + 255) / 256), 256, 0, st, x, C, sp, voxels, y, c_pitch,
                                                                                     c_off, groups);
  else if (C > 8 && (voxels + PK_VOX - 1) / PK_VOX < (1LL << 31))
    launch(pack_cl_tiled_kernel, dim3(static_cast<unsigned>((voxels + PK_VOX - 1) / PK_VOX), (groups + 7) / 8), 256, 0, st, 
        x, C, sp, voxels, y, c_pitch, c_off, groups);
  else
    launch(pack_cl_kernel, static_cast<unsigned>((total + 255) / 256), 256, 0, st, x, C, sp, voxels, y, c_pitch, c_off,
                                                                              groups);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_unpack_cl(const __nv_bfloat16* x, int N, int C, long long sp, int c_pitch, int c_off, float* y,
                         cudaStream_t st) {
  const long long voxels = static_cast<long long>(N) * sp;
  const int groups = (C + 7) >> 3;
  const long long total = voxels * groups;
  if (C > 8 && (voxels + PK_VOX - 1) / PK_VOX < (1LL << 31))
    launch(unpack_cl_tiled_kernel, dim3(static_cast<unsigned>((voxels + PK_VOX - 1) / PK_VOX), (groups + 7) / 8), 256, 0, st, 
        x, C, sp, voxels, c_pitch, c_off, y, groups);
  else
    launch(unpack_cl_kernel, static_cast<unsigned>((total + 255) / 256), 256, 0, st, x, C, sp, voxels, c_pitch, c_off, y,
                                                                                groups);
  LAUNCH_CHECK();
  return cudaSuccess;
}

static float axis_scale(int n_in, int n_out, int align) {
  if (align) return n_out > 1 ? static_cast<float>(n_in - 1) / static_cast<float>(n_out - 1) : 0.f;
  return static_cast<float>(n_in) / static_cast<float>(n_out);
}
static ResizeGeom make_geom(int Ti, int Hi, int Wi, int To, int Ho, int Wo, int align) {
  ResizeGeom g;
  g.Ti = Ti; g.Hi = Hi; g.Wi = Wi; g.To = To; g.Ho = Ho; g.Wo = Wo;
  g.st = axis_scale(Ti, To, align);
  g.sh = axis_scale(Hi, Ho, align);
  g.sw = axis_scale(Wi, Wo, align);
  g.align = align;
  return g;
}

void ew_linear_taps_host(int n_in, int n_out, int align, int32_t* i0, int32_t* i1, float* l0, float* l1) {
  const float s = axis_scale(n_in, n_out, align);
  for (int o = 0; o < n_out; ++o) {
    const Tap t = linear_tap(o, n_in, s, align);
    i0[o] = t.i0; i1[o] = t.i1; l0[o] = t.l0; l1[o] = t.l1;
  }
}
cudaError_t ew_linear_taps_dev(int n_in, int n_out, int align, int32_t* i0, int32_t* i1, float* l0, float* l1,
                               cudaStream_t st) {
  launch(linear_taps_kernel, (n_out + 127) / 128, 128, 0, st, n_in, n_out, axis_scale(n_in, n_out, align), align, i0, i1,
                                                         l0, l1);
  LAUNCH_CHECK();
  return cudaSuccess;
}
// Tile planning for the shared-memory-staged resize kernels.  Forward: `band` output rows need at most
// floor((band-1)*sh) + 3 source rows.  Returns false when even a 1-row band does not fit (-> generic kernel).
constexpr size_t RS_SMEM_BUDGET = 72 * 1024;   // backward: three CTAs per SM
constexpr size_t RS_SMEM_FWD = 110 * 1024;     // forward (two staged tiles): two CTAs per SM
static void plan_threads(int Wo, int max_ny, TiledGeom* tg) {
  tg->cpt = (Wo + 255) / 256;
  tg->nx = (Wo + tg->cpt - 1) / tg->cpt;
  int ny = 512 / tg->nx;
  if (ny < 1) ny = 1;
  if (ny > max_ny) ny = max_ny;
  tg->ny = ny;
}
// the register-only forward needs Ti <= 8 frames, To <= 64 taps, int-sized slices, an exact magic division of the
// column index; i0 must be non-decreasing along T (always true for a linear map) — make_twalk() verifies it
static bool colwalk_ok(const ResizeGeom& g, long long slices, int C = 1) {
  return g.Ti <= CW_MAX_TI && g.To <= 64 && slices <= 65535 &&
         static_cast<long long>(C) * g.To * g.Ho * g.Wo < (1LL << 31) &&
         static_cast<long long>(C) * g.Ti * g.Hi * g.Wi < (1LL << 31) &&
         static_cast<long long>(g.Ho) * g.Wo * g.Wo < (1LL << 32);
}
static bool make_twalk(const ResizeGeom& g, TWalk* w) {
  int to = 0;
  for (int f = 0; f < CW_MAX_TI; ++f) {
    int k = 0;
    for (int j = 0; j < CW_MAX_PER_FRAME; ++j) w->l0[f][j] = w->l1[f][j] = 0.f;
    while (f < g.Ti && to < g.To) {
      const Tap t = linear_tap(to, g.Ti, g.st, g.align);
      if (t.i0 != f) break;
      if (k == CW_MAX_PER_FRAME) return false;     // > 4x temporal up-sampling: the tiled kernel handles it
      w->l0[f][k] = t.l0;
      w->l1[f][k] = t.l1;
      ++k;
      ++to;
    }
    w->n[f] = k;
  }
  w->col_magic = g.Wo > 1 ? static_cast<unsigned>(((1ULL << 32) + g.Wo - 1) / g.Wo) : 0u;
  return to == g.To;
}
static bool plan_fwd_tiles(const ResizeGeom& g, TiledGeom* tg, size_t* smem) {
  for (int band = 16; band >= 1; --band) {
    const int n_hs = static_cast<int>((band - 1) * (g.sh > 0.f ? g.sh : 0.f)) + 3;
    const size_t bytes = (static_cast<size_t>(g.Ti) * (n_hs + band) * g.Wo + 2 * static_cast<size_t>(g.Wo)) * 4 +
                         (static_cast<size_t>(g.To) + band) * 16;
    if (bytes <= RS_SMEM_FWD) {
      tg->g = g; tg->band = band; tg->band_log2 = 0; tg->n_hs = n_hs;
      plan_threads(g.Wo, 64, tg);
      *smem = bytes;
      return true;
    }
  }
  return false;
}
static bool plan_bwd_tiles(const ResizeGeom& g, TiledGeom* tg, size_t* smem) {
  for (int lg = 3; lg >= 0; --lg) {
    const int band = 1 << lg;
    const size_t bytes = (static_cast<size_t>(g.Ti) * band * g.Wo + 2 * static_cast<size_t>(g.Wo) + g.Wi + g.Hi + 4) * 4 +
                         (static_cast<size_t>(g.Ho) + g.To) * 16;
    if (bytes <= RS_SMEM_BUDGET) {
      tg->g = g; tg->band = band; tg->band_log2 = lg; tg->n_hs = 0;
      plan_threads(g.Wo, band, tg);       // thread rows own source rows of the band
      *smem = bytes;
      return true;
    }
  }
  return false;
}
template <typename K>
static cudaError_t rs_allow_smem(K kernel, bool* done) {
  if (*done) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(RS_SMEM_FWD));
  if (e == cudaSuccess) *done = true;
  return e;
}

cudaError_t ew_resize3d_fwd(const float* x, long long NC, int Ti, int Hi, int Wi, float* y, int To, int Ho, int Wo,
                            int align, cudaStream_t st) {
  const ResizeGeom g = make_geom(Ti, Hi, Wi, To, Ho, Wo, align);
  TiledGeom tg;
  size_t smem;
  TWalk tw;
  if (colwalk_ok(g, NC) && make_twalk(g, &tw)) {
    const dim3 grid((Ho * Wo + CW_THREADS - 1) / CW_THREADS, static_cast<unsigned>(NC));
#define HPVG_CWF(T_) launch(resize3d_fwd_colwalk_kernel<T_>, grid, CW_THREADS, 0, st, x, g, tw, y)
    switch (Ti) {   // the frame count is a template parameter: no per-frame predicates in the unrolled body
      case 1: HPVG_CWF(1); break;
      case 2: HPVG_CWF(2); break;
      case 3: HPVG_CWF(3); break;
      case 4: HPVG_CWF(4); break;
      case 5: HPVG_CWF(5); break;
      case 6: HPVG_CWF(6); break;
      case 7: HPVG_CWF(7); break;
      default: HPVG_CWF(8); break;
    }
#undef HPVG_CWF
  } else if (NC <= 65535 && plan_fwd_tiles(g, &tg, &smem)) {
    static bool ok = false;
    cudaError_t e = rs_allow_smem(resize3d_fwd_tiled_kernel, &ok);
    if (e != cudaSuccess) return e;
    launch(resize3d_fwd_tiled_kernel, dim3((Ho + tg.band - 1) / tg.band, static_cast<unsigned>(NC)), dim3(tg.nx, tg.ny), smem, st, 
        x, tg, y);
  } else {
    launch(resize3d_fwd_kernel, grid_for(NC * To * Ho * Wo, 256), 256, 0, st, x, NC, g, y);
  }
  LAUNCH_CHECK();
  return cudaSuccess;
}
// window extent of a backward block (max over block rows / columns), from the same inverse-tap function the kernel uses
static bool plan_bwd_window(const ResizeGeom& g, BwdWin* win, size_t* bytes) {
  win->rows = win->cols = 1;
  for (int h0 = 0; h0 < g.Hi; h0 += BW_TY) {
    const int last = (h0 + BW_TY - 1 < g.Hi) ? h0 + BW_TY - 1 : g.Hi - 1;
    const InvTap a = make_inv_tap(h0, g.Hi, g.Ho, g.sh, g.align), b = make_inv_tap(last, g.Hi, g.Ho, g.sh, g.align);
    const int hi = (b.lo + b.n < g.Ho) ? b.lo + b.n : g.Ho;
    if (hi - a.lo > win->rows) win->rows = hi - a.lo;
  }
  for (int w0 = 0; w0 < g.Wi; w0 += BW_TX) {
    const int last = (w0 + BW_TX - 1 < g.Wi) ? w0 + BW_TX - 1 : g.Wi - 1;
    const InvTap a = make_inv_tap(w0, g.Wi, g.Wo, g.sw, g.align), b = make_inv_tap(last, g.Wi, g.Wo, g.sw, g.align);
    const int hi = (b.lo + b.n < g.Wo) ? b.lo + b.n : g.Wo;
    if (hi - a.lo > win->cols) win->cols = hi - a.lo;
  }
  win->cols |= 1;   // odd row pitch: the ~1.26-strided window reads spread over the banks
  *bytes = static_cast<size_t>(g.Ti) * win->rows * win->cols * sizeof(float);
  return *bytes <= 100 * 1024;
}

cudaError_t ew_resize3d_bwd(const float* gy, long long NC, int To, int Ho, int Wo, float* gx, int Ti, int Hi, int Wi,
                            int align, cudaStream_t st) {
  const ResizeGeom g = make_geom(Ti, Hi, Wi, To, Ho, Wo, align);
  TiledGeom tg;
  size_t smem;
  // register-only adjoint: needs <= 4 contributing outputs per source row / column (floor(2/s) + 1 <= 4), To <= 64
  const bool few_h = (Ho == 1) || (g.sh > 0.f && static_cast<int>(2.0f / g.sh) + 1 <= 4);
  const bool few_w = (Wo == 1) || (g.sw > 0.f && static_cast<int>(2.0f / g.sw) + 1 <= 4);
  TWalk tw;
  BwdWin win;
  size_t bw_smem = 0;
  if (few_h && few_w && Ti <= CW_MAX_TI && To <= 64 && NC <= 65535 &&
      static_cast<long long>(To) * Ho * Wo < (1LL << 29) && static_cast<long long>(Ti) * Hi * Wi < (1LL << 31) &&
      make_twalk(g, &tw) && plan_bwd_window(g, &win, &bw_smem)) {
    const dim3 grid((Wi + BW_TX - 1) / BW_TX, (Hi + BW_TY - 1) / BW_TY, static_cast<unsigned>(NC));
    const dim3 block(BW_TX, BW_TY);
    const int nh = (Ho == 1) ? 1 : static_cast<int>(2.0f / g.sh) + 1, nw = (Wo == 1) ? 1 : static_cast<int>(2.0f / g.sw) + 1;
    int max_n = 0;
    for (int f = 0; f < CW_MAX_TI; ++f) max_n = tw.n[f] > max_n ? tw.n[f] : max_n;
#define HPVG_BW(H_, W_, T_, K_)                                                                                    \
  {                                                                                                                \
    static bool ok_ = false;                                                                                       \
    if (!ok_) {                                                                                                    \
      cudaError_t e_ = cudaFuncSetAttribute(resize3d_bwd_colwalk_kernel<H_, W_, T_, K_>,                           \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);              \
      if (e_ != cudaSuccess) return e_;                                                                            \
      ok_ = true;                                                                                                  \
    }                                                                                                              \
    launch(resize3d_bwd_colwalk_kernel<H_, W_, T_, K_>, grid, block, bw_smem, st, gy, g, tw, win, gx);                 \
  }
#define HPVG_BWT(H_, W_, K_)                    \
    switch (Ti) {                                 \
      case 1: HPVG_BW(H_, W_, 1, K_); break;      \
      case 2: HPVG_BW(H_, W_, 2, K_); break;      \
      case 3: HPVG_BW(H_, W_, 3, K_); break;      \
      case 4: HPVG_BW(H_, W_, 4, K_); break;      \
      case 5: HPVG_BW(H_, W_, 5, K_); break;      \
      case 6: HPVG_BW(H_, W_, 6, K_); break;      \
      case 7: HPVG_BW(H_, W_, 7, K_); break;      \
      default: HPVG_BW(H_, W_, 8, K_); break;     \
    }
    if (nh <= 3 && nw <= 3) {
      if (max_n <= 2) { HPVG_BWT(3, 3, 2) } else { HPVG_BWT(3, 3, 4) }
    } else {
      if (max_n <= 2) { HPVG_BWT(4, 4, 2) } else { HPVG_BWT(4, 4, 4) }
    }
#undef HPVG_BWT
#undef HPVG_BW
  } else if (NC <= 65535 && plan_bwd_tiles(g, &tg, &smem)) {
    static bool ok = false;
    cudaError_t e = rs_allow_smem(resize3d_bwd_tiled_kernel, &ok);
    if (e != cudaSuccess) return e;
    launch(resize3d_bwd_tiled_kernel, dim3((Hi + tg.band - 1) / tg.band, static_cast<unsigned>(NC)), dim3(tg.nx, tg.ny), smem, st, 
        gy, tg, gx);
  } else {
    launch(resize3d_bwd_kernel, grid_for(NC * Ti * Hi * Wi, 128), 128, 0, st, gy, NC, g, gx);
  }
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_upsample_noise_pack(const float* x, int N, int C, int Ti, int Hi, int Wi, int To, int Ho, int Wo,
                                   const float* noise, float amp, unsigned long long seed,
                                   unsigned long long sample_base, const unsigned long long* d_sample_offset,
                                   float* up, __nv_bfloat16* xin, int xin_f32, cudaStream_t st) {
  const ResizeGeom g = make_geom(Ti, Hi, Wi, To, Ho, Wo, 1);
  TiledGeom tg;
  size_t smem;
  TWalk tw;
  if ((C == 1 || C == 3) && colwalk_ok(g, N, C) && make_twalk(g, &tw)) {
    const dim3 grid((Ho * Wo + CW_THREADS - 1) / CW_THREADS, N);
#define HPVG_CW(C_, T_) launch(upsample_noise_pack_colwalk_kernel<C_, T_>, grid, CW_THREADS, 0, st,  \
      x, g, tw, noise, amp, seed, sample_base, d_sample_offset, up, xin, xin_f32)
#define HPVG_CWT(C_)                       \
    switch (Ti) {                            \
      case 1: HPVG_CW(C_, 1); break;         \
      case 2: HPVG_CW(C_, 2); break;         \
      case 3: HPVG_CW(C_, 3); break;         \
      case 4: HPVG_CW(C_, 4); break;         \
      case 5: HPVG_CW(C_, 5); break;         \
      case 6: HPVG_CW(C_, 6); break;         \
      case 7: HPVG_CW(C_, 7); break;         \
      default: HPVG_CW(C_, 8); break;        \
    }
    if (C == 1) { HPVG_CWT(1) } else { HPVG_CWT(3) }
#undef HPVG_CWT
#undef HPVG_CW
  } else if (N <= 65535 && plan_fwd_tiles(g, &tg, &smem)) {
    static bool ok = false;
    cudaError_t e = rs_allow_smem(upsample_noise_pack_tiled_kernel, &ok);
    if (e != cudaSuccess) return e;
    launch(upsample_noise_pack_tiled_kernel, dim3((Ho + tg.band - 1) / tg.band, N), dim3(tg.nx, tg.ny), smem, st, 
        x, C, tg, noise, amp, seed, sample_base, d_sample_offset, up, xin, xin_f32);
  } else {
    launch(upsample_noise_pack_kernel, grid_for(static_cast<long long>(N) * To * Ho * Wo, 256), 256, 0, st, 
        x, N, C, g, noise, amp, seed, sample_base, d_sample_offset, up, xin, xin_f32);
  }
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_bn_stats_cl(const __nv_bfloat16* y, long long voxels, double* sum, double* sumsq, DetScratch det,
                           cudaStream_t st) {
  launch(bn_stats_cl_kernel, grid_for(voxels, 128, DET_STREAM_BLOCKS), 256, 0, st, y, voxels, sum, sumsq, det);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_bn_finalize(const double* sum, const double* sumsq, long long count, const float* gamma,
                           const float* beta, float eps, float momentum, float* mm, float* mv, float* scale,
                           float* shift, float* mean, float* invstd, cudaStream_t st) {
  launch(bn_finalize_kernel, 1, 64, 0, st, sum, sumsq, static_cast<double>(count), gamma, beta, eps, momentum, mm, mv,
                                       scale, shift, mean, invstd);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_bn_apply_cl(const __nv_bfloat16* y, long long voxels, const float* scale, const float* shift, int act,
                           __nv_bfloat16* x, cudaStream_t st) {
  launch(bn_apply_cl_kernel, grid_for(voxels * 8, 256), 256, 0, st, y, voxels * 8, scale, shift, act, x);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_sn_power_iter(const float* w, int cout, int k, float* u, float* v, float* sigma, float* inv_sigma,
                             cudaStream_t st) {
  const size_t smem = (2 * static_cast<size_t>(cout) + k) * sizeof(float);
  launch(sn_power_iter_kernel, 1, 512, smem, st, w, cout, k, u, v, sigma, inv_sigma);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_sn_power_iter_multi(const SnTable& tab, int n_layers, cudaStream_t st) {
  size_t smem = 0;
  for (int i = 0; i < n_layers; ++i) {
    const size_t b = (2 * static_cast<size_t>(tab.cout[i]) + tab.k[i]) * sizeof(float);
    if (b > smem) smem = b;
  }
  launch(sn_power_iter_multi_kernel, n_layers, 512, smem, st, tab);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_bn_moving_update_multi(const BnMovingTable& tab, int n_layers, float eps, float momentum,
                                      cudaStream_t st) {
  launch(bn_moving_update_multi_kernel, n_layers, 64, 0, st, tab, eps, momentum);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_bn_center_multi(const BnCenterTable& tab, int n_layers, float eps, cudaStream_t st) {
  launch(bn_center_multi_kernel, n_layers, 64, 0, st, tab, eps);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_bn_train_apply_cl(const __nv_bfloat16* y, long long voxels, const double* sums, const float* gamma,
                                 const float* beta, float eps, float momentum, float* mm, float* mv, float* saved,
                                 int act, __nv_bfloat16* x, const float* center, cudaStream_t st) {
  launch(bn_train_apply_cl_kernel, grid_for(voxels * 8, 256), 256, 0, st, y, voxels * 8, sums,
                                                                      static_cast<double>(voxels), gamma, beta, eps,
                                                                      momentum, mm, mv, saved, act, x, center);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_bn_fold_eval(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                            const float* bias, int C, float* scale, float* shift, cudaStream_t st) {
  launch(bn_fold_eval_kernel, (C + 63) / 64, 64, 0, st, gamma, beta, mean, var, eps, bias, C, scale, shift);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_affine_from_bias(const float* bias, const float* inv_sigma, int C, float* scale, float* shift,
                                cudaStream_t st) {
  launch(affine_from_bias_kernel, (C + 63) / 64, 64, 0, st, bias, inv_sigma, C, scale, shift);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_reduce(int op, const float* a, const float* b, long long n, float* out, DetScratch det,
                      cudaStream_t st) {
  const int grid = grid_for(n, 256, 148 * 4);
  const float inv = 1.0f / static_cast<float>(n);
  if (op == 0) launch(reduce_kernel<0>, grid, 256, 0, st, a, b, n, inv, out, det);
  else if (op == 1) launch(reduce_kernel<1>, grid, 256, 0, st, a, b, n, inv, out, det);
  else launch(reduce_kernel<2>, grid, 256, 0, st, a, b, n, inv, out, det);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_reparam(const float* mu, const float* lv, const float* eps, long long n, float* z, cudaStream_t st) {
  launch(reparam_kernel, grid_for(n, 256), 256, 0, st, mu, lv, eps, n, z);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_reparam_bwd(const float* gz, const float* eps, const float* lv, long long n, float* gmu, float* glv,
                           cudaStream_t st) {
  launch(reparam_bwd_kernel, grid_for(n, 256), 256, 0, st, gz, eps, lv, n, gmu, glv);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_adam_clip(const AdamTable& tab, int n_tensors, float* norms_scratch, float beta1, float beta2,
                         float eps, float bias_corr, float clip, const unsigned long long* d_step, cudaStream_t st,
                         const float* d_hyper) {
  // blocks per tensor: enough to cover the largest tensor with ~2 float4 per thread, and to fill the 148 SMs
  long long nmax = 1;
  for (int i = 0; i < n_tensors; ++i) nmax = tab.n[i] > nmax ? tab.n[i] : nmax;
  long long want = (nmax + 2047) / 2048;
  const long long cap = (148LL * 16 + n_tensors - 1) / n_tensors;
  if (want > cap) want = cap;
  const int gx = static_cast<int>(want < 1 ? 1 : want);
  const int nb = gx < ADAM_NORM_BLOCKS ? gx : ADAM_NORM_BLOCKS;
  if (clip > 0.f) {
    launch(adam_norm_kernel, dim3(nb, n_tensors), 256, 0, st, tab, norms_scratch);
    LAUNCH_CHECK();
  }
  launch(adam_apply_kernel, dim3(gx, n_tensors), 256, 0, st, tab, norms_scratch, beta1, beta2, eps, bias_corr, clip,
                                                         d_step, d_hyper, nb);
  LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t ew_lrelu_bwd_cl(const __nv_bfloat16* ga, const __nv_bfloat16* a, long long elems, __nv_bfloat16* gz,
                            cudaStream_t st) {
  launch(lrelu_bwd_cl_kernel, grid_for(elems / 8, 256), 256, 0, st, ga, a, elems / 8, gz);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_bn_bwd_cl(const __nv_bfloat16* ga, const __nv_bfloat16* y, long long voxels, const float* saved, int act,
                         double* sums, DetScratch det, __nv_bfloat16* gy, float* dgamma, float* dbeta, int accumulate,
                         cudaStream_t st) {
  launch(bn_bwd_reduce_cl_kernel, grid_for(voxels, 128, DET_STREAM_BLOCKS), 256, 0, st, ga, y, voxels, saved, act, sums, det);
  LAUNCH_CHECK();
  launch(bn_bwd_apply_cl_kernel, grid_for(voxels * 8, 256), 256, 0, st, ga, y, voxels * 8, saved, act, sums,
                                                                    1.0 / static_cast<double>(voxels), gy, dgamma, dbeta,
                                                                    accumulate);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_colsum_cl(const __nv_bfloat16* g, long long voxels, double* scratch, DetScratch det, float* out,
                         int accumulate, cudaStream_t st) {
  cudaError_t e = ew_bn_stats_cl(g, voxels, scratch, scratch + 64, det, st);
  if (e != cudaSuccess) return e;
  launch(d2f_kernel, 1, 64, 0, st, scratch, 64, 1.f, accumulate, out);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_diff_scale(const float* a, const float* b, long long n, float coef, int accumulate, float* g,
                          cudaStream_t st) {
  launch(diff_scale_kernel, grid_for(n, 256), 256, 0, st, a, b, n, coef, accumulate, g);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_tanh_bwd(const float* g, const float* out, long long n, float* gpre, cudaStream_t st) {
  launch(tanh_bwd_kernel, grid_for(n, 256), 256, 0, st, g, out, n, gpre);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_axpby(float a, const float* x, float b, float* y, long long n, cudaStream_t st) {
  launch(axpby_kernel, grid_for(n, 256), 256, 0, st, a, x, b, y, n);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_gather_strided(const float* src, long long n, long long stride, long long offset, float* dst,
                              cudaStream_t st) {
  launch(gather_strided_kernel, grid_for(n, 256), 256, 0, st, src, n, stride, offset, dst);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_frames_to_clip(const uint8_t* frames, int Hs, int Ws, int bgr, int start, int every, int T, int H, int W,
                              int hflip, float* clip, cudaStream_t st) {
  FrameGeom g;
  g.Hs = Hs; g.Ws = Ws; g.H = H; g.W = W; g.T = T; g.start = start; g.every = every; g.hflip = hflip; g.bgr = bgr;
  g.sx = static_cast<double>(Ws) / W;
  g.sy = static_cast<double>(Hs) / H;
  if ((H + FC_ROWS - 1) / FC_ROWS > 65535 || static_cast<long long>(Hs) * Ws * 3 >= (1LL << 31))
    return cudaErrorInvalidValue;
  launch(frames_to_clip_kernel, dim3((W + 255) / 256, (H + FC_ROWS - 1) / FC_ROWS), 256, 0, st, frames, g, clip);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_randn(float* z, long long n, unsigned long long seed, unsigned long long offset,
                     const unsigned long long* d_offset, cudaStream_t st) {
  launch(randn_kernel, grid_for((n + 3) / 4, 256), 256, 0, st, z, n, seed, offset, d_offset);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_counter_add(unsigned long long* c, unsigned long long inc, cudaStream_t st) {
  launch(counter_add_kernel, 1, 1, 0, st, c, inc);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_fill(float* y, float v, long long n, cudaStream_t st) {
  launch(fill_kernel, grid_for(n, 256), 256, 0, st, y, v, n);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_channel_sum_ncdhw(const float* g, int N, int C, long long sp, int accumulate, double* scratch,
                                 float* out, cudaStream_t st) {
  if (C > 128) return cudaErrorInvalidValue;
  const int nblk = grid_for(sp, 512, 128);        // scratch: [C <= 128][nblk <= 128] doubles (the DetScratch partials)
  launch(channel_sum_ncdhw_kernel, dim3(nblk, C), 512, 0, st, g, N, C, sp, scratch);
  LAUNCH_CHECK();
  launch(channel_sum_final_kernel, (C + 63) / 64, 64, 0, st, scratch, nblk, C, accumulate, out);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_kl_grad(const float* mu, const float* lv, long long n, float coef, float* gmu, float* glv,
                       cudaStream_t st) {
  launch(kl_grad_kernel, grid_for(n, 256), 256, 0, st, mu, lv, n, coef, gmu, glv);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_sn_grad(const float* G, const float* w, const float* u, const float* v, const float* sigma, int cout,
                       int k, int accumulate, float* scratch /*[SN_GRAD_BLOCKS]*/, float* gw, cudaStream_t st) {
  const int n = cout * k;
  launch(sn_grad_dot_kernel, SN_GRAD_BLOCKS, 256, 0, st, G, w, n, scratch);
  LAUNCH_CHECK();
  launch(sn_grad_apply_kernel, grid_for(n, 256, 148 * 2), 256, 0, st, G, scratch, u, v, sigma, cout, k, accumulate, gw);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_lerp(const float* a, const float* b, float alpha, long long n, float* out, cudaStream_t st) {
  launch(lerp_kernel, grid_for(n, 256), 256, 0, st, a, b, alpha, n, out);
  LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t ew_gp_grad(const float* g, int N, int C, long long sp, float lambda, float* Gout, float* gp,
                       DetScratch det, cudaStream_t st) {
  launch(gp_grad_kernel, grid_for(static_cast<long long>(N) * sp, 256, 148 * 4), 256, 0, st, g, N, C, sp, lambda, Gout, gp,
                                                                                         det);
  LAUNCH_CHECK();
  return cudaSuccess;
}

// ----------------------------------------------------------------------------------------------- strided window (+ ReLU)
// out[n][t][ho][wo][:] = act(in[n][t][h0 + ho*sh][w0 + wo*sw][:]) on bf16 channels-last tensors of C channels (C % 8 == 0):
// the crop of a "valid" convolution computed as a zero-padded one, the stride-2 sub-sampling of a strided convolution
// computed at stride 1, and the ReLU of the sinFID feature networks (src/sinFID/inception.py:66-72 Conv2d_1a/2a/2b are
// conv(no pad / stride 2) + BN + ReLU; the conv kernels here are stride-1 "same" convolutions with LeakyReLU epilogues).
__global__ void slice_act_cl_kernel(const uint4* __restrict__ in, int T, int Hi, int Wi, int Ho, int Wo, int h0, int w0,
                                    int sh, int sw, int c8, int relu, long long total, uint4* __restrict__ out) {
  pdl_grid_sync();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % c8);
    long long v = i / c8;
    const int wo = static_cast<int>(v % Wo); v /= Wo;
    const int ho = static_cast<int>(v % Ho); v /= Ho;      // v = n*T + t
    const long long src = ((v * Hi + (h0 + ho * sh)) * Wi + (w0 + wo * sw)) * c8 + g;
    uint4 q = __ldg(in + src);
    if (relu) {
      __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&q);
      const __nv_bfloat162 z = __float2bfloat162_rn(0.f);
#pragma unroll
      for (int e = 0; e < 4; ++e) h2[e] = __hmax2(h2[e], z);
    }
    out[i] = q;
  }
}

cudaError_t ew_slice_act_cl(const __nv_bfloat16* in, int NT, int Hi, int Wi, int C, int Ho, int Wo, int h0, int w0, int sh,
                            int sw, int relu, __nv_bfloat16* out, cudaStream_t st) {
  const int c8 = C / 8;
  const long long total = static_cast<long long>(NT) * Ho * Wo * c8;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch(slice_act_cl_kernel, static_cast<unsigned>(blocks), 256, 0, st, reinterpret_cast<const uint4*>(in), NT, Hi, Wi, Ho, Wo,
                                                                      h0, w0, sh, sw, c8, relu, total,
                                                                      reinterpret_cast<uint4*>(out));
  return cudaGetLastError();
}

// nn.Pad(mode='REFLECT') on a channels-last tensor (reference networks_3d.py:65-68, the bias-free `bn=False` branch of
// ConvBlock3DSN): out[n, t, h, w, :] = in[n, r(t - pt, T), r(h - ph, H), r(w - ph, W), :], r = mirror without repeating
// the edge.  One 16-byte group per thread (any element type: q16 = 16-byte groups per voxel).
__device__ __forceinline__ int reflect_index(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }
__global__ void reflect_pad_cl_kernel(const uint4* __restrict__ in, int T, int H, int W, int q16, int pt, int ph,
                                      long long total, uint4* __restrict__ out) {
  pdl_grid_sync();
  const int To = T + 2 * pt, Ho = H + 2 * ph, Wo = W + 2 * ph;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % q16);
    long long v = i / q16;
    const int wo = static_cast<int>(v % Wo); v /= Wo;
    const int ho = static_cast<int>(v % Ho); v /= Ho;
    const int to = static_cast<int>(v % To); v /= To;      // v = n
    const int t = reflect_index(to - pt, T), h = reflect_index(ho - ph, H), w = reflect_index(wo - ph, W);
    out[i] = __ldg(in + (((v * T + t) * H + h) * W + w) * q16 + g);
  }
}
cudaError_t ew_reflect_pad_cl(const void* in, int N, int T, int H, int W, int voxel_bytes, int pt, int ph, void* out,
                              cudaStream_t st) {
  const int q16 = voxel_bytes / 16;
  const long long total = static_cast<long long>(N) * (T + 2 * pt) * (H + 2 * ph) * (W + 2 * ph) * q16;
  launch(reflect_pad_cl_kernel, grid_for(total, 256), 256, 0, st, reinterpret_cast<const uint4*>(in), T, H, W, q16, pt, ph,
         total, reinterpret_cast<uint4*>(out));
  return cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------- host-drawn noise
// z <- N(0,1) from UNIFORMS drawn on the host: the host draws the random numbers (numpy, counter-based per sample —
// the reference draws its z on the host too, eval_video.py:67), 3.9x cheaper per value as uniforms than as ziggurat
// normals, and this kernel applies Box-Muller in place after the H2D copy: (u[2i], u[2i+1]) -> (z[2i], z[2i+1]).
// u in [0, 1) with 24 bits: 1 - u is exact and in (0, 1], so the logarithm is finite.
__global__ void box_muller_inplace_kernel(float* __restrict__ z, long long n) {
  pdl_grid_sync();
  const long long pairs = (n + 1) >> 1;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < pairs;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const bool two = 2 * i + 1 < n;
    const float u1 = 1.0f - z[2 * i], u2 = two ? z[2 * i + 1] : 0.5f;
    const float r = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    z[2 * i] = r * cs;
    if (two) z[2 * i + 1] = r * sn;
  }
}

cudaError_t ew_box_muller_inplace(float* z, long long n, cudaStream_t st) {
  launch(box_muller_inplace_kernel, grid_for((n + 1) / 2, 256), 256, 0, st, z, n);
  return cudaGetLastError();
}

}  // namespace hpvg
