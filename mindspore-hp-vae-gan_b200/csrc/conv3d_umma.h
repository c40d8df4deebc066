// Internal interface of the tcgen05 implicit-GEMM convolution (see conv3d_umma.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "det_reduce.cuh"

namespace hpvg {

enum ConvMode {
  CONV_MODE_64_64 = 0, CONV_MODE_64_16 = 1, CONV_MODE_8_64 = 2, CONV_MODE_64_T = 3,
  // kind::tf32 variants over fp32 channels-last activations.  Byte for byte the shared-memory images of the bf16
  // variants with half the channels per 128-byte row: 32 -> 64 (a 64 -> 64 layer = two launches, the second adding the
  // first's fp32 partial sums), <= 4 -> 64 (heads; 16-byte voxels) and 32 -> <= 3 (tails; two launches likewise).
  CONV_MODE_T32_64 = 4, CONV_MODE_T4_64 = 5, CONV_MODE_T32_T = 6
};
enum ConvAct { CONV_ACT_NONE = 0, CONV_ACT_LRELU = 1, CONV_ACT_TANH = 2, CONV_ACT_LRELU_MASK = 3 };
enum ConvOut { CONV_OUT_BF16_NDHWC = 0, CONV_OUT_F32_NCDHW = 1, CONV_OUT_F32_RAW = 2, CONV_OUT_F32_NDHWC = 3 };
inline bool conv_mode_is_tf32(int mode) { return mode >= CONV_MODE_T32_64 && mode <= CONV_MODE_T32_T; }

// device-side parameter block (passed as __grid_constant__)
struct ConvParams {
  int N, T, H, W;
  int w_tiles, h_pairs, n_units;
  const uint8_t* wimg;   // [2 ranks][W_BYTES] packed filter bank (conv3d_pack_weights)
  const float* scale;    // [Cout] epilogue multiplier  (gamma/sqrt(var+eps), 1/sigma, or 1)
  const float* shift;    // [Cout] epilogue offset      (bias / folded BN shift)
  int act;               // ConvAct
  int out_mode;          // ConvOut
  void* out;
  int out_pitch;         // channels per voxel of the channels-last (bf16 / fp32) output tensor
  int out_coff;          // first output channel inside that pitch
  DetScratch det;        // stats: block partials + ticket counter of the deterministic cross-CTA sum
  int tma_out;           // bf16 channels-last output through the TMA-store epilogue (tensor map passed to the kernel)
  int cout_real;         // CONV_OUT_F32_NCDHW: number of real output channels (<= 4)
  const float* addend;   // cl out: fp32 [V][64] partial sums added before scale/shift (split-Cin); tf32 RAW out: the same,
                         // added to the raw accumulator (may alias `out`);
                         // NCDHW out: fp32 NCDHW residual added after scale/shift, before the activation
  const void* mask;      // CONV_ACT_LRELU_MASK: bf16 cl tensor (same voxels, 64 channels at `mask`, mask_pitch per voxel)
  int mask_pitch;        //   y = v * LeakyReLU'(mask): the backward of a LeakyReLU fused into the data-gradient conv
  int in_merged;         // head variant: the tensor map has (C, W) merged (densely packed 8-channel input)
  double* stats;         // bf16 out, Cout 64: optional [2][64] fp64 accumulators (+= sum y, sum y^2 over all voxels of
                         // the stored output): the training-mode BatchNorm statistics, fused into the epilogue
};

// host-side launch description
struct ConvLaunch {
  int mode;              // ConvMode
  int N, T, H, W;
  const void* in;        // bf16 (fp32 for the tf32 modes) channels-last; may point at a 64- (32-) channel slice of a
                         // wider tensor
  int in_pitch;          // channels per voxel of the input tensor (64, 128, ... or 8; tf32: 32k or 4)
  const void* wimg;
  const float* scale;
  const float* shift;
  int act, out_mode;
  void* out;
  int out_pitch, out_coff, cout_real;
  const float* addend;
  double* stats;
  DetScratch det{nullptr, nullptr};   // required with stats: the calling stream's reduction scratch
  const void* mask;
  int mask_pitch;
  int max_pairs;         // CTA pairs to launch (<= 74 on a 148-SM B200)
};

// returns nullptr on success, else a static error string
const char* conv3d_umma_launch(const ConvLaunch& L, cudaStream_t stream);
int conv3d_umma_wimg_bytes(int mode);
const char* conv3d_pack_weights(const float* w, int w_cout, int w_cin, int kt, int mode, int transpose_flip,
                                int cout_off, int cout, int cin_off, int cin, void* img, cudaStream_t stream);

// 64 -> (<= 3) tail convolution with the in-plane taps folded into N (conv3d_tail.cu); uses ConvLaunch with
// mode == CONV_MODE_64_T, out_mode == CONV_OUT_F32_NCDHW
// table-driven packing of many filter banks in one launch (bf16 variants); mode < 0: an (1, bias) epilogue-vector pair
struct PackEntry {
  const float* w;      // weights (Cout, Cin, kt, 3, 3) — or the bias vector for an affine entry
  void* img;           // packed image (bf16) — or fp32 [2][64] for an affine entry
  int cout, cin, kt, mode, flip, cin_off, cout_off, w_cout, w_cin, total;
};
constexpr int PACK_MAX_ENTRIES = 64;
struct PackTable {
  PackEntry e[PACK_MAX_ENTRIES];
};
const char* conv3d_pack_weights_multi(const PackEntry* entries, int n, cudaStream_t stream);

const char* conv3d_tail_launch(const ConvLaunch& L, int sm_count, cudaStream_t stream);
int conv3d_tail_wimg_bytes();

// the same over fp32 channels-last operands on kind::tf32 (conv3d_wgrad_tf32.cu); pitches in fp32 channels (>= 32, or a
// narrow tensor of 4k channels)
const char* conv3d_wgrad_tf32_launch(const void* x, int x_pitch, const void* gy, int gy_pitch, int N, int T, int H,
                                     int W, float* dw, int w_cin, int kt, int co_off, int co_n, int ci_off, int ci_n,
                                     int accumulate, float scale, float* workspace, int sm_count, cudaStream_t stream);
size_t conv3d_wgrad_tf32_workspace_bytes(int sm_count);

// weight gradient (conv3d_wgrad.cu)
size_t conv3d_wgrad_workspace_bytes(int sm_count);
const char* conv3d_wgrad_launch(const void* x, int x_pitch, const void* gy, int gy_pitch, int N, int T, int H, int W,
                                float* dw, int w_cin, int kt, int co_off, int co_n, int ci_off, int ci_n, int accumulate,
                                float scale, float* workspace, int sm_count, cudaStream_t stream);

}  // namespace hpvg
