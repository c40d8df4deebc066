/* libhpvg — C ABI of the B200-native HP-VAE-GAN hot path.
 *
 * The reference (SakiRinn/mindspore-hp-vae-gan) has no FFI of its own: its operator surface is MindSpore
 * `nn.Cell`s and one `Primitive` (src/tools/trilinear.py:171-254).  This header is the boundary a maintainer binds
 * instead of the MindSpore library ops listed in SURVEY.md §2.2 — through `ops.Custom(func_type="aot")` (the
 * `Hpvg*` entry points at the bottom have exactly that signature) or through ctypes (everything else).
 * INTEGRATION.md shows both bindings.
 *
 * Conventions
 *   - every pointer named d_* / "device" is a CUDA device pointer owned by the caller; the library never frees it.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  Nothing synchronises inside.
 *   - thread safety: every entry point may be called concurrently from several host threads provided each uses its own
 *     stream (all internal scratch — reduction buffers, weight-gradient partials, AOT staging — is per stream).
 *   - return value: 0 on success, negative HPVG_E_* on failure; hpvg_last_error() gives a thread-local message.
 *   - "cl" tensors are channels-last bf16: (N, T, H, W, Cpitch); "ncdhw" tensors are fp32 (N, C, T, H, W) — the
 *     layout of the reference's Tensors.  2-D (N, C, H, W) data is the T == 1 case.
 */
#ifndef HPVG_H_
#define HPVG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPVG_OK 0
#define HPVG_E_CUDA (-1)
#define HPVG_E_ARG (-2)
#define HPVG_E_UNSUPPORTED (-3)

/* conv kernel variants / epilogues (values mirror csrc/conv3d_umma.h) */
#define HPVG_CONV_64_64 0 /* Cin 64 -> Cout 64                                   */
#define HPVG_CONV_64_16 1 /* Cin 64 -> Cout <= 4 real (tail convs 64->3, 64->1)  */
#define HPVG_CONV_8_64 2  /* Cin <= 8 (zero padded to 8) -> Cout 64 (head convs) */
#define HPVG_CONV_64_3 3  /* Cin 64 -> Cout <= 3, in-plane taps folded into N (fast tail convs; fp32 ncdhw out) */
/* kind::tf32 variants over fp32 channels-last activations ("f32cl": (N,T,H,W,Cpitch) fp32, values rounded to tf32).
 * The reference computes in fp32 (networks_3d.py:48-50 nn.Conv3d on fp32 Tensors); these are the fp32-accurate path
 * (per-layer rel-L2 <= 1e-3).  A 64-channel input is consumed 32 channels per launch: the first launch writes raw fp32
 * partial sums (HPVG_OUT_F32_RAW), the last adds them through d_addend and applies the epilogue. */
#define HPVG_CONV_T32_64 4 /* Cin 32 (a 32-channel slice at d_in) -> Cout 64            */
#define HPVG_CONV_T4_64 5  /* Cin <= 4 (zero padded to 4) -> Cout 64 (head convs)        */
#define HPVG_CONV_T32_3 6  /* Cin 32 -> Cout <= 3 (tail convs; fp32 ncdhw out)           */
#define HPVG_ACT_NONE 0
#define HPVG_ACT_LRELU 1 /* LeakyReLU(0.2): mindspore.nn.LeakyReLU default, networks_3d.py:20 */
#define HPVG_ACT_TANH 2
#define HPVG_ACT_LRELU_MASK 3 /* y = v * LeakyReLU'(mask): LeakyReLU backward fused into a data-gradient conv */
#define HPVG_OUT_BF16_CL 0
#define HPVG_OUT_F32_NCDHW 1
#define HPVG_OUT_F32_RAW 2
#define HPVG_OUT_F32_CL 3 /* fp32 channels-last (tf32 variants only) */

/* ---------------------------------------------------------------- runtime (ctypes route; MindSpore owns these itself) */
int hpvg_version(void);
const char* hpvg_last_error(void);
int hpvg_device_count(void);
int hpvg_init(int device);
int hpvg_sm_count(void);
int hpvg_malloc(void** d_ptr, size_t bytes);
int hpvg_free(void* d_ptr);
int hpvg_host_alloc(void** h_ptr, size_t bytes); /* pinned */
int hpvg_host_free(void* h_ptr);
int hpvg_memset(void* d_ptr, int value, size_t bytes, void* stream);
int hpvg_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream);
int hpvg_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream);
int hpvg_d2d(void* d_dst, const void* d_src, size_t bytes, void* stream);
int hpvg_stream_create(void** stream);
/* Scratch memory is per stream.  hpvg_stream_create registers its stream; a stream created elsewhere (MindSpore's) is
 * registered on first use, or explicitly with hpvg_stream_attach — required before that stream is first used INSIDE a
 * CUDA-graph capture (registration allocates).  hpvg_stream_detach releases the scratch of a foreign stream. */
int hpvg_stream_attach(void* stream);
int hpvg_stream_detach(void* stream);
int hpvg_stream_destroy(void* stream);
int hpvg_stream_sync(void* stream);
int hpvg_device_sync(void);
int hpvg_event_create(void** event);
int hpvg_event_destroy(void* event);
int hpvg_event_record(void* event, void* stream);
int hpvg_event_sync(void* event);
int hpvg_stream_wait_event(void* stream, void* event); /* later work on `stream` waits for `event` */
int hpvg_event_elapsed_ms(void* start, void* stop, float* ms);
/* CUDA-graph capture of everything enqueued on `stream` between begin and end */
int hpvg_graph_begin(void* stream);
int hpvg_graph_end(void* stream, void** graph_exec);
int hpvg_graph_launch(void* graph_exec, void* stream);
int hpvg_graph_destroy(void* graph_exec);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long hpvg_launch_count(void);
/* programmatic dependent launch for every kernel of the library (default on; environment HPVG_PDL=0 turns it off):
   returns the previous setting.  Kernels launched with it start their prologue while their predecessor in the
   stream drains and block in griddepcontrol.wait before touching global memory (csrc/launch.cuh) */
int hpvg_set_pdl(int on);

/* ---------------------------------------------------------------- layout changes at the API edge */
/* fp32 (N,C,T,H,W) -> bf16 (N,T,H,W,c_pitch) channels [c_off, c_off+C); channels [c_off+C, c_off+c_zero_to) := 0 */
int hpvg_pack_cl(const float* d_x, int N, int C, int T, int H, int W, void* d_y, int c_pitch, int c_off,
                 int c_zero_to, void* stream);
/* bf16 (N,T,H,W,c_pitch) channels [c_off, c_off+C) -> fp32 (N,C,T,H,W) */
int hpvg_unpack_cl(const void* d_x, int N, int C, int T, int H, int W, int c_pitch, int c_off, float* d_y,
                   void* stream);

/* the same for fp32 channels-last tensors (tf32 precision mode); c_off / c_pitch multiples of 4 */
int hpvg_pack_cl_f32(const float* d_x, int N, int C, int T, int H, int W, float* d_y, int c_pitch, int c_off,
                     int c_zero_to, void* stream);
int hpvg_unpack_cl_f32(const float* d_x, int N, int C, int T, int H, int W, int c_pitch, int c_off, float* d_y,
                       void* stream);

/* ---------------------------------------------------------------- convolution
 * Replaces nn.Conv3d / nn.Conv2d 3x3(x3), stride 1, zero pad 1 (networks_3d.py:48-50,380,399; networks_2d.py:47-49;
 * spectral_norm.py:152).  Weights are first packed into the kernel's shared-memory image. */
int hpvg_conv_wimg_bytes(int mode);
/* d_w: fp32 (w_cout, w_cin, kt, 3, 3), kt in {1,3}.  Packs the sub-filter [cout_off, +cout) x [cin_off, +cin).
 * transpose_flip = 1 packs the data-gradient filter (roles of cin/cout swapped, taps mirrored). */
int hpvg_conv_pack_weights(const float* d_w, int w_cout, int w_cin, int kt, int mode, int transpose_flip,
                           int cout_off, int cout, int cin_off, int cin, void* d_wimg, void* stream);
/* n banks in ONE launch (arrays of length n on the host; bf16 variants only).  mode[i] < 0 marks an epilogue-vector
 * entry instead of a filter bank: d_w[i] = bias (cout[i] values, nullable), d_wimg[i] = fp32 [2][64] := (1, bias). */
int hpvg_conv_pack_weights_multi(int n, const float* const* d_w, const int* w_cout, const int* w_cin, const int* kt,
                                 const int* mode, const int* transpose_flip, const int* cout_off, const int* cout,
                                 const int* cin_off, const int* cin, void* const* d_wimg, void* stream);
/* y = act((conv(x) [+ addend_raw]) * scale + shift [+ residual]) ; see HPVG_OUT_* for the output layouts.
 *  d_in     bf16 cl, in_pitch channels per voxel (the kernel reads 64 — or 8 — channels starting at d_in)
 *  d_scale/d_shift fp32 [Cout]: bias, folded BatchNorm, 1/sigma of spectral norm
 *  d_addend HPVG_OUT_BF16_CL: optional fp32 [voxels][64] partial sums (split-Cin accumulation)
 *           HPVG_OUT_F32_NCDHW: optional fp32 ncdhw residual (tanh(block(x)+up), networks_3d.py:450)
 *  d_stats  optional fp64 [2][64] (HPVG_OUT_BF16_CL, 64 output channels): += per-channel sum / sum of squares of the
 *           stored output — the batch statistics of a following training-mode BatchNorm3d (networks_3d.py:52),
 *           fused into the conv epilogue.  The caller zeroes it.
 *  d_mask   HPVG_ACT_LRELU_MASK only: bf16 cl tensor of the same voxels (64 channels at d_mask, mask_pitch channels
 *           per voxel) holding a stored LeakyReLU activation a; the epilogue applies v * (a > 0 ? 1 : 0.2), i.e. the
 *           backward of that LeakyReLU, so the data-gradient conv emits the next layer's pre-activation gradient.    */
int hpvg_conv_cl(int mode, int N, int T, int H, int W, const void* d_in, int in_pitch, const void* d_wimg,
                 const float* d_scale, const float* d_shift, int act, int out_mode, void* d_out, int out_pitch,
                 int out_coff, int cout_real, const float* d_addend, double* d_stats, const void* d_mask,
                 int mask_pitch, void* stream);

/* ---------------------------------------------------------------- linear resize
 * Replaces UpsampleTrilinear3D(output_size, align_corners) (src/tools/trilinear.py:171-254, called from
 * src/utils/images.py:54-61) and ops.ResizeBilinear (images.py:41).  Index / weight tables are computed on the host
 * in IEEE fp32 with the reference's rule and are exposed for bit-exact checking. */
int hpvg_linear_taps(int n_in, int n_out, int align_corners, int32_t* i0, int32_t* i1, float* l0, float* l1);
/* the same tables computed by the device code path (test hook: device arrays of n_out entries) */
int hpvg_linear_taps_dev(int n_in, int n_out, int align_corners, int32_t* d_i0, int32_t* d_i1, float* d_l0, float* d_l1,
                         void* stream);
int hpvg_resize3d_fwd(const float* d_x, int N, int C, int Ti, int Hi, int Wi, float* d_y, int To, int Ho, int Wo,
                      int align_corners, void* stream);
int hpvg_resize3d_bwd(const float* d_gy, int N, int C, int To, int Ho, int Wo, float* d_gx, int Ti, int Hi, int Wi,
                      int align_corners, void* stream);
/* fused block input stage (networks_3d.py:440-446): up = resize(x_prev); x_in = up + noise*amp;
 * writes up (fp32 ncdhw, the residual) and x_in (bf16 cl, 8-channel padded).  d_noise may be NULL (no noise), or
 * noise_seed != 0 generates N(0,1) on the device with Philox keyed by (seed, sample index, element). */
int hpvg_upsample_noise_pack(const float* d_x, int N, int C, int Ti, int Hi, int Wi, int To, int Ho, int Wo,
                             const float* d_noise, float amp, uint64_t noise_seed, uint64_t sample_base,
                             const uint64_t* d_sample_offset /* nullable: device draw counter added to sample_base */,
                             float* d_up, void* d_xin_cl, void* stream);
/* tf32 precision mode: x_in is fp32 channels-last, 4 channels (16 bytes) per voxel */
int hpvg_upsample_noise_pack_f32(const float* d_x, int N, int C, int Ti, int Hi, int Wi, int To, int Ho, int Wo,
                                 const float* d_noise, float amp, uint64_t noise_seed, uint64_t sample_base,
                                 const uint64_t* d_sample_offset, float* d_up, float* d_xin_cl, void* stream);
/* Device-side data path between the video decoder and the network.  d_frames: uint8 [F][Hs][Ws][3]
 * decoded frames (bgr != 0: decoder order, as cv2.VideoCapture returns them).  Writes the fp32 clip [1][3][T][H][W] the
 * reference's SingleVideoDataset.__getitem__ yields for window `start` and rate `every`:
 *   generate_frames.py:42-46  cv2.cvtColor(BGR2RGB) + cv2.resize(INTER_LINEAR) on uint8 (restated bit-exactly)
 *   video.py:52-59            frames[start : start + lcm + 1 : every], float32 / 255
 *   video.py:75-86            optional horizontal flip, Normalize(mean .5, std .5), (C, T, H, W) */
int hpvg_frames_to_clip(const uint8_t* d_frames, int F, int Hs, int Ws, int bgr, int start, int every, int T, int H,
                        int W, int hflip, float* d_clip, void* stream);
/* z ~ N(0,1) on the device (Philox4x32-10 + Box-Muller), keyed by (seed, offset [+ *d_offset], element): stand-in for
 * the reference's host numpy draws (images.py:17-21, networks_3d.py:28-34) inside CUDA-graph replays */
/* z <- N(0,1) in place from host-drawn uniforms u in [0,1) (pairs (u[2i], u[2i+1]) -> Box-Muller): the device half of
 * the sampling pipeline's host noise (utils.generate_noise_ref, eval_video.py:67, drawn on the host there too). */
int hpvg_box_muller_inplace(float* d_z, long long n, void* stream);
int hpvg_randn(float* d_z, long long n, uint64_t seed, uint64_t offset, const uint64_t* d_offset, void* stream);
int hpvg_counter_add(uint64_t* d_counter, uint64_t inc, void* stream);

/* ---------------------------------------------------------------- BatchNorm (training mode), channels-last bf16, C == 64
 * Replaces nn.BatchNorm3d in set_train() mode (networks_3d.py:52): batch mean / biased variance over N*T*H*W. */
int hpvg_bn_stats_cl(const void* d_y, long long voxels, double* d_sum /*[64]*/, double* d_sumsq /*[64]*/,
                     void* stream);
/* scale = gamma/sqrt(var+eps), shift = beta - mean*scale; also updates moving stats (momentum 0.9) */
int hpvg_bn_finalize(const double* d_sum, const double* d_sumsq, long long count, const float* d_gamma,
                     const float* d_beta, float eps, float momentum, float* d_moving_mean, float* d_moving_var,
                     float* d_scale, float* d_shift, float* d_mean, float* d_invstd, void* stream);
/* x = lrelu(y*scale + shift), elementwise on (voxels, 64) bf16 */
int hpvg_bn_apply_lrelu_cl(const void* d_y, long long voxels, const float* d_scale, const float* d_shift, int act,
                           void* d_x, void* stream);

/* one-pass training-mode BatchNorm: d_sums = the fp64 [2][64] statistics produced by hpvg_conv_cl(d_stats=...);
 * finalises (scale, shift), updates the moving statistics, writes d_saved = (scale, shift, mean, invstd) [4][64]
 * (nullable) and x = act(y*scale + shift) in a single launch.  d_center (nullable): y was stored minus this per-channel
 * offset (hpvg_bn_center_multi); d_saved stays in that centred frame, the moving mean gets the offset added back. */
int hpvg_bn_train_apply_cl(const void* d_y, long long voxels, const double* d_sums, const float* d_gamma,
                           const float* d_beta, float eps, float momentum, float* d_moving_mean, float* d_moving_var,
                           float* d_saved, int act, void* d_x, const float* d_center, void* stream);

/* deferred moving-statistics update of many BatchNorm layers in one launch (pointer arrays on the host): pass
 * d_moving_mean = d_moving_var = NULL to hpvg_bn_train_apply_cl, keep its d_saved, and replay the updates later in the
 * reference's order — needed when training-mode forwards of one network run concurrently on several streams */
int hpvg_bn_moving_update_multi(int n_layers, const float* const* d_saved, float* const* d_moving_mean,
                                float* const* d_moving_var, const float* const* d_center /* nullable (array / entries) */,
                                float eps, float momentum, void* stream);
/* Kink-centred bf16 storage of the pre-BatchNorm activation (training mode): for every layer i, from the layer's gamma,
 * beta, moving mean / variance and conv bias (64 floats each) compute center = moving_mean - beta*sqrt(moving_var+eps)/gamma
 * — the estimated position of the LeakyReLU kink in units of the conv output — and the conv epilogue vectors
 * aff[2][64] = (1, bias - center).  The conv then stores y - center (bf16 resolution is finest where the mask of the
 * backward is decided); hpvg_bn_train_apply_cl / hpvg_bn_moving_update_multi take d_center to keep the MOVING mean that
 * of the un-centred output.  BatchNorm being shift invariant, nothing else changes. */
int hpvg_bn_center_multi(int n_layers, const float* const* d_gamma, const float* const* d_beta,
                         const float* const* d_moving_mean, const float* const* d_moving_var,
                         const float* const* d_bias, float* const* d_center, float* const* d_aff, float eps,
                         void* stream);

/* ---------------------------------------------------------------- spectral norm (spectral_norm.py:142-151)
 * One power iteration on W viewed (Cout, K): v <- l2n(W^T u); u <- l2n(W v); sigma = u^T W v.
 * Updates d_u, d_v in place and writes sigma and 1/sigma. */
int hpvg_sn_power_iter(const float* d_w, int cout, int k, float* d_u, float* d_v, float* d_sigma,
                       float* d_inv_sigma, void* stream);

/* the same for every spectrally normalised layer of a network in ONE launch (arrays live on the host, <= 16 layers);
 * d_sigma2[i] -> 2 floats (sigma, 1/sigma); d_aff[i] (nullable array / entries) -> [2][64] conv epilogue vectors
 * (scale = 1/sigma, shift = bias); d_u_copy / d_v_copy (nullable) receive a snapshot of the updated u / v, which
 * the backward of THIS pass needs because u, v advance on every forward (spectral_norm.py:146-148). */
int hpvg_sn_power_iter_multi(int n_layers, const float* const* d_w, const int* cout, const int* k, float* const* d_u,
                             float* const* d_v, float* const* d_sigma2, const float* const* d_bias,
                             float* const* d_aff, float* const* d_u_copy, float* const* d_v_copy, void* stream);

/* ---------------------------------------------------------------- small fused elementwise / reductions */
/* scale[c] = a[c]*mul ; shift[c] = b[c]  helpers for epilogue vectors */
int hpvg_bn_fold_eval(const float* d_gamma, const float* d_beta, const float* d_mean, const float* d_var, float eps,
                      const float* d_bias, int C, float* d_scale, float* d_shift, void* stream);
int hpvg_affine_from_bias(const float* d_bias, const float* d_inv_sigma /*nullable*/, int C, float* d_scale,
                          float* d_shift, void* stream);
/* losses (src/modules/losses.py:5-7, train_video.py:163,339): results are fp32 scalars on the device */
int hpvg_mse(const float* d_a, const float* d_b, long long n, float* d_out, void* stream);
int hpvg_mean(const float* d_a, long long n, float* d_out, void* stream);
int hpvg_kl(const float* d_mu, const float* d_logvar, long long n, float* d_out, void* stream);
/* z = eps*exp(0.5*logvar)+mu (networks_3d.py:415-417) */
int hpvg_reparam(const float* d_mu, const float* d_logvar, const float* d_eps, long long n, float* d_z, void* stream);
/* its backward: d_gmu += gz ; d_glogvar += gz * eps * 0.5 * exp(0.5*logvar) */
int hpvg_reparam_bwd(const float* d_gz, const float* d_eps, const float* d_logvar, long long n, float* d_gmu,
                     float* d_glogvar, void* stream);

/* ---------------------------------------------------------------- optimiser (src/modules/optimizers.py:33-43)
 * Multi-tensor ClipByNorm(clip) + Adam in two launches.  Tensors are described by parallel arrays on the HOST. */
int hpvg_adam_clip_multi(int n_tensors, float* const* d_params, const float* const* d_grads, float* const* d_m,
                         float* const* d_v, const long long* sizes, const float* lrs, float beta1, float beta2,
                         float eps, int step, float clip_norm /* <=0: no clipping */,
                         const uint64_t* d_step /* nullable: 1-based step read on the device instead of `step` */,
                         void* stream);

/* ---------------------------------------------------------------- backward (hand-restated MindSpore autodiff)
 * Data gradient of a conv = hpvg_conv_cl with a filter bank packed with transpose_flip = 1.
 * Weight gradient: dW[(co_off+co)][(ci_off+ci)][tap] (+)= scale * sum_v gy[v][co] * x[v+tap][ci] for the 64x64 channel
 * block starting at d_x / d_gy (bf16 cl).  A pitch >= 64 selects the 64-channel slice at the pointer; a pitch of
 * 8..56 (multiple of 8) is a NARROW operand — the 8-channel block inputs of head convs, the 8-channel output
 * gradients of tail convs — read as it is: all of its channels, the rest of the 64-wide block as zero (TMA
 * out-of-bounds fill), cropped with co_n / ci_n.  d_dw: fp32 (Cout, w_cin, kt, 3, 3). */
int hpvg_conv_wgrad_cl(const void* d_x, int x_pitch, const void* d_gy, int gy_pitch, int N, int T, int H, int W,
                       float* d_dw, int w_cin, int kt, int co_off, int co_n, int ci_off, int ci_n, int accumulate,
                       float scale, void* stream);
/* The weight gradient on kind::tf32 over fp32 channels-last operands (pitches in fp32 channels: >= 64 selects the
 * 64-channel slice at the pointer, 4..28 is a narrow operand read as it is). */
int hpvg_conv_wgrad_cl_tf32(const float* d_x, int x_pitch, const float* d_gy, int gy_pitch, int N, int T, int H, int W,
                            float* d_dw, int w_cin, int kt, int co_off, int co_n, int ci_off, int ci_n, int accumulate,
                            float scale, void* stream);
/* fp32 channels-last twins of the BatchNorm / LeakyReLU / column-sum kernels (C == 64, tf32 precision mode) */
int hpvg_bn_stats_cl_f32(const float* d_y, long long voxels, double* d_sum, double* d_sumsq, void* stream);
int hpvg_bn_apply_lrelu_cl_f32(const float* d_y, long long voxels, const float* d_scale, const float* d_shift, int act,
                               float* d_x, void* stream);
int hpvg_bn_train_apply_cl_f32(const float* d_y, long long voxels, const double* d_sums, const float* d_gamma,
                               const float* d_beta, float eps, float momentum, float* d_moving_mean,
                               float* d_moving_var, float* d_saved, int act, float* d_x, const float* d_center, void* stream);
int hpvg_lrelu_bwd_cl_f32(const float* d_ga, const float* d_a, long long elems, float* d_gz, void* stream);
int hpvg_bn_bwd_cl_f32(const float* d_ga, const float* d_y, long long voxels, const float* d_saved, int act,
                       float* d_gy, float* d_dgamma, float* d_dbeta, int accumulate, void* stream);
int hpvg_colsum_cl_f32(const float* d_g, long long voxels, float* d_out, int accumulate, void* stream);
/* out[n,t,ho,wo,:] = act(in[n,t,h0+ho*sh,w0+wo*sw,:]) on bf16 channels-last tensors (C % 8 == 0; relu != 0 applies ReLU):
 * crop / stride-2 sub-sampling / ReLU of the sinFID feature networks (src/sinFID/inception.py:66-72, c3d.py:63-66),
 * whose convolutions run on the stride-1 zero-padded kernels above. NT = N*T planes of Hi x Wi voxels. */
int hpvg_slice_act_cl(const void* d_in, int NT, int Hi, int Wi, int C, int Ho, int Wo, int h0, int w0, int sh, int sw,
                      int relu, void* d_out, void* stream);
/* nn.Pad(mode='REFLECT') of a channels-last tensor (N,T,H,W,voxel_bytes) by pad_t frames and pad_hw rows / columns on
 * both sides (reference networks_3d.py:65-68, networks_2d.py: the bias-free `bn=False` branch of ConvBlock3DSN / 2DSN):
 * d_out is (N, T+2*pad_t, H+2*pad_hw, W+2*pad_hw, voxel_bytes).  The reference's pad_mode='valid' convolution of the
 * padded tensor is then hpvg_conv_cl on it followed by the interior crop hpvg_slice_act_cl(h0 = w0 = 1). */
int hpvg_reflect_pad_cl(const void* d_in, int N, int T, int H, int W, int voxel_bytes, int pad_t, int pad_hw, void* d_out,
                        void* stream);
/* gz = ga * LeakyReLU'(a), a = stored activation (bf16 cl, elems % 8 == 0) */
int hpvg_lrelu_bwd_cl(const void* d_ga, const void* d_a, long long elems, void* d_gz, void* stream);
/* BatchNorm(train)+act backward on (voxels, 64) bf16: d_saved = (scale, shift, mean, invstd) from the forward
 * (hpvg_bn_finalize); writes gy and (optionally accumulating) dgamma / dbeta. */
int hpvg_bn_bwd_cl(const void* d_ga, const void* d_y, long long voxels, const float* d_saved, int act, void* d_gy,
                   float* d_dgamma, float* d_dbeta, int accumulate, void* stream);
/* out[c] (+)= sum_v g[v][c]  (bias gradient), (voxels, 64) bf16 */
int hpvg_colsum_cl(const void* d_g, long long voxels, float* d_out, int accumulate, void* stream);
/* g (+)= coef*(out - target)  (nn.MSELoss gradient with coef = weight*2/n) */
int hpvg_mse_grad(const float* d_out, const float* d_target, long long n, float coef, int accumulate, float* d_g,
                  void* stream);
int hpvg_tanh_bwd(const float* d_g, const float* d_out, long long n, float* d_gpre, void* stream);
int hpvg_axpby(float a, const float* d_x, float b, float* d_y, long long n, void* stream); /* y = a*x + b*y */
int hpvg_fill(float* d_y, float value, long long n, void* stream);
/* dst[i] = src[i*stride + offset], i < n (e.g. one filter tap of a weight-gradient tensor: the 64x64 second-moment
 * matrix of a feature map for sinFID, src/sinFID/fid_score.py:176-177) */
int hpvg_gather_strided(const float* d_src, long long n, long long stride, long long offset, float* d_dst,
                        void* stream);
int hpvg_channel_sum(const float* d_g, int N, int C, long long spatial, int accumulate, float* d_out, void* stream);
int hpvg_kl_grad(const float* d_mu, const float* d_logvar, long long n, float coef, float* d_gmu, float* d_glogvar,
                 void* stream);
/* spectral-norm chain rule with u, v held constant: gW (+)= (G - <G, W/sigma> u v^T)/sigma */
int hpvg_sn_grad(const float* d_G, const float* d_w, const float* d_u, const float* d_v, const float* d_sigma,
                 int cout, int k, int accumulate, float* d_gw, void* stream);
/* WGAN-GP pieces (src/modules/losses.py:47-52) */
int hpvg_lerp(const float* d_a, const float* d_b, float alpha, long long n, float* d_out, void* stream);
int hpvg_gp_grad(const float* d_g, int N, int C, long long spatial, float lambda, float* d_G, float* d_gp,
                 void* stream);

/* ---------------------------------------------------------------- fused sampling entry (SURVEY.md §8b "granularity")
 * One call = GeneratorHPVAEGAN.construct(noise_init=z, isRandom=True) in eval mode for a batch of N samples
 * (reference networks_3d.py:406-451 as driven by eval_video.py:53-82): decoder block on z, then for every refinement
 * stage  up = trilinear(x_prev) ; x_in = up + noise * amp ; x = tanh(block(x_in) + up).
 * bf16 precision mode.  The description holds what the host prepared once per checkpoint: packed filter banks
 * (hpvg_conv_pack_weights*) and epilogue vectors (hpvg_bn_fold_eval: eval-mode BatchNorm folded into the conv) of
 * every layer.  No allocation, no synchronisation: everything is enqueued on `stream`; activations live in the
 * caller's workspace of hpvg_generator_sample_workspace() bytes. */
#define HPVG_MAX_LEVELS 16
#define HPVG_BLOCK_LAYERS 8
typedef struct {
  int n_layers;                             /* conv layers of the block; the last one is the 64 -> nc_im tail */
  int cin[HPVG_BLOCK_LAYERS];               /* nc_im (head of a refinement block), 64, or 128 (decoder head) */
  const void* wimg[HPVG_BLOCK_LAYERS][2];   /* packed banks; [l][1] is the second 64-channel half of a 128-channel input */
  const float* scale[HPVG_BLOCK_LAYERS];    /* [64] epilogue multiplier (tail: [nc_im]) */
  const float* shift[HPVG_BLOCK_LAYERS];    /* [64] epilogue offset */
} HpvgBlock;
typedef struct {
  int n_stages;                             /* refinement blocks = pyramid levels - 1 */
  int nc_im, latent_dim;
  int T[HPVG_MAX_LEVELS], H[HPVG_MAX_LEVELS], W[HPVG_MAX_LEVELS];   /* level 0 (z, decoder) .. n_stages */
  float noise_amp[HPVG_MAX_LEVELS];         /* level l: amplitude of the refinement noise; 0 = none (VAE levels) */
  uint64_t noise_seed[HPVG_MAX_LEVELS];     /* level l: Philox key of that noise */
  HpvgBlock decoder;
  HpvgBlock body[HPVG_MAX_LEVELS];          /* body[s] produces level s + 1 */
} HpvgGenerator;
/* One block in eval mode (ConvBlock3D x (n_layers - 1) + the 64 -> nc_im tail, networks_3d.py:377-381,395-401):
 * d_out = tanh(block(x) [+ d_residual]).  d_x_cl: bf16 channels-last (N,T,H,W,x_pitch): x_pitch = 8 for a refinement
 * block's padded input, 64 or 128 for the decoder.  Workspace: hpvg_block_fwd_eval_workspace() bytes. */
size_t hpvg_block_fwd_eval_workspace(const HpvgBlock* b, int N, int T, int H, int W);
int hpvg_block_fwd_eval(const HpvgBlock* b, int nc_im, int N, int T, int H, int W, const void* d_x_cl, int x_pitch,
                        const float* d_residual, float* d_out, void* d_workspace, size_t workspace_bytes, void* stream);
size_t hpvg_generator_sample_workspace(const HpvgGenerator* g, int N);
/* d_z: fp32 (N, latent_dim, T0, H0, W0); d_out: fp32 (N, nc_im, T, H, W) of the finest level; d_vae_out: optional fp32
 * (N, nc_im, T0, H0, W0) decoder output; sample_base: index of the batch's first sample (keys the device noise, so a
 * sample's noise does not depend on how samples are batched or sharded); d_sample_offset: optional device counter
 * added to sample_base when the kernels run — the call can then be captured ONCE into a CUDA graph (hpvg_graph_begin /
 * _end) and replayed for every batch with only that counter and d_z rewritten */
int hpvg_generator_sample(const HpvgGenerator* g, const float* d_z, int N, uint64_t sample_base,
                          const uint64_t* d_sample_offset, float* d_out, float* d_vae_out, void* d_workspace,
                          size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- MindSpore ops.Custom(func_type="aot") entry points
 * int Name(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes, void* stream, void* extra)
 * params = fp32 device pointers in the reference's layouts, inputs then outputs, pre-allocated by the framework.
 * No allocation per call and no synchronisation inside: staging buffers live in the calling stream's grow-only
 * workspace (the first call at a new size grows it).  0 = success, anything else makes MindSpore raise.
 * Each entry replaces the MindSpore library op named next to it. */
#define HPVG_AOT(name) \
  int name(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes, void* stream, void* extra)
HPVG_AOT(HpvgUpsampleTrilinear3D);       /* x -> y, align_corners=True            (src/tools/trilinear.py:171-254)  */
HPVG_AOT(HpvgUpsampleTrilinear3DGrad);   /* (dy, x) -> dx                          (its bprop)                       */
HPVG_AOT(HpvgConv3dBias);                /* (x, w, b) -> conv(x, w) + b            (nn.Conv3d, networks_3d.py:48-50)  */
HPVG_AOT(HpvgConv3dBiasLRelu);           /* ... -> LeakyReLU(0.2)                  (ConvBlock3DSN, :57-73)            */
HPVG_AOT(HpvgConv3dBiasTanh);            /* ... -> tanh (Cout <= 3 tails)          (networks_3d.py:399,450)           */
/*   channel shapes: (Cin <= 8 | 64) -> 64 and 64 -> (<= 3); x (N,Cin,T,H,W), w (Cout,Cin,3,3,3), b (Cout)          */
HPVG_AOT(HpvgConv3dBiasLReluGrad);       /* (x, w, y, dy) -> (dx, dw, db), 64 -> 64 (autodiff of the above)          */
HPVG_AOT(HpvgBatchNorm3dLReluTrain);     /* (x, gamma, beta, moving_mean, moving_var) -> (y, saved[4,64]); moving
                                            statistics updated in place           (nn.BatchNorm3d train, :52-53)    */
HPVG_AOT(HpvgBatchNorm3dLReluTrainGrad); /* (dy, x, saved) -> (dx, dgamma, dbeta)                                    */
HPVG_AOT(HpvgSpectralNormIter);          /* (w, u, v) -> (sigma2[2] = sigma, 1/sigma; u'; v') (spectral_norm.py:146-151) */
HPVG_AOT(HpvgClipAdam);                  /* (param, grad, m, v, hyper[6] = lr, beta1, beta2, eps, clip, step) ->
                                            (param', m', v')                      (optimizers.py:41-43, nn.Adam)    */
HPVG_AOT(HpvgMSELoss);                   /* (a, b) -> loss[1]                      (nn.MSELoss)                      */
HPVG_AOT(HpvgKLLoss);                    /* (mu, logvar) -> loss[1]                (losses.py:5-7)                   */

#ifdef __cplusplus
}
#endif
#endif /* HPVG_H_ */
