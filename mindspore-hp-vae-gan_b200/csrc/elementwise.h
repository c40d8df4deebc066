// Internal host-side launchers of the bandwidth-bound kernels (see elementwise.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "det_reduce.cuh"
#include <cstdint>

namespace hpvg {

constexpr int ADAM_MAX_TENSORS = 64;
constexpr int ADAM_NORM_BLOCKS = 64;   // block partials per tensor of the gradient-norm pass (scratch: 64 x 64 floats)
struct AdamTable {
  float* p[ADAM_MAX_TENSORS];
  const float* g[ADAM_MAX_TENSORS];
  float* m[ADAM_MAX_TENSORS];
  float* v[ADAM_MAX_TENSORS];
  long long n[ADAM_MAX_TENSORS];
  float lr[ADAM_MAX_TENSORS];
};

cudaError_t ew_pack_cl(const float* x, int N, int C, long long sp, __nv_bfloat16* y, int c_pitch, int c_off,
                       int c_zero_to, cudaStream_t st);
cudaError_t ew_unpack_cl(const __nv_bfloat16* x, int N, int C, long long sp, int c_pitch, int c_off, float* y,
                         cudaStream_t st);
void ew_linear_taps_host(int n_in, int n_out, int align, int32_t* i0, int32_t* i1, float* l0, float* l1);
cudaError_t ew_linear_taps_dev(int n_in, int n_out, int align, int32_t* i0, int32_t* i1, float* l0, float* l1,
                               cudaStream_t st);
cudaError_t ew_resize3d_fwd(const float* x, long long NC, int Ti, int Hi, int Wi, float* y, int To, int Ho, int Wo,
                            int align, cudaStream_t st);
cudaError_t ew_resize3d_bwd(const float* gy, long long NC, int To, int Ho, int Wo, float* gx, int Ti, int Hi, int Wi,
                            int align, cudaStream_t st);
cudaError_t ew_upsample_noise_pack(const float* x, int N, int C, int Ti, int Hi, int Wi, int To, int Ho, int Wo,
                                   const float* noise, float amp, unsigned long long seed,
                                   unsigned long long sample_base, const unsigned long long* d_sample_offset,
                                   float* up, __nv_bfloat16* xin, int xin_f32, cudaStream_t st);
cudaError_t ew_frames_to_clip(const uint8_t* frames, int Hs, int Ws, int bgr, int start, int every, int T, int H, int W,
                              int hflip, float* clip, cudaStream_t st);
cudaError_t ew_randn(float* z, long long n, unsigned long long seed, unsigned long long offset,
                     const unsigned long long* d_offset, cudaStream_t st);
cudaError_t ew_counter_add(unsigned long long* c, unsigned long long inc, cudaStream_t st);
cudaError_t ew_bn_stats_cl(const __nv_bfloat16* y, long long voxels, double* sum, double* sumsq, DetScratch det,
                           cudaStream_t st);
cudaError_t ew_bn_finalize(const double* sum, const double* sumsq, long long count, const float* gamma,
                           const float* beta, float eps, float momentum, float* mm, float* mv, float* scale,
                           float* shift, float* mean, float* invstd, cudaStream_t st);
cudaError_t ew_bn_apply_cl(const __nv_bfloat16* y, long long voxels, const float* scale, const float* shift, int act,
                           __nv_bfloat16* x, cudaStream_t st);
constexpr int SN_MAX_LAYERS = 16;
struct SnTable {
  const float* w[SN_MAX_LAYERS];
  float* u[SN_MAX_LAYERS];
  float* v[SN_MAX_LAYERS];
  float* sigma[SN_MAX_LAYERS];      // [2]: sigma, 1/sigma
  const float* bias[SN_MAX_LAYERS]; // nullable
  float* aff[SN_MAX_LAYERS];        // nullable [2][64]: (1/sigma, bias) conv epilogue vectors
  float* u_copy[SN_MAX_LAYERS];     // nullable: snapshot of the updated u / v for this pass's backward
  float* v_copy[SN_MAX_LAYERS];
  int cout[SN_MAX_LAYERS];
  int k[SN_MAX_LAYERS];
};
cudaError_t ew_sn_power_iter(const float* w, int cout, int k, float* u, float* v, float* sigma, float* inv_sigma,
                             cudaStream_t st);
cudaError_t ew_sn_power_iter_multi(const SnTable& tab, int n_layers, cudaStream_t st);
constexpr int BN_MOVING_MAX_LAYERS = 64;
struct BnMovingTable {
  const float* saved[BN_MOVING_MAX_LAYERS];
  float* mm[BN_MOVING_MAX_LAYERS];
  float* mv[BN_MOVING_MAX_LAYERS];
  const float* center[BN_MOVING_MAX_LAYERS];   // nullable entries: offset of the centred frame `saved` lives in
};
struct BnCenterTable {     // per layer: inputs gamma, beta, moving mean / variance, conv bias; outputs center[64], aff[2][64]
  const float* gamma[BN_MOVING_MAX_LAYERS];
  const float* beta[BN_MOVING_MAX_LAYERS];
  const float* mm[BN_MOVING_MAX_LAYERS];
  const float* mv[BN_MOVING_MAX_LAYERS];
  const float* bias[BN_MOVING_MAX_LAYERS];
  float* center[BN_MOVING_MAX_LAYERS];
  float* aff[BN_MOVING_MAX_LAYERS];
};
cudaError_t ew_bn_center_multi(const BnCenterTable& tab, int n_layers, float eps, cudaStream_t st);
cudaError_t ew_bn_moving_update_multi(const BnMovingTable& tab, int n_layers, float eps, float momentum,
                                      cudaStream_t st);
cudaError_t ew_bn_train_apply_cl(const __nv_bfloat16* y, long long voxels, const double* sums, const float* gamma,
                                 const float* beta, float eps, float momentum, float* mm, float* mv, float* saved,
                                 int act, __nv_bfloat16* x, const float* center,
                                 cudaStream_t st);
cudaError_t ew_bn_fold_eval(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                            const float* bias, int C, float* scale, float* shift, cudaStream_t st);
cudaError_t ew_affine_from_bias(const float* bias, const float* inv_sigma, int C, float* scale, float* shift,
                                cudaStream_t st);
cudaError_t ew_reduce(int op, const float* a, const float* b, long long n, float* out, DetScratch det, cudaStream_t st);
cudaError_t ew_reparam_bwd(const float* gz, const float* eps, const float* lv, long long n, float* gmu, float* glv,
                           cudaStream_t st);
cudaError_t ew_reparam(const float* mu, const float* lv, const float* eps, long long n, float* z, cudaStream_t st);
cudaError_t ew_adam_clip(const AdamTable& tab, int n_tensors, float* norms_scratch, float beta1, float beta2,
                         float eps, float bias_corr, float clip, const unsigned long long* d_step, cudaStream_t st,
                         const float* d_hyper = nullptr /* device (lr, beta1, beta2, eps, clip, step) overriding all */);

cudaError_t ew_lrelu_bwd_cl(const __nv_bfloat16* ga, const __nv_bfloat16* a, long long elems, __nv_bfloat16* gz,
                            cudaStream_t st);
cudaError_t ew_bn_bwd_cl(const __nv_bfloat16* ga, const __nv_bfloat16* y, long long voxels, const float* saved, int act,
                         double* sums, DetScratch det, __nv_bfloat16* gy, float* dgamma, float* dbeta, int accumulate,
                         cudaStream_t st);
cudaError_t ew_colsum_cl(const __nv_bfloat16* g, long long voxels, double* scratch, DetScratch det, float* out,
                         int accumulate, cudaStream_t st);
cudaError_t ew_diff_scale(const float* a, const float* b, long long n, float coef, int accumulate, float* g,
                          cudaStream_t st);
cudaError_t ew_tanh_bwd(const float* g, const float* out, long long n, float* gpre, cudaStream_t st);
cudaError_t ew_axpby(float a, const float* x, float b, float* y, long long n, cudaStream_t st);
cudaError_t ew_fill(float* y, float v, long long n, cudaStream_t st);
cudaError_t ew_gather_strided(const float* src, long long n, long long stride, long long offset, float* dst,
                              cudaStream_t st);
cudaError_t ew_channel_sum_ncdhw(const float* g, int N, int C, long long sp, int accumulate, double* scratch,
                                 float* out, cudaStream_t st);
cudaError_t ew_kl_grad(const float* mu, const float* lv, long long n, float coef, float* gmu, float* glv,
                       cudaStream_t st);
cudaError_t ew_sn_grad(const float* G, const float* w, const float* u, const float* v, const float* sigma, int cout,
                       int k, int accumulate, float* scratch, float* gw, cudaStream_t st);
cudaError_t ew_lerp(const float* a, const float* b, float alpha, long long n, float* out, cudaStream_t st);
cudaError_t ew_gp_grad(const float* g, int N, int C, long long sp, float lambda, float* Gout, float* gp,
                       DetScratch det, cudaStream_t st);


// ---- fp32 channels-last twins (tf32 precision mode; elementwise_f32.cu)
cudaError_t ew_pack_cl_f32(const float* x, int N, int C, long long sp, float* y, int c_pitch, int c_off, int c_zero_to,
                           cudaStream_t st);
cudaError_t ew_unpack_cl_f32(const float* x, int N, int C, long long sp, int c_pitch, int c_off, float* y,
                             cudaStream_t st);
cudaError_t ew_bn_stats_cl_f32(const float* y, long long voxels, double* sum, double* sumsq, DetScratch det,
                               cudaStream_t st);
cudaError_t ew_bn_apply_cl_f32(const float* y, long long voxels, const float* scale, const float* shift, int act,
                               float* x, cudaStream_t st);
cudaError_t ew_bn_train_apply_cl_f32(const float* y, long long voxels, const double* sums, const float* gamma,
                                     const float* beta, float eps, float momentum, float* mm, float* mv, float* saved,
                                     int act, float* x, const float* center, cudaStream_t st);
cudaError_t ew_lrelu_bwd_cl_f32(const float* ga, const float* a, long long elems, float* gz, cudaStream_t st);
cudaError_t ew_bn_bwd_cl_f32(const float* ga, const float* y, long long voxels, const float* saved, int act,
                             double* sums, DetScratch det, float* gy, float* dgamma, float* dbeta, int accumulate,
                             cudaStream_t st);
cudaError_t ew_colsum_cl_f32(const float* g, long long voxels, double* scratch, DetScratch det, float* out,
                             int accumulate, cudaStream_t st);

// strided window of a bf16 channels-last tensor (+ optional ReLU): out = act(in[.., h0 + ho*sh, w0 + wo*sw, :])
cudaError_t ew_reflect_pad_cl(const void* in, int N, int T, int H, int W, int voxel_bytes, int pt, int ph, void* out,
                              cudaStream_t st);
cudaError_t ew_slice_act_cl(const __nv_bfloat16* in, int NT, int Hi, int Wi, int C, int Ho, int Wo, int h0, int w0, int sh,
                            int sw, int relu, __nv_bfloat16* out, cudaStream_t st);

// Box-Muller in place over host-drawn uniforms (see elementwise.cu)
cudaError_t ew_box_muller_inplace(float* z, long long n, cudaStream_t st);

}  // namespace hpvg
