// extern "C" surface of libhpvg.so (declared in include/hpvg.h): a minimal CUDA runtime shim for the ctypes route
// plus one entry point per operator of the hot path, and the MindSpore ops.Custom(func_type="aot") wrappers.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/hpvg.h"
#include "conv3d_umma.h"
#include "elementwise.h"
#include "launch.cuh"

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
}  // namespace
namespace hpvg {
int g_pdl_mode = -1;
}
namespace {
int g_sm_count = 0;

// Scratch memory is PER STREAM: two host threads (or two branches of a captured graph) driving different streams never
// share a reduction buffer, so every entry point is safe to call concurrently as long as each caller uses its own
// stream.  A stream's context is created by hpvg_stream_create / hpvg_stream_attach (or lazily on first
// use, which allocates and therefore must not happen while that stream is being captured into a graph).
struct StreamCtx {
  float* adam_norms = nullptr;   // [ADAM_MAX_TENSORS] per-tensor gradient norms
  float* wgrad_ws = nullptr;     // [sm/3][27][64][64] partial weight gradients
  double* sums = nullptr;        // [128] reduction scratch (BatchNorm backward, column sums)
  float* fscratch = nullptr;     // [256] fp32 scratch (spectral-norm gradient partial dots)
  void* arena = nullptr;         // grow-only workspace of the ops.Custom(aot) entry points
  size_t arena_cap = 0;
  hpvg::DetScratch det{nullptr, nullptr};   // block partials + ticket counter of the deterministic reductions
};
std::mutex g_ctx_mu;
std::unordered_map<void*, StreamCtx*> g_ctx;

void ctx_free(StreamCtx* c) {
  if (!c) return;
  cudaFree(c->adam_norms); cudaFree(c->wgrad_ws); cudaFree(c->sums); cudaFree(c->fscratch); cudaFree(c->arena);
  cudaFree(c->det.partials); cudaFree(c->det.counter);
  delete c;
}

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return HPVG_E_CUDA;
}
#define CU(x)                                          \
  do {                                                 \
    cudaError_t e_ = (x);                              \
    if (e_ != cudaSuccess) return cuda_fail(e_, #x);   \
  } while (0)
#define KL(x, n)                                       \
  do {                                                 \
    cudaError_t e_ = (x);                              \
    if (e_ != cudaSuccess) return cuda_fail(e_, #x);   \
    g_launches += (n);                                 \
  } while (0)

inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

// The scratch context of `stream` (nullptr + g_err on failure).
StreamCtx* ctx_for(void* stream) {
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  auto it = g_ctx.find(stream);
  if (it != g_ctx.end()) return it->second;
  if (g_sm_count == 0) {
    g_err = "hpvg_init was not called";
    return nullptr;
  }
  StreamCtx* c = new StreamCtx();
  size_t wg = hpvg::conv3d_wgrad_workspace_bytes(g_sm_count);
  const size_t wg32 = hpvg::conv3d_wgrad_tf32_workspace_bytes(g_sm_count);
  if (wg32 > wg) wg = wg32;
  cudaError_t e = cudaMalloc(&c->adam_norms, hpvg::ADAM_MAX_TENSORS * hpvg::ADAM_NORM_BLOCKS * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&c->det.partials, static_cast<size_t>(hpvg::DET_MAX_BLOCKS) * 128 * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&c->det.counter, 256);
  if (e == cudaSuccess) e = cudaMemset(c->det.counter, 0, 256);
  if (e == cudaSuccess) e = cudaMalloc(&c->wgrad_ws, wg);
  if (e == cudaSuccess) e = cudaMalloc(&c->sums, 128 * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&c->fscratch, 256 * sizeof(float));
  if (e != cudaSuccess) {
    cudaGetLastError();
    g_err = std::string("scratch allocation for a new stream failed (") + cudaGetErrorString(e) +
            "); a stream first used inside a graph capture must be registered with hpvg_stream_attach before";
    ctx_free(c);
    return nullptr;
  }
  g_ctx[stream] = c;
  return c;
}

// `bytes` of 256-byte aligned workspace on `stream` (grow-only; growing synchronises that stream once).
void* arena_for(void* stream, size_t bytes) {
  StreamCtx* c = ctx_for(stream);
  if (!c) return nullptr;
  if (bytes <= c->arena_cap) return c->arena;
  if (c->arena) {
    if (cudaStreamSynchronize(S(stream)) != cudaSuccess) return nullptr;
    cudaFree(c->arena);
    c->arena = nullptr;
    c->arena_cap = 0;
  }
  const size_t cap = (bytes + (1u << 20) - 1) & ~static_cast<size_t>((1u << 20) - 1);
  if (cudaMalloc(&c->arena, cap) != cudaSuccess) {
    cudaGetLastError();
    g_err = "workspace allocation failed";
    return nullptr;
  }
  c->arena_cap = cap;
  return c->arena;
}
#define CTX(c, st)                        \
  StreamCtx* c = ctx_for(st);             \
  if (!c) return HPVG_E_ARG

}  // namespace

extern "C" {

int hpvg_version(void) { return 100; }
const char* hpvg_last_error(void) { return g_err.c_str(); }
long long hpvg_launch_count(void) { return g_launches.load(); }
int hpvg_set_pdl(int on) {
  const int before = hpvg::pdl_enabled() ? 1 : 0;
  hpvg::g_pdl_mode = on ? 1 : 0;
  return before;
}

int hpvg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
int hpvg_init(int device) {
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(HPVG_E_UNSUPPORTED, "libhpvg needs an sm_100a (B200) device");
  g_sm_count = prop.multiProcessorCount;
  if (!ctx_for(nullptr)) return HPVG_E_CUDA;      // the legacy default stream's scratch
  return HPVG_OK;
}
int hpvg_sm_count(void) { return g_sm_count; }
int hpvg_malloc(void** p, size_t bytes) {
  CU(cudaMalloc(p, bytes ? bytes : 16));
  return HPVG_OK;
}
int hpvg_free(void* p) {
  CU(cudaFree(p));
  return HPVG_OK;
}
int hpvg_host_alloc(void** p, size_t bytes) {
  CU(cudaMallocHost(p, bytes ? bytes : 16));
  return HPVG_OK;
}
int hpvg_host_free(void* p) {
  CU(cudaFreeHost(p));
  return HPVG_OK;
}
int hpvg_memset(void* p, int v, size_t bytes, void* st) {
  CU(cudaMemsetAsync(p, v, bytes, S(st)));
  return HPVG_OK;
}
int hpvg_h2d(void* d, const void* h, size_t bytes, void* st) {
  CU(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, S(st)));
  return HPVG_OK;
}
int hpvg_d2h(void* h, const void* d, size_t bytes, void* st) {
  CU(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, S(st)));
  return HPVG_OK;
}
int hpvg_d2d(void* d, const void* s, size_t bytes, void* st) {
  CU(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, S(st)));
  return HPVG_OK;
}
int hpvg_stream_create(void** st) {
  cudaStream_t s;
  CU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *st = s;
  if (!ctx_for(s)) return HPVG_E_CUDA;
  return HPVG_OK;
}
int hpvg_stream_attach(void* st) { return ctx_for(st) ? HPVG_OK : HPVG_E_CUDA; }
int hpvg_stream_detach(void* st) {
  StreamCtx* c = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    auto it = g_ctx.find(st);
    if (it == g_ctx.end()) return HPVG_OK;
    c = it->second;
    g_ctx.erase(it);
  }
  CU(cudaStreamSynchronize(S(st)));
  ctx_free(c);
  return HPVG_OK;
}
int hpvg_stream_destroy(void* st) {
  hpvg_stream_detach(st);
  CU(cudaStreamDestroy(S(st)));
  return HPVG_OK;
}
int hpvg_stream_sync(void* st) {
  CU(cudaStreamSynchronize(S(st)));
  return HPVG_OK;
}
int hpvg_device_sync(void) {
  CU(cudaDeviceSynchronize());
  return HPVG_OK;
}
int hpvg_event_create(void** ev) {
  cudaEvent_t e;
  CU(cudaEventCreate(&e));
  *ev = e;
  return HPVG_OK;
}
int hpvg_event_destroy(void* ev) {
  CU(cudaEventDestroy(static_cast<cudaEvent_t>(ev)));
  return HPVG_OK;
}
int hpvg_event_record(void* ev, void* st) {
  CU(cudaEventRecord(static_cast<cudaEvent_t>(ev), S(st)));
  return HPVG_OK;
}
int hpvg_stream_wait_event(void* st, void* ev) {
  CU(cudaStreamWaitEvent(S(st), static_cast<cudaEvent_t>(ev), 0));
  return HPVG_OK;
}
int hpvg_event_sync(void* ev) {
  CU(cudaEventSynchronize(static_cast<cudaEvent_t>(ev)));
  return HPVG_OK;
}
int hpvg_event_elapsed_ms(void* a, void* b, float* ms) {
  CU(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(a), static_cast<cudaEvent_t>(b)));
  return HPVG_OK;
}
int hpvg_graph_begin(void* st) {
  CU(cudaStreamBeginCapture(S(st), cudaStreamCaptureModeRelaxed));
  return HPVG_OK;
}
int hpvg_graph_end(void* st, void** exec) {
  cudaGraph_t g;
  CU(cudaStreamEndCapture(S(st), &g));
  cudaGraphExec_t e;
  cudaError_t r = cudaGraphInstantiate(&e, g, 0);
  cudaGraphDestroy(g);
  if (r != cudaSuccess) return cuda_fail(r, "cudaGraphInstantiate");
  *exec = e;
  return HPVG_OK;
}
int hpvg_graph_launch(void* exec, void* st) {
  CU(cudaGraphLaunch(static_cast<cudaGraphExec_t>(exec), S(st)));
  return HPVG_OK;
}
int hpvg_graph_destroy(void* exec) {
  CU(cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(exec)));
  return HPVG_OK;
}

// ------------------------------------------------------------------------------------------------ layout
int hpvg_pack_cl(const float* x, int N, int C, int T, int H, int W, void* y, int c_pitch, int c_off, int c_zero_to,
                 void* st) {
  if ((c_off & 7) || (c_pitch & 7)) return fail(HPVG_E_ARG, "pack_cl: c_off and c_pitch must be multiples of 8");
  if (c_zero_to > c_pitch || c_off + C > c_pitch) return fail(HPVG_E_ARG, "pack_cl: channels exceed pitch");
  if (N <= 0 || C <= 0 || T <= 0 || H <= 0 || W <= 0) return HPVG_OK;
  KL(hpvg::ew_pack_cl(x, N, C, static_cast<long long>(T) * H * W, static_cast<__nv_bfloat16*>(y), c_pitch, c_off,
                      c_zero_to, S(st)), 1);
  return HPVG_OK;
}
int hpvg_unpack_cl(const void* x, int N, int C, int T, int H, int W, int c_pitch, int c_off, float* y, void* st) {
  if ((c_off & 7) || (c_pitch & 7)) return fail(HPVG_E_ARG, "unpack_cl: c_off and c_pitch must be multiples of 8");
  if (N <= 0 || C <= 0 || T <= 0 || H <= 0 || W <= 0) return HPVG_OK;
  KL(hpvg::ew_unpack_cl(static_cast<const __nv_bfloat16*>(x), N, C, static_cast<long long>(T) * H * W, c_pitch,
                        c_off, y, S(st)), 1);
  return HPVG_OK;
}

int hpvg_pack_cl_f32(const float* x, int N, int C, int T, int H, int W, float* y, int c_pitch, int c_off,
                     int c_zero_to, void* st) {
  if ((c_off & 3) || (c_pitch & 3)) return fail(HPVG_E_ARG, "pack_cl_f32: c_off and c_pitch must be multiples of 4");
  if (c_zero_to > c_pitch || c_off + C > c_pitch) return fail(HPVG_E_ARG, "pack_cl_f32: channels exceed pitch");
  if (N <= 0 || C <= 0 || T <= 0 || H <= 0 || W <= 0) return HPVG_OK;
  KL(hpvg::ew_pack_cl_f32(x, N, C, static_cast<long long>(T) * H * W, y, c_pitch, c_off, c_zero_to, S(st)), 1);
  return HPVG_OK;
}
int hpvg_unpack_cl_f32(const float* x, int N, int C, int T, int H, int W, int c_pitch, int c_off, float* y, void* st) {
  if (c_off + C > c_pitch) return fail(HPVG_E_ARG, "unpack_cl_f32: channels exceed pitch");
  if (N <= 0 || C <= 0 || T <= 0 || H <= 0 || W <= 0) return HPVG_OK;
  KL(hpvg::ew_unpack_cl_f32(x, N, C, static_cast<long long>(T) * H * W, c_pitch, c_off, y, S(st)), 1);
  return HPVG_OK;
}

// ------------------------------------------------------------------------------------------------ conv
int hpvg_conv_wimg_bytes(int mode) { return hpvg::conv3d_umma_wimg_bytes(mode); }

int hpvg_conv_pack_weights(const float* w, int w_cout, int w_cin, int kt, int mode, int transpose_flip, int cout_off,
                           int cout, int cin_off, int cin, void* wimg, void* st) {
  const int max_co = (mode == HPVG_CONV_64_16) ? 16 : (mode == HPVG_CONV_64_3 || mode == HPVG_CONV_T32_3) ? 3 : 64;
  const int max_ci = (mode == HPVG_CONV_8_64) ? 8 : (mode == HPVG_CONV_T4_64) ? 4
                     : (mode == HPVG_CONV_T32_64 || mode == HPVG_CONV_T32_3) ? 32 : 64;
  if (cout > max_co || cin > max_ci || cout <= 0 || cin <= 0)
    return fail(HPVG_E_ARG, "conv_pack_weights: channel counts exceed the kernel variant");
  const char* e = hpvg::conv3d_pack_weights(w, w_cout, w_cin, kt, mode, transpose_flip, cout_off, cout, cin_off, cin,
                                            wimg, S(st));
  if (e) return fail(HPVG_E_CUDA, std::string("conv_pack_weights: ") + e);
  g_launches += 1;
  return HPVG_OK;
}

int hpvg_conv_pack_weights_multi(int n, const float* const* w, const int* w_cout, const int* w_cin, const int* kt,
                                 const int* mode, const int* transpose_flip, const int* cout_off, const int* cout,
                                 const int* cin_off, const int* cin, void* const* wimg, void* st) {
  if (n <= 0) return HPVG_OK;
  std::vector<hpvg::PackEntry> es(n);
  for (int i = 0; i < n; ++i) {
    hpvg::PackEntry& e = es[i];
    e.w = w[i]; e.img = wimg[i]; e.cout = cout[i]; e.cin = cin[i]; e.kt = kt[i]; e.mode = mode[i];
    e.flip = transpose_flip[i]; e.cin_off = cin_off[i]; e.cout_off = cout_off[i]; e.w_cout = w_cout[i];
    e.w_cin = w_cin[i]; e.total = 0;
    if (!e.img || (e.mode >= 0 && !e.w)) return fail(HPVG_E_ARG, "conv_pack_weights_multi: null pointer");
  }
  const char* err = hpvg::conv3d_pack_weights_multi(es.data(), n, S(st));
  if (err) return fail(HPVG_E_CUDA, std::string("conv_pack_weights_multi: ") + err);
  g_launches += (n + hpvg::PACK_MAX_ENTRIES - 1) / hpvg::PACK_MAX_ENTRIES;
  return HPVG_OK;
}
int hpvg_conv_cl(int mode, int N, int T, int H, int W, const void* in, int in_pitch, const void* wimg,
                 const float* scale, const float* shift, int act, int out_mode, void* out, int out_pitch,
                 int out_coff, int cout_real, const float* addend, double* stats, const void* mask, int mask_pitch,
                 void* st) {
  if (N <= 0 || T <= 0 || H <= 0 || W <= 0) return HPVG_OK;  // empty input: nothing to do
  if (!in || !wimg || !scale || !shift || !out) return fail(HPVG_E_ARG, "conv_cl: null pointer");
  if (out_mode == HPVG_OUT_BF16_CL && ((out_pitch & 7) || (out_coff & 7)))
    return fail(HPVG_E_ARG, "conv_cl: out_pitch/out_coff must be multiples of 8");
  if (g_sm_count == 0) return fail(HPVG_E_ARG, "hpvg_init was not called");
  const bool tf32 = hpvg::conv_mode_is_tf32(mode);
  if (tf32 != (out_mode == HPVG_OUT_F32_CL) && out_mode != HPVG_OUT_F32_RAW && out_mode != HPVG_OUT_F32_NCDHW)
    return fail(HPVG_E_ARG, "conv_cl: bf16 kernels write HPVG_OUT_BF16_CL, tf32 kernels HPVG_OUT_F32_CL");
  if (stats && (mode == HPVG_CONV_64_16 || mode == HPVG_CONV_64_3 || mode == HPVG_CONV_T32_3 ||
                (out_mode != HPVG_OUT_BF16_CL && out_mode != HPVG_OUT_F32_CL)))
    return fail(HPVG_E_ARG, "conv_cl: fused BatchNorm statistics need a 64-channel channels-last output");
  hpvg::ConvLaunch L;
  L.mode = mode;
  L.N = N; L.T = T; L.H = H; L.W = W;
  L.in = in; L.in_pitch = in_pitch;
  L.wimg = wimg; L.scale = scale; L.shift = shift;
  L.act = act; L.out_mode = out_mode;
  L.out = out; L.out_pitch = out_pitch; L.out_coff = out_coff; L.cout_real = cout_real;
  L.addend = addend;
  L.stats = stats;
  if (stats) {
    CTX(cx, st);
    L.det = cx->det;
  }
  L.mask = mask;
  L.mask_pitch = mask_pitch;
  L.max_pairs = g_sm_count / 2;
  const char* e = hpvg::conv3d_umma_launch(L, S(st));
  if (e) return fail(HPVG_E_CUDA, std::string("conv_cl: ") + e);
  g_launches += 1;
  return HPVG_OK;
}

// ------------------------------------------------------------------------------------------------ resize
int hpvg_linear_taps(int n_in, int n_out, int align, int32_t* i0, int32_t* i1, float* l0, float* l1) {
  if (n_in <= 0 || n_out <= 0) return fail(HPVG_E_ARG, "linear_taps: sizes must be positive");
  hpvg::ew_linear_taps_host(n_in, n_out, align, i0, i1, l0, l1);
  return HPVG_OK;
}
int hpvg_linear_taps_dev(int n_in, int n_out, int align, int32_t* i0, int32_t* i1, float* l0, float* l1, void* st) {
  if (n_in <= 0 || n_out <= 0) return fail(HPVG_E_ARG, "linear_taps: sizes must be positive");
  KL(hpvg::ew_linear_taps_dev(n_in, n_out, align, i0, i1, l0, l1, S(st)), 1);
  return HPVG_OK;
}
int hpvg_resize3d_fwd(const float* x, int N, int C, int Ti, int Hi, int Wi, float* y, int To, int Ho, int Wo,
                      int align, void* st) {
  if (Ti <= 0 || Hi <= 0 || Wi <= 0 || To <= 0 || Ho <= 0 || Wo <= 0)
    return fail(HPVG_E_ARG, "resize3d: sizes must be positive");   // trilinear.py:246-249 validator
  if (N * C == 0) return HPVG_OK;
  KL(hpvg::ew_resize3d_fwd(x, static_cast<long long>(N) * C, Ti, Hi, Wi, y, To, Ho, Wo, align, S(st)), 1);
  return HPVG_OK;
}
int hpvg_resize3d_bwd(const float* gy, int N, int C, int To, int Ho, int Wo, float* gx, int Ti, int Hi, int Wi,
                      int align, void* st) {
  if (Ti <= 0 || Hi <= 0 || Wi <= 0 || To <= 0 || Ho <= 0 || Wo <= 0)
    return fail(HPVG_E_ARG, "resize3d: sizes must be positive");
  if (N * C == 0) return HPVG_OK;
  KL(hpvg::ew_resize3d_bwd(gy, static_cast<long long>(N) * C, To, Ho, Wo, gx, Ti, Hi, Wi, align, S(st)), 1);
  return HPVG_OK;
}
int hpvg_upsample_noise_pack(const float* x, int N, int C, int Ti, int Hi, int Wi, int To, int Ho, int Wo,
                             const float* noise, float amp, uint64_t seed, uint64_t sample_base,
                             const uint64_t* d_sample_offset, float* up, void* xin, void* st) {
  if (C < 1 || C > 4) return fail(HPVG_E_ARG, "upsample_noise_pack: 1 <= C <= 4");
  if (N <= 0) return HPVG_OK;
  KL(hpvg::ew_upsample_noise_pack(x, N, C, Ti, Hi, Wi, To, Ho, Wo, noise, amp, seed, sample_base,
                                  reinterpret_cast<const unsigned long long*>(d_sample_offset), up,
                                  static_cast<__nv_bfloat16*>(xin), 0, S(st)), 1);
  return HPVG_OK;
}
int hpvg_upsample_noise_pack_f32(const float* x, int N, int C, int Ti, int Hi, int Wi, int To, int Ho, int Wo,
                                 const float* noise, float amp, uint64_t seed, uint64_t sample_base,
                                 const uint64_t* d_sample_offset, float* up, float* xin, void* st) {
  if (C < 1 || C > 4) return fail(HPVG_E_ARG, "upsample_noise_pack: 1 <= C <= 4");
  if (N <= 0) return HPVG_OK;
  KL(hpvg::ew_upsample_noise_pack(x, N, C, Ti, Hi, Wi, To, Ho, Wo, noise, amp, seed, sample_base,
                                  reinterpret_cast<const unsigned long long*>(d_sample_offset), up,
                                  reinterpret_cast<__nv_bfloat16*>(xin), 1, S(st)), 1);
  return HPVG_OK;
}

// ------------------------------------------------------------------------------------------------ batch norm
int hpvg_bn_stats_cl(const void* y, long long voxels, double* sum, double* sumsq, void* st) {
  if (voxels <= 0) return fail(HPVG_E_ARG, "bn_stats: empty batch");
  CTX(cx, st);
  KL(hpvg::ew_bn_stats_cl(static_cast<const __nv_bfloat16*>(y), voxels, sum, sumsq, cx->det, S(st)), 1);
  return HPVG_OK;
}
int hpvg_bn_finalize(const double* sum, const double* sumsq, long long count, const float* gamma, const float* beta,
                     float eps, float momentum, float* mm, float* mv, float* scale, float* shift, float* mean,
                     float* invstd, void* st) {
  KL(hpvg::ew_bn_finalize(sum, sumsq, count, gamma, beta, eps, momentum, mm, mv, scale, shift, mean, invstd, S(st)), 1);
  return HPVG_OK;
}
int hpvg_bn_apply_lrelu_cl(const void* y, long long voxels, const float* scale, const float* shift, int act, void* x,
                           void* st) {
  if (voxels <= 0) return HPVG_OK;
  KL(hpvg::ew_bn_apply_cl(static_cast<const __nv_bfloat16*>(y), voxels, scale, shift, act,
                          static_cast<__nv_bfloat16*>(x), S(st)), 1);
  return HPVG_OK;
}

int hpvg_bn_train_apply_cl(const void* y, long long voxels, const double* sums, const float* gamma, const float* beta,
                           float eps, float momentum, float* mm, float* mv, float* saved, int act, void* x,
                           const float* center, void* st) {
  if (voxels <= 0) return fail(HPVG_E_ARG, "bn_train_apply: empty batch");
  if (!y || !sums || !gamma || !beta || !x) return fail(HPVG_E_ARG, "bn_train_apply: null pointer");
  KL(hpvg::ew_bn_train_apply_cl(static_cast<const __nv_bfloat16*>(y), voxels, sums, gamma, beta, eps, momentum, mm, mv,
                                saved, act, static_cast<__nv_bfloat16*>(x), center, S(st)), 1);
  return HPVG_OK;
}
int hpvg_bn_center_multi(int n_layers, const float* const* gamma, const float* const* beta, const float* const* mm,
                         const float* const* mv, const float* const* bias, float* const* center, float* const* aff,
                         float eps, void* st) {
  if (n_layers < 0 || !gamma || !beta || !mm || !mv || !bias || !center || !aff)
    return fail(HPVG_E_ARG, "bn_center_multi: bad arguments");
  for (int base = 0; base < n_layers; base += hpvg::BN_MOVING_MAX_LAYERS) {
    hpvg::BnCenterTable tab;
    std::memset(&tab, 0, sizeof(tab));
    int cnt = 0;
    for (int i = base; i < n_layers && cnt < hpvg::BN_MOVING_MAX_LAYERS; ++i, ++cnt) {
      if (!gamma[i] || !beta[i] || !mm[i] || !mv[i] || !bias[i] || !center[i] || !aff[i])
        return fail(HPVG_E_ARG, "bn_center_multi: null pointer");
      tab.gamma[cnt] = gamma[i]; tab.beta[cnt] = beta[i]; tab.mm[cnt] = mm[i]; tab.mv[cnt] = mv[i];
      tab.bias[cnt] = bias[i]; tab.center[cnt] = center[i]; tab.aff[cnt] = aff[i];
    }
    KL(hpvg::ew_bn_center_multi(tab, cnt, eps, S(st)), 1);
  }
  return HPVG_OK;
}

int hpvg_bn_moving_update_multi(int n_layers, const float* const* saved, float* const* mm, float* const* mv,
                                const float* const* center, float eps, float momentum, void* st) {
  if (n_layers < 0 || !saved || !mm || !mv) return fail(HPVG_E_ARG, "bn_moving_update_multi: bad arguments");
  for (int base = 0; base < n_layers; base += hpvg::BN_MOVING_MAX_LAYERS) {
    hpvg::BnMovingTable tab;
    std::memset(&tab, 0, sizeof(tab));
    int cnt = 0;
    for (int i = base; i < n_layers && cnt < hpvg::BN_MOVING_MAX_LAYERS; ++i, ++cnt) {
      if (!saved[i] || !mm[i] || !mv[i]) return fail(HPVG_E_ARG, "bn_moving_update_multi: null pointer");
      tab.saved[cnt] = saved[i]; tab.mm[cnt] = mm[i]; tab.mv[cnt] = mv[i];
      tab.center[cnt] = center ? center[i] : nullptr;
    }
    KL(hpvg::ew_bn_moving_update_multi(tab, cnt, eps, momentum, S(st)), 1);
  }
  return HPVG_OK;
}

// ------------------------------------------------------------------------------------------------ spectral norm
int hpvg_sn_power_iter_multi(int n_layers, const float* const* w, const int* cout, const int* k, float* const* u,
                             float* const* v, float* const* sigma2, const float* const* bias, float* const* aff,
                             float* const* u_copy, float* const* v_copy, void* st) {
  if (n_layers < 1 || n_layers > hpvg::SN_MAX_LAYERS)
    return fail(HPVG_E_ARG, "sn_power_iter_multi: 1..16 layers per call");
  hpvg::SnTable tab;
  std::memset(&tab, 0, sizeof(tab));
  for (int i = 0; i < n_layers; ++i) {
    if (cout[i] <= 0 || k[i] <= 0 || (2 * cout[i] + k[i]) * 4 > 48 * 1024)
      return fail(HPVG_E_ARG, "sn_power_iter_multi: matrix too large for the single-CTA kernel");
    tab.w[i] = w[i]; tab.u[i] = u[i]; tab.v[i] = v[i]; tab.sigma[i] = sigma2[i];
    tab.bias[i] = bias ? bias[i] : nullptr;
    tab.aff[i] = aff ? aff[i] : nullptr;
    tab.u_copy[i] = u_copy ? u_copy[i] : nullptr;
    tab.v_copy[i] = v_copy ? v_copy[i] : nullptr;
    tab.cout[i] = cout[i]; tab.k[i] = k[i];
  }
  KL(hpvg::ew_sn_power_iter_multi(tab, n_layers, S(st)), 1);
  return HPVG_OK;
}
int hpvg_sn_power_iter(const float* w, int cout, int k, float* u, float* v, float* sigma, float* inv_sigma,
                       void* st) {
  if (cout <= 0 || k <= 0 || (2 * cout + k) * 4 > 48 * 1024)
    return fail(HPVG_E_ARG, "sn_power_iter: matrix too large for the single-CTA kernel");
  KL(hpvg::ew_sn_power_iter(w, cout, k, u, v, sigma, inv_sigma, S(st)), 1);
  return HPVG_OK;
}

// ------------------------------------------------------------------------------------------------ small ops
int hpvg_bn_fold_eval(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                      const float* bias, int C, float* scale, float* shift, void* st) {
  KL(hpvg::ew_bn_fold_eval(gamma, beta, mean, var, eps, bias, C, scale, shift, S(st)), 1);
  return HPVG_OK;
}
int hpvg_affine_from_bias(const float* bias, const float* inv_sigma, int C, float* scale, float* shift, void* st) {
  KL(hpvg::ew_affine_from_bias(bias, inv_sigma, C, scale, shift, S(st)), 1);
  return HPVG_OK;
}
int hpvg_mse(const float* a, const float* b, long long n, float* out, void* st) {
  if (n <= 0) return fail(HPVG_E_ARG, "mse: empty input");
  CTX(cx, st);
  KL(hpvg::ew_reduce(0, a, b, n, out, cx->det, S(st)), 1);
  return HPVG_OK;
}
int hpvg_mean(const float* a, long long n, float* out, void* st) {
  if (n <= 0) return fail(HPVG_E_ARG, "mean: empty input");
  CTX(cx, st);
  KL(hpvg::ew_reduce(1, a, nullptr, n, out, cx->det, S(st)), 1);
  return HPVG_OK;
}
int hpvg_kl(const float* mu, const float* lv, long long n, float* out, void* st) {
  if (n <= 0) return fail(HPVG_E_ARG, "kl: empty input");
  CTX(cx, st);
  KL(hpvg::ew_reduce(2, mu, lv, n, out, cx->det, S(st)), 1);
  return HPVG_OK;
}
int hpvg_reparam(const float* mu, const float* lv, const float* eps, long long n, float* z, void* st) {
  if (n <= 0) return HPVG_OK;
  KL(hpvg::ew_reparam(mu, lv, eps, n, z, S(st)), 1);
  return HPVG_OK;
}

int hpvg_reparam_bwd(const float* gz, const float* eps, const float* lv, long long n, float* gmu, float* glv, void* st) {
  if (n <= 0) return HPVG_OK;
  if (!gz || !eps || !lv || !gmu || !glv) return fail(HPVG_E_ARG, "reparam_bwd: null pointer");
  KL(hpvg::ew_reparam_bwd(gz, eps, lv, n, gmu, glv, S(st)), 1);
  return HPVG_OK;
}

int hpvg_adam_clip_multi(int n_tensors, float* const* params, const float* const* grads, float* const* m,
                         float* const* v, const long long* sizes, const float* lrs, float beta1, float beta2,
                         float eps, int step, float clip, const uint64_t* d_step, void* st) {
  if (step < 1 && !d_step) return fail(HPVG_E_ARG, "adam: step is 1-based");
  if (step < 1) step = 1;
  CTX(cx, st);
  const float bc = static_cast<float>(std::sqrt(1.0 - std::pow(static_cast<double>(beta2), step)) /
                                      (1.0 - std::pow(static_cast<double>(beta1), step)));
  for (int base = 0; base < n_tensors; base += hpvg::ADAM_MAX_TENSORS) {
    hpvg::AdamTable tab;
    std::memset(&tab, 0, sizeof(tab));
    int cnt = 0;
    for (int i = base; i < n_tensors && cnt < hpvg::ADAM_MAX_TENSORS; ++i, ++cnt) {
      tab.p[cnt] = params[i];
      tab.g[cnt] = grads[i];
      tab.m[cnt] = m[i];
      tab.v[cnt] = v[i];
      tab.n[cnt] = sizes[i];
      tab.lr[cnt] = lrs[i];
    }
    KL(hpvg::ew_adam_clip(tab, cnt, cx->adam_norms, beta1, beta2, eps, bc, clip,
                          reinterpret_cast<const unsigned long long*>(d_step), S(st), nullptr), clip > 0.f ? 2 : 1);
  }
  return HPVG_OK;
}

// ------------------------------------------------------------------------------------------------ backward
int hpvg_conv_wgrad_cl(const void* x, int x_pitch, const void* gy, int gy_pitch, int N, int T, int H, int W,
                       float* dw, int w_cin, int kt, int co_off, int co_n, int ci_off, int ci_n, int accumulate,
                       float scale, void* st) {
  if (N <= 0 || T <= 0 || H <= 0 || W <= 0) return HPVG_OK;
  CTX(cx, st);
  if ((x_pitch & 7) || (gy_pitch & 7) || x_pitch < 8 || gy_pitch < 8)
    return fail(HPVG_E_ARG, "conv_wgrad_cl: operands must be bf16 channels-last tensors with a pitch of 8k channels");
  if ((x_pitch < 64 && ci_n > x_pitch) || (gy_pitch < 64 && co_n > gy_pitch))
    return fail(HPVG_E_ARG, "conv_wgrad_cl: block extent exceeds the channels of a narrow operand");
  if (co_n < 1 || co_n > 64 || ci_n < 1 || ci_n > 64 || (kt != 1 && kt != 3))
    return fail(HPVG_E_ARG, "conv_wgrad_cl: bad block extents");
  const char* e = hpvg::conv3d_wgrad_launch(x, x_pitch, gy, gy_pitch, N, T, H, W, dw, w_cin, kt, co_off, co_n, ci_off,
                                            ci_n, accumulate, scale, cx->wgrad_ws, g_sm_count, S(st));
  if (e) return fail(HPVG_E_CUDA, std::string("conv_wgrad_cl: ") + e);
  g_launches += 2;
  return HPVG_OK;
}
int hpvg_conv_wgrad_cl_tf32(const float* x, int x_pitch, const float* gy, int gy_pitch, int N, int T, int H, int W,
                            float* dw, int w_cin, int kt, int co_off, int co_n, int ci_off, int ci_n, int accumulate,
                            float scale, void* st) {
  if (N <= 0 || T <= 0 || H <= 0 || W <= 0) return HPVG_OK;
  CTX(cx, st);
  if ((x_pitch & 3) || (gy_pitch & 3) || x_pitch < 4 || gy_pitch < 4 || (x_pitch >= 32 && x_pitch < 64) ||
      (gy_pitch >= 32 && gy_pitch < 64))
    return fail(HPVG_E_ARG, "conv_wgrad_cl_tf32: operands are fp32 channels-last tensors of >= 64 channels, or narrow "
                            "tensors of 4..28 channels");
  if ((x_pitch < 64 && ci_n > x_pitch) || (gy_pitch < 64 && co_n > gy_pitch))
    return fail(HPVG_E_ARG, "conv_wgrad_cl_tf32: block extent exceeds the channels of a narrow operand");
  if (co_n < 1 || co_n > 64 || ci_n < 1 || ci_n > 64 || (kt != 1 && kt != 3))
    return fail(HPVG_E_ARG, "conv_wgrad_cl_tf32: bad block extents");
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(gy) & 15))
    return fail(HPVG_E_ARG, "conv_wgrad_cl_tf32: operands must be 16-byte aligned");
  const char* e = hpvg::conv3d_wgrad_tf32_launch(x, x_pitch, gy, gy_pitch, N, T, H, W, dw, w_cin, kt, co_off, co_n,
                                                 ci_off, ci_n, accumulate, scale, cx->wgrad_ws, g_sm_count, S(st));
  if (e) return fail(HPVG_E_CUDA, std::string("conv_wgrad_cl_tf32: ") + e);
  g_launches += 2;
  return HPVG_OK;
}
int hpvg_bn_stats_cl_f32(const float* y, long long voxels, double* sum, double* sumsq, void* st) {
  if (voxels <= 0) return fail(HPVG_E_ARG, "bn_stats: empty batch");
  CTX(cx, st);
  KL(hpvg::ew_bn_stats_cl_f32(y, voxels, sum, sumsq, cx->det, S(st)), 1);
  return HPVG_OK;
}
int hpvg_bn_apply_lrelu_cl_f32(const float* y, long long voxels, const float* scale, const float* shift, int act,
                               float* x, void* st) {
  if (voxels <= 0) return HPVG_OK;
  KL(hpvg::ew_bn_apply_cl_f32(y, voxels, scale, shift, act, x, S(st)), 1);
  return HPVG_OK;
}
int hpvg_bn_train_apply_cl_f32(const float* y, long long voxels, const double* sums, const float* gamma,
                               const float* beta, float eps, float momentum, float* mm, float* mv, float* saved,
                               int act, float* x, const float* center, void* st) {
  if (voxels <= 0) return fail(HPVG_E_ARG, "bn_train_apply: empty batch");
  if (!y || !sums || !gamma || !beta || !x) return fail(HPVG_E_ARG, "bn_train_apply: null pointer");
  KL(hpvg::ew_bn_train_apply_cl_f32(y, voxels, sums, gamma, beta, eps, momentum, mm, mv, saved, act, x, center, S(st)), 1);
  return HPVG_OK;
}
int hpvg_lrelu_bwd_cl_f32(const float* ga, const float* a, long long elems, float* gz, void* st) {
  if (elems <= 0) return HPVG_OK;
  if (elems & 3) return fail(HPVG_E_ARG, "lrelu_bwd_cl_f32: element count must be a multiple of 4");
  KL(hpvg::ew_lrelu_bwd_cl_f32(ga, a, elems, gz, S(st)), 1);
  return HPVG_OK;
}
int hpvg_bn_bwd_cl_f32(const float* ga, const float* y, long long voxels, const float* saved, int act, float* gy,
                       float* dgamma, float* dbeta, int accumulate, void* st) {
  if (voxels <= 0) return fail(HPVG_E_ARG, "bn_bwd: empty batch");
  CTX(cx, st);
  KL(hpvg::ew_bn_bwd_cl_f32(ga, y, voxels, saved, act, cx->sums, cx->det, gy, dgamma, dbeta, accumulate, S(st)), 2);
  return HPVG_OK;
}
int hpvg_colsum_cl_f32(const float* g, long long voxels, float* out, int accumulate, void* st) {
  if (voxels <= 0) return fail(HPVG_E_ARG, "colsum: empty input");
  CTX(cx, st);
  KL(hpvg::ew_colsum_cl_f32(g, voxels, cx->sums, cx->det, out, accumulate, S(st)), 2);
  return HPVG_OK;
}
int hpvg_slice_act_cl(const void* in, int NT, int Hi, int Wi, int C, int Ho, int Wo, int h0, int w0, int sh, int sw,
                      int relu, void* out, void* st) {
  if (NT <= 0 || Ho <= 0 || Wo <= 0) return HPVG_OK;
  if (!in || !out || (C & 7) || sh < 1 || sw < 1 || h0 < 0 || w0 < 0 || h0 + (Ho - 1) * sh >= Hi ||
      w0 + (Wo - 1) * sw >= Wi)
    return fail(HPVG_E_ARG, "slice_act_cl: window outside the input, or channels not a multiple of 8");
  KL(hpvg::ew_slice_act_cl(static_cast<const __nv_bfloat16*>(in), NT, Hi, Wi, C, Ho, Wo, h0, w0, sh, sw, relu,
                           static_cast<__nv_bfloat16*>(out), S(st)), 1);
  return HPVG_OK;
}
int hpvg_reflect_pad_cl(const void* in, int N, int T, int H, int W, int voxel_bytes, int pad_t, int pad_hw, void* out,
                        void* st) {
  if (N <= 0) return HPVG_OK;
  if (!in || !out || T < 1 || H < 1 || W < 1 || voxel_bytes < 16 || (voxel_bytes & 15) || pad_t < 0 || pad_hw < 0)
    return fail(HPVG_E_ARG, "reflect_pad_cl: bad geometry (voxels are multiples of 16 bytes)");
  if (pad_t >= T || pad_hw >= H || pad_hw >= W) return fail(HPVG_E_ARG, "reflect_pad_cl: padding must be smaller than the axis");
  KL(hpvg::ew_reflect_pad_cl(in, N, T, H, W, voxel_bytes, pad_t, pad_hw, out, S(st)), 1);
  return HPVG_OK;
}
int hpvg_lrelu_bwd_cl(const void* ga, const void* a, long long elems, void* gz, void* st) {
  if (elems <= 0) return HPVG_OK;
  if (elems & 7) return fail(HPVG_E_ARG, "lrelu_bwd_cl: element count must be a multiple of 8");
  KL(hpvg::ew_lrelu_bwd_cl(static_cast<const __nv_bfloat16*>(ga), static_cast<const __nv_bfloat16*>(a), elems,
                           static_cast<__nv_bfloat16*>(gz), S(st)), 1);
  return HPVG_OK;
}
int hpvg_bn_bwd_cl(const void* ga, const void* y, long long voxels, const float* saved, int act, void* gy,
                   float* dgamma, float* dbeta, int accumulate, void* st) {
  if (voxels <= 0) return fail(HPVG_E_ARG, "bn_bwd: empty batch");
  CTX(cx, st);
  KL(hpvg::ew_bn_bwd_cl(static_cast<const __nv_bfloat16*>(ga), static_cast<const __nv_bfloat16*>(y), voxels, saved,
                        act, cx->sums, cx->det, static_cast<__nv_bfloat16*>(gy), dgamma, dbeta, accumulate, S(st)), 2);
  return HPVG_OK;
}
int hpvg_colsum_cl(const void* g, long long voxels, float* out, int accumulate, void* st) {
  if (voxels <= 0) return fail(HPVG_E_ARG, "colsum: empty input");
  CTX(cx, st);
  KL(hpvg::ew_colsum_cl(static_cast<const __nv_bfloat16*>(g), voxels, cx->sums, cx->det, out, accumulate, S(st)), 2);
  return HPVG_OK;
}
int hpvg_mse_grad(const float* out, const float* target, long long n, float coef, int accumulate, float* g, void* st) {
  if (n <= 0) return HPVG_OK;
  KL(hpvg::ew_diff_scale(out, target, n, coef, accumulate, g, S(st)), 1);
  return HPVG_OK;
}
int hpvg_tanh_bwd(const float* g, const float* out, long long n, float* gpre, void* st) {
  if (n <= 0) return HPVG_OK;
  KL(hpvg::ew_tanh_bwd(g, out, n, gpre, S(st)), 1);
  return HPVG_OK;
}
int hpvg_axpby(float a, const float* x, float b, float* y, long long n, void* st) {
  if (n <= 0) return HPVG_OK;
  KL(hpvg::ew_axpby(a, x, b, y, n, S(st)), 1);
  return HPVG_OK;
}
int hpvg_fill(float* y, float v, long long n, void* st) {
  if (n <= 0) return HPVG_OK;
  KL(hpvg::ew_fill(y, v, n, S(st)), 1);
  return HPVG_OK;
}
int hpvg_frames_to_clip(const uint8_t* frames, int F, int Hs, int Ws, int bgr, int start, int every, int T, int H, int W,
                        int hflip, float* clip, void* st) {
  if (F <= 0 || Hs <= 0 || Ws <= 0 || T <= 0 || H <= 0 || W <= 0 || every <= 0 || start < 0)
    return fail(HPVG_E_ARG, "frames_to_clip: sizes must be positive");
  if (static_cast<long long>(start) + static_cast<long long>(T - 1) * every >= F)
    return fail(HPVG_E_ARG, "frames_to_clip: frame window runs past the decoded frames");
  if (!frames || !clip) return fail(HPVG_E_ARG, "frames_to_clip: null pointer");
  KL(hpvg::ew_frames_to_clip(frames, Hs, Ws, bgr ? 1 : 0, start, every, T, H, W, hflip ? 1 : 0, clip, S(st)), 1);
  return HPVG_OK;
}
int hpvg_box_muller_inplace(float* z, long long n, void* st) {
  if (n <= 0) return HPVG_OK;
  if (!z) return fail(HPVG_E_ARG, "box_muller_inplace: null pointer");
  KL(hpvg::ew_box_muller_inplace(z, n, S(st)), 1);
  return HPVG_OK;
}
int hpvg_randn(float* z, long long n, uint64_t seed, uint64_t offset, const uint64_t* d_offset, void* st) {
  if (n <= 0) return HPVG_OK;
  if (!z) return fail(HPVG_E_ARG, "randn: null pointer");
  KL(hpvg::ew_randn(z, n, seed, offset, reinterpret_cast<const unsigned long long*>(d_offset), S(st)), 1);
  return HPVG_OK;
}
int hpvg_counter_add(uint64_t* d_counter, uint64_t inc, void* st) {
  if (!d_counter) return fail(HPVG_E_ARG, "counter_add: null pointer");
  KL(hpvg::ew_counter_add(reinterpret_cast<unsigned long long*>(d_counter), inc, S(st)), 1);
  return HPVG_OK;
}
int hpvg_gather_strided(const float* src, long long n, long long stride, long long offset, float* dst, void* st) {
  if (n <= 0) return HPVG_OK;
  if (!src || !dst || stride < 1 || offset < 0) return fail(HPVG_E_ARG, "gather_strided: bad arguments");
  KL(hpvg::ew_gather_strided(src, n, stride, offset, dst, S(st)), 1);
  return HPVG_OK;
}
int hpvg_channel_sum(const float* g, int N, int C, long long sp, int accumulate, float* out, void* st) {
  if (N <= 0 || C <= 0 || sp <= 0) return fail(HPVG_E_ARG, "channel_sum: empty input");
  CTX(cx, st);
  if (C > 128) return fail(HPVG_E_ARG, "channel_sum: at most 128 channels");
  KL(hpvg::ew_channel_sum_ncdhw(g, N, C, sp, accumulate, cx->det.partials, out, S(st)), 2);
  return HPVG_OK;
}
int hpvg_kl_grad(const float* mu, const float* lv, long long n, float coef, float* gmu, float* glv, void* st) {
  if (n <= 0) return HPVG_OK;
  KL(hpvg::ew_kl_grad(mu, lv, n, coef, gmu, glv, S(st)), 1);
  return HPVG_OK;
}
int hpvg_sn_grad(const float* G, const float* w, const float* u, const float* v, const float* sigma, int cout, int k,
                 int accumulate, float* gw, void* st) {
  if (cout <= 0 || k <= 0) return fail(HPVG_E_ARG, "sn_grad: empty matrix");
  CTX(cx, st);
  KL(hpvg::ew_sn_grad(G, w, u, v, sigma, cout, k, accumulate, cx->fscratch, gw, S(st)), 2);
  return HPVG_OK;
}
int hpvg_lerp(const float* a, const float* b, float alpha, long long n, float* out, void* st) {
  if (n <= 0) return HPVG_OK;
  KL(hpvg::ew_lerp(a, b, alpha, n, out, S(st)), 1);
  return HPVG_OK;
}
int hpvg_gp_grad(const float* g, int N, int C, long long sp, float lambda, float* Gout, float* gp, void* st) {
  if (N <= 0 || C <= 0 || sp <= 0) return fail(HPVG_E_ARG, "gp_grad: empty input");
  CTX(cx, st);
  KL(hpvg::ew_gp_grad(g, N, C, sp, lambda, Gout, gp, cx->det, S(st)), 1);
  return HPVG_OK;
}

}  // extern "C"
// ------------------------------------------------------------------------------------------------ fused sampling entry
// GeneratorHPVAEGAN.construct(noise_init=z, isRandom=True) in eval mode as ONE call (include/hpvg.h): the launches the
// Python layer (hpvg/networks_3d.py: construct / refinement_layers / _run_block) makes, in the same order with the same
// arguments — so the result is bit-identical to that path — without a host round trip per launch.
namespace {
struct SampleWs {
  size_t zcl, raw, act[2], up, xin, lvl[2], total;
};
inline size_t al256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }
SampleWs sample_ws_layout(const HpvgGenerator* g, int N) {
  size_t vmax = 0;
  for (int l = 0; l <= g->n_stages; ++l) {
    const size_t v = static_cast<size_t>(g->T[l]) * g->H[l] * g->W[l];
    if (v > vmax) vmax = v;
  }
  const size_t v0 = static_cast<size_t>(g->T[0]) * g->H[0] * g->W[0], n = static_cast<size_t>(N);
  SampleWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += al256(bytes);
    return o;
  };
  w.zcl = take(n * v0 * g->latent_dim * 2);
  w.raw = take(n * v0 * 64 * 4);
  w.act[0] = take(n * vmax * 64 * 2);
  w.act[1] = take(n * vmax * 64 * 2);
  w.up = take(n * vmax * g->nc_im * 4);
  w.xin = take(n * vmax * 16);
  w.lvl[0] = take(n * vmax * g->nc_im * 4);
  w.lvl[1] = take(n * vmax * g->nc_im * 4);
  w.total = off;
  return w;
}
// one block: hidden layers (conv + folded BN + LeakyReLU, bf16 channels-last ping-pong) and the tail
// (64 -> nc_im, + residual, tanh, fp32 ncdhw)
int sample_block(const HpvgBlock& b, int nc_im, int N, int T, int H, int W, const void* x, int x_pitch, char* act0,
                 char* act1, float* raw, const float* residual, float* out, void* st) {
  if (b.n_layers < 2 || b.n_layers > HPVG_BLOCK_LAYERS) return fail(HPVG_E_ARG, "generator_sample: 2..8 layers per block");
  const void* h = x;
  int pitch = x_pitch;
  for (int l = 0; l < b.n_layers - 1; ++l) {
    char* o = (l & 1) ? act1 : act0;
    int rc;
    if (b.cin[l] == 128) {   // two 64-channel halves: the first leaves raw fp32 partial sums, the second adds them
      rc = hpvg_conv_cl(HPVG_CONV_64_64, N, T, H, W, h, pitch, b.wimg[l][0], b.scale[l], b.shift[l], HPVG_ACT_NONE,
                        HPVG_OUT_F32_RAW, raw, 64, 0, 64, nullptr, nullptr, nullptr, 0, st);
      if (rc != HPVG_OK) return rc;
      rc = hpvg_conv_cl(HPVG_CONV_64_64, N, T, H, W, static_cast<const char*>(h) + 64 * 2, pitch, b.wimg[l][1],
                        b.scale[l], b.shift[l], HPVG_ACT_LRELU, HPVG_OUT_BF16_CL, o, 64, 0, 64, raw, nullptr, nullptr, 0,
                        st);
    } else if (b.cin[l] == 64 || b.cin[l] <= 8) {
      rc = hpvg_conv_cl(b.cin[l] == 64 ? HPVG_CONV_64_64 : HPVG_CONV_8_64, N, T, H, W, h, pitch, b.wimg[l][0],
                        b.scale[l], b.shift[l], HPVG_ACT_LRELU, HPVG_OUT_BF16_CL, o, 64, 0, 64, nullptr, nullptr,
                        nullptr, 0, st);
    } else {
      return fail(HPVG_E_ARG, "generator_sample: layer input channels must be <= 8, 64 or 128");
    }
    if (rc != HPVG_OK) return rc;
    h = o;
    pitch = 64;
  }
  const int t = b.n_layers - 1;
  return hpvg_conv_cl(nc_im <= 3 ? HPVG_CONV_64_3 : HPVG_CONV_64_16, N, T, H, W, h, pitch, b.wimg[t][0], b.scale[t],
                      b.shift[t], HPVG_ACT_TANH, HPVG_OUT_F32_NCDHW, out, 64, 0, nc_im, residual, nullptr, nullptr, 0,
                      st);
}
}  // namespace
extern "C" {
size_t hpvg_block_fwd_eval_workspace(const HpvgBlock* b, int N, int T, int H, int W) {
  if (!b || N <= 0 || T <= 0 || H <= 0 || W <= 0) return 0;
  const size_t v = static_cast<size_t>(N) * T * H * W;
  return 2 * al256(v * 64 * 2) + (b->cin[0] == 128 ? al256(v * 64 * 4) : 0);
}
int hpvg_block_fwd_eval(const HpvgBlock* b, int nc_im, int N, int T, int H, int W, const void* x, int x_pitch,
                        const float* residual, float* out, void* workspace, size_t workspace_bytes, void* st) {
  if (!b || !x || !out || !workspace) return fail(HPVG_E_ARG, "block_fwd_eval: null argument");
  if (N <= 0) return HPVG_OK;
  if (nc_im < 1 || nc_im > 4) return fail(HPVG_E_ARG, "block_fwd_eval: nc_im 1..4");
  if (workspace_bytes < hpvg_block_fwd_eval_workspace(b, N, T, H, W)) return fail(HPVG_E_ARG, "block_fwd_eval: workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(HPVG_E_ARG, "block_fwd_eval: workspace not 256-byte aligned");
  const size_t a = al256(static_cast<size_t>(N) * T * H * W * 64 * 2);
  char* ws = static_cast<char*>(workspace);
  return sample_block(*b, nc_im, N, T, H, W, x, x_pitch, ws, ws + a, reinterpret_cast<float*>(ws + 2 * a), residual, out,
                      st);
}
size_t hpvg_generator_sample_workspace(const HpvgGenerator* g, int N) {
  if (!g || N <= 0 || g->n_stages < 0 || g->n_stages >= HPVG_MAX_LEVELS) return 0;
  return sample_ws_layout(g, N).total;
}
int hpvg_generator_sample(const HpvgGenerator* g, const float* z, int N, uint64_t sample_base,
                          const uint64_t* d_sample_offset, float* out, float* vae_out, void* workspace,
                          size_t workspace_bytes, void* st) {
  if (!g || !z || !out || !workspace) return fail(HPVG_E_ARG, "generator_sample: null argument");
  if (N <= 0) return HPVG_OK;
  if (g->n_stages < 0 || g->n_stages >= HPVG_MAX_LEVELS) return fail(HPVG_E_ARG, "generator_sample: 0..15 stages");
  if (g->nc_im < 1 || g->nc_im > 4 || (g->latent_dim != 64 && g->latent_dim != 128))
    return fail(HPVG_E_ARG, "generator_sample: nc_im 1..4, latent_dim 64 or 128");
  const SampleWs w = sample_ws_layout(g, N);
  if (workspace_bytes < w.total) return fail(HPVG_E_ARG, "generator_sample: workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(HPVG_E_ARG, "generator_sample: workspace not 256-byte aligned");
  char* ws = static_cast<char*>(workspace);
  // ---- decoder on z (networks_3d.py:419-423)
  int rc = hpvg_pack_cl(z, N, g->latent_dim, g->T[0], g->H[0], g->W[0], ws + w.zcl, g->latent_dim, 0, g->latent_dim, st);
  if (rc != HPVG_OK) return rc;
  float* lvl[2] = {reinterpret_cast<float*>(ws + w.lvl[0]), reinterpret_cast<float*>(ws + w.lvl[1])};
  float* prev = (g->n_stages == 0) ? out : (vae_out ? vae_out : lvl[0]);
  rc = sample_block(g->decoder, g->nc_im, N, g->T[0], g->H[0], g->W[0], ws + w.zcl, g->latent_dim, ws + w.act[0],
                    ws + w.act[1], reinterpret_cast<float*>(ws + w.raw), nullptr, prev, st);
  if (rc != HPVG_OK) return rc;
  if (g->n_stages == 0 && vae_out) {
    const size_t bytes = static_cast<size_t>(N) * g->nc_im * g->T[0] * g->H[0] * g->W[0] * 4;
    CU(cudaMemcpyAsync(vae_out, out, bytes, cudaMemcpyDeviceToDevice, S(st)));
  }
  // ---- refinement stages (networks_3d.py:434-451)
  float* up = reinterpret_cast<float*>(ws + w.up);
  for (int s = 0; s < g->n_stages; ++s) {
    const int l = s + 1;
    const float amp = g->noise_amp[l];
    rc = hpvg_upsample_noise_pack(prev, N, g->nc_im, g->T[s], g->H[s], g->W[s], g->T[l], g->H[l], g->W[l], nullptr, amp,
                                  amp != 0.f ? g->noise_seed[l] : 0, sample_base, d_sample_offset, up, ws + w.xin, st);
    if (rc != HPVG_OK) return rc;
    float* o = (s == g->n_stages - 1) ? out : lvl[l & 1];
    rc = sample_block(g->body[s], g->nc_im, N, g->T[l], g->H[l], g->W[l], ws + w.xin, 8, ws + w.act[0], ws + w.act[1],
                      nullptr, up, o, st);
    if (rc != HPVG_OK) return rc;
    prev = o;
  }
  return HPVG_OK;
}
}  // extern "C"
extern "C" {
// ------------------------------------------------------------------------------------------------ MindSpore AOT
// Entry points with the signature MindSpore's ops.Custom(func_type="aot") calls:
//     int f(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes, void* stream, void* extra)
// params = inputs then outputs, all fp32 device tensors in the reference's layouts (NCDHW / parameter shapes), all
// owned by the caller.  Nothing allocates per call and nothing synchronises: the bf16 channels-last staging buffers,
// packed filter banks and epilogue vectors live in the calling stream's grow-only workspace (arena_for) — the first
// call at a new size grows it (one stream synchronise), every later call reuses it.  0 = success; any other value makes
// MindSpore raise (hpvg_last_error() has the reason).  Where the reference's cell updates a Parameter in place
// (BatchNorm moving statistics) the entry point writes through the corresponding INPUT pointer, as the MindSpore kernel
// does.  INTEGRATION.md lists each entry with the Custom(...) declaration that binds it.
}  // extern "C" (helpers with C++ linkage follow)

namespace {
bool is_f32(const char* s) { return s && std::strcmp(s, "float32") == 0; }
struct Carver {     // sequential 256-byte aligned sub-buffers of one workspace
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base(static_cast<char*>(b)) {}
  template <typename T>
  T* take(size_t bytes) {
    T* p = reinterpret_cast<T*>(base + off);
    off += (bytes + 255) & ~static_cast<size_t>(255);
    return p;
  }
};
inline size_t al(size_t b) { return (b + 255) & ~static_cast<size_t>(255); }
bool all_f32(int n, const char** dtypes) {
  for (int i = 0; i < n; ++i)
    if (!is_f32(dtypes[i])) return false;
  return true;
}
int aot_bad(const char* msg) {
  g_err = msg;
  return 1;
}

// y = act(conv3x3x3(x, w) + b): (Cin <= 8 | 64) -> 64 with fp32 NCDHW in/out, or 64 -> (<= 3) (tail; tanh/none)
int aot_conv(int act, int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes, void* stream) {
  if (nparam != 4 || !all_f32(4, dtypes) || ndims[0] != 5 || ndims[1] != 5 || ndims[2] != 1 || ndims[3] != 5)
    return aot_bad("conv: expects (x[N,Cin,T,H,W], w[Cout,Cin,3,3,3], b[Cout]) -> y[N,Cout,T,H,W], all float32");
  const int64_t* a = shapes[0];
  const int64_t* ws = shapes[1];
  const int N = (int)a[0], Cin = (int)a[1], T = (int)a[2], H = (int)a[3], W = (int)a[4];
  const int Cout = (int)ws[0];
  if (ws[1] != Cin || ws[2] != 3 || ws[3] != 3 || ws[4] != 3 || shapes[2][0] != Cout || shapes[3][1] != Cout ||
      shapes[3][0] != N || shapes[3][2] != T || shapes[3][3] != H || shapes[3][4] != W)
    return aot_bad("conv: inconsistent shapes");
  const bool head = Cin <= 8 && Cout == 64, body = Cin == 64 && Cout == 64, tail = Cin == 64 && Cout <= 3;
  if (!head && !body && !tail) return aot_bad("conv: supported channel shapes are (<=8|64)->64 and 64->(<=3)");
  const size_t vox = (size_t)N * T * H * W;
  const int mode = head ? HPVG_CONV_8_64 : (body ? HPVG_CONV_64_64 : HPVG_CONV_64_3);
  const int in_pitch = head ? 8 : 64;
  const size_t need = al(vox * in_pitch * 2) + al(tail ? 0 : vox * 128) + al(hpvg_conv_wimg_bytes(mode)) + al(512);
  void* wsp = arena_for(stream, need);
  if (!wsp) return 3;
  Carver cv(wsp);
  void* xcl = cv.take<void>(vox * in_pitch * 2);
  void* ycl = cv.take<void>(tail ? 0 : vox * 128);
  void* wimg = cv.take<void>(hpvg_conv_wimg_bytes(mode));
  float* aff = cv.take<float>(512);
  int rc = hpvg_pack_cl(static_cast<const float*>(params[0]), N, Cin, T, H, W, xcl, in_pitch, 0, in_pitch, stream);
  if (!rc) rc = hpvg_conv_pack_weights(static_cast<const float*>(params[1]), Cout, Cin, 3, mode, 0, 0, Cout, 0, Cin, wimg,
                                       stream);
  if (!rc) rc = hpvg_memset(aff, 0, 512, stream);
  if (!rc) rc = hpvg_affine_from_bias(static_cast<const float*>(params[2]), nullptr, Cout, aff, aff + 64, stream);
  if (tail) {
    if (!rc) rc = hpvg_conv_cl(mode, N, T, H, W, xcl, 64, wimg, aff, aff + 64, act, HPVG_OUT_F32_NCDHW, params[3], 64, 0,
                               Cout, nullptr, nullptr, nullptr, 0, stream);
  } else {
    if (!rc) rc = hpvg_conv_cl(mode, N, T, H, W, xcl, in_pitch, wimg, aff, aff + 64, act, HPVG_OUT_BF16_CL, ycl, 64, 0, 64,
                               nullptr, nullptr, nullptr, 0, stream);
    if (!rc) rc = hpvg_unpack_cl(ycl, N, 64, T, H, W, 64, 0, static_cast<float*>(params[3]), stream);
  }
  return rc;
}
}  // namespace

extern "C" {

int HpvgUpsampleTrilinear3D(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                            void* stream, void* /*extra*/) {
  if (nparam != 2 || ndims[0] != 5 || ndims[1] != 5 || !is_f32(dtypes[0]) || !is_f32(dtypes[1])) return 1;
  const int64_t* a = shapes[0];
  const int64_t* b = shapes[1];
  if (a[0] != b[0] || a[1] != b[1]) return 2;
  return hpvg_resize3d_fwd(static_cast<const float*>(params[0]), (int)a[0], (int)a[1], (int)a[2], (int)a[3],
                           (int)a[4], static_cast<float*>(params[1]), (int)b[2], (int)b[3], (int)b[4], 1, stream);
}
int HpvgUpsampleTrilinear3DGrad(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                                void* stream, void* /*extra*/) {
  // inputs: dy (N,C,To,Ho,Wo), x (N,C,Ti,Hi,Wi) [shape only]; output: dx
  if (nparam != 3 || ndims[0] != 5 || ndims[2] != 5 || !is_f32(dtypes[0]) || !is_f32(dtypes[2])) return 1;
  const int64_t* a = shapes[0];
  const int64_t* b = shapes[2];
  return hpvg_resize3d_bwd(static_cast<const float*>(params[0]), (int)a[0], (int)a[1], (int)a[2], (int)a[3],
                           (int)a[4], static_cast<float*>(params[2]), (int)b[2], (int)b[3], (int)b[4], 1, stream);
}

// nn.Conv3d(+bias) [+ LeakyReLU(0.2) | tanh] of networks_3d.py:48-53, 380, 399: (x, w, b) -> y
int HpvgConv3dBias(int n, void** p, int* nd, int64_t** sh, const char** dt, void* st, void*) {
  return aot_conv(HPVG_ACT_NONE, n, p, nd, sh, dt, st);
}
int HpvgConv3dBiasLRelu(int n, void** p, int* nd, int64_t** sh, const char** dt, void* st, void*) {
  return aot_conv(HPVG_ACT_LRELU, n, p, nd, sh, dt, st);
}
int HpvgConv3dBiasTanh(int n, void** p, int* nd, int64_t** sh, const char** dt, void* st, void*) {
  return aot_conv(HPVG_ACT_TANH, n, p, nd, sh, dt, st);
}

// bprop of y = LeakyReLU(conv(x, w) + b) for the 64 -> 64 layer: (x, w, y, dy) -> (dx, dw, db).
// The data gradient is the forward kernel on the transposed / mirrored bank (dgrad), the weight gradient the tcgen05
// wgrad kernel, the bias gradient a column sum; y (the stored activation) supplies the LeakyReLU mask.
int HpvgConv3dBiasLReluGrad(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes, void* stream,
                            void* /*extra*/) {
  if (nparam != 7 || !all_f32(7, dtypes) || ndims[0] != 5 || ndims[1] != 5 || ndims[2] != 5 || ndims[3] != 5)
    return aot_bad("conv grad: expects (x, w, y, dy) -> (dx, dw, db), float32");
  const int64_t* a = shapes[0];
  if (a[1] != 64 || shapes[1][0] != 64 || shapes[1][1] != 64 || shapes[2][1] != 64 || shapes[3][1] != 64 ||
      shapes[4][1] != 64 || shapes[5][0] != 64 || shapes[6][0] != 64)
    return aot_bad("conv grad: the 64 -> 64 layer only");
  const int N = (int)a[0], T = (int)a[2], H = (int)a[3], W = (int)a[4];
  const size_t vox = (size_t)N * T * H * W, act_b = vox * 128;
  const size_t need = 4 * al(act_b) + al(hpvg_conv_wimg_bytes(HPVG_CONV_64_64)) + al(512);
  void* wsp = arena_for(stream, need);
  if (!wsp) return 3;
  Carver cv(wsp);
  void* xcl = cv.take<void>(act_b);
  void* ycl = cv.take<void>(act_b);
  void* gcl = cv.take<void>(act_b);      // dy, then gz = dy * LeakyReLU'(y) in place
  void* dxcl = cv.take<void>(act_b);
  void* wimg = cv.take<void>(hpvg_conv_wimg_bytes(HPVG_CONV_64_64));
  float* aff = cv.take<float>(512);
  const float* x = static_cast<const float*>(params[0]);
  const float* w = static_cast<const float*>(params[1]);
  int rc = hpvg_pack_cl(x, N, 64, T, H, W, xcl, 64, 0, 64, stream);
  if (!rc) rc = hpvg_pack_cl(static_cast<const float*>(params[2]), N, 64, T, H, W, ycl, 64, 0, 64, stream);
  if (!rc) rc = hpvg_pack_cl(static_cast<const float*>(params[3]), N, 64, T, H, W, gcl, 64, 0, 64, stream);
  if (!rc) rc = hpvg_lrelu_bwd_cl(gcl, ycl, (long long)vox * 64, gcl, stream);
  if (!rc) rc = hpvg_colsum_cl(gcl, (long long)vox, static_cast<float*>(params[6]), 0, stream);
  if (!rc) rc = hpvg_conv_wgrad_cl(xcl, 64, gcl, 64, N, T, H, W, static_cast<float*>(params[5]), 64, 3, 0, 64, 0, 64, 0,
                                   1.0f, stream);
  if (!rc) rc = hpvg_conv_pack_weights(w, 64, 64, 3, HPVG_CONV_64_64, 1, 0, 64, 0, 64, wimg, stream);
  if (!rc) rc = hpvg_fill(aff, 1.0f, 64, stream);
  if (!rc) rc = hpvg_memset(aff + 64, 0, 256, stream);
  if (!rc) rc = hpvg_conv_cl(HPVG_CONV_64_64, N, T, H, W, gcl, 64, wimg, aff, aff + 64, HPVG_ACT_NONE, HPVG_OUT_BF16_CL,
                             dxcl, 64, 0, 64, nullptr, nullptr, nullptr, 0, stream);
  if (!rc) rc = hpvg_unpack_cl(dxcl, N, 64, T, H, W, 64, 0, static_cast<float*>(params[4]), stream);
  return rc;
}

// Training-mode nn.BatchNorm3d(64) + LeakyReLU (networks_3d.py:52-53):
//   (x, gamma, beta, moving_mean, moving_var) -> (y, saved[4,64] = scale, shift, mean, invstd)
// moving_mean / moving_var are updated IN PLACE (momentum 0.9), like the Parameters of the reference's cell.
int HpvgBatchNorm3dLReluTrain(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes, void* stream,
                              void* /*extra*/) {
  if (nparam != 7 || !all_f32(7, dtypes) || ndims[0] != 5 || ndims[5] != 5 || shapes[0][1] != 64 || shapes[1][0] != 64 ||
      ndims[6] != 2 || shapes[6][0] != 4 || shapes[6][1] != 64)
    return aot_bad("batchnorm train: expects (x[N,64,T,H,W], gamma, beta, moving_mean, moving_var) -> (y, saved[4,64])");
  const int64_t* a = shapes[0];
  const int N = (int)a[0], T = (int)a[2], H = (int)a[3], W = (int)a[4];
  const size_t vox = (size_t)N * T * H * W;
  void* wsp = arena_for(stream, 2 * al(vox * 128) + al(1024));
  if (!wsp) return 3;
  Carver cv(wsp);
  void* xcl = cv.take<void>(vox * 128);
  void* ycl = cv.take<void>(vox * 128);
  double* sums = cv.take<double>(1024);
  int rc = hpvg_pack_cl(static_cast<const float*>(params[0]), N, 64, T, H, W, xcl, 64, 0, 64, stream);
  if (!rc) rc = hpvg_bn_stats_cl(xcl, (long long)vox, sums, sums + 64, stream);
  if (!rc) rc = hpvg_bn_train_apply_cl(xcl, (long long)vox, sums, static_cast<const float*>(params[1]),
                                       static_cast<const float*>(params[2]), 1e-5f, 0.9f, static_cast<float*>(params[3]),
                                       static_cast<float*>(params[4]), static_cast<float*>(params[6]), HPVG_ACT_LRELU, ycl,
                                       nullptr, stream);
  if (!rc) rc = hpvg_unpack_cl(ycl, N, 64, T, H, W, 64, 0, static_cast<float*>(params[5]), stream);
  return rc;
}
// its bprop: (dy, x, saved) -> (dx, dgamma, dbeta)
int HpvgBatchNorm3dLReluTrainGrad(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                                  void* stream, void* /*extra*/) {
  if (nparam != 6 || !all_f32(6, dtypes) || ndims[0] != 5 || ndims[1] != 5 || shapes[0][1] != 64 || shapes[1][1] != 64 ||
      ndims[2] != 2 || shapes[2][0] != 4 || shapes[2][1] != 64 || shapes[4][0] != 64 || shapes[5][0] != 64)
    return aot_bad("batchnorm grad: expects (dy, x, saved[4,64]) -> (dx, dgamma[64], dbeta[64])");
  const int64_t* a = shapes[0];
  const int N = (int)a[0], T = (int)a[2], H = (int)a[3], W = (int)a[4];
  const size_t vox = (size_t)N * T * H * W;
  void* wsp = arena_for(stream, 3 * al(vox * 128));
  if (!wsp) return 3;
  Carver cv(wsp);
  void* gcl = cv.take<void>(vox * 128);
  void* xcl = cv.take<void>(vox * 128);
  void* dxcl = cv.take<void>(vox * 128);
  int rc = hpvg_pack_cl(static_cast<const float*>(params[0]), N, 64, T, H, W, gcl, 64, 0, 64, stream);
  if (!rc) rc = hpvg_pack_cl(static_cast<const float*>(params[1]), N, 64, T, H, W, xcl, 64, 0, 64, stream);
  if (!rc) rc = hpvg_bn_bwd_cl(gcl, xcl, (long long)vox, static_cast<const float*>(params[2]), HPVG_ACT_LRELU, dxcl,
                               static_cast<float*>(params[4]), static_cast<float*>(params[5]), 0, stream);
  if (!rc) rc = hpvg_unpack_cl(dxcl, N, 64, T, H, W, 64, 0, static_cast<float*>(params[3]), stream);
  return rc;
}

// One power iteration of SpectualNormConv3d (spectral_norm.py:146-151): (w, u, v) -> (sigma2 = [sigma, 1/sigma], u', v')
int HpvgSpectralNormIter(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes, void* stream,
                         void* /*extra*/) {
  if (nparam != 6 || !all_f32(6, dtypes) || ndims[0] < 2) return aot_bad("sn iter: expects (w, u, v) -> (sigma2, u', v')");
  const int cout = (int)shapes[0][0];
  long long k = 1;
  for (int i = 1; i < ndims[0]; ++i) k *= shapes[0][i];
  long long nu = 1, nv = 1;
  for (int i = 0; i < ndims[1]; ++i) nu *= shapes[1][i];
  for (int i = 0; i < ndims[2]; ++i) nv *= shapes[2][i];
  if (nu != cout || nv != k || shapes[3][0] != 2) return aot_bad("sn iter: u must have Cout, v Cin*taps, sigma2 2 elements");
  float* s2 = static_cast<float*>(params[3]);
  int rc = hpvg_d2d(params[4], params[1], (size_t)nu * 4, stream);
  if (!rc) rc = hpvg_d2d(params[5], params[2], (size_t)nv * 4, stream);
  if (!rc) rc = hpvg_sn_power_iter(static_cast<const float*>(params[0]), cout, (int)k, static_cast<float*>(params[4]),
                                   static_cast<float*>(params[5]), s2, s2 + 1, stream);
  return rc;
}

// ClippedAdam step of ONE parameter tensor (optimizers.py:41-43 + mindspore.nn.Adam):
//   (param, grad, m, v, hyper[6] = lr, beta1, beta2, eps, clip (0 = none), step (1-based)) -> (param', m', v')
// hyper is a DEVICE tensor (MindSpore keeps the learning rate and the global step as Tensors).
int HpvgClipAdam(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes, void* stream,
                 void* /*extra*/) {
  if (nparam != 8 || !all_f32(8, dtypes) || ndims[4] != 1 || shapes[4][0] != 6)
    return aot_bad("clip adam: expects (param, grad, m, v, hyper[6]) -> (param', m', v'), float32");
  long long n = 1;
  for (int i = 0; i < ndims[0]; ++i) n *= shapes[0][i];
  CTX(cx, stream);
  int rc = hpvg_d2d(params[5], params[0], (size_t)n * 4, stream);
  if (!rc) rc = hpvg_d2d(params[6], params[2], (size_t)n * 4, stream);
  if (!rc) rc = hpvg_d2d(params[7], params[3], (size_t)n * 4, stream);
  if (rc) return rc;
  hpvg::AdamTable tab;
  std::memset(&tab, 0, sizeof(tab));
  tab.p[0] = static_cast<float*>(params[5]);
  tab.g[0] = static_cast<const float*>(params[1]);
  tab.m[0] = static_cast<float*>(params[6]);
  tab.v[0] = static_cast<float*>(params[7]);
  tab.n[0] = n;
  KL(hpvg::ew_adam_clip(tab, 1, cx->adam_norms, 0.f, 0.f, 0.f, 1.f, 1.f, nullptr, S(stream),
                        static_cast<const float*>(params[4])), 2);
  return HPVG_OK;
}

// losses (losses.py:5-7, nn.MSELoss): (a, b) -> loss[1] ; (mu, logvar) -> loss[1]
int HpvgMSELoss(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes, void* stream, void*) {
  if (nparam != 3 || !all_f32(3, dtypes)) return aot_bad("mse: expects (a, b) -> loss[1]");
  long long n = 1, m = 1;
  for (int i = 0; i < ndims[0]; ++i) n *= shapes[0][i];
  for (int i = 0; i < ndims[1]; ++i) m *= shapes[1][i];
  if (n != m) return aot_bad("mse: operand sizes differ");
  return hpvg_mse(static_cast<const float*>(params[0]), static_cast<const float*>(params[1]), n,
                  static_cast<float*>(params[2]), stream);
}
int HpvgKLLoss(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes, void* stream, void*) {
  if (nparam != 3 || !all_f32(3, dtypes)) return aot_bad("kl: expects (mu, logvar) -> loss[1]");
  long long n = 1, m = 1;
  for (int i = 0; i < ndims[0]; ++i) n *= shapes[0][i];
  for (int i = 0; i < ndims[1]; ++i) m *= shapes[1][i];
  if (n != m) return aot_bad("kl: operand sizes differ");
  return hpvg_kl(static_cast<const float*>(params[0]), static_cast<const float*>(params[1]), n,
                 static_cast<float*>(params[2]), stream);
}

}  // extern "C"
