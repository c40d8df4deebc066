// Hardware probes for the tcgen05 / TMA semantics the convolution kernels rely on (development tool, not product).
//
//   hpvg_probe_umma : copy an arbitrary byte image into shared memory, issue a host-specified list of
//                     tcgen05.mma ops (descriptors built on the host, relative to the image start), dump TMEM.
//   hpvg_probe_tma  : encode a tiled tensor map from host-specified dims/strides/box/swizzle, load one box at
//                     host-specified (possibly negative) coordinates, dump the shared-memory bytes.
//
// Built as tools/probe/libhpvg_probe.so and driven by tools/probe/run_probe.py via ctypes.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../mindspore-hp-vae-gan_b200/csrc/ptx.cuh"

using namespace hpvg;

struct ProbeOp {
  uint64_t a_desc;
  uint64_t b_desc;
  uint32_t idesc;
  uint32_t accumulate;
  uint32_t d_col;
  uint32_t pad;
};

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      return -1;                                                                                \
    }                                                                                           \
  } while (0)

__global__ void __launch_bounds__(128, 1)
probe_umma_kernel(const uint8_t* __restrict__ image, int image_bytes, const ProbeOp* __restrict__ ops, int n_ops,
                  int n_cols, int alloc_cols, float* __restrict__ out, int repeat, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_base_s;

  // 1024-align the image inside dynamic smem
  uint32_t base = smem_u32(smem_raw);
  uint32_t aligned = (base + 1023u) & ~1023u;
  uint8_t* img = smem_raw + (aligned - base);

  for (int i = threadIdx.x * 16; i < image_bytes; i += blockDim.x * 16) {
    *reinterpret_cast<uint4*>(img + i) = *reinterpret_cast<const uint4*>(image + i);
  }
  fence_proxy_async();  // generic-proxy smem writes -> visible to the async proxy (UMMA reads)
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, alloc_cols);
  if (threadIdx.x == 32) {
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    if (elect_one()) {
      t0 = clock64();
      const uint64_t add = static_cast<uint64_t>(aligned >> 4);
      for (int r = 0; r < repeat; ++r) {
        for (int i = 0; i < n_ops; ++i) {
          ProbeOp op = ops[i];
          // start-address field is bits [0,14): image-relative -> absolute
          uint64_t a = (op.a_desc & ~0x3FFFull) | (((op.a_desc & 0x3FFF) + add) & 0x3FFF);
          uint64_t b = (op.b_desc & ~0x3FFFull) | (((op.b_desc & 0x3FFF) + add) & 0x3FFF);
          umma_bf16(tmem_base + op.d_col, a, b, op.idesc, (r > 0) ? 1u : op.accumulate);
        }
      }
      umma_commit(&done_bar);
    }
    __syncwarp();
  }
  mbar_wait(&done_bar, 0);
  t1 = clock64();
  tc_fence_after();
  if (threadIdx.x == 32 && cycles) *cycles = t1 - t0;  // lane 0 of warp 1 may not be the elected lane; see below

  // dump: warp w owns lanes [32w, 32w+32)
  for (int c = 0; c < n_cols; c += 8) {
    uint32_t r[8];
    tmem_ld8(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    const int lane = warp * 32 + (threadIdx.x & 31);
#pragma unroll
    for (int j = 0; j < 8; ++j) out[lane * n_cols + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, alloc_cols);
}

// Timing variant: issue `repeat` x n_ops MMAs from one thread and report cycles from first issue to completion,
// measured by the issuing thread itself.
__global__ void __launch_bounds__(128, 1)
probe_umma_time_kernel(const ProbeOp* __restrict__ ops, int n_ops, int alloc_cols, int repeat, int smem_used,
                       long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_base_s;
  uint32_t base = smem_u32(smem_raw);
  uint32_t aligned = (base + 1023u) & ~1023u;
  uint8_t* img = smem_raw + (aligned - base);
  for (int i = threadIdx.x * 16; i < smem_used; i += blockDim.x * 16) *reinterpret_cast<uint4*>(img + i) = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, alloc_cols);
  if (threadIdx.x == 32) {
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 32) {
    const uint64_t add = static_cast<uint64_t>(aligned >> 4);
    long long t0 = clock64();
    for (int r = 0; r < repeat; ++r) {
      for (int i = 0; i < n_ops; ++i) {
        ProbeOp op = ops[i];
        uint64_t a = (op.a_desc & ~0x3FFFull) | (((op.a_desc & 0x3FFF) + add) & 0x3FFF);
        uint64_t b = (op.b_desc & ~0x3FFFull) | (((op.b_desc & 0x3FFF) + add) & 0x3FFF);
        umma_bf16(tmem_base + op.d_col, a, b, op.idesc, 1u);
      }
    }
    umma_commit(&done_bar);
    mbar_wait(&done_bar, 0);
    long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, alloc_cols);
}

extern "C" int hpvg_probe_umma(const void* image, int image_bytes, const ProbeOp* ops, int n_ops, int n_cols,
                               float* out /* [128][n_cols] host */) {
  int alloc_cols = 32;
  while (alloc_cols < n_cols) alloc_cols <<= 1;
  uint8_t* d_img;
  ProbeOp* d_ops;
  float* d_out;
  int padded = (image_bytes + 15) & ~15;
  CK(cudaMalloc(&d_img, padded));
  CK(cudaMemset(d_img, 0, padded));
  CK(cudaMemcpy(d_img, image, image_bytes, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_ops, sizeof(ProbeOp) * n_ops));
  CK(cudaMemcpy(d_ops, ops, sizeof(ProbeOp) * n_ops, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_out, sizeof(float) * 128 * n_cols));
  CK(cudaMemset(d_out, 0xFF, sizeof(float) * 128 * n_cols));
  int smem = padded + 1024;
  CK(cudaFuncSetAttribute(probe_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_umma_kernel<<<1, 128, smem>>>(d_img, padded, d_ops, n_ops, n_cols, alloc_cols, d_out, 1, nullptr);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out, d_out, sizeof(float) * 128 * n_cols, cudaMemcpyDeviceToHost));
  cudaFree(d_img);
  cudaFree(d_ops);
  cudaFree(d_out);
  return 0;
}

// returns average cycles per MMA (x1000) over `repeat` x n_ops, on `grid` concurrent CTAs (max over CTAs)
extern "C" long long hpvg_probe_umma_time(const ProbeOp* ops, int n_ops, int alloc_cols, int repeat, int smem_used,
                                          int grid) {
  ProbeOp* d_ops;
  long long* d_cyc;
  if (cudaMalloc(&d_ops, sizeof(ProbeOp) * n_ops) != cudaSuccess) return -1;
  cudaMemcpy(d_ops, ops, sizeof(ProbeOp) * n_ops, cudaMemcpyHostToDevice);
  cudaMalloc(&d_cyc, sizeof(long long) * grid);
  int smem = smem_used + 1024;
  if (cudaFuncSetAttribute(probe_umma_time_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -2;
  for (int it = 0; it < 2; ++it) probe_umma_time_kernel<<<grid, 128, smem>>>(d_ops, n_ops, alloc_cols, repeat, smem_used, d_cyc);
  if (cudaDeviceSynchronize() != cudaSuccess) {
    fprintf(stderr, "time kernel failed: %s\n", cudaGetErrorString(cudaGetLastError()));
    return -3;
  }
  long long* h = new long long[grid];
  cudaMemcpy(h, d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  delete[] h;
  cudaFree(d_ops);
  cudaFree(d_cyc);
  return mx * 1000 / (static_cast<long long>(repeat) * n_ops);
}

// ------------------------------------------------------------------------------------------------ TMA probe
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeTiledFn)p;
  }
  return fn;
}

__global__ void probe_tma_kernel(const __grid_constant__ CUtensorMap tmap, int rank, int c0, int c1, int c2, int c3,
                                 int c4, int box_bytes, uint8_t* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  uint32_t base = smem_u32(smem_raw);
  uint32_t aligned = (base + 1023u) & ~1023u;
  uint8_t* img = smem_raw + (aligned - base);
  for (int i = threadIdx.x; i < box_bytes; i += blockDim.x) img[i] = 0xEE;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, box_bytes);
    if (rank == 5) tma_load_5d(img, &tmap, &bar, c0, c1, c2, c3, c4);
    else if (rank == 4) tma_load_4d(img, &tmap, &bar, c0, c1, c2, c3);
    else if (rank == 3) tma_load_3d(img, &tmap, &bar, c0, c1, c2);
    else tma_load_2d(img, &tmap, &bar, c0, c1);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < box_bytes; i += blockDim.x) out[i] = img[i];
}

// elem_bytes: 2 (bf16) or 4 (fp32). swizzle: 0 none, 1 32B, 2 64B, 3 128B. strides are bytes for dims 1..rank-1.
extern "C" int hpvg_probe_tma(const void* gsrc, long long gbytes, int elem_bytes, int rank, const long long* dims,
                              const long long* strides_bytes, const int* box, int swizzle, const int* coords,
                              void* out, int box_bytes) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -10;
  uint8_t* d_src;
  uint8_t* d_out;
  CK(cudaMalloc(&d_src, gbytes));
  CK(cudaMemcpy(d_src, gsrc, gbytes, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_out, box_bytes));
  CUtensorMap tmap;
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
  CUtensorMapSwizzle sw = swizzle == 3 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_32B
                                         : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(&tmap, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank,
                   d_src, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r);
    return -11;
  }
  int smem = box_bytes + 1024;
  CK(cudaFuncSetAttribute(probe_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_tma_kernel<<<1, 128, smem>>>(tmap, rank, coords[0], coords[1], coords[2], coords[3], coords[4], box_bytes,
                                     d_out);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out, d_out, box_bytes, cudaMemcpyDeviceToHost));
  cudaFree(d_src);
  cudaFree(d_out);
  return 0;
}
