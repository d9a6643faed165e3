// bf_microbench.cu -- roofline denominators the fold kernels are judged against.
//
// MEASURED_PEAKS.json (driver-written) only has HBM copy bandwidth and bf16 GEMM throughput;
// the fold kernels are bound by the INT32 ALU pipe (min-plus), the FP64 pipe (sum-product)
// and shared-memory bandwidth, so those three peaks are measured here on the same GPU.
#include <cuda_runtime.h>

#include "../../include/b200fold.h"

namespace {

// min-plus relaxations on registers: 1 IADD + 1 IMNMX per relaxation, 8 independent chains
__global__ void mb_int32(int *out, int iters, int seed) {
  int a0 = threadIdx.x + seed, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  int b = blockIdx.x + 3, c = seed | 1;
  for (int k = 0; k < iters; k++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      a0 = min(a0, a1 + b); a1 = min(a1, a2 + c); a2 = min(a2, a3 + b); a3 = min(a3, a4 + c);
      a4 = min(a4, a5 + b); a5 = min(a5, a6 + c); a6 = min(a6, a7 + b); a7 = min(a7, a0 + c);
    }
    b += k; c ^= k;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// sum-product on registers: 1 DFMA per relaxation (2 flops), 8 independent chains
__global__ void mb_fp64(double *out, int iters, double seed) {
  double a0 = threadIdx.x * 1e-9 + seed, a1 = a0 + 1e-9, a2 = a0 + 2e-9, a3 = a0 + 3e-9, a4 = a0 + 4e-9, a5 = a0 + 5e-9, a6 = a0 + 6e-9, a7 = a0 + 7e-9;
  const double b = 0.999999, c = 1e-12;
  for (int k = 0; k < iters; k++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
      a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// conflict-free 128-bit shared-memory loads
__global__ void mb_smem(int *out, int iters) {
  __shared__ int4 buf[2048];  // 32 KB
  for (int k = threadIdx.x; k < 2048; k += blockDim.x) buf[k] = make_int4(k, k + 1, k + 2, k + 3);
  __syncthreads();
  int4 acc = make_int4(0, 0, 0, 0);
  int idx = threadIdx.x;
  for (int k = 0; k < iters; k++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      int4 v = buf[(idx + r * 256) & 2047];
      acc.x += v.x; acc.y ^= v.y; acc.z += v.z; acc.w ^= v.w;
    }
    idx = (idx + 64) & 2047;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

}  // namespace

extern "C" int bf_microbench(double out[3]) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return BF_ERR_CUDA;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int threads = 256, blocks = sms * 8;
  void *buf = nullptr;
  if (cudaMalloc(&buf, (size_t)threads * blocks * sizeof(double)) != cudaSuccess) return BF_ERR_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  const int iters = 4096;
  for (int kind = 0; kind < 3; kind++) {
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
      cudaEventRecord(e0);
      if (kind == 0) mb_int32<<<blocks, threads>>>((int *)buf, iters, rep);
      else if (kind == 1) mb_fp64<<<blocks, threads>>>((double *)buf, iters, 1.0 + rep);
      else mb_smem<<<blocks, threads>>>((int *)buf, iters);
      cudaEventRecord(e1);
      if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(buf); return BF_ERR_CUDA; }
      cudaEventElapsedTime(&ms, e0, e1);
      double per_thread = (double)threads * blocks * iters;
      double rate = (kind == 2) ? per_thread * 8.0 * 16.0 / (ms * 1e-3)   // 8 x 16-byte loads per iteration -> bytes/s
                                : per_thread * 64.0 * 2.0 / (ms * 1e-3);  // 64 relaxations per iteration x (add+min | mul+add) -> ops/s
      if (rate > best) best = rate;
    }
    out[kind] = best;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf);
  return BF_OK;
}
