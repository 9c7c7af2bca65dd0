// Diagnostics only (not part of the product): raw FP64 pipe throughput of the device, used to put
// the DGEMM numbers in profiles/ next to what the hardware can issue.
//   mode 0: DMMA.8x8x4 (mma.sync m8n8k4 f64), 16 independent accumulator pairs per warp
//   mode 1: DFMA, 16 independent chains per thread
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void dmma_probe(double* out, int iters) {
  double c[16][2];
#pragma unroll
  for (int j = 0; j < 16; ++j) c[j][0] = c[j][1] = 0.0;
  double a = 1e-3 * (threadIdx.x & 7), b = 1.0 + 1e-3 * (threadIdx.x & 3);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += c[j][0] + c[j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_probe(double* out, int iters) {
  double c[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) c[j] = 1e-3 * j;
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) c[j] = fma(c[j], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += c[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mode 2: the operand pattern of gemm_f64_kernel's inner k-step (4 A fragments x 8 B fragments, 32
// accumulator pairs) with no shared-memory traffic: isolates register-file effects from LDS/barriers
__global__ void dmma_tile_probe(double* out, int iters) {
  double c[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) c[i][j][0] = c[i][j][1] = 0.0;
  double a[4], b[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = 1e-3 * (threadIdx.x & 7) + i;
#pragma unroll
  for (int j = 0; j < 8; ++j) b[j] = 1.0 + 1e-3 * (threadIdx.x & 3) + j;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(c[i][j][0]), "+d"(c[i][j][1]) : "d"(a[i]), "d"(b[j]));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += c[i][j][0] + c[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int probe_fp64(int mode, int threads, int blocks_per_sm, int iters, double* tflops) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = sms * blocks_per_sm;
  double* out;
  if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    if (mode == 0) dmma_probe<<<blocks, threads>>>(out, iters);
    else if (mode == 2) dmma_tile_probe<<<blocks, threads>>>(out, iters);
    else dfma_probe<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  double flops;
  if (mode == 0) flops = (double)blocks * (threads / 32) * iters * 16.0 * 512.0;
  else if (mode == 2) flops = (double)blocks * (threads / 32) * iters * 32.0 * 512.0;
  else flops = (double)blocks * threads * iters * 16.0 * 2.0;
  *tflops = flops / (ms * 1e-3) / 1e12;
  cudaFree(out);
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}
