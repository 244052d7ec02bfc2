// fg_peak.cu -- diagnostic: measured fp64 throughput of the device, DFMA (CUDA cores) and DMMA (mma.sync.m8n8k4.f64),
// all SMs busy, register operands only.  bench.py reports its fp64 rooflines against these numbers, measured in the
// same run (MEASURED_PEAKS.json holds HBM and bf16 only).
#include <cuda_runtime.h>
#include "../../include/fg_abi.h"

namespace {
__global__ void k_peak_dfma(double* out, int iters) {
  double a[8], b = 1.0000001, c = 0.9999999;
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_peak_dmma(double* out, int iters) {
  double c[4][2];
  for (int i = 0; i < 4; ++i) { c[i][0] = threadIdx.x; c[i][1] = 1.0; }
  const double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

extern "C" int fg_debug_fp64_peak(int device, double out[2]) {
  if (!out) return FG_ERR_INVALID;
  cudaDeviceProp p;
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&p, device) != cudaSuccess) return FG_ERR_CUDA;
  const int blocks = p.multiProcessorCount * 4, threads = 256, iters = 20000;
  double* buf = nullptr;
  if (cudaMalloc(&buf, sizeof(double) * blocks * threads) != cudaSuccess) return FG_ERR_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 2; ++rep) {          // the second pass is the measurement
    cudaEventRecord(e0); k_peak_dfma<<<blocks, threads>>>(buf, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    out[0] = 2.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    cudaEventRecord(e0); k_peak_dmma<<<blocks, threads>>>(buf, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    out[1] = 512.0 * 4 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12;   // m8n8k4: 512 flop per warp instruction
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  const cudaError_t e = cudaGetLastError();
  cudaFree(buf);
  return e == cudaSuccess ? FG_OK : FG_ERR_CUDA;
}
