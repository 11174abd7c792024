// DMMA (mma.sync.m8n8k4.f64) throughput probe against the plain DFMA pipe on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int CH>
__global__ void k_dmma(double* out, int iters, double a, double b) {
  double c[CH][2];
  for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma(c[i][0], c[i][1], a, b);
  double s = 0;
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double c[CH];
  for (int i = 0; i < CH; ++i) c[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i] = fma(a, c[i], b);
  double s = 0;
  for (int i = 0; i < CH; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out;
  cudaMalloc(&out, 148 * 1024 * 8 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int threads : {128, 256, 512, 1024}) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      k_dmma<16><<<148, threads>>>(out, iters, 1.0000001, 0.5);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma_dmma = 148.0 * (threads / 32) * iters * 16.0 * 256.0 / (ms * 1e-3);
      cudaEventRecord(e0);
      k_dfma<16><<<148, threads>>>(out, iters, 1.0000001, 0.5);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      double fma_dfma = 148.0 * threads * iters * 16.0 / (ms * 1e-3);
      if (rep) printf("threads/CTA %4d: DMMA %.3e FMA/s   DFMA %.3e FMA/s   (err %s)\n", threads, fma_dmma, fma_dfma, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
