// fp64_peak.cu -- DFMA vs DMMA (mma.sync.m8n8k4.f64) peak on B200, to decide the 20-state design.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double * out, int iters)
{
  double a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
  double s = 0;
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 12345.678) out[0] = s;
}

__global__ void k_dmma(double * out, int iters)
{
  double c[8][2];
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

int main()
{
  double * out; cudaMalloc(&out, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, blocks = 148 * 4, threads = 256;
  float ms;
  k_dfma<<<blocks, threads>>>(out, 10); k_dmma<<<blocks, threads>>>(out, 10); cudaDeviceSynchronize();
  cudaEventRecord(e0); k_dfma<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  printf("DFMA: %.2f TFLOP/s (%.3f ms)\n", 2.0 * blocks * threads * 16.0 * iters / ms / 1e9, ms);
  cudaEventRecord(e0); k_dmma<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  // per warp-level mma: 8x8x4 = 256 FMA = 512 flop
  printf("DMMA m8n8k4: %.2f TFLOP/s (%.3f ms)\n", 512.0 * blocks * (threads / 32) * 8.0 * iters / ms / 1e9, ms);
  return 0;
}
