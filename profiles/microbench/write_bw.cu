// write_bw.cu -- what does a store-only stream reach on B200?  (tree_kernel_s4 is ~95 % stores)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_bw write_bw.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k_store(double * __restrict__ p, size_t n4, int iters_per_thread)
{
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const double v = (double)threadIdx.x;
  for (; i < n4; i += stride)
  {
    double * q = p + i * 4;
    if (MODE == 0) asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%1,%1,%1};" :: "l"(q), "d"(v) : "memory");
    if (MODE == 1) asm volatile("st.global.v4.f64 [%0], {%1,%1,%1,%1};" :: "l"(q), "d"(v) : "memory");
    if (MODE == 2) asm volatile("st.global.cs.v4.f64 [%0], {%1,%1,%1,%1};" :: "l"(q), "d"(v) : "memory");
    if (MODE == 3) { asm volatile("st.global.v2.f64 [%0], {%1,%1};" :: "l"(q), "d"(v) : "memory");
                     asm volatile("st.global.v2.f64 [%0], {%1,%1};" :: "l"(q + 2), "d"(v) : "memory"); }
    if (MODE == 4) asm volatile("st.global.wt.v4.f64 [%0], {%1,%1,%1,%1};" :: "l"(q), "d"(v) : "memory");
  }
}

__global__ void k_copy(const double * __restrict__ a, double * __restrict__ b, size_t n4)
{
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride)
  {
    double x, y, z, w;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=d"(w) : "l"(a + i * 4));
    asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(b + i * 4), "d"(x), "d"(y), "d"(z), "d"(w) : "memory");
  }
}

__global__ void k_read(const double * __restrict__ a, double * __restrict__ out, size_t n4)
{
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  double acc = 0;
  for (; i < n4; i += stride)
  {
    double x, y, z, w;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=d"(w) : "l"(a + i * 4));
    acc += x + y + z + w;
  }
  if (acc == 123.456) out[0] = acc;
}

template <class F> float time_ms(F f, int reps)
{
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}

int main()
{
  const size_t bytes = (size_t)4 << 30;       // 4 GiB >> L2
  double * p, * q; cudaMalloc(&p, bytes); cudaMalloc(&q, bytes);
  cudaMemset(p, 0, bytes); cudaMemset(q, 0, bytes);
  const size_t n4 = bytes / 32;
  const int grids[] = {148 * 4, 148 * 8, 148 * 16, 148 * 64};
  for (int g : grids)
  {
    printf("grid %d x 256\n", g);
    printf("  st.L1::no_allocate.v4.f64 : %7.1f GB/s\n", bytes / 1e6 / time_ms([&] { k_store<0><<<g, 256>>>(p, n4, 0); }, 10));
    printf("  st.v4.f64                 : %7.1f GB/s\n", bytes / 1e6 / time_ms([&] { k_store<1><<<g, 256>>>(p, n4, 0); }, 10));
    printf("  st.cs.v4.f64              : %7.1f GB/s\n", bytes / 1e6 / time_ms([&] { k_store<2><<<g, 256>>>(p, n4, 0); }, 10));
    printf("  2 x st.v2.f64             : %7.1f GB/s\n", bytes / 1e6 / time_ms([&] { k_store<3><<<g, 256>>>(p, n4, 0); }, 10));
    printf("  st.wt.v4.f64              : %7.1f GB/s\n", bytes / 1e6 / time_ms([&] { k_store<4><<<g, 256>>>(p, n4, 0); }, 10));
    printf("  copy (read+write bytes)   : %7.1f GB/s\n", 2.0 * bytes / 1e6 / time_ms([&] { k_copy<<<g, 256>>>(p, q, n4); }, 10));
    printf("  read only                 : %7.1f GB/s\n", bytes / 1e6 / time_ms([&] { k_read<<<g, 256>>>(p, q, n4); }, 10));
  }
  printf("cudaMemsetAsync             : %7.1f GB/s\n", bytes / 1e6 / time_ms([&] { cudaMemsetAsync(p, 1, bytes); }, 10));
  printf("cudaMemcpyAsync D2D (r+w)   : %7.1f GB/s\n", 2.0 * bytes / 1e6 / time_ms([&] { cudaMemcpyAsync(q, p, bytes, cudaMemcpyDeviceToDevice); }, 10));
  return 0;
}
