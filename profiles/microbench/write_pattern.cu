// write_pattern.cu -- is tree_kernel_s4's HBM WRITE PATTERN itself slower than a linear stream?
// The kernel writes, per CTA, tile after tile (1024 cells = 32 KB) and within a tile op after op (15 ops = 15
// different CLV buffers of the locus, 128 000 B apart): 148 CTAs x 15 interleaved 32 KB streams.  This program
// issues exactly those stores (no arithmetic) and, for comparison, the same bytes as one linear stream per CTA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_pattern write_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void st256(double * p, double a)
{
  asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%1,%1,%1};" :: "l"(p), "d"(a) : "memory");
}

// mode 0: kernel pattern; mode 1: linear per CTA; mode 2: kernel pattern but ops outermost per locus (op-major)
template <int MODE, int CPT>
__global__ void __launch_bounds__(256) k(double * __restrict__ out, int n_loci, int ops, int cells, int buffers)
{
  const unsigned tid = threadIdx.x;
  const size_t buf_doubles = (size_t)cells * 4, locus_doubles = buf_doubles * buffers;
  const int tiles = (cells + 256 * CPT - 1) / (256 * CPT);
  const int l0 = (int)(((long long)n_loci * blockIdx.x) / gridDim.x), l1 = (int)(((long long)n_loci * (blockIdx.x + 1)) / gridDim.x);
  const double v = (double)tid;
  if (MODE == 1)
  {
    double * p = out + (size_t)l0 * locus_doubles;
    const size_t total_cells = (size_t)(l1 - l0) * ops * cells;
    for (size_t c = tid; c < total_cells; c += 256) st256(p + c * 4, v);
    return;
  }
  for (int l = l0; l < l1; ++l)
  {
    double * base = out + (size_t)l * locus_doubles;
    if (MODE == 0)
      for (int t = 0; t < tiles; ++t)
        for (int op = 0; op < ops; ++op)
#pragma unroll
          for (int j = 0; j < CPT; ++j)
          {
            const int cell = t * 256 * CPT + j * 256 + tid;
            if (cell < cells) st256(base + (size_t)op * buf_doubles + (size_t)cell * 4, v);
          }
    else
      for (int op = 0; op < ops; ++op)
        for (int t = 0; t < tiles; ++t)
#pragma unroll
          for (int j = 0; j < CPT; ++j)
          {
            const int cell = t * 256 * CPT + j * 256 + tid;
            if (cell < cells) st256(base + (size_t)op * buf_doubles + (size_t)cell * 4, v);
          }
  }
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
  const int n_loci = 10000, ops = 15, cells = 4000, buffers = 30;        // config 3
  const size_t bytes = (size_t)n_loci * buffers * cells * 32;
  double * p; if (cudaMalloc(&p, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(p, 0, bytes);
  const double written = (double)n_loci * ops * cells * 32;
  printf("config-3 write pattern: %d loci x %d ops x %d cells x 32 B = %.2f GB per pass (arena %.1f GB)\n", n_loci, ops, cells, written / 1e9, bytes / 1e9);
  for (int ctas : {1, 2, 4})
  {
    const int g = 148 * ctas;
    printf(" grid %d x 256:\n", g);
    printf("   kernel pattern (tile-major, 15 buffers), 4 cells/thread : %7.1f GB/s\n", written / 1e6 / time_ms([&] { k<0, 4><<<g, 256>>>(p, n_loci, ops, cells, buffers); }, 5));
    printf("   kernel pattern, 2 cells/thread                          : %7.1f GB/s\n", written / 1e6 / time_ms([&] { k<0, 2><<<g, 256>>>(p, n_loci, ops, cells, buffers); }, 5));
    printf("   op-major within a locus (buffer after buffer)           : %7.1f GB/s\n", written / 1e6 / time_ms([&] { k<2, 4><<<g, 256>>>(p, n_loci, ops, cells, buffers); }, 5));
    printf("   linear stream per CTA                                   : %7.1f GB/s\n", written / 1e6 / time_ms([&] { k<1, 4><<<g, 256>>>(p, n_loci, ops, cells, buffers); }, 5));
  }
  return 0;
}
