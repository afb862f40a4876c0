// store_paths.cu -- how should a compute-then-store kernel at LOW occupancy (tree_kernel_s4<4,.,4>: 8 warps per SM,
// 4 cells of 32 bytes per thread and op) hand its results to HBM?
//   mode 0: st.global.v4.f64 straight from registers (what the kernel does today)
//   mode 1: registers -> shared memory (STS.128 x2 per cell) -> one cp.async.bulk.global.shared::cta of 4 KB per warp
//           and op (TMA bulk store, double buffered per warp): stores leave the register file and the LSU queue
//   mode 2: like 1 but one 32 KB bulk store per CTA and op (one elected thread, __syncthreads)
// Each warp alternates K dependent FP64 multiply-adds per cell ("the op") with the store of its 128 cells.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_paths store_paths.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void st256(double * p, double a, double b, double c, double d)
{
  asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void bulk_store(void * gdst, const void * ssrc, unsigned bytes)
{
  const unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int MODE, int CPT, int K>
__global__ void __launch_bounds__(256) k(double * __restrict__ out, size_t cells_per_cta, int ops, double seed)
{
  extern __shared__ double4 sm[];
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  constexpr unsigned TILE = 256 * CPT;                     // cells per CTA and op
  double * base = out + (size_t)blockIdx.x * cells_per_cta * 4;
  const size_t tiles = cells_per_cta / TILE;
  double x[CPT][4];
#pragma unroll
  for (int j = 0; j < CPT; ++j) { x[j][0] = seed + tid; x[j][1] = seed * 2 + j; x[j][2] = seed * 3; x[j][3] = seed * 5 + lane; }
  unsigned buf = 0;
  for (size_t t = 0; t < tiles; ++t)
  {
    for (int op = 0; op < ops; ++op)
    {
      // the "op": K dependent multiply-adds on each of the 4 components of every cell
#pragma unroll
      for (int k2 = 0; k2 < K; ++k2)
#pragma unroll
        for (int j = 0; j < CPT; ++j)
        {
          x[j][0] = fma(x[j][0], 1.0000001, x[j][1] * 1e-9); x[j][1] = fma(x[j][1], 0.9999999, x[j][2] * 1e-9);
          x[j][2] = fma(x[j][2], 1.0000002, x[j][3] * 1e-9); x[j][3] = fma(x[j][3], 0.9999998, x[j][0] * 1e-9);
        }
      // destination of this (tile, op): ops buffers of cells_per_cta cells each would be the real layout; here the
      // ops of a tile simply follow each other
      double * dst = base + ((t * ops + op) % tiles) * (size_t)TILE * 4;
      if (MODE == 0)
      {
#pragma unroll
        for (int j = 0; j < CPT; ++j) st256(dst + ((size_t)j * 256 + tid) * 4, x[j][0], x[j][1], x[j][2], x[j][3]);
      }
      else if (MODE == 1)
      {
        // warp-private staging: [2 buffers][8 warps][CPT*32 cells]; the warp's cells are contiguous in HBM
        double4 * st = sm + ((size_t)buf * 8 + warp) * (CPT * 32);
        bulk_wait_read<1>();                               // the buffer used two ops ago has been read by the TMA unit
        __syncwarp();
#pragma unroll
        for (int j = 0; j < CPT; ++j) st[j * 32 + lane] = make_double4(x[j][0], x[j][1], x[j][2], x[j][3]);
        fence_async();
        __syncwarp();
        if (lane == 0) { bulk_store(dst + (size_t)warp * (CPT * 32) * 4, st, CPT * 32 * 32); bulk_commit(); }
        buf ^= 1u;
      }
      else
      {
        double4 * st = sm + (size_t)buf * TILE;
        if (tid == 0) bulk_wait_read<1>();
        __syncthreads();
#pragma unroll
        for (int j = 0; j < CPT; ++j) st[j * 256 + tid] = make_double4(x[j][0], x[j][1], x[j][2], x[j][3]);
        fence_async();
        __syncthreads();
        if (tid == 0) { bulk_store(dst, st, TILE * 32); bulk_commit(); }
        buf ^= 1u;
      }
    }
  }
  if (MODE != 0) { bulk_wait_read<0>(); asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
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

template <int MODE, int CPT, int K>
void run(double * p, size_t bytes, int ctas_per_sm, const char * what)
{
  const int grid = 148 * ctas_per_sm;
  const size_t tile_bytes = (size_t)256 * CPT * 32;
  size_t cells_per_cta = bytes / grid / 32;
  cells_per_cta -= cells_per_cta % (256 * CPT);
  const size_t smem = MODE == 0 ? 0 : 2 * tile_bytes;
  cudaFuncSetAttribute(k<MODE, CPT, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int ops = 15;
  const double total = (double)grid * cells_per_cta * 32.0 * ops;
  // a tile's ops rewrite the CTA's region `ops` times over; total bytes written = ops x region
  float ms = time_ms([&] { k<MODE, CPT, K><<<grid, 256, smem>>>(p, cells_per_cta, ops, 1.0); }, 3);
  cudaError_t err = cudaGetLastError();
  printf("  %-58s K=%3d CPT=%d CTAs/SM=%d : %7.1f GB/s%s\n", what, K, CPT, ctas_per_sm, total / 1e6 / ms,
         err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main()
{
  const size_t bytes = (size_t)2 << 30;
  double * p; cudaMalloc(&p, bytes); cudaMemset(p, 0, bytes);
  printf("compute-then-store at low occupancy, 256 threads per CTA, 15 ops per tile, %zu MiB region\n", bytes >> 20);
#define ROW(K) \
  run<0, 4, K>(p, bytes, 1, "st.v4.f64 from registers"); \
  run<0, 2, K>(p, bytes, 2, "st.v4.f64 from registers"); \
  run<0, 1, K>(p, bytes, 4, "st.v4.f64 from registers"); \
  run<1, 4, K>(p, bytes, 1, "smem + 4 KB bulk store per warp"); \
  run<1, 2, K>(p, bytes, 2, "smem + 2 KB bulk store per warp"); \
  run<2, 4, K>(p, bytes, 1, "smem + 32 KB bulk store per CTA");
  ROW(0) ROW(8) ROW(16) ROW(32)
  return 0;
}
