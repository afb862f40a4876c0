// reduce.cuh -- fixed-order reductions: per-locus lnL, batch sum, diploid phase mean
#pragma once
#include "common.cuh"

namespace bppgpu {

// ----------------------------------------------------------------------------- finish kernel
// lnl[locus] = sum of its tile partials in tile order; lnl_sum = sum over the loci.
// One thread per locus; each block reduces its 256 loci in a fixed tree, the LAST block to finish adds
// the block sums in block order.  Every order is a function of indices only (never of arrival), so the
// result is bitwise reproducible; no floating-point atomics.
__global__ void __launch_bounds__(256)
finish_kernel(const double * __restrict__ tile_partial, const unsigned int * __restrict__ tile_first,
              unsigned int n_loci, double * __restrict__ lnl, double * __restrict__ lnl_sum,
              double * __restrict__ block_sums, unsigned int * __restrict__ counter)
{
  __shared__ double s_red[8];
  __shared__ bool s_last;
  const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (i < n_loci)
  {
    for (unsigned int t = tile_first[i]; t < tile_first[i + 1]; ++t) v += tile_partial[t];
    lnl[i] = v;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
  if ((threadIdx.x & 31u) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    double t = 0.0;
    for (unsigned int w = 0; w < (blockDim.x >> 5); ++w) t += s_red[w];
    block_sums[blockIdx.x] = t;
    __threadfence();
    s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double acc = 0.0;
  for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) acc += block_sums[b];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
  __syncthreads();
  if ((threadIdx.x & 31u) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    double t = 0.0;
    for (unsigned int w = 0; w < (blockDim.x >> 5); ++w) t += s_red[w];
    *lnl_sum = t;
    *counter = 0;
  }
}

// ----------------------------------------------------------------------------- diploid kernel
// locus.c:2600-2614: logl = sum_i log(mean_j lh[map[k++]]) * weight[i], one block, fixed order.
__global__ void __launch_bounds__(256)
diploid_kernel(const LocusDev * __restrict__ loci, unsigned int locus_id, const double * __restrict__ lh,
               double * __restrict__ out)
{
  __shared__ double s_red[8];
  const LocusDev & L = loci[locus_id];
  double acc = 0.0;
  for (unsigned int i = threadIdx.x; i < L.unphased; i += blockDim.x)
  {
    const unsigned long long a = L.dip_off[i], b = L.dip_off[i + 1];
    double mean = 0.0;
    for (unsigned long long k = a; k < b; ++k) mean = __dadd_rn(mean, lh[L.dip_map[k]]);
    mean = mean / (double)(b - a);
    acc += __dmul_rn(log(mean), (double)L.dip_weights[i]);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
  if ((threadIdx.x & 31u) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    double t = 0.0;
    for (unsigned int w = 0; w < (blockDim.x >> 5); ++w) t += s_red[w];
    *out = t;
  }
}

}  // namespace bppgpu
