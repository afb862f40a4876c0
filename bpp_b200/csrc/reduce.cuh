// reduce.cuh -- fixed-order reductions: per-locus lnL, batch sum, diploid phase mean
#pragma once
#include "common.cuh"

namespace bppgpu {

// ----------------------------------------------------------------------------- finish kernel
// lnl[locus] = sum of its tile partials in tile order; lnl_sum = fixed-order sum over the loci.
__global__ void __launch_bounds__(1024)
finish_kernel(const double * __restrict__ tile_partial, const unsigned int * __restrict__ tile_first,
              unsigned int n_loci, double * __restrict__ lnl, double * __restrict__ lnl_sum)
{
  __shared__ double s_red[32];
  double acc = 0.0;
  for (unsigned int i = threadIdx.x; i < n_loci; i += blockDim.x)
  {
    double v = 0.0;
    for (unsigned int t = tile_first[i]; t < tile_first[i + 1]; ++t) v += tile_partial[t];
    lnl[i] = v;
    acc += v;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
  if ((threadIdx.x & 31u) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    double t = 0.0;
    for (unsigned int w = 0; w < (blockDim.x >> 5); ++w) t += s_red[w];
    *lnl_sum = t;
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
    acc += __dmul_rn(log(mean), (double)L.weights[i]);
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
