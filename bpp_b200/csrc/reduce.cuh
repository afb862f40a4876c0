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

// ----------------------------------------------------------------------------- diploid loci inside a batch
// A batched root evaluation treats every locus as haploid (site log-likelihoods).  For the loci of the batch that
// carry a diploid mapping (locus.c:2586-2615) this kernel, launched between the tree kernel and finish_kernel,
// redoes the root from the root CLV that the tree kernel has just written: site likelihoods
// (pll_core_root_likelihood_vector, core_likelihood.c:214-408: no log, no scaler, no weight), mean over the phase
// resolutions, log, weight of the unphased site, fixed-order sum.  The result replaces the locus' tile partials
// (first tile = the value, the others = 0), so finish_kernel needs no change.  One block per batch locus; blocks
// of haploid loci return at once.
__global__ void __launch_bounds__(256)
diploid_batch_kernel(const LocusDev * __restrict__ loci, const unsigned int * __restrict__ batch_locus,
                     const unsigned int * __restrict__ root_clv, const unsigned int * __restrict__ tile_first,
                     double * __restrict__ tile_partial)
{
  __shared__ double s_red[8];
  const unsigned int b = blockIdx.x;
  const LocusDev & L = loci[batch_locus[b]];
  if (!L.dip_off) return;
  const unsigned int S = L.states, R = L.rate_cats;
  const double * clv = L.clv + (size_t)(root_clv[b] - L.tips) * L.clv_stride;
  double acc = 0.0;
  for (unsigned int i = threadIdx.x; i < L.unphased; i += blockDim.x)
  {
    const unsigned long long a0 = L.dip_off[i], a1 = L.dip_off[i + 1];
    double mean = 0.0;
    for (unsigned long long k = a0; k < a1; ++k)
    {
      const double * c = clv + (size_t)L.dip_map[k] * L.site_stride;
      double term = 0.0;
      for (unsigned int r = 0; r < R; ++r, c += L.cat_stride)
      {
        double tr;
        if (S == 4)      // core_likelihood_avx.c:121-150: (p0 + p1) + (p2 + p3)
          tr = __dadd_rn(__dadd_rn(__dmul_rn(L.freqs[0], c[0]), __dmul_rn(L.freqs[1], c[1])),
                         __dadd_rn(__dmul_rn(L.freqs[2], c[2]), __dmul_rn(L.freqs[3], c[3])));
        else
        {
          tr = 0.0;
          for (unsigned int s = 0; s < S; ++s) tr = __dadd_rn(tr, __dmul_rn(c[s], L.freqs[s]));
        }
        term = __dadd_rn(term, __dmul_rn(tr, L.rate_weights[r]));
      }
      mean = __dadd_rn(mean, term);
    }
    mean = mean / (double)(a1 - a0);
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
    tile_partial[tile_first[b]] = t;
  }
  for (unsigned int t = tile_first[b] + 1 + threadIdx.x; t < tile_first[b + 1]; t += blockDim.x) tile_partial[t] = 0.0;
}

}  // namespace bppgpu
