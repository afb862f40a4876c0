// pmatrix.cuh -- transition-probability matrices (locus.c:2325-2415, core_pmatrix.c:674-783)
#pragma once
#include "common.cuh"

namespace bppgpu {

// ----------------------------------------------------------------------------- P-matrix kernel
// grid.x = loci of the batch; the threads of a block stride over (op, cat, row) of their locus.
// JC69: locus.c:2390-2391 (exp form).  Eigen: core_pmatrix.c:745-771 -- expm1, temp = Vinv*expd,
// P[j][k] = delta_jk + sum_m temp[j][m]*V[m][k], m-sum sequential with separate mul/add.
// one (branch, category, row) task: row j of P for category n of the branch with length t
__device__ __forceinline__ void pmatrix_row(const LocusDev & L, unsigned int pm_idx, double t, unsigned int n, unsigned int j)
{
  const unsigned int S = L.states, R = L.rate_cats;
  const double bt = t * L.rates[n];
  double * row = L.pmat + ((size_t)pm_idx * R + n) * S * S + (size_t)j * S;
  if (bt < 1e-100)
  {
    for (unsigned int k = 0; k < S; ++k) row[k] = (j == k) ? 1.0 : 0.0;
  }
  else if (L.model_kind == 0)
  {
    const double a = (1 + 3 * exp(-4 * bt / 3)) / 4;
    const double b = (1 - a) / 3;
    for (unsigned int k = 0; k < S; ++k) row[k] = (j == k) ? a : b;
  }
  else if (S == 4)
  {
    const double * __restrict__ V = L.eigenvecs;
    const double * __restrict__ Vi = L.inv_eigenvecs + (size_t)j * 4;
    double temp[4];
#pragma unroll
    for (int mm = 0; mm < 4; ++mm) temp[mm] = __dmul_rn(Vi[mm], expm1(L.eigenvals[mm] * bt));
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
      double acc = (j == (unsigned)k) ? 1.0 : 0.0;
#pragma unroll
      for (int mm = 0; mm < 4; ++mm) acc = __dadd_rn(acc, __dmul_rn(temp[mm], V[mm * 4 + k]));
      row[k] = acc;
    }
  }
  else
  {
    const double * __restrict__ V = L.eigenvecs;
    const double * __restrict__ Vi = L.inv_eigenvecs + (size_t)j * S;
    const double * __restrict__ ev = L.eigenvals;
    for (unsigned int k = 0; k < S; ++k)
    {
      double acc = (j == k) ? 1.0 : 0.0;
      for (unsigned int mm = 0; mm < S; ++mm)
      {
        const double temp = __dmul_rn(Vi[mm], expm1(ev[mm] * bt));
        acc = __dadd_rn(acc, __dmul_rn(temp, V[(size_t)mm * S + k]));
      }
      row[k] = acc;
    }
  }
}

// whole 4x4 matrix of one (branch, category): the expm1 / exp values are shared by the four rows
__device__ __forceinline__ void pmatrix_full4(const LocusDev & L, unsigned int pm_idx, double t, unsigned int n)
{
  const unsigned int R = L.rate_cats;
  const double bt = t * L.rates[n];
  double2 * P = reinterpret_cast<double2 *>(L.pmat + ((size_t)pm_idx * R + n) * 16);
  if (bt < 1e-100)
  {
#pragma unroll
    for (int j = 0; j < 4; ++j) { P[2 * j] = make_double2(j == 0 ? 1.0 : 0.0, j == 1 ? 1.0 : 0.0); P[2 * j + 1] = make_double2(j == 2 ? 1.0 : 0.0, j == 3 ? 1.0 : 0.0); }
  }
  else if (L.model_kind == 0)
  {
    const double a = (1 + 3 * exp(-4 * bt / 3)) / 4;
    const double b = (1 - a) / 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) { P[2 * j] = make_double2(j == 0 ? a : b, j == 1 ? a : b); P[2 * j + 1] = make_double2(j == 2 ? a : b, j == 3 ? a : b); }
  }
  else
  {
    const double * __restrict__ V = L.eigenvecs;
    const double * __restrict__ Vi = L.inv_eigenvecs;
    double ex[4], v[16];
#pragma unroll
    for (int mm = 0; mm < 4; ++mm) ex[mm] = expm1(L.eigenvals[mm] * bt);
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = V[e];
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
      double temp[4], row[4];
#pragma unroll
      for (int mm = 0; mm < 4; ++mm) temp[mm] = __dmul_rn(Vi[j * 4 + mm], ex[mm]);
#pragma unroll
      for (int k = 0; k < 4; ++k)
      {
        double acc = (j == k) ? 1.0 : 0.0;
#pragma unroll
        for (int mm = 0; mm < 4; ++mm) acc = __dadd_rn(acc, __dmul_rn(temp[mm], v[mm * 4 + k]));
        row[k] = acc;
      }
      P[2 * j] = make_double2(row[0], row[1]); P[2 * j + 1] = make_double2(row[2], row[3]);
    }
  }
}

__global__ void __launch_bounds__(128)
pmatrix_kernel(const LocusDev * __restrict__ loci, const unsigned int * __restrict__ batch_locus,
               const unsigned int * __restrict__ mat_off, const unsigned int * __restrict__ mat_idx,
               const double * __restrict__ mat_bl)
{
  const unsigned int bl = blockIdx.x;
  const LocusDev & L = loci[batch_locus[bl]];
  const unsigned int first = mat_off[bl], count = mat_off[bl + 1] - first;
  const unsigned int S = L.states, R = L.rate_cats;
  const unsigned int tasks = count * R * S;
  for (unsigned int t = threadIdx.x; t < tasks; t += blockDim.x)
  {
    const unsigned int j = t % S, n = (t / S) % R, m = t / (S * R);
    pmatrix_row(L, mat_idx[first + m], mat_bl[first + m], n, j);
  }
}

// 20-state variant of the eigen form: grid (locus, slice); a block handles every gridDim.y-th (branch, category)
// pair of its locus with V and V^-1 staged in shared memory and temp = V^-1 . diag(expm1) formed once per pair
__global__ void __launch_bounds__(128)
pmatrix_kernel_wide(const LocusDev * __restrict__ loci, const unsigned int * __restrict__ batch_locus,
                    const unsigned int * __restrict__ mat_off, const unsigned int * __restrict__ mat_idx,
                    const double * __restrict__ mat_bl)
{
  extern __shared__ double s_pm[];        // V[S*S] | Vinv[S*S] | temp[S*S]
  const unsigned int bl = blockIdx.x;
  const LocusDev & L = loci[batch_locus[bl]];
  const unsigned int first = mat_off[bl], count = mat_off[bl + 1] - first;
  const unsigned int S = L.states, R = L.rate_cats, SS = S * S;
  double * sV = s_pm, * sVi = s_pm + SS, * sT = s_pm + 2 * SS;
  if (blockIdx.y >= count * R) return;
  for (unsigned int t = threadIdx.x; t < SS; t += blockDim.x) { sV[t] = L.eigenvecs[t]; sVi[t] = L.inv_eigenvecs[t]; }
  __syncthreads();
  for (unsigned int g = blockIdx.y; g < count * R; g += gridDim.y)
  {
    const unsigned int n = g % R, m = g / R;
    const double bt = mat_bl[first + m] * L.rates[n];
    double * P = L.pmat + ((size_t)mat_idx[first + m] * R + n) * SS;
    if (bt < 1e-100)
    {
      for (unsigned int t = threadIdx.x; t < SS; t += blockDim.x) P[t] = (t / S == t % S) ? 1.0 : 0.0;
      continue;
    }
    __syncthreads();                       // the previous pair is done with temp
    for (unsigned int t = threadIdx.x; t < SS; t += blockDim.x)
      sT[t] = __dmul_rn(sVi[t], expm1(L.eigenvals[t % S] * bt));          // core_pmatrix.c:753-758
    __syncthreads();
    for (unsigned int t = threadIdx.x; t < SS; t += blockDim.x)
    {
      const unsigned int j = t / S, k = t % S;
      double acc = (j == k) ? 1.0 : 0.0;
      for (unsigned int mm = 0; mm < S; ++mm) acc = __dadd_rn(acc, __dmul_rn(sT[j * S + mm], sV[mm * S + k]));   // :760-771
      P[t] = acc;
    }
  }
}

}  // namespace bppgpu
