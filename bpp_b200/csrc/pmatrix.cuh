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

// Wide-state variant of the eigen form (20 states): grid (locus, slice), V and V^-1 of the locus staged in
// shared memory once per block; every WARP then takes its own (branch, category) pairs -- expm1 of the S
// eigenvalues, temp = V^-1 . diag(expm1) in a padded (S+1)-stride private buffer (conflict-free row reads),
// and P = I + temp . V in 1 x 4 register tiles -- with warp-level synchronisation only.
// ST = compile-time state count (full unrolling), 0 = read it from the locus.
template <int ST>
__global__ void __launch_bounds__(128)
pmatrix_kernel_wide(const LocusDev * __restrict__ loci, const unsigned int * __restrict__ batch_locus,
                    const unsigned int * __restrict__ mat_off, const unsigned int * __restrict__ mat_idx,
                    const double * __restrict__ mat_bl)
{
  extern __shared__ double s_pm[];        // V[S*S] | Vinv[S*S] | per warp: temp[S*(S+1)] | expm1[S]
  const unsigned int bl = blockIdx.x;
  const LocusDev & L = loci[batch_locus[bl]];
  const unsigned int first = mat_off[bl], count = mat_off[bl + 1] - first;
  const unsigned int S = ST ? (unsigned)ST : L.states, R = L.rate_cats, SS = S * S;
  const unsigned int lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double * sV = s_pm, * sVi = s_pm + SS;
  double * sT = s_pm + 2 * SS + (size_t)warp * (S * (S + 1) + S), * sE = sT + S * (S + 1);
  if (blockIdx.y * nw >= count * R) return;
  for (unsigned int t = threadIdx.x; t < SS; t += blockDim.x) { sV[t] = L.eigenvecs[t]; sVi[t] = L.inv_eigenvecs[t]; }
  __syncthreads();
  const bool quad = (S & 3u) == 0;         // 1 x 4 register tile per thread: one temp load feeds four columns
  for (unsigned int g = blockIdx.y * nw + warp; g < count * R; g += gridDim.y * nw)
  {
    const unsigned int n = g % R, m = g / R;
    const double bt = mat_bl[first + m] * L.rates[n];
    double * P = L.pmat + ((size_t)mat_idx[first + m] * R + n) * SS;
    if (bt < 1e-100)
    {
      for (unsigned int t = lane; t < SS; t += 32) P[t] = (t / S == t % S) ? 1.0 : 0.0;
      continue;
    }
    __syncwarp();                          // the previous pair is done with temp and expm1
    for (unsigned int t = lane; t < S; t += 32) sE[t] = expm1(L.eigenvals[t] * bt);                   // core_pmatrix.c:753-754
    __syncwarp();
    for (unsigned int t = lane; t < SS; t += 32) sT[(t / S) * (S + 1) + t % S] = __dmul_rn(sVi[t], sE[t % S]);   // :756-758
    __syncwarp();
    if (quad)
    {
      // 4 x 4 register tiles: per inner index four temp loads and one 32-byte V load feed 16 multiply-adds
      const unsigned int q4 = S >> 2;
      for (unsigned int t = lane; t < q4 * q4; t += 32)
      {
        const unsigned int j = (t / q4) * 4, k = (t % q4) * 4;
        double a[4][4];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y) a[x][y] = (j + x == k + y) ? 1.0 : 0.0;
#pragma unroll
        for (unsigned int mm = 0; mm < S; ++mm)                                                      // :760-771
        {
          const double2 v0 = *reinterpret_cast<const double2 *>(sV + mm * S + k);
          const double2 v1 = *reinterpret_cast<const double2 *>(sV + mm * S + k + 2);
#pragma unroll
          for (int x = 0; x < 4; ++x)
          {
            const double tv = sT[(j + x) * (S + 1) + mm];
            a[x][0] = __dadd_rn(a[x][0], __dmul_rn(tv, v0.x)); a[x][1] = __dadd_rn(a[x][1], __dmul_rn(tv, v0.y));
            a[x][2] = __dadd_rn(a[x][2], __dmul_rn(tv, v1.x)); a[x][3] = __dadd_rn(a[x][3], __dmul_rn(tv, v1.y));
          }
        }
#pragma unroll
        for (int x = 0; x < 4; ++x) st256(P + (j + x) * S + k, a[x][0], a[x][1], a[x][2], a[x][3]);
      }
    }
    else
      for (unsigned int t = lane; t < SS; t += 32)
      {
        const unsigned int j = t / S, k = t % S;
        double acc = (j == k) ? 1.0 : 0.0;
        for (unsigned int mm = 0; mm < S; ++mm) acc = __dadd_rn(acc, __dmul_rn(sT[j * (S + 1) + mm], sV[mm * S + k]));
        P[t] = acc;
      }
  }
}

}  // namespace bppgpu
