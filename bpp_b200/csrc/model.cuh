// model.cuh -- batched update of the loci's model blocks and the eigen-decomposition on the device.
//
// The callers that change models do so for many loci between two likelihood evaluations (propose_alpha
// prop_gamma.c:53-165: new category rates; propose_qrates / propose_freqs locus.c:2782-3354: new Q), so the
// dirty model blocks of a batch travel in ONE pinned blob and are scattered by one kernel instead of one
// small copy per locus, and pll_update_eigen (core_pmatrix.c:239-297 with create_ratematrix :186-237) runs
// here, one thread per locus, instead of on the host:
//   symmetrised rate matrix A_ij = r_ij sqrt(pi_i pi_j), A_ii = -sum_j r_ij pi_j, divided by the mean rate;
//   eigenvecs[i][j] = a[i][j] sqrt(pi_j), inv_eigenvecs[i][j] = a[j][i] / sqrt(pi_i), eigenvals = d.
// The reference diagonalises with tred2 / tqli; a symmetric matrix has one spectrum, so a cyclic Jacobi
// iteration (the same routine as the host-side jacobi_eigen in engine.cu) gives the same P-matrices to
// rounding (tests/test_gpu_parity.py compares them with the reference's).
#pragma once
#include "common.cuh"

namespace bppgpu {

enum : unsigned { EIGEN_KEEP = 0, EIGEN_COPY = 1, EIGEN_COMPUTE = 2 };

// a: S x S symmetric (destroyed), vec: S x S receives the eigenvectors as COLUMNS, lam: S eigenvalues
__device__ inline void jacobi_eigen_dev(double * a, double * vec, double * lam, int n)
{
  for (int i = 0; i < n * n; ++i) vec[i] = 0.0;
  for (int i = 0; i < n; ++i) vec[i * n + i] = 1.0;
  for (int sweep = 0; sweep < 100; ++sweep)
  {
    double off = 0;
    for (int p = 0; p < n; ++p) for (int q = p + 1; q < n; ++q) off += a[p * n + q] * a[p * n + q];
    if (off < 1e-300) break;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q)
      {
        const double apq = a[p * n + q];
        if (fabs(apq) < 1e-300) continue;
        const double app = a[p * n + p], aqq = a[q * n + q];
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k)
        {
          const double akp = a[k * n + p], akq = a[k * n + q];
          a[k * n + p] = c * akp - s * akq;
          a[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k)
        {
          const double apk = a[p * n + k], aqk = a[q * n + k];
          a[p * n + k] = c * apk - s * aqk;
          a[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k)
        {
          const double vkp = vec[k * n + p], vkq = vec[k * n + q];
          vec[k * n + p] = c * vkp - s * vkq;
          vec[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < n; ++i) lam[i] = a[i * n + i];
}

// One block per dirty locus.  Record of locus d at stage + rec_off[d]:
//   [freqs S][rates R][rate_weights R][subst S(S-1)/2] and, for EIGEN_COPY, [V S*S][V^-1 S*S][lambda S].
// scratch: 2*S*S doubles per record for EIGEN_COMPUTE.
__global__ void __launch_bounds__(64)
model_update_kernel(const LocusDev * __restrict__ loci, const unsigned int * __restrict__ ids,
                    const double * __restrict__ stage, const unsigned long long * __restrict__ rec_off,
                    const unsigned char * __restrict__ eigen_mode, double * __restrict__ scratch,
                    const unsigned long long * __restrict__ scratch_off)
{
  const unsigned int d = blockIdx.x;
  const LocusDev & L = loci[ids[d]];
  const unsigned int S = L.states, R = L.rate_cats, np = S * (S - 1) / 2;
  const double * rec = stage + rec_off[d];
  for (unsigned int i = threadIdx.x; i < S; i += blockDim.x) L.freqs[i] = rec[i];
  for (unsigned int i = threadIdx.x; i < R; i += blockDim.x) { L.rates[i] = rec[S + i]; L.rate_weights[i] = rec[S + R + i]; }
  for (unsigned int i = threadIdx.x; i < np; i += blockDim.x) L.subst[i] = rec[S + 2 * R + i];
  const unsigned int mode = eigen_mode[d];
  if (mode == EIGEN_COPY)
  {
    const double * e = rec + S + 2 * R + np;
    for (unsigned int i = threadIdx.x; i < S * S; i += blockDim.x) { L.eigenvecs[i] = e[i]; L.inv_eigenvecs[i] = e[S * S + i]; }
    for (unsigned int i = threadIdx.x; i < S; i += blockDim.x) L.eigenvals[i] = e[2 * S * S + i];
  }
  else if (mode == EIGEN_COMPUTE)
  {
    const double * f = rec;
    const double * p = rec + S + 2 * R;
    double * q = scratch + scratch_off[d];
    double * vec = q + S * S;
    // create_ratematrix (core_pmatrix.c:186-237): exchangeabilities normalised by the last one
    const double last = p[np - 1];
    for (unsigned int i = threadIdx.x; i < S * S; i += blockDim.x) q[i] = 0.0;
    __syncthreads();
    if (threadIdx.x == 0)
    {
      unsigned int k = 0;
      for (unsigned int i = 0; i < S; ++i)
        for (unsigned int j = i + 1; j < S; ++j)
        {
          const double factor = last > 0.0 ? p[k] / last : p[k];
          ++k;
          q[i * S + j] = q[j * S + i] = factor * sqrt(f[i] * f[j]);
          q[i * S + i] -= factor * f[j];
          q[j * S + j] -= factor * f[i];
        }
      double mean = 0;
      for (unsigned int i = 0; i < S; ++i) mean += f[i] * (-q[i * S + i]);
      for (unsigned int i = 0; i < S * S; ++i) q[i] /= mean;
      jacobi_eigen_dev(q, vec, L.eigenvals, (int)S);
    }
    __syncthreads();
    // rows of a = eigenvectors = columns of vec (pll_update_eigen :271-290)
    for (unsigned int e = threadIdx.x; e < S * S; e += blockDim.x)
    {
      const unsigned int i = e / S, j = e % S;
      L.eigenvecs[e] = vec[j * S + i] * sqrt(f[j]);
      L.inv_eigenvecs[e] = vec[i * S + j] / sqrt(f[i]);
    }
  }
}

}  // namespace bppgpu
