// pmatrix.cuh -- transition-probability matrices (locus.c:2325-2415, core_pmatrix.c:674-783)
#pragma once
#include "common.cuh"

namespace bppgpu {


// Closed-form DNA models other than JC69, one whole matrix per call.  Expressions and their order are
// the reference's: K80 locus.c:2300-2322, F81 :2236-2254, HKY / F84 / TN93 :2116-2162, T92 :2039-2066
// (no zero-length special case there: expm1(-0) makes the identity by itself).
__device__ __forceinline__ void pmatrix_closed4(const LocusDev & L, double bl, double * __restrict__ pmat)
{
  const double * __restrict__ freqs = L.freqs;
  const double * __restrict__ qrates = L.subst;
  if (L.model_kind == MODEL_K80)
  {
    const double kappa = qrates[1] / qrates[0];
    const double e1 = expm1(-4 * bl / (kappa + 2));
    if (fabs(kappa - 1) < 1e-20)
    {
      for (int j = 0; j < 4; ++j)
        for (int k = 0; k < 4; ++k) pmat[j * 4 + k] = (j == k) ? 1. + 3 / 4. * e1 : -e1 / 4;
    }
    else
    {
      const double e2 = expm1(-2 * bl * (kappa + 1) / (kappa + 2));
      const double dg = 1 + (e1 + 2 * e2) / 4, tv = -e1 / 4, ts = (e1 - 2 * e2) / 4;
      for (int j = 0; j < 4; ++j)
        for (int k = 0; k < 4; ++k) pmat[j * 4 + k] = (j == k) ? dg : (((j ^ k) == 2) ? ts : tv);
    }
  }
  else if (L.model_kind == MODEL_F81)
  {
    double beta = 1;
    for (int j = 0; j < 4; ++j) beta -= freqs[j] * freqs[j];
    beta = 1. / beta;
    const double e = exp(-beta * bl), em1 = expm1(-beta * bl);
    for (int j = 0; j < 4; ++j)
      for (int k = 0; k < 4; ++k) pmat[j * 4 + k] = (j == k) ? e - freqs[k] * em1 : -freqs[k] * em1;
  }
  else if (L.model_kind == MODEL_T92)
  {
    const double GC = freqs[3] + freqs[2];
    const double e1 = expm1(-bl);
    const double e2 = expm1(-(qrates[0] / qrates[1] + 1) * bl / 2);
    pmat[0]  = -(1 - GC) / 2 * e1;
    pmat[1]  = GC / 2 * e1 - GC * e2;
    pmat[2]  = -GC / 2 * e1;
    pmat[3]  = 1 + 0.5 * (1 - GC) * e1 + GC * e2;
    pmat[4]  = -(1 - GC) / 2 * e1;
    pmat[5]  = 1 + GC / 2 * e1 + (1 - GC) * e2;
    pmat[6]  = -GC / 2 * e1;
    pmat[7]  = (1 - GC) / 2 * e1 - (1 - GC) * e2;
    pmat[8]  = 1 + 0.5 * (1 - GC) * e1 + GC * e2;
    pmat[9]  = -GC / 2 * e1;
    pmat[10] = GC / 2 * e1 - GC * e2;
    pmat[11] = -(1 - GC) / 2 * e1;
    pmat[12] = (1 - GC) / 2 * e1 - (1 - GC) * e2;
    pmat[13] = -GC / 2 * e1;
    pmat[14] = 1 + GC / 2 * e1 + (1 - GC) * e2;
    pmat[15] = -(1 - GC) / 2 * e1;
  }
  else    // HKY, F84, TN93 share the TN93 formulas
  {
    const double A = freqs[0], C = freqs[1], G = freqs[2], T = freqs[3];
    const double Y = T + C, R = A + G;
    double bt, a1t, a2t;
    if (L.model_kind == MODEL_HKY)
    {
      const double kappa = qrates[1] / qrates[0];
      const double mr = 1 / (2 * T * C * kappa + 2 * A * G * kappa + 2 * Y * R);
      bt = bl * mr;
      a1t = a2t = kappa * bt;
    }
    else if (L.model_kind == MODEL_F84)
    {
      const double kappa = qrates[0] / qrates[1];
      const double mr = 1 / (2 * T * C * kappa + 2 * A * G * kappa + 2 * Y * R);
      bt = bl * mr;
      a1t = (1 + kappa / Y) * bt;
      a2t = (1 + kappa / R) * bt;
    }
    else
    {
      const double mr = 1 / (2 * T * C * qrates[0] + 2 * A * G * qrates[1] + 2 * Y * R);
      bt = bl * mr;
      a1t = (qrates[0] / qrates[2]) * bt;
      a2t = (qrates[1] / qrates[2]) * bt;
    }
    const double e1 = expm1(-bt);
    const double e2 = expm1(-(R * a2t + Y * bt));
    const double e3 = expm1(-(Y * a1t + R * bt));
    pmat[0]  = 1 + Y * A / R * e1 + G / R * e2;
    pmat[1]  = -C * e1;
    pmat[2]  = Y * G / R * e1 - G / R * e2;
    pmat[3]  = -T * e1;
    pmat[4]  = -A * e1;
    pmat[5]  = 1 + (R * C * e1 + T * e3) / Y;
    pmat[6]  = -G * e1;
    pmat[7]  = (R * e1 - e3) * T / Y;
    pmat[8]  = Y * A / R * e1 - A / R * e2;
    pmat[9]  = -C * e1;
    pmat[10] = 1 + Y * G / R * e1 + A / R * e2;
    pmat[11] = -T * e1;
    pmat[12] = -A * e1;
    pmat[13] = (R * e1 - e3) * C / Y;
    pmat[14] = -G * e1;
    pmat[15] = 1 + (R * T * e1 + C * e3) / Y;
  }
}

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
  if (L.model_kind >= MODEL_K80)
  {
    double m[16];
    pmatrix_closed4(L, bt, m);
    for (unsigned int k = 0; k < 4; ++k) row[k] = m[j * 4 + k];
  }
  else if (bt < 1e-100)
  {
    for (unsigned int k = 0; k < S; ++k) row[k] = (j == k) ? 1.0 : 0.0;
  }
  else if (L.model_kind == MODEL_JC69)
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
  if (L.model_kind >= MODEL_K80)
  {
    double m[16];
    pmatrix_closed4(L, bt, m);
#pragma unroll
    for (int j = 0; j < 8; ++j) P[j] = make_double2(m[2 * j], m[2 * j + 1]);
  }
  else if (bt < 1e-100)
  {
#pragma unroll
    for (int j = 0; j < 4; ++j) { P[2 * j] = make_double2(j == 0 ? 1.0 : 0.0, j == 1 ? 1.0 : 0.0); P[2 * j + 1] = make_double2(j == 2 ? 1.0 : 0.0, j == 3 ? 1.0 : 0.0); }
  }
  else if (L.model_kind == MODEL_JC69)
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
  if (S == 4)
  {
    for (unsigned int t = threadIdx.x; t < count * R; t += blockDim.x)
      pmatrix_full4(L, mat_idx[first + t / R], mat_bl[first + t / R], t % R);
    return;
  }
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

// 20 states on the FP64 tensor instruction: P = I + (V^-1 diag(expm1(lambda t))) . V is a 20x20x20 GEMM per
// (branch, category), 45 DMMA.8x8x4 on 24x24 padded tiles.  V (B fragments) and V^-1 (the base of the A
// fragments) are the same for every matrix of the locus and stay in registers; per matrix a warp computes the 20
// expm1 values (one per lane, handed round by shuffles), scales its A fragments, runs the DMMA chains from the
// identity and stores the accumulators as 16-byte pieces.  No shared memory at all -- pmatrix_kernel_wide is
// bound by its shared-memory loads (profiles/r1_pmatrix_wide_v2_config4_ncu_summary.txt).  The m-sum is a fused
// chain here instead of the reference's separate multiply and add (core_pmatrix.c:760-771): the matrices agree
// with the reference's to ~1e-16, like the CLVs of the 20-state tree kernels that consume them.
__global__ void __launch_bounds__(128)
pmatrix_kernel_dmma20(const LocusDev * __restrict__ loci, const unsigned int * __restrict__ batch_locus,
                      const unsigned int * __restrict__ mat_off, const unsigned int * __restrict__ mat_idx,
                      const double * __restrict__ mat_bl)
{
  constexpr unsigned int S = 20, SS = 400;
  const unsigned int bl = blockIdx.x;
  const LocusDev & L = loci[batch_locus[bl]];
  const unsigned int first = mat_off[bl], count = mat_off[bl + 1] - first;
  const unsigned int R = L.rate_cats;
  const unsigned int lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const unsigned int r = lane >> 2, q = lane & 3u;
  if ((blockIdx.y * nw + warp) >= count * R) return;
  const double * __restrict__ V = L.eigenvecs;
  const double * __restrict__ Vi = L.inv_eigenvecs;
  double vi[3][5], vb[3][5];
#pragma unroll
  for (int t = 0; t < 3; ++t)
#pragma unroll
    for (int ks = 0; ks < 5; ++ks)
    {
      const unsigned int i = 8 * t + r;
      vi[t][ks] = (i < S) ? Vi[i * S + 4 * ks + q] : 0.0;          // A: row 8t+r, k = 4ks+q
      vb[t][ks] = (i < S) ? V[(4 * ks + q) * S + i] : 0.0;         // B: k = 4ks+q, column 8t+r
    }
  const double lam = lane < S ? L.eigenvals[lane] : 0.0;
  for (unsigned int g = blockIdx.y * nw + warp; g < count * R; g += gridDim.y * nw)
  {
    const unsigned int n = g % R, m = g / R;
    const double bt = mat_bl[first + m] * L.rates[n];
    double * P = L.pmat + ((size_t)mat_idx[first + m] * R + n) * SS;
    if (bt < 1e-100)                                               // core_pmatrix.c:738-743
    {
      for (unsigned int t = lane; t < SS; t += 32) P[t] = (t / S == t % S) ? 1.0 : 0.0;
      continue;
    }
    const double e = expm1(lam * bt);                              // :753-754, lane = eigenvalue index
    double a[3][5];
#pragma unroll
    for (int ks = 0; ks < 5; ++ks)
    {
      const double ex = __shfl_sync(0xFFFFFFFFu, e, 4 * ks + q);
#pragma unroll
      for (int t = 0; t < 3; ++t) a[t][ks] = __dmul_rn(vi[t][ks], ex);      // :756-758
    }
    double c[3][3][2];
#pragma unroll
    for (int mt = 0; mt < 3; ++mt)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
      {
        c[mt][nt][0] = (8 * mt + r == 8 * nt + 2 * q) ? 1.0 : 0.0;
        c[mt][nt][1] = (8 * mt + r == 8 * nt + 2 * q + 1) ? 1.0 : 0.0;
      }
#pragma unroll
    for (int ks = 0; ks < 5; ++ks)
#pragma unroll
      for (int mt = 0; mt < 3; ++mt)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
          asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
              : "+d"(c[mt][nt][0]), "+d"(c[mt][nt][1]) : "d"(a[mt][ks]), "d"(vb[nt][ks]));
#pragma unroll
    for (int mt = 0; mt < 3; ++mt)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
      {
        const unsigned int i = 8 * mt + r, k = 8 * nt + 2 * q;
        if (i < S && k < S) *reinterpret_cast<double2 *>(P + i * S + k) = make_double2(c[mt][nt][0], c[mt][nt][1]);
      }
  }
}

}  // namespace bppgpu
