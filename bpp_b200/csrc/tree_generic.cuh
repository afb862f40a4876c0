// tree_generic.cuh -- semantic fallback kernel: any state count, any number of rate categories
#pragma once
#include "common.cuh"

namespace bppgpu {

// ----------------------------------------------------------------------------- generic tree kernel
// Any state count / any number of rate categories: one thread per PATTERN loops over categories and
// states, every operand comes from HBM (the plan is built with 0 slots and SRC_PREV disabled).
// This is the semantic fallback (R not a power of two, exotic state counts); the tuned paths are
// tree_kernel_s4 and the 20-state kernel.
__device__ __forceinline__ unsigned int tip_code(const LocusDev & L, unsigned int tip, unsigned int pattern)
{
  const unsigned int * w = reinterpret_cast<const unsigned int *>(L.tip_codes);
  if (L.states == 4) return (w[(size_t)pattern * L.tip_words + (tip >> 3)] >> ((tip & 7u) * 4)) & 0xFu;
  return w[(size_t)tip * L.sites + pattern];
}

template <bool EXACT>
__global__ void __launch_bounds__(128)
tree_kernel_generic(const TreeParams prm)
{
  __shared__ double s_red[4];
  const unsigned int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31u;
  const unsigned int tile = blockIdx.x;
  const unsigned int bl = prm.tile_locus[tile];
  __shared__ LocusDev L;
  {
    const unsigned int * src = reinterpret_cast<const unsigned int *>(prm.loci + prm.batch_locus[bl]);
    unsigned int * dst = reinterpret_cast<unsigned int *>(&L);
    for (unsigned int w = tid; w < sizeof(LocusDev) / 4; w += nthr) dst[w] = src[w];
  }
  __syncthreads();
  const unsigned int nops = prm.plan_count[bl];
  const PlanOp * __restrict__ gplan = prm.plan + prm.op_off[bl] + bl;
  const unsigned int S = L.states, R = L.rate_cats, CS = L.cat_stride;
  const unsigned int praw = prm.tile_cell0[tile] + tid;       // here a "cell" is a pattern
  const bool valid = praw < L.sites;
  const unsigned int pattern = valid ? praw : L.sites - 1;
  double site_val = 0.0;

  for (unsigned int k = 0; k < nops; ++k)
  {
    const PlanOp q = gplan[k];
    const unsigned int lk = q.lsrc >> 28, li = q.lsrc & 0x0FFFFFFFu;
    const unsigned int rk = q.rsrc >> 28, ri = q.rsrc & 0x0FFFFFFFu;
    const double * lp = nullptr, * rp = nullptr;
    unsigned int lcode = 0, rcode = 0;
    if (lk == SRC_TIP_PACKED) lcode = tip_code(L, li, pattern);
    else lp = ((lk == SRC_TIP_DENSE) ? L.tip_dense : L.clv) + (size_t)li * L.clv_stride + (size_t)pattern * L.site_stride;
    if (!(q.ctl & CTL_EVAL_ONLY))
    {
      if (rk == SRC_TIP_PACKED) rcode = tip_code(L, ri, pattern);
      else rp = ((rk == SRC_TIP_DENSE) ? L.tip_dense : L.clv) + (size_t)ri * L.clv_stride + (size_t)pattern * L.site_stride;
    }
    unsigned int osc = 0;
    if (lk == SRC_HBM && q.lsc >= 0) osc += L.scale[(size_t)q.lsc * L.sites + pattern];

    if (!(q.ctl & CTL_EVAL_ONLY))
    {
      if (rk == SRC_HBM && q.rsc >= 0) osc += L.scale[(size_t)q.rsc * L.sites + pattern];
      double * out = L.clv + (size_t)q.dst * L.clv_stride + (size_t)pattern * L.site_stride;
      bool below = true;
      for (unsigned int n = 0; n < R; ++n)
      {
        const double * __restrict__ Pl = L.pmat + ((size_t)q.lpm * R + n) * S * S;
        const double * __restrict__ Pr = L.pmat + ((size_t)q.rpm * R + n) * S * S;
        for (unsigned int i = 0; i < S; ++i)
        {
          // four lane sums over columns == 0..3 (mod 4), combined (s0+s1)+(s2+s3):
          // core_partials_avx.c:1330-1567 (mul+add) / core_partials_avx2.c:666-726 (fma)
          double xa[4] = {0, 0, 0, 0}, ya[4] = {0, 0, 0, 0};
          for (unsigned int j = 0; j < S; ++j)
          {
            const double lv = lp ? lp[(size_t)n * CS + j] : (((lcode >> j) & 1u) ? 1.0 : 0.0);
            const double rv = rp ? rp[(size_t)n * CS + j] : (((rcode >> j) & 1u) ? 1.0 : 0.0);
            if (EXACT)
            {
              xa[j & 3] = __dadd_rn(xa[j & 3], __dmul_rn(Pl[i * S + j], lv));
              ya[j & 3] = __dadd_rn(ya[j & 3], __dmul_rn(Pr[i * S + j], rv));
            }
            else
            {
              xa[j & 3] = fma(Pl[i * S + j], lv, xa[j & 3]);
              ya[j & 3] = fma(Pr[i * S + j], rv, ya[j & 3]);
            }
          }
          const double x = __dadd_rn(__dadd_rn(xa[0], xa[1]), __dadd_rn(xa[2], xa[3]));
          const double y = __dadd_rn(__dadd_rn(ya[0], ya[1]), __dadd_rn(ya[2], ya[3]));
          const double o = __dmul_rn(x, y);
          below = below && (o < BPPGPU_SCALE_THRESHOLD);
          if (valid) out[(size_t)n * CS + i] = o;
        }
      }
      if (q.dsc >= 0)
      {
        if (below)
        {
          if (valid) for (unsigned int n = 0; n < R; ++n) for (unsigned int e = 0; e < S; ++e) out[(size_t)n * CS + e] = __dmul_rn(out[(size_t)n * CS + e], BPPGPU_SCALE_FACTOR);
          osc += 1;
        }
        if (valid) L.scale[(size_t)q.dsc * L.sites + pattern] = osc;
      }
      else osc = 0;
    }

    if (q.ctl & CTL_ROOT)
    {
      const double * rc = (q.ctl & CTL_EVAL_ONLY) ? lp : (L.clv + (size_t)q.dst * L.clv_stride + (size_t)pattern * L.site_stride);
      double term = 0.0;
      for (unsigned int n = 0; n < R; ++n)
      {
        double la[4] = {0, 0, 0, 0};
        for (unsigned int j = 0; j < S; ++j)
        {
          const double cv = rc ? rc[(size_t)n * CS + j] : (((lcode >> j) & 1u) ? 1.0 : 0.0);
          la[j & 3] = __dadd_rn(la[j & 3], __dmul_rn(L.freqs[j], cv));
        }
        const double tr = __dadd_rn(__dadd_rn(la[0], la[1]), __dadd_rn(la[2], la[3]));
        term = __dadd_rn(term, __dmul_rn(tr, L.rate_weights[n]));
      }
      unsigned int rs = osc;
      if (q.ctl & CTL_EVAL_ONLY) rs = (q.root_sc >= 0) ? osc : 0;
      double s;
      if (prm.persite_mode == 2) s = term;
      else
      {
        s = log(term);
        if (rs) s = __dadd_rn(s, __dmul_rn((double)rs, prm.log_threshold));
        s = __dmul_rn(s, (double)L.weights[pattern]);
      }
      if (valid)
      {
        site_val = s;
        if (prm.persite) prm.persite[pattern] = s;
      }
    }
    // a later op of this list may read what this thread just wrote
    __threadfence_block();
  }

  if (prm.tile_partial)
  {
    double v = site_val;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    if (lane == 0) s_red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0)
    {
      double acc = 0.0;
      for (unsigned int w = 0; w < (nthr >> 5); ++w) acc += s_red[w];
      prm.tile_partial[tile] = acc;
    }
  }
}

}  // namespace bppgpu
