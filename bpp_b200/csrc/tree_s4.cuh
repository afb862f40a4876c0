// tree_s4.cuh -- the hot kernel: tree-fused Felsenstein pruning for 4 states.
//
// One persistent CTA walks a contiguous range of tiles; a tile is blockDim.x cells
// (cell = pattern*RL + cat, RL = rate categories, a power of two <= 32 so that the RL lanes of a site
// sit in one warp) of one locus.  For its tile a thread executes the locus' WHOLE planned op list:
//   - packed tip states (4 bits per tip and site) are fetched once per tile, one tile ahead, and stay
//     in registers;
//   - a child produced by an earlier op of the same list comes from registers (SRC_PREV) or from the
//     thread's shared-memory stack (SRC_SLOT) -- it is never re-read from HBM;
//   - every inner CLV is written to HBM exactly once with a 256-bit store, scalers with the same
//     pass, and the root's site log-likelihoods are reduced in the same kernel.
// HBM traffic per locus is therefore the compulsory (T-1) CLV writes + packed tips + weights +
// the staged plan/P-matrix block (SURVEY.md 8d "B_min").
//
// Arithmetic (reference file:line, /root/reference/src):
//   x_i = (P_i0 c0 + P_i1 c1) + (P_i2 c2 + P_i3 c3), separate mul/add  core_partials_avx.c:423-473
//   parent_i = x_i * y_i                                               :476
//   site rescaling: all 4*R entries < 2^-256 (strict, unscaled)        :493-529, core_partials.c:720-754
//   root: sum_j rw_j ((pi0 c0 + pi1 c1) + (pi2 c2 + pi3 c3)), log, + scaler*log(2^-256), * weight
//                                                                      core_likelihood_avx.c:121-150
#pragma once
#include "common.cuh"

namespace bppgpu {

template <bool EXACT>
__device__ __forceinline__ double dot4(const double2 pa, const double2 pb, const double c0, const double c1,
                                       const double c2, const double c3)
{
  if (EXACT)
    return __dadd_rn(__dadd_rn(__dmul_rn(pa.x, c0), __dmul_rn(pa.y, c1)),
                     __dadd_rn(__dmul_rn(pb.x, c2), __dmul_rn(pb.y, c3)));
  return fma(pa.x, c0, pa.y * c1) + fma(pb.x, c2, pb.y * c3);
}

// 0/1 double from bit j of a state mask without a conversion instruction
__device__ __forceinline__ double bit_to_double(unsigned int code, int j)
{
  return __hiloint2double((int)(((code >> j) & 1u) * 0x3FF00000u), 0);
}

struct Cell4 { double v0, v1, v2, v3; unsigned int sc; };

template <int RL, bool EXACT>
__global__ void __launch_bounds__(256, 3)
tree_kernel_s4(const TreeParams prm)
{
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr size_t RWB = ((size_t)RL * 8 + 15) & ~(size_t)15;
  constexpr size_t CHUNKB = (size_t)TREE_CHUNK * sizeof(PlanOp) + (size_t)TREE_CHUNK * 2 * RL * PM_STRIDE * 8;
  constexpr size_t STAGEB = sizeof(LocusHdr) + RWB + CHUNKB;
  const unsigned int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31u;

  // shared memory: [stage: hdr | rw | ops | P] [desc ring 4 x 32 B] [red 32 doubles] [stack] [sstack]
  const LocusHdr * H = reinterpret_cast<const LocusHdr *>(smem);
  const double * s_rw = reinterpret_cast<const double *>(smem + sizeof(LocusHdr));
  const PlanOp * s_ops = reinterpret_cast<const PlanOp *>(smem + sizeof(LocusHdr) + RWB);
  const double * s_pm = reinterpret_cast<const double *>(smem + sizeof(LocusHdr) + RWB + TREE_CHUNK * sizeof(PlanOp));
  TileDesc * s_desc = reinterpret_cast<TileDesc *>(smem + STAGEB);
  double * s_red = reinterpret_cast<double *>(smem + STAGEB + 4 * sizeof(TileDesc));
  double2 * s_stack = reinterpret_cast<double2 *>(smem + STAGEB + 4 * sizeof(TileDesc) + 32 * 8);
  unsigned int * s_sstack = reinterpret_cast<unsigned int *>(s_stack + (size_t)prm.n_slots * 2 * nthr);

  const unsigned int t_begin = (unsigned int)(((unsigned long long)prm.n_tiles * blockIdx.x) / gridDim.x);
  const unsigned int t_end = (unsigned int)(((unsigned long long)prm.n_tiles * (blockIdx.x + 1)) / gridDim.x);
  if (t_begin >= t_end) return;

  // descriptor ring: tiles t_begin and t_begin+1
  if (tid < 4)
  {
    const unsigned int t = t_begin + (tid >> 1);
    if (t < t_end)
      cp_async16(reinterpret_cast<unsigned char *>(&s_desc[t & 3u]) + (tid & 1u) * 16,
                 reinterpret_cast<const unsigned char *>(prm.tiles + t) + (tid & 1u) * 16);
  }
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();

  // tips / weight of the first tile
  unsigned int tw0 = 0, tw1 = 0, wgt = 0;
  {
    const TileDesc d = s_desc[t_begin & 3u];
    const unsigned int craw = d.cell0 + tid;
    const unsigned int pat = (craw < d.ncell ? craw : d.ncell - 1) / RL;
    tw0 = __ldg(d.tipwords + (size_t)pat * d.tip_words);
    if (d.tip_words > 1) tw1 = __ldg(d.tipwords + (size_t)pat * d.tip_words + 1);
    wgt = __ldg(d.weights + pat);
  }
  unsigned int cur_locus = 0xFFFFFFFFu;

  for (unsigned int t = t_begin; t < t_end; ++t)
  {
    const TileDesc d = s_desc[t & 3u];
    // ---- prefetch for the next tiles: tips/weight of t+1 into registers, descriptor of t+2 into the ring
    unsigned int ntw0 = 0, ntw1 = 0, nwgt = 0;
    if (t + 1 < t_end)
    {
      const TileDesc dn = s_desc[(t + 1) & 3u];
      const unsigned int craw = dn.cell0 + tid;
      const unsigned int pat = (craw < dn.ncell ? craw : dn.ncell - 1) / RL;
      ntw0 = __ldg(dn.tipwords + (size_t)pat * dn.tip_words);
      if (dn.tip_words > 1) ntw1 = __ldg(dn.tipwords + (size_t)pat * dn.tip_words + 1);
      nwgt = __ldg(dn.weights + pat);
    }
    if (tid < 2 && t + 2 < t_end)
      cp_async16(reinterpret_cast<unsigned char *>(&s_desc[(t + 2) & 3u]) + tid * 16,
                 reinterpret_cast<const unsigned char *>(prm.tiles + t + 2) + tid * 16);
    cp_async_commit();

    // ---- stage the locus block (header, rate weights, first chunk) when the locus changes
    const unsigned char * gblk = prm.blocks + prm.blk_off[d.locus];
    if (d.locus != cur_locus)
    {
      const uint4 * src = reinterpret_cast<const uint4 *>(gblk);
      uint4 * dst = reinterpret_cast<uint4 *>(smem);
      for (unsigned int w = tid; w < STAGEB / 16; w += nthr) dst[w] = __ldg(src + w);
      cur_locus = d.locus;
      __syncthreads();
    }
    const unsigned int nops = H->nops;
    const unsigned int ncell = d.ncell;
    const unsigned int cell_raw = d.cell0 + tid;
    const bool valid = cell_raw < ncell;
    const unsigned int cell = valid ? cell_raw : ncell - 1;
    const unsigned int pattern = cell / RL;
    const unsigned int cat = cell % RL;
    double * const clv_cell = H->clv + (size_t)cell * 4;
    const unsigned long long stride = H->clv_stride;
    const unsigned int sites = H->sites;

    double p0 = 0, p1 = 0, p2 = 0, p3 = 0;    // result of the previous op (SRC_PREV)
    unsigned int psc = 0;
    double site_val = 0.0;

    for (unsigned int base = 0; base < nops; base += TREE_CHUNK)
    {
      if (base)
      {
        // trees with more than TREE_CHUNK inner nodes: restage chunk by chunk
        __syncthreads();
        const uint4 * src = reinterpret_cast<const uint4 *>(gblk + sizeof(LocusHdr) + RWB + (size_t)(base / TREE_CHUNK) * CHUNKB);
        uint4 * dst = reinterpret_cast<uint4 *>(smem + sizeof(LocusHdr) + RWB);
        for (unsigned int w = tid; w < CHUNKB / 16; w += nthr) dst[w] = __ldg(src + w);
        cur_locus = 0xFFFFFFFFu;               // chunk 0 is gone
        __syncthreads();
      }
      const unsigned int cn = min((unsigned)TREE_CHUNK, nops - base);
      for (unsigned int k = 0; k < cn; ++k)
      {
        const PlanOp q = s_ops[k];
        const unsigned int spill = q.ctl & CTL_SPILL_MASK;
        if (spill)
        {
          s_stack[((spill - 1) * 2 + 0) * nthr + tid] = make_double2(p0, p1);
          s_stack[((spill - 1) * 2 + 1) * nthr + tid] = make_double2(p2, p3);
          s_sstack[(spill - 1) * nthr + tid] = psc;
        }
        double l0, l1, l2, l3, r0, r1, r2, r3;
        unsigned int lsc = 0, rsc = 0;
        // ---- left operand
        {
          const unsigned int kind = q.lsrc >> 28, idx = q.lsrc & 0x0FFFFFFFu;
          if (kind == SRC_TIP_PACKED)
          {
            unsigned int word = (idx < 8) ? tw0 : tw1;
            if (idx >= 16) word = __ldg(H->tipwords + (size_t)pattern * H->tip_words + (idx >> 3));
            const unsigned int code = word >> ((idx & 7u) * 4);
            l0 = bit_to_double(code, 0); l1 = bit_to_double(code, 1); l2 = bit_to_double(code, 2); l3 = bit_to_double(code, 3);
          }
          else if (kind == SRC_PREV) { l0 = p0; l1 = p1; l2 = p2; l3 = p3; lsc = psc; }
          else if (kind == SRC_SLOT)
          {
            const double2 a = s_stack[(idx * 2 + 0) * nthr + tid], b = s_stack[(idx * 2 + 1) * nthr + tid];
            l0 = a.x; l1 = a.y; l2 = b.x; l3 = b.y; lsc = s_sstack[idx * nthr + tid];
          }
          else if (kind == SRC_TIP_DENSE) ld256_nc(H->tip_dense + (size_t)idx * stride + (size_t)cell * 4, l0, l1, l2, l3);
          else
          {
            ld256(clv_cell + (size_t)idx * stride, l0, l1, l2, l3);
            if (q.lsc >= 0) lsc = H->scale[(size_t)q.lsc * sites + pattern];
          }
        }
        double o0, o1, o2, o3;
        unsigned int osc;
        if (q.ctl & CTL_EVAL_ONLY)
        {
          o0 = l0; o1 = l1; o2 = l2; o3 = l3; osc = lsc;
        }
        else
        {
          // ---- right operand
          const unsigned int kind = q.rsrc >> 28, idx = q.rsrc & 0x0FFFFFFFu;
          if (kind == SRC_TIP_PACKED)
          {
            unsigned int word = (idx < 8) ? tw0 : tw1;
            if (idx >= 16) word = __ldg(H->tipwords + (size_t)pattern * H->tip_words + (idx >> 3));
            const unsigned int code = word >> ((idx & 7u) * 4);
            r0 = bit_to_double(code, 0); r1 = bit_to_double(code, 1); r2 = bit_to_double(code, 2); r3 = bit_to_double(code, 3);
          }
          else if (kind == SRC_PREV) { r0 = p0; r1 = p1; r2 = p2; r3 = p3; rsc = psc; }
          else if (kind == SRC_SLOT)
          {
            const double2 a = s_stack[(idx * 2 + 0) * nthr + tid], b = s_stack[(idx * 2 + 1) * nthr + tid];
            r0 = a.x; r1 = a.y; r2 = b.x; r3 = b.y; rsc = s_sstack[idx * nthr + tid];
          }
          else if (kind == SRC_TIP_DENSE) ld256_nc(H->tip_dense + (size_t)idx * stride + (size_t)cell * 4, r0, r1, r2, r3);
          else
          {
            ld256(clv_cell + (size_t)idx * stride, r0, r1, r2, r3);
            if (q.rsc >= 0) rsc = H->scale[(size_t)q.rsc * sites + pattern];
          }
          // ---- parent = (P_l . l) * (P_r . r)
          const double2 * __restrict__ pl = reinterpret_cast<const double2 *>(s_pm + ((k * 2 + 0) * RL + cat) * PM_STRIDE);
          const double2 * __restrict__ pr = reinterpret_cast<const double2 *>(s_pm + ((k * 2 + 1) * RL + cat) * PM_STRIDE);
          const double x0 = dot4<EXACT>(pl[0], pl[1], l0, l1, l2, l3), y0 = dot4<EXACT>(pr[0], pr[1], r0, r1, r2, r3);
          const double x1 = dot4<EXACT>(pl[2], pl[3], l0, l1, l2, l3), y1 = dot4<EXACT>(pr[2], pr[3], r0, r1, r2, r3);
          const double x2 = dot4<EXACT>(pl[4], pl[5], l0, l1, l2, l3), y2 = dot4<EXACT>(pr[4], pr[5], r0, r1, r2, r3);
          const double x3 = dot4<EXACT>(pl[6], pl[7], l0, l1, l2, l3), y3 = dot4<EXACT>(pr[6], pr[7], r0, r1, r2, r3);
          o0 = __dmul_rn(x0, y0); o1 = __dmul_rn(x1, y1); o2 = __dmul_rn(x2, y2); o3 = __dmul_rn(x3, y3);
          osc = 0;
          if (q.dsc >= 0)
          {
            osc = lsc + rsc;
            unsigned int below = (o0 < BPPGPU_SCALE_THRESHOLD) & (o1 < BPPGPU_SCALE_THRESHOLD) &
                                 (o2 < BPPGPU_SCALE_THRESHOLD) & (o3 < BPPGPU_SCALE_THRESHOLD);
#pragma unroll
            for (int dd = 1; dd < RL; dd <<= 1) below &= __shfl_xor_sync(0xFFFFFFFFu, below, dd);
            if (below)
            {
              o0 = __dmul_rn(o0, BPPGPU_SCALE_FACTOR); o1 = __dmul_rn(o1, BPPGPU_SCALE_FACTOR);
              o2 = __dmul_rn(o2, BPPGPU_SCALE_FACTOR); o3 = __dmul_rn(o3, BPPGPU_SCALE_FACTOR);
              osc += 1;
            }
            if (valid && cat == 0) H->scale[(size_t)q.dsc * sites + pattern] = osc;
          }
          if (valid) st256(clv_cell + (size_t)q.dst * stride, o0, o1, o2, o3);
        }
        p0 = o0; p1 = o1; p2 = o2; p3 = o3; psc = osc;

        if (q.ctl & CTL_ROOT)
        {
          const double tr = __dadd_rn(__dadd_rn(__dmul_rn(H->freqs[0], o0), __dmul_rn(H->freqs[1], o1)),
                                      __dadd_rn(__dmul_rn(H->freqs[2], o2), __dmul_rn(H->freqs[3], o3)));
          double term = 0.0;
#pragma unroll
          for (int j = 0; j < RL; ++j)
          {
            const double v = __shfl_sync(0xFFFFFFFFu, tr, (lane & ~(unsigned)(RL - 1)) + j);
            term = __dadd_rn(term, __dmul_rn(v, s_rw[j]));
          }
          unsigned int rs = osc;
          if (q.ctl & CTL_EVAL_ONLY) rs = (q.root_sc >= 0) ? osc : 0;
          double s;
          if (prm.persite_mode == 2) s = term;
          else
          {
            s = log(term);
            if (rs) s = __dadd_rn(s, __dmul_rn((double)rs, prm.log_threshold));
            s = __dmul_rn(s, (double)wgt);
          }
          if (valid && cat == 0)
          {
            site_val = s;
            if (prm.persite) prm.persite[pattern] = s;
          }
        }
      }
    }

    // ---- deterministic tile reduction of the weighted site lnL values
    if (prm.tile_partial)
    {
      double v = site_val;
#pragma unroll
      for (int dd = 16; dd > 0; dd >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, dd);
      if (lane == 0) s_red[(t & 1u) * 16 + (tid >> 5)] = v;
    }
    cp_async_wait_all();
    __syncthreads();                           // s_red complete; descriptor t+2 visible; stage reusable
    if (prm.tile_partial && tid == 0)
    {
      double acc = 0.0;
      for (unsigned int w = 0; w < (nthr >> 5); ++w) acc += s_red[(t & 1u) * 16 + w];
      prm.tile_partial[t] = acc;
    }
    tw0 = ntw0; tw1 = ntw1; wgt = nwgt;
  }
}

}  // namespace bppgpu
