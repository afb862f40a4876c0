// tree_s20c.cuh -- 20-state tree kernel, category-major tiling with the P-matrices staged in shared memory.
//
// tree_s20.cuh walks tiles of 32 sites x RL categories and reads every P-matrix element through L1 from
// the locus' pmatrix block: per tile that is more P-matrix traffic (14 edges x RL x 3.2 kB) than CLV
// output, and its scattered 8-byte loads saturate the L1 wavefront pipe (ncu: r1_s20_v2).  Here a tile is
// 256 sites of ONE category of one locus, a CTA (16 warps, one per 16 sites) keeps all the matrices
// that (locus, category) needs in shared memory across the locus' site blocks, and every matrix access of
// the op loop -- DMMA A fragments, tip-column gathers -- is a conflict-free shared-memory load:
//   stage[m] = 20 rows x 28 doubles: columns 0..19 = P, 20..23 = the tip edge's ambiguity columns.
//   Row stride 28 doubles puts the four rows a half-warp touches on disjoint bank groups.
// Tip column ids are staged per warp (T x 16 bytes); a parked X lives in a shared-memory stack slot.
// The warp's CLV tile (16 sites x 20 states) rotates the states of sites 4..7 and 12..15 by 4 positions
// (tile_idx): accumulator-layout stores, B-fragment loads and the 16-byte row copies are then all free of
// bank conflicts (without it the stores are 2-way conflicted: sites 4 apart are 160 banks apart).
//
// Categories of a site are now in different CTAs, so
//   * the root only writes pi . clv per (category, site); root20_kernel (reduce.cuh) combines the
//     categories, takes the log, applies the pattern weight and reduces the locus in a fixed order;
//   * per-site scaling (an AND over all categories of a site, core_partials.c:739-754) cannot be done:
//     batches holding a locus with scale buffers run tree_kernel_s20 instead (engine.cu picks).
// Same reference semantics as tree_s20.cuh: core_partials.c:585-756, core_likelihood.c:24-212.
#pragma once
#include "tree_s20.cuh"

namespace bppgpu {

constexpr int S20C_NT = 512;                         // threads per CTA
constexpr int S20C_NW = S20C_NT / 32;                // warps
constexpr int S20C_SITES = S20C_NW * S20_WS;         // sites per tile (256)
constexpr int S20C_PST = 28;                         // staged row stride (doubles)
constexpr int S20C_PMAT = S20 * S20C_PST;            // doubles per staged matrix

__host__ inline size_t s20c_smem_bytes(unsigned cap, unsigned max_tips, int slots)
{
  return (size_t)cap * S20C_PMAT * 8 + (size_t)S20C_NW * S20_WS * S20 * 8 + (size_t)S20C_NW * max_tips * S20_WS +
         (size_t)slots * S20C_NW * 32 * S20_NG * 6 * 8 + 16;
}

// index (doubles) of state i of local site n in the warp's swizzled tile
__device__ __forceinline__ unsigned int tile_idx(unsigned int n, unsigned int i)
{
  const unsigned int j = i + (n & 4u);
  return n * S20 + (j >= S20 ? j - S20 : j);
}

// V = P . tile with P staged in shared memory (row stride S20C_PST)
__device__ __forceinline__ void matvec20_staged(const double * P, const double * tile, unsigned int r, unsigned int q,
                                                V20 & out)
{
  double a[3][5];
#pragma unroll
  for (int mt = 0; mt < 3; ++mt)
#pragma unroll
    for (int ks = 0; ks < 5; ++ks)
    {
      const unsigned int i = 8 * mt + r;
      a[mt][ks] = (i < S20) ? P[i * S20C_PST + 4 * ks + q] : 0.0;
    }
  // k-steps outermost: the S20_NG x 3 accumulator chains are independent (see matvec20)
  double b[S20_NG][5];
#pragma unroll
  for (int g = 0; g < S20_NG; ++g)
#pragma unroll
    for (int ks = 0; ks < 5; ++ks) b[g][ks] = tile[tile_idx(8 * g + r, 4 * ks + q)];
#pragma unroll
  for (int g = 0; g < S20_NG; ++g)
#pragma unroll
    for (int mt = 0; mt < 3; ++mt) out.v[g][mt][0] = out.v[g][mt][1] = 0.0;
#pragma unroll
  for (int ks = 0; ks < 5; ++ks)
#pragma unroll
    for (int g = 0; g < S20_NG; ++g)
#pragma unroll
      for (int mt = 0; mt < 3; ++mt) dmma(out.v[g][mt][0], out.v[g][mt][1], a[mt][ks], b[g][ks]);
}

template <int RL>
__global__ void __launch_bounds__(S20C_NT, 1)
tree_kernel_s20c(const TreeParams prm)
{
  extern __shared__ __align__(16) unsigned char smem20c[];
  const unsigned int tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const unsigned int r = lane >> 2, q = lane & 3u;
  const unsigned int cap = prm.lut_cap, maxT = prm.max_tips;

  double * s_stage = reinterpret_cast<double *>(smem20c);
  double * s_tile = s_stage + (size_t)cap * S20C_PMAT + (size_t)warp * S20_WS * S20;
  double * s_stack = s_stage + (size_t)cap * S20C_PMAT + (size_t)S20C_NW * S20_WS * S20;
  unsigned char * s_cols = reinterpret_cast<unsigned char *>(s_stack + (size_t)prm.n_slots * S20C_NW * 32 * S20_NG * 6) +
                           (size_t)warp * maxT * S20_WS;

  const unsigned int t_begin = (unsigned int)(((unsigned long long)prm.n_tiles * blockIdx.x) / gridDim.x);
  const unsigned int t_end = (unsigned int)(((unsigned long long)prm.n_tiles * (blockIdx.x + 1)) / gridDim.x);
  unsigned int staged_bl = 0xFFFFFFFFu, staged_cat = 0xFFFFFFFFu;

  for (unsigned int t = t_begin; t < t_end; ++t)
  {
    const unsigned int bl = prm.tile_locus[t];
    const unsigned int j = prm.tile_cell0[t];                 // tile index inside the locus: cat * nsb + site block
    const unsigned char * blk = prm.blocks + prm.tile_blk[2 * (size_t)t];
    const Hdr20 * H = reinterpret_cast<const Hdr20 *>(blk);
    const OpRec20 * recs = reinterpret_cast<const OpRec20 *>(blk + sizeof(Hdr20));
    const double * base = reinterpret_cast<const double *>(blk);
    const unsigned int sites = __ldg(&H->sites), nops = __ldg(&H->nops), T = __ldg(&H->tips);
    const unsigned int nsb = (sites + S20C_SITES - 1) / S20C_SITES;
    const unsigned int cat = j / nsb;
    const unsigned int site0 = (j % nsb) * S20C_SITES + warp * S20_WS;   // first site of this warp
    double * const clv = H->clv;
    const unsigned long long stride = __ldg(&H->clv_stride);

    // ---- stage the matrices of (locus, category): rows of 160 B from the pmatrix block, 32 B of extra columns
    if (bl != staged_bl || cat != staged_cat)
    {
      __syncthreads();                                         // every warp is done with the previous matrices
      const unsigned int n_stage = min(__ldg(&H->n_stage), cap);
      const uint2 * list = reinterpret_cast<const uint2 *>(blk + __ldg(&H->stage_off));
      const double * pmat = H->pmat;
      for (unsigned int idx = tid; idx < n_stage * (S20 * 12); idx += S20C_NT)
      {
        const unsigned int m = idx / (S20 * 12), rem = idx % (S20 * 12), row = rem / 12, ch = rem % 12;
        const uint2 ent = __ldg(list + m);
        double * dst = s_stage + (size_t)m * S20C_PMAT + row * S20C_PST + ch * 2;
        if (ch < 10) cp_async16(dst, pmat + ((size_t)ent.x * RL + cat) * (S20 * S20) + row * S20 + ch * 2);
        else if (ent.y) cp_async16(dst, base + ent.y + (size_t)cat * (S20 * S20_EXT) + row * S20_EXT + (ch - 10) * 2);
      }
      cp_async_commit();
      cp_async_wait_all();
      __syncthreads();
      staged_bl = bl; staged_cat = cat;
    }
    // ---- the warp's tip columns: s_cols[tip][16 sites]
    __syncwarp();
    for (unsigned int e = lane; e < T * S20_WS; e += 32)
    {
      const unsigned int tip = e / S20_WS, sl = e % S20_WS;
      s_cols[e] = __ldg(H->tip_cols + (size_t)tip * sites + min(site0 + sl, sites - 1));
    }
    __syncwarp();
    if (site0 >= sites) continue;                              // a warp beyond the end of the locus (no barriers below)

    unsigned int sitev[S20_NG][2];
    bool validv[S20_NG][2];
#pragma unroll
    for (int g = 0; g < S20_NG; ++g)
#pragma unroll
      for (int e = 0; e < 2; ++e)
      {
        const unsigned int s = site0 + 8 * g + 2 * q + e;
        validv[g][e] = s < sites;
        sitev[g][e] = validv[g][e] ? s : sites - 1;
      }

    V20 X;
#pragma unroll
    for (int g = 0; g < S20_NG; ++g)
#pragma unroll
      for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) X.v[g][mt][e] = 0.0;

    // tile -> global in 16-byte pieces, consecutive lanes on consecutive pieces: piece c of the warp = site c/10,
    // states 2*(c%10), 2*(c%10)+1 (the rotation moves whole pieces).  Shared-memory reads are then free of bank
    // conflicts (a lane reading 32 bytes made every LDS.128 two-way conflicted: 27 % of the kernel's wavefronts) and
    // every warp-wide 128-bit store covers 512 contiguous bytes except where it crosses to the next site.
    auto tile_to_global = [&](double * dst_buf)
    {
#pragma unroll
      for (int it = 0; it < (S20_WS * 10) / 32; ++it)
      {
        const unsigned int c = it * 32 + lane, n = c / 10, part = c % 10;
        const unsigned int s = site0 + n;
        const double2 u = *reinterpret_cast<const double2 *>(s_tile + tile_idx(n, part * 2));
        if (s < sites) st128(dst_buf + ((size_t)s * RL + cat) * S20 + part * 2, u.x, u.y);
      }
    };
    auto global_to_tile = [&](const double * src_buf, bool coherent)
    {
#pragma unroll
      for (int it = 0; it < (S20_WS * 5 + 31) / 32; ++it)
      {
        const unsigned int c = it * 32 + lane, n = c / 5, part = c % 5;
        if (c >= S20_WS * 5) break;
        const unsigned int s = min(site0 + n, sites - 1);
        double a, b, cc, d;
        if (coherent) ld256(src_buf + ((size_t)s * RL + cat) * S20 + part * 4, a, b, cc, d);
        else ld256_nc(src_buf + ((size_t)s * RL + cat) * S20 + part * 4, a, b, cc, d);
        double * dst = s_tile + tile_idx(n, part * 4);
        *reinterpret_cast<double2 *>(dst) = make_double2(a, b);
        *reinterpret_cast<double2 *>(dst + 2) = make_double2(cc, d);
      }
    };
    auto fetch = [&](unsigned int kind, unsigned int p0, unsigned int st, V20 & v)
    {
      const double * P = s_stage + (size_t)st * S20C_PMAT;
      if (kind == SRC_TIP_PACKED)
      {
        // X = column `col` of the staged matrix (ambiguity codes: columns 20..23)
#pragma unroll
        for (int g = 0; g < S20_NG; ++g)
#pragma unroll
          for (int e = 0; e < 2; ++e)
          {
            const unsigned int col = s_cols[p0 * S20_WS + 8 * g + 2 * q + e];
#pragma unroll
            for (int mt = 0; mt < 3; ++mt)
            {
              const unsigned int i = 8 * mt + r;
              v.v[g][mt][e] = (i < S20) ? P[i * S20C_PST + col] : 0.0;
            }
          }
      }
      else if (kind == SRC_SLOT)
      {
        const double * sk = s_stack + (size_t)(p0 * S20C_NW + warp) * (S20_NG * 6 * 32) + lane;
#pragma unroll
        for (int g = 0; g < S20_NG; ++g)
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int mt = 0; mt < 3; ++mt) v.v[g][mt][e] = sk[((g * 3 + mt) * 2 + e) * 32];
      }
      else
      {
        __syncwarp();
        global_to_tile((kind == SRC_TIP_DENSE ? H->tip_dense : clv) + (size_t)p0 * stride, kind == SRC_HBM);
        __syncwarp();
        matvec20_staged(P, s_tile, r, q, v);
      }
    };

    // op records are prefetched one op ahead (w2 = scalers / ext offsets is not used by this kernel)
    uint4 n0 = __ldg(reinterpret_cast<const uint4 *>(recs));
    uint4 n1 = __ldg(reinterpret_cast<const uint4 *>(recs) + 1);
    uint4 n3 = __ldg(reinterpret_cast<const uint4 *>(recs) + 3);
    for (unsigned int k = 0; k < nops; ++k)
    {
      const uint4 w0 = n0, w1 = n1, w3 = n3;
      if (k + 1 < nops)
      {
        n0 = __ldg(reinterpret_cast<const uint4 *>(recs + k + 1));
        n1 = __ldg(reinterpret_cast<const uint4 *>(recs + k + 1) + 1);
        n3 = __ldg(reinterpret_cast<const uint4 *>(recs + k + 1) + 3);
      }
      const unsigned int ctl = w0.x;
      const unsigned int akind = (ctl >> OP_AKIND_SHIFT) & 15u, bkind = (ctl >> OP_BKIND_SHIFT) & 15u;
      V20 O;

      if (ctl & OP_EVAL)
      {
        // root CLV that this list did not produce: read it as is
        if (akind == SRC_TIP_PACKED)
        {
#pragma unroll
          for (int g = 0; g < S20_NG; ++g)
#pragma unroll
            for (int e = 0; e < 2; ++e)
            {
              const unsigned int col = s_cols[w0.z * S20_WS + 8 * g + 2 * q + e];
              const unsigned int mask = (col < S20) ? (1u << col) : prm.loci[prm.batch_locus[bl]].colmask[col - S20];
#pragma unroll
              for (int mt = 0; mt < 3; ++mt) O.v[g][mt][e] = (double)((mask >> (8 * mt + r)) & 1u);
            }
        }
        else
        {
          __syncwarp();
          global_to_tile((akind == SRC_TIP_DENSE ? H->tip_dense : clv) + (size_t)w0.z * stride, akind == SRC_HBM);
          __syncwarp();
#pragma unroll
          for (int g = 0; g < S20_NG; ++g)
#pragma unroll
            for (int mt = 0; mt < 3; ++mt)
#pragma unroll
              for (int e = 0; e < 2; ++e)
              {
                const unsigned int i = 8 * mt + r;
                O.v[g][mt][e] = (i < S20) ? s_tile[tile_idx(8 * g + 2 * q + e, i)] : 0.0;
              }
        }
      }
      else
      {
        V20 A;
        fetch(akind, w0.z, w3.y, A);
        if (!(ctl & OP_BPREV)) fetch(bkind, w1.z, w3.z, X);           // B into the (dead) X registers
#pragma unroll
        for (int g = 0; g < S20_NG; ++g)
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int mt = 0; mt < 3; ++mt) O.v[g][mt][e] = A.v[g][mt][e] * X.v[g][mt][e];
        // ---- the CLV goes to HBM exactly once (through the tile for coalesced 256-bit stores)
        __syncwarp();
#pragma unroll
        for (int g = 0; g < S20_NG; ++g)
#pragma unroll
          for (int mt = 0; mt < 3; ++mt)
#pragma unroll
            for (int e = 0; e < 2; ++e)
            {
              const unsigned int i = 8 * mt + r;
              if (i < S20) s_tile[tile_idx(8 * g + 2 * q + e, i)] = O.v[g][mt][e];
            }
        __syncwarp();
        tile_to_global(clv + ((size_t)w0.y) * S20);
        // ---- push through the edge above with DMMA; the tile already holds the operand
        if (ctl & OP_PUSH)
        {
          matvec20_staged(s_stage + (size_t)w3.w * S20C_PMAT, s_tile, r, q, X);
          if (ctl & OP_PARKA)
          {
            double * sk = s_stack + (size_t)(w3.x * S20C_NW + warp) * (S20_NG * 6 * 32) + lane;
#pragma unroll
            for (int g = 0; g < S20_NG; ++g)
#pragma unroll
              for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int mt = 0; mt < 3; ++mt) sk[((g * 3 + mt) * 2 + e) * 32] = X.v[g][mt][e];
          }
        }
      }

      if (ctl & OP_ROOT)
      {
        // pi . clv of this category, reduced over the 8 row lanes; root20_kernel does the rest
        double * out = prm.rootdot + (size_t)prm.site_off[bl] * RL + (size_t)cat * sites;
#pragma unroll
        for (int g = 0; g < S20_NG; ++g)
#pragma unroll
          for (int e = 0; e < 2; ++e)
          {
            double s = 0.0;
#pragma unroll
            for (int mt = 0; mt < 3; ++mt)
              if (8 * mt + r < S20) s += __ldg(&H->freqs[8 * mt + r]) * O.v[g][mt][e];
            s += __shfl_xor_sync(0xFFFFFFFFu, s, 4);
            s += __shfl_xor_sync(0xFFFFFFFFu, s, 8);
            s += __shfl_xor_sync(0xFFFFFFFFu, s, 16);
            if (r == 0 && validv[g][e]) out[sitev[g][e]] = s;
          }
      }
    }
  }
}

// ----------------------------------------------------------------------------- root of the category-major kernel
// One CTA per locus: term = sum_cat rw_cat * rootdot[cat][site] (core_likelihood.c:179-196), log, pattern
// weight, fixed-order block reduction.  The locus' value goes into the first of its tile partials (the
// others are zeroed) so that finish_kernel sums the batch exactly as for the other kernels.
__global__ void __launch_bounds__(128)
root20_kernel(const LocusDev * __restrict__ loci, const unsigned int * __restrict__ batch_locus,
              const double * __restrict__ rootdot, const unsigned long long * __restrict__ site_off,
              const unsigned int * __restrict__ tile_first, double * __restrict__ tile_partial,
              double * __restrict__ persite, int persite_mode)
{
  __shared__ double s_red[4];
  const unsigned int bl = blockIdx.x;
  const LocusDev & L = loci[batch_locus[bl]];
  const unsigned int sites = L.sites, R = L.rate_cats;
  const double * base = rootdot + (size_t)site_off[bl] * R;
  double acc = 0.0;
  for (unsigned int s = threadIdx.x; s < sites; s += blockDim.x)
  {
    double term = 0.0;
    for (unsigned int c = 0; c < R; ++c) term += base[(size_t)c * sites + s] * L.rate_weights[c];
    double v;
    if (persite_mode == 2) v = term;
    else v = log(term) * (double)L.weights[s];
    if (persite) persite[s] = v;
    acc += v;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
  if ((threadIdx.x & 31u) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    tile_partial[tile_first[bl]] = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
    for (unsigned int t = tile_first[bl] + 1; t < tile_first[bl + 1]; ++t) tile_partial[t] = 0.0;
  }
}

}  // namespace bppgpu
