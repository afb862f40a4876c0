// tree_s4.cuh -- the hot kernel: tree-fused Felsenstein pruning for 4 states.
//
// One persistent CTA walks a contiguous range of tiles; a tile is TREE_NT cells
// (cell = pattern*RL + cat, RL = rate categories, a power of two <= 8 so that the RL lanes of a site
// sit in one warp) of one locus.  For its tile a thread executes the locus' WHOLE planned op list as
// a stack machine over TRANSFORMED vectors X = P_edge . clv:
//   - a packed tip child (4 bits per tip and site, fetched one tile ahead, kept in registers) is one
//     shared-memory lookup  X = LUT_edge[mask]  (16 masks x 4 doubles per edge and category, built
//     per staged chunk from the edge's P-matrix);
//   - an inner child produced earlier in the list is the register-resident X of the previous op or a
//     shared-memory stack slot -- it is never re-read from HBM;
//   - parent = X_a * X_b (4 multiplies), stored once with a 256-bit store, then pushed through the
//     P-matrix of the edge above it (the only 4x4 mat-vec of the op);
//   - per-site rescaling and the root's site log-likelihoods are fused into the same pass.
// The per-locus program (header, ops, P-matrices) is a contiguous block built by plan_kernel_blocks;
// the block of the NEXT locus is copied with cp.async into the second stage buffer while the current
// tile computes, and tile descriptors run two tiles ahead in a small ring.
// HBM traffic per locus is the compulsory (T-1) CLV writes + packed tips + weights + the block
// (SURVEY.md 8d "B_min").
//
// Arithmetic (reference file:line, /root/reference/src):
//   x_i = (P_i0 c0 + P_i1 c1) + (P_i2 c2 + P_i3 c3), separate mul/add  core_partials_avx.c:423-473
//   parent_i = x_i * y_i                                               :476
//   site rescaling: all 4*R entries < 2^-256 (strict, unscaled)        :493-529, core_partials.c:720-754
//   root: sum_j rw_j ((pi0 c0 + pi1 c1) + (pi2 c2 + pi3 c3)), log, + scaler*log(2^-256), * weight
//                                                                      core_likelihood_avx.c:121-150
// In EXACT mode every x_i is produced by exactly these operations in exactly this order (the lookup
// table rows are computed with real multiplications by 0.0 / 1.0), so CLVs are bit-identical to the
// reference's AVX kernels given bit-identical P-matrices.
#pragma once
#include "common.cuh"

namespace bppgpu {

// all views alias the dynamic shared memory; indexing them with integers (no generic pointers)
// keeps the address arithmetic out of the instruction stream
extern __shared__ uint4 s4[];
extern __shared__ double s8[];
extern __shared__ unsigned int s1[];

struct Vec4 { double a, b, c, d; };

__device__ __forceinline__ double2 as_d2(const uint4 v)
{
  return make_double2(__hiloint2double((int)v.y, (int)v.x), __hiloint2double((int)v.w, (int)v.z));
}
__device__ __forceinline__ uint4 as_u4(const double x, const double y)
{
  return make_uint4((unsigned)__double2loint(x), (unsigned)__double2hiint(x), (unsigned)__double2loint(y),
                    (unsigned)__double2hiint(y));
}

template <bool EXACT>
__device__ __forceinline__ double dot4(const double2 pa, const double2 pb, const double c0, const double c1,
                                       const double c2, const double c3)
{
  if (EXACT)
    return __dadd_rn(__dadd_rn(__dmul_rn(pa.x, c0), __dmul_rn(pa.y, c1)),
                     __dadd_rn(__dmul_rn(pb.x, c2), __dmul_rn(pb.y, c3)));
  return fma(pa.x, c0, pa.y * c1) + fma(pb.x, c2, pb.y * c3);
}

// X = P . v, P = 16 doubles (row-major) at uint4 index p of shared memory
template <bool EXACT>
__device__ __forceinline__ Vec4 matvec_s4(const unsigned int p, const double v0, const double v1, const double v2,
                                          const double v3)
{
  Vec4 x;
  x.a = dot4<EXACT>(as_d2(s4[p + 0]), as_d2(s4[p + 1]), v0, v1, v2, v3);
  x.b = dot4<EXACT>(as_d2(s4[p + 2]), as_d2(s4[p + 3]), v0, v1, v2, v3);
  x.c = dot4<EXACT>(as_d2(s4[p + 4]), as_d2(s4[p + 5]), v0, v1, v2, v3);
  x.d = dot4<EXACT>(as_d2(s4[p + 6]), as_d2(s4[p + 7]), v0, v1, v2, v3);
  return x;
}

// cold path: child CLV resident in HBM (partial updates, stack overflow) or a dense tip: load it and
// apply the edge's P-matrix straight from global memory (L1-cached).  kind/p0/p1/p2 as in OpRec.
template <int RL, bool EXACT>
__device__ __noinline__ Vec4 fetch_global(const LocusHdr * H, unsigned int kind, unsigned int p0, unsigned int p1,
                                          unsigned int p2, unsigned int cell, unsigned int pattern, unsigned int cat,
                                          unsigned int * sc)
{
  double v0, v1, v2, v3;
  if (kind == SRC_TIP_DENSE) ld256_nc(H->tip_dense + (size_t)p0 * H->clv_stride + (size_t)cell * 4, v0, v1, v2, v3);
  else ld256(H->clv + (size_t)p0 * H->clv_stride + (size_t)cell * 4, v0, v1, v2, v3);
  *sc = (kind == SRC_HBM && (int)p2 >= 0) ? H->scale[(size_t)p2 * H->sites + pattern] : 0u;
  const double2 * __restrict__ p = reinterpret_cast<const double2 *>(H->pmat + ((size_t)p1 * RL + cat) * 16);
  Vec4 x;
  x.a = dot4<EXACT>(__ldg(p + 0), __ldg(p + 1), v0, v1, v2, v3);
  x.b = dot4<EXACT>(__ldg(p + 2), __ldg(p + 3), v0, v1, v2, v3);
  x.c = dot4<EXACT>(__ldg(p + 4), __ldg(p + 5), v0, v1, v2, v3);
  x.d = dot4<EXACT>(__ldg(p + 6), __ldg(p + 7), v0, v1, v2, v3);
  return x;
}

// cold path: tip word beyond the two register-resident ones (more than 16 tips)
__device__ __noinline__ unsigned int fetch_tipword(const LocusHdr * H, unsigned int pattern, unsigned int wi)
{
  return __ldg(H->tipwords + (size_t)pattern * H->tip_words + wi);
}

template <int RL>
struct S4Layout               // everything in uint4 (16-byte) units
{
  static constexpr unsigned RW16 = (unsigned)((((size_t)RL * 8 + 15) & ~(size_t)15) / 16);
  static constexpr unsigned CAP = (unsigned)lut_cap(RL);
  static constexpr unsigned HDR = 0;
  static constexpr unsigned RW = 8;
  static constexpr unsigned CH = RW + RW16;                       // ChunkHdr
  static constexpr unsigned OPS = CH + 1;                         // OpRec[TREE_CHUNK]
  static constexpr unsigned PUP = OPS + TREE_CHUNK * 4;           // [TREE_CHUNK][RL] x 9
  static constexpr unsigned TIPP = PUP + TREE_CHUNK * RL * 9;     // [CAP][RL] x 9
  static constexpr unsigned STAGE = TIPP + CAP * RL * 9;          // one stage buffer
  static constexpr unsigned CHUNK = STAGE - CH;                   // chunk size
  static constexpr unsigned LUT = 2 * STAGE;                      // [CAP][RL] x 49
  static constexpr unsigned RING = LUT + CAP * RL * 49;           // 4 x (TileDesc 2 + blk 1)
  static constexpr unsigned RED = RING + 12;                      // 32 doubles
  static constexpr unsigned STACK = RED + 16;                     // [slots][2][TREE_NT] uint4, then [slots][TREE_NT] u32
  __host__ __device__ static constexpr size_t bytes(int slots)
  {
    return (size_t)STACK * 16 + (size_t)slots * (2 * TREE_NT * 16 + TREE_NT * 4);
  }
};

template <int RL, bool EXACT>
__global__ void __launch_bounds__(TREE_NT, 3)
tree_kernel_s4(const TreeParams prm)
{
  using Lay = S4Layout<RL>;
  const unsigned int tid = threadIdx.x, lane = tid & 31u;
  const unsigned int SST1 = (Lay::STACK + (unsigned)prm.n_slots * 2 * TREE_NT) * 4;      // u32 index of the scaler stack

  const unsigned int t_begin = (unsigned int)(((unsigned long long)prm.n_tiles * blockIdx.x) / gridDim.x);
  const unsigned int t_end = (unsigned int)(((unsigned long long)prm.n_tiles * (blockIdx.x + 1)) / gridDim.x);
  if (t_begin >= t_end) return;

  auto ring_fetch = [&](unsigned int t)      // descriptor + block offset of tile t -> ring slot t&3 (3 x 16 B)
  {
    if (tid < 3 && t < t_end)
    {
      const unsigned char * src = tid < 2 ? reinterpret_cast<const unsigned char *>(prm.tiles + t) + tid * 16
                                          : reinterpret_cast<const unsigned char *>(prm.tile_blk + 2 * (size_t)t);
      cp_async16(&s4[Lay::RING + (t & 3u) * 3 + tid], src);
    }
  };
  auto stage_fetch = [&](unsigned int buf, unsigned long long blk)   // header + rate weights + chunk 0
  {
    const uint4 * src = reinterpret_cast<const uint4 *>(prm.blocks + blk);
    for (unsigned int w = tid; w < Lay::STAGE; w += TREE_NT) cp_async16(&s4[buf * Lay::STAGE + w], src + w);
  };

  ring_fetch(t_begin);
  ring_fetch(t_begin + 1);
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();

  // tips / weight of the first tile
  unsigned int tw0 = 0, tw1 = 0, wgt = 0;
  {
    const TileDesc d = *reinterpret_cast<const TileDesc *>(&s4[Lay::RING + (t_begin & 3u) * 3]);
    const unsigned int craw = d.cell0 + tid;
    const unsigned int pat = (craw < d.ncell ? craw : d.ncell - 1) / RL;
    tw0 = __ldg(d.tipwords + (size_t)pat * d.tip_words);
    if (d.tip_words > 1) tw1 = __ldg(d.tipwords + (size_t)pat * d.tip_words + 1);
    wgt = __ldg(d.weights + pat);
  }
  unsigned int buf = 0;
  unsigned int staged_locus = 0xFFFFFFFFu;      // locus whose header + chunk 0 + LUT are valid in stage[buf]
  unsigned int prefetched_locus = 0xFFFFFFFFu;  // locus whose block is (being) copied into stage[buf ^ 1]

  for (unsigned int t = t_begin; t < t_end; ++t)
  {
    const unsigned int rs = Lay::RING + (t & 3u) * 3;
    const TileDesc d = *reinterpret_cast<const TileDesc *>(&s4[rs]);
    const unsigned long long blk = *reinterpret_cast<const unsigned long long *>(&s4[rs + 2]);
    bool build_lut = false;
    if (d.locus != staged_locus)
    {
      if (d.locus == prefetched_locus) buf ^= 1u;          // landed: wait_all + barrier at the end of the last tile
      else
      {
        stage_fetch(buf, blk);
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
      }
      prefetched_locus = 0xFFFFFFFFu;
      build_lut = true;
    }
    // ---- prefetch: tips/weight of tile t+1 -> registers; block of the next locus -> other stage buffer;
    //      descriptor of tile t+2 -> ring
    unsigned int ntw0 = 0, ntw1 = 0, nwgt = 0;
    if (t + 1 < t_end)
    {
      const unsigned int rn = Lay::RING + ((t + 1) & 3u) * 3;
      const TileDesc dn = *reinterpret_cast<const TileDesc *>(&s4[rn]);
      const unsigned int craw = dn.cell0 + tid;
      const unsigned int pat = (craw < dn.ncell ? craw : dn.ncell - 1) / RL;
      ntw0 = __ldg(dn.tipwords + (size_t)pat * dn.tip_words);
      if (dn.tip_words > 1) ntw1 = __ldg(dn.tipwords + (size_t)pat * dn.tip_words + 1);
      nwgt = __ldg(dn.weights + pat);
      if (dn.locus != d.locus && dn.locus != prefetched_locus)
      {
        const unsigned long long nblk = *reinterpret_cast<const unsigned long long *>(&s4[rn + 2]);
        stage_fetch(buf ^ 1u, nblk);
        prefetched_locus = dn.locus;
      }
    }
    ring_fetch(t + 2);
    cp_async_commit();

    const unsigned int sb = buf * Lay::STAGE;
    const LocusHdr * H = reinterpret_cast<const LocusHdr *>(&s4[sb]);

    const unsigned int ncell = d.ncell;
    const unsigned int cell_raw = d.cell0 + tid;
    const bool valid = cell_raw < ncell;
    const unsigned int cell = valid ? cell_raw : ncell - 1;
    const unsigned int pattern = cell / RL;
    const unsigned int cat = cell % RL;
    const unsigned int lut_t = Lay::LUT + cat * 49;
    const unsigned int stk_t = Lay::STACK + tid;

    double x0 = 0, x1 = 0, x2 = 0, x3 = 0;    // X of the previous op's result
    unsigned int psc = 0;
    double site_val = 0.0;
    const unsigned int n_chunks = H->n_chunks;
    unsigned char * const clv_cell = reinterpret_cast<unsigned char *>(H->clv) + ((size_t)cell << 5);

    for (unsigned int c = 0; c < n_chunks; ++c)
    {
      if (c > 0)
      {
        // trees with more ops than one chunk holds: later chunks are staged in place, synchronously
        __syncthreads();
        const uint4 * src = reinterpret_cast<const uint4 *>(prm.blocks + blk) + Lay::CH + (size_t)c * Lay::CHUNK;
        for (unsigned int w = tid; w < Lay::CHUNK; w += TREE_NT) s4[sb + Lay::CH + w] = __ldg(src + w);
        __syncthreads();
        build_lut = true;
      }
      if (build_lut)
      {
        // LUT[s][cat][mask] = P_tip-edge . bits(mask), same operation order as the mat-vec
        const unsigned int entries = s1[(sb + Lay::CH) * 4 + 1] * RL * 16;       // ChunkHdr.ntips
        for (unsigned int e = tid; e < entries; e += TREE_NT)
        {
          const unsigned int mask = e & 15u, sc = e >> 4;      // sc = slot*RL + cat
          const Vec4 v = matvec_s4<EXACT>(sb + Lay::TIPP + sc * 9, (double)(mask & 1u), (double)((mask >> 1) & 1u),
                                          (double)((mask >> 2) & 1u), (double)((mask >> 3) & 1u));
          s4[Lay::LUT + sc * 49 + mask * 3] = as_u4(v.a, v.b);
          s4[Lay::LUT + sc * 49 + mask * 3 + 1] = as_u4(v.c, v.d);
        }
        __syncthreads();
        build_lut = false;
        staged_locus = (c == 0) ? d.locus : 0xFFFFFFFFu;     // a later chunk overwrote chunk 0
      }
      const unsigned int cn = s1[(sb + Lay::CH) * 4];         // ChunkHdr.nops
      const unsigned int ops = sb + Lay::OPS;
      const unsigned int pup_t = sb + Lay::PUP + cat * 9;

      for (unsigned int k = 0; k < cn; ++k)
      {
        const uint4 w0 = s4[ops + 4 * k], w1 = s4[ops + 4 * k + 1], w2 = s4[ops + 4 * k + 2];
        const unsigned int ctl = w0.x;
        if (ctl & OP_PARK)
        {
          const unsigned int ps = s1[(ops + 4 * k + 3) * 4];
          s4[stk_t + ps * (2 * TREE_NT)] = as_u4(x0, x1);
          s4[stk_t + ps * (2 * TREE_NT) + TREE_NT] = as_u4(x2, x3);
          s1[SST1 + ps * TREE_NT + tid] = psc;
        }
        // X of an operand that is not the register-resident previous result
        auto fetch = [&](unsigned int kind, unsigned int p0, unsigned int p1, unsigned int p2, unsigned int & sc) -> Vec4
        {
          Vec4 r;
          if (kind == SRC_TIP_PACKED)
          {
            unsigned int word = p0 ? tw1 : tw0;
            if (p0 >= 2) word = fetch_tipword(H, pattern, p0);
            const unsigned int mask = (word >> p1) & 15u;
            const unsigned int li = lut_t + p2 + mask * 3;
            const double2 u = as_d2(s4[li]), w = as_d2(s4[li + 1]);
            r.a = u.x; r.b = u.y; r.c = w.x; r.d = w.y; sc = 0;
          }
          else if (kind == SRC_SLOT)
          {
            const double2 u = as_d2(s4[stk_t + p0 * (2 * TREE_NT)]), w = as_d2(s4[stk_t + p0 * (2 * TREE_NT) + TREE_NT]);
            r.a = u.x; r.b = u.y; r.c = w.x; r.d = w.y; sc = s1[SST1 + p0 * TREE_NT + tid];
          }
          else r = fetch_global<RL, EXACT>(H, kind, p0, p1, p2, cell, pattern, cat, &sc);
          return r;
        };

        double o0, o1, o2, o3;
        unsigned int osc = 0;
        if (ctl & OP_EVAL)
        {
          // root CLV that this list did not produce: read it (no P applied)
          const unsigned int kind = w0.z, p0 = w0.w;
          if (kind == SRC_TIP_PACKED)
          {
            unsigned int word = p0 ? tw1 : tw0;
            if (p0 >= 2) word = fetch_tipword(H, pattern, p0);
            const unsigned int mask = (word >> w1.x) & 15u;
            o0 = (double)(mask & 1u); o1 = (double)((mask >> 1) & 1u); o2 = (double)((mask >> 2) & 1u); o3 = (double)((mask >> 3) & 1u);
          }
          else if (kind == SRC_TIP_DENSE) ld256_nc(H->tip_dense + (size_t)p0 * H->clv_stride + (size_t)cell * 4, o0, o1, o2, o3);
          else
          {
            ld256(H->clv + (size_t)p0 * H->clv_stride + (size_t)cell * 4, o0, o1, o2, o3);
            if ((int)w1.y >= 0) osc = H->scale[(size_t)w1.y * H->sites + pattern];
          }
        }
        else
        {
          unsigned int asc = 0, bsc = 0;
          const Vec4 a = fetch(w0.z, w0.w, w1.x, w1.y, asc);
          if (ctl & OP_BPREV)
          {
            o0 = __dmul_rn(x0, a.a); o1 = __dmul_rn(x1, a.b); o2 = __dmul_rn(x2, a.c); o3 = __dmul_rn(x3, a.d);
            bsc = psc;
          }
          else
          {
            const Vec4 b = fetch(w1.z, w1.w, w2.x, w2.y, bsc);
            o0 = __dmul_rn(a.a, b.a); o1 = __dmul_rn(a.b, b.b); o2 = __dmul_rn(a.c, b.c); o3 = __dmul_rn(a.d, b.d);
          }
          // ---- per-site scaling (core_partials.c:720,739-754)
          if (ctl & OP_SCALE)
          {
            osc = asc + bsc;
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
            if (valid && cat == 0) H->scale[(size_t)(int)w2.z * H->sites + pattern] = osc;
          }
          if (valid) st256(reinterpret_cast<double *>(clv_cell + ((size_t)w0.y << 5)), o0, o1, o2, o3);
          // ---- push through the edge above: this X is what the parent's op consumes
          if (ctl & OP_PUSH)
          {
            const Vec4 x = matvec_s4<EXACT>(pup_t + k * (RL * 9), o0, o1, o2, o3);
            x0 = x.a; x1 = x.b; x2 = x.c; x3 = x.d; psc = osc;
          }
        }

        if (ctl & OP_ROOT)
        {
          const double tr = __dadd_rn(__dadd_rn(__dmul_rn(H->freqs[0], o0), __dmul_rn(H->freqs[1], o1)),
                                      __dadd_rn(__dmul_rn(H->freqs[2], o2), __dmul_rn(H->freqs[3], o3)));
          double term = 0.0;
#pragma unroll
          for (int j = 0; j < RL; ++j)
          {
            const double v = __shfl_sync(0xFFFFFFFFu, tr, (lane & ~(unsigned)(RL - 1)) + j);
            term = __dadd_rn(term, __dmul_rn(v, s8[(sb + Lay::RW) * 2 + j]));
          }
          unsigned int rsc = osc;
          if (ctl & OP_EVAL) rsc = ((int)w2.w >= 0) ? osc : 0;
          double s;
          if (prm.persite_mode == 2) s = term;
          else
          {
            s = log(term);
            if (rsc) s = __dadd_rn(s, __dmul_rn((double)rsc, prm.log_threshold));
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
      if (lane == 0) s8[Lay::RED * 2 + (t & 1u) * 16 + (tid >> 5)] = v;
    }
    cp_async_wait_all();
    __syncthreads();          // reduction inputs complete; prefetched block and descriptor t+2 visible
    if (prm.tile_partial && tid == 0)
    {
      double acc = 0.0;
      for (unsigned int w = 0; w < (TREE_NT >> 5); ++w) acc += s8[Lay::RED * 2 + (t & 1u) * 16 + w];
      prm.tile_partial[t] = acc;
    }
    tw0 = ntw0; tw1 = ntw1; wgt = nwgt;
  }
}

}  // namespace bppgpu
