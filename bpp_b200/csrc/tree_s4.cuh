// tree_s4.cuh -- the hot kernel: tree-fused Felsenstein pruning for 4 states.
//
// One persistent CTA walks a contiguous range of tiles; a tile is TREE_NT*CPT cells
// (cell = pattern*RL + cat, RL = rate categories, a power of two <= 8 so that the RL lanes of a site
// sit in one warp) of one locus; thread tid owns the CPT cells cell0 + tid + j*TREE_NT (same category).
// For its cells a thread executes the locus' WHOLE planned op list as a stack machine over
// TRANSFORMED vectors X = P_edge . clv:
//   - a packed tip child (4 bits per tip and site, fetched one tile ahead, kept in registers) is one
//     shared-memory lookup  X = LUT_edge[mask]  (16 masks x 4 doubles per edge and category, built
//     per staged chunk from the edge's P-matrix);
//   - an inner child produced earlier in the list is the register-resident X of the previous op or a
//     shared-memory stack slot -- it is never re-read from HBM;
//   - parent = X_a * X_b (4 multiplies), stored once with a 256-bit store, then pushed through the
//     P-matrix of the edge above it (the only 4x4 mat-vec of the op; its 16 shared-memory words are
//     loaded once for the thread's CPT cells);
//   - per-site rescaling and the root's site log-likelihoods are fused into the same pass.
// The per-locus program (header, ops, P-matrices) is a contiguous block built by plan_kernel_blocks;
// the block of the NEXT locus is copied with cp.async into the second stage buffer while the current
// tile computes, and tile descriptors run two tiles ahead in a small ring.
// Loci whose ops all use fast operands (HDR_FAST) run tile_fast; anything else (HBM-resident
// children of partial updates, dense tips, more than 16 tips, root-only evaluation) runs tile_general.
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

// all views alias the dynamic shared memory
extern __shared__ uint4 s4[];
extern __shared__ double2 sd2[];
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

// cell offset (within the CTA's TREE_NT cells of one j) owned by thread tid: lanes of a warp are ordered category-major
#ifndef BPPGPU_S4_PERM
#define BPPGPU_S4_PERM 0
#endif
#ifndef BPPGPU_S4_TMA_STAGE
#define BPPGPU_S4_TMA_STAGE 1        // locus blocks arrive by one TMA bulk copy (0: 16-byte cp.async by all threads)
#endif
// lanes of a site: with the identity mapping (cell = cell0 + tid) the RL lanes of a site are neighbours; the
// category-major permutation (BPPGPU_S4_PERM=1, measured and rejected: every quarter warp then writes eight
// 32-byte pieces of eight different 128-byte lines per store instead of 256 contiguous bytes, and the store path
// is transaction-bound: config 3 went from 3.84 to 4.82 ms) puts them 32/RL lanes apart
template <int RL> struct S4Lanes
{
  static constexpr unsigned int SPW = 32u / RL;       // sites per warp
  __device__ static __forceinline__ unsigned int perm(const unsigned int tid)
  {
#if BPPGPU_S4_PERM
    const unsigned int lane = tid & 31u;
    return (tid & ~31u) | ((lane % SPW) * RL + lane / SPW);
#else
    return tid;
#endif
  }
  // lane of category r of this lane's site; xor distance between the lanes of a site for step dd = 1, 2, 4
  __device__ static __forceinline__ unsigned int cat_lane(const unsigned int lane, const unsigned int r)
  {
#if BPPGPU_S4_PERM
    return (lane % SPW) + r * SPW;
#else
    return (lane & ~(unsigned int)(RL - 1)) + r;
#endif
  }
  __device__ static __forceinline__ unsigned int xor_step(const unsigned int dd)
  {
#if BPPGPU_S4_PERM
    return dd * SPW;
#else
    return dd;
#endif
  }
};
template <int RL>
__device__ __forceinline__ unsigned int s4_perm(const unsigned int tid) { return S4Lanes<RL>::perm(tid); }

template <int RL, int CPT>
struct S4Layout               // everything in uint4 (16-byte) units
{
  static constexpr unsigned RW16 = (unsigned)((((size_t)RL * 8 + 15) & ~(size_t)15) / 16);
  static constexpr unsigned CAP = (unsigned)lut_cap(RL);
  static constexpr unsigned RW = 8;
  static constexpr unsigned CH = RW + RW16;                       // ChunkHdr
  static constexpr unsigned OPS = CH + 1;                         // OpRec[TREE_CHUNK]
  static constexpr unsigned PUP = OPS + TREE_CHUNK * 4;           // [TREE_CHUNK][RL] x 9
  static constexpr unsigned TIPP = PUP + TREE_CHUNK * RL * 9;     // [CAP][RL] x 9
  static constexpr unsigned STAGE = TIPP + CAP * RL * 9;          // one stage buffer
  static constexpr unsigned CHUNK = STAGE - CH;                   // chunk size
#ifndef BPPGPU_S4_NSTAGE4
#define BPPGPU_S4_NSTAGE4 2
#endif
  static constexpr unsigned NSTAGE = RL >= 8 ? 1 : (RL == 4 ? BPPGPU_S4_NSTAGE4 : 2);   // double-buffered staging (the next locus' block
                                                                  // is copied while the current one computes) where shared memory allows
  // shared memory of the kernel: [RING 4 x (TileDesc 2 + blk 1)][RED 32 doubles][NSTAGE stage buffers]
  // [LUT cap x lut_slot_u4][STACK [slots][CPT*TREE_NT][2] uint4, then [slots][CPT*TREE_NT] u32].  cap is the
  // launch's tip-slot capacity (<= CAP, sized to the batch's largest tree): a stage buffer holds only
  // the first cap slots of the block's tipP area, which is why tipP comes last in the block.
  static constexpr unsigned RING = 0;
  static constexpr unsigned RED = 12;
  static constexpr unsigned MBAR = 28;                            // two mbarriers (one per stage buffer)
  static constexpr unsigned STAGE0 = 29;
  static constexpr unsigned SLOT = 2 * CPT * TREE_NT;             // uint4 per slot
  __host__ __device__ static constexpr unsigned stage_sz(unsigned cap) { return TIPP + cap * RL * 9; }
  __host__ __device__ static constexpr unsigned lut0(unsigned cap) { return STAGE0 + NSTAGE * stage_sz(cap); }
  __host__ __device__ static constexpr unsigned stack0(unsigned cap) { return lut0(cap) + cap * lut_slot_u4(RL); }
  // after the stack: packed tip words and pattern weights of the current and the next tile, one entry per SITE (the
  // RL lanes of a site read the same word: a broadcast),
  // [2 buffers][W tip words, then the weight][CPT][SROW] u32 (W = the launch's tip_words), filled by 4-byte cp.async
  static constexpr unsigned SROW = TREE_NT / RL;                  // sites per row of TREE_NT cells
  __host__ __device__ static constexpr unsigned tips_buf(unsigned words) { return (words + 1) * CPT * SROW; }   // u32 per buffer
  __host__ __device__ static constexpr unsigned tips0_u32(int slots, unsigned cap)
  {
    return (stack0(cap) + (unsigned)slots * SLOT) * 4 + (unsigned)slots * CPT * TREE_NT;
  }
  __host__ __device__ static constexpr size_t bytes(int slots, unsigned cap, unsigned words)
  {
    return (size_t)tips0_u32(slots, cap) * 4 + 2 * (size_t)tips_buf(words) * 4;
  }
};

// everything a tile function needs to know about the thread's cells
template <int CPT>
struct TileCtx
{
  unsigned int sb;               // stage buffer base (uint4 units)
  unsigned int lut0, stack0;     // lookup tables and stack (uint4 units)
  unsigned int sst1;             // u32 index of the scaler stack
  unsigned int cat;
  unsigned int cell[CPT];        // clamped cell index
  bool valid[CPT];
  unsigned int tw0[CPT];         // tip word 0 (tips 0..7) of the thread's cells
  unsigned int tips_s;           // u32 index of this thread's site in the tile's tip buffer: word w of cell j at
                                 // tips_s + (w * CPT + j) * SROW, pattern weight at tips_s + wgt_off + j * SROW (SROW = TREE_NT / RL)
  unsigned int wgt_off;          // = tip_words * CPT * SROW
  unsigned int cell0, ncell;     // the tile's first cell and the locus' cell count
};

// ------------------------------------------------------------------------------------------------
// fast path: every operand is a register-resident packed tip, a stack slot or the previous X
// ------------------------------------------------------------------------------------------------
// MODE 0 is the lean instantiation for loci flagged HDR_SIMPLE (no HBM-class operand, no scaler): full-tree
// passes without scaling, the headline workload.  MODE 1 adds per-site scaling (HDR_NOHBM: full passes with
// scale buffers), MODE 2 also HBM-class operands (partial updates).
// One call runs the ops of the chunk that is staged; x / psc (the register X and its scaler count) carry over
// from chunk to chunk of the same tile.
// MODE 3 is the SPECULATIVE form of MODE 1 (specialised scaled launches, SCALED_ONLY): per-site rescaling is rare (all
// 4 x RL entries of a site below 2^-256), so the pass runs like MODE 0 -- no scaler counts, no scaler stores, no
// per-op lane exchange -- and only records, one bit per (op, cell), where an op's values fell below the threshold.
// The caller combines the bits over the lanes of a site once per tile; if no site qualified (the common case) the
// scalers of the tile are zero-filled in one coalesced sweep, otherwise the warp repeats the tile in MODE 1.
template <int RL, bool EXACT, int CPT, int MODE>
__device__ __forceinline__ double tile_fast(const TreeParams & prm, const TileCtx<CPT> & tc, double (&x)[CPT][4],
                                            unsigned int (&psc)[CPT], unsigned int & wnz, unsigned long long & spec)
{
  constexpr bool SCALED = MODE == 1 || MODE == 2, FULL = MODE == 2, SPEC = MODE == 3;
  static_assert(!SPEC || TREE_CHUNK * CPT <= 64, "one bit per (op, cell) of a chunk");
  constexpr unsigned int LOG2RL = RL == 1 ? 0 : (RL == 2 ? 1 : (RL == 4 ? 2 : 3));
  using Lay = S4Layout<RL, CPT>;
  const unsigned int tid = threadIdx.x, lane = tid & 31u;
  const unsigned int sb = tc.sb;
  const LocusHdr * H = reinterpret_cast<const LocusHdr *>(&s4[sb]);
  // the two halves of an entry (states 0,1 / states 2,3) are loaded in the order that keeps a quarter warp's accesses
  // on disjoint banks: lanes with bit 2 set fetch the upper half first -- from replica B of the lookup tables (RL >= 4,
  // see lut_slot_u4) and from their own stack slots, which they fill in the same exchanged order
  // Shared-memory operands are addressed by BYTE offsets from the start of the dynamic shared memory, built from
  // per-thread invariants and per-op (warp-uniform) terms, so that a lookup costs shift / and / multiply-add per cell.
  constexpr unsigned int ROWB = lut_row_u4(RL) * 16u, SLOTJ = 2u * TREE_NT * 16u;
  const unsigned char * const sm = reinterpret_cast<const unsigned char *>(s4);
  static_assert(!(RL == 8 && BPPGPU_S4_PERM), "8 categories: lane bit 2 must be category bit 2");
  const unsigned int hs = (lane >> 2) & 1u, hl = RL >= 4 ? hs : 0u;
  const unsigned int lut_lo = (tc.lut0 + tc.cat * lut_cat_u4(RL) + hl * lut_rep_u4(RL) + hl) * 16u;   // first-fetched half
  const unsigned int stk_lo = (tc.stack0 + tid * 2 + hs) * 16u;
  const int lut_d = hl ? -16 : 16, stk_d = hs ? -16 : 16;                                        // ... to the other half
  const unsigned int pup_t = sb + Lay::PUP + tc.cat * 9;
  const unsigned int tipp_t = sb + Lay::TIPP + tc.cat * 9;
  const unsigned int ops = sb + Lay::OPS;
  const unsigned int cn = s1[(sb + Lay::CH) * 4];
  unsigned char * const clv0 = reinterpret_cast<unsigned char *>(H->clv);
  const unsigned int sites = H->sites;
  unsigned char * pc[CPT];                       // this thread's cells in CLV buffer 0
#pragma unroll
  for (int j = 0; j < CPT; ++j) pc[j] = clv0 + ((size_t)tc.cell[j] << 5);

  double site_sum = 0.0;

  // Partial updates (root paths of gene-tree moves): nearly every op has a sibling CLV that lives in HBM, and a
  // thread would meet those loads one after the other, a DRAM round trip per op with a handful of warps per SM to
  // hide it.  All of the chunk's HBM-resident operands are requested into L2 up front (no registers involved), so
  // the loads of ops 2..n overlap the first one's latency.
#ifndef BPPGPU_S4_L2PREFETCH
#define BPPGPU_S4_L2PREFETCH 1
#endif
  if (FULL && BPPGPU_S4_L2PREFETCH)
    for (unsigned int k = 0; k < cn; ++k)
    {
      const unsigned int ctl = s1[(ops + 4 * k) * 4];
      const unsigned int akind = (ctl >> OP_AKIND_SHIFT) & 15u, bkind = (ctl >> OP_BKIND_SHIFT) & 15u;
      if (akind == SRC_HBM)
      {
        const size_t off = (size_t)(s1[(ops + 4 * k + 2) * 4] * (sites * RL)) << 5;
#pragma unroll
        for (int j = 0; j < CPT; ++j) prefetch_l2(pc[j] + off);
      }
      if (bkind == SRC_HBM && !(ctl & OP_BPREV))
      {
        const size_t off = (size_t)(s1[(ops + 4 * k + 3) * 4] * (sites * RL)) << 5;
#pragma unroll
        for (int j = 0; j < CPT; ++j) prefetch_l2(pc[j] + off);
      }
    }

  for (unsigned int k = 0; k < cn; ++k)
  {
    const uint4 w0 = s4[ops + 4 * k], w1 = s4[ops + 4 * k + 1];
    const unsigned int ctl = w0.x;
    double o[CPT][4];
    unsigned int osc[CPT];
    // ---- operand A (tip lookup or stack slot), then the product with B
    {
      const unsigned int amask = w0.z & 15u, ash = (w0.z >> 8) & 31u;
      const unsigned int aword = (w0.z >> 4) & 15u;
      const unsigned int bmask = w1.x & 15u, bsh = (w1.x >> 8) & 31u;
      const unsigned int bword = (w1.x >> 4) & 15u;
      // HBM-class operands: a CLV re-read from global memory (written earlier in this pass by this very
      // thread, or resident from an earlier call) times its edge's staged P-matrix: the producer's Pup
      // (SRC_HBML) or a tipP slot (SRC_HBM)
      const unsigned int akind = (ctl >> OP_AKIND_SHIFT) & 15u, bkind = (ctl >> OP_BKIND_SHIFT) & 15u;
      const bool a_hbm = FULL && (akind == SRC_HBML || akind == SRC_HBM), b_hbm = FULL && (bkind == SRC_HBML || bkind == SRC_HBM);
      double2 a0[CPT], a1[CPT];
      if (a_hbm)
      {
        const size_t a_off = (size_t)(s1[(ops + 4 * k + 2) * 4] * (sites * RL)) << 5;       // a_p0 = buffer index
        const unsigned int a_pi = (akind == SRC_HBML ? pup_t : tipp_t) + w0.w * (RL * 9);
#pragma unroll
        for (int j = 0; j < CPT; ++j)
        {
          double v0, v1, v2, v3;
          ld256(reinterpret_cast<const double *>(pc[j] + a_off), v0, v1, v2, v3);
          a0[j].x = dot4<EXACT>(sd2[a_pi + 0], sd2[a_pi + 1], v0, v1, v2, v3); a0[j].y = dot4<EXACT>(sd2[a_pi + 2], sd2[a_pi + 3], v0, v1, v2, v3);
          a1[j].x = dot4<EXACT>(sd2[a_pi + 4], sd2[a_pi + 5], v0, v1, v2, v3); a1[j].y = dot4<EXACT>(sd2[a_pi + 6], sd2[a_pi + 7], v0, v1, v2, v3);
        }
      }
      else
      {
        // warp-uniform per op: table or stack, where the other half lies, how far apart the thread's cells are
        const unsigned int base = (amask ? lut_lo : stk_lo) + w0.w * 16u, jst = amask ? 0u : SLOTJ;
        const int dh = amask ? lut_d : stk_d;
#pragma unroll
        for (int j = 0; j < CPT; ++j)
        {
          const unsigned int wa = aword ? s1[tc.tips_s + (aword * CPT + j) * Lay::SROW] : tc.tw0[j];
          const unsigned int at = base + ((wa >> ash) & amask) * ROWB + j * jst;
          a0[j] = *reinterpret_cast<const double2 *>(sm + at);
          a1[j] = *reinterpret_cast<const double2 *>(sm + at + dh);
        }
      }
      if (!(ctl & OP_BPREV))
      {
        // operand B is not the register X: load it INTO the X registers (they are dead: a pushed X
        // is consumed by exactly one op, and that op has OP_BPREV)
        if (b_hbm)
        {
          const size_t b_off = (size_t)(s1[(ops + 4 * k + 3) * 4] * (sites * RL)) << 5;     // b_p0
          const unsigned int b_pi = (bkind == SRC_HBML ? pup_t : tipp_t) + w1.y * (RL * 9);
#pragma unroll
          for (int j = 0; j < CPT; ++j)
          {
            double v0, v1, v2, v3;
            ld256(reinterpret_cast<const double *>(pc[j] + b_off), v0, v1, v2, v3);
            x[j][0] = dot4<EXACT>(sd2[b_pi + 0], sd2[b_pi + 1], v0, v1, v2, v3); x[j][1] = dot4<EXACT>(sd2[b_pi + 2], sd2[b_pi + 3], v0, v1, v2, v3);
            x[j][2] = dot4<EXACT>(sd2[b_pi + 4], sd2[b_pi + 5], v0, v1, v2, v3); x[j][3] = dot4<EXACT>(sd2[b_pi + 6], sd2[b_pi + 7], v0, v1, v2, v3);
          }
        }
        else
        {
          const unsigned int base = (bmask ? lut_lo : stk_lo) + w1.y * 16u, jst = bmask ? 0u : SLOTJ;
          const int dh = bmask ? lut_d : stk_d;
#pragma unroll
          for (int j = 0; j < CPT; ++j)
          {
            const unsigned int wb = bword ? s1[tc.tips_s + (bword * CPT + j) * Lay::SROW] : tc.tw0[j];
            const unsigned int at = base + ((wb >> bsh) & bmask) * ROWB + j * jst;
            const double2 b0 = *reinterpret_cast<const double2 *>(sm + at);
            const double2 b1 = *reinterpret_cast<const double2 *>(sm + at + dh);
            x[j][0] = b0.x; x[j][1] = b0.y; x[j][2] = b1.x; x[j][3] = b1.y;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < CPT; ++j)
      {
        o[j][0] = __dmul_rn(x[j][0], a0[j].x); o[j][1] = __dmul_rn(x[j][1], a0[j].y);
        o[j][2] = __dmul_rn(x[j][2], a1[j].x); o[j][3] = __dmul_rn(x[j][3], a1[j].y);
        osc[j] = 0;
      }
    }
    // ---- per-site scaling (core_partials.c:720,739-754): a site is rescaled when all its 4 x RL entries are below
    // 2^-256.  That is rare, and the bookkeeping around it was 60 % more instructions than the unscaled op, so the
    // common case is kept to: one integer compare per cell (entries are non-negative, and 2^-256 has a zero
    // mantissa, so "all four below" is max(high words) < 0x2FF00000 -- off the FP64 pipe), the cells' bits packed
    // into one word and combined over the site's RL lanes once per op, and ONE warp-uniform branch.  Scaler
    // counts of staged children are read only once some count of the warp is non-zero (wnz).
    if (SPEC && (ctl & OP_SCALE))
    {
      unsigned int bits = 0;
#pragma unroll
      for (int j = 0; j < CPT; ++j)
      {
        const int h0 = __double2hiint(o[j][0]), h1 = __double2hiint(o[j][1]), h2 = __double2hiint(o[j][2]), h3 = __double2hiint(o[j][3]);
        if (max(max(h0, h1), max(h2, h3)) < 0x2FF00000) bits |= 1u << j;
      }
      spec |= (unsigned long long)bits << (k * CPT);
    }
    if (SCALED && (ctl & OP_SCALE))
    {
      const uint4 w2 = s4[ops + 4 * k + 2], w3 = s4[ops + 4 * k + 3];
      const unsigned int ak = (ctl >> OP_AKIND_SHIFT) & 15u, bk = (ctl >> OP_BKIND_SHIFT) & 15u;
      const bool a_glob = FULL && (ak == SRC_HBML || ak == SRC_HBM) && (int)w2.z >= 0;
      const bool b_glob = FULL && !(ctl & OP_BPREV) && (bk == SRC_HBML || bk == SRC_HBM) && (int)w3.z >= 0;
      unsigned int bits = 0;
#pragma unroll
      for (int j = 0; j < CPT; ++j)
      {
        const int h0 = __double2hiint(o[j][0]), h1 = __double2hiint(o[j][1]), h2 = __double2hiint(o[j][2]), h3 = __double2hiint(o[j][3]);
        if (max(max(h0, h1), max(h2, h3)) < 0x2FF00000) bits |= 1u << j;
      }
#pragma unroll
      for (int dd = 1; dd < RL; dd <<= 1) bits &= __shfl_xor_sync(0xFFFFFFFFu, bits, S4Lanes<RL>::xor_step(dd));
#pragma unroll
      for (int j = 0; j < CPT; ++j) osc[j] = 0;
      if (wnz | (FULL ? (unsigned)(a_glob || b_glob) : 0u))
      {
        // counts of the children: the previous X, parked values, HBM-resident CLVs of earlier calls
        const bool a_slot = ak == SRC_SLOT, b_slot = !(ctl & OP_BPREV) && bk == SRC_SLOT;
        const unsigned int a_sl = tc.sst1 + w2.x * (CPT * TREE_NT) + tid, b_sl = tc.sst1 + w3.x * (CPT * TREE_NT) + tid;
#pragma unroll
        for (int j = 0; j < CPT; ++j)
        {
          const unsigned int site = tc.cell[j] >> LOG2RL;
          unsigned int sc = (ctl & OP_BPREV) ? psc[j] : 0u;
          if (a_slot) sc += s1[a_sl + j * TREE_NT];
          if (b_slot) sc += s1[b_sl + j * TREE_NT];
          if (a_glob) sc += H->scale[(size_t)(int)w2.z * sites + site];
          if (b_glob) sc += H->scale[(size_t)(int)w3.z * sites + site];
          osc[j] = sc;
        }
        if (FULL)
        {
          unsigned int any = 0;
#pragma unroll
          for (int j = 0; j < CPT; ++j) any |= osc[j];
          wnz |= __any_sync(0xFFFFFFFFu, any != 0u) ? 1u : 0u;
        }
      }
      if (__any_sync(0xFFFFFFFFu, bits != 0u))
      {
        wnz = 1u;
#pragma unroll
        for (int j = 0; j < CPT; ++j)
          if ((bits >> j) & 1u)
          {
            o[j][0] = __dmul_rn(o[j][0], BPPGPU_SCALE_FACTOR); o[j][1] = __dmul_rn(o[j][1], BPPGPU_SCALE_FACTOR);
            o[j][2] = __dmul_rn(o[j][2], BPPGPU_SCALE_FACTOR); o[j][3] = __dmul_rn(o[j][3], BPPGPU_SCALE_FACTOR);
            osc[j] += 1;
          }
      }
      unsigned int * const sc_out = H->scale + (size_t)(int)w1.z * sites;
#pragma unroll
      for (int j = 0; j < CPT; ++j)
        if (tc.valid[j] && tc.cat == 0) sc_out[tc.cell[j] >> LOG2RL] = osc[j];
    }
    // ---- the CLV goes to HBM exactly once
    {
      const size_t d_off = (size_t)w0.y << 5;        // the parent's buffer
#pragma unroll
      for (int j = 0; j < CPT; ++j)
        if (tc.valid[j]) st256(reinterpret_cast<double *>(pc[j] + d_off), o[j][0], o[j][1], o[j][2], o[j][3]);
    }
    // ---- push through the edge above: this X is what the parent's op consumes
    if (ctl & OP_PUSH)
    {
      const unsigned int p = pup_t + k * (RL * 9);
#pragma unroll
      for (int h = 0; h < 4; ++h)
      {
        const double2 pa = sd2[p + 2 * h], pb = sd2[p + 2 * h + 1];
#pragma unroll
        for (int j = 0; j < CPT; ++j) x[j][h] = dot4<EXACT>(pa, pb, o[j][0], o[j][1], o[j][2], o[j][3]);
      }
#pragma unroll
      for (int j = 0; j < CPT; ++j) psc[j] = osc[j];
      if (ctl & OP_PARKA)
      {
        const unsigned int po = w1.w;
#pragma unroll
        for (int j = 0; j < CPT; ++j)
        {
          unsigned char * const dst = const_cast<unsigned char *>(sm) + stk_lo + j * SLOTJ + po * 16u;
          *reinterpret_cast<double2 *>(dst) = make_double2(x[j][0], x[j][1]);
          *reinterpret_cast<double2 *>(dst + stk_d) = make_double2(x[j][2], x[j][3]);
          if (SCALED && (ctl & OP_SCALE)) s1[tc.sst1 + (po / Lay::SLOT) * (CPT * TREE_NT) + j * TREE_NT + tid] = psc[j];
        }
      }
    }
    if (ctl & OP_ROOT)
    {
      const double f0 = H->freqs[0], f1 = H->freqs[1], f2 = H->freqs[2], f3 = H->freqs[3];
      double term[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j)
      {
        const double tr = __dadd_rn(__dadd_rn(__dmul_rn(f0, o[j][0]), __dmul_rn(f1, o[j][1])),
                                    __dadd_rn(__dmul_rn(f2, o[j][2]), __dmul_rn(f3, o[j][3])));
        term[j] = 0.0;
#pragma unroll
        for (int r = 0; r < RL; ++r)
        {
          const double v = __shfl_sync(0xFFFFFFFFu, tr, S4Lanes<RL>::cat_lane(lane, r));
          term[j] = __dadd_rn(term[j], __dmul_rn(v, s8[(sb + Lay::RW) * 2 + r]));
        }
      }
#pragma unroll
      for (int j = 0; j < CPT; ++j)
      {
        double sv;
        if (prm.persite_mode == 2) sv = term[j];
        else
        {
          sv = log(term[j]);
          if (SCALED && osc[j]) sv = __dadd_rn(sv, __dmul_rn((double)osc[j], prm.log_threshold));
          sv = __dmul_rn(sv, (double)s1[tc.tips_s + tc.wgt_off + j * Lay::SROW]);
        }
        if (tc.valid[j] && tc.cat == 0)
        {
          site_sum += sv;
          if (prm.persite) prm.persite[tc.cell[j] / RL] = sv;
        }
      }
    }
  }
  return site_sum;
}

// ------------------------------------------------------------------------------------------------
// general path: any operand kind, any number of chunks; one cell at a time, correctness first
// ------------------------------------------------------------------------------------------------
template <int RL, bool EXACT>
__device__ __forceinline__ Vec4 load_global_x(const LocusHdr * H, unsigned int kind, unsigned int p0, unsigned int pm,
                                              int scidx, unsigned int cell, unsigned int cat, unsigned int & sc)
{
  double v0, v1, v2, v3;
  if (kind == SRC_TIP_DENSE) ld256_nc(H->tip_dense + (size_t)p0 * H->clv_stride + (size_t)cell * 4, v0, v1, v2, v3);
  else ld256(H->clv + (size_t)p0 * H->clv_stride + (size_t)cell * 4, v0, v1, v2, v3);
  sc = (kind == SRC_HBM && scidx >= 0) ? H->scale[(size_t)scidx * H->sites + cell / RL] : 0u;
  const double2 * __restrict__ p = reinterpret_cast<const double2 *>(H->pmat + ((size_t)pm * RL + cat) * 16);
  Vec4 x;
  x.a = dot4<EXACT>(__ldg(p + 0), __ldg(p + 1), v0, v1, v2, v3);
  x.b = dot4<EXACT>(__ldg(p + 2), __ldg(p + 3), v0, v1, v2, v3);
  x.c = dot4<EXACT>(__ldg(p + 4), __ldg(p + 5), v0, v1, v2, v3);
  x.d = dot4<EXACT>(__ldg(p + 6), __ldg(p + 7), v0, v1, v2, v3);
  return x;
}

// One chunk of ops for cell slot j of the thread.  X / psc persist across chunks in xs / pscs.
template <int RL, bool EXACT, int CPT>
__device__ __noinline__ double chunk_general(const TreeParams prm, unsigned int sb, unsigned int lay, unsigned int sst1, unsigned int j,
                                             unsigned int cell, bool valid, unsigned int cat, unsigned int tw0,
                                             unsigned int tw1, unsigned int wgt, double * xs, unsigned int * pscs)
{
  using Lay = S4Layout<RL, CPT>;
  const unsigned int tid = threadIdx.x, lane = tid & 31u;
  const LocusHdr * H = reinterpret_cast<const LocusHdr *>(&s4[sb]);
  const unsigned int pattern = cell / RL;
  const unsigned int lut_t = (lay & 0xFFFFu) + cat * lut_cat_u4(RL);   // lay = lut0 | stack0 << 16; replica A of the tables
  const unsigned int hs = (lane >> 2) & 1u;                            // stack slots hold their halves exchanged for these lanes
  const unsigned int stk_t = (lay >> 16) + tid * 2 + j * (2 * TREE_NT);
  const unsigned int ops = sb + Lay::OPS;
  const unsigned int cn = s1[(sb + Lay::CH) * 4];
  double x0 = xs[0], x1 = xs[1], x2 = xs[2], x3 = xs[3];
  unsigned int psc = *pscs;
  double site_sum = 0.0;

  for (unsigned int k = 0; k < cn; ++k)
  {
    const uint4 w0 = s4[ops + 4 * k], w1 = s4[ops + 4 * k + 1], w2 = s4[ops + 4 * k + 2], w3 = s4[ops + 4 * k + 3];
    const unsigned int ctl = w0.x;
    auto tipmask = [&](unsigned int sel, unsigned int tip) -> unsigned int
    {
      unsigned int word = (sel & 16u) ? tw1 : tw0;
      if (tip >= 16) word = __ldg(H->tipwords + (size_t)pattern * H->tip_words + (tip >> 3));
      return (word >> ((tip & 7u) * 4)) & 15u;
    };
    auto fetch = [&](unsigned int kind, unsigned int sel, unsigned int off, unsigned int p0, unsigned int pm, int scidx,
                     unsigned int & sc) -> Vec4
    {
      Vec4 r;
      if (kind == SRC_TIP_PACKED)
      {
        const unsigned int li = lut_t + off + tipmask(sel, p0) * lut_row_u4(RL);
        const unsigned int xh = (RL == 8 && cat >= 4) ? 1u : 0u;          // (see lut_slot_u4: stored exchanged)
        const double2 u = as_d2(s4[li + xh]), w = as_d2(s4[li + (xh ^ 1u)]);
        r.a = u.x; r.b = u.y; r.c = w.x; r.d = w.y; sc = 0;
      }
      else if (kind == SRC_SLOT)
      {
        const double2 u = as_d2(s4[stk_t + off + hs]), w = as_d2(s4[stk_t + off + (hs ^ 1u)]);
        r.a = u.x; r.b = u.y; r.c = w.x; r.d = w.y;
        sc = s1[sst1 + p0 * (CPT * TREE_NT) + j * TREE_NT + tid];
      }
      else r = load_global_x<RL, EXACT>(H, kind, p0, pm, scidx, cell, cat, sc);
      return r;
    };
    unsigned int akind = (ctl >> OP_AKIND_SHIFT) & 15u, bkind = (ctl >> OP_BKIND_SHIFT) & 15u;
    if (akind == SRC_HBML) akind = SRC_HBM;          // the general path applies P from global memory
    if (bkind == SRC_HBML) bkind = SRC_HBM;
    double o0, o1, o2, o3;
    unsigned int osc = 0;
    if (ctl & OP_EVAL)
    {
      // root CLV that this list did not produce: read it (no P applied)
      if (akind == SRC_TIP_PACKED)
      {
        const unsigned int mask = tipmask(w0.z, w2.x);
        o0 = (double)(mask & 1u); o1 = (double)((mask >> 1) & 1u); o2 = (double)((mask >> 2) & 1u); o3 = (double)((mask >> 3) & 1u);
      }
      else if (akind == SRC_TIP_DENSE) ld256_nc(H->tip_dense + (size_t)w2.x * H->clv_stride + (size_t)cell * 4, o0, o1, o2, o3);
      else
      {
        ld256(H->clv + (size_t)w2.x * H->clv_stride + (size_t)cell * 4, o0, o1, o2, o3);
        if ((int)w1.z >= 0) osc = H->scale[(size_t)(int)w1.z * H->sites + pattern];
      }
    }
    else
    {
      unsigned int asc = 0, bsc = 0;
      const Vec4 a = fetch(akind, w0.z, w0.w, w2.x, w2.y, (int)w2.z, asc);
      if (ctl & OP_BPREV)
      {
        o0 = __dmul_rn(x0, a.a); o1 = __dmul_rn(x1, a.b); o2 = __dmul_rn(x2, a.c); o3 = __dmul_rn(x3, a.d);
        bsc = psc;
      }
      else
      {
        const Vec4 b = fetch(bkind, w1.x, w1.y, w3.x, w3.y, (int)w3.z, bsc);
        o0 = __dmul_rn(a.a, b.a); o1 = __dmul_rn(a.b, b.b); o2 = __dmul_rn(a.c, b.c); o3 = __dmul_rn(a.d, b.d);
      }
      if (ctl & OP_SCALE)
      {
        osc = asc + bsc;
        unsigned int below = (o0 < BPPGPU_SCALE_THRESHOLD) & (o1 < BPPGPU_SCALE_THRESHOLD) &
                             (o2 < BPPGPU_SCALE_THRESHOLD) & (o3 < BPPGPU_SCALE_THRESHOLD);
#pragma unroll
        for (int dd = 1; dd < RL; dd <<= 1) below &= __shfl_xor_sync(0xFFFFFFFFu, below, S4Lanes<RL>::xor_step(dd));
        if (below)
        {
          o0 = __dmul_rn(o0, BPPGPU_SCALE_FACTOR); o1 = __dmul_rn(o1, BPPGPU_SCALE_FACTOR);
          o2 = __dmul_rn(o2, BPPGPU_SCALE_FACTOR); o3 = __dmul_rn(o3, BPPGPU_SCALE_FACTOR);
          osc += 1;
        }
        if (valid && cat == 0) H->scale[(size_t)(int)w1.z * H->sites + pattern] = osc;
      }
      if (valid) st256(H->clv + (((size_t)w0.y + cell) << 2), o0, o1, o2, o3);
      if (ctl & OP_PUSH)
      {
        const Vec4 x = matvec_s4<EXACT>(sb + Lay::PUP + (k * RL + cat) * 9, o0, o1, o2, o3);
        x0 = x.a; x1 = x.b; x2 = x.c; x3 = x.d; psc = osc;
        if (ctl & OP_PARKA)
        {
          s4[stk_t + w1.w + hs] = as_u4(x0, x1);
          s4[stk_t + w1.w + (hs ^ 1u)] = as_u4(x2, x3);
          s1[sst1 + (w1.w / Lay::SLOT) * (CPT * TREE_NT) + j * TREE_NT + tid] = psc;
        }
      }
    }
    if (ctl & OP_ROOT)
    {
      const double tr = __dadd_rn(__dadd_rn(__dmul_rn(H->freqs[0], o0), __dmul_rn(H->freqs[1], o1)),
                                  __dadd_rn(__dmul_rn(H->freqs[2], o2), __dmul_rn(H->freqs[3], o3)));
      double term = 0.0;
#pragma unroll
      for (int r = 0; r < RL; ++r)
      {
        const double v = __shfl_sync(0xFFFFFFFFu, tr, S4Lanes<RL>::cat_lane(lane, r));
        term = __dadd_rn(term, __dmul_rn(v, s8[(sb + Lay::RW) * 2 + r]));
      }
      unsigned int rsc = osc;
      if (ctl & OP_EVAL) rsc = ((int)w1.z >= 0) ? osc : 0;
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
        site_sum += s;
        if (prm.persite) prm.persite[pattern] = s;
      }
    }
  }
  xs[0] = x0; xs[1] = x1; xs[2] = x2; xs[3] = x3; *pscs = psc;
  return site_sum;
}

// LUT[s][cat][mask] = P_tip-edge . bits(mask), same operation order as the mat-vec
template <int RL, bool EXACT, int CPT>
__device__ __forceinline__ void build_lut(unsigned int sb, unsigned int lut0)
{
  using Lay = S4Layout<RL, CPT>;
  const unsigned int entries = s1[(sb + Lay::CH) * 4 + 1] * RL * 16;       // ChunkHdr.ntips
  for (unsigned int e = threadIdx.x; e < entries; e += TREE_NT)
  {
    const unsigned int mask = e & 15u, sc = e >> 4;      // sc = slot*RL + cat
    const Vec4 v = matvec_s4<EXACT>(sb + Lay::TIPP + sc * 9, (double)(mask & 1u), (double)((mask >> 1) & 1u),
                                    (double)((mask >> 2) & 1u), (double)((mask >> 3) & 1u));
    const unsigned int at = lut0 + (sc / RL) * lut_slot_u4(RL) + (sc % RL) * lut_cat_u4(RL) + mask * lut_row_u4(RL);
    const unsigned int xh = (RL == 8 && (sc % RL) >= 4) ? 1u : 0u;   // 8 categories: 4..7 stored with the halves exchanged
    s4[at + xh] = as_u4(v.a, v.b);
    s4[at + (xh ^ 1u)] = as_u4(v.c, v.d);
    if (RL == 4)                                         // replica B: halves exchanged
    {
      s4[at + lut_rep_u4(RL)] = as_u4(v.c, v.d);
      s4[at + lut_rep_u4(RL) + 1] = as_u4(v.a, v.b);
    }
  }
}

// CTAs per SM: 3 / 2 / 1 for 1 / 2 / 4 cells per thread (85 / 128 / 255 registers).  Measured on B200:
// 2 cells at 2 CTAs beats 3 CTAs at 80 registers (spills) for R = 1; 4 cells win for R >= 4, where the
// shared-memory pipe is the limit and the P-matrix loads are amortised over twice the cells.
//
// SCALED_ONLY = false: the kernel with every path (lean / scaled / HBM-class fast instantiations, chunk-by-chunk
// restaging, the cell-at-a-time walker): freshly planned lists, partial updates, big trees, and unscaled full passes.
// SCALED_ONLY = true: ONLY the scaled one-chunk instantiation, in its speculative form, for runs on a cached plan that
// is known (plan_class_kernel) to hold nothing but HDR_NOHBM one-chunk loci: full passes with scale buffers in the
// steady state of the mixing / tau / alpha / qrates moves.  (A lean-only launch was built and measured too: 186
// instead of 242 registers and no stack frame, and not a microsecond faster -- DESIGN.md 4.3 -- so unscaled batches
// stay on the general kernel.)
template <int RL, bool EXACT, int CPT, bool SCALED_ONLY>
__global__ void __launch_bounds__(TREE_NT, s4_ctas_per_sm(CPT))
tree_kernel_s4(const TreeParams prm)
{
  using Lay = S4Layout<RL, CPT>;
  const unsigned int tid = threadIdx.x, lane = tid & 31u;
  const unsigned int ptid = s4_perm<RL>(tid);          // which of the warp's 32 cells this lane owns
  const unsigned int stage_sz = Lay::stage_sz(prm.lut_cap), lut0 = Lay::lut0(prm.lut_cap), stack0 = Lay::stack0(prm.lut_cap);

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
  // The locus block (header + rate weights + chunk 0: op records and the P-matrices of its edges, contiguous, built
  // by the planner) comes in as ONE TMA bulk copy issued by thread 0; every thread waits for it on the stage
  // buffer's mbarrier when the block is first used.  (Round 1 copied it with 16-byte cp.async by all threads.)
  unsigned long long * const mbar = reinterpret_cast<unsigned long long *>(&s4[Lay::MBAR]);
  unsigned int mphase[2] = {0u, 0u};
#if BPPGPU_S4_TMA_STAGE
  auto stage_fetch = [&](unsigned int buf, unsigned long long blk)   // header + rate weights + chunk 0
  {
    if (tid == 0)
    {
      mbar_expect_tx(mbar + buf, stage_sz * 16u);
      bulk_load(&s4[Lay::STAGE0 + buf * stage_sz], prm.blocks + blk, stage_sz * 16u, mbar + buf);
    }
  };
  auto stage_wait = [&](unsigned int buf, bool) { mbar_wait(mbar + buf, mphase[buf]); mphase[buf] ^= 1u; };
  // chunk c >= 1 of the locus at blk -> stage buffer b, behind a copy of the header and the rate weights (two bulk
  // copies on the buffer's mbarrier): the chunks of a big tree alternate between the two stage buffers
  auto chunk_fetch = [&](unsigned int b, unsigned long long blk, unsigned int c)
  {
    if (tid == 0)
    {
      mbar_expect_tx(mbar + b, stage_sz * 16u);
      bulk_load(&s4[Lay::STAGE0 + b * stage_sz], prm.blocks + blk, Lay::CH * 16u, mbar + b);
      bulk_load(&s4[Lay::STAGE0 + b * stage_sz + Lay::CH], prm.blocks + blk + ((size_t)Lay::CH + (size_t)c * Lay::CHUNK) * 16u,
                (stage_sz - Lay::CH) * 16u, mbar + b);
    }
  };
  constexpr bool PINGPONG = Lay::NSTAGE == 2;
#else
  constexpr bool PINGPONG = false;
  auto chunk_fetch = [&](unsigned int, unsigned long long, unsigned int) {};
  auto stage_fetch = [&](unsigned int buf, unsigned long long blk)
  {
    const uint4 * src = reinterpret_cast<const uint4 *>(prm.blocks + blk);
    for (unsigned int w = tid; w < stage_sz; w += TREE_NT) cp_async16(&s4[Lay::STAGE0 + buf * stage_sz + w], src + w);
  };
  // a prefetched block landed with the wait + barrier at the end of the previous tile; a fresh one is waited for here
  auto stage_wait = [&](unsigned int, bool fresh) { (void)mphase; if (fresh) { cp_async_commit(); cp_async_wait_all(); __syncthreads(); } };
#endif
  const unsigned int tips0 = Lay::tips0_u32(prm.n_slots, prm.lut_cap), tips_buf = Lay::tips_buf(prm.tip_words);
  // packed tip words + pattern weight of the sites of tile d -> tip buffer `tb` (asynchronous; the lane of category 0
  // copies, the site's RL lanes read: every read comes after a CTA barrier that follows the copying lane's wait)
  auto load_tips = [&](const TileDesc & d, unsigned int tb)
  {
#pragma unroll
    for (int j = 0; j < CPT; ++j)
    {
      const unsigned int craw = d.cell0 + ptid + j * TREE_NT;
      const unsigned int pat = (craw < d.ncell ? craw : d.ncell - 1) / RL;
      if (ptid % RL) continue;                         // one lane per site copies
      unsigned int * dst = &s1[tips0 + tb * tips_buf + j * Lay::SROW + ptid / RL];
      const unsigned int nw = d.tip_words < prm.tip_words ? d.tip_words : prm.tip_words;
      for (unsigned int w = 0; w < nw; ++w) cp_async4(dst + w * (CPT * Lay::SROW), d.tipwords + (size_t)pat * d.tip_words + w);
      cp_async4(dst + prm.tip_words * (CPT * Lay::SROW), d.weights + pat);
    }
  };

  if (tid == 0) { mbar_init(mbar, 1); mbar_init(mbar + 1, 1); mbar_init_fence(); }
  ring_fetch(t_begin);
  ring_fetch(t_begin + 1);
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();

  TileCtx<CPT> tc;
  tc.lut0 = lut0; tc.stack0 = stack0;
  tc.sst1 = (stack0 + (unsigned)prm.n_slots * Lay::SLOT) * 4;
  load_tips(*reinterpret_cast<const TileDesc *>(&s4[Lay::RING + (t_begin & 3u) * 3]), 0u);
  cp_async_commit();
  cp_async_wait_all();

  unsigned int buf = 0;
  unsigned int staged_locus = 0xFFFFFFFFu;      // locus whose header + chunk 0 + LUT are valid in stage[buf]
  unsigned int prefetched_locus = 0xFFFFFFFFu;  // locus whose block is (being) copied into stage[buf ^ 1]

  for (unsigned int t = t_begin; t < t_end; ++t)
  {
    const unsigned int rs = Lay::RING + (t & 3u) * 3;
    const TileDesc d = *reinterpret_cast<const TileDesc *>(&s4[rs]);
    const unsigned long long blk = *reinterpret_cast<const unsigned long long *>(&s4[rs + 2]);
    // ---- make the locus block current: either it was prefetched into the other buffer, or load it now
    if (d.locus != staged_locus)
    {
      const bool fresh = d.locus != prefetched_locus;
      if (!fresh) buf ^= 1u;                               // on its way (or landed) since the previous tile
      else stage_fetch(buf, blk);                          // (every reader of this buffer passed the last tile's barrier)
      stage_wait(buf, fresh);
      prefetched_locus = 0xFFFFFFFFu;
      build_lut<RL, EXACT, CPT>(Lay::STAGE0 + buf * stage_sz, lut0);
      __syncthreads();
      staged_locus = d.locus;
    }
    // ---- prefetch: tips/weight of tile t+1 -> registers; block of the next locus -> other stage buffer;
    //      descriptor of tile t+2 -> ring
    const unsigned int tb = (t - t_begin) & 1u;          // tip buffer of this tile
    const unsigned int sb = Lay::STAGE0 + buf * stage_sz;
    const LocusHdr * H = reinterpret_cast<const LocusHdr *>(&s4[sb]);
    // a locus of several chunks uses the other stage buffer for its own next chunk (below); the block of the tile
    // after this one is then requested while the last chunk runs
    // (nothing of this is kept in registers across the tile: the one-chunk kernels run at their register cap, and the
    // chunk loop below reads the flag and the next tile's ring entry again)
    if (t + 1 < t_end)
    {
      const unsigned int rn = Lay::RING + ((t + 1) & 3u) * 3;
      const TileDesc dn = *reinterpret_cast<const TileDesc *>(&s4[rn]);
      load_tips(dn, tb ^ 1u);
      const bool pingpong = PINGPONG && H->n_chunks > 1 && (H->flags & HDR_FAST);
      if (Lay::NSTAGE == 2 && !pingpong && dn.locus != d.locus && dn.locus != prefetched_locus)
      {
        stage_fetch(buf ^ 1u, *reinterpret_cast<const unsigned long long *>(&s4[rn + 2]));
        prefetched_locus = dn.locus;
      }
    }
    ring_fetch(t + 2);
    cp_async_commit();

    tc.sb = sb;
    tc.cat = (d.cell0 + ptid) % RL;
    tc.tips_s = tips0 + tb * tips_buf + ptid / RL;
    tc.wgt_off = prm.tip_words * (CPT * Lay::SROW);
#pragma unroll
    for (int j = 0; j < CPT; ++j) tc.tw0[j] = s1[tc.tips_s + j * Lay::SROW];
#pragma unroll
    for (int j = 0; j < CPT; ++j)
    {
      const unsigned int craw = d.cell0 + ptid + j * TREE_NT;
      tc.valid[j] = craw < d.ncell;
      tc.cell[j] = tc.valid[j] ? craw : d.ncell - 1;
    }
    tc.cell0 = d.cell0; tc.ncell = d.ncell;

    double site_sum;
    if constexpr (SCALED_ONLY)
    {
      // specialised launch: the host selected it from the class word of the cached plan; a locus of another class
      // cannot be here (and would poison its lnL instead of computing something wrong)
      double x[CPT][4];
      unsigned int psc[CPT];
      unsigned int wnz = 0;
#pragma unroll
      for (int j = 0; j < CPT; ++j) { x[j][0] = x[j][1] = x[j][2] = x[j][3] = 0.0; psc[j] = 0; }
      unsigned long long spec = 0;
      // one or two categories: (nearly) every lane stores a scaler anyway and there is little lane exchange to save;
      // the speculative form measured slower there (config 2 scaled: 0.515 against 0.476 ms)
      if constexpr (RL < 4) site_sum = tile_fast<RL, EXACT, CPT, 1>(prm, tc, x, psc, wnz, spec);
      else
      {
        constexpr unsigned int LOG2RL = RL == 1 ? 0 : (RL == 2 ? 1 : (RL == 4 ? 2 : 3));
        site_sum = tile_fast<RL, EXACT, CPT, 3>(prm, tc, x, psc, wnz, spec);
        // a site is rescaled when ALL its RL categories are below the threshold at the same op
#pragma unroll
        for (int dd = 1; dd < RL; dd <<= 1) spec &= __shfl_xor_sync(0xFFFFFFFFu, spec, S4Lanes<RL>::xor_step(dd));
        if (__any_sync(0xFFFFFFFFu, spec != 0ull))
        {
          // rare: this warp repeats its cells of the tile with the rescaling in place (no CTA barrier inside tile_fast;
          // the repeat overwrites every CLV, scaler and per-site value the speculative pass stored)
#pragma unroll
          for (int j = 0; j < CPT; ++j) { x[j][0] = x[j][1] = x[j][2] = x[j][3] = 0.0; psc[j] = 0; }
          wnz = 0;
          site_sum = tile_fast<RL, EXACT, CPT, 1>(prm, tc, x, psc, wnz, spec);
        }
        else
        {
          // no site of this warp was rescaled: the scalers of its sites are zero for every scaled op of the list --
          // one sweep, lane = (op, cell row, site of the warp's row), instead of CPT predicated stores per op
          constexpr unsigned int SPW = 32u >> LOG2RL, PER_OP = CPT * SPW;
          const unsigned int ops = sb + Lay::OPS, cn = s1[(sb + Lay::CH) * 4];
          for (unsigned int i = lane; i < cn * PER_OP; i += 32)
          {
            const unsigned int k = i / PER_OP, r = i % PER_OP, j = r / SPW, sw = r % SPW;
            const unsigned int ctl = s1[(ops + 4 * k) * 4];
            const int dsc = (int)s1[(ops + 4 * k + 1) * 4 + 2];
            const unsigned int cell = tc.cell0 + j * TREE_NT + (tid & ~31u) + sw * RL;
            if ((ctl & OP_SCALE) && cell < tc.ncell) H->scale[(size_t)dsc * H->sites + (cell >> LOG2RL)] = 0u;
          }
        }
      }
      if (H->n_chunks != 1 || !(H->flags & HDR_NOHBM)) site_sum = __longlong_as_double(0x7FF8000000000000ll);
    }
    else
    {
    // another chunk of the current locus -> the stage buffer (synchronously, in place of the one that is there)
    auto restage = [&](unsigned int c)
    {
      __syncthreads();
      const uint4 * src = reinterpret_cast<const uint4 *>(prm.blocks + blk) + Lay::CH + (size_t)c * Lay::CHUNK;
      for (unsigned int w = tid; w < stage_sz - Lay::CH; w += TREE_NT) s4[sb + Lay::CH + w] = __ldg(src + w);
      __syncthreads();
      build_lut<RL, EXACT, CPT>(sb, lut0);
      __syncthreads();
      staged_locus = 0xFFFFFFFFu;          // chunk 0 is gone
    };
#ifdef BPPGPU_NO_MULTICHUNK_FAST
    if ((H->flags & HDR_FAST) && H->n_chunks == 1)
#else
    if (H->flags & HDR_FAST)
#endif
    {
      const unsigned int flags = H->flags, n_chunks = H->n_chunks;
      double x[CPT][4];
      unsigned int psc[CPT];
      unsigned int wnz = 0;                    // some scaler count of this warp's cells is non-zero (this tile)
      unsigned long long spec = 0;             // (MODE 3 only)
#pragma unroll
      for (int j = 0; j < CPT; ++j) { x[j][0] = x[j][1] = x[j][2] = x[j][3] = 0.0; psc[j] = 0; }
      // no HBM-class operand and no scaler: the lean instantiation; no HBM-class operand: the scaled one; anything
      // else on the fast path runs the full one.  One chunk for trees of up to 16 ops whose tips fit the lookup
      // tables, chunk by chunk beyond (X, its scaler count and the parked values carry over)
      site_sum = 0.0;
      // chunk c >= 1 becomes current: with two stage buffers its copy was started while chunk c-1 ran, so what is
      // left is one barrier (every warp is done with chunk c-1, its buffer and the lookup tables), the wait for the
      // copy, the table rebuild and the barrier behind it; otherwise it is staged in place, synchronously
      const bool pingpong = PINGPONG && n_chunks > 1;          // (HDR_FAST holds in this branch)
      auto advance = [&](unsigned int c)
      {
        if (pingpong)
        {
          __syncthreads();
          buf ^= 1u;
          stage_wait(buf, false);
          tc.sb = Lay::STAGE0 + buf * stage_sz;
          build_lut<RL, EXACT, CPT>(tc.sb, lut0);
          __syncthreads();
          staged_locus = 0xFFFFFFFFu;          // chunk 0 is gone
        }
        else restage(c);
      };
      // before chunk c runs: start the copy of what comes next into the other buffer (whose last readers passed the
      // barrier in advance(c), or the end of the previous tile): the locus' next chunk, or after the last one the
      // block of the next tile (this locus again, or the next one)
      auto ahead = [&](unsigned int c)
      {
        if (!pingpong) return;
        if (c + 1 < n_chunks) chunk_fetch(buf ^ 1u, blk, c + 1);
        else if (t + 1 < t_end)                    // (ring slot t+1 stays valid through tile t: ring_fetch(t+2) fills another)
        {
          const unsigned int rn = Lay::RING + ((t + 1) & 3u) * 3;
          stage_fetch(buf ^ 1u, *reinterpret_cast<const unsigned long long *>(&s4[rn + 2]));
          prefetched_locus = reinterpret_cast<const TileDesc *>(&s4[rn])->locus;
        }
      };
      if (flags & HDR_SIMPLE)
        for (unsigned int c = 0; c < n_chunks; ++c)
        {
          if (c > 0) advance(c);
          ahead(c);
          site_sum += tile_fast<RL, EXACT, CPT, 0>(prm, tc, x, psc, wnz, spec);
        }
      else if (flags & HDR_NOHBM)
        for (unsigned int c = 0; c < n_chunks; ++c)
        {
          if (c > 0) advance(c);
          ahead(c);
          site_sum += tile_fast<RL, EXACT, CPT, 1>(prm, tc, x, psc, wnz, spec);
        }
      else
        for (unsigned int c = 0; c < n_chunks; ++c)
        {
          if (c > 0) advance(c);
          ahead(c);
          site_sum += tile_fast<RL, EXACT, CPT, 2>(prm, tc, x, psc, wnz, spec);
        }
    }
    else
    {
      // general path: chunk by chunk (later chunks are staged in place, synchronously), cell by cell
      double xs[CPT][4];
      unsigned int pscs[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j) { xs[j][0] = xs[j][1] = xs[j][2] = xs[j][3] = 0.0; pscs[j] = 0; }
      site_sum = 0.0;
      const unsigned int n_chunks = H->n_chunks;
      for (unsigned int c = 0; c < n_chunks; ++c)
      {
        if (c > 0) restage(c);
#pragma unroll
        for (int j = 0; j < CPT; ++j)
          site_sum += chunk_general<RL, EXACT, CPT>(prm, sb, lut0 | (stack0 << 16), tc.sst1, j, tc.cell[j], tc.valid[j], tc.cat, tc.tw0[j],
                                                    (d.tip_words > 1 && prm.tip_words > 1) ? s1[tc.tips_s + (CPT + j) * Lay::SROW] : 0u,
                                                    s1[tc.tips_s + tc.wgt_off + j * Lay::SROW], xs[j], &pscs[j]);
      }
    }

    }
    // ---- deterministic tile reduction of the weighted site lnL values
    if (prm.tile_partial)
    {
      double v = site_sum;
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
  }
}

}  // namespace bppgpu
