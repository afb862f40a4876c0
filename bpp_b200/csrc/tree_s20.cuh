// tree_s20.cuh -- tree-fused Felsenstein pruning for 20 states (amino acids) on the FP64 tensor path.
//
// Per (site, category) the CLV update is two 20x20 mat-vecs; across a tile of sites it is the dense
// GEMM [20x20] . [20 x sites] per category, which is what the FP64 DMMA instruction
// (mma.sync.aligned.m8n8k4.f64, SASS DMMA.8x8x4) is for.  tcgen05 has no FP64 kind, and DMMA reaches
// the same 37 TFLOP/s as DFMA on B200 (profiles/microbench/fp64_peak.cu) with 1/16 of the
// instructions and the P-matrix fragments held in registers.
//
// Same stack-machine / push model as tree_s4.cuh:
//   X = P_edge . clv travels between ops (registers, or a shared-memory stack slot), an op is the
//   elementwise product parent = X_a * X_b, the parent is written to HBM once and pushed through the
//   P-matrix of the edge above it with 15 DMMAs per 8 sites (M = 24 padded states, K = 20, N = 8).
//   A packed tip child needs no arithmetic at all: X = column `state` of the edge's P-matrix
//   (ambiguity codes use up to four precomputed extra columns = sums of columns over the mask).
// One warp owns 16 sites of ONE category (so P fragments are warp-uniform); a CTA of 8 warps covers
// 128/RL sites x RL categories.  All register-resident vectors are in the DMMA accumulator layout:
//   lane = 4*r + q:  V[g][mt][e] = state 8*mt + r  of site 8*g + 2*q + e      (g < S20_NG, mt < 3, e < 2)
// P-matrices are read through L1 straight from the locus' pmatrix block (they are uniform per warp
// and a few kB per edge); no shared-memory staging is needed for a tensor-bound kernel.
//
// Reference semantics: core_partials.c:585-756 (generic states), per-site scaling :739-754, root
// core_likelihood.c:24-212.  The reference's own 20-state kernels (AVX: mul+add, AVX2: FMA) differ
// from each other in rounding; parity here is the lnL bar (<= 1e-10 relative), not bit identity.
#pragma once
#include "common.cuh"
#include "plan.cuh"

namespace bppgpu {

constexpr int S20 = 20;
constexpr int S20_EXT = 4;          // extra tip columns (ambiguity masks) per tip edge and category
constexpr int S20_NG = 2;           // groups of 8 sites per warp (16 sites): keeps the kernel at <= 128 registers
constexpr int S20_WS = 8 * S20_NG;  // sites per warp
constexpr int S20_NT = 256;              // threads per CTA of the site-major 20-state kernel
constexpr int S20_TILE = (S20_NT / 32) * S20_WS;   // cells (site, category) per CTA tile

struct OpRec20                      // 64 bytes
{
  unsigned int ctl;                 // OP_* flags | a_kind << 8 | b_kind << 12
  unsigned int dst_cell;            // parent buffer offset in cells: dst * sites * RL
  unsigned int a_p0, a_pm; int a_sc; unsigned int a_ext;    // ext: offset (doubles) of the tip's extra columns in the block
  unsigned int b_p0, b_pm; int b_sc; unsigned int b_ext;
  unsigned int up_pm; int dsc; unsigned int park_slot;
  unsigned int a_st, b_st, up_st;   // staged-matrix indices of the three edges (category-major kernel)
};
static_assert(sizeof(OpRec20) == 64, "OpRec20 must be 64 bytes");

struct Hdr20                        // 384 bytes, start of the locus block
{
  double * clv;
  double * tip_dense;
  unsigned int * scale;
  const unsigned char * tip_cols;   // [tip][cols_pitch] column ids: 0..19 = state, 20.. = extra column
  const double * pmat;
  const unsigned int * weights;
  unsigned long long clv_stride;
  unsigned int sites, nops, tips, n_ext_ops;
  double freqs[S20];
  double rw[8];
  unsigned int n_stage, stage_off;  // staged-matrix list: n entries of (pmatrix index, ext offset | 0) at byte stage_off
  unsigned int cols_pitch, site_stride;    // row pitch of tip_cols (bytes); CLV strides (doubles), see LocusDev
  unsigned int cat_stride, pad0;
  double pad[8];
};
static_assert(sizeof(Hdr20) == 384, "Hdr20 must be 384 bytes");

// block = [Hdr20][staged-matrix list, S20_LIST_CAP entries][op records][long lists: the list moves here][extra tip columns]
constexpr unsigned int S20_LIST_CAP = 64;                                             // (pmatrix index, ext offset | 0) entries
constexpr unsigned int S20_RECS_OFF = (unsigned int)sizeof(Hdr20) + S20_LIST_CAP * 8;  // byte offset of the op records
__host__ __device__ inline size_t block20_bytes(unsigned RL, unsigned nops_max)
{
  // header + list + ops + a long list (three edges per op) + extra columns for at most two tip children per op
  return S20_RECS_OFF + (size_t)nops_max * sizeof(OpRec20) + (size_t)nops_max * 3 * 8 + 16 +
         (size_t)nops_max * 2 * RL * S20 * S20_EXT * 8;
}

// ---------------------------------------------------------------- planner (one warp per locus)
__global__ void __launch_bounds__(128)
plan_kernel_blocks20(const LocusDev * __restrict__ loci, const unsigned int * __restrict__ batch_locus,
                     unsigned int n_loci, const unsigned int * __restrict__ op_off,
                     const RawOp * __restrict__ ops, const unsigned int * __restrict__ root_clv,
                     const int * __restrict__ root_sc, int want_root,
                     unsigned char * __restrict__ blocks, const unsigned long long * __restrict__ blk_off,
                     const unsigned int * __restrict__ tile_first, unsigned long long * __restrict__ tile_blk,
                     unsigned int * __restrict__ plan_count,
                     unsigned char * __restrict__ scratch, const unsigned long long * __restrict__ scratch_off,
                     int max_slots, unsigned int RL, unsigned long long hdr_shift, int do_ext)
{
  // hdr_shift: bytes in front of the header inside the block (the category-major kernel keeps its matrix images
  // there); do_ext: sum the tips' ambiguity columns here (the category-major path does it while building images)
  const unsigned int bl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned int lane = threadIdx.x & 31u;
  if (bl >= n_loci) return;
  const LocusDev & L = loci[batch_locus[bl]];
  const unsigned int first = op_off[bl], n = op_off[bl + 1] - first;
  const RawOp * __restrict__ o = ops + first;
  unsigned char * blk = blocks + blk_off[bl] + hdr_shift;
  for (unsigned int t = tile_first[bl] + lane; t < tile_first[bl + 1]; t += 32)
  {
    tile_blk[2 * (size_t)t] = blk_off[bl] + hdr_shift;
    tile_blk[2 * (size_t)t + 1] = 0;
  }
  OpRec20 * const recs_g = reinterpret_cast<OpRec20 *>(blk + S20_RECS_OFF);
  const unsigned int T = L.tips;
  unsigned int cnt = 0, n_ext = 0, ns_out = 0, soff_out = 0;
  // small loci (the common case) are planned in shared memory: the serial planner below patches earlier records
  // (OP_PUSH, OP_PARKA) and looks buffers up in `where`, and every such access was a dependent global round trip
  __shared__ __align__(16) OpRec20 s_rec[4][33];
  __shared__ unsigned int s_where[4][64];
  __shared__ unsigned char s_slot[4][64];
  __shared__ uint2 s_list[4][96];
  const unsigned int wib = (threadIdx.x >> 5) & 3u;
  const bool small = (n + 1 <= 33) && (L.clv_buffers <= 64);
  OpRec20 * const recs = small ? s_rec[wib] : recs_g;

  // evaluation order: Sethi-Ullman DFS for lists of <= 32 ops (fewest parked values), as given otherwise
  __shared__ unsigned char s_ord[4][32];
  unsigned char * ord = s_ord[(threadIdx.x >> 5) & 3u];
  const bool reorder = n <= 32;
  if (reorder)
  {
    const SuLane su = su_order_small(o, n, T);
    if (lane < n) ord[su.pos] = (unsigned char)lane;
  }
  __syncwarp();

  if (lane == 0)
  {
    unsigned int * where = small ? s_where[wib] : reinterpret_cast<unsigned int *>(scratch + scratch_off[bl]);
    unsigned char * slot_of = small ? s_slot[wib] : reinterpret_cast<unsigned char *>(where + L.clv_buffers);
    for (unsigned int k = 0; k < n; ++k)
    {
      where[o[k].parent - T] = 0;
      if (o[k].left >= T) where[o[k].left - T] = 0;
      if (o[k].right >= T) where[o[k].right - T] = 0;
    }
    const unsigned int rootc = want_root ? root_clv[bl] : 0xFFFFFFFFu;
    const unsigned int total = n + (want_root ? 1u : 0u);
    // staged-matrix list: one entry per distinct (P-matrix, form) -- a tip edge is staged transposed with its
    // ambiguity columns, an inner edge as tensor fragments -- so a full pass has exactly 2T - 2 entries however
    // often a list revisits a node.  Short lists collect it in shared memory; it ends up at the fixed offset
    // sizeof(Hdr20) when it fits S20_LIST_CAP, behind the op records otherwise.
    const unsigned int long_off = (unsigned int)(S20_RECS_OFF + (size_t)total * sizeof(OpRec20));
    uint2 * stage = small ? s_list[wib] : reinterpret_cast<uint2 *>(blk + long_off);
    unsigned int n_stage = 0;
    const unsigned int ext0 = (unsigned int)(((long_off + (size_t)total * 3 * 8 + 15) & ~(size_t)15) / 8);   // doubles, 16-byte aligned
    auto stage_entry = [&](unsigned int pm, bool packed) -> unsigned int
    {
      for (unsigned int s = 0; s < n_stage; ++s)
      {
        const uint2 ent = stage[s];
        if (ent.x == pm && (ent.y != 0u) == packed) return s;
      }
      unsigned int ext = 0;
      if (packed) { ext = ext0 + n_ext * RL * S20 * S20_EXT; ++n_ext; }
      stage[n_stage] = make_uint2(pm, ext);
      return n_stage++;
    };
    unsigned int free_slots = (max_slots >= 32) ? 0xFFFFFFFFu : ((1u << max_slots) - 1u);
    unsigned int prev = 0xFFFFFFFFu, prev_k = 0;
    bool root_done = false;
    const unsigned int cells_per_buf = L.sites * RL;
    int last_root = -1;                        // the last op that writes the root CLV carries OP_ROOT
    for (unsigned int k = 0; k < n; ++k) if (want_root && o[reorder ? ord[k] : k].parent == rootc) last_root = (int)k;
    for (unsigned int k = 0; k < n; ++k)       // k = position in the evaluation order = index of the OpRec20
    {
      RawOp r = o[reorder ? ord[k] : k];
      if (L.scale_buffers == 0) r.psc = r.lsc = r.rsc = -1;        // a locus without scale buffers cannot scale
      const unsigned int child[2] = { r.left, r.right };
      unsigned int kind[2], p0[2], pm[2], ext[2], st[2]; int sc[2];
      int prev_child = -1;
      unsigned int consumed_slots = 0;
      for (int c = 0; c < 2; ++c)
      {
        const unsigned int idx = child[c];
        pm[c] = c ? r.rpm : r.lpm; p0[c] = 0; ext[c] = 0; sc[c] = -1; st[c] = 0;
        if (idx < T)
        {
          p0[c] = idx;
          kind[c] = L.tip_is_dense[idx] ? SRC_TIP_DENSE : SRC_TIP_PACKED;
          st[c] = stage_entry(pm[c], kind[c] == SRC_TIP_PACKED);
          ext[c] = stage[st[c]].y;
        }
        else
        {
          const unsigned int b = idx - T;
          if (b == prev && prev_child < 0)
          {
            kind[c] = SRC_PREV; prev_child = c;
            recs[prev_k].ctl |= OP_PUSH; recs[prev_k].up_pm = pm[c];
            recs[prev_k].up_st = stage_entry(pm[c], false);
          }
          else if (where[b])
          {
            const unsigned int s = slot_of[b];
            kind[c] = SRC_SLOT; p0[c] = s; consumed_slots |= 1u << s;
            recs[where[b] - 1].ctl |= OP_PUSH; recs[where[b] - 1].up_pm = pm[c];
            recs[where[b] - 1].up_st = stage_entry(pm[c], false);
            where[b] = 0;
          }
          else
          {
            kind[c] = SRC_HBM; p0[c] = b; sc[c] = c ? r.rsc : r.lsc;
            st[c] = stage_entry(pm[c], false);
          }
        }
      }
      if (prev != 0xFFFFFFFFu && prev_child < 0 && free_slots)
      {
        const int s = __ffs(free_slots) - 1;
        free_slots &= ~(1u << s);
        where[prev] = prev_k + 1; slot_of[prev] = (unsigned char)s;
        recs[prev_k].ctl |= OP_PARKA; recs[prev_k].park_slot = (unsigned)s;
      }
      free_slots |= consumed_slots;
      const int ia = prev_child == 0 ? 1 : 0, ib = prev_child == 0 ? 0 : 1;
      OpRec20 q;
      q.ctl = (kind[ia] << OP_AKIND_SHIFT) | (kind[ib] << OP_BKIND_SHIFT);
      q.dst_cell = (r.parent - T) * cells_per_buf; q.dsc = r.psc; q.park_slot = 0; q.up_pm = 0;
      q.a_st = st[ia]; q.b_st = st[ib]; q.up_st = 0;
      q.a_p0 = p0[ia]; q.a_pm = pm[ia]; q.a_sc = sc[ia]; q.a_ext = ext[ia];
      q.b_p0 = p0[ib]; q.b_pm = pm[ib]; q.b_sc = sc[ib]; q.b_ext = ext[ib];
      if (r.psc >= 0) q.ctl |= OP_SCALE;
      if (prev_child >= 0) q.ctl |= OP_BPREV;
      if ((int)k == last_root) { q.ctl |= OP_ROOT; root_done = true; }
      recs[k] = q;
      prev = r.parent - T; prev_k = k;
    }
    cnt = n;
    if (want_root && !root_done)
    {
      OpRec20 q;
      memset(&q, 0, sizeof(q));
      unsigned int kind;
      q.dsc = root_sc[bl]; q.a_sc = root_sc[bl];
      if (rootc < T) { q.a_p0 = rootc; kind = L.tip_is_dense[rootc] ? SRC_TIP_DENSE : SRC_TIP_PACKED; }
      else { kind = SRC_HBM; q.a_p0 = rootc - T; }
      q.ctl = OP_EVAL | OP_ROOT | (kind << OP_AKIND_SHIFT);
      recs[n] = q;
      cnt = n + 1;
    }
    Hdr20 * H = reinterpret_cast<Hdr20 *>(blk);
    H->clv = L.clv; H->tip_dense = L.tip_dense; H->scale = L.scale; H->tip_cols = L.tip_cols; H->pmat = L.pmat;
    H->weights = L.weights; H->clv_stride = L.clv_stride; H->sites = L.sites; H->nops = cnt; H->tips = L.tips;
    H->n_ext_ops = n_ext; H->n_stage = n_stage; H->cols_pitch = L.cols_pitch; H->site_stride = L.site_stride; H->cat_stride = L.cat_stride;
    H->stage_off = (small && n_stage <= S20_LIST_CAP) ? (unsigned int)sizeof(Hdr20) : long_off;
    ns_out = n_stage; soff_out = H->stage_off;
    for (int j = 0; j < S20; ++j) H->freqs[j] = L.freqs[j];
    for (unsigned int j = 0; j < 8; ++j) H->rw[j] = j < RL ? L.rate_weights[j] : 0.0;
    plan_count[bl] = cnt;
  }
  cnt = __shfl_sync(0xFFFFFFFFu, cnt, 0);
  __syncwarp();
  if (small)
  {
    const uint4 * src = reinterpret_cast<const uint4 *>(s_rec[wib]);
    uint4 * dst = reinterpret_cast<uint4 *>(recs_g);
    for (unsigned int w = lane; w < cnt * 4; w += 32) dst[w] = src[w];
    const unsigned int ns = __shfl_sync(0xFFFFFFFFu, ns_out, 0);
    uint2 * ldst = reinterpret_cast<uint2 *>(blk + __shfl_sync(0xFFFFFFFFu, soff_out, 0));
    for (unsigned int w = lane; w < ns; w += 32) ldst[w] = s_list[wib][w];
  }
  if (!do_ext) return;
  // extra tip columns: ext[cat][i][x] = sum over the states j of ambiguity mask x of P[i][j], j ascending
  double * base = reinterpret_cast<double *>(blk);
  const unsigned int n_ext_cols = L.n_ext_cols;
  unsigned int cmask[S20_EXT];
#pragma unroll
  for (int x = 0; x < S20_EXT; ++x) cmask[x] = (unsigned)x < n_ext_cols ? L.colmask[x] : 0u;
  for (unsigned int k = 0; k < cnt; ++k)
  {
    const OpRec20 q = recs[k];
    if (q.ctl & OP_EVAL) continue;
    for (int c = 0; c < 2; ++c)
    {
      const unsigned int kind = (q.ctl >> (c ? OP_BKIND_SHIFT : OP_AKIND_SHIFT)) & 15u;
      if (kind != SRC_TIP_PACKED) continue;
      const unsigned int pm = c ? q.b_pm : q.a_pm;
      double * ext = base + (c ? q.b_ext : q.a_ext);
      // one lane per (category, row): the row's 20 entries are read once (five 256-bit loads) and feed all the
      // ambiguity columns (the first version re-read them entry by entry: ~200 dependent 8-byte loads per lane and
      // tip operand, most of this kernel's time)
      for (unsigned int row = lane; row < RL * S20; row += 32)
      {
        const double * P = L.pmat + ((size_t)pm * RL + row / S20) * (S20 * S20) + (size_t)(row % S20) * S20;
        double p[S20];
#pragma unroll
        for (int v = 0; v < S20 / 4; ++v) ld256(P + 4 * v, p[4 * v], p[4 * v + 1], p[4 * v + 2], p[4 * v + 3]);
        double acc[S20_EXT];
#pragma unroll
        for (int x = 0; x < S20_EXT; ++x)
        {
          acc[x] = 0.0;
          const unsigned int mask = (unsigned)x < n_ext_cols ? cmask[x] : 0u;
#pragma unroll
          for (int j = 0; j < S20; ++j) if ((mask >> j) & 1u) acc[x] += p[j];
        }
        double2 * dst = reinterpret_cast<double2 *>(ext + (size_t)row * S20_EXT);
        dst[0] = make_double2(acc[0], acc[1]);
        dst[1] = make_double2(acc[2], acc[3]);
      }
    }
  }
}

// ---------------------------------------------------------------- DMMA helpers
__device__ __forceinline__ void dmma(double & d0, double & d1, const double a, const double b)
{
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

struct V20 { double v[S20_NG][3][2]; };

// V = P . tile, P = 20x20 row-major in global memory (read through L1), tile = the warp's 32 x 20 doubles
__device__ __forceinline__ void matvec20(const double * __restrict__ P, const double * tile, unsigned int r, unsigned int q,
                                         V20 & out)
{
  double a[3][5];
#pragma unroll
  for (int mt = 0; mt < 3; ++mt)
#pragma unroll
    for (int ks = 0; ks < 5; ++ks)
    {
      const unsigned int i = 8 * mt + r;
      a[mt][ks] = (i < S20) ? __ldg(P + i * S20 + 4 * ks + q) : 0.0;
    }
  // k-steps outermost: the S20_NG x 3 accumulator chains are independent, so consecutive DMMAs never wait
  // for each other (five dependent DMMAs back to back would expose the tensor pipe's latency)
  double b[S20_NG][5];
#pragma unroll
  for (int g = 0; g < S20_NG; ++g)
#pragma unroll
    for (int ks = 0; ks < 5; ++ks) b[g][ks] = tile[(8 * g + r) * S20 + 4 * ks + q];
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

// ---------------------------------------------------------------- the kernel
// grid: persistent; tile t = S20_NT cells = (S20_NT / RL) sites x RL categories of one locus
template <int RL>
__global__ void __launch_bounds__(S20_NT, 2)
tree_kernel_s20(const TreeParams prm)
{
  extern __shared__ __align__(16) unsigned char smem20[];
  constexpr int NW = S20_NT / 32;                 // warps
  constexpr int SPT = S20_TILE / RL;               // sites per tile
  const unsigned int tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const unsigned int r = lane >> 2, q = lane & 3u;
  const unsigned int cat = warp % RL, pg = warp / RL;

  // shared memory: [tiles NW x S20_WS x 20 doubles][exch SPT x RL doubles][flags SPT x RL u32][red 32 doubles]
  //                [stack: slots x NW x 32 lanes x 24 doubles][stack scalers: slots x NW x 32 x 8 u32]
  double * s_tile = reinterpret_cast<double *>(smem20) + (size_t)warp * S20_WS * S20;
  double * s_exch = reinterpret_cast<double *>(smem20) + (size_t)NW * S20_WS * S20;
  unsigned int * s_flag = reinterpret_cast<unsigned int *>(s_exch + SPT * RL);
  double * s_red = reinterpret_cast<double *>(s_flag + SPT * RL);
  double * s_stack = s_red + 32;
  unsigned int * s_sstack = reinterpret_cast<unsigned int *>(s_stack + (size_t)prm.n_slots * NW * 32 * S20_NG * 6);

  const unsigned int t_begin = (unsigned int)(((unsigned long long)prm.n_tiles * blockIdx.x) / gridDim.x);
  const unsigned int t_end = (unsigned int)(((unsigned long long)prm.n_tiles * (blockIdx.x + 1)) / gridDim.x);

  for (unsigned int t = t_begin; t < t_end; ++t)
  {
    const unsigned int bl = prm.tile_locus[t];
    const unsigned int site0 = prm.tile_cell0[t] / RL + pg * S20_WS;   // first site of this warp
    const unsigned char * blk = prm.blocks + prm.tile_blk[2 * (size_t)t];
    const Hdr20 * H = reinterpret_cast<const Hdr20 *>(blk);
    const OpRec20 * recs = reinterpret_cast<const OpRec20 *>(blk + S20_RECS_OFF);
    const double * base = reinterpret_cast<const double *>(blk);
    const unsigned int sites = __ldg(&H->sites), nops = __ldg(&H->nops), cols_pitch = __ldg(&H->cols_pitch);
    const unsigned int sstr = __ldg(&H->site_stride), cstr = __ldg(&H->cat_stride);
    double * const clv = H->clv;
    const double * const pmat = H->pmat;
    const unsigned long long stride = __ldg(&H->clv_stride);
    (void)bl;

    // the lane's sites in accumulator layout
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

    V20 X;                                   // X of the previous op (accumulator layout)
    unsigned int xsc[S20_NG][2];
    double site_sum = 0.0;
#pragma unroll
    for (int g = 0; g < S20_NG; ++g)
#pragma unroll
      for (int e = 0; e < 2; ++e)
      {
        xsc[g][e] = 0;
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) X.v[g][mt][e] = 0.0;
      }

    // global <-> tile, coalesced in 32-byte chunks: chunk c of the warp = site c/5, part c%5
    auto tile_to_global = [&](double * dst_buf)
    {
#pragma unroll
      for (int it = 0; it < (S20_WS * 5 + 31) / 32; ++it)
      {
        const unsigned int c = it * 32 + lane, n = c / 5, part = c % 5;
        if (c >= S20_WS * 5) break;
        const unsigned int s = site0 + n;
        const double2 u = *reinterpret_cast<const double2 *>(s_tile + n * S20 + part * 4);
        const double2 w = *reinterpret_cast<const double2 *>(s_tile + n * S20 + part * 4 + 2);
        if (s < sites) st256(dst_buf + (size_t)s * sstr + (size_t)cat * cstr + part * 4, u.x, u.y, w.x, w.y);
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
        if (coherent) ld256(src_buf + (size_t)s * sstr + (size_t)cat * cstr + part * 4, a, b, cc, d);
        else ld256_nc(src_buf + (size_t)s * sstr + (size_t)cat * cstr + part * 4, a, b, cc, d);
        *reinterpret_cast<double2 *>(s_tile + n * S20 + part * 4) = make_double2(a, b);
        *reinterpret_cast<double2 *>(s_tile + n * S20 + part * 4 + 2) = make_double2(cc, d);
      }
    };
    auto tile_to_v = [&](V20 & v)
    {
#pragma unroll
      for (int g = 0; g < S20_NG; ++g)
#pragma unroll
        for (int mt = 0; mt < 3; ++mt)
#pragma unroll
          for (int e = 0; e < 2; ++e)
          {
            const unsigned int i = 8 * mt + r;
            v.v[g][mt][e] = (i < S20) ? s_tile[(8 * g + 2 * q + e) * S20 + i] : 0.0;
          }
    };
    auto v_to_tile = [&](const V20 & v)
    {
#pragma unroll
      for (int g = 0; g < S20_NG; ++g)
#pragma unroll
        for (int mt = 0; mt < 3; ++mt)
#pragma unroll
          for (int e = 0; e < 2; ++e)
          {
            const unsigned int i = 8 * mt + r;
            if (i < S20) s_tile[(8 * g + 2 * q + e) * S20 + i] = v.v[g][mt][e];
          }
    };
    // X of one operand that is not the register-resident previous result
    auto fetch = [&](unsigned int kind, unsigned int p0, unsigned int pm, int scidx, unsigned int ext, V20 & v,
                     unsigned int (&sc)[S20_NG][2])
    {
      const double * P = pmat + ((size_t)pm * RL + cat) * (S20 * S20);
      if (kind == SRC_TIP_PACKED)
      {
        const double * E = base + ext + (size_t)cat * S20 * S20_EXT;
#pragma unroll
        for (int g = 0; g < S20_NG; ++g)
#pragma unroll
          for (int e = 0; e < 2; ++e)
          {
            const unsigned int col = __ldg(H->tip_cols + (size_t)p0 * cols_pitch + sitev[g][e]);
            sc[g][e] = 0;
#pragma unroll
            for (int mt = 0; mt < 3; ++mt)
            {
              const unsigned int i = 8 * mt + r;
              double x = 0.0;
              if (i < S20) x = (col < S20) ? __ldg(P + i * S20 + col) : __ldg(E + i * S20_EXT + (col - S20));
              v.v[g][mt][e] = x;
            }
          }
      }
      else if (kind == SRC_SLOT)
      {
        const double * st = s_stack + (size_t)(p0 * NW + warp) * (S20_NG * 6 * 32) + lane;
        const unsigned int * ss = s_sstack + (size_t)(p0 * NW + warp) * (S20_NG * 2 * 32) + lane;
#pragma unroll
        for (int g = 0; g < S20_NG; ++g)
#pragma unroll
          for (int e = 0; e < 2; ++e)
          {
            sc[g][e] = ss[(g * 2 + e) * 32];
#pragma unroll
            for (int mt = 0; mt < 3; ++mt) v.v[g][mt][e] = st[((g * 3 + mt) * 2 + e) * 32];
          }
      }
      else
      {
        // HBM-resident child CLV (or dense tip): load it into the tile and apply the edge's P-matrix
        __syncwarp();
        global_to_tile((kind == SRC_TIP_DENSE ? H->tip_dense : clv) + (size_t)p0 * stride, kind == SRC_HBM);
        __syncwarp();
        matvec20(P, s_tile, r, q, v);
#pragma unroll
        for (int g = 0; g < S20_NG; ++g)
#pragma unroll
          for (int e = 0; e < 2; ++e)
            sc[g][e] = (kind == SRC_HBM && scidx >= 0) ? H->scale[(size_t)scidx * sites + sitev[g][e]] : 0u;
      }
    };

    for (unsigned int k = 0; k < nops; ++k)
    {
      const uint4 w0 = __ldg(reinterpret_cast<const uint4 *>(recs + k));
      const uint4 w1 = __ldg(reinterpret_cast<const uint4 *>(recs + k) + 1);
      const uint4 w2 = __ldg(reinterpret_cast<const uint4 *>(recs + k) + 2);
      const uint4 w3 = __ldg(reinterpret_cast<const uint4 *>(recs + k) + 3);
      const unsigned int ctl = w0.x;
      const unsigned int akind = (ctl >> OP_AKIND_SHIFT) & 15u, bkind = (ctl >> OP_BKIND_SHIFT) & 15u;
      V20 O;
      unsigned int osc[S20_NG][2];

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
              const unsigned int col = __ldg(H->tip_cols + (size_t)w0.z * cols_pitch + sitev[g][e]);
              const unsigned int mask = (col < S20) ? (1u << col) : prm.loci[prm.batch_locus[bl]].colmask[col - S20];
#pragma unroll
              for (int mt = 0; mt < 3; ++mt) O.v[g][mt][e] = (double)((mask >> (8 * mt + r)) & 1u);
              osc[g][e] = 0;
            }
        }
        else
        {
          __syncwarp();
          global_to_tile((akind == SRC_TIP_DENSE ? H->tip_dense : clv) + (size_t)w0.z * stride, akind == SRC_HBM);
          __syncwarp();
          tile_to_v(O);
#pragma unroll
          for (int g = 0; g < S20_NG; ++g)
#pragma unroll
            for (int e = 0; e < 2; ++e)
              osc[g][e] = (akind == SRC_HBM && (int)w1.x >= 0) ? H->scale[(size_t)(int)w1.x * sites + sitev[g][e]] : 0u;
        }
      }
      else
      {
        V20 A;
        unsigned int asc[S20_NG][2];
        fetch(akind, w0.z, w0.w, (int)w1.x, w1.y, A, asc);
        if (!(ctl & OP_BPREV)) fetch(bkind, w1.z, w1.w, (int)w2.x, w2.y, X, xsc);     // B into the (dead) X registers
#pragma unroll
        for (int g = 0; g < S20_NG; ++g)
#pragma unroll
          for (int e = 0; e < 2; ++e)
          {
            osc[g][e] = asc[g][e] + xsc[g][e];
#pragma unroll
            for (int mt = 0; mt < 3; ++mt) O.v[g][mt][e] = A.v[g][mt][e] * X.v[g][mt][e];
          }
        // ---- per-site scaling: all 20*R entries of the site strictly below 2^-256 (core_partials.c:720-754)
        if (ctl & OP_SCALE)
        {
          unsigned int below[S20_NG][2];
#pragma unroll
          for (int g = 0; g < S20_NG; ++g)
#pragma unroll
            for (int e = 0; e < 2; ++e)
            {
              unsigned int b = 1u;
#pragma unroll
              for (int mt = 0; mt < 3; ++mt)
                if (8 * mt + r < S20) b &= (O.v[g][mt][e] < BPPGPU_SCALE_THRESHOLD) ? 1u : 0u;
              b &= __shfl_xor_sync(0xFFFFFFFFu, b, 4);
              b &= __shfl_xor_sync(0xFFFFFFFFu, b, 8);
              b &= __shfl_xor_sync(0xFFFFFFFFu, b, 16);
              below[g][e] = b;
            }
          if (RL > 1)
          {
            __syncthreads();
            if (r == 0)
#pragma unroll
              for (int g = 0; g < S20_NG; ++g)
#pragma unroll
                for (int e = 0; e < 2; ++e) s_flag[(pg * S20_WS + 8 * g + 2 * q + e) * RL + cat] = below[g][e];
            __syncthreads();
#pragma unroll
            for (int g = 0; g < S20_NG; ++g)
#pragma unroll
              for (int e = 0; e < 2; ++e)
              {
                unsigned int b = 1u;
                for (int c = 0; c < RL; ++c) b &= s_flag[(pg * S20_WS + 8 * g + 2 * q + e) * RL + c];
                below[g][e] = b;
              }
          }
#pragma unroll
          for (int g = 0; g < S20_NG; ++g)
#pragma unroll
            for (int e = 0; e < 2; ++e)
            {
              if (below[g][e])
              {
#pragma unroll
                for (int mt = 0; mt < 3; ++mt) O.v[g][mt][e] *= BPPGPU_SCALE_FACTOR;
                osc[g][e] += 1;
              }
              if (cat == 0 && r == 0 && validv[g][e]) H->scale[(size_t)(int)w2.w * sites + sitev[g][e]] = osc[g][e];
            }
        }
        else
        {
#pragma unroll
          for (int g = 0; g < S20_NG; ++g)
#pragma unroll
            for (int e = 0; e < 2; ++e) osc[g][e] = 0;
        }
        // ---- the CLV goes to HBM exactly once (through the tile for coalesced 256-bit stores)
        __syncwarp();
        v_to_tile(O);
        __syncwarp();
        tile_to_global(clv + ((size_t)w0.y) * S20);
        // ---- push through the edge above with DMMA; the tile already holds the operand
        if (ctl & OP_PUSH)
        {
          matvec20(pmat + ((size_t)w2.z * RL + cat) * (S20 * S20), s_tile, r, q, X);
#pragma unroll
          for (int g = 0; g < S20_NG; ++g)
#pragma unroll
            for (int e = 0; e < 2; ++e) xsc[g][e] = osc[g][e];
          if (ctl & OP_PARKA)
          {
            double * st = s_stack + (size_t)(w3.x * NW + warp) * (S20_NG * 6 * 32) + lane;
            unsigned int * ss = s_sstack + (size_t)(w3.x * NW + warp) * (S20_NG * 2 * 32) + lane;
#pragma unroll
            for (int g = 0; g < S20_NG; ++g)
#pragma unroll
              for (int e = 0; e < 2; ++e)
              {
                ss[(g * 2 + e) * 32] = xsc[g][e];
#pragma unroll
                for (int mt = 0; mt < 3; ++mt) st[((g * 3 + mt) * 2 + e) * 32] = X.v[g][mt][e];
              }
          }
        }
      }

      if (ctl & OP_ROOT)
      {
        // site term: sum_cat rw_cat * (pi . clv_cat); the dot product is reduced over the 8 row lanes
        double tr[S20_NG][2];
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
            tr[g][e] = s;
          }
        __syncthreads();
        if (r == 0)
#pragma unroll
          for (int g = 0; g < S20_NG; ++g)
#pragma unroll
            for (int e = 0; e < 2; ++e) s_exch[(pg * S20_WS + 8 * g + 2 * q + e) * RL + cat] = tr[g][e];
        __syncthreads();
        if (cat == 0 && r == 0)
        {
#pragma unroll
          for (int g = 0; g < S20_NG; ++g)
#pragma unroll
            for (int e = 0; e < 2; ++e)
            {
              double term = 0.0;
              for (int c = 0; c < RL; ++c) term += s_exch[(pg * S20_WS + 8 * g + 2 * q + e) * RL + c] * __ldg(&H->rw[c]);
              unsigned int rsc = osc[g][e];
              if (ctl & OP_EVAL) rsc = ((int)w2.w >= 0) ? osc[g][e] : 0;
              double s;
              if (prm.persite_mode == 2) s = term;
              else
              {
                s = log(term);
                if (rsc) s += (double)rsc * prm.log_threshold;
                s *= (double)__ldg(H->weights + sitev[g][e]);
              }
              if (validv[g][e])
              {
                site_sum += s;
                if (prm.persite) prm.persite[sitev[g][e]] = s;
              }
            }
        }
      }
    }

    // ---- deterministic tile reduction
    if (prm.tile_partial)
    {
      double v = site_sum;
#pragma unroll
      for (int dd = 16; dd > 0; dd >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, dd);
      __syncthreads();
      if (lane == 0) s_red[warp] = v;
      __syncthreads();
      if (tid == 0)
      {
        double acc = 0.0;
        for (int w = 0; w < NW; ++w) acc += s_red[w];
        prm.tile_partial[t] = acc;
      }
    }
    __syncthreads();
  }
}

template <int RL>
__host__ inline size_t s20_smem_bytes(int slots)
{
  constexpr int NW = S20_NT / 32, SPT = S20_TILE / RL;
  return (size_t)NW * S20_WS * S20 * 8 + (size_t)SPT * RL * 8 + (size_t)SPT * RL * 4 + 32 * 8 +
         (size_t)slots * NW * 32 * S20_NG * (6 * 8 + 2 * 4);
}

}  // namespace bppgpu
