// plan.cuh -- turn a locus' post-ordered op list (the traversal locus_update_partials walks,
// locus.c:2530-2571) into a stack-machine program.
//
// A child produced earlier in the same list is taken from a register (the previous op's result)
// or a shared-memory slot instead of being re-read from HBM; everything else is a packed tip, a
// dense tip or an HBM-resident CLV.
//
// 4-state blocks use the PUSH model: what travels in the register / slot is not the child's CLV c
// but X = P_edge . c, computed once right after c is produced (the op gets OP_PUSH and the
// P-matrix of the edge above it).  Packed tip children get a per-edge 16-entry lookup table of
// X = P_edge . bits(mask); the planner hands out the table slots and closes a chunk when the
// next op's tables would not fit.
#pragma once
#include "common.cuh"
#include "pmatrix.cuh"

namespace bppgpu {

// ---------------------------------------------------------------- flat plan (generic kernel)
// plan + op_off[bl] + bl, one spare entry per locus; every operand comes from HBM
__global__ void plan_kernel_flat(const LocusDev * __restrict__ loci, const unsigned int * __restrict__ batch_locus,
                                 unsigned int n_loci, const unsigned int * __restrict__ op_off,
                                 const RawOp * __restrict__ ops, const unsigned int * __restrict__ root_clv,
                                 const int * __restrict__ root_sc, int want_root,
                                 PlanOp * __restrict__ plan, unsigned int * __restrict__ plan_count)
{
  const unsigned int bl = blockIdx.x * blockDim.x + threadIdx.x;
  if (bl >= n_loci) return;
  const LocusDev & L = loci[batch_locus[bl]];
  const unsigned int first = op_off[bl], n = op_off[bl + 1] - first;
  const RawOp * o = ops + first;
  PlanOp * p = plan + first + bl;
  const unsigned int T = L.tips;
  const unsigned int rootc = want_root ? root_clv[bl] : 0xFFFFFFFFu;
  bool root_done = false;
  auto classify = [&](unsigned int idx) -> unsigned int
  {
    if (idx < T) return ((L.tip_is_dense[idx] ? SRC_TIP_DENSE : SRC_TIP_PACKED) << 28) | idx;
    return (SRC_HBM << 28) | (idx - T);
  };
  for (unsigned int k = 0; k < n; ++k)
  {
    const RawOp r = o[k];
    PlanOp q;
    q.dst = r.parent - T; q.lpm = r.lpm; q.rpm = r.rpm; q.dsc = r.psc; q.lsc = r.lsc; q.rsc = r.rsc;
    q.ctl = 0; q.root_sc = -1; q.pad[0] = q.pad[1] = 0;
    q.lsrc = classify(r.left); q.rsrc = classify(r.right);
    if (want_root && r.parent == rootc) { q.ctl |= CTL_ROOT; q.root_sc = r.psc; root_done = true; }
    p[k] = q;
  }
  unsigned int cnt = n;
  if (want_root && !root_done)
  {
    PlanOp q;
    q.dst = 0; q.lpm = q.rpm = 0; q.dsc = -1; q.rsc = -1; q.rsrc = 0; q.pad[0] = q.pad[1] = 0;
    q.lsc = root_sc[bl]; q.root_sc = root_sc[bl];
    q.ctl = CTL_EVAL_ONLY | CTL_ROOT;
    q.lsrc = classify(rootc);
    p[n] = q;
    cnt = n + 1;
  }
  plan_count[bl] = cnt;
}

// ---------------------------------------------------------------- staged blocks (4-state kernel)
// One WARP per locus: lane 0 plans sequentially, then all lanes gather the P-matrices the kernel
// needs (Pup of every pushed op, tipP of every packed tip child) into the block, already in the
// kernel's padded shared-memory layout, and publish the block offset of every tile of the locus.
// scratch per locus (global memory, only for loci too big for the shared-memory path):
// [unsigned int where[clv_buffers]]: 1 + op slot of the OpRec that produced the buffer and still holds its X
// in a stack slot, 0 = HBM only; [unsigned char slot_of[clv_buffers]].
struct Operand { unsigned int kind, sel, off, p0, pm; int sc; };

__global__ void __launch_bounds__(128)
plan_kernel_blocks(const LocusDev * __restrict__ loci, const unsigned int * __restrict__ batch_locus,
                   unsigned int n_loci, const unsigned int * __restrict__ op_off,
                   const RawOp * __restrict__ ops, const unsigned int * __restrict__ root_clv,
                   const int * __restrict__ root_sc, int want_root,
                   unsigned char * __restrict__ blocks, const unsigned long long * __restrict__ blk_off,
                   const unsigned int * __restrict__ tile_first, unsigned long long * __restrict__ tile_blk,
                   unsigned int * __restrict__ plan_count,
                   unsigned char * __restrict__ scratch, const unsigned long long * __restrict__ scratch_off,
                   int max_slots, unsigned int RL, unsigned int cpt,
                   const unsigned int * __restrict__ mat_off, const unsigned int * __restrict__ mat_idx,
                   const double * __restrict__ mat_bl)
{
  const unsigned int bl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned int lane = threadIdx.x & 31u;
  if (bl >= n_loci) return;
  const LocusDev & L = loci[batch_locus[bl]];
  // fused P-matrix build (mat_off != nullptr): the warp computes its locus' matrices first; the gather
  // below reads them back after the __syncwarp() that orders the warp's global writes
  if (mat_off)
  {
    const unsigned int mfirst = mat_off[bl], mcount = mat_off[bl + 1] - mfirst;
    for (unsigned int t = lane; t < mcount * RL * 4; t += 32)
    {
      const unsigned int j = t & 3u, n = (t >> 2) % RL, m = t / (4 * RL);
      pmatrix_row(L, mat_idx[mfirst + m], mat_bl[mfirst + m], n, j);
    }
    __syncwarp();
  }
  const unsigned int first = op_off[bl], n = op_off[bl + 1] - first;
  const RawOp * o = ops + first;
  unsigned char * blk = blocks + blk_off[bl];
  const size_t cb = chunk_bytes(RL);
  const unsigned int chunks0 = (unsigned int)(sizeof(LocusHdr) + rw_bytes(RL));
  const unsigned int cap = (unsigned int)lut_cap((int)RL);
  const unsigned int lut_unit = RL * (LUT_CAT / 2);      // uint4 units per tip table set
  const unsigned int slot_unit = 2 * cpt * TREE_NT;      // uint4 units per stack slot
  const unsigned int T = L.tips;
  unsigned int n_chunks = 0, cnt = 0;

  // small loci (the common case) are planned entirely in shared memory: raw ops, the producer map and
  // the OpRecs under construction; the block in HBM is written once, coalesced, at the end
  constexpr unsigned int SM_OPS = 2 * TREE_CHUNK, SM_BUF = 64;
  __shared__ __align__(16) RawOp s_raw[4][SM_OPS];
  __shared__ __align__(16) OpRec s_rec[4][SM_OPS];
  __shared__ unsigned int s_where[4][SM_BUF];
  __shared__ unsigned char s_slot[4][SM_BUF];
  __shared__ unsigned char s_prod[4][SM_BUF];       // 1 + op slot that produced the buffer in this list
  __shared__ unsigned char s_order[4][SM_OPS];      // evaluation order (Sethi-Ullman)
  __shared__ signed char s_kid[4][SM_OPS][2];       // producing op of the left / right child, or -1
  __shared__ unsigned char s_need[4][SM_OPS];
  const unsigned int wib = threadIdx.x >> 5;
  // every closed chunk holds >= m ops, so a list of n (+1 eval-only) ops needs <= n/m + 1 chunks
  const unsigned int m_ops = (cap / 2 < (unsigned)TREE_CHUNK) ? (cap / 2 ? cap / 2 : 1u) : (unsigned)TREE_CHUNK;
  const bool small = (n + 1 <= SM_OPS) && (L.clv_buffers <= SM_BUF) && (n / m_ops + 1 <= SM_OPS / TREE_CHUNK);
  if (small)
  {
    const uint4 * src = reinterpret_cast<const uint4 *>(o);
    uint4 * dst = reinterpret_cast<uint4 *>(s_raw[wib]);
    for (unsigned int w = lane; w < n * 2; w += 32) dst[w] = src[w];
    __syncwarp();
    o = s_raw[wib];
  }

  for (unsigned int t = tile_first[bl] + lane; t < tile_first[bl + 1]; t += 32)
  {
    tile_blk[2 * (size_t)t] = blk_off[bl];
    tile_blk[2 * (size_t)t + 1] = 0;
  }

  if (lane == 0)
  {
    unsigned int * where = small ? s_where[wib] : reinterpret_cast<unsigned int *>(scratch + scratch_off[bl]);
    unsigned char * slot_of = small ? s_slot[wib] : reinterpret_cast<unsigned char *>(where + L.clv_buffers);
    // OpRec r (global op slot c_idx*TREE_CHUNK + c_nops) lives in shared memory or directly in the block
    auto rec_at = [&](unsigned int r) -> OpRec *
    {
      if (small) return &s_rec[wib][r];
      return reinterpret_cast<OpRec *>(blk + chunks0 + (size_t)(r / TREE_CHUNK) * cb + sizeof(ChunkHdr)) + (r % TREE_CHUNK);
    };
    for (unsigned int k = 0; k < n; ++k)
    {
      where[o[k].parent - T] = 0;
      if (o[k].left >= T) where[o[k].left - T] = 0;
      if (o[k].right >= T) where[o[k].right - T] = 0;
    }
    // Evaluation order.  Any order that respects the dependencies gives bit-identical results, so small
    // lists are re-ordered Sethi-Ullman style (the child subtree that needs more parked values first):
    // a balanced tree of T tips then needs log2(T)-1 stack slots instead of up to T/3 in the caller's
    // left-first post-order.
    unsigned char * order = s_order[wib];
    unsigned char * prodmap = s_prod[wib];
    if (small)
    {
      signed char (*kid)[2] = s_kid[wib];
      unsigned char * need = s_need[wib];
      unsigned int consumed = 0;
      for (unsigned int k = 0; k < n; ++k)
      {
        prodmap[o[k].parent - T] = 0;
        if (o[k].left >= T) prodmap[o[k].left - T] = 0;
        if (o[k].right >= T) prodmap[o[k].right - T] = 0;
      }
      for (unsigned int k = 0; k < n; ++k)
      {
        int c0 = -1, c1 = -1;
        if (o[k].left >= T && prodmap[o[k].left - T]) c0 = prodmap[o[k].left - T] - 1;
        if (o[k].right >= T && prodmap[o[k].right - T]) c1 = prodmap[o[k].right - T] - 1;
        if (c0 >= 0 && (consumed >> c0) & 1u) c0 = -1;      // a value is pushed to one consumer only
        if (c1 >= 0 && ((consumed >> c1) & 1u || c1 == c0)) c1 = -1;
        kid[k][0] = (signed char)c0; kid[k][1] = (signed char)c1;
        if (c0 >= 0) consumed |= 1u << c0;
        if (c1 >= 0) consumed |= 1u << c1;
        const unsigned int n0 = c0 >= 0 ? need[c0] : 0u, n1 = c1 >= 0 ? need[c1] : 0u;
        need[k] = (unsigned char)((c0 >= 0 && c1 >= 0) ? max(max(n0, n1), 1u + min(n0, n1)) : max(n0, n1));
        prodmap[o[k].parent - T] = (unsigned char)(k + 1);
      }
      // post-order DFS from every list root (ops nobody in the list consumes), bigger need first
      unsigned int emitted = 0;
      unsigned char stack[SM_OPS];
      unsigned int done = 0;
      for (unsigned int root = 0; root < n; ++root)
      {
        if ((consumed >> root) & 1u) continue;
        int sp = 0;
        stack[sp++] = (unsigned char)root;
        while (sp > 0)
        {
          const unsigned int k = stack[sp - 1];
          const int c0 = kid[k][0], c1 = kid[k][1];
          const bool p0 = c0 >= 0 && !((done >> c0) & 1u), p1 = c1 >= 0 && !((done >> c1) & 1u);
          if (p0 || p1)
          {
            int nxt;
            if (p0 && p1) nxt = need[c0] >= need[c1] ? c0 : c1;
            else nxt = p0 ? c0 : c1;
            stack[sp++] = (unsigned char)nxt;
          }
          else { order[emitted++] = (unsigned char)k; done |= 1u << k; --sp; }
        }
      }
      for (unsigned int k = 0; k < n; ++k) prodmap[o[k].parent - T] = 0;      // reused below: op slot of the producer
    }
    const unsigned int rootc = want_root ? root_clv[bl] : 0xFFFFFFFFu;
    unsigned int free_slots = (max_slots >= 32) ? 0xFFFFFFFFu : ((1u << max_slots) - 1u);
    unsigned int prev = 0xFFFFFFFFu;        // buffer whose X sits in the register, or none
    unsigned int prev_off = 0;              // global op slot of the op that produced it
    bool root_done = false, fast = true;
    unsigned int c_idx = 0, c_nops = 0, c_ntips = 0;
    auto close_chunk = [&]()
    {
      ChunkHdr * h = reinterpret_cast<ChunkHdr *>(blk + chunks0 + (size_t)c_idx * cb);
      h->nops = c_nops; h->ntips = c_ntips; h->pad0 = h->pad1 = 0;
      ++c_idx; c_nops = 0; c_ntips = 0;
    };
    const unsigned int cells_per_buf = L.sites * RL;
    for (unsigned int kk = 0; kk < n; ++kk)
    {
      const RawOp r = o[small ? (unsigned int)order[kk] : kk];
      const unsigned int child[2] = { r.left, r.right };
      unsigned int ntip = 0;
      for (int c = 0; c < 2; ++c) if (child[c] < T && !L.tip_is_dense[child[c]]) ++ntip;
      if (c_nops == (unsigned)TREE_CHUNK || c_ntips + ntip > cap) close_chunk();
      Operand opd[2];
      int prev_child = -1;
      unsigned int consumed_slots = 0;
      for (int c = 0; c < 2; ++c)
      {
        const unsigned int idx = child[c];
        Operand & q = opd[c];
        q.pm = c ? r.rpm : r.lpm; q.sel = q.off = q.p0 = 0; q.sc = -1;
        if (idx < T)
        {
          q.p0 = idx;
          if (L.tip_is_dense[idx]) { q.kind = SRC_TIP_DENSE; fast = false; }
          else
          {
            q.kind = SRC_TIP_PACKED;
            q.sel = 15u | ((idx >> 3) << 4) | (((idx & 7u) * 4) << 8);
            q.off = c_ntips * lut_unit; ++c_ntips;
            if (idx >= 16) fast = false;
          }
        }
        else
        {
          const unsigned int b = idx - T;
          if (b == prev && prev_child < 0)
          {
            q.kind = SRC_PREV; prev_child = c;
            OpRec * prod = rec_at(prev_off);
            prod->ctl |= OP_PUSH; prod->up_pm = q.pm;
          }
          else if (where[b])
          {
            const unsigned int s = slot_of[b];
            q.kind = SRC_SLOT; q.p0 = s; q.off = s * slot_unit; consumed_slots |= 1u << s;
            OpRec * prod = rec_at(where[b] - 1);
            prod->ctl |= OP_PUSH; prod->up_pm = q.pm;
            where[b] = 0;
          }
          else
          {
            q.kind = SRC_HBM; q.p0 = b; q.sc = c ? r.rsc : r.lsc;
            // produced by an op of the chunk being filled: the fast path re-reads the CLV (an L2 hit)
            // and applies the producer's Pup, which is staged with this chunk
            if (small && prodmap[b] && (unsigned)(prodmap[b] - 1) / TREE_CHUNK == c_idx)
            {
              q.kind = SRC_HBML; q.off = (unsigned)(prodmap[b] - 1) % TREE_CHUNK;
              OpRec * prod = rec_at(prodmap[b] - 1);
              prod->ctl |= OP_PUSH; prod->up_pm = q.pm;
            }
          }
        }
      }
      // the previous result is not consumed by this op: its producer parks X in a free slot right after
      // the push; without a free slot it is dropped (the consumer re-reads the CLV from HBM)
      if (prev != 0xFFFFFFFFu && prev_child < 0 && free_slots)
      {
        const int s = __ffs(free_slots) - 1;
        free_slots &= ~(1u << s);
        where[prev] = prev_off + 1; slot_of[prev] = (unsigned char)s;
        OpRec * prod = rec_at(prev_off);
        prod->ctl |= OP_PARKA; prod->park_off = (unsigned)s * slot_unit;
      }
      free_slots |= consumed_slots;
      // operand A is never the register X (the product is commutative); a re-read CLV must be A
      int ia = prev_child == 0 ? 1 : 0;
      if (prev_child < 0 && opd[1].kind == SRC_HBML && opd[0].kind != SRC_HBML) ia = 1;
      const Operand & A = opd[ia];
      const Operand & B = opd[1 - ia];
      if (A.kind == SRC_HBM || B.kind == SRC_HBM || B.kind == SRC_HBML) fast = false;
      OpRec q;
      q.ctl = (A.kind << OP_AKIND_SHIFT) | (B.kind << OP_BKIND_SHIFT);
      q.dst_cell = (r.parent - T) * cells_per_buf; q.dsc = r.psc; q.park_off = 0; q.up_pm = 0; q.pad = 0;
      q.a_sel = A.sel; q.a_off = A.off; q.a_p0 = A.p0; q.a_pm = A.pm; q.a_sc = A.sc;
      q.b_sel = B.sel; q.b_off = B.off; q.b_p0 = B.p0; q.b_pm = B.pm; q.b_sc = B.sc;
      if (r.psc >= 0) q.ctl |= OP_SCALE;
      if (prev_child >= 0) q.ctl |= OP_BPREV;
      if (want_root && r.parent == rootc) { q.ctl |= OP_ROOT; root_done = true; }
      const unsigned int rix = c_idx * TREE_CHUNK + c_nops;
      *rec_at(rix) = q;
      ++c_nops;
      prev = r.parent - T; prev_off = rix;
      if (small) prodmap[r.parent - T] = (unsigned char)(rix + 1);
    }
    cnt = n;
    if (want_root && !root_done)
    {
      if (c_nops == (unsigned)TREE_CHUNK) close_chunk();
      OpRec q;
      memset(&q, 0, sizeof(q));
      unsigned int kind;
      q.dsc = root_sc[bl]; q.a_sc = root_sc[bl];
      if (rootc < T)
      {
        q.a_p0 = rootc;
        if (L.tip_is_dense[rootc]) kind = SRC_TIP_DENSE;
        else { kind = SRC_TIP_PACKED; q.a_sel = 15u | ((rootc >> 3) << 4) | (((rootc & 7u) * 4) << 8); }
      }
      else { kind = SRC_HBM; q.a_p0 = rootc - T; }
      q.ctl = OP_EVAL | OP_ROOT | (kind << OP_AKIND_SHIFT);
      fast = false;
      *rec_at(c_idx * TREE_CHUNK + c_nops) = q;
      ++c_nops;
      cnt = n + 1;
    }
    if (c_nops) close_chunk();
    n_chunks = c_idx;
    LocusHdr * H = reinterpret_cast<LocusHdr *>(blk);
    H->clv = L.clv; H->tip_dense = L.tip_dense; H->scale = L.scale;
    H->tipwords = reinterpret_cast<const unsigned int *>(L.tip_codes); H->pmat = L.pmat;
    H->clv_stride = L.clv_stride; H->sites = L.sites; H->nops = cnt; H->tip_words = L.tip_words;
    H->n_chunks = n_chunks; H->flags = (fast && n_chunks == 1) ? HDR_FAST : 0u; H->pad0 = 0;
    for (int j = 0; j < 4; ++j) H->freqs[j] = L.freqs[j];
    plan_count[bl] = cnt;
  }
  n_chunks = __shfl_sync(0xFFFFFFFFu, n_chunks, 0);
  __syncwarp();
  if (small)
  {
    // OpRecs: shared -> block, 16 bytes per lane and step
    for (unsigned int w = lane; w < n_chunks * TREE_CHUNK * 4; w += 32)
    {
      const unsigned int r = w >> 2;
      uint4 * dst = reinterpret_cast<uint4 *>(blk + chunks0 + (size_t)(r / TREE_CHUNK) * cb + sizeof(ChunkHdr)) + (r % TREE_CHUNK) * 4 + (w & 3u);
      *dst = reinterpret_cast<const uint4 *>(s_rec[wib])[w];
    }
  }
  double * rw = reinterpret_cast<double *>(blk + sizeof(LocusHdr));
  for (unsigned int j = lane; j < RL; j += 32) rw[j] = L.rate_weights[j];
  // gather: per chunk, per op: Pup (if pushed) and the tipP of its packed tip children
  const unsigned int mat = RL * 16;                     // doubles per matrix
  for (unsigned int c = 0; c < n_chunks; ++c)
  {
    unsigned char * ch = blk + chunks0 + (size_t)c * cb;
    const ChunkHdr hdr = *reinterpret_cast<const ChunkHdr *>(ch);
    const OpRec * cops = small ? &s_rec[wib][c * TREE_CHUNK] : reinterpret_cast<const OpRec *>(ch + sizeof(ChunkHdr));
    double * Pup = reinterpret_cast<double *>(ch + sizeof(ChunkHdr) + TREE_CHUNK * sizeof(OpRec));
    double * tipP = Pup + (size_t)TREE_CHUNK * RL * PM_STRIDE;
    for (unsigned int e = lane; e < hdr.nops * 3 * mat; e += 32)
    {
      const unsigned int k = e / (3 * mat), w = e % (3 * mat);
      const unsigned int which = w / mat, r = (w / 16) % RL, x = w & 15u;
      const OpRec & q = cops[k];
      if (q.ctl & OP_EVAL) continue;
      if (which == 0)
      {
        if (q.ctl & OP_PUSH) Pup[((size_t)k * RL + r) * PM_STRIDE + x] = L.pmat[((size_t)q.up_pm * RL + r) * 16 + x];
      }
      else
      {
        const unsigned int kind = (q.ctl >> (which == 1 ? OP_AKIND_SHIFT : OP_BKIND_SHIFT)) & 15u;
        if (kind == SRC_TIP_PACKED)
        {
          const unsigned int s = (which == 1 ? q.a_off : q.b_off) / lut_unit;
          const unsigned int pm = which == 1 ? q.a_pm : q.b_pm;
          tipP[((size_t)s * RL + r) * PM_STRIDE + x] = L.pmat[((size_t)pm * RL + r) * 16 + x];
        }
      }
    }
  }
}

}  // namespace bppgpu
