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
  int last_root = -1;                          // the last op that writes the root CLV carries CTL_ROOT
  for (unsigned int k = 0; k < n; ++k) if (want_root && o[k].parent == rootc) last_root = (int)k;
  for (unsigned int k = 0; k < n; ++k)
  {
    const RawOp r = o[k];
    PlanOp q;
    q.dst = r.parent - T; q.lpm = r.lpm; q.rpm = r.rpm; q.dsc = r.psc; q.lsc = r.lsc; q.rsc = r.rsc;
    q.ctl = 0; q.root_sc = -1; q.pad[0] = q.pad[1] = 0;
    q.lsrc = classify(r.left); q.rsrc = classify(r.right);
    if ((int)k == last_root) { q.ctl |= CTL_ROOT; q.root_sc = r.psc; root_done = true; }
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


// ---------------------------------------------------------------- lane-parallel planner (lists of <= 32 ops)
// lane k owns op k.  Producer / consumer links, Sethi-Ullman needs, subtree sizes and evaluation
// positions are computed with warp shuffles; only the slot / chunk bookkeeping walks the positions
// one by one (uniformly, ~20 instructions per op).  Produces exactly the program the serial planner
// below produces for the same evaluation order.
struct SmallPlanOut { unsigned int n_chunks, cnt; bool fast, simple, nohbm; };

// Evaluation order of a list of <= 32 ops, lane k = op k: DFS post-order over the forest the list forms,
// the child with the larger Sethi-Ullman need first (so that the fewest intermediate X values are alive
// at once).  Returns the lane's op, its producer links and its position.
struct SuLane { RawOp r; int kid[2]; int par; unsigned int pos; };

__device__ __forceinline__ SuLane su_order_small(const RawOp * o, unsigned int n, unsigned int T)
{
  const unsigned int FULL = 0xFFFFFFFFu;
  const unsigned int lane = threadIdx.x & 31u;
  const bool act = lane < n;
  SuLane out;
  RawOp r;
  if (act) r = o[lane];
  else { r.parent = 0xFFFFFFFFu; r.left = r.right = 0; r.lpm = r.rpm = 0; r.psc = r.lsc = r.rsc = -1; }
  const unsigned int child[2] = { r.left, r.right };
  const bool tip[2] = { act && child[0] < T, act && child[1] < T };
  // producer of each inner child: the latest earlier op that writes that buffer
  int kid[2] = { -1, -1 };
  for (unsigned int j = 0; j < n; ++j)
  {
    const unsigned int pj = __shfl_sync(FULL, r.parent, j);
    if (act && j < lane)
    {
      if (!tip[0] && pj == child[0]) kid[0] = (int)j;
      if (!tip[1] && pj == child[1]) kid[1] = (int)j;
    }
  }
  if (kid[1] == kid[0]) kid[1] = -1;
  // consumer of each op: the first op that takes its result (a pushed value has one consumer)
  int par = -1;
  for (unsigned int k2 = 0; k2 < n; ++k2)
  {
    const int a0 = __shfl_sync(FULL, kid[0], k2), a1 = __shfl_sync(FULL, kid[1], k2);
    if (par < 0 && (a0 == (int)lane || a1 == (int)lane)) par = (int)k2;
  }
#pragma unroll
  for (int c = 0; c < 2; ++c)
  {
    const int pk = __shfl_sync(FULL, par, kid[c] >= 0 ? kid[c] : 0);
    if (kid[c] >= 0 && pk != (int)lane) kid[c] = -1;
  }
  // Sethi-Ullman need and subtree size, bottom-up in index order (producers precede consumers)
  unsigned int need = 0, sz = 1;
  for (unsigned int it = 0; it < n; ++it)
  {
    const unsigned int n0 = __shfl_sync(FULL, need, kid[0] >= 0 ? kid[0] : 0), n1 = __shfl_sync(FULL, need, kid[1] >= 0 ? kid[1] : 0);
    const unsigned int s0 = __shfl_sync(FULL, sz, kid[0] >= 0 ? kid[0] : 0), s1 = __shfl_sync(FULL, sz, kid[1] >= 0 ? kid[1] : 0);
    if (lane == it)
    {
      const unsigned int m0 = kid[0] >= 0 ? n0 : 0u, m1 = kid[1] >= 0 ? n1 : 0u;
      need = (kid[0] >= 0 && kid[1] >= 0) ? max(max(m0, m1), 1u + min(m0, m1)) : max(m0, m1);
      sz = 1u + (kid[0] >= 0 ? s0 : 0u) + (kid[1] >= 0 ? s1 : 0u);
    }
  }
  int fk;                                       // the kid evaluated first: the one that needs more
  {
    const unsigned int n0 = __shfl_sync(FULL, need, kid[0] >= 0 ? kid[0] : 0), n1 = __shfl_sync(FULL, need, kid[1] >= 0 ? kid[1] : 0);
    fk = (kid[0] >= 0 && kid[1] >= 0) ? (n0 >= n1 ? kid[0] : kid[1]) : (kid[0] >= 0 ? kid[0] : kid[1]);
  }
  // evaluation positions: list roots in index order, then top-down (consumers have larger indices)
  const bool is_root = act && par < 0;
  unsigned int start = 0;
  for (unsigned int j = 0; j < n; ++j)
  {
    const unsigned int rj = __shfl_sync(FULL, is_root ? sz : 0u, j);
    if (is_root && j < lane) start += rj;
  }
  for (int it = (int)n - 1; it >= 0; --it)
  {
    const unsigned int st = __shfl_sync(FULL, start, it);
    const int fkk = __shfl_sync(FULL, fk, it);
    const unsigned int szf = __shfl_sync(FULL, sz, fkk >= 0 ? fkk : 0);
    if (par == it) start = st + ((int)lane == fkk ? 0u : szf);
  }
  out.r = r; out.kid[0] = kid[0]; out.kid[1] = kid[1]; out.par = par; out.pos = start + sz - 1;
  return out;
}

__device__ __forceinline__ SmallPlanOut
plan_small_parallel(const LocusDev & L, const RawOp * o, unsigned int n, unsigned int rootc, int rootsc, bool want_root,
                    unsigned char * blk, unsigned int chunks0, size_t cb, unsigned int cap, unsigned int lut_unit,
                    unsigned int slot_unit, int max_slots, unsigned int RL, unsigned int max_fast_tip, OpRec * rec,
                    unsigned char * s_order, unsigned int * s_push)
{
  const unsigned int FULL = 0xFFFFFFFFu;
  const unsigned int lane = threadIdx.x & 31u;
  const unsigned int T = L.tips;
  const bool act = lane < n;
  const SuLane su = su_order_small(o, n, T);
  const RawOp r = su.r;
  const unsigned int child[2] = { r.left, r.right };
  bool tip[2], dense[2];
#pragma unroll
  for (int c = 0; c < 2; ++c)
  {
    tip[c] = act && child[c] < T;
    dense[c] = tip[c] && L.tip_is_dense[child[c]] != 0;
  }
  const int kid[2] = { su.kid[0], su.kid[1] };
  const int par = su.par;
  const unsigned int pos = su.pos;
  if (act) s_order[pos] = (unsigned char)lane;
  s_push[lane] = 0;
  __syncwarp();
  // slots and chunks, position by position (uniform control flow)
  // children that need a staged P-matrix slot: packed tips (lookup table) and CLVs that this list does
  // not produce (HBM-resident children of partial updates)
  unsigned int ntip = 0;
#pragma unroll
  for (int c = 0; c < 2; ++c) if (act && ((tip[c] && !dense[c]) || (!tip[c] && kid[c] < 0))) ++ntip;
  unsigned int free_slots = (max_slots >= 32) ? FULL : ((1u << max_slots) - 1u);
  unsigned int c_idx = 0, c_nops = 0, c_ntips = 0;
  int myslot = -1;
  unsigned int mychunk = 0, myidx = 0, mylut = 0;
  for (unsigned int p = 0; p < n; ++p)
  {
    const unsigned int k = s_order[p];
    if (p > 0)
    {
      const unsigned int kp = s_order[p - 1];
      const int cons = __shfl_sync(FULL, par, kp);
      if (cons >= 0 && cons != (int)k && free_slots)
      {
        const int s = __ffs(free_slots) - 1;
        free_slots &= ~(1u << s);
        if (lane == kp) myslot = s;
      }
    }
    const int k0 = __shfl_sync(FULL, kid[0], k), k1 = __shfl_sync(FULL, kid[1], k);
    const int s0 = __shfl_sync(FULL, myslot, k0 >= 0 ? k0 : 0), s1 = __shfl_sync(FULL, myslot, k1 >= 0 ? k1 : 0);
    if (k0 >= 0 && s0 >= 0) free_slots |= 1u << s0;
    if (k1 >= 0 && s1 >= 0) free_slots |= 1u << s1;
    const unsigned int ntk = __shfl_sync(FULL, ntip, k);
    if (c_nops == (unsigned)TREE_CHUNK || c_ntips + ntk > cap)
    {
      if (lane == 0)
      {
        ChunkHdr * h = reinterpret_cast<ChunkHdr *>(blk + chunks0 + (size_t)c_idx * cb);
        h->nops = c_nops; h->ntips = c_ntips; h->pad0 = h->pad1 = 0;
      }
      ++c_idx; c_nops = 0; c_ntips = 0;
    }
    if (lane == k) { mychunk = c_idx; myidx = c_nops; mylut = c_ntips; }
    ++c_nops; c_ntips += ntk;
  }
  // operands
  const unsigned int cells_per_buf = L.sites * RL;
  unsigned int okind[2], osel[2], ooff[2], op0[2], opm[2]; int osc[2];
  int prev_child = -1;
  bool op_fast = true;
  unsigned int lutn = mylut;
#pragma unroll
  for (int c = 0; c < 2; ++c)
  {
    const int kc = kid[c] >= 0 ? kid[c] : 0;
    const unsigned int kpos = __shfl_sync(FULL, pos, kc);
    const int kslot = __shfl_sync(FULL, myslot, kc);
    const unsigned int kchunk = __shfl_sync(FULL, mychunk, kc), kidx = __shfl_sync(FULL, myidx, kc);
    opm[c] = c ? r.rpm : r.lpm; osel[c] = ooff[c] = op0[c] = 0; osc[c] = -1; okind[c] = SRC_HBM;
    if (!act) continue;
    if (tip[c])
    {
      op0[c] = child[c];
      if (dense[c]) { okind[c] = SRC_TIP_DENSE; op_fast = false; }
      else
      {
        okind[c] = SRC_TIP_PACKED;
        osel[c] = 15u | ((child[c] >> 3) << 4) | (((child[c] & 7u) * 4) << 8);
        ooff[c] = lutn * lut_unit; ++lutn;
        if (child[c] >= max_fast_tip) op_fast = false;             // its tip word is not staged
      }
    }
    else if (kid[c] >= 0 && kpos + 1 == pos && prev_child < 0) { okind[c] = SRC_PREV; prev_child = c; s_push[kid[c]] = opm[c] + 1; }
    else if (kid[c] >= 0 && kslot >= 0) { okind[c] = SRC_SLOT; op0[c] = (unsigned)kslot; ooff[c] = (unsigned)kslot * slot_unit; s_push[kid[c]] = opm[c] + 1; }
    else
    {
      op0[c] = child[c] - T; osc[c] = c ? r.rsc : r.lsc;
      if (kid[c] >= 0)
      {
        // produced by this list but neither in the register nor in a slot: re-read it with the producer's Pup
        if (kchunk == mychunk) { okind[c] = SRC_HBML; ooff[c] = kidx; s_push[kid[c]] = opm[c] + 1; }
        else { op_fast = false; ooff[c] = 0xFFFFFFFFu; }      // re-read across chunks: no staged matrix, general walker
      }
      else { ooff[c] = lutn; ++lutn; }         // HBM-resident child: its edge's P-matrix is staged in slot lutn
    }
  }
  __syncwarp();
  // the root's site log-likelihoods are summed where the root CLV is produced: by the LAST op that writes it (a list
  // may visit a node more than once; locus_root_loglikelihood evaluates the root once)
  const unsigned int rootmask = __ballot_sync(FULL, act && want_root && r.parent == rootc);
  const bool root_done = rootmask != 0;
  const bool is_root_op = act && want_root && r.parent == rootc && (lane == 31u || (rootmask >> (lane + 1u)) == 0u);
  if (act)
  {
    int ia = prev_child == 0 ? 1 : 0;
    if (prev_child < 0 && okind[1] == SRC_HBML && okind[0] != SRC_HBML) ia = 1;
    const int ib = 1 - ia;
    OpRec q;
    q.ctl = (okind[ia] << OP_AKIND_SHIFT) | (okind[ib] << OP_BKIND_SHIFT);
    q.dst_cell = (r.parent - T) * cells_per_buf; q.dsc = r.psc; q.pad = 0;
    q.a_sel = osel[ia]; q.a_off = ooff[ia]; q.a_p0 = op0[ia]; q.a_pm = opm[ia]; q.a_sc = osc[ia];
    q.b_sel = osel[ib]; q.b_off = ooff[ib]; q.b_p0 = op0[ib]; q.b_pm = opm[ib]; q.b_sc = osc[ib];
    q.park_off = 0; q.up_pm = 0;
    if (r.psc >= 0) q.ctl |= OP_SCALE;
    if (prev_child >= 0) q.ctl |= OP_BPREV;
    if (is_root_op) q.ctl |= OP_ROOT;
    if (s_push[lane]) { q.ctl |= OP_PUSH; q.up_pm = s_push[lane] - 1; }
    if (myslot >= 0) { q.ctl |= OP_PARKA; q.park_off = (unsigned)myslot * slot_unit; }
    rec[mychunk * TREE_CHUNK + myidx] = q;
  }
  bool fast = __ballot_sync(FULL, !act || op_fast) == FULL;
  const bool op_nohbm = okind[0] != SRC_HBM && okind[0] != SRC_HBML && okind[1] != SRC_HBM && okind[1] != SRC_HBML;
  const bool nohbm = __ballot_sync(FULL, !act || op_nohbm) == FULL;
  const bool simple = nohbm && __ballot_sync(FULL, !act || r.psc < 0) == FULL;
  unsigned int cnt = n;
  if (want_root && !root_done)
  {
    if (c_nops == (unsigned)TREE_CHUNK)
    {
      if (lane == 0)
      {
        ChunkHdr * h = reinterpret_cast<ChunkHdr *>(blk + chunks0 + (size_t)c_idx * cb);
        h->nops = c_nops; h->ntips = c_ntips; h->pad0 = h->pad1 = 0;
      }
      ++c_idx; c_nops = 0; c_ntips = 0;
    }
    if (lane == 0)
    {
      OpRec q;
      memset(&q, 0, sizeof(q));
      unsigned int kind;
      q.dsc = rootsc; q.a_sc = rootsc;
      if (rootc < T)
      {
        q.a_p0 = rootc;
        if (L.tip_is_dense[rootc]) kind = SRC_TIP_DENSE;
        else { kind = SRC_TIP_PACKED; q.a_sel = 15u | ((rootc >> 3) << 4) | (((rootc & 7u) * 4) << 8); }
      }
      else { kind = SRC_HBM; q.a_p0 = rootc - T; }
      q.ctl = OP_EVAL | OP_ROOT | (kind << OP_AKIND_SHIFT);
      rec[c_idx * TREE_CHUNK + c_nops] = q;
    }
    ++c_nops; cnt = n + 1; fast = false;
  }
  if (c_nops)
  {
    if (lane == 0)
    {
      ChunkHdr * h = reinterpret_cast<ChunkHdr *>(blk + chunks0 + (size_t)c_idx * cb);
      h->nops = c_nops; h->ntips = c_ntips; h->pad0 = h->pad1 = 0;
    }
    ++c_idx;
  }
  else if (c_idx == 0 && lane == 0)
  {
    // an empty list: the tree kernel still stages chunk 0 of every locus it has tiles of, so its header must say "nothing"
    ChunkHdr * h = reinterpret_cast<ChunkHdr *>(blk + chunks0);
    h->nops = 0; h->ntips = 0; h->pad0 = h->pad1 = 0;
  }
  __syncwarp();
  SmallPlanOut out;
  out.n_chunks = c_idx; out.cnt = cnt; out.fast = fast; out.simple = out.fast && simple; out.nohbm = out.fast && nohbm;
  return out;
}

// gather: per chunk, per op: Pup (if pushed) and the tipP of its packed tip children (and of HBM-resident children
// that were given a table slot: lists planned lane-parallel, hbm_slots).  A matrix set is RL*8 double2; `per` lanes
// copy one set, so a warp moves 32/per sets per step.  The op records are read from shared memory right after
// planning (srec) or from the block itself (refresh of a cached plan).
__device__ __forceinline__ void gather_block_matrices(const LocusDev & L, unsigned char * blk, unsigned int chunks0, size_t cb,
                                                      unsigned int n_chunks, unsigned int RL, unsigned int lut_unit,
                                                      bool hbm_slots, const OpRec * srec, unsigned int c_first = 0,
                                                      unsigned int c_step = 1)
{
  const unsigned int lane = threadIdx.x & 31u;
  const unsigned int per = RL * 8;                      // double2 per (op, which) matrix set; RL is a power of two
  const unsigned int per_sh = 31u - (unsigned)__clz((int)per);
  const float inv_lut_unit = 1.0f / (float)lut_unit;
  for (unsigned int c = c_first; c < n_chunks; c += c_step)       // (several warps of a big locus share the chunks)
  {
    unsigned char * ch = blk + chunks0 + (size_t)c * cb;
    const ChunkHdr hdr = *reinterpret_cast<const ChunkHdr *>(ch);
    const OpRec * cops = srec ? &srec[c * TREE_CHUNK] : reinterpret_cast<const OpRec *>(ch + sizeof(ChunkHdr));
    double * Pup = reinterpret_cast<double *>(ch + sizeof(ChunkHdr) + TREE_CHUNK * sizeof(OpRec));
    double * tipP = Pup + (size_t)TREE_CHUNK * RL * PM_STRIDE;
    const unsigned int total = hdr.nops * 3 * per;
    for (unsigned int idx = lane; idx < total; idx += 32)
    {
      const unsigned int task = idx >> per_sh, e = idx & (per - 1u);   // e: double2 index within the set
      const unsigned int k = task / 3, which = task % 3;
      const OpRec & q = cops[k];
      if (q.ctl & OP_EVAL) continue;
      unsigned int pm; double * dst;
      if (which == 0)
      {
        if (!(q.ctl & OP_PUSH)) continue;
        pm = q.up_pm; dst = Pup + (size_t)k * RL * PM_STRIDE;
      }
      else
      {
        const unsigned int kind = (q.ctl >> (which == 1 ? OP_AKIND_SHIFT : OP_BKIND_SHIFT)) & 15u;
        const unsigned int off = which == 1 ? q.a_off : q.b_off;
        unsigned int slot;
        if (kind == SRC_TIP_PACKED) slot = (unsigned int)((float)off * inv_lut_unit + 0.5f);   // off = slot * lut_unit, exact
        else if (kind == SRC_HBM && hbm_slots && off != 0xFFFFFFFFu) slot = off;
        else continue;
        pm = which == 1 ? q.a_pm : q.b_pm;
        dst = tipP + (size_t)slot * RL * PM_STRIDE;
      }
      const unsigned int r = e >> 3, x = (e & 7u) * 2;
      const double2 v = *reinterpret_cast<const double2 *>(L.pmat + ((size_t)pm * RL + r) * 16 + x);
      *reinterpret_cast<double2 *>(dst + (size_t)r * PM_STRIDE + x) = v;
    }
  }
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
                   int max_slots, unsigned int RL, unsigned int cpt, unsigned int cap,
                   const unsigned int * __restrict__ mat_off, const unsigned int * __restrict__ mat_idx,
                   const double * __restrict__ mat_bl, unsigned int max_fast_tip)
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
    for (unsigned int t = lane; t < mcount * RL; t += 32)
    {
      const unsigned int n = t % RL, m = t / RL;
      pmatrix_full4(L, mat_idx[mfirst + m], mat_bl[mfirst + m], n);
    }
    __syncwarp();
  }
  const unsigned int first = op_off[bl], n = op_off[bl + 1] - first;
  const RawOp * o = ops + first;
  unsigned char * blk = blocks + blk_off[bl];
  const size_t cb = chunk_bytes(RL);
  const unsigned int chunks0 = (unsigned int)(sizeof(LocusHdr) + rw_bytes(RL));
  const unsigned int lut_unit = lut_slot_u4((int)RL);      // uint4 units per tip table set
  const unsigned int slot_unit = 2 * cpt * TREE_NT;      // uint4 units per stack slot
  const unsigned int T = L.tips;
  unsigned int n_chunks = 0, cnt = 0;

  // small loci (the common case) are planned entirely in shared memory: raw ops, the producer map and
  // the OpRecs under construction; the block in HBM is written once, coalesced, at the end
  // SM_OPS: records of up to 4 chunks (a chunk closes after TREE_CHUNK ops or when its lookup tables are full: a
  // 31-op list over the 16 table slots of 4 or 8 rate categories can take 4); the lane-parallel planner itself
  // handles lists of up to 32 ops
  constexpr unsigned int SM_OPS = 4 * TREE_CHUNK, SM_RAW = 32, SM_BUF = 64;
  __shared__ __align__(16) RawOp s_raw[4][SM_RAW];
  __shared__ __align__(16) OpRec s_rec[4][SM_OPS];
  __shared__ unsigned int s_where[4][SM_BUF];
  __shared__ unsigned char s_slot[4][SM_BUF];
  __shared__ unsigned char s_order[4][SM_OPS];      // evaluation order (Sethi-Ullman)
  const unsigned int wib = threadIdx.x >> 5;
  // every closed chunk holds >= m ops, so a list of n (+1 eval-only) ops needs <= n/m + 1 chunks
  const unsigned int m_ops = (cap / 2 < (unsigned)TREE_CHUNK) ? (cap / 2 ? cap / 2 : 1u) : (unsigned)TREE_CHUNK;
  // (short lists over big trees too: root-path updates of 48-tip loci used to be turned away here by a limit on the
  // buffer count that only the serial planner's shared-memory tables once had, and ran on the cell-at-a-time walker)
  const bool small = (n <= SM_RAW) && (n / m_ops + 1 <= SM_OPS / TREE_CHUNK);
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

  if (small)
  {
    const SmallPlanOut sp = plan_small_parallel(L, o, n, want_root ? root_clv[bl] : 0xFFFFFFFFu, want_root ? root_sc[bl] : -1,
                                                want_root != 0, blk, chunks0, cb, cap, lut_unit, slot_unit, max_slots, RL,
                                                max_fast_tip, s_rec[wib], s_order[wib], s_where[wib]);
    n_chunks = sp.n_chunks; cnt = sp.cnt;
    if (lane == 0)
    {
      LocusHdr * H = reinterpret_cast<LocusHdr *>(blk);
      H->clv = L.clv; H->tip_dense = L.tip_dense; H->scale = L.scale;
      H->tipwords = reinterpret_cast<const unsigned int *>(L.tip_codes); H->pmat = L.pmat;
      H->clv_stride = L.clv_stride; H->sites = L.sites; H->nops = cnt; H->tip_words = L.tip_words;
      H->n_chunks = n_chunks; H->flags = (sp.fast ? HDR_FAST : 0u) | (sp.simple ? HDR_SIMPLE : 0u) | (sp.nohbm ? HDR_NOHBM : 0u) | HDR_LANEPLAN; H->pad0 = 0;
      for (int j = 0; j < 4; ++j) H->freqs[j] = L.freqs[j];
      plan_count[bl] = cnt;
    }
  }
  else if (lane == 0)
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
    // Lists of 33..PLAN_MED_OPS ops (trees of up to 129 tips: real data, the reference's frogs example has 42-60
    // sequences per locus) are ordered the same way here, serially, with the bookkeeping in thread-local arrays.
    constexpr unsigned int PLAN_MED_OPS = 128, PLAN_MED_BUF = 256;
    unsigned char order[PLAN_MED_OPS];
    unsigned char prodmap[PLAN_MED_BUF];       // 1 + op that produced the buffer in this list; later: 1 + op slot
    const bool ordered = n <= PLAN_MED_OPS && L.clv_buffers <= PLAN_MED_BUF;
    if (ordered)
    {
      signed char kid[PLAN_MED_OPS][2];
      unsigned char need[PLAN_MED_OPS];
      unsigned char consumed[PLAN_MED_OPS], done[PLAN_MED_OPS];
      for (unsigned int k = 0; k < n; ++k)
      {
        prodmap[o[k].parent - T] = 0;
        if (o[k].left >= T) prodmap[o[k].left - T] = 0;
        if (o[k].right >= T) prodmap[o[k].right - T] = 0;
        consumed[k] = done[k] = 0;
      }
      for (unsigned int k = 0; k < n; ++k)
      {
        int c0 = -1, c1 = -1;
        if (o[k].left >= T && prodmap[o[k].left - T]) c0 = prodmap[o[k].left - T] - 1;
        if (o[k].right >= T && prodmap[o[k].right - T]) c1 = prodmap[o[k].right - T] - 1;
        if (c0 >= 0 && consumed[c0]) c0 = -1;               // a value is pushed to one consumer only
        if (c1 >= 0 && (consumed[c1] || c1 == c0)) c1 = -1;
        kid[k][0] = (signed char)c0; kid[k][1] = (signed char)c1;
        if (c0 >= 0) consumed[c0] = 1;
        if (c1 >= 0) consumed[c1] = 1;
        const unsigned int n0 = c0 >= 0 ? need[c0] : 0u, n1 = c1 >= 0 ? need[c1] : 0u;
        need[k] = (unsigned char)((c0 >= 0 && c1 >= 0) ? max(max(n0, n1), 1u + min(n0, n1)) : max(n0, n1));
        prodmap[o[k].parent - T] = (unsigned char)(k + 1);
      }
      // post-order DFS from every list root (ops nobody in the list consumes), bigger need first
      unsigned int emitted = 0;
      unsigned char stack[PLAN_MED_OPS];
      for (unsigned int root = 0; root < n; ++root)
      {
        if (consumed[root]) continue;
        int sp = 0;
        stack[sp++] = (unsigned char)root;
        while (sp > 0)
        {
          const unsigned int k = stack[sp - 1];
          const int c0 = kid[k][0], c1 = kid[k][1];
          const bool p0 = c0 >= 0 && !done[c0], p1 = c1 >= 0 && !done[c1];
          if (p0 || p1)
          {
            int nxt;
            if (p0 && p1) nxt = need[c0] >= need[c1] ? c0 : c1;
            else nxt = p0 ? c0 : c1;
            stack[sp++] = (unsigned char)nxt;
          }
          else { order[emitted++] = (unsigned char)k; done[k] = 1; --sp; }
        }
      }
      for (unsigned int k = 0; k < n; ++k) prodmap[o[k].parent - T] = 0;      // reused below: op slot of the producer
    }
    const unsigned int rootc = want_root ? root_clv[bl] : 0xFFFFFFFFu;
    unsigned int free_slots = (max_slots >= 32) ? 0xFFFFFFFFu : ((1u << max_slots) - 1u);
    unsigned int prev = 0xFFFFFFFFu;        // buffer whose X sits in the register, or none
    unsigned int prev_off = 0;              // global op slot of the op that produced it
    bool root_done = false, fast = true;
    bool any_reread = false, any_scaled = false;     // an HBML operand / a scaled op: not the lean (scaled-lean) op loop
    unsigned int c_idx = 0, c_nops = 0, c_ntips = 0;
    auto close_chunk = [&]()
    {
      ChunkHdr * h = reinterpret_cast<ChunkHdr *>(blk + chunks0 + (size_t)c_idx * cb);
      h->nops = c_nops; h->ntips = c_ntips; h->pad0 = h->pad1 = 0;
      ++c_idx; c_nops = 0; c_ntips = 0;
    };
    const unsigned int cells_per_buf = L.sites * RL;
    int last_root = -1;                        // the last op that writes the root CLV carries OP_ROOT
    for (unsigned int kk = 0; kk < n; ++kk) if (want_root && o[ordered ? (unsigned int)order[kk] : kk].parent == rootc) last_root = (int)kk;
    for (unsigned int kk = 0; kk < n; ++kk)
    {
      const RawOp r = o[ordered ? (unsigned int)order[kk] : kk];
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
            if (idx >= max_fast_tip) fast = false;              // its tip word is not staged
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
            if (ordered && prodmap[b] && (unsigned)(prodmap[b] - 1) / TREE_CHUNK == c_idx)
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
      if (A.kind == SRC_HBML) any_reread = true;
      if (r.psc >= 0) any_scaled = true;
      OpRec q;
      q.ctl = (A.kind << OP_AKIND_SHIFT) | (B.kind << OP_BKIND_SHIFT);
      q.dst_cell = (r.parent - T) * cells_per_buf; q.dsc = r.psc; q.park_off = 0; q.up_pm = 0; q.pad = 0;
      q.a_sel = A.sel; q.a_off = A.off; q.a_p0 = A.p0; q.a_pm = A.pm; q.a_sc = A.sc;
      q.b_sel = B.sel; q.b_off = B.off; q.b_p0 = B.p0; q.b_pm = B.pm; q.b_sc = B.sc;
      if (r.psc >= 0) q.ctl |= OP_SCALE;
      if (prev_child >= 0) q.ctl |= OP_BPREV;
      if ((int)kk == last_root) { q.ctl |= OP_ROOT; root_done = true; }
      const unsigned int rix = c_idx * TREE_CHUNK + c_nops;
      *rec_at(rix) = q;
      ++c_nops;
      prev = r.parent - T; prev_off = rix;
      if (ordered && rix < 255u) prodmap[r.parent - T] = (unsigned char)(rix + 1);
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
    if (c_idx == 0)
    {
      ChunkHdr * h = reinterpret_cast<ChunkHdr *>(blk + chunks0);       // empty list, see plan_small_parallel
      h->nops = 0; h->ntips = 0; h->pad0 = h->pad1 = 0;
    }
    LocusHdr * H = reinterpret_cast<LocusHdr *>(blk);
    H->clv = L.clv; H->tip_dense = L.tip_dense; H->scale = L.scale;
    H->tipwords = reinterpret_cast<const unsigned int *>(L.tip_codes); H->pmat = L.pmat;
    H->clv_stride = L.clv_stride; H->sites = L.sites; H->nops = cnt; H->tip_words = L.tip_words;
    // chunk by chunk on the fast path as long as every operand is a staged tip, the register or a stack slot
    H->n_chunks = n_chunks;
    H->flags = fast ? (HDR_FAST | (any_reread ? 0u : HDR_NOHBM) | ((any_reread || any_scaled) ? 0u : HDR_SIMPLE)) : 0u;
    H->pad0 = 0;
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
  gather_block_matrices(L, blk, chunks0, cb, n_chunks, RL, lut_unit, small, small ? s_rec[wib] : nullptr);
}

// ---------------------------------------------------------------- refresh of a planned block set
// The program of a batch (op records, slots, chunks) depends only on the op lists; what changes from one proposal
// to the next with the same lists is the VALUE of the P-matrices.  When the batch still holds the planned blocks
// of the staged lists (bppgpu_batch_run twice on the same stage, or the other index parity after
// bppgpu_batch_flip_indices), the planner is skipped: this kernel rebuilds the matrices from the branch lengths
// (mat_off != nullptr) and copies them into the Pup / tipP areas of the blocks again.  WPL warps per locus (1, 2 or
// 4, all in one CTA): batches of many loci use one -- there are warps enough -- while a few hundred big trees (254
// matrices x RL categories each at 128 tips) would leave most of the GPU idle behind 250 serial warps, so the
// host gives each of those a whole CTA: the warps split the matrices, then the chunks.
template <int WPL>
__global__ void __launch_bounds__(128)
plan_refresh_blocks(const LocusDev * __restrict__ loci, const unsigned int * __restrict__ batch_locus,
                    unsigned int n_loci, unsigned char * __restrict__ blocks, const unsigned long long * __restrict__ blk_off,
                    unsigned int RL, const unsigned int * __restrict__ mat_off, const unsigned int * __restrict__ mat_idx,
                    const double * __restrict__ mat_bl)
{
  const unsigned int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned int bl = warp / WPL, sub = warp % WPL;
  const unsigned int lane = threadIdx.x & 31u;
  const bool act = bl < n_loci;                 // (no early return: the warps of a locus meet at a CTA barrier)
  if (act && mat_off)
  {
    const LocusDev & L = loci[batch_locus[bl]];
    const unsigned int mfirst = mat_off[bl], mcount = mat_off[bl + 1] - mfirst;
    for (unsigned int t = sub * 32 + lane; t < mcount * RL; t += 32 * WPL)
    {
      const unsigned int n = t % RL, m = t / RL;
      pmatrix_full4(L, mat_idx[mfirst + m], mat_bl[mfirst + m], n);
    }
  }
  if (WPL == 1) __syncwarp(); else __syncthreads();      // the matrices of the locus are written (block-scope ordering)
  if (!act) return;
  const LocusDev & L = loci[batch_locus[bl]];
  unsigned char * blk = blocks + blk_off[bl];
  const LocusHdr * H = reinterpret_cast<const LocusHdr *>(blk);
  const unsigned int chunks0 = (unsigned int)(sizeof(LocusHdr) + rw_bytes(RL));
  gather_block_matrices(L, blk, chunks0, chunk_bytes(RL), H->n_chunks, RL, lut_slot_u4((int)RL), (H->flags & HDR_LANEPLAN) != 0, nullptr,
                        sub, WPL);
}

// ---------------------------------------------------------------- device-side index flips
// SWAP_CLV_INDEX / SWAP_SCALER_INDEX / SWAP_PMAT_INDEX (locus.c:24-26) applied to EVERY inner node and EVERY edge
// of the staged step: what a whole-tree proposal does on the host before it calls the seam (prop_mixing.c:100-124).
// BPP's allocation is assumed (checked by the caller): clv_buffers = 2(T-1), prob_matrices = 2(2T-2),
// scale_buffers = 2(T-1) or 0.  One warp per locus.
__global__ void __launch_bounds__(128)
flip_indices_kernel(const LocusDev * __restrict__ loci, const unsigned int * __restrict__ batch_locus, unsigned int n_loci,
                    const unsigned int * __restrict__ op_off, RawOp * __restrict__ ops,
                    const unsigned int * __restrict__ mat_off, unsigned int * __restrict__ mat_idx,
                    unsigned int * __restrict__ root_clv, int * __restrict__ root_sc)
{
  const unsigned int bl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned int lane = threadIdx.x & 31u;
  if (bl >= n_loci) return;
  const unsigned int T = loci[batch_locus[bl]].tips, inner = T - 1, edges = 2 * T - 2;
  auto fclv = [&](unsigned int i) -> unsigned int { return i < T ? i : T + (i - 1) % (2 * T - 2); };
  auto fsc = [&](int i) -> int { return i < 0 ? i : (int)((T + (unsigned)i - 1) % (2 * T - 2)); };
  auto fpm = [&](unsigned int i) -> unsigned int { return (edges + i) % (2 * edges); };
  (void)inner;
  if (op_off)
    for (unsigned int k = op_off[bl] + lane; k < op_off[bl + 1]; k += 32)
    {
      RawOp r = ops[k];
      r.parent = fclv(r.parent); r.left = fclv(r.left); r.right = fclv(r.right);
      r.lpm = fpm(r.lpm); r.rpm = fpm(r.rpm);
      r.psc = fsc(r.psc); r.lsc = fsc(r.lsc); r.rsc = fsc(r.rsc);
      ops[k] = r;
    }
  if (mat_off)
    for (unsigned int k = mat_off[bl] + lane; k < mat_off[bl + 1]; k += 32) mat_idx[k] = fpm(mat_idx[k]);
  if (root_clv && lane == 0) { root_clv[bl] = fclv(root_clv[bl]); root_sc[bl] = fsc(root_sc[bl]); }
}

// Class of a planned batch (4 states): do ALL its loci run the lean one-chunk instantiation (HDR_SIMPLE), or all the
// scaled one-chunk one (HDR_NOHBM)?  One word per index parity, read back by the host and kept with the cached plan;
// runs on a plan that is all scaled (and not all lean) then launch the kernel that carries only that instantiation
// (tree_kernel_s4<.., SCALED_ONLY = true>).
enum : unsigned { PLAN_CLASS_KNOWN = 1u, PLAN_CLASS_LEAN = 2u, PLAN_CLASS_SCALED = 4u };
__global__ void __launch_bounds__(256)
plan_class_kernel(const unsigned char * __restrict__ blocks, const unsigned long long * __restrict__ blk_off, unsigned int n,
                  unsigned int * __restrict__ cls)
{
  // *cls starts as KNOWN | LEAN | SCALED (set by the host on the stream); a CTA that meets a locus of another class
  // clears the bit -- once: later CTAs see it gone and skip the atomic
  const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int lean = 1u, scaled = 1u;
  if (i < n)
  {
    const LocusHdr * H = reinterpret_cast<const LocusHdr *>(blocks + blk_off[i]);
    const bool one = H->n_chunks == 1;
    lean = (one && (H->flags & HDR_SIMPLE)) ? 1u : 0u;
    scaled = (one && (H->flags & HDR_NOHBM)) ? 1u : 0u;
  }
  const int all_lean = __syncthreads_and((int)lean), all_scaled = __syncthreads_and((int)scaled);
  if (threadIdx.x == 0)
  {
    const unsigned int clear = (all_lean ? 0u : PLAN_CLASS_LEAN) | (all_scaled ? 0u : PLAN_CLASS_SCALED);
    if (clear && (*reinterpret_cast<volatile unsigned int *>(cls) & clear)) atomicAnd(cls, ~clear);
  }
}

}  // namespace bppgpu
