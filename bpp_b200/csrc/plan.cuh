// plan.cuh -- turn a locus' post-ordered op list (the traversal locus_update_partials walks,
// locus.c:2530-2571) into a stack-machine program: a child produced earlier in the same list is
// taken from a register (SRC_PREV) or a shared-memory slot (SRC_SLOT) instead of being re-read
// from HBM; everything else is a packed tip, a dense tip or an HBM-resident CLV.
#pragma once
#include "common.cuh"

namespace bppgpu {

// Sequential planner (one thread per locus).  emit(k, op) stores op k; returns the op count
// (n, or n+1 when a CTL_EVAL_ONLY op for a root that this list does not produce is appended).
template <class Emit>
__device__ __forceinline__ unsigned int plan_locus(const LocusDev & L, const RawOp * __restrict__ o, unsigned int n,
                                                   unsigned int rootc, int rootsc, bool want_root,
                                                   unsigned char * __restrict__ loc, int max_slots, bool allow_prev,
                                                   Emit emit)
{
  const unsigned int T = L.tips;
  for (unsigned int k = 0; k < n; ++k)
  {
    loc[o[k].parent - T] = 0;
    if (o[k].left >= T) loc[o[k].left - T] = 0;
    if (o[k].right >= T) loc[o[k].right - T] = 0;
  }
  unsigned int free_slots = (max_slots >= 32) ? 0xFFFFFFFFu : ((1u << max_slots) - 1u);
  unsigned int prev = 0xFFFFFFFFu;
  bool root_done = false;
  for (unsigned int k = 0; k < n; ++k)
  {
    const RawOp r = o[k];
    PlanOp q;
    q.dst = r.parent - T; q.lpm = r.lpm; q.rpm = r.rpm;
    q.dsc = r.psc; q.lsc = r.lsc; q.rsc = r.rsc; q.ctl = 0; q.root_sc = -1; q.pad[0] = q.pad[1] = 0;
    unsigned int src[2]; unsigned int consumed_slots = 0; bool uses_prev = false;
    const unsigned int child[2] = { r.left, r.right };
    for (int c = 0; c < 2; ++c)
    {
      const unsigned int idx = child[c];
      if (idx < T) src[c] = ((L.tip_is_dense[idx] ? SRC_TIP_DENSE : SRC_TIP_PACKED) << 28) | idx;
      else
      {
        const unsigned int b = idx - T;
        if (allow_prev && b == prev && !uses_prev) { src[c] = (SRC_PREV << 28); uses_prev = true; }
        else if (loc[b]) { src[c] = (SRC_SLOT << 28) | (unsigned)(loc[b] - 1); consumed_slots |= 1u << (loc[b] - 1); loc[b] = 0; }
        else src[c] = (SRC_HBM << 28) | b;
      }
    }
    // the previous result is not consumed by this op: park it in a free slot (else it stays HBM-only)
    if (prev != 0xFFFFFFFFu && !uses_prev && free_slots)
    {
      const int s = __ffs(free_slots) - 1;
      free_slots &= ~(1u << s);
      loc[prev] = (unsigned char)(s + 1);
      q.ctl |= (unsigned)(s + 1);
    }
    free_slots |= consumed_slots;
    q.lsrc = src[0]; q.rsrc = src[1];
    if (want_root && r.parent == rootc) { q.ctl |= CTL_ROOT; q.root_sc = r.psc; root_done = true; }
    emit(k, q);
    prev = q.dst;
  }
  if (want_root && !root_done)
  {
    PlanOp q;
    q.dst = 0; q.lpm = q.rpm = 0; q.dsc = -1; q.rsc = -1; q.rsrc = 0; q.pad[0] = q.pad[1] = 0;
    q.lsc = rootsc; q.root_sc = rootsc;
    q.ctl = CTL_EVAL_ONLY | CTL_ROOT;
    if (rootc < T) q.lsrc = ((L.tip_is_dense[rootc] ? SRC_TIP_DENSE : SRC_TIP_PACKED) << 28) | rootc;
    else q.lsrc = (SRC_HBM << 28) | (rootc - T);       // not produced by this list: HBM-resident
    emit(n, q);
    return n + 1;
  }
  return n;
}

// flat plan for the generic kernel: plan + op_off[bl] + bl, one spare entry per locus
__global__ void plan_kernel_flat(const LocusDev * __restrict__ loci, const unsigned int * __restrict__ batch_locus,
                                 unsigned int n_loci, const unsigned int * __restrict__ op_off,
                                 const RawOp * __restrict__ ops, const unsigned int * __restrict__ root_clv,
                                 const int * __restrict__ root_sc, int want_root,
                                 PlanOp * __restrict__ plan, unsigned int * __restrict__ plan_count,
                                 unsigned char * __restrict__ scratch, const unsigned long long * __restrict__ scratch_off)
{
  const unsigned int bl = blockIdx.x * blockDim.x + threadIdx.x;
  if (bl >= n_loci) return;
  const LocusDev & L = loci[batch_locus[bl]];
  const unsigned int first = op_off[bl], n = op_off[bl + 1] - first;
  PlanOp * p = plan + first + bl;
  plan_count[bl] = plan_locus(L, ops + first, n, want_root ? root_clv[bl] : 0xFFFFFFFFu, want_root ? root_sc[bl] : -1,
                              want_root != 0, scratch + scratch_off[bl], 0, false,
                              [p](unsigned int k, const PlanOp & q) { p[k] = q; });
}

// staged blocks for the 4-state kernel: one WARP per locus; lane 0 plans, all lanes gather the
// P-matrices of every op into the block in the kernel's shared-memory layout
__global__ void __launch_bounds__(128)
plan_kernel_blocks(const LocusDev * __restrict__ loci, const unsigned int * __restrict__ batch_locus,
                   unsigned int n_loci, const unsigned int * __restrict__ op_off,
                   const RawOp * __restrict__ ops, const unsigned int * __restrict__ root_clv,
                   const int * __restrict__ root_sc, int want_root,
                   unsigned char * __restrict__ blocks, const unsigned long long * __restrict__ blk_off,
                   unsigned int * __restrict__ plan_count,
                   unsigned char * __restrict__ scratch, const unsigned long long * __restrict__ scratch_off,
                   int max_slots, unsigned int RL)
{
  const unsigned int bl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned int lane = threadIdx.x & 31u;
  if (bl >= n_loci) return;
  const LocusDev & L = loci[batch_locus[bl]];
  const unsigned int first = op_off[bl], n = op_off[bl + 1] - first;
  unsigned char * blk = blocks + blk_off[bl];
  const size_t cb = chunk_bytes(RL);
  unsigned char * chunks = blk + sizeof(LocusHdr) + rw_bytes(RL);
  unsigned int cnt = 0;
  if (lane == 0)
  {
    cnt = plan_locus(L, ops + first, n, want_root ? root_clv[bl] : 0xFFFFFFFFu, want_root ? root_sc[bl] : -1,
                     want_root != 0, scratch + scratch_off[bl], max_slots, true,
                     [chunks, cb](unsigned int k, const PlanOp & q)
                     { reinterpret_cast<PlanOp *>(chunks + (size_t)(k / TREE_CHUNK) * cb)[k % TREE_CHUNK] = q; });
    LocusHdr * H = reinterpret_cast<LocusHdr *>(blk);
    H->clv = L.clv; H->tip_dense = L.tip_dense; H->scale = L.scale;
    H->tipwords = reinterpret_cast<const unsigned int *>(L.tip_codes);
    H->clv_stride = L.clv_stride; H->sites = L.sites; H->nops = cnt; H->tip_words = L.tip_words;
    H->n_chunks = (cnt + TREE_CHUNK - 1) / TREE_CHUNK;
    for (int j = 0; j < 4; ++j) H->freqs[j] = L.freqs[j];
    plan_count[bl] = cnt;
  }
  cnt = __shfl_sync(0xFFFFFFFFu, cnt, 0);
  __syncwarp();
  double * rw = reinterpret_cast<double *>(blk + sizeof(LocusHdr));
  for (unsigned int j = lane; j < RL; j += 32) rw[j] = L.rate_weights[j];
  // gather: element e -> (op k, child c, cat r, entry x)
  const unsigned int per_op = 2 * RL * 16;
  for (unsigned int e = lane; e < cnt * per_op; e += 32)
  {
    const unsigned int k = e / per_op, w = e % per_op;
    const unsigned int c = w / (RL * 16), r = (w / 16) % RL, x = w & 15u;
    unsigned char * ch = chunks + (size_t)(k / TREE_CHUNK) * cb;
    const PlanOp & q = reinterpret_cast<const PlanOp *>(ch)[k % TREE_CHUNK];
    if (q.ctl & CTL_EVAL_ONLY) continue;
    double * P = reinterpret_cast<double *>(ch + TREE_CHUNK * sizeof(PlanOp));
    P[(((k % TREE_CHUNK) * 2 + c) * RL + r) * PM_STRIDE + x] = L.pmat[((size_t)(c ? q.rpm : q.lpm) * RL + r) * 16 + x];
  }
}

}  // namespace bppgpu
