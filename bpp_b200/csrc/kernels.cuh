// kernels.cuh -- sm_100a device code of the Felsenstein-pruning engine.
//
// Reference semantics (file:line relative to /root/reference/src):
//   P-matrix   locus.c:2325-2415 (JC69 closed form), core_pmatrix.c:674-783 (eigen form)
//   CLV update core_partials.c:585-756; association order of core_partials_avx.c:368-531
//   root lnL   core_likelihood.c:24-212, core_likelihood_avx.c:98-157; vector form :214-408
//
// Data layout in HBM (per locus, reference order so tips/inner CLVs are P*R*S contiguous doubles):
//   clv[buffer][pattern][cat][state]  pmat[idx][cat][row=parent state][col=child state]
//   scale[buffer][pattern] (u32)      tip codes[tip][pattern] (u8 for 4 states, u32 for 20)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bppgpu {

struct LocusDev
{
  double * clv;                 // inner buffers; buffer b (= clv_index - tips) at clv + b*clv_stride
  double * tip_dense;           // dense tip CLVs (pll_set_tip_clv with non-0/1 values) or nullptr
  void * tip_codes;             // packed tip state masks
  unsigned char * tip_is_dense; // [tips]
  double * pmat;                // idx at pmat + idx*R*S*S
  unsigned int * scale;         // buffer s at scale + s*sites
  unsigned int * weights;       // [sites]
  double * freqs;               // [S]
  double * rates;               // [R]
  double * rate_weights;        // [R]
  double * eigenvecs;           // [S*S]
  double * inv_eigenvecs;       // [S*S]
  double * eigenvals;           // [S]
  unsigned long long * dip_off; // diploid CSR offsets [unphased+1] or nullptr
  unsigned long long * dip_map; // diploid mapping
  unsigned long long clv_stride;
  unsigned int tips, sites, states, rate_cats;
  unsigned int clv_buffers, prob_matrices, scale_buffers, model_kind;   // model_kind 0 = JC69, 1 = eigen
  unsigned int unphased, pad0;
};

// operand kinds of a planned pruning step
enum : unsigned { SRC_TIP_PACKED = 0, SRC_TIP_DENSE = 1, SRC_HBM = 2, SRC_SLOT = 3, SRC_PREV = 4 };
enum : unsigned { CTL_SPILL_MASK = 0xFFu, CTL_ROOT = 1u << 8, CTL_EVAL_ONLY = 1u << 9 };

struct PlanOp                   // 48 bytes, uniform per CTA
{
  unsigned int dst;             // inner buffer index of the parent
  unsigned int lsrc, rsrc;      // kind << 28 | index
  unsigned int lpm, rpm;        // pmatrix indices
  int dsc, lsc, rsc;            // scaler buffer indices or -1
  unsigned int ctl;             // bits 0-7 spill slot + 1 (0 = none); CTL_ROOT; CTL_EVAL_ONLY
  int root_sc;                  // scaler buffer of the root for CTL_EVAL_ONLY
  unsigned int pad[2];
};

struct RawOp                    // == bppgpu_partial_op
{
  unsigned int parent, left, right, lpm, rpm;
  int psc, lsc, rsc;
};

#define BPPGPU_SCALE_FACTOR    115792089237316195423570985008687907853269984665640564039457584007913129639936.0
#define BPPGPU_SCALE_THRESHOLD (1.0 / BPPGPU_SCALE_FACTOR)

// ----------------------------------------------------------------------------- memory helpers
__device__ __forceinline__ void ld256_nc(const double * p, double & a, double & b, double & c, double & d)
{
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
// coherent variant: a CLV written earlier in the same kernel by the same thread may be re-read
__device__ __forceinline__ void ld256(const double * p, double & a, double & b, double & c, double & d)
{
  asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p) : "memory");
}
__device__ __forceinline__ void st256(double * p, double a, double b, double c, double d)
{
  asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// ----------------------------------------------------------------------------- P-matrix kernel
// grid.x = loci of the batch; the threads of a block stride over (op, cat, row) of their locus.
// JC69: locus.c:2390-2391 (exp form).  Eigen: core_pmatrix.c:745-771 -- expm1, temp = Vinv*expd,
// P[j][k] = delta_jk + sum_m temp[j][m]*V[m][k], m-sum sequential with separate mul/add.
__global__ void __launch_bounds__(128)
pmatrix_kernel(const LocusDev * __restrict__ loci, const unsigned int * __restrict__ batch_locus,
               const unsigned int * __restrict__ mat_off, const unsigned int * __restrict__ mat_idx,
               const double * __restrict__ mat_bl)
{
  const unsigned int bl = blockIdx.x;
  const LocusDev & L = loci[batch_locus[bl]];
  const unsigned int first = mat_off[bl], count = mat_off[bl + 1] - first;
  const unsigned int S = L.states, R = L.rate_cats;
  const unsigned int tasks = count * R * S;
  for (unsigned int t = threadIdx.x; t < tasks; t += blockDim.x)
  {
    const unsigned int j = t % S;
    const unsigned int n = (t / S) % R;
    const unsigned int m = t / (S * R);
    const double bt = mat_bl[first + m] * L.rates[n];
    double * row = L.pmat + ((size_t)mat_idx[first + m] * R + n) * S * S + (size_t)j * S;
    if (bt < 1e-100)
    {
      for (unsigned int k = 0; k < S; ++k) row[k] = (j == k) ? 1.0 : 0.0;
    }
    else if (L.model_kind == 0)
    {
      const double a = (1 + 3 * exp(-4 * bt / 3)) / 4;
      const double b = (1 - a) / 3;
      for (unsigned int k = 0; k < S; ++k) row[k] = (j == k) ? a : b;
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
}

// 20-state variant of the eigen form: expm1 hoisted into shared memory per (op, cat)
__global__ void __launch_bounds__(128)
pmatrix_kernel_wide(const LocusDev * __restrict__ loci, const unsigned int * __restrict__ batch_locus,
                    const unsigned int * __restrict__ mat_off, const unsigned int * __restrict__ mat_idx,
                    const double * __restrict__ mat_bl)
{
  extern __shared__ double s_pm[];        // V[S*S] | Vinv[S*S] | expd[S] per in-flight (op,cat) group
  const unsigned int bl = blockIdx.x;
  const LocusDev & L = loci[batch_locus[bl]];
  const unsigned int first = mat_off[bl], count = mat_off[bl + 1] - first;
  const unsigned int S = L.states, R = L.rate_cats, SS = S * S;
  double * sV = s_pm, * sVi = s_pm + SS, * sE = s_pm + 2 * SS;
  for (unsigned int t = threadIdx.x; t < SS; t += blockDim.x) { sV[t] = L.eigenvecs[t]; sVi[t] = L.inv_eigenvecs[t]; }
  __syncthreads();
  for (unsigned int g = 0; g < count * R; ++g)
  {
    const unsigned int n = g % R, m = g / R;
    const double bt = mat_bl[first + m] * L.rates[n];
    double * P = L.pmat + ((size_t)mat_idx[first + m] * R + n) * SS;
    if (bt < 1e-100)
    {
      for (unsigned int t = threadIdx.x; t < SS; t += blockDim.x) P[t] = (t / S == t % S) ? 1.0 : 0.0;
      continue;
    }
    if (threadIdx.x < S) sE[threadIdx.x] = expm1(L.eigenvals[threadIdx.x] * bt);
    __syncthreads();
    for (unsigned int t = threadIdx.x; t < SS; t += blockDim.x)
    {
      const unsigned int j = t / S, k = t % S;
      double acc = (j == k) ? 1.0 : 0.0;
      for (unsigned int mm = 0; mm < S; ++mm)
        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(sVi[j * S + mm], sE[mm]), sV[mm * S + k]));
      P[t] = acc;
    }
    __syncthreads();
  }
}

// ----------------------------------------------------------------------------- plan kernel
// One thread per locus turns the post-ordered op list (locus_update_partials' traversal) into a
// stack-machine program: a child produced earlier in the same list is taken from a register
// (SRC_PREV) or a shared-memory slot (SRC_SLOT) instead of being re-read from HBM.
__global__ void plan_kernel(const LocusDev * __restrict__ loci, const unsigned int * __restrict__ batch_locus,
                            unsigned int n_loci, const unsigned int * __restrict__ op_off,
                            const RawOp * __restrict__ ops, const unsigned int * __restrict__ root_clv,
                            const int * __restrict__ root_sc, int want_root,
                            PlanOp * __restrict__ plan, unsigned int * __restrict__ plan_count,
                            unsigned char * __restrict__ scratch, const unsigned long long * __restrict__ scratch_off,
                            int max_slots)
{
  const unsigned int bl = blockIdx.x * blockDim.x + threadIdx.x;
  if (bl >= n_loci) return;
  const LocusDev & L = loci[batch_locus[bl]];
  const unsigned int first = op_off[bl], n = op_off[bl + 1] - first;
  const RawOp * o = ops + first;
  PlanOp * p = plan + first + bl;             // one spare entry per locus for CTL_EVAL_ONLY
  unsigned char * loc = scratch + scratch_off[bl];   // [clv_buffers]: 0 = HBM, s+1 = slot s
  const unsigned int T = L.tips;
  const unsigned int rootc = want_root ? root_clv[bl] : 0xFFFFFFFFu;
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
        if (b == prev && !uses_prev) { src[c] = (SRC_PREV << 28); uses_prev = true; }
        else if (loc[b]) { src[c] = (SRC_SLOT << 28) | (unsigned)(loc[b] - 1); consumed_slots |= 1u << (loc[b] - 1); loc[b] = 0; }
        else src[c] = (SRC_HBM << 28) | b;
      }
    }
    if (prev != 0xFFFFFFFFu && !uses_prev && free_slots)
    {
      const int s = __ffs(free_slots) - 1;
      free_slots &= ~(1u << s);
      loc[prev] = (unsigned char)(s + 1);
      q.ctl |= (unsigned)(s + 1);
    }
    free_slots |= consumed_slots;
    q.lsrc = src[0]; q.rsrc = src[1];
    if (r.parent == rootc) { q.ctl |= CTL_ROOT; q.root_sc = r.psc; root_done = true; }
    p[k] = q;
    prev = q.dst;
  }
  unsigned int cnt = n;
  if (want_root && !root_done)
  {
    PlanOp q;
    q.dst = 0; q.lpm = q.rpm = 0; q.dsc = -1; q.rsc = -1; q.rsrc = 0; q.pad[0] = q.pad[1] = 0;
    q.lsc = root_sc[bl]; q.root_sc = root_sc[bl];
    q.ctl = CTL_EVAL_ONLY | CTL_ROOT;
    if (rootc < T) q.lsrc = ((L.tip_is_dense[rootc] ? SRC_TIP_DENSE : SRC_TIP_PACKED) << 28) | rootc;
    else
    {
      const unsigned int b = rootc - T;
      // the root may sit in a slot/register only if it was produced by this list, which is the
      // root_done case; here it is always HBM-resident
      q.lsrc = (SRC_HBM << 28) | b;
    }
    p[n] = q;
    cnt = n + 1;
  }
  plan_count[bl] = cnt;
}

// ----------------------------------------------------------------------------- 4-state tree kernel
template <bool EXACT>
__device__ __forceinline__ double dot4(const double2 pa, const double2 pb, const double c0, const double c1,
                                       const double c2, const double c3)
{
  if (EXACT)   // (p0+p1)+(p2+p3), separate mul/add: core_partials_avx.c:423-473
    return __dadd_rn(__dadd_rn(__dmul_rn(pa.x, c0), __dmul_rn(pa.y, c1)),
                     __dadd_rn(__dmul_rn(pb.x, c2), __dmul_rn(pb.y, c3)));
  return fma(pa.x, c0, pa.y * c1) + fma(pb.x, c2, pb.y * c3);
}

constexpr int TREE_CHUNK = 16;      // ops staged per shared-memory refill
constexpr int PM_STRIDE  = 18;      // doubles per (child, cat) matrix in shared memory (16 + 2 pad)

struct TreeParams
{
  const LocusDev * loci;
  const unsigned int * batch_locus;
  const unsigned int * tile_locus;    // batch-local locus of each tile
  const unsigned int * tile_cell0;    // first cell (pattern*R + cat) of each tile
  const unsigned int * op_off;
  const PlanOp * plan;
  const unsigned int * plan_count;
  double * tile_partial;              // per-tile weighted site-lnL sums
  double * persite;                   // optional per-site output of the (single) locus, or nullptr
  int persite_mode;                   // 1 = weighted site lnL, 2 = site likelihood (vector form)
  int n_slots;                        // shared-memory stack slots per thread
  double log_threshold;               // log(PLL_SCALE_THRESHOLD) as the host libm evaluates it
};

// One CTA = one tile of blockDim.x cells (cell = pattern*RL + cat) of one locus; it walks the whole
// planned op list of that locus.  RL = rate categories (power of two <= 32): the RL lanes of a site
// are adjacent lanes of one warp.
template <int RL, bool EXACT>
__global__ void __launch_bounds__(256, 3)
tree_kernel_s4(const TreeParams prm)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned int tid = threadIdx.x, nthr = blockDim.x;
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

  // shared memory carve-up
  PlanOp * s_plan = reinterpret_cast<PlanOp *>(smem_raw);                               // TREE_CHUNK
  double * s_pm = reinterpret_cast<double *>(smem_raw + TREE_CHUNK * sizeof(PlanOp));   // CHUNK*2*RL*PM_STRIDE
  double2 * s_stack = reinterpret_cast<double2 *>(s_pm + TREE_CHUNK * 2 * RL * PM_STRIDE);   // slots*2*nthr
  unsigned int * s_sstack = reinterpret_cast<unsigned int *>(s_stack + (size_t)prm.n_slots * 2 * nthr);
  double * s_red = reinterpret_cast<double *>(s_sstack + (size_t)prm.n_slots * nthr);   // 32 doubles
  s_red = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(s_red) + 7) & ~uintptr_t(7));

  const unsigned int ncell = L.sites * RL;
  const unsigned int cell_raw = prm.tile_cell0[tile] + tid;
  const bool valid = cell_raw < ncell;
  const unsigned int cell = valid ? cell_raw : ncell - 1;
  const unsigned int pattern = cell / RL;
  const unsigned int cat = cell % RL;
  const unsigned int lane = tid & 31u;
  const size_t cell_off = (size_t)cell * 4;

  double p0 = 0, p1 = 0, p2 = 0, p3 = 0;    // result of the previous op (SRC_PREV)
  unsigned int psc = 0;
  double site_val = 0.0;

  for (unsigned int base = 0; base < nops; base += TREE_CHUNK)
  {
    const unsigned int cn = min((unsigned)TREE_CHUNK, nops - base);
    __syncthreads();
    // stage the chunk's plan and its P-matrices
    {
      const unsigned int words = cn * (sizeof(PlanOp) / 4);
      const unsigned int * src = reinterpret_cast<const unsigned int *>(gplan + base);
      unsigned int * dst = reinterpret_cast<unsigned int *>(s_plan);
      for (unsigned int w = tid; w < words; w += nthr) dst[w] = src[w];
      const unsigned int elems = cn * 2 * RL * 16;
      for (unsigned int e = tid; e < elems; e += nthr)
      {
        const unsigned int x = e & 15u, r = (e >> 4) % RL, c = ((e >> 4) / RL) & 1u, k = (e >> 4) / (2 * RL);
        const PlanOp & q = gplan[base + k];
        if (q.ctl & CTL_EVAL_ONLY) continue;
        const unsigned int pm = c ? q.rpm : q.lpm;
        s_pm[((k * 2 + c) * RL + r) * PM_STRIDE + x] = __ldg(L.pmat + ((size_t)pm * RL + r) * 16 + x);
      }
    }
    __syncthreads();

    for (unsigned int k = 0; k < cn; ++k)
    {
      const PlanOp q = s_plan[k];
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
        if (kind == SRC_PREV) { l0 = p0; l1 = p1; l2 = p2; l3 = p3; lsc = psc; }
        else if (kind == SRC_SLOT)
        {
          const double2 a = s_stack[(idx * 2 + 0) * nthr + tid], b = s_stack[(idx * 2 + 1) * nthr + tid];
          l0 = a.x; l1 = a.y; l2 = b.x; l3 = b.y; lsc = s_sstack[idx * nthr + tid];
        }
        else if (kind == SRC_TIP_PACKED)
        {
          const unsigned int code = reinterpret_cast<const unsigned char *>(L.tip_codes)[(size_t)idx * L.sites + pattern];
          l0 = (code & 1u) ? 1.0 : 0.0; l1 = (code & 2u) ? 1.0 : 0.0; l2 = (code & 4u) ? 1.0 : 0.0; l3 = (code & 8u) ? 1.0 : 0.0;
        }
        else if (kind == SRC_TIP_DENSE) ld256_nc(L.tip_dense + (size_t)idx * L.clv_stride + cell_off, l0, l1, l2, l3);
        else
        {
          ld256(L.clv + (size_t)idx * L.clv_stride + cell_off, l0, l1, l2, l3);
          if (q.lsc >= 0) lsc = L.scale[(size_t)q.lsc * L.sites + pattern];
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
        if (kind == SRC_PREV) { r0 = p0; r1 = p1; r2 = p2; r3 = p3; rsc = psc; }
        else if (kind == SRC_SLOT)
        {
          const double2 a = s_stack[(idx * 2 + 0) * nthr + tid], b = s_stack[(idx * 2 + 1) * nthr + tid];
          r0 = a.x; r1 = a.y; r2 = b.x; r3 = b.y; rsc = s_sstack[idx * nthr + tid];
        }
        else if (kind == SRC_TIP_PACKED)
        {
          const unsigned int code = reinterpret_cast<const unsigned char *>(L.tip_codes)[(size_t)idx * L.sites + pattern];
          r0 = (code & 1u) ? 1.0 : 0.0; r1 = (code & 2u) ? 1.0 : 0.0; r2 = (code & 4u) ? 1.0 : 0.0; r3 = (code & 8u) ? 1.0 : 0.0;
        }
        else if (kind == SRC_TIP_DENSE) ld256_nc(L.tip_dense + (size_t)idx * L.clv_stride + cell_off, r0, r1, r2, r3);
        else
        {
          ld256(L.clv + (size_t)idx * L.clv_stride + cell_off, r0, r1, r2, r3);
          if (q.rsc >= 0) rsc = L.scale[(size_t)q.rsc * L.sites + pattern];
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
        // ---- per-site scaling (core_partials.c:720,739-754): all S*R entries strictly below 2^-256
        if (q.dsc >= 0)
        {
          osc = lsc + rsc;
          unsigned int below = (o0 < BPPGPU_SCALE_THRESHOLD) & (o1 < BPPGPU_SCALE_THRESHOLD) &
                               (o2 < BPPGPU_SCALE_THRESHOLD) & (o3 < BPPGPU_SCALE_THRESHOLD);
#pragma unroll
          for (int d = 1; d < RL; d <<= 1) below &= __shfl_xor_sync(0xFFFFFFFFu, below, d);
          if (below)
          {
            o0 = __dmul_rn(o0, BPPGPU_SCALE_FACTOR); o1 = __dmul_rn(o1, BPPGPU_SCALE_FACTOR);
            o2 = __dmul_rn(o2, BPPGPU_SCALE_FACTOR); o3 = __dmul_rn(o3, BPPGPU_SCALE_FACTOR);
            osc += 1;
          }
          if (valid && cat == 0) L.scale[(size_t)q.dsc * L.sites + pattern] = osc;
        }
        if (valid) st256(L.clv + (size_t)q.dst * L.clv_stride + cell_off, o0, o1, o2, o3);
      }
      p0 = o0; p1 = o1; p2 = o2; p3 = o3; psc = osc;

      if (q.ctl & CTL_ROOT)
      {
        // site likelihood: sum_j rw_j * ((pi0 c0 + pi1 c1) + (pi2 c2 + pi3 c3)), core_likelihood_avx.c:121-132
        const double f0 = __ldg(L.freqs + 0), f1 = __ldg(L.freqs + 1), f2 = __ldg(L.freqs + 2), f3 = __ldg(L.freqs + 3);
        const double tr = __dadd_rn(__dadd_rn(__dmul_rn(f0, o0), __dmul_rn(f1, o1)),
                                    __dadd_rn(__dmul_rn(f2, o2), __dmul_rn(f3, o3)));
        double term = 0.0;
#pragma unroll
        for (int j = 0; j < RL; ++j)
        {
          const double v = __shfl_sync(0xFFFFFFFFu, tr, (lane & ~(unsigned)(RL - 1)) + j);
          term = __dadd_rn(term, __dmul_rn(v, __ldg(L.rate_weights + j)));
        }
        unsigned int rs = osc;
        if (q.ctl & CTL_EVAL_ONLY) rs = (q.root_sc >= 0) ? osc : 0;
        double s;
        if (prm.persite_mode == 2) s = term;
        else
        {
          s = log(term);
          if (rs) s = __dadd_rn(s, __dmul_rn((double)rs, prm.log_threshold));
          s = __dmul_rn(s, (double)__ldg(L.weights + pattern));
        }
        if (valid && cat == 0)
        {
          site_val = s;
          if (prm.persite) prm.persite[pattern] = s;
        }
      }
    }
  }

  // deterministic tile reduction of the weighted site lnL values
  if (prm.tile_partial)
  {
    double v = site_val;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    __syncthreads();
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

// ----------------------------------------------------------------------------- generic tree kernel
// Any state count / any number of rate categories: one thread per PATTERN loops over categories and
// states, every operand comes from HBM (the plan is built with 0 slots and SRC_PREV disabled).
// This is the semantic fallback (R not a power of two, exotic state counts); the tuned paths are
// tree_kernel_s4 and the 20-state kernel.
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
  const unsigned int S = L.states, R = L.rate_cats;
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
    if (lk == SRC_TIP_PACKED)
      lcode = (S == 4) ? reinterpret_cast<const unsigned char *>(L.tip_codes)[(size_t)li * L.sites + pattern]
                       : reinterpret_cast<const unsigned int *>(L.tip_codes)[(size_t)li * L.sites + pattern];
    else lp = ((lk == SRC_TIP_DENSE) ? L.tip_dense : L.clv) + (size_t)li * L.clv_stride + (size_t)pattern * R * S;
    if (!(q.ctl & CTL_EVAL_ONLY))
    {
      if (rk == SRC_TIP_PACKED)
        rcode = (S == 4) ? reinterpret_cast<const unsigned char *>(L.tip_codes)[(size_t)ri * L.sites + pattern]
                         : reinterpret_cast<const unsigned int *>(L.tip_codes)[(size_t)ri * L.sites + pattern];
      else rp = ((rk == SRC_TIP_DENSE) ? L.tip_dense : L.clv) + (size_t)ri * L.clv_stride + (size_t)pattern * R * S;
    }
    unsigned int osc = 0;
    if (lk == SRC_HBM && q.lsc >= 0) osc += L.scale[(size_t)q.lsc * L.sites + pattern];

    if (!(q.ctl & CTL_EVAL_ONLY))
    {
      if (rk == SRC_HBM && q.rsc >= 0) osc += L.scale[(size_t)q.rsc * L.sites + pattern];
      double * out = L.clv + (size_t)q.dst * L.clv_stride + (size_t)pattern * R * S;
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
            const double lv = lp ? lp[n * S + j] : (((lcode >> j) & 1u) ? 1.0 : 0.0);
            const double rv = rp ? rp[n * S + j] : (((rcode >> j) & 1u) ? 1.0 : 0.0);
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
          if (valid) out[n * S + i] = o;
        }
      }
      if (q.dsc >= 0)
      {
        if (below)
        {
          if (valid) for (unsigned int e = 0; e < R * S; ++e) out[e] = __dmul_rn(out[e], BPPGPU_SCALE_FACTOR);
          osc += 1;
        }
        if (valid) L.scale[(size_t)q.dsc * L.sites + pattern] = osc;
      }
      else osc = 0;
    }

    if (q.ctl & CTL_ROOT)
    {
      const double * rc = (q.ctl & CTL_EVAL_ONLY) ? lp : (L.clv + (size_t)q.dst * L.clv_stride + (size_t)pattern * R * S);
      double term = 0.0;
      for (unsigned int n = 0; n < R; ++n)
      {
        double la[4] = {0, 0, 0, 0};
        for (unsigned int j = 0; j < S; ++j)
        {
          const double cv = rc ? rc[n * S + j] : (((lcode >> j) & 1u) ? 1.0 : 0.0);
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

// ----------------------------------------------------------------------------- finish kernel
// lnl[locus] = sum of its tile partials in tile order; lnl_sum = fixed-order sum over the loci.
__global__ void __launch_bounds__(1024)
finish_kernel(const double * __restrict__ tile_partial, const unsigned int * __restrict__ tile_first,
              unsigned int n_loci, double * __restrict__ lnl, double * __restrict__ lnl_sum)
{
  __shared__ double s_red[32];
  double acc = 0.0;
  for (unsigned int i = threadIdx.x; i < n_loci; i += blockDim.x)
  {
    double v = 0.0;
    for (unsigned int t = tile_first[i]; t < tile_first[i + 1]; ++t) v += tile_partial[t];
    lnl[i] = v;
    acc += v;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
  if ((threadIdx.x & 31u) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    double t = 0.0;
    for (unsigned int w = 0; w < (blockDim.x >> 5); ++w) t += s_red[w];
    *lnl_sum = t;
  }
}

// ----------------------------------------------------------------------------- diploid kernel
// locus.c:2600-2614: logl = sum_i log(mean_j lh[map[k++]]) * weight[i], one block, fixed order.
__global__ void __launch_bounds__(256)
diploid_kernel(const LocusDev * __restrict__ loci, unsigned int locus_id, const double * __restrict__ lh,
               double * __restrict__ out)
{
  __shared__ double s_red[8];
  const LocusDev & L = loci[locus_id];
  double acc = 0.0;
  for (unsigned int i = threadIdx.x; i < L.unphased; i += blockDim.x)
  {
    const unsigned long long a = L.dip_off[i], b = L.dip_off[i + 1];
    double mean = 0.0;
    for (unsigned long long k = a; k < b; ++k) mean = __dadd_rn(mean, lh[L.dip_map[k]]);
    mean = mean / (double)(b - a);
    acc += __dmul_rn(log(mean), (double)L.weights[i]);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
  if ((threadIdx.x & 31u) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    double t = 0.0;
    for (unsigned int w = 0; w < (blockDim.x >> 5); ++w) t += s_red[w];
    *out = t;
  }
}

}  // namespace bppgpu
