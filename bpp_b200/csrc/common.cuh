// common.cuh -- shared device-side types of the sm_100a Felsenstein-pruning engine.
//
// Reference semantics (file:line relative to /root/reference/src):
//   P-matrix   locus.c:2325-2415 (JC69 closed form), core_pmatrix.c:674-783 (eigen form)
//   CLV update core_partials.c:585-756; association order of core_partials_avx.c:368-531
//   root lnL   core_likelihood.c:24-212, core_likelihood_avx.c:98-157; vector form :214-408
//
// Data layout in HBM (per locus; CLVs in the reference's order so a CLV is P*R*S contiguous doubles; 20-state loci
// with several categories keep clv[buffer][cat][pattern][state], see LocusDev::site_stride):
//   clv[buffer][pattern][cat][state]     pmat[idx][cat][row = parent state][col = child state]
//   scale[buffer][pattern] (u32)         weights[pattern] (u32)
//   tips, 4 states : 4-bit state masks, 8 tips per u32 word: tipwords[pattern][tip/8] >> 4*(tip%8)
//   tips, >4 states: u32 masks tip_codes[tip][pattern]
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bppgpu {

struct LocusDev
{
  double * clv;                 // inner buffers; buffer b (= clv_index - tips) at clv + b*clv_stride
  double * tip_dense;           // dense tip CLVs (pll_set_tip_clv with non-0/1 values) or nullptr
  void * tip_codes;             // packed tip state masks (layout above)
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
  unsigned int clv_buffers, prob_matrices, scale_buffers, model_kind;   // MODEL_* below
  unsigned int unphased, tip_words;                                     // tip_words = ceil(tips/8) (4 states)
  // > 4 states: column ids of the tips (0..S-1 = one-hot state, S.. = ambiguity mask colmask[id-S])
  unsigned char * tip_cols;     // [tips][cols_pitch], cols_pitch = sites rounded up to 16 (padding = column 0)
  unsigned int * colmask;       // [4] ambiguity masks
  unsigned int n_ext_cols, cols_pitch;
  unsigned int * dip_weights;   // diploid loci: weights of the unphased sites [unphased]
  double * subst;               // [S(S-1)/2] substitution parameters (closed-form DNA models read qrates here)
  // element (site, cat, state) of a CLV buffer or dense tip sits at site * site_stride + cat * cat_stride + state:
  // site-major (the reference's order: R*S, S) everywhere except 20-state loci with 2..8 categories, which are
  // category-major (S, sites*S) -- the 20-state tree kernel works on one category at a time and then writes whole
  // site blocks contiguously (the host never reads inner CLVs except through bppgpu_get_clv, which transposes)
  unsigned int site_stride, cat_stride;
};

// how the P-matrices of a locus are built (locus_update_matrices dispatch, locus.c:2417-2479)
enum : unsigned { MODEL_JC69 = 0, MODEL_EIGEN = 1, MODEL_K80 = 2, MODEL_F81 = 3, MODEL_HKY = 4, MODEL_T92 = 5,
                  MODEL_TN93 = 6, MODEL_F84 = 7 };

// operand kinds of a planned pruning step
enum : unsigned { SRC_TIP_PACKED = 0, SRC_TIP_DENSE = 1, SRC_HBM = 2, SRC_SLOT = 3, SRC_PREV = 4,
                  SRC_HBML = 5 };   // HBM-resident CLV produced by an op of the same staged chunk: its edge's
                                    // P-matrix is the producer's Pup, already in shared memory (4-state fast path)
enum : unsigned { CTL_ROOT = 1u << 8, CTL_EVAL_ONLY = 1u << 9 };

struct PlanOp                   // flat plan of the generic kernel, 48 bytes
{
  unsigned int dst;             // inner buffer index of the parent
  unsigned int lsrc, rsrc;      // kind << 28 | index
  unsigned int lpm, rpm;        // pmatrix indices
  int dsc, lsc, rsc;            // scaler buffer indices or -1
  unsigned int ctl;             // CTL_ROOT; CTL_EVAL_ONLY
  int root_sc;                  // scaler buffer of the root for CTL_EVAL_ONLY
  unsigned int pad[2];
};

struct RawOp                    // == bppgpu_partial_op
{
  unsigned int parent, left, right, lpm, rpm;
  int psc, lsc, rsc;
};

// ---- staged per-locus block of the 4-state tree kernel (built by plan_kernel_blocks) ----------
// [LocusHdr 128 B][rate_weights RL doubles, padded to 16 B][chunk 0][chunk 1]...
// chunk = [ChunkHdr 16 B][OpRec ops[TREE_CHUNK]][double Pup[TREE_CHUNK][RL][PM_STRIDE]]
//         [double tipP[lut_cap(RL)][RL][PM_STRIDE]]
// Pup[k]  = P-matrix of the edge ABOVE the node op k computes (the consumer's lpm/rpm);
// tipP[s] = P-matrix of the edge above the packed tip child that was given LUT slot s.
#ifndef BPPGPU_TREE_NT
#define BPPGPU_TREE_NT 256
#endif
#ifndef BPPGPU_S4_CTAS2
#define BPPGPU_S4_CTAS2 (BPPGPU_TREE_NT > 256 ? 1 : 2 * (256 / BPPGPU_TREE_NT))   // CTAs per SM targeted by the 2-cells-per-thread kernel
#endif
constexpr int TREE_NT    = BPPGPU_TREE_NT;     // threads (= cells) per tile of the 4-state kernel
// CTAs per SM the 4-state kernel is compiled for, by cells per thread (register budget 85 / 128 / 255)
__host__ __device__ constexpr int s4_ctas_per_sm(int cpt)
{
  return cpt == 2 ? BPPGPU_S4_CTAS2 : (TREE_NT > 256 ? 1 : (cpt == 1 ? 3 : 1) * (256 / TREE_NT));
}
constexpr int TREE_CHUNK = 16;      // ops per staged chunk
constexpr int S4_MAX_TIP_WORDS = 16; // packed tip words (8 tips each) the 4-state fast path stages per cell: 128 tips
constexpr int PM_STRIDE  = 18;      // doubles per (matrix, cat) in shared memory (16 + 2 pad: the RL
                                    // categories of a site land in different banks)
constexpr int LUT_ROW    = 6;       // doubles per state-mask row of a tip lookup table (4 + 2 pad: the
                                    // one-hot masks 1,2,4,8 and 15 fall into different bank groups)
constexpr int LUT_CAT    = 16 * LUT_ROW + 2;   // doubles per (tip child, cat) table
__host__ __device__ constexpr int lut_cap(int RL) { return RL <= 2 ? 32 : 16; }   // tip children per chunk
// Layout of the tip lookup tables in shared memory (uint4 = 16-byte units; an entry X = P_edge . bits(mask) is two
// uint4: states 0,1 and states 2,3).
//   RL <= 2: [slot][cat][mask] with padded rows (LUT_ROW) -- a quarter warp holds 8 or 4 different sites of the same
//            one or two categories, the padding spreads their masks over the banks.
//   RL >= 4: [slot][replica][mask][cat], no padding.  A quarter warp holds TWO sites x 4 categories (or one site x 8):
//            with one table the two sites hit the same banks whenever their masks differ (31 % of all shared-memory
//            wavefronts of the round-1 kernel were such replays).  Replica B holds the same entries with the two
//            halves exchanged; the odd site of a quarter warp (lane bit 2) reads B and issues its loads in the
//            opposite order, so every load instruction of a quarter warp covers eight different 16-byte bank groups,
//            whatever the masks are.
//   RL == 8: [slot][mask][cat], ONE table.  A quarter warp is one site x 8 categories, i.e. one row of 256 bytes; the
//            lanes of categories 4..7 are exactly the lanes with bit 2 set, so instead of a replica their entries are
//            STORED with the halves exchanged and they fetch in the opposite order: conflict-free at half the space,
//            which is what lets 16 tip children (a whole 16-tip tree) share a chunk as with 4 categories -- with
//            5 to 8 categories a 16-tip tree used to take three chunks of 8 table slots.
__host__ __device__ constexpr unsigned int lut_row_u4(int RL)  { return RL >= 4 ? 2u * (unsigned)RL : (unsigned)(LUT_ROW / 2); }
__host__ __device__ constexpr unsigned int lut_cat_u4(int RL)  { return RL >= 4 ? 2u : (unsigned)(LUT_CAT / 2); }
__host__ __device__ constexpr unsigned int lut_rep_u4(int RL)  { return RL == 4 ? 32u * (unsigned)RL : 0u; }
__host__ __device__ constexpr unsigned int lut_slot_u4(int RL) { return RL == 4 ? 64u * (unsigned)RL : (RL == 8 ? 32u * (unsigned)RL : (unsigned)RL * (LUT_CAT / 2)); }

struct ChunkHdr { unsigned int nops, ntips, pad0, pad1; };

// pre-decoded op of the 4-state kernel (64 bytes = 4 x uint4, uniform per CTA).
// Operand A is never the register-resident X; operand B may be (OP_BPREV).
//   fast operands: packed tip whose word is register-resident (tip < 16) or a stack slot:
//     sel = mask (15 tip / 0 slot) | word index << 4 | nibble shift << 8
//     off = tip: lookup-table offset; slot: stack offset (uint4 units, relative to the thread's base)
//   cold operands (HBM-resident CLV, dense tip, tip >= 16): kind / p0 / pm / sc
enum : unsigned { OP_ROOT = 1u << 0, OP_EVAL = 1u << 1, OP_PUSH = 1u << 2, OP_BPREV = 1u << 3, OP_SCALE = 1u << 4,
                  OP_PARKA = 1u << 5,       // after the push, park X in stack slot park_off
                  OP_AKIND_SHIFT = 8, OP_BKIND_SHIFT = 12 };
struct OpRec
{
  unsigned int ctl;             // OP_* flags | a_kind << 8 | b_kind << 12
  unsigned int dst_cell;        // parent buffer offset in cells (32-byte units): dst * sites * RL
  unsigned int a_sel, a_off;
  unsigned int b_sel, b_off;
  int dsc;                      // parent scaler buffer or -1 (OP_EVAL: the root's scaler buffer)
  unsigned int park_off;        // OP_PARKA: stack offset of the slot
  unsigned int a_p0, a_pm; int a_sc; unsigned int up_pm;     // up_pm: pmatrix of the edge above dst (OP_PUSH)
  unsigned int b_p0, b_pm; int b_sc; unsigned int pad;
};
static_assert(sizeof(OpRec) == 64, "OpRec must be 64 bytes");

enum : unsigned { HDR_FAST = 1u,     // every op of the locus uses fast operands only (any number of chunks)
                  HDR_SIMPLE = 2u,   // ... and none is HBM-class or scaled: the lean instantiation of tile_fast
                  HDR_NOHBM = 4u,    // ... and none is HBM-class (scaling allowed): full passes with scale buffers
                  HDR_LANEPLAN = 8u };  // planned lane-parallel: HBM-resident children own a staged P-matrix slot

struct LocusHdr
{
  double * clv;
  double * tip_dense;
  unsigned int * scale;
  const unsigned int * tipwords;
  const double * pmat;
  unsigned long long clv_stride;
  unsigned int sites, nops, tip_words, n_chunks;
  double freqs[4];
  unsigned int flags, pad0;
  double pad[3];
};
static_assert(sizeof(LocusHdr) == 128, "LocusHdr must be 128 bytes");

__host__ __device__ inline size_t rw_bytes(unsigned RL) { return ((size_t)RL * 8 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t chunk_bytes(unsigned RL)
{
  return sizeof(ChunkHdr) + (size_t)TREE_CHUNK * sizeof(OpRec) + (size_t)TREE_CHUNK * RL * PM_STRIDE * 8 +
         (size_t)lut_cap((int)RL) * RL * PM_STRIDE * 8;
}
// upper bound on the chunks of a list of nops ops: a chunk closes after TREE_CHUNK ops or when the
// next op's tip children no longer fit its lookup tables
// cap = the launch's tip-slot capacity (<= lut_cap(RL)): every closed chunk holds at least min(cap/2, TREE_CHUNK) ops
__host__ __device__ inline unsigned int max_chunks(unsigned cap, unsigned nops)
{
  const unsigned m = cap / 2 < (unsigned)TREE_CHUNK ? (cap / 2 ? cap / 2 : 1) : (unsigned)TREE_CHUNK;
  return (nops + m - 1) / m + 1;
}
__host__ __device__ inline size_t block_bytes(unsigned RL, unsigned nops_max, unsigned cap)
{
  return sizeof(LocusHdr) + rw_bytes(RL) + (size_t)max_chunks(cap, nops_max) * chunk_bytes(RL);
}

// one tile of blockDim cells (cell = pattern*RL + cat) of one locus; static per batch
struct TileDesc
{
  const unsigned int * tipwords;   // the locus' packed tips
  const unsigned int * weights;    // the locus' pattern weights
  unsigned int locus;              // batch-local locus
  unsigned int cell0;              // first cell of the tile
  unsigned int tip_words;
  unsigned int ncell;              // sites * RL
};                                 // a tile covers TREE_NT * CPT cells: thread tid owns cells cell0 + perm(tid) + j*TREE_NT
static_assert(sizeof(TileDesc) == 32, "TileDesc must be 32 bytes");

struct TreeParams
{
  const LocusDev * loci;
  const unsigned int * batch_locus;
  const unsigned int * tile_locus;    // generic kernel: batch-local locus of each tile
  const unsigned int * tile_cell0;    // generic kernel: first pattern of each tile
  const unsigned int * op_off;
  const PlanOp * plan;                // generic kernel: flat plan
  const unsigned int * plan_count;
  const TileDesc * tiles;             // 4-state kernel
  const unsigned char * blocks;       // 4-state kernel: staged per-locus blocks
  const unsigned long long * tile_blk;// byte offset of the block of each tile's locus (written by the planner)
  unsigned int n_tiles;
  double * tile_partial;              // per-tile weighted site-lnL sums
  double * persite;                   // optional per-site output of the (single) locus, or nullptr
  int persite_mode;                   // 1 = weighted site lnL, 2 = site likelihood (vector form)
  int n_slots;                        // shared-memory stack slots per cell
  unsigned int tip_words;             // 4-state kernel: packed tip words staged per cell (<= S4_MAX_TIP_WORDS)
  unsigned int lut_cap;               // tip-slot capacity of the launch (4-state kernel), <= lut_cap(RL);
                                      // 20-state category-major kernel: staged P-matrix capacity
  double * rootdot;                   // 20-state category-major kernel: pi . clv_root per (locus, category, site)
  const unsigned long long * site_off; // ... and the first site of each batch locus in it (prefix sums)
  unsigned int max_tips;              // ... and the largest tip count of the batch (tip-column staging)
  double log_threshold;               // log(PLL_SCALE_THRESHOLD) as the host libm evaluates it
};

#define BPPGPU_SCALE_FACTOR    115792089237316195423570985008687907853269984665640564039457584007913129639936.0
#define BPPGPU_SCALE_THRESHOLD (1.0 / BPPGPU_SCALE_FACTOR)

// ----------------------------------------------------------------------------- memory helpers
__device__ __forceinline__ void prefetch_l2(const void * p)
{
  asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}
__device__ __forceinline__ void ld256_nc(const double * p, double & a, double & b, double & c, double & d)
{
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
// prefetch load: volatile so that the compiler issues it where it is written (a plain __ldg whose result
// is only used after a long loop gets sunk below the loop, which turns the prefetch into a stall)
__device__ __forceinline__ unsigned int ld_u32_prefetch(const unsigned int * p)
{
  unsigned int v;
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
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
__device__ __forceinline__ void st128(double * p, double a, double b)
{
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" :: "l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void cp_async16(void * smem_dst, const void * gsrc)
{
  const unsigned int s = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void * smem_dst, const void * gsrc)
{
  const unsigned int s = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(s), "l"(gsrc) : "memory");
}
// ---- TMA bulk copies (1-D, no tensor map): one elected thread moves a contiguous block between global and shared
// memory; completion of a load is signalled on an mbarrier (complete_tx), of a store through the bulk async-group
__device__ __forceinline__ void mbar_init(unsigned long long * bar, unsigned int count)
{
  const unsigned int a = (unsigned int)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long * bar, unsigned int bytes)
{
  const unsigned int a = (unsigned int)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long * bar, unsigned int parity)
{
  const unsigned int a = (unsigned int)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" :: "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void * smem_dst, const void * gsrc, unsigned int bytes, unsigned long long * bar)
{
  const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
  const unsigned int b = (unsigned int)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void bulk_store(void * gdst, const void * smem_src, unsigned int bytes)
{
  const unsigned int s = (unsigned int)__cvta_generic_to_shared(smem_src);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

}  // namespace bppgpu
