// tree_s20t.cuh -- 20-state tree kernel, sites on the M axis of the FP64 tensor instruction.
//
// tree_s20c.cuh (round 1) computed X = P . clv with the states on M: the product leaves the tensor unit in the
// accumulator layout (one state row per lane quad), but the next mat-vec wants it as a B fragment (one SITE per lane
// quad), so every node went registers -> shared-memory tile -> B fragments, and the same tile fed the global stores.
// ncu had that kernel at 75 % LSU-pipe utilisation and 0.40 of the HBM roofline: ~8 bytes of shared-memory traffic
// per byte of CLV written.  Here the product is transposed, X^T = clv^T . P^T:
//   A (8 x 4)  = 8 sites x 4 child states       -- the node's own values, straight from registers
//   B (4 x 8)  = P^T tile                        -- the same for every site: 15 registers per edge, loaded once per
//                                                   warp and op from a pre-permuted "fragment image" in shared memory
//   D (8 x 8)  = 8 sites x 8 parent states       -- lane (r, q) of the warp gets site r, two states
// A lane quad owns ONE site in both the A and the D layout, and the sum over k may run in any order, so with the
// state permutation below a D fragment IS the A fragment of the next product -- no shuffle, no shared memory:
//   lane (r = lane / 4, q = lane % 4) holds of site r:  w[0..3] = states 4q .. 4q+3,  w[4..5] = states 16+2q, 17+2q (q < 2)
//   n-tile 0 -> states 4(n/2) + n%2, n-tile 1 -> states 4(n/2) + 2 + n%2, n-tile 2 -> states 16 + n (n < 4; zero columns above)
//   k-step ks < 4 -> state 4q + ks = w[ks];  k-step 4 -> states 16, 18, 17, 19 for q = 0..3 (w[4], or the quad
//   neighbour's w[5] by one shuffle).  15 DMMA.8x8x4 per 8 sites and edge, as before.
// The same layout is what the HBM side wants: a lane stores its states 4q..4q+3 with one 256-bit store (a quad
// covers 128 contiguous bytes of the site's 160-byte CLV row) plus one 128-bit store for states 16..19, and an
// HBM-resident child is loaded the same way -- no tile, no transposition.
// A packed tip child is still "column `state` of the edge's P-matrix": the tip edges are staged TRANSPOSED
// (row = child state incl. the <= 4 ambiguity columns, 20 parent states contiguous), a lane reads its 48 bytes of
// the row with three 128-bit loads; quads of odd sites issue the first two in the opposite order, which keeps every
// quarter-warp on eight distinct 16-byte bank groups whatever the tip states are (the round-1 kernel: 1.9x replays).
//
// Staging.  A CTA works on one (locus, category) "group" at a time and keeps that category's matrices in shared
// memory: tip edges as the transposed image, inner edges as the fragment image, 3840 bytes each.  The images are
// built once per step by image20_kernel into the front of the locus' block (building them inside this kernel with
// 8-byte cp.async cost 15 % of its time plus a CTA barrier per group), so a group's stage is two contiguous pieces:
// the category's images and the "meta" block (header, matrix list, op records).  One elected thread fetches them
// with two TMA bulk copies (cp.async.bulk, mbarrier complete_tx) into one of two stage buffers; the warp that
// finishes a group last (shared-memory counter) issues the fetch of the group after next into the buffer it just
// freed.  There is no CTA-wide barrier in the kernel: a warp may run one group ahead of the slowest one.
//
// Per-site scaling (core_partials.c:739-754: a site is rescaled when ALL its 20 x R entries are below 2^-256) needs
// the categories of a site to talk: the scaled instantiation is launched as thread-block clusters of R CTAs, CTA
// rank = category, all working on the same (locus, site block); warp w of every CTA publishes the under-threshold
// bits of its 16 sites in its own shared memory and reads the R-1 others through distributed shared memory
// (st.release / ld.acquire at cluster scope, a sequence number in the word, two parity slots) -- a warp-to-warp
// handshake, no cluster-wide barrier in the op loop.
//
// Reference semantics: core_partials.c:585-756, core_likelihood.c:24-212; parity bar = lnL within 1e-10 relative
// (the reference's own AVX / AVX2 20-state kernels differ from each other in rounding).
#pragma once
#include "tree_s20.cuh"

namespace bppgpu {

#ifndef BPPGPU_S20T_NG
#define BPPGPU_S20T_NG 2
#endif
#ifndef BPPGPU_S20T_NT
#define BPPGPU_S20T_NT 512
#endif
constexpr int S20T_NG = BPPGPU_S20T_NG;              // groups of 8 sites per warp
constexpr int S20T_NT = BPPGPU_S20T_NT;              // threads per CTA
constexpr int S20T_NW = S20T_NT / 32;                // warps
constexpr int S20T_WS = 8 * S20T_NG;                 // sites per warp
constexpr int S20T_SITES = S20T_NW * S20T_WS;        // sites per tile (256)
constexpr int S20T_IMG = 480;                        // doubles per staged matrix image (either form)

// block of a locus: [RL x cap images][Hdr20][list][op records]...; tile_blk points at the header
__host__ __device__ inline size_t s20t_img_bytes(unsigned cap) { return (size_t)cap * S20T_IMG * 8; }        // per category
__host__ __device__ inline size_t s20t_meta_bytes(unsigned max_tips) { return S20_RECS_OFF + (size_t)2 * max_tips * sizeof(OpRec20); }
// scaling handshake ring: two tiles' worth of op words per warp (a tile publishes at most 2 * max_tips), power of two
__host__ __device__ inline unsigned int s20t_hand_slots(unsigned max_tips) { unsigned int h = 1; while (h < 4 * max_tips) h <<= 1; return h; }
__host__ __device__ inline size_t s20t_cols_bytes(unsigned max_tips) { return (size_t)S20T_NW * max_tips * S20T_WS; }   // one of four buffers

// shared memory: [2 x (cap images + meta)][parked X][tip columns x 4][handshake words][mbarriers, counters]
__host__ inline size_t s20t_smem_bytes(unsigned cap, unsigned max_tips, int slots, bool scaled)
{
  size_t b = 2 * (s20t_img_bytes(cap) + s20t_meta_bytes(max_tips));
  b += (size_t)slots * S20T_NW * S20T_NG * 6 * 32 * 8;                            // parked X
  if (scaled) b += (size_t)slots * S20T_NW * S20T_NG * 32 * 4;                    // ... and their scaler counts
  b += 4 * s20t_cols_bytes(max_tips);                                             // tip column ids
  b += (scaled ? (size_t)s20t_hand_slots(max_tips) : 2) * S20T_NW * 8 + 2 * 8 + 2 * 8;   // handshake words, mbarriers, counters
  b += (size_t)S20T_NW * 4 * 32;                                                  // group records
  return b + 16;
}

__device__ __forceinline__ void cp_async16_nc(void * smem_dst, const void * gsrc)
{
  const unsigned int s = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gsrc));
}
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ double2 lds128(unsigned int saddr)
{
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(saddr));
  return v;
}
__device__ __forceinline__ unsigned int cluster_ctarank()
{
  unsigned int r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ----------------------------------------------------------------------------- matrix images
// One warp per (locus, category, list entry): the matrix is read once (coalesced), permuted through shared memory
// and written as 3840 contiguous bytes at  block + (cat * cap + entry) * 3840:
//   tip edge   (ent.y != 0): img[col * 20 + i] = P[i][col], col < 20;  img[(20 + x) * 20 + i] = sum over the states j
//                            of ambiguity mask x of P[i][j], j ascending (what plan_kernel_blocks20 puts into `ext`)
//   inner edge (ent.y == 0): img[(nt * 5 + ks) * 32 + lane] = P[row(nt, lane / 4)][col(ks, lane % 4)], zero for nt = 2, r >= 4
__global__ void __launch_bounds__(256)
image20_kernel(const LocusDev * __restrict__ loci, const unsigned int * __restrict__ batch_locus,
               unsigned char * __restrict__ blocks, const unsigned long long * __restrict__ blk_off,
               unsigned long long hdr_shift, unsigned int RL, unsigned int cap)
{
  __shared__ double s_p[8][S20 * S20];
  const unsigned int bl = blockIdx.x, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const LocusDev & L = loci[batch_locus[bl]];
  unsigned char * blk0 = blocks + blk_off[bl];
  const Hdr20 * H = reinterpret_cast<const Hdr20 *>(blk0 + hdr_shift);
  const unsigned int n_stage = min(H->n_stage, cap);
  const uint2 * list = reinterpret_cast<const uint2 *>(blk0 + hdr_shift + H->stage_off);
  unsigned int cmask[S20_EXT];
#pragma unroll
  for (int x = 0; x < S20_EXT; ++x) cmask[x] = (unsigned)x < L.n_ext_cols ? L.colmask[x] : 0u;
  double * sp = s_p[warp];
  for (unsigned int w = blockIdx.y * nw + warp; w < n_stage * RL; w += gridDim.y * nw)
  {
    const unsigned int m = w / RL, cat = w % RL;
    const uint2 ent = list[m];
    const double * P = L.pmat + ((size_t)ent.x * RL + cat) * (S20 * S20);
    double * img = reinterpret_cast<double *>(blk0) + ((size_t)cat * cap + m) * S20T_IMG;
    __syncwarp();
    for (unsigned int e = 2 * lane; e < S20 * S20; e += 64)
      *reinterpret_cast<double2 *>(sp + e) = *reinterpret_cast<const double2 *>(P + e);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < S20T_IMG / 32; ++k)
    {
      const unsigned int o = k * 32 + lane;
      double v = 0.0;
      if (ent.y)
      {
        const unsigned int col = o / S20, i = o % S20;
        if (col < S20) v = sp[i * S20 + col];
        else
        {
          unsigned int mask = cmask[col - S20];                // no ambiguity codes in the locus: all four are zero
          while (mask) { const int j = __ffs(mask) - 1; mask &= mask - 1; v += sp[i * S20 + j]; }
        }
      }
      else
      {
        const unsigned int nt = k / 5, ks = k % 5, r = lane >> 2, q = lane & 3u;
        const unsigned int row = nt == 0 ? 4 * (r >> 1) + (r & 1u) : (nt == 1 ? 4 * (r >> 1) + 2 + (r & 1u) : 16 + r);
        const unsigned int col = ks < 4 ? 4 * q + ks : 16 + 2 * (q & 1u) + (q >> 1);
        if (row < S20) v = sp[row * S20 + col];
      }
      img[o] = v;
    }
  }
}

// tile word (tile_cell0): bits 0..11 site block, 12..23 site blocks of the locus, 24..31 category
struct Group20 { unsigned int bl, j, t0, pad; unsigned long long blk, pad1; };     // t0 = its first tile in this CTA's range

template <int RL, bool SCALED>
__global__ void __launch_bounds__(S20T_NT, 1)
tree_kernel_s20t(const TreeParams prm, unsigned int * __restrict__ rootsc)
{
  constexpr int NG = S20T_NG, NW = S20T_NW, WS = S20T_WS;
  extern __shared__ __align__(16) unsigned char smem20t[];
  const unsigned int tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const unsigned int r = lane >> 2, q = lane & 3u;
  const bool odd = (r & 1u) != 0;
  const unsigned int cap = prm.lut_cap, maxT = prm.max_tips, opcap = 2 * maxT;
  const size_t img_bytes = s20t_img_bytes(cap), meta_bytes = s20t_meta_bytes(maxT), cols_bytes = s20t_cols_bytes(maxT);
  const size_t buf_bytes = img_bytes + meta_bytes;

  double * s_stack = reinterpret_cast<double *>(smem20t + 2 * buf_bytes);
  unsigned int * s_sstack = reinterpret_cast<unsigned int *>(s_stack + (size_t)prm.n_slots * NW * NG * 6 * 32);
  unsigned char * s_cols_all = reinterpret_cast<unsigned char *>(s_sstack + (SCALED ? (size_t)prm.n_slots * NW * NG * 32 : 0));
  unsigned long long * s_hand = reinterpret_cast<unsigned long long *>(s_cols_all + 4 * cols_bytes);   // [parity][warp]
  const unsigned int HS = SCALED ? s20t_hand_slots(maxT) : 2u;
  unsigned long long * s_full = s_hand + HS * NW;                                                        // [2] stage buffer filled
  unsigned int * s_done = reinterpret_cast<unsigned int *>(s_full + 2);                                  // [2] warps done with the buffer
  unsigned char * s_ring = reinterpret_cast<unsigned char *>(s_done + 4);                                // [warp][4] group records

  // work distribution: unscaled, a tile is (locus, category, site block) and CTAs split the tile list; scaled, a
  // tile is (locus, site block), the CLUSTERS split the list and the CTA's rank in its cluster is the category
  unsigned int my_cat = 0, part = blockIdx.x, parts = gridDim.x;
  if (SCALED && RL > 1) { my_cat = cluster_ctarank(); part = blockIdx.x / RL; parts = gridDim.x / RL; }
  const unsigned int t_begin = (unsigned int)(((unsigned long long)prm.n_tiles * part) / parts);
  const unsigned int t_end = (unsigned int)(((unsigned long long)prm.n_tiles * (part + 1)) / parts);
  unsigned int hand_seq = 0;

  // (locus, category) groups of this CTA's tile range.  Their records (locus, tile word, first tile, block offset)
  // run three groups ahead of their use in a per-warp ring in shared memory, filled by lane 0 with cp.async: held in
  // registers they were spilled the moment they were loaded, and the spill store waited for the load.
  Group20 * const ring = reinterpret_cast<Group20 *>(s_ring) + warp * 4;
  auto group_end = [&](const Group20 & g) -> unsigned int
  {
    return g.t0 >= t_end ? t_end : min(t_end, g.t0 + ((g.j >> 12) & 0xFFFu) - (g.j & 0xFFFu));
  };
  auto group_cat = [&](const Group20 & g) -> unsigned int { return (SCALED && RL > 1) ? my_cat : (g.j >> 24); };
  auto request_group = [&](Group20 * dst, unsigned int t)     // lane 0
  {
    dst->t0 = t;
    if (t < t_end)
    {
      cp_async4(&dst->bl, prm.tile_locus + t);
      cp_async4(&dst->j, prm.tile_cell0 + t);
      const unsigned int s = (unsigned int)__cvta_generic_to_shared(&dst->blk);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(s), "l"(prm.tile_blk + 2 * (size_t)t) : "memory");
    }
    else { dst->bl = 0; dst->j = 1u << 12; dst->blk = 0; }
  };
  // one thread: images of the group's category + meta block into stage buffer b
  auto fetch_group = [&](const Group20 & g, unsigned int b)
  {
    unsigned char * dst = smem20t + b * buf_bytes;
    const unsigned char * hdr = prm.blocks + g.blk;
    fence_proxy_async();
    mbar_expect_tx(s_full + b, (unsigned int)buf_bytes);
    bulk_load(dst, hdr - (size_t)RL * img_bytes + (size_t)group_cat(g) * img_bytes, (unsigned int)img_bytes, s_full + b);
    bulk_load(dst + img_bytes, hdr, (unsigned int)meta_bytes, s_full + b);
  };
  // prologue: the first three records (a dependent chain, once per CTA), barriers, the first two fetches
  if (lane == 0)
  {
    for (int i = 0; i < 3; ++i)
    {
      request_group(ring + i, i == 0 ? t_begin : group_end(ring[i - 1]));
      cp_async_commit();
      cp_async_wait<0>();
    }
  }
  if (tid == 0)
  {
    mbar_init(s_full, 1); mbar_init(s_full + 1, 1);
    s_done[0] = s_done[1] = 0;
    mbar_init_fence();
    if (ring[0].t0 < t_end) fetch_group(ring[0], 0);
    if (ring[1].t0 < t_end) fetch_group(ring[1], 1);
  }
  if (SCALED && RL > 1) for (unsigned int i = tid; i < HS * NW; i += S20T_NT) s_hand[i] = 0ull;
  __syncthreads();
  if (SCALED && RL > 1) cluster_sync_all();                   // every CTA of the cluster runs and has cleared its words

  // the warp's tip columns of one tile: T x 16 bytes
  auto fetch_cols = [&](unsigned char * dst, const Hdr20 * H, unsigned int site0)
  {
    const unsigned int pitch = H->cols_pitch, T = H->tips;
    for (unsigned int e = lane; e < T * (WS / 16); e += 32)
    {
      const unsigned int tip = e / (WS / 16), off = (e % (WS / 16)) * 16;
      if (site0 + off < pitch) cp_async16_nc(dst + tip * WS + off, H->tip_cols + (size_t)tip * pitch + site0 + off);
    }
  };
  auto cols_buf = [&](unsigned int i) -> unsigned char * { return s_cols_all + i * cols_bytes + (size_t)warp * maxT * WS; };

  if (ring[0].t0 < t_end)
  {
    mbar_wait(s_full, 0);
    fetch_cols(cols_buf(0), reinterpret_cast<const Hdr20 *>(smem20t + img_bytes), (ring[0].j & 0xFFFu) * S20T_SITES + warp * WS);
    cp_async_commit();
  }

  for (unsigned int gi = 0; ring[gi & 3u].t0 < t_end; ++gi)
  {
    cp_async_wait<0>();                                        // the first tile's columns, the record of group gi + 2
    __syncwarp();
    const Group20 G0 = ring[gi & 3u];
    const unsigned int t_first = G0.t0, t_last = group_end(G0), b = gi & 1u;
    const unsigned int bl = G0.bl;
    const unsigned int cat = group_cat(G0);
    const unsigned char * const blk = prm.blocks + G0.blk;
    const unsigned char * const meta = smem20t + b * buf_bytes + img_bytes;
    const Hdr20 * const H = reinterpret_cast<const Hdr20 *>(meta);
    mbar_wait(s_full + b, (gi >> 1) & 1u);                     // normally satisfied long ago (see the last tile below)
    // the record of the group three ahead; it is committed with the first tile's columns and complete one tile later
    if (lane == 0) request_group(ring + ((gi + 3) & 3u), group_end(ring[(gi + 2) & 3u]));

    const unsigned int sites = H->sites, nops = H->nops;
    const uint4 * recs = nops <= opcap ? reinterpret_cast<const uint4 *>(meta + S20_RECS_OFF)
                                       : reinterpret_cast<const uint4 *>(blk + S20_RECS_OFF);
    const unsigned int stage_sa = (unsigned int)__cvta_generic_to_shared(smem20t + b * buf_bytes);
    const double * const stage = reinterpret_cast<const double *>(smem20t + b * buf_bytes);
    double * const clv = H->clv;
    const unsigned long long stride = H->clv_stride;
    const unsigned int sstr = H->site_stride, cstr = H->cat_stride;
    const bool next_valid = ring[(gi + 1) & 3u].t0 < t_end;

    for (unsigned int t = t_first; t < t_last; ++t)
    {
    const unsigned int rt = t - t_first;
    const unsigned int site0 = ((G0.j & 0xFFFu) + rt) * S20T_SITES + warp * WS;
    // tip columns run one tile ahead: the next tile of this group, or the first tile of the next group (whose
    // header is in the other stage buffer -- waiting for it here instead of at the top of the next group gives
    // the columns a whole tile to arrive)
    const unsigned char * const s_cols = cols_buf(rt == 0 ? b : 2 + (rt & 1u));
    if (t + 1 < t_last) fetch_cols(cols_buf(2 + ((rt + 1) & 1u)), H, site0 + S20T_SITES);
    else if (next_valid)
    {
      mbar_wait(s_full + (b ^ 1u), ((gi + 1) >> 1) & 1u);
      fetch_cols(cols_buf(b ^ 1u), reinterpret_cast<const Hdr20 *>(smem20t + (b ^ 1u) * buf_bytes + img_bytes),
                 (ring[(gi + 1) & 3u].j & 0xFFFu) * S20T_SITES + warp * WS);
    }
    cp_async_commit();
    cp_async_wait<1>();                                        // this tile's columns (committed one tile ago)
    __syncwarp();
    if (site0 >= sites) continue;                              // a warp beyond the end of the locus

    unsigned int sitev[NG];
    bool validv[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g)
    {
      const unsigned int s = site0 + 8 * g + r;
      validv[g] = s < sites;
      sitev[g] = validv[g] ? s : sites - 1;
    }

    // scaling handshake: a warp's bits of one op in its own shared memory, read by the same warp of the other CTAs.
    // The 64-bit word (sequence number, bits) is all that is communicated, so relaxed accesses suffice: a release
    // store at cluster scope is a full memory barrier that waits for every CLV store the thread has in flight, an
    // acquire load invalidates L1 -- together they cost this kernel 1.4 ms.
    auto publish = [&](unsigned int seq, unsigned int m)
    {
      if (lane == 0)
      {
        const unsigned int la = (unsigned int)__cvta_generic_to_shared(s_hand + (seq & (HS - 1u)) * NW + warp);
        const unsigned long long v = ((unsigned long long)seq << 32) | m;
        asm volatile("st.relaxed.cluster.shared::cta.u64 [%0], %1;" :: "r"(la), "l"(v) : "memory");
      }
      __syncwarp();      // tile_check reads the warp's own word too (lane = op x category): ordered within the warp
    };
    auto consume = [&](unsigned int seq) -> unsigned int        // AND of the other categories' bits of op `seq`
    {
      unsigned int theirs = 0xFFFFFFFFu;
      if (lane < RL && lane != cat)
      {
        const unsigned int la = (unsigned int)__cvta_generic_to_shared(s_hand + (seq & (HS - 1u)) * NW + warp);
        unsigned int ra;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(lane));
        unsigned long long v;
        do
        {
          asm volatile("ld.relaxed.cluster.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(ra) : "memory");
        } while ((unsigned int)(v >> 32) != seq);
        theirs = (unsigned int)v;
      }
#pragma unroll
      for (int d = 1; d < RL; d <<= 1) theirs &= __shfl_xor_sync(0xFFFFFFFFu, theirs, d);
      return __shfl_sync(0xFFFFFFFFu, theirs, 0);
    };
    // the ops seq0 + 1 .. seq0 + n of this tile at once: lane = (op, category); true if some site's AND is non-zero
    auto tile_check = [&](unsigned int seq0, unsigned int n) -> bool
    {
      bool ev = false;
      for (unsigned int base = 0; base < n; base += 32 / RL)
      {
        const unsigned int i = base + lane / RL, p = lane % RL;
        unsigned int v = 0u;
        if (i < n)
        {
          const unsigned int seq = seq0 + 1 + i;
          const unsigned int la = (unsigned int)__cvta_generic_to_shared(s_hand + (seq & (HS - 1u)) * NW + warp);
          unsigned int ra;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(p));
          unsigned long long w;
          do
          {
            asm volatile("ld.relaxed.cluster.shared::cluster.u64 %0, [%1];" : "=l"(w) : "r"(ra) : "memory");
          } while ((unsigned int)(w >> 32) != seq);
          v = (unsigned int)w;
        }
#pragma unroll
        for (int d = 1; d < RL; d <<= 1) v &= __shfl_xor_sync(0xFFFFFFFFu, v, d);
        ev = ev || __any_sync(0xFFFFFFFFu, v != 0u);
      }
      return ev;
    };
    // Scaled batches run a tile optimistically first: every op publishes its under-threshold bits but goes on as if
    // no site were rescaled; the other categories' bits are read once, at the end of the tile (by then they have
    // arrived, so nobody waits).  If the AND ever is non-zero (rare: all 20 x R entries of a site below 2^-256) the
    // warp -- and with it the same warp of the R-1 other CTAs, which see the same AND at the same op -- starts the
    // tile again in lock-step mode, where every op waits for its peers before it stores.  Stores are idempotent.
    bool sync_mode = false;
    for (;;)
    {
    bool redo = false;
    unsigned int n_pub = 0;
    double X[NG][6];
    unsigned int xsc[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g)
    {
      xsc[g] = 0;
#pragma unroll
      for (int i = 0; i < 6; ++i) X[g][i] = 0.0;
    }

    // a CLV row of the lane's site: states 4q..4q+3 (256 bits) and 16+2q, 17+2q (128 bits, q < 2)
    auto load_clv = [&](const double * buf, bool coherent, double (&v)[NG][6])
    {
#pragma unroll
      for (int g = 0; g < NG; ++g)
      {
        const double * p = buf + (size_t)sitev[g] * sstr + (size_t)cat * cstr;
        if (coherent) ld256(p + 4 * q, v[g][0], v[g][1], v[g][2], v[g][3]);
        else ld256_nc(p + 4 * q, v[g][0], v[g][1], v[g][2], v[g][3]);
        v[g][4] = v[g][5] = 0.0;
        if (q < 2)
          asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
                       : "=d"(v[g][4]), "=d"(v[g][5]) : "l"(p + 16 + 2 * q) : "memory");
      }
    };
    // out = (in . P^T) for the warp's NG x 8 sites; img = fragment image of the edge
    auto push = [&](const double * img, const double (&in)[NG][6], double (&out)[NG][6])
    {
      double a4[NG];
#pragma unroll
      for (int g = 0; g < NG; ++g)
      {
        const double t5 = __shfl_xor_sync(0xFFFFFFFFu, in[g][5], 2);
        a4[g] = q < 2 ? in[g][4] : t5;
#pragma unroll
        for (int i = 0; i < 6; ++i) out[g][i] = 0.0;
      }
      // k-steps outermost: NG x 3 independent accumulator chains; the three fragments of a k-step are loaded as
      // they are needed (15 live fragment registers would not leave room for the next op's prefetched operand)
#pragma unroll
      for (int ks = 0; ks < 5; ++ks)
      {
        const double b0 = img[ks * 32 + lane], b1 = img[(5 + ks) * 32 + lane], b2 = img[(10 + ks) * 32 + lane];
#ifdef S20T_ABL_NODMMA
#pragma unroll
        for (int g = 0; g < NG; ++g) { out[g][0] += in[g][ks] * b0; out[g][2] += a4[g] * b1; out[g][4] += in[g][5 - ks] * b2; }
#else
#pragma unroll
        for (int g = 0; g < NG; ++g)
        {
          const double a = ks < 4 ? in[g][ks] : a4[g];
          dmma(out[g][0], out[g][1], a, b0);
          dmma(out[g][2], out[g][3], a, b1);
          dmma(out[g][4], out[g][5], a, b2);
        }
#endif
      }
    };
    auto fetch = [&](unsigned int kind, unsigned int p0, unsigned int st, int scidx, double (&v)[NG][6],
                     unsigned int (&sc)[NG])
    {
      if (kind == SRC_TIP_PACKED)
      {
        // byte address of the lane's first piece in row 0 of the transposed image; odd sites read the two halves
        // of their 32 bytes in the opposite order (bank groups), the third piece is states 16 + 2q, 17 + 2q
        const unsigned int a0 = stage_sa + st * (S20T_IMG * 8) + 32 * q;
#pragma unroll
        for (int g = 0; g < NG; ++g)
        {
          const unsigned int row = a0 + (unsigned int)s_cols[p0 * WS + 8 * g + r] * (S20 * 8);
          const double2 u0 = lds128(row + (odd ? 16u : 0u));
          const double2 u1 = lds128(row + (odd ? 0u : 16u));
          v[g][0] = odd ? u1.x : u0.x; v[g][1] = odd ? u1.y : u0.y;
          v[g][2] = odd ? u0.x : u1.x; v[g][3] = odd ? u0.y : u1.y;
          v[g][4] = v[g][5] = 0.0;
          if (q < 2) { const double2 u2 = lds128(row + 128u - 16u * q); v[g][4] = u2.x; v[g][5] = u2.y; }
          sc[g] = 0;
        }
      }
      else if (kind == SRC_SLOT)
      {
        const double2 * sk = reinterpret_cast<const double2 *>(s_stack) + (size_t)(p0 * NW + warp) * (NG * 3 * 32) + lane;
#pragma unroll
        for (int g = 0; g < NG; ++g)
        {
#pragma unroll
          for (int h = 0; h < 3; ++h) { const double2 u = sk[(g * 3 + h) * 32]; v[g][2 * h] = u.x; v[g][2 * h + 1] = u.y; }
          sc[g] = SCALED ? s_sstack[(size_t)(p0 * NW + warp) * (NG * 32) + g * 32 + lane] : 0u;
        }
      }
      else
      {
        // HBM-resident child CLV (or dense tip): load it in the lane's layout and push it through its edge
        double c[NG][6];
        load_clv((kind == SRC_TIP_DENSE ? H->tip_dense : clv) + (size_t)p0 * stride, kind == SRC_HBM, c);
        push(stage + (size_t)st * S20T_IMG, c, v);
#pragma unroll
        for (int g = 0; g < NG; ++g)
          sc[g] = (SCALED && kind == SRC_HBM && scidx >= 0) ? H->scale[(size_t)scidx * sites + sitev[g]] : 0u;
      }
    };

    // operand A of the next op (a packed tip or a parked value, never the previous result) is fetched before the
    // current op's push, so its shared-memory latency hides behind the DMMAs
    double A[NG][6];
    unsigned int asc[NG];
    bool a_ready = false;
    for (unsigned int k = 0; k < nops; ++k)
    {
      const uint4 w0 = recs[4 * k], w1 = recs[4 * k + 1], w2 = recs[4 * k + 2], w3 = recs[4 * k + 3];
      const unsigned int ctl = w0.x;
      const unsigned int akind = (ctl >> OP_AKIND_SHIFT) & 15u, bkind = (ctl >> OP_BKIND_SHIFT) & 15u;
      double O[NG][6];
      unsigned int osc[NG];

      if (ctl & OP_EVAL)
      {
        // root CLV that this list did not produce: read it as is
        if (akind == SRC_TIP_PACKED)
        {
#pragma unroll
          for (int g = 0; g < NG; ++g)
          {
            const unsigned int col = s_cols[w0.z * WS + 8 * g + r];
            const unsigned int mask = (col < S20) ? (1u << col) : prm.loci[prm.batch_locus[bl]].colmask[col - S20];
#pragma unroll
            for (int i = 0; i < 4; ++i) O[g][i] = (double)((mask >> (4 * q + i)) & 1u);
            O[g][4] = q < 2 ? (double)((mask >> (16 + 2 * q)) & 1u) : 0.0;
            O[g][5] = q < 2 ? (double)((mask >> (17 + 2 * q)) & 1u) : 0.0;
            osc[g] = 0;
          }
        }
        else
        {
          load_clv((akind == SRC_TIP_DENSE ? H->tip_dense : clv) + (size_t)w0.z * stride, akind == SRC_HBM, O);
#pragma unroll
          for (int g = 0; g < NG; ++g)
            osc[g] = (SCALED && akind == SRC_HBM && (int)w1.x >= 0) ? H->scale[(size_t)(int)w1.x * sites + sitev[g]] : 0u;
        }
      }
      else
      {
        if (!a_ready) fetch(akind, w0.z, w3.y, (int)w1.x, A, asc);
        a_ready = false;
        if (!(ctl & OP_BPREV)) fetch(bkind, w1.z, w3.z, (int)w2.x, X, xsc);         // B into the (dead) X registers
#pragma unroll
        for (int g = 0; g < NG; ++g)
        {
          osc[g] = asc[g] + xsc[g];
#pragma unroll
          for (int i = 0; i < 6; ++i) O[g][i] = A[g][i] * X[g][i];
        }
        if (SCALED && (ctl & OP_SCALE))
        {
          // all 20 x R entries of a site strictly below 2^-256: entries are non-negative and 2^-256 has a zero
          // mantissa, so that is max(high words) < 0x2FF00000 (integer pipe; padding entries of lanes q >= 2 are zeros)
          unsigned int mine = 1u;
#pragma unroll
          for (int g = 0; g < NG; ++g)
          {
            int hm = __double2hiint(O[g][0]);
#pragma unroll
            for (int i = 1; i < 6; ++i) hm = max(hm, __double2hiint(O[g][i]));
            unsigned int b = hm < 0x2FF00000 ? 1u : 0u;
            b &= __shfl_xor_sync(0xFFFFFFFFu, b, 1);
            b &= __shfl_xor_sync(0xFFFFFFFFu, b, 2);
            if (q == (unsigned)g) mine = b;
          }
          unsigned int m = __ballot_sync(0xFFFFFFFFu, mine != 0u) & (NG >= 4 ? 0xFFFFFFFFu : (NG == 2 ? 0x33333333u : 0x11111111u));   // bit 4r + g = site r of group g
          if (RL > 1)
          {
            ++hand_seq;
            publish(hand_seq, m);
            if (sync_mode) m &= consume(hand_seq);
            else { ++n_pub; m = 0; }
          }
          if (m)                                                         // warp-uniform, rare
          {
#pragma unroll
            for (int g = 0; g < NG; ++g)
              if ((m >> (4 * r + g)) & 1u)
              {
#pragma unroll
                for (int i = 0; i < 6; ++i) O[g][i] *= BPPGPU_SCALE_FACTOR;
                osc[g] += 1;
              }
          }
          if (cat == 0 && q == 0)
          {
#pragma unroll
            for (int g = 0; g < NG; ++g)
              if (validv[g]) H->scale[(size_t)(int)w2.w * sites + sitev[g]] = osc[g];
          }
        }
        else
        {
#pragma unroll
          for (int g = 0; g < NG; ++g) osc[g] = 0;
        }
        // ---- the CLV goes to HBM exactly once, straight from the registers
#ifndef S20T_ABL_NOSTORE
#pragma unroll
        for (int g = 0; g < NG; ++g)
          if (validv[g])
          {
            double * p = clv + (size_t)w0.y * S20 + (size_t)sitev[g] * sstr + (size_t)cat * cstr;
            st256(p + 4 * q, O[g][0], O[g][1], O[g][2], O[g][3]);
            if (q < 2) st128(p + 16 + 2 * q, O[g][4], O[g][5]);
          }
#endif
        // ---- operand A of the next op
        if (k + 1 < nops)
        {
          const uint4 n0 = recs[4 * k + 4], n3 = recs[4 * k + 7];
          const unsigned int nkind = (n0.x >> OP_AKIND_SHIFT) & 15u;
          if (!(n0.x & OP_EVAL) && (nkind == SRC_TIP_PACKED || nkind == SRC_SLOT))
          {
            fetch(nkind, n0.z, n3.y, -1, A, asc);
            a_ready = true;
          }
        }
        // ---- push through the edge above
        if (ctl & OP_PUSH)
        {
          push(stage + (size_t)w3.w * S20T_IMG, O, X);
#pragma unroll
          for (int g = 0; g < NG; ++g) xsc[g] = osc[g];
          if (ctl & OP_PARKA)
          {
            double2 * sk = reinterpret_cast<double2 *>(s_stack) + (size_t)(w3.x * NW + warp) * (NG * 3 * 32) + lane;
#pragma unroll
            for (int g = 0; g < NG; ++g)
            {
#pragma unroll
              for (int h = 0; h < 3; ++h) sk[(g * 3 + h) * 32] = make_double2(X[g][2 * h], X[g][2 * h + 1]);
              if (SCALED) s_sstack[(size_t)(w3.x * NW + warp) * (NG * 32) + g * 32 + lane] = xsc[g];
            }
          }
        }
      }

      if (ctl & OP_ROOT)
      {
        // pi . clv of this category, reduced over the quad; root20_kernel does the rest
        double * out = prm.rootdot + (size_t)prm.site_off[bl] * RL + (size_t)cat * sites;
#pragma unroll
        for (int g = 0; g < NG; ++g)
        {
          double s = 0.0;
#pragma unroll
          for (int i = 0; i < 4; ++i) s += H->freqs[4 * q + i] * O[g][i];
          if (q < 2) s += H->freqs[16 + 2 * q] * O[g][4] + H->freqs[17 + 2 * q] * O[g][5];
          s += __shfl_xor_sync(0xFFFFFFFFu, s, 1);
          s += __shfl_xor_sync(0xFFFFFFFFu, s, 2);
          if (q == 0 && validv[g])
          {
            out[sitev[g]] = s;
            if (SCALED && cat == 0)
            {
              unsigned int rsc = osc[g];
              if ((ctl & OP_EVAL) && (int)w2.w < 0) rsc = 0;
              rootsc[prm.site_off[bl] + sitev[g]] = rsc;
            }
          }
        }
      }
    }
    if (SCALED && RL > 1 && n_pub && tile_check(hand_seq - n_pub, n_pub)) redo = true;
    if (!redo) break;
    sync_mode = true;                                          // a site has to be rescaled: again, in lock step
    }
    __syncwarp();                                              // the tile's reads of its column ids are over
    }
    // ---- done with the group: the last warp to get here refills the buffer with the group after next
    __syncwarp();
    if (lane == 0)
    {
      __threadfence_block();
      if (atomicAdd(s_done + b, 1u) == NW - 1)
      {
        s_done[b] = 0;
        __threadfence_block();
        const Group20 G2 = ring[(gi + 2) & 3u];
        if (G2.t0 < t_end) fetch_group(G2, b);
      }
    }
    __syncwarp();
  }
  cp_async_wait<0>();
  if (SCALED && RL > 1) cluster_sync_all();                    // nobody leaves while a peer may still poll its words
}

// ----------------------------------------------------------------------------- root of the category-major kernel
// One CTA per locus: term = sum_cat rw_cat * rootdot[cat][site] (core_likelihood.c:179-196), log, scaler
// correction, pattern weight, fixed-order block reduction.  The locus' value goes into the first of its tile
// partials (the others are zeroed) so that finish_kernel sums the batch exactly as for the other kernels.
__global__ void __launch_bounds__(128)
root20_kernel(const LocusDev * __restrict__ loci, const unsigned int * __restrict__ batch_locus,
              const double * __restrict__ rootdot, const unsigned long long * __restrict__ site_off,
              const unsigned int * __restrict__ rootsc, double log_threshold,
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
    else
    {
      v = log(term);
      const unsigned int rsc = rootsc ? rootsc[site_off[bl] + s] : 0u;
      if (rsc) v += (double)rsc * log_threshold;
      v *= (double)L.weights[s];
    }
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
