// engine.cu -- host side of the C-ABI (include/bpp_b200.h): device-resident locus store, batches,
// staging, launches.  No CPU compute fallback: without a CUDA device every entry point fails
// through the fatal handler.
//
// Reference interfaces mirrored here (file:line relative to /root/reference/src):
//   locus_create locus.c:622-870, locus_destroy :872, pll_set_tip_states :561, pll_set_tip_clv :596,
//   pll_set_frequencies :889, pll_set_subst_params :877, locus_update_matrices :2417,
//   locus_update_partials :2530, locus_root_loglikelihood :2573, pll_update_eigen core_pmatrix.c:239.
#include "../../include/bpp_b200.h"
#include "common.cuh"
#include "plan.cuh"
#include "tree_s4.cuh"
#include "tree_s20.cuh"
#include "tree_s20t.cuh"
#include "tree_generic.cuh"
#include "reduce.cuh"
#include "model.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <vector>

using namespace bppgpu;

#define BPPGPU_MAX_WAVES 8

// ------------------------------------------------------------------------------------ errors
static thread_local std::string g_last_error;
static void (*g_fatal_handler)(const char *) = nullptr;

static void fatal(const char * fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  if (g_fatal_handler) { g_fatal_handler(buf); return; }
  fprintf(stderr, "bppgpu: %s\n", buf);     // util.c:30 fatal(): message + exit(1)
  exit(1);
}

#define CUDA_CHECK(call)                                                                         \
  do {                                                                                           \
    cudaError_t err__ = (call);                                                                  \
    if (err__ != cudaSuccess)                                                                    \
      fatal("CUDA error %s at %s:%d: %s", cudaGetErrorName(err__), __FILE__, __LINE__,           \
            cudaGetErrorString(err__));                                                          \
  } while (0)

// ------------------------------------------------------------------------------------ arena
// Slab allocator over cudaMalloc: loci are created once and live for the whole run, so a bump
// pointer with exact-size reuse of freed blocks is enough (and avoids 10^4-10^5 cudaMalloc calls).
struct Arena
{
  static constexpr size_t kAlign = 256;
  size_t slab_bytes = (size_t)1 << 30;
  std::vector<char *> slabs;
  char * cur = nullptr;
  size_t cur_left = 0;
  size_t total = 0;
  std::multimap<size_t, char *> free_blocks;

  void * alloc(size_t bytes)
  {
    bytes = (bytes + kAlign - 1) / kAlign * kAlign;
    if (bytes == 0) bytes = kAlign;
    auto it = free_blocks.find(bytes);
    if (it != free_blocks.end()) { char * p = it->second; free_blocks.erase(it); return p; }
    if (bytes > cur_left)
    {
      size_t sz = std::max(slab_bytes, bytes);
      char * p = nullptr;
      cudaError_t err = cudaMalloc(&p, sz);
      if (err != cudaSuccess)
      {
        fatal("Unable to allocate enough memory on the device (%zu bytes): %s", sz, cudaGetErrorString(err));
        return nullptr;
      }
      slabs.push_back(p);
      total += sz;
      cur = p; cur_left = sz;
    }
    char * r = cur;
    cur += bytes; cur_left -= bytes;
    return r;
  }
  void release(void * p, size_t bytes)
  {
    if (!p) return;
    bytes = (bytes + kAlign - 1) / kAlign * kAlign;
    if (bytes == 0) bytes = kAlign;
    free_blocks.emplace(bytes, (char *)p);
  }
  void destroy()
  {
    for (char * s : slabs) cudaFree(s);
    slabs.clear(); free_blocks.clear(); cur = nullptr; cur_left = 0; total = 0;
  }
};

// ------------------------------------------------------------------------------------ objects
struct bppgpu_engine
{
  int device = 0;
  unsigned int math = BPPGPU_MATH_EXACT;
  cudaStream_t stream = nullptr;
  Arena arena;
  std::mutex mu;
  // device table of LocusDev, indexed by locus id
  LocusDev * d_loci = nullptr;
  size_t loci_cap = 0;
  std::vector<bppgpu_locus *> loci;      // by id (nullptr = free)
  std::vector<unsigned int> free_ids;
  std::atomic<unsigned long long> launches{0};   // kernels launched (host threads on disjoint batches count concurrently)
  std::atomic<unsigned long long> dirty_epoch{1};   // bumped whenever a locus' host mirrors change
  int sm_count = 148;
  size_t smem_optin = 0, smem_per_sm = 0;
  // profiling
  bool profiling = false;
  double prof_ms[BPPGPU_KERNEL_COUNT] = {0, 0, 0, 0};
  unsigned long long prof_n[BPPGPU_KERNEL_COUNT] = {0, 0, 0, 0};
  struct PendingEvent { cudaEvent_t a, b; int kind; };
  std::vector<PendingEvent> pending;
  std::vector<cudaEvent_t> event_pool;
  std::mutex prof_mu;                    // host threads on disjoint batches profile concurrently
  double log_threshold = 0;
  // kernels whose dynamic shared-memory limit has been raised on this device.  The limit is a property of the
  // FUNCTION, shared by every batch and host thread: it is raised once to the device maximum, never per launch
  // (two batches with different sizes would otherwise lower it under each other's feet).
  std::mutex attr_mu;
  std::set<const void *> attr_done;
};

template <typename K>
static void ensure_max_smem(bppgpu_engine * e, K kernel)
{
  std::lock_guard<std::mutex> lock(e->attr_mu);
  const void * key = reinterpret_cast<const void *>(kernel);
  if (e->attr_done.count(key)) return;
  // the opt-in limit covers static + dynamic shared memory
  cudaFuncAttributes fa;
  size_t stat = 0;
  if (cudaFuncGetAttributes(&fa, kernel) == cudaSuccess) stat = fa.sharedSizeBytes; else cudaGetLastError();
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(e->smem_optin - stat));
  if (err != cudaSuccess) { fprintf(stderr, "bppgpu: cudaFuncSetAttribute: %s\n", cudaGetErrorString(err)); cudaGetLastError(); }
  e->attr_done.insert(key);
}

struct bppgpu_locus
{
  bppgpu_engine * e = nullptr;
  unsigned int id = 0;
  unsigned int dtype = 0, model = 0, tips = 0, clv_buffers = 0, states = 0, sites = 0;
  unsigned int rate_matrices = 0, prob_matrices = 0, rate_cats = 0, scale_buffers = 0, attributes = 0;
  // rate_cats is what the DEVICE holds: 1, 2, 4 or 8 categories for 4- and 20-state loci, so that every category count
  // up to 8 runs the fast kernels.  user_cats is what the caller created the locus with; the categories beyond it are
  // copies of category 0 with weight 0 -- they add +0.0 to every site likelihood and, being copies, never change
  // whether ALL categories of a site are below the scaling threshold -- and the accessors strip them again.
  unsigned int user_cats = 0;
  LocusDev dev;                        // host copy of the device descriptor
  // sizes of the arena blocks (for release)
  size_t b_clv = 0, b_tipdense = 0, b_codes = 0, b_flags = 0, b_pmat = 0, b_scale = 0, b_weights = 0, b_model = 0;
  size_t b_dip_off = 0, b_dip_map = 0, b_dip_w = 0;
  // host mirrors of the small inputs
  std::vector<unsigned int> h_codes;    // 4 states: [pattern][tip/8] nibbles; else [tip][pattern] masks
  std::vector<unsigned char> h_cols;    // > 4 states: [tip][pattern] column ids
  unsigned int h_colmask[4] = {0, 0, 0, 0};
  unsigned int n_ext_cols = 0;
  bool col_overflow = false;            // more than 4 distinct ambiguity masks: generic kernel only
  size_t b_cols = 0, b_colmask = 0;
  std::vector<unsigned char> h_tip_dense_flag;
  std::vector<double> h_freqs, h_subst, h_rates, h_rate_weights, h_evecs, h_ievecs, h_evals;
  bool eigen_valid = false;            // locus->eigen_decomp_valid[0]
  bool eigen_on_device = false;        // ... computed by model_update_kernel: the host mirrors h_evecs.. are stale
  bool codes_dirty = false, model_dirty = false, flags_dirty = false;
  bppgpu_batch * self_batch = nullptr; // batch of one for the synchronous per-locus API
};

struct bppgpu_batch
{
  bppgpu_engine * e = nullptr;
  unsigned int n = 0;
  std::vector<bppgpu_locus *> loci;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int kernel_kind = 0;                 // 0 = 4-state kernel, 1 = generic fallback, 2 = 20-state DMMA kernel
  unsigned int RL = 1;                 // lanes per site of the 4-state kernel
  unsigned int cpt = 1;                // cells per thread of the 4-state kernel
  unsigned int tile_threads = 256;
  unsigned int n_tiles = 0;
  // static device tables
  unsigned int * d_batch_locus = nullptr, * d_tile_locus = nullptr, * d_tile_cell0 = nullptr, * d_tile_first = nullptr;
  unsigned long long * d_scratch_off = nullptr;
  unsigned char * d_scratch = nullptr;
  size_t scratch_bytes = 0;
  // per-step device inputs (one blob) and outputs
  char * d_in = nullptr; size_t d_in_cap = 0;
  char * h_in = nullptr; size_t h_in_cap = 0;       // pinned
  PlanOp * d_plan = nullptr; size_t plan_cap = 0;   // generic kernel: flat plan
  unsigned char * d_blocks = nullptr; size_t blocks_cap = 0;   // 4-state kernel: staged per-locus blocks (current parity)
  // plan cache: the planned blocks of the staged op lists are kept for BOTH index parities (the lists as staged
  // and after bppgpu_batch_flip_indices); a run whose blocks are still valid only refreshes the P-matrices
  unsigned char * d_blocks_par[2] = {nullptr, nullptr};
  int parity = 0;
  bool plan_valid[2] = {false, false};
  unsigned int plan_key[2] = {0, 0};     // want_root / slots / tip-slot capacity the cached plan was made for
  cudaEvent_t ev_inputs = nullptr;       // recorded behind the last H2D copy of the step's host arrays
  // class of a cached 4-state plan (plan_class_kernel, read back asynchronously): PLAN_CLASS_* bits once known.
  // Runs on a cached plan whose loci are all lean (or all scaled one-chunk) launch the specialised kernel.
  unsigned int * d_class = nullptr, * h_class = nullptr;   // 2 words each (per parity); h_class pinned
  cudaEvent_t ev_class[2] = {nullptr, nullptr};
  bool class_pending[2] = {false, false};
  unsigned int plan_class[2] = {0, 0};
  TileDesc * d_tiles = nullptr;
  unsigned long long * d_tile_blk = nullptr;        // per tile: {block offset, 0}, written by the planner
  unsigned int * d_plan_count = nullptr;
  int grid = 0; size_t tree_smem = 0; int slots = 0;
  double * d_tile_partial = nullptr, * d_lnl = nullptr, * d_lnl_sum = nullptr, * d_block_sums = nullptr;
  unsigned int * d_counter = nullptr;
  double * h_out = nullptr;                         // pinned: n lnl + 1 sum
  double * d_persite = nullptr; size_t persite_cap = 0;
  // layout of the staged blob
  size_t o_mat_off = 0, o_mat_idx = 0, o_mat_bl = 0, o_op_off = 0, o_ops = 0, o_root_clv = 0, o_root_sc = 0, o_blk_off = 0;
  unsigned int total_mats = 0, total_ops = 0;
  unsigned int lut_cap_rt = 0;         // tip-slot capacity of the 4-state launches: the batch's largest tree
  bool s20_cat = false;                // 20 states: category-major tiles (tree_kernel_s20t)
  bool s20_scaled = false;             // ... a locus of the batch scales per site: clusters of RL CTAs, tile = (locus, site block)
  double * d_rootdot = nullptr;        // ... its per (category, site) root dot products
  unsigned int * d_rootsc = nullptr;   // ... and the root's scaler count per site (scaled batches)
  unsigned long long * d_site_off = nullptr;
  bool staged_mats = false, staged_ops = false, staged_roots = false;
  bool all_in_blob = false; size_t blob_bytes = 0;   // the staged step lives entirely in h_in (device layout): one H2D copy
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  unsigned long long synced_epoch = 0; // engine dirty_epoch at the last batch_sync_loci
  unsigned long long diploid_epoch = 0; bool any_diploid = false;   // a locus of the batch carries a diploid mapping
  // batched model sync: pinned blob + device copy, record offsets, ids, eigen modes, Jacobi scratch
  char * h_model = nullptr, * d_model = nullptr; size_t model_cap = 0;
  double * d_eig_scratch = nullptr; size_t eig_scratch_cap = 0;
  bool synced_eigen = false;           // ... which also brought the eigen-decompositions up to date
  // pipelined step (4-state kernel, big batches): the step's host arrays are uploaded in waves of loci on
  // copy_stream while the planner / tree kernels of earlier waves already run (alternating between
  // `stream` and alt_stream so that a wave fills the tail of the previous one)
  cudaStream_t copy_stream = nullptr, alt_stream = nullptr;
  unsigned int n_waves = 1;            // waves of the CURRENT staged inputs
  unsigned int wave_first[BPPGPU_MAX_WAVES + 1] = {0};
  cudaEvent_t ev_copy[BPPGPU_MAX_WAVES] = {nullptr}, ev_fork = nullptr, ev_join = nullptr;
  bool inputs_pending = false;         // staged with waves and not run yet: run() issues the copies
  struct Pending { const unsigned int * midx; const double * mbl; const bppgpu_partial_op * ops;
                   const unsigned int * rclv; const int * rsc; } pend = {nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_tables = nullptr;
  bool tables_on_device = false;       // the three offset tables of the last stage are still valid in d_in
  size_t tables_tm = 0, tables_to = 0;
  bool have_m = false, have_o = false;
  std::vector<unsigned int> last_mcounts, last_ocounts;
  unsigned int max_tips = 0;
  int su_slots = 0;                    // 4 states, trees of more than 16 tips: stack slots the staged op lists need (0 = unknown)
  unsigned int tip_words_rt = 1;       // packed tip words the 4-state kernel stages per cell
  unsigned int wave_pref = 0;          // 0 = automatic
  std::vector<unsigned int> h_tile_first;
  // launch configuration of the tree kernel, resolved once per (shared-memory size)
  size_t cfg_smem[2] = {0, 0}; int cfg_per_sm[2] = {0, 0}; unsigned int cfg_key[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};   // [SCALED_ONLY]
  unsigned long long overflow_epoch = 0;   // engine dirty_epoch at the last col_overflow check (20 states)
  bool flip_checked = false;           // bppgpu_batch_flip_indices: the loci have BPP's 2x buffer allocation
  bool last_scaled_only = false;       // instantiation of the last 4-state tree launch (kernel name reporting)
};

// ------------------------------------------------------------------------------------ helpers
static cudaEvent_t get_event(bppgpu_engine * e)
{
  std::lock_guard<std::mutex> lock(e->prof_mu);
  if (!e->event_pool.empty()) { cudaEvent_t ev = e->event_pool.back(); e->event_pool.pop_back(); return ev; }
  cudaEvent_t ev;
  CUDA_CHECK(cudaEventCreate(&ev));
  return ev;
}

struct ProfScope
{
  bppgpu_engine * e; cudaStream_t s; int kind; cudaEvent_t a = nullptr, b = nullptr;
  ProfScope(bppgpu_engine * e_, cudaStream_t s_, int kind_) : e(e_), s(s_), kind(kind_)
  {
    e->launches++;
    if (e->profiling) { a = get_event(e); b = get_event(e); cudaEventRecord(a, s); }
  }
  ~ProfScope()
  {
    if (e->profiling && a) { cudaEventRecord(b, s); std::lock_guard<std::mutex> lock(e->prof_mu); e->pending.push_back({a, b, kind}); }
  }
};

static void drain_profile(bppgpu_engine * e)
{
  std::lock_guard<std::mutex> lock(e->prof_mu);
  for (auto & p : e->pending)
  {
    cudaEventSynchronize(p.b);
    float ms = 0;
    cudaEventElapsedTime(&ms, p.a, p.b);
    e->prof_ms[p.kind] += ms; e->prof_n[p.kind]++;
    e->event_pool.push_back(p.a); e->event_pool.push_back(p.b);
  }
  e->pending.clear();
}

// symmetric eigen-decomposition, cyclic Jacobi.  Used when the host did not hand over its own
// pll_update_eigen result.  Rows of `vec` are the eigenvectors.
static void jacobi_eigen(std::vector<double> & a, int n, std::vector<double> & lam, std::vector<double> & vec)
{
  vec.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) vec[(size_t)i * n + i] = 1.0;      // columns = eigenvectors during sweeps
  for (int sweep = 0; sweep < 100; ++sweep)
  {
    double off = 0;
    for (int p = 0; p < n; ++p) for (int q = p + 1; q < n; ++q) off += a[(size_t)p * n + q] * a[(size_t)p * n + q];
    if (off < 1e-300) break;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q)
      {
        const double apq = a[(size_t)p * n + q];
        if (std::fabs(apq) < 1e-300) continue;
        const double app = a[(size_t)p * n + p], aqq = a[(size_t)q * n + q];
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k)
        {
          const double akp = a[(size_t)k * n + p], akq = a[(size_t)k * n + q];
          a[(size_t)k * n + p] = c * akp - s * akq;
          a[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k)
        {
          const double apk = a[(size_t)p * n + k], aqk = a[(size_t)q * n + k];
          a[(size_t)p * n + k] = c * apk - s * aqk;
          a[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k)
        {
          const double vkp = vec[(size_t)k * n + p], vkq = vec[(size_t)k * n + q];
          vec[(size_t)k * n + p] = c * vkp - s * vkq;
          vec[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  lam.resize(n);
  for (int i = 0; i < n; ++i) lam[i] = a[(size_t)i * n + i];
  // transpose so that rows are eigenvectors
  std::vector<double> t((size_t)n * n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) t[(size_t)i * n + j] = vec[(size_t)j * n + i];
  vec.swap(t);
}

// pll_update_eigen, core_pmatrix.c:239-297 (+ create_ratematrix :186-237): symmetrised rate matrix
// with mean rate 1; eigenvecs[i][j] = a[i][j]*sqrt(pi_j), inv_eigenvecs[i][j] = a[j][i]/sqrt(pi_i).
static void host_update_eigen(bppgpu_locus * l)
{
  const int S = (int)l->states;
  const int np = S * (S - 1) / 2;
  std::vector<double> p(l->h_subst.begin(), l->h_subst.begin() + np);
  if (p[np - 1] > 0.0) for (int i = 0; i < np; ++i) p[i] /= l->h_subst[np - 1];
  const std::vector<double> & f = l->h_freqs;
  std::vector<double> q((size_t)S * S, 0.0);
  int k = 0;
  for (int i = 0; i < S; ++i)
    for (int j = i + 1; j < S; ++j)
    {
      const double factor = p[k++];
      q[(size_t)i * S + j] = q[(size_t)j * S + i] = factor * std::sqrt(f[i] * f[j]);
      q[(size_t)i * S + i] -= factor * f[j];
      q[(size_t)j * S + j] -= factor * f[i];
    }
  double mean = 0;
  for (int i = 0; i < S; ++i) mean += f[i] * (-q[(size_t)i * S + i]);
  for (auto & v : q) v /= mean;
  std::vector<double> lam, a;
  jacobi_eigen(q, S, lam, a);
  l->h_evals = lam;
  l->h_evecs.assign((size_t)S * S, 0.0);
  l->h_ievecs.assign((size_t)S * S, 0.0);
  for (int i = 0; i < S; ++i)
    for (int j = 0; j < S; ++j)
    {
      l->h_evecs[(size_t)i * S + j] = a[(size_t)i * S + j] * std::sqrt(f[j]);
      l->h_ievecs[(size_t)i * S + j] = a[(size_t)j * S + i] / std::sqrt(f[i]);
    }
  l->eigen_valid = true;
  l->model_dirty = true, l->e->dirty_epoch++;
}

static inline void set_code(bppgpu_locus * l, unsigned tip, size_t site, unsigned int c)
{
  if (l->states == 4)
  {
    unsigned int & w = l->h_codes[site * l->dev.tip_words + (tip >> 3)];
    const unsigned sh = (tip & 7u) * 4;
    w = (w & ~(0xFu << sh)) | ((c & 0xFu) << sh);
  }
  else
  {
    l->h_codes[(size_t)tip * l->sites + site] = c;
    // column id for the tensor-core path: one-hot -> state, else one of up to 4 ambiguity columns
    unsigned int col = 0xFF;
    if (c && !(c & (c - 1))) col = (unsigned int)__builtin_ctz(c);
    else
    {
      for (unsigned int x = 0; x < l->n_ext_cols; ++x) if (l->h_colmask[x] == c) col = l->states + x;
      if (col == 0xFF)
      {
        if (l->n_ext_cols < 4) { l->h_colmask[l->n_ext_cols] = c; col = l->states + l->n_ext_cols; ++l->n_ext_cols; }
        else { l->col_overflow = true; col = 0; }
      }
    }
    l->h_cols[(size_t)tip * l->dev.cols_pitch + site] = (unsigned char)col;
  }
}
static inline unsigned int get_code(const bppgpu_locus * l, unsigned tip, size_t site)
{
  if (l->states == 4) return (l->h_codes[site * l->dev.tip_words + (tip >> 3)] >> ((tip & 7u) * 4)) & 0xFu;
  return l->h_codes[(size_t)tip * l->sites + site];
}

static size_t model_doubles(unsigned S, unsigned R) { return (size_t)S + R + R + 2 * (size_t)S * S + S + (size_t)S * (S - 1) / 2; }

// push dirty host mirrors (tip codes, dense flags, model block) of a locus to the device
static void locus_sync(bppgpu_locus * l, cudaStream_t s)
{
  if (l->codes_dirty)
  {
    CUDA_CHECK(cudaMemcpyAsync(l->dev.tip_codes, l->h_codes.data(), l->h_codes.size() * 4, cudaMemcpyHostToDevice, s));
    if (l->dev.tip_cols)
    {
      CUDA_CHECK(cudaMemcpyAsync(l->dev.tip_cols, l->h_cols.data(), l->h_cols.size(), cudaMemcpyHostToDevice, s));
      CUDA_CHECK(cudaMemcpyAsync(l->dev.colmask, l->h_colmask, sizeof(l->h_colmask), cudaMemcpyHostToDevice, s));
      if (l->dev.n_ext_cols != l->n_ext_cols)
      {
        l->dev.n_ext_cols = l->n_ext_cols;
        CUDA_CHECK(cudaMemcpyAsync(l->e->d_loci + l->id, &l->dev, sizeof(LocusDev), cudaMemcpyHostToDevice, s));
      }
    }
    l->codes_dirty = false;
  }
  if (l->flags_dirty)
  {
    CUDA_CHECK(cudaMemcpyAsync(l->dev.tip_is_dense, l->h_tip_dense_flag.data(), l->tips, cudaMemcpyHostToDevice, s));
    l->flags_dirty = false;
  }
  if (l->model_dirty)
  {
    const unsigned S = l->states, R = l->rate_cats;
    std::vector<double> blk(model_doubles(S, R));
    double * w = blk.data();
    memcpy(w, l->h_freqs.data(), S * 8); w += S;
    memcpy(w, l->h_rates.data(), R * 8); w += R;
    memcpy(w, l->h_rate_weights.data(), R * 8); w += R;
    memcpy(w, l->h_evecs.data(), (size_t)S * S * 8); w += (size_t)S * S;
    memcpy(w, l->h_ievecs.data(), (size_t)S * S * 8); w += (size_t)S * S;
    memcpy(w, l->h_evals.data(), S * 8); w += S;
    memcpy(w, l->h_subst.data(), (size_t)S * (S - 1) / 2 * 8);
    // pageable source: the copy is staged by the runtime before the call returns
    if (l->eigen_on_device && l->eigen_valid)
    {
      // the decomposition on the device is the current one (the host mirror is stale): leave it alone
      CUDA_CHECK(cudaMemcpyAsync(l->dev.freqs, blk.data(), (size_t)(S + 2 * R) * 8, cudaMemcpyHostToDevice, s));
      CUDA_CHECK(cudaMemcpyAsync(l->dev.subst, l->h_subst.data(), (size_t)S * (S - 1) / 2 * 8, cudaMemcpyHostToDevice, s));
    }
    else
    {
      CUDA_CHECK(cudaMemcpyAsync(l->dev.freqs, blk.data(), blk.size() * 8, cudaMemcpyHostToDevice, s));
      l->eigen_on_device = false;
    }
    l->model_dirty = false;
  }
}

// locus_update_matrices dispatch (locus.c:2426-2455): every DNA model but GTR has a closed form
static unsigned model_kind_of(const bppgpu_locus * l)
{
  if (l->dtype != BPPGPU_DATA_DNA) return MODEL_EIGEN;
  switch (l->model)
  {
    case BPPGPU_DNA_MODEL_JC69: return MODEL_JC69;
    case BPPGPU_DNA_MODEL_K80:  return MODEL_K80;
    case BPPGPU_DNA_MODEL_F81:  return MODEL_F81;
    case BPPGPU_DNA_MODEL_HKY:  return MODEL_HKY;
    case BPPGPU_DNA_MODEL_T92:  return MODEL_T92;
    case BPPGPU_DNA_MODEL_TN93: return MODEL_TN93;
    case BPPGPU_DNA_MODEL_F84:  return MODEL_F84;
    default: return MODEL_EIGEN;
  }
}
static bool model_is_closed_form(const bppgpu_locus * l) { return model_kind_of(l) != MODEL_EIGEN; }

// ------------------------------------------------------------------------------------ engine API
extern "C" int bppgpu_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" const char * bppgpu_last_error(void) { return g_last_error.c_str(); }
extern "C" void bppgpu_set_fatal_handler(void (*handler)(const char *)) { g_fatal_handler = handler; }
extern "C" const char * bppgpu_version(void) { return "bpp_b200 0.1 (sm_100a)"; }

extern "C" bppgpu_engine * bppgpu_engine_create(int device, unsigned int flags)
{
  int n = bppgpu_device_count();
  if (n <= 0) { fatal("no CUDA device available: the bpp_b200 engine has no CPU fallback"); return nullptr; }
  if (device < 0 || device >= n) { fatal("invalid device ordinal %d (have %d)", device, n); return nullptr; }
  CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) { fatal("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor); return nullptr; }
  bppgpu_engine * e = new bppgpu_engine();
  e->device = device;
  e->math = flags & 1u;
  e->sm_count = prop.multiProcessorCount;
  e->smem_optin = prop.sharedMemPerBlockOptin;
  e->smem_per_sm = prop.sharedMemPerMultiprocessor;
  e->log_threshold = std::log(BPPGPU_SCALE_THRESHOLD);      // core_likelihood.c:200 evaluates it with libm
  CUDA_CHECK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  return e;
}

extern "C" void bppgpu_engine_destroy(bppgpu_engine * e)
{
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  std::vector<bppgpu_locus *> ls = e->loci;
  for (bppgpu_locus * l : ls) if (l) bppgpu_locus_destroy(l);
  drain_profile(e);
  for (cudaEvent_t ev : e->event_pool) cudaEventDestroy(ev);
  if (e->d_loci) cudaFree(e->d_loci);
  e->arena.destroy();
  cudaStreamDestroy(e->stream);
  delete e;
}

extern "C" int bppgpu_engine_device(const bppgpu_engine * e) { return e->device; }
extern "C" void bppgpu_engine_set_math(bppgpu_engine * e, unsigned int m) { e->math = m & 1u; }
extern "C" void bppgpu_engine_synchronize(bppgpu_engine * e) { cudaSetDevice(e->device); CUDA_CHECK(cudaDeviceSynchronize()); }
extern "C" void * bppgpu_engine_stream(bppgpu_engine * e) { return (void *)e->stream; }
extern "C" unsigned long long bppgpu_engine_launch_count(const bppgpu_engine * e) { return e->launches.load(); }
extern "C" unsigned long long bppgpu_engine_bytes_allocated(const bppgpu_engine * e) { return e->arena.total; }
extern "C" void bppgpu_engine_set_profiling(bppgpu_engine * e, int on) { drain_profile(e); e->profiling = on != 0; }
extern "C" void bppgpu_engine_reset_profile(bppgpu_engine * e)
{
  drain_profile(e);
  for (int i = 0; i < BPPGPU_KERNEL_COUNT; ++i) { e->prof_ms[i] = 0; e->prof_n[i] = 0; }
}
extern "C" void bppgpu_engine_get_profile(bppgpu_engine * e, double * ms, unsigned long long * cnt)
{
  drain_profile(e);
  for (int i = 0; i < BPPGPU_KERNEL_COUNT; ++i) { if (ms) ms[i] = e->prof_ms[i]; if (cnt) cnt[i] = e->prof_n[i]; }
}

// ------------------------------------------------------------------------------------ locus API
static void engine_publish_locus(bppgpu_engine * e, bppgpu_locus * l)
{
  if (l->id >= e->loci_cap)
  {
    size_t ncap = std::max<size_t>(65536, e->loci_cap * 2);       // 9 MB of descriptors: growing (sync + free) stays rare
    while (ncap <= l->id) ncap *= 2;
    LocusDev * nd = nullptr;
    CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaMalloc(&nd, ncap * sizeof(LocusDev)));
    if (e->d_loci)
    {
      CUDA_CHECK(cudaMemcpy(nd, e->d_loci, e->loci_cap * sizeof(LocusDev), cudaMemcpyDeviceToDevice));
      CUDA_CHECK(cudaFree(e->d_loci));
    }
    e->d_loci = nd; e->loci_cap = ncap;
  }
  CUDA_CHECK(cudaMemcpy(e->d_loci + l->id, &l->dev, sizeof(LocusDev), cudaMemcpyHostToDevice));
}

extern "C" bppgpu_locus * bppgpu_locus_create(bppgpu_engine * e, unsigned int dtype, unsigned int model,
                                              unsigned int tips, unsigned int clv_buffers,
                                              unsigned int states, unsigned int sites,
                                              unsigned int rate_matrices, unsigned int prob_matrices,
                                              unsigned int rate_cats, unsigned int scale_buffers,
                                              unsigned int attributes)
{
  if (!e) { fatal("bppgpu_locus_create: no engine"); return nullptr; }
  if (states < 2 || states > 32) { fatal("unsupported number of states %u", states); return nullptr; }
  if (rate_matrices != 1) { fatal("rate_matrices must be 1 (method.c:4143)"); return nullptr; }
  if (sites == 0 || tips < 2 || rate_cats == 0) { fatal("invalid locus dimensions"); return nullptr; }
  const unsigned int user_cats = rate_cats;
  // 3 -> 4, 5..7 -> 8 categories on the device (BPPGPU_PAD_CATS=0: keep the count and take the generic kernel)
  static const bool pad_cats = !(getenv("BPPGPU_PAD_CATS") && atoi(getenv("BPPGPU_PAD_CATS")) == 0);
  if (pad_cats && (states == 4 || states == S20) && rate_cats < 8 && (rate_cats & (rate_cats - 1)) != 0)
    rate_cats = rate_cats < 4 ? 4 : 8;
  std::lock_guard<std::mutex> lock(e->mu);
  CUDA_CHECK(cudaSetDevice(e->device));
  bppgpu_locus * l = new bppgpu_locus();
  l->e = e;
  l->dtype = dtype; l->model = model; l->tips = tips; l->clv_buffers = clv_buffers; l->states = states;
  l->sites = sites; l->rate_matrices = rate_matrices; l->prob_matrices = prob_matrices;
  l->rate_cats = rate_cats; l->user_cats = user_cats; l->scale_buffers = scale_buffers; l->attributes = attributes;
  const size_t S = states, R = rate_cats, P = sites;
  const size_t clv_doubles = P * R * S;
  LocusDev & d = l->dev;
  memset(&d, 0, sizeof(d));
  l->b_clv = clv_buffers * clv_doubles * 8;
  const size_t TW = (tips + 7) / 8;
  l->b_codes = (S == 4) ? P * TW * 4 : (size_t)tips * P * 4;
  l->b_flags = tips;
  l->b_pmat = (size_t)prob_matrices * R * S * S * 8;
  l->b_scale = (size_t)scale_buffers * P * 4;
  l->b_weights = P * 4;
  l->b_model = model_doubles(states, rate_cats) * 8;
  d.clv = (double *)e->arena.alloc(l->b_clv);
  d.tip_codes = e->arena.alloc(l->b_codes);
  if (S != 4)
  {
    d.cols_pitch = (P + 15u) & ~15u;
    l->b_cols = (size_t)tips * d.cols_pitch; l->b_colmask = 16;
    d.tip_cols = (unsigned char *)e->arena.alloc(l->b_cols);
    d.colmask = (unsigned int *)e->arena.alloc(l->b_colmask);
    l->h_cols.assign(l->b_cols, 0);
  }
  d.tip_is_dense = (unsigned char *)e->arena.alloc(l->b_flags);
  d.pmat = (double *)e->arena.alloc(l->b_pmat);
  d.scale = scale_buffers ? (unsigned int *)e->arena.alloc(l->b_scale) : nullptr;
  d.weights = (unsigned int *)e->arena.alloc(l->b_weights);
  double * m = (double *)e->arena.alloc(l->b_model);
  d.freqs = m; d.rates = m + S; d.rate_weights = d.rates + R; d.eigenvecs = d.rate_weights + R;
  d.inv_eigenvecs = d.eigenvecs + S * S; d.eigenvals = d.inv_eigenvecs + S * S; d.subst = d.eigenvals + S;
  if (!d.clv || !d.tip_codes || !d.tip_is_dense || !d.pmat || (scale_buffers && !d.scale) || !d.weights || !m ||
      (S != 4 && (!d.tip_cols || !d.colmask)))
  {
    // the arena's fatal() has reported the failed allocation; with a handler that returns, hand back nothing rather
    // than a half-built locus
    Arena & a = e->arena;
    a.release(d.clv, l->b_clv); a.release(d.tip_codes, l->b_codes); a.release(d.tip_cols, l->b_cols);
    a.release(d.colmask, l->b_colmask); a.release(d.tip_is_dense, l->b_flags); a.release(d.pmat, l->b_pmat);
    a.release(d.scale, l->b_scale); a.release(d.weights, l->b_weights); a.release(m, l->b_model);
    delete l;
    return nullptr;
  }
  d.clv_stride = clv_doubles;
  // 20-state loci with several categories keep their CLVs category-major (common.cuh, LocusDev::site_stride)
  const bool cat_major = states == S20 && rate_cats > 1 && rate_cats <= 8 && (rate_cats & (rate_cats - 1)) == 0;
  d.site_stride = cat_major ? (unsigned)S : (unsigned)(R * S);
  d.cat_stride = cat_major ? (unsigned)(P * S) : (unsigned)S;
  d.tips = tips; d.sites = sites; d.states = states; d.rate_cats = rate_cats;
  d.clv_buffers = clv_buffers; d.prob_matrices = prob_matrices; d.scale_buffers = scale_buffers;
  d.model_kind = model_kind_of(l);
  d.tip_words = (unsigned)TW;
  // zero what the reference zeroes (locus.c:745-755,765-771,859-867); weights default to 1 (:852)
  CUDA_CHECK(cudaMemsetAsync(d.clv, 0, l->b_clv, e->stream));
  CUDA_CHECK(cudaMemsetAsync(d.pmat, 0, l->b_pmat, e->stream));
  if (d.scale) CUDA_CHECK(cudaMemsetAsync(d.scale, 0, l->b_scale, e->stream));
  CUDA_CHECK(cudaMemsetAsync(d.tip_codes, 0, l->b_codes, e->stream));
  CUDA_CHECK(cudaMemsetAsync(d.tip_is_dense, 0, l->b_flags, e->stream));
  {
    std::vector<unsigned int> ones(P, 1u);
    CUDA_CHECK(cudaMemcpyAsync(d.weights, ones.data(), P * 4, cudaMemcpyHostToDevice, e->stream));
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
  }
  l->h_codes.assign(l->b_codes / 4, 0);
  l->h_tip_dense_flag.assign(tips, 0);
  l->h_freqs.assign(S, 0.0);                       // zero like locus.c:826 until pll_set_frequencies
  l->h_subst.assign(S * (S - 1) / 2, 0.0);
  l->h_rates.assign(R, 1.0);
  l->h_rate_weights.assign(R, 0.0);
  for (unsigned r = 0; r < user_cats; ++r) l->h_rate_weights[r] = 1.0 / (double)user_cats;    // locus.c:845-848
  l->h_evecs.assign(S * S, 0.0); l->h_ievecs.assign(S * S, 0.0); l->h_evals.assign(S, 0.0);
  l->model_dirty = true, l->e->dirty_epoch++;
  if (!e->free_ids.empty()) { l->id = e->free_ids.back(); e->free_ids.pop_back(); e->loci[l->id] = l; }
  else { l->id = (unsigned)e->loci.size(); e->loci.push_back(l); }
  engine_publish_locus(e, l);
  return l;
}

extern "C" void bppgpu_locus_destroy(bppgpu_locus * l)
{
  if (!l) return;
  bppgpu_engine * e = l->e;
  if (l->self_batch) { bppgpu_batch * b = l->self_batch; l->self_batch = nullptr; bppgpu_batch_destroy(b); }
  std::lock_guard<std::mutex> lock(e->mu);
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  Arena & a = e->arena;
  a.release(l->dev.clv, l->b_clv);
  a.release(l->dev.tip_dense, l->b_tipdense);
  a.release(l->dev.tip_codes, l->b_codes);
  a.release(l->dev.tip_cols, l->b_cols);
  a.release(l->dev.colmask, l->b_colmask);
  a.release(l->dev.tip_is_dense, l->b_flags);
  a.release(l->dev.pmat, l->b_pmat);
  a.release(l->dev.scale, l->b_scale);
  a.release(l->dev.weights, l->b_weights);
  a.release(l->dev.freqs, l->b_model);
  a.release(l->dev.dip_off, l->b_dip_off);
  a.release(l->dev.dip_map, l->b_dip_map);
  a.release(l->dev.dip_weights, l->b_dip_w);
  e->loci[l->id] = nullptr;
  e->free_ids.push_back(l->id);
  delete l;
}

extern "C" int bppgpu_set_tip_states(bppgpu_locus * l, unsigned int tip, const unsigned int * map, const char * seq)
{
  if (tip >= l->tips) { fatal("tip index %u out of range", tip); return BPPGPU_FAILURE; }
  const size_t P = l->sites;
  for (size_t i = 0; i < P; ++i)
  {
    const unsigned int c = map[(int)(unsigned char)seq[i]];
    if (c == 0) { fatal("Illegal state code in tip \"%c\"", seq[i]); return BPPGPU_FAILURE; }   // locus.c:538
    set_code(l, tip, i, c);
  }
  l->codes_dirty = true, l->e->dirty_epoch++;
  if (l->h_tip_dense_flag[tip]) { l->h_tip_dense_flag[tip] = 0; l->flags_dirty = true, l->e->dirty_epoch++; }
  return BPPGPU_SUCCESS;
}

extern "C" int bppgpu_set_tip_clv(bppgpu_locus * l, unsigned int tip, const double * clv, int padding)
{
  (void)padding;                                   // states_padded == states
  if (tip >= l->tips) { fatal("tip index %u out of range", tip); return BPPGPU_FAILURE; }
  const size_t P = l->sites, S = l->states, R = l->rate_cats;
  bool binary = true;
  for (size_t i = 0; i < P * S && binary; ++i) binary = (clv[i] == 0.0 || clv[i] == 1.0);
  if (binary)
  {
    for (size_t i = 0; i < P; ++i)                 // an all-zero tip vector is legal here (lnL = -inf)
    {
      unsigned int c = 0;
      for (size_t j = 0; j < S; ++j) if (clv[i * S + j] == 1.0) c |= 1u << j;
      set_code(l, tip, i, c);
    }
    l->codes_dirty = true, l->e->dirty_epoch++;
    if (l->h_tip_dense_flag[tip]) { l->h_tip_dense_flag[tip] = 0; l->flags_dirty = true, l->e->dirty_epoch++; }
    return BPPGPU_SUCCESS;
  }
  bppgpu_engine * e = l->e;
  std::lock_guard<std::mutex> lock(e->mu);
  CUDA_CHECK(cudaSetDevice(e->device));
  if (!l->dev.tip_dense)
  {
    l->b_tipdense = (size_t)l->tips * P * R * S * 8;
    l->dev.tip_dense = (double *)e->arena.alloc(l->b_tipdense);
    engine_publish_locus(e, l);
  }
  std::vector<double> full(P * R * S);             // replicate over categories, locus.c:609-616
  for (size_t i = 0; i < P; ++i) for (size_t r = 0; r < R; ++r) memcpy(&full[i * l->dev.site_stride + r * l->dev.cat_stride], clv + i * S, S * 8);
  CUDA_CHECK(cudaMemcpy(l->dev.tip_dense + (size_t)tip * P * R * S, full.data(), full.size() * 8, cudaMemcpyHostToDevice));
  l->h_tip_dense_flag[tip] = 1; l->flags_dirty = true, l->e->dirty_epoch++;
  return BPPGPU_SUCCESS;
}

extern "C" void bppgpu_set_pattern_weights(bppgpu_locus * l, const unsigned int * w)
{
  CUDA_CHECK(cudaSetDevice(l->e->device));
  CUDA_CHECK(cudaDeviceSynchronize());           // a running batch may still be reading the old weights
  CUDA_CHECK(cudaMemcpy(l->dev.weights, w, (size_t)l->sites * 4, cudaMemcpyHostToDevice));
}
extern "C" void bppgpu_set_frequencies(bppgpu_locus * l, unsigned int idx, const double * f)
{
  if (idx != 0) { fatal("freqs_index must be 0"); return; }
  l->h_freqs.assign(f, f + l->states); l->eigen_valid = false; l->model_dirty = true, l->e->dirty_epoch++;   // locus.c:889-897
}
extern "C" void bppgpu_set_subst_params(bppgpu_locus * l, unsigned int idx, const double * p)
{
  if (idx != 0) { fatal("params_index must be 0"); return; }
  l->h_subst.assign(p, p + l->states * (l->states - 1) / 2); l->eigen_valid = false; l->model_dirty = true, l->e->dirty_epoch++;    // locus.c:877-887
}
extern "C" void bppgpu_set_category_rates(bppgpu_locus * l, const double * r)
{
  l->h_rates.assign(l->rate_cats, r[0]);                     // (padding categories: copies of category 0)
  std::copy(r, r + l->user_cats, l->h_rates.begin());
  l->model_dirty = true, l->e->dirty_epoch++;
}
extern "C" void bppgpu_set_category_weights(bppgpu_locus * l, const double * w)
{
  l->h_rate_weights.assign(l->rate_cats, 0.0);               // (padding categories weigh nothing)
  std::copy(w, w + l->user_cats, l->h_rate_weights.begin());
  l->model_dirty = true, l->e->dirty_epoch++;
}
extern "C" void bppgpu_set_eigen(bppgpu_locus * l, unsigned int idx, const double * ev, const double * iev, const double * lam)
{
  if (idx != 0) { fatal("params_index must be 0"); return; }
  const size_t S = l->states;
  l->h_evecs.assign(ev, ev + S * S); l->h_ievecs.assign(iev, iev + S * S); l->h_evals.assign(lam, lam + S);
  l->eigen_valid = true; l->eigen_on_device = false; l->model_dirty = true, l->e->dirty_epoch++;
}
extern "C" void bppgpu_get_eigen(bppgpu_locus * l, unsigned int idx, double * ev, double * iev, double * lam)
{
  (void)idx;
  const size_t S = l->states;
  if (!l->eigen_valid) { host_update_eigen(l); l->eigen_on_device = false; }
  else if (l->eigen_on_device)
  {
    // decomposed by model_update_kernel: fetch it (the kernel ran on the stream of the batch that synced the locus;
    // a device-wide synchronisation is the simple, rarely needed answer)
    CUDA_CHECK(cudaSetDevice(l->e->device));
    CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaMemcpy(l->h_evecs.data(), l->dev.eigenvecs, S * S * 8, cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaMemcpy(l->h_ievecs.data(), l->dev.inv_eigenvecs, S * S * 8, cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaMemcpy(l->h_evals.data(), l->dev.eigenvals, S * 8, cudaMemcpyDeviceToHost));
    l->eigen_on_device = false;
  }
  memcpy(ev, l->h_evecs.data(), S * S * 8); memcpy(iev, l->h_ievecs.data(), S * S * 8); memcpy(lam, l->h_evals.data(), S * 8);
}

extern "C" int bppgpu_set_diploid(bppgpu_locus * l, unsigned int unphased, const unsigned long * cnt,
                                  const unsigned long * mapping, unsigned long maplen, const unsigned int * weights)
{
  bppgpu_engine * e = l->e;
  std::lock_guard<std::mutex> lock(e->mu);
  CUDA_CHECK(cudaSetDevice(e->device));
  std::vector<unsigned long long> off(unphased + 1, 0), mp(maplen);
  for (unsigned int i = 0; i < unphased; ++i) off[i + 1] = off[i] + cnt[i];
  if (off[unphased] != maplen) { fatal("diploid mapping length mismatch"); return BPPGPU_FAILURE; }
  for (unsigned long i = 0; i < maplen; ++i) mp[i] = mapping[i];
  e->arena.release(l->dev.dip_off, l->b_dip_off); e->arena.release(l->dev.dip_map, l->b_dip_map);
  e->arena.release(l->dev.dip_weights, l->b_dip_w);
  l->b_dip_w = std::max<size_t>(4, (size_t)unphased * 4);
  l->dev.dip_weights = (unsigned int *)e->arena.alloc(l->b_dip_w);
  CUDA_CHECK(cudaMemcpy(l->dev.dip_weights, weights, (size_t)unphased * 4, cudaMemcpyHostToDevice));
  l->b_dip_off = off.size() * 8; l->b_dip_map = std::max<size_t>(8, mp.size() * 8);
  l->dev.dip_off = (unsigned long long *)e->arena.alloc(l->b_dip_off);
  l->dev.dip_map = (unsigned long long *)e->arena.alloc(l->b_dip_map);
  CUDA_CHECK(cudaMemcpy(l->dev.dip_off, off.data(), off.size() * 8, cudaMemcpyHostToDevice));
  if (!mp.empty()) CUDA_CHECK(cudaMemcpy(l->dev.dip_map, mp.data(), mp.size() * 8, cudaMemcpyHostToDevice));
  l->dev.unphased = unphased;
  engine_publish_locus(e, l);
  e->dirty_epoch++;
  return BPPGPU_SUCCESS;
}

// ------------------------------------------------------------------------------------ batch
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// shared memory of a 4-state launch (S4Layout::bytes) for run-time RL / cells per thread
static size_t s4_smem_bytes(unsigned RL, unsigned cpt, int slots, unsigned cap, unsigned words)
{
  switch (RL * 10 + cpt)
  {
    case 11: return S4Layout<1, 1>::bytes(slots, cap, words);
    case 12: return S4Layout<1, 2>::bytes(slots, cap, words);
    case 14: return S4Layout<1, 4>::bytes(slots, cap, words);
    case 21: return S4Layout<2, 1>::bytes(slots, cap, words);
    case 22: return S4Layout<2, 2>::bytes(slots, cap, words);
    case 24: return S4Layout<2, 4>::bytes(slots, cap, words);
    case 41: return S4Layout<4, 1>::bytes(slots, cap, words);
    case 42: return S4Layout<4, 2>::bytes(slots, cap, words);
    case 44: return S4Layout<4, 4>::bytes(slots, cap, words);
    case 81: return S4Layout<8, 1>::bytes(slots, cap, words);
    case 82: return S4Layout<8, 2>::bytes(slots, cap, words);
    case 84: return S4Layout<8, 4>::bytes(slots, cap, words);
    default: return ~(size_t)0;
  }
}

// stack slots a full pass over a T-tip tree is given: ceil(log2 T) - 1 covers a balanced tree in Sethi-Ullman order
static int s4_slots_formula(unsigned maxT)
{
  int slots = 1;
  while ((1u << slots) < maxT) ++slots;
  return std::min(std::max(slots - 1, 1), 6);
}

static void batch_launch_cfg(bppgpu_batch * b)
{
  // all loci of a batch share states / rate_cats (checked at creation)
  const bppgpu_locus * l0 = b->loci[0];
  const unsigned R = l0->rate_cats;
  const bool pow2 = (R & (R - 1)) == 0 && R <= 8;
  bool overflow = false;
  for (auto * l : b->loci) overflow = overflow || l->col_overflow;
  if (l0->states == 4 && pow2) { b->kernel_kind = 0; b->RL = R; }
  else if (l0->states == S20 && pow2 && !overflow) { b->kernel_kind = 2; b->RL = R; }
  else { b->kernel_kind = 1; b->RL = 1; }
  size_t cells = 0;
  for (auto * l : b->loci) cells += (size_t)l->sites * (b->kernel_kind != 1 ? R : 1);
  const size_t mean = cells / b->n;
  // 4-state kernel: 2 cells per thread once the loci are big enough to fill such tiles
  b->cpt = (b->kernel_kind == 0 && mean >= 2 * TREE_NT) ? 2 : 1;
  // with rate categories the shared-memory pipe limits the kernel: 4 cells per thread halve the P-matrix loads per cell
  if (b->kernel_kind == 0 && R >= 4 && mean >= 7 * (TREE_NT / 2)) b->cpt = 4;
  if (const char * ev = getenv("BPPGPU_CPT"))                 // tuning knob
    if (b->kernel_kind == 0 && (atoi(ev) == 1 || atoi(ev) == 2 || atoi(ev) == 4)) b->cpt = (unsigned)atoi(ev);
  // big trees: lookup tables, stack slots and the packed tip words of the tile's sites share the SM's shared memory.
  // Cells per thread give way until the slots a tree of this size typically needs fit: random (coalescent) trees need
  // ceil(log2 T) - 2 slots or fewer (a balanced tree one more; the run sizes the stack from the staged lists' real
  // Sethi-Ullman need, tree_s4_slots), and a list that finds no slot leaves the fast path.
  if (b->kernel_kind == 0)
  {
    unsigned maxT = 0;
    for (auto * l : b->loci) maxT = std::max(maxT, l->tips);
    const unsigned words = std::min<unsigned>((maxT + 7) / 8, (unsigned)S4_MAX_TIP_WORDS);
    const unsigned cap = std::min<unsigned>((unsigned)lut_cap((int)R), std::max(4u, (maxT + 3u) & ~3u));
    const int est = maxT > 16 ? std::max(s4_slots_formula(maxT) - 1, 1) : s4_slots_formula(maxT);
    while (b->cpt > 1 && s4_smem_bytes(R, b->cpt, est, cap, words) + 1024 > b->e->smem_optin) b->cpt /= 2;
  }
  b->tile_threads = b->kernel_kind == 0 ? TREE_NT : (b->kernel_kind == 2 ? S20_NT : 128);
}

extern "C" bppgpu_batch * bppgpu_batch_create(bppgpu_engine * e, unsigned int n, bppgpu_locus * const * loci)
{
  if (!e || n == 0) { fatal("bppgpu_batch_create: empty batch"); return nullptr; }
  CUDA_CHECK(cudaSetDevice(e->device));
  bppgpu_batch * b = new bppgpu_batch();
  b->e = e; b->n = n; b->loci.assign(loci, loci + n);
  for (unsigned i = 0; i < n; ++i)
  {
    if (!loci[i] || loci[i]->e != e) { fatal("batch locus %u belongs to another engine", i); delete b; return nullptr; }
    if (loci[i]->states != loci[0]->states || loci[i]->rate_cats != loci[0]->rate_cats)
    { fatal("all loci of a batch must share states and rate_cats"); delete b; return nullptr; }
  }
  CUDA_CHECK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  b->own_stream = true;
  batch_launch_cfg(b);
  std::vector<unsigned int> ids(n), tile_locus, tile_cell0, tile_first(n + 1, 0);
  std::vector<TileDesc> tiles;
  std::vector<unsigned long long> site_off(n + 1, 0);
  if (b->kernel_kind == 2)
  {
    // category-major tiles with the matrices of one (locus, category) staged in shared memory, unless the tree is
    // too big for them to fit; a batch with a locus that scales per site runs the cluster form of the same kernel
    unsigned maxT = 0; bool scaled = false;
    for (unsigned i = 0; i < n; ++i) { maxT = std::max(maxT, loci[i]->tips); scaled = scaled || loci[i]->scale_buffers > 0; }
    b->s20_cat = s20t_smem_bytes(2 * maxT - 1, maxT, 0, scaled) + 1024 <= e->smem_optin;
    b->s20_scaled = b->s20_cat && scaled;
    if (const char * ev = getenv("BPPGPU_S20T")) b->s20_cat = b->s20_cat && atoi(ev) != 0;      // tuning knob
    if (!b->s20_cat) b->s20_scaled = false;
  }
  std::vector<unsigned long long> soff(n);
  size_t sbytes = 0;
  for (unsigned i = 0; i < n; ++i)
  {
    const bppgpu_locus * l = loci[i];
    ids[i] = l->id;
    const unsigned cells = l->sites * (b->kernel_kind != 1 ? b->RL : 1);
    tile_first[i] = (unsigned)tile_locus.size();
    const unsigned tile_cells = b->kernel_kind == 2 ? (unsigned)S20_TILE : b->tile_threads * b->cpt;
    site_off[i + 1] = site_off[i] + l->sites;
    if (b->s20_cat)
    {
      // tile word: site block | site blocks of the locus << 12 | category << 24 (scaled: the cluster rank is the category)
      const unsigned nsb = (l->sites + S20T_SITES - 1) / S20T_SITES;
      if (nsb >= 4096) { fatal("20-state locus with %u patterns: more than the category-major kernel's tile word holds", l->sites); delete b; return nullptr; }
      for (unsigned c = 0; c < (b->s20_scaled ? 1u : b->RL); ++c)
        for (unsigned sb = 0; sb < nsb; ++sb) { tile_locus.push_back(i); tile_cell0.push_back(sb | (nsb << 12) | (c << 24)); }
    }
    else
    for (unsigned c = 0; c < cells; c += tile_cells)
    {
      tile_locus.push_back(i); tile_cell0.push_back(c);
      if (b->kernel_kind == 0)
      {
        TileDesc td;
        td.tipwords = (const unsigned int *)l->dev.tip_codes; td.weights = l->dev.weights;
        td.locus = i; td.cell0 = c; td.tip_words = l->dev.tip_words; td.ncell = cells;
        tiles.push_back(td);
      }
    }
    soff[i] = sbytes; sbytes += align_up((size_t)l->clv_buffers * 5, 16);
  }
  tile_first[n] = (unsigned)tile_locus.size();
  b->n_tiles = (unsigned)tile_locus.size();
  b->scratch_bytes = sbytes;
  CUDA_CHECK(cudaMalloc(&b->d_batch_locus, n * 4));
  CUDA_CHECK(cudaMalloc(&b->d_tile_locus, b->n_tiles * 4));
  CUDA_CHECK(cudaMalloc(&b->d_tile_cell0, b->n_tiles * 4));
  CUDA_CHECK(cudaMalloc(&b->d_tile_first, (n + 1) * 4));
  CUDA_CHECK(cudaMalloc(&b->d_scratch_off, n * 8));
  CUDA_CHECK(cudaMalloc(&b->d_scratch, std::max<size_t>(16, sbytes)));
  CUDA_CHECK(cudaMalloc(&b->d_plan_count, n * 4));
  CUDA_CHECK(cudaMalloc(&b->d_tile_partial, b->n_tiles * 8));
  CUDA_CHECK(cudaMalloc(&b->d_lnl, (n + 1) * 8));
  b->d_lnl_sum = b->d_lnl + n;
  CUDA_CHECK(cudaMalloc(&b->d_block_sums, ((n + 255) / 256) * 8));
  CUDA_CHECK(cudaMalloc(&b->d_counter, 4));
  CUDA_CHECK(cudaMemset(b->d_counter, 0, 4));
  CUDA_CHECK(cudaHostAlloc(&b->h_out, (n + 1) * 8, cudaHostAllocDefault));
  CUDA_CHECK(cudaMemcpy(b->d_batch_locus, ids.data(), n * 4, cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(b->d_tile_locus, tile_locus.data(), b->n_tiles * 4, cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(b->d_tile_cell0, tile_cell0.data(), b->n_tiles * 4, cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(b->d_tile_first, tile_first.data(), (n + 1) * 4, cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(b->d_scratch_off, soff.data(), n * 8, cudaMemcpyHostToDevice));
  if (b->kernel_kind == 0)
  {
    CUDA_CHECK(cudaMalloc(&b->d_tiles, tiles.size() * sizeof(TileDesc)));
    CUDA_CHECK(cudaMemcpy(b->d_tiles, tiles.data(), tiles.size() * sizeof(TileDesc), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMalloc(&b->d_tile_blk, tiles.size() * 16));
  }
  if (b->kernel_kind == 2) CUDA_CHECK(cudaMalloc(&b->d_tile_blk, (size_t)b->n_tiles * 16));
  if (b->s20_cat)
  {
    CUDA_CHECK(cudaMalloc(&b->d_rootdot, (size_t)site_off[n] * b->RL * 8));
    CUDA_CHECK(cudaMalloc(&b->d_site_off, (size_t)(n + 1) * 8));
    CUDA_CHECK(cudaMemcpy(b->d_site_off, site_off.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice));
    if (b->s20_scaled) CUDA_CHECK(cudaMalloc(&b->d_rootsc, (size_t)site_off[n] * 4));
  }
  CUDA_CHECK(cudaMemset(b->d_plan_count, 0, n * 4));
  CUDA_CHECK(cudaMemset(b->d_tile_partial, 0, b->n_tiles * 8));
  CUDA_CHECK(cudaEventCreate(&b->t0));
  CUDA_CHECK(cudaEventCreate(&b->t1));
  CUDA_CHECK(cudaEventCreateWithFlags(&b->ev_inputs, cudaEventDisableTiming));
  if (b->kernel_kind == 0)
  {
    CUDA_CHECK(cudaMalloc(&b->d_class, 8));
    CUDA_CHECK(cudaHostAlloc(&b->h_class, 8, cudaHostAllocDefault));
    b->h_class[0] = b->h_class[1] = 0;
    for (auto & ev : b->ev_class) CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  }
  b->h_tile_first = tile_first;
  for (unsigned i = 0; i < n; ++i) b->max_tips = std::max(b->max_tips, loci[i]->tips);
  b->tip_words_rt = std::min<unsigned>((b->max_tips + 7) / 8, (unsigned)S4_MAX_TIP_WORDS);
  if (b->kernel_kind == 0 && n > 1)          // the per-locus API keeps a batch of one per locus: no side streams for those
  {
    CUDA_CHECK(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&b->alt_stream, cudaStreamNonBlocking));
    for (auto & ev : b->ev_copy) CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&b->ev_tables, cudaEventDisableTiming));
    if (const char * ev = getenv("BPPGPU_WAVES")) b->wave_pref = (unsigned)std::min(std::max(atoi(ev), 0), BPPGPU_MAX_WAVES);
  }
  return b;
}

extern "C" void bppgpu_batch_destroy(bppgpu_batch * b)
{
  if (!b) return;
  cudaSetDevice(b->e->device);
  cudaStreamSynchronize(b->stream);
  drain_profile(b->e);
  cudaFree(b->d_batch_locus); cudaFree(b->d_tile_locus); cudaFree(b->d_tile_cell0); cudaFree(b->d_tile_first);
  cudaFree(b->d_scratch_off); cudaFree(b->d_scratch); cudaFree(b->d_plan_count); cudaFree(b->d_tile_partial);
  cudaFree(b->d_lnl); cudaFree(b->d_in); cudaFree(b->d_plan); cudaFree(b->d_persite);
  cudaFree(b->d_rootdot); cudaFree(b->d_rootsc); cudaFree(b->d_site_off);
  cudaFree(b->d_blocks_par[0]); cudaFree(b->d_blocks_par[1]); cudaFree(b->d_tiles); cudaFree(b->d_tile_blk); cudaFree(b->d_block_sums); cudaFree(b->d_counter);
  cudaFreeHost(b->h_out); cudaFreeHost(b->h_in);
  if (b->d_class) { cudaFree(b->d_class); cudaFreeHost(b->h_class); for (auto & ev : b->ev_class) cudaEventDestroy(ev); }
  if (b->h_model) cudaFreeHost(b->h_model);
  cudaFree(b->d_model); cudaFree(b->d_eig_scratch);
  cudaEventDestroy(b->t0); cudaEventDestroy(b->t1); cudaEventDestroy(b->ev_inputs);
  if (b->copy_stream)
  {
    cudaStreamSynchronize(b->copy_stream); cudaStreamSynchronize(b->alt_stream);
    for (auto & ev : b->ev_copy) cudaEventDestroy(ev);
    cudaEventDestroy(b->ev_fork); cudaEventDestroy(b->ev_join); cudaEventDestroy(b->ev_tables);
    cudaStreamDestroy(b->copy_stream); cudaStreamDestroy(b->alt_stream);
  }
  if (b->own_stream) cudaStreamDestroy(b->stream);
  delete b;
}

extern "C" unsigned int bppgpu_batch_size(const bppgpu_batch * b) { return b->n; }
// name of the tree kernel instantiation this batch launches (bench.py ties its roofline and the committed ncu
// capture to it)
extern "C" const char * bppgpu_batch_kernel_name(bppgpu_batch * b)
{
  static thread_local char buf[96];
  const bool exact = b->e->math == BPPGPU_MATH_EXACT;
  if (b->kernel_kind == 0) snprintf(buf, sizeof(buf), "tree_kernel_s4<%u,%s,%u,%s>", b->RL, exact ? "true" : "false", b->cpt, b->last_scaled_only ? "scaled_only" : "all_paths");
  else if (b->kernel_kind == 2) snprintf(buf, sizeof(buf), b->s20_cat ? (b->s20_scaled ? "tree_kernel_s20t<%u,true>" : "tree_kernel_s20t<%u,false>") : "tree_kernel_s20<%u>", b->RL);
  else snprintf(buf, sizeof(buf), "tree_kernel_generic<%s>", exact ? "true" : "false");
  return buf;
}
extern "C" int bppgpu_batch_plan_stats(bppgpu_batch * b, unsigned int out[8])
{
  if (!b || b->kernel_kind != 0 || !b->d_blocks || !b->tables_on_device) return BPPGPU_FAILURE;
  CUDA_CHECK(cudaSetDevice(b->e->device));
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  const unsigned long long * boff = (const unsigned long long *)(b->h_in + b->o_blk_off);
  for (int k = 0; k < 8; ++k) out[k] = 0;
  LocusHdr H;
  for (unsigned i = 0; i < b->n; ++i)
  {
    CUDA_CHECK(cudaMemcpy(&H, b->d_blocks + boff[i], sizeof(H), cudaMemcpyDeviceToHost));
    if (H.flags & HDR_FAST)
    {
      out[0]++;
      if (H.n_chunks == 1 && (H.flags & HDR_SIMPLE)) out[1]++;
      else if (H.n_chunks == 1 && (H.flags & HDR_NOHBM)) out[2]++;
    }
    else out[3]++;
    out[4] = std::max(out[4], H.n_chunks);
  }
  out[5] = (unsigned)b->slots; out[6] = b->cpt; out[7] = (unsigned)b->tree_smem;
  return BPPGPU_SUCCESS;
}
extern "C" void bppgpu_batch_set_waves(bppgpu_batch * b, unsigned int waves)
{
  b->wave_pref = waves > BPPGPU_MAX_WAVES ? BPPGPU_MAX_WAVES : waves;
}
extern "C" void * bppgpu_batch_lnl_sum_dev(bppgpu_batch * b) { return (void *)b->d_lnl_sum; }
extern "C" void * bppgpu_batch_stream(bppgpu_batch * b) { return (void *)b->stream; }
extern "C" void bppgpu_batch_synchronize(bppgpu_batch * b) { CUDA_CHECK(cudaStreamSynchronize(b->stream)); }
extern "C" void bppgpu_batch_timer_start(bppgpu_batch * b) { CUDA_CHECK(cudaEventRecord(b->t0, b->stream)); }
extern "C" double bppgpu_batch_timer_stop_ms(bppgpu_batch * b)
{
  CUDA_CHECK(cudaEventRecord(b->t1, b->stream));
  CUDA_CHECK(cudaEventSynchronize(b->t1));
  float ms = 0;
  CUDA_CHECK(cudaEventElapsedTime(&ms, b->t0, b->t1));
  return ms;
}

// tip-slot capacity of a 4-state batch's launches: any op list over a T-tip tree that visits a node at most once has
// at most T tip or HBM-resident children; lists that need more simply take more chunks (the blocks are sized for it)
static unsigned s4_lut_cap_rt(const bppgpu_batch * b)
{
  return std::min<unsigned>((unsigned)lut_cap((int)b->RL), std::max(4u, (b->max_tips + 3u) & ~3u));
}

// Copy the slice of the staged step that belongs to loci [i0, i1) to the device on stream cs.
static void batch_issue_copies(bppgpu_batch * b, unsigned i0, unsigned i1, cudaStream_t cs)
{
  const bppgpu_batch::Pending & q = b->pend;
  const unsigned int * moff = (const unsigned int *)(b->h_in + b->o_mat_off);
  const unsigned int * ooff = (const unsigned int *)(b->h_in + b->o_op_off);
  auto put = [&](size_t dst, const void * from, size_t bytes)
  {
    if (bytes) CUDA_CHECK(cudaMemcpyAsync(b->d_in + dst, from, bytes, cudaMemcpyHostToDevice, cs));
  };
  if (b->all_in_blob && i0 == 0 && i1 == b->n)
  {
    // every array of the step was staged through the batch's own pinned blob, which has the device blob's layout:
    // one copy instead of five (what small per-locus steps are made of is launch and copy overhead)
    put(0, b->h_in, b->blob_bytes);
    return;
  }
  if (q.midx)
  {
    const size_t m0 = moff[i0], m1 = moff[i1];
    put(b->o_mat_idx + m0 * 4, q.midx + m0, (m1 - m0) * 4);
    put(b->o_mat_bl + m0 * 8, q.mbl + m0, (m1 - m0) * 8);
  }
  if (q.ops)
  {
    const size_t o0 = ooff[i0], o1 = ooff[i1];
    put(b->o_ops + o0 * sizeof(bppgpu_partial_op), q.ops + o0, (o1 - o0) * sizeof(bppgpu_partial_op));
  }
  if (q.rclv)
  {
    put(b->o_root_clv + (size_t)i0 * 4, q.rclv + i0, (size_t)(i1 - i0) * 4);
    put(b->o_root_sc + (size_t)i0 * 4, q.rsc + i0, (size_t)(i1 - i0) * 4);
  }
}

// every index of a step against the dimensions of its locus (what the per-locus wrappers check for one locus)
static bool locus_ops_valid(const bppgpu_locus * l, unsigned count, const bppgpu_partial_op * ops, unsigned * bad)
{
  const unsigned nb = l->tips + l->clv_buffers;
  for (unsigned i = 0; i < count; ++i)
  {
    const bppgpu_partial_op & o = ops[i];
    if (o.parent_clv_index < l->tips || o.parent_clv_index >= nb || o.left_clv_index >= nb || o.right_clv_index >= nb ||
        o.left_pmatrix_index >= l->prob_matrices || o.right_pmatrix_index >= l->prob_matrices ||
        o.parent_scaler_index >= (int)l->scale_buffers || o.left_scaler_index >= (int)l->scale_buffers ||
        o.right_scaler_index >= (int)l->scale_buffers || o.parent_scaler_index < -1 || o.left_scaler_index < -1 ||
        o.right_scaler_index < -1)
    { *bad = i; return false; }
  }
  return true;
}

// Stack slots the staged lists need when evaluated in Sethi-Ullman order (the planner's order): a tip or a child the
// list does not produce costs nothing (lookup / load), an inner node needs max(a, b) values alive, one more when both
// children need the same; one of them is the register X, the rest are stack slots.  Too few slots only send a locus
// to the slow walker, so this is a sizing hint, not a correctness input.
static int batch_su_slots(const bppgpu_batch * b, const unsigned int * ocounts, const bppgpu_partial_op * ops)
{
  std::vector<unsigned char> need;
  std::vector<unsigned int> stamp;
  int worst = 1;
  size_t o = 0;
  for (unsigned i = 0; i < b->n; ++i)
  {
    const bppgpu_locus * l = b->loci[i];
    const unsigned nb = l->tips + l->clv_buffers;
    if (need.size() < nb) { need.resize(nb); stamp.resize(nb, 0xFFFFFFFFu); }
    for (unsigned k = 0; k < ocounts[i]; ++k)
    {
      const bppgpu_partial_op & q = ops[o + k];
      if (q.parent_clv_index >= nb || q.left_clv_index >= nb || q.right_clv_index >= nb) continue;   // rejected by the validation
      const int na = stamp[q.left_clv_index] == i ? need[q.left_clv_index] : 0;
      const int nr = stamp[q.right_clv_index] == i ? need[q.right_clv_index] : 0;
      const int n = (na == 0 && nr == 0) ? 1 : (na == nr ? na + 1 : std::max(na, nr));
      need[q.parent_clv_index] = (unsigned char)std::min(n, 250); stamp[q.parent_clv_index] = i;
      worst = std::max(worst, n);
    }
    o += ocounts[i];
  }
  return std::max(worst - 1, 1);
}

static bool batch_validate(const bppgpu_batch * b, const unsigned int * mcounts, const unsigned int * midx,
                           const unsigned int * ocounts, const bppgpu_partial_op * ops, const unsigned int * rclv, const int * rsc)
{
  size_t m = 0, o = 0;
  for (unsigned i = 0; i < b->n; ++i)
  {
    const bppgpu_locus * l = b->loci[i];
    if (mcounts)
    {
      for (unsigned k = 0; k < mcounts[i]; ++k)
        if (midx[m + k] >= l->prob_matrices) { fatal("batch locus %u: pmatrix index %u out of range", i, midx[m + k]); return false; }
      m += mcounts[i];
    }
    if (ocounts)
    {
      unsigned bad = 0;
      if (!locus_ops_valid(l, ocounts[i], ops + o, &bad)) { fatal("batch locus %u: op %u has an index out of range", i, bad); return false; }
      o += ocounts[i];
    }
    if (rclv)
    {
      if (rclv[i] >= l->tips + l->clv_buffers) { fatal("batch locus %u: root clv index %u out of range", i, rclv[i]); return false; }
      if (rsc && (rsc[i] >= (int)l->scale_buffers || rsc[i] < -1)) { fatal("batch locus %u: root scaler index %d out of range", i, rsc[i]); return false; }
    }
  }
  return true;
}

// Stage the inputs of one step: build the per-locus offset tables and hand the caller's arrays to the copy
// engine.  Arrays in pinned memory (bppgpu_host_alloc / cudaHostRegister) are copied from where they are,
// anything else goes through the batch's pinned blob.  Any of the three groups may be absent (nullptr counts).
// A step that will run in waves (see bppgpu_batch_set_waves) only records the sources here: its copies are
// issued wave by wave in batch_run, interleaved with the launches, so that the first kernel starts after
// 1/waves of the upload.
// bytes in front of the header in a 20-state block: the category-major kernel's matrix images (RL x cap x 3840)
static size_t s20_hdr_shift(const bppgpu_batch * b)
{
  return b->s20_cat ? (size_t)b->RL * s20t_img_bytes(2 * b->max_tips - 1) : 0;
}

static int batch_stage(bppgpu_batch * b, const unsigned int * mcounts, const unsigned int * midx, const double * mbl,
                       const unsigned int * ocounts, const bppgpu_partial_op * ops,
                       const unsigned int * rclv, const int * rsc)
{
  bppgpu_engine * e = b->e;
  CUDA_CHECK(cudaSetDevice(e->device));
  const unsigned n = b->n;
  // unchanged counts (the usual case: same trees, new branch lengths / flipped indices): the offset tables
  // in the pinned blob and on the device are still right
  bool same = b->tables_on_device && (mcounts != nullptr) == b->have_m && (ocounts != nullptr) == b->have_o &&
              (!mcounts || memcmp(mcounts, b->last_mcounts.data(), (size_t)n * 4) == 0) &&
              (!ocounts || memcmp(ocounts, b->last_ocounts.data(), (size_t)n * 4) == 0);
  size_t tm = b->tables_tm, to = b->tables_to;
  if (!same)
  {
    tm = to = 0;
    if (mcounts) for (unsigned i = 0; i < n; ++i) tm += mcounts[i];
    if (ocounts) for (unsigned i = 0; i < n; ++i) to += ocounts[i];
    size_t off = 0;
    b->o_mat_off = off; off = align_up(off + (n + 1) * 4, 16);
    b->o_mat_idx = off; off = align_up(off + tm * 4, 16);
    b->o_mat_bl = off;  off = align_up(off + tm * 8, 16);
    b->o_op_off = off;  off = align_up(off + (n + 1) * 4, 16);
    b->o_ops = off;     off = align_up(off + to * sizeof(bppgpu_partial_op), 16);
    b->o_root_clv = off; off = align_up(off + n * 4, 16);
    b->o_root_sc = off;  off = align_up(off + n * 4, 16);
    b->o_blk_off = off;  off = align_up(off + (size_t)(n + 1) * 8, 16);
    b->blob_bytes = off;
    if (off > b->h_in_cap)
    {
      CUDA_CHECK(cudaStreamSynchronize(b->stream));
      if (b->copy_stream) CUDA_CHECK(cudaStreamSynchronize(b->copy_stream));
      if (b->h_in) cudaFreeHost(b->h_in);
      if (b->d_in) cudaFree(b->d_in);
      b->h_in_cap = b->d_in_cap = off + off / 4;
      CUDA_CHECK(cudaHostAlloc(&b->h_in, b->h_in_cap, cudaHostAllocDefault));
      CUDA_CHECK(cudaMalloc(&b->d_in, b->d_in_cap));
    }
    if (b->kernel_kind == 1 && to + n > b->plan_cap)
    {
      CUDA_CHECK(cudaStreamSynchronize(b->stream));
      if (b->d_plan) cudaFree(b->d_plan);
      b->plan_cap = (to + n) + (to + n) / 4;
      CUDA_CHECK(cudaMalloc(&b->d_plan, b->plan_cap * sizeof(PlanOp)));
    }
  }
  // Indices address a shared arena: one that is out of range would silently overwrite another locus' buffers.
  // Checked whenever the shape of the step changed (and on every stage with BPPGPU_CHECK_INDICES=1).
  static const bool check_always = getenv("BPPGPU_CHECK_INDICES") && atoi(getenv("BPPGPU_CHECK_INDICES")) != 0;
  if (!same || check_always)
    if (!batch_validate(b, mcounts, midx, ocounts, ops, rclv, rsc)) return BPPGPU_FAILURE;
  if (ocounts && ops && b->kernel_kind == 0 && b->max_tips > 16) b->su_slots = batch_su_slots(b, ocounts, ops);
  // the previous step's blob may still be in flight on the stream (a waved step joins alt_stream into it)
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  if (b->copy_stream) CUDA_CHECK(cudaStreamSynchronize(b->copy_stream));      // tables of a step that never ran
  b->inputs_pending = false;
  unsigned int * moff = (unsigned int *)(b->h_in + b->o_mat_off);
  unsigned int * ooff = (unsigned int *)(b->h_in + b->o_op_off);
  unsigned long long * boff = (unsigned long long *)(b->h_in + b->o_blk_off);
  if (!same)
  {
    unsigned int mo = 0, oo = 0;
    unsigned long long bo = 0;
    moff[0] = 0; ooff[0] = 0; boff[0] = 0;
    for (unsigned i = 0; i < n; ++i)
    {
      const unsigned oc = ocounts ? ocounts[i] : 0;
      mo += mcounts ? mcounts[i] : 0;
      oo += oc;
      // one spare op per locus for a root that the list does not produce (CTL_EVAL_ONLY)
      bo += b->kernel_kind == 2 ? align_up(s20_hdr_shift(b) + block20_bytes(b->RL, oc + 1), 256) : block_bytes(b->RL, oc + 1, s4_lut_cap_rt(b));
      moff[i + 1] = mo; ooff[i + 1] = oo; boff[i + 1] = bo;
    }
    b->have_m = mcounts != nullptr; b->have_o = ocounts != nullptr;
    if (mcounts) b->last_mcounts.assign(mcounts, mcounts + n);
    if (ocounts) b->last_ocounts.assign(ocounts, ocounts + n);
    b->tables_tm = tm; b->tables_to = to;
    if (b->kernel_kind != 1 && boff[n] > b->blocks_cap)
    {
      cudaFree(b->d_blocks_par[0]); cudaFree(b->d_blocks_par[1]);
      b->d_blocks_par[0] = b->d_blocks_par[1] = nullptr;
      b->blocks_cap = boff[n] + boff[n] / 4 + 8192;          // slack: the 20-state kernel fetches fixed-size meta blocks
      CUDA_CHECK(cudaMalloc(&b->d_blocks_par[0], b->blocks_cap));
    }
  }
  // new op lists or roots: whatever was planned is void (a matrices-only stage keeps the plan)
  if (ocounts || rclv || !same) { b->plan_valid[0] = b->plan_valid[1] = false; b->parity = 0; }
  b->d_blocks = b->d_blocks_par[b->parity];
  // waves: a full pass of a big 4-state batch uploads and runs in slices of loci (balanced by tiles)
  unsigned W = 1;
  if (b->kernel_kind == 0 && b->copy_stream && mcounts && ocounts && rclv)
  {
    if (b->wave_pref) W = b->wave_pref;
    else if (n >= 1024 && tm * 12 + to * sizeof(bppgpu_partial_op) >= (1u << 20)) W = 2;
    if (W > n) W = n;
  }
  if (W > BPPGPU_MAX_WAVES) W = BPPGPU_MAX_WAVES;
  if (W != b->n_waves || !same)
  {
    b->n_waves = W;
    b->wave_first[0] = 0;
    // geometric wave sizes (ratio r): the first wave is small so that the kernels start early, and each
    // later upload still hides under the compute of the wave before it (compute : upload is about 5 : 1).
    // Measured on config 2 (tools/run_waves.sh): 2 waves of 20 % + 80 % are as good as it gets (0.556 ms per
    // step against 0.603 ms unpipelined); every further wave costs about as much in launch gaps and kernel
    // tails as its earlier start saves.
    double ratio = 4.0;
    if (const char * ev = getenv("BPPGPU_WAVE_RATIO")) ratio = std::max(1.0, atof(ev));
    double total_w = 0, acc = 0, cur = 1;
    for (unsigned w = 0; w < W; ++w) { total_w += cur; cur *= ratio; }
    cur = 1;
    for (unsigned w = 1, i = 0; w <= W; ++w)
    {
      acc += cur; cur *= ratio;
      const unsigned long long goal = (unsigned long long)((double)b->n_tiles * (acc / total_w));
      while (i < n && b->h_tile_first[i] < goal) ++i;
      b->wave_first[w] = (w == W) ? n : i;
    }
  }
  cudaStream_t cs = W > 1 ? b->copy_stream : b->stream;
  // sources: the caller's pinned arrays, or their copy in the pinned blob
  auto is_pinned = [](const void * p) -> bool
  {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
  };
  b->all_in_blob = true;
  auto source = [&](size_t dst, const void * src, size_t bytes) -> const void *
  {
    if (!src || !bytes) return nullptr;
    // small arrays are cheaper to memcpy into the blob than to ask the driver what kind of memory they are
    if (bytes > 4096 && is_pinned(src)) { b->all_in_blob = false; return src; }
    memcpy(b->h_in + dst, src, bytes);
    return b->h_in + dst;
  };
  bppgpu_batch::Pending & q = b->pend;
  q.midx = (const unsigned int *)source(b->o_mat_idx, mcounts ? midx : nullptr, tm * 4);
  q.mbl = (const double *)source(b->o_mat_bl, mcounts ? mbl : nullptr, tm * 8);
  q.ops = (const bppgpu_partial_op *)source(b->o_ops, ocounts ? ops : nullptr, to * sizeof(bppgpu_partial_op));
  q.rclv = (const unsigned int *)source(b->o_root_clv, rclv, (size_t)n * 4);
  if (rclv && !rsc)
  {
    int * p = (int *)(b->h_in + b->o_root_sc);
    for (unsigned i = 0; i < n; ++i) p[i] = -1;
    q.rsc = p;
  }
  else q.rsc = (const int *)source(b->o_root_sc, rclv ? rsc : nullptr, (size_t)n * 4);
  if (!same && !(b->all_in_blob && W == 1))       // (a step that lives entirely in the blob travels as one copy, tables included)
  {
    CUDA_CHECK(cudaMemcpyAsync(b->d_in + b->o_mat_off, b->h_in + b->o_mat_off, (n + 1) * 4, cudaMemcpyHostToDevice, cs));
    CUDA_CHECK(cudaMemcpyAsync(b->d_in + b->o_op_off, b->h_in + b->o_op_off, (n + 1) * 4, cudaMemcpyHostToDevice, cs));
    CUDA_CHECK(cudaMemcpyAsync(b->d_in + b->o_blk_off, b->h_in + b->o_blk_off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, cs));
  }
  b->tables_on_device = true;
  if (W > 1)
  {
    // copies are issued by batch_run, wave by wave; the tables (if any) are already on their way
    CUDA_CHECK(cudaEventRecord(b->ev_tables, cs));
    b->inputs_pending = true;
  }
  else { batch_issue_copies(b, 0, n, cs); CUDA_CHECK(cudaEventRecord(b->ev_inputs, cs)); }
  b->total_mats = (unsigned)tm; b->total_ops = (unsigned)to;
  b->staged_mats = mcounts != nullptr; b->staged_ops = ocounts != nullptr; b->staged_roots = rclv != nullptr;
  return BPPGPU_SUCCESS;
}

extern "C" void * bppgpu_host_alloc(size_t bytes)
{
  void * p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess)
  {
    cudaGetLastError();
    fatal("Unable to allocate %zu bytes of pinned host memory", bytes);
    return nullptr;
  }
  return p;
}

extern "C" void bppgpu_host_free(void * p) { if (p) cudaFreeHost(p); }

static void batch_sync_loci(bppgpu_batch * b, bool need_eigen)
{
  // nothing of any locus changed since this batch last looked (and it then decomposed what it needed)
  if (b->synced_epoch == b->e->dirty_epoch.load() && (!need_eigen || b->synced_eigen)) return;
  bppgpu_engine * e = b->e;
  // loci whose model block must travel, and those that need pll_update_eigen first (locus.c:2462-2476)
  std::vector<bppgpu_locus *> dirty;
  for (bppgpu_locus * l : b->loci)
  {
    if (l->codes_dirty || l->flags_dirty)
    {
      const bool md = l->model_dirty;
      l->model_dirty = false;
      locus_sync(l, b->stream);
      l->model_dirty = md;
    }
    if (l->model_dirty || (need_eigen && !l->eigen_valid && !model_is_closed_form(l))) dirty.push_back(l);
  }
  size_t min_dirty = 8;
  if (const char * ev = getenv("BPPGPU_MODEL_BATCH_MIN")) min_dirty = (size_t)std::max(1, atoi(ev));   // tuning / test knob
  if (dirty.size() < min_dirty)
  {
    for (bppgpu_locus * l : dirty)
    {
      if (need_eigen && !l->eigen_valid && !model_is_closed_form(l)) { host_update_eigen(l); l->eigen_on_device = false; }
      if (l->model_dirty) locus_sync(l, b->stream);
    }
  }
  else
  {
    // one pinned blob: [ids u32][eigen modes u8][record offsets u64][scratch offsets u64][records f64]
    const size_t nd = dirty.size();
    size_t rec_doubles = 0, scr_doubles = 0;
    std::vector<unsigned long long> roff(nd), soff(nd);
    std::vector<unsigned char> mode(nd);
    for (size_t d = 0; d < nd; ++d)
    {
      bppgpu_locus * l = dirty[d];
      const size_t S = l->states, R = l->rate_cats;
      const bool compute = need_eigen && !l->eigen_valid && !model_is_closed_form(l);
      // a valid decomposition that lives only on the device stays where it is
      mode[d] = compute ? EIGEN_COMPUTE : ((l->eigen_valid && !l->eigen_on_device) ? EIGEN_COPY : EIGEN_KEEP);
      roff[d] = rec_doubles; soff[d] = scr_doubles;
      rec_doubles += S + 2 * R + S * (S - 1) / 2 + (mode[d] == EIGEN_COPY ? 2 * S * S + S : 0);
      if (compute) scr_doubles += 2 * S * S;
    }
    const size_t o_ids = 0, o_mode = align_up(o_ids + nd * 4, 16), o_roff = align_up(o_mode + nd, 16),
                 o_soff = o_roff + nd * 8, o_rec = o_soff + nd * 8, total = o_rec + rec_doubles * 8;
    if (total > b->model_cap)
    {
      CUDA_CHECK(cudaStreamSynchronize(b->stream));
      if (b->h_model) cudaFreeHost(b->h_model);
      if (b->d_model) cudaFree(b->d_model);
      b->model_cap = total + total / 4;
      CUDA_CHECK(cudaHostAlloc(&b->h_model, b->model_cap, cudaHostAllocDefault));
      CUDA_CHECK(cudaMalloc(&b->d_model, b->model_cap));
    }
    else CUDA_CHECK(cudaStreamSynchronize(b->stream));       // the blob of the previous sync may still be in flight
    if (scr_doubles * 8 > b->eig_scratch_cap)
    {
      if (b->d_eig_scratch) cudaFree(b->d_eig_scratch);
      b->eig_scratch_cap = scr_doubles * 8 + scr_doubles * 2;
      CUDA_CHECK(cudaMalloc(&b->d_eig_scratch, b->eig_scratch_cap));
    }
    unsigned int * ids = (unsigned int *)(b->h_model + o_ids);
    memcpy(b->h_model + o_mode, mode.data(), nd);
    memcpy(b->h_model + o_roff, roff.data(), nd * 8);
    memcpy(b->h_model + o_soff, soff.data(), nd * 8);
    double * recs = (double *)(b->h_model + o_rec);
    for (size_t d = 0; d < nd; ++d)
    {
      bppgpu_locus * l = dirty[d];
      const size_t S = l->states, R = l->rate_cats, np = S * (S - 1) / 2;
      ids[d] = l->id;
      double * w = recs + roff[d];
      memcpy(w, l->h_freqs.data(), S * 8); w += S;
      memcpy(w, l->h_rates.data(), R * 8); w += R;
      memcpy(w, l->h_rate_weights.data(), R * 8); w += R;
      memcpy(w, l->h_subst.data(), np * 8); w += np;
      if (mode[d] == EIGEN_COPY)
      {
        memcpy(w, l->h_evecs.data(), S * S * 8); w += S * S;
        memcpy(w, l->h_ievecs.data(), S * S * 8); w += S * S;
        memcpy(w, l->h_evals.data(), S * 8);
      }
      else if (mode[d] == EIGEN_COMPUTE) { l->eigen_valid = true; l->eigen_on_device = true; }
      l->model_dirty = false;
    }
    CUDA_CHECK(cudaMemcpyAsync(b->d_model, b->h_model, total, cudaMemcpyHostToDevice, b->stream));
    e->launches++;
    model_update_kernel<<<(unsigned)nd, 64, 0, b->stream>>>(
        e->d_loci, (const unsigned int *)(b->d_model + o_ids), (const double *)(b->d_model + o_rec),
        (const unsigned long long *)(b->d_model + o_roff), (const unsigned char *)(b->d_model + o_mode),
        b->d_eig_scratch, (const unsigned long long *)(b->d_model + o_soff));
    CUDA_CHECK(cudaGetLastError());
  }
  // LocusHdr carries frequencies and flags of the loci, the blocks their rate weights: replan
  if (b->synced_epoch != e->dirty_epoch.load()) b->plan_valid[0] = b->plan_valid[1] = false;
  b->synced_epoch = e->dirty_epoch.load();
  b->synced_eigen = need_eigen;
}

// persistent launch: as many CTAs as fit on the device at once, each walks a contiguous tile range.
// The stack-slot count is a performance knob only (a value that finds no slot is re-read from L2), so it
// is lowered until two CTAs fit on an SM.
template <int RL, int CPT>
static int tree_s4_slots(bppgpu_batch * b, int wanted, unsigned cap)
{
  // A list whose parked values do not all get a slot leaves the fast path (its loci run the cell-at-a-time walker,
  // 3-4x slower), so the slots a tree of this size can need come first: if they fit one CTA they are kept and the
  // occupancy gives way (48 tips with 4 categories used to end up with one slot at two CTAs per SM -- which the
  // shared memory did not allow anyway).  Only what does not even fit one CTA is cut.
  int slots = wanted;
  while (slots > 1 && S4Layout<RL, CPT>::bytes(slots, cap, b->tip_words_rt) + 1024 > b->e->smem_optin) --slots;
  return slots;
}

template <int RL, bool EXACT, int CPT, bool SCALED_ONLY>
static void launch_tree_s4_impl(bppgpu_batch * b, const TreeParams & prm, cudaStream_t st)
{
  bppgpu_engine * e = b->e;
  const size_t smem = S4Layout<RL, CPT>::bytes(prm.n_slots, prm.lut_cap, prm.tip_words);
  b->slots = prm.n_slots; b->tree_smem = smem;
  constexpr int K = SCALED_ONLY ? 1 : 0;
  const unsigned key = (unsigned)(K * 1000 + RL * 100 + CPT * 10 + (EXACT ? 1 : 0));
  if (b->cfg_key[K] != key || b->cfg_smem[K] != smem)
  {
    ensure_max_smem(e, tree_kernel_s4<RL, EXACT, CPT, SCALED_ONLY>);
    int per_sm = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tree_kernel_s4<RL, EXACT, CPT, SCALED_ONLY>, TREE_NT, smem));
    if (per_sm < 1) { fatal("tree kernel does not fit on an SM (smem %zu)", smem); return; }
    b->cfg_key[K] = key; b->cfg_smem[K] = smem; b->cfg_per_sm[K] = per_sm;
  }
  const unsigned grid = std::min<unsigned>(prm.n_tiles, (unsigned)(b->cfg_per_sm[K] * e->sm_count));
  tree_kernel_s4<RL, EXACT, CPT, SCALED_ONLY><<<grid, TREE_NT, smem, st>>>(prm);
}

template <int RL, bool EXACT, int CPT>
static void launch_tree_s4_kind(bppgpu_batch * b, const TreeParams & prm, cudaStream_t st, bool scaled_only)
{
  if (scaled_only) launch_tree_s4_impl<RL, EXACT, CPT, true>(b, prm, st);
  else launch_tree_s4_impl<RL, EXACT, CPT, false>(b, prm, st);
}

template <int RL>
static void launch_tree_s4_rl(bppgpu_batch * b, const TreeParams & prm, cudaStream_t st, bool scaled_only)
{
  const bool exact = b->e->math == BPPGPU_MATH_EXACT;
  if (b->cpt == 2) { if (exact) launch_tree_s4_kind<RL, true, 2>(b, prm, st, scaled_only); else launch_tree_s4_kind<RL, false, 2>(b, prm, st, scaled_only); }
  else if (b->cpt == 4) { if (exact) launch_tree_s4_kind<RL, true, 4>(b, prm, st, scaled_only); else launch_tree_s4_kind<RL, false, 4>(b, prm, st, scaled_only); }
  else { if (exact) launch_tree_s4_kind<RL, true, 1>(b, prm, st, scaled_only); else launch_tree_s4_kind<RL, false, 1>(b, prm, st, scaled_only); }
}

// scaled_only: the specialised launch for a cached plan of scaled one-chunk loci (otherwise the kernel with every path)
static int launch_tree_s4(bppgpu_batch * b, const TreeParams & prm, cudaStream_t st, bool scaled_only = false)
{
  b->last_scaled_only = scaled_only;
  switch (b->RL)
  {
    case 1: launch_tree_s4_rl<1>(b, prm, st, scaled_only); break;
    case 2: launch_tree_s4_rl<2>(b, prm, st, scaled_only); break;
    case 4: launch_tree_s4_rl<4>(b, prm, st, scaled_only); break;
    case 8: launch_tree_s4_rl<8>(b, prm, st, scaled_only); break;
    default: fatal("internal: RL=%u", b->RL); return BPPGPU_FAILURE;
  }
  return BPPGPU_SUCCESS;
}

template <int RL>
static void launch_tree_s20(bppgpu_batch * b, const TreeParams & prm)
{
  bppgpu_engine * e = b->e;
  const size_t smem = s20_smem_bytes<RL>(prm.n_slots);
  ensure_max_smem(e, tree_kernel_s20<RL>);
  int per_sm = 0;
  CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tree_kernel_s20<RL>, S20_NT, smem));
  if (per_sm < 1) { fatal("20-state tree kernel does not fit on an SM (smem %zu)", smem); return; }
  const unsigned grid = std::min<unsigned>(b->n_tiles, (unsigned)(per_sm * e->sm_count));
  tree_kernel_s20<RL><<<grid, S20_NT, smem, b->stream>>>(prm);
}

// category-major 20-state kernel: persistent CTAs, one per SM; scaled batches as clusters of RL CTAs (rank = category)
template <int RL, bool SCALED>
static int launch_tree_s20t_inst(bppgpu_batch * b, const TreeParams & prm)
{
  bppgpu_engine * e = b->e;
  const size_t smem = s20t_smem_bytes(prm.lut_cap, prm.max_tips, prm.n_slots, SCALED);
  ensure_max_smem(e, tree_kernel_s20t<RL, SCALED>);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(S20T_NT); cfg.dynamicSmemBytes = smem; cfg.stream = b->stream;
  cudaLaunchAttribute attr[1];
  unsigned grid = std::min<unsigned>(b->n_tiles, (unsigned)e->sm_count);
  if (SCALED && RL > 1)
  {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = RL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3(RL * (unsigned)(e->sm_count / RL));
    int max_clusters = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveClusters(&max_clusters, tree_kernel_s20t<RL, SCALED>, &cfg));
    if (max_clusters < 1) { fatal("20-state scaled tree kernel: no cluster of %d CTAs fits (smem %zu)", RL, smem); return BPPGPU_FAILURE; }
    grid = RL * std::min<unsigned>(b->n_tiles, (unsigned)max_clusters);
  }
  cfg.gridDim = dim3(grid);
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, tree_kernel_s20t<RL, SCALED>, prm, b->d_rootsc));
  return BPPGPU_SUCCESS;
}

static int launch_tree_s20t(bppgpu_batch * b, const TreeParams & prm)
{
  switch (b->RL * 2 + (b->s20_scaled ? 1 : 0))
  {
    case 2: return launch_tree_s20t_inst<1, false>(b, prm);
    case 3: return launch_tree_s20t_inst<1, true>(b, prm);
    case 4: return launch_tree_s20t_inst<2, false>(b, prm);
    case 5: return launch_tree_s20t_inst<2, true>(b, prm);
    case 8: return launch_tree_s20t_inst<4, false>(b, prm);
    case 9: return launch_tree_s20t_inst<4, true>(b, prm);
    case 16: return launch_tree_s20t_inst<8, false>(b, prm);
    case 17: return launch_tree_s20t_inst<8, true>(b, prm);
    default: fatal("internal: RL=%u", b->RL); return BPPGPU_FAILURE;
  }
}

// launches: [pmatrix] [plan + tree (+ finish)] on the batch stream, using the staged blob
static int batch_run(bppgpu_batch * b, bool do_mats, bool do_tree, bool want_root, double * persite, int persite_mode)
{
  bppgpu_engine * e = b->e;
  CUDA_CHECK(cudaSetDevice(e->device));
  const unsigned n = b->n;
  batch_sync_loci(b, do_mats);
  const unsigned int * d_mat_off = (const unsigned int *)(b->d_in + b->o_mat_off);
  const unsigned int * d_mat_idx = (const unsigned int *)(b->d_in + b->o_mat_idx);
  const double * d_mat_bl = (const double *)(b->d_in + b->o_mat_bl);
  const unsigned int * d_op_off = (const unsigned int *)(b->d_in + b->o_op_off);
  const RawOp * d_ops = (const RawOp *)(b->d_in + b->o_ops);
  const unsigned int * d_root_clv = (const unsigned int *)(b->d_in + b->o_root_clv);
  const int * d_root_sc = (const int *)(b->d_in + b->o_root_sc);

  // 4-state batches doing matrices and partials in one call build the P-matrices inside the planner
  const bool fuse_mats = do_mats && do_tree && b->kernel_kind == 0 && b->total_mats;
  if (do_mats && b->total_mats && !fuse_mats)
  {
    ProfScope ps(e, b->stream, BPPGPU_KERNEL_PMATRIX);
    const unsigned S = b->loci[0]->states;
    if (S > 8)
    {
      const size_t sm = (2 * (size_t)S * S + 4 * ((size_t)S * (S + 1) + S)) * 8;
      // BPPGPU_PMAT_DMMA=0 keeps the reference's separate multiply / add order (pmatrix_kernel_wide)
      static const bool use_dmma = !(getenv("BPPGPU_PMAT_DMMA") && atoi(getenv("BPPGPU_PMAT_DMMA")) == 0);
      if (S == 20 && use_dmma) pmatrix_kernel_dmma20<<<dim3(n, 4), 128, 0, b->stream>>>(e->d_loci, b->d_batch_locus, d_mat_off, d_mat_idx, d_mat_bl);
      else if (S == 20) pmatrix_kernel_wide<20><<<dim3(n, 8), 128, sm, b->stream>>>(e->d_loci, b->d_batch_locus, d_mat_off, d_mat_idx, d_mat_bl);
      else pmatrix_kernel_wide<0><<<dim3(n, 8), 128, sm, b->stream>>>(e->d_loci, b->d_batch_locus, d_mat_off, d_mat_idx, d_mat_bl);
    }
    else
      pmatrix_kernel<<<n, 64, 0, b->stream>>>(e->d_loci, b->d_batch_locus, d_mat_off, d_mat_idx, d_mat_bl);
    CUDA_CHECK(cudaGetLastError());
  }
  if (!do_tree) return BPPGPU_SUCCESS;

  // shared-memory stack slots per thread: ceil(log2 T) covers balanced trees in recursive post-order;
  // beyond that the plan falls back to re-reading the child from HBM (still correct)
  int slots = 0;
  unsigned lut_cap_rt = 0;
  if (b->kernel_kind == 2 && b->overflow_epoch != e->dirty_epoch.load())      // (tip states changed since the last look)
  {
    for (auto * l : b->loci)
      if (l->col_overflow)
      {
        fatal("a locus of this batch has more than 4 distinct ambiguity codes; create the batch after the tip states are set");
        return BPPGPU_FAILURE;
      }
    b->overflow_epoch = e->dirty_epoch.load();
  }
  if (b->kernel_kind != 1)
  {
    const unsigned maxT = b->max_tips;
    slots = s4_slots_formula(maxT);
    // big 4-state trees: what the staged lists really need (Sethi-Ullman, batch_validate) -- every slot costs 8-36 kB
    if (b->kernel_kind == 0 && b->su_slots > 0) slots = std::min(slots, b->su_slots);
    // 20 states: a parked X costs 3.4 kB of shared memory per warp and slot, which would push the
    // P-matrices (read through L1) out of the SM; re-reading the child's CLV (an L2 hit) and redoing one
    // DMMA mat-vec is cheaper, so nothing is parked
    if (b->kernel_kind == 2 && !b->s20_cat) slots = 0;
    if (b->s20_cat)
    {
      // one CTA of 16 warps per SM, two stage buffers (matrix images + meta block); the stack takes what is left
      lut_cap_rt = 2 * maxT - 1;
      while (slots > 0 && s20t_smem_bytes(lut_cap_rt, maxT, slots, b->s20_scaled) + 1024 > e->smem_optin) --slots;
    }
    if (b->kernel_kind == 0)
    {
      // tip-slot capacity: any op list over a T-tip tree has at most T tip or HBM-resident children
      lut_cap_rt = s4_lut_cap_rt(b);
      const int w = slots;
      switch (b->RL * 10 + b->cpt)
      {
        case 11: slots = tree_s4_slots<1, 1>(b, w, lut_cap_rt); break;
        case 12: slots = tree_s4_slots<1, 2>(b, w, lut_cap_rt); break;
        case 14: slots = tree_s4_slots<1, 4>(b, w, lut_cap_rt); break;
        case 24: slots = tree_s4_slots<2, 4>(b, w, lut_cap_rt); break;
        case 44: slots = tree_s4_slots<4, 4>(b, w, lut_cap_rt); break;
        case 84: slots = tree_s4_slots<8, 4>(b, w, lut_cap_rt); break;
        case 21: slots = tree_s4_slots<2, 1>(b, w, lut_cap_rt); break;
        case 22: slots = tree_s4_slots<2, 2>(b, w, lut_cap_rt); break;
        case 41: slots = tree_s4_slots<4, 1>(b, w, lut_cap_rt); break;
        case 42: slots = tree_s4_slots<4, 2>(b, w, lut_cap_rt); break;
        case 81: slots = tree_s4_slots<8, 1>(b, w, lut_cap_rt); break;
        case 82: slots = tree_s4_slots<8, 2>(b, w, lut_cap_rt); break;
        default: break;
      }
    }
  }
  const unsigned long long * d_blk_off = (const unsigned long long *)(b->d_in + b->o_blk_off);
  // inputs staged in waves and not consumed yet: either run wave by wave (full pass of the 4-state kernel),
  // or make the stream wait for the whole upload
  const bool waved = b->inputs_pending && b->n_waves > 1 && b->kernel_kind == 0 && fuse_mats && want_root && !persite;
  if (b->inputs_pending && !waved)
  {
    CUDA_CHECK(cudaStreamWaitEvent(b->stream, b->ev_tables, 0));
    batch_issue_copies(b, 0, n, b->stream);
    CUDA_CHECK(cudaEventRecord(b->ev_inputs, b->stream));
  }
  b->inputs_pending = false;
  // the planned program depends on the op lists (and on want_root / slots / table capacity) only: if the blocks of
  // this index parity are still those of the staged lists, rebuild the matrices and copy them into the blocks again
  const unsigned int pkey = 1u | (want_root ? 2u : 0u) | ((unsigned)slots << 2) | (lut_cap_rt << 8);
  static const bool plan_cache_on = !(getenv("BPPGPU_PLAN_CACHE") && atoi(getenv("BPPGPU_PLAN_CACHE")) == 0);
  // (20 states, category-major kernel: the block holds no matrices -- image20_kernel rebuilds the images from the
  // pmatrix block every run -- so a cached plan needs no refresh at all)
  const bool cacheable = b->kernel_kind == 0 || (b->kernel_kind == 2 && b->s20_cat);
  const bool cached = plan_cache_on && cacheable && b->plan_valid[b->parity] && b->plan_key[b->parity] == pkey && !persite;
  if (!waved && cached && b->kernel_kind == 0)
  {
    ProfScope ps(e, b->stream, BPPGPU_KERNEL_PLAN);
    // a CTA per locus for batches of few, big trees; a warp per locus otherwise
    if (n <= 4096 && b->max_tips > 16)
      plan_refresh_blocks<4><<<n, 128, 0, b->stream>>>(
          e->d_loci, b->d_batch_locus, n, b->d_blocks, d_blk_off, b->RL, fuse_mats ? d_mat_off : nullptr, d_mat_idx, d_mat_bl);
    else
      plan_refresh_blocks<1><<<(n * 32 + 127) / 128, 128, 0, b->stream>>>(
          e->d_loci, b->d_batch_locus, n, b->d_blocks, d_blk_off, b->RL, fuse_mats ? d_mat_off : nullptr, d_mat_idx, d_mat_bl);
    CUDA_CHECK(cudaGetLastError());
  }
  if (cacheable && !cached) { b->plan_valid[b->parity] = !persite; b->plan_key[b->parity] = pkey; b->class_pending[b->parity] = false; b->plan_class[b->parity] = 0; }
  // BPPGPU_S4_SCALED_ONLY=0: always the kernel with every path
  static const bool s4_kinds_on = !(getenv("BPPGPU_S4_SCALED_ONLY") && atoi(getenv("BPPGPU_S4_SCALED_ONLY")) == 0);
  if (!waved && !cached)
  {
    ProfScope ps(e, b->stream, BPPGPU_KERNEL_PLAN);
    if (b->kernel_kind == 0)
      plan_kernel_blocks<<<(n * 32 + 127) / 128, 128, 0, b->stream>>>(
          e->d_loci, b->d_batch_locus, n, d_op_off, d_ops, d_root_clv, d_root_sc, want_root ? 1 : 0,
          b->d_blocks, d_blk_off, b->d_tile_first, b->d_tile_blk, b->d_plan_count, b->d_scratch, b->d_scratch_off,
          slots, b->RL, b->cpt, lut_cap_rt, fuse_mats ? d_mat_off : nullptr, d_mat_idx, d_mat_bl, 8u * b->tip_words_rt);
    else if (b->kernel_kind == 2)
      plan_kernel_blocks20<<<(n * 32 + 127) / 128, 128, 0, b->stream>>>(
          e->d_loci, b->d_batch_locus, n, d_op_off, d_ops, d_root_clv, d_root_sc, want_root ? 1 : 0,
          b->d_blocks, d_blk_off, b->d_tile_first, b->d_tile_blk, b->d_plan_count, b->d_scratch, b->d_scratch_off,
          slots, b->RL, (unsigned long long)s20_hdr_shift(b), b->s20_cat ? 0 : 1);
    else
      plan_kernel_flat<<<(n + 127) / 128, 128, 0, b->stream>>>(
          e->d_loci, b->d_batch_locus, n, d_op_off, d_ops, d_root_clv, d_root_sc, want_root ? 1 : 0,
          b->d_plan, b->d_plan_count);
    CUDA_CHECK(cudaGetLastError());
  }
  // Class of a plan that IS being reused (first run on the cached blocks): one small kernel over the block headers,
  // read back behind it; known to the host once the caller has collected that step, and from then on the runs on
  // this plan launch the specialised kernel.  Steps that are planned afresh every time (partial updates) never pay.
  bool s4_scaled_only = false;
  if (b->kernel_kind == 0 && cached && !waved && s4_kinds_on)
  {
    const int par = b->parity;
    if (b->plan_class[par] == 0 && !b->class_pending[par])
    {
      e->launches++;
      CUDA_CHECK(cudaMemsetAsync(b->d_class + par, (int)(PLAN_CLASS_KNOWN | PLAN_CLASS_LEAN | PLAN_CLASS_SCALED), 4, b->stream));   // every byte 0x07
      plan_class_kernel<<<(n + 255) / 256, 256, 0, b->stream>>>(b->d_blocks, d_blk_off, n, b->d_class + par);
      CUDA_CHECK(cudaGetLastError());
      CUDA_CHECK(cudaMemcpyAsync(b->h_class + par, b->d_class + par, 4, cudaMemcpyDeviceToHost, b->stream));
      CUDA_CHECK(cudaEventRecord(b->ev_class[par], b->stream));
      b->class_pending[par] = true;
    }
    else if (b->class_pending[par])
    {
      const cudaError_t q = cudaEventQuery(b->ev_class[par]);
      if (q == cudaSuccess)
      {
        b->plan_class[par] = b->h_class[par] & (PLAN_CLASS_KNOWN | PLAN_CLASS_LEAN | PLAN_CLASS_SCALED);
        b->class_pending[par] = false;
      }
      else if (q == cudaErrorNotReady) (void)cudaGetLastError();      // "not ready" is an answer, not an error to keep
      else CUDA_CHECK(q);
    }
    // all loci scaled-capable one-chunk lists, and not all of them lean (an unscaled batch is both)
    if (!b->class_pending[par])
      s4_scaled_only = (b->plan_class[par] & PLAN_CLASS_SCALED) && !(b->plan_class[par] & PLAN_CLASS_LEAN);
  }
  if (b->kernel_kind == 2 && b->s20_cat)
  {
    // matrix images of every (locus, category, list entry) in front of the locus' block: what the tree kernel's
    // TMA fetches bring into shared memory
    ProfScope ps(e, b->stream, BPPGPU_KERNEL_PLAN);
    image20_kernel<<<dim3(n, 4), 256, 0, b->stream>>>(e->d_loci, b->d_batch_locus, b->d_blocks, d_blk_off,
                                                       (unsigned long long)s20_hdr_shift(b), b->RL, lut_cap_rt);
    CUDA_CHECK(cudaGetLastError());
  }
  TreeParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.loci = e->d_loci; prm.batch_locus = b->d_batch_locus; prm.tile_locus = b->d_tile_locus;
  prm.tile_cell0 = b->d_tile_cell0; prm.op_off = d_op_off; prm.plan = b->d_plan; prm.plan_count = b->d_plan_count;
  prm.tiles = b->d_tiles; prm.blocks = b->d_blocks; prm.tile_blk = b->d_tile_blk; prm.n_tiles = b->n_tiles;
  prm.tile_partial = want_root ? b->d_tile_partial : nullptr;
  prm.persite = persite; prm.persite_mode = persite_mode; prm.n_slots = slots; prm.lut_cap = lut_cap_rt;
  prm.tip_words = b->tip_words_rt;
  prm.log_threshold = e->log_threshold;
  if (waved)
  {
    // wave w: [upload on copy_stream] -> planner -> tree kernel on stream (even w) / alt_stream (odd w); the
    // two compute streams let wave w+1 start on the SMs that wave w's tail frees
    CUDA_CHECK(cudaEventRecord(b->ev_fork, b->stream));
    CUDA_CHECK(cudaStreamWaitEvent(b->alt_stream, b->ev_fork, 0));
    for (unsigned w = 0; w < b->n_waves; ++w)
    {
      const unsigned i0 = b->wave_first[w], i1 = b->wave_first[w + 1];
      if (i1 <= i0) continue;
      cudaStream_t st = (w & 1u) ? b->alt_stream : b->stream;
      batch_issue_copies(b, i0, i1, b->copy_stream);
      CUDA_CHECK(cudaEventRecord(b->ev_copy[w], b->copy_stream));
      if (w + 1 == b->n_waves) CUDA_CHECK(cudaEventRecord(b->ev_inputs, b->copy_stream));
      CUDA_CHECK(cudaStreamWaitEvent(st, b->ev_copy[w], 0));
      {
        ProfScope ps(e, st, BPPGPU_KERNEL_PLAN);
        plan_kernel_blocks<<<((i1 - i0) * 32 + 127) / 128, 128, 0, st>>>(
            e->d_loci, b->d_batch_locus + i0, i1 - i0, d_op_off + i0, d_ops, d_root_clv + i0, d_root_sc + i0, 1,
            b->d_blocks, d_blk_off + i0, b->d_tile_first + i0, b->d_tile_blk, b->d_plan_count + i0, b->d_scratch,
            b->d_scratch_off + i0, slots, b->RL, b->cpt, lut_cap_rt, d_mat_off + i0, d_mat_idx, d_mat_bl, 8u * b->tip_words_rt);
        CUDA_CHECK(cudaGetLastError());
      }
      const unsigned t0 = b->h_tile_first[i0], t1 = b->h_tile_first[i1];
      TreeParams pw = prm;
      pw.tiles = b->d_tiles + t0; pw.tile_blk = b->d_tile_blk + 2 * (size_t)t0; pw.tile_partial = b->d_tile_partial + t0;
      pw.n_tiles = t1 - t0;
      {
        ProfScope ps(e, st, BPPGPU_KERNEL_TREE);
        if (!launch_tree_s4(b, pw, st)) return BPPGPU_FAILURE;
        CUDA_CHECK(cudaGetLastError());
      }
    }
    CUDA_CHECK(cudaEventRecord(b->ev_join, b->alt_stream));
    CUDA_CHECK(cudaStreamWaitEvent(b->stream, b->ev_join, 0));
  }
  else
  {
    ProfScope ps(e, b->stream, BPPGPU_KERNEL_TREE);
    if (b->kernel_kind == 0)
    {
      if (!launch_tree_s4(b, prm, b->stream, s4_scaled_only)) return BPPGPU_FAILURE;
    }
    else if (b->kernel_kind == 2)
    {
      if (b->s20_cat)
      {
        const unsigned maxT = b->max_tips;
        prm.max_tips = maxT; prm.rootdot = b->d_rootdot; prm.site_off = b->d_site_off;
        if (!launch_tree_s20t(b, prm)) return BPPGPU_FAILURE;
        if (want_root)
          root20_kernel<<<n, 128, 0, b->stream>>>(e->d_loci, b->d_batch_locus, b->d_rootdot, b->d_site_off, b->d_rootsc,
                                                  e->log_threshold, b->d_tile_first, b->d_tile_partial, persite, persite_mode);
      }
      else
      switch (b->RL)
      {
        case 1: launch_tree_s20<1>(b, prm); break;
        case 2: launch_tree_s20<2>(b, prm); break;
        case 4: launch_tree_s20<4>(b, prm); break;
        case 8: launch_tree_s20<8>(b, prm); break;
        default: fatal("internal: RL=%u", b->RL); return BPPGPU_FAILURE;
      }
    }
    else
    {
      if (e->math == BPPGPU_MATH_EXACT) tree_kernel_generic<true><<<b->n_tiles, b->tile_threads, 0, b->stream>>>(prm);
      else tree_kernel_generic<false><<<b->n_tiles, b->tile_threads, 0, b->stream>>>(prm);
    }
    CUDA_CHECK(cudaGetLastError());
  }
  if (want_root && !persite)
  {
    // loci with a diploid mapping: their root is the phase-resolution mean (locus.c:2586-2615)
    if (b->diploid_epoch != e->dirty_epoch.load())
    {
      b->any_diploid = false;
      for (auto * l : b->loci) b->any_diploid = b->any_diploid || l->dev.dip_off != nullptr;
      b->diploid_epoch = e->dirty_epoch.load();
    }
    if (b->any_diploid)
    {
      ProfScope ps(e, b->stream, BPPGPU_KERNEL_FINISH);
      diploid_batch_kernel<<<n, 256, 0, b->stream>>>(e->d_loci, b->d_batch_locus, d_root_clv, b->d_tile_first, b->d_tile_partial);
      CUDA_CHECK(cudaGetLastError());
    }
  }
  if (want_root)
  {
    ProfScope ps(e, b->stream, BPPGPU_KERNEL_FINISH);
    finish_kernel<<<(n + 255) / 256, 256, 0, b->stream>>>(b->d_tile_partial, b->d_tile_first, n, b->d_lnl, b->d_lnl_sum,
                                                          b->d_block_sums, b->d_counter);
    CUDA_CHECK(cudaGetLastError());
  }
  return BPPGPU_SUCCESS;
}

static int batch_collect(bppgpu_batch * b, double * lnl_out, double * sum_out)
{
  CUDA_CHECK(cudaSetDevice(b->e->device));
  CUDA_CHECK(cudaMemcpyAsync(b->h_out, b->d_lnl, (b->n + 1) * 8, cudaMemcpyDeviceToHost, b->stream));
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  if (lnl_out) memcpy(lnl_out, b->h_out, b->n * 8);
  if (sum_out) *sum_out = b->h_out[b->n];
  return BPPGPU_SUCCESS;
}

extern "C" int bppgpu_batch_update_matrices(bppgpu_batch * b, const unsigned int * counts,
                                            const unsigned int * idx, const double * bl)
{
  if (!batch_stage(b, counts, idx, bl, nullptr, nullptr, nullptr, nullptr)) return BPPGPU_FAILURE;
  if (!batch_run(b, true, false, false, nullptr, 0)) return BPPGPU_FAILURE;
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  return BPPGPU_SUCCESS;
}

extern "C" int bppgpu_batch_update_partials(bppgpu_batch * b, const unsigned int * counts, const bppgpu_partial_op * ops)
{
  if (!batch_stage(b, nullptr, nullptr, nullptr, counts, ops, nullptr, nullptr)) return BPPGPU_FAILURE;
  if (!batch_run(b, false, true, false, nullptr, 0)) return BPPGPU_FAILURE;
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  return BPPGPU_SUCCESS;
}

extern "C" int bppgpu_batch_root_loglikelihood(bppgpu_batch * b, const unsigned int * rclv, const int * rsc, double * out)
{
  if (!batch_stage(b, nullptr, nullptr, nullptr, nullptr, nullptr, rclv, rsc)) return BPPGPU_FAILURE;
  if (!batch_run(b, false, true, true, nullptr, 0)) return BPPGPU_FAILURE;
  return batch_collect(b, out, nullptr);
}

extern "C" int bppgpu_batch_stage(bppgpu_batch * b, const unsigned int * mc, const unsigned int * mi, const double * mb,
                                  const unsigned int * oc, const bppgpu_partial_op * ops,
                                  const unsigned int * rclv, const int * rsc)
{
  return batch_stage(b, mc, mi, mb, oc, ops, rclv, rsc);
}

extern "C" int bppgpu_batch_run(bppgpu_batch * b)
{
  return batch_run(b, b->staged_mats, b->staged_ops || b->staged_roots, b->staged_roots, nullptr, 0);
}

extern "C" int bppgpu_batch_collect(bppgpu_batch * b, double * lnl_out, double * sum_out)
{
  return batch_collect(b, lnl_out, sum_out);
}

// blocks until the device has read the host arrays of the last stage / set_branch_lengths (they may be reused then)
extern "C" int bppgpu_batch_wait_inputs(bppgpu_batch * b)
{
  CUDA_CHECK(cudaSetDevice(b->e->device));
  if (b->inputs_pending) { fatal("bppgpu_batch_wait_inputs: the staged step has not been run yet (its copies are issued by run)"); return BPPGPU_FAILURE; }
  CUDA_CHECK(cudaEventSynchronize(b->ev_inputs));
  return BPPGPU_SUCCESS;
}

// Device-side SWAP_CLV_INDEX / SWAP_SCALER_INDEX / SWAP_PMAT_INDEX of every inner node and edge of the staged step
// (flip_indices_kernel).  The planned blocks of both parities are kept, so from the third step on a whole-tree
// proposal costs no planning at all.
extern "C" int bppgpu_batch_flip_indices(bppgpu_batch * b)
{
  bppgpu_engine * e = b->e;
  CUDA_CHECK(cudaSetDevice(e->device));
  if (!b->tables_on_device || !(b->staged_ops || b->staged_mats || b->staged_roots))
  { fatal("bppgpu_batch_flip_indices: no staged step to flip"); return BPPGPU_FAILURE; }
  if (!b->flip_checked)           // the dimensions of a locus never change: checked once per batch, not per proposal
  {
    for (auto * l : b->loci)
      if (l->clv_buffers != 2 * (l->tips - 1) || l->prob_matrices != 2 * (2 * l->tips - 2) ||
          (l->scale_buffers != 0 && l->scale_buffers != 2 * (l->tips - 1)))
      { fatal("bppgpu_batch_flip_indices: needs BPP's 2x buffer allocation (method.c:4137-4147)"); return BPPGPU_FAILURE; }
    b->flip_checked = true;
  }
  if (b->inputs_pending)
  {
    // staged in waves and never run: bring the whole step to the device first
    CUDA_CHECK(cudaStreamWaitEvent(b->stream, b->ev_tables, 0));
    batch_issue_copies(b, 0, b->n, b->stream);
    CUDA_CHECK(cudaEventRecord(b->ev_inputs, b->stream));
    b->inputs_pending = false;
  }
  e->launches++;
  flip_indices_kernel<<<(b->n * 32 + 127) / 128, 128, 0, b->stream>>>(
      e->d_loci, b->d_batch_locus, b->n,
      b->staged_ops ? (const unsigned int *)(b->d_in + b->o_op_off) : nullptr, (RawOp *)(b->d_in + b->o_ops),
      b->staged_mats ? (const unsigned int *)(b->d_in + b->o_mat_off) : nullptr, (unsigned int *)(b->d_in + b->o_mat_idx),
      b->staged_roots ? (unsigned int *)(b->d_in + b->o_root_clv) : nullptr, (int *)(b->d_in + b->o_root_sc));
  CUDA_CHECK(cudaGetLastError());
  b->parity ^= 1;
  if (b->kernel_kind != 1 && !b->d_blocks_par[b->parity]) CUDA_CHECK(cudaMalloc(&b->d_blocks_par[b->parity], b->blocks_cap));
  b->d_blocks = b->d_blocks_par[b->parity];
  return BPPGPU_SUCCESS;
}

// new branch lengths for the staged matrix list (same order and count): the only host data a whole-tree proposal
// has to send once its lists are on the device
extern "C" int bppgpu_batch_set_branch_lengths(bppgpu_batch * b, const double * branch_lengths)
{
  CUDA_CHECK(cudaSetDevice(b->e->device));
  if (!b->staged_mats || !b->tables_on_device) { fatal("bppgpu_batch_set_branch_lengths: no staged matrix list"); return BPPGPU_FAILURE; }
  const size_t bytes = (size_t)b->total_mats * 8;
  const void * src = branch_lengths;
  cudaPointerAttributes a;
  const bool pinned = cudaPointerGetAttributes(&a, src) == cudaSuccess && a.type == cudaMemoryTypeHost;
  if (!pinned)
  {
    cudaGetLastError();
    CUDA_CHECK(cudaStreamSynchronize(b->stream));           // the blob's previous content may still be in flight
    memcpy(b->h_in + b->o_mat_bl, branch_lengths, bytes);
    src = b->h_in + b->o_mat_bl;
  }
  if (b->inputs_pending) b->pend.mbl = (const double *)src;        // not uploaded yet: run() copies from here
  else
  {
    CUDA_CHECK(cudaMemcpyAsync(b->d_in + b->o_mat_bl, src, bytes, cudaMemcpyHostToDevice, b->stream));
    CUDA_CHECK(cudaEventRecord(b->ev_inputs, b->stream));
  }
  return BPPGPU_SUCCESS;
}

extern "C" int bppgpu_batch_full_pass(bppgpu_batch * b, const unsigned int * mc, const unsigned int * mi, const double * mb,
                                      const unsigned int * oc, const bppgpu_partial_op * ops,
                                      const unsigned int * rclv, const int * rsc, double * lnl_out, double * sum_out)
{
  if (!batch_stage(b, mc, mi, mb, oc, ops, rclv, rsc)) return BPPGPU_FAILURE;
  if (!batch_run(b, mc != nullptr, true, rclv != nullptr, nullptr, 0)) return BPPGPU_FAILURE;
  return batch_collect(b, lnl_out, sum_out);
}

// ------------------------------------------------------------------------------------ per-locus synchronous seam
static bppgpu_batch * self_batch(bppgpu_locus * l)
{
  if (!l->self_batch)
  {
    bppgpu_locus * arr[1] = { l };
    l->self_batch = bppgpu_batch_create(l->e, 1, arr);
  }
  return l->self_batch;
}

extern "C" int bppgpu_update_matrices(bppgpu_locus * l, unsigned int count, const unsigned int * idx, const double * bl)
{
  for (unsigned i = 0; i < count; ++i)
    if (idx[i] >= l->prob_matrices) { fatal("pmatrix index %u out of range", idx[i]); return BPPGPU_FAILURE; }
  return bppgpu_batch_update_matrices(self_batch(l), &count, idx, bl);
}

extern "C" int bppgpu_update_partials(bppgpu_locus * l, unsigned int count, const bppgpu_partial_op * ops)
{
  unsigned bad = 0;
  if (!locus_ops_valid(l, count, ops, &bad)) { fatal("update_partials: op %u has an index out of range", bad); return BPPGPU_FAILURE; }
  return bppgpu_batch_update_partials(self_batch(l), &count, ops);
}

static void ensure_persite(bppgpu_batch * b, size_t n)
{
  if (n > b->persite_cap)
  {
    CUDA_CHECK(cudaStreamSynchronize(b->stream));
    if (b->d_persite) cudaFree(b->d_persite);
    CUDA_CHECK(cudaMalloc(&b->d_persite, n * 8));
    b->persite_cap = n;
  }
}

extern "C" double bppgpu_root_loglikelihood(bppgpu_locus * l, unsigned int root_clv, int root_sc, double * persite)
{
  bppgpu_batch * b = self_batch(l);
  double out = 0;
  if (!batch_stage(b, nullptr, nullptr, nullptr, nullptr, nullptr, &root_clv, &root_sc)) return NAN;
  if (persite) ensure_persite(b, l->sites);
  if (!batch_run(b, false, true, true, persite ? b->d_persite : nullptr, 1)) return NAN;
  batch_collect(b, &out, nullptr);
  if (persite) CUDA_CHECK(cudaMemcpy(persite, b->d_persite, (size_t)l->sites * 8, cudaMemcpyDeviceToHost));
  return out;
}

extern "C" int bppgpu_root_likelihood_vector(bppgpu_locus * l, unsigned int root_clv, double * persite_lh)
{
  bppgpu_batch * b = self_batch(l);
  int none = -1;
  if (!batch_stage(b, nullptr, nullptr, nullptr, nullptr, nullptr, &root_clv, &none)) return BPPGPU_FAILURE;
  ensure_persite(b, l->sites);
  if (!batch_run(b, false, true, true, b->d_persite, 2)) return BPPGPU_FAILURE;
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  CUDA_CHECK(cudaMemcpy(persite_lh, b->d_persite, (size_t)l->sites * 8, cudaMemcpyDeviceToHost));
  return BPPGPU_SUCCESS;
}

extern "C" double bppgpu_root_loglikelihood_diploid(bppgpu_locus * l, unsigned int root_clv)
{
  if (!l->dev.dip_off) { fatal("locus has no diploid mapping"); return NAN; }
  bppgpu_batch * b = self_batch(l);
  int none = -1;
  if (!batch_stage(b, nullptr, nullptr, nullptr, nullptr, nullptr, &root_clv, &none)) return NAN;
  ensure_persite(b, l->sites);
  if (!batch_run(b, false, true, true, b->d_persite, 2)) return NAN;
  {
    ProfScope ps(b->e, b->stream, BPPGPU_KERNEL_FINISH);
    diploid_kernel<<<1, 256, 0, b->stream>>>(b->e->d_loci, l->id, b->d_persite, b->d_lnl_sum);
    CUDA_CHECK(cudaGetLastError());
  }
  double out = 0;
  CUDA_CHECK(cudaMemcpyAsync(b->h_out, b->d_lnl_sum, 8, cudaMemcpyDeviceToHost, b->stream));
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  out = b->h_out[0];
  return out;
}

// ------------------------------------------------------------------------------------ raw buffers
extern "C" int bppgpu_get_clv(bppgpu_locus * l, unsigned int clv_index, double * out)
{
  CUDA_CHECK(cudaSetDevice(l->e->device));
  const size_t P = l->sites, R = l->rate_cats, U = l->user_cats, S = l->states, nd = P * R * S;
  if (clv_index >= l->tips + l->clv_buffers) { fatal("clv index out of range"); return BPPGPU_FAILURE; }
  CUDA_CHECK(cudaDeviceSynchronize());
  if (clv_index >= l->tips || l->h_tip_dense_flag[clv_index])
  {
    const double * src = clv_index >= l->tips ? l->dev.clv + (size_t)(clv_index - l->tips) * nd : l->dev.tip_dense + (size_t)clv_index * nd;
    if (l->dev.site_stride == R * S && U == R) { CUDA_CHECK(cudaMemcpy(out, src, nd * 8, cudaMemcpyDeviceToHost)); return BPPGPU_SUCCESS; }
    // category-major on the device, or more categories there than the caller has: hand it out in the reference's
    // [site][cat][state] order with the caller's categories
    std::vector<double> raw(nd);
    CUDA_CHECK(cudaMemcpy(raw.data(), src, nd * 8, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < P; ++i) for (size_t r = 0; r < U; ++r)
      memcpy(out + (i * U + r) * S, &raw[i * l->dev.site_stride + r * l->dev.cat_stride], S * 8);
    return BPPGPU_SUCCESS;
  }
  for (size_t i = 0; i < P; ++i)          // expand the packed tip like set_tipclv, locus.c:540-555
  {
    const unsigned int c = get_code(l, clv_index, i);
    for (size_t r = 0; r < U; ++r) for (size_t j = 0; j < S; ++j) out[(i * U + r) * S + j] = (double)((c >> j) & 1u);
  }
  return BPPGPU_SUCCESS;
}

// The raw accessors are debug / test paths: they wait for everything in flight on the device (the engine's and the
// batches' streams are non-blocking, a plain cudaMemcpy would not order against them).
extern "C" int bppgpu_get_pmatrix(bppgpu_locus * l, unsigned int idx, double * out)
{
  CUDA_CHECK(cudaSetDevice(l->e->device));
  CUDA_CHECK(cudaDeviceSynchronize());
  if (idx >= l->prob_matrices) { fatal("pmatrix index out of range"); return BPPGPU_FAILURE; }
  const size_t ss = (size_t)l->states * l->states, nd = (size_t)l->rate_cats * ss;
  CUDA_CHECK(cudaMemcpy(out, l->dev.pmat + idx * nd, (size_t)l->user_cats * ss * 8, cudaMemcpyDeviceToHost));   // the caller's categories
  return BPPGPU_SUCCESS;
}

extern "C" int bppgpu_set_pmatrix(bppgpu_locus * l, unsigned int idx, const double * in)
{
  CUDA_CHECK(cudaSetDevice(l->e->device));
  CUDA_CHECK(cudaDeviceSynchronize());
  if (idx >= l->prob_matrices) { fatal("pmatrix index out of range"); return BPPGPU_FAILURE; }
  const size_t ss = (size_t)l->states * l->states, nd = (size_t)l->rate_cats * ss;
  CUDA_CHECK(cudaMemcpy(l->dev.pmat + idx * nd, in, (size_t)l->user_cats * ss * 8, cudaMemcpyHostToDevice));
  for (unsigned r = l->user_cats; r < l->rate_cats; ++r)       // padding categories: copies of category 0
    CUDA_CHECK(cudaMemcpy(l->dev.pmat + idx * nd + r * ss, in, ss * 8, cudaMemcpyHostToDevice));
  return BPPGPU_SUCCESS;
}

extern "C" int bppgpu_get_scaler(bppgpu_locus * l, unsigned int idx, unsigned int * out)
{
  CUDA_CHECK(cudaSetDevice(l->e->device));
  CUDA_CHECK(cudaDeviceSynchronize());
  if (idx >= l->scale_buffers) { fatal("scaler index out of range"); return BPPGPU_FAILURE; }
  CUDA_CHECK(cudaMemcpy(out, l->dev.scale + (size_t)idx * l->sites, (size_t)l->sites * 4, cudaMemcpyDeviceToHost));
  return BPPGPU_SUCCESS;
}

// ------------------------------------------------------------------------------------ collective (NCCL)
#include "comm.cuh"
