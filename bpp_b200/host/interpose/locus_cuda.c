/* locus_cuda.c -- bpp v4.8.7 itself on the B200 engine, with ZERO edits to the reference's sources.
 *
 * The reference has no plugin interface; its seam is the set of C functions of src/bpp.h:2032-2090
 * (implemented in src/locus.c).  This file defines functions with exactly those names and signatures:
 *
 *     locus_update_matrices      locus.c:2417        locus_update_partials     locus.c:2530
 *     locus_update_all_partials  locus.c:2523        locus_root_loglikelihood  locus.c:2573
 *     locus_destroy              locus.c:872         prop_mixing_update_gtrees prop_mixing.c:52
 *     propose_tau_update_gtrees  stree.c:4338        locus_propose_alpha_serial / _parallel  prop_gamma.c:175,197
 *
 * Linked into an executable in front of the reference built as a shared library (the recipe that compiles the reference, see INTEGRATION.md, builds
 * libbppref.so from the unmodified sources with -fPIC, so every call to these functions -- including the ones made
 * from inside locus.c -- goes through the PLT), ELF symbol interposition makes every caller of the seam
 * (method.c:4285-4297, prop_mixing.c:117-131, gtree.c, stree.c, prop_gamma.c, prop_rj.c, locus.c's own
 * propose_qrates / propose_freqs ...) land here.  With BPP_B200=1 in the environment the work is forwarded to the
 * engine through the C-ABI (include/bpp_b200.h); otherwise to the reference's own functions (dlsym RTLD_NEXT), so
 * the same binary is stock bpp.  INTEGRATION.md shows the equivalent source hooks (an arch bit next to
 * PLL_ATTRIB_ARCH_AVX2) a maintainer would add instead of interposition.
 *
 * What lives where:
 *   - locus_t stays the reference's struct, created by the reference's locus_create.  Its host CLV / scaler
 *     buffers are simply never written on the CUDA path (the host never reads inner CLVs).
 *   - per locus_t a bppgpu_locus handle is created lazily at the first call; tips come from the tip CLVs the
 *     reference has already filled (pll_set_tip_states -> locus->clv[tip], 0/1 doubles: the engine packs them),
 *     pattern weights / diploid mapping from the struct.
 *   - model parameters (frequencies, qrates, category rates, eigen-decomposition) are compared with a cached copy
 *     at every locus_update_matrices and re-sent when they changed: the proposals write them straight into the
 *     struct (locus.c:2703-3354), not through setters.
 *   - branch lengths: the reference's own locus_update_matrices is called first.  It writes node->length for every
 *     node of the traversal (strict clock core_pmatrix.c:711-715, relaxed clocks locus.c:1105-1193) and runs
 *     pll_update_eigen when needed; the engine then gets (pmatrix_index, length) and builds the matrices on the
 *     device.  (The host-side P-matrices the call also fills are unused.)
 *   - gnode_t index flips (SWAP_CLV_INDEX ...) stay host integers; every call passes them.
 *   - prop_mixing_update_gtrees and propose_tau_update_gtrees: the reference's loop body is kept (it is called as
 *     is), but while it runs the seam only RECORDS each locus' triplet and returns lnL = 0; afterwards all loci of
 *     the call go to the device as one batch (one upload, one planner + tree kernel + finish launch, one
 *     read-back) and gt->logl / lnacceptance / logl_diff get their lnL added.  BPP_B200_BATCH=0 turns that off.
 *   - everywhere else a locus' update_matrices / update_partials are queued and leave together with its next
 *     root_loglikelihood as one call (BPP_B200_FUSE=0: three synchronous calls, as the reference issues them).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>

#include "bpp.h"        /* the reference's own header (compile with -I<reference>/src) */
#include "bpp_b200.h"

/* ------------------------------------------------------------------ the reference's own implementations */
typedef void   (*fn_update_matrices)(locus_t *, gtree_t *, gnode_t **, stree_t *, long, unsigned int);
typedef void   (*fn_update_partials)(locus_t *, gnode_t **, unsigned int);
typedef void   (*fn_update_all_partials)(locus_t *, gtree_t *);
typedef double (*fn_root_loglikelihood)(locus_t *, gnode_t *, const unsigned int *, double *);
typedef void   (*fn_locus_destroy)(locus_t *);
typedef void   (*fn_mixing)(locus_t **, gtree_t **, stree_t *, long, long, double, long, double *);
typedef void   (*fn_tau)(locus_t **, gtree_t **, stree_t *, snode_t *, double, double, double, double, double, long, long,
                         snode_t **, unsigned int, unsigned int *, unsigned int *, double *, double *, long);

static fn_update_matrices     real_update_matrices;
static fn_update_partials     real_update_partials;
static fn_update_all_partials real_update_all_partials;
static fn_root_loglikelihood  real_root_loglikelihood;
static fn_locus_destroy       real_locus_destroy;
static fn_mixing              real_mixing;
static fn_tau                 real_tau;
typedef double (*fn_alpha_serial)(stree_t *, locus_t **, gtree_t **);
typedef void   (*fn_alpha_parallel)(stree_t *, locus_t **, gtree_t **, long, long, long, long *, long *);
static fn_alpha_serial        real_alpha_serial;
static fn_alpha_parallel      real_alpha_parallel;

static int g_enabled = -1, g_batching = 1, g_verbose = 0, g_fuse = 1, g_batch_alpha = 0;
static bppgpu_engine * g_engine;
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

/* statistics, printed at exit with BPP_B200_VERBOSE=1 */
static unsigned long long n_mat_calls, n_part_calls, n_root_calls, n_batches, n_batch_loci;

static void * next_sym(const char * name)
{
  void * p = dlsym(RTLD_NEXT, name);
  if (!p) fatal("locus_cuda: the reference's %s was not found (link libbppref.so behind this file)", name);
  return p;
}

static void report(void)
{
  if (g_verbose && g_enabled == 1)
    fprintf(stderr, "[bpp_b200] %s: update_matrices %llu, update_partials %llu, root_loglikelihood %llu calls; "
                    "%llu batched passes over %llu loci; %llu kernels launched\n", bppgpu_version(), n_mat_calls,
            n_part_calls, n_root_calls, n_batches, n_batch_loci, g_engine ? bppgpu_engine_launch_count(g_engine) : 0ULL);
}

static void init_once(void)
{
  const char * ev = getenv("BPP_B200");
  real_update_matrices = (fn_update_matrices)next_sym("locus_update_matrices");
  real_update_partials = (fn_update_partials)next_sym("locus_update_partials");
  real_update_all_partials = (fn_update_all_partials)next_sym("locus_update_all_partials");
  real_root_loglikelihood = (fn_root_loglikelihood)next_sym("locus_root_loglikelihood");
  real_locus_destroy = (fn_locus_destroy)next_sym("locus_destroy");
  real_mixing = (fn_mixing)next_sym("prop_mixing_update_gtrees");
  real_tau = (fn_tau)next_sym("propose_tau_update_gtrees");
  real_alpha_serial = (fn_alpha_serial)next_sym("locus_propose_alpha_serial");
  real_alpha_parallel = (fn_alpha_parallel)next_sym("locus_propose_alpha_parallel");
  g_enabled = ev && atoi(ev) != 0;
  if ((ev = getenv("BPP_B200_BATCH"))) g_batching = atoi(ev) != 0;
  if ((ev = getenv("BPP_B200_VERBOSE"))) g_verbose = atoi(ev);
  if ((ev = getenv("BPP_B200_FUSE"))) g_fuse = atoi(ev) != 0;
  if ((ev = getenv("BPP_B200_BATCH_ALPHA"))) g_batch_alpha = atoi(ev) != 0;
  if (g_enabled)
  {
    int dev = (ev = getenv("BPP_B200_DEVICE")) ? atoi(ev) : 0;
    unsigned int math = ((ev = getenv("BPP_B200_MATH")) && !strcmp(ev, "fma")) ? BPPGPU_MATH_FMA : BPPGPU_MATH_EXACT;
    g_engine = bppgpu_engine_create(dev, math);      /* no device -> the engine's fatal(): there is no CPU fallback */
    if (!g_engine) fatal("locus_cuda: cannot create the engine: %s", bppgpu_last_error());
    atexit(report);
  }
}

static int enabled(void)
{
  pthread_once(&g_once, init_once);
  return g_enabled;
}

/* ------------------------------------------------------------------ per-locus state, keyed by the locus_t pointer */
typedef struct
{
  locus_t * key;
  bppgpu_locus * h;
  int diploid_sent;
  double * model;              /* cached copy: freqs S, qrates S(S-1)/2, rates R, rate weights R, V, V^-1, lambda */
  int eigen_sent;
  /* scratch for one call */
  unsigned int cap;
  bppgpu_partial_op * ops;
  unsigned int * idx;
  double * bl;
  /* slot of this locus in the deferred batch that is being recorded (-1: none) */
  long slot;
  /* work recorded but not yet sent: the host never looks at P-matrices or CLVs, only at the next log-likelihood,
     so update_matrices / update_partials are queued and leave together with the root evaluation as ONE call
     (one upload, planner + tree kernel + finish, one read-back) */
  bppgpu_batch * one;          /* batch of this locus alone */
  unsigned int p_mats, p_ops;
} lstate_t;

#define TABLE_BITS 16
static lstate_t * g_table[1u << TABLE_BITS];
static pthread_rwlock_t g_table_lock = PTHREAD_RWLOCK_INITIALIZER;

static unsigned int hash_ptr(const void * p)
{
  uint64_t x = (uint64_t)(uintptr_t)p;
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 29;
  return (unsigned int)x & ((1u << TABLE_BITS) - 1);
}

static lstate_t * table_find(locus_t * l)
{
  unsigned int i = hash_ptr(l);
  lstate_t * s;
  pthread_rwlock_rdlock(&g_table_lock);
  while ((s = g_table[i]) && s->key != l) i = (i + 1) & ((1u << TABLE_BITS) - 1);
  pthread_rwlock_unlock(&g_table_lock);
  return s;
}

static void ensure_cap(lstate_t * s, unsigned int count)
{
  if (count <= s->cap) return;
  s->cap = 2 * count + 8;
  s->ops = (bppgpu_partial_op *)xrealloc(s->ops, s->cap * sizeof(bppgpu_partial_op));
  s->idx = (unsigned int *)xrealloc(s->idx, s->cap * sizeof(unsigned int));
  s->bl = (double *)xrealloc(s->bl, s->cap * sizeof(double));
}

/* send what is queued for a locus without evaluating a root (another consumer needs the buffers up to date) */
static void flush_pending(lstate_t * s)
{
  if (s->p_mats && !bppgpu_update_matrices(s->h, s->p_mats, s->idx, s->bl)) fatal("bppgpu_update_matrices: %s", bppgpu_last_error());
  if (s->p_ops && !bppgpu_update_partials(s->h, s->p_ops, s->ops)) fatal("bppgpu_update_partials: %s", bppgpu_last_error());
  s->p_mats = s->p_ops = 0;
}

static size_t model_doubles(const locus_t * l)
{
  const size_t S = l->states, R = l->rate_cats;
  return S + S * (S - 1) / 2 + 2 * R + 2 * S * S + S;
}

/* create the device mirror of a locus at its first use: same arguments as the reference's locus_create call
   (method.c:4137-4147), tips from the tip CLVs the reference has filled, weights, diploid mapping */
static lstate_t * state_of(locus_t * l)
{
  lstate_t * s = table_find(l);
  unsigned int i, t;
  if (s) return s;
  if (l->states_padded != l->states)
    fatal("locus_cuda: states_padded (%u) != states (%u) is not supported", l->states_padded, l->states);
  if (l->rate_matrices != 1) fatal("locus_cuda: rate_matrices must be 1");
  s = (lstate_t *)xcalloc(1, sizeof(lstate_t));
  s->key = l; s->slot = -1;
  s->h = bppgpu_locus_create(g_engine, l->dtype, l->model, l->tips, l->clv_buffers, l->states, l->sites,
                             l->rate_matrices, l->prob_matrices, l->rate_cats, l->scale_buffers,
                             l->attributes | BPPGPU_ATTRIB_ARCH_CUDA);
  if (!s->h) fatal("locus_cuda: bppgpu_locus_create failed: %s", bppgpu_last_error());
  {
    /* locus->clv[tip] is [site][cat][state] with every category a copy of the first (locus.c:540-555) */
    const size_t S = l->states, R = l->rate_cats, P = l->sites;
    double * tip = (double *)xmalloc(P * S * sizeof(double));
    for (t = 0; t < l->tips; ++t)
    {
      for (i = 0; i < P; ++i) memcpy(tip + i * S, l->clv[t] + i * R * S, S * sizeof(double));
      if (!bppgpu_set_tip_clv(s->h, t, tip, 0)) fatal("locus_cuda: bppgpu_set_tip_clv failed: %s", bppgpu_last_error());
    }
    free(tip);
  }
  if (!l->diploid) bppgpu_set_pattern_weights(s->h, l->pattern_weights);
  s->model = (double *)xcalloc(model_doubles(l), sizeof(double));
  s->model[0] = -1;            /* no valid frequency vector starts with -1: forces the first sync */
  ensure_cap(s, 2 * l->tips);
  pthread_rwlock_wrlock(&g_table_lock);
  i = hash_ptr(l);
  while (g_table[i]) i = (i + 1) & ((1u << TABLE_BITS) - 1);
  g_table[i] = s;
  pthread_rwlock_unlock(&g_table_lock);
  return s;
}

/* re-send what the proposals changed in the struct since the last call */
static void sync_model(lstate_t * s, locus_t * l)
{
  const size_t S = l->states, R = l->rate_cats, NP = S * (S - 1) / 2;
  double * c_freqs = s->model, * c_subst = c_freqs + S, * c_rates = c_subst + NP, * c_rw = c_rates + R,
         * c_ev = c_rw + R, * c_iev = c_ev + S * S, * c_lam = c_iev + S * S;
  const int eigen_model = !(l->dtype == BPP_DATA_DNA && l->model != BPP_DNA_MODEL_GTR);
  int params_changed = 0;
  if (memcmp(c_freqs, l->frequencies[0], S * sizeof(double)))
  {
    memcpy(c_freqs, l->frequencies[0], S * sizeof(double));
    bppgpu_set_frequencies(s->h, 0, c_freqs);
    params_changed = 1;
  }
  if (l->subst_params && l->subst_params[0] && memcmp(c_subst, l->subst_params[0], NP * sizeof(double)))
  {
    memcpy(c_subst, l->subst_params[0], NP * sizeof(double));
    bppgpu_set_subst_params(s->h, 0, c_subst);
    params_changed = 1;
  }
  if (memcmp(c_rates, l->rates, R * sizeof(double)))
  {
    memcpy(c_rates, l->rates, R * sizeof(double));
    bppgpu_set_category_rates(s->h, c_rates);
  }
  if (memcmp(c_rw, l->rate_weights, R * sizeof(double)))
  {
    memcpy(c_rw, l->rate_weights, R * sizeof(double));
    bppgpu_set_category_weights(s->h, c_rw);
  }
  /* the reference's own pll_update_eigen result (locus_update_matrices has just made it valid): handing it over
     keeps the P-matrices on the reference's decomposition instead of the engine's Jacobi one */
  if (eigen_model && l->eigen_decomp_valid[0] &&
      (params_changed || !s->eigen_sent || memcmp(c_lam, l->eigenvals[0], S * sizeof(double)) ||
       memcmp(c_ev, l->eigenvecs[0], S * S * sizeof(double))))
  {
    memcpy(c_ev, l->eigenvecs[0], S * S * sizeof(double));
    memcpy(c_iev, l->inv_eigenvecs[0], S * S * sizeof(double));
    memcpy(c_lam, l->eigenvals[0], S * sizeof(double));
    bppgpu_set_eigen(s->h, 0, c_ev, c_iev, c_lam);
    s->eigen_sent = 1;
  }
}

static void sync_diploid(lstate_t * s, locus_t * l)
{
  long i;
  unsigned long maplen = 0;
  if (!l->diploid || s->diploid_sent) return;
  for (i = 0; i < l->unphased_length; ++i) maplen += l->diploid_resolution_count[i];
  /* for diploid loci pattern_weights holds the weights of the UNPHASED sites (method.c:4172-4193) */
  if (!bppgpu_set_diploid(s->h, (unsigned int)l->unphased_length, l->diploid_resolution_count, l->diploid_mapping,
                          maplen, l->pattern_weights))
    fatal("locus_cuda: bppgpu_set_diploid failed: %s", bppgpu_last_error());
  s->diploid_sent = 1;
}

static unsigned int fill_ops(gnode_t ** trav, unsigned int count, bppgpu_partial_op * ops)
{
  unsigned int i;
  for (i = 0; i < count; ++i)        /* locus.c:2541-2570 */
  {
    gnode_t * node = trav[i], * lnode = node->left, * rnode = node->right;
    ops[i].parent_clv_index = node->clv_index;
    ops[i].left_clv_index = lnode->clv_index;
    ops[i].right_clv_index = rnode->clv_index;
    ops[i].left_pmatrix_index = lnode->pmatrix_index;
    ops[i].right_pmatrix_index = rnode->pmatrix_index;
    ops[i].parent_scaler_index = node->scaler_index;
    ops[i].left_scaler_index = lnode->scaler_index;
    ops[i].right_scaler_index = rnode->scaler_index;
  }
  return count;
}

/* ------------------------------------------------------------------ deferred batch (one per host thread) */
typedef struct
{
  int active;
  long first, count;                     /* loci [first, first+count) of the caller's array */
  locus_t ** loci;
  bppgpu_batch * batch;                  /* cached: same loci as last time */
  locus_t ** batch_key; long batch_n;
  /* step arrays, pinned */
  unsigned int * mcounts, * ocounts, * midx, * rclv; int * rsc; double * mbl; bppgpu_partial_op * ops;
  double * lnl;
  size_t cap_loci, cap_mats, cap_ops, n_mats, n_ops;
  long last_slot;                        /* slot of the locus whose calls are being recorded; slots only go up */
  unsigned char * touched;               /* the locus asked for its root log-likelihood */
  int overflow;                          /* a locus was visited out of order: the recording cannot be replayed */
} defer_t;

static __thread defer_t tl_defer;

static void defer_reserve(defer_t * d, size_t loci, size_t mats, size_t ops)
{
  if (loci > d->cap_loci)
  {
    bppgpu_host_free(d->mcounts); bppgpu_host_free(d->ocounts); bppgpu_host_free(d->rclv); bppgpu_host_free(d->rsc);
    free(d->lnl); free(d->touched);
    d->touched = (unsigned char *)xmalloc(loci);
    d->cap_loci = loci;
    d->mcounts = (unsigned int *)bppgpu_host_alloc(loci * sizeof(unsigned int));
    d->ocounts = (unsigned int *)bppgpu_host_alloc(loci * sizeof(unsigned int));
    d->rclv = (unsigned int *)bppgpu_host_alloc(loci * sizeof(unsigned int));
    d->rsc = (int *)bppgpu_host_alloc(loci * sizeof(int));
    d->lnl = (double *)xmalloc(loci * sizeof(double));
  }
  if (mats > d->cap_mats)
  {
    unsigned int * ni = (unsigned int *)bppgpu_host_alloc(2 * mats * sizeof(unsigned int));
    double * nb = (double *)bppgpu_host_alloc(2 * mats * sizeof(double));
    if (d->n_mats) { memcpy(ni, d->midx, d->n_mats * sizeof(unsigned int)); memcpy(nb, d->mbl, d->n_mats * sizeof(double)); }
    bppgpu_host_free(d->midx); bppgpu_host_free(d->mbl);
    d->midx = ni; d->mbl = nb; d->cap_mats = 2 * mats;
  }
  if (ops > d->cap_ops)
  {
    bppgpu_partial_op * no = (bppgpu_partial_op *)bppgpu_host_alloc(2 * ops * sizeof(bppgpu_partial_op));
    if (d->n_ops) memcpy(no, d->ops, d->n_ops * sizeof(bppgpu_partial_op));
    bppgpu_host_free(d->ops);
    d->ops = no; d->cap_ops = 2 * ops;
  }
}

/* ------------------------------------------------------------------ the seam */
void locus_update_matrices(locus_t * locus, gtree_t * gtree, gnode_t ** traversal, stree_t * stree, long msa_index,
                           unsigned int count)
{
  unsigned int i;
  lstate_t * s;
  defer_t * d = &tl_defer;
  if (!enabled()) { real_update_matrices(locus, gtree, traversal, stree, msa_index, count); return; }
  if (!opt_usedata) return;
  /* the reference: node->length for every node (any clock), pll_update_eigen if Q changed */
  real_update_matrices(locus, gtree, traversal, stree, msa_index, count);
  s = state_of(locus);
  sync_model(s, locus);
  __atomic_add_fetch(&n_mat_calls, 1, __ATOMIC_RELAXED);
  if (d->active && s->slot >= 0)
  {
    /* the arrays are concatenated in slot order: a locus may be visited once, later loci only afterwards; matrices
       behind this locus' partials would have to be replayed in between, which one fused pass cannot do */
    if (s->slot < d->last_slot || d->ocounts[s->slot]) { d->overflow = 1; return; }
    d->last_slot = s->slot;
    defer_reserve(d, 0, d->n_mats + count, 0);
    for (i = 0; i < count; ++i) { d->midx[d->n_mats + i] = traversal[i]->pmatrix_index; d->mbl[d->n_mats + i] = traversal[i]->length; }
    d->mcounts[s->slot] += count; d->n_mats += count;
    return;
  }
  /* partials queued behind earlier matrices must see those, not these: keep the order by flushing */
  if (s->p_ops) flush_pending(s);
  ensure_cap(s, s->p_mats + count);
  for (i = 0; i < count; ++i) { s->idx[s->p_mats + i] = traversal[i]->pmatrix_index; s->bl[s->p_mats + i] = traversal[i]->length; }
  s->p_mats += count;
  if (!g_fuse) flush_pending(s);
}

void locus_update_partials(locus_t * locus, gnode_t ** traversal, unsigned int count)
{
  lstate_t * s;
  defer_t * d = &tl_defer;
  if (!enabled()) { real_update_partials(locus, traversal, count); return; }
  if (!opt_usedata) return;
  s = state_of(locus);
  __atomic_add_fetch(&n_part_calls, 1, __ATOMIC_RELAXED);
  if (d->active && s->slot >= 0)
  {
    if (s->slot < d->last_slot) { d->overflow = 1; return; }
    d->last_slot = s->slot;
    defer_reserve(d, 0, 0, d->n_ops + count);
    fill_ops(traversal, count, d->ops + d->n_ops);
    d->ocounts[s->slot] += count; d->n_ops += count;
    return;
  }
  ensure_cap(s, s->p_ops + count);
  fill_ops(traversal, count, s->ops + s->p_ops);
  s->p_ops += count;
  if (!g_fuse) flush_pending(s);
}

static void all_partials_rec(gnode_t * node, gnode_t ** out, unsigned int * n)
{
  if (!node->left) return;
  all_partials_rec(node->left, out, n);
  all_partials_rec(node->right, out, n);
  out[(*n)++] = node;
}

void locus_update_all_partials(locus_t * locus, gtree_t * gtree)
{
  gnode_t ** trav;
  unsigned int n = 0;
  if (!enabled()) { real_update_all_partials(locus, gtree); return; }
  if (!opt_usedata) return;
  trav = (gnode_t **)xmalloc((gtree->tip_count + gtree->inner_count) * sizeof(gnode_t *));
  all_partials_rec(gtree->root, trav, &n);        /* locus.c:2482-2521: left, right, node */
  locus_update_partials(locus, trav, n);
  free(trav);
}

double locus_root_loglikelihood(locus_t * locus, gnode_t * root, const unsigned int * freqs_indices, double * persite_lnl)
{
  lstate_t * s;
  defer_t * d = &tl_defer;
  double logl;
  if (!enabled()) return real_root_loglikelihood(locus, root, freqs_indices, persite_lnl);
  if (!opt_usedata) return 0;
  s = state_of(locus);
  sync_diploid(s, locus);
  __atomic_add_fetch(&n_root_calls, 1, __ATOMIC_RELAXED);
  if (d->active && s->slot >= 0)
  {
    if (s->slot < d->last_slot || persite_lnl) { d->overflow = 1; return 0.0; }
    d->last_slot = s->slot;
    d->rclv[s->slot] = root->clv_index; d->rsc[s->slot] = root->scaler_index;
    d->touched[s->slot] = 1;
    return 0.0;                                   /* the batch adds the real value afterwards */
  }
  if (g_fuse && !persite_lnl)
  {
    /* matrices + partials + root of this locus in one call; a batch applies the diploid phase mean itself */
    const unsigned int rclv = root->clv_index;
    const int rsc = root->scaler_index;
    if (!s->one)
    {
      s->one = bppgpu_batch_create(g_engine, 1, &s->h);
      if (!s->one) fatal("bppgpu_batch_create: %s", bppgpu_last_error());
    }
    if (!bppgpu_batch_full_pass(s->one, s->p_mats ? &s->p_mats : NULL, s->idx, s->bl, s->p_ops ? &s->p_ops : NULL, s->ops,
                                &rclv, &rsc, &logl, NULL))
      fatal("bppgpu_batch_full_pass: %s", bppgpu_last_error());
    s->p_mats = s->p_ops = 0;
    return opt_bfbeta * logl;
  }
  flush_pending(s);
  if (locus->diploid) logl = bppgpu_root_loglikelihood_diploid(s->h, root->clv_index);     /* locus.c:2586-2615 */
  else logl = bppgpu_root_loglikelihood(s->h, root->clv_index, root->scaler_index, persite_lnl);
  return opt_bfbeta * logl;                       /* locus.c:2630 */
}

void locus_destroy(locus_t * locus)
{
  if (enabled())
  {
    lstate_t * s;
    unsigned int i = hash_ptr(locus);
    pthread_rwlock_wrlock(&g_table_lock);
    while ((s = g_table[i]) && s->key != locus) i = (i + 1) & ((1u << TABLE_BITS) - 1);
    if (s) s->key = (locus_t *)(uintptr_t)1;      /* tombstone: keeps the probe chains intact */
    pthread_rwlock_unlock(&g_table_lock);
    if (s)
    {
      pthread_mutex_lock(&g_mu);
      if (tl_defer.batch) { bppgpu_batch_destroy(tl_defer.batch); tl_defer.batch = NULL; tl_defer.batch_n = 0; }
      pthread_mutex_unlock(&g_mu);
      if (s->one) bppgpu_batch_destroy(s->one);
      bppgpu_locus_destroy(s->h);
      free(s->model); free(s->ops); free(s->idx); free(s->bl);
      s->h = NULL;
    }
  }
  else pthread_once(&g_once, init_once);
  real_locus_destroy(locus);
}

/* ------------------------------------------------------------------ batched caller loops
 * Both whole-data proposals walk `for each locus`, call the seam triplet for the loci they changed, add
 * logl - gtree->logl to a running sum and store gtree->logl = logl; the accept / reject decision is taken
 * afterwards for all loci together (prop_mixing.c:52-220 lnacceptance, stree.c:4338-4775 logl_diff).  The
 * reference's own loop is run as it is, with the seam recording instead of computing and returning logl = 0; then
 * ONE batch evaluates every recorded locus and the sums and gtree->logl get the real values added. */
static int defer_begin(defer_t * d, locus_t ** locus, long locus_start, long locus_count)
{
  long i;
  int same = d->batch && d->batch_n == locus_count;
  for (i = 0; same && i < locus_count; ++i) same = d->batch_key[i] == locus[locus_start + i];
  if (!same)
  {
    bppgpu_locus ** hs = (bppgpu_locus **)xmalloc(locus_count * sizeof(bppgpu_locus *));
    int uniform = 1;
    if (d->batch) bppgpu_batch_destroy(d->batch);
    d->batch = NULL;
    free(d->batch_key);
    d->batch_key = (locus_t **)xmalloc(locus_count * sizeof(locus_t *));
    for (i = 0; i < locus_count; ++i)
    {
      locus_t * l = locus[locus_start + i];
      d->batch_key[i] = l;
      hs[i] = state_of(l)->h;
      uniform = uniform && l->states == locus[locus_start]->states && l->rate_cats == locus[locus_start]->rate_cats;
    }
    d->batch_n = locus_count;
    if (uniform) d->batch = bppgpu_batch_create(g_engine, (unsigned int)locus_count, hs);
    free(hs);
    if (!d->batch) { d->batch_n = 0; return 0; }        /* mixed data types: a batch needs one (states, rate_cats) shape */
  }
  defer_reserve(d, (size_t)locus_count, 0, 0);
  for (i = 0; i < locus_count; ++i)
  {
    lstate_t * s = state_of(locus[locus_start + i]);
    if (s->p_mats || s->p_ops) flush_pending(s);
    s->slot = i;
    d->mcounts[i] = d->ocounts[i] = 0;
    d->touched[i] = 0;
  }
  d->n_mats = d->n_ops = 0; d->last_slot = 0; d->overflow = 0;
  d->first = locus_start; d->count = locus_count;
  d->active = 1;
  return 1;
}

/* returns the sum of the new log-likelihoods of the recorded loci, after storing them in gtree->logl */
static double defer_end(defer_t * d, locus_t ** locus, gtree_t ** gtree, const char * who)
{
  long i, touched = 0;
  double sum = 0;
  d->active = 0;
  for (i = 0; i < d->count; ++i) state_of(locus[d->first + i])->slot = -1;
  if (d->overflow)
    fatal("locus_cuda: %s visited its loci out of order or asked for per-site values; run with BPP_B200_BATCH=0", who);
  for (i = 0; i < d->count; ++i)
  {
    if (d->touched[i]) { ++touched; continue; }
    if (d->mcounts[i] || d->ocounts[i])
      fatal("locus_cuda: %s updated locus %ld without evaluating it; run with BPP_B200_BATCH=0", who, d->first + i);
    /* untouched: its root is evaluated along (and ignored) */
    d->rclv[i] = gtree[d->first + i]->root->clv_index; d->rsc[i] = gtree[d->first + i]->root->scaler_index;
  }
  if (!touched) return 0;
  if (!bppgpu_batch_full_pass(d->batch, d->mcounts, d->midx, d->mbl, d->ocounts, d->ops, d->rclv, d->rsc, d->lnl, NULL))
    fatal("bppgpu_batch_full_pass: %s", bppgpu_last_error());
  __atomic_add_fetch(&n_batches, 1, __ATOMIC_RELAXED);
  __atomic_add_fetch(&n_batch_loci, (unsigned long long)touched, __ATOMIC_RELAXED);
  for (i = 0; i < d->count; ++i)
    if (d->touched[i])
    {
      const double logl = opt_bfbeta * d->lnl[i];
      gtree[d->first + i]->logl = logl;       /* the loop stored 0 */
      sum += logl;
    }
  return sum;
}

/* mixing move, prop_mixing.c:52-220 */
void prop_mixing_update_gtrees(locus_t ** locus, gtree_t ** gtree, stree_t * stree, long locus_start, long locus_count,
                               double c, long thread_index, double * ret_lnacceptance)
{
  defer_t * d = &tl_defer;
  if (!enabled() || !g_batching || !opt_usedata || locus_count < 2 || d->active || !defer_begin(d, locus, locus_start, locus_count))
  {
    pthread_once(&g_once, init_once);
    real_mixing(locus, gtree, stree, locus_start, locus_count, c, thread_index, ret_lnacceptance);
    return;
  }
  real_mixing(locus, gtree, stree, locus_start, locus_count, c, thread_index, ret_lnacceptance);
  *ret_lnacceptance += defer_end(d, locus, gtree, "prop_mixing_update_gtrees");
}

/* species-tree node age move, stree.c:4338-4775: only the loci with gene-tree nodes in the affected populations'
   age window issue a triplet (a partial update along the marked nodes) */
void propose_tau_update_gtrees(locus_t ** loci, gtree_t ** gtree, stree_t * stree, snode_t * snode, double oldage,
                               double minage, double maxage, double minfactor, double maxfactor, long locus_start,
                               long locus_count, snode_t ** affected, unsigned int paffected_count,
                               unsigned int * ret_count_above, unsigned int * ret_count_below, double * ret_logl_diff,
                               double * ret_logpr_diff, long thread_index)
{
  defer_t * d = &tl_defer;
  if (!enabled() || !g_batching || !opt_usedata || locus_count < 2 || d->active || !defer_begin(d, loci, locus_start, locus_count))
  {
    pthread_once(&g_once, init_once);
    real_tau(loci, gtree, stree, snode, oldage, minage, maxage, minfactor, maxfactor, locus_start, locus_count, affected,
             paffected_count, ret_count_above, ret_count_below, ret_logl_diff, ret_logpr_diff, thread_index);
    return;
  }
  real_tau(loci, gtree, stree, snode, oldage, minage, maxage, minfactor, maxfactor, locus_start, locus_count, affected,
           paffected_count, ret_count_above, ret_count_below, ret_logl_diff, ret_logpr_diff, thread_index);
  *ret_logl_diff += defer_end(d, loci, gtree, "propose_tau_update_gtrees");
}

/* ------------------------------------------------------------------ batched alpha move (prop_gamma.c:53-226)
 * The reference proposes, evaluates and accepts locus by locus.  Loci are independent, so the move is the same
 * Markov chain when all loci propose first, ONE batch evaluates them (full-tree passes: new category rates change
 * every P-matrix), and every locus then takes its own accept / reject decision.  Only the order in which the
 * random numbers are consumed changes (all proposal draws, then the acceptance draws), so mcmc.txt is a different,
 * equally valid realisation -- which is why this is opt-in (BPP_B200_BATCH_ALPHA=1); the default keeps the
 * reference's draw order and evaluates locus by locus.  Index flips and roll-back are the reference's
 * (prop_gamma.c:100-158).
 */
#define IP_SWAP_CLV_INDEX(n,i)    ((n)+((i)-1)%(2*(n)-2))
#define IP_SWAP_SCALER_INDEX(n,i) (((n)+((i)-1))%(2*(n)-2))
#define IP_SWAP_PMAT_INDEX(e,i)   (((e)+(i))%((e)<<1))

static int batched_alpha(stree_t * stree, locus_t ** locus, gtree_t ** gtree, long start, long count, long thread_index,
                         long * p_candidates, long * p_accepted)
{
  defer_t * d = &tl_defer;
  long i, accepted = 0, candidates = 0;
  unsigned int m, n, maxn = 0;
  double * alpha_old, * lnacc, * old_logl, * old_rates;
  gnode_t ** gt_nodes;
  const double minv = -99, maxv = 99;
  if (d->active || !defer_begin(d, locus, start, count)) return 0;
  for (i = start; i < start + count; ++i)
    if (gtree[i]->tip_count + gtree[i]->inner_count > maxn) maxn = gtree[i]->tip_count + gtree[i]->inner_count;
  alpha_old = (double *)xmalloc(count * sizeof(double));
  lnacc = (double *)xmalloc(count * sizeof(double));
  old_logl = (double *)xmalloc(count * sizeof(double));
  old_rates = (double *)xmalloc(count * 64 * sizeof(double));
  gt_nodes = (gnode_t **)xmalloc(maxn * sizeof(gnode_t *));
  /* phase 1: every candidate proposes, flips its indices and records its full-tree pass */
  for (i = start; i < start + count; ++i)
  {
    locus_t * l = locus[i];
    gtree_t * gt = gtree[i];
    double loga_old, loga_new;
    if (!(l->dtype == BPP_DATA_DNA && l->rate_cats > 1 && l->rate_cats <= 64)) continue;
    ++candidates;
    alpha_old[i - start] = l->rates_alpha;
    old_logl[i - start] = gt->logl;
    loga_old = log(l->rates_alpha);
    loga_new = loga_old + opt_finetune_alpha * legacy_rnd_symmetrical(thread_index);
    loga_new = reflect(loga_new, minv, maxv, thread_index);
    lnacc[i - start] = loga_new - loga_old;
    l->rates_alpha = exp(loga_new);
    memcpy(old_rates + (i - start) * 64, l->rates, l->rate_cats * sizeof(double));
    pll_compute_gamma_cats(l->rates_alpha, l->rates_alpha, l->rate_cats, l->rates, PLL_GAMMA_RATES_MEAN);
    for (m = 0, n = 0; m < gt->tip_count + gt->inner_count; ++m)
    {
      gnode_t * p = gt->nodes[m];
      if (p->parent) { p->pmatrix_index = IP_SWAP_PMAT_INDEX(gt->edge_count, p->pmatrix_index); gt_nodes[n++] = p; }
    }
    locus_update_matrices(l, gt, gt_nodes, stree, i, n);
    n = 0;
    all_partials_rec(gt->root, gt_nodes, &n);
    for (m = 0; m < n; ++m)
    {
      gt_nodes[m]->clv_index = IP_SWAP_CLV_INDEX(gt->tip_count, gt_nodes[m]->clv_index);
      if (opt_scaling) gt_nodes[m]->scaler_index = IP_SWAP_SCALER_INDEX(gt->tip_count, gt_nodes[m]->scaler_index);
    }
    locus_update_partials(l, gt_nodes, n);
    (void)locus_root_loglikelihood(l, gt->root, l->param_indices, NULL);
  }
  /* phase 2: one batch; gtree->logl of every candidate now holds the PROPOSED log-likelihood */
  (void)defer_end(d, locus, gtree, "locus_propose_alpha");
  /* phase 3: per-locus decisions (prop_gamma.c:128-158) */
  for (i = start; i < start + count; ++i)
  {
    locus_t * l = locus[i];
    gtree_t * gt = gtree[i];
    double lnacceptance, alpha_new;
    if (!(l->dtype == BPP_DATA_DNA && l->rate_cats > 1 && l->rate_cats <= 64)) continue;
    alpha_new = l->rates_alpha;
    lnacceptance = lnacc[i - start] + (gt->logl - old_logl[i - start]);
    lnacceptance += (opt_alpha_alpha - 1) * log(alpha_new / alpha_old[i - start]) - opt_alpha_beta * (alpha_new - alpha_old[i - start]);
    if (lnacceptance >= -1e-10 || legacy_rndu(thread_index) < exp(lnacceptance)) { ++accepted; continue; }
    /* rejected: the old halves of the double buffers still hold the old state */
    l->rates_alpha = alpha_old[i - start];
    gt->logl = old_logl[i - start];
    for (m = 0; m < gt->tip_count + gt->inner_count; ++m)
    {
      gnode_t * p = gt->nodes[m];
      if (p->parent) p->pmatrix_index = IP_SWAP_PMAT_INDEX(gt->edge_count, p->pmatrix_index);
    }
    n = 0;
    all_partials_rec(gt->root, gt_nodes, &n);
    for (m = 0; m < n; ++m)
    {
      gt_nodes[m]->clv_index = IP_SWAP_CLV_INDEX(gt->tip_count, gt_nodes[m]->clv_index);
      if (opt_scaling) gt_nodes[m]->scaler_index = IP_SWAP_SCALER_INDEX(gt->tip_count, gt_nodes[m]->scaler_index);
    }
    pll_set_category_rates(l, old_rates + (i - start) * 64);
  }
  free(alpha_old); free(lnacc); free(old_logl); free(old_rates); free(gt_nodes);
  *p_candidates = candidates; *p_accepted = accepted;
  return 1;
}

double locus_propose_alpha_serial(stree_t * stree, locus_t ** locus, gtree_t ** gtree)
{
  long candidates = 0, accepted = 0;
  if (enabled() && g_batching && g_batch_alpha && opt_usedata && opt_locus_count > 1 &&
      batched_alpha(stree, locus, gtree, 0, opt_locus_count, 0, &candidates, &accepted))
    return accepted ? (double)accepted / candidates : 0;
  pthread_once(&g_once, init_once);
  return real_alpha_serial(stree, locus, gtree);
}

void locus_propose_alpha_parallel(stree_t * stree, locus_t ** locus, gtree_t ** gtree, long locus_start, long locus_count,
                                  long thread_index, long * p_proposal_count, long * p_accepted)
{
  if (enabled() && g_batching && g_batch_alpha && opt_usedata && locus_count > 1 &&
      batched_alpha(stree, locus, gtree, locus_start, locus_count, thread_index, p_proposal_count, p_accepted))
    return;
  pthread_once(&g_once, init_once);
  real_alpha_parallel(stree, locus, gtree, locus_start, locus_count, thread_index, p_proposal_count, p_accepted);
}
