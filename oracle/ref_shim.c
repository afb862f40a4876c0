/* oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A flat, ctypes-friendly driver around the UNMODIFIED reference (bpp v4.8.7).
 * It is compiled together with the reference sources (oracle/Makefile) into
 * oracle/_ref/libbppref.so and is used only by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  Nothing in the product
 * path (bpp_b200/, include/) links, loads or calls it.
 *
 * Everything numerical is done by the reference's own functions:
 *   locus_create                 locus.c:622
 *   pll_set_tip_states           locus.c:561   (set_tipclv locus.c:525)
 *   pll_set_tip_clv              locus.c:596
 *   pll_set_pattern_weights      locus.c
 *   pll_set_frequencies          locus.c:889
 *   pll_set_subst_params         locus.c:877
 *   pll_set_category_rates       locus.c
 *   locus_update_matrices        locus.c:2417  (jc69 :2325, eigen core_pmatrix.c:239,674)
 *   locus_update_partials        locus.c:2530  (core_partials*.c)
 *   locus_root_loglikelihood     locus.c:2573  (core_likelihood*.c)
 *   pll_compute_gamma_cats       gamma.c:221
 * This file only builds gnode_t / gtree_t scaffolding (bpp.h:692-774) the way
 * gtree.c:2395-2399,2664-2675 initialises it, and partitions loci over
 * pthreads the way threads.c:234-263 (load_balance_none) does.
 */
#include "bpp.h"
#include <pthread.h>
#include <time.h>

typedef struct
{
  locus_t * locus;
  gtree_t * gtree;
  gnode_t * nodes;       /* contiguous storage, node id == creation order */
  gnode_t ** trav;       /* scratch, 2T-1 entries */
  unsigned int tips;
} ref_locus_t;

typedef struct
{
  int n;
  unsigned int dtype, model, states, rate_cats, attributes;
  int scaling;
  ref_locus_t * l;
} ref_set_t;

static int g_random_ready = 0;

ref_set_t * ref_set_create(int n_loci, int dtype, int model, int states,
                           int rate_cats, int scaling, int attributes)
{
  ref_set_t * s = (ref_set_t *)calloc(1, sizeof(ref_set_t));
  s->n = n_loci;
  s->dtype = dtype; s->model = model; s->states = states;
  s->rate_cats = rate_cats; s->scaling = scaling; s->attributes = attributes;
  s->l = (ref_locus_t *)calloc(n_loci, sizeof(ref_locus_t));

  /* globals the locus seam reads */
  opt_alpha_cats  = rate_cats;
  opt_alpha_alpha = 1; opt_alpha_beta = 1;
  opt_usedata = 1;
  opt_bfbeta  = 1;
  opt_clock   = BPP_CLOCK_GLOBAL;
  opt_scaling = scaling;
  opt_arch    = attributes;
  opt_threads = 1;
  (void)g_random_ready;
  return s;
}

static void ref_locus_free(ref_locus_t * rl)
{
  if (!rl->locus) return;
  locus_destroy(rl->locus);
  free(rl->gtree->nodes);
  free(rl->gtree);
  free(rl->nodes);
  free(rl->trav);
  memset(rl, 0, sizeof(*rl));
}

void ref_set_destroy(ref_set_t * s)
{
  int i;
  for (i = 0; i < s->n; ++i) ref_locus_free(s->l + i);
  free(s->l);
  free(s);
}

/* mirrors method.c:4137-4147: clv_buffers = 2*inner, prob_matrices = 2*edges,
   scale_buffers = opt_scaling ? 2*inner : 0, rate_matrices = 1 */
int ref_locus_create(ref_set_t * s, int i, int tips, int sites)
{
  unsigned int k;
  ref_locus_t * rl = s->l + i;
  ref_locus_free(rl);
  unsigned int inner = tips - 1, edges = 2 * tips - 2, nn = 2 * tips - 1;

  rl->tips = tips;
  rl->locus = locus_create(s->dtype, s->model, tips, 2 * inner, s->states, sites,
                           1, 2 * edges, s->rate_cats,
                           s->scaling ? 2 * inner : 0, s->attributes);
  rl->nodes = (gnode_t *)calloc(nn, sizeof(gnode_t));
  rl->trav  = (gnode_t **)calloc(nn, sizeof(gnode_t *));
  rl->gtree = (gtree_t *)calloc(1, sizeof(gtree_t));
  rl->gtree->tip_count = tips;
  rl->gtree->inner_count = inner;
  rl->gtree->edge_count = edges;
  rl->gtree->rate_mui = 1;
  rl->gtree->nodes = (gnode_t **)calloc(nn, sizeof(gnode_t *));
  for (k = 0; k < nn; ++k)
  {
    gnode_t * x = rl->nodes + k;
    rl->gtree->nodes[k] = x;
    x->node_index = k;
    x->clv_index = k;                 /* gtree.c:2395,2664 */
    x->pmatrix_index = k;             /* gtree.c:2397,2666 */
    x->scaler_index = (k < (unsigned)tips || !s->scaling)
                        ? PLL_SCALE_BUFFER_NONE : (int)(k - tips); /* :2399,2675 */
  }
  rl->gtree->root = rl->nodes + nn - 1;
  return 1;
}

int ref_set_tip_states(ref_set_t * s, int i, int tip, const char * seq)
{
  const unsigned int * map = (s->states == 4) ? pll_map_nt : pll_map_aa;
  return pll_set_tip_states(s->l[i].locus, tip, map, seq);
}

int ref_set_tip_clv(ref_set_t * s, int i, int tip, const double * clv)
{
  return pll_set_tip_clv(s->l[i].locus, tip, clv, 0);
}

void ref_set_weights(ref_set_t * s, int i, const unsigned int * w)
{
  pll_set_pattern_weights(s->l[i].locus, w);
}

void ref_set_model(ref_set_t * s, int i, const double * freqs,
                   const double * subst_params, const double * rates)
{
  locus_t * L = s->l[i].locus;
  if (freqs) pll_set_frequencies(L, 0, freqs);
  if (subst_params) pll_set_subst_params(L, 0, subst_params);
  if (rates) pll_set_category_rates(L, rates);
}

int ref_gamma_rates(double alpha, int cats, double * out)
{
  return pll_compute_gamma_cats(alpha, alpha, cats, out, PLL_GAMMA_RATES_MEAN);
}

/* inner node k (0-based) has id tips+k; left/right are node ids; times has
   2T-1 entries; children must have smaller ids than their parent */
void ref_set_tree(ref_set_t * s, int i, const int * left, const int * right,
                  const double * times, double rate_mui)
{
  ref_locus_t * rl = s->l + i;
  unsigned int T = rl->tips, k, nn = 2 * T - 1;
  for (k = 0; k < nn; ++k)
  {
    rl->nodes[k].time = times[k];
    rl->nodes[k].left = rl->nodes[k].right = NULL;
    rl->nodes[k].parent = NULL;
  }
  for (k = 0; k < T - 1; ++k)
  {
    gnode_t * x = rl->nodes + T + k;
    x->left = rl->nodes + left[k];
    x->right = rl->nodes + right[k];
    x->left->parent = x;
    x->right->parent = x;
  }
  rl->gtree->rate_mui = rate_mui;
  for (k = 0; k < nn; ++k)
    if (!rl->nodes[k].parent) rl->gtree->root = rl->nodes + k;
}

void ref_set_times(ref_set_t * s, int i, const double * times)
{
  ref_locus_t * rl = s->l + i;
  unsigned int k, nn = 2 * rl->tips - 1;
  for (k = 0; k < nn; ++k) rl->nodes[k].time = times[k];
}

/* field: 0 clv_index, 1 scaler_index, 2 pmatrix_index */
int ref_node_get(ref_set_t * s, int i, int node, int field)
{
  gnode_t * x = s->l[i].nodes + node;
  return field == 0 ? (int)x->clv_index : field == 1 ? x->scaler_index
                                                     : (int)x->pmatrix_index;
}

void ref_node_set(ref_set_t * s, int i, int node, int field, int v)
{
  gnode_t * x = s->l[i].nodes + node;
  if (field == 0) x->clv_index = v;
  else if (field == 1) x->scaler_index = v;
  else x->pmatrix_index = v;
}

void ref_update_matrices(ref_set_t * s, int i, int count, const int * node_ids)
{
  ref_locus_t * rl = s->l + i;
  int k;
  for (k = 0; k < count; ++k) rl->trav[k] = rl->nodes + node_ids[k];
  locus_update_matrices(rl->locus, rl->gtree, rl->trav, NULL, i, count);
}

void ref_update_partials(ref_set_t * s, int i, int count, const int * node_ids)
{
  ref_locus_t * rl = s->l + i;
  int k;
  for (k = 0; k < count; ++k) rl->trav[k] = rl->nodes + node_ids[k];
  locus_update_partials(rl->locus, rl->trav, count);
}

double ref_root_loglikelihood(ref_set_t * s, int i, double * persite_or_null)
{
  ref_locus_t * rl = s->l + i;
  return locus_root_loglikelihood(rl->locus, rl->gtree->root,
                                  rl->locus->param_indices, persite_or_null);
}

/* diploid branch of locus_root_loglikelihood (locus.c:2586-2615); the arrays
   are copied and owned by the locus like diploid.c does */
void ref_set_diploid(ref_set_t * s, int i, int unphased_length,
                     const unsigned long * resolution_count,
                     const unsigned long * mapping, int mapping_len)
{
  locus_t * L = s->l[i].locus;
  L->diploid = 1;
  L->unphased_length = unphased_length;
  L->diploid_resolution_count = (unsigned long *)xmalloc(unphased_length * sizeof(unsigned long));
  memcpy(L->diploid_resolution_count, resolution_count, unphased_length * sizeof(unsigned long));
  L->diploid_mapping = (unsigned long *)xmalloc(mapping_len * sizeof(unsigned long));
  memcpy(L->diploid_mapping, mapping, mapping_len * sizeof(unsigned long));
  L->likelihood_vector = (double *)xmalloc(L->sites * sizeof(double));
}

static void post_order(gnode_t * x, gnode_t ** out, unsigned int * n)
{
  if (!x->left) return;
  post_order(x->left, out, n);
  post_order(x->right, out, n);
  out[(*n)++] = x;
}

/* one unit of the benchmark metric: all 2T-2 P-matrices, all T-1 inner CLVs in
   recursive post-order (prop_mixing.c:28-50), one root lnL */
double ref_full_pass(ref_set_t * s, int i)
{
  ref_locus_t * rl = s->l + i;
  unsigned int k, n = 0, nn = 2 * rl->tips - 1;
  for (k = 0; k < nn; ++k)
    if (rl->nodes[k].parent) rl->trav[n++] = rl->nodes + k;
  locus_update_matrices(rl->locus, rl->gtree, rl->trav, NULL, i, n);
  n = 0;
  post_order(rl->gtree->root, rl->trav, &n);
  locus_update_partials(rl->locus, rl->trav, n);
  return locus_root_loglikelihood(rl->locus, rl->gtree->root,
                                  rl->locus->param_indices, NULL);
}

typedef struct { ref_set_t * s; int first, count, passes; double * out; } work_t;

static void * worker(void * arg)
{
  work_t * w = (work_t *)arg;
  int p, i;
  for (p = 0; p < w->passes; ++p)
    for (i = w->first; i < w->first + w->count; ++i)
      w->out[i] = ref_full_pass(w->s, i);
  return NULL;
}

/* static contiguous partition of loci over pthreads (threads.c:234-263);
   returns wall seconds for `passes` passes over loci [first, first+count) */
double ref_full_pass_all(ref_set_t * s, int first, int count, int nthreads,
                         int passes, double * lnl_out)
{
  struct timespec t0, t1;
  int t, per, rem, start;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > count) nthreads = count;
  pthread_t * th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
  work_t * w = (work_t *)calloc(nthreads, sizeof(work_t));
  per = count / nthreads; rem = count % nthreads; start = first;
  for (t = 0; t < nthreads; ++t)
  {
    w[t].s = s; w[t].first = start; w[t].count = per + (t < rem ? 1 : 0);
    w[t].passes = passes; w[t].out = lnl_out;
    start += w[t].count;
  }
  clock_gettime(CLOCK_MONOTONIC, &t0);
  if (nthreads == 1) worker(w);
  else
  {
    for (t = 0; t < nthreads; ++t) pthread_create(th + t, NULL, worker, w + t);
    for (t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(th); free(w);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* The likelihood part of a gene-tree age move (gtree.c:5437-5467) for node `node` of locus i with its new age:
   flip and rebuild the P-matrices of the 2-3 edges around the node, flip the CLV / scaler indices of the root
   path, update those partials, evaluate the root.  The move is kept ("accepted"). */
#define SHIM_SWAP_CLV(n,i)    ((n)+((i)-1)%(2*(n)-2))
#define SHIM_SWAP_SCALER(n,i) (((n)+((i)-1))%(2*(n)-2))
#define SHIM_SWAP_PMAT(e,i)   (((e)+(i))%((e)<<1))
double ref_age_move(ref_set_t * s, int i, int node_id, double new_age)
{
  ref_locus_t * rl = s->l + i;
  gnode_t * node = rl->nodes + node_id, * temp;
  unsigned int k = 0, j, T = rl->tips;
  node->time = new_age;
  rl->trav[k++] = node->left;
  rl->trav[k++] = node->right;
  if (node->parent) rl->trav[k++] = node;
  for (j = 0; j < k; ++j) rl->trav[j]->pmatrix_index = SHIM_SWAP_PMAT(rl->gtree->edge_count, rl->trav[j]->pmatrix_index);
  locus_update_matrices(rl->locus, rl->gtree, rl->trav, NULL, i, k);
  for (k = 0, temp = node; temp; temp = temp->parent)
  {
    rl->trav[k++] = temp;
    temp->clv_index = SHIM_SWAP_CLV(T, temp->clv_index);
    if (s->scaling) temp->scaler_index = SHIM_SWAP_SCALER(T, temp->scaler_index);
  }
  locus_update_partials(rl->locus, rl->trav, k);
  return locus_root_loglikelihood(rl->locus, rl->gtree->root, rl->locus->param_indices, NULL);
}

typedef struct { ref_set_t * s; int first, count; const int * nodes; const double * ages; double * out; } agework_t;

static void * age_worker(void * arg)
{
  agework_t * w = (agework_t *)arg;
  int i;
  for (i = w->first; i < w->first + w->count; ++i) w->out[i] = ref_age_move(w->s, i, w->nodes[i], w->ages[i]);
  return NULL;
}

/* one age move per locus for loci [first, first+count), static partition over pthreads; returns wall seconds */
double ref_age_move_all(ref_set_t * s, int first, int count, int nthreads, const int * node_ids,
                        const double * new_ages, double * lnl_out)
{
  struct timespec t0, t1;
  int t, per, rem, start;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > count) nthreads = count;
  pthread_t * th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
  agework_t * w = (agework_t *)calloc(nthreads, sizeof(agework_t));
  per = count / nthreads; rem = count % nthreads; start = first;
  for (t = 0; t < nthreads; ++t)
  {
    w[t].s = s; w[t].first = start; w[t].count = per + (t < rem ? 1 : 0);
    w[t].nodes = node_ids; w[t].ages = new_ages; w[t].out = lnl_out;
    start += w[t].count;
  }
  clock_gettime(CLOCK_MONOTONIC, &t0);
  if (nthreads == 1) age_worker(w);
  else
  {
    for (t = 0; t < nthreads; ++t) pthread_create(th + t, NULL, age_worker, w + t);
    for (t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(th); free(w);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* raw buffer access for parity checks */
const double * ref_clv(ref_set_t * s, int i, int clv_index)
{ return s->l[i].locus->clv[clv_index]; }
const double * ref_pmatrix(ref_set_t * s, int i, int pmatrix_index)
{ return s->l[i].locus->pmatrix[pmatrix_index]; }
const unsigned int * ref_scaler(ref_set_t * s, int i, int scaler_index)
{ return s->l[i].locus->scale_buffer[scaler_index]; }
const double * ref_eigenvecs(ref_set_t * s, int i) { return s->l[i].locus->eigenvecs[0]; }
const double * ref_inv_eigenvecs(ref_set_t * s, int i) { return s->l[i].locus->inv_eigenvecs[0]; }
const double * ref_eigenvals(ref_set_t * s, int i) { return s->l[i].locus->eigenvals[0]; }
const double * ref_rates(ref_set_t * s, int i) { return s->l[i].locus->rates; }
const double * ref_freqs(ref_set_t * s, int i) { return s->l[i].locus->frequencies[0]; }
const double * ref_likelihood_vector(ref_set_t * s, int i) { return s->l[i].locus->likelihood_vector; }

/* empirical amino-acid model tables (maps.c:299,868), read-only data */
const double * ref_aa_rates_lg(void) { return pll_aa_rates_lg; }
const double * ref_aa_freqs_lg(void) { return pll_aa_freqs_lg; }
const unsigned int * ref_map_nt(void) { return pll_map_nt; }
const unsigned int * ref_map_aa(void) { return pll_map_aa; }
