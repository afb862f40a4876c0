/* bpp_gpu_host.c -- see bpp_gpu_host.h.  Pure marshalling: every number is computed on the GPU. */
#include <stdlib.h>
#include <string.h>
#include "bpp_gpu_host.h"

/* ------------------------------------------------------------------ gene tree scaffolding */
gtree_gpu_t * gtree_create_gpu(unsigned int tips, const int * left, const int * right, const double * times,
                               double rate_mui, int scaling)
{
  unsigned int k, nn = 2 * tips - 1;
  gtree_gpu_t * t = (gtree_gpu_t *)calloc(1, sizeof(gtree_gpu_t));
  gnode_gpu_t * store = (gnode_gpu_t *)calloc(nn, sizeof(gnode_gpu_t));
  t->tip_count = tips; t->inner_count = tips - 1; t->edge_count = 2 * tips - 2;
  t->rate_mui = rate_mui;
  t->nodes = (gnode_gpu_t **)calloc(nn, sizeof(gnode_gpu_t *));
  for (k = 0; k < nn; ++k)
  {
    gnode_gpu_t * x = store + k;
    t->nodes[k] = x;
    x->node_index = k;
    x->clv_index = k;                       /* gtree.c:2395,2664 */
    x->pmatrix_index = k;                   /* gtree.c:2397,2666 */
    x->scaler_index = (k < tips || !scaling) ? BPPGPU_SCALE_BUFFER_NONE : (int)(k - tips);   /* :2399,2675 */
    x->time = times[k];
  }
  for (k = 0; k + 1 < tips; ++k)
  {
    gnode_gpu_t * x = store + tips + k;
    x->left = store + left[k]; x->right = store + right[k];
    x->left->parent = x; x->right->parent = x;
  }
  t->root = store + nn - 1;
  for (k = 0; k < nn; ++k) if (!store[k].parent) t->root = store + k;
  return t;
}

void gtree_set_relaxed_clock_gpu(gtree_gpu_t * t, const stree_gpu_t * stree, const int * pops, double rate_scale)
{
  unsigned int k;
  t->stree = stree; t->rate_scale = rate_scale;
  if (stree && pops) for (k = 0; k < t->tip_count + t->inner_count; ++k) t->nodes[k]->pop = pops[k];
}

double gtree_branch_length_gpu(const gtree_gpu_t * t, const gnode_gpu_t * node)
{
  const stree_gpu_t * st = t->stree;
  double time, length = 0;
  int start, end;
  if (!st) return (node->parent->time - node->time) * t->rate_mui;
  /* walk up the species tree from the node's population to its parent's: every species branch crossed contributes
     its time span times its rate, the last span lies in the parent's population */
  time = node->time; start = node->pop; end = node->parent->pop;
  while (start != end)
  {
    const int pop = start;
    start = st->parent[start];
    length += (st->tau[start] - time) * st->brate[pop] * t->rate_scale;
    time = st->tau[start];
  }
  length += (node->parent->time - time) * st->brate[end] * t->rate_scale;
  return length;
}

void gtree_destroy_gpu(gtree_gpu_t * t)
{
  if (!t) return;
  free(t->nodes[0]);            /* contiguous storage */
  free(t->nodes);
  free(t);
}

static void all_partials_recursive(gnode_gpu_t * node, unsigned int * trav_size, gnode_gpu_t ** outbuffer)
{
  if (!node->left) return;
  all_partials_recursive(node->left, trav_size, outbuffer);
  all_partials_recursive(node->right, trav_size, outbuffer);
  outbuffer[(*trav_size)++] = node;
}

void gtree_all_partials_gpu(gnode_gpu_t * root, gnode_gpu_t ** travbuffer, unsigned int * trav_size)
{
  *trav_size = 0;
  if (!root->left) return;
  all_partials_recursive(root, trav_size, travbuffer);
}

/* ------------------------------------------------------------------ locus seam */
locus_gpu_t * locus_create_gpu(bppgpu_engine * e, unsigned int dtype, unsigned int model, unsigned int tips,
                               unsigned int clv_buffers, unsigned int states, unsigned int sites,
                               unsigned int rate_matrices, unsigned int prob_matrices, unsigned int rate_cats,
                               unsigned int scale_buffers, unsigned int attributes)
{
  locus_gpu_t * l = (locus_gpu_t *)calloc(1, sizeof(locus_gpu_t));
  l->handle = bppgpu_locus_create(e, dtype, model, tips, clv_buffers, states, sites, rate_matrices, prob_matrices,
                                  rate_cats, scale_buffers, attributes);
  if (!l->handle) { free(l); return NULL; }
  l->tips = tips; l->clv_buffers = clv_buffers; l->states = states; l->sites = sites;
  l->rate_matrices = rate_matrices; l->prob_matrices = prob_matrices; l->rate_cats = rate_cats;
  l->scale_buffers = scale_buffers; l->attributes = attributes; l->model = model; l->dtype = dtype;
  l->cap = 2 * tips;
  l->ops = (bppgpu_partial_op *)malloc(l->cap * sizeof(bppgpu_partial_op));
  l->idx = (unsigned int *)malloc(l->cap * sizeof(unsigned int));
  l->bl = (double *)malloc(l->cap * sizeof(double));
  return l;
}

void locus_destroy_gpu(locus_gpu_t * l)
{
  if (!l) return;
  bppgpu_locus_destroy(l->handle);
  free(l->ops); free(l->idx); free(l->bl);
  free(l);
}

int pll_set_tip_states_gpu(locus_gpu_t * l, unsigned int tip, const unsigned int * map, const char * seq)
{ return bppgpu_set_tip_states(l->handle, tip, map, seq); }
int pll_set_tip_clv_gpu(locus_gpu_t * l, unsigned int tip, const double * clv, int padding)
{ return bppgpu_set_tip_clv(l->handle, tip, clv, padding); }
void pll_set_pattern_weights_gpu(locus_gpu_t * l, const unsigned int * w) { bppgpu_set_pattern_weights(l->handle, w); }
void pll_set_frequencies_gpu(locus_gpu_t * l, unsigned int i, const double * f) { bppgpu_set_frequencies(l->handle, i, f); }
void pll_set_subst_params_gpu(locus_gpu_t * l, unsigned int i, const double * p) { bppgpu_set_subst_params(l->handle, i, p); }
void pll_set_category_rates_gpu(locus_gpu_t * l, const double * r) { bppgpu_set_category_rates(l->handle, r); }

static void ensure_cap(locus_gpu_t * l, unsigned int count)
{
  if (count <= l->cap) return;
  l->cap = count;
  l->ops = (bppgpu_partial_op *)realloc(l->ops, l->cap * sizeof(bppgpu_partial_op));
  l->idx = (unsigned int *)realloc(l->idx, l->cap * sizeof(unsigned int));
  l->bl = (double *)realloc(l->bl, l->cap * sizeof(double));
}

/* node->length as locus_update_matrices sets it before it builds the matrix: strict clock or relaxed clocks
   (gtree_branch_length_gpu) */
static unsigned int fill_matrix_ops(gtree_gpu_t * gtree, gnode_gpu_t ** trav, unsigned int count,
                                    unsigned int * idx, double * bl)
{
  unsigned int i;
  for (i = 0; i < count; ++i)
  {
    gnode_gpu_t * node = trav[i];
    node->length = gtree_branch_length_gpu(gtree, node);
    idx[i] = node->pmatrix_index;
    bl[i] = node->length;
  }
  return count;
}

/* locus.c:2541-2570: node, node->left, node->right resolved to buffer indices */
static unsigned int fill_partial_ops(gnode_gpu_t ** trav, unsigned int count, bppgpu_partial_op * ops)
{
  unsigned int i;
  for (i = 0; i < count; ++i)
  {
    gnode_gpu_t * node = trav[i], * lnode = node->left, * rnode = node->right;
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

void locus_update_matrices_gpu(locus_gpu_t * l, gtree_gpu_t * gtree, gnode_gpu_t ** trav, unsigned int count)
{
  ensure_cap(l, count);
  fill_matrix_ops(gtree, trav, count, l->idx, l->bl);
  bppgpu_update_matrices(l->handle, count, l->idx, l->bl);
}

void locus_update_partials_gpu(locus_gpu_t * l, gnode_gpu_t ** trav, unsigned int count)
{
  ensure_cap(l, count);
  fill_partial_ops(trav, count, l->ops);
  bppgpu_update_partials(l->handle, count, l->ops);
}

double locus_root_loglikelihood_gpu(locus_gpu_t * l, gnode_gpu_t * root, double * persite_lnl)
{
  return bppgpu_root_loglikelihood(l->handle, root->clv_index, root->scaler_index, persite_lnl);
}

/* ------------------------------------------------------------------ batch */
struct locus_batch_gpu_s
{
  bppgpu_batch * handle;
  locus_gpu_t ** loci;
  unsigned int n, mat_cap, op_cap;
  unsigned int * mcounts, * ocounts, * midx, * root_clv;
  int * root_sc;
  double * mbl;
  bppgpu_partial_op * ops;
  gnode_gpu_t ** trav;
  unsigned int trav_cap;
};

locus_batch_gpu_t * locus_batch_create_gpu(bppgpu_engine * e, locus_gpu_t ** loci, unsigned int n)
{
  unsigned int i, mats = 0, ops = 0, maxn = 0;
  locus_batch_gpu_t * b = (locus_batch_gpu_t *)calloc(1, sizeof(*b));
  bppgpu_locus ** h = (bppgpu_locus **)malloc(n * sizeof(*h));
  for (i = 0; i < n; ++i)
  {
    h[i] = loci[i]->handle;
    mats += 2 * loci[i]->tips - 2; ops += loci[i]->tips - 1;
    if (2 * loci[i]->tips > maxn) maxn = 2 * loci[i]->tips;
  }
  b->handle = bppgpu_batch_create(e, n, h);
  free(h);
  if (!b->handle) { free(b); return NULL; }
  b->n = n;
  b->loci = (locus_gpu_t **)malloc(n * sizeof(*b->loci));
  memcpy(b->loci, loci, n * sizeof(*b->loci));
  b->mat_cap = mats; b->op_cap = ops; b->trav_cap = maxn;
  b->mcounts = (unsigned int *)malloc(n * sizeof(unsigned int));
  b->ocounts = (unsigned int *)malloc(n * sizeof(unsigned int));
  /* the arrays that travel every step live in pinned memory: the engine copies them from where they are and
     pipelines the upload of a big full pass against its first wave of kernels (bppgpu_batch_set_waves) */
  b->root_clv = (unsigned int *)bppgpu_host_alloc(n * sizeof(unsigned int));
  b->root_sc = (int *)bppgpu_host_alloc(n * sizeof(int));
  b->midx = (unsigned int *)bppgpu_host_alloc(mats * sizeof(unsigned int));
  b->mbl = (double *)bppgpu_host_alloc(mats * sizeof(double));
  b->ops = (bppgpu_partial_op *)bppgpu_host_alloc(ops * sizeof(bppgpu_partial_op));
  b->trav = (gnode_gpu_t **)malloc(maxn * sizeof(gnode_gpu_t *));
  return b;
}

void locus_batch_destroy_gpu(locus_batch_gpu_t * b)
{
  if (!b) return;
  bppgpu_batch_destroy(b->handle);
  free(b->loci); free(b->mcounts); free(b->ocounts); free(b->trav);
  bppgpu_host_free(b->root_clv); bppgpu_host_free(b->root_sc);
  bppgpu_host_free(b->midx); bppgpu_host_free(b->mbl); bppgpu_host_free(b->ops);
  free(b);
}

double locus_batch_full_pass_gpu(locus_batch_gpu_t * b, gtree_gpu_t ** gtrees, double * logl_out)
{
  unsigned int i, j, k, m = 0, o = 0;
  double sum = 0;
  for (i = 0; i < b->n; ++i)
  {
    gtree_gpu_t * gt = gtrees[i];
    /* all branches (prop_mixing.c:108-117) */
    k = 0;
    for (j = 0; j < gt->tip_count + gt->inner_count; ++j)
      if (gt->nodes[j]->parent) b->trav[k++] = gt->nodes[j];
    fill_matrix_ops(gt, b->trav, k, b->midx + m, b->mbl + m);
    b->mcounts[i] = k; m += k;
    /* all inner nodes in post-order (prop_mixing.c:119-128) */
    gtree_all_partials_gpu(gt->root, b->trav, &k);
    fill_partial_ops(b->trav, k, b->ops + o);
    b->ocounts[i] = k; o += k;
    b->root_clv[i] = gt->root->clv_index;
    b->root_sc[i] = gt->root->scaler_index;
  }
  if (!bppgpu_batch_full_pass(b->handle, b->mcounts, b->midx, b->mbl, b->ocounts, b->ops, b->root_clv, b->root_sc,
                              logl_out, &sum))
    return 0;
  return sum;
}
