/* bpp_gpu_host.h -- C host-side mirror of the reference's locus seam over the C-ABI.
 *
 * Same function names (with a _gpu suffix), same argument meaning and the same error behaviour as
 * bpp v4.8.7 src/bpp.h:2032-2090 / src/locus.c, so that the call sites of the reference
 * (method.c:4137-4297, prop_mixing.c:52-220, gtree.c:5437-5467 ...) translate line by line; see
 * INTEGRATION.md for the hooks inside BPP itself.  gnode_gpu_t / gtree_gpu_t carry exactly the
 * fields of gnode_t / gtree_t (bpp.h:692-757) that the seam reads.
 */
#ifndef BPP_GPU_HOST_H
#define BPP_GPU_HOST_H

#include "bpp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gnode_gpu_s
{
  struct gnode_gpu_s * left;
  struct gnode_gpu_s * right;
  struct gnode_gpu_s * parent;
  double length;                /* written by locus_update_matrices like node->length (core_pmatrix.c:714) */
  double time;
  unsigned int node_index;
  unsigned int clv_index;
  int scaler_index;
  unsigned int pmatrix_index;
  int pop;                      /* species-tree node the gene node sits in (gnode_t::pop); relaxed clocks only */
} gnode_gpu_t;

/* the slice of stree_t / snode_t the relaxed-clock branch lengths read (locus.c:1105-1193): parent, tau and, for the
   locus at hand, the branch rate of every species node (snode_t::brate[msa_index]); no hybridisation nodes */
typedef struct stree_gpu_s
{
  unsigned int node_count;
  const int * parent;           /* parent species node or -1 for the root */
  const double * tau;
  const double * brate;
} stree_gpu_t;

typedef struct gtree_gpu_s
{
  unsigned int tip_count, inner_count, edge_count;
  gnode_gpu_t ** nodes;         /* tips first, then inner nodes; nodes[i]->node_index == i */
  gnode_gpu_t * root;
  double rate_mui;
  double logl;
  const stree_gpu_t * stree;    /* NULL = strict clock (opt_clock == BPP_CLOCK_GLOBAL) */
  double rate_scale;            /* relaxed clocks: 1, or the locus rate for BPP_CLOCK_SIMPLE (brate[0] * locusrate) */
} gtree_gpu_t;

typedef struct locus_gpu_s
{
  bppgpu_locus * handle;
  unsigned int tips, clv_buffers, states, sites, rate_matrices, prob_matrices, rate_cats, scale_buffers;
  unsigned int attributes, model, dtype;
  /* scratch for building op lists */
  bppgpu_partial_op * ops;
  unsigned int * idx;
  double * bl;
  unsigned int cap;
} locus_gpu_t;

/* index flips, locus.c:24-26 */
#define SWAP_CLV_INDEX_GPU(n,i)    ((n)+((i)-1)%(2*(n)-2))
#define SWAP_SCALER_INDEX_GPU(n,i) (((n)+((i)-1))%(2*(n)-2))
#define SWAP_PMAT_INDEX_GPU(e,i)   (((e)+(i))%((e)<<1))

/* gene-tree scaffolding with the reference's initial index assignment (gtree.c:2395-2399,2664-2675):
   left/right are node ids of inner node k = tips+k, children before parents */
gtree_gpu_t * gtree_create_gpu(unsigned int tips, const int * left, const int * right, const double * times,
                               double rate_mui, int scaling);
void gtree_destroy_gpu(gtree_gpu_t * t);
/* relaxed clock: pops[k] = species node of gene node k (gnode_t::pop); pass stree = NULL to go back to the strict clock */
void gtree_set_relaxed_clock_gpu(gtree_gpu_t * t, const stree_gpu_t * stree, const int * pops, double rate_scale);
/* length of the branch above `node` as locus_update_matrices computes it: (parent time - time) * rate_mui under the
   strict clock (locus.c:2347-2351), the rate-weighted sum over the species-tree branches it crosses under a relaxed
   clock (update_branchlength_relaxed_clock{,_simple}, locus.c:1105-1193) */
double gtree_branch_length_gpu(const gtree_gpu_t * t, const gnode_gpu_t * node);
/* recursive left,right,node traversal (prop_mixing.c:28-50) */
void gtree_all_partials_gpu(gnode_gpu_t * root, gnode_gpu_t ** travbuffer, unsigned int * trav_size);

/* locus seam */
locus_gpu_t * locus_create_gpu(bppgpu_engine * e, unsigned int dtype, unsigned int model, unsigned int tips,
                               unsigned int clv_buffers, unsigned int states, unsigned int sites,
                               unsigned int rate_matrices, unsigned int prob_matrices, unsigned int rate_cats,
                               unsigned int scale_buffers, unsigned int attributes);
void locus_destroy_gpu(locus_gpu_t * locus);
int  pll_set_tip_states_gpu(locus_gpu_t * locus, unsigned int tip_index, const unsigned int * map, const char * sequence);
int  pll_set_tip_clv_gpu(locus_gpu_t * locus, unsigned int tip_index, const double * clv, int padding);
void pll_set_pattern_weights_gpu(locus_gpu_t * locus, const unsigned int * pattern_weights);
void pll_set_frequencies_gpu(locus_gpu_t * locus, unsigned int freqs_index, const double * frequencies);
void pll_set_subst_params_gpu(locus_gpu_t * locus, unsigned int params_index, const double * params);
void pll_set_category_rates_gpu(locus_gpu_t * locus, const double * rates);
void locus_update_matrices_gpu(locus_gpu_t * locus, gtree_gpu_t * gtree, gnode_gpu_t ** traversal, unsigned int count);
void locus_update_partials_gpu(locus_gpu_t * locus, gnode_gpu_t ** traversal, unsigned int count);
double locus_root_loglikelihood_gpu(locus_gpu_t * locus, gnode_gpu_t * root, double * persite_lnl);

/* the callers' `for each locus` loop as one launch: full-tree pass of loci [0, n) the way
   prop_mixing_update_gtrees does it per locus (all 2T-2 P-matrices, all inner CLVs in post-order,
   root lnL); logl_out[i] receives locus i's lnL, the fixed-order sum is returned */
typedef struct locus_batch_gpu_s locus_batch_gpu_t;
locus_batch_gpu_t * locus_batch_create_gpu(bppgpu_engine * e, locus_gpu_t ** loci, unsigned int n);
void locus_batch_destroy_gpu(locus_batch_gpu_t * b);
double locus_batch_full_pass_gpu(locus_batch_gpu_t * b, gtree_gpu_t ** gtrees, double * logl_out);

/* pll_compute_gamma_cats, mean method (gamma.c:221-284) */
int bppgpu_compute_gamma_cats(double alpha, double beta, unsigned int categories, double * output_rates);

#ifdef __cplusplus
}
#endif
#endif
