/* oracle/ref_hook.c -- TEST INFRASTRUCTURE ONLY (fixture generation in the build container).
 *
 * Linked into oracle/_ref/libbppref_hook.so, in which the reference's method.c is compiled with
 *   -Dlocus_root_loglikelihood=ref_hook_root_lnl
 * so that the call at method.c:4297 (init(): first full-tree lnL of every locus, the "log-L0" line)
 * lands here.  The hook dumps everything the likelihood path saw for that locus -- REAL data after
 * BPP's own parsing, site-pattern compression and diploid phase resolution -- to the file named by
 * $BPP_HOOK_DUMP, calls the real function, records its result, and exits after the last locus.
 * Nothing of the reference is modified or copied.
 */
#include "bpp.h"

#undef locus_root_loglikelihood
double locus_root_loglikelihood(locus_t * locus, gnode_t * root, const unsigned int * freqs_indices, double * persite_lnl);

static FILE * fp = NULL;
static long dumped = 0;

static void wr(const void * p, size_t n) { fwrite(p, 1, n, fp); }
static void wi(long v) { wr(&v, sizeof(long)); }
static void wd(double v) { wr(&v, sizeof(double)); }

static long count_nodes(gnode_t * x) { return x->left ? 1 + count_nodes(x->left) + count_nodes(x->right) : 1; }

static void dump_node(gnode_t * x)
{
  wi(x->node_index); wi(x->left ? x->left->node_index : -1); wi(x->right ? x->right->node_index : -1);
  wi(x->parent ? x->parent->node_index : -1); wi(x->clv_index); wi(x->scaler_index); wi(x->pmatrix_index);
  wd(x->time); wd(x->length);
  if (x->left) { dump_node(x->left); dump_node(x->right); }
}

double ref_hook_root_lnl(locus_t * locus, gnode_t * root, const unsigned int * freqs_indices, double * persite_lnl)
{
  unsigned int i;
  double logl;
  if (!fp)
  {
    const char * path = getenv("BPP_HOOK_DUMP");
    fp = fopen(path ? path : "bpp_hook_dump.bin", "wb");
    wi(opt_locus_count);
  }
  logl = locus_root_loglikelihood(locus, root, freqs_indices, persite_lnl);
  wi(locus->tips); wi(locus->sites); wi(locus->states); wi(locus->rate_cats); wi(locus->model); wi(locus->dtype);
  wi(locus->diploid); wi(locus->diploid ? locus->unphased_length : 0); wi(locus->scale_buffers);
  wd(logl); wd(opt_bfbeta);
  wr(locus->frequencies[0], locus->states * sizeof(double));
  wr(locus->rates, locus->rate_cats * sizeof(double));
  wr(locus->rate_weights, locus->rate_cats * sizeof(double));
  /* tip CLVs as the reference stores them (0/1 doubles, first category) */
  for (i = 0; i < locus->tips; ++i)
  {
    unsigned int s;
    for (s = 0; s < locus->sites; ++s)
      wr(locus->clv[i] + (size_t)s * locus->states * locus->rate_cats, locus->states * sizeof(double));
  }
  if (locus->diploid)
  {
    long maplen = 0;
    for (i = 0; i < (unsigned int)locus->unphased_length; ++i) maplen += locus->diploid_resolution_count[i];
    wr(locus->pattern_weights, locus->unphased_length * sizeof(unsigned int));
    wr(locus->diploid_resolution_count, locus->unphased_length * sizeof(unsigned long));
    wi(maplen);
    wr(locus->diploid_mapping, maplen * sizeof(unsigned long));
    wr(locus->likelihood_vector, locus->sites * sizeof(double));
  }
  else wr(locus->pattern_weights, locus->sites * sizeof(unsigned int));
  wi(count_nodes(root));
  dump_node(root);
  /* root CLV for a direct check */
  wr(locus->clv[root->clv_index], (size_t)locus->sites * locus->states * locus->rate_cats * sizeof(double));
  if (++dumped == opt_locus_count) { fclose(fp); exit(0); }
  return logl;
}
