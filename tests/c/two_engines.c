/* two_engines.c -- GPU test program (built and run by tests/test_comm.py).
 *
 * The natural multi-GPU shape for BPP itself: ONE process, one engine per GPU, one host pthread per engine
 * working on a contiguous range of loci (load_balance_none, /root/reference/src/threads.c:234-263), and after
 * the threads are done the per-thread partial results are summed (threads.c:583-590 mixing: one scalar;
 * threads.c:544-558 tau: four scalars) -- here by bppgpu_allreduce_sum over NCCL.
 *
 * usage: two_engines <n_engines>
 *   n_engines = 1: one engine, communicator of one rank (bppgpu_comm_init_rank, the per-process shape)
 *   n_engines = 2: devices 0 and 1, bppgpu_comm_init_all, one pthread per engine
 * Both are compared with the fixed-order sum of all loci computed on engine 0 alone.  Prints "OK <sum>" on success.
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bpp_gpu_host.h"

#define N_LOCI 37
#define TIPS 6
#define SITES 211
#define CATS 4

static unsigned long long rng_state = 20261017ULL;
static unsigned int rnd(void)
{
  rng_state = rng_state * 6364136223846793005ULL + 1442695040888963407ULL;
  return (unsigned int)(rng_state >> 33);
}
static double rndu(void) { return (rnd() + 0.5) / 2147483648.0; }

static unsigned int nt_map[256];

typedef struct { char seq[TIPS][SITES + 1]; unsigned int w[SITES]; int left[TIPS - 1], right[TIPS - 1];
                 double times[2 * TIPS - 1], freqs[4], subst[6]; } locus_data;
static locus_data data[N_LOCI];

static void make_data(void)
{
  int i, t, s, k;
  memset(nt_map, 0, sizeof(nt_map));
  nt_map['A'] = 1; nt_map['C'] = 2; nt_map['G'] = 4; nt_map['T'] = 8; nt_map['N'] = 15;
  for (i = 0; i < N_LOCI; ++i)
  {
    locus_data * d = data + i;
    double age = 0, fs = 0;
    int active[TIPS], m = TIPS;
    for (t = 0; t < TIPS; ++t)
    {
      for (s = 0; s < SITES; ++s) d->seq[t][s] = (rnd() % 50 == 0) ? 'N' : "ACGT"[rnd() & 3];
      d->seq[t][SITES] = 0;
      active[t] = t;
    }
    for (s = 0; s < SITES; ++s) d->w[s] = 1 + rnd() % 5;
    memset(d->times, 0, sizeof(d->times));
    for (k = 0; k < TIPS - 1; ++k)            /* random joins, ages increasing towards the root */
    {
      int a = rnd() % m, b = rnd() % (m - 1);
      if (b >= a) ++b;
      d->left[k] = active[a]; d->right[k] = active[b];
      age += 0.001 + 0.03 * rndu();
      d->times[TIPS + k] = age;
      if (a > b) { int x = a; a = b; b = x; }
      active[a] = TIPS + k; active[b] = active[m - 1]; --m;
    }
    for (k = 0; k < 4; ++k) { d->freqs[k] = 0.8 + 0.4 * rndu(); fs += d->freqs[k]; }
    for (k = 0; k < 4; ++k) d->freqs[k] /= fs;
    for (k = 0; k < 6; ++k) d->subst[k] = 0.5 + rndu();
    d->subst[5] = 1.0;
  }
}

typedef struct
{
  bppgpu_engine * e;
  bppgpu_comm * comm;
  int first, count;
  locus_gpu_t * loci[N_LOCI];
  gtree_gpu_t * trees[N_LOCI];
  locus_batch_gpu_t * batch;
  double lnl[N_LOCI];
  double local, global, tau[4];
} shard_t;

static int shard_setup(shard_t * sh, int device, int first, int count)
{
  int i, t;
  double rates[CATS];
  sh->e = bppgpu_engine_create(device, BPPGPU_MATH_EXACT);
  if (!sh->e) return 0;
  sh->first = first; sh->count = count;
  bppgpu_compute_gamma_cats(0.5, 0.5, CATS, rates);
  for (i = 0; i < count; ++i)
  {
    locus_data * d = data + first + i;
    locus_gpu_t * l = locus_create_gpu(sh->e, BPPGPU_DATA_DNA, BPPGPU_DNA_MODEL_GTR, TIPS, 2 * (TIPS - 1), 4, SITES, 1,
                                       2 * (2 * TIPS - 2), CATS, 2 * (TIPS - 1), BPPGPU_ATTRIB_ARCH_CUDA);
    if (!l) return 0;
    for (t = 0; t < TIPS; ++t) if (!pll_set_tip_states_gpu(l, t, nt_map, d->seq[t])) return 0;
    pll_set_pattern_weights_gpu(l, d->w);
    pll_set_frequencies_gpu(l, 0, d->freqs);
    pll_set_subst_params_gpu(l, 0, d->subst);
    pll_set_category_rates_gpu(l, rates);
    sh->loci[i] = l;
    sh->trees[i] = gtree_create_gpu(TIPS, d->left, d->right, d->times, 1.0, 1);
  }
  sh->batch = locus_batch_create_gpu(sh->e, sh->loci, count);
  return sh->batch != NULL;
}

static void shard_free(shard_t * sh)
{
  int i;
  if (sh->comm) bppgpu_comm_destroy(sh->comm);
  locus_batch_destroy_gpu(sh->batch);
  for (i = 0; i < sh->count; ++i) { locus_destroy_gpu(sh->loci[i]); gtree_destroy_gpu(sh->trees[i]); }
  bppgpu_engine_destroy(sh->e);
}

static void * shard_run(void * arg)
{
  shard_t * sh = (shard_t *)arg;
  int k;
  sh->local = locus_batch_full_pass_gpu(sh->batch, sh->trees, sh->lnl);
  sh->global = sh->local;
  /* mixing: one scalar (threads.c:583-590) */
  if (!bppgpu_allreduce_sum(sh->comm, &sh->global, 1)) sh->global = NAN;
  /* tau: four scalars (threads.c:544-558) */
  for (k = 0; k < 4; ++k) sh->tau[k] = (k + 1) * sh->local;
  if (!bppgpu_allreduce_sum(sh->comm, sh->tau, 4)) sh->tau[0] = NAN;
  return NULL;
}

int main(int argc, char ** argv)
{
  int ne = argc > 1 ? atoi(argv[1]) : 1, i, k;
  static shard_t whole, part[2];
  double expect;
  make_data();
  if (ne < 1 || ne > 2) { fprintf(stderr, "n_engines must be 1 or 2\n"); return 2; }
  if (bppgpu_device_count() < ne) { printf("SKIP need %d devices, have %d\n", ne, bppgpu_device_count()); return 0; }

  /* the answer: all loci on engine 0, fixed-order sum */
  memset(&whole, 0, sizeof(whole));
  if (!shard_setup(&whole, 0, 0, N_LOCI)) { fprintf(stderr, "setup failed: %s\n", bppgpu_last_error()); return 1; }
  expect = locus_batch_full_pass_gpu(whole.batch, whole.trees, whole.lnl);

  if (ne == 1)
  {
    unsigned char id[BPPGPU_COMM_ID_BYTES];
    if (!bppgpu_comm_get_unique_id(id)) return 1;
    whole.comm = bppgpu_comm_init_rank(whole.e, 1, 0, id);
    if (!whole.comm) return 1;
    shard_run(&whole);
    if (whole.global != expect || whole.tau[3] != 4 * expect) { fprintf(stderr, "1-rank sum differs\n"); return 1; }
    if (bppgpu_comm_calls(whole.comm) != 2) return 1;
    printf("nccl %d\n", bppgpu_comm_nccl_version());
  }
  else
  {
    pthread_t th[2];
    bppgpu_engine * engines[2];
    bppgpu_comm * comms[2];
    double * vv[2], v0[2], v1[2];
    for (i = 0; i < 2; ++i)
    {
      /* load_balance_none, threads.c:234-263 */
      const int per = N_LOCI / 2, rem = N_LOCI % 2;
      const int first = i * per + (i < rem ? i : rem), count = per + (i < rem ? 1 : 0);
      memset(&part[i], 0, sizeof(part[i]));
      if (!shard_setup(&part[i], i, first, count)) { fprintf(stderr, "setup failed: %s\n", bppgpu_last_error()); return 1; }
      engines[i] = part[i].e;
    }
    if (!bppgpu_comm_init_all(engines, 2, comms)) return 1;
    part[0].comm = comms[0]; part[1].comm = comms[1];
    for (i = 0; i < 2; ++i) pthread_create(&th[i], NULL, shard_run, &part[i]);
    for (i = 0; i < 2; ++i) pthread_join(th[i], NULL);
    for (i = 0; i < 2; ++i)
    {
      if (part[i].global != part[0].global) { fprintf(stderr, "ranks disagree\n"); return 1; }
      if (fabs(part[i].global - expect) > 1e-12 * fabs(expect)) { fprintf(stderr, "sharded sum %.17g != %.17g\n", part[i].global, expect); return 1; }
      for (k = 0; k < 4; ++k)
        if (fabs(part[i].tau[k] - (k + 1) * expect) > 1e-12 * fabs(expect)) { fprintf(stderr, "tau sum differs\n"); return 1; }
      for (k = 0; k < part[i].count; ++k)
        if (part[i].lnl[k] != whole.lnl[part[i].first + k]) { fprintf(stderr, "locus %d differs between devices\n", part[i].first + k); return 1; }
    }
    /* the single-thread form: one host thread drives both engines */
    v0[0] = part[0].local; v0[1] = 1; v1[0] = part[1].local; v1[1] = 2;
    vv[0] = v0; vv[1] = v1;
    if (!bppgpu_allreduce_sum_all(comms, 2, vv, 2)) return 1;
    if (v0[0] != v1[0] || v0[1] != 3 || v1[1] != 3 || fabs(v0[0] - expect) > 1e-12 * fabs(expect)) { fprintf(stderr, "sum_all differs\n"); return 1; }
    for (i = 0; i < 2; ++i) shard_free(&part[i]);
  }
  shard_free(&whole);
  printf("OK %.17g\n", expect);
  return 0;
}
