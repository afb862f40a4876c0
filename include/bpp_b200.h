/* bpp_b200.h -- C-ABI of the B200-native Felsenstein-pruning likelihood engine.
 *
 * Drop-in boundary for ONE path of bpp v4.8.7: P-matrix build -> CLV update ->
 * root log-likelihood, i.e. the "locus seam" of the reference
 * (src/bpp.h:2032-2090, implemented in src/locus.c) and the core kernels below it
 * (src/bpp.h:2313-2404; core_pmatrix.c, core_partials*.c, core_likelihood*.c).
 *
 * Plain C: opaque handles, raw pointers and sizes only.  All pointer arguments are
 * HOST pointers unless a name ends in _dev.  All buffers of a locus (CLVs, P-matrices,
 * scalers, packed tip states, weights, model) live in the HBM of the engine's GPU and
 * keep the reference's 2x double-buffered INDEX scheme: the host keeps flipping
 * clv_index / scaler_index / pmatrix_index integers exactly as locus.c:24-26 does and
 * passes indices; nothing is ever copied back unless asked.
 *
 * Conventions kept from the reference:
 *   - BPPGPU_SUCCESS = 1 / BPPGPU_FAILURE = 0          (bpp.h:175-176)
 *   - unrecoverable errors call the fatal handler (default: message to stderr +
 *     exit(1), util.c:30); bppgpu_set_fatal_handler() lets a host install its own.
 *   - no silent CPU fallback: if no sm_100 device / no CUDA driver is present every
 *     compute entry point fails through the fatal handler.
 *   - thread safety: concurrent calls on DISJOINT loci/batches are safe
 *     (threads.c:87-200 calls the seam that way).  Creating or destroying loci, batches and communicators is
 *     set-up work and must not overlap compute calls on the same engine.
 *   - a fatal handler installed with bppgpu_set_fatal_handler should not return (longjmp / exit / throw across a C++
 *     host); if it does return, the failing call reports BPPGPU_FAILURE / NULL where it can, but the objects it
 *     touched must not be used further.
 *   - the raw accessors (bppgpu_get_clv / get_pmatrix / get_scaler / set_pmatrix, bppgpu_set_pattern_weights) wait
 *     for all work in flight on the device before they copy.
 */
#ifndef BPP_B200_H
#define BPP_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BPPGPU_SUCCESS 1
#define BPPGPU_FAILURE 0

/* bpp.h:208-222 */
#define BPPGPU_DATA_DNA        0
#define BPPGPU_DATA_AA         1
#define BPPGPU_DNA_MODEL_JC69  0
#define BPPGPU_DNA_MODEL_K80   1
#define BPPGPU_DNA_MODEL_F81   2
#define BPPGPU_DNA_MODEL_HKY   3
#define BPPGPU_DNA_MODEL_T92   4
#define BPPGPU_DNA_MODEL_TN93  5
#define BPPGPU_DNA_MODEL_F84   6
#define BPPGPU_DNA_MODEL_GTR   7

/* bpp.h:380 PLL_SCALE_BUFFER_NONE */
#define BPPGPU_SCALE_BUFFER_NONE (-1)

/* new arch bit for locus->attributes / opt_arch next to PLL_ATTRIB_ARCH_* (bpp.h:364-369);
   bit 3 (AVX512) is declared but unused there, bit 6 is free. */
#define BPPGPU_ATTRIB_ARCH_CUDA (1u << 6)

/* engine flags */
#define BPPGPU_MATH_EXACT 0u   /* mul/add in the association order of the AVX kernels: 4-state CLVs
                                  bit-identical to --arch avx/avx2 given identical P-matrices */
#define BPPGPU_MATH_FMA   1u   /* fused multiply-add (<= 1 ulp per dot product; lnL within 1e-10) */

typedef struct bppgpu_engine bppgpu_engine;   /* one per GPU */
typedef struct bppgpu_locus  bppgpu_locus;    /* device mirror of locus_t (bpp.h:863-920) */
typedef struct bppgpu_batch  bppgpu_batch;    /* an ordered set of loci launched together */

/* One pruning step; all fields are the reference's buffer indices
   (gnode_t.clv_index / scaler_index / pmatrix_index, bpp.h:715-717) exactly as
   locus_update_partials (locus.c:2530-2571) resolves them for node, node->left, node->right. */
typedef struct bppgpu_partial_op
{
  unsigned int parent_clv_index;
  unsigned int left_clv_index;
  unsigned int right_clv_index;
  unsigned int left_pmatrix_index;    /* pmatrix_index of the LEFT CHILD (locus.c:2565) */
  unsigned int right_pmatrix_index;
  int parent_scaler_index;            /* BPPGPU_SCALE_BUFFER_NONE = no scaling for this node */
  int left_scaler_index;
  int right_scaler_index;
} bppgpu_partial_op;

/* ------------------------------------------------------------------ engine */
int  bppgpu_device_count(void);
bppgpu_engine * bppgpu_engine_create(int device, unsigned int flags);
void bppgpu_engine_destroy(bppgpu_engine * e);
int  bppgpu_engine_device(const bppgpu_engine * e);
void bppgpu_engine_set_math(bppgpu_engine * e, unsigned int math_mode);
void bppgpu_engine_synchronize(bppgpu_engine * e);
void * bppgpu_engine_stream(bppgpu_engine * e);                 /* cudaStream_t of the engine */
unsigned long long bppgpu_engine_launch_count(const bppgpu_engine * e);   /* kernels launched so far */
unsigned long long bppgpu_engine_bytes_allocated(const bppgpu_engine * e);
const char * bppgpu_last_error(void);
void bppgpu_set_fatal_handler(void (*handler)(const char * msg));
const char * bppgpu_version(void);

/* pinned host memory: step arrays (branch lengths, ops, indices ...) that live here are copied to the
   device straight from the caller's buffers, anything else is staged through an internal pinned blob */
void * bppgpu_host_alloc(size_t bytes);
void   bppgpu_host_free(void * p);

/* per-kernel device timers (CUDA events on the launching stream), for bench.py's roofline */
#define BPPGPU_KERNEL_PMATRIX 0
#define BPPGPU_KERNEL_PLAN    1
#define BPPGPU_KERNEL_TREE    2
#define BPPGPU_KERNEL_FINISH  3
#define BPPGPU_KERNEL_COUNT   4
void bppgpu_engine_set_profiling(bppgpu_engine * e, int on);
void bppgpu_engine_get_profile(bppgpu_engine * e, double * ms_by_kernel, unsigned long long * launches_by_kernel);
void bppgpu_engine_reset_profile(bppgpu_engine * e);

/* ------------------------------------------------------------------ locus (bpp.h:2032-2090)
 * replaces locus_create / locus_destroy (locus.c:622-870, 872).  Same arguments, same meaning;
 * BPP calls it with clv_buffers = 2(T-1), prob_matrices = 2(2T-2), scale_buffers = 2(T-1) or 0,
 * rate_matrices = 1 (method.c:4137-4147).
 * Any rate_cats is accepted.  On the device a 4- or 20-state locus holds 1, 2, 4 or 8 categories (3 is kept as 4,
 * 5..7 as 8: the extra ones are copies of category 0 with weight 0, which changes neither a site likelihood nor a
 * scaling decision); every call below takes and returns arrays with the caller's rate_cats.  More than 8
 * categories, or another state count, run the generic kernel. */
bppgpu_locus * bppgpu_locus_create(bppgpu_engine * e,
                                   unsigned int dtype, unsigned int model,
                                   unsigned int tips, unsigned int clv_buffers,
                                   unsigned int states, unsigned int sites,
                                   unsigned int rate_matrices, unsigned int prob_matrices,
                                   unsigned int rate_cats, unsigned int scale_buffers,
                                   unsigned int attributes);
void bppgpu_locus_destroy(bppgpu_locus * l);

/* replaces pll_set_tip_states (locus.c:561; set_tipclv :525): map is the caller's 256-entry
   char->state-mask table (pll_map_nt / pll_map_aa).  Tips are stored PACKED (one mask per site);
   the kernels expand them to the 0/1 doubles the reference stores. */
int  bppgpu_set_tip_states(bppgpu_locus * l, unsigned int tip_index, const unsigned int * map, const char * sequence);
/* replaces pll_set_tip_clv (locus.c:596): `states` doubles per site (padding ignored: states_padded==states) */
int  bppgpu_set_tip_clv(bppgpu_locus * l, unsigned int tip_index, const double * clv, int padding);
/* replaces pll_set_pattern_weights, pll_set_frequencies (:889), pll_set_subst_params (:877),
   pll_set_category_rates.  The model setters only record the new values; the next call that needs them ships
   the changed model blocks of all its loci in one transfer. */
void bppgpu_set_pattern_weights(bppgpu_locus * l, const unsigned int * pattern_weights);
void bppgpu_set_frequencies(bppgpu_locus * l, unsigned int freqs_index, const double * frequencies);
void bppgpu_set_subst_params(bppgpu_locus * l, unsigned int params_index, const double * params);
void bppgpu_set_category_rates(bppgpu_locus * l, const double * rates);
void bppgpu_set_category_weights(bppgpu_locus * l, const double * rate_weights);   /* default 1/R, locus.c:845 */
/* optional: hand over a decomposition computed by the host's own pll_update_eigen
   (core_pmatrix.c:239); otherwise the engine decomposes lazily like locus.c:2462-2476 -- on the device, in one
   kernel for all loci of a batch whose Q changed (create_ratematrix + cyclic Jacobi), on the host for single loci.
   bppgpu_get_eigen returns the decomposition in use (fetched from the device if it was computed there). */
void bppgpu_set_eigen(bppgpu_locus * l, unsigned int params_index,
                      const double * eigenvecs, const double * inv_eigenvecs, const double * eigenvals);
void bppgpu_get_eigen(bppgpu_locus * l, unsigned int params_index,
                      double * eigenvecs, double * inv_eigenvecs, double * eigenvals);

/* replaces locus_update_matrices (locus.c:2417) for `count` branches: P-matrix of the edge above
   a node goes to pmatrix[pmatrix_indices[i]]; branch_lengths[i] = (parent.time-node.time)*rate_mui
   (core_pmatrix.c:711-715) is computed by the caller.  Dispatch by the locus' model like locus.c:2426-2455:
   closed forms for JC69 (locus.c:2325), K80 (:2256), F81 (:2190), HKY / F84 / TN93 (:2068) and T92 (:1981),
   whose parameters are read from subst_params / frequencies exactly as there; eigen form
   (core_pmatrix.c:674) for GTR and the amino-acid models. */
int  bppgpu_update_matrices(bppgpu_locus * l, unsigned int count,
                            const unsigned int * pmatrix_indices, const double * branch_lengths);
/* replaces locus_update_partials (locus.c:2530): ops must be post-ordered */
int  bppgpu_update_partials(bppgpu_locus * l, unsigned int count, const bppgpu_partial_op * ops);
/* replaces locus_root_loglikelihood, haploid branch (locus.c:2618-2630; opt_bfbeta is applied by
   the caller).  persite_lnl may be NULL. */
double bppgpu_root_loglikelihood(bppgpu_locus * l, unsigned int root_clv_index, int root_scaler_index,
                                 double * persite_lnl);
/* replaces pll_core_root_likelihood_vector (core_likelihood.c:214): per-site likelihood, no log,
   no scaler, no weight */
int  bppgpu_root_likelihood_vector(bppgpu_locus * l, unsigned int root_clv_index, double * persite_lh);
/* diploid branch of locus_root_loglikelihood (locus.c:2586-2615) evaluated on the device:
   mean over phase resolutions, log, weight, sum */
int  bppgpu_set_diploid(bppgpu_locus * l, unsigned int unphased_length,
                        const unsigned long * resolution_count, const unsigned long * mapping,
                        unsigned long mapping_length,
                        const unsigned int * unphased_weights);   /* unphased_length weights, method.c:4185-4193 */
double bppgpu_root_loglikelihood_diploid(bppgpu_locus * l, unsigned int root_clv_index);

/* raw buffer access (debug printers output.c:26-96, parity tests, kernel-seam use) */
int  bppgpu_get_clv(bppgpu_locus * l, unsigned int clv_index, double * out);         /* sites*rate_cats*states */
int  bppgpu_get_pmatrix(bppgpu_locus * l, unsigned int pmatrix_index, double * out);  /* rate_cats*states*states */
int  bppgpu_set_pmatrix(bppgpu_locus * l, unsigned int pmatrix_index, const double * in);
int  bppgpu_get_scaler(bppgpu_locus * l, unsigned int scaler_index, unsigned int * out); /* sites */

/* ------------------------------------------------------------------ batch
 * The reference walks `for each locus` on the host (e.g. prop_mixing.c:71-214); a batch is that
 * loop turned into one launch.  Per-locus arrays are concatenated in batch order; counts[i] is
 * the number of entries of locus i. */
bppgpu_batch * bppgpu_batch_create(bppgpu_engine * e, unsigned int n_loci, bppgpu_locus * const * loci);
void bppgpu_batch_destroy(bppgpu_batch * b);
unsigned int bppgpu_batch_size(const bppgpu_batch * b);
const char * bppgpu_batch_kernel_name(bppgpu_batch * b);      /* tree kernel instantiation the batch launches */
/* Diagnostics of the last planned step of a 4-state batch (synchronises the batch): how many loci run which path of
 * the tree kernel.  out[0] = loci on the fast path (stack machine over staged chunks), out[1] = of those, lean
 * one-chunk loci, out[2] = scaled one-chunk loci, out[3] = loci on the cell-at-a-time walker, out[4] = largest
 * chunk count, out[5] = stack slots per cell, out[6] = cells per thread, out[7] = shared memory of the launch (bytes).
 * Returns BPPGPU_FAILURE for other batch kinds or before the first run. */
int bppgpu_batch_plan_stats(bppgpu_batch * b, unsigned int out[8]);

int  bppgpu_batch_update_matrices(bppgpu_batch * b, const unsigned int * counts,
                                  const unsigned int * pmatrix_indices, const double * branch_lengths);
int  bppgpu_batch_update_partials(bppgpu_batch * b, const unsigned int * counts, const bppgpu_partial_op * ops);
int  bppgpu_batch_root_loglikelihood(bppgpu_batch * b, const unsigned int * root_clv_indices,
                                     const int * root_scaler_indices, double * lnl_out);
/* update_matrices + update_partials + root_loglikelihood of every locus of the batch in one call:
   one H2D copy of the step's inputs, P-matrix kernel, tree kernel with the root lnL fused, one
   D2H copy of n_loci doubles.  lnl_sum_out (may be NULL) receives the fixed-order sum over loci. */
int  bppgpu_batch_full_pass(bppgpu_batch * b,
                            const unsigned int * matrix_counts, const unsigned int * pmatrix_indices,
                            const double * branch_lengths,
                            const unsigned int * op_counts, const bppgpu_partial_op * ops,
                            const unsigned int * root_clv_indices, const int * root_scaler_indices,
                            double * lnl_out, double * lnl_sum_out);
/* the same step split so that a caller (bench.py `value`) can keep the inputs resident in HBM:
   stage = H2D only, run = kernels only (asynchronous on the batch stream), collect = D2H + sync */
int  bppgpu_batch_stage(bppgpu_batch * b,
                        const unsigned int * matrix_counts, const unsigned int * pmatrix_indices,
                        const double * branch_lengths,
                        const unsigned int * op_counts, const bppgpu_partial_op * ops,
                        const unsigned int * root_clv_indices, const int * root_scaler_indices);
int  bppgpu_batch_run(bppgpu_batch * b);
/* A full pass of a big 4-state batch is pipelined: the step's arrays are uploaded in `waves` slices of loci
   on a copy stream while planner and tree kernel of earlier slices already run (0 = automatic: 2 waves of
   20 % + 80 % of the loci once the step's arrays exceed 1 MB, else 1; 1 = off; at most 8).  Results do not depend
   on the setting.
   LIFETIME OF THE STEP ARRAYS: arrays that live in pinned memory (bppgpu_host_alloc) are read by the copy engine
   straight from the caller's buffers, asynchronously: they must stay valid and unchanged until
   bppgpu_batch_wait_inputs, bppgpu_batch_collect or bppgpu_batch_synchronize has returned after the
   bppgpu_batch_run that consumes them (run itself never blocks).  Pageable arrays are copied into the batch's own
   pinned blob before stage returns and may be reused at once. */
void bppgpu_batch_set_waves(bppgpu_batch * b, unsigned int waves);
int  bppgpu_batch_collect(bppgpu_batch * b, double * lnl_out, double * lnl_sum_out);
int  bppgpu_batch_wait_inputs(bppgpu_batch * b);   /* returns when the device has read the step's host arrays */
/* Whole-tree proposals (the mixing move, prop_mixing.c:52-220; debug_full_lh, method.c:4660-4696) keep their
   traversals and flip EVERY index: SWAP_PMAT_INDEX of all edges, SWAP_CLV_INDEX / SWAP_SCALER_INDEX of all inner
   nodes (locus.c:24-26).  bppgpu_batch_flip_indices applies exactly those flips to the staged step ON THE DEVICE
   (ops, matrix list, roots), and the batch keeps the planned program of both index parities; with
   bppgpu_batch_set_branch_lengths (same order and count as the staged matrix list) such a step uploads branch
   lengths only and runs without planning.  A rejected proposal is flip_indices again (nothing is recomputed: the
   other half of the double buffers still holds the accepted state).  Requires BPP's 2x allocation
   (clv_buffers = 2(T-1), prob_matrices = 2(2T-2), scale_buffers = 2(T-1) or 0). */
int  bppgpu_batch_flip_indices(bppgpu_batch * b);
int  bppgpu_batch_set_branch_lengths(bppgpu_batch * b, const double * branch_lengths);
/* device address of the batch's lnL sum (one double), valid after run; lets the caller hand it to
   an all-reduce (torch.distributed / NCCL) without a host round trip */
void * bppgpu_batch_lnl_sum_dev(bppgpu_batch * b);
void * bppgpu_batch_stream(bppgpu_batch * b);
/* device timers around a region of the batch stream */
void   bppgpu_batch_timer_start(bppgpu_batch * b);
double bppgpu_batch_timer_stop_ms(bppgpu_batch * b);          /* synchronizes the stream */
void   bppgpu_batch_synchronize(bppgpu_batch * b);

/* ------------------------------------------------------------------ multi-GPU: the path's one exchange step
 * Loci are sharded over GPUs; after a batched move the reference adds per-thread partial results on the main
 * thread (threads.c:583-590 mixing: lnacceptance; threads.c:544-558 tau: logl_diff, logpr_diff, count_above,
 * count_below).  Over GPUs that is one NCCL all-reduce(sum) of <= 4 doubles.  libnccl.so.2 is loaded with
 * dlopen at the first call (override with $BPPGPU_NCCL_LIB); without it these calls fail through the fatal
 * handler. */
typedef struct bppgpu_comm bppgpu_comm;       /* one per (engine, communicator) */
#define BPPGPU_COMM_ID_BYTES 128
int  bppgpu_comm_nccl_version(void);
/* one process per GPU: rank 0 makes the id, the host ships its 128 bytes to every rank, all call init_rank */
int  bppgpu_comm_get_unique_id(void * id128);
bppgpu_comm * bppgpu_comm_init_rank(bppgpu_engine * e, int nranks, int rank, const void * id128);
/* one process, n engines on n devices (BPP's pthreads over loci ranges, threads.c:234-263) */
int  bppgpu_comm_init_all(bppgpu_engine * const * engines, int n, bppgpu_comm ** comms_out);
void bppgpu_comm_destroy(bppgpu_comm * c);
int  bppgpu_comm_nranks(const bppgpu_comm * c);
int  bppgpu_comm_rank(const bppgpu_comm * c);
unsigned long long bppgpu_comm_calls(const bppgpu_comm * c);       /* all-reduces issued so far */
/* v[0..n) <- sum over ranks (host doubles, in place, n <= 256); every rank / every engine's thread calls it */
int  bppgpu_allreduce_sum(bppgpu_comm * c, double * v, int n);
/* the same when ONE host thread drives all engines of the process: v[i] belongs to comms[i] */
int  bppgpu_allreduce_sum_all(bppgpu_comm * const * comms, int ncomms, double * const * v, int n);
/* device-side: all-reduce the batch's lnL sum in place on the batch stream behind bppgpu_batch_run (no host
   round trip); bppgpu_batch_collect then returns the global sum in lnl_sum_out */
int  bppgpu_batch_allreduce_lnl_sum(bppgpu_batch * b, bppgpu_comm * c);

#ifdef __cplusplus
}
#endif
#endif /* BPP_B200_H */
