// comm.cuh -- the one exchange step of the path: a sum over ranks of a handful of doubles per MCMC step.
//
// Reference (file:line relative to /root/reference/src): after its thread barrier the reference adds the
// per-thread partial results of a move on the main thread -- threads.c:583-590 (mixing: lnacceptance, one
// scalar) and threads.c:544-558 (tau: logl_diff, logpr_diff, count_above, count_below).  With the loci
// sharded over GPUs the same sums run over ranks: one ncclAllReduce(sum) of <= 4 doubles.
//
// Two shapes are supported, both through NCCL:
//   - one process per GPU (torchrun / MPI style): rank 0 makes an id (bppgpu_comm_get_unique_id), the host
//     ships its 128 bytes to the other ranks by any channel it has, every rank calls bppgpu_comm_init_rank;
//   - one process, one engine per GPU (BPP's pthreads, threads.c:234-263): bppgpu_comm_init_all.
// libnccl is loaded with dlopen at the first use, so the library itself has no link-time dependency on it;
// a missing libnccl fails loudly through the fatal handler (no fallback).
#pragma once
#include <dlfcn.h>
#include <nccl.h>

struct NcclApi
{
  void * handle = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char * (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static bool nccl_load()
{
  std::lock_guard<std::mutex> lock(g_nccl_mu);
  if (g_nccl.handle) return true;
  const char * names[] = { getenv("BPPGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
  void * h = nullptr;
  for (const char * nm : names)
    if (nm && *nm && (h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
  if (!h) { fatal("cannot load libnccl (%s): multi-GPU sums need NCCL; set BPPGPU_NCCL_LIB", dlerror()); return false; }
  NcclApi a;
  a.handle = h;
#define BPPGPU_NCCL_SYM(field, name) \
  *(void **)(&a.field) = dlsym(h, name); \
  if (!a.field) { fatal("libnccl has no symbol %s", name); return false; }
  BPPGPU_NCCL_SYM(GetVersion, "ncclGetVersion")
  BPPGPU_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  BPPGPU_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  BPPGPU_NCCL_SYM(CommInitAll, "ncclCommInitAll")
  BPPGPU_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  BPPGPU_NCCL_SYM(AllReduce, "ncclAllReduce")
  BPPGPU_NCCL_SYM(GroupStart, "ncclGroupStart")
  BPPGPU_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  BPPGPU_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef BPPGPU_NCCL_SYM
  g_nccl = a;
  return true;
}

#define NCCL_CHECK(call, fail)                                                                     \
  do {                                                                                             \
    ncclResult_t r__ = (call);                                                                     \
    if (r__ != ncclSuccess)                                                                        \
    { fatal("NCCL error at %s:%d: %s", __FILE__, __LINE__, g_nccl.GetErrorString(r__)); fail; }    \
  } while (0)

struct bppgpu_comm
{
  bppgpu_engine * e = nullptr;
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  double * d_buf = nullptr;            // device scratch for host-side sums
  double * h_buf = nullptr;            // pinned
  cudaStream_t stream = nullptr;       // host-side sums run here (not on a batch stream)
  std::atomic<unsigned long long> calls{0};
};

static constexpr int COMM_MAX_DOUBLES = 256;

static bppgpu_comm * comm_wrap(bppgpu_engine * e, ncclComm_t c, int nranks, int rank)
{
  bppgpu_comm * cm = new bppgpu_comm();
  cm->e = e; cm->comm = c; cm->nranks = nranks; cm->rank = rank;
  cudaSetDevice(e->device);
  if (cudaMalloc(&cm->d_buf, COMM_MAX_DOUBLES * 8) != cudaSuccess ||
      cudaHostAlloc(&cm->h_buf, COMM_MAX_DOUBLES * 8, cudaHostAllocDefault) != cudaSuccess ||
      cudaStreamCreateWithFlags(&cm->stream, cudaStreamNonBlocking) != cudaSuccess)
  {
    fatal("bppgpu_comm: cannot allocate the reduction buffers: %s", cudaGetErrorString(cudaGetLastError()));
    delete cm;
    return nullptr;
  }
  return cm;
}

extern "C" int bppgpu_comm_nccl_version(void)
{
  if (!nccl_load()) return 0;
  int v = 0;
  g_nccl.GetVersion(&v);
  return v;
}

extern "C" int bppgpu_comm_get_unique_id(void * id128)
{
  if (!nccl_load()) return BPPGPU_FAILURE;
  ncclUniqueId id;
  NCCL_CHECK(g_nccl.GetUniqueId(&id), return BPPGPU_FAILURE);
  memcpy(id128, id.internal, NCCL_UNIQUE_ID_BYTES);
  return BPPGPU_SUCCESS;
}

extern "C" bppgpu_comm * bppgpu_comm_init_rank(bppgpu_engine * e, int nranks, int rank, const void * id128)
{
  if (!e || nranks < 1 || rank < 0 || rank >= nranks) { fatal("bppgpu_comm_init_rank: invalid arguments"); return nullptr; }
  if (!nccl_load()) return nullptr;
  CUDA_CHECK(cudaSetDevice(e->device));
  ncclUniqueId id;
  memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
  ncclComm_t c = nullptr;
  NCCL_CHECK(g_nccl.CommInitRank(&c, nranks, id, rank), return nullptr);
  return comm_wrap(e, c, nranks, rank);
}

extern "C" int bppgpu_comm_init_all(bppgpu_engine * const * engines, int n, bppgpu_comm ** comms_out)
{
  if (!engines || n < 1 || !comms_out) { fatal("bppgpu_comm_init_all: invalid arguments"); return BPPGPU_FAILURE; }
  if (!nccl_load()) return BPPGPU_FAILURE;
  std::vector<int> devs(n);
  for (int i = 0; i < n; ++i)
  {
    if (!engines[i]) { fatal("bppgpu_comm_init_all: engine %d is NULL", i); return BPPGPU_FAILURE; }
    devs[i] = engines[i]->device;
    for (int j = 0; j < i; ++j)
      if (devs[j] == devs[i]) { fatal("bppgpu_comm_init_all: engines %d and %d share device %d", j, i, devs[i]); return BPPGPU_FAILURE; }
  }
  std::vector<ncclComm_t> cs(n, nullptr);
  NCCL_CHECK(g_nccl.CommInitAll(cs.data(), n, devs.data()), return BPPGPU_FAILURE);
  for (int i = 0; i < n; ++i)
  {
    comms_out[i] = comm_wrap(engines[i], cs[i], n, i);
    if (!comms_out[i]) return BPPGPU_FAILURE;
  }
  return BPPGPU_SUCCESS;
}

extern "C" void bppgpu_comm_destroy(bppgpu_comm * c)
{
  if (!c) return;
  cudaSetDevice(c->e->device);
  cudaStreamSynchronize(c->stream);
  if (c->comm) g_nccl.CommDestroy(c->comm);
  cudaFree(c->d_buf); cudaFreeHost(c->h_buf); cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" int bppgpu_comm_nranks(const bppgpu_comm * c) { return c->nranks; }
extern "C" int bppgpu_comm_rank(const bppgpu_comm * c) { return c->rank; }
extern "C" unsigned long long bppgpu_comm_calls(const bppgpu_comm * c) { return c->calls.load(); }

// v[0..n) <- sum over ranks, in place, host doubles; returns when v holds the result.  Called by every rank
// (one process per GPU) or by every engine's host thread (one process, engines on pthreads).
extern "C" int bppgpu_allreduce_sum(bppgpu_comm * c, double * v, int n)
{
  if (!c || !v || n < 0 || n > COMM_MAX_DOUBLES) { fatal("bppgpu_allreduce_sum: invalid arguments (n <= %d)", COMM_MAX_DOUBLES); return BPPGPU_FAILURE; }
  if (n == 0) return BPPGPU_SUCCESS;
  CUDA_CHECK(cudaSetDevice(c->e->device));
  memcpy(c->h_buf, v, (size_t)n * 8);
  CUDA_CHECK(cudaMemcpyAsync(c->d_buf, c->h_buf, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
  NCCL_CHECK(g_nccl.AllReduce(c->d_buf, c->d_buf, (size_t)n, ncclDouble, ncclSum, c->comm, c->stream), return BPPGPU_FAILURE);
  CUDA_CHECK(cudaMemcpyAsync(c->h_buf, c->d_buf, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  memcpy(v, c->h_buf, (size_t)n * 8);
  c->calls++;
  return BPPGPU_SUCCESS;
}

// the same for a host that drives all its engines from ONE thread: v[i] belongs to comms[i]
extern "C" int bppgpu_allreduce_sum_all(bppgpu_comm * const * comms, int ncomms, double * const * v, int n)
{
  if (!comms || ncomms < 1 || !v || n < 0 || n > COMM_MAX_DOUBLES) { fatal("bppgpu_allreduce_sum_all: invalid arguments"); return BPPGPU_FAILURE; }
  if (n == 0) return BPPGPU_SUCCESS;
  for (int i = 0; i < ncomms; ++i)
  {
    bppgpu_comm * c = comms[i];
    CUDA_CHECK(cudaSetDevice(c->e->device));
    memcpy(c->h_buf, v[i], (size_t)n * 8);
    CUDA_CHECK(cudaMemcpyAsync(c->d_buf, c->h_buf, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
  }
  NCCL_CHECK(g_nccl.GroupStart(), return BPPGPU_FAILURE);
  for (int i = 0; i < ncomms; ++i)
  {
    bppgpu_comm * c = comms[i];
    NCCL_CHECK(g_nccl.AllReduce(c->d_buf, c->d_buf, (size_t)n, ncclDouble, ncclSum, c->comm, c->stream), return BPPGPU_FAILURE);
  }
  NCCL_CHECK(g_nccl.GroupEnd(), return BPPGPU_FAILURE);
  for (int i = 0; i < ncomms; ++i)
  {
    bppgpu_comm * c = comms[i];
    CUDA_CHECK(cudaSetDevice(c->e->device));
    CUDA_CHECK(cudaMemcpyAsync(c->h_buf, c->d_buf, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
  }
  for (int i = 0; i < ncomms; ++i)
  {
    bppgpu_comm * c = comms[i];
    CUDA_CHECK(cudaSetDevice(c->e->device));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    memcpy(v[i], c->h_buf, (size_t)n * 8);
    c->calls++;
  }
  return BPPGPU_SUCCESS;
}

// Device-side form for the batched step: the batch's lnL sum (one double, bppgpu_batch_lnl_sum_dev) is
// all-reduced in place on the batch's own stream, right behind the kernels of bppgpu_batch_run, with no host
// round trip; bppgpu_batch_collect then returns the GLOBAL sum in lnl_sum_out (the per-locus values stay local).
extern "C" int bppgpu_batch_allreduce_lnl_sum(bppgpu_batch * b, bppgpu_comm * c)
{
  if (!b || !c || b->e != c->e) { fatal("bppgpu_batch_allreduce_lnl_sum: batch and communicator belong to different engines"); return BPPGPU_FAILURE; }
  CUDA_CHECK(cudaSetDevice(b->e->device));
  NCCL_CHECK(g_nccl.AllReduce(b->d_lnl_sum, b->d_lnl_sum, 1, ncclDouble, ncclSum, c->comm, b->stream), return BPPGPU_FAILURE);
  c->calls++;
  return BPPGPU_SUCCESS;
}
