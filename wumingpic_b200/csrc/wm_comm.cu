// wm_comm.cu -- inter-GPU transport: the NCCL replacement of the reference's MPI_SENDRECV /
// MPI_ALLREDUCE call sites.  One process per GPU; the communicator is created from
// a ncclUniqueId the host broadcasts (MPI_Bcast in the Fortran driver, torch.distributed in bench.py).
//
// NCCL is resolved at run time with dlopen so that (a) single-GPU use needs no NCCL at all and
// (b) inside a Python process that already loaded torch's bundled libnccl.so.2 the same library
// instance is reused instead of a second copy.
#include "wm_internal.cuh"

#include <dlfcn.h>
#include <cstring>

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi& api() {
  static NcclApi a;
  if (a.lib) return a;
  const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; names[i] && !a.lib; ++i) a.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!a.lib) return a;
#define LOAD(field, sym) *(void**)(&a.field) = dlsym(a.lib, sym)
  LOAD(GetUniqueId, "ncclGetUniqueId");
  LOAD(CommInitRank, "ncclCommInitRank");
  LOAD(CommDestroy, "ncclCommDestroy");
  LOAD(Send, "ncclSend");
  LOAD(Recv, "ncclRecv");
  LOAD(AllReduce, "ncclAllReduce");
  LOAD(GroupStart, "ncclGroupStart");
  LOAD(GroupEnd, "ncclGroupEnd");
  LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
  a.ok = a.GetUniqueId && a.CommInitRank && a.Send && a.Recv && a.AllReduce && a.GroupStart && a.GroupEnd;
  return a;
}

#define WM_NCCL(call)                                                                          \
  do {                                                                                         \
    ncclResult_t r__ = (call);                                                                 \
    if (r__ != 0) {                                                                            \
      wm_set_error(std::string(#call) + ": " + (api().GetErrorString ? api().GetErrorString(r__) : "nccl error")); \
      return WM_ERR_CUDA;                                                                      \
    }                                                                                          \
  } while (0)

}  // namespace

extern "C" int wm_comm_unique_id(char* id_bytes128) {
  NcclApi& a = api();
  if (!a.ok) {
    wm_set_error("NCCL library not found (libnccl.so.2)");
    return WM_ERR_CUDA;
  }
  ncclUniqueId id;
  WM_NCCL(a.GetUniqueId(&id));
  std::memcpy(id_bytes128, id.internal, 128);
  return WM_OK;
}

extern "C" int wm_comm_init(wm_ctx* ctx, int nranks, int rank, const char* id_bytes128) {
  if (!ctx) return WM_ERR_ARG;
  if (nranks != ctx->prm.nproc_j * ctx->prm.nproc_k) {
    wm_set_error("error in proc no.");  // mpi_set.f90:34-43
    return WM_ERR_ARG;
  }
  if (rank != ctx->prm.rank_j * ctx->prm.nproc_k + ctx->prm.rank_k) {
    wm_set_error("rank does not match rank_j*nproc_k + rank_k");
    return WM_ERR_ARG;
  }
  if (nranks > 1 && ((ctx->g.dim == 3 && ctx->prm.nproc_j != 1) || (ctx->g.dim == 2 && ctx->prm.nproc_k != 1))) {
    wm_set_error("decompose along the last axis only (z-slabs in 3-D: nproc_j = 1; y-slabs in 2-D): on NVSwitch every "
                 "GPU pair has full bandwidth, so 1-D slabs replace the reference's 2-D rank grid");
    return WM_ERR_ARG;
  }
  ctx->nranks = nranks;
  ctx->rank = rank;
  if (nranks == 1) return WM_OK;
  NcclApi& a = api();
  if (!a.ok) {
    wm_set_error("NCCL library not found (libnccl.so.2)");
    return WM_ERR_CUDA;
  }
  ncclUniqueId id;
  std::memcpy(id.internal, id_bytes128, 128);
  ncclComm_t comm;
  WM_CUDA(cudaSetDevice(ctx->device));
  WM_NCCL(a.CommInitRank(&comm, nranks, id, rank));
  ctx->nccl_comm = comm;
  return wm_enable_slab_migration(ctx);
}

int wm_comm_destroy(wm_ctx* ctx) {
  if (ctx->nccl_comm && api().CommDestroy) api().CommDestroy((ncclComm_t)ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
  return WM_OK;
}

// One MPI_SENDRECV: send `snd` towards the down (dir_down=1) or up neighbour along `axis`
// (1 = y, 2 = z) and receive the matching message from the opposite neighbour into `rcv`.
// When the neighbour is this rank nothing moves: the caller unpacks straight from `snd`.
int wm_comm_sendrecv(wm_ctx* ctx, int axis, int dir_down, const double* snd, double* rcv, size_t n) {
  const int to = dir_down ? ctx->rank_down[axis - 1] : ctx->rank_up[axis - 1];
  const int from = dir_down ? ctx->rank_up[axis - 1] : ctx->rank_down[axis - 1];
  if (to == ctx->rank && from == ctx->rank) return WM_OK;
  if (!ctx->nccl_comm) {
    wm_set_error("multi-rank exchange requested but wm_comm_init was not called");
    return WM_ERR_ARG;
  }
  NcclApi& a = api();
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  WM_NCCL(a.GroupStart());
  WM_NCCL(a.Send(snd, n, ncclFloat64, to, comm, ctx->stream));
  WM_NCCL(a.Recv(rcv, n, ncclFloat64, from, comm, ctx->stream));
  WM_NCCL(a.GroupEnd());
  return WM_OK;
}

// Grouped point-to-point messages (byte counts) for the migration phases of wm_sort.cu: everything between
// begin and end is one ncclGroup, i.e. one fused transfer kernel over NVLink.
int wm_comm_group_begin(wm_ctx* ctx) {
  if (!ctx->nccl_comm) {
    wm_set_error("multi-rank exchange requested but wm_comm_init was not called");
    return WM_ERR_ARG;
  }
  WM_NCCL(api().GroupStart());
  return WM_OK;
}
int wm_comm_group_end(wm_ctx* ctx) {
  WM_NCCL(api().GroupEnd());
  ctx->launches++;
  return WM_OK;
}
int wm_comm_send(wm_ctx* ctx, int peer, const void* buf, size_t bytes) {
  WM_NCCL(api().Send(buf, bytes, ncclInt8, peer, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  return WM_OK;
}
int wm_comm_recv(wm_ctx* ctx, int peer, void* buf, size_t bytes) {
  WM_NCCL(api().Recv(buf, bytes, ncclInt8, peer, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  return WM_OK;
}

// MPI_ALLREDUCE(..., MPI_SUM) on n doubles, in place on the device
int wm_comm_allreduce_sum(wm_ctx* ctx, double* dev_buf, int n) {
  if (ctx->nranks == 1) return WM_OK;
  if (!ctx->nccl_comm) {
    wm_set_error("multi-rank reduction requested but wm_comm_init was not called");
    return WM_ERR_ARG;
  }
  NcclApi& a = api();
  WM_NCCL(a.AllReduce(dev_buf, dev_buf, (size_t)n, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  return WM_OK;
}
