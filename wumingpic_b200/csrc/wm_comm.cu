// wm_comm.cu -- inter-GPU transport: the NCCL replacement of the reference's MPI_SENDRECV /
// MPI_ALLREDUCE call sites.  One process per GPU; the communicator is created from
// a ncclUniqueId the host broadcasts (MPI_Bcast in the Fortran driver, torch.distributed in bench.py).
//
// NCCL is resolved at run time with dlopen so that (a) single-GPU use needs no NCCL at all and
// (b) inside a Python process that already loaded torch's bundled libnccl.so.2 the same library
// instance is reused instead of a second copy.
#include "wm_internal.cuh"

#include <dlfcn.h>
#include <cstdlib>
#include <cstring>
#include <vector>

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
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*) = nullptr;   // NCCL >= 2.18
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
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
  LOAD(CommSplit, "ncclCommSplit");
  LOAD(Send, "ncclSend");
  LOAD(Recv, "ncclRecv");
  LOAD(AllReduce, "ncclAllReduce");
  LOAD(AllGather, "ncclAllGather");
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


// ---------------------------------------------------------------------------------------------
// Peer-memory set-up of the multi-GPU cgm (PeerCG, wm_internal.cuh).  The four CG operand arrays that neighbours write into
// (phi, p, p2, r), the reduction mailbox and the reduction counter move into ONE allocation per rank, whose CUDA IPC handle
// (+ the offset inside the driver's underlying block: small allocations are sub-allocated) is all-gathered over NCCL; every
// rank then maps all peers' arenas.  A magic-number handshake through the mapped pointers verifies the mapping, and the ranks
// agree collectively (all-reduce) before k_cgm_coop<true> is used; otherwise the NCCL-sequenced solve stays in place.
// ---------------------------------------------------------------------------------------------
struct PeerBlob {
  cudaIpcMemHandle_t handle;    // 64 bytes
  unsigned long long offset;    // arena - base of the underlying allocation
  int nl;                       // slab thickness (nzl in 3-D, nyl in 2-D)
  int rank;
};

size_t round256(size_t b) { return (b + 255) / 256 * 256; }

int peer_setup(wm_ctx* ctx) {
  static const bool off = getenv("WM_CG_NCCL") != nullptr;   // measurement switch: keep the NCCL-sequenced solve
  NcclApi& a = api();
  const int R = ctx->nranks;
  ctx->peer_ok = false;
  if (off || R > WM_MAX_PEERS || !a.AllGather) return WM_OK;
  const Geo& g = ctx->g;
  const size_t arr = round256(g.nbox() * sizeof(double));
  const size_t mail_bytes = round256((size_t)2 * R * sizeof(WmMail));
  const size_t total = 4 * arr + mail_bytes + 256;
  unsigned char* arena = nullptr;
  WM_CUDA(cudaMalloc(&arena, total));
  WM_CUDA(cudaMemsetAsync(arena, 0, total, ctx->stream));
  // the operand arrays move into the arena
  double** old[4] = {&ctx->phi, &ctx->pcg, &ctx->pcg2, &ctx->rcg};
  for (int q = 0; q < 4; ++q) {
    if (*old[q]) cudaFree(*old[q]);
    *old[q] = reinterpret_cast<double*>(arena + q * arr);
  }
  ctx->peer_arena = arena;
  unsigned long long* my_seq = reinterpret_cast<unsigned long long*>(arena + 4 * arr + mail_bytes);

  // base of the underlying allocation (driver API, resolved at run time like NCCL)
  PeerBlob mine;
  std::memset(&mine, 0, sizeof(mine));
  bool ok = true;
  {
    typedef int (*GetRange)(unsigned long long*, size_t*, unsigned long long);
    void* cu = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    GetRange get = cu ? (GetRange)dlsym(cu, "cuMemGetAddressRange_v2") : nullptr;
    unsigned long long base = 0;
    size_t sz = 0;
    if (!get || get(&base, &sz, (unsigned long long)(uintptr_t)arena) != 0) ok = false;
    else mine.offset = (unsigned long long)(uintptr_t)arena - base;
  }
  if (ok && cudaIpcGetMemHandle(&mine.handle, arena) != cudaSuccess) { ok = false; cudaGetLastError(); }
  mine.nl = g.dim == 3 ? g.nzl : g.nyl;
  mine.rank = ok ? ctx->rank : -1;

  // all-gather the blobs
  unsigned char* dev = nullptr;
  WM_CUDA(cudaMalloc(&dev, sizeof(PeerBlob) * (R + 1)));
  WM_CUDA(cudaMemcpyAsync(dev, &mine, sizeof(PeerBlob), cudaMemcpyHostToDevice, ctx->stream));
  WM_NCCL(a.AllGather(dev, dev + sizeof(PeerBlob), sizeof(PeerBlob), ncclInt8, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  std::vector<PeerBlob> all(R);
  WM_CUDA(cudaMemcpyAsync(all.data(), dev + sizeof(PeerBlob), sizeof(PeerBlob) * R, cudaMemcpyDeviceToHost, ctx->stream));
  WM_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int r = 0; r < R; ++r) if (all[r].rank != r) ok = false;

  // map every peer's arena
  unsigned char* arenas[WM_MAX_PEERS] = {};
  if (ok) {
    for (int r = 0; r < R && ok; ++r) {
      if (r == ctx->rank) { arenas[r] = arena; continue; }
      void* base = nullptr;
      if (cudaIpcOpenMemHandle(&base, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = false;
        break;
      }
      ctx->peer_opened[r] = base;
      arenas[r] = static_cast<unsigned char*>(base) + all[r].offset;
    }
  }
  // handshake: every rank writes a magic number behind its counter and reads everybody else's through the mapping
  const unsigned long long magic = 0x574d5045455200ull;   // "WMPEER"
  unsigned long long mine_magic = magic + (unsigned long long)ctx->rank;
  WM_CUDA(cudaMemcpyAsync(my_seq + 1, &mine_magic, sizeof(mine_magic), cudaMemcpyHostToDevice, ctx->stream));
  double* flag = reinterpret_cast<double*>(dev);   // reuse: 1.0 = ok on this rank
  double one = 1.0;
  WM_CUDA(cudaMemcpyAsync(flag, &one, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  WM_NCCL(a.AllReduce(flag, flag, 1, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));   // barrier: magic written everywhere
  WM_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ok) {
    for (int r = 0; r < R && ok; ++r) {
      // a neighbour's arena differs in size when the slabs are uneven: its counter sits behind ITS four arrays
      const size_t nbox_r = g.dim == 3 ? (size_t)g.bx * g.by * (all[r].nl + 4) : (size_t)g.bx * (all[r].nl + 4);
      const size_t off_r = 4 * round256(nbox_r * sizeof(double)) + mail_bytes;
      unsigned long long got = 0;
      if (cudaMemcpy(&got, arenas[r] + off_r + 8, sizeof(got), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); ok = false; }
      else if (got != magic + (unsigned long long)r) ok = false;
    }
  }
  double okv = ok ? 0.0 : 1.0;   // number of ranks that failed
  WM_CUDA(cudaMemcpyAsync(flag, &okv, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  WM_NCCL(a.AllReduce(flag, flag, 1, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  WM_CUDA(cudaMemcpyAsync(&okv, flag, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  WM_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(dev);
  if (okv != 0.0) {
    fprintf(stderr, "[wuming_b200] rank %d: peer-memory cgm unavailable (CUDA IPC mapping failed on %d rank(s)); using the NCCL-sequenced solve\n",
            ctx->rank, (int)okv);
    return WM_OK;
  }
  // tables
  const int ax = g.dim == 3 ? 1 : 0;
  const int lo = ctx->rank_down[ax], hi = ctx->rank_up[ax];
  const long long ss = g.dim == 3 ? (long long)g.bx * g.by : (long long)g.bx;   // box stride of the slab axis
  PeerCG& pc = ctx->peer;
  auto arr_of = [&](int r) {
    const size_t nbox_r = g.dim == 3 ? (size_t)g.bx * g.by * (all[r].nl + 4) : (size_t)g.bx * (all[r].nl + 4);
    return round256(nbox_r * sizeof(double));
  };
  for (int q = 0; q < 4; ++q) {
    pc.lo[q] = reinterpret_cast<double*>(arenas[lo] + q * arr_of(lo));
    pc.hi[q] = reinterpret_cast<double*>(arenas[hi] + q * arr_of(hi));
  }
  pc.shift_lo = (long long)all[lo].nl * ss;
  pc.shift_hi = -(long long)mine.nl * ss;
  for (int r = 0; r < R; ++r) pc.mail[r] = reinterpret_cast<WmMail*>(arenas[r] + 4 * arr_of(r));
  pc.seq = my_seq;
  pc.nranks = R;
  pc.rank = ctx->rank;
  ctx->peer_ok = true;
  return WM_OK;
}

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
  // a second communicator over the same ranks for the sort / migration stream (wm_api.cu sort_after_fused): operations of one
  // communicator must not run concurrently from two streams
  ctx->nccl_comm2 = nullptr;
  if (a.CommSplit && ctx->overlap) {
    ncclComm_t comm2 = nullptr;
    if (a.CommSplit(comm, 0, rank, &comm2, nullptr) == 0 && comm2) ctx->nccl_comm2 = comm2;
  }
  WM_TRY(peer_setup(ctx));
  return wm_enable_slab_migration(ctx);
}

int wm_comm_destroy(wm_ctx* ctx) {
  for (int r = 0; r < WM_MAX_PEERS; ++r)
    if (ctx->peer_opened[r]) { cudaIpcCloseMemHandle(ctx->peer_opened[r]); ctx->peer_opened[r] = nullptr; }
  if (ctx->peer_arena) {
    cudaFree(ctx->peer_arena);
    ctx->peer_arena = nullptr;
    ctx->phi = ctx->pcg = ctx->pcg2 = ctx->rcg = nullptr;   // they lived in the arena
    ctx->peer_ok = false;
  }
  if (ctx->nccl_comm2 && api().CommDestroy) api().CommDestroy((ncclComm_t)ctx->nccl_comm2);
  ctx->nccl_comm2 = nullptr;
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
