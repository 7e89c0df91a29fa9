// wm_internal.cuh -- shared declarations of the B200 backend (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/wuming_b200.h"

// ---------------------------------------------------------------------------------------------
// Device-side geometry: everything the kernels need, passed by value.
// Box arrays (uf, df, uj, tmpf) keep the reference's layout (comp fastest, then x, y, z with two
// ghost layers) so host <-> device transfers of uf are plain copies:
//   box index of (i,j,k) = ((k-(nzs-2))*by + (j-(nys-2)))*bx + (i-(nxgs-2))
// 2-D runs use nzl = 1, bz = 1 and k = 0 everywhere.
// ---------------------------------------------------------------------------------------------
struct Geo {
  int dim, ndim, nsp, np;
  int nxgs, nxge, nygs, nyge, nzgs, nzge, nys, nye, nzs, nze;
  int nx, nyl, nzl;   // global x cells, local y / z cells
  int ny, nz;         // global y / z cells
  int bx, by, bz;     // box dims incl. ghosts
  int npen;           // pencils on this rank over all species: nsp*nyl*nzl
  int multi;          // 1: the slab-axis neighbours (z in 3-D, y in 2-D) are other ranks -> ghost destination rows
  int ngrow;          // pencils per ghost plane and species (nyl in 3-D, 1 in 2-D)
  int nrows;          // destination rows of the sort: npen local + (multi ? 2*nsp*ngrow ghost rows : 0)
  int bc;
  double delx, delt, c, gfac, d_delx, d_delt;
  double q[2], r[2];
  double f1, f2, f3, f4, f5;
  // per-species constants: fac1 = q/r*0.5*delt, fac2 = q*delt/r (particle.f90:101-103), qdxdt = q*delx*d_delt (field.f90:297)
  double fac1[2], fac2[2], qdxdt[2];

  __host__ __device__ inline size_t box(int i, int j, int k) const {
    return ((size_t)(dim == 3 ? (k - (nzs - 2)) : 0) * by + (j - (nys - 2))) * bx + (i - (nxgs - 2));
  }
  __host__ __device__ inline size_t nbox() const { return (size_t)bx * by * bz; }
  // pencil id of (j,k,isp) with isp 0-based: the order of the reference's np2(nys:nye,nzs:nze,nsp)
  __host__ __device__ inline int pen(int j, int k, int isp) const {
    return (isp * nzl + (dim == 3 ? (k - nzs) : 0)) * nyl + (j - nys);
  }
};

// Particle store: structure of arrays, globally ordered by (species, k, j, i-cell).
// Set A ("up") and set B ("gp") hold x,y,(z),ux,uy,uz; the 64-bit ID lives in its own ping-pong
// pair because the push never reorders and so never touches it.
struct Ptcl {
  double* c[6];  // 3-D: x y z ux uy uz ; 2-D: x y ux uy uz (c[5] unused)
};

// ---------------------------------------------------------------------------------------------
// Peer-memory tables of the multi-GPU cgm (wm_fields.cu k_cgm_coop<true>, set up by wm_comm.cu): every rank maps its two
// slab neighbours' CG operand arrays and every rank's reduction mailbox through CUDA IPC, so that the ghost-plane exchange
// and the two all-reduces of a CG iteration are stores / polls over NVLink inside ONE cooperative kernel per GPU.
// ---------------------------------------------------------------------------------------------
constexpr int WM_MAX_PEERS = 16;
struct WmMail { double a, b; unsigned long long seq, pad; };   // one slot per (parity, source rank)
struct PeerCG {
  double* lo[4];        // phi, p0, p1, r of the lower neighbour along the slab axis (peer pointers)
  double* hi[4];        // ... of the upper neighbour
  long long shift_lo;   // box-index shift from my first slab plane to the lower neighbour's upper ghost plane
  long long shift_hi;   // ... from my last slab plane to the upper neighbour's lower ghost plane
  WmMail* mail[WM_MAX_PEERS];   // mailbox [2][nranks] of every rank (mail[rank] is the local one)
  unsigned long long* seq;      // local: number of reductions done so far (same on all ranks)
  int nranks, rank;
};

struct wm_ctx {
  wm_params prm;       // what the kernels run on (y and z exchanged when swap_yz)
  // 3-D y-slabs (the reference's nproc_j > 1, nproc_k = 1: every shipped 3-D sample, 3d/proj/*/config_sample.json) run on the z-slab
  // machinery through an exact relabelling at the host boundary: the device works in (x, y' = z, z' = y).  An axis exchange is a
  // reflection, so the axial vector changes sign: B' = -(Bx, Bz, By), E' = (Ex, Ez, Ey), u' = (ux, uz, uy), J' = (Jx, Jz, Jy) -- with
  // these every equation of the loop (Lorentz force, Faraday, Ampere, the Yee staggering, the Esirkepov factors) is the same code in
  // the primed system; only the order of floating-point sums changes.  wm_upload / wm_download / wm_mom_calc translate pencil
  // order, record columns and field components; hprm keeps the caller's view.
  wm_params hprm;
  bool swap_yz = false;
  double* swapbuf = nullptr;   // nbox x 6 staging of the field translation
  Geo g;
  int device = 0;
  cudaStream_t stream = nullptr;
  // The sort / re-binning / migration of a step needs only what the fused push kernel left behind, the field solve only J: inside
  // wm_step (and the five-call sequence that reaches the same kernels) the sort runs on a second stream -- with its own NCCL
  // communicator -- concurrently with the field solve, and the main stream joins it before anything else touches the particles.
  // Both phases are latency-bound on thin slabs (strong scaling), so overlapping them is where 8-GPU efficiency comes from.
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_fused = nullptr, ev_sort = nullptr;
  void* nccl_comm2 = nullptr;
  int overlap = 1;
  // particles
  size_t cap = 0;      // capacity (particles) of each SoA array
  long long ntot = 0;  // active particles
  long long n_sp0 = 0; // particles of species 0 (they come first)
  Ptcl A, B;
  double* id[2] = {nullptr, nullptr};
  int cid = 0;  // which id array goes with set A
  // cell index: cs[pen*(nx+1) + (i-nxgs)] = absolute start of cell i of pencil pen; entry nx = pencil end
  int* cs = nullptr;
  int* cs_new = nullptr;   // histogram / scan target of the sort
  int* np2 = nullptr;      // particles per pencil (npen)
  int* poff = nullptr;     // pencil offsets (npen+1), absolute
  int* flags = nullptr;    // sticky device error flags
  int* cnt27 = nullptr;    // per (offset, species, source cell) counts -> offsets (wm_sort.cu)
  unsigned char* dst_off = nullptr;  // destination offset (0..26) of every pushed particle
  int* inv = nullptr;      // the sort's permutation: inv[new position] = old position (wm_sort.cu)
  int* goff = nullptr;     // per (destination cell, offset): absolute start of that source group
  int* inc = nullptr;      // multi-rank: arrivals per edge-plane cell announced by the neighbours [side][isp][t][0..nx]
  int* inc_off = nullptr;  // exclusive scan of inc (+ total)
  int* totals = nullptr;   // device copy of the six per-step totals (k_totals)
  size_t dst_off_cap = 0;
  int use_fused = 1;
  // lazy sort (wm_sort.cu): after a sort inside wm_step the permutation is NOT applied; set A then is the pushed set in
  // source order and the particle at sorted position pos is A[inv[pos]] (or, for arrivals from a neighbour rank,
  // R[-2 - inv[pos]]).  The next fused kernel reads through inv; every other consumer calls wm_materialize first.
  bool lazy = false;
  int allow_lazy = 0;            // set by wm_step around its sort
  int lazy_nxs = 0, lazy_nxe = 0;
  Ptcl R = {};                   // arrival store of a slab run (lazy mode): what the neighbours sent, in arrival order
  double* rid = nullptr;
  size_t rcap = 0;
  int pusher = 0;      // WM_PUSHER_BORIS (particle__solv) or WM_PUSHER_VAY (particle__solv_vay)
  void* scan_tmp = nullptr;
  size_t scan_tmp_bytes = 0;
  // fields
  double *uf = nullptr, *df = nullptr, *uj = nullptr, *gkl = nullptr, *tmpf = nullptr;
  double* mom = nullptr;   // moments N,V,T: nsp boxes of 7 components (allocated on first wm_mom_calc)
  // cg work arrays: phi, p (one ghost layer: (nx+2)(nyl+2)(nzl+2)), r, b, ap (interior)
  double *pcg2 = nullptr;  // second search-direction buffer of the cooperative cgm
  double *phi = nullptr, *pcg = nullptr, *rcg = nullptr, *bcg = nullptr, *apcg = nullptr;
  double* red = nullptr;        // reduction scratch (device)
  double* red_host = nullptr;   // pinned
  int cg_ite[3] = {0, 0, 0};
  bool cg_ite_on_device = false;   // the cooperative cgm leaves its iteration counts in totals[6..8]
  // halo buffers
  double* hbuf[4] = {nullptr, nullptr, nullptr, nullptr};  // send lo, send hi, recv lo, recv hi
  size_t hbuf_elems = 0;
  // staging for host <-> device layout conversion
  double* stage = nullptr;
  size_t stage_elems = 0;
  // state machine
  bool gp_valid = false;     // set B holds the pushed state
  bool keys_valid = false;   // migration pass done (histogram in cs_new)
  int last_nxs = 0, last_nxe = 0;  // x range of the last particle__solv (bc__particle_y[z] has no range argument)
  // The reference's five calls on the fast kernels (wm_api.cu): particle__solv on resident state is DEFERRED (nobody reads gp
  // between it and field__fdtd_i: 3d/proj/weibel/app.f90:102-104) together with a following reflecting-wall bc__particle_x /
  // bc__injection (reconnection and shock loops); field__fdtd_i then launches the fused push + boundary + deposit kernel, and the
  // later bc__particle_x (Weibel loop) / bc__particle_y[z] find their work done.  Any other reader of gp runs the deferred
  // procedures with the per-procedure kernels first (flush_deferred).
  bool defer_push = false;       // a particle__solv is pending
  int defer_pusher = 0, defer_nxs = 0, defer_nxe = 0;
  int defer_xbc = 0;             // 0 none, WM_ORDER_RECONNECTION: reflecting bc__particle_x pending, WM_ORDER_SHOCK: bc__injection
  double defer_u0 = 0.0;
  bool fused_done = false;       // the fused kernel ran for this step: set B is pushed, wrapped and counted
  int fused_order = 0;
  // comm
  PeerCG peer = {};
  bool peer_ok = false;            // the peer-memory cgm is usable (all ranks agreed)
  void* peer_arena = nullptr;      // phi | pcg | pcg2 | rcg | mailbox | seq in ONE allocation = one IPC handle per rank
  void* peer_opened[WM_MAX_PEERS] = {};   // bases returned by cudaIpcOpenMemHandle (to close)
  void* nccl_comm = nullptr;
  int nranks = 1, rank = 0;
  int rank_up[2] = {0, 0}, rank_down[2] = {0, 0};  // [0]: y neighbours, [1]: z neighbours
  // bookkeeping
  long long launches = 0;
  int timing = 0;
  // phase timing: five events per timed step from a pool, resolved (one host sync) only when the pool is full or the sums are
  // read -- wm_step never waits on the host between steps
  static constexpr int EV_STEPS = 64, EV_PER = 7;   // per step: start, fused end, field start, field end, step end, sort start, sort end
  cudaEvent_t ev[EV_PER * EV_STEPS] = {};
  int ev_used = 0;               // timed steps recorded and not yet folded into ms_sum
  float ms_phase[4] = {0, 0, 0, 0};
  double ms_sum[4] = {0, 0, 0, 0};
  int timed_steps = 0;
};

// Sort-beside-field-solve policy (wm_api.cu sort_after_fused, wm_fields.cu wm_k_cgm): returns 0 when the sort stays on the main
// stream, else the blocks per SM the cooperative cgm may take while the sort runs beside it.  The overlap pays where both phases
// are latency-bound, i.e. on thin slabs of a multi-GPU run (measured on 8 x B200, fixed 256x256x128 box: 20.2 -> 18.6 ms per step;
// on 4.2 M cells per rank it costs time).  The cgm grid is capped at <= 4 blocks of 256 threads per SM so that the sort's NCCL
// transfer kernels can always become resident beside it: a cooperative grid that fills every SM while the neighbours' solves wait
// for this rank's reductions and this rank's transfer kernel waits for the neighbours' would be a cross-rank dependency cycle.
static inline int wm_overlap_bps(const wm_ctx* c, long long ncells) {
  if (!c->overlap || c->nranks == 1 || !c->nccl_comm2 || !c->stream2) return 0;
  const long long per_bps = 148LL * 256 * 16;            // ~16 cells per thread and sweep
  const long long bps = (ncells + per_bps - 1) / per_bps;
  if (bps > 4) return 0;
  return (int)(bps < 2 ? 2 : bps);
}

// error plumbing -------------------------------------------------------------------------------
void wm_set_error(const std::string& msg);
#define WM_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      wm_set_error(std::string(#call) + ": " + cudaGetErrorString(e__) + " at " + __FILE__ + \
                   ":" + std::to_string(__LINE__));                                        \
      return WM_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)
#define WM_TRY(call)            \
  do {                          \
    int r__ = (call);           \
    if (r__ != WM_OK) return r__; \
  } while (0)
#define WM_LAUNCH_CHECK(ctx)               \
  do {                                     \
    (ctx)->launches++;                     \
    WM_CUDA(cudaGetLastError());           \
  } while (0)

static inline int wm_blocks(long long n, int threads) { return (int)((n + threads - 1) / threads); }

// sub-module entry points ------------------------------------------------------------------------
// particles (wm_particles.cu)
int wm_k_tmpf(wm_ctx* ctx, int nxs, int nxe);
int wm_k_push(wm_ctx* ctx, int nxs, int nxe);
int wm_k_deposit(wm_ctx* ctx, int nxs, int nxe);
int wm_k_mom(wm_ctx* ctx, int nxs, int nxe);
int wm_k_mom_fold(wm_ctx* ctx);
int wm_k_push_deposit_fused(wm_ctx* ctx, int nxs, int nxe, int order, double u0);
bool wm_fused_supported(const wm_ctx* ctx, int order);
int wm_k_bc_x(wm_ctx* ctx, int nxs, int nxe, int kind, double u0);
// sort / migration (wm_sort.cu)
int wm_sort_prepare(wm_ctx* ctx);
int wm_k_classify(wm_ctx* ctx, int nxs, int nxe);
int wm_k_sort(wm_ctx* ctx, int nxs, int nxe);
int wm_enable_slab_migration(wm_ctx* ctx);
int wm_k_refresh_np2(wm_ctx* ctx);
int wm_k_cells_consistent(wm_ctx* ctx, int* inconsistent);
int wm_materialize(wm_ctx* ctx);   // apply a pending (lazy) sort permutation: set A becomes the sorted set again
int wm_k_energy(wm_ctx* ctx, double* out_host);
int wm_k_gauss(wm_ctx* ctx, double* out_host);
int wm_k_load_weibel(wm_ctx* ctx, int n0, double v_thi, double v_the, double t_ani, double b0, unsigned long long seed);
int wm_k_aos_to_soa(wm_ctx* ctx, const double* stage, Ptcl dst, double* dst_id, int pen0, int npens, int maxcnt);
int wm_k_soa_to_aos(wm_ctx* ctx, double* stage, Ptcl src, const double* src_id, int pen0, int npens, int maxcnt);
int wm_k_swap_box6(wm_ctx* ctx, const double* in, double* out, bool to_device);
int wm_k_cs_from_cumcnt(wm_ctx* ctx, const int* cumcnt_dev);
int wm_k_cumcnt_from_cs(wm_ctx* ctx, int* cumcnt_dev);
// fields (wm_fields.cu)
int wm_k_zero_uj(wm_ctx* ctx, int nxs, int nxe);
int wm_k_curre(wm_ctx* ctx, int nxs, int nxe);
int wm_k_gkl(wm_ctx* ctx, int nxs, int nxe);
int wm_k_cgm(wm_ctx* ctx, int nxs, int nxe);
int wm_k_dfield(wm_ctx* ctx, int nxs, int nxe);
int wm_k_de(wm_ctx* ctx, int nxs, int nxe);
int wm_k_update(wm_ctx* ctx, int nxs, int nxe);
// comm (wm_comm.cu)
int wm_comm_sendrecv(wm_ctx* ctx, int axis, int dir_down, const double* snd, double* rcv, size_t n);
int wm_comm_allreduce_sum(wm_ctx* ctx, double* dev_buf, int n);
int wm_comm_group_begin(wm_ctx* ctx);
int wm_comm_group_end(wm_ctx* ctx);
int wm_comm_send(wm_ctx* ctx, int peer, const void* buf, size_t bytes);
int wm_comm_recv(wm_ctx* ctx, int peer, void* buf, size_t bytes);
int wm_comm_destroy(wm_ctx* ctx);
