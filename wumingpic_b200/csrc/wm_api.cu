// wm_api.cu -- the C ABI of include/wuming_b200.h: context life cycle, host <-> device state
// transfer in the reference's array layouts, and the per-procedure entry points that sequence the
// kernels of wm_particles.cu / wm_fields.cu exactly like the reference's time loop
// (3d/proj/weibel/app.f90:100-108).
#include "wm_internal.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

int wm_comm_destroy(wm_ctx* ctx);
void wm_sort_timing_report(int rank);
static double g_stage_ms[9] = {0};   // WM_FIELD_TIMING (see field_stages)
static long g_stage_calls = 0;

namespace {
thread_local std::string g_err;
constexpr double kPi = 3.14159265358979323846264338327950288;

int free_particles(wm_ctx* c) {
  for (int k = 0; k < 6; ++k) {
    if (c->A.c[k]) cudaFree(c->A.c[k]);
    if (c->B.c[k]) cudaFree(c->B.c[k]);
    c->A.c[k] = c->B.c[k] = nullptr;
  }
  for (int k = 0; k < 2; ++k) {
    if (c->id[k]) cudaFree(c->id[k]);
    c->id[k] = nullptr;
  }
  for (int k = 0; k < 6; ++k) {
    if (c->R.c[k]) cudaFree(c->R.c[k]);
    c->R.c[k] = nullptr;
  }
  if (c->rid) cudaFree(c->rid);
  c->rid = nullptr;
  c->rcap = 0;
  c->cap = 0;
  c->lazy = false;     // whatever permutation was pending referred to the freed arrays
  return WM_OK;
}

// make room for `need` particles (keeps no contents: callers refill)
int reserve_particles(wm_ctx* c, size_t need) {
  if (need <= c->cap) return WM_OK;
  free_particles(c);
  double factor = c->nranks > 1 ? 1.25 : 1.0;
  if (const char* e = getenv("WM_CAP_FACTOR")) factor = atof(e);
  size_t cap = (size_t)(need * factor) + 1024;
  const int ncomp = c->g.ndim - 1;
  for (int k = 0; k < ncomp; ++k) {
    WM_CUDA(cudaMalloc(&c->A.c[k], cap * sizeof(double)));
    WM_CUDA(cudaMalloc(&c->B.c[k], cap * sizeof(double)));
  }
  for (int k = 0; k < 2; ++k) WM_CUDA(cudaMalloc(&c->id[k], cap * sizeof(double)));
  c->cap = cap;
  return WM_OK;
}

int reserve_stage(wm_ctx* c, size_t elems) {
  if (elems <= c->stage_elems) return WM_OK;
  if (c->stage) cudaFree(c->stage);
  c->stage = nullptr;
  WM_CUDA(cudaMalloc(&c->stage, elems * sizeof(double)));
  c->stage_elems = elems;
  return WM_OK;
}

int check_flags(wm_ctx* c) {
  int f = 0;
  WM_CUDA(cudaMemcpyAsync(&f, c->flags, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  WM_CUDA(cudaStreamSynchronize(c->stream));
  if (f & 1) {
    wm_set_error("memory over (np2 > np)");
    return WM_ERR_MEMORY_OVER;
  }
  if (f & 4) {
    wm_set_error("********** stop at cgm after ite_max **********");
    return WM_ERR_CG_ITEMAX;
  }
  if (f & 8) {
    wm_set_error("peer-memory cgm: a neighbour rank never arrived at a reduction (rank died or diverged)");
    return WM_ERR_CUDA;
  }
  if (f & 2) {
    wm_set_error("a particle left the one-cell neighbourhood of its cell (|inc| > 1 or outside the slab)");
    return WM_ERR_PARTICLE_LOST;
  }
  return WM_OK;
}

// y-slab relabelling (wm_ctx::swap_yz): caller pencil (isp, k, j) <-> device pencil (isp, k' = j, j' = k)
int pen_host_to_dev(const wm_ctx* c, int pen_h) {
  const Geo& g = c->g;
  const int nyl_h = g.nzl, nzl_h = g.nyl;
  const int jj = pen_h % nyl_h, kk = (pen_h / nyl_h) % nzl_h, isp = pen_h / (nyl_h * nzl_h);
  return (isp * g.nzl + jj) * g.nyl + kk;
}
// rows of `w` ints per pencil from the caller's pencil order into the device's (to_dev) or back
void permute_pencil_rows(const wm_ctx* c, const int* in, int* out, int w, bool to_dev) {
  for (int ph = 0; ph < c->g.npen; ++ph) {
    const int pd = pen_host_to_dev(c, ph);
    const int* s = in + (size_t)(to_dev ? ph : pd) * w;
    int* d = out + (size_t)(to_dev ? pd : ph) * w;
    for (int i = 0; i < w; ++i) d[i] = s[i];
  }
}
int need_swapbuf(wm_ctx* c) {
  if (!c->swapbuf) WM_CUDA(cudaMalloc(&c->swapbuf, c->g.nbox() * 6 * sizeof(double)));
  return WM_OK;
}
int no_swap(const wm_ctx* c, const char* what) {
  if (!c->swap_yz) return WM_OK;
  wm_set_error(std::string(what) + " is not available with 3-D y-slabs (the device works in the relabelled (x, z, y) system): "
               "use z-slabs (nproc_j = 1), or the host form of this procedure between wm_download and wm_upload");
  return WM_ERR_STATE;
}

bool range_ok(const wm_ctx* c, int nxs, int nxe) {
  return nxs >= c->g.nxgs && nxe <= c->g.nxge && nxs <= nxe;
}
}  // namespace

void wm_set_error(const std::string& msg) { g_err = msg; }

extern "C" {
static int flush_deferred(wm_ctx* c);
static int fold_timing(wm_ctx* c);

const char* wm_last_error(void) { return g_err.c_str(); }
int wm_version(void) { return 100; }

int wm_para_range(int n1, int n2, int isize, int irank, int* ns, int* ne) {
  // 3d/common/mpi_set.f90:81-94
  if (isize <= 0 || irank < 0 || irank >= isize || !ns || !ne) return WM_ERR_ARG;
  int iwork1 = (n2 - n1 + 1) / isize;
  int iwork2 = (n2 - n1 + 1) % isize;
  *ns = irank * iwork1 + n1 + std::min(irank, iwork2);
  *ne = *ns + iwork1 - 1;
  if (iwork2 > irank) *ne = *ne + 1;
  return WM_OK;
}

int wm_create(const wm_params* prm, wm_ctx** out) {
  if (!prm || !out) return WM_ERR_ARG;
  *out = nullptr;
  {
  const wm_params& p = *prm;
  if (!((p.dim == 3 && p.ndim == 7) || (p.dim == 2 && p.ndim == 6)) || p.nsp != 2 || p.np <= 0) {
    wm_set_error("wm_create: need (dim,ndim) = (3,7) or (2,6), nsp = 2, np > 0");
    return WM_ERR_ARG;
  }
  if (p.nxge < p.nxgs || p.nye < p.nys || (p.dim == 3 && p.nze < p.nzs) || p.nproc_j < 1 || p.nproc_k < 1) {
    wm_set_error("wm_create: empty index range");
    return WM_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    wm_set_error("wm_create: no CUDA device (this backend has no CPU fallback)");
    return WM_ERR_CUDA;
  }
  }
  wm_ctx* c = new wm_ctx();
  c->hprm = *prm;
  // 3-D y-slabs (nproc_j > 1, nproc_k = 1) -> z'-slabs of the relabelled system (wm_internal.cuh, wm_ctx::swap_yz); WM_SWAP_YZ=1
  // forces the relabelling on any 3-D run (how the single-GPU tests exercise it)
  wm_params pp = *prm;
  if (pp.dim == 3 && pp.nproc_k == 1 && (pp.nproc_j > 1 || getenv("WM_SWAP_YZ") != nullptr)) {
    std::swap(pp.nygs, pp.nzgs); std::swap(pp.nyge, pp.nzge);
    std::swap(pp.nys, pp.nzs);   std::swap(pp.nye, pp.nze);
    std::swap(pp.nproc_j, pp.nproc_k); std::swap(pp.rank_j, pp.rank_k);
    c->swap_yz = true;
  }
  const wm_params& p = pp;
  c->prm = p;
  for (int k = 0; k < 6; ++k) c->A.c[k] = c->B.c[k] = nullptr;
  if (p.device >= 0) {
    c->device = p.device;
  } else {
    cudaGetDevice(&c->device);
  }
  WM_CUDA(cudaSetDevice(c->device));
  {
    // Main stream (push, field solve) and second stream (the sort that may run beside the field solve) have EQUAL priority by
    // default: measured on 8 GPUs (fixed 256x256x128 box), equal priorities give 18.6 ms per step, a low-priority sort stream 19.6 ms
    // (the starved sort, with its two NCCL phases and host read-back, becomes the longer chain).  WM_STREAM_PRIO=1: main high / sort
    // low, 2: sort high / main low (measurement switch).
    int lo = 0, hi = 0;
    WM_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    const int mode = getenv("WM_STREAM_PRIO") ? atoi(getenv("WM_STREAM_PRIO")) : 0;
    const int pm = mode == 1 ? hi : lo, ps = mode == 2 ? hi : lo;
    WM_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, pm));
    WM_CUDA(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, ps));
  }
  Geo& g = c->g;
  g.dim = p.dim; g.ndim = p.ndim; g.nsp = p.nsp; g.np = p.np;
  g.nxgs = p.nxgs; g.nxge = p.nxge; g.nygs = p.nygs; g.nyge = p.nyge;
  g.nys = p.nys; g.nye = p.nye;
  if (p.dim == 3) { g.nzgs = p.nzgs; g.nzge = p.nzge; g.nzs = p.nzs; g.nze = p.nze; }
  else { g.nzgs = g.nzge = g.nzs = g.nze = 0; }
  g.nx = g.nxge - g.nxgs + 1; g.ny = g.nyge - g.nygs + 1; g.nz = p.dim == 3 ? g.nzge - g.nzgs + 1 : 1;
  g.nyl = g.nye - g.nys + 1; g.nzl = p.dim == 3 ? g.nze - g.nzs + 1 : 1;
  g.bx = g.nx + 4; g.by = g.nyl + 4; g.bz = p.dim == 3 ? g.nzl + 4 : 1;
  g.npen = g.nsp * g.nyl * g.nzl;
  g.multi = 0;                              // set by wm_comm_init when the slab-axis neighbours are other ranks
  g.ngrow = p.dim == 3 ? g.nyl : 1;
  g.nrows = g.npen;
  g.bc = p.bc_kind;
  g.delx = p.delx; g.delt = p.delt; g.c = p.c; g.gfac = p.gfac;
  g.d_delx = 1.0 / p.delx; g.d_delt = 1.0 / p.delt;
  for (int s = 0; s < 2; ++s) { g.q[s] = p.q[s]; g.r[s] = p.r[s]; }
  for (int s = 0; s < 2; ++s) {
    g.fac1[s] = p.q[s] / p.r[s] * 5e-1 * p.delt;
    g.fac2[s] = p.q[s] * p.delt / p.r[s];
    g.qdxdt[s] = p.q[s] * p.delx * g.d_delt;
  }
  // field__init: 3d/common/field.f90:57-63, 2d/common/field.f90:53-59
  g.f1 = p.c * p.delt / p.delx;
  g.f2 = p.gfac * g.f1 * g.f1;
  g.f3 = 4.0 * kPi * p.delx / p.c;
  g.f5 = std::pow(p.delx / (p.c * p.delt * p.gfac), 2);
  g.f4 = (p.dim == 3 ? 6.0 : 4.0) + g.f5;
  // mpi_set__init neighbour table (3d/common/mpi_set.f90:63-76), periodic in y and z
  c->last_nxs = g.nxgs;
  c->last_nxe = g.nxge;
  c->nranks = 1;  // until wm_comm_init
  c->rank = p.rank_j * p.nproc_k + p.rank_k;
  auto rk = [&](int j, int k) { return ((j + p.nproc_j) % p.nproc_j) * p.nproc_k + ((k + p.nproc_k) % p.nproc_k); };
  c->rank_up[0] = rk(p.rank_j + 1, p.rank_k); c->rank_down[0] = rk(p.rank_j - 1, p.rank_k);
  c->rank_up[1] = rk(p.rank_j, p.rank_k + 1); c->rank_down[1] = rk(p.rank_j, p.rank_k - 1);

  const size_t nb = g.nbox();
  WM_CUDA(cudaMalloc(&c->uf, nb * 6 * sizeof(double)));
  WM_CUDA(cudaMalloc(&c->df, nb * 6 * sizeof(double)));
  WM_CUDA(cudaMalloc(&c->tmpf, nb * 6 * sizeof(double)));
  WM_CUDA(cudaMalloc(&c->uj, nb * 3 * sizeof(double)));
  WM_CUDA(cudaMalloc(&c->gkl, nb * 3 * sizeof(double)));
  double** cg[6] = {&c->phi, &c->pcg, &c->rcg, &c->bcg, &c->apcg, &c->pcg2};
  for (auto pp : cg) {
    WM_CUDA(cudaMalloc(pp, nb * sizeof(double)));
    WM_CUDA(cudaMemsetAsync(*pp, 0, nb * sizeof(double), c->stream));
  }
  // field.f90:109-123: df, gkl, uj start at zero (df is the CG warm start)
  WM_CUDA(cudaMemsetAsync(c->uf, 0, nb * 6 * sizeof(double), c->stream));
  WM_CUDA(cudaMemsetAsync(c->df, 0, nb * 6 * sizeof(double), c->stream));
  WM_CUDA(cudaMemsetAsync(c->tmpf, 0, nb * 6 * sizeof(double), c->stream));
  WM_CUDA(cudaMemsetAsync(c->uj, 0, nb * 3 * sizeof(double), c->stream));
  WM_CUDA(cudaMemsetAsync(c->gkl, 0, nb * 3 * sizeof(double), c->stream));
  // cell index rows: npen local pencils + the ghost rows of a slab run (wm_sort.cu), nx+1 entries each, + grand total
  const size_t ncs = ((size_t)g.npen + 2 * g.nsp * g.ngrow) * (g.nx + 1) + 1;
  WM_CUDA(cudaMalloc(&c->cs, ncs * sizeof(int)));
  WM_CUDA(cudaMalloc(&c->cs_new, ncs * sizeof(int)));
  WM_CUDA(cudaMemsetAsync(c->cs, 0, ncs * sizeof(int), c->stream));
  WM_CUDA(cudaMemsetAsync(c->cs_new, 0, ncs * sizeof(int), c->stream));
  const size_t ninc = (size_t)2 * g.nsp * g.ngrow * (g.nx + 1) + 1;
  WM_CUDA(cudaMalloc(&c->inc, ninc * sizeof(int)));
  WM_CUDA(cudaMalloc(&c->inc_off, ninc * sizeof(int)));
  WM_CUDA(cudaMalloc(&c->totals, 16 * sizeof(int)));
  WM_CUDA(cudaMemsetAsync(c->totals, 0, 16 * sizeof(int), c->stream));
  WM_CUDA(cudaMemsetAsync(c->inc, 0, ninc * sizeof(int), c->stream));
  WM_CUDA(cudaMalloc(&c->np2, (size_t)g.npen * sizeof(int)));
  WM_CUDA(cudaMalloc(&c->poff, ((size_t)g.npen + 1) * sizeof(int)));
  WM_CUDA(cudaMemsetAsync(c->np2, 0, (size_t)g.npen * sizeof(int), c->stream));
  WM_CUDA(cudaMemsetAsync(c->poff, 0, ((size_t)g.npen + 1) * sizeof(int), c->stream));
  WM_CUDA(cudaMalloc(&c->flags, 4 * sizeof(int)));
  WM_CUDA(cudaMemsetAsync(c->flags, 0, 4 * sizeof(int), c->stream));
  WM_CUDA(cudaMalloc(&c->red, (4096 + 64) * sizeof(double)));
  WM_CUDA(cudaMemsetAsync(c->red, 0, (4096 + 64) * sizeof(double), c->stream));
  WM_CUDA(cudaMallocHost(&c->red_host, 64 * sizeof(double)));
  c->hbuf_elems = (size_t)2 * 6 * g.bx * std::max(g.by, g.bz);
  WM_CUDA(cudaMalloc(&c->hbuf[0], c->hbuf_elems * sizeof(double)));
  WM_CUDA(cudaMalloc(&c->hbuf[2], c->hbuf_elems * sizeof(double)));
  for (int e = 0; e < wm_ctx::EV_PER * wm_ctx::EV_STEPS; ++e) WM_CUDA(cudaEventCreate(&c->ev[e]));
  WM_CUDA(cudaEventCreateWithFlags(&c->ev_fused, cudaEventDisableTiming));
  WM_CUDA(cudaEventCreateWithFlags(&c->ev_sort, cudaEventDisableTiming));
  if (const char* e = getenv("WM_OVERLAP_SORT")) c->overlap = atoi(e);
  WM_CUDA(cudaStreamSynchronize(c->stream));
  *out = c;
  return WM_OK;
}

int wm_destroy(wm_ctx* c) {
  if (!c) return WM_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (g_stage_calls > 0) {
    fprintf(stderr, "[wuming_b200] rank %d field__fdtd_i stages over %ld calls (ms/call): ele_cur %.3f curre %.3f gkl %.3f cgm %.3f dfield %.3f dE %.3f dfield %.3f update %.3f\n",
            c->rank, g_stage_calls, g_stage_ms[1] / g_stage_calls, g_stage_ms[2] / g_stage_calls, g_stage_ms[3] / g_stage_calls,
            g_stage_ms[4] / g_stage_calls, g_stage_ms[5] / g_stage_calls, g_stage_ms[6] / g_stage_calls, g_stage_ms[7] / g_stage_calls,
            g_stage_ms[8] / g_stage_calls);
    g_stage_calls = 0;
  }
  wm_sort_timing_report(c->rank);
  wm_comm_destroy(c);   // also releases the peer arena (and nulls the CG arrays that lived in it)
  free_particles(c);
  double* d[] = {c->swapbuf, c->mom, c->uf, c->df, c->uj, c->gkl, c->tmpf, c->phi, c->pcg, c->pcg2, c->rcg, c->bcg, c->apcg, c->red, c->hbuf[0],
                 c->hbuf[2], c->stage};
  for (double* p : d) if (p) cudaFree(p);
  int* ii[] = {c->cs, c->cs_new, c->np2, c->poff, c->flags, c->cnt27, c->inc, c->inc_off, c->totals, c->inv, c->goff};
  if (c->dst_off) cudaFree(c->dst_off);
  for (int* p : ii) if (p) cudaFree(p);
  if (c->scan_tmp) cudaFree(c->scan_tmp);
  if (c->red_host) cudaFreeHost(c->red_host);
  for (int e = 0; e < wm_ctx::EV_PER * wm_ctx::EV_STEPS; ++e) if (c->ev[e]) cudaEventDestroy(c->ev[e]);
  if (c->ev_fused) cudaEventDestroy(c->ev_fused);
  if (c->ev_sort) cudaEventDestroy(c->ev_sort);
  if (c->stream2) { cudaStreamSynchronize(c->stream2); cudaStreamDestroy(c->stream2); }
  cudaStreamDestroy(c->stream);
  delete c;
  return WM_OK;
}

// ---------------------------------------------------------------------------------------------
// state transfer
// ---------------------------------------------------------------------------------------------
int wm_upload(wm_ctx* c, const double* up, const int* np2, const int* cumcnt, const double* uf) {
  if (!c) return WM_ERR_ARG;
  WM_CUDA(cudaSetDevice(c->device));
  const Geo& g = c->g;
  std::vector<int> np2_d, cumcnt_d;      // y-slab relabelling: the caller's small integer arrays in the device's pencil order
  if (c->swap_yz) {
    if (np2) { np2_d.resize(g.npen); permute_pencil_rows(c, np2, np2_d.data(), 1, true); np2 = np2_d.data(); }
    if (cumcnt) { cumcnt_d.resize((size_t)g.npen * (g.nx + 1)); permute_pencil_rows(c, cumcnt, cumcnt_d.data(), g.nx + 1, true); cumcnt = cumcnt_d.data(); }
  }
  if (uf && c->swap_yz) {
    WM_TRY(need_swapbuf(c));
    WM_CUDA(cudaMemcpyAsync(c->swapbuf, uf, g.nbox() * 6 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    WM_TRY(wm_k_swap_box6(c, c->swapbuf, c->uf, true));
  } else if (uf) WM_CUDA(cudaMemcpyAsync(c->uf, uf, g.nbox() * 6 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (up) {
    c->lazy = false;   // the uploaded state supersedes a pending (lazy) sort permutation
    c->defer_push = false;   // ... and whatever was deferred on the old state
    c->defer_xbc = 0;
    c->fused_done = false;
    if (!np2 || !cumcnt) {
      wm_set_error("wm_upload: up needs np2 and cumcnt");
      return WM_ERR_ARG;
    }
    std::vector<int> poff(g.npen + 1, 0);
    int maxcnt = 0;
    for (int pen = 0; pen < g.npen; ++pen) {
      if (np2[pen] < 0 || np2[pen] > g.np) {
        wm_set_error("memory over (np2 > np)");
        return WM_ERR_MEMORY_OVER;
      }
      poff[pen + 1] = poff[pen] + np2[pen];
      maxcnt = std::max(maxcnt, np2[pen]);
    }
    c->ntot = poff[g.npen];
    c->n_sp0 = poff[g.npen / g.nsp];
    WM_TRY(reserve_particles(c, (size_t)c->ntot));
    std::vector<int> cs((size_t)g.npen * (g.nx + 1) + 1);
    // Cell i of a pencil holds the records (cumcnt(i), cumcnt(i+1)] (particle.f90:100-108); the pencil population is np2, NOT
    // cumcnt(nxe+1): the shock driver leaves cumcnt above nxe stale (inject / relocate bump np2 and cumcnt(nxe) only,
    // 2d/proj/shock/app.f90:836-838; init never fills cumcnt(nxe+1:), :346-359), and a stale entry below its predecessor is an
    // empty cell in the reference's loops.  The device index is the same membership made monotone and closed with np2.
    for (int pen = 0; pen < g.npen; ++pen) {
      int run = 0;
      for (int i = 0; i <= g.nx; ++i) {
        run = std::min(std::max(run, cumcnt[(size_t)pen * (g.nx + 1) + i]), np2[pen]);
        cs[(size_t)pen * (g.nx + 1) + i] = poff[pen] + run;
      }
    }
    cs.back() = poff[g.npen];
    c->keys_valid = false;
    WM_CUDA(cudaMemcpyAsync(c->cs, cs.data(), cs.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    WM_CUDA(cudaMemcpyAsync(c->np2, np2, (size_t)g.npen * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    WM_CUDA(cudaMemcpyAsync(c->poff, poff.data(), poff.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    WM_CUDA(cudaStreamSynchronize(c->stream));  // cs/poff are stack-owned host vectors
    if (maxcnt > 0) {
      const size_t per_pen = (size_t)maxcnt * g.ndim;
      size_t want = std::min<size_t>(per_pen * g.npen, (size_t)32 << 20);
      want = std::max(want, per_pen);
      WM_TRY(reserve_stage(c, want));
      const int chunk = (int)std::max<size_t>(1, c->stage_elems / per_pen);
      for (int pen0 = 0; pen0 < g.npen; pen0 += chunk) {
        const int n = std::min(chunk, g.npen - pen0);
        WM_CUDA(cudaMemcpy2DAsync(c->stage, per_pen * sizeof(double), up + (size_t)pen0 * g.np * g.ndim,
                                  (size_t)g.np * g.ndim * sizeof(double), per_pen * sizeof(double), n,
                                  cudaMemcpyHostToDevice, c->stream));
        WM_TRY(wm_k_aos_to_soa(c, c->stage, c->A, c->id[c->cid], pen0, n, maxcnt));
      }
    }
    c->gp_valid = false;
    // The cell index must agree with the particles (cell = int(x), sort.f90:65): the kernels' re-binning moves a particle at most
    // one cell from the cell the index puts it in.  Every state the reference passes around satisfies this EXCEPT the shock driver's
    // freshly loaded box, whose cumcnt is nominal (n0 per cell, 2d/proj/shock/app.f90:346-359) while the particles are spread
    // evenly over one more cell (:436-441), so that some sit one cell below their nominal cell.  The reference evaluates its first
    // step about the nominal cells (and loses the charge of the particles that end up two cells away); here the index is repaired
    // on upload by one sort__bucket of the uploaded set, i.e. the first step already runs on consistent cells.
    int bad = 0;
    WM_TRY(wm_k_cells_consistent(c, &bad));
    if (c->nranks > 1 && c->nccl_comm) {           // the repair sort exchanges counts with the neighbours: decide it together
      double v = bad ? 1.0 : 0.0;
      WM_CUDA(cudaMemcpyAsync(c->red, &v, sizeof(double), cudaMemcpyHostToDevice, c->stream));
      WM_TRY(wm_comm_allreduce_sum(c, c->red, 1));
      WM_CUDA(cudaMemcpyAsync(&v, c->red, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      WM_CUDA(cudaStreamSynchronize(c->stream));
      bad = v > 0.0;
    }
    if (bad) {
      std::swap(c->A, c->B);                       // the uploaded set plays the pushed set of a sort
      c->gp_valid = true;
      WM_TRY(wm_k_classify(c, g.nxgs, g.nxge));
      WM_TRY(wm_k_sort(c, g.nxgs, g.nxge));
      c->gp_valid = false;
      c->keys_valid = false;
    }
  }
  WM_CUDA(cudaStreamSynchronize(c->stream));
  return WM_OK;
}

int wm_download(wm_ctx* c, double* up, int* np2, int* cumcnt, double* uf, double* gp) {
  if (!c) return WM_ERR_ARG;
  WM_CUDA(cudaSetDevice(c->device));
  const Geo& g = c->g;
  if (gp) {
    if (c->fused_done) {
      wm_set_error("wm_download: after field__fdtd_i took the push over (fused kernel) gp holds the wrapped, re-binned set; "
                   "read gp between particle__solv and field__fdtd_i, or call wm_set_fused(ctx, 0)");
      return WM_ERR_STATE;
    }
    WM_TRY(flush_deferred(c));   // somebody wants the pushed set itself: run the deferred push now
  }
  WM_TRY(wm_materialize(c));
  WM_TRY(check_flags(c));
  if (uf && c->swap_yz) {
    WM_TRY(need_swapbuf(c));
    WM_TRY(wm_k_swap_box6(c, c->uf, c->swapbuf, false));
    WM_CUDA(cudaMemcpyAsync(uf, c->swapbuf, g.nbox() * 6 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  } else if (uf) WM_CUDA(cudaMemcpyAsync(uf, c->uf, g.nbox() * 6 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  std::vector<int> h_np2(g.npen);
  WM_CUDA(cudaMemcpyAsync(h_np2.data(), c->np2, (size_t)g.npen * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  WM_CUDA(cudaStreamSynchronize(c->stream));
  int maxcnt = 0;
  for (int v : h_np2) maxcnt = std::max(maxcnt, v);
  if (np2) {
    if (c->swap_yz) permute_pencil_rows(c, h_np2.data(), np2, 1, false);
    else std::memcpy(np2, h_np2.data(), (size_t)g.npen * sizeof(int));
  }
  if (cumcnt) {
    std::vector<int> cs((size_t)g.npen * (g.nx + 1));
    WM_CUDA(cudaMemcpyAsync(cs.data(), c->cs, cs.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    WM_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<int> rel(c->swap_yz ? cs.size() : 0);
    int* dst = c->swap_yz ? rel.data() : cumcnt;
    for (int pen = 0; pen < g.npen; ++pen) {
      const int base = cs[(size_t)pen * (g.nx + 1)];
      for (int i = 0; i <= g.nx; ++i) dst[(size_t)pen * (g.nx + 1) + i] = cs[(size_t)pen * (g.nx + 1) + i] - base;
    }
    if (c->swap_yz) permute_pencil_rows(c, rel.data(), cumcnt, g.nx + 1, false);
  }
  for (int which = 0; which < 2; ++which) {
    double* dst = which == 0 ? up : gp;
    if (!dst || maxcnt == 0) continue;
    if (which == 1 && (!c->gp_valid || c->keys_valid)) {
      wm_set_error("wm_download: gp is only defined between particle__solv and bc__particle_yz");
      return WM_ERR_STATE;
    }
    const size_t per_pen = (size_t)maxcnt * g.ndim;
    size_t want = std::min<size_t>(per_pen * g.npen, (size_t)32 << 20);
    want = std::max(want, per_pen);
    WM_TRY(reserve_stage(c, want));
    const int chunk = (int)std::max<size_t>(1, c->stage_elems / per_pen);
    for (int pen0 = 0; pen0 < g.npen; pen0 += chunk) {
      const int n = std::min(chunk, g.npen - pen0);
      WM_TRY(wm_k_soa_to_aos(c, c->stage, which == 0 ? c->A : c->B, c->id[c->cid], pen0, n, maxcnt));
      // only the first np2 records of a pencil are defined; rows are copied to the longest pencil
      WM_CUDA(cudaMemcpy2DAsync(dst + (size_t)pen0 * g.np * g.ndim, (size_t)g.np * g.ndim * sizeof(double), c->stage,
                                per_pen * sizeof(double), per_pen * sizeof(double), n, cudaMemcpyDeviceToHost,
                                c->stream));
    }
  }
  WM_CUDA(cudaStreamSynchronize(c->stream));
  return WM_OK;
}

int wm_download_work(wm_ctx* c, int which, double* out) {
  if (!c || !out) return WM_ERR_ARG;
  WM_TRY(no_swap(c, "wm_download_work"));
  WM_CUDA(cudaSetDevice(c->device));
  const Geo& g = c->g;
  const size_t nb = g.nbox();
  if (which == 0) {
    WM_CUDA(cudaMemcpyAsync(out, c->uj, nb * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  } else if (which == 1) {
    WM_CUDA(cudaMemcpyAsync(out, c->df, nb * 6 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  } else if (which == 2) {
    // gkl(3, nxgs:nxge, nys:nye, nzs:nze): the device keeps it on the box layout
    std::vector<double> tmp(nb * 3);
    WM_CUDA(cudaMemcpyAsync(tmp.data(), c->gkl, nb * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    WM_CUDA(cudaStreamSynchronize(c->stream));
    size_t t = 0;
    for (int k = g.nzs; k <= g.nze; ++k)
      for (int j = g.nys; j <= g.nye; ++j)
        for (int i = g.nxgs; i <= g.nxge; ++i)
          for (int cc = 0; cc < 3; ++cc) out[t++] = tmp[g.box(i, j, k) * 3 + cc];
  } else {
    return WM_ERR_ARG;
  }
  WM_CUDA(cudaStreamSynchronize(c->stream));
  return WM_OK;
}

int wm_upload_work(wm_ctx* c, int which, const double* in) {
  if (!c || !in || which != 1) return WM_ERR_ARG;
  WM_CUDA(cudaSetDevice(c->device));
  if (c->swap_yz) {      // df is a field increment: it translates like uf
    WM_TRY(need_swapbuf(c));
    WM_CUDA(cudaMemcpyAsync(c->swapbuf, in, c->g.nbox() * 6 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    WM_TRY(wm_k_swap_box6(c, c->swapbuf, c->df, true));
  } else {
    WM_CUDA(cudaMemcpyAsync(c->df, in, c->g.nbox() * 6 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  WM_CUDA(cudaStreamSynchronize(c->stream));
  return WM_OK;
}

// ---------------------------------------------------------------------------------------------
// the hot path, one entry per reference procedure
// ---------------------------------------------------------------------------------------------
static int particle_solv_with(wm_ctx* c, int nxs, int nxe, int pusher) {
  if (!c || !range_ok(c, nxs, nxe)) { wm_set_error("Initialize first by calling particle__init()"); return WM_ERR_ARG; }
  WM_CUDA(cudaSetDevice(c->device));
  WM_TRY(wm_materialize(c));
  WM_TRY(wm_k_tmpf(c, nxs, nxe));
  const int saved = c->pusher;
  c->pusher = pusher;
  const int rc = wm_k_push(c, nxs, nxe);
  c->pusher = saved;
  WM_TRY(rc);
  c->gp_valid = true;
  c->keys_valid = false;
  c->fused_done = false;
  c->last_nxs = nxs;
  c->last_nxe = nxe;
  return WM_OK;
}

// run what was deferred with the per-procedure kernels (somebody wants gp itself, or the call order is not one of the three
// time loops the fused kernel covers)
static int flush_deferred(wm_ctx* c) {
  if (!c->defer_push) return WM_OK;
  c->defer_push = false;
  WM_TRY(particle_solv_with(c, c->defer_nxs, c->defer_nxe, c->defer_pusher));
  const int xbc = c->defer_xbc;
  c->defer_xbc = 0;
  if (xbc == WM_ORDER_RECONNECTION) WM_TRY(wm_k_bc_x(c, c->defer_nxs, c->defer_nxe, WM_BC_RECONNECTION, 0.0));
  if (xbc == WM_ORDER_SHOCK) WM_TRY(wm_k_bc_x(c, c->defer_nxs, c->defer_nxe, WM_BC_SHOCK, c->defer_u0));
  return WM_OK;
}

// particle__solv / particle__solv_vay on device-resident state: deferred when the fused kernel may take it over
static int particle_solv_entry(wm_ctx* c, int nxs, int nxe, int pusher) {
  if (!c || !range_ok(c, nxs, nxe)) { wm_set_error("Initialize first by calling particle__init()"); return WM_ERR_ARG; }
  WM_CUDA(cudaSetDevice(c->device));
  WM_TRY(flush_deferred(c));     // two pushes in a row: the first one really happens
  if (!c->use_fused || c->ntot == 0) return particle_solv_with(c, nxs, nxe, pusher);
  c->defer_push = true;
  c->defer_pusher = pusher;
  c->defer_nxs = nxs;
  c->defer_nxe = nxe;
  c->defer_xbc = 0;
  c->gp_valid = true;            // logically the pushed set exists from here on
  c->keys_valid = false;
  c->fused_done = false;
  c->last_nxs = nxs;
  c->last_nxe = nxe;
  return WM_OK;
}
int wm_particle_solv(wm_ctx* c, int nxs, int nxe) { return particle_solv_entry(c, nxs, nxe, WM_PUSHER_BORIS); }
int wm_particle_solv_vay(wm_ctx* c, int nxs, int nxe) { return particle_solv_entry(c, nxs, nxe, WM_PUSHER_VAY); }

int wm_field_stage(wm_ctx* c, int nxs, int nxe, int stage) {
  if (!c || !range_ok(c, nxs, nxe)) { wm_set_error("Initialize first by calling field__init()"); return WM_ERR_ARG; }
  WM_CUDA(cudaSetDevice(c->device));
  if (stage == 1) WM_TRY(flush_deferred(c));
  c->last_nxs = nxs;
  c->last_nxe = nxe;
  switch (stage) {
    case 1:
      if (!c->gp_valid || c->keys_valid || c->fused_done) {
        wm_set_error("field__fdtd_i needs the pushed particles: call it after particle__solv and before bc__particle_y[z]");
        return WM_ERR_STATE;
      }
      WM_TRY(wm_k_zero_uj(c, nxs, nxe));
      return wm_k_deposit(c, nxs, nxe);
    case 2: return wm_k_curre(c, nxs, nxe);
    case 3: return wm_k_gkl(c, nxs, nxe);
    case 4: return wm_k_cgm(c, nxs, nxe);
    case 5: return wm_k_dfield(c, nxs, nxe);
    case 6: return wm_k_de(c, nxs, nxe);
    case 7: return wm_k_dfield(c, nxs, nxe);
    case 8: return wm_k_update(c, nxs, nxe);
  }
  return WM_ERR_ARG;
}

// stages first..8 of field__fdtd_i; WM_FIELD_TIMING=1 (measurement aid) accumulates per-stage device times, printed by wm_destroy
static int field_stages(wm_ctx* c, int nxs, int nxe, int first) {
  static const bool timing = getenv("WM_FIELD_TIMING") != nullptr;
  if (!timing) {
    for (int s = first; s <= 8; ++s) WM_TRY(wm_field_stage(c, nxs, nxe, s));
    return WM_OK;
  }
  cudaEvent_t ev[9];   // created per call on the context's device (a measurement aid: the cost does not matter)
  for (auto& e : ev) cudaEventCreate(&e);
  cudaEventRecord(ev[first - 1], c->stream);
  for (int s = first; s <= 8; ++s) {
    WM_TRY(wm_field_stage(c, nxs, nxe, s));
    cudaEventRecord(ev[s], c->stream);
  }
  cudaEventSynchronize(ev[8]);
  for (int s = first; s <= 8; ++s) { float ms = 0; cudaEventElapsedTime(&ms, ev[s - 1], ev[s]); g_stage_ms[s] += ms; }
  for (auto& e : ev) cudaEventDestroy(e);
  g_stage_calls++;
  return WM_OK;
}
// K1 + ONE kernel for push + boundaries + deposit + destination counting (wm_fused.cu); the pending lazy permutation of the
// previous sort is consumed by it when it covered the same x range
static int fused_push_deposit(wm_ctx* c, int nxs, int nxe, int order, double u0) {
  if (c->lazy && (nxs != c->lazy_nxs || nxe != c->lazy_nxe)) WM_TRY(wm_materialize(c));
  WM_TRY(wm_k_tmpf(c, nxs, nxe));
  WM_TRY(wm_k_zero_uj(c, nxs, nxe));
  WM_TRY(wm_k_push_deposit_fused(c, nxs, nxe, order, u0));
  WM_CUDA(cudaEventRecord(c->ev_fused, c->stream));   // what the sort waits for (not for the field solve that follows)
  c->gp_valid = true;
  c->keys_valid = false;
  c->fused_done = true;
  c->fused_order = order;
  c->last_nxs = nxs;
  c->last_nxe = nxe;
  return WM_OK;
}

// The sort that follows a fused push kernel (lazy: its permutation stays pending).  With overlap on it is enqueued on the second
// stream behind the fused kernel only, so it runs concurrently with the field solve the caller enqueued on the main stream; the main
// stream then waits for it.  t0 / t1: optional timing events recorded around it on the stream it runs on.
static int sort_after_fused(wm_ctx* c, int nxs, int nxe, cudaEvent_t t0, cudaEvent_t t1) {
  const bool ov = wm_overlap_bps(c, (long long)(nxe - nxs + 1) * c->g.nyl * c->g.nzl) > 0;
  cudaStream_t main_stream = c->stream;
  void* main_comm = c->nccl_comm;
  if (ov) {
    c->stream = c->stream2;
    if (c->nccl_comm2) c->nccl_comm = c->nccl_comm2;
    cudaStreamWaitEvent(c->stream, c->ev_fused, 0);
  }
  if (t0) cudaEventRecord(t0, c->stream);
  c->allow_lazy = 1;      // the next fused kernel (or wm_materialize) applies the permutation
  const int rc = wm_k_sort(c, nxs, nxe);
  c->allow_lazy = 0;
  if (t1) cudaEventRecord(t1, c->stream);
  if (ov) {
    cudaEventRecord(c->ev_sort, c->stream);
    c->stream = main_stream;
    c->nccl_comm = main_comm;
    cudaStreamWaitEvent(c->stream, c->ev_sort, 0);
  }
  return rc;
}

int wm_field_fdtd_i(wm_ctx* c, int nxs, int nxe) {
  if (!c || !range_ok(c, nxs, nxe)) { wm_set_error("Initialize first by calling field__init()"); return WM_ERR_ARG; }
  WM_CUDA(cudaSetDevice(c->device));
  if (c->defer_push) {
    // which of the reference's three time loops this call sequence is: nothing between solv and fdtd_i (Weibel / beam,
    // 3d/proj/weibel/app.f90:100-108), the reflecting bc__particle_x (reconnection :103-108) or bc__injection (shock)
    const int order = c->defer_xbc;
    const int saved = c->pusher;
    if (nxs == c->defer_nxs && nxe == c->defer_nxe && wm_fused_supported(c, order)) {
      c->defer_push = false;
      c->defer_xbc = 0;
      c->pusher = c->defer_pusher;
      const int rc = fused_push_deposit(c, nxs, nxe, order, c->defer_u0);
      c->pusher = saved;
      WM_TRY(rc);
      return field_stages(c, nxs, nxe, 2);
    }
  }
  return field_stages(c, nxs, nxe, 1);
}

int wm_bc_particle_x(wm_ctx* c, int nxs, int nxe) {
  if (!c) return WM_ERR_ARG;
  WM_CUDA(cudaSetDevice(c->device));
  if (!c->gp_valid) { wm_set_error("bc__particle_x acts on the pushed particles (call particle__solv first)"); return WM_ERR_STATE; }
  if (c->fused_done) {
    // Weibel loop: the periodic wrap was applied by the fused kernel after its deposit (ORDER 0); in the wall loops the fused
    // kernel already reflected before the deposit and a second call of the driver would be a no-op in the reference too
    // (a reflected particle lies inside the walls)
    return WM_OK;
  }
  if (c->defer_push && c->defer_xbc == 0 && c->g.bc != WM_BC_PERIODIC && nxs == c->defer_nxs && nxe == c->defer_nxe) {
    c->defer_xbc = WM_ORDER_RECONNECTION;    // reflecting walls before the field step: the reconnection loop
    return WM_OK;
  }
  WM_TRY(flush_deferred(c));
  // boundary_shock__particle_x is the same reflecting-wall rule as boundary_reconnection__particle_x
  return wm_k_bc_x(c, nxs, nxe, c->g.bc == WM_BC_PERIODIC ? WM_BC_PERIODIC : WM_BC_RECONNECTION, 0.0);
}

int wm_bc_injection(wm_ctx* c, int nxs, int nxe, double u0) {
  if (!c) return WM_ERR_ARG;
  WM_CUDA(cudaSetDevice(c->device));
  if (!c->gp_valid) { wm_set_error("bc__injection acts on the pushed particles"); return WM_ERR_STATE; }
  if (c->fused_done) { wm_set_error("bc__injection after field__fdtd_i: not a call order of the reference"); return WM_ERR_STATE; }
  if (c->defer_push && c->defer_xbc == 0 && c->g.bc == WM_BC_SHOCK && nxs == c->defer_nxs && nxe == c->defer_nxe) {
    c->defer_xbc = WM_ORDER_SHOCK;
    c->defer_u0 = u0;
    return WM_OK;
  }
  WM_TRY(flush_deferred(c));
  return wm_k_bc_x(c, nxs, nxe, WM_BC_SHOCK, u0);
}

int wm_bc_particle_yz(wm_ctx* c) {
  if (!c) return WM_ERR_ARG;
  WM_CUDA(cudaSetDevice(c->device));
  if (!c->gp_valid) { wm_set_error("bc__particle_yz acts on the pushed particles"); return WM_ERR_STATE; }
  if (c->fused_done) { c->keys_valid = true; return WM_OK; }   // wrapped and counted by the fused kernel
  WM_TRY(flush_deferred(c));
  // re-binning is classification here; the movers travel inside sort__bucket's scatter (wm_sort.cu)
  WM_TRY(wm_k_classify(c, c->last_nxs, c->last_nxe));
  c->keys_valid = true;
  return WM_OK;
}

int wm_sort_bucket(wm_ctx* c, int nxs, int nxe) {
  if (!c || !range_ok(c, nxs, nxe)) { wm_set_error("Initialize first by calling sort__init()"); return WM_ERR_ARG; }
  WM_CUDA(cudaSetDevice(c->device));
  if (!c->gp_valid) { wm_set_error("sort__bucket sorts the pushed particles"); return WM_ERR_STATE; }
  WM_TRY(flush_deferred(c));
  if (c->fused_done) {
    // the permutation stays pending for the next fused kernel (lazy sort), exactly as inside wm_step; the field solve this
    // driver enqueued with field__fdtd_i may still be running on the main stream: the sort overlaps it
    WM_TRY(sort_after_fused(c, nxs, nxe, nullptr, nullptr));
  } else {
    if (!c->keys_valid) WM_TRY(wm_k_classify(c, nxs, nxe));
    WM_TRY(wm_k_sort(c, nxs, nxe));
  }
  c->gp_valid = false;
  c->keys_valid = false;
  c->fused_done = false;
  return WM_OK;
}

// fold the recorded (not yet read) phase events into the sums: the only host synchronisation of the timing mode
static int fold_timing(wm_ctx* c) {
  if (c->ev_used == 0) return WM_OK;
  constexpr int P = wm_ctx::EV_PER;
  WM_CUDA(cudaEventSynchronize(c->ev[P * (c->ev_used - 1) + 4]));   // the step end: recorded after the main stream joined the sort
  for (int s = 0; s < c->ev_used; ++s) {
    for (int e = 0; e < 3; ++e) cudaEventElapsedTime(&c->ms_phase[e], c->ev[P * s + e], c->ev[P * s + e + 1]);
    cudaEventElapsedTime(&c->ms_phase[3], c->ev[P * s + 5], c->ev[P * s + 6]);   // the sort, on whichever stream it ran
    for (int e = 0; e < 4; ++e) c->ms_sum[e] += c->ms_phase[e];
    c->timed_steps++;
  }
  c->ev_used = 0;
  return WM_OK;
}

int wm_step(wm_ctx* c, int nxs, int nxe, int order, double u0, int nsteps) {
  if (!c || !range_ok(c, nxs, nxe)) return WM_ERR_ARG;
  WM_CUDA(cudaSetDevice(c->device));
  WM_TRY(flush_deferred(c));
  const bool fused = c->use_fused && wm_fused_supported(c, order);
  // a pending lazy sort is consumed by the fused kernel only if it covered the same x range
  if (c->lazy && (!fused || nxs != c->lazy_nxs || nxe != c->lazy_nxe)) WM_TRY(wm_materialize(c));
  for (int it = 0; it < nsteps; ++it) {
    if (c->timing && c->ev_used == wm_ctx::EV_STEPS) WM_TRY(fold_timing(c));
    cudaEvent_t* ev = c->ev + wm_ctx::EV_PER * c->ev_used;
    if (c->timing) WM_CUDA(cudaEventRecord(ev[0], c->stream));
    if (fused) {
      // K1, then ONE kernel for push + boundaries + deposit + destination counting (wm_fused.cu)
      WM_TRY(fused_push_deposit(c, nxs, nxe, order, u0));
      if (c->timing) { WM_CUDA(cudaEventRecord(ev[1], c->stream)); WM_CUDA(cudaEventRecord(ev[2], c->stream)); }
      WM_TRY(field_stages(c, nxs, nxe, 2));
      if (c->timing) WM_CUDA(cudaEventRecord(ev[3], c->stream));
      WM_TRY(sort_after_fused(c, nxs, nxe, c->timing ? ev[5] : nullptr, c->timing ? ev[6] : nullptr));
      c->gp_valid = false;
      c->keys_valid = false;
      c->fused_done = false;
    } else {
      WM_TRY(particle_solv_with(c, nxs, nxe, c->pusher));
      if (c->timing) WM_CUDA(cudaEventRecord(ev[1], c->stream));
      if (order == WM_ORDER_RECONNECTION) WM_TRY(wm_bc_particle_x(c, nxs, nxe));
      if (order == WM_ORDER_SHOCK) WM_TRY(wm_bc_injection(c, nxs, nxe, u0));
      WM_TRY(wm_field_stage(c, nxs, nxe, 1));
      if (c->timing) WM_CUDA(cudaEventRecord(ev[2], c->stream));
      WM_TRY(field_stages(c, nxs, nxe, 2));
      if (c->timing) WM_CUDA(cudaEventRecord(ev[3], c->stream));
      if (c->timing) WM_CUDA(cudaEventRecord(ev[5], c->stream));
      if (order == WM_ORDER_WEIBEL) WM_TRY(wm_bc_particle_x(c, nxs, nxe));
      WM_TRY(wm_bc_particle_yz(c));
      WM_TRY(wm_sort_bucket(c, nxs, nxe));
      if (c->timing) WM_CUDA(cudaEventRecord(ev[6], c->stream));
    }
    if (c->timing) {
      WM_CUDA(cudaEventRecord(ev[4], c->stream));
      c->ev_used++;
    }
  }
  return WM_OK;
}

}  // extern "C"

// called by wm_comm_init once the communicator exists: the slab-axis neighbours are other ranks from now on
int wm_enable_slab_migration(wm_ctx* c) {
  Geo& g = c->g;
  g.multi = 1;
  g.nrows = g.npen + 2 * g.nsp * g.ngrow;
  return WM_OK;
}

extern "C" {

int wm_set_fused(wm_ctx* c, int on) {
  if (!c) return WM_ERR_ARG;
  c->use_fused = on;
  return WM_OK;
}

int wm_settle(wm_ctx* c) {
  if (!c) return WM_ERR_ARG;
  WM_CUDA(cudaSetDevice(c->device));
  return wm_materialize(c);
}

int wm_set_pusher(wm_ctx* c, int pusher) {
  if (!c || (pusher != WM_PUSHER_BORIS && pusher != WM_PUSHER_VAY)) { wm_set_error("wm_set_pusher: WM_PUSHER_BORIS or WM_PUSHER_VAY"); return WM_ERR_ARG; }
  c->pusher = pusher;
  return WM_OK;
}

// ---------------------------------------------------------------------------------------------
// host-buffer forms
// ---------------------------------------------------------------------------------------------
int wm_h_particle_solv(wm_ctx* c, double* gp, const double* up, const double* uf, const int* cumcnt, const int* np2,
                       int nxs, int nxe) {
  WM_TRY(wm_upload(c, up, np2, cumcnt, uf));
  WM_TRY(wm_particle_solv(c, nxs, nxe));
  return wm_download(c, nullptr, nullptr, nullptr, nullptr, gp);
}

int wm_h_particle_solv_vay(wm_ctx* c, double* gp, const double* up, const double* uf, const int* cumcnt, const int* np2,
                           int nxs, int nxe) {
  WM_TRY(wm_upload(c, up, np2, cumcnt, uf));
  WM_TRY(wm_particle_solv_vay(c, nxs, nxe));
  return wm_download(c, nullptr, nullptr, nullptr, nullptr, gp);
}

int wm_h_field_fdtd_i(wm_ctx* c, double* uf, const double* up, const double* gp, const int* cumcnt, const int* np2,
                      int nxs, int nxe) {
  if (!c) return WM_ERR_ARG;
  WM_TRY(wm_upload(c, up, np2, cumcnt, uf));
  // gp arrives from the host as well: stage it into set B with the same pencil offsets
  {
    const Geo& g = c->g;
    int maxcnt = 0;
    for (int pen = 0; pen < g.npen; ++pen) maxcnt = std::max(maxcnt, np2[pen]);
    if (maxcnt > 0) {
      const size_t per_pen = (size_t)maxcnt * g.ndim;
      const int chunk = (int)std::max<size_t>(1, c->stage_elems / per_pen);
      for (int pen0 = 0; pen0 < g.npen; pen0 += chunk) {
        const int n = std::min(chunk, g.npen - pen0);
        WM_CUDA(cudaMemcpy2DAsync(c->stage, per_pen * sizeof(double), gp + (size_t)pen0 * g.np * g.ndim,
                                  (size_t)g.np * g.ndim * sizeof(double), per_pen * sizeof(double), n,
                                  cudaMemcpyHostToDevice, c->stream));
        // the ID column of gp equals up's (particle.f90:227-231); it is written to the spare ID array
        WM_TRY(wm_k_aos_to_soa(c, c->stage, c->B, c->id[1 - c->cid], pen0, n, maxcnt));
      }
    }
    c->gp_valid = true;
  }
  WM_TRY(wm_field_fdtd_i(c, nxs, nxe));
  return wm_download(c, nullptr, nullptr, nullptr, uf, nullptr);
}

int wm_h_step(wm_ctx* c, double* up, double* uf, int* np2, int* cumcnt, int nxs, int nxe, int order, double u0) {
  WM_TRY(wm_upload(c, up, np2, cumcnt, uf));
  WM_TRY(wm_step(c, nxs, nxe, order, u0, 1));
  return wm_download(c, up, np2, cumcnt, uf, nullptr);
}

// ---------------------------------------------------------------------------------------------
// diagnostics
// ---------------------------------------------------------------------------------------------
int wm_load_weibel(wm_ctx* c, int n0, double v_thi, double v_the, double t_ani, double b0, unsigned long long seed) {
  if (!c || n0 <= 0) return WM_ERR_ARG;
  WM_CUDA(cudaSetDevice(c->device));
  c->lazy = false;
  c->defer_push = false;
  c->defer_xbc = 0;
  c->fused_done = false;
  const Geo& g = c->g;
  if ((long long)n0 * g.nx > g.np) {
    wm_set_error("Error: Too large number of particles");  // 3d/proj/weibel/app.f90:315-320
    return WM_ERR_MEMORY_OVER;
  }
  c->ntot = (long long)n0 * g.nx * g.npen;
  c->n_sp0 = c->ntot / g.nsp;
  WM_TRY(reserve_particles(c, (size_t)c->ntot));
  WM_TRY(wm_k_load_weibel(c, n0, v_thi, v_the, t_ani, b0, seed));
  c->gp_valid = false;
  return WM_OK;
}

// mom_calc__accl + mom_calc__nvt + bc__mom of the drivers' output block (3d/proj/weibel/app.f90:121-124), on the sorted
// device-resident particles; only the (7, nx+2, nyl+2, [nzl+2,] nsp) moment array crosses to the host.
int wm_mom_calc(wm_ctx* c, int nxs, int nxe, double* mom) {
  if (!c || !mom || !range_ok(c, nxs, nxe)) { wm_set_error("Initialize first by calling mom_calc__init()"); return WM_ERR_ARG; }
  WM_CUDA(cudaSetDevice(c->device));
  const Geo& g = c->g;
  if (c->gp_valid) { wm_set_error("mom_calc acts on the sorted particles (call it after sort__bucket)"); return WM_ERR_STATE; }
  WM_TRY(wm_materialize(c));
  const size_t nb = g.nbox(), nel = nb * 7 * g.nsp;
  if (!c->mom) WM_CUDA(cudaMalloc(&c->mom, nel * sizeof(double)));
  WM_CUDA(cudaMemsetAsync(c->mom, 0, nel * sizeof(double), c->stream));
  WM_TRY(wm_k_tmpf(c, nxs, nxe));
  if (c->ntot > 0) WM_TRY(wm_k_mom(c, nxs, nxe));
  WM_TRY(wm_k_mom_fold(c));
  std::vector<double> tmp(nel);
  WM_CUDA(cudaMemcpyAsync(tmp.data(), c->mom, nel * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  WM_CUDA(cudaStreamSynchronize(c->stream));
  size_t t = 0;
  if (c->swap_yz) {
    // caller's mom(7, x, y, z, nsp) from the device's (x, y' = z, z' = y): V = (Vx, Vz', Vy'), T = (Txx, Tz'z', Ty'y')
    static const int lmap[7] = {0, 1, 3, 2, 4, 6, 5};
    for (int isp = 0; isp < g.nsp; ++isp)
      for (int kh = g.nys - 1; kh <= g.nye + 1; ++kh)          // the caller's z range is the device's y' range
        for (int jh = g.nzs - 1; jh <= g.nze + 1; ++jh)
          for (int i = g.nxgs - 1; i <= g.nxge + 1; ++i) {
            const double* s = &tmp[((size_t)isp * nb + g.box(i, kh, jh)) * 7];
            for (int l = 0; l < 7; ++l) mom[t++] = s[lmap[l]];
          }
    return check_flags(c);
  }
  const int k0 = g.dim == 3 ? g.nzs - 1 : 0, k1 = g.dim == 3 ? g.nze + 1 : 0;
  for (int isp = 0; isp < g.nsp; ++isp)
    for (int k = k0; k <= k1; ++k)
      for (int j = g.nys - 1; j <= g.nye + 1; ++j)
        for (int i = g.nxgs - 1; i <= g.nxge + 1; ++i) {
          const double* s = &tmp[((size_t)isp * nb + g.box(i, j, k)) * 7];
          for (int l = 0; l < 7; ++l) mom[t++] = s[l];
        }
  return check_flags(c);
}

int wm_energy(wm_ctx* c, double* out) {
  if (!c || !out) return WM_ERR_ARG;
  WM_CUDA(cudaSetDevice(c->device));
  WM_TRY(wm_materialize(c));
  return wm_k_energy(c, out);
}

int wm_gauss(wm_ctx* c, double* out) {
  if (!c || !out) return WM_ERR_ARG;
  WM_CUDA(cudaSetDevice(c->device));
  WM_TRY(wm_materialize(c));
  return wm_k_gauss(c, out);
}

int wm_get_stats(wm_ctx* c, wm_stats* out) {
  if (!c || !out) return WM_ERR_ARG;
  WM_CUDA(cudaSetDevice(c->device));
  const Geo& g = c->g;
  std::vector<int> h(g.npen);
  int f = 0;
  WM_CUDA(cudaMemcpyAsync(h.data(), c->np2, (size_t)g.npen * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  WM_CUDA(cudaMemcpyAsync(&f, c->flags, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (c->cg_ite_on_device)
    WM_CUDA(cudaMemcpyAsync(c->cg_ite, c->totals + 6, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  WM_CUDA(cudaStreamSynchronize(c->stream));
  WM_TRY(fold_timing(c));
  for (int l = 0; l < 3; ++l) out->cg_iterations[l] = c->cg_ite[l];
  out->n_particles = c->ntot;
  out->max_np2 = 0;
  for (int v : h) out->max_np2 = std::max(out->max_np2, v);
  out->error_flags = f;
  out->timed_steps = c->timed_steps;
  out->ms_push = c->ms_sum[0];
  out->ms_deposit = c->ms_sum[1];
  out->ms_field = c->ms_sum[2];
  out->ms_sort = c->ms_sum[3];
  return WM_OK;
}

int wm_sync(wm_ctx* c) {
  if (!c) return WM_ERR_ARG;
  WM_CUDA(cudaSetDevice(c->device));
  WM_CUDA(cudaStreamSynchronize(c->stream));
  return check_flags(c);
}

int wm_set_timing(wm_ctx* c, int on) {
  if (!c) return WM_ERR_ARG;
  c->timing = on;
  c->ev_used = 0;
  for (int e = 0; e < 4; ++e) c->ms_sum[e] = 0;
  c->timed_steps = 0;
  return WM_OK;
}

long long wm_launch_count(wm_ctx* c) { return c ? c->launches : 0; }
void* wm_stream(wm_ctx* c) { return c ? (void*)c->stream : nullptr; }

}  // extern "C"
