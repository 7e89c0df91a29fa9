// wm_shock.cu -- the particle source of the shock set-up on the device: the driver's inject() and
// relocate() (2d/proj/shock/app.f90:697-852, 615-692; 3d/proj/shock/app.f90:733-906, 644-728) mutate up / np2 / cumcnt /
// uf on the host after EVERY step, which would force a full state round trip per step.  Here only the integer
// bookkeeping stays with the caller (how many particles each row receives: app.f90:711-774, and the ID offsets of
// get_global_cumsum :769-781); the particles are created in HBM.
//
// The particle store is packed (no per-pencil slack), so appending n(row) particles to every pencil is one shifted
// copy of the sorted set A into the free set B (an HBM-rate pass, 2 x ndim x 8 B per resident particle), the new
// particles are written behind their pencils, and the cell index moves along:
//   cs'(pen, kk) = cs(pen, kk) + shift(pen) + (kk >= nxe - nxgs ? n(pen) : 0),   shift = exclusive prefix of n over pencils.
// The reference bumps cumcnt(nxe) only and leaves the entries above it stale (they are never read: the loops stop at
// nxe); the monotone form above is what wm_download returns for those entries.
//
// Random numbers: Fortran's random_number is not reproducible (utils/wuming_utils.f90:48-53); positions and momenta come
// from Philox-4x32-10 keyed by (seed; global row, particle index in the row, purpose, step), the convention of the
// Weibel loader (wm_diag.cu), so that the oracle restatement and this kernel draw the same numbers on any slab count.
#include "wm_internal.cuh"

#include <algorithm>
#include <vector>

namespace {

constexpr int TPB = 256;
constexpr double kPi = 3.14159265358979323846264338327950288;

__device__ inline void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ inline void uniform2(unsigned long long seed, uint32_t stream, uint32_t idx, uint32_t purpose, uint32_t epoch,
                                double& u0, double& u1) {
  uint32_t o[4];
  philox(idx, purpose, stream, epoch, (uint32_t)seed, (uint32_t)(seed >> 32), o);
  uint64_t a = ((uint64_t)o[1] << 32) | o[0], b = ((uint64_t)o[3] << 32) | o[2];
  u0 = (double)(a >> 11) * (1.0 / 9007199254740992.0);
  u1 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
}
// utils/wuming_utils.f90:83-86
__device__ inline void box_muller(double x1, double x2, double& ns, double& nc) {
  double rr = sqrt(-2.0 * log(1.0 - x1) + 1.0e-30);
  ns = rr * sin(2.0 * kPi * x2);
  nc = rr * cos(2.0 * kPi * x2);
}

// per pencil: n(pen) = cnt[row(pen)] (inject) or n0 (relocate); shift = exclusive prefix over pencils.  One block; the
// pencil count is at most a few 10^4.
__global__ void k_shock_shift(Geo g, const int* __restrict__ cnt_rows, int n_fixed, int* __restrict__ n_pen, int* __restrict__ shift) {
  __shared__ int part[TPB];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int nrow = g.nyl * g.nzl;
  for (int base = 0; base <= g.npen; base += TPB) {
    const int pen = base + threadIdx.x;
    int n = 0;
    if (pen < g.npen) n = cnt_rows ? cnt_rows[pen % nrow] : n_fixed;
    part[threadIdx.x] = n;
    __syncthreads();
    for (int o = 1; o < TPB; o <<= 1) {   // Hillis-Steele inclusive scan of the tile
      int v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
      __syncthreads();
      part[threadIdx.x] += v;
      __syncthreads();
    }
    if (pen <= g.npen) {
      shift[pen] = carry + part[threadIdx.x] - n;
      if (pen < g.npen) n_pen[pen] = n;
    }
    __syncthreads();
    if (threadIdx.x == TPB - 1) carry += part[TPB - 1];
    __syncthreads();
  }
}

// shifted copy of the sorted set: pencil pen moves from cs(pen,0) to cs(pen,0) + shift(pen); one block per pencil
__global__ void k_shock_move(Geo g, Ptcl A, Ptcl B, const double* __restrict__ id_in, double* __restrict__ id_out,
                             const int* __restrict__ cs, const int* __restrict__ shift) {
  const int ncomp = g.ndim - 1;
  for (int pen = blockIdx.x; pen < g.npen; pen += gridDim.x) {
    const int* row = cs + (size_t)pen * (g.nx + 1);
    const int beg = row[0], end = row[g.nx], sh = shift[pen];
    for (int p = beg + threadIdx.x; p < end; p += blockDim.x) {
      for (int c = 0; c < ncomp; ++c) B.c[c][p + sh] = A.c[c][p];
      id_out[p + sh] = id_in[p];
    }
  }
}

// new cell index + flags: particles beyond the injection boundary would be overwritten (there are none after
// boundary_shock__injection); np2 over the host pencil capacity = the reference's "memory over"
__global__ void k_shock_index(Geo g, const int* __restrict__ cs, int* __restrict__ cs_new, const int* __restrict__ n_pen,
                              const int* __restrict__ shift, int kk0, int* flags) {
  const int w = g.nx + 1;
  const long long n = (long long)g.npen * w + 1;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    if (e == n - 1) { cs_new[e] = cs[e] + shift[g.npen]; continue; }
    const int pen = (int)(e / w), kk = (int)(e % w);
    cs_new[e] = cs[e] + shift[pen] + (kk >= kk0 ? n_pen[pen] : 0);
    if (kk == g.nx) {
      if (cs[e] != cs[(size_t)pen * w + kk0]) atomicOr(flags, 2);
      if (cs[e] - cs[(size_t)pen * w] + n_pen[pen] > g.np) atomicOr(flags, 1);
    }
  }
}

// the new particles.  RELOC = false: inject (2d/proj/shock/app.f90:786-841, 3d :815-879); true: relocate (2d :637-676, 3d :664-711)
template <int D, bool RELOC>
__global__ void k_shock_fill(Geo g, Ptcl B, double* __restrict__ id_out, const int* __restrict__ cs_new, const int* __restrict__ n_pen,
                             const long long* __restrict__ id_first, wm_shock_params sp, int nxe, uint32_t epoch) {
  const int nrow = g.nyl * g.nzl;
  const uint32_t base = RELOC ? 16u : 0u;
  const double x0 = fabs(sp.v0) * g.delt;
  const double xd0 = sp.l_damp_ini + g.nxgs * g.delx, xds = sp.l_damp_ini * 0.1;   // vprofile, 2d/proj/shock/app.f90:883-893
  constexpr int U = D;
  for (int pen = blockIdx.x; pen < g.npen; pen += gridDim.x) {
    const int n = n_pen[pen];
    const int isp = pen / nrow, rl = pen % nrow;
    const int j = g.nys + rl % g.nyl, k = D == 3 ? g.nzs + rl / g.nyl : 0;
    const uint32_t row = (uint32_t)((j - g.nygs) + (D == 3 ? g.ny * (k - g.nzgs) : 0));
    const int first = cs_new[(size_t)pen * (g.nx + 1) + g.nx] - n;   // the new particles sit at the end of the pencil
    const double sd = isp == 0 ? sp.v_thi : sp.v_the;
    for (int ii = 1 + threadIdx.x; ii <= n; ii += blockDim.x) {
      double ur0, ur1, a0, a1, b0, b1, ns, nc, ms, mc;
      uniform2(sp.seed, row, (uint32_t)ii, base, epoch, ur0, ur1);
      uniform2(sp.seed, row, (uint32_t)ii, base + (uint32_t)(2 * isp + 1), epoch, a0, a1);
      uniform2(sp.seed, row, (uint32_t)ii, base + (uint32_t)(2 * isp + 2), epoch, b0, b1);
      box_muller(a0, a1, ns, nc);
      box_muller(b0, b1, ms, mc);
      double x = RELOC ? (nxe - 1) * g.delx + (ii - 5e-1) / n * g.delx : nxe * g.delx + (ii - 5e-1) / n * x0;
      double ux = sd * ns;
      const double uy = sd * nc, uz = sd * ms;
      if (!RELOC) x = x + (sp.v0 + ux) * g.delt;   // injection (non-relativistic approximation)
      const double v1 = 0.5 * sp.v0 * (1 + tanh((x - xd0) / xds));
      const double gam1 = 1.0 / sqrt(1.0 - (v1 / g.c) * (v1 / g.c));
      const double gamp = sqrt(1.0 + (ux * ux + uy * uy + uz * uz) / (g.c * g.c));
      ux = gam1 * (ux + v1 * gamp);
      const size_t d = (size_t)first + (ii - 1);
      B.c[0][d] = x;
      B.c[1][d] = (j + ur0) * g.delx;
      if (D == 3) B.c[2][d] = (k + ur1) * g.delx;
      B.c[U][d] = ux;
      B.c[U + 1][d] = uy;
      B.c[U + 2][d] = uz;
      const long long pid = (long long)ii + id_first[pen];
      id_out[d] = __longlong_as_double(-pid);
    }
  }
}

// upstream field columns nxe-1, nxe (2d/proj/shock/app.f90:680-689 = 840-849; 3d :715-726 = 893-904), ghosts included
__global__ void k_shock_field(Geo g, double* __restrict__ uf, wm_shock_params sp, int nxe) {
  const double by = sp.b0 * sin(sp.theta_bn) * cos(sp.phi_bn), bz = sp.b0 * sin(sp.theta_bn) * sin(sp.phi_bn);
  const int n = g.by * g.bz;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int j = g.nys - 2 + e % g.by, k = g.dim == 3 ? g.nzs - 2 + e / g.by : 0;
    double* f1 = uf + g.box(nxe - 1, j, k) * 6;
    double* f2 = uf + g.box(nxe, j, k) * 6;
    f1[1] = by;
    f1[2] = bz;
    f1[4] = +sp.v0 * bz / g.c;
    f1[5] = -sp.v0 * by / g.c;
    f2[1] = by;
    f2[2] = bz;
  }
}

// room for `need` particles, keeping the sorted set A and its IDs
int grow_particles(wm_ctx* c, size_t need) {
  if (need <= c->cap) return WM_OK;
  const size_t cap = need + need / 2 + 1024;
  const int ncomp = c->g.ndim - 1;
  for (int k = 0; k < ncomp; ++k) {
    double* a = nullptr;
    WM_CUDA(cudaMalloc(&a, cap * sizeof(double)));
    WM_CUDA(cudaMemcpyAsync(a, c->A.c[k], (size_t)c->ntot * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    WM_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(c->A.c[k]);
    cudaFree(c->B.c[k]);
    c->A.c[k] = a;
    WM_CUDA(cudaMalloc(&c->B.c[k], cap * sizeof(double)));
  }
  double* idn = nullptr;
  WM_CUDA(cudaMalloc(&idn, cap * sizeof(double)));
  WM_CUDA(cudaMemcpyAsync(idn, c->id[c->cid], (size_t)c->ntot * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  WM_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(c->id[0]);
  cudaFree(c->id[1]);
  c->id[c->cid] = idn;
  WM_CUDA(cudaMalloc(&c->id[1 - c->cid], cap * sizeof(double)));
  c->cap = cap;
  return WM_OK;
}

int shock_source(wm_ctx* c, const wm_shock_params* sp, int nxe, const int* cnt_rows_host, const long long* id_first_host,
                 long long epoch, bool reloc) {
  const Geo& g = c->g;
  if (g.bc != WM_BC_SHOCK) { wm_set_error("inject / relocate belong to the shock set-up (bc_kind = WM_BC_SHOCK)"); return WM_ERR_ARG; }
  if (nxe < g.nxgs + 1 || nxe > g.nxge) { wm_set_error("inject / relocate: nxe outside the box"); return WM_ERR_ARG; }
  if (c->gp_valid) { wm_set_error("inject / relocate act on the sorted particles: call them after sort__bucket"); return WM_ERR_STATE; }
  WM_CUDA(cudaSetDevice(c->device));
  WM_TRY(wm_materialize(c));   // the shifted copy below moves the cell-sorted set
  cudaStream_t st = c->stream;
  const int nrow = g.nyl * g.nzl;
  long long added_sp = 0;   // per species
  if (reloc) added_sp = (long long)sp->n0 * nrow;
  else for (int r = 0; r < nrow; ++r) {
    if (cnt_rows_host[r] < 0) { wm_set_error("inject: negative row count"); return WM_ERR_ARG; }
    added_sp += cnt_rows_host[r];
  }
  {
    // slab runs keep the head-room of reserve_particles (wm_api.cu): the sort parks the outgoing ghost rows behind the
    // local particles and receives the arrivals into the other set
    const size_t need = (size_t)(c->ntot + g.nsp * added_sp);
    WM_TRY(grow_particles(c, c->nranks > 1 ? need + need / 4 + 1024 : need));
  }
  // small device-side tables: [row counts | n per pencil | shift per pencil (+ total)] and the ID offsets
  int* tab = nullptr;
  long long* idf = nullptr;
  WM_CUDA(cudaMalloc(&tab, ((size_t)nrow + 2 * (size_t)g.npen + 2) * sizeof(int)));
  WM_CUDA(cudaMalloc(&idf, (size_t)g.npen * sizeof(long long)));
  int* cnt_rows = tab;
  int* n_pen = tab + nrow;
  int* shift = n_pen + g.npen;
  if (!reloc) WM_CUDA(cudaMemcpyAsync(cnt_rows, cnt_rows_host, (size_t)nrow * sizeof(int), cudaMemcpyHostToDevice, st));
  WM_CUDA(cudaMemcpyAsync(idf, id_first_host, (size_t)g.npen * sizeof(long long), cudaMemcpyHostToDevice, st));
  k_shock_shift<<<1, TPB, 0, st>>>(g, reloc ? nullptr : cnt_rows, sp->n0, n_pen, shift);
  WM_LAUNCH_CHECK(c);
  const int old_cid = c->cid;
  const int blocks = std::min(g.npen, 148 * 8);
  k_shock_move<<<blocks, TPB, 0, st>>>(g, c->A, c->B, c->id[old_cid], c->id[1 - old_cid], c->cs, shift);
  WM_LAUNCH_CHECK(c);
  const long long ncs = (long long)g.npen * (g.nx + 1) + 1;
  k_shock_index<<<wm_blocks(ncs, TPB), TPB, 0, st>>>(g, c->cs, c->cs_new, n_pen, shift, nxe - g.nxgs, c->flags);
  WM_LAUNCH_CHECK(c);
  const uint32_t ep = (uint32_t)epoch;
#define WM_FILL(D, R) k_shock_fill<D, R><<<blocks, 128, 0, st>>>(g, c->B, c->id[1 - old_cid], c->cs_new, n_pen, idf, *sp, nxe, ep)
  if (g.dim == 3) { if (reloc) WM_FILL(3, true); else WM_FILL(3, false); }
  else { if (reloc) WM_FILL(2, true); else WM_FILL(2, false); }
#undef WM_FILL
  WM_LAUNCH_CHECK(c);
  k_shock_field<<<wm_blocks((long long)g.by * g.bz, TPB), TPB, 0, st>>>(g, c->uf, *sp, nxe);
  WM_LAUNCH_CHECK(c);
  std::swap(c->A, c->B);
  std::swap(c->cs, c->cs_new);
  c->cid = 1 - old_cid;
  c->ntot += g.nsp * added_sp;
  c->n_sp0 += added_sp;
  WM_TRY(wm_k_refresh_np2(c));
  WM_CUDA(cudaStreamSynchronize(st));   // the host tables of the caller and tab / idf may go
  cudaFree(tab);
  cudaFree(idf);
  return WM_OK;
}

}  // namespace

extern "C" {

int wm_shock_inject(wm_ctx* c, const wm_shock_params* sp, int nxe, const int* nlinj, const long long* id_first, long long epoch) {
  if (!c || !sp || !nlinj || !id_first) return WM_ERR_ARG;
  if (c->swap_yz) {
    wm_set_error("the device-side shock source is not available with 3-D y-slabs (relabelled device system): use z-slabs, or the "
                 "driver's own inject() / relocate() between wm_download and wm_upload");
    return WM_ERR_STATE;
  }
  return shock_source(c, sp, nxe, nlinj, id_first, epoch, false);
}

int wm_shock_relocate(wm_ctx* c, const wm_shock_params* sp, int nxe_new, const long long* id_first, long long epoch) {
  if (!c || !sp || !id_first) return WM_ERR_ARG;
  if (c->swap_yz) {
    wm_set_error("the device-side shock source is not available with 3-D y-slabs (relabelled device system): use z-slabs, or the "
                 "driver's own inject() / relocate() between wm_download and wm_upload");
    return WM_ERR_STATE;
  }
  return shock_source(c, sp, nxe_new, nullptr, id_first, epoch, true);
}

}  // extern "C"
