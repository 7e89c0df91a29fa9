// wm_diag.cu -- synthetic Weibel load and the two diagnostics used by the parity/property tests:
//   energy_history       3d/proj/weibel/app.f90:509-577
//   Gauss-law residual   max|div E - 4 pi rho| with rho from the same quadratic spline at cell centres
//                        (forward differences implied by 3d/common/field.f90:171-187)
//   Weibel load          3d/proj/weibel/app.f90:311-338, 391-504  [2d/proj/weibel/app.f90:404-432]
#include "wm_internal.cuh"

#include <vector>

int wm_k_scalar_fold(wm_ctx* ctx, double* arr, int nxs, int nxe);  // wm_fields.cu

namespace {

constexpr int TPB = 256;
constexpr double kPi = 3.14159265358979323846264338327950288;

// Philox-4x32-10 (Salmon et al., SC'11), the stream the synthetic loaders are defined on
__host__ __device__ inline void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                       uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__host__ __device__ inline void uniform2(unsigned long long seed, uint32_t stream, uint32_t idx, uint32_t purpose,
                                         double& u0, double& u1) {
  uint32_t o[4];
  philox(idx, purpose, stream, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), o);
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

// swap: 3-D y-slabs -- the device works in (x, y' = z, z' = y) (wm_internal.cuh, wm_ctx::swap_yz): the CALLER's pencil (j, k) = (k', j')
// keys the random stream, its y goes to our z column and its anisotropic uz to our uy column, so that the state is the caller's load
__global__ void k_load_weibel(Geo g, Ptcl A, double* __restrict__ id, int n0, double v_thi, double v_the, double t_ani,
                              unsigned long long seed, int swap) {
  const int npp = n0 * g.nx;  // particles per pencil and species
  const long long n = (long long)npp * g.nyl * g.nzl;
  const int U = g.dim;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int ii = (int)(e % npp) + 1;
    const int jk = (int)(e / npp);
    const int j = g.nys + jk % g.nyl, k = g.dim == 3 ? g.nzs + jk / g.nyl : 0;
    const uint32_t pencil = swap ? (uint32_t)((k - g.nzgs) + (size_t)g.nz * (j - g.nygs))     // caller's (j_h, k_h) = (k, j), ny_h = nz
                                 : (uint32_t)((j - g.nygs) + (g.dim == 3 ? (size_t)g.ny * (k - g.nzgs) : 0));
    double u0, u1;
    uniform2(seed, pencil, (uint32_t)ii, 0u, u0, u1);
    const double x = (g.nxgs + (g.nxge - g.nxgs + 1) * (ii - 5e-1) / npp) * g.delx;
    const double y = (j + (swap ? u1 : u0)) * g.delx;
    const double z = (k + (swap ? u0 : u1)) * g.delx;
    for (int isp = 0; isp < 2; ++isp) {
      const size_t d = (size_t)g.pen(j, k, isp) * npp + (ii - 1);
      const double sd = isp == 0 ? v_thi : v_the;
      double a0, a1, b0, b1, ns, nc, ms, mc;
      uniform2(seed, pencil, (uint32_t)ii, (uint32_t)(2 * isp + 1), a0, a1);
      uniform2(seed, pencil, (uint32_t)ii, (uint32_t)(2 * isp + 2), b0, b1);
      box_muller(a0, a1, ns, nc);
      box_muller(b0, b1, ms, mc);
      A.c[0][d] = x;
      A.c[1][d] = y;
      if (g.dim == 3) A.c[2][d] = z;
      A.c[U][d] = sd * ns;
      A.c[U + (swap ? 2 : 1)][d] = sd * nc;
      A.c[U + (swap ? 1 : 2)][d] = t_ani * sd * ms;
      long long pid = (long long)isp * ((long long)npp * g.ny * g.nz) + (long long)pencil * npp + ii;
      id[d] = __longlong_as_double(-pid);
    }
  }
}

__global__ void k_load_index(Geo g, int* __restrict__ cs, int* __restrict__ np2, int* __restrict__ poff, int n0) {
  const int npp = n0 * g.nx;
  const long long n = (long long)g.npen * (g.nx + 1);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int pen = (int)(e / (g.nx + 1)), i = (int)(e % (g.nx + 1));
    cs[e] = pen * npp + i * n0;
    if (i == 0) {
      np2[pen] = npp;
      poff[pen] = pen * npp;
      if (pen == g.npen - 1) poff[g.npen] = g.npen * npp;
    }
  }
}

__global__ void k_load_uf(Geo g, double* __restrict__ uf, double b0, int swap) {
  const long long n = (long long)g.nbox();
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    double* f = uf + e * 6;
    f[0] = 0; f[1] = swap ? -b0 : 0; f[2] = swap ? 0 : b0; f[3] = 0; f[4] = 0; f[5] = 0;     // Bz = b0; relabelled: B'y' = -Bz
  }
}

__global__ void k_energy_ptcl(Geo g, Ptcl A, long long n, long long n_sp0, double* __restrict__ acc) {
  const int U = g.dim;
  double s[2] = {0, 0};
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
    double ux = A.c[U][p], uy = A.c[U + 1][p], uz = A.c[U + 2][p];
    double gam = sqrt(1.0 + (ux * ux + uy * uy + uz * uz) / (g.c * g.c));
    int isp = p < n_sp0 ? 0 : 1;
    s[isp] += g.r[isp] * (gam - 1.0);
  }
  for (int t = 0; t < 2; ++t) {
    double v = s[t];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(acc + t, v);
  }
}

__global__ void k_energy_field(Geo g, const double* __restrict__ uf, double* __restrict__ acc) {
  const long long n = (long long)g.nx * g.nyl * g.nzl;
  double eb = 0, ee = 0;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int i = g.nxgs + (int)(e % g.nx);
    long long r = e / g.nx;
    int j = g.nys + (int)(r % g.nyl), k = g.dim == 3 ? g.nzs + (int)(r / g.nyl) : 0;
    const double* f = uf + g.box(i, j, k) * 6;
    eb += f[0] * f[0] + f[1] * f[1] + f[2] * f[2];
    ee += f[3] * f[3] + f[4] * f[4] + f[5] * f[5];
  }
  for (int o = 16; o > 0; o >>= 1) {
    eb += __shfl_down_sync(0xffffffffu, eb, o);
    ee += __shfl_down_sync(0xffffffffu, ee, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(acc + 2, ee);
    atomicAdd(acc + 3, eb);
  }
}

// rho(i,j,k) += q S(i)S(j)S(k) about the particle's own cell int(x), into a scalar box array
__global__ void k_rho(Geo g, Ptcl A, long long n, long long n_sp0, double* __restrict__ rho) {
  const long long sY = g.bx, sZ = (long long)g.bx * g.by;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
    const double q = g.q[p < n_sp0 ? 0 : 1];
    double s[3][3];
    int c[3] = {0, 0, 0};
    for (int a = 0; a < g.dim; ++a) {
      double v = A.c[a][p] * g.d_delx;
      c[a] = (int)floor(v);
      double dh = v - 0.5 - c[a];
      s[a][0] = 0.5 * (0.5 - dh) * (0.5 - dh);
      s[a][1] = 0.75 - dh * dh;
      s[a][2] = 0.5 * (0.5 + dh) * (0.5 + dh);
    }
    double* base = rho + g.box(c[0], c[1], c[2]);
    if (g.dim == 3) {
      for (int kk = -1; kk <= 1; ++kk)
        for (int jj = -1; jj <= 1; ++jj)
          for (int ii = -1; ii <= 1; ++ii)
            atomicAdd(base + kk * sZ + jj * sY + ii, q * s[0][ii + 1] * s[1][jj + 1] * s[2][kk + 1]);
    } else {
      for (int jj = -1; jj <= 1; ++jj)
        for (int ii = -1; ii <= 1; ++ii) atomicAdd(base + jj * sY + ii, q * s[0][ii + 1] * s[1][jj + 1]);
    }
  }
}

// walls (bc != periodic): only the cells nxs+1 .. nxe-2 that no wall rule touches are checked
__global__ void k_gauss(Geo g, const double* __restrict__ uf, const double* __restrict__ rho,
                        unsigned long long* __restrict__ out, int i_lo, int i_hi) {
  const long long n = (long long)g.nx * g.nyl * g.nzl;
  const long long sY = (long long)g.bx * 6, sZ = (long long)g.bx * g.by * 6;
  double res = 0, mx = 0;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int i = g.nxgs + (int)(e % g.nx);
    long long r = e / g.nx;
    int j = g.nys + (int)(r % g.nyl), k = g.dim == 3 ? g.nzs + (int)(r / g.nyl) : 0;
    if (i < i_lo || i > i_hi) continue;
    const size_t o = g.box(i, j, k);
    const double* f = uf + o * 6;
    double div = (f[3 + 6] - f[3]) + (f[4 + sY] - f[4]);
    if (g.dim == 3) div += f[5 + sZ] - f[5];
    double rr = 4.0 * kPi * g.delx * rho[o];
    res = fmax(res, fabs(div - rr));
    mx = fmax(mx, fabs(rr));
  }
  atomicMax(out, (unsigned long long)__double_as_longlong(res));
  atomicMax(out + 1, (unsigned long long)__double_as_longlong(mx));
}

int grid_for(long long n) {
  long long b = (n + TPB - 1) / TPB;
  const long long cap = 148LL * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

int wm_k_load_weibel(wm_ctx* ctx, int n0, double v_thi, double v_the, double t_ani, double b0, unsigned long long seed) {
  const Geo& g = ctx->g;
  const long long n = (long long)n0 * g.nx * g.nyl * g.nzl;
  const int swap = ctx->swap_yz ? 1 : 0;
  k_load_uf<<<grid_for((long long)g.nbox()), TPB, 0, ctx->stream>>>(g, ctx->uf, b0, swap);
  WM_LAUNCH_CHECK(ctx);
  k_load_index<<<grid_for((long long)g.npen * (g.nx + 1)), TPB, 0, ctx->stream>>>(g, ctx->cs, ctx->np2, ctx->poff, n0);
  WM_LAUNCH_CHECK(ctx);
  k_load_weibel<<<grid_for(n), TPB, 0, ctx->stream>>>(g, ctx->A, ctx->id[ctx->cid], n0, v_thi, v_the, t_ani, seed, swap);
  WM_LAUNCH_CHECK(ctx);
  return WM_OK;
}

int wm_k_energy(wm_ctx* ctx, double* out_host) {
  const Geo& g = ctx->g;
  double* acc = ctx->red + 4096 + 16;
  WM_CUDA(cudaMemsetAsync(acc, 0, 4 * sizeof(double), ctx->stream));
  if (ctx->ntot > 0) {
    k_energy_ptcl<<<grid_for(ctx->ntot), TPB, 0, ctx->stream>>>(g, ctx->A, ctx->ntot, ctx->n_sp0, acc);
    WM_LAUNCH_CHECK(ctx);
  }
  k_energy_field<<<grid_for((long long)g.nx * g.nyl * g.nzl), TPB, 0, ctx->stream>>>(g, ctx->uf, acc);
  WM_LAUNCH_CHECK(ctx);
  double h[4];
  WM_CUDA(cudaMemcpyAsync(h, acc, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  WM_CUDA(cudaStreamSynchronize(ctx->stream));
  out_host[0] = h[0];
  out_host[1] = h[1];
  out_host[2] = h[2] / (8.0 * kPi);
  out_host[3] = h[3] / (8.0 * kPi);
  return WM_OK;
}

int wm_k_gauss(wm_ctx* ctx, double* out_host) {
  const Geo& g = ctx->g;
  double* rho = ctx->apcg;  // CG scratch is free outside cgm
  WM_CUDA(cudaMemsetAsync(rho, 0, g.nbox() * sizeof(double), ctx->stream));
  if (ctx->ntot > 0) {
    k_rho<<<grid_for(ctx->ntot), TPB, 0, ctx->stream>>>(g, ctx->A, ctx->ntot, ctx->n_sp0, rho);
    WM_LAUNCH_CHECK(ctx);
  }
  const bool per = g.bc == WM_BC_PERIODIC;
  const int gx0 = per ? g.nxgs : ctx->last_nxs, gx1 = per ? g.nxge : ctx->last_nxe;
  WM_TRY(wm_k_scalar_fold(ctx, rho, gx0, gx1));
  unsigned long long* acc = (unsigned long long*)(ctx->red + 4096 + 24);
  WM_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(unsigned long long), ctx->stream));
  k_gauss<<<grid_for((long long)g.nx * g.nyl * g.nzl), TPB, 0, ctx->stream>>>(g, ctx->uf, rho, acc, per ? g.nxgs : gx0 + 1,
                                                                             per ? g.nxge : gx1 - 2);
  WM_LAUNCH_CHECK(ctx);
  WM_CUDA(cudaMemcpyAsync(out_host, acc, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  WM_CUDA(cudaStreamSynchronize(ctx->stream));
  return WM_OK;
}
