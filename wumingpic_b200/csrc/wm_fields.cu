// wm_fields.cu -- implicit FDTD Maxwell update on the device.
//
//   field__fdtd_i           3d/common/field.f90:70-208   [2d/common/field.f90:66-186]
//   cgm                     3d/common/field.f90:409-560  [2d :319-461]
//   boundary_periodic__curre / __dfield / __phi
//                           3d/common/boundary_periodic.f90:676-978, 458-673, 981-1099
//                           [2d/common/boundary_periodic.f90:357-508, 251-354, 511-568]
//
// All arrays live on the "box" layout of Geo (two ghost layers, component fastest), which is the
// reference's own uf layout.  Every MPI_SENDRECV of the reference is one pack -> transport -> unpack
// triple; the transport is a pointer hand-over when the neighbour is this rank (periodic wrap on a
// single slab) and an NCCL send/recv pair otherwise (wm_comm.cu).  These kernels are HBM-bound
// stencil sweeps over (nx+4)(nyl+4)(nzl+4) cells; they are sized one thread per cell with x fastest
// so that every warp reads/writes contiguous rows.
#include "wm_internal.cuh"

#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace {

constexpr int TPB = 256;
constexpr double kPi = 3.14159265358979323846264338327950288;

// ---------------------------------------------------------------------------------------------
// plane pack / unpack: `nl` consecutive planes starting at p0 along `axis` (1 = y, 2 = z), x range
// [i0,i1], transverse range [t0,t1] (k for axis y, j for axis z), NC components.
// ---------------------------------------------------------------------------------------------
template <int NC>
__global__ void k_pack(const double* __restrict__ arr, double* __restrict__ buf, Geo g, int axis, int p0, int nl,
                       int i0, int i1, int t0, int t1) {
  const int nxr = i1 - i0 + 1, ntr = t1 - t0 + 1;
  const long long n = (long long)nl * nxr * ntr;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int i = i0 + (int)(e % nxr);
    long long rest = e / nxr;
    int t = t0 + (int)(rest % ntr);
    int l = (int)(rest / ntr);
    int j = axis == 1 ? p0 + l : t;
    int k = axis == 1 ? t : p0 + l;
    const double* s = arr + g.box(i, j, k) * NC;
    double* d = buf + e * NC;
#pragma unroll
    for (int c = 0; c < NC; ++c) d[c] = s[c];
  }
}

template <int NC, bool ADD>
__global__ void k_unpack(double* __restrict__ arr, const double* __restrict__ buf, Geo g, int axis, int p0, int nl,
                         int i0, int i1, int t0, int t1) {
  const int nxr = i1 - i0 + 1, ntr = t1 - t0 + 1;
  const long long n = (long long)nl * nxr * ntr;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int i = i0 + (int)(e % nxr);
    long long rest = e / nxr;
    int t = t0 + (int)(rest % ntr);
    int l = (int)(rest / ntr);
    int j = axis == 1 ? p0 + l : t;
    int k = axis == 1 ? t : p0 + l;
    double* d = arr + g.box(i, j, k) * NC;
    const double* s = buf + e * NC;
#pragma unroll
    for (int c = 0; c < NC; ++c) d[c] = ADD ? d[c] + s[c] : s[c];
  }
}

// One SENDRECV: planes [src_p0, src_p0+nl) are sent towards `down` (dir_down=1) or `up` and what
// arrives is unpacked (copy or add) into planes [dst_p0, dst_p0+nl).
template <int NC, bool ADD>
int exchange(wm_ctx* ctx, double* arr, int axis, bool dir_down, int src_p0, int dst_p0, int nl, int i0, int i1,
             int t0, int t1) {
  const Geo& g = ctx->g;
  const size_t n = (size_t)nl * (i1 - i0 + 1) * (t1 - t0 + 1) * NC;
  if (n == 0) return WM_OK;
  if (n > ctx->hbuf_elems) {
    wm_set_error("halo buffer too small");
    return WM_ERR_ARG;
  }
  double* snd = ctx->hbuf[0];
  double* rcv = ctx->hbuf[2];
  const int blocks = wm_blocks((long long)(n / NC), TPB);
  k_pack<NC><<<blocks, TPB, 0, ctx->stream>>>(arr, snd, g, axis, src_p0, nl, i0, i1, t0, t1);
  WM_LAUNCH_CHECK(ctx);
  WM_TRY(wm_comm_sendrecv(ctx, axis, dir_down ? 1 : 0, snd, rcv, n));
  const int peer = dir_down ? ctx->rank_down[axis - 1] : ctx->rank_up[axis - 1];
  const double* in = (peer == ctx->rank) ? snd : rcv;
  k_unpack<NC, ADD><<<blocks, TPB, 0, ctx->stream>>>(arr, in, g, axis, dst_p0, nl, i0, i1, t0, t1);
  WM_LAUNCH_CHECK(ctx);
  return WM_OK;
}

// x-direction periodic operations over all (j,k) of the given transverse ranges
template <int NC>
__global__ void k_x_fold_add(double* __restrict__ arr, Geo g, int nxs, int nxe, int j0, int j1, int k0, int k1) {
  // uj(nxe-1) += uj(nxs-2); uj(nxe) += uj(nxs-1); uj(nxs) += uj(nxe+1); uj(nxs+1) += uj(nxe+2)
  // then ghosts <- interior (boundary_periodic.f90:965-976)
  const int nj = j1 - j0 + 1, nk = k1 - k0 + 1;
  const int n = nj * nk * NC;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    int c = e % NC;
    int j = j0 + (e / NC) % nj;
    int k = k0 + (e / NC) / nj;
    double* row = arr + g.box(g.nxgs - 2, j, k) * NC + c;
    auto X = [&](int i) -> double& { return row[(size_t)(i - (g.nxgs - 2)) * NC]; };
    X(nxe - 1) = X(nxe - 1) + X(nxs - 2);
    X(nxe) = X(nxe) + X(nxs - 1);
    X(nxs) = X(nxs) + X(nxe + 1);
    X(nxs + 1) = X(nxs + 1) + X(nxe + 2);
    X(nxs - 2) = X(nxe - 1);
    X(nxs - 1) = X(nxe);
    X(nxe + 1) = X(nxs);
    X(nxe + 2) = X(nxs + 1);
  }
}

template <int NC>
__global__ void k_x_copy(double* __restrict__ arr, Geo g, int nxs, int nxe, int nl, int j0, int j1, int k0, int k1) {
  const int nj = j1 - j0 + 1, nk = k1 - k0 + 1;
  const int n = nj * nk * NC;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    int c = e % NC;
    int j = j0 + (e / NC) % nj;
    int k = k0 + (e / NC) / nj;
    double* row = arr + g.box(g.nxgs - 2, j, k) * NC + c;
    auto X = [&](int i) -> double& { return row[(size_t)(i - (g.nxgs - 2)) * NC]; };
    if (nl == 2) {
      X(nxs - 2) = X(nxe - 1);
      X(nxe + 2) = X(nxs + 1);
    }
    X(nxs - 1) = X(nxe);
    X(nxe + 1) = X(nxs);
  }
}

// conducting-wall x rule of boundary_{reconnection,shock}__dfield over all j,k incl. ghosts
// ({2d,3d}/proj/reconnection/boundary_reconnection.f90:350-359 / :672-682, {2d,3d}/proj/shock/boundary_shock.f90:396-405 / :674-686)
__global__ void k_x_wall_dfield(double* __restrict__ df, Geo g, int nxs, int nxe, int j0, int j1, int k0, int k1) {
  const int nj = j1 - j0 + 1, nk = k1 - k0 + 1;
  const int n = nj * nk;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int j = j0 + e % nj, k = k0 + e / nj;
    double* row = df + g.box(g.nxgs - 2, j, k) * 6;
    auto D = [&](int c, int i) -> double& { return row[(size_t)(i - (g.nxgs - 2)) * 6 + (c - 1)]; };
    D(1, nxs - 1) = -D(1, nxs);
    for (int c = 2; c <= 4; ++c) D(c, nxs - 1) = D(c, nxs + 1);
    for (int c = 5; c <= 6; ++c) D(c, nxs - 1) = -D(c, nxs);
    if (g.bc == WM_BC_RECONNECTION) {
      D(1, nxe) = -D(1, nxe - 1);
      for (int c = 2; c <= 4; ++c) D(c, nxe + 1) = D(c, nxe - 1);
      for (int c = 5; c <= 6; ++c) D(c, nxe) = -D(c, nxe - 1);
    } else {
      for (int c = 1; c <= 6; ++c) D(c, nxe + 1) = 0.0;
    }
  }
}

// x rule of boundary_{reconnection,shock}__phi for component l (1: odd about the wall face, 2,3: even about the wall cell)
// ({2d,3d}/proj/reconnection/boundary_reconnection.f90:557-577 / :1094-1116, {2d,3d}/proj/shock/boundary_shock.f90:603-623 / :1086-1108)
__global__ void k_x_wall_phi(double* __restrict__ a, Geo g, int nxs, int nxe, int l, int j0, int j1, int k0, int k1) {
  const int nj = j1 - j0 + 1, nk = k1 - k0 + 1;
  const int n = nj * nk;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int j = j0 + e % nj, k = k0 + e / nj;
    double* row = a + g.box(g.nxgs - 2, j, k);
    auto X = [&](int i) -> double& { return row[i - (g.nxgs - 2)]; };
    const bool rec = g.bc == WM_BC_RECONNECTION;
    if (l == 1) {
      X(nxs - 1) = -X(nxs);
      X(nxe + 1) = rec ? -X(nxe - 2) : 0.0;
    } else {
      X(nxs - 1) = X(nxs + 1);
      X(nxe + 1) = rec ? X(nxe - 1) : 0.0;
    }
  }
}

// one-layer periodic x fold of a scalar box array: a(nxe) += a(nxs-1); a(nxs) += a(nxe+1)
__global__ void k_x_fold1(double* __restrict__ arr, Geo g, int nxs, int nxe, int j0, int j1, int k0, int k1) {
  const int nj = j1 - j0 + 1, nk = k1 - k0 + 1;
  const int n = nj * nk;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    int j = j0 + e % nj, k = k0 + e / nj;
    double* row = arr + g.box(g.nxgs - 2, j, k);
    auto X = [&](int i) -> double& { return row[i - (g.nxgs - 2)]; };
    X(nxe) = X(nxe) + X(nxs - 1);
    X(nxs) = X(nxs) + X(nxe + 1);
  }
}

__global__ void k_zero_box(double* __restrict__ arr, Geo g, int nc, int i0, int i1) {
  // zero x in [i0,i1] for all j,k of the box (field.f90:227-229)
  const int nxr = i1 - i0 + 1;
  const long long n = (long long)nxr * g.by * g.bz;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int i = i0 + (int)(e % nxr);
    long long jk = e / nxr;
    double* d = arr + (jk * g.bx + (i - (g.nxgs - 2))) * nc;
    for (int c = 0; c < nc; ++c) d[c] = 0.0;
  }
}

// interior cell decode: e -> (i,j,k) with x fastest
__device__ inline void cell_of(const Geo& g, long long e, int nxs, int nxr, int& i, int& j, int& k) {
  i = nxs + (int)(e % nxr);
  long long r = e / nxr;
  j = g.nys + (int)(r % g.nyl);
  k = g.nzs + (int)(r / g.nyl);
}

// K7: RHS of the implicit equation (field.f90:128-159; 2d field.f90:125-146)
__global__ void k_gkl(const double* __restrict__ uf, const double* __restrict__ uj, double* __restrict__ gkl, Geo g,
                      int nxs, int nxe) {
  const int nxr = nxe - nxs + 1;
  const long long n = (long long)nxr * g.nyl * g.nzl;
  const double f1 = g.f1, f2 = g.f2, f3 = g.f3;
  const size_t sx = 1, sy = g.bx, sz = (size_t)g.bx * g.by;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int i, j, k;
    cell_of(g, e, nxs, nxr, i, j, k);
    const size_t o = g.box(i, j, k);
    auto F = [&](int c, long long d) { return uf[(o + d) * 6 + (c - 1)]; };
    auto J = [&](int c, long long d) { return uj[(o + d) * 3 + (c - 1)]; };
    const long long mx = -(long long)sx, px = sx, my = -(long long)sy, py = sy, mz = -(long long)sz, pz = sz;
    double g1, g2, g3;
    if (g.dim == 3) {
      g1 = +f2 * (+F(1, mz) + F(1, my) + F(1, mx) - 6.0 * F(1, 0) + F(1, px) + F(1, py) + F(1, pz)
                  + f3 * (-J(3, my) + J(3, 0) + J(2, mz) - J(2, 0)))
           - f1 * (-F(6, my) + F(6, 0) + F(5, mz) - F(5, 0));
      g2 = +f2 * (+F(2, mz) + F(2, my) + F(2, mx) - 6.0 * F(2, 0) + F(2, px) + F(2, py) + F(2, pz)
                  + f3 * (-J(1, mz) + J(1, 0) + J(3, mx) - J(3, 0)))
           - f1 * (-F(4, mz) + F(4, 0) + F(6, mx) - F(6, 0));
      g3 = +f2 * (+F(3, mz) + F(3, my) + F(3, mx) - 6.0 * F(3, 0) + F(3, px) + F(3, py) + F(3, pz)
                  + f3 * (-J(2, mx) + J(2, 0) + J(1, my) - J(1, 0)))
           - f1 * (-F(5, mx) + F(5, 0) + F(4, my) - F(4, 0));
    } else {
      g1 = +f2 * (+F(1, my) + F(1, mx) - 4.0 * F(1, 0) + F(1, px) + F(1, py) + f3 * (-J(3, my) + J(3, 0)))
           - f1 * (-F(6, my) + F(6, 0));
      g2 = +f2 * (+F(2, my) + F(2, mx) - 4.0 * F(2, 0) + F(2, px) + F(2, py) - f3 * (-J(3, mx) + J(3, 0)))
           + f1 * (-F(6, mx) + F(6, 0));
      g3 = +f2 * (+F(3, my) + F(3, mx) - 4.0 * F(3, 0) + F(3, px) + F(3, py)
                  + f3 * (-J(2, mx) + J(2, 0) + J(1, my) - J(1, 0)))
           - f1 * (-F(5, mx) + F(5, 0) + F(4, my) - F(4, 0));
    }
    double* out = gkl + o * 3;
    out[0] = g1;
    out[1] = g2;
    out[2] = g3;
  }
}

// K10: explicit dE (field.f90:167-191; 2d :154-171)
__global__ void k_de(const double* __restrict__ uf, const double* __restrict__ uj, double* __restrict__ df, Geo g,
                     int nxs, int nxe) {
  const int nxr = nxe - nxs + 1;
  const long long n = (long long)nxr * g.nyl * g.nzl;
  const double f1 = g.f1, gfac = g.gfac;
  const double fpd = 4.0 * kPi * g.delt;
  const size_t sx = 1, sy = g.bx, sz = (size_t)g.bx * g.by;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int i, j, k;
    cell_of(g, e, nxs, nxr, i, j, k);
    const size_t o = g.box(i, j, k);
    auto F = [&](int c, size_t d) { return uf[(o + d) * 6 + (c - 1)]; };
    auto D = [&](int c, size_t d) { return df[(o + d) * 6 + (c - 1)]; };
    double e4, e5, e6;
    if (g.dim == 3) {
      e4 = +f1 * (+gfac * (-D(3, 0) + D(3, sy) + D(2, 0) - D(2, sz)) + (-F(3, 0) + F(3, sy) + F(2, 0) - F(2, sz)))
           - fpd * uj[o * 3 + 0];
      e5 = +f1 * (+gfac * (-D(1, 0) + D(1, sz) + D(3, 0) - D(3, sx)) + (-F(1, 0) + F(1, sz) + F(3, 0) - F(3, sx)))
           - fpd * uj[o * 3 + 1];
      e6 = +f1 * (+gfac * (-D(2, 0) + D(2, sx) + D(1, 0) - D(1, sy)) + (-F(2, 0) + F(2, sx) + F(1, 0) - F(1, sy)))
           - fpd * uj[o * 3 + 2];
    } else {
      e4 = +f1 * (+gfac * (-D(3, 0) + D(3, sy)) + (-F(3, 0) + F(3, sy))) - fpd * uj[o * 3 + 0];
      e5 = -f1 * (+gfac * (-D(3, 0) + D(3, sx)) + (-F(3, 0) + F(3, sx))) - fpd * uj[o * 3 + 1];
      e6 = +f1 * (+gfac * (-D(2, 0) + D(2, sx) + D(1, 0) - D(1, sy)) + (-F(2, 0) + F(2, sx) + F(1, 0) - F(1, sy)))
           - fpd * uj[o * 3 + 2];
    }
    df[o * 6 + 3] = e4;
    df[o * 6 + 4] = e5;
    df[o * 6 + 5] = e6;
  }
}

// K11: uf += df over x in [nxs-2,nxe+2] and all j,k incl. ghosts (field.f90:196-206)
__global__ void k_update(double* __restrict__ uf, const double* __restrict__ df, Geo g, int nxs, int nxe) {
  const int nxr = (nxe + 2) - (nxs - 2) + 1;
  const long long n = (long long)nxr * 6 * g.by * g.bz;
  const int x0 = (nxs - 2) - (g.nxgs - 2);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    long long row = e / (nxr * 6);
    int w = (int)(e % (nxr * 6));
    size_t a = ((size_t)row * g.bx + x0) * 6 + w;
    uf[a] = uf[a] + df[a];
  }
}

// ---------------------------------------------------------------------------------------------
// cgm kernels.  phi, p, r, b, ap are scalar box arrays.  Block partial sums go to `part`
// (2 doubles per block) and are folded in a fixed order by k_fold, so results are deterministic.
// ---------------------------------------------------------------------------------------------
__device__ inline void block_sum2(double a, double b, double* part) {
  __shared__ double sa[TPB / 32], sb[TPB / 32];
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_down_sync(0xffffffffu, a, o);
    b += __shfl_down_sync(0xffffffffu, b, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sa[w] = a; sb[w] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0, tb = 0;
    for (int t = 0; t < TPB / 32; ++t) { ta += sa[t]; tb += sb[t]; }
    part[2 * blockIdx.x] = ta;
    part[2 * blockIdx.x + 1] = tb;
  }
}

// out[o0], out[o1] = sums of the per-block partials (single block, fixed order)
__global__ void k_fold(const double* __restrict__ part, int nblocks, double* __restrict__ out, int o0, int o1) {
  __shared__ double sa[TPB], sb[TPB];
  double a = 0, b = 0;
  for (int t = threadIdx.x; t < nblocks; t += TPB) { a += part[2 * t]; b += part[2 * t + 1]; }
  sa[threadIdx.x] = a; sb[threadIdx.x] = b;
  __syncthreads();
  for (int s = TPB / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) { sa[threadIdx.x] += sa[threadIdx.x + s]; sb[threadIdx.x] += sb[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[o0] = sa[0]; if (o1 >= 0) out[o1] = sb[0]; }
}

// phi = db(l); b = f5*gkl(l); sum b^2   (field.f90:442-452)
__global__ void k_cg_init(const double* __restrict__ df, const double* __restrict__ gkl, double* __restrict__ phi,
                          double* __restrict__ b, double* __restrict__ part, Geo g, int nxs, int nxe, int l) {
  const int nxr = nxe - nxs + 1;
  const long long n = (long long)nxr * g.nyl * g.nzl;
  double s = 0;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int i, j, k;
    cell_of(g, e, nxs, nxr, i, j, k);
    const size_t o = g.box(i, j, k);
    phi[o] = df[o * 6 + l];
    double bb = g.f5 * gkl[o * 3 + l];
    b[o] = bb;
    s = s + bb * bb;
  }
  block_sum2(s, 0.0, part);
}

// r = b + sum_neigh(phi) - f4*phi ; p = r ; sum r^2   (field.f90:460-473)
__global__ void k_cg_r0(const double* __restrict__ phi, const double* __restrict__ b, double* __restrict__ r,
                        double* __restrict__ p, double* __restrict__ part, Geo g, int nxs, int nxe) {
  const int nxr = nxe - nxs + 1;
  const long long n = (long long)nxr * g.nyl * g.nzl;
  const size_t sy = g.bx, sz = (size_t)g.bx * g.by;
  double s = 0;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int i, j, k;
    cell_of(g, e, nxs, nxr, i, j, k);
    const size_t o = g.box(i, j, k);
    double rr;
    if (g.dim == 3)
      rr = b[o] + phi[o - sz] + phi[o - sy] + phi[o - 1] - g.f4 * phi[o] + phi[o + 1] + phi[o + sy] + phi[o + sz];
    else
      rr = b[o] + phi[o - sy] + phi[o - 1] - g.f4 * phi[o] + phi[o + 1] + phi[o + sy];
    r[o] = rr;
    p[o] = rr;
    s = s + rr * rr;
  }
  block_sum2(s, 0.0, part);
}

// ap = f4*p - sum_neigh(p) ; sum r^2, sum p*ap   (field.f90:485-499)
__global__ void k_cg_ap(const double* __restrict__ p, const double* __restrict__ r, double* __restrict__ ap,
                        double* __restrict__ part, Geo g, int nxs, int nxe) {
  const int nxr = nxe - nxs + 1;
  const long long n = (long long)nxr * g.nyl * g.nzl;
  const size_t sy = g.bx, sz = (size_t)g.bx * g.by;
  double s1 = 0, s2 = 0;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int i, j, k;
    cell_of(g, e, nxs, nxr, i, j, k);
    const size_t o = g.box(i, j, k);
    double a;
    if (g.dim == 3)
      a = -p[o - sz] - p[o - sy] - p[o - 1] + g.f4 * p[o] - p[o + 1] - p[o + sy] - p[o + sz];
    else
      a = -p[o - sy] - p[o - 1] + g.f4 * p[o] - p[o + 1] - p[o + sy];
    ap[o] = a;
    s1 = s1 + r[o] * r[o];
    s2 = s2 + p[o] * a;
  }
  block_sum2(s1, s2, part);
}

// av = S[0]/S[1]; phi += av*p; r -= av*ap; sum r_new^2   (field.f90:507-518, 527-536)
__global__ void k_cg_update(double* __restrict__ phi, double* __restrict__ r, const double* __restrict__ p,
                            const double* __restrict__ ap, const double* __restrict__ S, double* __restrict__ part,
                            Geo g, int nxs, int nxe) {
  const int nxr = nxe - nxs + 1;
  const long long n = (long long)nxr * g.nyl * g.nzl;
  const double av = S[0] / S[1];
  double s = 0;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int i, j, k;
    cell_of(g, e, nxs, nxr, i, j, k);
    const size_t o = g.box(i, j, k);
    phi[o] = phi[o] + av * p[o];
    double rn = r[o] - av * ap[o];
    r[o] = rn;
    s = s + rn * rn;
  }
  block_sum2(s, 0.0, part);
}

// bv = S[2]/S[0]; p = r + bv*p   (field.f90:539-549)
__global__ void k_cg_p(double* __restrict__ p, const double* __restrict__ r, const double* __restrict__ S, Geo g,
                       int nxs, int nxe) {
  const int nxr = nxe - nxs + 1;
  const long long n = (long long)nxr * g.nyl * g.nzl;
  const double bv = S[2] / S[0];
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int i, j, k;
    cell_of(g, e, nxs, nxr, i, j, k);
    const size_t o = g.box(i, j, k);
    p[o] = r[o] + bv * p[o];
  }
}

// db(l) = phi   (field.f90:554-556)
__global__ void k_cg_store(double* __restrict__ df, const double* __restrict__ phi, Geo g, int nxs, int nxe, int l) {
  const int nxr = nxe - nxs + 1;
  const long long n = (long long)nxr * g.nyl * g.nzl;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int i, j, k;
    cell_of(g, e, nxs, nxr, i, j, k);
    const size_t o = g.box(i, j, k);
    df[o * 6 + l] = phi[o];
  }
}

// ---------------------------------------------------------------------------------------------
// cgm as ONE persistent cooperative kernel (single rank; periodic or wall boundaries): all three components, all
// iterations, no host round trip.  Control flow is field.f90:437-558 verbatim, including
// the reference's stop rule quirks; the periodic ghost fill of boundary_periodic__phi becomes index
// wrapping.  Dot products: per-block partials in a fixed order, then EVERY block folds all partials in the
// same fixed order after a grid sync, so all blocks take the same branch and results are deterministic.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_fold(double v, double* sh) {
  // fixed-order tree over the block; every thread returns the total
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double t = 0;
#pragma unroll
  for (int q = 0; q < TPB / 32; ++q) t += sh[q];
  return t;
}

// two-value grid reduction with ONE grid sync: the partial buffers ping-pong (a block can be at most one
// reduction ahead of the slowest block, because it cannot pass the next sync alone)
__device__ __forceinline__ void grid_sum2(cg::grid_group& grid, double a, double b, double* part, int& red_idx, double* sh,
                                          double& ta, double& tb) {
  a = block_fold(a, sh);
  b = block_fold(b, sh);
  double* buf = part + (size_t)(red_idx & 1) * 2 * gridDim.x;
  red_idx++;
  if (threadIdx.x == 0) { buf[2 * blockIdx.x] = a; buf[2 * blockIdx.x + 1] = b; }
  grid.sync();
  double x = 0, y = 0;
  for (int q = threadIdx.x; q < (int)gridDim.x; q += TPB) { x += __ldcg(buf + 2 * q); y += __ldcg(buf + 2 * q + 1); }
  ta = block_fold(x, sh);
  tb = block_fold(y, sh);
}

// ---- multi-GPU form of grid_sum2 (PeerCG, wm_internal.cuh): the all-reduce of the reference's MPI_ALLREDUCE over NVLink peer
// memory, inside the kernel.  After the local grid sync, block 0 folds the block partials and one thread per destination
// rank stores (a, b) and then -- behind a system fence -- the sequence number into that rank's mailbox slot [parity][my rank];
// thread 0 of every block then polls the LOCAL mailbox until all ranks' slots carry this sequence number and sums them in rank
// order, so every block of every rank gets bit-identical totals and takes the same branches.  The same fence + flag also
// publishes the ghost-plane stores a rank made into its neighbours' operand arrays before the reduction (thread 0 of every block
// runs __threadfence_system() behind a CTA barrier, before the grid sync), and the system fence after the poll orders the ghost reads behind it.
// Two parities suffice: a rank cannot start reduction n+2 before every rank has finished reading reduction n.
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void peer_sum2(cg::grid_group& grid, const PeerCG& pc, unsigned long long& seq, double a, double b,
                                          double* part, int& red_idx, double* sh, double* sh2, int* flags, double& ta, double& tb) {
  a = block_fold(a, sh);   // contains CTA barriers: every ghost-plane store of this block precedes thread 0's fence below
  b = block_fold(b, sh);
  double* buf = part + (size_t)(red_idx & 1) * 2 * gridDim.x;
  red_idx++;
  if (threadIdx.x == 0) {
    buf[2 * blockIdx.x] = a; buf[2 * blockIdx.x + 1] = b;
    __threadfence_system();   // one system fence per block (cumulative over the block's peer stores), not one per storing thread
  }
  grid.sync();
  seq++;
  const int par = (int)(seq & 1ull);
  if (blockIdx.x == 0) {
    double x = 0, y = 0;
    for (int q = threadIdx.x; q < (int)gridDim.x; q += TPB) { x += __ldcg(buf + 2 * q); y += __ldcg(buf + 2 * q + 1); }
    x = block_fold(x, sh);
    y = block_fold(y, sh);
    if ((int)threadIdx.x < pc.nranks) {
      WmMail* m = pc.mail[threadIdx.x] + (par * pc.nranks + pc.rank);
      st_volatile_u64(reinterpret_cast<unsigned long long*>(&m->a), (unsigned long long)__double_as_longlong(x));
      st_volatile_u64(reinterpret_cast<unsigned long long*>(&m->b), (unsigned long long)__double_as_longlong(y));
      __threadfence_system();
      st_volatile_u64(&m->seq, seq);
    }
  }
  if (threadIdx.x == 0) {
    const WmMail* mine = pc.mail[pc.rank] + par * pc.nranks;
    double x = 0, y = 0;
    bool dead = (*(volatile int*)flags & 8) != 0;
    for (int s = 0; s < pc.nranks; ++s) {
      long long spins = 0;
      while (!dead && ld_volatile_u64(&mine[s].seq) != seq) {
        if (++spins > (1ll << 27)) { dead = true; atomicOr(flags, 8); }   // a peer never arrived: sticky error, no hang
      }
    }
    __threadfence_system();
    for (int s = 0; s < pc.nranks; ++s) {
      x += __longlong_as_double((long long)ld_volatile_u64(reinterpret_cast<const unsigned long long*>(&mine[s].a)));
      y += __longlong_as_double((long long)ld_volatile_u64(reinterpret_cast<const unsigned long long*>(&mine[s].b)));
    }
    sh2[0] = x; sh2[1] = y;
  }
  __syncthreads();
  ta = sh2[0];
  tb = sh2[1];
  __syncthreads();
}

__device__ __forceinline__ void cell_of32(const Geo& g, int e, int nxs, int nxr, int& i, int& j, int& k) {
  i = nxs + e % nxr;
  const int r = e / nxr;
  j = g.nys + r % g.nyl;
  k = g.nzs + r / g.nyl;
}

// PEER = true: the slab-decomposed solve of a multi-GPU run as the same single launch per GPU.  The slab-axis neighbours
// (z in 3-D, y in 2-D) are the ghost planes of the local arrays, which the NEIGHBOUR RANKS fill: every sweep that writes phi, p
// or r also stores its two edge planes into the neighbours' ghost planes over NVLink, and the dot products are peer_sum2.
template <bool PEER>
__global__ void __launch_bounds__(TPB) k_cgm_coop(double* __restrict__ df, const double* __restrict__ gkl,
                                                  double* phi, double* p0, double* p1, double* r,
                                                  double* __restrict__ b, double* __restrict__ ap, double* __restrict__ part,
                                                  int* __restrict__ ite_out, int* flags, Geo g, int nxs, int nxe, PeerCG pc) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sh[TPB / 32];
  __shared__ double sh2[2];
  unsigned long long seq = 0;
  if (PEER) seq = *pc.seq;    // reductions done by earlier solves (identical on every rank)
  // edge-plane pushes: my first plane along the slab axis goes to the lower neighbour's upper ghost plane, my last one to the
  // upper neighbour's lower ghost plane (index shifts pc.shift_lo / pc.shift_hi inside the same box layout)
#define WM_PUSH(which, o, jj, kk, val)                                                          \
  if (PEER) {                                                                                   \
    const int sl = d3 ? (kk) : (jj);                                                            \
    if (sl == (d3 ? g.nzs : g.nys)) pc.lo[which][(long long)(o) + pc.shift_lo] = (val);                   \
    if (sl == (d3 ? g.nze : g.nye)) pc.hi[which][(long long)(o) + pc.shift_hi] = (val);                   \
  }
#define WM_SUM2(a_, b_, ta_, tb_)                                                               \
  do {                                                                                          \
    if (PEER) {                                                                                 \
      peer_sum2(grid, pc, seq, a_, b_, part, red_idx, sh, sh2, flags, ta_, tb_);                \
    } else {                                                                                    \
      grid_sum2(grid, a_, b_, part, red_idx, sh, ta_, tb_);                                     \
    }                                                                                           \
  } while (0)
  const int nxr = nxe - nxs + 1;
  const int n = nxr * g.nyl * g.nzl;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const long long sy = g.bx, sz = (long long)g.bx * g.by;
  const double err = 1e-6;
  const int ite_max = 100;
  const bool d3 = g.dim == 3;
  int red_idx = 0;
  // neighbours: periodic wrap inside [nys,nye] x [nzs,nze]; in x periodic wrap, or the wall rule of boundary_*__phi as an
  // (index, coefficient) pair: l = 1 (odd about the wall face): phi(nxs-1) = -phi(nxs), phi(nxe+1) = -phi(nxe-2) [reconnection]
  // or 0 [shock]; l = 2,3 (even about the wall cell): phi(nxs-1) = phi(nxs+1), phi(nxe+1) = phi(nxe-1) or 0
  const bool per = g.bc == WM_BC_PERIODIC, rec = g.bc == WM_BC_RECONNECTION;
#define WM_NBR(e)                                                                     \
  int i, j, k;                                                                        \
  cell_of32(g, e, nxs, nxr, i, j, k);                                                 \
  const long long o = (long long)g.box(i, j, k);                                      \
  long long xm = o - 1, xp = o + 1;                                                   \
  double cxm = 1.0, cxp = 1.0;                                                        \
  if (i == nxs) {                                                                     \
    if (per) xm = o + (nxr - 1);                                                      \
    else if (l == 0) { xm = o; cxm = -1.0; }                                          \
    else xm = o + 1;                                                                  \
  }                                                                                   \
  if (i == nxe) {                                                                     \
    if (per) xp = o - (nxr - 1);                                                      \
    else if (!rec) { xp = o; cxp = 0.0; }                                             \
    else if (l == 0) { xp = o - 2; cxp = -1.0; }                                      \
    else xp = o - 1;                                                                  \
  }                                                                                   \
  const bool ywrap = !(PEER && !d3), zwrap = !PEER;   /* the slab axis reads the ghost planes the neighbours fill */ \
  const long long ym = (ywrap && j == g.nys) ? o + (g.nyl - 1) * sy : o - sy;                                  \
  const long long yp = (ywrap && j == g.nye) ? o - (g.nyl - 1) * sy : o + sy;                                  \
  const long long zm = !d3 ? o : ((zwrap && k == g.nzs) ? o + (g.nzl - 1) * sz : o - sz);                      \
  const long long zp = !d3 ? o : ((zwrap && k == g.nze) ? o - (g.nzl - 1) * sz : o + sz);
  // slab-axis neighbours (ghost planes written by other GPUs) are read through L2
#define WM_LDS(arr, x) (PEER ? __ldcg((arr) + (x)) : (arr)[x])
  for (int l = 0; l < 3; ++l) {
    int ite = 0;
    double s = 0, dummy;
    double* pold = p0;   // the search direction of the previous iteration (complete)
    double* pnew = p1;
    int wold = 1, wnew = 2;   // the same two buffers in the neighbours' tables (pc.lo / pc.hi)
    for (int e = tid; e < n; e += nth) {
      int i, j, k;
      cell_of32(g, e, nxs, nxr, i, j, k);
      const size_t o = g.box(i, j, k);
      const double ph = df[o * 6 + l];
      phi[o] = ph;
      WM_PUSH(0, o, j, k, ph)
      const double bb = g.f5 * gkl[o * 3 + l];
      b[o] = bb;
      s = s + bb * bb;
    }
    double sum_g;
    WM_SUM2(s, 0.0, sum_g, dummy);      // its sync also orders phi before the stencil below
    const double eps = sqrt(sum_g) * err;
    s = 0;
    for (int e = tid; e < n; e += nth) {
      WM_NBR(e)
      double rr;
      if (d3) rr = b[o] + WM_LDS(phi, zm) + WM_LDS(phi, ym) + cxm * phi[xm] - g.f4 * phi[o] + cxp * phi[xp] + WM_LDS(phi, yp) + WM_LDS(phi, zp);
      else rr = b[o] + WM_LDS(phi, ym) + cxm * phi[xm] - g.f4 * phi[o] + cxp * phi[xp] + WM_LDS(phi, yp);
      r[o] = rr;
      pold[o] = rr;
      WM_PUSH(3, o, j, k, rr)
      WM_PUSH(wold, o, j, k, rr)
      s = s + rr * rr;
    }
    double sumr_g;
    WM_SUM2(s, 0.0, sumr_g, dummy);
    if (sqrt(sumr_g) > eps) {
      double bv = 0.0;   // first iteration: p = r
      while (sum_g > eps) {
        ite = ite + 1;
        // p = r + bv*p (field.f90:539-549 of the previous iteration) is evaluated on the fly at the seven stencil
        // points with the same fma everywhere, so no separate sweep / grid sync is needed for it
        double s1 = 0, s2 = 0;
        for (int e = tid; e < n; e += nth) {
          WM_NBR(e)
#define WM_P(x) fma(bv, pold[x], r[x])
#define WM_PS(x) fma(bv, WM_LDS(pold, x), WM_LDS(r, x))
          const double pc_ = WM_P(o);
          double a;
          if (d3) a = -WM_PS(zm) - WM_PS(ym) - cxm * WM_P(xm) + g.f4 * pc_ - cxp * WM_P(xp) - WM_PS(yp) - WM_PS(zp);
          else a = -WM_PS(ym) - cxm * WM_P(xm) + g.f4 * pc_ - cxp * WM_P(xp) - WM_PS(yp);
#undef WM_P
#undef WM_PS
          pnew[o] = pc_;
          WM_PUSH(wnew, o, j, k, pc_)
          ap[o] = a;
          s1 = s1 + r[o] * r[o];
          s2 = s2 + pc_ * a;
        }
        double sum2_g;
        WM_SUM2(s1, s2, sumr_g, sum2_g);
        const double av = sumr_g / sum2_g;
        s = 0;
        for (int e = tid; e < n; e += nth) {
          int i, j, k;
          cell_of32(g, e, nxs, nxr, i, j, k);
          const size_t o = g.box(i, j, k);
          phi[o] = phi[o] + av * pnew[o];
          const double rn = r[o] - av * ap[o];
          r[o] = rn;
          WM_PUSH(3, o, j, k, rn)
          s = s + rn * rn;
        }
        sum_g = sqrt(sumr_g);
        if (ite >= ite_max) {
          if (tid == 0) atomicOr(flags, 4);   // "stop at cgm after ite_max"  field.f90:522-525
          break;
        }
        double sum1_g;
        WM_SUM2(s, 0.0, sum1_g, dummy);   // its sync completes r and pnew for the next stencil
        bv = sum1_g / sumr_g;
        double* tsw = pold; pold = pnew; pnew = tsw;
        const int wsw = wold; wold = wnew; wnew = wsw;
      }
    }
    if (PEER) { __syncthreads(); if (threadIdx.x == 0) __threadfence_system(); }
    grid.sync();   // every block is past the stencil reads of phi / r before df and the next component's phi are written
    for (int e = tid; e < n; e += nth) {
      int i, j, k;
      cell_of32(g, e, nxs, nxr, i, j, k);
      const size_t o = g.box(i, j, k);
      df[o * 6 + l] = phi[o];
    }
    if (tid == 0) ite_out[l] = ite;
    grid.sync();
  }
  if (PEER && tid == 0) *pc.seq = seq;
#undef WM_LDS
#undef WM_PUSH
#undef WM_SUM2
#undef WM_NBR
}

// ---------------------------------------------------------------------------------------------
// Multi-rank cgm (slabs along the last axis): the host sequences the iterations (NCCL cannot be called from a kernel),
// but every iteration is 5 kernels + ONE send/recv group + 2 all-reduces instead of 14 kernels + 2 groups + 2 all-reduces:
//   * the slab-axis halo of p travels as one pack kernel, one NCCL group (both directions), one unpack kernel;
//   * the x rule (periodic wrap or wall) and the wrap of the non-decomposed transverse axis are index arithmetic inside
//     the stencil kernels, exactly as in k_cgm_coop;
//   * the block partials are folded by the last block to finish (fixed order -> deterministic), no separate fold launch.
// ---------------------------------------------------------------------------------------------
__device__ inline void block_sum2_fold(double a, double b, double* part, double* S, int o0, int o1, unsigned* counter) {
  __shared__ double sa[TPB], sb[TPB];
  __shared__ bool last;
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_down_sync(0xffffffffu, a, o);
    b += __shfl_down_sync(0xffffffffu, b, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sa[w] = a; sb[w] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0, tb = 0;
    for (int t = 0; t < TPB / 32; ++t) { ta += sa[t]; tb += sb[t]; }
    part[2 * blockIdx.x] = ta;
    part[2 * blockIdx.x + 1] = tb;
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  double x = 0, y = 0;
  for (int t = threadIdx.x; t < (int)gridDim.x; t += TPB) { x += __ldcg(part + 2 * t); y += __ldcg(part + 2 * t + 1); }
  sa[threadIdx.x] = x; sb[threadIdx.x] = y;
  __syncthreads();
  for (int s = TPB / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) { sa[threadIdx.x] += sa[threadIdx.x + s]; sb[threadIdx.x] += sb[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { S[o0] = sa[0]; if (o1 >= 0) S[o1] = sb[0]; *counter = 0; }
}

// neighbour indices of interior cell (i,j,k) for the slab-decomposed solve: slab axis (z in 3-D, y in 2-D) reads the ghost
// planes, the other transverse axis wraps periodically, x wraps or follows the wall rule of component l (0-based)
struct Nbr { long long xm, xp, ym, yp, zm, zp; double cxm, cxp; };
__device__ __forceinline__ Nbr slab_nbr(const Geo& g, long long o, int i, int j, int nxs, int nxe, int l) {
  const long long sy = g.bx, sz = (long long)g.bx * g.by;
  const int nxr = nxe - nxs + 1;
  Nbr n;
  n.xm = o - 1; n.xp = o + 1; n.cxm = 1.0; n.cxp = 1.0;
  const bool per = g.bc == WM_BC_PERIODIC, rec = g.bc == WM_BC_RECONNECTION;
  if (i == nxs) {
    if (per) n.xm = o + (nxr - 1);
    else if (l == 0) { n.xm = o; n.cxm = -1.0; }
    else n.xm = o + 1;
  }
  if (i == nxe) {
    if (per) n.xp = o - (nxr - 1);
    else if (!rec) { n.xp = o; n.cxp = 0.0; }
    else if (l == 0) { n.xp = o - 2; n.cxp = -1.0; }
    else n.xp = o - 1;
  }
  if (g.dim == 3) {
    n.ym = j == g.nys ? o + (g.nyl - 1) * sy : o - sy;
    n.yp = j == g.nye ? o - (g.nyl - 1) * sy : o + sy;
    n.zm = o - sz; n.zp = o + sz;
  } else {
    n.ym = o - sy; n.yp = o + sy; n.zm = o; n.zp = o;
  }
  return n;
}

__global__ void k_cgw_init(const double* __restrict__ df, const double* __restrict__ gkl, double* __restrict__ phi,
                           double* __restrict__ b, double* __restrict__ part, double* S, unsigned* counter, Geo g, int nxs,
                           int nxe, int l) {
  const int nxr = nxe - nxs + 1;
  const long long n = (long long)nxr * g.nyl * g.nzl;
  double s = 0;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int i, j, k;
    cell_of(g, e, nxs, nxr, i, j, k);
    const size_t o = g.box(i, j, k);
    phi[o] = df[o * 6 + l];
    const double bb = g.f5 * gkl[o * 3 + l];
    b[o] = bb;
    s = s + bb * bb;
  }
  block_sum2_fold(s, 0.0, part, S, 3, -1, counter);
}

__global__ void k_cgw_r0(const double* __restrict__ phi, const double* __restrict__ b, double* __restrict__ r,
                         double* __restrict__ p, double* __restrict__ part, double* S, unsigned* counter, Geo g, int nxs,
                         int nxe, int l) {
  const int nxr = nxe - nxs + 1;
  const long long n = (long long)nxr * g.nyl * g.nzl;
  double s = 0;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int i, j, k;
    cell_of(g, e, nxs, nxr, i, j, k);
    const long long o = (long long)g.box(i, j, k);
    const Nbr nb = slab_nbr(g, o, i, j, nxs, nxe, l);
    double rr;
    if (g.dim == 3)
      rr = b[o] + phi[nb.zm] + phi[nb.ym] + nb.cxm * phi[nb.xm] - g.f4 * phi[o] + nb.cxp * phi[nb.xp] + phi[nb.yp] + phi[nb.zp];
    else
      rr = b[o] + phi[nb.ym] + nb.cxm * phi[nb.xm] - g.f4 * phi[o] + nb.cxp * phi[nb.xp] + phi[nb.yp];
    r[o] = rr;
    p[o] = rr;
    s = s + rr * rr;
  }
  block_sum2_fold(s, 0.0, part, S, 0, -1, counter);
}

__global__ void k_cgw_ap(const double* __restrict__ p, const double* __restrict__ r, double* __restrict__ ap,
                         double* __restrict__ part, double* S, unsigned* counter, Geo g, int nxs, int nxe, int l) {
  const int nxr = nxe - nxs + 1;
  const long long n = (long long)nxr * g.nyl * g.nzl;
  double s1 = 0, s2 = 0;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int i, j, k;
    cell_of(g, e, nxs, nxr, i, j, k);
    const long long o = (long long)g.box(i, j, k);
    const Nbr nb = slab_nbr(g, o, i, j, nxs, nxe, l);
    double a;
    if (g.dim == 3)
      a = -p[nb.zm] - p[nb.ym] - nb.cxm * p[nb.xm] + g.f4 * p[o] - nb.cxp * p[nb.xp] - p[nb.yp] - p[nb.zp];
    else
      a = -p[nb.ym] - nb.cxm * p[nb.xm] + g.f4 * p[o] - nb.cxp * p[nb.xp] - p[nb.yp];
    ap[o] = a;
    s1 = s1 + r[o] * r[o];
    s2 = s2 + p[o] * a;
  }
  block_sum2_fold(s1, s2, part, S, 0, 1, counter);
}

__global__ void k_cgw_update(double* __restrict__ phi, double* __restrict__ r, const double* __restrict__ p,
                             const double* __restrict__ ap, double* S, double* __restrict__ part, unsigned* counter, Geo g,
                             int nxs, int nxe) {
  const int nxr = nxe - nxs + 1;
  const long long n = (long long)nxr * g.nyl * g.nzl;
  const double av = S[0] / S[1];
  double s = 0;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int i, j, k;
    cell_of(g, e, nxs, nxr, i, j, k);
    const size_t o = g.box(i, j, k);
    phi[o] = phi[o] + av * p[o];
    const double rn = r[o] - av * ap[o];
    r[o] = rn;
    s = s + rn * rn;
  }
  block_sum2_fold(s, 0.0, part, S, 2, -1, counter);
}

// both faces of the slab axis of a scalar box array: buf = [lo face | hi face], each (transverse interior) x (nxs..nxe)
__global__ void k_halo_pack2(const double* __restrict__ a, double* __restrict__ buf, Geo g, int nxs, int nxe) {
  const int nxr = nxe - nxs + 1;
  const int ntr = g.dim == 3 ? g.nyl : 1;
  const int n1 = nxr * ntr;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 2 * n1; e += gridDim.x * blockDim.x) {
    const int face = e / n1, r = e % n1;
    const int i = nxs + r % nxr, t = r / nxr;
    const size_t o = g.dim == 3 ? g.box(i, g.nys + t, face == 0 ? g.nzs : g.nze) : g.box(i, face == 0 ? g.nys : g.nye, 0);
    buf[e] = a[o];
  }
}
// buf = [from the up neighbour (its lo face) -> my hi ghost | from the down neighbour (its hi face) -> my lo ghost]
__global__ void k_halo_unpack2(double* __restrict__ a, const double* __restrict__ buf, Geo g, int nxs, int nxe) {
  const int nxr = nxe - nxs + 1;
  const int ntr = g.dim == 3 ? g.nyl : 1;
  const int n1 = nxr * ntr;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 2 * n1; e += gridDim.x * blockDim.x) {
    const int face = e / n1, r = e % n1;
    const int i = nxs + r % nxr, t = r / nxr;
    const size_t o = g.dim == 3 ? g.box(i, g.nys + t, face == 0 ? g.nze + 1 : g.nzs - 1) : g.box(i, face == 0 ? g.nye + 1 : g.nys - 1, 0);
    a[o] = buf[e];
  }
}

int halo_slab(wm_ctx* ctx, double* a, int nxs, int nxe) {
  const Geo& g = ctx->g;
  const int ax = g.dim == 3 ? 1 : 0;
  const size_t n1 = (size_t)(nxe - nxs + 1) * (g.dim == 3 ? g.nyl : 1);
  if (2 * n1 > ctx->hbuf_elems) { wm_set_error("halo buffer too small"); return WM_ERR_ARG; }
  double *snd = ctx->hbuf[0], *rcv = ctx->hbuf[2];
  const int blocks = wm_blocks((long long)2 * n1, TPB);
  k_halo_pack2<<<blocks, TPB, 0, ctx->stream>>>(a, snd, g, nxs, nxe);
  WM_LAUNCH_CHECK(ctx);
  WM_TRY(wm_comm_group_begin(ctx));
  WM_TRY(wm_comm_send(ctx, ctx->rank_down[ax], snd, n1 * sizeof(double)));
  WM_TRY(wm_comm_recv(ctx, ctx->rank_up[ax], rcv, n1 * sizeof(double)));
  WM_TRY(wm_comm_send(ctx, ctx->rank_up[ax], snd + n1, n1 * sizeof(double)));
  WM_TRY(wm_comm_recv(ctx, ctx->rank_down[ax], rcv + n1, n1 * sizeof(double)));
  WM_TRY(wm_comm_group_end(ctx));
  k_halo_unpack2<<<blocks, TPB, 0, ctx->stream>>>(a, rcv, g, nxs, nxe);
  WM_LAUNCH_CHECK(ctx);
  return WM_OK;
}

int grid_for(long long n) {
  long long b = (n + TPB - 1) / TPB;
  const long long cap = 148LL * 8;  // persistent-style grid: 8 CTAs of 256 threads per SM
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// boundary_*__phi: one ghost layer of a scalar box array (boundary_periodic.f90:981-1099)
int bc_phi(wm_ctx* ctx, double* a, int nxs, int nxe, int l) {
  const Geo& g = ctx->g;
  const int k0 = g.dim == 3 ? g.nzs : 0, k1 = g.dim == 3 ? g.nze : 0;
  // y: row nys -> jdown's nye+1 ; row nye -> jup's nys-1   (i interior, k interior)
  WM_TRY((exchange<1, false>(ctx, a, 1, true, g.nys, g.nye + 1, 1, nxs, nxe, k0, k1)));
  WM_TRY((exchange<1, false>(ctx, a, 1, false, g.nye, g.nys - 1, 1, nxs, nxe, k0, k1)));
  if (g.dim == 3) {
    // z: plane nzs -> kdown's nze+1 ; plane nze -> kup's nzs-1   (j in [nys-1,nye+1])
    WM_TRY((exchange<1, false>(ctx, a, 2, true, g.nzs, g.nze + 1, 1, nxs, nxe, g.nys - 1, g.nye + 1)));
    WM_TRY((exchange<1, false>(ctx, a, 2, false, g.nze, g.nzs - 1, 1, nxs, nxe, g.nys - 1, g.nye + 1)));
  }
  // x periodic, one layer, over the ghost-extended transverse ranges
  const int kk0 = g.dim == 3 ? g.nzs - 1 : 0, kk1 = g.dim == 3 ? g.nze + 1 : 0;
  const int n = (g.nye - g.nys + 3) * (kk1 - kk0 + 1);
  if (g.bc == WM_BC_PERIODIC)
    k_x_copy<1><<<wm_blocks(n, TPB), TPB, 0, ctx->stream>>>(a, g, nxs, nxe, 1, g.nys - 1, g.nye + 1, kk0, kk1);
  else
    k_x_wall_phi<<<wm_blocks(n, TPB), TPB, 0, ctx->stream>>>(a, g, nxs, nxe, l, g.nys - 1, g.nye + 1, kk0, kk1);
  WM_LAUNCH_CHECK(ctx);
  return WM_OK;
}

}  // namespace

int wm_k_zero_uj(wm_ctx* ctx, int nxs, int nxe) {
  const Geo& g = ctx->g;
  const long long n = (long long)(nxe - nxs + 5) * g.by * g.bz;
  k_zero_box<<<grid_for(n), TPB, 0, ctx->stream>>>(ctx->uj, g, 3, nxs - 2, nxe + 2);
  WM_LAUNCH_CHECK(ctx);
  return WM_OK;
}

// boundary_periodic__curre (boundary_periodic.f90:676-978; 2d :357-508)
int wm_k_curre(wm_ctx* ctx, int nxs, int nxe) {
  const Geo& g = ctx->g;
  double* uj = ctx->uj;
  if (g.dim == 3) {
    // 1) y add, 2 layers, all k incl. ghosts
    WM_TRY((exchange<3, true>(ctx, uj, 1, true, g.nys - 2, g.nye - 1, 2, nxs - 2, nxe + 2, g.nzs - 2, g.nze + 2)));
    WM_TRY((exchange<3, true>(ctx, uj, 1, false, g.nye + 1, g.nys, 2, nxs - 2, nxe + 2, g.nzs - 2, g.nze + 2)));
    // 2) z add, 2 layers, interior j only
    WM_TRY((exchange<3, true>(ctx, uj, 2, true, g.nzs - 2, g.nze - 1, 2, nxs - 2, nxe + 2, g.nys, g.nye)));
    WM_TRY((exchange<3, true>(ctx, uj, 2, false, g.nze + 1, g.nzs, 2, nxs - 2, nxe + 2, g.nys, g.nye)));
    // 3) y copy-back, 1 layer, interior k
    WM_TRY((exchange<3, false>(ctx, uj, 1, true, g.nys, g.nye + 1, 1, nxs - 2, nxe + 2, g.nzs, g.nze)));
    WM_TRY((exchange<3, false>(ctx, uj, 1, false, g.nye, g.nys - 1, 1, nxs - 2, nxe + 2, g.nzs, g.nze)));
    // 4) z copy-back, 1 layer, j in [nys-1,nye+1]
    WM_TRY((exchange<3, false>(ctx, uj, 2, true, g.nzs, g.nze + 1, 1, nxs - 2, nxe + 2, g.nys - 1, g.nye + 1)));
    WM_TRY((exchange<3, false>(ctx, uj, 2, false, g.nze, g.nzs - 1, 1, nxs - 2, nxe + 2, g.nys - 1, g.nye + 1)));
  } else {
    // 2-D: 2 layers add, then 2 layers copy-back (2d boundary_periodic.f90:357-508)
    WM_TRY((exchange<3, true>(ctx, uj, 1, true, g.nys - 2, g.nye - 1, 2, nxs - 2, nxe + 2, 0, 0)));
    WM_TRY((exchange<3, true>(ctx, uj, 1, false, g.nye + 1, g.nys, 2, nxs - 2, nxe + 2, 0, 0)));
    WM_TRY((exchange<3, false>(ctx, uj, 1, true, g.nys, g.nye + 1, 2, nxs - 2, nxe + 2, 0, 0)));
    WM_TRY((exchange<3, false>(ctx, uj, 1, false, g.nye - 1, g.nys - 2, 2, nxs - 2, nxe + 2, 0, 0)));
  }
  if (g.bc == WM_BC_PERIODIC) {
    const int k0 = g.dim == 3 ? g.nzs - 2 : 0, k1 = g.dim == 3 ? g.nze + 2 : 0;
    const int n = g.by * (k1 - k0 + 1) * 3;
    k_x_fold_add<3><<<wm_blocks(n, TPB), TPB, 0, ctx->stream>>>(uj, g, nxs, nxe, g.nys - 2, g.nye + 2, k0, k1);
    WM_LAUNCH_CHECK(ctx);
  }
  return WM_OK;
}

int wm_k_gkl(wm_ctx* ctx, int nxs, int nxe) {
  const Geo& g = ctx->g;
  const long long n = (long long)(nxe - nxs + 1) * g.nyl * g.nzl;
  k_gkl<<<grid_for(n), TPB, 0, ctx->stream>>>(ctx->uf, ctx->uj, ctx->gkl, g, nxs, nxe);
  WM_LAUNCH_CHECK(ctx);
  return WM_OK;
}

// boundary_periodic__dfield (boundary_periodic.f90:458-673; 2d :251-354)
int wm_k_dfield(wm_ctx* ctx, int nxs, int nxe) {
  const Geo& g = ctx->g;
  double* df = ctx->df;
  const int k0 = g.dim == 3 ? g.nzs : 0, k1 = g.dim == 3 ? g.nze : 0;
  WM_TRY((exchange<6, false>(ctx, df, 1, true, g.nys, g.nye + 1, 2, nxs, nxe, k0, k1)));
  WM_TRY((exchange<6, false>(ctx, df, 1, false, g.nye - 1, g.nys - 2, 2, nxs, nxe, k0, k1)));
  if (g.dim == 3) {
    WM_TRY((exchange<6, false>(ctx, df, 2, true, g.nzs, g.nze + 1, 2, nxs, nxe, g.nys - 2, g.nye + 2)));
    WM_TRY((exchange<6, false>(ctx, df, 2, false, g.nze - 1, g.nzs - 2, 2, nxs, nxe, g.nys - 2, g.nye + 2)));
  }
  if (g.bc == WM_BC_PERIODIC) {
    const int kk0 = g.dim == 3 ? g.nzs - 2 : 0, kk1 = g.dim == 3 ? g.nze + 2 : 0;
    const int n = g.by * (kk1 - kk0 + 1) * 6;
    k_x_copy<6><<<wm_blocks(n, TPB), TPB, 0, ctx->stream>>>(df, g, nxs, nxe, 2, g.nys - 2, g.nye + 2, kk0, kk1);
    WM_LAUNCH_CHECK(ctx);
  } else {
    const int kk0 = g.dim == 3 ? g.nzs - 2 : 0, kk1 = g.dim == 3 ? g.nze + 2 : 0;
    const int n = g.by * (kk1 - kk0 + 1);
    k_x_wall_dfield<<<wm_blocks(n, TPB), TPB, 0, ctx->stream>>>(df, g, nxs, nxe, g.nys - 2, g.nye + 2, kk0, kk1);
    WM_LAUNCH_CHECK(ctx);
  }
  return WM_OK;
}

int wm_k_de(wm_ctx* ctx, int nxs, int nxe) {
  const Geo& g = ctx->g;
  const long long n = (long long)(nxe - nxs + 1) * g.nyl * g.nzl;
  k_de<<<grid_for(n), TPB, 0, ctx->stream>>>(ctx->uf, ctx->uj, ctx->df, g, nxs, nxe);
  WM_LAUNCH_CHECK(ctx);
  return WM_OK;
}

int wm_k_update(wm_ctx* ctx, int nxs, int nxe) {
  const Geo& g = ctx->g;
  const long long n = (long long)(nxe - nxs + 5) * 6 * g.by * g.bz;
  k_update<<<grid_for(n), TPB, 0, ctx->stream>>>(ctx->uf, ctx->df, g, nxs, nxe);
  WM_LAUNCH_CHECK(ctx);
  return WM_OK;
}

// cgm (field.f90:409-560): host-driven loop, device-side scalars.  One 8-byte read-back per
// iteration decides the `do while(sum_g > eps)` test; av, bv never leave the device.
int wm_k_cgm(wm_ctx* ctx, int nxs, int nxe) {
  const Geo& g = ctx->g;
  const long long n = (long long)(nxe - nxs + 1) * g.nyl * g.nzl;
  static const bool no_coop = getenv("WM_CG_HOSTLOOP") != nullptr;
  const bool peer = ctx->nranks > 1 && ctx->peer_ok;
  if ((ctx->nranks == 1 || peer) && !no_coop && n < (1LL << 30)) {
    // the whole solve is one cooperative launch per GPU: single rank (periodic or walls), or slab-decomposed with the halo
    // planes and the all-reduces exchanged over NVLink peer memory inside the kernel (k_cgm_coop<true>)
    static int per_sm[2] = {0, 0};
    void* kern = peer ? (void*)k_cgm_coop<true> : (void*)k_cgm_coop<false>;
    if (per_sm[peer] == 0) WM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[peer], kern, TPB, 0));
    int nsm = 0;
    WM_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device));
    // blocks per SM: all that fit -- unless the sort runs beside this kernel on the second stream (wm_overlap_bps, wm_internal.cuh)
    int bps = per_sm[peer];
    static const int bps_env = getenv("WM_CG_BPS") ? atoi(getenv("WM_CG_BPS")) : 0;
    const int ov_bps = wm_overlap_bps(ctx, n);
    if (ov_bps > 0) bps = std::min(bps, ov_bps);
    if (bps_env > 0) bps = std::min(bps, bps_env);
    int nb = (int)std::min<long long>((long long)nsm * bps, (n + TPB - 1) / TPB);
    nb = std::max(1, std::min(nb, 1024));   // 2 ping-pong partial buffers of 2*nb doubles in ctx->red
    double *df = ctx->df, *gkl = ctx->gkl, *phi = ctx->phi, *p = ctx->pcg, *p2 = ctx->pcg2, *r = ctx->rcg, *b = ctx->bcg, *ap = ctx->apcg;
    double* part = ctx->red;
    int* ite_dev = ctx->totals + 6;
    int* flags = ctx->flags;
    Geo gg = g;
    PeerCG pc = ctx->peer;
    void* args[] = {&df, &gkl, &phi, &p, &p2, &r, &b, &ap, &part, &ite_dev, &flags, &gg, &nxs, &nxe, &pc};
    WM_CUDA(cudaLaunchCooperativeKernel(kern, dim3(nb), dim3(TPB), args, 0, ctx->stream));
    ctx->launches++;
    ctx->cg_ite_on_device = true;
    return WM_OK;
  }
  ctx->cg_ite_on_device = false;
  const int nb = grid_for(n);
  double* part = ctx->red;           // 2*nb partials
  double* S = ctx->red + 2 * 2048;   // S[0]=sumr S[1]=sum2 S[2]=sum1 S[3]=sum(b^2)
  double* h = ctx->red_host;
  const int ite_max = 100;
  const double err = 1e-6;
  cudaStream_t st = ctx->stream;
  static const bool generic = getenv("WM_CG_GENERIC") != nullptr;
  if (ctx->nranks > 1 && !generic) {
    // slab-decomposed solve: see the comment above block_sum2_fold
    unsigned* counter = reinterpret_cast<unsigned*>(ctx->totals + 12);
    for (int l = 0; l < 3; ++l) {
      int ite = 0;
      k_cgw_init<<<nb, TPB, 0, st>>>(ctx->df, ctx->gkl, ctx->phi, ctx->bcg, part, S, counter, g, nxs, nxe, l);
      WM_LAUNCH_CHECK(ctx);
      WM_TRY(wm_comm_allreduce_sum(ctx, S + 3, 1));
      WM_TRY(halo_slab(ctx, ctx->phi, nxs, nxe));
      k_cgw_r0<<<nb, TPB, 0, st>>>(ctx->phi, ctx->bcg, ctx->rcg, ctx->pcg, part, S, counter, g, nxs, nxe, l);
      WM_LAUNCH_CHECK(ctx);
      WM_TRY(wm_comm_allreduce_sum(ctx, S, 1));
      WM_CUDA(cudaMemcpyAsync(h, S, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
      WM_CUDA(cudaStreamSynchronize(st));
      double sum_g = h[3];
      const double eps = sqrt(sum_g) * err;
      double sumr_g = h[0];
      if (sqrt(sumr_g) > eps) {
        while (sum_g > eps) {
          ite = ite + 1;
          WM_TRY(halo_slab(ctx, ctx->pcg, nxs, nxe));
          k_cgw_ap<<<nb, TPB, 0, st>>>(ctx->pcg, ctx->rcg, ctx->apcg, part, S, counter, g, nxs, nxe, l);
          WM_LAUNCH_CHECK(ctx);
          WM_TRY(wm_comm_allreduce_sum(ctx, S, 2));
          WM_CUDA(cudaMemcpyAsync(h, S, sizeof(double), cudaMemcpyDeviceToHost, st));
          k_cgw_update<<<nb, TPB, 0, st>>>(ctx->phi, ctx->rcg, ctx->pcg, ctx->apcg, S, part, counter, g, nxs, nxe);
          WM_LAUNCH_CHECK(ctx);
          if (ite >= ite_max) {
            WM_CUDA(cudaStreamSynchronize(st));
            wm_set_error("********** stop at cgm after ite_max **********");
            return WM_ERR_CG_ITEMAX;
          }
          WM_TRY(wm_comm_allreduce_sum(ctx, S + 2, 1));
          k_cg_p<<<nb, TPB, 0, st>>>(ctx->pcg, ctx->rcg, S, g, nxs, nxe);
          WM_LAUNCH_CHECK(ctx);
          WM_CUDA(cudaStreamSynchronize(st));
          sumr_g = h[0];
          sum_g = sqrt(sumr_g);
        }
      }
      k_cg_store<<<nb, TPB, 0, st>>>(ctx->df, ctx->phi, g, nxs, nxe, l);
      WM_LAUNCH_CHECK(ctx);
      ctx->cg_ite[l] = ite;
    }
    return WM_OK;
  }
  for (int l = 0; l < 3; ++l) {
    int ite = 0;
    k_cg_init<<<nb, TPB, 0, st>>>(ctx->df, ctx->gkl, ctx->phi, ctx->bcg, part, g, nxs, nxe, l);
    WM_LAUNCH_CHECK(ctx);
    k_fold<<<1, TPB, 0, st>>>(part, nb, S, 3, -1);
    WM_LAUNCH_CHECK(ctx);
    WM_TRY(wm_comm_allreduce_sum(ctx, S + 3, 1));
    WM_TRY(bc_phi(ctx, ctx->phi, nxs, nxe, l + 1));
    k_cg_r0<<<nb, TPB, 0, st>>>(ctx->phi, ctx->bcg, ctx->rcg, ctx->pcg, part, g, nxs, nxe);
    WM_LAUNCH_CHECK(ctx);
    k_fold<<<1, TPB, 0, st>>>(part, nb, S, 0, -1);
    WM_LAUNCH_CHECK(ctx);
    WM_TRY(wm_comm_allreduce_sum(ctx, S, 1));
    WM_CUDA(cudaMemcpyAsync(h, S, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    WM_CUDA(cudaStreamSynchronize(st));
    double sum_g = h[3];
    const double eps = sqrt(sum_g) * err;
    double sumr_g = h[0];
    if (sqrt(sumr_g) > eps) {
      while (sum_g > eps) {
        ite = ite + 1;
        WM_TRY(bc_phi(ctx, ctx->pcg, nxs, nxe, l + 1));
        k_cg_ap<<<nb, TPB, 0, st>>>(ctx->pcg, ctx->rcg, ctx->apcg, part, g, nxs, nxe);
        WM_LAUNCH_CHECK(ctx);
        k_fold<<<1, TPB, 0, st>>>(part, nb, S, 0, 1);
        WM_LAUNCH_CHECK(ctx);
        WM_TRY(wm_comm_allreduce_sum(ctx, S, 2));
        WM_CUDA(cudaMemcpyAsync(h, S, sizeof(double), cudaMemcpyDeviceToHost, st));
        k_cg_update<<<nb, TPB, 0, st>>>(ctx->phi, ctx->rcg, ctx->pcg, ctx->apcg, S, part, g, nxs, nxe);
        WM_LAUNCH_CHECK(ctx);
        if (ite >= ite_max) {
          WM_CUDA(cudaStreamSynchronize(st));
          wm_set_error("********** stop at cgm after ite_max **********");
          return WM_ERR_CG_ITEMAX;
        }
        k_fold<<<1, TPB, 0, st>>>(part, nb, S, 2, -1);
        WM_LAUNCH_CHECK(ctx);
        WM_TRY(wm_comm_allreduce_sum(ctx, S + 2, 1));
        k_cg_p<<<nb, TPB, 0, st>>>(ctx->pcg, ctx->rcg, S, g, nxs, nxe);
        WM_LAUNCH_CHECK(ctx);
        WM_CUDA(cudaStreamSynchronize(st));
        sumr_g = h[0];
        sum_g = sqrt(sumr_g);
      }
    }
    k_cg_store<<<nb, TPB, 0, st>>>(ctx->df, ctx->phi, g, nxs, nxe, l);
    WM_LAUNCH_CHECK(ctx);
    ctx->cg_ite[l] = ite;
  }
  return WM_OK;
}

// boundary_*__mom: fold the one-cell ghost layer of the moment boxes into the owning cells, x first, then y, then z
// (3d/common/boundary_periodic.f90:1102-1235 [2d :571-636]; walls fold x onto the same side:
//  3d/proj/reconnection/boundary_reconnection.f90:1122-1258, 2d :582-647)
namespace {
__global__ void k_x_fold_mom(double* __restrict__ arr, Geo g, int j0, int j1, int k0, int k1) {
  const int nj = j1 - j0 + 1, nk = k1 - k0 + 1;
  const int n = nj * nk * 7;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int c = e % 7, j = j0 + (e / 7) % nj, k = k0 + (e / 7) / nj;
    double* row = arr + g.box(g.nxgs - 2, j, k) * 7 + c;
    auto X = [&](int i) -> double& { return row[(size_t)(i - (g.nxgs - 2)) * 7]; };
    if (g.bc == WM_BC_PERIODIC) {
      X(g.nxgs) = X(g.nxgs) + X(g.nxge + 1);
      X(g.nxge) = X(g.nxge) + X(g.nxgs - 1);
    } else {
      X(g.nxgs) = X(g.nxgs) + X(g.nxgs - 1);
      X(g.nxge) = X(g.nxge) + X(g.nxge + 1);
    }
  }
}

}  // namespace

int wm_k_mom_fold(wm_ctx* ctx) {
  const Geo& g = ctx->g;
  const int k0 = g.dim == 3 ? g.nzs - 1 : 0, k1 = g.dim == 3 ? g.nze + 1 : 0;
  for (int isp = 0; isp < g.nsp; ++isp) {
    double* a = ctx->mom + (size_t)isp * g.nbox() * 7;
    const int n = (g.nyl + 2) * (k1 - k0 + 1) * 7;
    k_x_fold_mom<<<wm_blocks(n, TPB), TPB, 0, ctx->stream>>>(a, g, g.nys - 1, g.nye + 1, k0, k1);
    WM_LAUNCH_CHECK(ctx);
    WM_TRY((exchange<7, true>(ctx, a, 1, true, g.nys - 1, g.nye, 1, g.nxgs - 1, g.nxge + 1, k0, k1)));
    WM_TRY((exchange<7, true>(ctx, a, 1, false, g.nye + 1, g.nys, 1, g.nxgs - 1, g.nxge + 1, k0, k1)));
    if (g.dim == 3) {
      WM_TRY((exchange<7, true>(ctx, a, 2, true, g.nzs - 1, g.nze, 1, g.nxgs - 1, g.nxge + 1, g.nys, g.nye)));
      WM_TRY((exchange<7, true>(ctx, a, 2, false, g.nze + 1, g.nzs, 1, g.nxgs - 1, g.nxge + 1, g.nys, g.nye)));
    }
  }
  return WM_OK;
}

// Diagnostic helper (Gauss check): fold the one-cell ghost layer of a scalar box array into the
// owning cells -- the 1-layer analogue of boundary_periodic__curre's add phase (y, then z, then x).
int wm_k_scalar_fold(wm_ctx* ctx, double* a, int nxs, int nxe) {
  const Geo& g = ctx->g;
  const int k0 = g.dim == 3 ? g.nzs - 1 : 0, k1 = g.dim == 3 ? g.nze + 1 : 0;
  WM_TRY((exchange<1, true>(ctx, a, 1, true, g.nys - 1, g.nye, 1, nxs - 1, nxe + 1, k0, k1)));
  WM_TRY((exchange<1, true>(ctx, a, 1, false, g.nye + 1, g.nys, 1, nxs - 1, nxe + 1, k0, k1)));
  if (g.dim == 3) {
    WM_TRY((exchange<1, true>(ctx, a, 2, true, g.nzs - 1, g.nze, 1, nxs - 1, nxe + 1, g.nys, g.nye)));
    WM_TRY((exchange<1, true>(ctx, a, 2, false, g.nze + 1, g.nzs, 1, nxs - 1, nxe + 1, g.nys, g.nye)));
  }
  if (g.bc == WM_BC_PERIODIC) {
    const int kk0 = g.dim == 3 ? g.nzs : 0, kk1 = g.dim == 3 ? g.nze : 0;
    const int n = g.nyl * (kk1 - kk0 + 1);
    k_x_fold1<<<wm_blocks(n, TPB), TPB, 0, ctx->stream>>>(a, g, nxs, nxe, g.nys, g.nye, kk0, kk1);
    WM_LAUNCH_CHECK(ctx);
  }
  return WM_OK;
}
