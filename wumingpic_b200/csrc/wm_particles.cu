// wm_particles.cu -- particle kernels: field staging, Buneman-Boris push, Esirkepov deposit,
// x boundary.  (y/z re-binning, migration and the cell-ordered sort live in wm_sort.cu.)
//
//   particle__solv                3d/common/particle.f90:52-233   [2d/common/particle.f90:48-179]
//   ele_cur                       3d/common/field.f90:211-406     [2d/common/field.f90:189-316]
//   boundary_periodic__particle_x 3d/common/boundary_periodic.f90:68-101
//   boundary_periodic__particle_yz                               :104-455
//   sort__bucket                  3d/common/sort.f90:40-88
//
// Data layout: particles are a structure of arrays ordered by (species, k, j, x-cell); the cell
// index `cs` plays the role of the reference's cumcnt (absolute offsets, nx+1 entries per pencil).
// Each warp owns one cell at a time and its lanes walk that cell's particles, so every global
// load/store is a contiguous 256-byte run and all lanes share the same 27-cell field stencil.
#include "wm_internal.cuh"
#include "wm_push.cuh"

namespace {

constexpr int TPB = 256;

__device__ __forceinline__ void shape3(double dh, double& sm, double& s0, double& sp) {
  sm = 5e-1 * (5e-1 - dh) * (5e-1 - dh);
  s0 = 7.5e-1 - dh * dh;
  sp = 5e-1 * (5e-1 + dh) * (5e-1 + dh);
}

// ---------------------------------------------------------------------------------------------
// K1: fields at (i+1/2, j+1/2, k+1/2)   particle.f90:75-91 [2d :71-83]
// ---------------------------------------------------------------------------------------------
__global__ void k_tmpf(const double* __restrict__ uf, double* __restrict__ tmpf, Geo g, int nxs, int nxe) {
  const int nxr = nxe - nxs + 3;
  const int nyr = g.nyl + 2, nzr = g.dim == 3 ? g.nzl + 2 : 1;
  const long long n = (long long)nxr * nyr * nzr;
  const size_t sx = 6, sy = (size_t)g.bx * 6, sz = (size_t)g.bx * g.by * 6;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int i = nxs - 1 + (int)(e % nxr);
    long long r = e / nxr;
    int j = g.nys - 1 + (int)(r % nyr);
    int k = g.dim == 3 ? g.nzs - 1 + (int)(r / nyr) : 0;
    const size_t o = g.box(i, j, k) * 6;
    const double* u = uf + o;
    double* t = tmpf + o;
    if (g.dim == 3) {
      t[0] = 2.5e-1 * (+u[0] + u[0 + sy] + u[0 + sz] + u[0 + sy + sz]);
      t[1] = 2.5e-1 * (+u[1] + u[1 + sx] + u[1 + sz] + u[1 + sx + sz]);
      t[2] = 2.5e-1 * (+u[2] + u[2 + sx] + u[2 + sy] + u[2 + sx + sy]);
      t[3] = 5e-1 * (+u[3] + u[3 + sx]);
      t[4] = 5e-1 * (+u[4] + u[4 + sy]);
      t[5] = 5e-1 * (+u[5] + u[5 + sz]);
    } else {
      t[0] = 0.5 * (+u[0] + u[0 + sy]);
      t[1] = 0.5 * (+u[1] + u[1 + sx]);
      t[2] = 0.25 * (+u[2] + u[2 + sx] + u[2 + sy] + u[2 + sx + sy]);
      t[3] = 0.5 * (+u[3] + u[3 + sx]);
      t[4] = 0.5 * (+u[4] + u[4 + sy]);
      t[5] = u[5];
    }
  }
}

// decode the c-th active cell (x fastest) -> (i,j,k)
__device__ __forceinline__ void active_cell(const Geo& g, long long c, int nxs, int nxr, int& i, int& j, int& k) {
  i = nxs + (int)(c % nxr);
  long long r = c / nxr;
  j = g.nys + (int)(r % g.nyl);
  k = g.dim == 3 ? g.nzs + (int)(r / g.nyl) : 0;
}

// 27-point (9-point in 2-D) gather of the six staged field components, in the reference's
// nesting: x-sum, then *shy summed over y, then *shz summed over z.
template <int D>
__device__ __forceinline__ void gather(const double* __restrict__ T, const Geo& g, const double sx[3],
                                       const double sy[3], const double sz[3], double f[6]) {
  const long long sY = (long long)g.bx * 6, sZ = (long long)g.bx * g.by * 6;
  if (D == 3) {
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
      double pl[6];
#pragma unroll
      for (int jj = 0; jj < 3; ++jj) {
        const double* t = T + (kk - 1) * sZ + (jj - 1) * sY - 6;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          double row = +__ldg(t + c) * sx[0] + __ldg(t + 6 + c) * sx[1] + __ldg(t + 12 + c) * sx[2];
          pl[c] = jj == 0 ? row * sy[0] : pl[c] + row * sy[jj];
        }
      }
#pragma unroll
      for (int c = 0; c < 6; ++c) f[c] = kk == 0 ? pl[c] * sz[0] : f[c] + pl[c] * sz[kk];
    }
  } else {
#pragma unroll
    for (int jj = 0; jj < 3; ++jj) {
      const double* t = T + (jj - 1) * sY - 6;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double row = +__ldg(t + c) * sx[0] + __ldg(t + 6 + c) * sx[1] + __ldg(t + 12 + c) * sx[2];
        f[c] = jj == 0 ? row * sy[0] : f[c] + row * sy[jj];
      }
    }
  }
}

// Buneman-Boris update of one particle   particle.f90:186-217
__device__ __forceinline__ void boris(const double f[6], double fac1, double fac2, double txxx, double c, double delt,
                                      double& ux, double& uy, double& uz, double& gam_out) {
  const double bpx = f[0], bpy = f[1], bpz = f[2], epx = f[3], epy = f[4], epz = f[5];
  double uvm1 = ux + fac1 * epx;
  double uvm2 = uy + fac1 * epy;
  double uvm3 = uz + fac1 * epz;
  double gam = sqrt(c * c + uvm1 * uvm1 + uvm2 * uvm2 + uvm3 * uvm3);
  double igam = 1.0 / gam;
  double fac1r = fac1 * igam;
  double fac2r = fac2 / (gam + txxx * (bpx * bpx + bpy * bpy + bpz * bpz) * igam);
  double uvm4 = uvm1 + fac1r * (+uvm2 * bpz - uvm3 * bpy);
  double uvm5 = uvm2 + fac1r * (+uvm3 * bpx - uvm1 * bpz);
  double uvm6 = uvm3 + fac1r * (+uvm1 * bpy - uvm2 * bpx);
  uvm1 = uvm1 + fac2r * (+uvm5 * bpz - uvm6 * bpy);
  uvm2 = uvm2 + fac2r * (+uvm6 * bpx - uvm4 * bpz);
  uvm3 = uvm3 + fac2r * (+uvm4 * bpy - uvm5 * bpx);
  ux = uvm1 + fac1 * epx;
  uy = uvm2 + fac1 * epy;
  uz = uvm3 + fac1 * epz;
  gam_out = 1.0 / sqrt(1.0 + (+ux * ux + uy * uy + uz * uz) / (c * c));
}

// ---------------------------------------------------------------------------------------------
// K2: push.  One warp per cell, lanes over that cell's particles.
// ---------------------------------------------------------------------------------------------
template <int D, bool VAY>
__global__ void __launch_bounds__(TPB) k_push(Geo g, Ptcl A, Ptcl B, const int* __restrict__ cs,
                                              const double* __restrict__ tmpf, int nxs, int nxe) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nxr = nxe - nxs + 1;
  const long long ncell = (long long)nxr * g.nyl * g.nzl;
  constexpr int U = D;  // index of ux in the SoA
  for (long long cidx = warp; cidx < ncell; cidx += nwarps) {
    int i, j, k;
    active_cell(g, cidx, nxs, nxr, i, j, k);
    const double* T = tmpf + g.box(i, j, k) * 6;
    for (int isp = 0; isp < g.nsp; ++isp) {
      const int* row = cs + (size_t)g.pen(j, k, isp) * (g.nx + 1) + (i - g.nxgs);
      const int beg = row[0], end = row[1];
      const double fac1 = g.q[isp] / g.r[isp] * 5e-1 * g.delt;
      const double txxx = fac1 * fac1;
      const double fac2 = g.q[isp] * g.delt / g.r[isp];
      for (int p = beg + lane; p < end; p += 32) {
        const double x = A.c[0][p], y = A.c[1][p];
        const double z = D == 3 ? A.c[2][p] : 0.0;
        double ux = A.c[U][p], uy = A.c[U + 1][p], uz = A.c[U + 2][p];
        double sx[3], sy[3], sz[3] = {0, 1, 0};
        shape3(x * g.d_delx - 5e-1 - i, sx[0], sx[1], sx[2]);
        shape3(y * g.d_delx - 5e-1 - j, sy[0], sy[1], sy[2]);
        if (D == 3) shape3(z * g.d_delx - 5e-1 - k, sz[0], sz[1], sz[2]);
        double f[6];
        gather<D>(T, g, sx, sy, sz, f);
        double gam;
        if (VAY) wm_vay_update(f, fac1, fac2, g.c, ux, uy, uz, gam);   // particle__solv_vay
        else boris(f, fac1, fac2, txxx, g.c, g.delt, ux, uy, uz, gam);
        B.c[U][p] = ux;
        B.c[U + 1][p] = uy;
        B.c[U + 2][p] = uz;
        B.c[0][p] = x + ux * g.delt * gam;
        B.c[1][p] = y + uy * g.delt * gam;
        if (D == 3) B.c[2][p] = z + uz * g.delt * gam;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K5: Esirkepov deposit (v1: one particle per lane, non-zero W terms go straight to global RED.F64)
// field.f90:252-396.  s0 is taken about the LOOP cell (from cs), s1 about int(x_new).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void s0ds(double xo, double xn, int cell, double d_delx, double s0[5], double ds[5],
                                     int* flags) {
  double dh = xo * d_delx - 5e-1 - cell;
  s0[0] = 0.0;
  shape3(dh, s0[1], s0[2], s0[3]);
  s0[4] = 0.0;
  int i2 = (int)(xn * d_delx);
  dh = xn * d_delx - 5e-1 - i2;
  int inc = i2 - cell;
  if (inc < -1 || inc > 1) {
    atomicOr(flags, 2);
    inc = inc < 0 ? -1 : 1;
  }
  double s1_1, s1_2, s1_3;
  shape3(dh, s1_1, s1_2, s1_3);
  const double smo_1 = inc == -1 ? 1.0 : 0.0, smo_2 = inc == 0 ? 1.0 : 0.0, smo_3 = inc == 1 ? 1.0 : 0.0;
  ds[0] = s1_1 * smo_1;
  ds[1] = s1_1 * smo_2 + s1_2 * smo_1;
  ds[2] = s1_2 * smo_2 + s1_3 * smo_1 + s1_1 * smo_3;
  ds[3] = s1_3 * smo_2 + s1_2 * smo_3;
  ds[4] = s1_3 * smo_3;
#pragma unroll
  for (int m = 0; m < 5; ++m) ds[m] = ds[m] - s0[m];
}

__device__ __forceinline__ void red_add(double* addr, double v) {
  if (v != 0.0) atomicAdd(addr, v);
}

template <int D>
__global__ void __launch_bounds__(TPB) k_deposit(Geo g, Ptcl A, Ptcl B, const int* __restrict__ cs,
                                                 double* __restrict__ uj, int* flags, int nxs, int nxe) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nxr = nxe - nxs + 1;
  const long long ncell = (long long)nxr * g.nyl * g.nzl;
  const double fac = 1.0 / 3.0;
  const long long sY = (long long)g.bx * 3, sZ = (long long)g.bx * g.by * 3;
  for (long long cidx = warp; cidx < ncell; cidx += nwarps) {
    int i, j, k;
    active_cell(g, cidx, nxs, nxr, i, j, k);
    double* J0 = uj + g.box(i, j, k) * 3;
    for (int isp = 0; isp < g.nsp; ++isp) {
      const int* row = cs + (size_t)g.pen(j, k, isp) * (g.nx + 1) + (i - g.nxgs);
      const int beg = row[0], end = row[1];
      const double qdxdt = g.q[isp] * g.delx * g.d_delt;
      for (int p = beg + lane; p < end; p += 32) {
        double s0x[5], s0y[5], dsx[5], dsy[5];
        s0ds(A.c[0][p], B.c[0][p], i, g.d_delx, s0x, dsx, flags);
        s0ds(A.c[1][p], B.c[1][p], j, g.d_delx, s0y, dsy, flags);
        if (D == 3) {
          double s0z[5], dsz[5];
          s0ds(A.c[2][p], B.c[2][p], k, g.d_delx, s0z, dsz, flags);
#pragma unroll
          for (int kp = 0; kp < 5; ++kp) {
#pragma unroll
            for (int jp = 0; jp < 5; ++jp) {
              // pjx(ip, jp, kp)
              double dstmp = ((s0y[jp] + 5e-1 * dsy[jp]) * s0z[kp] + (5e-1 * s0y[jp] + fac * dsy[jp]) * dsz[kp]) * qdxdt;
              double* Jx = J0 + (kp - 2) * sZ + (jp - 2) * sY;
              double pj = -dsx[0] * dstmp;
              red_add(Jx + (-1) * 3 + 0, pj);
              pj = pj - dsx[1] * dstmp;
              red_add(Jx + 0 * 3 + 0, pj);
              pj = pj - dsx[2] * dstmp;
              red_add(Jx + 1 * 3 + 0, pj);
              pj = pj - dsx[3] * dstmp;
              red_add(Jx + 2 * 3 + 0, pj);
              // pjy(run, ip=jp, kp): x offset = jp-2, y offset = run, z offset = kp-2
              dstmp = ((s0x[jp] + 5e-1 * dsx[jp]) * s0z[kp] + (5e-1 * s0x[jp] + fac * dsx[jp]) * dsz[kp]) * qdxdt;
              double* Jy = J0 + (kp - 2) * sZ + (jp - 2) * 3 + 1;
              pj = -dsy[0] * dstmp;
              red_add(Jy + (-1) * sY, pj);
              pj = pj - dsy[1] * dstmp;
              red_add(Jy, pj);
              pj = pj - dsy[2] * dstmp;
              red_add(Jy + sY, pj);
              pj = pj - dsy[3] * dstmp;
              red_add(Jy + 2 * sY, pj);
              // pjz(run, ip=jp, jp=kp): x offset = jp-2, y offset = kp-2, z offset = run
              dstmp = ((s0x[jp] + 5e-1 * dsx[jp]) * s0y[kp] + (5e-1 * s0x[jp] + fac * dsx[jp]) * dsy[kp]) * qdxdt;
              double* Jz = J0 + (kp - 2) * sY + (jp - 2) * 3 + 2;
              pj = -dsz[0] * dstmp;
              red_add(Jz + (-1) * sZ, pj);
              pj = pj - dsz[1] * dstmp;
              red_add(Jz, pj);
              pj = pj - dsz[2] * dstmp;
              red_add(Jz + sZ, pj);
              pj = pj - dsz[3] * dstmp;
              red_add(Jz + 2 * sZ, pj);
            }
          }
        } else {
          // 2-D (2d/common/field.f90:270-298): Jx, Jy by 2-D Esirkepov, Jz = q*vz*(S0S0 + ...)
          const double ux = B.c[2][p], uy = B.c[3][p], uz = B.c[4][p];
          const double gvz = uz / sqrt(1.0 + (+ux * ux + uy * uy + uz * uz) / (g.c * g.c));
          const double qq = g.q[isp];
#pragma unroll
          for (int jp = 0; jp < 5; ++jp) {
            double pj = 0.0;
            double* Jx = J0 + (jp - 2) * sY;
#pragma unroll
            for (int ip = 0; ip < 4; ++ip) {
              pj = pj - qq * g.delx * g.d_delt * dsx[ip] * (s0y[jp] + 0.5 * dsy[jp]);
              red_add(Jx + (ip - 1) * 3 + 0, pj);
            }
          }
#pragma unroll
          for (int ip = 0; ip < 5; ++ip) {
            double pj = 0.0;
            double* Jy = J0 + (ip - 2) * 3 + 1;
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
              pj = pj - qq * g.delx * g.d_delt * dsy[jp] * (s0x[ip] + 0.5 * dsx[ip]);
              red_add(Jy + (jp - 1) * sY, pj);
            }
          }
#pragma unroll
          for (int jp = 0; jp < 5; ++jp)
#pragma unroll
            for (int ip = 0; ip < 5; ++ip)
              red_add(J0 + (jp - 2) * sY + (ip - 2) * 3 + 2,
                      qq * gvz * (+s0x[ip] * s0y[jp] + 0.5 * dsx[ip] * s0y[jp] + 0.5 * s0x[ip] * dsy[jp]
                                  + fac * dsx[ip] * dsy[jp]));
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Moments: mom_calc__accl + mom_calc__nvt in one pass, no particle ever leaves the device.
//   mom_calc__accl  3d/common/mom_calc.f90:49-216 [2d :49-163]  the gather + Boris rotation of particle__solv with
//                   delt/2 (mom_calc__init :36) and no move: momenta re-centred to the time level of the positions
//   mom_calc__nvt   3d/common/mom_calc.f90:219-332 [2d :166-252] CIC deposit of N, V, T at (i+1/2, j+1/2, k+1/2);
//                   3-D: ih = floor(x*d_delx - 1/2), dx = x - 1/2 - ih (no d_delx, :246-251); 2-D: ih = int(x*d_delx - 1/2),
//                   dx = x*d_delx - 1/2 - ih
// mom lives on the box layout (two ghost layers, 7 components fastest), one box per species.  Output-cadence code:
// one RED.F64 per (moment, node).
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(TPB) k_mom(Geo g, Ptcl A, const int* __restrict__ cs, const double* __restrict__ tmpf,
                                             double* __restrict__ mom, int nxs, int nxe) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nxr = nxe - nxs + 1;
  const long long ncell = (long long)nxr * g.nyl * g.nzl;
  constexpr int U = D;
  const double delth = g.delt * 5e-1;
  const long long sY = (long long)g.bx * 7, sZ = (long long)g.bx * g.by * 7;
  for (long long cidx = warp; cidx < ncell; cidx += nwarps) {
    int i, j, k;
    active_cell(g, cidx, nxs, nxr, i, j, k);
    const double* T = tmpf + g.box(i, j, k) * 6;
    for (int isp = 0; isp < g.nsp; ++isp) {
      const int* row = cs + (size_t)g.pen(j, k, isp) * (g.nx + 1) + (i - g.nxgs);
      const int beg = row[0], end = row[1];
      const double fac1 = g.q[isp] / g.r[isp] * 5e-1 * delth;
      const double txxx = fac1 * fac1;
      const double fac2 = g.q[isp] * delth / g.r[isp];
      double* M = mom + (size_t)isp * g.nbox() * 7;
      for (int p = beg + lane; p < end; p += 32) {
        const double x = A.c[0][p], y = A.c[1][p];
        const double z = D == 3 ? A.c[2][p] : 0.0;
        double ux = A.c[U][p], uy = A.c[U + 1][p], uz = A.c[U + 2][p];
        double sx[3], sy[3], sz[3] = {0, 1, 0};
        shape3(x * g.d_delx - 5e-1 - i, sx[0], sx[1], sx[2]);
        shape3(y * g.d_delx - 5e-1 - j, sy[0], sy[1], sy[2]);
        if (D == 3) shape3(z * g.d_delx - 5e-1 - k, sz[0], sz[1], sz[2]);
        double f[6];
        gather<D>(T, g, sx, sy, sz, f);
        double gam;
        boris(f, fac1, fac2, txxx, g.c, delth, ux, uy, uz, gam);
        int ih, jh, kh = 0;
        double wx[2], wy[2], wz[2] = {1.0, 0.0};
        if (D == 3) {
          ih = (int)floor(x * g.d_delx - 5e-1); jh = (int)floor(y * g.d_delx - 5e-1); kh = (int)floor(z * g.d_delx - 5e-1);
          wx[1] = x - 5e-1 - ih; wy[1] = y - 5e-1 - jh; wz[1] = z - 5e-1 - kh;
          wz[0] = 1.0 - wz[1];
        } else {
          ih = (int)(x * g.d_delx - 0.5); jh = (int)(y * g.d_delx - 0.5);
          wx[1] = x * g.d_delx - 0.5 - ih; wy[1] = y * g.d_delx - 0.5 - jh;
        }
        wx[0] = 1.0 - wx[1]; wy[0] = 1.0 - wy[1];
        const double val[7] = {1.0, ux * gam, uy * gam, uz * gam, ux * ux * gam, uy * uy * gam, uz * uz * gam};
        double* base = M + g.box(ih, jh, D == 3 ? kh : 0) * 7;
#pragma unroll
        for (int c = 0; c < (D == 3 ? 2 : 1); ++c)
#pragma unroll
          for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int a = 0; a < 2; ++a) {
              double* node = base + c * sZ + b * sY + a * 7;
              const double w3 = D == 3 ? wx[a] * wy[b] * wz[c] : wx[a] * wy[b];
              atomicAdd(node, w3);
#pragma unroll
              for (int l = 1; l < 7; ++l) atomicAdd(node + l, D == 3 ? val[l] * wx[a] * wy[b] * wz[c] : val[l] * wx[a] * wy[b]);
            }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K12: x boundary on the pushed set
//   periodic     3d/common/boundary_periodic.f90:68-101 (int(x*d_delx), round-to-nearest)
//                2d/common/boundary_periodic.f90:61-96  (int(x/delx) and the wrap under ieee_down, :74)
//   reflecting   {2d,3d}/proj/reconnection/boundary_reconnection.f90:61-99 / :69-110, 2d/proj/shock/boundary_shock.f90:62-100
//   injection    2d/proj/shock/boundary_shock.f90:255-297, 3d/proj/shock/boundary_shock.f90:424-469
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void k_bc_x_periodic(Geo g, double* __restrict__ x, long long n) {
  const double len = D == 2 ? __dmul_rd((double)(g.nxge - g.nxgs + 1), g.delx) : (g.nxge - g.nxgs + 1) * g.delx;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
    double xx = x[p];
    int ipos = D == 2 ? (int)__ddiv_rd(xx, g.delx) : (int)(xx * g.d_delx);
    if (ipos < g.nxgs) x[p] = D == 2 ? __dadd_rd(xx, len) : xx + len;
    else if (ipos >= g.nxge + 1) x[p] = D == 2 ? __dadd_rd(xx, -len) : xx - len;
  }
}

// kind 1: reflecting walls at (nxs+1) delx and (nxe-1) delx; kind 2: left wall + moving right wall xend (injection)
template <int D>
__global__ void k_bc_x_walls(Geo g, Ptcl P, long long n, int nxs, int nxe, int kind, double u0) {
  constexpr int U = D;
  const double xend = nxe * g.delx + u0 / sqrt(1.0 + (u0 * u0) / (g.c * g.c)) * g.delt;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
    const double xx = P.c[0][p];
    // the 2-D modules and the 3-D reconnection module divide by delx; the 3-D shock module multiplies by d_delx
    const int ipos = (D == 3 && kind == 2) ? (int)(xx * g.d_delx) : (int)(xx / g.delx);
    if (ipos < nxs + 1) {
      P.c[0][p] = 2.0 * (nxs + 1) * g.delx - xx;
      P.c[U][p] = -P.c[U][p];
      P.c[U + 1][p] = -P.c[U + 1][p];
      P.c[U + 2][p] = -P.c[U + 2][p];
    } else if (kind == 1 ? ipos >= nxe - 1 : xx > xend) {
      if (kind == 1) {
        P.c[0][p] = 2.0 * (nxe - 1) * g.delx - xx;
        P.c[U][p] = -P.c[U][p];
      } else {
        P.c[0][p] = 2.0 * xend - xx;
        P.c[U][p] = 2.0 * u0 - P.c[U][p];
      }
      P.c[U + 1][p] = -P.c[U + 1][p];
      P.c[U + 2][p] = -P.c[U + 2][p];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host <-> device layout conversion: the reference's padded AoS pencils <-> SoA
// stage holds [pencil][slot < maxcnt][ndim]
// ---------------------------------------------------------------------------------------------
// swap (3-D y-slabs, wm_internal.cuh): the stage holds the CALLER's pencils (order (isp, k, j), columns x y z ux uy uz id); the device
// pencil of caller pencil (isp, k, j) is (isp, k' = j, j' = k) and the record columns y <-> z, uy <-> uz change places
__device__ __forceinline__ int wm_host_pen_to_dev(const Geo& g, int pen_h) {
  const int nyl_h = g.nzl, nzl_h = g.nyl;                      // the caller's slab extents
  const int jj = pen_h % nyl_h, kk = (pen_h / nyl_h) % nzl_h, isp = pen_h / (nyl_h * nzl_h);
  return (isp * g.nzl + jj) * g.nyl + kk;
}
__device__ __forceinline__ int wm_swap_col(int c) { return c == 1 ? 2 : (c == 2 ? 1 : (c == 4 ? 5 : (c == 5 ? 4 : c))); }

__global__ void k_aos_to_soa(Geo g, const double* __restrict__ stage, Ptcl dst, double* __restrict__ dst_id,
                             const int* __restrict__ poff, const int* __restrict__ np2, int pen0, int npens, int maxcnt, int swap) {
  const long long n = (long long)npens * maxcnt;
  const int nd = g.ndim;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int pl = (int)(e / maxcnt), ii = (int)(e % maxcnt);
    int pen = swap ? wm_host_pen_to_dev(g, pen0 + pl) : pen0 + pl;
    if (ii >= np2[pen]) continue;
    const double* s = stage + e * nd;
    size_t d = (size_t)poff[pen] + ii;
    for (int c = 0; c < nd - 1; ++c) dst.c[swap ? wm_swap_col(c) : c][d] = s[c];
    dst_id[d] = s[nd - 1];
  }
}

__global__ void k_soa_to_aos(Geo g, double* __restrict__ stage, Ptcl src, const double* __restrict__ src_id,
                             const int* __restrict__ poff, const int* __restrict__ np2, int pen0, int npens, int maxcnt, int swap) {
  const long long n = (long long)npens * maxcnt;
  const int nd = g.ndim;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int pl = (int)(e / maxcnt), ii = (int)(e % maxcnt);
    int pen = swap ? wm_host_pen_to_dev(g, pen0 + pl) : pen0 + pl;
    if (ii >= np2[pen]) continue;
    double* s = stage + e * nd;
    size_t d = (size_t)poff[pen] + ii;
    for (int c = 0; c < nd - 1; ++c) s[c] = src.c[swap ? wm_swap_col(c) : c][d];
    s[nd - 1] = src_id[d];
  }
}

// field translation of the y-slab relabelling, both directions (it is an involution): out(c', i, a, b) = sign(c') in(c, i, b, a) with
// c' = (0 2 1 3 5 4)[c], minus on the three B components; `in` has the OTHER system's box extents (its y extent = our z extent)
__global__ void k_swap_box6(const double* __restrict__ in, double* __restrict__ out, int bx, int by_out, int bz_out) {
  const long long n = (long long)bx * by_out * bz_out;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e % bx), a = (int)((e / bx) % by_out), b = (int)(e / ((long long)bx * by_out));
    const double* s = in + (((long long)a * bz_out + b) * bx + i) * 6;       // (i, y_in = b, z_in = a): in's y extent is bz_out
    double* d = out + e * 6;
    d[0] = -s[0]; d[1] = -s[2]; d[2] = -s[1];
    d[3] = s[3];  d[4] = s[5];  d[5] = s[4];
  }
}

int grid_for(long long n) {
  long long b = (n + TPB - 1) / TPB;
  const long long cap = 148LL * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

int wm_k_tmpf(wm_ctx* ctx, int nxs, int nxe) {
  const Geo& g = ctx->g;
  const long long n = (long long)(nxe - nxs + 3) * (g.nyl + 2) * (g.dim == 3 ? g.nzl + 2 : 1);
  k_tmpf<<<grid_for(n), TPB, 0, ctx->stream>>>(ctx->uf, ctx->tmpf, g, nxs, nxe);
  WM_LAUNCH_CHECK(ctx);
  return WM_OK;
}

int wm_k_push(wm_ctx* ctx, int nxs, int nxe) {
  const Geo& g = ctx->g;
  const long long ncell = (long long)(nxe - nxs + 1) * g.nyl * g.nzl;
  const int blocks = (int)std::min<long long>((ncell * 32 + TPB - 1) / TPB, 148LL * 8);
  const bool vay = ctx->pusher == WM_PUSHER_VAY;
  if (g.dim == 3) {
    if (vay) k_push<3, true><<<blocks, TPB, 0, ctx->stream>>>(g, ctx->A, ctx->B, ctx->cs, ctx->tmpf, nxs, nxe);
    else k_push<3, false><<<blocks, TPB, 0, ctx->stream>>>(g, ctx->A, ctx->B, ctx->cs, ctx->tmpf, nxs, nxe);
  } else {
    if (vay) k_push<2, true><<<blocks, TPB, 0, ctx->stream>>>(g, ctx->A, ctx->B, ctx->cs, ctx->tmpf, nxs, nxe);
    else k_push<2, false><<<blocks, TPB, 0, ctx->stream>>>(g, ctx->A, ctx->B, ctx->cs, ctx->tmpf, nxs, nxe);
  }
  WM_LAUNCH_CHECK(ctx);
  return WM_OK;
}

int wm_k_mom(wm_ctx* ctx, int nxs, int nxe) {
  const Geo& g = ctx->g;
  const long long ncell = (long long)(nxe - nxs + 1) * g.nyl * g.nzl;
  const int blocks = (int)std::min<long long>((ncell * 32 + TPB - 1) / TPB, 148LL * 8);
  if (g.dim == 3) k_mom<3><<<blocks, TPB, 0, ctx->stream>>>(g, ctx->A, ctx->cs, ctx->tmpf, ctx->mom, nxs, nxe);
  else k_mom<2><<<blocks, TPB, 0, ctx->stream>>>(g, ctx->A, ctx->cs, ctx->tmpf, ctx->mom, nxs, nxe);
  WM_LAUNCH_CHECK(ctx);
  return WM_OK;
}

int wm_k_deposit(wm_ctx* ctx, int nxs, int nxe) {
  const Geo& g = ctx->g;
  const long long ncell = (long long)(nxe - nxs + 1) * g.nyl * g.nzl;
  const int blocks = (int)std::min<long long>((ncell * 32 + TPB - 1) / TPB, 148LL * 8);
  if (g.dim == 3)
    k_deposit<3><<<blocks, TPB, 0, ctx->stream>>>(g, ctx->A, ctx->B, ctx->cs, ctx->uj, ctx->flags, nxs, nxe);
  else
    k_deposit<2><<<blocks, TPB, 0, ctx->stream>>>(g, ctx->A, ctx->B, ctx->cs, ctx->uj, ctx->flags, nxs, nxe);
  WM_LAUNCH_CHECK(ctx);
  return WM_OK;
}

// kind: WM_BC_PERIODIC wrap, WM_BC_RECONNECTION reflecting walls (also boundary_shock__particle_x), WM_BC_SHOCK injection
int wm_k_bc_x(wm_ctx* ctx, int nxs, int nxe, int kind, double u0) {
  const Geo& g = ctx->g;
  if (ctx->ntot == 0) return WM_OK;
  const int blocks = grid_for(ctx->ntot);
  if (kind == WM_BC_PERIODIC) {
    if (g.dim == 3) k_bc_x_periodic<3><<<blocks, TPB, 0, ctx->stream>>>(g, ctx->B.c[0], ctx->ntot);
    else k_bc_x_periodic<2><<<blocks, TPB, 0, ctx->stream>>>(g, ctx->B.c[0], ctx->ntot);
  } else {
    const int k = kind == WM_BC_SHOCK ? 2 : 1;
    if (g.dim == 3) k_bc_x_walls<3><<<blocks, TPB, 0, ctx->stream>>>(g, ctx->B, ctx->ntot, nxs, nxe, k, u0);
    else k_bc_x_walls<2><<<blocks, TPB, 0, ctx->stream>>>(g, ctx->B, ctx->ntot, nxs, nxe, k, u0);
  }
  WM_LAUNCH_CHECK(ctx);
  return WM_OK;
}

int wm_k_aos_to_soa(wm_ctx* ctx, const double* stage, Ptcl dst, double* dst_id, int pen0, int npens, int maxcnt) {
  const long long n = (long long)npens * maxcnt;
  if (n == 0) return WM_OK;
  k_aos_to_soa<<<grid_for(n), TPB, 0, ctx->stream>>>(ctx->g, stage, dst, dst_id, ctx->poff, ctx->np2, pen0, npens, maxcnt, ctx->swap_yz ? 1 : 0);
  WM_LAUNCH_CHECK(ctx);
  return WM_OK;
}

// field box between the caller's (x, y, z) and the device's (x, z, y) system (see wm_ctx::swap_yz); to_device: in = caller layout
int wm_k_swap_box6(wm_ctx* ctx, const double* in, double* out, bool to_device) {
  const Geo& g = ctx->g;
  const int by_out = to_device ? g.by : g.bz, bz_out = to_device ? g.bz : g.by;
  k_swap_box6<<<grid_for((long long)g.nbox()), TPB, 0, ctx->stream>>>(in, out, g.bx, by_out, bz_out);
  WM_LAUNCH_CHECK(ctx);
  return WM_OK;
}

int wm_k_soa_to_aos(wm_ctx* ctx, double* stage, Ptcl src, const double* src_id, int pen0, int npens, int maxcnt) {
  const long long n = (long long)npens * maxcnt;
  if (n == 0) return WM_OK;
  k_soa_to_aos<<<grid_for(n), TPB, 0, ctx->stream>>>(ctx->g, stage, src, src_id, ctx->poff, ctx->np2, pen0, npens, maxcnt, ctx->swap_yz ? 1 : 0);
  WM_LAUNCH_CHECK(ctx);
  return WM_OK;
}
