// oracle/oracle2d.cpp -- TEST INFRASTRUCTURE ONLY (see oracle_common.h header).
//
// CPU restatement of WumingPIC's 2-D per-timestep loop, following the reference loop nests and
// array shapes 1:1 (Fortran index bases are kept through the accessor functions below).
// Pinned bit for bit to the translated reference (oracle/f2cxx, tests/test_ref_transpiled.py); see oracle_common.h.
//
//   particle__solv                 2d/common/particle.f90:48-179
//   field__init / fdtd_i           2d/common/field.f90:22-63, 66-186
//   ele_cur                        2d/common/field.f90:189-316
//   cgm                            2d/common/field.f90:319-461
//   sort__bucket                   2d/common/sort.f90:36-82
//   boundary_periodic__particle_x  2d/common/boundary_periodic.f90:61-96    (round-down mode, :74)
//   boundary_periodic__particle_y                                :99-248    (round-down mode, :124)
//   boundary_periodic__dfield                                    :251-354
//   boundary_periodic__curre                                     :357-508
//   boundary_periodic__phi                                       :511-568
//   boundary_reconnection__*       2d/proj/reconnection/boundary_reconnection.f90:61-99 (particle_x),
//                                  :254-361 (dfield), :364-502 (curre), :505-579 (phi)
//   boundary_shock__*              2d/proj/shock/boundary_shock.f90:62-100 (particle_x), :255-297 (injection),
//                                  :300-407 (dfield), :410-548 (curre), :551-625 (phi)
//   mom_calc__accl / __nvt          2d/common/mom_calc.f90:49-163, 166-252 ; boundary_*__mom 2d/common/boundary_periodic.f90:571-636
//   mpi_set (rank table, slabs)    2d/common/mpi_set.f90:21-51
//   time loops                     2d/proj/weibel/app.f90:99-107, 2d/proj/reconnection/app.f90:99-106,
//                                  2d/proj/shock/app.f90:112-118
//   Weibel initial load            2d/proj/weibel/app.f90:292-328, 380-474
//
// MPI ranks (1-D slabs in y) are emulated in-process exactly like oracle3d.cpp: every MPI_SENDRECV
// becomes one lock-step phase over all ranks, MPI_ALLREDUCE a sum in rank order.
#include "oracle_common.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

const double kPi = 4.0 * std::atan(1.0);

// ---- IEEE round-toward-minus-infinity results of single operations, built from round-to-nearest
// arithmetic + exact error terms (independent of compiler flags / FENV support).  Used where the
// reference runs under ieee_set_rounding_mode(ieee_down).
inline double down1(double s) { return std::nextafter(s, -HUGE_VAL); }
inline double add_rd(double a, double b) {
  const double s = a + b;
  const double bb = s - a;
  const double err = (a - (s - bb)) + (b - bb);  // TwoSum: a + b = s + err exactly
  return err < 0.0 ? down1(s) : s;
}
inline double mul_rd(double a, double b) {
  const double p = a * b;
  const double err = std::fma(a, b, -p);
  return err < 0.0 ? down1(p) : p;
}
inline double div_rd(double a, double b) {
  const double q = a / b;
  const double rem = std::fma(-q, b, a);  // a - q*b exactly (sign decides which side q lies)
  return ((rem < 0.0) == (b > 0.0)) && rem != 0.0 ? down1(q) : q;
}

struct Rank2;

struct World2 {
  int ndim = 6, np = 0, nsp = 2;
  int nxgs = 2, nxge = 0, nygs = 2, nyge = 0;
  int nxs = 0, nxe = 0;
  int nproc = 1;
  int bc = 0;  // 0 periodic, 1 reconnection walls, 2 shock walls
  double delx = 1, delt = 1, c = 1, gfac = 0.501, d_delx = 1, d_delt = 1;
  double q[2] = {0, 0}, r[2] = {1, 1};
  double f1 = 0, f2 = 0, f3 = 0, f4 = 0, f5 = 0;
  std::vector<Rank2> ranks;
  int cg_ite[3] = {0, 0, 0};
  int err = 0;  // 1: cgm ite_max, 2: memory over, 3: moved more than one row, 4: x outside [nxs,nxe] at the sort
  int pusher = 0;  // the pusher step() calls: 0 particle__solv (Buneman-Boris), 1 particle__solv_vay
  int nx() const { return nxge - nxgs + 1; }
};

struct Rank2 {
  const World2* w = nullptr;
  int rank = 0;
  int nys = 0, nye = 0, nyl = 0;
  int nup = 0, ndown = 0;
  std::vector<double> up, gp, uf, df, gkl, uj, mom;
  std::vector<int> np2, cumcnt;
  inline size_t im(int l, int i, int j, int isp) const {  // mom(7,nxgs-1:nxge+1,nys-1:nye+1,nsp)
    return ((((size_t)(isp - 1) * (nyl + 2) + (j - (nys - 1))) * (w->nx() + 2) + (i - (w->nxgs - 1)))) * 7 + (l - 1);
  }
  inline size_t ip(int d, int ii, int j, int isp) const {  // up/gp(ndim,np,nys:nye,nsp)
    return (((size_t)(isp - 1) * nyl + (j - nys)) * w->np + (size_t)(ii - 1)) * w->ndim + (d - 1);
  }
  inline size_t i6(int cc, int i, int j) const {  // uf/df(6,nxgs-2:nxge+2,nys-2:nye+2)
    return ((size_t)(j - (nys - 2)) * (w->nx() + 4) + (i - (w->nxgs - 2))) * 6 + (cc - 1);
  }
  inline size_t i3(int cc, int i, int j) const {  // uj(3, same box)
    return ((size_t)(j - (nys - 2)) * (w->nx() + 4) + (i - (w->nxgs - 2))) * 3 + (cc - 1);
  }
  inline size_t ig(int cc, int i, int j) const {  // gkl(3,nxgs:nxge,nys:nye)
    return ((size_t)(j - nys) * w->nx() + (i - w->nxgs)) * 3 + (cc - 1);
  }
  inline size_t ic(int i, int j, int isp) const {  // cumcnt(nxgs:nxge+1,nys:nye,nsp)
    return ((size_t)(isp - 1) * nyl + (j - nys)) * (w->nx() + 1) + (i - w->nxgs);
  }
  inline size_t in2(int j, int isp) const { return (size_t)(isp - 1) * nyl + (j - nys); }  // np2(nys:nye,nsp)
};

struct Cg2 {
  int nxs, nxe, nys, nye;
  std::vector<double> phi, p, r, b, ap;
  inline size_t i1(int i, int j) const { return (size_t)(j - (nys - 1)) * (nxe - nxs + 3) + (i - (nxs - 1)); }
  inline size_t i0(int i, int j) const { return (size_t)(j - nys) * (nxe - nxs + 1) + (i - nxs); }
};

enum Dir { TO_DOWN, TO_UP };
template <class T, class Pack, class Unpack>
void sendrecv(World2& w, Dir d, Pack pack, Unpack unpack) {
  const int R = (int)w.ranks.size();
  std::vector<std::vector<T>> snd(R);
  for (int r = 0; r < R; ++r) pack(w.ranks[r], snd[r]);
  for (int r = 0; r < R; ++r) {
    const Rank2& me = w.ranks[r];
    const int src = d == TO_DOWN ? me.nup : me.ndown;
    unpack(w.ranks[r], snd[src]);
  }
}

// ---------------------------------------------------------------------------
// particle__solv -- 2d/common/particle.f90:48-179
// ---------------------------------------------------------------------------
// accl = true: mom_calc__accl (2d/common/mom_calc.f90:49-163): same gather and rotation with delt/2 (:34), no move
// vay = true: particle__solv_vay (2d/common/particle.f90:182-315): same staging and gather, Vay velocity update (:268-292)
void particle_solv(World2& w, Rank2& R, std::vector<double>& gp, const std::vector<double>& up, bool accl = false, bool vay = false) {
  const int nxs = w.nxs, nxe = w.nxe, nys = R.nys, nye = R.nye;
  const double d_delx = w.d_delx, delt = accl ? w.delt * 0.5 : w.delt, c = w.c;
  const int tx = nxe - nxs + 3, ty = nye - nys + 3;
  std::vector<double> tmp((size_t)6 * tx * ty);
  auto T = [&](int cc, int i, int j) -> double& {
    return tmp[((size_t)(j - (nys - 1)) * tx + (i - (nxs - 1))) * 6 + (cc - 1)];
  };
  const std::vector<double>& uf = R.uf;
  // fields at (i+1/2, j+1/2) -- particle.f90:71-83
#pragma omp parallel for
  for (int j = nys - 1; j <= nye + 1; ++j)
    for (int i = nxs - 1; i <= nxe + 1; ++i) {
      T(1, i, j) = 0.5 * (+uf[R.i6(1, i, j)] + uf[R.i6(1, i, j + 1)]);
      T(2, i, j) = 0.5 * (+uf[R.i6(2, i, j)] + uf[R.i6(2, i + 1, j)]);
      T(3, i, j) = 0.25 * (+uf[R.i6(3, i, j)] + uf[R.i6(3, i + 1, j)] + uf[R.i6(3, i, j + 1)] + uf[R.i6(3, i + 1, j + 1)]);
      T(4, i, j) = 0.5 * (+uf[R.i6(4, i, j)] + uf[R.i6(4, i + 1, j)]);
      T(5, i, j) = 0.5 * (+uf[R.i6(5, i, j)] + uf[R.i6(5, i, j + 1)]);
      T(6, i, j) = uf[R.i6(6, i, j)];
    }
  // particle.f90:85-171
#pragma omp parallel for schedule(static)
  for (int j = nys; j <= nye; ++j)
    for (int i = nxs; i <= nxe; ++i)
      for (int isp = 1; isp <= w.nsp; ++isp) {
        const double fac1 = w.q[isp - 1] / w.r[isp - 1] * 0.5 * delt;
        const double txxx = fac1 * fac1;
        const double fac2 = w.q[isp - 1] * delt / w.r[isp - 1];
        const int i_beg = R.cumcnt[R.ic(i, j, isp)] + 1, i_end = R.cumcnt[R.ic(i + 1, j, isp)];
        for (int ii = i_beg; ii <= i_end; ++ii) {
          const double* u = &up[R.ip(1, ii, j, isp)];
          double* g = &gp[R.ip(1, ii, j, isp)];
          double sx[3], sy[3];
          double dh = u[0] * d_delx - 0.5 - i;
          sx[0] = 0.5 * (0.5 - dh) * (0.5 - dh);
          sx[1] = 0.75 - dh * dh;
          sx[2] = 0.5 * (0.5 + dh) * (0.5 + dh);
          dh = u[1] * d_delx - 0.5 - j;
          sy[0] = 0.5 * (0.5 - dh) * (0.5 - dh);
          sy[1] = 0.75 - dh * dh;
          sy[2] = 0.5 * (0.5 + dh) * (0.5 + dh);
          double f[6];
          for (int cc = 1; cc <= 6; ++cc)
            f[cc - 1] = +(+T(cc, i - 1, j - 1) * sx[0] + T(cc, i, j - 1) * sx[1] + T(cc, i + 1, j - 1) * sx[2]) * sy[0]
                        + (+T(cc, i - 1, j) * sx[0] + T(cc, i, j) * sx[1] + T(cc, i + 1, j) * sx[2]) * sy[1]
                        + (+T(cc, i - 1, j + 1) * sx[0] + T(cc, i, j + 1) * sx[1] + T(cc, i + 1, j + 1) * sx[2]) * sy[2];
          const double bpx = f[0], bpy = f[1], bpz = f[2], epx = f[3], epy = f[4], epz = f[5];
          if (vay) {   // particle.f90:268-299
            double uvm1 = u[2], uvm2 = u[3], uvm3 = u[4];
            double gam = std::sqrt(c * c + uvm1 * uvm1 + uvm2 * uvm2 + uvm3 * uvm3);
            const double fac1r = fac1 / gam;
            const double uvm4 = uvm1 + fac2 * epx + fac1r * (+uvm2 * bpz - uvm3 * bpy);
            const double uvm5 = uvm2 + fac2 * epy + fac1r * (+uvm3 * bpx - uvm1 * bpz);
            const double uvm6 = uvm3 + fac2 * epz + fac1r * (+uvm1 * bpy - uvm2 * bpx);
            const double taux = fac1 * bpx / c, tauy = fac1 * bpy / c, tauz = fac1 * bpz / c;
            const double tau2 = taux * taux + tauy * tauy + tauz * tauz;
            const double ua = (uvm4 * taux + uvm5 * tauy + uvm6 * tauz) / c;
            const double sigma = 1.0 + (uvm4 * uvm4 + uvm5 * uvm5 + uvm6 * uvm6) / (c * c) - tau2;
            const double gam2 = 0.5 * (sigma + std::sqrt(sigma * sigma + 4.0 * (tau2 + ua * ua)));
            gam = std::sqrt(gam2);
            const double s_ = 1.0 / (tau2 + gam2);
            g[2] = s_ * (gam2 * uvm4 + c * ua * taux + gam * (uvm5 * tauz - uvm6 * tauy));
            g[3] = s_ * (gam2 * uvm5 + c * ua * tauy + gam * (uvm6 * taux - uvm4 * tauz));
            g[4] = s_ * (gam2 * uvm6 + c * ua * tauz + gam * (uvm4 * tauy - uvm5 * taux));
            gam = 1.0 / gam;
            g[0] = u[0] + g[2] * delt * gam;
            g[1] = u[1] + g[3] * delt * gam;
            continue;
          }
          double uvm1 = u[2] + fac1 * epx;
          double uvm2 = u[3] + fac1 * epy;
          double uvm3 = u[4] + fac1 * epz;
          double gam = std::sqrt(c * c + uvm1 * uvm1 + uvm2 * uvm2 + uvm3 * uvm3);
          double igam = 1.0 / gam;
          double fac1r = fac1 * igam;
          double fac2r = fac2 / (gam + txxx * (bpx * bpx + bpy * bpy + bpz * bpz) * igam);
          double uvm4 = uvm1 + fac1r * (+uvm2 * bpz - uvm3 * bpy);
          double uvm5 = uvm2 + fac1r * (+uvm3 * bpx - uvm1 * bpz);
          double uvm6 = uvm3 + fac1r * (+uvm1 * bpy - uvm2 * bpx);
          uvm1 = uvm1 + fac2r * (+uvm5 * bpz - uvm6 * bpy);
          uvm2 = uvm2 + fac2r * (+uvm6 * bpx - uvm4 * bpz);
          uvm3 = uvm3 + fac2r * (+uvm4 * bpy - uvm5 * bpx);
          g[2] = uvm1 + fac1 * epx;
          g[3] = uvm2 + fac1 * epy;
          g[4] = uvm3 + fac1 * epz;
          if (accl) { g[0] = u[0]; g[1] = u[1]; continue; }
          gam = 1.0 / std::sqrt(1.0 + (+g[2] * g[2] + g[3] * g[3] + g[4] * g[4]) / (c * c));
          g[0] = u[0] + g[2] * delt * gam;
          g[1] = u[1] + g[3] * delt * gam;
        }
      }
  // particle.f90:173-177 -- ID carry over the whole padded array
  if (w.ndim == 6 && !accl) {
    const size_t n = up.size() / 6;
#pragma omp parallel for
    for (size_t t = 0; t < n; ++t) gp[t * 6 + 5] = up[t * 6 + 5];
  }
}

// ---------------------------------------------------------------------------
// ele_cur -- 2d/common/field.f90:189-316
// ---------------------------------------------------------------------------
void ele_cur(World2& w, Rank2& R, const std::vector<double>& up, const std::vector<double>& gp) {
  const int nxs = w.nxs, nxe = w.nxe, nys = R.nys, nye = R.nye;
  const double d_delx = w.d_delx, d_delt = w.d_delt, delx = w.delx, c = w.c;
  const double fac = 1.0 / 3.0;
  std::vector<double>& uj = R.uj;
  for (int j = nys - 2; j <= nye + 2; ++j)
    for (int i = nxs - 2; i <= nxe + 2; ++i)
      for (int cc = 1; cc <= 3; ++cc) uj[R.i3(cc, i, j)] = 0.0;
  const size_t ujn = uj.size();
  int nth = 1;
#ifdef _OPENMP
  nth = omp_get_max_threads();
#endif
  std::vector<std::vector<double>> priv(nth);
#pragma omp parallel num_threads(nth)
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    priv[tid].assign(ujn, 0.0);
    double* ujp = priv[tid].data();
#pragma omp for schedule(static)
    for (int j = nys; j <= nye; ++j)
      for (int i = nxs; i <= nxe; ++i) {
        double pjx[5][5], pjy[5][5], pjz[5][5];  // [jp+2][ip+2] == (ip,jp)
        std::memset(pjx, 0, sizeof(pjx));
        std::memset(pjy, 0, sizeof(pjy));
        std::memset(pjz, 0, sizeof(pjz));
        for (int isp = 1; isp <= w.nsp; ++isp) {
          const double qq = w.q[isp - 1];
          const int i_beg = R.cumcnt[R.ic(i, j, isp)] + 1, i_end = R.cumcnt[R.ic(i + 1, j, isp)];
          for (int ii = i_beg; ii <= i_end; ++ii) {
            const double* u = &up[R.ip(1, ii, j, isp)];
            const double* g = &gp[R.ip(1, ii, j, isp)];
            double s0[2][5], ds[2][5];
            const int cell[2] = {i, j};
            for (int a = 0; a < 2; ++a) {
              double dh = u[a] * d_delx - 0.5 - cell[a];
              s0[a][0] = 0.0;
              s0[a][1] = 0.5 * (0.5 - dh) * (0.5 - dh);
              s0[a][2] = 0.75 - dh * dh;
              s0[a][3] = 0.5 * (0.5 + dh) * (0.5 + dh);
              s0[a][4] = 0.0;
            }
            for (int a = 0; a < 2; ++a) {
              int i2 = (int)(g[a] * d_delx);
              double dh = g[a] * d_delx - 0.5 - i2;
              int inc = i2 - cell[a];
              double s1_1 = 0.5 * (0.5 - dh) * (0.5 - dh);
              double s1_2 = 0.75 - dh * dh;
              double s1_3 = 0.5 * (0.5 + dh) * (0.5 + dh);
              double smo_1 = -(inc - std::abs(inc)) * 0.5 + 0;
              double smo_2 = -std::abs(inc) + 1;
              double smo_3 = (inc + std::abs(inc)) * 0.5 + 0;
              ds[a][0] = s1_1 * smo_1;
              ds[a][1] = s1_1 * smo_2 + s1_2 * smo_1;
              ds[a][2] = s1_2 * smo_2 + s1_3 * smo_1 + s1_1 * smo_3;
              ds[a][3] = s1_3 * smo_2 + s1_2 * smo_3;
              ds[a][4] = s1_3 * smo_3;
            }
            for (int a = 0; a < 2; ++a)
              for (int m = 0; m < 5; ++m) ds[a][m] = ds[a][m] - s0[a][m];
            const double gvz = g[4] / std::sqrt(1.0 + (+g[2] * g[2] + g[3] * g[3] + g[4] * g[4]) / (c * c));
            double pjtmp[5][5];
            // Jx: pjtmp(ip+1,jp) = pjtmp(ip,jp) - q*delx*d_delt*ds(ip,1)*(s0(jp,2)+0.5*ds(jp,2))   field.f90:270-277
            std::memset(pjtmp, 0, sizeof(pjtmp));
            for (int jp = 0; jp < 5; ++jp)
              for (int ipp = 0; ipp < 4; ++ipp)
                pjtmp[jp][ipp + 1] = pjtmp[jp][ipp] - qq * delx * d_delt * ds[0][ipp] * (s0[1][jp] + 0.5 * ds[1][jp]);
            for (int jp = 0; jp < 5; ++jp)
              for (int ipp = 0; ipp < 5; ++ipp) pjx[jp][ipp] = pjx[jp][ipp] + pjtmp[jp][ipp];
            // Jy: field.f90:279-286
            std::memset(pjtmp, 0, sizeof(pjtmp));
            for (int jp = 0; jp < 4; ++jp)
              for (int ipp = 0; ipp < 5; ++ipp)
                pjtmp[jp + 1][ipp] = pjtmp[jp][ipp] - qq * delx * d_delt * ds[1][jp] * (s0[0][ipp] + 0.5 * ds[0][ipp]);
            for (int jp = 0; jp < 5; ++jp)
              for (int ipp = 0; ipp < 5; ++ipp) pjy[jp][ipp] = pjy[jp][ipp] + pjtmp[jp][ipp];
            // Jz: field.f90:288-294
            for (int jp = 0; jp < 5; ++jp)
              for (int ipp = 0; ipp < 5; ++ipp)
                pjz[jp][ipp] = pjz[jp][ipp]
                               + qq * gvz * (+s0[0][ipp] * s0[1][jp] + 0.5 * ds[0][ipp] * s0[1][jp]
                                             + 0.5 * s0[0][ipp] * ds[1][jp] + fac * ds[0][ipp] * ds[1][jp]);
          }
        }
        for (int jp = -2; jp <= 2; ++jp)
          for (int ipp = -2; ipp <= 2; ++ipp) {
            ujp[R.i3(1, i + ipp, j + jp)] += pjx[jp + 2][ipp + 2];
            ujp[R.i3(2, i + ipp, j + jp)] += pjy[jp + 2][ipp + 2];
            ujp[R.i3(3, i + ipp, j + jp)] += pjz[jp + 2][ipp + 2];
          }
      }
  }
  for (int t = 0; t < nth; ++t)
    if (!priv[t].empty())
      for (size_t n = 0; n < ujn; ++n) uj[n] += priv[t][n];
}

// ---------------------------------------------------------------------------
// boundary_*__curre -- 2d/common/boundary_periodic.f90:357-508; walls: no x treatment
// (2d/proj/reconnection/boundary_reconnection.f90:364-502, 2d/proj/shock/boundary_shock.f90:410-548)
// ---------------------------------------------------------------------------
void bc_curre(World2& w) {
  const int nxs = w.nxs, nxe = w.nxe;
  auto pack2 = [&](int j0_off, bool top) {
    return [=](Rank2& R, std::vector<double>& b) {
      const int j0 = (top ? R.nye : R.nys) + j0_off;
      for (int i = nxs - 2; i <= nxe + 2; ++i) {
        for (int cc = 1; cc <= 3; ++cc) b.push_back(R.uj[R.i3(cc, i, j0)]);
        for (int cc = 1; cc <= 3; ++cc) b.push_back(R.uj[R.i3(cc, i, j0 + 1)]);
      }
    };
  };
  auto unpack2 = [&](int j0_off, bool top, bool add) {
    return [=](Rank2& R, const std::vector<double>& b) {
      const int j0 = (top ? R.nye : R.nys) + j0_off;
      size_t t = 0;
      for (int i = nxs - 2; i <= nxe + 2; ++i) {
        for (int cc = 1; cc <= 3; ++cc) { double& v = R.uj[R.i3(cc, i, j0)]; v = add ? v + b[t] : b[t]; ++t; }
        for (int cc = 1; cc <= 3; ++cc) { double& v = R.uj[R.i3(cc, i, j0 + 1)]; v = add ? v + b[t] : b[t]; ++t; }
      }
    };
  };
  sendrecv<double>(w, TO_DOWN, pack2(-2, false), unpack2(-1, true, true));   // nys-2,nys-1 -> += nye-1,nye
  sendrecv<double>(w, TO_UP, pack2(+1, true), unpack2(0, false, true));      // nye+1,nye+2 -> += nys,nys+1
  sendrecv<double>(w, TO_DOWN, pack2(0, false), unpack2(+1, true, false));   // nys,nys+1   -> nye+1,nye+2
  sendrecv<double>(w, TO_UP, pack2(-1, true), unpack2(-2, false, false));    // nye-1,nye   -> nys-2,nys-1
  if (w.bc != 0) return;
  for (Rank2& R : w.ranks) {
    for (int j = R.nys - 2; j <= R.nye + 2; ++j)
      for (int cc = 1; cc <= 3; ++cc) {
        R.uj[R.i3(cc, nxe - 1, j)] += R.uj[R.i3(cc, nxs - 2, j)];
        R.uj[R.i3(cc, nxe, j)] += R.uj[R.i3(cc, nxs - 1, j)];
        R.uj[R.i3(cc, nxs, j)] += R.uj[R.i3(cc, nxe + 1, j)];
        R.uj[R.i3(cc, nxs + 1, j)] += R.uj[R.i3(cc, nxe + 2, j)];
      }
    for (int j = R.nys - 2; j <= R.nye + 2; ++j)
      for (int cc = 1; cc <= 3; ++cc) {
        R.uj[R.i3(cc, nxs - 2, j)] = R.uj[R.i3(cc, nxe - 1, j)];
        R.uj[R.i3(cc, nxs - 1, j)] = R.uj[R.i3(cc, nxe, j)];
        R.uj[R.i3(cc, nxe + 1, j)] = R.uj[R.i3(cc, nxs, j)];
        R.uj[R.i3(cc, nxe + 2, j)] = R.uj[R.i3(cc, nxs + 1, j)];
      }
  }
}

// ---------------------------------------------------------------------------
// boundary_*__dfield -- 2d/common/boundary_periodic.f90:251-354; x rule of the walls:
// 2d/proj/reconnection/boundary_reconnection.f90:350-359, 2d/proj/shock/boundary_shock.f90:396-405
// ---------------------------------------------------------------------------
void bc_dfield(World2& w) {
  const int nxs = w.nxs, nxe = w.nxe;
  auto pack2 = [&](int j0_off, bool top) {
    return [=](Rank2& R, std::vector<double>& b) {
      const int j0 = (top ? R.nye : R.nys) + j0_off;
      for (int i = nxs; i <= nxe; ++i) {
        for (int cc = 1; cc <= 6; ++cc) b.push_back(R.df[R.i6(cc, i, j0)]);
        for (int cc = 1; cc <= 6; ++cc) b.push_back(R.df[R.i6(cc, i, j0 + 1)]);
      }
    };
  };
  auto unpack2 = [&](int j0_off, bool top) {
    return [=](Rank2& R, const std::vector<double>& b) {
      const int j0 = (top ? R.nye : R.nys) + j0_off;
      size_t t = 0;
      for (int i = nxs; i <= nxe; ++i) {
        for (int cc = 1; cc <= 6; ++cc) R.df[R.i6(cc, i, j0)] = b[t++];
        for (int cc = 1; cc <= 6; ++cc) R.df[R.i6(cc, i, j0 + 1)] = b[t++];
      }
    };
  };
  sendrecv<double>(w, TO_DOWN, pack2(0, false), unpack2(+1, true));   // nys,nys+1 -> nye+1,nye+2
  sendrecv<double>(w, TO_UP, pack2(-1, true), unpack2(-2, false));    // nye-1,nye -> nys-2,nys-1
  for (Rank2& R : w.ranks)
    for (int j = R.nys - 2; j <= R.nye + 2; ++j) {
      auto D = [&](int cc, int i) -> double& { return R.df[R.i6(cc, i, j)]; };
      if (w.bc == 0) {
        for (int cc = 1; cc <= 6; ++cc) {
          D(cc, nxs - 2) = D(cc, nxe - 1);
          D(cc, nxs - 1) = D(cc, nxe);
          D(cc, nxe + 1) = D(cc, nxs);
          D(cc, nxe + 2) = D(cc, nxs + 1);
        }
      } else {
        D(1, nxs - 1) = -D(1, nxs);
        for (int cc = 2; cc <= 4; ++cc) D(cc, nxs - 1) = D(cc, nxs + 1);
        for (int cc = 5; cc <= 6; ++cc) D(cc, nxs - 1) = -D(cc, nxs);
        if (w.bc == 1) {
          D(1, nxe) = -D(1, nxe - 1);
          for (int cc = 2; cc <= 4; ++cc) D(cc, nxe + 1) = D(cc, nxe - 1);
          for (int cc = 5; cc <= 6; ++cc) D(cc, nxe) = -D(cc, nxe - 1);
        } else {
          for (int cc = 1; cc <= 6; ++cc) D(cc, nxe + 1) = 0.0;
        }
      }
    }
}

// ---------------------------------------------------------------------------
// boundary_*__phi -- 2d/common/boundary_periodic.f90:511-568; walls:
// 2d/proj/reconnection/boundary_reconnection.f90:557-577, 2d/proj/shock/boundary_shock.f90:603-623
// ---------------------------------------------------------------------------
void bc_phi(World2& w, std::vector<Cg2>& cg, int sel, int l) {
  const int nxs = w.nxs, nxe = w.nxe;
  auto A = [&](Rank2& R) -> std::vector<double>& { return sel == 0 ? cg[R.rank].phi : cg[R.rank].p; };
  auto I = [&](Rank2& R, int i, int j) { return cg[R.rank].i1(i, j); };
  sendrecv<double>(w, TO_DOWN,
      [&](Rank2& R, std::vector<double>& b) { for (int i = nxs; i <= nxe; ++i) b.push_back(A(R)[I(R, i, R.nys)]); },
      [&](Rank2& R, const std::vector<double>& b) { size_t t = 0; for (int i = nxs; i <= nxe; ++i) A(R)[I(R, i, R.nye + 1)] = b[t++]; });
  sendrecv<double>(w, TO_UP,
      [&](Rank2& R, std::vector<double>& b) { for (int i = nxs; i <= nxe; ++i) b.push_back(A(R)[I(R, i, R.nye)]); },
      [&](Rank2& R, const std::vector<double>& b) { size_t t = 0; for (int i = nxs; i <= nxe; ++i) A(R)[I(R, i, R.nys - 1)] = b[t++]; });
  for (Rank2& R : w.ranks)
    for (int j = R.nys - 1; j <= R.nye + 1; ++j) {
      std::vector<double>& a = A(R);
      if (w.bc == 0) {
        a[I(R, nxs - 1, j)] = a[I(R, nxe, j)];
        a[I(R, nxe + 1, j)] = a[I(R, nxs, j)];
      } else if (l == 1) {
        a[I(R, nxs - 1, j)] = -a[I(R, nxs, j)];
        a[I(R, nxe + 1, j)] = w.bc == 1 ? -a[I(R, nxe - 2, j)] : 0.0;
      } else {
        a[I(R, nxs - 1, j)] = a[I(R, nxs + 1, j)];
        a[I(R, nxe + 1, j)] = w.bc == 1 ? a[I(R, nxe - 1, j)] : 0.0;
      }
    }
}

// ---------------------------------------------------------------------------
// cgm -- 2d/common/field.f90:319-461 (control flow: SURVEY.md 3.3)
// ---------------------------------------------------------------------------
void cgm(World2& w) {
  const int nxs = w.nxs, nxe = w.nxe;
  const int ite_max = 100;
  const double err = 1e-6;
  const int NR = (int)w.ranks.size();
  std::vector<Cg2> cg(NR);
  for (int r = 0; r < NR; ++r) {
    Rank2& R = w.ranks[r];
    Cg2& c = cg[r];
    c.nxs = nxs; c.nxe = nxe; c.nys = R.nys; c.nye = R.nye;
    const size_t n1 = (size_t)(nxe - nxs + 3) * (R.nyl + 2), n0 = (size_t)(nxe - nxs + 1) * R.nyl;
    c.phi.assign(n1, 0.0); c.p.assign(n1, 0.0);
    c.r.assign(n0, 0.0); c.b.assign(n0, 0.0); c.ap.assign(n0, 0.0);
  }
  const double f4 = w.f4, f5 = w.f5;
  for (int l = 1; l <= 3; ++l) {
    int ite = 0;
    double sum_g = 0.0;
    for (int r = 0; r < NR; ++r) {
      Rank2& R = w.ranks[r]; Cg2& c = cg[r];
      double sum = 0.0;
      for (int j = R.nys; j <= R.nye; ++j)
        for (int i = nxs; i <= nxe; ++i) {
          c.phi[c.i1(i, j)] = R.df[R.i6(l, i, j)];
          const double bb = f5 * R.gkl[R.ig(l, i, j)];
          c.b[c.i0(i, j)] = bb;
          sum = sum + bb * bb;
        }
      sum_g += sum;
    }
    const double eps = std::sqrt(sum_g) * err;
    bc_phi(w, cg, 0, l);
    double sumr_g = 0.0;
    for (int r = 0; r < NR; ++r) {
      Rank2& R = w.ranks[r]; Cg2& c = cg[r];
      double sumr = 0.0;
      for (int j = R.nys; j <= R.nye; ++j)
        for (int i = nxs; i <= nxe; ++i) {
          const double rr = c.b[c.i0(i, j)] + c.phi[c.i1(i, j - 1)] + c.phi[c.i1(i - 1, j)] - f4 * c.phi[c.i1(i, j)]
                            + c.phi[c.i1(i + 1, j)] + c.phi[c.i1(i, j + 1)];
          c.r[c.i0(i, j)] = rr;
          c.p[c.i1(i, j)] = rr;
          sumr = sumr + rr * rr;
        }
      sumr_g += sumr;
    }
    if (std::sqrt(sumr_g) > eps) {
      while (sum_g > eps) {
        ite = ite + 1;
        bc_phi(w, cg, 1, l);
        double s_r = 0.0, s_2 = 0.0;
        for (int r = 0; r < NR; ++r) {
          Rank2& R = w.ranks[r]; Cg2& c = cg[r];
          double sumr = 0.0, sum2 = 0.0;
          for (int j = R.nys; j <= R.nye; ++j)
            for (int i = nxs; i <= nxe; ++i) {
              const double a = -c.p[c.i1(i, j - 1)] - c.p[c.i1(i - 1, j)] + f4 * c.p[c.i1(i, j)] - c.p[c.i1(i + 1, j)]
                               - c.p[c.i1(i, j + 1)];
              c.ap[c.i0(i, j)] = a;
              sumr = sumr + c.r[c.i0(i, j)] * c.r[c.i0(i, j)];
              sum2 = sum2 + c.p[c.i1(i, j)] * a;
            }
          s_r += sumr; s_2 += sum2;
        }
        sumr_g = s_r;
        const double sum2_g = s_2;
        const double av = sumr_g / sum2_g;
        for (int r = 0; r < NR; ++r) {
          Rank2& R = w.ranks[r]; Cg2& c = cg[r];
          for (int j = R.nys; j <= R.nye; ++j)
            for (int i = nxs; i <= nxe; ++i) {
              c.phi[c.i1(i, j)] = c.phi[c.i1(i, j)] + av * c.p[c.i1(i, j)];
              c.r[c.i0(i, j)] = c.r[c.i0(i, j)] - av * c.ap[c.i0(i, j)];
            }
        }
        sum_g = std::sqrt(sumr_g);
        if (ite >= ite_max) {
          std::fprintf(stderr, "********** stop at cgm after ite_max **********\n");
          w.err = 1;
          return;
        }
        double sum1_g = 0.0;
        for (int r = 0; r < NR; ++r) {
          Rank2& R = w.ranks[r]; Cg2& c = cg[r];
          double sum1 = 0.0;
          for (int j = R.nys; j <= R.nye; ++j)
            for (int i = nxs; i <= nxe; ++i) sum1 = sum1 + c.r[c.i0(i, j)] * c.r[c.i0(i, j)];
          sum1_g += sum1;
        }
        const double bv = sum1_g / sumr_g;
        for (int r = 0; r < NR; ++r) {
          Rank2& R = w.ranks[r]; Cg2& c = cg[r];
          for (int j = R.nys; j <= R.nye; ++j)
            for (int i = nxs; i <= nxe; ++i) c.p[c.i1(i, j)] = c.r[c.i0(i, j)] + bv * c.p[c.i1(i, j)];
        }
      }
    }
    for (int r = 0; r < NR; ++r) {
      Rank2& R = w.ranks[r]; Cg2& c = cg[r];
      for (int j = R.nys; j <= R.nye; ++j)
        for (int i = nxs; i <= nxe; ++i) R.df[R.i6(l, i, j)] = c.phi[c.i1(i, j)];
    }
    w.cg_ite[l - 1] = ite;
  }
}

// ---------------------------------------------------------------------------
// field__fdtd_i -- 2d/common/field.f90:66-186; stages as in oracle3d.cpp
// ---------------------------------------------------------------------------
void stage_gkl(World2& w, Rank2& R) {
  const double f1 = w.f1, f2 = w.f2, f3 = w.f3;
  const std::vector<double>&uf = R.uf, &uj = R.uj;
#pragma omp parallel for
  for (int j = R.nys; j <= R.nye; ++j)
    for (int i = w.nxs; i <= w.nxe; ++i) {
      R.gkl[R.ig(1, i, j)] = +f2 * (+uf[R.i6(1, i, j - 1)] + uf[R.i6(1, i - 1, j)] - 4.0 * uf[R.i6(1, i, j)]
                                     + uf[R.i6(1, i + 1, j)] + uf[R.i6(1, i, j + 1)]
                                     + f3 * (-uj[R.i3(3, i, j - 1)] + uj[R.i3(3, i, j)]))
                             - f1 * (-uf[R.i6(6, i, j - 1)] + uf[R.i6(6, i, j)]);
      R.gkl[R.ig(2, i, j)] = +f2 * (+uf[R.i6(2, i, j - 1)] + uf[R.i6(2, i - 1, j)] - 4.0 * uf[R.i6(2, i, j)]
                                     + uf[R.i6(2, i + 1, j)] + uf[R.i6(2, i, j + 1)]
                                     - f3 * (-uj[R.i3(3, i - 1, j)] + uj[R.i3(3, i, j)]))
                             + f1 * (-uf[R.i6(6, i - 1, j)] + uf[R.i6(6, i, j)]);
      R.gkl[R.ig(3, i, j)] = +f2 * (+uf[R.i6(3, i, j - 1)] + uf[R.i6(3, i - 1, j)] - 4.0 * uf[R.i6(3, i, j)]
                                     + uf[R.i6(3, i + 1, j)] + uf[R.i6(3, i, j + 1)]
                                     + f3 * (-uj[R.i3(2, i - 1, j)] + uj[R.i3(2, i, j)] + uj[R.i3(1, i, j - 1)]
                                             - uj[R.i3(1, i, j)]))
                             - f1 * (-uf[R.i6(5, i - 1, j)] + uf[R.i6(5, i, j)] + uf[R.i6(4, i, j - 1)] - uf[R.i6(4, i, j)]);
    }
}

void stage_de(World2& w, Rank2& R) {
  const double f1 = w.f1, gfac = w.gfac, delt = w.delt;
  const std::vector<double>&uf = R.uf, &uj = R.uj;
  std::vector<double>& df = R.df;
#pragma omp parallel for
  for (int j = R.nys; j <= R.nye; ++j)
    for (int i = w.nxs; i <= w.nxe; ++i) {
      df[R.i6(4, i, j)] = +f1 * (+gfac * (-df[R.i6(3, i, j)] + df[R.i6(3, i, j + 1)]) + (-uf[R.i6(3, i, j)] + uf[R.i6(3, i, j + 1)]))
                          - 4.0 * kPi * delt * uj[R.i3(1, i, j)];
      df[R.i6(5, i, j)] = -f1 * (+gfac * (-df[R.i6(3, i, j)] + df[R.i6(3, i + 1, j)]) + (-uf[R.i6(3, i, j)] + uf[R.i6(3, i + 1, j)]))
                          - 4.0 * kPi * delt * uj[R.i3(2, i, j)];
      df[R.i6(6, i, j)] = +f1 * (+gfac * (-df[R.i6(2, i, j)] + df[R.i6(2, i + 1, j)] + df[R.i6(1, i, j)] - df[R.i6(1, i, j + 1)])
                                 + (-uf[R.i6(2, i, j)] + uf[R.i6(2, i + 1, j)] + uf[R.i6(1, i, j)] - uf[R.i6(1, i, j + 1)]))
                          - 4.0 * kPi * delt * uj[R.i3(3, i, j)];
    }
}

void stage_update(World2& w, Rank2& R) {
  for (int j = R.nys - 2; j <= R.nye + 2; ++j)
    for (int i = w.nxs - 2; i <= w.nxe + 2; ++i)
      for (int cc = 1; cc <= 6; ++cc) R.uf[R.i6(cc, i, j)] = R.uf[R.i6(cc, i, j)] + R.df[R.i6(cc, i, j)];
}

void field_fdtd_i(World2& w, int stage) {
  if (stage == 0 || stage == 1) for (Rank2& R : w.ranks) ele_cur(w, R, R.up, R.gp);
  if (stage == 0 || stage == 2) bc_curre(w);
  if (stage == 0 || stage == 3) for (Rank2& R : w.ranks) stage_gkl(w, R);
  if (stage == 0 || stage == 4) { cgm(w); if (w.err) return; }
  if (stage == 0 || stage == 5) bc_dfield(w);
  if (stage == 0 || stage == 6) for (Rank2& R : w.ranks) stage_de(w, R);
  if (stage == 0 || stage == 7) bc_dfield(w);
  if (stage == 0 || stage == 8) for (Rank2& R : w.ranks) stage_update(w, R);
}

// ---------------------------------------------------------------------------
// boundary_periodic__particle_x -- 2d/common/boundary_periodic.f90:61-96 (whole body under ieee_down)
// ---------------------------------------------------------------------------
void bc_particle_x_periodic(World2& w, Rank2& R, std::vector<double>& up) {
  const double len = mul_rd((double)(w.nxge - w.nxgs + 1), w.delx);
  for (int isp = 1; isp <= w.nsp; ++isp)
    for (int j = R.nys; j <= R.nye; ++j) {
      const int n = R.np2[R.in2(j, isp)];
      for (int ii = 1; ii <= n; ++ii) {
        double& x = up[R.ip(1, ii, j, isp)];
        const int ipos = (int)div_rd(x, w.delx);
        if (ipos < w.nxgs) x = add_rd(x, len);
        else if (ipos >= w.nxge + 1) x = add_rd(x, -len);
      }
    }
}

// boundary_reconnection__particle_x / boundary_shock__particle_x (reflecting walls, round-to-nearest)
// 2d/proj/reconnection/boundary_reconnection.f90:61-99, 2d/proj/shock/boundary_shock.f90:62-100
void bc_particle_x_reflect(World2& w, Rank2& R, std::vector<double>& up) {
  const int nxs = w.nxs, nxe = w.nxe;
  for (int isp = 1; isp <= w.nsp; ++isp)
    for (int j = R.nys; j <= R.nye; ++j) {
      const int n = R.np2[R.in2(j, isp)];
      for (int ii = 1; ii <= n; ++ii) {
        double* u = &up[R.ip(1, ii, j, isp)];
        const int ipos = (int)(u[0] / w.delx);
        if (ipos < nxs + 1) {
          u[0] = 2.0 * (nxs + 1) * w.delx - u[0];
          u[2] = -u[2]; u[3] = -u[3]; u[4] = -u[4];
        } else if (ipos >= nxe - 1) {
          u[0] = 2.0 * (nxe - 1) * w.delx - u[0];
          u[2] = -u[2]; u[3] = -u[3]; u[4] = -u[4];
        }
      }
    }
}

// boundary_shock__injection -- 2d/proj/shock/boundary_shock.f90:255-297
void bc_injection(World2& w, Rank2& R, std::vector<double>& up, double u0) {
  const int nxs = w.nxs, nxe = w.nxe;
  const double xend = nxe * w.delx + u0 / std::sqrt(1 + (u0 * u0) / (w.c * w.c)) * w.delt;
  for (int isp = 1; isp <= w.nsp; ++isp)
    for (int j = R.nys; j <= R.nye; ++j) {
      const int n = R.np2[R.in2(j, isp)];
      for (int ii = 1; ii <= n; ++ii) {
        double* u = &up[R.ip(1, ii, j, isp)];
        const int ipos = (int)(u[0] / w.delx);
        if (ipos < nxs + 1) {
          u[0] = +2.0 * (nxs + 1) * w.delx - u[0];
          u[2] = -u[2]; u[3] = -u[3]; u[4] = -u[4];
        } else if (u[0] > xend) {
          u[0] = +2.0 * xend - u[0];
          u[2] = +2.0 * u0 - u[2];
          u[3] = -u[3]; u[4] = -u[4];
        }
      }
    }
}

// ---------------------------------------------------------------------------
// boundary_*__particle_y -- 2d/common/boundary_periodic.f90:99-248 (identical in the reconnection and
// shock modules; all three run under ieee_down).  Serial per rank: j / ii order is one admissible
// outcome of the reference's lock-ordered appends.
// ---------------------------------------------------------------------------
void bc_particle_y(World2& w, int which /*0: gp, 1: up*/) {
  const int NR = (int)w.ranks.size();
  const int ndim = w.ndim;
  struct Mig { std::vector<std::vector<double>> bff; std::vector<int> cnt, cnt2; std::vector<std::vector<int>> flag; };
  std::vector<Mig> M(NR);
  auto P = [&](Rank2& R) -> std::vector<double>& { return which == 0 ? R.gp : R.up; };
  const double len = mul_rd((double)(w.nyge - w.nygs + 1), w.delx);
  for (int isp = 1; isp <= w.nsp; ++isp) {
    for (int r = 0; r < NR; ++r) {
      Rank2& R = w.ranks[r]; Mig& m = M[r];
      std::vector<double>& up = P(R);
      m.bff.assign((size_t)R.nyl + 2, {});
      m.cnt.assign((size_t)R.nyl + 2, 0);
      m.cnt2.assign((size_t)R.nyl, 0);
      m.flag.assign((size_t)R.nyl, {});
      for (int j = R.nys; j <= R.nye; ++j) {
        const int n = R.np2[R.in2(j, isp)];
        for (int ii = 1; ii <= n; ++ii) {
          double* u = &up[R.ip(1, ii, j, isp)];
          const int jpos = (int)div_rd(u[1], w.delx);
          if (jpos != j) {
            if (jpos <= w.nygs - 1) u[1] = add_rd(u[1], len);
            else if (jpos >= w.nyge + 1) u[1] = add_rd(u[1], -len);
            if (jpos < R.nys - 1 || jpos > R.nye + 1) {
              std::fprintf(stderr, "oracle2d: particle moved more than one row (jpos=%d j=%d)\n", jpos, j);
              w.err = 3;
              return;
            }
            std::vector<double>& b = m.bff[jpos - (R.nys - 1)];
            b.insert(b.end(), u, u + ndim);
            m.cnt[jpos - (R.nys - 1)] += 1;
            m.cnt2[j - R.nys] += 1;
            m.flag[j - R.nys].push_back(ii);
          }
        }
      }
    }
    // transfer to rank-1 (:174-180), then to rank+1 (:183-189)
    {
      std::vector<std::vector<double>> snd(NR);
      for (int r = 0; r < NR; ++r) snd[r] = M[r].bff[0];
      for (int r = 0; r < NR; ++r) {
        Rank2& R = w.ranks[r]; Mig& m = M[r];
        const std::vector<double>& in = snd[R.nup];
        std::vector<double>& b = m.bff[R.nye - (R.nys - 1)];
        b.insert(b.end(), in.begin(), in.end());
        m.cnt[R.nye - (R.nys - 1)] += (int)(in.size() / ndim);
      }
      for (int r = 0; r < NR; ++r) snd[r] = M[r].bff[w.ranks[r].nyl + 1];
      for (int r = 0; r < NR; ++r) {
        Rank2& R = w.ranks[r]; Mig& m = M[r];
        const std::vector<double>& in = snd[R.ndown];
        std::vector<double>& b = m.bff[1];
        b.insert(b.end(), in.begin(), in.end());
        m.cnt[1] += (int)(in.size() / ndim);
      }
    }
    // hole filling / append (:192-236)
    for (int r = 0; r < NR; ++r) {
      Rank2& R = w.ranks[r]; Mig& m = M[r];
      std::vector<double>& up = P(R);
      for (int j = R.nys; j <= R.nye; ++j) {
        int& np2 = R.np2[R.in2(j, isp)];
        int& cnt = m.cnt[j - (R.nys - 1)];
        const std::vector<double>& bff = m.bff[j - (R.nys - 1)];
        const std::vector<int>& flag = m.flag[j - R.nys];
        const int c2 = m.cnt2[j - R.nys];
        int iii = 0;
        int cnt_tmp = c2;
        bool done = false;
        for (int ii = 1; ii <= c2 && !done; ++ii) {
          if (cnt == 0) {
            if (np2 < flag[ii - 1]) break;
            while (np2 == flag[cnt_tmp - 1]) {
              np2 = np2 - 1;
              if (np2 < flag[ii - 1]) { done = true; break; }
              cnt_tmp = cnt_tmp - 1;
            }
            if (done) break;
            for (int d = 1; d <= ndim; ++d) up[R.ip(d, flag[ii - 1], j, isp)] = up[R.ip(d, np2, j, isp)];
            np2 = np2 - 1;
          } else {
            for (int d = 1; d <= ndim; ++d) up[R.ip(d, flag[ii - 1], j, isp)] = bff[(size_t)ndim * iii + (d - 1)];
            iii = iii + 1;
            cnt = cnt - 1;
          }
        }
        if (cnt > 0) {
          if (np2 + cnt > w.np) {
            std::fprintf(stderr, "memory over (np2 > np) %d %d %d %d\n", w.np, np2 + cnt, j, isp);
            w.err = 2;
            return;
          }
          for (int ii = 1; ii <= cnt; ++ii)
            for (int d = 1; d <= ndim; ++d)
              up[R.ip(d, np2 + ii, j, isp)] = bff[(size_t)ndim * iii + (d - 1) + (size_t)ndim * (ii - 1)];
        }
        np2 = np2 + cnt;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// sort__bucket -- 2d/common/sort.f90:36-82
// ---------------------------------------------------------------------------
void sort_bucket(World2& w, Rank2& R, std::vector<double>& dst, const std::vector<double>& src) {
  const int nxs = w.nxs, nxe = w.nxe, ndim = w.ndim;
  for (int isp = 1; isp <= w.nsp; ++isp) {
#pragma omp parallel for
    for (int j = R.nys; j <= R.nye; ++j) {
      std::vector<int> cnt(nxe - nxs + 1, 0), sum_cnt(nxe - nxs + 2, 0);
      const int n = R.np2[R.in2(j, isp)];
      bool bad = false;
      for (int ii = 1; ii <= n; ++ii) {
        const int i = (int)(src[R.ip(1, ii, j, isp)]);
        if (i < nxs || i > nxe) { bad = true; break; }  // the reference would index out of bounds here
        cnt[i - nxs] += 1;
      }
      if (bad) { w.err = 4; continue; }
      sum_cnt[0] = 0;
      R.cumcnt[R.ic(nxs, j, isp)] = 0;
      for (int i = nxs + 1; i <= nxe + 1; ++i) {
        sum_cnt[i - nxs] = sum_cnt[i - 1 - nxs] + cnt[i - 1 - nxs];
        R.cumcnt[R.ic(i, j, isp)] = sum_cnt[i - nxs];
      }
      for (int ii = 1; ii <= n; ++ii) {
        const int i = (int)(src[R.ip(1, ii, j, isp)]);
        for (int d = 1; d <= ndim; ++d) dst[R.ip(d, sum_cnt[i - nxs] + 1, j, isp)] = src[R.ip(d, ii, j, isp)];
        sum_cnt[i - nxs] += 1;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// mom_calc__nvt -- 2d/common/mom_calc.f90:166-252 (ih = int(x*d_delx - 0.5), dx = x*d_delx - 0.5 - ih)
// ---------------------------------------------------------------------------
void mom_nvt(World2& w, Rank2& R, const std::vector<double>& up) {
  std::fill(R.mom.begin(), R.mom.end(), 0.0);
  for (int isp = 1; isp <= w.nsp; ++isp)
    for (int j = R.nys; j <= R.nye; ++j) {
      const int n = R.np2[R.in2(j, isp)];
      for (int ii = 1; ii <= n; ++ii) {
        const double* u = &up[R.ip(1, ii, j, isp)];
        const int ih = (int)(u[0] * w.d_delx - 0.5);
        const int jh = (int)(u[1] * w.d_delx - 0.5);
        const double dx = u[0] * w.d_delx - 0.5 - ih, dxm = 1.0 - dx;
        const double dy = u[1] * w.d_delx - 0.5 - jh, dym = 1.0 - dy;
        const double gam = 1.0 / std::sqrt(1.0 + (+u[2] * u[2] + u[3] * u[3] + u[4] * u[4]) / (w.c * w.c));
        const double wx[2] = {dxm, dx}, wy[2] = {dym, dy};
        const double val[7] = {1.0, u[2] * gam, u[3] * gam, u[4] * gam, u[2] * u[2] * gam, u[3] * u[3] * gam, u[4] * u[4] * gam};
        for (int l = 1; l <= 7; ++l)
          for (int b = 0; b < 2; ++b)
            for (int a = 0; a < 2; ++a) {
              double& m = R.mom[R.im(l, ih + a, jh + b, isp)];
              m = l == 1 ? m + wx[a] * wy[b] : m + val[l - 1] * wx[a] * wy[b];
            }
      }
    }
}

// boundary_periodic__mom -- 2d/common/boundary_periodic.f90:571-636; walls fold x onto the same side
// (2d/proj/reconnection/boundary_reconnection.f90:582-647, 2d/proj/shock/boundary_shock.f90:628-693)
void bc_mom(World2& w) {
  for (Rank2& R : w.ranks)
    for (int isp = 1; isp <= w.nsp; ++isp)
      for (int j = R.nys - 1; j <= R.nye + 1; ++j)
        for (int l = 1; l <= 7; ++l) {
          if (w.bc == 0) {
            R.mom[R.im(l, w.nxgs, j, isp)] += R.mom[R.im(l, w.nxge + 1, j, isp)];
            R.mom[R.im(l, w.nxge, j, isp)] += R.mom[R.im(l, w.nxgs - 1, j, isp)];
          } else {
            R.mom[R.im(l, w.nxgs, j, isp)] += R.mom[R.im(l, w.nxgs - 1, j, isp)];
            R.mom[R.im(l, w.nxge, j, isp)] += R.mom[R.im(l, w.nxge + 1, j, isp)];
          }
        }
  for (int isp = 1; isp <= w.nsp; ++isp) {
    sendrecv<double>(w, TO_DOWN,
        [&](Rank2& R, std::vector<double>& b) {
          for (int i = w.nxgs - 1; i <= w.nxge + 1; ++i) for (int l = 1; l <= 7; ++l) b.push_back(R.mom[R.im(l, i, R.nys - 1, isp)]);
        },
        [&](Rank2& R, const std::vector<double>& b) {
          size_t t = 0;
          for (int i = w.nxgs - 1; i <= w.nxge + 1; ++i) for (int l = 1; l <= 7; ++l) R.mom[R.im(l, i, R.nye, isp)] += b[t++];
        });
    sendrecv<double>(w, TO_UP,
        [&](Rank2& R, std::vector<double>& b) {
          for (int i = w.nxgs - 1; i <= w.nxge + 1; ++i) for (int l = 1; l <= 7; ++l) b.push_back(R.mom[R.im(l, i, R.nye + 1, isp)]);
        },
        [&](Rank2& R, const std::vector<double>& b) {
          size_t t = 0;
          for (int i = w.nxgs - 1; i <= w.nxge + 1; ++i) for (int l = 1; l <= 7; ++l) R.mom[R.im(l, i, R.nys, isp)] += b[t++];
        });
  }
}

// the moment block of the drivers (2d/proj/weibel/app.f90:116-119): accl (gp <- up), nvt (mom <- gp), bc__mom
void mom_calc(World2& w) {
  for (Rank2& R : w.ranks) particle_solv(w, R, R.gp, R.up, true);
  for (Rank2& R : w.ranks) mom_nvt(w, R, R.gp);
  bc_mom(w);
}


// the upstream field columns both inject() and relocate() reset (2d/proj/shock/app.f90:680-689, 840-849)
void shock_boundary_field(World2& w, Rank2& R, const orc::ShockPrm& sp, int nxe) {
  const double by = sp.b0 * std::sin(sp.theta_bn) * std::cos(sp.phi_bn), bz = sp.b0 * std::sin(sp.theta_bn) * std::sin(sp.phi_bn);
  for (int j = R.nys - 2; j <= R.nye + 2; ++j) {
    R.uf[R.i6(2, nxe - 1, j)] = by;
    R.uf[R.i6(3, nxe - 1, j)] = bz;
    R.uf[R.i6(5, nxe - 1, j)] = +sp.v0 * R.uf[R.i6(3, nxe - 1, j)] / w.c;
    R.uf[R.i6(6, nxe - 1, j)] = -sp.v0 * R.uf[R.i6(2, nxe - 1, j)] / w.c;
    R.uf[R.i6(2, nxe, j)] = by;
    R.uf[R.i6(3, nxe, j)] = bz;
  }
}

// ---------------------------------------------------------------------------
// inject -- 2d/proj/shock/app.f90:697-852, with the integer bookkeeping of :711-774 (how many particles each row
// receives: nlinj_grid) supplied by the caller per GLOBAL row, so that the result does not depend on the rank count.
// nptotal (:769-770) = global population per species before the call, ncinj_grid (:778-781) = exclusive prefix of the
// row counts in rank order (= global row order for y slabs).
// ---------------------------------------------------------------------------
void shock_inject(World2& w, const orc::ShockPrm& sp, const int* nlinj_rows, uint32_t epoch) {
  const int nxe = w.nxe;
  const double delx = w.delx, delt = w.delt, c = w.c;
  long long nptotal[2] = {0, 0};
  for (Rank2& R : w.ranks)
    for (int isp = 1; isp <= 2; ++isp)
      for (int j = R.nys; j <= R.nye; ++j) nptotal[isp - 1] += R.np2[R.in2(j, isp)];
  const double x0 = std::fabs(sp.v0) * delt;
  for (Rank2& R : w.ranks) {
    for (int j = R.nys; j <= R.nye; ++j) {
      const uint32_t row = (uint32_t)(j - w.nygs);
      const int n = nlinj_rows[row];
      long long ncinj = 0;
      for (uint32_t rr = 0; rr < row; ++rr) ncinj += nlinj_rows[rr];
      for (int isp = 1; isp <= 2; ++isp) {
        const int base = R.np2[R.in2(j, isp)];
        if (base + n > w.np) { w.err = 2; return; }
        for (int ii = 1; ii <= n; ++ii) {
          double* u = &R.up[R.ip(1, base + ii, j, isp)];
          double ur0, ur1;
          orc::Philox::uniform2(sp.seed, row, (uint32_t)ii, 0u, ur0, ur1, epoch);
          u[0] = nxe * delx + (ii - 0.5) / n * x0;      // :790
          u[1] = (j + ur0) * delx;                      // :791
          double v[3];
          orc::shock_velocity(sp, row, (uint32_t)ii, isp, 0u, epoch, c, v);
          u[2] = v[0]; u[3] = v[1]; u[4] = v[2];
          u[0] = u[0] + (sp.v0 + u[2]) * delt;          // :817-818 injection (non-relativistic approximation)
          const double v1 = orc::vprofile(sp, u[0], w.nxgs, delx);
          const double gam1 = 1 / std::sqrt(1 - (v1 / c) * (v1 / c));
          const double gamp = std::sqrt(1 + (u[2] * u[2] + u[3] * u[3] + u[4] * u[4]) / (c * c));
          u[2] = gam1 * (u[2] + v1 * gamp);             // :836
          const int64_t pid = (int64_t)ii + ncinj + nptotal[isp - 1];   // :839
          const int64_t neg = -pid;
          std::memcpy(&u[5], &neg, 8);
        }
      }
      for (int isp = 1; isp <= 2; ++isp) {             // :846-851
        R.np2[R.in2(j, isp)] += n;
        R.cumcnt[R.ic(nxe, j, isp)] += n;
      }
    }
    shock_boundary_field(w, R, sp, nxe);
  }
}

// relocate -- 2d/proj/shock/app.f90:615-692: the box grows by one cell filled with n0 particles per row and species
void shock_relocate(World2& w, const orc::ShockPrm& sp, uint32_t epoch) {
  if (w.nxe == w.nxge) return;
  w.nxe = w.nxe + 1;
  const int nxe = w.nxe, n0 = sp.n0;
  const double delx = w.delx, c = w.c;
  long long nptotal[2] = {0, 0};
  for (Rank2& R : w.ranks)
    for (int isp = 1; isp <= 2; ++isp)
      for (int j = R.nys; j <= R.nye; ++j) nptotal[isp - 1] += R.np2[R.in2(j, isp)];
  for (Rank2& R : w.ranks) {
    for (int j = R.nys; j <= R.nye; ++j) {
      const uint32_t row = (uint32_t)(j - w.nygs);
      for (int isp = 1; isp <= 2; ++isp) {
        const int base = R.np2[R.in2(j, isp)];
        if (base + n0 > w.np) { w.err = 2; return; }
        for (int ii = 1; ii <= n0; ++ii) {
          double* u = &R.up[R.ip(1, base + ii, j, isp)];
          double ur0, ur1;
          orc::Philox::uniform2(sp.seed, row, (uint32_t)ii, 16u, ur0, ur1, epoch);
          u[0] = (nxe - 1) * delx + (ii - 0.5) / n0 * delx;   // :641
          u[1] = (j + ur0) * delx;
          double v[3];
          orc::shock_velocity(sp, row, (uint32_t)ii, isp, 16u, epoch, c, v);
          u[2] = v[0]; u[3] = v[1]; u[4] = v[2];
          const double v1 = orc::vprofile(sp, u[0], w.nxgs, delx);
          const double gam1 = 1 / std::sqrt(1 - (v1 / c) * (v1 / c));
          const double gamp = std::sqrt(1 + (u[2] * u[2] + u[3] * u[3] + u[4] * u[4]) / (c * c));
          u[2] = gam1 * (u[2] + v1 * gamp);
          const int64_t pid = (int64_t)ii + (int64_t)row * n0 + nptotal[isp - 1];   // :670
          const int64_t neg = -pid;
          std::memcpy(&u[5], &neg, 8);
        }
        R.np2[R.in2(j, isp)] += n0;                                                 // :673-674
        R.cumcnt[R.ic(nxe, j, isp)] = R.cumcnt[R.ic(nxe - 1, j, isp)] + n0;
      }
    }
    shock_boundary_field(w, R, sp, nxe);
  }
}

// one time step; order: 0 Weibel (2d/proj/weibel/app.f90:99-107), 1 reconnection (2d/proj/reconnection/app.f90:99-106),
// 2 shock without the driver's inject/relocate (2d/proj/shock/app.f90:112-118)
void step(World2& w, int order, double u0) {
  for (Rank2& R : w.ranks) particle_solv(w, R, R.gp, R.up, false, w.pusher == 1);
  if (order == 1) for (Rank2& R : w.ranks) bc_particle_x_reflect(w, R, R.gp);
  if (order == 2) for (Rank2& R : w.ranks) bc_injection(w, R, R.gp, u0);
  field_fdtd_i(w, 0);
  if (w.err) return;
  if (order == 0) for (Rank2& R : w.ranks) bc_particle_x_periodic(w, R, R.gp);
  bc_particle_y(w, 0);
  if (w.err) return;
  for (Rank2& R : w.ranks) sort_bucket(w, R, R.up, R.gp);
}

}  // namespace

// ===========================================================================
// C interface (ctypes)
// ===========================================================================
extern "C" {

void* orc2_create(int nx, int ny, int np, int nproc, double delx, double delt, double c, double gfac, const double* q,
                  const double* r, int bc) {
  World2* w = new World2();
  w->np = np;
  w->nxge = w->nxgs + nx - 1; w->nyge = w->nygs + ny - 1;
  w->nxs = w->nxgs; w->nxe = w->nxge;
  w->nproc = nproc; w->bc = bc;
  w->delx = delx; w->delt = delt; w->c = c; w->gfac = gfac;
  w->d_delx = 1.0 / delx; w->d_delt = 1.0 / delt;
  for (int s = 0; s < 2; ++s) { w->q[s] = q[s]; w->r[s] = r[s]; }
  // field__init: 2d/common/field.f90:53-59
  w->f1 = c * delt / delx;
  w->f2 = gfac * w->f1 * w->f1;
  w->f3 = 4.0 * kPi * delx / c;
  w->f4 = 4.0 + std::pow(delx / (c * delt * gfac), 2);
  w->f5 = std::pow(delx / (c * delt * gfac), 2);
  // mpi_set__init: 2d/common/mpi_set.f90:36-47 (1-D slabs in y, periodic neighbours)
  w->ranks.resize(nproc);
  for (int k = 0; k < nproc; ++k) {
    Rank2& R = w->ranks[k];
    R.w = w; R.rank = k;
    orc::para_range(R.nys, R.nye, w->nygs, w->nyge, nproc, k);
    R.nyl = R.nye - R.nys + 1;
    R.nup = (k + 1) % nproc; R.ndown = (k - 1 + nproc) % nproc;
    const size_t npart = (size_t)w->ndim * np * R.nyl * w->nsp;
    const size_t nbox = (size_t)(nx + 4) * (R.nyl + 4);
    R.up.assign(npart, 0.0); R.gp.assign(npart, 0.0);
    R.uf.assign(6 * nbox, 0.0); R.df.assign(6 * nbox, 0.0); R.uj.assign(3 * nbox, 0.0);
    R.gkl.assign((size_t)3 * nx * R.nyl, 0.0);
    R.mom.assign((size_t)7 * (nx + 2) * (R.nyl + 2) * w->nsp, 0.0);
    R.np2.assign((size_t)R.nyl * w->nsp, 0);
    R.cumcnt.assign((size_t)(nx + 1) * R.nyl * w->nsp, 0);
  }
  return w;
}

void orc2_destroy(void* h) { delete (World2*)h; }
int orc2_nranks(void* h) { return (int)((World2*)h)->ranks.size(); }
int orc2_error(void* h) { return ((World2*)h)->err; }
void orc2_clear_error(void* h) { ((World2*)h)->err = 0; }

// out[0..1] = nys,nye ; out[2..3] = nup,ndown
void orc2_rank_geom(void* h, int rank, int* out) {
  Rank2& R = ((World2*)h)->ranks[rank];
  out[0] = R.nys; out[1] = R.nye; out[2] = R.nup; out[3] = R.ndown;
}

// which: 0 up, 1 gp, 2 uf, 3 df, 4 uj, 5 gkl
double* orc2_dptr(void* h, int rank, int which) {
  Rank2& R = ((World2*)h)->ranks[rank];
  switch (which) {
    case 0: return R.up.data();
    case 1: return R.gp.data();
    case 2: return R.uf.data();
    case 3: return R.df.data();
    case 4: return R.uj.data();
    case 5: return R.gkl.data();
    case 6: return R.mom.data();
  }
  return nullptr;
}
int* orc2_iptr(void* h, int rank, int which) {
  Rank2& R = ((World2*)h)->ranks[rank];
  return which == 0 ? R.np2.data() : R.cumcnt.data();
}

void orc2_set_xrange(void* h, int nxs, int nxe) { ((World2*)h)->nxs = nxs; ((World2*)h)->nxe = nxe; }
void orc2_cg_iterations(void* h, int* out) { for (int l = 0; l < 3; ++l) out[l] = ((World2*)h)->cg_ite[l]; }

void orc2_particle_solv(void* h) { World2& w = *(World2*)h; for (Rank2& R : w.ranks) particle_solv(w, R, R.gp, R.up); }
void orc2_particle_solv_vay(void* h) { World2& w = *(World2*)h; for (Rank2& R : w.ranks) particle_solv(w, R, R.gp, R.up, false, true); }
// get_particle_count -- 2d/common/paraio.f90 (the routine of 3d/common/paraio.f90:1007-1085 without k)
long long orc2_pack_particles(void* h, int rank, int mode, double* buf, long long* lcount) {
  World2& w = *(World2*)h;
  Rank2& R = w.ranks[rank];
  long long ip = 0;
  for (int isp = 1; isp <= w.nsp; ++isp) {
    lcount[isp - 1] = 0;
    for (int j = R.nys; j <= R.nye; ++j)
      for (int i = 1; i <= R.np2[R.in2(j, isp)]; ++i) {
        const double* u = &R.up[R.ip(1, i, j, isp)];
        int64_t pid;
        std::memcpy(&pid, &u[w.ndim - 1], 8);
        if (mode == 0 || pid > 0) {
          lcount[isp - 1] += 1;
          if (buf) for (int jp = 0; jp < w.ndim; ++jp) buf[ip * w.ndim + jp] = u[jp];
          ip += 1;
        }
      }
  }
  return ip;
}
void orc2_set_pusher(void* h, int kind) { ((World2*)h)->pusher = kind; }
void orc2_shock_inject(void* h, const orc::ShockPrm* sp, const int* nlinj_rows, unsigned epoch) { shock_inject(*(World2*)h, *sp, nlinj_rows, epoch); }
void orc2_shock_relocate(void* h, const orc::ShockPrm* sp, unsigned epoch) { shock_relocate(*(World2*)h, *sp, epoch); }
int orc2_nxe(void* h) { return ((World2*)h)->nxe; }
void orc2_field_fdtd_i(void* h, int stage) { field_fdtd_i(*(World2*)h, stage); }
// kind: 0 periodic wrap, 1 reflecting walls
void orc2_bc_particle_x(void* h, int kind) {
  World2& w = *(World2*)h;
  for (Rank2& R : w.ranks) { if (kind == 0) bc_particle_x_periodic(w, R, R.gp); else bc_particle_x_reflect(w, R, R.gp); }
}
void orc2_bc_injection(void* h, double u0) { World2& w = *(World2*)h; for (Rank2& R : w.ranks) bc_injection(w, R, R.gp, u0); }
void orc2_bc_particle_y(void* h) { bc_particle_y(*(World2*)h, 0); }
void orc2_sort_bucket(void* h) { World2& w = *(World2*)h; for (Rank2& R : w.ranks) sort_bucket(w, R, R.up, R.gp); }
void orc2_mom_calc(void* h) { mom_calc(*(World2*)h); }
void orc2_step(void* h, int order, double u0) { step(*(World2*)h, order, u0); }

// Deterministic Weibel load -- 2d/proj/weibel/app.f90:311-328 (np2, cumcnt), :389-432 (uf, positions,
// Maxwellian), :437-474 (IDs); Philox stream keyed by the GLOBAL row index (slab-count independent), the
// same counters the 3-D loader and the device loader use (purpose 0: y offset, 2*isp-1 / 2*isp: Box-Muller pairs).
void orc2_load_weibel(void* h, int n0, double v_thi, double v_the, double t_ani, double b0, uint64_t seed) {
  World2& w = *(World2*)h;
  const int nx = w.nx(), ny = w.nyge - w.nygs + 1;
  const double sd[2] = {v_thi, v_the};
  for (Rank2& R : w.ranks) {
    for (size_t t = 0; t < R.uf.size() / 6; ++t) {
      double* f = &R.uf[t * 6];
      f[0] = 0; f[1] = 0; f[2] = b0; f[3] = 0; f[4] = 0; f[5] = 0;
    }
    for (int isp = 1; isp <= w.nsp; ++isp)
      for (int j = R.nys; j <= R.nye; ++j) {
        R.np2[R.in2(j, isp)] = n0 * nx;
        R.cumcnt[R.ic(w.nxgs, j, isp)] = 0;
        for (int i = w.nxgs + 1; i <= w.nxge + 1; ++i) R.cumcnt[R.ic(i, j, isp)] = R.cumcnt[R.ic(i - 1, j, isp)] + n0;
      }
    for (int j = R.nys; j <= R.nye; ++j) {
      const uint32_t row = (uint32_t)(j - w.nygs);
      const int n = R.np2[R.in2(j, 1)];
      for (int ii = 1; ii <= n; ++ii) {
        double u0, u1;
        orc::Philox::uniform2(seed, row, (uint32_t)ii, 0u, u0, u1);
        const double x = (w.nxgs + (w.nxge - w.nxgs + 1) * (ii - 0.5) / n) * w.delx;
        const double y = (j + u0) * w.delx;
        for (int isp = 1; isp <= 2; ++isp) {
          double* u = &R.up[R.ip(1, ii, j, isp)];
          u[0] = x; u[1] = y;
          double a0, a1, b0_, b1_, ns, nc, ms, mc;
          orc::Philox::uniform2(seed, row, (uint32_t)ii, (uint32_t)(2 * isp - 1), a0, a1);
          orc::Philox::uniform2(seed, row, (uint32_t)ii, (uint32_t)(2 * isp), b0_, b1_);
          orc::box_muller(a0, a1, ns, nc);
          orc::box_muller(b0_, b1_, ms, mc);
          u[2] = sd[isp - 1] * ns;
          u[3] = sd[isp - 1] * nc;
          u[4] = t_ani * sd[isp - 1] * ms;
          int64_t pid = (int64_t)(isp - 1) * ((int64_t)n0 * nx * ny) + (int64_t)row * ((int64_t)n0 * nx) + ii;
          int64_t neg = -pid;
          std::memcpy(&u[5], &neg, 8);
        }
      }
    }
    R.gp = R.up;
  }
}

// energy_history -- 2d/proj/weibel/app.f90 (same sums as 3-D): out[0..1] kinetic, out[2] E^2/8pi, out[3] B^2/8pi
void orc2_energy(void* h, double* out) {
  World2& w = *(World2*)h;
  double vene[2] = {0, 0}, efield = 0, bfield = 0;
  for (Rank2& R : w.ranks) {
    for (int isp = 1; isp <= 2; ++isp)
      for (int j = R.nys; j <= R.nye; ++j)
        for (int ii = 1; ii <= R.np2[R.in2(j, isp)]; ++ii) {
          const double* u = &R.up[R.ip(1, ii, j, isp)];
          const double u2 = u[2] * u[2] + u[3] * u[3] + u[4] * u[4];
          vene[isp - 1] += w.r[isp - 1] * (std::sqrt(1.0 + u2 / (w.c * w.c)) - 1.0);
        }
    for (int j = R.nys; j <= R.nye; ++j)
      for (int i = w.nxgs; i <= w.nxge; ++i) {
        const double* f = &R.uf[R.i6(1, i, j)];
        bfield += f[0] * f[0] + f[1] * f[1] + f[2] * f[2];
        efield += f[3] * f[3] + f[4] * f[4] + f[5] * f[5];
      }
  }
  out[0] = vene[0]; out[1] = vene[1]; out[2] = efield / (8.0 * kPi); out[3] = bfield / (8.0 * kPi);
}

// Gauss-law residual max|div E - 4 pi rho| (in-plane: Ex, Ey forward differences; rho with the same
// quadratic spline) over the global grid, periodic in y; in x periodic (bc 0) or over the cells
// nxs+1 .. nxe-2 that no wall rule touches.  which: 0 positions from up, 1 from gp.
void orc2_gauss(void* h, int which, double* out) {
  World2& w = *(World2*)h;
  const int nx = w.nx(), ny = w.nyge - w.nygs + 1;
  std::vector<double> rho((size_t)nx * ny, 0.0), ex(rho.size()), ey(rho.size());
  auto G = [&](int i, int j) {
    i = ((i - w.nxgs) % nx + nx) % nx; j = ((j - w.nygs) % ny + ny) % ny;
    return (size_t)j * nx + i;
  };
  for (Rank2& R : w.ranks) {
    const std::vector<double>& P = which == 0 ? R.up : R.gp;
    for (int isp = 1; isp <= 2; ++isp)
      for (int j = R.nys; j <= R.nye; ++j)
        for (int ii = 1; ii <= R.np2[R.in2(j, isp)]; ++ii) {
          const double* u = &P[R.ip(1, ii, j, isp)];
          int c2[2]; double s[2][3];
          for (int a = 0; a < 2; ++a) {
            c2[a] = (int)std::floor(u[a] * w.d_delx);
            const double dh = u[a] * w.d_delx - 0.5 - c2[a];
            s[a][0] = 0.5 * (0.5 - dh) * (0.5 - dh); s[a][1] = 0.75 - dh * dh; s[a][2] = 0.5 * (0.5 + dh) * (0.5 + dh);
          }
          for (int b = -1; b <= 1; ++b)
            for (int a = -1; a <= 1; ++a) {
              if (w.bc != 0 && (c2[0] + a < w.nxgs || c2[0] + a > w.nxge)) continue;
              rho[G(c2[0] + a, c2[1] + b)] += w.q[isp - 1] * s[0][a + 1] * s[1][b + 1];
            }
        }
    for (int j = R.nys; j <= R.nye; ++j)
      for (int i = w.nxgs; i <= w.nxge; ++i) {
        ex[G(i, j)] = R.uf[R.i6(4, i, j)];
        ey[G(i, j)] = R.uf[R.i6(5, i, j)];
      }
  }
  double res = 0, mx = 0;
  const int i_lo = w.bc == 0 ? w.nxgs : w.nxs + 1, i_hi = w.bc == 0 ? w.nxge : w.nxe - 2;
  for (int j = w.nygs; j <= w.nyge; ++j)
    for (int i = i_lo; i <= i_hi; ++i) {
      const double div = ex[G(i + 1, j)] - ex[G(i, j)] + ey[G(i, j + 1)] - ey[G(i, j)];
      const double rr = 4.0 * kPi * w.delx * rho[G(i, j)];
      res = std::max(res, std::fabs(div - rr));
      mx = std::max(mx, std::fabs(rr));
    }
  out[0] = res; out[1] = mx;
}

}  // extern "C"
