// oracle/oracle3d.cpp -- TEST INFRASTRUCTURE ONLY (see oracle_common.h header).
//
// CPU restatement of WumingPIC's 3-D per-timestep loop, following the reference
// loop nests and array shapes 1:1 (Fortran index bases are kept through the
// accessor functions below).  Pinned bit for bit to the translated reference
// (oracle/f2cxx, tests/test_ref_transpiled.py); see oracle_common.h.
//
//   particle__solv                3d/common/particle.f90:52-233
//   field__init / fdtd_i          3d/common/field.f90:22-67, 70-208
//   ele_cur                       3d/common/field.f90:211-406
//   cgm                           3d/common/field.f90:409-560
//   sort__bucket                  3d/common/sort.f90:40-88
//   boundary_periodic__particle_x 3d/common/boundary_periodic.f90:68-101
//   boundary_periodic__particle_yz                              :104-455
//   boundary_periodic__dfield                                   :458-673
//   boundary_periodic__curre                                    :676-978
//   boundary_periodic__phi                                      :981-1099
//   boundary_reconnection__*       3d/proj/reconnection/boundary_reconnection.f90:69-110 (particle_x), :672-682 (dfield x rule),
//                                  :689-978 (curre: no x treatment), :1094-1116 (phi x rule)
//   boundary_shock__*              3d/proj/shock/boundary_shock.f90:424-469 (injection), :674-686 (dfield), :1086-1108 (phi)
//   mom_calc__accl / __nvt         3d/common/mom_calc.f90:49-216, 219-332 ; boundary_*__mom 3d/common/boundary_periodic.f90:1102-1235
//   mpi_set (rank table, slabs)   3d/common/mpi_set.f90:21-97
//   time loop                     3d/proj/weibel/app.f90:100-108
//   Weibel initial load           3d/proj/weibel/app.f90:298-338, 391-504
//
// MPI ranks are emulated in-process: a World holds nproc_j*nproc_k Ranks and
// every MPI_SENDRECV of the reference becomes one sendrecv() phase executed in
// lock-step over all ranks; MPI_ALLREDUCE becomes a sum in rank order.
#include "oracle_common.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

const double kPi = 4.0 * std::atan(1.0);

struct Rank3;

struct World3 {
  // geometry / constants shared by all ranks (the module-level SAVEd copies)
  int ndim = 7, np = 0, nsp = 2;
  int nxgs = 2, nxge = 0, nygs = 2, nyge = 0, nzgs = 2, nzge = 0;
  int nxs = 0, nxe = 0;
  int nproc_j = 1, nproc_k = 1;
  int bc = 0;  // 0 periodic, 1 reconnection walls, 2 shock walls
  double delx = 1, delt = 1, c = 1, gfac = 0.501, d_delx = 1, d_delt = 1;
  double q[2] = {0, 0}, r[2] = {1, 1};
  double f1 = 0, f2 = 0, f3 = 0, f4 = 0, f5 = 0;
  std::vector<Rank3> ranks;
  int cg_ite[3] = {0, 0, 0};
  int err = 0;  // 1: cgm ite_max, 2: memory over (np2 > np)
  int pusher = 0;  // the pusher step() calls: 0 particle__solv (Buneman-Boris), 1 particle__solv_vay
  int nx() const { return nxge - nxgs + 1; }
};

struct Rank3 {
  const World3* w = nullptr;
  int rank = 0, rj = 0, rk = 0;
  int nys = 0, nye = 0, nzs = 0, nze = 0, nyl = 0, nzl = 0;
  int jup = 0, jdown = 0, kup = 0, kdown = 0;
  std::vector<double> up, gp, uf, df, gkl, uj, mom;
  std::vector<int> np2, cumcnt;
  inline size_t im(int l, int i, int j, int k, int isp) const {  // mom(7,nxgs-1:nxge+1,nys-1:nye+1,nzs-1:nze+1,nsp)
    return (((((size_t)(isp - 1) * (nzl + 2) + (k - (nzs - 1))) * (nyl + 2) + (j - (nys - 1))) * (w->nx() + 2) + (i - (w->nxgs - 1)))) * 7 + (l - 1);
  }

  // ---- Fortran-indexed accessors ------------------------------------------
  inline size_t ip(int d, int ii, int j, int k, int isp) const {  // up/gp(ndim,np,nys:nye,nzs:nze,nsp)
    return ((((size_t)(isp - 1) * nzl + (k - nzs)) * nyl + (j - nys)) * w->np + (size_t)(ii - 1)) * w->ndim + (d - 1);
  }
  inline size_t i6(int cc, int i, int j, int k) const {  // uf/df(6,nxgs-2:nxge+2,nys-2:nye+2,nzs-2:nze+2)
    return ((((size_t)(k - (nzs - 2))) * (nyl + 4) + (j - (nys - 2))) * (w->nx() + 4) + (i - (w->nxgs - 2))) * 6 + (cc - 1);
  }
  inline size_t i3(int cc, int i, int j, int k) const {  // uj(3, same box)
    return ((((size_t)(k - (nzs - 2))) * (nyl + 4) + (j - (nys - 2))) * (w->nx() + 4) + (i - (w->nxgs - 2))) * 3 + (cc - 1);
  }
  inline size_t ig(int cc, int i, int j, int k) const {  // gkl(3,nxgs:nxge,nys:nye,nzs:nze)
    return ((((size_t)(k - nzs)) * nyl + (j - nys)) * w->nx() + (i - w->nxgs)) * 3 + (cc - 1);
  }
  inline size_t ic(int i, int j, int k, int isp) const {  // cumcnt(nxgs:nxge+1,nys:nye,nzs:nze,nsp)
    return (((size_t)(isp - 1) * nzl + (k - nzs)) * nyl + (j - nys)) * (w->nx() + 1) + (i - w->nxgs);
  }
  inline size_t in2(int j, int k, int isp) const {  // np2(nys:nye,nzs:nze,nsp)
    return ((size_t)(isp - 1) * nzl + (k - nzs)) * nyl + (j - nys);
  }
};

// scalar work arrays of cgm: phi,p (1 ghost), r,b,ap (interior) -- field.f90:432-434
struct Cg3 {
  int nxs, nxe, nys, nye, nzs, nze;
  std::vector<double> phi, p, r, b, ap;
  inline size_t i1(int i, int j, int k) const {
    return (((size_t)(k - (nzs - 1))) * (nye - nys + 3) + (j - (nys - 1))) * (nxe - nxs + 3) + (i - (nxs - 1));
  }
  inline size_t i0(int i, int j, int k) const {
    return (((size_t)(k - nzs)) * (nye - nys + 1) + (j - nys)) * (nxe - nxs + 1) + (i - nxs);
  }
};

// One MPI_SENDRECV executed by every rank: rank r sends pack(r) towards `to(r)`
// and receives what `from(r)` sent.
enum Dir { TO_JDOWN, TO_JUP, TO_KDOWN, TO_KUP };
template <class T, class Pack, class Unpack>
void sendrecv(World3& w, Dir d, Pack pack, Unpack unpack) {
  const int R = (int)w.ranks.size();
  std::vector<std::vector<T>> snd(R);
  for (int r = 0; r < R; ++r) pack(w.ranks[r], snd[r]);
  for (int r = 0; r < R; ++r) {
    const Rank3& me = w.ranks[r];
    int src = (d == TO_JDOWN) ? me.jup : (d == TO_JUP) ? me.jdown : (d == TO_KDOWN) ? me.kup : me.kdown;
    unpack(w.ranks[r], snd[src]);
  }
}

// ---------------------------------------------------------------------------
// particle__solv -- 3d/common/particle.f90:52-233
// ---------------------------------------------------------------------------
// accl = true: mom_calc__accl (3d/common/mom_calc.f90:49-216) -- the same gather and Boris rotation with delt/2
// (mom_calc__init :36), no move: gp(1:3) = up(1:3), gp(4:6) = re-centred momenta; gp(7) is not written.
// vay = true: particle__solv_vay (3d/common/particle.f90:236-419, Vay, PoP 15, 056701 (2008)): the same staging and gather,
// then the Vay velocity update (:375-400) and the move with the new gamma (:402-406).
void particle_solv(World3& w, Rank3& R, std::vector<double>& gp, const std::vector<double>& up, bool accl = false, bool vay = false) {
  const int nxs = w.nxs, nxe = w.nxe, nys = R.nys, nye = R.nye, nzs = R.nzs, nze = R.nze;
  const double d_delx = w.d_delx, delt = accl ? w.delt * 5e-1 : w.delt, c = w.c;
  const int tx = nxe - nxs + 3, ty = nye - nys + 3, tz = nze - nzs + 3;
  std::vector<double> tmpf((size_t)6 * tx * ty * tz);
  auto T = [&](int cc, int i, int j, int k) -> double& {
    return tmpf[((((size_t)(k - (nzs - 1))) * ty + (j - (nys - 1))) * tx + (i - (nxs - 1))) * 6 + (cc - 1)];
  };
  const std::vector<double>& uf = R.uf;
  // fields at (i+1/2, j+1/2, k+1/2) -- particle.f90:75-91
#pragma omp parallel for
  for (int k = nzs - 1; k <= nze + 1; ++k)
    for (int j = nys - 1; j <= nye + 1; ++j)
      for (int i = nxs - 1; i <= nxe + 1; ++i) {
        T(1, i, j, k) = 2.5e-1 * (+uf[R.i6(1, i, j, k)] + uf[R.i6(1, i, j + 1, k)]
                                  + uf[R.i6(1, i, j, k + 1)] + uf[R.i6(1, i, j + 1, k + 1)]);
        T(2, i, j, k) = 2.5e-1 * (+uf[R.i6(2, i, j, k)] + uf[R.i6(2, i + 1, j, k)]
                                  + uf[R.i6(2, i, j, k + 1)] + uf[R.i6(2, i + 1, j, k + 1)]);
        T(3, i, j, k) = 2.5e-1 * (+uf[R.i6(3, i, j, k)] + uf[R.i6(3, i + 1, j, k)]
                                  + uf[R.i6(3, i, j + 1, k)] + uf[R.i6(3, i + 1, j + 1, k)]);
        T(4, i, j, k) = 5e-1 * (+uf[R.i6(4, i, j, k)] + uf[R.i6(4, i + 1, j, k)]);
        T(5, i, j, k) = 5e-1 * (+uf[R.i6(5, i, j, k)] + uf[R.i6(5, i, j + 1, k)]);
        T(6, i, j, k) = 5e-1 * (+uf[R.i6(6, i, j, k)] + uf[R.i6(6, i, j, k + 1)]);
      }

  // particle.f90:93-225
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = nzs; k <= nze; ++k)
    for (int j = nys; j <= nye; ++j)
      for (int i = nxs; i <= nxe; ++i)
        for (int isp = 1; isp <= w.nsp; ++isp) {
          const double fac1 = w.q[isp - 1] / w.r[isp - 1] * 5e-1 * delt;
          const double txxx = fac1 * fac1;
          const double fac2 = w.q[isp - 1] * delt / w.r[isp - 1];
          const int i_beg = R.cumcnt[R.ic(i, j, k, isp)] + 1, i_end = R.cumcnt[R.ic(i + 1, j, k, isp)];
          for (int ii = i_beg; ii <= i_end; ++ii) {
            const double* u = &up[R.ip(1, ii, j, k, isp)];
            double* g = &gp[R.ip(1, ii, j, k, isp)];
            double sx[3], sy[3], sz[3];
            double dh = u[0] * d_delx - 5e-1 - i;
            sx[0] = 5e-1 * (5e-1 - dh) * (5e-1 - dh);
            sx[1] = 7.5e-1 - dh * dh;
            sx[2] = 5e-1 * (5e-1 + dh) * (5e-1 + dh);
            dh = u[1] * d_delx - 5e-1 - j;
            sy[0] = 5e-1 * (5e-1 - dh) * (5e-1 - dh);
            sy[1] = 7.5e-1 - dh * dh;
            sy[2] = 5e-1 * (5e-1 + dh) * (5e-1 + dh);
            dh = u[2] * d_delx - 5e-1 - k;
            sz[0] = 5e-1 * (5e-1 - dh) * (5e-1 - dh);
            sz[1] = 7.5e-1 - dh * dh;
            sz[2] = 5e-1 * (5e-1 + dh) * (5e-1 + dh);

            // ((sum_x * shy) summed over y) * shz summed over z, in the reference's order
            double f[6];
            for (int cc = 1; cc <= 6; ++cc) {
              double acc = 0.0;
              for (int kk = -1; kk <= 1; ++kk) {
                double row = (+(+T(cc, i - 1, j - 1, k + kk) * sx[0] + T(cc, i, j - 1, k + kk) * sx[1] + T(cc, i + 1, j - 1, k + kk) * sx[2]) * sy[0]
                              + (+T(cc, i - 1, j, k + kk) * sx[0] + T(cc, i, j, k + kk) * sx[1] + T(cc, i + 1, j, k + kk) * sx[2]) * sy[1]
                              + (+T(cc, i - 1, j + 1, k + kk) * sx[0] + T(cc, i, j + 1, k + kk) * sx[1] + T(cc, i + 1, j + 1, k + kk) * sx[2]) * sy[2])
                             * sz[kk + 1];
                acc = (kk == -1) ? row : acc + row;
              }
              f[cc - 1] = acc;
            }
            const double bpx = f[0], bpy = f[1], bpz = f[2], epx = f[3], epy = f[4], epz = f[5];

            if (vay) {   // particle.f90:375-406
              double uvm1 = u[3], uvm2 = u[4], uvm3 = u[5];
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
              g[3] = s_ * (gam2 * uvm4 + c * ua * taux + gam * (uvm5 * tauz - uvm6 * tauy));
              g[4] = s_ * (gam2 * uvm5 + c * ua * tauy + gam * (uvm6 * taux - uvm4 * tauz));
              g[5] = s_ * (gam2 * uvm6 + c * ua * tauz + gam * (uvm4 * tauy - uvm5 * taux));
              gam = 1.0 / gam;
              g[0] = u[0] + g[3] * delt * gam;
              g[1] = u[1] + g[4] * delt * gam;
              g[2] = u[2] + g[5] * delt * gam;
              continue;
            }

            double uvm1 = u[3] + fac1 * epx;
            double uvm2 = u[4] + fac1 * epy;
            double uvm3 = u[5] + fac1 * epz;

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

            g[3] = uvm1 + fac1 * epx;
            g[4] = uvm2 + fac1 * epy;
            g[5] = uvm3 + fac1 * epz;
            if (accl) { g[0] = u[0]; g[1] = u[1]; g[2] = u[2]; continue; }

            gam = 1.0 / std::sqrt(1.0 + (+g[3] * g[3] + g[4] * g[4] + g[5] * g[5]) / (c * c));
            g[0] = u[0] + g[3] * delt * gam;
            g[1] = u[1] + g[4] * delt * gam;
            g[2] = u[2] + g[5] * delt * gam;
          }
        }

  // particle.f90:227-231 -- ID carry over the whole padded array
  if (w.ndim == 7 && !accl) {
    const size_t n = up.size() / 7;
#pragma omp parallel for
    for (size_t t = 0; t < n; ++t) gp[t * 7 + 6] = up[t * 7 + 6];
  }
}

// ---------------------------------------------------------------------------
// ele_cur -- 3d/common/field.f90:211-406
// ---------------------------------------------------------------------------
void ele_cur(World3& w, Rank3& R, const std::vector<double>& up, const std::vector<double>& gp) {
  const int nxs = w.nxs, nxe = w.nxe, nys = R.nys, nye = R.nye, nzs = R.nzs, nze = R.nze;
  const double d_delx = w.d_delx;
  const double fac = 1.0 / 3.0;
  std::vector<double>& uj = R.uj;
  // field.f90:227-229
  for (int k = nzs - 2; k <= nze + 2; ++k)
    for (int j = nys - 2; j <= nye + 2; ++j)
      for (int i = nxs - 2; i <= nxe + 2; ++i)
        for (int cc = 1; cc <= 3; ++cc) uj[R.i3(cc, i, j, k)] = 0.0;

  // REDUCTION(+:uj) (field.f90:237): one private copy of uj per thread, summed in thread order.
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
#pragma omp for collapse(2) schedule(static)
  for (int k = nzs; k <= nze; ++k)
    for (int j = nys; j <= nye; ++j)
      for (int i = nxs; i <= nxe; ++i) {
        double pjx[5][5][5], pjy[5][5][5], pjz[5][5][5];  // [kp+2][jp+2][ip+2] == (ip,jp,kp) Fortran order
        std::memset(pjx, 0, sizeof(pjx));
        std::memset(pjy, 0, sizeof(pjy));
        std::memset(pjz, 0, sizeof(pjz));
        for (int isp = 1; isp <= w.nsp; ++isp) {
          const double qdxdt = w.q[isp - 1] * w.delx * w.d_delt;
          const int i_beg = R.cumcnt[R.ic(i, j, k, isp)] + 1, i_end = R.cumcnt[R.ic(i + 1, j, k, isp)];
          for (int ii = i_beg; ii <= i_end; ++ii) {
            const double* u = &up[R.ip(1, ii, j, k, isp)];
            const double* g = &gp[R.ip(1, ii, j, k, isp)];
            double s0[3][5], ds[3][5];
            const int cell[3] = {i, j, k};
            for (int a = 0; a < 3; ++a) {
              double dh = u[a] * d_delx - 5e-1 - cell[a];
              s0[a][0] = 0.0;
              s0[a][1] = 5e-1 * (5e-1 - dh) * (5e-1 - dh);
              s0[a][2] = 7.5e-1 - dh * dh;
              s0[a][3] = 5e-1 * (5e-1 + dh) * (5e-1 + dh);
              s0[a][4] = 0.0;
            }
            for (int a = 0; a < 3; ++a) {
              int i2 = (int)(g[a] * d_delx);
              double dh = g[a] * d_delx - 5e-1 - i2;
              int inc = i2 - cell[a];
              double s1_1 = 5e-1 * (5e-1 - dh) * (5e-1 - dh);
              double s1_2 = 7.5e-1 - dh * dh;
              double s1_3 = 5e-1 * (5e-1 + dh) * (5e-1 + dh);
              double smo_1 = -(inc - std::abs(inc)) * 5e-1 + 0;
              double smo_2 = -std::abs(inc) + 1;
              double smo_3 = (inc + std::abs(inc)) * 5e-1 + 0;
              ds[a][0] = s1_1 * smo_1;
              ds[a][1] = s1_1 * smo_2 + s1_2 * smo_1;
              ds[a][2] = s1_2 * smo_2 + s1_3 * smo_1 + s1_1 * smo_3;
              ds[a][3] = s1_3 * smo_2 + s1_2 * smo_3;
              ds[a][4] = s1_3 * smo_3;
            }
            for (int a = 0; a < 3; ++a)
              for (int m = 0; m < 5; ++m) ds[a][m] = ds[a][m] - s0[a][m];
            const double *s0x = s0[0], *s0y = s0[1], *s0z = s0[2], *dsx = ds[0], *dsy = ds[1], *dsz = ds[2];

            for (int kp = 0; kp < 5; ++kp)
              for (int jp = 0; jp < 5; ++jp) {
                // pjx(ip,jp,kp)
                double dstmp = ((s0y[jp] + 5e-1 * dsy[jp]) * s0z[kp] + (5e-1 * s0y[jp] + fac * dsy[jp]) * dsz[kp]) * qdxdt;
                double pjtmp = -dsx[0] * dstmp;
                pjx[kp][jp][1] += pjtmp;
                pjtmp = pjtmp - dsx[1] * dstmp;
                pjx[kp][jp][2] += pjtmp;
                pjtmp = pjtmp - dsx[2] * dstmp;
                pjx[kp][jp][3] += pjtmp;
                pjtmp = pjtmp - dsx[3] * dstmp;
                pjx[kp][jp][4] += pjtmp;
                // pjy(jp_y, ip=jp, kp): running index is the FIRST subscript
                dstmp = ((s0x[jp] + 5e-1 * dsx[jp]) * s0z[kp] + (5e-1 * s0x[jp] + fac * dsx[jp]) * dsz[kp]) * qdxdt;
                pjtmp = -dsy[0] * dstmp;
                pjy[kp][jp][1] += pjtmp;
                pjtmp = pjtmp - dsy[1] * dstmp;
                pjy[kp][jp][2] += pjtmp;
                pjtmp = pjtmp - dsy[2] * dstmp;
                pjy[kp][jp][3] += pjtmp;
                pjtmp = pjtmp - dsy[3] * dstmp;
                pjy[kp][jp][4] += pjtmp;
                // pjz(kp_z, ip=jp, jp=kp)
                dstmp = ((s0x[jp] + 5e-1 * dsx[jp]) * s0y[kp] + (5e-1 * s0x[jp] + fac * dsx[jp]) * dsy[kp]) * qdxdt;
                pjtmp = -dsz[0] * dstmp;
                pjz[kp][jp][1] += pjtmp;
                pjtmp = pjtmp - dsz[1] * dstmp;
                pjz[kp][jp][2] += pjtmp;
                pjtmp = pjtmp - dsz[2] * dstmp;
                pjz[kp][jp][3] += pjtmp;
                pjtmp = pjtmp - dsz[3] * dstmp;
                pjz[kp][jp][4] += pjtmp;
              }
          }
        }
        // field.f90:389-399: uj(1,..)+=pjx(ip,jp,kp); uj(2,..)+=pjy(jp,ip,kp); uj(3,..)+=pjz(kp,ip,jp)
        for (int kp = -2; kp <= 2; ++kp)
          for (int jp = -2; jp <= 2; ++jp)
            for (int ipp = -2; ipp <= 2; ++ipp) {
              ujp[R.i3(1, i + ipp, j + jp, k + kp)] += pjx[kp + 2][jp + 2][ipp + 2];
              ujp[R.i3(2, i + ipp, j + jp, k + kp)] += pjy[kp + 2][ipp + 2][jp + 2];
              ujp[R.i3(3, i + ipp, j + jp, k + kp)] += pjz[jp + 2][ipp + 2][kp + 2];
            }
      }
  }
  for (int t = 0; t < nth; ++t)
    if (!priv[t].empty())
      for (size_t n = 0; n < ujn; ++n) uj[n] += priv[t][n];
}

// ---------------------------------------------------------------------------
// boundary_periodic__curre -- 3d/common/boundary_periodic.f90:676-978
// ---------------------------------------------------------------------------
void bc_curre(World3& w) {
  const int nxs = w.nxs, nxe = w.nxe;
  // 1) y-direction add (2 layers, all i and k incl. ghosts)
  sendrecv<double>(w, TO_JDOWN,
      [&](Rank3& R, std::vector<double>& b) {
        for (int k = R.nzs - 2; k <= R.nze + 2; ++k)
          for (int i = nxs - 2; i <= nxe + 2; ++i) {
            for (int cc = 1; cc <= 3; ++cc) b.push_back(R.uj[R.i3(cc, i, R.nys - 2, k)]);
            for (int cc = 1; cc <= 3; ++cc) b.push_back(R.uj[R.i3(cc, i, R.nys - 1, k)]);
          }
      },
      [&](Rank3& R, const std::vector<double>& b) {
        size_t t = 0;
        for (int k = R.nzs - 2; k <= R.nze + 2; ++k)
          for (int i = nxs - 2; i <= nxe + 2; ++i) {
            for (int cc = 1; cc <= 3; ++cc) R.uj[R.i3(cc, i, R.nye - 1, k)] += b[t++];
            for (int cc = 1; cc <= 3; ++cc) R.uj[R.i3(cc, i, R.nye, k)] += b[t++];
          }
      });
  sendrecv<double>(w, TO_JUP,
      [&](Rank3& R, std::vector<double>& b) {
        for (int k = R.nzs - 2; k <= R.nze + 2; ++k)
          for (int i = nxs - 2; i <= nxe + 2; ++i) {
            for (int cc = 1; cc <= 3; ++cc) b.push_back(R.uj[R.i3(cc, i, R.nye + 1, k)]);
            for (int cc = 1; cc <= 3; ++cc) b.push_back(R.uj[R.i3(cc, i, R.nye + 2, k)]);
          }
      },
      [&](Rank3& R, const std::vector<double>& b) {
        size_t t = 0;
        for (int k = R.nzs - 2; k <= R.nze + 2; ++k)
          for (int i = nxs - 2; i <= nxe + 2; ++i) {
            for (int cc = 1; cc <= 3; ++cc) R.uj[R.i3(cc, i, R.nys, k)] += b[t++];
            for (int cc = 1; cc <= 3; ++cc) R.uj[R.i3(cc, i, R.nys + 1, k)] += b[t++];
          }
      });
  // 2) z-direction add (2 layers, j interior only) (:767-837)
  sendrecv<double>(w, TO_KDOWN,
      [&](Rank3& R, std::vector<double>& b) {
        for (int j = R.nys; j <= R.nye; ++j)
          for (int i = nxs - 2; i <= nxe + 2; ++i) {
            for (int cc = 1; cc <= 3; ++cc) b.push_back(R.uj[R.i3(cc, i, j, R.nzs - 2)]);
            for (int cc = 1; cc <= 3; ++cc) b.push_back(R.uj[R.i3(cc, i, j, R.nzs - 1)]);
          }
      },
      [&](Rank3& R, const std::vector<double>& b) {
        size_t t = 0;
        for (int j = R.nys; j <= R.nye; ++j)
          for (int i = nxs - 2; i <= nxe + 2; ++i) {
            for (int cc = 1; cc <= 3; ++cc) R.uj[R.i3(cc, i, j, R.nze - 1)] += b[t++];
            for (int cc = 1; cc <= 3; ++cc) R.uj[R.i3(cc, i, j, R.nze)] += b[t++];
          }
      });
  sendrecv<double>(w, TO_KUP,
      [&](Rank3& R, std::vector<double>& b) {
        for (int j = R.nys; j <= R.nye; ++j)
          for (int i = nxs - 2; i <= nxe + 2; ++i) {
            for (int cc = 1; cc <= 3; ++cc) b.push_back(R.uj[R.i3(cc, i, j, R.nze + 1)]);
            for (int cc = 1; cc <= 3; ++cc) b.push_back(R.uj[R.i3(cc, i, j, R.nze + 2)]);
          }
      },
      [&](Rank3& R, const std::vector<double>& b) {
        size_t t = 0;
        for (int j = R.nys; j <= R.nye; ++j)
          for (int i = nxs - 2; i <= nxe + 2; ++i) {
            for (int cc = 1; cc <= 3; ++cc) R.uj[R.i3(cc, i, j, R.nzs)] += b[t++];
            for (int cc = 1; cc <= 3; ++cc) R.uj[R.i3(cc, i, j, R.nzs + 1)] += b[t++];
          }
      });
  // 3) y copy-back, 1 layer, k interior (:843-901)
  sendrecv<double>(w, TO_JDOWN,
      [&](Rank3& R, std::vector<double>& b) {
        for (int k = R.nzs; k <= R.nze; ++k)
          for (int i = nxs - 2; i <= nxe + 2; ++i)
            for (int cc = 1; cc <= 3; ++cc) b.push_back(R.uj[R.i3(cc, i, R.nys, k)]);
      },
      [&](Rank3& R, const std::vector<double>& b) {
        size_t t = 0;
        for (int k = R.nzs; k <= R.nze; ++k)
          for (int i = nxs - 2; i <= nxe + 2; ++i)
            for (int cc = 1; cc <= 3; ++cc) R.uj[R.i3(cc, i, R.nye + 1, k)] = b[t++];
      });
  sendrecv<double>(w, TO_JUP,
      [&](Rank3& R, std::vector<double>& b) {
        for (int k = R.nzs; k <= R.nze; ++k)
          for (int i = nxs - 2; i <= nxe + 2; ++i)
            for (int cc = 1; cc <= 3; ++cc) b.push_back(R.uj[R.i3(cc, i, R.nye, k)]);
      },
      [&](Rank3& R, const std::vector<double>& b) {
        size_t t = 0;
        for (int k = R.nzs; k <= R.nze; ++k)
          for (int i = nxs - 2; i <= nxe + 2; ++i)
            for (int cc = 1; cc <= 3; ++cc) R.uj[R.i3(cc, i, R.nys - 1, k)] = b[t++];
      });
  // 4) z copy-back, 1 layer, j in [nys-1,nye+1] (:905-963)
  sendrecv<double>(w, TO_KDOWN,
      [&](Rank3& R, std::vector<double>& b) {
        for (int j = R.nys - 1; j <= R.nye + 1; ++j)
          for (int i = nxs - 2; i <= nxe + 2; ++i)
            for (int cc = 1; cc <= 3; ++cc) b.push_back(R.uj[R.i3(cc, i, j, R.nzs)]);
      },
      [&](Rank3& R, const std::vector<double>& b) {
        size_t t = 0;
        for (int j = R.nys - 1; j <= R.nye + 1; ++j)
          for (int i = nxs - 2; i <= nxe + 2; ++i)
            for (int cc = 1; cc <= 3; ++cc) R.uj[R.i3(cc, i, j, R.nze + 1)] = b[t++];
      });
  sendrecv<double>(w, TO_KUP,
      [&](Rank3& R, std::vector<double>& b) {
        for (int j = R.nys - 1; j <= R.nye + 1; ++j)
          for (int i = nxs - 2; i <= nxe + 2; ++i)
            for (int cc = 1; cc <= 3; ++cc) b.push_back(R.uj[R.i3(cc, i, j, R.nze)]);
      },
      [&](Rank3& R, const std::vector<double>& b) {
        size_t t = 0;
        for (int j = R.nys - 1; j <= R.nye + 1; ++j)
          for (int i = nxs - 2; i <= nxe + 2; ++i)
            for (int cc = 1; cc <= 3; ++cc) R.uj[R.i3(cc, i, j, R.nzs - 1)] = b[t++];
      });
  // 5) x periodic fold + copy-back over all j,k incl. ghosts (:965-976); the wall modules do nothing in x
  if (w.bc != 0) return;
  for (Rank3& R : w.ranks) {
    for (int k = R.nzs - 2; k <= R.nze + 2; ++k)
      for (int j = R.nys - 2; j <= R.nye + 2; ++j)
        for (int cc = 1; cc <= 3; ++cc) {
          R.uj[R.i3(cc, nxe - 1, j, k)] += R.uj[R.i3(cc, nxs - 2, j, k)];
          R.uj[R.i3(cc, nxe, j, k)] += R.uj[R.i3(cc, nxs - 1, j, k)];
          R.uj[R.i3(cc, nxs, j, k)] += R.uj[R.i3(cc, nxe + 1, j, k)];
          R.uj[R.i3(cc, nxs + 1, j, k)] += R.uj[R.i3(cc, nxe + 2, j, k)];
        }
    for (int k = R.nzs - 2; k <= R.nze + 2; ++k)
      for (int j = R.nys - 2; j <= R.nye + 2; ++j)
        for (int cc = 1; cc <= 3; ++cc) {
          R.uj[R.i3(cc, nxs - 2, j, k)] = R.uj[R.i3(cc, nxe - 1, j, k)];
          R.uj[R.i3(cc, nxs - 1, j, k)] = R.uj[R.i3(cc, nxe, j, k)];
          R.uj[R.i3(cc, nxe + 1, j, k)] = R.uj[R.i3(cc, nxs, j, k)];
          R.uj[R.i3(cc, nxe + 2, j, k)] = R.uj[R.i3(cc, nxs + 1, j, k)];
        }
  }
}

// ---------------------------------------------------------------------------
// boundary_periodic__dfield -- 3d/common/boundary_periodic.f90:458-673
// ---------------------------------------------------------------------------
void bc_dfield(World3& w) {
  const int nxs = w.nxs, nxe = w.nxe;
  auto pack_y = [&](int off0, bool from_top) {
    return [=](Rank3& R, std::vector<double>& b) {
      int j0 = from_top ? R.nye + off0 : R.nys + off0;
      for (int k = R.nzs; k <= R.nze; ++k)
        for (int i = nxs; i <= nxe; ++i) {
          for (int cc = 1; cc <= 6; ++cc) b.push_back(R.df[R.i6(cc, i, j0, k)]);
          for (int cc = 1; cc <= 6; ++cc) b.push_back(R.df[R.i6(cc, i, j0 + 1, k)]);
        }
    };
  };
  auto unpack_y = [&](int off0, bool at_top) {
    return [=](Rank3& R, const std::vector<double>& b) {
      int j0 = at_top ? R.nye + off0 : R.nys + off0;
      size_t t = 0;
      for (int k = R.nzs; k <= R.nze; ++k)
        for (int i = nxs; i <= nxe; ++i) {
          for (int cc = 1; cc <= 6; ++cc) R.df[R.i6(cc, i, j0, k)] = b[t++];
          for (int cc = 1; cc <= 6; ++cc) R.df[R.i6(cc, i, j0 + 1, k)] = b[t++];
        }
    };
  };
  sendrecv<double>(w, TO_JDOWN, pack_y(0, false), unpack_y(+1, true));   // nys,nys+1 -> nye+1,nye+2
  sendrecv<double>(w, TO_JUP, pack_y(-1, true), unpack_y(-2, false));    // nye-1,nye -> nys-2,nys-1
  auto pack_z = [&](int off0, bool from_top) {
    return [=](Rank3& R, std::vector<double>& b) {
      int k0 = from_top ? R.nze + off0 : R.nzs + off0;
      for (int j = R.nys - 2; j <= R.nye + 2; ++j)
        for (int i = nxs; i <= nxe; ++i) {
          for (int cc = 1; cc <= 6; ++cc) b.push_back(R.df[R.i6(cc, i, j, k0)]);
          for (int cc = 1; cc <= 6; ++cc) b.push_back(R.df[R.i6(cc, i, j, k0 + 1)]);
        }
    };
  };
  auto unpack_z = [&](int off0, bool at_top) {
    return [=](Rank3& R, const std::vector<double>& b) {
      int k0 = at_top ? R.nze + off0 : R.nzs + off0;
      size_t t = 0;
      for (int j = R.nys - 2; j <= R.nye + 2; ++j)
        for (int i = nxs; i <= nxe; ++i) {
          for (int cc = 1; cc <= 6; ++cc) R.df[R.i6(cc, i, j, k0)] = b[t++];
          for (int cc = 1; cc <= 6; ++cc) R.df[R.i6(cc, i, j, k0 + 1)] = b[t++];
        }
    };
  };
  sendrecv<double>(w, TO_KDOWN, pack_z(0, false), unpack_z(+1, true));
  sendrecv<double>(w, TO_KUP, pack_z(-1, true), unpack_z(-2, false));
  // x periodic copy, all j,k incl. ghosts (:665-670); walls: 3d/proj/reconnection/boundary_reconnection.f90:672-682,
  // 3d/proj/shock/boundary_shock.f90:674-686
  for (Rank3& R : w.ranks)
    for (int k = R.nzs - 2; k <= R.nze + 2; ++k)
      for (int j = R.nys - 2; j <= R.nye + 2; ++j) {
        auto D = [&](int cc, int i) -> double& { return R.df[R.i6(cc, i, j, k)]; };
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
// boundary_periodic__phi -- 3d/common/boundary_periodic.f90:981-1099
// `sel` picks cg.phi (0) or cg.p (1).
// ---------------------------------------------------------------------------
void bc_phi(World3& w, std::vector<Cg3>& cg, int sel, int l) {
  const int nxs = w.nxs, nxe = w.nxe;
  auto A = [&](Rank3& R) -> std::vector<double>& { return sel == 0 ? cg[R.rank].phi : cg[R.rank].p; };
  auto I = [&](Rank3& R, int i, int j, int k) { return cg[R.rank].i1(i, j, k); };
  sendrecv<double>(w, TO_JDOWN,
      [&](Rank3& R, std::vector<double>& b) {
        for (int k = R.nzs; k <= R.nze; ++k)
          for (int i = nxs; i <= nxe; ++i) b.push_back(A(R)[I(R, i, R.nys, k)]);
      },
      [&](Rank3& R, const std::vector<double>& b) {
        size_t t = 0;
        for (int k = R.nzs; k <= R.nze; ++k)
          for (int i = nxs; i <= nxe; ++i) A(R)[I(R, i, R.nye + 1, k)] = b[t++];
      });
  sendrecv<double>(w, TO_JUP,
      [&](Rank3& R, std::vector<double>& b) {
        for (int k = R.nzs; k <= R.nze; ++k)
          for (int i = nxs; i <= nxe; ++i) b.push_back(A(R)[I(R, i, R.nye, k)]);
      },
      [&](Rank3& R, const std::vector<double>& b) {
        size_t t = 0;
        for (int k = R.nzs; k <= R.nze; ++k)
          for (int i = nxs; i <= nxe; ++i) A(R)[I(R, i, R.nys - 1, k)] = b[t++];
      });
  sendrecv<double>(w, TO_KDOWN,
      [&](Rank3& R, std::vector<double>& b) {
        for (int j = R.nys - 1; j <= R.nye + 1; ++j)
          for (int i = nxs; i <= nxe; ++i) b.push_back(A(R)[I(R, i, j, R.nzs)]);
      },
      [&](Rank3& R, const std::vector<double>& b) {
        size_t t = 0;
        for (int j = R.nys - 1; j <= R.nye + 1; ++j)
          for (int i = nxs; i <= nxe; ++i) A(R)[I(R, i, j, R.nze + 1)] = b[t++];
      });
  sendrecv<double>(w, TO_KUP,
      [&](Rank3& R, std::vector<double>& b) {
        for (int j = R.nys - 1; j <= R.nye + 1; ++j)
          for (int i = nxs; i <= nxe; ++i) b.push_back(A(R)[I(R, i, j, R.nze)]);
      },
      [&](Rank3& R, const std::vector<double>& b) {
        size_t t = 0;
        for (int j = R.nys - 1; j <= R.nye + 1; ++j)
          for (int i = nxs; i <= nxe; ++i) A(R)[I(R, i, j, R.nzs - 1)] = b[t++];
      });
  // x: periodic (:1091-1096) or the wall rule of component l (boundary_reconnection.f90:1094-1116, boundary_shock.f90:1086-1108)
  for (Rank3& R : w.ranks)
    for (int k = R.nzs - 1; k <= R.nze + 1; ++k)
      for (int j = R.nys - 1; j <= R.nye + 1; ++j) {
        std::vector<double>& a = A(R);
        if (w.bc == 0) {
          a[I(R, nxs - 1, j, k)] = a[I(R, nxe, j, k)];
          a[I(R, nxe + 1, j, k)] = a[I(R, nxs, j, k)];
        } else if (l == 1) {
          a[I(R, nxs - 1, j, k)] = -a[I(R, nxs, j, k)];
          a[I(R, nxe + 1, j, k)] = w.bc == 1 ? -a[I(R, nxe - 2, j, k)] : 0.0;
        } else {
          a[I(R, nxs - 1, j, k)] = a[I(R, nxs + 1, j, k)];
          a[I(R, nxe + 1, j, k)] = w.bc == 1 ? a[I(R, nxe - 1, j, k)] : 0.0;
        }
      }
}

// ---------------------------------------------------------------------------
// cgm -- 3d/common/field.f90:409-560 (control flow exactly as SURVEY.md §3.3)
// ---------------------------------------------------------------------------
void cgm(World3& w) {
  const int nxs = w.nxs, nxe = w.nxe;
  const int ite_max = 100;
  const double err = 1e-6;
  const int NR = (int)w.ranks.size();
  std::vector<Cg3> cg(NR);
  for (int r = 0; r < NR; ++r) {
    Rank3& R = w.ranks[r];
    Cg3& c = cg[r];
    c.nxs = nxs; c.nxe = nxe; c.nys = R.nys; c.nye = R.nye; c.nzs = R.nzs; c.nze = R.nze;
    size_t n1 = (size_t)(nxe - nxs + 3) * (R.nyl + 2) * (R.nzl + 2), n0 = (size_t)(nxe - nxs + 1) * R.nyl * R.nzl;
    c.phi.assign(n1, 0.0); c.p.assign(n1, 0.0);
    c.r.assign(n0, 0.0); c.b.assign(n0, 0.0); c.ap.assign(n0, 0.0);
  }
  const double f4 = w.f4, f5 = w.f5;
  for (int l = 1; l <= 3; ++l) {
    int ite = 0;
    double sum_g = 0.0;
    for (int r = 0; r < NR; ++r) {
      Rank3& R = w.ranks[r]; Cg3& c = cg[r];
      double sum = 0.0;
#pragma omp parallel for reduction(+ : sum)
      for (int k = R.nzs; k <= R.nze; ++k)
        for (int j = R.nys; j <= R.nye; ++j)
          for (int i = nxs; i <= nxe; ++i) {
            c.phi[c.i1(i, j, k)] = R.df[R.i6(l, i, j, k)];
            double bb = f5 * R.gkl[R.ig(l, i, j, k)];
            c.b[c.i0(i, j, k)] = bb;
            sum = sum + bb * bb;
          }
      sum_g += sum;
    }
    const double eps = std::sqrt(sum_g) * err;
    bc_phi(w, cg, 0, l);
    double sumr_g = 0.0;
    for (int r = 0; r < NR; ++r) {
      Rank3& R = w.ranks[r]; Cg3& c = cg[r];
      double sumr = 0.0;
#pragma omp parallel for reduction(+ : sumr)
      for (int k = R.nzs; k <= R.nze; ++k)
        for (int j = R.nys; j <= R.nye; ++j)
          for (int i = nxs; i <= nxe; ++i) {
            double rr = c.b[c.i0(i, j, k)] + c.phi[c.i1(i, j, k - 1)] + c.phi[c.i1(i, j - 1, k)]
                        + c.phi[c.i1(i - 1, j, k)] - f4 * c.phi[c.i1(i, j, k)] + c.phi[c.i1(i + 1, j, k)]
                        + c.phi[c.i1(i, j + 1, k)] + c.phi[c.i1(i, j, k + 1)];
            c.r[c.i0(i, j, k)] = rr;
            c.p[c.i1(i, j, k)] = rr;
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
          Rank3& R = w.ranks[r]; Cg3& c = cg[r];
          double sumr = 0.0, sum2 = 0.0;
#pragma omp parallel for reduction(+ : sumr, sum2)
          for (int k = R.nzs; k <= R.nze; ++k)
            for (int j = R.nys; j <= R.nye; ++j)
              for (int i = nxs; i <= nxe; ++i) {
                double a = -c.p[c.i1(i, j, k - 1)] - c.p[c.i1(i, j - 1, k)]
                           - c.p[c.i1(i - 1, j, k)] + f4 * c.p[c.i1(i, j, k)] - c.p[c.i1(i + 1, j, k)]
                           - c.p[c.i1(i, j + 1, k)] - c.p[c.i1(i, j, k + 1)];
                c.ap[c.i0(i, j, k)] = a;
                sumr = sumr + c.r[c.i0(i, j, k)] * c.r[c.i0(i, j, k)];
                sum2 = sum2 + c.p[c.i1(i, j, k)] * a;
              }
          s_r += sumr; s_2 += sum2;
        }
        sumr_g = s_r;
        const double sum2_g = s_2;
        const double av = sumr_g / sum2_g;
        for (int r = 0; r < NR; ++r) {
          Rank3& R = w.ranks[r]; Cg3& c = cg[r];
#pragma omp parallel for
          for (int k = R.nzs; k <= R.nze; ++k)
            for (int j = R.nys; j <= R.nye; ++j)
              for (int i = nxs; i <= nxe; ++i) {
                c.phi[c.i1(i, j, k)] = c.phi[c.i1(i, j, k)] + av * c.p[c.i1(i, j, k)];
                c.r[c.i0(i, j, k)] = c.r[c.i0(i, j, k)] - av * c.ap[c.i0(i, j, k)];
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
          Rank3& R = w.ranks[r]; Cg3& c = cg[r];
          double sum1 = 0.0;
#pragma omp parallel for reduction(+ : sum1)
          for (int k = R.nzs; k <= R.nze; ++k)
            for (int j = R.nys; j <= R.nye; ++j)
              for (int i = nxs; i <= nxe; ++i) sum1 = sum1 + c.r[c.i0(i, j, k)] * c.r[c.i0(i, j, k)];
          sum1_g += sum1;
        }
        const double bv = sum1_g / sumr_g;
        for (int r = 0; r < NR; ++r) {
          Rank3& R = w.ranks[r]; Cg3& c = cg[r];
#pragma omp parallel for
          for (int k = R.nzs; k <= R.nze; ++k)
            for (int j = R.nys; j <= R.nye; ++j)
              for (int i = nxs; i <= nxe; ++i)
                c.p[c.i1(i, j, k)] = c.r[c.i0(i, j, k)] + bv * c.p[c.i1(i, j, k)];
        }
      }
    }
    for (int r = 0; r < NR; ++r) {
      Rank3& R = w.ranks[r]; Cg3& c = cg[r];
      for (int k = R.nzs; k <= R.nze; ++k)
        for (int j = R.nys; j <= R.nye; ++j)
          for (int i = nxs; i <= nxe; ++i) R.df[R.i6(l, i, j, k)] = c.phi[c.i1(i, j, k)];
    }
    w.cg_ite[l - 1] = ite;
  }
}

// ---------------------------------------------------------------------------
// field__fdtd_i -- 3d/common/field.f90:70-208
// `which`: 0 = whole routine; the stages are also callable one by one for the
// stage-wise parity tests (1 ele_cur, 2 curre, 3 gkl, 4 cgm, 5 dfield, 6 dE,
// 7 dfield, 8 uf+=df).
// ---------------------------------------------------------------------------
void stage_gkl(World3& w, Rank3& R) {
  const double f1 = w.f1, f2 = w.f2, f3 = w.f3;
  const std::vector<double>&uf = R.uf, &uj = R.uj;
#pragma omp parallel for
  for (int k = R.nzs; k <= R.nze; ++k)
    for (int j = R.nys; j <= R.nye; ++j)
      for (int i = w.nxs; i <= w.nxe; ++i) {
        R.gkl[R.ig(1, i, j, k)] = +f2 * (+uf[R.i6(1, i, j, k - 1)] + uf[R.i6(1, i, j - 1, k)]
                                         + uf[R.i6(1, i - 1, j, k)] - 6.0 * uf[R.i6(1, i, j, k)] + uf[R.i6(1, i + 1, j, k)]
                                         + uf[R.i6(1, i, j + 1, k)] + uf[R.i6(1, i, j, k + 1)]
                                         + f3 * (-uj[R.i3(3, i, j - 1, k)] + uj[R.i3(3, i, j, k)]
                                                 + uj[R.i3(2, i, j, k - 1)] - uj[R.i3(2, i, j, k)]))
                                  - f1 * (-uf[R.i6(6, i, j - 1, k)] + uf[R.i6(6, i, j, k)]
                                          + uf[R.i6(5, i, j, k - 1)] - uf[R.i6(5, i, j, k)]);
        R.gkl[R.ig(2, i, j, k)] = +f2 * (+uf[R.i6(2, i, j, k - 1)] + uf[R.i6(2, i, j - 1, k)]
                                         + uf[R.i6(2, i - 1, j, k)] - 6.0 * uf[R.i6(2, i, j, k)] + uf[R.i6(2, i + 1, j, k)]
                                         + uf[R.i6(2, i, j + 1, k)] + uf[R.i6(2, i, j, k + 1)]
                                         + f3 * (-uj[R.i3(1, i, j, k - 1)] + uj[R.i3(1, i, j, k)]
                                                 + uj[R.i3(3, i - 1, j, k)] - uj[R.i3(3, i, j, k)]))
                                  - f1 * (-uf[R.i6(4, i, j, k - 1)] + uf[R.i6(4, i, j, k)]
                                          + uf[R.i6(6, i - 1, j, k)] - uf[R.i6(6, i, j, k)]);
        R.gkl[R.ig(3, i, j, k)] = +f2 * (+uf[R.i6(3, i, j, k - 1)] + uf[R.i6(3, i, j - 1, k)]
                                         + uf[R.i6(3, i - 1, j, k)] - 6.0 * uf[R.i6(3, i, j, k)] + uf[R.i6(3, i + 1, j, k)]
                                         + uf[R.i6(3, i, j + 1, k)] + uf[R.i6(3, i, j, k + 1)]
                                         + f3 * (-uj[R.i3(2, i - 1, j, k)] + uj[R.i3(2, i, j, k)]
                                                 + uj[R.i3(1, i, j - 1, k)] - uj[R.i3(1, i, j, k)]))
                                  - f1 * (-uf[R.i6(5, i - 1, j, k)] + uf[R.i6(5, i, j, k)]
                                          + uf[R.i6(4, i, j - 1, k)] - uf[R.i6(4, i, j, k)]);
      }
}

void stage_de(World3& w, Rank3& R) {
  const double f1 = w.f1, gfac = w.gfac, delt = w.delt;
  const std::vector<double>&uf = R.uf, &uj = R.uj;
  std::vector<double>& df = R.df;
#pragma omp parallel for
  for (int k = R.nzs; k <= R.nze; ++k)
    for (int j = R.nys; j <= R.nye; ++j)
      for (int i = w.nxs; i <= w.nxe; ++i) {
        df[R.i6(4, i, j, k)] = +f1 * (+gfac * (-df[R.i6(3, i, j, k)] + df[R.i6(3, i, j + 1, k)]
                                               + df[R.i6(2, i, j, k)] - df[R.i6(2, i, j, k + 1)])
                                      + (-uf[R.i6(3, i, j, k)] + uf[R.i6(3, i, j + 1, k)]
                                         + uf[R.i6(2, i, j, k)] - uf[R.i6(2, i, j, k + 1)]))
                               - 4.0 * kPi * delt * uj[R.i3(1, i, j, k)];
        df[R.i6(5, i, j, k)] = +f1 * (+gfac * (-df[R.i6(1, i, j, k)] + df[R.i6(1, i, j, k + 1)]
                                               + df[R.i6(3, i, j, k)] - df[R.i6(3, i + 1, j, k)])
                                      + (-uf[R.i6(1, i, j, k)] + uf[R.i6(1, i, j, k + 1)]
                                         + uf[R.i6(3, i, j, k)] - uf[R.i6(3, i + 1, j, k)]))
                               - 4.0 * kPi * delt * uj[R.i3(2, i, j, k)];
        df[R.i6(6, i, j, k)] = +f1 * (+gfac * (-df[R.i6(2, i, j, k)] + df[R.i6(2, i + 1, j, k)]
                                               + df[R.i6(1, i, j, k)] - df[R.i6(1, i, j + 1, k)])
                                      + (-uf[R.i6(2, i, j, k)] + uf[R.i6(2, i + 1, j, k)]
                                         + uf[R.i6(1, i, j, k)] - uf[R.i6(1, i, j + 1, k)]))
                               - 4.0 * kPi * delt * uj[R.i3(3, i, j, k)];
      }
}

void stage_update(World3& w, Rank3& R) {
#pragma omp parallel for
  for (int k = R.nzs - 2; k <= R.nze + 2; ++k)
    for (int j = R.nys - 2; j <= R.nye + 2; ++j)
      for (int i = w.nxs - 2; i <= w.nxe + 2; ++i)
        for (int cc = 1; cc <= 6; ++cc) R.uf[R.i6(cc, i, j, k)] = R.uf[R.i6(cc, i, j, k)] + R.df[R.i6(cc, i, j, k)];
}

void field_fdtd_i(World3& w, int stage) {
  if (stage == 0 || stage == 1) for (Rank3& R : w.ranks) ele_cur(w, R, R.up, R.gp);
  if (stage == 0 || stage == 2) bc_curre(w);
  if (stage == 0 || stage == 3) for (Rank3& R : w.ranks) stage_gkl(w, R);
  if (stage == 0 || stage == 4) { cgm(w); if (w.err) return; }
  if (stage == 0 || stage == 5) bc_dfield(w);
  if (stage == 0 || stage == 6) for (Rank3& R : w.ranks) stage_de(w, R);
  if (stage == 0 || stage == 7) bc_dfield(w);
  if (stage == 0 || stage == 8) for (Rank3& R : w.ranks) stage_update(w, R);
}

// ---------------------------------------------------------------------------
// boundary_periodic__particle_x -- 3d/common/boundary_periodic.f90:68-101
// ---------------------------------------------------------------------------
void bc_particle_x(World3& w, Rank3& R, std::vector<double>& up) {
  for (int isp = 1; isp <= w.nsp; ++isp) {
#pragma omp parallel for collapse(2)
    for (int k = R.nzs; k <= R.nze; ++k)
      for (int j = R.nys; j <= R.nye; ++j) {
        const int n = R.np2[R.in2(j, k, isp)];
        for (int ii = 1; ii <= n; ++ii) {
          double& x = up[R.ip(1, ii, j, k, isp)];
          int ipos = (int)(x * w.d_delx);
          if (ipos < w.nxgs) x = x + (w.nxge - w.nxgs + 1) * w.delx;
          else if (ipos >= w.nxge + 1) x = x - (w.nxge - w.nxgs + 1) * w.delx;
        }
      }
  }
}

// boundary_reconnection__particle_x -- 3d/proj/reconnection/boundary_reconnection.f90:69-110 (reflecting walls)
void bc_particle_x_reflect(World3& w, Rank3& R, std::vector<double>& up) {
  const int nxs = w.nxs, nxe = w.nxe;
  for (int isp = 1; isp <= w.nsp; ++isp)
    for (int k = R.nzs; k <= R.nze; ++k)
      for (int j = R.nys; j <= R.nye; ++j) {
        const int n = R.np2[R.in2(j, k, isp)];
        for (int ii = 1; ii <= n; ++ii) {
          double* u = &up[R.ip(1, ii, j, k, isp)];
          const int ipos = (int)(u[0] / w.delx);
          if (ipos < nxs + 1) {
            u[0] = 2.0 * (nxs + 1) * w.delx - u[0];
            u[3] = -u[3]; u[4] = -u[4]; u[5] = -u[5];
          } else if (ipos >= nxe - 1) {
            u[0] = 2.0 * (nxe - 1) * w.delx - u[0];
            u[3] = -u[3]; u[4] = -u[4]; u[5] = -u[5];
          }
        }
      }
}

// boundary_shock__injection -- 3d/proj/shock/boundary_shock.f90:424-469
void bc_injection(World3& w, Rank3& R, std::vector<double>& up, double u0) {
  const int nxs = w.nxs, nxe = w.nxe;
  const double xend = nxe * w.delx + u0 / std::sqrt(1.0 + (u0 * u0) / (w.c * w.c)) * w.delt;
  for (int isp = 1; isp <= w.nsp; ++isp)
    for (int k = R.nzs; k <= R.nze; ++k)
      for (int j = R.nys; j <= R.nye; ++j) {
        const int n = R.np2[R.in2(j, k, isp)];
        for (int ii = 1; ii <= n; ++ii) {
          double* u = &up[R.ip(1, ii, j, k, isp)];
          const int ipos = (int)(u[0] * w.d_delx);
          if (ipos < nxs + 1) {
            u[0] = 2.0 * (nxs + 1) * w.delx - u[0];
            u[3] = -u[3]; u[4] = -u[4]; u[5] = -u[5];
          } else if (u[0] > xend) {
            u[0] = 2.0 * xend - u[0];
            u[3] = 2.0 * u0 - u[3];
            u[4] = -u[4]; u[5] = -u[5];
          }
        }
      }
}

// ---------------------------------------------------------------------------
// boundary_periodic__particle_yz -- 3d/common/boundary_periodic.f90:104-455
// (boundary_reconnection__particle_yz and boundary_shock__particle_yz are verbatim copies of it)
// Serial per rank (the reference's arrival order under OpenMP locks is racy;
// the serial order k-outer / j / ii is one admissible outcome).
// ---------------------------------------------------------------------------
struct Mig3 {
  std::vector<std::vector<double>> bff;  // (nyl+2)*(nzl+2) destination pencils
  std::vector<int> cnt, cnt2;
  std::vector<std::vector<int>> flag;
};

void bc_particle_yz(World3& w, int which /*0: gp, 1: up*/) {
  const int NR = (int)w.ranks.size();
  const int ndim = w.ndim;
  std::vector<Mig3> M(NR);
  auto P = [&](Rank3& R) -> std::vector<double>& { return which == 0 ? R.gp : R.up; };
  auto ib = [&](const Rank3& R, int j, int k) { return (size_t)(k - (R.nzs - 1)) * (R.nyl + 2) + (j - (R.nys - 1)); };
  auto i2 = [&](const Rank3& R, int j, int k) { return (size_t)(k - R.nzs) * R.nyl + (j - R.nys); };
  for (int isp = 1; isp <= w.nsp; ++isp) {
    for (int r = 0; r < NR; ++r) {
      Rank3& R = w.ranks[r]; Mig3& m = M[r];
      std::vector<double>& up = P(R);
      m.bff.assign((size_t)(R.nyl + 2) * (R.nzl + 2), {});
      m.cnt.assign((size_t)(R.nyl + 2) * (R.nzl + 2), 0);
      m.cnt2.assign((size_t)R.nyl * R.nzl, 0);
      m.flag.assign((size_t)R.nyl * R.nzl, {});
      for (int k = R.nzs; k <= R.nze; ++k)
        for (int j = R.nys; j <= R.nye; ++j) {
          const int n = R.np2[R.in2(j, k, isp)];
          for (int ii = 1; ii <= n; ++ii) {
            double* u = &up[R.ip(1, ii, j, k, isp)];
            int jpos = (int)(u[1] * w.d_delx);
            int kpos = (int)(u[2] * w.d_delx);
            if (!(jpos == j && kpos == k)) {
              if (jpos <= w.nygs - 1) u[1] = u[1] + (w.nyge - w.nygs + 1) * w.delx;
              else if (jpos >= w.nyge + 1) u[1] = u[1] - (w.nyge - w.nygs + 1) * w.delx;
              if (kpos <= w.nzgs - 1) u[2] = u[2] + (w.nzge - w.nzgs + 1) * w.delx;
              else if (kpos >= w.nzge + 1) u[2] = u[2] - (w.nzge - w.nzgs + 1) * w.delx;
              if (jpos < R.nys - 1 || jpos > R.nye + 1 || kpos < R.nzs - 1 || kpos > R.nze + 1) {
                std::fprintf(stderr, "oracle3d: particle moved more than one pencil (jpos=%d kpos=%d)\n", jpos, kpos);
                w.err = 3;
                return;
              }
              std::vector<double>& b = m.bff[ib(R, jpos, kpos)];
              b.insert(b.end(), u, u + ndim);
              m.cnt[ib(R, jpos, kpos)] += 1;
              m.cnt2[i2(R, j, k)] += 1;
              m.flag[i2(R, j, k)].push_back(ii);
            }
          }
        }
    }
    // the four transfers; counts travel with the payload (the count messages of the
    // reference, :192/:243/:293/:343, are implied by the per-pencil vectors)
    auto xfer = [&](Dir d, bool along_j, bool src_at_top, bool dst_at_top) {
      // along_j: ghost pencils (j_src, nzs-1:nze+1); else (nys:nye, k_src)
      std::vector<std::vector<std::vector<double>>> snd(NR);
      for (int r = 0; r < NR; ++r) {
        Rank3& R = w.ranks[r]; Mig3& m = M[r];
        if (along_j) {
          int js = src_at_top ? R.nye + 1 : R.nys - 1;
          for (int k = R.nzs - 1; k <= R.nze + 1; ++k) snd[r].push_back(m.bff[ib(R, js, k)]);
        } else {
          int ks = src_at_top ? R.nze + 1 : R.nzs - 1;
          for (int j = R.nys; j <= R.nye; ++j) snd[r].push_back(m.bff[ib(R, j, ks)]);
        }
      }
      for (int r = 0; r < NR; ++r) {
        Rank3& R = w.ranks[r]; Mig3& m = M[r];
        int src = (d == TO_JDOWN) ? R.jup : (d == TO_JUP) ? R.jdown : (d == TO_KDOWN) ? R.kup : R.kdown;
        const auto& in = snd[src];
        if (along_j) {
          int jd = dst_at_top ? R.nye : R.nys;
          int t = 0;
          for (int k = R.nzs - 1; k <= R.nze + 1; ++k, ++t) {
            std::vector<double>& b = m.bff[ib(R, jd, k)];
            b.insert(b.end(), in[t].begin(), in[t].end());
            m.cnt[ib(R, jd, k)] += (int)(in[t].size() / ndim);
          }
        } else {
          int kd = dst_at_top ? R.nze : R.nzs;
          int t = 0;
          for (int j = R.nys; j <= R.nye; ++j, ++t) {
            std::vector<double>& b = m.bff[ib(R, j, kd)];
            b.insert(b.end(), in[t].begin(), in[t].end());
            m.cnt[ib(R, j, kd)] += (int)(in[t].size() / ndim);
          }
        }
      }
    };
    xfer(TO_JDOWN, true, false, true);   // (nys-1,*) -> jdown ; from jup appended to (nye,*)   :191-238
    xfer(TO_JUP, true, true, false);     // (nye+1,*) -> jup   ; from jdown appended to (nys,*) :242-289
    xfer(TO_KDOWN, false, false, true);  // (*,nzs-1) -> kdown ; appended to (*,nze)            :293-339
    xfer(TO_KUP, false, true, false);    // (*,nze+1) -> kup   ; appended to (*,nzs)            :343-389

    // hole filling / append -- :395-441
    for (int r = 0; r < NR; ++r) {
      Rank3& R = w.ranks[r]; Mig3& m = M[r];
      std::vector<double>& up = P(R);
      for (int k = R.nzs; k <= R.nze; ++k)
        for (int j = R.nys; j <= R.nye; ++j) {
          int& np2 = R.np2[R.in2(j, k, isp)];
          int& cnt = m.cnt[ib(R, j, k)];
          const std::vector<double>& bff = m.bff[ib(R, j, k)];
          const std::vector<int>& flag = m.flag[i2(R, j, k)];
          const int c2 = m.cnt2[i2(R, j, k)];
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
              for (int d = 1; d <= ndim; ++d) up[R.ip(d, flag[ii - 1], j, k, isp)] = up[R.ip(d, np2, j, k, isp)];
              np2 = np2 - 1;
            } else {
              for (int d = 1; d <= ndim; ++d) up[R.ip(d, flag[ii - 1], j, k, isp)] = bff[(size_t)ndim * iii + (d - 1)];
              iii = iii + 1;
              cnt = cnt - 1;
            }
          }
          if (cnt > 0) {
            if (np2 + cnt > w.np) {
              std::fprintf(stderr, "memory over (np2 > np) %d %d %d %d %d\n", w.np, np2 + cnt, j, k, isp);
              w.err = 2;
              return;
            }
            for (int ii = 1; ii <= cnt; ++ii)
              for (int d = 1; d <= ndim; ++d)
                up[R.ip(d, np2 + ii, j, k, isp)] = bff[(size_t)ndim * iii + (d - 1) + (size_t)ndim * (ii - 1)];
          }
          np2 = np2 + cnt;
        }
    }
  }
}

// ---------------------------------------------------------------------------
// sort__bucket -- 3d/common/sort.f90:40-88   (gp -> up, emits cumcnt)
// ---------------------------------------------------------------------------
void sort_bucket(World3& w, Rank3& R, std::vector<double>& dst, const std::vector<double>& src) {
  const int nxs = w.nxs, nxe = w.nxe, ndim = w.ndim;
  for (int isp = 1; isp <= w.nsp; ++isp) {
#pragma omp parallel for collapse(2)
    for (int k = R.nzs; k <= R.nze; ++k)
      for (int j = R.nys; j <= R.nye; ++j) {
        std::vector<int> cnt(nxe - nxs + 1, 0), sum_cnt(nxe - nxs + 2, 0);
        const int n = R.np2[R.in2(j, k, isp)];
        bool bad = false;
        for (int ii = 1; ii <= n; ++ii) {
          int i = (int)(src[R.ip(1, ii, j, k, isp)]);
          if (i < nxs || i > nxe) { bad = true; break; }  // the reference would index out of bounds here
          cnt[i - nxs] += 1;
        }
        if (bad) { w.err = 4; continue; }
        sum_cnt[0] = 0;
        R.cumcnt[R.ic(nxs, j, k, isp)] = 0;
        for (int i = nxs + 1; i <= nxe + 1; ++i) {
          sum_cnt[i - nxs] = sum_cnt[i - 1 - nxs] + cnt[i - 1 - nxs];
          R.cumcnt[R.ic(i, j, k, isp)] = sum_cnt[i - nxs];
        }
        for (int ii = 1; ii <= n; ++ii) {
          int i = (int)(src[R.ip(1, ii, j, k, isp)]);
          for (int d = 1; d <= ndim; ++d) dst[R.ip(d, sum_cnt[i - nxs] + 1, j, k, isp)] = src[R.ip(d, ii, j, k, isp)];
          sum_cnt[i - nxs] += 1;
        }
      }
  }
}

// ---------------------------------------------------------------------------
// mom_calc__nvt -- 3d/common/mom_calc.f90:219-332 (serial CIC deposit of N, V, T at (i+1/2, j+1/2, k+1/2))
// NB the reference forms dx = x - 0.5 - ih without d_delx (:246-251); reproduced.
// ---------------------------------------------------------------------------
void mom_nvt(World3& w, Rank3& R, const std::vector<double>& up) {
  std::fill(R.mom.begin(), R.mom.end(), 0.0);
  for (int isp = 1; isp <= w.nsp; ++isp)
    for (int k = R.nzs; k <= R.nze; ++k)
      for (int j = R.nys; j <= R.nye; ++j) {
        const int n = R.np2[R.in2(j, k, isp)];
        for (int ii = 1; ii <= n; ++ii) {
          const double* u = &up[R.ip(1, ii, j, k, isp)];
          const int ih = (int)std::floor(u[0] * w.d_delx - 5e-1);
          const int jh = (int)std::floor(u[1] * w.d_delx - 5e-1);
          const int kh = (int)std::floor(u[2] * w.d_delx - 5e-1);
          const double dx = u[0] - 5e-1 - ih, dxm = 1.0 - dx;
          const double dy = u[1] - 5e-1 - jh, dym = 1.0 - dy;
          const double dz = u[2] - 5e-1 - kh, dzm = 1.0 - dz;
          const double gam = 1.0 / std::sqrt(1.0 + (+u[3] * u[3] + u[4] * u[4] + u[5] * u[5]) / (w.c * w.c));
          const double wx[2] = {dxm, dx}, wy[2] = {dym, dy}, wz[2] = {dzm, dz};
          // per moment the eight nodes in the reference's order (x fastest, then y, then z), weights as wx*wy*wz
          const double val[7] = {1.0, u[3] * gam, u[4] * gam, u[5] * gam, u[3] * u[3] * gam, u[4] * u[4] * gam, u[5] * u[5] * gam};
          for (int l = 1; l <= 7; ++l)
            for (int c = 0; c < 2; ++c)
              for (int b = 0; b < 2; ++b)
                for (int a = 0; a < 2; ++a) {
                  double& m = R.mom[R.im(l, ih + a, jh + b, kh + c, isp)];
                  m = l == 1 ? m + wx[a] * wy[b] * wz[c] : m + val[l - 1] * wx[a] * wy[b] * wz[c];
                }
        }
      }
}

// boundary_periodic__mom -- 3d/common/boundary_periodic.f90:1102-1235; walls (x fold onto the same side):
// 3d/proj/reconnection/boundary_reconnection.f90:1122-1258, 3d/proj/shock/boundary_shock.f90:1112-1249
void bc_mom(World3& w) {
  for (Rank3& R : w.ranks)
    for (int isp = 1; isp <= w.nsp; ++isp)
      for (int k = R.nzs - 1; k <= R.nze + 1; ++k)
        for (int j = R.nys - 1; j <= R.nye + 1; ++j)
          for (int l = 1; l <= 7; ++l) {
            if (w.bc == 0) {
              R.mom[R.im(l, w.nxgs, j, k, isp)] += R.mom[R.im(l, w.nxge + 1, j, k, isp)];
              R.mom[R.im(l, w.nxge, j, k, isp)] += R.mom[R.im(l, w.nxgs - 1, j, k, isp)];
            } else {
              R.mom[R.im(l, w.nxgs, j, k, isp)] += R.mom[R.im(l, w.nxgs - 1, j, k, isp)];
              R.mom[R.im(l, w.nxge, j, k, isp)] += R.mom[R.im(l, w.nxge + 1, j, k, isp)];
            }
          }
  for (int isp = 1; isp <= w.nsp; ++isp) {
    auto xfer = [&](Dir d, bool along_j, int src_off, bool src_top, int dst_off, bool dst_top) {
      sendrecv<double>(w, d,
          [&](Rank3& R, std::vector<double>& b) {
            if (along_j) {
              const int js = (src_top ? R.nye : R.nys) + src_off;
              for (int k = R.nzs - 1; k <= R.nze + 1; ++k)
                for (int i = w.nxgs - 1; i <= w.nxge + 1; ++i)
                  for (int l = 1; l <= 7; ++l) b.push_back(R.mom[R.im(l, i, js, k, isp)]);
            } else {
              const int ks = (src_top ? R.nze : R.nzs) + src_off;
              for (int j = R.nys; j <= R.nye; ++j)
                for (int i = w.nxgs - 1; i <= w.nxge + 1; ++i)
                  for (int l = 1; l <= 7; ++l) b.push_back(R.mom[R.im(l, i, j, ks, isp)]);
            }
          },
          [&](Rank3& R, const std::vector<double>& b) {
            size_t t = 0;
            if (along_j) {
              const int jd = (dst_top ? R.nye : R.nys) + dst_off;
              for (int k = R.nzs - 1; k <= R.nze + 1; ++k)
                for (int i = w.nxgs - 1; i <= w.nxge + 1; ++i)
                  for (int l = 1; l <= 7; ++l) R.mom[R.im(l, i, jd, k, isp)] += b[t++];
            } else {
              const int kd = (dst_top ? R.nze : R.nzs) + dst_off;
              for (int j = R.nys; j <= R.nye; ++j)
                for (int i = w.nxgs - 1; i <= w.nxge + 1; ++i)
                  for (int l = 1; l <= 7; ++l) R.mom[R.im(l, i, j, kd, isp)] += b[t++];
            }
          });
    };
    xfer(TO_JDOWN, true, -1, false, 0, true);    // row nys-1 -> jdown, added into its nye
    xfer(TO_JUP, true, +1, true, 0, false);      // row nye+1 -> jup, added into its nys
    xfer(TO_KDOWN, false, -1, false, 0, true);   // plane nzs-1 -> kdown, added into its nze (j interior)
    xfer(TO_KUP, false, +1, true, 0, false);     // plane nze+1 -> kup, added into its nzs
  }
}

// the moment block of the drivers (3d/proj/weibel/app.f90:121-124): accl (gp <- up), nvt (mom <- gp), bc__mom
void mom_calc(World3& w) {
  for (Rank3& R : w.ranks) particle_solv(w, R, R.gp, R.up, true);
  for (Rank3& R : w.ranks) mom_nvt(w, R, R.gp);
  bc_mom(w);
}


// the upstream field columns both inject() and relocate() reset (3d/proj/shock/app.f90:715-726, 893-904)
void shock_boundary_field(World3& w, Rank3& R, const orc::ShockPrm& sp, int nxe) {
  const double by = sp.b0 * std::sin(sp.theta_bn) * std::cos(sp.phi_bn), bz = sp.b0 * std::sin(sp.theta_bn) * std::sin(sp.phi_bn);
  for (int k = R.nzs - 2; k <= R.nze + 2; ++k)
    for (int j = R.nys - 2; j <= R.nye + 2; ++j) {
      R.uf[R.i6(2, nxe - 1, j, k)] = by;
      R.uf[R.i6(3, nxe - 1, j, k)] = bz;
      R.uf[R.i6(5, nxe - 1, j, k)] = +sp.v0 * R.uf[R.i6(3, nxe - 1, j, k)] / w.c;
      R.uf[R.i6(6, nxe - 1, j, k)] = -sp.v0 * R.uf[R.i6(2, nxe - 1, j, k)] / w.c;
      R.uf[R.i6(2, nxe, j, k)] = by;
      R.uf[R.i6(3, nxe, j, k)] = bz;
    }
}

// ---------------------------------------------------------------------------
// inject -- 3d/proj/shock/app.f90:733-906; the row counts nlinj_grid (:747-790) come from the caller per GLOBAL row
// r = (k - nzgs) ny + (j - nygs), which is also the order of ncinj_grid (:799-812) for z slabs (nproc_j = 1).
// ---------------------------------------------------------------------------
void shock_inject(World3& w, const orc::ShockPrm& sp, const int* nlinj_rows, uint32_t epoch) {
  const int nxe = w.nxe, ny = w.nyge - w.nygs + 1;
  const double delx = w.delx, delt = w.delt, c = w.c;
  long long nptotal[2] = {0, 0};
  for (Rank3& R : w.ranks)
    for (int isp = 1; isp <= 2; ++isp)
      for (int k = R.nzs; k <= R.nze; ++k)
        for (int j = R.nys; j <= R.nye; ++j) nptotal[isp - 1] += R.np2[R.in2(j, k, isp)];
  const double x0 = std::fabs(sp.v0) * delt;
  for (Rank3& R : w.ranks) {
    for (int k = R.nzs; k <= R.nze; ++k)
      for (int j = R.nys; j <= R.nye; ++j) {
        const uint32_t row = (uint32_t)((k - w.nzgs) * ny + (j - w.nygs));
        const int n = nlinj_rows[row];
        long long ncinj = 0;
        for (uint32_t rr = 0; rr < row; ++rr) ncinj += nlinj_rows[rr];
        for (int isp = 1; isp <= 2; ++isp) {
          const int base = R.np2[R.in2(j, k, isp)];
          if (base + n > w.np) { w.err = 2; return; }
          for (int ii = 1; ii <= n; ++ii) {
            double* u = &R.up[R.ip(1, base + ii, j, k, isp)];
            double ur0, ur1;
            orc::Philox::uniform2(sp.seed, row, (uint32_t)ii, 0u, ur0, ur1, epoch);
            u[0] = nxe * delx + (ii - 5e-1) / n * x0;     // :822
            u[1] = (j + ur0) * delx;                      // :823
            u[2] = (k + ur1) * delx;                      // :824
            double v[3];
            orc::shock_velocity(sp, row, (uint32_t)ii, isp, 0u, epoch, c, v);
            u[3] = v[0]; u[4] = v[1]; u[5] = v[2];
            u[0] = u[0] + (sp.v0 + u[3]) * delt;          // :853-854
            const double v1 = orc::vprofile(sp, u[0], w.nxgs, delx);
            const double gam1 = 1e0 / std::sqrt(1e0 - (v1 / c) * (v1 / c));
            const double gamp = std::sqrt(1e0 + (u[3] * u[3] + u[4] * u[4] + u[5] * u[5]) / (c * c));
            u[3] = gam1 * (u[3] + v1 * gamp);             // :873
            const int64_t pid = (int64_t)ii + ncinj + nptotal[isp - 1];   // :876
            const int64_t neg = -pid;
            std::memcpy(&u[6], &neg, 8);
          }
        }
        for (int isp = 1; isp <= 2; ++isp) {              // :884-887
          R.np2[R.in2(j, k, isp)] += n;
          R.cumcnt[R.ic(nxe, j, k, isp)] += n;
        }
      }
    shock_boundary_field(w, R, sp, nxe);
  }
}

// relocate -- 3d/proj/shock/app.f90:644-728
void shock_relocate(World3& w, const orc::ShockPrm& sp, uint32_t epoch) {
  if (w.nxe == w.nxge) return;
  w.nxe = w.nxe + 1;
  const int nxe = w.nxe, n0 = sp.n0, ny = w.nyge - w.nygs + 1;
  const double delx = w.delx, c = w.c;
  long long nptotal[2] = {0, 0};
  for (Rank3& R : w.ranks)
    for (int isp = 1; isp <= 2; ++isp)
      for (int k = R.nzs; k <= R.nze; ++k)
        for (int j = R.nys; j <= R.nye; ++j) nptotal[isp - 1] += R.np2[R.in2(j, k, isp)];
  for (Rank3& R : w.ranks) {
    for (int k = R.nzs; k <= R.nze; ++k)
      for (int j = R.nys; j <= R.nye; ++j) {
        const uint32_t row = (uint32_t)((k - w.nzgs) * ny + (j - w.nygs));
        for (int isp = 1; isp <= 2; ++isp) {
          const int base = R.np2[R.in2(j, k, isp)];
          if (base + n0 > w.np) { w.err = 2; return; }
          for (int ii = 1; ii <= n0; ++ii) {
            double* u = &R.up[R.ip(1, base + ii, j, k, isp)];
            double ur0, ur1;
            orc::Philox::uniform2(sp.seed, row, (uint32_t)ii, 16u, ur0, ur1, epoch);
            u[0] = (nxe - 1) * delx + (ii - 5e-1) / n0 * delx;   // :670
            u[1] = (j + ur0) * delx;
            u[2] = (k + ur1) * delx;
            double v[3];
            orc::shock_velocity(sp, row, (uint32_t)ii, isp, 16u, epoch, c, v);
            u[3] = v[0]; u[4] = v[1]; u[5] = v[2];
            const double v1 = orc::vprofile(sp, u[0], w.nxgs, delx);
            const double gam1 = 1e0 / std::sqrt(1e0 - (v1 / c) * (v1 / c));
            const double gamp = std::sqrt(1e0 + (u[3] * u[3] + u[4] * u[4] + u[5] * u[5]) / (c * c));
            u[3] = gam1 * (u[3] + v1 * gamp);
            const int64_t pid = (int64_t)ii + (int64_t)row * n0 + nptotal[isp - 1];   // :704
            const int64_t neg = -pid;
            std::memcpy(&u[6], &neg, 8);
          }
          R.np2[R.in2(j, k, isp)] += n0;                                             // :707-708
          R.cumcnt[R.ic(nxe, j, k, isp)] = R.cumcnt[R.ic(nxe - 1, j, k, isp)] + n0;
        }
      }
    shock_boundary_field(w, R, sp, nxe);
  }
}

// one time step; order 0: Weibel/beam (3d/proj/weibel/app.f90:100-108), 1: reconnection (3d/proj/reconnection/app.f90:103-108),
// 2: shock without the driver's inject/relocate (3d/proj/shock/app.f90, same call order as 2d/proj/shock/app.f90:112-118)
void step(World3& w, int order = 0, double u0 = 0.0) {
  for (Rank3& R : w.ranks) particle_solv(w, R, R.gp, R.up, false, w.pusher == 1);
  if (order == 1) for (Rank3& R : w.ranks) bc_particle_x_reflect(w, R, R.gp);
  if (order == 2) for (Rank3& R : w.ranks) bc_injection(w, R, R.gp, u0);
  field_fdtd_i(w, 0);
  if (w.err) return;
  if (order == 0) for (Rank3& R : w.ranks) bc_particle_x(w, R, R.gp);
  bc_particle_yz(w, 0);
  if (w.err) return;
  for (Rank3& R : w.ranks) sort_bucket(w, R, R.up, R.gp);
}

}  // namespace

// ===========================================================================
// C interface (ctypes)
// ===========================================================================
extern "C" {

// nx,ny,nz: global cells; np: pencil capacity; q,r: per species.
void* orc3_create(int nx, int ny, int nz, int np, int nproc_j, int nproc_k, double delx, double delt, double c,
                  double gfac, const double* q, const double* r, int bc) {
  World3* w = new World3();
  w->np = np;
  w->nxge = w->nxgs + nx - 1; w->nyge = w->nygs + ny - 1; w->nzge = w->nzgs + nz - 1;
  w->nxs = w->nxgs; w->nxe = w->nxge;
  w->nproc_j = nproc_j; w->nproc_k = nproc_k; w->bc = bc;
  w->delx = delx; w->delt = delt; w->c = c; w->gfac = gfac;
  w->d_delx = 1.0 / delx; w->d_delt = 1.0 / delt;
  for (int s = 0; s < 2; ++s) { w->q[s] = q[s]; w->r[s] = r[s]; }
  // field__init: 3d/common/field.f90:57-63
  w->f1 = c * delt / delx;
  w->f2 = gfac * w->f1 * w->f1;
  w->f3 = 4.0 * kPi * delx / c;
  w->f4 = 6.0 + std::pow(delx / (c * delt * gfac), 2);
  w->f5 = std::pow(delx / (c * delt * gfac), 2);
  // mpi_set__init: 3d/common/mpi_set.f90:45-76 (rank = j*nproc_k + k, periodic neighbours)
  const int NR = nproc_j * nproc_k;
  w->ranks.resize(NR);
  auto rk_of = [&](int j, int k) { return ((j + nproc_j) % nproc_j) * nproc_k + ((k + nproc_k) % nproc_k); };
  for (int j = 0; j < nproc_j; ++j)
    for (int k = 0; k < nproc_k; ++k) {
      Rank3& R = w->ranks[j * nproc_k + k];
      R.w = w; R.rank = j * nproc_k + k; R.rj = j; R.rk = k;
      orc::para_range(R.nys, R.nye, w->nygs, w->nyge, nproc_j, j);
      orc::para_range(R.nzs, R.nze, w->nzgs, w->nzge, nproc_k, k);
      R.nyl = R.nye - R.nys + 1; R.nzl = R.nze - R.nzs + 1;
      R.jup = rk_of(j + 1, k); R.jdown = rk_of(j - 1, k); R.kup = rk_of(j, k + 1); R.kdown = rk_of(j, k - 1);
      size_t npart = (size_t)w->ndim * np * R.nyl * R.nzl * w->nsp;
      size_t nbox = (size_t)(nx + 4) * (R.nyl + 4) * (R.nzl + 4);
      R.up.assign(npart, 0.0); R.gp.assign(npart, 0.0);
      R.uf.assign(6 * nbox, 0.0); R.df.assign(6 * nbox, 0.0); R.uj.assign(3 * nbox, 0.0);
      R.gkl.assign((size_t)3 * nx * R.nyl * R.nzl, 0.0);
      R.mom.assign((size_t)7 * (nx + 2) * (R.nyl + 2) * (R.nzl + 2) * w->nsp, 0.0);
      R.np2.assign((size_t)R.nyl * R.nzl * w->nsp, 0);
      R.cumcnt.assign((size_t)(nx + 1) * R.nyl * R.nzl * w->nsp, 0);
    }
  return w;
}

void orc3_destroy(void* h) { delete (World3*)h; }

int orc3_nranks(void* h) { return (int)((World3*)h)->ranks.size(); }
// the OpenMP team the parallel regions of this library will really use (what bench.py reports as `cores`): a launcher such as
// torch.distributed.run exports OMP_NUM_THREADS=1, which the runtime obeys -- the count must be set and read back, not assumed
int orc_num_threads(void) { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
// threads that actually entered a parallel region (a cgroup / nested-parallelism limit can cut the team below the request)
int orc_team_size(void) {
  int n = 1;
#pragma omp parallel
  {
#pragma omp single
    n = omp_get_num_threads();
  }
  return n;
}
// the keyed draws of the oracle's loaders and particle source, for tests that feed the TRANSLATED reference's uniform_rand() /
// normal_rand() (utils/wuming_utils.f90) the values the oracle gives the same particle (tests/test_ref_driver_procs.py)
void orc_philox_uniform2(unsigned long long seed, unsigned stream, unsigned idx, unsigned purpose, unsigned epoch, double* out) {
  orc::Philox::uniform2(seed, stream, idx, purpose, out[0], out[1], epoch);
}
void orc_box_muller(double x1, double x2, double* out) { orc::box_muller(x1, x2, out[0], out[1]); }
int orc3_error(void* h) { return ((World3*)h)->err; }
void orc3_clear_error(void* h) { ((World3*)h)->err = 0; }

// out[0..3] = nys,nye,nzs,nze ; out[4..7] = jup,jdown,kup,kdown
void orc3_rank_geom(void* h, int rank, int* out) {
  Rank3& R = ((World3*)h)->ranks[rank];
  out[0] = R.nys; out[1] = R.nye; out[2] = R.nzs; out[3] = R.nze;
  out[4] = R.jup; out[5] = R.jdown; out[6] = R.kup; out[7] = R.kdown;
}

// which: 0 up, 1 gp, 2 uf, 3 df, 4 uj, 5 gkl
double* orc3_dptr(void* h, int rank, int which) {
  Rank3& R = ((World3*)h)->ranks[rank];
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
// which: 0 np2, 1 cumcnt
int* orc3_iptr(void* h, int rank, int which) {
  Rank3& R = ((World3*)h)->ranks[rank];
  return which == 0 ? R.np2.data() : R.cumcnt.data();
}

void orc3_set_xrange(void* h, int nxs, int nxe) { ((World3*)h)->nxs = nxs; ((World3*)h)->nxe = nxe; }
void orc3_cg_iterations(void* h, int* out) { for (int l = 0; l < 3; ++l) out[l] = ((World3*)h)->cg_ite[l]; }

void orc3_particle_solv(void* h) { World3& w = *(World3*)h; for (Rank3& R : w.ranks) particle_solv(w, R, R.gp, R.up); }
void orc3_particle_solv_vay(void* h) { World3& w = *(World3*)h; for (Rank3& R : w.ranks) particle_solv(w, R, R.gp, R.up, false, true); }
// get_particle_count -- 3d/common/paraio.f90:1007-1085: pack every active particle (mode 0) or the tracers with a positive ID
// (mode 1) of one rank in (isp, k, j, i) order; returns the number of packed records, lcount[isp] as in the reference
long long orc3_pack_particles(void* h, int rank, int mode, double* buf, long long* lcount) {
  World3& w = *(World3*)h;
  Rank3& R = w.ranks[rank];
  long long ip = 0;
  for (int isp = 1; isp <= w.nsp; ++isp) {
    lcount[isp - 1] = 0;
    for (int k = R.nzs; k <= R.nze; ++k)
      for (int j = R.nys; j <= R.nye; ++j)
        for (int i = 1; i <= R.np2[R.in2(j, k, isp)]; ++i) {
          const double* u = &R.up[R.ip(1, i, j, k, isp)];
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
void orc3_set_pusher(void* h, int kind) { ((World3*)h)->pusher = kind; }
void orc3_shock_inject(void* h, const orc::ShockPrm* sp, const int* nlinj_rows, unsigned epoch) { shock_inject(*(World3*)h, *sp, nlinj_rows, epoch); }
void orc3_shock_relocate(void* h, const orc::ShockPrm* sp, unsigned epoch) { shock_relocate(*(World3*)h, *sp, epoch); }
int orc3_nxe(void* h) { return ((World3*)h)->nxe; }
void orc3_field_fdtd_i(void* h, int stage) { field_fdtd_i(*(World3*)h, stage); }
void orc3_bc_particle_x(void* h) { World3& w = *(World3*)h; for (Rank3& R : w.ranks) bc_particle_x(w, R, R.gp); }
void orc3_bc_particle_yz(void* h) { bc_particle_yz(*(World3*)h, 0); }
void orc3_sort_bucket(void* h) { World3& w = *(World3*)h; for (Rank3& R : w.ranks) sort_bucket(w, R, R.up, R.gp); }
void orc3_mom_calc(void* h) { mom_calc(*(World3*)h); }
void orc3_step(void* h) { step(*(World3*)h); }
void orc3_step_order(void* h, int order, double u0) { step(*(World3*)h, order, u0); }
void orc3_bc_particle_x_reflect(void* h) { World3& w = *(World3*)h; for (Rank3& R : w.ranks) bc_particle_x_reflect(w, R, R.gp); }
void orc3_bc_injection(void* h, double u0) { World3& w = *(World3*)h; for (Rank3& R : w.ranks) bc_injection(w, R, R.gp, u0); }

// ---------------------------------------------------------------------------
// Deterministic Weibel load -- 3d/proj/weibel/app.f90:311-338 (np2, cumcnt),
// :391-453 (uf, positions, Maxwellian), :458-504 (IDs), with the Philox stream
// keyed by the GLOBAL pencil index so the state is slab-count independent.
// IDs: the reference numbers rank-major; here pid = global pencil offset + ii
// (also slab independent), stored as transfer(-pid, 1d0).
// ---------------------------------------------------------------------------
void orc3_load_weibel(void* h, int n0, double v_thi, double v_the, double t_ani, double b0, uint64_t seed) {
  World3& w = *(World3*)h;
  const int nx = w.nx(), ny = w.nyge - w.nygs + 1;
  const double sd[2] = {v_thi, v_the};
  for (Rank3& R : w.ranks) {
    for (size_t t = 0; t < R.uf.size() / 6; ++t) {
      double* f = &R.uf[t * 6];
      f[0] = 0; f[1] = 0; f[2] = b0; f[3] = 0; f[4] = 0; f[5] = 0;
    }
    for (int isp = 1; isp <= w.nsp; ++isp)
      for (int k = R.nzs; k <= R.nze; ++k)
        for (int j = R.nys; j <= R.nye; ++j) {
          R.np2[R.in2(j, k, isp)] = n0 * nx;
          R.cumcnt[R.ic(w.nxgs, j, k, isp)] = 0;
          for (int i = w.nxgs + 1; i <= w.nxge + 1; ++i)
            R.cumcnt[R.ic(i, j, k, isp)] = R.cumcnt[R.ic(i - 1, j, k, isp)] + n0;
        }
#pragma omp parallel for collapse(2)
    for (int k = R.nzs; k <= R.nze; ++k)
      for (int j = R.nys; j <= R.nye; ++j) {
        const uint32_t pencil = (uint32_t)((j - w.nygs) + (size_t)ny * (k - w.nzgs));
        const int n = R.np2[R.in2(j, k, 1)];
        for (int ii = 1; ii <= n; ++ii) {
          double u0, u1;
          orc::Philox::uniform2(seed, pencil, (uint32_t)ii, 0u, u0, u1);
          const double x = (w.nxgs + (w.nxge - w.nxgs + 1) * (ii - 5e-1) / n) * w.delx;
          const double y = (j + u0) * w.delx;
          const double z = (k + u1) * w.delx;
          for (int isp = 1; isp <= 2; ++isp) {
            double* u = &R.up[R.ip(1, ii, j, k, isp)];
            u[0] = x; u[1] = y; u[2] = z;
            double a0, a1, b0_, b1_, ns, nc, ms, mc;
            orc::Philox::uniform2(seed, pencil, (uint32_t)ii, (uint32_t)(2 * isp - 1), a0, a1);
            orc::Philox::uniform2(seed, pencil, (uint32_t)ii, (uint32_t)(2 * isp), b0_, b1_);
            orc::box_muller(a0, a1, ns, nc);
            orc::box_muller(b0_, b1_, ms, mc);
            u[3] = sd[isp - 1] * ns;
            u[4] = sd[isp - 1] * nc;
            u[5] = t_ani * sd[isp - 1] * ms;
            int64_t pid = (int64_t)(isp - 1) * ((int64_t)n0 * nx * ny * (w.nzge - w.nzgs + 1))
                          + (int64_t)pencil * ((int64_t)n0 * nx) + ii;
            int64_t neg = -pid;
            std::memcpy(&u[6], &neg, 8);
          }
        }
      }
    R.gp = R.up;  // app.f90:374
  }
}

// energy_history -- 3d/proj/weibel/app.f90:509-577: out[0..1] kinetic per species, out[2] E^2/8pi, out[3] B^2/8pi
void orc3_energy(void* h, double* out) {
  World3& w = *(World3*)h;
  double vene[2] = {0, 0}, efield = 0, bfield = 0;
  for (Rank3& R : w.ranks) {
    for (int isp = 1; isp <= 2; ++isp)
      for (int k = R.nzs; k <= R.nze; ++k)
        for (int j = R.nys; j <= R.nye; ++j)
          for (int ii = 1; ii <= R.np2[R.in2(j, k, isp)]; ++ii) {
            const double* u = &R.up[R.ip(1, ii, j, k, isp)];
            double u2 = u[3] * u[3] + u[4] * u[4] + u[5] * u[5];
            double gam = std::sqrt(1.0 + u2 / (w.c * w.c));
            vene[isp - 1] += w.r[isp - 1] * (gam - 1.0);
          }
    for (int k = R.nzs; k <= R.nze; ++k)
      for (int j = R.nys; j <= R.nye; ++j)
        for (int i = w.nxgs; i <= w.nxge; ++i) {
          const double* f = &R.uf[R.i6(1, i, j, k)];
          bfield += f[0] * f[0] + f[1] * f[1] + f[2] * f[2];
          efield += f[3] * f[3] + f[4] * f[4] + f[5] * f[5];
        }
  }
  out[0] = vene[0]; out[1] = vene[1]; out[2] = efield / (8.0 * kPi); out[3] = bfield / (8.0 * kPi);
}

// Gauss-law residual max|divE - 4 pi rho| over the global periodic grid, and max|4 pi rho|.
// rho(i,j,k) = sum q S(i)S(j)S(k), quadratic spline about the cell centre; divE is the forward
// difference implied by 3d/common/field.f90:171-187.
// which: 0 -> positions from up, 1 -> from gp (using np2 counts).
void orc3_gauss(void* h, int which, double* out) {
  World3& w = *(World3*)h;
  const int nx = w.nx(), ny = w.nyge - w.nygs + 1, nz = w.nzge - w.nzgs + 1;
  std::vector<double> rho((size_t)nx * ny * nz, 0.0), ex(rho.size()), ey(rho.size()), ez(rho.size());
  auto G = [&](int i, int j, int k) {
    i = ((i - w.nxgs) % nx + nx) % nx; j = ((j - w.nygs) % ny + ny) % ny; k = ((k - w.nzgs) % nz + nz) % nz;
    return ((size_t)k * ny + j) * nx + i;
  };
  for (Rank3& R : w.ranks) {
    const std::vector<double>& P = which == 0 ? R.up : R.gp;
    for (int isp = 1; isp <= 2; ++isp)
      for (int k = R.nzs; k <= R.nze; ++k)
        for (int j = R.nys; j <= R.nye; ++j)
          for (int ii = 1; ii <= R.np2[R.in2(j, k, isp)]; ++ii) {
            const double* u = &P[R.ip(1, ii, j, k, isp)];
            int c3[3]; double s[3][3];
            for (int a = 0; a < 3; ++a) {
              c3[a] = (int)std::floor(u[a] * w.d_delx);
              double dh = u[a] * w.d_delx - 0.5 - c3[a];
              s[a][0] = 0.5 * (0.5 - dh) * (0.5 - dh); s[a][1] = 0.75 - dh * dh; s[a][2] = 0.5 * (0.5 + dh) * (0.5 + dh);
            }
            for (int c = -1; c <= 1; ++c)
              for (int b = -1; b <= 1; ++b)
                for (int a = -1; a <= 1; ++a) {
                  if (w.bc != 0 && (c3[0] + a < w.nxgs || c3[0] + a > w.nxge)) continue;
                  rho[G(c3[0] + a, c3[1] + b, c3[2] + c)] += w.q[isp - 1] * s[0][a + 1] * s[1][b + 1] * s[2][c + 1];
                }
          }
    for (int k = R.nzs; k <= R.nze; ++k)
      for (int j = R.nys; j <= R.nye; ++j)
        for (int i = w.nxgs; i <= w.nxge; ++i) {
          ex[G(i, j, k)] = R.uf[R.i6(4, i, j, k)];
          ey[G(i, j, k)] = R.uf[R.i6(5, i, j, k)];
          ez[G(i, j, k)] = R.uf[R.i6(6, i, j, k)];
        }
  }
  double res = 0, mx = 0;
  // walls: only the cells nxs+1 .. nxe-2 that no wall rule touches
  const int i_lo = w.bc == 0 ? w.nxgs : w.nxs + 1, i_hi = w.bc == 0 ? w.nxge : w.nxe - 2;
  for (int k = w.nzgs; k <= w.nzge; ++k)
    for (int j = w.nygs; j <= w.nyge; ++j)
      for (int i = i_lo; i <= i_hi; ++i) {
        double div = ex[G(i + 1, j, k)] - ex[G(i, j, k)] + ey[G(i, j + 1, k)] - ey[G(i, j, k)] + ez[G(i, j, k + 1)] - ez[G(i, j, k)];
        double rr = 4.0 * kPi * w.delx * rho[G(i, j, k)];
        res = std::max(res, std::fabs(div - rr));
        mx = std::max(mx, std::fabs(rr));
      }
  out[0] = res; out[1] = mx;
}

}  // extern "C"
