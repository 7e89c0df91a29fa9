// hostside/weibel3d_main.cpp -- app__main of the 3-D Weibel set-up (3d/proj/weibel/app.f90:88-160) on the C++ host-side
// mirror: init, the five library calls per step, energy_history at the moment cadence.  The configuration constants are
// those of 3d/proj/weibel/config_sample.json's meaning (n_ppc, v_the, v_thi, t_ani, omega_pe, mass_ratio = 1) with
// c = delx = delt = 1 and gfac = 0.501 (app.f90:35-41); the initial load is the device-side Weibel loader
// (wm_load_weibel = app.f90:391-504 with Philox streams), output is the energy history on stdout.
//
//   weibel3d_main [nx ny nz n_ppc max_it intvl_mom [fused]]      fused = 1: the loop body as ONE wm_step call per step
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "wuming_b200.hpp"

int main(int argc, char** argv) {
  const int nx = argc > 1 ? std::atoi(argv[1]) : 32, ny = argc > 2 ? std::atoi(argv[2]) : 16, nz = argc > 3 ? std::atoi(argv[3]) : 8;
  const int n_ppc = argc > 4 ? std::atoi(argv[4]) : 8, max_it = argc > 5 ? std::atoi(argv[5]) : 10;
  const int intvl_mom = argc > 6 ? std::atoi(argv[6]) : 5;
  const bool fused = argc > 7 && std::atoi(argv[7]) != 0;
  const double pi = 4.0 * std::atan(1.0);
  const double c = 1.0, delx = 1.0, delt = 1.0, gfac = 0.501;
  const double omega_pe = 0.1, mass_ratio = 1.0, v_the = 0.1, v_thi = 0.1, t_ani = 5.0;

  wm_params p = {};
  p.dim = 3; p.ndim = 7; p.nsp = 2;
  p.nxgs = 2; p.nxge = nx + 1; p.nygs = 2; p.nyge = ny + 1; p.nzgs = 2; p.nzge = nz + 1;     // app.f90:33-34: indices start at 2
  p.nys = p.nygs; p.nye = p.nyge; p.nzs = p.nzgs; p.nze = p.nzge;                             // one rank
  p.np = 3 * n_ppc * nx;                                                                      // app.f90:285: pencil capacity
  p.nproc_j = p.nproc_k = 1; p.rank_j = p.rank_k = 0;
  p.bc_kind = WM_BC_PERIODIC; p.device = -1;
  p.delx = delx; p.delt = delt; p.c = c; p.gfac = gfac;
  // app.f90:298-305: r(1) = mass_ratio * r(2), q = +-sqrt(r / (4 pi n0)) * omega_p, n0 particles per cell
  p.r[0] = mass_ratio; p.r[1] = 1.0;
  const double wpe = omega_pe, wpi = wpe / std::sqrt(mass_ratio);
  p.q[0] = +std::sqrt(p.r[0] / (4.0 * pi * n_ppc)) * wpi;
  p.q[1] = -std::sqrt(p.r[1] / (4.0 * pi * n_ppc)) * wpe;
  wuming::init(p);
  wuming::check(wm_load_weibel(wuming::ctx(), n_ppc, v_thi, v_the, t_ani, 0.0, 20240601ull), "init");
  const int nxs = p.nxgs, nxe = p.nxge;
  wuming::energy_history(0.0, stdout);

  for (int it = 1; it <= max_it; ++it) {
    if (fused) {
      wuming::step(nxs, nxe, WM_ORDER_WEIBEL);
    } else {
      wuming::particle__solv(nxs, nxe);        // app.f90:102-108
      wuming::field__fdtd_i(nxs, nxe);
      wuming::bc__particle_x(nxs, nxe);
      wuming::bc__particle_yz();
      wuming::sort__bucket(nxs, nxe);
    }
    if (it % intvl_mom == 0) wuming::energy_history(it * delt, stdout);   // app.f90:120-126
  }
  wm_stats st;
  wuming::check(wm_get_stats(wuming::ctx(), &st), "wm_get_stats");
  std::printf("# %d steps, %lld particles, cg iterations of the last step %d %d %d\n", max_it, st.n_particles, st.cg_iterations[0],
              st.cg_iterations[1], st.cg_iterations[2]);
  wuming::finalize();
  return 0;
}
