// hostside/wuming_b200.hpp -- the host side above the C ABI in a compiled language (the reference's host language,
// Fortran 90, cannot be compiled in this image: no f951 / MPI): a header-only C++17 mirror of the reference's module
// procedures on the per-timestep path, same names, same argument meaning, same error behaviour
// ("write(6,*) message; stop" -> message on stderr, exit(1)).  Like the Fortran modules it keeps ONE simulation per
// process in module-level state (3d/common/particle.f90:10-16 `save`d module variables, SURVEY.md 8b "Threading").
//
//   particle__init / field__init / sort__init / bc__init   3d/proj/weibel/app.f90:341-353  -> wuming::init
//   particle__solv        3d/common/particle.f90:52-233         field__fdtd_i   3d/common/field.f90:70-208
//   particle__solv_vay    3d/common/particle.f90:236-419        sort__bucket    3d/common/sort.f90:40-88
//   bc__particle_x / bc__particle_yz / bc__injection            3d/common/boundary_periodic.f90:68-455, boundary_shock.f90
//   mom_calc (accl + nvt + bc__mom)  3d/common/mom_calc.f90     energy_history  3d/proj/weibel/app.f90:509-577
//
// The arrays live on the device between wm_upload and wm_download (WM_SHIM_RESIDENT); the
// procedures therefore take the index range only.  fortran/wuming_b200_shim.f90 is the same mirror in Fortran with the
// reference's full explicit-shape argument lists.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../include/wuming_b200.h"

namespace wuming {

inline wm_ctx*& ctx() {
  static wm_ctx* c = nullptr;
  return c;
}

// the reference's error convention: print the message, stop
inline void check(int ierr, const char* where) {
  if (ierr == WM_OK) return;
  std::fprintf(stderr, "%s: %s\n", where, wm_last_error());
  std::exit(1);
}

// the four *__init calls of init() (3d/proj/weibel/app.f90:341-353) -- their argument lists are the fields of wm_params
inline void init(const wm_params& p) {
  check(wm_create(&p, &ctx()), "particle__init / field__init / sort__init / bc__init");
}
inline void finalize() {
  if (ctx()) wm_destroy(ctx());
  ctx() = nullptr;
}

inline void require_init(const char* msg) {
  if (ctx()) return;
  std::fprintf(stderr, "%s\n", msg);   // e.g. 'Initialize first by calling particle__init()' (particle.f90:69-72)
  std::exit(1);
}

inline void particle__solv(int nxs, int nxe) {
  require_init("Initialize first by calling particle__init()");
  check(wm_particle_solv(ctx(), nxs, nxe), "particle__solv");
}
inline void particle__solv_vay(int nxs, int nxe) {
  require_init("Initialize first by calling particle__init()");
  check(wm_particle_solv_vay(ctx(), nxs, nxe), "particle__solv_vay");
}
inline void field__fdtd_i(int nxs, int nxe) {
  require_init("Initialize first by calling field__init()");
  check(wm_field_fdtd_i(ctx(), nxs, nxe), "field__fdtd_i");
}
inline void bc__particle_x(int nxs, int nxe) {
  require_init("Initialize first by calling boundary__init()");
  check(wm_bc_particle_x(ctx(), nxs, nxe), "bc__particle_x");
}
inline void bc__injection(int nxs, int nxe, double u0) {
  require_init("Initialize first by calling boundary__init()");
  check(wm_bc_injection(ctx(), nxs, nxe, u0), "bc__injection");
}
inline void bc__particle_yz() {
  require_init("Initialize first by calling boundary__init()");
  check(wm_bc_particle_yz(ctx()), "bc__particle_yz");
}
inline void sort__bucket(int nxs, int nxe) {
  require_init("Initialize first by calling sort__init()");
  check(wm_sort_bucket(ctx(), nxs, nxe), "sort__bucket");
}
// mom_calc__accl + mom_calc__nvt + bc__mom (3d/proj/weibel/app.f90:121-124); mom(7, nx+2, nyl+2, nzl+2, nsp)
inline void mom_calc(int nxs, int nxe, std::vector<double>& mom) {
  require_init("Initialize first by calling mom_calc__init()");
  check(wm_mom_calc(ctx(), nxs, nxe, mom.data()), "mom_calc");
}
// the five sums of energy_history (3d/proj/weibel/app.f90:509-577): kinetic per species, E^2/8pi, B^2/8pi, total
inline void energy_history(double it_delt, std::FILE* unit) {
  double e[4];
  check(wm_energy(ctx(), e), "energy_history");
  std::fprintf(unit, "%10.2f %12.5e %12.5e %12.5e %12.5e %12.5e\n", it_delt, e[0], e[1], e[2], e[3], e[0] + e[1] + e[2] + e[3]);
}
// the whole loop body in one call (wm_step: fused push + deposit, lazy sort)
inline void step(int nxs, int nxe, int order, double u0 = 0.0, int nsteps = 1) {
  require_init("Initialize first by calling particle__init()");
  check(wm_step(ctx(), nxs, nxe, order, u0, nsteps), "wm_step");
}

}  // namespace wuming
