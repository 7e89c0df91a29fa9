// wm_push.cuh -- velocity updates shared by the per-procedure push (wm_particles.cu) and the fused kernels (wm_fused.cu).
#pragma once

// particle__solv_vay: Vay pusher (PoP 15, 056701 (2008)), 3d/common/particle.f90:369-406 [2d :264-299].
// f = (bpx,bpy,bpz,epx,epy,epz) gathered at the particle; fac1 = q/r*delt/2, fac2 = q*delt/r (:290-291).
// On return (ux,uy,uz) is the new momentum and igam = 1/gamma_new, the factor of the move x += u*delt*igam.
__device__ __forceinline__ void wm_vay_update(const double f[6], double fac1, double fac2, double c, double& ux, double& uy,
                                              double& uz, double& igam) {
  const double bpx = f[0], bpy = f[1], bpz = f[2], epx = f[3], epy = f[4], epz = f[5];
  const double uvm1 = ux, uvm2 = uy, uvm3 = uz;
  double gam = sqrt(c * c + uvm1 * uvm1 + uvm2 * uvm2 + uvm3 * uvm3);
  const double fac1r = fac1 / gam;
  const double uvm4 = uvm1 + fac2 * epx + fac1r * (+uvm2 * bpz - uvm3 * bpy);
  const double uvm5 = uvm2 + fac2 * epy + fac1r * (+uvm3 * bpx - uvm1 * bpz);
  const double uvm6 = uvm3 + fac2 * epz + fac1r * (+uvm1 * bpy - uvm2 * bpx);
  const double taux = fac1 * bpx / c, tauy = fac1 * bpy / c, tauz = fac1 * bpz / c;
  const double tau2 = taux * taux + tauy * tauy + tauz * tauz;
  const double ua = (uvm4 * taux + uvm5 * tauy + uvm6 * tauz) / c;
  const double sigma = 1.0 + (uvm4 * uvm4 + uvm5 * uvm5 + uvm6 * uvm6) / (c * c) - tau2;
  const double gam2 = 0.5 * (sigma + sqrt(sigma * sigma + 4.0 * (tau2 + ua * ua)));
  gam = sqrt(gam2);
  const double s = 1.0 / (tau2 + gam2);
  ux = s * (gam2 * uvm4 + c * ua * taux + gam * (uvm5 * tauz - uvm6 * tauy));
  uy = s * (gam2 * uvm5 + c * ua * tauy + gam * (uvm6 * taux - uvm4 * tauz));
  uz = s * (gam2 * uvm6 + c * ua * tauz + gam * (uvm4 * tauy - uvm5 * taux));
  igam = 1.0 / gam;
}
