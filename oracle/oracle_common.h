// oracle/oracle_common.h -- TEST INFRASTRUCTURE ONLY.
//
// Shared helpers for the CPU restatement ("oracle") of WumingPIC's per-timestep
// PIC loop.  Nothing under oracle/ is part of the product: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load it, and only as the checker / timed CPU baseline.
//
// PARITY PIN: the reference ships no golden vectors, known-answer tests or
// fixtures for push / deposit / field solve / migration / sort (SURVEY.md §4,
// §8c), and no Fortran compiler or MPI exists in this image, so a gfortran
// build of the reference cannot be run here.  What pins the oracle instead is
// THE REFERENCE'S OWN SOURCE run through another front end: oracle/f2cxx
// translates the 14 hot-path Fortran files mechanically into C++ (oracle/_ref,
// built from /root/reference, never committed) and the oracle equals that
// translated reference BIT FOR BIT -- fields, np2, cumcnt, particle records in
// order, every step, 2-D and 3-D, periodic / reconnection / shock modules,
// both pushers, moments, 1 rank and y, z, y x z rank grids
// (tests/test_ref_transpiled.py; committed vectors tests/golden/ref_cases.npz,
// tests/test_ref_golden.py).  Not covered by that pin: OpenMP / a real MPI
// library's reduction order, and the drivers' loaders (random_number).
// The oracle is a line-by-line restatement of the Fortran loop nests (each
// function cites the file:line it follows) and is additionally checked through
// the physics invariants the algorithm guarantees (Gauss-law residual at
// round-off, particle-count / ID-multiset conservation, Boris |u| conservation,
// N-slab == 1-slab equivalence) and against a second, independent numpy
// restatement of push / deposit / field solve (tests/test_oracle_independent.py).
#pragma once
#include <cstdint>
#include <cmath>
#include <cstddef>

namespace orc {

// ---------------------------------------------------------------------------
// Philox-4x32-10 counter-based RNG (Salmon et al., SC'11; published constants).
// Used by the deterministic synthetic loaders so that the oracle and the GPU
// loader produce the same uniform stream independent of the slab count
// (the reference seeds non-reproducibly: utils/wuming_utils.f90:48-53).
// ---------------------------------------------------------------------------
struct Philox {
  static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
  }
  static inline void run(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(0xD2511F53u, c0, hi0, lo0);
      mulhilo(0xCD9E8D57u, c2, hi1, lo1);
      uint32_t n0 = hi1 ^ c1 ^ k0;
      uint32_t n1 = lo1;
      uint32_t n2 = hi0 ^ c3 ^ k1;
      uint32_t n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
  // two uniforms in [0,1) with 53 random bits each; `epoch` (the time-step counter of the shock injection) is the fourth
  // counter word, 0 for the initial loaders
  static inline void uniform2(uint64_t seed, uint32_t stream, uint32_t idx, uint32_t purpose,
                              double& u0, double& u1, uint32_t epoch = 0u) {
    uint32_t ctr[4] = {idx, purpose, stream, epoch};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t o[4];
    run(ctr, key, o);
    uint64_t a = ((uint64_t)o[1] << 32) | o[0];
    uint64_t b = ((uint64_t)o[3] << 32) | o[2];
    u0 = (double)(a >> 11) * (1.0 / 9007199254740992.0);
    u1 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
  }
};

// Box-Muller pair in the reference's form (utils/wuming_utils.f90:72-90):
//   rr = sqrt(-2*log(1-x1) + 1e-30); (rr*sin(2 pi x2), rr*cos(2 pi x2))
static inline void box_muller(double x1, double x2, double& n_sin, double& n_cos) {
  const double pi = 4.0 * std::atan(1.0);
  double rr = std::sqrt(-2.0 * std::log(1.0 - x1) + 1.0e-30);
  n_sin = rr * std::sin(2.0 * pi * x2);
  n_cos = rr * std::cos(2.0 * pi * x2);
}

// Parameters of the shock driver's inject() / relocate() (2d/proj/shock/app.f90:615-852, 3d :644-906) and its velocity
// profile vprofile (2d :883-893, 3d :939-949).
struct ShockPrm {
  int n0;
  double v0, v_thi, v_the, b0, theta_bn, phi_bn, l_damp_ini;
  uint64_t seed;
};
static inline double vprofile(const ShockPrm& s, double x, int nxgs, double delx) {
  const double x0 = s.l_damp_ini + nxgs * delx;
  const double xs = s.l_damp_ini * 0.1;
  return 0.5 * s.v0 * (1 + std::tanh((x - x0) / xs));
}
// Maxwellian in the fluid rest frame + Lorentz transform to the lab frame (2d/proj/shock/app.f90:811-826).  The reference
// draws from Fortran's random_number (not reproducible, utils/wuming_utils.f90:48-53); here the three normal deviates of
// particle (stream = global row, idx = ii, species isp) at step `epoch` are Box-Muller pairs of Philox uniforms with
// purpose base + 2 isp - 1 (-> ux, uy) and base + 2 isp (-> uz), the same convention as the Weibel loader.
static inline void shock_velocity(const ShockPrm& s, uint32_t row, uint32_t ii, int isp, uint32_t base, uint32_t epoch, double c,
                                  double u[3]) {
  const double sd = isp == 1 ? s.v_thi : s.v_the;
  double a0, a1, b0, b1, ns, nc, ms, mc;
  Philox::uniform2(s.seed, row, ii, base + (uint32_t)(2 * isp - 1), a0, a1, epoch);
  Philox::uniform2(s.seed, row, ii, base + (uint32_t)(2 * isp), b0, b1, epoch);
  box_muller(a0, a1, ns, nc);
  box_muller(b0, b1, ms, mc);
  u[0] = sd * ns; u[1] = sd * nc; u[2] = sd * ms;
  (void)mc; (void)c;
}

// start/end of a 1-D block decomposition (3d/common/mpi_set.f90:81-94, para_range)
static inline void para_range(int& ns, int& ne, int n1, int n2, int isize, int irank) {
  int iwork1 = (n2 - n1 + 1) / isize;
  int iwork2 = (n2 - n1 + 1) % isize;
  ns = irank * iwork1 + n1 + (irank < iwork2 ? irank : iwork2);
  ne = ns + iwork1 - 1;
  if (iwork2 > irank) ne = ne + 1;
}

}  // namespace orc
