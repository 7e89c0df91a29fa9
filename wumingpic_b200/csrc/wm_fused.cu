// wm_fused.cu -- the benchmark hot path: ONE kernel for K1-gather + Boris push + x/y/z boundary +
// Esirkepov deposit + destination-cell counting, followed by a deterministic stable counting sort.
//
//   particle__solv   3d/common/particle.f90:93-225      ele_cur  3d/common/field.f90:238-404
//   bc__particle_x   3d/common/boundary_periodic.f90:68-101
//   bc__particle_yz  :104-455 (coordinate wrap + re-binning)    sort__bucket  3d/common/sort.f90:55-86
//
// Why this shape (measured on B200):
//   * fp64 atomicAdd on shared memory is an ATOMS.CAS spin loop on sm_100a and warp shuffles issue at one
//     warp-instruction per clock per SM, so neither smem atomics nor shuffle reductions can carry the
//     ~300 current contributions per particle.  Instead the deposit is turned inside out:
//       phase A  one PARTICLE per thread: gather (field tile in smem, loaded by TMA bulk copies), Boris,
//                move, boundary wrap, and the 1-D Esirkepov factors (prefix sums c, pairs (S0,DS)) written
//                to a shared-memory record;
//       phase B  one OUTPUT STRIP per thread: (cell, component, transverse index) owns 4x5 current values
//                in REGISTERS and walks the records of its cell: acc[r][kp] += c[r] * (A*S0[kp] + B*DS[kp]).
//     No atomics and no shuffles on the per-particle path; each strip ends with <= 20 RED.F64 to global J.
//     Records are written two-ended inside a batch: particles that stay in their cell (inc = 0 on every axis, the
//     majority) from slot 0 up, cell-crossers from slot 15 down.  A stayer has S0 = DS = 0 on the two outer stencil
//     points of every axis and c0 = 0, c3 = O(eps), so phase B runs a lean 2x3 loop over the stayer slots (15 fp64
//     instructions per particle and strip instead of 31) and the full 4x5 loop only over the crossers.
//   * the next batch's particle loads are issued before the current batch is processed (register prefetch), so the
//     global-load latency hides behind ~1.5 k instructions of phase A/B work instead of stalling 4 warps per scheduler.
//   * a CTA owns G = 8 consecutive x cells of one (j,k) pencil (16 lanes per cell): their particles are one contiguous run
//     of the cell-sorted SoA (coalesced 128-byte loads per half-warp) and share a 10x3x3 field tile, loaded by 9 TMA bulk
//     row copies, plus a second tile of its x differences (the x sum of the gather is f(i) + sx(+1) d(i) - sx(-1) d(i-1)).
//   * lazy sort (wm_sort.cu): between two steps the sort's permutation stays pending and this kernel reads its particle
//     through inv[sorted position] (a two-stage prefetch requests the index one batch before the particle), so the
//     particle store is read and written once per step.
//   * the kernel also counts, per source cell, how many particles go to each of the 27 neighbour cells.
//     |dx| < c dt <= 1 cell, so that count matrix is all the sort needs: destination offsets follow from a
//     per-cell prefix over the 27 sources and ranks from the (stable) order inside the source cell, which
//     makes the scatter atomic-free and the particle order -- hence every later sum -- deterministic.
#include "wm_cells.cuh"
#include "wm_push.cuh"

#include <cstdlib>

namespace {

// G = cells per CTA (template parameter): 16 threads per cell, so a CTA has 16*G threads.  G = 8 gives three
// 128-thread CTAs per SM (168 registers, 57.5 KB of shared memory each) in different phases of their batch loop.
constexpr int SLOTS = 16;     // particles per cell and batch
constexpr int NF = 21;        // double2 fields per particle record
#ifndef WM_CSTR
#define WM_CSTR 17
#endif
constexpr int CSTR = WM_CSTR;      // slot stride between cells (17: bank spreading between the two cells of a warp)

template <int G>
struct __align__(16) Smem {
  static constexpr int FSTR = G * CSTR + 1;   // double2 stride between record fields (== 1 mod 8)
  static constexpr int TILE_X = G + 2;
  static constexpr int TILE_ROW = TILE_X * 6; // doubles per (jj,kk) row of the field tile
  double tile[9 * TILE_ROW];          // tmpf for cells i0-1..i0+G, j-1..j+1, k-1..k+1
  double dtile[9 * TILE_ROW];         // x differences of the tile: dtile[row][x] = tile[row][x+1] - tile[row][x]
  double2 rec[NF * FSTR];             // per-particle deposit factors, field-major
  int beg[2][G + 1];                  // cs row segments of both species
  int cnt27[2][27][G];                // destination-offset counts per (species, offset, cell)
  unsigned long long bar;             // mbarrier for the TMA bulk copies
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}

__device__ __forceinline__ void shape3(double dh, double& sm, double& s0, double& sp) {
  sm = 5e-1 * (5e-1 - dh) * (5e-1 - dh);
  s0 = 7.5e-1 - dh * dh;
  sp = 5e-1 * (5e-1 + dh) * (5e-1 + dh);
}

// S0 (about the loop cell) and DS = S1 - S0 (S1 about int(x_new) = cell + inc) on the 5-point stencil, field.f90:255-325
__device__ __forceinline__ void s0ds(double xo, double xn, int cell, int inc, double d_delx, double s0[5], double ds[5]) {
  double dh = xo * d_delx - 5e-1 - cell;
  s0[0] = 0.0;
  shape3(dh, s0[1], s0[2], s0[3]);
  s0[4] = 0.0;
  dh = xn * d_delx - 5e-1 - (cell + inc);
  double a, b, c;
  shape3(dh, a, b, c);
  ds[0] = inc == -1 ? a : 0.0;
  ds[1] = inc == -1 ? b : (inc == 0 ? a : 0.0);
  ds[2] = inc == -1 ? c : (inc == 0 ? b : a);
  ds[3] = inc == 0 ? c : (inc == 1 ? b : 0.0);
  ds[4] = inc == 1 ? c : 0.0;
#pragma unroll
  for (int m = 0; m < 5; ++m) ds[m] = ds[m] - s0[m];
}

// ---------------------------------------------------------------------------------------------
// fused push + boundary + deposit + destination counting (3-D)
//   ORDER 0 (Weibel/beam): deposit sees the un-wrapped new position, the periodic x wrap follows (3d/proj/weibel/app.f90:100-108)
//   ORDER 1 (reconnection): reflecting walls act on the pushed particle BEFORE the deposit (3d/proj/reconnection/app.f90:103-108,
//            boundary_reconnection.f90:69-110); ORDER 2 (shock): boundary_shock__injection likewise (boundary_shock.f90:424-469)
// ---------------------------------------------------------------------------------------------
#ifndef WM_FUSED_MINB
#define WM_FUSED_MINB 3   // 3 x 128-thread CTAs per SM: 168 registers, no spills (4 CTAs at 128 registers: 3 % slower)
#endif
template <int ORDER, int G, bool VAY>
__global__ void __launch_bounds__(16 * G, WM_FUSED_MINB)
k_fused3(Geo g, Ptcl A, Ptcl B, const double* __restrict__ id_in, double* __restrict__ id_out,
         const int* __restrict__ inv, Ptcl R, const double* __restrict__ rid, const int* __restrict__ cs, const double* __restrict__ tmpf, double* __restrict__ uj, int* __restrict__ cnt,
         int* __restrict__ hist, unsigned char* __restrict__ dst_off, int* flags, int nxs, int nxe, int ngx, double u0,
         double xend) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using SM = Smem<G>;
  constexpr int TPB = 16 * G, FSTR = SM::FSTR, TILE_ROW = SM::TILE_ROW;
  SM& S = *reinterpret_cast<SM*>(smem_raw);
  const int t = threadIdx.x;
  // group -> (i0, j, k)
  // block order: j-strips of 8 rows, then k, then the rows of the strip, x groups fastest -- the order wm_sort.cu
  // walks too, so that a CTA's neighbours in y and z run within a few MB of it (J RED traffic stays in L2)
  const int gx = blockIdx.x % ngx;
  int j, k;
  wm_strip_pencil(g, blockIdx.x / ngx, j, k);
  const int i0 = nxs + gx * G;
  const int ncg = min(G, nxe - i0 + 1);  // cells in this group

  // prologue order: the tile (TMA), the index rows and the first particles are three independent global round trips --
  // all three are in flight before anything waits (the tile is waited for after the first particle loads were issued)
  if (t == 0) {
    mbar_init(&S.bar, 1);
    // field tile by TMA bulk copies: 9 rows of (ncg+2) cells x 6 doubles (48 B per cell keeps 16-B alignment)
    const uint32_t row_bytes = (uint32_t)(ncg + 2) * 48u;
    mbar_expect_tx(&S.bar, 9u * row_bytes);
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int jj = r % 3 - 1, kk = r / 3 - 1;
      tma_bulk_g2s(&S.tile[r * TILE_ROW], tmpf + g.box(i0 - 1, j + jj, k + kk) * 6, row_bytes, &S.bar);
    }
  }
  if (t < 2 * (G + 1)) {
    const int isp = t / (G + 1), ii = t % (G + 1);
    const int* row = cs + (size_t)g.pen(j, k, isp) * (g.nx + 1) + (i0 - g.nxgs);
    S.beg[isp][ii] = row[min(ii, ncg)];
  }
  for (int e = t; e < 2 * 27 * G; e += TPB) (&S.cnt27[0][0][0])[e] = 0;
  __syncthreads();

  // phase-A identity: (cell, slot); phase-B identity: (cell, component, transverse index m)
  const int ca = t >> 4, sa = t & 15;
  const int cb = t >> 4, hb = t & 15;
  const int comp = hb / 5, mb = hb % 5;
  const bool b_active = hb < 15 && cb < ncg;
  const int n0a = ca < ncg ? S.beg[0][ca + 1] - S.beg[0][ca] : 0;
  const int n1a = ca < ncg ? S.beg[1][ca + 1] - S.beg[1][ca] : 0;
  const int ncb = cb < ncg ? (S.beg[0][cb + 1] - S.beg[0][cb]) + (S.beg[1][cb + 1] - S.beg[1][cb]) : 0;
  // A cell's records are written (phase A) and consumed (phase B) by the same half-warp, so the two phases only need
  // warp-level synchronisation: every warp walks the batches of its own two cells at its own pace, and the phase-A
  // load latency of one warp overlaps the phase-B math of the others.
  const int nbatch = (max(ncb, __shfl_xor_sync(0xffffffffu, ncb, 16)) + SLOTS - 1) / SLOTS;

  double acc[20];
#pragma unroll
  for (int e = 0; e < 20; ++e) acc[e] = 0.0;

  // record field indices: 2a = (c1,c2), 2a+1 = (c0,c3) prefix sums of axis a; 6+5a+m = (S0,DS)[m] of axis a
  const int f_c = 2 * comp;                       // (c1,c2) of the running axis; f_c + 1 = (c0,c3)
  const int ax1 = comp == 0 ? 1 : 0;              // first transverse axis (A,B): y for Jx, x for Jy and Jz
  const int ax2 = comp == 2 ? 1 : 2;              // second transverse axis: z for Jx and Jy, y for Jz
  const double fac = 1.0 / 3.0;
  const double inv_c2 = 1.0 / (g.c * g.c);
  const int ia = i0 + ca;
  const double len_x = (g.nxge - g.nxgs + 1) * g.delx;
  const double len_y = (g.nyge - g.nygs + 1) * g.delx;
  const double len_z = (g.nzge - g.nzgs + 1) * g.delx;

  int ns0 = 0, nl0 = 0, ns1 = 0, nl1 = 0;   // stayers / leavers of this thread's cell written so far, per species
  // register prefetch of the next batch's particle (index -1: none)
  double nx_ = 0, ny_ = 0, nz_ = 0, nux = 0, nuy = 0, nuz = 0, nid = 0;
  int np_ = -1, nisp = 0;
  // two-stage prefetch: the index of batch b+2 (the sorted position, or -- lazy sort -- its image under the pending
  // permutation) is requested while the particle of batch b+1 is loaded and batch b is processed, so neither load
  // stalls the warp.  qi: where the next particle sits (>= 0: set A, <= -2: arrival store of a slab run, -1: none)
  int qi = -1, qisp = 0;
  auto prep = [&](int batch) {
    const int idx = batch * SLOTS + sa;
    int pos = -1;
    qi = -1; qisp = 0;
    if (idx < n0a) { pos = S.beg[0][ca] + idx; }
    else if (idx < n0a + n1a) { pos = S.beg[1][ca] + (idx - n0a); qisp = 1; }
    if (pos >= 0) qi = inv ? inv[pos] : pos;
  };
  auto fetch = [&]() {
    np_ = qi == -1 ? -1 : 0; nisp = qisp;
    if (qi >= 0) {
      nx_ = A.c[0][qi]; ny_ = A.c[1][qi]; nz_ = A.c[2][qi];
      nux = A.c[3][qi]; nuy = A.c[4][qi]; nuz = A.c[5][qi];
      nid = id_in[qi];
    } else if (qi <= -2) {
      const int a = -2 - qi;
      nx_ = R.c[0][a]; ny_ = R.c[1][a]; nz_ = R.c[2][a];
      nux = R.c[3][a]; nuy = R.c[4][a]; nuz = R.c[5][a];
      nid = rid[a];
    }
  };
  prep(0);
  fetch();
  prep(1);

  mbar_wait(&S.bar, 0);
  // x differences of the field tile, once per CTA (shared by the ~128 particles of every cell): the x sum of the gather
  // becomes f(i) + sx(+1) d(i) - sx(-1) d(i-1), two DFMA instead of DMUL + two DFMA (sx(-1) + sx(0) + sx(+1) = 1)
  for (int e = t; e < 9 * (G + 1) * 6; e += TPB) {
    const int r = e / ((G + 1) * 6), q = e % ((G + 1) * 6);
    S.dtile[r * TILE_ROW + q] = S.tile[r * TILE_ROW + q + 6] - S.tile[r * TILE_ROW + q];
  }
  __syncthreads();

  for (int batch = 0; batch < nbatch; ++batch) {
    int nst = 0, ncr = 0;   // stayer / crosser records of this half-warp's cell in this batch
    // ------------------------------ phase A ------------------------------
    {
      const int p = np_, isp = nisp;
      const double x = nx_, y = ny_, z = nz_;
      double ux = nux, uy = nuy, uz = nuz;
      const double idv = nid;
      if (batch + 1 < nbatch) { fetch(); prep(batch + 2); }
      double xn = 0, yn = 0, zn = 0;
      int o = 13;
      int inc0 = 0, inc1 = 0, inc2 = 0;
      if (p >= 0) {
        double sx[3], sy[3], sz[3];
        shape3(x * g.d_delx - 5e-1 - ia, sx[0], sx[1], sx[2]);
        shape3(y * g.d_delx - 5e-1 - j, sy[0], sy[1], sy[2]);
        shape3(z * g.d_delx - 5e-1 - k, sz[0], sz[1], sz[2]);
        // gather from the smem tile, reference nesting (x-sum, *shy, *shz)  particle.f90:126-184; the x sum in the
        // difference form above (same value up to one rounding of 1e-16 |f|)
        const double nsx0 = -sx[0];
        double f[6];
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
          double pl[6];
#pragma unroll
          for (int jj = 0; jj < 3; ++jj) {
            const double2* tc = reinterpret_cast<const double2*>(&S.tile[(kk * 3 + jj) * TILE_ROW + (ca + 1) * 6]);
            const double2* td = reinterpret_cast<const double2*>(&S.dtile[(kk * 3 + jj) * TILE_ROW + ca * 6]);
            double v[6], d[12];
#pragma unroll
            for (int e = 0; e < 3; ++e) { double2 q2 = tc[e]; v[2 * e] = q2.x; v[2 * e + 1] = q2.y; }
#pragma unroll
            for (int e = 0; e < 6; ++e) { double2 q2 = td[e]; d[2 * e] = q2.x; d[2 * e + 1] = q2.y; }
#pragma unroll
            for (int c = 0; c < 6; ++c) {
              double row = fma(sx[2], d[6 + c], fma(nsx0, d[c], v[c]));
              pl[c] = jj == 0 ? row * sy[0] : pl[c] + row * sy[jj];
            }
          }
#pragma unroll
          for (int c = 0; c < 6; ++c) f[c] = kk == 0 ? pl[c] * sz[0] : f[c] + pl[c] * sz[kk];
        }
        // Buneman-Boris  particle.f90:186-217
        const double fac1 = g.fac1[isp];   // per-species constants computed once on the host (wm_create)
        const double txxx = fac1 * fac1;
        const double fac2 = g.fac2[isp];
        if (VAY) {
          // particle__solv_vay  particle.f90:369-406
          double igam;
          wm_vay_update(f, fac1, fac2, g.c, ux, uy, uz, igam);
          xn = x + ux * g.delt * igam;
          yn = y + uy * g.delt * igam;
          zn = z + uz * g.delt * igam;
        } else {
          const double bpx = f[0], bpy = f[1], bpz = f[2], epx = f[3], epy = f[4], epz = f[5];
          double uvm1 = ux + fac1 * epx, uvm2 = uy + fac1 * epy, uvm3 = uz + fac1 * epz;
          // gam = sqrt(q), igam = 1/gam through one rsqrt (two roundings instead of a correctly rounded sqrt and a
          // division: ~1e-16 relative, far inside the 1e-13 push tolerance; saves ~20 fp64 issue slots per particle)
          const double qg = g.c * g.c + uvm1 * uvm1 + uvm2 * uvm2 + uvm3 * uvm3;
          double igam = rsqrt(qg);
          double gam = qg * igam;
          double fac1r = fac1 * igam;
          // reciprocal + multiply instead of a division (one more rounding, 1e-16 relative)
          double fac2r = fac2 * __drcp_rn(gam + txxx * (bpx * bpx + bpy * bpy + bpz * bpz) * igam);
          double uvm4 = uvm1 + fac1r * (+uvm2 * bpz - uvm3 * bpy);
          double uvm5 = uvm2 + fac1r * (+uvm3 * bpx - uvm1 * bpz);
          double uvm6 = uvm3 + fac1r * (+uvm1 * bpy - uvm2 * bpx);
          uvm1 = uvm1 + fac2r * (+uvm5 * bpz - uvm6 * bpy);
          uvm2 = uvm2 + fac2r * (+uvm6 * bpx - uvm4 * bpz);
          uvm3 = uvm3 + fac2r * (+uvm4 * bpy - uvm5 * bpx);
          ux = uvm1 + fac1 * epx;
          uy = uvm2 + fac1 * epy;
          uz = uvm3 + fac1 * epz;
          gam = rsqrt(1.0 + (+ux * ux + uy * uy + uz * uz) * inv_c2);
          xn = x + ux * g.delt * gam;
          yn = y + uy * g.delt * gam;
          zn = z + uz * g.delt * gam;
        }
        if (ORDER != 0) {
          // x walls on the pushed particle, before the deposit sees it
          const int ipos = ORDER == 2 ? (int)(xn * g.d_delx) : (int)(xn / g.delx);
          if (ipos < nxs + 1) {
            xn = 2.0 * (nxs + 1) * g.delx - xn;
            ux = -ux; uy = -uy; uz = -uz;
          } else if (ORDER == 1 ? ipos >= nxe - 1 : xn > xend) {
            if (ORDER == 1) { xn = 2.0 * (nxe - 1) * g.delx - xn; ux = -ux; }
            else { xn = 2.0 * xend - xn; ux = 2.0 * u0 - ux; }
            uy = -uy; uz = -uz;
          }
        }
        // cell increments (field.f90:280-283): the destination offset of the sort and the stayer / crosser class
        inc0 = (int)(xn * g.d_delx) - ia;
        inc1 = (int)(yn * g.d_delx) - j;
        inc2 = (int)(zn * g.d_delx) - k;
        if (inc0 < -1 || inc0 > 1 || inc1 < -1 || inc1 > 1 || inc2 < -1 || inc2 > 1) {
          atomicOr(flags, 2);
          inc0 = max(-1, min(1, inc0)); inc1 = max(-1, min(1, inc1)); inc2 = max(-1, min(1, inc2));
        }
        o = (inc0 + 1) + 3 * (inc1 + 1) + 9 * (inc2 + 1);
        atomicAdd(&S.cnt27[isp][o][ca], 1);
      }
      // The 16 lanes of a half-warp share the cell; ranks come from ballots.  Stayers take record slots 0.. and the
      // front of the cell's run in the pushed set, leavers record slots 15.. downwards and the back of the run (what
      // wm_sort.cu's gather expects).
      const unsigned lane = t & 31u;
      const unsigned half = 0xffffu << (lane & 16u);
      const unsigned lower = half & ((1u << lane) - 1u);
      const bool v0 = p >= 0 && isp == 0, v1 = p >= 0 && isp == 1, st = o == 13;
      const unsigned bs0 = __ballot_sync(0xffffffffu, v0 && st) & half, bl0 = __ballot_sync(0xffffffffu, v0 && !st) & half;
      const unsigned bs1 = __ballot_sync(0xffffffffu, v1 && st) & half, bl1 = __ballot_sync(0xffffffffu, v1 && !st) & half;
      nst = __popc(bs0 | bs1);
      ncr = __popc(bl0 | bl1);
      if (p >= 0) {
        // Esirkepov 1-D factors -> record
        const double qdxdt = g.qdxdt[isp];
        const int slot = ca * CSTR + (st ? __popc((bs0 | bs1) & lower) : SLOTS - 1 - __popc((bl0 | bl1) & lower));
        const double xo[3] = {x, y, z}, xnw[3] = {xn, yn, zn};
        const int cell[3] = {ia, j, k};
        const int inc[3] = {inc0, inc1, inc2};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          double s0[5], ds[5];
          s0ds(xo[a], xnw[a], cell[a], inc[a], g.d_delx, s0, ds);
          const double c0 = -ds[0] * qdxdt;
          const double c1 = c0 - ds[1] * qdxdt;
          const double c2 = c1 - ds[2] * qdxdt;
          const double c3 = c2 - ds[3] * qdxdt;
          S.rec[(2 * a) * FSTR + slot] = make_double2(c1, c2);
          S.rec[(2 * a + 1) * FSTR + slot] = make_double2(c0, c3);
#pragma unroll
          for (int m = 0; m < 5; ++m) S.rec[(6 + 5 * a + m) * FSTR + slot] = make_double2(s0[m], ds[m]);
        }
        // boundaries: periodic x (boundary_periodic.f90:86-92) and periodic y,z wrap of the coordinate (:161-171).
        // The destination cell is fixed by the pre-wrap integer cell, as in the reference.
        {
          if (ORDER == 0) {
            int ipos = (int)(xn * g.d_delx);
            if (ipos < g.nxgs) xn = xn + len_x;
            else if (ipos >= g.nxge + 1) xn = xn - len_x;
          }
          int jpos = (int)(yn * g.d_delx);
          if (jpos <= g.nygs - 1) yn = yn + len_y;
          else if (jpos >= g.nyge + 1) yn = yn - len_y;
          int kpos = (int)(zn * g.d_delx);
          if (kpos <= g.nzgs - 1) zn = zn + len_z;
          else if (kpos >= g.nzge + 1) zn = zn - len_z;
        }
        int pw;
        if (isp == 0) pw = st ? S.beg[0][ca] + ns0 + __popc(bs0 & lower) : S.beg[0][ca + 1] - 1 - (nl0 + __popc(bl0 & lower));
        else          pw = st ? S.beg[1][ca] + ns1 + __popc(bs1 & lower) : S.beg[1][ca + 1] - 1 - (nl1 + __popc(bl1 & lower));
        B.c[0][pw] = xn; B.c[1][pw] = yn; B.c[2][pw] = zn;
        B.c[3][pw] = ux; B.c[4][pw] = uy; B.c[5][pw] = uz;
        id_out[pw] = idv;
        if (!st) dst_off[pw] = (unsigned char)o;
      }
      ns0 += __popc(bs0); nl0 += __popc(bl0); ns1 += __popc(bs1); nl1 += __popc(bl1);
    }
    __syncwarp();
    // ------------------------------ phase B ------------------------------
    if (b_active) {
      // stayers: S0 = DS = 0 at m = 0, 4 on every axis, c0 = 0 and c3 = -(sum of DS) q = O(eps) q (dropped: it is the
      // round-off residue of a term that is exactly zero, field.f90:341-349).  Two slots per trip: their load -> D ->
      // accumulate chains are independent, which halves the fixed-latency stalls of the 4 warps per scheduler.
      const double2* rc = S.rec + f_c * FSTR + cb * CSTR;
      const double2* r1 = S.rec + (6 + 5 * ax1 + mb) * FSTR + cb * CSTR;
      const double2* r2 = S.rec + (6 + 5 * ax2) * FSTR + cb * CSTR;
      auto stay = [&](int s) {
        const double2 c12 = rc[s];
        const double2 p1 = r1[s];
        const double Av = p1.x + 5e-1 * p1.y;
        const double Bv = 5e-1 * p1.x + fac * p1.y;
#pragma unroll
        for (int kp = 1; kp < 4; ++kp) {
          const double2 p2 = r2[kp * FSTR + s];
          const double D = Av * p2.x + Bv * p2.y;
          acc[1 * 5 + kp] += c12.x * D;
          acc[2 * 5 + kp] += c12.y * D;
        }
      };
      auto cross = [&](int s) {
        const double2 c12 = rc[s];
        const double2 c03 = rc[FSTR + s];
        const double2 p1 = r1[s];
        const double Av = p1.x + 5e-1 * p1.y;
        const double Bv = 5e-1 * p1.x + fac * p1.y;
#pragma unroll
        for (int kp = 0; kp < 5; ++kp) {
          const double2 p2 = r2[kp * FSTR + s];
          const double D = (kp == 0 || kp == 4) ? Bv * p2.y : Av * p2.x + Bv * p2.y;
          acc[0 * 5 + kp] += c03.x * D;
          acc[1 * 5 + kp] += c12.x * D;
          acc[2 * 5 + kp] += c12.y * D;
          acc[3 * 5 + kp] += c03.y * D;
        }
      };
      int s = 0;
      if (mb >= 1 && mb <= 3) {   // a stayer's S0 = DS = 0 at the outer points: the strips mb = 0, 4 receive exactly nothing
        for (; s + 1 < nst; s += 2) { stay(s); stay(s + 1); }
        if (s < nst) stay(s);
      }
      s = SLOTS - ncr;
      for (; s + 1 < SLOTS; s += 2) { cross(s); cross(s + 1); }
      if (s < SLOTS) cross(s);
    }
    __syncwarp();
  }

  // ---- flush: current strips -> global J (RED.F64, zeros skipped), counts -> cnt27 -----------------
  if (b_active) {
    const int ib = i0 + cb;
    const long long sY = (long long)g.bx * 3, sZ = (long long)g.bx * g.by * 3;
    // strides of the strip's three indices (running r, first transverse mb, second transverse kp) per component:
    //   Jx(i+r-1, j+mb-2, k+kp-2)   Jy(i+mb-2, j+r-1, k+kp-2)   Jz(i+mb-2, j+kp-2, k+r-1)
    const long long s_r = comp == 0 ? 3 : (comp == 1 ? sY : sZ);
    const long long s_m = comp == 0 ? sY : 3;
    const long long s_k = comp == 2 ? sY : sZ;
    double* J0 = uj + g.box(ib, j, k) * 3 + comp - s_r + (mb - 2) * s_m - 2 * s_k;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int kp = 0; kp < 5; ++kp) {
        const double v = acc[r * 5 + kp];
        if (v != 0.0) atomicAdd(J0 + r * s_r + kp * s_k, v);
      }
  }
  // ---- re-binning information for the sort: one count line per (cell, species), group sizes -> histogram ----
  // A cell's counts were accumulated by its own half-warp only, so every warp flushes its two cells and leaves: no CTA
  // barrier at the tail (the slowest warp of a CTA no longer holds the other three).
  __syncwarp();
  if (ca < ncg) {
    const size_t cell = wm_cell_index(g, i0, j, k) + ca;
    for (int e = sa; e < 2 * WM_CNT_LINE; e += 16) {
      const int o = e % WM_CNT_LINE, isp = e / WM_CNT_LINE;
      cnt[cell * 2 * WM_CNT_LINE + e] = o < 27 ? S.cnt27[isp][o][ca] : 0;
    }
    for (int e = sa; e < 2 * 27; e += 16) {
      const int o = e % 27, isp = e / 27;
      const int n = S.cnt27[isp][o][ca];
      if (n > 0) {
        int drow, ti;
        if (wm_dest_of(g, i0 + ca, j, k, o, isp, nxs, nxe, drow, ti)) atomicAdd(hist + (size_t)drow * (g.nx + 1) + (ti - g.nxgs), n);
        else atomicOr(flags, 2);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// fused push + boundary + deposit + destination counting (2-D)
//   particle__solv 2d/common/particle.f90:85-171, ele_cur 2d/common/field.f90:209-314 (Jx, Jy: 2-D Esirkepov prefix sums;
//   Jz = q vz (S0x S0y + 1/2 DSx S0y + 1/2 S0x DSy + 1/3 DSx DSy) with vz of the pushed particle, :262-264, 288-294),
//   boundary_periodic__particle_x / __particle_y (2d/common/boundary_periodic.f90:61-96, 99-248: int(x/delx) and the
//   wraps run under ieee_down -> __ddiv_rd / __dadd_rd), walls as in the 3-D kernel (2d/proj/reconnection/
//   boundary_reconnection.f90:61-99, 2d/proj/shock/boundary_shock.f90:255-297).
// Same structure as k_fused3: a CTA owns G x-cells of one row j, 16 lanes per cell; record fields (double2):
//   0 (cx1,cx2) 1 (cx0,cx3) 2 (cy1,cy2) 3 (cy0,cy3) 4+m (S0x,DSx)[m] 9+m (S0y,DSy)[m] 14 (q vz, -)
// phase-B lanes: (Jx, jp) own Jx(i-1..i+2, j+jp-2); (Jy, ip) own Jy(i+ip-2, j-1..j+2); (Jz, ip) own Jz(i+ip-2, j-2..j+2).
// ---------------------------------------------------------------------------------------------
constexpr int NF2 = 15;

template <int G>
struct __align__(16) Smem2 {
  static constexpr int FSTR = G * CSTR + 1;
  static constexpr int TILE_X = G + 2;
  static constexpr int TILE_ROW = TILE_X * 6;
  double tile[3 * TILE_ROW];
  double2 rec[NF2 * FSTR];
  int beg[2][G + 1];
  int cnt27[2][27][G];
  unsigned long long bar;
};

template <int ORDER, int G, bool VAY>
__global__ void __launch_bounds__(16 * G, 32 / G)
k_fused2(Geo g, Ptcl A, Ptcl B, const double* __restrict__ id_in, double* __restrict__ id_out,
         const int* __restrict__ inv, Ptcl R, const double* __restrict__ rid, const int* __restrict__ cs, const double* __restrict__ tmpf, double* __restrict__ uj, int* __restrict__ cnt,
         int* __restrict__ hist, unsigned char* __restrict__ dst_off, int* flags, int nxs, int nxe, int ngx, double u0,
         double xend) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using SM = Smem2<G>;
  constexpr int TPB = 16 * G, FSTR = SM::FSTR, TILE_ROW = SM::TILE_ROW;
  SM& S = *reinterpret_cast<SM*>(smem_raw);
  const int t = threadIdx.x;
  const int gx = blockIdx.x % ngx;
  int j, k;
  wm_strip_pencil(g, blockIdx.x / ngx, j, k);
  const int i0 = nxs + gx * G;
  const int ncg = min(G, nxe - i0 + 1);

  if (t == 0) mbar_init(&S.bar, 1);
  if (t < 2 * (G + 1)) {
    const int isp = t / (G + 1), ii = t % (G + 1);
    const int* row = cs + (size_t)g.pen(j, 0, isp) * (g.nx + 1) + (i0 - g.nxgs);
    S.beg[isp][ii] = row[min(ii, ncg)];
  }
  for (int e = t; e < 2 * 27 * G; e += TPB) (&S.cnt27[0][0][0])[e] = 0;
  __syncthreads();
  if (t == 0) {
    const uint32_t row_bytes = (uint32_t)(ncg + 2) * 48u;
    mbar_expect_tx(&S.bar, 3u * row_bytes);
#pragma unroll
    for (int r = 0; r < 3; ++r) tma_bulk_g2s(&S.tile[r * TILE_ROW], tmpf + g.box(i0 - 1, j + r - 1, 0) * 6, row_bytes, &S.bar);
  }
  const int ca = t >> 4, sa = t & 15;
  const int cb = t >> 4, hb = t & 15;
  const int comp = hb / 5, mb = hb % 5;
  const bool b_active = hb < 15 && cb < ncg;
  const int n0a = ca < ncg ? S.beg[0][ca + 1] - S.beg[0][ca] : 0;
  const int n1a = ca < ncg ? S.beg[1][ca + 1] - S.beg[1][ca] : 0;
  const int ncb = cb < ncg ? (S.beg[0][cb + 1] - S.beg[0][cb]) + (S.beg[1][cb + 1] - S.beg[1][cb]) : 0;
  const int nbatch = (max(ncb, __shfl_xor_sync(0xffffffffu, ncb, 16)) + SLOTS - 1) / SLOTS;

  double acc[5];
#pragma unroll
  for (int e = 0; e < 5; ++e) acc[e] = 0.0;
  const double fac = 1.0 / 3.0;
  const double inv_c2 = 1.0 / (g.c * g.c);
  const int ia = i0 + ca;
  const double len_x = __dmul_rd((double)(g.nxge - g.nxgs + 1), g.delx);
  const double len_y = __dmul_rd((double)(g.nyge - g.nygs + 1), g.delx);

  mbar_wait(&S.bar, 0);

  int ns0 = 0, nl0 = 0, ns1 = 0, nl1 = 0;
  for (int batch = 0; batch < nbatch; ++batch) {
    // ------------------------------ phase A ------------------------------
    {
      const int idx = batch * SLOTS + sa;
      int p = -1, isp = 0;
      if (idx < n0a) { p = S.beg[0][ca] + idx; }
      else if (idx < n0a + n1a) { p = S.beg[1][ca] + (idx - n0a); isp = 1; }
      double xn = 0, yn = 0, ux = 0, uy = 0, uz = 0, idv = 0;
      int o = 13;
      if (p >= 0) {
        const int q = inv ? inv[p] : p;          // lazy sort: see k_fused3
        const bool loc = q >= 0;
        const int a = loc ? q : -2 - q;
        const double x = loc ? A.c[0][a] : R.c[0][a], y = loc ? A.c[1][a] : R.c[1][a];
        ux = loc ? A.c[2][a] : R.c[2][a]; uy = loc ? A.c[3][a] : R.c[3][a]; uz = loc ? A.c[4][a] : R.c[4][a];
        idv = loc ? id_in[a] : rid[a];
        double sx[3], sy[3];
        shape3(x * g.d_delx - 0.5 - ia, sx[0], sx[1], sx[2]);
        shape3(y * g.d_delx - 0.5 - j, sy[0], sy[1], sy[2]);
        double f[6];
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) {
          const double2* tr = reinterpret_cast<const double2*>(&S.tile[jj * TILE_ROW + ca * 6]);
          double v[18];
#pragma unroll
          for (int e = 0; e < 9; ++e) { double2 d = tr[e]; v[2 * e] = d.x; v[2 * e + 1] = d.y; }
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            double row = +v[c] * sx[0] + v[6 + c] * sx[1] + v[12 + c] * sx[2];
            f[c] = jj == 0 ? row * sy[0] : f[c] + row * sy[jj];
          }
        }
        const double fac1 = g.fac1[isp];
        const double txxx = fac1 * fac1;
        const double fac2 = g.fac2[isp];
        if (VAY) {
          // particle__solv_vay  2d/common/particle.f90:264-299
          double igam;
          wm_vay_update(f, fac1, fac2, g.c, ux, uy, uz, igam);
          xn = x + ux * g.delt * igam;
          yn = y + uy * g.delt * igam;
        } else {
          const double bpx = f[0], bpy = f[1], bpz = f[2], epx = f[3], epy = f[4], epz = f[5];
          double uvm1 = ux + fac1 * epx, uvm2 = uy + fac1 * epy, uvm3 = uz + fac1 * epz;
          const double qg = g.c * g.c + uvm1 * uvm1 + uvm2 * uvm2 + uvm3 * uvm3;
          double igam = rsqrt(qg);
          double gam = qg * igam;
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
          gam = rsqrt(1.0 + (+ux * ux + uy * uy + uz * uz) * inv_c2);
          xn = x + ux * g.delt * gam;
          yn = y + uy * g.delt * gam;
        }
        if (ORDER != 0) {
          const int ipos = (int)(xn / g.delx);
          if (ipos < nxs + 1) {
            xn = 2.0 * (nxs + 1) * g.delx - xn;
            ux = -ux; uy = -uy; uz = -uz;
          } else if (ORDER == 1 ? ipos >= nxe - 1 : xn > xend) {
            if (ORDER == 1) { xn = 2.0 * (nxe - 1) * g.delx - xn; ux = -ux; }
            else { xn = 2.0 * xend - xn; ux = 2.0 * u0 - ux; }
            uy = -uy; uz = -uz;
          }
        }
        // deposit increments (field.f90:227-246: int(gp*d_delx)); the re-binning cell of y is int(y/delx) under ieee_down
        int inc0 = (int)(xn * g.d_delx) - ia;
        int inc1 = (int)(yn * g.d_delx) - j;
        const int jpos = (int)__ddiv_rd(yn, g.delx);
        int dj = jpos - j;
        if (inc0 < -1 || inc0 > 1 || inc1 < -1 || inc1 > 1 || dj < -1 || dj > 1) {
          atomicOr(flags, 2);
          inc0 = max(-1, min(1, inc0)); inc1 = max(-1, min(1, inc1)); dj = max(-1, min(1, dj));
        }
        const double qdxdt = g.qdxdt[isp];
        const int slot = ca * CSTR + sa;
        {
          double s0[5], ds[5];
          s0ds(x, xn, ia, inc0, g.d_delx, s0, ds);
          const double c0 = -ds[0] * qdxdt, c1 = c0 - ds[1] * qdxdt, c2 = c1 - ds[2] * qdxdt, c3 = c2 - ds[3] * qdxdt;
          S.rec[0 * FSTR + slot] = make_double2(c1, c2);
          S.rec[1 * FSTR + slot] = make_double2(c0, c3);
#pragma unroll
          for (int m = 0; m < 5; ++m) S.rec[(4 + m) * FSTR + slot] = make_double2(s0[m], ds[m]);
          s0ds(y, yn, j, inc1, g.d_delx, s0, ds);
          const double e0 = -ds[0] * qdxdt, e1 = e0 - ds[1] * qdxdt, e2 = e1 - ds[2] * qdxdt, e3 = e2 - ds[3] * qdxdt;
          S.rec[2 * FSTR + slot] = make_double2(e1, e2);
          S.rec[3 * FSTR + slot] = make_double2(e0, e3);
#pragma unroll
          for (int m = 0; m < 5; ++m) S.rec[(9 + m) * FSTR + slot] = make_double2(s0[m], ds[m]);
          // gvz = uz / sqrt(1 + u^2/c^2) of the pushed (and wall-reflected) particle, field.f90:262-264
          const double gvz = uz * rsqrt(1.0 + (+ux * ux + uy * uy + uz * uz) * inv_c2);
          S.rec[14 * FSTR + slot] = make_double2(g.q[isp] * gvz, 0.0);
        }
        // boundaries: periodic x wrap (Weibel order) and periodic y wrap, both under round-down as in the reference
        if (ORDER == 0) {
          const int ipos = (int)__ddiv_rd(xn, g.delx);
          if (ipos < g.nxgs) xn = __dadd_rd(xn, len_x);
          else if (ipos >= g.nxge + 1) xn = __dadd_rd(xn, -len_x);
        }
        if (jpos <= g.nygs - 1) yn = __dadd_rd(yn, len_y);
        else if (jpos >= g.nyge + 1) yn = __dadd_rd(yn, -len_y);
        o = (inc0 + 1) + 3 * (dj + 1) + 9;
        atomicAdd(&S.cnt27[isp][o][ca], 1);
      }
      {
        const unsigned lane = t & 31u;
        const unsigned half = 0xffffu << (lane & 16u);
        const unsigned lower = half & ((1u << lane) - 1u);
        const bool v0 = p >= 0 && isp == 0, v1 = p >= 0 && isp == 1, st = o == 13;
        const unsigned bs0 = __ballot_sync(0xffffffffu, v0 && st) & half, bl0 = __ballot_sync(0xffffffffu, v0 && !st) & half;
        const unsigned bs1 = __ballot_sync(0xffffffffu, v1 && st) & half, bl1 = __ballot_sync(0xffffffffu, v1 && !st) & half;
        if (p >= 0) {
          int pw;
          if (isp == 0) pw = st ? S.beg[0][ca] + ns0 + __popc(bs0 & lower) : S.beg[0][ca + 1] - 1 - (nl0 + __popc(bl0 & lower));
          else          pw = st ? S.beg[1][ca] + ns1 + __popc(bs1 & lower) : S.beg[1][ca + 1] - 1 - (nl1 + __popc(bl1 & lower));
          B.c[0][pw] = xn; B.c[1][pw] = yn;
          B.c[2][pw] = ux; B.c[3][pw] = uy; B.c[4][pw] = uz;
          id_out[pw] = idv;
          if (!st) dst_off[pw] = (unsigned char)o;
        }
        ns0 += __popc(bs0); nl0 += __popc(bl0); ns1 += __popc(bs1); nl1 += __popc(bl1);
      }
    }
    __syncwarp();
    // ------------------------------ phase B ------------------------------
    if (b_active) {
      const int nv = min(SLOTS, ncb - batch * SLOTS);
      if (comp < 2) {
        // Jx (comp 0): acc[r] += cx[r] (S0y + DSy/2)[jp] ; Jy (comp 1): acc[r] += cy[r] (S0x + DSx/2)[ip]
        const int f_c = 2 * comp, f_a = comp == 0 ? 9 + mb : 4 + mb;
        for (int s = 0; s < nv; ++s) {
          const int slot = cb * CSTR + s;
          const double2 c12 = S.rec[f_c * FSTR + slot];
          const double2 c03 = S.rec[(f_c + 1) * FSTR + slot];
          const double2 p1 = S.rec[f_a * FSTR + slot];
          const double Av = p1.x + 0.5 * p1.y;
          acc[0] += c03.x * Av;
          acc[1] += c12.x * Av;
          acc[2] += c12.y * Av;
          acc[3] += c03.y * Av;
        }
      } else {
        // Jz: acc[jp] += q vz (S0x[ip] (S0y + DSy/2)[jp] + DSx[ip] (S0y/2 + DSy/3)[jp])
        for (int s = 0; s < nv; ++s) {
          const int slot = cb * CSTR + s;
          const double2 px = S.rec[(4 + mb) * FSTR + slot];
          const double qg = S.rec[14 * FSTR + slot].x;
          const double qs = qg * px.x, qd = qg * px.y;
#pragma unroll
          for (int jp = 0; jp < 5; ++jp) {
            const double2 py = S.rec[(9 + jp) * FSTR + slot];
            acc[jp] += qs * (py.x + 0.5 * py.y) + qd * (0.5 * py.x + fac * py.y);
          }
        }
      }
    }
    __syncwarp();
  }

  if (b_active) {
    const int ib = i0 + cb;
    const long long sY = (long long)g.bx * 3;
    double* J0 = uj + g.box(ib, j, 0) * 3 + comp;
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      const double v = acc[e];
      if (v != 0.0 && (comp == 2 || e < 4)) {
        long long off;
        if (comp == 0) off = (e - 1) * 3 + (mb - 2) * sY;        // Jx(i+r-1, j+jp-2)
        else if (comp == 1) off = (mb - 2) * 3 + (e - 1) * sY;   // Jy(i+ip-2, j+r-1)
        else off = (mb - 2) * 3 + (e - 2) * sY;                  // Jz(i+ip-2, j+jp-2)
        atomicAdd(J0 + off, v);
      }
    }
  }
  __syncthreads();
  {
    const size_t cell0 = wm_cell_index(g, i0, j, 0);
    for (int e = t; e < G * 2 * WM_CNT_LINE; e += TPB) {
      const int o = e % WM_CNT_LINE, isp = (e / WM_CNT_LINE) % 2, c = e / (2 * WM_CNT_LINE);
      if (c < ncg) cnt[cell0 * 2 * WM_CNT_LINE + e] = o < 27 ? S.cnt27[isp][o][c] : 0;
    }
    for (int e = t; e < 2 * 27 * G; e += TPB) {
      const int c = e % G, o = (e / G) % 27, isp = e / (G * 27);
      const int n = c < ncg ? S.cnt27[isp][o][c] : 0;
      if (n > 0) {
        int drow, ti;
        if (wm_dest_of(g, i0 + c, j, 0, o, isp, nxs, nxe, drow, ti)) atomicAdd(hist + (size_t)drow * (g.nx + 1) + (ti - g.nxgs), n);
        else atomicOr(flags, 2);
      }
    }
  }
}

template <int ORDER, int G, bool VAY>
int launch_fused2(wm_ctx* ctx, int nxs, int nxe, double u0) {
  const Geo& g = ctx->g;
  // per device / context, so not cached in a process-wide static (a process may drive several devices, wm_params.device)
  WM_CUDA(cudaFuncSetAttribute(k_fused2<ORDER, G, VAY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem2<G>)));
  const int ngx = (nxe - nxs + 1 + G - 1) / G;
  const int blocks = ngx * g.nyl;
  const double xend = nxe * g.delx + u0 / sqrt(1 + (u0 * u0) / (g.c * g.c)) * g.delt;   // 2d/proj/shock/boundary_shock.f90:271
  k_fused2<ORDER, G, VAY><<<blocks, 16 * G, sizeof(Smem2<G>), ctx->stream>>>(g, ctx->A, ctx->B, ctx->id[ctx->cid], ctx->id[1 - ctx->cid],
                                                                       ctx->lazy ? ctx->inv : nullptr, ctx->R, ctx->rid, ctx->cs, ctx->tmpf, ctx->uj, ctx->cnt27, ctx->cs_new,
                                                                       ctx->dst_off, ctx->flags, nxs, nxe, ngx, u0, xend);
  WM_LAUNCH_CHECK(ctx);
  return WM_OK;
}

template <int ORDER, int G, bool VAY>
int launch_fused(wm_ctx* ctx, int nxs, int nxe, double u0) {
  const Geo& g = ctx->g;
  WM_CUDA(cudaFuncSetAttribute(k_fused3<ORDER, G, VAY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<G>)));
  const int ngx = (nxe - nxs + 1 + G - 1) / G;
  const int blocks = ngx * g.nyl * g.nzl;
  const double xend = nxe * g.delx + u0 / sqrt(1.0 + (u0 * u0) / (g.c * g.c)) * g.delt;   // boundary_shock.f90:438
  k_fused3<ORDER, G, VAY><<<blocks, 16 * G, sizeof(Smem<G>), ctx->stream>>>(g, ctx->A, ctx->B, ctx->id[ctx->cid], ctx->id[1 - ctx->cid],
                                                                      ctx->lazy ? ctx->inv : nullptr, ctx->R, ctx->rid, ctx->cs, ctx->tmpf, ctx->uj, ctx->cnt27, ctx->cs_new,
                                                                      ctx->dst_off, ctx->flags, nxs, nxe, ngx, u0, xend);
  WM_LAUNCH_CHECK(ctx);
  return WM_OK;
}

}  // namespace

// which (order, boundary kind) pairs the fused kernel covers: the three time loops of the reference with their own boundary module
bool wm_fused_supported(const wm_ctx* ctx, int order) {
  const Geo& g = ctx->g;
  return (order == WM_ORDER_WEIBEL && g.bc == WM_BC_PERIODIC) || (order == WM_ORDER_RECONNECTION && g.bc == WM_BC_RECONNECTION) ||
         (order == WM_ORDER_SHOCK && g.bc == WM_BC_SHOCK);
}

int wm_k_push_deposit_fused(wm_ctx* ctx, int nxs, int nxe, int order, double u0) {
  if (!wm_fused_supported(ctx, order)) {
    wm_set_error("fused path: the time loop (order) must be used with its own boundary module (bc_kind)");
    return WM_ERR_ARG;
  }
  WM_TRY(wm_sort_prepare(ctx));
  const bool vay = ctx->pusher == WM_PUSHER_VAY;
#define WM_DISPATCH(fn, ord, uu) (vay ? fn<ord, 8, true>(ctx, nxs, nxe, uu) : fn<ord, 8, false>(ctx, nxs, nxe, uu))
  if (ctx->g.dim == 2) {
    if (order == WM_ORDER_RECONNECTION) return WM_DISPATCH(launch_fused2, 1, 0.0);
    if (order == WM_ORDER_SHOCK) return WM_DISPATCH(launch_fused2, 2, u0);
    return WM_DISPATCH(launch_fused2, 0, 0.0);
  }
  // (round 2 measured a one-warp-per-cell form of this kernel: slower, see profiles/r02_fused_experiments.md)
  if (order == WM_ORDER_RECONNECTION) return WM_DISPATCH(launch_fused, 1, 0.0);
  if (order == WM_ORDER_SHOCK) return WM_DISPATCH(launch_fused, 2, u0);
  return WM_DISPATCH(launch_fused, 0, 0.0);
#undef WM_DISPATCH
}
