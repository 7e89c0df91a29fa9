// wm_sort.cu -- particle re-binning, inter-GPU migration and the cell-ordered sort, as ONE deterministic
// counting sort keyed by (species, k, j, i):
//
//   boundary_*__particle_y[z]  3d/common/boundary_periodic.f90:104-455  [2d :99-248]
//   sort__bucket               3d/common/sort.f90:40-88                 [2d/common/sort.f90:36-82]
//
// The reference moves a particle at most one cell per step in every direction (it assumes so at
// boundary_periodic.f90:152-185: movers go to pencil jpos,kpos in [nys-1,nye+1] x [nzs-1,nze+1]).  So the whole
// re-binning is described by one byte per particle -- the destination offset o = (di+1) + 3(dj+1) + 9(dk+1)
// relative to its source cell -- and a 27 x ncell count matrix.  From those:
//   producers   the fused push kernel (wm_fused.cu) or k_classify below write the pushed set B "two-ended" inside every
//               source cell -- stayers packed at the front, leavers at the back with their offset byte -- one count
//               line per (cell, species), and add each group's size to the histogram of its destination cell
//               (integer RED; this is the histogram sort.f90:62-69 builds with a per-pencil loop);
//   scan        exclusive sum over all cells = the new cumcnt (sort.f90:71-74), as absolute offsets;
//   k_goff      per DESTINATION cell: exclusive prefix over its <= 27 source groups (one warp, shuffle scan) -> the
//               absolute position where every (source cell, offset) group starts;
//   k_perm      per SOURCE cell (one warp): the permutation.  inv[new position] = old position for the stayers (a block)
//               and for every leaver (rank inside its group = number of earlier leavers of the zone with the same
//               offset byte, from match.any).  Only the offset bytes are read: 1 byte per leaver;
//   k_apply     A[pos] = B[inv[pos]] for the 6 coordinate arrays and the ID, destination-ordered: coalesced
//               full-sector stores, all loads of a thread independent, ~25 instructions per particle.
//               No atomics on the data path; the order inside every cell is deterministic:
//               [local sources by offset, each in zone order][arrivals from the low neighbour][from the high one].
//   (r01 history: a source-centric scatter cost 2.1x read / 1.5x write amplification from partial sectors; a
//    destination-centric gather that re-scanned every leaver zone from its 9 neighbouring rows was instruction bound at
//    36 warp-instructions per particle -- profiles/r01_s3_gather.md.)
// Multi-GPU (slabs along the last axis: z in 3-D, y in 2-D; the reference's rank grid with nproc_j = 1): cells one
// layer outside the slab are extra destination rows ("ghost rows") of the same sort, placed behind the local
// particles.  Their per-cell counts travel first (the reference's count message, boundary_periodic.f90:192,243), are
// added to the receiver's histogram before its scan, and the payload (the ghost rows of the 6 coordinate arrays and
// the ID array, already sorted by destination cell) follows while the local scatter runs; k_insert drops each
// arriving cell group into the slots reserved at the end of its cell.  The bulk particles are read once and
// written once per step; only movers across a slab face make a second trip.
#include "wm_cells.cuh"

#include <algorithm>
#include <cub/device/device_scan.cuh>

namespace {

constexpr int TPB = 256;

// ---------------------------------------------------------------------------------------------
// boundary_*__particle_y[z] on the pushed set (per-procedure path): movers are the particles with
// int(y/delx) != j or int(z/delx) != k (boundary_periodic.f90:152-160); their coordinate is wrapped by the
// GLOBAL periodic condition (:161-171).  x was wrapped/reflected by bc__particle_x; sort__bucket's key is int(x)
// (sort.f90:65).  One warp per (cell, species); the classified set is written two-ended into P (the dead old set).
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(TPB) k_classify(Geo g, Ptcl B, Ptcl P, const double* __restrict__ id_in,
                                                  double* __restrict__ id_out, const int* __restrict__ cs,
                                                  int* __restrict__ cnt, int* __restrict__ hist,
                                                  unsigned char* __restrict__ dst_off, int* flags, int nxs, int nxe) {
  __shared__ int s_cnt[TPB / 32][32];
  constexpr int NC = D == 3 ? 6 : 5;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nxr = nxe - nxs + 1;
  const int npl = g.nyl * g.nzl;
  const long long per_sp = (long long)nxr * npl;
  const double len_y = D == 2 ? __dmul_rd((double)(g.nyge - g.nygs + 1), g.delx) : (g.nyge - g.nygs + 1) * g.delx;
  const double len_z = (g.nzge - g.nzgs + 1) * g.delx;
  for (long long w = warp; w < per_sp * 2; w += nwarps) {
    const int isp = (int)(w / per_sp);
    const long long wi = w % per_sp;
    int j, k;
    wm_strip_pencil(g, (int)(wi / nxr), j, k);
    const int i = nxs + (int)(wi % nxr);
    const int* row = cs + (size_t)g.pen(j, k, isp) * (g.nx + 1) + (i - g.nxgs);
    const int beg = row[0], end = row[1];
    s_cnt[wib][lane] = 0;
    __syncwarp();
    int nfront = 0, nback = 0;
    for (int p0 = beg; p0 < end; p0 += 32) {
      const int p = p0 + lane;
      const bool act = p < end;
      double v[NC], idv = 0.0;
      int o = 13;
      if (act) {
#pragma unroll
        for (int c = 0; c < NC; ++c) v[c] = B.c[c][p];
        idv = id_in[p];
        int ix = (int)v[0];                          // sort.f90:65: no d_delx
        if (ix == nxe + 1) ix = nxe;                 // a periodic wrap that rounded onto the upper edge stays in the last cell
        int di = ix - i;
        if (g.bc == WM_BC_PERIODIC) { if (di > 1) di -= g.nx; else if (di < -1) di += g.nx; }
        const int jpos = D == 2 ? (int)__ddiv_rd(v[1], g.delx) : (int)(v[1] * g.d_delx);   // 2-D: int(y/delx) under ieee_down
        int dj = jpos - j, dk = 0;
        if (jpos <= g.nygs - 1) v[1] = D == 2 ? __dadd_rd(v[1], len_y) : v[1] + len_y;
        else if (jpos >= g.nyge + 1) v[1] = D == 2 ? __dadd_rd(v[1], -len_y) : v[1] - len_y;
        if (D == 3) {
          const int kpos = (int)(v[2] * g.d_delx);
          dk = kpos - k;
          if (kpos <= g.nzgs - 1) v[2] = v[2] + len_z;
          else if (kpos >= g.nzge + 1) v[2] = v[2] - len_z;
        }
        if (di < -1 || di > 1 || dj < -1 || dj > 1 || dk < -1 || dk > 1) {
          atomicOr(flags, 2);
          di = max(-1, min(1, di)); dj = max(-1, min(1, dj)); dk = max(-1, min(1, dk));
        }
        o = (di + 1) + 3 * (dj + 1) + 9 * (dk + 1);
        atomicAdd(&s_cnt[wib][o], 1);
      }
      const unsigned ms = __ballot_sync(0xffffffffu, act && o == 13);
      const unsigned ml = __ballot_sync(0xffffffffu, act && o != 13);
      if (act) {
        const unsigned lower = (1u << lane) - 1u;
        const int pos = o == 13 ? beg + nfront + __popc(ms & lower) : end - 1 - (nback + __popc(ml & lower));
#pragma unroll
        for (int c = 0; c < NC; ++c) P.c[c][pos] = v[c];
        id_out[pos] = idv;
        dst_off[pos] = (unsigned char)o;
      }
      nfront += __popc(ms);
      nback += __popc(ml);
    }
    __syncwarp();
    const size_t cell = wm_cell_index(g, i, j, k);
    const int c = lane < 27 ? s_cnt[wib][lane] : 0;
    cnt[(cell * 2 + isp) * WM_CNT_LINE + lane] = c;
    if (c > 0) {
      int drow, ti;
      if (wm_dest_of(g, i, j, k, lane, isp, nxs, nxe, drow, ti)) atomicAdd(hist + (size_t)drow * (g.nx + 1) + (ti - g.nxgs), c);
      else atomicOr(flags, 2);
    }
    __syncwarp();
  }
}

// hist(edge-plane cell) += arrivals announced by the neighbours; inc = [side][isp][t][0..nx]
__global__ void k_add_incoming(Geo g, const int* __restrict__ inc, int* __restrict__ hist) {
  const int per_side = g.nsp * g.ngrow * (g.nx + 1);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 2 * per_side; e += gridDim.x * blockDim.x) {
    const int side = e / per_side, r = e % per_side;
    const int ii = r % (g.nx + 1), t = (r / (g.nx + 1)) % g.ngrow, isp = r / ((g.nx + 1) * g.ngrow);
    if (ii == g.nx) continue;
    const int c = inc[e];
    if (c == 0) continue;
    int j, k;
    if (g.dim == 3) { j = g.nys + t; k = side == 0 ? g.nzs : g.nze; }
    else { j = side == 0 ? g.nys : g.nye; k = 0; }
    atomicAdd(hist + (size_t)g.pen(j, k, isp) * (g.nx + 1) + ii, c);
  }
}

// ---------------------------------------------------------------------------------------------
// The data movement of the sort: permutation (k_goff, k_perm), then one pass that applies it (k_apply).
// ---------------------------------------------------------------------------------------------
constexpr int GD = 8;       // destination cells per CTA of k_apply

// source cell (si,sj,sk) of destination (i, virtual j, virtual k) and offset o; need_dl: required slab-axis offset of
// a ghost row's sources, 2 = any (local rows)
__device__ __forceinline__ bool src_of(const Geo& g, int i, int j, int k, int need_dl, int o, int nxs, int nxe, int& si,
                                       int& sj, int& sk) {
  const int di = o % 3 - 1, dj = (o / 3) % 3 - 1, dk = o / 9 - 1;
  if (g.dim == 2 && dk != 0) return false;
  const int dl = g.dim == 3 ? dk : dj;
  if (need_dl != 2 && dl != need_dl) return false;
  si = i - di; sj = j - dj; sk = k - dk;
  if (g.bc == WM_BC_PERIODIC) si = wm_unwrap(si, g.nxgs, g.nxge, g.nx);
  if (si < nxs || si > nxe) return false;
  if (g.dim == 3) {
    sj = wm_unwrap(sj, g.nys, g.nye, g.nyl);
    if (sk < g.nzs || sk > g.nze) {
      if (g.multi) return false;                 // that source lives on the neighbour rank
      sk = wm_unwrap(sk, g.nzs, g.nze, g.nzl);
    }
  } else {
    sk = 0;
    if (sj < g.nys || sj > g.nye) {
      if (g.multi) return false;
      sj = wm_unwrap(sj, g.nys, g.nye, g.nyl);
    }
  }
  return true;
}

// destination row (local rows in strip order, ghost rows last) -> (isp, virtual j, virtual k, row id, need_dl)
__device__ __forceinline__ void dest_row(const Geo& g, int wrow, int& isp, int& j, int& k, int& row, int& need_dl) {
  const int npl = g.nyl * g.nzl;
  need_dl = 2;
  if (wrow < g.npen) {
    isp = wrow / npl;
    wm_strip_pencil(g, wrow % npl, j, k);
    row = g.pen(j, k, isp);
  } else {
    row = wrow;
    const int r = row - g.npen;
    const int side = r / (g.nsp * g.ngrow);
    isp = (r / g.ngrow) % g.nsp;
    const int tt = r % g.ngrow;
    need_dl = side == 0 ? -1 : 1;
    if (g.dim == 3) { j = g.nys + tt; k = side == 0 ? g.nzs - 1 : g.nze + 1; }
    else { j = side == 0 ? g.nys - 1 : g.nye + 1; k = 0; }
  }
}

// goff[(row*nx + i-nxgs)*32 + o] = absolute start of the group (source cell of offset o) inside destination cell (row, i).
// One block per destination row and chunk of XCH x-cells; everything that depends only on (row, o) is hoisted out of
// the loop over the cells (the integer divisions of the row decode cost more than the scan itself).
constexpr int XCH = 64;
__global__ void __launch_bounds__(TPB) k_goff(Geo g, const int* __restrict__ cs_new, const int* __restrict__ cnt,
                                              int* __restrict__ goff, int nxs, int nxe, int nch) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int isp, j, k, row, need_dl;
  dest_row(g, blockIdx.x / nch, isp, j, k, row, need_dl);
  const int i_lo = nxs + (blockIdx.x % nch) * XCH, i_hi = min(nxe, i_lo + XCH - 1);
  // lane o: the source row of offset o (cell index of its x = nxgs), or -1
  const int di = lane % 3 - 1;
  long long srow = -1;
  {
    int si, sj, sk;
    if (lane < 27 && src_of(g, nxs + (di > 0 ? 1 : 0), j, k, need_dl, lane, nxs - 1, nxe + 1, si, sj, sk))
      srow = (long long)wm_cell_index(g, g.nxgs, sj, sk);
  }
  const int* crow = cs_new + (size_t)row * (g.nx + 1) - g.nxgs;
  int* grow = goff + ((size_t)row * g.nx - g.nxgs) * WM_CNT_LINE + lane;
  for (int i = i_lo + wib; i <= i_hi; i += TPB / 32) {
    int c = 0;
    if (srow >= 0) {
      int si = i - di;
      if (g.bc == WM_BC_PERIODIC) si = wm_unwrap(si, g.nxgs, g.nxge, g.nx);
      if (si >= nxs && si <= nxe) c = cnt[((size_t)(srow + (si - g.nxgs)) * 2 + isp) * WM_CNT_LINE + lane];
    }
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    grow[(size_t)i * WM_CNT_LINE] = crow[i] + incl - c;
  }
}

// inv[new position] = old position, source-centric: one warp per (cell, species); one block per source row and chunk of
// XCH x-cells, the destination rows of the 27 offsets hoisted out of the loop over the cells
__global__ void __launch_bounds__(TPB) k_perm(Geo g, const int* __restrict__ cs, const int* __restrict__ cnt,
                                              const int* __restrict__ goff, const unsigned char* __restrict__ dst_off,
                                              int* __restrict__ inv, int* flags, int nxs, int nxe, int nch) {
  __shared__ int s_run[TPB / 32][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int npl = g.nyl * g.nzl;
  const int wrow = blockIdx.x / nch;
  const int isp = wrow / npl;
  int j, k;
  wm_strip_pencil(g, wrow % npl, j, k);
  const int i_lo = nxs + (blockIdx.x % nch) * XCH, i_hi = min(nxe, i_lo + XCH - 1);
  // lane o: destination row of offset o (x handled per cell), or -1 if that offset leaves the domain in y / z
  const int di = lane % 3 - 1;
  int drow = -1;
  {
    int ti;
    if (lane < 27 && wm_dest_of(g, nxs + (di < 0 ? 1 : 0), j, k, lane, isp, nxs - 1, nxe + 1, drow, ti)) {} else drow = -1;
  }
  const int* row = cs + (size_t)g.pen(j, k, isp) * (g.nx + 1) - g.nxgs;
  const size_t cell0 = wm_cell_index(g, g.nxgs, j, k);
  for (int i = i_lo + wib; i <= i_hi; i += TPB / 32) {
    const int beg = row[i], end = row[i + 1];
    if (beg == end) continue;
    // lane o: where the group of offset o starts in its destination cell
    const int c = lane < 27 ? cnt[((cell0 + (i - g.nxgs)) * 2 + isp) * WM_CNT_LINE + lane] : 0;
    int mygoff = -1;
    if (c > 0) {
      int ti = i + di;
      if (g.bc == WM_BC_PERIODIC) ti = wm_unwrap(ti, g.nxgs, g.nxge, g.nx);
      if (drow >= 0 && ti >= nxs && ti <= nxe) mygoff = goff[((size_t)drow * g.nx + (ti - g.nxgs)) * WM_CNT_LINE + lane];
      else atomicOr(flags, 2);
    }
    const int nstay = __shfl_sync(0xffffffffu, c, 13);
    const int g13 = __shfl_sync(0xffffffffu, mygoff, 13);
    for (int e = lane; e < nstay; e += 32) inv[g13 + e] = beg + e;
    if (beg + nstay == end) continue;
    s_run[wib][lane] = 0;
    __syncwarp();
    for (int p0 = beg + nstay; p0 < end; p0 += 32) {
      const int p = p0 + lane;
      const bool act = p < end;
      const int o = act ? (int)dst_off[p] : 31;                 // 31: idle lanes form their own group
      const unsigned m = __match_any_sync(0xffffffffu, o);
      const int rank = s_run[wib][o] + __popc(m & ((1u << lane) - 1u));
      const int gbase = __shfl_sync(0xffffffffu, mygoff, o & 31);
      __syncwarp();
      if (act && (m & ((1u << lane) - 1u)) == 0) s_run[wib][o] += __popc(m);   // group leader
      if (act && gbase >= 0) inv[gbase + rank] = p;
      __syncwarp();
    }
  }
}

// A[pos] = B[inv[pos]]: one CTA per run of GD destination cells, rows in strip order (sources of a run live within a few
// MB of it in the same traversal, so partially used source sectors are still in L2 when their other users arrive)
template <int D>
__global__ void __launch_bounds__(TPB) k_apply(Geo g, Ptcl B, Ptcl A, const double* __restrict__ id_in,
                                               double* __restrict__ id_out, const int* __restrict__ cs_new,
                                               const int* __restrict__ inv, int nxs, int nxe, int ngx, int row0, Ptcl R,
                                               const double* __restrict__ rid) {
  constexpr int NC = D == 3 ? 6 : 5;
  const int gx = blockIdx.x % ngx;
  int isp, j, k, row, need_dl;
  dest_row(g, row0 + blockIdx.x / ngx, isp, j, k, row, need_dl);
  const int ia = nxs + gx * GD;
  const int ncg = min(GD, nxe - ia + 1);
  const int* crow = cs_new + (size_t)row * (g.nx + 1) + (ia - g.nxgs);
  const int base = crow[0], end = crow[ncg];
  for (int pos = base + threadIdx.x; pos < end; pos += TPB) {
    const int p = inv[pos];
    if (p == -1) continue;                   // a slot reserved for an arrival from a neighbour rank (k_insert fills it)
    double v[NC + 1];
    if (p >= 0) {
#pragma unroll
      for (int c = 0; c < NC; ++c) v[c] = B.c[c][p];
      v[NC] = id_in[p];
    } else {                                 // lazy slab run: the arrival sits in the arrival store
      const int q = -2 - p;
#pragma unroll
      for (int c = 0; c < NC; ++c) v[c] = R.c[c][q];
      v[NC] = rid[q];
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) A.c[c][pos] = v[c];
    id_out[pos] = v[NC];
  }
}

// inv = -1 on the slots k_insert will fill (the tail of every edge-plane cell that receives arrivals): k_apply skips them.
// Same slot arithmetic as k_insert; replaces a memset of the whole permutation array (cap x 4 B per step).
// inc_off != nullptr (lazy sort): the slot instead points at the arrival itself, inv = -2 - (index in the arrival store).
__global__ void k_mark_arrivals(Geo g, const int* __restrict__ inc, const int* __restrict__ cs_new, int* __restrict__ inv,
                                const int* __restrict__ inc_off) {
  const int lane = threadIdx.x & 31;
  const int warp = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int nwarps = (int)(((long long)gridDim.x * blockDim.x) >> 5);
  const int per_side = g.nsp * g.ngrow * (g.nx + 1);
  const bool same_plane = g.dim == 3 ? g.nzs == g.nze : g.nys == g.nye;
  for (int e = warp; e < 2 * per_side; e += nwarps) {
    const int n = inc[e];
    if (n == 0) continue;
    const int side = e / per_side, r = e % per_side;
    const int ii = r % (g.nx + 1), t = (r / (g.nx + 1)) % g.ngrow, isp = r / ((g.nx + 1) * g.ngrow);
    if (ii == g.nx) continue;
    int j, k;
    if (g.dim == 3) { j = g.nys + t; k = side == 0 ? g.nzs : g.nze; }
    else { j = side == 0 ? g.nys : g.nye; k = 0; }
    const int cell_end = cs_new[(size_t)g.pen(j, k, isp) * (g.nx + 1) + ii + 1];
    int dst = cell_end - n;
    if (side == 0 && same_plane) dst -= inc[per_side + r];
    const int src = inc_off ? inc_off[e] : 0;
    for (int q = lane; q < n; q += 32) inv[dst + q] = inc_off ? -2 - (src + q) : -1;
  }
}

// arrivals: recv (in B / the spare ID array) holds [from the low neighbour | from the high neighbour], each sorted by
// destination cell with the counts `inc` and exclusive offsets `inc_off`.  One warp per arriving cell group.
template <int D>
__global__ void __launch_bounds__(TPB) k_insert(Geo g, Ptcl R, const double* __restrict__ rid, Ptcl A,
                                                double* __restrict__ id_out, const int* __restrict__ inc,
                                                const int* __restrict__ inc_off, const int* __restrict__ cs_new) {
  const int lane = threadIdx.x & 31;
  const int warp = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int nwarps = (int)(((long long)gridDim.x * blockDim.x) >> 5);
  const int per_side = g.nsp * g.ngrow * (g.nx + 1);
  constexpr int NC = D == 3 ? 6 : 5;
  const bool same_plane = g.dim == 3 ? g.nzs == g.nze : g.nys == g.nye;
  for (int e = warp; e < 2 * per_side; e += nwarps) {
    const int n = inc[e];
    if (n == 0) continue;
    const int side = e / per_side, r = e % per_side;
    const int ii = r % (g.nx + 1), t = (r / (g.nx + 1)) % g.ngrow, isp = r / ((g.nx + 1) * g.ngrow);
    int j, k;
    if (g.dim == 3) { j = g.nys + t; k = side == 0 ? g.nzs : g.nze; }
    else { j = side == 0 ? g.nys : g.nye; k = 0; }
    const int cell_end = cs_new[(size_t)g.pen(j, k, isp) * (g.nx + 1) + ii + 1];
    int dst = cell_end - n;
    if (side == 0 && same_plane) dst -= inc[per_side + r];   // the high neighbour's arrivals come last
    const int src = inc_off[e];
    for (int q = lane; q < n; q += 32) {
#pragma unroll
      for (int c = 0; c < NC; ++c) A.c[c][dst + q] = R.c[c][src + q];
      id_out[dst + q] = rid[src + q];
    }
  }
}

// `expect` >= 0: the population the sort must conserve (single rank); a mismatch means a particle pointed outside the
// domain (|displacement| > one cell, or through a wall) and no destination claimed it
__global__ void k_np2_poff(Geo g, const int* __restrict__ cs, int* __restrict__ np2, int* __restrict__ poff, int* flags,
                           long long expect) {
  for (int pen = blockIdx.x * blockDim.x + threadIdx.x; pen <= g.npen; pen += gridDim.x * blockDim.x) {
    if (pen < g.npen) {
      const int* row = cs + (size_t)pen * (g.nx + 1);
      int n = row[g.nx] - row[0];
      np2[pen] = n;
      poff[pen] = row[0];
      if (n > g.np) atomicOr(flags, 1);   // "memory over (np2 > np)"  boundary_periodic.f90:435-438
    } else {
      const int total = cs[(size_t)g.npen * (g.nx + 1)];
      poff[pen] = total;
      if (expect >= 0 && total != expect) atomicOr(flags, 2);
    }
  }
}

// totals the host needs in multi-rank mode: [0] local particles, [1] species-1 start, [2..3] ghost sends lo/hi,
// [4..5] arrivals lo/hi
__global__ void k_totals(Geo g, const int* __restrict__ cs_new, const int* __restrict__ inc_off, int* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int w = g.nx + 1;
  const int g0 = cs_new[(size_t)g.npen * w];
  const int g1 = cs_new[(size_t)(g.npen + g.nsp * g.ngrow) * w];
  const int gend = cs_new[(size_t)g.nrows * w];
  const int per_side = g.nsp * g.ngrow * w;
  out[0] = g0;
  out[1] = cs_new[(size_t)(g.npen / g.nsp) * w];
  out[2] = g1 - g0;
  out[3] = gend - g1;
  out[4] = inc_off[per_side];
  out[5] = inc_off[2 * per_side] - inc_off[per_side];
}

int grid_for(long long n) {
  long long b = (n + TPB - 1) / TPB;
  const long long cap = 148LL * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

int ensure_sort_buffers(wm_ctx* ctx) {
  const Geo& g = ctx->g;
  const size_t ncell = (size_t)g.nx * g.nyl * g.nzl;
  if (!ctx->cnt27) {
    WM_CUDA(cudaMalloc(&ctx->cnt27, ncell * 2 * WM_CNT_LINE * sizeof(int)));
    WM_CUDA(cudaMemsetAsync(ctx->cnt27, 0, ncell * 2 * WM_CNT_LINE * sizeof(int), ctx->stream));
  }
  if (ctx->dst_off_cap < ctx->cap) {
    if (ctx->dst_off) cudaFree(ctx->dst_off);
    if (ctx->inv) cudaFree(ctx->inv);
    ctx->dst_off = nullptr;
    ctx->inv = nullptr;
    WM_CUDA(cudaMalloc(&ctx->dst_off, ctx->cap));
    WM_CUDA(cudaMalloc(&ctx->inv, ctx->cap * sizeof(int)));
    WM_CUDA(cudaMemsetAsync(ctx->inv, 0, ctx->cap * sizeof(int), ctx->stream));   // stale entries must stay valid indices
    ctx->dst_off_cap = ctx->cap;
  }
  const size_t ngoff = ((size_t)g.npen + 2 * g.nsp * g.ngrow) * g.nx * WM_CNT_LINE;
  if (!ctx->goff) WM_CUDA(cudaMalloc(&ctx->goff, ngoff * sizeof(int)));
  return WM_OK;
}

int scan_ints(wm_ctx* ctx, int* data, size_t n) {
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, data, data, (int)n, ctx->stream);
  if (need > ctx->scan_tmp_bytes) {
    if (ctx->scan_tmp) cudaFree(ctx->scan_tmp);
    ctx->scan_tmp = nullptr;
    WM_CUDA(cudaMalloc(&ctx->scan_tmp, need));
    ctx->scan_tmp_bytes = need;
  }
  WM_CUDA(cub::DeviceScan::ExclusiveSum(ctx->scan_tmp, need, data, data, (int)n, ctx->stream));
  ctx->launches += 2;
  return WM_OK;
}

// flags[1] |= 1 if some particle's int(x) is not the cell the index puts it in (one thread per (pencil, cell); runs once per upload)
__global__ void k_cells_consistent(Geo g, const double* __restrict__ x, const int* __restrict__ cs, int* flags) {
  const long long n = (long long)g.npen * g.nx;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int pen = (int)(e / g.nx), ci = (int)(e % g.nx);
    const int* row = cs + (size_t)pen * (g.nx + 1) + ci;
    bool bad = false;
    for (int p = row[0]; p < row[1]; ++p) bad |= (int)x[p] != g.nxgs + ci;     // sort.f90:65: the cell of a particle is int(x)
    if (bad) atomicOr(flags + 1, 1);
  }
}

}  // namespace

// Does the uploaded cell index agree with the particles?  It always does after a sort__bucket; the one state of the reference where
// it does not is the shock driver's freshly loaded box (nominal cumcnt).  *inconsistent = 1 if not.
int wm_k_cells_consistent(wm_ctx* ctx, int* inconsistent) {
  const Geo& g = ctx->g;
  *inconsistent = 0;
  if (ctx->ntot == 0) return WM_OK;
  WM_CUDA(cudaMemsetAsync(ctx->flags + 1, 0, sizeof(int), ctx->stream));
  k_cells_consistent<<<grid_for((long long)g.npen * g.nx), TPB, 0, ctx->stream>>>(g, ctx->A.c[0], ctx->cs, ctx->flags);
  WM_LAUNCH_CHECK(ctx);
  WM_CUDA(cudaMemcpyAsync(inconsistent, ctx->flags + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  WM_CUDA(cudaStreamSynchronize(ctx->stream));
  return WM_OK;
}

// called by the producers (fused push kernel, k_classify) before they accumulate the histogram
int wm_sort_prepare(wm_ctx* ctx) {
  WM_TRY(ensure_sort_buffers(ctx));
  const size_t ncs = (size_t)ctx->g.nrows * (ctx->g.nx + 1);
  WM_CUDA(cudaMemsetAsync(ctx->cs_new, 0, (ncs + 1) * sizeof(int), ctx->stream));
  return WM_OK;
}

// boundary_*__particle_y[z]: classification + coordinate wrap of the pushed set (per-procedure path).  The classified
// copy is written into the storage of the old sorted set (dead after the deposit); the two sets then swap names.
int wm_k_classify(wm_ctx* ctx, int nxs, int nxe) {
  const Geo& g = ctx->g;
  WM_TRY(wm_sort_prepare(ctx));
  const long long nwork = (long long)(nxe - nxs + 1) * g.nyl * g.nzl * 2;
  const int blocks = (int)std::min<long long>((nwork * 32 + TPB - 1) / TPB, 148LL * 8);
  if (g.dim == 3)
    k_classify<3><<<blocks, TPB, 0, ctx->stream>>>(g, ctx->B, ctx->A, ctx->id[ctx->cid], ctx->id[1 - ctx->cid], ctx->cs,
                                                   ctx->cnt27, ctx->cs_new, ctx->dst_off, ctx->flags, nxs, nxe);
  else
    k_classify<2><<<blocks, TPB, 0, ctx->stream>>>(g, ctx->B, ctx->A, ctx->id[ctx->cid], ctx->id[1 - ctx->cid], ctx->cs,
                                                   ctx->cnt27, ctx->cs_new, ctx->dst_off, ctx->flags, nxs, nxe);
  WM_LAUNCH_CHECK(ctx);
  std::swap(ctx->A, ctx->B);   // B = classified pushed set, A = free for the sorted output
  return WM_OK;
}

// sort__bucket (+ the migration half of bc__particle_y[z]): needs dst_off / cnt27 of the pushed set B
// WM_SORT_TIMING=1 (measurement aid): device time of the phases of wm_k_sort, summed per rank and printed by wm_sort_timing_report
static double g_sort_ms[8] = {0};
static long g_sort_calls = 0;
void wm_sort_timing_report(int rank) {
  if (g_sort_calls == 0) return;
  fprintf(stderr, "[wuming_b200] rank %d sort phases over %ld calls (ms/call): counts-exchange %.3f scans+totals(host sync) %.3f goff+mark+perm %.3f "
                  "apply(ghost rows) %.3f payload-exchange %.3f insert+np2 %.3f\n", rank, g_sort_calls, g_sort_ms[0] / g_sort_calls,
          g_sort_ms[1] / g_sort_calls, g_sort_ms[2] / g_sort_calls, g_sort_ms[3] / g_sort_calls, g_sort_ms[4] / g_sort_calls,
          g_sort_ms[5] / g_sort_calls);
  g_sort_calls = 0;
}

int wm_k_sort(wm_ctx* ctx, int nxs, int nxe) {
  const Geo& g = ctx->g;
  cudaStream_t st = ctx->stream;
  static const bool timing = getenv("WM_SORT_TIMING") != nullptr;
  cudaEvent_t tev[7] = {};
  int tn = 0;
  auto mark = [&]() { if (timing && tn < 7) { if (!tev[tn]) cudaEventCreate(&tev[tn]); cudaEventRecord(tev[tn++], st); } };
  mark();
  const int w = g.nx + 1;
  const size_t ncs = (size_t)g.nrows * w;
  const int per_side = g.nsp * g.ngrow * w;
  // on entry: B / id[1-cid] hold the pushed set two-ended per source cell, cnt27 the count lines and cs_new the
  // destination histogram (all written by the fused push kernel or by k_classify)
  int tot[6] = {0, 0, 0, 0, 0, 0};
  if (g.multi) {
    // count message: my low ghost rows -> down neighbour (arrive in its high plane), high ghost rows -> up neighbour
    int* ghost_lo = ctx->cs_new + (size_t)g.npen * w;
    int* ghost_hi = ghost_lo + per_side;
    const int ax = g.dim == 3 ? 1 : 0;
    WM_TRY(wm_comm_group_begin(ctx));
    WM_TRY(wm_comm_send(ctx, ctx->rank_down[ax], ghost_lo, (size_t)per_side * sizeof(int)));
    WM_TRY(wm_comm_recv(ctx, ctx->rank_up[ax], ctx->inc + per_side, (size_t)per_side * sizeof(int)));
    WM_TRY(wm_comm_send(ctx, ctx->rank_up[ax], ghost_hi, (size_t)per_side * sizeof(int)));
    WM_TRY(wm_comm_recv(ctx, ctx->rank_down[ax], ctx->inc, (size_t)per_side * sizeof(int)));
    WM_TRY(wm_comm_group_end(ctx));
    k_add_incoming<<<grid_for(2 * per_side), TPB, 0, st>>>(g, ctx->inc, ctx->cs_new);
    WM_LAUNCH_CHECK(ctx);
    WM_CUDA(cudaMemcpyAsync(ctx->inc_off, ctx->inc, (size_t)2 * per_side * sizeof(int), cudaMemcpyDeviceToDevice, st));
    WM_CUDA(cudaMemsetAsync(ctx->inc_off + 2 * per_side, 0, sizeof(int), st));
    WM_TRY(scan_ints(ctx, ctx->inc_off, (size_t)2 * per_side + 1));
  }
  mark();
  WM_TRY(scan_ints(ctx, ctx->cs_new, ncs + 1));
  if (g.multi) {
    k_totals<<<1, 32, 0, st>>>(g, ctx->cs_new, ctx->inc_off, ctx->totals);
    WM_LAUNCH_CHECK(ctx);
    WM_CUDA(cudaMemcpyAsync(tot, ctx->totals, sizeof(tot), cudaMemcpyDeviceToHost, st));
    WM_CUDA(cudaStreamSynchronize(st));
    const size_t need = (size_t)tot[0] + tot[2] + tot[3];
    if (need > ctx->cap || (size_t)tot[4] + tot[5] > ctx->cap) {
      wm_set_error("particle arrays too small for the migrated population (raise WM_CAP_FACTOR)");
      return WM_ERR_MEMORY_OVER;
    }
  }
  mark();
  const int old_cid = 1 - ctx->cid;   // the producers moved the IDs along with the particles into the spare array
  static const bool no_lazy = getenv("WM_NO_LAZY_SORT") != nullptr;   // measurement switch
  const bool lazy = ctx->allow_lazy && !no_lazy;
  const int nxr = nxe - nxs + 1;
  const int ngx = (nxr + GD - 1) / GD;
  {
    const int nch = (nxr + XCH - 1) / XCH;
    k_goff<<<g.nrows * nch, TPB, 0, st>>>(g, ctx->cs_new, ctx->cnt27, ctx->goff, nxs, nxe, nch);
    WM_LAUNCH_CHECK(ctx);
    if (g.multi) {   // arrival slots: -1 for k_apply / k_insert, or the index into the arrival store (lazy)
      k_mark_arrivals<<<std::min(wm_blocks((long long)2 * per_side * 32, TPB), 148 * 8), TPB, 0, st>>>(g, ctx->inc, ctx->cs_new, ctx->inv,
                                                                                                      lazy ? ctx->inc_off : nullptr);
      WM_LAUNCH_CHECK(ctx);
    }
    k_perm<<<g.npen * nch, TPB, 0, st>>>(g, ctx->cs, ctx->cnt27, ctx->goff, ctx->dst_off, ctx->inv, ctx->flags, nxs, nxe, nch);
    WM_LAUNCH_CHECK(ctx);
    mark();
    // the permutation is applied now -- or, inside wm_step, left to the next fused kernel (which reads through inv); a slab
    // run still materialises its outgoing ghost rows, behind the local particles of the free set
    const int row0 = lazy ? g.npen : 0;
    const int rows = g.nrows - row0;
    if (rows > 0) {
      if (g.dim == 3)
        k_apply<3><<<rows * ngx, TPB, 0, st>>>(g, ctx->B, ctx->A, ctx->id[old_cid], ctx->id[1 - old_cid], ctx->cs_new, ctx->inv, nxs,
                                               nxe, ngx, row0, ctx->R, ctx->rid);
      else
        k_apply<2><<<rows * ngx, TPB, 0, st>>>(g, ctx->B, ctx->A, ctx->id[old_cid], ctx->id[1 - old_cid], ctx->cs_new, ctx->inv, nxs,
                                               nxe, ngx, row0, ctx->R, ctx->rid);
      WM_LAUNCH_CHECK(ctx);
    }
  }
  mark();
  if (g.multi) {
    // payload: ghost rows of A (behind the local particles) -> neighbours; arrivals land in B / the old ID array,
    // which are free once the scatter has read them (stream order) -- or, lazy, in the arrival store R
    const int ax = g.dim == 3 ? 1 : 0;
    const int ncomp = g.ndim - 1;
    const size_t lo0 = (size_t)tot[0], hi0 = (size_t)tot[0] + tot[2];
    if (lazy && (size_t)tot[4] + tot[5] > ctx->rcap) {
      const size_t rcap = ((size_t)tot[4] + tot[5]) * 3 / 2 + 4096;
      for (int c = 0; c < ncomp; ++c) {
        if (ctx->R.c[c]) cudaFree(ctx->R.c[c]);
        WM_CUDA(cudaMalloc(&ctx->R.c[c], rcap * sizeof(double)));
      }
      if (ctx->rid) cudaFree(ctx->rid);
      WM_CUDA(cudaMalloc(&ctx->rid, rcap * sizeof(double)));
      ctx->rcap = rcap;
    }
    WM_TRY(wm_comm_group_begin(ctx));
    for (int c = 0; c <= ncomp; ++c) {
      const double* src = c < ncomp ? ctx->A.c[c] : ctx->id[1 - old_cid];
      double* dst = lazy ? (c < ncomp ? ctx->R.c[c] : ctx->rid) : (c < ncomp ? ctx->B.c[c] : ctx->id[old_cid]);
      WM_TRY(wm_comm_send(ctx, ctx->rank_down[ax], src + lo0, (size_t)tot[2] * sizeof(double)));
      WM_TRY(wm_comm_recv(ctx, ctx->rank_up[ax], dst + tot[4], (size_t)tot[5] * sizeof(double)));
      WM_TRY(wm_comm_send(ctx, ctx->rank_up[ax], src + hi0, (size_t)tot[3] * sizeof(double)));
      WM_TRY(wm_comm_recv(ctx, ctx->rank_down[ax], dst, (size_t)tot[4] * sizeof(double)));
    }
    WM_TRY(wm_comm_group_end(ctx));
    mark();
    if (!lazy && tot[4] + tot[5] > 0) {
      const int blocks = std::min(wm_blocks((long long)2 * per_side * 32, TPB), 148 * 8);
      if (g.dim == 3)
        k_insert<3><<<blocks, TPB, 0, st>>>(g, ctx->B, ctx->id[old_cid], ctx->A, ctx->id[1 - old_cid], ctx->inc, ctx->inc_off,
                                            ctx->cs_new);
      else
        k_insert<2><<<blocks, TPB, 0, st>>>(g, ctx->B, ctx->id[old_cid], ctx->A, ctx->id[1 - old_cid], ctx->inc, ctx->inc_off,
                                            ctx->cs_new);
      WM_LAUNCH_CHECK(ctx);
    }
    const long long kept = (long long)tot[0] - tot[4] - tot[5] + tot[2] + tot[3];
    if (kept != ctx->ntot) {
      wm_set_error("a particle left the one-cell neighbourhood of its cell (no destination cell claimed it)");
      return WM_ERR_PARTICLE_LOST;
    }
    ctx->ntot = tot[0];
    ctx->n_sp0 = tot[1];
  }
  std::swap(ctx->cs, ctx->cs_new);
  if (lazy) {
    std::swap(ctx->A, ctx->B);   // set A = the pushed set (read through inv), set B = free
    ctx->cid = old_cid;
    ctx->lazy = true;
    ctx->lazy_nxs = nxs;
    ctx->lazy_nxe = nxe;
  } else {
    ctx->cid = 1 - old_cid;
  }
  k_np2_poff<<<wm_blocks(g.npen + 1, TPB), TPB, 0, st>>>(g, ctx->cs, ctx->np2, ctx->poff, ctx->flags,
                                                           g.multi ? -1LL : ctx->ntot);
  WM_LAUNCH_CHECK(ctx);
  if (timing) {
    mark();
    cudaEventSynchronize(tev[tn - 1]);
    for (int e = 0; e + 1 < tn; ++e) { float ms = 0; cudaEventElapsedTime(&ms, tev[e], tev[e + 1]); g_sort_ms[tn == 7 ? e : (e < 4 ? e : e + 1)] += ms; }
    for (int e = 0; e < tn; ++e) cudaEventDestroy(tev[e]);
    g_sort_calls++;
  }
  return WM_OK;
}

// np2 / poff from the cell index (after the index was edited outside the sort: wm_shock.cu)
int wm_k_refresh_np2(wm_ctx* ctx) {
  const Geo& g = ctx->g;
  k_np2_poff<<<wm_blocks(g.npen + 1, TPB), TPB, 0, ctx->stream>>>(g, ctx->cs, ctx->np2, ctx->poff, ctx->flags,
                                                                    g.multi ? -1LL : ctx->ntot);
  WM_LAUNCH_CHECK(ctx);
  return WM_OK;
}

// Apply a pending (lazy) permutation: afterwards set A is the cell-sorted set every other kernel expects.
int wm_materialize(wm_ctx* ctx) {
  if (!ctx->lazy) return WM_OK;
  const Geo& g = ctx->g;
  const int nxs = ctx->lazy_nxs, nxe = ctx->lazy_nxe;
  const int ngx = (nxe - nxs + 1 + GD - 1) / GD;
  if (g.dim == 3)
    k_apply<3><<<g.npen * ngx, TPB, 0, ctx->stream>>>(g, ctx->A, ctx->B, ctx->id[ctx->cid], ctx->id[1 - ctx->cid], ctx->cs, ctx->inv,
                                                      nxs, nxe, ngx, 0, ctx->R, ctx->rid);
  else
    k_apply<2><<<g.npen * ngx, TPB, 0, ctx->stream>>>(g, ctx->A, ctx->B, ctx->id[ctx->cid], ctx->id[1 - ctx->cid], ctx->cs, ctx->inv,
                                                      nxs, nxe, ngx, 0, ctx->R, ctx->rid);
  WM_LAUNCH_CHECK(ctx);
  std::swap(ctx->A, ctx->B);
  ctx->cid = 1 - ctx->cid;
  ctx->lazy = false;
  return WM_OK;
}
