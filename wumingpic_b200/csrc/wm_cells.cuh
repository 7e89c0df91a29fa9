// wm_cells.cuh -- cell / row arithmetic shared by the push kernels (which emit the re-binning information) and the
// sort (which consumes it).  See wm_sort.cu for the algorithm.
//
//   destination offset  o = (di+1) + 3 (dj+1) + 9 (dk+1),  di,dj,dk in {-1,0,1}  (boundary_periodic.f90:152-185 assumes
//                       |displacement| <= one cell); o = 13 is "stays in its cell"
//   cnt line            cnt[(cell*2 + isp)*32 + o]: how many particles of source cell `cell`, species isp, take offset o
//   rows                destination rows of the sort: npen local pencil rows (isp-major, the reference's np2 order), then
//                       -- slab runs only -- 2*nsp*ngrow ghost rows one layer outside the slab along the last axis
#pragma once
#include "wm_internal.cuh"

constexpr int WM_JT = 8;        // j-rows per traversal strip
constexpr int WM_CNT_LINE = 32; // ints per (cell, species) count line (27 used)

__device__ __forceinline__ int wm_unwrap(int v, int lo, int hi, int n) { return v < lo ? v + n : (v > hi ? v - n : v); }

__device__ __forceinline__ size_t wm_cell_index(const Geo& g, int i, int j, int k) {
  return ((size_t)(g.dim == 3 ? (k - g.nzs) : 0) * g.nyl + (j - g.nys)) * g.nx + (i - g.nxgs);
}

// Traversal order of the (j,k) pencils shared by the fused kernel and the sort: j-strips of WM_JT rows, then k, then
// the rows of the strip.  Cells that exchange particles (k+-1, j+-1 neighbours) are then processed within a few MB
// of each other, so their J updates and their re-read source sectors meet in L2.
__device__ __forceinline__ void wm_strip_pencil(const Geo& g, int w, int& j, int& k) {
  const int per_strip = WM_JT * g.nzl;
  const int nstrips = (g.nyl + WM_JT - 1) / WM_JT;
  int s = w / per_strip;
  if (s > nstrips - 1) s = nstrips - 1;
  const int rem = w - s * per_strip;
  const int rows = min(WM_JT, g.nyl - s * WM_JT);
  k = g.dim == 3 ? g.nzs + rem / rows : 0;
  j = g.nys + s * WM_JT + rem % rows;
}

// ghost row of (side, isp, t): t = j - nys in 3-D (ghost plane k = nzs-1 / nze+1), 0 in 2-D (ghost row j = nys-1 / nye+1)
__device__ __forceinline__ int wm_ghost_row(const Geo& g, int side, int isp, int t) {
  return g.npen + (side * g.nsp + isp) * g.ngrow + t;
}

// (row, x cell) a particle of source cell (i,j,k), species isp reaches with offset o; false if it leaves the domain
__device__ __forceinline__ bool wm_dest_of(const Geo& g, int i, int j, int k, int o, int isp, int nxs, int nxe, int& row,
                                           int& ti) {
  const int di = o % 3 - 1, dj = (o / 3) % 3 - 1, dk = o / 9 - 1;
  ti = i + di;
  if (g.bc == WM_BC_PERIODIC) ti = wm_unwrap(ti, g.nxgs, g.nxge, g.nx);
  if (ti < nxs || ti > nxe) return false;
  int tj = j + dj, tk = k + dk;
  if (g.dim == 3) {
    tj = wm_unwrap(tj, g.nys, g.nye, g.nyl);                             // y is periodic inside the slab (nproc_j = 1)
    if (tk < g.nzs) {
      if (g.multi) { row = wm_ghost_row(g, 0, isp, tj - g.nys); return true; }
      tk += g.nzl;
    } else if (tk > g.nze) {
      if (g.multi) { row = wm_ghost_row(g, 1, isp, tj - g.nys); return true; }
      tk -= g.nzl;
    }
  } else {
    if (dk != 0) return false;
    tk = 0;
    if (tj < g.nys) {
      if (g.multi) { row = wm_ghost_row(g, 0, isp, 0); return true; }
      tj += g.nyl;
    } else if (tj > g.nye) {
      if (g.multi) { row = wm_ghost_row(g, 1, isp, 0); return true; }
      tj -= g.nyl;
    }
  }
  row = g.pen(tj, tk, isp);
  return true;
}
