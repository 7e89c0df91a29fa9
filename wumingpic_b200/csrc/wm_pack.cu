// wm_pack.cu -- particle packing for the output procedures on the device: paraio's get_particle_count
// (3d/common/paraio.f90:1007-1085 [2d/common/paraio.f90, same routine without k]) walks up(:, 1:np2(j,k,isp), j, k, isp) in
// (species, k, j, i) order and packs either every active particle (mode 0: io__ptcl) or the tracers, i.e. the particles
// with a positive ID (mode 1: io__orb), as consecutive ndim-double records.  On the host that needs the whole particle
// store downloaded first; here the selection is a stable stream compaction over the cell-sorted SoA -- whose global order
// IS (species, k, j, cell) -- followed by one SoA -> AoS gather, and only the packed records cross PCIe.
// The order of the records inside one cell is this backend's (deterministic) cell order, which differs from the
// reference's insertion order inside a cell; the reference defines no order there either (its sort is a counting sort
// over OpenMP threads) and the readers of the _orb / _ptcl files identify particles by ID.
#include "wm_internal.cuh"

#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>
#include <vector>

namespace {

constexpr int TPB = 256;

struct IsTracer {
  const double* id;
  __device__ bool operator()(const int& p) const { return __double_as_longlong(id[p]) > 0; }
};

// records sel[first .. first+n) -> stage (ndim doubles each)
// swap: 3-D y-slabs (wm_ctx::swap_yz): the caller's y, uy are the device's z', uz' columns and vice versa
__global__ void k_pack_records(Geo g, Ptcl A, const double* __restrict__ id, const int* __restrict__ sel, long long first,
                               long long n, double* __restrict__ stage, int swap) {
  const int ncomp = g.ndim - 1;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long p = sel ? sel[first + e] : first + e;
    double* o = stage + e * g.ndim;
    for (int c = 0; c < ncomp; ++c) {
      const int cs = !swap ? c : (c == 1 ? 2 : (c == 2 ? 1 : (c == 4 ? 5 : (c == 5 ? 4 : c))));
      o[c] = A.c[cs][p];
    }
    o[ncomp] = id[p];
  }
}

__global__ void k_count_below(const int* __restrict__ sel, int nsel, long long bound, int* __restrict__ out) {
  int c = 0;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nsel; e += gridDim.x * blockDim.x) c += sel[e] < bound ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

}  // namespace

extern "C" int wm_pack_particles(wm_ctx* c, int mode, double* buf, long long cap_records, long long* lcount) {
  // 3-D y-slabs: the records come out in the caller's columns but in (species, j, k, cell) pencil order instead of paraio's
  // (species, k, j, cell) -- the readers of the _ptcl / _orb files identify particles by ID (see the header of this file)
  if (!c || !lcount || (mode != 0 && mode != 1)) {
    wm_set_error("Error: invalid mode specified for get_particle_count");   // paraio.f90:1065-1069
    return WM_ERR_ARG;
  }
  WM_CUDA(cudaSetDevice(c->device));
  if (c->gp_valid) { wm_set_error("particle output acts on the sorted particles (call it after sort__bucket)"); return WM_ERR_STATE; }
  WM_TRY(wm_materialize(c));
  const Geo& g = c->g;
  cudaStream_t st = c->stream;
  const long long ntot = c->ntot;
  int* sel = nullptr;
  long long nsel = ntot;
  lcount[0] = c->n_sp0;
  lcount[1] = ntot - c->n_sp0;
  if (mode == 1 && ntot > 0) {
    int* nsel_dev = nullptr;
    WM_CUDA(cudaMalloc(&sel, (size_t)ntot * sizeof(int)));
    WM_CUDA(cudaMalloc(&nsel_dev, 2 * sizeof(int)));
    WM_CUDA(cudaMemsetAsync(nsel_dev, 0, 2 * sizeof(int), st));
    thrust::counting_iterator<int> idx(0);
    IsTracer pred{c->id[c->cid]};
    size_t need = 0;
    cub::DeviceSelect::If(nullptr, need, idx, sel, nsel_dev, (int)ntot, pred, st);
    void* tmp = nullptr;
    WM_CUDA(cudaMalloc(&tmp, need));
    WM_CUDA(cub::DeviceSelect::If(tmp, need, idx, sel, nsel_dev, (int)ntot, pred, st));
    c->launches += 2;
    int h[2] = {0, 0};
    WM_CUDA(cudaMemcpyAsync(h, nsel_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    WM_CUDA(cudaStreamSynchronize(st));
    nsel = h[0];
    if (nsel > 0) {
      k_count_below<<<std::min(wm_blocks(nsel, TPB), 148 * 8), TPB, 0, st>>>(sel, (int)nsel, c->n_sp0, nsel_dev + 1);
      WM_LAUNCH_CHECK(c);
      WM_CUDA(cudaMemcpyAsync(h + 1, nsel_dev + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
      WM_CUDA(cudaStreamSynchronize(st));
    }
    lcount[0] = h[1];
    lcount[1] = nsel - h[1];
    cudaFree(tmp);
    cudaFree(nsel_dev);
  } else if (mode == 1) {
    nsel = 0;
    lcount[0] = lcount[1] = 0;
  }
  int rc = WM_OK;
  if (buf && nsel > 0) {
    if (nsel > cap_records) {
      wm_set_error("wm_pack_particles: buffer too small for the selected particles");
      rc = WM_ERR_ARG;
    } else {
      // chunks of <= 4 Mi records through a device staging buffer
      const long long chunk = std::min<long long>(nsel, 4ll << 20);
      double* stage = nullptr;
      WM_CUDA(cudaMalloc(&stage, (size_t)chunk * g.ndim * sizeof(double)));
      for (long long first = 0; first < nsel; first += chunk) {
        const long long n = std::min(chunk, nsel - first);
        k_pack_records<<<std::min(wm_blocks(n, TPB), 148 * 16), TPB, 0, st>>>(g, c->A, c->id[c->cid], sel, first, n, stage, c->swap_yz ? 1 : 0);
        WM_LAUNCH_CHECK(c);
        WM_CUDA(cudaMemcpyAsync(buf + first * g.ndim, stage, (size_t)n * g.ndim * sizeof(double), cudaMemcpyDeviceToHost, st));
        WM_CUDA(cudaStreamSynchronize(st));
      }
      cudaFree(stage);
    }
  }
  if (sel) cudaFree(sel);
  return rc;
}
