// stub_cabi.cpp -- a stand-in for libwuming_b200.so on machines without a GPU (TEST INFRASTRUCTURE ONLY; built into
// oracle/_ref/libwm_stub.so by shim_harness.py).  It defines the entry points of include/wuming_b200.h that the Fortran shim
// calls -- compiled AGAINST that header, so a signature that drifts from the ABI does not compile -- and forwards every call to a
// callback installed by the test (tests/test_shim_executed.py), where the CPU oracle plays the device.  What this checks is the
// shim: which C functions it calls, in which order, with which arrays, shapes and scalars, and what it hands back to the driver.
// It is never loaded by the product, and it computes nothing itself.
#include <cstring>
#include <string>

#include "wuming_b200.h"

extern "C" {
typedef int (*stub_callback)(const char* name, void** ptrs, const long long* ints, const double* dbls);
}
static stub_callback g_cb = nullptr;
static std::string g_err = "";
static int g_handle_storage[64];
static int g_handles = 0;

static int fwd(const char* name, void** p, const long long* i, const double* d) {
  if (!g_cb) {
    g_err = "stub: no callback installed";
    return WM_ERR_STATE;
  }
  return g_cb(name, p, i, d);
}

extern "C" {
void stub_set_callback(stub_callback cb) { g_cb = cb; }
void stub_set_error(const char* msg) { g_err = msg ? msg : ""; }
int stub_sizeof_params(void) { return (int)sizeof(wm_params); }
int stub_sizeof_shock_params(void) { return (int)sizeof(wm_shock_params); }

const char* wm_last_error(void) { return g_err.c_str(); }

int wm_create(const wm_params* prm, wm_ctx** out) {
  if (g_handles >= 64) return WM_ERR_STATE;
  *out = (wm_ctx*)&g_handle_storage[g_handles++];          // an opaque, non-null handle; the callback keys its state on it
  void* p[] = {(void*)prm, (void*)*out};
  return fwd("wm_create", p, nullptr, nullptr);
}
int wm_destroy(wm_ctx* c) {
  void* p[] = {c};
  return fwd("wm_destroy", p, nullptr, nullptr);
}
int wm_comm_unique_id(char* id) {
  std::memset(id, 0, 128);
  void* p[] = {id};
  return fwd("wm_comm_unique_id", p, nullptr, nullptr);
}
int wm_comm_init(wm_ctx* c, int nranks, int rank, const char* id) {
  void* p[] = {c, (void*)id};
  long long i[] = {nranks, rank};
  return fwd("wm_comm_init", p, i, nullptr);
}
int wm_upload(wm_ctx* c, const double* up, const int* np2, const int* cumcnt, const double* uf) {
  void* p[] = {c, (void*)up, (void*)np2, (void*)cumcnt, (void*)uf};
  return fwd("wm_upload", p, nullptr, nullptr);
}
int wm_download(wm_ctx* c, double* up, int* np2, int* cumcnt, double* uf, double* gp) {
  void* p[] = {c, up, np2, cumcnt, uf, gp};
  return fwd("wm_download", p, nullptr, nullptr);
}
#define RANGE_CALL(fn)                         \
  int fn(wm_ctx* c, int nxs, int nxe) {        \
    void* p[] = {c};                           \
    long long i[] = {nxs, nxe};                \
    return fwd(#fn, p, i, nullptr);            \
  }
RANGE_CALL(wm_particle_solv)
RANGE_CALL(wm_particle_solv_vay)
RANGE_CALL(wm_field_fdtd_i)
RANGE_CALL(wm_bc_particle_x)
RANGE_CALL(wm_sort_bucket)
int wm_bc_injection(wm_ctx* c, int nxs, int nxe, double u0) {
  void* p[] = {c};
  long long i[] = {nxs, nxe};
  double d[] = {u0};
  return fwd("wm_bc_injection", p, i, d);
}
int wm_bc_particle_yz(wm_ctx* c) {
  void* p[] = {c};
  return fwd("wm_bc_particle_yz", p, nullptr, nullptr);
}
int wm_mom_calc(wm_ctx* c, int nxs, int nxe, double* mom) {
  void* p[] = {c, mom};
  long long i[] = {nxs, nxe};
  return fwd("wm_mom_calc", p, i, nullptr);
}
int wm_shock_inject(wm_ctx* c, const wm_shock_params* prm, int nxe, const int* nlinj, const long long* id_first, long long epoch) {
  void* p[] = {c, (void*)prm, (void*)nlinj, (void*)id_first};
  long long i[] = {nxe, epoch};
  return fwd("wm_shock_inject", p, i, nullptr);
}
int wm_shock_relocate(wm_ctx* c, const wm_shock_params* prm, int nxe_new, const long long* id_first, long long epoch) {
  void* p[] = {c, (void*)prm, (void*)id_first};
  long long i[] = {nxe_new, epoch};
  return fwd("wm_shock_relocate", p, i, nullptr);
}
}
