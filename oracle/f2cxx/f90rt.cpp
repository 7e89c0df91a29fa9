// f90rt.cpp -- the non-inline part of the run-time behind the translated reference (TEST INFRASTRUCTURE ONLY): STOP
// bookkeeping and the `use mpi` transport.  One copy of this file is linked into every translated library, so every loaded copy
// of the library (= one emulated MPI rank) has its own state.
#include "f90rt.h"

static int g_stops = 0;
static std::string g_last;
static f90rt_sendrecv_fn g_sr = nullptr;
static f90rt_allreduce_fn g_ar = nullptr;

extern "C" {

int f90rt_stop_count() { return g_stops; }
const char* f90rt_last_stop() { return g_last.c_str(); }
void f90rt_note_stop(const char* what) {
  ++g_stops;
  g_last = what;
  std::fprintf(stderr, "[f2cxx] %s\n", what);
  std::fesetround(FE_TONEAREST);      // a STOP inside an ieee_down section must not leak the rounding mode into the caller
}
void f90rt_set_transport(f90rt_sendrecv_fn sr, f90rt_allreduce_fn ar) {
  g_sr = sr;
  g_ar = ar;
}
void f90rt_set_rounding_nearest() { std::fesetround(FE_TONEAREST); }

// MPI_SENDRECV(sendbuf, sendcount, sendtype, dest, sendtag, recvbuf, recvcount, recvtype, source, recvtag, comm, status, ierr)
void f90rt_mpi_sendrecv(const void* sbuf, int* scount, int* stype, int* dest, int* stag, void* rbuf, int* rcount, int* rtype,
                        int* src, int* rtag, int*, int*, int* ierr) {
  const int sb = *scount * *stype, rb = *rcount * *rtype;      // the datatype handle IS the element size (f90rt.h)
  if (g_sr) {
    g_sr(sbuf, sb, *dest, *stag, rbuf, rb, *src, *rtag);
  } else {
    // one rank: every neighbour is this rank; the message it sends to itself arrives in its own receive buffer
    if (sb > rb) throw std::runtime_error("MPI_SENDRECV: message longer than the receive buffer");
    std::memmove(rbuf, sbuf, (size_t)sb);
  }
  if (ierr) *ierr = 0;
}

// MPI_ALLREDUCE(sendbuf, recvbuf, count, datatype, op, comm, ierr)
void f90rt_mpi_allreduce(const void* sbuf, void* rbuf, int* count, int* type, int* op, int*, int* ierr) {
  if (g_ar)
    g_ar(sbuf, rbuf, *count, *type, *op);
  else
    std::memmove(rbuf, sbuf, (size_t)(*count * *type));
  if (ierr) *ierr = 0;
}

void f90rt_mpi_barrier(int*, int* ierr) {
  if (ierr) *ierr = 0;
}
// the driver-level collectives are used by single-rank tests only: one rank gathers / reduces / broadcasts to itself
static void need_one_rank(const char* what) {
  if (g_sr) throw std::runtime_error(std::string(what) + " is implemented for one rank only");
}
static f90rt_bcast_fn g_bc = nullptr;
void f90rt_set_bcast(f90rt_bcast_fn bc) { g_bc = bc; }
// MPI_BCAST(buffer, count, datatype, root, comm, ierr)
void f90rt_mpi_bcast(void* buf, int* count, int* type, int* root, int*, int* ierr) {
  if (g_bc)
    g_bc(buf, *count * *type, *root);
  else
    need_one_rank("MPI_BCAST");
  if (ierr) *ierr = 0;
}
void f90rt_mpi_allgather(const void* sbuf, int* scount, int* stype, void* rbuf, int*, int*, int*, int* ierr) {
  need_one_rank("MPI_ALLGATHER");
  std::memmove(rbuf, sbuf, (size_t)(*scount * *stype));
  if (ierr) *ierr = 0;
}
void f90rt_mpi_reduce(const void* sbuf, void* rbuf, int* count, int* type, int*, int*, int*, int* ierr) {
  need_one_rank("MPI_REDUCE");
  std::memmove(rbuf, sbuf, (size_t)(*count * *type));
  if (ierr) *ierr = 0;
}
void f90rt_mpi_finalize(int* ierr) {
  if (ierr) *ierr = 0;
}

static f90rt_rand_fn g_uniform = nullptr, g_normal = nullptr;
static f90rt_shuffle_fn g_shuffle = nullptr;
void f90rt_set_random(f90rt_rand_fn u, f90rt_rand_fn n, f90rt_shuffle_fn s) {
  g_uniform = u;
  g_normal = n;
  g_shuffle = s;
}
double f90rt_uniform_rand() {
  if (!g_uniform) throw std::runtime_error("uniform_rand() called but the test driver provided no stream");
  return g_uniform();
}
double f90rt_normal_rand() {
  if (!g_normal) throw std::runtime_error("normal_rand() called but the test driver provided no stream");
  return g_normal();
}
void f90rt_shuffle(int* a, int* n) {
  if (!g_shuffle) throw std::runtime_error("shuffle() called but the test driver provided no permutation");
  g_shuffle(a, *n);
}

static std::vector<double> g_cap;
void f90rt_capture(int n, const double* v) { g_cap.insert(g_cap.end(), v, v + n); }
int f90rt_captured(double* out, int max) {
  int n = (int)g_cap.size() < max ? (int)g_cap.size() : max;
  for (int i = 0; i < n; ++i) out[i] = g_cap[i];
  g_cap.clear();
  return n;
}
}
