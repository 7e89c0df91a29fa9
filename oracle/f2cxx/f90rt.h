// f90rt.h -- run-time support for the C++ that oracle/f2cxx/f2cxx.py generates from the reference's Fortran sources.
// TEST INFRASTRUCTURE ONLY.  Column-major arrays with declared lower bounds (bounds-checked with -DF90_BOUNDS), the handful of
// intrinsics the hot path uses with Fortran's typing rules, STOP as an exception caught at the entry point, and the transport
// hooks behind `use mpi` (one rank: a copy; N ranks: the test driver installs a rendezvous, oracle/f2cxx/pyref.py).
#pragma once
#include <cfenv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace f90 {

struct Stop {
  std::string where;
};

struct B {
  long lo, hi;
};

template <class T, int R>
struct Arr {
  typedef T value_type;
  T* p = nullptr;
  long lo_[R], n_[R], st_[R];
  std::vector<T> own;

  Arr() {
    for (int d = 0; d < R; ++d) lo_[d] = 1, n_[d] = 0, st_[d] = 0;
  }
  Arr(T* ptr, const B (&b)[R]) { set(ptr, b); }
  Arr(const Arr&) = delete;
  Arr& operator=(const Arr&) = delete;

  void set(T* ptr, const B (&b)[R]) {
    long tot = 1;
    for (int d = 0; d < R; ++d) {
      lo_[d] = b[d].lo;
      n_[d] = b[d].hi - b[d].lo + 1;
      if (n_[d] < 0) n_[d] = 0;
      st_[d] = tot;
      tot *= n_[d];
    }
    if (ptr) {
      p = ptr;
    } else {                      // automatic / allocated storage (Fortran leaves it undefined; zero here)
      own.assign((size_t)tot, T());
      p = own.data();
    }
  }
  void allocate(const B (&b)[R]) {
    if (p) throw std::runtime_error("allocate of an allocated array");
    set(nullptr, b);
    if (!p) {                     // zero-sized: still "allocated"
      own.reserve(1);
      p = own.data();
    }
  }
  void deallocate() {
    if (!p) throw std::runtime_error("deallocate of an unallocated array");
    std::vector<T>().swap(own);
    p = nullptr;
    for (int d = 0; d < R; ++d) n_[d] = 0;
  }
  long lb(int d) const { return lo_[d]; }
  long ub(int d) const { return lo_[d] + n_[d] - 1; }
  long extent(int d) const { return n_[d]; }
  long size() const {
    long t = 1;
    for (int d = 0; d < R; ++d) t *= n_[d];
    return t;
  }
  T* data() const { return p; }

  template <class... I>
  T& operator()(I... idx) const {
    static_assert(sizeof...(I) == R, "rank mismatch");
    const long ii[R] = {(long)idx...};
    long off = 0;
    for (int d = 0; d < R; ++d) {
#ifdef F90_BOUNDS
      if (ii[d] < lo_[d] || ii[d] >= lo_[d] + n_[d]) {
        char msg[160];
        std::snprintf(msg, sizeof msg, "subscript %d = %ld outside %ld:%ld", d + 1, ii[d], lo_[d], lo_[d] + n_[d] - 1);
        throw std::out_of_range(msg);
      }
#endif
      off += (ii[d] - lo_[d]) * st_[d];
    }
    return p[off];
  }
};

template <class T>
struct Tmp {
  T v;
  T* ptr() { return &v; }
};
template <class T>
inline Tmp<T> tmp(T v) {
  return Tmp<T>{v};
}

inline long trip_count(long lo, long hi, long st) {
  if (st == 0) throw std::runtime_error("do loop with zero step");
  long n = (hi - lo + st) / st;
  return n > 0 ? n : 0;
}

// ---- intrinsics (generic names resolve on the argument type, as in Fortran) ----
inline int int_(int x) { return x; }
inline int int_(long long x) { return (int)x; }
inline int int_(float x) { return (int)x; }          // truncation toward zero
inline int int_(double x) { return (int)x; }
inline float sqrt_(float x) { return std::sqrt(x); }
inline double sqrt_(double x) { return std::sqrt(x); }
inline int abs_(int x) { return x < 0 ? -x : x; }
inline long long abs_(long long x) { return x < 0 ? -x : x; }
inline float abs_(float x) { return std::fabs(x); }
inline double abs_(double x) { return std::fabs(x); }
inline float atan_(float x) { return std::atan(x); }
inline double atan_(double x) { return std::atan(x); }
inline double exp_(double x) { return std::exp(x); }
inline float exp_(float x) { return std::exp(x); }
inline double log_(double x) { return std::log(x); }
inline float log_(float x) { return std::log(x); }
inline double sin_(double x) { return std::sin(x); }
inline float sin_(float x) { return std::sin(x); }
inline double cos_(double x) { return std::cos(x); }
inline float cos_(float x) { return std::cos(x); }
inline double tanh_(double x) { return std::tanh(x); }
inline float tanh_(float x) { return std::tanh(x); }
inline double cosh_(double x) { return std::cosh(x); }
inline float cosh_(float x) { return std::cosh(x); }
inline int floor_(float x) { return (int)std::floor(x); }
inline int floor_(double x) { return (int)std::floor(x); }
inline int nint_(double x) { return (int)std::lround(x); }
inline double atan2_(double y, double x) { return std::atan2(y, x); }

template <class A, class Bt>
inline typename std::common_type<A, Bt>::type max_(A a, Bt b) {
  typedef typename std::common_type<A, Bt>::type C;
  return (C)a > (C)b ? (C)a : (C)b;
}
template <class A, class Bt, class... Rest>
inline auto max_(A a, Bt b, Rest... r) -> decltype(max_(max_(a, b), r...)) {
  return max_(max_(a, b), r...);
}
template <class A, class Bt>
inline typename std::common_type<A, Bt>::type min_(A a, Bt b) {
  typedef typename std::common_type<A, Bt>::type C;
  return (C)a < (C)b ? (C)a : (C)b;
}
template <class A, class Bt, class... Rest>
inline auto min_(A a, Bt b, Rest... r) -> decltype(min_(min_(a, b), r...)) {
  return min_(min_(a, b), r...);
}
inline int mod_(int a, int b) { return a % b; }       // sign of the dividend, like Fortran's MOD
inline double mod_(double a, double b) { return std::fmod(a, b); }
inline double sign_(double a, double b) { return std::signbit(b) ? -std::fabs(a) : std::fabs(a); }
inline int sign_(int a, int b) { return b < 0 ? -abs_(a) : abs_(a); }
inline long long sign_(long long a, long long b) { return b < 0 ? -abs_(a) : abs_(a); }
// TRANSFER(source, mold) between the two 8-byte kinds the reference uses it for (the particle ID bit-cast into a real(8) slot)
inline double transfer_double(long long v) {
  double d;
  std::memcpy(&d, &v, 8);
  return d;
}
inline double transfer_double(double v) { return v; }
inline long long transfer_i64(double v) {
  long long i;
  std::memcpy(&i, &v, 8);
  return i;
}
inline long long transfer_i64(long long v) { return v; }

// x ** n with an integer exponent: repeated multiplication by squaring, the expansion gfortran emits for small constant
// exponents (x**2 = x*x, x**3 = (x*x)*x, x**4 = (x*x)*(x*x)); real exponents go through pow
template <class T>
inline T powi(T x, long n) {
  if (n < 0) return (T)1 / powi(x, -n);
  if (n == 0) return (T)1;
  // left-to-right binary method: ((x*x)*x) for 3, ((x*x)*(x*x)) for 4
  int top = 0;
  for (long m = n; m > 1; m >>= 1) ++top;
  T r = x;
  for (int b = top - 1; b >= 0; --b) {
    r = r * r;
    if ((n >> b) & 1) r = r * x;
  }
  return r;
}
inline double pow_(double x, int n) { return powi<double>(x, n); }
inline float pow_(float x, int n) { return powi<float>(x, n); }
inline int pow_(int x, int n) { return powi<int>(x, n); }
inline double pow_(double x, double y) { return std::pow(x, y); }
inline double pow_(double x, float y) { return std::pow(x, (double)y); }
inline double pow_(float x, double y) { return std::pow((double)x, y); }
inline float pow_(float x, float y) { return std::pow(x, y); }
inline double pow_(int x, double y) { return std::pow((double)x, y); }

// gcc does not model the rounding mode as state (no FENV_ACCESS): -frounding-math stops constant folding, and the memory clobber
// keeps every value that is LOADED after the call from being computed before it -- which covers the reference's use (the
// operands of its ieee_down sections are particle coordinates read from `up` inside the loops that follow the call)
inline void set_rounding(int mode) {
  asm volatile("" ::: "memory");
  std::fesetround(mode);
  asm volatile("" ::: "memory");
}
struct RoundingScope {              // IEEE modes are restored when the procedure that changed them returns (F2003 14.4)
  int saved;
  RoundingScope() : saved(std::fegetround()) {}
  ~RoundingScope() { std::fesetround(saved); }
};

inline void message(const char* s) { std::fprintf(stderr, "[f2cxx] %s\n", s); }
[[noreturn]] inline void stop(const char* where) { throw Stop{where}; }

}  // namespace f90

// ---- entry-point guards: STOP and run-time errors must not unwind into ctypes ----
extern "C" {
int f90rt_stop_count();
const char* f90rt_last_stop();
void f90rt_note_stop(const char* what);
}
namespace f90 {
// nesting of translated procedures: only the outermost one swallows a STOP.  Internal linkage on purpose: an `inline` variable
// is a GNU-unique symbol, i.e. ONE object shared by all the private copies of the library that emulate the MPI ranks
static thread_local int g_depth = 0;
struct EntryGuard {
  EntryGuard() { ++g_depth; }
  ~EntryGuard() { --g_depth; }
  bool nested() const { return g_depth > 1; }
};
}  // namespace f90
#define F90_ENTRY_BEGIN   \
  f90::EntryGuard _guard; \
  try {
#define F90_ENTRY_END                                                                      \
  }                                                                                        \
  catch (const f90::Stop& s) {                                                             \
    if (_guard.nested()) throw;                                                            \
    f90rt_note_stop(("STOP at " + s.where).c_str());                                       \
  }                                                                                        \
  catch (const std::exception& e) {                                                        \
    if (_guard.nested()) throw;                                                            \
    f90rt_note_stop((std::string("run-time error: ") + e.what()).c_str());                 \
  }

// ---- `use mpi` ----
// Datatypes are passed around as the integers the driver hands to the *__init routines (mnpi, mnpr, opsum, ncomw); the stub
// gives them these values:
#define F90RT_MPI_INTEGER 4           /* = bytes per element */
#define F90RT_MPI_DOUBLE 8
#define F90RT_MPI_SUM 1
extern "C" {
// hooks a multi-rank driver installs (nullptr = single rank: sendrecv with oneself is a copy, allreduce the identity)
typedef void (*f90rt_sendrecv_fn)(const void* sbuf, int sbytes, int dest, int stag, void* rbuf, int rbytes, int src, int rtag);
typedef void (*f90rt_allreduce_fn)(const void* sbuf, void* rbuf, int count, int type, int op);
void f90rt_set_transport(f90rt_sendrecv_fn sr, f90rt_allreduce_fn ar);
typedef void (*f90rt_bcast_fn)(void* buf, int bytes, int root);
void f90rt_set_bcast(f90rt_bcast_fn bc);          // MPI_BCAST of a multi-rank run (the shim hands the NCCL id round with it)
void f90rt_mpi_sendrecv(const void* sbuf, int* scount, int* stype, int* dest, int* stag, void* rbuf, int* rcount, int* rtype,
                        int* src, int* rtag, int* comm, int* status, int* ierr);
void f90rt_mpi_allreduce(const void* sbuf, void* rbuf, int* count, int* type, int* op, int* comm, int* ierr);
void f90rt_mpi_barrier(int* comm, int* ierr);
void f90rt_mpi_bcast(void* buf, int* count, int* type, int* root, int* comm, int* ierr);
void f90rt_mpi_allgather(const void* sbuf, int* scount, int* stype, void* rbuf, int* rcount, int* rtype, int* comm, int* ierr);
void f90rt_mpi_reduce(const void* sbuf, void* rbuf, int* count, int* type, int* op, int* root, int* comm, int* ierr);
void f90rt_mpi_finalize(int* ierr);
// inputs the test driver provides in place of the reference's random_number consumers (utils/wuming_utils.f90: uniform_rand,
// normal_rand, shuffle): values are handed out in call order
typedef double (*f90rt_rand_fn)();
typedef void (*f90rt_shuffle_fn)(int* a, int n);
void f90rt_set_random(f90rt_rand_fn uniform, f90rt_rand_fn normal, f90rt_shuffle_fn shuffle);
double f90rt_uniform_rand();
double f90rt_normal_rand();
void f90rt_shuffle(int* a, int* n);
// records written by `write(unit, fmt) ...` to data files (energy.dat): the values, in order
void f90rt_capture(int n, const double* v);
int f90rt_captured(double* out, int max);       // -> number of values; clears the buffer
void f90rt_set_rounding_nearest();
// `wm_check` of the ISO_C_BINDING shim (fortran/wuming_b200_c.f90), provided natively by shim_rt.cpp: STOP with the C-ABI
// library's message when ierr is not 0
void f90rt_wm_check(int* ierr, const char* where);
}
