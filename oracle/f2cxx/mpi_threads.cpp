// mpi_threads.cpp -- an in-process MPI for the translated reference (TEST INFRASTRUCTURE ONLY; built into
// oracle/_ref/libf2cxx_mpi.so by build_ref.py).
//
// One private copy of libwuming_ref{2,3}d*.so is one MPI rank (f90rt.cpp), every rank runs on its own host thread, and the two
// communication calls the reference's hot path makes -- MPI_SENDRECV with a neighbour (boundary_periodic.f90) and MPI_ALLREDUCE
// of one or two doubles (field.f90 cgm) -- rendezvous here without going through the Python interpreter: this is what lets
// `bench.py --impl reference` time the reference's own source as a flat-MPI job with one rank per host core.  The functions have
// the hook signatures of f90rt.h; the rank a call comes from is the rank its THREAD was bound to (f2mpi_bind).
//
// Like the shared-memory transports of real MPI libraries, waiting ranks poll (pause, then yield): a step of the 3-D loop makes
// ~ 300 blocking calls per rank (40 CG iterations with two all-reduces and a ghost exchange each), and a sleeping wait would bill
// the reference for the kernel's wake-up latency.  Every rank has its own inbox (a sender touches only its neighbour's), the
// all-reduce is one arrival counter and a generation word.
//
// MPI_ALLREDUCE sums in rank order (deterministic; pyref.py's Python transport and the oracle's emulation do the same).
// A rank that never arrives makes its peers give up after `timeout_s` with an exception, which the entry-point guard of the
// translated procedure turns into a recorded STOP -- the job fails, it does not hang.
#include <sched.h>

#include <atomic>
#include <chrono>
#include <cstring>
#include <deque>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

struct Msg {
  int src, tag;
  std::vector<char> data;
};

struct alignas(64) Inbox {
  std::mutex m;
  std::deque<Msg> q;
  std::atomic<long> posted{0};     // messages ever put here: a waiting receiver looks again only when this moves
};

struct alignas(64) Slot {
  double v[6];
};

struct Hub {
  int n = 1;
  double timeout_s = 120.0;
  std::vector<Inbox> inbox;
  std::vector<Slot> slot;           // all-reduce operands, one cache line per rank
  Slot total;
  alignas(64) std::atomic<int> arrived{0};
  alignas(64) std::atomic<long> generation{0};
  std::atomic<bool> aborted{false};
  std::atomic<long> n_sendrecv{0}, n_allreduce{0}, bytes{0};
  std::vector<char> bc_buf;         // MPI_BCAST
  alignas(64) std::atomic<int> bc_arrived{0};
  alignas(64) std::atomic<long> bc_generation{0};
  explicit Hub(int nranks) : n(nranks), inbox((size_t)nranks), slot((size_t)nranks) {}
};

thread_local Hub* t_hub = nullptr;
thread_local int t_rank = -1;

Hub& hub() {
  if (!t_hub || t_rank < 0) throw std::runtime_error("mpi_threads: this thread is not bound to a rank (f2mpi_bind)");
  return *t_hub;
}

inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#endif
}

// polls ready() : a few thousand pauses, then yields (more ranks than cores still make progress); gives up after timeout_s
template <class Pred>
void wait(Hub& h, Pred ready, const char* what) {
  std::chrono::steady_clock::time_point t0;
  bool timing = false;
  for (long spin = 0;; ++spin) {
    if (ready()) return;
    if (h.aborted.load(std::memory_order_relaxed)) throw std::runtime_error(std::string("mpi_threads: job aborted while waiting in ") + what);
    if (spin < 4000) {
      cpu_relax();
      continue;
    }
    sched_yield();
    if ((spin & 1023) == 0) {
      const auto now = std::chrono::steady_clock::now();
      if (!timing) {
        t0 = now;
        timing = true;
      } else if (std::chrono::duration<double>(now - t0).count() > h.timeout_s) {
        h.aborted.store(true);
        throw std::runtime_error(std::string("mpi_threads: timeout in ") + what + " (a peer rank never arrived)");
      }
    }
  }
}

}  // namespace

extern "C" {

void* f2mpi_create(int nranks, double timeout_s) {
  Hub* h = new Hub(nranks);
  h->timeout_s = timeout_s;
  return h;
}
void f2mpi_destroy(void* p) { delete (Hub*)p; }
void f2mpi_bind(void* p, int rank) {
  t_hub = (Hub*)p;
  t_rank = rank;
}
void f2mpi_abort(void* p) { ((Hub*)p)->aborted.store(true); }
int f2mpi_aborted(void* p) { return ((Hub*)p)->aborted.load() ? 1 : 0; }
// traffic counters since creation: out[0] = sendrecv calls, out[1] = allreduce calls, out[2] = payload bytes sent
void f2mpi_stats(void* p, long* out) {
  Hub* h = (Hub*)p;
  out[0] = h->n_sendrecv.load(), out[1] = h->n_allreduce.load(), out[2] = h->bytes.load();
}

// f90rt_sendrecv_fn.  Messages between one (source, destination, tag) triple are received in the order they were sent.
void f2mpi_sendrecv(const void* sbuf, int sbytes, int dest, int stag, void* rbuf, int rbytes, int src, int rtag) {
  Hub& h = hub();
  const int me = t_rank;
  if (dest < 0 || dest >= h.n || src < 0 || src >= h.n) throw std::runtime_error("MPI_SENDRECV: rank outside the communicator");
  if (sbytes < 0) sbytes = 0;
  {
    Msg out{me, stag, std::vector<char>((const char*)sbuf, (const char*)sbuf + sbytes)};
    Inbox& to = h.inbox[(size_t)dest];
    std::lock_guard<std::mutex> lk(to.m);
    to.q.push_back(std::move(out));
    to.posted.fetch_add(1, std::memory_order_release);
  }
  h.n_sendrecv.fetch_add(1, std::memory_order_relaxed);
  h.bytes.fetch_add(sbytes, std::memory_order_relaxed);
  Inbox& mine = h.inbox[(size_t)me];
  std::vector<char> in;
  long seen = -1;
  wait(h, [&] {
    const long posted = mine.posted.load(std::memory_order_acquire);
    if (posted == seen) return false;           // nothing new since the last look
    std::lock_guard<std::mutex> lk(mine.m);
    seen = mine.posted.load(std::memory_order_relaxed);
    for (auto it = mine.q.begin(); it != mine.q.end(); ++it)
      if (it->src == src && it->tag == rtag) {
        in = std::move(it->data);
        mine.q.erase(it);
        return true;
      }
    return false;
  }, "MPI_SENDRECV");
  if ((long)in.size() > (long)rbytes) throw std::runtime_error("MPI_SENDRECV: message longer than the receive buffer");
  if (!in.empty()) std::memcpy(rbuf, in.data(), in.size());
}

// f90rt_bcast_fn: the root's bytes reach every rank (two barriers around one shared buffer)
void f2mpi_bcast(void* buf, int bytes, int root) {
  Hub& h = hub();
  if (root < 0 || root >= h.n || bytes < 0) throw std::runtime_error("MPI_BCAST: bad root or count");
  auto barrier = [&](const char* what) {
    const long gen = h.bc_generation.load(std::memory_order_acquire);
    if (h.bc_arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == h.n) {
      h.bc_arrived.store(0, std::memory_order_relaxed);
      h.bc_generation.store(gen + 1, std::memory_order_release);
    } else {
      wait(h, [&] { return h.bc_generation.load(std::memory_order_acquire) != gen; }, what);
    }
  };
  if (t_rank == root) h.bc_buf.assign((const char*)buf, (const char*)buf + bytes);
  barrier("MPI_BCAST");                        // the root's data is in place
  if (t_rank != root) std::memcpy(buf, h.bc_buf.data(), (size_t)bytes);
  barrier("MPI_BCAST");                        // everybody has copied: the buffer may be reused
}

// f90rt_allreduce_fn: doubles, MPI_SUM (what cgm uses); the sum runs in rank order on the last rank to arrive
void f2mpi_allreduce(const void* sbuf, void* rbuf, int count, int type, int op) {
  Hub& h = hub();
  if (type != 8 || op != 1) throw std::runtime_error("MPI_ALLREDUCE: only MPI_DOUBLE_PRECISION with MPI_SUM is implemented");
  if (count < 0 || count > 6) throw std::runtime_error("MPI_ALLREDUCE: more than 6 elements (the hot path reduces 1 or 2)");
  std::memcpy(h.slot[(size_t)t_rank].v, sbuf, sizeof(double) * (size_t)count);
  h.n_allreduce.fetch_add(1, std::memory_order_relaxed);
  const long gen = h.generation.load(std::memory_order_acquire);
  if (h.arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == h.n) {
    for (int c = 0; c < count; ++c) {
      double t = h.slot[0].v[c];
      for (int r = 1; r < h.n; ++r) t = t + h.slot[(size_t)r].v[c];
      h.total.v[c] = t;
    }
    h.arrived.store(0, std::memory_order_relaxed);
    h.generation.store(gen + 1, std::memory_order_release);
  } else {
    wait(h, [&] { return h.generation.load(std::memory_order_acquire) != gen; }, "MPI_ALLREDUCE");
  }
  // nobody can overwrite `total` (or this rank's slot) before every rank has copied it: the next reduction completes only when
  // all ranks have arrived at it, and a rank arrives there after this copy
  std::memcpy(rbuf, h.total.v, sizeof(double) * (size_t)count);
}
}
