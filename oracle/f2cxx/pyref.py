"""pyref.py -- drives the translated reference (oracle/_ref/libwuming_ref{2,3}d.so, built by build_ref.py from the Fortran
sources with f2cxx.py) the way the reference's drivers do.  TEST INFRASTRUCTURE ONLY: only tests/ may import it.

One loaded copy of the library is one MPI rank: module variables and SAVEd locals are library-static, so every rank of an
emulated N-rank run loads its OWN private copy of the .so (a temporary file name per copy), runs on its own Python thread, and
the `use mpi` hooks of f90rt.h (MPI_SENDRECV, MPI_ALLREDUCE) rendezvous through queues in this file.  MPI_ALLREDUCE sums in rank
order, which is what the oracle's emulation does as well (a real MPI library may associate differently: round-off).

The call sequences are the drivers': `*__init` as in 3d/proj/weibel/app.f90:341-353 (2d :331-341), one step as in
3d/proj/weibel/app.f90:100-108, 2d/proj/reconnection/app.f90:99-106 (3d :101-107), 2d/proj/shock/app.f90:112-118 (3d :111-117).
"""
import ctypes as C
import os
import queue
import re
import shutil
import tempfile
import threading

import numpy as np

from . import build_ref

ORDER_WEIBEL, ORDER_RECONNECTION, ORDER_SHOCK = 0, 1, 2
_BC = {0: "boundary_periodic", 1: "boundary_reconnection", 2: "boundary_shock"}
MPI_INTEGER, MPI_DOUBLE, MPI_SUM = 4, 8, 1          # f90rt.h: a datatype handle is the element size


def available(dim):
    return build_ref.build(dim) is not None


def para_range(n1, n2, isize, irank):
    """3d/common/mpi_set.f90:81-94"""
    iwork1, iwork2 = (n2 - n1 + 1) // isize, (n2 - n1 + 1) % isize
    ns = irank * iwork1 + n1 + min(irank, iwork2)
    ne = ns + iwork1 - 1 + (1 if iwork2 > irank else 0)
    return ns, ne


class _Rank:
    """one private copy of the translated library + this rank's arrays"""

    def __init__(self, dim, path):
        fd, self.tmp = tempfile.mkstemp(prefix=f"wuming_ref{dim}d_", suffix=".so")
        os.close(fd)
        shutil.copyfile(path, self.tmp)
        self.L = C.CDLL(self.tmp)
        os.unlink(self.tmp)              # the mapping stays; nothing is left behind
        self.L.f90rt_last_stop.restype = C.c_char_p
        self.keep = []

    def call(self, name, *args):
        """every argument by reference: int / float -> temporaries, numpy arrays -> their data, ctypes functions -> the address"""
        conv, hold = [], []
        for a in args:
            if isinstance(a, (bool, int, np.integer)):
                v = C.c_int(int(a))
                hold.append(v)
                conv.append(C.byref(v))
            elif isinstance(a, (float, np.floating)):
                v = C.c_double(float(a))
                hold.append(v)
                conv.append(C.byref(v))
            elif isinstance(a, np.ndarray):
                assert a.flags["C_CONTIGUOUS"]
                conv.append(C.c_void_p(a.ctypes.data))
            elif isinstance(a, str):         # a procedure of the library passed as an actual argument
                conv.append(C.cast(getattr(self.L, a), C.c_void_p))
            else:
                raise TypeError(type(a))
        before = self.L.f90rt_stop_count()
        getattr(self.L, name)(*conv)
        if self.L.f90rt_stop_count() != before:
            raise RuntimeError(f"{name}: {self.L.f90rt_last_stop().decode()}")


class RefWorld:
    """nproc_j x nproc_k ranks of the translated reference (2-D: nproc_j y-slabs).  Arrays per rank have the reference's layout;
    numpy sees them in C order with the index order reversed, exactly like oracle.pyoracle.World2 / World3."""

    def __init__(self, dim, nx, ny, nz, np_cap, nproc_j=1, nproc_k=1, delx=1.0, delt=1.0, c=1.0, gfac=0.501, q=(1.0, -1.0),
                 r=(1.0, 1.0), bc=0, bounds=False, fast=False, native_mpi=False, lib=None, before_init=None):
        """fast: the -O3 -march=native build of the same generated C++ (timing only, never parity).  native_mpi: the ranks'
        MPI_SENDRECV / MPI_ALLREDUCE rendezvous in mpi_threads.cpp instead of Python callbacks (same semantics, no interpreter
        lock on the communication path: what a timed flat-MPI run with one rank per host thread needs)"""
        self.lib_override = lib
        # lib: another library with the reference's module interface -- the translated ISO_C_BINDING shim (shim_harness.py), whose
        # procedures forward to the C ABI: the same driver then runs the CUDA backend (or the recording stub) instead
        path = lib or (build_ref.build_fast(dim) if fast else build_ref.build(dim, bounds=bounds))
        if path is None:
            raise RuntimeError("the translated reference is not built and /root/reference is absent")
        self.native_mpi, self._hub, self._mpi = bool(native_mpi), None, None
        self.dim, self.nx, self.ny, self.nz, self.np, self.bc = dim, nx, ny, nz if dim == 3 else 1, np_cap, bc
        self.ndim, self.nsp = (7 if dim == 3 else 6), 2
        self.q, self.r = np.ascontiguousarray(q, np.float64), np.ascontiguousarray(r, np.float64)
        self.delx, self.delt, self.c, self.gfac = float(delx), float(delt), float(c), float(gfac)
        self.nproc_j, self.nproc_k = nproc_j, (nproc_k if dim == 3 else 1)
        self.nranks = self.nproc_j * self.nproc_k
        self.nxgs, self.nxge, self.nygs, self.nyge, self.nzgs, self.nzge = 2, nx + 1, 2, ny + 1, 2, self.nz + 1
        self.nxs, self.nxe = self.nxgs, self.nxge
        self.ranks, self.g, self.a = [], [], []
        self._mail = {}                       # (src, dst, tag) -> Queue
        self._mail_lock = threading.Lock()
        self._red = None
        for rk in range(self.nranks):
            R = _Rank(dim, path)
            rj, rkk = (rk // self.nproc_k, rk % self.nproc_k)          # rank = j * nproc_k + k (3d/common/mpi_set.f90:52-60)
            nys, nye = para_range(self.nygs, self.nyge, self.nproc_j, rj)
            nzs, nze = para_range(self.nzgs, self.nzge, self.nproc_k, rkk) if dim == 3 else (0, 0)
            tab = lambda j, k: (j % self.nproc_j) * self.nproc_k + (k % self.nproc_k)  # noqa: E731  periodic rank table
            g = dict(nys=nys, nye=nye, nzs=nzs, nze=nze, jup=tab(rj + 1, rkk), jdown=tab(rj - 1, rkk), kup=tab(rj, rkk + 1),
                     kdown=tab(rj, rkk - 1))
            nyl, nzl = nye - nys + 1, nze - nzs + 1
            if dim == 3:
                shp = dict(up=(2, nzl, nyl, np_cap, 7), uf=(nzl + 4, nyl + 4, nx + 4, 6), np2=(2, nzl, nyl),
                           cumcnt=(2, nzl, nyl, nx + 1), mom=(2, nzl + 2, nyl + 2, nx + 2, 7))
            else:
                shp = dict(up=(2, nyl, np_cap, 6), uf=(nyl + 4, nx + 4, 6), np2=(2, nyl), cumcnt=(2, nyl, nx + 1),
                           mom=(2, nyl + 2, nx + 2, 7))
            a = dict(up=np.zeros(shp["up"]), gp=np.zeros(shp["up"]), uf=np.zeros(shp["uf"]), mom=np.zeros(shp["mom"]),
                     np2=np.zeros(shp["np2"], np.int32), cumcnt=np.zeros(shp["cumcnt"], np.int32))
            self.ranks.append(R)
            self.g.append(g)
            self.a.append(a)
        if self.nranks > 1:
            self._install_transport()
        if before_init is not None:          # what a driver does between mpi_set__init and the __init calls (the shim's
            self._all(lambda rk: before_init(rk, self.ranks[rk]))          # wm_shim_comm_init), on every rank
        self._init_modules()

    # ---- the MPI library of the emulated ranks ---------------------------------------------------------------------
    def _box(self, key):
        with self._mail_lock:
            if key not in self._mail:
                self._mail[key] = queue.Queue()
            return self._mail[key]

    def _install_transport(self):
        if self.native_mpi:
            M = self._mpi = C.CDLL(build_ref.build_mpi())
            M.f2mpi_create.restype = C.c_void_p
            M.f2mpi_create.argtypes = [C.c_int, C.c_double]
            for f in (M.f2mpi_destroy, M.f2mpi_abort, M.f2mpi_aborted):
                f.argtypes = [C.c_void_p]
            M.f2mpi_bind.argtypes = [C.c_void_p, C.c_int]
            M.f2mpi_stats.argtypes = [C.c_void_p, C.POINTER(C.c_long)]
            self._hub = M.f2mpi_create(self.nranks, 120.0)
            for R in self.ranks:
                R.L.f90rt_set_transport.argtypes = [C.c_void_p, C.c_void_p]
                R.L.f90rt_set_transport(C.cast(M.f2mpi_sendrecv, C.c_void_p), C.cast(M.f2mpi_allreduce, C.c_void_p))
                R.L.f90rt_set_bcast.argtypes = [C.c_void_p]
                R.L.f90rt_set_bcast(C.cast(M.f2mpi_bcast, C.c_void_p))
            return
        SR = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int)
        AR = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int)
        self._barrier = threading.Barrier(self.nranks)
        self._red_slots = [None] * self.nranks
        for me, R in enumerate(self.ranks):
            def sendrecv(sbuf, sbytes, dest, stag, rbuf, rbytes, src, rtag, me=me):
                self._box((me, dest, stag)).put(C.string_at(sbuf, sbytes) if sbytes else b"")
                msg = self._box((src, me, rtag)).get(timeout=120)
                if len(msg) > rbytes:
                    raise RuntimeError("MPI_SENDRECV: message longer than the receive buffer")
                if msg:
                    C.memmove(rbuf, msg, len(msg))

            def allreduce(sbuf, rbuf, count, typ, op, me=me):
                assert typ == MPI_DOUBLE and op == MPI_SUM
                self._red_slots[me] = np.frombuffer(C.string_at(sbuf, 8 * count), np.float64).copy()
                self._barrier.wait(timeout=120)
                tot = self._red_slots[0].copy()
                for v in self._red_slots[1:]:            # rank order, like the oracle's emulation
                    tot = tot + v
                C.memmove(rbuf, tot.ctypes.data, 8 * count)
                self._barrier.wait(timeout=120)

            R.keep += [SR(sendrecv), AR(allreduce)]
            R.L.f90rt_set_transport(R.keep[-2], R.keep[-1])

    def pin_ranks(self, cpus):
        """timed runs: rank r's threads run on host CPU cpus[r % len(cpus)] (like `mpiexec --bind-to core`), so that the memory a
        rank touches first -- seed its arrays from inside _all() -- is local to the core that works on it"""
        self._pin = list(cpus)

    def _all(self, fn):
        """run fn(rank) on every rank -- concurrently when there is more than one (they exchange messages)"""
        pin = getattr(self, "_pin", None)
        if self.nranks == 1 and not pin:
            fn(0)
            return
        err = []

        def body(rk):
            try:
                if pin:
                    try:
                        os.sched_setaffinity(0, {pin[rk % len(pin)]})       # 0 = the calling thread
                    except OSError:
                        pass
                if self._hub:
                    self._mpi.f2mpi_bind(self._hub, rk)       # the rank of a native MPI call is the rank its thread is bound to
                fn(rk)
            except BaseException as e:  # noqa: BLE001
                err.append(e)
                if self._hub:
                    self._mpi.f2mpi_abort(self._hub)
                elif getattr(self, "_barrier", None) is not None:
                    self._barrier.abort()

        th = [threading.Thread(target=body, args=(rk,)) for rk in range(self.nranks)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if err:
            raise err[0]

    # ---- the drivers' call sequences -------------------------------------------------------------------------------
    def _geom_args(self, rk):
        g = self.g[rk]
        if self.dim == 3:
            return [self.nxgs, self.nxge, self.nygs, self.nyge, self.nzgs, self.nzge, g["nys"], g["nye"], g["nzs"], g["nze"]]
        return [self.nxgs, self.nxge, self.nygs, self.nyge, g["nys"], g["nye"]]

    def _init_modules(self):
        """every rank on its own thread, like ranks of an MPI job: the reference's __init procedures are local, but the shim's last
        __init creates the device context and connects the communicator -- a collective (MPI_BCAST of the NCCL id)"""
        def one(rk):
            R, nstat = self.ranks[rk], np.zeros(6, np.int32)
            g, head = self.g[rk], [self.ndim, self.np, self.nsp] + self._geom_args(rk)
            nb = [g["jup"], g["jdown"], g["kup"], g["kdown"]] if self.dim == 3 else [g["jup"], g["jdown"]]
            # bc__init(..., jup, jdown, kup, kdown, mnpi, mnpr, ncomw, nerr, nstat, delx, delt, c) + the hidden extent of nstat(:)
            R.call(_BC[self.bc] + "__init", *head, *nb, MPI_INTEGER, MPI_DOUBLE, 0, 0, nstat, self.delx, self.delt, self.c,
                   len(nstat))
            R.call("particle__init", *head, self.delx, self.delt, self.c, self.q, self.r)
            R.call("field__init", *head, MPI_DOUBLE, 0, MPI_SUM, 0, self.delx, self.delt, self.c, self.q, self.r, self.gfac)
            R.call("sort__init", *head)
            R.call("mom_calc__init", *head, self.delx, self.delt, self.c, self.q, self.r)
        self._all(one)

    def arr(self, which, rank=0):
        return self.a[rank][which]

    def saved(self, proc, name, rank=0, dtype=np.float64):
        """numpy view (index order reversed, like arr()) of a SAVEd allocatable local of a translated procedure, e.g.
        saved("field__fdtd_i", "df"): the CG warm start; None before the procedure's first call allocated it"""
        f = getattr(self.ranks[rank].L, f"f2cxx_saved__{proc}__{name}")
        f.restype = C.c_void_p
        b = (C.c_long * 16)()
        p = f(b)
        if not p:
            return None
        ext = []
        for d in range(8):
            if b[2 * d + 1] == 0 and b[2 * d] == 0 and d > 0 and all(v == 0 for v in b[2 * d:]):
                break
            ext.append(b[2 * d + 1] - b[2 * d] + 1)
        n = int(np.prod(ext))
        ct = C.c_double if dtype == np.float64 else C.c_int
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(n,)).reshape(ext[::-1])

    def geom(self, rank=0):
        return dict(self.g[rank])

    def set_xrange(self, nxs, nxe):
        self.nxs, self.nxe = nxs, nxe

    def _bcname(self, what):
        return f"{_BC[self.bc]}__{what}"

    def particle_solv(self, vay=False):
        self._all(lambda rk: self.ranks[rk].call("particle__solv_vay" if vay else "particle__solv", self.a[rk]["gp"], self.a[rk]["up"],
                                                 self.a[rk]["uf"], self.a[rk]["cumcnt"], self.nxs, self.nxe))

    def bc_particle_x(self):
        extra = [] if self.bc == 0 else [self.nxs, self.nxe]
        self._all(lambda rk: self.ranks[rk].call(self._bcname("particle_x"), self.a[rk]["gp"], self.a[rk]["np2"], *extra))

    def bc_injection(self, u0):
        self._all(lambda rk: self.ranks[rk].call(self._bcname("injection"), self.a[rk]["gp"], self.a[rk]["np2"], self.nxs, self.nxe,
                                                 float(u0)))

    def field_fdtd_i(self):
        self._all(lambda rk: self.ranks[rk].call("field__fdtd_i", self.a[rk]["uf"], self.a[rk]["up"], self.a[rk]["gp"],
                                                 self.a[rk]["cumcnt"], self.nxs, self.nxe, self._bcname("dfield"),
                                                 self._bcname("curre"), self._bcname("phi")))

    def bc_particle_yz(self):
        name = self._bcname("particle_yz" if self.dim == 3 else "particle_y")
        self._all(lambda rk: self.ranks[rk].call(name, self.a[rk]["gp"], self.a[rk]["np2"]))

    def sort_bucket(self):
        # sort__bucket(gp, up, ...) : dummy names swap -- the first argument is the OUTPUT (the drivers pass `up, gp`)
        self._all(lambda rk: self.ranks[rk].call("sort__bucket", self.a[rk]["up"], self.a[rk]["gp"], self.a[rk]["cumcnt"],
                                                 self.a[rk]["np2"], self.nxs, self.nxe))

    def step(self, order=ORDER_WEIBEL, u0=0.0, vay=False):
        self.particle_solv(vay)
        if order == ORDER_RECONNECTION:
            self.bc_particle_x()
        elif order == ORDER_SHOCK:
            self.bc_injection(u0)
        self.field_fdtd_i()
        if order == ORDER_WEIBEL:
            self.bc_particle_x()
        self.bc_particle_yz()
        self.sort_bucket()

    def run_steps(self, n, order=ORDER_WEIBEL, u0=0.0, vay=False):
        """n whole steps with ONE host thread per rank for the whole run (step() starts a thread per rank and procedure): every rank
        goes through the driver's call sequence on its own and meets its neighbours only inside the MPI calls, as ranks of an MPI
        job do -- the form the timed CPU baseline uses"""
        def loop(rk):
            a, R = self.a[rk], self.ranks[rk]
            extra = [] if self.bc == 0 else [self.nxs, self.nxe]
            yz = self._bcname("particle_yz" if self.dim == 3 else "particle_y")
            for _ in range(n):
                R.call("particle__solv_vay" if vay else "particle__solv", a["gp"], a["up"], a["uf"], a["cumcnt"], self.nxs, self.nxe)
                if order == ORDER_RECONNECTION:
                    R.call(self._bcname("particle_x"), a["gp"], a["np2"], *extra)
                elif order == ORDER_SHOCK:
                    R.call(self._bcname("injection"), a["gp"], a["np2"], self.nxs, self.nxe, float(u0))
                R.call("field__fdtd_i", a["uf"], a["up"], a["gp"], a["cumcnt"], self.nxs, self.nxe, self._bcname("dfield"),
                       self._bcname("curre"), self._bcname("phi"))
                if order == ORDER_WEIBEL:
                    R.call(self._bcname("particle_x"), a["gp"], a["np2"], *extra)
                R.call(yz, a["gp"], a["np2"])
                R.call("sort__bucket", a["up"], a["gp"], a["cumcnt"], a["np2"], self.nxs, self.nxe)
        self._all(loop)

    def mpi_stats(self):
        """(MPI_SENDRECV calls, MPI_ALLREDUCE calls, payload bytes) of the native transport since the world was made"""
        if not self._hub:
            return None
        out = (C.c_long * 3)()
        self._mpi.f2mpi_stats(self._hub, out)
        return tuple(out)

    def close(self):
        if self._hub:
            self._mpi.f2mpi_destroy(self._hub)
            self._hub = None

    def mom_calc(self):
        """mom_calc__accl + mom_calc__nvt + bc__mom as the drivers call them (3d/proj/weibel/app.f90:121-124)"""
        def one(rk):
            a, R = self.a[rk], self.ranks[rk]
            R.call("mom_calc__accl", a["gp"], a["up"], a["uf"], a["cumcnt"], self.nxs, self.nxe)
            R.call("mom_calc__nvt", a["mom"], a["gp"], a["np2"])
            R.call(self._bcname("mom"), a["mom"])
        self._all(one)


# ------------------------------------------------------------------------------------------------------------------------------
# the drivers' own procedures (proj/*/app.f90), assembled and translated by app_harness.py
# ------------------------------------------------------------------------------------------------------------------------------
class RefApp:
    """one rank of a driver's module `app`: harness__configure / harness__init wrap the statement blocks of load_config / init,
    every other procedure (set_initial_condition, inject, relocate, energy_history, ...) is the reference's text.  The random
    inputs (`uniform_rand`, `normal_rand`, `shuffle` of utils/wuming_utils.f90) are handed out IN CALL ORDER from the sequences
    given to feed(); running out of values raises."""

    def __init__(self, name, path=None):
        """path: another library assembled around the same driver (mainloop_harness.build_full: the whole init() + main loop on top
        of the ISO_C_BINDING shim) -- same module `app`, same configuration and random-input hooks"""
        from . import app_harness
        path = path or app_harness.build(name)
        if path is None:
            raise RuntimeError("the translated driver procedures are not built and /root/reference is absent")
        self.name, self.cfg = name, app_harness.APPS[name]
        self.R = _Rank(2 if "2d" in name else 3, path)
        self.L = self.R.L
        self._u, self._n, self._perm = [], [], []
        self.calls = {"uniform": 0, "normal": 0, "shuffle": 0}
        RAND, SHUF = C.CFUNCTYPE(C.c_double), C.CFUNCTYPE(None, C.POINTER(C.c_int), C.c_int)

        def uniform():
            self.calls["uniform"] += 1
            return self._pop(self._u, "uniform_rand")

        def normal():
            self.calls["normal"] += 1
            return self._pop(self._n, "normal_rand")

        def shuffle(a, n):
            self.calls["shuffle"] += 1
            perm = self._pop(self._perm, "shuffle")
            vals = [a[i] for i in range(n)]
            assert sorted(perm) == list(range(n))
            for i in range(n):
                a[i] = vals[perm[i]]

        self.R.keep += [RAND(uniform), RAND(normal), SHUF(shuffle)]
        self.L.f90rt_set_random(*self.R.keep[-3:])

    @staticmethod
    def _pop(seq, what):
        if not seq:
            raise RuntimeError(f"{what}() called more often than the test provided values")   # -> std::terminate is avoided: ctypes prints it
        return seq.pop(0)

    def feed(self, uniform=(), normal=(), shuffles=()):
        self._u, self._n, self._perm = list(map(float, uniform)), list(map(float, normal)), [list(p) for p in shuffles]

    def leftover(self):
        return len(self._u), len(self._n), len(self._perm)

    def call(self, name, *args):
        self.R.call(name, *args)

    def scalar(self, name, ctype=None):
        """a ctypes view of a module scalar (read .value, assign .value)"""
        f = getattr(self.L, f"f2cxx_modvar__app__{name}")
        f.restype = C.c_void_p
        if ctype is None:
            ctype = C.c_int if re.match(r"(n|i|mpi)", name) else C.c_double
        return ctype.from_address(f())

    def array(self, name, dtype=np.float64):
        """numpy view (index order reversed) of a module array; None while unallocated"""
        f = getattr(self.L, f"f2cxx_modarr__app__{name}")
        f.restype = C.c_void_p
        b = (C.c_long * 16)()
        p = f(b)
        if not p:
            return None
        rank = {"np2": self.R_dim(), "cumcnt": self.R_dim() + 1, "uf": self.R_dim() + 1, "up": self.R_dim() + 2, "gp": self.R_dim() + 2,
                "mom": self.R_dim() + 2, "r": 1, "q": 1}[name]
        ext = [b[2 * d + 1] - b[2 * d] + 1 for d in range(rank)]
        ct = C.c_double if dtype == np.float64 else C.c_int
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(int(np.prod(ext)),)).reshape(ext[::-1])

    def R_dim(self):
        return 2 if "2d" in self.name else 3

    def configure(self, rank_args, **params):
        """harness__configure(<the "parameter" section of config.json>, nrank, nys, nye[, nzs, nze, nrank_j, nrank_k])"""
        vals = []
        for c, t in self.cfg["config"]:
            v = params[c]
            vals.append(int(v) if t == "i" else float(v))
        self.call("harness__configure", *vals, *[int(v) for v in rank_args])
