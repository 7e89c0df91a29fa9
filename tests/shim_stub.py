"""The recording stand-in for libwuming_b200.so used by tests/test_shim_executed.py: oracle/_ref/libwm_stub.so forwards every
C-ABI call the translated Fortran shim makes to StubDevice.dispatch, where the CPU oracle plays the device-resident state.
TEST INFRASTRUCTURE: nothing here is on the product path (the product fails loudly without its CUDA library)."""
import ctypes as C

import numpy as np

from oracle.f2cxx import shim_harness
from oracle.pyoracle import World2, World3
from wumingpic_b200.backend import ShockParams, _Params

CB = C.CFUNCTYPE(C.c_int, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), C.POINTER(C.c_double))
_STATE = {}


def load_stub():
    """the stub library, loaded ONCE with RTLD_GLOBAL so that the wm_* symbols of the translated shim resolve to it"""
    if "lib" not in _STATE:
        L = C.CDLL(shim_harness.build_stub(), mode=C.RTLD_GLOBAL)
        L.stub_set_callback.argtypes = [CB]
        L.stub_set_error.argtypes = [C.c_char_p]
        _STATE["lib"] = L
    return _STATE["lib"]


class StubDevice:
    """one fake device: wm_create builds an oracle world from the wm_params the shim filled in; every later call is logged as
    (name, integer arguments, which pointer arguments were non-null) and carried out by the oracle"""

    def __init__(self):
        self.L = load_stub()
        self.log, self.worlds, self.prm, self.fail_next = [], {}, {}, None
        self._cb = CB(self.dispatch)
        self.L.stub_set_callback(self._cb)

    # ---- helpers
    def world(self, handle):
        return self.worlds[handle]

    @staticmethod
    def view(ptr, like):
        ct = C.c_double if like.dtype == np.float64 else (C.c_longlong if like.dtype == np.int64 else C.c_int)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=like.shape)

    def dispatch(self, name, p, i, d):
        name = name.decode()
        try:
            return getattr(self, "do_" + name)(p, i, d)
        except Exception as ex:  # noqa: BLE001  -- an exception must not unwind through the C frames
            self.L.stub_set_error(f"stub {name}: {type(ex).__name__}: {ex}".encode())
            return 6

    def maybe_fail(self, name):
        if self.fail_next and self.fail_next[0] == name:
            code, msg = self.fail_next[1:]
            self.fail_next = None
            self.L.stub_set_error(msg.encode())
            return code
        return 0

    # ---- the C ABI
    def do_wm_create(self, p, i, d):
        prm = C.cast(p[0], C.POINTER(_Params)).contents
        f = {k: getattr(prm, k) for k, _ in _Params._fields_ if k not in ("q", "r")}
        f["q"], f["r"] = list(prm.q), list(prm.r)
        self.log.append(("wm_create", dict(f)))
        nx, ny = f["nxge"] - f["nxgs"] + 1, f["nyge"] - f["nygs"] + 1
        kw = dict(delx=f["delx"], delt=f["delt"], c=f["c"], gfac=f["gfac"], q=f["q"], r=f["r"], bc=f["bc_kind"])
        if f["dim"] == 3:
            w = World3(nx, ny, f["nzge"] - f["nzgs"] + 1, f["np"], **kw)
        else:
            w = World2(nx, ny, f["np"], **kw)
        self.worlds[p[1]] = w
        self.prm[p[1]] = f
        return self.maybe_fail("wm_create")

    def do_wm_destroy(self, p, i, d):
        self.log.append(("wm_destroy",))
        self.worlds.pop(p[0]).close()
        return 0

    def do_wm_upload(self, p, i, d):
        w = self.world(p[0])
        given = []
        for k, which in enumerate(("up", "np2", "cumcnt", "uf"), 1):
            if p[k]:
                given.append(which)
                w.arr(which)[...] = self.view(p[k], w.arr(which))
        self.log.append(("wm_upload", tuple(given)))
        return self.maybe_fail("wm_upload")

    def do_wm_download(self, p, i, d):
        w = self.world(p[0])
        given = []
        for k, which in enumerate(("up", "np2", "cumcnt", "uf", "gp"), 1):
            if p[k]:
                given.append(which)
                self.view(p[k], w.arr(which))[...] = w.arr(which)
        self.log.append(("wm_download", tuple(given)))
        return 0

    def _ranged(self, name, p, i, fn):
        w = self.world(p[0])
        self.log.append((name, int(i[0]), int(i[1])))
        w.set_xrange(int(i[0]), int(i[1]))
        fn(w)
        if w.error():
            self.L.stub_set_error(b"memory over (np2 > np)")
            return 4
        return self.maybe_fail(name)

    def do_wm_particle_solv(self, p, i, d):
        return self._ranged("wm_particle_solv", p, i, lambda w: w.particle_solv())

    def do_wm_particle_solv_vay(self, p, i, d):
        return self._ranged("wm_particle_solv_vay", p, i, lambda w: w.particle_solv_vay())

    def do_wm_field_fdtd_i(self, p, i, d):
        return self._ranged("wm_field_fdtd_i", p, i, lambda w: w.field_fdtd_i())

    def do_wm_bc_particle_x(self, p, i, d):
        return self._ranged("wm_bc_particle_x", p, i, lambda w: w.bc_particle_x())

    def do_wm_sort_bucket(self, p, i, d):
        return self._ranged("wm_sort_bucket", p, i, lambda w: w.sort_bucket())

    def do_wm_bc_injection(self, p, i, d):
        u0 = float(d[0])
        self.log.append(("wm_bc_injection.u0", u0))
        return self._ranged("wm_bc_injection", p, i, lambda w: w.bc_injection(u0))

    def do_wm_bc_particle_yz(self, p, i, d):
        w = self.world(p[0])
        self.log.append(("wm_bc_particle_yz",))
        (w.bc_particle_yz if isinstance(w, World3) else w.bc_particle_y)()
        return 4 if w.error() else 0

    def do_wm_mom_calc(self, p, i, d):
        w = self.world(p[0])
        self.log.append(("wm_mom_calc", int(i[0]), int(i[1])))
        w.set_xrange(int(i[0]), int(i[1]))
        w.mom_calc()
        self.view(p[1], w.arr("mom"))[...] = w.arr("mom")
        return 0

    def do_wm_shock_inject(self, p, i, d):
        prm = C.cast(p[1], C.POINTER(ShockParams)).contents
        w = self.world(p[0])
        nrows = int(np.prod(w.arr("np2").shape[1:]))
        nl = np.array(self.view(p[2], np.zeros(nrows, np.int32)))
        ids = np.array(self.view(p[3], np.zeros(2 * nrows, np.int64)))
        self.log.append(("wm_shock_inject", int(i[0]), int(i[1]), {k: getattr(prm, k) for k, _ in ShockParams._fields_}, nl, ids))
        return 0

    def do_wm_shock_relocate(self, p, i, d):
        prm = C.cast(p[1], C.POINTER(ShockParams)).contents
        w = self.world(p[0])
        nrows = int(np.prod(w.arr("np2").shape[1:]))
        ids = np.array(self.view(p[2], np.zeros(2 * nrows, np.int64)))
        self.log.append(("wm_shock_relocate", int(i[0]), int(i[1]), {k: getattr(prm, k) for k, _ in ShockParams._fields_}, ids))
        return 0

    def do_wm_comm_unique_id(self, p, i, d):
        self.log.append(("wm_comm_unique_id",))
        return 0

    def do_wm_comm_init(self, p, i, d):
        self.log.append(("wm_comm_init", int(i[0]), int(i[1])))
        return 0

    def names(self):
        return [e[0] for e in self.log]


class MultiRankStubDevice(StubDevice):
    """N fake devices behind one callback: ONE oracle world with N emulated ranks plays all of them.  Uploads and downloads touch
    the calling rank's arrays; a compute call is collective -- it waits until every rank has made it (each from its own thread,
    like ranks of an MPI job), then the oracle does the work once."""

    def __init__(self, nranks):
        import threading
        super().__init__()
        self.nranks, self.barrier, self.lock = nranks, threading.Barrier(nranks), threading.Lock()
        self.w, self.rank_of, self.logs, self.ids = None, {}, [[] for _ in range(nranks)], {}

    def collective(self, fn):
        if self.barrier.wait(timeout=60) == 0:
            # the OpenMP team size is a per-thread setting: the oracle must run serially HERE (a rank's thread) for bit-exact sums
            from oracle import pyoracle
            pyoracle.set_num_threads(1)
            fn()
        self.barrier.wait(timeout=60)

    def do_wm_create(self, p, i, d):
        prm = C.cast(p[0], C.POINTER(_Params)).contents
        f = {k: getattr(prm, k) for k, _ in _Params._fields_ if k not in ("q", "r")}
        f["q"], f["r"] = list(prm.q), list(prm.r)
        rank = f["rank_j"] * f["nproc_k"] + f["rank_k"]
        with self.lock:
            if self.w is None:
                nx, ny, nz = f["nxge"] - f["nxgs"] + 1, f["nyge"] - f["nygs"] + 1, f["nzge"] - f["nzgs"] + 1
                kw = dict(delx=f["delx"], delt=f["delt"], c=f["c"], gfac=f["gfac"], q=f["q"], r=f["r"], bc=f["bc_kind"])
                self.w = World3(nx, ny, nz, f["np"], nproc_j=f["nproc_j"], nproc_k=f["nproc_k"], **kw) if f["dim"] == 3 else \
                    World2(nx, ny, f["np"], nproc=f["nproc_j"], **kw)
            self.rank_of[p[1]] = rank
            self.prm[rank] = f
        self.logs[rank].append(("wm_create",))
        return 0

    def _r(self, handle):
        return self.rank_of[handle]

    def do_wm_upload(self, p, i, d):
        rk = self._r(p[0])
        for k, which in enumerate(("up", "np2", "cumcnt", "uf"), 1):
            if p[k]:
                self.w.arr(which, rk)[...] = self.view(p[k], self.w.arr(which, rk))
        self.logs[rk].append(("wm_upload",))
        return 0

    def do_wm_download(self, p, i, d):
        rk = self._r(p[0])
        for k, which in enumerate(("up", "np2", "cumcnt", "uf", "gp"), 1):
            if p[k]:
                self.view(p[k], self.w.arr(which, rk))[...] = self.w.arr(which, rk)
        self.logs[rk].append(("wm_download",))
        return 0

    def _ranged(self, name, p, i, fn):
        rk = self._r(p[0])
        self.logs[rk].append((name, int(i[0]), int(i[1])))

        def work():
            self.w.set_xrange(int(i[0]), int(i[1]))
            fn(self.w)
        self.collective(work)
        return 4 if self.w.error() else 0

    def do_wm_bc_particle_yz(self, p, i, d):
        rk = self._r(p[0])
        self.logs[rk].append(("wm_bc_particle_yz",))
        self.collective(lambda: (self.w.bc_particle_yz if isinstance(self.w, World3) else self.w.bc_particle_y)())
        return 4 if self.w.error() else 0

    def do_wm_comm_unique_id(self, p, i, d):
        C.memmove(p[0], bytes(range(1, 129)), 128)               # what rank 0 draws; the shim hands it round with MPI_BCAST
        self.logs[0].append(("wm_comm_unique_id",))
        return 0

    def do_wm_comm_init(self, p, i, d):
        rk = self._r(p[0])
        self.ids[rk] = C.string_at(p[1], 128)
        self.logs[rk].append(("wm_comm_init", int(i[0]), int(i[1])))
        return 0
