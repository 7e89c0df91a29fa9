"""ctypes binding of include/wuming_b200.h with the reference's procedure names.

Host arrays follow the reference's layouts exactly (column-major):
  3-D  up/gp(7,np,nys:nye,nzs:nze,nsp)  uf(6,nxgs-2:nxge+2,nys-2:nye+2,nzs-2:nze+2)
       np2(nys:nye,nzs:nze,nsp)         cumcnt(nxgs:nxge+1,nys:nye,nzs:nze,nsp)
  2-D  up/gp(6,np,nys:nye,nsp)          uf(6,nxgs-2:nxge+2,nys-2:nye+2) ...
Any contiguous numpy array with that memory layout is accepted (Fortran-ordered with the
reference shape, or C-ordered with the reversed shape).
"""
import ctypes as C
import os

import numpy as np

from .mpi_set import SlabLayout, para_range  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

WM_BC_PERIODIC, WM_BC_RECONNECTION, WM_BC_SHOCK = 0, 1, 2
WM_ORDER_WEIBEL, WM_ORDER_RECONNECTION, WM_ORDER_SHOCK = 0, 1, 2

_ERR = {1: "WM_ERR_ARG", 2: "WM_ERR_CUDA", 3: "WM_ERR_CG_ITEMAX", 4: "WM_ERR_MEMORY_OVER",
        5: "WM_ERR_PARTICLE_LOST", 6: "WM_ERR_STATE"}


class WmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{_ERR.get(code, code)}: {msg}")
        self.code = code


class _Params(C.Structure):
    _fields_ = [("dim", C.c_int), ("ndim", C.c_int), ("np", C.c_int), ("nsp", C.c_int),
                ("nxgs", C.c_int), ("nxge", C.c_int), ("nygs", C.c_int), ("nyge", C.c_int),
                ("nzgs", C.c_int), ("nzge", C.c_int),
                ("nys", C.c_int), ("nye", C.c_int), ("nzs", C.c_int), ("nze", C.c_int),
                ("nproc_j", C.c_int), ("nproc_k", C.c_int), ("rank_j", C.c_int), ("rank_k", C.c_int),
                ("bc_kind", C.c_int), ("device", C.c_int),
                ("delx", C.c_double), ("delt", C.c_double), ("c", C.c_double), ("gfac", C.c_double),
                ("q", C.c_double * 2), ("r", C.c_double * 2)]


class _Stats(C.Structure):
    _fields_ = [("cg_iterations", C.c_int * 3), ("n_particles", C.c_longlong), ("max_np2", C.c_int),
                ("error_flags", C.c_int), ("timed_steps", C.c_int), ("ms_push", C.c_double), ("ms_deposit", C.c_double),
                ("ms_field", C.c_double), ("ms_sort", C.c_double)]


# every symbol include/wuming_b200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "wm_last_error", "wm_version", "wm_para_range", "wm_create", "wm_destroy", "wm_comm_unique_id", "wm_comm_init",
    "wm_upload", "wm_download", "wm_download_work", "wm_upload_work", "wm_particle_solv", "wm_field_fdtd_i", "wm_field_stage",
    "wm_bc_particle_x", "wm_bc_injection", "wm_bc_particle_yz", "wm_sort_bucket", "wm_step", "wm_set_fused",
    "wm_h_particle_solv", "wm_h_field_fdtd_i", "wm_h_step", "wm_load_weibel", "wm_energy", "wm_gauss",
    "wm_get_stats", "wm_sync", "wm_set_timing", "wm_launch_count", "wm_stream", "wm_mom_calc",
    "wm_particle_solv_vay", "wm_h_particle_solv_vay", "wm_set_pusher", "wm_shock_inject", "wm_shock_relocate", "wm_settle", "wm_pack_particles",
]


class ShockParams(C.Structure):
    """wm_shock_params of include/wuming_b200.h: constants of 2d/proj/shock/app.f90's inject() / relocate() / vprofile()"""
    _fields_ = [("n0", C.c_int), ("v0", C.c_double), ("v_thi", C.c_double), ("v_the", C.c_double), ("b0", C.c_double),
                ("theta_bn", C.c_double), ("phi_bn", C.c_double), ("l_damp_ini", C.c_double), ("seed", C.c_ulonglong)]


def library_path():
    # WM_B200_LIB: an alternative build of the same library (kernel-tuning experiments); never a different backend
    return os.environ.get("WM_B200_LIB") or os.path.join(_HERE, "lib", "libwuming_b200.so")


def load_library():
    """Load the CUDA backend.  Fails loudly if it has not been built (no fallback)."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C wumingpic_b200/csrc). There is no CPU fallback.")
        L = C.CDLL(path)
        dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
        L.wm_last_error.restype = C.c_char_p
        L.wm_create.argtypes = [C.POINTER(_Params), C.POINTER(vp)]
        L.wm_destroy.argtypes = [vp]
        L.wm_para_range.argtypes = [C.c_int] * 4 + [ip, ip]
        L.wm_comm_unique_id.argtypes = [C.c_char_p]
        L.wm_comm_init.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
        L.wm_upload.argtypes = [vp, dp, ip, ip, dp]
        L.wm_download.argtypes = [vp, dp, ip, ip, dp, dp]
        L.wm_download_work.argtypes = [vp, C.c_int, dp]
        L.wm_upload_work.argtypes = [vp, C.c_int, dp]
        for name in ("wm_particle_solv", "wm_particle_solv_vay", "wm_field_fdtd_i", "wm_bc_particle_x", "wm_sort_bucket"):
            getattr(L, name).argtypes = [vp, C.c_int, C.c_int]
        L.wm_field_stage.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.wm_bc_injection.argtypes = [vp, C.c_int, C.c_int, C.c_double]
        L.wm_bc_particle_yz.argtypes = [vp]
        L.wm_step.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]
        L.wm_set_fused.argtypes = [vp, C.c_int]
        L.wm_set_pusher.argtypes = [vp, C.c_int]
        L.wm_settle.argtypes = [vp]
        L.wm_pack_particles.argtypes = [vp, C.c_int, dp, C.c_longlong, C.POINTER(C.c_longlong)]
        lp = C.POINTER(C.c_longlong)
        L.wm_shock_inject.argtypes = [vp, C.POINTER(ShockParams), C.c_int, ip, lp, C.c_longlong]
        L.wm_shock_relocate.argtypes = [vp, C.POINTER(ShockParams), C.c_int, lp, C.c_longlong]
        L.wm_h_particle_solv_vay.argtypes = [vp, dp, dp, dp, ip, ip, C.c_int, C.c_int]
        L.wm_h_particle_solv.argtypes = [vp, dp, dp, dp, ip, ip, C.c_int, C.c_int]
        L.wm_h_field_fdtd_i.argtypes = [vp, dp, dp, dp, ip, ip, C.c_int, C.c_int]
        L.wm_h_step.argtypes = [vp, dp, dp, ip, ip, C.c_int, C.c_int, C.c_int, C.c_double]
        L.wm_load_weibel.argtypes = [vp, C.c_int] + [C.c_double] * 4 + [C.c_ulonglong]
        L.wm_energy.argtypes = [vp, dp]
        L.wm_mom_calc.argtypes = [vp, C.c_int, C.c_int, dp]
        L.wm_gauss.argtypes = [vp, dp]
        L.wm_get_stats.argtypes = [vp, C.POINTER(_Stats)]
        L.wm_sync.argtypes = [vp]
        L.wm_set_timing.argtypes = [vp, C.c_int]
        L.wm_launch_count.argtypes = [vp]
        L.wm_launch_count.restype = C.c_longlong
        L.wm_stream.argtypes = [vp]
        L.wm_stream.restype = C.c_void_p
        _LIB = L
    return _LIB


def weibel_constants(n0, mass_ratio=1.0, sigma_e=0.0, omega_pe=0.1, c=1.0):
    """q, r, b0 of the Weibel set-ups (3d/proj/weibel/app.f90:298-309)."""
    wpe = omega_pe
    wge = omega_pe * np.sqrt(sigma_e)
    wpi = wpe / np.sqrt(mass_ratio)
    wgi = wge / mass_ratio
    r = np.array([mass_ratio, 1.0])
    q = np.array([+np.sqrt(r[0] / (4.0 * np.pi * n0)) * wpi, -np.sqrt(r[1] / (4.0 * np.pi * n0)) * wpe])
    b0 = r[0] * c / q[0] * wgi
    return q, r, b0


def _dptr(a, size=None):
    if a is None:
        return None
    if a.dtype != np.float64 or not (a.flags.c_contiguous or a.flags.f_contiguous):
        raise TypeError("expected a contiguous float64 array")
    if size is not None and a.size != size:
        raise ValueError(f"array has {a.size} elements, the reference shape needs {size}")
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _iptr(a, size=None):
    if a is None:
        return None
    if a.dtype != np.int32 or not (a.flags.c_contiguous or a.flags.f_contiguous):
        raise TypeError("expected a contiguous int32 array")
    if size is not None and a.size != size:
        raise ValueError(f"array has {a.size} elements, the reference shape needs {size}")
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Backend:
    """One rank's simulation state on one GPU.

    The constructor takes the union of the reference's ``*__init`` arguments
    (3d/proj/weibel/app.f90:341-353); the methods carry the reference's procedure names and operate on
    the device-resident state, while ``upload``/``download`` are the explicit sync points
    (SURVEY.md 8b).  ``h_*`` methods take host arrays with the reference's own argument lists.
    """

    def __init__(self, dim, np_cap, nxgs, nxge, nygs, nyge, nzgs=0, nzge=0, nys=None, nye=None, nzs=None, nze=None,
                 delx=1.0, delt=1.0, c=1.0, gfac=0.501, q=(1.0, -1.0), r=(1.0, 1.0), bc_kind=WM_BC_PERIODIC,
                 nproc_j=1, nproc_k=1, rank_j=0, rank_k=0, device=-1):
        self.L = load_library()
        self.dim, self.ndim, self.nsp, self.np = dim, (7 if dim == 3 else 6), 2, np_cap
        nys = nygs if nys is None else nys
        nye = nyge if nye is None else nye
        nzs = nzgs if nzs is None else nzs
        nze = nzge if nze is None else nze
        p = _Params(dim=dim, ndim=self.ndim, np=np_cap, nsp=2, nxgs=nxgs, nxge=nxge, nygs=nygs, nyge=nyge, nzgs=nzgs,
                    nzge=nzge, nys=nys, nye=nye, nzs=nzs, nze=nze, nproc_j=nproc_j, nproc_k=nproc_k, rank_j=rank_j,
                    rank_k=rank_k, bc_kind=bc_kind, device=device, delx=delx, delt=delt, c=c, gfac=gfac)
        p.q[0], p.q[1], p.r[0], p.r[1] = q[0], q[1], r[0], r[1]
        self.prm = p
        self.nx, self.nyl = nxge - nxgs + 1, nye - nys + 1
        self.nzl = (nze - nzs + 1) if dim == 3 else 1
        self.npen = self.nsp * self.nyl * self.nzl
        self.nbox = (self.nx + 4) * (self.nyl + 4) * ((self.nzl + 4) if dim == 3 else 1)
        self.h = C.c_void_p()
        self._ck(self.L.wm_create(C.byref(p), C.byref(self.h)))

    # -- plumbing --------------------------------------------------------------------------------
    def _ck(self, code):
        if code != 0:
            raise WmError(code, self.L.wm_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.wm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shapes(self):
        """C-order shapes of the host arrays (the reference's Fortran shapes reversed)."""
        if self.dim == 3:
            return {"up": (self.nsp, self.nzl, self.nyl, self.np, self.ndim),
                    "uf": (self.nzl + 4, self.nyl + 4, self.nx + 4, 6),
                    "uj": (self.nzl + 4, self.nyl + 4, self.nx + 4, 3),
                    "gkl": (self.nzl, self.nyl, self.nx, 3),
                    "np2": (self.nsp, self.nzl, self.nyl),
                    "cumcnt": (self.nsp, self.nzl, self.nyl, self.nx + 1)}
        return {"up": (self.nsp, self.nyl, self.np, self.ndim),
                "uf": (self.nyl + 4, self.nx + 4, 6),
                "uj": (self.nyl + 4, self.nx + 4, 3),
                "gkl": (self.nyl, self.nx, 3),
                "np2": (self.nsp, self.nyl),
                "cumcnt": (self.nsp, self.nyl, self.nx + 1)}

    def empty(self, which):
        s = self.shapes()
        key = {"gp": "up", "df": "uf"}.get(which, which)
        return np.zeros(s[key], dtype=np.int32 if key in ("np2", "cumcnt") else np.float64)

    # -- multi-GPU -------------------------------------------------------------------------------
    def comm_unique_id(self):
        buf = C.create_string_buffer(128)
        self._ck(self.L.wm_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, nranks, rank, unique_id):
        self._ck(self.L.wm_comm_init(self.h, nranks, rank, unique_id))

    # -- sync points -----------------------------------------------------------------------------
    def upload(self, up=None, np2=None, cumcnt=None, uf=None):
        n = self.npen
        self._ck(self.L.wm_upload(self.h, _dptr(up, n * self.np * self.ndim if up is not None else None),
                                  _iptr(np2, n if np2 is not None else None),
                                  _iptr(cumcnt, n * (self.nx + 1) if cumcnt is not None else None),
                                  _dptr(uf, self.nbox * 6 if uf is not None else None)))

    def download(self, up=None, np2=None, cumcnt=None, uf=None, gp=None):
        self._ck(self.L.wm_download(self.h, _dptr(up), _iptr(np2), _iptr(cumcnt), _dptr(uf), _dptr(gp)))

    def download_work(self, which):
        idx = ("uj", "df", "gkl").index(which)
        out = self.empty(which)
        self._ck(self.L.wm_download_work(self.h, idx, _dptr(out)))
        return out

    def upload_work(self, which, arr):
        idx = ("uj", "df", "gkl").index(which)
        self._ck(self.L.wm_upload_work(self.h, idx, _dptr(arr, self.nbox * 6)))

    # -- the reference's procedures on resident state ------------------------------------------
    def particle__solv(self, nxs, nxe):
        self._ck(self.L.wm_particle_solv(self.h, nxs, nxe))

    def particle__solv_vay(self, nxs, nxe):
        """particle__solv_vay (3d/common/particle.f90:236-419): the Vay pusher on resident state."""
        self._ck(self.L.wm_particle_solv_vay(self.h, nxs, nxe))

    def settle(self):
        """apply the sort permutation wm_step left pending (asynchronous); implied by every call that reads the sorted set"""
        self._ck(self.L.wm_settle(self.h))

    def set_pusher(self, kind):
        """which pusher step()/h_step() run: WM_PUSHER_BORIS (0, particle__solv) or WM_PUSHER_VAY (1, particle__solv_vay)"""
        self._ck(self.L.wm_set_pusher(self.h, kind))

    def field__fdtd_i(self, nxs, nxe, stage=0):
        if stage:
            self._ck(self.L.wm_field_stage(self.h, nxs, nxe, stage))
        else:
            self._ck(self.L.wm_field_fdtd_i(self.h, nxs, nxe))

    def bc__particle_x(self, nxs, nxe):
        self._ck(self.L.wm_bc_particle_x(self.h, nxs, nxe))

    def bc__injection(self, nxs, nxe, u0):
        self._ck(self.L.wm_bc_injection(self.h, nxs, nxe, u0))

    def bc__particle_yz(self):
        self._ck(self.L.wm_bc_particle_yz(self.h))

    bc__particle_y = bc__particle_yz

    def sort__bucket(self, nxs, nxe):
        self._ck(self.L.wm_sort_bucket(self.h, nxs, nxe))

    def step(self, nxs, nxe, nsteps=1, order=WM_ORDER_WEIBEL, u0=0.0):
        self._ck(self.L.wm_step(self.h, nxs, nxe, order, u0, nsteps))

    def time_loop(self, nxs, nxe, nsteps=1, order=WM_ORDER_WEIBEL, u0=0.0, vay=False):
        """nsteps iterations of the reference drivers' time loop, procedure by procedure, in the order of the set-up's app.f90
        (Weibel 3d/proj/weibel/app.f90:100-108; reconnection 3d/proj/reconnection/app.f90:103-108; shock 2d/proj/shock/app.f90:112-118).
        On resident state these five calls run the same fused kernel + lazy sort as step() (wm_api.cu: deferred particle__solv)."""
        for _ in range(nsteps):
            (self.particle__solv_vay if vay else self.particle__solv)(nxs, nxe)
            if order == WM_ORDER_RECONNECTION:
                self.bc__particle_x(nxs, nxe)
            elif order == WM_ORDER_SHOCK:
                self.bc__injection(nxs, nxe, u0)
            self.field__fdtd_i(nxs, nxe)
            if order == WM_ORDER_WEIBEL:
                self.bc__particle_x(nxs, nxe)
            self.bc__particle_yz()
            self.sort__bucket(nxs, nxe)

    def set_fused(self, on=True):
        self._ck(self.L.wm_set_fused(self.h, 1 if on else 0))

    # -- host-buffer forms (the reference's own argument lists) -----------------------------------
    def h_particle__solv(self, gp, up, uf, cumcnt, np2, nxs, nxe):
        self._ck(self.L.wm_h_particle_solv(self.h, _dptr(gp), _dptr(up), _dptr(uf), _iptr(cumcnt), _iptr(np2), nxs, nxe))

    def h_field__fdtd_i(self, uf, up, gp, cumcnt, np2, nxs, nxe):
        self._ck(self.L.wm_h_field_fdtd_i(self.h, _dptr(uf), _dptr(up), _dptr(gp), _iptr(cumcnt), _iptr(np2), nxs, nxe))

    def h_step(self, up, uf, np2, cumcnt, nxs, nxe, order=WM_ORDER_WEIBEL, u0=0.0):
        self._ck(self.L.wm_h_step(self.h, _dptr(up), _dptr(uf), _iptr(np2), _iptr(cumcnt), nxs, nxe, order, u0))

    # -- synthetic load, diagnostics ---------------------------------------------------------------
    def load_weibel(self, n0, v_thi=0.1, v_the=0.1, t_ani=5.0, b0=0.0, seed=20240601):
        self._ck(self.L.wm_load_weibel(self.h, n0, v_thi, v_the, t_ani, b0, seed))

    def mom_calc(self, nxs, nxe):
        """mom_calc__accl + mom_calc__nvt + bc__mom on the device; returns mom with the reference's shape (C order reversed)"""
        shp = (self.nsp, self.nzl + 2, self.nyl + 2, self.nx + 2, 7) if self.dim == 3 else (self.nsp, self.nyl + 2, self.nx + 2, 7)
        out = np.zeros(shp)
        self._ck(self.L.wm_mom_calc(self.h, nxs, nxe, _dptr(out)))
        return out

    # -- the shock driver's particle source on the device (2d/proj/shock/app.f90:615-852) --------
    def shock_inject(self, prm, nxe, nlinj, id_first, epoch):
        """inject(): nlinj[row] particles per species behind every local pencil (row = (j - nys) + nyl (k - nzs));
        id_first[isp, row] = ncinj_grid(row) + nptotal(isp)"""
        nlinj = np.ascontiguousarray(nlinj, dtype=np.int32).ravel()
        id_first = np.ascontiguousarray(id_first, dtype=np.int64).ravel()
        assert nlinj.size == self.nyl * self.nzl and id_first.size == self.nsp * self.nyl * self.nzl
        self._ck(self.L.wm_shock_inject(self.h, C.byref(prm), nxe, nlinj.ctypes.data_as(C.POINTER(C.c_int)),
                                        id_first.ctypes.data_as(C.POINTER(C.c_longlong)), epoch))

    def shock_relocate(self, prm, nxe_new, id_first, epoch):
        """relocate() after nxe = nxe + 1: n0 particles per row and species in the new cell; id_first[isp, row] =
        global_row * n0 + nptotal(isp)"""
        id_first = np.ascontiguousarray(id_first, dtype=np.int64).ravel()
        assert id_first.size == self.nsp * self.nyl * self.nzl
        self._ck(self.L.wm_shock_relocate(self.h, C.byref(prm), nxe_new, id_first.ctypes.data_as(C.POINTER(C.c_longlong)), epoch))

    def pack_particles(self, mode):
        """paraio's get_particle_count (3d/common/paraio.f90:1007-1085): mode 0 all active particles, mode 1 tracers (ID > 0);
        returns (records[n, ndim], lcount[nsp]) packed species-major in pencil order"""
        lc = (C.c_longlong * 2)()
        self._ck(self.L.wm_pack_particles(self.h, mode, None, 0, lc))
        n = int(lc[0] + lc[1])
        buf = np.zeros((max(n, 1), self.ndim))
        if n:
            self._ck(self.L.wm_pack_particles(self.h, mode, _dptr(buf), n, lc))
        return buf[:n], np.array([lc[0], lc[1]], dtype=np.int64)

    def energy(self):
        out = np.zeros(4)
        self._ck(self.L.wm_energy(self.h, _dptr(out)))
        return out

    def gauss(self):
        out = np.zeros(2)
        self._ck(self.L.wm_gauss(self.h, _dptr(out)))
        return out[0], out[1]

    def stats(self):
        s = _Stats()
        self._ck(self.L.wm_get_stats(self.h, C.byref(s)))
        return {"cg_iterations": list(s.cg_iterations), "n_particles": s.n_particles, "max_np2": s.max_np2,
                "error_flags": s.error_flags, "timed_steps": s.timed_steps, "ms_push": s.ms_push, "ms_deposit": s.ms_deposit,
                "ms_field": s.ms_field, "ms_sort": s.ms_sort}

    def sync(self):
        self._ck(self.L.wm_sync(self.h))

    def set_timing(self, on=True):
        self._ck(self.L.wm_set_timing(self.h, 1 if on else 0))

    def launch_count(self):
        return self.L.wm_launch_count(self.h)

    def stream(self):
        """raw cudaStream_t of this context (wrap with torch.cuda.ExternalStream to record events on it)"""
        return self.L.wm_stream(self.h)
