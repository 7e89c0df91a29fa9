"""ctypes view of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product path (wumingpic_b200) never does.
Pinned bit for bit to the reference's own source, translated (oracle/f2cxx -> oracle/_ref) -- see oracle/oracle_common.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
class ShockPrm(C.Structure):
    """orc::ShockPrm (oracle_common.h) = wm_shock_params (include/wuming_b200.h), field for field"""
    _fields_ = [("n0", C.c_int), ("v0", C.c_double), ("v_thi", C.c_double), ("v_the", C.c_double), ("b0", C.c_double),
                ("theta_bn", C.c_double), ("phi_bn", C.c_double), ("l_damp_ini", C.c_double), ("seed", C.c_uint64)]


_LIBS = {}


def _host_cpu_tag():
    """the CPU feature flags of this host: liboracle_fast.so is built with -march=native, so a copy built on another machine (the
    .so travels from the build container to the GPU box) must be rebuilt when the instruction sets differ"""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    import hashlib
                    return hashlib.sha1(" ".join(sorted(line.split(":", 1)[1].split())).encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def build(fast=False):
    target = "liboracle_fast.so" if fast else "liboracle.so"
    if fast:
        subprocess.run(["make", "-s", "-B", "-C", _HERE, target], check=True)
        with open(os.path.join(_HERE, "liboracle_fast.host"), "w") as f:
            f.write(_host_cpu_tag())
    else:
        subprocess.run(["make", "-s", "-C", _HERE, target], check=True)
    return os.path.join(_HERE, target)


def lib(fast=False):
    key = bool(fast)
    if key not in _LIBS:
        path = os.path.join(_HERE, "liboracle_fast.so" if fast else "liboracle.so")
        src_newer = (not os.path.exists(path)) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(path)
            for f in ("oracle3d.cpp", "oracle2d.cpp", "oracle_common.h"))
        if fast and not src_newer:
            try:
                with open(os.path.join(_HERE, "liboracle_fast.host")) as f:
                    src_newer = f.read().strip() != _host_cpu_tag()
            except OSError:
                src_newer = True
        if src_newer:
            path = build(fast)
        L = C.CDLL(path)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        for dim in ("3", "2"):
            if not hasattr(L, f"orc{dim}_create"):
                continue
            getattr(L, f"orc{dim}_create").restype = C.c_void_p
            getattr(L, f"orc{dim}_dptr").restype = dp
            getattr(L, f"orc{dim}_iptr").restype = ip
        L.orc3_create.argtypes = [C.c_int] * 6 + [C.c_double] * 4 + [dp, dp, C.c_int]
        L.orc3_dptr.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc3_iptr.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc3_load_weibel.argtypes = [C.c_void_p, C.c_int] + [C.c_double] * 4 + [C.c_uint64]
        for name in ("destroy", "particle_solv", "particle_solv_vay", "bc_particle_x", "bc_particle_yz", "sort_bucket", "step",
                     "clear_error"):
            getattr(L, "orc3_" + name).argtypes = [C.c_void_p]
        L.orc3_field_fdtd_i.argtypes = [C.c_void_p, C.c_int]
        L.orc3_mom_calc.argtypes = [C.c_void_p]
        for d in (2, 3):
            getattr(L, f"orc{d}_pack_particles").argtypes = [C.c_void_p, C.c_int, C.c_int, dp, C.POINTER(C.c_longlong)]
            getattr(L, f"orc{d}_pack_particles").restype = C.c_longlong
            getattr(L, f"orc{d}_shock_inject").argtypes = [C.c_void_p, C.POINTER(ShockPrm), ip, C.c_uint]
            getattr(L, f"orc{d}_shock_relocate").argtypes = [C.c_void_p, C.POINTER(ShockPrm), C.c_uint]
            getattr(L, f"orc{d}_nxe").argtypes = [C.c_void_p]
        L.orc3_set_pusher.argtypes = [C.c_void_p, C.c_int]
        L.orc2_set_pusher.argtypes = [C.c_void_p, C.c_int]
        L.orc2_mom_calc.argtypes = [C.c_void_p]
        L.orc3_step_order.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.orc3_bc_particle_x_reflect.argtypes = [C.c_void_p]
        L.orc3_bc_injection.argtypes = [C.c_void_p, C.c_double]
        L.orc3_nranks.argtypes = [C.c_void_p]
        L.orc3_error.argtypes = [C.c_void_p]
        L.orc3_rank_geom.argtypes = [C.c_void_p, C.c_int, ip]
        L.orc3_set_xrange.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc3_cg_iterations.argtypes = [C.c_void_p, ip]
        L.orc3_energy.argtypes = [C.c_void_p, dp]
        L.orc3_gauss.argtypes = [C.c_void_p, C.c_int, dp]
        L.orc2_create.argtypes = [C.c_int] * 4 + [C.c_double] * 4 + [dp, dp, C.c_int]
        L.orc2_dptr.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc2_iptr.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc2_load_weibel.argtypes = [C.c_void_p, C.c_int] + [C.c_double] * 4 + [C.c_uint64]
        for name in ("destroy", "particle_solv", "particle_solv_vay", "bc_particle_y", "sort_bucket", "clear_error"):
            getattr(L, "orc2_" + name).argtypes = [C.c_void_p]
        L.orc2_bc_particle_x.argtypes = [C.c_void_p, C.c_int]
        L.orc2_bc_injection.argtypes = [C.c_void_p, C.c_double]
        L.orc2_step.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.orc2_field_fdtd_i.argtypes = [C.c_void_p, C.c_int]
        L.orc2_nranks.argtypes = [C.c_void_p]
        L.orc2_error.argtypes = [C.c_void_p]
        L.orc2_rank_geom.argtypes = [C.c_void_p, C.c_int, ip]
        L.orc2_set_xrange.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc2_cg_iterations.argtypes = [C.c_void_p, ip]
        L.orc2_energy.argtypes = [C.c_void_p, dp]
        L.orc2_gauss.argtypes = [C.c_void_p, C.c_int, dp]
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_philox_uniform2.argtypes = [C.c_ulonglong, C.c_uint, C.c_uint, C.c_uint, C.c_uint, dp]
        L.orc_box_muller.argtypes = [C.c_double, C.c_double, dp]
        _LIBS[key] = L
    return _LIBS[key]


def set_num_threads(n, fast=False):
    """size the OpenMP team of the oracle library explicitly and return the team a parallel region really gets"""
    L = lib(fast)
    L.orc_set_num_threads(int(n))
    return L.orc_team_size()


def num_threads(fast=False):
    return lib(fast).orc_num_threads()


def philox_uniform2(seed, stream, idx, purpose, epoch=0):
    """the two uniforms of the oracle's keyed stream (oracle_common.h Philox::uniform2)"""
    out = (C.c_double * 2)()
    lib().orc_philox_uniform2(seed, stream, idx, purpose, epoch, out)
    return out[0], out[1]


def box_muller(x1, x2):
    """(rr sin, rr cos) in the reference's form (utils/wuming_utils.f90:72-90), evaluated by the oracle's libm"""
    out = (C.c_double * 2)()
    lib().orc_box_muller(x1, x2, out)
    return out[0], out[1]


def keyed_normals(seed, row, ii, isp, base, epoch=0):
    """the three normal deviates the oracle gives particle (row, ii) of species isp (oracle_common.h shock_velocity / the loaders)"""
    a0, a1 = philox_uniform2(seed, row, ii, base + 2 * isp - 1, epoch)
    b0, b1 = philox_uniform2(seed, row, ii, base + 2 * isp, epoch)
    ns, nc = box_muller(a0, a1)
    ms, _ = box_muller(b0, b1)
    return ns, nc, ms


def weibel_constants(n0, mass_ratio=1.0, sigma_e=0.0, omega_pe=0.1, c=1.0):
    """q, r, b0 as in 3d/proj/weibel/app.f90:298-309."""
    wpe = omega_pe
    wge = omega_pe * np.sqrt(sigma_e)
    wpi = wpe / np.sqrt(mass_ratio)
    wgi = wge / mass_ratio
    r = np.array([mass_ratio, 1.0])
    q = np.array([+np.sqrt(r[0] / (4.0 * np.pi * n0)) * wpi, -np.sqrt(r[1] / (4.0 * np.pi * n0)) * wpe])
    b0 = r[0] * c / q[0] * wgi
    return q, r, b0


class World3:
    """In-process emulation of an nproc_j x nproc_k MPI run of the 3-D reference loop."""

    def __init__(self, nx, ny, nz, np_cap, nproc_j=1, nproc_k=1, delx=1.0, delt=1.0, c=1.0, gfac=0.501,
                 q=(1.0, -1.0), r=(1.0, 1.0), bc=0, fast=False):
        self.L = lib(fast)
        self.nx, self.ny, self.nz, self.np, self.ndim, self.nsp, self.bc = nx, ny, nz, np_cap, 7, 2, bc
        self.q = np.ascontiguousarray(q, dtype=np.float64)
        self.r = np.ascontiguousarray(r, dtype=np.float64)
        self.delx, self.delt, self.c, self.gfac = delx, delt, c, gfac
        dp = C.POINTER(C.c_double)
        self.h = C.c_void_p(self.L.orc3_create(nx, ny, nz, np_cap, nproc_j, nproc_k, delx, delt, c, gfac,
                                               self.q.ctypes.data_as(dp), self.r.ctypes.data_as(dp), bc))
        self.nranks = self.L.orc3_nranks(self.h)

    def close(self):
        if self.h:
            self.L.orc3_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def geom(self, rank=0):
        g = (C.c_int * 8)()
        self.L.orc3_rank_geom(self.h, rank, g)
        return dict(zip(("nys", "nye", "nzs", "nze", "jup", "jdown", "kup", "kdown"), list(g)))

    # numpy views in Fortran index order (first index fastest): shapes reversed for C order
    def _shape(self, rank, which):
        g = self.geom(rank)
        nyl, nzl = g["nye"] - g["nys"] + 1, g["nze"] - g["nzs"] + 1
        if which in ("up", "gp"):
            return (self.nsp, nzl, nyl, self.np, self.ndim)
        if which in ("uf", "df"):
            return (nzl + 4, nyl + 4, self.nx + 4, 6)
        if which == "uj":
            return (nzl + 4, nyl + 4, self.nx + 4, 3)
        if which == "gkl":
            return (nzl, nyl, self.nx, 3)
        if which == "mom":
            return (self.nsp, nzl + 2, nyl + 2, self.nx + 2, 7)
        if which == "np2":
            return (self.nsp, nzl, nyl)
        if which == "cumcnt":
            return (self.nsp, nzl, nyl, self.nx + 1)
        raise KeyError(which)

    def arr(self, which, rank=0):
        shape = self._shape(rank, which)
        if which in ("np2", "cumcnt"):
            p = self.L.orc3_iptr(self.h, rank, 0 if which == "np2" else 1)
        else:
            p = self.L.orc3_dptr(self.h, rank, ("up", "gp", "uf", "df", "uj", "gkl", "mom").index(which))
        return np.ctypeslib.as_array(p, shape=shape)

    def load_weibel(self, n0, v_thi=0.1, v_the=0.1, t_ani=5.0, b0=0.0, seed=20240601):
        self.L.orc3_load_weibel(self.h, n0, v_thi, v_the, t_ani, b0, seed)

    def particle_solv(self):
        self.L.orc3_particle_solv(self.h)

    def particle_solv_vay(self):
        self.L.orc3_particle_solv_vay(self.h)

    def set_pusher(self, kind):
        """the pusher step() calls: 0 particle__solv (Buneman-Boris), 1 particle__solv_vay"""
        self.L.orc3_set_pusher(self.h, kind)

    def shock_inject(self, prm, nlinj_rows, epoch):
        """inject() of 3d/proj/shock/app.f90:733-906; nlinj_rows per GLOBAL row (k - nzgs) ny + (j - nygs)"""
        a = np.ascontiguousarray(nlinj_rows, dtype=np.int32)
        self.L.orc3_shock_inject(self.h, C.byref(prm), a.ctypes.data_as(C.POINTER(C.c_int)), epoch)

    def shock_relocate(self, prm, epoch):
        self.L.orc3_shock_relocate(self.h, C.byref(prm), epoch)

    def pack_particles(self, mode, rank=0):
        """get_particle_count of paraio (3d/common/paraio.f90:1007-1085): (records[n, ndim], lcount[nsp])"""
        lc = (C.c_longlong * 2)()
        n = self.L.orc3_pack_particles(self.h, rank, mode, None, lc)
        buf = np.zeros((max(n, 1), self.ndim))
        self.L.orc3_pack_particles(self.h, rank, mode, buf.ctypes.data_as(C.POINTER(C.c_double)), lc)
        return buf[:n], np.array([lc[0], lc[1]], dtype=np.int64)

    @property
    def nxe_now(self):
        return self.L.orc3_nxe(self.h)

    def field_fdtd_i(self, stage=0):
        self.L.orc3_field_fdtd_i(self.h, stage)

    def set_xrange(self, nxs, nxe):
        self.L.orc3_set_xrange(self.h, nxs, nxe)

    def bc_particle_x(self, kind=None):
        if (0 if self.bc == 0 else 1) if kind is None else kind:
            self.L.orc3_bc_particle_x_reflect(self.h)
        else:
            self.L.orc3_bc_particle_x(self.h)

    def bc_injection(self, u0):
        self.L.orc3_bc_injection(self.h, u0)

    def bc_particle_yz(self):
        self.L.orc3_bc_particle_yz(self.h)

    def sort_bucket(self):
        self.L.orc3_sort_bucket(self.h)

    def mom_calc(self):
        """mom_calc__accl + mom_calc__nvt + bc__mom (3d/proj/weibel/app.f90:121-124); overwrites gp like the reference"""
        self.L.orc3_mom_calc(self.h)

    def step(self, order=0, u0=0.0):
        self.L.orc3_step_order(self.h, order, u0)

    def error(self):
        return self.L.orc3_error(self.h)

    def cg_iterations(self):
        it = (C.c_int * 3)()
        self.L.orc3_cg_iterations(self.h, it)
        return list(it)

    def energy(self):
        e = (C.c_double * 4)()
        self.L.orc3_energy(self.h, e)
        return np.array(list(e))

    def gauss(self, which=0):
        e = (C.c_double * 2)()
        self.L.orc3_gauss(self.h, which, e)
        return e[0], e[1]


class World2:
    """In-process emulation of an nproc-rank (1-D y slabs) MPI run of the 2-D reference loop.
    bc: 0 periodic (Weibel), 1 reconnection walls, 2 shock walls; order (of step()): 0 Weibel, 1 reconnection, 2 shock."""

    def __init__(self, nx, ny, np_cap, nproc=1, delx=1.0, delt=1.0, c=1.0, gfac=0.501, q=(1.0, -1.0), r=(1.0, 1.0),
                 bc=0, fast=False):
        self.L = lib(fast)
        self.nx, self.ny, self.np, self.ndim, self.nsp, self.bc = nx, ny, np_cap, 6, 2, bc
        self.q = np.ascontiguousarray(q, dtype=np.float64)
        self.r = np.ascontiguousarray(r, dtype=np.float64)
        self.delx, self.delt, self.c, self.gfac = delx, delt, c, gfac
        dp = C.POINTER(C.c_double)
        self.h = C.c_void_p(self.L.orc2_create(nx, ny, np_cap, nproc, delx, delt, c, gfac, self.q.ctypes.data_as(dp),
                                               self.r.ctypes.data_as(dp), bc))
        self.nranks = self.L.orc2_nranks(self.h)

    def close(self):
        if self.h:
            self.L.orc2_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def geom(self, rank=0):
        g = (C.c_int * 4)()
        self.L.orc2_rank_geom(self.h, rank, g)
        d = dict(zip(("nys", "nye", "nup", "ndown"), list(g)))
        d["nzs"] = d["nze"] = 0
        return d

    def _shape(self, rank, which):
        g = self.geom(rank)
        nyl = g["nye"] - g["nys"] + 1
        if which in ("up", "gp"):
            return (self.nsp, nyl, self.np, self.ndim)
        if which in ("uf", "df"):
            return (nyl + 4, self.nx + 4, 6)
        if which == "uj":
            return (nyl + 4, self.nx + 4, 3)
        if which == "gkl":
            return (nyl, self.nx, 3)
        if which == "mom":
            return (self.nsp, nyl + 2, self.nx + 2, 7)
        if which == "np2":
            return (self.nsp, nyl)
        if which == "cumcnt":
            return (self.nsp, nyl, self.nx + 1)
        raise KeyError(which)

    def arr(self, which, rank=0):
        shape = self._shape(rank, which)
        if which in ("np2", "cumcnt"):
            p = self.L.orc2_iptr(self.h, rank, 0 if which == "np2" else 1)
        else:
            p = self.L.orc2_dptr(self.h, rank, ("up", "gp", "uf", "df", "uj", "gkl", "mom").index(which))
        return np.ctypeslib.as_array(p, shape=shape)

    def load_weibel(self, n0, v_thi=0.1, v_the=0.1, t_ani=5.0, b0=0.0, seed=20240601):
        self.L.orc2_load_weibel(self.h, n0, v_thi, v_the, t_ani, b0, seed)

    def set_xrange(self, nxs, nxe):
        self.L.orc2_set_xrange(self.h, nxs, nxe)

    def particle_solv(self):
        self.L.orc2_particle_solv(self.h)

    def particle_solv_vay(self):
        self.L.orc2_particle_solv_vay(self.h)

    def set_pusher(self, kind):
        self.L.orc2_set_pusher(self.h, kind)

    def shock_inject(self, prm, nlinj_rows, epoch):
        """inject() of 2d/proj/shock/app.f90:697-852; nlinj_rows per GLOBAL row j - nygs"""
        a = np.ascontiguousarray(nlinj_rows, dtype=np.int32)
        self.L.orc2_shock_inject(self.h, C.byref(prm), a.ctypes.data_as(C.POINTER(C.c_int)), epoch)

    def shock_relocate(self, prm, epoch):
        self.L.orc2_shock_relocate(self.h, C.byref(prm), epoch)

    def pack_particles(self, mode, rank=0):
        """get_particle_count of paraio (3d/common/paraio.f90:1007-1085): (records[n, ndim], lcount[nsp])"""
        lc = (C.c_longlong * 2)()
        n = self.L.orc2_pack_particles(self.h, rank, mode, None, lc)
        buf = np.zeros((max(n, 1), self.ndim))
        self.L.orc2_pack_particles(self.h, rank, mode, buf.ctypes.data_as(C.POINTER(C.c_double)), lc)
        return buf[:n], np.array([lc[0], lc[1]], dtype=np.int64)

    @property
    def nxe_now(self):
        return self.L.orc2_nxe(self.h)

    def field_fdtd_i(self, stage=0):
        self.L.orc2_field_fdtd_i(self.h, stage)

    def bc_particle_x(self, kind=None):
        self.L.orc2_bc_particle_x(self.h, (0 if self.bc == 0 else 1) if kind is None else kind)

    def bc_injection(self, u0):
        self.L.orc2_bc_injection(self.h, u0)

    def bc_particle_y(self):
        self.L.orc2_bc_particle_y(self.h)

    def sort_bucket(self):
        self.L.orc2_sort_bucket(self.h)

    def mom_calc(self):
        self.L.orc2_mom_calc(self.h)

    def step(self, order=0, u0=0.0):
        self.L.orc2_step(self.h, order, u0)

    def error(self):
        return self.L.orc2_error(self.h)

    def cg_iterations(self):
        it = (C.c_int * 3)()
        self.L.orc2_cg_iterations(self.h, it)
        return list(it)

    def energy(self):
        e = (C.c_double * 4)()
        self.L.orc2_energy(self.h, e)
        return np.array(list(e))

    def gauss(self, which=0):
        e = (C.c_double * 2)()
        self.L.orc2_gauss(self.h, which, e)
        return e[0], e[1]
