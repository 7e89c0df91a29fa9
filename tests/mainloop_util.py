"""Driver of oracle/_ref/libwuming_main_<setup><dim>d.so (oracle/f2cxx/mainloop_harness.py): the reference's own app__main loop,
patched for the device-resident mode, on top of the translated shim.  TEST INFRASTRUCTURE."""
import ctypes as C

import numpy as np

from oracle.f2cxx import mainloop_harness, pyref
from tests.util import active_mask

ORDER = {"weibel": 0, "reconnection": 1, "shock": 2}
BCNAME = {0: "boundary_periodic", 1: "boundary_reconnection", 2: "boundary_shock"}


class MainLoop:
    def __init__(self, setup, dim, w, max_it, intvl_ptcl, intvl_orb, intvl_mom, intvl_expand=1, u0=0.0):
        lib = mainloop_harness.build(setup, dim)
        if lib is None:
            raise RuntimeError("the main-loop library is not built and /root/reference is absent")
        self.setup, self.dim, self.w, self.u0 = setup, dim, w, u0
        self.cad = dict(max_it=max_it, intvl_ptcl=intvl_ptcl, intvl_orb=intvl_orb, intvl_mom=intvl_mom, intvl_expand=intvl_expand)
        self.R = R = pyref._Rank(dim, lib)
        # one driver (2d/proj/reconnection) ends without a final save_restart after the loop: read it off the assembled text
        import os
        src = os.path.join(os.path.dirname(lib), f"main_{setup}{dim}d.f90")
        self.final_save = open(src).read().count("call save_restart(") >= 2 if os.path.exists(src) else not (setup == "reconnection" and dim == 2)
        nx, ny, nz = w.nx, w.ny, (w.nz if dim == 3 else 1)
        ndim = w.ndim
        geom = [2, nx + 1, 2, ny + 1] + ([2, nz + 1] if dim == 3 else []) + [2, ny + 1] + ([2, nz + 1] if dim == 3 else [])
        head = [ndim, w.np, 2] + geom
        nstat = np.zeros(6, np.int32)
        q, r = np.ascontiguousarray(w.q), np.ascontiguousarray(w.r)
        nb = [0, 0, 0, 0] if dim == 3 else [0, 0]
        # the drivers' init sequence (3d/proj/weibel/app.f90:341-353)
        R.call(BCNAME[w.bc] + "__init", *head, *nb, 4, 8, 0, 0, nstat, w.delx, w.delt, w.c, len(nstat))
        R.call("particle__init", *head, w.delx, w.delt, w.c, q, r)
        R.call("field__init", *head, 8, 0, 1, 0, w.delx, w.delt, w.c, q, r, w.gfac)
        R.call("sort__init", *head)
        R.call("mom_calc__init", *head, w.delx, w.delt, w.c, q, r)
        R.call("harness__alloc", ndim, w.np, 2, 2, nx + 1, 2, ny + 1, 2, nz + 1, 2, ny + 1, 2, nz + 1)
        for k in ("up", "gp", "uf", "np2", "cumcnt"):
            self.array(k)[...] = w.arr(k)
        for k, v in dict(nxs=2, nxe=nx + 1, it0=0, verbose=0, nrank=0, **self.cad).items():
            self.scalar(k, C.c_int).value = v
        self.scalar("max_elapsed", C.c_double).value = 1e30
        self.scalar("u0", C.c_double).value = u0

    def scalar(self, name, ct):
        f = getattr(self.R.L, f"f2cxx_modvar__app__{name}")
        f.restype = C.c_void_p
        return ct.from_address(f())

    def array(self, name):
        f = getattr(self.R.L, f"f2cxx_modarr__app__{name}")
        f.restype = C.c_void_p
        b = (C.c_long * 16)()
        p = f(b)
        like = self.w.arr(name)
        ct = C.c_double if like.dtype == np.float64 else C.c_int
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(like.size,)).reshape(like.shape)

    def run(self):
        """-> [(kind name, it, values)] in the order the driver produced them"""
        self.R.call("harness__main")
        buf = (C.c_double * 100000)()
        self.R.L.f90rt_captured.argtypes = [C.POINTER(C.c_double), C.c_int]
        n = self.R.L.f90rt_captured(buf, len(buf))
        v, out, i = list(buf[:n]), [], 0
        while i < n:
            if v[i] != mainloop_harness.MAGIC:          # e.g. the driver's own write(restart_file, ...) it
                i += 1
                continue
            kind, it, cnt = int(v[i + 1]), int(v[i + 2]), int(v[i + 3])
            out.append((mainloop_harness.KINDS[kind], it, v[i + 4:i + 4 + cnt]))
            i += 4 + cnt
        return out

    # ---- what the oracle expects ---------------------------------------------------------------------------------------------
    def checksums(self):
        w = self.w
        up = w.arr("up")
        m = active_mask(w.arr("np2"), w.np)
        return [float(w.arr("np2").sum()), float(w.arr("uf").sum()), float((up[m][:, 0] + 3.0 * up[m][:, w.ndim - 2]).sum())]

    def host_edit(self):
        """what the harness's inject() / relocate() stand-ins do to the host arrays"""
        w = self.w
        uz = w.arr("up")[..., w.ndim - 2]
        m = active_mask(w.arr("np2"), w.np)
        uz[m] = 0.999 * uz[m]

    def expected(self):
        """step the oracle through the same schedule -> [(kind, it, values)]"""
        w, c, out = self.w, self.cad, []
        interior = (slice(None),) + (slice(1, -1),) * self.dim
        for it in range(1, c["max_it"] + 1):
            w.step(ORDER[self.setup], self.u0)
            assert w.error() == 0
            if self.setup == "shock":
                s = self.checksums()
                out.append(("inject", 0, [s[0], s[2]]))
                self.host_edit()
                if it % c["intvl_expand"] == 0:
                    s = self.checksums()
                    out.append(("relocate", 0, [s[0], s[2]]))
                    self.host_edit()
            if it % c["intvl_ptcl"] == 0:
                out.append(("io__ptcl", it, self.checksums()))
            if it % c["intvl_orb"] == 0:
                out.append(("io__orb", it, self.checksums()))
            if it % c["intvl_mom"] == 0:
                w.mom_calc()
                out.append(("io__mom", it, [float(w.arr("mom")[interior].sum()), float(w.arr("uf").sum())]))
                if self.setup != "shock":
                    out.append(("energy_history", it, self.checksums()))
        if self.final_save:
            out.append(("save_restart", c["max_it"] + 1, self.checksums()))
        return out


def assert_records_match(got, want, rtol=1e-11):
    assert [(k, it) for k, it, _ in got] == [(k, it) for k, it, _ in want], ([(k, it) for k, it, _ in got], [(k, it) for k, it, _ in want])
    for (k, it, a), (_, _, b) in zip(got, want):
        assert len(a) == len(b), (k, it)
        for x, y in zip(a, b):
            assert abs(x - y) <= rtol * max(abs(y), 1.0), (k, it, a, b)
