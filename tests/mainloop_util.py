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
        meta = os.path.join(os.path.dirname(lib), f"main_{setup}{dim}d.meta")
        self.final_save = open(meta).read().split()[1] == "1" if os.path.exists(meta) else not (setup == "reconnection" and dim == 2)
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


# ---- the whole drivers (init() + loop), shared by the CPU tests (stub device) and the GPU tests (real library) -------------------------
import pytest  # noqa: E402
from oracle import pyoracle  # noqa: E402


def records(A):
    from oracle.f2cxx import mainloop_harness as mh
    buf = (C.c_double * 200000)()
    A.L.f90rt_captured.argtypes = [C.POINTER(C.c_double), C.c_int]
    n = A.L.f90rt_captured(buf, len(buf))
    v, got, i = list(buf[:n]), [], 0
    while i < n:
        if v[i] != mh.MAGIC:
            i += 1
            continue
        got.append((mh.KINDS[int(v[i + 1])], int(v[i + 2]), v[i + 4:i + 4 + int(v[i + 3])]))
        i += 4 + int(v[i + 3])
    return got


def sums(w):
    up, m = w.arr("up"), active_mask(w.arr("np2"), w.np)
    return [float(w.arr("np2").sum()), float(w.arr("uf").sum()), float((up[m][:, 0] + 3.0 * up[m][:, w.ndim - 2]).sum())]


def whole_reconnection(dim, rtol=1e-11, after_init=None):
    """-> the RefApp after init() + 10 steps of the patched reconnection driver, its outputs checked against the oracle"""
    from oracle.f2cxx import mainloop_harness
    from oracle.pyoracle import World2, World3
    from tests.test_ref_driver_procs import REC_CFG
    lib = mainloop_harness.build_full(dim, "reconnection")
    if lib is None:
        pytest.skip("the full-driver library is not built and /root/reference is absent")
    nz = 3
    cfg = dict(REC_CFG)
    if dim == 3:
        cfg.update(num_process_j=1, n_z=nz)
    A = pyref.RefApp(f"reconnection{dim}d", path=lib)
    nx, ny = cfg["n_x"], cfg["n_y"]
    A.configure([0, 2, ny + 1] if dim == 2 else [0, 2, ny + 1, 2, nz + 1, 0, 0], **cfg)
    for k, v in dict(max_it=10, intvl_ptcl=4, intvl_orb=1000, intvl_mom=3, verbose=0).items():
        A.scalar(k, C.c_int).value = v
    A.scalar("max_elapsed", C.c_double).value = 1e30
    rng = np.random.default_rng(11)
    A.feed(uniform=0.02 + 0.96 * rng.random(200000), normal=rng.standard_normal(200000))       # more than the loader draws
    A.call("init")
    if after_init:
        after_init(A)
    npcap = A.scalar("np").value
    w = (World3(nx, ny, nz, npcap, q=A.array("q").copy(), r=A.array("r").copy(), bc=1, delt=0.5) if dim == 3 else
         World2(nx, ny, npcap, q=A.array("q").copy(), r=A.array("r").copy(), bc=1, delt=0.5))
    for name in ("up", "gp", "uf"):
        w.arr(name)[...] = A.array(name)
    w.arr("np2")[...] = A.array("np2", np.int32)
    w.arr("cumcnt")[...] = A.array("cumcnt", np.int32)
    # the load is sorted: every particle lies in the cell its index puts it in
    up, cc = w.arr("up"), w.arr("cumcnt")
    for idx in np.ndindex(w.arr("np2").shape):
        cell = np.searchsorted(cc[idx], np.arange(w.arr("np2")[idx]), side="right") - 1 + 2
        assert np.array_equal(cell, up[idx][:len(cell), 0].astype(int))
    records(A)                                            # drop the energy.dat record of it0
    A.call("harness__loop")
    got = records(A)
    want = []
    interior = (slice(None),) + (slice(1, -1),) * dim
    for it in range(1, 11):
        w.step(1, 0.0)
        assert w.error() == 0
        if it % 4 == 0:
            want.append(("io__ptcl", it, sums(w)))
        if it % 3 == 0:
            w.mom_calc()
            want.append(("io__mom", it, [float(w.arr("mom")[interior].sum()), float(w.arr("uf").sum())]))
    if dim == 3:                                           # the 2-D driver ends without a final save_restart
        want.append(("save_restart", 11, sums(w)))
    assert_records_match(got, want, rtol)
    return A


def whole_shock(dim, rtol=1e-11, after_init=None):
    """-> (RefApp, steps) after init() + 8 steps of the patched shock driver with its own inject() / relocate(), outputs checked"""
    from oracle.f2cxx import mainloop_harness
    from oracle.pyoracle import ShockPrm
    from tests.test_ref_driver_procs import SEED, SHOCK_CFG, oracle_world_from_app, rows_of
    lib = mainloop_harness.build_full(dim, "shock")
    if lib is None:
        pytest.skip("the full-driver library is not built and /root/reference is absent")
    nz, steps = 3, 8
    cfg = dict(SHOCK_CFG)
    if dim == 3:
        cfg.update(num_process_j=1, n_z=nz)
    A = pyref.RefApp(f"shock{dim}d", path=lib)
    ny, n0, nx = cfg["n_y"], cfg["n_ppc"], cfg["n_x"]
    A.configure([0, 2, ny + 1] if dim == 2 else [0, 2, ny + 1, 2, nz + 1, 0, 0], **cfg)
    for k, v in dict(max_it=steps, intvl_ptcl=3, intvl_orb=1000, intvl_mom=4, intvl_expand=1, verbose=0).items():
        A.scalar(k, C.c_int).value = v
    A.scalar("max_elapsed", C.c_double).value = 1e30
    rows = rows_of(dim, ny, nz)
    npr = n0 * (cfg["n_x_ini"] - 1)
    rng = np.random.default_rng(5)
    A.feed(uniform=rng.random(len(rows) * npr * (dim - 1)), normal=rng.standard_normal(2 * len(rows) * npr * 3))
    A.call("init")
    assert A.leftover() == (0, 0, 0)
    if after_init:
        after_init(A)
    w = oracle_world_from_app(A, dim, cfg, nz)
    prm = ShockPrm(n0=n0, v0=A.scalar("v0").value, v_thi=cfg["v_thi"], v_the=cfg["v_the"], b0=A.scalar("b0").value,
                   theta_bn=A.scalar("theta_bn").value, phi_bn=A.scalar("phi_bn").value, l_damp_ini=cfg["l_damp_ini"], seed=SEED)
    u0, v0, delt = A.scalar("u0").value, A.scalar("v0").value, A.scalar("delt").value
    pflux = n0 * abs(v0) * delt * 1.0 * (ny if dim == 2 else ny * nz)
    frac = 0.37
    nginj = int(pflux) + (1 if frac < pflux - int(pflux) else 0)
    counts = np.array([nginj // len(rows) + (1 if i < nginj % len(rows) else 0) for i in range(len(rows))], dtype=np.int32)
    # every random input of the 8 inject() / relocate() calls, in call order (the box grows by one cell per step until it is full)
    nxe0 = A.scalar("nxe").value
    uni, nrm, shuf, nxe = [], [], [], nxe0
    for it in range(1, steps + 1):
        u, n_ = [frac], {1: [], 2: []}
        for (j, k, row), n in zip(rows, counts):
            for ii in range(1, n + 1):
                a, b = pyoracle.philox_uniform2(SEED, row, ii, 0, it)
                u += [a] if dim == 2 else [a, b]
                for isp in (1, 2):
                    n_[isp] += list(pyoracle.keyed_normals(SEED, row, ii, isp, 0, it))
        uni += u
        nrm += n_[1] + n_[2]
        shuf += [[0], list(range(len(rows)))]
        if nxe < nx + 1:
            u, n_ = [], {1: [], 2: []}
            for j, k, row in rows:
                for ii in range(1, n0 + 1):
                    a, b = pyoracle.philox_uniform2(SEED, row, ii, 16, it)
                    u += [a] if dim == 2 else [a, b]
                    for isp in (1, 2):
                        n_[isp] += list(pyoracle.keyed_normals(SEED, row, ii, isp, 16, it))
            uni += u
            nrm += n_[1] + n_[2]
            nxe += 1
    A.feed(uniform=uni, normal=nrm, shuffles=shuf)
    records(A)
    A.call("harness__loop")
    assert A.leftover() == (0, 0, 0)                       # inject / relocate drew exactly what was predicted
    got = records(A)
    want = []
    interior = (slice(None),) + (slice(1, -1),) * dim
    for it in range(1, steps + 1):
        w.step(2, u0)
        w.shock_inject(prm, counts, it)
        w.shock_relocate(prm, it)
        assert w.error() == 0
        if it % 3 == 0:
            want.append(("io__ptcl", it, sums(w)))
        if it % 4 == 0:
            w.mom_calc()
            want.append(("io__mom", it, [float(w.arr("mom")[interior].sum()), float(w.arr("uf").sum())]))
    want.append(("save_restart", steps + 1, sums(w)))
    assert_records_match(got, want, rtol)
    assert w.nxe_now == A.scalar("nxe").value == nxe0 + steps               # the box grew by one cell per step
    return A, steps
