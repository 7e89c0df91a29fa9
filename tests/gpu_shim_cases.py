"""The Fortran ISO_C_BINDING shim (fortran/*.f90, translated by oracle/f2cxx like the reference itself: oracle/f2cxx/shim_harness.py)
on top of the REAL libwuming_b200.so: the reference's driver sequences (pyref.RefWorld) call the shim's `particle__solv`,
`field__fdtd_i`, `bc__*`, `sort__bucket` with the reference's host arrays, the shim calls the C ABI, the CUDA kernels do the work --
the whole drop-in path a maintainer gets by linking the shim, against the oracle.  The CPU half of this (same driver, same shim,
a recording stub under the C ABI) is tests/test_shim_executed.py.  NOT collected by name: tests/gpu_shim_cases.py runs every case
of this file in a process of its own (the shim binds to libwuming_b200.so through the process-global symbol scope, and a fault in
this newest, least-run path must not take the rest of the GPU suite with it)."""
import ctypes as C

import numpy as np
import pytest

from tests.util import active_mask, canonical_cells, make_world2, make_world3, rel_err

pytestmark = pytest.mark.gpu
NX = 18


@pytest.fixture(scope="module")
def real_backend():
    import wumingpic_b200
    from oracle.f2cxx import shim_harness
    from wumingpic_b200.backend import library_path
    wumingpic_b200.load_library()
    L = C.CDLL(library_path(), mode=C.RTLD_GLOBAL)         # the shim's undefined wm_* symbols resolve against the global scope
    glob = C.CDLL(None)
    try:
        same = C.cast(glob.wm_create, C.c_void_p).value == C.cast(L.wm_create, C.c_void_p).value
    except AttributeError:
        same = False
    if not same:
        pytest.skip("wm_* in this process do not resolve to libwuming_b200.so (the CPU stub was loaded first: run with -m gpu)")
    L.wm_destroy.argtypes = [C.c_void_p]
    return L, shim_harness


def shim_world(real_backend, dim, w, bc=0):
    from oracle.f2cxx import pyref
    L, shim_harness = real_backend
    R = pyref.RefWorld(dim, w.nx, w.ny, w.nz if dim == 3 else 0, w.np, q=w.q, r=w.r, bc=bc, lib=shim_harness.build(dim))
    for k in ("up", "gp", "uf", "np2", "cumcnt"):
        R.arr(k)[...] = w.arr(k)
    return R


def destroy(real_backend, R):
    """the shim keeps its context for the life of the program (like the reference's module state); the test releases it"""
    f = R.ranks[0].L.f2cxx_modvar__wuming_b200_c__ctx
    f.restype = C.c_void_p
    ctx = C.c_void_p.from_address(f())
    if ctx.value:
        real_backend[0].wm_destroy(ctx)
        ctx.value = None


def compare(R, w, tol_f=1e-8, tol_p=1e-9):      # the tolerances of tests/test_gpu_five_calls.py
    assert np.array_equal(R.arr("np2"), w.arr("np2")) and np.array_equal(R.arr("cumcnt"), w.arr("cumcnt"))
    assert rel_err(R.arr("uf"), w.arr("uf")) < tol_f
    for (cg, rg), (cr, rr) in zip(canonical_cells(R.arr("up"), R.arr("np2"), R.arr("cumcnt")),
                                  canonical_cells(w.arr("up"), w.arr("np2"), w.arr("cumcnt"))):
        assert np.array_equal(cg, cr)
        assert np.array_equal(rg[:, -1].view(np.int64), rr[:, -1].view(np.int64))
        if len(rg):
            assert np.abs(rg[:, :-1] - rr[:, :-1]).max() < tol_p


@pytest.mark.parametrize("dim,bc,order,u0", [(3, 0, 0, 0.0), (2, 0, 0, 0.0), (3, 1, 1, 0.0), (2, 2, 2, 0.3)],
                         ids=["3d-weibel", "2d-weibel", "3d-reconnection", "2d-shock"])
def test_driver_through_the_shim_sync_every_call(real_backend, dim, bc, order, u0):
    """the default mode of the shim: every procedure hands the reference's host-visible result back (an unmodified driver)"""
    w = make_world3(NX, 8, 6, 6, bc=bc) if dim == 3 else make_world2(NX, 14, 8, bc=bc)
    R = shim_world(real_backend, dim, w, bc=bc)
    try:
        # one step taken apart: what particle__solv hands back is the pushed set
        w.particle_solv()
        R.particle_solv()
        m = active_mask(w.arr("np2"), w.np)
        for c in range(w.ndim - 1):
            assert rel_err(R.arr("gp")[m][:, c], w.arr("gp")[m][:, c]) < 1e-12
        # ... then whole steps of the set-up's own call order, from the start state again
        w2 = make_world3(NX, 8, 6, 6, bc=bc) if dim == 3 else make_world2(NX, 14, 8, bc=bc)
        for k in ("up", "gp", "uf", "np2", "cumcnt"):
            R.arr(k)[...] = w2.arr(k)
        for it in range(4):
            w2.step(order, u0)
            R.step(order=order, u0=u0)
            assert w2.error() == 0
            compare(R, w2)
    finally:
        destroy(real_backend, R)


def test_driver_through_the_shim_resident(real_backend):
    """WM_SHIM_RESIDENT: the five calls run the fused kernel + lazy sort on device-resident state; the host arrays are refreshed
    by wm_shim_sync_to_host only"""
    w = make_world3(NX, 8, 6, 6)
    R = shim_world(real_backend, 3, w)
    try:
        R.ranks[0].call("wm_shim_set_mode", 1)
        for _ in range(5):
            w.step()
            R.step()
        assert not np.array_equal(R.arr("uf"), w.arr("uf"))          # nothing came back yet
        R.ranks[0].call("wm_shim_sync_to_host", R.arr("up"), R.arr("uf"), R.arr("np2"), R.arr("cumcnt"))
        compare(R, w)
    finally:
        destroy(real_backend, R)


def test_moments_through_the_shim(real_backend):
    """the drivers' output block -- mom_calc__accl, mom_calc__nvt, bc__mom -- through the shim: one wm_mom_calc on the device-resident
    particles; interior nodes against the oracle (tests/test_gpu_parity_variants.py::test_moments for the tolerance)"""
    w = make_world3(NX, 8, 6, 6)
    R = shim_world(real_backend, 3, w)
    try:
        for _ in range(2):
            w.step()
            R.step()
        w.mom_calc()
        R.mom_calc()
        got, ref = R.arr("mom"), w.arr("mom")
        inner = (slice(None),) + (slice(1, -1),) * 3
        for l in range(7):
            assert rel_err(got[inner][..., l], ref[inner][..., l]) < 1e-9, l
        assert abs(got[inner][..., 0].sum() - w.arr("np2").sum()) < 1e-9 * w.arr("np2").sum()
    finally:
        destroy(real_backend, R)


@pytest.mark.parametrize("setup,dim", [("weibel", 3), ("reconnection", 2), ("shock", 3)])
def test_patched_main_loop_on_the_gpu(real_backend, setup, dim):
    """the reference's OWN main loop (app__main after `call init()`, with the edits of tools/make_reference_patch.py --resident),
    translated, on top of the shim and the real library: 12 steps device-resident, outputs every 3 / 4 / 5 steps -- every record an
    output procedure is handed is the oracle's state of that step (checksums; the CPU twin with exact transfer counts is
    tests/test_shim_executed.py::test_patched_main_loop_of_every_driver)"""
    from oracle.f2cxx import mainloop_harness
    from tests.mainloop_util import MainLoop, assert_records_match
    if mainloop_harness.build(setup, dim) is None:
        pytest.skip("the main-loop library is not built and /root/reference is absent")
    bc = {"weibel": 0, "reconnection": 1, "shock": 2}[setup]
    w = make_world3(NX, 8, 6, 6, bc=bc) if dim == 3 else make_world2(NX, 14, 8, bc=bc)
    ml = MainLoop(setup, dim, w, max_it=12, intvl_ptcl=5, intvl_orb=4, intvl_mom=3, intvl_expand=2, u0=0.3 if setup == "shock" else 0.0)
    try:
        got = ml.run()
        want = ml.expected()
        assert_records_match(got, want, rtol=1e-7)
    finally:
        f = ml.R.L.f2cxx_modvar__wuming_b200_c__ctx
        f.restype = C.c_void_p
        ctx = C.c_void_p.from_address(f())
        if ctx.value:
            real_backend[0].wm_destroy(ctx)
            ctx.value = None


def test_the_whole_weibel_driver_on_the_gpu(real_backend):
    """init() + app__main of 3d/proj/weibel/app.f90 -- the reference's text, patched --resident, translated -- on top of the shim and
    the real library: the driver's own loader fills the host arrays (random inputs = the oracle's keyed draws), the context is created
    inside its last __init, 10 steps run device-resident, and the outputs at the driver's cadences are the oracle's states of those steps
    (CPU twin: tests/test_shim_executed.py::test_the_whole_weibel_driver_behind_load_config)"""
    from oracle.f2cxx import mainloop_harness as mh
    from tests.mainloop_util import assert_records_match
    from tests.test_shim_executed import full_driver
    A, w, cfg = full_driver(3)
    try:
        A.call("harness__main")
        assert A.leftover() == (0, 0, 0)
        buf = (C.c_double * 100000)()
        A.L.f90rt_captured.argtypes = [C.POINTER(C.c_double), C.c_int]
        n = A.L.f90rt_captured(buf, len(buf))
        v, got, i = list(buf[:n]), [], 0
        while i < n:
            if v[i] != mh.MAGIC:
                i += 1
                continue
            got.append((mh.KINDS[int(v[i + 1])], int(v[i + 2]), v[i + 4:i + 4 + int(v[i + 3])]))
            i += 4 + int(v[i + 3])

        def sums():
            up, m = w.arr("up"), active_mask(w.arr("np2"), w.np)
            return [float(w.arr("np2").sum()), float(w.arr("uf").sum()), float((up[m][:, 0] + 3.0 * up[m][:, w.ndim - 2]).sum())]
        want = []
        for it in range(1, 11):
            w.step()
            if it % 4 == 0:
                want.append(("io__ptcl", it, sums()))
            if it % 3 == 0:
                w.mom_calc()
                want.append(("io__mom", it, [float(w.arr("mom")[:, 1:-1, 1:-1, 1:-1].sum()), float(w.arr("uf").sum())]))
        want.append(("save_restart", 11, sums()))
        assert_records_match(got, want, rtol=1e-7)
    finally:
        f = A.L.f2cxx_modvar__wuming_b200_c__ctx
        f.restype = C.c_void_p
        ctx = C.c_void_p.from_address(f())
        if ctx.value:
            real_backend[0].wm_destroy(ctx)
            ctx.value = None


def test_the_whole_reconnection_driver_on_the_gpu(real_backend):
    """3d/proj/reconnection/app.f90 -- init() with the Harris-sheet loader and the one-off host-side sort, then 10 steps of its loop
    (reflecting walls, cfl 0.5), the reference's text patched --resident -- on top of the shim and the real library (CPU twin:
    tests/test_shim_executed.py::test_the_whole_reconnection_driver)"""
    from tests.mainloop_util import whole_reconnection
    made = []
    try:
        whole_reconnection(3, rtol=1e-7, after_init=made.append)
    finally:
        for A in made:
            f = A.L.f2cxx_modvar__wuming_b200_c__ctx
            f.restype = C.c_void_p
            ctx = C.c_void_p.from_address(f())
            if ctx.value:
                real_backend[0].wm_destroy(ctx)
                ctx.value = None
