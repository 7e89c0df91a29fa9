"""The Fortran ISO_C_BINDING shim (fortran/wuming_b200_c.f90 + wuming_b200_shim{2,3}d.f90 -- what a maintainer links instead of the
reference's common/*.f90) EXECUTED on CPU.  The image has no Fortran compiler; oracle/f2cxx translates the shim's text like it
translates the reference's (oracle/f2cxx/shim_harness.py), and the C ABI underneath is oracle/_ref/libwm_stub.so: a recorder compiled
against include/wuming_b200.h that forwards every call to tests/shim_stub.py, where the CPU oracle plays the device.  The driver on
top is pyref.RefWorld -- the call sequences of the reference's drivers, unchanged, because the shim's modules have the reference's
module interface.  Checked: the struct the shim fills, which C functions each procedure calls and in which order, which arrays
cross the boundary in both synchronisation modes, that the host-visible results of every call are the reference's (bit for bit:
the oracle does the arithmetic on both sides), the np2 inference from a shock-shaped cumcnt, the error path (a failing C call
becomes the reference's STOP with the library's message), the moments, the shock source, and the rank-grid / communicator set-up."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle
from oracle.f2cxx import pyref, shim_harness
from tests.shim_stub import StubDevice
from tests.util import active_mask, make_world2, make_world3

NX, NY, NZ, N0 = 12, 6, 5, 4


@pytest.fixture(autouse=True)
def one_thread():
    before = pyoracle.num_threads()
    pyoracle.set_num_threads(1)
    yield
    pyoracle.set_num_threads(before)


def shim_world(dim, w, bc=0):
    """the reference's driver (pyref.RefWorld) on top of the translated shim; host arrays seeded with the oracle's state"""
    D = StubDevice()
    R = pyref.RefWorld(dim, w.nx, w.ny, w.nz if dim == 3 else 0, w.np, q=w.q, r=w.r, bc=bc, lib=shim_harness.build(dim))
    for k in ("up", "gp", "uf", "np2", "cumcnt"):
        R.arr(k)[...] = w.arr(k)
    return D, R


def same_state(R, w, what):
    for k in ("np2", "cumcnt", "uf"):
        assert np.array_equal(R.arr(k), w.arr(k)), (what, k)
    m = active_mask(w.arr("np2"), w.np)
    assert np.array_equal(R.arr("up")[m].view(np.int64), w.arr("up")[m].view(np.int64)), (what, "up")


def test_context_is_created_by_the_last_init_with_the_drivers_numbers():
    w = make_world3(NX, NY, NZ, N0)
    D, R = shim_world(3, w)
    assert D.names() == ["wm_create"]                  # bc__init, particle__init, field__init, ... : one context
    f = D.log[0][1]
    assert (f["dim"], f["ndim"], f["np"], f["nsp"]) == (3, 7, w.np, 2)
    assert (f["nxgs"], f["nxge"], f["nygs"], f["nyge"], f["nzgs"], f["nzge"]) == (2, NX + 1, 2, NY + 1, 2, NZ + 1)
    assert (f["nys"], f["nye"], f["nzs"], f["nze"]) == (2, NY + 1, 2, NZ + 1)
    assert (f["nproc_j"], f["nproc_k"], f["rank_j"], f["rank_k"], f["bc_kind"], f["device"]) == (1, 1, 0, 0, 0, -1)
    assert (f["delx"], f["delt"], f["c"], f["gfac"]) == (1.0, 1.0, 1.0, 0.501)
    assert f["q"] == list(w.q) and f["r"] == list(w.r)
    # the C struct the translator derived from `type, bind(c) :: wm_params` has the header's size (the values above its layout)
    from wumingpic_b200.backend import ShockParams, _Params
    assert D.L.stub_sizeof_params() == C.sizeof(_Params) and D.L.stub_sizeof_shock_params() == C.sizeof(ShockParams)


@pytest.mark.parametrize("dim", [3, 2])
def test_weibel_loop_sync_every_call(dim):
    """the default mode: an unmodified driver sees the reference's host-visible results after every call"""
    w = make_world3(NX, NY, NZ, N0) if dim == 3 else make_world2(NX, NY, N0)
    D, R = shim_world(dim, w)
    for it in range(3):
        del D.log[:]
        w.step()
        R.step()
        same_state(R, w, f"step {it}")
        assert D.log == [("wm_upload", ("up", "np2", "cumcnt", "uf")), ("wm_particle_solv", 2, NX + 1), ("wm_download", ("gp",)),
                         ("wm_field_fdtd_i", 2, NX + 1), ("wm_download", ("uf",)),
                         ("wm_bc_particle_x", 2, NX + 1), ("wm_download", ("gp",)),
                         ("wm_bc_particle_yz",), ("wm_sort_bucket", 2, NX + 1), ("wm_download", ("up", "np2", "cumcnt"))], it
    # stage-wise: what particle__solv hands back is the pushed set
    w.particle_solv()
    R.particle_solv()
    m = active_mask(w.arr("np2"), w.np)
    assert np.array_equal(R.arr("gp")[m].view(np.int64), w.arr("gp")[m].view(np.int64))


@pytest.mark.parametrize("bc,order,u0", [(1, pyref.ORDER_RECONNECTION, 0.0), (2, pyref.ORDER_SHOCK, -0.2)])
def test_wall_loops(bc, order, u0):
    w = make_world3(NX, NY, NZ, N0, bc=bc)
    D, R = shim_world(3, w, bc=bc)
    assert D.log[0][1]["bc_kind"] == bc                 # boundary_reconnection__init / boundary_shock__init registered the rules
    for it in range(3):
        del D.log[:]
        w.step(order, u0)
        R.step(order=order, u0=u0)
        same_state(R, w, f"bc {bc} step {it}")
        names = D.names()
        second = "wm_bc_particle_x" if bc == 1 else "wm_bc_injection.u0"
        assert names[:3] == ["wm_upload", "wm_particle_solv", "wm_download"] and names[3] == second, names
        assert "wm_field_fdtd_i" in names and names.index("wm_field_fdtd_i") > names.index(second)
        if bc == 2:
            assert ("wm_bc_injection.u0", u0) in D.log and ("wm_bc_injection", 2, NX + 1) in D.log


def test_resident_mode_moves_nothing_until_asked():
    w = make_world3(NX, NY, NZ, N0)
    D, R = shim_world(3, w)
    L = R.ranks[0]
    L.call("wm_shim_set_mode", 1)                        # WM_SHIM_RESIDENT
    del D.log[:]
    for _ in range(3):
        w.step()
        R.step()
    assert D.names().count("wm_upload") == 1 and "wm_download" not in D.names()       # the first solv uploads, then nothing
    assert not np.array_equal(R.arr("uf"), w.arr("uf"))                                  # the host copy is stale by design
    L.call("wm_shim_sync_to_host", R.arr("up"), R.arr("uf"), R.arr("np2"), R.arr("cumcnt"))
    assert D.log[-1] == ("wm_download", ("up", "np2", "cumcnt", "uf"))
    same_state(R, w, "after wm_shim_sync_to_host")
    # the driver edits the host arrays (shock inject / relocate on the host) and says so: the next solv uploads again
    del D.log[:]
    R.step()
    assert "wm_upload" not in D.names()
    L.call("wm_shim_host_modified")
    R.step()
    assert D.names().count("wm_upload") == 1


def test_np2_from_a_shock_shaped_cumcnt():
    """the shock driver's inject() / relocate() bump np2 and cumcnt(nxe) and leave cumcnt(nxe+1:) stale (3d/proj/shock/app.f90,
    SURVEY.md App. A.8): the population the shim uploads is the largest prefix count of the active range, not cumcnt(nxe+1)"""
    w = make_world3(NX, NY, NZ, N0)
    D, R = shim_world(3, w)
    nxs, nxe = 2, NX - 2                                 # an active range that ends inside the box
    cc = R.arr("cumcnt")                                 # (nsp, nz, ny, nx + 1): index i - nxgs
    cc[...] = np.arange(NX + 1)[None, None, None, :] * 3
    cc[..., nxe - 2] += 5                                # cumcnt(nxe) bumped by the injection ...
    cc[..., nxe - 1:] = 0                                # ... cumcnt(nxe+1) and above stale
    expect = cc[..., :nxe - 1].max(axis=-1)
    R.ranks[0].call("shim_upload", R.arr("up"), R.arr("uf"), cc, nxs, nxe)
    assert D.log[-1] == ("wm_upload", ("up", "np2", "cumcnt", "uf"))
    dev = D.world(next(iter(D.worlds)))
    assert np.array_equal(dev.arr("np2"), expect) and np.array_equal(dev.arr("cumcnt"), cc)
    assert (expect == 3 * (nxe - 2) + 5).all()


def test_a_failing_call_is_the_references_stop():
    w = make_world3(NX, NY, NZ, N0)
    D, R = shim_world(3, w)
    R.particle_solv()
    D.fail_next = ("wm_field_fdtd_i", 3, "stop at cgm after ite_max")          # WM_ERR_CG_ITEMAX <-> field.f90:522-525
    with pytest.raises(RuntimeError, match=r"field__fdtd_i: stop at cgm after ite_max \(code 3\)"):
        R.field_fdtd_i()
    assert D.log[-1][0] == "wm_field_fdtd_i"            # nothing was downloaded after the failure


def test_procedures_refuse_to_run_before_init():
    """the reference's own guard: 'Initialize first by calling particle__init()' + stop (3d/common/particle.f90:69-72)"""
    StubDevice()
    R = pyref._Rank(3, shim_harness.build(3))
    z = np.zeros(8)
    with pytest.raises(RuntimeError, match="STOP"):
        R.call("particle__solv", z, z, z, np.zeros(8, np.int32), 2, 3)


def test_moments_cross_once():
    w = make_world3(NX, NY, NZ, N0)       # from step 0: the CG warm start (a SAVEd local of field__fdtd_i) is zero on both sides
    D, R = shim_world(3, w)
    R.step()
    w.step()
    same_state(R, w, "before the moments")
    del D.log[:]
    w.mom_calc()
    R.mom_calc()                                          # mom_calc__accl + mom_calc__nvt + bc__mom, as the drivers call them
    # default mode: the host arrays may have been edited since the sort, so they travel; then accl only records the range and
    # bc__mom is folded in on the device: ONE compute call, and mom is all that comes back
    assert D.log == [("wm_upload", ("up", "np2", "cumcnt", "uf")), ("wm_mom_calc", 2, NX + 1)]
    assert np.array_equal(R.arr("mom"), w.arr("mom"))


def test_moments_after_the_driver_edited_the_host_arrays():
    """the shock driver's order: sort__bucket, inject() / relocate() on the HOST arrays, then the moment block, then the next
    particle__solv.  In resident mode nothing travels for the moments unless the driver said it edited the arrays
    (wm_shim_host_modified) -- then they do, and the moments are those of the edited state"""
    w = make_world3(NX, NY, NZ, N0)
    D, R = shim_world(3, w)
    L = R.ranks[0]
    L.call("wm_shim_set_mode", 1)
    R.step()
    w.step()
    del D.log[:]
    R.mom_calc()
    w.mom_calc()
    assert D.log == [("wm_mom_calc", 2, NX + 1)] and np.array_equal(R.arr("mom"), w.arr("mom"))      # resident: no transfer but mom
    # the driver's host-side edit: bring the state back, drop the last particle of every pencil, say so
    L.call("wm_shim_sync_to_host", R.arr("up"), R.arr("uf"), R.arr("np2"), R.arr("cumcnt"))
    for a in (R, w):
        cc, n2 = a.arr("cumcnt"), a.arr("np2")
        last = cc[..., -1].copy()
        n2[...] = n2 - 1
        cc[...] = np.minimum(cc, (last - 1)[..., None])
    L.call("wm_shim_host_modified")
    del D.log[:]
    R.mom_calc()
    w.mom_calc()
    assert D.log == [("wm_upload", ("up", "np2", "cumcnt", "uf")), ("wm_mom_calc", 2, NX + 1)]
    assert np.array_equal(R.arr("mom"), w.arr("mom"))
    assert abs(R.arr("mom")[:, 1:-1, 1:-1, 1:-1, 0].sum() - w.arr("np2").sum()) < 1e-9 * w.arr("np2").sum()


def test_sync_to_host_downloads_once_per_step():
    """resident mode with several outputs firing in the same step (io__ptcl, io__orb, the moment block, save_restart all call
    wm_shim_sync_to_host in a patched driver): the first call brings the state back, the others find it current"""
    w = make_world3(NX, NY, NZ, N0)
    D, R = shim_world(3, w)
    L = R.ranks[0]
    L.call("wm_shim_set_mode", 1)
    R.step()
    w.step()
    del D.log[:]
    for _ in range(3):
        L.call("wm_shim_sync_to_host", R.arr("up"), R.arr("uf"), R.arr("np2"), R.arr("cumcnt"))
    assert D.names() == ["wm_download"]
    same_state(R, w, "after the first sync")
    R.step()
    w.step()
    L.call("wm_shim_sync_to_host", R.arr("up"), R.arr("uf"), R.arr("np2"), R.arr("cumcnt"))
    assert D.names().count("wm_download") == 2
    same_state(R, w, "after the next step's sync")


def test_shock_source_marshalling():
    w = make_world3(NX, NY, NZ, N0, bc=2)
    D, R = shim_world(3, w, bc=2)
    L = R.ranks[0]
    L.L.shock_source__init.argtypes = None
    seed = np.array([20240601], dtype=np.int64)
    L.call("shock_source__init", 7, -0.3, 0.01, 0.02, 0.5, 1.1, 0.2, 4.0, seed)
    nl = np.arange(NY * NZ, dtype=np.int32)
    ids = np.arange(2 * NY * NZ, dtype=np.int64) * 1000
    L.call("shock_source__inject", NX, nl, ids, 17)
    name, nxe, epoch, prm, got_nl, got_ids = D.log[-1]
    assert (name, nxe, epoch) == ("wm_shock_inject", NX, 17)
    assert prm == dict(n0=7, v0=-0.3, v_thi=0.01, v_the=0.02, b0=0.5, theta_bn=1.1, phi_bn=0.2, l_damp_ini=4.0, seed=20240601)
    assert np.array_equal(got_nl, nl) and np.array_equal(got_ids, ids)
    L.call("shock_source__relocate", NX + 1, ids, 18)
    name, nxe, epoch, prm, got_ids = D.log[-1]
    assert (name, nxe, epoch) == ("wm_shock_relocate", NX + 1, 18) and np.array_equal(got_ids, ids) and prm["n0"] == 7


def test_rank_grid_and_communicator_are_set_up_when_the_context_is_created():
    """wm_shim_comm_init right after mpi_set__init (before the __init calls): the context is created by the last __init WITH the
    rank grid, and the communicator is connected right behind it -- rank 0 draws the id, MPI_BCAST, wm_comm_init"""
    D = StubDevice()
    R = pyref._Rank(3, shim_harness.build(3))
    R.call("wm_shim_comm_init", 4, 1, 4, 0, 0)           # nproc, nproc_j, nproc_k, nrank, ncomw: rank 0 of a 1 x 4 grid
    assert D.log == []                                    # nothing can be created yet
    nstat = np.zeros(6, np.int32)
    q, r = np.array([1.0, -1.0]), np.array([1.0, 1.0])
    head = [7, 100, 2, 2, NX + 1, 2, NY + 1, 2, 9, 2, NY + 1, 2, 3]      # ... nzgs, nzge = 2, 9; this rank's nzs, nze = 2, 3
    R.call("boundary_periodic__init", *head, 0, 0, 1, 3, 4, 8, 0, 0, nstat, 1.0, 1.0, 1.0, len(nstat))
    R.call("particle__init", *head, 1.0, 1.0, 1.0, q, r)
    assert D.log == []
    R.call("field__init", *head, 8, 0, 1, 0, 1.0, 1.0, 1.0, q, r, 0.501)
    assert D.names() == ["wm_create", "wm_comm_unique_id", "wm_comm_init"]
    f = D.log[0][1]
    assert (f["nproc_j"], f["nproc_k"], f["rank_j"], f["rank_k"], f["nzs"], f["nze"], f["nzge"]) == (1, 4, 0, 0, 2, 3, 9)
    assert D.log[2] == ("wm_comm_init", 4, 0)
    with pytest.raises(RuntimeError, match="STOP"):      # too late to change the grid of an existing context
        R.call("wm_shim_comm_init", 4, 2, 2, 0, 0)


def test_a_non_root_rank_of_a_y_slab_grid():
    """rank 2 of a 4 x 1 grid (y-slabs, what every shipped 3-D sample uses): rank = j nproc_k + k (3d/common/mpi_set.f90:45-60), the
    id arrives by MPI_BCAST (the root alone draws it), the slab bounds the driver computed travel in wm_params"""
    D = StubDevice()
    R = pyref._Rank(3, shim_harness.build(3))
    R.call("wm_shim_comm_init", 4, 4, 1, 2, 0)
    nstat = np.zeros(6, np.int32)
    q, r = np.array([1.0, -1.0]), np.array([1.0, 1.0])
    head = [7, 100, 2, 2, NX + 1, 2, 17, 2, NZ + 1, 10, 13, 2, NZ + 1]       # ny = 16 in four slabs: this rank owns j = 10 .. 13
    R.call("boundary_periodic__init", *head, 3, 1, 2, 2, 4, 8, 0, 0, nstat, 1.0, 1.0, 1.0, len(nstat))
    R.call("field__init", *head, 8, 0, 1, 0, 1.0, 1.0, 1.0, q, r, 0.501)
    R.call("sort__init", *head)
    assert D.log == []                                   # charges and masses are still missing
    R.call("particle__init", *head, 1.0, 1.0, 1.0, q, r)
    assert D.names() == ["wm_create", "wm_comm_init"]    # no wm_comm_unique_id on a non-root rank
    f = D.log[0][1]
    assert (f["nproc_j"], f["nproc_k"], f["rank_j"], f["rank_k"]) == (4, 1, 2, 0)
    assert (f["nygs"], f["nyge"], f["nys"], f["nye"], f["nzs"], f["nze"]) == (2, 17, 10, 13, 2, NZ + 1)
    assert D.log[1] == ("wm_comm_init", 4, 2)


# ---- the reference's own main loop, patched for the resident mode, on top of the shim ---------------------------------------------
@pytest.mark.parametrize("setup,dim", [("weibel", 3), ("weibel", 2), ("reconnection", 3), ("reconnection", 2), ("shock", 3), ("shock", 2)])
def test_patched_main_loop_of_every_driver(setup, dim):
    """app__main of proj/<setup>/app.f90 after `call init()` -- the text of the reference, with the edits of
    tools/make_reference_patch.py --resident -- translated and run for 12 steps with outputs every 3 / 4 / 5 steps: every record an
    output procedure was handed is the state of ITS step (checksums of np2, uf, the particles and the moments against the oracle
    stepped through the same schedule), although the state left the device only at those steps; in the shock loop the host-side
    inject() / relocate() see the current state and their edits reach the device before the next push"""
    from oracle.f2cxx import mainloop_harness
    from tests.mainloop_util import MainLoop, assert_records_match
    if mainloop_harness.build(setup, dim) is None:
        pytest.skip("the main-loop library is not built and /root/reference is absent")
    bc = {"weibel": 0, "reconnection": 1, "shock": 2}[setup]
    w = make_world3(NX, NY, NZ, N0, bc=bc) if dim == 3 else make_world2(NX, NY, N0, bc=bc)
    D = StubDevice()
    ml = MainLoop(setup, dim, w, max_it=12, intvl_ptcl=5, intvl_orb=4, intvl_mom=3, intvl_expand=2, u0=-0.2 if setup == "shock" else 0.0)
    assert D.names() == ["wm_create"]
    got = ml.run()
    want = ml.expected()
    assert_records_match(got, want)
    names = D.names()
    if setup == "shock":
        # the host arrays are edited every step: one download (in front of inject(); the later syncs of the step and the final
        # save_restart find the host current) and one upload (in front of the moment block if there is one, else of the next push)
        # uploads: the initial state, then once behind each of the 12 edits that something on the device still follows (the pushes
        # of steps 2 .. 12 and the moment block of step 12)
        assert names.count("wm_download") == 12 and names.count("wm_upload") == 13
    else:
        # steps with an output: 3 4 5 6 8 9 10 12 -> eight downloads, and the final save_restart finds the host current; ONE upload
        assert names.count("wm_download") == 8 and names.count("wm_upload") == 1
    assert names.count("wm_mom_calc") == 4


# ---- the whole Weibel driver behind load_config: init() + app__main, the reference's text, on top of the shim ------------------------
def full_driver(dim, nz=4, max_it=10):
    """configured like 3d/proj/weibel/config_sample.json (small): harness__configure = the "parameter" section + the size block of
    load_config; the random inputs of the loader are the oracle's keyed draws for the same particles; -> (RefApp, oracle world)"""
    from oracle.f2cxx import mainloop_harness
    from tests.test_ref_driver_procs import SEED, WEIBEL_CFG, rows_of
    lib = mainloop_harness.build_full(dim)
    if lib is None:
        pytest.skip("the full-driver library is not built and /root/reference is absent")
    cfg = dict(WEIBEL_CFG)
    if dim == 3:
        cfg.update(num_process_j=1, n_z=nz)
    A = pyref.RefApp(f"weibel{dim}d", path=lib)
    nx, ny, n0 = cfg["n_x"], cfg["n_y"], cfg["n_ppc"]
    A.configure([0, 2, ny + 1] if dim == 2 else [0, 2, ny + 1, 2, nz + 1, 0, 0], **cfg)
    for k, v in dict(max_it=max_it, intvl_ptcl=4, intvl_orb=1000, intvl_mom=3, verbose=0).items():
        A.scalar(k, C.c_int).value = v
    A.scalar("max_elapsed", C.c_double).value = 1e30
    uni, nrm = [], {1: [], 2: []}
    for j, k, row in rows_of(dim, ny, nz):
        for ii in range(1, n0 * nx + 1):
            a, b = pyoracle.philox_uniform2(SEED, row, ii, 0)
            uni += [a] if dim == 2 else [a, b]
            for isp in (1, 2):
                nrm[isp] += list(pyoracle.keyed_normals(SEED, row, ii, isp, 0))
    A.feed(uniform=uni, normal=nrm[1] + nrm[2])
    q, r, b0 = pyoracle.weibel_constants(n0, mass_ratio=cfg["mass_ratio"], sigma_e=cfg["sigma_e"], omega_pe=cfg["omega_pe"])
    from oracle.pyoracle import World2, World3
    npcap = n0 * nx * (5 if dim == 2 else 3)
    w = World3(nx, ny, nz, npcap, q=q, r=r) if dim == 3 else World2(nx, ny, npcap, q=q, r=r)
    w.load_weibel(n0, v_thi=cfg["v_thi"], v_the=cfg["v_the"], t_ani=cfg["t_ani"], b0=b0, seed=SEED)
    return A, w, cfg


@pytest.mark.parametrize("dim", [3, 2])
def test_the_whole_weibel_driver_behind_load_config(dim):
    """init() -- mpi_set__init, the patch's wm_shim_comm_init, allocation, constants, the module __init calls in the driver's order,
    set_initial_condition + set_particle_ids, energy_history, gp = up -- and app__main's loop, all the reference's text (patched
    --resident), on top of the shim: the device context appears inside the LAST __init with the driver's numbers, the load is the
    oracle's load, and every output of 10 steps is the oracle's state of its step"""
    from tests.mainloop_util import assert_records_match
    from oracle.f2cxx import mainloop_harness as mh
    D = StubDevice()
    A, w, cfg = full_driver(dim)
    A.call("harness__main")
    assert A.leftover() == (0, 0, 0)                     # the loader consumed exactly the random inputs of its particles
    names = D.names()
    assert names[0] == "wm_create" and "wm_comm_init" not in names          # one rank: wm_shim_comm_init only recorded the grid
    f = D.log[0][1]
    nx, ny, n0 = cfg["n_x"], cfg["n_y"], cfg["n_ppc"]
    assert (f["dim"], f["np"], f["nxgs"], f["nxge"], f["nygs"], f["nyge"]) == (dim, n0 * nx * (5 if dim == 2 else 3), 2, nx + 1, 2, ny + 1)
    assert (f["nproc_j"], f["nproc_k"], f["rank_j"], f["rank_k"], f["bc_kind"]) == (1, 1, 0, 0, 0)
    assert f["q"] == pytest.approx(list(w.q), rel=3e-16) and f["r"] == list(w.r) and f["gfac"] == 0.501 and f["delt"] == 1.0
    # the records: MAGIC-framed ones from the output stand-ins; energy_history is the driver's own (its energy.dat records are unframed)
    buf = (C.c_double * 100000)()
    A.L.f90rt_captured.argtypes = [C.POINTER(C.c_double), C.c_int]
    n = A.L.f90rt_captured(buf, len(buf))
    v, got, i, energy = list(buf[:n]), [], 0, []
    while i < n:
        if v[i] != mh.MAGIC:
            energy.append(v[i])
            i += 1
            continue
        got.append((mh.KINDS[int(v[i + 1])], int(v[i + 2]), v[i + 4:i + 4 + int(v[i + 3])]))
        i += 4 + int(v[i + 3])
    # the oracle through the same schedule
    from tests.util import active_mask

    def sums():
        up, m = w.arr("up"), active_mask(w.arr("np2"), w.np)
        return [float(w.arr("np2").sum()), float(w.arr("uf").sum()), float((up[m][:, 0] + 3.0 * up[m][:, w.ndim - 2]).sum())]
    want, e_want = [], [list(w.energy())]
    interior = (slice(None),) + (slice(1, -1),) * dim
    for it in range(1, 11):
        w.step()
        if it % 4 == 0:
            want.append(("io__ptcl", it, sums()))
        if it % 3 == 0:
            w.mom_calc()
            want.append(("io__mom", it, [float(w.arr("mom")[interior].sum()), float(w.arr("uf").sum())]))
            e_want.append(list(w.energy()))
    want.append(("save_restart", 11, sums()))
    assert_records_match(got, want)
    # energy_history ran on the host arrays at it0 and at every moment cadence: energy.dat records = (t, kinetic..., E, B, total)
    rec = 6 if dim == 3 else 5
    # the driver also `write`s the restart file name (one number, the step) before the final save_restart: the trailing value
    assert len(energy) == rec * len(e_want) + 1 and energy[-1] == 11.0
    for k, e in enumerate(e_want):
        r_ = energy[rec * k:rec * (k + 1)]
        tot = ((e[0] + e[1]) + e[2]) + e[3]
        ref = [e[0], e[1], e[2], e[3], tot] if dim == 3 else [e[0] + e[1], e[2], e[3], tot]
        assert r_[0] == 3.0 * k and r_[1:] == pytest.approx(ref, rel=1e-12)
    # ONE upload (the first push); downloads at the output steps 3 4 6 8 9 and for the final save_restart after step 10
    assert names.count("wm_upload") == 1 and names.count("wm_download") == 6


@pytest.mark.parametrize("dim", [3, 2])
def test_sort_outside_the_time_loop_is_done_on_the_host(dim):
    """the drivers call sort__bucket once BEFORE the first push -- on the freshly loaded set (proj/reconnection/app.f90 init:
    `call sort__bucket(gp, up, ...); up = gp`) and on a restarted one (every driver: `io__input(gp, ...); sort__bucket(up, gp, ...)`).
    No pushed set exists on the device then; the shim sorts the host arrays itself (the reference's rule: key int(x), stable,
    cumcnt = exclusive prefix) and the result travels with the first particle__solv"""
    w = make_world3(NX, NY, NZ, N0) if dim == 3 else make_world2(NX, NY, N0)
    D, R = shim_world(dim, w)
    rng = np.random.default_rng(7)
    # scramble every pencil of the loaded set (what a loader that draws x at random leaves behind), same on both sides
    up, np2 = w.arr("up"), w.arr("np2")
    for idx in np.ndindex(np2.shape):
        n = np2[idx]
        up[idx][:n] = up[idx][rng.permutation(n)]
    R.arr("up")[...] = up
    R.arr("cumcnt")[...] = -7                              # intent(out): whatever was there is replaced
    w.arr("gp")[...] = up                                  # the oracle's sort__bucket reads gp and writes up
    w.sort_bucket()
    del D.log[:]
    R.ranks[0].call("sort__bucket", R.arr("gp"), R.arr("up"), R.arr("cumcnt"), R.arr("np2"), R.nxs, R.nxe)
    assert D.log == []                                     # nothing on the device
    m = active_mask(np2, w.np)
    assert np.array_equal(R.arr("cumcnt"), w.arr("cumcnt"))
    assert np.array_equal(R.arr("gp")[m].view(np.int64), w.arr("up")[m].view(np.int64))      # same order too (stable)
    # the driver's `up = gp`, then the time loop: the first push uploads the sorted set
    R.arr("up")[...] = R.arr("gp")
    R.ranks[0].call("wm_shim_set_mode", 1)
    for _ in range(2):
        w.step()
        R.step()
    assert D.names().count("wm_upload") == 1 and D.names().count("wm_sort_bucket") == 2          # in the loop: on the device
    R.ranks[0].call("wm_shim_sync_to_host", R.arr("up"), R.arr("uf"), R.arr("np2"), R.arr("cumcnt"))
    same_state(R, w, "two steps after the host-side sort")


@pytest.mark.parametrize("dim", [2, 3])
def test_the_whole_reconnection_driver(dim):
    """{2d,3d}/proj/reconnection/app.f90, patched --resident: init() -- constants, the Harris-sheet loader with its statement
    functions, the one-off sort__bucket of the load (host side in the shim), `up = gp`, energy_history -- then 10 steps of its loop
    (reflecting walls before the field solve, cfl 0.5) on top of the shim, against an oracle world started from the loaded state"""
    from tests.mainloop_util import whole_reconnection
    D = StubDevice()

    def after_init(A):
        assert D.names() == ["wm_create"] and D.log[0][1]["bc_kind"] == 1 and D.log[0][1]["delt"] == 0.5
    whole_reconnection(dim, after_init=after_init)
    names = D.names()
    assert names.count("wm_upload") == 1 and names.count("wm_bc_particle_x") == 10 and names.count("wm_sort_bucket") == 10


@pytest.mark.parametrize("dim", [2, 3])
def test_the_whole_shock_driver(dim):
    """{2d,3d}/proj/shock/app.f90, patched --resident: init() (cold upstream load, nominal cumcnt), then 8 steps of its loop with the
    driver's OWN inject() and relocate() -- host-side procedures that edit up, np2, cumcnt, uf and nxe every step, bracketed by the
    patch's sync / host_modified calls -- on top of the shim, the box growing; against the oracle's loop with its particle source
    (the random inputs of inject / relocate are the oracle's keyed draws, as in tests/test_ref_driver_procs.py)"""
    from tests.mainloop_util import whole_shock
    D = StubDevice()

    def after_init(A):
        assert D.names() == ["wm_create"] and D.log[0][1]["bc_kind"] == 2
    A, steps = whole_shock(dim, after_init=after_init)
    names = D.names()
    # one download per step (in front of inject()); uploads: the load, then once behind each step's edits when something on the
    # device follows (the pushes of steps 2 .. 8 and the moment block of step 8)
    assert names.count("wm_bc_injection") == steps and names.count("wm_download") == steps and names.count("wm_upload") == steps + 1


@pytest.mark.parametrize("nproc_j,nproc_k", [(2, 1), (1, 2)], ids=["y-slabs", "z-slabs"])
def test_two_ranks_through_the_shim(nproc_j, nproc_k):
    """two ranks of a driver, each with its private copy of the shim (module state = one rank), on two threads: wm_shim_comm_init after
    mpi_set__init, the context of every rank created by its last __init with its OWN slab bounds and rank coordinates, the NCCL id
    drawn on rank 0 and handed round with MPI_BCAST, then the time loop with every rank's host arrays (lower bounds nys / nzs that
    are not the global ones) -- against the oracle's two-rank world, bit for bit"""
    from tests.shim_stub import MultiRankStubDevice
    w = make_world3(NX, 8, 6, N0, nproc_j=nproc_j, nproc_k=nproc_k)
    D = MultiRankStubDevice(2)
    R = pyref.RefWorld(3, NX, 8, 6, w.np, nproc_j=nproc_j, nproc_k=nproc_k, q=w.q, r=w.r, lib=shim_harness.build(3), native_mpi=True,
                       before_init=lambda rk, L: L.call("wm_shim_comm_init", 2, nproc_j, nproc_k, rk, 0))
    for rk in range(2):
        g, f = w.geom(rk), D.prm[rk]
        assert (f["nproc_j"], f["nproc_k"], f["rank_j"], f["rank_k"]) == (nproc_j, nproc_k, rk // nproc_k, rk % nproc_k)
        assert (f["nys"], f["nye"], f["nzs"], f["nze"]) == (g["nys"], g["nye"], g["nzs"], g["nze"])
        assert (f["nygs"], f["nyge"], f["nzgs"], f["nzge"]) == (2, 9, 2, 7)
    assert [e[0] for e in D.logs[0]] == ["wm_create", "wm_comm_unique_id", "wm_comm_init"] and D.logs[0][2] == ("wm_comm_init", 2, 0)
    assert [e[0] for e in D.logs[1]] == ["wm_create", "wm_comm_init"] and D.logs[1][1] == ("wm_comm_init", 2, 1)
    assert D.ids[0] == D.ids[1] == bytes(range(1, 129))                       # the id rank 0 drew reached rank 1
    for rk in range(2):
        for k in ("up", "gp", "uf", "np2", "cumcnt"):
            R.arr(k, rk)[...] = w.arr(k, rk)
    for it in range(3):
        w.step()
        R.step()
        assert w.error() == 0
        for rk in range(2):
            for k in ("np2", "cumcnt", "uf"):
                assert np.array_equal(R.arr(k, rk), w.arr(k, rk)), (it, rk, k)
            m = active_mask(w.arr("np2", rk), w.np)
            assert np.array_equal(R.arr("up", rk)[m].view(np.int64), w.arr("up", rk)[m].view(np.int64)), (it, rk)
    R.close()
