"""The hand-written oracle against THE REFERENCE'S OWN SOURCE, translated and run here.

oracle/f2cxx/build_ref.py turns the reference's Fortran files (particle, field, sort, boundary_periodic, mom_calc of both trees and
the boundary modules of the reconnection and shock set-ups, read where they lie under /root/reference) into C++ with the mechanical
translator oracle/f2cxx/f2cxx.py (unit-tested on its own: tests/test_f2cxx_translator.py) and compiles them into oracle/_ref/.
oracle/f2cxx/pyref.py drives the result with the drivers' call sequences, one private copy of the library per emulated MPI rank.

Every comparison below is BIT FOR BIT (the oracle runs with one OpenMP thread; its multi-threaded deposit differs from the serial
order by round-off): fields incl. ghosts, `np2`, `cumcnt`, and the particle records of every pencil IN ORDER, after every step --
for the periodic, reconnection and shock boundary modules, both pushers, the moments, 2-D (incl. the `ieee_down` sections) and
3-D, on 1 rank and on y, z and y x z rank grids.  This is what pins the oracle; where the translated library is absent and cannot
be built (no /root/reference and no prebuilt oracle/_ref), the tests skip and the committed fixtures generated from it
(tests/golden/ref_cases.npz, tests/test_ref_golden.py) stand in."""
import numpy as np
import pytest

from oracle import pyoracle
from oracle.f2cxx import pyref
from tests.util import active_mask, make_world2, make_world3

pytestmark = pytest.mark.skipif(not (pyref.available(3) and pyref.available(2)),
                                reason="the translated reference is not built and /root/reference is absent")


@pytest.fixture(autouse=True)
def one_thread():
    """serial oracle = the serial reference; the team size is restored for the other tests"""
    before = pyoracle.num_threads()
    pyoracle.set_num_threads(1)
    yield
    pyoracle.set_num_threads(before)


def seed(R, w):
    for rk in range(w.nranks):
        for k in ("up", "gp", "uf", "np2", "cumcnt"):
            R.arr(k, rk)[...] = w.arr(k, rk)


def assert_identical(R, w, what, extra=()):
    for rk in range(w.nranks):
        for k in ("np2", "cumcnt", "uf") + tuple(extra):
            assert np.array_equal(w.arr(k, rk), R.arr(k, rk)), (what, "rank", rk, k, float(np.abs(w.arr(k, rk) - R.arr(k, rk)).max()))
        m = active_mask(w.arr("np2", rk), w.np)
        # the records bit-cast to integers: positions, momenta and the 64-bit IDs, in pencil order
        assert np.array_equal(w.arr("up", rk)[m].view(np.int64), R.arr("up", rk)[m].view(np.int64)), (what, "rank", rk, "up")


def wrapped_fraction(before, after, np2, np_cap, axis, length):
    """fraction of particles whose coordinate `axis` jumped by about the box length in one step"""
    m = active_mask(np2, np_cap)
    return float((np.abs(after[m][:, axis] - before[m][:, axis]) > 0.5 * length).mean())


@pytest.mark.parametrize("nproc_j,nproc_k", [(1, 1), (1, 2), (2, 1), (2, 2), (3, 2)])
def test_weibel_3d_rank_grids(nproc_j, nproc_k):
    nx, ny, nz, n0 = 10, 6, 7, 5
    w = make_world3(nx, ny, nz, n0, nproc_j=nproc_j, nproc_k=nproc_k)
    R = pyref.RefWorld(3, nx, ny, nz, w.np, nproc_j=nproc_j, nproc_k=nproc_k, q=w.q, r=w.r, bounds=True)
    for rk in range(w.nranks):
        g = R.geom(rk)
        assert w.geom(rk) == {k: g[k] for k in w.geom(rk)}            # para_range + the periodic rank table (mpi_set.f90:45-94)
    seed(R, w)
    for it in range(8):
        w.step()
        R.step()
        assert w.error() == 0
        assert_identical(R, w, f"3-D Weibel {nproc_j}x{nproc_k}, step {it}")


def test_weibel_3d_stage_by_stage():
    """one step taken apart: the push, the three work arrays of field__fdtd_i (SAVEd locals of the reference), every boundary
    procedure and the sort, each compared before the next one runs"""
    nx, ny, nz, n0 = 12, 8, 6, 6
    w = make_world3(nx, ny, nz, n0)
    R = pyref.RefWorld(3, nx, ny, nz, w.np, q=w.q, r=w.r, bounds=True)
    seed(R, w)
    for it in range(3):
        w.particle_solv()
        R.particle_solv()
        m = active_mask(w.arr("np2"), w.np)
        assert np.array_equal(w.arr("gp")[m].view(np.int64), R.arr("gp")[m].view(np.int64)), "particle__solv"
        w.field_fdtd_i()
        R.field_fdtd_i()
        for name in ("uj", "gkl", "df"):
            assert np.array_equal(w.arr(name), R.saved("field__fdtd_i", name)), ("field__fdtd_i", name)
        assert np.array_equal(w.arr("uf"), R.arr("uf"))
        w.bc_particle_x()
        R.bc_particle_x()
        assert np.array_equal(w.arr("gp")[m].view(np.int64), R.arr("gp")[m].view(np.int64)), "bc__particle_x"
        w.bc_particle_yz()
        R.bc_particle_yz()
        assert np.array_equal(w.arr("np2"), R.arr("np2")), "bc__particle_yz"
        m = active_mask(w.arr("np2"), w.np)
        assert np.array_equal(w.arr("gp")[m].view(np.int64), R.arr("gp")[m].view(np.int64)), "bc__particle_yz"
        w.sort_bucket()
        R.sort_bucket()
        assert_identical(R, w, f"sort__bucket, step {it}")


@pytest.mark.parametrize("nproc", [1, 2, 3])
def test_weibel_2d_with_round_down_wraps(nproc):
    """2-D: `bc__particle_x` / `bc__particle_y` run under ieee_set_rounding_mode(ieee_down) (2d/common/boundary_periodic.f90:74,124);
    the oracle reproduces that with exact error-term arithmetic -- they must agree on every wrapped particle"""
    nx, ny, n0 = 9, 7, 8
    w = make_world2(nx, ny, n0, nproc=nproc)
    R = pyref.RefWorld(2, nx, ny, 0, w.np, nproc_j=nproc, q=w.q, r=w.r, bounds=True)
    seed(R, w)
    wraps = 0
    for it in range(12):
        w.particle_solv()
        R.particle_solv()
        w.field_fdtd_i()
        R.field_fdtd_i()
        before = [w.arr("gp", rk).copy() for rk in range(nproc)]
        w.bc_particle_x()
        R.bc_particle_x()
        wraps += sum(int((np.abs(w.arr("gp", rk)[..., 0] - before[rk][..., 0]) > 1.0).sum()) for rk in range(nproc))
        w.bc_particle_y()
        R.bc_particle_yz()
        w.sort_bucket()
        R.sort_bucket()
        assert w.error() == 0
        assert_identical(R, w, f"2-D Weibel {nproc} ranks, step {it}")
    assert wraps > 0, "no particle crossed the periodic x boundary: the round-down section was not exercised"


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("bc,order,u0", [(1, pyref.ORDER_RECONNECTION, 0.0), (2, pyref.ORDER_SHOCK, -0.3)])
@pytest.mark.parametrize("nproc", [1, 2])
def test_wall_set_ups(dim, bc, order, u0, nproc):
    """the reconnection and shock boundary modules ({2d,3d}/proj/{reconnection,shock}/boundary_*.f90: reflecting walls, the
    injection wall, wall rules of curre / dfield / phi) in the drivers' call order"""
    nx, ny, nz, n0 = 14, 8, 6, 5
    if dim == 2:
        w = make_world2(nx, ny, n0, nproc=nproc, bc=bc)
        R = pyref.RefWorld(2, nx, ny, 0, w.np, nproc_j=nproc, q=w.q, r=w.r, bc=bc, bounds=True)
    else:
        w = make_world3(nx, ny, nz, n0, nproc_k=nproc, bc=bc)
        R = pyref.RefWorld(3, nx, ny, nz, w.np, nproc_k=nproc, q=w.q, r=w.r, bc=bc, bounds=True)
    seed(R, w)
    for it in range(8):
        w.step(order, u0)
        R.step(order, u0)
        assert w.error() == 0
        assert_identical(R, w, f"{dim}-D bc {bc} on {nproc} ranks, step {it}")


@pytest.mark.parametrize("dim", [2, 3])
def test_vay_pusher_and_moments(dim):
    nx, ny, nz, n0 = 12, 8, 6, 5
    w = make_world2(nx, ny, n0, b0=0.3) if dim == 2 else make_world3(nx, ny, nz, n0, b0=0.3)
    R = pyref.RefWorld(dim, nx, ny, nz, w.np, q=w.q, r=w.r, bounds=True)
    seed(R, w)
    w.set_pusher(1)
    for it in range(4):
        w.step()
        R.step(vay=True)
        assert_identical(R, w, f"{dim}-D particle__solv_vay, step {it}")
    w.mom_calc()
    R.mom_calc()
    assert np.abs(w.arr("mom")).max() > 0
    assert np.array_equal(w.arr("mom"), R.arr("mom")), float(np.abs(w.arr("mom") - R.arr("mom")).max())


def test_cfl_half_and_unequal_masses_3d():
    """the reconnection drivers' parameters: cfl = 0.5 (other CG coefficients and iteration counts) and mass ratio 16"""
    nx, ny, nz, n0 = 10, 6, 6, 4
    q, r, _ = pyoracle.weibel_constants(n0, mass_ratio=16.0)
    w = pyoracle.World3(nx, ny, nz, n0 * nx * 3, q=q, r=r, delt=0.5)
    w.load_weibel(n0)
    R = pyref.RefWorld(3, nx, ny, nz, w.np, q=q, r=r, delt=0.5, bounds=True)
    seed(R, w)
    for it in range(6):
        w.step()
        R.step()
        assert_identical(R, w, f"cfl 0.5, step {it}")
    assert w.error() == 0


def test_the_reference_aborts_of_the_path_are_reproduced():
    """pencil overflow `np2 > np` stops the reference (boundary_periodic.f90:435-438); the translated STOP is reported to the
    driver instead of killing the test process, and the oracle flags the same condition"""
    nx, ny, nz, n0 = 6, 4, 4, 4
    q, r, _ = pyoracle.weibel_constants(n0)
    npc = n0 * nx + 1                                              # hardly any head-room: the first migration overflows a pencil
    w = pyoracle.World3(nx, ny, nz, npc, q=q, r=r)
    w.load_weibel(n0, v_thi=0.4, v_the=0.4)
    R = pyref.RefWorld(3, nx, ny, nz, npc, q=q, r=r)
    seed(R, w)
    stopped = False
    for _ in range(20):
        w.step()
        try:
            R.step()
        except RuntimeError as e:
            assert "STOP at 3d/common/boundary_periodic.f90" in str(e)
            stopped = True
            break
        if w.error() != 0:
            break
    assert stopped and w.error() != 0
