"""The exchange y <-> z with B -> -B (an axis exchange is a reflection: the axial vector changes sign) is an exact symmetry of the
reference's time loop -- push, Esirkepov deposit, implicit field solve, boundaries, sort.  The device runs 3-D y-slabs through this
relabelling (wm_ctx::swap_yz); here the symmetry itself is checked on the CPU oracle, independent of any GPU code: a world
and its relabelled twin (nx, nz, ny), stepped separately, stay each other's relabelling to round-off (the order of sums differs)."""
import numpy as np
import pytest

from oracle.pyoracle import World3
from tests.util import active_mask, make_world3, squeeze_into_walls

NX, NY, NZ, N0 = 12, 10, 6, 5
COL = [0, 2, 1, 3, 5, 4, 6]                 # x, z, y, ux, uz, uy, id
FCOL = [0, 2, 1, 3, 5, 4]
FSIGN = np.array([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0])


def relabel_particles(up):
    return np.ascontiguousarray(up.transpose(0, 2, 1, 3, 4)[..., COL])


def relabel_field(uf):
    return np.ascontiguousarray(uf.transpose(1, 0, 2, 3)[..., FCOL] * FSIGN)


@pytest.mark.parametrize("bc,order,u0", [(0, 0, 0.0), (1, 1, 0.0), (2, 2, 0.3)], ids=["periodic", "reconnection", "shock"])
def test_relabelled_world_evolves_to_the_relabelled_state(bc, order, u0):
    w = make_world3(NX, NY, NZ, N0, steps=2, b0=0.2, bc=bc, order=order, u0=u0)
    t = World3(NX, NZ, NY, w.np, q=w.q, r=w.r, bc=bc)
    t.arr("up")[...] = relabel_particles(w.arr("up"))
    t.arr("gp")[...] = t.arr("up")
    t.arr("np2")[...] = w.arr("np2").transpose(0, 2, 1)
    t.arr("cumcnt")[...] = w.arr("cumcnt").transpose(0, 2, 1, 3)
    t.arr("uf")[...] = relabel_field(w.arr("uf"))
    t.arr("df")[...] = relabel_field(w.arr("df"))
    for _ in range(5):
        w.step(order, u0)
        t.step(order, u0)
    assert w.error() == 0 and t.error() == 0
    assert w.cg_iterations() == t.cg_iterations()
    ref = relabel_field(w.arr("uf"))
    assert np.abs(t.arr("uf") - ref).max() < 1e-11 * np.abs(ref).max()
    assert np.array_equal(t.arr("np2"), w.arr("np2").transpose(0, 2, 1))
    assert np.array_equal(t.arr("cumcnt"), w.arr("cumcnt").transpose(0, 2, 1, 3))
    # same particles in every cell (the order inside a cell may differ: migration order), coordinates to round-off
    a, b = t.arr("up"), relabel_particles(w.arr("up"))
    m = active_mask(t.arr("np2"), t.np)
    ia = np.where(m, a[..., 6].view(np.int64), np.iinfo(np.int64).max)
    ib = np.where(m, b[..., 6].view(np.int64), np.iinfo(np.int64).max)
    oa, ob = np.argsort(ia, axis=-1), np.argsort(ib, axis=-1)
    assert np.array_equal(np.take_along_axis(ia, oa, -1), np.take_along_axis(ib, ob, -1))
    xa = np.take_along_axis(a[..., :6], oa[..., None], -2)
    xb = np.take_along_axis(b[..., :6], ob[..., None], -2)
    assert np.abs(np.where(m[..., None], xa - xb, 0.0)).max() < 1e-11
    np.testing.assert_allclose(t.energy(), w.energy(), rtol=1e-11)
    w.close(); t.close()
