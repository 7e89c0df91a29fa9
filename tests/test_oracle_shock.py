"""CPU checks of the oracle's restatement of the shock driver's inject() / relocate() (2d/proj/shock/app.f90:615-852,
3d/proj/shock/app.f90:644-906): populations, IDs, placement, the upstream field columns, and slab-count independence."""
import numpy as np
import pytest

from tests.shock_util import V0, local_rows, make_shock_world, monotone, row_counts, shock_prm
from tests.util import active_mask


@pytest.mark.parametrize("dim", [2, 3])
def test_inject_and_relocate_properties(dim):
    n0, nx, ny, nz, nxe0 = 6, 20, 8, 4, 14
    w = make_shock_world(dim, nx, ny, nz, n0, nxe0)
    prm = shock_prm(n0)
    nrows = ny * (nz if dim == 3 else 1)
    before = w.arr("np2").copy()
    counts = row_counts(nrows, 1, n0)
    w.shock_inject(prm, counts, 1)
    assert w.error() == 0
    np2 = w.arr("np2")
    per_row = counts.reshape(before.shape[1:])
    assert np.array_equal(np2, before + per_row[None])
    m_new = active_mask(np2, w.np) & ~active_mask(before, w.np)
    new = w.arr("up")[m_new]
    # IDs continue the global numbering per species: -(nptotal + 1) ... -(nptotal + sum(counts))
    for isp in range(2):
        nptot = int(before[isp].sum())
        mine = np.sort(-w.arr("up")[isp][m_new[isp]][:, -1].view(np.int64))
        assert np.array_equal(mine, np.arange(nptot + 1, nptot + counts.sum() + 1))
    # placement: x in [nxe - x0 + ux*dt, nxe + ux*dt], both species of a pair at the same y (z) and x before the drift
    x0 = abs(V0) * w.delt
    assert new[:, 0].min() > nxe0 - x0 - 0.5 and new[:, 0].max() < nxe0 + 0.5
    # the injected particles belong to cell nxe-1: only cumcnt(nxe) moved
    cc = w.arr("cumcnt")
    assert np.array_equal(monotone(cc)[..., -1], np2)
    # upstream field columns
    uf = w.arr("uf")
    by, bz = prm.b0 * np.sin(prm.theta_bn) * np.cos(prm.phi_bn), prm.b0 * np.sin(prm.theta_bn) * np.sin(prm.phi_bn)
    col = uf[..., nxe0 - 1 - 2 + 2, :]            # x index of cell nxe-1 in the box (two ghosts, nxgs = 2)
    assert np.allclose(col[..., 1], by) and np.allclose(col[..., 2], bz)
    assert np.allclose(col[..., 4], V0 * bz / w.c) and np.allclose(col[..., 5], -V0 * by / w.c)
    # relocate: box grows by one cell holding exactly n0 particles per row and species
    before = np2.copy()
    w.shock_relocate(prm, 1)
    assert w.nxe_now == nxe0 + 1
    assert np.array_equal(w.arr("np2"), before + n0)
    cc = w.arr("cumcnt")
    assert np.all(cc[..., nxe0 + 1 - 2] - cc[..., nxe0 - 2] == n0)
    m_new = active_mask(w.arr("np2"), w.np) & ~active_mask(before, w.np)
    xr = w.arr("up")[m_new][:, 0]
    assert xr.min() >= nxe0 and xr.max() < nxe0 + 1
    w.close()


@pytest.mark.parametrize("dim", [2, 3])
def test_slab_count_independence(dim):
    """the same rows receive the same particles whether the world has one slab or two"""
    n0, nx, ny, nz, nxe0 = 4, 16, 8, 4, 12
    prm = shock_prm(n0)
    nrows = ny * (nz if dim == 3 else 1)
    counts = row_counts(nrows, 3, n0)
    recs = []
    for nproc in (1, 2):
        w = make_shock_world(dim, nx, ny, nz, n0, nxe0, nproc=nproc)
        w.shock_inject(prm, counts, 3)
        w.shock_relocate(prm, 3)
        rows = {}
        for rk in range(w.nranks):
            up, np2 = w.arr("up", rk), w.arr("np2", rk)
            gr = local_rows(w, rk, dim)
            flat_np2 = np2.reshape(2, -1)
            flat_up = up.reshape(2, len(gr), w.np, -1)
            for isp in range(2):
                for lr, g in enumerate(gr):
                    rows[(isp, int(g))] = flat_up[isp, lr, :flat_np2[isp, lr]].copy()
        recs.append(rows)
        w.close()
    assert recs[0].keys() == recs[1].keys()
    for key in recs[0]:
        a, b = recs[0][key], recs[1][key]
        assert a.shape == b.shape
        assert np.array_equal(a.view(np.int64), b.view(np.int64)), key
