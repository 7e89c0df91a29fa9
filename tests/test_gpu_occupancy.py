"""GPU-vs-oracle parity at the occupancy the benchmark runs at: 64 particles per cell and species = 128 per
cell = 8 sixteen-slot batches of k_fused3 per cell, so the two-stage prefetch, the running two-ended write cursors and the lazy-sort
indirection are compared with the oracle element-wise, not only through invariants; a ragged load (cells with 0, 1, > 16, > 32 and
> 200 particles); and BASELINE configs[0] at FULL size (2-D Weibel 256 x 256, 20 ppc, 2.62 M particles) GPU-vs-oracle.
Tolerances as in tests/test_gpu_parity3d.py."""
import numpy as np
import pytest

from oracle.pyoracle import World2, weibel_constants
from tests.util import active_mask, backend_for, canonical_cells, make_world3, rel_err, upload_from_world

pytestmark = pytest.mark.gpu


def _compare_particles(b, w, tol=1e-9):
    up, np2, cc = b.empty("up"), b.empty("np2"), b.empty("cumcnt")
    b.download(up, np2, cc)
    assert np.array_equal(np2, w.arr("np2")), "np2 differs"
    assert np.array_equal(cc, w.arr("cumcnt")), "cumcnt differs"
    worst = 0.0
    for (cg, rg), (cr, rr) in zip(canonical_cells(up, np2, cc), canonical_cells(w.arr("up"), w.arr("np2"), w.arr("cumcnt"))):
        assert np.array_equal(cg, cr)
        assert np.array_equal(rg[:, -1].view(np.int64), rr[:, -1].view(np.int64)), "particle ID sets differ"
        if len(rg):
            worst = max(worst, np.abs(rg[:, :-1] - rr[:, :-1]).max())
    assert worst < tol, worst


def _stagewise(w, b, nxe):
    w.particle_solv()
    b.particle__solv(2, nxe)
    gp = b.empty("gp")
    b.download(gp=gp)
    m = active_mask(w.arr("np2"), w.np)
    for c in range(6):
        assert rel_err(gp[m][:, c], w.arr("gp")[m][:, c]) < 1e-13, c
    w.field_fdtd_i(1)
    b.field__fdtd_i(2, nxe, 1)
    assert rel_err(b.download_work("uj"), w.arr("uj")) < 1e-12     # RED deposit of the per-procedure path


@pytest.mark.parametrize("path", ["wm_step", "five-calls"])
def test_64ppc_eight_batches_per_cell(path):
    nx, ny, nz, n0 = 16, 8, 8, 64                       # 131 072 particles, 128 per cell
    w = make_world3(nx, ny, nz, n0, steps=1, np_factor=2)
    b = backend_for(w)
    upload_from_world(b, w)
    if path == "wm_step":
        _stagewise(w, b, nx + 1)
        upload_from_world(b, w)                          # back to the common start state
        w.arr("gp")[...] = w.arr("up")
    # the fused kernel's deposit against the oracle's, on the same pushed state
    drift = []
    for it in range(8):
        w.step()
        (b.step if path == "wm_step" else b.time_loop)(2, nx + 1, 1)
        if it == 0:
            assert rel_err(b.download_work("uj")[2:-2, 2:-2, 2:-2], w.arr("uj")[2:-2, 2:-2, 2:-2]) < 1e-12
        uf = b.empty("uf")
        b.download(uf=uf)               # NB settles the lazy permutation on odd steps only below
        drift.append(rel_err(uf, w.arr("uf")))
        res, rho = b.gauss()
        assert res < 1e-13 * max(rho, 1.0)
        assert b.stats()["cg_iterations"] == w.cg_iterations()
    assert drift[0] < 1e-10 and drift[-1] < 1e-8, drift
    _compare_particles(b, w)
    # four more steps WITHOUT any reader in between: every fused launch reads through the previous step's pending permutation
    for _ in range(4):
        w.step()
    (b.step if path == "wm_step" else b.time_loop)(2, nx + 1, 4)
    _compare_particles(b, w)
    uf = b.empty("uf")
    b.download(uf=uf)
    assert rel_err(uf, w.arr("uf")) < 1e-8
    assert b.stats()["error_flags"] == 0
    b.close(); w.close()


def test_ragged_cells():
    """x -> 2 + (nx - 4) ((x - 2)/nx)^3 piles the load up near the low-x edge: cells with a hundred particles and more next to
    thinly filled and empty ones; the Weibel loop then runs as usual"""
    nx, ny, nz, n0 = 24, 6, 6, 12
    w = make_world3(nx, ny, nz, n0, np_factor=3)
    up, gp, np2 = w.arr("up"), w.arr("gp"), w.arr("np2")
    gp[...] = up
    m = active_mask(np2, w.np)
    x = gp[..., 0]
    x[m] = 2.0 + (nx - 4) * ((x[m] - 2.0) / nx) ** 3
    w.sort_bucket()
    w.arr("gp")[...] = w.arr("up")
    cc = w.arr("cumcnt")
    per_cell = np.diff(cc, axis=-1)
    assert per_cell.max() >= 100 and (per_cell == 0).any() and ((per_cell > 16) & (per_cell <= 32)).any()
    b = backend_for(w)
    upload_from_world(b, w)
    for it in range(6):
        w.step()
        b.step(2, nx + 1, 1)
        res, rho = b.gauss()
        assert res < 1e-13 * max(rho, 1.0)
        if it % 2 == 1:                # a reader every other step: pending and settled permutations both occur
            uf = b.empty("uf")
            b.download(uf=uf)
            assert rel_err(uf, w.arr("uf")) < 1e-8
    _compare_particles(b, w)
    assert b.stats()["error_flags"] == 0
    b.close(); w.close()


def test_c1_full_size_gpu_vs_oracle():
    """BASELINE.json configs[0]: 2d/proj/weibel/config_sample.json -- 256 x 256 cells, n_ppc = 20, np = 5 n_ppc nx
    (2d/proj/weibel/app.f90:248,275), omega_pe = 0.1, v_th = 0.1, t_ani = 5 -- at full size, GPU against the oracle."""
    nx = ny = 256
    nppc = 20
    q, r, _ = weibel_constants(nppc)
    w = World2(nx, ny, 5 * nppc * nx, q=q, r=r)
    w.load_weibel(nppc, v_thi=0.1, v_the=0.1, t_ani=5.0, b0=0.0)
    w.step()
    b = backend_for(w)
    upload_from_world(b, w)
    ntot = int(w.arr("np2").sum())
    assert ntot == 2 * nppc * nx * ny
    for it in range(3):
        w.step()
        b.step(2, nx + 1, 1)
        assert w.error() == 0
        uf = b.empty("uf")
        b.download(uf=uf)
        assert rel_err(uf, w.arr("uf")) < 1e-9, it
        res, rho = b.gauss()
        assert res < 1e-13 * max(rho, 1.0)
        assert b.stats()["cg_iterations"] == w.cg_iterations()
    up, np2, cc = b.empty("up"), b.empty("np2"), b.empty("cumcnt")
    b.download(up, np2, cc)
    assert np.array_equal(np2, w.arr("np2")) and np.array_equal(cc, w.arr("cumcnt"))
    m = active_mask(np2, w.np)
    # per-pencil multisets of IDs (vectorised: sort every pencil's IDs)
    ids_g = np.where(m, up[..., -1].view(np.int64), np.iinfo(np.int64).max)
    ids_r = np.where(m, w.arr("up")[..., -1].view(np.int64), np.iinfo(np.int64).max)
    assert np.array_equal(np.sort(ids_g, axis=-1), np.sort(ids_r, axis=-1))
    assert b.stats()["n_particles"] == ntot and b.stats()["error_flags"] == 0
    b.close(); w.close()
