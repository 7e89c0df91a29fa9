"""GPU parity tests (3-D, periodic): the CUDA backend against the CPU oracle, stage by stage and
over whole steps, all through the C ABI.  Tolerances are relative to the max-norm of the quantity
(SURVEY.md Appendix A.10):
    gp after push            <= 1e-13   (FMA contraction vs the oracle's unfused arithmetic)
    uj after deposit         <= 1e-12   (re-ordered sums over <= 128 x 27 contributions)
    dB / dE, equal CG count  <= 1e-10   (CG amplifies 1e-12 input differences by kappa ~ 4)
    sort / migration         exact equality of np2, cumcnt and of per-cell particle records
"""
import numpy as np
import pytest

from tests.util import backend_for, canonical_cells, make_world3, rel_err, upload_from_world, active_mask

pytestmark = pytest.mark.gpu

NX, NY, NZ, N0 = 16, 12, 10, 8


@pytest.fixture()
def pair():
    w = make_world3(NX, NY, NZ, N0, steps=3)     # a few oracle steps so that E, B != 0
    b = backend_for(w)
    upload_from_world(b, w)
    yield w, b
    b.close()
    w.close()


def test_upload_download_roundtrip(pair):
    w, b = pair
    up, np2, cc, uf = b.empty("up"), b.empty("np2"), b.empty("cumcnt"), b.empty("uf")
    b.download(up, np2, cc, uf)
    assert np.array_equal(np2, w.arr("np2")) and np.array_equal(cc, w.arr("cumcnt"))
    assert np.array_equal(uf, w.arr("uf"))
    m = active_mask(np2, w.np)
    assert np.array_equal(up[m].view(np.int64), w.arr("up")[m].view(np.int64))   # bit-exact incl. the 64-bit IDs


def test_push_matches_oracle(pair):
    w, b = pair
    w.particle_solv()
    b.particle__solv(2, NX + 1)
    gp = b.empty("gp")
    b.download(gp=gp)
    m = active_mask(w.arr("np2"), w.np)
    ref = w.arr("gp")[m]
    got = gp[m]
    assert np.array_equal(got[:, 6].view(np.int64), ref[:, 6].view(np.int64))
    for c in range(6):
        assert rel_err(got[:, c], ref[:, c]) < 1e-13, c


def test_field_solve_stagewise(pair):
    w, b = pair
    w.particle_solv()
    b.particle__solv(2, NX + 1)
    tol = {1: 1e-12, 2: 1e-12, 3: 1e-12, 4: 1e-10, 5: 1e-10, 6: 1e-10, 7: 1e-10, 8: 1e-10}
    for stage in range(1, 9):
        w.field_fdtd_i(stage)
        b.field__fdtd_i(2, NX + 1, stage)
        assert w.error() == 0
        if stage in (1, 2):
            got, ref = b.download_work("uj"), w.arr("uj")
            if stage == 2:   # after curre only interior + first ghost layer are defined identically
                got, ref = got[1:-1, 1:-1, 1:-1], ref[1:-1, 1:-1, 1:-1]
        elif stage == 3:
            got, ref = b.download_work("gkl"), w.arr("gkl")
        elif stage in (4, 5, 6, 7):
            got, ref = b.download_work("df"), w.arr("df")
            if stage in (4, 6):   # ghosts are refreshed by the following dfield call
                got, ref = got[2:-2, 2:-2, 2:-2], ref[2:-2, 2:-2, 2:-2]
        else:
            uf = b.empty("uf")
            b.download(uf=uf)
            got, ref = uf, w.arr("uf")
        assert rel_err(got, ref) < tol[stage], f"stage {stage}"
        if stage == 4:
            assert b.stats()["cg_iterations"] == w.cg_iterations()


def test_boundary_migration_sort_exact(pair):
    """x wrap + y/z re-binning + counting sort: index sets and records must match bit-exactly.  The pushed
    positions handed to the device are the oracle's own gp (through the host-buffer form of field__fdtd_i,
    which stages up AND gp), so that no 1e-16 push difference can flip a particle across a cell edge."""
    w, b = pair
    w.particle_solv()
    w.field_fdtd_i()
    uf_tmp = w.arr("uf").copy()
    b.h_field__fdtd_i(uf_tmp, w.arr("up"), w.arr("gp"), w.arr("cumcnt"), w.arr("np2"), 2, NX + 1)
    w.bc_particle_x(); w.bc_particle_yz(); w.sort_bucket()
    assert w.error() == 0
    b.bc__particle_x(2, NX + 1); b.bc__particle_yz(); b.sort__bucket(2, NX + 1)
    up, np2, cc = b.empty("up"), b.empty("np2"), b.empty("cumcnt")
    b.download(up, np2, cc)
    assert np.array_equal(np2, w.arr("np2"))
    assert np.array_equal(cc, w.arr("cumcnt"))
    moved = 0
    for (c_got, r_got), (c_ref, r_ref) in zip(canonical_cells(up, np2, cc),
                                              canonical_cells(w.arr("up"), w.arr("np2"), w.arr("cumcnt"))):
        assert np.array_equal(c_got, c_ref)
        assert np.array_equal(r_got.view(np.int64), r_ref.view(np.int64))   # bit-exact records
        moved += len(r_got)
    assert moved == int(np2.sum())


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "per-procedure"])
def test_multistep_drift_and_invariants(fused):
    """Whole steps against the oracle, through wm_step's fused kernel + deterministic sort and through the
    per-procedure kernels (what a driver calling the five entry points runs)."""
    w = make_world3(NX, NY, NZ, N0)
    b = backend_for(w)
    b.set_fused(fused)
    upload_from_world(b, w)
    ntot = int(w.arr("np2").sum())
    e0 = b.energy()
    np.testing.assert_allclose(e0, w.energy(), rtol=1e-12)
    drift = []
    for it in range(1, 9):
        w.step()
        b.step(2, NX + 1, 1)
        uf = b.empty("uf")
        b.download(uf=uf)
        drift.append(rel_err(uf, w.arr("uf")))
        res, rho = b.gauss()
        assert res < 1e-13 * max(rho, 1.0), f"Gauss residual {res} at step {it}"
        st = b.stats()
        assert st["n_particles"] == ntot and st["error_flags"] == 0
    # per-step drift of E,B against the oracle (CG tolerance 1e-6 bounds it, equal iteration counts keep it ~1e-12)
    assert drift[0] < 1e-10 and drift[-1] < 1e-8, drift
    up, np2, cc = b.empty("up"), b.empty("np2"), b.empty("cumcnt")
    b.download(up, np2, cc)
    assert np.array_equal(np2, w.arr("np2")), "particle index sets diverged"
    got = canonical_cells(up, np2, cc)
    ref = canonical_cells(w.arr("up"), w.arr("np2"), w.arr("cumcnt"))
    worst = 0.0
    for (cg, rg), (cr, rr) in zip(got, ref):
        assert np.array_equal(cg, cr)
        assert np.array_equal(rg[:, 6].view(np.int64), rr[:, 6].view(np.int64))
        if len(rg):
            worst = max(worst, np.abs(rg[:, :6] - rr[:, :6]).max())
    assert worst < 1e-9, worst
    np.testing.assert_allclose(b.energy(), w.energy(), rtol=1e-9)
    b.close()


def test_fused_path_is_deterministic():
    """The fused path's sort is a stable, atomic-free scatter: two runs from the same state give bit-identical
    particle arrays (J sums still use global RED.F64 whose order may differ, so fields agree to round-off)."""
    outs = []
    w = make_world3(NX, NY, NZ, N0, steps=1)    # one host state (the OpenMP oracle itself is not bit-reproducible)
    for rep in range(2):
        b = backend_for(w)
        upload_from_world(b, w)
        b.step(2, NX + 1, 1)
        up, np2, cc = b.empty("up"), b.empty("np2"), b.empty("cumcnt")
        b.download(up, np2, cc)
        outs.append((up[active_mask(np2, w.np)].view(np.int64).copy(), np2.copy(), cc.copy()))
        b.close()
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])
    assert np.array_equal(outs[0][0], outs[1][0])


def test_odd_grid_fused():
    """nx not a multiple of the 16-cell CTA group, tiny ny/nz (every cell touches the periodic wrap)."""
    w = make_world3(21, 3, 4, 5, steps=1)
    b = backend_for(w)
    upload_from_world(b, w)
    for _ in range(3):
        w.step()
        b.step(2, 22, 1)
    uf, np2 = b.empty("uf"), b.empty("np2")
    b.download(uf=uf, np2=np2)
    assert np.array_equal(np2, w.arr("np2"))
    assert rel_err(uf, w.arr("uf")) < 1e-9
    res, rho = b.gauss()
    assert res < 1e-13 * max(rho, 1.0)
    b.close()


def test_host_buffer_step_matches_resident():
    """wm_h_step (upload, step, download: the e2e path) equals the resident path."""
    w = make_world3(8, 6, 6, 4, steps=2)
    b = backend_for(w)
    up, uf, np2, cc = (w.arr(k).copy() for k in ("up", "uf", "np2", "cumcnt"))
    b.upload_work("df", w.arr("df"))
    b.h_step(up, uf, np2, cc, 2, 9)
    w.step()
    assert np.array_equal(np2, w.arr("np2"))
    assert rel_err(uf, w.arr("uf")) < 1e-10
    b.close()


def test_full_size_properties():
    """Size-independent properties at a GPU-sized load (device-generated Weibel state): particle count
    conserved, Gauss residual at round-off every step, energy conserved to the scheme's accuracy."""
    import wumingpic_b200 as wm
    nx, ny, nz, n0 = 64, 64, 32, 32
    q, r, _ = wm.weibel_constants(n0)
    b = wm.Backend(3, n0 * nx * 3, 2, nx + 1, 2, ny + 1, 2, nz + 1, q=q, r=r)
    b.load_weibel(n0)
    ntot = 2 * n0 * nx * ny * nz
    e0 = b.energy().sum()
    for it in range(5):
        b.step(2, nx + 1, 1)
        res, rho = b.gauss()
        assert res < 1e-13 * max(rho, 1.0)
        st = b.stats()
        assert st["n_particles"] == ntot and st["error_flags"] == 0 and st["max_np2"] <= n0 * nx * 3
        assert all(1 <= i < 100 for i in st["cg_iterations"])
    assert abs(b.energy().sum() - e0) / e0 < 1e-3
    b.close()
