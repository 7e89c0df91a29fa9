"""GPU parity tests for the other rows of SURVEY.md section 8(a): the 2-D code path (periodic Weibel, configs[0]) and the
wall set-ups (reconnection: reflecting particle walls + conducting-wall field rules; shock: injection wall + open right
boundary) in 2-D and 3-D, against the CPU oracle, through the C ABI (per-procedure kernels).
Same tolerances as tests/test_gpu_parity3d.py (relative to the max-norm of the quantity):
    gp <= 1e-13, uj <= 1e-12, dB/dE with equal CG iteration counts <= 1e-10, sort/migration index sets and records exact.
"""
import numpy as np
import pytest

from tests.util import (active_mask, backend_for, canonical_cells, make_world2, make_world3, rel_err, upload_from_world)

pytestmark = pytest.mark.gpu

# (dim, bc, order, u0)
VARIANTS = [(2, 0, 0, 0.0), (2, 1, 1, 0.0), (2, 2, 2, 0.3), (3, 1, 1, 0.0), (3, 2, 2, 0.3)]
IDS = ["2d-periodic", "2d-reconnection", "2d-shock", "3d-reconnection", "3d-shock"]
NX = 18


def make(dim, bc, order, u0, steps):
    if dim == 2:
        return make_world2(NX, 14, 8, steps=steps, bc=bc, order=order, u0=u0)
    return make_world3(NX, 8, 6, 6, steps=steps, bc=bc, order=order, u0=u0)


def x_bc(obj, order, u0, nxs, nxe, oracle):
    """the x-boundary call of this set-up's time loop"""
    if order == 2:
        obj.bc_injection(u0) if oracle else obj.bc__injection(nxs, nxe, u0)
    else:
        obj.bc_particle_x() if oracle else obj.bc__particle_x(nxs, nxe)


def yz(obj, oracle):
    if oracle:
        obj.bc_particle_y() if hasattr(obj, "bc_particle_y") else obj.bc_particle_yz()
    else:
        obj.bc__particle_yz()


@pytest.mark.parametrize("dim,bc,order,u0", VARIANTS, ids=IDS)
def test_stagewise(dim, bc, order, u0):
    w = make(dim, bc, order, u0, steps=3)
    b = backend_for(w)
    upload_from_world(b, w)
    nxs, nxe = 2, NX + 1
    U = dim   # index of ux in a record
    # --- push ---
    w.particle_solv()
    b.particle__solv(nxs, nxe)
    gp = b.empty("gp")
    b.download(gp=gp)
    m = active_mask(w.arr("np2"), w.np)
    ref, got = w.arr("gp")[m], gp[m]
    assert np.array_equal(got[:, -1].view(np.int64), ref[:, -1].view(np.int64))
    for c in range(w.ndim - 1):
        assert rel_err(got[:, c], ref[:, c]) < 1e-13, c
    # --- x boundary before the field step (reconnection / shock order) ---
    if order != 0:
        x_bc(w, order, u0, nxs, nxe, True)
        x_bc(b, order, u0, nxs, nxe, False)
        b.download(gp=gp)
        ref, got = w.arr("gp")[m], gp[m]
        assert rel_err(got[:, 0], ref[:, 0]) < 1e-13
        assert rel_err(got[:, U], ref[:, U]) < 1e-13
    # --- field solve, stage by stage ---
    tol = {1: 1e-12, 2: 1e-12, 3: 1e-12, 4: 1e-10, 5: 1e-10, 6: 1e-10, 7: 1e-10, 8: 1e-10}
    inner = (slice(1, -1),) * dim
    inner2 = (slice(2, -2),) * dim
    for stage in range(1, 9):
        w.field_fdtd_i(stage)
        b.field__fdtd_i(nxs, nxe, stage)
        assert w.error() == 0
        if stage in (1, 2):
            got, ref = b.download_work("uj"), w.arr("uj")
            if stage == 2:
                got, ref = got[inner], ref[inner]
        elif stage == 3:
            got, ref = b.download_work("gkl"), w.arr("gkl")
        elif stage in (4, 5, 6, 7):
            got, ref = b.download_work("df"), w.arr("df")
            if stage in (4, 6):
                got, ref = got[inner2], ref[inner2]
        else:
            uf = b.empty("uf")
            b.download(uf=uf)
            got, ref = uf, w.arr("uf")
        assert rel_err(got, ref) < tol[stage], f"stage {stage}"
        if stage == 4:
            assert b.stats()["cg_iterations"] == w.cg_iterations()
    b.close()


@pytest.mark.parametrize("dim,bc,order,u0", VARIANTS, ids=IDS)
def test_boundary_migration_sort_exact(dim, bc, order, u0):
    """x boundary + y(/z) re-binning + counting sort on the oracle's own pushed positions: bit-exact records."""
    w = make(dim, bc, order, u0, steps=3)
    b = backend_for(w)
    upload_from_world(b, w)
    nxs, nxe = 2, NX + 1
    w.particle_solv()
    if order != 0:
        x_bc(w, order, u0, nxs, nxe, True)
    w.field_fdtd_i()
    uf_tmp = w.arr("uf").copy()
    b.h_field__fdtd_i(uf_tmp, w.arr("up"), w.arr("gp"), w.arr("cumcnt"), w.arr("np2"), nxs, nxe)
    if order == 0:
        x_bc(w, order, u0, nxs, nxe, True)
        x_bc(b, order, u0, nxs, nxe, False)
    yz(w, True)
    w.sort_bucket()
    assert w.error() == 0
    yz(b, False)
    b.sort__bucket(nxs, nxe)
    up, np2, cc = b.empty("up"), b.empty("np2"), b.empty("cumcnt")
    b.download(up, np2, cc)
    assert np.array_equal(np2, w.arr("np2"))
    assert np.array_equal(cc, w.arr("cumcnt"))
    for (c_got, r_got), (c_ref, r_ref) in zip(canonical_cells(up, np2, cc),
                                              canonical_cells(w.arr("up"), w.arr("np2"), w.arr("cumcnt"))):
        assert np.array_equal(c_got, c_ref)
        assert np.array_equal(r_got.view(np.int64), r_ref.view(np.int64))
    b.close()


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "per-procedure"])
@pytest.mark.parametrize("dim,bc,order,u0", VARIANTS, ids=IDS)
def test_multistep(dim, bc, order, u0, fused):
    """Whole steps through wm_step: the fused push+deposit kernels (k_fused3 / k_fused2, all three time loops) and the
    per-procedure kernels (set_fused(False): what a driver calling the five entry points one by one runs)."""
    w = make(dim, bc, order, u0, steps=0)
    b = backend_for(w)
    b.set_fused(fused)
    upload_from_world(b, w)
    ntot = int(w.arr("np2").sum())
    np.testing.assert_allclose(b.energy(), w.energy(), rtol=1e-12)
    drift = []
    for it in range(1, 9):
        w.step(order, u0)
        b.step(2, NX + 1, 1, order, u0)
        assert w.error() == 0
        uf = b.empty("uf")
        b.download(uf=uf)
        drift.append(rel_err(uf, w.arr("uf")))
        res, rho = b.gauss()
        assert res < 1e-13 * max(rho, 1.0), f"Gauss residual {res} at step {it}"
        st = b.stats()
        assert st["n_particles"] == ntot and st["error_flags"] == 0
    assert drift[0] < 1e-10 and drift[-1] < 1e-8, drift
    up, np2, cc = b.empty("up"), b.empty("np2"), b.empty("cumcnt")
    b.download(up, np2, cc)
    assert np.array_equal(np2, w.arr("np2")), "particle index sets diverged"
    worst = 0.0
    for (cg, rg), (cr, rr) in zip(canonical_cells(up, np2, cc),
                                  canonical_cells(w.arr("up"), w.arr("np2"), w.arr("cumcnt"))):
        assert np.array_equal(cg, cr)
        assert np.array_equal(rg[:, -1].view(np.int64), rr[:, -1].view(np.int64))
        if len(rg):
            worst = max(worst, np.abs(rg[:, :-1] - rr[:, :-1]).max())
    assert worst < 1e-9, worst
    np.testing.assert_allclose(b.energy(), w.energy(), rtol=1e-9)
    b.close()


def test_2d_host_buffer_step_and_weibel_loader():
    """wm_h_step in 2-D (configs[0] is the reference's CPU-runnable 2-D Weibel case) and the device-side Weibel loader
    against the oracle's (same Philox counters): positions bit-identical, Maxwellian momenta to libm ulp."""
    import wumingpic_b200 as wm
    w = make_world2(16, 12, 6)
    b = backend_for(w)
    b.load_weibel(6)
    up, np2, cc, uf = b.empty("up"), b.empty("np2"), b.empty("cumcnt"), b.empty("uf")
    b.download(up, np2, cc, uf)
    m = active_mask(np2, w.np)
    assert np.array_equal(np2, w.arr("np2")) and np.array_equal(cc, w.arr("cumcnt"))
    assert np.array_equal(up[m][:, :2], w.arr("up")[m][:, :2])
    assert np.array_equal(up[m][:, 5].view(np.int64), w.arr("up")[m][:, 5].view(np.int64))
    np.testing.assert_allclose(up[m][:, 2:5], w.arr("up")[m][:, 2:5], rtol=0, atol=1e-15)
    b.close()
    b = backend_for(w)
    up, uf, np2, cc = (w.arr(k).copy() for k in ("up", "uf", "np2", "cumcnt"))
    b.h_step(up, uf, np2, cc, 2, 17)
    w.step()
    assert np.array_equal(np2, w.arr("np2"))
    assert rel_err(uf, w.arr("uf")) < 1e-10
    b.close()
    assert wm.backend.WM_BC_SHOCK == 2


@pytest.mark.parametrize("dim,bc,order,u0", [(3, 0, 0, 0.0)] + VARIANTS, ids=["3d-periodic"] + IDS)
def test_moments(dim, bc, order, u0):
    """wm_mom_calc = mom_calc__accl + mom_calc__nvt + bc__mom against the oracle: sums over <= 8 x ppc
    contributions per node in a different order -> 1e-12 relative to the max-norm of each moment."""
    w = make(dim, bc, order, u0, steps=3)
    b = backend_for(w)
    upload_from_world(b, w)
    got = b.mom_calc(2, NX + 1)
    w.mom_calc()
    ref = w.arr("mom")
    inner = (slice(None),) + (slice(1, -1),) * dim     # after bc__mom only the interior nodes are meaningful
    for l in range(7):
        assert rel_err(got[inner][..., l], ref[inner][..., l]) < 1e-12, l
    assert abs(got[inner][..., 0].sum() - w.arr("np2").sum()) < 1e-9 * w.arr("np2").sum()
    # the moment block leaves the particle state untouched
    up, np2, cc = b.empty("up"), b.empty("np2"), b.empty("cumcnt")
    b.download(up, np2, cc)
    m = active_mask(np2, w.np)
    assert np.array_equal(up[m].view(np.int64), w.arr("up")[m].view(np.int64))
    b.close()
