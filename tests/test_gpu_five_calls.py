"""The reference drivers' own call sequence -- five procedure calls per step (3d/proj/weibel/app.f90:100-108; reconnection
3d/proj/reconnection/app.f90:103-108; shock 2d/proj/shock/app.f90:112-118) -- on device-resident state must (a) give the oracle's
result for every set-up and (b) run the fast kernels: particle__solv is deferred and field__fdtd_i launches the fused
push + boundary + deposit kernel, which shows in the launch count."""
import numpy as np
import pytest

from tests.test_gpu_parity_variants import IDS, NX, VARIANTS, make
from tests.util import active_mask, backend_for, canonical_cells, rel_err, upload_from_world

pytestmark = pytest.mark.gpu
ALL = [(3, 0, 0, 0.0)] + VARIANTS
ALL_IDS = ["3d-periodic"] + IDS


def _make(dim, bc, order, u0, steps):
    if dim == 3 and bc == 0:
        from tests.util import make_world3
        return make_world3(NX, 8, 6, 6, steps=steps)
    return make(dim, bc, order, u0, steps)


@pytest.mark.parametrize("vay", [False, True], ids=["boris", "vay"])
@pytest.mark.parametrize("dim,bc,order,u0", ALL, ids=ALL_IDS)
def test_five_calls_match_oracle(dim, bc, order, u0, vay):
    w = _make(dim, bc, order, u0, steps=1)
    w.set_pusher(1 if vay else 0)
    b = backend_for(w)
    upload_from_world(b, w)
    nxe = NX + 1
    for it in range(6):
        w.step(order, u0)
        b.time_loop(2, nxe, 1, order, u0, vay=vay)
        assert w.error() == 0
        if it in (1, 4):      # readers at irregular intervals: the permutation is pending across the other steps
            uf = b.empty("uf")
            b.download(uf=uf)
            assert rel_err(uf, w.arr("uf")) < 1e-8, it
        res, rho = b.gauss() if it % 2 else (0.0, 1.0)
        assert res < 1e-13 * max(rho, 1.0)
    up, np2, cc, uf = b.empty("up"), b.empty("np2"), b.empty("cumcnt"), b.empty("uf")
    b.download(up, np2, cc, uf)
    assert np.array_equal(np2, w.arr("np2")) and np.array_equal(cc, w.arr("cumcnt"))
    assert rel_err(uf, w.arr("uf")) < 1e-8
    for (cg, rg), (cr, rr) in zip(canonical_cells(up, np2, cc), canonical_cells(w.arr("up"), w.arr("np2"), w.arr("cumcnt"))):
        assert np.array_equal(cg, cr)
        assert np.array_equal(rg[:, -1].view(np.int64), rr[:, -1].view(np.int64))
        if len(rg):
            assert np.abs(rg[:, :-1] - rr[:, :-1]).max() < 1e-9
    assert b.stats()["error_flags"] == 0
    b.close(); w.close()


def test_five_calls_launch_the_fused_kernels():
    """same number of kernel launches per step as wm_step, and far fewer than the per-procedure kernels"""
    from tests.util import make_world3
    w = make_world3(NX, 8, 6, 6, steps=1)
    counts = {}
    for mode in ("wm_step", "five", "slow"):
        b = backend_for(w)
        b.set_fused(mode != "slow")
        upload_from_world(b, w)
        (b.step if mode == "wm_step" else b.time_loop)(2, NX + 1, 2)
        n0 = b.launch_count()
        (b.step if mode == "wm_step" else b.time_loop)(2, NX + 1, 3)
        counts[mode] = b.launch_count() - n0
        b.close()
    assert counts["five"] == counts["wm_step"], counts
    assert counts["slow"] > counts["five"], counts


def test_gp_reader_between_solv_and_fdtd_runs_the_real_push():
    """a driver (or test) that reads gp after particle__solv gets the pushed set -- the deferral is invisible"""
    from tests.util import make_world3
    w = make_world3(NX, 8, 6, 6, steps=2)
    b = backend_for(w)
    upload_from_world(b, w)
    w.particle_solv()
    b.particle__solv(2, NX + 1)
    gp = b.empty("gp")
    b.download(gp=gp)
    m = active_mask(w.arr("np2"), w.np)
    for c in range(6):
        assert rel_err(gp[m][:, c], w.arr("gp")[m][:, c]) < 1e-13
    # ... and the step completes on the per-procedure kernels from there
    w.field_fdtd_i(); w.bc_particle_x(); w.bc_particle_yz(); w.sort_bucket()
    b.field__fdtd_i(2, NX + 1); b.bc__particle_x(2, NX + 1); b.bc__particle_yz(); b.sort__bucket(2, NX + 1)
    uf, np2 = b.empty("uf"), b.empty("np2")
    b.download(uf=uf, np2=np2)
    assert np.array_equal(np2, w.arr("np2")) and rel_err(uf, w.arr("uf")) < 1e-10
    b.close(); w.close()


def test_upload_with_shock_shaped_cumcnt():
    """The shock driver leaves cumcnt above nxe stale: inject()/relocate() bump np2 and cumcnt(nxe) only and init never fills
    cumcnt(nxe+1:) (2d/proj/shock/app.f90:346-359, 836-838).  wm_upload must take cell membership from the
    entries up to nxe and the pencil population from np2 -- never from cumcnt(nxe+1)."""
    from tests.shock_util import U0, make_shock_world
    nx, nxe = 22, 17
    w = make_shock_world(3, nx, 8, 6, 6, nxe)
    cc_stale = w.arr("cumcnt").copy()
    assert (cc_stale[..., nxe - 2] == w.arr("np2")).all()          # cumcnt(nxe) == np2: every particle is below cell nxe
    cc_stale[..., nxe + 1 - 2:] = 0                                # the tail the shock driver never writes
    outs = []
    for cc in (w.arr("cumcnt"), cc_stale):
        b = backend_for(w)
        b.upload(w.arr("up"), w.arr("np2"), np.ascontiguousarray(cc), w.arr("uf"))
        b.upload_work("df", w.arr("df"))
        b.step(2, nxe, 2, 2, U0)
        up, np2, c2, uf = b.empty("up"), b.empty("np2"), b.empty("cumcnt"), b.empty("uf")
        b.download(up, np2, c2, uf)
        assert b.stats()["error_flags"] == 0 and b.stats()["n_particles"] == int(w.arr("np2").sum())
        outs.append((up[active_mask(np2, w.np)].copy(), np2.copy(), c2[..., :nxe - 1].copy(), uf.copy()))
        b.close()
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])
    # same particles in the same (deterministic) order; coordinates to round-off (the J sums of step 1 are RED.F64 in any order)
    assert np.array_equal(outs[0][0][:, -1].view(np.int64), outs[1][0][:, -1].view(np.int64))
    assert np.abs(outs[0][0][:, :-1] - outs[1][0][:, :-1]).max() < 1e-12
    assert rel_err(outs[1][3], outs[0][3]) < 1e-12
    w.step(2, U0); w.step(2, U0)
    assert np.array_equal(outs[1][1], w.arr("np2"))
    w.close()
