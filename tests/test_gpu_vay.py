"""GPU parity of the Vay pusher: particle__solv_vay (3d/common/particle.f90:236-419, 2d/common/particle.f90:182-315)
against the oracle's restatement -- stage-wise on fields strong enough that Vay and Buneman-Boris differ visibly, and
over whole steps with the pusher selected for wm_step (fused kernel and per-procedure kernels).  Tolerances as in
test_gpu_parity3d.py: pushed particles <= 1e-13 relative, fields <= 1e-10 with equal CG counts, index sets exact."""
import numpy as np
import pytest

from tests.util import backend_for, canonical_cells, make_world2, make_world3, rel_err, upload_from_world, active_mask

pytestmark = pytest.mark.gpu
WM_PUSHER_VAY = 1


def _strong_fields(w, seed=7, amp=50.0):
    """overwrite uf by O(amp) fields: q dt B / 2m ~ 0.25 with the test charge, far from the small-angle limit"""
    rng = np.random.default_rng(seed)
    uf = w.arr("uf")
    uf[...] = amp * rng.standard_normal(uf.shape)


@pytest.mark.parametrize("dim", [3, 2])
def test_vay_push_matches_oracle(dim):
    w = make_world3(16, 12, 10, 8, steps=1) if dim == 3 else make_world2(24, 16, 10, steps=1)
    _strong_fields(w)
    b = backend_for(w)
    upload_from_world(b, w)
    nx = w.nx
    w.particle_solv()
    boris = w.arr("gp").copy()
    w.particle_solv_vay()
    b.particle__solv_vay(2, nx + 1)
    gp = b.empty("gp")
    b.download(gp=gp)
    m = active_mask(w.arr("np2"), w.np)
    ref, got = w.arr("gp")[m], gp[m]
    nd = ref.shape[1]
    assert np.array_equal(got[:, nd - 1].view(np.int64), ref[:, nd - 1].view(np.int64))
    for c in range(nd - 1):
        assert rel_err(got[:, c], ref[:, c]) < 1e-13, c
    # the two pushers really differ on these fields (so the test cannot pass by running Buneman-Boris)
    assert np.abs(boris[m][:, :nd - 1] - ref[:, :nd - 1]).max() > 1e-6
    # and the plain entry point still runs Buneman-Boris
    b.particle__solv(2, nx + 1)
    b.download(gp=gp)
    assert rel_err(gp[m][:, :nd - 1], boris[m][:, :nd - 1]) < 1e-13
    b.close(); w.close()


@pytest.mark.parametrize("fused", [1, 0], ids=["fused", "per-procedure"])
@pytest.mark.parametrize("dim", [3, 2])
def test_vay_whole_steps(dim, fused):
    w = make_world3(16, 12, 10, 8, steps=2) if dim == 3 else make_world2(24, 16, 10, steps=2)
    b = backend_for(w)
    upload_from_world(b, w)
    b.set_fused(bool(fused))
    b.set_pusher(WM_PUSHER_VAY)
    w.set_pusher(1)
    nx = w.nx
    for _ in range(4):
        w.step()
        b.step(2, nx + 1, 1)
        assert w.error() == 0
        uf = b.empty("uf")
        b.download(uf=uf)
        assert rel_err(uf, w.arr("uf")) < 1e-9
        res, rho = b.gauss()
        assert res < 1e-13 * max(rho, 1.0)
    up, np2, cc = b.empty("up"), b.empty("np2"), b.empty("cumcnt")
    b.download(up, np2, cc)
    assert np.array_equal(np2, w.arr("np2")) and np.array_equal(cc, w.arr("cumcnt"))
    for (cg, rg), (cr, rr) in zip(canonical_cells(up, np2, cc), canonical_cells(w.arr("up"), w.arr("np2"), w.arr("cumcnt"))):
        assert np.array_equal(cg, cr)
        assert np.array_equal(rg[:, -1].view(np.int64), rr[:, -1].view(np.int64))
        if len(rg):
            assert np.abs(rg[:, :-1] - rr[:, :-1]).max() < 1e-9
    assert b.stats()["error_flags"] == 0
    b.close(); w.close()
