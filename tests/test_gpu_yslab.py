"""3-D y-slabs (the reference's nproc_j > 1, nproc_k = 1 -- every shipped 3-D sample, 3d/proj/*/config_sample.json): the device runs
them as z-slabs of the exactly relabelled system (x, y' = z, z' = y), B' = -(Bx, Bz, By) (wm_internal.cuh, wm_ctx::swap_yz).  On one
GPU the relabelling is forced with WM_SWAP_YZ=1 and must be invisible: every host-visible result equals the oracle's in the caller's
own layout.  The two-rank y-slab run against the oracle's nproc_j = 2 emulation is in tests/test_gpu_multi.py."""
import os

import numpy as np
import pytest

from tests.util import active_mask, backend_for, canonical_cells, make_world3, rel_err, upload_from_world

pytestmark = pytest.mark.gpu
NX, NY, NZ, N0 = 14, 10, 6, 6            # ny != nz: a transposition slip cannot cancel


@pytest.fixture()
def swapped():
    os.environ["WM_SWAP_YZ"] = "1"
    yield
    os.environ.pop("WM_SWAP_YZ", None)


def test_roundtrip_is_bit_exact(swapped):
    w = make_world3(NX, NY, NZ, N0, steps=2)
    b = backend_for(w)
    upload_from_world(b, w)
    up, np2, cc, uf = b.empty("up"), b.empty("np2"), b.empty("cumcnt"), b.empty("uf")
    b.download(up, np2, cc, uf)
    assert np.array_equal(np2, w.arr("np2")) and np.array_equal(cc, w.arr("cumcnt")) and np.array_equal(uf, w.arr("uf"))
    m = active_mask(np2, w.np)
    assert np.array_equal(up[m].view(np.int64), w.arr("up")[m].view(np.int64))
    np.testing.assert_allclose(b.energy(), w.energy(), rtol=1e-12)
    b.close(); w.close()


@pytest.mark.parametrize("bc,order,u0", [(0, 0, 0.0), (1, 1, 0.0), (2, 2, 0.3)], ids=["periodic", "reconnection", "shock"])
def test_steps_match_oracle(swapped, bc, order, u0):
    w = make_world3(NX, NY, NZ, N0, steps=1, bc=bc, order=order, u0=u0)
    b = backend_for(w)
    upload_from_world(b, w)
    for it in range(8):
        w.step(order, u0)
        (b.step if it % 2 else b.time_loop)(2, NX + 1, 1, order, u0)
        uf = b.empty("uf")
        b.download(uf=uf)
        assert rel_err(uf, w.arr("uf")) < 1e-8, it
        res, rho = b.gauss()
        assert res < 1e-13 * max(rho, 1.0)
    up, np2, cc = b.empty("up"), b.empty("np2"), b.empty("cumcnt")
    b.download(up, np2, cc)
    assert np.array_equal(np2, w.arr("np2")) and np.array_equal(cc, w.arr("cumcnt"))
    for (cg, rg), (cr, rr) in zip(canonical_cells(up, np2, cc), canonical_cells(w.arr("up"), w.arr("np2"), w.arr("cumcnt"))):
        assert np.array_equal(cg, cr) and np.array_equal(rg[:, -1].view(np.int64), rr[:, -1].view(np.int64))
        if len(rg):
            assert np.abs(rg[:, :-1] - rr[:, :-1]).max() < 1e-9
    got = b.mom_calc(2, NX + 1)
    w.mom_calc()
    inner = (slice(None),) + (slice(1, -1),) * 3
    for l in range(7):
        assert rel_err(got[inner][..., l], w.arr("mom")[inner][..., l]) < 1e-9, l
    assert b.stats()["error_flags"] == 0
    b.close(); w.close()


def test_device_weibel_loader_under_relabelling(swapped):
    """wm_load_weibel generates the CALLER's load: same Philox stream per caller pencil, y / z and the anisotropic uz in the caller's
    columns -- positions bit-identical to the oracle's loader, Maxwellian momenta to libm ulp"""
    w = make_world3(NX, NY, NZ, N0, b0=0.3)
    b = backend_for(w)
    b.load_weibel(N0, b0=0.3)
    up, np2, cc, uf = b.empty("up"), b.empty("np2"), b.empty("cumcnt"), b.empty("uf")
    b.download(up, np2, cc, uf)
    m = active_mask(np2, w.np)
    assert np.array_equal(np2, w.arr("np2")) and np.array_equal(cc, w.arr("cumcnt")) and np.array_equal(uf, w.arr("uf"))
    assert np.array_equal(up[m][:, :3], w.arr("up")[m][:, :3])
    assert np.array_equal(up[m][:, 6].view(np.int64), w.arr("up")[m][:, 6].view(np.int64))
    np.testing.assert_allclose(up[m][:, 3:6], w.arr("up")[m][:, 3:6], rtol=0, atol=1e-15)
    b.close(); w.close()


def test_pack_and_unavailable_entry_points(swapped):
    """particle packing gives the caller's records (as a set per species: the pencil order is (j, k) under y-slabs); the device-side
    shock source and the stage-wise work arrays answer WM_ERR_STATE with a message"""
    import wumingpic_b200 as wm
    w = make_world3(NX, NY, NZ, N0, steps=1)
    b = backend_for(w)
    upload_from_world(b, w)
    rec, lc = b.pack_particles(0)
    ref, lcr = w.pack_particles(0)
    assert np.array_equal(lc, lcr)
    for lo, hi in ((0, lc[0]), (lc[0], lc[0] + lc[1])):
        a, r = rec[lo:hi], ref[lo:hi]
        a, r = a[np.argsort(a[:, 6].view(np.int64))], r[np.argsort(r[:, 6].view(np.int64))]
        assert np.array_equal(a.view(np.int64), r.view(np.int64))
    with pytest.raises(wm.WmError, match="y-slabs"):
        b.download_work("uj")
    prm = wm.ShockParams(n0=N0, v0=-0.3, v_thi=0.0, v_the=0.0, b0=0.0, theta_bn=0.0, phi_bn=0.0, l_damp_ini=4.0, seed=1)
    with pytest.raises(wm.WmError, match="y-slabs"):
        b.shock_inject(prm, NX, np.zeros(NY * NZ, dtype=np.int32), np.zeros((2, NY * NZ), dtype=np.int64), 1)
    b.close(); w.close()
