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


def test_unavailable_entry_points_say_so(swapped):
    import wumingpic_b200 as wm
    w = make_world3(NX, NY, NZ, N0)
    b = backend_for(w)
    upload_from_world(b, w)
    with pytest.raises(wm.WmError, match="y-slabs"):
        b.load_weibel(N0)
    with pytest.raises(wm.WmError, match="y-slabs"):
        b.pack_particles(0)
    b.close(); w.close()
