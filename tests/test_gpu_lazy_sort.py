"""The lazy sort's state machine (wm_sort.cu: wm_step leaves the permutation pending, the next fused kernel reads through it,
every other consumer applies it first): interleavings of wm_step with the entry points that read the sorted set must give
the oracle's result, and the pending and the settled form of the same state must be the same state."""
import numpy as np
import pytest

from tests.util import active_mask, backend_for, canonical_cells, make_world3, rel_err, upload_from_world

pytestmark = pytest.mark.gpu
NX, NY, NZ, N0 = 14, 8, 6, 6


def _same_as_oracle(b, w, tol=1e-9):
    up, np2, cc, uf = b.empty("up"), b.empty("np2"), b.empty("cumcnt"), b.empty("uf")
    b.download(up, np2, cc, uf)
    assert np.array_equal(np2, w.arr("np2")) and np.array_equal(cc, w.arr("cumcnt"))
    assert rel_err(uf, w.arr("uf")) < 1e-9
    for (cg, rg), (cr, rr) in zip(canonical_cells(up, np2, cc), canonical_cells(w.arr("up"), w.arr("np2"), w.arr("cumcnt"))):
        assert np.array_equal(cg, cr)
        assert np.array_equal(rg[:, -1].view(np.int64), rr[:, -1].view(np.int64))
        if len(rg):
            assert np.abs(rg[:, :-1] - rr[:, :-1]).max() < tol


def test_interleaved_consumers():
    w = make_world3(NX, NY, NZ, N0, steps=1)
    b = backend_for(w)
    upload_from_world(b, w)
    nxe = NX + 1
    # two steps in one call (the second reads through the first one's permutation), then a download (settles)
    w.step(); w.step()
    b.step(2, nxe, 2)
    _same_as_oracle(b, w)
    # step, moments (settle), step, energy (settle), step with the per-procedure kernels (settle at entry), step fused again
    w.step(); b.step(2, nxe, 1)
    got = b.mom_calc(2, nxe); w.mom_calc()
    inner = (slice(None),) + (slice(1, -1),) * 3
    assert rel_err(got[inner], w.arr("mom")[inner]) < 1e-9
    w.step(); b.step(2, nxe, 1)
    np.testing.assert_allclose(b.energy(), w.energy(), rtol=1e-9)
    b.set_fused(False)
    w.step(); b.step(2, nxe, 1)
    b.set_fused(True)
    w.step(); b.step(2, nxe, 1)
    # explicit settle, then the five per-procedure calls on the settled state
    b.settle()
    w.step()
    b.particle__solv(2, nxe); b.field__fdtd_i(2, nxe); b.bc__particle_x(2, nxe); b.bc__particle_yz(); b.sort__bucket(2, nxe)
    _same_as_oracle(b, w)
    assert b.stats()["error_flags"] == 0
    b.close(); w.close()


def test_pending_and_settled_forms_are_the_same_state():
    """two backends from the same state: one keeps stepping with the permutation pending, the other settles after every step;
    both must hold bit-identical particle records (the sort is deterministic) and fields equal to round-off of the J sums"""
    w = make_world3(NX, NY, NZ, N0, steps=1)
    a, b = backend_for(w), backend_for(w)
    for x in (a, b):
        upload_from_world(x, w)
    nxe = NX + 1
    for _ in range(4):
        a.step(2, nxe, 1)
        b.step(2, nxe, 1)
        b.settle()
    out = []
    for x in (a, b):
        up, np2, cc, uf = x.empty("up"), x.empty("np2"), x.empty("cumcnt"), x.empty("uf")
        x.download(up, np2, cc, uf)
        out.append((up[active_mask(np2, w.np)].copy(), np2.copy(), cc.copy(), uf.copy()))
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])
    assert np.abs(out[0][0][:, :6] - out[1][0][:, :6]).max() < 1e-12
    assert np.array_equal(out[0][0][:, 6].view(np.int64), out[1][0][:, 6].view(np.int64))
    assert rel_err(out[0][3], out[1][3]) < 1e-12
    a.close(); b.close(); w.close()
