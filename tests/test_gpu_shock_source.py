"""GPU parity of the shock driver's particle source on the device: wm_shock_inject / wm_shock_relocate
against the oracle's restatement of inject() / relocate() (2d/proj/shock/app.f90:615-852, 3d :644-906), inside the shock
time loop (solv, injection wall, fdtd_i, particle_y[z], sort, inject, relocate: app.f90:110-125) with a growing box.
Index sets, IDs and cumcnt (made monotone above nxe, see wm_shock.cu) must match exactly; coordinates to 1e-12, uf to 1e-9."""
import numpy as np
import pytest

import wumingpic_b200 as wm
from tests.shock_util import (U0, id_first_inject, id_first_relocate, local_rows, make_shock_world, monotone, row_counts,
                              shock_prm)
from tests.util import backend_for, canonical_cells, rel_err, upload_from_world

pytestmark = pytest.mark.gpu


def _prm_c(p):
    return wm.ShockParams(n0=p.n0, v0=p.v0, v_thi=p.v_thi, v_the=p.v_the, b0=p.b0, theta_bn=p.theta_bn, phi_bn=p.phi_bn,
                          l_damp_ini=p.l_damp_ini, seed=p.seed)


@pytest.mark.parametrize("fused", [1, 0], ids=["fused", "per-procedure"])
@pytest.mark.parametrize("dim", [2, 3])
def test_shock_loop_with_device_source(dim, fused):
    n0, nx, ny, nz, nxe = 6, 22, 10, 4, 14
    w = make_shock_world(dim, nx, ny, nz, n0, nxe)
    prm = shock_prm(n0)
    prm_c = _prm_c(prm)
    b = backend_for(w)
    upload_from_world(b, w)
    b.set_fused(bool(fused))
    rows = local_rows(w, 0, dim)
    nrows = len(rows)
    for it in range(1, 7):
        w.step(2, U0)
        b.step(2, nxe, 1, 2, U0)
        assert w.error() == 0
        # inject: the host keeps the integer bookkeeping
        counts = row_counts(nrows, it, n0)
        nptotal = w.arr("np2").reshape(2, -1).sum(axis=1)
        w.shock_inject(prm, counts, it)
        b.shock_inject(prm_c, nxe, counts[rows], id_first_inject(rows, counts, nptotal), it)
        if it % 2 == 0:          # intvl_expand = 2
            nptotal = w.arr("np2").reshape(2, -1).sum(axis=1)
            w.shock_relocate(prm, it)
            nxe += 1
            assert w.nxe_now == nxe
            b.shock_relocate(prm_c, nxe, id_first_relocate(rows, n0, nptotal), it)
        assert w.error() == 0
        up, np2, cc, uf = b.empty("up"), b.empty("np2"), b.empty("cumcnt"), b.empty("uf")
        b.download(up, np2, cc, uf)
        assert np.array_equal(np2, w.arr("np2")), it
        cc_ref = monotone(w.arr("cumcnt"))
        # entries i = nxgs .. nxe+1 are defined (sort.f90:62-86 writes no further; beyond them the reference keeps stale values)
        assert np.array_equal(cc[..., :nxe], cc_ref[..., :nxe]), it
        assert rel_err(uf, w.arr("uf")) < 1e-9, it
        for (cg, rg), (cr, rr) in zip(canonical_cells(up, np2, cc), canonical_cells(w.arr("up"), w.arr("np2"), cc_ref)):
            assert np.array_equal(cg, cr)
            assert np.array_equal(rg[:, -1].view(np.int64), rr[:, -1].view(np.int64)), it
            if len(rg):
                assert np.abs(rg[:, :-1] - rr[:, :-1]).max() < 1e-12, it
    assert b.stats()["error_flags"] == 0
    assert b.stats()["n_particles"] == int(w.arr("np2").sum())
    b.close(); w.close()


def test_source_argument_errors():
    w = make_shock_world(2, 16, 6, 1, 4, 10)
    b = backend_for(w)
    upload_from_world(b, w)
    prm_c = _prm_c(shock_prm(4))
    idf = np.zeros((2, 6), dtype=np.int64)
    with pytest.raises(wm.WmError):
        b.shock_inject(prm_c, 40, np.zeros(6, dtype=np.int32), idf, 1)      # nxe outside the box
    with pytest.raises(wm.WmError):
        b.shock_inject(prm_c, 10, -np.ones(6, dtype=np.int32), idf, 1)      # negative count
    b.particle__solv(2, 10)
    with pytest.raises(wm.WmError):
        b.shock_inject(prm_c, 10, np.ones(6, dtype=np.int32), idf, 1)       # pushed set pending: must follow sort__bucket
    b.close(); w.close()
