"""BASELINE.json configs[0] ("C1"): the reference's own CPU case, 2d/proj/weibel/config_sample.json -- 256 x 256 cells,
n_ppc = 20, two species of equal mass, omega_pe = 0.1, v_th = 0.1, t_ani = 5, four MPI ranks (y slabs).  The oracle runs it at full
size with the four ranks emulated and with one rank: identical cell-sorted particle sets, fields to round-off, Gauss residual at
round-off, particle count conserved."""
import numpy as np

from oracle.pyoracle import World2, weibel_constants
from tests.util import canonical_cells

NX = NY = 256
NPPC = 20
STEPS = 3


def _run(nproc):
    q, r, _ = weibel_constants(NPPC, mass_ratio=1.0, sigma_e=0.0, omega_pe=0.1)
    w = World2(NX, NY, 5 * NPPC * NX, nproc=nproc, q=q, r=r)          # np = 5 n_ppc nx, 2d/proj/weibel/app.f90:275
    w.load_weibel(NPPC, v_thi=0.1, v_the=0.1, t_ani=5.0, b0=0.0)
    n0 = sum(int(w.arr("np2", rk).sum()) for rk in range(w.nranks))
    res = []
    for _ in range(STEPS):
        w.step()
        assert w.error() == 0
        res.append(w.gauss())
    uf = np.zeros((NY, NX, 6))
    cells = {}
    for rk in range(w.nranks):
        g = w.geom(rk)
        uf[g["nys"] - 2:g["nye"] - 1] = w.arr("uf", rk)[2:-2, 2:-2]
        np2 = w.arr("np2", rk)
        keys = [(s, j + g["nys"]) for s in range(2) for j in range(np2.shape[1])]
        for (cell, rec), key in zip(canonical_cells(w.arr("up", rk), np2, w.arr("cumcnt", rk)), keys):
            cells[key] = (cell, rec)
    n1 = sum(int(w.arr("np2", rk).sum()) for rk in range(w.nranks))
    it = w.cg_iterations()
    w.close()
    return uf, cells, it, res, n0, n1


def test_c1_four_ranks_equal_one_rank():
    u1, c1, it1, res1, n0, n1 = _run(1)
    u4, c4, it4, res4, m0, m1 = _run(4)
    assert n0 == n1 == m0 == m1 == 2 * NPPC * NX * NY                   # 2.62 M particles
    assert it1 == it4
    for res, rho in res1 + res4:
        assert res < 1e-13 * max(rho, 1.0)
    assert np.abs(u1 - u4).max() < 1e-12 * np.abs(u1).max()
    assert c1.keys() == c4.keys()
    for key in c1:
        assert np.array_equal(c1[key][0], c4[key][0])
        assert np.array_equal(c1[key][1][:, 5].view(np.int64), c4[key][1][:, 5].view(np.int64))
        assert np.abs(c1[key][1][:, :5] - c4[key][1][:, :5]).max() < 1e-11
