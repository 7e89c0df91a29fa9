"""Shared by the set-up tests and tests/multigpu_check.py: an oracle world loaded with a wumingpic_b200.setups state, and the shock
driver's loop body (step, inject, relocate) run side by side on the oracle and the backend."""
import numpy as np

from oracle.pyoracle import ShockPrm, World2, World3
from wumingpic_b200 import setups


def world_for(s, nproc=1):
    """oracle world of set-up `s` with `nproc` slabs (y-slabs in 2-D, z-slabs in 3-D), loaded rank by rank from the slab loaders"""
    kw = dict(delx=1.0, delt=s.delt, c=s.c, gfac=s.gfac, q=s.q, r=s.r, bc=s.bc)
    w = World2(s.nx, s.ny, s.np_cap, nproc=nproc, **kw) if s.dim == 2 else World3(s.nx, s.ny, s.nz, s.np_cap, nproc_j=1, nproc_k=nproc, **kw)
    load = setups.reconnection_slab if s.name == "reconnection" else setups.shock_slab
    w.set_xrange(s.nxs, s.nxe)
    for rk in range(w.nranks):
        g = w.geom(rk)
        up, np2, cc, uf = load(s, g["nys"], g["nye"], g["nzs"] or 2, g["nze"] or 2)
        w.arr("up", rk)[...] = up
        w.arr("gp", rk)[...] = up
        w.arr("np2", rk)[...] = np2
        w.arr("cumcnt", rk)[...] = cc
        w.arr("uf", rk)[...] = uf
    return w


def oracle_shock_prm(s, seed=20240601):
    e = s.extra
    return ShockPrm(n0=s.n0, v0=e["v0"], v_thi=e["v_thi"], v_the=e["v_the"], b0=e["b0"], theta_bn=e["theta_bn"], phi_bn=e["phi_bn"],
                    l_damp_ini=e["l_damp_ini"], seed=seed)


def id_first_inject(rows, counts, nptotal):
    excl = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
    return np.stack([excl[rows] + nptotal[0], excl[rows] + nptotal[1]])


def id_first_relocate(rows, n0, nptotal):
    g = np.asarray(rows, dtype=np.int64)
    return np.stack([g * n0 + nptotal[0], g * n0 + nptotal[1]])
