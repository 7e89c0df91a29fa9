"""Shared helpers of the parity tests (oracle <-> CUDA backend)."""
import numpy as np

from oracle.pyoracle import World2, World3, weibel_constants


def make_world3(nx, ny, nz, n0, steps=0, nproc_j=1, nproc_k=1, np_factor=3, seed=20240601, b0=0.0, bc=0, order=0, u0=0.0,
                **kw):
    q, r, _ = weibel_constants(n0)
    w = World3(nx, ny, nz, n0 * nx * np_factor, nproc_j=nproc_j, nproc_k=nproc_k, q=q, r=r, bc=bc, **kw)
    w.load_weibel(n0, b0=b0, seed=seed)
    if bc != 0:
        squeeze_into_walls(w)
    for _ in range(steps):
        w.step(order, u0)
    assert w.error() == 0
    return w


def make_world2(nx, ny, n0, steps=0, nproc=1, np_factor=3, seed=20240601, b0=0.0, bc=0, order=0, u0=0.0, **kw):
    """2-D oracle world with the Weibel load.  bc != 0 (walls): the load is squeezed into the cells nxs+1 .. nxe-2 the
    reflecting walls confine particles to, and re-sorted with the oracle's own sort__bucket."""
    q, r, _ = weibel_constants(n0)
    w = World2(nx, ny, n0 * nx * np_factor, nproc=nproc, q=q, r=r, bc=bc, **kw)
    w.load_weibel(n0, b0=b0, seed=seed)
    if bc != 0:
        squeeze_into_walls(w)
    for _ in range(steps):
        w.step(order, u0)
    assert w.error() == 0
    return w


def squeeze_into_walls(w):
    """x -> nxs+1 + (x - nxgs) (nx-3)/nx for every loaded particle, then sort__bucket (gp -> up, new cumcnt)."""
    nx = w.nx
    for rk in range(w.nranks):
        up, gp, np2 = w.arr("up", rk), w.arr("gp", rk), w.arr("np2", rk)
        gp[...] = up
        m = active_mask(np2, w.np)
        x = gp[..., 0]
        x[m] = 3.0 + (x[m] - 2.0) * (nx - 3.0) / nx
    w.sort_bucket()
    for rk in range(w.nranks):
        w.arr("gp", rk)[...] = w.arr("up", rk)


def backend_for(world, rank=0, device=-1, nproc_j=1, nproc_k=1):
    """A CUDA Backend with the geometry of one oracle rank (3-D: rank = rank_j * nproc_k + rank_k, mpi_set.f90:45-60;
    2-D: y slabs, rank = rank_j)."""
    import wumingpic_b200 as wm
    g = world.geom(rank)
    if isinstance(world, World2):
        return wm.Backend(2, world.np, 2, world.nx + 1, 2, world.ny + 1, nys=g["nys"], nye=g["nye"], delx=world.delx,
                          delt=world.delt, c=world.c, gfac=world.gfac, q=world.q, r=world.r, device=device,
                          bc_kind=world.bc, nproc_j=world.nranks, nproc_k=1, rank_j=rank, rank_k=0)
    return wm.Backend(3, world.np, 2, world.nx + 1, 2, world.ny + 1, 2, world.nz + 1, nys=g["nys"], nye=g["nye"],
                      nzs=g["nzs"], nze=g["nze"], delx=world.delx, delt=world.delt, c=world.c, gfac=world.gfac,
                      q=world.q, r=world.r, device=device, bc_kind=world.bc, nproc_j=nproc_j, nproc_k=nproc_k,
                      rank_j=rank // nproc_k, rank_k=rank % nproc_k)


def upload_from_world(b, world, rank=0):
    b.upload(world.arr("up", rank), world.arr("np2", rank), world.arr("cumcnt", rank), world.arr("uf", rank))
    b.upload_work("df", world.arr("df", rank))   # the SAVEd CG warm start (field.f90:102) travels with the state


def rel_err(a, b):
    """max|a-b| / max|b| (relative to the max-norm of the quantity)."""
    scale = np.abs(b).max()
    return np.abs(a - b).max() / (scale if scale > 0 else 1.0)


def active_mask(np2, np_cap):
    """boolean mask (nsp, [nzl,] nyl, np) of the defined particle slots"""
    return np.arange(np_cap).reshape((1,) * np2.ndim + (-1,)) < np2[..., None]


def canonical_cells(up, np2, cumcnt):
    """Per-pencil records sorted by (x-cell, particle ID): the permutation-invariant form in which the
    reference defines the result of migration + sort.  Returns a list of arrays."""
    out = []
    if np2.ndim == 2:   # 2-D: (nsp, nyl) -> a single k plane
        up, np2, cumcnt = up[:, None], np2[:, None], cumcnt[:, None]
    nsp, nzl, nyl = np2.shape
    for isp in range(nsp):
        for k in range(nzl):
            for j in range(nyl):
                n = np2[isp, k, j]
                rec = up[isp, k, j, :n]
                cc = cumcnt[isp, k, j]
                cell = np.searchsorted(cc, np.arange(n), side="right") - 1
                ids = rec[:, -1].view(np.int64)
                o = np.lexsort((ids, cell))
                out.append((cell[o], rec[o]))
    return out
