"""Helpers of the shock particle-source tests: a shock world whose box can still grow, the per-row injection counts and the
ID offsets the driver computes on the host (2d/proj/shock/app.f90:711-781)."""
import numpy as np

from oracle.pyoracle import ShockPrm, World2, World3, weibel_constants
from tests.util import active_mask

U0 = -0.3                                   # u_inject < 0: upstream flows towards -x (app.f90:319-321)
V0 = U0 / np.sqrt(1 + U0 * U0)


def shock_prm(n0, seed=20240601):
    return ShockPrm(n0=n0, v0=V0, v_thi=0.02, v_the=0.03, b0=0.05, theta_bn=np.pi / 2, phi_bn=np.pi / 3, l_damp_ini=4.0, seed=seed)


def make_shock_world(dim, nx, ny, nz, n0, nxe0, nproc=1):
    """Weibel-loaded particles squeezed into the cells nxs+1 .. nxe0-2 of a box [2, nxe0] that may grow up to nx+1"""
    q, r, _ = weibel_constants(n0)
    if dim == 2:
        w = World2(nx, ny, n0 * nx * 3, nproc=nproc, q=q, r=r, bc=2)
    else:
        w = World3(nx, ny, nz, n0 * nx * 3, nproc_j=1, nproc_k=nproc, q=q, r=r, bc=2)
    w.load_weibel(n0, b0=0.0)
    w.set_xrange(2, nxe0)
    for rk in range(w.nranks):
        up, gp, np2 = w.arr("up", rk), w.arr("gp", rk), w.arr("np2", rk)
        gp[...] = up
        m = active_mask(np2, w.np)
        x = gp[..., 0]
        x[m] = 3.0 + (x[m] - 2.0) * (nxe0 - 2.0 - 3.0) / nx
    w.sort_bucket()
    for rk in range(w.nranks):
        w.arr("gp", rk)[...] = w.arr("up", rk)
    assert w.error() == 0
    return w


def row_counts(nrows_global, it, n0):
    """a deterministic, uneven pattern incl. empty rows (what nlinj_grid looks like after the random remainders)"""
    rng = np.random.default_rng(1000 + it)
    c = rng.integers(0, 2 * n0 + 1, size=nrows_global).astype(np.int32)
    c[rng.integers(0, nrows_global)] = 0
    return c


def id_first_inject(counts_local_rows_global_index, counts_global, nptotal):
    """id_first[isp, local row] = ncinj_grid(row) + nptotal(isp)   (app.f90:769-781, 839)"""
    excl = np.concatenate([[0], np.cumsum(counts_global)[:-1]]).astype(np.int64)
    return np.stack([excl[counts_local_rows_global_index] + nptotal[0], excl[counts_local_rows_global_index] + nptotal[1]])


def id_first_relocate(local_rows_global_index, n0, nptotal):
    """id_first[isp, local row] = global_row * n0 + nptotal(isp)   (app.f90:670)"""
    g = np.asarray(local_rows_global_index, dtype=np.int64)
    return np.stack([g * n0 + nptotal[0], g * n0 + nptotal[1]])


def local_rows(w, rank, dim):
    """global row index of every local row, in the backend's row order (j fastest, then k)"""
    g = w.geom(rank)
    ny = w.ny
    if dim == 2:
        return np.arange(g["nys"], g["nye"] + 1) - 2
    js = np.arange(g["nys"], g["nye"] + 1) - 2
    ks = np.arange(g["nzs"], g["nze"] + 1) - 2
    return (ks[:, None] * ny + js[None, :]).ravel()


def monotone(cc):
    """the reference leaves cumcnt above nxe stale after inject / relocate; the device keeps it monotone"""
    return np.maximum.accumulate(cc, axis=-1)
