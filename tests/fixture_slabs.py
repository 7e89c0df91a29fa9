"""Cutting per-rank z-slabs out of the global golden state of tests/golden/bench_parity3d.npz and comparing a rank's result
with the global end state.  Pure numpy, NO oracle import: bench.py uses this before its timed region so that every bench
line carries a parity figure against the oracle's committed output (`checks.parity`), on exactly the path that is timed."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bench_parity3d.npz")


def load(path=GOLDEN):
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def slab_state(fx, nzs, nze):
    """host arrays (reference layout, C order = reversed Fortran shape) of the rank that owns the global planes nzs..nze
    (global cell indices start at 2): up(nsp,nzl,nyl,np,7), np2, cumcnt, uf and df with two ghost planes on either side"""
    nx, ny, nz, npc = int(fx["nx"]), int(fx["ny"]), int(fx["nz"]), int(fx["np_cap"])
    k0, k1 = nzs - 2, nze - 2 + 1                        # 0-based plane range
    np2g, ccg = fx["np2_0"], fx["cumcnt_0"]
    off = np.concatenate([[0], np.cumsum(np2g.reshape(-1))])
    nzl = k1 - k0
    up = np.zeros((2, nzl, ny, npc, 7))
    for isp in range(2):
        for k in range(k0, k1):
            for j in range(ny):
                p = (isp * nz + k) * ny + j
                up[isp, k - k0, j, :np2g[isp, k, j]] = fx["rec0"][off[p]:off[p + 1]]
    # box planes: index kk of the global box <-> global plane kk (ghosts 0,1 and nz+2,nz+3); the slab needs nzs-2 .. nze+2
    planes = np.arange(nzs - 2, nze + 3)
    interior = (planes - 2) % nz + 2                     # periodic image inside the global interior
    return dict(up=up, np2=np.ascontiguousarray(np2g[:, k0:k1]), cumcnt=np.ascontiguousarray(ccg[:, k0:k1]),
                uf=np.ascontiguousarray(fx["uf_0"][interior]), df=np.ascontiguousarray(fx["df_0"][interior]))


def canonical_ids(up, np2, cumcnt):
    """particle IDs in (species, k, j, cell, ID) order"""
    out = []
    nsp, nzl, nyl = np2.shape
    for isp in range(nsp):
        for k in range(nzl):
            for j in range(nyl):
                n = np2[isp, k, j]
                ids = up[isp, k, j, :n, -1].view(np.int64)
                cell = np.searchsorted(cumcnt[isp, k, j], np.arange(n), side="right") - 1
                out.append(ids[np.lexsort((ids, cell))])
    return np.concatenate(out) if out else np.zeros(0, np.int64)


def compare_slab(fx, nzs, nze, up, np2, cumcnt, uf):
    """-> dict(uf_rel_err, np2_equal, cumcnt_equal, ids_equal) of this rank's slab against the golden end state"""
    ny, nz = int(fx["ny"]), int(fx["nz"])
    k0, k1 = nzs - 2, nze - 2 + 1
    ref_uf = fx["uf_1"][nzs:nze + 1, 2:-2, 2:-2]         # interior planes of the slab, interior in x and y
    got_uf = uf[2:-2, 2:-2, 2:-2]
    scale = np.abs(fx["uf_1"]).max()
    err = float(np.abs(got_uf - ref_uf).max() / (scale if scale > 0 else 1.0))
    np2g, ccg = fx["np2_1"], fx["cumcnt_1"]
    off = np.concatenate([[0], np.cumsum(np2g.reshape(-1))])
    ref_ids = np.concatenate([fx["ids_1"][off[(isp * nz + k0) * ny]:off[(isp * nz + k1) * ny]] for isp in range(2)])
    got_ids = canonical_ids(up, np2, cumcnt)
    return dict(uf_rel_err=err, np2_equal=bool(np.array_equal(np2, np2g[:, k0:k1])),
                cumcnt_equal=bool(np.array_equal(cumcnt, ccg[:, k0:k1])),
                ids_equal=bool(len(got_ids) == len(ref_ids) and np.array_equal(got_ids, ref_ids)))
