"""Host-side integer bookkeeping of the shock driver's inject() (2d/proj/shock/app.f90:711-781, 3d :747-812): how many
particles each rank and each row receive.  It stays on the host (a few integers per row); the particles themselves are
created on the device by Backend.shock_inject (wm_shock.cu)."""
import numpy as np


def inject_counts(n0, v0, delt, delx, ny, nz, nproc, rows_of_rank, rng):
    """Returns (nginj_proc[nproc], [nlinj_grid of every rank]) following the three steps of the reference:
    (1) total = int(pflux) (+1 with probability frac(pflux)), pflux = n0 |v0| delt delx ny nz   (app.f90:711-715)
    (2) equal share per rank, the remainder to randomly chosen ranks                            (:718-726)
    (3) equal share per row of the rank, the remainder to randomly chosen rows                  (:732-743)
    rows_of_rank[r] = number of rows (nyl*nzl) of rank r; rng = numpy Generator (the reference uses random_number)."""
    pflux = n0 * abs(v0) * delt * delx * ny * nz
    nginj = int(pflux)
    if rng.random() < pflux - int(pflux):
        nginj += 1
    per_rank = np.full(nproc, nginj // nproc, dtype=np.int64)
    per_rank[rng.permutation(nproc)[: nginj % nproc]] += 1
    grids = []
    for r in range(nproc):
        nrow = rows_of_rank[r]
        g = np.full(nrow, per_rank[r] // nrow, dtype=np.int32)
        g[rng.permutation(nrow)[: per_rank[r] % nrow]] += 1
        grids.append(g)
    return per_rank, grids
