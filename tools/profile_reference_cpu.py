#!/usr/bin/env python
"""profile_reference_cpu.py -- where the reference's CPU time goes, MEASURED (SURVEY.md 8a's last column is inferred from the loop
structure: the reference ships no profile).  The translated reference (oracle/_ref, -O3 -march=native) runs the bounded C2 sample
(3-D Weibel nx x ny x 4, 64 ppc x 2) as flat MPI with one rank per host thread; every procedure call of the drivers' time loop is
timed on every rank (a rank's time in a procedure includes its waits for neighbours inside MPI_SENDRECV / MPI_ALLREDUCE).  ele_cur
-- the deposit, a private procedure inside field__fdtd_i -- is separated as field__fdtd_i's time minus the same call on a box of
the same size WITHOUT particles but with white-noise fields (so that the three CG solves iterate as they do on a white right-hand
side: 13-14 iterations).   Uses oracle/: a measurement aid, not part of the product.
    python tools/profile_reference_cpu.py [nx ny ppc steps]"""
import collections
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import pyoracle  # noqa: E402
from oracle.f2cxx import pyref  # noqa: E402
from oracle.pyoracle import World3, weibel_constants  # noqa: E402

nx, ny, n0, steps = (int(v) for v in (sys.argv[1:5] + ["256", "256", "64", "3"][len(sys.argv) - 1:]))
nz = 4
cores = len(os.sched_getaffinity(0))
nj, nk = bench.rank_grid(cores, ny, nz)
q, r, _ = weibel_constants(n0)
cap = int(n0 * nx * 1.5)


def run(load):
    pyoracle.set_num_threads(cores, fast=True)
    w = World3(nx, ny, nz, cap, nproc_j=nj, nproc_k=nk, q=q, r=r, fast=True)
    if load:
        w.load_weibel(n0)
    else:
        import numpy as np
        rng = np.random.default_rng(1)
        for rk in range(nj * nk):
            w.arr("uf", rk)[...] = 1e-3 * rng.standard_normal(w.arr("uf", rk).shape)
    R = pyref.RefWorld(3, nx, ny, nz, cap, nproc_j=nj, nproc_k=nk, q=q, r=r, fast=True, native_mpi=nj * nk > 1)
    R.pin_ranks(sorted(os.sched_getaffinity(0)))
    R._all(lambda rk: [R.arr(k, rk).__setitem__(Ellipsis, w.arr(k, rk)) for k in ("up", "uf", "np2", "cumcnt")])
    npart = sum(int(w.arr("np2", rk).sum()) for rk in range(nj * nk))
    w.close()
    T = collections.defaultdict(float)
    orig = pyref._Rank.call

    def timed(self, name, *a):
        t = time.perf_counter()
        orig(self, name, *a)
        T[name] += time.perf_counter() - t
    R.run_steps(1)
    pyref._Rank.call = timed
    t0 = time.perf_counter()
    R.run_steps(steps)
    wall = time.perf_counter() - t0
    pyref._Rank.call = orig
    R.close()
    return npart, wall / steps, {k: v / steps / (nj * nk) for k, v in T.items()}


npart, wall, full = run(True)
_, wall0, empty = run(False)
dep = full["field__fdtd_i"] - empty["field__fdtd_i"]
rows = [("particle__solv (gather + push)", full["particle__solv"]),
        ("ele_cur (Esirkepov deposit; field__fdtd_i minus the same call on an empty box)", dep),
        ("field__fdtd_i without the deposit (curre, gkl, 3 x cgm + phi, dfield, dE)", empty["field__fdtd_i"]),
        ("boundary_periodic__particle_x", full["boundary_periodic__particle_x"]),
        ("boundary_periodic__particle_yz (re-binning + migration)", full["boundary_periodic__particle_yz"]),
        ("sort__bucket", full["sort__bucket"])]
tot = sum(v for _, v in rows)
out = {"sample": f"3-D Weibel {nx}x{ny}x{nz}, {n0} ppc x 2 = {npart} particles, flat MPI {nj} x {nk} ranks on {cores} host threads, "
                 f"{steps} steps; translated reference -O3 -march=native", "s_per_step_wall": wall,
       "particle_updates_per_s": npart / wall,
       "mean_seconds_per_rank_and_step": {k: v for k, v in rows}, "share": {k: v / tot for k, v in rows}}
print(json.dumps(out, indent=1))
