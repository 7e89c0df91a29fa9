#!/usr/bin/env python
"""time_c1_reference.py -- BASELINE.json configs[0] as the reference runs it, timed on this host's cores: 2d/proj/weibel/config_sample.json
(256 x 256 cells, n_ppc 20, np = 5 n_ppc nx, `mpiexec -np 4`) through oracle/_ref -- the reference's own 2-D source translated to C++
(oracle/f2cxx), compiled -O3 -march=native -frounding-math, four flat-MPI ranks on four host threads (oracle/f2cxx/mpi_threads.cpp).
Uses oracle/: a measurement aid beside bench.py's CPU legs, not part of the product.   python tools/time_c1_reference.py [steps] [ranks]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.f2cxx import pyref  # noqa: E402
from oracle.pyoracle import World2, weibel_constants  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
nproc = int(sys.argv[2]) if len(sys.argv) > 2 else 4
nx = ny = 256
nppc = 20
q, r, _ = weibel_constants(nppc, mass_ratio=1.0, sigma_e=0.0, omega_pe=0.1)
w = World2(nx, ny, 5 * nppc * nx, nproc=nproc, q=q, r=r)
w.load_weibel(nppc, v_thi=0.1, v_the=0.1, t_ani=5.0, b0=0.0)
R = pyref.RefWorld(2, nx, ny, 0, w.np, nproc_j=nproc, q=q, r=r, fast=True, native_mpi=nproc > 1)
npart = 0
for rk in range(nproc):
    for k in ("up", "uf", "np2", "cumcnt"):
        R.arr(k, rk)[...] = w.arr(k, rk)
    npart += int(w.arr("np2", rk).sum())
w.close()
R.run_steps(2)
t0 = time.perf_counter()
R.run_steps(steps)
dt = time.perf_counter() - t0
assert sum(int(R.arr("np2", rk).sum()) for rk in range(nproc)) == npart
print(json.dumps({"config": "C1: 2-D Weibel 256x256, n_ppc 20, 2 species (2d/proj/weibel/config_sample.json)", "ranks": nproc,
                  "host_threads_available": len(os.sched_getaffinity(0)), "particles": npart, "steps": steps,
                  "s_per_step": dt / steps, "particle_updates_per_s": npart * steps / dt,
                  "what": "oracle/_ref: the reference's 2-D Fortran source translated to C++, -O3 -march=native, flat MPI on host threads"}))
