"""Generates the committed regression fixtures from the oracle (run from the repo root):
    OMP_NUM_THREADS=1 python tests/golden/make_oracle_fixtures.py
These pin the oracle against accidental edits.  They are not reference goldens -- the reference has
none for this path and cannot be built here (no Fortran compiler / MPI)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.util import make_world2, make_world3  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    nx, ny, nz, n0, steps = 10, 8, 6, 4, 4
    w = make_world3(nx, ny, nz, n0, steps=steps)
    np.savez_compressed(os.path.join(HERE, "oracle3d_weibel_small.npz"), nx=nx, ny=ny, nz=nz, n0=n0, steps=steps,
                        uf=w.arr("uf").copy(), np2=w.arr("np2").copy(), energy=w.energy())
    print("wrote oracle3d_weibel_small.npz")
    nx, ny, n0, steps, u0 = 12, 8, 5, 4, 0.3
    out = dict(nx=nx, ny=ny, n0=n0, steps=steps, u0=u0)
    for name, bc, order in (("weibel", 0, 0), ("reconnection", 1, 1), ("shock", 2, 2)):
        w = make_world2(nx, ny, n0, steps=steps, bc=bc, order=order, u0=u0)
        out["uf_" + name], out["np2_" + name], out["energy_" + name] = w.arr("uf").copy(), w.arr("np2").copy(), w.energy()
    np.savez_compressed(os.path.join(HERE, "oracle2d_small.npz"), **out)
    print("wrote oracle2d_small.npz")
