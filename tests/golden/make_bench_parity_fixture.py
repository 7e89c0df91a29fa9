"""Generates tests/golden/bench_parity3d.npz (run from the repo root):
    OMP_NUM_THREADS=1 python tests/golden/make_bench_parity_fixture.py

A small 3-D Weibel box stepped by the ORACLE (one emulated rank): the state before (records packed per pencil, fields with
ghosts, the CG warm start df) and after STEPS steps (fields, np2, cumcnt, particle IDs in canonical (pencil, cell, ID) order).
bench.py loads it WITHOUT importing oracle/: every rank cuts its own z-slab out of the global start state, runs STEPS steps
through wm_step and compares its slab with the global end state -- so that every bench line (N = 1, 2, 4, 8, lazy sort,
peer-memory cgm, NCCL migration: exactly the path that is timed) carries a parity figure against the oracle in `checks`.
tests/test_bench_parity_fixture.py checks the fixture against a fresh oracle run (not gpu) and the slab cutting logic.
Not a reference golden: the reference has none for this path and cannot be built here (no Fortran compiler / MPI)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.util import canonical_cells, make_world3  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
NX, NY, NZ, N0, PRE, STEPS = 8, 6, 32, 3, 2, 3


def pack(w):
    """records of all pencils back to back (pencil order = np2's memory order) -- the padded `up` is mostly empty"""
    up, np2 = w.arr("up"), w.arr("np2")
    flat_n = np2.reshape(-1)
    rec = up.reshape(-1, w.np, 7)
    return np.concatenate([rec[p, :flat_n[p]] for p in range(len(flat_n))], axis=0)


def canonical_ids(w):
    return np.concatenate([r[:, -1].view(np.int64) for _, r in canonical_cells(w.arr("up"), w.arr("np2"), w.arr("cumcnt"))])


def build():
    w = make_world3(NX, NY, NZ, N0, steps=PRE, np_factor=3)
    out = dict(nx=NX, ny=NY, nz=NZ, n0=N0, np_cap=w.np, steps=STEPS, q=w.q, r=w.r,
               rec0=pack(w), np2_0=w.arr("np2").copy(), cumcnt_0=w.arr("cumcnt").copy(), uf_0=w.arr("uf").copy(),
               df_0=w.arr("df").copy())
    for _ in range(STEPS):
        w.step()
    assert w.error() == 0
    out.update(uf_1=w.arr("uf").copy(), np2_1=w.arr("np2").copy(), cumcnt_1=w.arr("cumcnt").copy(), ids_1=canonical_ids(w),
               cg_1=np.array(w.cg_iterations()))
    return out


if __name__ == "__main__":
    out = build()
    path = os.path.join(HERE, "bench_parity3d.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out["rec0"]), "particles")
