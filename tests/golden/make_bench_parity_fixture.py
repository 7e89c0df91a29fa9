"""Generates tests/golden/bench_parity3d.npz (run from the repo root):
    OMP_NUM_THREADS=1 python tests/golden/make_bench_parity_fixture.py

A small 3-D Weibel box stepped by THE TRANSLATED REFERENCE (oracle/_ref: the reference's Fortran procedures translated to C++ by
oracle/f2cxx and driven in the driver's call order; one rank) from the oracle loader's Philox load: the state after PRE steps
(records packed per pencil, fields with ghosts, the CG warm start df = the SAVEd `df` of field__fdtd_i) and after STEPS more steps
(fields, np2, cumcnt, particle IDs in canonical (pencil, cell, ID) order).  The reference does not report its CG iteration counts
(a local variable of cgm): `cg_1` is the count of the single-threaded oracle, whose fields equal the reference's bit for bit
(tests/test_ref_transpiled.py).  Without the translated reference (no /root/reference, no prebuilt oracle/_ref) build() falls back
to the oracle, which differs by the round-off of its threaded deposit.
bench.py loads it WITHOUT importing oracle/: every rank cuts its own z-slab out of the global start state, runs STEPS steps
through wm_step and compares its slab with the global end state -- so that every bench line (N = 1, 2, 4, 8, lazy sort,
peer-memory cgm, NCCL migration: exactly the path that is timed) carries a parity figure against the oracle in `checks`.
tests/test_bench_parity_fixture.py checks the fixture against a fresh run of the translated reference (exact) and of the oracle,
and the slab cutting logic."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pyoracle  # noqa: E402
from oracle.f2cxx import pyref  # noqa: E402
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


def build_from_oracle():
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


def build_from_reference():
    """the same case stepped by the translated reference; the oracle contributes the load at t = 0 and, run beside it on one
    thread, the CG iteration counts"""
    before = pyoracle.num_threads()
    pyoracle.set_num_threads(1)
    try:
        w = make_world3(NX, NY, NZ, N0, np_factor=3)
        R = pyref.RefWorld(3, NX, NY, NZ, w.np, q=w.q, r=w.r, bounds=True)
        for k in ("up", "gp", "uf", "np2", "cumcnt"):
            R.arr(k)[...] = w.arr(k)
        for _ in range(PRE):
            R.step()
            w.step()
        out = dict(nx=NX, ny=NY, nz=NZ, n0=N0, np_cap=w.np, steps=STEPS, q=w.q, r=w.r,
                   rec0=pack(R), np2_0=R.arr("np2").copy(), cumcnt_0=R.arr("cumcnt").copy(), uf_0=R.arr("uf").copy(),
                   df_0=R.saved("field__fdtd_i", "df").copy())
        for _ in range(STEPS):
            R.step()
            w.step()
        assert w.error() == 0 and np.array_equal(w.arr("uf"), R.arr("uf"))
        out.update(uf_1=R.arr("uf").copy(), np2_1=R.arr("np2").copy(), cumcnt_1=R.arr("cumcnt").copy(), ids_1=canonical_ids(R),
                   cg_1=np.array(w.cg_iterations()))
        return out
    finally:
        pyoracle.set_num_threads(before)


def build():
    return build_from_reference() if pyref.available(3) else build_from_oracle()


if __name__ == "__main__":
    out = build()
    path = os.path.join(HERE, "bench_parity3d.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out["rec0"]), "particles")
