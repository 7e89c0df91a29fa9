"""Generates tests/golden/ref_cases.npz FROM THE TRANSLATED REFERENCE (run from the repo root, in the container that holds
/root/reference):

    python tests/golden/make_ref_fixtures.py

For every case the start state is the oracle loader's (a Philox Maxwellian load, "identical initial particle loads and seeds";
the reference's own loader draws from the non-reproducible random_number) and the END state -- fields with ghosts, np2, cumcnt and
every particle record in the reference's own order -- is what the reference's Fortran procedures compute from it, translated by
oracle/f2cxx and driven in the driver's call order (oracle/f2cxx/pyref.py).  No oracle arithmetic enters the end state.
tests/test_ref_golden.py holds the oracle to these vectors bit for bit (anywhere: no /root/reference needed) and checks, where
the translated reference can be built, that the committed file is what it produces today; tests/test_zzz_gpu_ref_golden.py holds
the CUDA path to them within the stated tolerances."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.f2cxx import pyref  # noqa: E402
from tests.util import make_world2, make_world3  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "ref_cases.npz")

# name: dim, nx, ny, nz, n0, bc, order, u0, steps  -- the shapes tests/test_gpu_parity_variants.py and __graft_entry__.smoke() use
CASES = {
    "weibel3d": (3, 12, 8, 8, 6, 0, 0, 0.0, 4),
    "weibel2d": (2, 18, 14, 0, 8, 0, 0, 0.0, 6),
    "reconnection2d": (2, 18, 14, 0, 8, 1, 1, 0.0, 4),
    "shock2d": (2, 18, 14, 0, 8, 2, 2, 0.3, 4),
    "reconnection3d": (3, 18, 8, 6, 4, 1, 1, 0.0, 4),
    "shock3d": (3, 18, 8, 6, 4, 2, 2, 0.3, 4),
}


def start_world(name):
    dim, nx, ny, nz, n0, bc, order, u0, steps = CASES[name]
    return make_world2(nx, ny, n0, bc=bc) if dim == 2 else make_world3(nx, ny, nz, n0, bc=bc)


def pack(up, np2):
    """records of all pencils back to back, pencil order = np2's memory order"""
    flat_n = np2.reshape(-1)
    rec = up.reshape(len(flat_n), -1, up.shape[-1])
    return np.concatenate([rec[p, :flat_n[p]] for p in range(len(flat_n))], axis=0)


def unpack(rec, np2, np_cap):
    flat_n = np2.reshape(-1)
    up = np.zeros((len(flat_n), np_cap, rec.shape[-1]))
    off = np.concatenate([[0], np.cumsum(flat_n)])
    for p in range(len(flat_n)):
        up[p, :flat_n[p]] = rec[off[p]:off[p + 1]]
    return up.reshape(np2.shape + (np_cap, rec.shape[-1]))


def build_case(name):
    dim, nx, ny, nz, n0, bc, order, u0, steps = CASES[name]
    w = start_world(name)
    out = {"np_cap": w.np, "q": w.q.copy(), "r": w.r.copy(), "rec0": pack(w.arr("up"), w.arr("np2")), "np2_0": w.arr("np2").copy(),
           "cumcnt_0": w.arr("cumcnt").copy(), "uf_0": w.arr("uf").copy()}
    R = pyref.RefWorld(dim, nx, ny, nz, w.np, q=w.q, r=w.r, bc=bc, bounds=True)
    for k in ("up", "gp", "uf", "np2", "cumcnt"):
        R.arr(k)[...] = w.arr(k)
    for _ in range(steps):
        R.step(order, u0)
    out.update(rec1=pack(R.arr("up"), R.arr("np2")), np2_1=R.arr("np2").copy(), cumcnt_1=R.arr("cumcnt").copy(), uf_1=R.arr("uf").copy())
    w.close()
    return out


def build():
    out = {}
    for name in CASES:
        for k, v in build_case(name).items():
            out[f"{name}.{k}"] = v
    return out


def load(path=PATH):
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


if __name__ == "__main__":
    out = build()
    np.savez_compressed(PATH, **out)
    print("wrote", PATH, os.path.getsize(PATH), "bytes;", {n: len(out[n + ".rec0"]) for n in CASES})
