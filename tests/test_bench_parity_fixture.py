"""The golden state bench.py checks every rank against (tests/golden/bench_parity3d.npz): (1) it is what the translated
reference produces today, bit for bit, and what the oracle produces to the round-off of its threaded deposit; (2) the slab cutting / comparing logic of tests/fixture_slabs.py is right -- the oracle's OWN N-slab emulation, started
from the slabs cut out of the global state, must land on the global end state."""
import numpy as np
import pytest

from tests import fixture_slabs as fs
from oracle.f2cxx import pyref
from tests.golden.make_bench_parity_fixture import build_from_oracle, build_from_reference
from tests.util import make_world3


@pytest.mark.skipif(not pyref.available(3), reason="the translated reference cannot be built here")
def test_fixture_is_the_translated_reference_output():
    fx, now = fs.load(), build_from_reference()
    assert sorted(fx) == sorted(now)
    for k in now:
        a, b = np.asarray(now[k]), np.asarray(fx[k])
        assert a.shape == b.shape, k
        assert np.array_equal(a.view(np.int64) if a.dtype == np.float64 else a, b.view(np.int64) if b.dtype == np.float64 else b), k


def test_fixture_is_current_oracle_output():
    fx, now = fs.load(), build_from_oracle()
    for k in ("np2_0", "cumcnt_0", "np2_1", "cumcnt_1", "ids_1"):
        assert np.array_equal(fx[k], now[k]), k
    assert np.array_equal(fx["rec0"][:, -1].view(np.int64), now["rec0"][:, -1].view(np.int64))   # the IDs, bit-cast
    # the OpenMP deposit reduction is not bit-reproducible: fields, and particles pushed by them, agree to round-off
    assert np.abs(fx["rec0"][:, :-1] - now["rec0"][:, :-1]).max() <= 1e-12
    for k in ("uf_0", "df_0", "uf_1"):
        assert np.abs(fx[k] - now[k]).max() <= 1e-12 * np.abs(now[k]).max(), k


@pytest.mark.parametrize("nranks", [1, 2, 4, 8])
def test_slabs_of_the_fixture_step_to_its_end_state(nranks):
    fx = fs.load()
    nx, ny, nz, n0 = (int(fx[k]) for k in ("nx", "ny", "nz", "n0"))
    w = make_world3(nx, ny, nz, n0, nproc_k=nranks, np_factor=3)
    assert w.np == int(fx["np_cap"])
    for rk in range(nranks):
        g = w.geom(rk)
        st = fs.slab_state(fx, g["nzs"], g["nze"])
        for name in ("up", "np2", "cumcnt", "uf", "df"):
            w.arr(name, rk)[...] = st[name]
        w.arr("gp", rk)[...] = st["up"]
    for _ in range(int(fx["steps"])):
        w.step()
    assert w.error() == 0
    for rk in range(nranks):
        g = w.geom(rk)
        res = fs.compare_slab(fx, g["nzs"], g["nze"], w.arr("up", rk), w.arr("np2", rk), w.arr("cumcnt", rk), w.arr("uf", rk))
        assert res["np2_equal"] and res["cumcnt_equal"] and res["ids_equal"], (rk, res)
        assert res["uf_rel_err"] < 1e-9, (rk, res)
