"""The oracle against golden vectors produced by the reference's own (translated) source: tests/golden/ref_cases.npz, written by
tests/golden/make_ref_fixtures.py from oracle/_ref (the reference's Fortran files translated to C++ by oracle/f2cxx and run in the
container that holds /root/reference).  Runs anywhere -- the vectors are committed: from each case's start state the oracle (one
OpenMP thread = the serial program) must land on the reference's end state BIT FOR BIT: fields incl. ghosts, np2, cumcnt and every
particle record in order.  Where the translated reference can be (re)built, the committed file must be what it produces today."""
import numpy as np
import pytest

from oracle import pyoracle
from oracle.f2cxx import pyref
from tests.golden import make_ref_fixtures as mk
from tests.util import active_mask


@pytest.fixture(scope="module")
def golden():
    return mk.load()


@pytest.fixture()
def one_thread():
    before = pyoracle.num_threads()
    pyoracle.set_num_threads(1)
    yield
    pyoracle.set_num_threads(before)


def seeded_world(golden, name):
    """an oracle world holding the case's start state"""
    w = mk.start_world(name)
    g = {k.split(".", 1)[1]: v for k, v in golden.items() if k.startswith(name + ".")}
    assert w.np == int(g["np_cap"]) and np.array_equal(w.q, g["q"]) and np.array_equal(w.r, g["r"])
    w.arr("np2")[...] = g["np2_0"]
    w.arr("cumcnt")[...] = g["cumcnt_0"]
    w.arr("uf")[...] = g["uf_0"]
    w.arr("up")[...] = mk.unpack(g["rec0"], g["np2_0"], w.np)
    w.arr("gp")[...] = w.arr("up")
    w.arr("df")[...] = 0.0
    return w, g


@pytest.mark.parametrize("name", list(mk.CASES))
def test_oracle_lands_on_the_reference_end_state_bit_for_bit(golden, one_thread, name):
    dim, nx, ny, nz, n0, bc, order, u0, steps = mk.CASES[name]
    w, g = seeded_world(golden, name)
    for _ in range(steps):
        w.step(order, u0)
    assert w.error() == 0
    assert np.array_equal(w.arr("np2"), g["np2_1"]) and np.array_equal(w.arr("cumcnt"), g["cumcnt_1"])
    assert np.array_equal(w.arr("uf"), g["uf_1"]), float(np.abs(w.arr("uf") - g["uf_1"]).max())
    m = active_mask(w.arr("np2"), w.np)
    assert np.array_equal(w.arr("up")[m].view(np.int64), g["rec1"].view(np.int64))
    assert np.abs(g["uf_1"] - g["uf_0"]).max() > 1e-6          # the case is not a fixed point
    w.close()


@pytest.mark.parametrize("name", ["weibel3d", "weibel2d"])
def test_threaded_oracle_differs_from_the_reference_by_round_off_only(golden, name):
    """the OpenMP deposit reduction re-associates the sum over threads: same index sets, fields to 1e-13"""
    dim, nx, ny, nz, n0, bc, order, u0, steps = mk.CASES[name]
    before = pyoracle.num_threads()
    pyoracle.set_num_threads(4)
    try:
        w, g = seeded_world(golden, name)
        for _ in range(steps):
            w.step(order, u0)
        assert np.array_equal(w.arr("np2"), g["np2_1"]) and np.array_equal(w.arr("cumcnt"), g["cumcnt_1"])
        assert np.abs(w.arr("uf") - g["uf_1"]).max() <= 1e-13 * np.abs(g["uf_1"]).max()
        w.close()
    finally:
        pyoracle.set_num_threads(before)


@pytest.mark.skipif(not (pyref.available(3) and pyref.available(2)), reason="the translated reference cannot be built here")
def test_committed_vectors_are_what_the_translated_reference_produces_today(golden):
    now = mk.build()
    assert sorted(now) == sorted(golden)
    for k in now:
        a, b = np.asarray(now[k]), np.asarray(golden[k])
        assert a.shape == b.shape and a.dtype == b.dtype, k
        assert np.array_equal(a.view(np.int64) if a.dtype == np.float64 else a, b.view(np.int64) if b.dtype == np.float64 else b), k
