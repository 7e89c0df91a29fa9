"""The CUDA path against golden vectors produced by the reference's own (translated) source -- tests/golden/ref_cases.npz, see
tests/golden/make_ref_fixtures.py: the end states were computed by the reference's Fortran procedures (translated to C++ by
oracle/f2cxx, run in the container that holds /root/reference); no oracle arithmetic is in them.  The oracle classes serve here
only as containers for geometry and host arrays; nothing is stepped on the CPU.

Tolerances (relative to the max-norm; SURVEY.md Appendix A.10, the same as tests/test_gpu_parity3d.py): E, B <= 1e-8 after the
case's 4-6 steps (equal CG iteration counts keep it near 1e-12); np2 and cumcnt exact; per cell the same particle IDs, positions
and momenta to 1e-9.  Runs last (file name) so that the stage-wise tests localise a failure first."""
import numpy as np
import pytest

from tests.golden import make_ref_fixtures as mk
from tests.test_ref_golden import seeded_world
from tests.util import backend_for, canonical_cells, rel_err, upload_from_world

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    return mk.load()


@pytest.mark.parametrize("path", ["wm_step", "five-calls", "per-procedure"])
@pytest.mark.parametrize("name", list(mk.CASES))
def test_gpu_lands_on_the_reference_end_state(golden, name, path):
    dim, nx, ny, nz, n0, bc, order, u0, steps = mk.CASES[name]
    w, g = seeded_world(golden, name)
    b = backend_for(w)
    b.set_fused(path != "per-procedure")
    upload_from_world(b, w)
    ntot = int(g["np2_0"].sum())
    for it in range(steps):
        if path == "five-calls":
            b.time_loop(2, nx + 1, 1, order, u0)
        else:
            b.step(2, nx + 1, 1, order, u0)
        if bc == 0:
            res, rho = b.gauss()
            assert res < 1e-13 * max(rho, 1.0), f"Gauss residual {res} at step {it + 1}"
    st = b.stats()
    assert st["error_flags"] == 0 and st["n_particles"] == ntot
    up, np2, cc, uf = b.empty("up"), b.empty("np2"), b.empty("cumcnt"), b.empty("uf")
    b.download(up, np2, cc, uf)
    b.close()
    assert np.array_equal(np2, g["np2_1"]), "np2 differs from the reference"
    assert np.array_equal(cc, g["cumcnt_1"]), "cumcnt differs from the reference"
    assert rel_err(uf, g["uf_1"]) < 1e-8
    ref_up = mk.unpack(g["rec1"], g["np2_1"], w.np)
    worst = 0.0
    for (cg, rg), (cr, rr) in zip(canonical_cells(up, np2, cc), canonical_cells(ref_up, g["np2_1"], g["cumcnt_1"])):
        assert np.array_equal(cg, cr)
        assert np.array_equal(rg[:, -1].view(np.int64), rr[:, -1].view(np.int64))        # the same particles in every cell
        if len(rg):
            worst = max(worst, float(np.abs(rg[:, :-1] - rr[:, :-1]).max()))
    assert worst < 1e-9, worst
    w.close()
