"""GPU parity of the particle packing behind io__ptcl / io__orb (paraio's get_particle_count, 3d/common/paraio.f90:1007-1085):
mode 0 (all active particles) and mode 1 (tracers: ID > 0) against the oracle.  Counts per species exact; the records arrive
species-major and pencil by pencil in both; inside a pencil they are compared as ID-keyed sets (cell order inside a cell is
implementation-defined on both sides), bit-exact."""
import numpy as np
import pytest

from tests.util import active_mask, backend_for, make_world2, make_world3, upload_from_world

pytestmark = pytest.mark.gpu


def _mark_tracers(w, every=7):
    """flip the sign of every `every`-th ID: tracers are the particles with a positive ID (paraio.f90:1047-1049)"""
    up, np2 = w.arr("up"), w.arr("np2")
    m = active_mask(np2, w.np)
    ids = up[..., -1].view(np.int64)
    flip = m & (np.abs(ids) % every == 0)
    ids[flip] = -ids[flip]
    return int(flip.sum())


@pytest.mark.parametrize("dim", [3, 2])
def test_pack_matches_oracle(dim):
    w = make_world3(12, 6, 5, 6, steps=2) if dim == 3 else make_world2(16, 9, 6, steps=2)
    ntr = _mark_tracers(w)
    assert ntr > 0
    b = backend_for(w)
    upload_from_world(b, w)
    for _ in range(2):               # two more steps on both sides: the device state is in lazy-sorted form when packing
        w.step()
        b.step(2, w.nx + 1, 1)
    np2 = w.arr("np2")
    for mode in (0, 1):
        ref, lref = w.pack_particles(mode)
        got, lgot = b.pack_particles(mode)
        assert np.array_equal(lgot, lref), mode
        assert got.shape == ref.shape
        if mode == 1:
            assert (ref[:, -1].view(np.int64) > 0).all() and len(ref) == ntr
        else:
            assert len(ref) == int(np2.sum())
            # pencil by pencil, in the reference's (isp, k, j) order
            start = 0
            for n in np2.reshape(-1):
                a, r = got[start:start + n], ref[start:start + n]
                oa, orr = np.argsort(a[:, -1].view(np.int64)), np.argsort(r[:, -1].view(np.int64))
                assert np.array_equal(a[oa][:, -1].view(np.int64), r[orr][:, -1].view(np.int64))
                assert np.abs(a[oa][:, :-1] - r[orr][:, :-1]).max() < 1e-9 if n else True
                start += n
        # species blocks: IDs as sets
        s0 = int(lref[0])
        for lo, hi in ((0, s0), (s0, len(ref))):
            assert np.array_equal(np.sort(got[lo:hi, -1].view(np.int64)), np.sort(ref[lo:hi, -1].view(np.int64)))
    b.close(); w.close()


def test_pack_argument_errors():
    import wumingpic_b200 as wm
    w = make_world3(8, 4, 4, 3)
    b = backend_for(w)
    upload_from_world(b, w)
    with pytest.raises(wm.WmError):
        b.pack_particles(2)                      # "invalid mode specified for get_particle_count"
    b.particle__solv(2, 9)
    with pytest.raises(wm.WmError):
        b.pack_particles(0)                      # pushed set pending
    b.close(); w.close()
