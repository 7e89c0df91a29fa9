"""tools/restart_reslab.py: the reference's restart format (3d/common/paraio.f90:102-283, 2d/common/paraio.f90) written and
read back, re-cut between rank grids, and handed to the kernels.  State comes from the oracle's multi-rank emulation."""
import importlib.util
import json
import os
import sys

import numpy as np
import pytest

from tests.util import canonical_cells, make_world2, make_world3

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("restart_reslab", os.path.join(ROOT, "tools", "restart_reslab.py"))
rr = importlib.util.module_from_spec(spec)
spec.loader.exec_module(rr)

NX, NY, NZ, N0 = 10, 8, 8, 4


def _attrs3(w, nproc, it=7):
    return dict(dummy_attribute=8, it=it, nxs=2, nxe=w.nx + 1, ndim=7, np=w.np, nxgs=2, nxge=w.nx + 1, nygs=2, nyge=w.ny + 1, nzgs=2,
                nzge=getattr(w, 'nz', 1) + 1, nsp=2, nproc=nproc, delx=w.delx, delt=w.delt, c=w.c, r=w.r, q=w.q)


def _snapshot_of(w, dim=3):
    slabs = []
    for rk in range(w.nranks):
        g = w.geom(rk)
        slabs.append((g["nys"], g["nye"], g["nzs"], g["nze"], w.arr("up", rk), w.arr("np2", rk), w.arr("uf", rk)))
    a = _attrs3(w, w.nranks)
    if dim == 2:
        a = {k: v for k, v in a.items() if k not in ("nzgs", "nzge")}
        a.update(ndim=6)
    return rr.Snapshot.from_slabs(dim, a, slabs)


def test_3d_roundtrip_and_reslab(tmp_path):
    w = make_world3(NX, NY, NZ, N0, steps=2, nproc_j=2, nproc_k=2)
    snap = _snapshot_of(w)
    p22, p14, p11, p22b = (str(tmp_path / n) for n in ("r22", "r14", "r11", "r22b"))
    assert snap.write(p22, 2, 2) == 4
    rr.read_restart(p22).write(p14, 1, 4)
    rr.read_restart(p14).write(p11, 1, 1)
    rr.read_restart(p11).write(p22b, 2, 2)
    assert open(p22 + ".raw", "rb").read() == open(p22b + ".raw", "rb").read()         # 2x2 -> 1x4 -> 1x1 -> 2x2 is the identity, bit for bit
    j1, j2 = json.load(open(p22 + ".json")), json.load(open(p22b + ".json"))
    j2["meta"]["rawfile"] = j1["meta"]["rawfile"]
    assert j1 == j2
    # the 1x4 file holds what a 1x4 run of the same physical state holds: the same particles in every pencil, the same field
    s14 = rr.read_restart(p14)
    assert s14.attrs["nproc"] == 4 and s14.attrs["it"] == 7
    w14 = make_world3(NX, NY, NZ, N0, steps=2, nproc_j=1, nproc_k=4)
    for rk in range(4):
        g = w14.geom(rk)
        up, np2, cc, uf = s14.slab(g["nys"], g["nye"], g["nzs"], g["nze"])
        assert np.array_equal(np2, w14.arr("np2", rk)) and np.array_equal(cc, w14.arr("cumcnt", rk))
        assert np.abs(uf - w14.arr("uf", rk)).max() <= 1e-12 * np.abs(uf).max()        # incl. the periodic ghost rows
        for (c1, r1), (c2, r2) in zip(canonical_cells(up, np2, cc), canonical_cells(w14.arr("up", rk), np2, cc)):
            assert np.array_equal(c1, c2) and np.array_equal(r1[:, -1].view(np.int64), r2[:, -1].view(np.int64))
            assert len(r1) == 0 or np.abs(r1[:, :-1] - r2[:, :-1]).max() < 1e-12


def test_json_follows_the_reference_schema(tmp_path):
    w = make_world3(NX, NY, NZ, N0, steps=1, nproc_j=2, nproc_k=1)
    p = str(tmp_path / "snap")
    _snapshot_of(w).write(p, 2, 1)
    js = json.load(open(p + ".json"))
    assert list(js) == ["meta", "attribute", "dataset"] and js["meta"]["endian"] == 1 and js["meta"]["rawfile"] == "snap.raw"
    assert list(js["attribute"]) == rr.ATTR_ORDER_3D
    assert list(js["dataset"]) == ["np2", "up01", "up02", "poffset", "uf"]
    raw = open(p + ".raw", "rb").read()
    off = 0
    for name, e in list(js["attribute"].items()) + list(js["dataset"].items()):
        assert list(e)[:6] == ["datatype", "offset", "size", "ndim", "shape", "description"]
        assert e["offset"] == off and e["ndim"] == len(e["shape"])                     # back to back, in the reference's order
        assert e["size"] == int(np.prod(e["shape"])) * np.dtype(e["datatype"]).itemsize
        off += e["size"]
    assert off == len(raw)
    assert js["dataset"]["np2"]["shape"] == [NY // 2, NZ, 2, 2] and js["dataset"]["uf"]["shape"] == [6, NX + 4, NY // 2 + 4, NZ + 4, 2]
    # attribute values in the JSON equal the bytes (what python/jsoncheck.py of the reference verifies)
    for name, e in js["attribute"].items():
        v = np.frombuffer(raw, "<" + e["datatype"], int(np.prod(e["shape"])), e["offset"])
        assert np.array_equal(v, np.atleast_1d(e["data"])), name
    ref_tool = "/root/reference/python"
    if os.path.isdir(ref_tool):                                                        # the reference's own checker, where the tree exists
        sys.path.insert(0, ref_tool)
        try:
            import jsoncheck
            assert jsoncheck.check_attribute(p + ".json", verbose=0)
        finally:
            sys.path.remove(ref_tool)


def test_slab_sort_is_sort_bucket(tmp_path):
    """a freshly read restart is unsorted (paraio__input + sort__bucket, 3d/proj/weibel/app.f90:366-369): Snapshot.slab(sort=True) must give
    exactly what the oracle's sort__bucket gives for the same records"""
    w = make_world3(NX, NY, NZ, N0, steps=2)
    w.particle_solv(); w.field_fdtd_i(); w.bc_particle_x(); w.bc_particle_yz()       # gp: pushed, re-binned, NOT yet sorted in x
    gp, np2 = w.arr("gp").copy(), w.arr("np2").copy()
    snap = rr.Snapshot.from_slabs(3, _attrs3(w, 1), [(2, NY + 1, 2, NZ + 1, gp, np2, w.arr("uf"))])
    up, np2s, cc, _ = snap.slab(2, NY + 1, 2, NZ + 1)
    w.sort_bucket()
    assert np.array_equal(np2s, w.arr("np2")) and np.array_equal(cc, w.arr("cumcnt"))
    m = np.arange(w.np)[None, None, None, :] < np2s[..., None]
    assert np.array_equal(up[m].view(np.int64), w.arr("up")[m].view(np.int64))        # same stable order, bit for bit


def test_2d_roundtrip(tmp_path):
    w = make_world2(12, 8, 5, steps=2, nproc=4)
    snap = _snapshot_of(w, dim=2)
    a, b, c = (str(tmp_path / n) for n in ("a4", "b2", "c4"))
    snap.write(a, 4)
    rr.read_restart(a).write(b, 2)
    rr.read_restart(b).write(c, 4)
    assert open(a + ".raw", "rb").read() == open(c + ".raw", "rb").read()
    js = json.load(open(b + ".json"))
    assert list(js["dataset"]) == ["up", "np2", "uf"] and js["dataset"]["up"]["shape"] == [6, w.np, 4, 2, 2]
    assert list(js["attribute"]) == rr.ATTR_ORDER_2D


@pytest.mark.gpu
def test_restart_seeds_a_gpu_run(tmp_path):
    """a 2x2-rank snapshot, re-cut to one rank, uploaded and stepped on the GPU = the oracle's one-rank run from the same state"""
    from tests.util import backend_for, rel_err
    w4 = make_world3(NX, NY, NZ, N0, steps=2, nproc_j=2, nproc_k=2)
    p = str(tmp_path / "r22")
    _snapshot_of(w4).write(p, 2, 2)
    snap = rr.read_restart(p)
    up, np2, cc, uf = snap.slab(2, NY + 1, 2, NZ + 1)
    w1 = make_world3(NX, NY, NZ, N0, steps=2)
    b = backend_for(w1)
    b.upload(up, np2, cc, uf)
    b.upload_work("df", w1.arr("df"))
    for _ in range(3):
        w1.step()
        b.step(2, NX + 1, 1)
    uf1, np21 = b.empty("uf"), b.empty("np2")
    b.download(uf=uf1, np2=np21)
    assert np.array_equal(np21, w1.arr("np2")) and rel_err(uf1, w1.arr("uf")) < 1e-9
    b.close()
