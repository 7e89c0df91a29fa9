#!/usr/bin/env python
"""Restart re-slabbing tool: reads a WumingPIC restart snapshot in the reference's raw + JSON format, re-cuts it for another
rank count / rank grid, and writes it back in the same format -- so that a snapshot the Fortran code wrote on R ranks seeds a GPU run on any
number of GPUs (and the other way round), and so that reference-produced state can be fed to this backend and to the oracle.

Format (kept as is, nothing of it is re-designed here):
  3-D  paraio__output  3d/common/paraio.f90:102-283   reader paraio__input :288-515 (rank-count check :385-397)
       <name>.json = {"meta": {"endian", "rawfile"}, "attribute": {...}, "dataset": {...}}; <name>.raw = the bytes, in this order:
       attributes dummy_attribute, it, nxs, nxe, ndim, np, nxgs, nxge, nygs, nyge, nzgs, nzge, nsp, nproc (i4), delx, delt, c, r(nsp), q(nsp) (f8);
       datasets  np2(nyl, nzl, nsp, nproc) i4; up01, up02, ...(ndim, npg) f8: the ACTIVE particles of a species packed rank by rank, inside a
       rank pencil by pencil (k outer, j inner), get_particle_count mode 0, :1007-1037; poffset(nsp, nproc) i8 = element offset of each
       rank's block inside its species' dataset; uf(6, nxg, nyl+4, nzl+4, nproc) f8: every rank's field box INCLUDING ghost cells.
       Every JSON entry = {datatype, offset, size, ndim, shape, description[, data]} (utils/iocore/jsonio.f90:128-170); shapes are column-major.
  2-D  paraio__output  2d/common/paraio.f90: datasets up(ndim, np, nyl, nsp, nproc) f8 -- the whole PADDED particle array --, np2(nyl, nsp, nproc)
       i4, uf(6, nxg, nyl+4, nproc) f8; no nzgs / nzge attributes.
  rank = j * nproc_k + k (3d/common/mpi_set.f90:45-60); the format has one block size per dataset, i.e. even slabs (ny % nproc_j == 0).

    tools/restart_reslab.py IN_PREFIX OUT_PREFIX --nproc-j J [--nproc-k K]      (prefix = path without .json / .raw)

As a library: read_restart(), Snapshot.write(), Snapshot.slab(...) (host arrays of one rank in the backend's / reference's layout, cell-sorted
with their cumcnt: what `sort__bucket` makes of a freshly read restart, 3d/proj/weibel/app.f90:366-369) and Snapshot.from_slabs(...).
"""
import argparse
import json
import os
from collections import OrderedDict

import numpy as np

ATTR_ORDER_3D = ["dummy_attribute", "it", "nxs", "nxe", "ndim", "np", "nxgs", "nxge", "nygs", "nyge", "nzgs", "nzge", "nsp", "nproc",
                 "delx", "delt", "c", "r", "q"]
ATTR_ORDER_2D = [a for a in ATTR_ORDER_3D if a not in ("nzgs", "nzge")]


def _entry(datatype, offset, shape, desc, data=None):
    size = int(np.prod(shape)) * np.dtype(datatype).itemsize
    e = OrderedDict([("datatype", datatype), ("offset", int(offset)), ("size", int(size)), ("ndim", len(shape)),
                     ("shape", [int(s) for s in shape]), ("description", desc)])
    if data is not None:
        e["data"] = data
    return e, size


class Snapshot:
    """A restart snapshot in rank-independent form: particles per global pencil, the field on the global grid."""

    def __init__(self, dim, attrs, pencils, uf_interior, uf_xghost_ok=True):
        self.dim, self.attrs = dim, dict(attrs)
        self.pencils = pencils            # [isp][kk][jj] -> (n, ndim) float64 records (kk = 0 only in 2-D), file order kept
        self.uf = uf_interior             # (nz, ny, nx + 4, 6): all x (incl. the two x ghost layers on either side), interior y and z
        a = self.attrs
        self.ndim, self.np, self.nsp = int(a["ndim"]), int(a["np"]), int(a["nsp"])
        self.nx = int(a["nxge"]) - int(a["nxgs"]) + 1
        self.ny = int(a["nyge"]) - int(a["nygs"]) + 1
        self.nz = int(a["nzge"]) - int(a["nzgs"]) + 1 if dim == 3 else 1

    # ---- one rank's view ---------------------------------------------------------------------------------------------
    def field_box(self, j0, j1, k0, k1):
        """uf(6, nxgs-2:nxge+2, nys-2:nye+2[, nzs-2:nze+2]) of the rank owning global rows j0..j1-1 (0-based), planes k0..k1-1: ghost rows
        are the periodic images (what bc__dfield leaves there; the reference decomposes y and z periodically in every set-up)"""
        jj = (np.arange(j0 - 2, j1 + 2)) % self.ny
        if self.dim == 3:
            kk = (np.arange(k0 - 2, k1 + 2)) % self.nz
            return np.ascontiguousarray(self.uf[kk][:, jj])
        return np.ascontiguousarray(self.uf[0][jj])

    def slab(self, nys, nye, nzs=0, nze=0, sort=True):
        """host arrays of the rank nys..nye (, nzs..nze) in the reference layout (C order = reversed Fortran shape): up, np2, cumcnt, uf.
        sort=True orders every pencil by x cell (stable) and fills cumcnt -- sort__bucket after io__input; False keeps file order (cumcnt = 0)"""
        a = self.attrs
        j0, j1 = nys - int(a["nygs"]), nye - int(a["nygs"]) + 1
        k0, k1 = (nzs - int(a["nzgs"]), nze - int(a["nzgs"]) + 1) if self.dim == 3 else (0, 1)
        nyl, nzl, nx = j1 - j0, k1 - k0, self.nx
        up = np.zeros((self.nsp, nzl, nyl, self.np, self.ndim))
        np2 = np.zeros((self.nsp, nzl, nyl), dtype=np.int32)
        cumcnt = np.zeros((self.nsp, nzl, nyl, nx + 1), dtype=np.int32)
        nxgs = int(a["nxgs"])
        for isp in range(self.nsp):
            for k in range(k0, k1):
                for j in range(j0, j1):
                    rec = self.pencils[isp][k][j]
                    n = len(rec)
                    if n > self.np:
                        raise ValueError("memory over (np2 > np)")
                    if sort and n:
                        cell = rec[:, 0].astype(np.int64) - nxgs            # sort__bucket's key: int(x), delx = 1 (sort.f90:65)
                        o = np.argsort(cell, kind="stable")
                        rec = rec[o]
                        cnt = np.bincount(np.clip(cell, 0, nx - 1), minlength=nx)
                        cumcnt[isp, k - k0, j - j0, 1:] = np.cumsum(cnt)
                    up[isp, k - k0, j - j0, :n] = rec
                    np2[isp, k - k0, j - j0] = n
        uf = self.field_box(j0, j1, k0, k1)
        if self.dim == 2:
            return up[:, 0], np2[:, 0], cumcnt[:, 0], uf
        return up, np2, cumcnt, uf

    @classmethod
    def from_slabs(cls, dim, attrs, slabs):
        """slabs: list of (nys, nye, nzs, nze, up, np2, uf) of every rank (reference layout, as Snapshot.slab returns them)"""
        a = dict(attrs)
        nsp, ndim = int(a["nsp"]), int(a["ndim"])
        nx = int(a["nxge"]) - int(a["nxgs"]) + 1
        ny = int(a["nyge"]) - int(a["nygs"]) + 1
        nz = int(a["nzge"]) - int(a["nzgs"]) + 1 if dim == 3 else 1
        pencils = [[[None] * ny for _ in range(nz)] for _ in range(nsp)]
        uf = np.zeros((nz, ny, nx + 4, 6))
        for (nys, nye, nzs, nze, up, np2, ufl) in slabs:
            j0 = nys - int(a["nygs"])
            k0 = nzs - int(a["nzgs"]) if dim == 3 else 0
            if dim == 2:
                up, np2, ufl = up[:, None], np2[:, None], ufl[None]
                ufl_int = ufl[:, 2:-2]
            else:
                ufl_int = ufl[2:-2, 2:-2]
            nzl, nyl = np2.shape[1], np2.shape[2]
            uf[k0:k0 + nzl, j0:j0 + nyl] = ufl_int
            for isp in range(nsp):
                for k in range(nzl):
                    for j in range(nyl):
                        pencils[isp][k0 + k][j0 + j] = np.array(up[isp, k, j, :np2[isp, k, j]])
        return cls(dim, a, pencils, uf)

    # ---- writer: the reference's byte layout -------------------------------------------------------------------------------
    def write(self, prefix, nproc_j, nproc_k=1):
        dim, a = self.dim, dict(self.attrs)
        if self.ny % nproc_j or (dim == 3 and self.nz % nproc_k) or (dim == 2 and nproc_k != 1):
            raise ValueError("the restart format has one block size per dataset: ny (nz) must be a multiple of nproc_j (nproc_k)")
        nproc = nproc_j * nproc_k
        a["nproc"] = nproc
        nyl, nzl = self.ny // nproc_j, self.nz // nproc_k
        raw = bytearray()
        js = OrderedDict([("meta", OrderedDict([("endian", 1), ("rawfile", os.path.basename(prefix) + ".raw")])),
                          ("attribute", OrderedDict()), ("dataset", OrderedDict())])
        for name in (ATTR_ORDER_3D if dim == 3 else ATTR_ORDER_2D):
            v = a[name] if name != "dummy_attribute" else a.get("dummy_attribute", 8)
            if name in ("delx", "delt", "c", "r", "q"):
                arr = np.atleast_1d(np.asarray(v, dtype="<f8"))
                data = [float(x) for x in arr] if arr.size > 1 else float(arr[0])
                e, _ = _entry("f8", len(raw), [arr.size], "", data)
            else:
                arr = np.atleast_1d(np.asarray(v, dtype="<i4"))
                e, _ = _entry("i4", len(raw), [1], "", int(arr[0]))
            js["attribute"][name] = e
            raw += arr.tobytes()
        ranks = [(rj, rk) for rj in range(nproc_j) for rk in range(nproc_k)]     # rank = rj * nproc_k + rk
        nygs, nzgs = int(a["nygs"]), int(a.get("nzgs", 0))
        views = []
        for rj, rk in ranks:
            nys = nygs + rj * nyl
            nzs = nzgs + rk * nzl
            views.append(self.slab(nys, nys + nyl - 1, nzs, nzs + nzl - 1, sort=False))
        ds = js["dataset"]
        if dim == 3:
            # np2(nyl, nzl, nsp, nproc)
            e, _ = _entry("i4", len(raw), [nyl, nzl, self.nsp, nproc], "number of active particles")
            ds["np2"] = e
            for up, np2, _, _ in views:
                raw += np.ascontiguousarray(np2, dtype="<i4").tobytes()          # C order (nsp, nzl, nyl) == Fortran (nyl, nzl, nsp)
            poff = np.zeros((nproc, self.nsp), dtype="<i8")
            for isp in range(self.nsp):
                counts = [int(v[1][isp].sum()) for v in views]
                e, _ = _entry("f8", len(raw), [self.ndim, sum(counts)], "particle species #%02d" % (isp + 1))
                ds["up%02d" % (isp + 1)] = e
                off = 0
                for r, (up, np2, _, _) in enumerate(views):
                    poff[r, isp] = self.ndim * off
                    off += counts[r]
                    for k in range(nzl):                                          # pencil by pencil, k outer, j inner (paraio.f90:1025-1036)
                        for j in range(nyl):
                            raw += np.ascontiguousarray(up[isp, k, j, :np2[isp, k, j]], dtype="<f8").tobytes()
            e, _ = _entry("i8", len(raw), [self.nsp, nproc], "particle offset")
            ds["poffset"] = e
            raw += poff.tobytes()
            e, _ = _entry("f8", len(raw), [6, self.nx + 4, nyl + 4, nzl + 4, nproc], "electromagnetic fields including ghost cells")
            ds["uf"] = e
            for _, _, _, uf in views:
                raw += np.ascontiguousarray(uf, dtype="<f8").tobytes()
        else:
            e, _ = _entry("f8", len(raw), [self.ndim, self.np, nyl, self.nsp, nproc], "particles")
            ds["up"] = e
            for up, _, _, _ in views:
                raw += np.ascontiguousarray(up, dtype="<f8").tobytes()
            e, _ = _entry("i4", len(raw), [nyl, self.nsp, nproc], "number of active particles")
            ds["np2"] = e
            for _, np2, _, _ in views:
                raw += np.ascontiguousarray(np2, dtype="<i4").tobytes()
            e, _ = _entry("f8", len(raw), [6, self.nx + 4, nyl + 4, nproc], "electromagnetic fields including ghost cells")
            ds["uf"] = e
            for _, _, _, uf in views:
                raw += np.ascontiguousarray(uf, dtype="<f8").tobytes()
        with open(prefix + ".raw", "wb") as f:
            f.write(raw)
        with open(prefix + ".json", "w") as f:
            json.dump(js, f, indent=2)
        return nproc


def _read(rawf, e, bo="<"):
    rawf.seek(e["offset"])
    n = int(np.prod(e["shape"]))
    return np.fromfile(rawf, bo + e["datatype"], n)


def read_restart(prefix):
    with open(prefix + ".json") as f:
        js = json.load(f, object_pairs_hook=OrderedDict)
    endian = js["meta"]["endian"]
    if endian not in (1, 16777216):
        raise ValueError(f"unrecognized endian flag: {endian}")
    bo = "<" if endian == 1 else ">"
    rawpath = os.path.join(os.path.dirname(prefix) or ".", js["meta"]["rawfile"])
    attrs = OrderedDict()
    with open(rawpath, "rb") as rf:
        for name, e in js["attribute"].items():
            v = _read(rf, e, bo)
            attrs[name] = v.copy() if v.size > 1 else v[0].item()
        dim = 3 if "nzgs" in attrs else 2
        ds = js["dataset"]
        nsp, ndim, npc, nproc = int(attrs["nsp"]), int(attrs["ndim"]), int(attrs["np"]), int(attrs["nproc"])
        nx = int(attrs["nxge"]) - int(attrs["nxgs"]) + 1
        ny = int(attrs["nyge"]) - int(attrs["nygs"]) + 1
        nz = int(attrs["nzge"]) - int(attrs["nzgs"]) + 1 if dim == 3 else 1
        if ds["np2"]["shape"][-1] != nproc:                       # the reference's reader makes the same check (paraio.f90:385-397)
            raise ValueError("np2 dataset does not match the nproc attribute")
        if dim == 3:
            nyl, nzl = ds["np2"]["shape"][0], ds["np2"]["shape"][1]
            nproc_j, nproc_k = ny // nyl, nz // nzl
            if nproc_j * nproc_k != nproc:
                raise ValueError("rank grid cannot be derived from the np2 block shape")
            np2 = _read(rf, ds["np2"], bo).reshape(nproc, nsp, nzl, nyl)
            poff = _read(rf, ds["poffset"], bo).reshape(nproc, nsp)
            pencils = [[[None] * ny for _ in range(nz)] for _ in range(nsp)]
            for isp in range(nsp):
                data = _read(rf, ds["up%02d" % (isp + 1)], bo).reshape(-1, ndim)
                for r in range(nproc):
                    rj, rk = r // nproc_k, r % nproc_k
                    p = int(poff[r, isp]) // ndim
                    for k in range(nzl):
                        for j in range(nyl):
                            n = int(np2[r, isp, k, j])
                            pencils[isp][rk * nzl + k][rj * nyl + j] = data[p:p + n].astype(np.float64)
                            p += n
            ufr = _read(rf, ds["uf"], bo).reshape(nproc, nzl + 4, nyl + 4, nx + 4, 6)
            uf = np.zeros((nz, ny, nx + 4, 6))
            for r in range(nproc):
                rj, rk = r // nproc_k, r % nproc_k
                uf[rk * nzl:(rk + 1) * nzl, rj * nyl:(rj + 1) * nyl] = ufr[r, 2:-2, 2:-2]
        else:
            nyl = ds["np2"]["shape"][0]
            if nyl * nproc != ny:
                raise ValueError("rank count does not match the np2 block shape")
            up = _read(rf, ds["up"], bo).reshape(nproc, nsp, nyl, npc, ndim)
            np2 = _read(rf, ds["np2"], bo).reshape(nproc, nsp, nyl)
            pencils = [[[None] * ny] for _ in range(nsp)]
            for r in range(nproc):
                for isp in range(nsp):
                    for j in range(nyl):
                        pencils[isp][0][r * nyl + j] = up[r, isp, j, :np2[r, isp, j]].astype(np.float64)
            ufr = _read(rf, ds["uf"], bo).reshape(nproc, nyl + 4, nx + 4, 6)
            uf = np.zeros((1, ny, nx + 4, 6))
            for r in range(nproc):
                uf[0, r * nyl:(r + 1) * nyl] = ufr[r, 2:-2]
    return Snapshot(dim, attrs, pencils, uf)


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("src", help="input prefix (reads <src>.json + the raw file it names)")
    ap.add_argument("dst", help="output prefix (writes <dst>.json + <dst>.raw)")
    ap.add_argument("--nproc-j", type=int, required=True, help="ranks along y (2-D: the rank count)")
    ap.add_argument("--nproc-k", type=int, default=1, help="ranks along z (3-D)")
    args = ap.parse_args()
    snap = read_restart(args.src)
    n = snap.write(args.dst, args.nproc_j, args.nproc_k)
    tot = sum(len(p) for sp in snap.pencils for pl in sp for p in pl)
    print(f"{args.src} -> {args.dst}: {snap.dim}-D, {snap.nx} x {snap.ny}" + (f" x {snap.nz}" if snap.dim == 3 else "") +
          f" cells, {tot} particles, it = {snap.attrs['it']}, now {n} rank(s) ({args.nproc_j} x {args.nproc_k})")


if __name__ == "__main__":
    main()
