#!/usr/bin/env python
"""make_reference_patch.py -- the build-system side of the drop-in, as a patch a maintainer applies to a WumingPIC checkout.

    python tools/make_reference_patch.py --ref /path/to/WumingPIC [--out wuming_b200.patch] [--install] [--resident]

Reads the checkout's Makefiles and drivers WHERE THEY LIE, derives the edited versions mechanically and writes a unified diff
(nothing of the reference is stored in this repository; the patch is made from the maintainer's own tree, so it follows whatever
version they have).  With --install it also copies fortran/wuming_b200_c.f90 and wuming_b200_shim{2,3}d.f90 into
{2d,3d}/common/ -- the only new files.  The edits:

  common.mk                      WM_B200 (this repository) and WM_B200_LIBS (-L... -lwuming_b200 -Wl,-rpath,...)
  {2d,3d}/common/Makefile        compile wuming_b200_c.f90 + wuming_b200_shim{2,3}d.f90 INSTEAD of boundary_periodic / field /
                                 particle / mom_calc / sort .f90 into libwuming{2,3}d_common.a; the module files by name (one
                                 source file now holds several modules); mpi_set, paraio, fio, h5io, wuming{2,3}d stay
  {2d,3d}/proj/*/Makefile        $(WM_B200_LIBS) behind the static libraries on the link line; the reconnection / shock projects stop
                                 compiling their own boundary_*.f90 (the shim provides those modules)
  {2d,3d}/proj/*/app.f90         one line after mpi_set__init: call wm_shim_comm_init(nproc, nproc_j, nproc_k, nrank, ncomw)
                                 (+ its `use`): the rank grid and the communicator for the device side

  --resident (optional)          the fast mode: the state stays on the GPU between time steps (WM_SHIM_RESIDENT) and the five calls of the
                                 loop run the fused kernel + lazy sort.  In app__main: `call wm_shim_set_mode(WM_SHIM_RESIDENT)` after
                                 init(), `call wm_shim_sync_to_host(up,uf,np2,cumcnt)` in front of everything that reads the host arrays
                                 (io__ptcl, io__orb, the moment / energy block, save_restart -- one download per step however many of
                                 them fire), and in the shock drivers the same in front of inject() / relocate() with
                                 `call wm_shim_host_modified()` behind them (they edit the host arrays).  Without it the default
                                 WM_SHIM_SYNC_EVERY_CALL mode needs no further edit and hands every result back after every call.

Everything else -- main.f90, the time loops, JSON config, paraio / mpiio output, utils -- is untouched: main.out runs as before,
with the kernels behind the five calls of the time loop on the GPU.  tests/test_reference_patch.py applies the patch to a copy of
the reference tree and checks the result.
"""
import argparse
import difflib
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REPLACED = ("boundary_periodic.f90", "field.f90", "particle.f90", "mom_calc.f90", "sort.f90")
SHIM_MODULES = ("wuming_b200_c", "particle", "field", "sort", "boundary_periodic", "boundary_reconnection", "boundary_shock",
                "mom_calc", "shock_source")


def edit_common_mk(text):
    add = ("\n# wuming-b200: the CUDA backend behind the time loop (libwuming_b200.so, C ABI include/wuming_b200.h)\n"
           f"WM_B200      ?= {ROOT}\n"
           "WM_B200_LIBS  = -L$(WM_B200)/wumingpic_b200/lib -lwuming_b200 -Wl,-rpath,$(WM_B200)/wumingpic_b200/lib\n")
    return text.rstrip("\n") + "\n" + add


def edit_common_makefile(text, dim):
    # the SRCS continuation block: drop the five replaced files, add the two shim files in front
    m = re.search(r"^SRCS\s*=\s*\\\n((?:.*\\\n)*.*\n)", text, re.M)
    if not m:
        raise SystemExit(f"{dim}d/common/Makefile: SRCS block not found")
    files = [f for f in m.group(1).replace("\\", " ").split() if f not in REPLACED]
    kept_mods = [f[:-4] for f in files]
    files = files[:1] + ["wuming_b200_c.f90", f"wuming_b200_shim{dim}d.f90"] + files[1:]
    text = text[:m.start()] + "SRCS   = \\\n\t" + " ".join(files) + "\n" + text[m.end():]
    mods = " ".join(f"{n}.mod" for n in kept_mods[:1] + list(SHIM_MODULES) + kept_mods[1:])
    text, n = re.subn(r"^MODS\s*=.*$", "MODS   = " + mods, text, count=1, flags=re.M)
    if n != 1:
        raise SystemExit(f"{dim}d/common/Makefile: MODS line not found")
    # dependencies: the umbrella module needs the shim's modules; the shim needs the C-binding module
    text = re.sub(rf"^(wuming{dim}d\.o:)\s*field\.o particle\.o mom_calc\.o sort\.o\s*\\\n", rf"\1 wuming_b200_shim{dim}d.o \\\n", text, flags=re.M)
    return text.rstrip("\n") + f"\nwuming_b200_shim{dim}d.o: wuming_b200_c.o\n"


def edit_proj_makefile(text):
    text, n = re.subn(r"^([ \t]*\$\(FC\) -o \$@ \$\^ \$\(LDFLAGS\) -lwuming[23]d_common -lwuming_utils)(.*)$", r"\1\2 $(WM_B200_LIBS)", text,
                      count=1, flags=re.M)
    if n != 1:
        raise SystemExit("proj Makefile: link line not found")
    text = re.sub(r"^(SRCS[ \t]*=[ \t]*main\.f90 app\.f90) boundary_(reconnection|shock)\.f90[ \t]*$", r"\1", text, flags=re.M)
    text = re.sub(r"^(app\.o:) boundary_(reconnection|shock)\.o ", r"\1 ", text, flags=re.M)
    return text


USE_PLAIN = "use wuming_b200_c, only: wm_shim_comm_init"
USE_RESIDENT = ("use wuming_b200_c, only: wm_shim_comm_init, wm_shim_set_mode, wm_shim_sync_to_host, wm_shim_host_modified, &\n"
                "{ind}                         WM_SHIM_RESIDENT")
SYNC = "call wm_shim_sync_to_host(up,uf,np2,cumcnt)"
READERS = ("io__ptcl", "io__orb", "mom_calc__accl", "save_restart")       # mom_calc__accl opens the moment / io__mom / energy block
WRITERS = ("inject", "relocate")                                           # shock drivers: edit up, np2, cumcnt, uf, nxe on the host


def edit_app(text, dim, resident=False):
    def use_line(m):
        extra = USE_RESIDENT.format(ind=m.group(1)) if resident else USE_PLAIN
        return f"{m.group(0)}\n{m.group(1)}{extra}"
    text, n = re.subn(rf"^([ \t]*)use wuming{dim}d[ \t]*$", use_line, text, count=1, flags=re.M | re.I)
    if n != 1:
        raise SystemExit("app.f90: `use wuming?d` not found")
    grid = "nproc,nproc_j,nproc_k" if dim == 3 else "nproc,nproc,1"
    text, n = re.subn(r"^([ \t]*)call mpi_set__init\(.*\)[ \t]*$", rf"\g<0>\n\1call wm_shim_comm_init({grid},nrank,ncomw)", text, count=1,
                      flags=re.M | re.I)
    if n != 1:
        raise SystemExit("app.f90: call mpi_set__init not found")
    if not resident:
        return text
    # the remaining edits live inside app__main
    m = re.search(r"^[ \t]*subroutine app__main\b.*?^[ \t]*end subroutine app__main\b", text, re.M | re.S | re.I)
    if not m:
        raise SystemExit("app.f90: app__main not found")
    body = m.group(0)
    body, n = re.subn(r"^([ \t]*)call init\(\)[ \t]*$", r"\g<0>\n\1call wm_shim_set_mode(WM_SHIM_RESIDENT)", body, count=1, flags=re.M | re.I)
    if n != 1:
        raise SystemExit("app.f90: call init() not found in app__main")
    for name in READERS:
        body = re.sub(rf"^([ \t]*)(call {name}\()", rf"\1{SYNC}\n\1\2", body, flags=re.M | re.I)
    for name in WRITERS:
        body = re.sub(rf"^([ \t]*)(call {name}\(\))[ \t]*$", rf"\1{SYNC}\n\1\2\n\1call wm_shim_host_modified()", body, flags=re.M | re.I)
    return text[:m.start()] + body + text[m.end():]


def plan(ref, resident=False):
    """[(relative path, new text)]"""
    out = [("common.mk", edit_common_mk(open(os.path.join(ref, "common.mk")).read()))]
    for dim in (2, 3):
        rel = f"{dim}d/common/Makefile"
        out.append((rel, edit_common_makefile(open(os.path.join(ref, rel)).read(), dim)))
        proj = os.path.join(ref, f"{dim}d", "proj")
        for name in sorted(os.listdir(proj)):
            mk, app = os.path.join(proj, name, "Makefile"), os.path.join(proj, name, "app.f90")
            if os.path.exists(mk) and os.path.exists(app):
                out.append((f"{dim}d/proj/{name}/Makefile", edit_proj_makefile(open(mk).read())))
                out.append((f"{dim}d/proj/{name}/app.f90", edit_app(open(app).read(), dim, resident)))
    return out


def make_patch(ref, resident=False):
    chunks = []
    for rel, new in plan(ref, resident):
        old = open(os.path.join(ref, rel)).read()
        chunks += list(difflib.unified_diff(old.splitlines(True), new.splitlines(True), "a/" + rel, "b/" + rel, n=2))
    return "".join(chunks)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", required=True, help="a WumingPIC checkout")
    ap.add_argument("--out", default="wuming_b200.patch")
    ap.add_argument("--install", action="store_true", help="also copy the shim sources into {2d,3d}/common/ of the checkout")
    ap.add_argument("--resident", action="store_true", help="device-resident time loop: mode switch + sync points in app__main")
    a = ap.parse_args()
    text = make_patch(a.ref, a.resident)
    with open(a.out, "w") as f:
        f.write(text)
    print(f"wrote {a.out}: {text.count(chr(10) + '+++ ')+ (1 if text.startswith('--- ') else 0)} files; apply with  patch -p1 -d {a.ref} < {a.out}")
    if a.install:
        for dim in (2, 3):
            for src in ("wuming_b200_c.f90", f"wuming_b200_shim{dim}d.f90"):
                shutil.copy(os.path.join(ROOT, "fortran", src), os.path.join(a.ref, f"{dim}d", "common", src))
                print("copied", src, f"-> {dim}d/common/")


if __name__ == "__main__":
    sys.exit(main())
