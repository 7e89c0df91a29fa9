#!/usr/bin/env python
"""build_ref.py -- the recipe behind oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

Translates the reference's hot-path Fortran sources, WHERE THEY LIE under /root/reference, into C++ with f2cxx.py and compiles
them with the distro g++ (`-O2 -ffp-contract=off -frounding-math`: no FMA contraction, like gfortran's default x86-64 code, and
no folding or motion of floating-point operations across the `ieee_set_rounding_mode` calls of the 2-D tree) into

    oracle/_ref/libwuming_ref3d.so   3d/common/{particle,field,sort,boundary_periodic,mom_calc}.f90
                                     + 3d/proj/{reconnection,shock}/boundary_*.f90
    oracle/_ref/libwuming_ref2d.so   the same files of the 2-D tree

Outputs (generated C++, objects, libraries) go to oracle/_ref/ only, which is git-ignored but travels to the GPU box.  Nothing is
rebuilt when /root/reference is absent (the GPU box): the prebuilt libraries are used as they are.

    python oracle/f2cxx/build_ref.py [--bounds] [--force]
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref")
REF = os.environ.get("WUMING_REFERENCE", "/root/reference")
CXX = "/usr/bin/g++"       # the image's CXX=/opt/gcc/bin/g++ lacks libgomp; one compiler for everything under oracle/

SOURCES = {
    "3d": ["3d/common/particle.f90", "3d/common/field.f90", "3d/common/sort.f90", "3d/common/boundary_periodic.f90",
           "3d/common/mom_calc.f90", "3d/proj/reconnection/boundary_reconnection.f90", "3d/proj/shock/boundary_shock.f90"],
    "2d": ["2d/common/particle.f90", "2d/common/field.f90", "2d/common/sort.f90", "2d/common/boundary_periodic.f90",
           "2d/common/mom_calc.f90", "2d/proj/reconnection/boundary_reconnection.f90", "2d/proj/shock/boundary_shock.f90"],
}


def lib_path(dim, bounds=False):
    return os.path.join(OUT, f"libwuming_ref{dim}d{'_chk' if bounds else ''}.so")


def reference_present():
    return all(os.path.exists(os.path.join(REF, f)) for fs in SOURCES.values() for f in fs)


def _stamp(dim, bounds):
    h = hashlib.sha1()
    for f in SOURCES[f"{dim}d"]:
        h.update(open(os.path.join(REF, f), "rb").read())
    for f in ("f2cxx.py", "f90rt.h", "f90rt.cpp", "build_ref.py"):
        h.update(open(os.path.join(HERE, f), "rb").read())
    h.update(b"bounds" if bounds else b"plain")
    return h.hexdigest()


def build(dim, bounds=False, force=False, quiet=True):
    """-> path of the library, or None when neither the reference sources nor a prebuilt library exist"""
    lib = lib_path(dim, bounds)
    if not reference_present():
        return lib if os.path.exists(lib) else None
    os.makedirs(OUT, exist_ok=True)
    stamp_file = lib + ".stamp"
    stamp = _stamp(dim, bounds)
    if not force and os.path.exists(lib) and os.path.exists(stamp_file) and open(stamp_file).read().strip() == stamp:
        return lib
    sys.path.insert(0, HERE)
    import f2cxx
    cpp = os.path.join(OUT, f"ref{dim}d.cpp")
    text = f2cxx.translate([(f, open(os.path.join(REF, f)).read()) for f in SOURCES[f"{dim}d"]])
    with open(cpp, "w") as f:
        f.write(text)
    cmd = [CXX, "-std=c++17", "-O2", "-ffp-contract=off", "-frounding-math", "-fPIC", "-shared", "-I", HERE, "-o", lib, cpp,
           os.path.join(HERE, "f90rt.cpp")] + (["-DF90_BOUNDS"] if bounds else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed on the translated reference:\n" + r.stderr[-4000:])
    with open(stamp_file, "w") as f:
        f.write(stamp)
    if not quiet:
        print(f"built {lib} from {len(SOURCES[f'{dim}d'])} reference files ({len(text.splitlines())} lines of C++)")
    return lib


if __name__ == "__main__":
    b, force = "--bounds" in sys.argv, "--force" in sys.argv
    if not reference_present():
        print(f"{REF} is absent: nothing to translate (prebuilt libraries in {OUT} are used as they are)")
        sys.exit(0)
    for d in (3, 2):
        build(d, bounds=b, force=force, quiet=False)
