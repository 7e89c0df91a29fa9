#!/usr/bin/env python
"""build_ref.py -- the recipe behind oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

Translates the reference's hot-path Fortran sources, WHERE THEY LIE under /root/reference, into C++ with f2cxx.py and compiles
them with the distro g++ (`-O2 -ffp-contract=off -frounding-math`: no FMA contraction, like gfortran's default x86-64 code, and
no folding or motion of floating-point operations across the `ieee_set_rounding_mode` calls of the 2-D tree) into

    oracle/_ref/libwuming_ref3d.so   3d/common/{particle,field,sort,boundary_periodic,mom_calc}.f90
                                     + 3d/proj/{reconnection,shock}/boundary_*.f90
    oracle/_ref/libwuming_ref2d.so   the same files of the 2-D tree

Outputs (generated C++, objects, libraries) go to oracle/_ref/ only, which is git-ignored but travels to the GPU box.  Nothing is
translated when /root/reference is absent (the GPU box): the prebuilt libraries are used as they are.

Two more outputs serve the TIMED CPU baseline (`bench.py --impl reference`, `cpu_baseline.kind = "reference"`):

    oracle/_ref/libwuming_ref3d_fast.so   the same generated ref3d.cpp compiled `-O3 -march=native` (FMA contraction allowed, as
                                          `gfortran -O3 -march=native` would; never used for parity).  It is stamped with the
                                          host's CPU flags and recompiled from the generated C++ that travelled with it when the
                                          flags differ (the GPU box's host CPU is not this container's) -- g++ only, no reference needed
    oracle/_ref/libf2cxx_mpi.so           mpi_threads.cpp: the in-process MPI_SENDRECV / MPI_ALLREDUCE behind a flat-MPI run with
                                          one rank per host thread

    python oracle/f2cxx/build_ref.py [--bounds] [--fast] [--force]
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


def lib_path(dim, bounds=False, fast=False):
    return os.path.join(OUT, f"libwuming_ref{dim}d{'_chk' if bounds else ''}{'_fast' if fast else ''}.so")


def host_cpu_tag():
    """sha1 of this host's sorted CPU feature flags (what a -march=native build depends on)"""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return hashlib.sha1(" ".join(sorted(line.split(":", 1)[1].split())).encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


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


def _compile(cpp, lib, flags):
    cmd = [CXX, "-std=c++17"] + flags + ["-fPIC", "-shared", "-I", HERE, "-o", lib, cpp, os.path.join(HERE, "f90rt.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed on the translated reference:\n" + r.stderr[-4000:])


FAST_FLAGS = ["-O3", "-march=native", "-fno-math-errno"]


def build_fast(dim, force=False, quiet=True):
    """-> path of libwuming_ref{dim}d_fast.so (the timed CPU baseline's build of the SAME generated C++), or None.  Rebuilt when the
    generated C++ is newer or the host's CPU flags differ from the ones it was compiled for; needs only g++ and oracle/_ref/ref{dim}d.cpp"""
    if build(dim) is None:
        return None
    cpp, lib = os.path.join(OUT, f"ref{dim}d.cpp"), lib_path(dim, fast=True)
    if not os.path.exists(cpp):
        return None
    stamp_file = lib + ".stamp"
    h = hashlib.sha1(open(cpp, "rb").read())
    for f in ("f90rt.h", "f90rt.cpp"):
        h.update(open(os.path.join(HERE, f), "rb").read())
    # the 2-D tree changes the rounding mode (ieee_down sections of boundary_periodic.f90): no folding / motion across those calls
    flags = FAST_FLAGS + (["-frounding-math"] if dim == 2 else [])
    stamp = h.hexdigest() + " " + host_cpu_tag() + " " + " ".join(flags)
    if not force and os.path.exists(lib) and os.path.exists(stamp_file) and open(stamp_file).read().strip() == stamp:
        return lib
    _compile(cpp, lib, flags)
    with open(stamp_file, "w") as f:
        f.write(stamp)
    if not quiet:
        print(f"built {lib} ({' '.join(FAST_FLAGS)}) from {cpp}")
    return lib


def build_mpi(force=False):
    """-> path of libf2cxx_mpi.so (mpi_threads.cpp: no reference source involved, g++ only)"""
    os.makedirs(OUT, exist_ok=True)
    src, lib = os.path.join(HERE, "mpi_threads.cpp"), os.path.join(OUT, "libf2cxx_mpi.so")
    if force or not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
        r = subprocess.run([CXX, "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", "-o", lib, src], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed on mpi_threads.cpp:\n" + r.stderr[-4000:])
    return lib


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
    _compile(cpp, lib, ["-O2", "-ffp-contract=off", "-frounding-math"] + (["-DF90_BOUNDS"] if bounds else []))
    with open(stamp_file, "w") as f:
        f.write(stamp)
    if not quiet:
        print(f"built {lib} from {len(SOURCES[f'{dim}d'])} reference files ({len(text.splitlines())} lines of C++)")
    return lib


if __name__ == "__main__":
    b, force = "--bounds" in sys.argv, "--force" in sys.argv
    build_mpi(force)
    if "--fast" in sys.argv:
        print(build_fast(3, force=force, quiet=False))
        sys.exit(0)
    if not reference_present():
        print(f"{REF} is absent: nothing to translate (prebuilt libraries in {OUT} are used as they are)")
        sys.exit(0)
    for d in (3, 2):
        build(d, bounds=b, force=force, quiet=False)
