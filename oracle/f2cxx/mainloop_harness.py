#!/usr/bin/env python
"""mainloop_harness.py -- the reference drivers' OWN main loop, patched by tools/make_reference_patch.py --resident, executed on top
of the ISO_C_BINDING shim (TEST INFRASTRUCTURE ONLY).

`app__main` of every proj/*/app.f90 is: load_config, init, then the time loop -- the five library calls per step, the shock
driver's inject() / relocate(), and the output block (io__ptcl, io__orb, the moments, energy_history, save_restart) at their
cadences.  load_config / init / the I/O modules need JSON, MPI-IO and character handling, which the translator does not take; the
loop itself does not.  So this recipe takes the text of app__main from the line after `call init()` to its end, VERBATIM FROM THE
PATCHED DRIVER (the reference file where it lies -> make_reference_patch.edit_app(resident=True) -> here), together with the driver's
own `use boundary_*, bc__... => ...` statement, and wraps it into a module `app` whose declarations and output procedures are
written here: io__ptcl / io__orb / io__mom / energy_history / save_restart record WHAT THEY WERE HANDED (step number and checksums
of the host arrays they would write), inject() / relocate() make a small deterministic edit of the host arrays.  The module is
translated together with fortran/wuming_b200_c.f90 + wuming_b200_shim{2,3}d.f90 into

    oracle/_ref/libwuming_main_<setup><dim>d.so        (undefined wm_* symbols, like the shim library)

and run over the recording stub with the oracle as the device (tests/test_shim_executed.py) and over libwuming_b200.so on a GPU
(tests/gpu_shim_cases.py): every output a patched driver writes must be the state of that step, although nothing but those
outputs ever leaves the device.
"""
import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(os.path.dirname(HERE), "_ref")
REF = os.environ.get("WUMING_REFERENCE", "/root/reference")
CXX = "/usr/bin/g++"
MAGIC = 8.88e88            # record marker in the capture stream: MAGIC, kind, it, n, v1 .. vn
KINDS = {1: "io__ptcl", 2: "io__orb", 3: "io__mom", 4: "energy_history", 5: "save_restart", 6: "inject", 7: "relocate"}


def _decl(dim, shock):
    a2 = ("np2(:,:), cumcnt(:,:,:)", "uf(:,:,:), up(:,:,:,:), gp(:,:,:,:), mom(:,:,:,:)") if dim == 2 else \
        ("np2(:,:,:), cumcnt(:,:,:,:)", "uf(:,:,:,:), up(:,:,:,:,:), gp(:,:,:,:,:), mom(:,:,:,:,:)")
    return f"""
  integer :: ndim, np, nsp, nxgs, nxge, nygs, nyge, nzgs, nzge, nys, nye, nzs, nze
  integer :: nxs, nxe, it0, max_it, intvl_ptcl, intvl_orb, intvl_mom, intvl_expand, verbose, nrank
  integer :: restart_file = 0, hunit = 10          ! the driver's character variable / a unit: `write` to either is captured
  integer, parameter :: nroot = 0
  real(8) :: max_elapsed, u0
  integer, allocatable :: {a2[0]}
  real(8), allocatable :: {a2[1]}
"""


def _pencil_loops(dim, body):
    if dim == 3:
        return f"""    do isp = 1, nsp
    do k = nzs, nze
    do j = nys, nye
{body.replace('@P', 'j,k,isp')}
    enddo
    enddo
    enddo
"""
    return f"""    do isp = 1, nsp
    do j = nys, nye
{body.replace('@P', 'j,isp')}
    enddo
    enddo
"""


def _standins(dim):
    """harness-written: what the output procedures were handed, as checksums of the MODULE arrays (their actual arguments)"""
    jk = "j, k" if dim == 3 else "j"
    alloc = ("allocate(np2(nys:nye,nzs:nze,nsp)); allocate(cumcnt(nxgs:nxge+1,nys:nye,nzs:nze,nsp))\n"
             "    allocate(uf(6,nxgs-2:nxge+2,nys-2:nye+2,nzs-2:nze+2)); allocate(up(ndim,np,nys:nye,nzs:nze,nsp))\n"
             "    allocate(gp(ndim,np,nys:nye,nzs:nze,nsp)); allocate(mom(7,nxgs-1:nxge+1,nys-1:nye+1,nzs-1:nze+1,nsp))") if dim == 3 else \
            ("allocate(np2(nys:nye,nsp)); allocate(cumcnt(nxgs:nxge+1,nys:nye,nsp))\n"
             "    allocate(uf(6,nxgs-2:nxge+2,nys-2:nye+2)); allocate(up(ndim,np,nys:nye,nsp))\n"
             "    allocate(gp(ndim,np,nys:nye,nsp)); allocate(mom(7,nxgs-1:nxge+1,nys-1:nye+1,nsp))")
    psum = _pencil_loops(dim, "      do ii = 1, np2(@P)\n        s = s + up(1,ii,@P) + 3d0*up(ndim-1,ii,@P)\n      enddo")
    # the edit must not depend on the ORDER of the particles inside a cell (the device's order is deterministic but not the oracle's)
    edit = _pencil_loops(dim, "      do ii = 1, np2(@P)\n        up(ndim-1,ii,@P) = 0.999d0*up(ndim-1,ii,@P)\n      enddo")
    momin = ":,nxgs:nxge,nys:nye,nzs:nze,:" if dim == 3 else ":,nxgs:nxge,nys:nye,:"
    return f"""
  subroutine harness__alloc(ndim_in,np_in,nsp_in,nxgs_in,nxge_in,nygs_in,nyge_in,nzgs_in,nzge_in,nys_in,nye_in,nzs_in,nze_in)
    integer, intent(in) :: ndim_in,np_in,nsp_in,nxgs_in,nxge_in,nygs_in,nyge_in,nzgs_in,nzge_in,nys_in,nye_in,nzs_in,nze_in
    ndim = ndim_in; np = np_in; nsp = nsp_in; nxgs = nxgs_in; nxge = nxge_in; nygs = nygs_in; nyge = nyge_in
    nzgs = nzgs_in; nzge = nzge_in; nys = nys_in; nye = nye_in; nzs = nzs_in; nze = nze_in
    {alloc}
    np2 = 0; cumcnt = 0; uf = 0d0; up = 0d0; gp = 0d0; mom = 0d0
  end subroutine harness__alloc

  function get_etime() result(t)
    real(8) :: t
    t = 0d0
  end function get_etime

  function particle_checksum() result(s)
    real(8) :: s
    integer :: isp, {jk}, ii
    s = 0d0
{psum}  end function particle_checksum

  subroutine io__ptcl(a, b, c, it)
    real(8), intent(in) :: a(*), b(*)
    integer, intent(in) :: c(*), it
    write(hunit,*) {MAGIC:.3e}_8, 1d0, 1d0*it, 3d0, 1d0*sum(np2), sum(uf), particle_checksum()
  end subroutine io__ptcl

  subroutine io__orb(a, b, c, it)
    real(8), intent(in) :: a(*), b(*)
    integer, intent(in) :: c(*), it
    write(hunit,*) {MAGIC:.3e}_8, 2d0, 1d0*it, 3d0, 1d0*sum(np2), sum(uf), particle_checksum()
  end subroutine io__orb

  subroutine io__mom(a, b, it)          ! the interior nodes: what bc__mom leaves meaningful
    real(8), intent(in) :: a(*), b(*)
    integer, intent(in) :: it
    write(hunit,*) {MAGIC:.3e}_8, 3d0, 1d0*it, 2d0, sum(mom(@MOMIN)), sum(uf)
  end subroutine io__mom

  subroutine energy_history(a, b, c, it)
    real(8), intent(in) :: a(*), b(*)
    integer, intent(in) :: c(*), it
    write(hunit,*) {MAGIC:.3e}_8, 4d0, 1d0*it, 3d0, 1d0*sum(np2), sum(uf), particle_checksum()
  end subroutine energy_history

  subroutine save_restart(a, b, c, n1, n2, it, fname)
    real(8), intent(in) :: a(*), b(*)
    integer, intent(in) :: c(*), n1, n2, it, fname
    write(hunit,*) {MAGIC:.3e}_8, 5d0, 1d0*it, 3d0, 1d0*sum(np2), sum(uf), particle_checksum()
  end subroutine save_restart

  subroutine finalize()
  end subroutine finalize

  ! stand-ins for the shock driver's host-side particle source: they EDIT the host arrays (uz of every particle shrinks by 0.1 %),
  ! which is all that matters to the shim -- the state must have been brought back before, and must travel again afterwards
  subroutine inject()
    integer :: isp, {jk}, ii
    write(hunit,*) {MAGIC:.3e}_8, 6d0, 0d0, 2d0, 1d0*sum(np2), particle_checksum()
{edit}  end subroutine inject

  subroutine relocate()
    integer :: isp, {jk}, ii
    write(hunit,*) {MAGIC:.3e}_8, 7d0, 0d0, 2d0, 1d0*sum(np2), particle_checksum()
{edit}  end subroutine relocate
""".replace("@MOMIN", momin)


def assemble(setup, dim):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_reference_patch as mp
    from app_harness import procedure
    rel = f"{dim}d/proj/{setup}/app.f90"
    patched = mp.edit_app(open(os.path.join(REF, rel)).read(), dim, resident=True)
    lines = patched.splitlines()
    # the driver's own `use boundary_..., bc__init => ..., &` statement (continuation lines included)
    i = next(k for k, l in enumerate(lines) if re.match(r"\s*use boundary_", l))
    j = i
    while lines[j].rstrip().endswith("&"):
        j += 1
    use_bc = "\n".join(lines[i:j + 1])
    main = procedure(patched, "app__main").splitlines()
    a = next(k for k, l in enumerate(main) if re.match(r"\s*call init\(\)", l)) + 1
    body = "\n".join(main[a:-1])
    text = "\n".join([
        f"! ASSEMBLED by oracle/f2cxx/mainloop_harness.py from {rel} (patched by tools/make_reference_patch.py --resident): declarations and",
        "! output procedures by the harness, `harness__main` = the driver's app__main after `call init()`, verbatim",
        "module app", "  use particle", "  use field", "  use sort", "  use mom_calc", "  use wuming_b200_c", use_bc, "  implicit none",
        _decl(dim, setup == "shock"), "contains", _standins(dim),
        "  subroutine harness__main()", "    integer :: it", "    real(8) :: etime, etime0", body, "  end subroutine harness__main",
        "end module app"]) + "\n"
    return text


def lib_path(setup, dim):
    return os.path.join(OUT, f"libwuming_main_{setup}{dim}d.so")


def build(setup, dim, force=False):
    """-> path of the library, or None when /root/reference is absent and no prebuilt library exists"""
    lib = lib_path(setup, dim)
    ref_file = os.path.join(REF, f"{dim}d", "proj", setup, "app.f90")
    if not os.path.exists(ref_file):
        return lib if os.path.exists(lib) else None
    os.makedirs(OUT, exist_ok=True)
    shim = [os.path.join(ROOT, "fortran", "wuming_b200_c.f90"), os.path.join(ROOT, "fortran", f"wuming_b200_shim{dim}d.f90")]
    deps = shim + [ref_file, os.path.join(ROOT, "tools", "make_reference_patch.py")] + \
        [os.path.join(HERE, f) for f in ("f2cxx.py", "f90rt.h", "f90rt.cpp", "shim_rt.cpp", "mainloop_harness.py", "app_harness.py")]
    h = hashlib.sha1()
    for f in deps:
        h.update(open(f, "rb").read())
    stamp_file = lib + ".stamp"
    if not force and os.path.exists(lib) and os.path.exists(stamp_file) and open(stamp_file).read().strip() == h.hexdigest():
        return lib
    sys.path.insert(0, HERE)
    import f2cxx
    src = assemble(setup, dim)
    if os.environ.get("WM_KEEP_F90"):              # debugging aid only: the assembled text holds the reference's loop verbatim
        with open(os.path.join(OUT, f"main_{setup}{dim}d.f90"), "w") as f:
            f.write(src)
    with open(os.path.join(OUT, f"main_{setup}{dim}d.meta"), "w") as f:      # what the tests need to know about the driver's structure
        f.write(f"final_save_restart {int(src.count('call save_restart(') >= 2)}\n")
    files = [(os.path.relpath(f, ROOT), open(f).read()) for f in shim] + [(f"main_{setup}{dim}d.f90 <- {dim}d/proj/{setup}/app.f90", src)]
    cpp = os.path.join(OUT, f"main_{setup}{dim}d.cpp")
    with open(cpp, "w") as f:
        f.write(f2cxx.translate(files, skip=("wm_check",)))
    r = subprocess.run([CXX, "-std=c++17", "-O1", "-fPIC", "-shared", "-DF90_BOUNDS", "-I", HERE, "-o", lib, cpp,
                        os.path.join(HERE, "f90rt.cpp"), os.path.join(HERE, "shim_rt.cpp")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed on the translated main loop:\n" + r.stderr[-4000:])
    with open(stamp_file, "w") as f:
        f.write(h.hexdigest())
    return lib


SETUPS = [("weibel", 3), ("weibel", 2), ("reconnection", 3), ("reconnection", 2), ("shock", 3), ("shock", 2)]

if __name__ == "__main__":
    for s_, d_ in SETUPS:
        print(build(s_, d_, force="--force" in sys.argv))


# ------------------------------------------------------------------------------------------------------------------------------
# the WHOLE driver behind load_config: init() and app__main, verbatim, on top of the shim (Weibel)
# ------------------------------------------------------------------------------------------------------------------------------
def _standin_from_call(stmt, name, real_names=(), array_real=("up", "gp", "uf", "q", "r"), array_int=("np2", "cumcnt")):
    """a do-nothing harness procedure with the dummy list of the driver's own `call name(...)` statement (kinds by actual name)"""
    m = re.search(rf"call\s+{name}\s*\((.*)\)", stmt, re.S)
    acts = [a.strip() for a in m.group(1).replace("&", " ").split(",") if a.strip()]
    dums, decl = [], []
    for k, a in enumerate(acts):
        d = f"d{k + 1}"
        dums.append(d)
        if a in array_real:
            decl.append(f"    real(8) :: {d}(*)")
        elif a in array_int:
            decl.append(f"    integer :: {d}(*)")
        elif a in real_names:
            decl.append(f"    real(8) :: {d}")
        else:
            decl.append(f"    integer :: {d}")
    return f"  subroutine {name}({', '.join(dums)})\n" + "\n".join(decl) + f"\n  end subroutine {name}\n"


def _statement(text, first):
    """the (continued) statement starting at the first line matching `first`"""
    lines = text.splitlines()
    i = next(k for k, l in enumerate(lines) if re.search(first, l))
    j = i
    while lines[j].rstrip().endswith("&") or (j + 1 < len(lines) and lines[j + 1].lstrip().startswith("&")):
        j += 1
    return "\n".join(lines[i:j + 1])


def assemble_full(dim, setup="weibel"):
    """module app of {dim}d/proj/<setup>: the driver's init(), its loaders and particle source (set_initial_condition,
    set_particle_ids, get_global_cumsum, energy_history / inject, relocate, vprofile) and app__main VERBATIM from the patched file;
    the declarations (app_harness's), the configuration entry, the I/O procedures and mpi_set__init by the harness.  harness__main =
    app__main from `call init()` on; harness__loop = the same without that line (init() is an entry point of its own)"""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sys.path.insert(0, HERE)
    import make_reference_patch as mp
    import app_harness as ah
    a = ah.APPS[f"{setup}{dim}d"]
    rel = a["file"]
    patched = mp.edit_app(open(os.path.join(REF, rel)).read(), dim, resident=True)
    lines = patched.splitlines()
    i = next(k for k, l in enumerate(lines) if re.match(r"\s*use boundary_", l))
    j = i
    while lines[j].rstrip().endswith("&"):
        j += 1
    use_bc = "\n".join(lines[i:j + 1])
    init = ah.procedure(patched, "init")
    reals = ("delx", "delt", "c", "wpe", "wpi", "wge", "wgi", "vti", "vte")
    standins = [_standin_from_call(_statement(init, r"call io__init"), "io__init", reals),
                _standin_from_call(_statement(init, r"call io__input"), "io__input", reals),
                _standin_from_call(_statement(init, r"call save_param"), "save_param", reals)]
    yz = "nys = nygs; nye = nyge" + ("; nzs = nzgs; nze = nzge; nrank_j = 0; nrank_k = 0" if dim == 3 else "")
    mpi_args = "a1,a2,a3,a4,a5,a6,a7" if dim == 3 else "a1,a2,a3"
    cfg = [c for c, _ in a["config"]]
    ints = [c for c, t in a["config"] if t == "i"]
    main = ah.procedure(patched, "app__main").splitlines()
    k0 = next(k for k, l in enumerate(main) if re.match(r"\s*call init\(\)", l))
    momin = ":,nxgs:nxge,nys:nye,nzs:nze,:" if dim == 3 else ":,nxgs:nxge,nys:nye,:"
    jk = "j, k" if dim == 3 else "j"
    psum = _pencil_loops(dim, "      do ii = 1, np2(@P)\n        s = s + up(1,ii,@P) + 3d0*up(ndim-1,ii,@P)\n      enddo")
    out = [f"! ASSEMBLED by oracle/f2cxx/mainloop_harness.py (assemble_full) from {rel}, patched by tools/make_reference_patch.py --resident:",
           "! init(), the loaders, energy_history and app__main are the driver's text; declarations, configuration and I/O by the harness",
           "module app", "  use particle", "  use field", "  use sort", "  use mom_calc", "  use wuming_b200_c", use_bc, "  implicit none",
           a["decl"],
           "  integer :: mnpi = 4, jup = 0, jdown = 0, kup = 0, kdown = 0, nup = 0, ndown = 0, nstat(6)",
           "  integer :: max_it, intvl_ptcl, intvl_orb, intvl_mom, intvl_expand, verbose",
           "  integer :: restart_file = 0, hunit = 10, datadir = 0, param = 0",
           "  logical :: restart = .false.", "  real(8) :: max_elapsed", "contains", "",
           f"  subroutine harness__configure({', '.join(c + '_in' for c in cfg)}, {', '.join(r + '_in' for r in a['rank'])})",
           f"    integer, intent(in) :: {', '.join(c + '_in' for c in ints + a['rank'])}",
           f"    real(8), intent(in) :: {', '.join(c + '_in' for c, t in a['config'] if t == 'r')}"]
    out += [f"    {c} = {c}_in" for c in cfg + a["rank"]]
    out.append(ah.block(open(os.path.join(REF, rel)).read(), a["sizes"][1], a["sizes"][2], inside=a["sizes"][0]))
    out += ["  end subroutine harness__configure", "",
            f"  subroutine mpi_set__init({mpi_args})          ! one rank: the whole box, every neighbour is this rank",
            f"    integer, intent(in) :: {mpi_args}", f"    {yz}; nrank = 0", "  end subroutine mpi_set__init", "",
            "  subroutine init_random_seed()", "  end subroutine init_random_seed", ""] + standins + [f"""
  function get_etime() result(t)
    real(8) :: t
    t = 0d0
  end function get_etime

  function particle_checksum() result(s)
    real(8) :: s
    integer :: isp, {jk}, ii
    s = 0d0
{psum}  end function particle_checksum

  subroutine io__ptcl(a, b, c_, it)
    real(8), intent(in) :: a(*), b(*)
    integer, intent(in) :: c_(*), it
    write(hunit,*) {MAGIC:.3e}_8, 1d0, 1d0*it, 3d0, 1d0*sum(np2), sum(uf), particle_checksum()
  end subroutine io__ptcl

  subroutine io__orb(a, b, c_, it)
    real(8), intent(in) :: a(*), b(*)
    integer, intent(in) :: c_(*), it
    write(hunit,*) {MAGIC:.3e}_8, 2d0, 1d0*it, 3d0, 1d0*sum(np2), sum(uf), particle_checksum()
  end subroutine io__orb

  subroutine io__mom(a, b, it)
    real(8), intent(in) :: a(*), b(*)
    integer, intent(in) :: it
    write(hunit,*) {MAGIC:.3e}_8, 3d0, 1d0*it, 2d0, sum(mom({momin})), sum(uf)
  end subroutine io__mom

  subroutine save_restart(a, b, c_, n1, n2, it, fname)
    real(8), intent(in) :: a(*), b(*)
    integer, intent(in) :: c_(*), n1, n2, it, fname
    write(hunit,*) {MAGIC:.3e}_8, 5d0, 1d0*it, 3d0, 1d0*sum(np2), sum(uf), particle_checksum()
  end subroutine save_restart

  subroutine finalize()
  end subroutine finalize
""", init]
    for p_ in a["procs"]:
        out.append(ah.procedure(patched, p_))
    for nm, first in (("harness__main", k0), ("harness__loop", k0 + 1)):
        out += [f"  subroutine {nm}()", "    integer :: it", "    real(8) :: etime, etime0"] + main[first:-1] + [f"  end subroutine {nm}", ""]
    out.append("end module app")
    return "\n".join(out) + "\n"


def build_full(dim, setup="weibel", force=False):
    """-> path of oracle/_ref/libwuming_full_<setup>{dim}d.so (None without /root/reference and without a prebuilt library)"""
    lib = os.path.join(OUT, f"libwuming_full_{setup}{dim}d.so")
    ref_file = os.path.join(REF, f"{dim}d", "proj", setup, "app.f90")
    if not os.path.exists(ref_file):
        return lib if os.path.exists(lib) else None
    os.makedirs(OUT, exist_ok=True)
    shim = [os.path.join(ROOT, "fortran", "wuming_b200_c.f90"), os.path.join(ROOT, "fortran", f"wuming_b200_shim{dim}d.f90")]
    deps = shim + [ref_file, os.path.join(ROOT, "tools", "make_reference_patch.py")] + \
        [os.path.join(HERE, f) for f in ("f2cxx.py", "f90rt.h", "f90rt.cpp", "shim_rt.cpp", "mainloop_harness.py", "app_harness.py")]
    h = hashlib.sha1(b"full" + setup.encode())
    for f in deps:
        h.update(open(f, "rb").read())
    stamp_file = lib + ".stamp"
    if not force and os.path.exists(lib) and os.path.exists(stamp_file) and open(stamp_file).read().strip() == h.hexdigest():
        return lib
    sys.path.insert(0, HERE)
    import f2cxx
    src = assemble_full(dim, setup)
    if os.environ.get("WM_KEEP_F90"):
        with open(os.path.join(OUT, f"full_{setup}{dim}d.f90"), "w") as f:
            f.write(src)
    files = [(os.path.relpath(f, ROOT), open(f).read()) for f in shim] + [(f"full_{setup}{dim}d.f90 <- {dim}d/proj/{setup}/app.f90", src)]
    cpp = os.path.join(OUT, f"full_{setup}{dim}d.cpp")
    with open(cpp, "w") as f:
        f.write(f2cxx.translate(files, skip=("wm_check",)))
    r = subprocess.run([CXX, "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-DF90_BOUNDS", "-I", HERE, "-o", lib, cpp,
                        os.path.join(HERE, "f90rt.cpp"), os.path.join(HERE, "shim_rt.cpp")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed on the translated driver:\n" + r.stderr[-4000:])
    with open(stamp_file, "w") as f:
        f.write(h.hexdigest())
    return lib
