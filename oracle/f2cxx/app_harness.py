"""app_harness.py -- the DRIVER-level procedures of the reference (`proj/*/app.f90`: initial loads, the shock driver's inject() /
relocate(), energy_history) through the same translator.  TEST INFRASTRUCTURE ONLY.

`app.f90` as a whole needs the JSON, MPI-IO and HDF5 modules and character handling, none of which is on the hot path.  What the
oracle and wumingpic_b200/setups.py restate are single procedures of it, so this recipe EXTRACTS those procedures -- and the few
statement blocks of `load_config` / `init` that derive sizes and physical constants -- from the reference file WHERE IT LIES, verbatim,
and wraps them into a module `app` whose declaration part is written here (the names and kinds of the driver's module variables those
procedures use; the reference's own declarations are interleaved with `character` / JSON declarations the translator does not take).
Everything executable is the reference's text; f2cxx.py translates the assembled module like any other file.

The random-number consumers of utils/wuming_utils.f90 (`uniform_rand`, `normal_rand`, `shuffle`) are INPUTS of a comparison, not
translated: Fortran's `random_number` stream is not reproducible (the reference seeds it from the clock), so the test driver hands
out the values in call order (f90rt.h hooks) -- the very values the oracle's keyed Philox streams give the same particles.

    build(name) -> path of oracle/_ref/libwuming_app_<name>.so   (None without /root/reference and without a prebuilt library)
"""
import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref")
REF = os.environ.get("WUMING_REFERENCE", "/root/reference")
CXX = "/usr/bin/g++"


def procedure(text, name):
    """the text of `subroutine name` / `function name` ... `end subroutine|function name`, verbatim"""
    lines = text.splitlines()
    start = None
    for i, ln in enumerate(lines):
        if start is None and re.match(rf"\s*(subroutine|function)\s+{name}\b", ln, re.I):
            start = i
        elif start is not None and re.match(rf"\s*end\s*(subroutine|function)\s+{name}\b", ln, re.I):
            return "\n".join(lines[start:i + 1]) + "\n"
    raise KeyError(f"procedure {name} not found")


def block(text, first, last, inside=None):
    """the lines from the first one matching `first` to the first later one matching `last` (both included), verbatim;
    `inside`: only search within that procedure"""
    if inside:
        text = procedure(text, inside)
    lines = text.splitlines()
    for i, ln in enumerate(lines):
        if re.search(first, ln):
            for j in range(i, len(lines)):
                if re.search(last, lines[j]):
                    return "\n".join(lines[i:j + 1]) + "\n"
            break
    raise KeyError(f"block {first!r} .. {last!r} not found")


# ---- per-application assembly --------------------------------------------------------------------------------------------------
# decl: the declaration part of the wrapper module (written here).  config: the "parameter" section of the driver's config.json =
# the dummy list of harness__configure.  Everything executable comes from the reference through procedure() / block().
_MPI = "  integer :: mnpr = 8, opsum = 1, ncomw = 0, nerr = 0      ! what `use mpi_set` provides; datatype handle = element size\n"


def _common(dim, cfl):
    yz = "nys, nye" if dim == 2 else "nys, nye, nzs, nze, nrank_j, nrank_k"
    sizes = "ny, nygs, nyge" if dim == 2 else "ny, nygs, nyge, nz, nzgs, nzge, nproc_j, nproc_k"
    arr = ("np2(:,:), cumcnt(:,:,:)", "uf(:,:,:), up(:,:,:,:), gp(:,:,:,:), mom(:,:,:,:)") if dim == 2 else \
        ("np2(:,:,:), cumcnt(:,:,:,:)", "uf(:,:,:,:), up(:,:,:,:,:), gp(:,:,:,:,:), mom(:,:,:,:,:)")
    decl = f"""
  integer :: nproc, nrank, it0, np, n0, nx, nxgs, nxge, nxs, nxe, {sizes}, {yz}, mpierr
  integer, parameter :: ndim = {6 if dim == 2 else 7}, nsp = 2, nroot = 0
  real(8), parameter :: c = 1.0d0, gfac = 0.501d0, cfl = {cfl}, delx = 1.0d0, pi = 4.0d0*atan(1.0d0)
  integer, allocatable :: {arr[0]}
  real(8), allocatable :: {arr[1]}
""" + _MPI
    return decl, ["nrank"] + yz.split(", ")


def _shock(dim):
    decl, rank = _common(dim, "1.0d0")
    cfgi = ["num_process", "n_ppc", "n_x", "n_x_ini", "n_y"] if dim == 2 else ["num_process", "num_process_j", "n_ppc", "n_x", "n_x_ini", "n_y", "n_z"]
    cfgr = ["u_inject", "mass_ratio", "sigma_e", "omega_pe", "v_the", "v_thi", "theta_bn", "phi_bn", "l_damp_ini"]
    decl += f"""  integer :: {', '.join(cfgi)}
  real(8) :: {', '.join(cfgr)}
  real(8), parameter :: xrs = 450.0d0, xre = 500.0d0
  real(8) :: r(nsp), q(nsp), delt, b0, u0, v0, gam0
"""
    return dict(file=f"{dim}d/proj/shock/app.f90", decl=decl, config=[(c, "i") for c in cfgi] + [(c, "r") for c in cfgr], rank=rank,
                sizes=("load_config", r"^\s*nproc\s*=\s*num_process", r"phi_bn\s*=\s*phi_bn"),
                init_locals="integer :: isp, i, j" + (", k" if dim == 3 else "") + "\n    real(8) :: wpe, wpi, wge, wgi, vte, vti",
                init_blocks=[("init", r"allocate\(np2", r"^\s*mom\b.*=\s*0"), ("init", r"^\s*delt\s*=\s*cfl", r"^\s*b0\s*="),
                             ("init", r"! number of particles", r"! initialize modules")],
                init_tail=["call set_initial_condition()", "it0 = 0", "gp = up"],
                procs=["set_initial_condition", "set_particle_ids", "relocate", "inject", "get_global_cumsum", "vprofile"])


def _weibel(dim):
    decl, rank = _common(dim, "1.0d0")
    cfgi = ["num_process", "n_ppc", "n_x", "n_y"] if dim == 2 else ["num_process", "num_process_j", "n_ppc", "n_x", "n_y", "n_z"]
    cfgr = ["mass_ratio", "sigma_e", "omega_pe", "v_the", "v_thi", "t_ani"]
    decl += f"""  integer :: {', '.join(cfgi)}
  real(8) :: {', '.join(cfgr)}
  real(8) :: r(nsp), q(nsp), delt, b0
"""
    return dict(file=f"{dim}d/proj/weibel/app.f90", decl=decl, config=[(c, "i") for c in cfgi] + [(c, "r") for c in cfgr], rank=rank,
                sizes=("load_config", r"^\s*nproc\s*=\s*num_process", r"^\s*nxe\s*=\s*nxge"),
                init_locals="integer :: isp, i, j" + (", k" if dim == 3 else "") + "\n    real(8) :: wpe, wpi, wge, wgi, vte, vti",
                init_blocks=[("init", r"allocate\(np2", r"^\s*mom\b.*=\s*0"), ("init", r"^\s*delt\s*=\s*cfl", r"^\s*b0\s*="),
                             ("init", r"! number of particles", r"! initialize modules")],
                init_tail=["call set_initial_condition()", "it0 = 0", "gp = up"],
                procs=["set_initial_condition", "set_particle_ids", "get_global_cumsum", "energy_history"])


def _reconnection(dim):
    decl, rank = _common(dim, "0.5d0")
    cfgi = ["num_process", "n_x", "n_y"] if dim == 2 else ["num_process", "num_process_j", "n_x", "n_y", "n_z"]
    cfgr = ["mass_ratio", "alpha", "rtemp", "lcs"]
    decl += f"""  integer :: {', '.join(cfgi)}, nbg, ncs
  real(8) :: {', '.join(cfgr)}
  real(8) :: r(nsp), q(nsp), delt, b0, vte, vti, x0, y0{', z0' if dim == 3 else ''}
"""
    return dict(file=f"{dim}d/proj/reconnection/app.f90", decl=decl,
                config=[(c, "i") for c in cfgi] + [(c, "r") for c in cfgr] + [("nbg", "i"), ("ncs", "i")], rank=rank,
                sizes=("load_config", r"^\s*nproc\s*=\s*num_process", r"^\s*nxe\s*=\s*nxge"),
                init_locals="integer :: isp, i, j, ii" + (", k" if dim == 3 else "") + "\n    real(8) :: wpe, wpi, wge, wgi, ldb",
                init_blocks=[("init", r"allocate\(np2", r"^\s*mom\b.*=\s*0"), ("init", r"^\s*r\(1\)\s*=\s*mass_ratio", r"^\s*np2\(.*=\s*nbg")],
                # the driver sorts the load with sort__bucket (another module: the test calls the translated one) before `up = gp`
                init_tail=["call set_initial_condition()", "it0 = 0"],
                procs=["set_initial_condition", "set_particle_ids", "get_global_cumsum", "energy_history"])


def _pack(dim):
    """get_particle_count of paraio (what io__ptcl / io__orb pack and count): the module's geometry variables are set by a harness
    procedure, the procedure itself is the reference's"""
    names = ["ndim", "np", "nsp", "nys", "nye"] + (["nzs", "nze"] if dim == 3 else []) + ["nproc"]
    decl = f"\n  integer :: {', '.join(names)}, mpierr\n"
    setup = (f"  subroutine harness__set({', '.join(n + '_in' for n in names)})\n"
             f"    integer, intent(in) :: {', '.join(n + '_in' for n in names)}\n"
             + "".join(f"    {n} = {n}_in\n" for n in names) + "  end subroutine harness__set\n")
    return dict(file=f"{dim}d/common/paraio.f90", decl=decl, raw=setup, procs=["get_particle_count"])


APPS = {"pack2d": _pack(2), "pack3d": _pack(3), "shock2d": _shock(2), "shock3d": _shock(3), "weibel2d": _weibel(2), "weibel3d": _weibel(3),
        "reconnection2d": _reconnection(2), "reconnection3d": _reconnection(3)}


def assemble(name):
    a = APPS[name]
    text = open(os.path.join(REF, a["file"])).read()
    if "raw" in a:
        return "\n".join([f"! ASSEMBLED by oracle/f2cxx/app_harness.py from {a['file']}", "module app", "  implicit none", a["decl"],
                          "contains", "", a["raw"]] + [procedure(text, p) for p in a["procs"]] + ["end module app"]) + "\n"
    cfg = [c for c, _ in a["config"]]
    out = [f"! ASSEMBLED by oracle/f2cxx/app_harness.py from {a['file']}: declarations by the harness, every executable statement the reference's",
           "module app", "  implicit none", a["decl"], "contains", "",
           f"  subroutine harness__configure({', '.join(c + '_in' for c in cfg)}, {', '.join(r + '_in' for r in a['rank'])})"]
    ints = [c for c, t in a["config"] if t == "i"]
    out.append(f"    integer, intent(in) :: {', '.join(c + '_in' for c in ints + a['rank'])}")
    out.append(f"    real(8), intent(in) :: {', '.join(c + '_in' for c, t in a['config'] if t == 'r')}")
    out += [f"    {c} = {c}_in" for c in cfg + a["rank"]]
    out.append(block(text, a["sizes"][1], a["sizes"][2], inside=a["sizes"][0]))
    out += ["  end subroutine harness__configure", "", "  subroutine harness__init()", "    " + a["init_locals"]]
    for pr, first, last in a["init_blocks"]:
        out.append(block(text, first, last, inside=pr))
    out += ["    " + t for t in a["init_tail"]] + ["  end subroutine harness__init", ""]
    for p in a["procs"]:
        out.append(procedure(text, p))
    out.append("end module app")
    return "\n".join(out) + "\n"


def lib_path(name):
    return os.path.join(OUT, f"libwuming_app_{name}.so")


def reference_present(name):
    return os.path.exists(os.path.join(REF, APPS[name]["file"]))


def build(name, force=False):
    lib = lib_path(name)
    if not reference_present(name):
        return lib if os.path.exists(lib) else None
    os.makedirs(OUT, exist_ok=True)
    h = hashlib.sha1(open(os.path.join(REF, APPS[name]["file"]), "rb").read())
    for f in ("f2cxx.py", "f90rt.h", "f90rt.cpp", "app_harness.py"):
        h.update(open(os.path.join(HERE, f), "rb").read())
    stamp, stamp_file = h.hexdigest(), lib + ".stamp"
    if not force and os.path.exists(lib) and os.path.exists(stamp_file) and open(stamp_file).read().strip() == stamp:
        return lib
    sys.path.insert(0, HERE)
    import f2cxx
    src = assemble(name)
    if os.environ.get("WM_KEEP_F90"):              # debugging aid only: the assembled text holds the reference's procedures verbatim,
        with open(os.path.join(OUT, f"app_{name}.f90"), "w") as f:       # so it is not kept (only the generated C++ and the library are)
            f.write(src)
    cpp = os.path.join(OUT, f"app_{name}.cpp")
    with open(cpp, "w") as f:
        f.write(f2cxx.translate([(f"app_{name}.f90 <- {APPS[name]['file']}", src)]))
    r = subprocess.run([CXX, "-std=c++17", "-O2", "-ffp-contract=off", "-frounding-math", "-fPIC", "-shared", "-DF90_BOUNDS", "-I", HERE,
                        "-o", lib, cpp, os.path.join(HERE, "f90rt.cpp")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed on the translated driver procedures:\n" + r.stderr[-4000:])
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return lib


if __name__ == "__main__":
    for n in APPS:
        print(n, build(n, force="--force" in sys.argv))
